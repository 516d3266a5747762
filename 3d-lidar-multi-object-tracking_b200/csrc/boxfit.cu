// boxfit.cu -- per-cluster L-shape / minimum-area-rectangle box fitting on sm_100a.
//
// Replaces boxFitting (/root/reference/object_tracking/src/cluster/box_fitting.cpp:422-431):
//   getClusteredPoints (:46-72)   gather the elevated points of every cluster IN CLOUD ORDER
//   getBoundingBox (:212-418)     per cluster: pixelise @18 px/m around point #0, first-occurrence min/max
//                                 slope points, max z (:239-293); L-shape fit from 80 seeded random samples
//                                 (:296-356) or cv::minAreaRect (:358-365); ruleBasedFilter (:97-158);
//                                 8-corner box (:379-389); mark_cluster centroid + AABB (:161-209)
//
// The reference's results depend on the order of the points inside a cluster (point #0, first-occurrence
// ties, the random index picks the i-th point), so the gather is a STABLE counting sort by cluster id:
//   B1 tile_hist_kernel    per 1024-point tile: cluster id of each point (label grid lookup through the u16 cell
//                          id kept by clustering) + shared-memory histogram -> table[tile][cluster]
//   B2 seg_offsets_body    exclusive scan over tiles per cluster, then over clusters -> segment starts (run by B1's last CTA)
//   B3 scatter_kernel      stable in-tile ranks (warp match_any, warps in order) -> sorted point indices
//   B4 box_fit_kernel      one CTA per cluster (grid-stride), all of getBoundingBox; the last CTA to finish
//                          compacts the accepted boxes in cluster-id order (:355,365 skip rejected clusters).
//
// cv::minAreaRect is third-party arithmetic that is not in /root/reference: it is implemented as the exact
// integer contract documented in oracle/mar_contract.cpp / DESIGN.md (parity for that one call is UNPINNED
// against OpenCV itself, bit-exact against the contract).
#include <cfloat>
#include <climits>
#include <random>
#include "lmot_internal.cuh"

namespace lmot {

namespace {

constexpr int kTile = 1024;
constexpr int kFitThreadsFrame = 512;   // one CTA per cluster: the largest cluster IS the kernel of a single frame (16 us of point pass at 256 threads)
constexpr int kFitThreadsBatch = 256;   // batched ticks hold ~8x the clusters: more, smaller CTAs per SM finish the union sooner (51 us at 512, 45 at 256)
constexpr int kCols = 1800;        // pixel columns a cluster can span: offsetX-450 = picX-initPicX in [-899,899]
constexpr int kColShift = 899 - 450;
constexpr int kHullCap = 2048;

// Programmatic dependent launch inside a slot's detection stream: scatter behind tile_hist, box_fit behind scatter.  The dependent's
// CTAs become resident once EVERY CTA of the kernel before it has started (they all trigger at their first instruction) and wait there
// for it to finish -- its launch latency leaves the frame's critical path.  Deadlock-free next to the ground kernel's own barrier: a
// waiting CTA only ever waits for a grid that is completely resident and waits for nobody.  (Not used in front of tile_hist: its CTAs
// would hold SM slots for the ~12 us of the ground kernel, slots the tracker chain needs.)
__device__ __forceinline__ void fit_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void fit_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------- B1
// One frame of a launch (blockIdx.y): every buffer the four kernels touch for it (the Slot's, see lmot_internal.cuh)
struct FitFrame {
  const uint16_t* cart; const int* label_grid; int* counters;
  uint16_t* pcid; int* table; int* seg_start; int* seg_size;
  const float4* elev; float4* sorted_pts;
  float* cl_box; float* cl_marker; uint8_t* cl_ok;
  float* boxes; float* markers; int* done; int* det_sem;
  unsigned* hist_ctr;      // atomicInc counter of finished tile_hist CTAs (wraps at the grid size: needs no reset)
};
struct FitBatch { FitFrame f[kMaxBatch]; };

__device__ __forceinline__ void seg_offsets_body(const FitFrame& F, int max_clusters);

__global__ void __launch_bounds__(kTile)
tile_hist_kernel(const __grid_constant__ FitBatch B, int max_clusters) {
  extern __shared__ int s_hist[];
  fit_pdl_trigger();
  const FitFrame& F = B.f[blockIdx.y];
  const uint16_t* __restrict__ cart = F.cart; const int* __restrict__ label_grid = F.label_grid; const int* __restrict__ counters = F.counters;
  uint16_t* __restrict__ pcid = F.pcid; int* __restrict__ table = F.table;
  const int n = counters[CNT_N_ELEV];
  const int K = min(counters[CNT_NUM_CLUSTER], max_clusters);
  const int tile = blockIdx.x;
  if (tile * kTile < n) {
  for (int k = threadIdx.x; k <= K; k += kTile) s_hist[k] = 0;
  __syncthreads();
  const int i = tile * kTile + threadIdx.x;
  unsigned cid = 0;
  if (i < n) {
    const unsigned c = cart[i];
    if (c != kNoCell) cid = (unsigned)__ldg(&label_grid[c]);
    if (cid > (unsigned)K) cid = 0;                    // capacity overflow is reported by seg_offsets_kernel
    pcid[i] = (uint16_t)cid;
  }
  const unsigned grp = __match_any_sync(0xFFFFFFFFu, cid);
  if (cid != 0 && (threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(&s_hist[cid], __popc(grp));
  __syncthreads();
  int* row = table + (size_t)tile * (max_clusters + 1);
  for (int k = threadIdx.x; k <= K; k += kTile) row[k] = s_hist[k];
  }
  // B2 by the last CTA of the frame to get here (tiles past the end of the cloud count too)
  __shared__ int s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicInc(F.hist_ctr, gridDim.x - 1u) == gridDim.x - 1u) ? 1 : 0;
  __syncthreads();
  if (s_last) { __threadfence(); seg_offsets_body(F, max_clusters); }
}

// ---------------------------------------------------------------------------------------------- B2
// exclusive scan over tiles per cluster, then over clusters -> segment starts.  One CTA of 1,024 threads per frame: the LAST
// tile_hist CTA of the frame to finish runs it (no launch in between); seg_offsets_kernel wraps it for frames without points.
__device__ __forceinline__ void seg_offsets_body(const FitFrame& F, int max_clusters) {
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  int* __restrict__ table = F.table; int* counters = F.counters;
  int* __restrict__ seg_start = F.seg_start; int* __restrict__ seg_size = F.seg_size; int* __restrict__ done = F.done;
  const int n = counters[CNT_N_ELEV];
  int K = counters[CNT_NUM_CLUSTER];
  if (threadIdx.x == 0) {
    s_carry = 0;
    *done = 0;
    if (K > max_clusters) counters[CNT_ERROR] = LMOT_ERR_CAPACITY;
  }
  K = min(K, max_clusters);
  const int n_tiles = (n + kTile - 1) / kTile;
  const int stride = max_clusters + 1;
  __syncthreads();
  // P lanes per cluster (a power of two, aligned inside a warp): each takes a contiguous range of tiles, so that a frame of ~60 tiles
  // and ~100 clusters is ONE batch of independent loads per lane instead of four dependent batches on 100 of the 1,024 threads
  const int P = K <= 128 ? 8 : (K <= 256 ? 4 : (K <= 512 ? 2 : 1));
  const int per_round = 1024 / P;
  const int per = (n_tiles + P - 1) / P;                // tiles per lane
  for (int k0 = 1; k0 <= K; k0 += per_round) {
    const int k = k0 + (int)threadIdx.x / P, j = (int)threadIdx.x % P;
    const int t_beg = min(j * per, n_tiles), t_end = min(t_beg + per, n_tiles);
    const bool mine = k <= K;
    int v[16];
    int sum = 0;
    if (mine) {
      if (per <= 16) {
#pragma unroll
        for (int u = 0; u < 16; ++u) { v[u] = (t_beg + u < t_end) ? table[(size_t)(t_beg + u) * stride + k] : 0; }
#pragma unroll
        for (int u = 0; u < 16; ++u) sum += v[u];
      } else {
        for (int t0 = t_beg; t0 < t_end; t0 += 16) {
#pragma unroll
          for (int u = 0; u < 16; ++u) v[u] = (t0 + u < t_end) ? table[(size_t)(t0 + u) * stride + k] : 0;
#pragma unroll
          for (int u = 0; u < 16; ++u) sum += v[u];
        }
      }
    }
    // exclusive prefix of the lanes' sums inside the cluster's group of P lanes; the group total ends up on every lane
    int incl_l = sum;
    for (int o = 1; o < P; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl_l, o, P); if (j >= o) incl_l += t; }
    const int group_total = __shfl_sync(0xFFFFFFFFu, incl_l, P - 1, P);
    if (mine) {
      int run = incl_l - sum;
      if (per <= 16) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
          if (t_beg + u < t_end) { table[(size_t)(t_beg + u) * stride + k] = run; run += v[u]; }
      } else {
        for (int t0 = t_beg; t0 < t_end; t0 += 16) {       // 16 independent loads in flight, then the dependent prefix
#pragma unroll
          for (int u = 0; u < 16; ++u) v[u] = (t0 + u < t_end) ? table[(size_t)(t0 + u) * stride + k] : 0;
#pragma unroll
          for (int u = 0; u < 16; ++u)
            if (t0 + u < t_end) { table[(size_t)(t0 + u) * stride + k] = run; run += v[u]; }
        }
      }
    }
    const int total = (mine && j == 0) ? group_total : 0;   // one lane per cluster carries its size into the scan over clusters
    // block exclusive scan of `total` (+ carry from previous chunks of 1024 clusters)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int v = s_warp[lane];
      int wi = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, wi, o); if (lane >= o) wi += t; }
      s_warp[lane] = wi - v;
    }
    __syncthreads();
    const int excl = s_carry + s_warp[warp] + incl - total;
    if (mine && j == 0) { seg_start[k] = excl; seg_size[k] = total; }
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = excl + total;
    __syncthreads();
  }
  if (threadIdx.x == 0) counters[CNT_N_CLUSTERED] = s_carry;
}

__global__ void __launch_bounds__(1024)
seg_offsets_kernel(const __grid_constant__ FitBatch B, int max_clusters) { seg_offsets_body(B.f[blockIdx.x], max_clusters); }

// ---------------------------------------------------------------------------------------------- B3
__global__ void __launch_bounds__(kTile)
scatter_kernel(const __grid_constant__ FitBatch B, int max_clusters) {
  extern __shared__ int s_cur[];
  fit_pdl_trigger();
  fit_pdl_wait();                          // tile_hist_kernel (+ the segment offsets of its last CTA) is complete
  const FitFrame& F = B.f[blockIdx.y];
  const uint16_t* __restrict__ pcid = F.pcid; const float4* __restrict__ elev = F.elev; const int* __restrict__ counters = F.counters;
  const int* __restrict__ table = F.table; const int* __restrict__ seg_start = F.seg_start; float4* __restrict__ sorted_pts = F.sorted_pts;
  const int n = counters[CNT_N_ELEV];
  const int K = min(counters[CNT_NUM_CLUSTER], max_clusters);
  const int tile = blockIdx.x;
  if (tile * kTile >= n) return;
  const int* row = table + (size_t)tile * (max_clusters + 1);
  for (int k = threadIdx.x + 1; k <= K; k += kTile) s_cur[k] = seg_start[k] + row[k];
  __syncthreads();
  const int i = tile * kTile + threadIdx.x;
  const unsigned cid = (i < n) ? pcid[i] : 0u;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (cid != 0) q = __ldg(&elev[i]);            // the point itself travels: box fitting then reads its cluster contiguously
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned grp = __match_any_sync(0xFFFFFFFFu, cid);
  const int leader = __ffs(grp) - 1;
  const int rank = __popc(grp & ((1u << lane) - 1u));
  // Ranks follow the cloud order: the warps' groups must take their segment ranges in warp order.  Every warp lists its groups
  // (cluster id, size); warp 0 then walks the 32 lists in order, one group per lane -- a warp's groups have distinct ids, so only
  // consecutive lists depend on each other.  (The warps taking turns behind 32 CTA barriers was ~5 of this kernel's ~8 us.)
  __shared__ unsigned short s_gcid[kTile / 32][32];
  __shared__ unsigned char s_gcnt[kTile / 32][32], s_gn[kTile / 32];
  __shared__ int s_gbase[kTile / 32][32];
  const bool is_leader = cid != 0 && lane == leader;
  const unsigned lmask = __ballot_sync(0xFFFFFFFFu, is_leader);
  if (is_leader) {
    const int slot = __popc(lmask & ((1u << lane) - 1u));
    s_gcid[warp][slot] = (unsigned short)cid; s_gcnt[warp][slot] = (unsigned char)__popc(grp);
  }
  if (lane == 0) s_gn[warp] = (unsigned char)__popc(lmask);
  __syncthreads();
  if (warp == 0) {
    for (int w = 0; w < kTile / 32; ++w) {
      if (lane < (int)s_gn[w]) {
        const int c = s_gcid[w][lane];
        const int base = s_cur[c];
        s_cur[c] = base + (int)s_gcnt[w][lane];
        s_gbase[w][lane] = base;
      }
      __syncwarp();
    }
  }
  __syncthreads();
  if (cid != 0) sorted_pts[s_gbase[warp][__popc(lmask & ((1u << leader) - 1u))] + rank] = q;
}

// ---------------------------------------------------------------------------------------------- B4
struct BoxParams {
  float roi, pic_scale, sensor_height;
  int ram_points, l_slope_dist, l_num_points, min_points, rule_mode;
  float t_height_min, t_height_max, t_width_min, t_width_max, t_len_min, t_len_max, t_area_max, t_ratio_min, t_ratio_max,
      min_len_ratio, t_pt_per_m3;
};

__device__ __forceinline__ unsigned okey(float f) {   // order-preserving, -0 == +0
  if (f == 0.f) f = 0.f;
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T* s_buf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xFFFFFFFFu, v, o));
  __syncthreads();
  if (lane == 0) s_buf[warp] = v;
  __syncthreads();
  T r = s_buf[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) r = op(r, s_buf[w]);
  return r;
}

struct MinU64 { __device__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a < b ? a : b; } };
struct MaxU64 { __device__ unsigned long long operator()(unsigned long long a, unsigned long long b) const { return a > b ? a : b; } };
struct MaxF { __device__ float operator()(float a, float b) const { return fmaxf(a, b); } };
struct MinF { __device__ float operator()(float a, float b) const { return fminf(a, b); } };
struct AddD { __device__ double operator()(double a, double b) const { return a + b; } };
struct MinI { __device__ int operator()(int a, int b) const { return a < b ? a : b; } };
struct MaxI { __device__ int operator()(int a, int b) const { return a > b ? a : b; } };

// ruleBasedFilter, box_fitting.cpp:97-158 (see lmot_rule_filter in lmot.h for the two modes)
__device__ bool rule_filter(const float pc[4][2], float maxZ, int n, const BoxParams& P) {
  if (n < P.min_points) return false;
  const float x1 = pc[0][0], y1 = pc[0][1], x2 = pc[1][0], y2 = pc[1][1], x3 = pc[2][0], y3 = pc[2][1];
  const float dist1 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
  const float dist2 = sqrtf((x3 - x2) * (x3 - x2) + (y3 - y2) * (y3 - y2));
  float length, width;
  if (dist1 > dist2) { length = dist1; width = dist2; } else { length = dist2; width = dist1; }
  const float height = maxZ + P.sensor_height;
  const float area = dist1 * dist2;
  const float mass = area * height;
  const float ratio = length / width;
  if (height > P.t_height_min && height < P.t_height_max) {
    if (P.rule_mode == LMOT_RULE_GCC13_O2_COMPAT) return true;
    if (width > P.t_width_min && width < P.t_width_max)
      if (length > P.t_len_min && length < P.t_len_max)
        if (area < P.t_area_max)
          if ((float)n > mass * P.t_pt_per_m3) {
            if (length > P.min_len_ratio) { if (ratio > P.t_ratio_min && ratio < P.t_ratio_max) return true; }
            else return true;
          }
    return false;   // INTENDED: the reference falls off the end here (undefined behaviour)
  }
  return false;
}

__device__ __forceinline__ void fit_mark(unsigned long long* clk, int slot) {
  if (clk && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); const unsigned row = blockIdx.y * gridDim.x + blockIdx.x; if (row < (unsigned)kFitClockCtas) clk[row * 8 + slot] = t; }
}

template <int NT>
__global__ void __launch_bounds__(NT, 1024 / NT)     // <= 64 registers: 1,024 threads of box fitting per SM
box_fit_kernel(const __grid_constant__ FitBatch B, const __grid_constant__ BoxParams P,
               const unsigned long long* __restrict__ mt_raw, int n_raw, int max_clusters, int max_boxes,
               unsigned long long* __restrict__ clk) {
  constexpr int kFitThreads = NT;
  fit_pdl_trigger();
  fit_pdl_wait();                          // scatter_kernel is complete
  const FitFrame& F = B.f[blockIdx.y];
  const float4* __restrict__ sorted_pts = F.sorted_pts; const int* __restrict__ seg_start = F.seg_start; const int* __restrict__ seg_size = F.seg_size;
  int* __restrict__ counters = F.counters;
  float* __restrict__ cl_box = F.cl_box; float* __restrict__ cl_marker = F.cl_marker; uint8_t* __restrict__ cl_ok = F.cl_ok;
  float* __restrict__ boxes = F.boxes; float* __restrict__ markers = F.markers; int* __restrict__ done = F.done; int* __restrict__ det_sem = F.det_sem;
  __shared__ int s_lo[kCols], s_hi[kCols];
  __shared__ short s_hx[kHullCap], s_hy[kHullCap];
  __shared__ short s_cx[kCols], s_clo[kCols], s_chi[kCols];     // occupied pixel columns, compacted
  __shared__ uint8_t s_flag8[kCols];
  __shared__ unsigned long long s_red64[kFitThreads / 32];
  __shared__ unsigned long long s_part64[2][kFitThreads / 32];
  __shared__ double s_partd[3][kFitThreads / 32];
  __shared__ float s_partf[7][kFitThreads / 32];
  __shared__ int s_redi[kFitThreads / 32];
  __shared__ int s_m;                        // hull size
  __shared__ int s_best;                     // best edge
  __shared__ unsigned long long s_bnum[kFitThreads / 32], s_bl[kFitThreads / 32];
  __shared__ int s_bidx[kFitThreads / 32];
  __shared__ float s_pc[4][2];
  __shared__ int s_flag;
  __shared__ int s_src[kFitThreads];         // compaction: cluster id of the j-th accepted box of the current batch
  const int tid = threadIdx.x;
  const int K = min(counters[CNT_NUM_CLUSTER], max_clusters);
  const float half = P.roi / 2;
  const float pic = P.pic_scale * P.roi;     // 900

  fit_mark(clk, 0);
  for (int k = blockIdx.x + 1; k <= K; k += gridDim.x) {
    const int n = seg_size[k];
    const float4* seg = sorted_pts + seg_start[k];          // the cluster's points, cloud order
    if (tid == 0) cl_ok[k] = 0;
    if (n < P.min_points || n <= 0) continue;      // ruleBasedFilter's first test (:100) rejects it whatever the fit

    // ---- point #0 and the pixel offsets (:218-225)
    const float4 q0 = __ldg(&seg[0]);
    const int initX = (int)floorf((q0.x + half) * P.pic_scale);
    const int initY = (int)floorf((q0.y + half) * P.pic_scale);
    const int initPicX = initX;
    const int initPicY = (int)(pic - (float)initY);
    const int offsetInitX = (int)(P.roi * P.pic_scale / 2 - (float)initPicX);
    const int offsetInitY = (int)(P.roi * P.pic_scale / 2 - (float)initPicY);

    // ---- per-point pass (:239-293): column extremes for the hull, slope extremes, max z, centroid, AABB
    for (int c = tid; c < kCols; c += kFitThreads) { s_lo[c] = INT_MAX; s_hi[c] = INT_MIN; }
    __syncthreads();
    unsigned long long kmin = ~0ull, kmax = 0ull;
    float maxZ = -99.f;
    double sx = 0, sy = 0, sz = 0;
    float mnx = FLT_MAX, mny = FLT_MAX, mnz = FLT_MAX, mxx = -FLT_MAX, mxy = -FLT_MAX, mxz = -FLT_MAX;
    // four points per thread in flight: as a plain loop this was one L2 round trip per iteration (18 us for a 5000-point cluster)
    for (int j0 = tid; j0 < n; j0 += 4 * kFitThreads) {
      float4 qq[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) { const int j = j0 + u * kFitThreads; qq[u] = (j < n) ? __ldg(&seg[j]) : make_float4(0.f, 0.f, 0.f, 0.f); }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * kFitThreads;
        if (j >= n) break;
        const float4 q = qq[u];
        const int x = (int)floorf((q.x + half) * P.pic_scale);
        const int y = (int)floorf((q.y + half) * P.pic_scale);
        const int picX = x, picY = (int)(pic - (float)y);
        const int offX = picX + offsetInitX, offY = picY + offsetInitY;
        const int col = offX + kColShift;
        if (col >= 0 && col < kCols) { atomicMin(&s_lo[col], offY); atomicMax(&s_hi[col], offY); }
        const float m = q.y / q.x;
        if (m < 999.f) { const unsigned long long key = ((unsigned long long)okey(m) << 32) | (unsigned)j; kmin = kmin < key ? kmin : key; }
        if (m > -999.f) { const unsigned long long key = ((unsigned long long)okey(m) << 32) | (0xFFFFFFFFu - (unsigned)j); kmax = kmax > key ? kmax : key; }
        if (q.z > maxZ) maxZ = q.z;
        sx += q.x; sy += q.y; sz += q.z;
        mnx = fminf(mnx, q.x); mny = fminf(mny, q.y); mnz = fminf(mnz, q.z);
        mxx = fmaxf(mxx, q.x); mxy = fmaxf(mxy, q.y); mxz = fmaxf(mxz, q.z);
      }
    }
    // thirteen block-wide reductions as ONE: shuffles inside the warp, one exchange of the per-warp partials through shared
    // memory, threads 0..12 fold one quantity each (26 CTA barriers before, 2 now)
    {
      const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long a = __shfl_xor_sync(0xFFFFFFFFu, kmin, o), b = __shfl_xor_sync(0xFFFFFFFFu, kmax, o);
        kmin = kmin < a ? kmin : a; kmax = kmax > b ? kmax : b;
        maxZ = fmaxf(maxZ, __shfl_xor_sync(0xFFFFFFFFu, maxZ, o));
        sx += __shfl_xor_sync(0xFFFFFFFFu, sx, o); sy += __shfl_xor_sync(0xFFFFFFFFu, sy, o); sz += __shfl_xor_sync(0xFFFFFFFFu, sz, o);
        mnx = fminf(mnx, __shfl_xor_sync(0xFFFFFFFFu, mnx, o)); mny = fminf(mny, __shfl_xor_sync(0xFFFFFFFFu, mny, o)); mnz = fminf(mnz, __shfl_xor_sync(0xFFFFFFFFu, mnz, o));
        mxx = fmaxf(mxx, __shfl_xor_sync(0xFFFFFFFFu, mxx, o)); mxy = fmaxf(mxy, __shfl_xor_sync(0xFFFFFFFFu, mxy, o)); mxz = fmaxf(mxz, __shfl_xor_sync(0xFFFFFFFFu, mxz, o));
      }
      __syncthreads();                       // (the previous cluster's readers of s_part are done)
      if (lane == 0) {
        s_part64[0][warp] = kmin; s_part64[1][warp] = kmax;
        s_partd[0][warp] = sx; s_partd[1][warp] = sy; s_partd[2][warp] = sz;
        s_partf[0][warp] = maxZ; s_partf[1][warp] = mnx; s_partf[2][warp] = mny; s_partf[3][warp] = mnz;
        s_partf[4][warp] = mxx; s_partf[5][warp] = mxy; s_partf[6][warp] = mxz;
      }
      __syncthreads();
      constexpr int NW = kFitThreads / 32;
      if (tid < 2) {
        unsigned long long r = s_part64[tid][0];
        for (int w = 1; w < NW; ++w) { const unsigned long long v = s_part64[tid][w]; r = (tid == 0) ? (r < v ? r : v) : (r > v ? r : v); }
        s_part64[tid][0] = r;
      } else if (tid < 5) {
        double r = s_partd[tid - 2][0];
        for (int w = 1; w < NW; ++w) r += s_partd[tid - 2][w];
        s_partd[tid - 2][0] = r;
      } else if (tid < 12) {
        const int q = tid - 5;
        float r = s_partf[q][0];
        for (int w = 1; w < NW; ++w) r = (q == 0 || q >= 4) ? fmaxf(r, s_partf[q][w]) : fminf(r, s_partf[q][w]);
        s_partf[q][0] = r;
      }
      __syncthreads();
      kmin = s_part64[0][0]; kmax = s_part64[1][0];
      sx = s_partd[0][0]; sy = s_partd[1][0]; sz = s_partd[2][0];
      maxZ = s_partf[0][0]; mnx = s_partf[1][0]; mny = s_partf[2][0]; mnz = s_partf[3][0];
      mxx = s_partf[4][0]; mxy = s_partf[5][0]; mxz = s_partf[6][0];
    }

    if (k == (int)blockIdx.x + 1) fit_mark(clk, 1);
    // first-occurrence min / max slope points (:268-280).  If no slope beats the 999 / -999 seeds the reference
    // reads uninitialised floats; defined here as (0,0).
    float minMx = 0.f, minMy = 0.f, maxMx = 0.f, maxMy = 0.f;
    if (kmin != ~0ull) { const float4 q = __ldg(&seg[(unsigned)(kmin & 0xFFFFFFFFull)]); minMx = q.x; minMy = q.y; }
    if (kmax != 0ull) { const float4 q = __ldg(&seg[0xFFFFFFFFu - (unsigned)(kmax & 0xFFFFFFFFull)]); maxMx = q.x; maxMy = q.y; }
    const float xDist = maxMx - minMx, yDist = maxMy - minMy;
    const float slopeDist = sqrtf(xDist * xDist + yDist * yDist);
    const float slope = (maxMy - minMy) / (maxMx - minMx);

    float pc[4][2];
    if (slopeDist > (float)P.l_slope_dist && n > P.l_num_points && (maxMy > 8.f || maxMy < -5.f)) {
      // ---- L-shape (:303-356): 80 draws of uniform_int_distribution<>(0,n-1) over mt19937_64(0), Lemire mapping
      // (libstdc++ >= 11): idx = hi64(raw*n), redraw while lo64 < 2^64 mod n.  First strictly larger distance wins.
      const float den = sqrtf(slope * slope + 1);
      const unsigned long long un = (unsigned long long)n;
      const unsigned long long thr = (0ull - un) % un;
      // one draw per thread; a Lemire redraw (probability n / 2^64) shifts the stream, then thread 0 replays it in order
      bool redraw = false;
      unsigned long long key = 0ull;                       // (dist bits << 32) | (~draw index): max = first largest dist
      float myx = 0.f, myy = 0.f;
      if (tid < P.ram_points) {
        const unsigned long long raw = mt_raw[tid % n_raw];
        redraw = raw * un < thr;
        const unsigned pInd = (unsigned)__umul64hi(raw, un);
        const float4 q = __ldg(&seg[pInd]);
        const float dist = fabsf(slope * q.x - 1 * q.y + maxMy - slope * maxMx) / den;
        myx = q.x; myy = q.y;
        if (dist > 0.f) key = ((unsigned long long)__float_as_uint(dist) << 32) | (0xFFFFFFFFu - (unsigned)tid);
      }
      const bool slow = __syncthreads_or(redraw) || P.ram_points > kFitThreads;
      if (!slow) {
        const unsigned long long best = block_reduce(key, MaxU64(), s_red64);
        const unsigned win = 0xFFFFFFFFu - (unsigned)(best & 0xFFFFFFFFull);
        if (best == 0ull) { if (tid == 0) { s_pc[1][0] = 0.f; s_pc[1][1] = 0.f; } }   // no dist > 0: uninitialised in the reference
        else if ((unsigned)tid == win) { s_pc[1][0] = myx; s_pc[1][1] = myy; }
      } else if (tid == 0) {
        float maxDist = 0.f, maxDx = 0.f, maxDy = 0.f;
        int cur = 0;
        for (int i = 0; i < P.ram_points; ++i) {
          unsigned long long raw = mt_raw[cur % n_raw]; ++cur;
          unsigned long long lo = raw * un;
          while (lo < thr) { raw = mt_raw[cur % n_raw]; ++cur; lo = raw * un; }
          const unsigned pInd = (unsigned)__umul64hi(raw, un);
          const float4 q = __ldg(&seg[pInd]);
          const float dist = fabsf(slope * q.x - 1 * q.y + maxMy - slope * maxMx) / den;
          if (dist > maxDist) { maxDist = dist; maxDx = q.x; maxDy = q.y; }
        }
        s_pc[1][0] = maxDx; s_pc[1][1] = maxDy;
      }
      __syncthreads();
      if (tid == 0) {
        const float maxDx = s_pc[1][0], maxDy = s_pc[1][1];
        const float maxMvecX = maxMx - maxDx, maxMvecY = maxMy - maxDy;
        const float minMvecX = minMx - maxDx, minMvecY = minMy - maxDy;
        s_pc[0][0] = minMx; s_pc[0][1] = minMy;
        s_pc[2][0] = maxMx; s_pc[2][1] = maxMy;
        s_pc[3][0] = maxDx + maxMvecX + minMvecX; s_pc[3][1] = maxDy + maxMvecY + minMvecY;
      }
      __syncthreads();
    } else {
      // ---- MAR contract (oracle/mar_contract.cpp): strict hull of the pixel set from the column extremes
      // strict convex hull of the pixel set, same canonical sequence as Andrew's monotone chain in
      // oracle/mar_contract.cpp (start = lexicographic minimum, lower chain left to right, upper chain right to left),
      // computed in parallel from the per-column extremes:
      //   (c, lo[c]) is a lower-hull vertex  <=>  max slope from any column on its left  <  min slope to any column on its right
      //   (c, hi[c]) is an upper-hull vertex <=>  min slope from the left                >  max slope to the right
      // (first / last occupied columns always are); slopes are compared exactly as integer cross products.
      int W;
      {
        // compact the occupied columns: thread t owns columns [8t, 8t+8)
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int c = tid * 8 + u; if (c < kCols && s_lo[c] != INT_MAX) ++cnt; }
        int incl = cnt;
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
        if (lane == 31) s_redi[warp] = incl;
        __syncthreads();
        int wbase = 0, tot = 0;
        for (int w = 0; w < kFitThreads / 32; ++w) { if (w < warp) wbase += s_redi[w]; tot += s_redi[w]; }
        int pos = wbase + incl - cnt;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int c = tid * 8 + u;
          if (c < kCols && s_lo[c] != INT_MAX) { s_cx[pos] = (short)(c - kColShift); s_clo[pos] = (short)s_lo[c]; s_chi[pos] = (short)s_hi[c]; ++pos; }
        }
        W = tot;
        __syncthreads();
      }
      for (int i = tid; i < W; i += kFitThreads) {
        const int xi = s_cx[i], li = s_clo[i], hi = s_chi[i];
        bool lowv = true, upv = true;
        if (i > 0 && i < W - 1) {
          // left side: running extreme slopes as (dy, dx), dx > 0
          int lmax_n = 0, lmax_d = 0, umin_n = 0, umin_d = 0;
          for (int j = 0; j < i; ++j) {
            const int dx = xi - s_cx[j];
            const int dl = li - s_clo[j], du = hi - s_chi[j];
            if (lmax_d == 0 || (long long)dl * lmax_d > (long long)lmax_n * dx) { lmax_n = dl; lmax_d = dx; }
            if (umin_d == 0 || (long long)du * umin_d < (long long)umin_n * dx) { umin_n = du; umin_d = dx; }
          }
          for (int k = i + 1; k < W && (lowv || upv); ++k) {
            const int dx = s_cx[k] - xi;
            const int dl = s_clo[k] - li, du = s_chi[k] - hi;
            if (!((long long)lmax_n * dx < (long long)dl * lmax_d)) lowv = false;   // need lmax < slope to k
            if (!((long long)umin_n * dx > (long long)du * umin_d)) upv = false;    // need umin > slope to k
          }
        }
        s_flag8[i] = (lowv ? 1 : 0) | (upv ? 2 : 0);
      }
      __syncthreads();
      // hull sequence = lower-chain vertices left to right, then upper-chain vertices right to left: positions from two block-wide
      // counts (thread t owns the compacted columns [8t, 8t+8)); as a loop on thread 0 this was 2 W dependent shared-memory
      // round trips -- 12 us for a 400-column cluster, the longest phase of the slowest CTA
      if (W == 1) {
        if (tid == 0) {
          int m = 1;
          s_hx[0] = s_cx[0]; s_hy[0] = s_clo[0];
          if (s_chi[0] != s_clo[0]) { s_hx[1] = s_cx[0]; s_hy[1] = s_chi[0]; m = 2; }
          s_m = m;
        }
      } else {
        const int lane = tid & 31, warp = tid >> 5;
        unsigned lowm = 0, upm = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = tid * 8 + u;
          if (i < W) {
            const int fl = s_flag8[i];
            if (fl & 1) lowm |= 1u << u;
            // the lo point of the first / last column already sits in the lower chain
            if ((fl & 2) && !((i == 0 || i == W - 1) && s_chi[i] == s_clo[i])) upm |= 1u << u;
          }
        }
        const int cl = __popc(lowm), cu = __popc(upm);
        int il = cl, iu = cu;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int a = __shfl_up_sync(0xFFFFFFFFu, il, o), b = __shfl_up_sync(0xFFFFFFFFu, iu, o);
          if (lane >= o) { il += a; iu += b; }
        }
        if (lane == 31) { s_redi[warp] = il; s_bidx[warp] = iu; }
        __syncthreads();
        int bl = 0, bu = 0, tl = 0, tu = 0;
        for (int w = 0; w < kFitThreads / 32; ++w) { if (w < warp) { bl += s_redi[w]; bu += s_bidx[w]; } tl += s_redi[w]; tu += s_bidx[w]; }
        int pl = bl + il - cl;                        // lower-chain vertices before this thread's columns
        int pu_after = tu - (bu + iu);                // upper-chain vertices AFTER this thread's columns (they come first, right to left)
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int i = tid * 8 + u;
          if ((lowm >> u) & 1u) { if (pl < kHullCap) { s_hx[pl] = s_cx[i]; s_hy[pl] = s_clo[i]; } ++pl; }
        }
#pragma unroll
        for (int u = 7; u >= 0; --u) {
          const int i = tid * 8 + u;
          if ((upm >> u) & 1u) { const int pos = tl + pu_after; if (pos < kHullCap) { s_hx[pos] = s_cx[i]; s_hy[pos] = s_chi[i]; } ++pu_after; }
        }
        if (tid == 0) {
          int m = tl + tu;
          if (m > kHullCap) { counters[CNT_ERROR] = LMOT_ERR_CAPACITY; m = kHullCap; }
          s_m = m;
        }
      }
      __syncthreads();
      const int m = s_m;
      float rc[4][2];
      if (m == 1) {
        for (int c = 0; c < 4; ++c) { rc[c][0] = (float)s_hx[0]; rc[c][1] = (float)s_hy[0]; }
      } else if (m == 2) {
        rc[0][0] = rc[1][0] = (float)s_hx[0]; rc[0][1] = rc[1][1] = (float)s_hy[0];
        rc[2][0] = rc[3][0] = (float)s_hx[1]; rc[2][1] = rc[3][1] = (float)s_hy[1];
      } else {
        // exact min-area edge: area_i = W*T/l compared as rationals, ties -> lowest edge index
        unsigned long long bnum = 0, bl = 1; int bidx = INT_MAX;
        for (int i = tid; i < m; i += kFitThreads) {
          const int i1 = (i + 1 == m) ? 0 : i + 1;
          const int dx = s_hx[i1] - s_hx[i], dy = s_hy[i1] - s_hy[i];
          int smin = INT_MAX, smax = INT_MIN, tmin = INT_MAX, tmax = INT_MIN;
          for (int j = 0; j < m; ++j) {
            const int s = s_hx[j] * dx + s_hy[j] * dy, t = -s_hx[j] * dy + s_hy[j] * dx;
            smin = min(smin, s); smax = max(smax, s); tmin = min(tmin, t); tmax = max(tmax, t);
          }
          const unsigned long long num = (unsigned long long)(smax - smin) * (unsigned long long)(tmax - tmin);
          const unsigned long long l = (unsigned long long)((long long)dx * dx + (long long)dy * dy);
          if (bidx == INT_MAX || num * bl < bnum * l) { bnum = num; bl = l; bidx = i; }   // i ascending per thread
        }
        // warp then block argmin with the same comparator (index breaks ties)
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long onum = __shfl_xor_sync(0xFFFFFFFFu, bnum, o), ol = __shfl_xor_sync(0xFFFFFFFFu, bl, o);
          const int oidx = __shfl_xor_sync(0xFFFFFFFFu, bidx, o);
          if (oidx != INT_MAX) {
            const bool better = (bidx == INT_MAX) || (onum * bl < bnum * ol) || (onum * bl == bnum * ol && oidx < bidx);
            if (better) { bnum = onum; bl = ol; bidx = oidx; }
          }
        }
        if ((tid & 31) == 0) { s_bnum[tid >> 5] = bnum; s_bl[tid >> 5] = bl; s_bidx[tid >> 5] = bidx; }
        __syncthreads();
        if (tid == 0) {
          for (int w = 1; w < kFitThreads / 32; ++w) {
            if (s_bidx[w] == INT_MAX) continue;
            const bool better = (bidx == INT_MAX) || (s_bnum[w] * bl < bnum * s_bl[w]) || (s_bnum[w] * bl == bnum * s_bl[w] && s_bidx[w] < bidx);
            if (better) { bnum = s_bnum[w]; bl = s_bl[w]; bidx = s_bidx[w]; }
          }
          s_best = bidx;
        }
        __syncthreads();
        const int best = s_best, b1 = (best + 1 == m) ? 0 : best + 1;
        long long ux = s_hx[b1] - s_hx[best], uy = s_hy[b1] - s_hy[best];
        for (int r = 0; r < 4 && !(ux >= 0 && uy < 0); ++r) { const long long t = ux; ux = -uy; uy = t; }
        const long long l = ux * ux + uy * uy;
        int smin = INT_MAX, smax = INT_MIN, tmin = INT_MAX, tmax = INT_MIN;
        for (int j = tid; j < m; j += kFitThreads) {
          const int s = (int)(s_hx[j] * ux + s_hy[j] * uy), t = (int)(-s_hx[j] * uy + s_hy[j] * ux);
          smin = min(smin, s); smax = max(smax, s); tmin = min(tmin, t); tmax = max(tmax, t);
        }
        smin = block_reduce(smin, MinI(), s_redi); smax = block_reduce(smax, MaxI(), s_redi);
        tmin = block_reduce(tmin, MinI(), s_redi); tmax = block_reduce(tmax, MaxI(), s_redi);
        const long long S[4] = {smin, smin, smax, smax}, T[4] = {tmax, tmin, tmin, tmax};
        for (int c = 0; c < 4; ++c) {
          rc[c][0] = (float)((double)(S[c] * ux - T[c] * uy) / (double)l);
          rc[c][1] = (float)((double)(S[c] * uy + T[c] * ux) / (double)l);
        }
      }
      // getPointsInPcFrame (:75-95)
      if (tid == 0) {
        for (int c = 0; c < 4; ++c) {
          const float rOffsetX = rc[c][0] - (float)offsetInitX, rOffsetY = rc[c][1] - (float)offsetInitY;
          const float rX = rOffsetX, rY = P.pic_scale * P.roi - rOffsetY;
          const float rmX = rX / P.pic_scale, rmY = rY / P.pic_scale;
          s_pc[c][0] = rmX - P.roi / 2; s_pc[c][1] = rmY - P.roi / 2;
        }
      }
      __syncthreads();
    }
    if (k == (int)blockIdx.x + 1) fit_mark(clk, 2);
    for (int c = 0; c < 4; ++c) { pc[c][0] = s_pc[c][0]; pc[c][1] = s_pc[c][1]; }
    const bool ok = rule_filter(pc, maxZ, n, P);
    if (tid == 0) {
      cl_ok[k] = ok ? 1 : 0;
      if (ok) {
        float* b = cl_box + (size_t)k * 24;
        for (int h = 0; h < 2; ++h)
          for (int c = 0; c < 4; ++c) { b[(h * 4 + c) * 3] = pc[c][0]; b[(h * 4 + c) * 3 + 1] = pc[c][1]; b[(h * 4 + c) * 3 + 2] = h == 0 ? -P.sensor_height : maxZ; }
        float* mk = cl_marker + (size_t)k * 6;      // mark_cluster (:161-209); zero extents become 0.1
        mk[0] = (float)(sx / n); mk[1] = (float)(sy / n); mk[2] = (float)(sz / n);
        float ex = mxx - mnx, ey = mxy - mny, ez = mxz - mnz;
        mk[3] = ex == 0.f ? 0.1f : ex; mk[4] = ey == 0.f ? 0.1f : ey; mk[5] = ez == 0.f ? 0.1f : ez;
      }
    }
    __syncthreads();
  }

  fit_mark(clk, 3);
  // ---- the last CTA of the frame compacts the accepted boxes in cluster-id order
  __threadfence();
  __syncthreads();
  if (tid == 0) s_flag = (atomicAdd(done, 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (!s_flag) return;
  __threadfence();
  int carry = 0;
  for (int k0 = 1; k0 <= K; k0 += kFitThreads) {
    const int k = k0 + tid;
    const int f = (k <= K) ? __ldcg((const unsigned char*)&cl_ok[k]) : 0;
    const unsigned bal = __ballot_sync(0xFFFFFFFFu, f);
    if ((tid & 31) == 0) s_redi[tid >> 5] = __popc(bal);
    __syncthreads();
    int wbase = 0, tot = 0;
    for (int w = 0; w < kFitThreads / 32; ++w) { if (w < (tid >> 5)) wbase += s_redi[w]; tot += s_redi[w]; }
    const int pos = carry + wbase + __popc(bal & ((1u << (tid & 31)) - 1u));
    if (f && pos < max_boxes) s_src[pos - carry] = k;
    __syncthreads();
    // cooperative copy: consecutive threads move consecutive floats
    const int nacc = min(tot, max_boxes - carry);
    for (int e = tid; e < nacc * 24; e += kFitThreads) {
      const int j = e / 24, w = e - j * 24;
      const float v = __ldcg(&cl_box[(size_t)s_src[j] * 24 + w]);
      boxes[(size_t)(carry + j) * 24 + w] = v;
    }
    for (int e = tid; e < nacc * 6; e += kFitThreads) {
      const int j = e / 6, w = e - j * 6;
      markers[(size_t)(carry + j) * 6 + w] = __ldcg(&cl_marker[(size_t)s_src[j] * 6 + w]);
    }
    carry += tot;
    __syncthreads();
  }
  if (tid == 0) {
    if (carry > max_boxes) { counters[CNT_ERROR] = LMOT_ERR_CAPACITY; carry = max_boxes; }
    counters[CNT_N_BOXES] = carry;
  }
  // the frame's detection stages are complete: post the slot's semaphore for the tracker's gate kernel (tracker.cu), which takes
  // the place of a stream-level event wait in front of the tracker chain
  if (det_sem) {
    __syncthreads();
    if (tid == 0) { __threadfence(); atomicAdd(det_sem, 1); }
  }
  fit_mark(clk, 4);
}

// batched submissions (one frame per sensor stream): the frames' box lists, concatenated in stream order, are ONE measurement
// list for the tracker (tracking/main.cpp:98-141 unpacks whatever the trackbox message holds; one immUkfJpdaf call per tick)
__global__ void __launch_bounds__(256)
concat_boxes_kernel(const __grid_constant__ FitBatch B, int n_frames, int max_boxes, float* __restrict__ boxes, int* __restrict__ counters,
                    int* __restrict__ frame_counts, int* __restrict__ det_sem) {
  __shared__ int s_off[kMaxBatch + 1];
  __shared__ int s_fc[kMaxBatch][5];
  // thread f fetches frame f's five counters (independent loads: one L2 round trip for the whole tick; a single thread walking the
  // frames was 40 dependent round trips, 20 us)
  if (threadIdx.x < n_frames) {
    const int f = threadIdx.x;
    int* fc = B.f[f].counters;
    const int ne = fc[CNT_N_ELEV], ng = fc[CNT_N_GROUND], nc = fc[CNT_NUM_CLUSTER], nb = fc[CNT_N_BOXES], er = fc[CNT_ERROR];
    s_fc[f][0] = ne; s_fc[f][1] = ng; s_fc[f][2] = nc; s_fc[f][3] = nb; s_fc[f][4] = er;
    frame_counts[4 * f] = ne; frame_counts[4 * f + 1] = ng; frame_counts[4 * f + 2] = nc; frame_counts[4 * f + 3] = nb;
    fc[CNT_ERROR] = 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0, err = 0, ne = 0, ng = 0, nc = 0;
    for (int f = 0; f < n_frames; ++f) {
      s_off[f] = acc;
      acc += s_fc[f][3];
      if (s_fc[f][4]) err = s_fc[f][4];
      ne += s_fc[f][0]; ng += s_fc[f][1]; nc += s_fc[f][2];
    }
    s_off[n_frames] = acc;
    if (acc > max_boxes) err = LMOT_ERR_CAPACITY;
    counters[CNT_N_BOXES] = acc < max_boxes ? acc : max_boxes;
    counters[CNT_ERROR] = err;
    // what the single-frame path reports per frame, summed over the tick (the tracker copies them into the result header)
    counters[CNT_N_ELEV] = ne; counters[CNT_N_GROUND] = ng; counters[CNT_NUM_CLUSTER] = nc;
  }
  __syncthreads();
  {
    // warp w copies frame w's list (96-byte boxes: six 16-byte words each, source and destination 16-byte aligned), four words per
    // lane in flight.  (One flat loop over all floats with a per-element frame search was 40 dependent L2 round trips: 18 us.)
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (w < n_frames) {
      const int off = s_off[w], cnt = min(s_off[w + 1], max_boxes) - off;
      const uint4* src = reinterpret_cast<const uint4*>(B.f[w].boxes);
      uint4* dst = reinterpret_cast<uint4*>(boxes + (size_t)off * 24);
      const int nw = cnt > 0 ? cnt * 6 : 0;
      for (int e0 = lane; e0 < nw; e0 += 128) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int e = e0 + 32 * u; if (e < nw) v[u] = src[e]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int e = e0 + 32 * u; if (e < nw) dst[e] = v[u]; }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();                // frame_counts lives in pinned host memory: the host reads it after the tick's completion event
    if (det_sem) atomicAdd(det_sem, 1);
  }
}

}  // namespace

int boxfit_alloc_shared(Ctx* c) {
  // raw mt19937_64(0) stream (box_fitting.cpp:303 re-seeds per cluster, so every cluster sees the same draws)
  constexpr int kRaw = 256;
  std::mt19937_64 mt(0);
  unsigned long long raw[kRaw];
  for (int i = 0; i < kRaw; ++i) raw[i] = mt();
  c->n_mt_raw = kRaw;
  c->max_sort_tiles = (c->max_points + kTile - 1) / kTile;
  LMOT_CUDA(c, cudaMalloc(&c->d_mt_raw, sizeof(raw)));
  LMOT_CUDA(c, cudaMemcpy(c->d_mt_raw, raw, sizeof(raw), cudaMemcpyHostToDevice));
  // tile_hist / scatter keep one int per cluster id in dynamic shared memory (lmot_create bounds max_clusters accordingly)
  const int sh = (c->prm.max_clusters + 1) * (int)sizeof(int);
  if (sh > 48 * 1024) {
    LMOT_CUDA(c, cudaFuncSetAttribute(tile_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sh));
    LMOT_CUDA(c, cudaFuncSetAttribute(scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sh));
  }
  return LMOT_OK;
}

int boxfit_alloc(Ctx* c, Slot* s) {
  const size_t np = (size_t)c->max_points;
  const int K1 = c->prm.max_clusters + 1;
  LMOT_CUDA(c, cudaMalloc(&s->d_pcid, np * sizeof(uint16_t)));
  LMOT_CUDA(c, cudaMalloc(&s->d_table, (size_t)c->max_sort_tiles * K1 * sizeof(int)));
  LMOT_CUDA(c, cudaMalloc(&s->d_seg_start, K1 * sizeof(int)));
  LMOT_CUDA(c, cudaMalloc(&s->d_seg_size, K1 * sizeof(int)));
  LMOT_CUDA(c, cudaMalloc(&s->d_sorted_pts, np * sizeof(float4)));
  LMOT_CUDA(c, cudaMalloc(&s->d_cl_box, (size_t)K1 * 24 * sizeof(float)));
  LMOT_CUDA(c, cudaMalloc(&s->d_cl_marker, (size_t)K1 * 6 * sizeof(float)));
  LMOT_CUDA(c, cudaMalloc(&s->d_cl_ok, K1));
  LMOT_CUDA(c, cudaMalloc(&s->d_boxes, (size_t)c->prm.max_boxes * 24 * sizeof(float)));
  LMOT_CUDA(c, cudaMalloc(&s->d_boxes_g, (size_t)c->prm.max_boxes * 24 * sizeof(float)));
  LMOT_CUDA(c, cudaMalloc(&s->d_markers, (size_t)c->prm.max_boxes * 6 * sizeof(float)));
  LMOT_CUDA(c, cudaMalloc(&s->d_done, 2 * sizeof(int)));                 // [0] finished box_fit CTAs, [1] finished tile_hist CTAs
  LMOT_CUDA(c, cudaMemsetAsync(s->d_done, 0, 2 * sizeof(int), s->stream));
  LMOT_CUDA(c, cudaMalloc(&s->d_det_sem, sizeof(int)));
  LMOT_CUDA(c, cudaMemsetAsync(s->d_det_sem, 0, sizeof(int), s->stream));
  return LMOT_OK;
}

void boxfit_free(Slot* s) {
  cudaFree(s->d_pcid); cudaFree(s->d_table); cudaFree(s->d_seg_start); cudaFree(s->d_seg_size); cudaFree(s->d_sorted_pts);
  cudaFree(s->d_cl_box); cudaFree(s->d_cl_marker); cudaFree(s->d_cl_ok); cudaFree(s->d_boxes); cudaFree(s->d_boxes_g); cudaFree(s->d_markers);
  cudaFree(s->d_done); cudaFree(s->d_det_sem);
}

static void fit_frame_of(const Slot* s, FitFrame& f, bool post_sem) {
  f.cart = s->d_cart; f.label_grid = s->d_label_grid; f.counters = s->d_counters;
  f.pcid = s->d_pcid; f.table = s->d_table; f.seg_start = s->d_seg_start; f.seg_size = s->d_seg_size;
  f.elev = s->d_elev; f.sorted_pts = s->d_sorted_pts;
  f.cl_box = s->d_cl_box; f.cl_marker = s->d_cl_marker; f.cl_ok = s->d_cl_ok;
  f.boxes = s->d_boxes; f.markers = s->d_markers; f.done = s->d_done; f.det_sem = post_sem ? s->d_det_sem : nullptr;
  f.hist_ctr = reinterpret_cast<unsigned*>(s->d_done + 1);
}

// inputs per frame: s->d_elev / CNT_N_ELEV, s->d_cart (from clustering), s->d_label_grid / CNT_NUM_CLUSTER.
// F > 1: one frame per sensor stream; the four kernels run over the union of the frames' clusters (blockIdx.y = frame).
int boxfit_launch_batch(Ctx* c, Slot* const* slots, int F, cudaStream_t st, const int* n_upper, bool post_sem) {
  if (F < 1 || F > kMaxBatch) return LMOT_ERR_INVALID;
  const int K1 = c->prm.max_clusters + 1;
  int tiles = 0;
  for (int i = 0; i < F; ++i) { const int t = (n_upper[i] + kTile - 1) / kTile; if (t > tiles) tiles = t; }
  const size_t sh = (size_t)K1 * sizeof(int);
  FitBatch B;
  for (int i = 0; i < F; ++i) fit_frame_of(slots[i], B.f[i], post_sem && F == 1);
  for (int i = F; i < kMaxBatch; ++i) B.f[i] = B.f[0];
  Slot* s0 = slots[0];
  if (tiles > 0) tile_hist_kernel<<<dim3(tiles, F), kTile, sh, st>>>(B, c->prm.max_clusters);     // + segment offsets by its last CTA per frame
  else seg_offsets_kernel<<<F, 1024, 0, st>>>(B, c->prm.max_clusters);
  kernel_mark(c, s0, st);
  // scatter and box_fit as programmatic dependents of the kernel before them (timing mode records an event between the kernels,
  // which needs the ordinary full serialisation)
  cudaLaunchAttribute pdl;
  pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl.val.programmaticStreamSerializationAllowed = c->timing ? 0 : 1;
  const int max_clusters = c->prm.max_clusters;
  if (tiles > 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(tiles, F); cfg.blockDim = dim3(kTile); cfg.dynamicSmemBytes = sh; cfg.stream = st; cfg.attrs = &pdl; cfg.numAttrs = 1;
    LMOT_CUDA(c, cudaLaunchKernelEx(&cfg, scatter_kernel, B, max_clusters));
    kernel_mark(c, s0, st);
  }
  BoxParams P;
  const lmot_params& p = c->prm;
  P.roi = p.roi_m;
  { volatile float ps = 900 / p.roi_m; P.pic_scale = ps; }   // float picScale = 900/roiM (box_fitting.cpp:18)
  P.sensor_height = p.sensor_height; P.ram_points = p.ram_points; P.l_slope_dist = p.l_slope_dist;
  P.l_num_points = p.l_num_points; P.min_points = p.min_cluster_points; P.rule_mode = p.rule_filter;
  P.t_height_min = p.t_height_min; P.t_height_max = p.t_height_max; P.t_width_min = p.t_width_min; P.t_width_max = p.t_width_max;
  P.t_len_min = p.t_len_min; P.t_len_max = p.t_len_max; P.t_area_max = p.t_area_max; P.t_ratio_min = p.t_ratio_min;
  P.t_ratio_max = p.t_ratio_max; P.min_len_ratio = p.min_len_ratio; P.t_pt_per_m3 = p.t_pt_per_m3;
  // one CTA per cluster: a frame rarely holds more than a few hundred clusters, F frames share the grid
  const int gx = F == 1 ? c->fit_ctas : (c->fit_ctas / 2 > 8 ? c->fit_ctas / 2 : 8);
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(gx, F); cfg.blockDim = dim3(F == 1 ? kFitThreadsFrame : kFitThreadsBatch); cfg.dynamicSmemBytes = 0; cfg.stream = st; cfg.attrs = &pdl; cfg.numAttrs = 1;
    const unsigned long long* mt = c->d_mt_raw;
    const int n_raw = c->n_mt_raw, max_boxes = c->prm.max_boxes;
    unsigned long long* clk = c->d_fit_clock;
    if (F == 1) LMOT_CUDA(c, cudaLaunchKernelEx(&cfg, box_fit_kernel<kFitThreadsFrame>, B, P, mt, n_raw, max_clusters, max_boxes, clk));
    else LMOT_CUDA(c, cudaLaunchKernelEx(&cfg, box_fit_kernel<kFitThreadsBatch>, B, P, mt, n_raw, max_clusters, max_boxes, clk));
  }
  c->last_fit_ctas = gx * F;
  kernel_mark(c, s0, st);
  LMOT_CUDA(c, cudaGetLastError());
  return LMOT_OK;
}

// the frames' box lists concatenated in stream order -> d_boxes / counters of the batch (the tracker's measurement list)
int boxes_concat_launch(Ctx* c, Slot* const* slots, int F, cudaStream_t st, float* d_boxes, int* d_counters, int* d_frame_counts, int* det_sem) {
  FitBatch B;
  for (int i = 0; i < F; ++i) fit_frame_of(slots[i], B.f[i], false);
  for (int i = F; i < kMaxBatch; ++i) B.f[i] = B.f[0];
  concat_boxes_kernel<<<1, 256, 0, st>>>(B, F, c->prm.max_boxes, d_boxes, d_counters, d_frame_counts, det_sem);
  kernel_mark(c, slots[0], st);
  LMOT_CUDA(c, cudaGetLastError());
  return LMOT_OK;
}

// n_lists padded device box lists (capacity cap boxes each, lengths in d_counts) -> one contiguous list + the counter block the
// tracker reads (multi-GPU hand-over: the lists arrived by ncclAllGather)
__global__ void __launch_bounds__(256)
pack_lists_kernel(const float* __restrict__ lists, const int* __restrict__ counts, int n_lists, int cap, int max_boxes, float* __restrict__ boxes,
                  int* __restrict__ counters) {
  __shared__ int s_off[kMaxBatch + 1];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int l = 0; l < n_lists; ++l) { s_off[l] = acc; acc += min(max(counts[l], 0), cap); }
    s_off[n_lists] = acc;
    counters[CNT_N_BOXES] = acc < max_boxes ? acc : max_boxes;
    counters[CNT_ERROR] = acc > max_boxes ? (int)LMOT_ERR_CAPACITY : 0;
    counters[CNT_N_ELEV] = 0; counters[CNT_N_GROUND] = 0; counters[CNT_NUM_CLUSTER] = 0;
  }
  __syncthreads();
  for (int l = 0; l < n_lists; ++l) {
    const int off = s_off[l], cnt = min(s_off[l + 1], max_boxes) - off;
    const float* src = lists + (size_t)l * cap * 24;
    for (int e = threadIdx.x; e < cnt * 24; e += 256) boxes[(size_t)off * 24 + e] = src[e];
  }
}

int boxes_pack_lists_launch(Ctx* c, cudaStream_t st, const float* d_lists, const int* d_counts, int n_lists, int cap, float* d_boxes, int* d_counters) {
  pack_lists_kernel<<<1, 256, 0, st>>>(d_lists, d_counts, n_lists, cap, c->prm.max_boxes, d_boxes, d_counters);
  LMOT_CUDA(c, cudaGetLastError());
  return LMOT_OK;
}

int boxfit_launch(Ctx* c, Slot* s, cudaStream_t st, int n_upper, bool post_sem) {
  Slot* sl[1] = {s};
  const int nu[1] = {n_upper};
  return boxfit_launch_batch(c, sl, 1, st, nu, post_sem);
}

}  // namespace lmot

// ground.cu -- slope-based polar-grid ground removal on sm_100a.
//
// Replaces groundRemove (/root/reference/object_tracking/src/groundremove/ground_removal.cpp:177-249) and
// gaussSmoothen (src/groundremove/gaus_blur.cpp:52-68).  Three kernels per frame:
//
//   K1 polar_bin_kernel      N threads.  float4 XYZI load, range filter (:46-64), bit-exact point->cell
//                            (:67-76, exact_math.cuh), warp-aggregated atomicMin of an order-preserving
//                            key into the 80x120 min-z grid (:79-92).  Also stores the u16 cell id so the
//                            classification pass does not repeat atan2f/sqrtf like the reference does.
//   K2 polar_grid_kernel     ONE CTA, the whole 9,600-cell grid in shared memory: height clamp (:192-197),
//                            3-tap blur in fp64 (gaus_blur.cpp), hDiff (:95-117), ground flag (:205-214),
//                            median filter (:120-146, evaluated Jacobi-style -- provably order independent),
//                            outlier filter (:149-174, sequential along bin, one lane per channel).
//   K3 classify_partition_kernel  N threads.  ground / elevated decision (:221-247) + ORDER-PRESERVING
//                            compaction of both output clouds with a single-pass decoupled look-back scan.
//
// Everything that decides a cell index or a label is IEEE round-to-nearest without FMA contraction, so the
// results are bit-identical to the reference built for x86-64.
#include "lmot_internal.cuh"
#include "exact_math.cuh"
#include "tma.cuh"

namespace lmot {

namespace {

__device__ __forceinline__ unsigned fkey(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// ground_removal.cpp:46-64 (range filter) + :67-76 (getCellIndexFromPoints) + :89 (index guard)
//
// The channel index must equal floor(fl(fl((atan2f_glibc(y,x) + pi) / 2pi) * 80)) bit for bit.  Evaluating the exact
// fdlibm restatement (two IEEE divisions + an 11-term polynomial without FMA, ~150 instructions) for every point makes
// this kernel instruction-bound, so it is only used where it can matter: CUDA's own atan2f (<= 2 ulp) and glibc's
// (<= 2 ulp) differ by < 1e-6 rad, i.e. < 1.3e-5 channel widths, and the reference's float roundings of the scaled angle
// move it by < 1e-5 more.  A point whose fast scaled angle lies further than kChanGuard = 2e-4 channel widths from
// every integer therefore has the same floor() under both evaluations; the rest (about 4 points in 10,000) take the
// exact path.  The result is identical to the exact path for every point (tests/test_ground_gpu.py compares cells of
// random, HDL-64 and boundary-hugging clouds against the reference bit for bit).
constexpr float kChanGuard = 2.0e-4f;

__device__ __forceinline__ uint16_t polar_cell(float x, float y, const GroundParams& p) {
  const float d = fsqrt(fadd(fmul(x, x), fmul(y, y)));
  if (d <= p.r_min || d >= p.r_max || d != d) return kNoCell;
  const float binP = fdiv(fsub(d, p.r_min), p.r_span);
  const float binF = floorf(fmul(binP, (float)kNumBin));
  if (!(binF >= 0.f && binF < (float)kNumBin)) return kNoCell;
  // fast path
  const float t = (atan2f(y, x) + 3.14159265358979323846f) * (float)(kNumChannel / 6.28318530717958647692);
  const float tf = floorf(t);
  const float fr = t - tf;
  float chF = tf;
  if (!(fr > kChanGuard && fr < 1.0f - kChanGuard)) {
    // exact path: glibc's atan2f, double add / divide, narrowing, float multiply, floor -- as the reference evaluates it
    const float a = atan2f_fdlibm(y, x);
    const double chD = __ddiv_rn(__dadd_rn((double)a, 3.14159265358979323846), 6.28318530717958647692);
    chF = floorf(fmul((float)chD, (float)kNumChannel));
  }
  if (!(chF >= 0.f && chF < (float)kNumChannel)) return kNoCell;
  return (uint16_t)((int)chF * kNumBin + (int)binF);
}

constexpr int kBinTile = 256;                      // points per TMA tile (4 KB), two tiles in flight per CTA

// K1.  Persistent CTAs; the XYZI frame is staged through shared memory by the bulk async copy engine (TMA, UBLKCP in
// SASS): one elected thread arms an mbarrier with the tile's byte count and issues a single 4 KB cp.async.bulk, the
// other 255 threads never touch the LSU for input -- they pick their point up from shared memory when the barrier
// flips, and the copy of tile t+2 overlaps the binning arithmetic of tile t.
__global__ void __launch_bounds__(kBinTile) polar_bin_kernel(const float4* __restrict__ pts, int n, GroundParams p,
                                                             uint16_t* __restrict__ cell, unsigned* __restrict__ keys) {
  __shared__ alignas(128) float4 s_buf[2][kBinTile];
  __shared__ alignas(8) uint64_t s_full[2];
  const int tid = threadIdx.x, lane = tid & 31;
  const int n_tiles = (n + kBinTile - 1) / kBinTile;
  if (tid == 0) { mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1); fence_mbar_init(); }
  __syncthreads();
  if (tid == 0) {
#pragma unroll
    for (int st = 0; st < 2; ++st) {
      const int tile = blockIdx.x + st * gridDim.x;
      if (tile < n_tiles) {
        const uint32_t bytes = (uint32_t)min(kBinTile, n - tile * kBinTile) * 16u;
        mbar_arrive_expect_tx(&s_full[st], bytes);
        bulk_copy_g2s(s_buf[st], pts + (size_t)tile * kBinTile, bytes, &s_full[st]);
      }
    }
  }
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int st = it & 1;
    mbar_wait(&s_full[st], (uint32_t)(it >> 1) & 1u);
    const int i = tile * kBinTile + tid;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) q = s_buf[st][tid];
    __syncthreads();                                 // every thread holds its point: the stage can be refilled
    if (tid == 0) {
      const int next = tile + 2 * gridDim.x;
      if (next < n_tiles) {
        const uint32_t bytes = (uint32_t)min(kBinTile, n - next * kBinTile) * 16u;
        mbar_arrive_expect_tx(&s_full[st], bytes);
        bulk_copy_g2s(s_buf[st], pts + (size_t)next * kBinTile, bytes, &s_full[st]);
      }
    }
    unsigned c = kNoCell;
    unsigned key = 0xFFFFFFFFu;
    if (i < n) {
      c = polar_cell(q.x, q.y, p);
      cell[i] = (uint16_t)c;
      float z = q.z;
      if (z == 0.f) z = 0.f;                       // -0 -> +0 (`z < minZ` does not order them either)
      if (c != kNoCell && z == z) key = fkey(z);   // NaN z never wins `z < minZ` (ground_removal.cpp:41)
    }
    // consecutive HDL-64 returns fall into the same cell: one atomic per distinct cell per warp
    const unsigned grp = __match_any_sync(0xFFFFFFFFu, c);
    const unsigned kmin = __reduce_min_sync(grp, key);
    if (c != kNoCell && lane == __ffs(grp) - 1 && kmin != 0xFFFFFFFFu) atomicMin(&keys[c], kmin);
  }
}

// generic-stride input -> float4 (the hot path is stride 4 and never runs this)
__global__ void repack_kernel(const float* __restrict__ in, int n, int stride, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = make_float4(in[(size_t)i * stride], in[(size_t)i * stride + 1], in[(size_t)i * stride + 2], 1.f);
}

constexpr int kGridCtas = 8;                       // 10 channels per CTA (+1 halo channel each side, recomputed locally)
constexpr int kChanPerCta = kNumChannel / kGridCtas;
constexpr int kGridThreads = 512;
constexpr int kLocChan = kChanPerCta + 2;
constexpr int kLocCells = kLocChan * kNumBin;       // 1440
constexpr int kCellsPerThread = (kLocCells + kGridThreads - 1) / kGridThreads;   // 3

// The 80 channels are independent in every stage except the median filter, which looks one channel to each side:
// each CTA owns 10 channels and recomputes the clamp / blur / hDiff / flag stages of one halo channel per side, so no
// inter-CTA synchronisation is needed.  Keys are re-armed by classify_partition_kernel (a neighbour may still be
// reading this CTA's boundary channel here).
__global__ void __launch_bounds__(kGridThreads, 1)
polar_grid_kernel(GroundParams p, const unsigned* __restrict__ keys, float* __restrict__ o_minz, float* __restrict__ o_height,
                  float* __restrict__ o_smoothed, float* __restrict__ o_hdiff, float* __restrict__ o_hg,
                  unsigned long long* __restrict__ tile_desc, int n_tiles, int* __restrict__ counters) {
  __shared__ float H[kLocCells];
  __shared__ uint8_t G[kLocCells];
  const int tid = threadIdx.x;
  const int ch0 = blockIdx.x * kChanPerCta - 1;              // global channel of local channel 0 (may be -1)

  if (blockIdx.x == 0) {                                     // housekeeping for the kernels that follow in this frame
    for (int t = tid; t < n_tiles; t += kGridThreads) tile_desc[t] = 0ull;
    if (tid == 0) { counters[CNT_TICKET_A] = 0; counters[CNT_N_ELEV] = 0; counters[CNT_N_GROUND] = 0; }
  }

  // (a4) height clamp, ground_removal.cpp:192-197
#pragma unroll
  for (int j = 0; j < kCellsPerThread; ++j) {
    const int l = tid + j * kGridThreads;
    if (l < kLocCells) {
      const int lc = l / kNumBin, b = l - lc * kNumBin, ch = ch0 + lc;
      float h = 0.f;
      if (ch >= 0 && ch < kNumChannel) {
        const float zi = fkey_inv(__ldg(&keys[ch * kNumBin + b]));
        if (lc >= 1 && lc <= kChanPerCta) o_minz[ch * kNumBin + b] = zi;
        if (zi > p.t_hmin && zi < p.t_hmax) h = zi;
        else if (zi > p.t_hmax) h = p.h_sensor;
        else h = p.t_hmin;
      }
      H[l] = h;
    }
  }
  __syncthreads();

  // (a5) blur, (a6) hDiff, (a7) ground flag -- per channel, neighbours along bin
#pragma unroll
  for (int j = 0; j < kCellsPerThread; ++j) {
    const int l = tid + j * kGridThreads;
    if (l < kLocCells) {
      const int lc = l / kNumBin, b = l - lc * kNumBin, ch = ch0 + lc;
      uint8_t g = 0;
      if (ch >= 0 && ch < kNumChannel) {
        const float h = H[l];
        double acc = 0.0;                                       // gaus_blur.cpp:58-65, order j = i-1, i, i+1
        if (b > 0) acc = __dadd_rn(acc, __dmul_rn(p.tap[0], (double)H[l - 1]));
        acc = __dadd_rn(acc, __dmul_rn(p.tap[1], (double)h));
        if (b < kNumBin - 1) acc = __dadd_rn(acc, __dmul_rn(p.tap[2], (double)H[l + 1]));
        const float sm = (float)acc;
        float hd;                                               // ground_removal.cpp:95-117
        if (b == 0) hd = fsub(h, H[l + 1]);
        else if (b == kNumBin - 1) hd = fsub(h, H[l - 1]);
        else {
          const float pre = fsub(h, H[l - 1]), post = fsub(h, H[l + 1]);
          hd = (pre > post) ? pre : post;
        }
        if (lc >= 1 && lc <= kChanPerCta) { o_smoothed[ch * kNumBin + b] = sm; o_hdiff[ch * kNumBin + b] = hd; }
        g = ((sm < p.t_hmax && hd < p.t_hdiff) || (h < p.t_hmax && hd < p.t_hdiff)) ? 1 : 0;  // :205-214
      }
      G[l] = g;
    }
  }
  __syncthreads();

  // (a8) applyMedianFilter, ground_removal.cpp:120-146.  A cell flips only if its four neighbours are ground
  // already, and a neighbour that flips in this pass would have needed this cell to be ground: the in-place
  // sequential pass and this two-phase (decide, then apply) pass are identical.  Own channels only.
  {
    float newh[kCellsPerThread];
    unsigned flip = 0;
#pragma unroll
    for (int j = 0; j < kCellsPerThread; ++j) {
      const int l = tid + j * kGridThreads;
      newh[j] = 0.f;
      if (l < kLocCells) {
        const int lc = l / kNumBin, b = l - lc * kNumBin, ch = ch0 + lc;
        if (lc >= 1 && lc <= kChanPerCta && ch >= 1 && ch < kNumChannel - 1 && b >= 1 && b < kNumBin - 1 && !G[l] && G[l + 1] &&
            G[l - 1] && G[l + kNumBin] && G[l - kNumBin]) {
          const float a = H[l + 1], bb = H[l - 1], c = H[l + kNumBin], d = H[l - kNumBin];
          const float lo1 = fminf(a, bb), hi1 = fmaxf(a, bb), lo2 = fminf(c, d), hi2 = fmaxf(c, d);
          const float m1 = fmaxf(lo1, lo2), m2 = fminf(hi1, hi2);  // the two middle values of the sorted four
          newh[j] = fdiv(fadd(m1, m2), 2.f);
          flip |= 1u << j;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kCellsPerThread; ++j)
      if (flip & (1u << j)) { const int l = tid + j * kGridThreads; H[l] = newh[j]; G[l] = 1; }
  }
  __syncthreads();

  // (a8) outlierFilter, ground_removal.cpp:149-174: in place along bin, so cell b sees the value written at b-1.
  // With T = tHmin, a cell is rewritten iff  A(b): all of b-1..b+2 ground, H[b]==T and (H[b+1]!=T or H[b+2]!=T),
  // and its left value is not T -- either originally, or because b-1 was itself rewritten.  Two consecutive rewrites
  // force H[b-1]==H[b]==T and H[b+1]!=T, which rules out a rewrite at b-2 (its b+2 is H[b]==T, its b+1 is T): the chain
  // is at most two cells long, so every cell's final value is a closed form of the ORIGINAL H[b-2..b+2], G[b-2..b+2].
  {
    const float T = p.t_hmin;
    float newh[kCellsPerThread];
    unsigned mod = 0;
#pragma unroll
    for (int j = 0; j < kCellsPerThread; ++j) {
      const int l = tid + j * kGridThreads;
      newh[j] = 0.f;
      if (l < kLocCells) {
        const int lc = l / kNumBin, b = l - lc * kNumBin, ch = ch0 + lc;
        if (lc >= 1 && lc <= kChanPerCta && ch >= 1 && ch < kNumChannel - 1 && b >= 1 && b < kNumBin - 2 && G[l] && G[l + 1] &&
            G[l - 1] && G[l + 2]) {
          const float h1 = H[l - 1], h2 = H[l], h3 = H[l + 1], h4 = H[l + 2];
          if (h2 == T && (h3 != T || h4 != T)) {                  // A(b)
            float left = h1;
            bool ok = (h1 != T);
            if (!ok && b - 1 >= 1 && G[l - 2]) {                  // was b-1 rewritten?  A(b-1) with H[b]==T needs H[b+1]!=T
              const float h0 = H[l - 2];
              if (h3 != T && h0 != T) { left = fdiv(fadd(h0, h3), 2.f); ok = true; }   // b-1 took its second branch
            }
            if (ok) { newh[j] = (h3 != T) ? fdiv(fadd(left, h3), 2.f) : fdiv(fadd(left, h4), 2.f); mod |= 1u << j; }
          }
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kCellsPerThread; ++j)
      if (mod & (1u << j)) H[tid + j * kGridThreads] = newh[j];
  }
  __syncthreads();

  // hGround == height for every ground cell (updateGround() follows every height write of a ground cell)
#pragma unroll
  for (int j = 0; j < kCellsPerThread; ++j) {
    const int l = tid + j * kGridThreads;
    if (l < kLocCells) {
      const int lc = l / kNumBin, b = l - lc * kNumBin, ch = ch0 + lc;
      if (lc >= 1 && lc <= kChanPerCta) {
        o_height[ch * kNumBin + b] = H[l];
        o_hg[ch * kNumBin + b] = G[l] ? H[l] : -INFINITY;
      }
    }
  }
}

// status (2 bits) | elevated count (31 bits) | ground count (31 bits)
__device__ __forceinline__ unsigned long long pack_desc(unsigned st, unsigned e, unsigned g) {
  return ((unsigned long long)st << 62) | ((unsigned long long)e << 31) | (unsigned long long)g;
}

__global__ void __launch_bounds__(kScanTile)
classify_partition_kernel(const float4* __restrict__ pts, int n, const uint16_t* __restrict__ cell,
                          const float* __restrict__ hg, double tol, uint8_t* __restrict__ labels,
                          float4* __restrict__ elev, float4* __restrict__ ground,
                          unsigned long long* tile_desc, int* counters, float roi, uint16_t* __restrict__ cart,
                          int* __restrict__ cart_count, unsigned* __restrict__ keys) {
  __shared__ int s_tile;
  __shared__ unsigned s_we[32], s_wg[32];
  __shared__ unsigned s_base_e, s_base_g;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // re-arm the min-z grid for the next frame (Cell::Cell(): minZ = 1000); polar_grid_kernel has consumed it
  for (int k = blockIdx.x * kScanTile + tid; k < kPolarCells; k += gridDim.x * kScanTile) keys[k] = fkey(1000.f);
  if (tid == 0) s_tile = atomicAdd(&counters[CNT_TICKET_A], 1);
  __syncthreads();
  const int tile = s_tile;
  const int i = tile * kScanTile + tid;

  int lab = 0;
  float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < n) {
    q = __ldg(&pts[i]);
    const unsigned c = cell[i];
    if (c != kNoCell) {
      const float h = __ldg(&hg[c]);                       // -inf for non-ground cells -> elevated
      lab = ((double)q.z < __dadd_rn((double)h, tol)) ? 1 : 2;   // ground_removal.cpp:236-246
    }
    labels[i] = (uint8_t)lab;
  }
  const unsigned be = __ballot_sync(0xFFFFFFFFu, lab == 2);
  const unsigned bg = __ballot_sync(0xFFFFFFFFu, lab == 1);
  if (lane == 0) { s_we[warp] = __popc(be); s_wg[warp] = __popc(bg); }
  __syncthreads();

  if (warp == 0) {
    const unsigned ve = s_we[lane], vg = s_wg[lane];
    unsigned ie = ve, ig = vg;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned te = __shfl_up_sync(0xFFFFFFFFu, ie, o), tg = __shfl_up_sync(0xFFFFFFFFu, ig, o);
      if (lane >= o) { ie += te; ig += tg; }
    }
    s_we[lane] = ie - ve;
    s_wg[lane] = ig - vg;
    const unsigned agg_e = __shfl_sync(0xFFFFFFFFu, ie, 31), agg_g = __shfl_sync(0xFFFFFFFFu, ig, 31);
    volatile unsigned long long* desc = tile_desc;
    if (lane == 0 && tile > 0) desc[tile] = pack_desc(1u, agg_e, agg_g);
    // decoupled look-back, 32 predecessors per round
    unsigned ex_e = 0, ex_g = 0;
    int j = tile - 1 - lane;
    while (true) {
      unsigned long long d = (j >= 0) ? desc[j] : pack_desc(2u, 0u, 0u);
      while (__any_sync(0xFFFFFFFFu, (d >> 62) == 0ull)) {
        if ((d >> 62) == 0ull) d = desc[j];
      }
      const unsigned pm = __ballot_sync(0xFFFFFFFFu, (d >> 62) == 2ull);
      const int first = pm ? (__ffs(pm) - 1) : 31;
      unsigned ce = (lane <= first) ? (unsigned)((d >> 31) & 0x7FFFFFFFull) : 0u;
      unsigned cg = (lane <= first) ? (unsigned)(d & 0x7FFFFFFFull) : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) { ce += __shfl_xor_sync(0xFFFFFFFFu, ce, o); cg += __shfl_xor_sync(0xFFFFFFFFu, cg, o); }
      ex_e += ce; ex_g += cg;
      if (pm) break;
      j -= 32;
    }
    if (lane == 0) {
      desc[tile] = pack_desc(2u, ex_e + agg_e, ex_g + agg_g);
      s_base_e = ex_e; s_base_g = ex_g;
      if (tile == (int)gridDim.x - 1) { counters[CNT_N_ELEV] = (int)(ex_e + agg_e); counters[CNT_N_GROUND] = (int)(ex_g + agg_g); }
    }
  }
  __syncthreads();
  const unsigned lt = (1u << lane) - 1u;
  unsigned cc = kNoCell;
  if (lab == 2) {
    const unsigned pos = s_base_e + s_we[warp] + __popc(be & lt);
    elev[pos] = make_float4(q.x, q.y, q.z, 1.f);
    if (cart_count) { cc = cart_cell_of(q.x, q.y, roi, kNumGrid); cart[pos] = (uint16_t)cc; }
  } else if (lab == 1) ground[s_base_g + s_wg[warp] + __popc(bg & lt)] = make_float4(q.x, q.y, q.z, 1.f);
  // fused first pass of clustering (mapCartesianGrid, component_clustering.cpp:38-47): the elevated point is still in
  // registers, so bin it now instead of re-reading the elevated cloud in a separate kernel
  if (cart_count) {
    const unsigned grp = __match_any_sync(0xFFFFFFFFu, cc);
    if (cc != kNoCell && lane == __ffs(grp) - 1) atomicAdd(&cart_count[cc], __popc(grp));
  }
}

__global__ void init_keys_kernel(unsigned* keys) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < kPolarCells) keys[k] = fkey(1000.f);
}

}  // namespace

int ground_alloc(Ctx* c, Slot* s) {
  const size_t np = (size_t)c->max_points;
  c->max_tiles = (c->max_points + kScanTile - 1) / kScanTile;
  cudaStream_t st = s->stream;
  LMOT_CUDA(c, cudaMalloc(&s->d_points, np * sizeof(float4)));
  LMOT_CUDA(c, cudaMalloc(&s->d_stage_in, np * 4 * sizeof(float)));
  LMOT_CUDA(c, cudaMalloc(&s->d_cell, np * sizeof(uint16_t)));
  LMOT_CUDA(c, cudaMalloc(&s->d_polar_key, kPolarCells * sizeof(unsigned)));
  LMOT_CUDA(c, cudaMalloc(&s->d_minz, kPolarCells * sizeof(float)));
  LMOT_CUDA(c, cudaMalloc(&s->d_height, kPolarCells * sizeof(float)));
  LMOT_CUDA(c, cudaMalloc(&s->d_smoothed, kPolarCells * sizeof(float)));
  LMOT_CUDA(c, cudaMalloc(&s->d_hdiff, kPolarCells * sizeof(float)));
  LMOT_CUDA(c, cudaMalloc(&s->d_hg, kPolarCells * sizeof(float)));
  LMOT_CUDA(c, cudaMalloc(&s->d_labels, np));
  LMOT_CUDA(c, cudaMalloc(&s->d_elev, np * sizeof(float4)));
  LMOT_CUDA(c, cudaMalloc(&s->d_ground, np * sizeof(float4)));
  LMOT_CUDA(c, cudaMalloc(&s->d_tile_desc, (size_t)c->max_tiles * sizeof(unsigned long long)));
  LMOT_CUDA(c, cudaMalloc(&s->d_counters, CNT_COUNT * sizeof(int)));
  LMOT_CUDA(c, cudaMemsetAsync(s->d_counters, 0, CNT_COUNT * sizeof(int), st));
  LMOT_CUDA(c, cudaHostAlloc(&s->h_counters, CNT_COUNT * sizeof(int), cudaHostAllocDefault));
  LMOT_CUDA(c, cudaHostAlloc(&s->h_set, CNT_COUNT * sizeof(int), cudaHostAllocDefault));
  init_keys_kernel<<<(kPolarCells + 255) / 256, 256, 0, st>>>(s->d_polar_key);
  LMOT_CUDA(c, cudaGetLastError());
  return LMOT_OK;
}

void ground_free(Slot* s) {
  cudaFree(s->d_points); cudaFree(s->d_stage_in); cudaFree(s->d_cell); cudaFree(s->d_polar_key); cudaFree(s->d_minz);
  cudaFree(s->d_height); cudaFree(s->d_smoothed); cudaFree(s->d_hdiff); cudaFree(s->d_hg); cudaFree(s->d_labels);
  cudaFree(s->d_elev); cudaFree(s->d_ground); cudaFree(s->d_tile_desc); cudaFree(s->d_counters);
  if (s->h_counters) cudaFreeHost(s->h_counters);
  if (s->h_set) cudaFreeHost(s->h_set);
}

int ground_repack(Ctx* c, cudaStream_t st, const float* d_in, int n, int stride, float4* d_out) {
  if (n > 0) repack_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_in, n, stride, d_out);
  LMOT_CUDA(c, cudaGetLastError());
  return LMOT_OK;
}

int ground_launch(Ctx* c, Slot* s, cudaStream_t st, const float4* pts, int n, bool fuse_count) {
  s->cur_points = pts;
  s->cur_n = n;
  const int n_tiles = (n + kScanTile - 1) / kScanTile;
  if (n > 0) {
    const int bin_tiles = (n + kBinTile - 1) / kBinTile;
    const int bin_ctas = bin_tiles < c->bin_ctas ? bin_tiles : c->bin_ctas;     // persistent: <= 2 CTAs per SM
    polar_bin_kernel<<<bin_ctas, kBinTile, 0, st>>>(pts, n, c->gp, s->d_cell, s->d_polar_key);
    kernel_mark(c, s, st);
  }
  polar_grid_kernel<<<kGridCtas, kGridThreads, 0, st>>>(
      c->gp, s->d_polar_key, s->d_minz, s->d_height, s->d_smoothed, s->d_hdiff, s->d_hg, s->d_tile_desc, n_tiles, s->d_counters);
  kernel_mark(c, s, st);
  if (n > 0)
    classify_partition_kernel<<<n_tiles, kScanTile, 0, st>>>(pts, n, s->d_cell, s->d_hg, c->gp.tol, s->d_labels, s->d_elev,
                                                           s->d_ground, s->d_tile_desc, s->d_counters, c->prm.roi_m, s->d_cart,
                                                           fuse_count ? s->d_count : nullptr, s->d_polar_key);
  if (n > 0) kernel_mark(c, s, st);
  LMOT_CUDA(c, cudaGetLastError());
  return LMOT_OK;
}

}  // namespace lmot

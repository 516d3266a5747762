// cluster.cu -- 2-D occupancy-grid connected-component clustering on sm_100a.
//
// Replaces componentClustering (/root/reference/object_tracking/src/cluster/component_clustering.cpp:260-268):
//   mapCartesianGrid (:28-225)  count points per 0.2 m cell of the 250x250 grid over +-25 m; every cell holding
//                               MORE THAN ONE point and its (clipped) 3x3 neighbourhood becomes occupied (-1)
//   findComponent/search (:228-257)  raster scan (x outer, y inner); each unlabelled occupied cell starts
//                               id = ++numCluster and a recursive 8-connected flood fill
//
// B200 design: two launches per frame.
//   C1 cart_count_kernel   one thread per elevated point: fp32 binning exactly as :43-48, u16 cell id kept for
//                          box fitting, warp-aggregated atomicAdd (one atomic per distinct cell per warp).
//   C2 ccl_cluster_kernel  ONE thread-block cluster (8 CTAs x 1024 threads, cluster.sync between phases, the
//                          per-CTA root counts exchanged through distributed shared memory):
//                            seed = count>1 -> occupied = dilate3x3(seed) -> lock-free union-find with
//                            atomicMin (root = smallest linear index of the component) -> flatten ->
//                            id = 1 + rank of the root among all roots in linear (raster) order.
//                          The recursive fill labels components in the order the raster scan first meets
//                          them, which is exactly the order of their smallest linear index: identical ids.
#include <cooperative_groups.h>
#include "lmot_internal.cuh"
#include "exact_math.cuh"

namespace cg = cooperative_groups;

namespace lmot {

namespace {

constexpr int kCclCtas = 8;
constexpr int kCclThreads = 1024;
constexpr int kCclAll = kCclCtas * kCclThreads;                       // 8192 threads per frame
constexpr int kCclPerThread = 8;                                       // 32 rows x 250 cells / 1024 threads, rounded up

__device__ __forceinline__ unsigned cart_cell(float x, float y, float roi) { return cart_cell_of(x, y, roi, kNumGrid); }

__global__ void __launch_bounds__(256)
cart_count_kernel(const float4* __restrict__ elev, const int* __restrict__ counters, float roi,
                  uint16_t* __restrict__ cart, int* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = counters[CNT_N_ELEV];
  if (blockIdx.x * blockDim.x >= n) return;      // whole block past the end (launch is sized for the input cloud)
  unsigned c = kNoCell;
  if (i < n) {
    const float4 q = __ldg(&elev[i]);
    c = cart_cell(q.x, q.y, roi);
    cart[i] = (uint16_t)c;
  }
  const unsigned grp = __match_any_sync(0xFFFFFFFFu, c);
  if (c != kNoCell && (threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(&count[c], __popc(grp));
}

__global__ void __launch_bounds__(256)
cart_cells_kernel(const float4* __restrict__ elev, const int* __restrict__ counters, float roi, uint16_t* __restrict__ cart) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < counters[CNT_N_ELEV]) { const float4 q = __ldg(&elev[i]); cart[i] = (uint16_t)cart_cell(q.x, q.y, roi); }
}

// ---- union-find with root = smallest index (global-memory and shared-memory flavours) ----------------------
__device__ __forceinline__ int uf_find(volatile int* L, int x) {
  int p = L[x];
  while (p != x) { x = p; p = L[x]; }
  return x;
}

__device__ __forceinline__ void uf_union(volatile int* L, int* Lw, int a, int b) {
  while (true) {
    a = uf_find(L, a);
    b = uf_find(L, b);
    if (a == b) return;
    if (a > b) { const int t = a; a = b; b = t; }
    const int old = atomicMin(&Lw[b], a);
    if (old == b) return;
    b = old;
  }
}

constexpr int kTileRows = 32;                  // 8 CTAs x 32 rows cover the 250 rows of the grid
constexpr int kRowPitch = 256;                 // shared-memory row pitch (250 columns used)
constexpr int kSeedRows = kTileRows + 3;       // rows ts-2 .. ts+32
constexpr int kOccRows = kTileRows + 1;        // rows ts-1 .. ts+31
constexpr int kCclSmem = kSeedRows * kRowPitch + kOccRows * kRowPitch + 2 * kTileRows * kRowPitch * (int)sizeof(int);

// One thread-block cluster per frame; CTA r owns grid rows [32r, 32r+32).  Each CTA labels its tile entirely in shared
// memory (pointer chasing at ~30 cycles instead of ~600 through L2), only the 7 tile borders are merged through
// global memory, and the per-CTA root counts travel through distributed shared memory.
__global__ void __cluster_dims__(kCclCtas, 1, 1) __launch_bounds__(kCclThreads, 1)
ccl_cluster_kernel(int* __restrict__ count, int* G, int* rid, int* __restrict__ out, int* __restrict__ counters) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  extern __shared__ unsigned char ccl_smem[];
  uint8_t* s_seed = ccl_smem;                                         // [35][256] count > 1, rows ts-2..ts+32
  uint8_t* s_occ = s_seed + kSeedRows * kRowPitch;                    // [33][256] dilated,   rows ts-1..ts+31
  int* s_L = reinterpret_cast<int*>(s_occ + kOccRows * kRowPitch);    // [32][256] local parent (local index) or -1
  int* s_fin = s_L + kTileRows * kRowPitch;                           // [32][256] flattened local root
  __shared__ int s_warp[32];
  __shared__ int s_tot[kCclCtas];
  const int ts = crank * kTileRows;
  const int rows = min(kTileRows, kNumGrid - ts);                     // 32, last tile 26

  // A: seed = count > 1 (component_clustering.cpp:136) for the tile and its halo
  for (int l = tid; l < kSeedRows * kRowPitch; l += kCclThreads) {
    const int x = ts - 2 + (l >> 8), y = l & 255;
    uint8_t sd = 0;
    if (x >= 0 && x < kNumGrid && y < kNumGrid) sd = __ldcg(&count[x * kNumGrid + y]) > 1 ? 1 : 0;
    s_seed[l] = sd;
  }
  __syncthreads();
  // B: occupied = seed dilated 3x3, clipped at the border (:137-214), rows ts-1 .. ts+31
  for (int l = tid; l < kOccRows * kRowPitch; l += kCclThreads) {
    const int orow = l >> 8, y = l & 255;
    const int x = ts - 1 + orow;
    uint8_t occ = 0;
    if (x >= 0 && x < kNumGrid && y < kNumGrid) {
      const uint8_t* c = s_seed + (orow + 1) * kRowPitch + y;         // seed row of x
      const int ym = y > 0 ? -1 : 0, yp = y < kNumGrid - 1 ? 1 : 0;   // s_seed rows outside the grid are zero
      occ = c[ym] | c[0] | c[yp] | c[-kRowPitch + ym] | c[-kRowPitch] | c[-kRowPitch + yp] | c[kRowPitch + ym] | c[kRowPitch] |
            c[kRowPitch + yp];
    }
    s_occ[l] = occ;
  }
  __syncthreads();
  // C1: label = start of the horizontal run inside the 32-cell warp chunk
  for (int task = warp; task < kTileRows * 8; task += kCclThreads / 32) {
    const int lx = task >> 3, y = ((task & 7) << 5) + lane;
    const bool o = lx < rows && s_occ[(lx + 1) * kRowPitch + y];
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, o);
    const unsigned zb = ~mask & ((1u << lane) - 1u);                  // empty cells below this lane
    const int start = zb ? (32 - __clz(zb)) : 0;
    s_L[lx * kRowPitch + y] = o ? (lx * kRowPitch + (y - lane) + start) : -1;
  }
  __syncthreads();
  // C2/C3: runs continuing across chunk boundaries, and 8-connectivity to the row above inside the tile.  A cell only
  // needs the unions its left neighbour cannot have made: with N occupied, skip when W and NW are occupied too; with
  // N empty, NW only if W is empty, NE always.
  {
    volatile int* Lv = s_L;
    for (int l = tid; l < kTileRows * kRowPitch; l += kCclThreads) {
      const int lx = l >> 8, y = l & 255;
      if (lx >= rows || y >= kNumGrid || !s_occ[(lx + 1) * kRowPitch + y]) continue;
      const uint8_t* o = s_occ + (lx + 1) * kRowPitch + y;
      const bool W = y > 0 && o[-1];
      if ((y & 31) == 0 && W) uf_union(Lv, s_L, l, l - 1);
      if (lx > 0) {
        const bool N = o[-kRowPitch], NW = y > 0 && o[-kRowPitch - 1], NE = y < kNumGrid - 1 && o[-kRowPitch + 1];
        if (N) { if (!(W && NW)) uf_union(Lv, s_L, l, l - kRowPitch); }
        else {
          if (NW && !W) uf_union(Lv, s_L, l, l - kRowPitch - 1);
          if (NE) uf_union(Lv, s_L, l, l - kRowPitch + 1);
        }
      }
    }
  }
  __syncthreads();
  // D: flatten inside the tile (shared memory only).  Global memory holds a union-find node only where another CTA
  // can need one: at the tile-local ROOT cells and in the first / last row of the tile (the border merge looks the
  // neighbour's root up through them).  Everything else stays in shared memory -- pointer chasing through L2 costs
  // ~20x more per hop.
  {
    volatile int* Lv = s_L;
    for (int l = tid; l < kTileRows * kRowPitch; l += kCclThreads) {
      const int lx = l >> 8, y = l & 255;
      if (lx >= rows || y >= kNumGrid || Lv[l] < 0) continue;
      const int r = uf_find(Lv, l);
      s_fin[l] = r;                                                   // flattened local root (separate array: no races)
    }
  }
  __syncthreads();
  for (int l = tid; l < kTileRows * kRowPitch; l += kCclThreads) {
    const int lx = l >> 8, y = l & 255;
    if (lx >= rows || y >= kNumGrid || s_L[l] < 0) continue;
    const int r = s_fin[l];
    if (r == l || lx == 0 || lx == rows - 1) G[(ts + lx) * kNumGrid + y] = (ts + (r >> 8)) * kNumGrid + (r & 255);
  }
  cluster.sync();
  // every CTA has read its halo: zero this tile's counts for the next frame
  for (int l = tid; l < rows * kNumGrid; l += kCclThreads) count[ts * kNumGrid + l] = 0;
  // E: merge across the tile border (first row of the tile against the last row of the previous tile), same rule
  if (crank > 0 && tid < kNumGrid) {
    const int y = tid;
    const uint8_t* o = s_occ + 1 * kRowPitch + y;                     // row ts
    if (o[0]) {
      volatile int* Gv = G;
      const int k = ts * kNumGrid + y;
      const bool W = y > 0 && o[-1];
      const bool N = o[-kRowPitch], NW = y > 0 && o[-kRowPitch - 1], NE = y < kNumGrid - 1 && o[-kRowPitch + 1];
      if (N) { if (!(W && NW)) uf_union(Gv, G, k, k - kNumGrid); }
      else {
        if (NW && !W) uf_union(Gv, G, k, k - kNumGrid - 1);
        if (NE) uf_union(Gv, G, k, k - kNumGrid + 1);
      }
    }
  }
  cluster.sync();
  // F: only the tile-local roots ask global memory for their final root; every other cell reads it from its root
  // through shared memory
  for (int l = tid; l < kTileRows * kRowPitch; l += kCclThreads) {
    const int lx = l >> 8, y = l & 255;
    if (lx >= rows || y >= kNumGrid || s_L[l] < 0 || s_fin[l] != l) continue;
    volatile int* Gv = G;
    s_L[l] = uf_find(Gv, (ts + lx) * kNumGrid + y);                   // s_L of a local root now holds the FINAL global root
  }
  __syncthreads();
  // roots of this tile in linear order, thread t owns kCclPerThread consecutive cells
  const int cbeg = ts * kNumGrid, cend = cbeg + rows * kNumGrid;
  const int tbeg = cbeg + tid * kCclPerThread;
  int roots = 0;
  unsigned rootmask = 0;
  int fin[kCclPerThread];
#pragma unroll
  for (int j = 0; j < kCclPerThread; ++j) {
    const int k = tbeg + j;
    fin[j] = -1;
    if (k < cend) {
      const int off = k - cbeg, l = (off / kNumGrid) * kRowPitch + off % kNumGrid;
      if (s_L[l] >= 0) {
        fin[j] = s_L[s_fin[l]];                                       // final global root of the cell
        if (fin[j] == k) { ++roots; rootmask |= 1u << j; }
      }
    }
  }
  int incl = roots;
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
    if (lane == 31)       // publish this CTA's root count into every CTA's s_tot through distributed shared memory
      for (int r = 0; r < kCclCtas; ++r) cluster.map_shared_rank(s_tot, r)[crank] = wi;
  }
  __syncthreads();
  const int excl_in_cta = s_warp[warp] + incl - roots;
  cluster.sync();
  int base = 0, total = 0;
#pragma unroll
  for (int r = 0; r < kCclCtas; ++r) { if (r < crank) base += s_tot[r]; total += s_tot[r]; }
  {
    int rank = base + excl_in_cta;
#pragma unroll
    for (int j = 0; j < kCclPerThread; ++j)
      if (rootmask & (1u << j)) rid[tbeg + j] = ++rank;      // id = 1 + rank in raster order (:247-257)
  }
  if (crank == 0 && tid == 0) counters[CNT_NUM_CLUSTER] = total;
  cluster.sync();
  // G: label grid (the ids of roots that live in other tiles come through L2; the loads are independent)
  int ids[kCclPerThread];
#pragma unroll
  for (int j = 0; j < kCclPerThread; ++j) ids[j] = fin[j] >= 0 ? __ldcg(&rid[fin[j]]) : 0;
#pragma unroll
  for (int j = 0; j < kCclPerThread; ++j) if (tbeg + j < cend) out[tbeg + j] = ids[j];
}

}  // namespace

int cluster_alloc(Ctx* c, Slot* s) {
  const size_t np = (size_t)c->max_points;
  LMOT_CUDA(c, cudaMalloc(&s->d_cart, np * sizeof(uint16_t)));
  LMOT_CUDA(c, cudaMalloc(&s->d_count, kCartCells * sizeof(int)));
  LMOT_CUDA(c, cudaMalloc(&s->d_parent, kCartCells * sizeof(int)));
  LMOT_CUDA(c, cudaMalloc(&s->d_rid, kCartCells * sizeof(int)));
  LMOT_CUDA(c, cudaMalloc(&s->d_label_grid, kCartCells * sizeof(int)));
  LMOT_CUDA(c, cudaMemsetAsync(s->d_count, 0, kCartCells * sizeof(int), s->stream));
  LMOT_CUDA(c, cudaMemsetAsync(s->d_label_grid, 0, kCartCells * sizeof(int), s->stream));
  LMOT_CUDA(c, cudaFuncSetAttribute(ccl_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCclSmem));
  return LMOT_OK;
}

void cluster_free(Slot* s) {
  cudaFree(s->d_cart); cudaFree(s->d_count); cudaFree(s->d_parent); cudaFree(s->d_rid); cudaFree(s->d_label_grid);
}

// elevated cloud = s->d_elev with its length in d_counters[CNT_N_ELEV]; n_upper bounds that length on the host
int cluster_launch(Ctx* c, Slot* s, cudaStream_t st, int n_upper, bool counted) {
  if (n_upper > 0 && !counted)   // `counted`: classify_partition_kernel already binned the elevated points (fused frame path)
    cart_count_kernel<<<(n_upper + 255) / 256, 256, 0, st>>>(s->d_elev, s->d_counters, c->prm.roi_m, s->d_cart, s->d_count);
  if (n_upper > 0 && !counted) kernel_mark(c, s, st);
  ccl_cluster_kernel<<<kCclCtas, kCclThreads, kCclSmem, st>>>(s->d_count, s->d_parent, s->d_rid, s->d_label_grid, s->d_counters);
  kernel_mark(c, s, st);
  LMOT_CUDA(c, cudaGetLastError());
  return LMOT_OK;
}

int cluster_cells_only(Ctx* c, Slot* s, cudaStream_t st, int n_upper) {
  if (n_upper > 0)
    cart_cells_kernel<<<(n_upper + 255) / 256, 256, 0, st>>>(s->d_elev, s->d_counters, c->prm.roi_m, s->d_cart);
  LMOT_CUDA(c, cudaGetLastError());
  return LMOT_OK;
}

}  // namespace lmot

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
constexpr int kCclChunk = (kCartCells + kCclCtas - 1) / kCclCtas;     // 7813 cells per CTA (contiguous)
constexpr int kCclPerThread = (kCclChunk + kCclThreads - 1) / kCclThreads;  // 8

// component_clustering.cpp:40-48
__device__ __forceinline__ unsigned cart_cell(float x, float y, float roi) {
  const float half = fdiv(roi, 2.f);
  const float xC = fadd(x, half), yC = fadd(y, half);
  if (!(xC >= 0.f && xC < roi && yC >= 0.f && yC < roi)) return kNoCell;
  const int xI = (int)floorf(fdiv(fmul((float)kNumGrid, xC), roi));
  const int yI = (int)floorf(fdiv(fmul((float)kNumGrid, yC), roi));
  return (unsigned)(xI * kNumGrid + yI);
}

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

__global__ void __cluster_dims__(kCclCtas, 1, 1) __launch_bounds__(kCclThreads, 1)
ccl_cluster_kernel(int* __restrict__ count, uint8_t* __restrict__ seed, int* L, int* rid, int* __restrict__ out,
                   int* __restrict__ counters) {
  cg::cluster_group cluster = cg::this_cluster();
  const int crank = (int)cluster.block_rank();
  const int tid = threadIdx.x;
  const int gtid = crank * kCclThreads + tid;
  __shared__ int s_warp[32];
  __shared__ int s_tot[kCclCtas];

  // P0: seed = count > 1 (component_clustering.cpp:136); the count grid is zeroed for the next frame
  for (int k = gtid; k < kCartCells; k += kCclAll) {
    seed[k] = count[k] > 1 ? 1 : 0;
    count[k] = 0;
  }
  cluster.sync();

  // P1: occupied = seed dilated 3x3, clipped at the border (:137-214); parent = self
  for (int k = gtid; k < kCartCells; k += kCclAll) {
    const int x = k / kNumGrid, y = k % kNumGrid;
    int occ = 0;
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int xx = x + dx;
      if (xx < 0 || xx >= kNumGrid) continue;
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= kNumGrid) continue;
        occ |= __ldcg(&seed[xx * kNumGrid + yy]);
      }
    }
    L[k] = occ ? k : -1;
  }
  cluster.sync();

  // P2: 8-connectivity (:228-244): union with the four neighbours that precede the cell in raster order
  {
    volatile int* Lv = L;
    for (int k = gtid; k < kCartCells; k += kCclAll) {
      if (Lv[k] < 0) continue;
      const int x = k / kNumGrid, y = k % kNumGrid;
      if (y > 0 && Lv[k - 1] >= 0) uf_union(Lv, L, k, k - 1);
      if (x > 0) {
        const int u = k - kNumGrid;
        if (Lv[u] >= 0) uf_union(Lv, L, k, u);
        else {
          // (x-1,y-1) and (x-1,y+1) are both adjacent to (x-1,y) when that one is occupied
          if (y > 0 && Lv[u - 1] >= 0) uf_union(Lv, L, k, u - 1);
          if (y < kNumGrid - 1 && Lv[u + 1] >= 0) uf_union(Lv, L, k, u + 1);
        }
      }
    }
  }
  cluster.sync();

  // P3: flatten; count roots of this CTA's contiguous chunk, thread t owns kCclPerThread consecutive cells
  const int cbeg = crank * kCclChunk, cend = min(cbeg + kCclChunk, kCartCells);
  const int tbeg = cbeg + tid * kCclPerThread;
  int roots = 0;
  unsigned rootmask = 0;
  {
    volatile int* Lv = L;
#pragma unroll
    for (int j = 0; j < kCclPerThread; ++j) {
      const int k = tbeg + j;
      if (k < cend && Lv[k] >= 0) {
        const int r = uf_find(Lv, k);
        Lv[k] = r;
        if (r == k) { ++roots; rootmask |= 1u << j; }
      }
    }
  }
  // block exclusive scan of `roots`
  int incl = roots;
  const int lane = tid & 31, warp = tid >> 5;
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
    if (lane == 31) {
      // publish this CTA's root count into every CTA's s_tot through distributed shared memory
      for (int r = 0; r < kCclCtas; ++r) cluster.map_shared_rank(s_tot, r)[crank] = wi;
    }
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
  if (gtid == 0) counters[CNT_NUM_CLUSTER] = total;
  cluster.sync();

  // P4: label grid
  for (int k = gtid; k < kCartCells; k += kCclAll) {
    const int r = __ldcg(&L[k]);
    out[k] = r >= 0 ? __ldcg(&rid[r]) : 0;
  }
}

}  // namespace

int cluster_alloc(Ctx* c) {
  const size_t np = (size_t)c->max_points;
  LMOT_CUDA(c, cudaMalloc(&c->d_cart, np * sizeof(uint16_t)));
  LMOT_CUDA(c, cudaMalloc(&c->d_count, kCartCells * sizeof(int)));
  LMOT_CUDA(c, cudaMalloc(&c->d_seed, kCartCells));
  LMOT_CUDA(c, cudaMalloc(&c->d_parent, kCartCells * sizeof(int)));
  LMOT_CUDA(c, cudaMalloc(&c->d_rid, kCartCells * sizeof(int)));
  LMOT_CUDA(c, cudaMalloc(&c->d_label_grid, kCartCells * sizeof(int)));
  LMOT_CUDA(c, cudaMemsetAsync(c->d_count, 0, kCartCells * sizeof(int), c->stream));
  LMOT_CUDA(c, cudaMemsetAsync(c->d_label_grid, 0, kCartCells * sizeof(int), c->stream));
  return LMOT_OK;
}

void cluster_free(Ctx* c) {
  cudaFree(c->d_cart); cudaFree(c->d_count); cudaFree(c->d_seed); cudaFree(c->d_parent); cudaFree(c->d_rid);
  cudaFree(c->d_label_grid);
}

// elevated cloud = c->d_elev with its length in d_counters[CNT_N_ELEV]; n_upper bounds that length on the host
int cluster_launch(Ctx* c, int n_upper) {
  if (n_upper > 0)
    cart_count_kernel<<<(n_upper + 255) / 256, 256, 0, c->stream>>>(c->d_elev, c->d_counters, c->prm.roi_m, c->d_cart,
                                                                    c->d_count);
  ccl_cluster_kernel<<<kCclCtas, kCclThreads, 0, c->stream>>>(c->d_count, c->d_seed, c->d_parent, c->d_rid,
                                                              c->d_label_grid, c->d_counters);
  LMOT_CUDA(c, cudaGetLastError());
  return LMOT_OK;
}

int cluster_cells_only(Ctx* c, int n_upper) {
  if (n_upper > 0)
    cart_cells_kernel<<<(n_upper + 255) / 256, 256, 0, c->stream>>>(c->d_elev, c->d_counters, c->prm.roi_m, c->d_cart);
  LMOT_CUDA(c, cudaGetLastError());
  return LMOT_OK;
}

}  // namespace lmot

// cluster.cu -- 2-D occupancy-grid connected-component clustering on sm_100a.
//
// Replaces componentClustering (/root/reference/object_tracking/src/cluster/component_clustering.cpp:260-268):
//   mapCartesianGrid (:28-225)  count points per 0.2 m cell of the 250x250 grid over +-25 m; every cell holding
//                               MORE THAN ONE point and its (clipped) 3x3 neighbourhood becomes occupied (-1)
//   findComponent/search (:228-257)  raster scan (x outer, y inner); each unlabelled occupied cell starts
//                               id = ++numCluster and a recursive 8-connected flood fill
//
// B200 design: the grid is 250 rows of 250 BITS (8 words per row, 8 KB in all), not 62,500 counters.
//   * "count > 1" needs no counter: two bit planes, `once` and `twice`.  A point sets its cell's bit in `once`; if it
//     was set already (or two points of one warp share the cell) it sets the bit in `twice`.  seed == twice.
//     (cart_mark(), called by ground_fused_kernel for every elevated point while it is still in registers, or by
//     cart_mark_kernel for the stand-alone entry point.)
//   * ccl_bitmap_kernel, ONE CTA, everything in shared memory:
//       dilate3x3 on words  ->  "pieces" (maximal runs of 1-bits inside a word; <= 16 per word, node id = word*16 + k,
//       monotone in the raster index of the piece's first cell)  ->  union-find over pieces (atomicMin, root = smallest
//       id): a piece joins the previous word's last piece when the run continues across the word boundary, and every
//       piece of the row above that touches its 1-cell halo  ->  path flattening  ->  id = 1 + rank of the root among
//       the roots in node order.  The recursive fill labels components in the order the raster scan first meets them,
//       i.e. by their smallest linear index == smallest node id: identical ids.
//       The 250x250 int label grid (the reference's only carrier of labels) is updated SPARSELY: cells occupied in the
//       previous frame are cleared, cells occupied now get their id (stores only where something is or was).
#include <climits>
#include "lmot_internal.cuh"
#include "exact_math.cuh"

namespace lmot {

namespace {

constexpr int kRowWords = 8;                                   // 250 bits per row
constexpr int kBitWords = kNumGrid * kRowWords;                // 2000
constexpr unsigned kLastWordMask = (1u << (kNumGrid - 32 * (kRowWords - 1))) - 1u;   // bits 0..25 of word 7
constexpr int kPiecesPerWord = 16;
constexpr int kNodes = kBitWords * kPiecesPerWord;             // 32000
constexpr int kCclThreads = 1024;
constexpr int kCclSmem = kNodes * 4 + 3 * kBitWords * 4 + 64 * 4;

__global__ void __launch_bounds__(256)
cart_mark_kernel(const float4* __restrict__ elev, const int* __restrict__ counters, float roi, uint16_t* __restrict__ cart,
                 unsigned* __restrict__ once, unsigned* __restrict__ twice) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = counters[CNT_N_ELEV];
  if (blockIdx.x * blockDim.x >= n) return;      // whole block past the end (launch is sized for the input cloud)
  unsigned c = kNoCell;
  if (i < n) {
    const float4 q = __ldg(&elev[i]);
    c = cart_cell_of(q.x, q.y, roi, kNumGrid);
    cart[i] = (uint16_t)c;
  }
  cart_mark(c, once, twice);
}

__global__ void __launch_bounds__(256)
cart_cells_kernel(const float4* __restrict__ elev, const int* __restrict__ counters, float roi, uint16_t* __restrict__ cart) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < counters[CNT_N_ELEV]) { const float4 q = __ldg(&elev[i]); cart[i] = (uint16_t)cart_cell_of(q.x, q.y, roi, kNumGrid); }
}

// ---- union-find over shared-memory nodes, root = smallest id -------------------------------------------------
// find with path halving: every visited node is re-pointed at its grandparent.  Parents only ever move to an ancestor
// (a smaller id), so concurrent finds / unions by other threads stay correct, and the chains that the row-by-row links
// would otherwise build (one hop per grid row of a tall component) collapse while the unions are still being made.
__device__ __forceinline__ int uf_find(volatile int* L, int x) {
  int p = L[x];
  while (p != x) {
    const int g = L[p];
    if (g != p) L[x] = g;
    x = p; p = g;
  }
  return x;
}

// read-only find for the flattening pass: there, the only writes are final roots (a halving store by another thread could
// land after the owner's flattening store and point the node back at an intermediate ancestor)
__device__ __forceinline__ int uf_root(const volatile int* L, int x) {
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

// horizontal 3-tap dilation of one word of a row (neighbour words supply the carry bits)
__device__ __forceinline__ unsigned hdil(const unsigned* row, int k) {
  const unsigned s = row[k];
  unsigned d = s | (s << 1) | (s >> 1);
  if (k > 0) d |= row[k - 1] >> 31;
  if (k < kRowWords - 1) d |= row[k + 1] << 31;
  return d;
}

__device__ __forceinline__ unsigned piece_starts(unsigned m) { return m & ~(m << 1); }
// index (inside its word) of the piece that contains set bit p of word m
__device__ __forceinline__ int piece_of(unsigned m, int p) { return __popc(piece_starts(m) & ((2u << p) - 1u)) - 1; }

// One frame of a launch (blockIdx.x): bit planes in, label grid + cluster count out
struct CclFrame {
  unsigned* once; unsigned* twice; unsigned* prev_occ;
  int* out;
  int* counters;
};
struct CclBatch { CclFrame f[kMaxBatch]; };

// pointer jumping: re-point x at its grandparent until its parent is a root.  Only the owner of x stores to L[x] here and every
// store moves x to an ancestor, so the walks of all threads run concurrently and shorten each other: a chain of n pieces
// (a tall object: one link per grid row) collapses in ~log2(n) rounds instead of n dependent loads per thread.
__device__ __forceinline__ int uf_compress(volatile int* L, int x) {
  int p = L[x];
  while (true) {
    const int g = L[p];
    if (g == p) break;
    L[x] = g;
    p = g;
  }
  return p;
}

__device__ __forceinline__ void ccl_mark(unsigned long long* clk, int slot) {
  if (clk && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); clk[blockIdx.x * 16 + slot] = t; }
}

// ONE CTA per frame, everything in shared memory:
//   A  seed = `twice` plane; re-arm the planes          B  occupied = seed dilated 3x3
//   C  one node per piece; its first parent is the SMALLEST neighbour it touches: the leftmost touching piece of the row above,
//      else the piece it continues from the previous word, else itself (plain stores, no atomics: a forest, ids decrease upwards)
//   D  pointer jumping                                  E  the remaining adjacencies (a piece touching several pieces above, or
//      one above and one to the left) as union-find unions over the now flat forest
//   F  pointer jumping again                            G  id = 1 + rank of the root in raster order, sparse label-grid update
__global__ void __launch_bounds__(kCclThreads, 1)
ccl_bitmap_kernel(const __grid_constant__ CclBatch B, unsigned long long* __restrict__ clk) {
  extern __shared__ __align__(16) unsigned char ccl_smem[];
  int* s_par = reinterpret_cast<int*>(ccl_smem);                       // [32000] parent node, later -(cluster id) at roots
  unsigned* s_seed = reinterpret_cast<unsigned*>(s_par + kNodes);       // [2000]
  unsigned* s_occ = s_seed + kBitWords;                                 // [2000]
  unsigned* s_prev = s_occ + kBitWords;                                 // [2000]
  int* s_warp = reinterpret_cast<int*>(s_prev + kBitWords);             // [32] + total
  const CclFrame& F = B.f[blockIdx.x];
  unsigned* __restrict__ once = F.once; unsigned* __restrict__ twice = F.twice; unsigned* __restrict__ prev_occ = F.prev_occ;
  int* __restrict__ out = F.out;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ccl_mark(clk, 0);

  // Thread t owns the two consecutive words 2t, 2t+1 (t < 1000) in every step.
  const int w0 = 2 * tid;
  const bool own = tid < kBitWords / 2;
  // A: seed = cells with more than one point (component_clustering.cpp:136); the bit planes are re-armed for the next frame
  if (own) {
    const uint2 tw = __ldcg(reinterpret_cast<const uint2*>(twice) + tid);
    const uint2 pv = __ldcg(reinterpret_cast<const uint2*>(prev_occ) + tid);
    s_seed[w0] = tw.x; s_seed[w0 + 1] = tw.y;
    s_prev[w0] = pv.x; s_prev[w0 + 1] = pv.y;
    reinterpret_cast<uint2*>(once)[tid] = make_uint2(0u, 0u);
    reinterpret_cast<uint2*>(twice)[tid] = make_uint2(0u, 0u);
  }
  __syncthreads();
  ccl_mark(clk, 1);
  // B: occupied = seed dilated 3x3, clipped at the border (:137-214)
  unsigned occ[2] = {0u, 0u};
  if (own) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int w = w0 + h, x = w >> 3, k = w & 7;
      unsigned o = hdil(s_seed + x * kRowWords, k);
      if (x > 0) o |= hdil(s_seed + (x - 1) * kRowWords, k);
      if (x < kNumGrid - 1) o |= hdil(s_seed + (x + 1) * kRowWords, k);
      if (k == kRowWords - 1) o &= kLastWordMask;
      occ[h] = o;
      s_occ[w] = o;
    }
    reinterpret_cast<uint2*>(prev_occ)[tid] = make_uint2(occ[0], occ[1]);
  }
  __syncthreads();
  ccl_mark(clk, 2);
  // C: one node per piece, first parent = smallest touching neighbour.  `extra` (per word, one bit per piece): the piece has
  // further adjacencies that step E must union.
  unsigned extra[2] = {0u, 0u};
  if (own) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const unsigned m = occ[h];
      if (!m) continue;
      const int w = w0 + h, x = w >> 3, k = w & 7;
      unsigned up = 0u, upl = 0u, upr = 0u;
      if (x > 0) {
        up = s_occ[w - kRowWords];
        upl = k > 0 ? s_occ[w - kRowWords - 1] : 0u;
        upr = k < kRowWords - 1 ? s_occ[w - kRowWords + 1] : 0u;
      }
      const unsigned left = (k > 0) ? ((h == 1) ? occ[0] : s_occ[w - 1]) : 0u;
      unsigned rest = m;
      int j = 0;
      while (rest) {
        const int a = __ffs(rest) - 1;
        const unsigned t = ~(rest >> a);                       // first zero above a ends the piece
        const int len = t ? __ffs(t) - 1 : 32;
        const unsigned pm = (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << a;
        rest &= ~pm;
        const int me = w * kPiecesPerWord + j;
        int par = me, links = 0;
        // same row: the run continues from the previous word (its last piece)
        const bool cont = (a == 0) && (left >> 31);
        // row above, right to left so that the LAST assignment is the leftmost (smallest id): word to the right (its first
        // piece), the pieces of the word above, the word to the left (its last piece)
        if (a + len == 32 && (upr & 1u)) { par = (w - kRowWords + 1) * kPiecesPerWord; ++links; }
        const unsigned touched = up & (pm | (pm << 1) | (pm >> 1));
        if (touched) {
          par = (w - kRowWords) * kPiecesPerWord + piece_of(up, __ffs(touched) - 1);
          // number of distinct pieces of `up` under `touched`: starts inside it, plus one if its lowest bit continues a piece
          const unsigned st = piece_starts(up) & touched;
          links += __popc(st) + (((touched & (0u - touched)) & ~piece_starts(up)) ? 1 : 0);
        }
        if (a == 0 && (upl >> 31)) {
          // the last piece of the word above-left; if it runs on into bit 0 of `up` it IS the first touched piece (same node
          // reached through its continuation link): not a further adjacency
          par = (w - kRowWords - 1) * kPiecesPerWord + __popc(piece_starts(upl)) - 1;
          ++links;
        }
        if (cont) { if (links == 0) par = (w - 1) * kPiecesPerWord + __popc(piece_starts(left)) - 1; ++links; }
        s_par[me] = par;
        if (links > 1) extra[h] |= 1u << j;
        ++j;
      }
    }
  }
  __syncthreads();
  ccl_mark(clk, 3);
  // D: flatten the forest
  if (own) {
    volatile int* Lv = s_par;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int np = __popc(piece_starts(occ[h]));
      for (int j = 0; j < np; ++j) uf_compress(Lv, (w0 + h) * kPiecesPerWord + j);
    }
  }
  __syncthreads();
  ccl_mark(clk, 4);
  // E: the adjacencies step C did not use (all of them, for the few pieces that have more than one)
  if (own) {
    volatile int* Lv = s_par;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (!extra[h]) continue;
      const unsigned m = occ[h];
      const int w = w0 + h, x = w >> 3, k = w & 7;
      unsigned up = 0u, upl = 0u, upr = 0u;
      if (x > 0) {
        up = s_occ[w - kRowWords];
        upl = k > 0 ? s_occ[w - kRowWords - 1] : 0u;
        upr = k < kRowWords - 1 ? s_occ[w - kRowWords + 1] : 0u;
      }
      const unsigned left = (k > 0) ? ((h == 1) ? occ[0] : s_occ[w - 1]) : 0u;
      unsigned rest = m;
      int j = 0;
      while (rest) {
        const int a = __ffs(rest) - 1;
        const unsigned t = ~(rest >> a);
        const int len = t ? __ffs(t) - 1 : 32;
        const unsigned pm = (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << a;
        rest &= ~pm;
        const int me = w * kPiecesPerWord + j;
        if ((extra[h] >> j) & 1u) {
          unsigned touched = up & (pm | (pm << 1) | (pm >> 1));
          while (touched) {
            const int p = __ffs(touched) - 1;
            uf_union(Lv, s_par, me, (w - kRowWords) * kPiecesPerWord + piece_of(up, p));
            const unsigned tu = ~(up >> p);
            const int lu = tu ? __ffs(tu) - 1 : 32;
            touched &= ~((lu >= 32 ? 0xFFFFFFFFu : ((1u << lu) - 1u)) << p);
          }
          if (a == 0 && (upl >> 31)) uf_union(Lv, s_par, me, (w - kRowWords - 1) * kPiecesPerWord + __popc(piece_starts(upl)) - 1);
          if (a + len == 32 && (upr & 1u)) uf_union(Lv, s_par, me, (w - kRowWords + 1) * kPiecesPerWord);
          if (a == 0 && (left >> 31)) uf_union(Lv, s_par, me, (w - 1) * kPiecesPerWord + __popc(piece_starts(left)) - 1);
        }
        ++j;
      }
    }
  }
  __syncthreads();
  ccl_mark(clk, 5);
  // F: flatten again (the unions re-pointed some roots), count the roots of the thread's words
  int roots = 0;
  unsigned rootmask[2] = {0u, 0u};
  if (own) {
    volatile int* Lv = s_par;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const unsigned m = occ[h];
      if (!m) continue;
      const int np = __popc(piece_starts(m));
      for (int j = 0; j < np; ++j) {
        const int nd = (w0 + h) * kPiecesPerWord + j;
        if (Lv[nd] == nd) { ++roots; rootmask[h] |= 1u << j; }     // (roots stay roots: nobody unions any more)
        else uf_compress(Lv, nd);
      }
    }
  }
  // id = 1 + rank of the root among all roots in node (= raster) order (:247-257)
  int incl = roots;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  ccl_mark(clk, 6);
  if (warp == 0) {
    const int v = s_warp[lane];
    int wi = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xFFFFFFFFu, wi, o); if (lane >= o) wi += t; }
    s_warp[lane] = wi - v;
    if (lane == 31) F.counters[CNT_NUM_CLUSTER] = wi;
  }
  __syncthreads();
  {
    int rank = s_warp[warp] + incl - roots;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      unsigned rm = rootmask[h];
      while (rm) { const int j = __ffs(rm) - 1; rm &= rm - 1; s_par[(w0 + h) * kPiecesPerWord + j] = -(++rank); }
    }
  }
  __syncthreads();
  ccl_mark(clk, 7);
  // G: label grid, sparse: cells occupied now get their id, cells occupied only in the previous frame are cleared
  if (own) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int w = w0 + h;
      const unsigned m = occ[h];
      int* o = out + (w >> 3) * kNumGrid + (w & 7) * 32;
      unsigned gone = s_prev[w] & ~m;
      while (gone) { const int b = __ffs(gone) - 1; gone &= gone - 1; o[b] = 0; }
      unsigned rest = m;
      int j = 0;
      while (rest) {
        const int a = __ffs(rest) - 1;
        const unsigned t = ~(rest >> a);
        const int len = t ? __ffs(t) - 1 : 32;
        rest &= ~((len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << a);
        int r = s_par[w * kPiecesPerWord + j];
        ++j;
        if (r >= 0) r = s_par[r];                               // non-root: its root holds -(id)
        for (int b = a; b < a + len; ++b) o[b] = -r;
      }
    }
  }
  ccl_mark(clk, 8);
}

// ---------------------------------------------------------------------------------------------------------------
// ccl_dense_kernel: the same algorithm on DENSE node ids.  ccl_bitmap_kernel numbers a piece word * 16 + k: the 32 lanes of a
// warp then hit the same shared-memory bank on every parent access, each thread walks its own ragged list of pieces, and the
// flattening walks are data-dependent loops of dependent loads (measured 18-23 us per frame, 8 of them in the second flatten).
// Here a block-wide scan of the pieces per word gives every piece its raster-order rank as id (still monotone: the smallest id of
// a component is its first piece in raster order), so that
//   * flattening is pointer jumping with thread t on ids t, t + 1024, ...: conflict-free, ~log2(depth) steps per node;
//   * the adjacencies step C cannot express as a parent link go to a pair list and are united one pair per thread.
constexpr int kPairCap = 4096;
constexpr int kDenseSmem = kNodes * 4 + 3 * kBitWords * 4 + (kBitWords + 8) * 4 + kPairCap * 4 + 64 * 4;

// flatten: every node re-points itself at its grandparent until its parent is a root (uf_compress), thread t on ids t, t + 1024,
// ...; only the owner stores to L[id], every store moves id to an ancestor, roots do not change while this runs -- so the walks
// need no barrier between them and shorten each other: ~log2(depth) steps per node.  (Barrier-synchronised rounds with
// __syncthreads_or were measured at 0.64 us PER ROUND, 4.5 us for the 7 rounds of a depth-54 forest.)
__device__ __forceinline__ void ccl_flatten(volatile int* L, int P) {
  for (int id = threadIdx.x; id < P; id += kCclThreads) uf_compress(L, id);
  __syncthreads();
}

__global__ void __launch_bounds__(kCclThreads, 1)
ccl_dense_kernel(const __grid_constant__ CclBatch B, unsigned long long* __restrict__ clk) {
  extern __shared__ __align__(16) unsigned char ccl_smem[];
  int* s_par = reinterpret_cast<int*>(ccl_smem);                       // [<= 32000] parent id, later -(cluster id) at roots
  unsigned* s_seed = reinterpret_cast<unsigned*>(s_par + kNodes);       // [2000]
  unsigned* s_occ = s_seed + kBitWords;                                 // [2000]
  unsigned* s_prev = s_occ + kBitWords;                                 // [2000]
  int* s_wbase = reinterpret_cast<int*>(s_prev + kBitWords);            // [2000 + 1] id of the first piece of every word
  unsigned* s_pairs = reinterpret_cast<unsigned*>(s_wbase + kBitWords + 8);   // [kPairCap] a | b << 16
  int* s_warp = reinterpret_cast<int*>(s_pairs + kPairCap);             // [32] + total, [40] pair count
  const CclFrame& F = B.f[blockIdx.x];
  unsigned* __restrict__ once = F.once; unsigned* __restrict__ twice = F.twice; unsigned* __restrict__ prev_occ = F.prev_occ;
  int* __restrict__ out = F.out;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  ccl_mark(clk, 0);
  const int w0 = 2 * tid;
  const bool own = tid < kBitWords / 2;
  // A: seed = cells with more than one point (component_clustering.cpp:136); the bit planes are re-armed for the next frame
  if (own) {
    const uint2 tw = __ldcg(reinterpret_cast<const uint2*>(twice) + tid);
    const uint2 pv = __ldcg(reinterpret_cast<const uint2*>(prev_occ) + tid);
    s_seed[w0] = tw.x; s_seed[w0 + 1] = tw.y;
    s_prev[w0] = pv.x; s_prev[w0 + 1] = pv.y;
    reinterpret_cast<uint2*>(once)[tid] = make_uint2(0u, 0u);
    reinterpret_cast<uint2*>(twice)[tid] = make_uint2(0u, 0u);
  }
  if (tid == 0) s_warp[40] = 0;
  __syncthreads();
  ccl_mark(clk, 1);
  // B: occupied = seed dilated 3x3, clipped at the border (:137-214); ids: exclusive scan of the pieces per word
  unsigned occ[2] = {0u, 0u};
  int np0 = 0, np1 = 0;
  if (own) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int w = w0 + h, x = w >> 3, k = w & 7;
      unsigned o = hdil(s_seed + x * kRowWords, k);
      if (x > 0) o |= hdil(s_seed + (x - 1) * kRowWords, k);
      if (x < kNumGrid - 1) o |= hdil(s_seed + (x + 1) * kRowWords, k);
      if (k == kRowWords - 1) o &= kLastWordMask;
      occ[h] = o;
      s_occ[w] = o;
    }
    reinterpret_cast<uint2*>(prev_occ)[tid] = make_uint2(occ[0], occ[1]);
    np0 = __popc(piece_starts(occ[0])); np1 = __popc(piece_starts(occ[1]));
  }
  int P;
  {
    const int cnt = np0 + np1;
    int incl = cnt;
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
      if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    const int base = s_warp[warp] + incl - cnt;
    if (own) { s_wbase[w0] = base; s_wbase[w0 + 1] = base + np0; }
    P = s_warp[32];
  }
  __syncthreads();
  ccl_mark(clk, 2);
  // C: first parent = the smallest neighbour a piece touches (leftmost touching piece of the row above, else the piece it
  // continues from the previous word, else itself); every OTHER adjacency becomes a pair for step E
  if (own) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const unsigned m = occ[h];
      if (!m) continue;
      const int w = w0 + h, x = w >> 3, k = w & 7;
      unsigned up = 0u, upl = 0u, upr = 0u;
      int bu = 0, bul = 0, bur = 0;
      if (x > 0) {
        up = s_occ[w - kRowWords]; bu = s_wbase[w - kRowWords];
        if (k > 0) { upl = s_occ[w - kRowWords - 1]; bul = s_wbase[w - kRowWords - 1]; }
        if (k < kRowWords - 1) { upr = s_occ[w - kRowWords + 1]; bur = s_wbase[w - kRowWords + 1]; }
      }
      const unsigned left = (k > 0) ? ((h == 1) ? occ[0] : s_occ[w - 1]) : 0u;
      const int bl = (k > 0) ? s_wbase[w - 1] : 0;
      const int mybase = s_wbase[w];
      unsigned rest = m;
      int j = 0;
      while (rest) {
        const int a = __ffs(rest) - 1;
        const unsigned t = ~(rest >> a);                       // first zero above a ends the piece
        const int len = t ? __ffs(t) - 1 : 32;
        const unsigned pm = (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << a;
        rest &= ~pm;
        const int me = mybase + j;
        ++j;
        // neighbours in DEcreasing id order; the last one found is the parent, the others go to the pair list
        int par = me;
        auto link = [&](int nb) {
          if (par != me) {                                     // the previous candidate loses: remember it as a pair
            const int slot = atomicAdd(&s_warp[40], 1);
            if (slot < kPairCap) s_pairs[slot] = (unsigned)me | ((unsigned)par << 16);   // (list full: step E re-derives ALL adjacencies instead)
          }
          par = nb;
        };
        if (a == 0 && (left >> 31)) link(bl + __popc(piece_starts(left)) - 1);                 // same row, previous word
        if (a + len == 32 && (upr & 1u)) link(bur);                                             // row above, word to the right
        unsigned touched = up & (pm | (pm << 1) | (pm >> 1));                                   // row above, same word: right to left
        while (touched) {
          const int p = 31 - __clz(touched);
          const int q = piece_of(up, p);
          link(bu + q);
          // clear this whole piece from `touched`: all bits from its start upwards
          const unsigned startbit = piece_starts(up) & ((2u << p) - 1u);                       // starts at or below p
          const int sb = 31 - __clz(startbit);                                                  // the start of the piece containing p
          touched &= (1u << sb) - 1u;
        }
        if (a == 0 && (upl >> 31)) link(bul + __popc(piece_starts(upl)) - 1);                   // row above, word to the left
        s_par[me] = par;
      }
    }
  }
  __syncthreads();
  ccl_mark(clk, 3);
  volatile int* Lv = s_par;
  // D: flatten
  ccl_flatten(Lv, P);
  ccl_mark(clk, 4);
  // E: the remaining adjacencies, one pair per thread
  if (s_warp[40] <= kPairCap) {
    const int npairs = s_warp[40];
    for (int i = tid; i < npairs; i += kCclThreads) { const unsigned pr = s_pairs[i]; uf_union(Lv, s_par, (int)(pr & 0xFFFFu), (int)(pr >> 16)); }
  } else if (own) {
    // more pairs than the list holds (checkerboard-like occupancy): every thread unites ALL adjacencies of its own pieces
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const unsigned m = occ[h];
      if (!m) continue;
      const int w = w0 + h, x = w >> 3, k = w & 7;
      unsigned up = 0u, upl = 0u, upr = 0u;
      int bu = 0, bul = 0, bur = 0;
      if (x > 0) {
        up = s_occ[w - kRowWords]; bu = s_wbase[w - kRowWords];
        if (k > 0) { upl = s_occ[w - kRowWords - 1]; bul = s_wbase[w - kRowWords - 1]; }
        if (k < kRowWords - 1) { upr = s_occ[w - kRowWords + 1]; bur = s_wbase[w - kRowWords + 1]; }
      }
      const unsigned left = (k > 0) ? ((h == 1) ? occ[0] : s_occ[w - 1]) : 0u;
      const int bl = (k > 0) ? s_wbase[w - 1] : 0;
      unsigned rest = m;
      int me = s_wbase[w];
      while (rest) {
        const int a = __ffs(rest) - 1;
        const unsigned t = ~(rest >> a);
        const int len = t ? __ffs(t) - 1 : 32;
        const unsigned pm = (len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << a;
        rest &= ~pm;
        if (a == 0 && (left >> 31)) uf_union(Lv, s_par, me, bl + __popc(piece_starts(left)) - 1);
        if (a + len == 32 && (upr & 1u)) uf_union(Lv, s_par, me, bur);
        unsigned touched = up & (pm | (pm << 1) | (pm >> 1));
        while (touched) {
          const int p = 31 - __clz(touched);
          uf_union(Lv, s_par, me, bu + piece_of(up, p));
          const unsigned startbit = piece_starts(up) & ((2u << p) - 1u);
          touched &= (1u << (31 - __clz(startbit))) - 1u;
        }
        if (a == 0 && (upl >> 31)) uf_union(Lv, s_par, me, bul + __popc(piece_starts(upl)) - 1);
        ++me;
      }
    }
  }
  __syncthreads();
  ccl_mark(clk, 5);
  // F: flatten again
  ccl_flatten(Lv, P);
  ccl_mark(clk, 6);
  // G: id = 1 + rank of the root among all roots in id (= raster) order (:247-257); thread t ranks the ids [t*c, (t+1)*c)
  {
    const int c = (P + kCclThreads - 1) / kCclThreads;
    const int i0 = min(tid * c, P), i1 = min(i0 + c, P);
    int roots = 0;
    for (int id = i0; id < i1; ++id) roots += (s_par[id] == id) ? 1 : 0;
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
      if (lane == 31) F.counters[CNT_NUM_CLUSTER] = wi;
    }
    __syncthreads();
    int rank = s_warp[warp] + incl - roots;
    // (every non-root points at its root after F, so turning roots into -(id) cannot confuse a concurrent reader: nobody reads
    // parents between here and the barrier below)
    for (int id = i0; id < i1; ++id) if (s_par[id] == id) s_par[id] = -(++rank);
  }
  __syncthreads();
  ccl_mark(clk, 7);
  // H: label grid, sparse: cells occupied now get their id, cells occupied only in the previous frame are cleared
  if (own) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int w = w0 + h;
      const unsigned m = occ[h];
      int* o = out + (w >> 3) * kNumGrid + (w & 7) * 32;
      unsigned gone = s_prev[w] & ~m;
      while (gone) { const int b = __ffs(gone) - 1; gone &= gone - 1; o[b] = 0; }
      unsigned rest = m;
      int id = s_wbase[w];
      while (rest) {
        const int a = __ffs(rest) - 1;
        const unsigned t = ~(rest >> a);
        const int len = t ? __ffs(t) - 1 : 32;
        rest &= ~((len >= 32 ? 0xFFFFFFFFu : ((1u << len) - 1u)) << a);
        int r = s_par[id];
        ++id;
        if (r >= 0) r = s_par[r];                               // non-root: its root holds -(id)
        for (int b = a; b < a + len; ++b) o[b] = -r;
      }
    }
  }
  ccl_mark(clk, 8);
}

// ---- the cluster node's side outputs (src/cluster/main.cpp:62-99), SURVEY.md §8(f)3 ------------------------------
//   makeClusteredCloud (component_clustering.cpp:308-335): every elevated point inside the ROI whose cell carries a label
//       becomes the CENTRE of its cell at z = -1, in cloud order
//   setObsMsg (:337-375): the same, but the function zeroes the cell in its by-value grid copy once it has emitted it, so
//       only the FIRST point (cloud order) of every labelled cell yields an obstacle, tagged with the cluster id
//   createCostMap (:425-454): 50 x 50 cells of 1 m, 15 per point (capped at 100) for points with z <= 0.1 outside the car's
//       4.5 x 2 m footprint; indices by double arithmetic and truncation toward zero exactly as written
// One CTA; two order-preserving compactions (ballot + carried offset per 1024-point tile).
constexpr int kCostW = 50, kCostH = 50;

__global__ void __launch_bounds__(1024)
cluster_outputs_kernel(const float4* __restrict__ elev, const uint16_t* __restrict__ cart, const int* __restrict__ counters,
                       const int* __restrict__ label_grid, float roi, int* __restrict__ first_idx, float4* __restrict__ clustered,
                       float4* __restrict__ obstacles, int* __restrict__ cost_map, int* __restrict__ out_counts) {
  __shared__ int s_cost[kCostW * kCostH];
  __shared__ int s_warp[32], s_warp2[32];
  __shared__ int s_carry, s_carry2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = counters[CNT_N_ELEV];
  const float grid_size = 0.2f;                                   // component_clustering.h:15
  for (int k = tid; k < kCostW * kCostH; k += 1024) s_cost[k] = 0;
  if (tid == 0) { s_carry = 0; s_carry2 = 0; }
  __syncthreads();
  // pass 1: first point of every labelled cell; cost map counts
  const double res = 1.0, cx = (kCostW / 2.0) * res - 0.0, cy = (kCostH / 2.0) * res - 25.0;   // g_resolution, g_offset_x/y (:15-20)
  for (int i = tid; i < n; i += 1024) {
    const unsigned cc = cart[i];
    if (cc != kNoCell && label_grid[cc] != 0) atomicMin(&first_idx[cc], i);
    const float4 q = __ldg(&elev[i]);
    if (!((double)q.z > 0.1) && !(fabs((double)q.x) < 4.5 && fabs((double)q.y) < 2.0)) {
      const int gy = (int)(((double)q.x + cx) / res), gx = (int)(((double)q.y + cy) / res);
      if (!(gy < 0 || gy >= kCostW || gx < 0 || gx >= kCostH)) atomicAdd(&s_cost[kCostW * gx + gy], 1);
    }
  }
  __syncthreads();
  for (int k = tid; k < kCostW * kCostH; k += 1024) cost_map[k] = min(100, 15 * s_cost[k]);
  // pass 2: the two clouds, in cloud order
  for (int i0 = 0; i0 < n; i0 += 1024) {
    const int i = i0 + tid;
    int lab = 0, obs = 0;
    unsigned cc = kNoCell;
    if (i < n) {
      cc = cart[i];
      if (cc != kNoCell) { lab = label_grid[cc]; obs = (lab != 0 && first_idx[cc] == i) ? 1 : 0; }
    }
    const unsigned bl = __ballot_sync(0xFFFFFFFFu, lab != 0), bo = __ballot_sync(0xFFFFFFFFu, obs);
    if (lane == 0) { s_warp[warp] = __popc(bl); s_warp2[warp] = __popc(bo); }
    __syncthreads();
    int wb = 0, tot = 0, wb2 = 0, tot2 = 0;
    for (int w = 0; w < 32; ++w) { if (w < warp) { wb += s_warp[w]; wb2 += s_warp2[w]; } tot += s_warp[w]; tot2 += s_warp2[w]; }
    if (lab != 0) {
      const int xI = (int)(cc / (unsigned)kNumGrid), yI = (int)(cc % (unsigned)kNumGrid);
      const float half = fdiv(roi, 2.f), hg = fdiv(grid_size, 2.f);
      const float ox = fadd(fsub(fmul(grid_size, (float)xI), half), hg), oy = fadd(fsub(fmul(grid_size, (float)yI), half), hg);
      clustered[s_carry + wb + __popc(bl & ((1u << lane) - 1u))] = make_float4(ox, oy, -1.f, 1.f);
      if (obs) {
        obstacles[s_carry2 + wb2 + __popc(bo & ((1u << lane) - 1u))] = make_float4(ox, oy, -1.f, (float)lab);
        first_idx[cc] = INT_MAX;                                   // scratch back to its rest state
      }
    }
    __syncthreads();
    if (tid == 0) { s_carry += tot; s_carry2 += tot2; }
    __syncthreads();
  }
  if (tid == 0) { out_counts[0] = s_carry; out_counts[1] = s_carry2; }
}

__global__ void fill_i32_kernel(int* p, int n, int v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

__global__ void zero_u32_kernel(unsigned* p, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = 0u;
}

}  // namespace

int cluster_alloc(Ctx* c, Slot* s) {
  const size_t np = (size_t)c->max_points;
  LMOT_CUDA(c, cudaMalloc(&s->d_cart, np * sizeof(uint16_t)));
  LMOT_CUDA(c, cudaMalloc(&s->d_cart_bits, 3 * kBitWords * sizeof(unsigned)));      // once | twice | occupancy of the previous frame
  LMOT_CUDA(c, cudaMalloc(&s->d_label_grid, kCartCells * sizeof(int)));
  LMOT_CUDA(c, cudaMemsetAsync(s->d_cart_bits, 0, 3 * kBitWords * sizeof(unsigned), s->stream));
  LMOT_CUDA(c, cudaMemsetAsync(s->d_label_grid, 0, kCartCells * sizeof(int), s->stream));
  s->label_grid_foreign = false;
  LMOT_CUDA(c, cudaFuncSetAttribute(ccl_bitmap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kCclSmem));
  LMOT_CUDA(c, cudaFuncSetAttribute(ccl_dense_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDenseSmem));
  return LMOT_OK;
}

void cluster_free(Slot* s) {
  cudaFree(s->d_cart); cudaFree(s->d_cart_bits); cudaFree(s->d_label_grid);
  cudaFree(s->d_first_idx); cudaFree(s->d_clustered); cudaFree(s->d_obstacles); cudaFree(s->d_cost_map);
}

// elevated cloud = s->d_elev with its length in d_counters[CNT_N_ELEV]; n_upper bounds that length on the host
int cluster_launch(Ctx* c, Slot* s, cudaStream_t st, int n_upper, bool counted) {
  unsigned* once = s->d_cart_bits, *twice = once + kBitWords, *prev = twice + kBitWords;
  if (s->label_grid_foreign) {   // the caller uploaded its own label grid (lmot_box_fit): the sparse update needs a clean slate
    LMOT_CUDA(c, cudaMemsetAsync(s->d_label_grid, 0, kCartCells * sizeof(int), st));
    LMOT_CUDA(c, cudaMemsetAsync(prev, 0, kBitWords * sizeof(unsigned), st));
    s->label_grid_foreign = false;
  }
  if (n_upper > 0 && !counted) {   // `counted`: ground_fused_kernel already marked the elevated points (fused frame path)
    cart_mark_kernel<<<(n_upper + 255) / 256, 256, 0, st>>>(s->d_elev, s->d_counters, c->prm.roi_m, s->d_cart, once, twice);
    kernel_mark(c, s, st);
  }
  Slot* sl[1] = {s};
  return ccl_launch_batch(c, sl, 1, st);
}

// connected components of F frames in one launch, one CTA per frame (the bit planes were marked by the ground kernel)
int ccl_launch_batch(Ctx* c, Slot* const* slots, int F, cudaStream_t st) {
  if (F < 1 || F > kMaxBatch) return LMOT_ERR_INVALID;
  CclBatch B;
  for (int i = 0; i < F; ++i) {
    Slot* s = slots[i];
    B.f[i].once = s->d_cart_bits; B.f[i].twice = s->d_cart_bits + kBitWords; B.f[i].prev_occ = s->d_cart_bits + 2 * kBitWords;
    B.f[i].out = s->d_label_grid; B.f[i].counters = s->d_counters;
  }
  for (int i = F; i < kMaxBatch; ++i) B.f[i] = B.f[0];
  if (c->ccl_variant == 2) ccl_bitmap_kernel<<<F, kCclThreads, kCclSmem, st>>>(B, c->d_ccl_clock);     // LMOT_CCL=2: A/B only
  else ccl_dense_kernel<<<F, kCclThreads, kDenseSmem, st>>>(B, c->d_ccl_clock);
  kernel_mark(c, slots[0], st);
  LMOT_CUDA(c, cudaGetLastError());
  return LMOT_OK;
}

// side outputs of the cluster node for the slot's current elevated cloud + label grid; buffers are allocated on first use
int cluster_outputs_launch(Ctx* c, Slot* s, cudaStream_t st) {
  if (!s->d_first_idx) {
    LMOT_CUDA(c, cudaMalloc(&s->d_first_idx, kCartCells * sizeof(int)));
    LMOT_CUDA(c, cudaMalloc(&s->d_clustered, (size_t)c->max_points * sizeof(float4)));
    LMOT_CUDA(c, cudaMalloc(&s->d_obstacles, (size_t)kCartCells * sizeof(float4)));
    LMOT_CUDA(c, cudaMalloc(&s->d_cost_map, (kCostW * kCostH + 2) * sizeof(int)));
    fill_i32_kernel<<<(kCartCells + 255) / 256, 256, 0, st>>>(s->d_first_idx, kCartCells, INT_MAX);
  }
  cluster_outputs_kernel<<<1, 1024, 0, st>>>(s->d_elev, s->d_cart, s->d_counters, s->d_label_grid, c->prm.roi_m, s->d_first_idx,
                                             s->d_clustered, s->d_obstacles, s->d_cost_map, s->d_cost_map + kCostW * kCostH);
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

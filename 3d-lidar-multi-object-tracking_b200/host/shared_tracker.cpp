// shared_tracker.cpp -- liblmot_shared.so: several GPUs, ONE track table, NCCL over NVLink (include/lmot_shared.h).
//
// Host code in the reference's language over the C ABI of liblmot.so; the data path is device to device:
//   detection (lmot_detect_dev, this rank's frame)  ->  ncclAllGather of the box counts and of the padded box lists
//   ->  owner: lmot_track_step_lists_dev (one step on the concatenation, or one step per frame in frame order)
//   ->  ncclBroadcast of the track count, then of count x 1,648 B of the track table into every rank's own table.
// Reference shape: one tracking node consuming the boxes of several cluster nodes
// (/root/reference/object_tracking/tracking/main.cpp:98-141,166); the frame order / dt dependency of
// /root/reference/object_tracking/tracking/imm_ukf_jpda.cpp:807,812-961 is kept by folding the lists in rank (= frame) order.
#include <cuda_runtime.h>
#include <nccl.h>
#include <cstdio>
#include <cstring>
#include <string>
#include "lmot_shared.h"

struct lmot_shared {
  lmot_ctx* ctx = nullptr;
  int rank = 0, world = 1, owner = 0, device = 0;
  ncclComm_t comm = nullptr;
  cudaStream_t st = nullptr;
  cudaEvent_t ev[5] = {};
  int cap = 0;                       // boxes per rank in the gathered buffer (= the context's max_boxes)
  float* d_lists = nullptr;          // [world][cap][24]
  int* d_counts = nullptr;           // [world]
  int* h_T = nullptr;                // pinned: the track count, read back once per tick
  void* d_table = nullptr; int bytes_per_track = 0, table_cap = 0;
  int* d_ntracks = nullptr;
  float last_us[4] = {0, 0, 0, 0};
  std::string err;
};

#define SH_CUDA(s, call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { (s)->err = std::string(#call) + ": " + cudaGetErrorString(e__); return LMOT_ERR_CUDA; } } while (0)
#define SH_NCCL(s, call) do { ncclResult_t r__ = (call); if (r__ != ncclSuccess) { (s)->err = std::string(#call) + ": " + ncclGetErrorString(r__); return LMOT_ERR_CUDA; } } while (0)
#define SH_LMOT(s, call) do { int r__ = (call); if (r__ < 0) { (s)->err = std::string(#call) + ": " + lmot_strerror(r__) + " | " + lmot_last_error((s)->ctx); return r__; } } while (0)

extern "C" {

int lmot_shared_unique_id(void* id128) {
  if (!id128) return LMOT_ERR_INVALID;
  static_assert(sizeof(ncclUniqueId) <= LMOT_SHARED_ID_BYTES, "ncclUniqueId does not fit");
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return LMOT_ERR_CUDA;
  memset(id128, 0, LMOT_SHARED_ID_BYTES);
  memcpy(id128, &id, sizeof(id));
  return LMOT_OK;
}

const char* lmot_shared_last_error(const lmot_shared* s) { return s ? s->err.c_str() : ""; }

int lmot_shared_create(lmot_shared** out, lmot_ctx* ctx, int rank, int world, int owner, const void* id128) {
  if (!out || !ctx || !id128 || world < 1 || world > LMOT_MAX_BATCH || rank < 0 || rank >= world || owner < 0 || owner >= world) return LMOT_ERR_INVALID;
  *out = nullptr;
  lmot_shared* s = new lmot_shared();
  s->ctx = ctx; s->rank = rank; s->world = world; s->owner = owner;
  SH_CUDA(s, cudaGetDevice(&s->device));                      // lmot_create selected the context's device on this thread
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  SH_NCCL(s, ncclCommInitRank(&s->comm, world, id, rank));
  SH_CUDA(s, cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
  for (auto& e : s->ev) SH_CUDA(s, cudaEventCreate(&e));
  SH_LMOT(s, lmot_set_stream(ctx, (void*)s->st));
  lmot_params p;
  SH_LMOT(s, lmot_get_params(ctx, &p));
  s->cap = p.max_boxes;                                       // capacity of one rank's list in the gathered buffer
  SH_CUDA(s, cudaMalloc(&s->d_lists, (size_t)world * s->cap * 24 * sizeof(float)));
  SH_CUDA(s, cudaMalloc(&s->d_counts, (size_t)world * sizeof(int)));
  SH_CUDA(s, cudaMemset(s->d_counts, 0, (size_t)world * sizeof(int)));
  SH_CUDA(s, cudaHostAlloc(&s->h_T, sizeof(int), cudaHostAllocDefault));
  SH_LMOT(s, lmot_tracker_table(ctx, &s->d_table, &s->bytes_per_track, &s->table_cap));
  SH_LMOT(s, lmot_tracker_counters_dev(ctx, &s->d_ntracks));
  *out = s;
  return LMOT_OK;
}

void lmot_shared_destroy(lmot_shared* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  if (s->st) cudaStreamSynchronize(s->st);
  lmot_set_stream(s->ctx, nullptr);
  if (s->comm) ncclCommDestroy(s->comm);
  cudaFree(s->d_lists); cudaFree(s->d_counts);
  if (s->h_T) cudaFreeHost(s->h_T);
  for (auto& e : s->ev) if (e) cudaEventDestroy(e);
  if (s->st) cudaStreamDestroy(s->st);
  delete s;
}

int lmot_shared_tick_dev(lmot_shared* s, const float* d_points, int n, double ts, double v, double yaw, int mode, double frame_dt_us,
                         lmot_track_out* out) {
  if (!s || (mode != LMOT_SHARED_STREAMS && mode != LMOT_SHARED_FRAMES)) return LMOT_ERR_INVALID;
  SH_CUDA(s, cudaSetDevice(s->device));
  cudaStream_t st = s->st;
  SH_CUDA(s, cudaEventRecord(s->ev[0], st));
  // 1. detection of this rank's frame (ground -> cluster -> box on the context's slot stream), joined back into `st` on the device
  SH_LMOT(s, lmot_detect_dev(s->ctx, d_points, n));
  SH_LMOT(s, lmot_flush(s->ctx));
  const float* d_boxes = nullptr; const int* d_nb = nullptr;
  SH_LMOT(s, lmot_detect_boxes_dev(s->ctx, &d_boxes, &d_nb));
  SH_CUDA(s, cudaEventRecord(s->ev[1], st));
  // 2. every rank's count and padded box list to every rank (the owner needs them; symmetric all-gather keeps one code path and lets
  //    any rank associate against the other streams' boxes)
  SH_NCCL(s, ncclGroupStart());
  SH_NCCL(s, ncclAllGather(d_nb, s->d_counts, 1, ncclInt, s->comm, st));
  SH_NCCL(s, ncclAllGather(d_boxes, s->d_lists, (size_t)s->cap * 24, ncclFloat, s->comm, st));
  SH_NCCL(s, ncclGroupEnd());
  SH_CUDA(s, cudaEventRecord(s->ev[2], st));
  // 3. the owner folds the lists into the shared table: one step on the concatenation, or one step per frame in frame order
  if (s->rank == s->owner) {
    if (mode == LMOT_SHARED_STREAMS) SH_LMOT(s, lmot_track_step_lists_dev(s->ctx, s->d_lists, s->d_counts, s->world, s->cap, ts, v, yaw));
    else
      for (int r = 0; r < s->world; ++r)
        SH_LMOT(s, lmot_track_step_lists_dev(s->ctx, s->d_lists + (size_t)r * s->cap * 24, s->d_counts + r, 1, s->cap, ts + r * frame_dt_us, v, yaw));
  }
  SH_CUDA(s, cudaEventRecord(s->ev[3], st));
  // 4. the table travels: count first (4 bytes to the host of every rank: it sizes the second broadcast), then count x 1,648 B
  SH_NCCL(s, ncclBroadcast(s->d_ntracks, s->d_ntracks, 1, ncclInt, s->owner, s->comm, st));
  SH_CUDA(s, cudaMemcpyAsync(s->h_T, s->d_ntracks, sizeof(int), cudaMemcpyDeviceToHost, st));
  SH_CUDA(s, cudaStreamSynchronize(st));
  const int T = *s->h_T;
  if (T < 0 || T > s->table_cap) { s->err = "track count out of range"; return LMOT_ERR_STATE; }
  if (T > 0) SH_NCCL(s, ncclBroadcast(s->d_table, s->d_table, (size_t)T * s->bytes_per_track, ncclChar, s->owner, s->comm, st));
  SH_CUDA(s, cudaEventRecord(s->ev[4], st));
  if (s->rank != s->owner) SH_LMOT(s, lmot_tracker_table_received(s->ctx, T));
  int rc = LMOT_OK;
  if (s->rank == s->owner) {
    lmot_frame_out fo;
    memset(&fo, 0, sizeof(fo));
    if (out) fo.tracks = *out;
    rc = lmot_frame_fetch(s->ctx, &fo);            // waits for the tick's last step; its outputs sit in pinned host memory
    if (out) *out = fo.tracks;
    if (rc < 0) { s->err = std::string("lmot_frame_fetch: ") + lmot_strerror(rc); return rc; }
  } else if (out) { out->n_tracks = T; out->n_vis = 0; }
  SH_CUDA(s, cudaStreamSynchronize(st));
  for (int i = 0; i < 4; ++i) { float ms = 0; cudaEventElapsedTime(&ms, s->ev[i], s->ev[i + 1]); s->last_us[i] = 1e3f * ms; }
  return rc;
}

int lmot_shared_last_us(lmot_shared* s, float us[4]) {
  if (!s || !us) return LMOT_ERR_INVALID;
  for (int i = 0; i < 4; ++i) us[i] = s->last_us[i];
  return LMOT_OK;
}

}  // extern "C"

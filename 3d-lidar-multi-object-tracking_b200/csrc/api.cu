// api.cu -- the extern "C" boundary declared in include/lmot.h.  Host-side marshalling only: H2D/D2H copies,
// stream/event ordering of the frame pipeline and capacity checks; every algorithmic step is a CUDA kernel in
// ground.cu / cluster.cu / boxfit.cu / tracker.cu.  There is deliberately no CPU implementation of any stage here.
//
// Frame pipeline: a context owns `pipeline_depth` detection slots (own stream + own buffers each), one tracker stream
// and one publish stream.  Frame f runs ground -> cluster -> box on slot f % depth; the tracker chain (gate -> TA -> TB
// -> TC, tracker.cu) waits ON THE DEVICE for the slot's boxes and folds them into the track table; publish_kernel moves
// the frame's results (counts, boxes, per-track outputs) from their device block into the pinned, device-mapped host
// block of the result ring and releases the slot; the host only waits on an event.
#include <cmath>
#include <cstring>
#include <cstdint>
#include <cstdlib>
#include <new>
#include <chrono>
#include <vector>
#include "lmot_internal.cuh"
#include "exact_math.cuh"

using namespace lmot;

struct lmot_ctx {
  Ctx c;
};

namespace {

inline double now_ns() { return (double)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// gaussKernel(samples=3, sigma=1.0) exactly as gaus_blur.cpp:26-49 evaluates it (host libm, double)
void gauss_taps(double tap[3]) {
  const int samples = 3;
  const double sigma = 1.0;
  const double mean = samples / 2;  // integer division, = 1 (gaus_blur.cpp:28)
  double sum = 0.0;
  for (int x = 0; x < samples; ++x) {
    tap[x] = std::exp(-0.5 * (std::pow((x - mean) / sigma, 2.0))) / (2 * M_PI * sigma * sigma);
    sum += tap[x];
  }
  for (int x = 0; x < samples; ++x) tap[x] /= sum;
}

int upload_points(Ctx* c, Slot* s, cudaStream_t st, const float* points, int n, int stride, float4* d_dst) {
  if (n < 0 || (n > 0 && !points) || stride < 3) return LMOT_ERR_INVALID;
  if (n > c->max_points) return LMOT_ERR_CAPACITY;
  if (n == 0) return LMOT_OK;
  if (stride == 4) {
    LMOT_CUDA(c, cudaMemcpyAsync(d_dst, points, (size_t)n * 16, cudaMemcpyHostToDevice, st));
  } else {
    if (stride > 4) {  // wide host records: pack xyz on the host side of the copy
      std::vector<float> tmp((size_t)n * 3);
      for (int i = 0; i < n; ++i) { tmp[3*i] = points[(size_t)i*stride]; tmp[3*i+1] = points[(size_t)i*stride+1]; tmp[3*i+2] = points[(size_t)i*stride+2]; }
      LMOT_CUDA(c, cudaMemcpyAsync(s->d_stage_in, tmp.data(), (size_t)n * 12, cudaMemcpyHostToDevice, st));
      LMOT_CUDA(c, cudaStreamSynchronize(st));
      stride = 3;
    } else {
      LMOT_CUDA(c, cudaMemcpyAsync(s->d_stage_in, points, (size_t)n * stride * 4, cudaMemcpyHostToDevice, st));
    }
    int rc = ground_repack(c, st, s->d_stage_in, n, stride, d_dst);
    if (rc) return rc;
  }
  return LMOT_OK;
}

int fetch_counters(Ctx* c, Slot* s, cudaStream_t st) {
  LMOT_CUDA(c, cudaMemcpyAsync(s->h_counters, s->d_counters, CNT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, st));
  LMOT_CUDA(c, cudaStreamSynchronize(st));
  return LMOT_OK;
}

// write a host-known count into the slot's device counter block (entry points that start mid-pipeline)
int set_counter(Ctx* c, Slot* s, cudaStream_t st, int which, int value) {
  s->h_set[which] = value;
  LMOT_CUDA(c, cudaMemcpyAsync(s->d_counters + which, s->h_set + which, sizeof(int), cudaMemcpyHostToDevice, st));
  return LMOT_OK;
}

int check_device_error(Ctx* c, Slot* s, cudaStream_t st) {
  const int e = s->h_counters[CNT_ERROR];
  if (e != 0) {
    set_counter(c, s, st, CNT_ERROR, 0);
    cudaStreamSynchronize(st);
    return e;
  }
  return LMOT_OK;
}

// wait until nothing of this context is running on the device and forget uncollected results
int drain(Ctx* c) {
  for (int i = 0; i < c->n_slots; ++i) LMOT_CUDA(c, cudaStreamSynchronize(c->slots[i].stream));
  for (int i = 0; i < c->n_bslots; ++i) LMOT_CUDA(c, cudaStreamSynchronize(c->bslots[i].stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->trk_stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->pub_stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < c->n_results; ++i) c->results[i].in_flight = false;
  c->n_in_flight = 0;
  c->res_oldest = c->res_next;
  return LMOT_OK;
}

// next detection slot of the ring (reuse is ordered on the device by the slot's ev_trk_done)
Slot* acquire_slot(Ctx* c) {
  Slot* s = &c->slots[c->next_slot];
  c->last_slot = c->next_slot;
  c->next_slot = (c->next_slot + 1) % c->n_slots;
  return s;
}

// next result block of the ring; `drop_oldest`: an uncollected result still sitting in it is dropped, otherwise the
// caller gets nullptr (ring full: collect first)
Result* acquire_result(Ctx* c, bool drop_oldest) {
  Result* r = &c->results[c->res_next];
  if (r->in_flight) {
    if (!drop_oldest) return nullptr;
    cudaEventSynchronize(r->ev_done);            // its kernels may still be writing the block
    r->in_flight = false;
    --c->n_in_flight;
    c->res_oldest = (c->res_next + 1) % c->n_results;
  }
  c->res_next = (c->res_next + 1) % c->n_results;
  c->last_res = r;
  return r;
}

// the tracker chain + publication of one tick: boxes / counters of `s` (a detection slot, or a batch bank) -> track table -> result block r
int submit_tracker(Ctx* c, Slot* s, Result* r, const float* d_boxes, const int* d_counters, double ts, double v, double yaw) {
  int rc;
  // the tracker waits for this tick's detection on the DEVICE (gate kernel), or on ev_det_done where that is not possible
  bool gated = false;
  if ((rc = tracker_launch(c, s, c->trk_stream, d_boxes, d_counters, ts, v, yaw, true, &gated))) return rc;
  if (gated) {
    // Nothing but kernels on the tracker stream (an event record between spawn_output_kernel and the next frame's gate
    // kernel would serialise the two launches): publish_kernel polls the device-side step count, and "the slot is free"
    // is recorded behind it on the publish stream -- ~10 us later than strictly necessary, detection has >100 us of slack
    if ((rc = tracker_publish(c, r, c->pub_stream))) return rc;
    LMOT_CUDA(c, cudaEventRecord(r->ev_done, c->pub_stream));
    LMOT_CUDA(c, cudaEventRecord(s->ev_trk_done, c->pub_stream));
  } else {
    if (c->timing) cudaEventRecord(r->ev[4], c->trk_stream);
    LMOT_CUDA(c, cudaEventRecord(s->ev_trk_done, c->trk_stream));      // the slot is free: its boxes / counters were consumed
    LMOT_CUDA(c, cudaEventRecord(r->ev_tc, c->trk_stream));
    // device block -> pinned host block, off the tracker's sequential chain
    LMOT_CUDA(c, cudaStreamWaitEvent(c->pub_stream, r->ev_tc, 0));
    if ((rc = tracker_publish(c, r, c->pub_stream))) return rc;
    LMOT_CUDA(c, cudaEventRecord(r->ev_done, c->pub_stream));
  }
  return LMOT_OK;
}

// the asynchronous frame: detection on the slot stream, tracker on the tracker stream
int submit(Ctx* c, Slot* s, Result* r, const float4* d_pts, int n, bool with_tracker, double ts, double v, double yaw) {
  int rc;
  struct Tail { Ctx* c; double t0; ~Tail() { c->host_ns[0] += now_ns() - t0; } } tail{c, now_ns()};
  s->res = r;
  r->n_kev = 0;
  r->batch_frames = 0;
  // the slot's previous boxes / counters must have been consumed by the tracker
  LMOT_CUDA(c, cudaStreamWaitEvent(s->stream, s->ev_trk_done, 0));
  if (c->timing) cudaEventRecord(r->ev[0], s->stream);
  if ((rc = ground_launch(c, s, s->stream, d_pts, n, true, false, c->fuse_ccl))) return rc;
  if (c->timing) cudaEventRecord(r->ev[1], s->stream);
  if (!c->fuse_ccl && (rc = cluster_launch(c, s, s->stream, n, true))) return rc;      // (fused: the ground kernel's last CTA did it)
  if (c->timing) cudaEventRecord(r->ev[2], s->stream);
  const bool gf = with_tracker && c->prm.global_frame;
  if ((rc = boxfit_launch(c, s, s->stream, n, with_tracker && !gf))) return rc;      // with_tracker: posts the slot's detection semaphore
  if (gf && (rc = boxes_to_global_launch(c, s, s->stream, s->d_boxes, s->d_counters, ts, v, yaw, true))) return rc;   // ... or this one does
  if (c->timing) cudaEventRecord(r->ev[3], s->stream);
  LMOT_CUDA(c, cudaEventRecord(s->ev_det_done, s->stream));
  if (with_tracker) {
    if ((rc = submit_tracker(c, s, r, s->d_boxes, s->d_counters, ts, v, yaw))) return rc;
  } else {
    // no spawn_output_kernel will snapshot the counters / boxes: copy them before the slot is reused
    LMOT_CUDA(c, cudaMemcpyAsync(r->h_det, s->d_counters, CNT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, s->stream));
    if ((rc = boxes_publish(c, s, r, s->stream))) return rc;
    LMOT_CUDA(c, cudaEventRecord(s->ev_trk_done, s->stream));
    LMOT_CUDA(c, cudaEventRecord(r->ev_done, s->stream));
  }
  r->has_tracks = with_tracker;
  if (!r->in_flight) { r->in_flight = true; ++c->n_in_flight; }
  return LMOT_OK;
}

// ---- batched ticks: F sensor streams, one frame each, ONE track table (BASELINE.json configs[3]) -------------------------
// Their own detection slots (allocated on first use, two banks of F so that the detection of tick t+1 overlaps the tracker of
// tick t); a bank's kernels all run on its first slot's stream.  `bank` is a pseudo-slot: the concatenated box list, its
// counters, the detection semaphore and the events the tracker hand-over needs.
int slot_create(Ctx* c, Slot* s, int index);
void slot_destroy(Slot* s);

int batch_ensure(Ctx* c, int F) {
  if (F <= c->batch_frames) return LMOT_OK;
  // (re)allocate for the larger batch: nothing of the old banks may be in flight
  for (int i = 0; i < c->n_bslots; ++i) LMOT_CUDA(c, cudaStreamSynchronize(c->bslots[i].stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->trk_stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->pub_stream));
  for (int i = c->n_bslots; i < kBatchBanks * F; ++i) {
    int rc = slot_create(c, &c->bslots[i], 1000 + i);
    if (rc) return rc;
    c->n_bslots = i + 1;
  }
  for (int b = 0; b < kBatchBanks; ++b) {
    Slot* k = &c->bank[b];
    if (k->d_boxes) continue;
    k->index = 2000 + b;
    LMOT_CUDA(c, cudaMalloc(&k->d_boxes, (size_t)c->prm.max_boxes * 24 * sizeof(float)));
    LMOT_CUDA(c, cudaMalloc(&k->d_boxes_g, (size_t)c->prm.max_boxes * 24 * sizeof(float)));
    LMOT_CUDA(c, cudaMalloc(&k->d_counters, CNT_COUNT * sizeof(int)));
    LMOT_CUDA(c, cudaMemset(k->d_counters, 0, CNT_COUNT * sizeof(int)));
    LMOT_CUDA(c, cudaMalloc(&k->d_det_sem, sizeof(int)));
    LMOT_CUDA(c, cudaMemset(k->d_det_sem, 0, sizeof(int)));
    LMOT_CUDA(c, cudaEventCreateWithFlags(&k->ev_det_done, cudaEventDisableTiming));
    LMOT_CUDA(c, cudaEventCreateWithFlags(&k->ev_trk_done, cudaEventDisableTiming));
    LMOT_CUDA(c, cudaEventCreateWithFlags(&k->ev_fork, cudaEventDisableTiming));
    LMOT_CUDA(c, cudaEventRecord(k->ev_trk_done, c->stream));
    LMOT_CUDA(c, cudaEventRecord(k->ev_det_done, c->stream));
    k->res = &c->results[0];
  }
  LMOT_CUDA(c, cudaDeviceSynchronize());
  c->batch_frames = F;
  return LMOT_OK;
}

void batch_destroy(Ctx* c) {
  for (int i = 0; i < c->n_bslots; ++i) slot_destroy(&c->bslots[i]);
  for (int b = 0; b < kBatchBanks; ++b) {
    Slot* k = &c->bank[b];
    cudaFree(k->d_boxes); cudaFree(k->d_boxes_g); cudaFree(k->d_counters); cudaFree(k->d_det_sem);
    if (k->ev_det_done) cudaEventDestroy(k->ev_det_done);
    if (k->ev_trk_done) cudaEventDestroy(k->ev_trk_done);
    if (k->ev_fork) cudaEventDestroy(k->ev_fork);
  }
}

// one tick: F device-resident frames (slot buffers or caller pointers) -> batched ground / CCL / box fitting -> concatenated boxes
// -> (optionally) the tracker.  d_pts[i] == nullptr: the frame was uploaded into the slot's own buffer.
int batch_submit(Ctx* c, int bank_i, Result* r, const float4* const* d_pts, const int* n, int F, bool with_tracker, double ts, double v, double yaw) {
  int rc;
  struct Tail { Ctx* c; double t0; ~Tail() { c->host_ns[0] += now_ns() - t0; } } tail{c, now_ns()};
  Slot* k = &c->bank[bank_i];
  Slot* sl[kMaxBatch];
  const float4* pp[kMaxBatch];
  for (int i = 0; i < F; ++i) { sl[i] = &c->bslots[bank_i * c->batch_frames + i]; pp[i] = d_pts[i] ? d_pts[i] : sl[i]->d_points; }
  cudaStream_t st = sl[0]->stream;
  k->res = r; sl[0]->res = r;
  r->n_kev = 0;
  r->batch_frames = F;
  if (c->timing) cudaEventRecord(r->ev[0], st);
  if ((rc = ground_launch_batch(c, sl, F, pp, n, st, true, false, c->fuse_ccl))) return rc;
  if (c->timing) cudaEventRecord(r->ev[1], st);
  if (!c->fuse_ccl && (rc = ccl_launch_batch(c, sl, F, st))) return rc;
  if (c->timing) cudaEventRecord(r->ev[2], st);
  if ((rc = boxfit_launch_batch(c, sl, F, st, n, false))) return rc;
  int* d_fc = nullptr;                       // the per-frame counts go straight into the result's pinned, device-mapped block
  LMOT_CUDA(c, cudaHostGetDevicePointer((void**)&d_fc, r->h_frame_counts, 0));
  const bool gf = with_tracker && c->prm.global_frame;
  if ((rc = boxes_concat_launch(c, sl, F, st, k->d_boxes, k->d_counters, d_fc, (with_tracker && !gf) ? k->d_det_sem : nullptr))) return rc;
  if (gf && (rc = boxes_to_global_launch(c, k, st, k->d_boxes, k->d_counters, ts, v, yaw, true))) return rc;
  if (c->timing) cudaEventRecord(r->ev[3], st);
  LMOT_CUDA(c, cudaEventRecord(k->ev_det_done, st));
  if (with_tracker) {
    if ((rc = submit_tracker(c, k, r, k->d_boxes, k->d_counters, ts, v, yaw))) return rc;
  } else {
    LMOT_CUDA(c, cudaMemcpyAsync(r->h_det, k->d_counters, CNT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, st));
    if ((rc = boxes_publish(c, k, r, st))) return rc;
    LMOT_CUDA(c, cudaEventRecord(k->ev_trk_done, st));
    LMOT_CUDA(c, cudaEventRecord(r->ev_done, st));
  }
  r->has_tracks = with_tracker;
  if (!r->in_flight) { r->in_flight = true; ++c->n_in_flight; }
  return LMOT_OK;
}

int copy_track_outputs(const Result* s, lmot_track_out* out) {
  const int T = s->h_hdr[HDR_N_TRACKS], nv = s->h_hdr[HDR_N_VIS];
  if (!out) return LMOT_OK;
  out->n_tracks = T; out->n_vis = nv;
  const int n = T < out->cap ? T : out->cap;
  const int nvc = nv < out->cap ? nv : out->cap;
  if (n > 0) {
    if (out->targets) memcpy(out->targets, s->h_targets, (size_t)n * 3 * sizeof(float));
    if (out->vandyaw) memcpy(out->vandyaw, s->h_vandyaw, (size_t)n * 2 * sizeof(double));
    if (out->track_manage) memcpy(out->track_manage, s->h_manage, (size_t)n * sizeof(int));
    if (out->is_static) memcpy(out->is_static, s->h_static, (size_t)n);
    if (out->is_vis) memcpy(out->is_vis, s->h_vis, (size_t)n);
  }
  if (nvc > 0 && out->vis_bb) memcpy(out->vis_bb, s->h_visbb, (size_t)nvc * 24 * sizeof(float));
  const bool wants = out->targets || out->vandyaw || out->track_manage || out->is_static || out->is_vis || out->vis_bb;
  return (wants && T > out->cap) ? LMOT_ERR_CAPACITY : LMOT_OK;      // a zero-initialised lmot_track_out asks for the counts only
}

// results of a finished frame from its pinned host block
int collect_result(Ctx* c, Result* r, lmot_frame_out* out) {
  const double t0 = now_ns();
  LMOT_CUDA(c, cudaEventSynchronize(r->ev_done));
  const double t1 = now_ns();
  c->host_ns[1] += t1 - t0; c->host_ns[3] += 1;
  struct Tail { Ctx* c; double t1; ~Tail() { c->host_ns[2] += now_ns() - t1; } } tail{c, t1};
  if (c->timing && r->has_tracks) {
    for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&c->stage_ms[i], r->ev[i], r->ev[i + 1]);
    c->n_kernel_ms = r->n_kev;
    for (int i = 0; i < r->n_kev; ++i) cudaEventElapsedTime(&c->kernel_ms[i], i == 0 ? r->ev[0] : r->kev[i - 1], r->kev[i]);
  }
  if (!r->has_tracks) {   // detect only: no kernel wrote the header
    r->h_hdr[HDR_N_ELEV] = r->h_det[CNT_N_ELEV]; r->h_hdr[HDR_N_GROUND] = r->h_det[CNT_N_GROUND];
    r->h_hdr[HDR_NUM_CLUSTER] = r->h_det[CNT_NUM_CLUSTER]; r->h_hdr[HDR_N_BOXES] = r->h_det[CNT_N_BOXES];
    r->h_hdr[HDR_N_TRACKS] = 0; r->h_hdr[HDR_N_VIS] = 0; r->h_hdr[HDR_ERROR] = r->h_det[CNT_ERROR]; r->h_hdr[HDR_WARN] = 0;
  }
  const int err = r->h_hdr[HDR_ERROR];
  const int warn = r->has_tracks ? r->h_hdr[HDR_WARN] : 0;
  if (r->has_tracks) c->last_n_act = r->h_hdr[HDR_N_ACT];
  if (out) {
    out->n_elevated = r->h_hdr[HDR_N_ELEV]; out->n_ground = r->h_hdr[HDR_N_GROUND];
    out->num_cluster = r->h_hdr[HDR_NUM_CLUSTER]; out->n_boxes = r->h_hdr[HDR_N_BOXES];
    const int nb = out->n_boxes < out->max_boxes ? out->n_boxes : out->max_boxes;
    if (out->boxes && nb > 0) memcpy(out->boxes, r->h_boxes, (size_t)nb * 24 * sizeof(float));
    const int rc = copy_track_outputs(r, &out->tracks);
    if (rc && !err) return rc;
  }
  return err ? err : warn;      // warn > 0 (LMOT_WARN_*): the outputs are valid
}

int result_create(Ctx* c, Result* r) {
  const int TC = c->prm.max_tracks, MB = c->prm.max_boxes;
  const unsigned fl = cudaHostAllocMapped;
  LMOT_CUDA(c, cudaHostAlloc(&r->h_hdr, HDR_COUNT * sizeof(int) + 16, fl));
  memset(r->h_hdr, 0, HDR_COUNT * sizeof(int));
  LMOT_CUDA(c, cudaHostAlloc(&r->h_det, CNT_COUNT * sizeof(int), fl));
  memset(r->h_det, 0, CNT_COUNT * sizeof(int));
  LMOT_CUDA(c, cudaHostAlloc(&r->h_frame_counts, 4 * kMaxBatch * sizeof(int), fl));
  memset(r->h_frame_counts, 0, 4 * kMaxBatch * sizeof(int));
  LMOT_CUDA(c, cudaHostAlloc(&r->h_boxes, (size_t)MB * 24 * sizeof(float) + 16, fl));
  // + 16 bytes: spawn_output_kernel writes whole 16-byte words
  LMOT_CUDA(c, cudaHostAlloc(&r->h_targets, (size_t)TC * 3 * sizeof(float) + 16, fl));
  LMOT_CUDA(c, cudaHostAlloc(&r->h_vandyaw, (size_t)TC * 2 * sizeof(double) + 16, fl));
  LMOT_CUDA(c, cudaHostAlloc(&r->h_manage, (size_t)TC * sizeof(int) + 16, fl));
  LMOT_CUDA(c, cudaHostAlloc(&r->h_static, (size_t)TC + 16, fl));
  LMOT_CUDA(c, cudaHostAlloc(&r->h_vis, (size_t)TC + 16, fl));
  LMOT_CUDA(c, cudaHostAlloc(&r->h_visbb, (size_t)TC * 24 * sizeof(float) + 16, fl));
  {   // device copy of the block, one allocation, every array 256-byte aligned with 16 bytes of slack
    auto al = [](size_t b) { return (b + 16 + 255) & ~(size_t)255; };
    const size_t sz[8] = {al(HDR_COUNT * sizeof(int)), al((size_t)MB * 96), al((size_t)TC * 12), al((size_t)TC * 16), al((size_t)TC * 4),
                          al((size_t)TC), al((size_t)TC), al((size_t)TC * 96)};
    size_t tot = 0;
    for (size_t b : sz) tot += b;
    LMOT_CUDA(c, cudaMalloc(&r->d_block, tot));
    LMOT_CUDA(c, cudaMemsetAsync(r->d_block, 0, tot, c->stream));
    unsigned char* p = r->d_block;
    r->d_hdr = (int*)p; p += sz[0]; r->d_boxes = (float*)p; p += sz[1]; r->d_targets = (float*)p; p += sz[2];
    r->d_vandyaw = (double*)p; p += sz[3]; r->d_manage = (int*)p; p += sz[4]; r->d_static = p; p += sz[5]; r->d_vis = p; p += sz[6];
    r->d_visbb = (float*)p;
  }
  LMOT_CUDA(c, cudaEventCreateWithFlags(&r->ev_tc, cudaEventDisableTiming));
  LMOT_CUDA(c, cudaEventCreateWithFlags(&r->ev_done, cudaEventDisableTiming));
  LMOT_CUDA(c, cudaEventRecord(r->ev_done, c->stream));
  for (int i = 0; i < 5; ++i) LMOT_CUDA(c, cudaEventCreate(&r->ev[i]));
  for (int i = 0; i < kMaxKernelEvents; ++i) LMOT_CUDA(c, cudaEventCreate(&r->kev[i]));
  return LMOT_OK;
}

void result_destroy(Result* r) {
  if (r->h_hdr) cudaFreeHost(r->h_hdr);
  if (r->h_det) cudaFreeHost(r->h_det);
  if (r->h_frame_counts) cudaFreeHost(r->h_frame_counts);
  if (r->h_boxes) cudaFreeHost(r->h_boxes);
  if (r->h_targets) cudaFreeHost(r->h_targets);
  if (r->h_vandyaw) cudaFreeHost(r->h_vandyaw);
  if (r->h_manage) cudaFreeHost(r->h_manage);
  if (r->h_static) cudaFreeHost(r->h_static);
  if (r->h_vis) cudaFreeHost(r->h_vis);
  if (r->h_visbb) cudaFreeHost(r->h_visbb);
  if (r->d_block) cudaFree(r->d_block);
  if (r->ev_tc) cudaEventDestroy(r->ev_tc);
  if (r->ev_done) cudaEventDestroy(r->ev_done);
  for (int i = 0; i < 5; ++i) if (r->ev[i]) cudaEventDestroy(r->ev[i]);
  for (int i = 0; i < kMaxKernelEvents; ++i) if (r->kev[i]) cudaEventDestroy(r->kev[i]);
}

int slot_create(Ctx* c, Slot* s, int index) {
  s->index = index;
  LMOT_CUDA(c, cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  LMOT_CUDA(c, cudaEventCreateWithFlags(&s->ev_fork, cudaEventDisableTiming));
  LMOT_CUDA(c, cudaEventCreateWithFlags(&s->ev_det_done, cudaEventDisableTiming));
  LMOT_CUDA(c, cudaEventCreateWithFlags(&s->ev_trk_done, cudaEventDisableTiming));
  int rc = ground_alloc(c, s);
  if (rc == LMOT_OK) rc = cluster_alloc(c, s);
  if (rc == LMOT_OK) rc = boxfit_alloc(c, s);
  if (rc) return rc;
  LMOT_CUDA(c, cudaEventRecord(s->ev_trk_done, s->stream));   // "free" from the start
  LMOT_CUDA(c, cudaEventRecord(s->ev_det_done, s->stream));
  s->res = &c->results[0];
  return LMOT_OK;
}

void slot_destroy(Slot* s) {
  ground_free(s); cluster_free(s); boxfit_free(s);
  if (s->ev_fork) cudaEventDestroy(s->ev_fork);
  if (s->ev_det_done) cudaEventDestroy(s->ev_det_done);
  if (s->ev_trk_done) cudaEventDestroy(s->ev_trk_done);
  if (s->stream) cudaStreamDestroy(s->stream);
}

}  // namespace

extern "C" {

int lmot_default_params(lmot_params* p) {
  if (!p) return LMOT_ERR_INVALID;
  memset(p, 0, sizeof(*p));
  p->r_min = 3.4f; p->r_max = 120.f; p->t_hmin = -2.0f; p->t_hmax = -0.4f; p->t_hdiff = 0.4f; p->h_sensor = 2.f;
  p->ground_tolerance = 0.25;
  p->roi_m = 50.f;
  p->ram_points = 80; p->l_slope_dist = 1; p->l_num_points = 5; p->sensor_height = 2.f;
  p->t_height_min = 0.8f; p->t_height_max = 2.6f; p->t_width_min = 0.2f; p->t_width_max = 3.5f;
  p->t_len_min = 0.2f; p->t_len_max = 14.0f; p->t_area_max = 20.0f; p->t_ratio_min = 1.f; p->t_ratio_max = 8.0f;
  p->min_len_ratio = 3.0f; p->t_pt_per_m3 = 8.f; p->min_cluster_points = 30;
  p->rule_filter = LMOT_RULE_INTENDED;
  p->oracle_compat_first_frame = 1;
  p->max_points = 1 << 20; p->max_clusters = 4096; p->max_boxes = 1024; p->max_tracks = 8192;
  p->global_frame = 0;
  p->node_prefilter = 0; p->filter_z_min = -3.0f; p->filter_z_max = 1.0f;
  p->filter_x_min = -15.f; p->filter_x_max = 5.f; p->filter_y_min = -50.f; p->filter_y_max = 50.f;
  p->pipeline_depth = 8;   // detection takes ~150 us per frame: 8 frames in flight keep it ahead of the ~40 us tracker chain
  p->result_ring = 32;
  return LMOT_OK;
}

const char* lmot_strerror(int s) {
  switch (s) {
    case LMOT_OK: return "ok";
    case LMOT_ERR_INVALID: return "invalid argument";
    case LMOT_ERR_CUDA: return "CUDA error (no CPU fallback exists)";
    case LMOT_ERR_CAPACITY: return "capacity exceeded";
    case LMOT_ERR_STATE: return "invalid call sequence / not available";
    case LMOT_WARN_TRACK_TABLE_FULL: return "warning: track table full (max_tracks): unmatched boxes spawn no new track, existing tracks are still updated and reported";
    default: return "unknown status";
  }
}

const char* lmot_build_info(void) { return "liblmot sm_100a, nvcc " __DATE__ " -fmad=false"; }

int lmot_get_params(const lmot_ctx* ctx, lmot_params* out) {
  if (!ctx || !out) return LMOT_ERR_INVALID;
  *out = ctx->c.prm;
  return LMOT_OK;
}

const char* lmot_last_error(const lmot_ctx* ctx) { return ctx ? ctx->c.last_error.c_str() : ""; }

int lmot_create(lmot_ctx** out, const lmot_params* params, int device) {
  if (!out) return LMOT_ERR_INVALID;
  *out = nullptr;
  lmot_params p;
  if (params) p = *params; else lmot_default_params(&p);
  if (p.max_points <= 0 || p.max_clusters <= 0 || p.max_boxes <= 0 || p.max_tracks <= 0) return LMOT_ERR_INVALID;
  if (p.max_clusters > 50000 || p.max_boxes > 65535) return LMOT_ERR_INVALID;   // u16 cluster ids / box indices; one int per cluster id in the sort kernels' shared memory (<= 200 KB)
  if (p.pipeline_depth < 1) p.pipeline_depth = 1;
  if (p.pipeline_depth > kMaxSlots) p.pipeline_depth = kMaxSlots;
  if (p.result_ring < 1) p.result_ring = 1;
  if (p.result_ring > kMaxResults) p.result_ring = kMaxResults;
  // the frame pipeline uses up to 8 detection streams + tracker + publish stream per context; with the default of 8 hardware work
  // queues, streams alias and serialise each other.  Only effective if this process has not created its CUDA context yet.
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return LMOT_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return LMOT_ERR_CUDA;
  lmot_ctx* h = new (std::nothrow) lmot_ctx();
  if (!h) return LMOT_ERR_INVALID;
  Ctx* c = &h->c;
  c->prm = p;
  c->device = device;
  c->max_points = p.max_points;
  c->n_slots = p.pipeline_depth;
  c->n_results = p.result_ring;
  c->gp.r_min = p.r_min; c->gp.r_max = p.r_max; c->gp.t_hmin = p.t_hmin; c->gp.t_hmax = p.t_hmax;
  c->gp.t_hdiff = p.t_hdiff; c->gp.h_sensor = p.h_sensor;
  { volatile float span = p.r_max - p.r_min; c->gp.r_span = span; }
  c->gp.prefilter = p.node_prefilter ? 1 : 0;
  c->gp.fz0 = p.filter_z_min; c->gp.fz1 = p.filter_z_max; c->gp.fx0 = p.filter_x_min; c->gp.fx1 = p.filter_x_max;
  c->gp.fy0 = p.filter_y_min; c->gp.fy1 = p.filter_y_max;
  c->gp.bin_scale = (float)((double)LMOT_NUM_BIN / (double)c->gp.r_span);
  c->gp.tol = p.ground_tolerance;
  gauss_taps(c->gp.tap);
  if (const char* e = getenv("LMOT_COOP")) c->coop_launch = atoi(e) != 0;
  if (const char* e = getenv("LMOT_FUSE_CCL")) c->fuse_ccl = atoi(e) != 0;
  if (const char* e = getenv("LMOT_SPIN_LIMIT")) c->spin_limit = (unsigned)strtoul(e, nullptr, 0);   // 0: device-side waits never trap (debuggers, MPS)
  if (const char* e = getenv("LMOT_TRK_CTAS")) { const int v = atoi(e); if (v >= 8 && v <= 4096) c->trk_ctas = v; }     // tuning only
  if (const char* e = getenv("LMOT_TC_WIDE")) c->tc_force = atoi(e) != 0 ? 1 : 0;      // tests: pin spawn_output_kernel's variant (default: by active-track count)
  if (const char* e = getenv("LMOT_FIT_CTAS")) { const int v = atoi(e); if (v >= 8 && v <= 4096) c->fit_ctas = v; }
  if (const char* e = getenv("LMOT_ZERO_COPY")) c->zero_copy = atoi(e) != 0;
  if (const char* e = getenv("LMOT_GROUND_HALF")) c->ground_half_sms = atoi(e) != 0;
  if (const char* e = getenv("LMOT_PTS_PER_CTA")) { const int v = atoi(e); if (v >= 256 && v <= 16384) c->pts_per_cta = v; }
  int rc = LMOT_OK;
  // The tracker is the one sequential chain of the pipeline (frame f+1's tracker needs frame f's table): its CTAs get
  // the highest stream priority so they are never queued behind the detection kernels of later frames.
  int prio_lo = 0, prio_hi = 0;
  cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithPriority(&c->trk_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->pub_stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return LMOT_ERR_CUDA; }
  c->stream = c->own_stream;
  rc = boxfit_alloc_shared(c);
  for (int i = 0; i < c->n_results && rc == LMOT_OK; ++i) rc = result_create(c, &c->results[i]);
  for (int i = 0; i < c->n_slots && rc == LMOT_OK; ++i) rc = slot_create(c, &c->slots[i], i);
  c->last_res = &c->results[0];
  if (rc == LMOT_OK) rc = tracker_alloc(c);
  if (rc == LMOT_OK) rc = (cudaDeviceSynchronize() == cudaSuccess) ? LMOT_OK : LMOT_ERR_CUDA;
  if (rc != LMOT_OK) { lmot_destroy(h); return rc; }
  *out = h;
  return LMOT_OK;
}

void lmot_destroy(lmot_ctx* ctx) {
  if (!ctx) return;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  ground_chain_forget(c->device, true);
  for (int i = 0; i < c->n_slots; ++i) slot_destroy(&c->slots[i]);
  batch_destroy(c);
  for (int i = 0; i < c->n_results; ++i) result_destroy(&c->results[i]);
  tracker_free(c);
  cudaFree(c->d_phase_clock); cudaFree(c->d_trk_trace); cudaFree(c->d_ccl_clock); cudaFree(c->d_fit_clock);
  cudaFree(c->d_mt_raw);
  if (c->trk_stream) cudaStreamDestroy(c->trk_stream);
  if (c->pub_stream) cudaStreamDestroy(c->pub_stream);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete ctx;
}

void* lmot_pinned_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}

void lmot_pinned_free(void* p) { if (p) cudaFreeHost(p); }

int lmot_set_stream(lmot_ctx* ctx, void* s) {
  if (!ctx) return LMOT_ERR_INVALID;
  cudaSetDevice(ctx->c.device);
  ground_chain_forget(ctx->c.device, false);       // (the stream being replaced must still be valid, lmot.h)
  ctx->c.stream = s ? (cudaStream_t)s : ctx->c.own_stream;
  return LMOT_OK;
}

int lmot_sync(lmot_ctx* ctx) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  for (int i = 0; i < c->n_slots; ++i) LMOT_CUDA(c, cudaStreamSynchronize(c->slots[i].stream));
  for (int i = 0; i < c->n_bslots; ++i) LMOT_CUDA(c, cudaStreamSynchronize(c->bslots[i].stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->trk_stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->pub_stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
  return LMOT_OK;
}

// make the caller's stream wait for everything submitted so far (so CUDA events recorded on it bracket the work)
int lmot_flush(lmot_ctx* ctx) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  for (int i = 0; i < c->n_slots; ++i) {
    LMOT_CUDA(c, cudaStreamWaitEvent(c->stream, c->slots[i].ev_det_done, 0));
    LMOT_CUDA(c, cudaStreamWaitEvent(c->stream, c->slots[i].ev_trk_done, 0));
  }
  for (int b = 0; b < kBatchBanks && c->batch_frames > 0; ++b) {
    LMOT_CUDA(c, cudaStreamWaitEvent(c->stream, c->bank[b].ev_det_done, 0));
    LMOT_CUDA(c, cudaStreamWaitEvent(c->stream, c->bank[b].ev_trk_done, 0));
  }
  if (c->last_res && c->last_res->ev_done) LMOT_CUDA(c, cudaStreamWaitEvent(c->stream, c->last_res->ev_done, 0));   // publication is in order
  return LMOT_OK;
}

// ---------------------------------------------------------------------------------------------- stage entry points
int lmot_ground_remove_dev(lmot_ctx* ctx, const float* d_points, int n) {
  if (!ctx || n < 0 || (n > 0 && !d_points) || (reinterpret_cast<uintptr_t>(d_points) & 15u)) return LMOT_ERR_INVALID;   // cp.async.bulk: 16-byte aligned
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  if (n > c->max_points) return LMOT_ERR_CAPACITY;
  if (c->n_in_flight > 0) { int rc = drain(c); if (rc) return rc; }     // slot 0 may still be busy with a pipelined frame
  Slot* s = &c->slots[0];
  c->last_slot = 0;
  // (no label array, no debug grid: what the frame pipeline launches, minus the clustering bit planes)
  return ground_launch(c, s, c->stream, reinterpret_cast<const float4*>(d_points), n, false, false, false, true);
}

int lmot_ground_remove(lmot_ctx* ctx, const float* points, int n, int stride, uint8_t* labels, float* elevated,
                       int* n_elevated, float* ground, int* n_ground) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  int rc = drain(c);
  if (rc) return rc;
  Slot* s = &c->slots[0];
  c->last_slot = 0;
  cudaStream_t st = c->stream;
  if ((rc = upload_points(c, s, st, points, n, stride, s->d_points))) return rc;
  if ((rc = ground_launch(c, s, st, s->d_points, n))) return rc;
  if ((rc = fetch_counters(c, s, st))) return rc;
  const int ne = s->h_counters[CNT_N_ELEV], ng = s->h_counters[CNT_N_GROUND];
  if (n_elevated) *n_elevated = ne;
  if (n_ground) *n_ground = ng;
  if (labels && n > 0) LMOT_CUDA(c, cudaMemcpyAsync(labels, s->d_labels, (size_t)n, cudaMemcpyDeviceToHost, st));
  if (elevated && ne > 0) LMOT_CUDA(c, cudaMemcpyAsync(elevated, s->d_elev, (size_t)ne * 16, cudaMemcpyDeviceToHost, st));
  if (ground && ng > 0) LMOT_CUDA(c, cudaMemcpyAsync(ground, s->d_ground, (size_t)ng * 16, cudaMemcpyDeviceToHost, st));
  LMOT_CUDA(c, cudaStreamSynchronize(st));
  return LMOT_OK;
}

int lmot_component_cluster(lmot_ctx* ctx, const float* elevated, int n, int stride, int32_t* grid, int* num_cluster) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  int rc = drain(c);
  if (rc) return rc;
  Slot* s = &c->slots[0];
  c->last_slot = 0;
  cudaStream_t st = c->stream;
  if ((rc = upload_points(c, s, st, elevated, n, stride, s->d_elev))) return rc;
  if ((rc = set_counter(c, s, st, CNT_N_ELEV, n))) return rc;
  if ((rc = cluster_launch(c, s, st, n))) return rc;
  if ((rc = fetch_counters(c, s, st))) return rc;
  if (num_cluster) *num_cluster = s->h_counters[CNT_NUM_CLUSTER];
  if (grid) {
    LMOT_CUDA(c, cudaMemcpyAsync(grid, s->d_label_grid, kCartCells * sizeof(int), cudaMemcpyDeviceToHost, st));
    LMOT_CUDA(c, cudaStreamSynchronize(st));
  }
  return LMOT_OK;
}

// the cluster node's side outputs for the elevated cloud + label grid of the most recent clustering of this context
int lmot_cluster_outputs(lmot_ctx* ctx, float* clustered, int cap_clustered, int* n_clustered, float* obstacles, int cap_obstacles,
                         int* n_obstacles, int32_t* cost_map) {
  if (!ctx || cap_clustered < 0 || cap_obstacles < 0) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  int rc = lmot_sync(ctx);
  if (rc) return rc;
  Slot* s = &c->slots[c->last_slot];
  cudaStream_t st = c->stream;
  if ((rc = cluster_outputs_launch(c, s, st))) return rc;
  int cnt[2] = {0, 0};
  LMOT_CUDA(c, cudaMemcpyAsync(cnt, s->d_cost_map + 2500, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  if (cost_map) LMOT_CUDA(c, cudaMemcpyAsync(cost_map, s->d_cost_map, 2500 * sizeof(int), cudaMemcpyDeviceToHost, st));
  LMOT_CUDA(c, cudaStreamSynchronize(st));
  if (n_clustered) *n_clustered = cnt[0];
  if (n_obstacles) *n_obstacles = cnt[1];
  const int nc = cnt[0] < cap_clustered ? cnt[0] : cap_clustered, no = cnt[1] < cap_obstacles ? cnt[1] : cap_obstacles;
  if (clustered && nc > 0) LMOT_CUDA(c, cudaMemcpyAsync(clustered, s->d_clustered, (size_t)nc * 16, cudaMemcpyDeviceToHost, st));
  if (obstacles && no > 0) LMOT_CUDA(c, cudaMemcpyAsync(obstacles, s->d_obstacles, (size_t)no * 16, cudaMemcpyDeviceToHost, st));
  LMOT_CUDA(c, cudaStreamSynchronize(st));
  return (cnt[0] > cap_clustered && clustered) || (cnt[1] > cap_obstacles && obstacles) ? LMOT_ERR_CAPACITY : LMOT_OK;
}

int lmot_debug_label_grid(lmot_ctx* ctx, int32_t* grid, int* num_cluster) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  int rc = lmot_sync(ctx);
  if (rc) return rc;
  Slot* s = &c->slots[c->last_slot];
  if ((rc = fetch_counters(c, s, c->stream))) return rc;
  if (num_cluster) *num_cluster = s->h_counters[CNT_NUM_CLUSTER];
  if (grid) {
    LMOT_CUDA(c, cudaMemcpyAsync(grid, s->d_label_grid, kCartCells * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return LMOT_OK;
}

int lmot_box_fit(lmot_ctx* ctx, const float* elevated, int n, int stride, const int32_t* grid, int num_cluster,
                 float* boxes, int max_boxes, int* n_boxes, float* markers) {
  if (!ctx || !grid || num_cluster < 0 || max_boxes < 0) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  if (num_cluster > c->prm.max_clusters) return LMOT_ERR_CAPACITY;
  int rc = drain(c);
  if (rc) return rc;
  Slot* s = &c->slots[0];
  c->last_slot = 0;
  cudaStream_t st = c->stream;
  if ((rc = upload_points(c, s, st, elevated, n, stride, s->d_elev))) return rc;
  LMOT_CUDA(c, cudaMemcpyAsync(s->d_label_grid, grid, kCartCells * sizeof(int), cudaMemcpyHostToDevice, st));
  s->label_grid_foreign = true;
  if ((rc = set_counter(c, s, st, CNT_N_ELEV, n))) return rc;
  if ((rc = set_counter(c, s, st, CNT_NUM_CLUSTER, num_cluster))) return rc;
  s->res = &c->results[0];
  s->res->n_kev = 0;
  if ((rc = cluster_cells_only(c, s, st, n))) return rc;
  if ((rc = boxfit_launch(c, s, st, n))) return rc;
  if ((rc = fetch_counters(c, s, st))) return rc;
  if ((rc = check_device_error(c, s, st))) return rc;
  const int nb = s->h_counters[CNT_N_BOXES];
  if (n_boxes) *n_boxes = nb;
  const int ncopy = nb < max_boxes ? nb : max_boxes;
  if (boxes && ncopy > 0) LMOT_CUDA(c, cudaMemcpyAsync(boxes, s->d_boxes, (size_t)ncopy * 24 * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (markers && ncopy > 0) LMOT_CUDA(c, cudaMemcpyAsync(markers, s->d_markers, (size_t)ncopy * 6 * sizeof(float), cudaMemcpyDeviceToHost, st));
  LMOT_CUDA(c, cudaStreamSynchronize(st));
  return nb > max_boxes ? LMOT_ERR_CAPACITY : LMOT_OK;
}

int lmot_track_step(lmot_ctx* ctx, const float* boxes, int m, double timestamp_us, double v_gps, double yaw_gps,
                    lmot_track_out* out) {
  if (!ctx || m < 0 || (m > 0 && !boxes)) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  if (m > c->prm.max_boxes) return LMOT_ERR_CAPACITY;
  int rc = drain(c);
  if (rc) return rc;
  Slot* s = &c->slots[0];
  c->last_slot = 0;
  s->res = &c->results[0];
  c->last_res = s->res;
  s->res->n_kev = 0;
  cudaStream_t st = c->stream;
  if (m > 0) LMOT_CUDA(c, cudaMemcpyAsync(s->d_boxes, boxes, (size_t)m * 24 * sizeof(float), cudaMemcpyHostToDevice, st));
  if ((rc = set_counter(c, s, st, CNT_N_BOXES, m))) return rc;
  if (c->timing) cudaEventRecord(s->res->ev[3], st);
  if (c->prm.global_frame && (rc = boxes_to_global_launch(c, s, st, s->d_boxes, s->d_counters, timestamp_us, v_gps, yaw_gps, false))) return rc;
  if ((rc = tracker_launch(c, s, st, s->d_boxes, s->d_counters, timestamp_us, v_gps, yaw_gps))) return rc;
  if (c->timing) cudaEventRecord(s->res->ev[4], st);
  if ((rc = tracker_publish(c, s->res, st))) return rc;
  LMOT_CUDA(c, cudaStreamSynchronize(st));
  if (c->timing) {    // tracker stage only: stage_ms[3], kernel_ms[0..] = predict+gate, update, spawn/output
    c->stage_ms[0] = c->stage_ms[1] = c->stage_ms[2] = 0.f;
    cudaEventElapsedTime(&c->stage_ms[3], s->res->ev[3], s->res->ev[4]);
    c->n_kernel_ms = s->res->n_kev;
    for (int i = 0; i < s->res->n_kev; ++i) cudaEventElapsedTime(&c->kernel_ms[i], i == 0 ? s->res->ev[3] : s->res->kev[i - 1], s->res->kev[i]);
  }
  const int err = s->res->h_hdr[HDR_ERROR];
  c->last_n_act = s->res->h_hdr[HDR_N_ACT];
  rc = copy_track_outputs(s->res, out);
  return err ? err : (rc ? rc : s->res->h_hdr[HDR_WARN]);
}

// ---------------------------------------------------------------------------------------------- frame pipeline
int lmot_detect_dev(lmot_ctx* ctx, const float* d_points, int n) {
  if (!ctx || n < 0 || (n > 0 && !d_points) || (reinterpret_cast<uintptr_t>(d_points) & 15u)) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  if (n > c->max_points) return LMOT_ERR_CAPACITY;
  Slot* s = acquire_slot(c);
  Result* r = acquire_result(c, true);
  LMOT_CUDA(c, cudaEventRecord(s->ev_fork, c->stream));
  LMOT_CUDA(c, cudaStreamWaitEvent(s->stream, s->ev_fork, 0));
  return submit(c, s, r, reinterpret_cast<const float4*>(d_points), n, false, 0, 0, 0);
}

int lmot_frame_dev(lmot_ctx* ctx, const float* d_points, int n, double timestamp_us, double v_gps, double yaw_gps) {
  if (!ctx || n < 0 || (n > 0 && !d_points) || (reinterpret_cast<uintptr_t>(d_points) & 15u)) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  if (n > c->max_points) return LMOT_ERR_CAPACITY;
  Slot* s = acquire_slot(c);
  Result* r = acquire_result(c, true);
  LMOT_CUDA(c, cudaEventRecord(s->ev_fork, c->stream));          // the caller's stream produced d_points
  LMOT_CUDA(c, cudaStreamWaitEvent(s->stream, s->ev_fork, 0));
  return submit(c, s, r, reinterpret_cast<const float4*>(d_points), n, true, timestamp_us, v_gps, yaw_gps);
}

int lmot_frame_submit(lmot_ctx* ctx, const float* points, int n, int stride, double timestamp_us, double v_gps, double yaw_gps) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  if (c->n_in_flight >= c->n_results) return LMOT_ERR_STATE;    // result ring full: collect first
  Slot* s = acquire_slot(c);
  Result* r = acquire_result(c, false);
  if (!r) return LMOT_ERR_STATE;
  // (Measured: end-to-end frames/s is bound by these copies -- one 1.9 MB copy per frame sustains ~39 GB/s, 49 us per frame,
  // against 41 us for the tracker chain.  A copy stream of its own changed nothing, two alternating ones made cudaMemcpyAsync
  // block the host.)
  // ZERO COPY (opt-in, LMOT_ZERO_COPY=1): a frame in page-locked host memory (lmot_pinned_alloc, cudaHostAlloc, cudaHostRegister) is
  // not copied; under unified addressing the ground kernel's bulk copies read it over PCIe straight into shared memory, so every
  // input byte crosses the link once and never touches HBM (possible whenever the launch keeps each CTA's chunk resident,
  // ground_reads_input_once).  Bit-identical results (tests/test_pipeline_gpu.py) -- but MEASURED SLOWER than the copy engine on
  // B200: SM-issued reads reach ~26 GB/s of the link (74 us per 1.9 MB frame, 13.4 k frames/s end to end) where cudaMemcpyAsync
  // sustains ~39 GB/s at this copy size (49 us, 20.4 k frames/s).  Hence off by default.  The buffer must stay untouched until
  // the frame has been collected (the same contract as the asynchronous copy).
  if (c->zero_copy && stride == 4 && n > 0 && n <= c->max_points && (reinterpret_cast<uintptr_t>(points) & 15u) == 0 && ground_reads_input_once(c, n)) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, points) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer)
      return submit(c, s, r, reinterpret_cast<const float4*>(pa.devicePointer), n, true, timestamp_us, v_gps, yaw_gps);
    cudaGetLastError();      // (pageable memory: not an error, take the copy path)
  }
  LMOT_CUDA(c, cudaStreamWaitEvent(s->stream, s->ev_trk_done, 0));
  int rc = upload_points(c, s, s->stream, points, n, stride, s->d_points);
  if (rc) return rc;
  return submit(c, s, r, s->d_points, n, true, timestamp_us, v_gps, yaw_gps);
}

int lmot_frame_collect(lmot_ctx* ctx, lmot_frame_out* out) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  if (c->n_in_flight <= 0) return LMOT_ERR_STATE;
  Result* r = &c->results[c->res_oldest];
  int guard = 0;
  while (!r->in_flight && guard++ < c->n_results) { c->res_oldest = (c->res_oldest + 1) % c->n_results; r = &c->results[c->res_oldest]; }
  if (!r->in_flight) return LMOT_ERR_STATE;
  const int rc = collect_result(c, r, out);
  r->in_flight = false;
  --c->n_in_flight;
  c->res_oldest = (c->res_oldest + 1) % c->n_results;
  return rc;
}

int lmot_frames_in_flight(lmot_ctx* ctx, int* n) {
  if (!ctx || !n) return LMOT_ERR_INVALID;
  *n = ctx->c.n_in_flight;
  return LMOT_OK;
}

// 1 if the oldest submitted frame has finished (lmot_frame_collect would not block), 0 if not, <0 on error
int lmot_frame_ready(lmot_ctx* ctx) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  if (c->n_in_flight <= 0) return 0;
  Result* r = &c->results[c->res_oldest];
  if (!r->in_flight) return 0;
  return cudaEventQuery(r->ev_done) == cudaSuccess ? 1 : 0;
}

// results of the MOST RECENT submission; older uncollected ones are dropped
int lmot_frame_fetch(lmot_ctx* ctx, lmot_frame_out* out) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  Result* r = c->last_res;
  const int rc = collect_result(c, r, out);
  for (int i = 0; i < c->n_results; ++i) c->results[i].in_flight = false;
  c->n_in_flight = 0;
  c->res_oldest = c->res_next;
  return rc;
}

int lmot_frame(lmot_ctx* ctx, const float* points, int n, int stride, double timestamp_us, double v_gps, double yaw_gps,
               lmot_frame_out* out) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  if (c->n_in_flight > 0) { int rc = drain(c); if (rc) return rc; }
  int rc = lmot_frame_submit(ctx, points, n, stride, timestamp_us, v_gps, yaw_gps);
  if (rc) return rc;
  return lmot_frame_collect(ctx, out);
}

// ---------------------------------------------------------------------------------------------- batched ticks
static int batch_common(lmot_ctx* ctx, const float* const* points, const int* n, int n_frames, int stride, bool host, bool with_tracker,
                        double ts, double v, double yaw) {
  if (!ctx || !points || !n || n_frames < 1 || n_frames > kMaxBatch) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  for (int i = 0; i < n_frames; ++i) {
    if (n[i] < 0 || (n[i] > 0 && !points[i])) return LMOT_ERR_INVALID;
    if (n[i] > c->max_points) return LMOT_ERR_CAPACITY;
    if (!host && n[i] > 0 && (reinterpret_cast<uintptr_t>(points[i]) & 15u)) return LMOT_ERR_INVALID;   // bulk copies need 16-byte alignment
  }
  if (host && c->n_in_flight >= c->n_results) return LMOT_ERR_STATE;
  int rc = batch_ensure(c, n_frames);
  if (rc) return rc;
  Result* r = acquire_result(c, !host);
  if (!r) return LMOT_ERR_STATE;
  const int b = c->next_bank;
  c->next_bank = (c->next_bank + 1) % kBatchBanks;
  Slot* k = &c->bank[b];
  Slot* s0 = &c->bslots[b * c->batch_frames];
  cudaStream_t st = s0->stream;
  // the bank's previous tick must have been consumed by the tracker; device frames: ordered after the caller's stream
  LMOT_CUDA(c, cudaStreamWaitEvent(st, k->ev_trk_done, 0));
  const float4* dp[kMaxBatch];
  if (host) {
    for (int i = 0; i < n_frames; ++i) {
      Slot* s = &c->bslots[b * c->batch_frames + i];
      if ((rc = upload_points(c, s, st, points[i], n[i], stride, s->d_points))) return rc;
      dp[i] = nullptr;
    }
  } else {
    LMOT_CUDA(c, cudaEventRecord(k->ev_fork, c->stream));
    LMOT_CUDA(c, cudaStreamWaitEvent(st, k->ev_fork, 0));
    for (int i = 0; i < n_frames; ++i) dp[i] = reinterpret_cast<const float4*>(points[i]);
  }
  return batch_submit(c, b, r, dp, n, n_frames, with_tracker, ts, v, yaw);
}

int lmot_batch_submit(lmot_ctx* ctx, const float* const* points, const int* n, int n_frames, int stride_floats, double timestamp_us,
                      double v_gps, double yaw_gps) {
  return batch_common(ctx, points, n, n_frames, stride_floats, true, true, timestamp_us, v_gps, yaw_gps);
}

int lmot_batch_dev(lmot_ctx* ctx, const float* const* d_points, const int* n, int n_frames, double timestamp_us, double v_gps, double yaw_gps) {
  return batch_common(ctx, d_points, n, n_frames, 4, false, true, timestamp_us, v_gps, yaw_gps);
}

int lmot_batch_detect_dev(lmot_ctx* ctx, const float* const* d_points, const int* n, int n_frames) {
  return batch_common(ctx, d_points, n, n_frames, 4, false, false, 0, 0, 0);
}

// ground removal + connected components of F device frames in two launches on the CALLER's stream (bench.py times the
// ground_removal + CCL roofline of the batched configuration with it; results stay on the device)
int lmot_batch_ground_ccl_dev(lmot_ctx* ctx, const float* const* d_points, const int* n, int n_frames) {
  if (!ctx || !d_points || !n || n_frames < 1 || n_frames > kMaxBatch) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  int rc = batch_ensure(c, n_frames);
  if (rc) return rc;
  if (c->n_in_flight > 0 && (rc = drain(c))) return rc;
  Slot* sl[kMaxBatch];
  const float4* pp[kMaxBatch];
  for (int i = 0; i < n_frames; ++i) {
    if (n[i] < 0 || n[i] > c->max_points || (n[i] > 0 && (!d_points[i] || (reinterpret_cast<uintptr_t>(d_points[i]) & 15u)))) return LMOT_ERR_INVALID;
    sl[i] = &c->bslots[i]; pp[i] = reinterpret_cast<const float4*>(d_points[i]);
  }
  sl[0]->res = &c->results[0];
  c->results[0].n_kev = 0;
  if ((rc = ground_launch_batch(c, sl, n_frames, pp, n, c->stream, true, false, c->fuse_ccl, c->fuse_ccl))) return rc;
  return c->fuse_ccl ? LMOT_OK : ccl_launch_batch(c, sl, n_frames, c->stream);
}

static int batch_fill(Ctx* c, Result* r, int rc0, lmot_batch_out* out, const lmot_frame_out& fo) {
  if (!out) return rc0;
  const int F = r->batch_frames;
  out->n_frames = F;
  for (int i = 0; i < kMaxBatch; ++i) {
    const bool in = i < F;
    out->n_elevated[i] = in ? r->h_frame_counts[4 * i] : 0; out->n_ground[i] = in ? r->h_frame_counts[4 * i + 1] : 0;
    out->num_cluster[i] = in ? r->h_frame_counts[4 * i + 2] : 0; out->n_boxes[i] = in ? r->h_frame_counts[4 * i + 3] : 0;
  }
  out->n_boxes_total = fo.n_boxes;
  out->tracks = fo.tracks;
  return rc0;
}

int lmot_batch_collect(lmot_ctx* ctx, lmot_batch_out* out) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  if (c->n_in_flight <= 0) return LMOT_ERR_STATE;
  Result* r = &c->results[c->res_oldest];
  int guard = 0;
  while (!r->in_flight && guard++ < c->n_results) { c->res_oldest = (c->res_oldest + 1) % c->n_results; r = &c->results[c->res_oldest]; }
  if (!r->in_flight) return LMOT_ERR_STATE;
  lmot_frame_out fo{};
  if (out) { fo.boxes = out->boxes; fo.max_boxes = out->max_boxes; fo.tracks = out->tracks; }
  const int rc = collect_result(c, r, out ? &fo : nullptr);
  r->in_flight = false;
  --c->n_in_flight;
  c->res_oldest = (c->res_oldest + 1) % c->n_results;
  return batch_fill(c, r, rc, out, fo);
}

int lmot_batch_fetch(lmot_ctx* ctx, lmot_batch_out* out) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  Result* r = c->last_res;
  lmot_frame_out fo{};
  if (out) { fo.boxes = out->boxes; fo.max_boxes = out->max_boxes; fo.tracks = out->tracks; }
  const int rc = collect_result(c, r, out ? &fo : nullptr);
  for (int i = 0; i < c->n_results; ++i) c->results[i].in_flight = false;
  c->n_in_flight = 0;
  c->res_oldest = c->res_next;
  return batch_fill(c, r, rc, out, fo);
}

int lmot_batch(lmot_ctx* ctx, const float* const* points, const int* n, int n_frames, int stride_floats, double timestamp_us, double v_gps,
               double yaw_gps, lmot_batch_out* out) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  if (c->n_in_flight > 0) { int rc = drain(c); if (rc) return rc; }
  int rc = lmot_batch_submit(ctx, points, n, n_frames, stride_floats, timestamp_us, v_gps, yaw_gps);
  if (rc) return rc;
  return lmot_batch_collect(ctx, out);
}

// ---------------------------------------------------------------------------------------------- device-side hand-over (multi-GPU)
// Several sensor streams on several GPUs feeding ONE tracker (host/shared_tracker.cpp): detection leaves its box
// list on the device, the lists travel rank to rank with NCCL, the owner folds them in WITHOUT a host bounce.
int lmot_detect_boxes_dev(lmot_ctx* ctx, const float** d_boxes, const int** d_n_boxes) {
  if (!ctx || !d_boxes || !d_n_boxes) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  Slot* s = &c->slots[c->last_slot];       // the slot of the most recent lmot_detect_dev / lmot_frame_dev
  *d_boxes = s->d_boxes;
  *d_n_boxes = s->d_counters + CNT_N_BOXES;
  return LMOT_OK;
}

int lmot_tracker_counters_dev(lmot_ctx* ctx, int** d_n_tracks) {
  if (!ctx || !d_n_tracks) return LMOT_ERR_INVALID;
  *d_n_tracks = ctx->c.d_trk_counters + CNT_N_TRACKS;
  return LMOT_OK;
}

// n_lists box lists of capacity cap_per_list boxes each (d_lists[l * cap_per_list * 24], lengths d_counts[l], all DEVICE memory,
// ordered after the work queued on the caller stream) are concatenated in list order and folded into the track table as ONE
// immUkfJpdaf step.  Asynchronous on the caller stream; results with lmot_frame_fetch.
int lmot_track_step_lists_dev(lmot_ctx* ctx, const float* d_lists, const int* d_counts, int n_lists, int cap_per_list, double timestamp_us,
                              double v_gps, double yaw_gps) {
  if (!ctx || !d_lists || !d_counts || n_lists < 1 || n_lists > kMaxBatch || cap_per_list < 1) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  int rc = batch_ensure(c, 1);              // the bank's box buffer + counters are the tracker's input
  if (rc) return rc;
  Slot* k = &c->bank[0];                    // (ordering is the caller stream's; no host synchronisation here)
  Result* r = acquire_result(c, true);
  k->res = r;
  r->n_kev = 0; r->batch_frames = 0;
  cudaStream_t st = c->stream;
  if ((rc = boxes_pack_lists_launch(c, st, d_lists, d_counts, n_lists, cap_per_list, k->d_boxes, k->d_counters))) return rc;
  if (c->prm.global_frame && (rc = boxes_to_global_launch(c, k, st, k->d_boxes, k->d_counters, timestamp_us, v_gps, yaw_gps, false))) return rc;
  if ((rc = tracker_launch(c, k, st, k->d_boxes, k->d_counters, timestamp_us, v_gps, yaw_gps))) return rc;
  if ((rc = tracker_publish(c, r, st))) return rc;
  LMOT_CUDA(c, cudaEventRecord(r->ev_done, st));
  LMOT_CUDA(c, cudaEventRecord(k->ev_trk_done, st));
  LMOT_CUDA(c, cudaEventRecord(k->ev_det_done, st));
  r->has_tracks = true;
  if (!r->in_flight) { r->in_flight = true; ++c->n_in_flight; }
  return LMOT_OK;
}

// after the table was received from another rank (NCCL broadcast into lmot_tracker_table's pointer, count already in the device
// counter lmot_tracker_counters_dev points at): rebuild the side arrays before the next step.  n = the count, as the host knows it.
int lmot_tracker_table_received(lmot_ctx* ctx, int n) {
  if (!ctx || n < 0 || n > ctx->c.prm.max_tracks) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  c->h_trk_counters[CNT_N_TRACKS] = n;
  c->act_valid = false;
  c->last_trk_res = nullptr;
  return LMOT_OK;
}

int lmot_origin_points(lmot_ctx* ctx, double timestamp_us, double v_gps, double yaw_gps, double out6[6]) {
  if (!ctx || !out6) return LMOT_ERR_INVALID;
  TrackerHost h = ctx->c.th;                       // peek: fold on a copy
  origin_points_fold(h, timestamp_us, v_gps, yaw_gps);
  out6[0] = h.egoPoint[0]; out6[1] = h.egoPoint[1]; out6[2] = h.egoPoint[2];
  out6[3] = h.egoPoint[0]; out6[4] = h.egoPoint[1]; out6[5] = h.egoPoint[2] + M_PI / 2;
  return LMOT_OK;
}

// ---------------------------------------------------------------------------------------------- tracker state
int lmot_tracker_reset(lmot_ctx* ctx) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  int rc = drain(c);
  if (rc) return rc;
  c->th = TrackerHost();
  c->act_valid = false;
  LMOT_CUDA(c, cudaMemsetAsync(c->d_trk_counters, 0, CNT_COUNT * sizeof(int), c->stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
  return LMOT_OK;
}

static int fetch_trk_counters(Ctx* c) {
  int rc = drain(c);
  if (rc) return rc;
  LMOT_CUDA(c, cudaMemcpyAsync(c->h_trk_counters, c->d_trk_counters, CNT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
  return LMOT_OK;
}

int lmot_tracker_num_tracks(lmot_ctx* ctx, int* n) {
  if (!ctx || !n) return LMOT_ERR_INVALID;
  int rc = fetch_trk_counters(&ctx->c);
  if (rc) return rc;
  *n = ctx->c.h_trk_counters[CNT_N_TRACKS];
  return LMOT_OK;
}

int lmot_tracker_table(lmot_ctx* ctx, void** dev_ptr, int* bytes_per_track, int* capacity) {
  if (!ctx) return LMOT_ERR_INVALID;
  if (dev_ptr) *dev_ptr = ctx->c.d_tracks;
  if (bytes_per_track) *bytes_per_track = (int)sizeof(TrackState);
  if (capacity) *capacity = ctx->c.prm.max_tracks;
  return LMOT_OK;
}

int lmot_tracker_set_num_tracks(lmot_ctx* ctx, int n) {
  if (!ctx || n < 0 || n > ctx->c.prm.max_tracks) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  int rc = drain(c);
  if (rc) return rc;
  c->h_trk_counters[CNT_N_TRACKS] = n;
  c->act_valid = false;
  LMOT_CUDA(c, cudaMemcpyAsync(c->d_trk_counters + CNT_N_TRACKS, c->h_trk_counters + CNT_N_TRACKS, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
  return LMOT_OK;
}

// flat per-track dump, layout shared with oracle/ref_harness.cpp (documented in DESIGN.md)
enum { D_TRACKNUM = 0, D_LIFETIME = 1, D_STATIC = 2, D_VIS = 3, D_X = 4, D_P = 24, D_MODE = 124, D_ZPRED = 127, D_S = 133,
       D_K = 145, D_BESTYAW = 175, D_BBYAW = 176, D_BBAREA = 177, D_DISTINIT = 178, D_XMERGEYAW = 179, D_INITMEAS = 180,
       D_VELON = 182, D_VELO = 183, D_BBN = 186, D_BB = 187, D_BESTBBN = 211, D_BESTBB = 212, D_TOTAL = LMOT_TRACK_DUMP_DOUBLES };

int lmot_tracker_dump(lmot_ctx* ctx, double* dumps, int cap, int* n_out) {
  if (!ctx || cap < 0) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  int rc = fetch_trk_counters(c);
  if (rc) return rc;
  const int T = c->h_trk_counters[CNT_N_TRACKS];
  if (n_out) *n_out = T;
  const int n = T < cap ? T : cap;
  if (n == 0 || !dumps) return LMOT_OK;
  std::vector<TrackState> h((size_t)n);
  LMOT_CUDA(c, cudaMemcpyAsync(h.data(), c->d_tracks, (size_t)n * sizeof(TrackState), cudaMemcpyDeviceToHost, c->stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < n; ++i) {
    const TrackState& t = h[i];
    double* d = dumps + (size_t)i * D_TOTAL;
    memset(d, 0, sizeof(double) * D_TOTAL);
    d[D_TRACKNUM] = t.trackNum; d[D_LIFETIME] = t.lifetime; d[D_STATIC] = t.isStatic; d[D_VIS] = t.isVisBB;
    for (int m = 0; m < 4; ++m) { memcpy(d + D_X + 5 * m, t.x[m], 5 * sizeof(double)); memcpy(d + D_P + 25 * m, t.P[m], 25 * sizeof(double)); }
    for (int m = 0; m < 3; ++m) {
      d[D_MODE + m] = t.modeProb[m]; d[D_ZPRED + 2 * m] = t.zPred[m][0]; d[D_ZPRED + 2 * m + 1] = t.zPred[m][1];
      memcpy(d + D_S + 4 * m, t.S[m], 4 * sizeof(double)); memcpy(d + D_K + 10 * m, t.K[m], 10 * sizeof(double));
    }
    d[D_BESTYAW] = t.bestYaw; d[D_DISTINIT] = t.distFromInit; d[D_XMERGEYAW] = t.x_merge_yaw;
    d[D_INITMEAS] = t.initMeas[0]; d[D_INITMEAS + 1] = t.initMeas[1];
    d[D_VELON] = t.nVelo; for (int k = 0; k < t.nVelo && k < 3; ++k) d[D_VELO + k] = t.velo[k];
    d[D_BBN] = t.nBBox; for (int p = 0; p < t.nBBox && p < 8; ++p) for (int q = 0; q < 3; ++q) d[D_BB + 3 * p + q] = t.BBox[p][q];
    d[D_BESTBBN] = t.nBest; for (int p = 0; p < t.nBest && p < 8; ++p) for (int q = 0; q < 3; ++q) d[D_BESTBB + 3 * p + q] = t.bestBBox[p][q];
  }
  return T > cap ? LMOT_ERR_CAPACITY : LMOT_OK;
}

int lmot_tracker_load(lmot_ctx* ctx, const double* dumps, int n, int init, double timestamp_us, double ego_velo,
                      double ego_yaw, double ego_pre_yaw, double ego_point_yaw) {
  if (!ctx || n < 0 || (n > 0 && !dumps)) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  if (n > c->prm.max_tracks) return LMOT_ERR_CAPACITY;
  int rc = drain(c);
  if (rc) return rc;
  std::vector<TrackState> h((size_t)(n > 0 ? n : 1));
  for (int i = 0; i < n; ++i) {
    TrackState& t = h[i];
    memset(&t, 0, sizeof(t));
    const double* d = dumps + (size_t)i * D_TOTAL;
    t.trackNum = (int)d[D_TRACKNUM]; t.lifetime = (int)d[D_LIFETIME]; t.isStatic = d[D_STATIC] != 0; t.isVisBB = d[D_VIS] != 0;
    for (int m = 0; m < 4; ++m) { memcpy(t.x[m], d + D_X + 5 * m, 5 * sizeof(double)); memcpy(t.P[m], d + D_P + 25 * m, 25 * sizeof(double)); }
    for (int m = 0; m < 3; ++m) {
      t.modeProb[m] = d[D_MODE + m]; t.zPred[m][0] = d[D_ZPRED + 2 * m]; t.zPred[m][1] = d[D_ZPRED + 2 * m + 1];
      memcpy(t.S[m], d + D_S + 4 * m, 4 * sizeof(double)); memcpy(t.K[m], d + D_K + 10 * m, 10 * sizeof(double));
    }
    t.bestYaw = d[D_BESTYAW]; t.distFromInit = d[D_DISTINIT]; t.x_merge_yaw = d[D_XMERGEYAW];
    t.initMeas[0] = d[D_INITMEAS]; t.initMeas[1] = d[D_INITMEAS + 1];
    t.nVelo = (int)d[D_VELON]; for (int k = 0; k < t.nVelo && k < 3; ++k) t.velo[k] = d[D_VELO + k];
    t.nBBox = (int)d[D_BBN]; for (int p = 0; p < t.nBBox && p < 8; ++p) for (int q = 0; q < 3; ++q) t.BBox[p][q] = (float)d[D_BB + 3 * p + q];
    t.nBest = (int)d[D_BESTBBN]; for (int p = 0; p < t.nBest && p < 8; ++p) for (int q = 0; q < 3; ++q) t.bestBBox[p][q] = (float)d[D_BESTBB + 3 * p + q];
  }
  if (n > 0) LMOT_CUDA(c, cudaMemcpyAsync(c->d_tracks, h.data(), (size_t)n * sizeof(TrackState), cudaMemcpyHostToDevice, c->stream));
  c->h_trk_counters[CNT_N_TRACKS] = n;
  LMOT_CUDA(c, cudaMemsetAsync(c->d_trk_counters, 0, CNT_COUNT * sizeof(int), c->stream));
  LMOT_CUDA(c, cudaMemcpyAsync(c->d_trk_counters + CNT_N_TRACKS, c->h_trk_counters + CNT_N_TRACKS, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
  c->act_valid = false;
  c->th = TrackerHost();
  c->th.init = init != 0; c->th.timestamp = timestamp_us; c->th.egoVelo = ego_velo; c->th.egoYaw = ego_yaw;
  c->th.egoPreYaw = ego_pre_yaw; c->th.egoPoint[2] = ego_point_yaw;
  // (like the oracle's ref_tracker_load, which clears egoDeltaHis_: the dead-reckoning fold restarts from (0, 0, -pi/2) and
  //  ego_point_yaw is only egoPoints_[0][2] until the next step recomputes it.  A checkpoint that must keep the accumulated ego
  //  pose restores it with lmot_tracker_set_ego.)
  return LMOT_OK;
}

// the frame-level scalars of getOriginPoints / immUkfJpdaf (imm_ukf_jpda.cpp:19-24,56-58) incl. the accumulated dead-reckoning pose
int lmot_tracker_get_ego(lmot_ctx* ctx, double ego8[8]) {
  if (!ctx || !ego8) return LMOT_ERR_INVALID;
  const TrackerHost& h = ctx->c.th;
  ego8[0] = h.init ? 1.0 : 0.0; ego8[1] = h.timestamp; ego8[2] = h.egoVelo; ego8[3] = h.egoYaw; ego8[4] = h.egoPreYaw;
  ego8[5] = h.fold[0]; ego8[6] = h.fold[1]; ego8[7] = h.fold[2];
  return LMOT_OK;
}

int lmot_tracker_set_ego(lmot_ctx* ctx, const double ego8[8]) {
  if (!ctx || !ego8) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  int rc = drain(c);
  if (rc) return rc;
  TrackerHost& h = c->th;
  h.init = ego8[0] != 0.0; h.timestamp = ego8[1]; h.egoVelo = ego8[2]; h.egoYaw = ego8[3]; h.egoPreYaw = ego8[4];
  h.fold[0] = ego8[5]; h.fold[1] = ego8[6]; h.fold[2] = ego8[7];
  h.egoPoint[0] = h.fold[0]; h.egoPoint[1] = h.fold[1]; h.egoPoint[2] = h.fold[2];
  return LMOT_OK;
}

// ---------------------------------------------------------------------------------------------- inspection
int lmot_debug_polar_grid(lmot_ctx* ctx, float* minz, float* height, float* smoothed, float* hdiff, float* hground,
                          uint8_t* isground) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  int rc = lmot_sync(ctx);
  if (rc) return rc;
  Slot* s = &c->slots[c->last_slot];
  cudaStream_t st = c->stream;
  // The fused kernel writes only what its phase 3 needs; the four intermediate grids are recomputed here from the launch's min-z
  // keys.  hGround / isGround are the fused kernel's own (stage entry points leave them in d_hg).
  if ((rc = ground_grids_debug(c, s, st))) return rc;
  const size_t b = kPolarCells * sizeof(float);
  if (minz) LMOT_CUDA(c, cudaMemcpyAsync(minz, s->d_minz, b, cudaMemcpyDeviceToHost, st));
  if (height) LMOT_CUDA(c, cudaMemcpyAsync(height, s->d_height, b, cudaMemcpyDeviceToHost, st));
  if (smoothed) LMOT_CUDA(c, cudaMemcpyAsync(smoothed, s->d_smoothed, b, cudaMemcpyDeviceToHost, st));
  if (hdiff) LMOT_CUDA(c, cudaMemcpyAsync(hdiff, s->d_hdiff, b, cudaMemcpyDeviceToHost, st));
  std::vector<float> hg, hu;
  if (hground || isground) {
    hg.resize(kPolarCells); hu.resize(kPolarCells);
    LMOT_CUDA(c, cudaMemcpyAsync(hg.data(), s->d_hg_dbg, b, cudaMemcpyDeviceToHost, st));
    LMOT_CUDA(c, cudaMemcpyAsync(hu.data(), s->d_hg, b, cudaMemcpyDeviceToHost, st));
  }
  LMOT_CUDA(c, cudaStreamSynchronize(st));
  for (int k = 0; k < kPolarCells && (hground || isground); ++k) {
    const float v = s->hg_valid ? hu[k] : hg[k];
    const bool g = !(std::isinf(v) && v < 0);
    if (hground) hground[k] = g ? v : 0.f;
    if (isground) isground[k] = g ? 1 : 0;
  }
  return LMOT_OK;
}

int lmot_debug_cell_index(lmot_ctx* ctx, int32_t* ch, int32_t* bin, int n) {
  if (!ctx || n < 0) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  int rc = lmot_sync(ctx);
  if (rc) return rc;
  Slot* s = &c->slots[c->last_slot];
  if (n > s->cur_n) return LMOT_ERR_INVALID;
  std::vector<uint16_t> cell((size_t)n);
  if ((rc = ground_cells_debug(c, s, c->stream))) return rc;   // the fused kernel keeps the ids in shared memory: recompute them
  if (n) LMOT_CUDA(c, cudaMemcpyAsync(cell.data(), s->d_cell, (size_t)n * 2, cudaMemcpyDeviceToHost, c->stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < n; ++i) {
    if (cell[i] == kNoCell) { ch[i] = -1; bin[i] = -1; }
    else { ch[i] = cell[i] / kNumBin; bin[i] = cell[i] % kNumBin; }
  }
  return LMOT_OK;
}

// diagnostic: completion times (ms, relative to the oldest frame's start) of the stage boundaries and of every kernel of the
// frames still in the result ring, pipelined submissions included.  out[frame][0..4] = ev[0..4] (start, ground, cluster, box,
// tracker), out[frame][5..5+n) = the kernels in launch order; row stride 5 + kMaxKernelEvents floats; -1 = not recorded.
int lmot_debug_timeline(lmot_ctx* ctx, float* out, int cap_frames, int* n_frames, int* row_stride) {
  if (!ctx || !out) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  int rc = lmot_sync(ctx);
  if (rc) return rc;
  const int stride = 5 + kMaxKernelEvents;
  if (row_stride) *row_stride = stride;
  int n = 0;
  Result* base = nullptr;
  for (int k = 0; k < c->n_results && n < cap_frames; ++k) {
    Result* r = &c->results[(c->res_next + k) % c->n_results];      // oldest first
    if (!r->has_tracks || r->n_kev == 0) continue;
    if (!base) base = r;
    float* o = out + (size_t)n * stride;
    for (int i = 0; i < stride; ++i) o[i] = -1.f;
    for (int i = 0; i < 5; ++i) if (cudaEventElapsedTime(&o[i], base->ev[0], r->ev[i]) != cudaSuccess) { o[i] = -1.f; cudaGetLastError(); }
    for (int i = 0; i < r->n_kev; ++i) if (cudaEventElapsedTime(&o[5 + i], base->ev[0], r->kev[i]) != cudaSuccess) { o[5 + i] = -1.f; cudaGetLastError(); }
    ++n;
  }
  if (n_frames) *n_frames = n;
  return LMOT_OK;
}

// diagnostic: %globaltimer spans (ns) of the tracker kernels of the last 32 tracker steps: out[32][8] = TA first start, TA last end,
// TB start, TB end, TC start, TC end, latest start of a TA CTA that had a track, same for TB; then out[256..271] = phase stamps of the last spawn_output_kernel (fast path), out[272..287] = of the first track's warp of the last imm_update_kernel; *next = ring position of the NEXT step (oldest entry)
int lmot_debug_tracker_trace(lmot_ctx* ctx, unsigned long long* out, int* next) {
  if (!ctx || !out) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  int rc = lmot_sync(ctx);
  if (rc) return rc;
  if (!c->d_trk_trace) return LMOT_ERR_STATE;
  // the last 32 steps, oldest first (rows of steps that never ran stay zero), then the 32 phase stamps
  for (int i = 0; i < 32; ++i) {
    const long long f = (long long)c->trk_frames - 32 + i;
    if (f < 0) { memset(out + i * 8, 0, 8 * sizeof(unsigned long long)); continue; }
    LMOT_CUDA(c, cudaMemcpy(out + i * 8, c->d_trk_trace + (size_t)(f % kTraceRows) * 8, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  }
  LMOT_CUDA(c, cudaMemcpy(out + 256, c->d_trk_trace + (size_t)kTraceRows * 8, 64 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (next) *next = 0;
  return LMOT_OK;
}

// diagnostic: switch the phase clock of ground_fused_kernel on (allocates [CTAs][8] u64) and read the last launch's stamps
int lmot_debug_phase_clock(lmot_ctx* ctx, unsigned long long* out, int cap_ctas, int* n_ctas) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  int rc = lmot_sync(ctx);
  if (rc) return rc;
  if (!c->d_phase_clock) {
    LMOT_CUDA(c, cudaMalloc(&c->d_phase_clock, (size_t)c->fused_max_ctas * 8 * sizeof(unsigned long long)));
    LMOT_CUDA(c, cudaMemset(c->d_phase_clock, 0, (size_t)c->fused_max_ctas * 8 * sizeof(unsigned long long)));
    LMOT_CUDA(c, cudaMalloc(&c->d_trk_trace, ((size_t)kTraceRows * 8 + 64) * sizeof(unsigned long long)));
    LMOT_CUDA(c, cudaMemset(c->d_trk_trace, 0, ((size_t)kTraceRows * 8 + 64) * sizeof(unsigned long long)));
    LMOT_CUDA(c, cudaMalloc(&c->d_ccl_clock, (size_t)kMaxBatch * 16 * sizeof(unsigned long long)));
    LMOT_CUDA(c, cudaMemset(c->d_ccl_clock, 0, (size_t)kMaxBatch * 16 * sizeof(unsigned long long)));
    LMOT_CUDA(c, cudaMalloc(&c->d_fit_clock, (size_t)kFitClockCtas * 8 * sizeof(unsigned long long)));
    LMOT_CUDA(c, cudaMemset(c->d_fit_clock, 0, (size_t)kFitClockCtas * 8 * sizeof(unsigned long long)));
    c->trk_frames = 0;
    if (n_ctas) *n_ctas = 0;
    return LMOT_OK;
  }
  const int g = c->last_ground_ctas < cap_ctas ? c->last_ground_ctas : cap_ctas;
  if (out && g > 0) LMOT_CUDA(c, cudaMemcpy(out, c->d_phase_clock, (size_t)g * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (n_ctas) *n_ctas = g;
  return LMOT_OK;
}

// diagnostic (after the phase clock was switched on): %globaltimer stamps of the last ccl_bitmap_kernel launch (which = 1:
// out[frames][16]) or of the last box_fit_kernel launch (which = 2: out[CTAs][8]: start, first cluster's point pass done, first
// cluster fitted, all clusters done, end)
int lmot_debug_stage_clocks(lmot_ctx* ctx, int which, unsigned long long* out, int cap_rows, int* n_rows, int* row_words) {
  if (!ctx || !out || (which != 1 && which != 2)) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  int rc = lmot_sync(ctx);
  if (rc) return rc;
  if (!c->d_ccl_clock) return LMOT_ERR_STATE;
  const int words = which == 1 ? 16 : 8;
  int rows = which == 1 ? kMaxBatch : (c->last_fit_ctas < kFitClockCtas ? c->last_fit_ctas : kFitClockCtas);
  if (rows > cap_rows) rows = cap_rows;
  if (rows > 0) LMOT_CUDA(c, cudaMemcpy(out, which == 1 ? c->d_ccl_clock : c->d_fit_clock, (size_t)rows * words * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  if (n_rows) *n_rows = rows;
  if (row_words) *row_words = words;
  return LMOT_OK;
}

int lmot_selftest_atan2f(const float* y, const float* x, int n, float* out) {
  if (n < 0 || (n > 0 && (!y || !x || !out))) return LMOT_ERR_INVALID;
  for (int i = 0; i < n; ++i) out[i] = atan2f_fdlibm(y[i], x[i]);
  return LMOT_OK;
}

int lmot_enable_timing(lmot_ctx* ctx, int on) {
  if (!ctx) return LMOT_ERR_INVALID;
  ctx->c.timing = on != 0;
  return LMOT_OK;
}

int lmot_last_kernel_ms(lmot_ctx* ctx, float* ms, int cap, int* n) {
  if (!ctx || !ms || !n) return LMOT_ERR_INVALID;
  *n = ctx->c.n_kernel_ms;
  for (int i = 0; i < ctx->c.n_kernel_ms && i < cap; ++i) ms[i] = ctx->c.kernel_ms[i];
  return LMOT_OK;
}

int lmot_debug_host_ns(lmot_ctx* ctx, double ns[4], int reset) {
  if (!ctx || !ns) return LMOT_ERR_INVALID;
  for (int i = 0; i < 4; ++i) { ns[i] = ctx->c.host_ns[i]; if (reset) ctx->c.host_ns[i] = 0; }
  return LMOT_OK;
}

int lmot_last_stage_ms(lmot_ctx* ctx, float ms[4]) {
  if (!ctx || !ms) return LMOT_ERR_INVALID;
  for (int i = 0; i < 4; ++i) ms[i] = ctx->c.stage_ms[i];
  return LMOT_OK;
}

}  // extern "C"

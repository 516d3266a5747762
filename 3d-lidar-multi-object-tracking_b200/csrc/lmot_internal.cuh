// lmot_internal.cuh -- context layout and launch prototypes shared by the stage translation units.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include "../../include/lmot.h"

namespace lmot {

constexpr int kNumChannel = LMOT_NUM_CHANNEL;
constexpr int kNumBin = LMOT_NUM_BIN;
constexpr int kPolarCells = kNumChannel * kNumBin;  // 9600
constexpr int kNumGrid = LMOT_NUM_GRID;
constexpr int kCartCells = kNumGrid * kNumGrid;      // 62500
constexpr uint16_t kNoCell = 0xFFFFu;
constexpr uint16_t kPreFiltered = 0xFFFEu;           // removed by the ground node's pre-filters (never reaches groundRemove)
constexpr int kScanTile = 1024;                      // points per tile of the stable-partition scan
constexpr int kMaxBatch = 8;                         // frames (one per sensor stream) of one batched launch, lmot_batch_*

// device-side counters of one frame (one int each; written by kernels, read by later kernels and by fetch)
enum Counter {
  CNT_N_ELEV = 0, CNT_N_GROUND, CNT_NUM_CLUSTER, CNT_N_BOXES, CNT_N_CLUSTERED, CNT_ERROR, CNT_N_TRACKS, CNT_N_VIS,
  CNT_N_ACT,        // tracks the next frame has to visit (live, or dead with a stale isVisBB_ flag): length of d_act_list
  CNT_COUNT = 16
};

struct GroundParams {
  float r_min, r_max, t_hmin, t_hmax, t_hdiff, h_sensor;
  float r_span;          // r_max - r_min evaluated in float (ground_removal.cpp:71)
  float bin_scale;       // 120 / r_span: bins per metre of the guarded fast path (ground.cu polar_cell)
  int prefilter;         // the `ground` node's PassThrough(z) + ConditionalRemoval(x, y) in front of groundRemove
  float fz0, fz1, fx0, fx1, fy0, fy1;
  double tol;            // 0.25
  double tap[3];         // gaussKernel(3, 1.0) computed on the host with the host libm (gaus_blur.cpp:26-49)
};

// one track == one UKF object of the reference (ukf.h:15-263) + its trackNumVec_ entry; x/P index 0 = merge,
// 1 = cv, 2 = ctrv, 3 = rm.  Xsig_pred_* is transient (recomputed by every Prediction) and not stored.
struct TrackState {
  double x[4][5];
  double P[4][25];
  double modeProb[3];
  double zPred[3][2];
  double S[3][4];
  double K[3][10];
  double bestYaw, distFromInit, x_merge_yaw;
  double initMeas[2];
  double velo[3];
  float BBox[8][3];
  float bestBBox[8][3];
  int trackNum, lifetime, nVelo, nBBox, nBest;
  uint8_t isStatic, isVisBB;
};

// planar rigid transform, rows (m0 m1 . m2) / (m3 m4 . m5) of a 4x4 whose third row is (0 0 1 0): lmot_params.global_frame
struct Xf2 { int on = 0; float m[6] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f}; };

// frame-level scalars of imm_ukf_jpda.cpp:19-24,56-58 (host side: a handful of doubles per frame)
struct TrackerHost {
  bool init = false;
  double timestamp = 0, egoVelo = 0, egoYaw = 0, egoPreYaw = 0;
  double egoPoint[3] = {0, 0, 0};
  double fold[3] = {0, 0, -1.5707963267948966};
};

constexpr int kMaxKernelEvents = 20;
constexpr int kTraceRows = 1024;      // diagnostic ring of per-frame tracker kernel spans (8 words each), then 32 phase stamps

// Result block: where one frame's results land.  Pinned, device-mapped host memory written by the kernels themselves
// (box compaction, spawn_output_kernel) -- the stores ARE the D2H transfer; the host waits on ev_done and reads.
// The ring of result blocks is independent of the detection slots, so the host can run many frames ahead of the GPU.
struct Result {
  int* h_hdr = nullptr;                // [16] n_elev, n_ground, num_cluster, n_boxes, n_tracks, n_vis, error
  int* h_det = nullptr;                // [CNT_COUNT] raw detection counters (detect-only submissions)
  int* h_frame_counts = nullptr;       // [kMaxBatch][4] batched ticks: n_elev, n_ground, num_cluster, n_boxes of every frame
  int batch_frames = 0;                // frames of the tick this block holds (0: a single-frame submission)
  float* h_boxes = nullptr;            // [max_boxes][24]
  float* h_targets = nullptr; double* h_vandyaw = nullptr; int* h_manage = nullptr;
  uint8_t* h_static = nullptr; uint8_t* h_vis = nullptr; float* h_visbb = nullptr;
  // device copy of the same block: written by spawn_output_kernel on the tracker stream, moved to the host block by
  // publish_kernel on the publish stream (off the tracker's sequential chain)
  unsigned char* d_block = nullptr;
  int* d_hdr = nullptr; float* d_boxes = nullptr; float* d_targets = nullptr; double* d_vandyaw = nullptr; int* d_manage = nullptr;
  uint8_t* d_static = nullptr; uint8_t* d_vis = nullptr; float* d_visbb = nullptr;
  cudaEvent_t ev_tc = nullptr;         // recorded on the tracker stream after spawn_output_kernel
  cudaEvent_t ev_done = nullptr;
  bool in_flight = false;              // submitted and not yet collected / dropped
  bool has_tracks = false;             // went through the tracker (frame) or not (detect only)
  Xf2 back;                            // global_frame: global -> sensor transform of this frame, applied by publish_kernel
  // timing (lmot_enable_timing): stage boundaries and one event after every kernel
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t kev[kMaxKernelEvents] = {};
  int n_kev = 0;
};

// Detection slot: every buffer the three detection stages (ground -> cluster -> box) touch for ONE frame, plus the
// stream they run on.  A context owns `pipeline_depth` slots so the detection stages of frame f+1.. run while the
// tracker (a sequential fold over frames, on its own stream) is still busy with frame f.
struct Slot {
  int index = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_fork = nullptr;       // recorded on the caller's stream: the frame's input is ready
  cudaEvent_t ev_det_done = nullptr;   // recorded on the slot stream after box fitting
  cudaEvent_t ev_trk_done = nullptr;   // recorded on the tracker stream when the tracker has consumed the slot

  // ---- frame input (host-buffer entry points copy here; *_dev entry points use the caller's pointer)
  float4* d_points = nullptr;
  float* d_stage_in = nullptr;         // raw staging for stride != 4 inputs
  const float4* cur_points = nullptr;
  int cur_n = 0;

  // ---- ground removal
  uint16_t* d_cell = nullptr;          // per point polar cell (ch*120+bin) or kNoCell
  unsigned* d_polar_key = nullptr;     // [2][9600] order-preserving uint key of min z; frame k uses grid k&1 and re-arms the other
  float* d_minz = nullptr;             // [9600] debug / parity
  float* d_height = nullptr;           // [9600]
  float* d_smoothed = nullptr;         // [9600]
  float* d_hdiff = nullptr;            // [9600]
  float* d_hg = nullptr;               // [9600] stage entry points: hGround of ground cells as the fused kernel evaluated it, -inf for non-ground cells
  bool hg_valid = false;               // the last launch wrote d_hg
  float* d_lim = nullptr;              // [9600] label limit of every cell (ground.cu label_limit): ground phase 2 -> phase 3
  float* d_hg_dbg = nullptr;           // [9600] the same grid recomputed as a whole by polar_grid_debug_kernel (allocated on first use)
  uint8_t* d_labels = nullptr;         // per point 0/1/2
  float4* d_elev = nullptr;            // compacted elevated cloud
  float4* d_ground = nullptr;          // compacted ground cloud
  unsigned long long* d_gdesc = nullptr;  // [2*CTAs] epoch-tagged per-CTA output counts of the fused ground kernel
  unsigned* d_gbar = nullptr;          // grid-barrier arrival counter of the fused ground kernel (never reset)
  unsigned bar_base = 0;               // host: arrivals of all earlier launches on this slot
  unsigned epoch = 0;                  // host: launches so far on this slot (tag of d_gdesc, parity of d_polar_key)
  int* d_counters = nullptr;           // [CNT_COUNT]
  int* h_counters = nullptr;           // pinned mirror
  int* h_set = nullptr;                // pinned staging for host-written counters

  // ---- clustering
  uint16_t* d_cart = nullptr;          // per elevated point cartesian cell (x*250+y) or kNoCell
  unsigned* d_cart_bits = nullptr;     // [3][2000] bit planes of the 250x250 grid (8 words per row): cells seen once | cells
                                       // seen more than once (re-armed by the CCL kernel) | occupancy of the previous frame
  int* d_label_grid = nullptr;         // [62500] final labels (0 = empty), x-major; updated sparsely from frame to frame
  // side outputs of the cluster node (allocated on first use by lmot_cluster_outputs)
  int* d_first_idx = nullptr;          // [62500] scratch: first point of each labelled cell, INT_MAX at rest
  float4* d_clustered = nullptr;       // makeClusteredCloud
  float4* d_obstacles = nullptr;       // setObsMsg: x, y, z, cluster id
  int* d_cost_map = nullptr;           // [2500] createCostMap, then the two output counts
  bool label_grid_foreign = false;     // the caller uploaded its own grid (lmot_box_fit): next clustering starts from zero

  // ---- box fitting
  uint16_t* d_pcid = nullptr;          // per elevated point cluster id (0 = none)
  int* d_table = nullptr;              // [tiles][max_clusters+1] tile histograms -> exclusive tile offsets
  int* d_seg_start = nullptr;          // [max_clusters+1] first slot of each cluster in d_sorted_idx
  int* d_seg_size = nullptr;           // [max_clusters+1] points per cluster
  float4* d_sorted_pts = nullptr;      // elevated points grouped by cluster, cloud order inside a cluster
  float* d_cl_box = nullptr;           // [max_clusters+1][24] per-cluster box (valid where d_cl_ok)
  float* d_cl_marker = nullptr;        // [max_clusters+1][6]
  uint8_t* d_cl_ok = nullptr;          // [max_clusters+1] rule filter verdict
  float* d_boxes = nullptr;            // [max_boxes][8][3] accepted boxes, cluster-id order
  float* d_boxes_g = nullptr;          // [max_boxes][8][3] the same list in the dead-reckoned global frame (lmot_params.global_frame)
  float* d_markers = nullptr;          // [max_boxes][6]
  int* d_done = nullptr;               // last-CTA-done counter
  int* d_det_sem = nullptr;            // semaphore: +1 by box_fit_kernel's last CTA (frame submissions), -1 by spawn_output_kernel

  struct Result* res = nullptr;        // result block of the frame currently (or last) processed on this slot
};

enum HostHdr { HDR_N_ELEV = 0, HDR_N_GROUND, HDR_NUM_CLUSTER, HDR_N_BOXES, HDR_N_TRACKS, HDR_N_VIS, HDR_ERROR, HDR_WARN, HDR_N_ACT, HDR_COUNT = 16 };

constexpr int kMaxSlots = 8;
constexpr int kFitClockCtas = 4096;                  // rows of the box-fitting phase clock (diagnostic)
constexpr int kBatchBanks = 2;                       // batched ticks: detection of tick t+1 overlaps the tracker of tick t
constexpr int kMaxResults = 64;

struct Ctx {
  lmot_params prm;
  int device = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;       // the caller's stream: stage-by-stage entry points, fork/join of the pipeline
  std::string last_error;
  GroundParams gp;

  // ---- capacities
  int max_points = 0, max_sort_tiles = 0, fit_ctas = 296, n_mt_raw = 0;
  bool coop_launch = false;            // LMOT_COOP=1: cudaLaunchCooperativeKernel for the ground kernel (A/B diagnostics)
  int fused_max_ctas = 0;              // co-residency limit of the cooperative ground kernel on this device
  unsigned long long* d_trk_trace = nullptr;     // diagnostic: [kTraceRows][8] per-frame kernel spans of the tracker chain + 32 phase stamps (same switch)
  unsigned long long trk_frames = 0;
  unsigned long long* d_phase_clock = nullptr;   // diagnostic: [CTAs][8] %globaltimer stamps of the last ground launch (lmot_debug_phase_clock)
  int last_ground_ctas = 0;
  unsigned long long* d_fit_clock = nullptr;     // diagnostic: [CTAs][8] stamps of the last box_fit_kernel launch (same switch)
  int last_fit_ctas = 0;
  bool fuse_ccl = true;                // frame / batch path: clustering runs in the ground kernel's tail (LMOT_FUSE_CCL=0: separate launch, A/B + phase stamps)
  unsigned long long* d_ccl_clock = nullptr;     // diagnostic: [frames][16] %globaltimer stamps of the last ccl_bitmap_kernel launch (same switch)
  bool zero_copy = false;              // lmot_frame_submit: pinned host frames are read by the ground kernel directly (LMOT_ZERO_COPY=1; off: slower than the copy engine)
  bool ground_half_sms = true;         // frame pipeline: ground kernel on half of the SMs (ground.cu ground_launch; LMOT_GROUND_HALF=0 disables, A/B only)
  unsigned spin_limit = 1u << 24;      // polls after which a device-side wait traps instead of hanging the GPU (LMOT_SPIN_LIMIT, 0 = wait for ever)
  int pts_per_cta = 768;               // target chunk of the fused ground kernel (LMOT_PTS_PER_CTA overrides, tuning only)
  unsigned long long* d_mt_raw = nullptr;  // raw mt19937_64(0) outputs (shared, read only)

  // ---- detection slots
  int n_slots = 1;
  Slot slots[kMaxSlots];
  int next_slot = 0;                   // slot of the next submission
  int last_slot = 0;                   // slot of the most recent submission (debug getters read it)
  // ---- batched ticks (lmot_batch_*): own detection slots, kBatchBanks x batch_frames, allocated on first use
  Slot bslots[kBatchBanks * kMaxBatch];
  int n_bslots = 0, batch_frames = 0, next_bank = 0;
  Slot bank[kBatchBanks];              // pseudo-slots: concatenated boxes + counters + semaphore + events of a bank

  // ---- result ring
  int n_results = 1;
  Result results[kMaxResults];
  int res_next = 0, res_oldest = 0, n_in_flight = 0;
  Result* last_res = nullptr;

  // ---- tracker
  TrackerHost th;
  cudaStream_t trk_stream = nullptr;
  cudaStream_t pub_stream = nullptr;   // device -> host publication of finished frames
  int* d_act_list = nullptr;           // [max_tracks] tracks to visit next frame (built by spawn_output_kernel)
  double4* d_pos = nullptr;            // [max_tracks] packed (x, y, yaw, -) of every track's merged state
  int* d_meas_n = nullptr;             // [max_tracks] TA -> TB: number of gated boxes of the track (-1: not provided)
  void* d_meas_ctr = nullptr;          // [max_tracks][32] double2: centre points of the first 32 gated boxes, box order
  unsigned* d_tc_seq = nullptr;        // tracker steps completed (counted by spawn_output_kernel, polled by publish_kernel)
  unsigned tc_launched = 0;            // host: tracker steps launched
  void* d_summary = nullptr;           // [max_tracks] ActSummary (tracker.cu): what TC needs of each active track, written by TB
  Result* last_trk_res = nullptr;      // result block of the previous tracker step (its device copy seeds the next one)
  bool act_valid = false;              // false after the table was written from the host: rebuilt before the next step
  int trk_ctas = 592, gate_words = 0;
  int last_n_act = 0;          // active tracks (live or visible) in the last result the host has read: picks spawn_output_kernel's variant
  bool tc_wide = false;
  int tc_force = -1;           // LMOT_TC_WIDE: -1 by count, 0 / 1 pinned
  TrackState* d_tracks = nullptr;      // [max_tracks] append-only table; dead tracks keep their slot
  int* d_trk_counters = nullptr;       // [CNT_COUNT] CNT_N_TRACKS / CNT_N_VIS / CNT_ERROR of the track table
  int* h_trk_counters = nullptr;       // pinned mirror
  unsigned* d_gate = nullptr;          // [max_tracks][gate_words] chi-square gate bits per (track, box)
  unsigned* d_setter = nullptr;        // [max_tracks][gate_words] boxes this track marks as matched
  int* d_first_setter = nullptr;       // [max_boxes] lowest track index that matched the box (INT_MAX = unmatched)
  uint8_t* d_skip = nullptr;           // [max_tracks] track did not reach measurementValidation this frame
  int* d_new_num = nullptr;            // [max_tracks] mergeOverSegmentation: largest visible container index per live track
  int* d_live_list = nullptr;          // [max_tracks] scratch: indices of live tracks
  int* d_vis_list = nullptr;           // [max_tracks] scratch: indices of tracks with a visible box

  // ---- timing
  bool timing = false;
  float stage_ms[4] = {0, 0, 0, 0};
  float kernel_ms[kMaxKernelEvents] = {};
  int n_kernel_ms = 0;
  double host_ns[4] = {0, 0, 0, 0};    // accumulated host time: submit, collect: event wait, collect: copies, calls
};

// error helper: records the CUDA error text in the context and returns LMOT_ERR_CUDA
#define LMOT_CUDA(ctx, call)                                                                     \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess) {                                                                    \
      (ctx)->last_error = std::string(#call) + ": " + cudaGetErrorString(e__);                   \
      return LMOT_ERR_CUDA;                                                                      \
    }                                                                                            \
  } while (0)

// timing mode only: mark the end of the kernel just launched on `st`
inline void kernel_mark(Ctx* c, Slot* s, cudaStream_t st) {
  if (c->timing && s->res && s->res->n_kev < kMaxKernelEvents) cudaEventRecord(s->res->kev[s->res->n_kev++], st);
}

#ifdef __CUDACC__
// mapCartesianGrid (component_clustering.cpp:38-47) + the "more than one point" test (:136) without a counter: bit planes
// `once` / `twice` of the 250x250 grid, 8 words per row.  c = x*250+y or 0xFFFF; ALL 32 lanes of the warp must call.
__device__ __forceinline__ void cart_mark(unsigned c, unsigned* __restrict__ once, unsigned* __restrict__ twice) {
  const unsigned grp = __match_any_sync(0xFFFFFFFFu, c);
  if (c != 0xFFFFu && (int)(threadIdx.x & 31) == __ffs(grp) - 1) {
    const unsigned x = c / (unsigned)kNumGrid, y = c - x * (unsigned)kNumGrid;
    const unsigned w = x * 8u + (y >> 5), bit = 1u << (y & 31u);
    if (__popc(grp) > 1) atomicOr(&twice[w], bit);
    else if (atomicOr(&once[w], bit) & bit) atomicOr(&twice[w], bit);
  }
}
#endif

// ---- stage launchers (asynchronous on the given stream) -------------------------------------------------
int ground_alloc(Ctx* c, Slot* s);
void ground_free(Slot* s);
// pts: device float4 array of n points; fuse_count: also bin the elevated points into the slot's cartesian count grid
// want_labels: also write the per-point u8 label array (stage entry point); the frame pipeline skips it
// `chain`: stand-alone entry points -- nothing of the library follows on `st` before the next ground launch; consecutive launches on
// one stream then run as programmatic dependents without an event between them (ground.cu, GroundChain)
int ground_launch(Ctx* c, Slot* s, cudaStream_t st, const float4* pts, int n, bool fuse_count = false, bool want_labels = true, bool fuse_ccl = false, bool chain = false);
int ground_launch_batch(Ctx* c, Slot* const* slots, int F, const float4* const* pts, const int* n, cudaStream_t st, bool fuse_count, bool want_labels, bool fuse_ccl = false, bool chain = false);
void ground_chain_forget(int device, bool device_idle);
int ground_cells_debug(Ctx* c, Slot* s, cudaStream_t st);
int ground_grids_debug(Ctx* c, Slot* s, cudaStream_t st);   // d_minz / d_height / d_smoothed / d_hdiff / d_hg_dbg from the last launch's keys
bool ground_reads_input_once(const Ctx* c, int n);
int ground_repack(Ctx* c, cudaStream_t st, const float* d_in, int n, int stride, float4* d_out);
int cluster_alloc(Ctx* c, Slot* s);
void cluster_free(Slot* s);
int cluster_launch(Ctx* c, Slot* s, cudaStream_t st, int n_upper, bool counted = false);
int ccl_launch_batch(Ctx* c, Slot* const* slots, int F, cudaStream_t st);
int cluster_cells_only(Ctx* c, Slot* s, cudaStream_t st, int n_upper);
int cluster_outputs_launch(Ctx* c, Slot* s, cudaStream_t st);      // makeClusteredCloud / setObsMsg / createCostMap  // d_cart for a cloud whose label grid comes from the caller
int boxfit_alloc(Ctx* c, Slot* s);
int boxfit_alloc_shared(Ctx* c);
void boxfit_free(Slot* s);
int boxfit_launch(Ctx* c, Slot* s, cudaStream_t st, int n_upper, bool post_sem = false);
int boxfit_launch_batch(Ctx* c, Slot* const* slots, int F, cudaStream_t st, const int* n_upper, bool post_sem);
int boxes_pack_lists_launch(Ctx* c, cudaStream_t st, const float* d_lists, const int* d_counts, int n_lists, int cap, float* d_boxes, int* d_counters);
int boxes_concat_launch(Ctx* c, Slot* const* slots, int F, cudaStream_t st, float* d_boxes, int* d_counters, int* d_frame_counts, int* det_sem);
void origin_points_fold(TrackerHost& h, double timestamp, double v_gps, double yaw_gps);
int tracker_alloc(Ctx* c);
void tracker_free(Ctx* c);
// boxes: device float[M][8][3] with M in det_counters[CNT_N_BOXES]; results into the slot's pinned host block
int tracker_launch(Ctx* c, Slot* s, cudaStream_t st, const float* d_boxes, const int* det_counters, double timestamp, double v_gps,
                   double yaw_gps, bool gate = false, bool* gated = nullptr);   // gate: wait for the slot's detection semaphore on the device (tracker.cu)
int boxes_to_global_launch(Ctx* c, Slot* sl, cudaStream_t st, const float* d_boxes, const int* det_counters, double timestamp, double v_gps,
                           double yaw_gps, bool post_sem);
int tracker_publish(Ctx* c, Result* r, cudaStream_t st);            // device block of r -> pinned host block
int boxes_publish(Ctx* c, Slot* s, Result* r, cudaStream_t st);     // detection-only: slot box list -> pinned host block

}  // namespace lmot

/* lmot.h -- C ABI of the B200-native LiDAR multi-object-tracking hot path (liblmot.so).
 *
 * Drop-in boundary.  The reference has no FFI; its boundary is four C++ free functions
 * called from the three ROS node callbacks.  Each entry point below replaces one of them:
 *
 *   lmot_ground_remove      <- groundRemove         object_tracking/include/ground_removal.h:62-64
 *                                                   (def. src/groundremove/ground_removal.cpp:177)
 *   lmot_component_cluster  <- componentClustering  object_tracking/include/component_clustering.h:20-22
 *                                                   (def. src/cluster/component_clustering.cpp:260)
 *   lmot_box_fit            <- boxFitting           object_tracking/include/box_fitting.h:34-36
 *                                                   (def. src/cluster/box_fitting.cpp:422)
 *   lmot_track_step         <- getOriginPoints + immUkfJpdaf
 *                                                   object_tracking/include/imm_ukf_jpda.h:15,19-22
 *                                                   (def. tracking/imm_ukf_jpda.cpp:74,704)
 *   lmot_frame              <- the four calls back to back, as object_tracking0/src/main.cpp:51-121 does in
 *                              one callback (device-resident between stages, one H2D and one D2H per frame)
 *
 * Conventions: plain pointers and sizes, caller-owned buffers, no C++/torch types.  Every function returns
 * LMOT_OK (0), a negative status (error), or a positive one (warning: the outputs are valid).  A context owns one CUDA device, one stream and all scratch memory;
 * it is NOT thread-safe (the reference's functions are non-re-entrant too); use one context per sensor stream.
 * There is no CPU fallback: without a usable CUDA device lmot_create fails with LMOT_ERR_CUDA.
 *
 * Point clouds are arrays of float with a caller-given stride in floats (3 = packed XYZ, 4 = XYZI or the
 * 16-byte pcl::PointXYZ layout the reference's PCL containers use).  Clouds RETURNED by the library are
 * always stride 4 (x, y, z, 1.0f) == pcl::PointXYZ, so a ROS shell can memcpy them into a PCL cloud.
 */
#ifndef LMOT_H
#define LMOT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LMOT_NUM_CHANNEL 80   /* ground_removal.h:16 */
#define LMOT_NUM_BIN 120      /* ground_removal.h:17 */
#define LMOT_NUM_GRID 250     /* component_clustering.h:13 */
#define LMOT_TRACK_DUMP_DOUBLES 236

typedef enum lmot_status {
  LMOT_OK = 0,
  LMOT_ERR_INVALID = -1,   /* bad argument */
  LMOT_ERR_CUDA = -2,      /* CUDA runtime/driver error, or no device (there is no CPU fallback) */
  LMOT_ERR_CAPACITY = -3,  /* input exceeds a capacity fixed at lmot_create (points, clusters, boxes, tracks) */
  LMOT_ERR_STATE = -4,     /* call sequence error (e.g. fetch before run) */
  /* positive = warning: the call did its work and every output is valid */
  LMOT_WARN_TRACK_TABLE_FULL = 1   /* max_tracks tracks exist (dead ones keep their slot, like the reference's targets_): boxes no track
                                      matched spawn no new track this frame; existing tracks are updated and reported as usual.  At the
                                      default 8192 and ~3 spawns per frame that is ~4.5 min of a 10 Hz stream; size max_tracks for the
                                      session (1.6 KB of device memory per track) or call lmot_tracker_reset between sequences. */
} lmot_status;

/* ruleBasedFilter (box_fitting.cpp:97-158) falls off its end without a return when a nested size check fails:
 * undefined behaviour.  INTENDED = fall-through means `false`; GCC13_O2_COMPAT = what g++ 13 -O2 makes of the
 * unmodified file (only the n>=30 test and the height window survive).  See SURVEY.md §8c. */
typedef enum lmot_rule_filter { LMOT_RULE_INTENDED = 0, LMOT_RULE_GCC13_O2_COMPAT = 1 } lmot_rule_filter;

/* Defaults (lmot_default_params) reproduce the reference's constants; citations are file:line under
 * /root/reference/object_tracking. */
typedef struct lmot_params {
  /* ground removal -- src/groundremove/ground_removal.cpp:24-33,239 */
  float r_min, r_max;        /* 3.4, 120 */
  float t_hmin, t_hmax;      /* -2.0, -0.4 */
  float t_hdiff;             /* 0.4 */
  float h_sensor;            /* 2.0 ("hSeonsor") */
  double ground_tolerance;   /* 0.25 */
  /* clustering -- src/cluster/component_clustering.cpp:11 */
  float roi_m;               /* 50 */
  /* box fitting -- src/cluster/box_fitting.cpp:18-44,100,308 */
  int ram_points;            /* 80 */
  int l_slope_dist;          /* 1 */
  int l_num_points;          /* 5 */
  float sensor_height;       /* 2 */
  float t_height_min, t_height_max;  /* 0.8, 2.6 */
  float t_width_min, t_width_max;    /* 0.2, 3.5 */
  float t_len_min, t_len_max;        /* 0.2, 14 */
  float t_area_max;          /* 20 */
  float t_ratio_min, t_ratio_max;    /* 1, 8 */
  float min_len_ratio;       /* 3 */
  float t_pt_per_m3;         /* 8 */
  int min_cluster_points;    /* 30 (box_fitting.cpp:100) */
  int rule_filter;           /* lmot_rule_filter, default LMOT_RULE_INTENDED */
  /* tracker -- tracking/imm_ukf_jpda.cpp:26-51,70 ; first-frame quirk :741-760 */
  int oracle_compat_first_frame; /* 1: first frame spawns ONE track from box #1 at the hard-coded (-1.5125,-8.975) */
  /* capacities (the reference's std::vectors grow without bound; fixed here, LMOT_ERR_CAPACITY when exceeded) */
  int max_points;            /* per frame, default 1<<20 */
  int max_clusters;          /* default 4096 (a 250x250 grid with 3x3 dilation cannot hold more) */
  int max_boxes;             /* default 1024 */
  int max_tracks;            /* tracks ever created (dead ones keep their slot), default 8192 */
  /* pre-filters of the reference's `ground` ROS node, applied BEFORE groundRemove (src/groundremove/main.cpp:56-89,104-112):
   * pcl::PassThrough on z (limits inclusive, non-finite points removed) then pcl::ConditionalRemoval x in (x_min,x_max) AND
   * y in (y_min,y_max) (strict).  Off by default: the hot path is groundRemove itself.  Fused into the first load. */
  int node_prefilter;        /* 0 */
  float filter_z_min, filter_z_max;   /* -3.0, 1.0  (ROS params filter_z_min / filter_z_max, main.cpp:147-148) */
  float filter_x_min, filter_x_max;   /* -15, 5     (main.cpp:68-71) */
  float filter_y_min, filter_y_max;   /* -50, 50    (main.cpp:73-76) */
  /* The `tracking` ROS node does not hand the boxes to immUkfJpdaf as they come: it moves them into a dead-reckoned "global" frame
   * first (pose of that frame in the sensor frame = egoPoints[0] of getOriginPoints, broadcast as tf velodyne -> global) and moves
   * targetPoints / visBBs back afterwards (tracking/main.cpp:76-83,142-158,182-195).  1: lmot_track_step / lmot_frame* / lmot_batch*
   * do the same on the device: boxes in, targets / vis_bb out stay in the SENSOR frame, the tracker state (and what lmot_tracker_dump
   * shows) is in the global frame.  0 (default): immUkfJpdaf as a function, boxes used as given.  The transform arithmetic is
   * pcl_ros::transformPointCloud's (float 4x4 from a double tf::Transform); tf and pcl_ros are not part of the reference tree. */
  int global_frame;          /* 0 */
  /* frames in flight inside one context: detection stages of frame f+1.. overlap the tracker of frame f (1..8, default 4) */
  int pipeline_depth;
  /* result blocks (pinned host memory) = how many submitted frames may wait to be collected (1..64, default 32) */
  int result_ring;
} lmot_params;

typedef struct lmot_ctx lmot_ctx;

int lmot_default_params(lmot_params* p);
int lmot_create(lmot_ctx** out, const lmot_params* params, int device);
void lmot_destroy(lmot_ctx* ctx);
int lmot_get_params(const lmot_ctx* ctx, lmot_params* out);   /* the parameters the context runs with (capacities after clamping) */
const char* lmot_strerror(int status);
const char* lmot_last_error(const lmot_ctx* ctx); /* text of the last CUDA error seen by this context */
const char* lmot_build_info(void);

/* Caller stream (cudaStream_t passed as void*): stage-by-stage entry points run on it, and the frame pipeline forks from /
 * joins into it (lmot_frame_dev orders a frame after the work already queued on it; lmot_flush makes it wait for the
 * pipeline).  NULL restores the context's own stream -- the legacy default stream (handle 0) cannot be selected, create a
 * stream instead (CUDA events recorded on stream 0 are not ordered with the pipeline's non-blocking streams).  The stream being
 * replaced must still be valid at this call (the library may record an event on it), and a caller stream must be replaced
 * (or the context destroyed) before that stream is destroyed. */
int lmot_set_stream(lmot_ctx* ctx, void* cuda_stream);

/* Page-locked host memory for frames handed to lmot_frame / lmot_frame_submit (a pageable buffer makes the H2D copy
 * synchronous and half as fast).  A ROS node allocates its PointCloud2 staging buffer with this; NULL without a CUDA device. */
void* lmot_pinned_alloc(size_t bytes);
void lmot_pinned_free(void* p);

/* ---- stage entry points, HOST buffers (synchronous: H2D, kernels, D2H, stream sync) -------------------- */

/* groundRemove.  labels (nullable) gets one byte per input point: 0 = in neither output (range filter /
 * cell index out of range, ground_removal.cpp:54,233), 1 = ground, 2 = elevated.  With node_prefilter on, 0 = removed by
 * the node's pre-filters and 3 = survived them (i.e. belongs to the node's `aux_points` cloud, as 1 and 2 do) but in neither
 * output of groundRemove.  elevated / ground (nullable)
 * receive the order-preserving output clouds, stride 4 floats, capacity n points each. */
int lmot_ground_remove(lmot_ctx* ctx, const float* points, int n, int stride_floats, uint8_t* labels,
                       float* elevated, int* n_elevated, float* ground, int* n_ground);

/* componentClustering.  grid = int32[250*250], x-major like array<array<int,250>,250> (written entirely;
 * the caller's zero-initialisation at src/cluster/main.cpp:72-73 is implied). */
int lmot_component_cluster(lmot_ctx* ctx, const float* elevated, int n, int stride_floats, int32_t* grid,
                           int* num_cluster);

/* The cluster node's side outputs (src/cluster/main.cpp:62-99) for the elevated cloud and label grid of the most recent
 * lmot_component_cluster / lmot_frame of this context (synchronous; call it before submitting further frames):
 *   clustered  <- makeClusteredCloud  component_clustering.cpp:308-335   float[cap][4] (cell centre x, y, -1, 1), cloud order
 *   obstacles  <- setObsMsg           component_clustering.cpp:337-375   float[cap][4] (cell centre x, y, -1, cluster id):
 *                                      ONE obstacle per labelled cell that holds a point, in order of its first point
 *   cost_map   <- createCostMap       component_clustering.cpp:425-454   int32[50*50], row = y cell, column = x cell
 * all nullable; LMOT_ERR_CAPACITY if a requested list does not fit (the counts are still returned). */
int lmot_cluster_outputs(lmot_ctx* ctx, float* clustered, int cap_clustered, int* n_clustered, float* obstacles, int cap_obstacles,
                         int* n_obstacles, int32_t* cost_map);

/* boxFitting.  boxes = float[max_boxes*8*3] (4 bottom corners z=-sensor_height, then 4 top corners z=maxZ,
 * box_fitting.cpp:379-389), in cluster-id order.  markers (nullable) = float[max_boxes*6]: centroid xyz and
 * AABB extent xyz of the cluster (mark_cluster, box_fitting.cpp:161-209). */
int lmot_box_fit(lmot_ctx* ctx, const float* elevated, int n, int stride_floats, const int32_t* grid,
                 int num_cluster, float* boxes, int max_boxes, int* n_boxes, float* markers);

/* Outputs of immUkfJpdaf: one entry per track ever created (imm_ukf_jpda.cpp:995-1041).  All pointers are
 * caller-owned with capacity `cap` entries (vis_bb: cap boxes); nullable ones are skipped. */
typedef struct lmot_track_out {
  int cap;
  int n_tracks;        /* out */
  int n_vis;           /* out: number of boxes in vis_bb */
  float* targets;      /* [cap*3]  targetPoints (x, y, -1.73/2) */
  double* vandyaw;     /* [cap*2]  targetVandYaw (v, yaw + ego yaw) */
  int32_t* track_manage; /* [cap]  trackNumVec_ */
  uint8_t* is_static;  /* [cap] */
  uint8_t* is_vis;     /* [cap] */
  float* vis_bb;       /* [cap*8*3] boxes of the tracks with is_vis, in track order */
} lmot_track_out;

/* getOriginPoints(timestamp, v_gps, yaw_gps) then immUkfJpdaf(boxes, timestamp, ...).  timestamp in
 * microseconds (dt = (ts - ts_prev)/1e6, imm_ukf_jpda.cpp:807).  boxes = float[m*8*3]. */
int lmot_track_step(lmot_ctx* ctx, const float* boxes, int m, double timestamp_us, double v_gps, double yaw_gps,
                    lmot_track_out* out);

/* getOriginPoints (imm_ukf_jpda.cpp:74-172) on its own: what the ego origin points WOULD be for this frame, without
 * advancing the tracker (lmot_track_step / lmot_frame do the same fold internally).  out6 = {x, y, yaw, x, y, yaw+pi/2}. */
int lmot_origin_points(lmot_ctx* ctx, double timestamp_us, double v_gps, double yaw_gps, double out6[6]);

typedef struct lmot_frame_out {
  int n_elevated, n_ground, num_cluster, n_boxes; /* out */
  float* boxes;        /* nullable, [max_boxes*8*3] */
  int max_boxes;
  lmot_track_out tracks;
} lmot_frame_out;

/* The whole hot path on one frame: one H2D of the XYZI frame, four device-resident stages, one D2H. */
int lmot_frame(lmot_ctx* ctx, const float* points, int n, int stride_floats, double timestamp_us, double v_gps,
               double yaw_gps, lmot_frame_out* out);

/* ---- pipelined frames (asynchronous) ----------------------------------------------------------------------
 * A context keeps up to `pipeline_depth` frames in flight: the three detection stages of a frame run on one of
 * `pipeline_depth` internal streams (own buffers each) while the tracker -- a sequential fold over frames -- runs on
 * its own stream and consumes the box lists in submission order.  Results are identical to frame-at-a-time calls.
 *
 * lmot_frame_submit  : HOST frame (pinned memory recommended) -> H2D + all four stages, returns immediately.
 *                      LMOT_ERR_STATE if result_ring frames are already waiting to be collected (the result ring is
 *                      independent of pipeline_depth, so the host may run far ahead of the GPU).
 * lmot_frame_collect : blocks until the OLDEST submitted frame is done and returns its results
 *                      (the kernels wrote them into pinned host memory; no device copy is issued here).
 * lmot_frame_dev     : like submit, but the frame is already on the device (stride 4 floats, 16-byte aligned); ordered
 *                      after the work already queued on the context's caller stream (lmot_set_stream).  Uncollected
 *                      results older than result_ring submissions are dropped.
 * lmot_frame_fetch   : waits for everything submitted so far and returns the results of the MOST RECENT frame.
 * lmot_flush         : makes the caller stream wait (on the device, not the host) for everything submitted so far, so
 *                      that CUDA events recorded on the caller stream bracket the work. */
int lmot_frame_submit(lmot_ctx* ctx, const float* points, int n, int stride_floats, double timestamp_us, double v_gps,
                      double yaw_gps);
int lmot_frame_collect(lmot_ctx* ctx, lmot_frame_out* out);
int lmot_frames_in_flight(lmot_ctx* ctx, int* n);
/* 1 if lmot_frame_collect would return without waiting, 0 if the oldest frame is still running or nothing is in flight */
int lmot_frame_ready(lmot_ctx* ctx);
int lmot_frame_dev(lmot_ctx* ctx, const float* d_points, int n, double timestamp_us, double v_gps, double yaw_gps);
int lmot_frame_fetch(lmot_ctx* ctx, lmot_frame_out* out);
int lmot_flush(lmot_ctx* ctx);
int lmot_ground_remove_dev(lmot_ctx* ctx, const float* d_points, int n);
int lmot_detect_dev(lmot_ctx* ctx, const float* d_points, int n); /* ground + cluster + box, no tracker */
int lmot_sync(lmot_ctx* ctx);

/* ---- batched ticks: several sensor streams, ONE shared track table (BASELINE.json configs[3]) ------------------------------
 * One tick = one frame from each of n_frames (<= LMOT_MAX_BATCH) sensors.  Ground removal, clustering and box fitting of all
 * frames run as ONE launch per stage (CTA groups own frames); the frames' box lists, concatenated in stream order, are the
 * measurement list of ONE immUkfJpdaf step -- what the reference does when one tracking node receives the boxes of several
 * cluster nodes in one trackbox message (tracking/main.cpp:98-141,166; the monolithic callback object_tracking0/src/main.cpp:51-121
 * per stream).  Per-frame results are identical to lmot_ground_remove / lmot_component_cluster / lmot_box_fit on that frame, the
 * track outputs to lmot_track_step on the concatenation.  Submission / collection follow lmot_frame_submit / lmot_frame_collect
 * (same result ring; ticks and single frames may be mixed, the tracker folds them in submission order). */
#define LMOT_MAX_BATCH 8
typedef struct lmot_batch_out {
  int n_frames;                                   /* out */
  int n_elevated[LMOT_MAX_BATCH], n_ground[LMOT_MAX_BATCH], num_cluster[LMOT_MAX_BATCH], n_boxes[LMOT_MAX_BATCH];   /* out, per frame */
  int n_boxes_total;                              /* out: boxes in `boxes` (frame 0's first, then frame 1's, ...) */
  float* boxes;                                   /* nullable, [max_boxes*8*3] */
  int max_boxes;
  lmot_track_out tracks;
} lmot_batch_out;
int lmot_batch_submit(lmot_ctx* ctx, const float* const* points, const int* n, int n_frames, int stride_floats, double timestamp_us,
                      double v_gps, double yaw_gps);          /* HOST frames (pinned memory recommended) */
int lmot_batch_dev(lmot_ctx* ctx, const float* const* d_points, const int* n, int n_frames, double timestamp_us, double v_gps,
                   double yaw_gps);                           /* DEVICE frames, stride 4 floats, 16-byte aligned; ordered after the caller stream */
int lmot_batch_detect_dev(lmot_ctx* ctx, const float* const* d_points, const int* n, int n_frames);   /* no tracker step */
int lmot_batch_collect(lmot_ctx* ctx, lmot_batch_out* out);   /* oldest submitted tick */
int lmot_batch_fetch(lmot_ctx* ctx, lmot_batch_out* out);     /* most recent tick; older uncollected results are dropped */
int lmot_batch(lmot_ctx* ctx, const float* const* points, const int* n, int n_frames, int stride_floats, double timestamp_us,
               double v_gps, double yaw_gps, lmot_batch_out* out);   /* submit + collect */
/* ground removal + connected components only, two launches on the caller stream, results stay on the device (roofline measurement) */
int lmot_batch_ground_ccl_dev(lmot_ctx* ctx, const float* const* d_points, const int* n, int n_frames);

/* ---- tracker state (checkpoint / teacher-forced parity tests) ------------------------------------------ */
int lmot_tracker_reset(lmot_ctx* ctx);
int lmot_tracker_num_tracks(lmot_ctx* ctx, int* n);
/* dumps = double[n*LMOT_TRACK_DUMP_DOUBLES], layout documented in DESIGN.md (same as oracle/ref_harness.cpp) */
int lmot_tracker_dump(lmot_ctx* ctx, double* dumps, int cap, int* n);
int lmot_tracker_load(lmot_ctx* ctx, const double* dumps, int n, int init, double timestamp_us, double ego_velo,
                      double ego_yaw, double ego_pre_yaw, double ego_point_yaw);
/* The frame-level state of getOriginPoints (imm_ukf_jpda.cpp:74-172), i.e. the other half of a checkpoint:
 * ego8 = {init, timestamp_us, egoVelo, egoYaw, egoPreYaw, x, y, yaw of the accumulated dead-reckoning pose}.
 * lmot_tracker_load restarts the dead reckoning at (0, 0, -pi/2) like the test oracle does; call lmot_tracker_set_ego AFTER it to
 * continue a sequence with a moving ego exactly where lmot_tracker_get_ego left it. */
int lmot_tracker_get_ego(lmot_ctx* ctx, double ego8[8]);
int lmot_tracker_set_ego(lmot_ctx* ctx, const double ego8[8]);

/* Device view of the track table for multi-GPU use (several sensor streams feeding ONE tracker): the owner
 * rank broadcasts `bytes_per_track * n` bytes from *dev_ptr with NCCL, the other ranks receive into their own table and
 * then call lmot_tracker_set_num_tracks.  The pointer stays valid for the life of the context. */
int lmot_tracker_table(lmot_ctx* ctx, void** dev_ptr, int* bytes_per_track, int* capacity);
int lmot_tracker_set_num_tracks(lmot_ctx* ctx, int n);

/* ---- device-side hand-over for several GPUs feeding ONE tracker (host/shared_tracker.cpp uses these with NCCL) ---
 * lmot_detect_boxes_dev      : device pointers to the box list (float[max_boxes*24]) and its length (int) of the most recent
 *                              lmot_detect_dev on this context -- the send buffers of the ncclAllGather
 * lmot_track_step_lists_dev  : n_lists (<= LMOT_MAX_BATCH) padded DEVICE box lists of cap_per_list boxes each, lengths in d_counts,
 *                              concatenated in list order -> ONE immUkfJpdaf step (imm_ukf_jpda.cpp:704), asynchronous on the caller
 *                              stream and ordered after the work already queued on it; results with lmot_frame_fetch
 * lmot_tracker_counters_dev  : device pointer to the number of tracks in the table (the count that travels with the table)
 * lmot_tracker_table_received: after an ncclBroadcast into lmot_tracker_table's pointer (and of the count): the table holds n tracks now */
int lmot_detect_boxes_dev(lmot_ctx* ctx, const float** d_boxes, const int** d_n_boxes);
int lmot_track_step_lists_dev(lmot_ctx* ctx, const float* d_lists, const int* d_counts, int n_lists, int cap_per_list, double timestamp_us,
                              double v_gps, double yaw_gps);
int lmot_tracker_counters_dev(lmot_ctx* ctx, int** d_n_tracks);
int lmot_tracker_table_received(lmot_ctx* ctx, int n);

/* ---- inspection of the last frame's device state (parity tests, debugging) ----------------------------- */
/* polar grid after ground removal: each float[80*120] / uint8[80*120], nullable */
int lmot_debug_polar_grid(lmot_ctx* ctx, float* minz, float* height, float* smoothed, float* hdiff, float* hground,
                          uint8_t* isground);
/* per-point polar cell of the last ground_remove: ch/bin int32[n]; -1/-1 for range-filtered points */
int lmot_debug_cell_index(lmot_ctx* ctx, int32_t* ch, int32_t* bin, int n);
/* label grid and per-elevated-point cluster id of the last clustering / box fitting */
int lmot_debug_label_grid(lmot_ctx* ctx, int32_t* grid, int* num_cluster);

/* diagnostic (after lmot_enable_timing): completion times in ms, relative to the start of the oldest frame still in the result
 * ring, of the stage boundaries [0..4] and of every kernel [5..] of each of those frames -- pipelined submissions included */
int lmot_debug_timeline(lmot_ctx* ctx, float* out, int cap_frames, int* n_frames, int* row_stride);

/* diagnostic (after the phase clock was switched on): %globaltimer spans of the tracker kernels of the last 32 tracker steps,
 * out[32][8], followed by 64 phase stamps of the last launches of the chain's kernels: out must hold 320 words */
int lmot_debug_tracker_trace(lmot_ctx* ctx, unsigned long long* out, int* next);

/* diagnostic: first call switches on the phase clock of ground_fused_kernel; later calls return the %globaltimer stamps
 * (ns) thread 0 of every CTA took at its 8 phase boundaries during the last launch: out[n_ctas][8] */
int lmot_debug_phase_clock(lmot_ctx* ctx, unsigned long long* out, int cap_ctas, int* n_ctas);

/* diagnostic (phase clock on): %globaltimer stamps of the last ccl_bitmap_kernel (which = 1, rows = frames, 16 words each) or
 * box_fit_kernel launch (which = 2, rows = CTAs, 8 words each) */
int lmot_debug_stage_clocks(lmot_ctx* ctx, int which, unsigned long long* out, int cap_rows, int* n_rows, int* row_words);

/* host-side self test of the bit-exact atan2f restatement against the host libm (no GPU needed) */
int lmot_selftest_atan2f(const float* y, const float* x, int n, float* out);

/* Per-stage device time of the last lmot_frame / lmot_frame_dev in milliseconds (cudaEvent), for bench.py:
 * ms[0] ground, ms[1] cluster, ms[2] box, ms[3] tracker.  Only recorded after lmot_enable_timing(ctx,1). */
int lmot_enable_timing(lmot_ctx* ctx, int on);
int lmot_last_stage_ms(lmot_ctx* ctx, float ms[4]);
/* finer: device time between consecutive kernels of the last timed frame, in launch order (9 kernels: ground_fused,
 * ccl_bitmap, tile_hist, seg_offsets, scatter, box_fit, imm_predict_gate, imm_update, spawn_output) */
int lmot_last_kernel_ms(lmot_ctx* ctx, float* ms, int cap, int* n);
/* accumulated host-side nanoseconds inside the library: [0] submit (launch calls), [1] collect: waiting for the frame's
 * event, [2] collect: copying results out of the pinned block, [3] number of collects */
int lmot_debug_host_ns(lmot_ctx* ctx, double ns[4], int reset);

#ifdef __cplusplus
}
#endif
#endif /* LMOT_H */

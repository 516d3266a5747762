/* lmot_shared.h -- several GPUs feeding ONE track table (liblmot_shared.so = host/shared_tracker.cpp over liblmot.so + NCCL).
 *
 * The path shards only at whole-frame / sensor-stream granularity; the one real exchange step is "several streams
 * feed ONE tracker": every rank detects on its own frame, the box lists are all-gathered (device to device, NCCL over NVLink), the
 * owner rank runs getOriginPoints + immUkfJpdaf (/root/reference/object_tracking/tracking/imm_ukf_jpda.cpp:74,704) on them and
 * broadcasts the updated track table (T x 1,648 B) so that every rank can associate its own boxes with the shared tracks.
 *   mode LMOT_SHARED_STREAMS : world sensors, one frame each per tick; boxes concatenated in rank order -> ONE tracker step
 *                              (BASELINE.json configs[3] across GPUs)
 *   mode LMOT_SHARED_FRAMES  : ONE sensor, world consecutive frames per tick, frame r on rank r; the owner folds the box lists in
 *                              frame order, one tracker step each at timestamp + r * frame_dt (the serial dependency of
 *                              imm_ukf_jpda.cpp:807,812-961 -- dt from the previous frame -- is preserved) (configs[4])
 * No box or track data crosses the host: the only host synchronisation per tick is one 4-byte read of the track count, which sizes
 * the table broadcast. */
#ifndef LMOT_SHARED_H
#define LMOT_SHARED_H
#include "lmot.h"
#ifdef __cplusplus
extern "C" {
#endif

#define LMOT_SHARED_STREAMS 0
#define LMOT_SHARED_FRAMES 1
#define LMOT_SHARED_ID_BYTES 128

typedef struct lmot_shared lmot_shared;

/* rank 0: a fresh NCCL unique id (ncclGetUniqueId) to hand to every rank out of band (file, MPI, torch.distributed store, ...) */
int lmot_shared_unique_id(void* id128);
/* collective: every rank calls it with its own context (one GPU per rank), the same id and the same owner */
int lmot_shared_create(lmot_shared** out, lmot_ctx* ctx, int rank, int world, int owner, const void* id128);
void lmot_shared_destroy(lmot_shared* s);
/* collective: one tick.  d_points: this rank's frame on ITS device (stride 4 floats, 16-byte aligned).  out (nullable): on the owner
 * the outputs of the tick's last tracker step; on the other ranks only n_tracks is set. */
int lmot_shared_tick_dev(lmot_shared* s, const float* d_points, int n, double timestamp_us, double v_gps, double yaw_gps, int mode,
                         double frame_dt_us, lmot_track_out* out);
/* device time of the last tick by CUDA events on the exchange stream, microseconds:
 * [0] detection (this rank), [1] all-gather of counts + box lists, [2] tracker step(s) (owner; 0 elsewhere), [3] broadcast of count + table */
int lmot_shared_last_us(lmot_shared* s, float us[4]);
const char* lmot_shared_last_error(const lmot_shared* s);

#ifdef __cplusplus
}
#endif
#endif

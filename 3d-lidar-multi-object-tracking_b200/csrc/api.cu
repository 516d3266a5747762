// api.cu -- the extern "C" boundary declared in include/lmot.h.  Host-side marshalling only: H2D/D2H copies,
// stream ordering and capacity checks; every algorithmic step is a CUDA kernel in ground.cu / cluster.cu /
// boxfit.cu / tracker.cu.  There is deliberately no CPU implementation of any stage here.
#include <cmath>
#include <cstring>
#include <new>
#include <vector>
#include "lmot_internal.cuh"
#include "exact_math.cuh"

using namespace lmot;

struct lmot_ctx {
  Ctx c;
};


namespace {

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

int upload_points(Ctx* c, const float* points, int n, int stride, float4* d_dst) {
  if (n < 0 || (n > 0 && !points) || stride < 3) return LMOT_ERR_INVALID;
  if (n > c->max_points) return LMOT_ERR_CAPACITY;
  if (n == 0) return LMOT_OK;
  if (stride == 4) {
    LMOT_CUDA(c, cudaMemcpyAsync(d_dst, points, (size_t)n * 16, cudaMemcpyHostToDevice, c->stream));
  } else {
    if (stride > 4) {  // wide host records: pack xyz on the host side of the copy
      std::vector<float> tmp((size_t)n * 3);
      for (int i = 0; i < n; ++i) { tmp[3*i] = points[(size_t)i*stride]; tmp[3*i+1] = points[(size_t)i*stride+1]; tmp[3*i+2] = points[(size_t)i*stride+2]; }
      LMOT_CUDA(c, cudaMemcpyAsync(c->d_stage_in, tmp.data(), (size_t)n * 12, cudaMemcpyHostToDevice, c->stream));
      LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
      stride = 3;
    } else {
      LMOT_CUDA(c, cudaMemcpyAsync(c->d_stage_in, points, (size_t)n * stride * 4, cudaMemcpyHostToDevice, c->stream));
    }
    int rc = ground_repack(c, c->d_stage_in, n, stride, d_dst);
    if (rc) return rc;
  }
  return LMOT_OK;
}

int fetch_counters(Ctx* c) {
  LMOT_CUDA(c, cudaMemcpyAsync(c->h_counters, c->d_counters, CNT_COUNT * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
  return LMOT_OK;
}

// write host-known counts into the device counter block (stage entry points that start mid-pipeline)
int set_counter(Ctx* c, int which, int value) {
  c->h_set[which] = value;
  LMOT_CUDA(c, cudaMemcpyAsync(c->d_counters + which, c->h_set + which, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  return LMOT_OK;
}

int check_device_error(Ctx* c) {
  const int e = c->h_counters[CNT_ERROR];
  if (e != 0) {
    set_counter(c, CNT_ERROR, 0);
    cudaStreamSynchronize(c->stream);
    return e;
  }
  return LMOT_OK;
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
  return LMOT_OK;
}

const char* lmot_strerror(int s) {
  switch (s) {
    case LMOT_OK: return "ok";
    case LMOT_ERR_INVALID: return "invalid argument";
    case LMOT_ERR_CUDA: return "CUDA error (no CPU fallback exists)";
    case LMOT_ERR_CAPACITY: return "capacity exceeded";
    case LMOT_ERR_STATE: return "invalid call sequence / not available";
    default: return "unknown status";
  }
}

const char* lmot_build_info(void) { return "liblmot sm_100a, nvcc " __DATE__ " -fmad=false"; }

const char* lmot_last_error(const lmot_ctx* ctx) { return ctx ? ctx->c.last_error.c_str() : ""; }

int lmot_create(lmot_ctx** out, const lmot_params* params, int device) {
  if (!out) return LMOT_ERR_INVALID;
  *out = nullptr;
  lmot_params p;
  if (params) p = *params; else lmot_default_params(&p);
  if (p.max_points <= 0 || p.max_clusters <= 0 || p.max_boxes <= 0 || p.max_tracks <= 0) return LMOT_ERR_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return LMOT_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return LMOT_ERR_CUDA;
  lmot_ctx* h = new (std::nothrow) lmot_ctx();
  if (!h) return LMOT_ERR_INVALID;
  Ctx* c = &h->c;
  c->prm = p;
  c->device = device;
  c->max_points = p.max_points;
  c->gp.r_min = p.r_min; c->gp.r_max = p.r_max; c->gp.t_hmin = p.t_hmin; c->gp.t_hmax = p.t_hmax;
  c->gp.t_hdiff = p.t_hdiff; c->gp.h_sensor = p.h_sensor;
  { volatile float span = p.r_max - p.r_min; c->gp.r_span = span; }
  c->gp.tol = p.ground_tolerance;
  gauss_taps(c->gp.tap);
  if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return LMOT_ERR_CUDA; }
  c->stream = c->own_stream;
  int rc = ground_alloc(c);
  if (rc == LMOT_OK) rc = cluster_alloc(c);
  if (rc == LMOT_OK) rc = boxfit_alloc(c);
  if (rc == LMOT_OK) rc = tracker_alloc(c);
  if (rc == LMOT_OK) rc = (cudaStreamSynchronize(c->stream) == cudaSuccess) ? LMOT_OK : LMOT_ERR_CUDA;
  if (rc != LMOT_OK) { lmot_destroy(h); return rc; }
  for (int i = 0; i < 5; ++i) cudaEventCreate(&c->ev[i]);
  *out = h;
  return LMOT_OK;
}

void lmot_destroy(lmot_ctx* ctx) {
  if (!ctx) return;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  ground_free(c);
  cluster_free(c);
  boxfit_free(c);
  tracker_free(c);
  for (int i = 0; i < 5; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete ctx;
}

int lmot_set_stream(lmot_ctx* ctx, void* s) {
  if (!ctx) return LMOT_ERR_INVALID;
  ctx->c.stream = s ? (cudaStream_t)s : ctx->c.own_stream;
  return LMOT_OK;
}

int lmot_sync(lmot_ctx* ctx) {
  if (!ctx) return LMOT_ERR_INVALID;
  LMOT_CUDA(&ctx->c, cudaStreamSynchronize(ctx->c.stream));
  return LMOT_OK;
}

int lmot_ground_remove_dev(lmot_ctx* ctx, const float* d_points, int n) {
  if (!ctx || n < 0 || (n > 0 && !d_points)) return LMOT_ERR_INVALID;
  if (n > ctx->c.max_points) return LMOT_ERR_CAPACITY;
  return ground_launch(&ctx->c, reinterpret_cast<const float4*>(d_points), n);
}

int lmot_ground_remove(lmot_ctx* ctx, const float* points, int n, int stride, uint8_t* labels, float* elevated,
                       int* n_elevated, float* ground, int* n_ground) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  int rc = upload_points(c, points, n, stride, c->d_points);
  if (rc) return rc;
  rc = ground_launch(c, c->d_points, n);
  if (rc) return rc;
  rc = fetch_counters(c);
  if (rc) return rc;
  const int ne = c->h_counters[CNT_N_ELEV], ng = c->h_counters[CNT_N_GROUND];
  if (n_elevated) *n_elevated = ne;
  if (n_ground) *n_ground = ng;
  if (labels && n > 0) LMOT_CUDA(c, cudaMemcpyAsync(labels, c->d_labels, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  if (elevated && ne > 0) LMOT_CUDA(c, cudaMemcpyAsync(elevated, c->d_elev, (size_t)ne * 16, cudaMemcpyDeviceToHost, c->stream));
  if (ground && ng > 0) LMOT_CUDA(c, cudaMemcpyAsync(ground, c->d_ground, (size_t)ng * 16, cudaMemcpyDeviceToHost, c->stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
  return LMOT_OK;
}

int lmot_component_cluster(lmot_ctx* ctx, const float* elevated, int n, int stride, int32_t* grid, int* num_cluster) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  int rc = upload_points(c, elevated, n, stride, c->d_elev);
  if (rc) return rc;
  if ((rc = set_counter(c, CNT_N_ELEV, n))) return rc;
  if ((rc = cluster_launch(c, n))) return rc;
  if ((rc = fetch_counters(c))) return rc;
  if (num_cluster) *num_cluster = c->h_counters[CNT_NUM_CLUSTER];
  if (grid) {
    LMOT_CUDA(c, cudaMemcpyAsync(grid, c->d_label_grid, kCartCells * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return LMOT_OK;
}

int lmot_debug_label_grid(lmot_ctx* ctx, int32_t* grid, int* num_cluster) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  int rc = fetch_counters(c);
  if (rc) return rc;
  if (num_cluster) *num_cluster = c->h_counters[CNT_NUM_CLUSTER];
  if (grid) {
    LMOT_CUDA(c, cudaMemcpyAsync(grid, c->d_label_grid, kCartCells * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
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
  int rc = upload_points(c, elevated, n, stride, c->d_elev);
  if (rc) return rc;
  LMOT_CUDA(c, cudaMemcpyAsync(c->d_label_grid, grid, kCartCells * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  if ((rc = set_counter(c, CNT_N_ELEV, n))) return rc;
  if ((rc = set_counter(c, CNT_NUM_CLUSTER, num_cluster))) return rc;
  if ((rc = cluster_cells_only(c, n))) return rc;
  if ((rc = boxfit_launch(c, n))) return rc;
  if ((rc = fetch_counters(c))) return rc;
  if ((rc = check_device_error(c))) return rc;
  const int nb = c->h_counters[CNT_N_BOXES];
  if (n_boxes) *n_boxes = nb;
  const int ncopy = nb < max_boxes ? nb : max_boxes;
  if (boxes && ncopy > 0) LMOT_CUDA(c, cudaMemcpyAsync(boxes, c->d_boxes, (size_t)ncopy * 24 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  if (markers && ncopy > 0) LMOT_CUDA(c, cudaMemcpyAsync(markers, c->d_markers, (size_t)ncopy * 6 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
  return nb > max_boxes ? LMOT_ERR_CAPACITY : LMOT_OK;
}

int lmot_detect_dev(lmot_ctx* ctx, const float* d_points, int n) {
  if (!ctx || n < 0 || (n > 0 && !d_points)) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  if (n > c->max_points) return LMOT_ERR_CAPACITY;
  int rc = ground_launch(c, reinterpret_cast<const float4*>(d_points), n);
  if (rc) return rc;
  if ((rc = cluster_launch(c, n))) return rc;
  return boxfit_launch(c, n);
}

// ---------------------------------------------------------------------------------------------- tracker
static int fetch_track_outputs(Ctx* c, lmot_track_out* out) {
  // h_counters already holds this frame's counters
  const int T = c->h_counters[CNT_N_TRACKS], nv = c->h_counters[CNT_N_VIS];
  if (!out) return LMOT_OK;
  out->n_tracks = T; out->n_vis = nv;
  const int n = T < out->cap ? T : out->cap;
  const int nvc = nv < out->cap ? nv : out->cap;
  if (n > 0) {
    if (out->targets) LMOT_CUDA(c, cudaMemcpyAsync(out->targets, c->d_out_targets, (size_t)n * 3 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    if (out->vandyaw) LMOT_CUDA(c, cudaMemcpyAsync(out->vandyaw, c->d_out_vandyaw, (size_t)n * 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (out->track_manage) LMOT_CUDA(c, cudaMemcpyAsync(out->track_manage, c->d_out_manage, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (out->is_static) LMOT_CUDA(c, cudaMemcpyAsync(out->is_static, c->d_out_static, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
    if (out->is_vis) LMOT_CUDA(c, cudaMemcpyAsync(out->is_vis, c->d_out_vis, (size_t)n, cudaMemcpyDeviceToHost, c->stream));
  }
  if (nvc > 0 && out->vis_bb) LMOT_CUDA(c, cudaMemcpyAsync(out->vis_bb, c->d_out_visbb, (size_t)nvc * 24 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
  return (T > out->cap) ? LMOT_ERR_CAPACITY : LMOT_OK;
}

int lmot_track_step(lmot_ctx* ctx, const float* boxes, int m, double timestamp_us, double v_gps, double yaw_gps,
                    lmot_track_out* out) {
  if (!ctx || m < 0 || (m > 0 && !boxes)) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  if (m > c->prm.max_boxes) return LMOT_ERR_CAPACITY;
  if (m > 0) LMOT_CUDA(c, cudaMemcpyAsync(c->d_boxes_in, boxes, (size_t)m * 24 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  int rc = set_counter(c, CNT_N_BOXES, m);
  if (rc) return rc;
  if ((rc = tracker_launch(c, c->d_boxes_in, timestamp_us, v_gps, yaw_gps))) return rc;
  if ((rc = fetch_counters(c))) return rc;
  if ((rc = check_device_error(c))) return rc;
  return fetch_track_outputs(c, out);
}

int lmot_frame_dev(lmot_ctx* ctx, const float* d_points, int n, double timestamp_us, double v_gps, double yaw_gps) {
  if (!ctx || n < 0 || (n > 0 && !d_points)) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  if (n > c->max_points) return LMOT_ERR_CAPACITY;
  int rc;
  if (c->timing) cudaEventRecord(c->ev[0], c->stream);
  if ((rc = ground_launch(c, reinterpret_cast<const float4*>(d_points), n))) return rc;
  if (c->timing) cudaEventRecord(c->ev[1], c->stream);
  if ((rc = cluster_launch(c, n))) return rc;
  if (c->timing) cudaEventRecord(c->ev[2], c->stream);
  if ((rc = boxfit_launch(c, n))) return rc;
  if (c->timing) cudaEventRecord(c->ev[3], c->stream);
  if ((rc = tracker_launch(c, c->d_boxes, timestamp_us, v_gps, yaw_gps))) return rc;
  if (c->timing) cudaEventRecord(c->ev[4], c->stream);
  return LMOT_OK;
}

int lmot_frame_fetch(lmot_ctx* ctx, lmot_frame_out* out) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  int rc = fetch_counters(c);
  if (rc) return rc;
  if (c->timing) for (int i = 0; i < 4; ++i) cudaEventElapsedTime(&c->stage_ms[i], c->ev[i], c->ev[i + 1]);
  if ((rc = check_device_error(c))) return rc;
  if (!out) return LMOT_OK;
  out->n_elevated = c->h_counters[CNT_N_ELEV]; out->n_ground = c->h_counters[CNT_N_GROUND];
  out->num_cluster = c->h_counters[CNT_NUM_CLUSTER]; out->n_boxes = c->h_counters[CNT_N_BOXES];
  const int nb = out->n_boxes < out->max_boxes ? out->n_boxes : out->max_boxes;
  if (out->boxes && nb > 0) LMOT_CUDA(c, cudaMemcpyAsync(out->boxes, c->d_boxes, (size_t)nb * 24 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  return fetch_track_outputs(c, &out->tracks);
}

int lmot_frame(lmot_ctx* ctx, const float* points, int n, int stride, double timestamp_us, double v_gps, double yaw_gps,
               lmot_frame_out* out) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  cudaSetDevice(c->device);
  int rc = upload_points(c, points, n, stride, c->d_points);
  if (rc) return rc;
  if ((rc = lmot_frame_dev(ctx, reinterpret_cast<const float*>(c->d_points), n, timestamp_us, v_gps, yaw_gps))) return rc;
  return lmot_frame_fetch(ctx, out);
}

int lmot_tracker_reset(lmot_ctx* ctx) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  c->th = TrackerHost();
  return set_counter(c, CNT_N_TRACKS, 0);
}

int lmot_tracker_num_tracks(lmot_ctx* ctx, int* n) {
  if (!ctx || !n) return LMOT_ERR_INVALID;
  int rc = fetch_counters(&ctx->c);
  if (rc) return rc;
  *n = ctx->c.h_counters[CNT_N_TRACKS];
  return LMOT_OK;
}

// flat per-track dump, layout shared with oracle/ref_harness.cpp (documented in DESIGN.md)
enum { D_TRACKNUM = 0, D_LIFETIME = 1, D_STATIC = 2, D_VIS = 3, D_X = 4, D_P = 24, D_MODE = 124, D_ZPRED = 127, D_S = 133,
       D_K = 145, D_BESTYAW = 175, D_BBYAW = 176, D_BBAREA = 177, D_DISTINIT = 178, D_XMERGEYAW = 179, D_INITMEAS = 180,
       D_VELON = 182, D_VELO = 183, D_BBN = 186, D_BB = 187, D_BESTBBN = 211, D_BESTBB = 212, D_TOTAL = LMOT_TRACK_DUMP_DOUBLES };

int lmot_tracker_dump(lmot_ctx* ctx, double* dumps, int cap, int* n_out) {
  if (!ctx || cap < 0) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  int rc = fetch_counters(c);
  if (rc) return rc;
  const int T = c->h_counters[CNT_N_TRACKS];
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
  if (n > 0) {
    LMOT_CUDA(c, cudaMemcpyAsync(c->d_tracks, h.data(), (size_t)n * sizeof(TrackState), cudaMemcpyHostToDevice, c->stream));
    LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  c->th = TrackerHost();
  c->th.init = init != 0; c->th.timestamp = timestamp_us; c->th.egoVelo = ego_velo; c->th.egoYaw = ego_yaw;
  c->th.egoPreYaw = ego_pre_yaw; c->th.egoPoint[2] = ego_point_yaw;
  return set_counter(c, CNT_N_TRACKS, n);
}

int lmot_debug_polar_grid(lmot_ctx* ctx, float* minz, float* height, float* smoothed, float* hdiff, float* hground,
                          uint8_t* isground) {
  if (!ctx) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  const size_t b = kPolarCells * sizeof(float);
  if (minz) LMOT_CUDA(c, cudaMemcpyAsync(minz, c->d_minz, b, cudaMemcpyDeviceToHost, c->stream));
  if (height) LMOT_CUDA(c, cudaMemcpyAsync(height, c->d_height, b, cudaMemcpyDeviceToHost, c->stream));
  if (smoothed) LMOT_CUDA(c, cudaMemcpyAsync(smoothed, c->d_smoothed, b, cudaMemcpyDeviceToHost, c->stream));
  if (hdiff) LMOT_CUDA(c, cudaMemcpyAsync(hdiff, c->d_hdiff, b, cudaMemcpyDeviceToHost, c->stream));
  std::vector<float> hg;
  if (hground || isground) {
    hg.resize(kPolarCells);
    LMOT_CUDA(c, cudaMemcpyAsync(hg.data(), c->d_hg, b, cudaMemcpyDeviceToHost, c->stream));
  }
  LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < kPolarCells && (hground || isground); ++k) {
    const bool g = !(std::isinf(hg[k]) && hg[k] < 0);
    if (hground) hground[k] = g ? hg[k] : 0.f;
    if (isground) isground[k] = g ? 1 : 0;
  }
  return LMOT_OK;
}

int lmot_debug_cell_index(lmot_ctx* ctx, int32_t* ch, int32_t* bin, int n) {
  if (!ctx || n < 0 || n > ctx->c.cur_n) return LMOT_ERR_INVALID;
  Ctx* c = &ctx->c;
  std::vector<uint16_t> cell((size_t)n);
  if (n) LMOT_CUDA(c, cudaMemcpyAsync(cell.data(), c->d_cell, (size_t)n * 2, cudaMemcpyDeviceToHost, c->stream));
  LMOT_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int i = 0; i < n; ++i) {
    if (cell[i] == kNoCell) { ch[i] = -1; bin[i] = -1; }
    else { ch[i] = cell[i] / kNumBin; bin[i] = cell[i] % kNumBin; }
  }
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

int lmot_last_stage_ms(lmot_ctx* ctx, float ms[4]) {
  if (!ctx || !ms) return LMOT_ERR_INVALID;
  for (int i = 0; i < 4; ++i) ms[i] = ctx->c.stage_ms[i];
  return LMOT_OK;
}

}  // extern "C"

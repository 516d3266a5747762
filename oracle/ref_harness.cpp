// TEST INFRASTRUCTURE ONLY -- C-ABI harness around the UNMODIFIED reference sources.
//
// Linked (by oracle/Makefile) with the six hot-path translation units of
// /root/reference/object_tracking compiled where they lie:
//   src/groundremove/{ground_removal,gaus_blur}.cpp  src/cluster/{component_clustering,box_fitting}.cpp
//   tracking/{ukf,imm_ukf_jpda}.cpp
// into oracle/_ref/libref_{o2,intended}.so.  Only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load those libraries.
//
// Every function here only marshals flat arrays into the reference's own containers and calls the
// reference entry point named in the comment; no algorithm is implemented in this file.
#include <cstdint>
#include <cstring>
#include <chrono>
#include "ground_removal.h"
#include "gaus_blur.h"
#include "component_clustering.h"
#include "box_fitting.h"
#include "ukf.h"
#include "imm_ukf_jpda.h"

// non-static helpers of ground_removal.cpp that are not declared in its header
void filterCloud(PointCloud<PointXYZ>::Ptr cloud, PointCloud<PointXYZ>& filteredCloud);
void getCellIndexFromPoints(float x, float y, int& chI, int& binI);
void applyMedianFilter(array<array<Cell, numBin>, numChannel>& polarData);
void outlierFilter(array<array<Cell, numBin>, numChannel>& polarData);

// tracker state: externally linked globals of imm_ukf_jpda.cpp:19-24,56-58,66-70
extern bool init_;
extern double timestamp_, egoVelo_, egoYaw_, egoPreYaw_;
extern int countIt;
extern vector<UKF> targets_;
extern vector<int> trackNumVec_;
extern vector<vector<double>> egoPoints_;
extern vector<vector<double>> egoDeltaHis_;
extern vector<double> egoDiffYaw_;

namespace {
// The reference prints per frame / per track (ground_removal.cpp:182,248, imm_ukf_jpda.cpp:497,609-628).  std::cout of
// this process is detached from its buffer once, at load time, and stays there (thread-safe: the reference
// arm of bench.py runs the three "nodes" on three threads).
// (a null rdbuf makes every operator<< set badbit and return; no buffer object of this library is involved)
struct QuietForever { QuietForever() { std::cout.rdbuf(nullptr); } } g_quiet_forever;
struct Quiet {};
PointCloud<PointXYZ>::Ptr make_cloud(const float* xyz, int n, int stride) {
  PointCloud<PointXYZ>::Ptr c(new PointCloud<PointXYZ>());
  c->points.resize(n);
  for (int i = 0; i < n; ++i) { c->points[i].x = xyz[i * stride]; c->points[i].y = xyz[i * stride + 1]; c->points[i].z = xyz[i * stride + 2]; }
  return c;
}
}  // namespace

extern "C" {

const char* ref_build_info() {
#ifdef LMOT_REF_INTENDED
  return "reference sources, g++ -O2, ruleBasedFilter patched copy (return false on fall-through)";
#else
  return "reference sources UNMODIFIED, g++ -O2";
#endif
}

// ground_removal.cpp:67-76 getCellIndexFromPoints
void ref_cell_index(const float* xy, int n, int stride, int* ch, int* bin) {
  for (int i = 0; i < n; ++i) getCellIndexFromPoints(xy[i * stride], xy[i * stride + 1], ch[i], bin[i]);
}

// ground_removal.cpp:177 groundRemove.  Outputs: compacted clouds (xyz, stride 3).
void ref_ground_remove(const float* xyz, int n, int stride, float* elev, int* n_elev, float* ground, int* n_ground) {
  Quiet q;
  PointCloud<PointXYZ>::Ptr cloud = make_cloud(xyz, n, stride);
  PointCloud<PointXYZ>::Ptr e(new PointCloud<PointXYZ>()), g(new PointCloud<PointXYZ>());
  groundRemove(cloud, e, g);
  *n_elev = (int)e->size(); *n_ground = (int)g->size();
  for (size_t i = 0; i < e->size(); ++i) { elev[3*i] = (*e)[i].x; elev[3*i+1] = (*e)[i].y; elev[3*i+2] = (*e)[i].z; }
  for (size_t i = 0; i < g->size(); ++i) { ground[3*i] = (*g)[i].x; ground[3*i+1] = (*g)[i].y; ground[3*i+2] = (*g)[i].z; }
}

// The polar grid after each stage, produced by calling the reference's own stage functions in the
// order groundRemove does (ground_removal.cpp:185-218).  For kernel-level debugging of the CUDA path.
// out arrays are [80*120]: minz, height, smoothed, hdiff, hground, isground
void ref_polar_grid(const float* xyz, int n, int stride, float* minz, float* height, float* smoothed, float* hdiff,
                    float* hground, uint8_t* isground) {
  Quiet q;
  PointCloud<PointXYZ>::Ptr cloud = make_cloud(xyz, n, stride);
  PointCloud<PointXYZ> filtered;
  filterCloud(cloud, filtered);
  static array<array<Cell, numBin>, numChannel> polar;
  polar = array<array<Cell, numBin>, numChannel>();
  createAndMapPolarGrid(filtered, polar);
  for (int c = 0; c < numChannel; ++c) {
    for (int b = 0; b < numBin; ++b) {
      float zi = polar[c][b].getMinZ();
      minz[c * numBin + b] = zi;
      if (zi > tHmin && zi < tHmax) polar[c][b].updataHeight(zi);
      else if (zi > tHmax) polar[c][b].updataHeight(hSeonsor);
      else polar[c][b].updataHeight(tHmin);
    }
    gaussSmoothen(polar[c], 1, 3);
    computeHDiffAdjacentCell(polar[c]);
    for (int b = 0; b < numBin; ++b) {
      if (polar[c][b].getSmoothed() < tHmax && polar[c][b].getHDiff() < tHDiff) polar[c][b].updateGround();
      else if (polar[c][b].getHeight() < tHmax && polar[c][b].getHDiff() < tHDiff) polar[c][b].updateGround();
    }
  }
  applyMedianFilter(polar);
  outlierFilter(polar);
  for (int c = 0; c < numChannel; ++c)
    for (int b = 0; b < numBin; ++b) {
      int k = c * numBin + b;
      height[k] = polar[c][b].getHeight(); smoothed[k] = polar[c][b].getSmoothed(); hdiff[k] = polar[c][b].getHDiff();
      isground[k] = polar[c][b].isThisGround() ? 1 : 0;
      hground[k] = polar[c][b].isThisGround() ? polar[c][b].getHGround() : 0.f;
    }
}

// component_clustering.cpp:260 componentClustering.  grid = int32[250*250] (x-major), zeroed by us like the caller does.
void ref_component_clustering(const float* xyz, int n, int stride, int32_t* grid, int* num_cluster) {
  Quiet q;
  PointCloud<PointXYZ>::Ptr cloud = make_cloud(xyz, n, stride);
  static array<array<int, numGrid>, numGrid> cart;
  cart = array<array<int, numGrid>, numGrid>{};
  int nc = 0;
  componentClustering(cloud, cart, nc);
  *num_cluster = nc;
  for (int x = 0; x < numGrid; ++x) for (int y = 0; y < numGrid; ++y) grid[x * numGrid + y] = cart[x][y];
}

// component_clustering.cpp:308-335 makeClusteredCloud, :337-375 setObsMsg, :425-454 createCostMap -- the cluster node's
// side outputs (src/cluster/main.cpp:62-99).  clustered = float[cap*3]; obstacles = double[cap*4] (x, y, z, cluster);
// cost_map = int32[50*50].
void ref_cluster_outputs(const float* xyz, int n, int stride, const int32_t* grid, int cap, float* clustered, int* n_clustered,
                         double* obstacles, int* n_obstacles, int32_t* cost_map) {
  Quiet q;
  PointCloud<PointXYZ>::Ptr cloud = make_cloud(xyz, n, stride);
  static array<array<int, numGrid>, numGrid> cart;
  for (int x = 0; x < numGrid; ++x) for (int y = 0; y < numGrid; ++y) cart[x][y] = grid[x * numGrid + y];
  PointCloud<PointXYZ>::Ptr cc(new PointCloud<PointXYZ>());
  makeClusteredCloud(cloud, cart, cc);
  *n_clustered = (int)cc->points.size();
  for (int i = 0; i < *n_clustered && i < cap; ++i) { clustered[3*i] = cc->points[i].x; clustered[3*i+1] = cc->points[i].y; clustered[3*i+2] = cc->points[i].z; }
  object_tracking::ObstacleList ol;
  setObsMsg(cloud, cart, ol);
  *n_obstacles = (int)ol.obstacles.size();
  for (int i = 0; i < *n_obstacles && i < cap; ++i) {
    obstacles[4*i] = ol.obstacles[i].x; obstacles[4*i+1] = ol.obstacles[i].y; obstacles[4*i+2] = ol.obstacles[i].z; obstacles[4*i+3] = ol.obstacles[i].cluster;
  }
  std::vector<int> cm = createCostMap(*cloud);
  for (size_t i = 0; i < cm.size() && i < 2500; ++i) cost_map[i] = cm[i];
}

// box_fitting.cpp:422 boxFitting.  boxes = float[n_boxes*8*3]; markers = float[n_boxes*6] (centroid xyz, scale xyz)
void ref_box_fitting(const float* xyz, int n, int stride, const int32_t* grid, int num_cluster, int max_boxes,
                     float* boxes, int* n_boxes, float* markers) {
  Quiet q;
  PointCloud<PointXYZ>::Ptr cloud = make_cloud(xyz, n, stride);
  static array<array<int, numGrid>, numGrid> cart;
  for (int x = 0; x < numGrid; ++x) for (int y = 0; y < numGrid; ++y) cart[x][y] = grid[x * numGrid + y];
  visualization_msgs::MarkerArray ma;
  vector<PointCloud<PointXYZ>> bb = boxFitting(cloud, cart, num_cluster, ma);
  *n_boxes = (int)bb.size();
  for (int b = 0; b < (int)bb.size() && b < max_boxes; ++b) {
    for (int p = 0; p < 8; ++p) { boxes[(b*8+p)*3] = bb[b][p].x; boxes[(b*8+p)*3+1] = bb[b][p].y; boxes[(b*8+p)*3+2] = bb[b][p].z; }
    if (markers) {
      const visualization_msgs::Marker& m = ma.markers[b];
      markers[b*6] = (float)m.pose.position.x; markers[b*6+1] = (float)m.pose.position.y; markers[b*6+2] = (float)m.pose.position.z;
      markers[b*6+3] = (float)m.scale.x; markers[b*6+4] = (float)m.scale.y; markers[b*6+5] = (float)m.scale.z;
    }
  }
}

// ---------------------------------------------------------------- tracker (imm_ukf_jpda.cpp:704 immUkfJpdaf)
void ref_tracker_reset() {
  targets_.clear(); trackNumVec_.clear(); egoPoints_.clear(); egoDeltaHis_.clear(); egoDiffYaw_.clear();
  init_ = false; timestamp_ = 0; egoVelo_ = 0; egoYaw_ = 0; egoPreYaw_ = 0; countIt = 0;
}

int ref_tracker_num_tracks() { return (int)targets_.size(); }

// one frame: getOriginPoints (imm_ukf_jpda.cpp:74) then immUkfJpdaf (:704), as tracking/main.cpp:74,166 does.
// boxes = float[m*8*3].  Outputs sized for n_tracks_out entries (query ref_tracker_num_tracks after the call):
//   targets float[T*3], vandyaw double[T*2], track_manage int[T], is_static u8[T], is_vis u8[T], vis_bb float[nvis*8*3]
void ref_tracker_step(const float* boxes, int m, double timestamp, double v_gps, double yaw_gps, int cap,
                      float* targets, double* vandyaw, int* track_manage, uint8_t* is_static, uint8_t* is_vis,
                      float* vis_bb, int* n_vis, int* n_tracks_out) {
  Quiet q;
  vector<PointCloud<PointXYZ>> bBoxes(m);
  for (int b = 0; b < m; ++b)
    for (int p = 0; p < 8; ++p) bBoxes[b].push_back(PointXYZ(boxes[(b*8+p)*3], boxes[(b*8+p)*3+1], boxes[(b*8+p)*3+2]));
  vector<vector<double>> origin;
  getOriginPoints(timestamp, origin, v_gps, yaw_gps);
  PointCloud<PointXYZ> tg; vector<vector<double>> vy; vector<int> tm; vector<bool> st, vis; vector<PointCloud<PointXYZ>> vbb;
  immUkfJpdaf(bBoxes, timestamp, tg, vy, tm, st, vis, vbb);
  int T = (int)tm.size();
  *n_tracks_out = T; *n_vis = (int)vbb.size();
  for (int i = 0; i < T && i < cap; ++i) {
    if (i < (int)tg.size()) { targets[3*i] = tg[i].x; targets[3*i+1] = tg[i].y; targets[3*i+2] = tg[i].z; }
    if (i < (int)vy.size()) { vandyaw[2*i] = vy[i][0]; vandyaw[2*i+1] = vy[i][1]; }
    track_manage[i] = tm[i];
    is_static[i] = (i < (int)st.size() && st[i]) ? 1 : 0;
    is_vis[i] = (i < (int)vis.size() && vis[i]) ? 1 : 0;
  }
  for (int b = 0; b < (int)vbb.size() && b < cap; ++b)
    for (int p = 0; p < 8 && p < (int)vbb[b].size(); ++p) { vis_bb[(b*8+p)*3] = vbb[b][p].x; vis_bb[(b*8+p)*3+1] = vbb[b][p].y; vis_bb[(b*8+p)*3+2] = vbb[b][p].z; }
}

// flat dump of one track (layout shared with include/lmot.h LMOT_TRACK_DUMP_DOUBLES = 236)
enum { D_TRACKNUM = 0, D_LIFETIME = 1, D_STATIC = 2, D_VIS = 3, D_X = 4, D_P = 24, D_MODE = 124, D_ZPRED = 127, D_S = 133,
       D_K = 145, D_BESTYAW = 175, D_BBYAW = 176, D_BBAREA = 177, D_DISTINIT = 178, D_XMERGEYAW = 179, D_INITMEAS = 180,
       D_VELON = 182, D_VELO = 183, D_BBN = 186, D_BB = 187, D_BESTBBN = 211, D_BESTBB = 212, D_TOTAL = 236 };

static void put_mat(double* o, const MatrixXd& m, int r, int c) {
  for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) o[i * c + j] = (m.rows() == r && m.cols() == c) ? m(i, j) : 0.0;
}
static void get_mat(const double* o, MatrixXd& m, int r, int c) {
  m = MatrixXd(r, c);
  for (int i = 0; i < r; ++i) for (int j = 0; j < c; ++j) m(i, j) = o[i * c + j];
}

void ref_tracker_dump(int i, double* d) {
  memset(d, 0, sizeof(double) * D_TOTAL);
  UKF& u = targets_[i];
  d[D_TRACKNUM] = trackNumVec_[i]; d[D_LIFETIME] = u.lifetime_; d[D_STATIC] = u.isStatic_; d[D_VIS] = u.isVisBB_;
  put_mat(d + D_X, u.x_merge_, 5, 1); put_mat(d + D_X + 5, u.x_cv_, 5, 1); put_mat(d + D_X + 10, u.x_ctrv_, 5, 1); put_mat(d + D_X + 15, u.x_rm_, 5, 1);
  put_mat(d + D_P, u.P_merge_, 5, 5); put_mat(d + D_P + 25, u.P_cv_, 5, 5); put_mat(d + D_P + 50, u.P_ctrv_, 5, 5); put_mat(d + D_P + 75, u.P_rm_, 5, 5);
  d[D_MODE] = u.modeProbCV_; d[D_MODE + 1] = u.modeProbCTRV_; d[D_MODE + 2] = u.modeProbRM_;
  for (int k = 0; k < 2; ++k) { d[D_ZPRED + k] = u.zPredCVl_(k); d[D_ZPRED + 2 + k] = u.zPredCTRVl_(k); d[D_ZPRED + 4 + k] = u.zPredRMl_(k); }
  put_mat(d + D_S, u.lS_cv_, 2, 2); put_mat(d + D_S + 4, u.lS_ctrv_, 2, 2); put_mat(d + D_S + 8, u.lS_rm_, 2, 2);
  put_mat(d + D_K, u.K_cv_, 5, 2); put_mat(d + D_K + 10, u.K_ctrv_, 5, 2); put_mat(d + D_K + 20, u.K_rm_, 5, 2);
  d[D_BESTYAW] = u.bestYaw_; d[D_BBYAW] = u.bb_yaw_; d[D_BBAREA] = u.bb_area_; d[D_DISTINIT] = u.distFromInit_; d[D_XMERGEYAW] = u.x_merge_yaw_;
  d[D_INITMEAS] = u.initMeas_(0); d[D_INITMEAS + 1] = u.initMeas_(1);
  d[D_VELON] = (double)u.velo_history_.size();
  for (size_t k = 0; k < u.velo_history_.size() && k < 3; ++k) d[D_VELO + k] = u.velo_history_[k];
  d[D_BBN] = (double)u.BBox_.size();
  for (size_t p = 0; p < u.BBox_.size() && p < 8; ++p) { d[D_BB + 3*p] = u.BBox_[p].x; d[D_BB + 3*p + 1] = u.BBox_[p].y; d[D_BB + 3*p + 2] = u.BBox_[p].z; }
  d[D_BESTBBN] = (double)u.bestBBox_.size();
  for (size_t p = 0; p < u.bestBBox_.size() && p < 8; ++p) { d[D_BESTBB + 3*p] = u.bestBBox_[p].x; d[D_BESTBB + 3*p + 1] = u.bestBBox_[p].y; d[D_BESTBB + 3*p + 2] = u.bestBBox_[p].z; }
}

// teacher forcing: overwrite the reference's track table and frame globals from a dump
void ref_tracker_load(int n_tracks, const double* dumps, int init, double timestamp, double ego_velo, double ego_yaw,
                      double ego_pre_yaw, double ego_point_yaw) {
  targets_.clear(); trackNumVec_.clear();
  for (int i = 0; i < n_tracks; ++i) {
    const double* d = dumps + (size_t)i * D_TOTAL;
    UKF u;
    VectorXd z(2); z << d[D_X], d[D_X + 1];
    u.Initialize(z, timestamp);  // sets the sigma weights; everything else is overwritten below
    u.lifetime_ = (int)d[D_LIFETIME]; u.isStatic_ = d[D_STATIC] != 0; u.isVisBB_ = d[D_VIS] != 0;
    get_mat(d + D_X, u.x_merge_, 5, 1); get_mat(d + D_X + 5, u.x_cv_, 5, 1); get_mat(d + D_X + 10, u.x_ctrv_, 5, 1); get_mat(d + D_X + 15, u.x_rm_, 5, 1);
    get_mat(d + D_P, u.P_merge_, 5, 5); get_mat(d + D_P + 25, u.P_cv_, 5, 5); get_mat(d + D_P + 50, u.P_ctrv_, 5, 5); get_mat(d + D_P + 75, u.P_rm_, 5, 5);
    u.modeProbCV_ = d[D_MODE]; u.modeProbCTRV_ = d[D_MODE + 1]; u.modeProbRM_ = d[D_MODE + 2];
    for (int k = 0; k < 2; ++k) { u.zPredCVl_(k) = d[D_ZPRED + k]; u.zPredCTRVl_(k) = d[D_ZPRED + 2 + k]; u.zPredRMl_(k) = d[D_ZPRED + 4 + k]; }
    get_mat(d + D_S, u.lS_cv_, 2, 2); get_mat(d + D_S + 4, u.lS_ctrv_, 2, 2); get_mat(d + D_S + 8, u.lS_rm_, 2, 2);
    get_mat(d + D_K, u.K_cv_, 5, 2); get_mat(d + D_K + 10, u.K_ctrv_, 5, 2); get_mat(d + D_K + 20, u.K_rm_, 5, 2);
    u.bestYaw_ = d[D_BESTYAW]; u.bb_yaw_ = d[D_BBYAW]; u.bb_area_ = d[D_BBAREA]; u.distFromInit_ = d[D_DISTINIT]; u.x_merge_yaw_ = d[D_XMERGEYAW];
    u.initMeas_(0) = d[D_INITMEAS]; u.initMeas_(1) = d[D_INITMEAS + 1];
    u.velo_history_.clear();
    for (int k = 0; k < (int)d[D_VELON]; ++k) u.velo_history_.push_back(d[D_VELO + k]);
    u.BBox_.clear(); u.bestBBox_.clear();
    for (int p = 0; p < (int)d[D_BBN]; ++p) u.BBox_.push_back(PointXYZ((float)d[D_BB + 3*p], (float)d[D_BB + 3*p + 1], (float)d[D_BB + 3*p + 2]));
    for (int p = 0; p < (int)d[D_BESTBBN]; ++p) u.bestBBox_.push_back(PointXYZ((float)d[D_BESTBB + 3*p], (float)d[D_BESTBB + 3*p + 1], (float)d[D_BESTBB + 3*p + 2]));
    targets_.push_back(u);
    trackNumVec_.push_back((int)d[D_TRACKNUM]);
  }
  init_ = init != 0; timestamp_ = timestamp; egoVelo_ = ego_velo; egoYaw_ = ego_yaw; egoPreYaw_ = ego_pre_yaw;
  egoPoints_.clear(); egoDeltaHis_.clear(); egoDiffYaw_.clear();
  vector<double> ep; ep.push_back(0); ep.push_back(0); ep.push_back(ego_point_yaw); egoPoints_.push_back(ep);
}

// ---------------------------------------------------------------- CPU-baseline timing helpers (bench.py --impl reference)
// Runs the four reference entry points on one frame and returns per-stage seconds (steady_clock).
// boxes_out receives the boxes (for feeding ref_tracker_step).  The tracker stage is timed separately by the caller.
void ref_time_detect(const float* xyz, int n, int stride, double* sec3, int max_boxes, float* boxes, int* n_boxes) {
  Quiet q;
  using clk = std::chrono::steady_clock;
  PointCloud<PointXYZ>::Ptr cloud = make_cloud(xyz, n, stride);
  PointCloud<PointXYZ>::Ptr e(new PointCloud<PointXYZ>()), g(new PointCloud<PointXYZ>());
  auto t0 = clk::now();
  groundRemove(cloud, e, g);
  auto t1 = clk::now();
  static array<array<int, numGrid>, numGrid> cart;
  cart = array<array<int, numGrid>, numGrid>{};
  int nc = 0;
  componentClustering(e, cart, nc);
  auto t2 = clk::now();
  visualization_msgs::MarkerArray ma;
  vector<PointCloud<PointXYZ>> bb = boxFitting(e, cart, nc, ma);
  auto t3 = clk::now();
  sec3[0] = std::chrono::duration<double>(t1 - t0).count();
  sec3[1] = std::chrono::duration<double>(t2 - t1).count();
  sec3[2] = std::chrono::duration<double>(t3 - t2).count();
  *n_boxes = (int)bb.size();
  for (int b = 0; b < (int)bb.size() && b < max_boxes; ++b)
    for (int p = 0; p < 8; ++p) { boxes[(b*8+p)*3] = bb[b][p].x; boxes[(b*8+p)*3+1] = bb[b][p].y; boxes[(b*8+p)*3+2] = bb[b][p].z; }
}

}  // extern "C"

// TEST INFRASTRUCTURE ONLY -- shared declarations of the plain-C++ restatement (oracle/port/*.cpp).
#pragma once
#include <cstdint>
#include <vector>

namespace port {

// one track == one UKF object of the reference (ukf.h:15-263); x/P index 0 = merge, 1 = cv, 2 = ctrv, 3 = rm
struct Track {
  double x[4][5];
  double P[4][25];
  double modeProb[3];
  double zPred[3][2];
  double S[3][4];
  double K[3][10];
  int lifetime;
  bool isStatic, isVisBB;
  float BBox[8][3]; int nBBox;
  float bestBBox[8][3]; int nBest;
  double bestYaw, bb_yaw, bb_area;
  double initMeas[2];
  double distFromInit;
  double x_merge_yaw;
  double velo[4]; int nVelo;
};

// file-scope globals of imm_ukf_jpda.cpp:19-24,56-58,66-70
struct Tracker {
  std::vector<Track> targets;
  std::vector<int> trackNum;
  bool init = false;
  double timestamp = 0, egoVelo = 0, egoYaw = 0, egoPreYaw = 0;
  double egoPoint[3] = {0, 0, 0};   // egoPoints_[0]
  double fold[3] = {0, 0, -1.5707963267948966};  // running (x, y, egoYaw) of getOriginPoints' replay
};

struct StepOut {
  std::vector<float> targets; std::vector<double> vandyaw; std::vector<int> track_manage;
  std::vector<uint8_t> is_static, is_vis; std::vector<float> vis_bb;
  void clear() { targets.clear(); vandyaw.clear(); track_manage.clear(); is_static.clear(); is_vis.clear(); vis_bb.clear(); }
};

struct GridDump { float *minz, *height, *smoothed, *hdiff, *hground; uint8_t* isground; };
void cell_index(float x, float y, int& chI, int& binI);
void ground_remove(const float* xyz, int n, int stride, std::vector<float>& elev, std::vector<float>& ground, GridDump* dump);
void component_clustering(const float* xyz, int n, int stride, int* grid, int& numCluster);
void box_fitting(const float* xyz, int n, int stride, const int* grid, int numCluster, int mode, std::vector<float>& boxes,
                 std::vector<float>& markers);
void ukf_initialize(Track& t, double zx, double zy);
void get_origin_points(Tracker& T, double timestamp, double v_gps, double yaw_gps);
void imm_ukf_jpdaf(Tracker& T, const float* boxes, int nb, double timestamp, StepOut& out);

}  // namespace port

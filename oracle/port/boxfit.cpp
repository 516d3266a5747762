// TEST INFRASTRUCTURE ONLY -- plain-C++ restatement of boxFitting
// (/root/reference/object_tracking/src/cluster/box_fitting.cpp).  cv::minAreaRect is the MAR contract of
// oracle/mar_contract.cpp (OpenCV is not in /root/reference: parity for that call is unpinned).
// rule_mode 0 = INTENDED (fall-through of ruleBasedFilter means false), 1 = GCC13_O2_COMPAT.
#include <cmath>
#include <random>
#include <vector>
#include "port.h"

extern "C" int lmot_oracle_mar(const int* xy, int n, float* corners8);

namespace port {

static const int numGrid = 250;
static float roiM = 50;
static float picScale = 900 / roiM;   // box_fitting.cpp:18
static int ramPoints = 80, lSlopeDist = 1, lnumPoints = 5;   // :19-23
static float sensorHeight = 2, tHeightMin = 0.8f, tHeightMax = 2.6f, tWidthMin = 0.2f, tWidthMax = 3.5f, tLenMin = 0.2f, tLenMax = 14.0f,
             tAreaMax = 20.0f, tRatioMin = 1, tRatioMax = 8.0f, minLenRatio = 3.0f, tPtPerM3 = 8;   // :26-44

bool cart_index(float x, float y, int& xI, int& yI);

static bool rule_based_filter(const float pc[4][2], float maxZ, int numPoints, int mode) {   // :97-158
  if (numPoints < 30) return false;
  const float x1 = pc[0][0], y1 = pc[0][1], x2 = pc[1][0], y2 = pc[1][1], x3 = pc[2][0], y3 = pc[2][1];
  const float dist1 = sqrtf((x1 - x2) * (x1 - x2) + (y1 - y2) * (y1 - y2));
  const float dist2 = sqrtf((x3 - x2) * (x3 - x2) + (y3 - y2) * (y3 - y2));
  float length, width;
  if (dist1 > dist2) { length = dist1; width = dist2; } else { length = dist2; width = dist1; }
  const float height = maxZ + sensorHeight;
  const float area = dist1 * dist2;
  const float mass = area * height;
  const float ratio = length / width;
  if (height > tHeightMin && height < tHeightMax) {
    if (mode == 1) return true;
    if (width > tWidthMin && width < tWidthMax)
      if (length > tLenMin && length < tLenMax)
        if (area < tAreaMax)
          if (numPoints > mass * tPtPerM3) {
            if (length > minLenRatio) { if (ratio > tRatioMin && ratio < tRatioMax) return true; }
            else return true;
          }
    return false;
  }
  return false;
}

void box_fitting(const float* xyz, int n, int stride, const int* grid, int numCluster, int mode, std::vector<float>& boxes,
                 std::vector<float>& markers) {
  std::vector<std::vector<float>> cl(numCluster);   // getClusteredPoints :46-72
  for (int i = 0; i < n; ++i) {
    const float x = xyz[(size_t)i * stride], y = xyz[(size_t)i * stride + 1], z = xyz[(size_t)i * stride + 2];
    int xI, yI;
    if (!cart_index(x, y, xI, yI)) continue;
    const int id = grid[xI * numGrid + yI];
    if (id != 0) { cl[id - 1].push_back(x); cl[id - 1].push_back(y); cl[id - 1].push_back(z); }
  }
  for (int ic = 0; ic < numCluster; ++ic) {          // getBoundingBox :212-418
    const std::vector<float>& P = cl[ic];
    const int numPoints = (int)P.size() / 3;
    if (numPoints == 0) continue;
    const float initPX = P[0] + roiM / 2, initPY = P[1] + roiM / 2;
    const int initX = (int)floorf(initPX * picScale), initY = (int)floorf(initPY * picScale);
    const int initPicX = initX;
    const int initPicY = (int)(picScale * roiM - initY);
    const int offsetInitX = (int)(roiM * picScale / 2 - initPicX);
    const int offsetInitY = (int)(roiM * picScale / 2 - initPicY);
    std::vector<int> pix((size_t)numPoints * 2);
    float minMx = 0, minMy = 0, maxMx = 0, maxMy = 0;   // uninitialised in the reference when no slope beats the seeds
    float minM = 999, maxM = -999, maxZ = -99;
    for (int i = 0; i < numPoints; ++i) {
      const float pX = P[3 * i], pY = P[3 * i + 1], pZ = P[3 * i + 2];
      const float roiX = pX + roiM / 2, roiY = pY + roiM / 2;
      const int x = (int)floorf(roiX * picScale), y = (int)floorf(roiY * picScale);
      const int picX = x, picY = (int)(picScale * roiM - y);
      pix[2 * i] = picX + offsetInitX; pix[2 * i + 1] = picY + offsetInitY;
      const float m = pY / pX;
      if (m < minM) { minM = m; minMx = pX; minMy = pY; }
      if (m > maxM) { maxM = m; maxMx = pX; maxMy = pY; }
      if (pZ > maxZ) maxZ = pZ;
    }
    const float xDist = maxMx - minMx, yDist = maxMy - minMy;
    const float slopeDist = sqrtf(xDist * xDist + yDist * yDist);
    const float slope = (maxMy - minMy) / (maxMx - minMx);
    std::mt19937_64 mt(0);
    std::uniform_int_distribution<> randPoints(0, numPoints - 1);
    float pc[4][2];
    if (slopeDist > lSlopeDist && numPoints > lnumPoints && (maxMy > 8 || maxMy < -5)) {   // :308-356
      float maxDist = 0, maxDx = 0, maxDy = 0;   // maxDx/maxDy uninitialised in the reference if no dist > 0
      for (int i = 0; i < ramPoints; ++i) {
        const int pInd = randPoints(mt);
        const float xI = P[3 * pInd], yI = P[3 * pInd + 1];
        const float dist = fabsf(slope * xI - 1 * yI + maxMy - slope * maxMx) / sqrtf(slope * slope + 1);
        if (dist > maxDist) { maxDist = dist; maxDx = xI; maxDy = yI; }
      }
      const float maxMvecX = maxMx - maxDx, maxMvecY = maxMy - maxDy, minMvecX = minMx - maxDx, minMvecY = minMy - maxDy;
      pc[0][0] = minMx; pc[0][1] = minMy; pc[1][0] = maxDx; pc[1][1] = maxDy; pc[2][0] = maxMx; pc[2][1] = maxMy;
      pc[3][0] = maxDx + maxMvecX + minMvecX; pc[3][1] = maxDy + maxMvecY + minMvecY;
    } else {                                                                                  // :358-365
      float rc[8];
      lmot_oracle_mar(pix.data(), numPoints, rc);
      for (int c = 0; c < 4; ++c) {   // getPointsInPcFrame :75-95
        const float rOffsetX = rc[2 * c] - offsetInitX, rOffsetY = rc[2 * c + 1] - offsetInitY;
        const float rX = rOffsetX, rY = picScale * roiM - rOffsetY;
        const float rmX = rX / picScale, rmY = rY / picScale;
        pc[c][0] = rmX - roiM / 2; pc[c][1] = rmY - roiM / 2;
      }
    }
    if (!rule_based_filter(pc, maxZ, numPoints, mode)) continue;
    for (int h = 0; h < 2; ++h) for (int c = 0; c < 4; ++c) { boxes.push_back(pc[c][0]); boxes.push_back(pc[c][1]); boxes.push_back(h == 0 ? -sensorHeight : maxZ); }
    // mark_cluster :161-209 (pcl::compute3DCentroid float accumulation in cloud order, getMinMax3D)
    float cs[3] = {0, 0, 0}, mn[3] = {P[0], P[1], P[2]}, mx[3] = {P[0], P[1], P[2]};
    for (int i = 0; i < numPoints; ++i) for (int a = 0; a < 3; ++a) { cs[a] += P[3 * i + a]; mn[a] = std::fmin(mn[a], P[3 * i + a]); mx[a] = std::fmax(mx[a], P[3 * i + a]); }
    for (int a = 0; a < 3; ++a) markers.push_back(cs[a] / (float)numPoints);
    for (int a = 0; a < 3; ++a) { float s = mx[a] - mn[a]; if (s == 0) s = 0.1f; markers.push_back(s); }
  }
}

}  // namespace port

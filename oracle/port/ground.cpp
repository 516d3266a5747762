// TEST INFRASTRUCTURE ONLY -- plain-C++ restatement of the reference ground removal
// (/root/reference/object_tracking/src/groundremove/ground_removal.cpp, gaus_blur.cpp), sequential like the
// reference.  float / double types follow the reference expression by expression; atan2f is the host libm's
// (which is what the reference links).  Validated bit-for-bit against oracle/_ref (tests/test_oracle_port.py).
#include <cmath>
#include <cstdint>
#include <vector>
#include "port.h"

namespace port {

static const int numChannel = 80, numBin = 120;                    // ground_removal.h:16-17
static float rMin = 3.4f, rMax = 120.f, tHmin = -2.0f, tHmax = -0.4f, tHDiff = 0.4f, hSeonsor = 2.f;   // ground_removal.cpp:24-33

struct Cell { float smoothed, height, hDiff, hGround, minZ; bool isGround; };   // ground_removal.h:30-52

// getCellIndexFromPoints, ground_removal.cpp:67-76
void cell_index(float x, float y, int& chI, int& binI) {
  const float distance = sqrtf(x * x + y * y);
  const float chP = (float)((atan2f(y, x) + M_PI) / (2 * M_PI));
  const float binP = (distance - rMin) / (rMax - rMin);
  chI = (int)floorf(chP * numChannel);
  binI = (int)floorf(binP * numBin);
}

// groundRemove ground_removal.cpp:177-249; grid (optional) receives the final cells
void ground_remove(const float* xyz, int n, int stride, std::vector<float>& elev, std::vector<float>& ground, GridDump* dump) {
  // filterCloud :46-64
  std::vector<float> f;
  f.reserve((size_t)n * 3);
  for (int i = 0; i < n; ++i) {
    const float x = xyz[(size_t)i * stride], y = xyz[(size_t)i * stride + 1], z = xyz[(size_t)i * stride + 2];
    const float distance = sqrtf(x * x + y * y);
    if (distance <= rMin || distance >= rMax) continue;
    f.push_back(x); f.push_back(y); f.push_back(z);
  }
  const int nf = (int)f.size() / 3;
  static Cell polar[numChannel][numBin];
  for (int c = 0; c < numChannel; ++c) for (int b = 0; b < numBin; ++b) { polar[c][b].minZ = 1000; polar[c][b].isGround = false; }   // Cell::Cell :35-38
  // createAndMapPolarGrid :79-92
  for (int i = 0; i < nf; ++i) {
    int chI, binI;
    cell_index(f[3 * i], f[3 * i + 1], chI, binI);
    if (chI < 0 || chI >= numChannel || binI < 0 || binI >= numBin) continue;
    if (f[3 * i + 2] < polar[chI][binI].minZ) polar[chI][binI].minZ = f[3 * i + 2];
  }
  // gaussKernel(3, 1) gaus_blur.cpp:26-49
  double kernel[3];
  {
    const int samples = 3; const double sigma = 1; const double mean = samples / 2; double sum = 0.0;
    for (int x = 0; x < samples; ++x) { kernel[x] = exp(-0.5 * (pow((x - mean) / sigma, 2.0))) / (2 * M_PI * sigma * sigma); sum += kernel[x]; }
    for (int x = 0; x < samples; ++x) kernel[x] /= sum;
  }
  for (int c = 0; c < numChannel; ++c) {
    for (int b = 0; b < numBin; ++b) {                      // :192-197
      const float zi = polar[c][b].minZ;
      if (zi > tHmin && zi < tHmax) polar[c][b].height = zi;
      else if (zi > tHmax) polar[c][b].height = hSeonsor;
      else polar[c][b].height = tHmin;
    }
    for (long i = 0; i < numBin; ++i) {                    // gaussSmoothen gaus_blur.cpp:52-68
      double smoothed = 0;
      for (long j = i - 1; j <= i + 1; ++j)
        if (j >= 0 && j < numBin) smoothed += kernel[1 + (j - i)] * polar[c][j].height;
      polar[c][i].smoothed = (float)smoothed;
    }
    for (int i = 0; i < numBin; ++i) {                     // computeHDiffAdjacentCell :95-117
      if (i == 0) polar[c][i].hDiff = polar[c][i].height - polar[c][i + 1].height;
      else if (i == numBin - 1) polar[c][i].hDiff = polar[c][i].height - polar[c][i - 1].height;
      else {
        const float pre = polar[c][i].height - polar[c][i - 1].height, post = polar[c][i].height - polar[c][i + 1].height;
        polar[c][i].hDiff = (pre > post) ? pre : post;
      }
    }
    for (int b = 0; b < numBin; ++b) {                     // :205-214
      Cell& k = polar[c][b];
      if (k.smoothed < tHmax && k.hDiff < tHDiff) { k.isGround = true; k.hGround = k.height; }
      else if (k.height < tHmax && k.hDiff < tHDiff) { k.isGround = true; k.hGround = k.height; }
    }
  }
  // applyMedianFilter :120-146 (in place, raster order)
  for (int c = 1; c < numChannel - 1; ++c)
    for (int b = 1; b < numBin - 1; ++b)
      if (!polar[c][b].isGround && polar[c][b + 1].isGround && polar[c][b - 1].isGround && polar[c + 1][b].isGround && polar[c - 1][b].isGround) {
        float s[4] = {polar[c][b + 1].height, polar[c][b - 1].height, polar[c + 1][b].height, polar[c - 1][b].height};
        for (int i = 1; i < 4; ++i) { float v = s[i]; int j = i - 1; while (j >= 0 && s[j] > v) { s[j + 1] = s[j]; --j; } s[j + 1] = v; }
        const float median = (s[1] + s[2]) / 2;
        polar[c][b].height = median; polar[c][b].isGround = true; polar[c][b].hGround = median;
      }
  // outlierFilter :149-174 (in place, order dependent along bin)
  for (int c = 1; c < numChannel - 1; ++c)
    for (int b = 1; b < numBin - 2; ++b)
      if (polar[c][b].isGround && polar[c][b + 1].isGround && polar[c][b - 1].isGround && polar[c][b + 2].isGround) {
        const float h1 = polar[c][b - 1].height, h2 = polar[c][b].height, h3 = polar[c][b + 1].height, h4 = polar[c][b + 2].height;
        if (h1 != tHmin && h2 == tHmin && h3 != tHmin) { polar[c][b].height = (h1 + h3) / 2; polar[c][b].hGround = polar[c][b].height; }
        else if (h1 != tHmin && h2 == tHmin && h3 == tHmin && h4 != tHmin) { polar[c][b].height = (h1 + h4) / 2; polar[c][b].hGround = polar[c][b].height; }
      }
  // classification :221-247
  for (int i = 0; i < nf; ++i) {
    const float x = f[3 * i], y = f[3 * i + 1], z = f[3 * i + 2];
    int chI, binI;
    cell_index(x, y, chI, binI);
    if (chI < 0 || chI >= numChannel || binI < 0 || binI >= numBin) continue;
    std::vector<float>* dst = &elev;
    if (polar[chI][binI].isGround) { if (z < (polar[chI][binI].hGround + 0.25)) dst = &ground; }
    dst->push_back(x); dst->push_back(y); dst->push_back(z);
  }
  if (dump)
    for (int c = 0; c < numChannel; ++c) for (int b = 0; b < numBin; ++b) {
      const int k = c * numBin + b; const Cell& q = polar[c][b];
      dump->minz[k] = q.minZ; dump->height[k] = q.height; dump->smoothed[k] = q.smoothed; dump->hdiff[k] = q.hDiff;
      dump->isground[k] = q.isGround; dump->hground[k] = q.isGround ? q.hGround : 0.f;
    }
}

}  // namespace port

// TEST INFRASTRUCTURE ONLY -- plain-C++ restatement of componentClustering
// (/root/reference/object_tracking/src/cluster/component_clustering.cpp:28-268).  The recursive flood fill is
// restated with an explicit stack that visits neighbours in the same order (depth-first, kX outer / kY inner),
// so labels are identical while a 62,500-cell component cannot overflow the call stack.
#include <cmath>
#include <vector>
#include "port.h"

namespace port {

static const int numGrid = 250;            // component_clustering.h:13
static float roiM = 50;                    // component_clustering.cpp:11

bool cart_index(float x, float y, int& xI, int& yI) {   // :40-48
  const float xC = x + roiM / 2, yC = y + roiM / 2;
  if (xC < 0 || xC >= roiM || yC < 0 || yC >= roiM) return false;
  xI = (int)floorf(numGrid * xC / roiM);
  yI = (int)floorf(numGrid * yC / roiM);
  return true;
}

void component_clustering(const float* xyz, int n, int stride, int* grid, int& numCluster) {
  std::vector<int> gridNum((size_t)numGrid * numGrid, 0);
  for (int i = 0; i < numGrid * numGrid; ++i) grid[i] = 0;
  for (int i = 0; i < n; ++i) {            // mapCartesianGrid :38-47
    int xI, yI;
    if (!cart_index(xyz[(size_t)i * stride], xyz[(size_t)i * stride + 1], xI, yI)) continue;
    gridNum[xI * numGrid + yI] += 1;
  }
  for (int xI = 0; xI < numGrid; ++xI)     // :134-214: cells with more than one point + clipped 3x3 neighbourhood
    for (int yI = 0; yI < numGrid; ++yI)
      if (gridNum[xI * numGrid + yI] > 1)
        for (int dx = -1; dx <= 1; ++dx) for (int dy = -1; dy <= 1; ++dy) {
          const int xx = xI + dx, yy = yI + dy;
          if (xx >= 0 && xx < numGrid && yy >= 0 && yy < numGrid) grid[xx * numGrid + yy] = -1;
        }
  numCluster = 0;                           // findComponent :247-257 + search :228-244
  struct Frame { int x, y, k; };
  std::vector<Frame> st;
  for (int cx = 0; cx < numGrid; ++cx)
    for (int cy = 0; cy < numGrid; ++cy) {
      if (grid[cx * numGrid + cy] != -1) continue;
      const int id = ++numCluster;
      grid[cx * numGrid + cy] = id;
      st.push_back({cx, cy, 0});
      while (!st.empty()) {
        Frame& fr = st.back();
        if (fr.k == 9) { st.pop_back(); continue; }
        const int kX = fr.k / 3 - 1, kY = fr.k % 3 - 1;
        ++fr.k;
        const int nx = fr.x + kX, ny = fr.y + kY;
        if (nx < 0 || nx >= numGrid || ny < 0 || ny >= numGrid) continue;
        if (grid[nx * numGrid + ny] == -1) { grid[nx * numGrid + ny] = id; st.push_back({nx, ny, 0}); }
      }
    }
}

}  // namespace port

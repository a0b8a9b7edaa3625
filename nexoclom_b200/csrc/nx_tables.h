// Host-side preparation of lookup tables (pure C++, no CUDA).
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>

namespace nx {

struct HostInterp {
  std::vector<double> x, f, slope;
  std::vector<unsigned short> bucket;
  double blo = 0.0, binvw = 0.0;
  int nbucket = 0;
};

// slopes as np.interp computes them, and a uniform bucket index:
// bucket[b] = largest j with x[j] <= blo + b*w  (0 if none).
inline HostInterp make_interp(const double* x, const double* f, int n, int nbucket = 4096) {
  HostInterp t;
  t.x.assign(x, x + n);
  t.f.assign(f, f + n);
  t.slope.resize(n > 1 ? n : 1, 0.0);
  for (int j = 0; j + 1 < n; ++j) t.slope[j] = (f[j + 1] - f[j]) / (x[j + 1] - x[j]);
  t.nbucket = nbucket;
  t.bucket.resize(nbucket);
  const double lo = x[0], hi = x[n - 1];
  const double w = (hi > lo) ? (hi - lo) / nbucket : 1.0;
  t.blo = lo;
  t.binvw = 1.0 / w;
  for (int b = 0; b < nbucket; ++b) {
    const double edge = lo + b * w;
    int j = (int)(std::upper_bound(x, x + n, edge) - x) - 1;
    t.bucket[b] = (unsigned short)std::max(j, 0);
  }
  return t;
}

}  // namespace nx

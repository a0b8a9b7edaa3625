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

// Record form of the same table for the fast path: one 32-byte record per
// interval (plus a clamp record at either end) so a lookup is one bucket load and
// one record load.  rec[0] = left clamp, rec[j+1] = [x[j], x[j+1]), rec[n] = right clamp.
struct HostFastTable {
  std::vector<double> rec;               // 4 doubles per record: lo, hi, f, slope
  std::vector<unsigned short> bucket;    // record containing the bucket's lower edge
  double blo = 0.0, binvw = 0.0;
  int nrec = 0, nbucket = 0;
};

inline HostFastTable make_fast_table(const double* x, const double* f, int n, int nbucket = 4096) {
  HostFastTable t;
  t.nrec = n + 1;
  t.rec.resize((size_t)4 * t.nrec);
  auto put = [&](int r, double lo, double hi, double fv, double sl) {
    t.rec[4 * r] = lo; t.rec[4 * r + 1] = hi; t.rec[4 * r + 2] = fv; t.rec[4 * r + 3] = sl;
  };
  put(0, -1e300, x[0], f[0], 0.0);
  for (int j = 0; j + 1 < n; ++j)
    put(j + 1, x[j], x[j + 1], f[j], (f[j + 1] - f[j]) / (x[j + 1] - x[j]));
  put(n, x[n - 1], 1.7976931348623157e308, f[n - 1], 0.0);
  t.nbucket = nbucket;
  t.bucket.resize(nbucket);
  const double lo = x[0], hi = x[n - 1];
  const double w = (hi > lo) ? (hi - lo) / nbucket : 1.0;
  t.blo = lo;
  t.binvw = 1.0 / w;
  for (int b = 0; b < nbucket; ++b) {
    const double edge = lo + b * w;
    int cnt = (int)(std::upper_bound(x, x + n, edge) - x);     // nodes <= edge
    t.bucket[b] = (unsigned short)cnt;                          // rec[cnt] = [x[cnt-1], x[cnt])
  }
  return t;
}

}  // namespace nx

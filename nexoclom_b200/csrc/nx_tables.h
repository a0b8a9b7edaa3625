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
// one record load, plus AT MOST TWO steps to the following records -- no search loop:
//   rec[0] = left clamp, rec[j+1] = [x[j], x[j+1]), rec[n] = right clamp, rec[n+1..n+2] = pad.
// bucket[b] is the record containing the point  edge_b - margin  (margin = 1e-6 bucket
// widths, far above the rounding of the device's bucket computation).  The builder
// doubles the bucket count until no bucket (widened by the margin) holds more than two
// nodes (union grids of two emission lines have near-coincident node PAIRS), so the
// record of any v that maps to bucket b is bucket[b], +1 or +2.
#define NX_FAST_TABLE_MAX_STEPS 2
struct HostFastTable {
  std::vector<double> rec;               // 4 doubles per record: lo, hi, f, slope
  std::vector<unsigned short> bucket;
  double blo = 0.0, binvw = 0.0;
  int nrec = 0, nbucket = 0;
  int max_steps = 0;                     // largest number of nodes any widened bucket holds
};

inline HostFastTable make_fast_table(const double* x, const double* f, int n, int nbucket = 4096) {
  HostFastTable t;
  t.nrec = n + 3;
  t.rec.resize((size_t)4 * t.nrec);
  auto put = [&](int r, double lo, double hi, double fv, double sl) {
    t.rec[4 * r] = lo; t.rec[4 * r + 1] = hi; t.rec[4 * r + 2] = fv; t.rec[4 * r + 3] = sl;
  };
  put(0, -1e300, x[0], f[0], 0.0);
  for (int j = 0; j + 1 < n; ++j)
    put(j + 1, x[j], x[j + 1], f[j], (f[j + 1] - f[j]) / (x[j + 1] - x[j]));
  for (int r = n; r < n + 3; ++r) put(r, x[n - 1], 1.7976931348623157e308, f[n - 1], 0.0);
  const double lo = x[0], hi = x[n - 1];
  for (;;) {
    const double w = (hi > lo) ? (hi - lo) / nbucket : 1.0;
    t.nbucket = nbucket;
    t.bucket.resize(nbucket);
    t.blo = lo;
    t.binvw = 1.0 / w;
    t.max_steps = 0;
    for (int b = 0; b < nbucket; ++b) {
      const double p0 = lo + (b - 1e-6) * w, p1 = lo + (b + 1 + 1e-6) * w;
      const int c0 = (int)(std::upper_bound(x, x + n, p0) - x);   // nodes <= p0
      const int c1 = (int)(std::upper_bound(x, x + n, p1) - x);
      t.bucket[b] = (unsigned short)c0;                           // rec[c0] = [x[c0-1], x[c0])
      t.max_steps = std::max(t.max_steps, c1 - c0);
    }
    if (t.max_steps <= NX_FAST_TABLE_MAX_STEPS || nbucket >= (1 << 22)) break;
    nbucket *= 2;
  }
  return t;
}

}  // namespace nx

// nexoclom_b200 -- K6: source maps (reference data_simulation/make_source_map.py:11-175).
//
// The reference histograms the initial states X0 over the surface and, for each of the
// nlon x nlat grid points, gathers every packet within a haversine radius
// smear_radius * cos(lat_point) (sklearn BallTree.query_radius) to build local speed /
// altitude / azimuth distributions -- a Python loop over 16 200 points with pandas
// selections.  Here the loop is turned inside out: one thread per packet walks the few
// latitude rows and longitude columns whose points can reach it, applies sklearn's own
// reduced-distance test
//     sin^2((lat_p - lat)/2) + cos(lat_p) cos(lat) sin^2((lon_p - lon)/2) <= sin^2(r_p / 2)
// in that operation order, and scatters into the per-point histograms with f64 / u64
// atomics (the 16 200 x 171 output block is L2 resident).  Bin indices follow
// np.histogram exactly (edges = linspace, right edge inclusive; nx_image.cuh::hist_bin).
#include <cuda_runtime.h>
#include <stdint.h>

#include "nx_image.cuh"
#include "nx_kernels.h"

namespace nx {

__global__ void __launch_bounds__(256)
k_source_map(long long n, SourceMapParams sp, const double* __restrict__ lon_,
             const double* __restrict__ lat_, const double* __restrict__ v_,
             const double* __restrict__ alt_, const double* __restrict__ az_,
             const double* __restrict__ frac_, const double* __restrict__ plon,
             const double* __restrict__ plat, const double* __restrict__ pcos,
             const double* __restrict__ pthr, SourceMapOut out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double lon = lon_[i], lat = lat_[i], v = v_[i], alt = alt_[i], az = az_[i];
  const double frac = frac_[i];
  const bool incl = frac > 0.0;                                   // make_source_map.py:58
  const double w = sp.weight_is_frac ? frac : 1.0;                // :59-65

  const double two_pi = 2.0 * NX_PI, half_pi = NX_PI / 2.0;
  const double step_v = sp.vmax / sp.nvel, step_alt = half_pi / sp.nalt, step_az = two_pi / sp.naz;
  const int bv = hist_bin(v, sp.nvel, 0.0, sp.vmax, step_v);
  const int balt = hist_bin(alt, sp.nalt, 0.0, half_pi, step_alt);
  const int baz = hist_bin(az, sp.naz, 0.0, two_pi, step_az);

  if (incl) {
    // whole-planet histograms (:72-109)
    const double step_lon = two_pi / sp.nlon, step_lat = NX_PI / sp.nlat;
    const int blon = hist_bin(lon, sp.nlon, 0.0, two_pi, step_lon);
    const int blat = hist_bin(lat, sp.nlat, -half_pi, half_pi, step_lat);
    if (blon >= 0 && blat >= 0) atomicAdd(&out.abundance_hist[blon * sp.nlat + blat], w);
    if (bv >= 0) atomicAdd(&out.speed_dist[bv], w);
    if (balt >= 0) atomicAdd(&out.altitude_dist[balt], w);
    if (baz >= 0) atomicAdd(&out.azimuth_dist[baz], w);
  }

  // grid points that reach this packet (:111-160)
  const double dlat = NX_PI / sp.nlat, dlon = two_pi / sp.nlon;
  const double coslat = cos(lat);
  int j0 = (int)floor((lat - sp.smear + half_pi) / dlat - 0.5) - 1;
  int j1 = (int)ceil((lat + sp.smear + half_pi) / dlat - 0.5) + 1;
  j0 = j0 < 0 ? 0 : j0;
  j1 = j1 > sp.nlat - 1 ? sp.nlat - 1 : j1;
  for (int j = j0; j <= j1; ++j) {
    const double s0 = sin(0.5 * sub_rn(plat[j], lat));
    const double a = mul_rn(s0, s0);
    const double thr = pthr[j];
    if (a > thr) continue;
    const double cc = mul_rn(pcos[j], coslat);
    // columns: c sin^2(dlon/2) <= thr - a
    int ncol = sp.nlon, ic = 0;
    if (cc > 0.0) {
      const double q = (thr - a) / cc * (1.0 + 1e-9) + 1e-15;
      if (q < 1.0) {
        const double dmax = 2.0 * asin(sqrt(q));
        ic = (int)floor((lon - dmax) / dlon - 0.5) - 1;
        const int ie = (int)ceil((lon + dmax) / dlon - 0.5) + 1;
        ncol = ie - ic + 1;
        if (ncol > sp.nlon) { ncol = sp.nlon; ic = 0; }
      }
    }
    for (int t = 0; t < ncol; ++t) {
      int ii = (ic + t) % sp.nlon;
      if (ii < 0) ii += sp.nlon;
      const double s1 = sin(0.5 * sub_rn(plon[ii], lon));
      const double rd = add_rn(a, mul_rn(mul_rn(cc, s1), s1));
      if (!(rd <= thr)) continue;
      const size_t p = (size_t)ii * sp.nlat + j;
      atomicAdd(&out.n_total[p], 1ull);
      atomicAdd(&out.abundance[p], w);                              // sub_weight.sum() (:137)
      if (incl) {
        atomicAdd(&out.n_included[p], 1ull);
        if (bv >= 0) atomicAdd(&out.speed_map[p * sp.nvel + bv], w);
        if (balt >= 0) atomicAdd(&out.altitude_map[p * sp.nalt + balt], w);
        if (baz >= 0) atomicAdd(&out.azimuth_map[p * sp.naz + baz], w);
      }
    }
  }
}

cudaError_t launch_source_map(cudaStream_t st, long long n, const SourceMapParams& sp,
                              const double* lon, const double* lat, const double* v,
                              const double* alt, const double* az, const double* frac,
                              const double* plon, const double* plat, const double* pcos,
                              const double* pthr, const SourceMapOut& out) {
  if (n <= 0) return cudaSuccess;
  k_source_map<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, sp, lon, lat, v, alt, az, frac,
                                                            plon, plat, pcos, pthr, out);
  return cudaGetLastError();
}

}  // namespace nx

// TEST-ONLY: compiles the kernels' shared per-packet physics (nx_physics.cuh)
// with g++ so that logic can be debugged against the oracle on machines without
// a GPU.  Never loaded by nexoclom_b200 (the product fails loudly without CUDA).
#include "../../nexoclom_b200/csrc/nx_physics.cuh"
#include "../../nexoclom_b200/csrc/nx_fast.cuh"
#include "../../nexoclom_b200/csrc/nx_tables.h"
#include "../../nexoclom_b200/csrc/nx_surface.cuh"

using namespace nx;

static InterpTable view(const HostInterp& h) {
  InterpTable t;
  t.x = h.x.data(); t.f = h.f.data(); t.slope = h.slope.data(); t.bucket = h.bucket.data();
  t.n = (int)h.x.size(); t.nbucket = h.nbucket; t.blo = h.blo; t.binvw = h.binvw;
  return t;
}

extern "C" int hc_dp_step(long n, const double* X, const double* h, double* out, double* delta,
                          const RunParams* p, const double* rpv, const double* rpa, int nrp,
                          int strict) {
  HostInterp hi; InterpTable T{};
  if (nrp > 0) { hi = make_interp(rpv, rpa, nrp); T = view(hi); }
  for (long i = 0; i < n; ++i) {
    double d[7];
    if (strict) dp_step<true, true>(*p, T, X + 8 * i, h[i], out + 8 * i, d);
    else dp_step<false, true>(*p, T, X + 8 * i, h[i], out + 8 * i, d);
    delta[8 * i] = 0.0;
    for (int k = 0; k < 7; ++k) delta[8 * i + 1 + k] = d[k];
  }
  return 0;
}

extern "C" int hc_integrate_adaptive(long n, double* X /* n x 8 row-major */, double* step,
                                     unsigned* att, unsigned* acc, const RunParams* p,
                                     const double* rpv, const double* rpa, int nrp, int strict) {
  HostInterp hi; InterpTable T{};
  HostFastTable hf; FastTable F{};
  if (nrp > 0) {
    hi = make_interp(rpv, rpa, nrp); T = view(hi);
    hf = make_fast_table(rpv, rpa, nrp, 32768);
    F.rec = reinterpret_cast<const InterpRec*>(hf.rec.data()); F.bucket = hf.bucket.data();
    F.nrec = hf.nrec; F.nbucket = hf.nbucket; F.blo = hf.blo; F.binvw = hf.binvw; F.boff = -hf.blo * hf.binvw;
  }
  int status = 0;
  for (long i = 0; i < n; ++i) {
    double* s = X + 8 * i;
    att[i] = acc[i] = 0;
    bool live = (s[0] > p->resolution) && (s[7] > 0.0);
    while (live) {
      // strict = 1: NumPy operation order; 0: the product's fast path; 2: FMA-contracted
      // variant of the strict template (kept for A/B comparisons)
      int fl = strict == 1 ? adaptive_attempt<true>(*p, T, s, step[i])
               : strict == 2 ? adaptive_attempt<false>(*p, T, s, step[i])
                             : adaptive_attempt_fast_rt(*p, F, s, step[i]);
      att[i]++;
      if (fl & ATT_ACCEPTED) acc[i]++;
      status |= fl & ~(ATT_ACCEPTED | ATT_LIVE);
      live = fl & ATT_LIVE;
    }
  }
  return status;
}

extern "C" int hc_integrate_constant(long n, double* X /* n x 8 */, double* traj /* n x 8 x nsteps or null */,
                                     int nsteps, unsigned long long seed, unsigned long long first_id,
                                     const RunParams* p, const double* rpv, const double* rpa, int nrp,
                                     const double* tx, int ntx, const double* ty, int nty,
                                     const double* c, int strict) {
  HostInterp hi; InterpTable T{};
  HostFastTable hf; FastTable F{};
  if (nrp > 0) {
    hi = make_interp(rpv, rpa, nrp); T = view(hi);
    hf = make_fast_table(rpv, rpa, nrp, 32768);
    F.rec = reinterpret_cast<const InterpRec*>(hf.rec.data()); F.bucket = hf.bucket.data();
    F.nrec = hf.nrec; F.nbucket = hf.nbucket; F.blo = hf.blo; F.binvw = hf.binvw; F.boff = -hf.blo * hf.binvw;
  }
  Spline2D S{tx, ty, c, ntx, nty};
  for (long i = 0; i < n; ++i) {
    double* s = X + 8 * i;
    if (traj) for (int k = 0; k < 8; ++k) traj[(i * 8 + k) * (long)nsteps + 0] = s[k];
    bool live = s[7] > 0.0;
    double curtime = p->endtime;
    int ct = 1;
    while (curtime > 0 && ct < nsteps) {
      if (live) {
        live = strict ? constant_step<true>(*p, T, S, s, seed, first_id + i, (uint32_t)ct)
                      : constant_step_fast_rt(*p, F, S, s, seed, first_id + i, (uint32_t)ct);
        if (traj) for (int k = 0; k < 8; ++k) traj[(i * 8 + k) * (long)nsteps + ct] = s[k];
      }
      ++ct; curtime -= p->step_size;
    }
  }
  return 0;
}

extern "C" void hc_uniform_pairs(long n, unsigned long long seed, unsigned long long first_id,
                                 unsigned stream, unsigned draw, double* u0, double* u1) {
  for (long i = 0; i < n; ++i) uniform_pair(seed, first_id + i, stream, draw, u0[i], u1[i]);
}

extern "C" void hc_spline_ev(long n, const double* x, const double* y, double* out,
                             const double* tx, int ntx, const double* ty, int nty, const double* c) {
  Spline2D S{tx, ty, c, ntx, nty};
  for (long i = 0; i < n; ++i) out[i] = spline2d_ev(S, x[i], y[i]);
}

#include "../../nexoclom_b200/csrc/nx_init.cuh"
#include "../../nexoclom_b200/csrc/nx_image.cuh"

extern "C" void hc_init_state(long n, const SourceParams* sp, unsigned long long seed,
                              unsigned long long first_id, const double* fmap,
                              const double* axes /* x_lo,x_hi,y_lo,y_hi */,
                              const double* cdf, const double* vtab, int ntab,
                              double* out /* n x 14 */) {
  HostInterp hs; InterpTable speed{};
  if (ntab > 0) { hs = make_interp(cdf, vtab, ntab); speed = view(hs); }
  SourceMap map{};
  if (fmap) { map.f = fmap; map.x_lo = axes[0]; map.x_hi = axes[1]; map.y_lo = axes[2]; map.y_hi = axes[3]; }
  for (long i = 0; i < n; ++i) init_packet(*sp, map, speed, seed, first_id + i, out + 14 * i);
}

// the pure transform of K1 on caller-supplied deviates (lon_in/lat_in: surface points
// sampled elsewhere, or null for the uniform band computed from u_sinlat/u_lon)
extern "C" void hc_init_from_deviates(long n, const SourceParams* sp, const double* cdf,
                                      const double* vtab, int ntab, const double* u_time,
                                      const double* u_sinlat, const double* u_lon,
                                      const double* lon_in, const double* lat_in,
                                      const double* u_speed, const double* z_normal,
                                      const double* u_alt, const double* u_az,
                                      double* out /* n x 14 */, const double* lon_cdf,
                                      const double* lon_tab, int nlon) {
  HostInterp hs; InterpTable speed{};
  if (ntab > 0) { hs = make_interp(cdf, vtab, ntab); speed = view(hs); }
  HostInterp hl; InterpTable lon1d{};
  if (nlon > 0) { hl = make_interp(lon_cdf, lon_tab, nlon); lon1d = view(hl); }
  for (long i = 0; i < n; ++i) {
    double lon, lat;
    if (lon_in) { lon = lon_in[i]; lat = lat_in[i]; }
    else if (sp->spatial_type == SPATIAL_LON1D) { lon = interp(lon1d, u_lon[i]); lat = 0.0; }
    else uniform_lonlat(*sp, u_sinlat[i], u_lon[i], lon, lat);
    init_packet_finish(*sp, speed, u_time[i], lon, lat, u_speed[i], z_normal[i], u_alt[i],
                       u_az[i], out + 14 * i);
  }
}

static GTables make_gtables(std::vector<HostInterp>& store, int nt, const int* sizes,
                            const double* v, const double* g) {
  GTables G{};
  G.n = nt;
  size_t off = 0;
  store.resize(nt);
  for (int t = 0; t < nt; ++t) {
    store[t] = make_interp(v + off, g + off, sizes[t]);
    G.t[t] = view(store[t]);
    G.f[t] = FastTable{};
    off += sizes[t];
  }
  return G;
}

extern "C" void hc_image(long n, const double* X /* n x 8 */, const ImageParams* ip, int nt,
                         const int* sizes, const double* v, const double* g,
                         double* image, long long* counts) {
  std::vector<HostInterp> store;
  GTables G = make_gtables(store, nt, sizes, v, g);
  const double sx = (ip->x1 - ip->x0) / ip->nx, sz = (ip->z1 - ip->z0) / ip->nz;
  for (long i = 0; i < n; ++i) {
    const double* s = X + 8 * i;
    if (ip->skip_dead && !(s[7] > 0.0)) continue;
    double w;
    int pix = image_packet(*ip, G, sx, sz, s[1], s[2], s[3], s[5], s[7], w);
    if (pix >= 0) { image[pix] += w; counts[pix] += 1; }
  }
}

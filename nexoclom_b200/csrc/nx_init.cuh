// nexoclom_b200 -- initial packet state (K1): surface position, speed and
// direction draws.  Reference: initial_state/source_distribution.py:12-283,
// math/randomdeviates.py:8-83, particle_tracking/Output.py:136-147.
// The reference draws from NumPy's PCG64 / global RandomState; here every packet
// owns a Philox4x32-10 stream keyed by (seed, global packet id), so results do
// not depend on launch geometry or on how packets are sharded over GPUs.
#pragma once
#include "nx_surface.cuh"

namespace nx {

enum SpatialType { SPATIAL_UNIFORM = 0, SPATIAL_MAP = 1, SPATIAL_LON1D = 2 };
enum SpeedType { SPEED_FLAT = 0, SPEED_GAUSSIAN = 1, SPEED_TABLE = 2 };
enum AngularType { ANGULAR_RADIAL = 0, ANGULAR_ISOTROPIC = 1, ANGULAR_2D = 2 };

struct SourceParams {
  int32_t spatial_type, speed_type, angular_type, is_planet;
  double exobase;
  double sinlat0, sinlat1;
  double lon0, lon1;
  double vprob, vsigma, delv;
  double v_scale;
  double sinalt0, sinalt1;
  double az0, az1;
  double endtime;
  int32_t random_time;
  int32_t map_nx, map_ny;
  int32_t map_lat_is_sin;
  double map_fmax;
  // StartPoint = a moon (extension, see RunParams): the source sits on the moon's surface;
  // exobase is in moon radii, the satellite-local frame of xyz_from_lonlat (:21-28: sub-planet
  // point at (0,-1,0), leading point at (-1,0,0)) is turned by the moon's phase at the packet's
  // ejection time and moved to the moon, whose orbital velocity is added.
  int32_t start_is_moon, reserved;
  double moon_a, moon_omega, moon_phi, moon_radius;     // R_p, rad/s, rad, R_p
};

// 2-D source map on a uniform (x, y) grid, row-major [nx][ny]
// (random_deviates_2d re-grids onto linspace(min, max, n), randomdeviates.py:58-59)
struct SourceMap {
  const double* f;
  double x_lo, x_hi, y_lo, y_hi;
};

NX_HD double bilinear(const SourceMap& m, int nx, int ny, double x, double y) {
  const double dx = (m.x_hi - m.x_lo) / (nx - 1), dy = (m.y_hi - m.y_lo) / (ny - 1);
  int i = (int)((x - m.x_lo) / dx), j = (int)((y - m.y_lo) / dy);
  i = i < 0 ? 0 : (i > nx - 2 ? nx - 2 : i);
  j = j < 0 ? 0 : (j > ny - 2 ? ny - 2 : j);
  const double tx = (x - (m.x_lo + i * dx)) / dx, ty = (y - (m.y_lo + j * dy)) / dy;
  const double* r0 = m.f + (size_t)i * ny + j;
  const double* r1 = r0 + ny;
  return r0[0] * (1 - tx) * (1 - ty) + r0[1] * (1 - tx) * ty + r1[0] * tx * (1 - ty) + r1[1] * tx * ty;
}

// uniform surface band (source_distribution.py:47-62)
NX_HD void uniform_lonlat(const SourceParams& sp, double u_sinlat, double u_lon, double& lon,
                          double& lat, double* sinlat_out = nullptr) {
  const double sinlat = add_rn(sp.sinlat0, mul_rn(sub_rn(sp.sinlat1, sp.sinlat0), u_sinlat));
  lat = asin(sinlat);
  lon = fmod(add_rn(sp.lon0, mul_rn(sub_rn(sp.lon1, sp.lon0), u_lon)), NX_TWO_PI);
  if (sinlat_out) *sinlat_out = sinlat;
}

// Everything after the surface point and the deviates are known: a pure transform
// (replayed against the reference's recorded draws in tests/test_hostcheck.py).
// z_normal is only used by the gaussian speed distribution.
NX_HD void init_packet_finish(const SourceParams& sp, const InterpTable& speed, double u_time,
                              double lon, double lat, double u_speed, double z_normal,
                              double u_alt, double u_az, double* x0) {
  // time until the image is taken (Output.py:136-139)
  const double time = sp.random_time ? mul_rn(u_time, sp.endtime) : sp.endtime;

  const double cl = cos(lat);
  const double sx = sp.is_planet ? sp.exobase : -sp.exobase;      // xyz_from_lonlat :18-28
  const double px = mul_rn(mul_rn(sx, sin(lon)), cl);
  const double py = mul_rn(mul_rn(-sp.exobase, cos(lon)), cl);
  const double pz = mul_rn(sp.exobase, sin(lat));
  const double local_time = fmod(add_rn(div_rn(mul_rn(lon, 12.0), NX_PI), 12.0), 24.0);

  // ---- speed (source_distribution.py:137-189) ----
  double v;
  if (sp.speed_type == SPEED_FLAT) {
    v = sub_rn(add_rn(mul_rn(mul_rn(u_speed, 2.0), sp.delv), sp.vprob), sp.delv);
  } else if (sp.speed_type == SPEED_GAUSSIAN) {
    v = (sp.vsigma == 0.0) ? sp.vprob : add_rn(mul_rn(z_normal, sp.vsigma), sp.vprob);
  } else {
    v = interp(speed, u_speed);           // inverse CDF (randomdeviates.py:29-33)
  }
  v = mul_rn(v, sp.v_scale);

  // ---- direction (source_distribution.py:192-252) ----
  double alt, az;
  double d[3];
  if (sp.angular_type == ANGULAR_2D) {
    // '2d' (source_distribution.py:213-222, 253-283): cos(alt) uniform, motion in the
    // equatorial plane; sinalt0/1 carry cos(altitude) of the two limits
    const double cosalt = add_rn(mul_rn(u_alt, sub_rn(sp.sinalt1, sp.sinalt0)), sp.sinalt0);
    alt = acos(cosalt); az = 0.0;
    const double v_rad = sin(alt), v_tan = cos(alt);
    const double rn = sqrt(add_rn(mul_rn(px, px), mul_rn(py, py)));
    const double rx = div_rn(px, rn), ry = div_rn(py, rn);
    const double tx = div_rn(py, rn), ty = div_rn(-px, rn);
    d[0] = add_rn(mul_rn(v_tan, tx), mul_rn(v_rad, rx));
    d[1] = add_rn(mul_rn(v_tan, ty), mul_rn(v_rad, ry));
    d[2] = 0.0;
  } else {
    if (sp.angular_type == ANGULAR_RADIAL) {
      alt = NX_PI / 2.; az = 0.0;
    } else {
      const double sinalt = add_rn(mul_rn(u_alt, sub_rn(sp.sinalt1, sp.sinalt0)), sp.sinalt0);
      alt = asin(sinalt);
      az = add_rn(sp.az0, mul_rn(sub_rn(sp.az1, sp.az0), u_az));
    }
    local_direction(px, py, pz, alt, az, d);
  }

  x0[0] = time; x0[1] = px; x0[2] = py; x0[3] = pz;
  x0[4] = mul_rn(d[0], v); x0[5] = mul_rn(d[1], v); x0[6] = mul_rn(d[2], v);
  x0[7] = 1.0; x0[8] = v; x0[9] = lon; x0[10] = lat; x0[11] = local_time;
  x0[12] = alt; x0[13] = az;
  if (sp.start_is_moon) {
    const double phi = sp.moon_phi - sp.moon_omega * time;
    const double c = cos(phi), s = sin(phi);
    const double lx = x0[1] * sp.moon_radius, ly = x0[2] * sp.moon_radius;
    const double lvx = x0[4], lvy = x0[5];
    const double vorb = sp.moon_a * sp.moon_omega;
    x0[1] = (lx * c - ly * s) - sp.moon_a * s;
    x0[2] = (lx * s + ly * c) + sp.moon_a * c;
    x0[3] = x0[3] * sp.moon_radius;
    x0[4] = (lvx * c - lvy * s) - vorb * c;
    x0[5] = (lvx * s + lvy * c) - vorb * s;
  }
}

// The same transform with the trigonometry folded (device K1 of the fast-arithmetic build; the
// exact-order version above is what the reference's recorded deviates are replayed through):
// sin(asin s) = s, cos(asin s) = sqrt(1 - s^2), one sincos per angle, the modulo of a value known
// to lie in [12, 36) as a subtraction, reciprocals instead of the nine divisions of the local
// frame.  Agrees with the exact version to a few ulp (gate of the parity tests: 1e-12).
// sinlat: sin(lat) when the caller has it (uniform band, sin-latitude maps), else NaN.
NX_HD void init_packet_finish_fast(const SourceParams& sp, const InterpTable& speed, double u_time,
                                   double lon, double lat, double sinlat, double u_speed,
                                   double z_normal, double u_alt, double u_az, double* x0) {
  const double time = sp.random_time ? u_time * sp.endtime : sp.endtime;
  double sl, cl, slon, clon;
  if (sinlat == sinlat) { sl = sinlat; cl = sqrt(fmax(1.0 - sl * sl, 0.0)); }
  else sincos(lat, &sl, &cl);
  sincos(lon, &slon, &clon);
  const double sx = sp.is_planet ? sp.exobase : -sp.exobase;
  const double px = sx * slon * cl, py = -sp.exobase * clon * cl, pz = sp.exobase * sl;
  double local_time = lon * 12.0 / NX_PI + 12.0;                  // lon in [0, 2 pi)
  if (local_time >= 24.0) local_time -= 24.0;

  double v;
  if (sp.speed_type == SPEED_FLAT) v = u_speed * 2.0 * sp.delv + sp.vprob - sp.delv;
  else if (sp.speed_type == SPEED_GAUSSIAN) v = (sp.vsigma == 0.0) ? sp.vprob : z_normal * sp.vsigma + sp.vprob;
  else v = interp(speed, u_speed);
  v *= sp.v_scale;

  double alt, az, d[3];
  if (sp.angular_type == ANGULAR_2D) {
    const double cosalt = u_alt * (sp.sinalt1 - sp.sinalt0) + sp.sinalt0;
    alt = acos(cosalt); az = 0.0;
    const double v_rad = sqrt(fmax(1.0 - cosalt * cosalt, 0.0)), v_tan = cosalt;   // alt in [0, pi]
    const double irn = 1.0 / sqrt(px * px + py * py);
    const double rx = px * irn, ry = py * irn;
    d[0] = v_tan * ry + v_rad * rx;
    d[1] = -v_tan * rx + v_rad * ry;
    d[2] = 0.0;
  } else {
    double v_rad, ca, saz = 0.0, caz = 1.0;
    if (sp.angular_type == ANGULAR_RADIAL) {
      alt = NX_PI / 2.; az = 0.0;
      sincos(alt, &v_rad, &ca);                       // cos(pi/2) as the exact version rounds it
    } else {
      const double sinalt = u_alt * (sp.sinalt1 - sp.sinalt0) + sp.sinalt0;
      alt = asin(sinalt);
      az = sp.az0 + (sp.az1 - sp.az0) * u_az;
      v_rad = sinalt; ca = sqrt(fmax(1.0 - sinalt * sinalt, 0.0));
      sincos(az, &saz, &caz);
    }
    // local_direction (nx_surface.cuh): r = p / |p|, east = (y, -x, 0) / |.|, north = (-zx, -zy, x^2 + y^2) / |.|
    const double v_tan0 = ca * caz, v_tan1 = ca * saz;
    const double h2 = px * px + py * py;
    const double irn = 1.0 / sqrt(h2 + pz * pz), ien = 1.0 / sqrt(h2);
    const double nx_ = -pz * px, ny_ = -pz * py;
    const double inn = 1.0 / sqrt(nx_ * nx_ + ny_ * ny_ + h2 * h2);
    d[0] = v_tan0 * (nx_ * inn) + v_tan1 * (py * ien) + v_rad * (px * irn);
    d[1] = v_tan0 * (ny_ * inn) + v_tan1 * (-px * ien) + v_rad * (py * irn);
    d[2] = v_tan0 * (h2 * inn) + v_rad * (pz * irn);
  }
  x0[0] = time; x0[1] = px; x0[2] = py; x0[3] = pz;
  x0[4] = d[0] * v; x0[5] = d[1] * v; x0[6] = d[2] * v;
  x0[7] = 1.0; x0[8] = v; x0[9] = lon; x0[10] = lat; x0[11] = local_time;
  x0[12] = alt; x0[13] = az;
}

// Fills x0[NCOL_X0] = time,x,y,z,vx,vy,vz,frac,v,longitude,latitude,local_time,
// altitude,azimuth for packet `id`.
// lon1d: inverse CDF of a longitude-only source map (source_distribution.py:72-76 ->
// random_deviates_1d, randomdeviates.py:29-33); latitude is 0 for those maps.
template <bool FAST = false>
NX_HD void init_packet(const SourceParams& sp, const SourceMap& map, const InterpTable& speed,
                       uint64_t seed, uint64_t id, double* x0,
                       const InterpTable& lon1d = InterpTable{}) {
  double u_time, u_sinlat, u_lon, u_speed, u_alt, u_az;
  uniform_pair(seed, id, STREAM_INIT, 0, u_time, u_sinlat);
  uniform_pair(seed, id, STREAM_INIT, 1, u_lon, u_speed);
  uniform_pair(seed, id, STREAM_INIT, 2, u_alt, u_az);

  // ---- position (source_distribution.py:47-62, 96-121) ----
  double lon, lat;
  double sinlat = __builtin_nan("");                  // sin(lat) where it is known without a sin()
  if (sp.spatial_type == SPATIAL_UNIFORM) {
    uniform_lonlat(sp, u_sinlat, u_lon, lon, lat, &sinlat);
  } else if (sp.spatial_type == SPATIAL_LON1D) {
    lon = interp(lon1d, u_lon);
    lat = 0.0;
  } else {
    // acceptance / rejection on the map (randomdeviates.py:61-72)
    uint32_t draw = 4;
    for (;;) {
      double ux, uy, uf, unused;
      uniform_pair(seed, id, STREAM_INIT, draw, ux, uy);
      uniform_pair(seed, id, STREAM_INIT, draw + 1, uf, unused);
      draw += 2;
      const double x = add_rn(mul_rn(ux, sub_rn(map.x_hi, map.x_lo)), map.x_lo);
      const double y = add_rn(mul_rn(uy, sub_rn(map.y_hi, map.y_lo)), map.y_lo);
      if (mul_rn(uf, sp.map_fmax) < bilinear(map, sp.map_nx, sp.map_ny, x, y) || draw > 4000) {
        lon = x;
        lat = sp.map_lat_is_sin ? asin(y) : y;
        if (sp.map_lat_is_sin) sinlat = y;
        break;
      }
    }
  }
  double z = 0.0;
  if (sp.speed_type == SPEED_GAUSSIAN && sp.vsigma != 0.0) {
    double g0, g1;
    uniform_pair(seed, id, STREAM_INIT, 3, g0, g1);
    z = sqrt(-2.0 * log(1.0 - g0)) * cos(NX_TWO_PI * g1);     // Box-Muller on (1-g0) in (0,1]
  }
  // the fast transform covers a planet's own surface (the moon frame and the time a packet of
  // a map that reaches 2 pi would need stay on the exact path)
  if (FAST && !sp.start_is_moon && lon >= 0.0 && lon < NX_TWO_PI)
    init_packet_finish_fast(sp, speed, u_time, lon, lat, sinlat, u_speed, z, u_alt, u_az, x0);
  else
    init_packet_finish(sp, speed, u_time, lon, lat, u_speed, z, u_alt, u_az, x0);
}

}  // namespace nx

// nexoclom_b200 -- K5 with spatial culling.
//
// The reference evaluates every spectrum against the union of ~425 KD-tree balls
// along the boresight and then applies a cone test (compute_iteration.py:151-222).
// The membership rule depends only on (line of sight, packet), so any candidate
// generator that does not miss a member gives identical results.  Here:
//
//   1. live packets are binned into a uniform G^3 cell grid over their bounding
//      cube and copied into cell order (counting sort: count / scan / scatter);
//      the cell id is (ix*G + iy)*G + iz, so the packets of a run of cells along z
//      are one contiguous range;
//   2. one WARP per line of sight marches along the boresight in segments of one
//      cell size; the axis-aligned box around a segment, inflated by the cone
//      radius at its far end, yields a rectangle of (ix, iy) columns and a z-range;
//      the warp streams the contiguous packet range of every column (coalesced);
//   3. a packet is tested in exactly one segment -- the one that contains its
//      axial coordinate losrad -- so overlapping boxes never double count;
//   4. survivors of that ownership test go through the same exact membership /
//      weighting code as the brute-force kernel (nx_image.cuh: los_hit, los_weight).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "nx_image.cuh"
#include "nx_kernels.h"

namespace nx {

#define FULL_MASK 0xffffffffu
#ifndef NX_LOS_SEG_CELLS
#define NX_LOS_SEG_CELLS 6.0      // march segment length in finest local cells (1: 41 ms, 4: 35.2, 8: 35.4, 16: 46.7)
#endif

// ---- 1. bounding cube: max |coordinate| over live packets ----------------------
__global__ void __launch_bounds__(256)
k_los_extent(StateCols P, long long n, int skip_dead, int round32, unsigned long long* out) {
  // out[0] = bits of max |coordinate|, out[1] = number of packets that enter the grid
  double m = 0.0;
  unsigned long long live = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    if (skip_dead && !(P.c[7][i] > 0.0)) continue;
    double x = P.c[1][i], y = P.c[2][i], z = P.c[3][i];
    if (round32) { x = round_f32(x); y = round_f32(y); z = round_f32(z); }
    m = fmax(m, fmax(fabs(x), fmax(fabs(y), fabs(z))));
    ++live;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = fmax(m, __shfl_xor_sync(FULL_MASK, m, o));
    live += __shfl_xor_sync(FULL_MASK, live, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(out, (unsigned long long)__double_as_longlong(m));
    if (live) atomicAdd(out + 1, live);
  }
}

__device__ __forceinline__ int cell_of(double v, const LosGrid& g) {
  int c = __double2int_rd(fma(g.k, asinh(v * g.inv_scale), 0.5 * g.G));
  return max(0, min(c, g.G - 1));
}
// width of the cell that holds coordinate v
__device__ __forceinline__ double cell_width(double v, const LosGrid& g) {
  return sqrt(fma(v, v, g.scale * g.scale)) / g.k;
}

// ---- 2. counting sort into cell order -------------------------------------------
__global__ void __launch_bounds__(256)
k_los_cell_count(StateCols P, long long n, LosGrid g, int skip_dead, int round32,
                 unsigned* __restrict__ cell_id, unsigned* __restrict__ count) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    unsigned id = 0xffffffffu;
    if (!(skip_dead && !(P.c[7][i] > 0.0))) {
      double x = P.c[1][i], y = P.c[2][i], z = P.c[3][i];
      if (round32) { x = round_f32(x); y = round_f32(y); z = round_f32(z); }
      const int ix = cell_of(x, g), iy = cell_of(y, g), iz = cell_of(z, g);
      id = (unsigned)((ix * g.G + iy) * g.G + iz);
      atomicAdd(&count[id], 1u);
    }
    cell_id[i] = id;
  }
}

// exclusive scan of `count` (ncell entries) into start[0..ncell]; three passes
#define NX_SCAN_ELEMS 4096
__global__ void __launch_bounds__(1024)
k_scan_partial(const unsigned* __restrict__ in, unsigned* __restrict__ out,
               unsigned* __restrict__ block_sum, int n) {
  __shared__ unsigned warp_tot[32];
  const int base = blockIdx.x * NX_SCAN_ELEMS + threadIdx.x * 4;
  unsigned v[4], s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) { v[k] = (base + k < n) ? in[base + k] : 0u; s += v[k]; }
  unsigned incl = s;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(FULL_MASK, incl, o);
    if (lane >= (unsigned)o) incl += t;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    unsigned w = warp_tot[lane], wi = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(FULL_MASK, wi, o);
      if (lane >= (unsigned)o) wi += t;
    }
    warp_tot[lane] = wi - w;                  // exclusive warp offsets
    if (lane == 31) block_sum[blockIdx.x] = wi;
  }
  __syncthreads();
  unsigned run = warp_tot[warp] + incl - s;   // exclusive prefix of this thread
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
}

__global__ void __launch_bounds__(1024)
k_scan_top(unsigned* __restrict__ block_sum, int nblocks, unsigned* __restrict__ total) {
  // single block: sequential over chunks of 1024 (nblocks <= a few thousand)
  __shared__ unsigned warp_tot[32];
  __shared__ unsigned carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = 0; b0 < nblocks; b0 += 1024) {
    const int i = b0 + threadIdx.x;
    const unsigned v = (i < nblocks) ? block_sum[i] : 0u;
    unsigned incl = v;
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(FULL_MASK, incl, o);
      if (lane >= (unsigned)o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      unsigned w = warp_tot[lane], wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(FULL_MASK, wi, o);
        if (lane >= (unsigned)o) wi += t;
      }
      warp_tot[lane] = wi - w;
    }
    __syncthreads();
    const unsigned excl = carry + warp_tot[warp] + incl - v;
    if (i < nblocks) block_sum[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(1024)
k_scan_add(unsigned* __restrict__ out, const unsigned* __restrict__ block_sum, int n,
           const unsigned* __restrict__ total) {
  const int base = blockIdx.x * NX_SCAN_ELEMS + threadIdx.x * 4;
  const unsigned add = block_sum[blockIdx.x];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (base + k < n) out[base + k] += add;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *total;
}

__global__ void __launch_bounds__(256)
k_los_cell_scatter(StateCols P, long long n, int round32, const unsigned* __restrict__ cell_id,
                   const unsigned* __restrict__ start, unsigned* __restrict__ cursor,
                   LosSorted S) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const unsigned id = cell_id[i];
    if (id == 0xffffffffu) continue;
    const unsigned pos = start[id] + atomicAdd(&cursor[id], 1u);
    double x = P.c[1][i], y = P.c[2][i], z = P.c[3][i], vy = P.c[5][i], fr = P.c[7][i];
    if (round32) { x = round_f32(x); y = round_f32(y); z = round_f32(z); vy = round_f32(vy); fr = round_f32(fr); }
    S.pos[pos] = make_double4(x, y, z, vy); S.frac[pos] = fr;
    S.idx[pos] = (unsigned)i;
  }
}

// ---- 3. one warp per line of sight ------------------------------------------------
__global__ void __launch_bounds__(128)
k_los_grid(LosSorted S, LosGrid g, const unsigned* __restrict__ start, long long nlos,
           const double* __restrict__ los, const double* __restrict__ dist_plan,
           const int* __restrict__ nball, const double* __restrict__ ladder,
           const double* __restrict__ wid2, LosParams lp, LosConsts lc, GTables G,
           double* __restrict__ radiance, unsigned long long* __restrict__ npack,
           unsigned char* __restrict__ included,
           unsigned long long* __restrict__ nused, const long long* __restrict__ used_off,
           unsigned long long* __restrict__ used_cursor, unsigned* __restrict__ used_idx,
           const unsigned* __restrict__ order) {
  const long long w_ = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w_ >= nlos) return;
  // lines of sight are processed in a spatially coherent order (host: Morton order of the
  // points of closest approach) so that neighbouring warps stream the same cells out of L2
  const long long l = order ? (long long)order[w_] : w_;
  const unsigned lane = threadIdx.x & 31u;
  LosRay L;
  L.xs = los[l]; L.ys = los[nlos + l]; L.zs = los[2 * nlos + l];
  L.bx = los[3 * nlos + l]; L.by = los[4 * nlos + l]; L.bz = los[5 * nlos + l];
  L.dist_plan = dist_plan[l];
  L.nball = nball[l];

  // farthest axial distance at which a packet can still be a member: planet
  // truncation, last KD ball, and the far side of the packet cube
  const double s2 = sin(2.0 * lp.dphi);
  double t_end = ladder[L.nball - 1] * (1.0 + s2);
  if (L.dist_plan < t_end) t_end = L.dist_plan;
  const double reach = sqrt(L.xs * L.xs + L.ys * L.ys + L.zs * L.zs) + 1.7320508075688772 * g.half;
  if (reach < t_end) t_end = reach;
  const double tan_phi = tan(lp.dphi) * (1.0 + 1e-9);

  double rad = 0.0;
  unsigned long long cnt = 0, used = 0;
  double t1 = 0.0;
  while (t1 < t_end) {
    // segment [t0, t1): about one cell of the finest axis long, never shorter than the
    // cone is wide there; t1 of one segment IS t0 of the next (same double)
    const double t0 = t1;
    const double cx = L.xs + L.bx * t0, cy = L.ys + L.by * t0, cz = L.zs + L.bz * t0;
    const double dt = fmax(NX_LOS_SEG_CELLS *
                               fmin(cell_width(cx, g), fmin(cell_width(cy, g), cell_width(cz, g))),
                           2.0 * t0 * tan_phi);
    t1 = t0 + dt;
    const double rho = t1 * tan_phi + 1e-9 * (1.0 + t1);
    const double ax = L.xs + L.bx * t0, bx = L.xs + L.bx * t1;
    const double ay = L.ys + L.by * t0, by = L.ys + L.by * t1;
    const double az = L.zs + L.bz * t0, bz = L.zs + L.bz * t1;
    const double xlo = fmin(ax, bx) - rho, xhi = fmax(ax, bx) + rho;
    const double ylo = fmin(ay, by) - rho, yhi = fmax(ay, by) + rho;
    const double zlo = fmin(az, bz) - rho, zhi = fmax(az, bz) + rho;
    // boxes entirely outside the packet cube hold nothing
    if (xlo > g.half || xhi < -g.half || ylo > g.half || yhi < -g.half || zlo > g.half ||
        zhi < -g.half)
      continue;
    const int ix0 = cell_of(xlo, g), ix1 = cell_of(xhi, g);
    const int iy0 = cell_of(ylo, g), iy1 = cell_of(yhi, g);
    const int iz0 = cell_of(zlo, g), iz1 = cell_of(zhi, g);
    for (int ix = ix0; ix <= ix1; ++ix) {
      for (int iy = iy0; iy <= iy1; ++iy) {
        const unsigned row = (unsigned)((ix * g.G + iy) * g.G);
        const unsigned p0 = start[row + iz0], p1 = start[row + iz1 + 1];
        for (unsigned q = p0 + lane; q < p1; q += 32) {
          const double4 rec = S.pos[q];
          const double px = rec.x, py = rec.y, pz = rec.z;
          // ownership: the segment that contains the packet's axial coordinate
          // (same arithmetic as los_hit so that every packet has one owner)
          const double rx = sub_rn(px, L.xs), ry = sub_rn(py, L.ys), rz = sub_rn(pz, L.zs);
          const double lr = add_rn(add_rn(mul_rn(rx, L.bx), mul_rn(ry, L.by)), mul_rn(rz, L.bz));
          if (!(lr >= t0 && lr < t1)) continue;
          double losrad, dist;
          if (los_hit(L, lp.dphi, lc.cos_margin2, lc.cos_accept2, lc.cover, ladder, wid2,
                    lc.inv_log_ratio, lc.log_t0,
                      lc.kwin, px, py, pz, losrad, dist)) {
            ++cnt;
            const double w = los_weight(L, lp, G, lc.sin_dphi, S.frac[q], rec.w, losrad, dist);
            rad += w;
            if (included) included[S.idx[q]] = 1;
            if (w > 0.0) {                    // `used` packets (compute_iteration.py:210-211)
              ++used;
              if (used_idx) used_idx[used_off[l] + atomicAdd(&used_cursor[l], 1ull)] = S.idx[q];
            }
          }
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    rad += __shfl_xor_sync(FULL_MASK, rad, o);
    cnt += __shfl_xor_sync(FULL_MASK, cnt, o);
    used += __shfl_xor_sync(FULL_MASK, used, o);
  }
  if (lane == 0) {
    if (radiance) radiance[l] += rad;
    if (npack) npack[l] += cnt;
    if (nused) nused[l] = used;
  }
}

// ---- host-side launch sequence ------------------------------------------------------
cudaError_t launch_los_grid_build(cudaStream_t st, int device, StateCols P, long long n,
                                  const LosParams& lp, LosGridWork& w) {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const int blocks = sms * 8;
  cudaError_t e;
  if ((e = cudaMemsetAsync(w.extent_bits, 0, 2 * sizeof(unsigned long long), st)) != cudaSuccess) return e;
  k_los_extent<<<blocks, 256, 0, st>>>(P, n, lp.skip_dead, lp.round_f32, w.extent_bits);
  unsigned long long bits[2] = {0, 0};
  if ((e = cudaMemcpyAsync(bits, w.extent_bits, sizeof(bits), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return e;
  if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return e;
  double ext;
  memcpy(&ext, &bits[0], sizeof(ext));
  if (!(ext > 0.0)) ext = 1.0;
  // cells per axis: fixed by the option "los_grid", else ~0.55 n^(1/3) in steps of 32
  // (measured on 1e5 lines of sight: 1.9e6 packets -> 64 best, 1e7 -> 96..128)
  int G = w.G_fixed;
  if (G <= 0) {
    G = (int)(0.55 * cbrt((double)(bits[1] ? bits[1] : 1)) / 32.0 + 0.5) * 32;
    G = G < 32 ? 32 : (G > NX_LOS_GRID_MAX ? NX_LOS_GRID_MAX : G);
  }
  w.G = G;
  w.grid.G = w.G;
  w.grid.half = ext * (1.0 + 1e-9) + 1e-9;
  // scale: one planet radius (cells ~0.06 R_p at the planet for G = 128, half = 25);
  // a cloud smaller than that gets a nearly uniform grid
  w.grid.scale = w.grid.half < w.scale ? w.grid.half : w.scale;
  w.grid.inv_scale = 1.0 / w.grid.scale;
  w.grid.k = 0.5 * w.G / asinh(w.grid.half * w.grid.inv_scale);
  const int ncell = w.G * w.G * w.G;
  if ((e = cudaMemsetAsync(w.count, 0, (size_t)ncell * sizeof(unsigned), st)) != cudaSuccess) return e;
  k_los_cell_count<<<blocks, 256, 0, st>>>(P, n, w.grid, lp.skip_dead, lp.round_f32, w.cell_id, w.count);
  const int nb = (ncell + NX_SCAN_ELEMS - 1) / NX_SCAN_ELEMS;
  k_scan_partial<<<nb, 1024, 0, st>>>(w.count, w.start, w.block_sum, ncell);
  k_scan_top<<<1, 1024, 0, st>>>(w.block_sum, nb, w.total);
  k_scan_add<<<nb, 1024, 0, st>>>(w.start, w.block_sum, ncell, w.total);
  if ((e = cudaMemsetAsync(w.count, 0, (size_t)ncell * sizeof(unsigned), st)) != cudaSuccess) return e;
  k_los_cell_scatter<<<blocks, 256, 0, st>>>(P, n, lp.round_f32, w.cell_id, w.start, w.count, w.sorted);
  return cudaGetLastError();
}

cudaError_t launch_los_grid(cudaStream_t st, const LosGridWork& w, long long nlos,
                            const double* los, const double* dist_plan, const int* nball,
                            const double* ladder, const double* wid2, const LosParams& lp,
                            const LosConsts& lc, const GTables& G, double* radiance,
                            unsigned long long* npack, unsigned char* included,
                            unsigned long long* nused, const long long* used_off,
                            unsigned long long* used_cursor, unsigned* used_idx,
                            const unsigned* order) {
  const long long threads = nlos * 32;
  const long long blocks = (threads + 127) / 128;
  k_los_grid<<<(unsigned)blocks, 128, 0, st>>>(w.sorted, w.grid, w.start, nlos, los, dist_plan,
                                               nball, ladder, wid2, lp, lc, G, radiance, npack,
                                               included, nused, used_off, used_cursor, used_idx, order);
  return cudaGetLastError();
}

}  // namespace nx

// nexoclom_b200 -- device-side `Output.save` (reference particle_tracking/Output.py:522-543):
// drop the frac == 0 rows when `compress`, round every column to float32, keep the packet
// index -- done on the GPU, so that a run's packets can stay resident for ModelImage /
// LOSResult in the same process and only the surviving ~1 % ever crosses PCIe (as f32) when
// the run is written to disk.
//
// Stable stream compaction in three launches: live rows per 2048-row tile, exclusive scan of
// the tile counts (one block), scatter with ballot prefix sums.  HBM-bound: reads the frac
// column twice and the seven other columns of the live rows once.
#include <cuda_runtime.h>
#include <stdint.h>

#include "nx_kernels.h"

namespace nx {

#define NX_CMP_THREADS 256
#define NX_CMP_ITEMS 8
#define NX_CMP_TILE (NX_CMP_THREADS * NX_CMP_ITEMS)

__global__ void __launch_bounds__(NX_CMP_THREADS)
k_compact_count(const double* __restrict__ frac, long long n, int skip_dead,
                unsigned* __restrict__ tile_count) {
  __shared__ unsigned warp_sum[NX_CMP_THREADS / 32];
  const long long base = (long long)blockIdx.x * NX_CMP_TILE;
  unsigned c = 0;
#pragma unroll
  for (int j = 0; j < NX_CMP_ITEMS; ++j) {
    const long long i = base + (long long)j * NX_CMP_THREADS + threadIdx.x;
    if (i < n && (!skip_dead || frac[i] > 0.0)) ++c;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned t = 0;
    for (int w = 0; w < NX_CMP_THREADS / 32; ++w) t += warp_sum[w];
    tile_count[blockIdx.x] = t;
  }
}

// exclusive scan of tile_count[ntiles] in place; total -> *total  (one 1024-thread block)
__global__ void __launch_bounds__(1024)
k_compact_scan(unsigned* __restrict__ tile_count, long long ntiles,
               unsigned long long* __restrict__ total) {
  __shared__ unsigned warp_tot[32];
  __shared__ unsigned carry, chunk_total;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (long long first = 0; first < ntiles; first += 1024) {
    const long long i = first + threadIdx.x;
    const unsigned v = i < ntiles ? tile_count[i] : 0u;
    unsigned incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (unsigned)o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const unsigned w = warp_tot[lane];
      unsigned wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= (unsigned)o) wi += t;
      }
      warp_tot[lane] = wi - w;                    // exclusive offset of each warp
      if (lane == 31) chunk_total = wi;
    }
    __syncthreads();
    if (i < ntiles) tile_count[i] = carry + warp_tot[warp] + incl - v;
    __syncthreads();
    if (threadIdx.x == 0) carry += chunk_total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(NX_CMP_THREADS)
k_compact_scatter(StateCols P, long long n, int skip_dead, int to_f32, int have_step,
                  const unsigned* __restrict__ tile_offset, double* __restrict__ out,
                  size_t out_stride, unsigned* __restrict__ index) {
  __shared__ unsigned warp_cnt[NX_CMP_THREADS / 32];
  __shared__ unsigned running;
  const long long base = (long long)blockIdx.x * NX_CMP_TILE;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) running = tile_offset[blockIdx.x];
  __syncthreads();
#pragma unroll 1
  for (int j = 0; j < NX_CMP_ITEMS; ++j) {
    const long long i = base + (long long)j * NX_CMP_THREADS + threadIdx.x;
    const double f = i < n ? P.c[7][i] : 0.0;
    const bool keep = i < n && (!skip_dead || f > 0.0);
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    unsigned off = running;
    for (unsigned w = 0; w < warp; ++w) off += warp_cnt[w];
    if (keep) {
      const size_t o = (size_t)off + __popc(m & ((1u << lane) - 1u));
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        double v = (k == 7) ? f : P.c[k][i];
        if (to_f32) v = (double)(float)v;
        out[(size_t)k * out_stride + o] = v;
      }
      {                              // 9th column: the adaptive driver's step size (Output.py:246)
        double v = have_step ? P.c[8][i] : 1000.0;
        if (to_f32) v = (double)(float)v;
        out[(size_t)8 * out_stride + o] = v;
      }
      index[o] = (unsigned)i;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned t = 0;
      for (int w = 0; w < NX_CMP_THREADS / 32; ++w) t += warp_cnt[w];
      running += t;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
k_to_f32(const double* __restrict__ src, size_t src_stride, long long n, int ncols,
         float* __restrict__ dst) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < ncols; ++k) dst[(size_t)k * n + i] = (float)src[(size_t)k * src_stride + i];
}

long long compact_tiles(long long n) { return (n + NX_CMP_TILE - 1) / NX_CMP_TILE; }

cudaError_t launch_compact_count(cudaStream_t st, const double* frac, long long n, int skip_dead,
                                 unsigned* tile_count, unsigned long long* total) {
  const long long nt = compact_tiles(n);
  k_compact_count<<<(unsigned)nt, NX_CMP_THREADS, 0, st>>>(frac, n, skip_dead, tile_count);
  k_compact_scan<<<1, 1024, 0, st>>>(tile_count, nt, total);
  return cudaGetLastError();
}

cudaError_t launch_compact_scatter(cudaStream_t st, StateCols P, long long n, int skip_dead,
                                   int to_f32, int have_step, const unsigned* tile_offset,
                                   double* out, size_t out_stride, unsigned* index) {
  k_compact_scatter<<<(unsigned)compact_tiles(n), NX_CMP_THREADS, 0, st>>>(
      P, n, skip_dead, to_f32, have_step, tile_offset, out, out_stride, index);
  return cudaGetLastError();
}

cudaError_t launch_to_f32(cudaStream_t st, const double* src, size_t src_stride, long long n,
                          int ncols, float* dst) {
  k_to_f32<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, src_stride, n, ncols, dst);
  return cudaGetLastError();
}

}  // namespace nx

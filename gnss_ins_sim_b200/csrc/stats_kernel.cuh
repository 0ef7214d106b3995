// K3: ensemble error statistics over runs -- InsDataMgr.__array_stats
// (ins_data_manager.py:797-808): max|e|, mean, std (ddof 0, two-pass like np.std).
// Deterministic two-stage reductions (no floating-point atomics): stage 1 writes one
// partial per block, stage 2 (one block) folds them in a fixed order.
#pragma once
#include "common.cuh"

namespace b2ins {

constexpr int kStatBlocks = 128;   // stage-1 grid
constexpr int kStatMaxComp = 32;

__host__ __device__ inline int stat_threads(int ncomp) { return ncomp * (1024 / ncomp >= 32 ? 32 : 1024 / ncomp); }

// MODE 0: sum e and max|e| ; MODE 1: sum (e - mean)^2
template <int MODE>
__global__ void err_stage1_kernel(int64_t runs, int ncomp, const double* __restrict__ err,
                                  const double* __restrict__ mean, double* __restrict__ ws) {
  extern __shared__ double sh[];  // [threads] sums, [threads] maxes
  const int threads = blockDim.x;
  const int c = threadIdx.x % ncomp;   // blockDim.x and the grid stride are multiples of ncomp
  const int64_t total = runs * ncomp;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * threads;
  double acc = 0.0, mx = 0.0;
  const double mu = (MODE == 1) ? mean[c] : 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * threads + threadIdx.x; i < total; i += stride) {
    const double e = err[i];
    if (MODE == 0) {
      acc += e;
      mx = fmax(mx, fabs(e));
    } else {
      const double d = e - mu;
      acc += d * d;
    }
  }
  sh[threadIdx.x] = acc;
  sh[threads + threadIdx.x] = mx;
  __syncthreads();
  if (threadIdx.x < ncomp) {
    double s = 0.0, m = 0.0;
    for (int k = threadIdx.x; k < threads; k += ncomp) {
      s += sh[k];
      m = fmax(m, sh[threads + k]);
    }
    ws[(static_cast<int64_t>(blockIdx.x) * 2) * ncomp + c] = s;
    ws[(static_cast<int64_t>(blockIdx.x) * 2 + 1) * ncomp + c] = m;
  }
}

// fold the per-block partials: out[0..ncomp) = sum, out[ncomp..2ncomp) = max (MODE 0 only)
template <int MODE>
__global__ void err_stage2_kernel(int nblocks, int ncomp, const double* __restrict__ ws,
                                  double* __restrict__ out) {
  const int c = threadIdx.x;
  if (c >= ncomp) return;
  double s = 0.0, m = 0.0;
  for (int b = 0; b < nblocks; ++b) {
    s += ws[(static_cast<int64_t>(b) * 2) * ncomp + c];
    m = fmax(m, ws[(static_cast<int64_t>(b) * 2 + 1) * ncomp + c]);
  }
  out[c] = s;
  if (MODE == 0) out[ncomp + c] = m;
}

// single-shard finalisation helpers
__global__ void stats_mean_kernel(int64_t runs, int ncomp, const double* __restrict__ partial,
                                  double* __restrict__ stats) {
  const int c = threadIdx.x;
  if (c >= ncomp) return;
  stats[c] = partial[ncomp + c];                                  // max|e|
  stats[ncomp + c] = partial[c] / static_cast<double>(runs);      // mean
}
__global__ void stats_std_kernel(int64_t runs, int ncomp, const double* __restrict__ partial2,
                                 double* __restrict__ stats) {
  const int c = threadIdx.x;
  if (c >= ncomp) return;
  stats[2 * ncomp + c] = sqrt(partial2[c] / static_cast<double>(runs));
}

// Small ensembles (runs * ncomp <= kStatSmallMax): everything in ONE block and one launch --
// sum/max, mean, second pass, std -- with the same fixed-order folding as the staged path.
constexpr int kStatSmallMax = 1 << 17;
constexpr int kStatSmallThreads = 1024;

__global__ void __launch_bounds__(kStatSmallThreads)
stats_small_kernel(int64_t runs, int ncomp, const double* __restrict__ err, double* __restrict__ stats) {
  __shared__ double sh[2 * kStatSmallThreads];
  __shared__ double mean_sh[kStatMaxComp];
  const int threads = (kStatSmallThreads / ncomp) * ncomp;   // a multiple of ncomp
  const int c = threadIdx.x % ncomp;
  const int64_t total = runs * ncomp;
  const bool on = threadIdx.x < threads;
  double acc = 0.0, mx = 0.0;
  if (on)
    for (int64_t i = threadIdx.x; i < total; i += threads) {
      const double e = err[i];
      acc += e;
      mx = fmax(mx, fabs(e));
    }
  sh[threadIdx.x] = acc;
  sh[kStatSmallThreads + threadIdx.x] = mx;
  __syncthreads();
  if (threadIdx.x < ncomp) {
    double s = 0.0, m = 0.0;
    for (int k = threadIdx.x; k < threads; k += ncomp) {
      s += sh[k];
      m = fmax(m, sh[kStatSmallThreads + k]);
    }
    stats[c] = m;
    const double mean = s / static_cast<double>(runs);
    stats[ncomp + c] = mean;
    mean_sh[c] = mean;
  }
  __syncthreads();
  const double mu = mean_sh[c];
  acc = 0.0;
  if (on)
    for (int64_t i = threadIdx.x; i < total; i += threads) {
      const double d = err[i] - mu;
      acc += d * d;
    }
  sh[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < ncomp) {
    double s = 0.0;
    for (int k = threadIdx.x; k < threads; k += ncomp) s += sh[k];
    stats[2 * ncomp + c] = sqrt(s / static_cast<double>(runs));
  }
}

}  // namespace b2ins

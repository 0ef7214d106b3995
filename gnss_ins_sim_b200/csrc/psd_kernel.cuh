// K5: vibration time series from a single-sided PSD -- time_series_from_psd
// (gnss_ins_sim/psd/time_series_from_psd.py:17-65), for runs x 3 axes of one sensor.
//
//   N   = n rounded up to even, capped at 16384 (:36-43);  L = N/2 + 1
//   sxx = np.interp(linspace(0, fs/2, L), freq, sxx) unless len(freq) == L (:45-50)
//   sxx[1:L-1] *= 0.5 ; ax = sqrt(sxx N fs) ; phi = pi randn(L) ; xk = ax exp(j phi) (:51-54)
//   x = real(ifft([xk, conj(xk[-2:0:-1])])) (:55-57), tiled to n by the consumer (t % N)
//
// The real part of that inverse DFT is a cosine synthesis,
//   x[m] = (1/N) [ A_0 + (-1)^m A_{L-1} + 2 sum_{k=1}^{L-2} (A_k cos(2 pi k m / N) - B_k sin(2 pi k m / N)) ]
// with A_k + j B_k = ax_k exp(j phi_k).  Kernel 1 draws the phases (Philox normal z0 of
// (t = k, draw = kDrawPsd + 3 sensor + axis, run)) and writes (A_k, B_k); kernel 2 evaluates the
// sum for one output sample per thread, rotating (cos, sin) by 2 pi m / N per k and
// re-seeding it exactly (integer k m mod N, sincospi) every kPsdReseed terms.  N need not be a
// power of two (N = n for short runs), which is why this is a direct synthesis and not an FFT.
#pragma once
#include "common.cuh"

namespace b2ins {

constexpr int kPsdThreads = 256;
constexpr int kPsdReseed = 32;
constexpr int kPsdChunk = 1024;  // (A,B) pairs staged per shared-memory chunk

struct PsdParams {
  double fs;
  int64_t runs, run_offset;
  int N, L, L0, sensor;
  int interp;          // 1: interpolate the table to L points
  uint32_t k0, k1;
  const double* freq;  // [L0]
  const double* sxx;   // [3][L0]
  double* ab;          // workspace [runs][3][L][2]
  double* series;      // [runs][3][N]
};

// np.interp(x, xp, fp): linear, clamped to the end values
__device__ __forceinline__ double interp_clamped(double x, const double* xp, const double* fp, int n) {
  if (x <= xp[0]) return fp[0];
  if (x >= xp[n - 1]) return fp[n - 1];
  int lo = 0, hi = n - 1;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (xp[mid] <= x)
      lo = mid;
    else
      hi = mid;
  }
  const double slope = (fp[hi] - fp[lo]) / (xp[hi] - xp[lo]);
  return slope * (x - xp[lo]) + fp[lo];
}

__global__ void __launch_bounds__(kPsdThreads) psd_phase_kernel(const __grid_constant__ PsdParams p) {
  const int64_t series = blockIdx.y;  // run * 3 + axis
  const int64_t run = series / 3;
  const int axis = static_cast<int>(series % 3);
  const int k = blockIdx.x * kPsdThreads + threadIdx.x;
  if (k >= p.L) return;
  const int64_t grun = p.run_offset + run;
  const double* tab = p.sxx + static_cast<int64_t>(axis) * p.L0;
  double sx;
  if (p.interp) {
    // np.linspace(0, fs/2, L)[k] = k * step, step = (fs/2) / (L - 1) (endpoint exact)
    const double stop = p.fs / 2.0;
    const double fk = (k == p.L - 1) ? stop : k * (stop / static_cast<double>(p.L - 1));
    sx = interp_clamped(fk, p.freq, tab, p.L0);
  } else {
    sx = tab[k];
  }
  if (k >= 1 && k < p.L - 1) sx = 0.5 * sx;
  const double ax = sqrt(sx * static_cast<double>(p.N) * p.fs);
  const Normal2 z = normal_pair(static_cast<uint32_t>(k), kDrawPsd + 3 * p.sensor + axis,
                                static_cast<uint32_t>(grun), static_cast<uint32_t>(grun >> 32), p.k0, p.k1);
  double s, c;
  sincos(kPi * z.z0, &s, &c);  // phi = math.pi * randn
  double* o = p.ab + (series * p.L + k) * 2;
  o[0] = ax * c;
  o[1] = ax * s;
}

__global__ void __launch_bounds__(kPsdThreads) psd_synth_kernel(const __grid_constant__ PsdParams p) {
  __shared__ double sh[kPsdChunk * 2];
  const int64_t series = blockIdx.y;
  const int m = blockIdx.x * kPsdThreads + threadIdx.x;
  const bool live = m < p.N;
  const double* ab = p.ab + series * p.L * 2;
  const int mm = live ? m : 0;
  // rotation by delta = 2 pi m / N
  double sd, cd;
  sincospi(2.0 * static_cast<double>(mm) / static_cast<double>(p.N), &sd, &cd);
  double acc = 0.0;
  double c = 1.0, s = 0.0;
  for (int k0 = 0; k0 < p.L; k0 += kPsdChunk) {
    const int cnt = min(kPsdChunk, p.L - k0);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * 2; i += kPsdThreads) sh[i] = ab[k0 * 2 + i];
    __syncthreads();
    for (int i = 0; i < cnt; ++i) {
      const int k = k0 + i;
      if ((k % kPsdReseed) == 0) {
        // exact angle: 2 pi ((k m) mod N) / N
        const int64_t r = (static_cast<int64_t>(k) * mm) % p.N;
        sincospi(2.0 * static_cast<double>(r) / static_cast<double>(p.N), &s, &c);
      }
      const double w = (k == 0 || k == p.L - 1) ? 1.0 : 2.0;
      // Re((A + jB)(c + js)) = A c - B s
      acc = fma(w, fma(sh[2 * i], c, -sh[2 * i + 1] * s), acc);
      const double cn = fma(c, cd, -s * sd);
      s = fma(s, cd, c * sd);
      c = cn;
    }
  }
  if (live) p.series[series * p.N + m] = acc / static_cast<double>(p.N);
}

inline int psd_series_len(int64_t n) {
  int64_t N = n;
  if (n % 2 != 0) N = n + 1;
  if (N > 16384) N = 16384;
  return static_cast<int>(N);
}

}  // namespace b2ins

// K1: IMU sensor-error generator, materialised -- pathgen.acc_gen / gyro_gen / bias_drift
// (pathgen.py:441-594).  Time-parallel: one CTA per run walks the samples in tiles of
// kNoiseThreads; thread i draws the normals of sample tile0+i (Philox4x32-10, Box-Muller),
// the first-order Gauss-Markov drift d[t+1] = a d[t] + b z[t] is an affine block scan
// (warp shuffles + one shared-memory hop) with the carry kept across tiles, and the
// measurements are written with the caller's strides.
#pragma once
#include "mc_kernel.cuh"

namespace b2ins {

constexpr int kNoiseThreads = 256;
constexpr int kNoiseWarps = kNoiseThreads / 32;

struct NoiseParams {
  int64_t n, runs, run_offset;
  double dt;
  uint32_t k0, k1;
  TriadNoise gyro, accel;
  const double* ref_gyro;
  const double* ref_accel;
  double* out_gyro;
  double* out_accel;
  int64_t osr, ost, osc;
  double* z_dump;  // [runs][n][12] or null
  // time segmentation (few runs, long series): blockIdx.x = run * nseg + seg, samples
  // [seg*seg_len, min(n, (seg+1)*seg_len)).  pass 0: write outputs, GM state at the segment
  // start taken from seg_carry[run][seg][6] (all zero for seg 0); pass 1: no output, only the
  // zero-state GM response at the segment end -> seg_end[run][seg][6].  Pass 1 covers segments
  // 0 .. nseg-2 only (the end value of the last one is never used: blockIdx.x = run * (nseg-1) +
  // seg) and only their last pass1_len samples: older drives have decayed below 1e-20 of the
  // state (pass1_len = seg_len if the correlation time is too long for that)
  int64_t seg_len;
  int64_t pass1_len;
  int nseg;
  int pass;
  double* seg_carry;
  double* seg_end;
};

__global__ void __launch_bounds__(kNoiseThreads, 2) imu_noise_kernel(const __grid_constant__ NoiseParams p) {
  __shared__ double apow[6][kNoiseThreads + 1];
  __shared__ double sh_w[6][kNoiseWarps];     // inclusive warp totals of the six channels
  __shared__ double sh_pre[6][kNoiseWarps + 1];   // zero-state value at the end of warps < w; [..][8] = tile
  const int segs = (p.pass == 1) ? p.nseg - 1 : p.nseg;
  const int64_t run = blockIdx.x / segs;
  const int seg = static_cast<int>(blockIdx.x % segs);
  const int64_t seg_hi = min64(p.n, (seg + 1) * p.seg_len);
  const int64_t seg_lo = (p.pass == 1) ? seg_hi - p.pass1_len : seg * p.seg_len;
  const int64_t grun = p.run_offset + run;
  const uint32_t run_lo = static_cast<uint32_t>(grun), run_hi = static_cast<uint32_t>(grun >> 32);
  const int i = threadIdx.x;
  for (int k = threadIdx.x; k <= kNoiseThreads; k += kNoiseThreads) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      apow[c][k] = pow(p.accel.gm_a[c], static_cast<double>(k));
      apow[3 + c][k] = pow(p.gyro.gm_a[c], static_cast<double>(k));
    }
  }
  double phase[3] = {0.0, 0.0, 0.0};
  if (p.gyro.vib_type == 2) {
#pragma unroll
    for (int c = 0; c < 3; ++c)
      phase[c] = (uniform01(0xFFFFFFFFu, kDrawPhase + c, run_lo, run_hi, p.k0, p.k1) * 2.0) * kPi;
  }
  double carry[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};  // d at the first sample of the tile
  if (p.pass == 0 && p.seg_carry) {
#pragma unroll
    for (int c = 0; c < 6; ++c) carry[c] = p.seg_carry[(run * p.nseg + seg) * 6 + c];
  }
  __syncthreads();

  for (int64_t tile0 = seg_lo; tile0 < seg_hi; tile0 += kNoiseThreads) {
    const int64_t t = tile0 + i;
    const bool live = t < seg_hi;
    double m[6], z[6];
    if (live && p.pass == 1) {   // only the Gauss-Markov drives matter
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        z[c] = normal_pair(static_cast<uint32_t>(t), kDrawAccel + c, run_lo, run_hi, p.k0, p.k1).z0;
        z[3 + c] = normal_pair(static_cast<uint32_t>(t), kDrawGyro + c, run_lo, run_hi, p.k0, p.k1).z0;
        m[c] = m[3 + c] = 0.0;
      }
    } else if (live) {
      noisy_sample(p, p.ref_accel + t * 3, p.ref_gyro + t * 3, static_cast<uint32_t>(t), run_lo,
                   run_hi, run, phase, m, m + 3, z, z + 3);
    } else {
#pragma unroll
      for (int c = 0; c < 6; ++c) m[c] = z[c] = 0.0;
    }
    // The six Gauss-Markov recurrences d[t+1] = a d[t] + b z[t] as affine block scans, together:
    // warp-level inclusive scans (shuffles with powers of a), the warp totals combined by 48
    // threads, two barriers per tile.  y_i = sum_{q<=i} a^{i-q} b z_q (zero state at the tile
    // start), d[tile0 + i] = a^i carry + y_{i-1}.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double y[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const TriadNoise& e = (c < 3) ? p.accel : p.gyro;
      y[c] = e.gm_b[c % 3] * z[c];
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
#pragma unroll
      for (int c = 0; c < 6; ++c) {
        const double u = __shfl_up_sync(0xffffffffu, y[c], off);
        if (lane >= off) y[c] += apow[c][off] * u;
      }
    }
    if (lane == 31) {
#pragma unroll
      for (int c = 0; c < 6; ++c) sh_w[c][warp] = y[c];
    }
    __syncthreads();
    if (threadIdx.x < 6 * (kNoiseWarps + 1)) {
      // thread (c, w): value at the end of warp w-1 from a zero state at the tile start
      const int c = threadIdx.x / (kNoiseWarps + 1), w = threadIdx.x % (kNoiseWarps + 1);
      double pre = 0.0;
      for (int q = 0; q < w; ++q) pre = apow[c][32] * pre + sh_w[c][q];
      sh_pre[c][w] = pre;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 6; ++c) {
      const TriadNoise& e = (c < 3) ? p.accel : p.gyro;
      const int a = c % 3;
      const double pre = sh_pre[c][warp];
      const double yi = y[c] + apow[c][lane + 1] * pre;      // inclusive, whole tile
      const double up = __shfl_up_sync(0xffffffffu, yi, 1);
      const double ym1 = (lane == 0) ? pre : up;             // y_{i-1}; 0 for the first sample
      const double d = apow[c][i] * carry[c] + ym1;
      m[c] += d + e.wd[a] * z[c];
      carry[c] = apow[c][kNoiseThreads] * carry[c] + sh_pre[c][kNoiseWarps];
    }
    if (live && p.pass == 0) {
      const int64_t o = run * p.osr + t * p.ost;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        p.out_accel[o + c * p.osc] = m[c];
        p.out_gyro[o + c * p.osc] = m[3 + c];
      }
      if (p.z_dump) {
        // (acc_gm[3], acc_w[3], gyr_gm[3], gyr_w[3]); the white normals are recovered from
        // the same Philox draws
        double* zd = p.z_dump + (run * p.n + t) * 12;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const Normal2 za = normal_pair(static_cast<uint32_t>(t), kDrawAccel + c, run_lo, run_hi, p.k0, p.k1);
          const Normal2 zg = normal_pair(static_cast<uint32_t>(t), kDrawGyro + c, run_lo, run_hi, p.k0, p.k1);
          zd[c] = za.z0;
          zd[3 + c] = za.z1;
          zd[6 + c] = zg.z0;
          zd[9 + c] = zg.z1;
        }
      }
    }
  }
  if (p.pass == 1 && threadIdx.x == 0) {
    // the tiles of a segment are whole (seg_len is a multiple of the tile) except in the last
    // segment, whose end value is never used
#pragma unroll
    for (int c = 0; c < 6; ++c) p.seg_end[(run * p.nseg + seg) * 6 + c] = carry[c];
  }
}

// carry-in of every segment from the zero-state segment responses: c[s+1] = a^L c[s] + E[s]
__global__ void noise_carry_kernel(NoiseParams p) {
  const int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= p.runs * 6) return;
  const int64_t run = idx / 6;
  const int c = static_cast<int>(idx % 6);
  const double a = (c < 3) ? p.accel.gm_a[c] : p.gyro.gm_a[c - 3];
  const double aL = pow(a, static_cast<double>(p.seg_len));
  double cin = 0.0;
  for (int s = 0; s < p.nseg; ++s) {
    p.seg_carry[(run * p.nseg + s) * 6 + c] = cin;
    if (s + 1 < p.nseg) cin = aL * cin + p.seg_end[(run * p.nseg + s) * 6 + c];
  }
}

}  // namespace b2ins

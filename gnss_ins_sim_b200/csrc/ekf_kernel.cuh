// K7: Monte-Carlo loosely-coupled GNSS/INS filter (BASELINE config 5).
//
// The reference's demo_algorithms/ins_loose.py is a stub (prediction() / correction() are `pass`,
// ins_loose.py:124-134): there is nothing to be parity-checked against.  The filter implemented here
// is specified from first principles in DESIGN.md section 11 (read that first: state order, error
// convention, Phi, Q, the six scalar GPS updates, the closed-loop correction); this kernel restates
// it for the device and is held to that spec on identical Philox draws.  What comes from the reference:
// the sensor models that feed it (pathgen.acc_gen / gyro_gen :441-594, gps_gen :596-625, the same
// Philox streams as K12 / K6) and the strapdown step (free_integration.py:133-172 = nav_step<0>).
//
// One thread owns one Monte-Carlo run for the whole series (the filter is serial in time and its
// GPS epochs are common to all runs, so a warp never diverges): the nominal state, the bias
// estimates and the generator's Gauss-Markov states live in registers, the 15 x 15 covariance in
// shared memory as P[element][thread] (bank-conflict-free; 57.6 KB per 32-thread CTA, three CTAs per
// SM).  Per IMU sample: six Box-Muller pairs, one strapdown step, and P <- Phi P Phi^T + Q done as two
// in-place sweeps that use the block structure of Phi (about 1100 FMA instead of 6750).  Per GPS
// sample (every fs / fs_gps steps): three more pairs, six scalar updates, the correction.
#pragma once
#include "mc_kernel.cuh"
#include "gps_kernel.cuh"

namespace b2ins {

constexpr int kEkfThreads = 32;
constexpr int kEkfN = 15;
constexpr uint32_t kDrawIni = 27;   // + j, t = 0xFFFFFFFE: the initial-state errors

struct EkfParams {
  int64_t n, runs, run_offset, m;
  double dt;
  int earth_rot;
  uint32_t k0, k1;
  TriadNoise gyro, accel;          // the generator (pre-digested as for K12)
  const double* ref_gyro;          // [n][3]
  const double* ref_accel;         // [n][3]
  const double* ref_nav;           // [n][9] att, pos (LLA), vel (NED)
  const double* ref_gps;           // [m][6]
  const int64_t* gps_idx;          // [m] IMU sample index of every GPS row (ascending)
  const double* gps_vis;           // [m] 1 = visible
  double stdp[3], stdv[3];         // GPS noise of the generator [m], [m/s] = the filter's R
  double p0[15];
  double ini[9];                   // true initial LLA, body velocity, Euler angles
  double ag[3], qg[3], aa[3], qa[3];   // bias model: a and b^2 per axis
  double arw2dt[3], vrw2dt[3];
  double qv_extra, qphi_extra;     // model-mismatch random walks: vel_rw^2 dt, att_rw^2 dt
  int64_t stats_start;
  double* end_err;                 // [runs][9]
  double* end_bias;                // [runs][6]
  double* consist;                 // [runs][19]: NEES sums (pos, vel, att), inside-3-sigma counts [15], epochs
  double* out_att;                 // histories of runs [0, dump_runs): [dump_runs][rows][3]
  double* out_pos;
  double* out_vel;
  double* out_wb;
  double* out_ab;
  int64_t dump_runs, dump_stride, dump_rows;
};

// 3 x 3 symmetric-positive NEES  e^T A^-1 e  via the adjugate
__device__ __forceinline__ double nees3(const double* a /* row-major 3x3 */, const double* e) {
  const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
  const double c11 = a[0] * a[8] - a[2] * a[6], c12 = a[1] * a[6] - a[0] * a[7];
  const double c22 = a[0] * a[4] - a[1] * a[3];
  const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
  const double q = e[0] * (c00 * e[0] + c01 * e[1] + c02 * e[2]) + e[1] * (c01 * e[0] + c11 * e[1] + c12 * e[2]) +
                   e[2] * (c02 * e[0] + c12 * e[1] + c22 * e[2]);
  return q / det;
}

// n -> b DCM (attitude.euler2dcm 'zyx' layout) -> [yaw, pitch, roll]
__device__ __forceinline__ void dcm2euler(const Dcm& c, double* yaw, double* pitch, double* roll) {
  *yaw = atan2(c.c01, c.c00);
  *pitch = -asin(fmin(1.0, fmax(-1.0, c.c02)));
  *roll = atan2(c.c12, c.c22);
}

// c (n -> b) times (I + s [phi x]), s = +-1
__device__ __forceinline__ Dcm dcm_times_small(const Dcm& c, const double* phi, double s) {
  const double x = s * phi[0], y = s * phi[1], z = s * phi[2];
  // (I + [phi x]) = [[1, -z, y], [z, 1, -x], [-y, x, 1]]
  Dcm r;
  r.c00 = c.c00 + c.c01 * z - c.c02 * y;
  r.c01 = -c.c00 * z + c.c01 + c.c02 * x;
  r.c02 = c.c00 * y - c.c01 * x + c.c02;
  r.c10 = c.c10 + c.c11 * z - c.c12 * y;
  r.c11 = -c.c10 * z + c.c11 + c.c12 * x;
  r.c12 = c.c10 * y - c.c11 * x + c.c12;
  r.c20 = c.c20 + c.c21 * z - c.c22 * y;
  r.c21 = -c.c20 * z + c.c21 + c.c22 * x;
  r.c22 = c.c20 * y - c.c21 * x + c.c22;
  return r;
}

__device__ __forceinline__ void set_attitude(NavState& st, double yaw, double pitch, double roll, double dt) {
  st.yaw = yaw;
  st.pitch = pitch;
  st.roll = roll;
  resync_exact<0>(st);
  st.icp = rcp_nr(st.sc.cp) * dt;
}

__global__ void __launch_bounds__(kEkfThreads) ekf_kernel(const __grid_constant__ EkfParams p) {
  extern __shared__ double Psm[];           // [225][kEkfThreads]
  const int tid = threadIdx.x;
  const int64_t run_raw = static_cast<int64_t>(blockIdx.x) * kEkfThreads + tid;
  const bool active = run_raw < p.runs;
  const int64_t run = active ? run_raw : p.runs - 1;
  const int64_t grun = p.run_offset + run;
  const uint32_t run_lo = static_cast<uint32_t>(grun), run_hi = static_cast<uint32_t>(grun >> 32);
  const bool dump = active && run < p.dump_runs && p.out_att;
  const double dt = p.dt;
  auto P = [&](int i, int j) -> double& { return Psm[(i * kEkfN + j) * kEkfThreads + tid]; };

  // GPS noise of the generator: horizontal sigmas in radians with the radii at the FIRST reference
  // sample, as pathgen.gps_gen does (pathgen.py:617-620)
  double sdp0 = p.stdp[0], sdp1 = p.stdp[1];
  {
    const GeoParam gp = geo_param(p.ref_gps[0], p.ref_gps[2]);
    sdp0 = div_nr(sdp0, gp.rm);
    sdp1 = div_nr(div_nr(sdp1, gp.rn), gp.cl);
  }

  // ---- initial covariance and nominal state: truth + a draw from P0 -----------------------------
#pragma unroll 1
  for (int e = 0; e < kEkfN * kEkfN; ++e) Psm[e * kEkfThreads + tid] = 0.0;
#pragma unroll
  for (int i = 0; i < kEkfN; ++i) P(i, i) = p.p0[i];
  double e0[10];
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    const Normal2 z = normal_pair(0xFFFFFFFEu, kDrawIni + j, run_lo, run_hi, p.k0, p.k1);
    e0[2 * j] = z.z0;
    e0[2 * j + 1] = z.z1;
  }
#pragma unroll
  for (int i = 0; i < 9; ++i) e0[i] *= sqrt(p.p0[i]);
  NavState st;
  {
    double ini[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) ini[i] = p.ini[i];
    nav_init<0>(st, ini, 9, dt);          // the TRUE initial state: st.vel = C^T v_body
    const GeoParam gp = geo_param(ini[0], ini[2]);
    st.pos.x += e0[0] / (gp.rm + ini[2]);
    st.pos.y += e0[1] / ((gp.rn + ini[2]) * gp.cl);
    st.pos.z -= e0[2];
    st.vel.x += e0[3];
    st.vel.y += e0[4];
    st.vel.z += e0[5];
    // C_hat(n->b) = C(n->b) (I + [phi x])
    const Dcm c = dcm_times_small(dcm_from_sincos(st.sc), e0 + 6, 1.0);
    double y, pt, r;
    dcm2euler(c, &y, &pt, &r);
    set_attitude(st, y, pt, r, dt);
  }
  double bg[3] = {0.0, 0.0, 0.0}, ba[3] = {0.0, 0.0, 0.0};       // bias estimates
  double carry[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};              // generator: GM drift d[i] (accel, gyro)
  double phase[3] = {0.0, 0.0, 0.0};
  double nees[3] = {0.0, 0.0, 0.0};
  int inside[kEkfN];
#pragma unroll
  for (int i = 0; i < kEkfN; ++i) inside[i] = 0;
  int epochs = 0;
  int64_t jg = 0;                          // next GPS row
  int64_t next_gps = p.m > 0 ? p.gps_idx[0] : -1;

  for (int64_t i = 0; i < p.n; ++i) {
    // ================= GPS sample of IMU sample i: update, then the consistency record ==========
    if (i == next_gps) {
      if (p.gps_vis[jg] > 0.0) {
        double zn[6];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const Normal2 zz = normal_pair(static_cast<uint32_t>(jg), kPairGps + j, run_lo, run_hi, p.k0, p.k1);
          zn[2 * j] = zz.z0;
          zn[2 * j + 1] = zz.z1;
        }
        const double* rg = p.ref_gps + jg * 6;
        const GeoParam gp = geo_param_sc(st.sl, st.cl, st.pos.z);
        const double rmh = gp.rm + st.pos.z, rnh = (gp.rn + st.pos.z) * gp.cl;
        double zm[6];
        zm[0] = (st.pos.x - (rg[0] + sdp0 * zn[0])) * rmh;
        zm[1] = (st.pos.y - (rg[1] + sdp1 * zn[1])) * rnh;
        zm[2] = -(st.pos.z - (rg[2] + p.stdp[2] * zn[2]));
        zm[3] = st.vel.x - (rg[3] + p.stdv[0] * zn[3]);
        zm[4] = st.vel.y - (rg[4] + p.stdv[1] * zn[4]);
        zm[5] = st.vel.z - (rg[5] + p.stdv[2] * zn[5]);
        double x[kEkfN];
#pragma unroll
        for (int q = 0; q < kEkfN; ++q) x[q] = 0.0;
#pragma unroll 1
        for (int k = 0; k < 6; ++k) {
          const double rk = (k < 3) ? p.stdp[k] * p.stdp[k] : p.stdv[k - 3] * p.stdv[k - 3];
          const double inv_s = 1.0 / (P(k, k) + rk);
          double K[kEkfN], row[kEkfN];
#pragma unroll
          for (int q = 0; q < kEkfN; ++q) {
            K[q] = P(q, k) * inv_s;
            row[q] = P(k, q);
          }
          double xk = 0.0, zk = 0.0;   // x[k], zm[k] with a run-time index: selects, not local-memory arrays
#pragma unroll
          for (int q = 0; q < 6; ++q) {
            xk = (q == k) ? x[q] : xk;
            zk = (q == k) ? zm[q] : zk;
          }
          const double innov = zk - xk;
#pragma unroll
          for (int q = 0; q < kEkfN; ++q) x[q] += K[q] * innov;
          // P <- P - K P[k,:], then (P + P^T)/2 as the spec does
#pragma unroll 1
          for (int a = 0; a < kEkfN; ++a) {
            double Ka = 0.0;
#pragma unroll
            for (int q = 0; q < kEkfN; ++q) Ka = (q == a) ? K[q] : Ka;
#pragma unroll
            for (int b = 0; b < kEkfN; ++b) P(a, b) -= Ka * row[b];
          }
#pragma unroll 1
          for (int a = 0; a < kEkfN; ++a)
            for (int b = a + 1; b < kEkfN; ++b) {
              const double s = 0.5 * (P(a, b) + P(b, a));
              P(a, b) = s;
              P(b, a) = s;
            }
        }
        // ---- close the loop -----------------------------------------------------------------
        st.pos.x -= x[0] / rmh;
        st.pos.y -= x[1] / rnh;
        st.pos.z += x[2];
        st.vel.x -= x[3];
        st.vel.y -= x[4];
        st.vel.z -= x[5];
        // C(n->b) = C_hat(n->b) (I - [phi x])
        const Dcm c = dcm_times_small(dcm_from_sincos(st.sc), x + 6, -1.0);
        double y, pt, r;
        dcm2euler(c, &y, &pt, &r);
        set_attitude(st, y, pt, r, dt);     // also refreshes sin/cos of the corrected latitude
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) {
          bg[c3] -= x[9 + c3];
          ba[c3] -= x[12 + c3];
        }
      }
      if (i >= p.stats_start) {
        const double* rn9 = p.ref_nav + i * 9;
        const GeoParam gp = geo_param(rn9[3], rn9[5]);
        double e[kEkfN];
        e[0] = (st.pos.x - rn9[3]) * (gp.rm + rn9[5]);
        e[1] = (st.pos.y - rn9[4]) * (gp.rn + rn9[5]) * gp.cl;
        e[2] = -(st.pos.z - rn9[5]);
        e[3] = st.vel.x - rn9[6];
        e[4] = st.vel.y - rn9[7];
        e[5] = st.vel.z - rn9[8];
        const Dcm ct = dcm_from_sincos(sincos3(rn9[0], rn9[1], rn9[2]));    // true n -> b
        const Dcm ce = dcm_from_sincos(st.sc);                              // estimated n -> b
        // M = C_hat(b->n) C(n->b) = ce^T ct = I - [phi x]
        const double m21 = ce.c02 * ct.c01 + ce.c12 * ct.c11 + ce.c22 * ct.c21;
        const double m12 = ce.c01 * ct.c02 + ce.c11 * ct.c12 + ce.c21 * ct.c22;
        const double m02 = ce.c00 * ct.c02 + ce.c10 * ct.c12 + ce.c20 * ct.c22;
        const double m20 = ce.c02 * ct.c00 + ce.c12 * ct.c10 + ce.c22 * ct.c20;
        const double m10 = ce.c01 * ct.c00 + ce.c11 * ct.c10 + ce.c21 * ct.c20;
        const double m01 = ce.c00 * ct.c01 + ce.c10 * ct.c11 + ce.c20 * ct.c21;
        e[6] = -0.5 * (m21 - m12);
        e[7] = -0.5 * (m02 - m20);
        e[8] = -0.5 * (m10 - m01);
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) {
          e[9 + c3] = bg[c3] - (p.gyro.b[c3] + carry[3 + c3]);
          e[12 + c3] = ba[c3] - (p.accel.b[c3] + carry[c3]);
        }
#pragma unroll
        for (int b = 0; b < 3; ++b) {
          double a9[9];
#pragma unroll
          for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c3 = 0; c3 < 3; ++c3) a9[r * 3 + c3] = P(3 * b + r, 3 * b + c3);
          nees[b] += nees3(a9, e + 3 * b);
        }
#pragma unroll
        for (int q = 0; q < kEkfN; ++q) inside[q] += (fabs(e[q]) <= 3.0 * sqrt(P(q, q))) ? 1 : 0;
        ++epochs;
      }
      ++jg;
      next_gps = jg < p.m ? p.gps_idx[jg] : -1;
    }
    // ================= histories ==================================================================
    int64_t row;
    if (dump && dump_row_generic(p.dump_stride, i, &row)) {
      const int64_t o = (run * p.dump_rows + row) * 3;
      p.out_att[o] = wrap_once(st.yaw); p.out_att[o + 1] = st.pitch; p.out_att[o + 2] = wrap_once(st.roll);
      p.out_pos[o] = st.pos.x; p.out_pos[o + 1] = st.pos.y; p.out_pos[o + 2] = st.pos.z;
      p.out_vel[o] = st.vel.x; p.out_vel[o + 1] = st.vel.y; p.out_vel[o + 2] = st.vel.z;
#pragma unroll
      for (int c3 = 0; c3 < 3; ++c3) {
        p.out_wb[o + c3] = bg[c3];
        p.out_ab[o + c3] = ba[c3];
      }
    }
    if (i == p.n - 1) break;
    // ================= the measurements of sample i (the K12 generator) ==========================
    double ma[3], mg[3], za[3], zg[3];
    noisy_sample(p, p.ref_accel + i * 3, p.ref_gyro + i * 3, static_cast<uint32_t>(i), run_lo, run_hi, run, phase,
                 ma, mg, za, zg);
#pragma unroll
    for (int c3 = 0; c3 < 3; ++c3) {
      ma[c3] += carry[c3] + p.accel.wd[c3] * za[c3];
      mg[c3] += carry[3 + c3] + p.gyro.wd[c3] * zg[c3];
      carry[c3] = fma(p.accel.gm_a[c3], carry[c3], p.accel.gm_b[c3] * za[c3]);
      carry[3 + c3] = fma(p.gyro.gm_a[c3], carry[3 + c3], p.gyro.gm_b[c3] * zg[c3]);
    }
    const Vec3 w{mg[0] - bg[0], mg[1] - bg[1], mg[2] - bg[2]};
    const Vec3 f{ma[0] - ba[0], ma[1] - ba[1], ma[2] - ba[2]};
    // ================= covariance: P <- Phi P Phi^T + Q with the blocks of Phi ====================
    {
      const Dcm c = dcm_from_sincos(st.sc);          // n -> b of sample i; b -> n is its transpose
      const double cb[9] = {c.c00 * dt, c.c10 * dt, c.c20 * dt, c.c01 * dt, c.c11 * dt, c.c21 * dt,
                            c.c02 * dt, c.c12 * dt, c.c22 * dt};      // C(b->n) dt, row-major
      const Vec3 fn = rot_b2n(st.sc, f);
      const double sx = fn.x * dt, sy = fn.y * dt, sz = fn.z * dt;    // [f_n x] dt = [[0,-sz,sy],[sz,0,-sx],[-sy,sx,0]]
      // sweep 1, column by column: A = Phi P
#pragma unroll 1
      for (int j = 0; j < kEkfN; ++j) {
        double col[kEkfN];
#pragma unroll
        for (int q = 0; q < kEkfN; ++q) col[q] = P(q, j);
        P(0, j) = fma(dt, col[3], col[0]);
        P(1, j) = fma(dt, col[4], col[1]);
        P(2, j) = fma(dt, col[5], col[2]);
        P(3, j) = col[3] + (-sz * col[7] + sy * col[8]) - (cb[0] * col[12] + cb[1] * col[13] + cb[2] * col[14]);
        P(4, j) = col[4] + (sz * col[6] - sx * col[8]) - (cb[3] * col[12] + cb[4] * col[13] + cb[5] * col[14]);
        P(5, j) = col[5] + (-sy * col[6] + sx * col[7]) - (cb[6] * col[12] + cb[7] * col[13] + cb[8] * col[14]);
        P(6, j) = col[6] + (cb[0] * col[9] + cb[1] * col[10] + cb[2] * col[11]);
        P(7, j) = col[7] + (cb[3] * col[9] + cb[4] * col[10] + cb[5] * col[11]);
        P(8, j) = col[8] + (cb[6] * col[9] + cb[7] * col[10] + cb[8] * col[11]);
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) {
          P(9 + c3, j) = p.ag[c3] * col[9 + c3];
          P(12 + c3, j) = p.aa[c3] * col[12 + c3];
        }
      }
      // sweep 2, row by row: P = A Phi^T
#pragma unroll 1
      for (int r = 0; r < kEkfN; ++r) {
        double a[kEkfN];
#pragma unroll
        for (int q = 0; q < kEkfN; ++q) a[q] = P(r, q);
        P(r, 0) = fma(dt, a[3], a[0]);
        P(r, 1) = fma(dt, a[4], a[1]);
        P(r, 2) = fma(dt, a[5], a[2]);
        P(r, 3) = a[3] + (-sz * a[7] + sy * a[8]) - (cb[0] * a[12] + cb[1] * a[13] + cb[2] * a[14]);
        P(r, 4) = a[4] + (sz * a[6] - sx * a[8]) - (cb[3] * a[12] + cb[4] * a[13] + cb[5] * a[14]);
        P(r, 5) = a[5] + (-sy * a[6] + sx * a[7]) - (cb[6] * a[12] + cb[7] * a[13] + cb[8] * a[14]);
        P(r, 6) = a[6] + (cb[0] * a[9] + cb[1] * a[10] + cb[2] * a[11]);
        P(r, 7) = a[7] + (cb[3] * a[9] + cb[4] * a[10] + cb[5] * a[11]);
        P(r, 8) = a[8] + (cb[6] * a[9] + cb[7] * a[10] + cb[8] * a[11]);
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) {
          P(r, 9 + c3) = p.ag[c3] * a[9 + c3];
          P(r, 12 + c3) = p.aa[c3] * a[12 + c3];
        }
      }
      // Q: C diag(vrw^2 dt) C^T and C diag(arw^2 dt) C^T (C = b -> n, here cb / dt), bias drives
      const double inv_dt2 = 1.0 / (dt * dt);
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c3 = 0; c3 < 3; ++c3) {
          const double qv = (cb[r * 3] * p.vrw2dt[0] * cb[c3 * 3] + cb[r * 3 + 1] * p.vrw2dt[1] * cb[c3 * 3 + 1] +
                             cb[r * 3 + 2] * p.vrw2dt[2] * cb[c3 * 3 + 2]) * inv_dt2;
          const double qa = (cb[r * 3] * p.arw2dt[0] * cb[c3 * 3] + cb[r * 3 + 1] * p.arw2dt[1] * cb[c3 * 3 + 1] +
                             cb[r * 3 + 2] * p.arw2dt[2] * cb[c3 * 3 + 2]) * inv_dt2;
          P(3 + r, 3 + c3) += qv;
          P(6 + r, 6 + c3) += qa;
        }
#pragma unroll
      for (int c3 = 0; c3 < 3; ++c3) {
        P(3 + c3, 3 + c3) += p.qv_extra;
        P(6 + c3, 6 + c3) += p.qphi_extra;
        P(9 + c3, 9 + c3) += p.qg[c3];
        P(12 + c3, 12 + c3) += p.qa[c3];
      }
    }
    // ================= nominal state ==========================================================
    const bool resync = ((i + 1) & (kResync - 1)) == 0;
    nav_step<0, false, 0>(st, w, f, dt, p.earth_rot != 0, 0, resync);
#pragma unroll
    for (int c3 = 0; c3 < 3; ++c3) {
      bg[c3] *= p.ag[c3];
      ba[c3] *= p.aa[c3];
    }
  }

  if (active) {
    const double* r = p.ref_nav + (p.n - 1) * 9;
    double* e = p.end_err + run * 9;
    e[0] = angle_range_pi(st.yaw - r[0]);
    e[1] = angle_range_pi(st.pitch - r[1]);
    e[2] = angle_range_pi(st.roll - r[2]);
    e[3] = st.pos.x - r[3];
    e[4] = st.pos.y - r[4];
    e[5] = st.pos.z - r[5];
    e[6] = st.vel.x - r[6];
    e[7] = st.vel.y - r[7];
    e[8] = st.vel.z - r[8];
    if (p.end_bias) {
#pragma unroll
      for (int c3 = 0; c3 < 3; ++c3) {
        p.end_bias[run * 6 + c3] = bg[c3];
        p.end_bias[run * 6 + 3 + c3] = ba[c3];
      }
    }
    if (p.consist) {
      double* o = p.consist + run * 19;
      o[0] = nees[0]; o[1] = nees[1]; o[2] = nees[2];
#pragma unroll
      for (int q = 0; q < kEkfN; ++q) o[3 + q] = static_cast<double>(inside[q]);
      o[18] = static_cast<double>(epochs);
    }
  }
}

}  // namespace b2ins

// K2 / K12: the Monte-Carlo strapdown kernel.
//
// One LANE GROUP of G lanes (G = 1,2,4,...,32) owns one Monte-Carlo run; G = 32 is
// "one warp owns one run".  Per block of G consecutive samples:
//   phase A (time-parallel): lane j prepares sample base+j -- Philox4x32-10 + Box-Muller
//           normals, white noise, constant bias, vibration added to the true IMU sample
//           (read from the TMA-staged shared-memory tile), or the fed gyro/accel sample.
//   phase B (serial in time): for k = 0..G-1 every lane of the group pulls sample base+k
//           from lane k by warp shuffle, advances the Gauss-Markov bias and the 9-DoF
//           strapdown state (registers, replicated across the group) by one step.
// The shared true trajectory (ref gyro/accel [n][3], optionally ref nav [n][9]) is staged
// through shared memory in tiles of kTile samples by 1-D bulk async copies (TMA) completing
// on mbarriers, kStages deep, one pipeline per CTA shared by all its runs.
#pragma once
#include "mech.cuh"

namespace b2ins {

// Optional phase clocks (tools only: -DB2INS_PHASE_CLOCKS builds libb2ins_prof.so)
#ifdef B2INS_PHASE_CLOCKS
__device__ unsigned long long g_phase_clocks[8];
#define B2_CLK(var) const long long var = clock64()
#define B2_ACC(i, t0, t1) \
  if ((threadIdx.x & 31) == 0) atomicAdd(&g_phase_clocks[i], static_cast<unsigned long long>((t1) - (t0)))
#else
#define B2_CLK(var)
#define B2_ACC(i, t0, t1)
#endif

constexpr int kWarps = 4;
constexpr int kThreads = kWarps * 32;
constexpr int kTile = 128;   // samples per shared-memory tile (multiple of 32)
constexpr int kStagesFast = 3;   // pipeline depth; 2 when the 72 B/sample nav tile is staged too
template <bool PROC>
struct Stages {
  static constexpr int value = PROC ? 2 : kStagesFast;
};

// error model of one triad, pre-digested on the host (bias_drift: pathgen.py:583-586)
struct TriadNoise {
  double b[3];        // constant bias
  double gm_a[3];     // 1 - dt/tau            (0 if the drift is white)
  double gm_b[3];     // drift*sqrt(1-exp(-2dt/tau))  (0 if the drift is white)
  double wd[3];       // drift sigma if the drift is white (corr = inf), else 0
  double w[3];        // rw / sqrt(dt)
  int vib_type;
  int series_len;
  double vib_amp[3];
  double vib_w;       // ((2 pi) f) dt
  const double* series;  // [runs][3][series_len]
};

struct McParams {
  int64_t n, runs, run_offset, ini_offset;
  double dt;
  int earth_rot;
  uint32_t k0, k1;
  TriadNoise gyro, accel;
  const double* ref_gyro;   // [n][3]
  const double* ref_accel;  // [n][3]
  const double* ref_nav;    // [n][9] att, pos, vel
  const double* ini;        // [ini_sets][ini_rows]
  int ini_sets, ini_rows;
  // fed measurements (K2) -- element (r,t,c) at r*sr + t*st + c*sc
  const double* fed_gyro;
  const double* fed_accel;
  int64_t sr, st, sc;
  // odometer variant (free_integration_odo): algo = 1
  int algo;
  const double* ref_odo;   // [n] true forward speed (pathgen 'odo')
  double odo_scale, odo_stdv;
  const double* fed_odo;   // K2: element (r,t) at r*so_r + t*so_t
  int64_t so_r, so_t;
  // histories for runs [0, dump_runs): same stride convention
  double* out_att;
  double* out_pos;
  double* out_vel;
  double* out_gyro;
  double* out_accel;
  double* out_odo;     // [dump_runs][n] (algo 1)
  double* out_quat;    // [dump_runs][rows][4] scalar-first quaternion of every kept attitude sample
  int64_t osr, ost, osc;
  int64_t dump_runs;
  int64_t dump_stride; // >= 1: histories keep samples 0, s, 2s, ... (rows = ceil(n / s))
  int64_t dump_rows;
  // per-run results
  double* end_err;     // [runs][9]
  double* end_state;   // [runs][9]
  double* proc_stats;  // [runs][3][9]
  int64_t stats_start;
  int debug;           // tools (B2INS_PHASE_CLOCKS builds only): 1 = producers idle, 2 = integrators idle
};

// Prepared samples of one block, one slot per lane: phase A stores (gyro xyz, accel xyz),
// phase B reads sample k of its group with three 128-bit broadcast loads instead of twelve
// 32-bit shuffles.  48 B per lane; row padding keeps the group bases on distinct banks.
struct alignas(16) SampleSlot {
  double g[3], a[3];
};

template <bool PROC>
struct TileSmem {
  static constexpr int kStages = Stages<PROC>::value;
  alignas(128) double gyro[kStages][kTile * 3];
  alignas(128) double accel[kStages][kTile * 3];
  alignas(128) double nav[kStages][PROC ? kTile * 9 : 2];  // only staged for process statistics
  alignas(16) SampleSlot slot[kWarps][32];
  alignas(8) uint64_t full[kStages];
  alignas(8) uint64_t empty[kStages];
};

// Issue the copies of one tile: the 16-byte-multiple part by bulk async copy (TMA)
// completing on full[s]; an odd sample count leaves one 8-byte tail copied by hand.
template <bool FED, bool PROC>
__device__ __forceinline__ void issue_tile(TileSmem<PROC>& sm, const McParams& p, int64_t tile, int s) {
  const int64_t t0 = tile * kTile;
  const uint32_t cnt = static_cast<uint32_t>(min64(kTile, p.n - t0));
  uint32_t tx = 0;
  if (!FED) tx += 2u * ((cnt * 24u) & ~15u);
  if (PROC) tx += (cnt * 72u) & ~15u;
  // the hand-copied tails are ordered before the arrive (release) below
  if (!FED) {
    if ((cnt * 24u) & 8u) {
      const uint32_t o = ((cnt * 24u) & ~15u) / 8;
      sm.gyro[s][o] = p.ref_gyro[t0 * 3 + o];
      sm.accel[s][o] = p.ref_accel[t0 * 3 + o];
    }
  }
  if (PROC) {
    if ((cnt * 72u) & 8u) {
      const uint32_t o = ((cnt * 72u) & ~15u) / 8;
      sm.nav[s][o] = p.ref_nav[t0 * 9 + o];
    }
  }
  mbar_arrive_expect_tx(&sm.full[s], tx);
  if (!FED) {
    const uint32_t b = (cnt * 24u) & ~15u;
    if (b) {
      bulk_g2s(sm.gyro[s], p.ref_gyro + t0 * 3, b, &sm.full[s]);
      bulk_g2s(sm.accel[s], p.ref_accel + t0 * 3, b, &sm.full[s]);
    }
  }
  if (PROC) {
    const uint32_t b = (cnt * 72u) & ~15u;
    if (b) bulk_g2s(sm.nav[s], p.ref_nav + t0 * 9, b, &sm.full[s]);
  }
}

// Vibration term of one axis (pathgen.py:477-493 / :540-555); only called when a model is on.
__device__ __forceinline__ double vib_term(const TriadNoise& e, int c, int sensor, uint32_t t,
                                           uint32_t run_lo, uint32_t run_hi, uint32_t k0,
                                           uint32_t k1, int64_t run_local, const double* phase) {
  if (e.vib_type == 1) {
    const Normal2 zv = normal_pair(t, kDrawVib + c, run_lo, run_hi, k0, k1);
    return e.vib_amp[c] * (sensor == 0 ? zv.z0 : zv.z1);
  }
  if (e.vib_type == 2) {
    const double arg = e.vib_w * static_cast<double>(t) + (sensor == 0 ? 0.0 : phase[c]);
    return e.vib_amp[c] * sin(arg);
  }
  if (e.vib_type == 3)
    return e.series[(run_local * 3 + c) * e.series_len + (t % static_cast<uint32_t>(e.series_len))];
  return 0.0;
}

// phase A in Monte-Carlo mode for one sample: the six Box-Muller pairs are drawn in ONE
// straight-line block (six independent Philox -> log/sqrt/sincospi chains the scheduler can
// interleave), then (ref + b) + white [+ vib] per axis, and the GM drive normals.
template <class P>
__device__ __forceinline__ void noisy_sample(const P& p, const double* ref_a3,
                                             const double* ref_g3, uint32_t t, uint32_t run_lo,
                                             uint32_t run_hi, int64_t run_local,
                                             const double* phase, double* ma, double* mg,
                                             double* za, double* zg) {
  Normal2 z[6];
#pragma unroll
  for (int c = 0; c < 6; ++c) z[c] = normal_pair(t, c, run_lo, run_hi, p.k0, p.k1);  // draws 0..5
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    za[c] = z[c].z0;
    zg[c] = z[3 + c].z0;
    ma[c] = (ref_a3[c] + p.accel.b[c]) + p.accel.w[c] * z[c].z1;
    mg[c] = (ref_g3[c] + p.gyro.b[c]) + p.gyro.w[c] * z[3 + c].z1;
  }
  if (p.accel.vib_type | p.gyro.vib_type) {   // uniform branch, off in the BASELINE configs
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      ma[c] += vib_term(p.accel, c, 0, t, run_lo, run_hi, p.k0, p.k1, run_local, phase);
      mg[c] += vib_term(p.gyro, c, 1, t, run_lo, run_hi, p.k0, p.k1, run_local, phase);
    }
  }
}

// Gauss-Markov drift of the G samples of a block, time-parallel: d[t+1] = a d[t] + b z[t] is an
// affine recurrence, so the group runs an inclusive scan y_j = sum_{q<=j} a^(j-q) b z_q with
// warp shuffles; d_j = a^j carry + y_(j-1) and the carry moves on by a^G carry + y_(G-1).
template <int G>
__device__ __forceinline__ double gm_block(double x, double a, double apj, double aG, int j,
                                           double& carry) {
  double y = x;
  if (G > 1) {
    double ap = a;
#pragma unroll
    for (int off = 1; off < G; off <<= 1) {
      const double u = __shfl_up_sync(0xffffffffu, y, off, G);
      if (j >= off) y = fma(ap, u, y);
      ap *= ap;
    }
  }
  double ym1 = 0.0, ylast = y;
  if (G > 1) {
    ym1 = __shfl_up_sync(0xffffffffu, y, 1, G);
    if (j == 0) ym1 = 0.0;
    ylast = __shfl_sync(0xffffffffu, y, G - 1, G);
  }
  const double d = (G > 1) ? fma(apj, carry, ym1) : carry;
  carry = fma(aG, carry, ylast);
  return d;
}

// Row of sample t in the (possibly decimated) histories; false: the sample is not kept.
__device__ __forceinline__ bool dump_row_generic(int64_t stride, int64_t t, int64_t* row) {
  if (stride <= 1) {
    *row = t;
    return true;
  }
  const int64_t q = t / stride;
  *row = q;
  return q * stride == t;
}
__device__ __forceinline__ bool dump_row(const McParams& p, int64_t t, int64_t* row) {
  return dump_row_generic(p.dump_stride, t, row);
}

// attitude.euler2quat, 'zyx' (attitude.py:188-205): [yaw, pitch, roll] -> scalar-first quaternion
__device__ __forceinline__ void write_quat(double* q, double yaw, double pitch, double roll) {
  double sy, cy, sp, cp, sr, cr;
  sincos_angle(0.5 * yaw, &sy, &cy);
  sincos_angle(0.5 * pitch, &sp, &cp);
  sincos_angle(0.5 * roll, &sr, &cr);
  q[0] = cy * cp * cr + sy * sp * sr;
  q[1] = cy * cp * sr - sy * sp * cr;
  q[2] = cy * sp * cr + sy * cp * sr;
  q[3] = sy * cp * cr - cy * sp * sr;
}

// process-error accumulation of one sample (ins_data_manager.py:536-541, :761-808)
__device__ __forceinline__ void proc_accumulate(const NavState& st, const double* r, double* pe_max,
                                                double* pe_sum, double* pe_sq, double* pe_k,
                                                int64_t& pe_cnt) {
  double e[9];
  e[0] = angle_range_pi(st.yaw - r[0]);
  e[1] = angle_range_pi(st.pitch - r[1]);
  e[2] = angle_range_pi(st.roll - r[2]);
  e[3] = st.pos.x - r[3];
  e[4] = st.pos.y - r[4];
  e[5] = st.pos.z - r[5];
  e[6] = st.vel.x - r[6];
  e[7] = st.vel.y - r[7];
  e[8] = st.vel.z - r[8];
  if (pe_cnt == 0) {
#pragma unroll
    for (int c = 0; c < 9; ++c) pe_k[c] = e[c];
  }
#pragma unroll
  for (int c = 0; c < 9; ++c) {
    pe_max[c] = fmax(pe_max[c], fabs(e[c]));
    const double d = e[c] - pe_k[c];
    pe_sum[c] += d;
    pe_sq[c] += d * d;
  }
  ++pe_cnt;
}

// a^e for a small non-negative integer e (binary powering; no libm call, no stack frame)
__device__ __forceinline__ double ipow(double a, int e) {
  double r = 1.0, b = a;
#pragma unroll
  for (int bit = 0; bit < 6; ++bit) {
    if ((e >> bit) & 1) r *= b;
    b *= b;
  }
  return r;
}

// resident CTAs per SM the register allocation must allow: the throughput configuration
// (G = 1) wants many warps per scheduler to cover FP64 latency; wide groups are latency bound
// by the serial recurrence and keep their registers
#ifndef B2INS_G1_MINBLOCKS
#define B2INS_G1_MINBLOCKS 5   // measured: 1.614 / 1.642 / 1.658e10 run-steps/s for 3 / 4 / 5 (profiles/variants_r01.jsonl)
#endif
template <int G>
struct MinBlocks {
  static constexpr int value = (G == 1) ? B2INS_G1_MINBLOCKS : (G == 2 ? 3 : 2);
};

// (The warp-specialised forms of the fused launch are mc_spec_kernel.cuh and mc_av_kernel.cuh.)
template <int G, int RF, bool FED, bool PROC>
__global__ void __launch_bounds__(kThreads, MinBlocks<G>::value)
mc_kernel(const __grid_constant__ McParams p) {
  __shared__ TileSmem<PROC> sm;
  constexpr int kRunsPerWarp = 32 / G;
  constexpr bool kSplit = (G >= 4);
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int j = lane % G;
  const int role = lane & 3;
  const int64_t run_raw =
      (static_cast<int64_t>(blockIdx.x) * kWarps + warp) * kRunsPerWarp + lane / G;
  const bool active = run_raw < p.runs;
  const int64_t run = active ? run_raw : p.runs - 1;  // idle groups shadow the last run
  const int64_t grun = p.run_offset + run;            // global run id
  const uint32_t run_lo = static_cast<uint32_t>(grun), run_hi = static_cast<uint32_t>(grun >> 32);
  const bool dump = active && run < p.dump_runs;
  const bool warp_dumps = __any_sync(0xffffffffu, dump);
  const bool odo_mode = p.algo == 1;
  constexpr bool kStaged = !FED || PROC;
  constexpr int kStages = Stages<PROC>::value;

  const int64_t num_tiles = (p.n + kTile - 1) / kTile;
  if (kStaged) {
    if (threadIdx.x == 0) {
      for (int s = 0; s < kStages; ++s) {
        mbar_init(&sm.full[s], 1);
        mbar_init(&sm.empty[s], kWarps);
      }
      mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int s = 0; s < kStages && s < num_tiles; ++s) issue_tile<FED, PROC>(sm, p, s, s);
    }
  }

  // ---- sample 0 ----------------------------------------------------------
  NavState st;
  {
    const int64_t irun = p.ini_offset + run;
    const int64_t set = (irun < p.ini_sets) ? irun : 0;  // free_integration.py:85-87
    nav_init<RF>(st, p.ini + set * p.ini_rows, p.ini_rows, p.dt);
  }
  // Gauss-Markov drift carried across blocks (d[0] = 0), and the powers a^j, a^G
  double carry[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  double apj[6], aG[6];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double aa = p.accel.gm_a[c], ag = p.gyro.gm_a[c];
    apj[c] = (G > 1) ? ipow(aa, j) : 1.0;
    apj[3 + c] = (G > 1) ? ipow(ag, j) : 1.0;
    aG[c] = (G > 1) ? ipow(aa, G) : aa;
    aG[3 + c] = (G > 1) ? ipow(ag, G) : ag;
  }
  double phase[3] = {0.0, 0.0, 0.0};
  if (!FED && p.gyro.vib_type == 2) {
#pragma unroll
    for (int c = 0; c < 3; ++c)  // np.random.rand(1)*2*pi, pathgen.py:553-555
      phase[c] = (uniform01(0xFFFFFFFFu, kDrawPhase + c, run_lo, run_hi, p.k0, p.k1) * 2.0) * kPi;
  }
  if (dump && j == 0 && p.out_att) {
    const int64_t o = run * p.osr;
    p.out_att[o] = st.yaw;
    p.out_att[o + p.osc] = st.pitch;
    p.out_att[o + 2 * p.osc] = st.roll;
    p.out_pos[o] = st.pos.x;
    p.out_pos[o + p.osc] = st.pos.y;
    p.out_pos[o + 2 * p.osc] = st.pos.z;
    p.out_vel[o] = st.vel.x;
    p.out_vel[o + p.osc] = st.vel.y;
    p.out_vel[o + 2 * p.osc] = st.vel.z;
    if (p.out_quat) write_quat(p.out_quat + run * p.dump_rows * 4, st.yaw, st.pitch, st.roll);
  }
  // process-error accumulators (shifted sums: K = first error sample)
  double pe_max[9], pe_sum[9], pe_sq[9], pe_k[9];
  int64_t pe_cnt = 0;
  if (PROC) {
#pragma unroll
    for (int c = 0; c < 9; ++c) pe_max[c] = pe_sum[c] = pe_sq[c] = pe_k[c] = 0.0;
  }

  // ---- time loop ---------------------------------------------------------
  for (int64_t tile = 0; tile < num_tiles; ++tile) {
    const int s = static_cast<int>(tile % kStages);
    const uint32_t parity = static_cast<uint32_t>((tile / kStages) & 1);
    const int64_t t0 = tile * kTile;
    const int cnt = static_cast<int>(min64(kTile, p.n - t0));
    if (kStaged) {
      // refill the stage the PREVIOUS tile used (all warps have had a whole tile to release
      // it, so the issuing thread hardly ever waits), then wait for this tile's data
      if (threadIdx.x == 0 && tile >= 1 && tile - 1 + kStages < num_tiles) {
        const int sp = static_cast<int>((tile - 1) % kStages);
        mbar_wait(&sm.empty[sp], static_cast<uint32_t>(((tile - 1) / kStages) & 1));
        issue_tile<FED, PROC>(sm, p, tile - 1 + kStages, sp);
      }
      B2_CLK(cw0);
      mbar_wait(&sm.full[s], parity);
      B2_CLK(cw1);
      B2_ACC(0, cw0, cw1);
    }

    for (int base = 0; base < cnt; base += G) {
      B2_CLK(ca0);
      // ---------------- phase A: lane j prepares sample t0 + base + j --------------
      double mg[3], ma[3];  // the complete measurement of sample base + j
      double mo = 0.0;      // odometer measurement (algo 1)
      const int tj = base + j;
      const int64_t t = t0 + tj;
      {
      if (FED) {
        if (tj < cnt) {
          const int64_t o = run * p.sr + t * p.st;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            mg[c] = p.fed_gyro[o + c * p.sc];
            ma[c] = odo_mode ? 0.0 : p.fed_accel[o + c * p.sc];
          }
          if (odo_mode) mo = p.fed_odo[run * p.so_r + t * p.so_t];
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) mg[c] = ma[c] = 0.0;
        }
      } else {
        double zg[3], za[3];
        if (tj < cnt) {
          noisy_sample(p, &sm.accel[s][tj * 3], &sm.gyro[s][tj * 3], static_cast<uint32_t>(t), run_lo,
                       run_hi, run, phase, ma, mg, za, zg);
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) mg[c] = ma[c] = zg[c] = za[c] = 0.0;
        }
        B2_CLK(ca1);
        B2_ACC(1, ca0, ca1);
        // + drift: the GM state d[t] (pathgen.py:583-590) or drift*z[t] if tau = inf (:591-593)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double da = gm_block<G>(p.accel.gm_b[c] * za[c], p.accel.gm_a[c], apj[c], aG[c], j,
                                        carry[c]);
          const double dg = gm_block<G>(p.gyro.gm_b[c] * zg[c], p.gyro.gm_a[c], apj[3 + c],
                                        aG[3 + c], j, carry[3 + c]);
          ma[c] += da + p.accel.wd[c] * za[c];
          mg[c] += dg + p.gyro.wd[c] * zg[c];
        }
        if (odo_mode) {   // pathgen.odo_gen, pathgen.py:627-641: scale*ref + stdv*randn
          const double zo = (tj < cnt) ? normal_pair(static_cast<uint32_t>(t), kDrawOdo, run_lo, run_hi,
                                                     p.k0, p.k1).z0 : 0.0;
          mo = (tj < cnt) ? p.odo_scale * p.ref_odo[t] + p.odo_stdv * zo : 0.0;
        }
      }
      int64_t row;
      if (warp_dumps && dump && tj < cnt && p.out_gyro && dump_row(p, t, &row)) {
        const int64_t o = run * p.osr + row * p.ost;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          p.out_gyro[o + c * p.osc] = mg[c];
          p.out_accel[o + c * p.osc] = ma[c];
        }
        if (odo_mode && p.out_odo) p.out_odo[run * p.dump_rows + row] = mo;
      }
      if (odo_mode) {   // the odometer sample rides to phase B in the accel.x slot
        ma[0] = mo;
        ma[1] = ma[2] = 0.0;
      }
      if (G > 1) {
        SampleSlot& mine = sm.slot[warp][lane];
        mine.g[0] = mg[0]; mine.g[1] = mg[1]; mine.g[2] = mg[2];
        mine.a[0] = ma[0]; mine.a[1] = ma[1]; mine.a[2] = ma[2];
      }
      }
      // hand-over from phase A to phase B: one warp doing both only needs its own lanes
      if (G > 1) __syncwarp();

      B2_CLK(cb0);
      B2_ACC(2, ca0, cb0);
      // ---------------- phase B: serial over the G samples of the block ------------
      double keep[9];  // lane k keeps the state after sample base+k (history output)
#pragma unroll
      for (int c = 0; c < 9; ++c) keep[c] = 0.0;
      // samples of this block that are followed by a step (the last sample of the series is not)
      const int kmax = static_cast<int>(min64(min64(G, cnt - base), p.n - 1 - (t0 + base)));
      const SampleSlot* grp = &sm.slot[warp][lane - j];
      // One step of the recurrence; HIST keeps the state after sample base+k in lane k.
      auto one_step = [&](int k, bool hist) {
        Vec3 w, f;
        if (G == 1) {
          w = Vec3{mg[0], mg[1], mg[2]};
          f = Vec3{ma[0], ma[1], ma[2]};
        } else {
          const SampleSlot& sl = grp[k];
          w = Vec3{sl.g[0], sl.g[1], sl.g[2]};
          f = Vec3{sl.a[0], sl.a[1], sl.a[2]};
        }
        if (PROC) {
          // error of sample t0+base+k (state BEFORE the step), ins_data_manager.py:536-541
          if (t0 + base + k >= p.stats_start)
            proc_accumulate(st, &sm.nav[s][(base + k) * 9], pe_max, pe_sum, pe_sq, pe_k, pe_cnt);
        }
        // exact trigonometry again after every kResync-th sample (a rule in absolute time: the same for
        // every lane-group width)
        const bool resync = ((t0 + base + k + 1) & (kResync - 1)) == 0;
        nav_step<RF, kSplit, 2>(st, w, f, p.dt, p.earth_rot != 0, role, resync, odo_mode);
        if (hist && j == k) {
          keep[0] = wrap_once(st.yaw); keep[1] = st.pitch; keep[2] = wrap_once(st.roll);
          keep[3] = st.pos.x; keep[4] = st.pos.y; keep[5] = st.pos.z;
          keep[6] = st.vel.x; keep[7] = st.vel.y; keep[8] = st.vel.z;
        }
      };
      if (warp_dumps) {            // history output: the rare path keeps the simple loop
#pragma unroll 1
        for (int k = 0; k < kmax; ++k) one_step(k, true);
      } else {
        // two steps per iteration: the off-chain tail of step k (velocity, position) overlaps the
        // dependency chain of step k+1 (ptxas does not pipeline across iterations by itself)
        int k = 0;
#pragma unroll 1
        for (; k + 1 < kmax; k += 2) {
          one_step(k, false);
          one_step(k + 1, false);
        }
        if (k < kmax) one_step(k, false);
      }
      if (G > 1) __syncwarp();   // slots are rewritten by the next block
      B2_CLK(cb1);
      B2_ACC(3, cb0, cb1);
      // ---------------- histories: lane j writes the state of sample base+j+1 ---------
      int64_t hrow;
      if (warp_dumps && dump && tj < cnt && p.out_att && t + 1 < p.n && dump_row(p, t + 1, &hrow)) {
        const int64_t row = hrow;
        const int64_t o = run * p.osr + row * p.ost;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          p.out_att[o + c * p.osc] = keep[c];
          p.out_pos[o + c * p.osc] = keep[3 + c];
          p.out_vel[o + c * p.osc] = keep[6 + c];
        }
        if (p.out_quat) write_quat(p.out_quat + (run * p.dump_rows + row) * 4, keep[0], keep[1], keep[2]);
      }
    }

    if (kStaged) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&sm.empty[s]);
    }
  }

  // ---- per-run results -----------------------------------------------------
  if (PROC) {   // the last sample has no step after it: its error is accumulated here
    if (p.n - 1 >= p.stats_start)
      proc_accumulate(st, p.ref_nav + (p.n - 1) * 9, pe_max, pe_sum, pe_sq, pe_k, pe_cnt);
  }
  if (active && j == 0) {
    if (p.end_err) {
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
    }
    if (p.end_state) {
      double* e = p.end_state + run * 9;
      e[0] = wrap_once(st.yaw); e[1] = st.pitch; e[2] = wrap_once(st.roll);
      e[3] = st.pos.x; e[4] = st.pos.y; e[5] = st.pos.z;
      e[6] = st.vel.x; e[7] = st.vel.y; e[8] = st.vel.z;
    }
    if (PROC && p.proc_stats) {
      double* o = p.proc_stats + run * 27;
      const double inv = pe_cnt > 0 ? 1.0 / static_cast<double>(pe_cnt) : 0.0;
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        const double m = pe_sum[c] * inv;  // mean of (e - K)
        o[c] = pe_max[c];
        o[9 + c] = pe_k[c] + m;
        o[18 + c] = sqrt(fmax(pe_sq[c] * inv - m * m, 0.0));
      }
    }
  }
}

}  // namespace b2ins

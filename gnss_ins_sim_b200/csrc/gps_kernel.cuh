// K6: GPS measurement generator.  Replaces pathgen.gps_gen (gnss_ins_sim/pathgen/pathgen.py:596-625)
// for all Monte-Carlo runs at once: gps[r][k][0:3] = ref[k][0:3] + pos_err * N(0,1),
// gps[r][k][3:6] = ref[k][3:6] + stdv * N(0,1).  With LLA positions (gps_type 0) the horizontal
// sigmas are converted from metres to radians with the radii of curvature at the FIRST reference
// sample, as the reference does (pathgen.py:617-620).
// Noise spec: the three Box-Muller pairs (k, draw 24 + j, global run) give
// (pos0, pos1), (pos2, vel0), (vel1, vel2).  One thread per (run, GPS sample): 48 B read (shared
// reference, L2-resident) and 48 B written per unit: an HBM-write-bound map.
#pragma once
#include "common.cuh"
#include "mech.cuh"

namespace b2ins {

constexpr uint32_t kPairGps = 24;

struct GpsParams {
  int64_t m, runs, run_offset;
  const double* ref;   // [m][6]
  double* out;         // [runs][m][6]
  double stdp[3], stdv[3];
  uint32_t k0, k1;
  int gps_type;
};

__global__ void __launch_bounds__(256) gps_noise_kernel(const __grid_constant__ GpsParams p) {
  double sd0 = p.stdp[0], sd1 = p.stdp[1];
  if (p.gps_type == 0) {
    const GeoParam gp = geo_param(p.ref[0], p.ref[2]);
    sd0 = div_nr(sd0, gp.rm);
    sd1 = div_nr(div_nr(sd1, gp.rn), gp.cl);
  }
  const int64_t total = p.m * p.runs;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t r = i / p.m;
    const int64_t k = i - r * p.m;
    const uint64_t run = static_cast<uint64_t>(p.run_offset + r);
    const uint32_t rl = static_cast<uint32_t>(run), rh = static_cast<uint32_t>(run >> 32);
    double z[6];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const Normal2 zz = normal_pair(static_cast<uint32_t>(k), kPairGps + j, rl, rh, p.k0, p.k1);
      z[2 * j] = zz.z0;
      z[2 * j + 1] = zz.z1;
    }
    const double* ref = p.ref + k * 6;
    double* o = p.out + i * 6;
    o[0] = ref[0] + sd0 * z[0];
    o[1] = ref[1] + sd1 * z[1];
    o[2] = ref[2] + p.stdp[2] * z[2];
    o[3] = ref[3] + p.stdv[0] * z[3];
    o[4] = ref[4] + p.stdv[1] * z[4];
    o[5] = ref[5] + p.stdv[2] * z[5];
  }
}

}  // namespace b2ins

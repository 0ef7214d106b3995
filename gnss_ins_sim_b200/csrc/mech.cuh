// Strapdown free-integration mechanization: the per-timestep recurrence of
// FreeIntegration.run (demo_algorithms/free_integration.py:63-174) with the L1 math it
// calls (attitude.py:344-371 euler2dcm zyx, :679-721 euler_update_zyx, :758-770 cross3;
// geoparams.py:25-53 geo_param, :70-87 lla2ecef), written once for all kernels.
// The whole state lives in registers; everything is double.
#pragma once
#include "common.cuh"

namespace b2ins {

struct Vec3 {
  double x, y, z;
};

// n -> b direction cosine matrix, row-major
struct Dcm {
  double c00, c01, c02, c10, c11, c12, c20, c21, c22;
};

struct SinCos3 {
  double sy, cy, sp, cp, sr, cr;  // yaw, pitch, roll
};

// sin/cos of an angle that normally lives in [-pi, pi] (Euler angles after their wrap,
// latitude, longitude).  sincos_bounded's three-term reduction stays accurate far beyond that
// (|x| < 1e6: error < 1e-10); larger magnitudes are only reachable after the Euler-angle
// singularity at pitch = +-pi/2 has blown a rate up, where the recurrence is meaningless
// anyway -- they are mapped to the angle 0 by a select (branch-free, off the critical path).
__device__ __forceinline__ void sincos_angle(double x, double* s, double* c) {
  sincos_bounded(fabs(x) <= 1.0e6 ? x : 0.0, s, c);
}

__device__ __forceinline__ SinCos3 sincos3(double yaw, double pitch, double roll) {
  SinCos3 t;
  sincos_angle(yaw, &t.sy, &t.cy);
  sincos_angle(pitch, &t.sp, &t.cp);
  sincos_angle(roll, &t.sr, &t.cr);
  return t;
}

// attitude.euler2dcm, 'zyx' branch: attitude.py:361-371
__device__ __forceinline__ Dcm dcm_from_sincos(const SinCos3& t) {
  Dcm c;
  c.c00 = t.cp * t.cy;
  c.c01 = t.cp * t.sy;
  c.c02 = -t.sp;
  c.c10 = t.sr * t.sp * t.cy - t.cr * t.sy;
  c.c11 = t.sr * t.sp * t.sy + t.cr * t.cy;
  c.c12 = t.cp * t.sr;
  c.c20 = t.sp * t.cr * t.cy + t.sy * t.sr;
  c.c21 = t.sp * t.cr * t.sy - t.cy * t.sr;
  c.c22 = t.cp * t.cr;
  return c;
}

__device__ __forceinline__ Vec3 mul_t(const Dcm& c, const Vec3& v) {  // c^T . v
  return Vec3{c.c00 * v.x + c.c10 * v.y + c.c20 * v.z, c.c01 * v.x + c.c11 * v.y + c.c21 * v.z,
              c.c02 * v.x + c.c12 * v.y + c.c22 * v.z};
}
// c_bn^T . v and c_bn . v without forming the matrix: the ZYX dcm is Rx(roll) Ry(pitch) Rz(yaw),
// so each product is three planar rotations (12 multiply-adds instead of 16 + 9).  Same value as
// dcm_from_sincos + mul / mul_t up to rounding.
__device__ __forceinline__ Vec3 rot_b2n(const SinCos3& t, const Vec3& v) {   // c^T . v
  // undo roll (about x)
  const double y1 = t.cr * v.y - t.sr * v.z;
  const double z1 = t.sr * v.y + t.cr * v.z;
  // undo pitch (about y)
  const double x2 = t.cp * v.x + t.sp * z1;
  const double z2 = -t.sp * v.x + t.cp * z1;
  // undo yaw (about z)
  return Vec3{t.cy * x2 - t.sy * y1, t.sy * x2 + t.cy * y1, z2};
}
__device__ __forceinline__ Vec3 rot_n2b(const SinCos3& t, const Vec3& v) {   // c . v
  const double x1 = t.cy * v.x + t.sy * v.y;
  const double y1 = -t.sy * v.x + t.cy * v.y;
  const double x2 = t.cp * x1 - t.sp * v.z;
  const double z2 = t.sp * x1 + t.cp * v.z;
  return Vec3{x2, t.cr * y1 + t.sr * z2, -t.sr * y1 + t.cr * z2};
}

// attitude.cross3: attitude.py:758-770
__device__ __forceinline__ Vec3 cross3(const Vec3& a, const Vec3& b) {
  return Vec3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// geoparams.geo_param: geoparams.py:25-53, given sin/cos of the latitude
struct GeoParam {
  double rm, rn, g, sl, cl;
};
__device__ __forceinline__ GeoParam geo_param_sc(double sl, double cl, double h) {
  GeoParam p;
  p.sl = sl;
  p.cl = cl;
  const double sl_sqr = p.sl * p.sl;
  const double q = 1.0 - kESqr * sl_sqr;       // in [0.9933, 1]: no special cases
  const double sq = sqrt_nr(q);
  const double inv_sq = rcp_nr(sq);
  p.rm = (kRe * (1 - kESqr)) * rcp_nr(sq * q);
  p.rn = kRe * inv_sq;
  const double g1 = kNormalGravity * (1 + kGravK * sl_sqr) * inv_sq;
  p.g = g1 * (1.0 - (2.0 / kRe) * (1.0 + kFlat + kGravM - 2.0 * kFlat * sl_sqr) * h +
              3.0 * h * h / kRe / kRe);
  return p;
}
__device__ __forceinline__ GeoParam geo_param(double lat, double h) {
  double sl, cl;
  sincos_angle(lat, &sl, &cl);
  return geo_param_sc(sl, cl, h);
}

// geoparams.lla2ecef: geoparams.py:70-87
__device__ __forceinline__ Vec3 lla2ecef(double lat, double lon, double alt) {
  double sl, cl, so, co;
  sincos_angle(lat, &sl, &cl);
  sincos_angle(lon, &so, &co);
  const double r = kRe / sqrt(1.0 - kESqr * sl * sl);
  const double rho = (r + alt) * cl;
  return Vec3{rho * co, rho * so, (r * (1.0 - kESqr) + alt) * sl};
}

// attitude.angle_range_pi: attitude.py:799-812 (python float % : result has the sign of 2pi)
__device__ __forceinline__ double angle_range_pi(double x) {
  double m = fmod(x, kTwoPi);
  if (m < 0.0) m += kTwoPi;
  if (m > kPi) m -= kTwoPi;
  return m;
}

// The navigation state of one Monte-Carlo run.
struct NavState {
  double yaw, pitch, roll;
  SinCos3 sc;  // sin/cos of (yaw, pitch, roll): euler2dcm(att[i]) of step i IS the
               // cos/sin euler_update_zyx needs at step i+1, so it is computed once
  double sl, cl;  // sin/cos of the latitude (ref_frame 0): geo_param of the NEXT step
  Vec3 vel_b;  // body velocity   (ref_frame 1 state)
  Vec3 vel;    // NED velocity    (ref_frame 0 state; ref_frame 1 output)
  Vec3 pos;    // ECEF-offset xyz (ref_frame 1) or lat, lon, alt (ref_frame 0)
  double g;    // gravity: geo_param(r0) or the ini override
  bool fixed_g;  // false: ref_frame 0 without override -> geo_param(pos) every step
};

// Refresh the cached sin/cos after the angles (and the latitude) moved.
// SPLIT (lane groups of >= 4 lanes): the state is replicated across the group, so the
// three (four with the latitude) independent sincos evaluations are spread over the lanes
// of each 4-lane subgroup -- lane role q evaluates angle q -- and exchanged by shuffles:
// one sincos worth of instruction issue instead of three or four.
template <int RF, bool SPLIT>
__device__ __forceinline__ void refresh_trig(NavState& s, int role) {
  if (!SPLIT) {
    s.sc = sincos3(s.yaw, s.pitch, s.roll);
    if (RF == 0) sincos_angle(s.pos.x, &s.sl, &s.cl);
  } else {
    double a = s.roll;
    if (role == 0) a = s.yaw;
    if (role == 1) a = s.pitch;
    if (RF == 0 && role == 3) a = s.pos.x;
    double sv, cv;
    sincos_angle(a, &sv, &cv);
    s.sc.sy = __shfl_sync(0xffffffffu, sv, 0, 4);
    s.sc.cy = __shfl_sync(0xffffffffu, cv, 0, 4);
    s.sc.sp = __shfl_sync(0xffffffffu, sv, 1, 4);
    s.sc.cp = __shfl_sync(0xffffffffu, cv, 1, 4);
    s.sc.sr = __shfl_sync(0xffffffffu, sv, 2, 4);
    s.sc.cr = __shfl_sync(0xffffffffu, cv, 2, 4);
    if (RF == 0) {
      s.sl = __shfl_sync(0xffffffffu, sv, 3, 4);
      s.cl = __shfl_sync(0xffffffffu, cv, 3, 4);
    }
  }
}

// free_integration.py:96-102 / :126-132 -- sample 0
template <int RF>
__device__ __forceinline__ void nav_init(NavState& s, const double* __restrict__ ini,
                                         int ini_rows) {
  const double lat = ini[0], lon = ini[1], alt = ini[2];
  s.vel_b = Vec3{ini[3], ini[4], ini[5]};
  s.yaw = ini[6];
  s.pitch = ini[7];
  s.roll = ini[8];
  s.sl = s.cl = 0.0;
  if (RF == 1) {
    s.pos = lla2ecef(lat, lon, alt);
    s.g = (ini_rows > 9) ? ini[9] : geo_param(lat, alt).g;  // free_integration.py:89-93
    s.fixed_g = true;
  } else {
    s.pos = Vec3{lat, lon, alt};
    s.fixed_g = ini_rows > 9;  // free_integration.py:143-146
    s.g = s.fixed_g ? ini[9] : 0.0;
  }
  refresh_trig<RF, false>(s, 0);
  const Dcm c = dcm_from_sincos(s.sc);
  s.vel = mul_t(c, s.vel_b);
}

// attitude.euler_update_zyx (attitude.py:679-721) using the cached sin/cos of the
// current angles.  t*tan(pitch) is evaluated as (t/cos(pitch))*sin(pitch).
__device__ __forceinline__ void euler_update(NavState& s, const Vec3& w, double dt) {
  const double t = w.z * s.sc.cr + w.y * s.sc.sr;
  const double phi_dot = div_nr(t, s.sc.cp);
  const double theta_dot = w.y * s.sc.cr - w.z * s.sc.sr;
  const double psi_dot = w.x + phi_dot * s.sc.sp;
  double y0 = s.yaw + phi_dot * dt;
  double y1 = s.pitch + theta_dot * dt;
  double y2 = s.roll + psi_dot * dt;
  // pitch reflection (attitude.py:703-710), as selects: one straight-line block per step
  const bool hi = y1 > kHalfPi, lo = y1 < -kHalfPi;
  y1 = hi ? (kPi - y1) : (lo ? (-kPi - y1) : y1);
  const bool flip = hi || lo;
  y0 = flip ? y0 + kPi : y0;
  y2 = flip ? y2 + kPi : y2;
  // ONE +-2pi wrap of yaw and roll (:712-720)
  y0 = (y0 > kPi) ? (y0 - kTwoPi) : ((y0 < -kPi) ? (y0 + kTwoPi) : y0);
  y2 = (y2 > kPi) ? (y2 - kTwoPi) : ((y2 < -kPi) ? (y2 + kTwoPi) : y2);
  s.yaw = y0;
  s.pitch = y1;
  s.roll = y2;
}

// One step i-1 -> i with the measurements of sample i-1.
//
// odo = false: FreeIntegration.run (free_integration.py:104-116 / :133-172).
// odo = true : free_integration_odo (free_integration_odo.py:104-112 / :121-158): same attitude
//              recurrence, body velocity = [odometer, 0, 0]; the odometer sample rides in accel.x.
template <int RF, bool SPLIT>
__device__ __forceinline__ void nav_step(NavState& s, const Vec3& gyro, const Vec3& accel,
                                         double dt, bool earth_rot, int role, bool odo) {
  if (RF == 1) {
    // free_integration.py:104-116
    // c_bn.dot(g_n) with g_n = [0,0,g]: third column of the OLD dcm, from the old sin/cos
    const Vec3 cg{-s.sc.sp * s.g, s.sc.cp * s.sc.sr * s.g, s.sc.cp * s.sc.cr * s.g};
    const Vec3 wxv = cross3(gyro, s.vel_b);
    const Vec3 vel_old = s.vel;
    euler_update(s, gyro, dt);
    if (odo) {
      s.vel_b = Vec3{accel.x, 0.0, 0.0};
    } else {
      s.vel_b.x = s.vel_b.x + (accel.x + cg.x) * dt - wxv.x * dt;
      s.vel_b.y = s.vel_b.y + (accel.y + cg.y) * dt - wxv.y * dt;
      s.vel_b.z = s.vel_b.z + (accel.z + cg.z) * dt - wxv.z * dt;
    }
    refresh_trig<RF, SPLIT>(s, role);
    s.vel = rot_b2n(s.sc, s.vel_b);
    s.pos.x += vel_old.x * dt;
    s.pos.y += vel_old.y * dt;
    s.pos.z += vel_old.z * dt;
  } else {
    // free_integration.py:133-172
    const GeoParam p = geo_param_sc(s.sl, s.cl, s.pos.z);
    const double rm_e = p.rm + s.pos.z;
    const double rn_e = p.rn + s.pos.z;
    const double g = s.fixed_g ? s.g : p.g;
    const double inv_rn = rcp_nr(rn_e), inv_rm = rcp_nr(rm_e), inv_cl = rcp_nr(p.cl);
    const Vec3 w_en{s.vel.y * inv_rn, -s.vel.x * inv_rm, -s.vel.y * p.sl * inv_cl * inv_rn};
    Vec3 w_ie{0.0, 0.0, 0.0};
    if (earth_rot) {
      w_ie.x = kWie * p.cl;
      w_ie.z = -kWie * p.sl;
    }
    const Vec3 w_sum{w_en.x + w_ie.x, w_en.y + w_ie.y, w_en.z + w_ie.z};
    const Vec3 cw = rot_n2b(s.sc, w_sum);   // c_bn of step i-1
    const Vec3 w_nb{gyro.x - cw.x, gyro.y - cw.y, gyro.z - cw.z};
    const Vec3 fa = rot_b2n(s.sc, accel);
    const Vec3 w2{2 * w_ie.x + w_en.x, 2 * w_ie.y + w_en.y, 2 * w_ie.z + w_en.z};
    const Vec3 cor = cross3(w2, s.vel);
    const Vec3 vel_old = s.vel;
    euler_update(s, w_nb, dt);
    s.pos.x += vel_old.x * inv_rm * dt;
    s.pos.y += vel_old.y * inv_rn * inv_cl * dt;
    s.pos.z += (-vel_old.z) * dt;
    refresh_trig<RF, SPLIT>(s, role);
    if (odo) {
      s.vel = rot_b2n(s.sc, Vec3{accel.x, 0.0, 0.0});   // c_bn of step i
    } else {
      s.vel.x = vel_old.x + (fa.x - cor.x) * dt;
      s.vel.y = vel_old.y + (fa.y - cor.y) * dt;
      s.vel.z = vel_old.z + (fa.z + g - cor.z) * dt;
    }
    // vel_b[i] = c_bn(i).dot(vel[i]) (:172) is not an output of the plugin; not computed
  }
}

}  // namespace b2ins

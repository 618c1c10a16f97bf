// Per-particle NON-LINEAR tracking: the "drift_kick_drift" (Bmad-X) maps of Drift, Quadrupole,
// Dipole / RBend and TransverseDeflectingCavity and the "second_order" maps (T_ijk p_j p_k) of
// Drift, Quadrupole, Sextupole, Dipole / RBend.
//
// The reference tracks each such element with 50-150 elementwise PyTorch kernels over the whole
// beam.  Here a RUN of consecutive non-linear elements (markers and inactive monitors in between
// are no-ops) is one streaming pass: a CTA stages a tile of particles with a TMA bulk copy, every
// thread keeps its particles' six coordinates in registers, walks the ops of the run with the
// per-(setting, op) constants broadcast from shared memory, and the tile leaves through a bulk
// store -- 28 B read + 28 B written per (particle, setting) for the whole run.
//
//   ch_nonlinear_constants: one CTA per setting, one thread per op: element parameters (slot
//     table) -> constants in fp64 (sin/cos of tilts and bend angles, fringe kicks, the 39
//     second-order coefficients with the reference's singularity-free compound functions),
//     rounded once to the beam dtype.
//   ch_track_nonlinear: the particle pass.
//
// Numerics.  Bmad-X coordinates (z, pz) are obtained from (tau, delta) with cancellation-free
// forms of the reference's expressions (p^2 - p0c^2 = delta p0c (2 E0 + delta p0c), etc.), so
// the float32 path keeps ~1e-7 relative accuracy where the reference's float32 path cancels to
// ~1e-4; consecutive drift_kick_drift ops stay in Bmad-X coordinates (the reference converts
// back and forth between every element; same result up to rounding).  The bend body and the TDC
// kick, whose formulas subtract path lengths of order L, are always evaluated in fp64.
//
// Reference behaviour restated here (desy-ml/cheetah @ 60d1053):
//   cheetah/utils/bmadx.py:7-318, cheetah/accelerator/drift.py:106-154,
//   quadrupole.py:168-251, dipole.py:183-370, transverse_deflecting_cavity.py:122-209,
//   cheetah/track_methods.py:80-281 (base_ttensor), cheetah/utils/autograd.py:108-670,
//   cheetah/accelerator/element.py:195-225, drift.py:67-83, quadrupole.py:112-143,
//   sextupole.py:90-116, dipole.py:396-466
#include <type_traits>

#include "ch_common.cuh"

namespace ch {
namespace {

constexpr double kPi = 3.141592653589793;
constexpr double kC = 299792458.0;

struct Program {
  const int32_t* opcodes;
  const int32_t* op_flags;
  const int32_t* slot_begin;
  const ScalarRef* slots;
};

__device__ __forceinline__ double slot_value(const Program& prog, int32_t slot, int64_t b) {
  const ScalarRef ref = prog.slots[slot];
  return load_scalar(ref.ptr, b * ref.stride, ref.dtype);
}

// cos(sqrt x) and si(sqrt x) = sin(sqrt x) / sqrt x continued to x < 0: the real closed form of
// the reference's complex sqrt / cos / sinc
__device__ __forceinline__ void cos_si(double x, double& c, double& s) {
  if (x > 0.0) {
    const double r = sqrt(x);
    double sn;
    sincos(r, &sn, &c);
    s = sn / r;
  } else if (x < 0.0) {
    const double r = sqrt(-x);
    c = cosh(r);
    s = sinh(r) / r;
  } else {
    c = 1.0;
    s = 1.0;
  }
}

// the compound functions of cheetah/utils/autograd.py with their coded limits
__device__ __forceinline__ double si1mdiv(double x) {  // :108-128
  if (x == 0.0) return 1.0 / 6.0;
  if (fabs(x) < 1e-3)
    return 1.0 / 6.0 + x * (-1.0 / 120.0 + x * (1.0 / 5040.0 + x * (-1.0 / 362880.0)));
  double c, s;
  cos_si(x, c, s);
  return (1.0 - s) / x;
}
__device__ __forceinline__ double sicos1mdiv(double x) {  // :149-174 (limit 1/6 as coded)
  if (x == 0.0) return 1.0 / 6.0;
  double c, s;
  cos_si(x, c, s);
  return (1.0 - s * c) / x;
}
__device__ __forceinline__ double sipsicos3mdiv(double x) {  // :209-235
  if (x == 0.0) return 0.0;
  double c, s;
  cos_si(x, c, s);
  return (3.0 - 4.0 * s + s * c) / (2.0 * x);
}
__device__ __forceinline__ double cossqrtmcosdivdiff(double a, double b) {  // :361-388
  double ca, sa, cb, sb;
  cos_si(a, ca, sa);
  cos_si(b, cb, sb);
  return a != b ? (cb - ca) / (a - b) : 0.5 * sa;
}
__device__ __forceinline__ double simsidivdiff(double a, double b) {  // :433-461
  double ca, sa, cb, sb;
  cos_si(a, ca, sa);
  cos_si(b, cb, sb);
  if (a != b) return (sa - sb) / (b - a);
  return b != 0.0 ? 0.5 * (sb - cb) / b : 1.0 / 6.0;
}
__device__ __forceinline__ double si2msi2divdiff(double a, double b) {  // :546-579
  double ca, sa, cb, sb;
  cos_si(a, ca, sa);
  cos_si(b, cb, sb);
  if (a != b) return (sb * sb - sa * sa) / (a - b);
  return b != 0.0 ? (1.0 - cb * cb - b * sb * cb) / (b * b) : 1.0 / 3.0;
}

// ---- constant-block layouts ------------------------------------------------------------------
// header (CH_NL_HEADER scalars per setting)
enum {
  H_P0C = 0, H_MC2 = 1, H_E0 = 2, H_BETA0 = 3, H_MC2_E0_SQ = 4, H_LENGTH = 5, H_CHARGE = 6,
  H_INV_P0C = 7, H_TWO_E0_OVER_P0C = 8
};
// drift_kick_drift ops
enum { D_L = 0 };
enum { Q_L = 0, Q_K1 = 1, Q_COS = 2, Q_SIN = 3, Q_XOFF = 4, Q_YOFF = 5, Q_STEP = 6 };
enum {
  B_L = 0, B_ANGLE = 1, B_COS = 2, B_SIN = 3, B_G = 4, B_HX1 = 5, B_HY1 = 6, B_HX2 = 7, B_HY2 = 8,
  B_COSA = 9, B_SINA = 10, B_LSINC = 11, B_L2GCOSC = 12
};
enum { T_HALF = 0, T_COS = 1, T_SIN = 2, T_XOFF = 3, T_YOFF = 4, T_V = 5, T_KRF = 6, T_PHASE = 7 };
// second-order ops: frame change (10), body R (9), T (39)
enum {
  S_COS = 0, S_SIN = 1, S_OX = 2, S_OY = 3, S_KX1 = 4, S_KY1 = 5, S_KX2 = 6, S_KY2 = 7, S_MX = 8,
  S_MY = 9, S_R = 10, S_T = 19
};
enum {  // body R entries (track_methods.py:62-75)
  R_CX = 0, R_SX = 1, R_10 = 2, R_CY = 3, R_SY = 4, R_32 = 5, R_05 = 6, R_15 = 7, R_56 = 8
};

// order of the T entries in the block (i, j, k as in track_methods.py:147-279)
enum {
  T000, T001, T011, T005, T015, T055, T022, T023, T033,
  T100, T101, T111, T105, T115, T155, T122, T123, T133,
  T202, T203, T212, T213, T225, T235,
  T302, T303, T312, T313, T325, T335,
  T400, T401, T411, T405, T415, T455, T422, T423, T433,
  T_COUNT
};
static_assert(S_T + T_COUNT <= CH_NL_BLOCK_SECOND_ORDER, "second-order block too small");

// base_ttensor (track_methods.py:80-281) in fp64; out[T_COUNT]
__device__ void second_order_coefficients(double L, double k1, double k2, double hx, double beta,
                                          double igamma2, double* t) {
  const double kx2 = k1 + hx * hx, ky2 = -k1;
  double cx, six, cy, siy;
  cos_si(kx2 * L * L, cx, six);
  cos_si(ky2 * L * L, cy, siy);
  const double sx = six * L, sy = siy * L;
  double ch, sih;
  cos_si(0.25 * kx2 * L * L, ch, sih);
  const double dx = 0.5 * L * L * sih * sih;
  const double L3 = L * L * L;
  const double a = kx2 * L * L, b = ky2 * L * L;
  const double fx = L3 * si1mdiv(a);
  const double f2y = L3 * sicos1mdiv(b);
  const double j1 = fx;
  const double j2 = L3 * sipsicos3mdiv(a);
  const double j3 = kx2 != 0.0 ? (15.0 * L - 22.5 * sx + 9.0 * sx * cx - 1.5 * sx * cx * cx +
                                  kx2 * sx * sx * sx) / (6.0 * kx2 * kx2 * kx2)
                               : L3 * L3 * L / 56.0;
  const double jden = kx2 - 4.0 * ky2;
  const double jc = L * L * cossqrtmcosdivdiff(a, b);
  const double js = L3 * simsidivdiff(a, b);
  const double jd = L3 * L * si2msi2divdiff(a, b);
  const double jf = jden != 0.0 ? (f2y - fx) / jden : L3 * L * L / 120.0;
  const double khk = k2 + 2.0 * hx * k1;
  const double ib = 1.0 / beta, ib2 = ib * ib, ib3 = ib2 * ib;
  const double hx2 = hx * hx, dx2 = dx * dx;

  t[T000] = -khk * (sx * sx + dx) / 6.0 - 0.5 * hx * kx2 * sx * sx;
  t[T001] = 2.0 * (-khk * sx * dx / 6.0 + 0.5 * hx * sx * cx);
  t[T011] = -khk * dx2 / 6.0 + 0.5 * hx * dx * cx;
  t[T005] = 2.0 * (-hx / 12.0 * ib * khk * (3.0 * sx * j1 - dx2) + 0.5 * hx2 * ib * sx * sx +
                   0.25 * ib * k1 * L * sx);
  t[T015] = 2.0 * (-hx / 12.0 * ib * khk * (sx * dx2 - 2.0 * cx * j2) +
                   0.25 * hx2 * ib * (sx * dx + cx * j1) - 0.25 * ib * (sx + L * cx));
  t[T055] = -hx2 / 6.0 * ib2 * khk * (dx2 * dx - 2.0 * sx * j2) + 0.5 * hx2 * hx * ib2 * sx * j1 -
            0.5 * hx * ib2 * L * sx - 0.5 * hx * ib2 * igamma2 * dx;
  t[T022] = k1 * k2 * jd + 0.5 * (k2 + hx * k1) * dx;
  t[T023] = 2.0 * (0.5 * k2 * js);
  t[T033] = k2 * jd - 0.5 * hx * dx;
  t[T100] = -khk * sx * (1.0 + 2.0 * cx) / 6.0;
  t[T101] = -2.0 * khk * dx * (1.0 + 2.0 * cx) / 6.0;
  t[T111] = -khk * sx * dx / 3.0 - 0.5 * hx * sx;
  t[T105] = 2.0 * (-hx / 12.0 * ib * khk * (3.0 * cx * j1 + sx * dx) -
                   0.25 * ib * k1 * (sx - L * cx));
  t[T115] = 2.0 * (-hx / 12.0 * ib * khk * (3.0 * sx * j1 + dx2) + 0.25 * ib * k1 * L * sx);
  t[T155] = -hx2 / 6.0 * ib2 * khk * (sx * dx2 - 2.0 * cx * j2) -
            0.5 * hx * ib2 * k1 * (cx * j1 - sx * dx) - 0.5 * hx * ib2 * igamma2 * sx;
  t[T122] = k1 * k2 * js + 0.5 * (k2 + hx * k1) * sx;
  t[T123] = 2.0 * (0.5 * k2 * jc);
  t[T133] = k2 * js - 0.5 * hx * sx;
  t[T202] = 2.0 * (0.5 * k2 * (cy * jc - 2.0 * k1 * sy * js) + 0.5 * hx * k1 * sx * sy);
  t[T203] = 2.0 * (0.5 * k2 * (sy * jc - 2.0 * cy * js) + 0.5 * hx * sx * cy);
  t[T212] = 2.0 * (0.5 * k2 * (cy * js - 2.0 * k1 * sy * jd) + 0.5 * hx * k1 * dx * sy);
  t[T213] = 2.0 * (0.5 * k2 * (sy * js - 2.0 * cy * jd) + 0.5 * hx * dx * cy);
  t[T225] = 2.0 * (0.5 * hx * ib * k2 * (cy * jd - 2.0 * k1 * sy * jf) +
                   0.5 * hx2 * ib * k1 * j1 * sy - 0.25 * ib * k1 * L * sy);
  t[T235] = 2.0 * (0.5 * hx * ib * k2 * (sy * jd - 2.0 * cy * jf) + 0.5 * hx2 * ib * j1 * cy -
                   0.25 * ib * (sy + L * cy));
  t[T302] = 2.0 * (0.5 * k1 * k2 * (2.0 * cy * js - sy * jc) + 0.5 * (k2 + hx * k1) * sx * cy);
  t[T303] = 2.0 * (0.5 * k2 * (2.0 * k1 * sy * js - cy * jc) + 0.5 * (k2 + hx * k1) * sx * sy);
  t[T312] = 2.0 * (0.5 * k1 * k2 * (2.0 * cy * jd - sy * js) + 0.5 * (k2 + hx * k1) * dx * cy);
  t[T313] = 2.0 * (0.5 * k2 * (2.0 * k1 * sy * jd - cy * js) + 0.5 * (k2 + hx * k1) * dx * sy);
  t[T325] = 2.0 * (0.5 * hx * ib * k1 * k2 * (2.0 * cy * jf - sy * jd) +
                   0.5 * hx * ib * (k2 + hx * k1) * j1 * cy + 0.25 * ib * k1 * (sy - L * cy));
  t[T335] = 2.0 * (0.5 * hx * ib * k2 * (2.0 * k1 * sy * jf - cy * jd) +
                   0.5 * hx * ib * (k2 + hx * k1) * j1 * sy - 0.25 * ib * k1 * L * sy);
  t[T400] = -(hx / 12.0 * ib * khk * (sx * dx + 3.0 * j1) - 0.25 * ib * k1 * (L - sx * cx));
  t[T401] = -2.0 * (hx / 12.0 * ib * khk * dx2 + 0.25 * ib * k1 * sx * sx);
  t[T411] = -(hx / 6.0 * ib * khk * j2 - 0.5 * ib * sx - 0.25 * ib * k1 * (j1 - sx * dx));
  t[T405] = -2.0 * (hx2 / 12.0 * ib2 * khk * (3.0 * dx * j1 - 4.0 * j2) +
                    0.25 * hx * ib2 * k1 * j1 * (1.0 + cx) + 0.5 * hx * ib2 * igamma2 * sx);
  t[T415] = -2.0 * (hx2 / 12.0 * ib2 * khk * (dx * dx2 - 2.0 * sx * j2) +
                    0.25 * hx * ib2 * k1 * sx * j1 + 0.5 * hx * ib2 * igamma2 * dx);
  t[T455] = -(hx2 * hx / 6.0 * ib3 * khk * (3.0 * j3 - 2.0 * dx * j2) +
              hx2 / 6.0 * ib3 * k1 * (sx * dx2 - j2 * (1.0 + 2.0 * cx)) +
              1.5 * ib3 * igamma2 * (hx2 * j1 - L));
  t[T422] = -(-hx * ib * k1 * k2 * jf - 0.5 * hx * ib * (k2 + hx * k1) * j1 +
              0.25 * ib * k1 * (L - cy * sy));
  t[T423] = -2.0 * (-0.5 * hx * ib * k2 * jd - 0.25 * ib * k1 * sy * sy);
  t[T433] = -(-hx * ib * k2 * jf + 0.5 * hx2 * ib * j1 - 0.25 * ib * (L + cy * sy));
}

// body R of base_rmatrix (track_methods.py:17-77) in fp64; r[9]
__device__ void body_rmatrix(double L, double k1, double hx, double beta, double igamma2,
                             double* r) {
  const double kx2 = k1 + hx * hx, ky2 = -k1;
  double cx, six, cy, siy;
  cos_si(kx2 * L * L, cx, six);
  cos_si(ky2 * L * L, cy, siy);
  const double sx = six * L, sy = siy * L;
  double ch, sih;
  cos_si(0.25 * kx2 * L * L, ch, sih);
  const double dx = hx * 0.5 * L * L * sih * sih;
  r[R_CX] = cx;
  r[R_SX] = sx;
  r[R_10] = -kx2 * sx;
  r[R_CY] = cy;
  r[R_SY] = sy;
  r[R_32] = -ky2 * sy;
  r[R_05] = dx / beta;
  r[R_15] = sx * hx / beta;
  r[R_56] = hx * hx * L * L * L * si1mdiv(kx2 * L * L) / (beta * beta) - L / (beta * beta) * igamma2;
}

// pole-face kicks of the linear / second-order dipole (dipole.py:430-466)
__device__ __forceinline__ void edge_kicks(double hx, double e, double fint, double gap, double& kx,
                                           double& ky) {
  const double se = sin(e);
  const double phi = fint * hx * gap / cos(e) * (1.0 + se * se);
  kx = hx * tan(e);
  ky = -hx * tan(e - phi);
}

// The constants table is always fp64: the bend body and the TDC kick are evaluated in fp64 from
// unrounded constants, everything else is rounded to the beam dtype while being staged.
__global__ void __launch_bounds__(64)
nonlinear_constants_kernel(Program prog, int32_t op_begin, int32_t n_ops, ScalarRef energy,
                           ScalarRef mass, ScalarRef charge, int32_t block, double* __restrict__ out) {
  __shared__ double lengths[64];
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  double* base = out + b * (CH_NL_HEADER + static_cast<int64_t>(n_ops) * block);

  const double E0 = load_scalar(energy.ptr, b * energy.stride, energy.dtype);
  const double mc2 = load_scalar(mass.ptr, 0, mass.dtype);
  const double q = charge.ptr ? load_scalar(charge.ptr, 0, charge.dtype) : -1.0;
  const double p0c = sqrt(E0 * E0 - mc2 * mc2);
  const double gamma = E0 / mc2, igamma2 = 1.0 / (gamma * gamma), beta = sqrt(1.0 - igamma2);

  double length = 0.0;
  if (tid < n_ops) {
    const int32_t op = op_begin + tid;
    const int32_t code = prog.opcodes[op];
    const int32_t flags = prog.op_flags[op];
    const int32_t s0 = prog.slot_begin[op];
    double c[CH_NL_BLOCK_SECOND_ORDER];
    for (int i = 0; i < block; ++i) c[i] = 0.0;
    switch (code) {
      case CH_OP_DKD_DRIFT:
        length = slot_value(prog, s0, b);
        c[D_L] = length;
        break;
      case CH_OP_DKD_QUADRUPOLE: {
        length = slot_value(prog, s0, b);
        const double tilt = slot_value(prog, s0 + 2, b);
        double sn, cs;
        sincos(tilt, &sn, &cs);
        c[Q_L] = length;
        c[Q_K1] = slot_value(prog, s0 + 1, b);
        c[Q_COS] = cs;
        c[Q_SIN] = sn;
        c[Q_XOFF] = slot_value(prog, s0 + 3, b);
        c[Q_YOFF] = slot_value(prog, s0 + 4, b);
        c[Q_STEP] = length / static_cast<double>(flags > 0 ? flags : 1);
        break;
      }
      case CH_OP_DKD_DIPOLE: {
        // slots: length, angle, e1, e2, fint, fint_exit, gap, gap_exit, tilt
        length = slot_value(prog, s0, b);
        const double angle = slot_value(prog, s0 + 1, b);
        const double tilt = slot_value(prog, s0 + 8, b);
        const double g = angle / length;
        double sn, cs, sa, ca;
        sincos(tilt, &sn, &cs);
        sincos(angle, &sa, &ca);
        c[B_L] = length;
        c[B_ANGLE] = angle;
        c[B_COS] = cs;
        c[B_SIN] = sn;
        c[B_G] = g;
        for (int side = 0; side < 2; ++side) {  // dipole.py:338-370
          const double e = slot_value(prog, s0 + 2 + side, b);
          const double fint = slot_value(prog, s0 + 4 + side, b);
          const double h_gap = 0.5 * slot_value(prog, s0 + 6 + side, b);
          const double se = sin(e);
          c[side ? B_HX2 : B_HX1] = g * tan(e);
          c[side ? B_HY2 : B_HY1] =
              -g * tan(e - 2.0 * fint * h_gap * g * (1.0 + se * se) / cos(e));
        }
        c[B_COSA] = ca;
        c[B_SINA] = sa;
        const double sinc_a = angle != 0.0 ? sa / angle : 1.0;
        double sh, chh;
        sincos(0.5 * angle, &sh, &chh);
        const double sinc_h = angle != 0.0 ? sh / (0.5 * angle) : 1.0;
        c[B_LSINC] = length * sinc_a;
        c[B_L2GCOSC] = length * length * g * (-0.5 * sinc_h * sinc_h);
        break;
      }
      case CH_OP_DKD_TDC: {
        // slots: length, voltage, phase, frequency, tilt, mis_x, mis_y
        length = slot_value(prog, s0, b);
        const double tilt = slot_value(prog, s0 + 4, b);
        const double frequency = slot_value(prog, s0 + 3, b);
        double sn, cs;
        sincos(tilt, &sn, &cs);
        c[T_HALF] = 0.5 * length;
        c[T_COS] = cs;
        c[T_SIN] = sn;
        c[T_XOFF] = slot_value(prog, s0 + 5, b);
        c[T_YOFF] = slot_value(prog, s0 + 6, b);
        c[T_V] = slot_value(prog, s0 + 1, b) * -1.0 * q / p0c;
        c[T_KRF] = 2.0 * kPi * frequency / kC;
        c[T_PHASE] = 2.0 * kPi * slot_value(prog, s0 + 2, b);
        break;
      }
      case CH_OP_SECOND_ORDER: {
        // slots: length, k1, k2, angle, e1, e2, fint, fint_exit, gap, tilt, mis_x, mis_y;
        // op_flags bit0: bend (pole faces + rotation, no offsets)
        length = slot_value(prog, s0, b);
        const double k1 = slot_value(prog, s0 + 1, b);
        const double k2 = slot_value(prog, s0 + 2, b);
        const double tilt = slot_value(prog, s0 + 9, b);
        double sn = 0.0, cs = 1.0;
        if (tilt != 0.0) sincos(tilt, &sn, &cs);
        c[S_COS] = cs;
        c[S_SIN] = sn;
        double hx = 0.0;
        if (flags & 1) {
          hx = slot_value(prog, s0 + 3, b) / length;
          const double gap = slot_value(prog, s0 + 8, b);
          edge_kicks(hx, slot_value(prog, s0 + 4, b), slot_value(prog, s0 + 6, b), gap, c[S_KX1],
                     c[S_KY1]);
          edge_kicks(hx, slot_value(prog, s0 + 5, b), slot_value(prog, s0 + 7, b), gap, c[S_KX2],
                     c[S_KY2]);
        } else {
          const double mx = slot_value(prog, s0 + 10, b), my = slot_value(prog, s0 + 11, b);
          c[S_OX] = -mx * cs - my * sn;  // track_methods.py:374-376
          c[S_OY] = mx * sn - my * cs;
          c[S_MX] = mx;
          c[S_MY] = my;
        }
        body_rmatrix(length, k1, hx, beta, igamma2, c + S_R);
        second_order_coefficients(length, k1, k2, hx, beta, igamma2, c + S_T);
        break;
      }
      default:
        break;
    }
    double* dst = base + CH_NL_HEADER + static_cast<int64_t>(tid) * block;
    for (int i = 0; i < block; ++i) dst[i] = static_cast<double>(c[i]);
  }
  lengths[tid] = length;
  __syncthreads();
  if (tid == 0) {
    double total = 0.0;
    for (int i = 0; i < n_ops; ++i) total += lengths[i];
    base[H_P0C] = static_cast<double>(p0c);
    base[H_MC2] = static_cast<double>(mc2);
    base[H_E0] = static_cast<double>(E0);
    base[H_BETA0] = static_cast<double>(p0c / E0);
    base[H_MC2_E0_SQ] = static_cast<double>((mc2 / E0) * (mc2 / E0));
    base[H_LENGTH] = static_cast<double>(total);
    base[H_CHARGE] = static_cast<double>(q);
    base[H_INV_P0C] = 1.0 / p0c;
    base[H_TWO_E0_OVER_P0C] = 2.0 * E0 / p0c;
    for (int i = H_TWO_E0_OVER_P0C + 1; i < CH_NL_HEADER; ++i) base[i] = 0.0;
  }
}

// ---- per-particle maps ------------------------------------------------------------------------
__device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }
__device__ __forceinline__ void sincos_t(float x, float& s, float& c) { sincosf(x, &s, &c); }
__device__ __forceinline__ void sincos_t(double x, double& s, double& c) { sincos(x, &s, &c); }
__device__ __forceinline__ float cosh_t(float x) { return coshf(x); }
__device__ __forceinline__ double cosh_t(double x) { return cosh(x); }
__device__ __forceinline__ float sinh_t(float x) { return sinhf(x); }
__device__ __forceinline__ double sinh_t(double x) { return sinh(x); }

// sqrt(1 + x) - 1 without cancellation (bmadx.py:263-268)
template <typename C>
__device__ __forceinline__ C sqrt_one(C x) {
  return x / (sqrt_t(C(1) + x) + C(1));
}

template <typename C>
struct Beam0 {  // reference-particle constants of one setting
  C p0c, mc2, E0, beta0, mc2_e0_sq, inv_p0c, two_e0_over_p0c;
};

// particle state: transverse coordinates + either (tau, delta) or Bmad-X (z, pz).  In Bmad-X
// mode the quantities that only depend on pz are carried along, so that a run of elements
// evaluates them once per particle: iP = 1 / (1 + pz), rb = beta / beta0 - 1, inv_beta = 1 / beta
// and the delta that corresponds to pz (for the way back).
template <typename C>
struct State {
  C x, px, y, py, l, d;
  C iP, rb, inv_beta, delta;
};

// (tau, delta) -> (z, pz): bmadx.py:7-30 with p^2 - p0c^2 = delta p0c (2 E0 + delta p0c)
template <typename C>
__device__ __forceinline__ void to_bmad(State<C>& s, const Beam0<C>& r) {
  const C energy = r.E0 + s.d * r.p0c;
  const C inv_energy = C(1) / energy;
  const C pz = sqrt_one(s.d * (r.two_e0_over_p0c + s.d));
  const C P = C(1) + pz;
  const C m_over_e = r.mc2 * inv_energy;
  s.iP = C(1) / P;
  s.rb = sqrt_one(m_over_e * m_over_e * pz * (C(2) + pz));  // (beta/beta0)^2 - 1 -> beta/beta0 - 1
  s.inv_beta = energy * s.iP * r.inv_p0c;
  s.delta = s.d;
  s.l = -(P * r.p0c * inv_energy) * s.l;
  s.d = pz;
}

// refresh the pz-dependent quantities after an element that changed pz (TDC)
template <typename C>
__device__ __forceinline__ void refresh_from_pz(State<C>& s, const Beam0<C>& r) {
  const C P = C(1) + s.d;
  const C p = P * r.p0c;
  const C energy = sqrt_t(p * p + r.mc2 * r.mc2);
  const C m_over_e = r.mc2 / energy;
  s.iP = C(1) / P;
  s.rb = sqrt_one(m_over_e * m_over_e * s.d * (C(2) + s.d));
  s.inv_beta = energy / p;
  s.delta = r.p0c * s.d * (C(2) + s.d) / (energy + r.E0);  // E^2 - E0^2 = p0c^2 pz (2 + pz)
}

// (z, pz) -> (tau, delta): bmadx.py:33-55
template <typename C>
__device__ __forceinline__ void from_bmad(State<C>& s) {
  s.l = -s.l * s.inv_beta;
  s.d = s.delta;
}

template <typename C>
__device__ __forceinline__ void offset_set(State<C>& s, C cs, C sn, C x_off, C y_off) {
  const C xi = s.x - x_off, yi = s.y - y_off;  // bmadx.py:115-146
  const C px = s.px, py = s.py;
  s.x = xi * cs + yi * sn;
  s.y = -xi * sn + yi * cs;
  s.px = px * cs + py * sn;
  s.py = -px * sn + py * cs;
}

template <typename C>
__device__ __forceinline__ void offset_unset(State<C>& s, C cs, C sn, C x_off, C y_off) {
  const C x = s.x, y = s.y, px = s.px, py = s.py;  // bmadx.py:149-180
  s.x = x * cs - y * sn + x_off;
  s.y = x * sn + y * cs + y_off;
  s.px = px * cs - py * sn;
  s.py = px * sn + py * cs;
}

// exact drift (bmadx.py:271-302) with the cached 1 / P and beta / beta0 - 1:
//   dz = L (sqrt_one((beta/beta0)^2 - 1) + sqrt_one(-Pxy2) / Pl),  sqrt_one(-Pxy2) = -Pxy2 / (Pl + 1)
template <typename C>
__device__ __forceinline__ void track_a_drift(State<C>& s, C L) {
  const C Px = s.px * s.iP, Py = s.py * s.iP;
  const C Pxy2 = Px * Px + Py * Py;
  const C Pl = sqrt_t(C(1) - Pxy2);
  const C t = C(1) / (Pl * (Pl + C(1)));
  const C L_over_Pl = L * t * (Pl + C(1));
  s.x += L_over_Pl * Px;
  s.y += L_over_Pl * Py;
  s.l += L * (s.rb - Pxy2 * t);
}

// bmadx.py:183-220; the high-|pz| branch ds (beta - beta0) / beta0 is ds * rb, which does not
// cancel
template <typename C>
__device__ __forceinline__ C low_energy_z_correction(const State<C>& s, C ds, const Beam0<C>& r) {
  const C pz = s.d;
  const C evaluation = r.mc2 * (r.beta0 * pz) * (r.beta0 * pz);
  const C b2 = r.beta0 * r.beta0;
  // both forms are a handful of multiply-adds: evaluated unconditionally and selected, so the
  // warp does not diverge on a per-particle comparison
  const C low = ds * pz * (C(1) - C(3) * (pz * b2) / C(2) +
                           pz * pz * b2 * (C(2) * b2 - r.mc2_e0_sq / C(2))) * r.mc2_e0_sq;
  return evaluation < C(3e-7) * r.E0 ? low : ds * s.rb;
}

// one plane of bmadx.py:223-260 for the argument `k1` (kx^2 = -k1), step length l
template <typename C>
struct QuadPlane {
  C a11, a12, a21, c1, c2, c3;
};
// cos / cosh and sin / sinh of k l as ONE pair of power series in w = k1 l^2 (either sign):
//   cx = sum w^n / (2n)!      = cos(sqrt(-w)) for w < 0, cosh(sqrt(w)) for w > 0
//   sx = l sum w^n / (2n+1)!  = sin(k l) / k resp. sinh(k l) / k
// For |w| <= 2.25 (|k l| <= 1.5, i.e. every realistic integration step) 8 terms are exact to
// 3e-11 and 12 terms to 3e-20, with no branch on the sign, no square root, no division by k and no
// trigonometric range reduction; per particle because k1 carries the particle's 1 / (1 + pz).
// (sqrt + sincosf + coshf + sinhf + 2 divisions per plane were ~45 % of a drift_kick_drift
// quadrupole.)  Larger |w| takes the closed forms.
__device__ constexpr double kCosSeries[12] = {
    1.0 / 1.0, 1.0 / 2.0, 1.0 / 24.0, 1.0 / 720.0, 1.0 / 40320.0, 1.0 / 3628800.0,
    1.0 / 479001600.0, 1.0 / 87178291200.0, 1.0 / 20922789888000.0, 1.0 / 6402373705728000.0,
    1.0 / 2432902008176640000.0, 1.0 / 1124000727777607680000.0};
__device__ constexpr double kSinSeries[12] = {
    1.0 / 1.0, 1.0 / 6.0, 1.0 / 120.0, 1.0 / 5040.0, 1.0 / 362880.0, 1.0 / 39916800.0,
    1.0 / 6227020800.0, 1.0 / 1307674368000.0, 1.0 / 355687428096000.0,
    1.0 / 121645100408832000.0, 1.0 / 51090942171709440000.0, 1.0 / 25852016738884976640000.0};
template <typename C>
__device__ __forceinline__ void series_cos_sin(C w, C& c, C& s) {
  constexpr int N = sizeof(C) == 4 ? 8 : 12;
  c = static_cast<C>(kCosSeries[N - 1]);
  s = static_cast<C>(kSinSeries[N - 1]);
#pragma unroll
  for (int n = N - 2; n >= 0; --n) {
    c = c * w + static_cast<C>(kCosSeries[n]);
    s = s * w + static_cast<C>(kSinSeries[n]);
  }
}

template <typename C>
__device__ __forceinline__ QuadPlane<C> quadrupole_plane(C k1, C l, C rel_p, C inv_rel_p) {
  C cx, sx;
  const C w = k1 * l * l;
  // the series is evaluated unconditionally; one warp-uniform test (every lane reaches this
  // point) skips the closed forms when all 32 particles are inside its range -- the normal case
  const bool in_range = w <= C(2.25) && w >= C(-2.25);
  series_cos_sin(w, cx, sx);
  sx *= l;
  if (!__all_sync(0xffffffffu, in_range)) {  // rare: some lane needs a closed form
    if (!in_range) {
      const C k = sqrt_t(k1 < C(0) ? -k1 : k1);
      if (k1 < C(0)) {  // kx real: focusing
        C sn;
        sincos_t(k * l, sn, cx);
        sx = sn / k;
      } else {
        cx = cosh_t(k * l);
        sx = sinh_t(k * l) / k;
      }
    }
  }
  QuadPlane<C> q;
  q.a11 = cx;
  q.a12 = sx * inv_rel_p;
  q.a21 = k1 * sx * rel_p;
  q.c1 = k1 * (-cx * sx + l) * C(0.25);
  q.c2 = -k1 * sx * sx * C(0.5) * inv_rel_p;
  q.c3 = -(cx * sx + l) * C(0.25) * inv_rel_p * inv_rel_p;
  return q;
}

// quadrupole.py:168-251; the per-step coefficients only depend on pz, which is constant
template <typename C>
__device__ __forceinline__ void track_quadrupole(State<C>& s, const C* c, int num_steps,
                                                 const Beam0<C>& r) {
  offset_set(s, c[Q_COS], c[Q_SIN], c[Q_XOFF], c[Q_YOFF]);
  const C rel_p = C(1) + s.d;
  const C k1 = c[Q_K1] * s.iP;  // b1 / (L rel_p)
  const QuadPlane<C> tx = quadrupole_plane(-k1, c[Q_STEP], rel_p, s.iP);
  const QuadPlane<C> ty = quadrupole_plane(k1, c[Q_STEP], rel_p, s.iP);
  const C dz_low = low_energy_z_correction(s, c[Q_STEP], r);
  for (int step = 0; step < num_steps; ++step) {
    // c1 u^2 + c2 u pu + c3 pu^2 per plane as u (c1 u + c2 pu) + (c3 pu) pu: 5 instead of 8 ops
    s.l += s.x * (tx.c1 * s.x + tx.c2 * s.px) + (tx.c3 * s.px) * s.px +
           s.y * (ty.c1 * s.y + ty.c2 * s.py) + (ty.c3 * s.py) * s.py;
    const C x = s.x, y = s.y;
    s.x = tx.a11 * x + tx.a12 * s.px;
    s.px = tx.a21 * x + tx.a11 * s.px;
    s.y = ty.a11 * y + ty.a12 * s.py;
    s.py = ty.a21 * y + ty.a11 * s.py;
    s.l += dz_low;
  }
  offset_unset(s, c[Q_COS], c[Q_SIN], c[Q_XOFF], c[Q_YOFF]);
}

__device__ __forceinline__ double sinc_d(double x) { return x != 0.0 ? sin(x) / x : 1.0; }

// dipole.py:183-370 in fp64 (entrance fringe, sector body, exit fringe in the tilted frame)
template <typename C>
__device__ __forceinline__ void track_dipole(State<C>& st, const double* cc, int flags,
                                             const Beam0<double>& rr) {
  const double cs = cc[B_COS], sn = cc[B_SIN];
  State<double> s{st.x, st.px, st.y, st.py, st.l, st.d};
  offset_set(s, cs, sn, 0.0, 0.0);
  if (flags & 1) {
    s.px += s.x * cc[B_HX1];
    s.py += s.y * cc[B_HY1];
  }
  {  // _bmadx_body, dipole.py:244-336
    const double L = cc[B_L], angle = cc[B_ANGLE], g = cc[B_G];
    const double cos_a = cc[B_COSA], sin_a = cc[B_SINA], l_sinc = cc[B_LSINC];
    const double l2gcosc = cc[B_L2GCOSC];
    const double x = s.x, pz = s.d;
    const double px_norm = sqrt((1.0 + pz) * (1.0 + pz) - s.py * s.py);
    const double phi1 = asin(s.px / px_norm);
    const double gp = g / px_norm;
    double sin_ap, cos_ap;
    sincos(angle + phi1, &sin_ap, &cos_ap);
    const double gx1 = 1.0 + g * x;
    const double alpha = 2.0 * gx1 * sin_ap * l_sinc - gp * (gx1 * l_sinc) * (gx1 * l_sinc);
    const double x2_t1 = x * cos_a + l2gcosc;
    const double x2_t2 = sqrt(cos_ap * cos_ap + gp * alpha);
    const double ga = gp * alpha;
    // Lcu = x2 - x2_t1 is the correction term itself (no subtraction of nearly equal numbers)
    double Lcu;
    if (fabs(angle + phi1) < 0.5 * kPi) {
      Lcu = alpha / (x2_t2 + cos_ap);
    } else {
      Lcu = alpha * (ga != 0.0 ? (sqrt(cos_ap * cos_ap + ga) - cos_ap) / ga : 1.0 / (2.0 * cos_ap));
    }
    const double x2 = x2_t1 + Lcu;
    const double Lcv = -l_sinc - x * sin_a;
    const double theta_p = 2.0 * (angle + phi1 - 0.5 * kPi - atan2(Lcv, Lcu));
    const double Lc = sqrt(Lcu * Lcu + Lcv * Lcv);
    const double Lp = Lc / sinc_d(0.5 * theta_p);
    const double mc2 = rr.mc2, p0c = rr.p0c;
    const double P = p0c * (1.0 + pz);
    const double E = sqrt(P * P + mc2 * mc2);
    const double beta = P / E, beta0 = rr.beta0;  // beta0 = p0c / E0 from the constants header
    s.x = x2;
    s.px = px_norm * sin(angle + phi1 - theta_p);
    s.y = s.y + s.py * Lp / px_norm;
    s.l = s.l + (beta * L / beta0) - ((1.0 + pz) * Lp / px_norm);
  }
  if (flags & 2) {
    s.px += s.x * cc[B_HX2];
    s.py += s.y * cc[B_HY2];
  }
  offset_unset(s, cs, sn, 0.0, 0.0);
  st.x = static_cast<C>(s.x);
  st.px = static_cast<C>(s.px);
  st.y = static_cast<C>(s.y);
  st.py = static_cast<C>(s.py);
  st.l = static_cast<C>(s.l);
}

// transverse_deflecting_cavity.py:122-209 in fp64
template <typename C>
__device__ __forceinline__ void track_tdc(State<C>& st, const C* c, const double* cc,
                                          const Beam0<C>& rc, const Beam0<double>& r) {
  // frame change and the two half drifts in the beam dtype, like a Drift op (the cached
  // pz-dependent quantities are valid on entry and refreshed after the kick); only the kick,
  // whose energy change is a 1e-5 correction on E, is evaluated in fp64
  offset_set(st, c[T_COS], c[T_SIN], c[T_XOFF], c[T_YOFF]);
  track_a_drift(st, c[T_HALF]);
  {
    const double x = st.x, z = st.l, pz = st.d;
    const double voltage = cc[T_V], k_rf = cc[T_KRF];
    const double pc_old = (1.0 + pz) * r.p0c;
    const double E_old = sqrt(pc_old * pc_old + r.mc2 * r.mc2);
    const double beta_old = pc_old / E_old;
    // phase = 2 pi (phase0 - t f) with t = -z / (beta c)
    const double phase = cc[T_PHASE] + k_rf * z / beta_old;
    double sp, cp;
    sincos(phase, &sp, &cp);
    const double E_new = E_old + voltage * cp * k_rf * x * r.p0c;
    const double pc = sqrt(E_new * E_new - r.mc2 * r.mc2);
    const double beta = pc / E_new;
    // pz = (pc - p0c) / p0c with pc^2 - p0c^2 = (E_new - E0)(E_new + E0)
    st.px = static_cast<C>(static_cast<double>(st.px) + voltage * sp);
    st.d = static_cast<C>((E_new - r.E0) * (E_new + r.E0) / (r.p0c * (pc + r.p0c)));
    st.l = static_cast<C>(z * beta / beta_old);
  }
  refresh_from_pz(st, rc);
  track_a_drift(st, c[T_HALF]);
  offset_unset(st, c[T_COS], c[T_SIN], c[T_XOFF], c[T_YOFF]);
}

// The two fp64 bodies are several hundred instructions each: called (not inlined) once per
// particle, so that the P particles of a thread do not multiply the code size and the register
// pressure of the FP64_OPS instantiation.
template <typename C>
__device__ __noinline__ void track_dipole_call(State<C>& st, const double* cc, int flags,
                                               const Beam0<double>& r) {
  track_dipole(st, cc, flags, r);
}
template <typename C>
__device__ __noinline__ void track_tdc_call(State<C>& st, const C* c, const double* cc,
                                            const Beam0<C>& rc, const Beam0<double>& r) {
  track_tdc(st, c, cc, rc, r);
}

// Elements [A, B) of a 16-byte aligned coefficient block through 128-bit shared-memory loads (all
// lanes read the same address: one broadcast wavefront per load).  Scalar loads of the 58
// second-order coefficients were 31 % of the executed instructions and 55 % of the stall samples
// of a second_order run.
template <typename C>
struct Wide;
template <>
struct Wide<float> {
  using type = float4;
  static constexpr int lanes = 4;
};
template <>
struct Wide<double> {
  using type = double2;
  static constexpr int lanes = 2;
};
template <int A, int B, typename C>
__device__ __forceinline__ void load_span(C (&dst)[B - A], const C* c) {
  using V = typename Wide<C>::type;
  constexpr int L = Wide<C>::lanes;
#pragma unroll
  for (int w = A / L; w <= (B - 1) / L; ++w) {
    const V v = reinterpret_cast<const V*>(c)[w];
    C e[4];
    if constexpr (L == 4) {
      e[0] = v.x, e[1] = v.y, e[2] = v.z, e[3] = v.w;
    } else {
      e[0] = v.x, e[1] = v.y, e[2] = C(0), e[3] = C(0);
    }
#pragma unroll
    for (int i = 0; i < L; ++i) {
      const int index = w * L + i;
      if (index >= A && index < B) dst[index - A] = e[i];
    }
  }
}

// element.py:195-225 with the frame changes of quadrupole.py:136-143 / dipole.py:417-426 applied
// to the particle instead of being folded into T: q = entry(p), t = R q + T q q, out = exit(t).
// The thread's P particles go through the element together, so every coefficient row is loaded
// once (T rows: 0..8, 9..17, 18..23, 24..29, 30..38 of the enum above).
template <typename C, int P>
__device__ __forceinline__ void track_second_order(State<C> (&s)[P], const C* c) {
  C h[S_T];  // frame, kicks and R
  load_span<0, S_T>(h, c);
  const C cs = h[S_COS], sn = h[S_SIN];
  const C* r = h + S_R;
  C x[P], y[P], px[P], py[P], dl[P];
  C xx[P], xp[P], pp[P], xd[P], pd[P], dd[P], yy[P], yq[P], qq[P];
#pragma unroll
  for (int k = 0; k < P; ++k) {
    x[k] = s[k].x * cs + s[k].y * sn + h[S_OX];
    y[k] = -s[k].x * sn + s[k].y * cs + h[S_OY];
    px[k] = s[k].px * cs + s[k].py * sn;
    py[k] = -s[k].px * sn + s[k].py * cs;
    px[k] += h[S_KX1] * x[k];
    py[k] += h[S_KY1] * y[k];
    dl[k] = s[k].d;
    xx[k] = x[k] * x[k], xp[k] = x[k] * px[k], pp[k] = px[k] * px[k];
    xd[k] = x[k] * dl[k], pd[k] = px[k] * dl[k], dd[k] = dl[k] * dl[k];
    yy[k] = y[k] * y[k], yq[k] = y[k] * py[k], qq[k] = py[k] * py[k];
  }
  // the three rows with the same nine products: x, px, tau
  C ox[P], opx[P];
  {
    C t[9];
    load_span<S_T + T000, S_T + T000 + 9>(t, c);
#pragma unroll
    for (int k = 0; k < P; ++k)
      ox[k] = r[R_CX] * x[k] + r[R_SX] * px[k] + r[R_05] * dl[k] + t[0] * xx[k] + t[1] * xp[k] +
              t[2] * pp[k] + t[3] * xd[k] + t[4] * pd[k] + t[5] * dd[k] + t[6] * yy[k] +
              t[7] * yq[k] + t[8] * qq[k];
  }
  {
    C t[9];
    load_span<S_T + T100, S_T + T100 + 9>(t, c);
#pragma unroll
    for (int k = 0; k < P; ++k)
      opx[k] = r[R_10] * x[k] + r[R_CX] * px[k] + r[R_15] * dl[k] + t[0] * xx[k] + t[1] * xp[k] +
               t[2] * pp[k] + t[3] * xd[k] + t[4] * pd[k] + t[5] * dd[k] + t[6] * yy[k] +
               t[7] * yq[k] + t[8] * qq[k];
  }
  {
    C t[9];
    load_span<S_T + T400, S_T + T400 + 9>(t, c);
#pragma unroll
    for (int k = 0; k < P; ++k)
      s[k].l = s[k].l + r[R_15] * x[k] + r[R_05] * px[k] + r[R_56] * dl[k] + t[0] * xx[k] +
               t[1] * xp[k] + t[2] * pp[k] + t[3] * xd[k] + t[4] * pd[k] + t[5] * dd[k] +
               t[6] * yy[k] + t[7] * yq[k] + t[8] * qq[k];
  }
  // the two vertical rows share six mixed products
  C t2[6], t3[6];
  load_span<S_T + T202, S_T + T202 + 6>(t2, c);
  load_span<S_T + T302, S_T + T302 + 6>(t3, c);
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const C xy = x[k] * y[k], xq = x[k] * py[k], py_ = px[k] * y[k], pq = px[k] * py[k];
    const C yd = y[k] * dl[k], qd = py[k] * dl[k];
    const C oy = r[R_CY] * y[k] + r[R_SY] * py[k] + t2[0] * xy + t2[1] * xq + t2[2] * py_ +
                 t2[3] * pq + t2[4] * yd + t2[5] * qd;
    C opy = r[R_32] * y[k] + r[R_CY] * py[k] + t3[0] * xy + t3[1] * xq + t3[2] * py_ +
            t3[3] * pq + t3[4] * yd + t3[5] * qd;
    const C opx_k = opx[k] + h[S_KX2] * ox[k];
    opy += h[S_KY2] * oy;
    s[k].x = ox[k] * cs - oy * sn + h[S_MX];
    s[k].y = ox[k] * sn + oy * cs + h[S_MY];
    s[k].px = opx_k * cs - opy * sn;
    s[k].py = opx_k * sn + opy * cs;
  }
}

template <typename T>
struct TrackArgs {
  const T* particles_in;
  const double* constants;
  T* particles_out;
  const int32_t* particle_index;
  const int32_t* constants_index;
  const int32_t* opcodes;   // device, already offset to the first op of the run
  const int32_t* op_flags;
  int64_t particle_stride;
  int64_t constants_stride;
  int64_t n_particles;
  int64_t n_settings;
  int32_t n_ops;
  int32_t block;
  int32_t bulk_in;
  int32_t bulk_out;
};

// FP64_OPS: the run holds a bend body or a TDC kick (evaluated in fp64); runs without them use a
// leaner instantiation with more resident CTAs per SM (more tiles in flight).
template <typename T, int P, int THREADS, bool FP64_OPS>
__global__ void __launch_bounds__(THREADS, (FP64_OPS || sizeof(T) == 8) ? 1 : 6)
nonlinear_track_kernel(const TrackArgs<T> a) {
  constexpr int TP = P * THREADS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // stage: outgoing rows (read by the bulk store); in0 / in1: double-buffered incoming tiles, so
  // that the tile of the NEXT setting is in flight while this one is computed
  T* stage = reinterpret_cast<T*>(smem_raw);
  T* in0 = stage + TP * 7;
  T* in1 = in0 + TP * 7;
  double* consts64 = reinterpret_cast<double*>(in1 + TP * 7);
  const int n_consts = CH_NL_HEADER + a.n_ops * a.block;
  T* consts = reinterpret_cast<T*>(consts64 + n_consts);  // the same, rounded to the beam dtype
  uint64_t* bar = reinterpret_cast<uint64_t*>(consts + (n_consts + 3) / 4 * 4);  // bar[0..1]
  int32_t* codes = reinterpret_cast<int32_t*>(bar + 2);
  int32_t* flags = codes + a.n_ops;

  const int tid = threadIdx.x;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * TP;
  const int count = static_cast<int>(min(static_cast<int64_t>(TP), a.n_particles - n0));

  if (a.bulk_in && tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    fence_mbar_init();
  }
  for (int i = tid; i < a.n_ops; i += THREADS) {
    codes[i] = a.opcodes[i];
    flags[i] = a.op_flags[i];
  }
  __syncthreads();

  auto tile_offset = [&](int64_t b) {
    return (a.particle_index ? a.particle_index[b] : b) * a.particle_stride + n0 * 7;
  };
  // one elected thread asks the TMA engine for the tile at `offset` (bulk path only)
  auto request = [&](int buf, int64_t offset) {
    if (tid == 0) {
      const uint32_t bytes = static_cast<uint32_t>(count) * 7u * sizeof(T);
      mbar_expect_tx(bar + buf, bytes);
      bulk_load(buf ? in1 : in0, a.particles_in + offset, bytes, bar + buf);
    }
  };

  T p[P][7];
  int64_t loaded = -1;     // offset of the tile held in registers
  int64_t requested = -1;  // offset of the tile most recently requested (into in[buf])
  uint32_t phase[2] = {0u, 0u};
  int buf = 0;             // buffer the next needed tile is (being) loaded into
  if (a.bulk_in && blockIdx.y < a.n_settings) {
    requested = tile_offset(blockIdx.y);
    request(0, requested);
  }
  for (int64_t b = blockIdx.y; b < a.n_settings; b += gridDim.y) {
    // the staging tile may still be read by the bulk store of the previous setting
    if (a.bulk_out && tid == 0) bulk_wait_read<0>();
    __syncthreads();
    // (fetching the NEXT setting's table with a bulk copy into a second buffer, off the per-tile
    // critical path, was measured and is not faster: 0.71 vs 0.69 ms for the single Drift)
    const double* src_c =
        a.constants + (a.constants_index ? a.constants_index[b] : b) * a.constants_stride;
    for (int i = tid; i < n_consts; i += THREADS) {
      const double v = src_c[i];
      consts64[i] = v;
      consts[i] = static_cast<T>(v);
    }
    const int64_t p_off = tile_offset(b);
    if (p_off != loaded) {
      const T* tile;
      if (a.bulk_in) {
        // normally the tile was prefetched during the previous iteration; if not (the previous
        // settings shared one tile), ask for it now.  Then prefetch the following tile into the
        // other buffer (every thread finished reading it before the barrier at the top of this
        // iteration).
        if (requested != p_off) {
          requested = p_off;
          request(buf, p_off);
        }
        const int64_t b_next = b + gridDim.y;
        const int cur = buf;
        if (b_next < a.n_settings) {
          const int64_t next_off = tile_offset(b_next);
          if (next_off != p_off) {
            requested = next_off;
            buf ^= 1;
            request(buf, next_off);
          }
        }
        mbar_wait(bar + cur, phase[cur]);
        phase[cur] ^= 1u;
        tile = cur ? in1 : in0;
      } else {
        const T* src = a.particles_in + p_off;
        for (int i = tid; i < count * 7; i += THREADS) in0[i] = src[i];
        __syncthreads();
        tile = in0;
      }
#pragma unroll
      for (int k = 0; k < P; ++k) {
        const int local = tid + k * THREADS;
#pragma unroll
        for (int j = 0; j < 7; ++j) p[k][j] = local < count ? tile[local * 7 + j] : T(0);
      }
      loaded = p_off;
    }
    __syncthreads();  // constants visible, tile consumed

    const Beam0<T> ref{consts[H_P0C], consts[H_MC2], consts[H_E0], consts[H_BETA0],
                       consts[H_MC2_E0_SQ], consts[H_INV_P0C], consts[H_TWO_E0_OVER_P0C]};
    const Beam0<double> ref64{consts64[H_P0C], consts64[H_MC2], consts64[H_E0],
                              consts64[H_BETA0], consts64[H_MC2_E0_SQ], consts64[H_INV_P0C],
                              consts64[H_TWO_E0_OVER_P0C]};
    // ops outer, the thread's P particles inner: opcode dispatch, flags and coefficient loads are
    // paid once per op instead of once per (op, particle)
    State<T> s[P];
#pragma unroll
    for (int k = 0; k < P; ++k) {
      s[k].x = p[k][0], s[k].px = p[k][1], s[k].y = p[k][2], s[k].py = p[k][3];
      s[k].l = p[k][4], s[k].d = p[k][5];
      s[k].iP = s[k].rb = s[k].inv_beta = s[k].delta = T(0);
    }
    bool bmad = false;  // representation of (s.l, s.d): uniform over the CTA
    for (int op = 0; op < a.n_ops; ++op) {
      const int code = codes[op];
      const T* c = consts + CH_NL_HEADER + op * a.block;
      const double* c64 = consts64 + CH_NL_HEADER + op * a.block;
      if (code == CH_OP_IDENTITY) continue;
      const bool wants_bmad = code != CH_OP_SECOND_ORDER;
      if (wants_bmad && !bmad) {
#pragma unroll
        for (int k = 0; k < P; ++k) to_bmad(s[k], ref);
      }
      if (!wants_bmad && bmad) {
#pragma unroll
        for (int k = 0; k < P; ++k) from_bmad(s[k]);
      }
      bmad = wants_bmad;
      switch (code) {
        case CH_OP_DKD_DRIFT: {
          const T length = c[D_L];
#pragma unroll
          for (int k = 0; k < P; ++k) track_a_drift(s[k], length);
          break;
        }
        case CH_OP_DKD_QUADRUPOLE: {
          const int steps = flags[op] > 0 ? flags[op] : 1;
#pragma unroll
          for (int k = 0; k < P; ++k) track_quadrupole(s[k], c, steps, ref);
          break;
        }
        case CH_OP_DKD_DIPOLE:
          if constexpr (FP64_OPS) {
#pragma unroll
            for (int k = 0; k < P; ++k) track_dipole_call(s[k], c64, flags[op], ref64);
          }
          break;
        case CH_OP_DKD_TDC:
          if constexpr (FP64_OPS) {
#pragma unroll
            for (int k = 0; k < P; ++k) track_tdc_call(s[k], c, c64, ref, ref64);
          }
          break;
        case CH_OP_SECOND_ORDER:
          track_second_order(s, c);
          break;
        default:
          break;
      }
    }
#pragma unroll
    for (int k = 0; k < P; ++k) {
      if (bmad) from_bmad(s[k]);
      T* row = stage + (tid + k * THREADS) * 7;
      row[0] = s[k].x;
      row[1] = s[k].px;
      row[2] = s[k].y;
      row[3] = s[k].py;
      row[4] = s[k].l;
      row[5] = s[k].d;
      row[6] = p[k][6];
    }

    T* out = a.particles_out + (b * a.n_particles + n0) * 7;
    if (a.bulk_out) {
      fence_async_shared();
      __syncthreads();
      if (tid == 0) {
        bulk_store(out, stage, static_cast<uint32_t>(count) * 7u * sizeof(T));
        bulk_commit();
      }
    } else {
      __syncthreads();
      for (int i = tid; i < count * 7; i += THREADS) out[i] = stage[i];
    }
  }
  if (a.bulk_out && tid == 0) bulk_wait<0>();
}

int block_size(const ch_program* program, int32_t op_begin, int32_t op_end) {
  for (int32_t i = op_begin; i < op_end; ++i)
    if (program->opcodes_host[i] == CH_OP_SECOND_ORDER) return CH_NL_BLOCK_SECOND_ORDER;
  return CH_NL_BLOCK_DKD;
}

int check_run(const ch_program* program, int32_t op_begin, int32_t op_end, const char* who) {
  CH_REQUIRE(program != nullptr, "%s: program is NULL", who);
  CH_REQUIRE(op_begin >= 0 && op_begin < op_end && op_end <= program->n_ops,
             "%s: op range [%d, %d) outside program of %d ops", who, op_begin, op_end,
             program->n_ops);
  CH_REQUIRE(op_end - op_begin <= CH_NL_MAX_OPS, "%s: a run holds at most %d ops, got %d", who,
             CH_NL_MAX_OPS, op_end - op_begin);
  for (int32_t i = op_begin; i < op_end; ++i) {
    const int32_t code = program->opcodes_host[i];
    CH_REQUIRE(code == CH_OP_IDENTITY || (code >= CH_OP_DKD_DRIFT && code <= CH_OP_SECOND_ORDER),
               "%s: op %d (opcode %d) is not a non-linear tracking op", who, i, code);
  }
  return CH_OK;
}

template <typename T>
int launch_track(const ch_program* program, int32_t op_begin, int32_t op_end, const double* constants,
                 int64_t constants_stride, const int32_t* constants_index,
                 const void* particles_in, int64_t particle_stride, const int32_t* particle_index,
                 int64_t n_particles, int64_t n_settings, void* particles_out,
                 cudaStream_t stream) {
  constexpr int P = sizeof(T) == 4 ? 2 : 1;
  constexpr int THREADS = 128;
  constexpr int TP = P * THREADS;
  TrackArgs<T> a;
  a.particles_in = static_cast<const T*>(particles_in);
  a.constants = static_cast<const double*>(constants);
  a.particles_out = static_cast<T*>(particles_out);
  a.particle_index = particle_index;
  a.constants_index = constants_index;
  a.opcodes = program->opcodes + op_begin;
  a.op_flags = program->op_flags + op_begin;
  a.particle_stride = particle_stride;
  a.constants_stride = constants_stride;
  a.n_particles = n_particles;
  a.n_settings = n_settings;
  a.n_ops = op_end - op_begin;
  a.block = block_size(program, op_begin, op_end);
  a.bulk_in = bulk_compatible<T>(particles_in, n_particles, particle_stride) ? 1 : 0;
  a.bulk_out = bulk_compatible<T>(particles_out, n_particles, n_particles * 7) ? 1 : 0;
  const int n_consts = CH_NL_HEADER + a.n_ops * a.block;
  const size_t smem = sizeof(T) * (3 * TP * 7 + (n_consts + 3) / 4 * 4) + sizeof(double) * n_consts +
                      2 * sizeof(uint64_t) + 2 * sizeof(int32_t) * a.n_ops;
  const int64_t tiles = (n_particles + TP - 1) / TP;
  CH_REQUIRE(tiles <= 2147483647LL, "ch_track_nonlinear: too many particles");
  // enough CTAs for ~8 waves of 148 SMs x 6 resident CTAs; beyond that a CTA loops over settings
  // (the beam tile stays in registers when it is shared, and is prefetched when it is not)
  int64_t rows = (148 * 6 * 8 + tiles - 1) / tiles;
  rows = rows < 1 ? 1 : (rows > n_settings ? n_settings : rows);
  dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(rows < 65535 ? rows : 65535));
  bool fp64_ops = false;
  for (int32_t i = op_begin; i < op_end; ++i)
    fp64_ops |= program->opcodes_host[i] == CH_OP_DKD_DIPOLE ||
                program->opcodes_host[i] == CH_OP_DKD_TDC;
  auto launch = [&](auto kernel) -> int {
    CH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
    kernel<<<grid, THREADS, smem, stream>>>(a);
    return CH_OK;
  };
  const int status = fp64_ops ? launch(nonlinear_track_kernel<T, P, THREADS, true>)
                              : launch(nonlinear_track_kernel<T, P, THREADS, false>);
  if (status != CH_OK) return status;
  CH_LAUNCH_CHECK();
  return CH_OK;
}

}  // namespace
}  // namespace ch

extern "C" int64_t ch_nonlinear_constants_len(const ch_program* program, int32_t op_begin,
                                              int32_t op_end) {
  if (ch::check_run(program, op_begin, op_end, "ch_nonlinear_constants_len") != CH_OK) return -1;
  return CH_NL_HEADER +
         static_cast<int64_t>(op_end - op_begin) * ch::block_size(program, op_begin, op_end);
}

extern "C" int ch_nonlinear_constants(const ch_program* program, int32_t op_begin, int32_t op_end,
                                      int64_t n_settings, const void* energy,
                                      int64_t energy_stride, int32_t energy_dtype,
                                      const void* mass_eV, int32_t mass_dtype,
                                      const void* num_elementary_charges, int32_t charge_dtype,
                                      double* constants, void* stream) {
  const int status = ch::check_run(program, op_begin, op_end, "ch_nonlinear_constants");
  if (status != CH_OK) return status;
  CH_REQUIRE(n_settings > 0 && n_settings <= 2147483647LL,
             "ch_nonlinear_constants: n_settings must be in [1, 2^31)");
  CH_REQUIRE(energy && mass_eV && constants, "ch_nonlinear_constants: NULL pointer argument");
  ch::Program prog{program->opcodes, program->op_flags, program->slot_begin, program->slots};
  ch::ScalarRef e{energy, energy_stride, energy_dtype};
  ch::ScalarRef m{mass_eV, 0, mass_dtype};
  ch::ScalarRef q{num_elementary_charges, 0, charge_dtype};
  const int32_t block = ch::block_size(program, op_begin, op_end);
  const unsigned blocks = static_cast<unsigned>(n_settings);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  ch::nonlinear_constants_kernel<<<blocks, 64, 0, s>>>(prog, op_begin, op_end - op_begin, e, m, q,
                                                       block, constants);
  CH_LAUNCH_CHECK();
  return CH_OK;
}

extern "C" int ch_track_nonlinear(const ch_program* program, int32_t op_begin, int32_t op_end,
                                  const double* constants, int64_t constants_stride,
                                  const int32_t* constants_index, const void* particles_in,
                                  int64_t particle_stride, const int32_t* particle_index,
                                  int64_t n_particles, int64_t n_settings, void* particles_out,
                                  int32_t dtype, void* stream) {
  const int status = ch::check_run(program, op_begin, op_end, "ch_track_nonlinear");
  if (status != CH_OK) return status;
  CH_REQUIRE(constants && particles_in && particles_out,
             "ch_track_nonlinear: NULL pointer argument");
  CH_REQUIRE(n_particles > 0 && n_settings > 0, "ch_track_nonlinear: empty beam or batch");
  CH_REQUIRE(dtype == CH_F32 || dtype == CH_F64, "ch_track_nonlinear: bad dtype %d", dtype);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == CH_F32)
    return ch::launch_track<float>(program, op_begin, op_end, constants, constants_stride,
                                   constants_index, particles_in, particle_stride, particle_index,
                                   n_particles, n_settings, particles_out, s);
  return ch::launch_track<double>(program, op_begin, op_end, constants, constants_stride,
                                  constants_index, particles_in, particle_stride, particle_index,
                                  n_particles, n_settings, particles_out, s);
}

// Map composition: builds every first-order map of a lattice section from the vectorised
// element parameters and multiplies them into the cumulative maps ch_apply_maps consumes.
//
// One CTA (256 threads) per lattice setting; per chunk of <= 256 elements:
//   A (one element per thread): read the element's parameters through the slot table and
//     reduce them, in fp64 with closed real forms, to <= 16 coefficients (cos/sin-like terms,
//     r56, edge kicks, tilt rotation, misalignment shifts).  All memory latency and all
//     transcendental math lives here, fully parallel.
//   B1 (32 groups of 8 lanes, 8 consecutive elements each): left-multiplying by an element map
//     acts on the 7 columns of a map independently, so lane j of a group owns column j of the
//     group's product (6 fp64 registers) and applies each element as 3-20 FMAs with
//     coefficients broadcast from shared memory; lane 7 sums the lengths.  All groups run at
//     once: 8 dependent steps instead of 256.
//   B2 (7 lanes): the exclusive prefix over the <= 32 group products, again column-wise (36
//     FMAs per lane and group, no communication between lanes).
//   B3: cut points (apertures, an active cavity) recorded in B1 relative to their group are
//     multiplied by the prefix of the groups before them and written to the record.
// Snapshots are rounded ONCE from fp64 to the beam dtype -- closer to the fp64 truth than the
// reference's chain of fp32 7x7 products (BASELINE.md section 2 noise-floor figures).
// Latency for ARES (195 elements) at one setting: 46 us with the serial walk -> see DESIGN.md.
//
// Reference behaviour restated here (desy-ml/cheetah @ 60d1053):
//   cheetah/track_methods.py:17-77      base_rmatrix (quadrupole / sector-bend body)
//   cheetah/track_methods.py:284-299    drift_matrix
//   cheetah/track_methods.py:302-382    rotation / misalignment entry+exit maps
//   cheetah/accelerator/{quadrupole.py:93-110, dipole.py:372-394, :430-466,
//     horizontal_corrector.py:60-78, vertical_corrector.py:60-78,
//     combined_corrector.py:76-98, solenoid.py:74-116, undulator.py:78-125,
//     cavity.py:253-358 (voltage == 0), custom_transfer_map.py:111-114}
//   cheetah/accelerator/segment.py:534-541   tm = R_n @ ... @ R_1
//   cheetah/utils/physics.py:4-19            gamma, 1/gamma^2, beta
#include "ch_common.cuh"

namespace ch {

namespace {

constexpr int kChunk = 256;      // elements per shared-memory chunk
constexpr int kCoef = 16;        // fp64 coefficients per element
constexpr int kLengthSlot = 15;  // coefficient index that always holds the element length
constexpr int kThreads = 256;
constexpr int kGroup = 8;                  // elements per group (phase B1)
constexpr int kGroups = kChunk / kGroup;   // 32 groups = 256 threads of 8 lanes
constexpr int kMaxCuts = CH_MAX_APERTURES + 1;  // apertures + one active cavity per section
constexpr uint32_t kAllFlags = CH_FLAG_XY_UNCOUPLED | CH_FLAG_NO_TAU_COLUMN |
                               CH_FLAG_NO_Y_DISPERSION | CH_FLAG_DELTA_IDENTITY;

struct Relativistic {
  double gamma, igamma2, beta;
};

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

// C(k2, L) = cos(sqrt(k2) L), S(k2, L) = sin(sqrt(k2) L)/sqrt(k2) continued to k2 <= 0:
// the real closed form of the reference's complex sqrt / cos / sinc (track_methods.py:42-49)
__device__ __forceinline__ void cos_sin_like(double k2, double L, double& c, double& s) {
  if (k2 > 0.0) {
    const double k = sqrt(k2);
    double sn, cs;
    sincos(k * L, &sn, &cs);
    c = cs;
    s = sn / k;
  } else if (k2 < 0.0) {
    const double k = sqrt(-k2);
    c = cosh(k * L);
    s = sinh(k * L) / k;
  } else {
    c = 1.0;
    s = L;
  }
}

// (L - S(k2, L)) / k2 = L^3 * si1mdiv(k2 L^2), with the series near 0 (autograd.py:108-128)
__device__ __forceinline__ double l_minus_s_over_k2(double k2, double L, double s) {
  const double x = k2 * L * L;
  if (fabs(x) < 1e-3) {
    // (1 - sinc(sqrt x)) / x = 1/6 - x/120 + x^2/5040 - x^3/362880 + ...
    const double series =
        1.0 / 6.0 + x * (-1.0 / 120.0 + x * (1.0 / 5040.0 + x * (-1.0 / 362880.0)));
    return L * L * L * series;
  }
  return (L - s) / k2;
}

// (1 - C(k2, L)) / k2 = 0.5 L^2 sinc^2(0.5 sqrt(k2) L)   (track_methods.py:51-52)
__device__ __forceinline__ double one_minus_c_over_k2(double k2, double L) {
  double c_half, s_half;
  cos_sin_like(k2, 0.5 * L, c_half, s_half);
  return 2.0 * s_half * s_half;
}

__device__ __forceinline__ double drift_r56(double L, const Relativistic& rel) {
  return -L / (rel.beta * rel.beta) * rel.igamma2;  // track_methods.py:297
}

// coefficient layout of a body (quadrupole / dipole):
//   [4] cx [5] sx [6] -kx2 sx [7] cy [8] sy [9] -ky2 sy [10] r56 [11] dx/beta [12] sx hx/beta
__device__ __forceinline__ void body_coefficients(double* c, double L, double k1, double hx,
                                                  const Relativistic& rel) {
  const double kx2 = k1 + hx * hx;
  const double ky2 = -k1;
  double cx, sx, cy, sy;
  cos_sin_like(kx2, L, cx, sx);
  cos_sin_like(ky2, L, cy, sy);
  const double ibeta = 1.0 / rel.beta;
  double r56 = -L * ibeta * ibeta * rel.igamma2;
  double dx_b = 0.0, sh_b = 0.0;
  if (hx != 0.0) {
    dx_b = hx * one_minus_c_over_k2(kx2, L) * ibeta;
    sh_b = sx * hx * ibeta;
    r56 += hx * hx * l_minus_s_over_k2(kx2, L, sx) * ibeta * ibeta;
  }
  c[4] = cx;
  c[5] = sx;
  c[6] = -kx2 * sx;
  c[7] = cy;
  c[8] = sy;
  c[9] = -ky2 * sy;
  c[10] = r56;
  c[11] = dx_b;
  c[12] = sh_b;
}

// ---- phase A: element parameters -> coefficients -------------------------------------
__device__ void element_coefficients(const Program& prog, int32_t op, int64_t b,
                                     const Relativistic& rel, double mass, double charge,
                                     double* c) {
  const int32_t code = prog.opcodes[op];
  const int32_t s0 = prog.slot_begin[op];
#pragma unroll
  for (int i = 0; i < kCoef; ++i) c[i] = 0.0;
  switch (code) {
    case CH_OP_DRIFT: {
      const double L = slot_value(prog, s0, b);
      c[0] = L;
      c[1] = drift_r56(L, rel);
      c[kLengthSlot] = L;
      break;
    }
    case CH_OP_CORRECTOR: {
      const double L = slot_value(prog, s0, b);
      const int32_t which = prog.op_flags[op];  // bit0: horizontal angle, bit1: vertical angle
      int32_t s = s0 + 1;
      c[0] = L;
      c[1] = drift_r56(L, rel);
      if (which & 1) c[2] = slot_value(prog, s++, b);
      if (which & 2) c[3] = slot_value(prog, s++, b);
      c[kLengthSlot] = L;
      break;
    }
    case CH_OP_QUADRUPOLE: {
      const double L = slot_value(prog, s0, b);
      const double k1 = slot_value(prog, s0 + 1, b);
      const double tilt = slot_value(prog, s0 + 2, b);
      const double mx = slot_value(prog, s0 + 3, b);
      const double my = slot_value(prog, s0 + 4, b);
      double sn = 0.0, cs = 1.0;
      if (tilt != 0.0) sincos(tilt, &sn, &cs);
      c[0] = cs;
      c[1] = sn;
      c[2] = -mx * cs - my * sn;  // entry shift (track_methods.py:374-376)
      c[3] = mx * sn - my * cs;
      body_coefficients(c, L, k1, 0.0, rel);
      c[13] = mx;  // exit shift (track_methods.py:378)
      c[14] = my;
      c[kLengthSlot] = L;
      break;
    }
    case CH_OP_DIPOLE: {
      const double L = slot_value(prog, s0, b);
      const double angle = slot_value(prog, s0 + 1, b);
      const double k1 = slot_value(prog, s0 + 2, b);
      const double e1 = slot_value(prog, s0 + 3, b);
      const double e2 = slot_value(prog, s0 + 4, b);
      const double fint = slot_value(prog, s0 + 5, b);
      const double fint_exit = slot_value(prog, s0 + 6, b);
      const double gap = slot_value(prog, s0 + 7, b);
      const double tilt = slot_value(prog, s0 + 8, b);
      const double hx = angle / L;
      double sn = 0.0, cs = 1.0;
      if (tilt != 0.0) sincos(tilt, &sn, &cs);
      c[0] = cs;
      c[1] = sn;
      {  // entrance pole face (dipole.py:430-447)
        const double se = sin(e1);
        const double phi = fint * hx * gap / cos(e1) * (1.0 + se * se);
        c[2] = hx * tan(e1);
        c[3] = -hx * tan(e1 - phi);
      }
      body_coefficients(c, L, k1, hx, rel);
      {  // exit pole face (dipole.py:449-466) -- uses `gap`, not gap_exit, like the reference
        const double se = sin(e2);
        const double phi = fint_exit * hx * gap / cos(e2) * (1.0 + se * se);
        c[13] = hx * tan(e2);
        c[14] = -hx * tan(e2 - phi);
      }
      c[kLengthSlot] = L;
      break;
    }
    case CH_OP_SOLENOID: {
      const double L = slot_value(prog, s0, b);
      const double k = slot_value(prog, s0 + 1, b);
      double s, co;
      sincos(L * k, &s, &co);
      const double s_k = (k != 0.0) ? s / k : L;
      c[0] = slot_value(prog, s0 + 2, b);  // mx
      c[1] = slot_value(prog, s0 + 3, b);  // my
      c[2] = co * co;
      c[3] = co * s_k;
      c[4] = s * co;
      c[5] = s * s_k;
      c[6] = k * s * co;
      c[7] = k * s * s;
      c[8] = L / (1.0 - rel.gamma * rel.gamma);
      c[kLengthSlot] = L;
      break;
    }
    case CH_OP_UNDULATOR: {
      const double L = slot_value(prog, s0, b);
      const double period = slot_value(prog, s0 + 1, b);
      const double kx = slot_value(prog, s0 + 2, b);
      const double ky = slot_value(prog, s0 + 3, b);
      const double ibeta2 = 1.0 / (rel.beta * rel.beta);
      c[0] = -L * rel.igamma2 * (ibeta2 + 0.5 * (kx * kx + ky * ky));
      const double freq =
          period > 0.0 ? 1.4142135623730951 * 3.141592653589793 / (period * rel.gamma * rel.beta)
                       : 0.0;
      {  // horizontal-plane focusing from ky (undulator.py:114-121)
        const double w = freq * ky;
        double s, co;
        sincos(w * L, &s, &co);
        c[1] = co;
        c[2] = (w != 0.0) ? s / w : L;
        c[3] = -s * w;
      }
      {  // vertical-plane focusing from kx (undulator.py:106-112)
        const double w = freq * kx;
        double s, co;
        sincos(w * L, &s, &co);
        c[4] = co;
        c[5] = (w != 0.0) ? s / w : L;
        c[6] = -s * w;
      }
      c[kLengthSlot] = L;
      break;
    }
    case CH_OP_CAVITY_OFF: {
      // cavity.py:253-358 at voltage == 0: alpha = 0 -> r11 = r22 = 1, r12 = L, r21 = 0,
      // r55 = r66 = 1, r65 = 0.  Standing wave: r56 = -L (Ef+Ei)/(Ef^2 Ei b1 (b1+b0)) with
      // Ef = Ei, b1 = b0; traveling wave: r56 = 0.
      const double L = slot_value(prog, s0, b);
      const double g = rel.gamma;
      c[0] = L;
      c[1] = (prog.op_flags[op] & 1) ? 0.0
                                     : -L / (g * g * g * rel.beta) * (2.0 * g) / (2.0 * rel.beta);
      c[kLengthSlot] = L;
      break;
    }
    case CH_OP_CUSTOM_MAP: {
      if (prog.slot_begin[op + 1] - s0 >= 2) c[kLengthSlot] = slot_value(prog, s0 + 1, b);
      break;
    }
    case CH_OP_APERTURE: {
      c[0] = slot_value(prog, s0, b);
      c[1] = slot_value(prog, s0 + 1, b);
      break;
    }
    case CH_OP_CAVITY: {
      // cavity.py:253-358 (R matrix) and :113-220 (exact delta update, T566/T556/T555)
      const double L = slot_value(prog, s0, b);
      const double V = slot_value(prog, s0 + 1, b);
      const double phi = slot_value(prog, s0 + 2, b) * (3.141592653589793 / 180.0);
      const double k = 2.0 * 3.141592653589793 * slot_value(prog, s0 + 3, b) / 299792458.0;
      const int32_t flags = prog.op_flags[op];
      // slot 4: device flag = the reference's batch-wide `(delta_energy > 0).any()`
      const bool energy_gain_terms = slot_value(prog, s0 + 4, b) != 0.0;
      const double E = rel.gamma * mass, m = mass;
      const double v_eff = -V * charge;
      double sphi, cphi;
      sincos(phi, &sphi, &cphi);
      const double delta_energy = v_eff * cphi;
      const double Ei = E / m, dE = delta_energy / m, Ef = Ei + dE, Ep = dE / L;
      const double E1 = E + delta_energy;
      const double beta0 = sqrt(1.0 - 1.0 / (Ei * Ei)), beta1 = sqrt(1.0 - 1.0 / (Ef * Ef));
      auto log1pdiv = [](double x) { return x != 0.0 ? log1p(x) / x : 1.0; };
      if (!(flags & 1)) {  // standing wave
        const double lg = log1pdiv(delta_energy / E);
        const double alpha = 0.3535533905932738 * v_eff / E * lg;  // sqrt(1/8)
        double sa, ca;
        sincos(alpha, &sa, &ca);
        const double rt2 = 1.4142135623730951;
        c[0] = ca - rt2 * cphi * sa;
        c[1] = (alpha != 0.0 ? sa / alpha : 1.0) * lg * L;
        c[2] = -(v_eff / (E1 * rt2 * L) * (0.5 + cphi * cphi) * sa);
        c[3] = Ei / Ef * (ca + rt2 * cphi * sa);
        c[4] = 1.0 + (dE != 0.0 ? k * L * beta0 * tan(phi) * (Ei * Ef * (beta0 * beta1 - 1.0) + 1.0) /
                                      (beta1 * Ef * dE)
                                : 0.0);
        c[5] = -L / (Ef * Ef * Ei * beta1) * (Ef + Ei) / (beta1 + beta0);
        c[6] = k * sphi * v_eff / (beta1 * E1);
        c[7] = Ei / Ef * beta0 / beta1;
      } else {  // traveling wave (Rosenzweig & Serafini)
        const double f = L * log1pdiv(dE / Ei);
        const double fi = -Ep / (2.0 * Ei), fo = Ep / (2.0 * Ef), body22 = Ei / Ef;
        c[0] = 1.0 + f * fi;
        c[1] = f;
        c[2] = fo * c[0] + body22 * fi;
        c[3] = fo * f + body22;
        c[4] = 1.0;
        c[5] = 0.0;
        c[6] = k * sphi * v_eff / E1;
        c[7] = c[3];
      }
      c[8] = E * rel.beta / (E1 * beta1);
      c[9] = V * rel.beta / (E1 * beta1);
      c[10] = rel.beta * k;
      c[11] = phi;
      const double g0 = rel.gamma, g1 = E1 / m, b0 = rel.beta, b1 = beta1;
      if (energy_gain_terms) {
        const double dgamma = V / m;
        const double b13 = b1 * b1 * b1, g13 = g1 * g1 * g1, b03 = b0 * b0 * b0, g03 = g0 * g0 * g0;
        c[12] = L * (b03 * g03 - b13 * g13) / (2.0 * b0 * b13 * g0 * (g0 - g1) * g13);
        c[13] = b0 * k * L * dgamma * g0 * (b13 * g13 + b0 * (g0 - g13)) * sphi /
                (b13 * g13 * (g0 - g1) * (g0 - g1));
        c[14] = b0 * b0 * k * k * L * dgamma / 2.0 *
                (dgamma * (2.0 * g0 * g13 * (b0 * b13 - 1.0) + g0 * g0 + 3.0 * g1 * g1 - 2.0) /
                     (b13 * g13 * (g0 - g1) * (g0 - g1) * (g0 - g1)) * sphi * sphi -
                 (g1 * g0 * (b1 * b0 - 1.0) + 1.0) / (b1 * g1 * (g0 - g1) * (g0 - g1)) * cphi);
      } else {
        c[12] = 1.5 * L * rel.igamma2 / (b0 * b0 * b0);
      }
      c[kLengthSlot] = L;
      break;
    }
    default:
      break;
  }
}

// ---- phase B: apply one element to this lane's column --------------------------------
// v[0..5] is column `lane` of the cumulative map; `one` is 1.0 on lane 6 (the affine
// column, where row 6 of the map contributes its 1) and 0.0 elsewhere.
__device__ __forceinline__ void mix(double& a, double& b, double c00, double c01, double c10,
                                    double c11) {
  const double ra = a, rb = b;
  a = fma(c00, ra, c01 * rb);
  b = fma(c10, ra, c11 * rb);
}

__device__ __forceinline__ void apply_body_column(double* v, const double* c, bool bend) {
  if (bend) {
    // row 4 uses the OLD rows 0, 1 and 5
    v[4] = fma(c[12], v[0], fma(c[11], v[1], fma(c[10], v[5], v[4])));
    const double r0 = v[0], r1 = v[1];
    v[0] = fma(c[4], r0, fma(c[5], r1, c[11] * v[5]));
    v[1] = fma(c[6], r0, fma(c[4], r1, c[12] * v[5]));
  } else {
    v[4] = fma(c[10], v[5], v[4]);
    mix(v[0], v[1], c[4], c[5], c[6], c[4]);
  }
  mix(v[2], v[3], c[7], c[8], c[9], c[7]);
}

__device__ __forceinline__ void rotate_column(double* v, double cs, double sn) {
  mix(v[0], v[2], cs, sn, -sn, cs);
  mix(v[1], v[3], cs, sn, -sn, cs);
}

template <typename T>
__device__ __forceinline__ T pack_flags(uint32_t f);
template <>
__device__ __forceinline__ float pack_flags<float>(uint32_t f) {
  return __uint_as_float(f);
}
template <>
__device__ __forceinline__ double pack_flags<double>(uint32_t f) {
  return __longlong_as_double(static_cast<long long>(f));
}


// Apply the map of one element to this lane's column (CH_OP_APERTURE and the snapshot part of
// CH_OP_CAVITY are handled by the caller).  `one` is 1.0 on lane 6 (the affine column).
__device__ __forceinline__ void apply_element_to_column(const Program& prog, int32_t op,
                                                        int32_t code, const double* c, int64_t b,
                                                        double* v, double one) {
  switch (code) {
    case CH_OP_DRIFT:
    case CH_OP_CAVITY_OFF:
      v[0] = fma(c[0], v[1], v[0]);
      v[2] = fma(c[0], v[3], v[2]);
      v[4] = fma(c[1], v[5], v[4]);
      break;
    case CH_OP_CORRECTOR:
      v[0] = fma(c[0], v[1], v[0]);
      v[2] = fma(c[0], v[3], v[2]);
      v[4] = fma(c[1], v[5], v[4]);
      v[1] = fma(c[2], one, v[1]);
      v[3] = fma(c[3], one, v[3]);
      break;
    case CH_OP_QUADRUPOLE:
      if (c[1] != 0.0 || c[0] != 1.0) rotate_column(v, c[0], c[1]);
      v[0] = fma(c[2], one, v[0]);
      v[2] = fma(c[3], one, v[2]);
      apply_body_column(v, c, false);
      if (c[1] != 0.0 || c[0] != 1.0) rotate_column(v, c[0], -c[1]);
      v[0] = fma(c[13], one, v[0]);
      v[2] = fma(c[14], one, v[2]);
      break;
    case CH_OP_DIPOLE:
      if (c[1] != 0.0 || c[0] != 1.0) rotate_column(v, c[0], c[1]);
      v[1] = fma(c[2], v[0], v[1]);
      v[3] = fma(c[3], v[2], v[3]);
      apply_body_column(v, c, true);
      v[1] = fma(c[13], v[0], v[1]);
      v[3] = fma(c[14], v[2], v[3]);
      if (c[1] != 0.0 || c[0] != 1.0) rotate_column(v, c[0], -c[1]);
      break;
    case CH_OP_SOLENOID: {
      v[0] = fma(-c[0], one, v[0]);
      v[2] = fma(-c[1], one, v[2]);
      v[4] = fma(c[8], v[5], v[4]);
      const double r0 = v[0], r1 = v[1], r2 = v[2], r3 = v[3];
      v[0] = c[2] * r0 + c[3] * r1 + c[4] * r2 + c[5] * r3;
      v[1] = -c[6] * r0 + c[2] * r1 - c[7] * r2 + c[4] * r3;
      v[2] = -c[4] * r0 - c[5] * r1 + c[2] * r2 + c[3] * r3;
      v[3] = c[7] * r0 - c[4] * r1 - c[6] * r2 + c[2] * r3;
      v[0] = fma(c[0], one, v[0]);
      v[2] = fma(c[1], one, v[2]);
      break;
    }
    case CH_OP_UNDULATOR:
      v[4] = fma(c[0], v[5], v[4]);
      mix(v[0], v[1], c[1], c[2], c[3], c[1]);
      mix(v[2], v[3], c[4], c[5], c[6], c[4]);
      break;
    case CH_OP_CUSTOM_MAP: {
      // dense user map: read straight from the parameter tensor (rare, latency-tolerant)
      const ScalarRef ref = prog.slots[prog.slot_begin[op]];
      double w[6];
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        double acc = one * load_scalar(ref.ptr, b * ref.stride + r * 7 + 6, ref.dtype);
#pragma unroll
        for (int k = 0; k < 6; ++k)
          acc = fma(load_scalar(ref.ptr, b * ref.stride + r * 7 + k, ref.dtype), v[k], acc);
        w[r] = acc;
      }
#pragma unroll
      for (int r = 0; r < 6; ++r) v[r] = w[r];
      break;
    }
    case CH_OP_CAVITY:
      mix(v[0], v[1], c[0], c[1], c[2], c[3]);
      mix(v[2], v[3], c[0], c[1], c[2], c[3]);
      mix(v[4], v[5], c[4], c[5], c[6], c[7]);
      break;
    default:
      break;
  }
}

// w = G . v for one column v of a 7x7 affine map (row 6 of G is 0 0 0 0 0 0 1), G row-major 6x7
// in shared memory; `one` is 1.0 for the affine column.
__device__ __forceinline__ void left_multiply_column(const double* g, double* v, double one) {
  double w[6];
#pragma unroll
  for (int r = 0; r < 6; ++r) {
    double acc = g[r * 7 + 6] * one;
#pragma unroll
    for (int k = 5; k >= 0; --k) acc = fma(g[r * 7 + k], v[k], acc);
    w[r] = acc;
  }
#pragma unroll
  for (int r = 0; r < 6; ++r) v[r] = w[r];
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
compose_maps_kernel(Program prog, int32_t op_begin, int32_t op_end, ScalarRef energy,
                    ScalarRef mass, ScalarRef charge, int32_t n_apertures, T* __restrict__ records,
                    int64_t record_len) {
  extern __shared__ __align__(16) unsigned char compose_smem[];
  double (*coef)[kCoef] = reinterpret_cast<double (*)[kCoef]>(compose_smem);  // [kChunk][kCoef]
  __shared__ int32_t codes[kChunk];
  __shared__ int32_t warp_cuts[kThreads / 32];
  __shared__ double group_map[kGroups][44];       // 6x7 product of a group, [42] = its length
  __shared__ double prefix_map[kGroups + 1][42];  // product of all groups BEFORE group g
  __shared__ double cut_rows[kMaxCuts][2][8];     // two rows of the group-partial map at a cut
  __shared__ int32_t cut_element[kMaxCuts];       // element index (in the chunk) of each cut
  __shared__ int32_t n_cuts_shared;
  __shared__ uint32_t flags_shared;

  // Programmatic dependent launch: the apply kernel that follows in the stream may start its
  // prologue (particle tile -> registers) now; it waits (griddepcontrol.wait) for this grid to
  // finish before it reads a record.
  asm volatile("griddepcontrol.launch_dependents;");
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  const int lane = tid & 7;     // column owned inside a group (7: length accumulator)
  const int group = tid >> 3;

  Relativistic rel;
  const double mass_value = load_scalar(mass.ptr, 0, mass.dtype);
  const double charge_value = charge.ptr ? load_scalar(charge.ptr, 0, charge.dtype) : -1.0;
  {
    const double e = load_scalar(energy.ptr, b * energy.stride, energy.dtype);
    const double m = mass_value;
    rel.gamma = e / m;
    rel.igamma2 = 1.0 / (rel.gamma * rel.gamma);
    rel.beta = sqrt(1.0 - rel.igamma2);
  }

  T* rec = records + b * record_len;
  T* aperture_rec = rec + CH_RECORD_HEADER + CH_RECORD_MAP;
  T* cavity_rec = aperture_rec + n_apertures * CH_RECORD_APERTURE;
  const double one = (lane == 6) ? 1.0 : 0.0;

  // running product of everything before the current chunk, column `tid` held by thread tid < 7
  double cum[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) cum[i] = (i == tid) ? 1.0 : 0.0;
  double total_length = 0.0;  // thread 7
  int32_t apertures_done = 0;
  if (tid == 0) flags_shared = kAllFlags;

  for (int32_t chunk = op_begin; chunk < op_end; chunk += kChunk) {
    const int32_t n = min(kChunk, op_end - chunk);
    // ---- A: coefficients; cut points in element order --------------------------------------
    if (tid < n) {
      codes[tid] = prog.opcodes[chunk + tid];
      element_coefficients(prog, chunk + tid, b, rel, mass_value, charge_value, coef[tid]);
    }
    {  // ordered list of the cut elements: ballot inside a warp, prefix over the warp totals
      const bool is_cut =
          tid < n && (codes[tid] == CH_OP_APERTURE || codes[tid] == CH_OP_CAVITY);
      const uint32_t mask = __ballot_sync(0xffffffffu, is_cut);
      if ((tid & 31) == 0) warp_cuts[tid >> 5] = __popc(mask);
      __syncthreads();
      int32_t before = 0, total = 0;
#pragma unroll
      for (int w = 0; w < kThreads / 32; ++w) {
        const int32_t count = warp_cuts[w];
        if (w < (tid >> 5)) before += count;
        total += count;
      }
      if (is_cut) cut_element[before + __popc(mask & ((1u << (tid & 31)) - 1u))] = tid;
      if (tid == 0) n_cuts_shared = total;
    }
    __syncthreads();
    const int32_t n_cuts = n_cuts_shared;

    // ---- B1: product of each group of 8 elements, column per lane ---------------------------
    {
      double v[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) v[i] = (i == lane) ? 1.0 : 0.0;
      double group_length = 0.0;
      const int32_t first = group * kGroup;
      int32_t cut = 0;  // ordinal of the next cut at or after `first`
      while (cut < n_cuts && cut_element[cut] < first) ++cut;
      for (int32_t i = first; i < min(first + kGroup, n); ++i) {
        const double* c = coef[i];
        const int32_t code = codes[i];
        group_length += c[kLengthSlot];
        if (code == CH_OP_APERTURE || code == CH_OP_CAVITY) {
          // rows (x, y) at an aperture, (tau, delta) at a cavity entrance, relative to the group
          const int r0 = code == CH_OP_APERTURE ? 0 : 4, r1 = code == CH_OP_APERTURE ? 2 : 5;
          if (lane < 7) {
            cut_rows[cut][0][lane] = v[r0];
            cut_rows[cut][1][lane] = v[r1];
          }
          ++cut;
        }
        if (code != CH_OP_APERTURE) apply_element_to_column(prog, chunk + i, code, c, b, v, one);
      }
      if (lane < 7) {
#pragma unroll
        for (int r = 0; r < 6; ++r) group_map[group][r * 7 + lane] = v[r];
      } else {
        group_map[group][42] = group_length;
      }
    }
    __syncthreads();

    // ---- B2: exclusive prefix over the groups (thread c < 7 owns column c) -------------------
    const int32_t n_groups = (n + kGroup - 1) / kGroup;
    if (tid < 7) {
      const double one_c = (tid == 6) ? 1.0 : 0.0;
      for (int32_t g = 0; g < n_groups; ++g) {
#pragma unroll
        for (int r = 0; r < 6; ++r) prefix_map[g][r * 7 + tid] = cum[r];
        left_multiply_column(group_map[g], cum, one_c);
      }
    } else if (tid == 7) {
      for (int32_t g = 0; g < n_groups; ++g) total_length += group_map[g][42];
    }
    __syncthreads();

    // ---- B3: cut snapshots = (rows relative to the group) . (prefix of the groups before) ------
    for (int32_t task = tid; task < n_cuts * 16; task += kThreads) {
      const int32_t cut = task >> 4, slot = task & 15;
      const int32_t element = cut_element[cut];
      const int32_t code = codes[element];
      const double* c = coef[element];
      // aperture ordinal within the section: cuts of this chunk before it that are apertures
      int32_t ordinal = apertures_done;
      for (int32_t k = 0; k < cut; ++k) ordinal += codes[cut_element[k]] == CH_OP_APERTURE;
      T* block = code == CH_OP_APERTURE ? aperture_rec + ordinal * CH_RECORD_APERTURE : cavity_rec;
      if (slot < 14) {
        const int row = slot / 7, col = slot - row * 7;
        const double* p = cut_rows[cut][row];
        const double* pre = prefix_map[element / kGroup];
        double acc = (col == 6) ? p[6] : 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) acc = fma(p[k], pre[k * 7 + col], acc);
        const T value = static_cast<T>(acc);
        block[slot] = value;
        if (code == CH_OP_APERTURE) {
          uint32_t f = kAllFlags;
          const bool is_x = row == 0;
          if (is_x && (col == 2 || col == 3) && value != T(0)) f &= ~CH_FLAG_XY_UNCOUPLED;
          if (!is_x && (col == 0 || col == 1) && value != T(0)) f &= ~CH_FLAG_XY_UNCOUPLED;
          if (col == 4 && value != T(0)) f &= ~CH_FLAG_NO_TAU_COLUMN;
          if (!is_x && col == 5 && value != T(0)) f &= ~CH_FLAG_NO_Y_DISPERSION;
          if (f != kAllFlags) atomicAnd(&flags_shared, f);
        }
      } else if (code == CH_OP_APERTURE) {
        block[slot] = static_cast<T>(c[slot - 14]);  // x_max, y_max
      } else if (slot == 14) {  // cavity constants (cavity.py:113-220), see the record layout
        double sphi, cphi;
        sincos(c[11], &sphi, &cphi);
        block[14] = static_cast<T>(c[8]);
        block[15] = static_cast<T>(c[9]);
        block[16] = static_cast<T>(c[10]);
        block[17] = static_cast<T>(sphi);
        block[18] = static_cast<T>(cphi);
        block[19] = static_cast<T>(c[12]);
        block[20] = static_cast<T>(c[13]);
        block[21] = static_cast<T>(c[14]);
        block[22] = block[23] = T(0);
      }
    }
    for (int32_t k = 0; k < n_cuts; ++k) apertures_done += codes[cut_element[k]] == CH_OP_APERTURE;
    __syncthreads();  // shared tables are rebuilt by the next chunk
  }

  // ---- final map, flags, length ----------------------------------------------------------------
  if (tid < 7) {
    uint32_t f = kAllFlags;
    T w[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      w[r] = static_cast<T>(cum[r]);
      rec[CH_RECORD_HEADER + r * 7 + tid] = w[r];
    }
    const bool z01 = w[0] != T(0) || w[1] != T(0);
    const bool z23 = w[2] != T(0) || w[3] != T(0);
    if ((tid == 2 || tid == 3) && z01) f &= ~CH_FLAG_XY_UNCOUPLED;
    if ((tid == 0 || tid == 1) && z23) f &= ~CH_FLAG_XY_UNCOUPLED;
    if (tid == 4 && (z01 || z23 || w[5] != T(0))) f &= ~CH_FLAG_NO_TAU_COLUMN;
    if (w[5] != ((tid == 5) ? T(1) : T(0))) f &= ~CH_FLAG_DELTA_IDENTITY;
    if (tid == 5 && z23) f &= ~CH_FLAG_NO_Y_DISPERSION;
    if ((tid == 2 || tid == 3) && w[4] != T(0)) f &= ~CH_FLAG_NO_Y_DISPERSION;
    if (f != kAllFlags) atomicAnd(&flags_shared, f);
  } else if (tid == 7) {
    rec[1] = static_cast<T>(total_length);
  }
  __syncthreads();
  if (tid == 0) rec[0] = pack_flags<T>(flags_shared);
}

}  // namespace
}  // namespace ch

extern "C" int ch_compose_maps(const ch_program* program, int32_t op_begin, int32_t op_end,
                               int64_t n_settings, const void* energy, int64_t energy_stride,
                               int32_t energy_dtype, const void* mass_eV, int32_t mass_dtype,
                               const void* num_elementary_charges, int32_t charge_dtype,
                               void* records, int64_t record_len, int32_t record_dtype,
                               void* stream) {
  CH_REQUIRE(program != nullptr, "ch_compose_maps: program is NULL");
  CH_REQUIRE(op_begin >= 0 && op_begin <= op_end && op_end <= program->n_ops,
             "ch_compose_maps: op range [%d, %d) outside program of %d ops", op_begin, op_end,
             program->n_ops);
  CH_REQUIRE(n_settings > 0 && n_settings <= 2147483647LL,
             "ch_compose_maps: n_settings must be in [1, 2^31)");
  CH_REQUIRE(energy && mass_eV && records, "ch_compose_maps: NULL pointer argument");
  CH_REQUIRE(record_len >= CH_RECORD_LEN(0), "ch_compose_maps: record_len %lld too small",
             static_cast<long long>(record_len));
  CH_REQUIRE(record_dtype == CH_F32 || record_dtype == CH_F64, "ch_compose_maps: bad dtype");

  ch::Program prog{program->opcodes, program->op_flags, program->slot_begin, program->slots};
  ch::ScalarRef e{energy, energy_stride, energy_dtype};
  ch::ScalarRef m{mass_eV, 0, mass_dtype};
  ch::ScalarRef q{num_elementary_charges, 0, charge_dtype};
  // apertures in the record: everything beyond header + map, minus an optional cavity block
  const int64_t extra = record_len - CH_RECORD_LEN(0);
  const int32_t n_apertures = static_cast<int32_t>(
      (extra % CH_RECORD_APERTURE == 0 ? extra : extra - CH_RECORD_CAVITY) / CH_RECORD_APERTURE);
  const unsigned blocks = static_cast<unsigned>(n_settings);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int smem = ch::kChunk * ch::kCoef * sizeof(double);  // + ~27 KB static: opt in above 48 KB
  if (record_dtype == CH_F32) {
    CH_CUDA(ch::allow_dynamic_smem(
        reinterpret_cast<const void*>(ch::compose_maps_kernel<float>), smem));
    ch::compose_maps_kernel<float><<<blocks, ch::kThreads, smem, s>>>(
        prog, op_begin, op_end, e, m, q, n_apertures, static_cast<float*>(records), record_len);
  } else {
    CH_CUDA(ch::allow_dynamic_smem(
        reinterpret_cast<const void*>(ch::compose_maps_kernel<double>), smem));
    ch::compose_maps_kernel<double><<<blocks, ch::kThreads, smem, s>>>(
        prog, op_begin, op_end, e, m, q, n_apertures, static_cast<double*>(records), record_len);
  }
  CH_LAUNCH_CHECK();
  return CH_OK;
}

// Map composition: one thread per lattice setting walks the lowered element program,
// builds every first-order map in fp64 registers with closed real forms and
// left-multiplies it into the cumulative 6x7 affine map using each element's sparsity
// (a drift costs 21 FMAs, not 343).  Cut points (apertures, section end) snapshot the
// cumulative map into the per-setting record consumed by ch_apply_maps.
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

struct Map {
  double m[6][7];  // rows 0-5 of the cumulative map; row 6 is implicit (0 0 0 0 0 0 1)
};

__device__ __forceinline__ void set_identity(Map& M) {
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 7; ++j) M.m[i][j] = (i == j) ? 1.0 : 0.0;
}

// row[a] += f * row[b]
__device__ __forceinline__ void axpy_row(Map& M, int a, int b, double f) {
#pragma unroll
  for (int j = 0; j < 7; ++j) M.m[a][j] = fma(f, M.m[b][j], M.m[a][j]);
}

// (row[a], row[b]) <- (c00 row[a] + c01 row[b], c10 row[a] + c11 row[b])
__device__ __forceinline__ void mix_rows(Map& M, int a, int b, double c00, double c01, double c10,
                                         double c11) {
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const double ra = M.m[a][j], rb = M.m[b][j];
    M.m[a][j] = fma(c00, ra, c01 * rb);
    M.m[b][j] = fma(c10, ra, c11 * rb);
  }
}

// x-y rotation by `angle` (entry) or its transpose (exit): track_methods.py:302-323
__device__ __forceinline__ void rotate(Map& M, double cs, double sn) {
  mix_rows(M, 0, 2, cs, sn, -sn, cs);
  mix_rows(M, 1, 3, cs, sn, -sn, cs);
}

struct Relativistic {
  double gamma, igamma2, beta;
};

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

// Body of a thick quadrupole / combined-function sector bend: track_methods.py:17-77
__device__ __forceinline__ void apply_body(Map& M, double L, double k1, double hx,
                                           const Relativistic& rel) {
  const double kx2 = k1 + hx * hx;
  const double ky2 = -k1;
  double cx, sx, cy, sy;
  cos_sin_like(kx2, L, cx, sx);
  cos_sin_like(ky2, L, cy, sy);
  const double ibeta = 1.0 / rel.beta;
  const double ibeta2 = ibeta * ibeta;
  double r56 = -L * ibeta2 * rel.igamma2;
  if (hx != 0.0) {
    const double dx_b = hx * one_minus_c_over_k2(kx2, L) * ibeta;  // R05 = R41
    const double sh_b = sx * hx * ibeta;                           // R15 = R40
    r56 += hx * hx * l_minus_s_over_k2(kx2, L, sx) * ibeta2;
    // row 4 uses the OLD rows 0, 1 and 5
    axpy_row(M, 4, 0, sh_b);
    axpy_row(M, 4, 1, dx_b);
    axpy_row(M, 4, 5, r56);
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const double r0 = M.m[0][j], r1 = M.m[1][j], r5 = M.m[5][j];
      M.m[0][j] = fma(cx, r0, fma(sx, r1, dx_b * r5));
      M.m[1][j] = fma(-kx2 * sx, r0, fma(cx, r1, sh_b * r5));
    }
  } else {
    axpy_row(M, 4, 5, r56);
    mix_rows(M, 0, 1, cx, sx, -kx2 * sx, cx);
  }
  mix_rows(M, 2, 3, cy, sy, -ky2 * sy, cy);
}

__device__ __forceinline__ void apply_drift(Map& M, double L, const Relativistic& rel) {
  axpy_row(M, 0, 1, L);
  axpy_row(M, 2, 3, L);
  axpy_row(M, 4, 5, -L / (rel.beta * rel.beta) * rel.igamma2);
}

struct Program {
  const int32_t* opcodes;
  const int32_t* op_flags;
  const int32_t* slot_begin;
  const ScalarRef* slots;
};

__device__ __forceinline__ double slot_value(const Program& prog, int32_t slot, int64_t b,
                                             int64_t inner = 0) {
  const ScalarRef ref = prog.slots[slot];
  return load_scalar(ref.ptr, b * ref.stride + inner, ref.dtype);
}

template <typename T>
__device__ __forceinline__ void store_rows(T* dst, const Map& M, int row_a, int row_b) {
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    dst[j] = static_cast<T>(M.m[row_a][j]);
    dst[7 + j] = static_cast<T>(M.m[row_b][j]);
  }
}

template <typename T>
struct FlagBits;
template <>
struct FlagBits<float> {
  static __device__ __forceinline__ float pack(uint32_t f) { return __uint_as_float(f); }
};
template <>
struct FlagBits<double> {
  static __device__ __forceinline__ double pack(uint32_t f) {
    return __longlong_as_double(static_cast<long long>(f));
  }
};

// sparsity of one 2-row (aperture) or 6-row (final) snapshot, on the values AS STORED
template <typename T>
__device__ __forceinline__ uint32_t aperture_flags(const T* rows) {
  uint32_t f = CH_FLAG_XY_UNCOUPLED | CH_FLAG_NO_TAU_COLUMN | CH_FLAG_NO_Y_DISPERSION |
               CH_FLAG_DELTA_IDENTITY;
  const T* x = rows;
  const T* y = rows + 7;
  if (x[2] != T(0) || x[3] != T(0) || y[0] != T(0) || y[1] != T(0)) f &= ~CH_FLAG_XY_UNCOUPLED;
  if (x[4] != T(0) || y[4] != T(0)) f &= ~CH_FLAG_NO_TAU_COLUMN;
  if (y[5] != T(0)) f &= ~CH_FLAG_NO_Y_DISPERSION;
  return f;
}

template <typename T>
__device__ __forceinline__ uint32_t map_flags(const T* m) {  // m: 6x7 row-major
  uint32_t f = CH_FLAG_XY_UNCOUPLED | CH_FLAG_NO_TAU_COLUMN | CH_FLAG_NO_Y_DISPERSION |
               CH_FLAG_DELTA_IDENTITY;
  auto at = [&](int i, int j) { return m[i * 7 + j]; };
  if (at(0, 2) != T(0) || at(0, 3) != T(0) || at(1, 2) != T(0) || at(1, 3) != T(0) ||
      at(2, 0) != T(0) || at(2, 1) != T(0) || at(3, 0) != T(0) || at(3, 1) != T(0))
    f &= ~CH_FLAG_XY_UNCOUPLED;
  if (at(0, 4) != T(0) || at(1, 4) != T(0) || at(2, 4) != T(0) || at(3, 4) != T(0) ||
      at(5, 4) != T(0))
    f &= ~CH_FLAG_NO_TAU_COLUMN;
  if (at(5, 0) != T(0) || at(5, 1) != T(0) || at(5, 2) != T(0) || at(5, 3) != T(0) ||
      at(5, 4) != T(0) || at(5, 5) != T(1) || at(5, 6) != T(0))
    f &= ~CH_FLAG_DELTA_IDENTITY;
  if (at(2, 5) != T(0) || at(3, 5) != T(0) || at(4, 2) != T(0) || at(4, 3) != T(0))
    f &= ~CH_FLAG_NO_Y_DISPERSION;
  return f;
}

template <typename T>
__global__ void __launch_bounds__(32)
compose_maps_kernel(Program prog, int32_t op_begin, int32_t op_end, int64_t n_settings,
                    ScalarRef energy, ScalarRef mass, T* __restrict__ records,
                    int64_t record_len) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= n_settings) return;

  Relativistic rel;
  {
    const double e = load_scalar(energy.ptr, b * energy.stride, energy.dtype);
    const double m = load_scalar(mass.ptr, 0, mass.dtype);
    rel.gamma = e / m;
    rel.igamma2 = 1.0 / (rel.gamma * rel.gamma);
    rel.beta = sqrt(1.0 - rel.igamma2);
  }

  Map M;
  set_identity(M);
  T* rec = records + b * record_len;
  T* aperture_rec = rec + CH_RECORD_HEADER + CH_RECORD_MAP;
  uint32_t flags = CH_FLAG_XY_UNCOUPLED | CH_FLAG_NO_TAU_COLUMN | CH_FLAG_NO_Y_DISPERSION |
                   CH_FLAG_DELTA_IDENTITY;
  double total_length = 0.0;  // sum of element lengths: s_out = s_in + total (element.py:183)

  for (int32_t op = op_begin; op < op_end; ++op) {
    const int32_t code = prog.opcodes[op];
    const int32_t s0 = prog.slot_begin[op];
    switch (code) {
      case CH_OP_IDENTITY:
        break;
      case CH_OP_DRIFT: {
        const double L = slot_value(prog, s0, b);
        total_length += L;
        apply_drift(M, L, rel);
        break;
      }
      case CH_OP_CORRECTOR: {
        const double L = slot_value(prog, s0, b);
        total_length += L;
        apply_drift(M, L, rel);
        // angle slots may be absent (horizontal-only / vertical-only correctors)
        const int32_t n = prog.slot_begin[op + 1] - s0;
        const int32_t which = prog.op_flags[op];  // bit0: has horizontal, bit1: has vertical
        int32_t s = s0 + 1;
        if ((which & 1) && s < s0 + n) M.m[1][6] += slot_value(prog, s++, b);
        if ((which & 2) && s < s0 + n) M.m[3][6] += slot_value(prog, s++, b);
        break;
      }
      case CH_OP_QUADRUPOLE: {
        const double L = slot_value(prog, s0, b);
        total_length += L;
        const double k1 = slot_value(prog, s0 + 1, b);
        const double tilt = slot_value(prog, s0 + 2, b);
        const double mx = slot_value(prog, s0 + 3, b);
        const double my = slot_value(prog, s0 + 4, b);
        double sn = 0.0, cs = 1.0;
        const bool tilted = tilt != 0.0;
        if (tilted) sincos(tilt, &sn, &cs);
        // entry: shift by the misalignment, then rotate (track_methods.py:345-382)
        if (tilted) rotate(M, cs, sn);
        M.m[0][6] += -mx * cs - my * sn;
        M.m[2][6] += mx * sn - my * cs;
        apply_body(M, L, k1, 0.0, rel);
        if (tilted) rotate(M, cs, -sn);
        M.m[0][6] += mx;
        M.m[2][6] += my;
        break;
      }
      case CH_OP_DIPOLE: {
        const double L = slot_value(prog, s0, b);
        total_length += L;
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
        const bool tilted = tilt != 0.0;
        if (tilted) {
          sincos(tilt, &sn, &cs);
          rotate(M, cs, sn);
        }
        {  // entrance pole face (dipole.py:430-447)
          const double se = sin(e1);
          const double phi = fint * hx * gap / cos(e1) * (1.0 + se * se);
          axpy_row(M, 1, 0, hx * tan(e1));
          axpy_row(M, 3, 2, -hx * tan(e1 - phi));
        }
        apply_body(M, L, k1, hx, rel);
        {  // exit pole face (dipole.py:449-466) -- uses `gap`, not gap_exit, like the reference
          const double se = sin(e2);
          const double phi = fint_exit * hx * gap / cos(e2) * (1.0 + se * se);
          axpy_row(M, 1, 0, hx * tan(e2));
          axpy_row(M, 3, 2, -hx * tan(e2 - phi));
        }
        if (tilted) rotate(M, cs, -sn);
        break;
      }
      case CH_OP_SOLENOID: {
        const double L = slot_value(prog, s0, b);
        total_length += L;
        const double k = slot_value(prog, s0 + 1, b);
        const double mx = slot_value(prog, s0 + 2, b);
        const double my = slot_value(prog, s0 + 3, b);
        M.m[0][6] -= mx;
        M.m[2][6] -= my;
        double s, c;
        sincos(L * k, &s, &c);
        const double s_k = (k != 0.0) ? s / k : L;
        axpy_row(M, 4, 5, L / (1.0 - rel.gamma * rel.gamma));
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          const double r0 = M.m[0][j], r1 = M.m[1][j], r2 = M.m[2][j], r3 = M.m[3][j];
          M.m[0][j] = c * c * r0 + c * s_k * r1 + s * c * r2 + s * s_k * r3;
          M.m[1][j] = -k * s * c * r0 + c * c * r1 - k * s * s * r2 + s * c * r3;
          M.m[2][j] = -s * c * r0 - s * s_k * r1 + c * c * r2 + c * s_k * r3;
          M.m[3][j] = k * s * s * r0 - s * c * r1 - k * s * c * r2 + c * c * r3;
        }
        M.m[0][6] += mx;
        M.m[2][6] += my;
        break;
      }
      case CH_OP_UNDULATOR: {
        const double L = slot_value(prog, s0, b);
        total_length += L;
        const double period = slot_value(prog, s0 + 1, b);
        const double kx = slot_value(prog, s0 + 2, b);
        const double ky = slot_value(prog, s0 + 3, b);
        const double ibeta2 = 1.0 / (rel.beta * rel.beta);
        axpy_row(M, 4, 5, -L * rel.igamma2 * (ibeta2 + 0.5 * (kx * kx + ky * ky)));
        const double freq =
            period > 0.0 ? 1.4142135623730951 * 3.141592653589793 / (period * rel.gamma * rel.beta)
                         : 0.0;
        {  // vertical-plane focusing from kx (undulator.py:106-112)
          const double w = freq * kx;
          double s, c;
          sincos(w * L, &s, &c);
          mix_rows(M, 2, 3, c, (w != 0.0) ? s / w : L, -s * w, c);
        }
        {  // horizontal-plane focusing from ky (undulator.py:114-121)
          const double w = freq * ky;
          double s, c;
          sincos(w * L, &s, &c);
          mix_rows(M, 0, 1, c, (w != 0.0) ? s / w : L, -s * w, c);
        }
        break;
      }
      case CH_OP_CAVITY_OFF: {
        // cavity.py:253-358 at voltage == 0: alpha = 0 -> r11 = r22 = 1, r12 = L, r21 = 0,
        // r55 = r66 = 1, r65 = 0.  Standing wave: r56 = -L (Ef+Ei)/(Ef^2 Ei b1 (b1+b0)) with
        // Ef = Ei, b1 = b0; traveling wave: r56 = 0.
        const double L = slot_value(prog, s0, b);
        total_length += L;
        axpy_row(M, 0, 1, L);
        axpy_row(M, 2, 3, L);
        if (!(prog.op_flags[op] & 1)) {
          const double g = rel.gamma;
          axpy_row(M, 4, 5, -L / (g * g * g * rel.beta) * (2.0 * g) / (2.0 * rel.beta));
        }
        break;
      }
      case CH_OP_CUSTOM_MAP: {
        const ScalarRef ref = prog.slots[s0];
        if (prog.slot_begin[op + 1] - s0 >= 2) total_length += slot_value(prog, s0 + 1, b);
        double c[6][7];
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 7; ++j)
            c[i][j] = load_scalar(ref.ptr, b * ref.stride + i * 7 + j, ref.dtype);
        Map N;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 7; ++j) {
            double acc = (j == 6) ? c[i][6] : 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) acc = fma(c[i][k], M.m[k][j], acc);
            N.m[i][j] = acc;
          }
        M = N;
        break;
      }
      case CH_OP_APERTURE: {
        store_rows(aperture_rec, M, 0, 2);
        aperture_rec[14] = static_cast<T>(slot_value(prog, s0, b));
        aperture_rec[15] = static_cast<T>(slot_value(prog, s0 + 1, b));
        flags &= aperture_flags(aperture_rec) | CH_FLAG_DELTA_IDENTITY;
        aperture_rec += CH_RECORD_APERTURE;
        break;
      }
      default:
        break;
    }
  }

  T* out = rec + CH_RECORD_HEADER;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j < 7; ++j) out[i * 7 + j] = static_cast<T>(M.m[i][j]);
  flags &= map_flags(out);
  rec[0] = FlagBits<T>::pack(flags);
  rec[1] = static_cast<T>(total_length);
}

}  // namespace
}  // namespace ch

extern "C" int ch_compose_maps(const ch_program* program, int32_t op_begin, int32_t op_end,
                               int64_t n_settings, const void* energy, int64_t energy_stride,
                               int32_t energy_dtype, const void* mass_eV, int32_t mass_dtype,
                               void* records, int64_t record_len, int32_t record_dtype,
                               void* stream) {
  CH_REQUIRE(program != nullptr, "ch_compose_maps: program is NULL");
  CH_REQUIRE(op_begin >= 0 && op_begin <= op_end && op_end <= program->n_ops,
             "ch_compose_maps: op range [%d, %d) outside program of %d ops", op_begin, op_end,
             program->n_ops);
  CH_REQUIRE(n_settings > 0, "ch_compose_maps: n_settings must be positive");
  CH_REQUIRE(energy && mass_eV && records, "ch_compose_maps: NULL pointer argument");
  CH_REQUIRE(record_len >= CH_RECORD_LEN(0), "ch_compose_maps: record_len %lld too small",
             static_cast<long long>(record_len));
  CH_REQUIRE(record_dtype == CH_F32 || record_dtype == CH_F64, "ch_compose_maps: bad dtype");

  ch::Program prog{program->opcodes, program->op_flags, program->slot_begin, program->slots};
  ch::ScalarRef e{energy, energy_stride, energy_dtype};
  ch::ScalarRef m{mass_eV, 0, mass_dtype};
  const int threads = 32;
  const unsigned blocks = static_cast<unsigned>((n_settings + threads - 1) / threads);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (record_dtype == CH_F32) {
    ch::compose_maps_kernel<float><<<blocks, threads, 0, s>>>(
        prog, op_begin, op_end, n_settings, e, m, static_cast<float*>(records), record_len);
  } else {
    ch::compose_maps_kernel<double><<<blocks, threads, 0, s>>>(
        prog, op_begin, op_end, n_settings, e, m, static_cast<double*>(records), record_len);
  }
  CH_LAUNCH_CHECK();
  return CH_OK;
}

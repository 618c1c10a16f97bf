// Pieces shared by apply.cu (general kernels, C-ABI entry points) and apply_lean.cu (the kernels
// specialised for one beam under many consecutive settings).
#pragma once

#include <type_traits>

#include "ch_common.cuh"

namespace ch {

template <typename T>
struct ApplyArgs {
  const T* particles_in;
  const T* survival_in;  // may be null -> ones
  const T* records;
  T* particles_out;
  T* survival_out;  // may be null iff n_apertures == 0
  const int32_t* particle_index;
  const int32_t* survival_index;
  const int32_t* record_index;
  int64_t particle_stride;  // elements per batch entry of particles_in (0 = shared)
  int64_t survival_stride;
  int64_t record_stride;
  int64_t n_particles;
  int64_t n_settings;
  int32_t record_len;
  int32_t n_apertures;
  uint32_t elliptical_mask;
  int32_t settings_per_cta;
  int32_t bulk_in;   // particles_in tiles satisfy the 16-byte rules of cp.async.bulk
  int32_t bulk_out;  // particles_out tiles do
  double* moments_out;  // [n_settings][CH_MOMENTS] survival-weighted sums (MOMENTS kernels)
  int32_t has_cavity;   // the record ends with a CH_RECORD_CAVITY block
  int32_t covariance;   // moments_out has CH_MOMENTS_COV entries per setting (full 6x6 sums)
  // COMPACT kernels (ch_apply_maps_compact): particles_out holds 6 coordinates per row and
  // the survival mask may be written as one byte per particle instead of a T
  int32_t compact;
  uint8_t* survival_u8;
};

// apply_lean.cu
template <typename T>
bool shared_beam_call(const ApplyArgs<T>& args, bool unit_seventh);
int launch_observe_shared_beam(const ApplyArgs<float>& args, cudaStream_t stream);
template <typename T>
int launch_apply_shared_beam(const ApplyArgs<T>& args, cudaStream_t stream);

namespace {

template <typename T>
struct Vec4;
template <>
struct Vec4<float> {
  using type = float4;
  static constexpr int lanes = 4;
};
template <>
struct Vec4<double> {
  using type = double2;
  static constexpr int lanes = 2;
};

// exact IEEE helpers so that masks follow the reference's unfused elementwise ops
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ void sincos_t(float x, float& s, float& c) { sincosf(x, &s, &c); }
__device__ __forceinline__ void sincos_t(double x, double& s, double& c) { sincos(x, &s, &c); }
__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }

// load `n` scalars (n % lanes == 0, 16-byte aligned) from shared memory as 128-bit words
template <typename T, int N>
__device__ __forceinline__ void load_coefficients(T (&dst)[N], const T* src) {
  using V = typename Vec4<T>::type;
  constexpr int L = Vec4<T>::lanes;
  static_assert(N % L == 0, "coefficient block must be a whole number of 128-bit words");
#pragma unroll
  for (int i = 0; i < N / L; ++i) {
    const V v = reinterpret_cast<const V*>(src)[i];
    if constexpr (L == 4) {
      dst[4 * i + 0] = v.x;
      dst[4 * i + 1] = v.y;
      dst[4 * i + 2] = v.z;
      dst[4 * i + 3] = v.w;
    } else {
      dst[2 * i + 0] = v.x;
      dst[2 * i + 1] = v.y;
    }
  }
}

template <typename T, bool UNIT7>
__device__ __forceinline__ T affine_row(const T* c, const T (&p)[7]) {
  // c[0..6] . (p0..p5, p6) ; with UNIT7 the seventh coordinate is known to be 1
  T acc = UNIT7 ? c[6] : c[6] * p[6];
#pragma unroll
  for (int j = 5; j >= 0; --j) acc = fma_t(c[j], p[j], acc);
  return acc;
}


// the same without the tau column (CH_FLAG_NO_TAU_COLUMN: c[4] == 0)
template <typename T, bool UNIT7>
__device__ __forceinline__ T affine_row_no_tau(const T* c, const T (&p)[7]) {
  T acc = UNIT7 ? c[6] : c[6] * p[6];
  acc = fma_t(c[5], p[5], acc);
#pragma unroll
  for (int j = 3; j >= 0; --j) acc = fma_t(c[j], p[j], acc);
  return acc;
}

__device__ __forceinline__ uint32_t record_flags(float header) { return __float_as_uint(header); }
__device__ __forceinline__ uint32_t record_flags(double header) {
  return static_cast<uint32_t>(__double_as_longlong(header));
}

// |v| < bound  <=>  -bound < v < bound for bound >= 0 (NaNs compare false either way)
__device__ __forceinline__ bool inside(float v, float bound) { return fabsf(v) < bound; }
__device__ __forceinline__ bool inside(double v, double bound) { return fabs(v) < bound; }

// accumulator type of the fused moments: the beam dtype (float32 partial sums per thread about the
// pilot for float32 beams, fp64 for float64 beams -- the reference's golden dtype); fp64 across
// threads and tiles in both cases
template <typename T>
using Acc = T;

// Sum 16 per-lane values over the 32 lanes of a warp with 16 shuffles instead of 80: every
// round halves the number of values a lane is responsible for.  Afterwards lane L (L even)
// holds the warp total of value index ((L >> 4) & 1) * 8 + ((L >> 3) & 1) * 4 +
// ((L >> 2) & 1) * 2 + ((L >> 1) & 1).
template <typename A>
__device__ __forceinline__ A packed_warp_sum(A (&v)[16], int lane) {
#pragma unroll
  for (int round = 0; round < 4; ++round) {
    const int offset = 16 >> round;      // 16, 8, 4, 2
    const int keep = 8 >> round;         // 8, 4, 2, 1 values kept
    const bool upper = (lane & offset) != 0;
#pragma unroll
    for (int i = 0; i < keep; ++i) {
      const A send = upper ? v[i] : v[i + keep];
      const A mine = upper ? v[i + keep] : v[i];
      v[i] = mine + __shfl_xor_sync(0xffffffffu, send, offset);
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

// Same for 32 per-lane values (5 rounds, 31 shuffles): afterwards lane L holds the warp total of
// value index L.
template <typename A>
__device__ __forceinline__ A packed_warp_sum(A (&v)[32], int lane) {
#pragma unroll
  for (int round = 0; round < 5; ++round) {
    const int offset = 16 >> round;  // 16, 8, 4, 2, 1 = number of values kept
    const bool upper = (lane & offset) != 0;
#pragma unroll
    for (int i = 0; i < offset; ++i) {
      const A send = upper ? v[i] : v[i + offset];
      const A mine = upper ? v[i + offset] : v[i];
      v[i] = mine + __shfl_xor_sync(0xffffffffu, send, offset);
    }
  }
  return v[0];
}

using f2 = float2;
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }

// `n` duplicated coefficients (n even) from shared memory, two per 128-bit word
template <int N>
__device__ __forceinline__ void load_pairs(f2 (&dst)[N], const f2* src) {
  static_assert(N % 2 == 0, "whole 128-bit words");
#pragma unroll
  for (int i = 0; i < N / 2; ++i) {
    const float4 v = reinterpret_cast<const float4*>(src)[i];
    dst[2 * i] = f2{v.x, v.y};
    dst[2 * i + 1] = f2{v.z, v.w};
  }
}

template <bool UNIT7>
__device__ __forceinline__ f2 affine_row2(const f2* c, const f2 (&p)[7]) {
  f2 acc = UNIT7 ? c[6] : mul2(c[6], p[6]);
#pragma unroll
  for (int j = 5; j >= 0; --j) acc = fma2(c[j], p[j], acc);
  return acc;
}

}  // namespace
}  // namespace ch

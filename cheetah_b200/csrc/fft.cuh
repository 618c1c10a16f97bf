// Shared-memory FFT building blocks for the Hockney/IGF Poisson solve (power-of-two lengths).
//
// The (2n)^3 convolution of space_charge_kick.py:293-322 (rfftn . rfftn -> irfftn) is done
// as three axis passes over a [x][y][kz] complex spectrum.  Every pass keeps a tile of
// columns in shared memory; forward transforms are radix-2 decimation-in-frequency (natural
// order in, bit-reversed out) and inverse transforms decimation-in-time (bit-reversed in,
// natural out), so the fused forward -> multiply -> inverse pass along x needs no
// permutation at all.  Zero padding (7/8 of the charge array) is never materialised: columns
// are loaded with `in_len` valid entries and only `out_len` outputs are stored.
#pragma once

#include "ch_common.cuh"

namespace ch {
namespace fft {

template <typename T>
struct Complex;
template <>
struct Complex<float> {
  using type = float2;
};
template <>
struct Complex<double> {
  using type = double2;
};

template <typename C>
__device__ __forceinline__ C cmul(C a, C b) {
  return C{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename C>
__device__ __forceinline__ C cmul_conj(C a, C b) {  // a * conj(b)
  return C{a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y};
}
template <typename C>
__device__ __forceinline__ C cadd(C a, C b) {
  return C{a.x + b.x, a.y + b.y};
}
template <typename C>
__device__ __forceinline__ C csub(C a, C b) {
  return C{a.x - b.x, a.y - b.y};
}

__device__ __forceinline__ int bit_reverse(int i, int log2_len) {
  return static_cast<int>(__brev(static_cast<unsigned>(i)) >> (32 - log2_len));
}

// twiddle[k] = exp(-2 pi i k / len), k < len / 2
__device__ __forceinline__ void fill_twiddles(float2* tw, int len) {
  for (int k = threadIdx.x; k < len / 2; k += blockDim.x) {
    float s, c;
    sincospif(-2.0f * static_cast<float>(k) / static_cast<float>(len), &s, &c);
    tw[k] = float2{c, s};
  }
}
__device__ __forceinline__ void fill_twiddles(double2* tw, int len) {
  for (int k = threadIdx.x; k < len / 2; k += blockDim.x) {
    double s, c;
    sincospi(-2.0 * static_cast<double>(k) / static_cast<double>(len), &s, &c);
    tw[k] = double2{c, s};
  }
}

// Forward DIF over `columns` columns of length `len` stored as v[col * pitch + i].
// Natural-order input, bit-reversed output.  Ends with a __syncthreads().
template <typename C>
__device__ __forceinline__ void forward_dif(C* v, const C* tw, int len, int log2_len, int columns,
                                            int pitch) {
  const int half_total = len >> 1;
  for (int span_log = log2_len; span_log >= 1; --span_log) {
    const int half = 1 << (span_log - 1);
    const int tw_step = len >> span_log;
    for (int t = threadIdx.x; t < columns * half_total; t += blockDim.x) {
      const int col = t / half_total;
      const int j = t - col * half_total;
      const int pos = j & (half - 1);
      const int i0 = ((j >> (span_log - 1)) << span_log) + pos;
      C* base = v + col * pitch;
      const C a = base[i0];
      const C b = base[i0 + half];
      base[i0] = cadd(a, b);
      base[i0 + half] = cmul(csub(a, b), tw[pos * tw_step]);
    }
    __syncthreads();
  }
}

// Inverse DIT (unnormalised): bit-reversed input, natural-order output.
template <typename C>
__device__ __forceinline__ void inverse_dit(C* v, const C* tw, int len, int log2_len, int columns,
                                            int pitch) {
  const int half_total = len >> 1;
  for (int span_log = 1; span_log <= log2_len; ++span_log) {
    const int half = 1 << (span_log - 1);
    const int tw_step = len >> span_log;
    for (int t = threadIdx.x; t < columns * half_total; t += blockDim.x) {
      const int col = t / half_total;
      const int j = t - col * half_total;
      const int pos = j & (half - 1);
      const int i0 = ((j >> (span_log - 1)) << span_log) + pos;
      C* base = v + col * pitch;
      const C a = base[i0];
      const C b = cmul_conj(base[i0 + half], tw[pos * tw_step]);
      base[i0] = cadd(a, b);
      base[i0 + half] = csub(a, b);
    }
    __syncthreads();
  }
}

}  // namespace fft
}  // namespace ch

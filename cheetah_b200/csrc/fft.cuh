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

// Thread mapping of every stage: the COLUMN index runs fastest over the threads (COLS = 16 or
// 8 columns per tile), so the lanes of a half-warp hold the same butterfly of different columns.
// With an odd column pitch (len + 1 for 16 columns; len + 2 for 8 columns, where a half-warp
// covers two neighbouring butterflies) their 64-bit shared-memory accesses fall into distinct
// banks for every span, and the twiddle factor is a broadcast.  (Mapping consecutive threads to
// consecutive butterflies of ONE column -- the textbook layout -- gives 2- to 4-way bank
// conflicts for spans below 32: ncu showed the passes 95 % L1/shared-bound.)
//
// One radix-2 DIF stage of span 2^span_log on all columns (no barrier).
template <int COLS, typename C>
__device__ __forceinline__ void dif_stage(C* v, const C* tw, int len, int span_log, int pitch) {
  const int half_total = len >> 1;
  const int half = 1 << (span_log - 1);
  const int tw_step = len >> span_log;
  for (int t = threadIdx.x; t < COLS * half_total; t += blockDim.x) {
    const int col = t % COLS;
    const int j = t / COLS;
    const int pos = j & (half - 1);
    const int i0 = ((j >> (span_log - 1)) << span_log) + pos;
    C* base = v + col * pitch;
    const C a = base[i0];
    const C b = base[i0 + half];
    base[i0] = cadd(a, b);
    base[i0 + half] = cmul(csub(a, b), tw[pos * tw_step]);
  }
}

template <int COLS, typename C>
__device__ __forceinline__ void dit_stage(C* v, const C* tw, int len, int span_log, int pitch) {
  const int half_total = len >> 1;
  const int half = 1 << (span_log - 1);
  const int tw_step = len >> span_log;
  for (int t = threadIdx.x; t < COLS * half_total; t += blockDim.x) {
    const int col = t % COLS;
    const int j = t / COLS;
    const int pos = j & (half - 1);
    const int i0 = ((j >> (span_log - 1)) << span_log) + pos;
    C* base = v + col * pitch;
    const C a = base[i0];
    const C b = cmul_conj(base[i0 + half], tw[pos * tw_step]);
    base[i0] = cadd(a, b);
    base[i0 + half] = csub(a, b);
  }
}

// Forward DIF over `columns` columns of length `len` stored as v[col * pitch + i].
// Natural-order input, bit-reversed output.  Two consecutive radix-2 stages (spans 2^s and
// 2^(s-1)) act on closed groups {i, i+q, i+2q, i+3q}, so each thread carries four points
// through both stages in registers: half the shared-memory round trips and barriers of a
// plain radix-2 loop with the identical data flow (the output order stays bit-reversed).
// Ends with a __syncthreads().
template <int COLS, typename C>
__device__ __forceinline__ void forward_dif(C* v, const C* tw, int len, int log2_len, int pitch) {
  const int quarter_total = len >> 2;
  int s = log2_len;
  for (; s >= 2; s -= 2) {
    const int quarter = 1 << (s - 2);
    const int step_a = len >> s;
    for (int t = threadIdx.x; t < COLS * quarter_total; t += blockDim.x) {
      const int col = t % COLS;
      const int j = t / COLS;
      const int pos = j & (quarter - 1);
      const int i0 = ((j >> (s - 2)) << s) + pos;
      C* base = v + col * pitch;
      const C x0 = base[i0], x1 = base[i0 + quarter], x2 = base[i0 + 2 * quarter],
              x3 = base[i0 + 3 * quarter];
      const C a0 = cadd(x0, x2), a2 = cmul(csub(x0, x2), tw[pos * step_a]);
      const C a1 = cadd(x1, x3), a3 = cmul(csub(x1, x3), tw[(pos + quarter) * step_a]);
      const C wb = tw[pos * 2 * step_a];
      base[i0] = cadd(a0, a1);
      base[i0 + quarter] = cmul(csub(a0, a1), wb);
      base[i0 + 2 * quarter] = cadd(a2, a3);
      base[i0 + 3 * quarter] = cmul(csub(a2, a3), wb);
    }
    __syncthreads();
  }
  if (s == 1) {
    dif_stage<COLS>(v, tw, len, 1, pitch);
    __syncthreads();
  }
}

// Inverse DIT (unnormalised): bit-reversed input, natural-order output; mirror image of
// forward_dif (stage pairs fused in registers).
template <int COLS, typename C>
__device__ __forceinline__ void inverse_dit(C* v, const C* tw, int len, int log2_len, int pitch) {
  const int quarter_total = len >> 2;
  int s = 2;
  if (log2_len & 1) {
    dit_stage<COLS>(v, tw, len, 1, pitch);
    __syncthreads();
    s = 3;
  }
  for (; s <= log2_len; s += 2) {
    const int quarter = 1 << (s - 2);
    const int step_a = len >> s;
    for (int t = threadIdx.x; t < COLS * quarter_total; t += blockDim.x) {
      const int col = t % COLS;
      const int j = t / COLS;
      const int pos = j & (quarter - 1);
      const int i0 = ((j >> (s - 2)) << s) + pos;
      C* base = v + col * pitch;
      const C x0 = base[i0], x2 = base[i0 + 2 * quarter];
      const C wb = tw[pos * 2 * step_a];
      const C b1 = cmul_conj(base[i0 + quarter], wb);
      const C b3 = cmul_conj(base[i0 + 3 * quarter], wb);
      const C y0 = cadd(x0, b1), y1 = csub(x0, b1), y2 = cadd(x2, b3), y3 = csub(x2, b3);
      const C c2 = cmul_conj(y2, tw[pos * step_a]);
      const C c3 = cmul_conj(y3, tw[(pos + quarter) * step_a]);
      base[i0] = cadd(y0, c2);
      base[i0 + 2 * quarter] = csub(y0, c2);
      base[i0 + quarter] = cadd(y1, c3);
      base[i0 + 3 * quarter] = csub(y1, c3);
    }
    __syncthreads();
  }
}

}  // namespace fft
}  // namespace ch

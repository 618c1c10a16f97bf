// Register-resident FFT core for the Poisson solve (float, lengths 32 .. 256).
//
// A length-LEN transform (LEN = 16 N2) is shared by N2 threads holding 16 points each:
//   n = N2 n1 + n2,  k = k1 + 16 k2      (n1, k1 < 16;  n2, k2 < N2)
//   X[k1 + 16 k2] = sum_n2 W_LEN^(n2 k1) [ sum_n1 x[N2 n1 + n2] W_16^(n1 k1) ] W_N2^(n2 k2)
// pass 1: thread n2 transforms its 16 points (radix-16 in registers) and applies W_LEN^(n2 k1);
// one exchange through shared memory; pass 2: thread u takes k1 = u + N2 m (m < 16 / N2) and
// does 16 / N2 transforms of length N2.  Input and output are both in natural order and in the
// SAME distribution: thread u holds element u + N2 j in slot j -- so forward -> multiply ->
// inverse chains need no permutation and global memory is touched straight from registers.
// Compared with the shared-memory radix-4 passes of fft.cuh (4 round trips per 128-point
// transform, ~110 instructions per point) this is one round trip and ~30 instructions per point.
//
// The butterflies are a compile-time unrolled decimation-in-time recursion; rotations by
// multiples of 45 degrees are special-cased.  Everything is __host__ __device__ so that
// tests/native/fft_regs_test.cu can check the index algebra on the CPU against a naive DFT.
#pragma once

#include <cuda_runtime.h>

#ifndef CH_HD
#define CH_HD __host__ __device__ __forceinline__
#endif

namespace ch {
namespace fftr {

using C = float2;

CH_HD C add(C a, C b) { return C{a.x + b.x, a.y + b.y}; }
CH_HD C sub(C a, C b) { return C{a.x - b.x, a.y - b.y}; }
CH_HD C mul(C a, C b) { return C{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
CH_HD C mul_conj(C a, C b) { return C{a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y}; }  // a conj(b)

// v * exp(-+ 2 pi i q / 16) (forward: minus) for a q known at compile time after unrolling
template <bool INV>
CH_HD C rotate16(C v, int q) {
  constexpr float kCos[16] = {1.0f, 0.92387953251128674f, 0.70710678118654752f,
                              0.38268343236508977f, 0.0f, -0.38268343236508977f,
                              -0.70710678118654752f, -0.92387953251128674f, -1.0f,
                              -0.92387953251128674f, -0.70710678118654752f,
                              -0.38268343236508977f, 0.0f, 0.38268343236508977f,
                              0.70710678118654752f, 0.92387953251128674f};
  constexpr float kSin[16] = {0.0f, 0.38268343236508977f, 0.70710678118654752f,
                              0.92387953251128674f, 1.0f, 0.92387953251128674f,
                              0.70710678118654752f, 0.38268343236508977f, 0.0f,
                              -0.38268343236508977f, -0.70710678118654752f,
                              -0.92387953251128674f, -1.0f, -0.92387953251128674f,
                              -0.70710678118654752f, -0.38268343236508977f};
  q &= 15;
  if (INV) q = (16 - q) & 15;  // exp(+i a) = exp(-i (2 pi - a))
  // now multiply by exp(-2 pi i q / 16) = cos - i sin
  if (q == 0) return v;
  if (q == 4) return C{v.y, -v.x};
  if (q == 8) return C{-v.x, -v.y};
  if (q == 12) return C{-v.y, v.x};
  constexpr float h = 0.70710678118654752f;
  if (q == 2) return C{(v.x + v.y) * h, (v.y - v.x) * h};
  if (q == 6) return C{(v.y - v.x) * h, -(v.x + v.y) * h};
  if (q == 10) return C{-(v.x + v.y) * h, (v.x - v.y) * h};
  if (q == 14) return C{(v.x - v.y) * h, (v.x + v.y) * h};
  const float c = kCos[q], s = kSin[q];
  return C{v.x * c + v.y * s, v.y * c - v.x * s};
}

// out[k] = sum_n in[OFFSET + n STRIDE] exp(-+ 2 pi i n k / R), natural order (R <= 16)
template <int R, bool INV>
struct Dft {
  template <int STRIDE, int OFFSET, int TOTAL>
  static CH_HD void run(const C (&in)[TOTAL], C (&out)[R]) {
    C even[R / 2], odd[R / 2];
    Dft<R / 2, INV>::template run<2 * STRIDE, OFFSET, TOTAL>(in, even);
    Dft<R / 2, INV>::template run<2 * STRIDE, OFFSET + STRIDE, TOTAL>(in, odd);
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
      const C t = rotate16<INV>(odd[k], k * (16 / R));
      out[k] = add(even[k], t);
      out[k + R / 2] = sub(even[k], t);
    }
  }
};
template <bool INV>
struct Dft<1, INV> {
  template <int STRIDE, int OFFSET, int TOTAL>
  static CH_HD void run(const C (&in)[TOTAL], C (&out)[1]) {
    out[0] = in[OFFSET];
  }
};

template <int LEN>
struct Plan {
  static_assert(LEN == 32 || LEN == 64 || LEN == 128 || LEN == 256, "register FFT: 32 .. 256");
  static constexpr int N2 = LEN / 16;       // threads per transform
  static constexpr int M = 16 / N2;         // pass-2 transforms per thread
  static constexpr int PITCH = LEN + 1;     // complex words per column of the exchange buffer
};

// twiddle table: tw[k] = exp(-2 pi i k / LEN), k < LEN (shared memory, filled once per CTA)
template <int LEN>
__device__ __forceinline__ void fill_twiddles(C* tw) {
  for (int k = threadIdx.x; k < LEN; k += blockDim.x) {
    float s, c;
    sincospif(-2.0f * static_cast<float>(k) / static_cast<float>(LEN), &s, &c);
    tw[k] = C{c, s};
  }
}

// Pass 1 on the 16 points of thread n2: v[k1] <- W_LEN^(n2 k1) * DFT16(v)[k1].
template <int LEN, bool INV>
CH_HD void pass1(C (&v)[16], int n2, const C* tw) {
  C y[16];
  Dft<16, INV>::template run<1, 0, 16>(v, y);
  v[0] = y[0];
#pragma unroll
  for (int k1 = 1; k1 < 16; ++k1) {
    const C w = tw[n2 * k1];  // n2 k1 <= 15 (N2 - 1) < LEN
    v[k1] = INV ? mul_conj(y[k1], w) : mul(y[k1], w);
  }
}

// Pass 2 of thread u on z[m][n2] = Z[n2][k1 = u + N2 m]: v[m + M k2] <- X[u + N2 (m + M k2)].
template <int LEN, bool INV>
CH_HD void pass2(const C (&z)[16], C (&v)[16]) {
  constexpr int N2 = Plan<LEN>::N2, M = Plan<LEN>::M;
#pragma unroll
  for (int m = 0; m < M; ++m) {
    C in[N2], out[N2];
#pragma unroll
    for (int n2 = 0; n2 < N2; ++n2) in[n2] = z[m * N2 + n2];
    Dft<N2, INV>::template run<1, 0, N2>(in, out);
#pragma unroll
    for (int k2 = 0; k2 < N2; ++k2) v[m + M * k2] = out[k2];
  }
}

// Whole transform of one column in two halves around a CTA-wide barrier.  `column` = this
// column's PITCH words of exchange space; the N2 threads of a column call both halves with
// their n2.      in : v[n1] = x[N2 n1 + n2]        out: v[j] = X[n2 + N2 j]
template <int LEN, bool INV>
CH_HD void transform_scatter(C (&v)[16], C* column, int n2, const C* tw) {
  pass1<LEN, INV>(v, n2, tw);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) column[n2 * 16 + k1] = v[k1];
}
template <int LEN, bool INV>
CH_HD void transform_gather(C (&v)[16], const C* column, int n2) {
  constexpr int N2 = Plan<LEN>::N2, M = Plan<LEN>::M;
  C z[16];
#pragma unroll
  for (int m = 0; m < M; ++m)
#pragma unroll
    for (int s = 0; s < N2; ++s) z[m * N2 + s] = column[s * 16 + n2 + N2 * m];
  pass2<LEN, INV>(z, v);
}
#ifdef __CUDACC__
template <int LEN, bool INV>
__device__ __forceinline__ void transform(C (&v)[16], C* column, int n2, const C* tw) {
  transform_scatter<LEN, INV>(v, column, n2, tw);
  __syncthreads();
  transform_gather<LEN, INV>(v, column, n2);
}
#endif

// Layout B of the exchange, for passes along the CONTIGUOUS axis: there the N2 threads of a
// transform are adjacent lanes (n2 in the low bits of threadIdx.x), so that a warp reads N2
// consecutive elements of each of its 32 / N2 rows straight from global memory.  The exchange
// word of (n2, k1) then sits at k1 (N2 + 1) + n2 and columns are PITCH_B = 16 (N2 + 1) + N2
// words apart (= N2 mod 16): scatter (fixed k1) and gather (lane u reads word
// (u + N2 m)(N2 + 1) + s) are both bank-conflict free for N2 = 2, 4, 8, 16.
template <int LEN>
struct PlanB {
  static constexpr int N2 = LEN / 16;
  static constexpr int PITCH = 16 * (N2 + 1) + N2;
};
template <int LEN, bool INV>
CH_HD void transform_scatter_b(C (&v)[16], C* column, int n2, const C* tw) {
  constexpr int N2 = Plan<LEN>::N2;
  pass1<LEN, INV>(v, n2, tw);
#pragma unroll
  for (int k1 = 0; k1 < 16; ++k1) column[k1 * (N2 + 1) + n2] = v[k1];
}
template <int LEN, bool INV>
CH_HD void transform_gather_b(C (&v)[16], const C* column, int n2) {
  constexpr int N2 = Plan<LEN>::N2, M = Plan<LEN>::M;
  C z[16];
#pragma unroll
  for (int m = 0; m < M; ++m)
#pragma unroll
    for (int s = 0; s < N2; ++s) z[m * N2 + s] = column[(n2 + N2 * m) * (N2 + 1) + s];
  pass2<LEN, INV>(z, v);
}
#ifdef __CUDACC__
template <int LEN, bool INV>
__device__ __forceinline__ void transform_b(C (&v)[16], C* column, int n2, const C* tw) {
  transform_scatter_b<LEN, INV>(v, column, n2, tw);
  __syncthreads();
  transform_gather_b<LEN, INV>(v, column, n2);
}
#endif

}  // namespace fftr
}  // namespace ch

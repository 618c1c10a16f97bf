// Segment.track / Segment.track_moments of ONE beam under many consecutive settings (the ARES
// x 4096 settings case): kernels specialised at compile time on the number of apertures and on
// whether any of them is elliptical.  Same arithmetic and the same fma chains as the general
// kernels of apply.cu (aperture masks are bit-identical); what goes away is the bookkeeping per
// setting.  ch_apply_maps* dispatch here when shared_beam_call() holds.
#include "apply_common.cuh"

namespace ch {
namespace {

// ---- observables only ---------------------------------------------------------------------------
// Segment.track_moments of ONE beam under many settings (shared beam and incoming survival,
// consecutive records, unit seventh column, at most three apertures: the ARES case).  Same
// arithmetic, same chains and the same masks as observe_maps_kernel (apply.cu); what goes away is the
// bookkeeping per setting: the aperture count and the record length are compile-time constants
// (records live in a static shared array, every LDS has an immediate offset, the aperture block
// is unrolled and its selects overlap the next block's FFMA2s), the beam is loaded once before
// the loop, the record pointer advances by one addition and the masks use a chained predicate
// (2 FSETP + 1 FSEL per particle and aperture).
__device__ __forceinline__ float keep_if_inside(float sv, float x, float x_max, float y,
                                                float y_max) {
  float out;
  asm("{\n"
      "  .reg .pred p;\n"
      "  .reg .f32 ax, ay;\n"
      "  abs.f32 ax, %1;\n"
      "  abs.f32 ay, %3;\n"
      "  setp.lt.f32 p, ax, %2;\n"
      "  setp.lt.and.f32 p, ay, %4, p;\n"
      "  selp.f32 %0, %5, 0f00000000, p;\n"
      "}"
      : "=f"(out)
      : "f"(x), "f"(x_max), "f"(y), "f"(y_max), "f"(sv));
  return out;
}

// c[0..3] . p[0..3] + c[5] p[5] + c[6]: a row without tau dependence (CH_FLAG_NO_TAU_COLUMN), the
// chain of affine_row_no_tau
__device__ __forceinline__ f2 affine_row_no_tau2(const f2* c, const f2 (&p)[7], f2 constant) {
  f2 acc = fma2(c[5], p[5], constant);
#pragma unroll
  for (int j = 3; j >= 0; --j) acc = fma2(c[j], p[j], acc);
  return acc;
}

// MODE as in process_setting (apply.cu): 0 dense, 1 sparse, 2 coupled (x-y coupling and dispersion
// allowed, no tau column in rows 0-3, delta untouched)
template <int NAP, int MODE, int MOMENTS, bool ELLIPTICAL>
__device__ __forceinline__ void observe_setting_lean(const f2* rec2, uint32_t elliptical_mask,
                                                     const f2 (&p)[4][7], f2 (&sv)[4],
                                                     const float (&first_particle)[7],
                                                     float (&pilot)[6],
                                                     f2 (&acc)[MOMENTS == 2 ? 29 : 14]) {
  constexpr int PAIRS = 4;
#pragma unroll
  for (int ap = 0; ap < NAP; ++ap) {
    f2 q[16];
    load_pairs(q, rec2 + CH_RECORD_HEADER + CH_RECORD_MAP + ap * CH_RECORD_APERTURE);
    const float x_max = q[14].x, y_max = q[15].x;
    f2 x[PAIRS], y[PAIRS];
#pragma unroll
    for (int k = 0; k < PAIRS; ++k) {
      if constexpr (MODE == 1) {
        x[k] = fma2(q[0], p[k][0], fma2(q[1], p[k][1], fma2(q[5], p[k][5], q[6])));
        y[k] = fma2(q[9], p[k][2], fma2(q[10], p[k][3], q[13]));
      } else if constexpr (MODE == 2) {
        x[k] = affine_row_no_tau2(q, p[k], q[6]);
        y[k] = affine_row_no_tau2(q + 7, p[k], q[13]);
      } else {
        x[k] = affine_row2<true>(q, p[k]);
        y[k] = affine_row2<true>(q + 7, p[k]);
      }
    }
    // ELLIPTICAL = false: no aperture of the section is elliptical, the whole setting is one
    // basic block and the selects of one aperture are scheduled between the FFMA2s of the next
    if (ELLIPTICAL && ((elliptical_mask >> ap) & 1u)) {
      const float xx = mul_rn(x_max, x_max), yy = mul_rn(y_max, y_max);
#pragma unroll
      for (int k = 0; k < PAIRS; ++k) {
        const bool lo = add_rn(div_rn(mul_rn(x[k].x, x[k].x), xx),
                               div_rn(mul_rn(y[k].x, y[k].x), yy)) <= 1.0f;
        const bool hi = add_rn(div_rn(mul_rn(x[k].y, x[k].y), xx),
                               div_rn(mul_rn(y[k].y, y[k].y), yy)) <= 1.0f;
        sv[k].x = lo ? sv[k].x : 0.0f;
        sv[k].y = hi ? sv[k].y : 0.0f;
      }
    } else {
#pragma unroll
      for (int k = 0; k < PAIRS; ++k) {
        sv[k].x = keep_if_inside(sv[k].x, x[k].x, x_max, y[k].x, y_max);
        sv[k].y = keep_if_inside(sv[k].y, x[k].y, x_max, y[k].y, y_max);
      }
    }
  }

  f2 c[44];
  load_pairs(c, rec2);
  const f2* m = c + CH_RECORD_HEADER;
  {
    const float(&in)[7] = first_particle;
    if constexpr (MODE == 1) {
      pilot[0] = fmaf(m[0].x, in[0], fmaf(m[1].x, in[1], fmaf(m[5].x, in[5], m[6].x)));
      pilot[1] = fmaf(m[7].x, in[0], fmaf(m[8].x, in[1], fmaf(m[12].x, in[5], m[13].x)));
      pilot[2] = fmaf(m[16].x, in[2], fmaf(m[17].x, in[3], m[20].x));
      pilot[3] = fmaf(m[23].x, in[2], fmaf(m[24].x, in[3], m[27].x));
      pilot[4] = fmaf(m[28].x, in[0],
                      fmaf(m[29].x, in[1], fmaf(m[32].x, in[4], fmaf(m[33].x, in[5], m[34].x))));
      pilot[5] = in[5];
    } else if constexpr (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float acc1 = fmaf(m[i * 7 + 5].x, in[5], m[i * 7 + 6].x);
#pragma unroll
        for (int j = 3; j >= 0; --j) acc1 = fmaf(m[i * 7 + j].x, in[j], acc1);
        pilot[i] = acc1;
      }
      float acc4 = m[34].x;
#pragma unroll
      for (int j = 5; j >= 0; --j) acc4 = fmaf(m[28 + j].x, in[j], acc4);
      pilot[4] = acc4;
      pilot[5] = in[5];
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        float acc1 = m[i * 7 + 6].x;
#pragma unroll
        for (int j = 5; j >= 0; --j) acc1 = fmaf(m[i * 7 + j].x, in[j], acc1);
        pilot[i] = acc1;
      }
    }
  }
  // Distances from the pilot come straight out of the fma chains: the constant of row i is
  // replaced by (constant - pilot_i), so the chain ends on u_i - c_i instead of u_i (one rounding
  // less than subtracting afterwards, and 20 packed additions fewer per setting).
  f2 shifted[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    const float c = m[i * 7 + 6].x - pilot[i];
    shifted[i] = f2{c, c};
  }
#pragma unroll
  for (int k = 0; k < PAIRS; ++k) {
    f2 d[6];
    if constexpr (MODE == 1) {
      d[0] = fma2(m[0], p[k][0], fma2(m[1], p[k][1], fma2(m[5], p[k][5], shifted[0])));
      d[1] = fma2(m[7], p[k][0], fma2(m[8], p[k][1], fma2(m[12], p[k][5], shifted[1])));
      d[2] = fma2(m[16], p[k][2], fma2(m[17], p[k][3], shifted[2]));
      d[3] = fma2(m[23], p[k][2], fma2(m[24], p[k][3], shifted[3]));
      d[4] = fma2(m[28], p[k][0],
                  fma2(m[29], p[k][1], fma2(m[32], p[k][4], fma2(m[33], p[k][5], shifted[4]))));
      d[5] = add2(p[k][5], f2{-pilot[5], -pilot[5]});
    } else if constexpr (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) d[i] = affine_row_no_tau2(m + i * 7, p[k], shifted[i]);
      {
        f2 acc1 = shifted[4];
#pragma unroll
        for (int j = 5; j >= 0; --j) acc1 = fma2(m[28 + j], p[k][j], acc1);
        d[4] = acc1;
      }
      d[5] = add2(p[k][5], f2{-pilot[5], -pilot[5]});
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        f2 acc1 = shifted[i];
#pragma unroll
        for (int j = 5; j >= 0; --j) acc1 = fma2(m[i * 7 + j], p[k][j], acc1);
        d[i] = acc1;
      }
    }
    const f2 w = sv[k];
    f2 wd[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) wd[i] = mul2(w, d[i]);
    if (k == 0) {  // first pair: the sums start here (no zero-filled accumulators)
      acc[0] = w;
      acc[1] = mul2(w, w);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        acc[2 + i] = wd[i];
        acc[8 + i] = mul2(wd[i], d[i]);
      }
    } else {
      acc[0] = add2(acc[0], w);
      acc[1] = fma2(w, w, acc[1]);
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        acc[2 + i] = add2(acc[2 + i], wd[i]);
        acc[8 + i] = fma2(wd[i], d[i], acc[8 + i]);
      }
    }
    if constexpr (MOMENTS == 2) {
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = i + 1; j < 6; ++j) {
          const int slot = 14 + i * (11 - i) / 2 + (j - i - 1);
          acc[slot] = k == 0 ? mul2(wd[i], d[j]) : fma2(wd[i], d[j], acc[slot]);
        }
    }
  }
}

constexpr int kLeanThreads = 128, kLeanP = 8;

template <int NAP, int MOMENTS, bool ELLIPTICAL>
__global__ void __launch_bounds__(kLeanThreads, MOMENTS == 2 ? (ELLIPTICAL ? 2 : 3) : 4)
observe_shared_beam_kernel(const ApplyArgs<float> a) {
  constexpr int THREADS = kLeanThreads, PAIRS = kLeanP / 2, TP = kLeanP * THREADS;
  constexpr int RECLEN = CH_RECORD_LEN(NAP);
  constexpr int NACC = MOMENTS == 2 ? 32 : 16;
  constexpr int NSUM = MOMENTS == 2 ? 29 : 14;
  constexpr int NOUT = MOMENTS == 2 ? CH_MOMENTS_COV : CH_MOMENTS;
  static_assert(RECLEN % 2 == 0 && RECLEN <= THREADS, "one record entry per thread");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(16) f2 recs[2][RECLEN];
  __shared__ float partial[2][THREADS / 32][NACC];
  __shared__ float pilot_shared[2][8];
  __shared__ uint64_t bar;

  const int tid = threadIdx.x;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * TP;
  const int count = static_cast<int>(min(static_cast<int64_t>(TP), a.n_particles - n0));
  const int64_t b_begin = static_cast<int64_t>(blockIdx.y) * a.settings_per_cta;
  const int n_local =
      static_cast<int>(min(a.n_settings - b_begin, static_cast<int64_t>(a.settings_per_cta)));
  if (n_local <= 0) return;

  // ---- the beam: once per CTA (before the wait for the compose kernel: see below) -----------
  {
    const float* src = a.particles_in + n0 * 7;
    if (a.bulk_in) {
      if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        const uint32_t bytes = static_cast<uint32_t>(count) * 7u * sizeof(float);
        mbar_expect_tx(&bar, bytes);
        bulk_load(tile, src, bytes, &bar);
      }
      __syncthreads();  // the barrier is initialised before anybody waits on it
      mbar_wait(&bar, 0);
    } else {
      for (int i = tid; i < count * 7; i += THREADS) tile[i] = src[i];
      __syncthreads();
    }
  }
  f2 p[PAIRS][7], sv_in[PAIRS];
#pragma unroll
  for (int k = 0; k < PAIRS; ++k) {
    const int lo = tid + (2 * k) * THREADS, hi = lo + THREADS;
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      p[k][j].x = lo < count ? tile[lo * 7 + j] : 0.0f;
      p[k][j].y = hi < count ? tile[hi * 7 + j] : 0.0f;
    }
    // lanes past the end of the beam must not count
    sv_in[k].x = lo < count ? (a.survival_in ? a.survival_in[n0 + lo] : 1.0f) : 0.0f;
    sv_in[k].y = hi < count ? (a.survival_in ? a.survival_in[n0 + hi] : 1.0f) : 0.0f;
  }
  float first[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) first[j] = a.particles_in[j];
  // the records come from the compose kernel launched just before: this kernel may have started
  // while that one was still running
  grid_dependency_wait();
  const float* rec_src = a.records + b_begin * a.record_stride;
  float fetched = tid < RECLEN ? rec_src[tid] : 0.0f;
  if (tid < RECLEN) recs[0][tid] = f2{fetched, fetched};

  double* out = a.moments_out + b_begin * NOUT;
  auto flush_moments = [&](int buf, double* dst) {
    if (tid < NSUM) {
      double total = 0.0;
#pragma unroll
      for (int wi = 0; wi < THREADS / 32; ++wi) total += static_cast<double>(partial[buf][wi][tid]);
      atomicAdd(dst + (tid < 14 ? tid : tid + 6), total);
    } else if (tid >= 32 && tid < 38 && blockIdx.x == 0) {
      dst[14 + (tid - 32)] = static_cast<double>(pilot_shared[buf][tid - 32]);
    }
  };
  constexpr uint32_t kSparse = CH_FLAG_XY_UNCOUPLED | CH_FLAG_NO_TAU_COLUMN |
                               CH_FLAG_NO_Y_DISPERSION | CH_FLAG_DELTA_IDENTITY;
  constexpr uint32_t kCoupled = CH_FLAG_NO_TAU_COLUMN | CH_FLAG_DELTA_IDENTITY;
  const int lane = tid & 31;
  const int slot = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 +
                   ((lane >> 1) & 1);

#pragma unroll 1
  for (int it = 0; it < n_local; ++it) {
    const int buf = it & 1;
    __syncthreads();  // record `it` is complete; so is partial[buf ^ 1]
    if (it > 0) {
      flush_moments(buf ^ 1, out);
      out += NOUT;
    }
    if (it + 1 < n_local) {  // in flight during the arithmetic below
      rec_src += a.record_stride;
      fetched = tid < RECLEN ? rec_src[tid] : 0.0f;
    }
    const f2* rec2 = recs[buf];
    f2 sv[PAIRS];
#pragma unroll
    for (int k = 0; k < PAIRS; ++k) sv[k] = sv_in[k];
    float pilot[6];
    f2 acc2[NSUM];
    const uint32_t flags = record_flags(rec2[0].x);
    if ((flags & kSparse) == kSparse)
      observe_setting_lean<NAP, 1, MOMENTS, ELLIPTICAL>(rec2, a.elliptical_mask, p, sv, first,
                                                        pilot, acc2);
    else if ((flags & kCoupled) == kCoupled)
      observe_setting_lean<NAP, 2, MOMENTS, ELLIPTICAL>(rec2, a.elliptical_mask, p, sv, first,
                                                        pilot, acc2);
    else
      observe_setting_lean<NAP, 0, MOMENTS, ELLIPTICAL>(rec2, a.elliptical_mask, p, sv, first,
                                                        pilot, acc2);
    if (it + 1 < n_local && tid < RECLEN) recs[buf ^ 1][tid] = f2{fetched, fetched};
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < NSUM; ++i) acc[i] = acc2[i].x + acc2[i].y;
    const float total = packed_warp_sum(acc, lane);
    if constexpr (MOMENTS == 2) {
      partial[buf][tid >> 5][lane] = total;
    } else if ((lane & 1) == 0) {
      partial[buf][tid >> 5][slot] = total;
    }
    if (tid == 0) {
#pragma unroll
      for (int i = 0; i < 6; ++i) pilot_shared[buf][i] = pilot[i];
    }
  }
  __syncthreads();
  flush_moments((n_local - 1) & 1, out);
}


// ---- particles out --------------------------------------------------------------------------
// apply_maps_kernel for the same call shape (float32, unit seventh column, rows of 7, no cavity):
// P = 4 particles per thread, 256 threads.  The sparse 29-FMA branch of the general kernel is
// HBM-bound already; the coupled (56) and dense (72) branches are issue-bound there (155
// instructions per particle and setting for 72 multiply-adds).  Here the aperture blocks are
// unrolled, records sit in a static shared array, and ONE barrier per setting serves three
// purposes: the rows of this setting are complete (the elected thread hands the tile to the TMA
// engine right after it), the next record is committed, and the tile of the previous setting has
// been read by the engine (the elected thread waits for that just before the barrier).
constexpr int kApplyLeanThreads = 256;
// particles per thread: tiles of 1024 float32 / 512 float64 particles (28 KB either way)
template <typename T>
constexpr int kApplyLeanP = sizeof(T) == 4 ? 4 : 2;

__device__ __forceinline__ double keep_if_inside(double sv, double x, double x_max, double y,
                                                 double y_max) {
  return (inside(x, x_max) && inside(y, y_max)) ? sv : 0.0;
}

template <typename T, int NAP, int MODE, bool ELLIPTICAL>
__device__ __forceinline__ void apply_setting_lean(const T* rec, uint32_t elliptical_mask,
                                                   const T (&p)[kApplyLeanP<T>][7],
                                                   T (&sv)[kApplyLeanP<T>], T* stage, int tid) {
  constexpr int P = kApplyLeanP<T>, THREADS = kApplyLeanThreads;
#pragma unroll
  for (int ap = 0; ap < NAP; ++ap) {
    T q[16];
    load_coefficients(q, rec + CH_RECORD_HEADER + CH_RECORD_MAP + ap * CH_RECORD_APERTURE);
    const T x_max = q[14], y_max = q[15];
#pragma unroll
    for (int k = 0; k < P; ++k) {
      T x, y;
      if constexpr (MODE == 1) {
        x = fma_t(q[0], p[k][0], fma_t(q[1], p[k][1], fma_t(q[5], p[k][5], q[6])));
        y = fma_t(q[9], p[k][2], fma_t(q[10], p[k][3], q[13]));
      } else if constexpr (MODE == 2) {
        x = affine_row_no_tau<T, true>(q, p[k]);
        y = affine_row_no_tau<T, true>(q + 7, p[k]);
      } else {
        x = affine_row<T, true>(q, p[k]);
        y = affine_row<T, true>(q + 7, p[k]);
      }
      if (ELLIPTICAL && ((elliptical_mask >> ap) & 1u)) {
        const T ex = div_rn(mul_rn(x, x), mul_rn(x_max, x_max));
        const T ey = div_rn(mul_rn(y, y), mul_rn(y_max, y_max));
        sv[k] = add_rn(ex, ey) <= T(1) ? sv[k] : T(0);
      } else {
        sv[k] = keep_if_inside(sv[k], x, x_max, y, y_max);
      }
    }
  }
  T c[44];
  load_coefficients(c, rec);
  const T* m = c + CH_RECORD_HEADER;
#pragma unroll
  for (int k = 0; k < P; ++k) {
    T* row = stage + (tid + k * THREADS) * 7;
    if constexpr (MODE == 1) {
      row[0] = fma_t(m[0], p[k][0], fma_t(m[1], p[k][1], fma_t(m[5], p[k][5], m[6])));
      row[1] = fma_t(m[7], p[k][0], fma_t(m[8], p[k][1], fma_t(m[12], p[k][5], m[13])));
      row[2] = fma_t(m[16], p[k][2], fma_t(m[17], p[k][3], m[20]));
      row[3] = fma_t(m[23], p[k][2], fma_t(m[24], p[k][3], m[27]));
      row[4] = fma_t(m[28], p[k][0],
                    fma_t(m[29], p[k][1], fma_t(m[32], p[k][4], fma_t(m[33], p[k][5], m[34]))));
      row[5] = p[k][5];
    } else if constexpr (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) row[i] = affine_row_no_tau<T, true>(m + i * 7, p[k]);
      row[4] = affine_row<T, true>(m + 28, p[k]);
      row[5] = p[k][5];
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) row[i] = affine_row<T, true>(m + i * 7, p[k]);
    }
    row[6] = T(1);
  }
}

template <typename T, int NAP, bool ELLIPTICAL>
__global__ void __launch_bounds__(kApplyLeanThreads, sizeof(T) == 4 ? 3 : 2)
apply_shared_beam_kernel(const ApplyArgs<T> a) {
  constexpr int THREADS = kApplyLeanThreads, P = kApplyLeanP<T>, TP = P * THREADS;
  constexpr int RECLEN = CH_RECORD_LEN(NAP);
  static_assert(RECLEN % 4 == 0 && RECLEN <= THREADS, "one record entry per thread");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* stage0 = reinterpret_cast<T*>(smem_raw);
  T* stage1 = stage0 + TP * 7;
  __shared__ __align__(16) T recs[2][RECLEN];
  __shared__ uint64_t bar;

  const int tid = threadIdx.x;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * TP;
  const int count = static_cast<int>(min(static_cast<int64_t>(TP), a.n_particles - n0));
  const int64_t b_begin = static_cast<int64_t>(blockIdx.y) * a.settings_per_cta;
  const int n_local =
      static_cast<int>(min(a.n_settings - b_begin, static_cast<int64_t>(a.settings_per_cta)));
  if (n_local <= 0) return;

  {
    const T* src = a.particles_in + n0 * 7;
    if (a.bulk_in) {
      if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
        const uint32_t bytes = static_cast<uint32_t>(count) * 7u * sizeof(T);
        mbar_expect_tx(&bar, bytes);
        bulk_load(stage1, src, bytes, &bar);
      }
      __syncthreads();
      mbar_wait(&bar, 0);
    } else {
      for (int i = tid; i < count * 7; i += THREADS) stage1[i] = src[i];
      __syncthreads();
    }
  }
  T p[P][7], sv_in[P];
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const int local = tid + k * THREADS;
#pragma unroll
    for (int j = 0; j < 7; ++j) p[k][j] = local < count ? stage1[local * 7 + j] : T(0);
    sv_in[k] = (a.survival_in != nullptr && local < count) ? a.survival_in[n0 + local] : T(1);
  }
  // the records come from the compose kernel launched just before (see the observables kernel)
  grid_dependency_wait();
  const T* rec_src = a.records + b_begin * a.record_stride;
  T fetched = tid < RECLEN ? rec_src[tid] : T(0);
  if (tid < RECLEN) recs[0][tid] = fetched;
  __syncthreads();  // record 0 complete; everybody holds its particles (stage1 is reused later)

  T* out = a.particles_out + (b_begin * a.n_particles + n0) * 7;
  T* survival_out =
      a.survival_out != nullptr ? a.survival_out + b_begin * a.n_particles + n0 : nullptr;
  const int64_t out_step = a.n_particles * 7;
  const uint32_t tile_bytes = static_cast<uint32_t>(count) * 7u * sizeof(T);
  constexpr uint32_t kSparse = CH_FLAG_XY_UNCOUPLED | CH_FLAG_NO_TAU_COLUMN |
                               CH_FLAG_NO_Y_DISPERSION | CH_FLAG_DELTA_IDENTITY;
  constexpr uint32_t kCoupled = CH_FLAG_NO_TAU_COLUMN | CH_FLAG_DELTA_IDENTITY;

#pragma unroll 1
  for (int it = 0; it < n_local; ++it) {
    const int buf = it & 1;
    T* stage = buf ? stage1 : stage0;
    const T* rec = recs[buf];
    if (it + 1 < n_local) {  // in flight during the arithmetic below
      rec_src += a.record_stride;
      fetched = tid < RECLEN ? rec_src[tid] : T(0);
    }
    T sv[P];
#pragma unroll
    for (int k = 0; k < P; ++k) sv[k] = sv_in[k];
    const uint32_t flags = record_flags(rec[0]);
    if ((flags & kSparse) == kSparse)
      apply_setting_lean<T, NAP, 1, ELLIPTICAL>(rec, a.elliptical_mask, p, sv, stage, tid);
    else if ((flags & kCoupled) == kCoupled)
      apply_setting_lean<T, NAP, 2, ELLIPTICAL>(rec, a.elliptical_mask, p, sv, stage, tid);
    else
      apply_setting_lean<T, NAP, 0, ELLIPTICAL>(rec, a.elliptical_mask, p, sv, stage, tid);
    if (survival_out != nullptr) {
#pragma unroll
      for (int k = 0; k < P; ++k) {
        const int local = tid + k * THREADS;
        if (local < count) survival_out[local] = sv[k];
      }
      survival_out += a.n_particles;
    }
    if (it + 1 < n_local && tid < RECLEN) recs[buf ^ 1][tid] = fetched;
    if (a.bulk_out) {
      // the engine has finished READING the other tile (handed over one setting ago) before
      // anybody passes the barrier and writes to it again
      if (tid == 0) bulk_wait_read<0>();
      fence_async_shared();
      __syncthreads();
      if (tid == 0) {
        bulk_store(out, stage, tile_bytes);
        bulk_commit();
      }
    } else {
      __syncthreads();
      for (int i = tid; i < count * 7; i += THREADS) out[i] = stage[i];
      __syncthreads();
    }
    out += out_step;
  }
  if (a.bulk_out && tid == 0) bulk_wait<0>();
}

}  // namespace

// One beam (and one incoming survival vector) under consecutive records with at most three
// apertures and no cavity tail: the call shape the kernels of this file are specialised for.
template <typename T>
bool shared_beam_call(const ApplyArgs<T>& args, bool unit_seventh) {
  return unit_seventh && args.particle_stride == 0 && args.record_index == nullptr &&
         (args.survival_in == nullptr || args.survival_stride == 0) && args.n_apertures <= 3 &&
         args.record_len == CH_RECORD_LEN(args.n_apertures) && !args.has_cavity && !args.compact;
}

template bool shared_beam_call<float>(const ApplyArgs<float>&, bool);
template bool shared_beam_call<double>(const ApplyArgs<double>&, bool);

namespace {
template <typename T, typename Launch>
int with_apertures(const ApplyArgs<T>& args, Launch&& launch) {
  const bool elliptical = (args.elliptical_mask & ((1u << args.n_apertures) - 1u)) != 0;
  using std::integral_constant;
  auto pick = [&](auto nap) -> int {
    return elliptical ? launch(nap, std::true_type{}) : launch(nap, std::false_type{});
  };
  switch (args.n_apertures) {
    case 0: return launch(integral_constant<int, 0>{}, std::false_type{});
    case 1: return pick(integral_constant<int, 1>{});
    case 2: return pick(integral_constant<int, 2>{});
    default: return pick(integral_constant<int, 3>{});
  }
}
}  // namespace

int launch_observe_shared_beam(const ApplyArgs<float>& args, cudaStream_t stream) {
  constexpr int TP = kLeanP * kLeanThreads;
  const int64_t tiles = (args.n_particles + TP - 1) / TP;
  const int64_t chunks = (args.n_settings + args.settings_per_cta - 1) / args.settings_per_cta;
  CH_REQUIRE(tiles <= 2147483647LL && chunks <= 65535, "ch_apply_maps_moments: grid too large");
  const dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(chunks));
  const size_t tile_bytes = sizeof(float) * TP * 7;
  auto run = [&](auto moments) -> int {
    return with_apertures(args, [&](auto nap, auto elliptical) -> int {
      CH_CUDA(launch_dependent(
          observe_shared_beam_kernel<decltype(nap)::value, decltype(moments)::value,
                                     decltype(elliptical)::value>,
          grid, dim3(kLeanThreads), tile_bytes, stream, args));
      CH_LAUNCH_CHECK();
      return CH_OK;
    });
  };
  return args.covariance ? run(std::integral_constant<int, 2>{})
                         : run(std::integral_constant<int, 1>{});
}

template <typename T>
int launch_apply_shared_beam(const ApplyArgs<T>& args, cudaStream_t stream) {
  constexpr int TP = kApplyLeanP<T> * kApplyLeanThreads;
  const int64_t tiles = (args.n_particles + TP - 1) / TP;
  const int64_t chunks = (args.n_settings + args.settings_per_cta - 1) / args.settings_per_cta;
  CH_REQUIRE(tiles <= 2147483647LL && chunks <= 65535, "ch_apply_maps: grid too large");
  const dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(chunks));
  const size_t smem = sizeof(T) * 2 * TP * 7;
  return with_apertures(args, [&](auto nap, auto elliptical) -> int {
    auto kernel = apply_shared_beam_kernel<T, decltype(nap)::value, decltype(elliptical)::value>;
    CH_CUDA(allow_dynamic_smem(reinterpret_cast<const void*>(kernel), static_cast<int>(smem)));
    CH_CUDA(launch_dependent(kernel, grid, dim3(kApplyLeanThreads), smem, stream, args));
    CH_LAUNCH_CHECK();
    return CH_OK;
  });
}
template int launch_apply_shared_beam<float>(const ApplyArgs<float>&, cudaStream_t);
template int launch_apply_shared_beam<double>(const ApplyArgs<double>&, cudaStream_t);

}  // namespace ch

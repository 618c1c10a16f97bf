// Map application: the one streaming pass over the particles of a linear lattice section.
//
// Replaces every `particles @ tm.mT` (cheetah/accelerator/element.py:181-191) and every
// aperture mask update (cheetah/accelerator/aperture.py:108-132) between two non-linear
// elements with a single read of the incoming particles and a single write of the
// outgoing particles and survival probabilities.
//
// Work decomposition (B200): a CTA owns a tile of TP = THREADS*P particles and keeps their
// six phase-space coordinates in registers ("tile-stationary"); it then loops over a chunk
// of lattice settings.  For each setting the per-setting record (cumulative maps from
// ch_compose_maps) is read from a double-buffered shared-memory copy as broadcast
// 128-bit loads, every thread evaluates the aperture rows and the final 6x7 map for its P
// particles, writes the 7-wide rows into a shared staging tile (stride-7 stores are
// bank-conflict free) and one elected thread hands the contiguous tile to the TMA engine
// (cp.async.bulk shared->global, SASS UBLKCP) so the LSU never sees the 28-byte-strided
// row layout.  Two staging tiles overlap the bulk store of setting b with the math of b+1.
// HBM traffic per (particle, setting): 28 B + 4 B written, the shared beam is read once per
// CTA.  No tensor cores: K = 7 is far below any MMA tile.
//
// Kernels in this file: apply_maps_kernel (particles out, optionally with the fused moments /
// covariance epilogue and the cavity tail), observe_maps_kernel (float32 observables only: the
// same tiling on packed FFMA2 pairs, nothing written but 20 / 36 sums per setting) and
// apply_maps_parameter_kernel (ParameterBeam: mu' = M mu, cov' = M cov M^T).
#include "apply_common.cuh"

namespace ch {
namespace {


// One lattice setting for this thread's P particles: survival masks at every aperture, then
// the final map into the staging tile.  SPARSE drops the structurally-zero terms (flags).
// With MOMENTS the outgoing coordinates are also accumulated into `acc` (see the kernel) about
// `pilot`, the image of the beam's first particle under the same map.
// MODE: 0 dense (72 multiply-adds with three apertures), 1 sparse (29; all four CH_FLAG_* hold),
// 2 coupled (56: x-y coupling and dispersion allowed, but no tau column in rows 0-3 and delta
// untouched -- solenoids, tilted magnets, dipoles without RF)
template <typename T, int P, int THREADS, bool UNIT7, int MODE, int MOMENTS, bool WRITE,
          bool CAVITY, int ROW = 7>
__device__ __forceinline__ void process_setting(const T* rec, int n_apertures,
                                                uint32_t elliptical_mask, const T (&p)[P][7],
                                                T (&sv)[P], T* stage, int tid,
                                                const T (&first_particle)[7], T (&pilot)[6],
                                                Acc<T> (&acc)[MOMENTS == 2 ? 32 : 16],
                                                const T* cavity) {
  for (int ap = 0; ap < n_apertures; ++ap) {
    T q[16];
    load_coefficients(q, rec + CH_RECORD_HEADER + CH_RECORD_MAP + ap * CH_RECORD_APERTURE);
    const T x_max = q[14], y_max = q[15];
    const bool elliptical = (elliptical_mask >> ap) & 1u;
#pragma unroll
    for (int k = 0; k < P; ++k) {
      T x, y;
      if constexpr (MODE == 1) {
        const T w = UNIT7 ? T(1) : p[k][6];
        x = fma_t(q[0], p[k][0], fma_t(q[1], p[k][1], fma_t(q[5], p[k][5], UNIT7 ? q[6] : q[6] * w)));
        y = fma_t(q[9], p[k][2], fma_t(q[10], p[k][3], UNIT7 ? q[13] : q[13] * w));
      } else if constexpr (MODE == 2) {
        x = affine_row_no_tau<T, UNIT7>(q, p[k]);
        y = affine_row_no_tau<T, UNIT7>(q + 7, p[k]);
      } else {
        x = affine_row<T, UNIT7>(q, p[k]);
        y = affine_row<T, UNIT7>(q + 7, p[k]);
      }
      bool keep;
      if (elliptical) {
        const T ex = div_rn(mul_rn(x, x), mul_rn(x_max, x_max));
        const T ey = div_rn(mul_rn(y, y), mul_rn(y_max, y_max));
        keep = add_rn(ex, ey) <= T(1);
      } else {
        keep = inside(x, x_max) && inside(y, y_max);
      }
      // survival * mask with mask in {0, 1}: exact for every finite survival probability
      sv[k] = keep ? sv[k] : T(0);
    }
  }

  T c[44];
  load_coefficients(c, rec);
  const T* m = c + CH_RECORD_HEADER;
  // Active cavity at the end of the section (cavity.py:113-220): delta is replaced by the exact
  // update a delta_in + bV (cos(phi - tau_in b0 k) - cos phi) and tau gets the second-order
  // terms, both from the coordinates at the cavity ENTRANCE (rows 4, 5 of the block).  The
  // cosine difference is evaluated as -2 sin(phi + e/2) sin(e/2), which does not cancel.
  T cav[CAVITY ? CH_RECORD_CAVITY : 2];
  if constexpr (CAVITY) load_coefficients(cav, cavity);
  auto cavity_tail = [&](const T (&in)[7], T& r4, T& r5) {
    const T tau_in = affine_row<T, UNIT7>(cav, in);
    const T delta_in = affine_row<T, UNIT7>(cav + 7, in);
    const T half = T(-0.5) * tau_in * cav[16];
    T sh, ch;
    sincos_t(half, sh, ch);
    const T dcos = T(-2) * (cav[17] * ch + cav[18] * sh) * sh;
    r5 = fma_t(cav[14], delta_in, cav[15] * dcos);
    r4 += cav[19] * delta_in * delta_in + cav[20] * tau_in * delta_in + cav[21] * tau_in * tau_in;
  };
  // the six outgoing coordinates of one particle
  auto map_rows = [&](const T (&in)[7], T (&out)[6]) {
    if constexpr (MODE == 1) {
      const T w = UNIT7 ? T(1) : in[6];
      auto constant = [&](int i) { return UNIT7 ? m[i * 7 + 6] : m[i * 7 + 6] * w; };
      out[0] = fma_t(m[0], in[0], fma_t(m[1], in[1], fma_t(m[5], in[5], constant(0))));
      out[1] = fma_t(m[7], in[0], fma_t(m[8], in[1], fma_t(m[12], in[5], constant(1))));
      out[2] = fma_t(m[16], in[2], fma_t(m[17], in[3], constant(2)));
      out[3] = fma_t(m[23], in[2], fma_t(m[24], in[3], constant(3)));
      out[4] = fma_t(m[28], in[0],
                     fma_t(m[29], in[1], fma_t(m[32], in[4], fma_t(m[33], in[5], constant(4)))));
      out[5] = in[5];
    } else if constexpr (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 4; ++i) out[i] = affine_row_no_tau<T, UNIT7>(m + i * 7, in);
      out[4] = affine_row<T, UNIT7>(m + 28, in);
      out[5] = in[5];
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) out[i] = affine_row<T, UNIT7>(m + i * 7, in);
    }
  };
  if constexpr (!MOMENTS) {
    // plain path (the headline kernel): rows go straight to the staging tile
#pragma unroll
    for (int k = 0; k < P; ++k) {
      T* row = stage + (tid + k * THREADS) * ROW;
      if constexpr (MODE == 1) {
        const T w = UNIT7 ? T(1) : p[k][6];
        auto constant = [&](int i) { return UNIT7 ? m[i * 7 + 6] : m[i * 7 + 6] * w; };
        row[0] = fma_t(m[0], p[k][0], fma_t(m[1], p[k][1], fma_t(m[5], p[k][5], constant(0))));
        row[1] = fma_t(m[7], p[k][0], fma_t(m[8], p[k][1], fma_t(m[12], p[k][5], constant(1))));
        row[2] = fma_t(m[16], p[k][2], fma_t(m[17], p[k][3], constant(2)));
        row[3] = fma_t(m[23], p[k][2], fma_t(m[24], p[k][3], constant(3)));
        row[4] = fma_t(m[28], p[k][0],
                       fma_t(m[29], p[k][1],
                             fma_t(m[32], p[k][4], fma_t(m[33], p[k][5], constant(4)))));
        row[5] = p[k][5];
      } else if constexpr (MODE == 2) {
#pragma unroll
        for (int i = 0; i < 4; ++i) row[i] = affine_row_no_tau<T, UNIT7>(m + i * 7, p[k]);
        row[4] = affine_row<T, UNIT7>(m + 28, p[k]);
        row[5] = p[k][5];
      } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) row[i] = affine_row<T, UNIT7>(m + i * 7, p[k]);
      }
      if constexpr (ROW == 7) row[6] = UNIT7 ? T(1) : p[k][6];
      if constexpr (CAVITY) {
        T r4 = row[4], r5;
        cavity_tail(p[k], r4, r5);
        row[4] = r4;
        row[5] = r5;
      }
    }
    return;
  }
  if constexpr (MOMENTS) {
    map_rows(first_particle, pilot);
    if constexpr (CAVITY) cavity_tail(first_particle, pilot[4], pilot[5]);
  }
#pragma unroll
  for (int k = 0; k < P; ++k) {
    T out[6];
    map_rows(p[k], out);
    if constexpr (CAVITY) cavity_tail(p[k], out[4], out[5]);
    if constexpr (WRITE) {
      T* row = stage + (tid + k * THREADS) * 7;
#pragma unroll
      for (int i = 0; i < 6; ++i) row[i] = out[i];
      row[6] = UNIT7 ? T(1) : p[k][6];
    }
    if constexpr (MOMENTS) {
      // survival-weighted fp32 sums about the pilot (no cancellation); tail lanes carry w = 0
      const Acc<T> w = sv[k];
      acc[0] += w;
      acc[1] = fma_t(w, w, acc[1]);
      Acc<T> d[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        d[i] = out[i] - pilot[i];
        acc[2 + i] = fma_t(w, d[i], acc[2 + i]);
        acc[8 + i] = fma_t(w * d[i], d[i], acc[8 + i]);
      }
      if constexpr (MOMENTS == 2) {  // the 15 off-diagonal second moments, (0,1), (0,2), ... (4,5)
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const Acc<T> wd = w * d[i];
#pragma unroll
          for (int j = i + 1; j < 6; ++j) {
            constexpr int kBase = 14;
            const int slot = kBase + i * (11 - i) / 2 + (j - i - 1);
            acc[slot] = fma_t(wd, d[j], acc[slot]);
          }
        }
      }
    }
  }
}

template <typename T, int P, int THREADS, bool UNIT7, int MOMENTS, bool WRITE, bool CAVITY,
          bool COMPACT = false>
__global__ void __launch_bounds__(THREADS, sizeof(T) == 4 ? (MOMENTS == 2 ? 2 : 3) : 1)
apply_maps_kernel(const ApplyArgs<T> a) {
  constexpr int TP = P * THREADS;
  // COMPACT (host-bound output, ch_apply_maps_compact): the seventh column is known to be 1 and
  // is not written -- 24-byte rows -- and the survival mask can leave as one byte per particle
  constexpr int ROW = COMPACT ? 6 : 7;
  static_assert(!COMPACT || (UNIT7 && MOMENTS == 0 && WRITE), "compact output: plain path only");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* stage0 = reinterpret_cast<T*>(smem_raw);
  T* stage1 = stage0 + TP * 7;
  T* rec0 = stage1 + TP * 7;
  T* rec1 = rec0 + a.record_len;
  uint64_t* bar = reinterpret_cast<uint64_t*>(rec1 + a.record_len);

  const int tid = threadIdx.x;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * TP;
  const int count = static_cast<int>(min(static_cast<int64_t>(TP), a.n_particles - n0));
  const int64_t b_begin = static_cast<int64_t>(blockIdx.y) * a.settings_per_cta;
  const int64_t b_end = min(a.n_settings, b_begin + a.settings_per_cta);

  if (a.bulk_in && tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }

  auto record_offset = [&](int64_t b) {
    return (a.record_index ? a.record_index[b] : b) * a.record_stride;
  };
  auto copy_record = [&](T* dst, int64_t b) {
    const T* src = a.records + record_offset(b);
    for (int i = tid; i < a.record_len; i += THREADS) dst[i] = src[i];
  };

  // ---- fused observables (MOMENTS): survival-weighted sums of the OUTGOING coordinates
  // about a per-setting pilot (the image of particle 0), so that mu / sigma never need the
  // (B, N, 7) array in HBM (ParticleBeam.mu_* / sigma_*,
  // cheetah/particles/particle_beam.py:1699-1805, cheetah/utils/statistics.py:30-62)
  constexpr int NACC = MOMENTS == 2 ? 32 : 16;           // per-lane accumulators
  constexpr int NSUM = MOMENTS == 2 ? 29 : 14;           // of which are sums
  constexpr int NOUT = MOMENTS == 2 ? CH_MOMENTS_COV : CH_MOMENTS;
  __shared__ Acc<T> partial[2][THREADS / 32][NACC];
  __shared__ T pilot_shared[2][8];
  auto flush_moments = [&](int buf, int64_t b) {  // threads 0..34 after a barrier
    if (tid < NSUM) {
      double total = 0.0;
#pragma unroll
      for (int wi = 0; wi < THREADS / 32; ++wi) total += static_cast<double>(partial[buf][wi][tid]);
      // sums 0..13 keep their place, the pilot sits at 14..19, the off-diagonal sums follow
      atomicAdd(&a.moments_out[b * NOUT + (tid < 14 ? tid : tid + 6)], total);
    } else if (tid >= 32 && tid < 38 && blockIdx.x == 0) {
      a.moments_out[b * NOUT + 14 + (tid - 32)] = static_cast<double>(pilot_shared[buf][tid - 32]);
    }
  };

  copy_record(rec0, b_begin);
  __syncthreads();  // record 0 + mbarrier init visible

  T p[P][7];
  T first[7];  // the beam's particle 0 (pilot of the fused moments)
#pragma unroll
  for (int j = 0; j < 7; ++j) first[j] = T(0);
  T sv_in[P];
#pragma unroll
  for (int k = 0; k < P; ++k) sv_in[k] = T(1);
  int64_t loaded_particles = -1, loaded_survival = -1;
  uint32_t phase = 0;

  for (int64_t b = b_begin; b < b_end; ++b) {
    const int it = static_cast<int>(b - b_begin);
    T* stage = (it & 1) ? stage1 : stage0;
    const T* rec = (it & 1) ? rec1 : rec0;
    T* rec_next = (it & 1) ? rec0 : rec1;

    // the staging tile we are about to reuse was handed to the TMA engine two iterations
    // ago: wait until the engine has finished READING it (at most 1 newer group in flight)
    if (a.bulk_out && tid == 0) bulk_wait_read<1>();
    __syncthreads();
    if (MOMENTS && it > 0) flush_moments((it - 1) & 1, b - 1);

    // ---- incoming particle tile -> registers (once per CTA when the beam is shared) ----
    const int64_t p_off =
        (a.particle_index ? a.particle_index[b] : b) * a.particle_stride + n0 * 7;
    if (p_off != loaded_particles) {
      const T* src = a.particles_in + p_off;
      if (a.bulk_in) {
        if (tid == 0) {
          const uint32_t bytes = static_cast<uint32_t>(count) * 7u * sizeof(T);
          mbar_expect_tx(bar, bytes);
          bulk_load(stage, src, bytes, bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
      } else {
        for (int i = tid; i < count * 7; i += THREADS) stage[i] = src[i];
        __syncthreads();
      }
#pragma unroll
      for (int k = 0; k < P; ++k) {
        const int local = tid + k * THREADS;
        if (local < count) {
#pragma unroll
          for (int j = 0; j < 7; ++j) p[k][j] = stage[local * 7 + j];
        } else {
#pragma unroll
          for (int j = 0; j < 7; ++j) p[k][j] = T(0);
        }
      }
      if (MOMENTS) {
        const T* head = a.particles_in + (p_off - n0 * 7);
#pragma unroll
        for (int j = 0; j < 7; ++j) first[j] = head[j];
      }
      loaded_particles = p_off;
      __syncthreads();  // everyone has its registers before the tile is overwritten
    }
    if (a.survival_out != nullptr || MOMENTS || (COMPACT && a.survival_u8 != nullptr)) {
      const int64_t s_off =
          (a.survival_index ? a.survival_index[b] : b) * a.survival_stride + n0;
      if (s_off != loaded_survival) {
#pragma unroll
        for (int k = 0; k < P; ++k) {
          const int local = tid + k * THREADS;
          sv_in[k] = (a.survival_in != nullptr && local < count) ? a.survival_in[s_off + local]
                                                                 : T(1);
        }
        loaded_survival = s_off;
      }
    }

    // ---- prefetch the next setting's record (visible after the next barrier) ----------
    if (b + 1 < b_end) copy_record(rec_next, b + 1);

    // ---- apertures + final map -> staging tile -------------------------------------------
    // The record's sparsity flags are uniform over the CTA: when the lattice has no x-y
    // coupling, no tau dependence, no vertical dispersion and leaves delta untouched (every
    // uncoupled, cavity-off lattice such as ARES) 29 of the 72 multiply-adds remain.
    T sv[P];
#pragma unroll
    for (int k = 0; k < P; ++k) sv[k] = sv_in[k];
    const uint32_t flags = record_flags(rec[0]);
    constexpr uint32_t kSparse = CH_FLAG_XY_UNCOUPLED | CH_FLAG_NO_TAU_COLUMN |
                                 CH_FLAG_NO_Y_DISPERSION | CH_FLAG_DELTA_IDENTITY;
    T pilot[6];
    Acc<T> acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = Acc<T>(0);
    if (MOMENTS) {  // lanes past the end of the beam must not count
#pragma unroll
      for (int k = 0; k < P; ++k)
        if (tid + k * THREADS >= count) sv[k] = T(0);
    }
    const T* cavity =
        rec + CH_RECORD_HEADER + CH_RECORD_MAP + a.n_apertures * CH_RECORD_APERTURE;
    constexpr uint32_t kCoupled = CH_FLAG_NO_TAU_COLUMN | CH_FLAG_DELTA_IDENTITY;
    if ((flags & kSparse) == kSparse)
      process_setting<T, P, THREADS, UNIT7, 1, MOMENTS, WRITE, CAVITY, ROW>(
          rec, a.n_apertures, a.elliptical_mask, p, sv, stage, tid, first, pilot, acc, cavity);
    else if ((flags & kCoupled) == kCoupled && !CAVITY)
      process_setting<T, P, THREADS, UNIT7, 2, MOMENTS, WRITE, CAVITY, ROW>(
          rec, a.n_apertures, a.elliptical_mask, p, sv, stage, tid, first, pilot, acc, cavity);
    else
      process_setting<T, P, THREADS, UNIT7, 0, MOMENTS, WRITE, CAVITY, ROW>(
          rec, a.n_apertures, a.elliptical_mask, p, sv, stage, tid, first, pilot, acc, cavity);
    if constexpr (COMPACT) {
      if (a.survival_u8 != nullptr) {
        uint8_t* dst = a.survival_u8 + b * a.n_particles + n0;
#pragma unroll
        for (int k = 0; k < P; ++k) {
          const int local = tid + k * THREADS;
          if (local < count) dst[local] = sv[k] != T(0) ? 1 : 0;
        }
      }
    }
    if (a.survival_out != nullptr) {
      T* dst = a.survival_out + b * a.n_particles + n0;
#pragma unroll
      for (int k = 0; k < P; ++k) {
        const int local = tid + k * THREADS;
        if (local < count) dst[local] = sv[k];
      }
    }
    if (MOMENTS) {
      // packed warp reduction, then one fp64 atomic per (tile, setting, statistic) issued at
      // the top of the next iteration
      const int lane = tid & 31;
      const Acc<T> total = packed_warp_sum(acc, lane);
      if constexpr (MOMENTS == 2) {
        partial[it & 1][tid >> 5][lane] = total;
      } else if ((lane & 1) == 0) {
        const int index = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 +
                          ((lane >> 1) & 1);
        partial[it & 1][tid >> 5][index] = total;
      }
      if (tid < 6) pilot_shared[it & 1][tid] = pilot[tid];
    }

    // ---- hand the finished tile to the TMA engine (or copy it out cooperatively) -------
    if (!WRITE) continue;  // observables only: nothing leaves the SM
    T* out = a.particles_out + (b * a.n_particles + n0) * ROW;
    if (a.bulk_out) {
      fence_async_shared();
      __syncthreads();
      if (tid == 0) {
        bulk_store(out, stage, static_cast<uint32_t>(count) * ROW * sizeof(T));
        bulk_commit();
      }
    } else {
      __syncthreads();
      for (int i = tid; i < count * ROW; i += THREADS) out[i] = stage[i];
    }
  }
  if (MOMENTS && b_end > b_begin) {
    __syncthreads();
    flush_moments(static_cast<int>((b_end - b_begin - 1) & 1), b_end - 1);
  }
  if (a.bulk_out && tid == 0) bulk_wait<0>();
}

// ---- observables only, float32: packed-pair arithmetic ------------------------------------------
// With nothing to write, Segment.track_moments is bound by the FP32 pipe, not by HBM: per
// (particle, setting) ~66 multiply-adds (3 apertures, the map, 12 weighted sums) and the shared
// beam stays in registers.  On sm_100 a 3-register FFMA issues every second cycle per SM
// sub-partition; the packed FFMA2 / FADD2 / FMUL2 forms do two lanes' worth of work in the same
// slot (the scalar kernel measured 21.2 ms for ARES x 1e6 particles x 4096 settings, i.e. the
// scalar issue limit).  This kernel therefore keeps two particles per 64-bit register pair:
// P = 8 particles per thread as 4 pairs, every coefficient of the record duplicated in shared
// memory so that one broadcast LDS.128 delivers two ready-made (c, c) operands.  The chains are
// the same fma sequences as in process_setting, so aperture masks are bit-identical to the
// particle-writing kernel's.
template <int PAIRS, bool UNIT7, bool SPARSE, int MOMENTS>
__device__ __forceinline__ void observe_setting(const f2* rec2, int n_apertures,
                                                uint32_t elliptical_mask,
                                                const f2 (&p)[PAIRS][7], f2 (&sv)[PAIRS],
                                                const float (&first_particle)[7],
                                                float (&pilot)[6],
                                                f2 (&acc)[MOMENTS == 2 ? 29 : 14]) {
#pragma unroll 1
  for (int ap = 0; ap < n_apertures; ++ap) {
    f2 q[16];
    load_pairs(q, rec2 + CH_RECORD_HEADER + CH_RECORD_MAP + ap * CH_RECORD_APERTURE);
    const float x_max = q[14].x, y_max = q[15].x;
    f2 x[PAIRS], y[PAIRS];
#pragma unroll
    for (int k = 0; k < PAIRS; ++k) {
      if constexpr (SPARSE) {
        x[k] = fma2(q[0], p[k][0],
                    fma2(q[1], p[k][1], fma2(q[5], p[k][5], UNIT7 ? q[6] : mul2(q[6], p[k][6]))));
        y[k] = fma2(q[9], p[k][2], fma2(q[10], p[k][3], UNIT7 ? q[13] : mul2(q[13], p[k][6])));
      } else {
        x[k] = affine_row2<UNIT7>(q, p[k]);
        y[k] = affine_row2<UNIT7>(q + 7, p[k]);
      }
    }
    // one uniform branch per aperture; inside, predicates are combined without short-circuit
    // jumps (2 FSETP + 1 FSEL per particle)
    if ((elliptical_mask >> ap) & 1u) {
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
        const bool lo = inside(x[k].x, x_max) & inside(y[k].x, y_max);
        const bool hi = inside(x[k].y, x_max) & inside(y[k].y, y_max);
        sv[k].x = lo ? sv[k].x : 0.0f;
        sv[k].y = hi ? sv[k].y : 0.0f;
      }
    }
  }

  f2 c[44];
  load_pairs(c, rec2);
  const f2* m = c + CH_RECORD_HEADER;
  // pilot: image of the beam's particle 0 (scalar, same chains as process_setting), from the
  // low halves of the coefficient pairs
  {
    const float w = UNIT7 ? 1.0f : first_particle[6];
    const float(&in)[7] = first_particle;
    if constexpr (SPARSE) {
      auto constant = [&](int i) { return UNIT7 ? m[i * 7 + 6].x : m[i * 7 + 6].x * w; };
      pilot[0] = fmaf(m[0].x, in[0], fmaf(m[1].x, in[1], fmaf(m[5].x, in[5], constant(0))));
      pilot[1] = fmaf(m[7].x, in[0], fmaf(m[8].x, in[1], fmaf(m[12].x, in[5], constant(1))));
      pilot[2] = fmaf(m[16].x, in[2], fmaf(m[17].x, in[3], constant(2)));
      pilot[3] = fmaf(m[23].x, in[2], fmaf(m[24].x, in[3], constant(3)));
      pilot[4] = fmaf(m[28].x, in[0],
                      fmaf(m[29].x, in[1],
                           fmaf(m[32].x, in[4], fmaf(m[33].x, in[5], constant(4)))));
      pilot[5] = in[5];
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        float acc1 = UNIT7 ? m[i * 7 + 6].x : m[i * 7 + 6].x * w;
#pragma unroll
        for (int j = 5; j >= 0; --j) acc1 = fmaf(m[i * 7 + j].x, in[j], acc1);
        pilot[i] = acc1;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < PAIRS; ++k) {
    f2 out[6];
    if constexpr (SPARSE) {
      auto constant = [&](int i) { return UNIT7 ? m[i * 7 + 6] : mul2(m[i * 7 + 6], p[k][6]); };
      out[0] = fma2(m[0], p[k][0], fma2(m[1], p[k][1], fma2(m[5], p[k][5], constant(0))));
      out[1] = fma2(m[7], p[k][0], fma2(m[8], p[k][1], fma2(m[12], p[k][5], constant(1))));
      out[2] = fma2(m[16], p[k][2], fma2(m[17], p[k][3], constant(2)));
      out[3] = fma2(m[23], p[k][2], fma2(m[24], p[k][3], constant(3)));
      out[4] = fma2(m[28], p[k][0],
                    fma2(m[29], p[k][1],
                         fma2(m[32], p[k][4], fma2(m[33], p[k][5], constant(4)))));
      out[5] = p[k][5];
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) out[i] = affine_row2<UNIT7>(m + i * 7, p[k]);
    }
    const f2 w = sv[k];
    acc[0] = add2(acc[0], w);
    acc[1] = fma2(w, w, acc[1]);
    f2 d[6], wd[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      d[i] = add2(out[i], f2{-pilot[i], -pilot[i]});
      wd[i] = mul2(w, d[i]);
      acc[2 + i] = add2(acc[2 + i], wd[i]);
      acc[8 + i] = fma2(wd[i], d[i], acc[8 + i]);
    }
    if constexpr (MOMENTS == 2) {
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = i + 1; j < 6; ++j) {
          const int slot = 14 + i * (11 - i) / 2 + (j - i - 1);
          acc[slot] = fma2(wd[i], d[j], acc[slot]);
        }
    }
  }
}

template <int P, int THREADS, bool UNIT7, int MOMENTS>
__global__ void __launch_bounds__(THREADS, MOMENTS == 2 ? 2 : 4)
observe_maps_kernel(const ApplyArgs<float> a) {
  static_assert(P % 2 == 0, "particles are processed in pairs");
  constexpr int PAIRS = P / 2;
  constexpr int TP = P * THREADS;
  constexpr int NACC = MOMENTS == 2 ? 32 : 16;
  constexpr int NSUM = MOMENTS == 2 ? 29 : 14;
  constexpr int NOUT = MOMENTS == 2 ? CH_MOMENTS_COV : CH_MOMENTS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);
  const int rec_pitch = (a.record_len + 3) & ~3;  // keeps every buffer 16-byte aligned
  f2* pairs0 = reinterpret_cast<f2*>(tile + TP * 7);
  f2* pairs1 = pairs0 + rec_pitch;
  uint64_t* bar = reinterpret_cast<uint64_t*>(pairs1 + rec_pitch);
  __shared__ float partial[2][THREADS / 32][NACC];
  __shared__ float pilot_shared[2][8];

  const int tid = threadIdx.x;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * TP;
  const int count = static_cast<int>(min(static_cast<int64_t>(TP), a.n_particles - n0));
  const int64_t b_begin = static_cast<int64_t>(blockIdx.y) * a.settings_per_cta;
  const int64_t b_end = min(a.n_settings, b_begin + a.settings_per_cta);

  if (a.bulk_in && tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  // The next setting's record and beam / survival offsets are FETCHED into registers before the
  // arithmetic of the current setting and COMMITTED to shared memory after it, so that their L2
  // latency is covered by ~250 packed instructions instead of stalling every warp once per setting.
  constexpr int kFetch = 2;  // record entries per thread held in registers (the rest is copied)
  float fetched[kFetch];
  const bool beam_moves = a.particle_stride != 0;
  const bool survival_moves = a.survival_in != nullptr && a.survival_stride != 0;
  int64_t next_p_off = n0 * 7, next_s_off = n0;
  auto fetch = [&](int64_t b) {
    const float* src =
        a.records + (a.record_index ? a.record_index[b] : b) * a.record_stride;
#pragma unroll
    for (int r = 0; r < kFetch; ++r) {
      const int i = tid + r * THREADS;
      fetched[r] = i < a.record_len ? src[i] : 0.0f;
    }
    // a shared beam / shared (or absent) incoming survival has one offset for every setting
    if (beam_moves)
      next_p_off = (a.particle_index ? a.particle_index[b] : b) * a.particle_stride + n0 * 7;
    if (survival_moves)
      next_s_off = (a.survival_index ? a.survival_index[b] : b) * a.survival_stride + n0;
  };
  auto commit = [&](f2* pairs, int64_t b) {
#pragma unroll
    for (int r = 0; r < kFetch; ++r) {
      const int i = tid + r * THREADS;
      if (i < a.record_len) pairs[i] = f2{fetched[r], fetched[r]};
    }
    if (a.record_len > kFetch * THREADS) {  // more than 13 apertures
      const float* src =
          a.records + (a.record_index ? a.record_index[b] : b) * a.record_stride;
      for (int i = tid + kFetch * THREADS; i < a.record_len; i += THREADS) {
        const float v = src[i];
        pairs[i] = f2{v, v};
      }
    }
  };
  auto flush_moments = [&](int buf, int64_t b) {
    if (tid < NSUM) {
      double total = 0.0;
#pragma unroll
      for (int wi = 0; wi < THREADS / 32; ++wi) total += static_cast<double>(partial[buf][wi][tid]);
      atomicAdd(&a.moments_out[b * NOUT + (tid < 14 ? tid : tid + 6)], total);
    } else if (tid >= 32 && tid < 38 && blockIdx.x == 0) {
      a.moments_out[b * NOUT + 14 + (tid - 32)] = static_cast<double>(pilot_shared[buf][tid - 32]);
    }
  };

  fetch(b_begin);
  commit(pairs0, b_begin);
  __syncthreads();

  f2 p[PAIRS][7];
  float first[7];
#pragma unroll
  for (int j = 0; j < 7; ++j) first[j] = 0.0f;
  f2 sv_in[PAIRS];
#pragma unroll
  for (int k = 0; k < PAIRS; ++k) sv_in[k] = f2{1.0f, 1.0f};
  int64_t loaded_particles = -1, loaded_survival = -1;
  uint32_t phase = 0;

  for (int64_t b = b_begin; b < b_end; ++b) {
    const int it = static_cast<int>(b - b_begin);
    const f2* rec2 = (it & 1) ? pairs1 : pairs0;
    __syncthreads();  // this setting's record is complete; partial[(it - 1) & 1] is too
    if (it > 0) flush_moments((it - 1) & 1, b - 1);

    const int64_t p_off = next_p_off;
    if (p_off != loaded_particles) {
      const float* src = a.particles_in + p_off;
      if (a.bulk_in) {
        if (tid == 0) {
          const uint32_t bytes = static_cast<uint32_t>(count) * 7u * sizeof(float);
          mbar_expect_tx(bar, bytes);
          bulk_load(tile, src, bytes, bar);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
      } else {
        for (int i = tid; i < count * 7; i += THREADS) tile[i] = src[i];
        __syncthreads();
      }
#pragma unroll
      for (int k = 0; k < PAIRS; ++k) {
        const int lo = tid + (2 * k) * THREADS, hi = lo + THREADS;
#pragma unroll
        for (int j = 0; j < 7; ++j) {
          p[k][j].x = lo < count ? tile[lo * 7 + j] : 0.0f;
          p[k][j].y = hi < count ? tile[hi * 7 + j] : 0.0f;
        }
      }
      const float* head = a.particles_in + (p_off - n0 * 7);
#pragma unroll
      for (int j = 0; j < 7; ++j) first[j] = head[j];
      loaded_particles = p_off;
      __syncthreads();  // registers filled before the tile is loaded again
    }
    const int64_t s_off = next_s_off;
    if (s_off != loaded_survival) {
#pragma unroll
      for (int k = 0; k < PAIRS; ++k) {
        const int lo = tid + (2 * k) * THREADS, hi = lo + THREADS;
        // lanes past the end of the beam must not count
        sv_in[k].x = lo < count ? (a.survival_in ? a.survival_in[s_off + lo] : 1.0f) : 0.0f;
        sv_in[k].y = hi < count ? (a.survival_in ? a.survival_in[s_off + hi] : 1.0f) : 0.0f;
      }
      loaded_survival = s_off;
    }

    if (b + 1 < b_end) fetch(b + 1);

    f2 sv[PAIRS];
#pragma unroll
    for (int k = 0; k < PAIRS; ++k) sv[k] = sv_in[k];
    const uint32_t flags = record_flags(rec2[0].x);
    constexpr uint32_t kSparse = CH_FLAG_XY_UNCOUPLED | CH_FLAG_NO_TAU_COLUMN |
                                 CH_FLAG_NO_Y_DISPERSION | CH_FLAG_DELTA_IDENTITY;
    float pilot[6];
    f2 acc2[NSUM];
#pragma unroll
    for (int i = 0; i < NSUM; ++i) acc2[i] = f2{0.0f, 0.0f};
    if ((flags & kSparse) == kSparse)
      observe_setting<PAIRS, UNIT7, true, MOMENTS>(rec2, a.n_apertures, a.elliptical_mask, p,
                                                    sv, first, pilot, acc2);
    else
      observe_setting<PAIRS, UNIT7, false, MOMENTS>(rec2, a.n_apertures, a.elliptical_mask,
                                                     p, sv, first, pilot, acc2);
    if (b + 1 < b_end) commit((it & 1) ? pairs0 : pairs1, b + 1);
    if (a.survival_out != nullptr) {
      float* dst = a.survival_out + b * a.n_particles + n0;
#pragma unroll
      for (int k = 0; k < PAIRS; ++k) {
        const int lo = tid + (2 * k) * THREADS, hi = lo + THREADS;
        if (lo < count) dst[lo] = sv[k].x;
        if (hi < count) dst[hi] = sv[k].y;
      }
    }
    // fold the two halves, then the packed butterfly over the warp (a shared-memory transpose
    // instead of the shuffle tree executed fewer instructions but was not faster: the kernel
    // is bound by dependent-issue latency at 4 warps per scheduler, not by instruction count)
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = 0.0f;
#pragma unroll
    for (int i = 0; i < NSUM; ++i) acc[i] = acc2[i].x + acc2[i].y;
    const int lane = tid & 31;
    const float total = packed_warp_sum(acc, lane);
    if constexpr (MOMENTS == 2) {
      partial[it & 1][tid >> 5][lane] = total;
    } else if ((lane & 1) == 0) {
      const int index = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 +
                        ((lane >> 1) & 1);
      partial[it & 1][tid >> 5][index] = total;
    }
    if (tid == 0) {
#pragma unroll
      for (int i = 0; i < 6; ++i) pilot_shared[it & 1][i] = pilot[i];
    }
  }
  if (b_end > b_begin) {
    __syncthreads();
    flush_moments(static_cast<int>((b_end - b_begin - 1) & 1), b_end - 1);
  }
}

// (covariance with 2 pairs x 256 threads -- 124 registers, 16 warps/SM instead of 8 -- measured
// 23.8 ms against 21.3 ms: the 31-shuffle reduction per warp and setting doubles per particle)
constexpr int kObserveP = 8, kObserveThreads = 128;

int launch_observe(const ApplyArgs<float>& args, bool unit_seventh, cudaStream_t stream) {
  constexpr int TP = kObserveP * kObserveThreads;
  const size_t rec_pitch = (static_cast<size_t>(args.record_len) + 3) & ~size_t(3);
  const size_t smem = sizeof(float) * (TP * 7 + 4 * rec_pitch) + sizeof(uint64_t);
  const int64_t tiles = (args.n_particles + TP - 1) / TP;
  const int64_t chunks = (args.n_settings + args.settings_per_cta - 1) / args.settings_per_cta;
  CH_REQUIRE(tiles <= 2147483647LL && chunks <= 65535, "ch_apply_maps_moments: grid too large");
  dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(chunks));
  if (shared_beam_call(args, unit_seventh) && args.survival_out == nullptr)
    return launch_observe_shared_beam(args, stream);
  auto launch = [&](auto kernel) -> int {
    CH_CUDA(allow_dynamic_smem(reinterpret_cast<const void*>(kernel), static_cast<int>(smem)));
    kernel<<<grid, kObserveThreads, smem, stream>>>(args);
    return CH_OK;
  };
  int status;
  if (args.covariance)
    status = unit_seventh ? launch(observe_maps_kernel<kObserveP, kObserveThreads, true, 2>)
                          : launch(observe_maps_kernel<kObserveP, kObserveThreads, false, 2>);
  else
    status = unit_seventh ? launch(observe_maps_kernel<kObserveP, kObserveThreads, true, 1>)
                          : launch(observe_maps_kernel<kObserveP, kObserveThreads, false, 1>);
  if (status != CH_OK) return status;
  CH_LAUNCH_CHECK();
  return CH_OK;
}

template <typename T, int P, int THREADS>
int launch_apply(const ApplyArgs<T>& args, bool unit_seventh, cudaStream_t stream) {
  constexpr int TP = P * THREADS;
  const size_t smem =
      sizeof(T) * (2 * TP * 7 + 2 * static_cast<size_t>(args.record_len)) + sizeof(uint64_t);
  const int64_t tiles = (args.n_particles + TP - 1) / TP;
  const int64_t chunks = (args.n_settings + args.settings_per_cta - 1) / args.settings_per_cta;
  CH_REQUIRE(tiles <= 2147483647LL && chunks <= 65535, "ch_apply_maps: grid too large");
  dim3 grid(static_cast<unsigned>(tiles), static_cast<unsigned>(chunks));
  auto launch = [&](auto kernel) -> int {
    CH_CUDA(allow_dynamic_smem(reinterpret_cast<const void*>(kernel), static_cast<int>(smem)));
    kernel<<<grid, THREADS, smem, stream>>>(args);
    return CH_OK;
  };
  // compile-time variants: (unit 7th column) x (moments epilogue) x (write particles) x (cavity)
  auto pick = [&](auto unit, auto moments, auto write, auto cavity) -> int {
    return launch(apply_maps_kernel<T, P, THREADS, decltype(unit)::value, decltype(moments)::value,
                                    decltype(write)::value, decltype(cavity)::value>);
  };
  auto with_cavity = [&](auto unit, auto moments, auto write) -> int {
    return args.has_cavity ? pick(unit, moments, write, std::true_type{})
                           : pick(unit, moments, write, std::false_type{});
  };
  using M0 = std::integral_constant<int, 0>;
  using M1 = std::integral_constant<int, 1>;
  using M2 = std::integral_constant<int, 2>;
  auto with_outputs = [&](auto unit) -> int {
    if (args.moments_out != nullptr && args.covariance) {
      if (args.particles_out == nullptr) return with_cavity(unit, M2{}, std::false_type{});
      return with_cavity(unit, M2{}, std::true_type{});
    }
    if (args.moments_out != nullptr && args.particles_out == nullptr)
      return with_cavity(unit, M1{}, std::false_type{});
    if (args.moments_out != nullptr) return with_cavity(unit, M1{}, std::true_type{});
    return with_cavity(unit, M0{}, std::true_type{});
  };
  int status;
  // particles out, nothing else: the kernels specialised for one beam under many settings
  if (args.moments_out == nullptr && args.particles_out != nullptr &&
      shared_beam_call(args, unit_seventh))
    return launch_apply_shared_beam(args, stream);
  if (args.compact)
    status = launch(apply_maps_kernel<T, P, THREADS, true, 0, true, false, true>);
  else
    status = unit_seventh ? with_outputs(std::true_type{}) : with_outputs(std::false_type{});
  if (status != CH_OK) return status;
  CH_LAUNCH_CHECK();
  return CH_OK;
}

template <typename T>
int apply_typed(const void* particles_in, int64_t particle_stride, const int32_t* particle_index,
                const void* survival_in, int64_t survival_stride, const int32_t* survival_index,
                const void* records, int64_t record_stride, const int32_t* record_index,
                int64_t record_len, int32_t n_apertures, uint32_t elliptical_mask,
                int64_t n_particles, int64_t n_settings, void* particles_out, void* survival_out,
                int32_t unit_seventh, double* moments_out, int32_t covariance,
                cudaStream_t stream, uint8_t* survival_u8 = nullptr, bool compact = false) {
  ApplyArgs<T> a;
  a.compact = compact ? 1 : 0;
  a.survival_u8 = survival_u8;
  a.moments_out = moments_out;
  a.covariance = covariance;
  a.particles_in = static_cast<const T*>(particles_in);
  a.survival_in = static_cast<const T*>(survival_in);
  a.records = static_cast<const T*>(records);
  a.particles_out = static_cast<T*>(particles_out);
  a.survival_out = static_cast<T*>(survival_out);
  a.particle_index = particle_index;
  a.survival_index = survival_index;
  a.record_index = record_index;
  a.particle_stride = particle_stride;
  a.survival_stride = survival_stride;
  a.record_stride = record_stride;
  a.n_particles = n_particles;
  a.n_settings = n_settings;
  a.record_len = static_cast<int32_t>(record_len);
  a.has_cavity = record_len == CH_RECORD_LEN(n_apertures) + CH_RECORD_CAVITY ? 1 : 0;
  a.n_apertures = n_apertures;
  a.elliptical_mask = elliptical_mask;

  // cp.async.bulk needs 16-byte aligned addresses and sizes for every tile
  auto tiles_aligned = [&](const void* base, int64_t batch_stride_elems, int row = 7) {
    return reinterpret_cast<uintptr_t>(base) % 16 == 0 &&
           (static_cast<size_t>(n_particles) * row * sizeof(T)) % 16 == 0 &&
           (static_cast<size_t>(batch_stride_elems) * sizeof(T)) % 16 == 0;
  };
  const int row_out = compact ? 6 : 7;
  a.bulk_in = tiles_aligned(particles_in, particle_stride) ? 1 : 0;
  a.bulk_out = (particles_out != nullptr &&
                tiles_aligned(particles_out, n_particles * row_out, row_out)) ? 1 : 0;

  // settings per CTA: amortise the tile load over many settings but keep >= ~8 waves of CTAs
  constexpr int P = sizeof(T) == 4 ? 4 : 2;
  constexpr int THREADS = 256;
  const int64_t tiles = (n_particles + P * THREADS - 1) / (P * THREADS);
  int64_t per_cta = 64;
  while (per_cta > 1 && tiles * ((n_settings + per_cta - 1) / per_cta) < 148 * 16) per_cta /= 2;
  a.settings_per_cta = static_cast<int32_t>(per_cta);
  if constexpr (sizeof(T) == 4) {
    static_assert(kObserveP * kObserveThreads == P * THREADS, "same tiling for both kernels");
    // observables only (no cavity tail): the packed-pair kernel
    if (moments_out != nullptr && particles_out == nullptr && !a.has_cavity)
      return launch_observe(a, unit_seventh != 0, stream);
  }
  return launch_apply<T, P, THREADS>(a, unit_seventh != 0, stream);
}

}  // namespace
}  // namespace ch

namespace {
int apply_dispatch(const void* particles_in, int64_t particle_stride,
                   const int32_t* particle_index, const void* survival_in,
                   int64_t survival_stride, const int32_t* survival_index, const void* records,
                   int64_t record_stride, const int32_t* record_index, int64_t record_len,
                   int32_t n_apertures, uint32_t elliptical_mask, int64_t n_particles,
                   int64_t n_settings, void* particles_out, void* survival_out, int32_t dtype,
                   int32_t unit_seventh, double* moments_out, int32_t covariance, void* stream) {
  CH_REQUIRE(particles_in && records, "ch_apply_maps: NULL pointer argument");
  CH_REQUIRE(particles_out || moments_out, "ch_apply_maps: no output requested");
  CH_REQUIRE(n_particles > 0 && n_settings > 0, "ch_apply_maps: empty beam or batch");
  CH_REQUIRE(n_apertures >= 0 && n_apertures <= CH_MAX_APERTURES,
             "ch_apply_maps: n_apertures %d outside [0, %d]", n_apertures, CH_MAX_APERTURES);
  CH_REQUIRE(record_len == CH_RECORD_LEN(n_apertures) ||
                 record_len == CH_RECORD_LEN(n_apertures) + CH_RECORD_CAVITY,
             "ch_apply_maps: record_len %lld does not match %d apertures",
             static_cast<long long>(record_len), n_apertures);
  CH_REQUIRE(n_apertures == 0 || survival_out != nullptr || particles_out == nullptr,
             "ch_apply_maps: survival_out is required when apertures are present");
  CH_REQUIRE(dtype == CH_F32 || dtype == CH_F64, "ch_apply_maps: bad dtype %d", dtype);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (moments_out != nullptr)
    CH_CUDA(cudaMemsetAsync(moments_out, 0,
                            sizeof(double) * (covariance ? CH_MOMENTS_COV : CH_MOMENTS) * n_settings,
                            s));
  if (dtype == CH_F32)
    return ch::apply_typed<float>(particles_in, particle_stride, particle_index, survival_in,
                                  survival_stride, survival_index, records, record_stride,
                                  record_index, record_len, n_apertures, elliptical_mask,
                                  n_particles, n_settings, particles_out, survival_out,
                                  unit_seventh, moments_out, covariance, s);
  return ch::apply_typed<double>(particles_in, particle_stride, particle_index, survival_in,
                                 survival_stride, survival_index, records, record_stride,
                                 record_index, record_len, n_apertures, elliptical_mask,
                                 n_particles, n_settings, particles_out, survival_out,
                                 unit_seventh, moments_out, covariance, s);
}
}  // namespace

extern "C" int ch_apply_maps(const void* particles_in, int64_t particle_stride,
                             const int32_t* particle_index, const void* survival_in,
                             int64_t survival_stride, const int32_t* survival_index,
                             const void* records, int64_t record_stride,
                             const int32_t* record_index, int64_t record_len,
                             int32_t n_apertures, uint32_t elliptical_mask, int64_t n_particles,
                             int64_t n_settings, void* particles_out, void* survival_out,
                             int32_t dtype, int32_t unit_seventh, void* stream) {
  CH_REQUIRE(particles_out != nullptr, "ch_apply_maps: particles_out is NULL");
  return apply_dispatch(particles_in, particle_stride, particle_index, survival_in,
                        survival_stride, survival_index, records, record_stride, record_index,
                        record_len, n_apertures, elliptical_mask, n_particles, n_settings,
                        particles_out, survival_out, dtype, unit_seventh, nullptr, 0, stream);
}

extern "C" int ch_apply_maps_compact(const void* particles_in, int64_t particle_stride,
                                     const int32_t* particle_index, const void* survival_in,
                                     int64_t survival_stride, const int32_t* survival_index,
                                     const void* records, int64_t record_stride,
                                     const int32_t* record_index, int64_t record_len,
                                     int32_t n_apertures, uint32_t elliptical_mask,
                                     int64_t n_particles, int64_t n_settings, void* coordinates_out,
                                     void* survival_out, uint8_t* survival_mask_out, int32_t dtype,
                                     void* stream) {
  CH_REQUIRE(particles_in && records && coordinates_out,
             "ch_apply_maps_compact: NULL pointer argument");
  CH_REQUIRE(n_particles > 0 && n_settings > 0, "ch_apply_maps_compact: empty beam or batch");
  CH_REQUIRE(n_apertures >= 0 && n_apertures <= CH_MAX_APERTURES,
             "ch_apply_maps_compact: n_apertures %d outside [0, %d]", n_apertures,
             CH_MAX_APERTURES);
  CH_REQUIRE(record_len == CH_RECORD_LEN(n_apertures),
             "ch_apply_maps_compact: record_len %lld does not match %d apertures (sections ending "
             "in an active cavity are not supported)",
             static_cast<long long>(record_len), n_apertures);
  CH_REQUIRE(n_apertures == 0 || survival_out != nullptr || survival_mask_out != nullptr,
             "ch_apply_maps_compact: a survival output is required when apertures are present");
  CH_REQUIRE(dtype == CH_F32 || dtype == CH_F64, "ch_apply_maps_compact: bad dtype %d", dtype);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == CH_F32)
    return ch::apply_typed<float>(particles_in, particle_stride, particle_index, survival_in,
                                  survival_stride, survival_index, records, record_stride,
                                  record_index, record_len, n_apertures, elliptical_mask,
                                  n_particles, n_settings, coordinates_out, survival_out, 1,
                                  nullptr, 0, s, survival_mask_out, true);
  return ch::apply_typed<double>(particles_in, particle_stride, particle_index, survival_in,
                                 survival_stride, survival_index, records, record_stride,
                                 record_index, record_len, n_apertures, elliptical_mask,
                                 n_particles, n_settings, coordinates_out, survival_out, 1,
                                 nullptr, 0, s, survival_mask_out, true);
}

extern "C" int ch_apply_maps_moments(const void* particles_in, int64_t particle_stride,
                                     const int32_t* particle_index, const void* survival_in,
                                     int64_t survival_stride, const int32_t* survival_index,
                                     const void* records, int64_t record_stride,
                                     const int32_t* record_index, int64_t record_len,
                                     int32_t n_apertures, uint32_t elliptical_mask,
                                     int64_t n_particles, int64_t n_settings, void* particles_out,
                                     void* survival_out, double* moments_out, int32_t dtype,
                                     int32_t unit_seventh, void* stream) {
  CH_REQUIRE(moments_out != nullptr, "ch_apply_maps_moments: moments_out is NULL");
  return apply_dispatch(particles_in, particle_stride, particle_index, survival_in,
                        survival_stride, survival_index, records, record_stride, record_index,
                        record_len, n_apertures, elliptical_mask, n_particles, n_settings,
                        particles_out, survival_out, dtype, unit_seventh, moments_out, 0, stream);
}

extern "C" int ch_apply_maps_covariance(const void* particles_in, int64_t particle_stride,
                                        const int32_t* particle_index, const void* survival_in,
                                        int64_t survival_stride, const int32_t* survival_index,
                                        const void* records, int64_t record_stride,
                                        const int32_t* record_index, int64_t record_len,
                                        int32_t n_apertures, uint32_t elliptical_mask,
                                        int64_t n_particles, int64_t n_settings,
                                        void* particles_out, void* survival_out,
                                        double* moments_out, int32_t dtype, int32_t unit_seventh,
                                        void* stream) {
  CH_REQUIRE(moments_out != nullptr, "ch_apply_maps_covariance: moments_out is NULL");
  return apply_dispatch(particles_in, particle_stride, particle_index, survival_in,
                        survival_stride, survival_index, records, record_stride, record_index,
                        record_len, n_apertures, elliptical_mask, n_particles, n_settings,
                        particles_out, survival_out, dtype, unit_seventh, moments_out, 1, stream);
}

// ---- ParameterBeam: mu' = M mu, cov' = M cov M^T (cheetah/accelerator/element.py:166-179) --------
namespace ch {
namespace {

template <typename T>
__global__ void __launch_bounds__(64)
apply_maps_parameter_kernel(const T* __restrict__ mu_in, int64_t mu_stride,
                            const int32_t* __restrict__ mu_index, const T* __restrict__ cov_in,
                            int64_t cov_stride, const T* __restrict__ records,
                            int64_t record_stride, const int32_t* __restrict__ record_index,
                            int32_t cavity_offset, T* __restrict__ mu_out,
                            T* __restrict__ cov_out) {
  __shared__ double m[7][7], mu[7], cov[7][7], half[7][7];
  __shared__ double entrance[2][7], entrance_cov[2][7];
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  const int64_t beam = mu_index ? mu_index[b] : b;
  const T* rec = records + (record_index ? record_index[b] : b) * record_stride + CH_RECORD_HEADER;
  if (tid < 49) {
    const int i = tid / 7, j = tid - i * 7;
    m[i][j] = i < 6 ? static_cast<double>(rec[i * 7 + j]) : (j == 6 ? 1.0 : 0.0);
    cov[i][j] = static_cast<double>(cov_in[beam * cov_stride + tid]);
  }
  if (tid < 7) mu[tid] = static_cast<double>(mu_in[beam * mu_stride + tid]);
  __syncthreads();
  if (tid < 49) {  // half = M cov
    const int i = tid / 7, j = tid - i * 7;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) acc = fma(m[i][k], cov[k][j], acc);
    half[i][j] = acc;
  }
  __syncthreads();
  if (tid < 49) {  // cov' = half M^T
    const int i = tid / 7, j = tid - i * 7;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) acc = fma(half[i][k], m[j][k], acc);
    cov_out[b * 49 + tid] = static_cast<T>(acc);
  } else if (tid < 56) {
    const int i = tid - 49;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) acc = fma(m[i][k], mu[k], acc);
    mu_out[b * 7 + i] = static_cast<T>(acc);
  }
  if (cavity_offset < 0) return;
  // Active cavity at the end of the section, ParameterBeam branch of Cavity.track
  // (cavity.py:129-135, :203-217): the longitudinal entries are replaced by expressions in the
  // moments at the cavity ENTRANCE, i.e. under the rows (tau, delta) of the map up to there.
  const T* cav = rec - CH_RECORD_HEADER + cavity_offset;
  __syncthreads();  // the plain results above are written
  if (tid < 14) entrance[tid / 7][tid % 7] = static_cast<double>(cav[tid]);
  __syncthreads();
  if (tid < 14) {  // entrance_cov[r] = row_r . cov
    const int r = tid / 7, j = tid % 7;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 7; ++k) acc = fma(entrance[r][k], cov[k][j], acc);
    entrance_cov[r][j] = acc;
  }
  __syncthreads();
  if (tid == 0) {
    double tau = 0.0, delta = 0.0, c44 = 0.0, c45 = 0.0, c55 = 0.0;
    for (int k = 0; k < 7; ++k) {
      tau = fma(entrance[0][k], mu[k], tau);
      delta = fma(entrance[1][k], mu[k], delta);
      c44 = fma(entrance_cov[0][k], entrance[0][k], c44);
      c45 = fma(entrance_cov[0][k], entrance[1][k], c45);
      c55 = fma(entrance_cov[1][k], entrance[1][k], c55);
    }
    const double a = cav[14], bv = cav[15], b0k = cav[16], sphi = cav[17], cphi = cav[18];
    const double t566 = cav[19], t556 = cav[20], t555 = cav[21];
    double se, ce;
    sincos(tau * b0k, &se, &ce);
    const double dcos = cphi * ce + sphi * se - cphi;  // cos(phi - tau b0 k) - cos(phi)
    mu_out[b * 7 + 5] = static_cast<T>(a * delta + bv * dcos);
    double m4 = 0.0;
    for (int k = 0; k < 7; ++k) m4 = fma(m[4][k], mu[k], m4);
    mu_out[b * 7 + 4] = static_cast<T>(m4 + t566 * delta * delta + t556 * tau * delta +
                                       t555 * tau * tau);
    const double longitudinal = t566 * c55 * c55 + t556 * c45 * c55 + t555 * c44 * c44;
    cov_out[b * 49 + 4 * 7 + 4] = static_cast<T>(longitudinal);
    cov_out[b * 49 + 4 * 7 + 5] = static_cast<T>(longitudinal);
    cov_out[b * 49 + 5 * 7 + 4] = static_cast<T>(longitudinal);
    cov_out[b * 49 + 5 * 7 + 5] = static_cast<T>(c55);
  }
}

}  // namespace
}  // namespace ch

extern "C" int ch_apply_maps_parameter(const void* mu_in, int64_t mu_stride,
                                       const int32_t* mu_index, const void* cov_in,
                                       int64_t cov_stride, const void* records,
                                       int64_t record_stride, const int32_t* record_index,
                                       int32_t cavity_offset, int64_t n_settings, void* mu_out,
                                       void* cov_out, int32_t dtype, void* stream) {
  CH_REQUIRE(mu_in && cov_in && records && mu_out && cov_out,
             "ch_apply_maps_parameter: NULL pointer argument");
  CH_REQUIRE(n_settings > 0 && n_settings <= 2147483647LL,
             "ch_apply_maps_parameter: n_settings must be in [1, 2^31)");
  CH_REQUIRE(dtype == CH_F32 || dtype == CH_F64, "ch_apply_maps_parameter: bad dtype %d", dtype);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned blocks = static_cast<unsigned>(n_settings);
  if (dtype == CH_F32)
    ch::apply_maps_parameter_kernel<float><<<blocks, 64, 0, s>>>(
        static_cast<const float*>(mu_in), mu_stride, mu_index, static_cast<const float*>(cov_in),
        cov_stride, static_cast<const float*>(records), record_stride, record_index,
        cavity_offset, static_cast<float*>(mu_out), static_cast<float*>(cov_out));
  else
    ch::apply_maps_parameter_kernel<double><<<blocks, 64, 0, s>>>(
        static_cast<const double*>(mu_in), mu_stride, mu_index,
        static_cast<const double*>(cov_in), cov_stride, static_cast<const double*>(records),
        record_stride, record_index, cavity_offset, static_cast<double*>(mu_out),
        static_cast<double*>(cov_out));
  CH_LAUNCH_CHECK();
  return CH_OK;
}

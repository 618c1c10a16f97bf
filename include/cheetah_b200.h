/*
 * cheetah_b200 -- C ABI of the B200-native backend for Cheetah's
 * Segment.track(ParticleBeam) hot path.
 *
 * The reference (desy-ml/cheetah @ 60d1053) is pure Python/PyTorch and has no FFI; the
 * drop-in boundary is its public Python API.  Each entry point below
 * replaces the body of one group of reference functions (cited as file:line relative to
 * the reference root) and is what a ctypes binding inside those functions would call
 * (see INTEGRATION.md).
 *
 * Conventions
 *  - every function returns 0 on success and a negative CH_E* code on failure; the
 *    message for the calling thread is available from ch_last_error().  Nothing throws
 *    across the ABI.
 *  - all data pointers are DEVICE pointers unless the name ends in _host; the caller
 *    owns every buffer it passes in, including outputs.  The library only owns
 *    ch_program objects; scratch buffers are passed in by the caller.
 *  - `stream` is a cudaStream_t passed as void* (PyTorch:
 *    torch.cuda.current_stream().cuda_stream); all work is enqueued on it, no call
 *    synchronises the device.
 *  - dtype arguments are CH_F32 or CH_F64.
 *  - "settings" (B) is the flattened vectorised batch of lattice settings and/or beams.
 */
#ifndef CHEETAH_B200_H
#define CHEETAH_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* 2: non-linear tracking, diagnostics, fused gather, quad-block charge grid (ch_sc_deposit) */
#define CH_ABI_VERSION 3

enum { CH_F32 = 0, CH_F64 = 1 };

enum {
  CH_OK = 0,
  CH_EINVAL = -1,   /* bad argument */
  CH_ECUDA = -2,    /* CUDA runtime error (message carries cudaGetErrorString) */
  CH_ENOMEM = -3,
  CH_EUNSUPPORTED = -4
};

/* Element opcodes of a lowered lattice program (one op per flattened element).        */
enum {
  CH_OP_IDENTITY = 0,     /* Marker, inactive BPM/Screen/Aperture: marker.py:44-50,      */
                          /*   bpm.py:69-75, screen.py:176-185, aperture.py:82-88        */
  CH_OP_DRIFT = 1,        /* slots: length            track_methods.py:284-299           */
                          /*   (also Sextupole with tracking_method="linear",            */
                          /*   sextupole.py:84-88)                                       */
  CH_OP_CORRECTOR = 2,    /* slots: length, angle_h, angle_v (either may be -1)          */
                          /*   horizontal_corrector.py:60-78, vertical_corrector.py:60-78*/
                          /*   combined_corrector.py:76-98                               */
  CH_OP_QUADRUPOLE = 3,   /* slots: length, k1, tilt, mis_x, mis_y  quadrupole.py:93-110 */
  CH_OP_DIPOLE = 4,       /* slots: length, angle, k1, e1, e2, fint, fint_exit, gap,     */
                          /*   tilt   dipole.py:372-394, :430-466 (RBend: rbend.py:83-101*/
                          /*   lowers to the same op with e1/e2 already offset)          */
  CH_OP_SOLENOID = 5,     /* slots: length, k, mis_x, mis_y        solenoid.py:74-116    */
  CH_OP_UNDULATOR = 6,    /* slots: length, period, kx, ky         undulator.py:78-125   */
  CH_OP_CAVITY_OFF = 7,   /* slots: length; op_flags bit0 = traveling wave               */
                          /*   cavity.py:253-358 evaluated at voltage == 0               */
  CH_OP_CUSTOM_MAP = 8,   /* slots: map (7x7 row-major, slot stride = 49 per setting),   */
                          /*   optional length                                           */
                          /*   custom_transfer_map.py:111-114                            */
  CH_OP_APERTURE = 9,     /* slots: x_max, y_max; op_flags bit0 = elliptical             */
                          /*   aperture.py:90-132 (cut point, not a map)                 */
  CH_OP_CAVITY = 10,      /* ACTIVE cavity (voltage != 0), must be the LAST op of its     */
                          /*   section.  slots: length, voltage, phase [deg], frequency,  */
                          /*   gain flag (device scalar, non-zero = use the energy-gain   */
                          /*   second-order terms: the reference's batch-wide             */
                          /*   `(delta_energy > 0).any()`, cavity.py:155);                */
                          /*   op_flags bit0 = traveling wave                             */
                          /*   cavity.py:100-251 (track), :253-358 (R matrix)             */
  /* per-particle NON-LINEAR ops (ch_track_nonlinear; never part of a linear section)      */
  CH_OP_DKD_DRIFT = 11,   /* Drift, tracking_method="drift_kick_drift"; slots: length      */
                          /*   drift.py:106-154, bmadx.py:271-302                          */
  CH_OP_DKD_QUADRUPOLE = 12, /* slots: length, k1, tilt, mis_x, mis_y; op_flags = num_steps */
                          /*   quadrupole.py:168-251, bmadx.py:115-260                     */
  CH_OP_DKD_DIPOLE = 13,  /* slots: length, angle, e1, e2, fint, fint_exit, gap, gap_exit, */
                          /*   tilt; op_flags bit0 = fringe at entrance, bit1 = at exit    */
                          /*   dipole.py:183-370                                           */
  CH_OP_DKD_TDC = 14,     /* TransverseDeflectingCavity; slots: length, voltage, phase     */
                          /*   [rad / 2 pi], frequency, tilt, mis_x, mis_y                 */
                          /*   transverse_deflecting_cavity.py:122-209                     */
  CH_OP_SECOND_ORDER = 15 /* tracking_method="second_order" of Drift / Quadrupole /        */
                          /*   Sextupole / Dipole / RBend; slots: length, k1, k2, angle,   */
                          /*   e1, e2, fint, fint_exit, gap, tilt, mis_x, mis_y (unused    */
                          /*   ones point at a zero); op_flags bit0 = bend (pole faces +   */
                          /*   rotation; otherwise rotation + misalignment)                */
                          /*   track_methods.py:80-281, element.py:195-225, drift.py:67-83,*/
                          /*   quadrupole.py:112-143, sextupole.py:90-116, dipole.py:396-466*/
};

/* Per-setting record written by ch_compose_maps, `record_len` scalars of the beam dtype:
 *   [0]            sparsity flags (integer stored in the scalar's bit pattern, see
 *                  CH_FLAG_*), [1] sum of the element lengths of the section
 *   [2 .. 43]      rows 0-5 of the cumulative 7x7 map from the start of the section to
 *                  its end, row-major 6x7 (row 6 is always 0 0 0 0 0 0 1)
 *   then for each aperture a (in lattice order), 16 scalars:
 *     [0..6] row 0 (x) and [7..13] row 2 (y) of the cumulative map from the start of the
 *     section to the aperture, [14] x_max, [15] y_max
 *   and, if the section ends with an active cavity, CH_RECORD_CAVITY more scalars:
 *     [0..6] row 4 (tau) and [7..13] row 5 (delta) of the cumulative map up to the cavity
 *     ENTRANCE, [14] E0 b0 / (E1 b1), [15] V b0 / (E1 b1), [16] b0 k, [17] sin(phi),
 *     [18] cos(phi), [19] T566, [20] T556, [21] T555 (cavity.py:113-220); the 6x7 map of the
 *     record then includes the cavity's R matrix and ch_apply_maps replaces delta by the exact
 *     update and adds the second-order terms to tau.
 */
#define CH_RECORD_HEADER 2
#define CH_RECORD_MAP 42
#define CH_RECORD_APERTURE 16
#define CH_RECORD_LEN(n_apertures) (CH_RECORD_HEADER + CH_RECORD_MAP + CH_RECORD_APERTURE * (n_apertures))
#define CH_RECORD_CAVITY 24
#define CH_MAX_APERTURES 32

/* sparsity flags: a set bit means the named group of map entries is exactly zero in
 * every map of the record, so ch_apply_maps may skip its multiplies                     */
#define CH_FLAG_XY_UNCOUPLED 1u    /* (0..1 x 2..3) and (2..3 x 0..1) blocks             */
#define CH_FLAG_NO_TAU_COLUMN 2u   /* column 4 of rows 0-3 and row 5                     */
#define CH_FLAG_DELTA_IDENTITY 4u  /* row 5 is 0 0 0 0 0 1 0                             */
#define CH_FLAG_NO_Y_DISPERSION 8u /* (2..3, 5) and (4, 2..3)                            */

typedef struct ch_program ch_program;

/* library ------------------------------------------------------------------------- */
int ch_abi_version(void);
const char* ch_last_error(void);
/* number of kernels this library has launched since load (monotonic) */
int64_t ch_kernel_launch_count(void);

/* lattice program ----------------------------------------------------------------- */
/* Replaces the per-call walk over element objects in Segment.track
 * (cheetah/accelerator/segment.py:545-574).  Host arrays describe the flattened lattice:
 * op i uses slots [slot_begin[i], slot_begin[i+1]) (slot_begin has n_ops+1 entries).
 * Slot s reads setting b at  ((T*)slot_ptrs[s])[b * slot_strides[s]]  with T given by
 * slot_dtypes[s]; stride 0 broadcasts one value to all settings.  The slot pointers are
 * borrowed: they must stay valid while the program is used (they are the live parameter
 * tensors of the element objects, so in-place updates are seen without re-lowering).  */
int ch_program_create(const int32_t* opcodes_host, const int32_t* op_flags_host,
                      const int32_t* slot_begin_host, int32_t n_ops,
                      const void* const* slot_ptrs_host, const int64_t* slot_strides_host,
                      const int32_t* slot_dtypes_host, int32_t n_slots,
                      void* stream, ch_program** program_out);
int ch_program_destroy(ch_program* program);

/* map composition ----------------------------------------------------------------- */
/* Replaces Element.first_order_transfer_map for every element type
 * (cheetah/track_methods.py:17-77, :284-382 and the wrappers listed at the opcodes) and
 * the merged-map product of Segment.first_order_transfer_map
 * (cheetah/accelerator/segment.py:534-541), for ops [op_begin, op_end) of `program`,
 * for n_settings settings.  One thread per setting walks the ops, keeps the cumulative
 * 6x7 map in fp64 registers, and writes one record (layout above) per setting, rounded
 * once to `record_dtype`.  energy / mass_eV are read like slots
 * (energy[b * energy_stride]); num_elementary_charges (a device scalar, may be NULL when
 * the program has no active cavity) is the species' charge in e.  gamma, beta follow
 * cheetah/utils/physics.py:4-19.                                                       */
int ch_compose_maps(const ch_program* program, int32_t op_begin, int32_t op_end,
                    int64_t n_settings,
                    const void* energy, int64_t energy_stride, int32_t energy_dtype,
                    const void* mass_eV, int32_t mass_dtype,
                    const void* num_elementary_charges, int32_t charge_dtype,
                    void* records, int64_t record_len, int32_t record_dtype,
                    void* stream);

/* map application ----------------------------------------------------------------- */
/* Replaces, in ONE pass over the particles, every `particles @ tm.mT` of
 * Element._track_first_order (cheetah/accelerator/element.py:181-191) and every
 * Aperture.track mask update (cheetah/accelerator/aperture.py:108-132) of one linear
 * section of the lattice.
 *
 *   particles_out[b, n, :] = M_b . particles_in[pidx(b), n, :]
 *   survival_out[b, n]     = survival_in[sidx(b), n] * prod_a mask_a(x_a, y_a)
 *
 * where M_b and the aperture rows come from records[ridx(b)], and for X in {p, r, s}
 * Xidx(b) = (X_index ? X_index[b] : b) * X_stride (stride 0 shares one entry among all
 * settings; strides are in units of one batch entry).  Rectangular masks use strict
 * comparisons, elliptical ones x^2/x_max^2 + y^2/y_max^2 <= 1, evaluated with IEEE
 * round-to-nearest operations in the beam dtype exactly as the reference does.
 * record_len is CH_RECORD_LEN(n_apertures), plus CH_RECORD_CAVITY when the section ends with
 * an active cavity.  survival_out may be NULL iff n_apertures == 0.  unit_seventh != 0 asserts that
 * particles_in[..., 6] == 1 (then the kernel adds column 6 instead of multiplying).   */
int ch_apply_maps(const void* particles_in, int64_t particle_stride, const int32_t* particle_index,
                  const void* survival_in, int64_t survival_stride, const int32_t* survival_index,
                  const void* records, int64_t record_stride, const int32_t* record_index,
                  int64_t record_len, int32_t n_apertures, uint32_t elliptical_mask,
                  int64_t n_particles, int64_t n_settings,
                  void* particles_out, void* survival_out,
                  int32_t dtype, int32_t unit_seventh, void* stream);

/* ch_apply_maps for results that are about to cross PCIe (the host-buffer path, host.py): the
 * seventh phase-space column is the constant 1 of cheetah/particles/particle_beam.py:60-106 and is
 * not written -- coordinates_out is [n_settings][n_particles][6] (24 instead of 28 bytes per row
 * for float32) -- and the survival probabilities can leave as a byte mask
 * survival_mask_out[n_settings][n_particles] in {0, 1} (exact whenever survival_in is NULL or holds
 * only zeros and ones, as after any chain of Aperture.track calls on a fresh beam,
 * aperture.py:108-132) instead of / next to survival_out in the beam dtype: 25 B per (particle,
 * setting) instead of 32.  Requires particles_in[..., 6] == 1 and a section without a cavity tail;
 * either survival output may be NULL. */
int ch_apply_maps_compact(const void* particles_in, int64_t particle_stride, const int32_t* particle_index,
                          const void* survival_in, int64_t survival_stride, const int32_t* survival_index,
                          const void* records, int64_t record_stride, const int32_t* record_index,
                          int64_t record_len, int32_t n_apertures, uint32_t elliptical_mask,
                          int64_t n_particles, int64_t n_settings,
                          void* coordinates_out, void* survival_out, uint8_t* survival_mask_out,
                          int32_t dtype, void* stream);

/* ParameterBeam tracking through one linear section (cheetah/accelerator/element.py:166-179):
 *   mu_out[b] = M_b mu_in[midx(b)],  cov_out[b] = M_b cov_in[midx(b)] M_b^T
 * with M_b the 7x7 map of records[ridx(b)] (products accumulated in fp64); mu [..][7],
 * cov [..][7][7]; index / stride conventions as in ch_apply_maps (mu_index serves both).
 * cavity_offset: position of the CH_RECORD_CAVITY block inside a record when the section ends with
 * an active cavity (CH_RECORD_LEN(n_apertures)), else -1; with it the longitudinal entries follow
 * the ParameterBeam branch of Cavity.track (cavity.py:129-135, :203-217).                       */
int ch_apply_maps_parameter(const void* mu_in, int64_t mu_stride, const int32_t* mu_index,
                            const void* cov_in, int64_t cov_stride,
                            const void* records, int64_t record_stride, const int32_t* record_index,
                            int32_t cavity_offset, int64_t n_settings,
                            void* mu_out, void* cov_out, int32_t dtype, void* stream);

/* ch_apply_maps with a fused observables epilogue: additionally
 * accumulates, per setting, the survival-weighted sums that ParticleBeam.mu_* / sigma_*
 * (cheetah/particles/particle_beam.py:1699-1805 -> cheetah/utils/statistics.py:30-62) are made
 * of, for the OUTGOING coordinates u_i (i = 0..5), about a pilot c_i (the image of particle 0):
 *   moments_out[b][0] = sum w, [1] = sum w^2, [2+i] = sum w (u_i - c_i),
 *   [8+i] = sum w (u_i - c_i)^2, [14+i] = c_i               (w = outgoing survival)
 * so mu_i = c_i + S1_i / S0 and var_i = (S2_i - S1_i^2 / S0) / (S0 - sum w^2 / S0).
 * particles_out and survival_out may be NULL: then nothing but the 20 doubles per setting
 * leaves the SM (no (B, N, 7) array in HBM); float32 sections without a cavity tail then run the
 * packed-pair kernel (observe_maps_kernel, two particles per FFMA2).  moments_out is zeroed here. */
#define CH_MOMENTS 20
int ch_apply_maps_moments(const void* particles_in, int64_t particle_stride, const int32_t* particle_index,
                          const void* survival_in, int64_t survival_stride, const int32_t* survival_index,
                          const void* records, int64_t record_stride, const int32_t* record_index,
                          int64_t record_len, int32_t n_apertures, uint32_t elliptical_mask,
                          int64_t n_particles, int64_t n_settings,
                          void* particles_out, void* survival_out, double* moments_out,
                          int32_t dtype, int32_t unit_seventh, void* stream);

/* non-linear tracking ------------------------------------------------------------------- */
/* A RUN of consecutive CH_OP_DKD_* / CH_OP_SECOND_ORDER ops (CH_OP_IDENTITY ops in between are
 * skipped) is tracked in ONE pass over the particles: 28 B read + 28 B written per (particle,
 * setting) for the whole run.  Replaces Drift/Quadrupole/Dipole._track_drift_kick_drift,
 * TransverseDeflectingCavity._track_drift_kick_drift and Element._track_second_order (files
 * cited at the opcodes).  Per setting the constants are
 *   [CH_NL_HEADER]  p0c, mc^2, E0, beta0, (mc^2 / E0)^2, sum of the element lengths, charge,
 *                   1 / p0c, 2 E0 / p0c, 0, 0, 0
 *   then one block per op: CH_NL_BLOCK_SECOND_ORDER scalars if the run holds a second-order op,
 *   else CH_NL_BLOCK_DKD (ch_nonlinear_constants_len gives the total).
 * Second-order block (what Element.second_order_transfer_map, element.py:134-147, is assembled
 * from on the host): [0] cos tilt, [1] sin tilt, [2..3] entry offsets, [4..5] entrance edge kicks
 * (px += k x, py += k y), [6..7] exit edge kicks, [8..9] exit offsets, [10..18] the body's
 * R: cx, sx, R10, cy, sy, R32, R05 (= R41), R15 (= R40), R45, [19..57] T_ijk in the order
 * 000 001 011 005 015 055 022 023 033 | 100 101 111 105 115 155 122 123 133 | 202 203 212 213
 * 225 235 | 302 303 312 313 325 335 | 400 401 411 405 415 455 422 423 433 (track_methods.py:147-279). */
#define CH_NL_HEADER 12
#define CH_NL_BLOCK_DKD 16
#define CH_NL_BLOCK_SECOND_ORDER 64
#define CH_NL_MAX_OPS 64

/* scalars per setting that ch_nonlinear_constants writes for ops [op_begin, op_end); -1 on error */
int64_t ch_nonlinear_constants_len(const ch_program* program, int32_t op_begin, int32_t op_end);

/* Element parameters -> per-(setting, op) constants in fp64 (sin / cos of tilts and bend angles,
 * fringe kicks, body R and the 39 second-order coefficients of base_ttensor with the compound
 * functions of cheetah/utils/autograd.py).  The particle pass rounds them to the beam dtype,
 * except for the bend body and the TDC kick, which it evaluates in fp64.  energy / mass / charge
 * are read as in ch_compose_maps.                                                           */
int ch_nonlinear_constants(const ch_program* program, int32_t op_begin, int32_t op_end,
                           int64_t n_settings,
                           const void* energy, int64_t energy_stride, int32_t energy_dtype,
                           const void* mass_eV, int32_t mass_dtype,
                           const void* num_elementary_charges, int32_t charge_dtype,
                           double* constants, void* stream);

/* particles_out[b, n, :] = (op_{end-1} o ... o op_begin)(particles_in[pidx(b), n, :]) with the
 * constants of setting cidx(b); index / stride conventions as in ch_apply_maps
 * (constants_stride = ch_nonlinear_constants_len, or 0 for one shared setting).  The beam
 * energy is unchanged (the reference recomputes sqrt(p0c^2 + m^2), equal up to rounding).   */
int ch_track_nonlinear(const ch_program* program, int32_t op_begin, int32_t op_end,
                       const double* constants, int64_t constants_stride,
                       const int32_t* constants_index,
                       const void* particles_in, int64_t particle_stride,
                       const int32_t* particle_index,
                       int64_t n_particles, int64_t n_settings, void* particles_out,
                       int32_t dtype, void* stream);

/* diagnostics -------------------------------------------------------------------------- */
/* Image recorded by an active Screen for a ParticleBeam: Screen.reading
 * (cheetah/accelerator/screen.py:241-344) with the read beam shifted by the screen misalignment
 * (:199-215), weights |particle_charges| * survival, extent -+ resolution * pixel_size / 2
 * (:137-146) and bins = resolution / binning.  method 0: 2-D cloud-in-cell deposit
 * (cheetah/utils/cloud_in_cell.py:129-241); method 1: torch.histogramdd semantics on the given
 * bin edges (edges_x [nx + 1], edges_y [ny + 1]: Screen.pixel_bin_edges, :148-166).  All arrays in
 * the beam dtype; misalignment [B or 1][2] (stride 0 or 2), pixel_size [2].  image
 * [B][resolution_y / binning][resolution_x / binning] (height, width) is zeroed here.         */
int ch_screen_image(const void* particles, int64_t particle_stride,
                    const void* charges, int64_t charge_stride,
                    const void* survival, int64_t survival_stride,
                    const void* misalignment, int64_t misalignment_stride,
                    const void* pixel_size,
                    int32_t resolution_x, int32_t resolution_y, int32_t binning, int32_t method,
                    const void* edges_x, const void* edges_y,
                    int64_t n_particles, int64_t n_beams, int32_t dtype,
                    void* image, void* stream);

/* Screen.reading with method="kde" (screen.py:312-326 -> cheetah/utils/kde.py:160-204): Gaussian
 * kernel density image on the pixel centres centers_x [nx], centers_y [ny]
 * (Screen.pixel_bin_centers, screen.py:168-174; uniformly spaced) with bandwidth[0] (device scalar,
 * Screen.kde_bandwidth), weights |particle_charges| * survival on the x factor only, kernel values
 * clamped to the smallest normal number, image normalised by (its sum + 1e-10).  Every particle is
 * spread over the pixels within 7 (float32) / 8 (float64) bandwidths, beyond which the kernel is
 * below the rounding of the sums.  totals: [B] doubles (zeroed here; the un-normalised sums on
 * return).  image [B][ny][nx] (height, width); other conventions as ch_screen_image.             */
int ch_screen_kde(const void* particles, int64_t particle_stride,
                  const void* charges, int64_t charge_stride,
                  const void* survival, int64_t survival_stride,
                  const void* misalignment, int64_t misalignment_stride,
                  const void* centers_x, int32_t nx, const void* centers_y, int32_t ny,
                  const void* bandwidth, int64_t n_particles, int64_t n_beams, int32_t dtype,
                  void* image, double* totals, void* stream);

/* Screen.reading for a (non-vectorised) ParameterBeam (screen.py:251-289): the bivariate normal
 * density of (x, y) ~ N((mu[0], mu[2]) - misalignment, cov[{0,2}][{0,2}]) at the grid points
 * float32(left + i * step_x), float32(bottom + j * step_y) -- torch.arange(left, right, step) with
 * tensor bounds yields float32 points computed in double, whatever the screen dtype; the caller
 * passes nx = ceil((right - left) / step_x) likewise.  mu [7], cov [7][7], misalignment [2] in the
 * beam dtype; image [ny][nx].                                                                  */
int ch_screen_gaussian(const void* mu, const void* cov, const void* misalignment,
                       double left, double step_x, int32_t nx,
                       double bottom, double step_y, int32_t ny,
                       int32_t dtype, void* image, void* stream);

/* ch_apply_maps_moments with the full second-moment matrix: moments_out[b] has CH_MOMENTS_COV
 * doubles, the first CH_MOMENTS as above, then [20 + k] = sum w (u_i - c_i)(u_j - c_j) for the 15
 * pairs i < j in the order (0,1), (0,2), ..., (0,5), (1,2), ..., (4,5), [35] unused.  With
 * cov_ij = (S_ij - S1_i S1_j / S0) / (S0 - sum w^2 / S0) this is
 * unbiased_weighted_covariance_matrix (cheetah/utils/statistics.py:65-88) of the outgoing beam,
 * i.e. ParticleBeam.as_parameter_beam's cov and the cov_xpx / cov_ypy / cov_taup properties.   */
#define CH_MOMENTS_COV 36
int ch_apply_maps_covariance(const void* particles_in, int64_t particle_stride, const int32_t* particle_index,
                             const void* survival_in, int64_t survival_stride, const int32_t* survival_index,
                             const void* records, int64_t record_stride, const int32_t* record_index,
                             int64_t record_len, int32_t n_apertures, uint32_t elliptical_mask,
                             int64_t n_particles, int64_t n_settings,
                             void* particles_out, void* survival_out, double* moments_out,
                             int32_t dtype, int32_t unit_seventh, void* stream);

/* space charge ----------------------------------------------------------------------- */
/* One SpaceChargeKick (cheetah/accelerator/space_charge_kick.py:477-586) is the sequence
 *   ch_sc_beam_moments -> ch_sc_grid_params -> ch_sc_deposit -> ch_sc_green_function ->
 *   ch_sc_green_spectrum -> ch_sc_poisson_solve -> ch_sc_field -> ch_sc_gather_kick
 * over B independent beams (the flattened vector dims, space_charge_kick.py:493-528).
 * All per-beam scalars live on the device in fp64 tables so that no step synchronises
 * with the host:
 *   stats  [B][CH_SC_STATS]   0 sum w, 1 sum w^2, 2-4 sum w (u - u0), 5-7 sum w (u - u0)^2
 *                             for u = x, y, tau; 8-10 the pilot u0 (particle 0), 11 ticket
 *                             counter of ch_sc_moments_and_params
 *   params [B][CH_SC_PARAMS]  0-2 grid half-extent (extent * sigma), 3-5 cell size,
 *                             6 gamma, 7 beta, 8 dt = L / (c beta), 9 1 / cell volume,
 *                             10 1 / gamma^2 (0 if gamma == 0), 11-13 sigma x, y, tau,
 *                             14 sum w, 15 mass in eV
 * Grid sizes nx, ny, nz: any value in [2, 256].  The FFT length of an axis is L = the next power
 * of two >= 2 n (>= 8; the convolution is aperiodic, so any padding >= 2 n - 1 gives the same
 * potential), K = L / 2 + 1 its number of non-redundant frequencies.
 * Batch strides are in elements per beam; 0 shares one array among all beams.           */
#define CH_SC_STATS 12
#define CH_SC_PARAMS 24
/* params 16-23 (derived, written with the rest): 16 gamma beta, 17 1 / (gamma beta),
 * 18 e dt / (m c) (momentum change in units of m c per unit of field), 19-21 1 / cell size in
 * the beam dtype, 22-23 reserved.                                                           */

/* Layout of the field array read by ch_sc_gather_kick*:
 *   CH_SC_FIELD_NODES   ch_sc_field:        [B][nx*ny*nz][2][4], z corner pairs (any dtype)
 *   CH_SC_FIELD_BRICKS  ch_sc_field_bricks: [B][nx*ny*nz][3][8], all 8 corners of a cell side by
 *                       side (float32 only)                                                  */
#define CH_SC_FIELD_NODES 0
#define CH_SC_FIELD_BRICKS 1

/* Survival-weighted sums for the unbiased weighted standard deviations of x, y, tau:
 * ParticleBeam.sigma_{x,y,tau} (cheetah/particles/particle_beam.py:1709-1717, :1761-1765,
 * :1801-1805) -> unbiased_weighted_variance (cheetah/utils/statistics.py:30-48).  One pass,
 * fp64 accumulation about a pilot particle.  survival may be NULL (all ones).           */
int ch_sc_beam_moments(const void* particles, int64_t particle_stride,
                       const void* survival, int64_t survival_stride,
                       int64_t n_particles, int64_t n_beams, int32_t dtype,
                       double* stats, void* stream);

/* ch_sc_beam_moments + ch_sc_grid_params in ONE launch: the last CTA to finish a beam's sums
 * computes its parameters (saves a dependent launch per kick).  Arguments as in the two
 * functions below / above.                                                                */
int ch_sc_moments_and_params(const void* particles, int64_t particle_stride,
                             const void* survival, int64_t survival_stride,
                             int64_t n_particles, int64_t n_beams,
                             const void* energy, int64_t energy_stride, int32_t energy_dtype,
                             const void* mass_eV, int32_t mass_dtype,
                             const void* effect_length, int64_t length_stride, int32_t length_dtype,
                             const void* extent_x, int64_t extent_x_stride,
                             const void* extent_y, int64_t extent_y_stride,
                             const void* extent_tau, int64_t extent_tau_stride, int32_t extent_dtype,
                             int32_t nx, int32_t ny, int32_t nz, int32_t dtype,
                             double* stats, double* params, void* stream);

/* sigma -> grid_dimensions, cell_size, dt, gamma, beta: space_charge_kick.py:531-550,
 * cheetah/particles/beam.py:323-336.  energy / effect_length / extents are read like
 * slots: value(b) = ptr[b * stride] with the given dtype.                               */
int ch_sc_grid_params(const double* stats, int64_t n_beams,
                      const void* energy, int64_t energy_stride, int32_t energy_dtype,
                      const void* mass_eV, int32_t mass_dtype,
                      const void* effect_length, int64_t length_stride, int32_t length_dtype,
                      const void* extent_x, int64_t extent_x_stride,
                      const void* extent_y, int64_t extent_y_stride,
                      const void* extent_tau, int64_t extent_tau_stride, int32_t extent_dtype,
                      int32_t nx, int32_t ny, int32_t nz, int32_t dtype,
                      double* params, void* stream);

/* Cloud-in-cell deposit of charge * survival at (x, y, -beta tau) on the (nx, ny, nz) grid
 * spanning +-grid half-extent: _array_rho (space_charge_kick.py:125-146) ->
 * _cloud_in_cell_3d (cheetah/utils/cloud_in_cell.py:244-384; cell-centred, inclusive
 * extent test, clamped corners with zeroed weights).  rho is zeroed here and holds CHARGE per
 * cell (the 1/cell-volume factor is applied in ch_sc_poisson_solve) in QUAD BLOCKS
 * rho[B][nx][4][ny/2 + 1][nz/2 + 1][4]: part p = 2 py + pz holds the 2 x 2 (y, z) blocks whose
 * lower corner (y0, z0) has parity (py, pz) (-1 counts as odd) at block ((y0 + py) / 2,
 * (z0 + pz) / 2), entries (dy, dz) = (0,0), (0,1), (1,0), (1,1).  The four (y, z) corners of a
 * particle are then one 16-byte aligned vector reduction (REDG.E.ADD.F32x4): 2 L2 reductions per
 * particle.  Logical grid: the sum over the four parts (nx, ny, nz even).                      */
int ch_sc_deposit(const void* particles, int64_t particle_stride,
                  const void* charges, int64_t charge_stride,
                  const void* survival, int64_t survival_stride,
                  const double* params, int64_t n_particles, int64_t n_beams,
                  int32_t nx, int32_t ny, int32_t nz, int32_t dtype,
                  void* rho, void* stream);

/* General 3-D cloud-in-cell deposit: cloud_in_cell_charge_deposition
 * (cheetah/utils/cloud_in_cell.py:8-64) for positions [B][N][3], extent [B][3][2]
 * (left, right per axis), charges [B][N] (NULL = ones); grid [B][nx*ny*nz] is zeroed here. */
int ch_cic_deposit3d(const void* positions, const void* extent, const void* charges,
                     int64_t n_particles, int64_t n_beams,
                     int32_t nx, int32_t ny, int32_t nz, int32_t dtype,
                     void* grid, void* stream);

/* The same for 1, 2 or 3 position dimensions (`dims`; the 1-D / 2-D specialisations,
 * cheetah/utils/cloud_in_cell.py:67-241): positions [B][N][dims], extent [B][dims][2], grid
 * [B][nx (* ny (* nz))]; unused bin counts are ignored.                                         */
int ch_cic_deposit(const void* positions, const void* extent, const void* charges,
                   int64_t n_particles, int64_t n_beams, int32_t dims,
                   int32_t nx, int32_t ny, int32_t nz, int32_t dtype,
                   void* grid, void* stream);

/* ch_sc_moments_and_params with the per-CTA sums added in a fixed order by the last CTA instead
 * of with float64 atomics (deterministic mode; partials: n_beams * ceil(N / 1024) * 8 doubles). */
int ch_sc_moments_and_params_deterministic(
    const void* particles, int64_t particle_stride, const void* survival, int64_t survival_stride,
    int64_t n_particles, int64_t n_beams, const void* energy, int64_t energy_stride,
    int32_t energy_dtype, const void* mass_eV, int32_t mass_dtype, const void* effect_length,
    int64_t length_stride, int32_t length_dtype, const void* extent_x, int64_t extent_x_stride,
    const void* extent_y, int64_t extent_y_stride, const void* extent_tau,
    int64_t extent_tau_stride, int32_t extent_dtype, int32_t nx, int32_t ny, int32_t nz,
    int32_t dtype, double* partials, double* stats, double* params, void* stream);

/* Deterministic variants of ch_sc_deposit and ch_cic_deposit: the same deposits accumulated in
 * 64-bit fixed point (integer atomics commute, so the result does not depend on the order in
 * which the hardware retires them: bit-identical from run to run).  The reference's
 * scatter_add_ is order-dependent on CUDA; its tests ask torch for deterministic algorithms
 * (tests/conftest.py:204), and the Python layer selects these entry points when
 * torch.are_deterministic_algorithms_enabled().  One unit is ~2^-41 of the largest particle
 * charge.  scratch: n_beams * (cells + 1) * 8 bytes.  About 3x the time of the float atomics. */
int ch_sc_deposit_deterministic(const void* particles, int64_t particle_stride,
                                const void* charges, int64_t charge_stride,
                                const void* survival, int64_t survival_stride,
                                const double* params, int64_t n_particles, int64_t n_beams,
                                int32_t nx, int32_t ny, int32_t nz, int32_t dtype,
                                void* scratch, void* rho, void* stream);
int ch_cic_deposit_deterministic(const void* positions, const void* extent, const void* charges,
                                 int64_t n_particles, int64_t n_beams, int32_t dims,
                                 int32_t nx, int32_t ny, int32_t nz, int32_t dtype,
                                 void* scratch, void* grid, void* stream);

/* Integrated Green function: _integrated_potential + _integrated_green_function
 * (space_charge_kick.py:103-123, :163-291).  The antiderivative is evaluated ONCE per
 * half-shifted lattice point in fp64 (lattice [B][(nx+1)(ny+1)(nz+1)] doubles; the reference
 * evaluates it 8 n^3 times).  If `green` is not NULL the differenced, mirrored array
 * [B][2nx][2ny][2nz] of the reference is also written (plane n of every axis zero) -- used by
 * the parity tests; the solver itself only needs the lattice.  d_tau is scaled by gamma.
 * float32 beams: lattice points further than 5 x the largest cell size from the origin are not
 * written -- their Green function is the 4th-order far-field series of the cell integral
 * (2-3e-7 of the value in float32), evaluated by the consumers of the lattice.               */
int ch_sc_green_function(const double* params, int64_t n_beams,
                         int32_t nx, int32_t ny, int32_t nz, int32_t dtype,
                         double* lattice, void* green, void* stream);

/* rfftn of the mirrored Green array without ever building it: the array is even in every axis,
 * so its spectrum is real and even and is stored compactly as spectrum [B][Kx][Ky][Kz]
 * (3 passes of packed real-even FFTs, ~4x less work than a general rfftn).
 * scratch: B * (nx*ny*Kz + nx*Ky*Kz) scalars of the beam dtype.  `params` as given
 * to ch_sc_green_function (cell sizes decide where the far-field series applies).            */
int ch_sc_green_spectrum(const double* lattice, const double* params, int64_t n_beams,
                         int32_t nx, int32_t ny, int32_t nz, int32_t dtype,
                         void* scratch, void* spectrum, void* stream);

/* phi = irfftn(rfftn(rho_padded) * rfftn(green)) / (4 pi eps0 * cell volume), cropped to
 * the physical octant: _solve_poisson_equation (space_charge_kick.py:293-322).  Hand-written
 * shared-memory FFT passes (no cuFFT): z real<->complex with two rows packed per transform,
 * y and x strided passes; zero padding is never materialised; the x pass fuses forward FFT,
 * the multiply by the compact Green spectrum and the inverse FFT.
 * rho: the quad-block charge grid written by ch_sc_deposit (its four parts are summed while
 * loading);
 * rho_spectrum: [B][Lx][Ly][Kz] complex scratch.                                        */
int ch_sc_poisson_solve(const void* rho, const void* green_spectrum, const double* params,
                        int64_t n_beams, int32_t nx, int32_t ny, int32_t nz, int32_t dtype,
                        void* rho_spectrum, void* phi, void* stream);

/* -(1/gamma^2) grad(phi) by central differences, zero on the boundary cells:
 * _E_plus_vB_field (space_charge_kick.py:324-365).  field [B][nx*ny*nz][2][4]: per cell
 * (gx, gy, gz, 0) of the cell itself and of its +z neighbour (zero beyond the grid), so that
 * the gather reads a z corner pair as one 32-byte sector.                                   */
int ch_sc_field(const void* phi, const double* params, int64_t n_beams,
                int32_t nx, int32_t ny, int32_t nz, int32_t dtype, void* field, void* stream);

/* The same field as "bricks" (float32 only): bricks [B][nx][ny][nz][3][8] holds, per cell
 * (cx, cy, cz), component s of the field at the eight nodes (cx + dx, cy + dy, cz + dz), corner
 * index 4 dx + 2 dy + dz, zero for nodes beyond the grid: the 96 bytes one particle of that cell
 * gathers, contiguous (3 sectors in 1-2 cache lines instead of 4 sectors in 4 lines).       */
int ch_sc_field_bricks(const void* phi, const double* params, int64_t n_beams,
                       int32_t nx, int32_t ny, int32_t nz, int32_t dtype, void* bricks,
                       void* stream);

/* Node-centred trilinear gather of the field at the 8 surrounding grid points x elementary
 * charge (_compute_forces, space_charge_kick.py:367-475), momentum kick P += F dt
 * (:557-565) and both coordinate conversions ParticleBeam.to_xyz_pxpypz /
 * from_xyz_pxpypz (cheetah/particles/particle_beam.py:1262-1346), fused: one read and one
 * write of the particles.  float64 beams evaluate the SI round trip as the reference does;
 * float32 beams use the algebraically equal difference form (the change of each coordinate;
 * the reference's fp32 version squares momenta of 1e-20 kg m/s into the subnormal range,
 * SURVEY.md 7.3).  `field` / `field_layout`: see CH_SC_FIELD_*.
 * forces_out (optional, [B][N][3]) receives the interpolated forces for tests.          */
int ch_sc_gather_kick(const void* particles_in, int64_t particle_stride,
                      const void* field, int32_t field_layout, const double* params,
                      int64_t n_particles, int64_t n_beams,
                      int32_t nx, int32_t ny, int32_t nz, int32_t dtype,
                      void* particles_out, void* forces_out, void* stream);

/* ch_sc_gather_kick with the neighbouring stages of the lattice fused in, while the particles are
 * in registers (both parts optional, at least one must be given):
 *  - records != NULL: the linear section that FOLLOWS the kick (a skippable run without apertures
 *    or cavities): particles_out = M . kicked, M the 6x7 map of the ch_compose_maps record
 *    records[b * record_stride] (stride 0 = one map for all beams).  Replaces one ch_apply_maps
 *    pass (cheetah/accelerator/element.py:181-191).
 *  - next_stats != NULL: the survival-weighted sums (about the origin) of the NEXT
 *    SpaceChargeKick on the outgoing coordinates and, from the last CTA of each beam, its grid
 *    parameters next_params -- what ch_sc_moments_and_params would compute in a separate pass
 *    (space_charge_kick.py:531-550).  next_* describe that next kick; energy / mass as there.    */
int ch_sc_gather_kick_fused(const void* particles_in, int64_t particle_stride,
                            const void* field, int32_t field_layout, const double* params,
                            int64_t n_particles, int64_t n_beams,
                            int32_t nx, int32_t ny, int32_t nz, int32_t dtype,
                            const void* records, int64_t record_stride,
                            const void* survival, int64_t survival_stride,
                            double* next_stats, double* next_params,
                            const void* energy, int64_t energy_stride, int32_t energy_dtype,
                            const void* mass_eV, int32_t mass_dtype,
                            const void* next_effect_length, int64_t next_length_stride,
                            int32_t next_length_dtype,
                            const void* next_extent_x, int64_t next_extent_x_stride,
                            const void* next_extent_y, int64_t next_extent_y_stride,
                            const void* next_extent_tau, int64_t next_extent_tau_stride,
                            int32_t next_extent_dtype,
                            int32_t next_nx, int32_t next_ny, int32_t next_nz,
                            void* particles_out, void* stream);

/* The grid half of one kick in one call: ch_sc_green_function + ch_sc_green_spectrum on an
 * internal high-priority side stream next to ch_sc_deposit on `stream`, joined inside the
 * Poisson solve right before its x convolution, the first kernel that reads the Green spectrum
 * (the Python layer's four calls and two events join before ch_sc_poisson_solve; at one beam an
 * eager kick is bound by the host's call rate).  Buffers as in those four functions.          */
int ch_sc_solve(const void* particles, int64_t particle_stride,
                const void* charges, int64_t charge_stride,
                const void* survival, int64_t survival_stride,
                const double* params, int64_t n_particles, int64_t n_beams,
                int32_t nx, int32_t ny, int32_t nz, int32_t dtype,
                void* rho, double* lattice, void* green_scratch, void* green_spectrum,
                void* rho_spectrum, void* phi, void* stream);

/* float32: ch_sc_field_bricks + ch_sc_gather_kick[_fused] interleaved per group of `group_beams`
 * beams: the bricks of a group are built from phi [B][nx*ny*nz] into `bricks`
 * (group_beams * nx*ny*nz * 24 floats, reused by every group) right before the group's gather
 * pass, so the gather finds them in L2 instead of HBM (the bricks of all beams of a large batch
 * exceed the L2 several times) and they never travel to HBM at all.  records and next_stats may
 * both be NULL (plain kick); forces_out as in ch_sc_gather_kick; the other arguments as in
 * ch_sc_gather_kick_fused.                                                                      */
int ch_sc_field_gather(const void* particles_in, int64_t particle_stride,
                       const void* phi, void* bricks, int32_t group_beams, const double* params,
                       int64_t n_particles, int64_t n_beams,
                       int32_t nx, int32_t ny, int32_t nz, int32_t dtype,
                       const void* records, int64_t record_stride,
                       const void* survival, int64_t survival_stride,
                       double* next_stats, double* next_params,
                       const void* energy, int64_t energy_stride, int32_t energy_dtype,
                       const void* mass_eV, int32_t mass_dtype,
                       const void* next_effect_length, int64_t next_length_stride,
                       int32_t next_length_dtype,
                       const void* next_extent_x, int64_t next_extent_x_stride,
                       const void* next_extent_y, int64_t next_extent_y_stride,
                       const void* next_extent_tau, int64_t next_extent_tau_stride,
                       int32_t next_extent_dtype,
                       int32_t next_nx, int32_t next_ny, int32_t next_nz,
                       void* particles_out, void* forces_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CHEETAH_B200_H */

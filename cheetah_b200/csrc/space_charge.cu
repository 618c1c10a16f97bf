// SpaceChargeKick kernel family: moments -> grid parameters -> cloud-in-cell deposit ->
// integrated Green function -> FFT Poisson solve -> field -> gather + kick.
//
// Reference (desy-ml/cheetah @ 60d1053): cheetah/accelerator/space_charge_kick.py:103-586,
// cheetah/utils/cloud_in_cell.py:244-384, cheetah/particles/particle_beam.py:1262-1346,
// :1709-1805, cheetah/utils/statistics.py:30-62, cheetah/particles/beam.py:323-336.
//
// Everything is batched over B independent beams; per-beam scalars travel between the
// kernels in small fp64 device tables (include/cheetah_b200.h) so nothing synchronises with
// the host.  Particle passes read the 28-byte AoS rows through TMA-staged shared-memory
// tiles; the charge histogram lives in L2 (a 64^3 fp32 grid is 1 MB) and is built with
// fire-and-forget RED atomics (N/cells is ~4: privatising tiles would cost more than the
// atomics they save); the 3-D convolution is three shared-memory FFT passes.
#include <math_constants.h>

#include <atomic>
#include <cstdlib>
#include <type_traits>

#include "ch_common.cuh"
#include "fft.cuh"
#include "fft_regs.cuh"

namespace ch {
namespace {

constexpr double kSpeedOfLight = 299792458.0;
constexpr double kElementaryCharge = 1.602176634e-19;
constexpr double kEpsilon0 = 8.8541878188e-12;
constexpr double kEvToKg = 1.7826619216278975e-36;
constexpr double kPi = 3.14159265358979323846;

// ---------------------------------------------------------------------------------------
// 1. survival-weighted moments
// ---------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------
// 2. per-beam grid parameters
// ---------------------------------------------------------------------------------------
struct GridInputs {
  ScalarRef energy, mass, length, extent_x, extent_y, extent_tau;
};

template <typename T>
__device__ void grid_params_for_beam(const double* s, int64_t b, const GridInputs& in, int nx,
                                     int ny, int nz, double* out) {
  const double sum_w = s[0];
  const double correction = sum_w - s[1] / sum_w;  // statistics.py:44
  const double extent[3] = {load_scalar(in.extent_x.ptr, b * in.extent_x.stride, in.extent_x.dtype),
                            load_scalar(in.extent_y.ptr, b * in.extent_y.stride, in.extent_y.dtype),
                            load_scalar(in.extent_tau.ptr, b * in.extent_tau.stride,
                                        in.extent_tau.dtype)};
  const int n[3] = {nx, ny, nz};
  double volume = 1.0;
  for (int d = 0; d < 3; ++d) {
    // sum w (u - mean)^2 = S2 - S1^2 / S0 about the pilot
    const double centred = s[5 + d] - s[2 + d] * s[2 + d] / sum_w;
    // the reference carries sigma, grid_dimensions and cell_size in the beam dtype
    const T sigma = static_cast<T>(sqrt(fmax(centred, 0.0) / correction));
    const T half_extent = static_cast<T>(extent[d]) * sigma;
    const T cell = T(2) * half_extent / static_cast<T>(n[d]);
    out[d] = static_cast<double>(half_extent);
    out[3 + d] = static_cast<double>(cell);
    out[11 + d] = static_cast<double>(sigma);
    volume *= static_cast<double>(cell);
  }
  const double mass = load_scalar(in.mass.ptr, 0, in.mass.dtype);
  const double gamma = load_scalar(in.energy.ptr, b * in.energy.stride, in.energy.dtype) / mass;
  const double beta = (fabs(gamma) > 0.0 && gamma > 0.0) ? sqrt(1.0 - 1.0 / (gamma * gamma)) : 1.0;
  const double length = load_scalar(in.length.ptr, b * in.length.stride, in.length.dtype);
  out[6] = gamma;
  out[7] = beta;
  out[8] = length / (kSpeedOfLight * beta);
  out[9] = 1.0 / volume;
  out[10] = gamma != 0.0 ? 1.0 / (gamma * gamma) : 0.0;
  out[14] = sum_w;
  out[15] = mass;
  // constants of the float32 gather (sc_gather_brick_kernel): u = P / (m c) per unit of px,
  // du per unit of field (du = e E dt / (m c)) and the reciprocal cell sizes in the beam dtype
  const double mc = mass * kEvToKg * kSpeedOfLight;
  out[16] = gamma * beta;
  out[17] = gamma * beta != 0.0 ? 1.0 / (gamma * beta) : 0.0;
  out[18] = kElementaryCharge * out[8] / mc;
  for (int d = 0; d < 3; ++d) out[19 + d] = static_cast<double>(T(1) / static_cast<T>(out[3 + d]));
  out[22] = out[23] = 0.0;
}

template <typename T>
__global__ void sc_grid_params_kernel(const double* __restrict__ stats, int64_t n_beams,
                                      GridInputs in, int nx, int ny, int nz,
                                      double* __restrict__ params) {
  const int64_t b = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (b >= n_beams) return;
  grid_params_for_beam<T>(stats + b * CH_SC_STATS, b, in, nx, ny, nz, params + b * CH_SC_PARAMS);
}

// ---- 1 (continued). moments kernel (after the parameter helpers it may call) -----------
template <typename T>
__global__ void __launch_bounds__(256)
sc_moments_kernel(const T* __restrict__ particles, int64_t particle_stride,
                  const T* __restrict__ survival, int64_t survival_stride, int64_t n_particles,
                  int bulk_in, double* __restrict__ stats, GridInputs in, int nx, int ny, int nz,
                  double* __restrict__ params, double* __restrict__ partials) {
  constexpr int P = 4, THREADS = 256, TP = P * THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);
  __shared__ uint64_t bar;
  __shared__ double partial[8][8];
  const int64_t b = blockIdx.y;
  const T* p = particles + b * particle_stride;
  const T* w = survival ? survival + b * survival_stride : nullptr;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * TP;
  const int count = static_cast<int>(min(static_cast<int64_t>(TP), n_particles - n0));
  if (bulk_in && threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  uint32_t phase = 0;
  cta_load_tile(tile, p + n0 * 7, count * 7, bulk_in != 0, &bar, phase);
  // pilot = particle 0 of the beam: the sums are taken about it to avoid cancellation
  const double x0 = static_cast<double>(p[0]);
  const double y0 = static_cast<double>(p[2]);
  const double t0 = static_cast<double>(p[4]);

  double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const int local = threadIdx.x + k * THREADS;
    if (local < count) {
      const double wi = w ? static_cast<double>(w[n0 + local]) : 1.0;
      const double dx = static_cast<double>(tile[local * 7 + 0]) - x0;
      const double dy = static_cast<double>(tile[local * 7 + 2]) - y0;
      const double dt = static_cast<double>(tile[local * 7 + 4]) - t0;
      acc[0] += wi;
      acc[1] = fma(wi, wi, acc[1]);
      acc[2] = fma(wi, dx, acc[2]);
      acc[3] = fma(wi, dy, acc[3]);
      acc[4] = fma(wi, dt, acc[4]);
      acc[5] = fma(wi * dx, dx, acc[5]);
      acc[6] = fma(wi * dy, dy, acc[6]);
      acc[7] = fma(wi * dt, dt, acc[7]);
    }
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double s = warp_sum(acc[k]);
    if (lane == 0) partial[warp][k] = s;
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    double s = 0.0;
    for (int wi = 0; wi < 8; ++wi) s += partial[wi][threadIdx.x];
    // deterministic mode: the CTA sums are parked and added in CTA order by the last CTA
    if (partials != nullptr)
      partials[(b * gridDim.x + blockIdx.x) * 8 + threadIdx.x] = s;
    else
      atomicAdd(&stats[b * CH_SC_STATS + threadIdx.x], s);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    stats[b * CH_SC_STATS + 8] = x0;
    stats[b * CH_SC_STATS + 9] = y0;
    stats[b * CH_SC_STATS + 10] = t0;
  }
  if (params == nullptr && partials == nullptr) return;
  // the last CTA of this beam to finish turns the sums into the grid parameters, so the
  // chain needs no separate launch (threadfence-reduction pattern; stats[11] is the ticket)
  __shared__ bool last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0)
    last = atomicAdd(&stats[b * CH_SC_STATS + 11], 1.0) == static_cast<double>(gridDim.x - 1);
  __syncthreads();
  if (last && partials != nullptr) {
    if (threadIdx.x < 8) {
      double s = 0.0;
      for (unsigned c = 0; c < gridDim.x; ++c)
        s += __ldcg(&partials[(b * gridDim.x + c) * 8 + threadIdx.x]);
      stats[b * CH_SC_STATS + threadIdx.x] = s;
    }
    __threadfence();
    __syncthreads();
  }
  if (last && threadIdx.x == 0 && params != nullptr) {
    double sums[CH_SC_STATS];
    for (int i = 0; i < CH_SC_STATS; ++i) sums[i] = __ldcg(&stats[b * CH_SC_STATS + i]);
    grid_params_for_beam<T>(sums, b, in, nx, ny, nz, params + b * CH_SC_PARAMS);
  }
}


// ---------------------------------------------------------------------------------------
// 3. cloud-in-cell deposit
// ---------------------------------------------------------------------------------------
// One axis of cloud_in_cell.py:275-362, in the beam dtype like the reference.
template <typename T>
struct AxisDeposit {
  int base;        // unclamped lower corner index (floor of the bin-space position)
  int lo, hi;      // clamped corner indices
  T w_lo, w_hi;    // corner weights (0 when the corner is off-grid)
  bool inside;     // inclusive extent test
};

template <typename T>
__device__ __forceinline__ AxisDeposit<T> deposit_axis(T pos, T left, T right, int bins) {
  AxisDeposit<T> a;
  a.inside = (pos >= left) && (pos <= right);
  const T width = (right - left) / static_cast<T>(bins);
  const T q = (pos - left) / width - T(0.5);
  const T fl = floor(q);
  const T frac = q - fl;
  // clamp before the int conversion so that far-away particles cannot overflow
  const T lim = static_cast<T>(bins + 1);
  const int qi = static_cast<int>(fmin(fmax(fl, -lim), lim));
  a.base = qi;
  a.lo = min(max(qi, 0), bins - 1);
  a.hi = min(max(qi + 1, 0), bins - 1);
  a.w_lo = (qi >= 0 && qi < bins) ? T(1) - frac : T(0);
  a.w_hi = (qi + 1 >= 0 && qi + 1 < bins) ? frac : T(0);
  return a;
}

template <typename T>
__device__ __forceinline__ void deposit_particle(T* __restrict__ grid, T x, T y, T z, T charge,
                                                 const T* lo, const T* hi, int nx, int ny, int nz) {
  const AxisDeposit<T> ax = deposit_axis(x, lo[0], hi[0], nx);
  const AxisDeposit<T> ay = deposit_axis(y, lo[1], hi[1], ny);
  const AxisDeposit<T> az = deposit_axis(z, lo[2], hi[2], nz);
  if (!(ax.inside && ay.inside && az.inside)) return;  // masked_charges = charges * in_extent
  const int ix[2] = {ax.lo, ax.hi}, iy[2] = {ay.lo, ay.hi}, iz[2] = {az.lo, az.hi};
  const T wx[2] = {ax.w_lo, ax.w_hi}, wy[2] = {ay.w_lo, ay.w_hi}, wz[2] = {az.w_lo, az.w_hi};
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const T w = wx[a] * wy[b] * wz[c];
        if (w != T(0)) atomicAdd(&grid[(ix[a] * ny + iy[b]) * nz + iz[c]], charge * w);
      }
}

// Four cells (2 x 2 in y, z) in ONE L2 transaction (SASS REDG.E.ADD.F32x4): needs a 16-byte
// aligned address -- see the quad-block layout below.  The L2 processes ~200 G reductions / s
// whatever their width (tools/micro/red_vec.cu), so the count of reductions is what matters.
__device__ __forceinline__ void add_quad(float* address, float a, float b, float c, float d) {
  atomicAdd(reinterpret_cast<float4*>(address), make_float4(a, b, c, d));
}
__device__ __forceinline__ void add_quad(double* address, double a, double b, double c, double d) {
  atomicAdd(address, a);
  atomicAdd(address + 1, b);
  atomicAdd(address + 2, c);
  atomicAdd(address + 3, d);
}

// Space-charge deposit.  Each particle touches 2 x-planes x (2 x 2) cells in (y, z).  The grid is
// kept as QUAD BLOCKS  rho[B][nx][4][ny/2 + 1][nz/2 + 1][4]:  part p = 2 py + pz holds the 2 x 2
// blocks whose lower corner (y0, z0) has parity (py, pz) (-1 counts as odd), block
// ((y0 + py) / 2, (z0 + pz) / 2), entries (dy, dz) = (0,0), (0,1), (1,0), (1,1).  Every corner
// quadruple is therefore contiguous and 16-byte aligned and ONE vector RED serves four cells:
// 2 L2 reductions per particle instead of 8 scalar ones (4 with z pairs only).  The logical grid
// is the sum of the four parts; the z pass of the FFT adds them while loading (quad_value).
template <typename T>
__device__ __forceinline__ T quad_value(const T* __restrict__ plane, int y, int z, int qy, int qz) {
  // plane = rho[b][x]; the cell belongs to one block of each part
  T sum = T(0);
#pragma unroll
  for (int py = 0; py < 2; ++py) {
    const int by = (y + py) >> 1, dy = (y + py) & 1;
#pragma unroll
    for (int pz = 0; pz < 2; ++pz) {
      const int bz = (z + pz) >> 1, dz = (z + pz) & 1;
      sum += plane[((static_cast<int64_t>(py * 2 + pz) * qy + by) * qz + bz) * 4 + dy * 2 + dz];
    }
  }
  return sum;
}

template <typename T>
__global__ void __launch_bounds__(256)
sc_deposit_kernel(const T* __restrict__ particles, int64_t particle_stride,
                  const T* __restrict__ charges, int64_t charge_stride,
                  const T* __restrict__ survival, int64_t survival_stride,
                  const double* __restrict__ params, int64_t n_particles, int nx, int ny, int nz,
                  int bulk_in, T* __restrict__ rho) {
  constexpr int P = 4, THREADS = 256, TP = P * THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);
  __shared__ uint64_t bar;
  const int64_t b = blockIdx.y;
  const double* prm = params + b * CH_SC_PARAMS;
  const T lo[3] = {-static_cast<T>(prm[0]), -static_cast<T>(prm[1]), -static_cast<T>(prm[2])};
  const T hi[3] = {static_cast<T>(prm[0]), static_cast<T>(prm[1]), static_cast<T>(prm[2])};
  const T minus_beta = -static_cast<T>(prm[7]);
  const T* q = charges + b * charge_stride;
  const T* w = survival ? survival + b * survival_stride : nullptr;
  const int qy = ny / 2 + 1, qz = nz / 2 + 1;
  const int64_t plane = static_cast<int64_t>(4) * qy * qz * 4;  // scalars per x plane
  T* grid = rho + b * nx * plane;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * TP;
  const int count = static_cast<int>(min(static_cast<int64_t>(TP), n_particles - n0));

  if (bulk_in && threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  // charges and survival probabilities of this thread's particles: issued before the wait for
  // the tile, so that their DRAM latency overlaps the copy instead of stalling every iteration
  T weight[P];
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const int local = threadIdx.x + k * THREADS;
    weight[k] = local < count ? q[n0 + local] : T(0);
    if (w != nullptr && local < count) weight[k] *= w[n0 + local];
  }
  uint32_t phase = 0;
  cta_load_tile(tile, particles + b * particle_stride + n0 * 7, count * 7, bulk_in != 0, &bar,
                phase);
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const int local = threadIdx.x + k * THREADS;
    if (local >= count) continue;
    const T charge = weight[k];
    // positions are (x, y, z = -beta tau): particle_beam.py:1335
    const AxisDeposit<T> ax = deposit_axis(tile[local * 7 + 0], lo[0], hi[0], nx);
    const AxisDeposit<T> ay = deposit_axis(tile[local * 7 + 2], lo[1], hi[1], ny);
    const T z = tile[local * 7 + 4] * minus_beta;
    const AxisDeposit<T> az = deposit_axis(z, lo[2], hi[2], nz);
    if (!(ax.inside && ay.inside && az.inside)) continue;  // charges * in_extent
    // lower corner indices: in [-1, n - 1] for a particle inside the extent (off-grid corners
    // carry zero weight, so the clamps only guard against rounding at the very edge)
    const int iy = min(max(ay.base, -1), ny - 1);
    const int iz = min(max(az.base, -1), nz - 1);
    const int py = iy & 1, pz = iz & 1;  // -1 & 1 == 1
    const int64_t block =
        ((static_cast<int64_t>(py * 2 + pz) * qy + ((iy + py) >> 1)) * qz + ((iz + pz) >> 1)) * 4;
    const int ix[2] = {ax.lo, ax.hi};
    const T wx[2] = {ax.w_lo, ax.w_hi};
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      if (wx[a] == T(0)) continue;
      const T lower = wx[a] * ay.w_lo * charge, upper = wx[a] * ay.w_hi * charge;
      add_quad(grid + ix[a] * plane + block, lower * az.w_lo, lower * az.w_hi, upper * az.w_lo,
               upper * az.w_hi);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
cic_deposit3d_kernel(const T* __restrict__ positions, const T* __restrict__ extent,
                     const T* __restrict__ charges, int64_t n_particles, int nx, int ny, int nz,
                     T* __restrict__ out) {
  const int64_t b = blockIdx.y;
  const T* e = extent + b * 6;
  const T lo[3] = {e[0], e[2], e[4]};
  const T hi[3] = {e[1], e[3], e[5]};
  const T* p = positions + b * n_particles * 3;
  const T* q = charges ? charges + b * n_particles : nullptr;
  T* grid = out + b * static_cast<int64_t>(nx) * ny * nz;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_particles;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    deposit_particle(grid, p[i * 3 + 0], p[i * 3 + 1], p[i * 3 + 2], q ? q[i] : T(1), lo, hi, nx,
                     ny, nz);
}

// 1-D and 2-D general deposits (cloud_in_cell.py:67-241): same per-axis rule, 2 / 4 corners.
template <typename T, int D>
__global__ void __launch_bounds__(256)
cic_deposit_low_kernel(const T* __restrict__ positions, const T* __restrict__ extent,
                       const T* __restrict__ charges, int64_t n_particles, int nx, int ny,
                       T* __restrict__ out) {
  const int64_t b = blockIdx.y;
  const T* e = extent + b * 2 * D;
  const T* p = positions + b * n_particles * D;
  const T* q = charges ? charges + b * n_particles : nullptr;
  T* grid = out + b * static_cast<int64_t>(nx) * (D == 2 ? ny : 1);
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_particles;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const T charge = q ? q[i] : T(1);
    const AxisDeposit<T> ax = deposit_axis(p[i * D], e[0], e[1], nx);
    if (D == 1) {
      if (!ax.inside) continue;
      if (ax.w_lo != T(0)) atomicAdd(&grid[ax.lo], charge * ax.w_lo);
      if (ax.w_hi != T(0)) atomicAdd(&grid[ax.hi], charge * ax.w_hi);
    } else {
      const AxisDeposit<T> ay = deposit_axis(p[i * D + 1], e[2], e[3], ny);
      if (!(ax.inside && ay.inside)) continue;
      const int ix[2] = {ax.lo, ax.hi}, iy[2] = {ay.lo, ay.hi};
      const T wx[2] = {ax.w_lo, ax.w_hi}, wy[2] = {ay.w_lo, ay.w_hi};
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const T w = wx[a] * wy[c];
          if (w != T(0)) atomicAdd(&grid[ix[a] * ny + iy[c]], charge * w);
        }
    }
  }
}

// ---------------------------------------------------------------------------------------
// 3b. deterministic deposits (the reference's scatter_add_ is order-dependent on
// CUDA and its tests only ask torch for determinism, tests/conftest.py:204)
// ---------------------------------------------------------------------------------------
// Float atomics make a histogram depend on the order in which the hardware retires them.  The
// deterministic variants accumulate in 64-bit FIXED POINT instead: integer addition is
// associative, so any order gives the same bits.  Scale per beam: 2^(61 - ceil(log2 N)) /
// max |q_i| (found by a first pass with an integer atomicMax on the bit pattern), i.e. no sum
// of N contributions can overflow and one unit is ~2^-41 of the largest particle charge --
// finer than the float32 (and float64 after 1e6 additions) rounding of the atomic version.
// scratch per beam: cells int64 + 1 uint64.
struct FixedGrid {
  unsigned long long* cells;
  double scale;
  __device__ __forceinline__ void add(int64_t index, double value) const {
    atomicAdd(cells + index, static_cast<unsigned long long>(__double2ll_rn(value * scale)));
  }
};

__device__ __forceinline__ double fixed_scale(unsigned long long max_bits, int64_t n_particles) {
  const double max_abs = __longlong_as_double(static_cast<long long>(max_bits));
  int shift = 61;
  for (int64_t n = n_particles; n > 0; n >>= 1) --shift;
  return max_abs > 0.0 ? ldexp(1.0, shift) / max_abs : 0.0;
}

template <typename T>
__global__ void __launch_bounds__(256)
cic_max_charge_kernel(const T* __restrict__ charges, int64_t charge_stride,
                      const T* __restrict__ survival, int64_t survival_stride,
                      int64_t n_particles, unsigned long long* __restrict__ max_bits) {
  const int64_t b = blockIdx.y;
  const T* q = charges ? charges + b * charge_stride : nullptr;
  const T* w = survival ? survival + b * survival_stride : nullptr;
  double local = 0.0;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_particles;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    T v = q ? q[i] : T(1);
    if (w) v *= w[i];
    local = fmax(local, fabs(static_cast<double>(v)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local = fmax(local, __shfl_xor_sync(0xffffffffu, local, o));
  // non-negative doubles order like their bit patterns
  if ((threadIdx.x & 31) == 0)
    atomicMax(max_bits + b, static_cast<unsigned long long>(__double_as_longlong(local)));
}

// Space-charge deposit into the fixed-point grid [B][nx][ny][nz]: the arithmetic of
// sc_deposit_kernel up to the accumulation.
template <typename T>
__global__ void __launch_bounds__(256)
sc_deposit_fixed_kernel(const T* __restrict__ particles, int64_t particle_stride,
                        const T* __restrict__ charges, int64_t charge_stride,
                        const T* __restrict__ survival, int64_t survival_stride,
                        const double* __restrict__ params, int64_t n_particles, int nx, int ny,
                        int nz, const unsigned long long* __restrict__ max_bits,
                        unsigned long long* __restrict__ fixed) {
  const int64_t b = blockIdx.y;
  const double* prm = params + b * CH_SC_PARAMS;
  const T lo[3] = {-static_cast<T>(prm[0]), -static_cast<T>(prm[1]), -static_cast<T>(prm[2])};
  const T hi[3] = {static_cast<T>(prm[0]), static_cast<T>(prm[1]), static_cast<T>(prm[2])};
  const T minus_beta = -static_cast<T>(prm[7]);
  const T* p = particles + b * particle_stride;
  const T* q = charges + b * charge_stride;
  const T* w = survival ? survival + b * survival_stride : nullptr;
  const FixedGrid grid{fixed + b * static_cast<int64_t>(nx) * ny * nz,
                       fixed_scale(max_bits[b], n_particles)};
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_particles;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const T charge = w ? q[i] * w[i] : q[i];
    const AxisDeposit<T> ax = deposit_axis(p[i * 7 + 0], lo[0], hi[0], nx);
    const AxisDeposit<T> ay = deposit_axis(p[i * 7 + 2], lo[1], hi[1], ny);
    const AxisDeposit<T> az = deposit_axis(p[i * 7 + 4] * minus_beta, lo[2], hi[2], nz);
    if (!(ax.inside && ay.inside && az.inside)) continue;  // charges * in_extent
    const int ix[2] = {ax.lo, ax.hi}, iy[2] = {ay.lo, ay.hi}, iz[2] = {az.lo, az.hi};
    const T wx[2] = {ax.w_lo, ax.w_hi}, wy[2] = {ay.w_lo, ay.w_hi}, wz[2] = {az.w_lo, az.w_hi};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const T row = wx[a] * wy[c] * charge;  // as sc_deposit_kernel: (wx wy q) wz
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const T value = row * wz[d];
          if (value != T(0))
            grid.add((static_cast<int64_t>(ix[a]) * ny + iy[c]) * nz + iz[d],
                     static_cast<double>(value));
        }
      }
  }
}

// fixed point -> the quad-block charge grid of sc_deposit_kernel: every cell (y, z) appears once
// in part 0 (blocks with even lower corners); the other three parts stay zero.
template <typename T>
__global__ void __launch_bounds__(256)
sc_fixed_to_quad_kernel(const unsigned long long* __restrict__ fixed,
                        const unsigned long long* __restrict__ max_bits, int64_t n_particles,
                        int nx, int ny, int nz, T* __restrict__ rho) {
  const int64_t b = blockIdx.y;
  const int64_t cells = static_cast<int64_t>(nx) * ny * nz;
  const double scale = fixed_scale(max_bits[b], n_particles);
  const double inverse = scale > 0.0 ? 1.0 / scale : 0.0;
  const int qy = ny / 2 + 1, qz = nz / 2 + 1;
  const int64_t plane = static_cast<int64_t>(4) * qy * qz * 4;
  T* out = rho + b * nx * plane;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < cells;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int z = static_cast<int>(idx % nz), y = static_cast<int>((idx / nz) % ny);
    const int x = static_cast<int>(idx / (static_cast<int64_t>(nz) * ny));
    const double value = static_cast<double>(static_cast<long long>(fixed[b * cells + idx])) * inverse;
    out[x * plane + (static_cast<int64_t>(y >> 1) * qz + (z >> 1)) * 4 + (y & 1) * 2 + (z & 1)] =
        static_cast<T>(value);
  }
}

// General 1-, 2- and 3-D deposits into a fixed-point grid (per-axis rule of deposit_axis).
template <typename T>
__global__ void __launch_bounds__(256)
cic_deposit_fixed_kernel(const T* __restrict__ positions, const T* __restrict__ extent,
                         const T* __restrict__ charges, int64_t n_particles, int dims, int nx,
                         int ny, int nz, const unsigned long long* __restrict__ max_bits,
                         unsigned long long* __restrict__ fixed) {
  const int64_t b = blockIdx.y;
  const T* e = extent + b * 2 * dims;
  const T* p = positions + b * n_particles * dims;
  const T* q = charges ? charges + b * n_particles : nullptr;
  const int n[3] = {nx, dims > 1 ? ny : 1, dims > 2 ? nz : 1};
  const FixedGrid grid{fixed + b * static_cast<int64_t>(n[0]) * n[1] * n[2],
                       fixed_scale(max_bits[b], n_particles)};
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_particles;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    AxisDeposit<T> axis[3];
    bool inside = true;
    for (int d = 0; d < 3; ++d) {
      if (d < dims) {
        axis[d] = deposit_axis(p[i * dims + d], e[2 * d], e[2 * d + 1], n[d]);
        inside = inside && axis[d].inside;
      } else {
        axis[d].lo = axis[d].hi = 0;
        axis[d].w_lo = T(1);
        axis[d].w_hi = T(0);
      }
    }
    if (!inside) continue;
    const T charge = q ? q[i] : T(1);
    for (int a = 0; a < 2; ++a)
      for (int c = 0; c < 2; ++c)
        for (int d = 0; d < 2; ++d) {
          // the order of the products follows the atomic kernels: wx wy wz, then the charge
          T weight = a ? axis[0].w_hi : axis[0].w_lo;
          if (dims > 1) weight *= c ? axis[1].w_hi : axis[1].w_lo;
          if (dims > 2) weight *= d ? axis[2].w_hi : axis[2].w_lo;
          if ((dims < 2 && c) || (dims < 3 && d) || weight == T(0)) continue;
          const int64_t cell = (static_cast<int64_t>(a ? axis[0].hi : axis[0].lo) * n[1] +
                                (c ? axis[1].hi : axis[1].lo)) * n[2] + (d ? axis[2].hi : axis[2].lo);
          grid.add(cell, static_cast<double>(charge * weight));
        }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
cic_fixed_to_grid_kernel(const unsigned long long* __restrict__ fixed,
                         const unsigned long long* __restrict__ max_bits, int64_t n_particles,
                         int64_t cells, T* __restrict__ out) {
  const int64_t b = blockIdx.y;
  const double scale = fixed_scale(max_bits[b], n_particles);
  const double inverse = scale > 0.0 ? 1.0 / scale : 0.0;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < cells;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[b * cells + idx] = static_cast<T>(
        static_cast<double>(static_cast<long long>(fixed[b * cells + idx])) * inverse);
}

// ---------------------------------------------------------------------------------------
// 4. integrated Green function
// ---------------------------------------------------------------------------------------
// Antiderivative F of 1/r (space_charge_kick.py:103-123), fp64, in a form that costs a quarter of
// the reference's expression (3 atan + 3 asinh + 4 sqrt, ~970 fp64 instructions per point):
//   * the Green function is the 8-corner signed difference of F over a cell (:195-236), which
//     annihilates every term that does not depend on all three coordinates.  With
//     asinh(x / rho) = log(x + r) - log(rho) the three rho terms are such terms, and so is the
//     (pi/4) x^2 left over when the third arctangent is replaced through
//     atan(xy/(tr)) + atan(xt/(yr)) + atan(yt/(xr)) = pi/2  (x, y, t > 0).  What remains is
//       F^ = (x^2 - t^2)/2 atan(xy/(tr)) + (x^2 - y^2)/2 atan(xt/(yr))
//            + y t log(x + r) + x t log(y + r) + x y log(t + r):
//     one sqrt, one division, two atan, three log.
//   * F is odd in each coordinate and the lattice point with index 0 sits at -1/2 cell.  The
//     dropped terms D = d1(y,t) + d2(x,t) + d3(x,y) + d4(x) keep cancelling across that plane only
//     if they are continued as functions of their own arguments, so lattice = F - D with the true
//     odd F gives, with s = sx sy st, the corrections (s - sy st) d1 + ... + (s - 1) d4 on the
//     index-0 planes (O(n^2) points).
//   * coordinates are taken in units of the grid diagonal R: the dropped logarithms then vanish
//     where the differences cancel most (far cells), and the rounding error of the differences
//     stays at the level of the reference's own fp64 expression (2e-9 of a far-corner value at
//     64^3, measured against 80-bit arithmetic; 8e-10 for the reference form).
// lattice[i][j][k] = R^2 F^(|i - 1/2| cx, |j - 1/2| cy, |k - 1/2| ct) (+ corrections), c = cell / R.
__device__ __forceinline__ double igf_lattice_value(int i, int j, int k, double cx, double cy,
                                                    double ct) {
  const double x = fabs(i - 0.5) * cx, y = fabs(j - 0.5) * cy, t = fabs(k - 0.5) * ct;
  const double x2 = x * x, y2 = y * y, t2 = t * t;
  const double r = sqrt(x2 + y2 + t2);
  const double inv = 1.0 / (y * t * r);
  const double a1 = atan(x * y2 * inv);  // atan(x y / (t r))
  const double a2 = atan(x * t2 * inv);  // atan(x t / (y r))
  double f = 0.5 * (x2 - t2) * a1 + 0.5 * (x2 - y2) * a2 + y * t * log(x + r) +
             x * t * log(y + r) + x * y * log(t + r);
  if (i == 0 || j == 0 || k == 0) {
    const double sx = i == 0 ? -1.0 : 1.0, sy = j == 0 ? -1.0 : 1.0, st = k == 0 ? -1.0 : 1.0;
    const double s = sx * sy * st;
    f *= s;
    if (i == 0) f += sy * st * y * t * log(y2 + t2);  // (s - sy st) d1, d1 = -1/2 y t log(y2 + t2)
    if (j == 0) f += sx * st * x * t * log(x2 + t2);
    if (k == 0) f += sx * sy * x * y * log(x2 + y2);
    f -= (s - 1.0) * (0.25 * kPi) * x2;               // (s - 1) d4, d4 = -(pi/4) x^2
  }
  return f;
}

// The reference's expression, kept for the parity tests of the new form (ch_sc_green_function
// with CH_GREEN_REFERENCE_FORM=1 in the environment).
__device__ __forceinline__ double igf_antiderivative(double x, double y, double t) {
  const double r = sqrt(x * x + y * y + t * t);
  return -0.5 * t * t * atan(x * y / (t * r)) - 0.5 * y * y * atan(x * t / (y * r)) -
         0.5 * x * x * atan(y * t / (x * r)) + y * t * asinh(x / sqrt(y * y + t * t)) +
         x * t * asinh(y / sqrt(x * x + t * t)) + x * y * asinh(t / sqrt(x * x + y * y));
}

// Far field.  Beyond a few cells the integral of 1/r over a cell is its Taylor series about the cell
// centre; odd orders vanish and, with t_a = x_a / r, e_a = h_a / r,
//   G = (V / r) [1 + 1/24 sum_a e_a^2 (3 t_a^2 - 1)
//                  + 1/1920 sum_a e_a^4 (105 t_a^4 - 90 t_a^2 + 9)
//                  + 1/576 sum_{a<b} e_a^2 e_b^2 (105 t_a^2 t_b^2 - 15 (t_a^2 + t_b^2) + 3)] + O(e^6)
// (derivatives of 1/r: d^n/dx^n = (-1)^n n! P_n(x / r) / r^(n+1)).  For r >= kFarRatio h_max the
// truncation error is <= 1.3e-7 of the value (tools/dev/igf_far_field.py, against 50-digit
// arithmetic) -- float32 resolution -- at ~60 float32 operations instead of the ~1000 fp64 ones
// of an exact antiderivative corner.  Float32 beams use it (the spectrum is float32 anyway);
// float64 beams keep the exact 8-corner difference everywhere.  With 9:1 cells (BASELINE
// configs[3]) 90 % of the lattice is far field.
constexpr double kFarRatio = 5.0;

struct GreenGeometry {
  double dx, dy, dt;  // cell sizes, d_tau scaled by gamma (space_charge_kick.py:185-189)
  double near_r2;     // points with |r|^2 below this use the exact difference (inf: all)
};

__device__ __forceinline__ GreenGeometry green_geometry(const double* prm, bool far_field) {
  GreenGeometry g;
  g.dx = prm[3];
  g.dy = prm[4];
  g.dt = prm[5] * prm[6];
  const double h = fmax(g.dx, fmax(g.dy, g.dt)) * kFarRatio;
  g.near_r2 = far_field ? h * h : CUDART_INF;
  return g;
}

// `slack` > 1 widens the near region: the lattice kernel computes slightly more points than the
// consumers ask for, so a borderline point can never be read without having been written.
__device__ __forceinline__ bool green_is_near(const GreenGeometry& g, int i, int j, int k,
                                              double slack = 1.0) {
  const double x = i * g.dx, y = j * g.dy, t = k * g.dt;
  return x * x + y * y + t * t < g.near_r2 * slack;
}

// (float32: only float32 beams take the far field, and the value is stored as float32 anyway;
// measured error of this evaluation 2-3e-7 of the value, tools/dev/igf_far_field.py)
__device__ __forceinline__ float igf_far_field(const GreenGeometry& g, int i, int j, int k) {
  const float x = static_cast<float>(i * g.dx), y = static_cast<float>(j * g.dy),
              t = static_cast<float>(k * g.dt);
  const float hx2 = static_cast<float>(g.dx * g.dx), hy2 = static_cast<float>(g.dy * g.dy),
              ht2 = static_cast<float>(g.dt * g.dt);
  const float x2 = x * x, y2 = y * y, t2 = t * t;
  const float inv_r2 = 1.0f / (x2 + y2 + t2);
  const float tx = x2 * inv_r2, ty = y2 * inv_r2, tt = t2 * inv_r2;
  const float ex = hx2 * inv_r2, ey = hy2 * inv_r2, et = ht2 * inv_r2;
  const float s2 = ex * (3.0f * tx - 1.0f) + ey * (3.0f * ty - 1.0f) + et * (3.0f * tt - 1.0f);
  const float s4a = ex * ex * ((105.0f * tx - 90.0f) * tx + 9.0f) +
                    ey * ey * ((105.0f * ty - 90.0f) * ty + 9.0f) +
                    et * et * ((105.0f * tt - 90.0f) * tt + 9.0f);
  const float s4b = ex * ey * (105.0f * tx * ty - 15.0f * (tx + ty) + 3.0f) +
                    ex * et * (105.0f * tx * tt - 15.0f * (tx + tt) + 3.0f) +
                    ey * et * (105.0f * ty * tt - 15.0f * (ty + tt) + 3.0f);
  return static_cast<float>(g.dx * g.dy * g.dt) * sqrtf(inv_r2) *
         (1.0f + s2 * (1.0f / 24.0f) + s4a * (1.0f / 1920.0f) + s4b * (1.0f / 576.0f));
}

// Green function value at grid point (i, j, k): far field, or the 8-corner signed difference of
// the antiderivative lattice (:195-236); `corner` = &lattice[i][j][k].
__device__ __forceinline__ double green_value(const GreenGeometry& g, const double* corner, int sx,
                                              int sy, int i, int j, int k) {
  if (!green_is_near(g, i, j, k)) return igf_far_field(g, i, j, k);
  return corner[sx + sy + 1] - corner[sy + 1] - corner[sx + 1] - corner[sx + sy] + corner[sx] +
         corner[sy] + corner[1] - corner[0];
}

// lattice[i][j][k] = F((i - 1/2) dx, (j - 1/2) dy, (k - 1/2) dt), 0 <= i <= nx, ...  (up to terms
// the 8-corner difference annihilates, see above)
template <bool REFERENCE_FORM>
__global__ void __launch_bounds__(256)
sc_green_lattice_kernel(const double* __restrict__ params, int nx, int ny, int nz, int far_field,
                        double* __restrict__ lattice) {
  const int64_t b = blockIdx.y;
  const double* prm = params + b * CH_SC_PARAMS;
  const GreenGeometry g = green_geometry(prm, far_field != 0);
  const double dx = g.dx, dy = g.dy, dt = g.dt;
  const double ex = nx * dx, ey = ny * dy, et = nz * dt;
  const double diagonal2 = ex * ex + ey * ey + et * et;
  const double inv_diagonal = rsqrt(diagonal2);
  const double cx = dx * inv_diagonal, cy = dy * inv_diagonal, ct = dt * inv_diagonal;
  const int64_t total = static_cast<int64_t>(nx + 1) * (ny + 1) * (nz + 1);
  double* out = lattice + b * total;
  // Threads run along y (the lattice is [i][j][k], k contiguous): the near region is a prefix
  // of every y row, so warps are either busy or skip at once.  Lattice point (i, j, k) is a corner
  // of the grid points (i-1..i, j-1..j, k-1..k); it is needed when the innermost of them is near.
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int j = static_cast<int>(idx % (ny + 1));
    const int k = static_cast<int>((idx / (ny + 1)) % (nz + 1));
    const int i = static_cast<int>(idx / (static_cast<int64_t>(ny + 1) * (nz + 1)));
    if (!green_is_near(g, max(i - 1, 0), max(j - 1, 0), max(k - 1, 0), 1.02)) continue;
    const int64_t at = (static_cast<int64_t>(i) * (ny + 1) + j) * (nz + 1) + k;
    if (REFERENCE_FORM)
      out[at] = igf_antiderivative((i - 0.5) * dx, (j - 0.5) * dy, (k - 0.5) * dt);
    else
      out[at] = diagonal2 * igf_lattice_value(i, j, k, cx, cy, ct);
  }
}

// green[2nx][2ny][2nz]: 8-corner signed difference of the lattice (:195-236), mirrored
// (:247-289); the planes with index n stay zero.
template <typename T>
__global__ void __launch_bounds__(256)
sc_green_mirror_kernel(const double* __restrict__ lattice, const double* __restrict__ params,
                       int nx, int ny, int nz, int far_field, T* __restrict__ green) {
  const int64_t b = blockIdx.y;
  const GreenGeometry g = green_geometry(params + b * CH_SC_PARAMS, far_field != 0);
  const int64_t total = static_cast<int64_t>(8) * nx * ny * nz;
  const double* f = lattice + b * static_cast<int64_t>(nx + 1) * (ny + 1) * (nz + 1);
  T* out = green + b * total;
  const int sy = nz + 1, sx = (ny + 1) * (nz + 1);
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int k2 = static_cast<int>(idx % (2 * nz));
    const int j2 = static_cast<int>((idx / (2 * nz)) % (2 * ny));
    const int i2 = static_cast<int>(idx / (static_cast<int64_t>(4) * ny * nz));
    if (i2 == nx || j2 == ny || k2 == nz) {
      out[idx] = T(0);
      continue;
    }
    const int i = i2 < nx ? i2 : 2 * nx - i2;
    const int j = j2 < ny ? j2 : 2 * ny - j2;
    const int k = k2 < nz ? k2 : 2 * nz - k2;
    out[idx] = static_cast<T>(green_value(g, f + i * sx + j * sy + k, sx, sy, i, j, k));
  }
}

// ---------------------------------------------------------------------------------------
// 5. FFT passes of the Poisson solve
// ---------------------------------------------------------------------------------------
constexpr int kFftThreads = 256;
constexpr int kRowPairs = 8;   // z pass: 8 pairs of real rows per CTA
constexpr int kColumns = 16;   // strided passes: 16 adjacent columns per CTA

// z pass, real -> complex.  Two real rows are packed as (a + i b), transformed once and
// separated with the Hermitian symmetry.  in: [B][in_x][in_y][in_z] real (in_z <= len valid
// entries, the rest of the length-`len` transform is zero padding);
// out: [B][2nx][2ny][len/2 + 1] complex, rows (x < in_x, y < in_y).
template <typename T>
__global__ void __launch_bounds__(kFftThreads)
fft_r2c_z_kernel(const T* __restrict__ in, int in_x, int in_y, int in_z, int in_pitch,
                 int quad, int len, int log2_len, int out_nx, int out_ny,
                 typename fft::Complex<T>::type* __restrict__ out) {
  using C = typename fft::Complex<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* v = reinterpret_cast<C*>(smem_raw);
  const int pitch = len + 2;  // 8 columns: see the bank note in fft.cuh
  C* tw = v + kRowPairs * pitch;
  const int64_t b = blockIdx.y;
  const int rows = in_x * in_y;
  const int kz = len / 2 + 1;
  const int first_row = blockIdx.x * 2 * kRowPairs;
  // plain rows are `in_pitch` apart; with quad != 0 the input is the quad-block charge grid of
  // sc_deposit_kernel and a value is the sum of four partial grids (quad_value)
  const int qy = in_y / 2 + 1, qz = in_z / 2 + 1;
  const int64_t plane = static_cast<int64_t>(4) * qy * qz * 4;
  const T* src = in + b * (quad ? in_x * plane : static_cast<int64_t>(rows) * in_pitch);
  C* dst = out + b * static_cast<int64_t>(out_nx) * out_ny * kz;

  fft::fill_twiddles(tw, len);
  for (int t = threadIdx.x; t < kRowPairs * len; t += kFftThreads) {
    const int pair = t / len, i = t - pair * len;
    const int r0 = first_row + 2 * pair, r1 = r0 + 1;
    C value{T(0), T(0)};
    if (i < in_z) {
      if (r0 < rows)
        value.x = quad ? quad_value(src + (r0 / in_y) * plane, r0 % in_y, i, qy, qz)
                       : src[static_cast<int64_t>(r0) * in_pitch + i];
      if (r1 < rows)
        value.y = quad ? quad_value(src + (r1 / in_y) * plane, r1 % in_y, i, qy, qz)
                       : src[static_cast<int64_t>(r1) * in_pitch + i];
    }
    v[pair * pitch + i] = value;
  }
  __syncthreads();
  fft::forward_dif<kRowPairs>(v, tw, len, log2_len, pitch);

  for (int t = threadIdx.x; t < kRowPairs * kz; t += kFftThreads) {
    const int pair = t / kz, k = t - pair * kz;
    const int r0 = first_row + 2 * pair, r1 = r0 + 1;
    if (r0 >= rows) continue;
    const C zk = v[pair * pitch + fft::bit_reverse(k, log2_len)];
    const C zm = v[pair * pitch + fft::bit_reverse((len - k) & (len - 1), log2_len)];
    // A = (Z[k] + conj Z[-k]) / 2, B = (Z[k] - conj Z[-k]) / (2i)
    const C a{T(0.5) * (zk.x + zm.x), T(0.5) * (zk.y - zm.y)};
    const C bb{T(0.5) * (zk.y + zm.y), T(-0.5) * (zk.x - zm.x)};
    const int x0 = r0 / in_y, y0 = r0 - x0 * in_y;
    dst[(static_cast<int64_t>(x0) * out_ny + y0) * kz + k] = a;
    if (r1 < rows) {
      const int x1 = r1 / in_y, y1 = r1 - x1 * in_y;
      dst[(static_cast<int64_t>(x1) * out_ny + y1) * kz + k] = bb;
    }
  }
}

// z pass, complex -> real, for rows (x < out_x, y < out_y); stores the first out_z samples of
// each length-`len` inverse transform times scale(b) = params[9] / (4 pi eps0 * 8 nx ny nz).
template <typename T>
__global__ void __launch_bounds__(kFftThreads)
fft_c2r_z_kernel(const typename fft::Complex<T>::type* __restrict__ in, int in_nx, int in_ny,
                 int len, int log2_len, int out_x, int out_y, int out_z,
                 const double* __restrict__ params, double norm, T* __restrict__ out) {
  using C = typename fft::Complex<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* v = reinterpret_cast<C*>(smem_raw);
  const int pitch = len + 2;
  C* tw = v + kRowPairs * pitch;
  const int64_t b = blockIdx.y;
  const int rows = out_x * out_y;
  const int kz = len / 2 + 1;
  const int first_row = blockIdx.x * 2 * kRowPairs;
  const C* src = in + b * static_cast<int64_t>(in_nx) * in_ny * kz;
  T* dst = out + b * static_cast<int64_t>(rows) * out_z;
  const T scale = static_cast<T>(params[b * CH_SC_PARAMS + 9] * norm);

  fft::fill_twiddles(tw, len);
  for (int t = threadIdx.x; t < kRowPairs * len; t += kFftThreads) {
    const int pair = t / len, k = t - pair * len;
    const int r0 = first_row + 2 * pair, r1 = r0 + 1;
    const int kk = k < kz ? k : len - k;  // Hermitian partner for the upper half
    C a{T(0), T(0)}, bb{T(0), T(0)};
    if (r0 < rows) {
      const int x0 = r0 / out_y, y0 = r0 - x0 * out_y;
      a = src[(static_cast<int64_t>(x0) * in_ny + y0) * kz + kk];
    }
    if (r1 < rows) {
      const int x1 = r1 / out_y, y1 = r1 - x1 * out_y;
      bb = src[(static_cast<int64_t>(x1) * in_ny + y1) * kz + kk];
    }
    if (k >= kz) {  // conj for the mirrored half
      a.y = -a.y;
      bb.y = -bb.y;
    }
    if (k == 0 || k == len / 2) a.y = bb.y = T(0);  // c2r ignores Im of DC / Nyquist
    // Z = A + i B
    v[pair * pitch + fft::bit_reverse(k, log2_len)] = C{a.x - bb.y, a.y + bb.x};
  }
  __syncthreads();
  fft::inverse_dit<kRowPairs>(v, tw, len, log2_len, pitch);

  for (int t = threadIdx.x; t < kRowPairs * out_z; t += kFftThreads) {
    const int pair = t / out_z, i = t - pair * out_z;
    const int r0 = first_row + 2 * pair, r1 = r0 + 1;
    const C z = v[pair * pitch + i];
    if (r0 < rows) dst[static_cast<int64_t>(r0) * out_z + i] = z.x * scale;
    if (r1 < rows) dst[static_cast<int64_t>(r1) * out_z + i] = z.y * scale;
  }
}

// Strided complex pass over columns of `data`.  Column (outer, inner) starts at
// outer * outer_stride + inner and steps by axis_stride; a CTA owns kColumns adjacent `inner`
// values so that every global access is a contiguous run.
//   MODE 0: forward, in place: in_len valid inputs (rest zero), `len` outputs
//   MODE 1: inverse, in place: `len` inputs, out_len outputs
//   MODE 2: forward, multiply by the Green spectrum, inverse: in_len in, out_len out.  The
//           spectrum of the mirrored (even) Green array is real and even in every index, so it
//           is stored compactly as green[B][len/2 + 1][green_ny][green_kz] (x pass only:
//           inner = ky * green_kz + kz, ky in [0, 2 (green_ny - 1)))
template <typename T, int MODE>
__global__ void __launch_bounds__(kFftThreads)
fft_strided_kernel(typename fft::Complex<T>::type* __restrict__ data,
                   const T* __restrict__ green, int green_ny, int green_kz, int len, int log2_len,
                   int in_len, int out_len, int64_t axis_stride, int inner_count,
                   int64_t outer_stride, int64_t batch_stride) {
  using C = typename fft::Complex<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* v = reinterpret_cast<C*>(smem_raw);
  const int pitch = len + 1;  // odd pitch: the 16 columns of a row fall into distinct banks
  C* tw = v + kColumns * pitch;
  const int64_t base = blockIdx.z * batch_stride + blockIdx.y * outer_stride;
  const int inner0 = blockIdx.x * kColumns;
  const int columns = min(kColumns, inner_count - inner0);
  C* col0 = data + base + inner0;

  fft::fill_twiddles(tw, len);
  for (int t = threadIdx.x; t < kColumns * len; t += kFftThreads) {
    const int i = t / kColumns, c = t - i * kColumns;
    C value{T(0), T(0)};
    if (c < columns && i < in_len) value = col0[i * axis_stride + c];
    const int slot = (MODE == 1) ? fft::bit_reverse(i, log2_len) : i;
    v[c * pitch + slot] = value;
  }
  __syncthreads();

  if (MODE == 0 || MODE == 2) fft::forward_dif<kColumns>(v, tw, len, log2_len, pitch);
  if (MODE == 2) {
    const int64_t plane = static_cast<int64_t>(green_ny) * green_kz;
    const T* g0 = green + blockIdx.z * (len / 2 + 1) * plane;
    const int full_ny = 2 * (green_ny - 1);
    for (int t = threadIdx.x; t < kColumns * len; t += kFftThreads) {
      const int k = t / kColumns, c = t - k * kColumns;
      if (c < columns) {
        const int inner = inner0 + c;
        const int ky = inner / green_kz, kz = inner - ky * green_kz;
        const int kx_even = k <= len / 2 ? k : len - k;
        const int ky_even = ky <= full_ny / 2 ? ky : full_ny - ky;
        const T g = g0[kx_even * plane + ky_even * green_kz + kz];
        C* slot = &v[c * pitch + fft::bit_reverse(k, log2_len)];
        slot->x *= g;
        slot->y *= g;
      }
    }
    __syncthreads();
  }
  if (MODE == 1 || MODE == 2) fft::inverse_dit<kColumns>(v, tw, len, log2_len, pitch);

  const int n_store = (MODE == 0) ? len : out_len;
  for (int t = threadIdx.x; t < kColumns * n_store; t += kFftThreads) {
    const int i = t / kColumns, c = t - i * kColumns;
    if (c >= columns) continue;
    const int slot = (MODE == 0) ? fft::bit_reverse(i, log2_len) : i;
    col0[i * axis_stride + c] = v[c * pitch + slot];
  }
}


// Transform of real EVEN sequences (the mirrored Green function, space_charge_kick.py:247-289):
// the column g[0..n) stands for the length-2n sequence g[0..n), 0, g[n-1..1]; its DFT is real
// and even, so only outputs k = 0..n are stored.  Two columns are packed as real and imaginary
// part of one complex transform (both spectra are real, so they separate for free).
// Column c -> (outer, inner) = divmod(c, inner_count); element i of the column lives at
//   outer * outer_stride + inner + i * axis_stride      (input and output each have their own)
// With FROM_LATTICE the input is the 8-corner difference of the antiderivative lattice
// (:195-236) evaluated on the fly (z pass: outer = x * ny + y, axis = z).
template <typename T, bool FROM_LATTICE>
__global__ void __launch_bounds__(kFftThreads)
fft_even_pass_kernel(const void* __restrict__ in, T* __restrict__ out, int n, int len,
                     int log2_len, int total_columns, int inner_count, int64_t in_outer_stride,
                     int64_t in_axis_stride, int64_t in_batch_stride, int64_t out_outer_stride,
                     int64_t out_axis_stride, int64_t out_batch_stride, int lattice_ny,
                     int lattice_nz, const double* __restrict__ params, int far_field) {
  using C = typename fft::Complex<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  C* v = reinterpret_cast<C*>(smem_raw);
  const int pitch = len + 2;  // 8 packed columns: see the bank note in fft.cuh
  constexpr int kPairs = kColumns / 2;
  C* tw = v + kPairs * pitch;
  const int c0 = blockIdx.x * kColumns;
  const int columns = min(kColumns, total_columns - c0);
  const bool axis_contiguous = in_axis_stride == 1;
  GreenGeometry geometry{};
  if (FROM_LATTICE) geometry = green_geometry(params + blockIdx.y * CH_SC_PARAMS, far_field != 0);

  fft::fill_twiddles(tw, len);
  for (int t = threadIdx.x; t < kColumns * n; t += kFftThreads) {
    // fastest index follows the contiguous direction of the input
    const int c = axis_contiguous ? t / n : t % kColumns;
    const int i = axis_contiguous ? t - c * n : t / kColumns;
    T value = T(0);
    if (c < columns) {
      const int col = c0 + c;
      const int outer = col / inner_count, inner = col - outer * inner_count;
      if (FROM_LATTICE) {
        const double* f = static_cast<const double*>(in) + blockIdx.y * in_batch_stride;
        const int x = outer / lattice_ny, y = outer - x * lattice_ny;
        const int sy = lattice_nz + 1, sx = (lattice_ny + 1) * (lattice_nz + 1);
        const double* q = f + static_cast<int64_t>(x) * sx + y * sy + i;
        value = static_cast<T>(green_value(geometry, q, sx, sy, x, y, i));
      } else {
        const T* src = static_cast<const T*>(in) + blockIdx.y * in_batch_stride;
        value = src[outer * in_outer_stride + inner + i * in_axis_stride];
      }
    }
    // even extension: position i and its mirror len - i (position n stays zero)
    C* column = v + (c >> 1) * pitch;
    if (c & 1) {
      column[i].y = value;
      if (i > 0) column[len - i].y = value;
    } else {
      column[i].x = value;
      if (i > 0) column[len - i].x = value;
    }
  }
  // positions n .. len - n are zero (position n alone when len == 2 n)
  const int gap = len - 2 * n + 1;
  for (int t = threadIdx.x; t < kPairs * gap; t += kFftThreads)
    v[(t / gap) * pitch + n + t % gap] = C{T(0), T(0)};
  __syncthreads();
  fft::forward_dif<kPairs>(v, tw, len, log2_len, pitch);

  T* dst = out + blockIdx.y * out_batch_stride;
  const int n_out = len / 2 + 1;
  for (int t = threadIdx.x; t < kColumns * n_out; t += kFftThreads) {
    const int c = axis_contiguous ? t / n_out : t % kColumns;
    const int k = axis_contiguous ? t - c * n_out : t / kColumns;
    if (c >= columns) continue;
    const int col = c0 + c;
    const int outer = col / inner_count, inner = col - outer * inner_count;
    const C z = v[(c >> 1) * pitch + fft::bit_reverse(k, log2_len)];
    dst[outer * out_outer_stride + inner + k * out_axis_stride] = (c & 1) ? z.y : z.x;
  }
}

// ---------------------------------------------------------------------------------------
// 5b. the same passes on the register-resident FFT core (fft_regs.cuh): float, lengths 32 .. 256
// ---------------------------------------------------------------------------------------
// Every pass moves its data global -> registers -> global; shared memory only carries the one
// exchange inside a transform (plus the pairing of mirrored / packed elements where a pass needs
// it).  Two thread layouts:
//   A, passes along a STRIDED axis: COLS adjacent columns per CTA, c = tid % COLS in the low bits
//      (a half-warp touches 16 adjacent columns = 128 contiguous bytes), n2 = tid / COLS;
//   B, passes along the CONTIGUOUS axis: n2 = tid % N2 in the low bits (a warp reads N2
//      consecutive elements of each of its rows), row pair = tid / N2.
// COLS / ROWS = 16 in general and 8 for launches that would not fill the GPU otherwise.
using fftr::C;

// rho[B][nx][ny][nz] = sum of the four quad-block parts of sc_deposit_kernel: one thread per
// 2 x 2 (y, z) quad reads the nine 16-byte blocks that overlap it.
__global__ void __launch_bounds__(256)
sc_quad_sum_kernel(const float* __restrict__ quad, int nx, int ny, int nz, float* __restrict__ rho) {
  const int qy = ny / 2 + 1, qz = nz / 2 + 1;
  const int hy = (ny + 1) / 2, hz = (nz + 1) / 2;
  const int64_t b = blockIdx.y;
  const int64_t quads = static_cast<int64_t>(nx) * hy * hz;
  const int64_t part = static_cast<int64_t>(qy) * qz;  // float4 blocks per part
  const float4* base = reinterpret_cast<const float4*>(quad) + b * nx * 4 * part;
  float* out = rho + b * static_cast<int64_t>(nx) * ny * nz;
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < quads;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int Z = static_cast<int>(idx % hz);
    const int Y = static_cast<int>((idx / hz) % hy);
    const int x = static_cast<int>(idx / (static_cast<int64_t>(hz) * hy));
    const float4* p00 = base + (static_cast<int64_t>(x) * 4 + 0) * part + Y * qz + Z;
    const float4* p01 = p00 + part;
    const float4* p10 = p01 + part;
    const float4* p11 = p10 + part;
    // (the +1 neighbours of the last quad of an odd-sized axis carry no charge: clamp them)
    const int nz1 = Z + 1 < qz ? 1 : 0, ny1 = Y + 1 < qy ? qz : 0;
    const float4 a = p00[0];
    const float4 b0 = p01[0], b1 = p01[nz1];
    const float4 c0 = p10[0], c1 = p10[ny1];
    const float4 d00 = p11[0], d01 = p11[nz1], d10 = p11[ny1], d11 = p11[ny1 + nz1];
    const int y0 = 2 * Y, z0 = 2 * Z;
    float* row0 = out + (static_cast<int64_t>(x) * ny + y0) * nz + z0;
    const float v00 = a.x + b0.y + c0.z + d00.w;
    const float v01 = a.y + b1.x + c0.w + d01.z;
    const float v10 = a.z + b0.w + c1.x + d10.y;
    const float v11 = a.w + b1.z + c1.y + d11.x;
    const bool z1 = z0 + 1 < nz, y1 = y0 + 1 < ny;
    row0[0] = v00;
    if (z1) row0[1] = v01;
    if (y1) {
      row0[nz] = v10;
      if (z1) row0[nz + 1] = v11;
    }
  }
}

template <int LEN, int COLS>
struct FftrSharedA {
  C exchange[COLS * fftr::Plan<LEN>::PITCH];
  C twiddles[LEN];
};
template <int LEN, int ROWS>
struct FftrSharedB {
  C exchange[ROWS * fftr::PlanB<LEN>::PITCH];
  C twiddles[LEN];
};

// Strided complex pass (MODE as in fft_strided_kernel).
template <int LEN, int MODE, int COLS>
__global__ void __launch_bounds__(COLS * fftr::Plan<LEN>::N2)
fftr_strided_kernel(C* __restrict__ data, const float* __restrict__ green, int green_ny,
                    int green_kz, int in_len, int out_len, int64_t axis_stride, int inner_count,
                    int64_t outer_stride, int64_t batch_stride) {
  constexpr int N2 = fftr::Plan<LEN>::N2, PITCH = fftr::Plan<LEN>::PITCH;
  __shared__ FftrSharedA<LEN, COLS> sh;
  const int c = threadIdx.x % COLS, n2 = threadIdx.x / COLS;
  const int inner = blockIdx.x * COLS + c;
  const bool live = inner < inner_count;
  C* col = data + blockIdx.z * batch_stride + blockIdx.y * outer_stride + inner;
  fftr::fill_twiddles<LEN>(sh.twiddles);
  C v[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    const int n = N2 * n1 + n2;
    v[n1] = (live && n < in_len) ? col[n * axis_stride] : C{0.0f, 0.0f};
  }
  __syncthreads();
  C* column = sh.exchange + c * PITCH;
  if (MODE == 0 || MODE == 2) fftr::transform<LEN, false>(v, column, n2, sh.twiddles);
  if (MODE == 2) {
    // the spectrum of the mirrored Green array is real and even in every index: compact storage
    const int64_t plane = static_cast<int64_t>(green_ny) * green_kz;
    const float* g0 = green + blockIdx.z * (LEN / 2 + 1) * plane;
    const int full_ny = 2 * (green_ny - 1);
    const int ky = live ? inner / green_kz : 0, kz = live ? inner - ky * green_kz : 0;
    const int ky_even = ky <= full_ny / 2 ? ky : full_ny - ky;
    const float* g1 = g0 + ky_even * green_kz + kz;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int kx = n2 + N2 * j;
      const int kx_even = kx <= LEN / 2 ? kx : LEN - kx;
      const float g = g1[kx_even * plane];
      v[j].x *= g;
      v[j].y *= g;
    }
    __syncthreads();  // every gather of the forward transform is done: the buffer is free
  }
  if (MODE == 1 || MODE == 2) fftr::transform<LEN, true>(v, column, n2, sh.twiddles);
  const int n_store = (MODE == 0) ? LEN : out_len;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int k = n2 + N2 * j;
    if (live && k < n_store) col[k * axis_stride] = v[j];
  }
}

// Real-even pass of the Green function along a strided axis (the y and x passes of
// green_spectrum): two adjacent real columns packed as one complex transform.  The column
// g[0..n) stands for g[0..n), 0 ..., g[n-1..1] of length LEN; outputs k = 0 .. LEN / 2 (real).
template <int LEN, int COLS>
__global__ void __launch_bounds__(COLS * fftr::Plan<LEN>::N2)
fftr_even_strided_kernel(const float* __restrict__ in, float* __restrict__ out, int n,
                         int total_columns, int inner_count, int64_t in_outer_stride,
                         int64_t in_axis_stride, int64_t in_batch_stride,
                         int64_t out_outer_stride, int64_t out_axis_stride,
                         int64_t out_batch_stride) {
  constexpr int N2 = fftr::Plan<LEN>::N2, PITCH = fftr::Plan<LEN>::PITCH;
  __shared__ FftrSharedA<LEN, COLS> sh;
  const int p = threadIdx.x % COLS, n2 = threadIdx.x / COLS;
  const int col_a = blockIdx.x * 2 * COLS + 2 * p, col_b = col_a + 1;
  const bool live_a = col_a < total_columns, live_b = col_b < total_columns;
  const int outer_a = col_a / inner_count, inner_a = col_a - outer_a * inner_count;
  const int outer_b = col_b / inner_count, inner_b = col_b - outer_b * inner_count;
  const float* src = in + blockIdx.y * in_batch_stride;
  const float* src_a = src + outer_a * in_outer_stride + inner_a;
  const float* src_b = src + outer_b * in_outer_stride + inner_b;
  fftr::fill_twiddles<LEN>(sh.twiddles);
  C v[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    const int pos = N2 * n1 + n2;
    const int i = pos < n ? pos : (pos > LEN - n ? LEN - pos : -1);  // even extension
    v[n1] = C{(live_a && i >= 0) ? src_a[i * in_axis_stride] : 0.0f,
              (live_b && i >= 0) ? src_b[i * in_axis_stride] : 0.0f};
  }
  __syncthreads();
  fftr::transform<LEN, false>(v, sh.exchange + p * PITCH, n2, sh.twiddles);
  float* dst = out + blockIdx.y * out_batch_stride;
  float* dst_a = dst + outer_a * out_outer_stride + inner_a;
  float* dst_b = dst + outer_b * out_outer_stride + inner_b;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int k = n2 + N2 * j;
    if (k <= LEN / 2) {
      if (live_a) dst_a[k * out_axis_stride] = v[j].x;
      if (live_b) dst_b[k * out_axis_stride] = v[j].y;
    }
  }
}

// Real-even z pass of the Green function straight from the antiderivative lattice: the input
// of column (x, y) is the 8-corner difference (:195-236) at z = 0 .. nz - 1, evaluated once per
// element by the thread that owns it and handed to the owner of its mirror image through
// shared memory.  out: s1[(x ny + y)][LEN / 2 + 1].
template <int LEN, int ROWS>
__global__ void __launch_bounds__(ROWS * fftr::Plan<LEN>::N2)
fftr_even_z_kernel(const double* __restrict__ lattice, const double* __restrict__ params,
                   int far_field, float* __restrict__ out, int nx, int ny, int nz) {
  constexpr int N2 = fftr::Plan<LEN>::N2, PITCH = fftr::PlanB<LEN>::PITCH, KZ = LEN / 2 + 1;
  static_assert(PITCH >= LEN / 2, "the mirror copy borrows a column of the exchange buffer");
  __shared__ FftrSharedB<LEN, ROWS> sh;
  const int n2 = threadIdx.x % N2, pr = threadIdx.x / N2;
  C* mirror = sh.exchange + pr * PITCH;  // own elements, read back by the owner of the mirror image
  const int total = nx * ny;
  const int col_a = (blockIdx.x * ROWS + pr) * 2, col_b = col_a + 1;
  const bool live_a = col_a < total, live_b = col_b < total;
  const int64_t points = static_cast<int64_t>(nx + 1) * (ny + 1) * (nz + 1);
  const double* f = lattice + blockIdx.y * points;
  const int sy = nz + 1, sx = (ny + 1) * (nz + 1);
  const GreenGeometry geometry = green_geometry(params + blockIdx.y * CH_SC_PARAMS, far_field != 0);
  const int xa = (live_a ? col_a : 0) / ny, ya = (live_a ? col_a : 0) - xa * ny;
  const int xb = (live_b ? col_b : 0) / ny, yb = (live_b ? col_b : 0) - xb * ny;
  const double* qa = f + static_cast<int64_t>(xa) * sx + ya * sy;
  const double* qb = f + static_cast<int64_t>(xb) * sx + yb * sy;
  fftr::fill_twiddles<LEN>(sh.twiddles);
  C v[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    const int pos = N2 * n1 + n2;
    v[n1] = C{0.0f, 0.0f};
    if (pos < nz) {
      v[n1] = C{live_a ? static_cast<float>(green_value(geometry, qa + pos, sx, sy, xa, ya, pos))
                       : 0.0f,
                live_b ? static_cast<float>(green_value(geometry, qb + pos, sx, sy, xb, yb, pos))
                       : 0.0f};
      mirror[pos] = v[n1];
    }
  }
  __syncthreads();
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    const int pos = N2 * n1 + n2;
    if (pos > LEN - nz) v[n1] = mirror[LEN - pos];
  }
  __syncthreads();  // the exchange of the transform overwrites the mirror copy
  fftr::transform_b<LEN, false>(v, sh.exchange + pr * PITCH, n2, sh.twiddles);
  float* dst = out + blockIdx.y * static_cast<int64_t>(total) * KZ;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int k = n2 + N2 * j;
    if (k < KZ) {
      if (live_a) dst[static_cast<int64_t>(col_a) * KZ + k] = v[j].x;
      if (live_b) dst[static_cast<int64_t>(col_b) * KZ + k] = v[j].y;
    }
  }
}

// z pass, real -> complex, on plain rows rho[rows][in_z] (sc_quad_sum_kernel): two rows packed
// per transform; out: [B][out_nx][out_ny][LEN / 2 + 1], rows (x < in_x, y < in_y).
template <int LEN, int ROWS>
__global__ void __launch_bounds__(ROWS * fftr::Plan<LEN>::N2)
fftr_r2c_z_kernel(const float* __restrict__ in, int in_x, int in_y, int in_z, int out_nx,
                  int out_ny, C* __restrict__ out) {
  constexpr int N2 = fftr::Plan<LEN>::N2, PITCH = fftr::PlanB<LEN>::PITCH, KZ = LEN / 2 + 1;
  constexpr int THREADS = ROWS * N2;
  static_assert(PITCH >= LEN, "the natural-order copy reuses a column of the exchange buffer");
  __shared__ FftrSharedB<LEN, ROWS> sh;
  const int n2 = threadIdx.x % N2, pr = threadIdx.x / N2;
  const int64_t b = blockIdx.y;
  const int rows = in_x * in_y;
  const int first_row = blockIdx.x * 2 * ROWS;
  const int r0 = first_row + 2 * pr, r1 = r0 + 1;
  const float* src = in + b * static_cast<int64_t>(rows) * in_z;
  const float* row0 = src + static_cast<int64_t>(r0) * in_z;
  const float* row1 = src + static_cast<int64_t>(r1) * in_z;
  C* dst = out + b * static_cast<int64_t>(out_nx) * out_ny * KZ;

  fftr::fill_twiddles<LEN>(sh.twiddles);
  C v[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    const int n = N2 * n1 + n2;
    v[n1] = C{(r0 < rows && n < in_z) ? row0[n] : 0.0f, (r1 < rows && n < in_z) ? row1[n] : 0.0f};
  }
  __syncthreads();
  C* column = sh.exchange + pr * PITCH;
  fftr::transform_b<LEN, false>(v, column, n2, sh.twiddles);
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 16; ++j) column[n2 + N2 * j] = v[j];
  __syncthreads();
  // separate the two real rows: A = (Z[k] + conj Z[-k]) / 2, B = (Z[k] - conj Z[-k]) / (2i)
  for (int t = threadIdx.x; t < ROWS * KZ; t += THREADS) {
    const int pair = t / KZ, k = t - pair * KZ;
    const int a0 = first_row + 2 * pair, a1 = a0 + 1;
    if (a0 >= rows) continue;
    const C zk = sh.exchange[pair * PITCH + k];
    const C zm = sh.exchange[pair * PITCH + ((LEN - k) & (LEN - 1))];
    const C a{0.5f * (zk.x + zm.x), 0.5f * (zk.y - zm.y)};
    const C bb{0.5f * (zk.y + zm.y), -0.5f * (zk.x - zm.x)};
    const int x0 = a0 / in_y, y0 = a0 - x0 * in_y;
    dst[(static_cast<int64_t>(x0) * out_ny + y0) * KZ + k] = a;
    if (a1 < rows) {
      const int x1 = a1 / in_y, y1 = a1 - x1 * in_y;
      dst[(static_cast<int64_t>(x1) * out_ny + y1) * KZ + k] = bb;
    }
  }
}

// z pass, complex -> real (semantics of fft_c2r_z_kernel): spectrum rows in, the first out_z
// samples of each inverse transform out, both straight from / to global memory.
template <int LEN, int ROWS>
__global__ void __launch_bounds__(ROWS * fftr::Plan<LEN>::N2)
fftr_c2r_z_kernel(const C* __restrict__ in, int in_nx, int in_ny, int out_x, int out_y, int out_z,
                  const double* __restrict__ params, double norm, float* __restrict__ out) {
  constexpr int N2 = fftr::Plan<LEN>::N2, PITCH = fftr::PlanB<LEN>::PITCH, KZ = LEN / 2 + 1;
  __shared__ FftrSharedB<LEN, ROWS> sh;
  const int n2 = threadIdx.x % N2, pr = threadIdx.x / N2;
  const int64_t b = blockIdx.y;
  const int rows = out_x * out_y;
  const int r0 = (blockIdx.x * ROWS + pr) * 2, r1 = r0 + 1;
  const bool live0 = r0 < rows, live1 = r1 < rows;
  const C* src = in + b * static_cast<int64_t>(in_nx) * in_ny * KZ;
  const int x0 = live0 ? r0 / out_y : 0, y0 = live0 ? r0 - x0 * out_y : 0;
  const int x1 = live1 ? r1 / out_y : 0, y1 = live1 ? r1 - x1 * out_y : 0;
  const C* row_a = src + (static_cast<int64_t>(x0) * in_ny + y0) * KZ;
  const C* row_b = src + (static_cast<int64_t>(x1) * in_ny + y1) * KZ;
  float* dst = out + b * static_cast<int64_t>(rows) * out_z;
  const float scale = static_cast<float>(params[b * CH_SC_PARAMS + 9] * norm);

  fftr::fill_twiddles<LEN>(sh.twiddles);
  C v[16];
#pragma unroll
  for (int n1 = 0; n1 < 16; ++n1) {
    const int k = N2 * n1 + n2;
    const int kk = k <= LEN / 2 ? k : LEN - k;  // Hermitian partner for the upper half
    C a = live0 ? row_a[kk] : C{0.0f, 0.0f};
    C bb = live1 ? row_b[kk] : C{0.0f, 0.0f};
    if (k > LEN / 2) {  // conj for the mirrored half
      a.y = -a.y;
      bb.y = -bb.y;
    }
    if (k == 0 || k == LEN / 2) a.y = bb.y = 0.0f;  // c2r ignores Im of DC / Nyquist
    v[n1] = C{a.x - bb.y, a.y + bb.x};             // Z = A + i B
  }
  __syncthreads();
  fftr::transform_b<LEN, true>(v, sh.exchange + pr * PITCH, n2, sh.twiddles);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int i = n2 + N2 * j;
    if (i < out_z) {
      if (live0) dst[static_cast<int64_t>(r0) * out_z + i] = v[j].x * scale;
      if (live1) dst[static_cast<int64_t>(r1) * out_z + i] = v[j].y * scale;
    }
  }
}

// ---------------------------------------------------------------------------------------
// 6. field
// ---------------------------------------------------------------------------------------
template <typename T>
struct Field4;
template <>
struct Field4<float> {
  using type = float4;
};
template <>
struct Field4<double> {
  using type = double4;
};

// One z corner pair {E(i, j, k), E(i, j, k + 1)} of the paired field layout.  For float the pair
// is one 32-byte sector and is fetched with ONE 256-bit load (LDG.E.256, sm_100+) instead of two
// 128-bit ones: half the L1 wavefronts of the gather, which is bound by L1/L2 sector traffic.
__device__ __forceinline__ void load_pair(const float4* pair, float4& lo, float4& hi) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(lo.x), "=f"(lo.y), "=f"(lo.z), "=f"(lo.w), "=f"(hi.x), "=f"(hi.y),
                 "=f"(hi.z), "=f"(hi.w)
               : "l"(pair));
}
__device__ __forceinline__ void load_pair(const double4* pair, double4& lo, double4& hi) {
  lo = pair[0];
  hi = pair[1];
}

template <typename T>
__global__ void __launch_bounds__(256)
sc_field_kernel(const T* __restrict__ phi, const double* __restrict__ params, int nx, int ny,
                int nz, typename Field4<T>::type* __restrict__ field) {
  const int64_t b = blockIdx.y;
  const double* prm = params + b * CH_SC_PARAMS;
  const int64_t total = static_cast<int64_t>(nx) * ny * nz;
  const T* f = phi + b * total;
  // reference: (phi[i+1] - phi[i-1]) * (0.5 * inv_cell), then * (-igamma2), in the beam dtype
  const T hx = T(0.5) * (T(1) / static_cast<T>(prm[3]));
  const T hy = T(0.5) * (T(1) / static_cast<T>(prm[4]));
  const T hz = T(0.5) * (T(1) / static_cast<T>(prm[5]));
  const T scale = -static_cast<T>(prm[10]);
  for (int64_t idx = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(idx % nz);
    const int j = static_cast<int>((idx / nz) % ny);
    const int i = static_cast<int>(idx / (static_cast<int64_t>(nz) * ny));
    typename Field4<T>::type e;
    e.x = (i > 0 && i < nx - 1) ? scale * ((f[idx + ny * nz] - f[idx - ny * nz]) * hx) : T(0);
    e.y = (j > 0 && j < ny - 1) ? scale * ((f[idx + nz] - f[idx - nz]) * hy) : T(0);
    e.z = (k > 0 && k < nz - 1) ? scale * ((f[idx + 1] - f[idx - 1]) * hz) : T(0);
    e.w = T(0);
    // paired layout field[cell][2] = {E(i, j, k), E(i, j, k + 1)}: the gather reads both z
    // neighbours of a corner pair with one 32-byte sector
    typename Field4<T>::type* cell = field + (b * total + idx) * 2;
    cell[0] = e;
    if (k > 0) cell[-1] = e;
    if (k == nz - 1) cell[1] = typename Field4<T>::type{T(0), T(0), T(0), T(0)};
  }
}

// ---------------------------------------------------------------------------------------
// 7. gather + kick
// ---------------------------------------------------------------------------------------
// What the FUSED instantiation does on top of the kick, while the particles are in registers:
//  * the linear section that follows the kick in the lattice (a skippable run without
//    apertures): out = M . kicked, M = the 6x7 map of a ch_compose_maps record -- saves the
//    separate ch_apply_maps pass (28 B read + 28 B written per particle);
//  * the survival-weighted sums of the NEXT SpaceChargeKick (ch_sc_moments_and_params) on the
//    outgoing coordinates, and, in the last CTA of a beam, its grid parameters -- saves the
//    moments pass (32 B read per particle) and a launch.
// Maps of the linear section fused into the float32 brick gather: copied device-to-device into
// constant memory before the launch, so that the 42 coefficients reach the FMAs through the
// uniform datapath (ULDC) instead of 12 shared-memory loads per particle -- the shared-memory /
// L1 data stage was the busiest unit of the kernel (72 %, a third of it these loads).  One
// buffer per device, owned by the first stream that uses it (gather_fused): other streams keep
// the shared-memory path, so concurrent tracks on one device cannot overwrite each other's maps.
constexpr int kConstMaps = 256;
constexpr int kConstMapPitch = 44;  // the record's header (flags, length) + 42 coefficients
__constant__ float c_gather_maps[kConstMaps * kConstMapPitch];

template <typename T>
struct GatherFusion {
  const T* records;       // null: no map
  int64_t record_stride;  // 0: one record for all beams
  int const_maps;         // the maps are in c_gather_maps (float32 brick gather)
  const T* survival;      // weights of the fused moments (null: ones)
  int64_t survival_stride;
  double* next_stats;     // null: no fused moments
  double* next_params;
  GridInputs next_in;
  int nnx, nny, nnz;
};

template <typename T, bool FUSED>
__global__ void __launch_bounds__(256, sizeof(T) == 4 ? 3 : 1)
sc_gather_kick_kernel(const T* __restrict__ particles_in, int64_t particle_stride,
                      const typename Field4<T>::type* __restrict__ field,
                      const double* __restrict__ params, int64_t n_particles, int nx, int ny,
                      int nz, int bulk_in, int bulk_out, T* __restrict__ particles_out,
                      T* __restrict__ forces_out, const GatherFusion<T> fusion) {
  constexpr int P = 4, THREADS = 256, TP = P * THREADS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);
  __shared__ uint64_t bar;
  __shared__ __align__(16) T map_s[FUSED ? 48 : 1];  // rows padded to 8 for 128-bit loads
  __shared__ double partial[FUSED ? 8 : 1][8];
  if constexpr (FUSED) {
    if (fusion.records != nullptr && threadIdx.x < 42)
      map_s[(threadIdx.x / 7) * 8 + threadIdx.x % 7] =
          fusion.records[blockIdx.y * fusion.record_stride + CH_RECORD_HEADER + threadIdx.x];
  }
  const int64_t b = blockIdx.y;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * TP;
  const int count = static_cast<int>(min(static_cast<int64_t>(TP), n_particles - n0));
  const double* prm = params + b * CH_SC_PARAMS;
  const typename Field4<T>::type* grid = field + b * static_cast<int64_t>(nx) * ny * nz * 2;

  if (bulk_in && threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  // weights of the fused moments: fetched now, their DRAM latency overlaps the tile copy
  T survival[P];
#pragma unroll
  for (int k = 0; k < P; ++k) {
    survival[k] = T(1);
    if constexpr (FUSED) {
      const int local = threadIdx.x + k * THREADS;
      if (fusion.next_stats != nullptr && fusion.survival != nullptr && local < count)
        survival[k] = fusion.survival[b * fusion.survival_stride + n0 + local];
    }
  }
  uint32_t phase = 0;
  cta_load_tile(tile, particles_in + b * particle_stride + n0 * 7, count * 7, bulk_in != 0, &bar,
                phase);

  // per-beam constants; grid geometry in the beam dtype (as the reference computes it)
  const T gd[3] = {static_cast<T>(prm[0]), static_cast<T>(prm[1]), static_cast<T>(prm[2])};
  const T cell[3] = {static_cast<T>(prm[3]), static_cast<T>(prm[4]), static_cast<T>(prm[5])};
  const int n[3] = {nx, ny, nz};
  const double gamma0 = prm[6], beta0 = prm[7], dt = prm[8];
  const double mc = prm[15] * kEvToKg * kSpeedOfLight;  // mass * c in kg m / s
  const double bg0 = gamma0 * beta0, inv_bg0 = 1.0 / bg0, dt_over_mc = dt / mc;

  T force[P][3];
  // ---- trilinear gather on the node-centred grid (space_charge_kick.py:388-475), two
  // particles at a time so that their 16 sector loads are in flight together ------------
#pragma unroll
  for (int half = 0; half < P; half += 2) {
    using F4 = typename Field4<T>::type;
    F4 lower[2][4], upper[2][4];
    T w_row[2][4], wz_lo[2], wz_hi[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int local = threadIdx.x + (half + u) * THREADS;
      const bool live = local < count;
      const T pos[3] = {live ? tile[local * 7 + 0] : T(0), live ? tile[local * 7 + 2] : T(0),
                        (live ? tile[local * 7 + 4] : T(0)) * -static_cast<T>(beta0)};
      int base[3];
      T w_lo[3], w_hi[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        const T norm = (pos[d] + gd[d]) / cell[d];
        const T fl = floor(norm);
        const T lim = static_cast<T>(n[d] + 1);
        base[d] = static_cast<int>(fmin(fmax(fl, -lim), lim));
        w_lo[d] = T(1) - fabs(norm - fl);  // 1 - |normalised - corner|  (:411-413)
        w_hi[d] = T(1) - fabs(norm - (fl + T(1)));
        // corners outside the grid contribute nothing (valid_mask, :425-433)
        if (base[d] < 0 || base[d] >= n[d]) w_lo[d] = T(0);
        if (base[d] + 1 < 0 || base[d] + 1 >= n[d]) w_hi[d] = T(0);
      }
      wz_lo[u] = w_lo[2];
      wz_hi[u] = w_hi[2];
      const bool from_first = base[2] < 0;  // base_z == -1: the upper corner is cell 0 itself
      const int cz = min(max(base[2], 0), nz - 1);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int ox = r >> 1, oy = r & 1;
        const T w = (ox ? w_hi[0] : w_lo[0]) * (oy ? w_hi[1] : w_lo[1]);
        w_row[u][r] = w * static_cast<T>(kElementaryCharge);
        const int ix = min(max(base[0] + ox, 0), nx - 1);
        const int iy = min(max(base[1] + oy, 0), ny - 1);
        const F4* pair = grid + ((static_cast<int64_t>(ix) * ny + iy) * nz + cz) * 2;
        load_pair(pair, lower[u][r], upper[u][r]);
        if (from_first) upper[u][r] = lower[u][r];
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      T fx = T(0), fy = T(0), fz = T(0);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        fx += w_row[u][r] * (wz_lo[u] * lower[u][r].x + wz_hi[u] * upper[u][r].x);
        fy += w_row[u][r] * (wz_lo[u] * lower[u][r].y + wz_hi[u] * upper[u][r].y);
        fz += w_row[u][r] * (wz_lo[u] * lower[u][r].z + wz_hi[u] * upper[u][r].z);
      }
      force[half + u][0] = fx;
      force[half + u][1] = fy;
      force[half + u][2] = fz;
    }
  }

  // fused moments of the outgoing particles: per-thread sums over its P particles in the beam
  // dtype, fp64 across threads and CTAs
  T acc8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int k = 0; k < P; ++k) {
    const int local = threadIdx.x + k * THREADS;
    T p[7];
#pragma unroll
    for (int j = 0; j < 7; ++j) p[j] = (local < count) ? tile[local * 7 + j] : T(0);
    const T fx = force[k][0], fy = force[k][1], fz = force[k][2];
    T row[7];  // outgoing particle; written back over the thread's own input row below
    if (forces_out != nullptr && local < count) {
      T* f = forces_out + (b * n_particles + n0 + local) * 3;
      f[0] = fx;
      f[1] = fy;
      f[2] = fz;
    }

    if constexpr (sizeof(T) == 4) {
      // ---- float32: the kick in difference form.  With u = P / (m c): u_x = px bg0,
      // gamma = g0 (1 + delta b0), du = F dt / (m c):
      //   px' = px + du_x / bg0,   gamma'^2 - gamma^2 = 2 u . du + |du|^2,
      //   delta' = delta + (2 u . du + |du|^2) / ((gamma' + gamma) bg0)
      // -- the algebra of ParticleBeam.to_xyz_pxpypz / from_xyz_pxpypz around P += F dt
      // (particle_beam.py:1262-1346, space_charge_kick.py:557-565) written for the CHANGE of each
      // coordinate: nothing cancels, so float32 carries the kick to ~1e-7 of itself (the
      // reference's float32 version squares SI momenta of 1e-20 kg m/s into the subnormal range,
      // SURVEY 7.3; the fp64 evaluation used before cost ~100 issue slots per particle more)
      const float bg = static_cast<float>(bg0), inv_bg = static_cast<float>(inv_bg0);
      const float du_scale = static_cast<float>(dt_over_mc);
      const float dux = fx * du_scale, duy = fy * du_scale, duz = fz * du_scale;
      const float ux = p[1] * bg, uy = p[3] * bg;
      const float gam = fmaf(p[5], bg, static_cast<float>(gamma0));
      const float uz = sqrtf(fmaxf(fmaf(gam, gam, -1.0f) - ux * ux - uy * uy, 0.0f));
      const float dg2 = 2.0f * (ux * dux + uy * duy + uz * duz) +
                        (dux * dux + duy * duy + duz * duz);
      const float gam_new = sqrtf(fmaf(gam, gam, dg2));
      row[0] = p[0];
      row[1] = fmaf(dux, inv_bg, p[1]);
      row[2] = p[2];
      row[3] = fmaf(duy, inv_bg, p[3]);
      row[4] = p[4];  // tau = -z / beta with z = -beta tau: unchanged
      row[5] = p[5] + dg2 / ((gam_new + gam) * bg);
      row[6] = p[6];
    } else {
      // ---- float64: Cheetah -> SI, kick, SI -> Cheetah (particle_beam.py:1262-1346) in units
      // of m c (u = P / (m c)): P_x = px p0, p0 / (m c) = gamma0 beta0; |u|^2 = gamma^2 - 1.
      const double gamma = gamma0 * (1.0 + static_cast<double>(p[5]) * beta0);
      double ux = static_cast<double>(p[1]) * bg0;
      double uy = static_cast<double>(p[3]) * bg0;
      double uz = sqrt(gamma * gamma - 1.0 - ux * ux - uy * uy);
      ux = fma(static_cast<double>(fx), dt_over_mc, ux);
      uy = fma(static_cast<double>(fy), dt_over_mc, uy);
      uz = fma(static_cast<double>(fz), dt_over_mc, uz);
      const double gamma_new = sqrt(1.0 + ux * ux + uy * uy + uz * uz);
      row[0] = p[0];
      row[1] = static_cast<T>(ux * inv_bg0);
      row[2] = p[2];
      row[3] = static_cast<T>(uy * inv_bg0);
      row[4] = p[4];  // tau = -z / beta with z = -beta tau: unchanged
      row[5] = static_cast<T>((gamma_new - gamma0) * inv_bg0);
      row[6] = p[6];
    }
    if constexpr (FUSED) {
      if (fusion.records != nullptr) {  // particles @ tm.mT of the following linear section
        // the gather is bound by L1 / shared-memory traffic: fetch each map row with two
        // (four for double) 128-bit loads instead of seven scalar ones
        T mapped[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          T c[8];
          if constexpr (sizeof(T) == 4) {
            const float4 lo = reinterpret_cast<const float4*>(map_s)[i * 2];
            const float4 hi = reinterpret_cast<const float4*>(map_s)[i * 2 + 1];
            c[0] = lo.x, c[1] = lo.y, c[2] = lo.z, c[3] = lo.w;
            c[4] = hi.x, c[5] = hi.y, c[6] = hi.z, c[7] = hi.w;
          } else {
#pragma unroll
            for (int h = 0; h < 4; ++h) {
              const double2 v = reinterpret_cast<const double2*>(map_s)[i * 4 + h];
              c[2 * h] = v.x, c[2 * h + 1] = v.y;
            }
          }
          T acc = c[6] * row[6];
#pragma unroll
          for (int j = 5; j >= 0; --j) acc = fma(c[j], row[j], acc);
          mapped[i] = acc;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) row[i] = mapped[i];
      }
      if (fusion.next_stats != nullptr && local < count) {
        // sums about the origin (ch_sc_beam_moments uses a pilot particle): four terms per thread
        // in the beam dtype, then fp64 -- the rounding of the partial sums averages out over the
        // ~N / 4 threads of a beam (relative error of the variance ~1e-10 (mean / sigma)^2)
        const T wi = survival[k];
        const T dx = row[0], dy = row[2], dt = row[4];
        acc8[0] += wi;
        acc8[1] = fma(wi, wi, acc8[1]);
        acc8[2] = fma(wi, dx, acc8[2]);
        acc8[3] = fma(wi, dy, acc8[3]);
        acc8[4] = fma(wi, dt, acc8[4]);
        acc8[5] = fma(wi * dx, dx, acc8[5]);
        acc8[6] = fma(wi * dy, dy, acc8[6]);
        acc8[7] = fma(wi * dt, dt, acc8[7]);
      }
    }
    // a row of the tile is only ever touched by the thread that owns the particle, so the
    // outgoing row can replace the incoming one right away (no block-wide staging of all rows)
#pragma unroll
    for (int j = 0; j < 7; ++j) tile[local * 7 + j] = row[j];
  }
  T* dst = particles_out + (b * n_particles + n0) * 7;
  if (bulk_out) {
    fence_async_shared();
    __syncthreads();
    if (threadIdx.x == 0) {
      bulk_store(dst, tile, static_cast<uint32_t>(count) * 7u * sizeof(T));
      bulk_commit();
      bulk_wait<0>();
    }
  } else {
    __syncthreads();
    for (int i = threadIdx.x; i < count * 7; i += THREADS) dst[i] = tile[i];
  }
  if constexpr (FUSED) {
    if (fusion.next_stats == nullptr) return;
    double* stats = fusion.next_stats + b * CH_SC_STATS;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const double sum = warp_sum(static_cast<double>(acc8[k]));
      if (lane == 0) partial[warp][k] = sum;
    }
    __syncthreads();
    if (threadIdx.x < 8) {
      double sum = 0.0;
      for (int wi = 0; wi < 8; ++wi) sum += partial[wi][threadIdx.x];
      atomicAdd(&stats[threadIdx.x], sum);
    }
    // With many CTAs per beam the caller derives the grid parameters in a separate tiny launch
    // (next_params == NULL here): the reductions above are fire-and-forget, whereas the
    // fence + ticket round trip below would keep every CTA's slot idle for ~2 us at its end
    // (measured: +25 % on a 64-beam gather).  With few CTAs the last one does it in place.
    if (fusion.next_params == nullptr) return;
    // last CTA of the beam: sums -> grid parameters of the next kick (as in sc_moments_kernel;
    // stats[8..10], the pilot, stay 0 from the memset)
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
      last = atomicAdd(&stats[11], 1.0) == static_cast<double>(gridDim.x - 1);
    __syncthreads();
    if (last && threadIdx.x == 0) {
      double sums[CH_SC_STATS];
      for (int i = 0; i < CH_SC_STATS; ++i) sums[i] = __ldcg(&stats[i]);
      grid_params_for_beam<T>(sums, b, fusion.next_in, fusion.nnx, fusion.nny, fusion.nnz,
                              fusion.next_params + b * CH_SC_PARAMS);
    }
  }
}

// ---------------------------------------------------------------------------------------
// 6b / 7b. float32: field bricks and the gather that reads them
// ---------------------------------------------------------------------------------------
// ncu on sc_gather_kick_kernel<float> (profiles/r02_gather_nodes_deposit_ncu_full.txt): 514
// instructions per particle at 45 % issue utilisation, the L1 data stage 71 % busy -- every
// particle pulls four 32-byte sectors out of four different 128-byte lines.  Bricks put what one
// particle needs side by side:
//   brick[b][cx][cy][cz] = float[3][8]  (96 bytes): E_component[s] at the eight nodes
//   (cx + dx, cy + dy, cz + dz), corner q = 4 dx + 2 dy + dz, zero for nodes beyond the grid:
//   3 sectors in 1.5 lines on average, three 256-bit loads, no clamping of corner indices.
// Together with (a) the per-particle survival weights fetched before the tile wait instead of
// inside the loop (one exposed DRAM latency per particle before), (b) the map of the fused
// linear section read through the uniform datapath from constant memory, (c) float32 warp sums
// of the fused moments, the fused pass went from 3.9 to 2.6 ms for 128 beams x 1e6 particles.
// Measured and rejected on the way (128 beams): four lanes per particle, one sector each, weights
// by shuffle (2.7 ms: fewer L1 wavefronts but 60 shuffles per particle); a two-stage software
// pipeline at 2 CTAs per SM (3.3 ms), and at 3 CTAs per SM with one or two sets of sector
// registers (fused 3.4-3.5 ms with the field pass against 3.26: spills; the plain kick without
// fusion gains, 3.9 -> 3.0 ms, but it only runs for the last kick of a lattice); 4 CTAs per SM at
// 64 registers (2.7-2.9 ms); the bricks of
// a thread's particles requested up front with 16-byte cp.async copies into shared-memory slots
// (512-particle tiles: 4.8 instead of 3.3 ms with the field pass); building and
// consuming the bricks in L2-sized groups of beams (1 beam per group 5.1 ms, 2: 4.4, all: 3.3
// including the field pass -- short launches pay their tails).  What bounds the kernel now
// (ncu at 128 beams, profiles/r02_sc_kernels_b128_ncu_full.txt): warps wait on their brick and
// tile loads (long scoreboard 7.3 per issue) with 24 warps per SM at 80 registers; neither fewer
// instructions (the sparse map saves 28 FMAs per particle: no change) nor L2-resident bricks
// (all beams reading one array: 3.14 instead of 3.26 ms with the field pass) move it -- it is
// the number of loads in flight per SM that a register-staged gather can keep.
constexpr int kBrickFloats = 24;
constexpr int kBrickRows = 8;

// One CTA: one x plane of cells, kBrickRows rows of y, all z.  The node fields of the
// 2 x (rows + 1) x (nz + 1) nodes it needs are computed once into shared memory (central
// differences of phi, zero on the boundary nodes and beyond the grid:
// space_charge_kick.py:324-365), then written out as bricks, consecutive lanes on consecutive
// 16-byte pieces.
__global__ void __launch_bounds__(256)
sc_field_brick_kernel(const float* __restrict__ phi, const double* __restrict__ params, int nx,
                      int ny, int nz, int beam0, float* __restrict__ bricks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* nodes = reinterpret_cast<float*>(smem_raw);  // [3][2][kBrickRows + 1][nz + 1]
  // beam blockIdx.z of this launch's group is beam beam0 + blockIdx.z of phi / params; the
  // bricks buffer holds the group only
  const int64_t b = blockIdx.z + beam0;
  const int cx = blockIdx.y, y0 = blockIdx.x * kBrickRows;
  const double* prm = params + b * CH_SC_PARAMS;
  const int64_t total = static_cast<int64_t>(nx) * ny * nz;
  const float* f = phi + b * total;
  // reference: (phi[i+1] - phi[i-1]) * (0.5 * inv_cell), then * (-igamma2), in the beam dtype
  const float hx = 0.5f * (1.0f / static_cast<float>(prm[3]));
  const float hy = 0.5f * (1.0f / static_cast<float>(prm[4]));
  const float hz = 0.5f * (1.0f / static_cast<float>(prm[5]));
  const float scale = -static_cast<float>(prm[10]);
  const int pz = nz + 1, py = kBrickRows + 1;
  const int node_count = 2 * py * pz;
  const int plane = ny * nz;
  // one warp per node row (a, j): lanes run along z, no integer divisions (the kernel is bound
  // by instruction issue, ncu: 81 %)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  for (int row = warp; row < 2 * py; row += warps) {
    const int a = row >= py ? 1 : 0, j = row - a * py;
    const int i = cx + a, jj = y0 + j;
    const bool inside_xy = i < nx && jj < ny;
    const bool x_ok = i > 0 && i < nx - 1, y_ok = jj > 0 && jj < ny - 1;
    const int base = (i * ny + jj) * nz;
    for (int k = lane; k < pz; k += 32) {
      float ex = 0.0f, ey = 0.0f, ez = 0.0f;
      if (inside_xy && k < nz) {
        const int idx = base + k;
        if (x_ok) ex = scale * ((f[idx + plane] - f[idx - plane]) * hx);
        if (y_ok) ey = scale * ((f[idx + nz] - f[idx - nz]) * hy);
        if (k > 0 && k < nz - 1) ez = scale * ((f[idx + 1] - f[idx - 1]) * hz);
      }
      const int t = row * pz + k;
      nodes[t] = ex;
      nodes[node_count + t] = ey;
      nodes[2 * node_count + t] = ez;
    }
  }
  __syncthreads();
  // item = (cz, component, half h) of one cell row: 4 corner values = one 16-byte store;
  // consecutive items are consecutive in memory.  Half h holds corners q = 4 h + (2 dy + dz).
  const int rows = min(kBrickRows, ny - y0);
  float4* out = reinterpret_cast<float4*>(
      bricks + (blockIdx.z * total + (static_cast<int64_t>(cx) * ny + y0) * nz) * kBrickFloats);
  const int row_items = nz * 6;
  for (int r = 0; r < rows; ++r) {
    for (int t = threadIdx.x; t < row_items; t += blockDim.x) {
      const int cz = t / 6, sub = t - cz * 6;
      const int h = sub & 1, comp = sub >> 1;
      const float* src = nodes + comp * node_count + (h * py + r) * pz + cz;
      out[r * row_items + t] = make_float4(src[0], src[1], src[pz], src[pz + 1]);
    }
  }
}

// One axis of the node-centred corner search (space_charge_kick.py:388-433) for the brick
// layout: the cell index whose brick holds both corners and the two weights.  The normalised
// position is (pos + half_extent) * (1 / cell) -- the reference divides; the product differs by
// an ulp at most, which can move a particle that sits on a node into the neighbouring cell, where
// the (continuous) trilinear weights give the same force to rounding.
__device__ __forceinline__ int brick_axis(float pos, float half_extent, float inv_cell, int n,
                                          float& w0, float& w1) {
  const float norm = (pos + half_extent) * inv_cell;
  const float fl = floorf(norm);
  const int base = __float2int_rd(norm);             // saturates for far-away particles
  const float w_lo = 1.0f - (norm - fl);             // 1 - |normalised - corner|  (:411-413)
  const float w_hi = 1.0f - ((fl + 1.0f) - norm);
  // corners outside the grid contribute nothing (valid_mask, :425-433)
  const bool lo_ok = static_cast<unsigned>(base) < static_cast<unsigned>(n);
  const bool hi_ok = static_cast<unsigned>(base + 1) < static_cast<unsigned>(n);
  // brick c holds nodes c and c + 1 (node n: zeros); for base == -1 node 0 is the UPPER corner:
  // brick 0 with that weight in the lower slot
  const bool shifted = base == -1;
  w0 = shifted ? w_hi : (lo_ok ? w_lo : 0.0f);
  w1 = (hi_ok && !shifted) ? w_hi : 0.0f;
  return min(max(base, 0), n - 1);
}

// 256-bit load of one brick sector, predicated.  Dead particle slots skip the load and compute on
// whatever the registers hold: their results are never stored.
__device__ __forceinline__ void load_sector(const float* src, bool active, float (&e)[8]) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.s32 p, %9, 0;\n"
      "@p ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
      "}\n"
      : "=f"(e[0]), "=f"(e[1]), "=f"(e[2]), "=f"(e[3]), "=f"(e[4]), "=f"(e[5]), "=f"(e[6]),
        "=f"(e[7])
      : "l"(src), "r"(static_cast<int>(active)));
}

// Gather + kick on bricks (see sc_gather_kick_kernel for what FUSED adds).  The kick is the
// float32 difference form of that kernel.
// P particles per thread (tiles of 256 P particles): chosen per launch, see launch_gather_bricks
template <bool FUSED, int P = 4>
__global__ void __launch_bounds__(256, 3)
sc_gather_brick_kernel(const float* __restrict__ particles_in, int64_t particle_stride,
                       const float* __restrict__ bricks, const double* __restrict__ params,
                       int64_t n_particles, int nx, int ny, int nz, int bulk_in, int bulk_out,
                       float* __restrict__ particles_out, float* __restrict__ forces_out,
                       const GatherFusion<float> fusion, int beam0) {
  constexpr int THREADS = 256, TP = P * THREADS;
  constexpr unsigned kFull = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* tile = reinterpret_cast<float*>(smem_raw);  // [TP][7]
  __shared__ uint64_t bar;
  __shared__ __align__(16) float map_s[FUSED ? 48 : 1];  // rows padded to 8 for 128-bit loads
  __shared__ double partial[FUSED ? 8 : 1][8];
  if constexpr (FUSED) {
    if (fusion.records != nullptr && !fusion.const_maps && threadIdx.x < 42)
      map_s[(threadIdx.x / 7) * 8 + threadIdx.x % 7] =
          fusion.records[(blockIdx.y + beam0) * fusion.record_stride + CH_RECORD_HEADER +
                         threadIdx.x];
  }
  // beam blockIdx.y of this launch's group is beam b of every per-beam array; the bricks buffer
  // holds the group only
  const int64_t b = blockIdx.y + beam0;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * TP;
  const int count = static_cast<int>(min(static_cast<int64_t>(TP), n_particles - n0));
  const double* prm = params + b * CH_SC_PARAMS;
  const float* grid = bricks + blockIdx.y * static_cast<int64_t>(nx) * ny * nz * kBrickFloats;
  // this beam's map in constant memory (uniform address: ULDC, no shared-memory traffic)
  const float* cmap =
      c_gather_maps + (fusion.record_stride == 0 ? 0 : static_cast<int>(b)) * kConstMapPitch;

  if (bulk_in && threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  uint32_t phase = 0;
  // per-beam constants (the loads overlap the tile copy); geometry in the beam dtype
  const float gdx = static_cast<float>(prm[0]), gdy = static_cast<float>(prm[1]),
              gdz = static_cast<float>(prm[2]);
  const float icx = static_cast<float>(prm[19]), icy = static_cast<float>(prm[20]),
              icz = static_cast<float>(prm[21]);
  const float minus_beta = -static_cast<float>(prm[7]), gamma0 = static_cast<float>(prm[6]);
  const float bg = static_cast<float>(prm[16]), inv_bg = static_cast<float>(prm[17]);
  const float du_per_field = static_cast<float>(prm[18]);
  // weights of the fused moments: fetched now, their DRAM latency overlaps the tile copy
  float survival[P];
#pragma unroll
  for (int k = 0; k < P; ++k) {
    survival[k] = 1.0f;
    if constexpr (FUSED) {
      const int local = threadIdx.x + k * THREADS;
      if (fusion.next_stats != nullptr && fusion.survival != nullptr && local < count)
        survival[k] = fusion.survival[b * fusion.survival_stride + n0 + local];
    }
  }
  cta_load_tile(tile, particles_in + b * particle_stride + n0 * 7, count * 7, bulk_in != 0, &bar,
                phase);

  const int lane = threadIdx.x & 31;
  float acc8[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // fused moments of the outgoing particles

#pragma unroll
  for (int k = 0; k < P; ++k) {
    const int local = threadIdx.x + k * THREADS;
    const bool live = local < count;
    float* mine = tile + local * 7;
    float p[7];
    p[0] = live ? mine[0] : 0.0f;
    p[2] = live ? mine[2] : 0.0f;
    p[4] = live ? mine[4] : 0.0f;
    // ---- own particle: brick index and corner weights ----------------------------------------
    float w[6];
    const int ix = brick_axis(p[0], gdx, icx, nx, w[0], w[1]);
    const int iy = brick_axis(p[2], gdy, icy, ny, w[2], w[3]);
    const int iz = brick_axis(p[4] * minus_beta, gdz, icz, nz, w[4], w[5]);
    const int cell = live ? (ix * ny + iy) * nz + iz : -1;
    float fx, fy, fz;
    {
      // one thread per particle: its three sectors with three 256-bit loads
      float e[3][8];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        load_sector(grid + (static_cast<unsigned>(max(cell, 0)) * kBrickFloats + c * 8), live,
                    e[c]);
      p[1] = live ? mine[1] : 0.0f;
      p[3] = live ? mine[3] : 0.0f;
      p[5] = live ? mine[5] : 0.0f;
      p[6] = live ? mine[6] : 0.0f;
      const float xy00 = w[0] * w[2], xy01 = w[0] * w[3], xy10 = w[1] * w[2], xy11 = w[1] * w[3];
      float f[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float acc = xy00 * fmaf(w[5], e[c][1], w[4] * e[c][0]);
        acc = fmaf(xy01, fmaf(w[5], e[c][3], w[4] * e[c][2]), acc);
        acc = fmaf(xy10, fmaf(w[5], e[c][5], w[4] * e[c][4]), acc);
        acc = fmaf(xy11, fmaf(w[5], e[c][7], w[4] * e[c][6]), acc);
        f[c] = live ? acc : 0.0f;
      }
      fx = f[0];
      fy = f[1];
      fz = f[2];
    }
    // ---- own particle: kick (difference form, see sc_gather_kick_kernel), map, moments -------
    if (forces_out != nullptr && live) {
      float* f = forces_out + (b * n_particles + n0 + local) * 3;
      f[0] = fx * static_cast<float>(kElementaryCharge);
      f[1] = fy * static_cast<float>(kElementaryCharge);
      f[2] = fz * static_cast<float>(kElementaryCharge);
    }
    const float dux = fx * du_per_field, duy = fy * du_per_field, duz = fz * du_per_field;
    const float ux = p[1] * bg, uy = p[3] * bg;
    const float gam = fmaf(p[5], bg, gamma0);  // g0 (1 + delta b0)
    // square roots and the division through the special-function unit (2 ulp): u_z and
    // gamma' + gamma only scale the small change of delta
    const float uz2 = fmaxf(fmaf(gam, gam, -1.0f) - ux * ux - uy * uy, 1e-30f);
    const float uz = uz2 * rsqrtf(uz2);
    const float dg2 =
        2.0f * (ux * dux + uy * duy + uz * duz) + (dux * dux + duy * duy + duz * duz);
    const float g2 = fmaxf(fmaf(gam, gam, dg2), 1e-30f);
    const float gam_new = g2 * rsqrtf(g2);
    float row[7];
    row[0] = p[0];
    row[1] = fmaf(dux, inv_bg, p[1]);
    row[2] = p[2];
    row[3] = fmaf(duy, inv_bg, p[3]);
    row[4] = p[4];  // tau = -z / beta with z = -beta tau: unchanged
    row[5] = p[5] + __fdividef(dg2, (gam_new + gam) * bg);
    row[6] = p[6];
    bool rewrite_positions = false;
    if constexpr (FUSED) {
      if (fusion.records != nullptr) {  // particles @ tm.mT of the following linear section
        float mapped[6];
        if (fusion.const_maps) {
          const float* m = cmap + CH_RECORD_HEADER;
          // the record's sparsity flags (ch_compose_maps; uniform over the CTA): an uncoupled
          // section without tau dependence -- drifts, upright quadrupoles, correctors -- needs 14
          // of the 42 multiply-adds (the chains of apply_maps_kernel's sparse branch)
          constexpr uint32_t kSparse = CH_FLAG_XY_UNCOUPLED | CH_FLAG_NO_TAU_COLUMN |
                                       CH_FLAG_NO_Y_DISPERSION | CH_FLAG_DELTA_IDENTITY;
          if ((__float_as_uint(cmap[0]) & kSparse) == kSparse) {
            const float w = row[6];
            mapped[0] = fmaf(m[0], row[0], fmaf(m[1], row[1], fmaf(m[5], row[5], m[6] * w)));
            mapped[1] = fmaf(m[7], row[0], fmaf(m[8], row[1], fmaf(m[12], row[5], m[13] * w)));
            mapped[2] = fmaf(m[16], row[2], fmaf(m[17], row[3], m[20] * w));
            mapped[3] = fmaf(m[23], row[2], fmaf(m[24], row[3], m[27] * w));
            mapped[4] = fmaf(m[28], row[0],
                             fmaf(m[29], row[1],
                                  fmaf(m[32], row[4], fmaf(m[33], row[5], m[34] * w))));
            mapped[5] = row[5];
          } else {
#pragma unroll
            for (int i = 0; i < 6; ++i) {
              float a = m[i * 7 + 6] * row[6];
#pragma unroll
              for (int j = 5; j >= 0; --j) a = fmaf(m[i * 7 + j], row[j], a);
              mapped[i] = a;
            }
          }
        } else {
#pragma unroll
          for (int i = 0; i < 6; ++i) {
            const float4 lo = reinterpret_cast<const float4*>(map_s)[i * 2];
            const float4 hi = reinterpret_cast<const float4*>(map_s)[i * 2 + 1];
            float a = hi.z * row[6];
            a = fmaf(hi.y, row[5], a);
            a = fmaf(hi.x, row[4], a);
            a = fmaf(lo.w, row[3], a);
            a = fmaf(lo.z, row[2], a);
            a = fmaf(lo.y, row[1], a);
            a = fmaf(lo.x, row[0], a);
            mapped[i] = a;
          }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) row[i] = mapped[i];
        rewrite_positions = true;
      }
      if (fusion.next_stats != nullptr && live) {
        const float wi = survival[k];
        const float dx = row[0], dy = row[2], dt = row[4];
        acc8[0] += wi;
        acc8[1] = fmaf(wi, wi, acc8[1]);
        acc8[2] = fmaf(wi, dx, acc8[2]);
        acc8[3] = fmaf(wi, dy, acc8[3]);
        acc8[4] = fmaf(wi, dt, acc8[4]);
        acc8[5] = fmaf(wi * dx, dx, acc8[5]);
        acc8[6] = fmaf(wi * dy, dy, acc8[6]);
        acc8[7] = fmaf(wi * dt, dt, acc8[7]);
      }
    }
    // a row of the tile is only ever touched by the thread that owns the particle; without a
    // map only the momenta change
    if (live) {
      mine[1] = row[1];
      mine[3] = row[3];
      mine[5] = row[5];
      if (rewrite_positions) {
        mine[0] = row[0];
        mine[2] = row[2];
        mine[4] = row[4];
      }
    }
  }
  float* dst = particles_out + (b * n_particles + n0) * 7;
  if (bulk_out) {
    fence_async_shared();
    __syncthreads();
    if (threadIdx.x == 0) {
      bulk_store(dst, tile, static_cast<uint32_t>(count) * 7u * sizeof(float));
      bulk_commit();
      bulk_wait<0>();
    }
  } else {
    __syncthreads();
    for (int i = threadIdx.x; i < count * 7; i += THREADS) dst[i] = tile[i];
  }
  if constexpr (FUSED) {
    if (fusion.next_stats == nullptr) return;
    double* stats = fusion.next_stats + b * CH_SC_STATS;
    // 128 terms per warp in float32 (their rounding averages out over the ~N / 128 warps of a
    // beam like that of the per-thread sums), float64 across warps and CTAs
    const int warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float sum = acc8[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(kFull, sum, o);
      if (lane == 0) partial[warp][k] = static_cast<double>(sum);
    }
    __syncthreads();
    if (threadIdx.x < 8) {
      double sum = 0.0;
      for (int wi = 0; wi < 8; ++wi) sum += partial[wi][threadIdx.x];
      atomicAdd(&stats[threadIdx.x], sum);
    }
    if (fusion.next_params == nullptr) return;  // see sc_gather_kick_kernel
    __shared__ bool last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0)
      last = atomicAdd(&stats[11], 1.0) == static_cast<double>(gridDim.x - 1);
    __syncthreads();
    if (last && threadIdx.x == 0) {
      double sums[CH_SC_STATS];
      for (int i = 0; i < CH_SC_STATS; ++i) sums[i] = __ldcg(&stats[i]);
      grid_params_for_beam<float>(sums, b, fusion.next_in, fusion.nnx, fusion.nny, fusion.nnz,
                                  fusion.next_params + b * CH_SC_PARAMS);
    }
  }
}

// ---------------------------------------------------------------------------------------
// host-side helpers
// ---------------------------------------------------------------------------------------
int log2_exact(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return ((1 << l) == v) ? l : -1;
}

// Any grid size in [2, 256] per axis.  The convolution with the Green function is aperiodic, so
// the padded transform length only has to be >= 2 n - 1: the next power of two >= 2 n (>= 8),
// which is 2 n itself for the power-of-two grids of the benchmarks.
bool grid_ok(int nx, int ny, int nz) {
  for (int v : {nx, ny, nz})
    if (v < 2 || v > 256) return false;
  return true;
}

int fft_len(int n) {
  int len = 8;
  while (len < 2 * n) len *= 2;
  return len;
}

template <typename K>
int allow_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024)
    CH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(bytes)));
  return CH_OK;
}

unsigned blocks_for(int64_t work, int threads, int64_t cap = 148 * 16) {
  const int64_t blocks = (work + threads - 1) / threads;
  return static_cast<unsigned>(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

// Compact Green spectrum [B][nx+1][ny+1][nz+1] from the antiderivative lattice: three even
// passes through two scratch arrays s1 [B][nx][ny][nz+1], s2 [B][nx][ny+1][nz+1].
// LEN -> std::integral_constant for the register-FFT kernels; false when LEN is not covered
template <typename F>
bool fftr_dispatch(int len, F&& f) {
  switch (len) {
    case 32: f(std::integral_constant<int, 32>{}); return true;
    case 64: f(std::integral_constant<int, 64>{}); return true;
    case 128: f(std::integral_constant<int, 128>{}); return true;
    case 256: f(std::integral_constant<int, 256>{}); return true;
    default: return false;
  }
}
// 16 columns (row pairs) per CTA, 8 when the launch would not give every SM two CTAs otherwise
template <typename F>
void fftr_tile(int64_t units, F&& f) {
  if ((units + 15) / 16 >= 2 * 148)
    f(std::integral_constant<int, 16>{});
  else
    f(std::integral_constant<int, 8>{});
}
// Test knobs: CH_GREEN_REFERENCE_FORM=1 evaluates the reference's antiderivative expression,
// CH_GREEN_NO_FAR_FIELD=1 the exact 8-corner difference everywhere.
bool green_reference_form() {
  static const bool on = [] {
    const char* v = getenv("CH_GREEN_REFERENCE_FORM");
    return v != nullptr && atoi(v) != 0;
  }();
  return on;
}
int green_far_field(int dtype) {
  static const bool off = [] {
    const char* v = getenv("CH_GREEN_NO_FAR_FIELD");
    return v != nullptr && atoi(v) != 0;
  }();
  return (dtype == CH_F32 && !off && !green_reference_form()) ? 1 : 0;
}
bool fftr_covers(int len) { return len == 32 || len == 64 || len == 128 || len == 256; }
bool fftr_enabled() {
  static const bool on = [] {
    const char* v = getenv("CH_FFT_SHARED_ONLY");  // test knob: force the fft.cuh passes
    return !(v != nullptr && atoi(v) != 0);
  }();
  return on;
}

// Compact Green spectrum [B][Kx][Ky][Kz] (K = fft_len / 2 + 1) from the antiderivative lattice:
// three even passes through two scratch arrays s1 [B][nx][ny][Kz], s2 [B][nx][Ky][Kz].
template <typename T>
int green_spectrum(const double* lattice, const double* params, int far_field, int64_t B, int nx,
                   int ny, int nz, T* s1, T* s2, T* spectrum, cudaStream_t stream) {
  using C = typename fft::Complex<T>::type;
  const int Lx = fft_len(nx), Ly = fft_len(ny), Lz = fft_len(nz);
  const int Kx = Lx / 2 + 1, Ky = Ly / 2 + 1, Kz = Lz / 2 + 1;
  const unsigned nb = static_cast<unsigned>(B);
  auto smem = [&](int len) { return sizeof(C) * ((kColumns / 2) * (len + 2) + len / 2); };
  const int64_t lattice_points = static_cast<int64_t>(nx + 1) * (ny + 1) * (nz + 1);
  constexpr bool kFloat = std::is_same<T, float>::value;
  const bool regs = kFloat && fftr_enabled() && fftr_covers(Lx) && fftr_covers(Ly) &&
                    fftr_covers(Lz);
  {  // z: rows (x, y) of the lattice difference -> s1[x][y][kz]
    const int columns = nx * ny;
    if (regs) {
      if constexpr (kFloat) {
        fftr_dispatch(Lz, [&](auto L) {
          constexpr int LEN = decltype(L)::value;
          fftr_tile((static_cast<int64_t>(columns) + 1) / 2 * B, [&](auto R) {
            constexpr int ROWS = decltype(R)::value;
            dim3 grid((columns + 2 * ROWS - 1) / (2 * ROWS), nb);
            fftr_even_z_kernel<LEN, ROWS><<<grid, ROWS * fftr::Plan<LEN>::N2, 0, stream>>>(
                lattice, params, far_field, s1, nx, ny, nz);
          });
        });
      }
    } else {
      auto k = fft_even_pass_kernel<T, true>;
      if (allow_smem(k, smem(Lz)) != CH_OK) return CH_ECUDA;
      dim3 grid((columns + kColumns - 1) / kColumns, nb);
      k<<<grid, kFftThreads, smem(Lz), stream>>>(
          lattice, s1, nz, Lz, log2_exact(Lz), columns, 1, 0, 1, lattice_points, Kz, 1,
          static_cast<int64_t>(nx) * ny * Kz, ny, nz, params, far_field);
    }
    CH_LAUNCH_CHECK();
  }
  {  // y: columns (x, kz) -> s2[x][ky][kz]
    const int columns = nx * Kz;
    if (regs) {
      if constexpr (kFloat) {
        fftr_dispatch(Ly, [&](auto L) {
          constexpr int LEN = decltype(L)::value;
          fftr_tile((static_cast<int64_t>(columns) + 1) / 2 * B, [&](auto R) {
            constexpr int COLS = decltype(R)::value;
            dim3 grid((columns + 2 * COLS - 1) / (2 * COLS), nb);
            fftr_even_strided_kernel<LEN, COLS><<<grid, COLS * fftr::Plan<LEN>::N2, 0, stream>>>(
                s1, s2, ny, columns, Kz, static_cast<int64_t>(ny) * Kz, Kz,
                static_cast<int64_t>(nx) * ny * Kz, static_cast<int64_t>(Ky) * Kz, Kz,
                static_cast<int64_t>(nx) * Ky * Kz);
          });
        });
      }
    } else {
      auto k = fft_even_pass_kernel<T, false>;
      if (allow_smem(k, smem(Ly)) != CH_OK) return CH_ECUDA;
      dim3 grid((columns + kColumns - 1) / kColumns, nb);
      k<<<grid, kFftThreads, smem(Ly), stream>>>(
          s1, s2, ny, Ly, log2_exact(Ly), columns, Kz, static_cast<int64_t>(ny) * Kz, Kz,
          static_cast<int64_t>(nx) * ny * Kz, static_cast<int64_t>(Ky) * Kz, Kz,
          static_cast<int64_t>(nx) * Ky * Kz, 0, 0, nullptr, 0);
    }
    CH_LAUNCH_CHECK();
  }
  {  // x: columns (ky, kz) -> spectrum[kx][ky][kz]
    const int columns = Ky * Kz;
    if (regs) {
      if constexpr (kFloat) {
        fftr_dispatch(Lx, [&](auto L) {
          constexpr int LEN = decltype(L)::value;
          fftr_tile((static_cast<int64_t>(columns) + 1) / 2 * B, [&](auto R) {
            constexpr int COLS = decltype(R)::value;
            dim3 grid((columns + 2 * COLS - 1) / (2 * COLS), nb);
            fftr_even_strided_kernel<LEN, COLS><<<grid, COLS * fftr::Plan<LEN>::N2, 0, stream>>>(
                s2, spectrum, nx, columns, columns, 0, columns,
                static_cast<int64_t>(nx) * columns, 0, columns,
                static_cast<int64_t>(Kx) * columns);
          });
        });
      }
    } else {
      auto k = fft_even_pass_kernel<T, false>;
      if (allow_smem(k, smem(Lx)) != CH_OK) return CH_ECUDA;
      dim3 grid((columns + kColumns - 1) / kColumns, nb);
      k<<<grid, kFftThreads, smem(Lx), stream>>>(
          s2, spectrum, nx, Lx, log2_exact(Lx), columns, columns, 0, columns,
          static_cast<int64_t>(nx) * columns, 0, columns, static_cast<int64_t>(Kx) * columns, 0,
          0, nullptr, 0);
    }
    CH_LAUNCH_CHECK();
  }
  return CH_OK;
}

// `green_ready` (optional): event after which green_spectrum_compact is valid when it is computed
// on another stream; only the x convolution waits for it, the charge's z and y passes do not.
template <typename T>
int poisson_solve(const T* rho, const T* green_spectrum_compact, const double* params, int64_t B,
                  int nx, int ny, int nz, typename fft::Complex<T>::type* rs, T* phi,
                  cudaStream_t stream, cudaEvent_t green_ready = nullptr) {
  using C = typename fft::Complex<T>::type;
  const int Nx = fft_len(nx), Ny = fft_len(ny), Nz = fft_len(nz), Kz = Nz / 2 + 1;
  const int lx = log2_exact(Nx), ly = log2_exact(Ny), lz = log2_exact(Nz);
  const int64_t spectrum = static_cast<int64_t>(Nx) * Ny * Kz;
  auto z_smem = [&](int len) { return sizeof(C) * (kRowPairs * (len + 2) + len / 2); };
  auto s_smem = [&](int len) { return sizeof(C) * (kColumns * (len + 1) + len / 2); };
  const unsigned nb = static_cast<unsigned>(B);
  const double norm = 1.0 / (4.0 * kPi * kEpsilon0) / (static_cast<double>(Nx) * Ny * Nz);
  constexpr bool kFloat = std::is_same<T, float>::value;
  if constexpr (kFloat) {
    if (fftr_enabled() && fftr_covers(Nx) && fftr_covers(Ny) && fftr_covers(Nz)) {
      // ---- register-FFT passes: rho z, y; fused x convolution; inverse y, z ----------------
      // plain charge rows first (phi is free until the last pass writes it)
      {
        const int64_t quads = static_cast<int64_t>(nx) * ((ny + 1) / 2) * ((nz + 1) / 2);
        dim3 grid(blocks_for(quads, 256), nb);
        sc_quad_sum_kernel<<<grid, 256, 0, stream>>>(rho, nx, ny, nz, phi);
        CH_LAUNCH_CHECK();
      }
      const int64_t row_pairs = (static_cast<int64_t>(nx) * ny + 1) / 2 * B;
      fftr_dispatch(Nz, [&](auto L) {
        constexpr int LEN = decltype(L)::value;
        fftr_tile(row_pairs, [&](auto R) {
          constexpr int ROWS = decltype(R)::value;
          dim3 grid((nx * ny + 2 * ROWS - 1) / (2 * ROWS), nb);
          fftr_r2c_z_kernel<LEN, ROWS><<<grid, ROWS * fftr::Plan<LEN>::N2, 0, stream>>>(
              phi, nx, ny, nz, Nx, Ny, rs);
        });
      });
      CH_LAUNCH_CHECK();
      fftr_dispatch(Ny, [&](auto L) {
        constexpr int LEN = decltype(L)::value;
        fftr_tile(static_cast<int64_t>(Kz) * nx * B, [&](auto R) {
          constexpr int COLS = decltype(R)::value;
          dim3 grid((Kz + COLS - 1) / COLS, nx, nb);
          fftr_strided_kernel<LEN, 0, COLS><<<grid, COLS * fftr::Plan<LEN>::N2, 0, stream>>>(
              rs, nullptr, 0, 0, ny, Ny, Kz, Kz, static_cast<int64_t>(Ny) * Kz, spectrum);
        });
      });
      CH_LAUNCH_CHECK();
      if (green_ready != nullptr) CH_CUDA(cudaStreamWaitEvent(stream, green_ready, 0));
      fftr_dispatch(Nx, [&](auto L) {
        constexpr int LEN = decltype(L)::value;
        const int inner = Ny * Kz;
        fftr_tile(static_cast<int64_t>(inner) * B, [&](auto R) {
          constexpr int COLS = decltype(R)::value;
          dim3 grid((inner + COLS - 1) / COLS, 1, nb);
          fftr_strided_kernel<LEN, 2, COLS><<<grid, COLS * fftr::Plan<LEN>::N2, 0, stream>>>(
              rs, green_spectrum_compact, Ny / 2 + 1, Kz, nx, nx, inner, inner, 0, spectrum);
        });
      });
      CH_LAUNCH_CHECK();
      fftr_dispatch(Ny, [&](auto L) {
        constexpr int LEN = decltype(L)::value;
        fftr_tile(static_cast<int64_t>(Kz) * nx * B, [&](auto R) {
          constexpr int COLS = decltype(R)::value;
          dim3 grid((Kz + COLS - 1) / COLS, nx, nb);
          fftr_strided_kernel<LEN, 1, COLS><<<grid, COLS * fftr::Plan<LEN>::N2, 0, stream>>>(
              rs, nullptr, 0, 0, Ny, ny, Kz, Kz, static_cast<int64_t>(Ny) * Kz, spectrum);
        });
      });
      CH_LAUNCH_CHECK();
      fftr_dispatch(Nz, [&](auto L) {
        constexpr int LEN = decltype(L)::value;
        fftr_tile(row_pairs, [&](auto R) {
          constexpr int ROWS = decltype(R)::value;
          dim3 grid((nx * ny + 2 * ROWS - 1) / (2 * ROWS), nb);
          fftr_c2r_z_kernel<LEN, ROWS><<<grid, ROWS * fftr::Plan<LEN>::N2, 0, stream>>>(
              rs, Nx, Ny, nx, ny, nz, params, norm, phi);
        });
      });
      CH_LAUNCH_CHECK();
      return CH_OK;
    }
  }

  // ---- rho: z (zero-padded rows of the physical octant), then y ------------------------
  {
    auto k = fft_r2c_z_kernel<T>;
    if (allow_smem(k, z_smem(Nz)) != CH_OK) return CH_ECUDA;
    dim3 grid((nx * ny + 2 * kRowPairs - 1) / (2 * kRowPairs), nb);
    k<<<grid, kFftThreads, z_smem(Nz), stream>>>(rho, nx, ny, nz, 0, 1, Nz, lz, Nx, Ny, rs);
    CH_LAUNCH_CHECK();
  }
  {
    auto k = fft_strided_kernel<T, 0>;
    if (allow_smem(k, s_smem(Ny)) != CH_OK) return CH_ECUDA;
    dim3 grid((Kz + kColumns - 1) / kColumns, nx, nb);
    k<<<grid, kFftThreads, s_smem(Ny), stream>>>(rs, nullptr, 0, 0, Ny, ly, ny, Ny, Kz, Kz,
                                                  static_cast<int64_t>(Ny) * Kz, spectrum);
    CH_LAUNCH_CHECK();
  }
  // ---- x: forward . multiply by the Green spectrum . inverse, fused; only x < nx stored ----
  if (green_ready != nullptr) CH_CUDA(cudaStreamWaitEvent(stream, green_ready, 0));
  {
    auto k = fft_strided_kernel<T, 2>;
    if (allow_smem(k, s_smem(Nx)) != CH_OK) return CH_ECUDA;
    const int inner = Ny * Kz;
    dim3 grid((inner + kColumns - 1) / kColumns, 1, nb);
    k<<<grid, kFftThreads, s_smem(Nx), stream>>>(rs, green_spectrum_compact, Ny / 2 + 1, Kz, Nx,
                                                  lx,
                                                  nx, nx, inner, inner, 0, spectrum);
    CH_LAUNCH_CHECK();
  }
  // ---- inverse y (keep y < ny) and inverse z (keep z < nz) -------------------------------
  {
    auto k = fft_strided_kernel<T, 1>;
    if (allow_smem(k, s_smem(Ny)) != CH_OK) return CH_ECUDA;
    dim3 grid((Kz + kColumns - 1) / kColumns, nx, nb);
    k<<<grid, kFftThreads, s_smem(Ny), stream>>>(rs, nullptr, 0, 0, Ny, ly, Ny, ny, Kz, Kz,
                                                  static_cast<int64_t>(Ny) * Kz, spectrum);
    CH_LAUNCH_CHECK();
  }
  {
    auto k = fft_c2r_z_kernel<T>;
    if (allow_smem(k, z_smem(Nz)) != CH_OK) return CH_ECUDA;
    dim3 grid((nx * ny + 2 * kRowPairs - 1) / (2 * kRowPairs), nb);
    k<<<grid, kFftThreads, z_smem(Nz), stream>>>(rs, Nx, Ny, Nz, lz, nx, ny, nz, params, norm, phi);
    CH_LAUNCH_CHECK();
  }
  return CH_OK;
}

}  // namespace
}  // namespace ch

// =========================================================================================
// C ABI
// =========================================================================================
#define CH_SC_COMMON_CHECKS(fn)                                                          \
  CH_REQUIRE(dtype == CH_F32 || dtype == CH_F64, fn ": bad dtype %d", dtype);            \
  CH_REQUIRE(n_beams > 0 && n_beams <= 65535, fn ": n_beams must be in [1, 65535]")

namespace {
int launch_moments(const void* particles, int64_t particle_stride, const void* survival,
                   int64_t survival_stride, int64_t n_particles, int64_t n_beams, int32_t dtype,
                   double* stats, const ch::GridInputs& inputs, int nx, int ny, int nz,
                   double* params, void* stream, double* partials = nullptr);
}

extern "C" int ch_sc_beam_moments(const void* particles, int64_t particle_stride,
                                  const void* survival, int64_t survival_stride,
                                  int64_t n_particles, int64_t n_beams, int32_t dtype,
                                  double* stats, void* stream) {
  return launch_moments(particles, particle_stride, survival, survival_stride, n_particles, n_beams,
                        dtype, stats, ch::GridInputs{}, 0, 0, 0, nullptr, stream);
}

extern "C" int ch_sc_moments_and_params(
    const void* particles, int64_t particle_stride, const void* survival, int64_t survival_stride,
    int64_t n_particles, int64_t n_beams, const void* energy, int64_t energy_stride,
    int32_t energy_dtype, const void* mass_eV, int32_t mass_dtype, const void* effect_length,
    int64_t length_stride, int32_t length_dtype, const void* extent_x, int64_t extent_x_stride,
    const void* extent_y, int64_t extent_y_stride, const void* extent_tau,
    int64_t extent_tau_stride, int32_t extent_dtype, int32_t nx, int32_t ny, int32_t nz,
    int32_t dtype, double* stats, double* params, void* stream) {
  CH_REQUIRE(energy && mass_eV && effect_length && extent_x && extent_y && extent_tau && params,
             "ch_sc_moments_and_params: NULL pointer argument");
  CH_REQUIRE(ch::grid_ok(nx, ny, nz),
             "ch_sc_moments_and_params: grid sizes (%d, %d, %d) must be in [2, 256]", nx,
             ny, nz);
  ch::GridInputs in{{energy, energy_stride, energy_dtype},
                    {mass_eV, 0, mass_dtype},
                    {effect_length, length_stride, length_dtype},
                    {extent_x, extent_x_stride, extent_dtype},
                    {extent_y, extent_y_stride, extent_dtype},
                    {extent_tau, extent_tau_stride, extent_dtype}};
  return launch_moments(particles, particle_stride, survival, survival_stride, n_particles, n_beams,
                        dtype, stats, in, nx, ny, nz, params, stream);
}

extern "C" int ch_sc_moments_and_params_deterministic(
    const void* particles, int64_t particle_stride, const void* survival, int64_t survival_stride,
    int64_t n_particles, int64_t n_beams, const void* energy, int64_t energy_stride,
    int32_t energy_dtype, const void* mass_eV, int32_t mass_dtype, const void* effect_length,
    int64_t length_stride, int32_t length_dtype, const void* extent_x, int64_t extent_x_stride,
    const void* extent_y, int64_t extent_y_stride, const void* extent_tau,
    int64_t extent_tau_stride, int32_t extent_dtype, int32_t nx, int32_t ny, int32_t nz,
    int32_t dtype, double* partials, double* stats, double* params, void* stream) {
  CH_REQUIRE(energy && mass_eV && effect_length && extent_x && extent_y && extent_tau && params &&
                 partials,
             "ch_sc_moments_and_params_deterministic: NULL pointer argument");
  CH_REQUIRE(ch::grid_ok(nx, ny, nz),
             "ch_sc_moments_and_params_deterministic: grid sizes (%d, %d, %d) must be in [2, 256]",
             nx, ny, nz);
  ch::GridInputs in{{energy, energy_stride, energy_dtype},
                    {mass_eV, 0, mass_dtype},
                    {effect_length, length_stride, length_dtype},
                    {extent_x, extent_x_stride, extent_dtype},
                    {extent_y, extent_y_stride, extent_dtype},
                    {extent_tau, extent_tau_stride, extent_dtype}};
  return launch_moments(particles, particle_stride, survival, survival_stride, n_particles, n_beams,
                        dtype, stats, in, nx, ny, nz, params, stream, partials);
}

namespace {
int launch_moments(const void* particles, int64_t particle_stride, const void* survival,
                   int64_t survival_stride, int64_t n_particles, int64_t n_beams, int32_t dtype,
                   double* stats, const ch::GridInputs& inputs, int nx, int ny, int nz,
                   double* params, void* stream, double* partials) {
  CH_SC_COMMON_CHECKS("ch_sc_beam_moments");
  CH_REQUIRE(particles && stats && n_particles > 0, "ch_sc_beam_moments: bad arguments");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  CH_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * CH_SC_STATS * n_beams, s));
  dim3 grid(static_cast<unsigned>((n_particles + 1023) / 1024), static_cast<unsigned>(n_beams));
  if (dtype == CH_F32) {
    const int bulk = ch::bulk_compatible<float>(particles, n_particles, particle_stride);
    ch::sc_moments_kernel<float><<<grid, 256, 1024 * 7 * sizeof(float), s>>>(
        static_cast<const float*>(particles), particle_stride, static_cast<const float*>(survival),
        survival_stride, n_particles, bulk, stats, inputs, nx, ny, nz, params, partials);
  } else {
    const int bulk = ch::bulk_compatible<double>(particles, n_particles, particle_stride);
    auto kernel = ch::sc_moments_kernel<double>;
    const size_t smem = 1024 * 7 * sizeof(double);
    CH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
    kernel<<<grid, 256, smem, s>>>(static_cast<const double*>(particles), particle_stride,
                                   static_cast<const double*>(survival), survival_stride,
                                   n_particles, bulk, stats, inputs, nx, ny, nz, params, partials);
  }
  CH_LAUNCH_CHECK();
  return CH_OK;
}
}  // namespace

extern "C" int ch_sc_grid_params(const double* stats, int64_t n_beams, const void* energy,
                                 int64_t energy_stride, int32_t energy_dtype, const void* mass_eV,
                                 int32_t mass_dtype, const void* effect_length,
                                 int64_t length_stride, int32_t length_dtype, const void* extent_x,
                                 int64_t extent_x_stride, const void* extent_y,
                                 int64_t extent_y_stride, const void* extent_tau,
                                 int64_t extent_tau_stride, int32_t extent_dtype, int32_t nx,
                                 int32_t ny, int32_t nz, int32_t dtype, double* params,
                                 void* stream) {
  CH_SC_COMMON_CHECKS("ch_sc_grid_params");
  CH_REQUIRE(stats && energy && mass_eV && effect_length && extent_x && extent_y && extent_tau &&
                 params,
             "ch_sc_grid_params: NULL pointer argument");
  CH_REQUIRE(ch::grid_ok(nx, ny, nz),
             "ch_sc_grid_params: grid sizes (%d, %d, %d) must be in [2, 256]", nx, ny, nz);
  ch::GridInputs in{{energy, energy_stride, energy_dtype},
                    {mass_eV, 0, mass_dtype},
                    {effect_length, length_stride, length_dtype},
                    {extent_x, extent_x_stride, extent_dtype},
                    {extent_y, extent_y_stride, extent_dtype},
                    {extent_tau, extent_tau_stride, extent_dtype}};
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned blocks = static_cast<unsigned>((n_beams + 127) / 128);
  if (dtype == CH_F32)
    ch::sc_grid_params_kernel<float><<<blocks, 128, 0, s>>>(stats, n_beams, in, nx, ny, nz, params);
  else
    ch::sc_grid_params_kernel<double><<<blocks, 128, 0, s>>>(stats, n_beams, in, nx, ny, nz, params);
  CH_LAUNCH_CHECK();
  return CH_OK;
}

extern "C" int ch_sc_deposit(const void* particles, int64_t particle_stride, const void* charges,
                             int64_t charge_stride, const void* survival, int64_t survival_stride,
                             const double* params, int64_t n_particles, int64_t n_beams,
                             int32_t nx, int32_t ny, int32_t nz, int32_t dtype, void* rho,
                             void* stream) {
  CH_SC_COMMON_CHECKS("ch_sc_deposit");
  CH_REQUIRE(particles && charges && params && rho && n_particles > 0,
             "ch_sc_deposit: bad arguments");
  CH_REQUIRE(ch::grid_ok(nx, ny, nz), "ch_sc_deposit: bad grid (%d, %d, %d)", nx, ny, nz);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t elem = dtype == CH_F32 ? 4 : 8;
  CH_CUDA(cudaMemsetAsync(rho, 0, elem * nx * 16 * (ny / 2 + 1) * (nz / 2 + 1) * n_beams, s));
  dim3 grid(static_cast<unsigned>((n_particles + 1023) / 1024), static_cast<unsigned>(n_beams));
  if (dtype == CH_F32) {
    const int bulk = ch::bulk_compatible<float>(particles, n_particles, particle_stride);
    ch::sc_deposit_kernel<float><<<grid, 256, 1024 * 7 * sizeof(float), s>>>(
        static_cast<const float*>(particles), particle_stride, static_cast<const float*>(charges),
        charge_stride, static_cast<const float*>(survival), survival_stride, params, n_particles,
        nx, ny, nz, bulk, static_cast<float*>(rho));
  } else {
    const int bulk = ch::bulk_compatible<double>(particles, n_particles, particle_stride);
    auto kernel = ch::sc_deposit_kernel<double>;
    const size_t smem = 1024 * 7 * sizeof(double);
    CH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
    kernel<<<grid, 256, smem, s>>>(
        static_cast<const double*>(particles), particle_stride,
        static_cast<const double*>(charges), charge_stride, static_cast<const double*>(survival),
        survival_stride, params, n_particles, nx, ny, nz, bulk, static_cast<double*>(rho));
  }
  CH_LAUNCH_CHECK();
  return CH_OK;
}

extern "C" int ch_cic_deposit3d(const void* positions, const void* extent, const void* charges,
                                int64_t n_particles, int64_t n_beams, int32_t nx, int32_t ny,
                                int32_t nz, int32_t dtype, void* grid_out, void* stream) {
  CH_SC_COMMON_CHECKS("ch_cic_deposit3d");
  CH_REQUIRE(positions && extent && grid_out && n_particles > 0, "ch_cic_deposit3d: bad arguments");
  CH_REQUIRE(nx > 0 && ny > 0 && nz > 0, "ch_cic_deposit3d: bad grid (%d, %d, %d)", nx, ny, nz);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t elem = dtype == CH_F32 ? 4 : 8;
  CH_CUDA(cudaMemsetAsync(grid_out, 0, elem * nx * ny * nz * n_beams, s));
  dim3 grid(ch::blocks_for(n_particles, 256), static_cast<unsigned>(n_beams));
  if (dtype == CH_F32)
    ch::cic_deposit3d_kernel<float><<<grid, 256, 0, s>>>(
        static_cast<const float*>(positions), static_cast<const float*>(extent),
        static_cast<const float*>(charges), n_particles, nx, ny, nz, static_cast<float*>(grid_out));
  else
    ch::cic_deposit3d_kernel<double><<<grid, 256, 0, s>>>(
        static_cast<const double*>(positions), static_cast<const double*>(extent),
        static_cast<const double*>(charges), n_particles, nx, ny, nz,
        static_cast<double*>(grid_out));
  CH_LAUNCH_CHECK();
  return CH_OK;
}

extern "C" int ch_cic_deposit(const void* positions, const void* extent, const void* charges,
                              int64_t n_particles, int64_t n_beams, int32_t dims, int32_t nx,
                              int32_t ny, int32_t nz, int32_t dtype, void* grid_out,
                              void* stream) {
  CH_SC_COMMON_CHECKS("ch_cic_deposit");
  CH_REQUIRE(dims >= 1 && dims <= 3, "ch_cic_deposit: dims must be 1, 2 or 3, got %d", dims);
  if (dims == 3)
    return ch_cic_deposit3d(positions, extent, charges, n_particles, n_beams, nx, ny, nz, dtype,
                            grid_out, stream);
  CH_REQUIRE(positions && extent && grid_out && n_particles > 0, "ch_cic_deposit: bad arguments");
  CH_REQUIRE(nx > 0 && (dims == 1 || ny > 0), "ch_cic_deposit: bad grid (%d, %d)", nx, ny);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t elem = dtype == CH_F32 ? 4 : 8;
  const int64_t cells = static_cast<int64_t>(nx) * (dims == 2 ? ny : 1);
  CH_CUDA(cudaMemsetAsync(grid_out, 0, elem * cells * n_beams, s));
  dim3 grid(ch::blocks_for(n_particles, 256), static_cast<unsigned>(n_beams));
  auto launch = [&](auto zero) {
    using T = decltype(zero);
    if (dims == 1)
      ch::cic_deposit_low_kernel<T, 1><<<grid, 256, 0, s>>>(
          static_cast<const T*>(positions), static_cast<const T*>(extent),
          static_cast<const T*>(charges), n_particles, nx, 1, static_cast<T*>(grid_out));
    else
      ch::cic_deposit_low_kernel<T, 2><<<grid, 256, 0, s>>>(
          static_cast<const T*>(positions), static_cast<const T*>(extent),
          static_cast<const T*>(charges), n_particles, nx, ny, static_cast<T*>(grid_out));
  };
  if (dtype == CH_F32) launch(0.0f); else launch(0.0);
  CH_LAUNCH_CHECK();
  return CH_OK;
}

extern "C" int ch_sc_deposit_deterministic(
    const void* particles, int64_t particle_stride, const void* charges, int64_t charge_stride,
    const void* survival, int64_t survival_stride, const double* params, int64_t n_particles,
    int64_t n_beams, int32_t nx, int32_t ny, int32_t nz, int32_t dtype, void* scratch, void* rho,
    void* stream) {
  CH_SC_COMMON_CHECKS("ch_sc_deposit_deterministic");
  CH_REQUIRE(particles && charges && params && rho && scratch && n_particles > 0,
             "ch_sc_deposit_deterministic: bad arguments");
  CH_REQUIRE(ch::grid_ok(nx, ny, nz), "ch_sc_deposit_deterministic: bad grid (%d, %d, %d)", nx, ny,
             nz);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t elem = dtype == CH_F32 ? 4 : 8;
  const int64_t cells = static_cast<int64_t>(nx) * ny * nz;
  auto* fixed = static_cast<unsigned long long*>(scratch);
  unsigned long long* max_bits = fixed + n_beams * cells;
  CH_CUDA(cudaMemsetAsync(scratch, 0, sizeof(unsigned long long) * n_beams * (cells + 1), s));
  CH_CUDA(cudaMemsetAsync(rho, 0, elem * nx * 16 * (ny / 2 + 1) * (nz / 2 + 1) * n_beams, s));
  dim3 grid(ch::blocks_for(n_particles, 256), static_cast<unsigned>(n_beams));
  dim3 cell_grid(ch::blocks_for(cells, 256), static_cast<unsigned>(n_beams));
  auto run = [&](auto zero) {
    using T = decltype(zero);
    ch::cic_max_charge_kernel<T><<<grid, 256, 0, s>>>(
        static_cast<const T*>(charges), charge_stride, static_cast<const T*>(survival),
        survival_stride, n_particles, max_bits);
    ch::count_launch();
    ch::sc_deposit_fixed_kernel<T><<<grid, 256, 0, s>>>(
        static_cast<const T*>(particles), particle_stride, static_cast<const T*>(charges),
        charge_stride, static_cast<const T*>(survival), survival_stride, params, n_particles, nx,
        ny, nz, max_bits, fixed);
    ch::count_launch();
    ch::sc_fixed_to_quad_kernel<T><<<cell_grid, 256, 0, s>>>(fixed, max_bits, n_particles, nx, ny,
                                                             nz, static_cast<T*>(rho));
  };
  if (dtype == CH_F32) run(0.0f); else run(0.0);
  CH_LAUNCH_CHECK();
  return CH_OK;
}

extern "C" int ch_cic_deposit_deterministic(const void* positions, const void* extent,
                                            const void* charges, int64_t n_particles,
                                            int64_t n_beams, int32_t dims, int32_t nx, int32_t ny,
                                            int32_t nz, int32_t dtype, void* scratch,
                                            void* grid_out, void* stream) {
  CH_SC_COMMON_CHECKS("ch_cic_deposit_deterministic");
  CH_REQUIRE(dims >= 1 && dims <= 3, "ch_cic_deposit_deterministic: dims must be 1, 2 or 3");
  CH_REQUIRE(positions && extent && grid_out && scratch && n_particles > 0,
             "ch_cic_deposit_deterministic: bad arguments");
  CH_REQUIRE(nx > 0 && (dims < 2 || ny > 0) && (dims < 3 || nz > 0),
             "ch_cic_deposit_deterministic: bad grid (%d, %d, %d)", nx, ny, nz);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t cells = static_cast<int64_t>(nx) * (dims > 1 ? ny : 1) * (dims > 2 ? nz : 1);
  auto* fixed = static_cast<unsigned long long*>(scratch);
  unsigned long long* max_bits = fixed + n_beams * cells;
  CH_CUDA(cudaMemsetAsync(scratch, 0, sizeof(unsigned long long) * n_beams * (cells + 1), s));
  dim3 grid(ch::blocks_for(n_particles, 256), static_cast<unsigned>(n_beams));
  dim3 cell_grid(ch::blocks_for(cells, 256), static_cast<unsigned>(n_beams));
  auto run = [&](auto zero) {
    using T = decltype(zero);
    ch::cic_max_charge_kernel<T><<<grid, 256, 0, s>>>(
        static_cast<const T*>(charges), n_particles, nullptr, 0, n_particles, max_bits);
    ch::count_launch();
    ch::cic_deposit_fixed_kernel<T><<<grid, 256, 0, s>>>(
        static_cast<const T*>(positions), static_cast<const T*>(extent),
        static_cast<const T*>(charges), n_particles, dims, nx, ny, nz, max_bits, fixed);
    ch::count_launch();
    ch::cic_fixed_to_grid_kernel<T><<<cell_grid, 256, 0, s>>>(fixed, max_bits, n_particles, cells,
                                                              static_cast<T*>(grid_out));
  };
  if (dtype == CH_F32) run(0.0f); else run(0.0);
  CH_LAUNCH_CHECK();
  return CH_OK;
}

extern "C" int ch_sc_green_function(const double* params, int64_t n_beams, int32_t nx, int32_t ny,
                                    int32_t nz, int32_t dtype, double* lattice, void* green,
                                    void* stream) {
  CH_SC_COMMON_CHECKS("ch_sc_green_function");
  CH_REQUIRE(params && lattice, "ch_sc_green_function: NULL pointer argument");
  CH_REQUIRE(ch::grid_ok(nx, ny, nz), "ch_sc_green_function: bad grid (%d, %d, %d)", nx, ny, nz);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t points = static_cast<int64_t>(nx + 1) * (ny + 1) * (nz + 1);
  // The lattice kernel is bound by fp64 transcendentals and runs on a side stream next to the
  // deposit, which is bound by L2 reduction requests.  Capping it at ~1 resident CTA per SM
  // (grid-stride loop) leaves the SM's registers and thread slots to the deposit's CTAs, so the
  // two overlap instead of queueing behind each other (128 beams, 100 kicks: 878 ms with 3 CTAs
  // per SM, 847 ms with 1; capping the FFT passes of the chain as well made them the critical
  // path and was not kept).
  static const int ctas_per_sm = [] {
    const char* v = getenv("CH_GREEN_CTAS_PER_SM");  // tuning knob
    const int n = v ? atoi(v) : 1;
    return n > 0 ? n : 1;
  }();
  const int64_t per_beam = (148 * ctas_per_sm + n_beams - 1) / n_beams;
  dim3 grid_a(ch::blocks_for(points, 256, per_beam < 1 ? 1 : per_beam),
              static_cast<unsigned>(n_beams));
  const int far_field = ch::green_far_field(dtype);
  if (ch::green_reference_form())
    ch::sc_green_lattice_kernel<true><<<grid_a, 256, 0, s>>>(params, nx, ny, nz, far_field,
                                                              lattice);
  else
    ch::sc_green_lattice_kernel<false><<<grid_a, 256, 0, s>>>(params, nx, ny, nz, far_field,
                                                               lattice);
  CH_LAUNCH_CHECK();
  if (green == nullptr) return CH_OK;  // the solver only needs the lattice
  dim3 grid_b(ch::blocks_for(static_cast<int64_t>(8) * nx * ny * nz, 256, 148 * 32),
              static_cast<unsigned>(n_beams));
  if (dtype == CH_F32)
    ch::sc_green_mirror_kernel<float><<<grid_b, 256, 0, s>>>(lattice, params, nx, ny, nz,
                                                             far_field, static_cast<float*>(green));
  else
    ch::sc_green_mirror_kernel<double><<<grid_b, 256, 0, s>>>(
        lattice, params, nx, ny, nz, far_field, static_cast<double*>(green));
  CH_LAUNCH_CHECK();
  return CH_OK;
}

extern "C" int ch_sc_green_spectrum(const double* lattice, const double* params, int64_t n_beams,
                                    int32_t nx, int32_t ny, int32_t nz, int32_t dtype,
                                    void* scratch, void* spectrum, void* stream) {
  CH_SC_COMMON_CHECKS("ch_sc_green_spectrum");
  CH_REQUIRE(lattice && params && scratch && spectrum,
             "ch_sc_green_spectrum: NULL pointer argument");
  CH_REQUIRE(ch::grid_ok(nx, ny, nz), "ch_sc_green_spectrum: bad grid (%d, %d, %d)", nx, ny, nz);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t s1 = n_beams * static_cast<int64_t>(nx) * ny * (ch::fft_len(nz) / 2 + 1);
  if (dtype == CH_F32) {
    float* base = static_cast<float*>(scratch);
    return ch::green_spectrum<float>(lattice, params, ch::green_far_field(dtype), n_beams, nx, ny,
                                     nz, base, base + s1, static_cast<float*>(spectrum), s);
  }
  double* base = static_cast<double*>(scratch);
  return ch::green_spectrum<double>(lattice, params, ch::green_far_field(dtype), n_beams, nx, ny,
                                    nz, base, base + s1, static_cast<double*>(spectrum), s);
}

extern "C" int ch_sc_poisson_solve(const void* rho, const void* green_spectrum,
                                   const double* params, int64_t n_beams, int32_t nx, int32_t ny,
                                   int32_t nz, int32_t dtype, void* rho_spectrum, void* phi,
                                   void* stream) {
  CH_SC_COMMON_CHECKS("ch_sc_poisson_solve");
  CH_REQUIRE(rho && green_spectrum && params && rho_spectrum && phi,
             "ch_sc_poisson_solve: NULL pointer argument");
  CH_REQUIRE(ch::grid_ok(nx, ny, nz), "ch_sc_poisson_solve: bad grid (%d, %d, %d)", nx, ny, nz);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == CH_F32)
    return ch::poisson_solve<float>(static_cast<const float*>(rho),
                                    static_cast<const float*>(green_spectrum), params, n_beams,
                                    nx, ny, nz, static_cast<float2*>(rho_spectrum),
                                    static_cast<float*>(phi), s);
  return ch::poisson_solve<double>(static_cast<const double*>(rho),
                                   static_cast<const double*>(green_spectrum), params, n_beams, nx,
                                   ny, nz, static_cast<double2*>(rho_spectrum),
                                   static_cast<double*>(phi), s);
}

extern "C" int ch_sc_field(const void* phi, const double* params, int64_t n_beams, int32_t nx,
                           int32_t ny, int32_t nz, int32_t dtype, void* field, void* stream) {
  CH_SC_COMMON_CHECKS("ch_sc_field");
  CH_REQUIRE(phi && params && field, "ch_sc_field: NULL pointer argument");
  CH_REQUIRE(ch::grid_ok(nx, ny, nz), "ch_sc_field: bad grid (%d, %d, %d)", nx, ny, nz);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  dim3 grid(ch::blocks_for(static_cast<int64_t>(nx) * ny * nz, 256),
            static_cast<unsigned>(n_beams));
  if (dtype == CH_F32)
    ch::sc_field_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(phi), params, nx, ny,
                                                    nz, static_cast<float4*>(field));
  else
    ch::sc_field_kernel<double><<<grid, 256, 0, s>>>(static_cast<const double*>(phi), params, nx,
                                                     ny, nz, static_cast<double4*>(field));
  CH_LAUNCH_CHECK();
  return CH_OK;
}

namespace {
int launch_field_bricks(const float* phi, const double* params, int beam0, int group, int nx,
                        int ny, int nz, float* bricks, cudaStream_t s);
}

extern "C" int ch_sc_field_bricks(const void* phi, const double* params, int64_t n_beams,
                                  int32_t nx, int32_t ny, int32_t nz, int32_t dtype, void* bricks,
                                  void* stream) {
  CH_SC_COMMON_CHECKS("ch_sc_field_bricks");
  CH_REQUIRE(dtype == CH_F32, "ch_sc_field_bricks: float32 only (float64 uses ch_sc_field)");
  CH_REQUIRE(phi && params && bricks, "ch_sc_field_bricks: NULL pointer argument");
  CH_REQUIRE(ch::grid_ok(nx, ny, nz), "ch_sc_field_bricks: bad grid (%d, %d, %d)", nx, ny, nz);
  return launch_field_bricks(static_cast<const float*>(phi), params, 0, static_cast<int>(n_beams),
                             nx, ny, nz, static_cast<float*>(bricks),
                             static_cast<cudaStream_t>(stream));
}

namespace {
// Brick gather of `group` beams starting at beam0 (the bricks buffer holds that group).
int launch_gather_bricks(const float* particles_in, int64_t particle_stride, const float* bricks,
                         const double* params, int64_t n_particles, int beam0, int group, int nx,
                         int ny, int nz, float* particles_out, float* forces_out,
                         const ch::GatherFusion<float>* fusion, cudaStream_t s) {
  // Particles per thread: 4 (tiles of 1024) unless the launch is only a few waves of CTAs long
  // (3 CTAs of 256 threads per SM): one beam of 1e6 particles is 977 tiles = 2.2 waves with the
  // last one a fifth full; tiles of 768 (3 full waves of shorter CTAs) or 1280 (2 waves) finish
  // sooner -- config 4 as a CUDA graph 11.96 -> 11.35 ms.
  int per_thread = 4;
  {
    const int64_t slots = 148 * 3;
    auto waves = [&](int p) {
      const int64_t ctas = (n_particles + 256 * p - 1) / (256 * p) * group;
      return (ctas + slots - 1) / slots;
    };
    if (waves(4) < 20) {
      int64_t best = waves(4) * 4;
      for (int p : {3, 5})
        if (waves(p) * p < best) {
          best = waves(p) * p;
          per_thread = p;
        }
    }
  }
  const int tile = 256 * per_thread;
  dim3 grid(static_cast<unsigned>((n_particles + tile - 1) / tile), static_cast<unsigned>(group));
  const int bulk_in = ch::bulk_compatible<float>(particles_in, n_particles, particle_stride);
  const int bulk_out = ch::bulk_compatible<float>(particles_out, n_particles, n_particles * 7);
  const size_t smem = static_cast<size_t>(tile) * 7 * sizeof(float);
  auto launch = [&](auto kernel, const ch::GatherFusion<float>& f) {
    kernel<<<grid, 256, smem, s>>>(particles_in, particle_stride, bricks, params, n_particles, nx,
                                   ny, nz, bulk_in, bulk_out, particles_out, forces_out, f, beam0);
  };
  const ch::GatherFusion<float> none{};
  auto with_tile = [&](auto p) {
    constexpr int P = decltype(p)::value;
    fusion ? launch(ch::sc_gather_brick_kernel<true, P>, *fusion)
           : launch(ch::sc_gather_brick_kernel<false, P>, none);
  };
  if (per_thread == 5) with_tile(std::integral_constant<int, 5>{});
  else if (per_thread == 3) with_tile(std::integral_constant<int, 3>{});
  else with_tile(std::integral_constant<int, 4>{});
  CH_LAUNCH_CHECK();
  return CH_OK;
}

int launch_field_bricks(const float* phi, const double* params, int beam0, int group, int nx,
                        int ny, int nz, float* bricks, cudaStream_t s) {
  const size_t smem = sizeof(float) * 3 * 2 * (ch::kBrickRows + 1) * (nz + 1);
  if (ch::allow_smem(ch::sc_field_brick_kernel, smem) != CH_OK) return CH_ECUDA;
  dim3 grid((ny + ch::kBrickRows - 1) / ch::kBrickRows, nx, static_cast<unsigned>(group));
  ch::sc_field_brick_kernel<<<grid, 256, smem, s>>>(phi, params, nx, ny, nz, beam0, bricks);
  CH_LAUNCH_CHECK();
  return CH_OK;
}

template <typename T>
int launch_gather(const void* particles_in, int64_t particle_stride, const void* field,
                  int field_layout, const double* params, int64_t n_particles, int64_t n_beams,
                  int nx, int ny, int nz, void* particles_out, void* forces_out,
                  const ch::GatherFusion<T>* fusion, cudaStream_t s) {
  using F4 = typename ch::Field4<T>::type;
  if constexpr (std::is_same<T, float>::value) {
    if (field_layout == CH_SC_FIELD_BRICKS)
      return launch_gather_bricks(static_cast<const float*>(particles_in), particle_stride,
                                  static_cast<const float*>(field), params, n_particles, 0,
                                  static_cast<int>(n_beams), nx, ny, nz,
                                  static_cast<float*>(particles_out),
                                  static_cast<float*>(forces_out), fusion, s);
  }
  dim3 grid(static_cast<unsigned>((n_particles + 1023) / 1024), static_cast<unsigned>(n_beams));
  const int bulk_in = ch::bulk_compatible<T>(particles_in, n_particles, particle_stride);
  const int bulk_out = ch::bulk_compatible<T>(particles_out, n_particles, n_particles * 7);
  const size_t smem = 1024 * 7 * sizeof(T);
  auto launch = [&](auto kernel, const ch::GatherFusion<T>& f) -> int {
    CH_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(smem)));
    kernel<<<grid, 256, smem, s>>>(static_cast<const T*>(particles_in), particle_stride,
                                   static_cast<const F4*>(field), params, n_particles, nx, ny, nz,
                                   bulk_in, bulk_out, static_cast<T*>(particles_out),
                                   static_cast<T*>(forces_out), f);
    return CH_OK;
  };
  const int status = fusion ? launch(ch::sc_gather_kick_kernel<T, true>, *fusion)
                            : launch(ch::sc_gather_kick_kernel<T, false>, ch::GatherFusion<T>{});
  if (status != CH_OK) return status;
  CH_LAUNCH_CHECK();
  return CH_OK;
}

// What the two fused entry points share.  `phi` != NULL: float32, field bricks built per group of
// `group_beams` beams into `field` (a buffer for one group) right before the group's gather, so
// that the bricks are read back from L2, not from HBM.
struct FusedArgs {
  const void* particles_in;
  int64_t particle_stride;
  const void* field;
  int32_t field_layout;
  const float* phi;
  int32_t group_beams;
  const double* params;
  int64_t n_particles, n_beams;
  int32_t nx, ny, nz, dtype;
  const void* records;
  int64_t record_stride;
  const void* survival;
  int64_t survival_stride;
  double* next_stats;
  double* next_params;
  ch::GridInputs next_in;
  int32_t next_nx, next_ny, next_nz;
  void* particles_out;
  void* forces_out;
};

int gather_fused(const FusedArgs& a, cudaStream_t s) {
  const bool fused = a.records != nullptr || a.next_stats != nullptr;
  if (a.next_stats != nullptr)
    CH_CUDA(cudaMemsetAsync(a.next_stats, 0, sizeof(double) * CH_SC_STATS * a.n_beams, s));
  int const_maps = 0;
  if (a.dtype == CH_F32 && a.field_layout == CH_SC_FIELD_BRICKS && a.records != nullptr) {
    const int64_t n_maps = a.record_stride == 0 ? 1 : a.n_beams;
    // One constant buffer per device: it belongs to the first stream that uses it.  Work on
    // that stream is ordered, so the copy below cannot overtake a gather that still reads the
    // previous maps; any other stream (a second host thread tracking on the same device) takes
    // the shared-memory path instead of racing for the buffer.
    static std::atomic<uintptr_t> owner[64];  // 0: unclaimed, else stream handle + 1
    int device = 0;
    CH_CUDA(cudaGetDevice(&device));
    bool mine = false;
    if (device >= 0 && device < 64) {
      const uintptr_t me = reinterpret_cast<uintptr_t>(s) + 1;
      uintptr_t expected = 0;
      mine = owner[device].compare_exchange_strong(expected, me) || expected == me;
    }
    if (mine && n_maps <= ch::kConstMaps) {
      void* symbol = nullptr;
      CH_CUDA(cudaGetSymbolAddress(&symbol, ch::c_gather_maps));
      CH_CUDA(cudaMemcpy2DAsync(
          symbol, sizeof(float) * ch::kConstMapPitch, a.records,
          sizeof(float) * (a.record_stride == 0 ? ch::kConstMapPitch : a.record_stride),
          sizeof(float) * (CH_RECORD_HEADER + CH_RECORD_MAP), static_cast<size_t>(n_maps),
          cudaMemcpyDeviceToDevice, s));
      const_maps = 1;
    }
  }
  // several waves of CTAs: grid parameters in a follow-up launch instead of a last-CTA epilogue
  const int64_t ctas = ((a.n_particles + 1023) / 1024) * a.n_beams;
  const bool separate_params = a.next_stats != nullptr && ctas > 148 * 3 * 4;
  double* in_kernel_params = separate_params ? nullptr : a.next_params;
  int status;
  if (a.dtype == CH_F32) {
    const ch::GatherFusion<float> f{static_cast<const float*>(a.records), a.record_stride,
                                    const_maps, static_cast<const float*>(a.survival),
                                    a.survival_stride, a.next_stats, in_kernel_params, a.next_in,
                                    a.next_nx, a.next_ny, a.next_nz};
    if (a.phi != nullptr) {
      status = CH_OK;
      const int group = a.group_beams < 1 ? 1 : a.group_beams;
      for (int64_t beam0 = 0; beam0 < a.n_beams && status == CH_OK; beam0 += group) {
        const int g = static_cast<int>(a.n_beams - beam0 < group ? a.n_beams - beam0 : group);
        float* bricks = static_cast<float*>(const_cast<void*>(a.field));
        status = launch_field_bricks(a.phi, a.params, static_cast<int>(beam0), g, a.nx, a.ny,
                                     a.nz, bricks, s);
        if (status != CH_OK) break;
        status = launch_gather_bricks(static_cast<const float*>(a.particles_in),
                                      a.particle_stride, bricks, a.params, a.n_particles,
                                      static_cast<int>(beam0), g, a.nx, a.ny, a.nz,
                                      static_cast<float*>(a.particles_out),
                                      static_cast<float*>(a.forces_out), fused ? &f : nullptr, s);
      }
    } else {
      status = launch_gather<float>(a.particles_in, a.particle_stride, a.field, a.field_layout,
                                    a.params, a.n_particles, a.n_beams, a.nx, a.ny, a.nz,
                                    a.particles_out, a.forces_out, fused ? &f : nullptr, s);
    }
  } else {
    const ch::GatherFusion<double> f{static_cast<const double*>(a.records), a.record_stride, 0,
                                     static_cast<const double*>(a.survival), a.survival_stride,
                                     a.next_stats, in_kernel_params, a.next_in, a.next_nx,
                                     a.next_ny, a.next_nz};
    status = launch_gather<double>(a.particles_in, a.particle_stride, a.field, a.field_layout,
                                   a.params, a.n_particles, a.n_beams, a.nx, a.ny, a.nz,
                                   a.particles_out, a.forces_out, fused ? &f : nullptr, s);
  }
  if (status != CH_OK || !separate_params) return status;
  const unsigned blocks = static_cast<unsigned>((a.n_beams + 127) / 128);
  if (a.dtype == CH_F32)
    ch::sc_grid_params_kernel<float><<<blocks, 128, 0, s>>>(a.next_stats, a.n_beams, a.next_in,
                                                            a.next_nx, a.next_ny, a.next_nz,
                                                            a.next_params);
  else
    ch::sc_grid_params_kernel<double><<<blocks, 128, 0, s>>>(a.next_stats, a.n_beams, a.next_in,
                                                             a.next_nx, a.next_ny, a.next_nz,
                                                             a.next_params);
  CH_LAUNCH_CHECK();
  return CH_OK;
}
}  // namespace

#define CH_SC_LAYOUT_CHECK(fn)                                                              \
  CH_REQUIRE(field_layout == CH_SC_FIELD_NODES ||                                            \
                 (field_layout == CH_SC_FIELD_BRICKS && dtype == CH_F32),                    \
             fn ": field_layout %d not available for dtype %d", field_layout, dtype)

extern "C" int ch_sc_gather_kick(const void* particles_in, int64_t particle_stride,
                                 const void* field, int32_t field_layout, const double* params,
                                 int64_t n_particles, int64_t n_beams, int32_t nx, int32_t ny,
                                 int32_t nz, int32_t dtype, void* particles_out, void* forces_out,
                                 void* stream) {
  CH_SC_COMMON_CHECKS("ch_sc_gather_kick");
  CH_REQUIRE(particles_in && field && params && particles_out && n_particles > 0,
             "ch_sc_gather_kick: bad arguments");
  CH_REQUIRE(ch::grid_ok(nx, ny, nz), "ch_sc_gather_kick: bad grid (%d, %d, %d)", nx, ny, nz);
  CH_SC_LAYOUT_CHECK("ch_sc_gather_kick");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == CH_F32)
    return launch_gather<float>(particles_in, particle_stride, field, field_layout, params,
                                n_particles, n_beams, nx, ny, nz, particles_out, forces_out,
                                nullptr, s);
  return launch_gather<double>(particles_in, particle_stride, field, field_layout, params,
                               n_particles, n_beams, nx, ny, nz, particles_out, forces_out,
                               nullptr, s);
}

#define CH_SC_NEXT_CHECKS(fn)                                                                   \
  if (next_stats != nullptr) {                                                                  \
    CH_REQUIRE(next_params && energy && mass_eV && next_effect_length && next_extent_x &&       \
                   next_extent_y && next_extent_tau,                                            \
               fn ": NULL pointer among the next kick's inputs");                               \
    CH_REQUIRE(ch::grid_ok(next_nx, next_ny, next_nz), fn ": bad next grid (%d, %d, %d)",       \
               next_nx, next_ny, next_nz);                                                      \
  }                                                                                             \
  const ch::GridInputs next_in{{energy, energy_stride, energy_dtype},                           \
                               {mass_eV, 0, mass_dtype},                                        \
                               {next_effect_length, next_length_stride, next_length_dtype},     \
                               {next_extent_x, next_extent_x_stride, next_extent_dtype},        \
                               {next_extent_y, next_extent_y_stride, next_extent_dtype},        \
                               {next_extent_tau, next_extent_tau_stride, next_extent_dtype}}

extern "C" int ch_sc_gather_kick_fused(
    const void* particles_in, int64_t particle_stride, const void* field, int32_t field_layout,
    const double* params, int64_t n_particles, int64_t n_beams, int32_t nx, int32_t ny, int32_t nz,
    int32_t dtype, const void* records, int64_t record_stride, const void* survival,
    int64_t survival_stride, double* next_stats, double* next_params, const void* energy,
    int64_t energy_stride, int32_t energy_dtype, const void* mass_eV, int32_t mass_dtype,
    const void* next_effect_length, int64_t next_length_stride, int32_t next_length_dtype,
    const void* next_extent_x, int64_t next_extent_x_stride, const void* next_extent_y,
    int64_t next_extent_y_stride, const void* next_extent_tau, int64_t next_extent_tau_stride,
    int32_t next_extent_dtype, int32_t next_nx, int32_t next_ny, int32_t next_nz,
    void* particles_out, void* stream) {
  CH_SC_COMMON_CHECKS("ch_sc_gather_kick_fused");
  CH_REQUIRE(particles_in && field && params && particles_out && n_particles > 0,
             "ch_sc_gather_kick_fused: bad arguments");
  CH_REQUIRE(ch::grid_ok(nx, ny, nz), "ch_sc_gather_kick_fused: bad grid (%d, %d, %d)", nx, ny, nz);
  CH_SC_LAYOUT_CHECK("ch_sc_gather_kick_fused");
  CH_REQUIRE(records != nullptr || next_stats != nullptr,
             "ch_sc_gather_kick_fused: nothing to fuse (use ch_sc_gather_kick)");
  CH_SC_NEXT_CHECKS("ch_sc_gather_kick_fused");
  const FusedArgs a{particles_in, particle_stride, field,          field_layout, nullptr,
                    0,            params,          n_particles,    n_beams,      nx,
                    ny,           nz,              dtype,          records,      record_stride,
                    survival,     survival_stride, next_stats,     next_params,  next_in,
                    next_nx,      next_ny,         next_nz,        particles_out, nullptr};
  return gather_fused(a, static_cast<cudaStream_t>(stream));
}

extern "C" int ch_sc_field_gather(
    const void* particles_in, int64_t particle_stride, const void* phi, void* bricks,
    int32_t group_beams, const double* params, int64_t n_particles, int64_t n_beams, int32_t nx,
    int32_t ny, int32_t nz, int32_t dtype, const void* records, int64_t record_stride,
    const void* survival, int64_t survival_stride, double* next_stats, double* next_params,
    const void* energy, int64_t energy_stride, int32_t energy_dtype, const void* mass_eV,
    int32_t mass_dtype, const void* next_effect_length, int64_t next_length_stride,
    int32_t next_length_dtype, const void* next_extent_x, int64_t next_extent_x_stride,
    const void* next_extent_y, int64_t next_extent_y_stride, const void* next_extent_tau,
    int64_t next_extent_tau_stride, int32_t next_extent_dtype, int32_t next_nx, int32_t next_ny,
    int32_t next_nz, void* particles_out, void* forces_out, void* stream) {
  CH_SC_COMMON_CHECKS("ch_sc_field_gather");
  CH_REQUIRE(dtype == CH_F32, "ch_sc_field_gather: float32 only");
  CH_REQUIRE(particles_in && phi && bricks && params && particles_out && n_particles > 0 &&
                 group_beams > 0,
             "ch_sc_field_gather: bad arguments");
  CH_REQUIRE(ch::grid_ok(nx, ny, nz), "ch_sc_field_gather: bad grid (%d, %d, %d)", nx, ny, nz);
  CH_SC_NEXT_CHECKS("ch_sc_field_gather");
  const FusedArgs a{particles_in, particle_stride, bricks,      CH_SC_FIELD_BRICKS,
                    static_cast<const float*>(phi), group_beams, params, n_particles, n_beams,
                    nx, ny, nz, dtype, records, record_stride, survival, survival_stride,
                    next_stats, next_params, next_in, next_nx, next_ny, next_nz, particles_out,
                    forces_out};
  return gather_fused(a, static_cast<cudaStream_t>(stream));
}

// ---------------------------------------------------------------------------------------
// one call for the grid half of a kick
// ---------------------------------------------------------------------------------------
namespace {
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};

// A high-priority side stream and two events per device, created on first use (before any graph
// capture: GraphedTrack warms the path up first).
int side_stream(SideStream** out) {
  static SideStream table[64];
  static std::atomic<int> ready[64];
  int device = 0;
  CH_CUDA(cudaGetDevice(&device));
  CH_REQUIRE(device >= 0 && device < 64, "ch_sc_solve: device index %d out of range", device);
  SideStream& entry = table[device];
  if (ready[device].load(std::memory_order_acquire) != 2) {
    int expected = 0;
    if (ready[device].compare_exchange_strong(expected, 1)) {
      int least = 0, greatest = 0;
      CH_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
      CH_CUDA(cudaStreamCreateWithPriority(&entry.stream, cudaStreamNonBlocking, greatest));
      CH_CUDA(cudaEventCreateWithFlags(&entry.fork, cudaEventDisableTiming));
      CH_CUDA(cudaEventCreateWithFlags(&entry.join, cudaEventDisableTiming));
      ready[device].store(2, std::memory_order_release);
    } else {
      while (ready[device].load(std::memory_order_acquire) != 2) {
      }
    }
  }
  *out = &entry;
  return CH_OK;
}
}  // namespace

extern "C" int ch_sc_solve(const void* particles, int64_t particle_stride, const void* charges,
                           int64_t charge_stride, const void* survival, int64_t survival_stride,
                           const double* params, int64_t n_particles, int64_t n_beams, int32_t nx,
                           int32_t ny, int32_t nz, int32_t dtype, void* rho, double* lattice,
                           void* green_scratch, void* green_spectrum, void* rho_spectrum,
                           void* phi, void* stream) {
  CH_SC_COMMON_CHECKS("ch_sc_solve");
  SideStream* side = nullptr;
  int status = side_stream(&side);
  if (status != CH_OK) return status;
  cudaStream_t main = static_cast<cudaStream_t>(stream);
  // The Green-function chain only needs the grid parameters: it runs on the side stream next to
  // the deposit and joins right before the x convolution, the first kernel that reads its
  // spectrum (not before the Poisson solve: the charge's quad sum, z and y passes do not need it;
  // config 4 as a CUDA graph 11.8 -> 11.2 ms).  Starting the chain AFTER the deposit when there
  // are many beams -- the deposit, bound by the L2's reduction rate, takes 2.6 instead of 1.97 ms
  // for 128 beams while the chain shares the SMs -- measured slower (716 instead of 710 ms per 100
  // kicks).
  CH_CUDA(cudaEventRecord(side->fork, main));
  CH_CUDA(cudaStreamWaitEvent(side->stream, side->fork, 0));
  status = ch_sc_green_function(params, n_beams, nx, ny, nz, dtype, lattice, nullptr, side->stream);
  if (status != CH_OK) return status;
  status = ch_sc_green_spectrum(lattice, params, n_beams, nx, ny, nz, dtype, green_scratch,
                                green_spectrum, side->stream);
  if (status != CH_OK) return status;
  CH_CUDA(cudaEventRecord(side->join, side->stream));
  status = ch_sc_deposit(particles, particle_stride, charges, charge_stride, survival,
                         survival_stride, params, n_particles, n_beams, nx, ny, nz, dtype, rho,
                         stream);
  if (status != CH_OK) return status;
  CH_REQUIRE(rho && green_spectrum && params && rho_spectrum && phi,
             "ch_sc_solve: NULL pointer argument");
  if (dtype == CH_F32)
    return ch::poisson_solve<float>(static_cast<const float*>(rho),
                                    static_cast<const float*>(green_spectrum), params, n_beams,
                                    nx, ny, nz, static_cast<float2*>(rho_spectrum),
                                    static_cast<float*>(phi), main, side->join);
  return ch::poisson_solve<double>(static_cast<const double*>(rho),
                                   static_cast<const double*>(green_spectrum), params, n_beams, nx,
                                   ny, nz, static_cast<double2*>(rho_spectrum),
                                   static_cast<double*>(phi), main, side->join);
}

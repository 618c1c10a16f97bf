// Beam diagnostics on the output side of the hot path: the image an active
// Screen records, as one pass over the particles.
//
// Replaces Screen.reading for a ParticleBeam (cheetah/accelerator/screen.py:241-344): the
// misalignment shift of the read beam (:199-215), the charge weights |q| * survival, the 2-D
// cloud-in-cell deposit (cheetah/utils/cloud_in_cell.py:129-241) or the torch.histogramdd call
// with the screen's pixel bin edges (:296-315), and the final transpose to (height, width).
// Each CTA streams a tile of particles (TMA bulk copy) and issues float atomics into the image,
// which (<= 20 MB at the largest screens) lives in L2.
#include <algorithm>

#include "ch_common.cuh"

namespace ch {
namespace {

template <typename T>
struct ScreenArgs {
  const T* particles;
  const T* charges;
  const T* survival;  // may be null (ones)
  const T* misalignment;
  const T* pixel_size;
  const T* edges_x;  // histogram only: nx + 1 / ny + 1 bin edges
  const T* edges_y;
  T* image;
  int64_t particle_stride, charge_stride, survival_stride, misalignment_stride;
  int64_t n_particles;
  int32_t resolution_x, resolution_y, nx, ny, method, bulk;
};

__device__ __forceinline__ float floor_t(float x) { return floorf(x); }
__device__ __forceinline__ double floor_t(double x) { return floor(x); }
__device__ __forceinline__ float abs_t(float x) { return fabsf(x); }
__device__ __forceinline__ double abs_t(double x) { return fabs(x); }

// index of the bin holding v for sorted edges e[0..n]: number of edges <= v, minus one; the
// right-most edge belongs to the last bin, values outside [e[0], e[n]] are dropped (-1)
// (ATen histogramdd with explicit bin edges: binary search)
template <typename T>
__device__ __forceinline__ int bin_of(const T* e, int n, T v) {
  if (!(v >= e[0]) || !(v <= e[n])) return -1;
  int lo = 0, hi = n + 1;  // first index with e[idx] > v
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (e[mid] <= v) lo = mid + 1; else hi = mid;
  }
  const int pos = lo - 1;
  return pos == n ? n - 1 : pos;
}

template <typename T>
__global__ void __launch_bounds__(256) screen_image_kernel(const ScreenArgs<T> a) {
  constexpr int TP = 1024;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);
  __shared__ uint64_t bar;
  const int64_t b = blockIdx.y;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * TP;
  const int count = static_cast<int>(min(static_cast<int64_t>(TP), a.n_particles - n0));
  if (a.bulk && threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  uint32_t phase = 0;
  cta_load_tile(tile, a.particles + b * a.particle_stride + n0 * 7, count * 7, a.bulk != 0, &bar,
                phase);

  const T mis_x = a.misalignment[b * a.misalignment_stride];
  const T mis_y = a.misalignment[b * a.misalignment_stride + 1];
  // screen.py:137-146: extent = -+ resolution * pixel_size / 2 in the pixel dtype
  const T right = T(a.resolution_x) * a.pixel_size[0] / T(2), left = -right;
  const T top = T(a.resolution_y) * a.pixel_size[1] / T(2), bottom = -top;
  const T bwx = (right - left) / T(a.nx), bwy = (top - bottom) / T(a.ny);
  const T* q = a.charges + b * a.charge_stride + n0;
  const T* w = a.survival ? a.survival + b * a.survival_stride + n0 : nullptr;
  T* image = a.image + b * static_cast<int64_t>(a.nx) * a.ny;

  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    const T x = tile[i * 7 + 0] - mis_x;
    const T y = tile[i * 7 + 2] - mis_y;
    const T weight = abs_t(q[i]) * (w ? w[i] : T(1));
    if (a.method == 1) {  // histogram
      const int ix = bin_of(a.edges_x, a.nx, x), iy = bin_of(a.edges_y, a.ny, y);
      if (ix >= 0 && iy >= 0) atomicAdd(image + static_cast<int64_t>(iy) * a.nx + ix, weight);
      continue;
    }
    // cloud-in-cell (cloud_in_cell.py:129-241); NaNs compare false and are dropped
    if (!(x >= left && x <= right && y >= bottom && y <= top)) continue;
    const T px = (x - left) / bwx - T(0.5), py = (y - bottom) / bwy - T(0.5);
    const T fx0 = floor_t(px), fy0 = floor_t(py);
    const int ix = static_cast<int>(fx0), iy = static_cast<int>(fy0);
    const T fx = px - fx0, fy = py - fy0;
    const T wx[2] = {(ix >= 0 && ix < a.nx) ? T(1) - fx : T(0),
                     (ix + 1 >= 0 && ix + 1 < a.nx) ? fx : T(0)};
    const T wy[2] = {(iy >= 0 && iy < a.ny) ? T(1) - fy : T(0),
                     (iy + 1 >= 0 && iy + 1 < a.ny) ? fy : T(0)};
#pragma unroll
    for (int dx = 0; dx < 2; ++dx)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const T value = weight * wx[dx] * wy[dy];
        if (value != T(0))
          atomicAdd(image + static_cast<int64_t>(iy + dy) * a.nx + (ix + dx), value);
      }
  }
}

// ---- method = "kde" (cheetah/utils/kde.py:6-204 through screen.py:312-326) ---------------------
// image[j][i] = sum_n Kx[n][i] Ky[n][j] / (sum_ij ... + 1e-10),
//   Kx[n][i] = max(w_n exp(-((x_n - cx_i) / sigma)^2 / 2) / sqrt(2 pi sigma^2), tiny),  Ky without w.
// The reference materialises Kx, Ky as (N, bins) arrays and multiplies them (N x nx x ny
// multiply-adds: 5e15 for 1e6 particles on a full 2448 x 2040 screen).  A Gaussian is below
// 2e-11 (float32) / 1e-14 (float64) of its peak beyond 7 / 8 sigma, so every particle only
// touches the window of pixels within that radius: one WARP per particle, lanes across the window's
// columns (each computes its Kx once), the row factors Ky computed by one lane each and broadcast
// by shuffle, one coalesced row of reductions per window row.  What the truncation drops is below
// the rounding of the sums; the clamp to `tiny` only matters where the reference's image is ~1e-38
// of its peak.  Totals for the normalisation are accumulated in fp64.
template <typename T>
struct KdeArgs {
  const T* particles;
  const T* charges;
  const T* survival;
  const T* misalignment;
  const T* centers_x;
  const T* centers_y;
  const T* bandwidth;
  T* image;
  double* totals;
  int64_t particle_stride, charge_stride, survival_stride, misalignment_stride;
  int64_t n_particles;
  int32_t nx, ny, bulk;
};

__device__ __forceinline__ float exp_t(float x) { return expf(x); }
__device__ __forceinline__ double exp_t(double x) { return exp(x); }
__device__ __forceinline__ float tiny_of(float) { return 1.17549435e-38f; }
__device__ __forceinline__ double tiny_of(double) { return 2.2250738585072014e-308; }

template <typename T>
__global__ void __launch_bounds__(256) screen_kde_kernel(const KdeArgs<T> a) {
  constexpr int TP = 1024;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);
  __shared__ uint64_t bar;
  const int64_t b = blockIdx.y;
  const int64_t n0 = static_cast<int64_t>(blockIdx.x) * TP;
  const int count = static_cast<int>(min(static_cast<int64_t>(TP), a.n_particles - n0));
  if (a.bulk && threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  __syncthreads();
  uint32_t phase = 0;
  cta_load_tile(tile, a.particles + b * a.particle_stride + n0 * 7, count * 7, a.bulk != 0, &bar,
                phase);

  const T mis_x = a.misalignment[b * a.misalignment_stride];
  const T mis_y = a.misalignment[b * a.misalignment_stride + 1];
  const T sigma = a.bandwidth[0];
  const T inv_norm = T(1) / sqrt(T(2) * T(3.14159265358979323846) * sigma * sigma);
  const T radius = (sizeof(T) == 4 ? T(7) : T(8)) * sigma;
  const T x0 = a.centers_x[0], y0 = a.centers_y[0];
  const T step_x = a.nx > 1 ? (a.centers_x[a.nx - 1] - x0) / T(a.nx - 1) : T(1);
  const T step_y = a.ny > 1 ? (a.centers_y[a.ny - 1] - y0) / T(a.ny - 1) : T(1);
  const T* q = a.charges + b * a.charge_stride + n0;
  const T* w = a.survival ? a.survival + b * a.survival_stride + n0 : nullptr;
  T* image = a.image + b * static_cast<int64_t>(a.nx) * a.ny;
  const T tiny = tiny_of(T(0));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double total = 0.0;

  for (int i = warp; i < count; i += 8) {
    const T x = tile[i * 7 + 0] - mis_x;
    const T y = tile[i * 7 + 2] - mis_y;
    const T weight = abs_t(q[i]) * (w ? w[i] : T(1));
    if (!(x == x) || !(y == y)) continue;  // NaN coordinates never reach a pixel
    // window of pixel centres within `radius` (one guard pixel for the rounding of the division)
    const int ix_lo = max(0, static_cast<int>(floor_t((x - radius - x0) / step_x)) - 1);
    const int ix_hi = min(a.nx - 1, static_cast<int>(floor_t((x + radius - x0) / step_x)) + 2);
    const int iy_lo = max(0, static_cast<int>(floor_t((y - radius - y0) / step_y)) - 1);
    const int iy_hi = min(a.ny - 1, static_cast<int>(floor_t((y + radius - y0) / step_y)) + 2);
    if (ix_lo > ix_hi || iy_lo > iy_hi) continue;
    for (int yb = iy_lo; yb <= iy_hi; yb += 32) {
      const int jy = yb + lane;
      T ky = T(0);
      if (jy <= iy_hi) {
        const T r = (y - a.centers_y[jy]) / sigma;
        ky = max(exp_t(T(-0.5) * r * r) * inv_norm, tiny);
      }
      const int rows = min(32, iy_hi - yb + 1);
      for (int xb = ix_lo; xb <= ix_hi; xb += 32) {
        const int jx = xb + lane;
        T kx = T(0);
        if (jx <= ix_hi) {
          const T r = (x - a.centers_x[jx]) / sigma;
          kx = max(weight * exp_t(T(-0.5) * r * r) * inv_norm, tiny);
        }
        T row_sum = T(0);
        for (int r = 0; r < rows; ++r) {
          const T factor = __shfl_sync(0xffffffffu, ky, r);
          const T value = kx * factor;
          if (jx <= ix_hi) {
            atomicAdd(image + static_cast<int64_t>(yb + r) * a.nx + jx, value);
            row_sum += value;
          }
        }
        total += static_cast<double>(row_sum);
      }
    }
  }
  total = warp_sum(total);
  if (lane == 0 && total != 0.0) atomicAdd(a.totals + b, total);
}

template <typename T>
__global__ void __launch_bounds__(256)
screen_kde_normalise_kernel(T* image, const double* totals, int64_t pixels) {
  const int64_t b = blockIdx.y;
  // kde.py:107-110: joint / (joint.sum() + epsilon) in the image dtype
  const T normalisation = static_cast<T>(totals[b]) + T(1e-10);
  T* im = image + b * pixels;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < pixels;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    im[i] = im[i] / normalisation;
}

// ---- ParameterBeam on a Screen (screen.py:251-289): bivariate normal density of (x, y) on the
// pixel grid arange(left, right, step).  torch.arange with tensor bounds returns the DEFAULT dtype
// (float32) whatever the screen's dtype and evaluates start + i * step in double, so the grid
// points are float32-rounded here too before they meet the beam's mu / cov.
template <typename T>
__global__ void __launch_bounds__(256)
screen_gaussian_kernel(const T* mu, const T* cov, const T* misalignment, double left,
                       double step_x, int nx, double bottom, double step_y, int ny, T* image) {
  const T mx = mu[0] - misalignment[0], my = mu[2] - misalignment[1];
  const T sxx = cov[0], sxy = cov[2], syy = cov[2 * 7 + 2];
  const T det = sxx * syy - sxy * sxy;
  const T norm = T(1) / (T(2) * T(3.14159265358979323846) * sqrt(det));
  const int64_t pixels = static_cast<int64_t>(nx) * ny;
  for (int64_t p = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; p < pixels;
       p += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int i = static_cast<int>(p % nx), j = static_cast<int>(p / nx);
    const T dx = static_cast<T>(static_cast<float>(left + i * step_x)) - mx;
    const T dy = static_cast<T>(static_cast<float>(bottom + j * step_y)) - my;
    const T quad = (syy * dx * dx - T(2) * sxy * dx * dy + sxx * dy * dy) / det;
    image[p] = exp_t(T(-0.5) * quad) * norm;
  }
}

}  // namespace
}  // namespace ch

extern "C" int ch_screen_image(const void* particles, int64_t particle_stride, const void* charges,
                               int64_t charge_stride, const void* survival,
                               int64_t survival_stride, const void* misalignment,
                               int64_t misalignment_stride, const void* pixel_size,
                               int32_t resolution_x, int32_t resolution_y, int32_t binning,
                               int32_t method, const void* edges_x, const void* edges_y,
                               int64_t n_particles, int64_t n_beams, int32_t dtype, void* image,
                               void* stream) {
  CH_REQUIRE(dtype == CH_F32 || dtype == CH_F64, "ch_screen_image: bad dtype %d", dtype);
  CH_REQUIRE(particles && charges && misalignment && pixel_size && image,
             "ch_screen_image: NULL pointer argument");
  CH_REQUIRE(n_particles > 0 && n_beams > 0 && n_beams <= 65535,
             "ch_screen_image: need particles and 1..65535 beams");
  CH_REQUIRE(resolution_x > 0 && resolution_y > 0 && binning > 0 &&
                 resolution_x / binning > 0 && resolution_y / binning > 0,
             "ch_screen_image: bad resolution (%d, %d) / binning %d", resolution_x, resolution_y,
             binning);
  CH_REQUIRE(method == 0 || (method == 1 && edges_x && edges_y),
             "ch_screen_image: method must be 0 (cloud-in-cell) or 1 (histogram, with bin edges)");
  const int nx = resolution_x / binning, ny = resolution_y / binning;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t elem = dtype == CH_F32 ? 4 : 8;
  CH_CUDA(cudaMemsetAsync(image, 0, elem * nx * ny * n_beams, s));
  dim3 grid(static_cast<unsigned>((n_particles + 1023) / 1024), static_cast<unsigned>(n_beams));
  auto launch = [&](auto zero) {
    using T = decltype(zero);
    ch::ScreenArgs<T> a;
    a.particles = static_cast<const T*>(particles);
    a.charges = static_cast<const T*>(charges);
    a.survival = static_cast<const T*>(survival);
    a.misalignment = static_cast<const T*>(misalignment);
    a.pixel_size = static_cast<const T*>(pixel_size);
    a.edges_x = static_cast<const T*>(edges_x);
    a.edges_y = static_cast<const T*>(edges_y);
    a.image = static_cast<T*>(image);
    a.particle_stride = particle_stride;
    a.charge_stride = charge_stride;
    a.survival_stride = survival_stride;
    a.misalignment_stride = misalignment_stride;
    a.n_particles = n_particles;
    a.resolution_x = resolution_x;
    a.resolution_y = resolution_y;
    a.nx = nx;
    a.ny = ny;
    a.method = method;
    a.bulk = ch::bulk_compatible<T>(particles, n_particles, particle_stride) ? 1 : 0;
    const int smem = 1024 * 7 * sizeof(T);
    cudaFuncSetAttribute(ch::screen_image_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         smem);
    ch::screen_image_kernel<T><<<grid, 256, smem, s>>>(a);
  };
  if (dtype == CH_F32) launch(0.0f); else launch(0.0);
  CH_LAUNCH_CHECK();
  return CH_OK;
}

extern "C" int ch_screen_kde(const void* particles, int64_t particle_stride, const void* charges,
                             int64_t charge_stride, const void* survival, int64_t survival_stride,
                             const void* misalignment, int64_t misalignment_stride,
                             const void* centers_x, int32_t nx, const void* centers_y, int32_t ny,
                             const void* bandwidth, int64_t n_particles, int64_t n_beams,
                             int32_t dtype, void* image, double* totals, void* stream) {
  CH_REQUIRE(dtype == CH_F32 || dtype == CH_F64, "ch_screen_kde: bad dtype %d", dtype);
  CH_REQUIRE(particles && charges && misalignment && centers_x && centers_y && bandwidth &&
                 image && totals,
             "ch_screen_kde: NULL pointer argument");
  CH_REQUIRE(n_particles > 0 && n_beams > 0 && n_beams <= 65535,
             "ch_screen_kde: need particles and 1..65535 beams");
  CH_REQUIRE(nx > 0 && ny > 0, "ch_screen_kde: bad image size (%d, %d)", nx, ny);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t elem = dtype == CH_F32 ? 4 : 8;
  const int64_t pixels = static_cast<int64_t>(nx) * ny;
  CH_CUDA(cudaMemsetAsync(image, 0, elem * pixels * n_beams, s));
  CH_CUDA(cudaMemsetAsync(totals, 0, sizeof(double) * n_beams, s));
  dim3 grid(static_cast<unsigned>((n_particles + 1023) / 1024), static_cast<unsigned>(n_beams));
  auto launch = [&](auto zero) {
    using T = decltype(zero);
    ch::KdeArgs<T> a;
    a.particles = static_cast<const T*>(particles);
    a.charges = static_cast<const T*>(charges);
    a.survival = static_cast<const T*>(survival);
    a.misalignment = static_cast<const T*>(misalignment);
    a.centers_x = static_cast<const T*>(centers_x);
    a.centers_y = static_cast<const T*>(centers_y);
    a.bandwidth = static_cast<const T*>(bandwidth);
    a.image = static_cast<T*>(image);
    a.totals = totals;
    a.particle_stride = particle_stride;
    a.charge_stride = charge_stride;
    a.survival_stride = survival_stride;
    a.misalignment_stride = misalignment_stride;
    a.n_particles = n_particles;
    a.nx = nx;
    a.ny = ny;
    a.bulk = ch::bulk_compatible<T>(particles, n_particles, particle_stride) ? 1 : 0;
    const int smem = 1024 * 7 * sizeof(T);
    cudaFuncSetAttribute(ch::screen_kde_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         smem);
    ch::screen_kde_kernel<T><<<grid, 256, smem, s>>>(a);
    dim3 ngrid(static_cast<unsigned>(std::min<int64_t>((pixels + 255) / 256, 1184)),
               static_cast<unsigned>(n_beams));
    ch::screen_kde_normalise_kernel<T><<<ngrid, 256, 0, s>>>(static_cast<T*>(image), totals, pixels);
  };
  if (dtype == CH_F32) launch(0.0f); else launch(0.0);
  CH_LAUNCH_CHECK();
  return CH_OK;
}

extern "C" int ch_screen_gaussian(const void* mu, const void* cov, const void* misalignment,
                                  double left, double step_x, int32_t nx, double bottom,
                                  double step_y, int32_t ny, int32_t dtype, void* image,
                                  void* stream) {
  CH_REQUIRE(dtype == CH_F32 || dtype == CH_F64, "ch_screen_gaussian: bad dtype %d", dtype);
  CH_REQUIRE(mu && cov && misalignment && image, "ch_screen_gaussian: NULL pointer argument");
  CH_REQUIRE(nx > 0 && ny > 0, "ch_screen_gaussian: bad image size (%d, %d)", nx, ny);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int64_t pixels = static_cast<int64_t>(nx) * ny;
  const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((pixels + 255) / 256, 1184));
  if (dtype == CH_F32)
    ch::screen_gaussian_kernel<float><<<blocks, 256, 0, s>>>(
        static_cast<const float*>(mu), static_cast<const float*>(cov),
        static_cast<const float*>(misalignment), left, step_x, nx, bottom, step_y, ny,
        static_cast<float*>(image));
  else
    ch::screen_gaussian_kernel<double><<<blocks, 256, 0, s>>>(
        static_cast<const double*>(mu), static_cast<const double*>(cov),
        static_cast<const double*>(misalignment), left, step_x, nx, bottom, step_y, ny,
        static_cast<double*>(image));
  CH_LAUNCH_CHECK();
  return CH_OK;
}

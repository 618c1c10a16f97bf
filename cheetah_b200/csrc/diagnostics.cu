// Beam diagnostics on the output side of the hot path: the image an active
// Screen records, as one pass over the particles.
//
// Replaces Screen.reading for a ParticleBeam (cheetah/accelerator/screen.py:241-344): the
// misalignment shift of the read beam (:199-215), the charge weights |q| * survival, the 2-D
// cloud-in-cell deposit (cheetah/utils/cloud_in_cell.py:129-241) or the torch.histogramdd call
// with the screen's pixel bin edges (:296-315), and the final transpose to (height, width).
// Each CTA streams a tile of particles (TMA bulk copy) and issues float atomics into the image,
// which (<= 20 MB at the largest screens) lives in L2.
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

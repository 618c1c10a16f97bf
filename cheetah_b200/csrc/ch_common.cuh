// Shared helpers for the cheetah_b200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "cheetah_b200.h"

namespace ch {

// ---- error channel (thread local, never throws across the ABI) -----------------------
void set_error(const char* fmt, ...);
void count_launch(int n = 1);
// opt a kernel in to `bytes` of dynamic shared memory (cached per kernel and device)
cudaError_t allow_dynamic_smem(const void* kernel, int bytes);

#define CH_REQUIRE(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      ch::set_error(__VA_ARGS__);      \
      return CH_EINVAL;                \
    }                                  \
  } while (0)

#define CH_CUDA(expr)                                                                  \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ch::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__,  \
                    __LINE__);                                                         \
      return CH_ECUDA;                                                                 \
    }                                                                                  \
  } while (0)

#define CH_LAUNCH_CHECK()                                                              \
  do {                                                                                 \
    cudaError_t _e = cudaGetLastError();                                               \
    if (_e != cudaSuccess) {                                                           \
      ch::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),        \
                    __FILE__, __LINE__);                                               \
      return CH_ECUDA;                                                                 \
    }                                                                                  \
    ch::count_launch();                                                                \
  } while (0)

// Launch `kernel` so that it may overlap the tail of the previous kernel in `stream`
// (programmatic stream serialization); the kernel calls grid_dependency_wait() before it touches
// anything that kernel wrote.
template <typename... KernelArgs, typename... Args>
inline cudaError_t launch_dependent(void (*kernel)(KernelArgs...), dim3 grid, dim3 block,
                                    size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t config = {};
  config.gridDim = grid;
  config.blockDim = block;
  config.dynamicSmemBytes = smem;
  config.stream = stream;
  cudaLaunchAttribute attribute[1];
  attribute[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attribute[0].val.programmaticStreamSerializationAllowed = 1;
  static const bool disabled = std::getenv("CH_NO_DEPENDENT_LAUNCH") != nullptr;  // A/B timing
  config.attrs = attribute;
  config.numAttrs = disabled ? 0 : 1;
  return cudaLaunchKernelEx(&config, kernel, static_cast<KernelArgs>(args)...);
}

// ---- a scalar read through a (pointer, stride, dtype) triple -------------------------
struct ScalarRef {
  const void* ptr;
  int64_t stride;
  int32_t dtype;
};

__device__ __forceinline__ double load_scalar(const void* ptr, int64_t index, int32_t dtype) {
  return dtype == CH_F64 ? static_cast<const double*>(ptr)[index]
                         : static_cast<double>(static_cast<const float*>(ptr)[index]);
}

// ---- PTX wrappers: mbarrier + 1-D bulk async copies (TMA engine, SASS UBLKCP) ---------
__device__ __forceinline__ uint32_t smem_addr(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_addr(bar)),
      "r"(phase)
      : "memory");
}

// global -> shared bulk copy, completion signalled on an mbarrier (bytes % 16 == 0,
// both addresses 16-byte aligned)
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, uint32_t bytes,
                                          uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_addr(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_addr(bar))
      : "memory");
}

// shared -> global bulk copy tracked by the per-thread bulk async-group
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
               "r"(smem_addr(smem_src)), "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void bulk_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}

template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// Programmatic dependent launch (griddepcontrol): a kernel launched with
// cudaLaunchAttributeProgrammaticStreamSerialization may begin before the previous kernel of the
// stream has finished; everything it reads from that kernel must come after this wait (a no-op
// for a normal launch).
__device__ __forceinline__ void grid_dependency_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}

// make generic-proxy shared-memory writes visible to the async proxy (TMA engine)
__device__ __forceinline__ void fence_async_shared() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- cooperative tile movers used by the particle-streaming kernels --------------------
// Load `n_elems` contiguous elements into shared memory with the whole CTA.  With
// `bulk` (16-byte aligned source and size) one elected thread issues a TMA bulk copy and
// everybody waits on the mbarrier; otherwise plain coalesced loads.  On return every
// thread may read the tile.  `bar` must have been initialised with count 1.
template <typename T>
__device__ __forceinline__ void cta_load_tile(T* smem_dst, const T* gmem_src, int n_elems,
                                              bool bulk, uint64_t* bar, uint32_t& phase) {
  if (bulk) {
    if (threadIdx.x == 0) {
      const uint32_t bytes = static_cast<uint32_t>(n_elems) * sizeof(T);
      mbar_expect_tx(bar, bytes);
      bulk_load(smem_dst, gmem_src, bytes, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1u;
  } else {
    for (int i = threadIdx.x; i < n_elems; i += blockDim.x) smem_dst[i] = gmem_src[i];
    __syncthreads();
  }
}

template <typename T>
__host__ __device__ inline bool bulk_compatible(const void* base, int64_t n_particles,
                                                int64_t batch_stride_elems) {
  return reinterpret_cast<uintptr_t>(base) % 16 == 0 &&
         (static_cast<size_t>(n_particles) * 7 * sizeof(T)) % 16 == 0 &&
         (static_cast<size_t>(batch_stride_elems) * sizeof(T)) % 16 == 0;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace ch

// opaque program object (device-resident copy of the lowered lattice)
struct ch_program {
  int32_t n_ops;
  int32_t n_slots;
  int32_t* opcodes;      // device [n_ops]
  int32_t* op_flags;     // device [n_ops]
  int32_t* slot_begin;   // device [n_ops + 1]
  ch::ScalarRef* slots;  // device [n_slots]
  int32_t* opcodes_host; // host copy of opcodes (argument checks, constant-block sizes)
};

// Library plumbing: error channel, launch counter and the device-resident lattice program.
#include <atomic>
#include <mutex>
#include <unordered_map>
#include <vector>

#include "ch_common.cuh"

namespace ch {

static thread_local char g_error[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) costs a driver call; a kernel keeps the
// attribute per device, so it is only raised when a launch needs more than it was given before.
cudaError_t allow_dynamic_smem(const void* kernel, int bytes) {
  static std::mutex guard;
  static std::unordered_map<uint64_t, int> granted;
  int device = 0;
  cudaError_t err = cudaGetDevice(&device);
  if (err != cudaSuccess) return err;
  const uint64_t key = reinterpret_cast<uint64_t>(kernel) ^ (static_cast<uint64_t>(device) << 56);
  std::lock_guard<std::mutex> lock(guard);
  auto it = granted.find(key);
  if (it != granted.end() && it->second >= bytes) return cudaSuccess;
  err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (err == cudaSuccess) granted[key] = bytes;
  return err;
}

}  // namespace ch

extern "C" int ch_abi_version(void) { return CH_ABI_VERSION; }
extern "C" const char* ch_last_error(void) { return ch::g_error; }
extern "C" int64_t ch_kernel_launch_count(void) {
  return ch::g_launches.load(std::memory_order_relaxed);
}

extern "C" int ch_program_create(const int32_t* opcodes_host, const int32_t* op_flags_host,
                                 const int32_t* slot_begin_host, int32_t n_ops,
                                 const void* const* slot_ptrs_host,
                                 const int64_t* slot_strides_host,
                                 const int32_t* slot_dtypes_host, int32_t n_slots, void* stream,
                                 ch_program** program_out) {
  CH_REQUIRE(program_out != nullptr, "ch_program_create: program_out is NULL");
  *program_out = nullptr;
  CH_REQUIRE(n_ops >= 0 && n_slots >= 0, "ch_program_create: negative size");
  CH_REQUIRE(n_ops == 0 || (opcodes_host && op_flags_host), "ch_program_create: NULL op arrays");
  CH_REQUIRE(slot_begin_host != nullptr, "ch_program_create: NULL slot_begin");
  CH_REQUIRE(n_slots == 0 || (slot_ptrs_host && slot_strides_host && slot_dtypes_host),
             "ch_program_create: NULL slot arrays");
  CH_REQUIRE(slot_begin_host[0] == 0 && slot_begin_host[n_ops] == n_slots,
             "ch_program_create: slot_begin must run from 0 to n_slots");
  for (int32_t i = 0; i < n_ops; ++i) {
    CH_REQUIRE(slot_begin_host[i] <= slot_begin_host[i + 1],
               "ch_program_create: slot_begin not monotonic at op %d", i);
    CH_REQUIRE(opcodes_host[i] >= CH_OP_IDENTITY && opcodes_host[i] <= CH_OP_SECOND_ORDER,
               "ch_program_create: unknown opcode %d at op %d", opcodes_host[i], i);
    static const int kMinSlots[] = {0, 1, 1, 5, 9, 4, 4, 1, 1, 2, 5, 1, 5, 9, 7, 12};
    CH_REQUIRE(slot_begin_host[i + 1] - slot_begin_host[i] >= kMinSlots[opcodes_host[i]],
               "ch_program_create: op %d (opcode %d) has too few slots", i, opcodes_host[i]);
  }
  std::vector<ch::ScalarRef> slots(static_cast<size_t>(n_slots));
  for (int32_t s = 0; s < n_slots; ++s) {
    CH_REQUIRE(slot_ptrs_host[s] != nullptr, "ch_program_create: slot %d has a NULL pointer", s);
    CH_REQUIRE(slot_dtypes_host[s] == CH_F32 || slot_dtypes_host[s] == CH_F64,
               "ch_program_create: slot %d has bad dtype %d", s, slot_dtypes_host[s]);
    slots[s] = ch::ScalarRef{slot_ptrs_host[s], slot_strides_host[s], slot_dtypes_host[s]};
  }

  ch_program* prog = new (std::nothrow) ch_program();
  if (!prog) {
    ch::set_error("ch_program_create: out of host memory");
    return CH_ENOMEM;
  }
  prog->n_ops = n_ops;
  prog->n_slots = n_slots;
  prog->opcodes = prog->op_flags = prog->slot_begin = nullptr;
  prog->slots = nullptr;
  prog->opcodes_host = new (std::nothrow) int32_t[n_ops > 0 ? n_ops : 1];
  if (!prog->opcodes_host) {
    delete prog;
    ch::set_error("ch_program_create: out of host memory");
    return CH_ENOMEM;
  }
  for (int32_t i = 0; i < n_ops; ++i) prog->opcodes_host[i] = opcodes_host[i];

  // one device allocation for the four tables
  const size_t ints = static_cast<size_t>(n_ops) * 2 + static_cast<size_t>(n_ops) + 1;
  const size_t ints_bytes = (ints * sizeof(int32_t) + 15) / 16 * 16;
  const size_t bytes = ints_bytes + slots.size() * sizeof(ch::ScalarRef);
  std::vector<unsigned char> staging(bytes, 0);
  int32_t* h_ints = reinterpret_cast<int32_t*>(staging.data());
  for (int32_t i = 0; i < n_ops; ++i) {
    h_ints[i] = opcodes_host[i];
    h_ints[n_ops + i] = op_flags_host[i];
  }
  for (int32_t i = 0; i <= n_ops; ++i) h_ints[2 * n_ops + i] = slot_begin_host[i];
  if (!slots.empty())
    memcpy(staging.data() + ints_bytes, slots.data(), slots.size() * sizeof(ch::ScalarRef));

  unsigned char* device = nullptr;
  cudaError_t err = cudaMalloc(&device, bytes);
  if (err != cudaSuccess) {
    delete[] prog->opcodes_host;
    delete prog;
    ch::set_error("ch_program_create: cudaMalloc(%zu) failed: %s", bytes,
                  cudaGetErrorString(err));
    return CH_ENOMEM;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  err = cudaMemcpyAsync(device, staging.data(), bytes, cudaMemcpyHostToDevice, s);
  // the staging vector is pageable host memory: the copy has been staged by the driver
  // when cudaMemcpyAsync returns, so it may be freed at scope exit
  if (err != cudaSuccess) {
    cudaFree(device);
    delete[] prog->opcodes_host;
    delete prog;
    ch::set_error("ch_program_create: upload failed: %s", cudaGetErrorString(err));
    return CH_ECUDA;
  }
  prog->opcodes = reinterpret_cast<int32_t*>(device);
  prog->op_flags = prog->opcodes + n_ops;
  prog->slot_begin = prog->op_flags + n_ops;
  prog->slots = reinterpret_cast<ch::ScalarRef*>(device + ints_bytes);
  *program_out = prog;
  return CH_OK;
}

extern "C" int ch_program_destroy(ch_program* program) {
  if (!program) return CH_OK;
  // freed with cudaFree: implicitly waits for kernels still using the tables
  cudaError_t err = cudaFree(program->opcodes);
  delete[] program->opcodes_host;
  delete program;
  if (err != cudaSuccess) {
    ch::set_error("ch_program_destroy: cudaFree failed: %s", cudaGetErrorString(err));
    return CH_ECUDA;
  }
  return CH_OK;
}

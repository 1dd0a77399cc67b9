// b200bls.cu -- C ABI (include/b200bls.h) over the field-VM kernel.
//
// Host side of the library: owns the CUDA context objects (one device per process), the
// embedded field programs (csrc/gen/programs.bin, linked in as a binary object), staging
// buffers for the host-pointer entry points, and the launch logic.  No CPU compute path
// exists here: every entry point either launches vm_kernel on the GPU or fails.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200bls.h"
#include "vm_kernel.cuh"

using namespace b200bls;

extern "C" {
extern const unsigned char _binary_programs_bin_start[];
extern const unsigned char _binary_programs_bin_end[];
}

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess)                                                                \
      return fail(B200BLS_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), \
                  __FILE__, __LINE__);                                                    \
  } while (0)

struct BlobEntry {
  char name[32];
  uint32_t n_ins, body_start, epi_start, n_consts, n_slots, n_cold;
  uint64_t code_off, consts_off;
};
static_assert(sizeof(BlobEntry) == 32 + 24 + 16, "blob entry layout");

struct DevProgram {
  uint2* code = nullptr;
  uint4* consts = nullptr;
  int n_ins = 0, body_start = 0, epi_start = 0, n_slots = 0, n_cold = 0;
};

struct Staging {
  void* ptr = nullptr;
  size_t cap = 0;
};

struct Context {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::map<std::string, DevProgram> programs;
  uint4* cold = nullptr;
  size_t cold_bytes = 0;
  Staging staging[VM_MAX_BUFS];
  uint64_t launches = 0;
  int ctas_per_sm = 1;   // 1: 18-slot programs, 2: the "#9" variants, two CTAs per SM
};

Context g_ctx;
std::mutex g_mu;

int ensure_staging(int i, size_t bytes) {
  Staging& s = g_ctx.staging[i];
  if (s.cap >= bytes) return 0;
  if (s.ptr) cudaFree(s.ptr);
  s.ptr = nullptr;
  s.cap = 0;
  size_t cap = bytes + bytes / 4 + 4096;
  cudaError_t e = cudaMalloc(&s.ptr, cap);
  if (e != cudaSuccess) return fail(B200BLS_E_NOMEM, "cudaMalloc(%zu) failed: %s", cap, cudaGetErrorString(e));
  s.cap = cap;
  return 0;
}

// grid: ctas_per_sm CTAs per SM at most; fewer when the batch is small
int launch_program(const DevProgram& pr, size_t n_items, const VmBuf* bufs, int n_bufs, int grid_override = 0) {
  Context& c = g_ctx;
  long long blocks_needed = (long long)((n_items + VM_NT - 1) / VM_NT);
  long long max_grid = (long long)c.sm_count * (pr.n_slots <= 9 ? 2 : 1);
  int grid = (int)(blocks_needed < max_grid ? blocks_needed : max_grid);
  if (grid < 1) grid = 1;
  if (grid_override > 0) grid = grid_override;
  long long total = (long long)grid * VM_NT;
  size_t cold_need = (size_t)pr.n_cold * 6 * sizeof(uint4) * total;
  if (cold_need > c.cold_bytes) {
    if (c.cold) cudaFree(c.cold);
    c.cold = nullptr;
    c.cold_bytes = 0;
    size_t want = (size_t)pr.n_cold * 6 * sizeof(uint4) * (size_t)max_grid * VM_NT;
    if (want < cold_need) want = cold_need;
    cudaError_t e = cudaMalloc(&c.cold, want);
    if (e != cudaSuccess) return fail(B200BLS_E_NOMEM, "cold area cudaMalloc(%zu) failed", want);
    c.cold_bytes = want;
  }
  VmParams p;
  memset(&p, 0, sizeof(p));
  p.code = pr.code;
  p.body_start = pr.body_start;
  p.epi_start = pr.epi_start;
  p.n_ins = pr.n_ins;
  p.consts = pr.consts;
  p.cold = c.cold;
  p.n_items = (long long)n_items;
  p.iters = (long long)((n_items + total - 1) / total);
  for (int i = 0; i < n_bufs && i < VM_MAX_BUFS; i++) p.bufs[i] = bufs[i];
  size_t smem = (size_t)pr.n_slots * 2 * 3 * sizeof(uint4) * VM_NT;
  vm_kernel<<<grid, VM_NT, smem, c.stream>>>(p);
  CU(cudaGetLastError());
  c.launches++;
  return 0;
}

const DevProgram* find_program(const char* base) {
  std::string name(base);
  if (g_ctx.ctas_per_sm == 2 && name.find('#') == std::string::npos) name += "#9";
  auto it = g_ctx.programs.find(name);
  if (it == g_ctx.programs.end()) {
    fail(B200BLS_E_PROGRAM, "unknown program '%s'", name.c_str());
    return nullptr;
  }
  return &it->second;
}

struct HostBuf {
  const void* in;   // host source (nullptr for pure outputs)
  void* out;        // host destination (nullptr for pure inputs)
  size_t stride;    // bytes per item
};

// copy inputs up, run, copy outputs back, synchronise
int run_host(const char* name, size_t n, const HostBuf* hb, int n_bufs) {
  if (!g_ctx.ready) return fail(B200BLS_E_NOT_INIT, "b200bls_init() has not succeeded");
  const DevProgram* pr = find_program(name);
  if (!pr) return B200BLS_E_PROGRAM;
  if (n == 0) return 0;
  VmBuf bufs[VM_MAX_BUFS];
  memset(bufs, 0, sizeof(bufs));
  for (int i = 0; i < n_bufs; i++) {
    int rc = ensure_staging(i, hb[i].stride * n);
    if (rc) return rc;
    bufs[i].ptr = (unsigned char*)g_ctx.staging[i].ptr;
    bufs[i].stride = (long long)hb[i].stride;
    if (hb[i].in) CU(cudaMemcpyAsync(bufs[i].ptr, hb[i].in, hb[i].stride * n, cudaMemcpyHostToDevice, g_ctx.stream));
  }
  int rc = launch_program(*pr, n, bufs, n_bufs);
  if (rc) return rc;
  for (int i = 0; i < n_bufs; i++)
    if (hb[i].out) CU(cudaMemcpyAsync(hb[i].out, bufs[i].ptr, hb[i].stride * n, cudaMemcpyDeviceToHost, g_ctx.stream));
  CU(cudaStreamSynchronize(g_ctx.stream));
  return 0;
}

struct DevBuf {
  const void* ptr;
  size_t stride;
};

int run_dev(const char* name, size_t n, const DevBuf* db, int n_bufs) {
  if (!g_ctx.ready) return fail(B200BLS_E_NOT_INIT, "b200bls_init() has not succeeded");
  const DevProgram* pr = find_program(name);
  if (!pr) return B200BLS_E_PROGRAM;
  if (n == 0) return 0;
  VmBuf bufs[VM_MAX_BUFS];
  memset(bufs, 0, sizeof(bufs));
  for (int i = 0; i < n_bufs; i++) {
    bufs[i].ptr = (unsigned char*)db[i].ptr;
    bufs[i].stride = (long long)db[i].stride;
  }
  return launch_program(*pr, n, bufs, n_bufs);
}

const char* kFieldOps[] = {"add", "sub", "mul", "sqr", "neg", "inv"};

int field_prog_name(int level, int op, char* out, size_t cap) {
  if (!(level == 1 || level == 2 || level == 6 || level == 12) || op < 0 || op > 5)
    return fail(B200BLS_E_ARG, "field_op: bad level %d or op %d", level, op);
  snprintf(out, cap, "f%d_%s", level, kFieldOps[op]);
  return 0;
}

}  // namespace

extern "C" {

const char* b200bls_last_error(void) { return g_err; }

int b200bls_init(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  Context& c = g_ctx;
  if (c.ready) {
    if (c.device == device) return 0;
    return fail(B200BLS_E_ARG, "already initialised on device %d", c.device);
  }
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(B200BLS_E_CUDA, "no CUDA device available (%s); this library has no CPU path",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= n_dev) return fail(B200BLS_E_ARG, "device %d out of range (0..%d)", device, n_dev - 1);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  c.sm_count = prop.multiProcessorCount;
  CU(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  CU(cudaEventCreate(&c.ev0));
  CU(cudaEventCreate(&c.ev1));
  CU(cudaFuncSetAttribute(vm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  // parse the embedded program blob
  const unsigned char* blob = _binary_programs_bin_start;
  size_t blob_len = (size_t)(_binary_programs_bin_end - _binary_programs_bin_start);
  if (blob_len < 16 || memcmp(blob, "B2BLSPRG", 8) != 0) return fail(B200BLS_E_PROGRAM, "bad program blob");
  uint32_t n_prog;
  memcpy(&n_prog, blob + 12, 4);
  for (uint32_t i = 0; i < n_prog; i++) {
    BlobEntry en;
    memcpy(&en, blob + 16 + i * sizeof(BlobEntry), sizeof(en));
    DevProgram dp;
    dp.n_ins = (int)en.n_ins;
    dp.body_start = (int)en.body_start;
    dp.epi_start = (int)en.epi_start;
    dp.n_slots = (int)en.n_slots;
    dp.n_cold = (int)en.n_cold;
    size_t code_bytes = (size_t)(en.n_ins + 1) * sizeof(uint2);
    size_t const_bytes = (size_t)en.n_consts * 3 * sizeof(uint4);
    CU(cudaMalloc(&dp.code, code_bytes));
    CU(cudaMalloc(&dp.consts, const_bytes));
    CU(cudaMemcpy(dp.code, blob + en.code_off, code_bytes, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dp.consts, blob + en.consts_off, const_bytes, cudaMemcpyHostToDevice));
    char nm[33];
    memcpy(nm, en.name, 32);
    nm[32] = 0;
    c.programs[nm] = dp;
  }
  const char* env = getenv("B200BLS_CTAS_PER_SM");
  if (env && (env[0] == '1' || env[0] == '2')) c.ctas_per_sm = env[0] - '0';
  c.device = device;
  c.ready = true;
  return 0;
}

void b200bls_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  Context& c = g_ctx;
  if (!c.ready) return;
  cudaStreamSynchronize(c.stream);
  for (auto& kv : c.programs) {
    cudaFree(kv.second.code);
    cudaFree(kv.second.consts);
  }
  c.programs.clear();
  for (auto& s : c.staging) {
    if (s.ptr) cudaFree(s.ptr);
    s.ptr = nullptr;
    s.cap = 0;
  }
  if (c.cold) cudaFree(c.cold);
  c.cold = nullptr;
  c.cold_bytes = 0;
  cudaEventDestroy(c.ev0);
  cudaEventDestroy(c.ev1);
  cudaStreamDestroy(c.stream);
  c.ready = false;
  c.device = -1;
}

int b200bls_sm_count(void) { return g_ctx.ready ? g_ctx.sm_count : 0; }

int b200bls_set_ctas_per_sm(int n) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (n != 1 && n != 2) return fail(B200BLS_E_ARG, "ctas_per_sm must be 1 or 2");
  g_ctx.ctas_per_sm = n;
  return 0;
}

int b200bls_get_ctas_per_sm(void) { return g_ctx.ctas_per_sm; }

int b200bls_sync(void) {
  if (!g_ctx.ready) return fail(B200BLS_E_NOT_INIT, "not initialised");
  CU(cudaStreamSynchronize(g_ctx.stream));
  return 0;
}

void* b200bls_malloc(size_t bytes) {
  if (!g_ctx.ready) {
    fail(B200BLS_E_NOT_INIT, "not initialised");
    return nullptr;
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
  if (e != cudaSuccess) {
    fail(B200BLS_E_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  return p;
}

void b200bls_free(void* p) {
  if (p) cudaFree(p);
}

void* b200bls_host_alloc(size_t bytes) {
  if (!g_ctx.ready) {
    fail(B200BLS_E_NOT_INIT, "not initialised");
    return nullptr;
  }
  void* p = nullptr;
  cudaError_t e = cudaMallocHost(&p, bytes ? bytes : 1);
  if (e != cudaSuccess) {
    fail(B200BLS_E_NOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e));
    return nullptr;
  }
  return p;
}

void b200bls_host_free(void* p) {
  if (p) cudaFreeHost(p);
}

int b200bls_h2d(void* dst, const void* src, size_t bytes) {
  if (!g_ctx.ready) return fail(B200BLS_E_NOT_INIT, "not initialised");
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, g_ctx.stream));
  return 0;
}

int b200bls_d2h(void* dst, const void* src, size_t bytes) {
  if (!g_ctx.ready) return fail(B200BLS_E_NOT_INIT, "not initialised");
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, g_ctx.stream));
  return 0;
}

int b200bls_timer_start(void) {
  if (!g_ctx.ready) return fail(B200BLS_E_NOT_INIT, "not initialised");
  CU(cudaEventRecord(g_ctx.ev0, g_ctx.stream));
  return 0;
}

int b200bls_timer_stop(float* ms) {
  if (!g_ctx.ready) return fail(B200BLS_E_NOT_INIT, "not initialised");
  CU(cudaEventRecord(g_ctx.ev1, g_ctx.stream));
  CU(cudaEventSynchronize(g_ctx.ev1));
  CU(cudaEventElapsedTime(ms, g_ctx.ev0, g_ctx.ev1));
  return 0;
}

uint64_t b200bls_launch_count(void) { return g_ctx.launches; }

int b200bls_run_program_dev(const char* name, size_t n_items, void* const* bufs, const int64_t* strides, int n_bufs) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (n_bufs < 0 || n_bufs > VM_MAX_BUFS) return fail(B200BLS_E_ARG, "n_bufs out of range");
  DevBuf db[VM_MAX_BUFS];
  for (int i = 0; i < n_bufs; i++) {
    db[i].ptr = bufs[i];
    db[i].stride = (size_t)strides[i];
  }
  return run_dev(name, n_items, db, n_bufs);
}

int b200bls_program_info(const char* name, int* n_ins, int* n_slots, int* n_cold) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g_ctx.ready) return fail(B200BLS_E_NOT_INIT, "not initialised");
  const DevProgram* pr = find_program(name);
  if (!pr) return B200BLS_E_PROGRAM;
  if (n_ins) *n_ins = pr->n_ins;
  if (n_slots) *n_slots = pr->n_slots;
  if (n_cold) *n_cold = pr->n_cold;
  return 0;
}

int b200bls_field_op_batch(int level, int op, const uint8_t* a, const uint8_t* b, uint8_t* out, size_t n) {
  std::lock_guard<std::mutex> lk(g_mu);
  char name[32];
  int rc = field_prog_name(level, op, name, sizeof(name));
  if (rc) return rc;
  size_t w = 48 * (size_t)level;
  HostBuf hb[3] = {{a, nullptr, w}, {op <= 2 ? b : nullptr, nullptr, w}, {nullptr, out, w}};
  if (!a || !out || (op <= 2 && !b)) return fail(B200BLS_E_ARG, "null buffer");
  return run_host(name, n, hb, 3);
}

int b200bls_field_op_batch_dev(int level, int op, const void* a, const void* b, void* out, size_t n) {
  std::lock_guard<std::mutex> lk(g_mu);
  char name[32];
  int rc = field_prog_name(level, op, name, sizeof(name));
  if (rc) return rc;
  size_t w = 48 * (size_t)level;
  DevBuf db[3] = {{a, w}, {b ? b : a, w}, {out, w}};
  return run_dev(name, n, db, 3);
}

int b200bls_pairing_batch(const uint8_t* P, const uint8_t* Q, uint8_t* out, size_t n) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!P || !Q || !out) return fail(B200BLS_E_ARG, "null buffer");
  HostBuf hb[3] = {{P, nullptr, 96}, {Q, nullptr, 192}, {nullptr, out, 576}};
  return run_host("pairing", n, hb, 3);
}

int b200bls_pairing_batch_dev(const void* P, const void* Q, void* out, size_t n) {
  std::lock_guard<std::mutex> lk(g_mu);
  DevBuf db[3] = {{P, 96}, {Q, 192}, {out, 576}};
  return run_dev("pairing", n, db, 3);
}

int b200bls_final_exp_batch(const uint8_t* in, uint8_t* out, size_t n) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!in || !out) return fail(B200BLS_E_ARG, "null buffer");
  HostBuf hb[2] = {{in, nullptr, 576}, {nullptr, out, 576}};
  return run_host("final_exp", n, hb, 2);
}

int b200bls_final_exp_batch_dev(const void* in, void* out, size_t n) {
  std::lock_guard<std::mutex> lk(g_mu);
  DevBuf db[2] = {{in, 576}, {out, 576}};
  return run_dev("final_exp", n, db, 2);
}

int b200bls_miller_loop_batch(const uint8_t* P, const uint8_t* Q, uint8_t* out, size_t n) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!P || !Q || !out) return fail(B200BLS_E_ARG, "null buffer");
  HostBuf hb[3] = {{P, nullptr, 96}, {Q, nullptr, 192}, {nullptr, out, 576}};
  return run_host("miller_loop", n, hb, 3);
}

int b200bls_miller_loop_batch_dev(const void* P, const void* Q, void* out, size_t n) {
  std::lock_guard<std::mutex> lk(g_mu);
  DevBuf db[3] = {{P, 96}, {Q, 192}, {out, 576}};
  return run_dev("miller_loop", n, db, 3);
}

}  // extern "C"

// qibo_b200: C-ABI entry points (include/qibo_b200.h).  Host-side dispatch only; kernels live in *.cuh.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/qibo_b200.h"
#include "qb_canon.hpp"
#include "qb_families.hpp"
#include "qb_common.cuh"
#include "qb_gate_kernels.cuh"
#include "qb_measure_kernels.cuh"
#include "qb_planner.hpp"
#include "qb_sweep.cuh"
#include "qb_permute.cuh"

using namespace qb;

// ---------------------------------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------------------------------
struct qb_context {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  std::mutex mu;  // one backend object may be entered from several joblib threads (parallel.py:53)
  // scratch (grown on demand, stream-ordered reuse)
  void* scratch = nullptr;
  size_t scratch_bytes = 0;
  double* cdf = nullptr;
  size_t cdf_bytes = 0;
  // device + pinned-host program buffers for the sweep kernel
  void* prog_dev = nullptr;
  void* prog_host = nullptr;
  size_t prog_bytes = 0;
  cudaEvent_t prog_done = nullptr;  // last kernel that read prog_dev
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
static int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  cudaGetLastError();  // clear sticky-less errors
  return e == cudaErrorMemoryAllocation ? QB_ERR_OOM : QB_ERR_CUDA;
}
#define QB_CUDA(call)                                   \
  do {                                                  \
    cudaError_t _e = (call);                            \
    if (_e != cudaSuccess) return cuda_fail(_e, #call); \
  } while (0)
#define QB_CHECK_LAUNCH(what)                             \
  do {                                                    \
    cudaError_t _e = cudaGetLastError();                  \
    if (_e != cudaSuccess) return cuda_fail(_e, what);    \
  } while (0)

struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

static int ensure_scratch(qb_context* h, size_t bytes) {
  if (h->scratch_bytes >= bytes) return QB_OK;
  if (h->scratch) {
    QB_CUDA(cudaStreamSynchronize(h->stream));
    cudaFree(h->scratch);
    h->scratch = nullptr;
    h->scratch_bytes = 0;
  }
  size_t want = bytes < (size_t(1) << 20) ? (size_t(1) << 20) : bytes;
  QB_CUDA(cudaMalloc(&h->scratch, want));
  h->scratch_bytes = want;
  return QB_OK;
}

static bool valid_state_args(const void* state, int nqubits, int dtype) {
  return state != nullptr && nqubits >= 1 && nqubits <= QB_MAX_QUBITS && (dtype == QB_C64 || dtype == QB_C128);
}

int qb_version(void) { return QB_VERSION; }
const char* qb_last_error(void) { return g_err.c_str(); }

int qb_device_count(int* count) {
  if (!count) return fail(QB_ERR_INVALID, "null count");
  int c = 0;
  cudaError_t e = cudaGetDeviceCount(&c);
  if (e != cudaSuccess) {
    *count = 0;
    return cuda_fail(e, "cudaGetDeviceCount");
  }
  *count = c;
  return QB_OK;
}

int qb_create(int device, void* cuda_stream, qb_handle* out) {
  if (!out) return fail(QB_ERR_INVALID, "null handle pointer");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(QB_ERR_CUDA, "qibo_b200 needs a CUDA device (no CPU fallback): " +
                                 std::string(e == cudaSuccess ? "no device found" : cudaGetErrorString(e)));
  if (device < 0 || device >= count) return fail(QB_ERR_INVALID, "device index out of range");
  DeviceGuard guard(device);
  qb_context* h = new qb_context();
  h->device = device;
  cudaDeviceProp prop;
  QB_CUDA(cudaGetDeviceProperties(&prop, device));
  h->sm_count = prop.multiProcessorCount;
  if (cuda_stream) {
    h->stream = (cudaStream_t)cuda_stream;
  } else {
    QB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  QB_CUDA(cudaEventCreateWithFlags(&h->prog_done, cudaEventDisableTiming));
  QB_CUDA(cudaEventCreate(&h->ev0));
  QB_CUDA(cudaEventCreate(&h->ev1));
  int rc = sweep_configure(prop);
  if (rc != QB_OK) {
    delete h;
    return fail(rc, "sweep kernel configuration failed (needs sm_100a, 227 KB shared memory per block)");
  }
  *out = h;
  return QB_OK;
}

int qb_destroy(qb_handle h) {
  if (!h) return QB_OK;
  DeviceGuard guard(h->device);
  cudaStreamSynchronize(h->stream);
  if (h->scratch) cudaFree(h->scratch);
  if (h->cdf) cudaFree(h->cdf);
  if (h->prog_dev) cudaFree(h->prog_dev);
  if (h->prog_host) cudaFreeHost(h->prog_host);
  if (h->prog_done) cudaEventDestroy(h->prog_done);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  if (h->own_stream) cudaStreamDestroy(h->stream);
  delete h;
  return QB_OK;
}

int qb_set_stream(qb_handle h, void* cuda_stream) {
  if (!h) return fail(QB_ERR_INVALID, "null handle");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  QB_CUDA(cudaStreamSynchronize(h->stream));
  if (h->own_stream) cudaStreamDestroy(h->stream);
  h->own_stream = false;
  if (cuda_stream) {
    h->stream = (cudaStream_t)cuda_stream;
  } else {
    QB_CUDA(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    h->own_stream = true;
  }
  return QB_OK;
}

int qb_sync(qb_handle h) {
  if (!h) return fail(QB_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  QB_CUDA(cudaStreamSynchronize(h->stream));
  return QB_OK;
}

int qb_mem_info(qb_handle h, size_t* free_bytes, size_t* total_bytes) {
  if (!h || !free_bytes || !total_bytes) return fail(QB_ERR_INVALID, "null argument");
  DeviceGuard guard(h->device);
  QB_CUDA(cudaMemGetInfo(free_bytes, total_bytes));
  return QB_OK;
}

int qb_malloc(qb_handle h, size_t bytes, void** dptr) {
  if (!h || !dptr) return fail(QB_ERR_INVALID, "null argument");
  DeviceGuard guard(h->device);
  QB_CUDA(cudaMalloc(dptr, bytes));
  return QB_OK;
}

int qb_free(qb_handle h, void* dptr) {
  if (!h) return fail(QB_ERR_INVALID, "null handle");
  DeviceGuard guard(h->device);
  QB_CUDA(cudaStreamSynchronize(h->stream));
  QB_CUDA(cudaFree(dptr));
  return QB_OK;
}

int qb_memcpy(qb_handle h, void* dst, const void* src, size_t bytes, int kind) {
  if (!h || (!dst && bytes) || (!src && bytes)) return fail(QB_ERR_INVALID, "null argument");
  DeviceGuard guard(h->device);
  cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  QB_CUDA(cudaMemcpyAsync(dst, src, bytes, k, h->stream));
  if (kind != 2) QB_CUDA(cudaStreamSynchronize(h->stream));
  return QB_OK;
}

int qb_memcpy_async(qb_handle h, void* dst, const void* src, size_t bytes, int kind, void* cuda_stream) {
  if (!h || (!dst && bytes) || (!src && bytes) || kind < 0 || kind > 2) return fail(QB_ERR_INVALID, "bad copy arguments");
  DeviceGuard guard(h->device);
  cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
  QB_CUDA(cudaMemcpyAsync(dst, src, bytes, k, cuda_stream ? (cudaStream_t)cuda_stream : h->stream));
  return QB_OK;
}

// ---------------------------------------------------------------------------------------------------
// K6
// ---------------------------------------------------------------------------------------------------
static int grid_for(uint64_t count, int threads, int sm_count, int waves = 16) {
  uint64_t blocks = (count + threads - 1) / threads;
  uint64_t cap = uint64_t(sm_count) * waves;
  return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

static int set_basis_locked(qb_context* h, void* state, int nqubits, int dtype, uint64_t index);

int qb_state_set_basis(qb_handle h, void* state, int nqubits, int dtype, uint64_t index) {
  if (!h || !valid_state_args(state, nqubits, dtype)) return fail(QB_ERR_INVALID, "bad state arguments");
  if (index >= (uint64_t(1) << nqubits)) return fail(QB_ERR_INVALID, "basis index out of range");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  return set_basis_locked(h, state, nqubits, dtype, index);
}

static int set_basis_locked(qb_context* h, void* state, int nqubits, int dtype, uint64_t index) {
  uint64_t count = uint64_t(1) << nqubits;
  int grid = grid_for(count, 256, h->sm_count);
  if (dtype == QB_C128)
    k6_fill<double2><<<grid, 256, 0, h->stream>>>((double2*)state, count, cmake<double2>(0, 0), index, 1);
  else
    k6_fill<float2><<<grid, 256, 0, h->stream>>>((float2*)state, count, cmake<float2>(0, 0), index, 1);
  QB_CHECK_LAUNCH("k6_fill");
  return QB_OK;
}

int qb_state_fill(qb_handle h, void* state, int nqubits, int dtype, double re, double im) {
  if (!h || !valid_state_args(state, nqubits, dtype)) return fail(QB_ERR_INVALID, "bad state arguments");
  uint64_t count = uint64_t(1) << nqubits;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  int grid = grid_for(count, 256, h->sm_count);
  if (dtype == QB_C128)
    k6_fill<double2><<<grid, 256, 0, h->stream>>>((double2*)state, count, cmake<double2>(re, im), 0, 0);
  else
    k6_fill<float2><<<grid, 256, 0, h->stream>>>((float2*)state, count, cmake<float2>((float)re, (float)im), 0, 0);
  QB_CHECK_LAUNCH("k6_fill");
  return QB_OK;
}

int qb_state_cast(qb_handle h, void* dst, int dst_dtype, const void* src, int src_dtype, uint64_t count) {
  if (!h || !dst || !src) return fail(QB_ERR_INVALID, "null argument");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  int grid = grid_for(count, 256, h->sm_count);
  if (dst_dtype == QB_C128 && src_dtype == QB_C64)
    k6_cast<double2, float2><<<grid, 256, 0, h->stream>>>((double2*)dst, (const float2*)src, count);
  else if (dst_dtype == QB_C64 && src_dtype == QB_C128)
    k6_cast<float2, double2><<<grid, 256, 0, h->stream>>>((float2*)dst, (const double2*)src, count);
  else if (dst_dtype == src_dtype && (dst_dtype == QB_C64 || dst_dtype == QB_C128)) {
    QB_CUDA(cudaMemcpyAsync(dst, src, count * (dst_dtype == QB_C128 ? 16 : 8), cudaMemcpyDeviceToDevice, h->stream));
    return QB_OK;
  } else
    return fail(QB_ERR_INVALID, "bad dtype");
  QB_CHECK_LAUNCH("k6_cast");
  return QB_OK;
}

// sum |amp|^2 over the slice (idx & mask) == val into scratch[0] (device), deterministic
static int slice_norm2_device(qb_context* h, const void* state, int nqubits, int dtype, const std::vector<int>& pos_sorted,
                              uint64_t val, double** result_dev) {
  int m = (int)pos_sorted.size();
  uint64_t ngroups = uint64_t(1) << (nqubits - m);
  int grid = grid_for(ngroups, RED_THREADS, h->sm_count, 8);
  int rc = ensure_scratch(h, (size_t)(grid + 8) * sizeof(double));
  if (rc != QB_OK) return rc;
  double* partial = (double*)h->scratch + 8;
  double* out = (double*)h->scratch;
  InsertList ins;
  ins.n = m;
  for (int i = 0; i < m; ++i) ins.pos[i] = (uint8_t)pos_sorted[i];
  if (dtype == QB_C128)
    k5_slice_norm2<double2><<<grid, RED_THREADS, 0, h->stream>>>((const double2*)state, ngroups, ins, val, partial);
  else
    k5_slice_norm2<float2><<<grid, RED_THREADS, 0, h->stream>>>((const float2*)state, ngroups, ins, val, partial);
  QB_CHECK_LAUNCH("k5_slice_norm2");
  k5_sum_partials<<<1, 32, 0, h->stream>>>(partial, grid, out);
  QB_CHECK_LAUNCH("k5_sum_partials");
  *result_dev = out;
  return QB_OK;
}

int qb_state_norm2(qb_handle h, const void* state, int nqubits, int dtype, double* out_host) {
  if (!h || !valid_state_args(state, nqubits, dtype) || !out_host) return fail(QB_ERR_INVALID, "bad state arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  double* dev = nullptr;
  int rc = slice_norm2_device(h, state, nqubits, dtype, {}, 0, &dev);
  if (rc != QB_OK) return rc;
  QB_CUDA(cudaMemcpyAsync(out_host, dev, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  QB_CUDA(cudaStreamSynchronize(h->stream));
  return QB_OK;
}

// ---------------------------------------------------------------------------------------------------
// K1
// ---------------------------------------------------------------------------------------------------
template <typename C> static C to_c(cd v) {
  return cmake<C>((typename real_of<C>::type)v.real(), (typename real_of<C>::type)v.imag());
}

static void fill_insert_and_offsets(const CanonOp& op, InsertList& ins, uint64_t* off) {
  std::vector<int> all(op.tpos);
  all.insert(all.end(), op.cpos.begin(), op.cpos.end());
  std::sort(all.begin(), all.end());
  ins.n = (int)all.size();
  for (int i = 0; i < ins.n; ++i) ins.pos[i] = (uint8_t)all[i];
  int k = (int)op.tpos.size();
  for (int j = 0; j < (1 << k); ++j) {
    uint64_t o = 0;
    for (int i = 0; i < k; ++i)
      if ((j >> (k - 1 - i)) & 1) o |= uint64_t(1) << op.tpos[i];
    off[j] = o;
  }
}

template <typename C, int K, int UNROLL>
static int launch_dense(qb_context* h, void* state, int nqubits, const CanonOp& op) {
  static DenseParams<C, K> p;  // large (up to 16 KB): keep off the stack; guarded by h->mu per context
  static std::mutex pm;
  std::lock_guard<std::mutex> lk(pm);
  fill_insert_and_offsets(op, p.ins, p.off);
  p.cmask = op.cmask();
  p.ngroups = uint64_t(1) << (nqubits - p.ins.n);
  for (int i = 0; i < (1 << (2 * K)); ++i) p.m[i] = to_c<C>(op.data[i]);
  uint64_t per_block = uint64_t(K1_THREADS) * UNROLL;
  uint64_t grid = (p.ngroups + per_block - 1) / per_block;
  k1_dense<C, K, UNROLL><<<(unsigned)grid, K1_THREADS, 0, h->stream>>>((C*)state, p);
  QB_CHECK_LAUNCH("k1_dense");
  return QB_OK;
}

template <typename C, int K, int UNROLL>
static int launch_diag(qb_context* h, void* state, int nqubits, const CanonOp& op) {
  DiagParams<C, K> p;
  fill_insert_and_offsets(op, p.ins, p.off);
  p.cmask = op.cmask();
  p.ngroups = uint64_t(1) << (nqubits - p.ins.n);
  for (int i = 0; i < (1 << K); ++i) p.d[i] = to_c<C>(op.data[i]);
  uint64_t per_block = uint64_t(K1_THREADS) * UNROLL;
  uint64_t grid = (p.ngroups + per_block - 1) / per_block;
  k1_diag<C, K, UNROLL><<<(unsigned)grid, K1_THREADS, 0, h->stream>>>((C*)state, p);
  QB_CHECK_LAUNCH("k1_diag");
  return QB_OK;
}

template <typename C> static int launch_slice(qb_context* h, void* state, int nqubits, const CanonOp& op) {
  SliceParams<C> p;
  uint64_t off[4] = {0, 0, 0, 0};
  fill_insert_and_offsets(op, p.ins, off);
  p.cmask = op.cmask();
  p.ngroups = uint64_t(1) << (nqubits - p.ins.n);
  constexpr int UNROLL = 4;
  uint64_t per_block = uint64_t(K1_THREADS) * UNROLL;
  uint64_t grid = (p.ngroups + per_block - 1) / per_block;
  if (op.kind == CK_PHASE) {
    p.phase = to_c<C>(op.data[0]);
    p.off01 = p.off10 = 0;
    k1_phase<C, UNROLL><<<(unsigned)grid, K1_THREADS, 0, h->stream>>>((C*)state, p);
  } else {
    p.phase = cmake<C>(1, 0);
    p.off01 = off[1];
    p.off10 = off[2];
    k1_swap<C, UNROLL><<<(unsigned)grid, K1_THREADS, 0, h->stream>>>((C*)state, p);
  }
  QB_CHECK_LAUNCH("k1_slice");
  return QB_OK;
}

template <typename C> static int apply_canon_k1(qb_context* h, void* state, int nqubits, const CanonOp& op) {
  int k = (int)op.tpos.size();
  switch (op.kind) {
    case CK_NOOP:
      return QB_OK;
    case CK_PHASE:
    case CK_SWAP:
      return launch_slice<C>(h, state, nqubits, op);
    case CK_DENSE:
      switch (k) {
        case 1: return launch_dense<C, 1, 4>(h, state, nqubits, op);
        case 2: return launch_dense<C, 2, 2>(h, state, nqubits, op);
        case 3: return launch_dense<C, 3, 1>(h, state, nqubits, op);
        case 4: return launch_dense<C, 4, 1>(h, state, nqubits, op);
        case 5: return launch_dense<C, 5, 1>(h, state, nqubits, op);
        default: return fail(QB_ERR_UNSUPPORTED, "dense gates on more than 5 target qubits are not supported");
      }
    case CK_DIAG:
      switch (k) {
        case 1: return launch_diag<C, 1, 4>(h, state, nqubits, op);
        case 2: return launch_diag<C, 2, 2>(h, state, nqubits, op);
        case 3: return launch_diag<C, 3, 1>(h, state, nqubits, op);
        case 4: return launch_diag<C, 4, 1>(h, state, nqubits, op);
        case 5: return launch_diag<C, 5, 1>(h, state, nqubits, op);
        case 6: return launch_diag<C, 6, 1>(h, state, nqubits, op);
        default: return fail(QB_ERR_UNSUPPORTED, "diagonal gates on more than 6 target qubits are not supported");
      }
  }
  return fail(QB_ERR_INVALID, "bad canonical op");
}

static int apply_gate_common(qb_handle h, void* state, int nqubits, int dtype, const double* data, bool is_diag,
                             int ntargets, const int* targets, int ncontrols, const int* controls) {
  if (!h || !valid_state_args(state, nqubits, dtype)) return fail(QB_ERR_INVALID, "bad state arguments");
  if (!data || (ntargets > 0 && !targets) || (ncontrols > 0 && !controls)) return fail(QB_ERR_INVALID, "null gate argument");
  if (ntargets > QB_MAX_OP_TARGETS) return fail(QB_ERR_UNSUPPORTED, "too many target qubits");
  CanonOp op;
  std::string err;
  if (!canonicalize(nqubits, data, is_diag, ntargets, targets, ncontrols, controls, op, err)) return fail(QB_ERR_INVALID, err);
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  return dtype == QB_C128 ? apply_canon_k1<double2>(h, state, nqubits, op) : apply_canon_k1<float2>(h, state, nqubits, op);
}

int qb_apply_matrix(qb_handle h, void* state, int nqubits, int dtype, const double* matrix, int ntargets, const int* targets,
                    int ncontrols, const int* controls) {
  return apply_gate_common(h, state, nqubits, dtype, matrix, false, ntargets, targets, ncontrols, controls);
}

int qb_apply_diagonal(qb_handle h, void* state, int nqubits, int dtype, const double* diag, int ntargets, const int* targets,
                      int ncontrols, const int* controls) {
  return apply_gate_common(h, state, nqubits, dtype, diag, true, ntargets, targets, ncontrols, controls);
}

// ---------------------------------------------------------------------------------------------------
// K2: whole programs
// ---------------------------------------------------------------------------------------------------
static int canonicalize_program(int nqubits, const qb_op* ops, int nops, std::vector<CanonOp>& out) {
  out.clear();
  out.reserve(nops);
  std::string err;
  for (int i = 0; i < nops; ++i) {
    const qb_op& o = ops[i];
    if (o.ntargets < 0 || o.ntargets > QB_MAX_OP_TARGETS || o.ncontrols < 0 || o.ncontrols > QB_MAX_OP_CONTROLS || !o.data)
      return fail(QB_ERR_INVALID, "bad op " + std::to_string(i));
    CanonOp c;
    if (!canonicalize(nqubits, o.data, o.is_diagonal != 0, o.ntargets, o.targets, o.ncontrols, o.controls, c, err))
      return fail(QB_ERR_INVALID, "op " + std::to_string(i) + ": " + err);
    out.push_back(std::move(c));
  }
  return QB_OK;
}

int qb_plan_program(int nqubits, int dtype, const qb_op* ops, int nops, int flags, qb_program_stats* stats,
                    int32_t* sweep_of_op) {
  if (nqubits < 1 || nqubits > QB_MAX_QUBITS || (dtype != QB_C64 && dtype != QB_C128) || nops < 0 || (nops && !ops))
    return fail(QB_ERR_INVALID, "bad program arguments");
  std::vector<CanonOp> canon;
  int rc = canonicalize_program(nqubits, ops, nops, canon);
  if (rc != QB_OK) return rc;
  Plan plan;
  std::string err;
  if (!plan_program(nqubits, dtype, canon, (flags & QB_PROGRAM_NO_FUSE) != 0, plan, err)) return fail(QB_ERR_UNSUPPORTED, err);
  if (stats) fill_stats(plan, nqubits, dtype, nops, stats);
  if (sweep_of_op)
    for (int i = 0; i < nops; ++i) sweep_of_op[i] = plan.sweep_of_op[i];
  return QB_OK;
}

// dest_of_qubit (Qibo qubit ids) -> PermSpec (bit positions); false when it is not a permutation
static bool perm_from_qubits(int nqubits, const int* dest_of_qubit, PermSpec& perm) {
  uint64_t seen = 0;
  for (int q = 0; q < nqubits; ++q) {
    const int d = dest_of_qubit[q];
    if (d < 0 || d >= nqubits || ((seen >> d) & 1)) return false;
    seen |= uint64_t(1) << d;
    perm.pi[nqubits - 1 - q] = nqubits - 1 - d;
  }
  return true;
}

// K8 launch (the caller holds the context's mutex)
static int permute_locked(qb_context* h, const void* src, void* dst, int nqubits, int dtype, const int* pi) {
  PermParams p;
  memset(&p, 0, sizeof(p));
  // 2^6 amplitudes: 1 KiB (complex128) / 512 B (complex64) contiguous on both sides (QB_PERM_LOW_BITS: tuning knob)
  int lowbits = env_int("QB_PERM_LOW_BITS", 6);
  if (lowbits < 3) lowbits = 3;
  if (lowbits > 6) lowbits = 6;
  perm_setup(nqubits, lowbits, lowbits, pi, p);
  if (launch_permute(h->stream, h->sm_count, src, dst, dtype, p) != QB_OK) return cuda_fail(cudaGetLastError(), "k8_permute");
  return QB_OK;
}

// a permutation that rides on a sweep needs the tensor-map encoder at launch time
static const PermSpec* fusable(const PermSpec* perm) { return perm && tma_encoder() != nullptr ? perm : nullptr; }

// the sweeps of a plan, then -- when a permutation was asked for and does not ride on the last sweep -- K8 into `dst`
static int launch_plan(qb_context* h, void* state, void* dst, int nqubits, int dtype, const Plan& plan, const char* prog_dev,
                       const PermSpec* perm, bool input_zero = false) {
  if (input_zero && plan.sweeps.empty()) {  // nothing will make the state on the way: write it
    const int rc = set_basis_locked(h, state, nqubits, dtype, 0);
    if (rc != QB_OK) return rc;
  }
  for (size_t s = 0; s < plan.sweeps.size(); ++s) {
    const int rc = launch_sweep(h->stream, h->sm_count, state, nqubits, dtype, plan.sweeps[s], prog_dev, dst, input_zero && s == 0);
    if (rc != QB_OK) {
      const std::string what = cudaGetErrorString(cudaGetLastError());
      return fail(rc, "sweep launch failed: " + what + " [" + sweep_resources(dtype) + "]");
    }
  }
  if (perm && !plan.perm_fused) return permute_locked(h, state, dst, nqubits, dtype, perm->pi);
  return QB_OK;
}

static void count_k8(const Plan& plan, const PermSpec* perm, int nqubits, int dtype, qb_program_stats* st) {
  if (!perm || plan.perm_fused) return;
  st->nsweeps += 1;
  st->ndense_passes += 1;
  st->bytes_moved += 2.0 * (dtype == QB_C128 ? 16.0 : 8.0) * (double)(uint64_t(1) << nqubits);
}

int qb_plan_program_permuted(int nqubits, int dtype, const qb_op* ops, int nops, const int* dest_of_qubit, int flags,
                             qb_program_stats* stats, int32_t* sweep_of_op) {
  if (nqubits < 1 || nqubits > QB_MAX_QUBITS || (dtype != QB_C64 && dtype != QB_C128) || nops < 0 || (nops && !ops) || !dest_of_qubit)
    return fail(QB_ERR_INVALID, "bad program arguments");
  PermSpec perm;
  if (!perm_from_qubits(nqubits, dest_of_qubit, perm)) return fail(QB_ERR_INVALID, "dest_of_qubit is not a permutation");
  std::vector<CanonOp> canon;
  int rc = canonicalize_program(nqubits, ops, nops, canon);
  if (rc != QB_OK) return rc;
  Plan plan;
  std::string err;
  if (!plan_program(nqubits, dtype, canon, (flags & QB_PROGRAM_NO_FUSE) != 0, plan, err, nullptr, &perm)) return fail(QB_ERR_UNSUPPORTED, err);
  if (stats) {
    fill_stats(plan, nqubits, dtype, nops, stats);
    count_k8(plan, &perm, nqubits, dtype, stats);
  }
  if (sweep_of_op)
    for (int i = 0; i < nops; ++i) sweep_of_op[i] = plan.sweep_of_op[i];
  return QB_OK;
}

static int apply_program_impl(qb_handle h, void* state, void* dst, int nqubits, int dtype, const qb_op* ops, int nops,
                              const PermSpec* perm, int flags, qb_program_stats* stats);

int qb_apply_program(qb_handle h, void* state, int nqubits, int dtype, const qb_op* ops, int nops, int flags,
                     qb_program_stats* stats) {
  if (!h || !valid_state_args(state, nqubits, dtype) || nops < 0 || (nops && !ops))
    return fail(QB_ERR_INVALID, "bad program arguments");
  return apply_program_impl(h, state, nullptr, nqubits, dtype, ops, nops, nullptr, flags, stats);
}

int qb_apply_program_permuted(qb_handle h, void* state, void* dst, int nqubits, int dtype, const qb_op* ops, int nops,
                              const int* dest_of_qubit, int flags, qb_program_stats* stats) {
  if (!h || !valid_state_args(state, nqubits, dtype) || !dst || dst == state || nops < 0 || (nops && !ops) || !dest_of_qubit)
    return fail(QB_ERR_INVALID, "bad program arguments");
  PermSpec perm;
  if (!perm_from_qubits(nqubits, dest_of_qubit, perm)) return fail(QB_ERR_INVALID, "dest_of_qubit is not a permutation");
  return apply_program_impl(h, state, dst, nqubits, dtype, ops, nops, &perm, flags, stats);
}

static int apply_program_impl(qb_handle h, void* state, void* dst, int nqubits, int dtype, const qb_op* ops, int nops,
                              const PermSpec* perm, int flags, qb_program_stats* stats) {
  std::vector<CanonOp> canon;
  int rc = canonicalize_program(nqubits, ops, nops, canon);
  if (rc != QB_OK) return rc;
  if (nqubits < 4 && perm && (flags & QB_PROGRAM_PERM_FUSED_ONLY)) return fail(QB_ERR_UNSUPPORTED, "no sweeps below 4 qubits");
  if (nqubits < 4) {
    // up to 8 amplitudes: nothing to tile (a register group of the sweep kernel is 2^4 amplitudes); the K1 kernels
    // apply the queue gate by gate
    std::lock_guard<std::mutex> lk(h->mu);
    DeviceGuard guard(h->device);
    if (flags & QB_PROGRAM_INPUT_ZERO) {
      rc = set_basis_locked(h, state, nqubits, dtype, 0);
      if (rc != QB_OK) return rc;
    }
    if (stats) {
      memset(stats, 0, sizeof(*stats));
      stats->nops = nops;
      stats->nsweeps = nops;
      stats->bytes_moved = (double)nops * 2.0 * (dtype == QB_C128 ? 16.0 : 8.0) * (double)(uint64_t(1) << nqubits);
    }
    for (auto& c : canon) {
      rc = dtype == QB_C128 ? apply_canon_k1<double2>(h, state, nqubits, c) : apply_canon_k1<float2>(h, state, nqubits, c);
      if (rc != QB_OK) return rc;
    }
    if (perm) {
      if (stats) stats->nsweeps += 1;
      return permute_locked(h, state, dst, nqubits, dtype, perm->pi);
    }
    return QB_OK;
  }
  Plan plan;
  std::string err;
  if (!plan_program(nqubits, dtype, canon, (flags & QB_PROGRAM_NO_FUSE) != 0, plan, err, nullptr, fusable(perm))) return fail(QB_ERR_UNSUPPORTED, err);
  if (perm && !plan.perm_fused && (flags & QB_PROGRAM_PERM_FUSED_ONLY)) return fail(QB_ERR_UNSUPPORTED, "the permutation cannot ride on a sweep of this program");
  if (stats) {
    fill_stats(plan, nqubits, dtype, nops, stats);
    count_k8(plan, perm, nqubits, dtype, stats);
  }

  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  // upload the compiled sweeps (pinned staging -> device program buffer)
  size_t need = plan.blob.size();
  if (need > h->prog_bytes) {
    QB_CUDA(cudaStreamSynchronize(h->stream));
    if (h->prog_dev) cudaFree(h->prog_dev);
    if (h->prog_host) cudaFreeHost(h->prog_host);
    h->prog_dev = h->prog_host = nullptr;
    h->prog_bytes = 0;
    size_t want = need < (size_t(1) << 20) ? (size_t(1) << 20) : need * 2;
    QB_CUDA(cudaMalloc(&h->prog_dev, want));
    QB_CUDA(cudaMallocHost(&h->prog_host, want));
    h->prog_bytes = want;
  } else {
    QB_CUDA(cudaEventSynchronize(h->prog_done));  // previous program may still be reading the buffers
  }
  if (need) {
    memcpy(h->prog_host, plan.blob.data(), need);
    QB_CUDA(cudaMemcpyAsync(h->prog_dev, h->prog_host, need, cudaMemcpyHostToDevice, h->stream));
  }
  if (flags & QB_PROGRAM_TIME) QB_CUDA(cudaEventRecord(h->ev0, h->stream));
  rc = launch_plan(h, state, dst, nqubits, dtype, plan, (const char*)h->prog_dev, perm, (flags & QB_PROGRAM_INPUT_ZERO) != 0);
  if (rc != QB_OK) return rc;
  QB_CUDA(cudaEventRecord(h->prog_done, h->stream));
  if (flags & QB_PROGRAM_TIME) {
    QB_CUDA(cudaEventRecord(h->ev1, h->stream));
    QB_CUDA(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    QB_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    if (stats) stats->elapsed_ms = ms;
  }
  return QB_OK;
}

// ---- compiled programs: plan once, keep the sweep programs resident in device memory, launch many times ----------
struct OpDesc {  // what qb_program_set_params needs to rebuild one op: its qubits and its current host data
  int ntargets = 0, ncontrols = 0, is_diagonal = 0;
  int targets[QB_MAX_OP_TARGETS], controls[QB_MAX_OP_CONTROLS];
  std::vector<double> data;  // interleaved complex128
};
struct qb_program_s {
  int nqubits = 0, dtype = 0, nops = 0, device = 0, flags = 0;
  Plan plan;                    // sweep descriptors + the schedule (the blob itself lives in `dev`)
  std::vector<CanonOp> canon;   // canonical ops (re-planned on a parameter update; nqubits < 4: applied gate by gate)
  std::vector<OpDesc> descs;
  void* dev = nullptr;
  size_t dev_bytes = 0;
  void* staging = nullptr;      // pinned host copy of the blob for stream-ordered re-uploads
  size_t staging_bytes = 0;
  qb_program_stats stats;
  bool has_perm = false;        // qb_program_create_permuted: the ops are followed by `perm`, the result goes to another buffer
  PermSpec perm;
};

static int program_create_impl(qb_handle h, int nqubits, int dtype, const qb_op* ops, int nops, const PermSpec* perm, int flags,
                               qb_program* out, qb_program_stats* stats);

int qb_program_create(qb_handle h, int nqubits, int dtype, const qb_op* ops, int nops, int flags, qb_program* out,
                      qb_program_stats* stats) {
  if (!h || !out || nqubits < 1 || nqubits > QB_MAX_QUBITS || (dtype != QB_C64 && dtype != QB_C128) || nops < 0 || (nops && !ops))
    return fail(QB_ERR_INVALID, "bad program arguments");
  return program_create_impl(h, nqubits, dtype, ops, nops, nullptr, flags, out, stats);
}

int qb_program_create_permuted(qb_handle h, int nqubits, int dtype, const qb_op* ops, int nops, const int* dest_of_qubit, int flags,
                               qb_program* out, qb_program_stats* stats) {
  if (!h || !out || nqubits < 1 || nqubits > QB_MAX_QUBITS || (dtype != QB_C64 && dtype != QB_C128) || nops < 0 || (nops && !ops) ||
      !dest_of_qubit)
    return fail(QB_ERR_INVALID, "bad program arguments");
  PermSpec perm;
  if (!perm_from_qubits(nqubits, dest_of_qubit, perm)) return fail(QB_ERR_INVALID, "dest_of_qubit is not a permutation");
  return program_create_impl(h, nqubits, dtype, ops, nops, &perm, flags, out, stats);
}

static int program_create_impl(qb_handle h, int nqubits, int dtype, const qb_op* ops, int nops, const PermSpec* perm, int flags,
                               qb_program* out, qb_program_stats* stats) {
  *out = nullptr;
  std::unique_ptr<qb_program_s> p(new qb_program_s());
  if (perm) {
    p->has_perm = true;
    p->perm = *perm;
  }
  p->nqubits = nqubits;
  p->dtype = dtype;
  p->nops = nops;
  p->device = h->device;
  p->flags = flags;
  int rc = canonicalize_program(nqubits, ops, nops, p->canon);
  if (rc != QB_OK) return rc;
  p->descs.resize(nops);
  for (int i = 0; i < nops; ++i) {
    OpDesc& d = p->descs[i];
    d.ntargets = ops[i].ntargets;
    d.ncontrols = ops[i].ncontrols;
    d.is_diagonal = ops[i].is_diagonal != 0;
    for (int t = 0; t < d.ntargets; ++t) d.targets[t] = ops[i].targets[t];
    for (int c = 0; c < d.ncontrols; ++c) d.controls[c] = ops[i].controls[c];
    const size_t dim = size_t(1) << d.ntargets;
    d.data.assign(ops[i].data, ops[i].data + 2 * (d.is_diagonal ? dim : dim * dim));
  }
  memset(&p->stats, 0, sizeof(p->stats));
  p->stats.nops = nops;
  if (nqubits < 4 && perm && (flags & QB_PROGRAM_PERM_FUSED_ONLY)) return fail(QB_ERR_UNSUPPORTED, "no sweeps below 4 qubits");
  if (nqubits < 4) {
    p->stats.nsweeps = nops + (perm ? 1 : 0);
    p->stats.bytes_moved = (double)nops * 2.0 * (dtype == QB_C128 ? 16.0 : 8.0) * (double)(uint64_t(1) << nqubits);
  } else {
    std::string err;
    if (!plan_program(nqubits, dtype, p->canon, (flags & QB_PROGRAM_NO_FUSE) != 0, p->plan, err, nullptr, fusable(perm)))
      return fail(QB_ERR_UNSUPPORTED, err);
    if (perm && !p->plan.perm_fused && (flags & QB_PROGRAM_PERM_FUSED_ONLY))
      return fail(QB_ERR_UNSUPPORTED, "the permutation cannot ride on a sweep of this program");
    fill_stats(p->plan, nqubits, dtype, nops, &p->stats);
    count_k8(p->plan, perm, nqubits, dtype, &p->stats);
    if (!p->plan.blob.empty()) {
      std::lock_guard<std::mutex> lk(h->mu);
      DeviceGuard guard(h->device);
      QB_CUDA(cudaMalloc(&p->dev, p->plan.blob.size()));
      p->dev_bytes = p->plan.blob.size();
      cudaError_t e = cudaMemcpyAsync(p->dev, p->plan.blob.data(), p->plan.blob.size(), cudaMemcpyHostToDevice, h->stream);
      if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);  // the host blob is released below
      if (e != cudaSuccess) {
        cudaFree(p->dev);
        return cuda_fail(e, "program upload");
      }
      std::vector<char>().swap(p->plan.blob);
    }
  }
  if (stats) *stats = p->stats;
  *out = p.release();
  return QB_OK;
}

static int program_run_impl(qb_handle h, qb_program p, void* state, void* dst, int flags, qb_program_stats* stats);

int qb_program_run(qb_handle h, qb_program p, void* state, int flags, qb_program_stats* stats) {
  if (!h || !p || !valid_state_args(state, p->nqubits, p->dtype)) return fail(QB_ERR_INVALID, "bad program arguments");
  if (p->has_perm) return fail(QB_ERR_INVALID, "the program ends with a permutation: run it with qb_program_run_permuted");
  return program_run_impl(h, p, state, nullptr, flags, stats);
}

int qb_program_run_permuted(qb_handle h, qb_program p, void* state, void* dst, int flags, qb_program_stats* stats) {
  if (!h || !p || !valid_state_args(state, p->nqubits, p->dtype) || !dst || dst == state) return fail(QB_ERR_INVALID, "bad program arguments");
  if (!p->has_perm) return fail(QB_ERR_INVALID, "the program was not created with qb_program_create_permuted");
  return program_run_impl(h, p, state, dst, flags, stats);
}

static int program_run_impl(qb_handle h, qb_program p, void* state, void* dst, int flags, qb_program_stats* stats) {
  if (p->device != h->device) return fail(QB_ERR_INVALID, "the program was compiled for another device");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  if (stats) *stats = p->stats;
  if (flags & QB_PROGRAM_TIME) QB_CUDA(cudaEventRecord(h->ev0, h->stream));
  if (p->nqubits < 4) {
    if (flags & QB_PROGRAM_INPUT_ZERO) {
      int rc = set_basis_locked(h, state, p->nqubits, p->dtype, 0);
      if (rc != QB_OK) return rc;
    }
    for (auto& c : p->canon) {
      int rc = p->dtype == QB_C128 ? apply_canon_k1<double2>(h, state, p->nqubits, c) : apply_canon_k1<float2>(h, state, p->nqubits, c);
      if (rc != QB_OK) return rc;
    }
    if (p->has_perm) {
      int rc = permute_locked(h, state, dst, p->nqubits, p->dtype, p->perm.pi);
      if (rc != QB_OK) return rc;
    }
  } else {
    int rc = launch_plan(h, state, dst, p->nqubits, p->dtype, p->plan, (const char*)p->dev, p->has_perm ? &p->perm : nullptr,
                         (flags & QB_PROGRAM_INPUT_ZERO) != 0);
    if (rc != QB_OK) return rc;
  }
  if (flags & QB_PROGRAM_TIME) {
    QB_CUDA(cudaEventRecord(h->ev1, h->stream));
    QB_CUDA(cudaEventSynchronize(h->ev1));
    float ms = 0.f;
    QB_CUDA(cudaEventElapsedTime(&ms, h->ev0, h->ev1));
    if (stats) stats->elapsed_ms = ms;
  }
  return QB_OK;
}

// ---- parameter slots: new matrices for some ops of a compiled program, schedule kept --------------------------------
// Gate families follow backends/npmatrices.py (RX :79, RY :84, RZ :89, U1 :96-ish, CU1 :230, CRX/CRY/CRZ): evaluated in
// double here so that a variational loop sends angles, not matrices.
int qb_program_set_params(qb_handle h, qb_program p, const qb_param_update* updates, int nupdates) {
  if (!h || !p || nupdates < 0 || (nupdates && !updates)) return fail(QB_ERR_INVALID, "bad parameter-update arguments");
  if (p->device != h->device) return fail(QB_ERR_INVALID, "the program was compiled for another device");
  if (nupdates == 0) return QB_OK;
  std::string err;
  for (int u = 0; u < nupdates; ++u) {
    const qb_param_update& up = updates[u];
    if (up.op_index < 0 || up.op_index >= p->nops) return fail(QB_ERR_INVALID, "parameter update: op index out of range");
    OpDesc& d = p->descs[up.op_index];
    std::vector<double> data;
    if (up.family == QB_GATE_MATRIX) {
      if (!up.matrix) return fail(QB_ERR_INVALID, "parameter update: null matrix");
      data.assign(up.matrix, up.matrix + d.data.size());
    } else if (!family_matrix(up.family, up.theta, d.ntargets, d.is_diagonal != 0, data) || data.size() != d.data.size()) {
      return fail(QB_ERR_INVALID, "parameter update: gate family does not match the op it updates");
    }
    if (up.conjugate)
      for (size_t i = 1; i < data.size(); i += 2) data[i] = -data[i];
    d.data.swap(data);
    CanonOp c;
    if (!canonicalize(p->nqubits, d.data.data(), d.is_diagonal != 0, d.ntargets, d.targets, d.ncontrols, d.controls, c, err))
      return fail(QB_ERR_INVALID, "parameter update, op " + std::to_string(up.op_index) + ": " + err);
    p->canon[up.op_index] = std::move(c);
  }
  if (p->nqubits < 4) return QB_OK;  // applied gate by gate from `canon`
  const bool no_fuse = (p->flags & QB_PROGRAM_NO_FUSE) != 0;
  Plan fresh;
  const PermSpec* perm = p->has_perm ? fusable(&p->perm) : nullptr;
  if (!plan_program(p->nqubits, p->dtype, p->canon, no_fuse, fresh, err, &p->plan, perm)) {
    // the structure changed (a rotation became an identity, a real matrix complex...): schedule from scratch
    if (!plan_program(p->nqubits, p->dtype, p->canon, no_fuse, fresh, err, nullptr, perm)) return fail(QB_ERR_UNSUPPORTED, err);
  }
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  const size_t need = fresh.blob.size();
  if (need > p->dev_bytes) {
    QB_CUDA(cudaStreamSynchronize(h->stream));  // launches may still read the old program
    if (p->dev) cudaFree(p->dev);
    p->dev = nullptr;
    p->dev_bytes = 0;
    QB_CUDA(cudaMalloc(&p->dev, need));
    p->dev_bytes = need;
  }
  if (need > p->staging_bytes) {
    QB_CUDA(cudaStreamSynchronize(h->stream));
    if (p->staging) cudaFreeHost(p->staging);
    p->staging = nullptr;
    p->staging_bytes = 0;
    QB_CUDA(cudaMallocHost(&p->staging, need));
    p->staging_bytes = need;
  } else {
    QB_CUDA(cudaStreamSynchronize(h->stream));  // (the previous re-upload may still be reading the staging copy)
  }
  if (need) {
    memcpy(p->staging, fresh.blob.data(), need);
    // stream-ordered: earlier launches of this program have read the old blob by the time the copy runs
    QB_CUDA(cudaMemcpyAsync(p->dev, p->staging, need, cudaMemcpyHostToDevice, h->stream));
  }
  std::vector<char>().swap(fresh.blob);
  p->plan = std::move(fresh);
  fill_stats(p->plan, p->nqubits, p->dtype, p->nops, &p->stats);
  count_k8(p->plan, p->has_perm ? &p->perm : nullptr, p->nqubits, p->dtype, &p->stats);
  return QB_OK;
}

int qb_program_destroy(qb_handle h, qb_program p) {
  if (!p) return QB_OK;
  if (p->staging) cudaFreeHost(p->staging);
  if (p->dev) {
    if (h) {
      std::lock_guard<std::mutex> lk(h->mu);
      DeviceGuard guard(h->device);
      cudaStreamSynchronize(h->stream);  // a launch may still be reading the program
      cudaFree(p->dev);
    } else {
      cudaFree(p->dev);
    }
  }
  delete p;
  return QB_OK;
}

// ---------------------------------------------------------------------------------------------------
// K8
// ---------------------------------------------------------------------------------------------------
int qb_permute_qubits(qb_handle h, const void* src, void* dst, int nqubits, int dtype, const int* dest_of_qubit) {
  if (!h || !valid_state_args(src, nqubits, dtype) || !dst || !dest_of_qubit || src == dst)
    return fail(QB_ERR_INVALID, "bad permutation arguments");
  int pi[64];
  uint64_t seen = 0;
  for (int q = 0; q < nqubits; ++q) {
    int d = dest_of_qubit[q];
    if (d < 0 || d >= nqubits || ((seen >> d) & 1)) return fail(QB_ERR_INVALID, "dest_of_qubit is not a permutation");
    seen |= uint64_t(1) << d;
    pi[nqubits - 1 - q] = nqubits - 1 - d;
  }
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  return permute_locked(h, src, dst, nqubits, dtype, pi);
}

// ---------------------------------------------------------------------------------------------------
// K9: expectation values
// ---------------------------------------------------------------------------------------------------
static int reduce2_to_host(qb_context* h, int grid, double* out_host) {
  double* partial = (double*)h->scratch + 8;
  double* out = (double*)h->scratch;
  k9_sum_partials2<<<1, 32, 0, h->stream>>>(partial, partial + grid, grid, out);
  QB_CHECK_LAUNCH("k9_sum_partials2");
  QB_CUDA(cudaMemcpyAsync(out_host, out, 2 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  QB_CUDA(cudaStreamSynchronize(h->stream));
  return QB_OK;
}

int qb_expval_pauli(qb_handle h, const void* state, int nqubits, int dtype, const char* paulis, const int* qubits, int nterms_qubits,
                    double* out_host) {
  if (!h || !valid_state_args(state, nqubits, dtype) || !out_host || nterms_qubits < 0 || (nterms_qubits && (!paulis || !qubits)))
    return fail(QB_ERR_INVALID, "bad expectation-value arguments");
  uint64_t xmask = 0, zmask = 0, seen = 0;
  int ny = 0;
  for (int i = 0; i < nterms_qubits; ++i) {
    const int q = qubits[i];
    if (q < 0 || q >= nqubits || ((seen >> q) & 1)) return fail(QB_ERR_INVALID, "bad or repeated qubit in Pauli string");
    seen |= uint64_t(1) << q;
    const uint64_t bit = uint64_t(1) << (nqubits - 1 - q);
    switch (paulis[i]) {
      case 'I': break;
      case 'X': xmask |= bit; break;
      case 'Y': xmask |= bit; zmask |= bit; ++ny; break;
      case 'Z': zmask |= bit; break;
      default: return fail(QB_ERR_INVALID, "Pauli strings are made of I, X, Y, Z");
    }
  }
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  const uint64_t count = uint64_t(1) << nqubits;
  const int grid = grid_for(count, RED_THREADS, h->sm_count, 8);
  int rc = ensure_scratch(h, (size_t)(2 * grid + 8) * sizeof(double));
  if (rc != QB_OK) return rc;
  double* partial = (double*)h->scratch + 8;
  if (dtype == QB_C128) k9_pauli_expval<double2><<<grid, RED_THREADS, 0, h->stream>>>((const double2*)state, count, xmask, zmask, partial, partial + grid);
  else k9_pauli_expval<float2><<<grid, RED_THREADS, 0, h->stream>>>((const float2*)state, count, xmask, zmask, partial, partial + grid);
  QB_CHECK_LAUNCH("k9_pauli_expval");
  double v[2];
  rc = reduce2_to_host(h, grid, v);
  if (rc != QB_OK) return rc;
  // times i^{nY}
  switch (ny & 3) {
    case 0: out_host[0] = v[0]; out_host[1] = v[1]; break;
    case 1: out_host[0] = -v[1]; out_host[1] = v[0]; break;
    case 2: out_host[0] = -v[0]; out_host[1] = -v[1]; break;
    default: out_host[0] = v[1]; out_host[1] = -v[0]; break;
  }
  return QB_OK;
}

int qb_state_vdot(qb_handle h, const void* a, const void* b, int nqubits, int dtype, double* out_host) {
  if (!h || !valid_state_args(a, nqubits, dtype) || !b || !out_host) return fail(QB_ERR_INVALID, "bad inner-product arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  const uint64_t count = uint64_t(1) << nqubits;
  const int grid = grid_for(count, RED_THREADS, h->sm_count, 8);
  int rc = ensure_scratch(h, (size_t)(2 * grid + 8) * sizeof(double));
  if (rc != QB_OK) return rc;
  double* partial = (double*)h->scratch + 8;
  if (dtype == QB_C128) k9_vdot<double2><<<grid, RED_THREADS, 0, h->stream>>>((const double2*)a, (const double2*)b, count, partial, partial + grid);
  else k9_vdot<float2><<<grid, RED_THREADS, 0, h->stream>>>((const float2*)a, (const float2*)b, count, partial, partial + grid);
  QB_CHECK_LAUNCH("k9_vdot");
  return reduce2_to_host(h, grid, out_host);
}

// ---------------------------------------------------------------------------------------------------
// K3
// ---------------------------------------------------------------------------------------------------
int qb_probabilities(qb_handle h, const void* state, int nqubits, int dtype, const int* qubits, int nmeasured,
                     void* probs_out) {
  if (!h || !valid_state_args(state, nqubits, dtype) || !probs_out) return fail(QB_ERR_INVALID, "bad state arguments");
  if (nmeasured < 0 || nmeasured > nqubits || (nmeasured && !qubits)) return fail(QB_ERR_INVALID, "bad measured qubits");
  ProbParams p;
  memset(&p, 0, sizeof(p));
  p.n = nqubits;
  p.nlow = nqubits < 5 ? nqubits : 5;
  for (int i = 0; i < 48; ++i) p.outbit[i] = -1;
  uint64_t mmask = 0;
  for (int i = 0; i < nmeasured; ++i) {
    int q = qubits[i];
    if (q < 0 || q >= nqubits) return fail(QB_ERR_INVALID, "measured qubit out of range");
    int pos = nqubits - 1 - q;
    if ((mmask >> pos) & 1) return fail(QB_ERR_INVALID, "repeated measured qubit");
    mmask |= uint64_t(1) << pos;
    p.outbit[pos] = (int8_t)(nmeasured - 1 - i);
  }
  uint64_t lowmask = (uint64_t(1) << p.nlow) - 1;
  uint64_t all = nqubits == 64 ? ~uint64_t(0) : ((uint64_t(1) << nqubits) - 1);
  p.low_unmeasured = (uint32_t)(~mmask & lowmask);
  p.mhigh_mask = mmask & ~lowmask;
  p.uhigh_mask = ~mmask & all & ~lowmask;
  p.n_mhigh = __builtin_popcountll(p.mhigh_mask);
  p.n_uhigh = __builtin_popcountll(p.uhigh_mask);
  p.nbins = uint64_t(1) << nmeasured;
  // want >= ~2^14 warps; never split below 4 iterations per warp
  int want = 14 - p.n_mhigh;
  if (want < 0) want = 0;
  int maxsplit = p.n_uhigh - 2;
  if (maxsplit < 0) maxsplit = 0;
  p.log_split = want < maxsplit ? want : maxsplit;
  // bound the partial buffer (2^log_split * nbins doubles) to 64 MiB
  while (p.log_split > 0 && ((p.nbins << p.log_split) * sizeof(double)) > (size_t(64) << 20)) --p.log_split;

  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  bool identity = nmeasured == nqubits;
  for (int i = 0; i < nmeasured && identity; ++i) identity = qubits[i] == i;
  if (identity) {  // every qubit, ascending: a plain elementwise pass
    uint64_t count = uint64_t(1) << nqubits;
    int grid = grid_for((count + 1) / 2, 256, h->sm_count, 32);
    if (dtype == QB_C128) k3_probs_full<double2, double><<<grid, 256, 0, h->stream>>>((const double2*)state, (double*)probs_out, count);
    else k3_probs_full<float2, float><<<grid, 256, 0, h->stream>>>((const float2*)state, (float*)probs_out, count);
    QB_CHECK_LAUNCH("k3_probs_full");
    return QB_OK;
  }
  double* partial = nullptr;
  if (p.log_split > 0) {
    int rc = ensure_scratch(h, (size_t)(p.nbins << p.log_split) * sizeof(double));
    if (rc != QB_OK) return rc;
    partial = (double*)h->scratch;
  }
  uint64_t nwarps = uint64_t(1) << (p.n_mhigh + p.log_split);
  uint64_t grid = (nwarps + 7) / 8;
  if (dtype == QB_C128)
    k3_probs<double2, double><<<(unsigned)grid, 256, 0, h->stream>>>((const double2*)state, (double*)probs_out, partial, p);
  else
    k3_probs<float2, float><<<(unsigned)grid, 256, 0, h->stream>>>((const float2*)state, (float*)probs_out, partial, p);
  QB_CHECK_LAUNCH("k3_probs");
  if (p.log_split > 0) {
    int g2 = grid_for(p.nbins, 256, h->sm_count);
    if (dtype == QB_C128)
      k3_finish<double><<<g2, 256, 0, h->stream>>>(partial, (double*)probs_out, p.nbins, 1 << p.log_split);
    else
      k3_finish<float><<<g2, 256, 0, h->stream>>>(partial, (float*)probs_out, p.nbins, 1 << p.log_split);
    QB_CHECK_LAUNCH("k3_finish");
  }
  return QB_OK;
}

// ---------------------------------------------------------------------------------------------------
// K4
// ---------------------------------------------------------------------------------------------------
static int cdf_locked(qb_context* h, const void* probs, int rdtype, uint64_t nbins, double* cdf_out, int mode,
                      double* total_dev = nullptr) {
  if (mode == QB_SCAN_EXACT) {
    if (rdtype == QB_F64) k4_scan_exact<double><<<1, 256, 0, h->stream>>>((const double*)probs, cdf_out, nbins);
    else k4_scan_exact<float><<<1, 256, 0, h->stream>>>((const float*)probs, cdf_out, nbins);
    QB_CHECK_LAUNCH("k4_scan_exact");
  } else {
    uint64_t nblocks = (nbins + PSCAN_BLOCK - 1) / PSCAN_BLOCK;
    int rc = ensure_scratch(h, (size_t)nblocks * 2 * sizeof(double));
    if (rc != QB_OK) return rc;
    double* tot = (double*)h->scratch;
    double* off = tot + nblocks;
    if (rdtype == QB_F64) {
      k4_scan_block<double, false><<<(unsigned)nblocks, PSCAN_THREADS, 0, h->stream>>>((const double*)probs, cdf_out, tot, off, nbins);
      k4_scan_totals<<<1, 32, 0, h->stream>>>(tot, off, nblocks);
      k4_scan_block<double, true><<<(unsigned)nblocks, PSCAN_THREADS, 0, h->stream>>>((const double*)probs, cdf_out, tot, off, nbins);
    } else {
      k4_scan_block<float, false><<<(unsigned)nblocks, PSCAN_THREADS, 0, h->stream>>>((const float*)probs, cdf_out, tot, off, nbins);
      k4_scan_totals<<<1, 32, 0, h->stream>>>(tot, off, nblocks);
      k4_scan_block<float, true><<<(unsigned)nblocks, PSCAN_THREADS, 0, h->stream>>>((const float*)probs, cdf_out, tot, off, nbins);
    }
    QB_CHECK_LAUNCH("k4_scan_block");
  }
  if (total_dev) QB_CUDA(cudaMemcpyAsync(total_dev, cdf_out + (nbins - 1), sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  k4_normalize<<<grid_for(nbins, 256, h->sm_count), 256, 0, h->stream>>>(cdf_out, nbins);
  k4_normalize_last<<<1, 32, 0, h->stream>>>(cdf_out, nbins);
  QB_CHECK_LAUNCH("k4_normalize");
  return QB_OK;
}

int qb_cdf(qb_handle h, const void* probs, int rdtype, uint64_t nbins, double* cdf_out, int mode) {
  if (!h || !probs || !cdf_out || nbins == 0 || (rdtype != QB_F32 && rdtype != QB_F64) ||
      (mode != QB_SCAN_EXACT && mode != QB_SCAN_PARALLEL))
    return fail(QB_ERR_INVALID, "bad cdf arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  return cdf_locked(h, probs, rdtype, nbins, cdf_out, mode);
}

int qb_sample_cdf(qb_handle h, const double* cdf, uint64_t nbins, const double* uniforms, uint64_t nshots, int64_t* out) {
  if (!h || !cdf || nbins == 0 || (nshots && (!uniforms || !out))) return fail(QB_ERR_INVALID, "bad sampling arguments");
  if (nshots == 0) return QB_OK;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  k4_search<<<grid_for(nshots, 256, h->sm_count), 256, 0, h->stream>>>(cdf, nbins, uniforms, nshots, (long long*)out);
  QB_CHECK_LAUNCH("k4_search");
  return QB_OK;
}

int qb_sample(qb_handle h, const void* probs, int rdtype, uint64_t nbins, const double* uniforms_host, uint64_t nshots,
              int64_t* out_host, int mode, double* total_out) {
  if (!h || !probs || nbins == 0 || (rdtype != QB_F32 && rdtype != QB_F64) || (nshots && (!uniforms_host || !out_host)) ||
      (mode != QB_SCAN_EXACT && mode != QB_SCAN_PARALLEL))
    return fail(QB_ERR_INVALID, "bad sampling arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  size_t need = nbins * sizeof(double) + nshots * 16 + 16;
  if (h->cdf_bytes < need) {
    if (h->cdf) {
      QB_CUDA(cudaStreamSynchronize(h->stream));
      cudaFree(h->cdf);
      h->cdf = nullptr;
      h->cdf_bytes = 0;
    }
    QB_CUDA(cudaMalloc((void**)&h->cdf, need));
    h->cdf_bytes = need;
  }
  double* cdf = h->cdf;
  double* u_dev = cdf + nbins;
  long long* out_dev = (long long*)(u_dev + nshots);
  double* total_dev = (double*)(out_dev + nshots);
  int rc = cdf_locked(h, probs, rdtype, nbins, cdf, mode, total_dev);
  if (rc != QB_OK) return rc;
  if (total_out) QB_CUDA(cudaMemcpyAsync(total_out, total_dev, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (nshots) {
    QB_CUDA(cudaMemcpyAsync(u_dev, uniforms_host, nshots * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    k4_search<<<grid_for(nshots, 256, h->sm_count), 256, 0, h->stream>>>(cdf, nbins, u_dev, nshots, out_dev);
    QB_CHECK_LAUNCH("k4_search");
    QB_CUDA(cudaMemcpyAsync(out_host, out_dev, nshots * sizeof(int64_t), cudaMemcpyDeviceToHost, h->stream));
  }
  QB_CUDA(cudaStreamSynchronize(h->stream));
  return QB_OK;
}

// ---------------------------------------------------------------------------------------------------
// K5
// ---------------------------------------------------------------------------------------------------
int qb_collapse(qb_handle h, void* state, int nqubits, int dtype, const int* qubits, int nmeasured, uint64_t outcome,
                int normalize) {
  if (!h || !valid_state_args(state, nqubits, dtype)) return fail(QB_ERR_INVALID, "bad state arguments");
  if (nmeasured < 1 || nmeasured > nqubits || !qubits) return fail(QB_ERR_INVALID, "bad measured qubits");
  if (outcome >> nmeasured) return fail(QB_ERR_INVALID, "outcome out of range");
  uint64_t mask = 0, val = 0;
  std::vector<int> pos;
  for (int i = 0; i < nmeasured; ++i) {
    int q = qubits[i];
    if (q < 0 || q >= nqubits) return fail(QB_ERR_INVALID, "measured qubit out of range");
    int p = nqubits - 1 - q;
    if ((mask >> p) & 1) return fail(QB_ERR_INVALID, "repeated measured qubit");
    mask |= uint64_t(1) << p;
    if ((outcome >> (nmeasured - 1 - i)) & 1) val |= uint64_t(1) << p;
    pos.push_back(p);
  }
  std::sort(pos.begin(), pos.end());
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  double* norm_dev = nullptr;
  if (normalize) {
    int rc = slice_norm2_device(h, state, nqubits, dtype, pos, val, &norm_dev);
    if (rc != QB_OK) return rc;
  } else {
    int rc = ensure_scratch(h, 64);
    if (rc != QB_OK) return rc;
    norm_dev = (double*)h->scratch;
  }
  uint64_t count = uint64_t(1) << nqubits;
  int grid = grid_for(count, 256, h->sm_count);
  if (dtype == QB_C128)
    k5_project<double2><<<grid, 256, 0, h->stream>>>((double2*)state, count, mask, val, norm_dev, normalize);
  else
    k5_project<float2><<<grid, 256, 0, h->stream>>>((float2*)state, count, mask, val, norm_dev, normalize);
  QB_CHECK_LAUNCH("k5_project");
  return QB_OK;
}

// ---------------------------------------------------------------------------------------------------
// X1: density-matrix probabilities and collapse
// ---------------------------------------------------------------------------------------------------
int qb_probabilities_dm(qb_handle h, const void* rho, int nqubits, int dtype, const int* qubits, int nmeasured, void* probs_out) {
  if (!h || !rho || !probs_out || nqubits < 1 || 2 * nqubits > QB_MAX_QUBITS || (dtype != QB_C64 && dtype != QB_C128))
    return fail(QB_ERR_INVALID, "bad density-matrix arguments");
  if (nmeasured < 0 || nmeasured > nqubits || (nmeasured && !qubits)) return fail(QB_ERR_INVALID, "bad measured qubits");
  DmProbParams p;
  memset(&p, 0, sizeof(p));
  p.n = nqubits;
  p.m = nmeasured;
  uint64_t mmask = 0;
  for (int i = 0; i < nmeasured; ++i) {
    const int q = qubits[i];
    if (q < 0 || q >= nqubits) return fail(QB_ERR_INVALID, "measured qubit out of range");
    const int pos = nqubits - 1 - q;
    if ((mmask >> pos) & 1) return fail(QB_ERR_INVALID, "repeated measured qubit");
    mmask |= uint64_t(1) << pos;
    p.pos[i] = (uint8_t)pos;
  }
  p.umask = ((uint64_t(1) << nqubits) - 1) & ~mmask;
  p.n_u = nqubits - nmeasured;
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  const uint64_t nbins = uint64_t(1) << nmeasured;
  const unsigned grid = (unsigned)((nbins * 32 + 255) / 256);
  if (dtype == QB_C128) k3_probs_dm<double2, double><<<grid, 256, 0, h->stream>>>((const double2*)rho, (double*)probs_out, p);
  else k3_probs_dm<float2, float><<<grid, 256, 0, h->stream>>>((const float2*)rho, (float*)probs_out, p);
  QB_CHECK_LAUNCH("k3_probs_dm");
  return QB_OK;
}

int qb_collapse_dm(qb_handle h, void* rho, int nqubits, int dtype, const int* qubits, int nmeasured, uint64_t outcome, int normalize) {
  if (!h || !rho || nqubits < 1 || 2 * nqubits > QB_MAX_QUBITS || (dtype != QB_C64 && dtype != QB_C128))
    return fail(QB_ERR_INVALID, "bad density-matrix arguments");
  if (nmeasured < 1 || nmeasured > nqubits || !qubits) return fail(QB_ERR_INVALID, "bad measured qubits");
  if (outcome >> nmeasured) return fail(QB_ERR_INVALID, "outcome out of range");
  uint64_t mask = 0, val = 0;
  std::vector<int> pos;
  for (int i = 0; i < nmeasured; ++i) {
    const int q = qubits[i];
    if (q < 0 || q >= nqubits) return fail(QB_ERR_INVALID, "measured qubit out of range");
    const int p = nqubits - 1 - q;
    if ((mask >> p) & 1) return fail(QB_ERR_INVALID, "repeated measured qubit");
    mask |= uint64_t(1) << p;
    if ((outcome >> (nmeasured - 1 - i)) & 1) val |= uint64_t(1) << p;
    pos.push_back(p);
  }
  std::sort(pos.begin(), pos.end());
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  const uint64_t ngroups = uint64_t(1) << (nqubits - nmeasured);
  const int grid = grid_for(ngroups, RED_THREADS, h->sm_count, 8);
  int rc = ensure_scratch(h, (size_t)(2 * grid + 8) * sizeof(double));
  if (rc != QB_OK) return rc;
  double* partial = (double*)h->scratch + 8;
  double* trace = (double*)h->scratch;
  if (normalize) {
    InsertList ins;
    ins.n = nmeasured;
    for (int i = 0; i < nmeasured; ++i) ins.pos[i] = (uint8_t)pos[i];
    if (dtype == QB_C128) k5_diag_slice_sum<double2><<<grid, RED_THREADS, 0, h->stream>>>((const double2*)rho, nqubits, ngroups, ins, val, partial, partial + grid);
    else k5_diag_slice_sum<float2><<<grid, RED_THREADS, 0, h->stream>>>((const float2*)rho, nqubits, ngroups, ins, val, partial, partial + grid);
    QB_CHECK_LAUNCH("k5_diag_slice_sum");
    k9_sum_partials2<<<1, 32, 0, h->stream>>>(partial, partial + grid, grid, trace);
    QB_CHECK_LAUNCH("k9_sum_partials2");
  }
  const uint64_t count = uint64_t(1) << (2 * nqubits);
  const uint64_t mask2 = (mask << nqubits) | mask, val2 = (val << nqubits) | val;
  const int g2 = grid_for(count, 256, h->sm_count);
  if (dtype == QB_C128) k5_project_dm<double2><<<g2, 256, 0, h->stream>>>((double2*)rho, count, mask2, val2, trace, normalize);
  else k5_project_dm<float2><<<g2, 256, 0, h->stream>>>((float2*)rho, count, mask2, val2, trace, normalize);
  QB_CHECK_LAUNCH("k5_project_dm");
  return QB_OK;
}

// ---------------------------------------------------------------------------------------------------
// K7 plumbing (exchange kernels live in qb_sweep.cuh)
// ---------------------------------------------------------------------------------------------------
int qb_pack_half(qb_handle h, const void* state, int nqubits, int dtype, int local_qubit, int bit, void* staging) {
  if (!h || !valid_state_args(state, nqubits, dtype) || !staging) return fail(QB_ERR_INVALID, "bad state arguments");
  if (local_qubit < 0 || local_qubit >= nqubits || (bit != 0 && bit != 1)) return fail(QB_ERR_INVALID, "bad qubit");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  int rc = launch_half_copy(h->stream, h->sm_count, (void*)state, staging, nqubits, dtype, nqubits - 1 - local_qubit, bit, 0);
  if (rc != QB_OK) return cuda_fail(cudaGetLastError(), "pack_half");
  return QB_OK;
}

int qb_unpack_half(qb_handle h, void* state, int nqubits, int dtype, int local_qubit, int bit, const void* staging) {
  if (!h || !valid_state_args(state, nqubits, dtype) || !staging) return fail(QB_ERR_INVALID, "bad state arguments");
  if (local_qubit < 0 || local_qubit >= nqubits || (bit != 0 && bit != 1)) return fail(QB_ERR_INVALID, "bad qubit");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  int rc = launch_half_copy(h->stream, h->sm_count, state, (void*)staging, nqubits, dtype, nqubits - 1 - local_qubit, bit, 1);
  if (rc != QB_OK) return cuda_fail(cudaGetLastError(), "unpack_half");
  return QB_OK;
}

int qb_swap_half_p2p(qb_handle h, void* state, void* peer_state, int nqubits, int dtype, int local_qubit, int my_bit, int part,
                     int nparts) {
  if (!h || !valid_state_args(state, nqubits, dtype) || !peer_state) return fail(QB_ERR_INVALID, "bad state arguments");
  if (local_qubit < 0 || local_qubit >= nqubits || (my_bit != 0 && my_bit != 1) || nparts < 1 || part < 0 || part >= nparts)
    return fail(QB_ERR_INVALID, "bad exchange arguments");
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  int rc = launch_swap_half_p2p(h->stream, h->sm_count, state, peer_state, nqubits, dtype, nqubits - 1 - local_qubit, my_bit, part, nparts);
  if (rc != QB_OK) return cuda_fail(cudaGetLastError(), "swap_half_p2p");
  return QB_OK;
}

static int alltoall_common(qb_handle h, const void* state, int dtype, int npeers, void* const* peer_states, const uint64_t* my_offsets,
                           const uint64_t* peer_offsets, const uint64_t* begins, const uint64_t* ends, bool push) {
  if (!h || !state || (dtype != QB_C128 && dtype != QB_C64)) return fail(QB_ERR_INVALID, "bad state arguments");
  if (npeers < 1 || npeers > 8 || !peer_states || !my_offsets || !peer_offsets || !begins || !ends)
    return fail(QB_ERR_INVALID, "bad all-to-all arguments (1..8 entries)");
  A2ATable tab;
  memset(&tab, 0, sizeof(tab));
  tab.npeers = npeers;
  for (int i = 0; i < npeers; ++i) {
    if (!peer_states[i] || begins[i] > ends[i]) return fail(QB_ERR_INVALID, "bad all-to-all entry");
    tab.peer[i] = peer_states[i];
    tab.my_off[i] = my_offsets[i];
    tab.peer_off[i] = peer_offsets[i];
    tab.begin[i] = begins[i];
    tab.end[i] = ends[i];
  }
  std::lock_guard<std::mutex> lk(h->mu);
  DeviceGuard guard(h->device);
  int rc = push ? launch_alltoall_push(h->stream, h->sm_count, state, dtype, tab)
                : launch_alltoall_p2p(h->stream, h->sm_count, const_cast<void*>(state), dtype, tab);
  if (rc != QB_OK) return cuda_fail(cudaGetLastError(), push ? "alltoall_push" : "alltoall_p2p");
  return QB_OK;
}

int qb_alltoall_p2p(qb_handle h, void* state, int dtype, int npeers, void* const* peer_states, const uint64_t* my_offsets,
                    const uint64_t* peer_offsets, const uint64_t* begins, const uint64_t* ends) {
  return alltoall_common(h, state, dtype, npeers, peer_states, my_offsets, peer_offsets, begins, ends, false);
}

int qb_alltoall_push_p2p(qb_handle h, const void* state, int dtype, int nentries, void* const* dest_buffers, const uint64_t* my_offsets,
                         const uint64_t* dest_offsets, const uint64_t* begins, const uint64_t* ends) {
  return alltoall_common(h, state, dtype, nentries, dest_buffers, my_offsets, dest_offsets, begins, ends, true);
}

int qb_ipc_get_handle(qb_handle h, void* dptr, void* handle_out_64bytes) {
  if (!h || !dptr || !handle_out_64bytes) return fail(QB_ERR_INVALID, "null argument");
  DeviceGuard guard(h->device);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
  cudaIpcMemHandle_t hd;
  QB_CUDA(cudaIpcGetMemHandle(&hd, dptr));
  memcpy(handle_out_64bytes, &hd, 64);
  return QB_OK;
}

int qb_ipc_open_handle(qb_handle h, const void* handle_64bytes, void** dptr_out) {
  if (!h || !handle_64bytes || !dptr_out) return fail(QB_ERR_INVALID, "null argument");
  DeviceGuard guard(h->device);
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle_64bytes, 64);
  QB_CUDA(cudaIpcOpenMemHandle(dptr_out, hd, cudaIpcMemLazyEnablePeerAccess));
  return QB_OK;
}

int qb_ipc_close_handle(qb_handle h, void* dptr) {
  if (!h || !dptr) return fail(QB_ERR_INVALID, "null argument");
  DeviceGuard guard(h->device);
  QB_CUDA(cudaIpcCloseMemHandle(dptr));
  return QB_OK;
}


// extern "C" entry points of libmlstm_b200.so (declared in include/mlstm_b200.h):
// argument validation, kernel-family dispatch, error reporting.  No torch types.
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "mlstm_common.cuh"

namespace mlstm {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
void clear_error() { g_err[0] = 0; }

namespace {

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Null check for every family; for the tcgen05 family (bf16, TMA tensor maps over the caller's strided storage) also
// the layout the tensor maps need, so a bad view is reported by name here and not as a cuTensorMapEncodeTiled code later.
int check_act(const char* name, const mlstm_act& a, bool required, bool tma, const mlstm_params& p, int dh) {
  if (!a.ptr) {
    if (required) { set_error("%s: null pointer", name); return MLSTM_ERR_INVALID_ARG; }
    return MLSTM_OK;
  }
  if (!tma) return MLSTM_OK;
  if (!aligned16(a.ptr)) {
    set_error("%s: base pointer %p is not 16-byte aligned (TMA needs it; pass an aligned view or a copy)", name, a.ptr);
    return MLSTM_ERR_INVALID_ARG;
  }
  const int64_t st[3] = {a.stride_b, a.stride_h, a.stride_s};
  const int32_t ext[3] = {p.B, p.NH, p.S};
  const char* which[3] = {"batch", "head", "token"};
  for (int d = 0; d < 3; ++d) {
    if (ext[d] <= 1) continue;   // the stride of a size-1 dimension is never used
    if (st[d] % 8 != 0 || st[d] < dh) {
      set_error("%s: %s stride %lld elements must be a multiple of 8 and >= the head dim %d (bf16 TMA rows are 16-byte units)",
                name, which[d], (long long)st[d], dh);
      return MLSTM_ERR_INVALID_ARG;
    }
  }
  return MLSTM_OK;
}

int validate(const mlstm_params* p, int is_bwd) {
  if (!p) { set_error("params is NULL"); return MLSTM_ERR_INVALID_ARG; }
  if (p->abi_version != MLSTM_B200_ABI_VERSION) {
    set_error("abi_version %d != %d", p->abi_version, MLSTM_B200_ABI_VERSION);
    return MLSTM_ERR_INVALID_ARG;
  }
  if (p->B < 0 || p->NH <= 0 || p->S < 0 || p->DHQK <= 0 || p->DHV <= 0) {
    set_error("bad sizes B=%d NH=%d S=%d DHQK=%d DHV=%d", p->B, p->NH, p->S, p->DHQK, p->DHV);
    return MLSTM_ERR_INVALID_ARG;
  }
  if (p->dtype != MLSTM_F32 && p->dtype != MLSTM_BF16) {
    set_error("unsupported dtype %d", p->dtype);
    return MLSTM_ERR_UNSUPPORTED;
  }
  if (p->gate_mode != MLSTM_IGATE_EXP && p->gate_mode != MLSTM_IGATE_SIGMOID) {
    set_error("unknown gate_mode %d", p->gate_mode);
    return MLSTM_ERR_INVALID_ARG;
  }
  if (p->B == 0 || p->S == 0) return MLSTM_OK;  // empty input: nothing to do
  int rc;
  const bool tma = p->dtype == MLSTM_BF16 && tc_supported(*p);
  if ((rc = check_act("q", p->q, true, tma, *p, p->DHQK)) || (rc = check_act("k", p->k, true, tma, *p, p->DHQK)) ||
      (rc = check_act("v", p->v, true, tma, *p, p->DHV)) || (rc = check_act("h", p->h, true, tma, *p, p->DHV)))
    return rc;
  if (!p->i.ptr || !p->f.ptr) { set_error("gate pre-activations i/f: null pointer"); return MLSTM_ERR_INVALID_ARG; }
  if ((p->c_last != nullptr) != (p->n_last != nullptr) || (p->c_last != nullptr) != (p->m_last != nullptr)) {
    set_error("c_last/n_last/m_last must be given together");
    return MLSTM_ERR_INVALID_ARG;
  }
  if ((p->n_row != nullptr) != (p->m_row != nullptr)) {
    set_error("n_row/m_row must be given together");
    return MLSTM_ERR_INVALID_ARG;
  }
  if (is_bwd) {
    if ((rc = check_act("dh", p->dh, true, tma, *p, p->DHV)) || (rc = check_act("dq", p->dq, true, tma, *p, p->DHQK)) ||
        (rc = check_act("dk", p->dk, true, tma, *p, p->DHQK)) || (rc = check_act("dv", p->dv, true, tma, *p, p->DHV)))
      return rc;
    if (!p->di.ptr || !p->df.ptr) { set_error("di/df: null pointer"); return MLSTM_ERR_INVALID_ARG; }
    if (!p->n_row || !p->m_row) { set_error("backward needs n_row and m_row from the forward"); return MLSTM_ERR_INVALID_ARG; }
    size_t need = mlstm_b200_workspace_bytes(p, 1);
    if (need && (!p->workspace || p->workspace_bytes < need)) {
      set_error("workspace too small: have %zu, need %zu", p->workspace ? p->workspace_bytes : (size_t)0, need);
      return MLSTM_ERR_WORKSPACE;
    }
    if (need && !aligned16(p->workspace)) { set_error("workspace must be 16-byte aligned"); return MLSTM_ERR_INVALID_ARG; }
  }
  return MLSTM_OK;
}

}  // namespace

// Bind the device that owns the buffers to the calling thread.  PyTorch runs the backward on
// an autograd worker thread that may not have a current CUDA context yet, and the driver-side
// tensor-map encoder needs one.
int bind_device(const void* dev_ptr) {
  cudaPointerAttributes attr;
  cudaError_t e = cudaPointerGetAttributes(&attr, dev_ptr);
  if (e != cudaSuccess || attr.type != cudaMemoryTypeDevice) {
    (void)cudaGetLastError();
    set_error("q does not point to device memory (%s)", e == cudaSuccess ? "host/unregistered pointer" : cudaGetErrorString(e));
    return MLSTM_ERR_INVALID_ARG;
  }
  // Once per thread and device: later calls only check that the device is still current.  (cudaFree is not allowed while a
  // stream of the process is being captured into a CUDA graph; with the binding cached, forward and backward are plain kernel
  // launches and can be captured.)
  thread_local int bound = -1;
  int cur = -1;
  if (bound == attr.device && cudaGetDevice(&cur) == cudaSuccess && cur == attr.device) return MLSTM_OK;
  e = cudaSetDevice(attr.device);
  if (e == cudaSuccess) e = cudaFree(nullptr);   // forces the primary context current on this thread
  if (e == cudaSuccess) bound = attr.device;
  if (e != cudaSuccess) {
    set_error("cudaSetDevice(%d): %s", attr.device, cudaGetErrorString(e));
    return MLSTM_ERR_CUDA;
  }
  return MLSTM_OK;
}

namespace {

enum Family { FAM_NONE = 0, FAM_SIMT = 1, FAM_TC = 2 };

Family pick(const mlstm_params& p) {
  if (p.dtype == MLSTM_BF16 && tc_supported(p)) return FAM_TC;
  if (simt_supported(p)) return FAM_SIMT;
  return FAM_NONE;
}

}  // namespace
}  // namespace mlstm

using namespace mlstm;

extern "C" {

int mlstm_b200_abi_version(void) { return MLSTM_B200_ABI_VERSION; }

size_t mlstm_b200_workspace_bytes(const mlstm_params* p, int is_backward) {
  if (!p || p->B <= 0 || p->S <= 0) return 0;
  if (!is_backward) return 0;
  switch (pick(*p)) {
    case FAM_TC: return tc_bwd_workspace(*p);
    case FAM_SIMT: return simt_bwd_workspace(*p);
    default: return 0;
  }
}

size_t mlstm_b200_state_bytes(const mlstm_params* p) {
  if (!p || p->B <= 0 || p->S <= 0) return 0;
  return pick(*p) == FAM_TC ? tc_state_bytes(*p) : 0;
}

const char* mlstm_b200_kernel_name(const mlstm_params* p, int /*is_backward*/) {
  if (!p) return nullptr;
  switch (pick(*p)) {
    case FAM_TC: return "tcgen05";
    case FAM_SIMT: return "simt";
    default: return nullptr;
  }
}

const char* mlstm_b200_kernel_variant(const mlstm_params* p, int is_backward) {
  if (!p) return nullptr;
  switch (pick(*p)) {
    case FAM_TC:
      if (is_backward) return tc_use_fused_bwd(*p) ? "fused_walk" : (tc_use_single_pass_bwd(*p) ? "single_pass" : "chunk_parallel");
      return tc_use_two_phase(*p) ? "two_phase" : "single_pass";
    case FAM_SIMT: return "simt";
    default: return nullptr;
  }
}

int mlstm_b200_fwd(const mlstm_params* p, void* cuda_stream) {
  g_err[0] = 0;
  int rc = validate(p, 0);
  if (rc) return rc;
  if (p->B == 0 || p->S == 0) return MLSTM_OK;
  if ((rc = bind_device(p->q.ptr))) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  switch (pick(*p)) {
    case FAM_TC: return tc_fwd(*p, st);
    case FAM_SIMT: return simt_fwd(*p, st);
    default:
      set_error("no kernel for dtype=%d DHQK=%d DHV=%d", p->dtype, p->DHQK, p->DHV);
      return MLSTM_ERR_UNSUPPORTED;
  }
}

static int bwd_impl(const mlstm_params* p, int part, void* cuda_stream) {
  g_err[0] = 0;
  int rc = validate(p, 1);
  if (rc) return rc;
  if (part < -1 || part > 1) { set_error("bad backward part %d", part); return MLSTM_ERR_INVALID_ARG; }
  if (p->B == 0 || p->S == 0) return MLSTM_OK;
  if ((rc = bind_device(p->q.ptr))) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  switch (pick(*p)) {
    case FAM_TC: return tc_bwd(*p, st, part);
    case FAM_SIMT: return simt_bwd(*p, st, part);
    default:
      set_error("no kernel for dtype=%d DHQK=%d DHV=%d", p->dtype, p->DHQK, p->DHV);
      return MLSTM_ERR_UNSUPPORTED;
  }
}

int mlstm_b200_bwd(const mlstm_params* p, void* cuda_stream) { return bwd_impl(p, -1, cuda_stream); }

int mlstm_b200_bwd_part(const mlstm_params* p, int part, void* cuda_stream) { return bwd_impl(p, part, cuda_stream); }

uint64_t mlstm_b200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

const char* mlstm_b200_last_error(void) { return g_err; }

}  // extern "C"

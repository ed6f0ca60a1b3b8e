// Host-side construction of TMA tensor maps for (B, NH, S, DH) activation tensors with
// arbitrary B/NH/S strides (innermost DH contiguous).  The driver entry point is resolved at
// run time through cudaGetDriverEntryPoint, so the library does not link against libcuda.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <mutex>

namespace mlstm {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled get_tmap_encoder() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  return fn;
}

// A tensor map is a pure function of (pointer, extents, strides, box): encoding one costs a driver call, and a
// forward + backward needs up to ten of them.  PyTorch's caching allocator hands the same blocks back step after
// step, so the encoded maps are kept in a small direct-mapped cache keyed on everything that enters the encoding
// (no driver-side state is attached to a map: reusing one for a re-allocated block at the same address is sound).
struct TmapKey {
  const void* ptr;
  int64_t s0, s1, s2;
  int32_t d0, d1, d2, d3, box_rows, kind;
};
struct TmapSlot { TmapKey key; CUtensorMap map; bool used; };
constexpr int TMAP_CACHE_SLOTS = 1024;
inline bool tmap_cache_lookup(const TmapKey& k, CUtensorMap* out, bool store) {
  static TmapSlot slots[TMAP_CACHE_SLOTS];
  static std::mutex mu;
  uint64_t hsh = (uint64_t)(uintptr_t)k.ptr * 0x9E3779B97F4A7C15ull;
  hsh ^= (uint64_t)k.kind * 0xC2B2AE3D27D4EB4Full + (uint64_t)k.box_rows * 0x165667B19E3779F9ull + (uint64_t)k.d1;
  TmapSlot& s = slots[(hsh >> 20) & (TMAP_CACHE_SLOTS - 1)];
  std::lock_guard<std::mutex> g(mu);
  if (store) { s.key = k; s.map = *out; s.used = true; return true; }
  if (s.used && memcmp(&s.key, &k, sizeof(TmapKey)) == 0) { *out = s.map; return true; }
  return false;
}

// Narrow head dims without copies (mlstm_api.cu, zero-padded problems): the caller registers "this tensor's rows really hold
// `extent` elements" for the duration of a launch sequence; make_act_tmap then encodes that extent as the innermost dimension
// while the kernels keep asking for 64-column boxes — TMA reads the missing columns as zeros and clips them on store, which
// is exactly the zero padding, at no HBM cost.  Per thread (the launch sequence runs on the calling thread).
struct ExtentOverride { const void* ptr; int extent; };
constexpr int MAX_EXTENT_OVERRIDES = 8;
inline ExtentOverride* extent_overrides() {
  static thread_local ExtentOverride t[MAX_EXTENT_OVERRIDES];
  return t;
}
inline int true_extent(const void* ptr, int DH) {
  const ExtentOverride* t = extent_overrides();
  for (int i = 0; i < MAX_EXTENT_OVERRIDES; ++i)
    if (t[i].ptr == ptr && ptr != nullptr) return t[i].extent;
  return DH;
}
struct ExtentScope {   // RAII: the registrations of one padded call
  ExtentScope() { clear(); }
  ~ExtentScope() { clear(); }
  void add(const void* ptr, int extent) {
    ExtentOverride* t = extent_overrides();
    for (int i = 0; i < MAX_EXTENT_OVERRIDES; ++i)
      if (!t[i].ptr) { t[i].ptr = ptr; t[i].extent = extent; return; }
  }
  static void clear() {
    ExtentOverride* t = extent_overrides();
    for (int i = 0; i < MAX_EXTENT_OVERRIDES; ++i) { t[i].ptr = nullptr; t[i].extent = 0; }
  }
};

// bf16 tensor (B, NH, S, DH); strides in elements.  Box = 64 (DH) x box_rows (S) x 1 x 1,
// 128-byte swizzle, out-of-bounds rows read as zero / are clipped on store.
// Returns 0 on success, else the CUresult (or -1 if the encoder is unavailable).
inline int make_act_tmap(CUtensorMap* out, const void* ptr, int B, int NH, int S, int DH, int64_t stride_b,
                         int64_t stride_h, int64_t stride_s, int box_rows) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.s0 = stride_s; key.s1 = stride_h; key.s2 = stride_b;
  const int ext = true_extent(ptr, DH);   // < DH: columns [ext, DH) are out of bounds = zeros (see ExtentOverride)
  key.d0 = ext; key.d1 = S; key.d2 = NH; key.d3 = B; key.box_rows = box_rows; key.kind = 4;
  if (tmap_cache_lookup(key, out, false)) return 0;
  PFN_tmapEncodeTiled enc = get_tmap_encoder();
  if (!enc) return -1;
  cuuint64_t dims[4] = {(cuuint64_t)ext, (cuuint64_t)S, (cuuint64_t)NH, (cuuint64_t)B};
  // size-1 dimensions may come with arbitrary strides; give them a legal one
  int64_t ss = (S > 1) ? stride_s : (int64_t)DH;
  int64_t sh = (NH > 1) ? stride_h : (int64_t)DH * S;
  int64_t sb = (B > 1) ? stride_b : (int64_t)DH * S * NH;
  cuuint64_t strides[3] = {(cuuint64_t)ss * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
  cuuint32_t box[4] = {64u, (cuuint32_t)box_rows, 1u, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == CUDA_SUCCESS) tmap_cache_lookup(key, out, true);
  return (int)r;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device, size) instead of on every launch.
inline cudaError_t set_max_smem_once(const void* kernel, size_t smem) {
  struct Slot { const void* fn; int dev; size_t smem; };
  static Slot done[128];
  static int n_done = 0;
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> g(mu);
    for (int i = 0; i < n_done; ++i)
      if (done[i].fn == kernel && done[i].dev == dev && done[i].smem >= smem) return cudaSuccess;
  }
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e == cudaSuccess) {
    std::lock_guard<std::mutex> g(mu);
    if (n_done < 128) done[n_done++] = Slot{kernel, dev, smem};
  }
  return e;
}

// SM count of the device that owns `dev_ptr` (the current device if the pointer is NULL or unknown): the kernel
// variant must not depend on which device happens to be current when a size query is made.
inline int sm_count_of(const void* dev_ptr) {
  static int cached[64];
  int dev = -1;
  if (dev_ptr) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, dev_ptr) == cudaSuccess && attr.type == cudaMemoryTypeDevice) dev = attr.device;
    else (void)cudaGetLastError();
  }
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return 148; }
  if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev];
  int sms = 148;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { (void)cudaGetLastError(); return 148; }
  if (dev >= 0 && dev < 64) cached[dev] = sms;
  return sms;
}

}  // namespace mlstm

// Stand-alone probe: tcgen05.mma issue/execute rate for the operand layouts the mLSTM kernels use.
// One CTA, tiles hold arbitrary data (timing only).  Prints cycles for NREP x 8 MMAs of
// M=128,N=128,K=16 per (A major, B major) combination, and for N=16.
#include <cstdio>
#include "tc_ptx.cuh"
using namespace mlstm::ptx;
constexpr int TILE = 128 * 128;
__global__ void __launch_bounds__(128) rate_kernel(long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 6 * TILE / 4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0x3c003c00u;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_base_s;
  uint32_t phase = 0;
  if (tid == 0) {
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem + 2 * TILE);
    int slot = 0;
    for (int amn = 0; amn < 2; ++amn) for (int bmn = 0; bmn < 2; ++bmn) for (int n : {128, 64, 16}) {
      const uint32_t idesc = make_idesc_bf16(128, n, amn, bmn);
      uint64_t a[8], b[8];   // descriptors pre-built, as in the kernels: the loop below times issue + execution only
      for (int ks = 0; ks < 8; ++ks) {
        a[ks] = amn ? make_sdesc(a0 + ks * 2048, TILE, 1024) : make_sdesc(a0 + (ks >> 2) * TILE + (ks & 3) * 32, 16, 1024);
        b[ks] = bmn ? make_sdesc(b0 + ks * 2048, TILE, 1024) : make_sdesc(b0 + (ks >> 2) * TILE + (ks & 3) * 32, 16, 1024);
      }
      long long t0 = clock64();
#pragma unroll
      for (int rep = 0; rep < 4; ++rep)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) umma_bf16_ss(tm, a[ks], b[ks], idesc, (rep | ks) > 0);
      long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, phase); phase ^= 1;
      long long t2 = clock64();
      out[slot * 2] = t1 - t0; out[slot * 2 + 1] = t2 - t0; ++slot;
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}
int main() {
  long long* d; cudaMalloc(&d, 64 * 8);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * TILE);
  for (int it = 0; it < 2; ++it) rate_kernel<<<1, 128, 6 * TILE>>>(d);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
  long long h[64]; cudaMemcpy(h, d, 24 * 8, cudaMemcpyDeviceToHost);
  int slot = 0;
  for (int amn = 0; amn < 2; ++amn) for (int bmn = 0; bmn < 2; ++bmn) for (int n : {128, 64, 16}) {
    printf("A %s  B %s  N=%3d : issue %6lld cyc, complete %6lld cyc for 32 MMAs -> %.1f cyc/MMA\n", amn ? "MN" : "K ", bmn ? "MN" : "K ", n,
           h[slot * 2], h[slot * 2 + 1], h[slot * 2 + 1] / 32.0);
    ++slot;
  }
  return 0;
}

// Stand-alone probe (not part of the library): tcgen05.mma with the A operand in TMEM ("TS" form).
//   T1  S = Q K^T (SS form, reference for the rest)
//   T2  P = bf16(tril(S/16)) written with tcgen05.st (packed bf16x2, 64 columns) into a SEPARATE TMEM
//       region, then H = P V with A = [tmem]
//   T3  the same with P written over S's own first 64 columns (in-place aliasing, after a CTA barrier)
//   timing: 8 MMAs (K = 128) SS vs TS, clock64 around issue..commit-wait
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_ts probe_ts.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_ptx.cuh"
#include "tc_tmap.cuh"

using namespace mlstm;
using namespace mlstm::ptx;

constexpr int S = 128, NH = 2, DH = 128, HEAD = 1;
constexpr int TILE = 128 * 128;

struct Maps { CUtensorMap q, k, v; };

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ Maps maps, float* outS, float* outH,
                                                     float* outH2, long long* cyc) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 2 * TILE;
  uint8_t* sV = smem + 4 * TILE;
  uint8_t* sP = smem + 6 * TILE;
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar_tma, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s, tS = tm, tH = tm + 128, tP = tm + 256, tH2 = tm + 384;
  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_tma, 6 * TILE);
    for (int kb = 0; kb < 2; ++kb) {
      tma_load_4d(sQ + kb * TILE, &maps.q, &bar_tma, kb * 64, 0, HEAD, 0);
      tma_load_4d(sK + kb * TILE, &maps.k, &bar_tma, kb * 64, 0, HEAD, 0);
      tma_load_4d(sV + kb * TILE, &maps.v, &bar_tma, kb * 64, 0, HEAD, 0);
    }
  }
  mbar_wait(&bar_tma, 0);
  uint32_t ph = 0;
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t off = (ks >> 2) * TILE + (ks & 3) * 32;
      umma_bf16_ss(tS, make_sdesc(smem_u32(sQ) + off, 16, 1024), make_sdesc(smem_u32(sK) + off, 16, 1024), idesc, ks > 0);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, ph); ph ^= 1;
  tc_fence_after();
  const int t = tid;
  const uint32_t la = (uint32_t)(warp * 32) << 16;
  uint32_t packed[4][16];
  for (int cb = 0; cb < 4; ++cb) {
    float r[32];
    tmem_ld32(tS + la + cb * 32, r);
    tmem_ld_wait();
    for (int c = 0; c < 32; ++c) outS[t * 128 + cb * 32 + c] = r[c];
    for (int c = 0; c < 32; c += 2) {
      const int j0 = cb * 32 + c;
      const float a = (j0 <= t) ? r[c] * 0.0625f : 0.f, b = (j0 + 1 <= t) ? r[c + 1] * 0.0625f : 0.f;
      packed[cb][c / 2] = pack_bf16x2(a, b);
    }
    // SS copy of P for the timing comparison
    for (int c = 0; c < 32; c += 8) {
      const int j = cb * 32 + c;
      *reinterpret_cast<uint4*>(sP + (j >> 6) * TILE + swz128(t, j & 63)) =
          make_uint4(packed[cb][c / 2], packed[cb][c / 2 + 1], packed[cb][c / 2 + 2], packed[cb][c / 2 + 3]);
    }
  }
  // T2: separate region
  for (int cb = 0; cb < 4; ++cb) tmem_st16(tP + la + cb * 16, packed[cb]);
  tmem_st_wait();
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  const uint32_t idH = make_idesc_bf16(128, 128, 0, 1);   // A K-major, B MN-major (V)
  long long c0 = 0, c1 = 0, c2 = 0;
  if (tid == 0) {
    tc_fence_after();
    c0 = clock64();
    for (int ks = 0; ks < 8; ++ks) umma_bf16_ts(tH, tP + ks * 8, make_sdesc(smem_u32(sV) + ks * 2048, TILE, 1024), idH, ks > 0);
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, ph); ph ^= 1;
  tc_fence_after();
  if (tid == 0) c1 = clock64();
  for (int cb = 0; cb < 4; ++cb) {
    float r[32];
    tmem_ld32(tH + la + cb * 32, r);
    tmem_ld_wait();
    for (int c = 0; c < 32; ++c) outH[t * 128 + cb * 32 + c] = r[c];
  }
  // T3: P over S's own columns [0, 64)
  tc_fence_before();
  __syncthreads();
  for (int cb = 0; cb < 4; ++cb) tmem_st16(tS + la + cb * 16, packed[cb]);
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    for (int ks = 0; ks < 8; ++ks) umma_bf16_ts(tH2, tS + ks * 8, make_sdesc(smem_u32(sV) + ks * 2048, TILE, 1024), idH, ks > 0);
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, ph); ph ^= 1;
  tc_fence_after();
  for (int cb = 0; cb < 4; ++cb) {
    float r[32];
    tmem_ld32(tH2 + la + cb * 32, r);
    tmem_ld_wait();
    for (int c = 0; c < 32; ++c) outH2[t * 128 + cb * 32 + c] = r[c];
  }
  // timing of the SS form of the same product
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    c2 = clock64();
    for (int ks = 0; ks < 8; ++ks)
      umma_bf16_ss(tH2, make_sdesc(smem_u32(sP) + (ks >> 2) * TILE + (ks & 3) * 32, 16, 1024),
                   make_sdesc(smem_u32(sV) + ks * 2048, TILE, 1024), idH, ks > 0);
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, ph); ph ^= 1;
  tc_fence_after();
  if (tid == 0) { cyc[0] = c1 - c0; cyc[1] = clock64() - c2; }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)
static void report(const char* name, const std::vector<float>& got, const std::vector<double>& ref) {
  double me = 0, mr = 0;
  for (size_t i = 0; i < ref.size(); ++i) { me = fmax(me, fabs(got[i] - ref[i])); mr = fmax(mr, fabs(ref[i])); }
  printf("%-4s max_abs_err %.4e  max_ref %.4e  rel %.3e  %s\n", name, me, mr, me / mr, (me / mr < 1e-3) ? "OK" : "FAIL");
}

int main() {
  const size_t n = (size_t)S * NH * DH;
  std::vector<__nv_bfloat16> hq(n), hk(n), hv(n);
  std::vector<float> Q(S * DH), K(S * DH), V(S * DH);
  srand(1);
  for (int s = 0; s < S; ++s)
    for (int h = 0; h < NH; ++h)
      for (int d = 0; d < DH; ++d) {
        float a = (rand() % 2001 - 1000) / 1000.f, b = (rand() % 2001 - 1000) / 1000.f, c = (rand() % 2001 - 1000) / 1000.f;
        size_t idx = ((size_t)s * NH + h) * DH + d;
        hq[idx] = __float2bfloat16_rn(a); hk[idx] = __float2bfloat16_rn(b); hv[idx] = __float2bfloat16_rn(c);
        if (h == HEAD) { Q[s * DH + d] = bf(a); K[s * DH + d] = bf(b); V[s * DH + d] = bf(c); }
      }
  __nv_bfloat16 *dq, *dk, *dv;
  float *oS, *oH, *oH2;
  long long* dc;
  CK(cudaMalloc(&dq, n * 2)); CK(cudaMalloc(&dk, n * 2)); CK(cudaMalloc(&dv, n * 2)); CK(cudaMalloc(&dc, 16));
  CK(cudaMemcpy(dq, hq.data(), n * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dk, hk.data(), n * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv, hv.data(), n * 2, cudaMemcpyHostToDevice));
  for (float** p : {&oS, &oH, &oH2}) CK(cudaMalloc(p, 128 * 128 * 4));
  Maps maps;
  int r = 0;
  r |= make_act_tmap(&maps.q, dq, 1, NH, S, DH, (int64_t)S * NH * DH, DH, NH * DH, 128);
  r |= make_act_tmap(&maps.k, dk, 1, NH, S, DH, (int64_t)S * NH * DH, DH, NH * DH, 128);
  r |= make_act_tmap(&maps.v, dv, 1, NH, S, DH, (int64_t)S * NH * DH, DH, NH * DH, 128);
  if (r) { printf("tensor map encode failed: %d\n", r); return 3; }
  const int smem = 8 * TILE;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<1, 128, smem>>>(maps, oS, oH, oH2, dc);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> gS(128 * 128), gH(128 * 128), gH2(128 * 128);
  long long cyc[2];
  CK(cudaMemcpy(gS.data(), oS, 128 * 128 * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gH.data(), oH, 128 * 128 * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gH2.data(), oH2, 128 * 128 * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cyc, dc, 16, cudaMemcpyDeviceToHost));
  std::vector<double> rS(128 * 128), rH(128 * 128);
  std::vector<float> P(128 * 128);
  for (int t = 0; t < 128; ++t)
    for (int j = 0; j < 128; ++j) {
      double a = 0;
      for (int d = 0; d < DH; ++d) a += (double)Q[t * DH + d] * K[j * DH + d];
      rS[t * 128 + j] = a;
      P[t * 128 + j] = (j <= t) ? bf((float)a * 0.0625f) : 0.f;
    }
  for (int t = 0; t < 128; ++t)
    for (int d = 0; d < 128; ++d) {
      double a = 0;
      for (int j = 0; j < 128; ++j) a += (double)P[t * 128 + j] * V[j * DH + d];
      rH[t * 128 + d] = a;
    }
  report("T1", gS, rS);
  report("T2", gH, rH);
  report("T3", gH2, rH);
  printf("cycles issue..done, 8 MMAs K=128 N=128: TS %lld  SS %lld\n", cyc[0], cyc[1]);
  return 0;
}

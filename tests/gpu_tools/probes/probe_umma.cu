// Stand-alone probe (not part of the library): checks every tcgen05 / TMA building block the
// mLSTM kernels rely on against a host reference, one small case each:
//   T1  S  = Q K^T            K-major A, K-major B (TMA SWIZZLE_128B tiles)
//   T2  H  = P V              P written by threads into a swizzled K-major tile after a
//                             TMEM load; V as MN-major B (two 64-wide N blocks, LBO)
//   T3  C  = X + K^T V        tcgen05.st of X then accumulating MMA; MN-major A and B
//   T4  G  = Q Cb             Cb = bf16(C/64) written by threads as an MN-major B tile
//   T5  Gt = V Cb^T           the same Cb buffer read as a K-major B operand
//   T6  TMA store of bf16(H) from a swizzled staging tile into (B,S,NH,DH) storage
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_umma probe_umma.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tc_ptx.cuh"
#include "tc_tmap.cuh"

using namespace mlstm;
using namespace mlstm::ptx;

constexpr int S = 128, NH = 2, DH = 128, HEAD = 1;
constexpr int TILE = 128 * 128;  // bytes of one [128][64] bf16 tile = 16384

struct Maps { CUtensorMap q, k, v, hout; };

__global__ void __launch_bounds__(128) probe_kernel(const __grid_constant__ Maps maps, float* outS, float* outH,
                                                     float* outC, float* outG, float* outGt) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                 // 2 tiles
  uint8_t* sK = smem + 2 * TILE;      // 2 tiles
  uint8_t* sV = smem + 4 * TILE;      // 2 tiles
  uint8_t* sP = smem + 6 * TILE;      // 2 tiles (P, later staging for H)
  uint8_t* sC = smem + 8 * TILE;      // 2 tiles (Cb)
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&bar_tma, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tmem_base_s;
  const uint32_t tS = tm, tH = tm + 128, tC = tm + 256, tG = tm + 384;

  if (tid == 0) {
    mbar_arrive_expect_tx(&bar_tma, 6 * TILE);
    for (int kb = 0; kb < 2; ++kb) {
      tma_load_4d(sQ + kb * TILE, &maps.q, &bar_tma, kb * 64, 0, HEAD, 0);
      tma_load_4d(sK + kb * TILE, &maps.k, &bar_tma, kb * 64, 0, HEAD, 0);
      tma_load_4d(sV + kb * TILE, &maps.v, &bar_tma, kb * 64, 0, HEAD, 0);
    }
  }
  mbar_wait(&bar_tma, 0);
  uint32_t mma_phase = 0;

  // ---- T1: S = Q K^T -------------------------------------------------------------------
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t off = (ks >> 2) * TILE + (ks & 3) * 32;
      umma_bf16_ss(tS, make_sdesc(smem_u32(sQ) + off, 16, 1024), make_sdesc(smem_u32(sK) + off, 16, 1024), idesc, ks > 0);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, mma_phase); mma_phase ^= 1;
  tc_fence_after();
  {
    const int t = tid;  // row == TMEM lane
    const uint32_t lane_addr = tS + ((uint32_t)(warp * 32) << 16);
    for (int cb = 0; cb < 4; ++cb) {
      float r[32];
      tmem_ld32(lane_addr + cb * 32, r);
      tmem_ld_wait();
      for (int c = 0; c < 32; ++c) outS[t * 128 + cb * 32 + c] = r[c];
      // P = tril(S / 16) as bf16 into the swizzled K-major tile
      for (int c = 0; c < 32; c += 8) {
        uint32_t w[4];
        for (int e = 0; e < 4; ++e) {
          int j0 = cb * 32 + c + 2 * e;
          float a = (j0 <= t) ? r[c + 2 * e] * 0.0625f : 0.f;
          float b = (j0 + 1 <= t) ? r[c + 2 * e + 1] * 0.0625f : 0.f;
          w[e] = pack_bf16x2(a, b);
        }
        int j = cb * 32 + c;
        uint8_t* dst = sP + (j >> 6) * TILE + swz128(t, j & 63);
        *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();

  // ---- T2: H = P V ; T3: C = X + K^T V -------------------------------------------------
  {  // X into TMEM: X[dk][dv] = 0.5*dk - 0.25*dv
    const uint32_t lane_addr = tC + ((uint32_t)(warp * 32) << 16);
    for (int cb = 0; cb < 4; ++cb) {
      float r[32];
      for (int c = 0; c < 32; ++c) r[c] = 0.5f * tid - 0.25f * (cb * 32 + c);
      tmem_st32(lane_addr + cb * 32, r);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idH = make_idesc_bf16(128, 128, 0, 1);   // A K-major (P), B MN-major (V)
    for (int ks = 0; ks < 8; ++ks) {
      uint64_t a = make_sdesc(smem_u32(sP) + (ks >> 2) * TILE + (ks & 3) * 32, 16, 1024);
      uint64_t b = make_sdesc(smem_u32(sV) + ks * 2048, TILE, 1024);  // 16 j-rows per k-step
      umma_bf16_ss(tH, a, b, idH, ks > 0);
    }
    const uint32_t idC = make_idesc_bf16(128, 128, 1, 1);   // A MN-major (K^T), B MN-major (V)
    for (int ks = 0; ks < 8; ++ks) {
      uint64_t a = make_sdesc(smem_u32(sK) + ks * 2048, TILE, 1024);
      uint64_t b = make_sdesc(smem_u32(sV) + ks * 2048, TILE, 1024);
      umma_bf16_ss(tC, a, b, idC, 1);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, mma_phase); mma_phase ^= 1;
  tc_fence_after();
  {
    const int t = tid;
    const uint32_t la = ((uint32_t)(warp * 32) << 16);
    for (int cb = 0; cb < 4; ++cb) {
      float r[32];
      tmem_ld32(tH + la + cb * 32, r);
      tmem_ld_wait();
      for (int c = 0; c < 32; ++c) outH[t * 128 + cb * 32 + c] = r[c];
      // stage bf16(H) for the TMA store (reuses the P tiles: the MMA reading them is done)
      for (int c = 0; c < 32; c += 8) {
        int j = cb * 32 + c;
        uint8_t* dst = sP + (j >> 6) * TILE + swz128(t, j & 63);
        *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(r[c], r[c + 1]), pack_bf16x2(r[c + 2], r[c + 3]),
                                                    pack_bf16x2(r[c + 4], r[c + 5]), pack_bf16x2(r[c + 6], r[c + 7]));
      }
      float q[32];
      tmem_ld32(tC + la + cb * 32, q);
      tmem_ld_wait();
      for (int c = 0; c < 32; ++c) outC[t * 128 + cb * 32 + c] = q[c];
      // Cb = bf16(C / 64): row dk = t, 64-wide dv blocks -> MN-major B tile [dk][dv]
      for (int c = 0; c < 32; c += 8) {
        int j = cb * 32 + c;
        uint8_t* dst = sC + (j >> 6) * TILE + swz128(t, j & 63);
        const float s = 1.f / 64.f;
        *reinterpret_cast<uint4*>(dst) =
            make_uint4(pack_bf16x2(q[c] * s, q[c + 1] * s), pack_bf16x2(q[c + 2] * s, q[c + 3] * s),
                       pack_bf16x2(q[c + 4] * s, q[c + 5] * s), pack_bf16x2(q[c + 6] * s, q[c + 7] * s));
      }
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();

  // ---- T6: TMA store of the staged bf16(H) ---------------------------------------------
  if (tid == 0) {
    tma_store_4d(&maps.hout, sP, 0, 0, HEAD, 0);
    tma_store_4d(&maps.hout, sP + TILE, 64, 0, HEAD, 0);
    tma_store_commit();
  }
  // ---- T4: G = Q Cb (Cb MN-major) ; T5: Gt = V Cb^T (Cb K-major) ------------------------
  if (tid == 0) {
    tc_fence_after();
    const uint32_t idG = make_idesc_bf16(128, 128, 0, 1);
    for (int ks = 0; ks < 8; ++ks) {
      uint64_t a = make_sdesc(smem_u32(sQ) + (ks >> 2) * TILE + (ks & 3) * 32, 16, 1024);
      uint64_t b = make_sdesc(smem_u32(sC) + ks * 2048, TILE, 1024);
      umma_bf16_ss(tG, a, b, idG, ks > 0);
    }
    const uint32_t idGt = make_idesc_bf16(128, 128, 0, 0);
    for (int ks = 0; ks < 8; ++ks) {
      uint32_t off = (ks >> 2) * TILE + (ks & 3) * 32;
      uint64_t a = make_sdesc(smem_u32(sV) + off, 16, 1024);
      uint64_t b = make_sdesc(smem_u32(sC) + off, 16, 1024);
      umma_bf16_ss(tS, a, b, idGt, ks > 0);
    }
    umma_commit(&bar_mma);
  }
  mbar_wait(&bar_mma, mma_phase); mma_phase ^= 1;
  tc_fence_after();
  {
    const int t = tid;
    const uint32_t la = ((uint32_t)(warp * 32) << 16);
    for (int cb = 0; cb < 4; ++cb) {
      float r[32];
      tmem_ld32(tG + la + cb * 32, r);
      tmem_ld_wait();
      for (int c = 0; c < 32; ++c) outG[t * 128 + cb * 32 + c] = r[c];
      tmem_ld32(tS + la + cb * 32, r);
      tmem_ld_wait();
      for (int c = 0; c < 32; ++c) outGt[t * 128 + cb * 32 + c] = r[c];
    }
  }
  if (tid == 0) tma_store_wait_all<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 2; } } while (0)

static void report(const char* name, const std::vector<float>& got, const std::vector<double>& ref) {
  double me = 0, mr = 0;
  for (size_t i = 0; i < ref.size(); ++i) { me = fmax(me, fabs(got[i] - ref[i])); mr = fmax(mr, fabs(ref[i])); }
  printf("%-4s max_abs_err %.4e  max_ref %.4e  rel %.3e  %s\n", name, me, mr, me / mr, (me / mr < 2e-2) ? "OK" : "FAIL");
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
  __nv_bfloat16 *dq, *dk, *dv, *dho;
  float *oS, *oH, *oC, *oG, *oGt;
  CK(cudaMalloc(&dq, n * 2)); CK(cudaMalloc(&dk, n * 2)); CK(cudaMalloc(&dv, n * 2)); CK(cudaMalloc(&dho, n * 2));
  CK(cudaMemset(dho, 0, n * 2));
  CK(cudaMemcpy(dq, hq.data(), n * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dk, hk.data(), n * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dv, hv.data(), n * 2, cudaMemcpyHostToDevice));
  for (float** p : {&oS, &oH, &oC, &oG, &oGt}) CK(cudaMalloc(p, 128 * 128 * 4));
  Maps maps;
  // (B,S,NH,DH) storage viewed as (B,NH,S,DH): stride_b = S*NH*DH, stride_h = DH, stride_s = NH*DH
  int r = 0;
  r |= make_act_tmap(&maps.q, dq, 1, NH, S, DH, (int64_t)S * NH * DH, DH, NH * DH, 128);
  r |= make_act_tmap(&maps.k, dk, 1, NH, S, DH, (int64_t)S * NH * DH, DH, NH * DH, 128);
  r |= make_act_tmap(&maps.v, dv, 1, NH, S, DH, (int64_t)S * NH * DH, DH, NH * DH, 128);
  r |= make_act_tmap(&maps.hout, dho, 1, NH, S, DH, (int64_t)S * NH * DH, DH, NH * DH, 128);
  if (r) { printf("tensor map encode failed: %d\n", r); return 3; }
  const int smem = 10 * TILE + 1024;
  CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  probe_kernel<<<1, 128, smem>>>(maps, oS, oH, oC, oG, oGt);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  std::vector<float> gS(128 * 128), gH(128 * 128), gC(128 * 128), gG(128 * 128), gGt(128 * 128);
  CK(cudaMemcpy(gS.data(), oS, 128 * 128 * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gH.data(), oH, 128 * 128 * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gC.data(), oC, 128 * 128 * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gG.data(), oG, 128 * 128 * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(gGt.data(), oGt, 128 * 128 * 4, cudaMemcpyDeviceToHost));
  std::vector<__nv_bfloat16> hho(n);
  CK(cudaMemcpy(hho.data(), dho, n * 2, cudaMemcpyDeviceToHost));

  std::vector<double> rS(128 * 128), rH(128 * 128), rC(128 * 128), rG(128 * 128), rGt(128 * 128);
  std::vector<float> P(128 * 128), Cb(128 * 128);
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
  for (int dk_ = 0; dk_ < 128; ++dk_)
    for (int d = 0; d < 128; ++d) {
      double a = 0.5 * dk_ - 0.25 * d;
      for (int j = 0; j < 128; ++j) a += (double)K[j * DH + dk_] * V[j * DH + d];
      rC[dk_ * 128 + d] = a;
      Cb[dk_ * 128 + d] = bf((float)a / 64.f);
    }
  for (int t = 0; t < 128; ++t)
    for (int d = 0; d < 128; ++d) {
      double a = 0, b = 0;
      for (int x = 0; x < 128; ++x) {
        a += (double)Q[t * DH + x] * Cb[x * 128 + d];     // G[t][dv] = sum_dk Q[t][dk] Cb[dk][dv]
        b += (double)V[t * DH + x] * Cb[d * 128 + x];     // Gt[t][dk] = sum_dv V[t][dv] Cb[dk][dv]
      }
      rG[t * 128 + d] = a;
      rGt[t * 128 + d] = b;
    }
  report("T1", gS, rS);
  report("T2", gH, rH);
  report("T3", gC, rC);
  report("T4", gG, rG);
  report("T5", gGt, rGt);
  {
    double me = 0, other = 0;
    for (int s = 0; s < S; ++s)
      for (int d = 0; d < DH; ++d) {
        float got = __bfloat162float(hho[((size_t)s * NH + HEAD) * DH + d]);
        me = fmax(me, fabs(got - bf(gH[s * 128 + d])));
        other = fmax(other, fabs(__bfloat162float(hho[((size_t)s * NH + (1 - HEAD)) * DH + d])));
      }
    printf("T6   store max_abs_err %.4e  other-head max %.4e  %s\n", me, other, (me == 0 && other == 0) ? "OK" : "FAIL");
  }
  return 0;
}

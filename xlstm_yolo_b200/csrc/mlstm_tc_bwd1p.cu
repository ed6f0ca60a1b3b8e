// Single-pass tcgen05 / TMEM / TMA backward kernels of the mLSTM cell (bf16 I/O, DH in {64, 128}),
// used for SHORT sequences (few chunks per head, many heads): one CTA per (batch, head) keeps the
// state on chip, so no chunk state ever goes through HBM.  Long sequences / small batches take the
// chunk-parallel family in mlstm_tc_bwd.cu.
//
// The backward recomputes every gate/decay matrix per chunk from q,k,v,i,f and the per-row
// (n_t, m_t) the forward saved; nothing of size S x S or S x DH x DH is ever stored.  The
// adjoint (derived in tests/emu_kernel_dataflow.py, checked there against autograd through
// the oracle) splits into three chunk walks, each with one DH x DH state resident in TMEM:
//
//   kernel A  (scan order)      carries C (recomputed like the forward)        -> dq, dn_t, R_t
//   kernel B1 (reverse order)   carries dC                                     -> dv
//   kernel B2 (reverse order)   carries dC and dn                              -> dk, di, df
//
// with (s = qk scale, N_t = max(|n_t|, e^-m_t) + eps, D_tj = exp(u_j - M_t) causal):
//   dn_t = -[|n_t| >= e^-m_t] sign(n_t) (dh_t . h_t) / N_t
//   dS   = (dH V^T / N + dn) * D            dq = s [ dS K + w (dH C^T / N + dn n) ]
//   dv   = (s S^T D / N) dH + kw (K dC)     dk = s dS^T Q + kw (V dC^T + dn_state)
//   dC  <- decay dC + (w s / N  Q)^T dH     dn_state <- decay dn_state + Q^T (w s dn)
//   di_j = k_j . dk_j ;  df = sigmoid(-f) * suffix_sum(q . dq - k . dk)
// All contractions are tcgen05 MMAs (bf16 operands from 128B-swizzled shared memory tiles,
// fp32 accumulators in TMEM); gating / masks / normalisers are fused SIMT between them.
#include "mlstm_common.cuh"
#include "tc_ptx.cuh"
#include "tc_tmap.cuh"

namespace mlstm {
namespace {

using namespace ptx;

constexpr int L = 128;
constexpr int NT = 128;
constexpr int TILE = L * 128;
constexpr float LOG2E = 1.4426950408889634f;

struct BwdMaps { CUtensorMap q, k, v, dh, out0, out1; };  // out0/out1: dq | dv | dk

struct alignas(16) GateBufB {   // all indexed by tile row
  float u2[L];      // u * log2e
  float M2[L];      // M * log2e
  float w[L];       // exp(m_prev - M)
  float invN[L];    // 1 / N_t
  float dn[L];      // dn_t
  float kw[L];      // exp(u - M_L)
  float R[L];       // q . dq   (kernel B2)
  float sig[L];     // sigmoid(-f)
  float decay;
  float pad[3];
};

template <int DH, int NIN>
struct SmemB {
  static constexpr int KT = DH / 64;
  static constexpr int TILE_C = DH * 128;
  alignas(1024) uint8_t in[NIN][KT * TILE];      // TMA-loaded operand tiles
  alignas(1024) uint8_t x[2 * TILE];             // dS / E^T / dS^T tile (K-major), then staging
  alignas(1024) uint8_t cb[KT * TILE_C];         // bf16 state (C or dC), [dk][dv]
  alignas(1024) uint8_t vec[2 * 2048];           // K-major [16][128] bf16 vector tile (kw or w s dn)
  GateBufB g[2];
  float nvec[DH];                                // n_prev (A) / dn_state (B2), fp32
  float scan[8];
  uint64_t bar_in[NIN], bar_m1, bar_s, bar_m2;
  uint32_t tmem_base;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ uint64_t dK(uint32_t base, int ks, uint32_t atom_stride) {  // K-major operand, k-step ks
  return make_sdesc(base + (ks >> 2) * atom_stride + (ks & 3) * 32, 16, 1024);
}
__device__ __forceinline__ uint64_t dMN(uint32_t base, int ks, uint32_t lbo) {         // MN-major operand
  return make_sdesc(base + ks * 2048, lbo, 1024);
}

struct ChunkGeom { int mc, tok0, nvalid; };
__device__ __forceinline__ ChunkGeom geom(int mc, int S) {
  ChunkGeom g;
  g.mc = mc; g.tok0 = mc * L; g.nvalid = min(L, S - g.tok0);
  return g;
}

// Gate vectors of memory chunk `mc`, rebuilt from i, f and the saved rows (n_t, m_t).
// Thread t = scan-local index t.  `need_dn`: also load dn_t and R_t from the workspace.
template <class SM>
__device__ __forceinline__ void gates_from_rows(SM& sm, GateBufB& G, const mlstm_params& p, int b, int h, int bh, int mc,
                                                const float* ws_dn, const float* ws_R) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const ChunkGeom cg = geom(mc, p.S);
  const bool rev = p.reverse != 0;
  const bool valid = t < cg.nvalid;
  const int r = (rev && valid) ? (cg.nvalid - 1 - t) : t;
  float ii = -INFINITY, logf = 0.f, fi = 0.f, mrow = 0.f, nrow = 0.f, dn = 0.f, R = 0.f;
  if (valid) {
    const int tok = cg.tok0 + r;
    fi = p.f.ptr[(int64_t)b * p.f.stride_b + (int64_t)h * p.f.stride_h + (int64_t)tok * p.f.stride_s];
    ii = igate_log(p, p.i.ptr[(int64_t)b * p.i.stride_b + (int64_t)h * p.i.stride_h + (int64_t)tok * p.i.stride_s]);
    logf = log_sigmoid(fi);
    mrow = p.m_row[(int64_t)bh * p.S + tok];
    nrow = p.n_row[(int64_t)bh * p.S + tok];
    if (ws_dn) { dn = ws_dn[(int64_t)bh * p.S + tok]; R = ws_R[(int64_t)bh * p.S + tok]; }
  }
  float bs = warp_scan_add(logf, lane);
  if (lane == 31) sm.scan[warp] = bs;
  __syncthreads();
  float pre = 0.f;
#pragma unroll
  for (int w = 0; w < 4; ++w) pre += (w < warp) ? sm.scan[w] : 0.f;
  bs += pre;
  const float u = ii - bs;
  float M = mrow - bs;
  if (t == cg.nvalid - 1) sm.scan[4] = M;        // M_L: M at the last valid scan index
  // m of the row that precedes this chunk in scan order (or the initial m)
  float m_prev;
  {
    const int ptok = rev ? (cg.tok0 + cg.nvalid) : (cg.tok0 - 1);
    m_prev = p.gate_mode ? 0.f
             : ((ptok >= 0 && ptok < p.S) ? p.m_row[(int64_t)bh * p.S + ptok] : (p.m_initial ? p.m_initial[bh] : 0.f));
  }
  __syncthreads();
  const float ML = sm.scan[4];
  if (!valid) M = ML;
  G.u2[r] = u * LOG2E;
  G.M2[r] = M * LOG2E;
  G.w[r] = __expf(m_prev - M);
  G.invN[r] = valid ? 1.f / (fmaxf(fabsf(nrow), __expf(-mrow)) + p.eps) : 0.f;
  G.dn[r] = dn;
  G.kw[r] = __expf(u - ML);
  G.R[r] = R;
  G.sig[r] = 1.f / (1.f + __expf(fi));
  if (t == 0) G.decay = __expf(m_prev - ML);
  __syncthreads();
}

// rows of a [128][DH] swizzled bf16 tile set scaled in place by rowscale[row]
template <int DH>
__device__ __forceinline__ void scale_rows(uint8_t* tile, const float* rowscale) {
  constexpr int KT = DH / 64;
  for (int it = 0; it < KT * TILE / 16 / NT; ++it) {
    const uint32_t o = (uint32_t)(threadIdx.x + it * NT) * 16u;
    const float s = rowscale[(o >> 7) & (L - 1)];
    uint4 w = *reinterpret_cast<uint4*>(tile + o);
    __nv_bfloat162* kk = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 f2 = __bfloat1622float2(kk[e]);
      kk[e] = __floats2bfloat162_rn(f2.x * s, f2.y * s);
    }
    *reinterpret_cast<uint4*>(tile + o) = w;
  }
}

// K-major [16 rows][128] bf16 tile whose every row is the vector val[0..127] (thread j writes val_j)
__device__ __forceinline__ void write_vec_tile(uint8_t* vec, float val) {
  const int j = threadIdx.x;
  const __nv_bfloat16 bv = __float2bfloat16_rn(val);
#pragma unroll
  for (int row = 0; row < 16; ++row)
    *reinterpret_cast<__nv_bfloat16*>(vec + (j >> 6) * 2048 + swz128(row, j & 63)) = bv;
}

// dot of this thread's row of a swizzled [128][DH] bf16 tile with 32-float blocks held in regs is done inline.
template <int DH>
__device__ __forceinline__ void load_row_f32(const uint8_t* tile, int row, int cb32, float (&out)[32]) {
  // columns [cb32*32, cb32*32+32) of `row`
#pragma unroll
  for (int x = 0; x < 32; x += 8) {
    const int col = cb32 * 32 + x;
    const uint4 w = *reinterpret_cast<const uint4*>(tile + (col >> 6) * TILE + swz128(row, col & 63));
    const __nv_bfloat162* qq = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float2 f2 = __bfloat1622float2(qq[e]);
      out[x + 2 * e] = f2.x;
      out[x + 2 * e + 1] = f2.y;
    }
  }
}

// state pass: cb <- bf16(T), T <- dnext * T (skipped when last); nvec <- n column
template <int DH, class SM>
__device__ __forceinline__ void state_pass(SM& sm, uint32_t tC, uint32_t tN, bool has_n, float dnext, bool last,
                                           uint32_t lane_sel) {
  constexpr int TILE_C = DH * 128;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp < DH / 32) {
#pragma unroll 1
    for (int cbk = 0; cbk < DH / 32; ++cbk) {
      float r[32];
      tmem_ld32(tC + lane_sel + cbk * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int x = 0; x < 32; x += 8) {
        const int dv = cbk * 32 + x;
        *reinterpret_cast<uint4*>(sm.cb + (dv >> 6) * TILE_C + swz128(tid, dv & 63)) =
            make_uint4(pack_bf16x2(r[x], r[x + 1]), pack_bf16x2(r[x + 2], r[x + 3]), pack_bf16x2(r[x + 4], r[x + 5]),
                       pack_bf16x2(r[x + 6], r[x + 7]));
      }
      if (!last) {
#pragma unroll
        for (int x = 0; x < 32; ++x) r[x] *= dnext;
        tmem_st32(tC + lane_sel + cbk * 32, r);
      }
    }
    if (has_n) {
      float rn[16];
      tmem_ld16(tN + lane_sel, rn);
      tmem_ld_wait();
      sm.nvec[tid] = rn[0];
      if (!last) {
        float r32[32];
#pragma unroll
        for (int x = 0; x < 32; ++x) r32[x] = rn[0] * dnext;
        tmem_st32(tN + lane_sel, r32);
      }
    }
    if (!last) tmem_st_wait();
  }
}

template <class SM>
__device__ __forceinline__ void setup(SM& sm, int nbar_in, uint32_t tmem_cols = 512) {
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int i = 0; i < nbar_in; ++i) mbar_init(&sm.bar_in[i], 1);
    mbar_init(&sm.bar_m1, 1); mbar_init(&sm.bar_s, 1); mbar_init(&sm.bar_m2, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, tmem_cols);
}

// ======================================================================================
// Kernel A: scan-order walk, carries C and n.  Tiles: in[0]=q in[1]=k in[2]=v in[3]=dh.
// ======================================================================================
template <int DH>
__global__ void __launch_bounds__(NT, 1) tc_bwd_dq_kernel(const __grid_constant__ BwdMaps maps, const mlstm_params p,
                                                          const float scale, float* __restrict__ ws_dn,
                                                          float* __restrict__ ws_R) {
  constexpr int KT = DH / 64;
  constexpr int TILE_C = DH * 128;
  constexpr uint32_t A_LBO_STATE = (DH == 128) ? TILE : 0;
  using SM = SmemB<DH, 4>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  uint8_t *sq = sm.in[0], *sk = sm.in[1], *sv = sm.in[2], *sdh = sm.in[3];

  const int tid = threadIdx.x, warp = tid >> 5;
  const int bh = blockIdx.x, b = bh / p.NH, h = bh % p.NH;
  const int S = p.S, NC = (S + L - 1) / L;
  const bool has_init = p.c_initial != nullptr;
  const bool rev = p.reverse != 0;

  setup(sm, 4);
  if (tid == 0) { tma_prefetch_desc(&maps.q); tma_prefetch_desc(&maps.k); tma_prefetch_desc(&maps.v); tma_prefetch_desc(&maps.dh); tma_prefetch_desc(&maps.out0); }
  if (!has_init) {
    for (int e = tid; e < KT * TILE_C / 16; e += NT) reinterpret_cast<uint4*>(sm.cb)[e] = make_uint4(0, 0, 0, 0);
    for (int e = tid; e < DH; e += NT) sm.nvec[e] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = sm.tmem_base;
  const uint32_t tZ = tm, tG = tm + 128, tC = tm + 128 + DH, tN = tm + 128 + 2 * DH;
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;

  auto mc_of = [&](int c) { return rev ? (NC - 1 - c) : c; };
  auto load = [&](int slot, const CUtensorMap* map, int c) {
    mbar_arrive_expect_tx(&sm.bar_in[slot], KT * TILE);
    for (int kt = 0; kt < KT; ++kt) tma_load_4d(sm.in[slot] + kt * TILE, map, &sm.bar_in[slot], kt * 64, mc_of(c) * L, h, b);
  };
  if (tid == 0) { load(3, &maps.dh, 0); load(2, &maps.v, 0); load(1, &maps.k, 0); load(0, &maps.q, 0); }
  gates_from_rows(sm, sm.g[0], p, b, h, bh, mc_of(0), nullptr, nullptr);

  if (has_init) {
    const float d0 = sm.g[0].decay;
    if (tid < DH) {
      const float* crow = p.c_initial + ((int64_t)bh * DH + tid) * DH;
      for (int cbk = 0; cbk < DH / 32; ++cbk) {
        float r[32];
#pragma unroll
        for (int x = 0; x < 32; ++x) r[x] = crow[cbk * 32 + x];
#pragma unroll
        for (int x = 0; x < 32; x += 8) {
          const int dv = cbk * 32 + x;
          *reinterpret_cast<uint4*>(sm.cb + (dv >> 6) * TILE_C + swz128(tid, dv & 63)) =
              make_uint4(pack_bf16x2(r[x], r[x + 1]), pack_bf16x2(r[x + 2], r[x + 3]), pack_bf16x2(r[x + 4], r[x + 5]),
                         pack_bf16x2(r[x + 6], r[x + 7]));
        }
#pragma unroll
        for (int x = 0; x < 32; ++x) r[x] *= d0;
        tmem_st32(tC + lane_sel + cbk * 32, r);
      }
      const float n0 = p.n_initial[(int64_t)bh * DH + tid];
      sm.nvec[tid] = n0;
      float r32[32];
#pragma unroll
      for (int x = 0; x < 32; ++x) r32[x] = n0 * d0;
      tmem_st32(tN + lane_sel, r32);
      tmem_st_wait();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  for (int c = 0; c < NC; ++c) {
    GateBufB& G = sm.g[c & 1];
    GateBufB& Gn = sm.g[(c + 1) & 1];
    const uint32_t ph = c & 1;
    const ChunkGeom cg = geom(mc_of(c), S);
    const int tok = cg.tok0 + tid;
    const bool row_ok = tid < cg.nvalid;

    // ---- MMA1: Z = dH V^T, G = dH Cb^T ---------------------------------------------------
    mbar_wait(&sm.bar_in[3], ph);
    mbar_wait(&sm.bar_in[2], ph);
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idZ = make_idesc_bf16(128, 128, 0, 0);
      for (int ks = 0; ks < DH / 16; ++ks)
        umma_bf16_ss(tZ, dK(smem_u32(sdh), ks, TILE), dK(smem_u32(sv), ks, TILE), idZ, ks > 0);
      constexpr uint32_t idG = make_idesc_bf16(128, DH, 0, 0);
      for (int ks = 0; ks < DH / 16; ++ks)
        umma_bf16_ss(tG, dK(smem_u32(sdh), ks, TILE), dK(smem_u32(sm.cb), ks, TILE_C), idG, ks > 0);
      umma_commit(&sm.bar_m1);
    }
    // ---- shadow: next gates; dn_t = -[active] sign(n) (dh . h) / N ------------------------
    if (c + 1 < NC) gates_from_rows(sm, Gn, p, b, h, bh, mc_of(c + 1), nullptr, nullptr);
    float dn = 0.f;
    const float invN = G.invN[tid];
    if (row_ok) {
      const __nv_bfloat16* hrow = reinterpret_cast<const __nv_bfloat16*>(p.h.ptr) + (int64_t)b * p.h.stride_b +
                                  (int64_t)h * p.h.stride_h + (int64_t)tok * p.h.stride_s;
      float hd = 0.f;
#pragma unroll 1
      for (int cbk = 0; cbk < DH / 32; ++cbk) {
        float dhr[32];
        load_row_f32<DH>(sdh, tid, cbk, dhr);
#pragma unroll
        for (int x = 0; x < 32; x += 8) {
          const uint4 w = *reinterpret_cast<const uint4*>(hrow + cbk * 32 + x);
          const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 f2 = __bfloat1622float2(hh[e]);
            hd += dhr[x + 2 * e] * f2.x + dhr[x + 2 * e + 1] * f2.y;
          }
        }
      }
      const float nr = p.n_row[(int64_t)bh * S + tok];
      const float mr = p.m_row[(int64_t)bh * S + tok];
      dn = (fabsf(nr) >= __expf(-mr)) ? -copysignf(1.f, nr) * hd * invN : 0.f;
      ws_dn[(int64_t)bh * S + tok] = dn;
    }
    mbar_wait(&sm.bar_m1, ph);
    tc_fence_after();

    // ---- Vbar = kw * V in place; kw vector tile -------------------------------------------
    scale_rows<DH>(sv, G.kw);
    write_vec_tile(sm.vec, G.kw[tid]);
    fence_proxy_async_smem();
    if (tid == 0) tma_store_wait_read<0>();
    tc_fence_before();
    __syncthreads();

    // ---- state MMAs: C += K^T Vbar, n += K^T kw ; prefetch dH(c+1) -------------------------
    if (tid == 0) {
      if (c + 1 < NC) load(3, &maps.dh, c + 1);
      mbar_wait(&sm.bar_in[1], ph);
      tc_fence_after();
      constexpr uint32_t idC = make_idesc_bf16(128, DH, 1, 1);
      constexpr uint32_t idN = make_idesc_bf16(128, 16, 1, 0);
      const uint32_t acc0 = (c > 0 || has_init) ? 1u : 0u;
      for (int ks = 0; ks < L / 16; ++ks) {
        const uint64_t a = dMN(smem_u32(sk), ks, A_LBO_STATE);
        umma_bf16_ss(tC, a, dMN(smem_u32(sv), ks, TILE), idC, (ks > 0) ? 1u : acc0);
        umma_bf16_ss(tN, a, dK(smem_u32(sm.vec), ks, 2048), idN, (ks > 0) ? 1u : acc0);
      }
      umma_commit(&sm.bar_s);
    }

    // ---- dS = (Z / N + dn) * exp2(u2_j - M2_t), causal -> bf16 tile -----------------------
    const float M2t = G.M2[tid];
#pragma unroll 1
    for (int cbk = 0; cbk < 4; ++cbk) {
      const bool full = rev ? (cbk > warp) : (cbk < warp);
      const bool diag = (cbk == warp);
      uint32_t packed[16];
      if (full || diag) {
        float z[32];
        tmem_ld32(tZ + lane_sel + cbk * 32, z);
        const uint32_t cbits = causal_bits(full, !rev, tid & 31);
        tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; x += 4) {
          const float4 u4 = *reinterpret_cast<const float4*>(&G.u2[cbk * 32 + x]);
          const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
          float pv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool keep = (cbits >> (x + e)) & 1u;
            pv[e] = keep ? fmaf(z[x + e], invN, dn) * ex2(uu[e] - M2t) : 0.f;
          }
          packed[x / 2] = pack_bf16x2(pv[0], pv[1]);
          packed[x / 2 + 1] = pack_bf16x2(pv[2], pv[3]);
        }
      } else {
#pragma unroll
        for (int x = 0; x < 16; ++x) packed[x] = 0u;
      }
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int j = cbk * 32 + x * 8;
        *reinterpret_cast<uint4*>(sm.x + (j >> 6) * TILE + swz128(tid, j & 63)) =
            make_uint4(packed[4 * x], packed[4 * x + 1], packed[4 * x + 2], packed[4 * x + 3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    // ---- MMA2: dQ = dS K (into the Z columns) ; prefetch V(c+1) ---------------------------
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idQ = make_idesc_bf16(128, DH, 0, 1);
      for (int ks = 0; ks < L / 16; ++ks)
        umma_bf16_ss(tZ, dK(smem_u32(sm.x), ks, TILE), dMN(smem_u32(sk), ks, TILE), idQ, ks > 0);
      umma_commit(&sm.bar_m2);
      mbar_wait(&sm.bar_s, ph);
      if (c + 1 < NC) load(2, &maps.v, c + 1);
    }
    mbar_wait(&sm.bar_in[0], ph);
    mbar_wait(&sm.bar_m2, ph);
    tc_fence_after();
    if (tid == 0 && c + 1 < NC) load(1, &maps.k, c + 1);

    // ---- epilogue: dq = s [ dQ + w (G / N + dn n_prev) ]; R = q . dq ----------------------
    const float wt = G.w[tid];
    float Rt = 0.f;
#pragma unroll 1
    for (int cbk = 0; cbk < DH / 32; ++cbk) {
      float dq_[32], gg[32], qr[32];
      tmem_ld32(tZ + lane_sel + cbk * 32, dq_);
      tmem_ld32(tG + lane_sel + cbk * 32, gg);
      tmem_ld_wait();
      load_row_f32<DH>(sq, tid, cbk, qr);
#pragma unroll
      for (int x = 0; x < 32; x += 8) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          o[e] = scale * (dq_[x + e] + wt * fmaf(gg[x + e], invN, dn * sm.nvec[cbk * 32 + x + e]));
          Rt = fmaf(qr[x + e], o[e], Rt);
        }
        const int dk_ = cbk * 32 + x;
        *reinterpret_cast<uint4*>(sm.x + (dk_ >> 6) * TILE + swz128(tid, dk_ & 63)) =
            make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
      }
    }
    if (row_ok) ws_R[(int64_t)bh * S + tok] = Rt;
    __syncthreads();   // every thread is done with nvec (n_prev) before the state pass rewrites it
    const bool last = (c + 1 == NC);
    state_pass<DH>(sm, tC, tN, true, last ? 1.f : Gn.decay, last, lane_sel);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      for (int kt = 0; kt < KT; ++kt) tma_store_4d(&maps.out0, sm.x + kt * TILE, kt * 64, cg.tok0, h, b);
      tma_store_commit();
      if (c + 1 < NC) load(0, &maps.q, c + 1);
    }
  }
  if (tid == 0) tma_store_wait_all<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ======================================================================================
// Kernels B1 / B2: reverse-order walk carrying dC.  Thread j <-> key row j.
//   MODE 1 (dv): tiles in[0]=q in[1]=k  in[2]=dh ; out0 = dv
//   MODE 2 (dk): tiles in[0]=q in[1]=v  in[2]=dh ; out0 = dk ; also di, df, carries dn_state
// ======================================================================================
// TMEM: S/Z 128 + O DH + dC DH columns.  At DH = 64 that is 256, so a dv CTA and a dk CTA share an SM
// (two 96 KB CTAs) and the two reverse walks run side by side in one launch (tc_bwd_dkv12_kernel).
// The dn_state recurrence of the dk walk is a [DH] vector: it lives in registers (SIMT column sums of
// the Q tile) rather than in a fourth TMEM accumulator.
template <int DH, int MODE>
__device__ __forceinline__ void dkv_body(const BwdMaps& maps, const mlstm_params& p, const float scale,
                                         const float* __restrict__ ws_dn, const float* __restrict__ ws_R, const int bh) {
  constexpr int KT = DH / 64;
  constexpr int TILE_C = DH * 128;
  constexpr uint32_t TCOLS = (DH == 64) ? 256 : 512;
  constexpr uint32_t A_LBO_STATE = (DH == 128) ? TILE : 0;
  using SM = SmemB<DH, 3>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SM& sm = *reinterpret_cast<SM*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
  uint8_t *sq = sm.in[0], *skv = sm.in[1], *sdh = sm.in[2];
  __shared__ __align__(16) float colv[3][L];     // per-query-row vectors: [0] exponent offset, [1] 1/N, [2] dn
  __shared__ __align__(16) float rowscale[L];    // (w s / N)_t for the dC update operand
  __shared__ __align__(16) float kwos[L];        // kw_j / s
  __shared__ __align__(16) float ncoef[L];       // (w s dn)_t : weights of the dn_state update (dk walk)
  __shared__ float npart[2][DH];
  __shared__ float df_carry;

  const int tid = threadIdx.x, warp = tid >> 5;
  const int b = bh / p.NH, h = bh % p.NH;
  float nstate = 0.f;                            // thread dk < DH: decayed dn_state entering this step
  const int S = p.S, NC = (S + L - 1) / L;
  const bool rev = p.reverse != 0;
  const CUtensorMap* map_kv = (MODE == 1) ? &maps.k : &maps.v;

  setup(sm, 3, TCOLS);
  if (tid == 0) { tma_prefetch_desc(&maps.q); tma_prefetch_desc(map_kv); tma_prefetch_desc(&maps.dh); tma_prefetch_desc(&maps.out0); df_carry = 0.f; }
  for (int e = tid; e < KT * TILE_C / 16; e += NT) reinterpret_cast<uint4*>(sm.cb)[e] = make_uint4(0, 0, 0, 0);
  for (int e = tid; e < DH; e += NT) sm.nvec[e] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = sm.tmem_base;
  const uint32_t tS = tm, tO = tm + 128, tC = tm + 128 + DH;
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;

  // processing step c (0 = scan-last chunk) -> memory chunk
  auto mc_of = [&](int c) { return rev ? c : (NC - 1 - c); };
  auto load = [&](int slot, const CUtensorMap* map, int c) {
    mbar_arrive_expect_tx(&sm.bar_in[slot], KT * TILE);
    for (int kt = 0; kt < KT; ++kt) tma_load_4d(sm.in[slot] + kt * TILE, map, &sm.bar_in[slot], kt * 64, mc_of(c) * L, h, b);
  };
  if (tid == 0) { load(1, map_kv, 0); load(2, &maps.dh, 0); load(0, &maps.q, 0); }
  gates_from_rows(sm, sm.g[0], p, b, h, bh, mc_of(0), MODE == 2 ? ws_dn : nullptr, ws_R);

  for (int c = 0; c < NC; ++c) {
    GateBufB& G = sm.g[c & 1];
    GateBufB& Gn = sm.g[(c + 1) & 1];
    const uint32_t ph = c & 1;
    const ChunkGeom cg = geom(mc_of(c), S);
    const int tok = cg.tok0 + tid;
    const bool row_ok = tid < cg.nvalid;

    // ---- MMA1: S^T = K Q^T (dv)  |  Z^T = V dH^T (dk) -------------------------------------
    mbar_wait(&sm.bar_in[1], ph);
    mbar_wait(&sm.bar_in[MODE == 1 ? 0 : 2], ph);
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t id1 = make_idesc_bf16(128, 128, 0, 0);
      const uint32_t bb = smem_u32(MODE == 1 ? sq : sdh);
      for (int ks = 0; ks < DH / 16; ++ks)
        umma_bf16_ss(tS, dK(smem_u32(skv), ks, TILE), dK(bb, ks, TILE), id1, ks > 0);
      umma_commit(&sm.bar_m1);
    }
    // ---- shadow: next gates; per-query-row vectors ---------------------------------------
    if (c + 1 < NC) gates_from_rows(sm, Gn, p, b, h, bh, mc_of(c + 1), MODE == 2 ? ws_dn : nullptr, ws_R);
    {
      const float inv = G.invN[tid];
      // dv: E^T = S^T * exp2(u2_j + log2 s - (M2_t - log2(invN_t)))   (1/N folded into the exponent)
      // dk: dS^T = (Z^T invN_t + dn_t) * exp2(u2_j + log2 s - M2_t)
      colv[0][tid] = (MODE == 1) ? (G.M2[tid] - ((inv > 0.f) ? log2f(inv) : -INFINITY)) : G.M2[tid];
      colv[1][tid] = inv;
      colv[2][tid] = G.dn[tid];
      rowscale[tid] = G.w[tid] * scale * inv;
      kwos[tid] = G.kw[tid] / scale;
      ncoef[tid] = G.w[tid] * scale * G.dn[tid];
    }
    mbar_wait(&sm.bar_in[MODE == 1 ? 2 : 0], ph);   // third tile (needed by the state MMAs)
    mbar_wait(&sm.bar_m1, ph);
    tc_fence_after();
    __syncthreads();   // colv / rowscale visible
    if (MODE == 2) {   // dn_state += Q^T (w s dn): column dk of the (un-scaled) Q tile, half of the rows per thread
      const int dk = tid % DH, part = tid / DH;
      constexpr int PARTS = NT / DH, ROWS = L / PARTS;
      float acc = 0.f;
#pragma unroll 8
      for (int t = part * ROWS; t < (part + 1) * ROWS; ++t)
        acc = fmaf(ncoef[t], __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(sq + (dk >> 6) * TILE + swz128(t, dk & 63))), acc);
      if (PARTS == 2) npart[part][dk] = acc;
      else npart[0][dk] = acc;
    }

    // ---- in-place operand scaling ---------------------------------------------------------
    if (MODE == 1) scale_rows<DH>(skv, G.kw);                   // Kbar = kw_j k_j
    else scale_rows<DH>(skv, kwos);                             // Vbar = (kw_j / s) v_j  (s applied in the epilogue)
    if (MODE == 1) scale_rows<DH>(sq, rowscale);                // Qtilde = (w s / N)_t q_t
    else {
      scale_rows<DH>(sdh, rowscale);                            // dHtilde = (w s / N)_t dh_t
    }
    fence_proxy_async_smem();
    if (tid == 0) tma_store_wait_read<0>();
    tc_fence_before();
    __syncthreads();

    // ---- inter-chunk MMA (overwrites tO) and the dC / dn_state updates ---------------------
    if (tid == 0) {
      tc_fence_after();
      if (MODE == 1) {   // dV = Kbar dCb : A K-major [j][dk], B MN-major [dk][dv]
        constexpr uint32_t idI = make_idesc_bf16(128, DH, 0, 1);
        for (int ks = 0; ks < DH / 16; ++ks)
          umma_bf16_ss(tO, dK(smem_u32(skv), ks, TILE), dMN(smem_u32(sm.cb), ks, TILE_C), idI, ks > 0);
      } else {           // dK = Vbar dCb^T : A K-major [j][dv], B K-major view of dCb (rows dk)
        constexpr uint32_t idI = make_idesc_bf16(128, DH, 0, 0);
        for (int ks = 0; ks < DH / 16; ++ks)
          umma_bf16_ss(tO, dK(smem_u32(skv), ks, TILE), dK(smem_u32(sm.cb), ks, TILE_C), idI, ks > 0);
      }
      constexpr uint32_t idC = make_idesc_bf16(128, DH, 1, 1);
      const uint32_t acc0 = (c > 0) ? 1u : 0u;
      for (int ks = 0; ks < L / 16; ++ks) {
        const uint64_t a = dMN(smem_u32(sq), ks, A_LBO_STATE);
        umma_bf16_ss(tC, a, dMN(smem_u32(sdh), ks, TILE), idC, (ks > 0) ? 1u : acc0);
      }
      umma_commit(&sm.bar_s);
    }

    // ---- E^T / dS^T tile: thread j owns key row j, columns t -------------------------------
    // dv: s folded into the exponent.  dk: NOT folded — kernel A rounds the same un-scaled dS to
    // bf16, so the rounding noise of R = q.dq and K = k.dk stays correlated and cancels in df.
    const float u2j = G.u2[tid] + ((MODE == 1) ? log2f(scale) : 0.f);
#pragma unroll 1
    for (int cbk = 0; cbk < 4; ++cbk) {
      // key j contributes to query t iff t is at/after j in scan order: forward t >= j, reverse t <= j
      const bool full = rev ? (cbk < warp) : (cbk > warp);
      const bool diag = (cbk == warp);
      uint32_t packed[16];
      if (full || diag) {
        float s_[32];
        tmem_ld32(tS + lane_sel + cbk * 32, s_);
        const uint32_t cbits = causal_bits(full, rev, tid & 31);
        tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; x += 4) {
          const float4 c4 = *reinterpret_cast<const float4*>(&colv[0][cbk * 32 + x]);
          const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
          float in4[4] = {0.f, 0.f, 0.f, 0.f}, dn4[4] = {0.f, 0.f, 0.f, 0.f};
          if (MODE == 2) {
            const float4 i4 = *reinterpret_cast<const float4*>(&colv[1][cbk * 32 + x]);
            const float4 d4 = *reinterpret_cast<const float4*>(&colv[2][cbk * 32 + x]);
            in4[0] = i4.x; in4[1] = i4.y; in4[2] = i4.z; in4[3] = i4.w;
            dn4[0] = d4.x; dn4[1] = d4.y; dn4[2] = d4.z; dn4[3] = d4.w;
          }
          float pv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool keep = (cbits >> (x + e)) & 1u;
            const float dd = ex2(u2j - cc[e]);
            const float val = (MODE == 1) ? s_[x + e] * dd : fmaf(s_[x + e], in4[e], dn4[e]) * dd;
            pv[e] = keep ? val : 0.f;
          }
          packed[x / 2] = pack_bf16x2(pv[0], pv[1]);
          packed[x / 2 + 1] = pack_bf16x2(pv[2], pv[3]);
        }
      } else {
#pragma unroll
        for (int x = 0; x < 16; ++x) packed[x] = 0u;
      }
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int t = cbk * 32 + x * 8;
        *reinterpret_cast<uint4*>(sm.x + (t >> 6) * TILE + swz128(tid, t & 63)) =
            make_uint4(packed[4 * x], packed[4 * x + 1], packed[4 * x + 2], packed[4 * x + 3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    // ---- MMA2: dV += E^T dH  |  dK += dS^T Q ----------------------------------------------
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(&sm.bar_s, ph);   // dv: dH must not be... (dH unscaled is read here; dk: Q unscaled) and tO complete
      constexpr uint32_t id2 = make_idesc_bf16(128, DH, 0, 1);
      const uint32_t bb = smem_u32(MODE == 1 ? sdh : sq);
      for (int ks = 0; ks < L / 16; ++ks)
        umma_bf16_ss(tO, dK(smem_u32(sm.x), ks, TILE), dMN(bb, ks, TILE), id2, 1u);
      umma_commit(&sm.bar_m2);
      if (c + 1 < NC) load(1, map_kv, c + 1);   // K/V tile free (inter MMA done)
      if (c + 1 < NC) load(MODE == 1 ? 0 : 2, MODE == 1 ? &maps.q : &maps.dh, c + 1);  // scaled tile free (dC MMA done)
    }
    mbar_wait(&sm.bar_m2, ph);
    tc_fence_after();
    if (tid == 0 && c + 1 < NC) load(MODE == 1 ? 2 : 0, MODE == 1 ? &maps.dh : &maps.q, c + 1);

    // ---- epilogue --------------------------------------------------------------------------
    float Kj = 0.f;
    const float kwj = G.kw[tid];
    const __nv_bfloat16* krow = nullptr;
    if (MODE == 2 && row_ok)
      krow = reinterpret_cast<const __nv_bfloat16*>(p.k.ptr) + (int64_t)b * p.k.stride_b + (int64_t)h * p.k.stride_h +
             (int64_t)tok * p.k.stride_s;
#pragma unroll 1
    for (int cbk = 0; cbk < DH / 32; ++cbk) {
      float o_[32];
      tmem_ld32(tO + lane_sel + cbk * 32, o_);
      tmem_ld_wait();
#pragma unroll
      for (int x = 0; x < 32; x += 8) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (MODE == 2) ? fmaf(kwj, sm.nvec[cbk * 32 + x + e], scale * o_[x + e]) : o_[x + e];
        if (MODE == 2 && row_ok) {
          const uint4 w = *reinterpret_cast<const uint4*>(krow + cbk * 32 + x);
          const __nv_bfloat162* kk = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            float2 f2 = __bfloat1622float2(kk[e]);
            Kj = fmaf(f2.x, o[2 * e], Kj);
            Kj = fmaf(f2.y, o[2 * e + 1], Kj);
          }
        }
        const int d_ = cbk * 32 + x;
        *reinterpret_cast<uint4*>(sm.x + (d_ >> 6) * TILE + swz128(tid, d_ & 63)) =
            make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
      }
    }
    if (MODE == 2) {
      // di_j = K_j ; df_j = sigmoid(-f_j) * (sum over scan positions >= j of (R - K) + carry)
      const float dB = row_ok ? (G.R[tid] - Kj) : 0.f;
      const int lane = tid & 31;
      float pre = warp_scan_add(dB, lane);
      if (lane == 31) sm.scan[warp] = pre;
      __syncthreads();
      float off = 0.f, tot = 0.f;
#pragma unroll
      for (int w = 0; w < 4; ++w) { off += (w < warp) ? sm.scan[w] : 0.f; tot += sm.scan[w]; }
      pre += off;
      const float carry = df_carry;
      const float suf = rev ? pre : (tot - pre + dB);
      if (row_ok) {
        const float i_raw = p.i.ptr[(int64_t)b * p.i.stride_b + (int64_t)h * p.i.stride_h + (int64_t)tok * p.i.stride_s];
        p.di.ptr[(int64_t)b * p.di.stride_b + (int64_t)h * p.di.stride_h + (int64_t)tok * p.di.stride_s] = Kj * igate_dlog(p, i_raw);
        p.df.ptr[(int64_t)b * p.df.stride_b + (int64_t)h * p.df.stride_h + (int64_t)tok * p.df.stride_s] =
            (suf + carry) * G.sig[tid];
      }
      __syncthreads();
      if (tid == 0) df_carry = carry + tot;
    } else {
      __syncthreads();   // nvec readers / uniform barrier count
    }
    const bool last = (c + 1 == NC);
    state_pass<DH>(sm, tC, 0u, false, last ? 1.f : Gn.decay, last, lane_sel);
    if (MODE == 2 && tid < DH) {   // the epilogue above read the previous dn_state: publish this step's
      const float nv = nstate + npart[0][tid] + ((NT / DH == 2) ? npart[1][tid] : 0.f);
      sm.nvec[tid] = nv;
      nstate = last ? 0.f : nv * Gn.decay;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      for (int kt = 0; kt < KT; ++kt) tma_store_4d(&maps.out0, sm.x + kt * TILE, kt * 64, cg.tok0, h, b);
      tma_store_commit();
    }
  }
  if (tid == 0) tma_store_wait_all<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, TCOLS);
}

template <int DH, int MODE>
__global__ void __launch_bounds__(NT, 2) tc_bwd_dkv_kernel(const __grid_constant__ BwdMaps maps, const mlstm_params p,
                                                           const float scale, const float* __restrict__ ws_dn,
                                                           const float* __restrict__ ws_R) {
  dkv_body<DH, MODE>(maps, p, scale, ws_dn, ws_R, blockIdx.x);
}

// dv and dk walks of the same (batch, head) as neighbouring CTAs of one launch (DH = 64: both fit one SM)
template <int DH>
__global__ void __launch_bounds__(NT, 2) tc_bwd_dkv12_kernel(const __grid_constant__ BwdMaps maps_v, const __grid_constant__ BwdMaps maps_k,
                                                             const mlstm_params p, const float scale,
                                                             const float* __restrict__ ws_dn, const float* __restrict__ ws_R) {
  if (blockIdx.x & 1) dkv_body<DH, 2>(maps_k, p, scale, ws_dn, ws_R, blockIdx.x >> 1);
  else dkv_body<DH, 1>(maps_v, p, scale, ws_dn, ws_R, blockIdx.x >> 1);
}

template <class K>
int prep_kernel(K kernel, size_t smem, const char* name) {
  cudaError_t e = set_max_smem_once(reinterpret_cast<const void*>(kernel), smem);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%s, %zu B): %s", name, smem, cudaGetErrorString(e));
    return MLSTM_ERR_CUDA;
  }
  return MLSTM_OK;
}
int launched(const char* name) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s launch failed: %s", name, cudaGetErrorString(e));
    return MLSTM_ERR_CUDA;
  }
  return MLSTM_OK;
}

template <int DH>
int launch_bwd(const mlstm_params& p, cudaStream_t st, int part) {
  BwdMaps m;
  int r = 0;
  r |= make_act_tmap(&m.q, p.q.ptr, p.B, p.NH, p.S, DH, p.q.stride_b, p.q.stride_h, p.q.stride_s, L);
  r |= make_act_tmap(&m.k, p.k.ptr, p.B, p.NH, p.S, DH, p.k.stride_b, p.k.stride_h, p.k.stride_s, L);
  r |= make_act_tmap(&m.v, p.v.ptr, p.B, p.NH, p.S, DH, p.v.stride_b, p.v.stride_h, p.v.stride_s, L);
  r |= make_act_tmap(&m.dh, p.dh.ptr, p.B, p.NH, p.S, DH, p.dh.stride_b, p.dh.stride_h, p.dh.stride_s, L);
  m.out1 = m.q;
  const size_t rows = (size_t)p.B * p.NH * p.S;
  float* ws_dn = reinterpret_cast<float*>(p.workspace);
  float* ws_R = ws_dn + rows;
  const float scale = resolve_scale(p);
  dim3 grid(p.B * p.NH), block(NT);
  int rc;
  if (part != 1) {
    r |= make_act_tmap(&m.out0, p.dq.ptr, p.B, p.NH, p.S, DH, p.dq.stride_b, p.dq.stride_h, p.dq.stride_s, L);
    if (r) { set_error("cuTensorMapEncodeTiled failed (%d)", r); return r == -1 ? MLSTM_ERR_NO_DEVICE : MLSTM_ERR_INVALID_ARG; }
    const size_t smA = sizeof(SmemB<DH, 4>) + 1024;
    if ((rc = prep_kernel(tc_bwd_dq_kernel<DH>, smA, "tc_bwd_dq"))) return rc;
    tc_bwd_dq_kernel<DH><<<grid, block, smA, st>>>(m, p, scale, ws_dn, ws_R);
    if ((rc = launched("tc_bwd_dq"))) return rc;
  }
  if (part != 0) {
    const size_t smB = sizeof(SmemB<DH, 3>) + 1024;
    BwdMaps mv = m, mk = m;
    r |= make_act_tmap(&mv.out0, p.dv.ptr, p.B, p.NH, p.S, DH, p.dv.stride_b, p.dv.stride_h, p.dv.stride_s, L);
    r |= make_act_tmap(&mk.out0, p.dk.ptr, p.B, p.NH, p.S, DH, p.dk.stride_b, p.dk.stride_h, p.dk.stride_s, L);
    if (r) { set_error("cuTensorMapEncodeTiled failed (%d)", r); return r == -1 ? MLSTM_ERR_NO_DEVICE : MLSTM_ERR_INVALID_ARG; }
    if (DH == 64) {   // the two reverse walks side by side: 2 CTAs (96 KB, 256 TMEM columns each) per SM
      if ((rc = prep_kernel(tc_bwd_dkv12_kernel<DH>, smB, "tc_bwd_dkdv"))) return rc;
      tc_bwd_dkv12_kernel<DH><<<dim3(2 * p.B * p.NH), block, smB, st>>>(mv, mk, p, scale, ws_dn, ws_R);
      if ((rc = launched("tc_bwd_dkdv"))) return rc;
    } else {
      if ((rc = prep_kernel(tc_bwd_dkv_kernel<DH, 1>, smB, "tc_bwd_dv"))) return rc;
      if ((rc = prep_kernel(tc_bwd_dkv_kernel<DH, 2>, smB, "tc_bwd_dk"))) return rc;
      tc_bwd_dkv_kernel<DH, 1><<<grid, block, smB, st>>>(mv, p, scale, ws_dn, ws_R);
      if ((rc = launched("tc_bwd_dv"))) return rc;
      tc_bwd_dkv_kernel<DH, 2><<<grid, block, smB, st>>>(mk, p, scale, ws_dn, ws_R);
      if ((rc = launched("tc_bwd_dk"))) return rc;
    }
  }
  return MLSTM_OK;
}

}  // namespace

size_t tc_bwd1p_workspace(const mlstm_params& p) { return sizeof(float) * 2 * (size_t)p.B * p.NH * p.S; }

int tc_bwd1p(const mlstm_params& p, cudaStream_t st, int part) {
  if (p.DHQK == 64) return launch_bwd<64>(p, st, part);
  return launch_bwd<128>(p, st, part);
}

}  // namespace mlstm

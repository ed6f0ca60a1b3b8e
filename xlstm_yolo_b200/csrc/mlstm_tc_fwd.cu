// tcgen05 / TMEM / TMA forward kernel of the mLSTM cell for bf16 I/O, DH in {64, 128}.
//
// One CTA (128 threads, thread t <-> tile row t <-> TMEM lane t) per (batch, head) walks the
// sequence in chunks of L = 128 tokens with the (C, n, m) state resident on chip:
//   C  : fp32 accumulator in TMEM, updated by an accumulating MMA  C += Kbar^T V
//   Cb : bf16 copy of C in shared memory (MN-major B operand of  G = Q Cb)
//   n  : fp32 in TMEM (MMA against a ones tile) + an fp32 copy in shared memory
//   m  : scalar carried in shared memory
// Per chunk (reference: backends.py:149-263; variables as in DESIGN.md / oracle):
//   MMA1  S = Q K^T,  G = Q Cb                          (TMA-loaded swizzled tiles)
//   SIMT  gates of the NEXT chunk (log-sigmoid cumsum, running max) ; q.n_prev
//   SIMT  Kbar = kw * K in place;  MMA  C += Kbar^T V,  n += Kbar^T 1
//   SIMT  P = S * exp2(u - M) (causal) -> bf16 swizzled tile ; row sums
//   MMA2  H = P V   (into the TMEM columns S occupied)
//   SIMT  h = (H + w G) / (max(|n|, e^-m) + eps) -> bf16 staging tile -> TMA store
//   SIMT  Cb <- bf16(C),  C <- decay_next * C   (one TMEM pass per chunk)
// No intermediate (D matrix, gates, P) ever reaches HBM.  `reverse` walks the tokens from
// the end: tiles stay in memory order, the causal mask and the gate scans flip.
#include "mlstm_common.cuh"
#include "tc_ptx.cuh"
#include "tc_tmap.cuh"

namespace mlstm {
namespace {

using namespace ptx;

constexpr int L = 128;                 // chunk rows
constexpr int NT = 128;                // threads per CTA
constexpr int TILE = L * 128;          // bytes of one [128 rows][64 bf16] swizzled tile
constexpr float LOG2E = 1.4426950408889634f;

struct FwdMaps { CUtensorMap q, k, v, h; };

struct alignas(16) GateBuf {       // indexed by tile row
  float u2[L];         // u * log2e + log2(scale)   (exponent of P, scale folded in)
  float M2[L];         // M * log2e
  float wq[L];         // exp(m_prev - M) * scale   (row weight of the inter-chunk term)
  float mrow[L];       // m_t = b_t + M_t
  float kw[L];         // exp(u - M_L)              (key weight for the state update)
  float decay;         // exp(m_prev - M_L)
  float m_next;        // g + M_L
};

template <int DH>
struct Smem {
  static constexpr int KT = DH / 64;               // 64-wide tiles per operand
  static constexpr int TILE_C = DH * 128;          // bytes of one [DH rows][64] Cb tile
  alignas(1024) uint8_t q[KT * TILE];
  alignas(1024) uint8_t k[KT * TILE];
  alignas(1024) uint8_t v[KT * TILE];
  alignas(1024) uint8_t p[2 * TILE];               // P (K-major, 2 tiles over j); h staging reuses it
  alignas(1024) uint8_t cb[KT * TILE_C];           // bf16 C, MN-major [dk][dv]
  alignas(1024) uint8_t ones[2048];                // bf16 1.0 (B operand of n += Kbar^T 1)
  GateBuf g[2];
  float n_prev[DH];
  float scan_a[4], scan_b[4];
  uint64_t bar_q, bar_k, bar_v, bar_m1, bar_kv, bar_m2;
  uint32_t tmem_base;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Gate vectors of chunk `c` into `G` (all 128 threads; thread t = scan-local index t).
template <int DH>
__device__ __forceinline__ void compute_gates(Smem<DH>& sm, GateBuf& G, const mlstm_params& p, int b, int h, int c,
                                              float m_prev, float scale) {
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  // Memory chunk and its valid rows.  Chunks are anchored at token 0 in both directions, so a
  // partial chunk always has its invalid rows at the tile's end (TMA zero-fills / clips them);
  // in reverse mode the partial chunk is simply the first one processed.
  const int NCc = (p.S + L - 1) / L;
  const int mc = p.reverse ? (NCc - 1 - c) : c;
  const int tok0 = mc * L;
  const int nvalid = min(L, p.S - tok0);
  const bool valid = t < nvalid;
  const int r = (p.reverse && valid) ? (nvalid - 1 - t) : t;   // tile row of scan-local index t
  float ii = -INFINITY, logf = 0.f;
  if (valid) {
    const int tok = tok0 + r;
    const int64_t off = (int64_t)b * p.f.stride_b + (int64_t)h * p.f.stride_h + (int64_t)tok * p.f.stride_s;
    const int64_t offi = (int64_t)b * p.i.stride_b + (int64_t)h * p.i.stride_h + (int64_t)tok * p.i.stride_s;
    logf = log_sigmoid(p.f.ptr[off]);
    ii = p.i.ptr[offi];
  }
  float bs = warp_scan_add(logf, lane);
  if (lane == 31) sm.scan_a[warp] = bs;
  __syncthreads();
  float pre = 0.f;
#pragma unroll
  for (int w = 0; w < 4; ++w) pre += (w < warp) ? sm.scan_a[w] : 0.f;
  bs += pre;
  const float g_tot = sm.scan_a[0] + sm.scan_a[1] + sm.scan_a[2] + sm.scan_a[3];
  const float u = ii - bs;
  float cm = warp_scan_max(u, lane);
  if (lane == 31) sm.scan_b[warp] = cm;
  __syncthreads();
  float cmax_all = -INFINITY;
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    if (w < warp) cm = fmaxf(cm, sm.scan_b[w]);
    cmax_all = fmaxf(cmax_all, sm.scan_b[w]);
  }
  const float M = fmaxf(m_prev, cm);
  const float ML = fmaxf(m_prev, cmax_all);
  G.u2[r] = u * LOG2E + log2f(scale);
  G.M2[r] = M * LOG2E;
  G.wq[r] = __expf(m_prev - M) * scale;
  G.mrow[r] = bs + M;
  G.kw[r] = __expf(u - ML);
  if (t == 0) {
    G.decay = __expf(m_prev - ML);
    G.m_next = g_tot + ML;
  }
  __syncthreads();
}

template <int DH>
__global__ void __launch_bounds__(NT, 1) tc_fwd_kernel(const __grid_constant__ FwdMaps maps, const mlstm_params p,
                                                       const float scale) {
  constexpr int KT = DH / 64;
  constexpr int TILE_C = Smem<DH>::TILE_C;
  constexpr uint32_t A_LBO_STATE = (DH == 128) ? TILE : 0;  // DH=64: the 2nd 64-row M block aliases the 1st
  extern __shared__ uint8_t smem_raw[];
  Smem<DH>& sm = *reinterpret_cast<Smem<DH>*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  const int tid = threadIdx.x, warp = tid >> 5;
  const int bh = blockIdx.x, b = bh / p.NH, h = bh % p.NH;
  const int S = p.S, NC = (S + L - 1) / L;
  const bool has_init = p.c_initial != nullptr;
  const bool rev = p.reverse != 0;

  if (tid == 0) {
    tma_prefetch_desc(&maps.q); tma_prefetch_desc(&maps.k); tma_prefetch_desc(&maps.v); tma_prefetch_desc(&maps.h);
    mbar_init(&sm.bar_q, 1); mbar_init(&sm.bar_k, 1); mbar_init(&sm.bar_v, 1);
    mbar_init(&sm.bar_m1, 1); mbar_init(&sm.bar_kv, 1); mbar_init(&sm.bar_m2, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, 512);
  // ones tile, zero Cb / n_prev when there is no initial state
  for (int e = tid; e < 2048 / 4; e += NT) reinterpret_cast<uint32_t*>(sm.ones)[e] = 0x3F803F80u;
  if (!has_init) {
    for (int e = tid; e < KT * TILE_C / 16; e += NT) reinterpret_cast<uint4*>(sm.cb)[e] = make_uint4(0, 0, 0, 0);
    for (int e = tid; e < DH; e += NT) sm.n_prev[e] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = sm.tmem_base;
  const uint32_t tS = tm, tG = tm + 128, tC = tm + 128 + DH, tN = tm + 128 + 2 * DH;
  const uint32_t lane_sel = (uint32_t)(warp * 32) << 16;

  auto tok0_of = [&](int c) { return (rev ? (NC - 1 - c) : c) * L; };
  auto issue_loads = [&](uint8_t* dst, const CUtensorMap* map, uint64_t* bar, int c) {
    mbar_arrive_expect_tx(bar, KT * TILE);
    for (int kt = 0; kt < KT; ++kt) tma_load_4d(dst + kt * TILE, map, bar, kt * 64, tok0_of(c), h, b);
  };

  float m_prev = p.m_initial ? p.m_initial[bh] : 0.f;
  if (tid == 0) {
    issue_loads(sm.q, &maps.q, &sm.bar_q, 0);
    issue_loads(sm.k, &maps.k, &sm.bar_k, 0);
    issue_loads(sm.v, &maps.v, &sm.bar_v, 0);
  }
  compute_gates<DH>(sm, sm.g[0], p, b, h, 0, m_prev, scale);

  if (has_init) {  // TMEM C <- decay_0 * C_0 ; Cb <- bf16(C_0) ; n likewise
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
      sm.n_prev[tid] = n0;
      float r16[32];
#pragma unroll
      for (int x = 0; x < 32; ++x) r16[x] = n0 * d0;
      tmem_st32(tN + lane_sel, r16);  // tN occupies 16 columns; the next 16 are unused scratch
      tmem_st_wait();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  for (int c = 0; c < NC; ++c) {
    GateBuf& G = sm.g[c & 1];
    GateBuf& Gn = sm.g[(c + 1) & 1];
    const uint32_t ph = c & 1;
    const int tok0 = tok0_of(c);

    // ---- MMA1: S = Q K^T, G = Q Cb ------------------------------------------------------
    mbar_wait(&sm.bar_q, ph);
    mbar_wait(&sm.bar_k, ph);
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idS = make_idesc_bf16(128, 128, 0, 0);
      for (int ks = 0; ks < DH / 16; ++ks) {
        const uint32_t off = (ks >> 2) * TILE + (ks & 3) * 32;
        umma_bf16_ss(tS, make_sdesc(smem_u32(sm.q) + off, 16, 1024), make_sdesc(smem_u32(sm.k) + off, 16, 1024), idS, ks > 0);
      }
      constexpr uint32_t idG = make_idesc_bf16(128, DH, 0, 1);
      for (int ks = 0; ks < DH / 16; ++ks) {
        const uint32_t off = (ks >> 2) * TILE + (ks & 3) * 32;
        umma_bf16_ss(tG, make_sdesc(smem_u32(sm.q) + off, 16, 1024),
                     make_sdesc(smem_u32(sm.cb) + ks * 2048, TILE_C, 1024), idG, ks > 0);
      }
      umma_commit(&sm.bar_m1);
    }
    // ---- in the MMA shadow: gates of the next chunk, q . n_prev --------------------------
    if (c + 1 < NC) compute_gates<DH>(sm, Gn, p, b, h, c + 1, G.m_next, scale);
    float qn = 0.f;
    for (int kt = 0; kt < KT; ++kt) {
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8) {
        const uint4 w = *reinterpret_cast<const uint4*>(sm.q + kt * TILE + swz128(tid, c8 * 8));
        const float4 n0 = *reinterpret_cast<const float4*>(&sm.n_prev[kt * 64 + c8 * 8]);
        const float4 n1 = *reinterpret_cast<const float4*>(&sm.n_prev[kt * 64 + c8 * 8 + 4]);
        const __nv_bfloat162* qq = reinterpret_cast<const __nv_bfloat162*>(&w);
        float2 a = __bfloat1622float2(qq[0]), bq = __bfloat1622float2(qq[1]), cq = __bfloat1622float2(qq[2]),
               dq = __bfloat1622float2(qq[3]);
        qn += a.x * n0.x + a.y * n0.y + bq.x * n0.z + bq.y * n0.w + cq.x * n1.x + cq.y * n1.y + dq.x * n1.z + dq.y * n1.w;
      }
    }
    mbar_wait(&sm.bar_m1, ph);
    tc_fence_after();

    // ---- Kbar = kw * K, in place ---------------------------------------------------------
    for (int it = 0; it < KT * TILE / 16 / NT; ++it) {
      const uint32_t o = (uint32_t)(tid + it * NT) * 16u;
      const int row = (o >> 7) & (L - 1);
      const float s = G.kw[row];
      uint4 w = *reinterpret_cast<uint4*>(sm.k + o);
      __nv_bfloat162* kk = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        float2 f2 = __bfloat1622float2(kk[e]);
        kk[e] = __floats2bfloat162_rn(f2.x * s, f2.y * s);
      }
      *reinterpret_cast<uint4*>(sm.k + o) = w;
    }
    fence_proxy_async_smem();
    if (tid == 0) tma_store_wait_read<0>();   // the previous chunk's h staging (= P tile) has been read
    tc_fence_before();
    __syncthreads();

    // ---- state MMAs: C += Kbar^T V, n += Kbar^T 1 ; prefetch Q(c+1) ----------------------
    if (tid == 0) {
      if (c + 1 < NC) issue_loads(sm.q, &maps.q, &sm.bar_q, c + 1);
      mbar_wait(&sm.bar_v, ph);
      tc_fence_after();
      constexpr uint32_t idC = make_idesc_bf16(128, DH, 1, 1);
      constexpr uint32_t idN = make_idesc_bf16(128, 16, 1, 1);
      const uint32_t acc0 = (c > 0 || has_init) ? 1u : 0u;
      for (int ks = 0; ks < L / 16; ++ks) {
        const uint64_t a = make_sdesc(smem_u32(sm.k) + ks * 2048, A_LBO_STATE, 1024);
        umma_bf16_ss(tC, a, make_sdesc(smem_u32(sm.v) + ks * 2048, TILE, 1024), idC, (ks > 0) ? 1u : acc0);
        umma_bf16_ss(tN, a, make_sdesc(smem_u32(sm.ones), 1024, 1024), idN, (ks > 0) ? 1u : acc0);
      }
      umma_commit(&sm.bar_kv);
    }

    // ---- P = S * exp2(u2_j - M2_t), causal; row sums -------------------------------------
    const float M2t = G.M2[tid];
    float rowsum = 0.f;
#pragma unroll 1
    for (int cbk = 0; cbk < 4; ++cbk) {
      // forward: keys j <= t ; reverse: keys j >= t   (tile rows; warp w owns rows 32w..32w+31)
      const bool full = rev ? (cbk > warp) : (cbk < warp);
      const bool diag = (cbk == warp);
      uint32_t packed[16];
      if (full || diag) {
        float s[32];
        tmem_ld32(tS + lane_sel + cbk * 32, s);
        tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; x += 4) {
          const float4 u4 = *reinterpret_cast<const float4*>(&G.u2[cbk * 32 + x]);
          const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
          float pv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int j = cbk * 32 + x + e;
            const bool keep = full || (rev ? (j >= tid) : (j <= tid));
            pv[e] = keep ? s[x + e] * ex2(uu[e] - M2t) : 0.f;
            rowsum += pv[e];
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
        *reinterpret_cast<uint4*>(sm.p + (j >> 6) * TILE + swz128(tid, j & 63)) =
            make_uint4(packed[4 * x], packed[4 * x + 1], packed[4 * x + 2], packed[4 * x + 3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();

    // ---- MMA2: H = P V (into the S columns) ; prefetch K(c+1) ----------------------------
    if (tid == 0) {
      tc_fence_after();
      constexpr uint32_t idH = make_idesc_bf16(128, DH, 0, 1);
      for (int ks = 0; ks < L / 16; ++ks) {
        umma_bf16_ss(tS, make_sdesc(smem_u32(sm.p) + (ks >> 2) * TILE + (ks & 3) * 32, 16, 1024),
                     make_sdesc(smem_u32(sm.v) + ks * 2048, TILE, 1024), idH, ks > 0);
      }
      umma_commit(&sm.bar_m2);
      mbar_wait(&sm.bar_kv, ph);
      if (c + 1 < NC) issue_loads(sm.k, &maps.k, &sm.bar_k, c + 1);
    }
    // row normaliser (backends.py:249-252)
    const float wq = G.wq[tid];
    const float mrow = G.mrow[tid];
    const float nr = rowsum + wq * qn;
    const float Nrow = fmaxf(fabsf(nr), __expf(-mrow)) + p.eps;
    const float inv = 1.f / Nrow;
    {
      const int tok = tok0 + tid;
      if (p.n_row && tok >= 0 && tok < S) {
        p.n_row[(int64_t)bh * S + tok] = nr;
        p.m_row[(int64_t)bh * S + tok] = mrow;
      }
    }
    mbar_wait(&sm.bar_m2, ph);
    tc_fence_after();
    if (tid == 0 && c + 1 < NC) issue_loads(sm.v, &maps.v, &sm.bar_v, c + 1);

    // ---- epilogue: h = (H + wq G) / N -> bf16 staging (P tile) ---------------------------
#pragma unroll 1
    for (int cbk = 0; cbk < DH / 32; ++cbk) {
      float hi[32], gg[32];
      tmem_ld32(tS + lane_sel + cbk * 32, hi);
      tmem_ld32(tG + lane_sel + cbk * 32, gg);
      tmem_ld_wait();
#pragma unroll
      for (int x = 0; x < 32; x += 8) {
        float o[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = (hi[x + e] + wq * gg[x + e]) * inv;
        const int dv = cbk * 32 + x;
        *reinterpret_cast<uint4*>(sm.p + (dv >> 6) * TILE + swz128(tid, dv & 63)) =
            make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
      }
    }
    // ---- state pass: Cb <- bf16(C); C <- decay_next C; n likewise ------------------------
    const bool last = (c + 1 == NC);
    const float dnext = last ? 1.f : Gn.decay;
    if (warp < DH / 32) {
#pragma unroll 1
      for (int cbk = 0; cbk < DH / 32; ++cbk) {
        float r[32];
        tmem_ld32(tC + lane_sel + cbk * 32, r);
        tmem_ld_wait();
        if (last && p.c_last) {
          float* dst = p.c_last + ((int64_t)bh * DH + tid) * DH + cbk * 32;
#pragma unroll
          for (int x = 0; x < 32; x += 4) *reinterpret_cast<float4*>(dst + x) = make_float4(r[x], r[x + 1], r[x + 2], r[x + 3]);
        }
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
      float rn[16];
      tmem_ld16(tN + lane_sel, rn);
      tmem_ld_wait();
      sm.n_prev[tid] = rn[0];
      if (last && p.n_last) p.n_last[(int64_t)bh * DH + tid] = rn[0];
      if (!last) {
        float r32[32];
#pragma unroll
        for (int x = 0; x < 32; ++x) r32[x] = rn[0] * dnext;
        tmem_st32(tN + lane_sel, r32);
        tmem_st_wait();
      }
    }
    if (last && p.m_last && tid == 0) p.m_last[bh] = G.m_next;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      for (int kt = 0; kt < KT; ++kt) tma_store_4d(&maps.h, sm.p + kt * TILE, kt * 64, tok0, h, b);
      tma_store_commit();
    }
  }

  if (tid == 0) tma_store_wait_all<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <int DH>
int launch_fwd(const mlstm_params& p, cudaStream_t st) {
  FwdMaps maps;
  int r = 0;
  r |= make_act_tmap(&maps.q, p.q.ptr, p.B, p.NH, p.S, DH, p.q.stride_b, p.q.stride_h, p.q.stride_s, L);
  r |= make_act_tmap(&maps.k, p.k.ptr, p.B, p.NH, p.S, DH, p.k.stride_b, p.k.stride_h, p.k.stride_s, L);
  r |= make_act_tmap(&maps.v, p.v.ptr, p.B, p.NH, p.S, DH, p.v.stride_b, p.v.stride_h, p.v.stride_s, L);
  r |= make_act_tmap(&maps.h, p.h.ptr, p.B, p.NH, p.S, DH, p.h.stride_b, p.h.stride_h, p.h.stride_s, L);
  if (r) {
    set_error("cuTensorMapEncodeTiled failed (%d): pointers must be 16-byte aligned, strides multiples of 8 elements", r);
    return r == -1 ? MLSTM_ERR_NO_DEVICE : MLSTM_ERR_INVALID_ARG;
  }
  const size_t smem = sizeof(Smem<DH>) + 1024;
  cudaError_t e = cudaFuncSetAttribute(tc_fwd_kernel<DH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(tc_fwd, %zu B): %s", smem, cudaGetErrorString(e));
    return MLSTM_ERR_CUDA;
  }
  tc_fwd_kernel<DH><<<dim3(p.B * p.NH), dim3(NT), smem, st>>>(maps, p, resolve_scale(p));
  count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("tc_fwd launch failed: %s", cudaGetErrorString(e));
    return MLSTM_ERR_CUDA;
  }
  return MLSTM_OK;
}

}  // namespace

int tc_fwd(const mlstm_params& p, cudaStream_t st) {
  if (p.DHQK == 64) return launch_fwd<64>(p, st);
  return launch_fwd<128>(p, st);
}

}  // namespace mlstm

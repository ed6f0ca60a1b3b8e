// Fused single-walk backward at DH = 128 (the north-star shape: B=32, NH=4, S=1600 / 6400, DH=128).
//
// One CTA per (batch, head) walks the 128-token chunks once, in reverse scan order, and produces dq, dk, dv, di, df
// together: q, k, v, dh are read exactly once (TMA), h once (row dots only, plain loads), the forward's entry states Cs
// once; nothing of the adjoint state ever goes through HBM.  Against the chunk-parallel family (mlstm_tc_bwd.cu: kernels
// A, SB, B1 | B2, DF) this removes the dCs round trip, the second and third reads of q, k, v, dh and the partial-sum
// workspace: ~480 MB of DRAM traffic instead of ~920 MB at B32 NH4 S1600.
//
// At DH = 128 every operand tile is 32 KB and every accumulator 128 TMEM columns, so the DH = 64 layout (gated tiles in
// shared memory, one accumulator per product) does not fit.  What makes it fit:
//   * the three gated tiles live in TMEM only.  dS is needed with rows = queries (dQ = dS K) and with rows = keys
//     (dK = dS^T Q); a TMEM A operand has its M rows on the lanes, so Z = dH V^T and Z^T = V dH^T are both computed
//     (one extra 128x128x128 MMA) and gated separately with the same arithmetic — what kernels A and B2 do;
//   * the inter-chunk products are folded into the intra-chunk accumulators through row-scaled bf16 copies of their A
//     operand, also held in TMEM:  dQ = [dS' | dHs] [K ; Cs^T],  dV = [E^T | Ks] [dH ; dCb],  dK = [dS'^T | Vs] [Q ; dCb^T]
//     with dHs = (w s / N) dH, Ks = kw K, Vs = kw V.  No G / Ik / Iv accumulators;
//   * the three chains share two 128-column accumulators (X0: Z -> dQ, then Z^T -> dK; X1: S^T -> dV) and one pair of
//     64-column operand regions (Pg gated tile, Ps scaled copy);
//   * dC += Q^T dHs uses dh scaled IN PLACE once its last unscaled reader (dV = E^T dH) has completed, so no tile is
//     double buffered and no scratch tile exists: 6 x 32 KB (q, k, v, dh, Cs, dCb) + gate vectors;
//   * outputs are staged for TMA stores in the Cs tile, which is dead once dQ has its state term: dq, dv and dk pass
//     through it one after the other (a third of a step apart), so k can be refilled for the next chunk right behind
//     its last MMA; the row dot K = k.dk, which comes after that, reads its k rows from global memory (L2 hits).
//
// Per chunk (reverse scan order; chains Q, V, K):
//   in(Q)   X0 = Z   = dH V^T            in(V)  X1 = S^T = K Q^T           (issued one step ahead)
//   S1      dn_t = dnf_t (dh_t . h_t) ; Ps = dHs, Pg = dS' = s (Z / N + dn) D            out(Q)  X0 = Ps Cs^T + Pg K
//           Ks, E^T = s S^T D / N into registers (the two tiles balance the causal work across the schedulers)
//   S2      dn_state column sums (under out(Q)) ; Ps = Ks, Pg = E^T                       out(V)  X1 = Ps dCb  + Pg dH
//   S3      dq = X0 + s w dn n_prev  -> stage, R = q.dq                                   in(K)   X0 = Z^T = V dH^T
//   S4      Ps = Vs,  Pg = dS'^T ; dh <- dHs in place                      dC += Q^T dHs ; out(K)  X0 = Ps dCb^T + Pg Q
//   S5      dv = X1 -> stage
//   S6      dk = X0 + kw dn_state -> stage, K = k.dk ; state pass: dCb <- bf16(dC), dC <- decay dC
//   gate warp, one step behind: di = K, df = sigmoid(-f) (suffix sum of R - K carried along the walk)
#include "tc_common.cuh"

namespace mlstm {
namespace {

using namespace tc;

constexpr int DH = 128;
constexpr int KT = DH / 64;            // 64-wide sub-tiles per operand
constexpr int TILE2 = KT * TILE;       // bytes of one [128][128] bf16 operand (two swizzled sub-tiles)
constexpr int TILE_C = DH * 128;       // bytes of one [DH rows][64] sub-tile of a state matrix
#ifndef MLSTM_F128_LB
#define MLSTM_F128_LB 128
#endif
constexpr int LB = MLSTM_F128_LB;      // rows per TMA load box (a tile is fetched as L / LB boxes per 64-column half).  Measured
                                       // at B32 NH4 S1600: 128 rows 191.5 us, 64: 192.0, 32: 196.0, 16: 205.9 — more, smaller
                                       // boxes only cost issue slots; the ~4 us a tile load takes under load is not per-box work

#ifdef MLSTM_TIMELINE
#define TLG(k) do { if (blockIdx.x == 0 && c < 6 && (threadIdx.x == 0 || threadIdx.x == CT)) \
    tlg[c * 64 + (threadIdx.x == 0 ? 0 : 32) + (k)] = clock64(); } while (0)
#else
#define TLG(k) do { } while (0)
#endif

struct F128Maps { CUtensorMap q, k, v, dh, cs, dq, dk, dv; };

struct SmemF128 {
  alignas(1024) uint8_t q[TILE2];
  alignas(1024) uint8_t k[TILE2];
  alignas(1024) uint8_t v[TILE2];
  alignas(1024) uint8_t dh[TILE2];           // scaled in place to dHs = (w s / N) dH for the state update
  alignas(1024) uint8_t cs[KT * TILE_C];     // forward entry state of the chunk, bf16 [dk][dv]; then staging of dq, dv, dk in turn
  alignas(1024) uint8_t dcb[KT * TILE_C];    // bf16 copy of the adjoint state leaving the chunk, [dk][dv]
  GateBuf g[3];
  alignas(16) float ns[3][DH];               // n_prev of the chunk (ring with g)
  alignas(16) float nvec[DH];                // dn_state leaving the chunk
  alignas(16) float ncoef[L];                // (w s dn)_t
  alignas(16) float rowscale[L];             // (w s / N)_t
  alignas(16) float npart[16][DH];           // S2 -> S3: per-warp partial column sums of Q^T (w s dn)
  float partR[4][L], partK[4][L];            // q . dq, k . dk partials per 32-column block
  alignas(16) float dbuf[2][2][L];           // [step parity][R - K | K][row]: handed to the gate warp for di, df
  uint64_t bar_q, bar_k, bar_v, bar_dh, bar_cs, bar_in[3], bar_out[3], bar_dc;
  uint32_t tmem_base;
};

static_assert(sizeof(SmemF128) <= 232448, "fused128 backward: shared memory budget (227 KB per CTA) exceeded");

// 32 columns [32 cb, 32 cb + 32) of row `row` of a [128][128] swizzled bf16 operand (two sub-tiles) -> fp32
__device__ __forceinline__ void tile_row32_128(const uint8_t* tile, int row, int cb, float (&out)[32]) {
#pragma unroll
  for (int x = 0; x < 32; x += 8) {
    const int col = cb * 32 + x;
    const uint4 w = *reinterpret_cast<const uint4*>(tile + (col >> 6) * TILE + swz128(row, col & 63));
    const __nv_bfloat162* qq = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f2 = __bfloat1622float2(qq[e]);
      out[x + 2 * e] = f2.x;
      out[x + 2 * e + 1] = f2.y;
    }
  }
}
// the same block scaled by s and rounded to packed bf16 pairs: the A operand layout tcgen05.st puts into TMEM
// (one packed bf16 multiply per pair: the scale is rounded to bf16 first, a 2^-9 relative error common to the row —
// the same order as the rounding of the scaled operand itself, which the MMA needs in bf16 anyway)
__device__ __forceinline__ void scaled_block(const uint8_t* tile, int row, int cb, float s, uint32_t (&pk)[16]) {
  const __nv_bfloat162 s2 = __float2bfloat162_rn(s);
#pragma unroll
  for (int x = 0; x < 32; x += 8) {
    const int col = cb * 32 + x;
    const uint4 w = *reinterpret_cast<const uint4*>(tile + (col >> 6) * TILE + swz128(row, col & 63));
    const __nv_bfloat162* qq = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const __nv_bfloat162 r = __hmul2(qq[e], s2);
      pk[x / 2 + e] = *reinterpret_cast<const uint32_t*>(&r);
    }
  }
}
// rows of a [128][128] swizzled bf16 operand scaled in place by rowscale[row], same arithmetic as scaled_block
__device__ __forceinline__ void scale_rows_bf16(uint8_t* tile, const float* rowscale, int tid) {
#pragma unroll
  for (int it = 0; it < TILE2 / 16 / CT; ++it) {
    const uint32_t o = (uint32_t)(tid + it * CT) * 16u;
    const __nv_bfloat162 s2 = __float2bfloat162_rn(rowscale[(o >> 7) & (L - 1)]);
    uint4 w = *reinterpret_cast<uint4*>(tile + o);
    __nv_bfloat162* kk = reinterpret_cast<__nv_bfloat162*>(&w);
#pragma unroll
    for (int e = 0; e < 4; ++e) kk[e] = __hmul2(kk[e], s2);
    *reinterpret_cast<uint4*>(tile + o) = w;
  }
}
__device__ __forceinline__ void stage_block(uint8_t* tile, int row, int cb, const uint32_t (&pk)[16]) {
#pragma unroll
  for (int x4 = 0; x4 < 4; ++x4) {
    const int col = cb * 32 + x4 * 8;
    *reinterpret_cast<uint4*>(tile + (col >> 6) * TILE + swz128(row, col & 63)) =
        make_uint4(pk[4 * x4], pk[4 * x4 + 1], pk[4 * x4 + 2], pk[4 * x4 + 3]);
  }
}

// (18 warps are allocated as 20: 96 registers per thread is the ceiling, ptxas finds it from the launch bounds)
// dk_true / dv_true: columns of a k / h row that exist in memory (< 128 for a zero-padded problem running on the caller's
// narrow tensors, tc_tmap.cuh: ExtentOverride): the two places that read rows with plain loads supply the zeros themselves.
// NARROW = false is the full-width kernel without those tests (the BASELINE shapes).
template <bool NARROW>
__global__ void __launch_bounds__(NT, 1) tc_bwd_fused128_kernel(const __grid_constant__ F128Maps maps, const mlstm_params p,
                                                                const float scale, const int dk_true, const int dv_true) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemF128& sm = *reinterpret_cast<SmemF128*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
#ifdef MLSTM_TIMELINE
  long long* tlg = reinterpret_cast<long long*>(p.workspace);
#endif

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool compute = tid < CT, issuer = tid == CT, gatew = tid >= GT0;
  const int rg = warp & 3, cq = compute ? (warp >> 2) : 0, row = rg * 32 + lane;
  const int bh = blockIdx.x, b = bh / p.NH, h = bh % p.NH;
  const int S = p.S, NC = num_chunks(S);
  const bool rev = p.reverse != 0;
  const StateLayout slay(p.B, p.NH, S, DH);
  const float* ns_all = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(p.states) + slay.ns_off) + (size_t)bh * NC * DH;
  const float l2s = log2f(scale);
  const __nv_bfloat16* h_base = reinterpret_cast<const __nv_bfloat16*>(p.h.ptr) + (int64_t)b * p.h.stride_b + (int64_t)h * p.h.stride_h;
  const __nv_bfloat16* k_base = reinterpret_cast<const __nv_bfloat16*>(p.k.ptr) + (int64_t)b * p.k.stride_b + (int64_t)h * p.k.stride_h;

  if (issuer) {
    tma_prefetch_desc(&maps.q); tma_prefetch_desc(&maps.k); tma_prefetch_desc(&maps.v); tma_prefetch_desc(&maps.dh);
    tma_prefetch_desc(&maps.cs); tma_prefetch_desc(&maps.dq); tma_prefetch_desc(&maps.dk); tma_prefetch_desc(&maps.dv);
    mbar_init(&sm.bar_q, 1); mbar_init(&sm.bar_k, 1); mbar_init(&sm.bar_v, 1); mbar_init(&sm.bar_dh, 1); mbar_init(&sm.bar_cs, 1);
    for (int x = 0; x < 3; ++x) { mbar_init(&sm.bar_in[x], 1); mbar_init(&sm.bar_out[x], 1); }
    mbar_init(&sm.bar_dc, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, 512);
  for (int e = tid; e < KT * TILE_C / 16; e += NT) reinterpret_cast<uint4*>(sm.dcb)[e] = make_uint4(0, 0, 0, 0);
  for (int e = tid; e < DH; e += NT) sm.nvec[e] = 0.f;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = sm.tmem_base;
  const uint32_t tX0 = tm, tX1 = tm + 128, tdC = tm + 256, tPg = tm + 384, tPs = tm + 448;
  const uint32_t lane_sel = (uint32_t)(rg * 32) << 16;

  // processing step c handles scan chunk sc = NC-1-c, memory chunk mem_chunk(sc)
  auto sc_of = [&](int c) { return NC - 1 - c; };
  auto tok0_of = [&](int c) { return mem_chunk(sc_of(c), NC, rev) * L; };
  auto load_act = [&](uint8_t* dst, const CUtensorMap* map, uint64_t* bar, int c) {
    mbar_arrive_expect_tx(bar, TILE2);
#pragma unroll
    for (int r0 = 0; r0 < L; r0 += LB)
      for (int kt = 0; kt < KT; ++kt) tma_load_4d(dst + kt * TILE + r0 * 128, map, bar, kt * 64, tok0_of(c) + r0, h, b);
  };
  auto load_cs = [&](int c) {
    mbar_arrive_expect_tx(&sm.bar_cs, KT * TILE_C);
    for (int kt = 0; kt < KT; ++kt) tma_load_2d(sm.cs + kt * TILE_C, &maps.cs, &sm.bar_cs, kt * 64, (bh * NC + sc_of(c)) * DH);
  };   // (the state rows are contiguous in memory: one box per half)
  auto gates_of = [&](int c) {   // one warp
    const int slot = c % 3;
    gates_warp_bwd(sm.g[slot], p, b, h, bh, mem_chunk(sc_of(c), NC, rev), lane, nullptr);
    for (int d = lane; d < DH; d += 32) sm.ns[slot][d] = ns_all[(size_t)sc_of(c) * DH + d];
    __syncwarp();
  };

  // operand descriptors: every tile is single buffered, so all of them are loop invariant
  const uint64_t dQk = make_sdesc(smem_u32(sm.q), 16, 1024), dQmn = make_sdesc(smem_u32(sm.q), TILE, 1024);
  const uint64_t dKk = make_sdesc(smem_u32(sm.k), 16, 1024), dKmn = make_sdesc(smem_u32(sm.k), TILE, 1024);
  const uint64_t dVk = make_sdesc(smem_u32(sm.v), 16, 1024);
  const uint64_t dHk = make_sdesc(smem_u32(sm.dh), 16, 1024), dHmn = make_sdesc(smem_u32(sm.dh), TILE, 1024);
  const uint64_t dCsk = make_sdesc(smem_u32(sm.cs), 16, 1024);
  const uint64_t dCbk = make_sdesc(smem_u32(sm.dcb), 16, 1024), dCbmn = make_sdesc(smem_u32(sm.dcb), TILE_C, 1024);
  constexpr uint32_t idKK = make_idesc_bf16(128, 128, 0, 0);   // A K-major, B K-major
  constexpr uint32_t idKM = make_idesc_bf16(128, 128, 0, 1);   // A K-major (smem or TMEM), B MN-major
  constexpr uint32_t idMM = make_idesc_bf16(128, 128, 1, 1);   // A MN-major, B MN-major

  // input product of a chain: 0 (Q): X0 = Z = dH V^T, 1 (V): X1 = S^T = K Q^T, 2 (K): X0 = Z^T = V dH^T
  auto issue_in = [&](int chain) {
    if (chain == 0) {
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) umma_bf16_ss(tX0, dHk + kstep(ks), dVk + kstep(ks), idKK, ks > 0);
    } else if (chain == 1) {
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) umma_bf16_ss(tX1, dKk + kstep(ks), dQk + kstep(ks), idKK, ks > 0);
    } else {
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) umma_bf16_ss(tX0, dVk + kstep(ks), dHk + kstep(ks), idKK, ks > 0);
    }
    umma_commit(&sm.bar_in[chain]);
  };
  // output product of a chain: state term from the scaled copy Ps, then the intra-chunk term from the gated tile Pg
  auto issue_out = [&](int chain) {
    if (chain == 0) {          // dQ = dHs Cs^T + dS' K
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) umma_bf16_ts(tX0, tPs + ks * 8, dCsk + kstep(ks, TILE_C), idKK, ks > 0);
#pragma unroll
      for (int ks = 0; ks < L / 16; ++ks) umma_bf16_ts(tX0, tPg + ks * 8, dKmn + mnstep(ks), idKM, 1u);
    } else if (chain == 1) {   // dV = Ks dCb + E^T dH
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) umma_bf16_ts(tX1, tPs + ks * 8, dCbmn + mnstep(ks), idKM, ks > 0);
#pragma unroll
      for (int ks = 0; ks < L / 16; ++ks) umma_bf16_ts(tX1, tPg + ks * 8, dHmn + mnstep(ks), idKM, 1u);
    } else {                   // dK = Vs dCb^T + dS'^T Q
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) umma_bf16_ts(tX0, tPs + ks * 8, dCbk + kstep(ks, TILE_C), idKK, ks > 0);
#pragma unroll
      for (int ks = 0; ks < L / 16; ++ks) umma_bf16_ts(tX0, tPg + ks * 8, dQmn + mnstep(ks), idKM, 1u);
    }
    umma_commit(&sm.bar_out[chain]);
  };

  if (issuer) {
    load_act(sm.dh, &maps.dh, &sm.bar_dh, 0); load_act(sm.v, &maps.v, &sm.bar_v, 0);
    load_act(sm.q, &maps.q, &sm.bar_q, 0); load_act(sm.k, &maps.k, &sm.bar_k, 0);
    load_cs(0);
  }
  if (gatew) gates_of(0);
  if (warp == 1 && NC > 1) gates_of(1);   // the compute warps are idle here: the first two chunks' gates side by side
  __syncthreads();
  if (issuer) {
    mbar_wait(&sm.bar_dh, 0); mbar_wait(&sm.bar_v, 0);
    tc_fence_after();
    issue_in(0);
    mbar_wait(&sm.bar_q, 0); mbar_wait(&sm.bar_k, 0);
    tc_fence_after();
    issue_in(1);
  }

  // di, df of processing step c (gate warp, one step behind the compute warps): di_j = K_j, df_j = sigmoid(-f_j) * (suffix
  // sum in scan order of R - K, carried along the walk).  Lane l owns tile rows 4l..4l+3.
  float df_carry = 0.f;
  auto scan_of = [&](int c) {
    const GateBuf& Gs = sm.g[c % 3];
    const float4 d4 = *reinterpret_cast<const float4*>(&sm.dbuf[c & 1][0][lane * 4]);
    const float4 k4 = *reinterpret_cast<const float4*>(&sm.dbuf[c & 1][1][lane * 4]);
    const float dB[4] = {d4.x, d4.y, d4.z, d4.w}, Kj[4] = {k4.x, k4.y, k4.z, k4.w};
    float pre[4], run = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) { run += dB[e]; pre[e] = run; }
    const float incl = warp_scan_add(run, lane);
    const float excl = incl - run, tot = __shfl_sync(0xffffffffu, incl, 31);
    const int tok0 = tok0_of(c);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = lane * 4 + e, tok = tok0 + r;
      const float pin = excl + pre[e];
      const float suf = rev ? pin : (tot - pin + dB[e]);
      if (tok < S) {
        const float i_raw = p.i.ptr[(int64_t)b * p.i.stride_b + (int64_t)h * p.i.stride_h + (int64_t)tok * p.i.stride_s];
        p.di.ptr[(int64_t)b * p.di.stride_b + (int64_t)h * p.di.stride_h + (int64_t)tok * p.di.stride_s] = Kj[e] * igate_dlog(p, i_raw);
        p.df.ptr[(int64_t)b * p.df.stride_b + (int64_t)h * p.df.stride_h + (int64_t)tok * p.df.stride_s] = (suf + df_carry) * Gs.sig[r];
      }
    }
    df_carry += tot;
  };

  float nstate = 0.f;   // thread dk < DH: decayed dn_state entering the step
  for (int c = 0; c < NC; ++c) {
    const uint32_t ph = c & 1;
    const bool last = (c + 1 == NC);
    if (gatew) {
      if (c > 0) scan_of(c - 1);             // before gates_of(c + 2) reuses that chunk's ring slot
      if (c + 2 < NC) gates_of(c + 2);
      named_sync(7, GT0);   // with the compute warps: gates two steps ahead are complete, step c's R - K, K rows are published
      continue;
    }
    const GateBuf& G = sm.g[c % 3];
    const GateBuf& Gn = sm.g[(c + 1) % 3];
    const float* nsv = sm.ns[c % 3];
    const int tok0 = tok0_of(c);
    const int tok = tok0 + row;
    const bool row_ok = compute && tok < S;
    const bool fullT = rev ? (cq > rg) : (cq < rg);   // tiles with rows = queries t (dS'): full blocks
    const bool fullJ = rev ? (cq < rg) : (cq > rg);   // tiles with rows = keys j (E^T, dS'^T)
    const bool diag = (cq == rg);
    uint32_t pg[16], ps[16];
    float dn_row = 0.f;
    TLG(0);

    // ---- S1: dn_t = dnf_t (dh_t . h_t) ; chain Q operands straight into TMEM ; chain V operands into registers.  The two
    //      gated tiles have opposite orientations (rows = queries / rows = keys), so every scheduler (= TMEM lane quadrant)
    //      gets three full and two diagonal 32x32 blocks between them.
    if (compute) {
      // h rows for the row dots dh_t . h_t, read the way they lie in memory: a half-warp takes one 256-byte row (16 bytes per
      // lane), warp w owns tile rows 8w .. 8w+7 — full cache lines per request (a thread reading its own row's 64 bytes, 32
      // rows per request, serialises in the L1 and holds up the control warp's MMA / TMA issue behind it)
      uint4 hw[4];
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int t_ = tok0 + warp * 8 + 2 * x + (lane >> 4);
        hw[x] = (t_ < S && (!NARROW || (lane & 15) * 8 < dv_true)) ? *reinterpret_cast<const uint4*>(h_base + (int64_t)t_ * p.h.stride_s + (lane & 15) * 8)
                                                     : make_uint4(0, 0, 0, 0);
      }
      mbar_wait(&sm.bar_dh, ph);
      TLG(1);
      const float rs = G.w[row] * scale * G.invN[row];
      scaled_block(sm.dh, row, cq, rs, ps);             // Ps = dHs = (w s / N) dH: the same rounding as the in-place copy in S4
      // Pg / Ps are free: their last reader, out(K) of the previous step, was waited for in that step's S6
      tmem_st16(tPs + lane_sel + cq * 16, ps);
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int r_ = warp * 8 + 2 * x + (lane >> 4), ch = lane & 15;
        const uint4 wd = *reinterpret_cast<const uint4*>(sm.dh + (ch >> 3) * TILE + swz128(r_, (ch & 7) * 8));
        const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&hw[x]);
        const __nv_bfloat162* dd = reinterpret_cast<const __nv_bfloat162*>(&wd);
        float part = 0.f;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = __bfloat1622float2(hh[e]), c2 = __bfloat1622float2(dd[e]);
          part = fmaf(a.x, c2.x, fmaf(a.y, c2.y, part));
        }
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if (ch == 0) {
          const float dn_ = G.dnf[r_] * part;
          sm.g[c % 3].dn[r_] = dn_;                     // read by row (chain Q) and by column (chain K)
          sm.ncoef[r_] = G.w[r_] * scale * dn_;
          sm.rowscale[r_] = G.w[r_] * scale * G.invN[r_];
        }
      }
      named_sync(3, CT);
      dn_row = G.dn[row];
      TLG(2);
      // Pg = dS'[t][j] = (Z invN_t + dn_t) 2^(u2_j + log2 s - M2_t), keep j <= t (reverse: j >= t)
      mbar_wait(&sm.bar_in[0], ph);
      tc_fence_after();
      if (fullT || diag) {
        float z[32];
        tmem_ld32(tX0 + lane_sel + cq * 32, z);
        const uint32_t bits = causal_bits(fullT, !rev, lane);
        tmem_ld_wait();
        const float M2t = G.M2[row] - l2s, invN = G.invN[row];
#pragma unroll
        for (int x = 0; x < 32; x += 4) {
          const float4 u4 = *reinterpret_cast<const float4*>(&G.u2[cq * 32 + x]);
          const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
          float ds[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool keep = (bits >> (x + e)) & 1u;
            ds[e] = keep ? fmaf(z[x + e], invN, dn_row) * ex2(uu[e] - M2t) : 0.f;
          }
          pg[x / 2] = pack_bf16x2(ds[0], ds[1]); pg[x / 2 + 1] = pack_bf16x2(ds[2], ds[3]);
        }
      } else {
#pragma unroll
        for (int x = 0; x < 16; ++x) pg[x] = 0u;
      }
      tmem_st16(tPg + lane_sel + cq * 16, pg);
      TLG(3);
      // chain V: Ps = Ks = kw K, Pg = E^T[j][t] = S^T 2^(u2_j + log2 s - c2_t), keep t >= j (reverse: t <= j) — kept in
      // registers until out(Q) has consumed the chain Q operands
      mbar_wait(&sm.bar_k, ph);
      scaled_block(sm.k, row, cq, G.kw[row], ps);
      mbar_wait(&sm.bar_in[1], ph);
      tc_fence_after();
      if (fullJ || diag) {
        float s_[32];
        tmem_ld32(tX1 + lane_sel + cq * 32, s_);
        const uint32_t bits = causal_bits(fullJ, rev, lane);
        tmem_ld_wait();
        const float u2j = G.u2[row] + l2s;
#pragma unroll
        for (int x = 0; x < 32; x += 4) {
          const float4 c4 = *reinterpret_cast<const float4*>(&G.c2[cq * 32 + x]);
          const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
          float ev[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool keep = (bits >> (x + e)) & 1u;
            ev[e] = keep ? s_[x + e] * ex2(u2j - cc[e]) : 0.f;
          }
          pg[x / 2] = pack_bf16x2(ev[0], ev[1]); pg[x / 2 + 1] = pack_bf16x2(ev[2], ev[3]);
        }
      } else {
#pragma unroll
        for (int x = 0; x < 16; ++x) pg[x] = 0u;
      }
      tmem_st_wait();
      TLG(4);
    }
    tc_fence_before();
    named_sync(2, GT0);
    TLG(5);
    if (issuer) {
      tc_fence_after();
      mbar_wait(&sm.bar_cs, ph);
      mbar_wait(&sm.bar_k, ph);
      tc_fence_after();
      TLG(22);
      issue_out(0);
      // the next chunk's tiles towards L2 now: their TMA loads, issued as each tile dies later in this step, then find them
      // there (191.5 -> 186.7 us at B32 NH4 S1600)
      if (!last) {
        for (int kt = 0; kt < KT; ++kt) {
          tma_prefetch_4d(&maps.k, kt * 64, tok0_of(c + 1), h, b); tma_prefetch_4d(&maps.v, kt * 64, tok0_of(c + 1), h, b);
          tma_prefetch_4d(&maps.dh, kt * 64, tok0_of(c + 1), h, b); tma_prefetch_4d(&maps.q, kt * 64, tok0_of(c + 1), h, b);
          tma_prefetch_2d(&maps.cs, kt * 64, (bh * NC + sc_of(c + 1)) * DH);
        }
      }
    }

    // ---- S2 (in the shadow of out(Q)): dn_state column sums of Q ; then the chain V operands go to TMEM --------------------
    if (compute) {
      if (!last) {
        mbar_wait(&sm.bar_q, ph);
        // lane: 4 adjacent dk columns, warp: 8 rows -> per-warp partial sums, reduced in S3
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const int col = lane * 4;
#pragma unroll
        for (int t = warp * 8; t < warp * 8 + 8; ++t) {
          const uint2 w = *reinterpret_cast<const uint2*>(sm.q + (col >> 6) * TILE + swz128(t, col & 63));
          const float2 f0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.x));
          const float2 f1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w.y));
          const float cf = sm.ncoef[t];
          a0 = fmaf(cf, f0.x, a0); a1 = fmaf(cf, f0.y, a1); a2 = fmaf(cf, f1.x, a2); a3 = fmaf(cf, f1.y, a3);
        }
        *reinterpret_cast<float4*>(&sm.npart[warp][col]) = make_float4(a0, a1, a2, a3);
      }
      TLG(6);
      mbar_wait(&sm.bar_out[0], ph);   // out(Q) complete: Pg / Ps free, X0 = dQ, the Cs tile is dead
      tc_fence_after();
      TLG(7);
      tmem_st16(tPs + lane_sel + cq * 16, ps);
      tmem_st16(tPg + lane_sel + cq * 16, pg);
      tmem_st_wait();
    }
    tc_fence_before();
    named_sync(2, GT0);
    TLG(8);
    if (issuer) {
      tc_fence_after();
      issue_out(1);   // dCb: written by the previous step's state pass (generic proxy, fenced); dh landed before in(Q)
    }

    // ---- S3: epilogue of chain Q.  dq = X0 + s w dn n_prev ; R = q . dq ; staged in the (dead) Cs tile -----------------
    float nadd = 0.f;   // thread dk < DH: this chunk's contribution to dn_state (the partials' buffer is reused from here on)
    if (compute) {
      if (!last && tid < DH) {
#pragma unroll
        for (int pt = 0; pt < 16; ++pt) nadd += sm.npart[pt][tid];
      }
      float acc[32], qr[32];
      tmem_ld32(tX0 + lane_sel + cq * 32, acc);
      tmem_ld_wait();
      const float cf = scale * G.w[row] * dn_row;
      mbar_wait(&sm.bar_q, ph);
      tile_row32_128(sm.q, row, cq, qr);
      float psum = 0.f;
#pragma unroll
      for (int x = 0; x < 32; x += 2) {
        const float o0 = fmaf(cf, nsv[cq * 32 + x], acc[x]);
        const float o1 = fmaf(cf, nsv[cq * 32 + x + 1], acc[x + 1]);
        if (row_ok) psum = fmaf(qr[x], o0, fmaf(qr[x + 1], o1, psum));
        pg[x / 2] = pack_bf16x2(o0, o1);
      }
      sm.partR[cq][row] = psum;
      stage_block(sm.cs, row, cq, pg);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    named_sync(2, GT0);
    TLG(9);
    if (issuer) {
      tc_fence_after();
      for (int kt = 0; kt < KT; ++kt) tma_store_4d(&maps.dq, sm.cs + kt * TILE, kt * 64, tok0, h, b);
      tma_store_commit();
      issue_in(2);    // X0 = Z^T = V dH^T (v landed before in(Q))
      if (!last) load_act(sm.k, &maps.k, &sm.bar_k, c + 1);   // S^T, out(Q) complete and the Ks copies taken: k is dead
      tma_store_wait_read<0>();   // dq has left the Cs tile before this warp joins the barrier that lets dv be staged there
    }

    // ---- S4: chain K.  Ps = Vs = kw V, Pg = dS'^T[j][t] = (Z^T invN_t + dn_t) 2^(u2_j + log2 s - M2_t) ; dh <- dHs ------
    if (compute) {
      mbar_wait(&sm.bar_v, ph);
      scaled_block(sm.v, row, cq, G.kw[row], ps);
      TLG(10);
      mbar_wait(&sm.bar_in[2], ph);
      tc_fence_after();
      TLG(11);
      if (fullJ || diag) {
        float z[32];
        tmem_ld32(tX0 + lane_sel + cq * 32, z);
        const uint32_t bits = causal_bits(fullJ, rev, lane);
        tmem_ld_wait();
        const float u2j = G.u2[row] + l2s;
#pragma unroll
        for (int x = 0; x < 32; x += 4) {
          const int c0 = cq * 32 + x;
          const float4 m4 = *reinterpret_cast<const float4*>(&G.M2[c0]);
          const float4 i4 = *reinterpret_cast<const float4*>(&G.invN[c0]);
          const float4 d4 = *reinterpret_cast<const float4*>(&G.dn[c0]);
          const float mm[4] = {m4.x, m4.y, m4.z, m4.w}, in4[4] = {i4.x, i4.y, i4.z, i4.w}, dn4[4] = {d4.x, d4.y, d4.z, d4.w};
          float dsv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool keep = (bits >> (x + e)) & 1u;
            dsv[e] = keep ? fmaf(z[x + e], in4[e], dn4[e]) * ex2(u2j - mm[e]) : 0.f;
          }
          pg[x / 2] = pack_bf16x2(dsv[0], dsv[1]); pg[x / 2 + 1] = pack_bf16x2(dsv[2], dsv[3]);
        }
      } else {
#pragma unroll
        for (int x = 0; x < 16; ++x) pg[x] = 0u;
      }
      TLG(12);
      mbar_wait(&sm.bar_out[1], ph);   // out(V) complete: Pg / Ps free, X1 = dV, dh has no unscaled reader left (Z, Z^T done too)
      tc_fence_after();
      TLG(13);
      tmem_st16(tPs + lane_sel + cq * 16, ps);
      tmem_st16(tPg + lane_sel + cq * 16, pg);
      if (!last) scale_rows_bf16(sm.dh, sm.rowscale, tid);   // dHs in place: the B operand of dC += Q^T dHs
      tmem_st_wait();
    }
    fence_proxy_async_smem();
    tc_fence_before();
    named_sync(2, GT0);
    TLG(14);
    if (issuer) {
      tc_fence_after();
      if (!last) {   // state update first: it frees dh for the next chunk's load as early as possible
#pragma unroll
        for (int ks = 0; ks < L / 16; ++ks) umma_bf16_ss(tdC, dQmn + mnstep(ks), dHmn + mnstep(ks), idMM, (ks > 0 || c > 0) ? 1u : 0u);
        umma_commit(&sm.bar_dc);
      }
      issue_out(2);
      if (!last) {
        load_act(sm.v, &maps.v, &sm.bar_v, c + 1);   // Z, Z^T complete and the Vs copies taken: v is dead
        mbar_wait(&sm.bar_dc, ph);                    // state update complete: dh is dead
        TLG(23);
        load_act(sm.dh, &maps.dh, &sm.bar_dh, c + 1);
      }
    }

    // ---- S5: epilogue of chain V (dv = X1, staged in the Cs tile); k rows for K = k . dk on their way ----------------------
    uint4 kw4[4];
    if (compute) {
      {
        float acc[32];
        tmem_ld32(tX1 + lane_sel + cq * 32, acc);
        tmem_ld_wait();
        TLG(15);
#pragma unroll
        for (int x = 0; x < 32; x += 2) pg[x / 2] = pack_bf16x2(acc[x], acc[x + 1]);
        stage_block(sm.cs, row, cq, pg);
      }
      // k rows for K = k . dk (the k tile already holds the next chunk): this thread's 32-column block of its row, L2 hits,
      // issued behind the staging so nothing is live across the TMEM read; in flight across the barrier and out(K)'s tail
      if (row_ok) {
        const uint4* src = reinterpret_cast<const uint4*>(k_base + (int64_t)tok * p.k.stride_s + cq * 32);
#pragma unroll
        for (int x = 0; x < 4; ++x) kw4[x] = (!NARROW || cq * 32 + x * 8 < dk_true) ? src[x] : make_uint4(0, 0, 0, 0);
      } else {
#pragma unroll
        for (int x = 0; x < 4; ++x) kw4[x] = make_uint4(0, 0, 0, 0);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    named_sync(2, GT0);
    TLG(16);
    if (issuer) {
      tc_fence_after();
      for (int kt = 0; kt < KT; ++kt) tma_store_4d(&maps.dv, sm.cs + kt * TILE, kt * 64, tok0, h, b);
      tma_store_commit();
      mbar_wait(&sm.bar_out[2], ph);     // out(K) complete (and dC, above): with the column sums and the R dots long done, q is dead
      TLG(24);
      if (!last) load_act(sm.q, &maps.q, &sm.bar_q, c + 1);
      tma_store_wait_read<0>();          // dv has left the Cs tile
      TLG(25);
    }

    // ---- S6: epilogue of chain K.  dk = X0 + kw dn_state ; K = k . dk ; state pass ; R - K and K rows for the gate warp ----
    if (compute) {
      mbar_wait(&sm.bar_out[2], ph);
      tc_fence_after();
      TLG(17);
      float acc[32];
      tmem_ld32(tX0 + lane_sel + cq * 32, acc);
      tmem_ld_wait();
      const float kwj = G.kw[row];
      float psum = 0.f;   // fp32 products of the un-rounded dk, like R = q . dq: the rounding noise of the two cancels in df
#pragma unroll
      for (int x = 0; x < 32; x += 2) {
        const float o0 = fmaf(kwj, sm.nvec[cq * 32 + x], acc[x]);
        const float o1 = fmaf(kwj, sm.nvec[cq * 32 + x + 1], acc[x + 1]);
        const float2 kk = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(&kw4[x / 8])[(x / 2) & 3]);
        psum = fmaf(kk.x, o0, fmaf(kk.y, o1, psum));
        pg[x / 2] = pack_bf16x2(o0, o1);
      }
      sm.partK[cq][row] = psum;
      TLG(18);
    }
    // the Cs tile takes the staged dk: the control warp has seen the dv store leave it (wait_read above) before this barrier
    named_sync(5, GT0);
    if (compute) {
      TLG(19);
      stage_block(sm.cs, row, cq, pg);
      if (cq == 0) {
        const float Kj = (sm.partK[0][row] + sm.partK[1][row]) + (sm.partK[2][row] + sm.partK[3][row]);
        const float Rj = (sm.partR[0][row] + sm.partR[1][row]) + (sm.partR[2][row] + sm.partR[3][row]);
        sm.dbuf[c & 1][0][row] = row_ok ? (Rj - Kj) : 0.f;
        sm.dbuf[c & 1][1][row] = row_ok ? Kj : 0.f;
      }
      // state pass: dCb <- bf16(dC), dC <- decay_next dC ; dn_state likewise
      if (!last) {
        const float dnext = Gn.decay;
        mbar_wait(&sm.bar_dc, ph);
        tc_fence_after();
        float r[32];
        tmem_ld32(tdC + lane_sel + cq * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; x += 8) {
          const int dv = cq * 32 + x;
          *reinterpret_cast<uint4*>(sm.dcb + (dv >> 6) * TILE_C + swz128(row, dv & 63)) =
              make_uint4(pack_bf16x2(r[x], r[x + 1]), pack_bf16x2(r[x + 2], r[x + 3]), pack_bf16x2(r[x + 4], r[x + 5]),
                         pack_bf16x2(r[x + 6], r[x + 7]));
        }
#pragma unroll
        for (int x = 0; x < 32; ++x) r[x] *= dnext;
        tmem_st32(tdC + lane_sel + cq * 32, r);
        tmem_st_wait();
        if (tid < DH) {
          const float nv = nstate + nadd;
          sm.nvec[tid] = nv;
          nstate = nv * dnext;
        }
      }
      TLG(20);
    }
    fence_proxy_async_smem();
    tc_fence_before();
    named_sync(2, GT0);
    TLG(21);
    if (issuer) {
      tc_fence_after();
      for (int kt = 0; kt < KT; ++kt) tma_store_4d(&maps.dk, sm.cs + kt * TILE, kt * 64, tok0, h, b);
      tma_store_commit();
      if (!last) {
        tma_store_wait_read<0>();        // dk has left the Cs tile
        load_cs(c + 1);
        TLG(26);
        // input products of the next chunk: X0 (dK) and X1 (dV) were consumed by the epilogues above
        mbar_wait(&sm.bar_dh, ph ^ 1); mbar_wait(&sm.bar_v, ph ^ 1);
        tc_fence_after();
        TLG(27);
        issue_in(0);
        mbar_wait(&sm.bar_q, ph ^ 1); mbar_wait(&sm.bar_k, ph ^ 1);
        tc_fence_after();
        TLG(28);
        issue_in(1);
      }
    }
    if (compute) named_sync(7, GT0);   // with the gate warp: gates two steps ahead are complete; it may now scan this step's rows
  }
  if (gatew) scan_of(NC - 1);
  if (issuer) tma_store_wait_read<0>();   // the staged tiles have been read; the global writes complete on their own
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

}  // namespace

#ifdef MLSTM_TIMELINE
size_t tc_bwd_fused128_workspace(const mlstm_params&) { return 16384; }
#else
size_t tc_bwd_fused128_workspace(const mlstm_params&) { return 0; }
#endif

int tc_bwd_fused128(const mlstm_params& p, cudaStream_t st, int part) {
  if (part == 0) return MLSTM_OK;   // one kernel: everything runs as "part 1"
  const StateLayout slay(p.B, p.NH, p.S, DH);
  if (!p.states || p.states_bytes < slay.total) {
    set_error("backward needs the forward's chunk-state buffer (%zu bytes)", slay.total);
    return MLSTM_ERR_WORKSPACE;
  }
  F128Maps m;
  int r = 0;
  r |= make_act_tmap(&m.q, p.q.ptr, p.B, p.NH, p.S, DH, p.q.stride_b, p.q.stride_h, p.q.stride_s, LB);
  r |= make_act_tmap(&m.k, p.k.ptr, p.B, p.NH, p.S, DH, p.k.stride_b, p.k.stride_h, p.k.stride_s, LB);
  r |= make_act_tmap(&m.v, p.v.ptr, p.B, p.NH, p.S, DH, p.v.stride_b, p.v.stride_h, p.v.stride_s, LB);
  r |= make_act_tmap(&m.dh, p.dh.ptr, p.B, p.NH, p.S, DH, p.dh.stride_b, p.dh.stride_h, p.dh.stride_s, LB);
  r |= make_act_tmap(&m.dq, p.dq.ptr, p.B, p.NH, p.S, DH, p.dq.stride_b, p.dq.stride_h, p.dq.stride_s, L);
  r |= make_act_tmap(&m.dk, p.dk.ptr, p.B, p.NH, p.S, DH, p.dk.stride_b, p.dk.stride_h, p.dk.stride_s, L);
  r |= make_act_tmap(&m.dv, p.dv.ptr, p.B, p.NH, p.S, DH, p.dv.stride_b, p.dv.stride_h, p.dv.stride_s, L);
  const size_t n_items = (size_t)p.B * p.NH * num_chunks(p.S);
  r |= make_state_tmap(&m.cs, reinterpret_cast<uint8_t*>(p.states) + slay.cs_off, n_items * DH, DH);
  if (r) {
    set_error("cuTensorMapEncodeTiled failed (%d)", r);
    return r == -1 ? MLSTM_ERR_NO_DEVICE : MLSTM_ERR_INVALID_ARG;
  }
  const size_t smem = sizeof(SmemF128);
  const int dk_true = true_extent(p.k.ptr, 128), dv_true = true_extent(p.h.ptr, 128);
  const bool narrow = dk_true < 128 || dv_true < 128;
  cudaError_t e = narrow ? set_max_smem_once(reinterpret_cast<const void*>(tc_bwd_fused128_kernel<true>), smem)
                         : set_max_smem_once(reinterpret_cast<const void*>(tc_bwd_fused128_kernel<false>), smem);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(tc_bwd_fused128, %zu B): %s", smem, cudaGetErrorString(e));
    return MLSTM_ERR_CUDA;
  }
  if (narrow) tc_bwd_fused128_kernel<true><<<dim3(p.B * p.NH), dim3(NT), smem, st>>>(m, p, resolve_scale(p), dk_true, dv_true);
  else tc_bwd_fused128_kernel<false><<<dim3(p.B * p.NH), dim3(NT), smem, st>>>(m, p, resolve_scale(p), dk_true, dv_true);
  count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("tc_bwd_fused128 launch failed: %s", cudaGetErrorString(e));
    return MLSTM_ERR_CUDA;
  }
  return MLSTM_OK;
}

}  // namespace mlstm

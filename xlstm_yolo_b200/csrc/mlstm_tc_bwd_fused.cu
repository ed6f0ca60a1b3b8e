// Fused single-walk backward for short sequences at DH = 64 (the BASELINE cfg2 shape: B=32, NH=4, S=400).
//
// One CTA per (batch, head) walks the chunks once, in reverse scan order, and produces dq, dk, dv, di, df
// together: q, k, v, dh, h are read exactly once.  What makes the single walk possible is the forward's
// per-chunk entry states (Cs, ns in `states`): with them the inter-chunk term of dq no longer needs a
// forward-order recomputation of C, so the old two walks (dq in scan order, dk/dv in reverse) collapse into one.
//
// Per chunk (rows = queries t for every gated tile, so one thread owns one row throughout):
//   MMA group 1   Z = dH V^T            S = Q K^T             G = dH Cs^T
//   SIMT          dn_t = dnf_t (dh_t.h_t);  dS = (Z/N + dn) * D  and  E = s S * D / N  -> two bf16 tiles in smem
//   MMA group 2   dQ = dS K             dK = dS^T Q           dV = E^T dH            (dS^T / E^T: the same tiles
//                 Ik = V dCb^T          Iv = K dCb                                    read as MN-major A operands)
//   epilogues     dq = s (dQ + w (G/N + dn n_prev))   dk = s dK + kw (Ik + dn_state)   dv = dV + kw Iv
//                 R = q.dq, K = k.dk -> di, df (suffix sums fall out of the reverse walk, carry in smem)
//   MMA group 3   dC += Qtilde^T dH     (Qtilde = (w s / N) Q, rows rescaled in place once q is no longer needed)
//   state pass    dCb <- bf16(dC), dC <- decay dC ;  dn_state likewise (SIMT column sums of Q)
// TMEM (512): Z|dQ 128, S|dK,dV 128, G 64, Ik 64, Iv 64, dC 64.   Shared memory 212 KB: q, dh double buffered.
#include "tc_common.cuh"

namespace mlstm {
namespace {

using namespace tc;

constexpr int DH = 64;
constexpr int NB = DH / 32;   // 32-column blocks of a DH-wide accumulator

// Developer aid (-DMLSTM_TIMELINE): CTA 0 stamps clock64() per phase into the workspace (tests/gpu_tools/timeline_fused.py)
#ifdef MLSTM_TIMELINE
#define TLF(k) do { if (blockIdx.x == 0 && c < 8) { \
    const int who_ = threadIdx.x == 0 ? 0 : threadIdx.x == 96 ? 1 : threadIdx.x == 480 ? 2 : threadIdx.x == CT ? 3 : -1; \
    if (who_ >= 0) tlf[c * 96 + who_ * 24 + (k)] = clock64(); } } while (0)
#else
#define TLF(k) do { } while (0)
#endif

struct FMaps { CUtensorMap q, k, v, dh, h, cs, dq, dk, dv; };

// 32 columns [32 cb, 32 cb + 32) of row `row` of one swizzled [128][64] bf16 tile -> fp32
__device__ __forceinline__ void tile_row32_64(const uint8_t* tile, int row, int cb, float (&out)[32]) {
#pragma unroll
  for (int x = 0; x < 32; x += 8) {
    const uint4 w = *reinterpret_cast<const uint4*>(tile + swz128(row, cb * 32 + x));
    const __nv_bfloat162* qq = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f2 = __bfloat1622float2(qq[e]);
      out[x + 2 * e] = f2.x;
      out[x + 2 * e + 1] = f2.y;
    }
  }
}

struct SmemF {
  alignas(1024) uint8_t q[2][TILE];
  alignas(1024) uint8_t dh[2][TILE];
  alignas(1024) uint8_t k[2][TILE];
  alignas(1024) uint8_t v[TILE];
  alignas(1024) uint8_t xs[2 * TILE];        // dS  [t][j] (two 64-column tiles); then staging of dq | dk
  alignas(1024) uint8_t xe[2 * TILE];        // E^T [j][t]; then tile 0 = staging of dv, tile 1 = the next chunk's h tile
  alignas(1024) uint8_t cs[DH * 128];        // forward entry state of the chunk, bf16 [dk][dv]
  alignas(1024) uint8_t dcb[DH * 128];       // bf16 copy of the adjoint state leaving the chunk, [dk][dv]
  GateBuf g[3];
  alignas(16) float ns[3][DH];               // n_prev of the chunk (ring with g)
  alignas(16) float nvec[DH];                // dn_state leaving the chunk
  alignas(16) float ncoef[L];                // (w s dn)_t
  alignas(16) float rowscale[L];             // (w s / N)_t
  float npart[8][DH];
  float part[2][L];                          // dn partials
  float partR[2][L], partK[2][L];
  float scan[8];
  float df_carry;
  uint64_t bar_q[2], bar_dh[2], bar_k[2], bar_v, bar_h, bar_cs, bar_m1, bar_i, bar_m2, bar_m3;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(NT, 1) tc_bwd_fused_kernel(const __grid_constant__ FMaps maps, const mlstm_params p,
                                                             const float scale) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemF& sm = *reinterpret_cast<SmemF*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();
#ifdef MLSTM_TIMELINE
  long long* tlf = reinterpret_cast<long long*>(p.workspace);
  const int who_h = threadIdx.x == 0 ? 0 : threadIdx.x == 96 ? 1 : threadIdx.x == 480 ? 2 : threadIdx.x == 512 ? 3 : -1;
#define TLH(k) do { if (blockIdx.x == 0 && who_h >= 0) tlf[800 + who_h * 8 + (k)] = clock64(); } while (0)
#else
#define TLH(k) do { } while (0)
#endif
  TLH(0);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool compute = tid < CT, issuer = tid == CT, gatew = tid >= GT0;
  const int rg = warp & 3, cq = compute ? (warp >> 2) : 4, row = rg * 32 + lane;
  const int bh = blockIdx.x, b = bh / p.NH, h = bh % p.NH;
  const int S = p.S, NC = num_chunks(S);
  const bool rev = p.reverse != 0;
  const StateLayout slay(p.B, p.NH, S, DH);
  const float* ns_all = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(p.states) + slay.ns_off) + (size_t)bh * NC * DH;
  const float l2s = log2f(scale);

  if (issuer) {
    tma_prefetch_desc(&maps.q); tma_prefetch_desc(&maps.k); tma_prefetch_desc(&maps.v); tma_prefetch_desc(&maps.dh);
    tma_prefetch_desc(&maps.h); tma_prefetch_desc(&maps.cs);
    mbar_init(&sm.bar_q[0], 1); mbar_init(&sm.bar_q[1], 1); mbar_init(&sm.bar_dh[0], 1); mbar_init(&sm.bar_dh[1], 1);
    mbar_init(&sm.bar_k[0], 1); mbar_init(&sm.bar_k[1], 1); mbar_init(&sm.bar_v, 1); mbar_init(&sm.bar_h, 1); mbar_init(&sm.bar_cs, 1);
    mbar_init(&sm.bar_m1, 3); mbar_init(&sm.bar_i, 2); mbar_init(&sm.bar_m2, 3); mbar_init(&sm.bar_m3, 1);
    fence_mbar_init();
    sm.df_carry = 0.f;
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, 512);
  for (int e = tid; e < DH * 128 / 16; e += NT) reinterpret_cast<uint4*>(sm.dcb)[e] = make_uint4(0, 0, 0, 0);
  for (int e = tid; e < DH; e += NT) sm.nvec[e] = 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  TLH(1);
  const uint32_t tm = sm.tmem_base;
  const uint32_t tZ = tm, tS = tm + 128, tG = tm + 256, tIk = tm + 320, tIv = tm + 384, tdC = tm + 448;
  const uint32_t lane_sel = (uint32_t)(rg * 32) << 16;

  // processing step c handles scan chunk sc = NC-1-c, memory chunk mem_chunk(sc)
  auto sc_of = [&](int c) { return NC - 1 - c; };
  auto tok0_of = [&](int c) { return mem_chunk(sc_of(c), NC, rev) * L; };
  auto load1 = [&](uint8_t* dst, const CUtensorMap* map, uint64_t* bar, int c) {
    mbar_arrive_expect_tx(bar, TILE);
    tma_load_4d(dst, map, bar, 0, tok0_of(c), h, b);
  };
  auto load_cs = [&](int c) {
    mbar_arrive_expect_tx(&sm.bar_cs, DH * 128);
    tma_load_2d(sm.cs, &maps.cs, &sm.bar_cs, 0, (bh * NC + sc_of(c)) * DH);
  };
  auto gates_of = [&](int c) {   // gate warp
    const int slot = c % 3;
    gates_warp_bwd(sm.g[slot], p, b, h, bh, mem_chunk(sc_of(c), NC, rev), lane, nullptr);
    for (int d = lane; d < DH; d += 32) sm.ns[slot][d] = ns_all[(size_t)sc_of(c) * DH + d];
    __syncwarp();
  };

  // operand descriptors (all tiles are [rows][64 bf16], 128-byte swizzled)
  // double-buffered tiles: descriptor of buffer 0 plus a constant per buffer (a descriptor array indexed by `buf` would live in
  // local memory and cost four register-to-uniform moves in front of every MMA)
  constexpr uint64_t BUF_STEP = (uint64_t)TILE >> 4;
  const uint64_t dQk0 = make_sdesc(smem_u32(sm.q[0]), 16, 1024), dQmnB0 = make_sdesc(smem_u32(sm.q[0]), TILE, 1024);
  const uint64_t dQmnA0 = make_sdesc(smem_u32(sm.q[0]), 0, 1024);   // M = dk = 64: 2nd M block aliases the 1st
  const uint64_t dHk0 = make_sdesc(smem_u32(sm.dh[0]), 16, 1024), dHmn0 = make_sdesc(smem_u32(sm.dh[0]), TILE, 1024);
  const uint64_t dKk0 = make_sdesc(smem_u32(sm.k[0]), 16, 1024), dKmn0 = make_sdesc(smem_u32(sm.k[0]), TILE, 1024);
  const uint64_t dVk = make_sdesc(smem_u32(sm.v), 16, 1024);
  const uint64_t dXSk = make_sdesc(smem_u32(sm.xs), 16, 1024), dXSmn = make_sdesc(smem_u32(sm.xs), TILE, 1024);
  const uint64_t dXEk = make_sdesc(smem_u32(sm.xe), 16, 1024);
  uint8_t* const sh = sm.xe + TILE;   // h tile of the chunk: parked in the second E^T tile between MMA group 2 and the next P2
  const uint64_t dCsk = make_sdesc(smem_u32(sm.cs), 16, 1024);
  const uint64_t dCbk = make_sdesc(smem_u32(sm.dcb), 16, 1024), dCbmn = make_sdesc(smem_u32(sm.dcb), DH * 128, 1024);

  // MMA group 1 of a chunk, one product per call (each arrives once on bar_m1): 0: Z = dH V^T (rows t), 1: S^T = K Q^T (rows j),
  // 2: G = dH Cs^T.  ptxas wraps every tcgen05.mma whose descriptors depend on the step in a register-to-uniform waterfall
  // (~100 cycles per MMA from one lane, the tensor pipe needs ~51), so independent products are issued by different lanes.
  auto issue_g1 = [&](int which, int buf) {
    constexpr uint32_t id128 = make_idesc_bf16(128, 128, 0, 0), id64 = make_idesc_bf16(128, DH, 0, 0);
    if (which == 0) {
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) umma_bf16_ss(tZ, (dHk0 + buf * BUF_STEP) + kstep(ks), dVk + kstep(ks), id128, ks > 0);
    } else if (which == 1) {
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) umma_bf16_ss(tS, (dKk0 + buf * BUF_STEP) + kstep(ks), (dQk0 + buf * BUF_STEP) + kstep(ks), id128, ks > 0);
    } else {
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) umma_bf16_ss(tG, (dHk0 + buf * BUF_STEP) + kstep(ks), dCsk + kstep(ks), id64, ks > 0);
    }
    umma_commit(&sm.bar_m1);
  };

  // products with the adjoint state leaving the chunk (0: Ik = V dCb^T, 1: Iv = K dCb; each arrives once on bar_i): needed by
  // the epilogues only, issued as soon as the state pass of the previous step has refreshed dCb — V is then dead early and
  // the next V tile has a whole step to land
  auto issue_g1b = [&](int which, int buf) {
    constexpr uint32_t idKK = make_idesc_bf16(128, DH, 0, 0), idKmn_ = make_idesc_bf16(128, DH, 0, 1);
    if (which == 0) {
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) umma_bf16_ss(tIk, dVk + kstep(ks), dCbk + kstep(ks), idKK, ks > 0);
    } else {
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks) umma_bf16_ss(tIv, (dKk0 + buf * BUF_STEP) + kstep(ks), dCbmn + mnstep(ks), idKmn_, ks > 0);
    }
    umma_commit(&sm.bar_i);
  };

  if (issuer) {
    load1(sm.dh[0], &maps.dh, &sm.bar_dh[0], 0); load1(sm.v, &maps.v, &sm.bar_v, 0);
    load1(sm.q[0], &maps.q, &sm.bar_q[0], 0); load1(sm.k[0], &maps.k, &sm.bar_k[0], 0);
    load_cs(0); load1(sh, &maps.h, &sm.bar_h, 0);
    if (NC > 1) { load1(sm.dh[1], &maps.dh, &sm.bar_dh[1], 1); load1(sm.q[1], &maps.q, &sm.bar_q[1], 1); load1(sm.k[1], &maps.k, &sm.bar_k[1], 1); }
  }
  if (gatew) gates_of(0);
  if (warp == 1 && NC > 1) gates_of(1);   // the compute warps are idle here: the first two chunks' gates side by side
  __syncthreads();
  TLH(2);
  if (issuer) {
    mbar_wait(&sm.bar_dh[0], 0); mbar_wait(&sm.bar_v, 0); mbar_wait(&sm.bar_q[0], 0); mbar_wait(&sm.bar_k[0], 0);
    mbar_wait(&sm.bar_cs, 0);
    tc_fence_after();
    issue_g1(0, 0); issue_g1(1, 0); issue_g1(2, 0);
    issue_g1b(0, 0); issue_g1b(1, 0);   // dCb = 0: zero products
  }

  // dn pass of processing step c (compute warps): dn_t = dnf_t (dh_t . h_t), plus the row coefficients the state update and the
  // dn_state sums use.  Runs one step ahead: step 0 here, step c + 1 in the shadow of step c's state MMA.
  auto dn_pass = [&](int c) -> float {
    const int bf = c & 1;
    const GateBuf& Gc = sm.g[c % 3];
    mbar_wait(&sm.bar_h, c & 1);
    mbar_wait(&sm.bar_dh[bf], (c >> 1) & 1);
    if (cq < NB) {
      float part = 0.f;
#pragma unroll
      for (int x8 = 0; x8 < 32; x8 += 8) {
        const uint32_t off = swz128(row, cq * 32 + x8);
        const uint4 wh = *reinterpret_cast<const uint4*>(sh + off);
        const uint4 wd = *reinterpret_cast<const uint4*>(sm.dh[bf] + off);
        const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&wh);
        const __nv_bfloat162* dd = reinterpret_cast<const __nv_bfloat162*>(&wd);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = __bfloat1622float2(hh[e]), c2 = __bfloat1622float2(dd[e]);
          part = fmaf(a.x, c2.x, fmaf(a.y, c2.y, part));
        }
      }
      sm.part[cq][row] = part;
    }
    named_sync(3, CT);
    const float dn = Gc.dnf[row] * (sm.part[0][row] + sm.part[1][row]);
    if (cq == 0) {
      sm.ncoef[row] = Gc.w[row] * scale * dn;
      sm.rowscale[row] = Gc.w[row] * scale * Gc.invN[row];
    }
    return dn;
  };
  float dn_next = compute ? dn_pass(0) : 0.f;
  float nstate = 0.f;   // thread dk < DH: decayed dn_state entering the step
  TLH(3);
  for (int c = 0; c < NC; ++c) {
    const uint32_t ph = c & 1;
    const int buf = c & 1;
    const bool last = (c + 1 == NC);
    if (gatew) {
      if (c + 2 < NC) gates_of(c + 2);
      named_sync(7, GT0);   // with the compute warps: gates two steps ahead are complete
      continue;
    }
    const GateBuf& G = sm.g[c % 3];
    const GateBuf& Gn = sm.g[(c + 1) % 3];
    const float* nsv = sm.ns[c % 3];
    const int tok0 = tok0_of(c);
    const int tok = tok0 + row;
    const bool row_ok = compute && tok < S;

    // ---- P1 ran one step ahead (in the shadow of the previous step's state MMA, or in the prologue)
    TLF(0);
    const float dn_row = dn_next;
    TLF(1);
    mbar_wait(&sm.bar_m1, ph);
    tc_fence_after();
    TLF(2);
    if (issuer) {
      mbar_wait(&sm.bar_i, ph);             // V dCb^T, K dCb complete: with group 1 done, V and Cs are dead
      if (!last) { load1(sm.v, &maps.v, &sm.bar_v, c + 1); load_cs(c + 1); }
    }
    TLF(3);

    // ---- P2: gated tiles.  dS -> xs with rows = queries t (causal work grows with the row group), E^T -> xe with
    //      rows = keys j (causal work shrinks with the row group): together every scheduler gets five 32x32 blocks
    if (compute) {
      const bool fullA = rev ? (cq > rg) : (cq < rg), fullB = rev ? (cq < rg) : (cq > rg);
      const bool diag = (cq == rg);
      const uint32_t bitsA = causal_bits(fullA, !rev, lane), bitsB = causal_bits(fullB, rev, lane);
      uint32_t pk[16], pe[16];
      if (fullA || diag) {   // dS[t][j] = (Z invN_t + dn_t) 2^(u2_j - M2_t), keep j <= t (reverse: j >= t)
        float z[32];
        tmem_ld32(tZ + lane_sel + cq * 32, z);
        tmem_ld_wait();
        TLF(16);
        const float M2t = G.M2[row], invN = G.invN[row];
#pragma unroll
        for (int x = 0; x < 32; x += 4) {
          const float4 u4 = *reinterpret_cast<const float4*>(&G.u2[cq * 32 + x]);
          const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
          float ds[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool keep = (bitsA >> (x + e)) & 1u;
            ds[e] = keep ? fmaf(z[x + e], invN, dn_row) * ex2(uu[e] - M2t) : 0.f;
          }
          pk[x / 2] = pack_bf16x2(ds[0], ds[1]); pk[x / 2 + 1] = pack_bf16x2(ds[2], ds[3]);
        }
      } else {
#pragma unroll
        for (int x = 0; x < 16; ++x) pk[x] = 0u;
      }
      TLF(17);
      if (fullB || diag) {   // E^T[j][t] = S^T 2^(u2_j + log2 s - c2_t), keep t >= j (reverse: t <= j)
        float s_[32];
        tmem_ld32(tS + lane_sel + cq * 32, s_);
        tmem_ld_wait();
        TLF(18);
        const float u2j = G.u2[row] + l2s;
#pragma unroll
        for (int x = 0; x < 32; x += 4) {
          const float4 c4 = *reinterpret_cast<const float4*>(&G.c2[cq * 32 + x]);
          const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
          float ev[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool keep = (bitsB >> (x + e)) & 1u;
            ev[e] = keep ? s_[x + e] * ex2(u2j - cc[e]) : 0.f;
          }
          pe[x / 2] = pack_bf16x2(ev[0], ev[1]); pe[x / 2 + 1] = pack_bf16x2(ev[2], ev[3]);
        }
      } else {
#pragma unroll
        for (int x = 0; x < 16; ++x) pe[x] = 0u;
      }
      TLF(14);
      // xs / xe still feed the previous step's output stores: the control warp drains them before it joins this
      // barrier, and that wait hides behind the tile math above (the packed tiles sit in registers meanwhile)
      named_sync(5, GT0);
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int col = cq * 32 + x * 8;
        const uint32_t off = (col >> 6) * TILE + swz128(row, col & 63);
        *reinterpret_cast<uint4*>(sm.xs + off) = make_uint4(pk[4 * x], pk[4 * x + 1], pk[4 * x + 2], pk[4 * x + 3]);
        *reinterpret_cast<uint4*>(sm.xe + off) = make_uint4(pe[4 * x], pe[4 * x + 1], pe[4 * x + 2], pe[4 * x + 3]);
      }
    }
    if (!compute) named_sync(5, GT0);   // control warp: its store drain (end of the previous step) is complete
    TLF(4);
    fence_proxy_async_smem();
    tc_fence_before();
    named_sync(2, GT0);
    TLF(5);

    // ---- MMA group 2: three independent products, one issuing lane each.  ptxas wraps every tcgen05.mma whose descriptors
    //      depend on the step in a register-to-uniform waterfall (~100 cycles per MMA from one lane; the tensor pipe needs
    //      ~51), and every compute warp is idle here: lanes 0 of compute warps 1 and 2 issue dK and dV next to the control lane
    const int g2 = issuer ? 0 : (tid == 32 ? 1 : (tid == 64 ? 2 : -1));
    if (g2 >= 0) {
      tc_fence_after();
      constexpr uint32_t idKmn = make_idesc_bf16(128, DH, 0, 1);   // A K-major, B MN-major
      constexpr uint32_t idMM = make_idesc_bf16(128, DH, 1, 1);    // A MN-major, B MN-major
      if (g2 == 0) {
#pragma unroll
        for (int ks = 0; ks < L / 16; ++ks) umma_bf16_ss(tZ, dXSk + kstep(ks), (dKmn0 + buf * BUF_STEP) + mnstep(ks), idKmn, ks > 0);         // dQ = dS K
      } else if (g2 == 1) {
#pragma unroll
        for (int ks = 0; ks < L / 16; ++ks) umma_bf16_ss(tS, dXSmn + mnstep(ks), (dQmnB0 + buf * BUF_STEP) + mnstep(ks), idMM, ks > 0);     // dK = dS^T Q
      } else {
#pragma unroll
        for (int ks = 0; ks < L / 16; ++ks) umma_bf16_ss(tS + 64, dXEk + kstep(ks), (dHmn0 + buf * BUF_STEP) + mnstep(ks), idKmn, ks > 0);   // dV = E^T dH
      }
      umma_commit(&sm.bar_m2);   // three arrivals complete the phase
    }
    TLF(6);
    // in the shadow of group 2: dn_state contribution = column sums of the (un-scaled) Q tile, 16 rows per thread
    if (compute) {
      {
        const int dk = tid & 63, pt = tid >> 6;
        float acc = 0.f;
#pragma unroll 4
        for (int t = pt * 16; t < pt * 16 + 16; ++t)
          acc = fmaf(sm.ncoef[t], __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(sm.q[buf] + swz128(t, dk))), acc);
        sm.npart[pt][dk] = acc;
      }
    }
    TLF(15);
    mbar_wait(&sm.bar_m2, ph);
    tc_fence_after();
    TLF(7);
    if (issuer && !last) load1(sh, &maps.h, &sm.bar_h, c + 1);   // second E^T tile is dead (dv is staged in the first): park the next h there

    // ---- P3: epilogues.  cq 0,1: dq blocks ; cq 2,3: dk then dv blocks -------------------------------
    if (compute) {
      uint32_t opk[16];
      float psum = 0.f;
      const float kwj = G.kw[row];
      if (cq < 2) {
        float acc[32], gg[32], qr[32];
        tmem_ld32(tZ + lane_sel + cq * 32, acc);
        tmem_ld32(tG + lane_sel + cq * 32, gg);
        tmem_ld_wait();
        const float wt = G.w[row], invN = G.invN[row];
        tile_row32_64(sm.q[buf], row, cq, qr);
#pragma unroll
        for (int x = 0; x < 32; x += 2) {
          const float o0 = scale * (acc[x] + wt * fmaf(gg[x], invN, dn_row * nsv[cq * 32 + x]));
          const float o1 = scale * (acc[x + 1] + wt * fmaf(gg[x + 1], invN, dn_row * nsv[cq * 32 + x + 1]));
          if (row_ok) psum = fmaf(qr[x], o0, fmaf(qr[x + 1], o1, psum));
          opk[x / 2] = pack_bf16x2(o0, o1);
        }
        sm.partR[cq][row] = psum;
      } else {
        const int cb = cq - 2;
        float acc[32], gi[32], kr[32];
        tmem_ld32(tS + lane_sel + cb * 32, acc);
        tmem_ld32(tIk + lane_sel + cb * 32, gi);
        tmem_ld_wait();
        tile_row32_64(sm.k[buf], row, cb, kr);
#pragma unroll
        for (int x = 0; x < 32; x += 2) {
          const float o0 = fmaf(kwj, gi[x] + sm.nvec[cb * 32 + x], scale * acc[x]);
          const float o1 = fmaf(kwj, gi[x + 1] + sm.nvec[cb * 32 + x + 1], scale * acc[x + 1]);
          if (row_ok) psum = fmaf(kr[x], o0, fmaf(kr[x + 1], o1, psum));
          opk[x / 2] = pack_bf16x2(o0, o1);
        }
        sm.partK[cb][row] = psum;
      }
      named_sync(3, CT);   // xs / xe (operands of group 2, complete) can now take the staged outputs; partials visible
      {
        uint8_t* stage = sm.xs + (cq < 2 ? 0 : TILE);   // dq | dk
        const int cb = cq & 1;
#pragma unroll
        for (int x4 = 0; x4 < 4; ++x4)
          *reinterpret_cast<uint4*>(stage + swz128(row, cb * 32 + x4 * 8)) =
              make_uint4(opk[4 * x4], opk[4 * x4 + 1], opk[4 * x4 + 2], opk[4 * x4 + 3]);
      }
      if (cq >= 2) {   // dv = dV + kw Iv
        const int cb = cq - 2;
        float acc[32], gi[32];
        tmem_ld32(tS + 64 + lane_sel + cb * 32, acc);
        tmem_ld32(tIv + lane_sel + cb * 32, gi);
        tmem_ld_wait();
#pragma unroll
        for (int x4 = 0; x4 < 4; ++x4) {
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int x = x4 * 8 + 2 * e;
            w[e] = pack_bf16x2(fmaf(kwj, gi[x], acc[x]), fmaf(kwj, gi[x + 1], acc[x + 1]));
          }
          *reinterpret_cast<uint4*>(sm.xe + swz128(row, cb * 32 + x4 * 8)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      } else if (cq == 0) {
        // di_j = K_j ; df_j = sigmoid(-f_j) (suffix sum in scan order of (R - K) + carry from the later chunks)
        const float Kj = sm.partK[0][row] + sm.partK[1][row];
        const float dB = row_ok ? (sm.partR[0][row] + sm.partR[1][row] - Kj) : 0.f;
        float pre = warp_scan_add(dB, lane);
        if (lane == 31) sm.scan[rg] = pre;
        named_sync(4, 128);
        float off = 0.f, tot = 0.f;
#pragma unroll
        for (int w = 0; w < 4; ++w) { off += (w < rg) ? sm.scan[w] : 0.f; tot += sm.scan[w]; }
        pre += off;
        const float carry = sm.df_carry;
        const float suf = rev ? pre : (tot - pre + dB);
        if (row_ok) {
          const float i_raw = p.i.ptr[(int64_t)b * p.i.stride_b + (int64_t)h * p.i.stride_h + (int64_t)tok * p.i.stride_s];
          p.di.ptr[(int64_t)b * p.di.stride_b + (int64_t)h * p.di.stride_h + (int64_t)tok * p.di.stride_s] = Kj * igate_dlog(p, i_raw);
          p.df.ptr[(int64_t)b * p.df.stride_b + (int64_t)h * p.df.stride_h + (int64_t)tok * p.df.stride_s] = (suf + carry) * G.sig[row];
        }
        named_sync(4, 128);
        if (tid == 0) sm.df_carry = carry + tot;
      }
      // ---- P4: Qtilde = (w s / N) Q in place (q rows were consumed above) -------------------------------
      named_sync(3, CT);
      if (!last) scale_rows<DH>(sm.q[buf], sm.rowscale, tid);   // the adjoint state leaving the first chunk is not an output
    }
    TLF(8);
    fence_proxy_async_smem();
    tc_fence_before();
    named_sync(2, GT0);
    TLF(9);
    if (issuer) {
      tc_fence_after();
      if (!last) {
        constexpr uint32_t idMM = make_idesc_bf16(128, DH, 1, 1);
#pragma unroll
        for (int ks = 0; ks < L / 16; ++ks) umma_bf16_ss(tdC, (dQmnA0 + buf * BUF_STEP) + mnstep(ks), (dHmn0 + buf * BUF_STEP) + mnstep(ks), idMM, (ks > 0 || c > 0) ? 1u : 0u);
        umma_commit(&sm.bar_m3);
      }
      // the staged outputs are final: store them now, so the drain is over long before the next gated tiles overwrite xs / xe
      tma_store_4d(&maps.dq, sm.xs, 0, tok0, h, b);
      tma_store_4d(&maps.dk, sm.xs + TILE, 0, tok0, h, b);
      tma_store_4d(&maps.dv, sm.xe, 0, tok0, h, b);
      tma_store_commit();
      if (c + 2 < NC) load1(sm.k[buf], &maps.k, &sm.bar_k[buf], c + 2);   // k rows were consumed in the epilogue
    }
    if (compute && !last) dn_next = dn_pass(c + 1);   // in the shadow of the state MMA
    // group 1 of the next chunk (Z, S^T, G were consumed by the epilogues above), one product per lane: lanes 0 of compute warps
    // 2, 3 and 6, which have no rows in the state pass below — their issue time hides behind it
    const int g1w = (warp == 2) ? 0 : (warp == 3) ? 1 : (warp == 6) ? 2 : -1;
    if (!last && lane == 0 && g1w >= 0) {
      const int nb = buf ^ 1;
      const uint32_t pn = ((c + 1) >> 1) & 1;
      if (g1w == 0) { mbar_wait(&sm.bar_dh[nb], pn); mbar_wait(&sm.bar_v, ph ^ 1); }
      else if (g1w == 1) { mbar_wait(&sm.bar_k[nb], pn); mbar_wait(&sm.bar_q[nb], pn); }
      else { mbar_wait(&sm.bar_dh[nb], pn); mbar_wait(&sm.bar_cs, ph ^ 1); }
      tc_fence_after();
      issue_g1(g1w, nb);
    }
    if (!last) {
      mbar_wait(&sm.bar_m3, ph);
      tc_fence_after();
    }
    TLF(10);
    if (issuer && c + 2 < NC) { load1(sm.dh[buf], &maps.dh, &sm.bar_dh[buf], c + 2); load1(sm.q[buf], &maps.q, &sm.bar_q[buf], c + 2); }

    // ---- P5: state pass: dCb <- bf16(dC), dC <- decay_next dC ; dn_state likewise ----------------------
    const float dnext = last ? 1.f : Gn.decay;
    if (compute && row < DH && cq < NB && !last) {
      float r[32];
      tmem_ld32(tdC + lane_sel + cq * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int x = 0; x < 32; x += 8)
        *reinterpret_cast<uint4*>(sm.dcb + swz128(row, cq * 32 + x)) =
            make_uint4(pack_bf16x2(r[x], r[x + 1]), pack_bf16x2(r[x + 2], r[x + 3]), pack_bf16x2(r[x + 4], r[x + 5]),
                       pack_bf16x2(r[x + 6], r[x + 7]));
#pragma unroll
      for (int x = 0; x < 32; ++x) r[x] *= dnext;
      tmem_st32(tdC + lane_sel + cq * 32, r);
      tmem_st_wait();
    }
    if (compute && tid < DH && !last) {
      float nv = nstate;
#pragma unroll
      for (int pt = 0; pt < 8; ++pt) nv += sm.npart[pt][tid];
      sm.nvec[tid] = nv;
      nstate = nv * dnext;
    }
    TLF(11);
    fence_proxy_async_smem();
    tc_fence_before();
    // end of step: the compute warps meet the gate warp (gates two steps ahead are complete).  The control warp takes no part:
    // it stored the outputs behind the state MMA and the products with dCb are issued by compute lanes, so it runs ahead to
    // the next step's loads and meets the compute warps again at barrier 5.
    if (compute) named_sync(7, GT0);
    TLF(12);
    if (!last && (tid == 32 || tid == 64)) {   // V, K of the next chunk landed before group 1 of that chunk was issued
      tc_fence_after();
      issue_g1b(tid == 32 ? 0 : 1, buf ^ 1);
    }
    if (issuer) tma_store_wait_read<0>();   // before this warp joins barrier 5 of the next step (xs / xe are rewritten after it)
    TLF(13);
  }
  TLH(4);
  if (issuer) tma_store_wait_read<0>();   // the staged tiles have been read; the global writes complete on their own
  tc_fence_before();
  __syncthreads();
  TLH(5);
  if (warp == 0) tmem_dealloc(tm, 512);
  TLH(6);
}

}  // namespace

bool tc_use_fused_bwd(const mlstm_params& p);

#ifdef MLSTM_TIMELINE
size_t tc_bwd_fused_workspace(const mlstm_params&) { return 16384; }
#else
size_t tc_bwd_fused_workspace(const mlstm_params&) { return 0; }
#endif

int tc_bwd_fused(const mlstm_params& p, cudaStream_t st, int part) {
  if (part == 0) return MLSTM_OK;   // one kernel: everything runs as "part 1"
  const StateLayout slay(p.B, p.NH, p.S, DH);
  if (!p.states || p.states_bytes < slay.total) {
    set_error("backward needs the forward's chunk-state buffer (%zu bytes)", slay.total);
    return MLSTM_ERR_WORKSPACE;
  }
  FMaps m;
  int r = 0;
  r |= make_act_tmap(&m.q, p.q.ptr, p.B, p.NH, p.S, DH, p.q.stride_b, p.q.stride_h, p.q.stride_s, L);
  r |= make_act_tmap(&m.k, p.k.ptr, p.B, p.NH, p.S, DH, p.k.stride_b, p.k.stride_h, p.k.stride_s, L);
  r |= make_act_tmap(&m.v, p.v.ptr, p.B, p.NH, p.S, DH, p.v.stride_b, p.v.stride_h, p.v.stride_s, L);
  r |= make_act_tmap(&m.dh, p.dh.ptr, p.B, p.NH, p.S, DH, p.dh.stride_b, p.dh.stride_h, p.dh.stride_s, L);
  r |= make_act_tmap(&m.h, p.h.ptr, p.B, p.NH, p.S, DH, p.h.stride_b, p.h.stride_h, p.h.stride_s, L);
  r |= make_act_tmap(&m.dq, p.dq.ptr, p.B, p.NH, p.S, DH, p.dq.stride_b, p.dq.stride_h, p.dq.stride_s, L);
  r |= make_act_tmap(&m.dk, p.dk.ptr, p.B, p.NH, p.S, DH, p.dk.stride_b, p.dk.stride_h, p.dk.stride_s, L);
  r |= make_act_tmap(&m.dv, p.dv.ptr, p.B, p.NH, p.S, DH, p.dv.stride_b, p.dv.stride_h, p.dv.stride_s, L);
  const size_t n_items = (size_t)p.B * p.NH * num_chunks(p.S);
  r |= make_state_tmap(&m.cs, reinterpret_cast<uint8_t*>(p.states) + slay.cs_off, n_items * DH, DH);
  if (r) {
    set_error("cuTensorMapEncodeTiled failed (%d)", r);
    return r == -1 ? MLSTM_ERR_NO_DEVICE : MLSTM_ERR_INVALID_ARG;
  }
  const size_t smem = sizeof(SmemF);
  cudaError_t e = set_max_smem_once(reinterpret_cast<const void*>(tc_bwd_fused_kernel), smem);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(tc_bwd_fused, %zu B): %s", smem, cudaGetErrorString(e));
    return MLSTM_ERR_CUDA;
  }
  tc_bwd_fused_kernel<<<dim3(p.B * p.NH), dim3(NT), smem, st>>>(m, p, resolve_scale(p));
  count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("tc_bwd_fused launch failed: %s", cudaGetErrorString(e));
    return MLSTM_ERR_CUDA;
  }
  return MLSTM_OK;
}

}  // namespace mlstm

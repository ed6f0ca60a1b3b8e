// Two-phase tcgen05 / TMEM / TMA forward of the mLSTM cell (bf16 I/O, DH in {64, 128}), used when
// B*NH is too small for the single-pass kernel (mlstm_tc_fwd.cu) to fill the GPU: two kernels.
//
//   tc_state_fwd   one CTA per (batch, head) walks the chunks IN ORDER and only carries the
//                  inter-chunk state (reference: backends.py:196-218): per 128-token chunk one
//                  in-place Kbar = kw*K, one accumulating MMA  C += Kbar^T V (+ n += Kbar^T 1),
//                  one TMEM pass that applies the next decay and writes the chunk's entry state
//                  (bf16 C, fp32 n, m) to the state workspace.  Light and latency-bound by design.
//   tc_fwd_par     persistent CTAs over all (batch, head, chunk) items, no serial dependency:
//                  S = Q K^T, G = Q Cs, qn = Q n   (MMA1, entry state Cs/n/m from the workspace)
//                  P = S * exp2(u - M) causal, bf16, row sums  (SIMT, one 32x32 block per warp)
//                  H = P V                                      (MMA2, into S's TMEM columns)
//                  h = (H + w G) / (max(|n|, e^-m) + eps)       (backends.py:220-263)
//
// 18 warps per CTA: 16 compute warps (warp w: tile rows / TMEM lanes 32*(w%4)..+31, column block
// w/4), one control warp whose lane 0 issues every TMA and tcgen05.mma from pre-built
// descriptors, one gate warp that prepares the gate vectors of the next item.  No D matrix, gate
// vector or P ever reaches HBM; the only extra traffic over the fused minimum is the chunk-entry
// state (DH^2/128 bf16 per token-head) and a second read of K,V by the state kernel.
#include "tc_common.cuh"

namespace mlstm {
bool tc_use_two_phase(const mlstm_params& p);
namespace {

using namespace tc;

struct FwdMaps { CUtensorMap q, k, v, cs; };

// =============================================================================================
// State kernel
// =============================================================================================
template <int DH>
struct SmemS {
  static constexpr int KT = DH / 64;
  alignas(1024) uint8_t k[2][KT * TILE];
  alignas(1024) uint8_t v[2][KT * TILE];
  alignas(1024) uint8_t stage[2][KT * DH * 128];   // bf16 entry-state tiles staged for TMA stores (double buffered)
  alignas(1024) uint8_t ones[2048];                // bf16 1.0 (B operand of n += Kbar^T 1)
  GateBuf g[3];                                    // ring: chunk sc uses g[sc % 3]
  uint64_t bar_k[2], bar_v[2], bar_mma;
  uint32_t tmem_base;
};

// gridDim.y value slices: the state C [dk][dv] splits by dv columns into independent recurrences (same K, same
// gates), so a small batch can still put a CTA on most SMs: a sequential kernel is bound by what ONE SM can pull
// from HBM per step (K + V + Cs tiles), and a slice pulls less.  Slice 0 also carries n and m.
// Head dims above the template's DH (DHF = p.DHQK = 256 with DH = 128): the state also splits by dk rows into
// DHF / DH independent row blocks (same V, same gates): blockIdx.y = row_block * nsl + slice, the CTA owns
// C[row0 .. row0 + DH)[col0 .. col0 + DVs) and reads the K columns [row0, row0 + DH) only.
template <int DH>
__global__ void __launch_bounds__(NT, 1) tc_state_fwd_kernel(const __grid_constant__ FwdMaps maps, const mlstm_params p) {
  constexpr int KT = DH / 64;
  const int DHF = p.DHQK, nrb = DHF / DH;
  // block launches (DHF > DH) put the block index in blockIdx.x: the CTAs that share a K or V tile are neighbours in launch
  // order, run in the same round and find each other's tiles in L2 (405 -> 210 MB DRAM reads at B32 NH4 S1600 DH256)
  const bool blocks = DHF != DH;
  const int by = blocks ? blockIdx.x : blockIdx.y, ny = blocks ? gridDim.x : gridDim.y;
  const int nsl = ny / nrb, sl = by % nsl, row0 = (by / nsl) * DH;
  const int DVs = DHF / nsl, col0 = sl * DVs;   // this CTA's value columns [col0, col0 + DVs)
  const int NB = DVs / 32, KTV = DVs / 64;     // active 32-column blocks / 64-column tiles of V and of the state
  constexpr uint32_t A_LBO = (DH == 128) ? TILE : 0;   // DH=64: the 2nd 64-row M block aliases the 1st
  constexpr uint32_t TCOLS = 256;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemS<DH>& sm = *reinterpret_cast<SmemS<DH>*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool compute = tid < CT, issuer = tid == CT, gatew = tid >= GT0;
  const int rg = warp & 3, cq = compute ? (warp >> 2) : 4, row = rg * 32 + lane;
  const int bh = blocks ? blockIdx.y : blockIdx.x, b = bh / p.NH, h = bh % p.NH;
  const int S = p.S, NC = num_chunks(S);
  const bool rev = p.reverse != 0, has_init = p.c_initial != nullptr;
  const StateLayout lay(p.B, p.NH, S, DHF);
  __nv_bfloat16* Cs = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(p.states) + lay.cs_off) + (size_t)bh * NC * DHF * DHF;
  float* ns = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(p.states) + lay.ns_off) + (size_t)bh * NC * DHF;
  float* ms = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(p.states) + lay.ms_off) + (size_t)bh * NC;

  if (issuer) {
    tma_prefetch_desc(&maps.k); tma_prefetch_desc(&maps.v);
    mbar_init(&sm.bar_k[0], 1); mbar_init(&sm.bar_k[1], 1); mbar_init(&sm.bar_v[0], 1); mbar_init(&sm.bar_v[1], 1);
    mbar_init(&sm.bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, TCOLS);
  for (int e = tid; e < 2048 / 4; e += NT) reinterpret_cast<uint32_t*>(sm.ones)[e] = 0x3F803F80u;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = sm.tmem_base, tC = tm, tN = tm + DH;
  const uint32_t lane_sel = (uint32_t)(rg * 32) << 16;

  auto load_kv = [&](int sc) {
    const int buf = sc & 1, tok0 = mem_chunk(sc, NC, rev) * L;
    mbar_arrive_expect_tx(&sm.bar_k[buf], KT * TILE);
    for (int kt = 0; kt < KT; ++kt) tma_load_4d(sm.k[buf] + kt * TILE, &maps.k, &sm.bar_k[buf], row0 + kt * 64, tok0, h, b);
    mbar_arrive_expect_tx(&sm.bar_v[buf], KTV * TILE);
    for (int kt = 0; kt < KTV; ++kt) tma_load_4d(sm.v[buf] + kt * TILE, &maps.v, &sm.bar_v[buf], col0 + kt * 64, tok0, h, b);
  };
  const uint64_t dK0 = make_sdesc(smem_u32(sm.k[0]), A_LBO, 1024), dV0 = make_sdesc(smem_u32(sm.v[0]), TILE, 1024);
  const uint64_t dOnes = make_sdesc(smem_u32(sm.ones), 1024, 1024);
  constexpr uint64_t BUF_STEP = (uint64_t)(KT * TILE) >> 4;

  if (issuer) { load_kv(0); if (NC > 1) load_kv(1); }
  if (gatew) {
    gates_warp_fwd(sm.g[0], p, b, h, mem_chunk(0, NC, rev), lane, p.m_initial ? p.m_initial[bh] : 0.f);
    if (NC > 1) gates_warp_fwd(sm.g[1], p, b, h, mem_chunk(1, NC, rev), lane, sm.g[0].m_next);
  }
  __syncthreads();

  // entry state of chunk 0 -> workspace; TMEM C <- decay_0 * C_0 when an initial state is given
  if (row < DH && cq < NB) {
    float r[32];
    const float* crow = has_init ? p.c_initial + ((int64_t)bh * DHF + row0 + row) * DHF + col0 + cq * 32 : nullptr;
#pragma unroll
    for (int x = 0; x < 32; ++x) r[x] = has_init ? crow[x] : 0.f;
    uint32_t pk[16];
#pragma unroll
    for (int x = 0; x < 32; x += 2) pk[x / 2] = pack_bf16x2(r[x], r[x + 1]);
    store_row32(Cs + (size_t)(row0 + row) * DHF + col0 + cq * 32, pk);
    if (has_init) {
      const float d0 = sm.g[0].decay;
#pragma unroll
      for (int x = 0; x < 32; ++x) r[x] *= d0;
      tmem_st32(tC + lane_sel + cq * 32, r);
    }
    if (cq == 0 && sl == 0) {
      const float n0 = has_init ? p.n_initial[(int64_t)bh * DHF + row0 + row] : 0.f;
      ns[row0 + row] = n0;
      if (has_init) {
#pragma unroll
        for (int x = 0; x < 32; ++x) r[x] = n0 * sm.g[0].decay;
        tmem_st32(tN + lane_sel, r);   // tN occupies 16 columns; the next 16 are scratch
      }
    }
    if (has_init) tmem_st_wait();
  }
  if (issuer && sl == 0 && row0 == 0) ms[0] = sm.g[0].m_prev;
  // Kbar(0) = kw * K(0), in place
  if (!gatew) mbar_wait(&sm.bar_k[0], 0);
  if (compute) scale_rows<DH>(sm.k[0], sm.g[0].kw, tid);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();

  for (int sc = 0; sc < NC; ++sc) {
    const bool last = (sc + 1 == NC);
    const int buf = sc & 1;
    if (gatew) {
      if (sc + 2 < NC) gates_warp_fwd(sm.g[(sc + 2) % 3], p, b, h, mem_chunk(sc + 2, NC, rev), lane, sm.g[(sc + 1) % 3].m_next);
      __syncthreads();
      continue;
    }
    // ---- state MMA of chunk sc; Kbar(sc+1) is rewritten in its shadow -----------------------
    if (issuer) {
      mbar_wait(&sm.bar_v[buf], (sc >> 1) & 1);
      tc_fence_after();
      const uint64_t dK = dK0 + buf * BUF_STEP, dV = dV0 + buf * BUF_STEP;
      const uint32_t idC = make_idesc_bf16(128, DVs, 1, 1);
      constexpr uint32_t idN = make_idesc_bf16(128, 16, 1, 1);
      const uint32_t acc0 = (sc > 0 || has_init) ? 1u : 0u;
#pragma unroll
      for (int ks = 0; ks < L / 16; ++ks) umma_bf16_ss(tC, dK + mnstep(ks), dV + mnstep(ks), idC, (ks > 0) ? 1u : acc0);
      if (sl == 0) {
#pragma unroll
        for (int ks = 0; ks < L / 16; ++ks) umma_bf16_ss(tN, dK + mnstep(ks), dOnes, idN, (ks > 0) ? 1u : acc0);
      }
      umma_commit(&sm.bar_mma);
    }
    if (!last) {
      mbar_wait(&sm.bar_k[buf ^ 1], ((sc + 1) >> 1) & 1);
      if (compute) scale_rows<DH>(sm.k[buf ^ 1], sm.g[(sc + 1) % 3].kw, tid);
      fence_proxy_async_smem();
    }
    mbar_wait(&sm.bar_mma, sc & 1);
    tc_fence_after();
    if (issuer && sc + 2 < NC) load_kv(sc + 2);   // this chunk's K/V buffers are free again

    // ---- state pass: entry state of chunk sc+1 -> workspace; C <- decay_{sc+1} C ------------
    const float dnext = last ? 1.f : sm.g[(sc + 1) % 3].decay;
    if (row < DH && cq < NB) {
      float r[32];
      tmem_ld32(tC + lane_sel + cq * 32, r);
      tmem_ld_wait();
      if (!last) {
        // entry state of chunk sc+1: bf16 into a swizzled staging tile, one TMA tile store after the barrier
        // (per-thread 64-byte row stores to global were the slowest part of this step)
#pragma unroll
        for (int x = 0; x < 32; x += 8) {
          const int dv = cq * 32 + x;
          *reinterpret_cast<uint4*>(sm.stage[buf] + (dv >> 6) * (DH * 128) + swz128(row, dv & 63)) =
              make_uint4(pack_bf16x2(r[x], r[x + 1]), pack_bf16x2(r[x + 2], r[x + 3]), pack_bf16x2(r[x + 4], r[x + 5]),
                         pack_bf16x2(r[x + 6], r[x + 7]));
        }
#pragma unroll
        for (int x = 0; x < 32; ++x) r[x] *= dnext;
        tmem_st32(tC + lane_sel + cq * 32, r);
      } else if (p.c_last) {
        float* dst = p.c_last + ((int64_t)bh * DHF + row0 + row) * DHF + col0 + cq * 32;
#pragma unroll
        for (int x = 0; x < 32; x += 4) *reinterpret_cast<float4*>(dst + x) = make_float4(r[x], r[x + 1], r[x + 2], r[x + 3]);
      }
      if (cq == 0 && sl == 0) {
        float rn[16];
        tmem_ld16(tN + lane_sel, rn);
        tmem_ld_wait();
        if (!last) {
          ns[(size_t)(sc + 1) * DHF + row0 + row] = rn[0];
#pragma unroll
          for (int x = 0; x < 32; ++x) r[x] = rn[0] * dnext;
          tmem_st32(tN + lane_sel, r);
        } else if (p.n_last) {
          p.n_last[(int64_t)bh * DHF + row0 + row] = rn[0];
        }
      }
      if (!last) tmem_st_wait();
    }
    if (issuer && sl == 0 && row0 == 0) {
      if (!last) ms[sc + 1] = sm.g[sc % 3].m_next;
      else if (p.m_last) p.m_last[bh] = sm.g[sc % 3].m_next;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    if (issuer && !last) {
      for (int kt = 0; kt < KTV; ++kt)
        tma_store_2d(&maps.cs, sm.stage[buf] + kt * (DH * 128), col0 + kt * 64, (bh * NC + sc + 1) * DHF + row0);
      tma_store_commit();
      tma_store_wait_read<1>();   // the other staging buffer (written again in the next step) has been read
    }
  }
  if (issuer) tma_store_wait_all<0>();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, TCOLS);
}

// =============================================================================================
// Chunk-parallel kernel
// =============================================================================================
template <int DH>
struct SmemP {
  static constexpr int KT = DH / 64;
  static constexpr int TILE_C = DH * 128;          // bytes of one [DH rows][64] state tile
  alignas(1024) uint8_t q[KT * TILE];
  alignas(1024) uint8_t k[KT * TILE];
  alignas(1024) uint8_t v[KT * TILE];
  alignas(1024) uint8_t cs[KT * TILE_C];           // bf16 entry state, [dk][dv]
  alignas(1024) uint8_t p[2 * TILE];               // P (K-major, 2 tiles over j)
  alignas(1024) uint8_t nvec[3][KT * 2048];        // K-major [16][DH]: row 0 = hi(n), row 1 = lo(n); ring of 3
  GateBuf g[3];                                    // the gate warp runs two items ahead
  float part_rs[4][L];
  uint64_t bar_q, bar_k, bar_v, bar_cs, bar_m1, bar_m2;
  uint32_t tmem_base;
};

template <int DH>
__global__ void __launch_bounds__(NT, 1) tc_fwd_par_kernel(const __grid_constant__ FwdMaps maps, const mlstm_params p,
                                                           const float scale, const int n_items, const int dv_true) {
  // dv_true: columns of an h row that exist in memory (< DH for a zero-padded problem running on the caller's narrow
  // tensors, tc_tmap.cuh: ExtentOverride) — the h rows leave through plain stores here, so the clipping TMA would do is explicit
  constexpr int KT = DH / 64;
  constexpr int TILE_C = SmemP<DH>::TILE_C;
  constexpr int NB = DH / 32;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemP<DH>& sm = *reinterpret_cast<SmemP<DH>*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool compute = tid < CT, issuer = tid == CT, gatew = tid >= GT0;
  const int rg = warp & 3, cq = compute ? (warp >> 2) : 4, row = rg * 32 + lane;
  const int S = p.S, NC = num_chunks(S);
  const bool rev = p.reverse != 0;
  const StateLayout lay(p.B, p.NH, S, DH);
  const float* ns_all = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(p.states) + lay.ns_off);
  const float* ms_all = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(p.states) + lay.ms_off);
  const float l2s = log2f(scale);

  if (issuer) {
    tma_prefetch_desc(&maps.q); tma_prefetch_desc(&maps.k); tma_prefetch_desc(&maps.v); tma_prefetch_desc(&maps.cs);
    mbar_init(&sm.bar_q, 1); mbar_init(&sm.bar_k, 1); mbar_init(&sm.bar_v, 1); mbar_init(&sm.bar_cs, 1);
    mbar_init(&sm.bar_m1, 1); mbar_init(&sm.bar_m2, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, 512);
  for (int e = tid; e < 3 * KT * 2048 / 16; e += NT) reinterpret_cast<uint4*>(sm.nvec)[e] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = sm.tmem_base, tS = tm, tG = tm + 128, tQN = tm + 128 + DH;
  const uint32_t lane_sel = (uint32_t)(rg * 32) << 16;

  const uint64_t dQ = make_sdesc(smem_u32(sm.q), 16, 1024), dKk = make_sdesc(smem_u32(sm.k), 16, 1024);
  const uint64_t dCs = make_sdesc(smem_u32(sm.cs), TILE_C, 1024), dVmn = make_sdesc(smem_u32(sm.v), TILE, 1024);
  const uint64_t dP = make_sdesc(smem_u32(sm.p), 16, 1024), dNv0 = make_sdesc(smem_u32(sm.nvec[0]), 16, 1024);
  constexpr uint64_t NV_STEP = (uint64_t)(KT * 2048) >> 4;

  auto load_qkc = [&](int item) {   // Q, K and the entry-state tile of an item
    const int bh = item / NC, sc = item % NC, b = bh / p.NH, h = bh % p.NH, tok0 = mem_chunk(sc, NC, rev) * L;
    mbar_arrive_expect_tx(&sm.bar_q, KT * TILE);
    for (int kt = 0; kt < KT; ++kt) tma_load_4d(sm.q + kt * TILE, &maps.q, &sm.bar_q, kt * 64, tok0, h, b);
    mbar_arrive_expect_tx(&sm.bar_k, KT * TILE);
    for (int kt = 0; kt < KT; ++kt) tma_load_4d(sm.k + kt * TILE, &maps.k, &sm.bar_k, kt * 64, tok0, h, b);
    mbar_arrive_expect_tx(&sm.bar_cs, KT * TILE_C);
    for (int kt = 0; kt < KT; ++kt) tma_load_2d(sm.cs + kt * TILE_C, &maps.cs, &sm.bar_cs, kt * 64, item * DH);
  };
  auto load_v = [&](int item) {
    const int bh = item / NC, sc = item % NC, b = bh / p.NH, h = bh % p.NH, tok0 = mem_chunk(sc, NC, rev) * L;
    mbar_arrive_expect_tx(&sm.bar_v, KT * TILE);
    for (int kt = 0; kt < KT; ++kt) tma_load_4d(sm.v + kt * TILE, &maps.v, &sm.bar_v, kt * 64, tok0, h, b);
  };
  auto issue_mma1 = [&](int n) {   // S = Q K^T, G = Q Cs, qn = Q [n_hi n_lo]
    constexpr uint32_t idS = make_idesc_bf16(128, 128, 0, 0), idG = make_idesc_bf16(128, DH, 0, 1);
    constexpr uint32_t idN = make_idesc_bf16(128, 16, 0, 0);
    const uint64_t dNv = dNv0 + (n % 3) * NV_STEP;
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) umma_bf16_ss(tS, dQ + kstep(ks), dKk + kstep(ks), idS, ks > 0);
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) umma_bf16_ss(tG, dQ + kstep(ks), dCs + mnstep(ks), idG, ks > 0);
#pragma unroll
    for (int ks = 0; ks < DH / 16; ++ks) umma_bf16_ss(tQN, dQ + kstep(ks), dNv + kstep(ks, 2048), idN, ks > 0);
    umma_commit(&sm.bar_m1);
  };
  // gate vectors and the n operand tile of an item, by the gate warp
  auto prep_item = [&](int item, int slot) {
    const int bh = item / NC, sc = item % NC, b = bh / p.NH, h = bh % p.NH;
    gates_warp_fwd(sm.g[slot], p, b, h, mem_chunk(sc, NC, rev), lane, ms_all[item]);
    for (int d = lane; d < DH; d += 32) {
      const float nv = ns_all[(size_t)item * DH + d];
      const __nv_bfloat16 hi = __float2bfloat16_rn(nv);
      const __nv_bfloat16 lo = __float2bfloat16_rn(nv - __bfloat162float(hi));
      uint8_t* base = sm.nvec[slot] + (d >> 6) * 2048;
      *reinterpret_cast<__nv_bfloat16*>(base + swz128(0, d & 63)) = hi;
      *reinterpret_cast<__nv_bfloat16*>(base + swz128(1, d & 63)) = lo;
    }
    fence_proxy_async_smem();
  };

  const int item0 = blockIdx.x;   // the grid is never larger than n_items
  if (issuer) { load_qkc(item0); load_v(item0); }
  if (gatew) {
    prep_item(item0, 0);
    if (item0 + (int)gridDim.x < n_items) prep_item(item0 + gridDim.x, 1);
  }
  __syncthreads();
  if (issuer) {
    mbar_wait(&sm.bar_q, 0); mbar_wait(&sm.bar_k, 0); mbar_wait(&sm.bar_cs, 0);
    tc_fence_after();
    issue_mma1(0);
  }

  int n = 0;
  for (int item = item0; item < n_items; item += gridDim.x, ++n) {
    const uint32_t ph = n & 1;
    const int next = item + gridDim.x;
    const bool has_next = next < n_items;
    if (gatew) {
      if (next + (int)gridDim.x < n_items) prep_item(next + gridDim.x, (n + 2) % 3);
      __syncthreads();
      continue;
    }
    const GateBuf& G = sm.g[n % 3];
    const int bh = item / NC, sc = item % NC, b = bh / p.NH, h = bh % p.NH;
    const int tok0 = mem_chunk(sc, NC, rev) * L;

    mbar_wait(&sm.bar_m1, ph);       // S, G, qn of this item are in TMEM; Q, K, Cs are dead
    tc_fence_after();
    if (issuer && has_next) load_qkc(next);

    // ---- P = s * S * exp2(u2_j - M2_t), causal: one 32x32 block per warp; partial row sums ----
    if (compute) {
      const float M2t = G.M2[row] - l2s;
      const bool full = rev ? (cq > rg) : (cq < rg);   // forward: keys j <= t ; reverse: keys j >= t
      const bool diag = (cq == rg);
      uint32_t packed[16];
      float rowsum = 0.f;
      if (full || diag) {
        float s[32];
        tmem_ld32(tS + lane_sel + cq * 32, s);
        const uint32_t cbits = causal_bits(full, !rev, lane);
        tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; x += 4) {
          const float4 u4 = *reinterpret_cast<const float4*>(&G.u2[cq * 32 + x]);
          const float uu[4] = {u4.x, u4.y, u4.z, u4.w};
          float pv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool keep = (cbits >> (x + e)) & 1u;
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
      sm.part_rs[cq][row] = rowsum;
#pragma unroll
      for (int x = 0; x < 4; ++x) {
        const int j = cq * 32 + x * 8;
        *reinterpret_cast<uint4*>(sm.p + (j >> 6) * TILE + swz128(row, j & 63)) =
            make_uint4(packed[4 * x], packed[4 * x + 1], packed[4 * x + 2], packed[4 * x + 3]);
      }
    }
    fence_proxy_async_smem();
    tc_fence_before();
    named_sync(2, GT0);

    // ---- MMA2: H = P V (into the S columns) --------------------------------------------------
    if (issuer) {
      mbar_wait(&sm.bar_v, ph);
      tc_fence_after();
      constexpr uint32_t idH = make_idesc_bf16(128, DH, 0, 1);
#pragma unroll
      for (int ks = 0; ks < L / 16; ++ks) umma_bf16_ss(tS, dP + kstep(ks), dVmn + mnstep(ks), idH, ks > 0);
      umma_commit(&sm.bar_m2);
    }
    // row normaliser (backends.py:249-252): n_t = sum_j P_tj + w s (q_t . n_prev)
    float inv = 0.f, ws = 0.f;
    if (compute) {
      float qn[16];
      tmem_ld16(tQN + lane_sel, qn);
      tmem_ld_wait();
      ws = G.w[row] * scale;
      const float mrow = G.mrow[row];
      const float nr = (sm.part_rs[0][row] + sm.part_rs[1][row] + sm.part_rs[2][row] + sm.part_rs[3][row]) + ws * (qn[0] + qn[1]);
      inv = 1.f / (fmaxf(fabsf(nr), __expf(-mrow)) + p.eps);
      const int tok = tok0 + row;
      if (cq == 0 && p.n_row && tok < S) {
        p.n_row[(int64_t)bh * S + tok] = nr;
        p.m_row[(int64_t)bh * S + tok] = mrow;
      }
    }
    mbar_wait(&sm.bar_m2, ph);
    tc_fence_after();
    if (issuer && has_next) load_v(next);

    // ---- epilogue: h = (H + w s G) / N, packed in registers; then MMA1 of the next item ------
    uint32_t hpk[16];
    if (cq < NB) {
      float hi[32], gg[32];
      tmem_ld32(tS + lane_sel + cq * 32, hi);
      tmem_ld32(tG + lane_sel + cq * 32, gg);
      tmem_ld_wait();
#pragma unroll
      for (int x = 0; x < 32; x += 2)
        hpk[x / 2] = pack_bf16x2((hi[x] + ws * gg[x]) * inv, (hi[x + 1] + ws * gg[x + 1]) * inv);
    }
    tc_fence_before();
    __syncthreads();   // end of item: TMEM free, next gates / n tile published (gate warp joins here)
    if (issuer && has_next) {
      mbar_wait(&sm.bar_q, ph ^ 1); mbar_wait(&sm.bar_k, ph ^ 1); mbar_wait(&sm.bar_cs, ph ^ 1);
      tc_fence_after();
      issue_mma1(n + 1);
    }
    if (cq < NB) {
      const int tok = tok0 + row;
      if (tok < S) {
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.h.ptr) + (int64_t)b * p.h.stride_b + (int64_t)h * p.h.stride_h +
                             (int64_t)tok * p.h.stride_s + cq * 32;
        if (dv_true >= DH) {
          store_row32(dst, hpk);
        } else {
#pragma unroll
          for (int x = 0; x < 4; ++x)
            if (cq * 32 + x * 8 < dv_true)
              *reinterpret_cast<uint4*>(dst + x * 8) = make_uint4(hpk[4 * x], hpk[4 * x + 1], hpk[4 * x + 2], hpk[4 * x + 3]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

template <class K>
int prep(K kernel, size_t smem, const char* name) {
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
int launch_fwd(const mlstm_params& p, cudaStream_t st) {
  const StateLayout lay(p.B, p.NH, p.S, DH);
  if (!p.states || p.states_bytes < lay.total) {
    set_error("forward needs a state workspace of %zu bytes (mlstm_b200_state_bytes), got %zu", lay.total,
              p.states ? p.states_bytes : (size_t)0);
    return MLSTM_ERR_WORKSPACE;
  }
  FwdMaps maps;
  int r = 0;
  r |= make_act_tmap(&maps.q, p.q.ptr, p.B, p.NH, p.S, DH, p.q.stride_b, p.q.stride_h, p.q.stride_s, L);
  r |= make_act_tmap(&maps.k, p.k.ptr, p.B, p.NH, p.S, DH, p.k.stride_b, p.k.stride_h, p.k.stride_s, L);
  r |= make_act_tmap(&maps.v, p.v.ptr, p.B, p.NH, p.S, DH, p.v.stride_b, p.v.stride_h, p.v.stride_s, L);
  const int NC = num_chunks(p.S);
  const int n_items = p.B * p.NH * NC;
  r |= make_state_tmap(&maps.cs, reinterpret_cast<uint8_t*>(p.states) + lay.cs_off, (size_t)n_items * DH, DH);
  if (r) {
    set_error("cuTensorMapEncodeTiled failed (%d): pointers must be 16-byte aligned, strides multiples of 8 elements", r);
    return r == -1 ? MLSTM_ERR_NO_DEVICE : MLSTM_ERR_INVALID_ARG;
  }
  int rc;
  const size_t smS = sizeof(SmemS<DH>), smP = sizeof(SmemP<DH>);
  if ((rc = prep(tc_state_fwd_kernel<DH>, smS, "tc_state_fwd"))) return rc;
  if ((rc = prep(tc_fwd_par_kernel<DH>, smP, "tc_fwd_par"))) return rc;
  const int sms_s = sm_count_of(p.q.ptr);
  const int nsl = (DH == 128 && 2 * p.B * p.NH <= sms_s) ? 2 : 1;   // 64-column value slices while they fit the SMs
  tc_state_fwd_kernel<DH><<<dim3(p.B * p.NH, nsl), dim3(NT), smS, st>>>(maps, p);
  if ((rc = launched("tc_state_fwd"))) return rc;
  const int sms = sms_s;
  const int grid = n_items < sms ? n_items : sms;
  tc_fwd_par_kernel<DH><<<dim3(grid), dim3(NT), smP, st>>>(maps, p, resolve_scale(p), n_items, true_extent(p.h.ptr, DH));
  return launched("tc_fwd_par");
}

}  // namespace

// State walk for head dims above 128 (mlstm_tc_256.cu): DHF / 128 row blocks x nsl value slices per (batch, head), each an
// independent recurrence on a [128][DHF / nsl] block of C.  `cs_store` has a box of 128 rows.
int tc_state_fwd_blocks(const mlstm_params& p, cudaStream_t st, const CUtensorMap& mk, const CUtensorMap& mv,
                        const CUtensorMap& cs_store, int nsl) {
  FwdMaps maps;
  maps.q = mk; maps.k = mk; maps.v = mv; maps.cs = cs_store;
  int rc;
  const size_t smS = sizeof(SmemS<128>);
  if ((rc = prep(tc_state_fwd_kernel<128>, smS, "tc_state_fwd"))) return rc;
  tc_state_fwd_kernel<128><<<dim3((p.DHQK / 128) * nsl, p.B * p.NH), dim3(NT), smS, st>>>(maps, p);
  return launched("tc_state_fwd");
}

int tc256_fwd(const mlstm_params& p, cudaStream_t st);

int tc_fwd_two_phase(const mlstm_params& p, cudaStream_t st) {
  if (p.DHQK == 256) return tc256_fwd(p, st);
  if (p.DHQK == 64) return launch_fwd<64>(p, st);
  return launch_fwd<128>(p, st);
}

}  // namespace mlstm

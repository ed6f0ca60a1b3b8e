// tcgen05 / TMEM / TMA kernels of the mLSTM cell for head dim 256 (bf16 I/O; reference shapes: qkv_block_size -> DH,
// vision_lstm2.py:416-417; math: backends.py:149-263).
//
// At DH = 256 one chunk's operands no longer fit a CTA: q, k tiles are 64 KB each, the [dk][dv] state 128 KB in bf16.
// The family is therefore chunk-parallel throughout (per-chunk states in HBM, as mlstm_tc_fwd2p.cu / mlstm_tc_bwd.cu
// at DH <= 128), and every 128 x 128 product whose contraction runs over a 256-wide dimension STREAMS that dimension
// through a three-stage TMA ring in four 64-wide slices:
//
//   item = (batch, head, chunk, output half hf): 128 rows x 128 output columns
//     MMA1   tS  (+)= T0_s T1_s^T                         s = 0..3      (the 128 x 128 score-like tile)
//            tX  (+)= T0_s St_s                            s = 0..3      (the product with the chunk state, half hf)
//     SIMT   gated bf16 tile P from tS  -> TMEM            (same arithmetic as the DH <= 128 kernels)
//     MMA2   out = P T2[:, hf]                             (T2 half resident, 32 KB)
//     SIMT   epilogue -> staged in the dead T2 half -> one TMA tile store
//
//            T0 (slices)  T1 (slices)  St (slice of the state)          T2 half     out
//   F        Q            K            Cs [dk slice][dv half]           V           h    = (P V + w s Q Cs) / N
//   A        dH           V            Cs [dk half][dv slice]           K           dq
//   B1       K            Q            dCs[dk slice][dv half]           dH          dv
//   B2       V            dH           dCs[dk half][dv slice]           Q           dk
//
// The chunk states come from the DH = 128 state walks run on (256/128)^2 independent 128 x 128 blocks of C / dC
// (tc_state_fwd_blocks, tc_state_bwd_blocks); di, df from the shared scan kernel.  Both halves of a chunk are
// neighbouring items, so the slices the second one loads are L2 hits.
#include "tc_common.cuh"

namespace mlstm {

int tc_state_fwd_blocks(const mlstm_params& p, cudaStream_t st, const CUtensorMap& mk, const CUtensorMap& mv,
                        const CUtensorMap& cs_store, int nsl);
int tc_state_bwd_blocks(const mlstm_params& p, cudaStream_t st, const CUtensorMap& mq, const CUtensorMap& mdh,
                        const CUtensorMap& mcs, const CUtensorMap& mdcs);
int tc_dfscan_launch(const mlstm_params& p, cudaStream_t st);
void tc_bwd_layout(const mlstm_params& p, size_t* dn_off, size_t* rpart_off, size_t* kpart_off, size_t* dcs_off, size_t* dns_off,
                   size_t* total);

namespace {

using namespace tc;

constexpr int DHF = 256;          // head dim
constexpr int HW = 128;           // output half width
constexpr int NSL = DHF / 64;     // 64-wide slices of the streamed dimension
constexpr int NST = 3;            // ring stages
constexpr int STAGE = 3 * TILE;   // T0 slice | T1 slice | state slice, 16 KB each

enum { MODE_F = 0, MODE_A = 1, MODE_B1 = 2, MODE_B2 = 3 };

struct Maps256 { CUtensorMap t0, t1, t2, st, out, t3; };   // t3: q (A) / k (B2), the rows of R = q.dq / K = k.dk

struct Scratch256 {   // device pointers into p.workspace / p.states
  float* dn; float* rpart; float* kpart;
  const float* ns; const float* ms; const float* dns;
  size_t rows_total;
};

template <int MODE>
struct Smem256 {
  // F: nvec, K-major [16][256] tiles (row 0 = hi(n), row 1 = lo(n)), ring of 3 | A, B2: t3, the q / k half tile [128][128]
  static constexpr int AUX = (MODE == MODE_F) ? 3 * NSL * 2048 : ((MODE == MODE_B1) ? 1024 : 2 * TILE);
  alignas(1024) uint8_t ring[NST][STAGE];
  alignas(1024) uint8_t t2[2 * TILE];              // resident MMA2 operand half [128][128]; then the output staging tile
  alignas(1024) uint8_t aux[AUX];
  GateBuf g[3];                                    // the gate warp runs two items ahead
  alignas(16) float vecf[3][HW];                   // ns half (A) / dns half (B2) of the item
  float part[4][L];
  uint64_t full[NST], empty[NST], bar_t2, bar_t3, bar_m1, bar_m2;
  uint32_t tmem_base;
};

template <int MODE>
__global__ void __launch_bounds__(NT, 1) tc256_par_kernel(const __grid_constant__ Maps256 maps, const mlstm_params p,
                                                          const Scratch256 sx, const float scale, const int n_items) {
  constexpr bool IS_F = (MODE == MODE_F), IS_A = (MODE == MODE_A), IS_B1 = (MODE == MODE_B1), IS_B2 = (MODE == MODE_B2);
  constexpr bool ROWQ = IS_F || IS_A;             // thread row = query t (else key j)
  constexpr bool ST_MN = IS_F || IS_B1;           // state slice: [64 streamed rows][128 half columns], MN-major B operand
  // K-major state slices sit right behind the T1 slice and continue its row pattern: S and the state product are ONE
  // N = 256 MMA per k-step (A, B2).  With an MN-major state slice (F, B1) the two products are separate instructions,
  // issued from two lanes (a single lane issues ~100 cycles per MMA, the tensor pipe needs 64).
  constexpr bool MERGED = !ST_MN;
  constexpr bool HAS_T3 = IS_A || IS_B2;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  Smem256<MODE>& sm = *reinterpret_cast<Smem256<MODE>*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool compute = tid < CT, issuer = tid == CT, gatew = tid >= GT0;
  const bool issuer2 = !MERGED && tid == 32;      // second MMA lane (lane 0 of compute warp 1): the state product
  const int rg = warp & 3, cq = compute ? (warp >> 2) : 4, row = rg * 32 + lane;
  const int S = p.S, NC = num_chunks(S);
  const bool rev = p.reverse != 0;
  const float l2s = log2f(scale);
  const int cta = blockIdx.x, ncta = gridDim.x;

  if (issuer) {
    tma_prefetch_desc(&maps.t0); tma_prefetch_desc(&maps.t1); tma_prefetch_desc(&maps.t2); tma_prefetch_desc(&maps.st);
    tma_prefetch_desc(&maps.out);
    for (int i = 0; i < NST; ++i) { mbar_init(&sm.full[i], 1); mbar_init(&sm.empty[i], MERGED ? 1 : 2); }
    mbar_init(&sm.bar_t2, 1); mbar_init(&sm.bar_t3, 1); mbar_init(&sm.bar_m1, MERGED ? 1 : 2); mbar_init(&sm.bar_m2, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, 512);
  if (IS_F)
    for (int e = tid; e < 3 * NSL * 2048 / 16; e += NT) reinterpret_cast<uint4*>(sm.aux)[e] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // tS: MMA1 tile (F, A: MMA2 writes its result over it) | tX: product with the chunk state | tO: MMA2 result (B) |
  // tP: gated tile, packed bf16 | tQN: q . n (F)
  const uint32_t tm = sm.tmem_base, tS = tm, tX = tm + 128, tO = tm + 256, tP = tm + 384, tQN = tm + 448;
  const uint32_t lane_sel = (uint32_t)(rg * 32) << 16;

  // item it2 -> (chunk item ci = (b*NH + h)*NC + sc, half hf)
  auto coords = [&](int it2, int& ci, int& hf, int& b, int& h, int& tok0) {
    ci = it2 >> 1; hf = it2 & 1;
    const int bh = ci / NC, sc = ci % NC;
    b = bh / p.NH; h = bh % p.NH; tok0 = mem_chunk(sc, NC, rev) * L;
  };

  // ---- flat slice sequence of this CTA: slice g = 4 * n + s of its n-th item ---------------------------------
  const int my_items = (n_items - cta + ncta - 1) / ncta;
  const int n_slices = my_items * NSL;
  auto load_slice = [&](int g) {
    const int slot = g % NST, n = g / NSL, s = g % NSL;
    if (g >= NST) mbar_wait(&sm.empty[slot], ((g / NST) - 1) & 1);   // the MMAs that read the slot's previous slice are complete
    int ci, hf, b, h, tok0; coords(cta + n * ncta, ci, hf, b, h, tok0);
    uint8_t* dst = sm.ring[slot];
    mbar_arrive_expect_tx(&sm.full[slot], STAGE);
    tma_load_4d(dst, &maps.t0, &sm.full[slot], s * 64, tok0, h, b);
    tma_load_4d(dst + TILE, &maps.t1, &sm.full[slot], s * 64, tok0, h, b);
    if (ST_MN) {   // rows: streamed slice of dk; columns: the half's two 64-wide blocks
      tma_load_2d(dst + 2 * TILE, &maps.st, &sm.full[slot], hf * HW, ci * DHF + s * 64);
      tma_load_2d(dst + 2 * TILE + 8192, &maps.st, &sm.full[slot], hf * HW + 64, ci * DHF + s * 64);
    } else {       // rows: the half's 128 dk rows; columns: streamed slice of dv
      tma_load_2d(dst + 2 * TILE, &maps.st, &sm.full[slot], s * 64, ci * DHF + hf * HW);
    }
    if (s == NSL - 2) {   // the item's last slice can only land once its first one has been consumed (3-slot ring, 4 slices):
      const int s1 = NSL - 1;   // pull it into L2 now, so that the late load is an L2 hit
      tma_prefetch_4d(&maps.t0, s1 * 64, tok0, h, b);
      tma_prefetch_4d(&maps.t1, s1 * 64, tok0, h, b);
      if (ST_MN) {
        tma_prefetch_2d(&maps.st, hf * HW, ci * DHF + s1 * 64);
        tma_prefetch_2d(&maps.st, hf * HW + 64, ci * DHF + s1 * 64);
      } else {
        tma_prefetch_2d(&maps.st, s1 * 64, ci * DHF + hf * HW);
      }
    }
  };
  int next_load = 0;
  auto pump_loads = [&](int upto) {   // issue every load up to slice index `upto` (inclusive) that has not gone out yet
    for (; next_load <= upto && next_load < n_slices; ++next_load) load_slice(next_load);
  };
  auto load_t2 = [&](int it2) {
    int ci, hf, b, h, tok0; coords(it2, ci, hf, b, h, tok0);
    mbar_arrive_expect_tx(&sm.bar_t2, 2 * TILE);
    for (int kt = 0; kt < 2; ++kt) tma_load_4d(sm.t2 + kt * TILE, &maps.t2, &sm.bar_t2, hf * HW + kt * 64, tok0, h, b);
  };
  auto load_t3 = [&](int it2) {
    int ci, hf, b, h, tok0; coords(it2, ci, hf, b, h, tok0);
    mbar_arrive_expect_tx(&sm.bar_t3, 2 * TILE);
    for (int kt = 0; kt < 2; ++kt) tma_load_4d(sm.aux + kt * TILE, &maps.t3, &sm.bar_t3, hf * HW + kt * 64, tok0, h, b);
  };
  // MMA1 of the CTA's n-th item, slice by slice as the ring fills.  lane2 = false: the control lane (S, or S | state
  // product merged; it also keeps the ring loads going); lane2 = true: the second lane (state product, F: q.n).
  auto stream_mma1 = [&](int n, bool lane2) {
    constexpr uint32_t idS = make_idesc_bf16(128, MERGED ? 256 : 128, 0, 0);
    constexpr uint32_t idX = make_idesc_bf16(128, HW, 0, 1);
    constexpr uint32_t idN = make_idesc_bf16(128, 16, 0, 0);
    for (int s = 0; s < NSL; ++s) {
      const int g = n * NSL + s, slot = g % NST;
      mbar_wait(&sm.full[slot], (g / NST) & 1);
      tc_fence_after();
      const uint32_t base = smem_u32(sm.ring[slot]);
      const uint64_t d0 = make_sdesc(base, 16, 1024);
      if (!lane2) {
        const uint64_t d1 = make_sdesc(base + TILE, 16, 1024);   // MERGED: 256 rows = T1 slice, then the K-major state slice
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_bf16_ss(tS, d0 + kstep(ks), d1 + kstep(ks), idS, (s > 0 || ks > 0) ? 1u : 0u);
      } else {
        const uint64_t dst_ = make_sdesc(base + 2 * TILE, 8192, 1024);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_bf16_ss(tX, d0 + kstep(ks), dst_ + mnstep(ks), idX, (s > 0 || ks > 0) ? 1u : 0u);
        if (IS_F) {
          const uint64_t dNv = make_sdesc(smem_u32(sm.aux) + (n % 3) * (NSL * 2048) + s * 2048, 16, 1024);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) umma_bf16_ss(tQN, d0 + kstep(ks), dNv + kstep(ks), idN, (s > 0 || ks > 0) ? 1u : 0u);
        }
      }
      umma_commit(&sm.empty[slot]);
      if (!lane2) pump_loads(g + 2);
    }
    umma_commit(&sm.bar_m1);
  };
  // gate vectors and per-item vectors, by the gate warp
  auto prep_item = [&](int it2, int slot) {
    int ci, hf, b, h, tok0; coords(it2, ci, hf, b, h, tok0);
    const int bh = ci / NC, sc = ci % NC;
    if (IS_F) {
      gates_warp_fwd(sm.g[slot], p, b, h, mem_chunk(sc, NC, rev), lane, sx.ms[ci]);
      for (int d = lane; d < DHF; d += 32) {
        const float nv = sx.ns[(size_t)ci * DHF + d];
        const __nv_bfloat16 hi = __float2bfloat16_rn(nv);
        const __nv_bfloat16 lo = __float2bfloat16_rn(nv - __bfloat162float(hi));
        uint8_t* base = sm.aux + slot * (NSL * 2048) + (d >> 6) * 2048;
        *reinterpret_cast<__nv_bfloat16*>(base + swz128(0, d & 63)) = hi;
        *reinterpret_cast<__nv_bfloat16*>(base + swz128(1, d & 63)) = lo;
      }
      fence_proxy_async_smem();
    } else {
      gates_warp_bwd(sm.g[slot], p, b, h, bh, mem_chunk(sc, NC, rev), lane, IS_B2 ? sx.dn : nullptr);
      if (IS_A || IS_B2)
        for (int d = lane; d < HW; d += 32) sm.vecf[slot][d] = (IS_A ? sx.ns : sx.dns)[(size_t)ci * DHF + hf * HW + d];
    }
    __syncwarp();
  };

  if (issuer) { pump_loads(NST - 1); load_t2(cta); if (HAS_T3) load_t3(cta); }
  if (gatew) {
    prep_item(cta, 0);
    if (cta + ncta < n_items) prep_item(cta + ncta, 1);
  }
  __syncthreads();
  if (issuer) stream_mma1(0, false);
  if (issuer2) stream_mma1(0, true);

  int n = 0;
  for (int it2 = cta; it2 < n_items; it2 += ncta, ++n) {
    const uint32_t ph = n & 1;
    const bool has_next = it2 + ncta < n_items;
    if (gatew) {
      if (it2 + 2 * ncta < n_items) prep_item(it2 + 2 * ncta, (n + 2) % 3);
      __syncthreads();
      continue;
    }
    const GateBuf& G = sm.g[n % 3];
    const float* vecf = sm.vecf[n % 3];
    int ci, hf, b, h, tok0; coords(it2, ci, hf, b, h, tok0);
    const int bh = ci / NC;
    const int tok = tok0 + row;
    const bool row_ok = compute && tok < S;
    const size_t grow = (size_t)bh * S + tok;

    // ---- A: dn_t = dnf_t (dh_t . h_t) over all 256 columns, rows read straight from global (L2) ----------------
    float dn_row = 0.f;
    if (IS_A) {
      // a warp per 512-byte row (16 bytes per lane: fully coalesced), 8 rows per warp, 4 at a time
      if (compute) {
        const __nv_bfloat16* hbase = reinterpret_cast<const __nv_bfloat16*>(p.h.ptr) + (int64_t)b * p.h.stride_b + (int64_t)h * p.h.stride_h + lane * 8;
        const __nv_bfloat16* dbase = reinterpret_cast<const __nv_bfloat16*>(p.dh.ptr) + (int64_t)b * p.dh.stride_b + (int64_t)h * p.dh.stride_h + lane * 8;
#pragma unroll
        for (int r0 = 0; r0 < 8; r0 += 4) {
          uint4 wh[4], wd[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const int t = tok0 + warp * 8 + r0 + r;
            wh[r] = wd[r] = make_uint4(0, 0, 0, 0);
            if (t < S) {
              wh[r] = *reinterpret_cast<const uint4*>(hbase + (int64_t)t * p.h.stride_s);
              wd[r] = *reinterpret_cast<const uint4*>(dbase + (int64_t)t * p.dh.stride_s);
            }
          }
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&wh[r]);
            const __nv_bfloat162* dd = reinterpret_cast<const __nv_bfloat162*>(&wd[r]);
            float part = 0.f;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 a = __bfloat1622float2(hh[e]), c2 = __bfloat1622float2(dd[e]);
              part = fmaf(a.x, c2.x, fmaf(a.y, c2.y, part));
            }
            part = warp_sum(part);
            const int rr = warp * 8 + r0 + r;
            if (lane == 0) {
              const float dnr = G.dnf[rr] * part;
              sm.part[0][rr] = dnr;
              if (hf == 0 && tok0 + rr < S) sx.dn[(size_t)bh * S + tok0 + rr] = dnr;
            }
          }
        }
        named_sync(3, CT);
        dn_row = sm.part[0][row];
      }
    }
    mbar_wait(&sm.bar_m1, ph);
    tc_fence_after();
    if (issuer) pump_loads((n + 1) * NSL + NST - 1);   // the ring is free: the next item's first slices

    // ---- gated bf16 tile: one 32x32 block per warp ------------------------------------------------------------
    float rowsum = 0.f;
    if (compute) {
      const bool full = ROWQ ? (rev ? (cq > rg) : (cq < rg)) : (rev ? (cq < rg) : (cq > rg));
      const bool diag = (cq == rg);
      uint32_t packed[16];
      if (full || diag) {
        float a[32];
        tmem_ld32(tS + lane_sel + cq * 32, a);
        const uint32_t cbits = causal_bits(full, ROWQ ? !rev : rev, lane);
        tmem_ld_wait();
        const float r0 = IS_F ? G.M2[row] - l2s : (IS_A ? G.M2[row] : (IS_B1 ? G.u2[row] + l2s : G.u2[row]));
        const float r1 = IS_A ? G.invN[row] : 0.f;
#pragma unroll
        for (int x = 0; x < 32; x += 4) {
          const int c0 = cq * 32 + x;
          const float4 e4 = *reinterpret_cast<const float4*>(ROWQ ? &G.u2[c0] : (IS_B1 ? &G.c2[c0] : &G.M2[c0]));
          const float ee[4] = {e4.x, e4.y, e4.z, e4.w};
          float in4[4] = {0.f, 0.f, 0.f, 0.f}, dn4[4] = {0.f, 0.f, 0.f, 0.f};
          if (IS_B2) {
            const float4 i4 = *reinterpret_cast<const float4*>(&G.invN[c0]);
            const float4 d4 = *reinterpret_cast<const float4*>(&G.dn[c0]);
            in4[0] = i4.x; in4[1] = i4.y; in4[2] = i4.z; in4[3] = i4.w;
            dn4[0] = d4.x; dn4[1] = d4.y; dn4[2] = d4.z; dn4[3] = d4.w;
          }
          float pv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const bool keep = (cbits >> (x + e)) & 1u;
            float val;
            if (IS_F) val = a[x + e] * ex2(ee[e] - r0);
            else if (IS_A) val = fmaf(a[x + e], r1, dn_row) * ex2(ee[e] - r0);
            else if (IS_B1) val = a[x + e] * ex2(r0 - ee[e]);
            else val = fmaf(a[x + e], in4[e], dn4[e]) * ex2(r0 - ee[e]);
            pv[e] = keep ? val : 0.f;
            rowsum += pv[e];
          }
          packed[x / 2] = pack_bf16x2(pv[0], pv[1]);
          packed[x / 2 + 1] = pack_bf16x2(pv[2], pv[3]);
        }
      } else {
#pragma unroll
        for (int x = 0; x < 16; ++x) packed[x] = 0u;
      }
      if (IS_F) sm.part[cq][row] = rowsum;
      (void)rowsum;
      tmem_st16(tP + lane_sel + cq * 16, packed);
      tmem_st_wait();
    }
    tc_fence_before();
    named_sync(2, GT0);

    // ---- MMA2: out = P T2half --------------------------------------------------------------------------------
    if (issuer) {
      mbar_wait(&sm.bar_t2, ph);
      tc_fence_after();
      constexpr uint32_t id2 = make_idesc_bf16(128, HW, 0, 1);
      const uint64_t d2mn = make_sdesc(smem_u32(sm.t2), TILE, 1024);
#pragma unroll
      for (int ks = 0; ks < L / 16; ++ks) umma_bf16_ts((IS_F || IS_A) ? tS : tO, tP + ks * 8, d2mn + mnstep(ks), id2, ks > 0);
      umma_commit(&sm.bar_m2);
    }
    // F: row normaliser (backends.py:249-252): n_t = sum_j P_tj + w s (q_t . n_prev)
    float inv = 0.f, ws = 0.f;
    if (IS_F && compute) {
      float qn[16];
      tmem_ld16(tQN + lane_sel, qn);
      tmem_ld_wait();
      ws = G.w[row] * scale;
      const float mrow = G.mrow[row];
      const float nr = (sm.part[0][row] + sm.part[1][row] + sm.part[2][row] + sm.part[3][row]) + ws * (qn[0] + qn[1]);
      inv = 1.f / (fmaxf(fabsf(nr), __expf(-mrow)) + p.eps);
      if (cq == 0 && hf == 0 && p.n_row && row_ok) {
        p.n_row[grow] = nr;
        p.m_row[grow] = mrow;
      }
    }
    mbar_wait(&sm.bar_m2, ph);
    tc_fence_after();

    // ---- epilogue, 16 columns at a time: outputs staged in the dead T2 half ------------------------------------------
    if (compute) {
      float psum = 0.f;
      const float wt = G.w[row], invN = G.invN[row], kwj = G.kw[row];
      if (HAS_T3) mbar_wait(&sm.bar_t3, ph);
#pragma unroll
      for (int hc = 0; hc < 2; ++hc) {
        const int c0 = cq * 32 + hc * 16;   // first column (within the half) of this pass
        float acc[16], gg[16];
        tmem_ld16(((IS_F || IS_A) ? tS : tO) + lane_sel + c0, acc);
        tmem_ld16(tX + lane_sel + c0, gg);
        tmem_ld_wait();
        uint32_t opk[8];
        if (IS_F) {
#pragma unroll
          for (int x = 0; x < 16; x += 2)
            opk[x / 2] = pack_bf16x2((acc[x] + ws * gg[x]) * inv, (acc[x + 1] + ws * gg[x + 1]) * inv);
        } else if (IS_B1) {   // dv = E^T dH + kw (K dC)
#pragma unroll
          for (int x = 0; x < 16; x += 2) opk[x / 2] = pack_bf16x2(fmaf(kwj, gg[x], acc[x]), fmaf(kwj, gg[x + 1], acc[x + 1]));
        } else {              // A: dq, R = q . dq   |   B2: dk, K = k . dk   (q / k rows from the T3 tile)
#pragma unroll
          for (int x8 = 0; x8 < 16; x8 += 8) {
            const int col = c0 + x8;
            const uint4 w = *reinterpret_cast<const uint4*>(sm.aux + (col >> 6) * TILE + swz128(row, col & 63));
            const __nv_bfloat162* qq = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int x = x8 + 2 * e;
              const float2 xr = __bfloat1622float2(qq[e]);
              float o0, o1;
              if (IS_A) {
                o0 = scale * (acc[x] + wt * fmaf(gg[x], invN, dn_row * vecf[c0 + x]));
                o1 = scale * (acc[x + 1] + wt * fmaf(gg[x + 1], invN, dn_row * vecf[c0 + x + 1]));
              } else {
                o0 = fmaf(kwj, gg[x] + vecf[c0 + x], scale * acc[x]);
                o1 = fmaf(kwj, gg[x + 1] + vecf[c0 + x + 1], scale * acc[x + 1]);
              }
              psum = fmaf(xr.x, o0, fmaf(xr.y, o1, psum));
              opk[x / 2] = pack_bf16x2(o0, o1);
            }
          }
        }
#pragma unroll
        for (int x4 = 0; x4 < 2; ++x4) {
          const int col = c0 + x4 * 8;
          *reinterpret_cast<uint4*>(sm.t2 + (col >> 6) * TILE + swz128(row, col & 63)) =
              make_uint4(opk[4 * x4], opk[4 * x4 + 1], opk[4 * x4 + 2], opk[4 * x4 + 3]);
        }
      }
      if (HAS_T3 && row_ok) (IS_A ? sx.rpart : sx.kpart)[(size_t)(hf * 4 + cq) * sx.rows_total + grow] = psum;
      (void)wt; (void)invN; (void)kwj; (void)psum;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();   // end of item: accumulators consumed, next gates published (the gate warp joins here)
    if (issuer) {
      for (int kt = 0; kt < 2; ++kt) tma_store_4d(&maps.out, sm.t2 + kt * TILE, hf * HW + kt * 64, tok0, h, b);
      tma_store_commit();
      if (has_next) {
        if (HAS_T3) load_t3(it2 + ncta);   // consumed by the epilogue just finished
        stream_mma1(n + 1, false);
        tma_store_wait_read<0>();     // the staged output has left shared memory: T2 of the next item may land
        load_t2(it2 + ncta);
      }
    }
    if (issuer2 && has_next) stream_mma1(n + 1, true);
  }
  if (issuer) tma_store_wait_all<0>();
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

template <int MODE>
int launch_par(const Maps256& m, const mlstm_params& p, const Scratch256& sx, cudaStream_t st, const char* name) {
  int rc;
  const size_t smem = sizeof(Smem256<MODE>);
  if ((rc = prep(tc256_par_kernel<MODE>, smem, name))) return rc;
  const int n_items = p.B * p.NH * num_chunks(p.S) * 2;
  const int sms = sm_count_of(p.q.ptr);
  const int grid = n_items < sms ? n_items : (sms & ~1);   // even: a CTA keeps one output half, neighbours share a chunk
  tc256_par_kernel<MODE><<<dim3(grid), dim3(NT), smem, st>>>(m, p, sx, resolve_scale(p), n_items);
  return launched(name);
}

int tmap_fail(int r) {
  set_error("cuTensorMapEncodeTiled failed (%d): pointers must be 16-byte aligned, strides multiples of 8 elements", r);
  return r == -1 ? MLSTM_ERR_NO_DEVICE : MLSTM_ERR_INVALID_ARG;
}

}  // namespace

int tc256_fwd(const mlstm_params& p, cudaStream_t st) {
  const StateLayout lay(p.B, p.NH, p.S, DHF);
  if (!p.states || p.states_bytes < lay.total) {
    set_error("forward needs a state workspace of %zu bytes (mlstm_b200_state_bytes), got %zu", lay.total,
              p.states ? p.states_bytes : (size_t)0);
    return MLSTM_ERR_WORKSPACE;
  }
  const int n_chunks = p.B * p.NH * num_chunks(p.S);
  uint8_t* sb = reinterpret_cast<uint8_t*>(p.states);
  Maps256 m;
  CUtensorMap cs128;
  int r = 0;
  r |= make_act_tmap(&m.t0, p.q.ptr, p.B, p.NH, p.S, DHF, p.q.stride_b, p.q.stride_h, p.q.stride_s, L);
  r |= make_act_tmap(&m.t1, p.k.ptr, p.B, p.NH, p.S, DHF, p.k.stride_b, p.k.stride_h, p.k.stride_s, L);
  r |= make_act_tmap(&m.t2, p.v.ptr, p.B, p.NH, p.S, DHF, p.v.stride_b, p.v.stride_h, p.v.stride_s, L);
  r |= make_act_tmap(&m.out, p.h.ptr, p.B, p.NH, p.S, DHF, p.h.stride_b, p.h.stride_h, p.h.stride_s, L);
  m.t3 = m.t0;
  r |= make_state_tmap(&m.st, sb + lay.cs_off, (size_t)n_chunks * DHF, DHF, 64);
  r |= make_state_tmap(&cs128, sb + lay.cs_off, (size_t)n_chunks * DHF, DHF, 128);
  if (r) return tmap_fail(r);
  int rc;
  if ((rc = tc_state_fwd_blocks(p, st, m.t1, m.t2, cs128, 2))) return rc;
  Scratch256 sx{};
  sx.ns = reinterpret_cast<const float*>(sb + lay.ns_off);
  sx.ms = reinterpret_cast<const float*>(sb + lay.ms_off);
  sx.rows_total = (size_t)p.B * p.NH * p.S;
  return launch_par<MODE_F>(m, p, sx, st, "tc256_fwd_par");
}

int tc256_bwd(const mlstm_params& p, cudaStream_t st, int part) {
  const StateLayout lay(p.B, p.NH, p.S, DHF);
  if (!p.states || p.states_bytes < lay.total) {
    set_error("backward needs the forward's chunk-state buffer (%zu bytes)", lay.total);
    return MLSTM_ERR_WORKSPACE;
  }
  size_t dn_off, rpart_off, kpart_off, dcs_off, dns_off, total;
  tc_bwd_layout(p, &dn_off, &rpart_off, &kpart_off, &dcs_off, &dns_off, &total);
  const int n_chunks = p.B * p.NH * num_chunks(p.S);
  uint8_t* sb = reinterpret_cast<uint8_t*>(p.states);
  uint8_t* ws = reinterpret_cast<uint8_t*>(p.workspace);
  CUtensorMap mq, mk, mv, mdh, mdq, mdk, mdv, cs128, dcs64, dcs128;
  int r = 0;
  r |= make_act_tmap(&mq, p.q.ptr, p.B, p.NH, p.S, DHF, p.q.stride_b, p.q.stride_h, p.q.stride_s, L);
  r |= make_act_tmap(&mk, p.k.ptr, p.B, p.NH, p.S, DHF, p.k.stride_b, p.k.stride_h, p.k.stride_s, L);
  r |= make_act_tmap(&mv, p.v.ptr, p.B, p.NH, p.S, DHF, p.v.stride_b, p.v.stride_h, p.v.stride_s, L);
  r |= make_act_tmap(&mdh, p.dh.ptr, p.B, p.NH, p.S, DHF, p.dh.stride_b, p.dh.stride_h, p.dh.stride_s, L);
  r |= make_act_tmap(&mdq, p.dq.ptr, p.B, p.NH, p.S, DHF, p.dq.stride_b, p.dq.stride_h, p.dq.stride_s, L);
  r |= make_act_tmap(&mdk, p.dk.ptr, p.B, p.NH, p.S, DHF, p.dk.stride_b, p.dk.stride_h, p.dk.stride_s, L);
  r |= make_act_tmap(&mdv, p.dv.ptr, p.B, p.NH, p.S, DHF, p.dv.stride_b, p.dv.stride_h, p.dv.stride_s, L);
  r |= make_state_tmap(&cs128, sb + lay.cs_off, (size_t)n_chunks * DHF, DHF, 128);
  r |= make_state_tmap(&dcs64, ws + dcs_off, (size_t)n_chunks * DHF, DHF, 64);
  r |= make_state_tmap(&dcs128, ws + dcs_off, (size_t)n_chunks * DHF, DHF, 128);
  if (r) return tmap_fail(r);
  Scratch256 sx{};
  sx.dn = reinterpret_cast<float*>(ws + dn_off);
  sx.rpart = reinterpret_cast<float*>(ws + rpart_off);
  sx.kpart = reinterpret_cast<float*>(ws + kpart_off);
  sx.ns = reinterpret_cast<const float*>(sb + lay.ns_off);
  sx.ms = reinterpret_cast<const float*>(sb + lay.ms_off);
  sx.dns = reinterpret_cast<const float*>(ws + dns_off);
  sx.rows_total = (size_t)p.B * p.NH * p.S;
  int rc;
  if (part != 1) {
    Maps256 m{mdh, mv, mk, cs128, mdq, mq};
    if ((rc = launch_par<MODE_A>(m, p, sx, st, "tc256_bwd_dq"))) return rc;
  }
  if (part != 0) {
    if ((rc = tc_state_bwd_blocks(p, st, mq, mdh, cs128, dcs128))) return rc;
    Maps256 m1{mk, mq, mdh, dcs64, mdv, mq};
    if ((rc = launch_par<MODE_B1>(m1, p, sx, st, "tc256_bwd_dv"))) return rc;
    Maps256 m2{mv, mdh, mq, dcs128, mdk, mk};
    if ((rc = launch_par<MODE_B2>(m2, p, sx, st, "tc256_bwd_dk"))) return rc;
    if ((rc = tc_dfscan_launch(p, st))) return rc;
  }
  return MLSTM_OK;
}

}  // namespace mlstm

// Producer of the cell's operands: depthwise 3x3 conv + SiLU + the three block-diagonal projections of a ViL layer
// in one kernel.
//
// Reference (ViLLayer.forward, vision_lstm2.py:482-491), four separate round trips over (B,S,inner):
//     x_conv = SequenceConv2d(x_mlstm)     vision_lstm_util.py:96-129   depthwise 3x3 on the H x W token grid, zero padded
//     c      = silu(x_conv)
//     q, k   = LinearHeadwiseExpand(c)     vision_lstm2.py:987-1022     y[h, o] = sum_d x[h, d] W[h, o, d] + b
//     v      = LinearHeadwiseExpand(x_mlstm)
//
// One persistent CTA per SM walks (128-token tile, projection block) pairs of ONE block h (its three d x d weight
// matrices stay in shared memory for the whole launch):
//   TMA    the x tile [128 + 2 HALO rows][d] (halo = one grid row + 1 on each side, rounded up to 8 rows so the centre
//          rows are a legal MMA operand; rows outside the sequence arrive as zeros)
//   SIMT   conv + bias + SiLU from the staged rows (column wrap masked by grid position) -> bf16 tile c, K-major
//   TMA    store of c
//   MMA    q = c Wq^T, k = c Wk^T, v = x_centre Wv^T       (tcgen05, fp32 accumulators in TMEM)
//   SIMT   + bias -> bf16 -> staged in the c tile, one output at a time -> TMA stores (the next tile's x rows stream in meanwhile)
// x is read once (plus L2-resident halo rows), c, q, k, v are written once: 5 tensor passes instead of 10.
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace mlstm {
namespace {

using namespace tc;

constexpr int QK_NT = CT + 32;        // 16 compute warps + the control warp

// Developer aid (-DQKV_TIMELINE): CTA 0 keeps clock64() stamps of the phases of its first tiles (thread 0 and the control lane) in
// shared memory and dumps them over the first bytes of v at the end (tests/gpu_tools/timeline_qkv.py).
#ifdef QKV_TIMELINE
#define QTL(k) do { if (blockIdx.x == 0 && n < 8 && (threadIdx.x == 0 || threadIdx.x == CT)) \
    qtl[(threadIdx.x == 0 ? 0 : 64) + n * 8 + (k)] = clock64(); } while (0)
#else
#define QTL(k) do { } while (0)
#endif
constexpr int RMAX = 304;             // staged rows: 128 + 2 * HALO, HALO <= 88 (grid width <= 80)

struct QkvMaps { CUtensorMap x, w[3], out[4]; };   // out: c, q, k, v

template <int DBLK>
struct SmemQ {
  static constexpr int KT = DBLK / 64;
  alignas(1024) uint8_t xh[KT][RMAX * 128];       // x rows with halo, 128-byte swizzled lines of 64 channels
  alignas(1024) uint8_t xc[KT * TILE];            // c tile: A operand of q, k and source of the c store
  alignas(1024) uint8_t w[3][KT * DBLK * 128];    // Wq, Wk, Wv of this CTA's block: [out][in] = K-major B operands
  alignas(16) float cw[9][DBLK];                  // conv taps of the block's channels, tap-major
  alignas(16) float cb[DBLK];
  alignas(16) float pb[3][DBLK];                  // projection biases
  uint64_t bar_w, bar_x, bar_mma;
  uint32_t tmem_base;
};

template <bool FP16>
__device__ __forceinline__ void cvt8(const uint4& w, float (&x)[8]) {
  if (FP16) {
    const __half2* h = reinterpret_cast<const __half2*>(&w);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f2 = __half22float2(h[e]); x[2 * e] = f2.x; x[2 * e + 1] = f2.y; }
  } else {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f2 = __bfloat1622float2(h[e]); x[2 * e] = f2.x; x[2 * e + 1] = f2.y; }
  }
}

template <int DBLK, bool FP16>
__global__ void __launch_bounds__(QK_NT, 1) qkv_fwd_kernel(const __grid_constant__ QkvMaps maps, const mlstm_qkv_params p,
                                                           const int halo, const int tiles_per_batch) {
  constexpr int KT = DBLK / 64;
  constexpr int CH = DBLK / 8;                    // 16-byte channel groups per row
  constexpr int RPT = 128 * CH / CT;              // conv row-items per thread
  constexpr int NB = DBLK / 32;                   // 32-column blocks of an output tile
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemQ<DBLK>& sm = *reinterpret_cast<SmemQ<DBLK>*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool compute = tid < CT, issuer = tid == CT;
  const int rg = warp & 3, cq = compute ? (warp >> 2) : NB, row = rg * 32 + lane;
  const int S = p.GH * p.GW, GW = p.GW, GH = p.GH;
  const int NHB = p.NH;
  const int hb = blockIdx.x % NHB;                           // this CTA's projection block
  const int n_tiles = p.B * tiles_per_batch;
  const int tile0 = blockIdx.x / NHB, tstep = gridDim.x / NHB;
  const int R = 128 + 2 * halo, RB = R / 2;                  // staged rows, rows per TMA box

  if (issuer) {
    tma_prefetch_desc(&maps.x);
    for (int j = 0; j < 3; ++j) tma_prefetch_desc(&maps.w[j]);
    for (int j = 0; j < 4; ++j) tma_prefetch_desc(&maps.out[j]);
    mbar_init(&sm.bar_w, 1); mbar_init(&sm.bar_x, 1); mbar_init(&sm.bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, 512);
  // conv taps (rotated by 180 degrees for the bottom-right layer), biases
  for (int e = tid; e < 9 * DBLK; e += QK_NT) {
    const int tap = e / DBLK, ch = e % DBLK;
    sm.cw[tap][ch] = p.conv_w[(size_t)(hb * DBLK + ch) * 9 + (p.rotate ? 8 - tap : tap)];
  }
  for (int e = tid; e < DBLK; e += QK_NT) {
    sm.cb[e] = p.conv_b ? p.conv_b[hb * DBLK + e] : 0.f;
    sm.pb[0][e] = p.bq ? p.bq[hb * DBLK + e] : 0.f;
    sm.pb[1][e] = p.bk ? p.bk[hb * DBLK + e] : 0.f;
    sm.pb[2][e] = p.bv ? p.bv[hb * DBLK + e] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = sm.tmem_base;
  const uint32_t lane_sel = (uint32_t)(rg * 32) << 16;

  auto tile_coords = [&](int t, int& b, int& tok0) { b = t / tiles_per_batch; tok0 = (t % tiles_per_batch) * 128; };
  auto load_x = [&](int t) {
    int b, tok0; tile_coords(t, b, tok0);
    mbar_arrive_expect_tx(&sm.bar_x, KT * R * 128);
    for (int kt = 0; kt < KT; ++kt)
      for (int hbx = 0; hbx < 2; ++hbx)
        tma_load_4d(sm.xh[kt] + hbx * RB * 128, &maps.x, &sm.bar_x, hb * DBLK + kt * 64, tok0 - halo + hbx * RB, b, 0);
  };
  auto prefetch_x = [&](int t) {
    int b, tok0; tile_coords(t, b, tok0);
    for (int kt = 0; kt < KT; ++kt)
      for (int hbx = 0; hbx < 2; ++hbx) tma_prefetch_4d(&maps.x, hb * DBLK + kt * 64, tok0 - halo + hbx * RB, b, 0);
  };
  if (issuer && tile0 < n_tiles) {
    mbar_arrive_expect_tx(&sm.bar_w, 3 * KT * DBLK * 128);
    for (int j = 0; j < 3; ++j)
      for (int kt = 0; kt < KT; ++kt) tma_load_2d(sm.w[j] + kt * DBLK * 128, &maps.w[j], &sm.bar_w, kt * 64, hb * DBLK);
    load_x(tile0);
  }
  // instruction descriptors: q, k are bf16 x bf16; v is x_dtype x x_dtype
  constexpr uint32_t idQK = make_idesc_bf16(128, DBLK, 0, 0);
  constexpr uint32_t idV = FP16 ? ((1u << 4) | ((uint32_t)(DBLK >> 3) << 17) | ((uint32_t)(128 >> 4) << 24)) : idQK;

#ifdef QKV_TIMELINE
  __shared__ long long qtl[128];
#endif
  int n = 0;
  for (int t = tile0; t < n_tiles; t += tstep, ++n) {
    const uint32_t ph = n & 1;
    int b, tok0; tile_coords(t, b, tok0);
    const bool has_next = t + tstep < n_tiles;
    if (issuer && has_next) prefetch_x(t + tstep);   // towards L2: the next tile's load has to wait for this tile's MMAs
    QTL(0);
    mbar_wait(&sm.bar_x, ph);
    QTL(1);

    // ---- conv + bias + SiLU: thread = 8 channels x RPT consecutive tokens.  Per grid row dy the RPT + 2 staged rows
    //      t0 - 1 .. t0 + RPT are read once each and feed the three taps dx they belong to (sliding window). ----------------
    if (compute) {
      const int c8 = tid % CH, ch0 = c8 * 8, kt = ch0 >> 6, cc = ch0 & 63;
      const int r0 = (tid / CH) * RPT;
      float acc[RPT][8];
      uint32_t ok[RPT];                 // bit (dy+1)*3 + (dx+1): the tap lies inside the grid for this token
      {
        const float4 b0 = *reinterpret_cast<const float4*>(&sm.cb[ch0]), b1 = *reinterpret_cast<const float4*>(&sm.cb[ch0 + 4]);
#pragma unroll
        for (int it = 0; it < RPT; ++it) {
          const int tok = tok0 + r0 + it;
          const int gy = tok / GW, gx = tok - gy * GW;
          const uint32_t my = (gy > 0 ? 1u : 0u) | 2u | (gy + 1 < GH ? 4u : 0u);       // dy = -1, 0, +1
          const uint32_t mx = (gx > 0 ? 1u : 0u) | 2u | (gx + 1 < GW ? 4u : 0u);
          ok[it] = (gy < GH) ? (((my & 1u) ? mx : 0u) | ((my & 2u) ? mx << 3 : 0u) | ((my & 4u) ? mx << 6 : 0u)) : 0u;
          acc[it][0] = b0.x; acc[it][1] = b0.y; acc[it][2] = b0.z; acc[it][3] = b0.w;
          acc[it][4] = b1.x; acc[it][5] = b1.y; acc[it][6] = b1.z; acc[it][7] = b1.w;
        }
      }
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        float wt[3][8];
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float4 w0 = *reinterpret_cast<const float4*>(&sm.cw[(dy + 1) * 3 + dx][ch0]);
          const float4 w1 = *reinterpret_cast<const float4*>(&sm.cw[(dy + 1) * 3 + dx][ch0 + 4]);
          wt[dx][0] = w0.x; wt[dx][1] = w0.y; wt[dx][2] = w0.z; wt[dx][3] = w0.w;
          wt[dx][4] = w1.x; wt[dx][5] = w1.y; wt[dx][6] = w1.z; wt[dx][7] = w1.w;
        }
#pragma unroll
        for (int j = -1; j <= RPT; ++j) {
          const uint4 w = *reinterpret_cast<const uint4*>(sm.xh[kt] + swz128(halo + r0 + dy * GW + j, cc));
          float xv[8];
          cvt8<FP16>(w, xv);
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const int it = j - dx;     // the token this staged row is the dx-neighbour of
            if (it < 0 || it >= RPT) continue;
            if (!((ok[it] >> ((dy + 1) * 3 + (dx + 1))) & 1u)) continue;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[it][e] = fmaf(xv[e], wt[dx + 1][e], acc[it][e]);
          }
        }
      }
#pragma unroll
      for (int it = 0; it < RPT; ++it) {
        float* a_ = acc[it];
        if (p.sp) {   // training: silu'(u) = sg (1 + u (1 - sg)) for the conv's backward, 16 bytes per thread and row
          float d_[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float sg = __fdividef(1.f, 1.f + __expf(-a_[e]));
            d_[e] = sg * fmaf(a_[e], 1.f - sg, 1.f);
            a_[e] *= sg;
          }
          if (tok0 + r0 + it < S)
            *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.sp) + ((size_t)b * S + tok0 + r0 + it) * p.D + hb * DBLK + ch0) =
                make_uint4(pack_bf16x2(d_[0], d_[1]), pack_bf16x2(d_[2], d_[3]), pack_bf16x2(d_[4], d_[5]), pack_bf16x2(d_[6], d_[7]));
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) a_[e] = __fdividef(a_[e], 1.f + __expf(-a_[e]));   // silu
        }
        *reinterpret_cast<uint4*>(sm.xc + kt * TILE + swz128(r0 + it, cc)) =
            make_uint4(pack_bf16x2(a_[0], a_[1]), pack_bf16x2(a_[2], a_[3]), pack_bf16x2(a_[4], a_[5]), pack_bf16x2(a_[6], a_[7]));
      }
    }
    QTL(2);
    fence_proxy_async_smem();
    tc_fence_before();
    named_sync(2, QK_NT);
    QTL(3);

    // ---- c leaves; q = c Wq^T, k = c Wk^T, v = x Wv^T ----------------------------------------------------------
    if (issuer) {
      for (int kt = 0; kt < KT; ++kt) tma_store_4d(&maps.out[0], sm.xc + kt * TILE, hb * DBLK + kt * 64, tok0, b, 0);
      tma_store_commit();
      if (n == 0) mbar_wait(&sm.bar_w, 0);
      tc_fence_after();
      const uint64_t dC = make_sdesc(smem_u32(sm.xc), 16, 1024);
      const uint64_t dX = make_sdesc(smem_u32(sm.xh[0]) + halo * 128, 16, 1024);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const uint64_t dW = make_sdesc(smem_u32(sm.w[j]), 16, 1024);
#pragma unroll
        for (int ks = 0; ks < DBLK / 16; ++ks)
          umma_bf16_ss(tm + j * DBLK, (j < 2 ? dC + kstep(ks) : dX + kstep(ks, RMAX * 128)), dW + kstep(ks, DBLK * 128),
                       j < 2 ? idQK : idV, ks > 0);
      }
      umma_commit(&sm.bar_mma);
    }
    QTL(4);
    mbar_wait(&sm.bar_mma, ph);
    tc_fence_after();
    QTL(5);
    if (issuer) {
      tma_store_wait_read<0>();   // c has left its tile: it becomes the staging tile of q, k, v in turn
      if (has_next) load_x(t + tstep);   // the x rows are dead: the next tile streams in under the epilogue
    }

    // ---- epilogue: + bias -> bf16 -> staged in the c tile (conflict-free swizzled 16-byte stores) -> one TMA store per output ----
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      named_sync(3, QK_NT);   // the staging tile is free (c / the previous output has been read by the TMA engine)
      if (cq < NB) {
        float a[32];
        tmem_ld32(tm + j * DBLK + lane_sel + cq * 32, a);
        tmem_ld_wait();
#pragma unroll
        for (int x = 0; x < 32; x += 8) {
          const int col = cq * 32 + x;
          const float4 b0 = *reinterpret_cast<const float4*>(&sm.pb[j][col]), b1 = *reinterpret_cast<const float4*>(&sm.pb[j][col + 4]);
          *reinterpret_cast<uint4*>(sm.xc + (col >> 6) * TILE + swz128(row, col & 63)) =
              make_uint4(pack_bf16x2(a[x] + b0.x, a[x + 1] + b0.y), pack_bf16x2(a[x + 2] + b0.z, a[x + 3] + b0.w),
                         pack_bf16x2(a[x + 4] + b1.x, a[x + 5] + b1.y), pack_bf16x2(a[x + 6] + b1.z, a[x + 7] + b1.w));
        }
      }
      fence_proxy_async_smem();
      named_sync(4, QK_NT);
      if (issuer) {
        for (int kt = 0; kt < KT; ++kt) tma_store_4d(&maps.out[1 + j], sm.xc + kt * TILE, hb * DBLK + kt * 64, tok0, b, 0);
        tma_store_commit();
        tma_store_wait_read<0>();
      }
    }
    QTL(6);
    tc_fence_before();
    named_sync(2, QK_NT);   // accumulators consumed: the next tile's MMAs may overwrite them
  }
  if (issuer) tma_store_wait_all<0>();
  tc_fence_before();
  __syncthreads();
#ifdef QKV_TIMELINE
  if (blockIdx.x == 0 && tid < 128) reinterpret_cast<long long*>(p.v)[tid] = qtl[tid];
#endif
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ---- column sums of up to three (T, D) bf16 matrices: partials[range][j][D], then a fixed-order reduction ----------------
constexpr int CS_NT = 256, CS_RANGES = 296;      // 2 CTAs per SM per 256-column slab and source
struct ColsumArgs { const __nv_bfloat16* src[3]; };

__global__ void __launch_bounds__(CS_NT) colsum_kernel(const ColsumArgs a, const int T, const int D, const int64_t ld,
                                                       float* __restrict__ partials, const int n_src) {
  __shared__ float red[8][256];
  const int j = blockIdx.z, cg = threadIdx.x & 31, rl = threadIdx.x >> 5;   // 32 groups of 8 columns x 8 row lanes
  const int col = blockIdx.y * 256 + cg * 8;
  const int per = (T + gridDim.x - 1) / gridDim.x;
  const int t0 = blockIdx.x * per, t1 = min(T, t0 + per);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  if (col < D) {
    const __nv_bfloat16* base = a.src[j] + col;
    for (int t = t0 + rl; t < t1; t += 8) {
      const uint4 w = *reinterpret_cast<const uint4*>(base + (int64_t)t * ld);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f2 = __bfloat1622float2(h[e]); acc[2 * e] += f2.x; acc[2 * e + 1] += f2.y; }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) red[rl][cg * 8 + e] = acc[e];
  __syncthreads();
  const int c = threadIdx.x;   // one column per thread, row lanes summed in fixed order
  if (blockIdx.y * 256 + c < D) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += red[r][c];
    partials[((size_t)blockIdx.x * n_src + j) * D + blockIdx.y * 256 + c] = s;
  }
}
__global__ void colsum_reduce_kernel(const float* __restrict__ partials, float* __restrict__ out, const int n, const int ranges) {
  const int e0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;   // eight lanes per output element
  const int e = min(e0, n - 1);
  const float s = ordered_sum8(partials + e, ranges, (size_t)n);
  if (e0 < n && (threadIdx.x & 7) == 0) out[e] = s;
}

// 3-D map (channels, tokens, batch) of a (B*S, D) bf16/fp16 matrix with row stride ld; box = 64 x box_rows x 1.  Encoded as a 4-D
// map with a unit outer dimension so the kernels use the one 4-D load / store wrapper.
int make_rows_tmap(CUtensorMap* out, const void* ptr, int D, int S, int B, int64_t ld, int box_rows) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.ptr = ptr; key.s0 = ld; key.d0 = D; key.d1 = S; key.d2 = B; key.box_rows = box_rows; key.kind = 7;
  if (tmap_cache_lookup(key, out, false)) return 0;
  PFN_tmapEncodeTiled enc = get_tmap_encoder();
  if (!enc) return -1;
  cuuint64_t dims[4] = {(cuuint64_t)D, (cuuint64_t)S, (cuuint64_t)B, 1};
  cuuint64_t strides[3] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * S, (cuuint64_t)ld * 2 * S * B};
  cuuint32_t box[4] = {64u, (cuuint32_t)box_rows, 1u, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  const int r = (int)enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r == 0) tmap_cache_lookup(key, out, true);
  return r;
}

bool qkv_shape_ok(int D, int NH, int GH, int GW, int64_t ld_x) {
  if (D <= 0 || NH <= 0 || D % NH != 0 || GH <= 0 || GW <= 0) return false;
  const int d = D / NH;
  return (d == 64 || d == 128) && GW <= 80 && ld_x >= D && ld_x % 8 == 0;
}

template <int DBLK, bool FP16>
int launch_qkv(const mlstm_qkv_params& p, const QkvMaps& maps, int halo, cudaStream_t st) {
  const size_t smem = sizeof(SmemQ<DBLK>);
  cudaError_t e = set_max_smem_once(reinterpret_cast<const void*>(qkv_fwd_kernel<DBLK, FP16>), smem);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(qkv_fwd, %zu B): %s", smem, cudaGetErrorString(e));
    return MLSTM_ERR_CUDA;
  }
  const int S = p.GH * p.GW, tiles_per_batch = (S + 127) / 128;
  const int sms = sm_count_of(p.x);
  int per_block = sms / p.NH;                       // CTAs per projection block
  if (per_block < 1) per_block = 1;
  if (per_block > p.B * tiles_per_batch) per_block = p.B * tiles_per_batch;
  qkv_fwd_kernel<DBLK, FP16><<<dim3(per_block * p.NH), dim3(QK_NT), smem, st>>>(maps, p, halo, tiles_per_batch);
  count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("qkv_fwd launch failed: %s", cudaGetErrorString(e));
    return MLSTM_ERR_CUDA;
  }
  return MLSTM_OK;
}

}  // namespace
}  // namespace mlstm

using namespace mlstm;

extern "C" {

int mlstm_b200_qkv_supported(int D, int NH, int GH, int GW, int64_t ld_x) { return qkv_shape_ok(D, NH, GH, GW, ld_x) ? 1 : 0; }

size_t mlstm_b200_colsum_workspace_bytes(int D, int n_src) {
  return (D > 0 && n_src > 0) ? sizeof(float) * (size_t)CS_RANGES * (size_t)n_src * (size_t)D : 0;
}

int mlstm_b200_colsum(const void* const* src, int n_src, int T, int D, int64_t ld, float* out, void* workspace,
                      size_t workspace_bytes, void* cuda_stream) {
  clear_error();
  if (!src || n_src < 1 || n_src > 3 || T < 0 || D < 8 || D % 8 != 0 || ld % 8 != 0 || ld < D || !out) {
    set_error("colsum: needs 1..3 sources, D and ld multiples of 8, ld >= D (n_src=%d T=%d D=%d ld=%lld)", n_src, T, D, (long long)ld);
    return MLSTM_ERR_INVALID_ARG;
  }
  ColsumArgs a{};
  for (int j = 0; j < n_src; ++j) {
    if (!src[j] || ((uintptr_t)src[j] & 15u)) { set_error("colsum: source %d null or not 16-byte aligned", j); return MLSTM_ERR_INVALID_ARG; }
    a.src[j] = reinterpret_cast<const __nv_bfloat16*>(src[j]);
  }
  if (!workspace || workspace_bytes < mlstm_b200_colsum_workspace_bytes(D, n_src)) {
    set_error("colsum: workspace too small (%zu < %zu)", workspace ? workspace_bytes : (size_t)0, mlstm_b200_colsum_workspace_bytes(D, n_src));
    return MLSTM_ERR_WORKSPACE;
  }
  int rc;
  if ((rc = bind_device(src[0]))) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  const int n = n_src * D;
  if (T == 0) {
    cudaMemsetAsync(out, 0, sizeof(float) * n, st);
    return MLSTM_OK;
  }
  int ranges = (T + 63) / 64;
  if (ranges > CS_RANGES) ranges = CS_RANGES;
  float* partials = reinterpret_cast<float*>(workspace);
  colsum_kernel<<<dim3(ranges, (D + 255) / 256, n_src), CS_NT, 0, st>>>(a, T, D, ld, partials, n_src);
  count_launch();
  colsum_reduce_kernel<<<(n * 8 + 255) / 256, 256, 0, st>>>(partials, out, n, ranges);
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("colsum launch failed: %s", cudaGetErrorString(e)); return MLSTM_ERR_CUDA; }
  return MLSTM_OK;
}

int mlstm_b200_qkv_fwd(const mlstm_qkv_params* p, void* cuda_stream) {
  clear_error();
  if (!p) { set_error("params is NULL"); return MLSTM_ERR_INVALID_ARG; }
  if (p->abi_version != MLSTM_B200_ABI_VERSION) {
    set_error("abi_version %d != %d", p->abi_version, MLSTM_B200_ABI_VERSION);
    return MLSTM_ERR_INVALID_ARG;
  }
  if (p->B < 0 || p->x_dtype < 0 || p->x_dtype > 1) { set_error("bad B / x_dtype"); return MLSTM_ERR_INVALID_ARG; }
  if (!qkv_shape_ok(p->D, p->NH, p->GH, p->GW, p->ld_x)) {
    set_error("qkv producer: D / NH must be 64 or 128, grid width <= 80, ld_x a multiple of 8 and >= D (D=%d NH=%d GH=%d GW=%d ld_x=%lld)",
              p->D, p->NH, p->GH, p->GW, (long long)p->ld_x);
    return MLSTM_ERR_UNSUPPORTED;
  }
  if (p->B == 0) return MLSTM_OK;
  if (!p->x || !p->conv_w || !p->wq || !p->wk || !p->wv || !p->c || !p->q || !p->k || !p->v) {
    set_error("qkv producer: null pointer");
    return MLSTM_ERR_INVALID_ARG;
  }
  if (((uintptr_t)p->x | (uintptr_t)p->wq | (uintptr_t)p->wk | (uintptr_t)p->wv | (uintptr_t)p->c | (uintptr_t)p->q | (uintptr_t)p->k |
       (uintptr_t)p->v | (uintptr_t)p->sp) & 15u) {
    set_error("qkv producer: x, weights and outputs must be 16-byte aligned (TMA)");
    return MLSTM_ERR_INVALID_ARG;
  }
  int rc;
  if ((rc = bind_device(p->x))) return rc;
  const int S = p->GH * p->GW, d = p->D / p->NH;
  const int halo = ((p->GW + 1 + 7) / 8) * 8;
  QkvMaps maps;
  int r = 0;
  r |= make_rows_tmap(&maps.x, p->x, p->D, S, p->B, p->ld_x, (128 + 2 * halo) / 2);
  const void* ws[3] = {p->wq, p->wk, p->wv};
  for (int j = 0; j < 3; ++j) r |= tc::make_state_tmap(&maps.w[j], ws[j], (size_t)p->D, d, d);
  void* outs[4] = {p->c, p->q, p->k, p->v};
  for (int j = 0; j < 4; ++j) r |= make_rows_tmap(&maps.out[j], outs[j], p->D, S, p->B, p->D, 128);
  if (r) {
    set_error("cuTensorMapEncodeTiled failed (%d): pointers must be 16-byte aligned", r);
    return r == -1 ? MLSTM_ERR_NO_DEVICE : MLSTM_ERR_INVALID_ARG;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  if (d == 128) return p->x_dtype ? launch_qkv<128, true>(*p, maps, halo, st) : launch_qkv<128, false>(*p, maps, halo, st);
  return p->x_dtype ? launch_qkv<64, true>(*p, maps, halo, st) : launch_qkv<64, false>(*p, maps, halo, st);
}

}  // extern "C"

// =====================================================================================================================
// Backward of the three projections: dxc = dc + dq Wq + dk Wk, dxv = dv Wv, dW* = dy^T in, db = column sums.
// Same tiling as the forward (a persistent CTA per SM walks 128-row tiles of ONE block h, weights resident).  The weight
// gradients accumulate in TMEM over all tiles of the CTA (three d x d fp32 accumulators) and leave once, as per-CTA partials
// reduced in fixed order by a second kernel.  Two MMA phases per tile share the activation buffers:
//   phase 1   acc  = dq Wq + dk Wk ;  dWq += dq^T c ;  dWk += dk^T c        buffers: dq | dk | c | stage <- dc
//   phase 2   acc  = dv Wv ;          dWv += dv^T x                          buffers: dv (was dq) | x (was c)
// =====================================================================================================================
namespace mlstm {
namespace {

struct QkvBwdMaps { CUtensorMap x, c, dq, dk, dv, dc, w[3], dxc, dxv; };

template <int DBLK>
struct SmemQB {
  static constexpr int KT = DBLK / 64;
  alignas(1024) uint8_t w[3][KT * DBLK * 128];
  alignas(1024) uint8_t a0[KT * TILE];            // dq, then dv
  alignas(1024) uint8_t a1[KT * TILE];            // dk
  alignas(1024) uint8_t a2[KT * TILE];            // c, then x
  alignas(1024) uint8_t stage[KT * TILE];         // dc on the way in, dxc / dxv on the way out
  uint64_t bar_w, bar_l1, bar_l2, bar_m1, bar_m2;
  uint32_t tmem_base;
};

// column partial sums of a [128][DBLK] swizzled bf16 tile: thread = (8 columns, row lane of 32)
template <int DBLK>
__device__ __forceinline__ void tile_colsum(const uint8_t* tile, int tid, float (&acc)[8]) {
  constexpr int CH = DBLK / 8;
  const int c8 = tid % CH, ch0 = c8 * 8, kt = ch0 >> 6, cc = ch0 & 63;
#pragma unroll
  for (int r = tid / CH; r < 128; r += CT / CH) {
    const uint4 w = *reinterpret_cast<const uint4*>(tile + kt * TILE + swz128(r, cc));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
    for (int e = 0; e < 4; ++e) { const float2 f2 = __bfloat1622float2(h[e]); acc[2 * e] += f2.x; acc[2 * e + 1] += f2.y; }
  }
}

template <int DBLK, bool FP16>
__global__ void __launch_bounds__(QK_NT, 1) qkv_bwd_kernel(const __grid_constant__ QkvBwdMaps maps, const mlstm_qkv_bwd_params p,
                                                           const int n_tiles, float* __restrict__ ws) {
  constexpr int KT = DBLK / 64;
  constexpr int NB = DBLK / 32;
  constexpr int CH = DBLK / 8;
  constexpr uint32_t A_LBO = (DBLK == 128) ? TILE : 0;    // d = 64: the second 64-row M block of a d x d product aliases the first
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemQB<DBLK>& sm = *reinterpret_cast<SmemQB<DBLK>*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool compute = tid < CT, issuer = tid == CT;
  const int rg = warp & 3, cq = compute ? (warp >> 2) : NB, row = rg * 32 + lane;
  const int hb = blockIdx.x % p.NH;
  const int tile0 = blockIdx.x / p.NH, tstep = gridDim.x / p.NH;
  const bool has_dc = p.dc != nullptr;

  if (issuer) {
    mbar_init(&sm.bar_w, 1); mbar_init(&sm.bar_l1, 1); mbar_init(&sm.bar_l2, 1); mbar_init(&sm.bar_m1, 1); mbar_init(&sm.bar_m2, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(&sm.tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = sm.tmem_base, tAcc = tm + 3 * DBLK;
  const uint32_t lane_sel = (uint32_t)(rg * 32) << 16;

  auto load_tile = [&](uint8_t* dst, const CUtensorMap* map, uint64_t* bar, int t) {
    for (int kt = 0; kt < KT; ++kt) tma_load_2d(dst + kt * TILE, map, bar, hb * DBLK + kt * 64, t * 128);
  };
  auto load_phase1 = [&](int t) {
    mbar_arrive_expect_tx(&sm.bar_l1, (has_dc ? 4 : 3) * KT * TILE);
    load_tile(sm.a0, &maps.dq, &sm.bar_l1, t);
    load_tile(sm.a1, &maps.dk, &sm.bar_l1, t);
    load_tile(sm.a2, &maps.c, &sm.bar_l1, t);
    if (has_dc) load_tile(sm.stage, &maps.dc, &sm.bar_l1, t);
  };
  if (issuer && tile0 < n_tiles) {
    mbar_arrive_expect_tx(&sm.bar_w, 3 * KT * DBLK * 128);
    for (int j = 0; j < 3; ++j)
      for (int kt = 0; kt < KT; ++kt) tma_load_2d(sm.w[j] + kt * DBLK * 128, &maps.w[j], &sm.bar_w, kt * 64, hb * DBLK);
    load_phase1(tile0);
  }
  // dy W : A = dy tile, K-major (K = out); B = W [K = out rows][N = in contiguous], MN-major
  constexpr uint32_t idX = make_idesc_bf16(128, DBLK, 0, 1);
  // dy^T in : A = dy tile, MN-major (M = out); B = in tile, MN-major (N = in); K = the tile's 128 rows
  constexpr uint32_t idW = make_idesc_bf16(128, DBLK, 1, 1);
  auto mma_dx = [&](const uint8_t* a, const uint8_t* w, uint32_t acc_first) {
    const uint64_t dA = make_sdesc(smem_u32(a), 16, 1024), dB = make_sdesc(smem_u32(w), DBLK * 128, 1024);
#pragma unroll
    for (int ks = 0; ks < DBLK / 16; ++ks) umma_bf16_ss(tAcc, dA + kstep(ks), dB + mnstep(ks), idX, ks > 0 ? 1u : acc_first);
  };
  auto mma_dw = [&](uint32_t d_tmem, const uint8_t* a, const uint8_t* b, uint32_t acc_first) {
    const uint64_t dA = make_sdesc(smem_u32(a), A_LBO, 1024), dB = make_sdesc(smem_u32(b), TILE, 1024);
#pragma unroll
    for (int ks = 0; ks < 128 / 16; ++ks) umma_bf16_ss(d_tmem, dA + mnstep(ks), dB + mnstep(ks), idW, ks > 0 ? 1u : acc_first);
  };
  // + optional dc, -> bf16 -> stage (each thread rewrites exactly the 64 bytes it has just read)
  auto epilogue = [&](bool add_dc) {
    if (cq < NB) {
      float a[32];
      tmem_ld32(tAcc + lane_sel + cq * 32, a);
      tmem_ld_wait();
#pragma unroll
      for (int x = 0; x < 32; x += 8) {
        const int col = cq * 32 + x;
        uint4* slot = reinterpret_cast<uint4*>(sm.stage + (col >> 6) * TILE + swz128(row, col & 63));
        if (add_dc) {
          const uint4 w = *slot;
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&w);
#pragma unroll
          for (int e = 0; e < 4; ++e) { const float2 f2 = __bfloat1622float2(h[e]); a[x + 2 * e] += f2.x; a[x + 2 * e + 1] += f2.y; }
        }
        *slot = make_uint4(pack_bf16x2(a[x], a[x + 1]), pack_bf16x2(a[x + 2], a[x + 3]), pack_bf16x2(a[x + 4], a[x + 5]),
                           pack_bf16x2(a[x + 6], a[x + 7]));
      }
    }
  };

  float dbq[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dbk[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dbv[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int n = 0;
  for (int t = tile0; t < n_tiles; t += tstep, ++n) {
    const uint32_t ph = n & 1;
    const bool has_next = t + tstep < n_tiles;
    const uint32_t accw = n > 0 ? 1u : 0u;
    // ---- phase 1 ----------------------------------------------------------------------------------------------------
    mbar_wait(&sm.bar_l1, ph);
    if (issuer) {
      if (n == 0) mbar_wait(&sm.bar_w, 0);
      tc_fence_after();
      mma_dx(sm.a0, sm.w[0], 0u);
      mma_dx(sm.a1, sm.w[1], 1u);
      mma_dw(tm, sm.a0, sm.a2, accw);
      mma_dw(tm + DBLK, sm.a1, sm.a2, accw);
      umma_commit(&sm.bar_m1);
    }
    if (compute && p.db) { tile_colsum<DBLK>(sm.a0, tid, dbq); tile_colsum<DBLK>(sm.a1, tid, dbk); }
    mbar_wait(&sm.bar_m1, ph);
    tc_fence_after();
    named_sync(2, QK_NT);          // every warp is past its column sums: dq and c may be overwritten
    if (issuer) {
      mbar_arrive_expect_tx(&sm.bar_l2, 2 * KT * TILE);
      load_tile(sm.a0, &maps.dv, &sm.bar_l2, t);
      load_tile(sm.a2, &maps.x, &sm.bar_l2, t);
    }
    epilogue(has_dc);
    fence_proxy_async_smem();
    tc_fence_before();
    named_sync(3, QK_NT);          // dxc staged, acc consumed
    if (issuer) {
      for (int kt = 0; kt < KT; ++kt) tma_store_2d(&maps.dxc, sm.stage + kt * TILE, hb * DBLK + kt * 64, t * 128);
      tma_store_commit();
    }
    // ---- phase 2 ----------------------------------------------------------------------------------------------------
    mbar_wait(&sm.bar_l2, ph);
    if (FP16) {                    // x arrives as fp16: bf16 in place (the MMAs of this kernel are bf16 x bf16)
      if (compute) {
#pragma unroll
        for (int it = 0; it < KT * TILE / 16 / CT; ++it) {
          uint4* q4 = reinterpret_cast<uint4*>(sm.a2) + tid + it * CT;
          uint4 w = *q4;
          const __half2* h = reinterpret_cast<const __half2*>(&w);
          uint32_t o[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) { const float2 f2 = __half22float2(h[e]); o[e] = pack_bf16x2(f2.x, f2.y); }
          *q4 = make_uint4(o[0], o[1], o[2], o[3]);
        }
      }
      fence_proxy_async_smem();
      named_sync(2, QK_NT);
    }
    if (issuer) {
      tc_fence_after();
      mma_dx(sm.a0, sm.w[2], 0u);
      mma_dw(tm + 2 * DBLK, sm.a0, sm.a2, accw);
      umma_commit(&sm.bar_m2);
      tma_store_wait_read<0>();    // dxc has left the staging tile
    }
    if (compute && p.db) tile_colsum<DBLK>(sm.a0, tid, dbv);
    mbar_wait(&sm.bar_m2, ph);
    tc_fence_after();
    named_sync(2, QK_NT);          // staging tile free (the control lane waited for the store), dv and x dead
    if (issuer && has_next) {      // the next tile's dq, dk, c stream in under the second epilogue; its dc after the store below
      mbar_arrive_expect_tx(&sm.bar_l1, (has_dc ? 4 : 3) * KT * TILE);
      load_tile(sm.a0, &maps.dq, &sm.bar_l1, t + tstep);
      load_tile(sm.a1, &maps.dk, &sm.bar_l1, t + tstep);
      load_tile(sm.a2, &maps.c, &sm.bar_l1, t + tstep);
    }
    epilogue(false);
    fence_proxy_async_smem();
    tc_fence_before();
    named_sync(3, QK_NT);
    if (issuer) {
      for (int kt = 0; kt < KT; ++kt) tma_store_2d(&maps.dxv, sm.stage + kt * TILE, hb * DBLK + kt * 64, t * 128);
      tma_store_commit();
      if (has_next) {
        tma_store_wait_read<0>();   // the staging tile is rewritten by the next tile (its dc now, or its first epilogue)
        if (has_dc) load_tile(sm.stage, &maps.dc, &sm.bar_l1, t + tstep);
      }
    }
  }
  if (issuer) tma_store_wait_all<0>();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // ---- weight-gradient partials of this CTA: [3][DBLK][DBLK] fp32, then the bias partials [3][DBLK] -----------------------
  float* wsw = ws + (size_t)blockIdx.x * (3 * DBLK * DBLK + 3 * DBLK);
  if (cq < NB && n > 0) {
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float a[32];
      tmem_ld32(tm + j * DBLK + lane_sel + cq * 32, a);
      tmem_ld_wait();
      if (row < DBLK) {
        float* dst = wsw + ((size_t)j * DBLK + row) * DBLK + cq * 32;
#pragma unroll
        for (int x = 0; x < 32; x += 4) *reinterpret_cast<float4*>(dst + x) = make_float4(a[x], a[x + 1], a[x + 2], a[x + 3]);
      }
    }
  }
  if (p.db) {   // reduce the 32 row lanes in fixed order through shared memory (the activation buffers are dead)
    float* red = reinterpret_cast<float*>(sm.a0);   // [32 row lanes][3][DBLK] floats <= a0 + a1
    if (compute) {
      const int c8 = tid % CH, rl = tid / CH;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        red[(rl * 3 + 0) * DBLK + c8 * 8 + e] = dbq[e];
        red[(rl * 3 + 1) * DBLK + c8 * 8 + e] = dbk[e];
        red[(rl * 3 + 2) * DBLK + c8 * 8 + e] = dbv[e];
      }
    }
    __syncthreads();
    if (tid < 3 * DBLK) {
      float s = 0.f;
      for (int rl = 0; rl < CT / CH; ++rl) s += red[rl * 3 * DBLK + tid];
      wsw[3 * DBLK * DBLK + tid] = s;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// out[j][h][o][i] = sum over the CTAs of block h, in launch order
__global__ void qkv_bwd_reduce_kernel(const float* __restrict__ ws, const mlstm_qkv_bwd_params p, const int per_block, const int d) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int stride = 3 * d * d + 3 * d;
  if (e < 3 * p.NH * d * d) {
    const int j = e / (p.NH * d * d), r = e % (p.NH * d * d), h = r / (d * d), oi = r % (d * d);
    float s = 0.f;
    for (int k = 0; k < per_block; ++k) s += ws[(size_t)(k * p.NH + h) * stride + (size_t)j * d * d + oi];
    (j == 0 ? p.dwq : (j == 1 ? p.dwk : p.dwv))[(size_t)h * d * d + oi] = s;
  }
  if (p.db && e < 3 * p.D) {
    const int j = e / p.D, col = e % p.D, h = col / d, c = col % d;
    float s = 0.f;
    for (int k = 0; k < per_block; ++k) s += ws[(size_t)(k * p.NH + h) * stride + 3 * d * d + j * d + c];
    p.db[e] = s;
  }
}

int qkv_bwd_ctas_per_block(const mlstm_qkv_bwd_params& p) {
  int per_block = 148 / p.NH;    // fixed (not the device's SM count): the workspace size must not depend on the device
  const int n_tiles = (p.T + 127) / 128;
  if (per_block < 1) per_block = 1;
  if (per_block > n_tiles) per_block = n_tiles;
  return per_block;
}

template <int DBLK, bool FP16>
int launch_qkv_bwd(const mlstm_qkv_bwd_params& p, const QkvBwdMaps& maps, cudaStream_t st) {
  const size_t smem = sizeof(SmemQB<DBLK>);
  cudaError_t e = set_max_smem_once(reinterpret_cast<const void*>(qkv_bwd_kernel<DBLK, FP16>), smem);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(qkv_bwd, %zu B): %s", smem, cudaGetErrorString(e));
    return MLSTM_ERR_CUDA;
  }
  const int per_block = qkv_bwd_ctas_per_block(p), n_tiles = (p.T + 127) / 128;
  float* ws = reinterpret_cast<float*>(p.workspace);
  qkv_bwd_kernel<DBLK, FP16><<<dim3(per_block * p.NH), dim3(QK_NT), smem, st>>>(maps, p, n_tiles, ws);
  count_launch();
  const int n = 3 * p.NH * DBLK * DBLK;
  qkv_bwd_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>(ws, p, per_block, DBLK);
  count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("qkv_bwd launch failed: %s", cudaGetErrorString(e));
    return MLSTM_ERR_CUDA;
  }
  return MLSTM_OK;
}

}  // namespace
}  // namespace mlstm

extern "C" {

size_t mlstm_b200_qkv_bwd_workspace_bytes(const mlstm_qkv_bwd_params* p) {
  if (!p || p->T <= 0 || p->NH <= 0 || p->D % p->NH != 0) return 0;
  const size_t d = (size_t)(p->D / p->NH);
  return sizeof(float) * (size_t)qkv_bwd_ctas_per_block(*p) * (size_t)p->NH * (3 * d * d + 3 * d);
}

int mlstm_b200_qkv_bwd(const mlstm_qkv_bwd_params* p, void* cuda_stream) {
  clear_error();
  if (!p) { set_error("params is NULL"); return MLSTM_ERR_INVALID_ARG; }
  if (p->abi_version != MLSTM_B200_ABI_VERSION) {
    set_error("abi_version %d != %d", p->abi_version, MLSTM_B200_ABI_VERSION);
    return MLSTM_ERR_INVALID_ARG;
  }
  if (p->T < 0 || p->x_dtype < 0 || p->x_dtype > 1) { set_error("bad T / x_dtype"); return MLSTM_ERR_INVALID_ARG; }
  if (!qkv_shape_ok(p->D, p->NH, 1, 1, p->ld_x)) {
    set_error("qkv backward: D / NH must be 64 or 128, ld_x a multiple of 8 and >= D (D=%d NH=%d ld_x=%lld)", p->D, p->NH, (long long)p->ld_x);
    return MLSTM_ERR_UNSUPPORTED;
  }
  if (!p->dwq || !p->dwk || !p->dwv) { set_error("qkv backward: null weight-gradient pointer"); return MLSTM_ERR_INVALID_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  const int d = p->D / p->NH;
  if (p->T == 0) {
    int rc0 = bind_device(p->dwq);
    if (rc0) return rc0;
    for (float* w : {p->dwq, p->dwk, p->dwv}) cudaMemsetAsync(w, 0, sizeof(float) * (size_t)p->NH * d * d, st);
    if (p->db) cudaMemsetAsync(p->db, 0, sizeof(float) * 3 * p->D, st);
    return MLSTM_OK;
  }
  if (!p->x || !p->c || !p->dq || !p->dk || !p->dv || !p->wq || !p->wk || !p->wv || !p->dxc || !p->dxv) {
    set_error("qkv backward: null pointer");
    return MLSTM_ERR_INVALID_ARG;
  }
  if (!p->workspace || p->workspace_bytes < mlstm_b200_qkv_bwd_workspace_bytes(p)) {
    set_error("qkv backward: workspace too small (%zu < %zu)", p->workspace ? p->workspace_bytes : (size_t)0,
              mlstm_b200_qkv_bwd_workspace_bytes(p));
    return MLSTM_ERR_WORKSPACE;
  }
  int rc;
  if ((rc = bind_device(p->x))) return rc;
  QkvBwdMaps maps;
  int r = 0;
  r |= make_mat_tmap(&maps.x, p->x, p->D, p->T, p->ld_x);
  r |= make_mat_tmap(&maps.c, p->c, p->D, p->T, p->D);
  r |= make_mat_tmap(&maps.dq, p->dq, p->D, p->T, p->D);
  r |= make_mat_tmap(&maps.dk, p->dk, p->D, p->T, p->D);
  r |= make_mat_tmap(&maps.dv, p->dv, p->D, p->T, p->D);
  maps.dc = maps.c;
  if (p->dc) r |= make_mat_tmap(&maps.dc, p->dc, p->D, p->T, p->D);
  r |= make_mat_tmap(&maps.dxc, p->dxc, p->D, p->T, p->D);
  r |= make_mat_tmap(&maps.dxv, p->dxv, p->D, p->T, p->D);
  const void* wsrc[3] = {p->wq, p->wk, p->wv};
  for (int j = 0; j < 3; ++j) r |= tc::make_state_tmap(&maps.w[j], wsrc[j], (size_t)p->D, d, d);
  if (r) {
    set_error("cuTensorMapEncodeTiled failed (%d): pointers must be 16-byte aligned", r);
    return r == -1 ? MLSTM_ERR_NO_DEVICE : MLSTM_ERR_INVALID_ARG;
  }
  if (d == 128) return p->x_dtype ? launch_qkv_bwd<128, true>(*p, maps, st) : launch_qkv_bwd<128, false>(*p, maps, st);
  return p->x_dtype ? launch_qkv_bwd<64, true>(*p, maps, st) : launch_qkv_bwd<64, false>(*p, maps, st);
}

}  // extern "C"

// =====================================================================================================================
// Backward of the depthwise conv + SiLU: du = dxc * sp ; dx = conv^T(du) + dxv ; dwc, dbc.  SIMT (depthwise: nothing for the
// tensor cores), same tiling and staging as the forward: x and dxc rows of a 128-token tile with a halo of one grid row + 1,
// du formed in place over the staged dxc rows (sp read straight from global, 16 coalesced bytes per thread), dx written
// straight to global.  Weight-gradient accumulators live in registers over all tiles of the CTA (thread = 4 channels x rows).
// =====================================================================================================================
namespace mlstm {
namespace {

struct ConvBwdMaps { CUtensorMap x, dxc; };

template <int DBLK>
struct SmemCB {
  static constexpr int KT = DBLK / 64;
  alignas(1024) uint8_t xh[KT][RMAX * 128];       // x rows with halo
  alignas(1024) uint8_t dh[KT][RMAX * 128];       // dxc rows with halo -> du (bf16) in place
  alignas(16) float cw[9][DBLK];                  // taps as the forward applied them (rotated for the bottom-right layer)
  uint64_t bar_x, bar_d;
};

template <int DBLK, bool FP16>
__global__ void __launch_bounds__(CT, 1) conv_bwd_kernel(const __grid_constant__ ConvBwdMaps maps, const mlstm_conv_bwd_params p,
                                                         const int halo, const int tiles_per_batch, float* __restrict__ ws) {
  constexpr int KT = DBLK / 64;
  constexpr int CH = DBLK / 8;                    // 16-byte channel groups per row (steps A, B)
  constexpr int RPT = 128 * CH / CT;              // rows per thread in step B (consecutive tokens)
  constexpr int CG = DBLK / 4;                    // 8-byte channel groups per row (step C)
  constexpr int RL = CT / CG;                     // row lanes in step C
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  SmemCB<DBLK>& sm = *reinterpret_cast<SmemCB<DBLK>*>(smem_raw);
  if ((smem_u32(smem_raw) & 1023u) != 0) __trap();

  const int tid = threadIdx.x;
  const int S = p.GH * p.GW, GW = p.GW, GH = p.GH;
  const int hb = blockIdx.x % p.NH;
  const int n_tiles = p.B * tiles_per_batch;
  const int tile0 = blockIdx.x / p.NH, tstep = gridDim.x / p.NH;
  const int R = 128 + 2 * halo, RB = R / 2;

  if (tid == 0) {
    tma_prefetch_desc(&maps.x); tma_prefetch_desc(&maps.dxc);
    mbar_init(&sm.bar_x, 1); mbar_init(&sm.bar_d, 1);
    fence_mbar_init();
  }
  for (int e = tid; e < 9 * DBLK; e += CT) {
    const int tap = e / DBLK, ch = e % DBLK;
    sm.cw[tap][ch] = p.conv_w[(size_t)(hb * DBLK + ch) * 9 + (p.rotate ? 8 - tap : tap)];
  }
  __syncthreads();

  auto tile_coords = [&](int t, int& b, int& tok0) { b = t / tiles_per_batch; tok0 = (t % tiles_per_batch) * 128; };
  auto load_rows = [&](uint8_t* dst, const CUtensorMap* map, uint64_t* bar, int t) {   // dst: [KT][RMAX * 128]
    int b, tok0; tile_coords(t, b, tok0);
    mbar_arrive_expect_tx(bar, KT * R * 128);
    for (int kt = 0; kt < KT; ++kt)
      for (int hbx = 0; hbx < 2; ++hbx)
        tma_load_4d(dst + kt * (RMAX * 128) + hbx * RB * 128, map, bar, hb * DBLK + kt * 64, tok0 - halo + hbx * RB, b, 0);
  };
  auto prefetch_tile = [&](int t) {
    int b, tok0; tile_coords(t, b, tok0);
    for (int kt = 0; kt < KT; ++kt)
      for (int hbx = 0; hbx < 2; ++hbx) {
        tma_prefetch_4d(&maps.x, hb * DBLK + kt * 64, tok0 - halo + hbx * RB, b, 0);
        tma_prefetch_4d(&maps.dxc, hb * DBLK + kt * 64, tok0 - halo + hbx * RB, b, 0);
      }
  };
  // bit (dy+1)*3 + (dx+1): token (gy, gx) has a neighbour at (gy+dy, gx+dx) inside the grid
  auto tap_mask = [&](int tok) -> uint32_t {
    const int gy = tok / GW, gx = tok - gy * GW;
    if (gy >= GH) return 0u;
    const uint32_t my = (gy > 0 ? 1u : 0u) | 2u | (gy + 1 < GH ? 4u : 0u);
    const uint32_t mx = (gx > 0 ? 1u : 0u) | 2u | (gx + 1 < GW ? 4u : 0u);
    return ((my & 1u) ? mx : 0u) | ((my & 2u) ? mx << 3 : 0u) | ((my & 4u) ? mx << 6 : 0u);
  };
  if (tid == 0 && tile0 < n_tiles) {
    load_rows(&sm.dh[0][0], &maps.dxc, &sm.bar_d, tile0);
    load_rows(&sm.xh[0][0], &maps.x, &sm.bar_x, tile0);
  }

  float wacc[9][4], bacc[4];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int e = 0; e < 4; ++e) wacc[k][e] = 0.f;
#pragma unroll
  for (int e = 0; e < 4; ++e) bacc[e] = 0.f;

  int n = 0;
  for (int t = tile0; t < n_tiles; t += tstep, ++n) {
    int b, tok0; tile_coords(t, b, tok0);
    const bool has_next = t + tstep < n_tiles;
    if (tid == 0 && has_next) prefetch_tile(t + tstep);
    mbar_wait(&sm.bar_d, n & 1);

    // ---- A: du = dxc * sp over every staged row that is a token of this batch element (zero rows stay zero); four rows'
    //      sp loads in flight per thread -----------------------------------------------------------------------------------
    {
      const int c8 = tid % CH, ch0 = c8 * 8, kt = ch0 >> 6, cc = ch0 & 63;
      const __nv_bfloat16* spb = reinterpret_cast<const __nv_bfloat16*>(p.sp) + (size_t)b * S * p.D + hb * DBLK + ch0;
      for (int rb = tid / CH; rb < R; rb += 4 * (CT / CH)) {
        uint4 wsp[4];
        bool on[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int r = rb + u * (CT / CH), tok = tok0 - halo + r;
          on[u] = r < R && tok >= 0 && tok < S;
          wsp[u] = make_uint4(0, 0, 0, 0);
          if (on[u]) wsp[u] = *reinterpret_cast<const uint4*>(spb + (size_t)tok * p.D);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (!on[u]) continue;
          uint4* slot = reinterpret_cast<uint4*>(sm.dh[kt] + swz128(rb + u * (CT / CH), cc));
          float d[8], s8[8];
          cvt8<false>(*slot, d);
          cvt8<false>(wsp[u], s8);
          *slot = make_uint4(pack_bf16x2(d[0] * s8[0], d[1] * s8[1]), pack_bf16x2(d[2] * s8[2], d[3] * s8[3]),
                             pack_bf16x2(d[4] * s8[4], d[5] * s8[5]), pack_bf16x2(d[6] * s8[6], d[7] * s8[7]));
        }
      }
    }
    __syncthreads();
    mbar_wait(&sm.bar_x, n & 1);

    // ---- C: weight gradients: thread = 4 channels x strided rows; dwc[tap] += du[s] x[s + off(tap)], dbc += du[s] ----------------
    {
      const int c4 = tid % CG, ch0 = c4 * 4, kt = ch0 >> 6, cc = ch0 & 63;
      for (int r = tid / CG; r < 128; r += RL) {
        const int tok = tok0 + r;
        if (tok >= S) break;
        const uint32_t ok = tap_mask(tok);
        const uint2 wd = *reinterpret_cast<const uint2*>(sm.dh[kt] + swz128(halo + r, cc));
        const __nv_bfloat162* hd = reinterpret_cast<const __nv_bfloat162*>(&wd);
        const float2 d01 = __bfloat1622float2(hd[0]), d23 = __bfloat1622float2(hd[1]);
        bacc[0] += d01.x; bacc[1] += d01.y; bacc[2] += d23.x; bacc[3] += d23.y;
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          if (!((ok >> tap) & 1u)) continue;
          const int rr = halo + r + (tap / 3 - 1) * GW + (tap % 3 - 1);
          const uint2 wx = *reinterpret_cast<const uint2*>(sm.xh[kt] + swz128(rr, cc));
          float x0, x1, x2, x3;
          if (FP16) {
            const __half2* hx = reinterpret_cast<const __half2*>(&wx);
            const float2 a = __half22float2(hx[0]), c2 = __half22float2(hx[1]);
            x0 = a.x; x1 = a.y; x2 = c2.x; x3 = c2.y;
          } else {
            const __nv_bfloat162* hx = reinterpret_cast<const __nv_bfloat162*>(&wx);
            const float2 a = __bfloat1622float2(hx[0]), c2 = __bfloat1622float2(hx[1]);
            x0 = a.x; x1 = a.y; x2 = c2.x; x3 = c2.y;
          }
          wacc[tap][0] = fmaf(d01.x, x0, wacc[tap][0]); wacc[tap][1] = fmaf(d01.y, x1, wacc[tap][1]);
          wacc[tap][2] = fmaf(d23.x, x2, wacc[tap][2]); wacc[tap][3] = fmaf(d23.y, x3, wacc[tap][3]);
        }
      }
    }
    __syncthreads();   // the x rows are dead: the next tile's stream in under step B
    if (tid == 0 && has_next) load_rows(&sm.xh[0][0], &maps.x, &sm.bar_x, t + tstep);

    // ---- B: dx[t] = dxv[t] + sum_taps w[tap] du[t - off(tap)]: thread = 8 channels x RPT consecutive tokens, sliding window -------
    {
      const int c8 = tid % CH, ch0 = c8 * 8, kt = ch0 >> 6, cc = ch0 & 63;
      const int r0 = (tid / CH) * RPT;
      float acc[RPT][8];
      uint32_t ok[RPT];
#pragma unroll
      for (int it = 0; it < RPT; ++it) {
        const int tok = tok0 + r0 + it;
        ok[it] = tap_mask(tok);
        uint4 wv = make_uint4(0, 0, 0, 0);
        if (tok < S) wv = *reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(p.dxv) + ((size_t)b * S + tok) * p.D + hb * DBLK + ch0);
        cvt8<false>(wv, acc[it]);
      }
#pragma unroll
      for (int dy = -1; dy <= 1; ++dy) {
        float wt[3][8];
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
          const float4 w0 = *reinterpret_cast<const float4*>(&sm.cw[(dy + 1) * 3 + dx][ch0]);
          const float4 w1 = *reinterpret_cast<const float4*>(&sm.cw[(dy + 1) * 3 + dx][ch0 + 4]);
          wt[dx][0] = w0.x; wt[dx][1] = w0.y; wt[dx][2] = w0.z; wt[dx][3] = w0.w;
          wt[dx][4] = w1.x; wt[dx][5] = w1.y; wt[dx][6] = w1.z; wt[dx][7] = w1.w;
        }
        // the source of tap (dy, dx) for output token it is the staged du row  it - dy*GW - dx ; j = it - dx
#pragma unroll
        for (int j = -1; j <= RPT; ++j) {
          const uint4 w = *reinterpret_cast<const uint4*>(sm.dh[kt] + swz128(halo + r0 - dy * GW + j, cc));
          float dv8[8];
          cvt8<false>(w, dv8);
#pragma unroll
          for (int dx = -1; dx <= 1; ++dx) {
            const int it = j + dx;
            if (it < 0 || it >= RPT) continue;
            // the source token's tap (dy, dx) lands on this token iff this token has a neighbour at (-dy, -dx)
            if (!((ok[it] >> (8 - ((dy + 1) * 3 + (dx + 1)))) & 1u)) continue;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[it][e] = fmaf(dv8[e], wt[dx + 1][e], acc[it][e]);
          }
        }
      }
#pragma unroll
      for (int it = 0; it < RPT; ++it) {
        const int tok = tok0 + r0 + it;
        if (tok >= S) continue;
        const size_t off = ((size_t)b * S + tok) * p.D + hb * DBLK + ch0;
        float* a_ = acc[it];
        uint4 o;
        if (FP16) {
          __half2 h[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(a_[2 * e], a_[2 * e + 1]);
          o = *reinterpret_cast<const uint4*>(h);
        } else {
          o = make_uint4(pack_bf16x2(a_[0], a_[1]), pack_bf16x2(a_[2], a_[3]), pack_bf16x2(a_[4], a_[5]), pack_bf16x2(a_[6], a_[7]));
        }
        *reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(p.dx) + off) = o;
      }
    }
    __syncthreads();   // the du rows are dead
    if (tid == 0 && has_next) load_rows(&sm.dh[0][0], &maps.dxc, &sm.bar_d, t + tstep);
  }

  // ---- per-CTA partials [10][DBLK] (9 taps in the forward's tap order + bias), row lanes summed in fixed order; two passes of
  //      five rows each through the (dead) staging buffers ------------------------------------------------------------------
  float* red = reinterpret_cast<float*>(sm.xh);   // [RL][5][DBLK] floats per pass
  for (int k0 = 0; k0 < 10; k0 += 5) {
    __syncthreads();
    {
      const int c4 = tid % CG, rl = tid / CG;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
#pragma unroll
        for (int kk = 0; kk < 5; ++kk) {
          const int k = k0 + kk;
          red[(rl * 5 + kk) * DBLK + c4 * 4 + e] = (k < 9) ? wacc[k < 9 ? k : 0][e] : bacc[e];
        }
      }
    }
    __syncthreads();
    for (int e = tid; e < 5 * DBLK; e += CT) {
      float s = 0.f;
      for (int rl = 0; rl < RL; ++rl) s += red[rl * 5 * DBLK + e];
      ws[(size_t)blockIdx.x * 10 * DBLK + k0 * DBLK + e] = s;
    }
  }
}

// dwc[(h*d + ch)*9 + orig_tap], dbc[h*d + ch] = fixed-order sums over the CTAs of block h
__global__ void conv_bwd_reduce_kernel(const float* __restrict__ ws, const mlstm_conv_bwd_params p, const int per_block, const int d) {
  const int e0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;   // eight lanes per output element
  const int n = 10 * p.D;
  const int e = min(e0, n - 1);
  const int k = e / p.D, col = e % p.D, h = col / d, ch = col % d;   // k: tap (0..8) or 9 = bias
  const float s = ordered_sum8(ws + (size_t)h * 10 * d + (size_t)k * d + ch, per_block, (size_t)p.NH * 10 * d);
  if (e0 < n && (threadIdx.x & 7) == 0) {
    if (k < 9) p.dwc[(size_t)col * 9 + (p.rotate ? 8 - k : k)] = s;
    else if (p.dbc) p.dbc[col] = s;
  }
}

int conv_bwd_ctas_per_block(const mlstm_conv_bwd_params& p) {
  int per_block = 148 / p.NH;    // fixed: the workspace size must not depend on the device
  const int n_tiles = p.B * ((p.GH * p.GW + 127) / 128);
  if (per_block < 1) per_block = 1;
  if (per_block > n_tiles) per_block = n_tiles;
  return per_block;
}

template <int DBLK, bool FP16>
int launch_conv_bwd(const mlstm_conv_bwd_params& p, const ConvBwdMaps& maps, int halo, cudaStream_t st) {
  const size_t smem = sizeof(SmemCB<DBLK>);
  cudaError_t e = set_max_smem_once(reinterpret_cast<const void*>(conv_bwd_kernel<DBLK, FP16>), smem);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(conv_bwd, %zu B): %s", smem, cudaGetErrorString(e));
    return MLSTM_ERR_CUDA;
  }
  const int S = p.GH * p.GW, tiles_per_batch = (S + 127) / 128;
  const int per_block = conv_bwd_ctas_per_block(p);
  float* ws = reinterpret_cast<float*>(p.workspace);
  conv_bwd_kernel<DBLK, FP16><<<dim3(per_block * p.NH), dim3(CT), smem, st>>>(maps, p, halo, tiles_per_batch, ws);
  count_launch();
  const int n = 10 * p.D;
  conv_bwd_reduce_kernel<<<(n * 8 + 255) / 256, 256, 0, st>>>(ws, p, per_block, DBLK);
  count_launch();
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("conv_bwd launch failed: %s", cudaGetErrorString(e));
    return MLSTM_ERR_CUDA;
  }
  return MLSTM_OK;
}

}  // namespace
}  // namespace mlstm

extern "C" {

size_t mlstm_b200_conv_bwd_workspace_bytes(const mlstm_conv_bwd_params* p) {
  if (!p || p->B <= 0 || p->NH <= 0 || p->D % p->NH != 0 || p->GH <= 0 || p->GW <= 0) return 0;
  return sizeof(float) * (size_t)conv_bwd_ctas_per_block(*p) * (size_t)p->NH * 10 * (size_t)(p->D / p->NH);
}

int mlstm_b200_conv_bwd(const mlstm_conv_bwd_params* p, void* cuda_stream) {
  clear_error();
  if (!p) { set_error("params is NULL"); return MLSTM_ERR_INVALID_ARG; }
  if (p->abi_version != MLSTM_B200_ABI_VERSION) {
    set_error("abi_version %d != %d", p->abi_version, MLSTM_B200_ABI_VERSION);
    return MLSTM_ERR_INVALID_ARG;
  }
  if (p->B < 0 || p->x_dtype < 0 || p->x_dtype > 1) { set_error("bad B / x_dtype"); return MLSTM_ERR_INVALID_ARG; }
  if (!qkv_shape_ok(p->D, p->NH, p->GH, p->GW, p->ld_x)) {
    set_error("conv backward: D / NH must be 64 or 128, grid width <= 80, ld_x a multiple of 8 and >= D (D=%d NH=%d GH=%d GW=%d ld_x=%lld)",
              p->D, p->NH, p->GH, p->GW, (long long)p->ld_x);
    return MLSTM_ERR_UNSUPPORTED;
  }
  if (!p->dwc) { set_error("conv backward: dwc is NULL"); return MLSTM_ERR_INVALID_ARG; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  if (p->B == 0) {
    int rc0 = bind_device(p->dwc);
    if (rc0) return rc0;
    cudaMemsetAsync(p->dwc, 0, sizeof(float) * 9 * (size_t)p->D, st);
    if (p->dbc) cudaMemsetAsync(p->dbc, 0, sizeof(float) * (size_t)p->D, st);
    return MLSTM_OK;
  }
  if (!p->x || !p->dxc || !p->dxv || !p->sp || !p->conv_w || !p->dx) { set_error("conv backward: null pointer"); return MLSTM_ERR_INVALID_ARG; }
  if (((uintptr_t)p->x | (uintptr_t)p->dxc | (uintptr_t)p->dxv | (uintptr_t)p->sp | (uintptr_t)p->dx) & 15u) {
    set_error("conv backward: x, dxc, dxv, sp, dx must be 16-byte aligned (16-byte vector and TMA accesses)");
    return MLSTM_ERR_INVALID_ARG;
  }
  if (!p->workspace || p->workspace_bytes < mlstm_b200_conv_bwd_workspace_bytes(p)) {
    set_error("conv backward: workspace too small (%zu < %zu)", p->workspace ? p->workspace_bytes : (size_t)0,
              mlstm_b200_conv_bwd_workspace_bytes(p));
    return MLSTM_ERR_WORKSPACE;
  }
  int rc;
  if ((rc = bind_device(p->x))) return rc;
  const int S = p->GH * p->GW, d = p->D / p->NH;
  const int halo = ((p->GW + 1 + 7) / 8) * 8;
  ConvBwdMaps maps;
  int r = 0;
  r |= make_rows_tmap(&maps.x, p->x, p->D, S, p->B, p->ld_x, (128 + 2 * halo) / 2);
  r |= make_rows_tmap(&maps.dxc, p->dxc, p->D, S, p->B, p->D, (128 + 2 * halo) / 2);
  if (r) {
    set_error("cuTensorMapEncodeTiled failed (%d): pointers must be 16-byte aligned", r);
    return r == -1 ? MLSTM_ERR_NO_DEVICE : MLSTM_ERR_INVALID_ARG;
  }
  if (d == 128) return p->x_dtype ? launch_conv_bwd<128, true>(*p, maps, halo, st) : launch_conv_bwd<128, false>(*p, maps, halo, st);
  return p->x_dtype ? launch_conv_bwd<64, true>(*p, maps, halo, st) : launch_conv_bwd<64, false>(*p, maps, halo, st);
}

}  // extern "C"

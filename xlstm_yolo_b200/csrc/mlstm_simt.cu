// fp32 SIMT kernels for the mLSTM cell: the "fp32 mode" (tolerance 1e-4) and the
// small-head-dim path (DH not a multiple of 64, e.g. the reference's default
// qkv_block_size=16, nn/modules/block.py:1823).  Same chunkwise algorithm as the tcgen05
// kernels (chunk tile L = 32 = one warp of gate lanes), all arithmetic in fp32 on CUDA
// cores, state (C, n, m) resident in shared memory for the whole scan.
//
// Follows the reference's chunkwise form, nn/modules/vision_lstm/xlstm/blocks/mlstm/
// backends.py:149-263 (state carry :196-218, intra-chunk :220-263); the backward is the
// hand-derived adjoint documented in tests/emu_kernel_dataflow.py and DESIGN.md.
//
// One CTA per (batch, head).  `reverse` walks the tokens from S-1 down to 0 by index
// arithmetic only (no flipped copy; vision_lstm2.py:479-480,505-506).
#include "mlstm_common.cuh"

namespace mlstm {
namespace {

constexpr int L = 32;     // chunk tile (rows per step)
constexpr int NT = 256;   // threads per CTA
constexpr int LP = L + 1; // padded row of the L x L matrices

// Value-dimension slicing.  The state C [DK][DV] must fit in shared memory; beyond DH = 128 the
// launcher runs the kernels once per DV slice (<= 64 columns) on offset v/h/dh/dv pointers.  The
// forward is independent per slice.  The backward is linear in the slices: with
// dn_s = -coef (dh_s . h_s), every slice yields its share of dq, dk, R, di and df; the shares are
// summed in fp32 (workspace accumulators for dq/dk, the fp32 outputs for di/df) and the last slice
// rounds once to the output dtype.
struct Slice {
  int cld;          // leading dimension (full DV) of c_initial / c_last
  int first, last;  // first / last slice of the launch sequence
  float* dq_acc;    // fp32 [B*NH*S][DK] accumulators (nullptr when there is a single slice)
  float* dk_acc;
};

struct Gates {  // per-chunk gate-derived vectors, filled by warp 0
  float u[L], M[L], w[L], mrow[L], kw[L], N[L], dn[L], fpre[L], aux[L];
  float decay, m_next;
};

__device__ __forceinline__ int tok_of(int pos, int S, int reverse) { return reverse ? (S - 1 - pos) : pos; }

template <typename T>
__device__ __forceinline__ const T* act_ptr(const mlstm_act& a, int b, int h, int tok) {
  return reinterpret_cast<const T*>(a.ptr) + (int64_t)b * a.stride_b + (int64_t)h * a.stride_h + (int64_t)tok * a.stride_s;
}
template <typename T>
__device__ __forceinline__ T* act_ptr_w(const mlstm_act& a, int b, int h, int tok) {
  return reinterpret_cast<T*>(a.ptr) + (int64_t)b * a.stride_b + (int64_t)h * a.stride_h + (int64_t)tok * a.stride_s;
}
__device__ __forceinline__ float* gate_ptr(const mlstm_gate& g, int b, int h, int tok) {
  return g.ptr + (int64_t)b * g.stride_b + (int64_t)h * g.stride_h + (int64_t)tok * g.stride_s;
}

// dst[r*ld + d] = mul * a[b,h,tok(pos0+r),d]; rows >= nvalid are zero.
template <typename T>
__device__ __forceinline__ void load_rows(float* dst, int ld, const mlstm_act& a, int b, int h, int pos0,
                                          int nvalid, int D, float mul, int S, int reverse) {
  for (int e = threadIdx.x; e < L * D; e += NT) {
    int r = e / D, d = e - r * D;
    float x = 0.f;
    if (r < nvalid) x = mul * to_f32<T>(act_ptr<T>(a, b, h, tok_of(pos0 + r, S, reverse))[d]);
    dst[r * ld + d] = x;
  }
}

// Gate vectors for a chunk, forward recurrence (m_prev known).  Called by warp 0.
__device__ __forceinline__ void gates_forward(Gates& G, const mlstm_params& p, int b, int h, int pos0, int nvalid,
                                              float m_prev, int lane) {
  bool valid = lane < nvalid;
  float fi = 0.f, ii = -INFINITY, logf = 0.f;
  if (valid) {
    int tok = tok_of(pos0 + lane, p.S, p.reverse);
    fi = *gate_ptr(p.f, b, h, tok);
    ii = igate_log(p, *gate_ptr(p.i, b, h, tok));
    logf = log_sigmoid(fi);
  }
  if (p.gate_mode) m_prev = 0.f;   // sigmoid input gate: no stabiliser, m_t == 0
  float bsum = warp_scan_add(logf, lane);
  float u = ii - bsum;
  float M = p.gate_mode ? -bsum : fmaxf(m_prev, warp_scan_max(u, lane));
  float ML = __shfl_sync(0xffffffffu, M, 31);
  float g = __shfl_sync(0xffffffffu, bsum, 31);
  G.u[lane] = u;
  G.M[lane] = M;
  G.w[lane] = __expf(m_prev - M);
  G.mrow[lane] = bsum + M;
  G.kw[lane] = __expf(u - ML);
  G.fpre[lane] = fi;
  if (lane == 0) {
    G.decay = __expf(m_prev - ML);
    G.m_next = g + ML;
  }
}

// ---------------------------------------------------------------------------------------
// Forward
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT) simt_fwd_kernel(const mlstm_params p, const float scale, const Slice sl) {
  extern __shared__ float sm[];
  const int DK = p.DHQK, DV = p.DHV, LQ = DK + 1, LV = DV + 1, LC = DV + 1;
  float* Cs = sm;                  // [DK][LC]
  float* ns = Cs + DK * LC;        // [DK]
  float* qs = ns + DK;             // [L][LQ]  (pre-multiplied by scale)
  float* ks = qs + L * LQ;         // [L][LQ]
  float* vs = ks + L * LQ;         // [L][LV]
  float* Ps = vs + L * LV;         // [L][LP]
  Gates& G = *reinterpret_cast<Gates*>(Ps + L * LP);
  __shared__ float m_carry;

  const int bh = blockIdx.x, b = bh / p.NH, h = bh % p.NH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.S;

  for (int e = tid; e < DK * DV; e += NT) {
    int dk = e / DV, dv = e - dk * DV;
    Cs[dk * LC + dv] = p.c_initial ? p.c_initial[((int64_t)bh * DK + dk) * sl.cld + dv] : 0.f;
  }
  for (int e = tid; e < DK; e += NT) ns[e] = p.n_initial ? p.n_initial[(int64_t)bh * DK + e] : 0.f;
  if (tid == 0) m_carry = p.m_initial ? p.m_initial[bh] : 0.f;
  __syncthreads();

  const int NC = (S + L - 1) / L;
  for (int c = 0; c < NC; ++c) {
    const int pos0 = c * L, nvalid = min(L, S - pos0);
    load_rows<T>(qs, LQ, p.q, b, h, pos0, nvalid, DK, scale, S, p.reverse);
    load_rows<T>(ks, LQ, p.k, b, h, pos0, nvalid, DK, 1.f, S, p.reverse);
    load_rows<T>(vs, LV, p.v, b, h, pos0, nvalid, DV, 1.f, S, p.reverse);
    if (warp == 0) gates_forward(G, p, b, h, pos0, nvalid, m_carry, lane);
    __syncthreads();

    // P = (Q K^T) * D, row normaliser (backends.py:242-252)
    for (int t = warp * 4; t < warp * 4 + 4; ++t) {
      float acc = 0.f;
      for (int d = 0; d < DK; ++d) acc = fmaf(qs[t * LQ + d], ks[lane * LQ + d], acc);
      float pv = (lane <= t) ? acc * __expf(G.u[lane] - G.M[t]) : 0.f;
      Ps[t * LP + lane] = pv;
      float rs = warp_sum(pv);
      float qn = 0.f;
      for (int d = lane; d < DK; d += 32) qn = fmaf(qs[t * LQ + d], ns[d], qn);
      qn = warp_sum(qn);
      if (lane == 0) {
        float nr = rs + G.w[t] * qn;
        float N = fmaxf(fabsf(nr), __expf(-G.mrow[t])) + p.eps;
        G.N[t] = N;
        if (t < nvalid && p.n_row) {
          int tok = tok_of(pos0 + t, S, p.reverse);
          p.n_row[(int64_t)bh * S + tok] = nr;
          p.m_row[(int64_t)bh * S + tok] = G.mrow[t];
        }
      }
    }
    __syncthreads();

    // h = (P V + w * Q C_prev) / N   (backends.py:254-263)
    for (int e = tid; e < L * DV; e += NT) {
      int t = e / DV, dv = e - t * DV;
      if (t >= nvalid) continue;
      float a1 = 0.f, a2 = 0.f;
      for (int j = 0; j <= t; ++j) a1 = fmaf(Ps[t * LP + j], vs[j * LV + dv], a1);
      for (int d = 0; d < DK; ++d) a2 = fmaf(qs[t * LQ + d], Cs[d * LC + dv], a2);
      float hv = (a1 + G.w[t] * a2) / G.N[t];
      act_ptr_w<T>(p.h, b, h, tok_of(pos0 + t, S, p.reverse))[dv] = from_f32<T>(hv);
    }
    // kbar_j = kw_j * k_j (in place; q,k products are done)
    for (int e = tid; e < L * DK; e += NT) {
      int j = e / DK, d = e - j * DK;
      ks[j * LQ + d] *= G.kw[j];
    }
    __syncthreads();

    // state carry (backends.py:196-218)
    const float decay = G.decay;
    for (int e = tid; e < DK * DV; e += NT) {
      int dk = e / DV, dv = e - dk * DV;
      float acc = 0.f;
#pragma unroll 8
      for (int j = 0; j < L; ++j) acc = fmaf(ks[j * LQ + dk], vs[j * LV + dv], acc);
      Cs[dk * LC + dv] = fmaf(decay, Cs[dk * LC + dv], acc);
    }
    for (int d = tid; d < DK; d += NT) {
      float acc = 0.f;
      for (int j = 0; j < L; ++j) acc += ks[j * LQ + d];
      ns[d] = fmaf(decay, ns[d], acc);
    }
    if (tid == 0) m_carry = G.m_next;
    __syncthreads();
  }

  if (p.c_last) {
    for (int e = tid; e < DK * DV; e += NT) {
      int dk = e / DV, dv = e - dk * DV;
      p.c_last[((int64_t)bh * DK + dk) * sl.cld + dv] = Cs[dk * LC + dv];
    }
    for (int e = tid; e < DK; e += NT) p.n_last[(int64_t)bh * DK + e] = ns[e];
    if (tid == 0) p.m_last[bh] = m_carry;
  }
}

// ---------------------------------------------------------------------------------------
// Backward kernel A: forward walk, produces dq, dn_row, R_row = q . dq
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT) simt_bwd_dq_kernel(const mlstm_params p, const float scale, float* __restrict__ ws_dn,
                                                         float* __restrict__ ws_R, const Slice sl) {
  extern __shared__ float sm[];
  const int DK = p.DHQK, DV = p.DHV, LQ = DK + 1, LV = DV + 1, LC = DV + 1;
  float* Cs = sm;
  float* ns = Cs + DK * LC;
  float* qs = ns + DK;             // scaled
  float* ks = qs + L * LQ;
  float* vs = ks + L * LQ;
  float* dhs = vs + L * LV;
  float* Gs = dhs + L * LV;        // [L][LQ]  G = dH C^T, then dq in place
  float* dSs = Gs + L * LQ;        // [L][LP]
  Gates& G = *reinterpret_cast<Gates*>(dSs + L * LP);
  __shared__ float m_carry;

  const int bh = blockIdx.x, b = bh / p.NH, h = bh % p.NH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.S;

  for (int e = tid; e < DK * DV; e += NT) {
    int dk = e / DV, dv = e - dk * DV;
    Cs[dk * LC + dv] = p.c_initial ? p.c_initial[((int64_t)bh * DK + dk) * sl.cld + dv] : 0.f;
  }
  for (int e = tid; e < DK; e += NT) ns[e] = p.n_initial ? p.n_initial[(int64_t)bh * DK + e] : 0.f;
  if (tid == 0) m_carry = p.m_initial ? p.m_initial[bh] : 0.f;
  __syncthreads();

  const int NC = (S + L - 1) / L;
  for (int c = 0; c < NC; ++c) {
    const int pos0 = c * L, nvalid = min(L, S - pos0);
    load_rows<T>(qs, LQ, p.q, b, h, pos0, nvalid, DK, scale, S, p.reverse);
    load_rows<T>(ks, LQ, p.k, b, h, pos0, nvalid, DK, 1.f, S, p.reverse);
    load_rows<T>(vs, LV, p.v, b, h, pos0, nvalid, DV, 1.f, S, p.reverse);
    load_rows<T>(dhs, LV, p.dh, b, h, pos0, nvalid, DV, 1.f, S, p.reverse);
    if (warp == 0) gates_forward(G, p, b, h, pos0, nvalid, m_carry, lane);
    __syncthreads();

    // per-row: N_t, dn_t (from dh.h), then dS = (Z/N + dn) * D
    for (int t = warp * 4; t < warp * 4 + 4; ++t) {
      float N = 1.f, dn = 0.f;
      if (t < nvalid) {
        int tok = tok_of(pos0 + t, S, p.reverse);
        const T* hrow = act_ptr<T>(p.h, b, h, tok);
        float hd = 0.f;
        for (int d = lane; d < DV; d += 32) hd = fmaf(dhs[t * LV + d], to_f32<T>(hrow[d]), hd);
        hd = warp_sum(hd);
        float nr = p.n_row[(int64_t)bh * S + tok];
        float floor_ = __expf(-G.mrow[t]);
        N = fmaxf(fabsf(nr), floor_) + p.eps;
        dn = (fabsf(nr) >= floor_) ? -copysignf(1.f, nr) * hd / N : 0.f;
        if (lane == 0) ws_dn[(int64_t)bh * S + tok] = dn;
      }
      if (lane == 0) { G.N[t] = N; G.dn[t] = dn; }
      float z = 0.f;
      for (int d = 0; d < DV; ++d) z = fmaf(dhs[t * LV + d], vs[lane * LV + d], z);
      float ds = (lane <= t) ? (z / N + dn) * __expf(G.u[lane] - G.M[t]) : 0.f;
      dSs[t * LP + lane] = ds;
    }
    // G[t][dk] = sum_dv dh[t][dv] C[dk][dv]
    for (int e = tid; e < L * DK; e += NT) {
      int t = e / DK, dk = e - t * DK;
      float acc = 0.f;
      for (int d = 0; d < DV; ++d) acc = fmaf(dhs[t * LV + d], Cs[dk * LC + d], acc);
      Gs[t * LQ + dk] = acc;
    }
    __syncthreads();

    // dq = s * [ dS K + w (G / N + dn n_prev) ]
    for (int e = tid; e < L * DK; e += NT) {
      int t = e / DK, dk = e - t * DK;
      float acc = 0.f;
      for (int j = 0; j <= t; ++j) acc = fmaf(dSs[t * LP + j], ks[j * LQ + dk], acc);
      float dqv = scale * (acc + G.w[t] * (Gs[t * LQ + dk] / G.N[t] + G.dn[t] * ns[dk]));
      Gs[t * LQ + dk] = dqv;   // this slice's share (R below is linear in it)
      if (t < nvalid) {
        const int tok = tok_of(pos0 + t, S, p.reverse);
        float out = dqv;
        if (sl.dq_acc) {
          float* a = sl.dq_acc + ((int64_t)bh * S + tok) * DK + dk;
          if (!sl.first) out += *a;
          if (!sl.last) *a = out;
        }
        if (sl.last) act_ptr_w<T>(p.dq, b, h, tok)[dk] = from_f32<T>(out);
      }
    }
    __syncthreads();

    // R_t = q_t . dq_t ; kbar in place
    for (int t = warp * 4; t < warp * 4 + 4; ++t) {
      float r = 0.f;
      for (int d = lane; d < DK; d += 32) r = fmaf(qs[t * LQ + d], Gs[t * LQ + d], r);
      r = warp_sum(r);
      if (lane == 0 && t < nvalid) {
        float* R = ws_R + (int64_t)bh * S + tok_of(pos0 + t, S, p.reverse);
        *R = (sl.first ? 0.f : *R) + r / scale;
      }
    }
    for (int e = tid; e < L * DK; e += NT) {
      int j = e / DK, d = e - j * DK;
      ks[j * LQ + d] *= G.kw[j];
    }
    __syncthreads();

    const float decay = G.decay;
    for (int e = tid; e < DK * DV; e += NT) {
      int dk = e / DV, dv = e - dk * DV;
      float acc = 0.f;
#pragma unroll 8
      for (int j = 0; j < L; ++j) acc = fmaf(ks[j * LQ + dk], vs[j * LV + dv], acc);
      Cs[dk * LC + dv] = fmaf(decay, Cs[dk * LC + dv], acc);
    }
    for (int d = tid; d < DK; d += NT) {
      float acc = 0.f;
      for (int j = 0; j < L; ++j) acc += ks[j * LQ + d];
      ns[d] = fmaf(decay, ns[d], acc);
    }
    if (tid == 0) m_carry = G.m_next;
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------
// Backward kernel B: reverse walk carrying (dC, dn), produces dk, dv, di, df
// ---------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(NT) simt_bwd_dkv_kernel(const mlstm_params p, const float scale,
                                                          const float* __restrict__ ws_dn, const float* __restrict__ ws_R,
                                                          const Slice sl) {
  extern __shared__ float sm[];
  const int DK = p.DHQK, DV = p.DHV, LQ = DK + 1, LV = DV + 1, LC = DV + 1;
  float* dCs = sm;                 // [DK][LC]
  float* dnv = dCs + DK * LC;      // [DK]
  float* qs = dnv + DK;            // scaled
  float* ks = qs + L * LQ;
  float* vs = ks + L * LQ;
  float* dhs = vs + L * LV;
  float* dks = dhs + L * LV;       // [L][LQ] dk tile
  float* Ets = dks + L * LQ;       // [L][LP] (E^T / N)[j][t]
  float* dSts = Ets + L * LP;      // [L][LP] dS^T[j][t]
  Gates& G = *reinterpret_cast<Gates*>(dSts + L * LP);
  __shared__ float df_carry;

  const int bh = blockIdx.x, b = bh / p.NH, h = bh % p.NH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = p.S;

  for (int e = tid; e < DK * LC; e += NT) dCs[e] = 0.f;
  for (int e = tid; e < DK; e += NT) dnv[e] = 0.f;
  if (tid == 0) df_carry = 0.f;
  __syncthreads();

  const int NC = (S + L - 1) / L;
  for (int c = NC - 1; c >= 0; --c) {
    const int pos0 = c * L, nvalid = min(L, S - pos0);
    load_rows<T>(qs, LQ, p.q, b, h, pos0, nvalid, DK, scale, S, p.reverse);
    load_rows<T>(ks, LQ, p.k, b, h, pos0, nvalid, DK, 1.f, S, p.reverse);
    load_rows<T>(vs, LV, p.v, b, h, pos0, nvalid, DV, 1.f, S, p.reverse);
    load_rows<T>(dhs, LV, p.dh, b, h, pos0, nvalid, DV, 1.f, S, p.reverse);
    if (warp == 0) {
      bool valid = lane < nvalid;
      float fi = 0.f, ii = -INFINITY, logf = 0.f, mrow = 0.f, nr = 0.f, dn = 0.f, R = 0.f;
      if (valid) {
        int tok = tok_of(pos0 + lane, S, p.reverse);
        fi = *gate_ptr(p.f, b, h, tok);
        ii = igate_log(p, *gate_ptr(p.i, b, h, tok));
        logf = log_sigmoid(fi);
        mrow = p.m_row[(int64_t)bh * S + tok];
        nr = p.n_row[(int64_t)bh * S + tok];
        dn = ws_dn[(int64_t)bh * S + tok];
        R = sl.first ? ws_R[(int64_t)bh * S + tok] : 0.f;   // the complete R enters df once
      }
      float bsum = warp_scan_add(logf, lane);
      float u = ii - bsum;
      float M = mrow - bsum;
      float ML = __shfl_sync(0xffffffffu, M, nvalid - 1);
      if (!valid) M = ML;
      float m_prev = (pos0 == 0) ? (p.m_initial ? p.m_initial[bh] : 0.f)
                                 : p.m_row[(int64_t)bh * S + tok_of(pos0 - 1, S, p.reverse)];
      if (p.gate_mode) m_prev = 0.f;
      G.u[lane] = u;
      G.M[lane] = M;
      G.w[lane] = __expf(m_prev - M);
      G.kw[lane] = __expf(u - ML);
      G.N[lane] = valid ? fmaxf(fabsf(nr), __expf(-mrow)) + p.eps : 1.f;
      G.dn[lane] = dn;
      G.fpre[lane] = fi;
      G.aux[lane] = R;
      if (lane == 0) G.decay = __expf(m_prev - ML);
    }
    __syncthreads();

    // (E^T / N)[j][t] and dS^T[j][t]; rows j owned by the warp, lane = t
    for (int j = warp * 4; j < warp * 4 + 4; ++j) {
      float s_ = 0.f, z = 0.f;
      for (int d = 0; d < DK; ++d) s_ = fmaf(ks[j * LQ + d], qs[lane * LQ + d], s_);
      for (int d = 0; d < DV; ++d) z = fmaf(vs[j * LV + d], dhs[lane * LV + d], z);
      float D = (j <= lane) ? __expf(G.u[j] - G.M[lane]) : 0.f;
      float N = G.N[lane];
      Ets[j * LP + lane] = s_ * D / N;
      dSts[j * LP + lane] = (z / N + G.dn[lane]) * D;
    }
    __syncthreads();

    // dv_j = sum_t (E^T/N) dh_t + kw_j (k_j dC)
    for (int e = tid; e < L * DV; e += NT) {
      int j = e / DV, dv = e - j * DV;
      if (j >= nvalid) continue;
      float a1 = 0.f, a2 = 0.f;
      for (int t = j; t < L; ++t) a1 = fmaf(Ets[j * LP + t], dhs[t * LV + dv], a1);
      for (int d = 0; d < DK; ++d) a2 = fmaf(ks[j * LQ + d], dCs[d * LC + dv], a2);
      act_ptr_w<T>(p.dv, b, h, tok_of(pos0 + j, S, p.reverse))[dv] = from_f32<T>(a1 + G.kw[j] * a2);
    }
    // dk_j = sum_t dS^T q_t(scaled) + kw_j (dC v_j + dnv)
    for (int e = tid; e < L * DK; e += NT) {
      int j = e / DK, dk = e - j * DK;
      float a1 = 0.f, a2 = 0.f;
      for (int t = j; t < L; ++t) a1 = fmaf(dSts[j * LP + t], qs[t * LQ + dk], a1);
      for (int d = 0; d < DV; ++d) a2 = fmaf(dCs[dk * LC + d], vs[j * LV + d], a2);
      float dkv = a1 + G.kw[j] * (a2 + dnv[dk]);
      dks[j * LQ + dk] = dkv;   // this slice's share (K_j below is linear in it)
      if (j < nvalid) {
        const int tok = tok_of(pos0 + j, S, p.reverse);
        float out = dkv;
        if (sl.dk_acc) {
          float* a = sl.dk_acc + ((int64_t)bh * S + tok) * DK + dk;
          if (!sl.first) out += *a;
          if (!sl.last) *a = out;
        }
        if (sl.last) act_ptr_w<T>(p.dk, b, h, tok)[dk] = from_f32<T>(out);
      }
    }
    __syncthreads();

    // K_j = k_j . dk_j  -> di, and the reverse cumsum for df
    for (int j = warp * 4; j < warp * 4 + 4; ++j) {
      float r = 0.f;
      for (int d = lane; d < DK; d += 32) r = fmaf(ks[j * LQ + d], dks[j * LQ + d], r);
      r = warp_sum(r);
      if (lane == 0) G.mrow[j] = r;  // reuse as K_j
    }
    // state update: dC <- decay dC + sum_t (w_t / N_t) q_t(scaled) (x) dh_t ; dnv likewise
    const float decay = G.decay;
    for (int e = tid; e < DK * DV; e += NT) {
      int dk = e / DV, dv = e - dk * DV;
      float acc = 0.f;
#pragma unroll 8
      for (int t = 0; t < L; ++t) acc = fmaf(G.w[t] / G.N[t] * qs[t * LQ + dk], dhs[t * LV + dv], acc);
      dCs[dk * LC + dv] = fmaf(decay, dCs[dk * LC + dv], acc);
    }
    for (int d = tid; d < DK; d += NT) {
      float acc = 0.f;
      for (int t = 0; t < L; ++t) acc = fmaf(G.w[t] * G.dn[t], qs[t * LQ + d], acc);
      dnv[d] = fmaf(decay, dnv[d], acc);
    }
    __syncthreads();
    if (warp == 0) {
      bool valid = lane < nvalid;
      float Kj = valid ? G.mrow[lane] : 0.f;
      float dB = valid ? (G.aux[lane] - Kj) : 0.f;
      // suffix sum over lanes: reverse, inclusive scan, reverse back
      float x = __shfl_sync(0xffffffffu, dB, 31 - lane);
      x = warp_scan_add(x, lane);
      float rc = __shfl_sync(0xffffffffu, x, 31 - lane) + df_carry;
      if (valid) {
        int tok = tok_of(pos0 + lane, S, p.reverse);
        float fi = G.fpre[lane];
        float* di = gate_ptr(p.di, b, h, tok);
        float* df = gate_ptr(p.df, b, h, tok);
        const float dfv = rc / (1.f + __expf(fi));  // sigmoid(-f)
        *di = (sl.first ? 0.f : *di) + Kj * igate_dlog(p, *gate_ptr(p.i, b, h, tok));
        *df = (sl.first ? 0.f : *df) + dfv;
      }
      float c0 = __shfl_sync(0xffffffffu, rc, 0);
      if (lane == 0) df_carry = c0;
    }
    __syncthreads();
  }
}

size_t fwd_smem(int DK, int DV) {
  return sizeof(float) * ((size_t)DK * (DV + 1) + DK + 2 * L * (DK + 1) + L * (DV + 1) + L * LP) + sizeof(Gates);
}
size_t bwd_dq_smem(int DK, int DV) {
  return sizeof(float) * ((size_t)DK * (DV + 1) + DK + 3 * L * (DK + 1) + 2 * L * (DV + 1) + L * LP) + sizeof(Gates);
}
size_t bwd_dkv_smem(int DK, int DV) {
  return sizeof(float) * ((size_t)DK * (DV + 1) + DK + 3 * L * (DK + 1) + 2 * L * (DV + 1) + 2 * L * LP) + sizeof(Gates);
}

template <typename K>
int set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) {
    set_error("cudaFuncSetAttribute(%zu B smem): %s", bytes, cudaGetErrorString(e));
    return MLSTM_ERR_CUDA;
  }
  return MLSTM_OK;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s launch failed: %s", what, cudaGetErrorString(e));
    return MLSTM_ERR_CUDA;
  }
  return MLSTM_OK;
}

constexpr int SLICE_DV = 64;       // value columns per launch once the state no longer fits
constexpr size_t SMEM_MAX = 227 * 1024;

int slice_width(const mlstm_params& p) {
  return (p.DHQK <= 128 && p.DHV <= 128) ? p.DHV : (p.DHV < SLICE_DV ? p.DHV : SLICE_DV);
}
int num_slices(const mlstm_params& p) {
  const int w = slice_width(p);
  return (p.DHV + w - 1) / w;
}
size_t elem_bytes(const mlstm_params& p) { return p.dtype == MLSTM_F32 ? 4 : 2; }

// params of value-slice s: v, h, dh, dv advanced to its first column, DHV = its width
mlstm_params slice_params(const mlstm_params& p, int s, Slice* sl, float* dq_acc, float* dk_acc) {
  mlstm_params q = p;
  const int w = slice_width(p), dv0 = s * w, ns = num_slices(p);
  q.DHV = (dv0 + w <= p.DHV) ? w : (p.DHV - dv0);
  const size_t off = (size_t)dv0 * elem_bytes(p);
  auto adv = [&](mlstm_act& a) { if (a.ptr) a.ptr = reinterpret_cast<uint8_t*>(a.ptr) + off; };
  adv(q.v); adv(q.h); adv(q.dh); adv(q.dv);
  if (q.c_initial) q.c_initial += dv0;
  if (q.c_last) q.c_last += dv0;
  sl->cld = p.DHV;
  sl->first = (s == 0);
  sl->last = (s == ns - 1);
  sl->dq_acc = ns > 1 ? dq_acc : nullptr;
  sl->dk_acc = ns > 1 ? dk_acc : nullptr;
  return q;
}

}  // namespace

bool simt_supported(const mlstm_params& p) {
  if (p.DHQK < 1 || p.DHV < 1 || p.DHQK > 256 || p.DHV > 256) return false;
  return bwd_dkv_smem(p.DHQK, slice_width(p)) <= SMEM_MAX;
}

// dn per slice, R, and (sliced shapes) the fp32 dq / dk accumulators
size_t simt_bwd_workspace(const mlstm_params& p) {
  const size_t rows = (size_t)p.B * p.NH * p.S;
  const int ns = num_slices(p);
  return sizeof(float) * (rows * (ns + 1) + (ns > 1 ? 2 * rows * p.DHQK : 0));
}

int simt_fwd(const mlstm_params& p, cudaStream_t st) {
  const float scale = resolve_scale(p);
  dim3 grid(p.B * p.NH), block(NT);
  int rc;
  for (int s = 0; s < num_slices(p); ++s) {
    Slice sl;
    const mlstm_params ps = slice_params(p, s, &sl, nullptr, nullptr);
    const size_t smem = fwd_smem(ps.DHQK, ps.DHV);
    if (p.dtype == MLSTM_F32) {
      if ((rc = set_smem(simt_fwd_kernel<float>, smem))) return rc;
      simt_fwd_kernel<float><<<grid, block, smem, st>>>(ps, scale, sl);
    } else {
      if ((rc = set_smem(simt_fwd_kernel<__nv_bfloat16>, smem))) return rc;
      simt_fwd_kernel<__nv_bfloat16><<<grid, block, smem, st>>>(ps, scale, sl);
    }
    count_launch();
    if ((rc = check_launch("simt_fwd"))) return rc;
  }
  return MLSTM_OK;
}

int simt_bwd(const mlstm_params& p, cudaStream_t st, int part) {
  const float scale = resolve_scale(p);
  const size_t rows = (size_t)p.B * p.NH * p.S;
  const int ns = num_slices(p);
  float* ws_dn = reinterpret_cast<float*>(p.workspace);   // [ns][rows]
  float* ws_R = ws_dn + (size_t)ns * rows;
  float* dq_acc = ws_R + rows;
  float* dk_acc = dq_acc + rows * p.DHQK;
  dim3 grid(p.B * p.NH), block(NT);
  int rc;
  if (part != 1) {
    for (int s = 0; s < ns; ++s) {
      Slice sl;
      const mlstm_params ps = slice_params(p, s, &sl, dq_acc, dk_acc);
      const size_t smA = bwd_dq_smem(ps.DHQK, ps.DHV);
      if (p.dtype == MLSTM_F32) {
        if ((rc = set_smem(simt_bwd_dq_kernel<float>, smA))) return rc;
        simt_bwd_dq_kernel<float><<<grid, block, smA, st>>>(ps, scale, ws_dn + (size_t)s * rows, ws_R, sl);
      } else {
        if ((rc = set_smem(simt_bwd_dq_kernel<__nv_bfloat16>, smA))) return rc;
        simt_bwd_dq_kernel<__nv_bfloat16><<<grid, block, smA, st>>>(ps, scale, ws_dn + (size_t)s * rows, ws_R, sl);
      }
      count_launch();
      if ((rc = check_launch("simt_bwd_dq"))) return rc;
    }
  }
  if (part != 0) {
    for (int s = 0; s < ns; ++s) {
      Slice sl;
      const mlstm_params ps = slice_params(p, s, &sl, dq_acc, dk_acc);
      const size_t smB = bwd_dkv_smem(ps.DHQK, ps.DHV);
      if (p.dtype == MLSTM_F32) {
        if ((rc = set_smem(simt_bwd_dkv_kernel<float>, smB))) return rc;
        simt_bwd_dkv_kernel<float><<<grid, block, smB, st>>>(ps, scale, ws_dn + (size_t)s * rows, ws_R, sl);
      } else {
        if ((rc = set_smem(simt_bwd_dkv_kernel<__nv_bfloat16>, smB))) return rc;
        simt_bwd_dkv_kernel<__nv_bfloat16><<<grid, block, smB, st>>>(ps, scale, ws_dn + (size_t)s * rows, ws_R, sl);
      }
      count_launch();
      if ((rc = check_launch("simt_bwd_dkv"))) return rc;
    }
  }
  return MLSTM_OK;
}

}  // namespace mlstm

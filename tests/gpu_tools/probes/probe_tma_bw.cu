// Stand-alone probe: HBM bandwidth reachable by the kernels' TMA tile loads ([128 rows][64 bf16]
// boxes, 128B swizzle) from (B,S,NH,DH) storage (the reference's layout, rows of one head are
// NH*DH*2 bytes apart) versus head-contiguous (B,NH,S,DH) storage.  148 persistent CTAs, each item
// = 3 tensors x 2 tiles x 16 KB, double buffered, no compute.
#include <cstdio>
#include <vector>
#include "tc_ptx.cuh"
#include "tc_tmap.cuh"
using namespace mlstm;
using namespace mlstm::ptx;
constexpr int TILE = 128 * 128;
struct Maps { CUtensorMap a, b, c; };
__global__ void __launch_bounds__(128) bw_kernel(const __grid_constant__ Maps maps, int NH, int NC, int n_items, int depth) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar[2];
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  __syncthreads();
  if (threadIdx.x != 0) return;
  auto issue = [&](int item, int buf) {
    const int bh = item / NC, sc = item % NC, b = bh / NH, h = bh % NH;
    mbar_arrive_expect_tx(&bar[buf], 6 * TILE);
    uint8_t* dst = smem + buf * 6 * TILE;
    for (int kt = 0; kt < 2; ++kt) {
      tma_load_4d(dst + (0 + kt) * TILE, &maps.a, &bar[buf], kt * 64, sc * 128, h, b);
      tma_load_4d(dst + (2 + kt) * TILE, &maps.b, &bar[buf], kt * 64, sc * 128, h, b);
      tma_load_4d(dst + (4 + kt) * TILE, &maps.c, &bar[buf], kt * 64, sc * 128, h, b);
    }
  };
  int n = 0;
  const int item0 = blockIdx.x;
  if (item0 < n_items) issue(item0, 0);
  for (int item = item0; item < n_items; item += gridDim.x, ++n) {
    const int next = item + gridDim.x;
    if (depth > 1 && next < n_items) issue(next, (n + 1) & 1);
    mbar_wait(&bar[n & 1], (n >> 1) & 1);
    if (depth == 1 && next < n_items) issue(next, (n + 1) & 1);
  }
}
int main() {
  const int B = 32, NH = 4, S = 1600, DH = 128, NC = 13;
  const size_t n = (size_t)B * S * NH * DH;
  __nv_bfloat16 *a, *b, *c;
  cudaMalloc(&a, n * 2); cudaMalloc(&b, n * 2); cudaMalloc(&c, n * 2);
  cudaMemset(a, 0, n * 2); cudaMemset(b, 0, n * 2); cudaMemset(c, 0, n * 2);
  char* flush; cudaMalloc(&flush, 256 << 20);
  cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 12 * TILE);
  for (int layout = 0; layout < 2; ++layout) {
    Maps m;
    int r = 0;
    const int64_t sb = (int64_t)S * NH * DH, sh = layout ? (int64_t)S * DH : DH, ss = layout ? DH : NH * DH;
    r |= make_act_tmap(&m.a, a, B, NH, S, DH, sb, sh, ss, 128);
    r |= make_act_tmap(&m.b, b, B, NH, S, DH, sb, sh, ss, 128);
    r |= make_act_tmap(&m.c, c, B, NH, S, DH, sb, sh, ss, 128);
    if (r) { printf("tmap error %d\n", r); return 1; }
    for (int depth = 1; depth <= 2; ++depth) {
      float best = 1e9;
      for (int it = 0; it < 5; ++it) {
        cudaMemset(flush, it, 256 << 20);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        bw_kernel<<<148, 128, 12 * TILE>>>(m, NH, NC, B * NH * NC, depth);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
      }
      if (cudaGetLastError() != cudaSuccess) { printf("cuda error\n"); return 1; }
      const double bytes = 3.0 * B * NH * NC * 128 * DH * 2;
      printf("layout %s  prefetch depth %d : %.1f us  %.0f GB/s\n", layout ? "(B,NH,S,DH)" : "(B,S,NH,DH)", depth, best * 1e3,
             bytes / best / 1e6);
    }
  }
  return 0;
}

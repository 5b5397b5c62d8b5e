// Bring-up probe: tensor-pipe time of the per-tile MMA sequences of the Toeplitz-FIR activation (kind::tf32, A in TMEM).
#include <cstdio>
#include "fir_probe_common.cuh"
using namespace alcm;
struct Args { long long* cyc; int n1, b1, k1, n2, b2, k2, parts, tiles, order; };
__global__ void __launch_bounds__(128) probe(Args p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 64 * 1024 / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot, b = smem_u32(&bar);
  if (tid == 0) {
    const uint32_t id1 = umma_idesc(2, p.n1), id2 = umma_idesc(2, p.n2);
    const uint64_t f1 = umma_desc_kmajor(smem_u32(smem), 128, 1024), f2 = umma_desc_kmajor(smem_u32(smem) + 16384, 128, 2560);
    const long long t0 = clock64();
    for (int t = 0; t < p.tiles; ++t) {
      const uint32_t base = tm + (t & 1) * 224;
      if (p.order == 0) {  // k-step outermost: consecutive MMAs hit different accumulators
        for (int ks = 0; ks < p.k1; ++ks)
          for (int pt = 0; pt < p.parts; ++pt)
            for (int bl = 0; bl < p.b1; ++bl)
              umma_ts_tf32(base + 80 + bl * p.n1, base + (p.n1 / 2) * bl + 8 * ks, f1 + (uint64_t)(16 * ks + 512 * pt), id1, (ks | pt) != 0);
        for (int ks = 0; ks < p.k2; ++ks)
          for (int pt = 0; pt < p.parts; ++pt)
            for (int bl = 0; bl < p.b2; ++bl)
              umma_ts_tf32(base + bl * p.n2, base + 80 + 2 * p.n2 * bl + 8 * ks, f2 + (uint64_t)(16 * ks + 1024 * pt), id2, (ks | pt) != 0);
      } else {             // block outermost: dependent chains
        for (int bl = 0; bl < p.b1; ++bl)
          for (int ks = 0; ks < p.k1; ++ks)
            for (int pt = 0; pt < p.parts; ++pt)
              umma_ts_tf32(base + 80 + bl * p.n1, base + (p.n1 / 2) * bl + 8 * ks, f1 + (uint64_t)(16 * ks + 512 * pt), id1, (ks | pt) != 0);
        for (int bl = 0; bl < p.b2; ++bl)
          for (int ks = 0; ks < p.k2; ++ks)
            for (int pt = 0; pt < p.parts; ++pt)
              umma_ts_tf32(base + bl * p.n2, base + 80 + 2 * p.n2 * bl + 8 * ks, f2 + (uint64_t)(16 * ks + 1024 * pt), id2, (ks | pt) != 0);
      }
    }
    const long long t1 = clock64();
    tc_commit(b);
    mbar_wait(b, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { p.cyc[0] = t1 - t0; p.cyc[1] = t2 - t0; }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}
int main() {
  long long* cyc; cudaMalloc(&cyc, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  // {n1, blocks1, ksteps1, n2, blocks2, ksteps2}
  const int cfg[][6] = {{48, 3, 4, 32, 2, 10}, {48, 3, 4, 16, 4, 6}, {16, 9, 2, 16, 4, 6}, {48, 3, 4, 64, 1, 18}, {144, 1, 10, 64, 1, 18}};
  for (auto& c : cfg)
    for (int parts = 1; parts <= 2; ++parts)
      for (int order = 0; order < 2; ++order) {
        Args a{cyc, c[0], c[1], c[2], c[3], c[4], c[5], parts, 64, order};
        probe<<<148, 128, 64 * 1024>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("ERR %s\n", cudaGetErrorString(e)); return 1; }
        long long h[2];
        cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
        const int nm = (c[1] * c[2] + c[4] * c[5]) * parts;
        printf("up N=%3d x%d blocks x%2d ksteps, down N=%3d x%d x%2d, parts=%d, %s: %3d MMAs/tile, %7.1f cycles/tile issue, %7.1f complete (%.1f per MMA)\n",
               c[0], c[1], c[2], c[3], c[4], c[5], parts, order ? "block-major" : "kstep-major", nm, h[0] / 64.0, h[1] / 64.0, h[1] / 64.0 / nm);
      }
  return 0;
}

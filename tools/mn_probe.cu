// Bring-up probe: which shared-memory element does a kind::tf32 MMA with an MN-major A operand read for (m, k)?
// A = 32 planes x 80 rows x 4 floats filled with the element's own m (pass 0) or row (pass 1); B = identity on k.
#include <cstdio>
#include <vector>
#include "fir_probe_common.cuh"
using namespace alcm;
constexpr int kRows = 80, kPlaneB = kRows * 16;
struct Args { float* d; int lbo, sbo, pass, amn, layout; };
__global__ void __launch_bounds__(128) probe(Args p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  float* X = reinterpret_cast<float*>(smem);
  float* F = reinterpret_cast<float*>(smem + 48 * 1024);   // [16][8] K-major: 2 core matrices along K (LBO 128), n groups SBO 256
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32 * kRows * 4; i += 128) {
    const int pl = i / (kRows * 4), r = (i / 4) % kRows, c = i % 4;
    X[i] = p.pass == 0 ? (float)(pl * 4 + c) : (float)r;
  }
  for (int i = tid; i < 16 * 8; i += 128) {
    const int n = i / 8, k = i % 8;
    F[((n >> 3) * 2 + (k >> 2)) * 32 + (n & 7) * 4 + (k & 3)] = (n == k) ? 1.f : 0.f;
  }
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 32); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot, b = smem_u32(&bar);
  if (tid == 0) {
    const uint32_t id = umma_idesc(2, 16) | (p.amn ? kIdescAMajorMN : 0u);
    const uint64_t ad = umma_desc_kmajor(smem_u32(X) + 8 * 16, p.lbo, p.sbo) | ((uint64_t)p.layout << 61);
    const uint64_t bd = umma_desc_kmajor(smem_u32(F), 128, 256);
    umma_ss<1>(tm, ad, bd, id, 0);
    tc_commit(b);
  }
  mbar_wait(b, 0);
  tc_fence_after();
  uint32_t v[16];
  tmem_ld_x16(tm + ((uint32_t)(warp * 32) << 16), v);
  tmem_ld_wait();
  for (int k = 0; k < 16; ++k) p.d[tid * 16 + k] = __uint_as_float(v[k]);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 32);
}
int main() {
  float* d; cudaMalloc(&d, 128 * 16 * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int variants[][4] = {{128, kPlaneB, 1, 1}, {128, kPlaneB, 1, 2}, {128, kPlaneB, 1, 4}, {128, kPlaneB, 1, 6}, {kPlaneB, 128, 1, 2}, {1024, 2048, 1, 2}, {1024, 2048, 1, 1}};
  for (auto& v : variants) {
    std::vector<float> hm(128 * 16), hr(128 * 16);
    for (int pass = 0; pass < 2; ++pass) {
      Args a{d, v[0], v[1], pass, v[2], v[3]};
      probe<<<1, 128, 64 * 1024>>>(a);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("ERR %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(pass ? hr.data() : hm.data(), d, 128 * 16 * 4, cudaMemcpyDeviceToHost);
    }
    printf("lbo=%d sbo=%d a_mn=%d layout=%d: (m,k) -> (signal,row) read [start row 8]\n", v[0], v[1], v[2], v[3]);
    for (int m : {0, 1, 3, 4, 5, 8, 9, 31, 32, 64, 127}) {
      printf("  m=%3d:", m);
      for (int k = 0; k < 8; ++k) printf(" (%g,%g)", hm[m * 16 + k], hr[m * 16 + k]);
      printf("\n");
    }
  }
  return 0;
}

// Bring-up probe for the tensor-core FIR used by act1d_umma_kernel (not product code):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../audiolcm_b200/csrc fir_probe.cu -o fir_probe
// 1. semantics of a kind::tf32 MMA whose A operand is MN-major in shared memory - the staged image of 32 fp32 planes
//    [plane][row][4 channels], M = 128 signals, K = time rows - against integer test patterns (exact in tf32);
// 2. semantics of a kind::tf32 MMA whose A operand lives in tensor memory (written with tcgen05.st);
// 3. cycles per MMA for the small N tiles the banded (Toeplitz) FIR blocks use.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "fir_probe_common.cuh"

using namespace alcm;

constexpr int kRows = 80, kPlaneB = kRows * 16, kXBytes = 32 * kPlaneB;   // 40 KB
constexpr int kN1 = 48, kK1 = 32, kN2 = 32, kK2 = 80;

__host__ __device__ inline int xval(int m, int r) { return ((m * 7 + r * 13) % 31) - 15; }
__host__ __device__ inline int f1val(int n, int k) { return ((n * 3 + k * 5) % 7) - 3; }
__host__ __device__ inline int f2val(int n, int k) { return ((n * 5 + k * 3) % 5) - 2; }
__host__ __device__ inline int a2val(int m, int c) { return ((m * 3 + c * 11) % 13) - 6; }

// K-major no-swizzle operand [N][K] of fp32: element (n,k) at ((n/8)*(K/4) + k/4)*128 + (n%8)*16 + (k%4)*4
__device__ inline uint32_t foff(int n, int k, int K) { return ((n >> 3) * (K >> 2) + (k >> 2)) * 128 + (n & 7) * 16 + (k & 3) * 4; }

struct Args { float* d1; float* d2; long long* cyc; int N, reps, ts; };

__global__ void __launch_bounds__(128) probe(Args p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  float* X = reinterpret_cast<float*>(smem);
  uint8_t* F1 = smem + kXBytes;
  uint8_t* F2 = F1 + kN1 * kK1 * 4;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 32 * kRows * 4; i += 128) {
    const int pl = i / (kRows * 4), r = (i / 4) % kRows, c = i % 4;
    X[i] = (float)xval(pl * 4 + c, r);
  }
  for (int i = tid; i < kN1 * kK1; i += 128) *reinterpret_cast<float*>(F1 + foff(i / kK1, i % kK1, kK1)) = (float)f1val(i / kK1, i % kK1);
  for (int i = tid; i < kN2 * kK2; i += 128) *reinterpret_cast<float*>(F2 + foff(i / kK2, i % kK2, kK2)) = (float)f2val(i / kK2, i % kK2);
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); fence_mbar_init(); }
  fence_proxy_async_smem();
  if (warp == 0) { tmem_alloc(smem_u32(&slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot, b = smem_u32(&bar);
  const uint32_t lane_t = tm + ((uint32_t)(warp * 32) << 16);
  uint32_t ph = 0;
  // ---- 1: D1[m][n] = sum_k X[m][24 + k] F1[n][k]  (A MN-major: SBO = plane stride, LBO = 128 B per 8 rows) ----
  if (tid == 0) {
    const uint32_t id = umma_idesc(2, kN1) | kIdescAMajorMN;
    for (int ks = 0; ks < kK1 / 8; ++ks) {
      const uint64_t ad = umma_desc_kmajor(smem_u32(X) + (24 + 8 * ks) * 16, 128, kPlaneB);
      const uint64_t bd = umma_desc_kmajor(smem_u32(F1) + ks * 256, 128, (kK1 / 4) * 128);
      umma_ss<1>(tm, ad, bd, id, ks > 0);
    }
    tc_commit(b);
  }
  mbar_wait(b, ph); ph ^= 1;
  tc_fence_after();
  for (int c = 0; c < kN1; c += 16) {
    uint32_t v[16];
    tmem_ld_x16(lane_t + c, v);
    tmem_ld_wait();
    for (int k = 0; k < 16; ++k) p.d1[(size_t)tid * kN1 + c + k] = __uint_as_float(v[k]);
  }
  // ---- 2: A2[m][c] written with tcgen05.st at columns 256.., D2[m][n] = sum_k A2[m][k] F2[n][k] at columns 128.. ----
  for (int c = 0; c < kK2; c += 16) {
    uint32_t v[16];
    for (int k = 0; k < 16; ++k) v[k] = __float_as_uint((float)a2val(tid, c + k));
    tmem_st_x16(lane_t + 256 + c, v);
  }
  tmem_st_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    const uint32_t id = umma_idesc(2, kN2);
    for (int ks = 0; ks < kK2 / 8; ++ks) {
      const uint64_t bd = umma_desc_kmajor(smem_u32(F2) + ks * 256, 128, (kK2 / 4) * 128);
      umma_ts_tf32(tm + 128, tm + 256 + 8 * ks, bd, id, ks > 0);
    }
    tc_commit(b);
  }
  mbar_wait(b, ph); ph ^= 1;
  tc_fence_after();
  for (int c = 0; c < kN2; c += 16) {
    uint32_t v[16];
    tmem_ld_x16(lane_t + 128 + c, v);
    tmem_ld_wait();
    for (int k = 0; k < 16; ++k) p.d2[(size_t)tid * kN2 + c + k] = __uint_as_float(v[k]);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // ---- 3: timing: `reps` MMAs of M=128 x N x K=8, SS (MN-major A) or TS ----
  if (tid == 0) {
    const uint32_t id_ss = umma_idesc(2, p.N) | kIdescAMajorMN, id_ts = umma_idesc(2, p.N);
    const uint64_t bd = umma_desc_kmajor(smem_u32(F2), 128, (kK2 / 4) * 128);
    const long long t0 = clock64();
    for (int i = 0; i < p.reps; ++i) {
      if (p.ts) umma_ts_tf32(tm + 128, tm + 256 + 8 * (i & 7), bd + (uint64_t)(16 * (i & 3)), id_ts, 1);
      else umma_ss<1>(tm + 128, umma_desc_kmajor(smem_u32(X) + (8 * (i & 7)) * 16, 128, kPlaneB), bd + (uint64_t)(16 * (i & 3)), id_ss, 1);
    }
    const long long t1 = clock64();
    tc_commit(b);
    mbar_wait(b, ph);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { p.cyc[0] = t1 - t0; p.cyc[1] = t2 - t0; }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
  float *d1, *d2; long long* cyc;
  cudaMalloc(&d1, 128 * kN1 * 4); cudaMalloc(&d2, 128 * kN2 * 4); cudaMalloc(&cyc, 16);
  const size_t sm = 100 * 1024;   // operands (56 KB) + slack for the N = 128 timing runs reading past F2
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  Args a{d1, d2, cyc, 32, 64, 0};
  probe<<<1, 128, sm>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("ERR %s\n", cudaGetErrorString(e)); return 1; }
  std::vector<float> h1(128 * kN1), h2(128 * kN2);
  cudaMemcpy(h1.data(), d1, h1.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(h2.data(), d2, h2.size() * 4, cudaMemcpyDeviceToHost);
  int bad1 = 0, bad2 = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < kN1; ++n) {
      int s = 0;
      for (int k = 0; k < kK1; ++k) s += xval(m, 24 + k) * f1val(n, k);
      if (h1[m * kN1 + n] != (float)s && bad1++ < 5) printf("D1[%d][%d] = %g, want %d\n", m, n, h1[m * kN1 + n], s);
    }
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < kN2; ++n) {
      int s = 0;
      for (int k = 0; k < kK2; ++k) s += a2val(m, k) * f2val(n, k);
      if (h2[m * kN2 + n] != (float)s && bad2++ < 5) printf("D2[%d][%d] = %g, want %d\n", m, n, h2[m * kN2 + n], s);
    }
  printf("MN-major SS tf32: %d mismatches of %d;  TMEM-A TS tf32: %d mismatches of %d\n", bad1, 128 * kN1, bad2, 128 * kN2);
  for (int ts = 0; ts < 2; ++ts)
    for (int N : {16, 32, 48, 64, 128}) {
      Args t{d1, d2, cyc, N, 512, ts};
      probe<<<148, 128, sm>>>(t);
      e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("ERR %s\n", cudaGetErrorString(e)); return 1; }
      long long h[2];
      cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
      printf("%s N=%3d: %.1f cycles/MMA issue, %.1f until complete (128xNx8 tf32: %.0f FMA/clk)\n", ts ? "TS" : "SS", N, h[0] / 512.0, h[1] / 512.0,
             128.0 * N * 8 * 512 / h[1]);
    }
  return (bad1 || bad2) ? 2 : 0;
}

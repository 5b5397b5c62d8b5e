// Fused anti-aliased activation: UpSample1d(2x Kaiser-sinc FIR) -> SnakeBeta -> DownSample1d.
// Restates vocoder/bigvgan/alias_free_torch/act.py:23-28, resample.py:25-33,46-49,
// filter.py:86-95 and activations.py:107-120 as ONE memory-bound kernel:
//   up   : y[2n]   = 2*sum_{q<6} x[c(n-3+q)] f[11-2q],  y[2n+1] = 2*sum_{q<6} x[c(n-2+q)] f[10-2q]
//   act  : y      += sin^2(y*e^alpha) / (e^beta + 1e-9)
//   down : out[m]  = sum_{k<12} y[clamp(2m+k-5, 0, 2T-1)] f[k]
// c() clamps to [0,T-1] (replicate padding on both FIRs).  Input: fp32 planes (E=4).  Output:
// fp32 planes (optionally RNE-rounded to tf32 for the tf32 MMA) or bf16 planes (E=8, two input
// planes per output plane).  One thread = one time step x one 16-byte output unit: all global
// accesses are 128-bit and a warp touches 512 contiguous bytes per plane.
#pragma once
#include "common.cuh"

namespace alcm {

// kaiser_sinc_filter1d(cutoff=0.25, half_width=0.3, kernel_size=12) - filter.py:28-57; symmetric.
__constant__ float c_fir[12] = {0.0020289647f, 0.0093894657f,  -0.0255434588f, -0.0576573834f, 0.1285725832f, 0.4432097971f,
                                0.4432097971f, 0.1285725832f, -0.0576573834f, -0.0255434588f, 0.0093894657f, 0.0020289647f};

constexpr int kActThreads = 256;

struct ActArgs {
  const float* x;   // fp32 planes
  PlaneGeom xg;
  void* out;        // fp32 or bf16 planes
  PlaneGeom og;
  const float* ea;  // exp(alpha)            [Cpad]
  const float* ib;  // 1/(exp(beta)+1e-9)    [Cpad]
  int T;
  int round_tf32;
};

__device__ __forceinline__ float snake1(float v, float ea, float ib) {
  // sin^2 has period pi: reduce a*v to [-pi/2, pi/2] (2-term Cody-Waite) then MUFU.SIN
  const float t = v * ea;
  const float k = rintf(t * 0.31830988618379067f);
  float r = fmaf(-k, 3.14159274101257324f, t);
  r = fmaf(-k, -8.74227765734758577e-8f, r);
  const float s = __sinf(r);
  return fmaf(ib * s, s, v);
}

template <int NPL, int kActTT>  // NPL input planes per thread: 1 -> fp32 out, 2 -> bf16 out; kActTT steps per block
__global__ void __launch_bounds__(kActThreads) act1d_kernel(const __grid_constant__ ActArgs a) {
  __shared__ float4 sx[NPL][kActTT + 12];
  __shared__ float4 se[NPL][kActTT + 6];
  __shared__ float4 so[NPL][kActTT + 6];
  const int tid = threadIdx.x;
  const int t0 = blockIdx.x * kActTT;
  const int oc = blockIdx.y, b = blockIdx.z;
  const int T = a.T;
  float f[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) f[i] = c_fir[i];

#pragma unroll
  for (int p = 0; p < NPL; ++p) {
    const int chunk = oc * NPL + p;
    const float4* xp = reinterpret_cast<const float4*>(a.x) + ((size_t)b * a.xg.nchunk + chunk) * a.xg.Tp + a.xg.pad;
    for (int i = tid; i < kActTT + 12; i += kActThreads) {
      const int t = min(max(t0 - 6 + i, 0), T - 1);
      sx[p][i] = xp[t];
    }
  }
  __syncthreads();

#pragma unroll
  for (int p = 0; p < NPL; ++p) {
    const int chunk = oc * NPL + p;
    const float4 ea = *reinterpret_cast<const float4*>(a.ea + chunk * 4);
    const float4 ib = *reinterpret_cast<const float4*>(a.ib + chunk * 4);
    for (int i = tid; i < kActTT + 6; i += kActThreads) {
      const int n = t0 - 3 + i;           // up-sampled pair index (unclamped)
      const int nc = min(max(n, 0), T - 1);
      // sx index of x[nc + d] with clamp: position of time t is (t - (t0-6))
      float4 ev = make_float4(0.f, 0.f, 0.f, 0.f), ov = ev;
#pragma unroll
      for (int q = 0; q < 7; ++q) {
        const int t = min(max(nc - 3 + q, 0), T - 1);
        const float4 xv = sx[p][t - (t0 - 6)];
        if (q < 6) {  // even: x[c(n-3+q)] * f[11-2q]
          const float w = f[11 - 2 * q];
          ev.x = fmaf(xv.x, w, ev.x); ev.y = fmaf(xv.y, w, ev.y); ev.z = fmaf(xv.z, w, ev.z); ev.w = fmaf(xv.w, w, ev.w);
        }
        if (q > 0) {  // odd: x[c(n-2+q')] * f[10-2q'], q' = q-1
          const float w = f[12 - 2 * q];
          ov.x = fmaf(xv.x, w, ov.x); ov.y = fmaf(xv.y, w, ov.y); ov.z = fmaf(xv.z, w, ov.z); ov.w = fmaf(xv.w, w, ov.w);
        }
      }
      ev.x = snake1(2.f * ev.x, ea.x, ib.x); ev.y = snake1(2.f * ev.y, ea.y, ib.y);
      ev.z = snake1(2.f * ev.z, ea.z, ib.z); ev.w = snake1(2.f * ev.w, ea.w, ib.w);
      ov.x = snake1(2.f * ov.x, ea.x, ib.x); ov.y = snake1(2.f * ov.y, ea.y, ib.y);
      ov.z = snake1(2.f * ov.z, ea.z, ib.z); ov.w = snake1(2.f * ov.w, ea.w, ib.w);
      // replicate padding of the down filter acts on y: y[j<0] = y[0], y[j>=2T] = y[2T-1]
      if (n < 0) ov = ev;       // both slots = y[0]   (nc == 0)
      if (n >= T) ev = ov;      // both slots = y[2T-1] (nc == T-1)
      se[p][i] = ev;
      so[p][i] = ov;
    }
  }
  __syncthreads();

  for (int i = tid; i < kActTT; i += kActThreads) {
    const int m = t0 + i;
    if (m >= T) break;
    float4 r[NPL];
#pragma unroll
    for (int p = 0; p < NPL; ++p) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      // k odd  -> even slot of n = m + (k-5)/2 ; k even -> odd slot of n = m + (k-6)/2 ; slot idx = n - (t0-3)
#pragma unroll
      for (int k = 0; k < 12; ++k) {
        const int n_rel = (k & 1) ? (i + 3 + (k - 5) / 2) : (i + 3 + (k - 6) / 2);
        const float4 yv = (k & 1) ? se[p][n_rel] : so[p][n_rel];
        const float w = f[k];
        acc.x = fmaf(yv.x, w, acc.x); acc.y = fmaf(yv.y, w, acc.y); acc.z = fmaf(yv.z, w, acc.z); acc.w = fmaf(yv.w, w, acc.w);
      }
      r[p] = acc;
    }
    if (NPL == 1) {
      float4 o = r[0];
      if (a.round_tf32) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
      float4* op = reinterpret_cast<float4*>(a.out) + ((size_t)b * a.og.nchunk + oc) * a.og.Tp + a.og.pad;
      op[m] = o;
    } else {
      __nv_bfloat162 h0 = __floats2bfloat162_rn(r[0].x, r[0].y), h1 = __floats2bfloat162_rn(r[0].z, r[0].w);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(r[NPL - 1].x, r[NPL - 1].y), h3 = __floats2bfloat162_rn(r[NPL - 1].z, r[NPL - 1].w);
      uint4 o;
      o.x = *reinterpret_cast<uint32_t*>(&h0); o.y = *reinterpret_cast<uint32_t*>(&h1);
      o.z = *reinterpret_cast<uint32_t*>(&h2); o.w = *reinterpret_cast<uint32_t*>(&h3);
      uint4* op = reinterpret_cast<uint4*>(a.out) + ((size_t)b * a.og.nchunk + oc) * a.og.Tp + a.og.pad;
      op[m] = o;
    }
  }
}

}  // namespace alcm

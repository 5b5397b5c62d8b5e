// Layout conversion, weight preparation, conv_post+tanh, GroupNorm+swish and the VAE's single
// attention block.  All of these are small memory-bound helpers around the two hot kernels
// (conv.cuh, act1d.cuh).
#pragma once
#include "common.cuh"

namespace alcm {

// ------------------------------------------------------------------------------- layout
// [B][C][T] fp32 channel-first  ->  planes (fp32 E=4, optional tf32 rounding, or bf16 E=8); x*mul
template <int E>
__global__ void pack_cf_kernel(const float* __restrict__ in, void* __restrict__ out, PlaneGeom og, int C, int T, float mul,
                               int rtf32) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  float v[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int c = chunk * E + e;
    v[e] = (c < C) ? in[((size_t)b * C + c) * T + t] * mul : 0.f;
  }
  uint8_t* dst = reinterpret_cast<uint8_t*>(out) + plane_row_off(og, b, chunk, t);
  if (E == 4) {
    if (rtf32) {
#pragma unroll
      for (int e = 0; e < E; ++e) v[e] = round_tf32(v[e]);
    }
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    uint4 o;
    o.x = pack16x2(og.fmt, v[0], v[1]); o.y = pack16x2(og.fmt, v[2], v[3]);
    o.z = pack16x2(og.fmt, v[4 % E], v[5 % E]); o.w = pack16x2(og.fmt, v[6 % E], v[7 % E]);
    *reinterpret_cast<uint4*>(dst) = o;
  }
}

// fp32 planes -> [B][C][T] fp32 channel-first
__global__ void unpack_cf_kernel(const float* __restrict__ in, PlaneGeom ig, float* __restrict__ out, int C, int T) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(in) + plane_row_off(ig, b, chunk, t));
  const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = chunk * 4 + e;
    if (c < C) out[((size_t)b * C + c) * T + t] = vv[e];
  }
}

// 16-bit planes (E=8, bf16 or fp16 by ig.fmt) -> [B][C][T] fp32 channel-first (test entry points only)
__global__ void unpack_cf_bf16_kernel(const uint8_t* __restrict__ in, PlaneGeom ig, float* __restrict__ out, int C, int T) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  const uint4 v = *reinterpret_cast<const uint4*>(in + plane_row_off(ig, b, chunk, t));
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = chunk * 8 + e;
    float f;
    if (ig.fmt == kFmtF16) {
      f = __half2float(__ushort_as_half((unsigned short)((e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xffffu))));
    } else {
      f = __uint_as_float((e & 1) ? (w[e >> 1] & 0xffff0000u) : (w[e >> 1] << 16));
    }
    if (c < C) out[((size_t)b * C + c) * T + t] = f;
  }
}

// fp32 planes -> operand planes: bf16 (E=8, two in-planes per out-plane) or tf32-rounded fp32 copy
template <int E>
__global__ void cast_planes_kernel(const float* __restrict__ in, PlaneGeom ig, void* __restrict__ out, PlaneGeom og, int T) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int oc = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  uint8_t* dst = reinterpret_cast<uint8_t*>(out) + plane_row_off(og, b, oc, t);
  if (E == 8) {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(in) + plane_row_off(ig, b, 2 * oc, t));
    const float4 c = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(in) + plane_row_off(ig, b, 2 * oc + 1, t));
    uint4 o;
    o.x = pack16x2(og.fmt, a.x, a.y); o.y = pack16x2(og.fmt, a.z, a.w); o.z = pack16x2(og.fmt, c.x, c.y); o.w = pack16x2(og.fmt, c.z, c.w);
    *reinterpret_cast<uint4*>(dst) = o;
  } else {
    float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(in) + plane_row_off(ig, b, oc, t));
    a.x = round_tf32(a.x); a.y = round_tf32(a.y); a.z = round_tf32(a.z); a.w = round_tf32(a.w);
    *reinterpret_cast<float4*>(dst) = a;
  }
}

// nn.LayerNorm(C) over the channel axis of a CHANNELS-FIRST tensor x[b][c][t] (the DiT keeps its token stream in the conv
// layout; the reference normalises (B,T,C) tokens, ldm/modules/new_attention.py:246-248).  One thread per (b, t) column,
// threads along t (coalesced rows); mean, then centred variance, then the affine output: the second and third sweeps of
// a column tile (C x 128 x 4 bytes) come from L2.
__global__ void __launch_bounds__(128) layernorm_cf_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float* __restrict__ y, int C, int T, float eps) {
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t >= T) return;
  const float* xp = x + (size_t)blockIdx.y * C * T + t;
  float* yp = y + (size_t)blockIdx.y * C * T + t;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int c = 0;
  for (; c + 4 <= C; c += 4) {
    s0 += xp[(size_t)c * T]; s1 += xp[(size_t)(c + 1) * T]; s2 += xp[(size_t)(c + 2) * T]; s3 += xp[(size_t)(c + 3) * T];
  }
  for (; c < C; ++c) s0 += xp[(size_t)c * T];
  const float mean = ((s0 + s1) + (s2 + s3)) / (float)C;
  s0 = s1 = s2 = s3 = 0.f;
  for (c = 0; c + 4 <= C; c += 4) {
    const float d0 = xp[(size_t)c * T] - mean, d1 = xp[(size_t)(c + 1) * T] - mean, d2 = xp[(size_t)(c + 2) * T] - mean,
                d3 = xp[(size_t)(c + 3) * T] - mean;
    s0 = fmaf(d0, d0, s0); s1 = fmaf(d1, d1, s1); s2 = fmaf(d2, d2, s2); s3 = fmaf(d3, d3, s3);
  }
  for (; c < C; ++c) { const float d = xp[(size_t)c * T] - mean; s0 = fmaf(d, d, s0); }
  const float rstd = rsqrtf(((s0 + s1) + (s2 + s3)) / (float)C + eps);
#pragma unroll 4
  for (c = 0; c < C; ++c) yp[(size_t)c * T] = (xp[(size_t)c * T] - mean) * rstd * gamma[c] + beta[c];
}

// GEGLU between the two convs of the DiT's Conv1dFeedForward (ldm/modules/new_attention.py:48-55):
// out[c] = in[c] * gelu(in[inner + c]) with the exact (erf) GELU of F.gelu; fp32 planes of 2*inner channels ->
// operand planes of inner channels (bf16 E=8, or fp32 E=4 optionally rounded to tf32).  inner % E == 0.
__device__ __forceinline__ float geglu1(float v, float g) { return v * (0.5f * g * (1.f + erff(g * 0.70710678118654752f))); }
// OP 1 (the mel front-end, NAT_mel.py:75-78): out[c] = sqrt(in[c]^2 + in[inner + c]^2 + 1e-9), the magnitude of an STFT whose
// real parts are channels [0, inner) and imaginary parts [inner, 2*inner).
template <int OP>
__device__ __forceinline__ float pair_op(float v, float g) { return OP == 0 ? geglu1(v, g) : sqrtf(fmaf(v, v, fmaf(g, g, 1e-9f))); }
template <int E, int OP>
__global__ void geglu_planes_kernel(const float* __restrict__ in, PlaneGeom ig, void* __restrict__ out, PlaneGeom og, int T, int inner,
                                    int round_tf) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int oc = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  const uint8_t* src = reinterpret_cast<const uint8_t*>(in);
  uint8_t* dst = reinterpret_cast<uint8_t*>(out) + plane_row_off(og, b, oc, t);
  const int gch = inner / 4;   // first gate chunk
  if (E == 8) {
    const float4 v0 = *reinterpret_cast<const float4*>(src + plane_row_off(ig, b, 2 * oc, t));
    const float4 v1 = *reinterpret_cast<const float4*>(src + plane_row_off(ig, b, 2 * oc + 1, t));
    const float4 g0 = *reinterpret_cast<const float4*>(src + plane_row_off(ig, b, gch + 2 * oc, t));
    const float4 g1 = *reinterpret_cast<const float4*>(src + plane_row_off(ig, b, gch + 2 * oc + 1, t));
    uint4 o;
    o.x = pack16x2(og.fmt, pair_op<OP>(v0.x, g0.x), pair_op<OP>(v0.y, g0.y)); o.y = pack16x2(og.fmt, pair_op<OP>(v0.z, g0.z), pair_op<OP>(v0.w, g0.w));
    o.z = pack16x2(og.fmt, pair_op<OP>(v1.x, g1.x), pair_op<OP>(v1.y, g1.y)); o.w = pack16x2(og.fmt, pair_op<OP>(v1.z, g1.z), pair_op<OP>(v1.w, g1.w));
    *reinterpret_cast<uint4*>(dst) = o;
  } else {
    const float4 v = *reinterpret_cast<const float4*>(src + plane_row_off(ig, b, oc, t));
    const float4 g = *reinterpret_cast<const float4*>(src + plane_row_off(ig, b, gch + oc, t));
    float4 o = make_float4(pair_op<OP>(v.x, g.x), pair_op<OP>(v.y, g.y), pair_op<OP>(v.z, g.z), pair_op<OP>(v.w, g.w));
    if (round_tf) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
    *reinterpret_cast<float4*>(dst) = o;
  }
}

// Mel front-end input (ldm/data/preprocess/NAT_mel.py:68-73): clamp the waveform to [-1,1], reflect-pad it by `padw` samples
// on both sides and fold it into hop-sized rows, X[b][c][n] = yp[hop*n + c], written as operand planes (C = hop channels,
// T = rows): the STFT is then a stride-1 conv over n (audiolcm_b200/melspec.py).
template <int E>
__global__ void mel_fold_kernel(const float* __restrict__ y, int L, int hop, int padw, void* __restrict__ out, PlaneGeom og, int rows,
                                int rtf32) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y, b = blockIdx.z;
  if (n >= rows) return;
  float v[E];
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int c = chunk * E + e;
    int i = hop * n + c - padw;
    i = i < 0 ? -i : (i >= L ? 2 * (L - 1) - i : i);
    v[e] = (c < hop) ? fminf(fmaxf(y[(size_t)b * L + i], -1.f), 1.f) : 0.f;
  }
  uint8_t* dst = reinterpret_cast<uint8_t*>(out) + plane_row_off(og, b, chunk, n);
  if (E == 4) {
    if (rtf32) {
#pragma unroll
      for (int e = 0; e < E; ++e) v[e] = round_tf32(v[e]);
    }
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    uint4 o;
    o.x = pack16x2(og.fmt, v[0], v[1]); o.y = pack16x2(og.fmt, v[2], v[3]);
    o.z = pack16x2(og.fmt, v[4 % E], v[5 % E]); o.w = pack16x2(og.fmt, v[6 % E], v[7 % E]);
    *reinterpret_cast<uint4*>(dst) = o;
  }
}

// fp32 planes -> [B][C][T] channel-first, rows [t0, t0+T), as log10(max(x, floor)) (NAT_mel.py:83-84)
__global__ void unpack_log10_kernel(const float* __restrict__ in, PlaneGeom ig, float* __restrict__ out, int C, int T, int t0, float floor_) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int chunk = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(in) + plane_row_off(ig, b, chunk, t + t0));
  const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int c = chunk * 4 + e;
    if (c < C) out[((size_t)b * C + c) * T + t] = log10f(fmaxf(vv[e], floor_));
  }
}

// n-way sum of fp32 planes -> fp32 planes and/or operand planes (bf16 E=8 or tf32-rounded fp32).
// One thread handles two adjacent fp32 chunks (= one bf16 chunk) of one time step.
struct SumArgs {
  const float* in[4];
  int n;
  PlaneGeom g;       // geometry of the fp32 inputs and of out32
  float* out32;      // may be null
  void* out_op;      // may be null
  PlaneGeom og;      // geometry of out_op
  int op_bf16, round_tf32, T;
};
__global__ void sum_planes_kernel(SumArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int pc = blockIdx.y, b = blockIdx.z;  // pair of fp32 chunks 2*pc, 2*pc+1
  if (t >= a.T) return;
  float4 s[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const size_t off = plane_row_off(a.g, b, 2 * pc + h, t);
    float4 acc = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(a.in[0]) + off);
    for (int i = 1; i < a.n; ++i) {
      const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(a.in[i]) + off);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    s[h] = acc;
    if (a.out32) *reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(a.out32) + off) = acc;
  }
  if (a.out_op) {
    if (a.op_bf16) {
      uint4 o;
      o.x = pack16x2(a.og.fmt, s[0].x, s[0].y); o.y = pack16x2(a.og.fmt, s[0].z, s[0].w);
      o.z = pack16x2(a.og.fmt, s[1].x, s[1].y); o.w = pack16x2(a.og.fmt, s[1].z, s[1].w);
      *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(a.out_op) + plane_row_off(a.og, b, pc, t)) = o;
    } else {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float4 v = s[h];
        if (a.round_tf32) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
        *reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(a.out_op) + plane_row_off(a.og, b, 2 * pc + h, t)) = v;
      }
    }
  }
}

// Downsample1D (autoencoder1d.py:296-316: right zero pad + Conv1d k3 stride 2) reads x at 2n, 2n+1, 2n+2.  With the time
// axis folded into channels - x2[parity*C + c][n] = x[c][2n + parity] - it is a 2-tap stride-1 conv (rows n, n+1) over 2C
// channels, which conv_umma_kernel handles.  This kernel writes x2 as operand planes (bf16 cast / tf32 rounding
// included).  C must be a multiple of E.
template <int E>
__global__ void s2d_cast_kernel(const float* __restrict__ in, PlaneGeom ig, void* __restrict__ out, PlaneGeom og, int C, int Tout,
                                int rtf32) {
  pdl_launch_dependents();
  pdl_wait();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int oc = blockIdx.y, b = blockIdx.z;
  if (n >= Tout) return;
  const int c0 = oc * E, parity = c0 / C, c = c0 - parity * C, t = 2 * n + parity;
  uint8_t* dst = reinterpret_cast<uint8_t*>(out) + plane_row_off(og, b, oc, n);
  const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(in) + plane_row_off(ig, b, c >> 2, t));
  if (E == 8) {
    const float4 d = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(in) + plane_row_off(ig, b, (c >> 2) + 1, t));
    uint4 o;
    o.x = pack16x2(og.fmt, a.x, a.y); o.y = pack16x2(og.fmt, a.z, a.w); o.z = pack16x2(og.fmt, d.x, d.y); o.w = pack16x2(og.fmt, d.z, d.w);
    *reinterpret_cast<uint4*>(dst) = o;
  } else {
    float4 v = a;
    if (rtf32) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
    *reinterpret_cast<float4*>(dst) = v;
  }
}

// ------------------------------------------------------------------------------- weights
// Effective weights of Downsample1D on the folded input (see s2d_cast_kernel): src (Cout,C,3) -> dst [2 taps][Cout][2C]
//   tap 0 (row n)  : [ w[.,.,0] | w[.,.,1] ]      tap 1 (row n+1): [ w[.,.,2] | 0 ]
__global__ void weff_down2_kernel(const float* __restrict__ src, float* __restrict__ dst, int Cout, int C) {
  const size_t n = (size_t)2 * Cout * 2 * C;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int ci2 = idx % (2 * C);
    const int co = (idx / (2 * C)) % Cout;
    const int tap = idx / ((size_t)2 * C * Cout);
    const int parity = ci2 / C, ci = ci2 - parity * C;
    const int k = tap == 0 ? parity : (parity == 0 ? 2 : -1);
    dst[idx] = k >= 0 ? src[((size_t)co * C + ci) * 3 + k] : 0.f;
  }
}

// weight_norm fold (torch.nn.utils.weight_norm dim=0; models.py:36-51,143,152,174):
// w[i,:,:] = g[i] * v[i,:,:] / ||v[i,:,:]||.  One block per dim-0 slice.
__global__ void wn_fold_kernel(const float* __restrict__ v, const float* __restrict__ g, float* __restrict__ w, int inner) {
  __shared__ double red[256];
  const int i = blockIdx.x;
  const float* vi = v + (size_t)i * inner;
  double s = 0.0;
  for (int k = threadIdx.x; k < inner; k += blockDim.x) s += (double)vi[k] * (double)vi[k];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  const float scale = g[i] / (float)sqrt(red[0]);
  for (int k = threadIdx.x; k < inner; k += blockDim.x) w[(size_t)i * inner + k] = vi[k] * scale;
}

// Effective per-(phase,tap) weight matrices Weff[phase][tap][Cout][Cin] (fp32) from a folded conv
// weight.  transposed=0: src is Conv1d (Cout,Cin,K); transposed=1: ConvTranspose1d (Cin,Cout,K).
// src_k[phase][tap][2]: up to two source taps summed (-1 = none) - the sum is used by the
// nearest-2x-upsample + k3 conv polyphase form (autoencoder1d.py:291-295).
struct WeffRecipe {
  int nphase, ntaps;
  int src_k[kMaxPhase][kMaxTaps][2];
};
__global__ void weff_kernel(const float* __restrict__ src, float* __restrict__ dst, WeffRecipe r, int Cout, int Cin, int K,
                            int transposed) {
  const size_t n = (size_t)r.nphase * r.ntaps * Cout * Cin;
  for (size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int ci = idx % Cin;
    const int co = (idx / Cin) % Cout;
    const int tp = (idx / ((size_t)Cin * Cout)) % r.ntaps;
    const int ph = idx / ((size_t)Cin * Cout * r.ntaps);
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int k = r.src_k[ph][tp][u];
      if (k >= 0) s += transposed ? src[((size_t)ci * Cout + co) * K + k] : src[((size_t)co * Cin + ci) * K + k];
    }
    dst[idx] = s;
  }
}

// Weff -> UMMA weight blobs [phase][n_tile][kb][tap][kc][NT][16B] (the smem image of a K-major,
// no-swizzle B operand: LBO = NT*16, SBO = 128).  One thread per 16-byte unit.
template <int E>
__global__ void pack_w_kernel(const float* __restrict__ weff, void* __restrict__ dst, int nphase, int ntaps, int Cout, int Cin,
                              int NT, int n_tiles, int kblk, int nkb, int f16) {
  const size_t units = (size_t)nphase * n_tiles * nkb * ntaps * kblk * NT;
  for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < units; u += (size_t)gridDim.x * blockDim.x) {
    size_t r = u;
    const int n = r % NT; r /= NT;
    const int kc = r % kblk; r /= kblk;
    const int tp = r % ntaps; r /= ntaps;
    const int kb = r % nkb; r /= nkb;
    const int nt = r % n_tiles; r /= n_tiles;
    const int ph = (int)r;
    const int co = nt * NT + n;
    float v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int ci = (kb * kblk + kc) * E + e;
      v[e] = (co < Cout && ci < Cin) ? weff[(((size_t)ph * ntaps + tp) * Cout + co) * Cin + ci] : 0.f;
    }
    uint8_t* d = reinterpret_cast<uint8_t*>(dst) + u * 16;
    if (E == 4) {
      *reinterpret_cast<float4*>(d) = make_float4(round_tf32(v[0]), round_tf32(v[1]), round_tf32(v[2]), round_tf32(v[3]));
    } else {
      uint4 o;
      o.x = pack16x2(f16 ? kFmtF16 : kFmtBF16, v[0], v[1]); o.y = pack16x2(f16 ? kFmtF16 : kFmtBF16, v[2], v[3]);
      o.z = pack16x2(f16 ? kFmtF16 : kFmtBF16, v[4 % E], v[5 % E]); o.w = pack16x2(f16 ? kFmtF16 : kFmtBF16, v[6 % E], v[7 % E]);
      *reinterpret_cast<uint4*>(d) = o;
    }
  }
}

// Re-tile packed weight blobs to another N tile (pure permutation of 16-byte units): launches whose output
// has few time tiles get narrower N tiles (more CTAs, smaller split-K fix-up) without keeping the fp32 weights.
// Units of output channels >= n_tiles1*NT1 are zero.
__global__ void repack_nt_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int nphase, int ntaps, int kblk, int nkb,
                                 int NT1, int n_tiles1, int NT2, int n_tiles2) {
  const size_t units = (size_t)nphase * n_tiles2 * nkb * ntaps * kblk * NT2;
  for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < units; u += (size_t)gridDim.x * blockDim.x) {
    size_t r = u;
    const int n = r % NT2; r /= NT2;
    const int kc = r % kblk; r /= kblk;
    const int tp = r % ntaps; r /= ntaps;
    const int kb = r % nkb; r /= nkb;
    const int nt = r % n_tiles2; r /= n_tiles2;
    const int ph = (int)r;
    const int co = nt * NT2 + n;
    const int nt1 = co / NT1, n1 = co - nt1 * NT1;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (nt1 < n_tiles1) v = src[(((((size_t)ph * n_tiles1 + nt1) * nkb + kb) * ntaps + tp) * kblk + kc) * NT1 + n1];
    dst[u] = v;
  }
}

// Snake / SnakeBeta parameters (activations.py:56-62,113-120): ea = exp(alpha), ib = 1/(exp(beta)+1e-9) with
// alpha_logscale, ea = alpha, ib = 1/(beta+1e-9) without.  Snake is the beta = alpha case (the caller passes alpha twice).
__global__ void snake_params_kernel(const float* __restrict__ alpha, const float* __restrict__ beta, float* __restrict__ ea,
                                    float* __restrict__ ib, int C, int Cpad, int linear) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cpad) return;
  if (c >= C) { ea[c] = 1.f; ib[c] = 1.f; return; }
  ea[c] = linear ? alpha[c] : expf(alpha[c]);
  ib[c] = 1.0f / ((linear ? beta[c] : expf(beta[c])) + 1e-9f);
}

// ------------------------------------------------------------------------------- conv_post + tanh
// models.py:199-201: Conv1d(C -> 1, k=7, p=3) + tanh on fp32 planes; w is [7][nchunk*4] (tap-major).
// PCM16: the sample is written as 16-bit PCM, rint(x * 32767) - what soundfile.write(path, wav, sr) stores for a
// float waveform (libsndfile's default float -> PCM_16 conversion; pythonscripts/InferAPI.py:98) - so the WAV payload
// leaves the GPU ready to be written and the device->host copy is half the size.
template <bool PCM16>
__global__ void conv_post_tanh_kernel(const float* __restrict__ x, PlaneGeom xg, const float* __restrict__ w, float bias,
                                      void* __restrict__ wav, int T, int ktaps) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float sw[];
  const int nw = ktaps * xg.nchunk * 4;
  for (int i = threadIdx.x; i < nw; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (t >= T) return;
  float acc = bias;
  const int half = ktaps / 2;
  for (int ch = 0; ch < xg.nchunk; ++ch) {
    const float4* xp = reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(x) + plane_row_off(xg, b, ch, 0));
    for (int j = 0; j < ktaps; ++j) {
      const float4 v = xp[t + j - half];  // zero rows of the plane supply the conv padding
      const float* wj = sw + (j * xg.nchunk + ch) * 4;
      acc = fmaf(v.x, wj[0], acc); acc = fmaf(v.y, wj[1], acc); acc = fmaf(v.z, wj[2], acc); acc = fmaf(v.w, wj[3], acc);
    }
  }
  const float y = tanhf(acc);
  if (PCM16) reinterpret_cast<short*>(wav)[(size_t)b * T + t] = (short)__float2int_rn(y * 32767.0f);
  else reinterpret_cast<float*>(wav)[(size_t)b * T + t] = y;
}

// ------------------------------------------------------------------------------- attention on the tensor cores
// AttnBlock1D (autoencoder1d.py:257-278) as two conv_umma_kernel launches with PER-ITEM "weights":
//   S[i][j] = C^-0.5 * sum_c q[c][i] k[c][j]   = a 1x1 conv over queries i whose output channels are the keys j,  W[j][c] = k[c][j]
//   h[c][i] = sum_j P[i][j] v[c][j]            = a 1x1 conv over queries i whose input channels are the keys j,   W[c][j] = v[c][j]
// pack_dyn_w_kernel writes such a per-item B operand (the same blob layout as pack_w_kernel, one tap) from fp32 planes:
//   mode 0: W[n][k] = src[b][channel k][time n]   (K of QK^T: n = key, k = channel)
//   mode 1: W[n][k] = src[b][channel n][time k]   (V of PV:   n = channel, k = key)
template <int E>
__global__ void pack_dyn_w_kernel(const float* __restrict__ src, PlaneGeom sg, int C, int T, void* __restrict__ dst, int mode, int NT,
                                  int n_tiles, int kblk, int nkb, size_t units_per_item, int f16) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int N = mode == 0 ? T : C, Kd = mode == 0 ? C : T;
  for (size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x; u < units_per_item; u += (size_t)gridDim.x * blockDim.x) {
    size_t r = u;
    const int n = r % NT; r /= NT;
    const int kc = r % kblk; r /= kblk;
    const int kb = r % nkb; r /= nkb;
    const int nt = (int)r;
    const int co = nt * NT + n;
    float v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
      const int ci = (kb * kblk + kc) * E + e;
      float x = 0.f;
      if (co < N && ci < Kd) {
        const int ch = mode == 0 ? ci : co, t = mode == 0 ? co : ci;
        x = *(reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(src) + plane_row_off(sg, b, ch >> 2, t)) + (ch & 3));
      }
      v[e] = x;
    }
    uint8_t* d = reinterpret_cast<uint8_t*>(dst) + ((size_t)b * units_per_item + u) * 16;
    if (E == 4) {
      *reinterpret_cast<float4*>(d) = make_float4(round_tf32(v[0]), round_tf32(v[1]), round_tf32(v[2]), round_tf32(v[3]));
    } else {
      uint4 o;
      o.x = pack16x2(f16 ? kFmtF16 : kFmtBF16, v[0], v[1]); o.y = pack16x2(f16 ? kFmtF16 : kFmtBF16, v[2], v[3]);
      o.z = pack16x2(f16 ? kFmtF16 : kFmtBF16, v[4 % E], v[5 % E]); o.w = pack16x2(f16 ? kFmtF16 : kFmtBF16, v[6 % E], v[7 % E]);
      *reinterpret_cast<uint4*>(d) = o;
    }
  }
}

// P = softmax over the keys (channels j < T) of the score planes S[b][j/4][i][4] (row = query i), written as operand planes
// P[b][j/E][i][E] for the PV conv; channels j >= T (padding) are written as zeros.  One thread per query.
template <int E>
__global__ void softmax_planes_kernel(const float* __restrict__ S, PlaneGeom sg, void* __restrict__ P, PlaneGeom pg, int T, int rtf32) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
  if (i >= T) return;
  const int nc = (T + 3) >> 2;
  auto ld = [&](int c) { return *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(S) + plane_row_off(sg, b, c, i)); };
  float m = -INFINITY;
  for (int c = 0; c < nc; ++c) {
    const float4 v = ld(c);
    const int j = 4 * c;
    m = fmaxf(m, v.x);
    if (j + 1 < T) m = fmaxf(m, v.y);
    if (j + 2 < T) m = fmaxf(m, v.z);
    if (j + 3 < T) m = fmaxf(m, v.w);
  }
  float sum = 0.f;
  for (int c = 0; c < nc; ++c) {
    const float4 v = ld(c);
    const int j = 4 * c;
    sum += expf(v.x - m);
    if (j + 1 < T) sum += expf(v.y - m);
    if (j + 2 < T) sum += expf(v.z - m);
    if (j + 3 < T) sum += expf(v.w - m);
  }
  const float inv = 1.f / sum;
  for (int pc = 0; pc < pg.nchunk; ++pc) {
    float o[E];
#pragma unroll
    for (int h = 0; h < E / 4; ++h) {
      const int c = pc * (E / 4) + h;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c < nc) v = ld(c);
      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) o[4 * h + e] = (4 * c + e < T) ? expf(vv[e] - m) * inv : 0.f;
    }
    uint8_t* d = reinterpret_cast<uint8_t*>(P) + plane_row_off(pg, b, pc, i);
    if (E == 4) {
      if (rtf32) {
#pragma unroll
        for (int e = 0; e < E; ++e) o[e] = round_tf32(o[e]);
      }
      *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      uint4 w;
      w.x = pack16x2(pg.fmt, o[0], o[1]); w.y = pack16x2(pg.fmt, o[2], o[3]);
      w.z = pack16x2(pg.fmt, o[4 % E], o[5 % E]); w.w = pack16x2(pg.fmt, o[6 % E], o[7 % E]);
      *reinterpret_cast<uint4*>(d) = w;
    }
  }
}

// ------------------------------------------------------------------------------- LCM sampler step
// LCMSampler.step (ldm/models/diffusion/scheduling_lcm.py:411-494), epsilon prediction, as ONE elementwise pass:
//   x0       = (sample - sqrt(1-abar_t) * eps) / sqrt(abar_t)                      (:455-456)
//   denoised = c_out * x0 + c_skip * sample                                        (:469)
//   prev     = sqrt(abar_prev) * denoised + sqrt(1-abar_prev) * noise   (not on the last step: prev = denoised)  (:474-478)
// The reference runs this as ~8 separate ATen kernels per step; the coefficients are host scalars of the schedule.
struct LcmStepCoef { float b_t_sqrt, inv_a_t_sqrt, c_out, c_skip, a_prev_sqrt, b_prev_sqrt; int last; };
__global__ void lcm_step_kernel(const float4* __restrict__ sample, const float4* __restrict__ eps, const float4* __restrict__ noise,
                                float4* __restrict__ prev, float4* __restrict__ denoised, size_t n4, LcmStepCoef k) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    const float4 s = sample[i], e = eps[i];
    float4 d;
    d.x = k.c_out * ((s.x - k.b_t_sqrt * e.x) * k.inv_a_t_sqrt) + k.c_skip * s.x;
    d.y = k.c_out * ((s.y - k.b_t_sqrt * e.y) * k.inv_a_t_sqrt) + k.c_skip * s.y;
    d.z = k.c_out * ((s.z - k.b_t_sqrt * e.z) * k.inv_a_t_sqrt) + k.c_skip * s.z;
    d.w = k.c_out * ((s.w - k.b_t_sqrt * e.w) * k.inv_a_t_sqrt) + k.c_skip * s.w;
    denoised[i] = d;
    if (k.last) {
      prev[i] = d;
    } else {
      const float4 z = noise[i];
      prev[i] = make_float4(k.a_prev_sqrt * d.x + k.b_prev_sqrt * z.x, k.a_prev_sqrt * d.y + k.b_prev_sqrt * z.y,
                            k.a_prev_sqrt * d.z + k.b_prev_sqrt * z.z, k.a_prev_sqrt * d.w + k.b_prev_sqrt * z.w);
    }
  }
}

// ALCM_GUARD self-check: number of non-zero bytes in a guard zone (n 16-byte units)
__global__ void count_nonzero_kernel(const uint4* __restrict__ p, size_t n, unsigned long long* __restrict__ out) {
  unsigned long long c = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const uint4 v = p[i];
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int b = 0; b < 4; ++b) c += ((w[k] >> (8 * b)) & 0xffu) != 0u;
  }
  if (c) atomicAdd(out, c);
}

// Seeded uniform fill in [lo, hi) (micro-benchmark operands): counter-based hash, fp32 or bf16 elements.
__global__ void fill_uniform_kernel(void* __restrict__ p, size_t n, int fmt, float lo, float hi, unsigned seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    unsigned h = (unsigned)i * 2654435761u ^ (unsigned)(i >> 32) * 40503u ^ seed * 0x9E3779B9u;
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    const float u = (float)(h >> 8) * (1.0f / 16777216.0f);
    const float v = lo + (hi - lo) * u;
    if (fmt == kFmtBF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16(v);
    else if (fmt == kFmtF16) reinterpret_cast<__half*>(p)[i] = __float2half(v);
    else reinterpret_cast<float*>(p)[i] = v;
  }
}

// ------------------------------------------------------------------------------- GroupNorm
// Normalize = GroupNorm(32, C, eps 1e-6, affine) (autoencoder1d.py:169-170); biased variance over
// (C/32 channels x T).  stats[b][g] = {mean, rstd}.  One block per (group, b).
__device__ __forceinline__ double gn_block_sum(float s, double* red) {
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __syncthreads();  // red[] may still be read from the previous pass
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = (double)s;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
  return t;
}

__global__ void __launch_bounds__(512) gn_stats_kernel(const float* __restrict__ x, PlaneGeom xg, int C, int T, int groups,
                                                         float eps, float2* __restrict__ stats) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[16];
  const int g = blockIdx.x, b = blockIdx.y;
  const int cpg = C / groups;
  const double n = (double)cpg * (double)T;
  const bool vec = (cpg & 3) == 0;  // the group is a whole number of 4-channel planes -> float4 loads
  // two passes (mean, then centred sum of squares): E[x^2]-mean^2 cancels badly for small groups.
  // Per-thread partial sums stay short (n / 512 terms), the cross-thread combination is in double.
  float mean = 0.f;
  double dmean = 0.0;
  for (int pass = 0; pass < 2; ++pass) {
    float s = 0.f;
    if (vec) {
      const int npl = cpg >> 2;
      for (int pl = 0; pl < npl; ++pl) {
        const float4* xp = reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(x) + plane_row_off(xg, b, (g * cpg >> 2) + pl, 0));
        float sp = 0.f;
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
          const float4 v = xp[t];
          if (pass) {
            const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
            sp += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
          } else {
            sp += (v.x + v.y) + (v.z + v.w);
          }
        }
        s += sp;
      }
    } else {
      for (int cc = 0; cc < cpg; ++cc) {
        const int c = g * cpg + cc;
        const float* xp = reinterpret_cast<const float*>(reinterpret_cast<const uint8_t*>(x) + plane_row_off(xg, b, c >> 2, 0)) + (c & 3);
        float sp = 0.f;
        for (int t = threadIdx.x; t < T; t += blockDim.x) {
          const float d = xp[(size_t)t * 4] - mean;
          sp += pass ? d * d : d;
        }
        s += sp;
      }
    }
    const double tot = gn_block_sum(s, red);
    if (pass == 0) {
      dmean = tot / n;
      mean = (float)dmean;
    } else if (threadIdx.x == 0) {
      // the second pass centred on the rounded mean: var = E[(x-m)^2] - (mean-m)^2, the correction is ~1e-16
      stats[b * groups + g] = make_float2(mean, (float)(1.0 / sqrt(tot / n + (double)eps)));
    }
  }
}

// y = (x-mean)*rstd*gamma + beta, optional swish (x*sigmoid(x), autoencoder1d.py:172-174), written
// as operand planes (E=4 fp32/tf32 or E=8 bf16).
template <int E>
__global__ void gn_apply_kernel(const float* __restrict__ x, PlaneGeom xg, void* __restrict__ out, PlaneGeom og, int C, int T,
                                int groups, const float2* __restrict__ stats, const float* __restrict__ gamma,
                                const float* __restrict__ beta, int swish, int rtf32) {
  pdl_launch_dependents();
  pdl_wait();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int oc = blockIdx.y, b = blockIdx.z;
  if (t >= T) return;
  const int cpg = C / groups;
  float v[E];
#pragma unroll
  for (int h = 0; h < E / 4; ++h) {
    const float4 a = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(x) + plane_row_off(xg, b, oc * (E / 4) + h, t));
    v[4 * h] = a.x; v[4 * h + 1] = a.y; v[4 * h + 2] = a.z; v[4 * h + 3] = a.w;
  }
#pragma unroll
  for (int e = 0; e < E; ++e) {
    const int c = oc * E + e;
    float y = 0.f;
    if (c < C) {
      const float2 st = stats[b * groups + c / cpg];
      y = (v[e] - st.x) * st.y * gamma[c] + beta[c];
      if (swish) y = y / (1.f + __expf(-y));
    }
    v[e] = y;
  }
  uint8_t* dst = reinterpret_cast<uint8_t*>(out) + plane_row_off(og, b, oc, t);
  if (E == 4) {
    if (rtf32) {
#pragma unroll
      for (int e = 0; e < E; ++e) v[e] = round_tf32(v[e]);
    }
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    uint4 o;
    o.x = pack16x2(og.fmt, v[0], v[1]); o.y = pack16x2(og.fmt, v[2], v[3]);
    o.z = pack16x2(og.fmt, v[4 % E], v[5 % E]); o.w = pack16x2(og.fmt, v[6 % E], v[7 % E]);
    *reinterpret_cast<uint4*>(dst) = o;
  }
}

// Single-launch GroupNorm(+swish) for activations that fit in shared memory (every GroupNorm of the 10 s
// clip does): one block owns `gpb` consecutive groups (gpb*cpg channels = a whole number of output units),
// stages them once (npl fp32 planes x T rows), takes mean and centred variance from shared memory and writes
// the normalised operand planes.  Same arithmetic as gn_stats_kernel + gn_apply_kernel, one HBM/L2 read
// instead of three and one launch instead of two.  dynamic smem = gpb*(cpg/4)*T*16 bytes.
template <int E>
__global__ void __launch_bounds__(512) gn_fused_kernel(const float* __restrict__ x, PlaneGeom xg, void* __restrict__ out, PlaneGeom og,
                                                         int C, int T, int groups, float eps, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, int swish, int rtf32, int gpb) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ float4 sg[];
  __shared__ double red[16];
  __shared__ float s_mean[4], s_rstd[4];
  const int gb = blockIdx.x, b = blockIdx.y;
  const int cpg = C / groups, nplg = cpg >> 2, npl = gpb * nplg;
  const int plane0 = gb * gpb * nplg;
  {  // stage the block's planes: 8 independent loads in flight per thread before the first shared-memory store
    const float4* xp0 = reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(x) + plane_row_off(xg, b, plane0, 0));
    const int total = npl * T;
    for (int i0 = threadIdx.x; i0 < total; i0 += 8 * blockDim.x) {
      float4 v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = i0 + k * blockDim.x;
        if (i < total) {
          const int pl = i / T, t = i - pl * T;
          v[k] = xp0[(size_t)pl * xg.Tp + t];
        }
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int i = i0 + k * blockDim.x;
        if (i < total) sg[i] = v[k];
      }
    }
  }
  __syncthreads();
  const double n = (double)cpg * (double)T;
  for (int g = 0; g < gpb; ++g) {
    const float4* gp = sg + g * nplg * T;
    const int cnt = nplg * T;
    float s = 0.f;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) { const float4 v = gp[i]; s += (v.x + v.y) + (v.z + v.w); }
    const float mean = (float)(gn_block_sum(s, red) / n);
    s = 0.f;
    for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
      const float4 v = gp[i];
      const float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
      s += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    const double var = gn_block_sum(s, red) / n;
    if (threadIdx.x == 0) { s_mean[g] = mean; s_rstd[g] = (float)(1.0 / sqrt(var + (double)eps)); }
  }
  __syncthreads();
  constexpr int PPU = E / 4;  // fp32 planes per output unit
  const int units = npl / PPU;
  for (int i = threadIdx.x; i < units * T; i += blockDim.x) {
    const int u = i / T, t = i - u * T;
    float v[E];
#pragma unroll
    for (int h = 0; h < PPU; ++h) {
      const int pl = u * PPU + h;
      const float4 a = sg[pl * T + t];
      const int g = pl / nplg;
      const float mean = s_mean[g], rstd = s_rstd[g];
      const int c0 = (plane0 + pl) * 4;
      const float4 gm = *reinterpret_cast<const float4*>(gamma + c0), bt = *reinterpret_cast<const float4*>(beta + c0);
      v[4 * h] = (a.x - mean) * rstd * gm.x + bt.x; v[4 * h + 1] = (a.y - mean) * rstd * gm.y + bt.y;
      v[4 * h + 2] = (a.z - mean) * rstd * gm.z + bt.z; v[4 * h + 3] = (a.w - mean) * rstd * gm.w + bt.w;
    }
    if (swish) {
#pragma unroll
      for (int e = 0; e < E; ++e) v[e] = v[e] / (1.f + __expf(-v[e]));
    }
    uint8_t* dst = reinterpret_cast<uint8_t*>(out) + plane_row_off(og, b, plane0 / PPU + u, t);
    if (E == 4) {
      if (rtf32) {
#pragma unroll
        for (int e = 0; e < E; ++e) v[e] = round_tf32(v[e]);
      }
      *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      uint4 o;
      o.x = pack16x2(og.fmt, v[0], v[1]); o.y = pack16x2(og.fmt, v[2], v[3]);
      o.z = pack16x2(og.fmt, v[4 % E], v[5 % E]); o.w = pack16x2(og.fmt, v[6 % E], v[7 % E]);
      *reinterpret_cast<uint4*>(dst) = o;
    }
  }
}

// ------------------------------------------------------------------------------- attention (AttnBlock1D)
// autoencoder1d.py:257-278: S[i][j] = scale * sum_c q[c][i] k[c][j]; P = softmax_j S;
// h[c][i] = sum_j v[c][j] P[i][j].  q,k,v fp32 planes; S row-major [B][T][T] fp32.
// float4 plane loads, 64x64 / 64x128 register-blocked tiles, and a channel split for the score GEMM (T = 312
// gives only 25 output tiles; the partial score planes are summed - in split order - by the softmax kernel).
// fp32 FFMA throughout (0.6 of the path's 1193 GFLOP).
constexpr int kAttnSplit = 8;

__global__ void __launch_bounds__(256) attn_scores2_kernel(const float* __restrict__ q, const float* __restrict__ k, PlaneGeom g, int C,
                                                             int T, float scale, float* __restrict__ Sp, int nsplit, int B) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float4 Qs[16][64], Ks[16][64];
  const int j0 = blockIdx.x * 64, i0 = blockIdx.y * 64;
  const int b = blockIdx.z / nsplit, z = blockIdx.z % nsplit;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int nch = (C + 3) >> 2, cps = (nch + nsplit - 1) / nsplit;
  const int c_begin = z * cps, c_end = min(nch, c_begin + cps);
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int s = 0; s < 4; ++s) acc[r][s] = 0.f;
  for (int c0 = c_begin; c0 < c_end; c0 += 16) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 2048; idx += 256) {
      const int which = idx >> 10, kc = (idx >> 6) & 15, r = idx & 63;
      const int chunk = c0 + kc, row = (which ? j0 : i0) + r;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (chunk < c_end && row < T)
        v = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(which ? k : q) + plane_row_off(g, b, chunk, row));
      if (which) Ks[kc][r] = v; else Qs[kc][r] = v;
    }
    __syncthreads();
#pragma unroll 4
    for (int kc = 0; kc < 16; ++kc) {
      float4 qv[4], kv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) { qv[r] = Qs[kc][ty * 4 + r]; kv[r] = Ks[kc][tx * 4 + r]; }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int s = 0; s < 4; ++s)
          acc[r][s] = fmaf(qv[r].x, kv[s].x, fmaf(qv[r].y, kv[s].y, fmaf(qv[r].z, kv[s].z, fmaf(qv[r].w, kv[s].w, acc[r][s]))));
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    if (i >= T) continue;
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const int j = j0 + tx * 4 + s;
      if (j < T) Sp[(((size_t)z * B + b) * T + i) * T + j] = acc[r][s] * scale;
    }
  }
}

// P[row][j] = softmax_j( sum_z Sp[z][row][j] ), row = b*T + i
__global__ void softmax_rows2_kernel(const float* __restrict__ Sp, int nsplit, int BT, int T, float* __restrict__ P) {
  pdl_launch_dependents();
  pdl_wait();
  const int row = blockIdx.x;
  __shared__ float red[32];
  extern __shared__ float srow[];  // T floats
  float m = -INFINITY;
  for (int j = threadIdx.x; j < T; j += blockDim.x) {
    float v = 0.f;
    for (int z = 0; z < nsplit; ++z) v += Sp[((size_t)z * BT + row) * T + j];
    srow[j] = v;
    m = fmaxf(m, v);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
  __syncthreads();
  m = red[0];
  for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  float s = 0.f;
  for (int j = threadIdx.x; j < T; j += blockDim.x) {
    const float e = expf(srow[j] - m);
    srow[j] = e;
    s += e;
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  s = 0.f;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
  const float inv = 1.f / s;
  for (int j = threadIdx.x; j < T; j += blockDim.x) P[(size_t)row * T + j] = srow[j] * inv;
}

// h[c][i] = sum_j v[c][j] P[i][j]; block = 64 time steps i x 32 chunks (128 channels); thread = 4 i x 2 chunks
__global__ void __launch_bounds__(256) attn_pv2_kernel(const float* __restrict__ v, PlaneGeom g, const float* __restrict__ P, int C,
                                                         int T, float* __restrict__ h, PlaneGeom hg) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float4 Vs[32][32];   // [j][chunk]
  __shared__ float Pt[32][65];    // [j][i]
  const int i0 = blockIdx.x * 64, ch0 = blockIdx.y * 32, b = blockIdx.z;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int nch = (C + 3) >> 2;
  float4 a0[4], a1[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) a0[r] = a1[r] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int j0 = 0; j0 < T; j0 += 32) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < 1024; idx += 256) {
      const int cc = idx >> 5, jj = idx & 31;
      float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ch0 + cc < nch && j0 + jj < T)
        val = *reinterpret_cast<const float4*>(reinterpret_cast<const uint8_t*>(v) + plane_row_off(g, b, ch0 + cc, j0 + jj));
      Vs[jj][cc] = val;
    }
    for (int idx = threadIdx.x; idx < 2048; idx += 256) {
      const int ii = idx >> 5, jj = idx & 31;
      Pt[jj][ii] = (i0 + ii < T && j0 + jj < T) ? P[((size_t)b * T + i0 + ii) * T + j0 + jj] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int jj = 0; jj < 32; ++jj) {
      const float4 v0 = Vs[jj][ty * 2], v1 = Vs[jj][ty * 2 + 1];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const float p = Pt[jj][tx * 4 + r];
        a0[r].x = fmaf(p, v0.x, a0[r].x); a0[r].y = fmaf(p, v0.y, a0[r].y); a0[r].z = fmaf(p, v0.z, a0[r].z); a0[r].w = fmaf(p, v0.w, a0[r].w);
        a1[r].x = fmaf(p, v1.x, a1[r].x); a1[r].y = fmaf(p, v1.y, a1[r].y); a1[r].z = fmaf(p, v1.z, a1[r].z); a1[r].w = fmaf(p, v1.w, a1[r].w);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + tx * 4 + r;
    if (i >= T) continue;
    const int c0 = ch0 + ty * 2;
    if (c0 < nch) *reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(h) + plane_row_off(hg, b, c0, i)) = a0[r];
    if (c0 + 1 < nch) *reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(h) + plane_row_off(hg, b, c0 + 1, i)) = a1[r];
  }
}

}  // namespace alcm

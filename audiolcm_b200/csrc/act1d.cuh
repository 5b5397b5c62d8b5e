// Fused anti-aliased activation: UpSample1d(2x Kaiser-sinc FIR) -> SnakeBeta -> DownSample1d.
// Restates vocoder/bigvgan/alias_free_torch/act.py:23-28, resample.py:25-33,46-49,
// filter.py:86-95 and activations.py:107-120 as ONE memory-bound kernel:
//   up   : y[2n]   = 2*sum_{q<6} x[c(n-3+q)] f[11-2q],  y[2n+1] = 2*sum_{q<6} x[c(n-2+q)] f[10-2q]
//   act  : y      += sin^2(y*e^alpha) / (e^beta + 1e-9)
//   down : out[m]  = sum_{k<12} y[clamp(2m+k-5, 0, 2T-1)] f[k]
// c() clamps to [0,T-1] (replicate padding on both FIRs).  Input: fp32 planes (E=4).  Output:
// fp32 planes (optionally RNE-rounded to tf32 for the tf32 MMA) or bf16 planes (E=8, two input
// planes per output plane).  No intermediate (2T-long) signal ever reaches HBM.
#pragma once
#include "common.cuh"

namespace alcm {

// kaiser_sinc_filter1d(cutoff=0.25, half_width=0.3, kernel_size=12) - filter.py:28-57; symmetric:
// f[k] = f[11-k]; only f[0..5] are stored.
__constant__ float c_fir[12] = {0.0020289647f, 0.0093894657f,  -0.0255434588f, -0.0576573834f, 0.1285725832f, 0.4432097971f,
                                0.4432097971f, 0.1285725832f, -0.0576573834f, -0.0255434588f, 0.0093894657f, 0.0020289647f};

constexpr int kActThreads = 128;   // large launches: 640-output tiles; small launches use 64 threads (320-output tiles)

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

// accurate snake: sin^2 has period pi -> reduce a*v to [-pi/2, pi/2] (2-term Cody-Waite), MUFU.SIN
__device__ __forceinline__ float snake_acc(float v, float ea, float ib) {
  const float t = v * ea;
  const float k = rintf(t * 0.31830988618379067f);
  float r = fmaf(-k, 3.14159274101257324f, t);
  r = fmaf(-k, -8.74227765734758577e-8f, r);
  const float s = __sinf(r);
  return fmaf(ib * s, s, v);
}
// fast snake: v + ib*sin^2(ea v) = (v + ib/2) - (ib/2) cos(2 ea v); ea2 = 2 ea, hb = ib/2
__device__ __forceinline__ float snake_fast(float v, float ea2, float hb) { return fmaf(-hb, __cosf(v * ea2), v + hb); }

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// sum of the 12 taps (the filter is normalised to 1 up to rounding): the constant part of the fast snake,
// v + ib/2, is added once per OUTPUT (scaled by this sum) instead of once per up-sampled value.
__device__ __forceinline__ float fir_sum() {
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 12; ++k) s += c_fir[k];
  return s;
}

// ---------------------------------------------------------------------------------------------------------------
// Two-phase kernel.  Round 1's register-blocked form kept R consecutive outputs per thread in registers and recomputed
// the up-sampled halo in every thread ((R+5)/R = 1.8x of the up-FIR + snake work at R = 6); it was issue-bound (ncu:
// issue slots 60-69 % busy at 41 % of the HBM peak).  Here every up-sampled value is computed exactly once per block:
//   stage : ONE bulk (TMA) copy per input plane brings x[t0-5 .. t0+TILE+9] into shared memory - the rows of a plane
//           are contiguous, so there is no per-thread load / store code at all; replicate padding is patched in by
//           the first / last block of a plane only.
//   phase1: thread i computes U consecutive "shifted pairs" P_s = (y[2t0-5+2s], y[2t0-4+2s]) - both values come from
//           the same 6-row window x[t0-5+s .. t0+s] - applies snake and stores them to shared memory (yo[s], ye[s]).
//   phase2: thread i produces R consecutive outputs: out[t0+Ri+r] = sum_{d<6} yo[Ri+r+d] f[2d] + ye[Ri+r+d] f[2d+1].
// U and R are ODD: a thread's rows are then an odd number of 16-byte units apart, which makes the 128-bit shared
// accesses of each 8-thread phase hit 8 different bank groups without any padding slots (and keeps the x tile a
// plain contiguous image of the plane, as the bulk copy needs).
// Per (time step, 4-channel plane): 24 + 24 packed FMAs for the two FIRs, 24 instructions of snake, ~8 shared-memory
// accesses - about 0.7x of the instruction count of the R = 6 register-blocked form.
// ---------------------------------------------------------------------------------------------------------------
template <int UR, int THREADS>
struct ActGeom {
  static constexpr int kPairs = THREADS * UR;       // shifted pairs per block: UR per thread, no remainder
  static constexpr int kTile = kPairs - 5;          // outputs per block (the last thread of a UR = 5 block only feeds its neighbours)
  static constexpr int kRows = kPairs + 5;          // staged x rows: x[t0-5 .. t0+kPairs-1]
  static constexpr int kXBytes = kRows * 16;
  static constexpr int kYBytes = kPairs * 16;       // one of yo / ye
  static constexpr size_t smem(int) { return (size_t)kXBytes + 2 * (size_t)kYBytes + 16; }   // ONE x buffer: plane 1 reuses it
};

// one shifted pair from a 6-row window -> (odd, even) snake'd values of a 4-channel plane
template <bool FAST>
__device__ __forceinline__ void act_pair(const float2 (&xlo)[6], const float2 (&xhi)[6], const float2 (&g2)[6], float2 ea_lo, float2 ea_hi,
                                         float2 ib_lo, float2 ib_hi, float4& yo, float4& ye) {
  float2 olo = make_float2(0.f, 0.f), ohi = olo, elo = olo, ehi = olo;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    const int ko = 10 - 2 * q, ke = 11 - 2 * q;
    const float2 wo = g2[ko < 6 ? ko : 11 - ko], we = g2[ke < 6 ? ke : 11 - ke];
    olo = ffma2(xlo[q], wo, olo); ohi = ffma2(xhi[q], wo, ohi);
    elo = ffma2(xlo[q], we, elo); ehi = ffma2(xhi[q], we, ehi);
  }
  if (FAST) {  // y = v - (ib/2) cos(2 ea v); the constant + ib/2 is added once per output (sum-1 down filter)
    const float2 nlo = make_float2(-ib_lo.x, -ib_lo.y), nhi = make_float2(-ib_hi.x, -ib_hi.y);
    float2 t, c;
    t = __fmul2_rn(olo, ea_lo); c = make_float2(__cosf(t.x), __cosf(t.y)); olo = ffma2(nlo, c, olo);
    t = __fmul2_rn(ohi, ea_hi); c = make_float2(__cosf(t.x), __cosf(t.y)); ohi = ffma2(nhi, c, ohi);
    t = __fmul2_rn(elo, ea_lo); c = make_float2(__cosf(t.x), __cosf(t.y)); elo = ffma2(nlo, c, elo);
    t = __fmul2_rn(ehi, ea_hi); c = make_float2(__cosf(t.x), __cosf(t.y)); ehi = ffma2(nhi, c, ehi);
  } else {
    olo.x = snake_acc(olo.x, ea_lo.x, ib_lo.x); olo.y = snake_acc(olo.y, ea_lo.y, ib_lo.y);
    ohi.x = snake_acc(ohi.x, ea_hi.x, ib_hi.x); ohi.y = snake_acc(ohi.y, ea_hi.y, ib_hi.y);
    elo.x = snake_acc(elo.x, ea_lo.x, ib_lo.x); elo.y = snake_acc(elo.y, ea_lo.y, ib_lo.y);
    ehi.x = snake_acc(ehi.x, ea_hi.x, ib_hi.x); ehi.y = snake_acc(ehi.y, ea_hi.y, ib_hi.y);
  }
  yo = make_float4(olo.x, olo.y, ohi.x, ohi.y);
  ye = make_float4(elo.x, elo.y, ehi.x, ehi.y);
}

template <int NPL, bool FAST, int UR, int THREADS, int MINB>  // NPL input planes per output unit: 1 -> fp32 out, 2 -> bf16 out
__global__ void __launch_bounds__(THREADS, MINB) act1d_kernel(const __grid_constant__ ActArgs a) {
  using G = ActGeom<UR, THREADS>;
  static_assert(UR == 5, "the last thread of a block has UR - 5 outputs: the tail handling below assumes none");
  extern __shared__ __align__(128) uint8_t act_smem[];
  float4* sx = reinterpret_cast<float4*>(act_smem);                                  // [kRows], shared by the NPL planes
  float4* yo = reinterpret_cast<float4*>(act_smem + (size_t)G::kXBytes);             // [kPairs]
  float4* ye = yo + G::kPairs;                                                        // [kPairs]
  const uint32_t bar = smem_u32(act_smem + (size_t)G::kXBytes + 2 * (size_t)G::kYBytes);   // one mbarrier per plane
  const int tid = threadIdx.x;
  const int t0 = blockIdx.x * G::kTile;
  const int oc = blockIdx.y, b = blockIdx.z;
  const int T = a.T;

  if (tid == 0) {
#pragma unroll
    for (int p = 0; p < NPL; ++p) mbar_init(bar + 8 * p, 1);
    fence_mbar_init();
  }
  __syncthreads();
  pdl_launch_dependents();
  pdl_wait();
  // ---- stage: one bulk copy per plane, each on its own barrier; rows outside [0,T) come from the plane's zero halo for
  //      now.  Both planes of a bf16 output unit go through the SAME buffer (15 KB per block instead of 21: 14 resident
  //      blocks per SM instead of 10): plane 1 is requested as soon as phase 1 of plane 0 has consumed the tile and
  //      lands while phase 2 of plane 0 runs ----
  const int row0 = a.xg.pad + t0 - 5;                                   // >= pad - 5 > 0
  const int nrows = min(G::kRows, a.xg.Tp - row0);
  auto request_plane = [&](int p) {
    const float4* src = reinterpret_cast<const float4*>(a.x) + ((size_t)b * a.xg.nchunk + (oc * NPL + p)) * a.xg.Tp + row0;
    mbar_expect_tx(bar + 8 * p, (uint32_t)(nrows * 16));
    bulk_g2s(smem_u32(sx), src, (uint32_t)(nrows * 16), bar + 8 * p);
  };
  if (tid == 0) request_plane(0);
  float2 f2[6], g2[6];  // broadcast taps: f[k] (down) and 2 f[k] (up), k = 0..5 (symmetric filter)
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    f2[k] = make_float2(c_fir[k], c_fir[k]);
    g2[k] = make_float2(2.f * c_fir[k], 2.f * c_fir[k]);
  }
  const bool first = (t0 == 0), last = t0 + G::kPairs - 1 > T - 1;      // block-uniform (t0 is a multiple of the tile)
  const int m0 = t0 + UR * tid;
  const bool has_out = UR * tid < G::kTile;                             // false only for the last thread when UR == 5
  uint2 held[UR];
#pragma unroll 1
  for (int p = 0; p < NPL; ++p) {
    float4* xp = sx;
    mbar_wait(bar + 8 * p, 0);
    if (first || last) {  // replicate padding of the up-sampling FIR (resample.py:28): rows outside [0,T) take the edge sample
      for (int lr = tid; lr < G::kRows; lr += THREADS) {
        const int t = t0 - 5 + lr;
        const int tc = min(max(t, 0), T - 1);
        const int src = tc - (t0 - 5);
        if (tc != t && src >= 0 && src < G::kRows) xp[lr] = xp[src];
      }
      if (NPL == 2 && p == 0) fence_proxy_async_smem();   // these generic writes precede the bulk copy of plane 1 into the same rows
      __syncthreads();
    }
    const int chunk = oc * NPL + p;
    const float4 ea = *reinterpret_cast<const float4*>(a.ea + chunk * 4);
    const float4 ib = *reinterpret_cast<const float4*>(a.ib + chunk * 4);
    const float2 ea_lo = FAST ? make_float2(2.f * ea.x, 2.f * ea.y) : make_float2(ea.x, ea.y);
    const float2 ea_hi = FAST ? make_float2(2.f * ea.z, 2.f * ea.w) : make_float2(ea.z, ea.w);
    const float2 ib_lo = FAST ? make_float2(0.5f * ib.x, 0.5f * ib.y) : make_float2(ib.x, ib.y);
    const float2 ib_hi = FAST ? make_float2(0.5f * ib.z, 0.5f * ib.w) : make_float2(ib.z, ib.w);
    {  // ---- phase 1: UR shifted pairs per thread ----
      float2 xlo[UR + 5], xhi[UR + 5];
#pragma unroll
      for (int k = 0; k < UR + 5; ++k) {
        const float4 v = xp[UR * tid + k];
        xlo[k] = make_float2(v.x, v.y);
        xhi[k] = make_float2(v.z, v.w);
      }
#pragma unroll
      for (int j = 0; j < UR; ++j) {
        float2 wl[6], wh[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) { wl[q] = xlo[j + q]; wh[q] = xhi[j + q]; }
        float4 o, e;
        act_pair<FAST>(wl, wh, g2, ea_lo, ea_hi, ib_lo, ib_hi, o, e);
        yo[UR * tid + j] = o;
        ye[UR * tid + j] = e;
      }
    }
    __syncthreads();
    if (NPL == 2 && p == 0 && tid == 0) request_plane(1);   // every thread is done reading the x tile of plane 0
    // replicate padding of the down filter acts on y (filter.py:89-91): y[j<0] = y[0], y[j>=2T] = y[2T-1].
    // yo[s] = y[2t0-5+2s], ye[s] = y[2t0-4+2s];  y[0] = ye[2-t0],  y[2T-1] = yo[T-t0+2].
    if (first || last) {
      if (first) {
        const float4 y0 = ye[2 - t0];
        for (int s = tid; s < 3 - t0; s += THREADS) {   // s <= 2-t0: jo = 2t0-5+2s < 0 ; je < 0 for s < 2-t0
          yo[s] = y0;
          if (s < 2 - t0) ye[s] = y0;
        }
      }
      if (last) {
        const int sl = T - t0 + 2;                      // yo[sl] = y[2T-1]
        if (sl >= 0 && sl < G::kPairs) {
          const float4 yl = yo[sl];
          for (int s = sl + tid; s < G::kPairs; s += THREADS) {  // je = 2t0-4+2s >= 2T for s >= sl ; jo >= 2T for s > sl
            ye[s] = yl;
            if (s > sl) yo[s] = yl;
          }
        }
      }
      __syncthreads();
    }
    // ---- phase 2: UR consecutive outputs per thread ----
    float2 alo[UR], ahi[UR];
#pragma unroll
    for (int r = 0; r < UR; ++r) alo[r] = ahi[r] = make_float2(0.f, 0.f);
    if (has_out) {
#pragma unroll
      for (int s = 0; s < UR + 5; ++s) {
        const float4 o = yo[UR * tid + s], e = ye[UR * tid + s];
        const float2 olo = make_float2(o.x, o.y), ohi = make_float2(o.z, o.w), elo = make_float2(e.x, e.y), ehi = make_float2(e.z, e.w);
#pragma unroll
        for (int r = 0; r < UR; ++r) {
          const int d = s - r;
          if (d >= 0 && d <= 5) {
            const int k0 = 2 * d, k1 = 2 * d + 1;
            const float2 w0 = f2[k0 < 6 ? k0 : 11 - k0], w1 = f2[k1 < 6 ? k1 : 11 - k1];
            alo[r] = ffma2(olo, w0, alo[r]); ahi[r] = ffma2(ohi, w0, ahi[r]);
            alo[r] = ffma2(elo, w1, alo[r]); ahi[r] = ffma2(ehi, w1, ahi[r]);
          }
        }
      }
    }
    float4 res[UR];
    if (FAST) {
      const float fs = fir_sum();
      const float2 add_lo = make_float2(ib_lo.x * fs, ib_lo.y * fs), add_hi = make_float2(ib_hi.x * fs, ib_hi.y * fs);
#pragma unroll
      for (int r = 0; r < UR; ++r) res[r] = make_float4(alo[r].x + add_lo.x, alo[r].y + add_lo.y, ahi[r].x + add_hi.x, ahi[r].y + add_hi.y);
    } else {
#pragma unroll
      for (int r = 0; r < UR; ++r) res[r] = make_float4(alo[r].x, alo[r].y, ahi[r].x, ahi[r].y);
    }
    if (NPL == 1) {
      float4* op = reinterpret_cast<float4*>(a.out) + ((size_t)b * a.og.nchunk + oc) * a.og.Tp + a.og.pad;
#pragma unroll
      for (int r = 0; r < UR; ++r) {
        if (UR * tid + r >= G::kTile || m0 + r >= T) break;
        float4 o = res[r];
        if (a.round_tf32) { o.x = round_tf32(o.x); o.y = round_tf32(o.y); o.z = round_tf32(o.z); o.w = round_tf32(o.w); }
        op[m0 + r] = o;
      }
    } else {
      uint2 pk[UR];
#pragma unroll
      for (int r = 0; r < UR; ++r) {
        pk[r] = make_uint2(pack16x2(a.og.fmt, res[r].x, res[r].y), pack16x2(a.og.fmt, res[r].z, res[r].w));
      }
      if (p == 0) {
#pragma unroll
        for (int r = 0; r < UR; ++r) held[r] = pk[r];
        __syncthreads();  // phase 2 of this plane is done with yo / ye before phase 1 of the next plane overwrites them
      } else {
        uint4* op = reinterpret_cast<uint4*>(a.out) + ((size_t)b * a.og.nchunk + oc) * a.og.Tp + a.og.pad;
#pragma unroll
        for (int r = 0; r < UR; ++r) {
          if (UR * tid + r >= G::kTile || m0 + r >= T) break;
          op[m0 + r] = make_uint4(held[r].x, held[r].y, pk[r].x, pk[r].y);
        }
      }
    }
  }
}

}  // namespace alcm

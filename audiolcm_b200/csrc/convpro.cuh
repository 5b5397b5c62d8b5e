// Activation1d fused into the conv's operand producer ("act -> conv" in one kernel) for the narrow stages.
//
// AMPBlock1.forward (vocoder/bigvgan/models.py:72-81) alternates Activation1d and Conv1d:  xt = c1(a1(x)),
// x = c2(a2(xt)) + x.  For the last three vocoder stages (C = 96 / 48 / 24) these launches are bound by HBM traffic and
// by the activation's ALU work, not by the tensor pipe, and the un-fused chain moves 28 bytes per element and layer
// (fp32 read + bf16 write per activation, bf16 read + fp32 write (+ fp32 residual read) per conv).  One N tile covers all
// output channels there, so a CTA that owns a 128-row time tile can compute the conv's A operand itself:
//
//   activation warps : each warp owns (K chunk, row segment) items of the tile's A slab: it bulk-copies (TMA) the fp32
//                      rows of its one / two input planes (with the +-5 row FIR halo) into a private double-buffered
//                      staging area, runs UpSample1d -> SnakeBeta -> DownSample1d (the two-phase form of act1d.cuh: every
//                      up-sampled value once, through a private shared-memory strip), and stores the result as bf16 /
//                      tf32 operand rows straight into the UMMA A slab - zeros outside [0,T): the conv zero-pads the
//                      ACTIVATED signal (models.py:27 get_padding), while the FIRs replicate-pad x and y (resample.py:28,
//                      filter.py:89-91).
//   warp 0           : weight producer (bulk copies of the pre-packed tap blobs, ring of stages) - as conv_umma_kernel
//   warp 1           : tcgen05.mma issuer, two TMEM accumulators
//   warps 2-5        : epilogue (bias, residual, scale, accumulate; fp32 planes out)
//
// The kernel is persistent (one CTA per SM): the activation of tile i+1 overlaps the MMAs and the epilogue of tile i
// through the two A slabs / two accumulators.  HBM traffic of an (activation, conv) pair drops from 6 + 6 (+4) to
// 4 + 4 (+4) bytes per element, and one launch replaces two.  Requires a single k-block (bf16: C <= 96, tf32: C <= 48).
#pragma once
#include "act1d.cuh"
#include "common.cuh"
#include "conv.cuh"

namespace alcm {

constexpr int kProActWarps = 8;
constexpr int kProUR = 3;                          // rows per lane and item
constexpr int kProSeg = 32 * kProUR;               // A-slab rows per work item
constexpr int kProXRows = kProSeg + 10;            // staged x rows per plane
constexpr int kProPairs = kProSeg + 5;             // shifted pairs per plane
constexpr int kProThreads = 32 * (6 + kProActWarps);

struct ProSmem {
  uint32_t a_stage, w_blob, w_stage, a_off, w_off, bias_off, bar_off, act_off, act_warp_bytes, total;
};
__host__ __device__ inline ProSmem pro_smem_layout(int kblk, int span, int NT, int w_stages, int tpg, int a_stages, int npl) {
  ProSmem L;
  L.a_stage = (uint32_t)kblk * (kTileM + span) * 16;
  L.w_blob = (uint32_t)kblk * NT * 16;
  L.w_stage = L.w_blob * tpg;
  L.a_off = 0;
  L.w_off = a_stages * L.a_stage;
  L.bias_off = L.w_off + w_stages * L.w_stage;
  L.bar_off = L.bias_off + NT * 4;
  const uint32_t nbars = 2 * a_stages + 2 * w_stages + 4;
  L.act_off = (L.bar_off + 8 * nbars + 16 + 127) & ~127u;
  // per activation warp: x staging [2 buffers][npl planes][kProXRows] float4, yo / ye [kProPairs] float4, 2 mbarriers
  L.act_warp_bytes = (uint32_t)(2 * npl * kProXRows * 16 + 2 * kProPairs * 16 + 16 + 127) & ~127u;
  L.total = L.act_off + kProActWarps * L.act_warp_bytes;
  return L;
}

// One work item of one activation warp: slab rows [r0, r0 + kProSeg) of operand chunk `chunk`.
// xs: staged fp32 rows of the NPL input planes (row lr <-> time ts - 5 + lr); yo / ye: the warp's pair strip.
template <int NPL>
__device__ __forceinline__ void pro_act_item(const ConvArgs& a, const float4* xs, float4* yo, float4* ye, uint8_t* slab, int rowsA,
                                             int chunk, int r0, int ts, int lane, const float2 (&f2)[6], const float2 (&g2)[6]) {
  constexpr int UR = kProUR;
  const int T = a.M;
  // replicate padding of the up-sampling FIR: staged rows outside [0,T) take the edge sample (when it is in the window)
  const bool lo_edge = ts - 5 < 0, hi_edge = ts + kProSeg + 4 > T - 1;   // warp-uniform
  if (lo_edge || hi_edge) {
    for (int i = lane; i < NPL * kProXRows; i += 32) {
      const int p = i / kProXRows, lr = i - p * kProXRows;
      const int t = ts - 5 + lr;
      const int tc = min(max(t, 0), T - 1);
      const int src = tc - (ts - 5);
      if (tc != t && src >= 0 && src < kProXRows) const_cast<float4*>(xs)[p * kProXRows + lr] = xs[p * kProXRows + src];
    }
    __syncwarp();
  }
  uint2 held[UR];
#pragma unroll 1
  for (int p = 0; p < NPL; ++p) {
    const int plane = chunk * NPL + p;
    const float4 ea = *reinterpret_cast<const float4*>(a.ea + plane * 4);
    const float4 ib = *reinterpret_cast<const float4*>(a.ib + plane * 4);
    const float2 ea_lo = make_float2(2.f * ea.x, 2.f * ea.y), ea_hi = make_float2(2.f * ea.z, 2.f * ea.w);
    const float2 ib_lo = make_float2(0.5f * ib.x, 0.5f * ib.y), ib_hi = make_float2(0.5f * ib.z, 0.5f * ib.w);
    const float4* xp = xs + p * kProXRows;
    {  // phase 1: UR shifted pairs per lane, + the 5 pairs beyond the segment (one each by lanes 0..4)
      float2 xlo[UR + 5], xhi[UR + 5];
#pragma unroll
      for (int k = 0; k < UR + 5; ++k) {
        const float4 v = xp[UR * lane + k];
        xlo[k] = make_float2(v.x, v.y);
        xhi[k] = make_float2(v.z, v.w);
      }
#pragma unroll
      for (int j = 0; j < UR; ++j) {
        float2 wl[6], wh[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) { wl[q] = xlo[j + q]; wh[q] = xhi[j + q]; }
        float4 o, e;
        act_pair<true>(wl, wh, g2, ea_lo, ea_hi, ib_lo, ib_hi, o, e);
        yo[UR * lane + j] = o;
        ye[UR * lane + j] = e;
      }
      if (lane < 5) {
        const int s = kProSeg + lane;
        float2 wl[6], wh[6];
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          const float4 v = xp[s + q];
          wl[q] = make_float2(v.x, v.y);
          wh[q] = make_float2(v.z, v.w);
        }
        float4 o, e;
        act_pair<true>(wl, wh, g2, ea_lo, ea_hi, ib_lo, ib_hi, o, e);
        yo[s] = o;
        ye[s] = e;
      }
    }
    __syncwarp();
    // replicate padding of the down filter on y: yo[s] = y[2ts-5+2s], ye[s] = y[2ts-4+2s]; y[0] = ye[2-ts], y[2T-1] = yo[T-ts+2]
    if (ts < 3) {
      const int s0 = 2 - ts;
      if (s0 < kProPairs) {
        const float4 y0 = ye[s0];
        __syncwarp();
        for (int s = lane; s <= s0; s += 32) {
          yo[s] = y0;
          if (s < s0) ye[s] = y0;
        }
      }
      __syncwarp();
    }
    {
      const int sl = T - ts + 2;
      if (sl < kProPairs && sl >= 0) {
        const float4 yl = yo[sl];
        __syncwarp();
        for (int s = sl + lane; s < kProPairs; s += 32) {
          ye[s] = yl;
          if (s > sl) yo[s] = yl;
        }
        __syncwarp();
      }
    }
    // phase 2: UR consecutive outputs per lane
    float2 alo[UR], ahi[UR];
#pragma unroll
    for (int r = 0; r < UR; ++r) alo[r] = ahi[r] = make_float2(0.f, 0.f);
#pragma unroll
    for (int s = 0; s < UR + 5; ++s) {
      const float4 o = yo[UR * lane + s], e = ye[UR * lane + s];
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
    const float fs = fir_sum();
    const float2 add_lo = make_float2(ib_lo.x * fs, ib_lo.y * fs), add_hi = make_float2(ib_hi.x * fs, ib_hi.y * fs);
    __syncwarp();  // every lane is done with yo / ye before the next plane's phase 1 overwrites them
    if (NPL == 1) {  // tf32 operand rows
#pragma unroll
      for (int r = 0; r < UR; ++r) {
        const int row = r0 + UR * lane + r, t = ts + UR * lane + r;
        if (row >= rowsA) break;
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t >= 0 && t < T)
          o = make_float4(round_tf32(alo[r].x + add_lo.x), round_tf32(alo[r].y + add_lo.y), round_tf32(ahi[r].x + add_hi.x),
                          round_tf32(ahi[r].y + add_hi.y));
        *reinterpret_cast<float4*>(slab + ((size_t)chunk * rowsA + row) * 16) = o;
      }
    } else {
      uint2 pk[UR];
#pragma unroll
      for (int r = 0; r < UR; ++r) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(alo[r].x + add_lo.x, alo[r].y + add_lo.y);
        __nv_bfloat162 h1 = __floats2bfloat162_rn(ahi[r].x + add_hi.x, ahi[r].y + add_hi.y);
        pk[r] = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
      }
      if (p == 0) {
#pragma unroll
        for (int r = 0; r < UR; ++r) held[r] = pk[r];
      } else {
#pragma unroll
        for (int r = 0; r < UR; ++r) {
          const int row = r0 + UR * lane + r, t = ts + UR * lane + r;
          if (row >= rowsA) break;
          uint4 o = make_uint4(0u, 0u, 0u, 0u);
          if (t >= 0 && t < T) o = make_uint4(held[r].x, held[r].y, pk[r].x, pk[r].y);
          *reinterpret_cast<uint4*>(slab + ((size_t)chunk * rowsA + row) * 16) = o;
        }
      }
    }
  }
}

// a.x / a.xg: the fp32 planes the Activation1d reads; a.ea / a.ib: its SnakeBeta parameters (per input channel);
// everything else as in conv_umma_kernel (nphase = 1, ksplit = 1, nkb = 1, acc_stages = 2).
template <int KIND>  // 0: bf16 operands (two fp32 planes per K chunk), 1: tf32 (one)
__global__ void __launch_bounds__(kProThreads, 1) conv_actpro_kernel(const __grid_constant__ ConvArgs a) {
  constexpr int NPL = (KIND == 0) ? 2 : 1;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = a.w_stages;
  const int rowsA = kTileM + a.span;
  const int AS = a.a_stages;
  const ProSmem L = pro_smem_layout(a.kblk, a.span, a.NT, S, a.tpg, AS, NPL);
  const uint32_t sA = smem_u32(smem) + L.a_off;
  const uint32_t sW = smem_u32(smem) + L.w_off;
  const uint32_t bars = smem_u32(smem) + L.bar_off;
  float* s_bias = reinterpret_cast<float*>(smem + L.bias_off);
  const uint32_t a_full = bars, a_empty = bars + 8 * AS, w_full = bars + 16 * AS, w_empty = w_full + 8 * S;
  const uint32_t acc_full = w_empty + 8 * S, acc_empty = acc_full + 16;
  const int nbars = 2 * AS + 2 * S + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bar_off + 8 * nbars);

  if (threadIdx.x == 0) {
    for (int i = 0; i < AS; ++i) mbar_init(a_full + 8 * i, kProActWarps);   // one arrival per activation warp
    for (int i = 0; i < AS; ++i) mbar_init(a_empty + 8 * i, 1);
    for (int i = 0; i < 2 * S + 2; ++i) mbar_init(w_full + 8 * i, 1);       // w_full, w_empty, acc_full[2]
    mbar_init(acc_empty, 4);
    mbar_init(acc_empty + 8, 4);
    for (int w = 0; w < kProActWarps; ++w) {
      const uint32_t xb = smem_u32(smem) + L.act_off + w * L.act_warp_bytes + 2 * NPL * kProXRows * 16 + 2 * kProPairs * 16;
      mbar_init(xb, 1);
      mbar_init(xb + 8, 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), a.tmem_cols);
    tmem_relinquish();
  }
  const int chunks_valid = a.xg.nchunk / NPL;    // operand K chunks that exist (plane_cpad: a multiple of 8 channels)
  if (chunks_valid < a.kblk) {                   // the missing K chunk(s) of a narrow operand: an all-zero slab, written once
    const uint32_t lo = (uint32_t)chunks_valid * (uint32_t)rowsA, hi = (uint32_t)a.kblk * (uint32_t)rowsA;
    for (int s = 0; s < AS; ++s) {
      uint4* base = reinterpret_cast<uint4*>(smem + L.a_off + (size_t)s * L.a_stage);
      for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) base[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_launch_dependents();

  if (warp == 0) {
    // ---- weight producer -------------------------------------------------------------------------------------
    const bool leader = elect_one();
    const int ntaps = a.ntaps, tpg = a.tpg;
    int ws = 0;
    uint32_t wpar = 1;
    for (int tile = blockIdx.x; tile < a.tiles_total; tile += gridDim.x) {
      const ConvTile T = conv_tile<false>(a, tile);
      const uint8_t* wsrc = a.w + ((size_t)T.nt * a.nkb) * ntaps * L.w_blob;
      for (int j0 = 0; j0 < ntaps; j0 += tpg) {
        const uint32_t bytes = (uint32_t)min(tpg, ntaps - j0) * L.w_blob;
        mbar_wait(w_empty + 8 * ws, wpar);
        if (leader) {
          mbar_expect_tx(w_full + 8 * ws, bytes);
          bulk_g2s(sW + ws * L.w_stage, wsrc, bytes, w_full + 8 * ws);
        }
        wsrc += bytes;
        if (++ws == S) { ws = 0; wpar ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---- MMA issuer (same burst structure as conv_mma_loop; one k-block per tile) ---------------------------------
    auto run = [&](auto nk2_tag) {
      constexpr int NK2 = decltype(nk2_tag)::value;
      const bool leader = elect_one();
      const uint32_t lead = leader ? 1u : 0u;
      const uint64_t a_desc0 = umma_desc_kmajor(sA, rowsA * 16, 128);
      const uint64_t w_desc0 = umma_desc_kmajor(sW, a.NT * 16, 128);
      const uint32_t a_step = (uint32_t)(2 * rowsA), w_step = (uint32_t)(2 * a.NT);
      const uint32_t a_stage16 = L.a_stage >> 4, w_stage16 = L.w_stage >> 4, w_blob16 = L.w_blob >> 4;
      const int ntaps = a.ntaps, tpg = a.tpg;
      const uint32_t idesc = a.idesc;
      int ws = 0, as = 0, it = 0;
      uint32_t wpar = 0, apar = 0;
      for (int tile = blockIdx.x; tile < a.tiles_total; tile += gridDim.x, ++it) {
        const uint32_t shift0 = (uint32_t)(a.tap_off[0][0] - a.min_off[0]);
        const uint64_t dshift = (uint64_t)(int64_t)(ntaps > 1 ? a.tap_off[0][1] - a.tap_off[0][0] : 0);
        const int st = it & 1;
        const uint32_t tmem_d = tmem_base + (uint32_t)(st * a.NT);
        mbar_wait(acc_empty + 8 * st, ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        mbar_wait(a_full + 8 * as, apar);
        tc_fence_after();
        uint32_t acc = 0;
        uint64_t a_tap = a_desc0 + (uint64_t)(as * a_stage16 + shift0);
        for (int j0 = 0; j0 < ntaps; j0 += tpg) {
          const int g = min(tpg, ntaps - j0);
          mbar_wait(w_full + 8 * ws, wpar);
          tc_fence_after();
          uint64_t bd = w_desc0 + (uint64_t)(ws * w_stage16);
          for (int t = 0; t < g; ++t, a_tap += dshift, bd += w_blob16) {
#pragma unroll
            for (int i = 0; i < NK2; ++i)
              umma_ss_pred<KIND>(tmem_d, a_tap + (uint64_t)(i * a_step), bd + (uint64_t)(i * w_step), idesc, (i > 0) ? 1u : acc, lead);
            acc = 1;
          }
          tc_commit_pred(w_empty + 8 * ws, lead);
          if (j0 + g == ntaps) tc_commit_pred(a_empty + 8 * as, lead);
          if (++ws == S) { ws = 0; wpar ^= 1; }
        }
        if (++as == AS) { as = 0; apar ^= 1; }
        tc_commit_pred(acc_full + 8 * st, lead);
      }
      __syncwarp();
    };
    switch (a.kblk >> 1) {
      case 1: run(std::integral_constant<int, 1>{}); break;
      case 2: run(std::integral_constant<int, 2>{}); break;
      case 3: run(std::integral_constant<int, 3>{}); break;
      case 4: run(std::integral_constant<int, 4>{}); break;
      case 5: run(std::integral_constant<int, 5>{}); break;
      default: run(std::integral_constant<int, 6>{}); break;
    }
  } else if (warp < 6) {
    // ---- epilogue ----------------------------------------------------------------------------------------------
    const int et = threadIdx.x - 64;
    const int qd = warp & 3;
    const int row = qd * 32 + lane;
    const float scale = a.scale;
    const int nq = a.NT >> 2;
    const size_t plane4 = (size_t)a.og.Tp;
    const float4* res4 = reinterpret_cast<const float4*>(a.res);
    float4* out4 = reinterpret_cast<float4*>(a.out);
    for (int i = et; i < a.NT; i += 128) s_bias[i] = a.bias ? __ldg(a.bias + i) : 0.f;   // single N tile
    asm volatile("bar.sync 1, 128;" ::: "memory");
    pdl_wait();
    int it = 0;
    for (int tile = blockIdx.x; tile < a.tiles_total; tile += gridDim.x, ++it) {
      const ConvTile T = conv_tile<false>(a, tile);
      const int st = it & 1;
      const uint32_t tmem_d = tmem_base + (uint32_t)(st * a.NT);
      const int q = T.q0 + row;
      const bool valid = q >= 0 && q < a.M;
      const int nq_valid = min(nq, a.og.nchunk);
      const size_t off0 = ((size_t)T.b * a.og.nchunk) * a.og.Tp + a.og.pad + (size_t)q;
      const bool has_res = (a.res != nullptr) && valid, accum = (a.accum != 0) && valid;
      mbar_wait(acc_full + 8 * st, (it >> 1) & 1);
      tc_fence_after();
      for (int c0 = 0; c0 < a.NT; c0 += 16) {
        float4 rr[4], oo[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int cq = (c0 >> 2) + g;
          rr[g] = oo[g] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cq < nq_valid) {
            if (has_res) rr[g] = res4[off0 + (size_t)cq * plane4];
            if (accum) oo[g] = out4[off0 + (size_t)cq * plane4];
          }
        }
        uint32_t u[16];
        tmem_ld_x16(tmem_d + ((uint32_t)(qd * 32) << 16) + c0, u);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int cq = (c0 >> 2) + g;
          const float4 bb = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * g);
          float4 r = make_float4(__uint_as_float(u[4 * g]) + bb.x, __uint_as_float(u[4 * g + 1]) + bb.y, __uint_as_float(u[4 * g + 2]) + bb.z,
                                 __uint_as_float(u[4 * g + 3]) + bb.w);
          r.x = (r.x + rr[g].x) * scale + oo[g].x; r.y = (r.y + rr[g].y) * scale + oo[g].y;
          r.z = (r.z + rr[g].z) * scale + oo[g].z; r.w = (r.w + rr[g].w) * scale + oo[g].w;
          if (valid && cq < nq_valid) out4[off0 + (size_t)cq * plane4] = r;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + 8 * st);
    }
  } else {
    // ---- activation warps: produce the A slabs ---------------------------------------------------------------------
    const int aw = warp - 6;
    uint8_t* area = smem + L.act_off + (size_t)aw * L.act_warp_bytes;
    float4* xbuf = reinterpret_cast<float4*>(area);                                  // [2][NPL][kProXRows]
    float4* yo = xbuf + 2 * NPL * kProXRows;
    float4* ye = yo + kProPairs;
    const uint32_t xbar = smem_u32(ye + kProPairs);                                  // two mbarriers
    float2 f2[6], g2[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      f2[k] = make_float2(c_fir[k], c_fir[k]);
      g2[k] = make_float2(2.f * c_fir[k], 2.f * c_fir[k]);
    }
    const int segs = (rowsA + kProSeg - 1) / kProSeg;
    const int nitems = chunks_valid * segs;          // item = chunk * segs + seg
    const size_t plane_rows = (size_t)a.xg.Tp;
    // item geometry: slab rows [seg*kProSeg, ...) of `chunk`; staged x rows start at time ts - 5
    auto issue = [&](int tile, int item, int buf) {
      const ConvTile T = conv_tile<false>(a, tile);
      const int chunk = item / segs, seg = item - chunk * segs;
      const int ts = T.q0 + a.min_off[0] + seg * kProSeg;
      const int row0 = a.xg.pad + ts - 5;                               // >= 0: pad >= |min_off| + 5
      const int nrows = max(0, min(kProXRows, a.xg.Tp - row0));
      fence_proxy_async_smem();   // this warp's earlier (generic) reads of the buffer are ordered before the async write
      __syncwarp();
      if (lane == 0 && nrows > 0) {
        mbar_expect_tx(xbar + 8 * buf, (uint32_t)(NPL * nrows * 16));
#pragma unroll
        for (int p = 0; p < NPL; ++p) {
          const float4* src = reinterpret_cast<const float4*>(a.x) + ((size_t)T.b * a.xg.nchunk + (chunk * NPL + p)) * plane_rows + row0;
          bulk_g2s(smem_u32(xbuf + (buf * NPL + p) * kProXRows), src, (uint32_t)(nrows * 16), xbar + 8 * buf);
        }
      } else if (lane == 0) {
        mbar_arrive(xbar + 8 * buf);
      }
    };
    pdl_wait();
    int buf = 0;
    uint32_t xpar[2] = {0u, 0u};
    int as = 0;
    uint32_t apar = 1;
    // prime: this warp's first item of the CTA's first tile (items of a tile: aw, aw + kProActWarps, ...)
    const int cur_tile = blockIdx.x, cur_item = aw;
    const bool has_work = aw < nitems;
    if (has_work && cur_tile < a.tiles_total) issue(cur_tile, cur_item, 0);
    for (int tile = blockIdx.x; tile < a.tiles_total; tile += gridDim.x) {
      mbar_wait(a_empty + 8 * as, apar);     // the MMAs that read this slab two tiles ago have retired
      uint8_t* slab = smem + L.a_off + (size_t)as * L.a_stage;
      if (has_work) {
        for (int item = aw; item < nitems; item += kProActWarps) {
          // prefetch this warp's next item (same tile, or the first item of its next tile) into the other buffer
          int nt = tile, ni = item + kProActWarps;
          if (ni >= nitems) { nt = tile + gridDim.x; ni = aw; }
          if (nt < a.tiles_total) issue(nt, ni, buf ^ 1);
          mbar_wait(xbar + 8 * buf, xpar[buf]);
          xpar[buf] ^= 1u;
          const ConvTile T = conv_tile<false>(a, tile);
          const int chunk = item / segs, seg = item - chunk * segs;
          const int ts = T.q0 + a.min_off[0] + seg * kProSeg;
          pro_act_item<NPL>(a, xbuf + buf * NPL * kProXRows, yo, ye, slab, rowsA, chunk, seg * kProSeg, ts, lane, f2, g2);
          buf ^= 1;
        }
      }
      fence_proxy_async_smem();              // generic-proxy slab writes -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full + 8 * as);
      if (++as == AS) { as = 0; apar ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

}  // namespace alcm

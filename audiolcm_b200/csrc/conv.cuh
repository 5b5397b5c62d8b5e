// Implicit-GEMM 1-D convolution on the plane layout.
//
//   conv_umma_kernel : tcgen05.mma (TMEM accumulator), operands staged by 1-D bulk TMA copies,
//                      warp-specialised (TMA producer / MMA issuer / 4 epilogue warps).
//   conv_simt_kernel : exact-fp32 FFMA version of the same contract (precision "fp32" mode and
//                      the on-device cross-check for the tensor-core kernel).
//
// One launch covers Conv1d (any k, dilation), and - through "phases" - the polyphase forms of
// ConvTranspose1d(k=2u, stride u) and nearest-2x-upsample+Conv1d(k=3):
//   out[b][co][q*ostride + phase] = scale * ( bias[co] + sum_tap sum_ci  W[phase][tap][co][ci] *
//                                    x[b][ci][q + tap_off[phase][tap]] + res[...] ) (+ out[...] if accum)
// Reference ops restated: vocoder/bigvgan/models.py:36-53,143,150-155 (Conv1d / ConvTranspose1d),
// ldm/models/autoencoder1d.py:186-213,291-295 (Conv1d, Upsample1D); bias / residual add
// (models.py:79) / block mean (models.py:193-196) are fused in the epilogue.
#pragma once
#include "common.cuh"

namespace alcm {

constexpr int kTileM = 128;

struct ConvArgs {
  const uint8_t* x;   // input planes (operand dtype)
  PlaneGeom xg;
  const uint8_t* w;   // UMMA: packed blobs [phase][n_tile][kb][tap][kc][NT][16B]; SIMT: fp32 [phase][tap][Cout][Cin]
  const float* bias;  // padded to n_tiles*NT (UMMA) / Cout (SIMT); may be null
  float* out;         // fp32 planes
  PlaneGeom og;
  const float* res;   // residual, geometry og, may alias out; may be null
  int M;              // rows q per batch item
  int ostride;        // = nphase
  int nphase, ntaps;
  int tap_off[kMaxPhase][kMaxTaps];
  int min_off[kMaxPhase];
  int span;           // max over phases of (max_off - min_off)
  int Cin, Cout;      // logical channels (SIMT) ; Cin padded to planes
  int kchunks, kblk, nkb;  // 16-byte K chunks: total, per k-block (even), number of k-blocks
  int NT, n_tiles, tmem_cols, w_stages;
  uint32_t idesc;
  unsigned long long w_phase_stride;  // bytes between phases in w
  float scale;
  int accum;
  int dbg;  // micro-benchmark only: bit0 skip weight copies, bit1 skip activation copies (results are garbage)
};

// ---------------------------------------------------------------------------------------------
// tcgen05 kernel.  grid = (ceil(M/128), n_tiles, B*nphase), block = 192 threads.
//   warp 0   : producer - per k-block one A slab per K chunk ([128+span rows] x 16 B, contiguous in
//              the plane) and per tap one pre-packed weight blob, all cp.async.bulk + mbarrier tx.
//   warp 1   : allocates TMEM, one lane issues tcgen05.mma; the tap shift is a 16 B*offset bump of
//              the A descriptor's start address (no im2col copy is ever materialised).
//   warps 2-5: epilogue - tcgen05.ld 32x32b (thread = time row), bias/residual/scale/accumulate,
//              float4 stores: each warp-level store is 512 contiguous bytes of one output plane.
// ---------------------------------------------------------------------------------------------
struct ConvSmemLayout {
  uint32_t a_stage, w_stage, a_off, w_off, bar_off, total;
};
__host__ __device__ inline ConvSmemLayout conv_smem_layout(int kblk, int span, int NT, int w_stages) {
  ConvSmemLayout L;
  L.a_stage = (uint32_t)kblk * (kTileM + span) * 16;
  L.w_stage = (uint32_t)kblk * NT * 16;
  L.a_off = 0;
  L.w_off = 2 * L.a_stage;
  L.bar_off = L.w_off + w_stages * L.w_stage;
  L.total = L.bar_off + 8 * (5 + 2 * w_stages) + 16;
  return L;
}

template <int KIND>
__global__ void __launch_bounds__(192) conv_umma_kernel(const __grid_constant__ ConvArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x, nt = blockIdx.y;
  const int b = blockIdx.z / a.nphase, ph = blockIdx.z % a.nphase;
  const int S = a.w_stages;
  const int rowsA = kTileM + a.span;
  const ConvSmemLayout L = conv_smem_layout(a.kblk, a.span, a.NT, S);
  const uint32_t sA = smem_u32(smem) + L.a_off;
  const uint32_t sW = smem_u32(smem) + L.w_off;
  const uint32_t bars = smem_u32(smem) + L.bar_off;
  // barrier slots: [0,1] a_full, [2,3] a_empty, [4..4+S) w_full, [4+S..4+2S) w_empty, [4+2S] acc_full
  const uint32_t a_full = bars, a_empty = bars + 16, w_full = bars + 32, w_empty = bars + 32 + 8 * S;
  const uint32_t acc_full = bars + 32 + 16 * S;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bar_off + 8 * (5 + 2 * S));

  if (threadIdx.x == 0) {
    for (int i = 0; i < 5 + 2 * S; ++i) mbar_init(bars + 8 * i, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), a.tmem_cols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int q0 = mt * kTileM;

  // Both asynchronous roles keep their warp CONVERGENT and predicate the async instructions with
  // elect.sync: UBLKCP / UTCHMMA / UTCBAR take warp-uniform operands, and issuing them from a
  // divergent `if (lane == 0)` makes the compiler wrap each one in an ELECT/R2UR/BRA serialisation
  // loop, which left the tensor pipe waiting on the issuing thread.
  if (warp == 0) {
    const bool leader = elect_one();
    const int row0 = q0 + a.min_off[ph] + a.xg.pad;  // >= 0 because |min_off| <= pad
    const int nrows = min(rowsA, a.xg.Tp - row0);
    const uint8_t* wsrc = a.w + (size_t)ph * a.w_phase_stride + (size_t)nt * a.nkb * a.ntaps * L.w_stage;
    const size_t plane_bytes = (size_t)a.xg.Tp * 16;
    const uint8_t* xsrc = a.x + ((size_t)b * a.xg.nchunk * a.xg.Tp + row0) * 16;
    const uint32_t a_bytes = (uint32_t)nrows * 16, a_pitch = (uint32_t)rowsA * 16;
    int ws = 0;
    uint32_t wpar = 1;  // producer waits on the "previous" phase of the empty barriers first
    for (int kb = 0; kb < a.nkb; ++kb) {
      const int as = kb & 1;
      mbar_wait(a_empty + 8 * as, ((kb >> 1) & 1) ^ 1);
      if (a.dbg & 2) {
        if (leader) mbar_arrive(a_full + 8 * as);
        xsrc += plane_bytes * a.kblk;
      } else {
        if (leader) mbar_expect_tx(a_full + 8 * as, (uint32_t)a.kblk * a_bytes);
        uint32_t dst = sA + as * L.a_stage;
        for (int c = 0; c < a.kblk; ++c, dst += a_pitch, xsrc += plane_bytes)
          if (leader) bulk_g2s(dst, xsrc, a_bytes, a_full + 8 * as);
      }
      for (int j = 0; j < a.ntaps; ++j) {
        mbar_wait(w_empty + 8 * ws, wpar);
        if (leader) {
          if (a.dbg & 1) {
            mbar_arrive(w_full + 8 * ws);
          } else {
            mbar_expect_tx(w_full + 8 * ws, L.w_stage);
            bulk_g2s(sW + ws * L.w_stage, wsrc, L.w_stage, w_full + 8 * ws);
          }
        }
        wsrc += L.w_stage;
        if (++ws == S) { ws = 0; wpar ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    const bool leader = elect_one();
    // Descriptors differ only in their 14-bit start-address field: build the constant part once
    // and add (byte offset >> 4) per MMA.
    const uint64_t a_desc0 = umma_desc_kmajor(sA, rowsA * 16, 128);
    const uint64_t w_desc0 = umma_desc_kmajor(sW, a.NT * 16, 128);
    const uint32_t a_step = (uint32_t)(2 * rowsA), w_step = (uint32_t)(2 * a.NT);  // two K chunks, in 16 B units
    const uint32_t a_stage16 = L.a_stage >> 4, w_stage16 = L.w_stage >> 4;
    // every parameter the loop needs lives in a register: the asm memory clobbers would otherwise
    // make the compiler re-read the constant bank on each iteration of the single issuing warp.
    // The tap shift (row offset of tap j inside the A slab) is linear in j for every conv form here.
    const int nkb = a.nkb, ntaps = a.ntaps, nk2 = a.kblk >> 1;
    const uint32_t idesc = a.idesc;
    const uint32_t shift0 = (uint32_t)(a.tap_off[ph][0] - a.min_off[ph]);
    const int dshift = ntaps > 1 ? a.tap_off[ph][1] - a.tap_off[ph][0] : 0;
    int ws = 0;
    uint32_t wpar = 0, acc = 0;
    for (int kb = 0; kb < nkb; ++kb) {
      const int as = kb & 1;
      mbar_wait(a_full + 8 * as, (kb >> 1) & 1);
      uint64_t a_tap = a_desc0 + (uint64_t)(as * a_stage16 + shift0);
      for (int j = 0; j < ntaps; ++j, a_tap += (uint64_t)(int64_t)dshift) {
        mbar_wait(w_full + 8 * ws, wpar);
        tc_fence_after();
        if (leader) {
          uint64_t ad = a_tap, bd = w_desc0 + (uint64_t)(ws * w_stage16);
#pragma unroll 4
          for (int i = 0; i < nk2; ++i, ad += a_step, bd += w_step) {
            umma_ss<KIND>(tmem_base, ad, bd, idesc, acc);
            acc = 1;
          }
          tc_commit(w_empty + 8 * ws);  // frees the weight slot when these MMAs retire
          if (j == ntaps - 1) tc_commit(a_empty + 8 * as);
        }
        if (++ws == S) { ws = 0; wpar ^= 1; }
      }
    }
    if (leader) tc_commit(acc_full);
    __syncwarp();
  } else {
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int qd = warp & 3;  // TMEM lane quarter this warp may access
    const int q = q0 + qd * 32 + lane;
    const bool valid = q < a.M;
    const size_t orow = (size_t)q * a.ostride + ph;
    const float scale = a.scale;
    for (int c0 = 0; c0 < a.NT; c0 += 16) {
      uint32_t v[16];
      tmem_ld_x16(tmem_base + ((uint32_t)(qd * 32) << 16) + c0, v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = nt * a.NT + c0 + 4 * g;
          const int chunk = col >> 2;
          if (chunk < a.og.nchunk) {
            float4 r = make_float4(__uint_as_float(v[4 * g]), __uint_as_float(v[4 * g + 1]), __uint_as_float(v[4 * g + 2]),
                                   __uint_as_float(v[4 * g + 3]));
            if (a.bias) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(a.bias + col));
              r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w;
            }
            const size_t off = (((size_t)b * a.og.nchunk + chunk) * a.og.Tp + a.og.pad + orow) * 4;
            if (a.res) {
              const float4 rr = *reinterpret_cast<const float4*>(a.res + off);
              r.x += rr.x; r.y += rr.y; r.z += rr.z; r.w += rr.w;
            }
            r.x *= scale; r.y *= scale; r.z *= scale; r.w *= scale;
            if (a.accum) {
              const float4 oo = *reinterpret_cast<const float4*>(a.out + off);
              r.x += oo.x; r.y += oo.y; r.z += oo.z; r.w += oo.w;
            }
            *reinterpret_cast<float4*>(a.out + off) = r;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, a.tmem_cols);
  }
}

// ---------------------------------------------------------------------------------------------
// Exact fp32 kernel (CUDA-core FFMA).  grid = (ceil(M/64), ceil(Cout/64), B*nphase), block 256.
// Input planes fp32 (E=4), weights fp32 [phase][tap][Cout][Cin].
// ---------------------------------------------------------------------------------------------
constexpr int kSimtTM = 64, kSimtTN = 64, kSimtKC = 16;
constexpr int kSimtRows = kSimtTM + 52;  // tile rows + largest span (50), padded

__global__ void __launch_bounds__(256) conv_simt_kernel(const __grid_constant__ ConvArgs a) {
  __shared__ float Xs[kSimtKC][kSimtRows];
  __shared__ float Ws[kSimtKC][kSimtTN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int q0 = blockIdx.x * kSimtTM, n0 = blockIdx.y * kSimtTN;
  const int b = blockIdx.z / a.nphase, ph = blockIdx.z % a.nphase;
  const float* xb = reinterpret_cast<const float*>(a.x) + (size_t)b * a.xg.nchunk * a.xg.Tp * 4;
  const float* wp = reinterpret_cast<const float*>(a.w) + (size_t)ph * a.ntaps * a.Cout * a.Cin;
  const int rows = kSimtTM + a.span;
  const int row0 = q0 + a.min_off[ph] + a.xg.pad;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int cin_pad = a.xg.nchunk * 4;
  for (int c0 = 0; c0 < cin_pad; c0 += kSimtKC) {
    __syncthreads();
    // X tile: 4 planes x rows float4 -> Xs[c][r]
    for (int idx = tid; idx < 4 * rows; idx += 256) {
      const int pl = idx / rows, r = idx % rows;
      const int chunk = c0 / 4 + pl;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (chunk < a.xg.nchunk && row0 + r < a.xg.Tp)
        v = *reinterpret_cast<const float4*>(xb + ((size_t)chunk * a.xg.Tp + row0 + r) * 4);
      Xs[pl * 4 + 0][r] = v.x; Xs[pl * 4 + 1][r] = v.y; Xs[pl * 4 + 2][r] = v.z; Xs[pl * 4 + 3][r] = v.w;
    }
    for (int j = 0; j < a.ntaps; ++j) {
      __syncthreads();
      for (int idx = tid; idx < kSimtKC * kSimtTN; idx += 256) {
        const int n = idx / kSimtKC, c = idx % kSimtKC;
        const int co = n0 + n, ci = c0 + c;
        Ws[c][n] = (co < a.Cout && ci < a.Cin) ? wp[((size_t)j * a.Cout + co) * a.Cin + ci] : 0.f;
      }
      __syncthreads();
      const int sh = a.tap_off[ph][j] - a.min_off[ph];
#pragma unroll
      for (int c = 0; c < kSimtKC; ++c) {
        float xv[4], wv[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xv[i] = Xs[c][ty * 4 + i + sh];
#pragma unroll
        for (int jn = 0; jn < 4; ++jn) wv[jn] = Ws[c][tx * 4 + jn];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int jn = 0; jn < 4; ++jn) acc[i][jn] = fmaf(xv[i], wv[jn], acc[i][jn]);
      }
    }
  }
  const int col = n0 + tx * 4;
  const int chunk = col >> 2;
  if (chunk >= a.og.nchunk) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= a.M) continue;
    float r[4];
#pragma unroll
    for (int jn = 0; jn < 4; ++jn) {
      const int co = col + jn;
      r[jn] = (co < a.Cout) ? acc[i][jn] + (a.bias ? a.bias[co] : 0.f) : 0.f;
    }
    const size_t off = (((size_t)b * a.og.nchunk + chunk) * a.og.Tp + a.og.pad + (size_t)q * a.ostride + ph) * 4;
    float4 o = make_float4(r[0], r[1], r[2], r[3]);
    if (a.res) {
      const float4 rr = *reinterpret_cast<const float4*>(a.res + off);
      o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
    }
    o.x *= a.scale; o.y *= a.scale; o.z *= a.scale; o.w *= a.scale;
    if (a.accum) {
      const float4 oo = *reinterpret_cast<const float4*>(a.out + off);
      o.x += oo.x; o.y += oo.y; o.z += oo.z; o.w += oo.w;
    }
    *reinterpret_cast<float4*>(a.out + off) = o;
  }
}

}  // namespace alcm

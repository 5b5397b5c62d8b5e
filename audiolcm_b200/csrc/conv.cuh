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
  int NT, n_tiles, tmem_cols;
  int w_stages;       // weight ring depth
  int a_stages;       // A-slab ring depth (2..4): short k-blocks (few taps) are bound by the slab's load latency
  int tpg;            // taps per weight stage (one bulk copy + one mbarrier round trip per `tpg` taps)
  int w_resident;     // persistent launch whose whole weight set (all taps, single N tile, single k-block) stays in shared
                      // memory: w_stages == number of tap groups, loaded by the CTA's first tile only
  uint32_t idesc;
  unsigned long long w_phase_stride;  // bytes between phases in w
  unsigned long long w_batch_stride;  // bytes between batch items in w (0: one weight set for all; > 0: per-item operands, e.g. attention K / V)
  float scale;
  int accum;
  int dbg;  // micro-benchmark only: bit0 skip weight copies, bit1 skip activation copies (results are garbage)
  // split-K (launches with too few output tiles to fill the GPU): tile index = (...)*ksplit + s (see tiles_total below),
  // split s reduces k-blocks [s*nkb/ksplit, (s+1)*nkb/ksplit); raw partial tiles go to `ws`, the last CTA
  // to arrive at a tile (counter in `tile_ctr`, self-resetting) sums them in split order and runs the
  // epilogue, so the result does not depend on arrival order.
  int ksplit;
  int cluster_splitk;      // 1: the ksplit CTAs of a tile form a thread-block cluster and reduce through distributed shared
                           //    memory (each CTA owns NT/ksplit columns); 0: global workspace + last-arriver fix-up
  float* ws;               // [tile][ksplit][NT/4][128] float4
  unsigned int* tile_ctr;  // [tile]
  long long* trace;        // micro-benchmark only: 8 timestamps per CTA (tools/bench_conv.py with ALCM_TRACE=1)
  // Tiles: linear index = (z * n_tiles + nt) * tiles_m + mt, z = (b * nphase + ph) * ksplit + split.  The grid is
  // 1-D; a CTA processes tiles blockIdx.x, blockIdx.x + gridDim.x, ... (persistent launch when gridDim.x < tiles_total,
  // with two TMEM accumulators so that the epilogue of one tile overlaps the main loop of the next).
  int tiles_m, tiles_total, acc_stages;
};

// ---------------------------------------------------------------------------------------------
// tcgen05 kernel.  1-D grid over output tiles (or over resident CTA slots for a persistent launch), block = 192 threads.
//   warp 0   : producer - per k-block one A slab per K chunk ([128+span rows] x 16 B, contiguous in
//              the plane) and per group of `tpg` taps one bulk copy of their (contiguous) pre-packed
//              weight blobs, all cp.async.bulk + mbarrier tx.
//   warp 1   : allocates TMEM, one lane issues tcgen05.mma; the tap shift is a 16 B*offset bump of
//              the A descriptor's start address (no im2col copy is ever materialised).
//   warps 2-5: epilogue - tcgen05.ld 32x32b (thread = time row), bias/residual/scale/accumulate,
//              float4 stores: each warp-level store is 512 contiguous bytes of one output plane.
// Measured on B200 (tools/mma_probe2.cu, ALCM_TRACE): one M=128,K=16 MMA costs max(N/2, 32+N/4)
// cycles (tensor floor vs. 128 B/clk shared-memory operand reads); the MMA queue behind the issuing
// thread is only ~2 deep, so every mbarrier round trip (~200 cycles of wait/fence/commit/loop code)
// must be amortised over >= ~512 tensor cycles of MMAs - hence several taps per weight stage.
// ---------------------------------------------------------------------------------------------
struct ConvSmemLayout {
  uint32_t a_stage, w_blob, w_stage, a_off, w_off, bias_off, bar_off, total;
};
__host__ __device__ inline ConvSmemLayout conv_smem_layout(int kblk, int span, int NT, int w_stages, int tpg, int a_stages) {
  ConvSmemLayout L;
  L.a_stage = (uint32_t)kblk * (kTileM + span) * 16;
  L.w_blob = (uint32_t)kblk * NT * 16;
  L.w_stage = L.w_blob * tpg;
  L.a_off = 0;
  L.w_off = a_stages * L.a_stage;
  L.bias_off = L.w_off + w_stages * L.w_stage;
  L.bar_off = L.bias_off + NT * 4;
  L.total = L.bar_off + 8 * (2 * a_stages + 2 * w_stages + 4) + 16;
  return L;
}

struct ConvTile {
  int mt, nt, zb, zs, b, ph, kb0, kb1, q0;
};
__device__ __forceinline__ ConvTile conv_tile(const ConvArgs& a, int tile) {
  ConvTile t;
  // the K splits of one output tile are consecutive CTAs (= one cluster when the reduction goes through DSMEM)
  t.zs = tile % a.ksplit;
  const int tl = tile / a.ksplit;
  t.mt = tl % a.tiles_m;
  const int r = tl / a.tiles_m;
  t.nt = r % a.n_tiles;
  t.zb = r / a.n_tiles;
  t.b = t.zb / a.nphase;
  t.ph = t.zb % a.nphase;
  t.kb0 = (int)((long)t.zs * a.nkb / a.ksplit);
  t.kb1 = (int)((long)(t.zs + 1) * a.nkb / a.ksplit);
  t.q0 = t.mt * kTileM;
  return t;
}

// MMA-issuing warp.  NK2 = MMAs per (k-block, tap) (two 16-byte K chunks per MMA); static so that
// the burst is straight-line code (a rolled loop re-writes the uniform descriptor registers of
// in-flight UTCHMMAs and stalls ~300 cycles per trip).
template <int KIND, int NK2>
__device__ __forceinline__ void conv_mma_loop(const ConvArgs& a, uint32_t sA, uint32_t sW, const ConvSmemLayout& L, int rowsA,
                                              uint32_t tmem_base, uint32_t a_full, uint32_t a_empty, uint32_t w_full,
                                              uint32_t w_empty, uint32_t acc_full, uint32_t acc_empty, long long* trace) {
  const bool leader = elect_one();
  const uint32_t lead = leader ? 1u : 0u;
  // Descriptors differ only in their 14-bit start-address field: build the constant part once
  // and add (byte offset >> 4) per MMA.
  const uint64_t a_desc0 = umma_desc_kmajor(sA, rowsA * 16, 128);
  const uint64_t w_desc0 = umma_desc_kmajor(sW, a.NT * 16, 128);
  const uint32_t a_step = (uint32_t)(2 * rowsA), w_step = (uint32_t)(2 * a.NT);  // two K chunks, in 16 B units
  const uint32_t a_stage16 = L.a_stage >> 4, w_stage16 = L.w_stage >> 4, w_blob16 = L.w_blob >> 4;
  // every parameter the loop needs lives in a register: the asm memory clobbers would otherwise
  // make the compiler re-read the constant bank on each iteration of the single issuing warp.
  // The tap shift (row offset of tap j inside the A slab) is linear in j for every conv form here.
  const int ntaps = a.ntaps, S = a.w_stages, tpg = a.tpg, AS = a.a_stages;
  const uint32_t idesc = a.idesc;
  int ws = 0, as = 0;
  uint32_t wpar = 0, apar = 0;
  int it = 0;  // tiles done by this CTA
  for (int tile = blockIdx.x; tile < a.tiles_total; tile += gridDim.x, ++it) {
    const ConvTile T = conv_tile(a, tile);
    const uint32_t shift0 = (uint32_t)(a.tap_off[T.ph][0] - a.min_off[T.ph]);
    const uint64_t dshift = (uint64_t)(int64_t)(ntaps > 1 ? a.tap_off[T.ph][1] - a.tap_off[T.ph][0] : 0);
    const int st = (a.acc_stages > 1) ? (it & 1) : 0;
    const uint32_t tmem_d = tmem_base + (uint32_t)(st * a.NT);
    if (a.acc_stages > 1) {  // the epilogue must have drained this accumulator (two tiles ago)
      mbar_wait(acc_empty + 8 * st, ((it >> 1) & 1) ^ 1);
      tc_fence_after();
    }
    uint32_t acc = 0;
    for (int kb = T.kb0; kb < T.kb1; ++kb) {
      mbar_wait(a_full + 8 * as, apar);
      uint64_t a_tap = a_desc0 + (uint64_t)(as * a_stage16 + shift0);
      for (int j0 = 0; j0 < ntaps; j0 += tpg) {
        const int g = min(tpg, ntaps - j0);
        if (!a.w_resident || it == 0) mbar_wait(w_full + 8 * ws, wpar);
        tc_fence_after();
        if (trace && it == 0 && acc == 0 && leader) trace[3] = clock64();
        // branch-free, warp-uniform issue: the election is the predicate of the async instructions themselves
        // (a lane-private descriptor or a divergent region makes ptxas wrap every UTCHMMA in an
        // ELECT/R2UR.BROADCAST/BRA.U.ANY loop or shuttle descriptors through R2UR)
        uint64_t bd = w_desc0 + (uint64_t)(ws * w_stage16);
        for (int t = 0; t < g; ++t, a_tap += dshift, bd += w_blob16) {
#pragma unroll
          for (int i = 0; i < NK2; ++i)
            umma_ss_pred<KIND>(tmem_d, a_tap + (uint64_t)(i * a_step), bd + (uint64_t)(i * w_step), idesc, (i > 0) ? 1u : acc, lead);
          acc = 1;
        }
        if (!a.w_resident) tc_commit_pred(w_empty + 8 * ws, lead);  // frees the weight slot when these MMAs retire
        if (j0 + g == ntaps) tc_commit_pred(a_empty + 8 * as, lead);
        if (++ws == S) { ws = 0; wpar ^= 1; }
      }
      if (++as == AS) { as = 0; apar ^= 1; }
    }
    tc_commit_pred(acc_full + 8 * st, lead);
    if (trace && it == 0 && leader) trace[4] = clock64();
  }
  __syncwarp();
}

template <int KIND, int MINB>
__global__ void __launch_bounds__(192, MINB) conv_umma_kernel(const __grid_constant__ ConvArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ksplit = a.ksplit;
  const int S = a.w_stages;
  const int rowsA = kTileM + a.span;
  const int AS = a.a_stages;
  const ConvSmemLayout L = conv_smem_layout(a.kblk, a.span, a.NT, S, a.tpg, AS);
  const uint32_t sA = smem_u32(smem) + L.a_off;
  const uint32_t sW = smem_u32(smem) + L.w_off;
  const uint32_t bars = smem_u32(smem) + L.bar_off;
  float* s_bias = reinterpret_cast<float*>(smem + L.bias_off);
  // barrier slots: a_full[AS], a_empty[AS], w_full[S], w_empty[S], acc_full[2], acc_empty[2]
  const uint32_t a_full = bars, a_empty = bars + 8 * AS, w_full = bars + 16 * AS, w_empty = w_full + 8 * S;
  const uint32_t acc_full = w_empty + 8 * S, acc_empty = acc_full + 16;
  const int nbars = 2 * AS + 2 * S + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + L.bar_off + 8 * nbars);
  volatile uint32_t* s_last = tmem_slot + 1;  // split-K: "this CTA arrived last at its tile"

  long long* trace = a.trace ? a.trace + 8 * (size_t)blockIdx.x : nullptr;  // timestamps of the CTA's first tile
  if (threadIdx.x == 0) {
    if (trace) { trace[0] = (long long)global_timer_ns(); trace[1] = clock64(); }
    for (int i = 0; i < nbars - 2; ++i) mbar_init(bars + 8 * i, 1);
    mbar_init(acc_empty, 4);       // one arrival per epilogue warp
    mbar_init(acc_empty + 8, 4);
    fence_mbar_init();
    *s_last = 1;
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), a.tmem_cols);
    tmem_relinquish();
  }
  if (a.xg.nchunk < a.kchunks) {
    // Narrow operands (24 / 20 channels): the planes hold fewer 16-byte K chunks than the MMA's K (a multiple of two
    // chunks).  The missing chunk is never read from HBM - its slab is zeroed here, once, in every ring slot (such
    // launches have a single k-block, so no copy ever lands there), and made visible to the tensor core's async proxy.
    const uint32_t lo = (uint32_t)a.xg.nchunk * (uint32_t)rowsA, hi = (uint32_t)a.kblk * (uint32_t)rowsA;  // 16-byte units
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
  if (trace && threadIdx.x == 0) trace[2] = clock64();
  pdl_launch_dependents();  // the next kernel may start its own setup / weight prefetch now

  // Both asynchronous roles keep their warp CONVERGENT and predicate the async instructions with
  // elect.sync: UBLKCP / UTCHMMA / UTCBAR take warp-uniform operands, and issuing them from a
  // divergent `if (lane == 0)` makes the compiler wrap each one in an ELECT/R2UR/BRA serialisation
  // loop, which left the tensor pipe waiting on the issuing thread.
  if (warp == 0) {
    const bool leader = elect_one();
    const size_t plane_bytes = (size_t)a.xg.Tp * 16;
    const uint32_t a_pitch = (uint32_t)rowsA * 16;
    const int ntaps = a.ntaps, tpg = a.tpg;
    int ws = 0, as = 0;
    uint32_t wpar = 1, apar = 1;  // producer waits on the "previous" phase of the empty barriers first
    bool waited = false;
    bool w_needed = true;  // weights-resident launches copy the weights for the CTA's first tile only
    for (int tile = blockIdx.x; tile < a.tiles_total; tile += gridDim.x) {
      const ConvTile T = conv_tile(a, tile);
      const int row0 = T.q0 + a.min_off[T.ph] + a.xg.pad;  // >= 0: |min_off| <= pad
      const int nrows = min(rowsA, a.xg.Tp - row0);
      const uint8_t* wsrc = a.w + (size_t)T.b * a.w_batch_stride + (size_t)T.ph * a.w_phase_stride + ((size_t)T.nt * a.nkb + T.kb0) * ntaps * L.w_blob;
      const uint8_t* xsrc = a.x + (((size_t)T.b * a.xg.nchunk + (size_t)T.kb0 * a.kblk) * a.xg.Tp + row0) * 16;
      const uint32_t a_bytes = (uint32_t)nrows * 16;
      // one weight stage (a group of taps): wait for the slot, arm the barrier, one bulk copy
      auto load_w = [&](int j0) {
        const uint32_t bytes = (uint32_t)min(tpg, ntaps - j0) * L.w_blob;
        mbar_wait(w_empty + 8 * ws, wpar);
        if (leader) {
          if (a.dbg & 1) {
            mbar_arrive(w_full + 8 * ws);
          } else {
            mbar_expect_tx(w_full + 8 * ws, bytes);
            bulk_g2s(sW + ws * L.w_stage, wsrc, bytes, w_full + 8 * ws);
          }
        }
        wsrc += bytes;
        if (++ws == S) { ws = 0; wpar ^= 1; }
      };
      // Only the first weight group goes out before the first A slab: the first MMA needs exactly those two, and
      // everything queued ahead of the slab delays it (12 prefetched groups cost ~1.3 us of first-MMA latency).
      if (w_needed) load_w(0);
      if (!waited) { pdl_wait(); waited = true; }
      int gi = 0;
      for (int kb = T.kb0; kb < T.kb1; ++kb) {
        mbar_wait(a_empty + 8 * as, apar);
        if (a.dbg & 2) {
          if (leader) mbar_arrive(a_full + 8 * as);
          xsrc += plane_bytes * a.kblk;
        } else {
          const int cv = min(a.kblk, a.xg.nchunk - kb * a.kblk);  // K chunks of this k-block that exist in memory
          if (leader) mbar_expect_tx(a_full + 8 * as, (uint32_t)cv * a_bytes);
          uint32_t dst = sA + as * L.a_stage;
          for (int c = 0; c < cv; ++c, dst += a_pitch, xsrc += plane_bytes)
            if (leader) bulk_g2s(dst, xsrc, a_bytes, a_full + 8 * as);
          xsrc += plane_bytes * (a.kblk - cv);
        }
        for (int j0 = 0; j0 < ntaps; j0 += tpg, ++gi)
          if (gi >= 1 && w_needed) load_w(j0);
        if (++as == AS) { as = 0; apar ^= 1; }
      }
      if (a.w_resident) w_needed = false;
    }
    __syncwarp();
  } else if (warp == 1) {
    switch (a.kblk >> 1) {
      case 1: conv_mma_loop<KIND, 1>(a, sA, sW, L, rowsA, tmem_base, a_full, a_empty, w_full, w_empty, acc_full, acc_empty, trace); break;
      case 2: conv_mma_loop<KIND, 2>(a, sA, sW, L, rowsA, tmem_base, a_full, a_empty, w_full, w_empty, acc_full, acc_empty, trace); break;
      case 3: conv_mma_loop<KIND, 3>(a, sA, sW, L, rowsA, tmem_base, a_full, a_empty, w_full, w_empty, acc_full, acc_empty, trace); break;
      case 4: conv_mma_loop<KIND, 4>(a, sA, sW, L, rowsA, tmem_base, a_full, a_empty, w_full, w_empty, acc_full, acc_empty, trace); break;
      case 5: conv_mma_loop<KIND, 5>(a, sA, sW, L, rowsA, tmem_base, a_full, a_empty, w_full, w_empty, acc_full, acc_empty, trace); break;
      default: conv_mma_loop<KIND, 6>(a, sA, sW, L, rowsA, tmem_base, a_full, a_empty, w_full, w_empty, acc_full, acc_empty, trace); break;
    }
  } else if (warp < 6) {
    const int et = threadIdx.x - 64;
    const int qd = warp & 3;  // TMEM lane quarter this warp may access
    const int row = qd * 32 + lane;
    const float scale = a.scale;
    const int nq = a.NT >> 2;                              // float4 column groups in a tile
    const size_t plane4 = (size_t)a.og.Tp;                 // float4 units between consecutive chunks
    const float4* res4 = reinterpret_cast<const float4*>(a.res);
    float4* out4 = reinterpret_cast<float4*>(a.out);
    int it = 0, nt_staged = -1;
    for (int tile = blockIdx.x; tile < a.tiles_total; tile += gridDim.x, ++it) {
      const ConvTile T = conv_tile(a, tile);
      const int st = (a.acc_stages > 1) ? (it & 1) : 0;
      const uint32_t tmem_d = tmem_base + (uint32_t)(st * a.NT);
      // while the main loop runs: stage this N tile's bias in shared memory (the epilogue reads it as broadcasts)
      if (T.nt != nt_staged) {
        asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only: nobody still reads the old bias
        for (int i = et; i < a.NT; i += 128) s_bias[i] = a.bias ? __ldg(a.bias + T.nt * a.NT + i) : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        nt_staged = T.nt;
      }
      if (it == 0) pdl_wait();  // residual / accumulate reads, output and split-K workspace writes come after the previous kernel
      const int q = T.q0 + row;
      const bool in_seq = q >= 0 && q < a.M;
      const bool valid = in_seq;
      const size_t orow = (size_t)q * a.ostride + T.ph;
      const int nq_valid = min(nq, a.og.nchunk - T.nt * nq);   // column groups that exist in the output planes
      const size_t off0 = ((size_t)T.b * a.og.nchunk + (size_t)T.nt * nq) * a.og.Tp + a.og.pad + orow;  // float4 units
      const bool has_res = (a.res != nullptr) && in_seq, accum = (a.accum != 0) && valid, store = (a.out != nullptr) && valid;
      mbar_wait(acc_full + 8 * st, (it >> 1) & 1);
      tc_fence_after();
      if (trace && it == 0 && threadIdx.x == 64) trace[5] = clock64();
      // bias / residual / scale / accumulate of 16 consecutive output channels of this thread's row; fp32 store
      auto emit = [&](int c0, const float (&v)[16], const float4 (&rr)[4], const float4 (&oo)[4]) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int cq = (c0 >> 2) + g;
          const float4 bb = *reinterpret_cast<const float4*>(s_bias + c0 + 4 * g);
          float4 r = make_float4(v[4 * g] + bb.x, v[4 * g + 1] + bb.y, v[4 * g + 2] + bb.z, v[4 * g + 3] + bb.w);
          r.x = (r.x + rr[g].x) * scale + oo[g].x; r.y = (r.y + rr[g].y) * scale + oo[g].y;
          r.z = (r.z + rr[g].z) * scale + oo[g].z; r.w = (r.w + rr[g].w) * scale + oo[g].w;
          if (store && cq < nq_valid) out4[off0 + (size_t)cq * plane4] = r;
        }
      };
      // residual / accumulate operands of 16 channels (issued before the TMEM load they are combined with)
      auto fetch = [&](int c0, float4 (&rr)[4], float4 (&oo)[4]) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int cq = (c0 >> 2) + g;
          rr[g] = oo[g] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (cq < nq_valid) {
            if (has_res) rr[g] = res4[off0 + (size_t)cq * plane4];
            if (accum) oo[g] = out4[off0 + (size_t)cq * plane4];
          }
        }
      };
      if (ksplit == 1) {
        for (int c0 = 0; c0 < a.NT; c0 += 16) {
          float4 rr[4], oo[4];
          fetch(c0, rr, oo);
          uint32_t u[16];
          tmem_ld_x16(tmem_d + ((uint32_t)(qd * 32) << 16) + c0, u);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(u[i]);
          emit(c0, v, rr, oo);
        }
      } else if (a.cluster_splitk) {
        // K splits of this tile = the CTAs of this cluster.  Reduce-scatter through distributed shared memory: every CTA
        // owns NT/ksplit columns and receives the other CTAs' partial sums for them in its own (drained) pipeline
        // buffers: staging[source split][column group][row] float4.  Barrier A: all accumulators of the cluster are
        // complete, so every CTA's pipeline buffers may be overwritten; barrier B (after the role branches): all
        // pushes have landed.
        cluster_arrive();
        cluster_wait();
        const int nqs = nq / ksplit;
        const uint32_t stg = smem_u32(smem);
        for (int c0 = 0; c0 < a.NT; c0 += 16) {
          uint32_t u[16];
          tmem_ld_x16(tmem_d + ((uint32_t)(qd * 32) << 16) + c0, u);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int cq = (c0 >> 2) + g, owner = cq / nqs, cql = cq - owner * nqs;
            const uint32_t laddr = stg + (uint32_t)(((T.zs * nqs + cql) * kTileM + row) * 16);
            st_shared_cluster_f4(map_shared_rank(laddr, (uint32_t)owner),
                                 make_float4(__uint_as_float(u[4 * g]), __uint_as_float(u[4 * g + 1]), __uint_as_float(u[4 * g + 2]),
                                             __uint_as_float(u[4 * g + 3])));
          }
        }
        cluster_arrive();
        cluster_wait();
        // my column slice: sum the ksplit partials in split order (deterministic), then the usual epilogue
        const float4* st4 = reinterpret_cast<const float4*>(smem);
        for (int cql = 0; cql < nqs; ++cql) {
          const int cq = T.zs * nqs + cql;
          float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int s = 0; s < ksplit; ++s) {
            const float4 p = st4[(s * nqs + cql) * kTileM + row];
            r.x += p.x; r.y += p.y; r.z += p.z; r.w += p.w;
          }
          if (!(valid && cq < nq_valid)) continue;
          const float4 bb = *reinterpret_cast<const float4*>(s_bias + cq * 4);
          r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w;
          const size_t off = off0 + (size_t)cq * plane4;
          if (has_res) { const float4 rr = res4[off]; r.x += rr.x; r.y += rr.y; r.z += rr.z; r.w += rr.w; }
          r.x *= scale; r.y *= scale; r.z *= scale; r.w *= scale;
          if (accum) { const float4 oo = out4[off]; r.x += oo.x; r.y += oo.y; r.z += oo.z; r.w += oo.w; }
          if (a.out != nullptr) out4[off] = r;
        }
      } else {
        // partial tile -> workspace, [col/4][row] float4 so that a warp writes 512 contiguous bytes
        const size_t wtile = ((size_t)T.zb * a.n_tiles + T.nt) * a.tiles_m + T.mt;
        float4* wst = reinterpret_cast<float4*>(a.ws) + (wtile * ksplit) * (size_t)(nq * kTileM);
        float4* mine = wst + (size_t)T.zs * (nq * kTileM);
        for (int c0 = 0; c0 < a.NT; c0 += 16) {
          uint32_t u[16];
          tmem_ld_x16(tmem_d + ((uint32_t)(qd * 32) << 16) + c0, u);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g)
            mine[(size_t)((c0 >> 2) + g) * kTileM + row] =
                make_float4(__uint_as_float(u[4 * g]), __uint_as_float(u[4 * g + 1]), __uint_as_float(u[4 * g + 2]),
                            __uint_as_float(u[4 * g + 3]));
        }
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (threadIdx.x == 64) {
          const unsigned int old = atomicAdd(a.tile_ctr + wtile, 1u);
          const unsigned int last = (old == (unsigned int)(ksplit - 1));
          if (last) a.tile_ctr[wtile] = 0;  // every split has arrived: safe to re-arm for the next launch
          *s_last = last;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (*s_last) {
          __threadfence();
          for (int c0 = 0; c0 < a.NT; c0 += 16) {
            float4 rr[4], oo[4];
            fetch(c0, rr, oo);
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = 0.f;
            for (int s = 0; s < ksplit; ++s) {
              const float4* src = wst + (size_t)s * (nq * kTileM);
#pragma unroll
              for (int g = 0; g < 4; ++g) {
                const float4 p = __ldcg(src + (size_t)((c0 >> 2) + g) * kTileM + row);
                v[4 * g] += p.x; v[4 * g + 1] += p.y; v[4 * g + 2] += p.z; v[4 * g + 3] += p.w;
              }
            }
            emit(c0, v, rr, oo);
          }
        }
      }
      if (a.acc_stages > 1) {  // hand the accumulator back to the MMA warp
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + 8 * st);
      }
      if (trace && it == 0 && threadIdx.x == 64) trace[6] = clock64();
    }
  }
  if (ksplit > 1 && a.cluster_splitk && !(warp >= 2 && warp < 6)) {
    cluster_arrive();  // barrier A and B of the DSMEM split-K reduction (the epilogue warps arrive inside their branch)
    cluster_wait();
    cluster_arrive();
    cluster_wait();
  }
  if (trace && threadIdx.x == 64) trace[7] = (long long)global_timer_ns();
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
  pdl_launch_dependents();
  pdl_wait();
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

// Host side of libaudiolcm_b200: weight preparation, per-(B,T) plans (buffers + kernel list +
// CUDA graph) and the C-ABI declared in include/audiolcm_b200.h.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/audiolcm_b200.h"
#include "act1d.cuh"
#include "common.cuh"
#include "conv.cuh"
#include "misc_kernels.cuh"

using namespace alcm;

// ------------------------------------------------------------------------------------------ errors
static thread_local std::string g_err = "";
struct AlcmError : std::runtime_error {
  int code;
  AlcmError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
#define CUDA_CHECK(expr)                                                                                   \
  do {                                                                                                     \
    cudaError_t _e = (expr);                                                                               \
    if (_e != cudaSuccess)                                                                                 \
      throw AlcmError(ALCM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + ":" + \
                                         std::to_string(__LINE__) + ")");                                  \
  } while (0)
#define REQUIRE(cond, msg)                                              \
  do {                                                                  \
    if (!(cond)) throw AlcmError(ALCM_ERR_INVALID, std::string(msg));   \
  } while (0)

template <class F>
static int guarded(F&& f) {
  try {
    f();
    return ALCM_OK;
  } catch (const AlcmError& e) {
    g_err = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_err = e.what();
    return ALCM_ERR_INTERNAL;
  }
}

// ------------------------------------------------------------------------------------------ ctx
static int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

// Run-time knobs.  Read ONCE - when a model handle is created, or at the top of a
// single-op entry point - and carried by value from there: plan-time decisions (N tile, split-K, workspace sizes)
// and launch-time geometry come from the same snapshot, and no launch ever calls getenv.
struct Knobs {
  int pdl, graph, lanes, lane_waves, persist, cluster_splitk, splitk, retile, retile_mode;
  int smem_budget, smem_budget_1w, smem_budget_mw, tpg, astages, kblk_max, nt, tmem2;
  int act_variant, w_resident, attn_tc, gn_fused, max_plans, trace, guard, nt192;
  static Knobs from_env() {
    Knobs k;
    k.pdl = env_int("ALCM_PDL", -1);  // -1 unset (per-plan default), 0 never, 1 always
    k.graph = env_int("ALCM_GRAPH", 1);
    k.lanes = env_int("ALCM_LANES", 1);
    k.lane_waves = env_int("ALCM_LANE_WAVES", 20);
    k.persist = env_int("ALCM_PERSIST", 1);
    k.cluster_splitk = env_int("ALCM_CLUSTER_SPLITK", 1);
    k.splitk = env_int("ALCM_SPLITK", 1);
    k.retile = env_int("ALCM_RETILE", 1);
    k.retile_mode = env_int("ALCM_RETILE_MODE", 1);
    k.smem_budget = env_int("ALCM_SMEM_BUDGET", 0);
    k.smem_budget_1w = env_int("ALCM_SMEM_BUDGET_1W", 120 * 1000);
    k.smem_budget_mw = env_int("ALCM_SMEM_BUDGET_MW", 100 * 1024);
    k.tpg = env_int("ALCM_TPG", 0);
    k.astages = env_int("ALCM_ASTAGES", 0);
    k.kblk_max = env_int("ALCM_KBLK_MAX", 8);
    k.nt = env_int("ALCM_NT", 128);
    k.tmem2 = env_int("ALCM_TMEM2", 0);
    k.act_variant = env_int("ALCM_ACT_VARIANT", -1);
    k.w_resident = env_int("ALCM_W_RESIDENT", 1);         // persistent narrow-stage convs keep all their weights in shared memory
    k.attn_tc = env_int("ALCM_ATTN_TC", 1);               // attention GEMMs on conv_umma_kernel (0: CUDA-core kernels)
    k.gn_fused = env_int("ALCM_GN_FUSED", 1);
    k.max_plans = std::max(1, env_int("ALCM_MAX_PLANS", 16));
    k.trace = env_int("ALCM_TRACE", 0);
    k.nt192 = env_int("ALCM_NT192", 96);                  // N tile of the 192-channel layers: two persistent 96-wide tiles (stage-3 convs at batch 64: 710 -> 893 TFLOP/s) or one 192-wide
    k.guard = env_int("ALCM_GUARD", 0);                   // 1: 4 KB zero guard zones between device buffers, verified by alcm_*_check_guards
    return k;
  }
};

// Per-device context.  Everything the host side caches about the device lives here (no process-global mutable
// state): two ctxs - two devices, or two host threads on one device - never share anything.
struct alcm_ctx {
  int device = 0;
  int sm_count = 148;
  std::mutex mu;                             // guards cluster_cap (a ctx may be shared by host threads)
  long cluster_cap[2][9] = {{0}};            // co-resident CTAs when launched as clusters of n (per operand type)
  std::atomic<unsigned long long> plan_clock{0};  // LRU stamps of the per-shape plans
  // micro-benchmark instrumentation (alcm_bench_conv only)
  int conv_dbg = 0;
  long long* conv_trace = nullptr;
  int conv_last_grid = 0;
};

// Device memory owner.  Three modes:
//   * plain: one cudaMalloc per allocation (weights, single-op entry points);
//   * measuring: hands out fake addresses and only adds up sizes (the sizing pass of a plan);
//   * slab: sub-allocates from ONE stream-ordered allocation (cudaMallocAsync) that was zero-filled once - the
//     buffers of a (B,T) plan.  Retiring a plan is a cudaFreeAsync: no device-wide synchronisation anywhere.
struct Arena {
  std::vector<void*> ptrs;
  size_t total = 0;
  bool measuring = false;
  uint8_t* slab = nullptr;
  size_t slab_bytes = 0, off = 0;
  cudaStream_t slab_stream = nullptr;
  // ALCM_GUARD=1 (self-check mode; compute-sanitizer is not available on every pool): every buffer is followed - and,
  // outside slabs, preceded - by a kGuard-byte zone of zeros that no kernel may touch; guard_violations() counts the
  // bytes that changed.  An out-of-bounds WRITE of any kernel shows up there (or in a neighbour's zero halo rows, which
  // the repeat-call bit-identity tests catch).
  static constexpr size_t kGuard = 4096;
  bool guard = false;
  bool f16 = false;   // 16-bit planes made from this arena are fp16 (precision "fp16"), else bf16
  std::vector<std::pair<uint8_t*, size_t>> gaps;
  static size_t align_up(size_t b) { return (std::max<size_t>(b, 16) + 255) & ~(size_t)255; }
  void* alloc(size_t bytes, bool zero = true) {
    const size_t g = guard ? kGuard : 0;
    if (measuring) {
      void* p = reinterpret_cast<void*>((uintptr_t)0x10000 + off);
      off += align_up(bytes) + g;
      total = off;
      return p;
    }
    if (slab) {
      const size_t b = align_up(bytes);
      if (off + b + g > slab_bytes) throw AlcmError(ALCM_ERR_INTERNAL, "plan slab overflow: the sizing pass and the build pass disagree");
      uint8_t* p = slab + off;
      off += b + g;
      if (g) gaps.emplace_back(p + b, g);   // the slab was zero-filled as a whole
      return p;
    }
    void* p = nullptr;
    bytes = std::max<size_t>(bytes, 16);
    const size_t b = guard ? align_up(bytes) : bytes;
    CUDA_CHECK(cudaMalloc(&p, b + 2 * g));
    if (zero) CUDA_CHECK(cudaMemset(p, 0, b + 2 * g));
    else if (g) { CUDA_CHECK(cudaMemset(p, 0, g)); CUDA_CHECK(cudaMemset(static_cast<uint8_t*>(p) + g + b, 0, g)); }
    ptrs.push_back(p);
    total += b + 2 * g;
    if (g) { gaps.emplace_back(static_cast<uint8_t*>(p), g); gaps.emplace_back(static_cast<uint8_t*>(p) + g + b, g); }
    return static_cast<uint8_t*>(p) + g;
  }
  long long guard_violations() const;  // bytes of the guard zones that are no longer zero (synchronises)
  void reserve(size_t bytes, cudaStream_t st) {  // slab mode: one stream-ordered allocation, zeroed once
    bytes = align_up(bytes);
    void* p = nullptr;
    CUDA_CHECK(cudaMallocAsync(&p, bytes, st));
    slab = static_cast<uint8_t*>(p);
    slab_bytes = bytes; slab_stream = st; total = bytes; off = 0;
    CUDA_CHECK(cudaMemsetAsync(p, 0, bytes, st));
  }
  void release_async(cudaStream_t st) {  // after the plan's last use has been ordered before `st`
    if (slab) cudaFreeAsync(slab, st);
    slab = nullptr;
  }
  void release() {
    for (void* p : ptrs) cudaFree(p);
    ptrs.clear();
    if (slab) cudaFreeAsync(slab, nullptr);  // owners synchronise the device before destroying a plan this way
    slab = nullptr;
    total = 0;
  }
  ~Arena() { release(); }
};

struct PlaneT {
  uint8_t* p = nullptr;
  int B = 0, C = 0, T = 0, esz = 4;
  PlaneGeom g{0, 0, 0};
  size_t bytes = 0;
  float* f() const { return reinterpret_cast<float*>(p); }
};

// Channels held in memory per time step.  The MMA consumes K in 16-channel steps (bf16) and the bf16 planes pack 8
// channels per unit, so planes are padded to a multiple of 8 - NOT 16: for the narrow tensors (C <= 48, i.e. the
// 24-channel last vocoder stage and the 20-channel latent) the missing K chunk of a conv operand is never stored
// in, written to or read from HBM; the conv kernel keeps an all-zero shared-memory slab for it instead (every such
// launch has a single k-block).  Wider tensors are padded to 16 (at most 8 of > 48 channels).
static const bool g_unpadded = env_int("ALCM_UNPADDED", 1) != 0;  // 0: 16-channel padding everywhere (A/B measurements); constant for the process
static inline int plane_cpad(int C) { return (g_unpadded && round_up(C, 16) <= 48) ? round_up(C, 8) : round_up(C, 16); }

static PlaneT make_planes(Arena& ar, int B, int C, int T, int esz) {
  PlaneT t;
  t.B = B; t.C = C; t.T = T; t.esz = esz;
  const int E = 16 / esz;
  t.g.nchunk = plane_cpad(C) / E;
  t.g.pad = kPad;
  t.g.Tp = T + 2 * kPad;
  t.g.fmt = esz == 2 ? (ar.f16 ? kFmtF16 : kFmtBF16) : kFmtF32;
  t.bytes = (size_t)B * t.g.nchunk * t.g.Tp * 16;
  t.p = static_cast<uint8_t*>(ar.alloc(t.bytes, true));
  return t;
}

static inline bool is16(int prec) { return prec == ALCM_PREC_BF16 || prec == ALCM_PREC_FP16; }   // 16-bit operands, kind::f16
static inline int opnd_esz(int prec) { return is16(prec) ? 2 : 4; }
static inline int umma_fmt(int prec) { return prec == ALCM_PREC_FP16 ? 0 : prec == ALCM_PREC_BF16 ? 1 : 2; }   // idesc a/b format

long long Arena::guard_violations() const {
  if (gaps.empty()) return 0;
  unsigned long long* d = nullptr;
  CUDA_CHECK(cudaMalloc(&d, sizeof(*d)));
  CUDA_CHECK(cudaMemset(d, 0, sizeof(*d)));
  CUDA_CHECK(cudaStreamSynchronize(nullptr));
  for (const auto& g : gaps) {
    count_nonzero_kernel<<<4, 256>>>(reinterpret_cast<const uint4*>(g.first), g.second / 16, d);
    CUDA_CHECK(cudaGetLastError());
  }
  unsigned long long h = 0;
  CUDA_CHECK(cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost));
  cudaFree(d);
  return (long long)h;
}

// ------------------------------------------------------------------------------------------ launches
// Every kernel of a plan is launched with the programmatic-stream-serialization attribute (PDL): it may be
// scheduled while its stream predecessor drains and synchronises itself with griddepcontrol.wait (pdl_wait()).
// Under stream capture these become programmatic edges of the CUDA graph.  ALCM_PDL=0 turns it off.
// Measured (round 1, batch-1 decode, CUDA-graph replay): 4.00 ms with PDL on every kernel, 3.97 ms with PDL on
// the VAE chain only, 3.89 ms without - graph kernel->kernel edges are already cheap, and early-resident
// dependents (up to 200 KB of shared memory each, idle in griddepcontrol.wait) keep other lanes' CTAs off
// the SMs.  So it is off by default and a per-plan choice (OpList::pdl, or ALCM_PDL=1 to force it).
static thread_local int t_pdl = 0;        // set by OpList::run / run_lanes around the launches of a plan
static thread_local int t_cluster_x = 1;  // >1: the next launch_k() launches thread-block clusters of this many CTAs (x)
template <typename... KArgs, typename... Args>
static void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (t_pdl) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (t_cluster_x > 1) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = (unsigned)t_cluster_x; at[n].val.clusterDim.y = 1; at[n].val.clusterDim.z = 1;
    ++n;
    t_cluster_x = 1;
  }
  cfg.attrs = at;
  cfg.numAttrs = n;
  CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...));
}

static cudaError_t sync_setup();

// ------------------------------------------------------------------------------------------ conv layers
enum ConvKind { KIND_CONV = 0, KIND_CONVT = 1, KIND_UPCONV3 = 2, KIND_DOWN2 = 3 };

struct ConvLayer {
  int Cin = 0, Cout = 0, nphase = 1, ntaps = 1;
  int tap_off[kMaxPhase][kMaxTaps];
  int min_off[kMaxPhase];
  int span = 0;
  float* weff = nullptr;   // fp32 [nphase][ntaps][Cout][Cin]   (kept only for ALCM_PREC_FP32)
  uint8_t* wpack = nullptr;
  size_t phase_stride = 0;
  float* bias = nullptr;   // padded
  int NT = 0, n_tiles = 0, kchunks = 0, kblk = 0, nkb = 0, tmem_cols = 0, w_stages = 0;
  uint32_t idesc = 0, smem = 0;
  int prec = 0;
  size_t w_batch_stride = 0;  // > 0: per-batch-item weights (dynamic_conv)
};

// Tile geometry of a tcgen05 conv layer (N tile, K chunks / blocks, TMEM columns, instruction descriptor) from its
// channel counts, taps and arithmetic mode.
static void conv_tiling(ConvLayer& L, const Knobs& K_) {
  const int prec = L.prec, Cout = L.Cout, Cin = L.Cin;
  const int E = is16(prec) ? 8 : 4;
  const int cout_pad = round_up(Cout, 16);
  int nt_pref = K_.nt;
  if (cout_pad == 192 && K_.nt192 > 0 && 192 % K_.nt192 == 0) nt_pref = K_.nt192 < 192 ? K_.nt192 : nt_pref;
  if (cout_pad <= 256 && (cout_pad <= nt_pref || cout_pad % nt_pref != 0)) L.NT = cout_pad;
  else L.NT = nt_pref;
  REQUIRE(L.NT % 16 == 0 && L.NT >= 16 && L.NT <= 256, "bad N tile");
  L.n_tiles = (cout_pad + L.NT - 1) / L.NT;
  L.tmem_cols = 32;
  while (L.tmem_cols < L.NT * (K_.tmem2 ? 2 : 1)) L.tmem_cols *= 2;
  L.kchunks = round_up(Cin, 16) / E;
  L.kblk = 0;
  if (L.kchunks <= 12) L.kblk = L.kchunks;
  else for (int d = std::min(12, K_.kblk_max); d >= 2; d -= 2) if (L.kchunks % d == 0) { L.kblk = d; break; }
  REQUIRE(L.kblk >= 2 && L.kblk % 2 == 0 && L.kblk <= 12, "bad k-block");
  L.nkb = L.kchunks / L.kblk;
  L.idesc = umma_idesc(umma_fmt(prec), L.NT);
  L.w_stages = 0;  // chosen per launch (pick_pipeline)
  L.smem = 0;
  L.phase_stride = (size_t)L.n_tiles * L.nkb * L.ntaps * L.kblk * L.NT * 16;
}

// A 1x1 "conv" whose weights are a per-batch-item operand written by pack_dyn_w_kernel at run time (attention): only the
// geometry is fixed here; wpack points into the plan's slab, w_batch_stride separates the items.
static ConvLayer dynamic_conv(const Knobs& K_, int prec, int Cout, int Cin) {
  ConvLayer L;
  L.Cin = Cin; L.Cout = Cout; L.prec = prec;
  L.nphase = 1; L.ntaps = 1;
  memset(L.tap_off, 0, sizeof(L.tap_off));
  memset(L.min_off, 0, sizeof(L.min_off));
  L.span = 0;
  conv_tiling(L, K_);
  L.w_batch_stride = L.phase_stride;
  return L;
}

// w: folded weight on device (Conv1d [Cout,Cin,K] or ConvTranspose1d [Cin,Cout,K]); bias may be null
static ConvLayer prepare_conv(Arena& ar, const Knobs& K_, int prec, ConvKind kind, const float* w, const float* bias, int Cout, int Cin, int K,
                              int dil_or_stride) {
  const int Cin_src = Cin;
  if (kind == KIND_DOWN2) Cin = 2 * Cin;   // Downsample1D on the time-folded input (s2d_cast_kernel): 2C channels, taps n and n+1
  ConvLayer L;
  L.Cin = Cin; L.Cout = Cout; L.prec = prec;
  WeffRecipe rc;
  memset(&rc, 0xff, sizeof(rc));  // all -1
  if (kind == KIND_CONV) {
    const int d = dil_or_stride;
    REQUIRE(K >= 1 && K <= kMaxTaps && (K % 2) == 1, "conv1d: odd kernel size <= 11 required");
    const int pad = (K * d - d) / 2;
    REQUIRE(pad <= kPad, "conv1d: padding exceeds plane halo");
    L.nphase = 1; L.ntaps = K;
    for (int j = 0; j < K; ++j) { L.tap_off[0][j] = j * d - pad; rc.src_k[0][j][0] = j; }
  } else if (kind == KIND_CONVT) {
    const int u = dil_or_stride;
    REQUIRE(K == 2 * u && (u % 2) == 0 && u <= kMaxPhase, "conv_transpose1d: kernel 2u, even stride u<=4, padding u/2 only");
    L.nphase = u; L.ntaps = 2;
    for (int r = 0; r < u; ++r) {
      const int s = r + u / 2;
      if (s < u) { L.tap_off[r][0] = 0; rc.src_k[r][0][0] = s; L.tap_off[r][1] = -1; rc.src_k[r][1][0] = s + u; }
      else       { L.tap_off[r][0] = 1; rc.src_k[r][0][0] = s - u; L.tap_off[r][1] = 0; rc.src_k[r][1][0] = s; }
    }
  } else if (kind == KIND_DOWN2) {
    REQUIRE(K == 3 && dil_or_stride == 2, "downsample conv: k=3, stride 2 only");
    L.nphase = 1; L.ntaps = 2;
    L.tap_off[0][0] = 0; L.tap_off[0][1] = 1;
  } else {
    REQUIRE(K == 3, "upsample conv: k=3 only");
    L.nphase = 2; L.ntaps = 2;
    L.tap_off[0][0] = -1; rc.src_k[0][0][0] = 0;
    L.tap_off[0][1] = 0;  rc.src_k[0][1][0] = 1; rc.src_k[0][1][1] = 2;
    L.tap_off[1][0] = 0;  rc.src_k[1][0][0] = 0; rc.src_k[1][0][1] = 1;
    L.tap_off[1][1] = 1;  rc.src_k[1][1][0] = 2;
  }
  rc.nphase = L.nphase; rc.ntaps = L.ntaps;
  L.span = 0;
  for (int p = 0; p < L.nphase; ++p) {
    int mn = L.tap_off[p][0], mx = L.tap_off[p][0];
    for (int j = 1; j < L.ntaps; ++j) { mn = std::min(mn, L.tap_off[p][j]); mx = std::max(mx, L.tap_off[p][j]); }
    L.min_off[p] = mn;
    L.span = std::max(L.span, mx - mn);
  }
  // effective weights
  const size_t nweff = (size_t)L.nphase * L.ntaps * Cout * Cin;
  Arena tmp;
  Arena& weff_owner = (prec == ALCM_PREC_FP32) ? ar : tmp;
  L.weff = static_cast<float*>(weff_owner.alloc(nweff * 4, false));
  if (kind == KIND_DOWN2) weff_down2_kernel<<<(unsigned)std::min<size_t>((nweff + 255) / 256, 4096), 256>>>(w, L.weff, Cout, Cin_src);
  else weff_kernel<<<(unsigned)std::min<size_t>((nweff + 255) / 256, 4096), 256>>>(w, L.weff, rc, Cout, Cin, K, kind == KIND_CONVT);
  CUDA_CHECK(cudaGetLastError());

  if (prec == ALCM_PREC_FP32) {
    L.bias = nullptr;
    if (bias) {
      L.bias = static_cast<float*>(ar.alloc((size_t)Cout * 4, false));
      CUDA_CHECK(cudaMemcpy(L.bias, bias, (size_t)Cout * 4, cudaMemcpyDeviceToDevice));
    }
    CUDA_CHECK(sync_setup());
    return L;
  }
  const int E = is16(prec) ? 8 : 4;
  conv_tiling(L, K_);
  L.wpack = static_cast<uint8_t*>(ar.alloc(L.phase_stride * L.nphase, false));
  const size_t units = L.phase_stride * L.nphase / 16;
  const unsigned blocks = (unsigned)std::min<size_t>((units + 255) / 256, 8192);
  if (E == 8) pack_w_kernel<8><<<blocks, 256>>>(L.weff, L.wpack, L.nphase, L.ntaps, Cout, Cin, L.NT, L.n_tiles, L.kblk, L.nkb, prec == ALCM_PREC_FP16);
  else pack_w_kernel<4><<<blocks, 256>>>(L.weff, L.wpack, L.nphase, L.ntaps, Cout, Cin, L.NT, L.n_tiles, L.kblk, L.nkb, 0);
  CUDA_CHECK(cudaGetLastError());
  L.bias = static_cast<float*>(ar.alloc((size_t)L.n_tiles * L.NT * 4, true));
  if (bias) CUDA_CHECK(cudaMemcpy(L.bias, bias, (size_t)Cout * 4, cudaMemcpyDeviceToDevice));
  CUDA_CHECK(sync_setup());
  L.weff = nullptr;  // tmp arena frees it
  return L;
}

// Everything a planning decision depends on: the device context and the knob snapshot of the model (or entry point).
struct Env {
  alcm_ctx* cx = nullptr;
  Knobs k;
  int sms() const { return cx->sm_count; }
};

// ---- per-launch N tile -----------------------------------------------------------------------------------
// Weights are packed for N = 128 (or the whole padded Cout when smaller).  A launch with few time tiles
// (the VAE at T = 312/624, conv_pre) fills the GPU better with narrower N tiles: more CTAs per launch and a
// split-K factor of at most 2, so the last-arriver fix-up reads 2 small partial tiles instead of 4-8 big ones.
static int pick_nt(const Env& env, const ConvLayer& L, int M, int B) {
  if (L.prec == ALCM_PREC_FP32 || !env.k.retile) return L.NT;
  const int cout_pad = round_up(L.Cout, 16);
  const long m = (long)((M + kTileM - 1) / kTileM) * B * L.nphase;
  const long want = (long)(0.6 * env.sms());
  // Measured (ALCM_TRACE, dbg flags): these launches are bound by fixed per-CTA costs and the split-K fix-up, not by
  // operand traffic - so first look for the widest tile that fills the GPU WITHOUT a K split, then with a split of 2.
  int best = L.NT;
  for (int pass = (env.k.retile_mode ? 0 : 1); pass < 2; ++pass) {
    for (int nt : {L.NT, 64, 32}) {
      if (nt > L.NT || cout_pad % nt != 0) continue;
      best = nt;
      const long ctas = m * (cout_pad / nt);
      if (ctas >= want || (pass == 1 && L.nkb >= 2 && 2 * ctas >= want)) return best;
    }
  }
  return best;
}

// Tile shape of a launch whose output has few time tiles (the VAE at T = 312/624, conv_pre): N tile and K split
// chosen together with a small cost model of one CTA - main loop (MMAs at max(N/2, 32+N/4) cycles, ~300 cycles per
// mbarrier round trip) plus the cluster reduce-scatter (two cluster barriers, N/16 TMEM->DSMEM pushes) - among the
// shapes that fit one wave.  Measured: at these sizes the launch is bound by per-CTA fixed costs, not by operand
// traffic, so wide tiles (efficient MMAs) with a split of 4-8 reduced through DSMEM beat narrow un-split ones.
static void conv_kernel_for(int prec, void (**kern)(ConvArgs), int* threads);

// CTAs that can be co-resident when launched as clusters of `ks` (GPC boundaries cost a few SMs): one big-smem CTA per SM.
// Cached in the ctx (per device).
static long cluster_capacity(const Env& env, int prec, int ks) {
  long (&cap)[2][9] = env.cx->cluster_cap;
  const int pi = is16(prec) ? 0 : 1;
  if (ks <= 1) return env.sms();
  std::lock_guard<std::mutex> lock(env.cx->mu);
  if (cap[pi][ks] == 0) {
    void (*kern)(ConvArgs) = nullptr;
    int threads = 192;
    conv_kernel_for(prec, &kern, &threads);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(ks * 64); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 200 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)ks; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) { cudaGetLastError(); n = env.sms() / ks * 3 / 4; }
    cap[pi][ks] = (long)n * ks;
  }
  return cap[pi][ks];
}

static bool choose_cluster_tile(const Env& env, const ConvLayer& L, int M, int B, int* nt_out, int* ks_out) {
  if (L.prec == ALCM_PREC_FP32 || !env.k.cluster_splitk) return false;
  const int cout_pad = round_up(L.Cout, 16);
  const long m = (long)((M + kTileM - 1) / kTileM) * B * L.nphase;
  if (m * ((cout_pad + L.NT - 1) / L.NT) * 2 > (long)env.sms()) return false;  // the default tiling already fills the GPU
  double best = 1e30;
  bool found = false;
  for (int nt : {256, 128, 64, 32}) {
    if (cout_pad % nt != 0) continue;
    for (int ks : {1, 2, 4, 8}) {
      if (ks > L.nkb || (nt / 4) % ks != 0) continue;
      const long ctas = m * (cout_pad / nt) * ks;
      if (ctas > cluster_capacity(env, L.prec, ks)) continue;
      const double cyc_mma = std::max(nt / 2.0, 32.0 + nt / 4.0);
      const int kbs = (L.nkb + ks - 1) / ks;
      double cost = kbs * (L.ntaps * (L.kblk / 2) * cyc_mma + 300.0);
      if (ks > 1) cost += 900.0 + (nt / 16) * 70.0 + (nt / 4 / ks) * ks * 8.0;
      cost += (nt / 16) * 40.0 / (ks > 1 ? ks : 1);                                  // epilogue stores
      cost *= std::max(1.0, 0.6 * env.sms() / (double)ctas);                         // idle SMs
      if (cost < best) { best = cost; *nt_out = nt; *ks_out = ks; found = true; }
    }
  }
  return found;
}

struct RetileCache {  // owned by a model: re-tiled copies of its layers, keyed by (packed weights, N tile)
  std::map<std::pair<const void*, int>, ConvLayer> m;
};

static const ConvLayer& retile(Arena& ar, RetileCache& cache, const ConvLayer& L, int NT2) {
  if (NT2 == L.NT) return L;
  auto key = std::make_pair((const void*)L.wpack, NT2);
  auto it = cache.m.find(key);
  if (it != cache.m.end()) return it->second;
  ConvLayer R = L;
  const int cout_pad = round_up(L.Cout, 16);
  R.NT = NT2;
  R.n_tiles = (cout_pad + NT2 - 1) / NT2;
  R.tmem_cols = 32;
  while (R.tmem_cols < NT2) R.tmem_cols *= 2;
  R.idesc = umma_idesc(umma_fmt(L.prec), NT2);
  R.phase_stride = (size_t)R.n_tiles * L.nkb * L.ntaps * L.kblk * NT2 * 16;
  R.wpack = static_cast<uint8_t*>(ar.alloc(R.phase_stride * L.nphase, false));
  const size_t units = R.phase_stride * L.nphase / 16;
  repack_nt_kernel<<<(unsigned)std::min<size_t>((units + 255) / 256, 8192), 256>>>(
      reinterpret_cast<const uint4*>(L.wpack), reinterpret_cast<uint4*>(R.wpack), L.nphase, L.ntaps, L.kblk, L.nkb, L.NT, L.n_tiles,
      NT2, R.n_tiles);
  CUDA_CHECK(cudaGetLastError());
  R.bias = static_cast<float*>(ar.alloc((size_t)R.n_tiles * NT2 * 4, true));
  CUDA_CHECK(cudaMemcpy(R.bias, L.bias, (size_t)std::min(L.n_tiles * L.NT, R.n_tiles * NT2) * 4, cudaMemcpyDeviceToDevice));
  CUDA_CHECK(sync_setup());
  return cache.m.emplace(key, R).first->second;
}

// weight_norm fold into a temp buffer (dim0 x inner)
static float* fold_wn(Arena& tmp, const float* g, const float* v, int dim0, int inner) {
  float* w = static_cast<float*>(tmp.alloc((size_t)dim0 * inner * 4, false));
  wn_fold_kernel<<<dim0, 256>>>(v, g, w, inner);
  CUDA_CHECK(cudaGetLastError());
  return w;
}

// ------------------------------------------------------------------------------------------ ops
enum { OP_FORK = -1, OP_JOIN = -2 };  // markers: ops between them carry a lane (independent chains)
constexpr int kMaxLanes = 4;
struct Op {
  int cls;
  double flops, bytes;
  std::function<void(cudaStream_t)> fn;
  int lane = 0;
  int stage = -1;  // vocoder stage (0-based) / -1: VAE or glue; for the per-stage roofline of alcm_profile_stages
};


// Pipeline shape of one launch.  Taps per weight stage: enough MMAs per mbarrier round trip to cover
// ~1024 tensor cycles (see conv.cuh).  Ring depth: ~100 KB (2 CTAs/SM) for multi-wave grids; single-wave grids
// get 120 KB - measured on the batch-1 decode (tools/sweep_env.sh): 200 KB 3.450 ms, 150 KB 3.496, 120 KB 3.424
// (operand traffic is not the limiter at that size, and a smaller footprint leaves room for the other lanes' blocks).
static void pick_pipeline(const Env& env, const ConvLayer& L, long ctas, int ksplit, bool cluster, int* stages, int* tpg,
                          int* a_stages, uint32_t* smem) {
  const uint32_t budget = (uint32_t)(env.k.smem_budget > 0 ? env.k.smem_budget
                                                           : (ctas <= (long)env.sms() ? env.k.smem_budget_1w : env.k.smem_budget_mw));
  const double cyc_mma = std::max(L.NT / 2.0, 32.0 + L.NT / 4.0);
  const double cyc_tap = (L.kblk / 2) * cyc_mma;
  // ~1024 tensor cycles per weight stage (measured: stage-1 conv at batch 1, taps per stage 1/2/3/4 -> 802/978/1039/1046 TFLOP/s)
  int tmin = std::max(1, std::min(L.ntaps, (int)std::ceil(1024.0 / cyc_tap)));
  const int ngrp = (L.ntaps + tmin - 1) / tmin;
  int t = (L.ntaps + ngrp - 1) / ngrp;
  if (env.k.tpg > 0) t = std::min(L.ntaps, env.k.tpg);
  // A-slab ring: a k-block whose MMAs take less than the slab's load latency (~1300 cycles) needs more than 2 slabs in flight
  const uint32_t a1 = (uint32_t)L.kblk * (kTileM + L.span) * 16, blob = (uint32_t)L.kblk * L.NT * 16;
  int AS = (L.ntaps * cyc_tap < 1500.0) ? 4 : 2;
  if (env.k.astages > 0) AS = env.k.astages;
  AS = std::max(2, std::min(4, AS));
  while (AS > 2 && AS * a1 > budget / 2) --AS;
  const int nkb_local = (L.nkb + ksplit - 1) / ksplit;
  AS = std::max(2, std::min(AS, nkb_local));
  const uint32_t a2 = AS * a1, fixed = a2 + L.NT * 4 + 512;
  while (t > 1 && fixed + 2u * t * blob > budget) --t;
  int S = (budget > fixed) ? (int)((budget - fixed) / (t * blob)) : 0;
  S = std::max(2, std::min(12, S));
  S = std::min(S, std::max(2, nkb_local * ((L.ntaps + t - 1) / t)));
  if (cluster) {  // DSMEM split-K: the partial tiles pushed by the other CTAs overlay the (drained) A and W buffers
    const uint32_t staging = (uint32_t)L.NT * 512u;
    while (a2 + (uint32_t)S * t * blob < staging && S < 12 && fixed + (uint32_t)(S + 1) * t * blob <= 220u * 1024u) ++S;
    REQUIRE(a2 + (uint32_t)S * t * blob >= staging, "cluster split-K: staging does not fit under the pipeline buffers");
  }
  *stages = S;
  *tpg = t;
  *a_stages = AS;
  *smem = conv_smem_layout(L.kblk, L.span, L.NT, S, t, AS).total;
  REQUIRE(*smem <= 227 * 1024, "conv tile does not fit shared memory");
}

struct SplitK {  // per-launch split-K resources (see ConvArgs::ksplit)
  int ksplit = 1;
  int cluster = 0;  // reduce through a thread-block cluster's distributed shared memory instead of the global workspace
  float* ws = nullptr;
  unsigned int* ctr = nullptr;
};

static inline int conv_m_tiles(int M) { return (M + kTileM - 1) / kTileM; }

// How many K splits a conv launch gets: only launches whose output tiles cannot fill the GPU are split.
static int pick_ksplit(const Env& env, const ConvLayer& L, int M, int B) {
  if (L.prec == ALCM_PREC_FP32 || !env.k.splitk || L.nkb < 2) return 1;
  const long ctas = (long)conv_m_tiles(M) * L.n_tiles * B * L.nphase;
  if (ctas * 2 > (long)env.sms()) return 1;
  return (int)std::max<long>(1, std::min<long>(std::min<long>(L.nkb, env.sms() / ctas), 8));
}

static void conv_kernel_for(int prec, void (**kern)(ConvArgs), int* threads) {
  // one register budget for every tile width (128/thread, 2 CTAs of 192 threads per SM): a narrower, spilling
  // 80-register build for N < 128 (4 CTAs/SM) measured slower at every batch size (batch 1: 3.31 -> 3.24 ms)
  *threads = 192;
  *kern = is16(prec) ? conv_umma_kernel<0, 2> : conv_umma_kernel<1, 2>;
}

// One conv launch, fully decided at PLAN time (kernel, grid, shared memory, pipeline shape, cluster size): replaying
// the plan - eagerly or as a CUDA graph - can never disagree with the split-K / workspace decisions taken when it was built.
struct ConvLaunch {
  ConvArgs a;
  void (*kern)(ConvArgs) = nullptr;
  dim3 grid, block;
  uint32_t smem = 0;
  int cluster_x = 1;
  alcm_ctx* cx = nullptr;
  void run(cudaStream_t st) const {
    ConvArgs args = a;
    args.dbg = cx->conv_dbg;        // micro-benchmark instrumentation only (0 / null otherwise)
    args.trace = cx->conv_trace;
    cx->conv_last_grid = (int)grid.x;
    if (cluster_x > 1) t_cluster_x = cluster_x;
    launch_k(kern, grid, block, smem, st, args);
  }
};

static ConvLaunch plan_conv(const Env& env, const ConvLayer& L, const PlaneT& x, const PlaneT& out, const float* res, int M, float scale,
                            int accum, const SplitK& sk) {
  ConvLaunch cl;
  cl.cx = env.cx;
  ConvArgs& a = cl.a;
  memset(&a, 0, sizeof(a));
  a.x = x.p; a.xg = x.g;
  a.bias = L.bias;
  a.out = out.f(); a.og = out.g;
  a.res = res;
  a.M = M;
  a.ostride = L.nphase; a.nphase = L.nphase; a.ntaps = L.ntaps;
  memcpy(a.tap_off, L.tap_off, sizeof(a.tap_off));
  memcpy(a.min_off, L.min_off, sizeof(a.min_off));
  a.span = L.span;
  a.Cin = L.Cin; a.Cout = L.Cout;
  a.scale = scale; a.accum = accum;
  a.ksplit = 1;
  const int B = x.B;
  if (L.prec == ALCM_PREC_FP32) {
    a.w = reinterpret_cast<const uint8_t*>(L.weff);
    cl.kern = conv_simt_kernel;
    cl.grid = dim3((M + kSimtTM - 1) / kSimtTM, (L.Cout + kSimtTN - 1) / kSimtTN, B * L.nphase);
    cl.block = dim3(256);
    cl.smem = 0;
    return cl;
  }
  a.w = L.wpack;
  a.kchunks = L.kchunks; a.kblk = L.kblk; a.nkb = L.nkb;
  a.NT = L.NT; a.n_tiles = L.n_tiles; a.tmem_cols = L.tmem_cols;
  a.idesc = L.idesc; a.w_phase_stride = L.phase_stride; a.w_batch_stride = L.w_batch_stride;
  a.ksplit = sk.ksplit; a.ws = sk.ws; a.tile_ctr = sk.ctr; a.cluster_splitk = sk.cluster;
  // K chunks that exist in memory (plane_cpad): the others are an all-zero shared-memory slab
  REQUIRE(x.g.nchunk >= L.kchunks || L.nkb == 1, "conv: operand planes narrower than K need a single k-block");
  a.tiles_m = conv_m_tiles(M);
  a.tiles_total = a.tiles_m * L.n_tiles * B * L.nphase * sk.ksplit;
  a.acc_stages = 1;
  uint32_t smem = 0;
  pick_pipeline(env, L, (long)a.tiles_total, sk.ksplit, sk.cluster != 0, &a.w_stages, &a.tpg, &a.a_stages, &smem);
  int threads = 192;
  conv_kernel_for(L.prec, &cl.kern, &threads);
  // Persistent launch for multi-wave grids: one CTA per resident slot loops over tiles with two TMEM accumulators,
  // so barrier/TMEM setup is paid once per CTA and the epilogue of tile i overlaps the main loop of tile i+1.
  int grid = a.tiles_total;
  if (sk.ksplit == 1 && 2 * L.NT <= 512 && env.k.persist) {
    int tcols2 = 32;
    while (tcols2 < 2 * L.NT) tcols2 *= 2;
    // resident CTAs per SM: shared memory (1 KB reserved per CTA), registers (128/thread -> 2 CTAs of 192 threads),
    // TMEM columns.  Persistent only if doubling the TMEM columns does
    // not cost a resident CTA (N = 192 would drop from 2 to 1 CTA/SM) and there is more than one wave of tiles.
    int occ = (int)((227u * 1024u) / (smem + 1024u));
    occ = std::min(occ, 2);
    const int occ_np = std::min(occ, 512 / L.tmem_cols), occ_p = std::min(occ, 512 / tcols2);
    if (occ_p >= 1 && occ_p == occ_np && a.tiles_total > occ_p * env.sms()) {
      grid = occ_p * env.sms();
      a.acc_stages = 2;
      a.tmem_cols = tcols2;
    }
  }
  // Weights-stationary form for the narrow stages.  A 128-row tile of a C-channel conv streams C*C*k*2 bytes of weights
  // from L2 but only ~128*C*(2+4[+4]) bytes of activations: at C <= 96 the weights are 2-3x the activation traffic and the
  // launch is bound by L2 -> SM bandwidth (ncu launch list, profiles/r2_launches_bf16_b64.txt: same HBM bytes, k = 11
  // takes 20 % longer than k = 3).  When every tap of the (single) N tile fits beside the A ring, a persistent CTA loads
  // them once and keeps them for all its tiles.
  if (env.k.w_resident && a.acc_stages == 2 && L.n_tiles == 1 && L.nkb == 1 && L.nphase == 1 && L.w_batch_stride == 0) {
    const int groups = (L.ntaps + a.tpg - 1) / a.tpg;
    const uint32_t smem_r = conv_smem_layout(L.kblk, L.span, L.NT, groups, a.tpg, a.a_stages).total;
    if (groups <= 12 && smem_r + 1024u <= (227u * 1024u) / 2u) {
      a.w_stages = groups;
      a.w_resident = 1;
      smem = smem_r;
    }
  }
  cl.grid = dim3(grid);
  cl.block = dim3(threads);
  cl.smem = smem;
  cl.cluster_x = sk.cluster ? sk.ksplit : 1;
  return cl;
}

struct OpList {
  std::vector<Op> ops;
  Env env;              // device context + knob snapshot every planning decision below depends on
  Arena* ar = nullptr;  // where split-K workspaces come from (null: never split)
  int pdl = 0;          // launch this plan's kernels with programmatic dependent launch
  int cur_stage = -1;   // tag for the per-stage profile
  Arena* war = nullptr;          // model-owned arena + cache for re-tiled weights (null: keep the packed N tile)
  RetileCache* cache = nullptr;
  float* ws[kMaxLanes] = {nullptr, nullptr, nullptr, nullptr};
  size_t ws_bytes[kMaxLanes] = {0, 0, 0, 0};
  // conv (+bias, +residual, scale, accumulate): `out` fp32 planes, `x` operand planes
  void conv(const ConvLayer& L0, const PlaneT& x, const PlaneT& out, const PlaneT* res, float scale = 1.f, int accum = 0) {
    const int M = x.T;  // rows per batch item are input time steps (== output steps / nphase)
    int nt_c = 0, ks_c = 1;
    const bool dyn = L0.w_batch_stride != 0;   // per-item weights: packed for exactly this tiling, one K pass
    const bool clustered = !dyn && war && cache && ar && choose_cluster_tile(env, L0, M, x.B, &nt_c, &ks_c);
    const ConvLayer& L = (!dyn && war && cache) ? retile(*war, *cache, L0, clustered ? nt_c : pick_nt(env, L0, M, x.B)) : L0;
    REQUIRE(x.esz == opnd_esz(L.prec), "conv: operand dtype mismatch");
    REQUIRE(out.esz == 4 && out.T == x.T * L.nphase, "conv: bad output planes");
    REQUIRE(out.g.nchunk * 4 >= L.Cout, "conv: channel mismatch");
    REQUIRE(x.g.nchunk * (16 / x.esz) >= L.Cin, "conv: channel mismatch");
    Op op;
    op.cls = ALCM_CLS_CONV;
    op.flops = 2.0 * L.Cin * L.Cout * L.ntaps * L.nphase * (double)M * x.B;
    op.bytes = (double)x.B * x.T * ((double)L.Cin * x.esz + (double)L.Cout * L.nphase * 4 + (res ? (double)L.Cout * 4 : 0.0));
    const float* rp = res ? res->f() : nullptr;
    SplitK sk;
    if (clustered) { sk.ksplit = ks_c; sk.cluster = ks_c > 1; }
    else if (ar && !dyn) sk.ksplit = pick_ksplit(env, L, M, x.B);
    if (sk.ksplit > 1 && !sk.cluster) {
      const size_t tiles = (size_t)conv_m_tiles(M) * L.n_tiles * x.B * L.nphase;
      const size_t need = tiles * sk.ksplit * (size_t)L.NT * kTileM * 4;
      if (need > ws_bytes[cur_lane]) {  // ops of one lane run in stream order and can share the partial-tile workspace
        ws_bytes[cur_lane] = std::max<size_t>(need, (size_t)16 << 20);
        ws[cur_lane] = static_cast<float*>(ar->alloc(ws_bytes[cur_lane], false));
      }
      sk.ws = ws[cur_lane];
      sk.ctr = static_cast<unsigned int*>(ar->alloc(tiles * sizeof(unsigned int), true));
    }
    const ConvLaunch cl = plan_conv(env, L, x, out, rp, M, scale, accum, sk);
    op.fn = [cl](cudaStream_t st) { cl.run(st); };
    push(op);
  }
  void act(const PlaneT& x, const PlaneT& out, const float* ea, const float* ib, int round_tf32, bool fast) {
    REQUIRE(x.esz == 4 && x.T == out.T && x.B == out.B, "act: bad planes");
    ActArgs a;
    a.x = x.f(); a.xg = x.g; a.out = out.p; a.og = out.g; a.ea = ea; a.ib = ib; a.T = x.T; a.round_tf32 = round_tf32;
    const int B = x.B, T = x.T;
    const int nch = out.g.nchunk, oesz = out.esz;
    Op op;
    op.cls = ALCM_CLS_ACT;
    op.flops = 0;
    op.bytes = (double)B * T * x.C * (4.0 + oesz);  // algorithmic: unpadded channels, one read + one write
    // One kernel (act1d.cuh); block sizes for measurements (ALCM_ACT_VARIANT): 0 = 128 threads (635 outputs per block),
    // 1 = 64 threads (315) compiled for 14 blocks per SM (72 registers, no spills), 2 = 32 threads (155), 3 = 64 threads
    // at 96 registers / 10 blocks.  Default 1 - measured best at every launch size; against 3 both forms (15 KB of shared
    // memory per block) gain 3-5 % from the extra resident warps.
    int variant = 1;
    if (env.k.act_variant >= 0) variant = env.k.act_variant;
    op.fn = [=](cudaStream_t st) {
      const int npl = oesz == 4 ? 1 : 2;
      auto go = [&](auto threads_tag, auto minb_tag) {
        constexpr int TH = decltype(threads_tag)::value, MB = decltype(minb_tag)::value;
        using G = ActGeom<5, TH>;
        const dim3 grid((T + G::kTile - 1) / G::kTile, nch, B);
        const size_t sm = G::smem(npl);
        if (oesz == 4) {
          if (fast) launch_k(act1d_kernel<1, true, 5, TH, MB>, grid, dim3(TH), sm, st, a);
          else launch_k(act1d_kernel<1, false, 5, TH, MB>, grid, dim3(TH), sm, st, a);
        } else {
          launch_k(act1d_kernel<2, true, 5, TH, MB>, grid, dim3(TH), sm, st, a);
        }
      };
      if (variant == 0) go(std::integral_constant<int, 128>{}, std::integral_constant<int, 5>{});
      else if (variant == 2) go(std::integral_constant<int, 32>{}, std::integral_constant<int, 20>{});
      else if (variant == 3) go(std::integral_constant<int, 64>{}, std::integral_constant<int, 10>{});
      else go(std::integral_constant<int, 64>{}, std::integral_constant<int, 14>{});
    };
    push(op);
  }
  // fp32 planes -> operand planes (bf16 / tf32-rounded)
  void cast(const PlaneT& x, const PlaneT& out) {
    REQUIRE(x.esz == 4 && x.T == out.T, "cast: bad planes");
    const int B = x.B, T = x.T, nch = out.g.nchunk, oesz = out.esz;
    PlaneT xc = x, oc = out;
    Op op;
    op.cls = ALCM_CLS_MISC; op.flops = 0; op.bytes = (double)B * T * x.C * (4.0 + oesz);
    op.fn = [=](cudaStream_t st) {
      dim3 grid((T + 255) / 256, nch, B);
      if (oesz == 2) launch_k(cast_planes_kernel<8>, dim3(grid), dim3(256), 0, st, xc.f(), xc.g, oc.p, oc.g, T);
      else launch_k(cast_planes_kernel<4>, dim3(grid), dim3(256), 0, st, xc.f(), xc.g, oc.p, oc.g, T);
    };
    push(op);
  }
  // value/gate pair ops on fp32 planes of 2*inner channels -> operand planes of inner channels: GEGLU of the DiT feed-forward
  // (op 0), STFT magnitude of the mel front-end (op 1)
  void pair(const PlaneT& x, const PlaneT& out, int inner, int round_tf, int opk) {
    const int E = 16 / out.esz;
    REQUIRE(x.esz == 4 && x.T == out.T && x.B == out.B, "pair op: bad planes");
    REQUIRE(inner % E == 0 && x.g.nchunk * 4 >= 2 * inner && out.g.nchunk * E >= inner, "pair op: channel mismatch");
    const int B = x.B, T = x.T, oesz = out.esz;
    PlaneT xc = x, oc = out;
    Op op;
    op.cls = ALCM_CLS_MISC; op.flops = 0; op.bytes = (double)B * T * inner * (8.0 + oesz);
    op.fn = [=](cudaStream_t st) {
      dim3 grid((T + 127) / 128, inner / E, B);
      if (oesz == 2) launch_k(opk ? geglu_planes_kernel<8, 1> : geglu_planes_kernel<8, 0>, grid, dim3(128), 0, st, xc.f(), xc.g, oc.p, oc.g, T, inner, 0);
      else launch_k(opk ? geglu_planes_kernel<4, 1> : geglu_planes_kernel<4, 0>, grid, dim3(128), 0, st, xc.f(), xc.g, oc.p, oc.g, T, inner, round_tf);
    };
    push(op);
  }
  void geglu(const PlaneT& x, const PlaneT& out, int inner, int round_tf) { pair(x, out, inner, round_tf, 0); }
  // n-way sum of fp32 planes (block mean of the AMP blocks, models.py:190-196, when the blocks ran as
  // parallel lanes); writes fp32 planes and/or operand planes for the next conv
  void sum(const std::vector<PlaneT>& in, const PlaneT* out32, const PlaneT* out_op, int round_tf) {
    REQUIRE(in.size() >= 2 && in.size() <= 4 && (out32 || out_op), "sum: bad arguments");
    SumArgs a;
    memset(&a, 0, sizeof(a));
    a.n = (int)in.size();
    for (int i = 0; i < a.n; ++i) a.in[i] = in[i].f();
    a.g = in[0].g;
    a.T = in[0].T;
    a.out32 = out32 ? out32->f() : nullptr;
    a.out_op = out_op ? out_op->p : nullptr;
    a.og = out_op ? out_op->g : in[0].g;
    a.op_bf16 = out_op ? (out_op->esz == 2) : 0;
    a.round_tf32 = round_tf;
    const int B = in[0].B, T = in[0].T, nch = in[0].g.nchunk;
    Op op;
    op.cls = ALCM_CLS_MISC; op.flops = 0;
    op.bytes = (double)B * T * nch * 4 * (4.0 * a.n + (out32 ? 4 : 0) + (out_op ? out_op->esz : 0));
    op.fn = [=](cudaStream_t st) { launch_k(sum_planes_kernel, dim3(dim3((T + 255) / 256, nch / 2, B)), dim3(256), 0, st, a); };
    push(op);
  }
  void fork() { Op m; m.cls = OP_FORK; m.flops = m.bytes = 0; ops.push_back(m); }
  void join() { Op m; m.cls = OP_JOIN; m.flops = m.bytes = 0; cur_lane = 0; ops.push_back(m); }
  void lane(int l) { cur_lane = l; }
  void push(Op& op) { op.lane = cur_lane; op.stage = cur_stage; ops.push_back(op); }
  int cur_lane = 0;
  int launches() const {
    int n = 0;
    for (const Op& o : ops) n += o.cls >= 0;
    return n;
  }
  // serial execution on one stream (eager mode / profiling): lanes simply run one after another
  int pdl_eff() const { return env.k.pdl < 0 ? pdl : env.k.pdl; }
  void run(cudaStream_t st) const {
    t_pdl = pdl_eff();
    for (const Op& o : ops) if (o.cls >= 0) o.fn(st);
    t_pdl = 0;
  }
  // execution with fork/join across side streams (used under stream capture -> parallel graph branches)
  void run_lanes(cudaStream_t st, cudaStream_t* side, std::vector<cudaEvent_t>& evs) const {
    size_t ei = 0;
    auto next_ev = [&]() {
      if (ei == evs.size()) {
        cudaEvent_t e;
        CUDA_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        evs.push_back(e);
      }
      return evs[ei++];
    };
    t_pdl = pdl_eff();
    for (const Op& o : ops) {
      if (o.cls == OP_FORK) {
        cudaEvent_t e = next_ev();
        CUDA_CHECK(cudaEventRecord(e, st));
        for (int l = 1; l < kMaxLanes; ++l) CUDA_CHECK(cudaStreamWaitEvent(side[l - 1], e, 0));
      } else if (o.cls == OP_JOIN) {
        for (int l = 1; l < kMaxLanes; ++l) {
          cudaEvent_t e = next_ev();
          CUDA_CHECK(cudaEventRecord(e, side[l - 1]));
          CUDA_CHECK(cudaStreamWaitEvent(st, e, 0));
        }
      } else {
        o.fn(o.lane == 0 ? st : side[o.lane - 1]);
      }
    }
    t_pdl = 0;
  }
};

static void launch_pack(const float* in, const PlaneT& out, int C, int T, float mul, int prec, cudaStream_t st) {
  dim3 grid((T + 255) / 256, out.g.nchunk, out.B);
  if (out.esz == 2) launch_k(pack_cf_kernel<8>, dim3(grid), dim3(256), 0, st, in, out.p, out.g, C, T, mul, 0);
  else launch_k(pack_cf_kernel<4>, dim3(grid), dim3(256), 0, st, in, out.p, out.g, C, T, mul, prec == ALCM_PREC_TF32);
}
static void launch_unpack(const PlaneT& in, float* out, int C, int T, cudaStream_t st) {
  dim3 grid((T + 255) / 256, (C + 3) / 4, in.B);
  launch_k(unpack_cf_kernel, dim3(grid), dim3(256), 0, st, in.f(), in.g, out, C, T);
}

struct GraphExec {
  cudaGraphExec_t exec = nullptr;
  ~GraphExec() { if (exec) cudaGraphExecDestroy(exec); }
};

static void capture_graph(const OpList& ol, GraphExec& ge) {
  cudaStream_t cs, side[kMaxLanes - 1];
  CUDA_CHECK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  for (auto& sd : side) CUDA_CHECK(cudaStreamCreateWithFlags(&sd, cudaStreamNonBlocking));
  std::vector<cudaEvent_t> evs;
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal);
  if (e == cudaSuccess) {
    try {
      ol.run_lanes(cs, side, evs);
      e = cudaStreamEndCapture(cs, &graph);
    } catch (...) {
      cudaStreamEndCapture(cs, &graph);
      e = cudaErrorUnknown;
    }
  }
  if (e == cudaSuccess) e = cudaGraphInstantiate(&ge.exec, graph, 0);
  if (graph) cudaGraphDestroy(graph);
  for (auto ev : evs) cudaEventDestroy(ev);
  for (auto sd : side) cudaStreamDestroy(sd);
  cudaStreamDestroy(cs);
  if (e != cudaSuccess) {
    ge.exec = nullptr;
    throw AlcmError(ALCM_ERR_CUDA, std::string("graph capture failed: ") + cudaGetErrorString(e));
  }
}

// ------------------------------------------------------------------------------------------ plans
// A plan = everything one (B,T) shape needs: ONE device slab (activation planes, split-K workspaces and counters,
// zero-filled once - the plane halos are never written again), the kernel list with every launch geometry decided,
// and its CUDA graph.  Plans are built by alcm_*_plan() (or on the first call for a shape) and cached per model;
// alcm_*_workspace_bytes() runs only the sizing pass.  Once a shape is planned, alcm_vocode / alcm_vae_decode /
// alcm_decode_to_wav do no allocation and no synchronisation: they enqueue on the caller's stream and return.
//   * The slab is stream-ordered memory (cudaMallocAsync / cudaFreeAsync): building or retiring a plan never
//     synchronises the device.
//   * A model keeps at most ALCM_MAX_PLANS (default 16) plans.  The least recently used one is RETIRED: it leaves the
//     cache at once and is destroyed later, when the event recorded after its last launch has completed (checked with
//     cudaEventQuery on later calls) - never by waiting.
//   * The buffers of a plan are shared by all calls for that shape: a call on another stream than the previous one
//     first waits (on the device) for that event, so two streams cannot race on one plan.
//   * A plan that was launched while the caller's stream was being captured is pinned (the caller's graph holds its
//     addresses) and is never retired.
struct PlanBase {
  Arena ar;
  unsigned long long stamp = 0;  // last use (LRU)
  int B = 0, T = 0, Tout = 0;
  OpList ol;
  GraphExec ge;
  cudaEvent_t done = nullptr;    // recorded after the last launch of this plan
  cudaStream_t last_stream = nullptr;
  bool used = false, pinned = false;
  virtual ~PlanBase() { if (done) cudaEventDestroy(done); }
  bool idle() const { return !used || cudaEventQuery(done) == cudaSuccess; }
};

struct PlanCache {
  std::vector<std::unique_ptr<PlanBase>> retired;
  void reap(cudaStream_t st) {  // destroy retired plans whose last launch has finished (non-blocking)
    for (size_t i = 0; i < retired.size();) {
      if (retired[i]->idle()) {
        retired[i]->ar.release_async(st);
        retired.erase(retired.begin() + i);
      } else {
        ++i;
      }
    }
  }
  template <class Map>
  void make_room(Map& plans, size_t cap, cudaStream_t st) {
    reap(st);
    while (plans.size() >= cap) {
      auto victim = plans.end();
      for (auto it = plans.begin(); it != plans.end(); ++it)
        if (!it->second->pinned && (victim == plans.end() || it->second->stamp < victim->second->stamp)) victim = it;
      if (victim == plans.end()) break;  // everything is pinned by caller graphs
      retired.emplace_back(std::move(victim->second));
      plans.erase(victim);
    }
    reap(st);
  }
};

// Weight preparation, plan building and the single-op entry points run their set-up kernels on the legacy default
// stream and wait for THAT stream only - never cudaDeviceSynchronize(): a device-wide wait would also touch streams
// that another host thread is capturing into a CUDA graph (which fails both the wait and the other thread's capture).
static cudaError_t sync_setup() { return cudaStreamSynchronize(nullptr); }

// Model destruction: wait for the last launch of every plan (cached or retired) through the plans' own events.
template <class Map>
static void wait_plans(Map& plans, PlanCache& pc) {
  for (auto& kv : plans)
    if (kv.second->used) cudaEventSynchronize(kv.second->done);
  for (auto& r : pc.retired)
    if (r->used) cudaEventSynchronize(r->done);
  pc.retired.clear();
}

static bool stream_capturing(cudaStream_t st) {
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  return cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
}

// Replay a plan: its own CUDA graph normally; kernel by kernel (serial lanes) when the caller's stream is being captured -
// the launches then become nodes of the CALLER's graph (a graph launch inside a capture invalidated it on this driver).
static void run_plan(const PlanBase& P, cudaStream_t st) {
  if (P.ge.exec && !stream_capturing(st)) CUDA_CHECK(cudaGraphLaunch(P.ge.exec, st));
  else P.ol.run(st);
}

// Order a plan's launches after its previous use (other stream) and keep the retire event current.
struct PlanUse {
  PlanBase* P;
  cudaStream_t st;
  bool capturing = false;
  PlanUse(PlanBase* p, cudaStream_t s) : P(p), st(s) {
    capturing = stream_capturing(st);
    if (capturing) { P->pinned = true; return; }
    if (P->used && P->last_stream != st) CUDA_CHECK(cudaStreamWaitEvent(st, P->done, 0));
  }
  void finish() {
    if (capturing) return;
    CUDA_CHECK(cudaEventRecord(P->done, st));
    P->used = true;
    P->last_stream = st;
  }
};

// Sizing pass + build pass.  `build(plan)` must be deterministic in its allocation sequence.
template <class Plan, class Build>
static std::unique_ptr<Plan> build_plan(const Env& env, Arena* war, RetileCache* cache, int B, int T, cudaStream_t st, bool size_only,
                                        size_t* bytes_out, Build&& build) {
  auto setup = [&](Plan& P) {
    P.B = B; P.T = T;
    P.ol.env = env;
    P.ol.ar = &P.ar;
    P.ol.war = war; P.ol.cache = cache;
  };
  size_t need = 0;
  {
    Plan probe;
    probe.ar.guard = env.k.guard != 0;
    probe.ar.measuring = true;
    setup(probe);
    build(probe);
    need = probe.ar.off;
  }
  if (bytes_out) *bytes_out = need;
  if (size_only) return nullptr;
  std::unique_ptr<Plan> pl(new Plan());
  setup(*pl);
  pl->ar.guard = env.k.guard != 0;
  pl->ar.reserve(need, st);
  build(*pl);
  CUDA_CHECK(cudaEventCreateWithFlags(&pl->done, cudaEventDisableTiming));
  // the slab's zero-fill is ordered on `st`: a first use on another stream must wait for it (PlanUse)
  cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusNone) {
    CUDA_CHECK(cudaEventRecord(pl->done, st));
    pl->used = true;
    pl->last_stream = st;
  }
  if (env.k.graph) capture_graph(pl->ol, pl->ge);
  return pl;
}

struct SnakeP { float* ea; float* ib; };

struct AmpBlock {
  ConvLayer c1[3], c2[3];
  SnakeP a[6];
};
struct VocStage {
  ConvLayer up;
  std::vector<AmpBlock> blocks;
  int C, u;
};

struct VocPlan : PlanBase {
  PlaneT mel_in;   // operand planes
  PlaneT post_in;  // fp32 planes feeding conv_post
};

struct alcm_vocoder {
  alcm_ctx* ctx;
  Env env;
  alcm_bigvgan_cfg cfg;
  int prec;
  Arena war;  // weights
  ConvLayer conv_pre;
  std::vector<VocStage> stages;
  SnakeP act_post;
  float* post_w = nullptr;  // [7][plane_cpad(C)] tap-major fp32
  float post_bias = 0.f;
  int post_C = 0, hop = 1;
  RetileCache retiled;
  std::map<std::pair<int, int>, std::unique_ptr<VocPlan>> plans;
  PlanCache pcache;
};

static SnakeP make_snake(Arena& ar, const float* alpha, const float* beta, int C, int linear) {
  const int Cpad = round_up(C, 16);
  SnakeP s;
  s.ea = static_cast<float*>(ar.alloc((size_t)Cpad * 4, false));
  s.ib = static_cast<float*>(ar.alloc((size_t)Cpad * 4, false));
  snake_params_kernel<<<(Cpad + 127) / 128, 128>>>(alpha, beta, s.ea, s.ib, C, Cpad, linear);
  CUDA_CHECK(cudaGetLastError());
  return s;
}

// Kernel list of one vocode of shape (B,T): models.py:181-203.
static void voc_build(const alcm_vocoder* v, VocPlan& P) {
  const int B = P.B, T = P.T;
  const Knobs& K = v->env.k;
  const int prec = v->prec, oe = opnd_esz(prec);
  P.ar.f16 = (prec == ALCM_PREC_FP16);
  const int rtf = (prec == ALCM_PREC_TF32);
  const bool fast = (prec != ALCM_PREC_FP32);  // MUFU.COS snake; the exact-fp32 mode keeps the range-reduced sin
  const int nk = v->cfg.num_kernels;
  P.mel_in = make_planes(P.ar, B, v->cfg.num_mels, T, oe);
  int C = v->cfg.upsample_initial_channel, Tc = T;
  PlaneT xprev = make_planes(P.ar, B, C, Tc, 4);
  P.ol.cur_stage = 1;  // stage tags of the per-stage profile: 0 VAE, 1 conv_pre, 2..7 vocoder stages 1..6, 8 post
  P.ol.conv(v->conv_pre, P.mel_in, xprev, nullptr);
  PlaneT next_up_in;
  bool have_next_up_in = false;
  for (size_t i = 0; i < v->stages.size(); ++i) {
    const VocStage& S = v->stages[i];
    P.ol.cur_stage = 2 + (int)i;
    PlaneT up_in = xprev;
    if (have_next_up_in) {
      up_in = next_up_in;
      have_next_up_in = false;
    } else if (is16(prec)) {
      up_in = make_planes(P.ar, B, C, Tc, 2);
      P.ol.cast(xprev, up_in);
    }
    C = S.C; Tc *= S.u;
    PlaneT X = make_planes(P.ar, B, C, Tc, 4);
    P.ol.conv(S.up, up_in, X, nullptr);
    // The nk AMP blocks of a stage are independent chains.  Up to ~20 waves of conv CTAs per launch they run as
    // parallel graph lanes with private buffers (measured, tools/sweep_env.sh / sweep_lanes.sh: batch-1 decode 3.43 ms
    // with a 3-wave threshold, 3.30 ms with every stage in lanes; batch 4: 10.84 -> 10.05 ms); beyond that (batch 64)
    // they run back to back, share buffers and accumulate straight into XS.
    const long conv_ctas = (long)((Tc + kTileM - 1) / kTileM) * std::max(1, round_up(C, 16) / 128) * B;
    const bool parallel = nk > 1 && nk <= kMaxLanes && K.lanes && conv_ctas < (long)K.lane_waves * v->env.sms();
    PlaneT XS = make_planes(P.ar, B, C, Tc, 4);
    std::vector<PlaneT> Z;
    PlaneT R, Y, A;
    if (parallel) P.ol.fork();
    for (int j = 0; j < nk; ++j) {
      const AmpBlock& bk = S.blocks[j];
      if (parallel || j == 0) {
        R = make_planes(P.ar, B, C, Tc, 4); A = make_planes(P.ar, B, C, Tc, oe);
        Y = make_planes(P.ar, B, C, Tc, 4);
      }
      if (parallel) P.ol.lane(j);
      const PlaneT* cur = &X;
      const bool rb2 = v->cfg.resblock2 != 0;
      const int nl = rb2 ? 2 : 3;
      for (int l = 0; l < nl; ++l) {  // AMPBlock1 models.py:72-81: x += c2(a(c1(a(x)))); AMPBlock2 :119-126: x += c(a(x))
        const ConvLayer* tail;
        if (rb2) {
          P.ol.act(*cur, A, bk.a[l].ea, bk.a[l].ib, rtf, fast);
          tail = &bk.c1[l];
        } else {
          P.ol.act(*cur, A, bk.a[2 * l].ea, bk.a[2 * l].ib, rtf, fast);
          P.ol.conv(bk.c1[l], A, Y, nullptr);
          P.ol.act(Y, A, bk.a[2 * l + 1].ea, bk.a[2 * l + 1].ib, rtf, fast);
          tail = &bk.c2[l];
        }
        if (l < nl - 1) {
          P.ol.conv(*tail, A, R, cur);
          cur = &R;
        } else if (parallel) {  // block output in place of its residual stream (same thread reads then writes)
          P.ol.conv(*tail, A, R, cur, 1.0f / nk, 0);
          Z.push_back(R);
        } else {  // x = xs / num_kernels, models.py:190-196, folded into the last conv of each block
          P.ol.conv(*tail, A, XS, cur, 1.0f / nk, j > 0);
        }
      }
    }
    if (parallel) {
      P.ol.join();
      const bool last = (i + 1 == v->stages.size());
      if (is16(prec) && !last) {  // the only consumer is the next upsampler: emit its bf16 operand directly
        next_up_in = make_planes(P.ar, B, C, Tc, 2);
        P.ol.sum(Z, nullptr, &next_up_in, 0);
        have_next_up_in = true;
      } else {
        P.ol.sum(Z, &XS, nullptr, 0);
      }
    }
    xprev = XS;
  }
  P.ol.cur_stage = 2 + (int)v->stages.size();
  P.post_in = make_planes(P.ar, B, C, Tc, 4);
  P.ol.act(xprev, P.post_in, v->act_post.ea, v->act_post.ib, 0, fast);
  P.Tout = Tc;
}

static VocPlan* voc_plan(alcm_vocoder* v, int B, int T, cudaStream_t st) {
  auto key = std::make_pair(B, T);
  auto it = v->plans.find(key);
  if (it != v->plans.end()) { it->second->stamp = ++v->ctx->plan_clock; return it->second.get(); }
  v->pcache.make_room(v->plans, (size_t)v->env.k.max_plans, st);
  std::unique_ptr<VocPlan> pl = build_plan<VocPlan>(v->env, &v->war, &v->retiled, B, T, st, false, nullptr,
                                                    [&](VocPlan& P) { voc_build(v, P); });
  pl->stamp = ++v->ctx->plan_clock;
  VocPlan* raw = pl.get();
  v->plans[key] = std::move(pl);
  return raw;
}

// out: fp32 waveform [B, Tout] (pcm16 = 0) or 16-bit PCM [B, Tout] (pcm16 = 1, the WAV payload soundfile.write produces)
static void voc_run(alcm_vocoder* v, VocPlan* P, const float* mel, void* out, int pcm16, cudaStream_t st) {
  // mel either as [B,C,T] device tensor (packed here) or already resident in P->mel_in (decode_to_wav)
  if (mel) launch_pack(mel, P->mel_in, v->cfg.num_mels, P->T, 1.f, v->prec, st);
  run_plan(*P, st);
  const int threads = 256;
  dim3 grid((P->Tout + threads - 1) / threads, P->B);
  const size_t sm = (size_t)7 * P->post_in.g.nchunk * 4 * sizeof(float);
  if (pcm16) launch_k(conv_post_tanh_kernel<true>, dim3(grid), dim3(threads), sm, st, P->post_in.f(), P->post_in.g, v->post_w, v->post_bias, out, P->Tout, 7);
  else launch_k(conv_post_tanh_kernel<false>, dim3(grid), dim3(threads), sm, st, P->post_in.f(), P->post_in.g, v->post_w, v->post_bias, out, P->Tout, 7);
  CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------ VAE
struct GnP { float* gamma; float* beta; int C; };
struct ResBlock {
  GnP n1, n2;
  ConvLayer c1, c2, nin;
  bool has_nin = false;
  int Cin, Cout;
};
struct AttnBlk { GnP norm; ConvLayer qkv, proj; int C; };  // q, k, v 1x1 convs run as one launch (Cout = 3C)
struct VaeLevel { std::vector<ResBlock> blocks; std::vector<AttnBlk> attn; bool has_up = false; ConvLayer up; int C; };   // attn: one per block, or none

struct VaePlan : PlanBase {
  PlaneT z_in, mel_out;
};

struct alcm_vae {
  alcm_ctx* ctx;
  Env env;
  alcm_vae_cfg cfg;
  int prec;
  Arena war;
  ConvLayer post_quant, conv_in, conv_out;
  ResBlock mid1, mid2;
  AttnBlk attn;
  std::vector<VaeLevel> levels;  // in execution order (highest level first)
  GnP norm_out;
  int up_factor = 1;
  RetileCache retiled;
  std::map<std::pair<int, int>, std::unique_ptr<VaePlan>> plans;
  PlanCache pcache;
};

static void op_gn(OpList& ol, Arena& ar, const PlaneT& x, const PlaneT& out, const GnP& n, int swish, int prec) {
  const int B = x.B, C = x.C, T = x.T, groups = 32;
  REQUIRE(C % groups == 0, "GroupNorm: C must be a multiple of 32");
  PlaneT xc = x, oc = out;
  GnP nn = n;
  {  // one launch when a block's groups fit in shared memory (gn_fused_kernel)
    const int cpg = C / groups, E = 16 / out.esz;
    int gpb = 1;
    while (gpb <= 4 && (gpb * cpg) % E != 0) ++gpb;
    const size_t smem = (size_t)gpb * (cpg / 4) * T * 16;
    if ((cpg % 4) == 0 && gpb <= 4 && groups % gpb == 0 && smem <= 200 * 1024 && ol.env.k.gn_fused) {
      Op f;
      f.cls = ALCM_CLS_NORM; f.flops = 0; f.bytes = (double)B * C * T * (4.0 + out.esz);
      const int oesz = out.esz, rtf = (prec == ALCM_PREC_TF32);
      f.fn = [=](cudaStream_t st) {
        dim3 grid(groups / gpb, B);
        if (oesz == 2) launch_k(gn_fused_kernel<8>, grid, dim3(512), smem, st, xc.f(), xc.g, oc.p, oc.g, C, T, groups, 1e-6f, nn.gamma, nn.beta, swish, 0, gpb);
        else launch_k(gn_fused_kernel<4>, grid, dim3(512), smem, st, xc.f(), xc.g, oc.p, oc.g, C, T, groups, 1e-6f, nn.gamma, nn.beta, swish, rtf, gpb);
      };
      ol.ops.push_back(f);
      return;
    }
  }
  float2* stats = static_cast<float2*>(ar.alloc((size_t)B * groups * sizeof(float2), false));
  Op a;
  a.cls = ALCM_CLS_NORM; a.flops = 0; a.bytes = (double)B * C * T * 4;
  a.fn = [=](cudaStream_t st) { launch_k(gn_stats_kernel, dim3(dim3(groups, B)), dim3(512), 0, st, xc.f(), xc.g, C, T, groups, 1e-6f, stats); };
  ol.ops.push_back(a);
  Op b;
  b.cls = ALCM_CLS_NORM; b.flops = 0; b.bytes = (double)B * C * T * (4.0 + out.esz);
  const int oesz = out.esz, nch = out.g.nchunk, rtf = (prec == ALCM_PREC_TF32);
  b.fn = [=](cudaStream_t st) {
    dim3 grid((T + 127) / 128, nch, B);
    if (oesz == 2) launch_k(gn_apply_kernel<8>, dim3(grid), dim3(128), 0, st, xc.f(), xc.g, oc.p, oc.g, C, T, groups, stats, nn.gamma, nn.beta, swish, 0);
    else launch_k(gn_apply_kernel<4>, dim3(grid), dim3(128), 0, st, xc.f(), xc.g, oc.p, oc.g, C, T, groups, stats, nn.gamma, nn.beta, swish, rtf);
  };
  ol.ops.push_back(b);
}

// operand-dtype view of an fp32 tensor: bf16 needs a cast copy, tf32/fp32 read the fp32 planes directly
static PlaneT as_operand(OpList& ol, Arena& ar, const PlaneT& x, int prec) {
  if (!is16(prec)) return x;
  PlaneT o = make_planes(ar, x.B, x.C, x.T, 2);
  ol.cast(x, o);
  return o;
}

static PlaneT op_resblock(OpList& ol, Arena& ar, const ResBlock& rb, const PlaneT& x, int prec) {  // autoencoder1d.py:215-235
  const int oe = opnd_esz(prec);
  PlaneT a1 = make_planes(ar, x.B, rb.Cin, x.T, oe);
  op_gn(ol, ar, x, a1, rb.n1, 1, prec);
  PlaneT h = make_planes(ar, x.B, rb.Cout, x.T, 4);
  ol.conv(rb.c1, a1, h, nullptr);
  PlaneT a2 = make_planes(ar, x.B, rb.Cout, x.T, oe);
  op_gn(ol, ar, h, a2, rb.n2, 1, prec);
  PlaneT sc = x;
  if (rb.has_nin) {
    sc = make_planes(ar, x.B, rb.Cout, x.T, 4);
    ol.conv(rb.nin, as_operand(ol, ar, x, prec), sc, nullptr);
  }
  PlaneT out = make_planes(ar, x.B, rb.Cout, x.T, 4);
  ol.conv(rb.c2, a2, out, &sc);
  return out;
}

// scores (channel-split partials) -> softmax -> PV; q, k, v, h are fp32 planes, Sp/Pm workspaces
static void push_attention(OpList& ol, const PlaneT& q, const PlaneT& k, const PlaneT& v, const PlaneT& h, float* Sp, float* Pm,
                           int B, int C, int T) {
  const float scale = 1.0f / sqrtf((float)C);  // reference unpacks (b,c,t) as (b,t,c): scale = C^-0.5
  const int nsplit = (C >= 256) ? kAttnSplit : 1;
  Op s1;
  s1.cls = ALCM_CLS_ATTN; s1.flops = 2.0 * B * (double)T * T * C; s1.bytes = (double)B * (2.0 * C * T + (double)T * T) * 4;
  s1.fn = [=](cudaStream_t st) {
    launch_k(attn_scores2_kernel, dim3((T + 63) / 64, (T + 63) / 64, B * nsplit), dim3(256), 0, st, q.f(), k.f(), q.g, C, T, scale, Sp, nsplit, B);
  };
  ol.ops.push_back(s1);
  Op s2;
  s2.cls = ALCM_CLS_ATTN; s2.flops = 0; s2.bytes = (double)B * T * T * 4 * (nsplit + 1.0);
  s2.fn = [=](cudaStream_t st) { launch_k(softmax_rows2_kernel, dim3(B * T), dim3(128), (size_t)T * 4, st, Sp, nsplit, B * T, T, Pm); };
  ol.ops.push_back(s2);
  Op s3;
  s3.cls = ALCM_CLS_ATTN; s3.flops = 2.0 * B * (double)T * T * C; s3.bytes = (double)B * (2.0 * C * T + (double)T * T) * 4;
  s3.fn = [=](cudaStream_t st) {
    launch_k(attn_pv2_kernel, dim3((T + 63) / 64, ((C + 3) / 4 + 31) / 32, B), dim3(256), 0, st, v.f(), v.g, Pm, C, T, h.f(), h.g);
  };
  ol.ops.push_back(s3);
}

// channel sub-range [c0, c0+C) of a plane tensor as a view (same batch stride)
static PlaneT plane_view(const PlaneT& x, int c0, int C) {
  PlaneT v = x;
  v.C = C;
  v.p = x.p + (size_t)(c0 / (16 / x.esz)) * x.g.Tp * 16;
  return v;
}

// QK^T, softmax and PV of AttnBlock1D on the tensor cores: the scores are a 1x1 conv of q whose per-item "weights" are k,
// the output a 1x1 conv of the probabilities whose per-item weights are v (misc_kernels.cuh).  q: fp32 planes [C][T].
static void push_attention_tc(OpList& ol, Arena& ar, const PlaneT& q, const PlaneT& k, const PlaneT& v, const PlaneT& h, int B, int C, int T,
                              int prec) {
  const int oe = opnd_esz(prec), rtf = prec == ALCM_PREC_TF32;
  const float scale = 1.0f / sqrtf((float)C);  // reference unpacks (b,c,t) as (b,t,c): scale = C^-0.5
  ConvLayer Ls = dynamic_conv(ol.env.k, prec, /*Cout=keys*/ T, /*Cin=*/C);
  ConvLayer Lp = dynamic_conv(ol.env.k, prec, /*Cout=*/C, /*Cin=keys*/ T);
  Ls.wpack = static_cast<uint8_t*>(ar.alloc(Ls.phase_stride * B, false));
  Lp.wpack = static_cast<uint8_t*>(ar.alloc(Lp.phase_stride * B, false));
  PlaneT qo = make_planes(ar, B, C, T, oe);
  PlaneT S = make_planes(ar, B, T, T, 4);      // channels = keys, rows = queries
  PlaneT P = make_planes(ar, B, T, T, oe);
  auto pack = [&](const PlaneT& src, const ConvLayer& L, int mode) {
    Op op;
    op.cls = ALCM_CLS_ATTN; op.flops = 0; op.bytes = (double)B * C * T * (4.0 + oe);
    const size_t units = L.phase_stride / 16;
    const PlaneT sc = src;
    const ConvLayer Lc = L;
    op.fn = [=](cudaStream_t st) {
      const dim3 grid((unsigned)std::min<size_t>((units + 255) / 256, 2048), B);
      if (oe == 2) launch_k(pack_dyn_w_kernel<8>, grid, dim3(256), 0, st, sc.f(), sc.g, C, T, (void*)Lc.wpack, mode, Lc.NT, Lc.n_tiles, Lc.kblk, Lc.nkb, units, (int)(prec == ALCM_PREC_FP16));
      else launch_k(pack_dyn_w_kernel<4>, grid, dim3(256), 0, st, sc.f(), sc.g, C, T, (void*)Lc.wpack, mode, Lc.NT, Lc.n_tiles, Lc.kblk, Lc.nkb, units, 0);
    };
    ol.push(op);
  };
  ol.cast(q, qo);                     // q as operand planes (bf16 / tf32-rounded)
  ol.ops.back().cls = ALCM_CLS_ATTN;
  pack(k, Ls, 0);
  pack(v, Lp, 1);
  const size_t n0 = ol.ops.size();
  ol.conv(Ls, qo, S, nullptr, scale);
  {
    Op sm;
    sm.cls = ALCM_CLS_ATTN; sm.flops = 0; sm.bytes = (double)B * T * T * (3 * 4.0 + oe);
    const PlaneT Sc = S, Pc = P;
    sm.fn = [=](cudaStream_t st) {
      const dim3 grid((T + 127) / 128, B);
      if (oe == 2) launch_k(softmax_planes_kernel<8>, grid, dim3(128), 0, st, Sc.f(), Sc.g, (void*)Pc.p, Pc.g, T, 0);
      else launch_k(softmax_planes_kernel<4>, grid, dim3(128), 0, st, Sc.f(), Sc.g, (void*)Pc.p, Pc.g, T, rtf);
    };
    ol.push(sm);
  }
  ol.conv(Lp, P, h, nullptr);
  for (size_t i = n0; i < ol.ops.size(); ++i) ol.ops[i].cls = ALCM_CLS_ATTN;   // the two GEMMs count as attention, not conv
}

static PlaneT op_attn(OpList& ol, Arena& ar, const AttnBlk& at, const PlaneT& x, int prec) {  // autoencoder1d.py:257-278
  const int oe = opnd_esz(prec), B = x.B, C = at.C, T = x.T;
  PlaneT hn = make_planes(ar, B, C, T, oe);
  op_gn(ol, ar, x, hn, at.norm, 0, prec);
  PlaneT qkv = make_planes(ar, B, 3 * C, T, 4);
  ol.conv(at.qkv, hn, qkv, nullptr);
  const PlaneT q = plane_view(qkv, 0, C), k = plane_view(qkv, C, C), v = plane_view(qkv, 2 * C, C);
  PlaneT h = make_planes(ar, B, C, T, 4);
  if (prec != ALCM_PREC_FP32 && ol.env.k.attn_tc) {
    push_attention_tc(ol, ar, q, k, v, h, B, C, T, prec);
  } else {
    REQUIRE((size_t)T * 4 <= 200 * 1024, "attention: sequence too long for the row-softmax kernel");
    float* Sp = static_cast<float*>(ar.alloc((size_t)kAttnSplit * B * T * T * 4, false));
    float* Pm = static_cast<float*>(ar.alloc((size_t)B * T * T * 4, false));
    push_attention(ol, q, k, v, h, Sp, Pm, B, C, T);
  }
  PlaneT out = make_planes(ar, B, C, T, 4);
  ol.conv(at.proj, as_operand(ol, ar, h, prec), out, &x);
  return out;
}

// Kernel list of one decode of shape (B,T): autoencoder1d.py:59-62,484-517.
static void vae_build(const alcm_vae* v, VaePlan& P) {
  const int B = P.B, T = P.T;
  P.ol.cur_stage = 0;
  const int prec = v->prec, oe = opnd_esz(prec);
  P.ar.f16 = (prec == ALCM_PREC_FP16);
  P.z_in = make_planes(P.ar, B, v->cfg.embed_dim, T, oe);
  PlaneT h0 = make_planes(P.ar, B, v->cfg.z_channels, T, 4);
  P.ol.conv(v->post_quant, P.z_in, h0, nullptr);
  PlaneT h = make_planes(P.ar, B, v->conv_in.Cout, T, 4);
  P.ol.conv(v->conv_in, as_operand(P.ol, P.ar, h0, prec), h, nullptr);
  h = op_resblock(P.ol, P.ar, v->mid1, h, prec);
  h = op_attn(P.ol, P.ar, v->attn, h, prec);
  h = op_resblock(P.ol, P.ar, v->mid2, h, prec);
  for (const VaeLevel& lv : v->levels) {
    for (size_t i = 0; i < lv.blocks.size(); ++i) {   // autoencoder1d.py:500-504
      h = op_resblock(P.ol, P.ar, lv.blocks[i], h, prec);
      if (!lv.attn.empty()) h = op_attn(P.ol, P.ar, lv.attn[i], h, prec);
    }
    if (lv.has_up) {
      PlaneT up = make_planes(P.ar, B, lv.C, h.T * 2, 4);
      P.ol.conv(lv.up, as_operand(P.ol, P.ar, h, prec), up, nullptr);
      h = up;
    }
  }
  PlaneT a = make_planes(P.ar, B, h.C, h.T, oe);
  op_gn(P.ol, P.ar, h, a, v->norm_out, 1, prec);
  P.mel_out = make_planes(P.ar, B, v->cfg.out_ch, h.T, 4);
  P.ol.conv(v->conv_out, a, P.mel_out, nullptr);
  P.Tout = h.T;
}

static VaePlan* vae_plan(alcm_vae* v, int B, int T, cudaStream_t st) {
  auto key = std::make_pair(B, T);
  auto it = v->plans.find(key);
  if (it != v->plans.end()) { it->second->stamp = ++v->ctx->plan_clock; return it->second.get(); }
  v->pcache.make_room(v->plans, (size_t)v->env.k.max_plans, st);
  std::unique_ptr<VaePlan> pl = build_plan<VaePlan>(v->env, &v->war, &v->retiled, B, T, st, false, nullptr,
                                                    [&](VaePlan& P) { vae_build(v, P); });
  pl->stamp = ++v->ctx->plan_clock;
  VaePlan* raw = pl.get();
  v->plans[key] = std::move(pl);
  return raw;
}

static void vae_run(alcm_vae* v, VaePlan* P, const float* z, float inv_scale, cudaStream_t st) {
  launch_pack(z, P->z_in, v->cfg.embed_dim, P->T, inv_scale, v->prec, st);
  run_plan(*P, st);
  CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------ VAE encoder
// AutoencoderKL.encode up to the posterior's parameters (autoencoder1d.py:52-56): Encoder1D.forward (:391-413) + quant_conv.
// Same kernel families as the decoder; the only new op is Downsample1D (:296-316), run as a 2-tap conv on the
// time-folded input (s2d_cast_kernel).  The encoder's ResnetBlocks DO get kernel_size (k = 5), unlike the decoder's.
struct EncLevel { std::vector<ResBlock> blocks; std::vector<AttnBlk> attn; bool has_down = false; ConvLayer down; int C; };
struct EncPlan : PlanBase {
  PlaneT x_in, mom_out;
};
struct alcm_vae_encoder {
  alcm_ctx* ctx;
  Env env;
  alcm_vae_enc_cfg cfg;
  int prec;
  Arena war;
  ConvLayer conv_in, conv_out, quant;
  std::vector<EncLevel> levels;
  ResBlock mid1, mid2;
  AttnBlk attn;
  GnP norm_out;
  int down_factor = 1;
  RetileCache retiled;
  std::map<std::pair<int, int>, std::unique_ptr<EncPlan>> plans;
  PlanCache pcache;
};

static void enc_build(const alcm_vae_encoder* v, EncPlan& P) {
  const int B = P.B, T = P.T;
  P.ol.cur_stage = 0;
  const int prec = v->prec, oe = opnd_esz(prec), rtf = prec == ALCM_PREC_TF32;
  P.ar.f16 = (prec == ALCM_PREC_FP16);
  P.x_in = make_planes(P.ar, B, v->cfg.in_channels, T, oe);
  PlaneT h = make_planes(P.ar, B, v->conv_in.Cout, T, 4);
  P.ol.conv(v->conv_in, P.x_in, h, nullptr);
  for (const EncLevel& lv : v->levels) {
    for (size_t i = 0; i < lv.blocks.size(); ++i) {   // autoencoder1d.py:391-396
      h = op_resblock(P.ol, P.ar, lv.blocks[i], h, prec);
      if (!lv.attn.empty()) h = op_attn(P.ol, P.ar, lv.attn[i], h, prec);
    }
    if (lv.has_down) {
      REQUIRE(h.T % 2 == 0, "vae_encode: the sequence length must be even at every Downsample1D");
      const int To = h.T / 2, C = lv.C;
      PlaneT x2 = make_planes(P.ar, B, 2 * C, To, oe);
      {
        const PlaneT hc = h, xc = x2;
        Op op;
        op.cls = ALCM_CLS_MISC; op.flops = 0; op.bytes = (double)B * h.T * C * (4.0 + oe);
        op.fn = [=](cudaStream_t st) {
          dim3 grid((To + 255) / 256, xc.g.nchunk, B);
          if (oe == 2) launch_k(s2d_cast_kernel<8>, grid, dim3(256), 0, st, hc.f(), hc.g, (void*)xc.p, xc.g, C, To, 0);
          else launch_k(s2d_cast_kernel<4>, grid, dim3(256), 0, st, hc.f(), hc.g, (void*)xc.p, xc.g, C, To, rtf);
        };
        P.ol.push(op);
      }
      PlaneT d = make_planes(P.ar, B, C, To, 4);
      P.ol.conv(lv.down, x2, d, nullptr);
      h = d;
    }
  }
  h = op_resblock(P.ol, P.ar, v->mid1, h, prec);
  h = op_attn(P.ol, P.ar, v->attn, h, prec);
  h = op_resblock(P.ol, P.ar, v->mid2, h, prec);
  PlaneT a = make_planes(P.ar, B, h.C, h.T, oe);
  op_gn(P.ol, P.ar, h, a, v->norm_out, 1, prec);
  PlaneT m0 = make_planes(P.ar, B, v->conv_out.Cout, h.T, 4);
  P.ol.conv(v->conv_out, a, m0, nullptr);
  P.mom_out = make_planes(P.ar, B, v->quant.Cout, h.T, 4);
  P.ol.conv(v->quant, as_operand(P.ol, P.ar, m0, prec), P.mom_out, nullptr);
  P.Tout = h.T;
}

// ------------------------------------------------------------------------------------------ C-ABI
static void set_kernel_attrs() {
  const int mx = 227 * 1024;
  CUDA_CHECK(cudaFuncSetAttribute(softmax_rows2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CUDA_CHECK(cudaFuncSetAttribute(gn_fused_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CUDA_CHECK(cudaFuncSetAttribute(gn_fused_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CUDA_CHECK(cudaFuncSetAttribute(conv_umma_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
  CUDA_CHECK(cudaFuncSetAttribute(conv_umma_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, mx));
  // the activation blocks are small (15-21 KB, 64 threads): their occupancy is bounded by shared memory, so ask for the
  // largest shared-memory carve-out instead of leaving the split to the driver's heuristic
  auto carve = [](auto kern) { CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)); };
  carve(act1d_kernel<1, true, 5, 128, 5>); carve(act1d_kernel<1, false, 5, 128, 5>); carve(act1d_kernel<2, true, 5, 128, 5>);
  carve(act1d_kernel<1, true, 5, 64, 10>); carve(act1d_kernel<1, false, 5, 64, 10>); carve(act1d_kernel<2, true, 5, 64, 10>);
  carve(act1d_kernel<1, true, 5, 64, 14>); carve(act1d_kernel<1, false, 5, 64, 14>); carve(act1d_kernel<2, true, 5, 64, 14>);
  carve(act1d_kernel<1, true, 5, 32, 20>); carve(act1d_kernel<1, false, 5, 32, 20>); carve(act1d_kernel<2, true, 5, 32, 20>);
}

extern "C" {

const char* alcm_last_error(void) { return g_err.c_str(); }

int alcm_ctx_create(alcm_ctx** out, int device) {
  return guarded([&] {
    REQUIRE(out != nullptr, "ctx_create: out is NULL");
    int n = 0;
    CUDA_CHECK(cudaGetDeviceCount(&n));
    REQUIRE(device >= 0 && device < n, "ctx_create: no such CUDA device");
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    REQUIRE(prop.major == 10, "audiolcm_b200 requires an sm_100 (Blackwell B200) GPU; there is no fallback path");
    set_kernel_attrs();
    alcm_ctx* c = new alcm_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    *out = c;
  });
}
void alcm_ctx_destroy(alcm_ctx* ctx) { delete ctx; }

int alcm_vocoder_num_tensors(const alcm_bigvgan_cfg* c) {
  if (!c) return -1;
  const int per_block = c->resblock2 ? (2 * 3 + 2 * 2) : (18 + 12);   // AMPBlock2: 2 convs + 2 activations (models.py:90-126)
  return 3 + c->num_upsamples * (3 + c->num_kernels * per_block) + 2 + 3;
}

int alcm_vocoder_create(alcm_ctx* ctx, const alcm_bigvgan_cfg* cfg, const float* const* t, int n_tensors, int precision,
                        alcm_vocoder** out) {
  return guarded([&] {
    REQUIRE(ctx && cfg && t && out, "vocoder_create: NULL argument");
    REQUIRE(precision >= 0 && precision <= 3, "vocoder_create: bad precision");
    REQUIRE(cfg->num_upsamples >= 1 && cfg->num_upsamples <= 8 && cfg->num_kernels >= 1 && cfg->num_kernels <= 4,
            "vocoder_create: bad config");
    REQUIRE(n_tensors == alcm_vocoder_num_tensors(cfg), "vocoder_create: wrong tensor count");
    for (int i = 0; i < n_tensors; ++i) REQUIRE(t[i] != nullptr, "vocoder_create: NULL tensor");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    std::unique_ptr<alcm_vocoder> v(new alcm_vocoder());
    v->ctx = ctx; v->cfg = *cfg; v->prec = precision;
    v->env.cx = ctx; v->env.k = Knobs::from_env();  // the knob snapshot of this model
    v->war.guard = v->env.k.guard != 0;
    const Knobs& K = v->env.k;
    Arena tmp;
    int ti = 0;
    const int c0 = cfg->upsample_initial_channel;
    {
      const float* w = fold_wn(tmp, t[ti], t[ti + 1], c0, cfg->num_mels * 7);
      v->conv_pre = prepare_conv(v->war, K, precision, KIND_CONV, w, t[ti + 2], c0, cfg->num_mels, 7, 1);
      ti += 3;
    }
    int C = c0;
    v->hop = 1;
    for (int i = 0; i < cfg->num_upsamples; ++i) {
      const int u = cfg->upsample_rates[i], k = cfg->upsample_kernel_sizes[i];
      REQUIRE(C % 2 == 0, "vocoder_create: channel count must halve per stage");
      VocStage S;
      S.u = u; S.C = C / 2;
      v->hop *= u;
      {  // ConvTranspose1d weight (Cin,Cout,k): weight_norm dim 0 = input channel
        const float* w = fold_wn(tmp, t[ti], t[ti + 1], C, S.C * k);
        S.up = prepare_conv(v->war, K, precision, KIND_CONVT, w, t[ti + 2], S.C, C, k, u);
        ti += 3;
      }
      C = S.C;
      for (int j = 0; j < cfg->num_kernels; ++j) {
        AmpBlock bk;
        const int kk = cfg->resblock_kernel_sizes[j];
        const int nl = cfg->resblock2 ? 2 : 3;
        for (int l = 0; l < nl; ++l) {
          const float* w = fold_wn(tmp, t[ti], t[ti + 1], C, C * kk);
          bk.c1[l] = prepare_conv(v->war, K, precision, KIND_CONV, w, t[ti + 2], C, C, kk, cfg->resblock_dilation_sizes[j][l]);
          ti += 3;
        }
        for (int l = 0; l < 3 && !cfg->resblock2; ++l) {
          const float* w = fold_wn(tmp, t[ti], t[ti + 1], C, C * kk);
          bk.c2[l] = prepare_conv(v->war, K, precision, KIND_CONV, w, t[ti + 2], C, C, kk, 1);
          ti += 3;
        }
        for (int m = 0; m < (cfg->resblock2 ? 2 : 6); ++m) {
          bk.a[m] = make_snake(v->war, t[ti], t[ti + 1], C, cfg->snake_linear);
          ti += 2;
        }
        S.blocks.push_back(bk);
        tmp.release();
      }
      v->stages.push_back(std::move(S));
    }
    v->act_post = make_snake(v->war, t[ti], t[ti + 1], C, cfg->snake_linear);
    ti += 2;
    {  // conv_post (1,C,7) -> tap-major [7][Cpad] fp32 on device
      const float* w = fold_wn(tmp, t[ti], t[ti + 1], 1, C * 7);
      const int cp = plane_cpad(C);  // row pitch = channels held by the fp32 planes feeding conv_post
      std::vector<float> hw((size_t)C * 7), hp((size_t)7 * cp, 0.f);
      CUDA_CHECK(cudaMemcpy(hw.data(), w, hw.size() * 4, cudaMemcpyDeviceToHost));
      for (int c = 0; c < C; ++c)
        for (int j = 0; j < 7; ++j) hp[(size_t)j * cp + c] = hw[(size_t)c * 7 + j];
      v->post_w = static_cast<float*>(v->war.alloc(hp.size() * 4, false));
      CUDA_CHECK(cudaMemcpy(v->post_w, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice));
      CUDA_CHECK(cudaMemcpy(&v->post_bias, t[ti + 2], 4, cudaMemcpyDeviceToHost));
      v->post_C = C;
      ti += 3;
    }
    CUDA_CHECK(sync_setup());
    *out = v.release();
  });
}
void alcm_vocoder_destroy(alcm_vocoder* v) {
  if (!v) return;
  cudaSetDevice(v->ctx->device);
  wait_plans(v->plans, v->pcache);  // plans may still be in flight on caller streams
  delete v;
}

static void check_voc_shape(alcm_vocoder* v, int B, int T) {
  REQUIRE(B >= 1 && T >= 1, "vocode: B and T must be positive");
  REQUIRE((long long)T * v->hop < (1ll << 30), "vocode: clip too long for one call; shard it along time");
}

int alcm_vocoder_plan(alcm_vocoder* v, int B, int T, void* stream) {
  return guarded([&] {
    REQUIRE(v, "vocoder_plan: NULL argument");
    check_voc_shape(v, B, T);
    CUDA_CHECK(cudaSetDevice(v->ctx->device));
    voc_plan(v, B, T, static_cast<cudaStream_t>(stream));
  });
}

int alcm_vocoder_workspace_bytes(alcm_vocoder* v, int B, int T, size_t* bytes) {
  return guarded([&] {
    REQUIRE(v && bytes, "vocoder_workspace_bytes: NULL argument");
    check_voc_shape(v, B, T);
    CUDA_CHECK(cudaSetDevice(v->ctx->device));
    auto it = v->plans.find(std::make_pair(B, T));
    if (it != v->plans.end()) { *bytes = it->second->ar.total; return; }
    build_plan<VocPlan>(v->env, &v->war, &v->retiled, B, T, nullptr, true, bytes, [&](VocPlan& P) { voc_build(v, P); });
  });
}

static void vocode_any(alcm_vocoder* v, const float* mel, int B, int T, void* out, int pcm16, void* stream) {
  REQUIRE(v && mel && out, "vocode: NULL argument");
  check_voc_shape(v, B, T);
  CUDA_CHECK(cudaSetDevice(v->ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VocPlan* P = voc_plan(v, B, T, st);
  PlanUse use(P, st);
  voc_run(v, P, mel, out, pcm16, st);
  use.finish();
}

int alcm_vocode(alcm_vocoder* v, const float* mel, int B, int T, float* wav, void* stream) {
  return guarded([&] { vocode_any(v, mel, B, T, wav, 0, stream); });
}
int alcm_vocode_pcm16(alcm_vocoder* v, const float* mel, int B, int T, short* pcm, void* stream) {
  return guarded([&] { vocode_any(v, mel, B, T, pcm, 1, stream); });
}

// ---- VAE
static GnP make_gn(Arena& ar, const float* w, const float* b, int C) {
  GnP g;
  g.C = C;
  g.gamma = static_cast<float*>(ar.alloc((size_t)C * 4, false));
  g.beta = static_cast<float*>(ar.alloc((size_t)C * 4, false));
  CUDA_CHECK(cudaMemcpy(g.gamma, w, (size_t)C * 4, cudaMemcpyDeviceToDevice));
  CUDA_CHECK(cudaMemcpy(g.beta, b, (size_t)C * 4, cudaMemcpyDeviceToDevice));
  return g;
}

static int vae_resblock_tensors(int cin, int cout) { return 8 + (cin != cout ? 2 : 0); }

// AttnBlock1D parameters (autoencoder1d.py:241-254) from 10 tensors: norm (w,b), q, k, v, proj_out (w,b each).  The three
// [C,C,1] projections are concatenated into ONE conv with Cout = 3C.
static AttnBlk load_attn(Arena& war, const Knobs& K, int precision, const float* const* t, int& ti, int C) {
  AttnBlk a;
  a.C = C;
  a.norm = make_gn(war, t[ti], t[ti + 1], C);
  ti += 2;
  const size_t wn = (size_t)C * C, bn = (size_t)C;
  Arena tmpq;
  float* wcat = static_cast<float*>(tmpq.alloc(3 * wn * 4, false));
  float* bcat = static_cast<float*>(tmpq.alloc(3 * bn * 4, false));
  for (int i = 0; i < 3; ++i) {
    CUDA_CHECK(cudaMemcpy(wcat + i * wn, t[ti + 2 * i], wn * 4, cudaMemcpyDeviceToDevice));
    CUDA_CHECK(cudaMemcpy(bcat + i * bn, t[ti + 2 * i + 1], bn * 4, cudaMemcpyDeviceToDevice));
  }
  a.qkv = prepare_conv(war, K, precision, KIND_CONV, wcat, bcat, 3 * C, C, 1, 1);
  ti += 6;
  a.proj = prepare_conv(war, K, precision, KIND_CONV, t[ti], t[ti + 1], C, C, 1, 1);
  ti += 2;
  CUDA_CHECK(sync_setup());   // tmpq is freed on return
  return a;
}

int alcm_vae_num_tensors(const alcm_vae_cfg* c) {
  if (!c) return -1;
  int n = 2 + 2;
  int block_in = c->ch * c->ch_mult[c->n_levels - 1];
  n += vae_resblock_tensors(block_in, block_in) * 2 + 10;
  for (int lv = c->n_levels - 1; lv >= 0; --lv) {
    const int block_out = c->ch * c->ch_mult[lv];
    for (int i = 0; i <= c->num_res_blocks; ++i) {
      n += vae_resblock_tensors(block_in, block_out) + (c->attn_levels[lv] ? 10 : 0);
      block_in = block_out;
    }
    if (c->upsample_levels[lv]) n += 2;
  }
  return n + 2 + 2;
}

int alcm_vae_create(alcm_ctx* ctx, const alcm_vae_cfg* cfg, const float* const* t, int n_tensors, int precision, alcm_vae** out) {
  return guarded([&] {
    REQUIRE(ctx && cfg && t && out, "vae_create: NULL argument");
    REQUIRE(precision >= 0 && precision <= 3, "vae_create: bad precision");
    REQUIRE(cfg->n_levels >= 1 && cfg->n_levels <= 8, "vae_create: bad n_levels");
    REQUIRE(n_tensors == alcm_vae_num_tensors(cfg), "vae_create: wrong tensor count");
    for (int i = 0; i < n_tensors; ++i) REQUIRE(t[i] != nullptr, "vae_create: NULL tensor");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    std::unique_ptr<alcm_vae> v(new alcm_vae());
    v->ctx = ctx; v->cfg = *cfg; v->prec = precision;
    v->env.cx = ctx; v->env.k = Knobs::from_env();
    v->war.guard = v->env.k.guard != 0;
    const Knobs& K = v->env.k;
    int ti = 0;
    auto conv = [&](int cout, int cin, int k, ConvKind kind = KIND_CONV) {
      ConvLayer L = prepare_conv(v->war, K, precision, kind, t[ti], t[ti + 1], cout, cin, k, 1);
      ti += 2;
      return L;
    };
    auto gn = [&](int C) {
      GnP g = make_gn(v->war, t[ti], t[ti + 1], C);
      ti += 2;
      return g;
    };
    auto resblock = [&](int cin, int cout) {
      ResBlock rb;
      rb.Cin = cin; rb.Cout = cout;
      rb.n1 = gn(cin);
      rb.c1 = conv(cout, cin, 3);
      rb.n2 = gn(cout);
      rb.c2 = conv(cout, cout, 3);
      rb.has_nin = cin != cout;
      if (rb.has_nin) rb.nin = conv(cout, cin, 1);
      return rb;
    };
    v->post_quant = conv(cfg->z_channels, cfg->embed_dim, 1);
    int block_in = cfg->ch * cfg->ch_mult[cfg->n_levels - 1];
    v->conv_in = conv(block_in, cfg->z_channels, cfg->kernel_size);
    v->mid1 = resblock(block_in, block_in);
    v->attn = load_attn(v->war, K, precision, t, ti, block_in);
    v->mid2 = resblock(block_in, block_in);
    v->up_factor = 1;
    for (int lv = cfg->n_levels - 1; lv >= 0; --lv) {
      VaeLevel L;
      const int block_out = cfg->ch * cfg->ch_mult[lv];
      for (int i = 0; i <= cfg->num_res_blocks; ++i) {
        L.blocks.push_back(resblock(block_in, block_out));
        block_in = block_out;
        if (cfg->attn_levels[lv]) L.attn.push_back(load_attn(v->war, K, precision, t, ti, block_in));   // autoencoder1d.py:466-468
      }
      L.C = block_in;
      L.has_up = cfg->upsample_levels[lv] != 0;
      if (L.has_up) {
        L.up = conv(block_in, block_in, 3, KIND_UPCONV3);
        v->up_factor *= 2;
      }
      v->levels.push_back(std::move(L));
    }
    v->norm_out = gn(block_in);
    v->conv_out = conv(cfg->out_ch, block_in, cfg->kernel_size);
    REQUIRE(ti == n_tensors, "vae_create: tensor walk mismatch");
    CUDA_CHECK(sync_setup());
    *out = v.release();
  });
}
void alcm_vae_destroy(alcm_vae* v) {
  if (!v) return;
  cudaSetDevice(v->ctx->device);
  wait_plans(v->plans, v->pcache);
  delete v;
}

int alcm_vae_plan(alcm_vae* v, int B, int T, void* stream) {
  return guarded([&] {
    REQUIRE(v, "vae_plan: NULL argument");
    REQUIRE(B >= 1 && T >= 1, "vae_plan: B and T must be positive");
    CUDA_CHECK(cudaSetDevice(v->ctx->device));
    vae_plan(v, B, T, static_cast<cudaStream_t>(stream));
  });
}

int alcm_vae_workspace_bytes(alcm_vae* v, int B, int T, size_t* bytes) {
  return guarded([&] {
    REQUIRE(v && bytes, "vae_workspace_bytes: NULL argument");
    REQUIRE(B >= 1 && T >= 1, "vae_workspace_bytes: B and T must be positive");
    CUDA_CHECK(cudaSetDevice(v->ctx->device));
    auto it = v->plans.find(std::make_pair(B, T));
    if (it != v->plans.end()) { *bytes = it->second->ar.total; return; }
    build_plan<VaePlan>(v->env, &v->war, &v->retiled, B, T, nullptr, true, bytes, [&](VaePlan& P) { vae_build(v, P); });
  });
}

int alcm_vae_decode(alcm_vae* v, const float* z, int B, int T, float inv_scale, float* mel, void* stream) {
  return guarded([&] {
    REQUIRE(v && z && mel, "vae_decode: NULL argument");
    REQUIRE(B >= 1 && T >= 1, "vae_decode: B and T must be positive");
    CUDA_CHECK(cudaSetDevice(v->ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    VaePlan* P = vae_plan(v, B, T, st);
    PlanUse use(P, st);
    vae_run(v, P, z, inv_scale, st);
    launch_unpack(P->mel_out, mel, v->cfg.out_ch, P->Tout, st);
    CUDA_CHECK(cudaGetLastError());
    use.finish();
  });
}

// ---- VAE encoder
static int enc_resblock_tensors(int cin, int cout) { return 8 + (cin != cout ? 2 : 0); }
int alcm_vae_encoder_num_tensors(const alcm_vae_enc_cfg* c) {
  if (!c) return -1;
  int n = 2, block_in = c->ch;
  for (int lv = 0; lv < c->n_levels; ++lv) {
    const int block_out = c->ch * c->ch_mult[lv];
    for (int i = 0; i < c->num_res_blocks; ++i) { n += enc_resblock_tensors(block_in, block_out) + (c->attn_levels[lv] ? 10 : 0); block_in = block_out; }
    if (c->downsample_levels[lv]) n += 2;
  }
  return n + 2 * enc_resblock_tensors(block_in, block_in) + 10 + 2 + 2 + 2;
}

int alcm_vae_encoder_create(alcm_ctx* ctx, const alcm_vae_enc_cfg* cfg, const float* const* t, int n_tensors, int precision,
                            alcm_vae_encoder** out) {
  return guarded([&] {
    REQUIRE(ctx && cfg && t && out, "vae_encoder_create: NULL argument");
    REQUIRE(precision >= 0 && precision <= 3, "vae_encoder_create: bad precision");
    REQUIRE(cfg->n_levels >= 1 && cfg->n_levels <= 8 && cfg->num_res_blocks >= 1, "vae_encoder_create: bad config");
    REQUIRE(n_tensors == alcm_vae_encoder_num_tensors(cfg), "vae_encoder_create: wrong tensor count");
    for (int i = 0; i < n_tensors; ++i) REQUIRE(t[i] != nullptr, "vae_encoder_create: NULL tensor");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    std::unique_ptr<alcm_vae_encoder> v(new alcm_vae_encoder());
    v->ctx = ctx; v->cfg = *cfg; v->prec = precision;
    v->env.cx = ctx; v->env.k = Knobs::from_env();
    v->war.guard = v->env.k.guard != 0;
    const Knobs& K = v->env.k;
    const int ks = cfg->kernel_size;
    int ti = 0;
    auto conv = [&](int cout, int cin, int k, ConvKind kind = KIND_CONV, int p = 1) {
      ConvLayer L = prepare_conv(v->war, K, precision, kind, t[ti], t[ti + 1], cout, cin, k, p);
      ti += 2;
      return L;
    };
    auto gn = [&](int C) {
      GnP g = make_gn(v->war, t[ti], t[ti + 1], C);
      ti += 2;
      return g;
    };
    auto resblock = [&](int cin, int cout) {
      ResBlock rb;
      rb.Cin = cin; rb.Cout = cout;
      rb.n1 = gn(cin);
      rb.c1 = conv(cout, cin, ks);
      rb.n2 = gn(cout);
      rb.c2 = conv(cout, cout, ks);
      rb.has_nin = cin != cout;
      if (rb.has_nin) rb.nin = conv(cout, cin, 1);
      return rb;
    };
    v->conv_in = conv(cfg->ch, cfg->in_channels, ks);
    int block_in = cfg->ch;
    for (int lv = 0; lv < cfg->n_levels; ++lv) {
      EncLevel L;
      const int block_out = cfg->ch * cfg->ch_mult[lv];
      for (int i = 0; i < cfg->num_res_blocks; ++i) {
        L.blocks.push_back(resblock(block_in, block_out));
        block_in = block_out;
        if (cfg->attn_levels[lv]) L.attn.push_back(load_attn(v->war, K, precision, t, ti, block_in));   // autoencoder1d.py:356-358
      }
      L.C = block_in;
      L.has_down = cfg->downsample_levels[lv] != 0;
      if (L.has_down) {
        REQUIRE(block_in % 8 == 0, "vae_encoder_create: Downsample1D needs a multiple of 8 channels");
        L.down = conv(block_in, block_in, 3, KIND_DOWN2, 2);
        v->down_factor *= 2;
      }
      v->levels.push_back(std::move(L));
    }
    v->mid1 = resblock(block_in, block_in);
    v->attn = load_attn(v->war, K, precision, t, ti, block_in);
    v->mid2 = resblock(block_in, block_in);
    v->norm_out = gn(block_in);
    const int zc2 = (cfg->double_z ? 2 : 1) * cfg->z_channels;
    v->conv_out = conv(zc2, block_in, ks);
    v->quant = conv(2 * cfg->embed_dim, zc2, 1);
    REQUIRE(ti == n_tensors, "vae_encoder_create: tensor walk mismatch");
    CUDA_CHECK(sync_setup());
    *out = v.release();
  });
}
void alcm_vae_encoder_destroy(alcm_vae_encoder* v) {
  if (!v) return;
  cudaSetDevice(v->ctx->device);
  wait_plans(v->plans, v->pcache);
  delete v;
}
int alcm_vae_encode(alcm_vae_encoder* v, const float* x, int B, int T, float* moments, void* stream) {
  return guarded([&] {
    REQUIRE(v && x && moments, "vae_encode: NULL argument");
    REQUIRE(B >= 1 && T >= 1 && T % v->down_factor == 0, "vae_encode: T must be a positive multiple of the down-sampling factor");
    CUDA_CHECK(cudaSetDevice(v->ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto key = std::make_pair(B, T);
    auto it = v->plans.find(key);
    EncPlan* P = nullptr;
    if (it != v->plans.end()) {
      P = it->second.get();
    } else {
      v->pcache.make_room(v->plans, (size_t)v->env.k.max_plans, st);
      std::unique_ptr<EncPlan> pl = build_plan<EncPlan>(v->env, &v->war, &v->retiled, B, T, st, false, nullptr, [&](EncPlan& R) { enc_build(v, R); });
      P = pl.get();
      v->plans[key] = std::move(pl);
    }
    P->stamp = ++v->ctx->plan_clock;
    PlanUse use(P, st);
    launch_pack(x, P->x_in, v->cfg.in_channels, T, 1.f, v->prec, st);
    run_plan(*P, st);
    launch_unpack(P->mom_out, moments, v->quant.Cout, P->Tout, st);
    CUDA_CHECK(cudaGetLastError());
    use.finish();
  });
}

static void decode_any(alcm_vae* vae, alcm_vocoder* voc, const float* z, int B, int T, float inv_scale, float* mel_out, void* out,
                       int pcm16, void* stream) {
  REQUIRE(vae && voc && z && out, "decode_to_wav: NULL argument");
  REQUIRE(B >= 1 && T >= 1, "decode_to_wav: B and T must be positive");
  REQUIRE(vae->cfg.out_ch == voc->cfg.num_mels, "decode_to_wav: VAE out_ch != vocoder num_mels");
  REQUIRE(vae->ctx->device == voc->ctx->device, "decode_to_wav: handles live on different devices");
  CUDA_CHECK(cudaSetDevice(vae->ctx->device));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VaePlan* PV = vae_plan(vae, B, T, st);
  check_voc_shape(voc, B, PV->Tout);
  VocPlan* PW = voc_plan(voc, B, PV->Tout, st);
  PlanUse uv(PV, st), uw(PW, st);
  vae_run(vae, PV, z, inv_scale, st);
  if (mel_out) launch_unpack(PV->mel_out, mel_out, vae->cfg.out_ch, PV->Tout, st);
  // mel stays on the device in plane form: fp32 planes -> the vocoder's operand planes
  {
    const PlaneT& src = PV->mel_out;
    const PlaneT& dst = PW->mel_in;
    dim3 grid((src.T + 255) / 256, dst.g.nchunk, B);
    if (dst.esz == 2) launch_k(cast_planes_kernel<8>, dim3(grid), dim3(256), 0, st, src.f(), src.g, dst.p, dst.g, src.T);
    else if (voc->prec == ALCM_PREC_TF32) launch_k(cast_planes_kernel<4>, dim3(grid), dim3(256), 0, st, src.f(), src.g, dst.p, dst.g, src.T);
    else CUDA_CHECK(cudaMemcpyAsync(dst.p, src.p, src.bytes, cudaMemcpyDeviceToDevice, st));
  }
  voc_run(voc, PW, nullptr, out, pcm16, st);
  uv.finish();
  uw.finish();
}

int alcm_decode_to_wav(alcm_vae* vae, alcm_vocoder* voc, const float* z, int B, int T, float inv_scale, float* mel_out,
                       float* wav, void* stream) {
  return guarded([&] { decode_any(vae, voc, z, B, T, inv_scale, mel_out, wav, 0, stream); });
}
int alcm_decode_to_pcm16(alcm_vae* vae, alcm_vocoder* voc, const float* z, int B, int T, float inv_scale, float* mel_out,
                         short* pcm, void* stream) {
  return guarded([&] { decode_any(vae, voc, z, B, T, inv_scale, mel_out, pcm, 1, stream); });
}

// ---- LCM sampler step (the scalar half of SURVEY 8f row 2; the DiT denoiser stays reference PyTorch) ----------
int alcm_lcm_step(alcm_ctx* ctx, const float* sample, const float* eps, const float* noise, float* prev, float* denoised, long long n,
                  float sqrt_alpha_prod_t, float sqrt_beta_prod_t, float c_out, float c_skip, float sqrt_alpha_prod_prev,
                  float sqrt_beta_prod_prev, int last_step, void* stream) {
  return guarded([&] {
    REQUIRE(ctx && sample && eps && prev && denoised, "lcm_step: NULL argument");
    REQUIRE(last_step || noise, "lcm_step: noise is required on every step but the last");
    REQUIRE(n > 0 && (n % 4) == 0, "lcm_step: element count must be a positive multiple of 4");
    REQUIRE(sqrt_alpha_prod_t > 0.f, "lcm_step: sqrt(alpha_prod_t) must be positive");
    REQUIRE(((uintptr_t)sample | (uintptr_t)eps | (uintptr_t)(noise ? noise : sample) | (uintptr_t)prev | (uintptr_t)denoised) % 16 == 0,
            "lcm_step: tensors must be 16-byte aligned");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    LcmStepCoef k{sqrt_beta_prod_t, 1.0f / sqrt_alpha_prod_t, c_out, c_skip, sqrt_alpha_prod_prev, sqrt_beta_prod_prev, last_step};
    const size_t n4 = (size_t)n / 4;
    const unsigned blocks = (unsigned)std::min<size_t>((n4 + 255) / 256, (size_t)ctx->sm_count * 8);
    launch_k(lcm_step_kernel, dim3(blocks), dim3(256), 0, static_cast<cudaStream_t>(stream), reinterpret_cast<const float4*>(sample),
             reinterpret_cast<const float4*>(eps), reinterpret_cast<const float4*>(noise), reinterpret_cast<float4*>(prev),
             reinterpret_cast<float4*>(denoised), n4, k);
    CUDA_CHECK(cudaGetLastError());
  });
}

// nn.LayerNorm(C) of a channels-first tensor (the DiT token stream in conv layout), see layernorm_cf_kernel
int alcm_layernorm_cf(alcm_ctx* ctx, const float* x, const float* gamma, const float* beta, float* y, int B, int C, int T, float eps,
                      void* stream) {
  return guarded([&] {
    REQUIRE(ctx && x && gamma && beta && y, "layernorm_cf: NULL argument");
    REQUIRE(B >= 1 && C >= 1 && T >= 1 && B <= 65535, "layernorm_cf: bad shape");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    launch_k(layernorm_cf_kernel, dim3((T + 127) / 128, B), dim3(128), 0, static_cast<cudaStream_t>(stream), x, gamma, beta, y, C, T, eps);
    CUDA_CHECK(cudaGetLastError());
  });
}

// ---- stand-alone Conv1d layer handle (the 9-tap Conv1dFeedForward convs of the DiT denoiser,
//      ldm/modules/new_attention.py:48-74, are 93 % of its FLOPs and have exactly the shape conv_umma_kernel handles) ----
struct ConvRunPlan : PlanBase {
  PlaneT x_in, res_in, out;
};
struct alcm_conv1d {
  alcm_ctx* ctx;
  Env env;
  int prec, Cin, Cout;
  Arena war;
  ConvLayer L;
  RetileCache retiled;
  std::map<std::tuple<int, int, int>, std::unique_ptr<ConvRunPlan>> plans;   // (B, T, has_res)
  PlanCache pcache;
};

int alcm_conv1d_create(alcm_ctx* ctx, const float* w, const float* bias, int Cout, int Cin, int K, int dilation, int precision,
                       alcm_conv1d** out) {
  return guarded([&] {
    REQUIRE(ctx && w && out, "conv1d_create: NULL argument");
    REQUIRE(Cout >= 1 && Cin >= 1 && dilation >= 1, "conv1d_create: bad shape");
    REQUIRE(precision >= 0 && precision <= 3, "conv1d_create: bad precision");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    std::unique_ptr<alcm_conv1d> c(new alcm_conv1d());
    c->ctx = ctx; c->prec = precision; c->Cin = Cin; c->Cout = Cout;
    c->env.cx = ctx; c->env.k = Knobs::from_env();
    c->war.guard = c->env.k.guard != 0;
    c->L = prepare_conv(c->war, c->env.k, precision, KIND_CONV, w, bias, Cout, Cin, K, dilation);
    *out = c.release();
  });
}
void alcm_conv1d_destroy(alcm_conv1d* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  wait_plans(c->plans, c->pcache);
  delete c;
}
int alcm_conv1d_run(alcm_conv1d* c, const float* x, const float* res, float* y, int B, int T, void* stream) {
  return guarded([&] {
    REQUIRE(c && x && y, "conv1d_run: NULL argument");
    REQUIRE(B >= 1 && T >= 1, "conv1d_run: B and T must be positive");
    CUDA_CHECK(cudaSetDevice(c->ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto key = std::make_tuple(B, T, res ? 1 : 0);
    auto it = c->plans.find(key);
    ConvRunPlan* P = nullptr;
    if (it != c->plans.end()) {
      P = it->second.get();
    } else {
      c->pcache.make_room(c->plans, (size_t)c->env.k.max_plans, st);
      const bool has_res = res != nullptr;
      std::unique_ptr<ConvRunPlan> pl = build_plan<ConvRunPlan>(c->env, &c->war, &c->retiled, B, T, st, false, nullptr, [&](ConvRunPlan& R) {
        R.ar.f16 = (c->prec == ALCM_PREC_FP16);
        R.x_in = make_planes(R.ar, B, c->Cin, T, opnd_esz(c->prec));
        R.out = make_planes(R.ar, B, c->Cout, T, 4);
        if (has_res) R.res_in = make_planes(R.ar, B, c->Cout, T, 4);
        R.ol.conv(c->L, R.x_in, R.out, has_res ? &R.res_in : nullptr);
        R.Tout = T;
      });
      P = pl.get();
      c->plans[key] = std::move(pl);
    }
    P->stamp = ++c->ctx->plan_clock;
    PlanUse use(P, st);
    launch_pack(x, P->x_in, c->Cin, T, 1.f, c->prec, st);
    if (res) launch_pack(res, P->res_in, c->Cout, T, 1.f, ALCM_PREC_FP32, st);
    run_plan(*P, st);
    launch_unpack(P->out, y, c->Cout, T, st);
    CUDA_CHECK(cudaGetLastError());
    use.finish();
  });
}

// ---- the DiT's Conv1dFeedForward (glu=True) as ONE plan: Conv1d(dim -> 2*inner, k) -> x * gelu(gate) -> Conv1d(inner -> dim_out, k)
//      (+ residual), ldm/modules/new_attention.py:48-74; the 2*inner-channel intermediate stays in planes ----
struct alcm_ffn1d {
  alcm_ctx* ctx;
  Env env;
  int prec, dim, inner, dim_out;
  Arena war;
  ConvLayer L1, L2;
  RetileCache retiled;
  std::map<std::tuple<int, int, int>, std::unique_ptr<ConvRunPlan>> plans;   // (B, T, has_res)
  PlanCache pcache;
};

int alcm_ffn1d_create(alcm_ctx* ctx, const float* w_in, const float* b_in, const float* w_out, const float* b_out, int dim, int inner,
                      int dim_out, int K, int precision, alcm_ffn1d** out) {
  return guarded([&] {
    REQUIRE(ctx && w_in && w_out && out, "ffn1d_create: NULL argument");
    REQUIRE(dim >= 1 && inner >= 8 && inner % 8 == 0 && dim_out >= 1 && K >= 1 && (K & 1), "ffn1d_create: bad shape (inner % 8 == 0, odd K)");
    REQUIRE(precision >= 0 && precision <= 3, "ffn1d_create: bad precision");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    std::unique_ptr<alcm_ffn1d> c(new alcm_ffn1d());
    c->ctx = ctx; c->prec = precision; c->dim = dim; c->inner = inner; c->dim_out = dim_out;
    c->env.cx = ctx; c->env.k = Knobs::from_env();
    c->war.guard = c->env.k.guard != 0;
    c->L1 = prepare_conv(c->war, c->env.k, precision, KIND_CONV, w_in, b_in, 2 * inner, dim, K, 1);
    c->L2 = prepare_conv(c->war, c->env.k, precision, KIND_CONV, w_out, b_out, dim_out, inner, K, 1);
    *out = c.release();
  });
}
void alcm_ffn1d_destroy(alcm_ffn1d* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  wait_plans(c->plans, c->pcache);
  delete c;
}
int alcm_ffn1d_run(alcm_ffn1d* c, const float* x, const float* res, float* y, int B, int T, void* stream) {
  return guarded([&] {
    REQUIRE(c && x && y, "ffn1d_run: NULL argument");
    REQUIRE(B >= 1 && T >= 1, "ffn1d_run: B and T must be positive");
    CUDA_CHECK(cudaSetDevice(c->ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto key = std::make_tuple(B, T, res ? 1 : 0);
    auto it = c->plans.find(key);
    ConvRunPlan* P = nullptr;
    if (it != c->plans.end()) {
      P = it->second.get();
    } else {
      c->pcache.make_room(c->plans, (size_t)c->env.k.max_plans, st);
      const bool has_res = res != nullptr;
      std::unique_ptr<ConvRunPlan> pl = build_plan<ConvRunPlan>(c->env, &c->war, &c->retiled, B, T, st, false, nullptr, [&](ConvRunPlan& R) {
        R.ar.f16 = (c->prec == ALCM_PREC_FP16);
        R.x_in = make_planes(R.ar, B, c->dim, T, opnd_esz(c->prec));
        PlaneT mid = make_planes(R.ar, B, 2 * c->inner, T, 4);
        PlaneT act = make_planes(R.ar, B, c->inner, T, opnd_esz(c->prec));
        R.out = make_planes(R.ar, B, c->dim_out, T, 4);
        if (has_res) R.res_in = make_planes(R.ar, B, c->dim_out, T, 4);
        R.ol.conv(c->L1, R.x_in, mid, nullptr);
        R.ol.geglu(mid, act, c->inner, c->prec == ALCM_PREC_TF32);
        R.ol.conv(c->L2, act, R.out, has_res ? &R.res_in : nullptr);
        R.Tout = T;
      });
      P = pl.get();
      c->plans[key] = std::move(pl);
    }
    P->stamp = ++c->ctx->plan_clock;
    PlanUse use(P, st);
    launch_pack(x, P->x_in, c->dim, T, 1.f, c->prec, st);
    if (res) launch_pack(res, P->res_in, c->dim_out, T, 1.f, ALCM_PREC_FP32, st);
    run_plan(*P, st);
    launch_unpack(P->out, y, c->dim_out, T, st);
    CUDA_CHECK(cudaGetLastError());
    use.finish();
  });
}

// ---- log10-mel front-end (ldm/data/preprocess/NAT_mel.py:64-85) as one plan: clamp + reflect pad + fold -> STFT as a
//      (taps+1)-tap conv -> magnitude -> mel filterbank as a 1x1 conv -> log10(clamp); SURVEY 8f row 4 ----
struct MelPlan : PlanBase {
  PlaneT x_in, mel_out;
};
struct alcm_melspec {
  alcm_ctx* ctx;
  Env env;
  int prec, hop, taps, nb_pad, n_mels;
  Arena war;
  ConvLayer stft, mel;
  RetileCache retiled;
  std::map<std::pair<int, int>, std::unique_ptr<MelPlan>> plans;   // (B, L)
  PlanCache pcache;
};

int alcm_melspec_create(alcm_ctx* ctx, const float* stft_w, const float* mel_w, int hop, int taps, int nb_pad, int n_mels, int precision,
                        alcm_melspec** out) {
  return guarded([&] {
    REQUIRE(ctx && stft_w && mel_w && out, "melspec_create: NULL argument");
    REQUIRE(hop >= 8 && taps >= 1 && taps <= 10 && (taps & 1) == 0 && nb_pad >= 8 && nb_pad % 8 == 0 && n_mels >= 1,
            "melspec_create: bad shape (even taps <= 10, nb_pad % 8 == 0)");
    REQUIRE(precision >= 0 && precision <= 3, "melspec_create: bad precision");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    std::unique_ptr<alcm_melspec> c(new alcm_melspec());
    c->ctx = ctx; c->prec = precision; c->hop = hop; c->taps = taps; c->nb_pad = nb_pad; c->n_mels = n_mels;
    c->env.cx = ctx; c->env.k = Knobs::from_env();
    c->war.guard = c->env.k.guard != 0;
    c->stft = prepare_conv(c->war, c->env.k, precision, KIND_CONV, stft_w, nullptr, 2 * nb_pad, hop, taps + 1, 1);
    c->mel = prepare_conv(c->war, c->env.k, precision, KIND_CONV, mel_w, nullptr, n_mels, nb_pad, 1, 1);
    *out = c.release();
  });
}
void alcm_melspec_destroy(alcm_melspec* c) {
  if (!c) return;
  cudaSetDevice(c->ctx->device);
  wait_plans(c->plans, c->pcache);
  delete c;
}
int alcm_melspec_run(alcm_melspec* c, const float* y, float* mel, int B, int L, void* stream) {
  return guarded([&] {
    REQUIRE(c && y && mel, "melspec_run: NULL argument");
    REQUIRE(B >= 1 && L >= c->hop * c->taps && L % c->hop == 0, "melspec_run: L must be a positive multiple of hop (>= one window)");
    CUDA_CHECK(cudaSetDevice(c->ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int frames = L / c->hop, rows = frames + c->taps - 1, padw = (c->taps - 1) * c->hop / 2;
    auto key = std::make_pair(B, L);
    auto it = c->plans.find(key);
    MelPlan* P = nullptr;
    if (it != c->plans.end()) {
      P = it->second.get();
    } else {
      c->pcache.make_room(c->plans, (size_t)c->env.k.max_plans, st);
      std::unique_ptr<MelPlan> pl = build_plan<MelPlan>(c->env, &c->war, &c->retiled, B, rows, st, false, nullptr, [&](MelPlan& R) {
        R.ar.f16 = (c->prec == ALCM_PREC_FP16);
        R.x_in = make_planes(R.ar, B, c->hop, rows, opnd_esz(c->prec));
        PlaneT spec = make_planes(R.ar, B, 2 * c->nb_pad, rows, 4);
        PlaneT mag = make_planes(R.ar, B, c->nb_pad, rows, opnd_esz(c->prec));
        R.mel_out = make_planes(R.ar, B, c->n_mels, rows, 4);
        R.ol.conv(c->stft, R.x_in, spec, nullptr);
        R.ol.pair(spec, mag, c->nb_pad, c->prec == ALCM_PREC_TF32, 1);
        R.ol.conv(c->mel, mag, R.mel_out, nullptr);
        R.Tout = frames;
      });
      P = pl.get();
      c->plans[key] = std::move(pl);
    }
    P->stamp = ++c->ctx->plan_clock;
    PlanUse use(P, st);
    {
      const PlaneT& X = P->x_in;
      dim3 grid((rows + 127) / 128, X.g.nchunk, B);
      if (X.esz == 2) launch_k(mel_fold_kernel<8>, grid, dim3(128), 0, st, y, L, c->hop, padw, (void*)X.p, X.g, rows, 0);
      else launch_k(mel_fold_kernel<4>, grid, dim3(128), 0, st, y, L, c->hop, padw, (void*)X.p, X.g, rows, (int)(c->prec == ALCM_PREC_TF32));
    }
    run_plan(*P, st);
    {  // frame m = sum_i X[m+i] D[i] is row m + taps/2 of the 'same' (taps+1)-tap conv whose first tap is zero
      const PlaneT& M = P->mel_out;
      dim3 grid((frames + 255) / 256, (c->n_mels + 3) / 4, B);
      launch_k(unpack_log10_kernel, grid, dim3(256), 0, st, M.f(), M.g, mel, c->n_mels, frames, c->taps / 2 - 1, 1e-5f);
    }
    CUDA_CHECK(cudaGetLastError());
    use.finish();
  });
}

// ---- single-op entry points ---------------------------------------------------------------
static void sync_free(const Arena& ar, cudaStream_t st) {
  CUDA_CHECK(cudaStreamSynchronize(st));
  if (ar.guard) REQUIRE(ar.guard_violations() == 0, "ALCM_GUARD: a kernel wrote outside its buffers (guard zone modified)");
}

int alcm_activation1d_fwd(alcm_ctx* ctx, const float* x, const float* alpha, const float* beta, float* y, int B, int C, int T,
                          int precision, void* stream) {
  return guarded([&] {
    REQUIRE(ctx && x && alpha && beta && y, "activation1d: NULL argument");
    REQUIRE(B >= 1 && C >= 1 && T >= 1, "activation1d: empty tensor");
    REQUIRE(precision >= 0 && precision <= 3, "activation1d: bad precision");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Arena ar;
    ar.guard = env_int("ALCM_GUARD", 0) != 0;
    ar.f16 = (precision == ALCM_PREC_FP16);
    PlaneT xin = make_planes(ar, B, C, T, 4), out = make_planes(ar, B, C, T, opnd_esz(precision));
    SnakeP sp = make_snake(ar, alpha, beta, C, 0);
    CUDA_CHECK(sync_setup());
    launch_pack(x, xin, C, T, 1.f, ALCM_PREC_FP32, st);
    OpList ol;
    ol.env = Env{ctx, Knobs::from_env()};
    ol.act(xin, out, sp.ea, sp.ib, precision == ALCM_PREC_TF32, precision != ALCM_PREC_FP32);
    ol.run(st);
    if (is16(precision)) {
      dim3 grid((T + 255) / 256, out.g.nchunk, B);
      launch_k(unpack_cf_bf16_kernel, dim3(grid), dim3(256), 0, st, out.p, out.g, y, C, T);
    } else {
      launch_unpack(out, y, C, T, st);
    }
    CUDA_CHECK(cudaGetLastError());
    sync_free(ar, st);
  });
}

static void run_conv_test(alcm_ctx* ctx, ConvKind kind, const float* x, const float* w, const float* bias, const float* res,
                          float* y, int B, int Cin, int Cout, int T, int K, int p, int precision, cudaStream_t st) {
  REQUIRE(ctx && x && w && y, "conv: NULL argument");
  REQUIRE(B >= 1 && Cin >= 1 && Cout >= 1 && T >= 1, "conv: empty tensor");
  REQUIRE(precision >= 0 && precision <= 3, "conv: bad precision");
  CUDA_CHECK(cudaSetDevice(ctx->device));
  Arena ar;
  ar.guard = env_int("ALCM_GUARD", 0) != 0;
  ar.f16 = (precision == ALCM_PREC_FP16);
  const Env env{ctx, Knobs::from_env()};
  ConvLayer L = prepare_conv(ar, env.k, precision, kind, w, bias, Cout, Cin, K, p);
  PlaneT xin = make_planes(ar, B, Cin, T, opnd_esz(precision));
  PlaneT out = make_planes(ar, B, Cout, T * L.nphase, 4);
  PlaneT rp;
  if (res) rp = make_planes(ar, B, Cout, T * L.nphase, 4);
  CUDA_CHECK(sync_setup());
  launch_pack(x, xin, Cin, T, 1.f, precision, st);
  if (res) launch_pack(res, rp, Cout, T * L.nphase, 1.f, ALCM_PREC_FP32, st);
  OpList ol;
  RetileCache rcache;
  ol.env = env;
  ol.ar = &ar;
  ol.war = &ar; ol.cache = &rcache;  // same per-launch tile choice (N tile, K split, cluster reduction) as the plans
  ol.conv(L, xin, out, res ? &rp : nullptr);
  CUDA_CHECK(sync_setup());  // workspace memsets
  ol.run(st);
  launch_unpack(out, y, Cout, T * L.nphase, st);
  CUDA_CHECK(cudaGetLastError());
  sync_free(ar, st);
}

int alcm_conv1d_fwd(alcm_ctx* ctx, const float* x, const float* w, const float* bias, const float* res, float* y, int B, int Cin,
                    int Cout, int T, int K, int dilation, int precision, void* stream) {
  return guarded([&] {
    REQUIRE(dilation >= 1, "conv1d: dilation must be >= 1");
    run_conv_test(ctx, KIND_CONV, x, w, bias, res, y, B, Cin, Cout, T, K, dilation, precision, static_cast<cudaStream_t>(stream));
  });
}
int alcm_conv_transpose1d_fwd(alcm_ctx* ctx, const float* x, const float* w, const float* bias, float* y, int B, int Cin, int Cout,
                              int T, int stride, int precision, void* stream) {
  return guarded([&] {
    run_conv_test(ctx, KIND_CONVT, x, w, bias, nullptr, y, B, Cin, Cout, T, 2 * stride, stride, precision,
                  static_cast<cudaStream_t>(stream));
  });
}
int alcm_upsample_conv3_fwd(alcm_ctx* ctx, const float* x, const float* w, const float* bias, float* y, int B, int Cin, int Cout,
                            int T, int precision, void* stream) {
  return guarded([&] {
    run_conv_test(ctx, KIND_UPCONV3, x, w, bias, nullptr, y, B, Cin, Cout, T, 3, 1, precision, static_cast<cudaStream_t>(stream));
  });
}

int alcm_groupnorm_swish_fwd(alcm_ctx* ctx, const float* x, const float* gamma, const float* beta, float* y, int B, int C, int T,
                             int groups, float eps, int swish, void* stream) {
  return guarded([&] {
    REQUIRE(ctx && x && gamma && beta && y, "groupnorm: NULL argument");
    REQUIRE(B >= 1 && C >= 1 && T >= 1, "groupnorm: empty tensor");
    REQUIRE(groups == 32 && fabsf(eps - 1e-6f) < 1e-12f, "groupnorm: only GroupNorm(32, eps=1e-6) is on the path");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Arena ar;
    ar.guard = env_int("ALCM_GUARD", 0) != 0;
    PlaneT xin = make_planes(ar, B, C, T, 4), out = make_planes(ar, B, C, T, 4);
    GnP g = make_gn(ar, gamma, beta, C);
    OpList ol;
    ol.env = Env{ctx, Knobs::from_env()};
    op_gn(ol, ar, xin, out, g, swish, ALCM_PREC_FP32);
    CUDA_CHECK(sync_setup());
    launch_pack(x, xin, C, T, 1.f, ALCM_PREC_FP32, st);
    ol.run(st);
    launch_unpack(out, y, C, T, st);
    CUDA_CHECK(cudaGetLastError());
    sync_free(ar, st);
  });
}

int alcm_attn1d_fwd(alcm_ctx* ctx, const float* q, const float* k, const float* v, float* out, int B, int C, int T, int precision,
                    void* stream) {
  return guarded([&] {
    REQUIRE(ctx && q && k && v && out, "attn: NULL argument");
    REQUIRE(B >= 1 && C >= 1 && T >= 1, "attn: empty tensor");
    REQUIRE(precision >= 0 && precision <= 3, "attn: bad precision");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    Arena ar;
    ar.guard = env_int("ALCM_GUARD", 0) != 0;
    ar.f16 = (precision == ALCM_PREC_FP16);
    PlaneT pq = make_planes(ar, B, C, T, 4), pk = make_planes(ar, B, C, T, 4), pv = make_planes(ar, B, C, T, 4);
    PlaneT ph = make_planes(ar, B, C, T, 4);
    OpList ol;
    ol.env = Env{ctx, Knobs::from_env()};
    ol.ar = &ar;
    if (precision != ALCM_PREC_FP32 && ol.env.k.attn_tc) {
      push_attention_tc(ol, ar, pq, pk, pv, ph, B, C, T, precision);
    } else {
      float* Pm = static_cast<float*>(ar.alloc((size_t)B * T * T * 4, false));
      float* Sp = static_cast<float*>(ar.alloc((size_t)kAttnSplit * B * T * T * 4, false));
      push_attention(ol, pq, pk, pv, ph, Sp, Pm, B, C, T);
    }
    CUDA_CHECK(sync_setup());
    launch_pack(q, pq, C, T, 1.f, ALCM_PREC_FP32, st);
    launch_pack(k, pk, C, T, 1.f, ALCM_PREC_FP32, st);
    launch_pack(v, pv, C, T, 1.f, ALCM_PREC_FP32, st);
    ol.run(st);
    launch_unpack(ph, out, C, T, st);
    CUDA_CHECK(cudaGetLastError());
    sync_free(ar, st);
  });
}

// ---- measurement ----------------------------------------------------------------------------
// per-kernel CUDA-event timing of an op list (eager, serialised); accumulates into per-class and per-(stage, class) bins
static void profile_ops(const OpList& ol, int iters, alcm_profile* out, alcm_stage_profile* stages, int max_stages, cudaStream_t st) {
  std::vector<cudaEvent_t> ev(ol.ops.size() + 1);
  for (auto& e : ev) CUDA_CHECK(cudaEventCreate(&e));
  auto bin = [&](const Op& o) -> alcm_stage_profile* {
    if (!stages || o.stage < 0 || o.stage >= max_stages) return nullptr;
    return &stages[o.stage];
  };
  for (int it = 0; it < iters; ++it) {
    t_pdl = 0;  // per-kernel timing: no overlap between neighbours
    CUDA_CHECK(cudaEventRecord(ev[0], st));
    for (size_t i = 0; i < ol.ops.size(); ++i) {
      if (ol.ops[i].cls >= 0) ol.ops[i].fn(st);
      CUDA_CHECK(cudaEventRecord(ev[i + 1], st));
    }
    CUDA_CHECK(cudaStreamSynchronize(st));
    for (size_t i = 0; i < ol.ops.size(); ++i) {
      float ms = 0.f;
      CUDA_CHECK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
      const Op& o = ol.ops[i];
      if (o.cls < 0) continue;
      if (out) out->ms[o.cls] += ms;
      if (alcm_stage_profile* b = bin(o)) b->ms[o.cls] += ms;
    }
  }
  for (const Op& o : ol.ops) {
    if (o.cls < 0) continue;
    if (out) { out->flops[o.cls] += o.flops; out->bytes[o.cls] += o.bytes; out->launches[o.cls] += 1; }
    if (alcm_stage_profile* b = bin(o)) { b->flops[o.cls] += o.flops; b->bytes[o.cls] += o.bytes; b->launches[o.cls] += 1; }
  }
  for (auto& e : ev) cudaEventDestroy(e);
}

static void profile_any(alcm_vae* vae, alcm_vocoder* voc, int B, int T, int iters, alcm_profile* out, alcm_stage_profile* stages,
                        int max_stages, void* stream) {
  REQUIRE(voc && iters >= 1, "profile: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CUDA_CHECK(cudaSetDevice(voc->ctx->device));
  int Tmel = T;
  if (vae) {
    VaePlan* PV = vae_plan(vae, B, T, st);
    PlanUse u(PV, st);
    profile_ops(PV->ol, iters, out, stages, max_stages, st);
    u.finish();
    Tmel = PV->Tout;
  }
  VocPlan* PW = voc_plan(voc, B, Tmel, st);
  PlanUse u(PW, st);
  profile_ops(PW->ol, iters, out, stages, max_stages, st);
  u.finish();
}

int alcm_profile_decode(alcm_vae* vae, alcm_vocoder* voc, int B, int T, int iters, alcm_profile* out, void* stream) {
  return guarded([&] {
    REQUIRE(out, "profile: bad argument");
    memset(out, 0, sizeof(*out));
    profile_any(vae, voc, B, T, iters, out, nullptr, 0, stream);
  });
}

int alcm_profile_stages(alcm_vae* vae, alcm_vocoder* voc, int B, int T, int iters, alcm_stage_profile* stages, int max_stages,
                        void* stream) {
  return guarded([&] {
    REQUIRE(stages && max_stages >= 1, "profile_stages: bad argument");
    memset(stages, 0, sizeof(*stages) * (size_t)max_stages);
    profile_any(vae, voc, B, T, iters, nullptr, stages, max_stages, stream);
  });
}

// Micro-benchmarks run on RANDOM operands (seeded, generated on the device): zero-filled operands draw far less
// power, so the SM clock - and with it the measured rate - would not be the one real data sees.
static void fill_uniform(void* p, size_t n, int fmt, float lo, float hi, unsigned seed) {
  const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 16384);
  fill_uniform_kernel<<<blocks, 256>>>(p, n, fmt, lo, hi, seed);
  CUDA_CHECK(cudaGetLastError());
}

int alcm_bench_conv(alcm_ctx* ctx, int B, int Cin, int Cout, int T, int K, int dilation, int precision, int iters, int dbg,
                    float* ms_per_launch) {
  return guarded([&] {
    REQUIRE(ctx && ms_per_launch && iters >= 1, "bench_conv: bad argument");
    REQUIRE(precision != ALCM_PREC_FP32 || dbg == 0, "bench_conv: dbg flags need a tcgen05 mode");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    Arena ar;
    ar.guard = env_int("ALCM_GUARD", 0) != 0;
    ar.f16 = (precision == ALCM_PREC_FP16);
    const Env env{ctx, Knobs::from_env()};
    // Kaiming-uniform-like weights, N(0,1)-like activations: the magnitudes of the real model
    const float wb = 1.0f / sqrtf((float)Cin * K);
    float* w = static_cast<float*>(ar.alloc((size_t)Cout * Cin * K * 4, false));
    float* bias = static_cast<float*>(ar.alloc((size_t)Cout * 4, false));
    fill_uniform(w, (size_t)Cout * Cin * K, 0, -wb, wb, 1u);
    fill_uniform(bias, (size_t)Cout, 0, -wb, wb, 2u);
    ConvLayer L = prepare_conv(ar, env.k, precision, KIND_CONV, w, bias, Cout, Cin, K, dilation);
    PlaneT x = make_planes(ar, B, Cin, T, opnd_esz(precision)), out = make_planes(ar, B, Cout, T, 4);
    // fill whole planes (pads included - they only feed the first/last rows of each clip; irrelevant for timing)
    fill_uniform(x.p, x.bytes / (size_t)x.esz, x.g.fmt, -1.7f, 1.7f, 3u);
    OpList ol;
    RetileCache rcache;
    ol.env = env;
    ol.ar = &ar;
    ol.war = &ar; ol.cache = &rcache;  // same per-launch N tile choice as the plans
    ol.conv(L, x, out, nullptr);
    CUDA_CHECK(sync_setup());
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
    ctx->conv_dbg = dbg;
    for (int i = 0; i < 3; ++i) ol.run(0);
    CUDA_CHECK(cudaEventRecord(e0, 0));
    for (int i = 0; i < iters; ++i) ol.run(0);
    CUDA_CHECK(cudaEventRecord(e1, 0));
    cudaError_t err = cudaEventSynchronize(e1);
    ctx->conv_dbg = 0;
    CUDA_CHECK(err);
    if (env.k.trace && precision != ALCM_PREC_FP32) {  // one more launch with per-CTA timestamps
      int nt_c = 0, ks_c = 1;
      const bool clustered = choose_cluster_tile(env, L, T, B, &nt_c, &ks_c);
      const ConvLayer& Lt = retile(ar, rcache, L, clustered ? nt_c : pick_nt(env, L, T, B));
      const int ks = clustered ? ks_c : pick_ksplit(env, Lt, T, B);
      const size_t nctas = (size_t)conv_m_tiles(T) * Lt.n_tiles * B * Lt.nphase * ks;
      long long* tr = static_cast<long long*>(ar.alloc(nctas * 8 * sizeof(long long), true));
      CUDA_CHECK(sync_setup());
      ctx->conv_trace = tr;
      ol.run(0);
      ctx->conv_trace = nullptr;
      CUDA_CHECK(sync_setup());
      std::vector<long long> h(nctas * 8);
      CUDA_CHECK(cudaMemcpy(h.data(), tr, h.size() * 8, cudaMemcpyDeviceToHost));
      const size_t launched = (size_t)ctx->conv_last_grid;
      long long t_min = h[0], t_max = h[7], s_max = h[0];
      double d[5] = {0, 0, 0, 0, 0}, life = 0;
      for (size_t c = 0; c < launched; ++c) {
        const long long* t = &h[c * 8];
        t_min = std::min(t_min, t[0]); t_max = std::max(t_max, t[7]); s_max = std::max(s_max, t[0]);
        for (int i = 0; i < 5; ++i) d[i] += (double)(t[i + 2] - t[i + 1]);
        life += (double)(t[7] - t[0]);
      }
      fprintf(stderr, "  trace: %zu CTAs (ksplit %d, NT %d, kblk %d, stages ~), span %.1f us, last start +%.1f us, mean CTA life %.1f us; "
              "mean cycles: setup %.0f, first-operands %.0f, mma-issue %.0f, drain %.0f, epilogue %.0f\n",
              launched, ks, Lt.NT, Lt.kblk, (t_max - t_min) / 1e3, (s_max - t_min) / 1e3, life / launched / 1e3, d[0] / launched, d[1] / launched,
              d[2] / launched, d[3] / launched, d[4] / launched);
    }
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_launch = ms / iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  });
}

int alcm_bench_act(alcm_ctx* ctx, int B, int C, int T, int precision, int iters, float* ms_per_launch) {
  return guarded([&] {
    REQUIRE(ctx && ms_per_launch && iters >= 1, "bench_act: bad argument");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    Arena ar;
    ar.guard = env_int("ALCM_GUARD", 0) != 0;
    ar.f16 = (precision == ALCM_PREC_FP16);
    PlaneT x = make_planes(ar, B, C, T, 4), out = make_planes(ar, B, C, T, opnd_esz(precision));
    fill_uniform(x.p, x.bytes / 4, 0, -1.7f, 1.7f, 5u);
    float* al = static_cast<float*>(ar.alloc((size_t)round_up(C, 16) * 4, false));
    float* be = static_cast<float*>(ar.alloc((size_t)round_up(C, 16) * 4, false));
    fill_uniform(al, (size_t)round_up(C, 16), 0, -0.9f, 0.9f, 6u);   // alpha, beta ~ spread of N(0, 0.5)
    fill_uniform(be, (size_t)round_up(C, 16), 0, -0.9f, 0.9f, 7u);
    SnakeP sp = make_snake(ar, al, be, C, 0);
    OpList ol;
    ol.env = Env{ctx, Knobs::from_env()};
    ol.act(x, out, sp.ea, sp.ib, precision == ALCM_PREC_TF32, precision != ALCM_PREC_FP32);
    CUDA_CHECK(sync_setup());
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) ol.run(0);
    CUDA_CHECK(cudaEventRecord(e0, 0));
    for (int i = 0; i < iters; ++i) ol.run(0);
    CUDA_CHECK(cudaEventRecord(e1, 0));
    cudaError_t err = cudaEventSynchronize(e1);
    CUDA_CHECK(err);
    float ms = 0.f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_launch = ms / iters;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
  });
}

// ALCM_GUARD=1 self-check: bytes of the guard zones around the model's weights and between the buffers of its plans
// that are no longer zero (0 = no kernel wrote out of bounds).  Waits for the plans' pending launches.
int alcm_vocoder_check_guards(alcm_vocoder* v, long long* bad) {
  return guarded([&] {
    REQUIRE(v && bad, "check_guards: NULL argument");
    CUDA_CHECK(cudaSetDevice(v->ctx->device));
    long long n = v->war.guard_violations();
    for (auto& kv : v->plans) {
      if (kv.second->used) CUDA_CHECK(cudaEventSynchronize(kv.second->done));
      n += kv.second->ar.guard_violations();
    }
    *bad = n;
  });
}
int alcm_vae_check_guards(alcm_vae* v, long long* bad) {
  return guarded([&] {
    REQUIRE(v && bad, "check_guards: NULL argument");
    CUDA_CHECK(cudaSetDevice(v->ctx->device));
    long long n = v->war.guard_violations();
    for (auto& kv : v->plans) {
      if (kv.second->used) CUDA_CHECK(cudaEventSynchronize(kv.second->done));
      n += kv.second->ar.guard_violations();
    }
    *bad = n;
  });
}

int alcm_vocoder_launches(alcm_vocoder* v, int B, int T) {
  int n = -1;
  guarded([&] {
    REQUIRE(v, "NULL vocoder");
    n = voc_plan(v, B, T, nullptr)->ol.launches() + 2;  // + mel pack + conv_post
  });
  return n;
}
int alcm_vae_launches(alcm_vae* v, int B, int T) {
  int n = -1;
  guarded([&] {
    REQUIRE(v, "NULL vae");
    n = vae_plan(v, B, T, nullptr)->ol.launches() + 2;  // + latent pack + mel unpack
  });
  return n;
}

}  // extern "C"

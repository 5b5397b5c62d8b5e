/* audiolcm_b200 - C-ABI of the B200-native latent->waveform decode path of AudioLCM.
 *
 * The reference (Text-to-Audio/AudioLCM) is pure PyTorch and has no FFI; these entry points are
 * what a binding for its two decode call sites would bind:
 *
 *   alcm_vocoder_create / alcm_vocode      <- VocoderBigVGAN.__init__ / .vocode
 *                                             /root/reference/vocoder/bigvgan/models.py:393-414
 *                                             (generator forward: models.py:181-203)
 *   alcm_vae_create / alcm_vae_decode      <- AutoencoderKL.decode as called by
 *                                             LCM_audio.decode_first_stage
 *                                             /root/reference/ldm/models/autoencoder1d.py:59-62,484-517
 *                                             /root/reference/ldm/models/diffusion/lcm_audio.py:392-406
 *   alcm_decode_to_wav                     <- the caller loop that chains both
 *                                             /root/reference/pythonscripts/InferAPI.py:87-96
 *
 * Conventions: plain pointers and sizes only.  All tensor pointers are DEVICE pointers to
 * contiguous fp32 arrays in the reference's own layouts ([B,C,T], weights as in the state_dict).
 * The caller owns inputs and outputs; the library owns packed weights and one workspace slab per
 * (B,T) shape ("plan").  alcm_*_plan() builds the plan of a shape ahead of time and
 * alcm_*_workspace_bytes() reports its size; a shape that was not planned is planned by the first
 * call that uses it.  Once a shape is planned, the decode calls perform NO allocation and NO
 * synchronisation: they enqueue kernels on `stream` and return.  Plan memory is stream-ordered
 * (cudaMallocAsync); the least recently used plans beyond ALCM_MAX_PLANS are retired without waiting.
 * A pre-planned call may be recorded into the caller's own CUDA graph (stream capture): plan first, capture after.
 * `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Functions return
 * 0 on success or a negative code; alcm_last_error() gives the message (per calling thread).  Nothing
 * aborts or throws across this boundary.
 * Threading: the library keeps no process-global mutable state.  An alcm_ctx may be shared by
 * threads; a model handle (alcm_vocoder / alcm_vae) may be used by one thread at a time.  Calls on
 * one handle from different streams are ordered on the device (they share the plan's buffers).
 * Run-time knobs (ALCM_* environment variables) are read once, when a model handle is
 * created.
 */
#ifndef AUDIOLCM_B200_H
#define AUDIOLCM_B200_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct alcm_ctx alcm_ctx;
typedef struct alcm_vocoder alcm_vocoder;
typedef struct alcm_vae alcm_vae;
typedef struct alcm_conv1d alcm_conv1d;
typedef struct alcm_ffn1d alcm_ffn1d;
typedef struct alcm_melspec alcm_melspec;
typedef struct alcm_vae_encoder alcm_vae_encoder;

/* arithmetic used by the conv GEMMs */
enum {
  ALCM_PREC_FP32 = 0, /* CUDA-core FFMA, fp32 end to end (exact mode; slow)                        */
  ALCM_PREC_TF32 = 1, /* tcgen05 kind::tf32, fp32 storage, fp32 accumulate  ("fp32 mode")          */
  ALCM_PREC_BF16 = 2, /* tcgen05 kind::f16 (bf16 operands), fp32 residual stream and accumulators  */
  ALCM_PREC_FP16 = 3  /* tcgen05 kind::f16 (fp16 operands: tf32's 10-bit mantissa at bf16's speed and bytes; conversions
                         saturate at +-65504 - for networks whose activations stay in fp16 range), fp32 elsewhere */
};

enum {
  ALCM_OK = 0,
  ALCM_ERR_INVALID = -1, /* bad argument / unsupported configuration */
  ALCM_ERR_CUDA = -2,    /* CUDA runtime error                      */
  ALCM_ERR_INTERNAL = -3
};

int alcm_ctx_create(alcm_ctx** out, int device);
void alcm_ctx_destroy(alcm_ctx* ctx);
/* message of the last failing call on this thread (never NULL) */
const char* alcm_last_error(void);

/* ---- BigVGAN vocoder ------------------------------------------------------------------------
 * cfg mirrors the generator keys of bigvgan_audioset16khz_80band.json.  resblock "1" (AMPBlock1, models.py:28-81) or
 * "2" (AMPBlock2, :90-126; `resblock2` = 1, two dilations per kernel size); activation "snakebeta" or "snake"
 * (activations.py:9-120; for "snake" pass each alpha twice, as alpha and as beta); snake_logscale true or false
 * (`snake_linear` = 1).  A zero-filled tail of the struct is the 16k config: AMPBlock1, log-scale parameters.
 * `tensors` lists device fp32 pointers in this order (state_dict names in brackets):
 *   conv_pre:   weight_g, weight_v, bias
 *   for i in 0..num_upsamples-1:
 *     ups.i.0:  weight_g, weight_v, bias
 *     for j in 0..num_kernels-1   [resblocks.(i*num_kernels+j)]:
 *       convs1.0..2: (weight_g, weight_v, bias) x3
 *       convs2.0..2: (weight_g, weight_v, bias) x3
 *       activations.0..5.act: (alpha, beta) x6
 *       [resblock2: convs.0..1 x (weight_g, weight_v, bias), activations.0..1.act x (alpha, beta) instead]
 *   activation_post.act: alpha, beta
 *   conv_post:  weight_g, weight_v, bias
 */
typedef struct {
  int num_mels;
  int upsample_initial_channel;
  int num_upsamples;
  int num_kernels;
  int upsample_rates[8];
  int upsample_kernel_sizes[8];
  int resblock_kernel_sizes[4];
  int resblock_dilation_sizes[4][3];
  int resblock2;     /* 0: AMPBlock1 (3 dilations), 1: AMPBlock2 (resblock_dilation_sizes[j][0..1]) */
  int snake_linear;  /* 0: alpha_logscale (exp of the stored parameters), 1: parameters used as stored */
} alcm_bigvgan_cfg;

int alcm_vocoder_num_tensors(const alcm_bigvgan_cfg* cfg);
int alcm_vocoder_create(alcm_ctx* ctx, const alcm_bigvgan_cfg* cfg, const float* const* tensors, int n_tensors,
                        int precision, alcm_vocoder** out);
void alcm_vocoder_destroy(alcm_vocoder* v);
/* mel [B,num_mels,T] -> wav [B, T*prod(upsample_rates)] (conv_post + tanh applied) */
int alcm_vocode(alcm_vocoder* v, const float* mel, int B, int T, float* wav, void* stream);
/* same, output as 16-bit PCM: pcm[b][t] = rint(wav * 32767) - the payload soundfile.write(path, wav, 16000) stores
 * (pythonscripts/InferAPI.py:98) */
int alcm_vocode_pcm16(alcm_vocoder* v, const float* mel, int B, int T, short* pcm, void* stream);
/* build the plan (workspace slab + kernel list + CUDA graph) of shape (B,T) now, on `stream` */
int alcm_vocoder_plan(alcm_vocoder* v, int B, int T, void* stream);
/* device bytes the plan of shape (B,T) occupies (sizing pass only when the shape is not planned yet) */
int alcm_vocoder_workspace_bytes(alcm_vocoder* v, int B, int T, size_t* bytes);

/* ---- 1-D KL-VAE decoder -----------------------------------------------------------------------
 * cfg mirrors ddconfig of configs/audiolcm.yaml:54-70.  upsample_levels[l] = 1 when level l ends
 * with an Upsample1D (reference: l in [d+1 for d in down_layers]).
 * `tensors` order (each conv = weight,bias; each norm = weight,bias; resblock = norm1, conv1,
 * norm2, conv2, then nin_shortcut iff in!=out):
 *   post_quant_conv, decoder.conv_in, mid.block_1, mid.attn_1 (norm, q, k, v, proj_out),
 *   mid.block_2, then for level = n_levels-1 .. 0: block.0 .. block.num_res_blocks,
 *   [upsample.conv], finally norm_out, conv_out.
 */
typedef struct {
  int ch;
  int out_ch;
  int z_channels;
  int embed_dim;
  int kernel_size;
  int num_res_blocks;
  int n_levels;
  int ch_mult[8];
  int upsample_levels[8];
  int attn_levels[8];     /* attn_levels[l] = 1 when level l is in attn_layers: an AttnBlock1D (norm, q, k, v, proj_out) follows
                             every ResnetBlock1D of that level (autoencoder1d.py:466-468,500-504); zero-filled = shipped config */
} alcm_vae_cfg;

int alcm_vae_num_tensors(const alcm_vae_cfg* cfg);
int alcm_vae_create(alcm_ctx* ctx, const alcm_vae_cfg* cfg, const float* const* tensors, int n_tensors, int precision,
                    alcm_vae** out);
void alcm_vae_destroy(alcm_vae* v);
/* z [B,embed_dim,T] -> mel [B,out_ch,T*2^n_up];  inv_scale = 1/scale_factor (lcm_audio.py:400) */
int alcm_vae_decode(alcm_vae* v, const float* z, int B, int T, float inv_scale, float* mel, void* stream);
int alcm_vae_plan(alcm_vae* v, int B, int T, void* stream);
int alcm_vae_workspace_bytes(alcm_vae* v, int B, int T, size_t* bytes);

/* ---- 1-D KL-VAE encoder (the step on the other side of the path, used by reconstruct_audio.py:115):
 * AutoencoderKL.encode up to the posterior's parameters (ldm/models/autoencoder1d.py:52-56,319-413).
 * downsample_levels[l] = 1 when level l ends with a Downsample1D (l in down_layers).  `tensors` order (conv = weight,bias;
 * norm = weight,bias; resblock = norm1, conv1, norm2, conv2[, nin_shortcut]):
 *   encoder.conv_in, for level = 0 .. n_levels-1: block.0 .. block.num_res_blocks-1, [downsample.conv],
 *   mid.block_1, mid.attn_1 (norm, q, k, v, proj_out), mid.block_2, norm_out, conv_out, quant_conv. */
typedef struct {
  int ch;
  int in_channels;
  int z_channels;
  int embed_dim;
  int kernel_size;
  int num_res_blocks;
  int n_levels;
  int double_z;
  int ch_mult[8];
  int downsample_levels[8];
  int attn_levels[8];     /* same meaning as in the decoder configuration above: autoencoder1d.py:356-358,391-396 */
} alcm_vae_enc_cfg;
int alcm_vae_encoder_num_tensors(const alcm_vae_enc_cfg* cfg);
int alcm_vae_encoder_create(alcm_ctx* ctx, const alcm_vae_enc_cfg* cfg, const float* const* tensors, int n_tensors, int precision,
                            alcm_vae_encoder** out);
void alcm_vae_encoder_destroy(alcm_vae_encoder* v);
/* mel [B,in_channels,T] -> moments [B, 2*embed_dim, T / 2^n_down] = (mean | logvar) of the posterior */
int alcm_vae_encode(alcm_vae_encoder* v, const float* x, int B, int T, float* moments, void* stream);

/* latent -> waveform with the mel kept on the device (mel_out may be NULL) */
int alcm_decode_to_wav(alcm_vae* vae, alcm_vocoder* voc, const float* z, int B, int T, float inv_scale, float* mel_out,
                       float* wav, void* stream);
/* same with the waveform packed as 16-bit PCM on the device (the batched driver that replaces InferAPI.py:87-98) */
int alcm_decode_to_pcm16(alcm_vae* vae, alcm_vocoder* voc, const float* z, int B, int T, float inv_scale, float* mel_out,
                         short* pcm, void* stream);

/* ---- LCM sampler step: LCMSampler.step (ldm/models/diffusion/scheduling_lcm.py:411-494, epsilon prediction) as one
 * elementwise kernel over n fp32 elements (device pointers, 16-byte aligned, n % 4 == 0):
 *   x0 = (sample - sqrt_beta_prod_t*eps)/sqrt_alpha_prod_t;  denoised = c_out*x0 + c_skip*sample;
 *   prev = last_step ? denoised : sqrt_alpha_prod_prev*denoised + sqrt_beta_prod_prev*noise.
 * The coefficients are the host scalars the reference computes per step (:441-452,401-409). */
int alcm_lcm_step(alcm_ctx* ctx, const float* sample, const float* eps, const float* noise, float* prev, float* denoised, long long n,
                  float sqrt_alpha_prod_t, float sqrt_beta_prod_t, float c_out, float c_skip, float sqrt_alpha_prod_prev,
                  float sqrt_beta_prod_prev, int last_step, void* stream);

/* nn.LayerNorm(C) (new_attention.py:246-248 norm1/2/3, eps 1e-5) applied over the channel axis of a channels-first
 * tensor: y[b,:,t] = (x[b,:,t] - mean) * rsqrt(var + eps) * gamma + beta.  x, y [B,C,T], gamma, beta [C]: device fp32. */
int alcm_layernorm_cf(alcm_ctx* ctx, const float* x, const float* gamma, const float* beta, float* y, int B, int C, int T, float eps,
                      void* stream);

/* ---- a Conv1d layer as a persistent handle: weights packed once, one plan per (B,T) shape.  Used for the 9-tap
 * Conv1dFeedForward convs of the DiT denoiser (ldm/modules/new_attention.py:48-74; ConcatDiT2MLP blocks,
 * concatDiT.py:108-130), 93 % of the denoiser's FLOPs.  w [Cout,Cin,K] and bias [Cout] are device fp32; padding is
 * (K*dilation-dilation)/2, K odd <= 11.  x [B,Cin,T], res (may be NULL) and y [B,Cout,T] are device fp32:
 * y = conv(x) + bias (+ res). */
int alcm_conv1d_create(alcm_ctx* ctx, const float* w, const float* bias, int Cout, int Cin, int K, int dilation, int precision,
                       alcm_conv1d** out);
void alcm_conv1d_destroy(alcm_conv1d* c);
int alcm_conv1d_run(alcm_conv1d* c, const float* x, const float* res, float* y, int B, int T, void* stream);

/* ---- Conv1dFeedForward(dim, mult, glu=True, kernel_size=K) of the DiT blocks as ONE handle / one plan per (B,T)
 * (ldm/modules/new_attention.py:38-74): h = Conv1d(dim -> 2*inner, K)(x); y = Conv1d(inner -> dim_out, K)(h[:inner] *
 * gelu(h[inner:])) (+ res).  The 2*inner-channel intermediate never leaves the operand layout (GEGLU is one kernel
 * between the two tcgen05 convs).  w_in [2*inner,dim,K], b_in [2*inner], w_out [dim_out,inner,K], b_out [dim_out]:
 * device fp32 (biases may be NULL); inner % 8 == 0, K odd <= 11.  x [B,dim,T], res (may be NULL), y [B,dim_out,T]. */
int alcm_ffn1d_create(alcm_ctx* ctx, const float* w_in, const float* b_in, const float* w_out, const float* b_out, int dim, int inner,
                      int dim_out, int K, int precision, alcm_ffn1d** out);
void alcm_ffn1d_destroy(alcm_ffn1d* c);
int alcm_ffn1d_run(alcm_ffn1d* c, const float* x, const float* res, float* y, int B, int T, void* stream);

/* ---- log10-mel front-end, MelNet.forward (ldm/data/preprocess/NAT_mel.py:64-85), as one plan per (B,L):
 * clamp to [-1,1] + reflect pad (n_fft-hop)/2 + fold into hop-sized rows (one kernel) -> |STFT| as a (taps+1)-tap Conv1d
 * over the folded rows + magnitude kernel -> mel filterbank as a 1x1 Conv1d -> log10(max(., 1e-5)).  n_fft = taps*hop.
 * stft_w [2*nb_pad, hop, taps+1]: the windowed DFT basis, real rows [0,nb) and imaginary rows [nb_pad, nb_pad+nb), tap 0
 * zero, tap j = basis columns [(j-1)*hop, j*hop); mel_w [n_mels, nb_pad, 1]; nb_pad = n_fft/2+1 rounded up to 8.
 * y [B,L] (L a multiple of hop), mel [B,n_mels,L/hop]: device fp32. */
int alcm_melspec_create(alcm_ctx* ctx, const float* stft_w, const float* mel_w, int hop, int taps, int nb_pad, int n_mels, int precision,
                        alcm_melspec** out);
void alcm_melspec_destroy(alcm_melspec* c);
int alcm_melspec_run(alcm_melspec* c, const float* y, float* mel, int B, int L, void* stream);

/* ---- single-op entry points (tests / micro-benchmarks); tensors are [B,C,T] fp32 on device -----*/
/* Activation1d(SnakeBeta logscale): act.py:23-28.  precision BF16 returns bf16-rounded values. */
int alcm_activation1d_fwd(alcm_ctx* ctx, const float* x, const float* alpha, const float* beta, float* y, int B, int C,
                          int T, int precision, void* stream);
/* Conv1d(Cin,Cout,K,dilation, padding=(K*d-d)/2) (+bias, +res if non-NULL); w [Cout,Cin,K] */
int alcm_conv1d_fwd(alcm_ctx* ctx, const float* x, const float* w, const float* bias, const float* res, float* y, int B,
                    int Cin, int Cout, int T, int K, int dilation, int precision, void* stream);
/* ConvTranspose1d(Cin,Cout,K=2*stride,stride,padding=stride/2); w [Cin,Cout,K]; y [B,Cout,T*stride] */
int alcm_conv_transpose1d_fwd(alcm_ctx* ctx, const float* x, const float* w, const float* bias, float* y, int B, int Cin,
                              int Cout, int T, int stride, int precision, void* stream);
/* nearest x2 + Conv1d(C,C,3,p=1): autoencoder1d.py:291-295; y [B,Cout,2T] */
int alcm_upsample_conv3_fwd(alcm_ctx* ctx, const float* x, const float* w, const float* bias, float* y, int B, int Cin,
                            int Cout, int T, int precision, void* stream);
/* GroupNorm(groups, eps) [+ swish] */
int alcm_groupnorm_swish_fwd(alcm_ctx* ctx, const float* x, const float* gamma, const float* beta, float* y, int B, int C,
                             int T, int groups, float eps, int swish, void* stream);
/* softmax_j(q^T k * C^-0.5) applied to v: autoencoder1d.py:264-275; q,k,v,out [B,C,T].  TF32 / BF16: both GEMMs on
 * conv_umma_kernel with per-item operands; FP32: CUDA-core kernels */
int alcm_attn1d_fwd(alcm_ctx* ctx, const float* q, const float* k, const float* v, float* out, int B, int C, int T,
                    int precision, void* stream);

/* ---- measurement ------------------------------------------------------------------------------
 * Kernel classes for per-class device timing. */
enum { ALCM_CLS_CONV = 0, ALCM_CLS_ACT = 1, ALCM_CLS_NORM = 2, ALCM_CLS_ATTN = 3, ALCM_CLS_MISC = 4, ALCM_NUM_CLS = 5 };
typedef struct {
  double ms[ALCM_NUM_CLS];       /* summed CUDA-event time per class over `iters` runs */
  double flops[ALCM_NUM_CLS];    /* algorithmic FLOPs per run (conv: 2*Cin*Cout*k*T)     */
  double bytes[ALCM_NUM_CLS];    /* algorithmic HBM bytes per run (one read + one write) */
  int launches[ALCM_NUM_CLS];    /* kernel launches per run                              */
} alcm_profile;
/* runs the vocoder (vae may be NULL) / vae+vocoder op lists eagerly with an event pair around every
 * kernel; the normal path replays a CUDA graph and has no events inside. */
int alcm_profile_decode(alcm_vae* vae, alcm_vocoder* voc, int B, int T, int iters, alcm_profile* out, void* stream);
/* the same measurement binned by pipeline stage: stages[0] = VAE decoder, [1] = conv_pre, [2..7] = vocoder stages
 * 1..6 (upsampler + 3 AMP blocks each), [8] = activation_post (conv_post+tanh is outside the op list) */
typedef alcm_profile alcm_stage_profile;
int alcm_profile_stages(alcm_vae* vae, alcm_vocoder* voc, int B, int T, int iters, alcm_stage_profile* stages, int max_stages,
                        void* stream);
/* times `iters` back-to-back launches of one Conv1d(Cin,Cout,K,dilation) on seeded RANDOM weights and
 * activations of shape [B,Cin,T] (generated on the device; zero operands would run at a different power / clock
 * point); dbg is for kernel bring-up (bit0/bit1 skip the weight/activation copies) */
int alcm_bench_conv(alcm_ctx* ctx, int B, int Cin, int Cout, int T, int K, int dilation, int precision, int iters, int dbg,
                    float* ms_per_launch);
/* same for one Activation1d launch on [B,C,T] (random x, alpha, beta) */
int alcm_bench_act(alcm_ctx* ctx, int B, int C, int T, int precision, int iters, float* ms_per_launch);
/* ALCM_GUARD=1 self-check (set before the handle is created): every device buffer the handle owns is separated from
 * its neighbours by 4 KB zones of zeros that no kernel may touch; *bad = bytes of those zones that changed (0 = no
 * out-of-bounds write since the handle was created).  Waits for the handle's pending launches. */
int alcm_vocoder_check_guards(alcm_vocoder* v, long long* bad);
int alcm_vae_check_guards(alcm_vae* v, long long* bad);
/* kernels launched by one alcm_vocode / alcm_vae_decode call for this shape (after planning) */
int alcm_vocoder_launches(alcm_vocoder* v, int B, int T);
int alcm_vae_launches(alcm_vae* v, int B, int T);

#ifdef __cplusplus
}
#endif
#endif /* AUDIOLCM_B200_H */

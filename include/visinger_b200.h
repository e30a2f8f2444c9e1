/*
 * visinger_b200.h -- C ABI of the B200-native VISinger inference hot path.
 *
 * The reference (jisang93/VISinger) is pure Python/PyTorch and has NO FFI, plugin or
 * operator interface of its own: the drop-in boundary is its nn.Module surface
 * (SURVEY.md section 8b).  This header is the C boundary underneath our mirror of that
 * surface; each entry point names the reference interface it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross this boundary;
 *   - activations are the reference's own tensors: fp32, [B, C, T] contiguous
 *     (time fastest), resident on the pack's CUDA device; `mask` is [B, 1, T] in {0,1};
 *     `g` is the speaker condition [B, gin, 1] or NULL (modules/visinger/flow.py:33,
 *     modules/visinger/decoder.py:40);
 *   - weights enter as the reference's state-dict entries (HOST fp32 pointers, the
 *     names and shapes of SURVEY.md Appendix A, either `weight_g`+`weight_v` or a
 *     plain `weight` after remove_weight_norm); vsg_pack_create folds weight-norm and
 *     the Flip permutations, re-lays the weights out for the kernels and uploads them;
 *   - the caller owns every input, output and workspace buffer; the library allocates
 *     device memory only inside vsg_pack_create and frees it in vsg_pack_destroy.
 *     No call synchronises the host or allocates: all run calls are asynchronous on
 *     `stream` (a cudaStream_t passed as void*) and CUDA-graph capturable;
 *   - every function returns 0 on success or a negative VSG_E* code and never throws;
 *     vsg_last_error() returns a thread-local, human readable message for the last
 *     failing call;
 *   - a VsgPack is immutable after creation and may be shared by several streams of
 *     its device; calls that share one workspace must be stream-ordered.
 */
#ifndef VISINGER_B200_H_
#define VISINGER_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSG_ABI_VERSION 1

enum {
  VSG_OK = 0,
  VSG_EINVAL = -1,   /* bad argument / shape / missing weight             */
  VSG_ECUDA = -2,    /* a CUDA runtime or driver call failed               */
  VSG_ENOMEM = -3,   /* workspace too small / allocation failed            */
  VSG_EUNSUPPORTED = -4 /* configuration outside what the kernels implement */
};

/* Arithmetic mode of a run call. */
enum {
  VSG_PRECISION_FP32 = 0, /* fp32 storage, fp32 FFMA accumulate: the parity mode
                             (waveform max-abs <= 1e-4, flow z <= 1e-5 vs the reference) */
  VSG_PRECISION_BF16 = 1, /* bf16 operands + storage, fp32 accumulate on tcgen05 tensor
                             cores (TMEM accumulators, TMA-fed): the throughput mode     */
  VSG_PRECISION_BF16X3 = 2 /* split-bf16 on the same tensor-core kernels, fp32-mode tolerances on tcgen05.  Decoder:
                             every activation and weight is a (hi, lo) bf16 pair (~16 mantissa bits), every
                             product three MMAs (hi*hi + hi*lo + lo*hi).  Flow (z <= 1e-5 on a state of magnitude ~5):
                             three planes (hi, mid, lo = the fp32 value exactly), six plane products per MMA product,
                             issued small-first because the tensor pipe's fp32 accumulation truncates            */
};

#define VSG_MAX_UPS 8
#define VSG_MAX_RESBLOCK_KERNELS 4
#define VSG_MAX_RESBLOCK_DILATIONS 4

/* Constructor arguments of the two reference modules.  A part with n_flows == 0 /
 * n_ups == 0 is absent from the pack. */
typedef struct VsgConfig {
  /* ResidualCouplingBlock(channels, hidden_channels, kernel_size, dilation_rate,
   * n_layers, n_flows, gin_channels)            modules/visinger/flow.py:16-31 */
  int32_t flow_channels;
  int32_t flow_hidden;
  int32_t flow_kernel_size;
  int32_t flow_dilation_rate;
  int32_t flow_n_layers;
  int32_t flow_n_flows;
  int32_t flow_gin;
  /* Generator(initial_channel, resblock, resblock_kernel_sizes, resblock_dilation_sizes,
   * upsample_rates, upsample_initial_channel, upsample_kernel_sizes, gin_channels)
   *                                             modules/visinger/decoder.py:14-38 */
  int32_t dec_initial_channel;
  int32_t dec_resblock;      /* 1 = ResBlock1, 2 = ResBlock2 */
  int32_t dec_n_kernels;     /* len(resblock_kernel_sizes) */
  int32_t dec_resblock_kernel_sizes[VSG_MAX_RESBLOCK_KERNELS];
  int32_t dec_n_dilations[VSG_MAX_RESBLOCK_KERNELS];
  int32_t dec_resblock_dilations[VSG_MAX_RESBLOCK_KERNELS][VSG_MAX_RESBLOCK_DILATIONS];
  int32_t dec_n_ups;         /* len(upsample_rates) */
  int32_t dec_upsample_rates[VSG_MAX_UPS];
  int32_t dec_upsample_kernel_sizes[VSG_MAX_UPS];
  int32_t dec_upsample_initial_channel;
  int32_t dec_gin;
} VsgConfig;

/* One state-dict entry (host memory, fp32, contiguous). */
typedef struct VsgTensor {
  const char* name;   /* e.g. "flow.flows.0.enc.in_layers.2.weight_v", "decoder.ups.0.bias" */
  const float* data;
  int32_t ndim;
  int64_t shape[4];
} VsgTensor;

typedef struct VsgPack VsgPack;

/* ABI version of the loaded library (== VSG_ABI_VERSION of the header it was built from). */
int vsg_abi_version(void);

/* Thread-local message of the last failing call ("" if none). */
const char* vsg_last_error(void);

/*
 * One-time weight pre-pack.  Replaces what the reference does implicitly on every
 * forward: the weight_norm pre-hook (torch.nn.utils.weight_norm at
 * modules/visinger/encoder.py:147,154,164 and modules/visinger/decoder.py:24-26,72-87)
 * and Flip (modules/visinger/flow.py:88-95, folded into channel permutations of
 * pre/post).  `weights` uses the key layout of the reference checkpoint
 * ["state_dict"]["model"] (utils/commons/ckpt_utils.py:37-56) restricted to the two
 * modules: flow weights under prefix `flow_prefix`, decoder weights under `dec_prefix`
 * (e.g. "flow." / "decoder.", or "" when packing a stand-alone module).
 */
int vsg_pack_create(const VsgConfig* cfg, const VsgTensor* weights, int32_t n_weights,
                    const char* flow_prefix, const char* dec_prefix, int32_t device, VsgPack** out);
void vsg_pack_destroy(VsgPack* pack);

/* Bytes of caller-provided scratch each run call needs for (B, T) in `precision`
 * (the max over the three run calls).  0 on invalid arguments. */
size_t vsg_workspace_bytes(const VsgPack* pack, int32_t B, int32_t T, int32_t precision);

/*
 * Prior sampling, models/visinger.py:107:
 *   z_p = (mu_p + noise * exp(logs_p)) * mask         all [B, C, T], mask [B, 1, T]
 * `noise` is injected by the caller (the reference draws torch.randn_like(mu_p)).
 */
int vsg_prior_sample(const float* mu_p, const float* logs_p, const float* noise, const float* mask,
                     float* z_p, int32_t B, int32_t C, int32_t T, void* stream);

/*
 * ResidualCouplingBlock.forward(x, x_mask, g, reverse)   modules/visinger/flow.py:33-40
 * (coupling layers flow.py:66-85, WaveNet modules/visinger/encoder.py:167-195 with
 * fused_add_tanh_sigmoid_multiply encoder.py:206-213).  x, y: [B, channels, T]; y may
 * alias x.  `g` NULL iff flow_gin == 0.  reverse != 0 is the inference direction.
 */
int vsg_flow_forward(const VsgPack* pack, const float* x, const float* mask, const float* g, float* y,
                     int32_t B, int32_t T, int32_t reverse, int32_t precision,
                     void* workspace, size_t workspace_bytes, void* stream);

/*
 * Generator.forward(x, g)                                 modules/visinger/decoder.py:40-59
 * z: [B, initial_channel, T] -> wav: [B, 1, T * prod(upsample_rates)].
 */
int vsg_generator_forward(const VsgPack* pack, const float* z, const float* g, float* wav,
                          int32_t B, int32_t T, int32_t precision,
                          void* workspace, size_t workspace_bytes, void* stream);

/*
 * The whole inference hot path, models/visinger.py:107-111:
 *   z_p = (mu_p + noise*exp(logs_p))*mask ; z_q = flow(z_p, mask, g, reverse=True)*mask ;
 *   wav = decoder(z_q*mask, g)
 * `z_q_out` (may be NULL) receives z_q [B, channels, T].
 */
int vsg_infer(const VsgPack* pack, const float* mu_p, const float* logs_p, const float* noise,
              const float* mask, const float* g, float* wav, float* z_q_out,
              int32_t B, int32_t T, int32_t precision,
              void* workspace, size_t workspace_bytes, void* stream);

/*
 * Output stage, `save_wav(wav, path, sr, norm)` up to the file write        utils/audio/io.py:8-14
 * (called with norm = out_wav_norm: true at inference/visinger.py:100, tasks/visinger.py:258):
 *   pcm = int16( (norm ? wav / max|wav| : wav) * 32767 )     fp32 arithmetic, conversion truncating toward zero
 * -- the integers numpy produces, bit for bit.  wav: device fp32 [B, L] (the Generator's output); `lengths` (device
 * int32 [B], valid SAMPLES per utterance, or NULL = L): the peak is taken over the valid samples of each utterance
 * only (the reference runs one utterance per call) and samples beyond them are written as 0; pcm: device int16 [B, L];
 * peak: device fp32 [B], receives max|wav| per utterance.  Asynchronous on `stream`, no allocation, capturable.
 */
int vsg_wav_to_int16(const float* wav, const int32_t* lengths, int16_t* pcm, float* peak, int32_t B, int32_t L,
                     int32_t norm, void* stream);

/*
 * PosteriorEncoder(in_channels, out_channels, hidden_channels, kernel_size, dilation_rate, n_layers, gin_channels)
 *                                                                         modules/visinger/encoder.py:76-101
 * (constructed at models/visinger.py:59-60 as (num_linear_bins, 192, 192, 5, 1, 16, gin); the training / voice-conversion
 * side of the model: SURVEY.md section 8 row f4).  It re-uses the WaveNet kernels of the flow.
 */
typedef struct VsgEncConfig {
  int32_t in_channels;
  int32_t out_channels;
  int32_t hidden_channels;
  int32_t kernel_size;
  int32_t dilation_rate;
  int32_t n_layers;
  int32_t gin_channels;
} VsgEncConfig;

/* Pre-pack of a PosteriorEncoder's state-dict entries (`prefix` + "pre.weight", "enc.in_layers.0.weight_v", ...,
 * "proj.bias"); the result is an ordinary VsgPack (vsg_pack_destroy frees it) that only vsg_posterior_* accept. */
int vsg_enc_pack_create(const VsgEncConfig* cfg, const VsgTensor* weights, int32_t n_weights, const char* prefix,
                        int32_t device, VsgPack** out);
size_t vsg_posterior_workspace_bytes(const VsgPack* pack, int32_t B, int32_t T, int32_t precision);

/*
 * PosteriorEncoder.forward(x, nonpadding, g)                              modules/visinger/encoder.py:92-98
 *   h = pre(x) * mask ; h = WaveNet(h, mask, g) ; stats = proj(h) * mask ; (mu_q, logs_q) = split(stats)
 *   z_q = (mu_q + noise * exp(logs_q)) * mask
 * x: [B, in_channels, T]; noise: [B, out_channels, T] (the reference draws torch.randn_like(mu_q));
 * stats: [B, 2 * out_channels, T] receives proj(h) * mask -- mu_q and logs_q are its two channel halves, the views
 * torch.split returns; z_q: [B, out_channels, T].  FP32: FFMA kernels (reference parity); BF16: tcgen05 kernels
 * (BF16X3 runs as FP32, like the flow).
 */
int vsg_posterior_forward(const VsgPack* pack, const float* x, const float* mask, const float* g, const float* noise,
                          float* z_q, float* stats, int32_t B, int32_t T, int32_t precision,
                          void* workspace, size_t workspace_bytes, void* stream);

/*
 * RelativeEncoder(hidden_channels, filter_channels, n_heads, n_layers, kernel_size, p_dropout, window_size, block_length,
 * pre_ln=False, gin_channels)                                             modules/rel_transformer.py:257-320
 * -- the frame-rate stacks of the prior network (PitchPredictor, modules/visinger/predictor.py:7-19; FramePriorNetwork,
 * modules/visinger/encoder.py:58-73) and the phoneme-rate stack of the TextEncoder.  Post-LN
 * only (pre_ln = False, the reference's default and the only variant VISinger builds), block_length = None, eval mode.
 */
typedef struct VsgRelEncConfig {
  int32_t hidden_channels;
  int32_t filter_channels;
  int32_t n_heads;
  int32_t n_layers;
  int32_t kernel_size;     /* FFN conv_1 taps */
  int32_t window_size;     /* relative-position window (4) */
  int32_t gin_channels;    /* 0 = no pre_net */
} VsgRelEncConfig;

int vsg_relenc_pack_create(const VsgRelEncConfig* cfg, const VsgTensor* weights, int32_t n_weights, const char* prefix,
                           int32_t device, VsgPack** out);
size_t vsg_relenc_workspace_bytes(const VsgPack* pack, int32_t B, int32_t T, int32_t g_per_frame, int32_t precision);

/*
 * RelativeEncoder.forward(x, x_mask, g)                                    modules/rel_transformer.py:286-320
 * x, y: [B, hidden, T] (y may alias x); mask: [B, 1, T]; g: NULL, or the condition the reference feeds to pre_net:
 * [B, gin, 1] (g_per_frame == 0, e.g. the speaker embedding of the PitchPredictor) or [B, gin, T] (g_per_frame != 0,
 * e.g. the frame-level f0 of the FramePriorNetwork).  Attention: full T x T softmax per head with +-window relative-
 * position key / value embeddings, masked scores filled with -1e4 (:167), channel LayerNorm eps 1e-4.
 */
int vsg_relenc_forward(const VsgPack* pack, const float* x, const float* mask, const float* g, int32_t g_per_frame,
                       float* y, int32_t B, int32_t T, int32_t precision, void* workspace, size_t workspace_bytes,
                       void* stream);

/*
 * FramePriorNetwork(hidden, filter, n_heads, n_layers, kernel_size, gin_channels = 1, p_dropout)
 *                                                                         modules/visinger/encoder.py:58-73
 * with prior sampling fused behind it.  Weights: `prefix` + "encoder.*" (a RelativeEncoder)
 * and `prefix` + "proj.{weight,bias}".
 *   h = encoder(x, mask, g) ; stats = proj(h) * mask ; (mu_p, logs_p) = split(stats)        encoder.py:70-73
 *   z_p = (mu_p + noise * exp(logs_p)) * mask                                               models/visinger.py:107
 * x: [B, hidden, T]; g: [B, 1, T] frame-level condition (log-f0 of voiced frames) or NULL; noise, z_p: [B, hidden, T];
 * stats: [B, 2 * hidden, T] or NULL (mu_p / logs_p are its channel halves).  z_p is what vsg_infer_zp consumes.
 */
int vsg_frame_prior_pack_create(const VsgRelEncConfig* cfg, const VsgTensor* weights, int32_t n_weights, const char* prefix,
                                int32_t device, VsgPack** out);
size_t vsg_frame_prior_workspace_bytes(const VsgPack* pack, int32_t B, int32_t T, int32_t precision);
int vsg_frame_prior_forward(const VsgPack* pack, const float* x, const float* mask, const float* g, const float* noise,
                            float* stats, float* z_p, int32_t B, int32_t T, int32_t precision,
                            void* workspace, size_t workspace_bytes, void* stream);

/*
 * Length regulator + frame positions, one kernel:
 *   y[b, :, t] = enc[b, :, mel2ph[b, t] - 1]  (0 where mel2ph == 0)       expand_states, models/commons/align_ops.py:22-26
 *   y[b, :, t] += pos_table[pos_t, :],  pos_t = running count of frames with y[b, 0, t] != 0 (0 on the others)
 *                                                    models/visinger.py:79-82, modules/rel_transformer.py:78-100
 * enc: [B, H, T_ph]; mel2ph: int64 [B, T]; pos_table: [pos_rows, H] sinusoidal table (row 0 zero) or NULL; y: [B, H, T].
 */
int vsg_length_regulate(const float* enc, const int64_t* mel2ph, const float* pos_table, int32_t pos_rows, float* y,
                        int32_t B, int32_t H, int32_t T_ph, int32_t T, void* stream);

/* vsg_infer from an already sampled z_p (models/visinger.py:109-111): z_q = flow(z_p, reverse) * mask; wav = decoder(z_q). */
int vsg_infer_zp(const VsgPack* pack, const float* z_p, const float* mask, const float* g, float* wav, float* z_q_out,
                 int32_t B, int32_t T, int32_t precision, void* workspace, size_t workspace_bytes, void* stream);

/* Samples produced per latent frame (prod(upsample_rates)); 0 if the pack has no decoder. */
int32_t vsg_hop_size(const VsgPack* pack);

/* Number of kernels launched by the most recent run call on this thread (bench bookkeeping). */
int32_t vsg_last_launch_count(void);

/*
 * Per-layer parity hook (tests and tuning only; it allocates and synchronises, unlike the run calls):
 * one Conv1d on the bf16 tcgen05 kernel with its fused epilogue -- the unit the reference dispatches as
 * nn.Conv1d -> F.conv1d followed by the residual add / leaky_relu (modules/visinger/decoder.py:91-104).
 *   x_bf16: device bf16 [B, L, Cin] channels-last; w: HOST fp32 [Cout][Cin][k]; bias: HOST fp32 [Cout] or NULL;
 *   add0/add1: device bf16 [B, L, Cout] or NULL;
 *   v = (conv1d(x, w, dilation, padding = (k-1)*dilation/2) + bias + add0 + add1) * scale
 *   out_f32 (device fp32 [B, L, Cout] or NULL) = v; out_raw_bf16 = bf16(v); out_act_bf16 = bf16(leaky_relu(v, 0.1)).
 *   flags bit 0: HALO mode (one activation box per channel chunk, taps through row-shifted UMMA descriptors);
 *   bit 1: keep the weights resident in shared memory; bit 2: split-bf16 (x, add0, add1, out_raw, out_act then
 *   carry two bf16 planes per row: [B, L, 2*C] = [hi | lo]); bit 3: add0 holds leaky_relu(r, 0.1) of the residual r and
 *   is inverted in the epilogue (r = a > 0 ? a : 10 a); bits 4..: cap on 128-row blocks per tile (0 = 4).
 */
int vsg_debug_conv1d_bf16(const void* x_bf16, const float* w, const float* bias, const void* add0_bf16,
                          const void* add1_bf16, float scale, float* out_f32, void* out_raw_bf16, void* out_act_bf16,
                          int32_t B, int32_t L, int32_t Cin, int32_t Cout, int32_t k, int32_t dilation, int32_t flags,
                          int32_t device);

/*
 * Per-layer parity hook for the fused ResBlock1 pair kernel (tests and tuning only; allocates and synchronises):
 *   v = (conv1d(leaky_relu(conv1d(xa, w1, dilation d1) + b1, 0.1), w2) + b2 + add0 + add1) * scale
 * -- one (c1, c2) step of ResBlock1.forward (modules/visinger/decoder.py:91-104) with xa = leaky_relu(x) and add0 = x.
 *   xa_bf16, add0, add1, out_raw, out_act: device bf16 [B, L, C] channels-last; w1, w2: HOST fp32 [C][C][k];
 *   b1, b2: HOST fp32 [C]; out_f32: device fp32 [B, L, C] or NULL.  C in {16, 32}.
 */
int vsg_debug_pair_bf16(const void* xa_bf16, const float* w1, const float* b1, const float* w2, const float* b2,
                        const void* add0_bf16, const void* add1_bf16, float scale, float* out_f32, void* out_raw_bf16,
                        void* out_act_bf16, int32_t B, int32_t L, int32_t C, int32_t k, int32_t d1, int32_t device);

/*
 * Per-layer parity hook for the whole-ResBlock1 kernel (tests and tuning only; allocates and synchronises):
 *   for q in 0..n_pairs-1:  x = conv1d(leaky_relu(conv1d(leaky_relu(x), w[2q], dilation d_q) + b[2q]), w[2q+1]) + b[2q+1] + x
 *   out = (x [+ add1]) * scale
 * -- ResBlock1.forward (modules/visinger/decoder.py:91-104) in ONE kernel.  xa_bf16 = leaky_relu(x_0), add1, out_raw,
 * out_act: device bf16 [B, L, C] channels-last; w: HOST fp32 [2 * n_pairs][C][C][k]; b: HOST fp32 [2 * n_pairs][C];
 * dilations: HOST int32 [n_pairs]; out_f32: device fp32 [B, L, C] or NULL.  C in {16, 32, 64}.
 * max_mb caps the 128-row blocks per tile (0 = as many as fit), sets = epilogue warp sets (0 = default).
 */
int vsg_debug_resblock_bf16(const void* xa_bf16, const float* w, const float* b, int32_t n_pairs,
                            const int32_t* dilations, const void* add1_bf16, float scale, float* out_f32,
                            void* out_raw_bf16, void* out_act_bf16, int32_t B, int32_t L, int32_t C, int32_t k,
                            int32_t max_mb, int32_t sets, int32_t device);

/* Tuning aids for tools/tune_plans.py: force the tile plan (0 / -1 = automatic) and the number of timed repetitions of
 * the following vsg_debug_conv1d_bf16 calls; average kernel milliseconds of the last such call. */
int vsg_debug_set_plan(int32_t mb, int32_t cw, int32_t two_ctas, int32_t resident, int32_t reps);
float vsg_debug_last_ms(void);

/* Host-only tuning aid: print (stderr) the tile plan the launcher would choose for one convolution.
 * x3: 0 plain bf16, 1 two bf16 planes per value (split-bf16), 2 three planes (the flow at the fp32 tolerance). */
int vsg_debug_plan(int32_t Cin, int32_t Cout, int32_t k, int32_t dilation, int32_t B, int32_t L, int32_t n_adds,
                   int32_t n_outs, int32_t x3);

/* Host-only (no GPU): the tile plan of the row-packed whole-ResBlock1 kernel for one resblock shape -- C channels, k taps,
 * n_pairs (conv1, conv2) pairs with the given dilations, L time steps per utterance.  variant: 0 plain bf16, 1 split-bf16
 * (bf16x3 mode), 2 plain bf16 as two CTAs per SM.  out[8] = { 128-row blocks per tile, halo time steps per side, valid
 * time steps per tile, weight-ring stages, dynamic shared memory in bytes, mask of the convolutions in the block-Toeplitz
 * form, tiles per utterance, tensor-memory columns }.  VSG_EUNSUPPORTED: the shape runs on the per-convolution kernels. */
int vsg_debug_rp_plan(int32_t C, int32_t k, int32_t n_pairs, const int32_t* dilations, int32_t L, int32_t variant,
                      int32_t* out);

/* Process-wide tuning defaults of the tensor-core path: activation-operand feeding mode, resident weights, and the
 * L2-resident batch tiling of the decoder (target MB of one intermediate tensor per sub-batch, 0 = no tiling, <0 =
 * keep; minimum tiles per launch, <=0 = keep).
 * halo_mode is a bit field (A/B switches for bench.py and the tests; all off = the shipped configuration, halo_mode = 1):
 *   bit 0 HALO activation tiles | bit 1 bf16x3 mode without the row-packed resblock kernel | bit 2 row-packed kernel as
 *   two CTAs per SM (opt-in, measured equal) | bits 4-7 cap on 128-row blocks per tile | bit 8 no programmatic dependent launch |
 *   bit 9 no fused resblock pairs | bit 10 one launch per upsampler polyphase | bit 11 never split N = 256 tiles |
 *   bit 12 store raw AND activated resblock streams | bit 13 generic epilogue images only |
 *   bit 14 resblock chains of a stage on one stream | bit 15 whole-resblock kernel rb_tc (opt-in) | bits 16-20 cap on
 *   128-row blocks per resblock tile | bits 21-23 resblock epilogue warp sets (0 = 4) | bit 24 no row-packed resblock
 *   kernel | bit 25 row-packed kernel at C = 64 too | bit 26 no block-Toeplitz form | bit 27 no specialised images for
 *   the last conv2 of a resblock (running-sum epilogues) | bits 28-29 conv_post on the CUDA-core kernels (1 register
 *   window, 2 shared-memory window) instead of tcgen05 | bit 30 row-packed kernel with two epilogue warp sets per block |
 *   bit 31 flow with separate res / skip launches and one cond_layer GEMV per flow. */
int vsg_set_tc_options(int32_t halo_mode, int32_t w_resident, int32_t l2_tensor_mb, int32_t min_tiles);

#ifdef __cplusplus
}
#endif
#endif /* VISINGER_B200_H_ */

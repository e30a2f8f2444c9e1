// Internal declarations shared by the run drivers and the C-ABI entry points.
#pragma once
#include "vsg_common.cuh"

namespace vsg {

int launch_cond(const float* W, const float* bias, const float* g, float* out, int O, int I, int B, cudaStream_t st);
int prior_sample(const float* mu, const float* logs, const float* noise, const float* mask, float* z, int B, int C,
                 int T, cudaStream_t st);
int mask_mul(const float* x, const float* mask, float* y, int B, int C, int T, cudaStream_t st);

// output stage (output.cu): peak-normalise -> x 32767 -> int16 (utils/audio/io.py:8-14)
int wav_to_int16(const float* wav, const int32_t* lengths, int16_t* pcm, float* peak, int B, int L, int norm, cudaStream_t st);

// fp32 parity mode (run_f32.cu)
size_t flow_ws_bytes_f32(const VsgPack* P, int B, int T);
size_t dec_ws_bytes_f32(const VsgPack* P, int B, int T);
int flow_forward_f32(const VsgPack* P, const float* x, const float* mask, const float* g, float* y, int B, int T,
                     int reverse, Workspace& ws, cudaStream_t st);
int generator_forward_f32(const VsgPack* P, const float* z, const float* g, float* wav, int B, int T, Workspace& ws,
                          cudaStream_t st);

// bf16 tensor-core mode (run_tc.cu)
size_t flow_ws_bytes_tc(const VsgPack* P, int B, int T, int planes = 1);
size_t dec_ws_bytes_tc(const VsgPack* P, int B, int T, bool x3);
int flow_forward_tc(const VsgPack* P, const float* x, const float* mask, const float* g, float* y, int B, int T,
                    int reverse, Workspace& ws, cudaStream_t st, int planes = 1, const float* out_mask = nullptr,   // out_mask: y * mask
                    const float* ps_logs = nullptr, const float* ps_noise = nullptr);   // x = mu_p: prior sampling at the entry   // planes = 3: fp32 tolerance (bf16x3 mode)
// VSG_X3_FLOW_FFMA=1 in the environment: the bf16x3 mode runs the flow on the fp32 CUDA-core kernels (A/B measurements)
bool x3_flow_on_tensor_cores();
int generator_forward_tc(const VsgPack* P, const float* z, const float* g, float* wav, int B, int T, Workspace& ws,
                         cudaStream_t st, bool x3);


// PosteriorEncoder (f32: run_f32.cu, bf16: run_tc.cu)
size_t posterior_ws_bytes_f32(const VsgPack* P, int B, int T);
size_t posterior_ws_bytes_tc(const VsgPack* P, int B, int T);
int posterior_forward_f32(const VsgPack* P, const float* x, const float* mask, const float* g, const float* noise,
                          float* z, float* stats, int B, int T, Workspace& ws, cudaStream_t st);
int posterior_forward_tc(const VsgPack* P, const float* x, const float* mask, const float* g, const float* noise,
                         float* z, float* stats, int B, int T, Workspace& ws, cudaStream_t st);


// y [B, Cout, T] = Conv1d(x [B, Cin, T]) with the fp32 kernel, stride 1, 'same' padding (run_f32.cu)
int conv_f32_plain(const ConvW32& w, const float* x, int B, int T, float* y, cudaStream_t st);

// RelativeEncoder (f32: run_f32.cu, bf16: run_tc.cu)
size_t relenc_ws_bytes_tc(const VsgPack* P, int B, int T, int g_t);
int relenc_forward_tc(const VsgPack* P, const float* x, const float* mask, const float* g, int g_t, float* y, int B, int T,
                      Workspace& ws, cudaStream_t st);
size_t relenc_ws_bytes_f32(const VsgPack* P, int B, int T, int g_t);
int relenc_forward_f32(const VsgPack* P, const float* x, const float* mask, const float* g, int g_t, float* y, int B, int T,
                       Workspace& ws, cudaStream_t st);


// FramePriorNetwork + prior sampling (encoder -> proj -> sample); stats [B, 2H, T] may be null
size_t frame_prior_ws_bytes(const VsgPack* P, int B, int T, int precision);
int frame_prior_forward_f32(const VsgPack* P, const float* x, const float* mask, const float* g, const float* noise,
                            float* stats, float* z, int B, int T, Workspace& ws, cudaStream_t st);
int frame_prior_forward_tc(const VsgPack* P, const float* x, const float* mask, const float* g, const float* noise,
                           float* stats, float* z, int B, int T, Workspace& ws, cudaStream_t st);
int length_regulate(const float* enc, const long long* mel2ph, const float* table, int table_rows, float* y, int B, int H,
                    int T_ph, int T, cudaStream_t st);

}  // namespace vsg

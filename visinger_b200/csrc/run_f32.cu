// fp32 parity-mode drivers: sequence the fused FFMA kernels of kernels_f32.cuh over one batch.
#include "vsg_common.cuh"
#include "kernels_f32.cuh"
#include "relenc_f32.cuh"
#include "run.cuh"

namespace vsg {

namespace {

int launch_conv_f32(const ConvF32& p, int B, cudaStream_t st) {
  if (p.ktaps < 1) return fail(VSG_EUNSUPPORTED, "convolution with no taps");
  const int halo = (p.ktaps - 1) * p.dil;
  auto smem_for = [&](int co_tile, int q_tile) {
    return (size_t)F32_CI_CHUNK * ((q_tile + halo) + p.ktaps * co_tile) * sizeof(float);
  };
#define VSG_F32_LAUNCH(WC, WQ)                                                                              \
  do {                                                                                                      \
    const size_t sm = smem_for(8 * WC, 256 * WQ);                                                           \
    if (sm > 200 * 1024) return fail(VSG_EUNSUPPORTED, "conv window too large for shared memory (%zu B)", sm); \
    if (sm > 48 * 1024)                                                                                     \
      VSG_CUDA_TRY(cudaFuncSetAttribute(conv_f32_kernel<WC, WQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
    dim3 grid((p.Lq + 256 * WQ - 1) / (256 * WQ), (p.Cout + 8 * WC - 1) / (8 * WC), B);                     \
    conv_f32_kernel<WC, WQ><<<grid, 256, sm, st>>>(p);                                                      \
    VSG_LAUNCH_CHECK("conv_f32_kernel");                                                                    \
  } while (0)
  if (p.Lq <= 0) return VSG_OK;
  if (p.Cout <= 16) VSG_F32_LAUNCH(2, 4);
  else if (p.Cout <= 32) VSG_F32_LAUNCH(4, 2);
  else VSG_F32_LAUNCH(8, 1);
#undef VSG_F32_LAUNCH
  return VSG_OK;
}

ConvF32 base_conv(const ConvW32& w, const float* x, long long x_bs, int x_cs, int Lin, int in_off0, int dil) {
  ConvF32 p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.x_bs = x_bs; p.x_cs = x_cs; p.Cin = w.Cin; p.Lin = Lin;
  p.in_off0 = in_off0; p.dil = dil; p.ktaps = w.ktaps;
  p.slope = 0.1f;   // LRELU_SLOPE, modules/visinger/decoder.py:10
  p.w = w.w; p.CoutP = w.CoutP; p.bias = w.bias; p.Cout = w.Cout;
  p.Lq = Lin; p.out_stride = 1; p.out_phase = 0; p.Lout = Lin;
  p.epi = EPI_LINEAR; p.acc_mode = ACC_SET; p.acc_div = 1.f;
  return p;
}

}  // namespace

int launch_cond(const float* W, const float* bias, const float* g, float* out, int O, int I, int B, cudaStream_t st) {
  const long long warps = (long long)O * B;
  const int threads = 256;
  const int blocks = (int)((warps * 32 + threads - 1) / threads);
  cond_gemv_kernel<<<blocks, threads, 0, st>>>(W, bias, g, out, O, I, B);
  VSG_LAUNCH_CHECK("cond_gemv_kernel");
  return VSG_OK;
}

size_t flow_ws_bytes_f32(const VsgPack* P, int B, int T) {
  const VsgConfig& c = P->cfg;
  size_t n = 0;
  n += align256((size_t)B * c.flow_n_flows * 2 * c.flow_hidden * c.flow_n_layers * sizeof(float));  // cond
  n += 3 * align256((size_t)B * c.flow_hidden * T * sizeof(float));                                  // h, acts, out
  return n;
}

// ResidualCouplingBlock.forward, modules/visinger/flow.py:33-40, with the Flips folded away.
int flow_forward_f32(const VsgPack* P, const float* x, const float* mask, const float* g, float* y, int B, int T,
                     int reverse, Workspace& ws, cudaStream_t st) {
  const VsgConfig& c = P->cfg;
  const int C = c.flow_channels, H = c.flow_hidden, NL = c.flow_n_layers, NF = c.flow_n_flows, half = C / 2;
  const int condO = 2 * H * NL;
  float* cond = ws.take<float>((size_t)B * NF * condO);
  float* h = ws.take<float>((size_t)B * H * T);
  float* acts = ws.take<float>((size_t)B * H * T);
  float* out = ws.take<float>((size_t)B * H * T);
  if (ws.overflow) return fail(VSG_ENOMEM, "flow workspace too small: need %zu bytes", ws.off);
  if (y != x) VSG_CUDA_TRY(cudaMemcpyAsync(y, x, (size_t)B * C * T * sizeof(float), cudaMemcpyDeviceToDevice, st));
  if (c.flow_gin > 0) {
    if (!g) return fail(VSG_EINVAL, "flow was built with gin_channels=%d but g is NULL", c.flow_gin);
    for (int f = 0; f < NF; ++f)
      VSG_TRY(launch_cond(P->flow_layers[f].cond_w, P->flow_layers[f].cond_b, g, cond + (size_t)f * condO * B, condO,
                          c.flow_gin, B, st));
  }
  const long long CT = (long long)C * T, HT = (long long)H * T;
  for (int step = 0; step < NF; ++step) {
    const int f = reverse ? NF - 1 - step : step;
    const int flipped = reverse ? (NF - f) & 1 : f & 1;   // parity of Flips applied so far
    const FlowLayer& fl = P->flow_layers[f];
    const float* x0 = y + (flipped ? (long long)half * T : 0);   // conditioning half (never modified here)
    float* x1 = y + (flipped ? 0 : (long long)half * T);         // half that is updated in place
    // h = pre(x0) * mask                                           flow.py:68
    {
      ConvF32 p = base_conv(fl.pre[flipped], x0, CT, T, T, 0, 1);
      p.y = h; p.y_bs = HT; p.y_cs = T; p.mask = mask; p.mask_bs = T;
      VSG_TRY(launch_conv_f32(p, B, st));
    }
    // WaveNet                                                       encoder.py:174-195
    int dil = 1;
    for (int i = 0; i < NL; ++i) {
      const int K = c.flow_kernel_size;
      const int pad = (K * dil - dil) / 2;
      {
        ConvF32 p = base_conv(fl.in_layers[i], h, HT, T, T, -pad, dil);
        p.epi = EPI_GATE;
        if (c.flow_gin > 0) { p.bcond = cond + (size_t)f * condO * B + (size_t)i * 2 * H; p.bcond_bs = condO; }
        p.y = acts; p.y_bs = HT; p.y_cs = T;
        VSG_TRY(launch_conv_f32(p, B, st));
      }
      {
        ConvF32 p = base_conv(fl.res_skip[i], acts, HT, T, T, 0, 1);
        p.epi = EPI_RES_SKIP;
        const bool last = (i == NL - 1);
        p.rs_split = last ? 0 : H;
        p.y = h; p.y_bs = HT; p.y_cs = T;
        p.y2 = out; p.y2_bs = HT; p.y2_cs = T;
        p.y2_first = (i == 0); p.y2_mask = last;
        p.mask = mask; p.mask_bs = T;
        VSG_TRY(launch_conv_f32(p, B, st));
      }
      dil *= c.flow_dilation_rate;
    }
    // m = post(h) * mask ; x1 = (x1 - m) * mask | m + x1 * mask     flow.py:70,78,83
    {
      ConvF32 p = base_conv(fl.post[flipped], out, HT, T, T, 0, 1);
      p.epi = EPI_COUPLE; p.couple_sign = reverse ? -1 : 1;
      p.y = x1; p.y_bs = CT; p.y_cs = T; p.mask = mask; p.mask_bs = T;
      VSG_TRY(launch_conv_f32(p, B, st));
    }
  }
  if (NF & 1) {   // an odd number of Flips does not cancel
    const long long n = (long long)B * half * T;
    flip_channels_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(y, C, T, n);
    VSG_LAUNCH_CHECK("flip_channels_kernel");
  }
  return VSG_OK;
}

size_t posterior_ws_bytes_f32(const VsgPack* P, int B, int T) {
  const EncPack& e = P->enc;
  return align256((size_t)B * 2 * e.hidden * e.n_layers * sizeof(float)) + 3 * align256((size_t)B * e.hidden * T * sizeof(float));
}

// PosteriorEncoder.forward, modules/visinger/encoder.py:92-98.
int posterior_forward_f32(const VsgPack* P, const float* x, const float* mask, const float* g, const float* noise,
                          float* z, float* stats, int B, int T, Workspace& ws, cudaStream_t st) {
  const EncPack& e = P->enc;
  const int H = e.hidden, NL = e.n_layers, K = e.kernel, Cin = e.in_channels, Co = e.out_channels;
  const int condO = 2 * H * NL;
  float* cond = ws.take<float>((size_t)B * condO);
  float* h = ws.take<float>((size_t)B * H * T);
  float* acts = ws.take<float>((size_t)B * H * T);
  float* out = ws.take<float>((size_t)B * H * T);
  if (ws.overflow) return fail(VSG_ENOMEM, "posterior workspace too small: need %zu bytes", ws.off);
  if (e.gin > 0) {
    if (!g) return fail(VSG_EINVAL, "posterior encoder was built with gin_channels=%d but g is NULL", e.gin);
    VSG_TRY(launch_cond(e.wn.cond_w, e.wn.cond_b, g, cond, condO, e.gin, B, st));
  }
  const long long HT = (long long)H * T;
  {  // h = pre(x) * mask                                            encoder.py:93
    ConvF32 p = base_conv(e.pre, x, (long long)Cin * T, T, T, 0, 1);
    p.y = h; p.y_bs = HT; p.y_cs = T; p.mask = mask; p.mask_bs = T;
    VSG_TRY(launch_conv_f32(p, B, st));
  }
  int dil = 1;
  for (int i = 0; i < NL; ++i) {   // WaveNet                        encoder.py:174-195
    const int pad = (K * dil - dil) / 2;
    {
      ConvF32 p = base_conv(e.wn.in_layers[i], h, HT, T, T, -pad, dil);
      p.epi = EPI_GATE;
      if (e.gin > 0) { p.bcond = cond + (size_t)i * 2 * H; p.bcond_bs = condO; }
      p.y = acts; p.y_bs = HT; p.y_cs = T;
      VSG_TRY(launch_conv_f32(p, B, st));
    }
    {
      ConvF32 p = base_conv(e.wn.res_skip[i], acts, HT, T, T, 0, 1);
      p.epi = EPI_RES_SKIP;
      const bool last = (i == NL - 1);
      p.rs_split = last ? 0 : H;
      p.y = h; p.y_bs = HT; p.y_cs = T;
      p.y2 = out; p.y2_bs = HT; p.y2_cs = T;
      p.y2_first = (i == 0); p.y2_mask = last;
      p.mask = mask; p.mask_bs = T;
      VSG_TRY(launch_conv_f32(p, B, st));
    }
    dil *= e.dil_rate;
  }
  {  // stats = proj(h) * mask                                        encoder.py:95
    ConvF32 p = base_conv(e.proj, out, HT, T, T, 0, 1);
    p.y = stats; p.y_bs = (long long)2 * Co * T; p.y_cs = T; p.mask = mask; p.mask_bs = T;
    VSG_TRY(launch_conv_f32(p, B, st));
  }
  const long long n = (long long)B * Co * T;   // z = (mu + noise * exp(logs)) * mask   encoder.py:96-97
  posterior_sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(stats, noise, mask, z, Co, T, n);
  VSG_LAUNCH_CHECK("posterior_sample_kernel");
  return VSG_OK;
}

int conv_f32_plain(const ConvW32& w, const float* x, int B, int T, float* y, cudaStream_t st) {
  ConvF32 p = base_conv(w, x, (long long)w.Cin * T, T, T, -(w.ktaps / 2), 1);
  p.y = y; p.y_bs = (long long)w.Cout * T; p.y_cs = T;
  return launch_conv_f32(p, B, st);
}

size_t relenc_ws_bytes_f32(const VsgPack* P, int B, int T, int g_t) {
  const RelEncPack& e = P->relenc;
  size_t n = align256((size_t)B * 3 * e.hidden * T * sizeof(float)) + align256((size_t)B * e.hidden * T * sizeof(float)) +
             align256((size_t)B * e.filter * T * sizeof(float));
  n += align256((size_t)B * e.hidden * (g_t ? T : 1) * sizeof(float));
  return n;
}

namespace {
template <int DK>
int launch_attention_f32(const float* qkv, const float* mask, const float* ek, const float* ev, float* o, int B, int n_heads,
                         int T, int w, cudaStream_t st) {
  const size_t sm = relenc_attention_f32_smem(DK, w);
  if (sm > 200 * 1024) return fail(VSG_EUNSUPPORTED, "attention tile does not fit shared memory (%zu B)", sm);
  VSG_CUDA_TRY(cudaFuncSetAttribute(relenc_attention_f32_kernel<DK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  dim3 grid((T + kAttTile - 1) / kAttTile, n_heads, B);
  relenc_attention_f32_kernel<DK><<<grid, 256, sm, st>>>(qkv, mask, ek, ev, o, n_heads, T, w);
  VSG_LAUNCH_CHECK("relenc_attention_f32_kernel");
  return VSG_OK;
}
}  // namespace

// RelativeEncoder.forward, modules/rel_transformer.py:286-320 (post-LN; dropout is the identity in eval mode).
// g: NULL, [B, gin] (g_t == 0: one condition vector per utterance, e.g. the speaker embedding) or [B, gin, T] (g_t != 0).
int relenc_forward_f32(const VsgPack* P, const float* x, const float* mask, const float* g, int g_t, float* y, int B, int T,
                       Workspace& ws, cudaStream_t st) {
  const RelEncPack& e = P->relenc;
  const int H = e.hidden, F = e.filter, NL = e.n_layers, K = e.kernel, dk = H / e.n_heads;
  float* qkv = ws.take<float>((size_t)B * 3 * H * T);
  float* o = ws.take<float>((size_t)B * H * T);
  float* f = ws.take<float>((size_t)B * F * T);
  float* gp = ws.take<float>((size_t)B * H * (g_t ? T : 1));
  if (ws.overflow) return fail(VSG_ENOMEM, "encoder workspace too small: need %zu bytes", ws.off);
  const long long HT = (long long)H * T, n = (long long)B * H * T;
  if (g && e.gin <= 0) return fail(VSG_EINVAL, "this encoder has no pre_net (gin_channels is None) but g was given");
  const float* gadd = nullptr;
  if (g) {   // g = pre_net(g)                                        :289-290
    if (g_t) {
      ConvF32 p = base_conv(e.pre_net, g, (long long)e.gin * T, T, T, 0, 1);
      p.y = gp; p.y_bs = HT; p.y_cs = T;
      VSG_TRY(launch_conv_f32(p, B, st));
    } else {
      VSG_TRY(launch_cond(e.pre_w, e.pre_b, g, gp, H, e.gin, B, st));
    }
    gadd = gp;
  }
  if (y != x) VSG_CUDA_TRY(cudaMemcpyAsync(y, x, (size_t)n * sizeof(float), cudaMemcpyDeviceToDevice, st));
  relenc_addg_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(y, gadd, g_t, mask, H, T, n);   // x = (x + g) * mask
  VSG_LAUNCH_CHECK("relenc_addg_mask_kernel");
  const unsigned ln_blocks = (unsigned)(((long long)B * T + 127) / 128);
  for (int i = 0; i < NL; ++i) {
    const RelEncLayer& L = e.layers[i];
    {  // q | k | v = conv_{q,k,v}(x)                                :124-126
      ConvF32 p = base_conv(L.qkv, y, HT, T, T, 0, 1);
      p.y = qkv; p.y_bs = 3 * HT; p.y_cs = T;
      VSG_TRY(launch_conv_f32(p, B, st));
    }
    switch (dk) {                                                   // attention :137-177
      case 16: VSG_TRY(launch_attention_f32<16>(qkv, mask, L.ek, L.ev, o, B, e.n_heads, T, e.window, st)); break;
      case 32: VSG_TRY(launch_attention_f32<32>(qkv, mask, L.ek, L.ev, o, B, e.n_heads, T, e.window, st)); break;
      case 48: VSG_TRY(launch_attention_f32<48>(qkv, mask, L.ek, L.ev, o, B, e.n_heads, T, e.window, st)); break;
      case 64: VSG_TRY(launch_attention_f32<64>(qkv, mask, L.ek, L.ev, o, B, e.n_heads, T, e.window, st)); break;
      case 96: VSG_TRY(launch_attention_f32<96>(qkv, mask, L.ek, L.ev, o, B, e.n_heads, T, e.window, st)); break;
      case 128: VSG_TRY(launch_attention_f32<128>(qkv, mask, L.ek, L.ev, o, B, e.n_heads, T, e.window, st)); break;
      default: return fail(VSG_EUNSUPPORTED, "attention head width %d (supported: 16, 32, 48, 64, 96, 128)", dk);
    }
    {  // x = x + conv_o(attn)                                        :130, :301
      ConvF32 p = base_conv(L.o, o, HT, T, T, 0, 1);
      p.res = y; p.res_bs = HT; p.res_cs = T;
      p.y = y; p.y_bs = HT; p.y_cs = T;
      VSG_TRY(launch_conv_f32(p, B, st));
    }
    relenc_layernorm_kernel<<<ln_blocks, 128, 0, st>>>(y, L.g1, L.b1, nullptr, 0, mask, 1e-4f, H, T, B);   // :303
    VSG_LAUNCH_CHECK("relenc_layernorm_kernel");
    {  // FFN: conv_1(x * mask) -> relu -> conv_2(. * mask)            :337-345
      ConvF32 p = base_conv(L.ffn1, y, HT, T, T, -(K / 2), 1);
      p.y = f; p.y_bs = (long long)F * T; p.y_cs = T; p.mask = mask; p.mask_bs = T;   // relu(h) * mask == relu(h * mask)
      VSG_TRY(launch_conv_f32(p, B, st));
    }
    {
      ConvF32 p = base_conv(L.ffn2, f, (long long)F * T, T, T, 0, 1);
      p.pre_lrelu = 1; p.slope = 0.f;                               // relu on load
      p.res = y; p.res_bs = HT; p.res_cs = T;
      p.y = y; p.y_bs = HT; p.y_cs = T;
      VSG_TRY(launch_conv_f32(p, B, st));
    }
    // x = LayerNorm(x); then the next layer's `x = (x + g) * mask` (:293-295) or the final `x * mask` (:318)
    relenc_layernorm_kernel<<<ln_blocks, 128, 0, st>>>(y, L.g2, L.b2, (i + 1 < NL) ? gadd : nullptr, g_t, mask, 1e-4f, H, T, B);
    VSG_LAUNCH_CHECK("relenc_layernorm_kernel");
  }
  return VSG_OK;
}

// FramePriorNetwork.forward (modules/visinger/encoder.py:67-73) + prior sampling (models/visinger.py:107), fp32 mode.
int frame_prior_forward_f32(const VsgPack* P, const float* x, const float* mask, const float* g, const float* noise,
                            float* stats, float* z, int B, int T, Workspace& ws, cudaStream_t st) {
  const RelEncPack& e = P->relenc;
  const int H = e.hidden;
  float* h = ws.take<float>((size_t)B * H * T);
  if (ws.overflow) return fail(VSG_ENOMEM, "frame prior workspace too small: need %zu bytes", ws.off);
  VSG_TRY(relenc_forward_f32(P, x, mask, g, g ? 1 : 0, h, B, T, ws, st));
  const size_t sm = (size_t)H * 32 * sizeof(float);
  dim3 grid((T + 31) / 32, B);
  prior_head_f32_kernel<<<grid, 256, sm, st>>>(h, e.proj_w, e.proj_b, noise, mask, stats, z, H, T);
  VSG_LAUNCH_CHECK("prior_head_f32_kernel");
  return VSG_OK;
}

int length_regulate(const float* enc, const long long* mel2ph, const float* table, int table_rows, float* y, int B, int H,
                    int T_ph, int T, cudaStream_t st) {
  const size_t sm = (size_t)2 * T * sizeof(int);
  if (sm > 96 * 1024) return fail(VSG_EUNSUPPORTED, "sequence too long for the length regulator (%d frames)", T);
  if (sm > 48 * 1024)
    VSG_CUDA_TRY(cudaFuncSetAttribute(length_regulate_pos_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
  length_regulate_pos_kernel<<<B, 512, sm, st>>>(enc, mel2ph, table, table_rows, y, H, T_ph, T);
  VSG_LAUNCH_CHECK("length_regulate_pos_kernel");
  return VSG_OK;
}

static size_t dec_max_elems(const VsgPack* P, int B, int T) {
  size_t m = (size_t)B * P->cfg.dec_upsample_initial_channel * T;
  long long L = T;
  for (const UpStage& st : P->ups) {
    L *= st.rate;
    m = std::max(m, (size_t)B * st.Cout * (size_t)L);
  }
  return m;
}

size_t dec_ws_bytes_f32(const VsgPack* P, int B, int T) {
  return align256((size_t)B * P->cfg.dec_upsample_initial_channel * sizeof(float)) +
         4 * align256(dec_max_elems(P, B, T) * sizeof(float));
}

// Generator.forward, modules/visinger/decoder.py:40-59.
int generator_forward_f32(const VsgPack* P, const float* z, const float* g, float* wav, int B, int T, Workspace& ws,
                          cudaStream_t st) {
  const VsgConfig& c = P->cfg;
  const int UIC = c.dec_upsample_initial_channel, NK = c.dec_n_kernels;
  const size_t E = dec_max_elems(P, B, T);
  float* cond = ws.take<float>((size_t)B * UIC);
  float* bufS = ws.take<float>(E);   // stage input / running resblock sum
  float* bufU = ws.take<float>(E);   // upsampled x shared by the NK resblocks
  float* bufR = ws.take<float>(E);   // resblock running x
  float* bufT = ws.take<float>(E);   // conv1 output
  if (ws.overflow) return fail(VSG_ENOMEM, "generator workspace too small: need %zu bytes", ws.off);
  if (c.dec_gin > 0) {
    if (!g) return fail(VSG_EINVAL, "generator was built with gin_channels=%d but g is NULL", c.dec_gin);
    VSG_TRY(launch_cond(P->dec_cond_w, P->dec_cond_b, g, cond, UIC, c.dec_gin, B, st));
  }
  // x = conv_pre(z) + cond(g)                                        decoder.py:41-43
  {
    ConvF32 p = base_conv(P->conv_pre, z, (long long)c.dec_initial_channel * T, T, T, -3, 1);
    if (c.dec_gin > 0) { p.bcond = cond; p.bcond_bs = UIC; }
    p.y = bufS; p.y_bs = (long long)UIC * T; p.y_cs = T;
    VSG_TRY(launch_conv_f32(p, B, st));
  }
  int L = T, ch = UIC;
  for (int i = 0; i < c.dec_n_ups; ++i) {
    const UpStage& us = P->ups[i];
    const int Lout = L * us.rate;
    // x = ups[i](leaky_relu(x))  as `rate` polyphase convolutions    decoder.py:45-46
    for (int r = 0; r < us.rate; ++r) {
      ConvF32 p = base_conv(us.phases[r].f32, bufS, (long long)ch * L, L, L, us.phases[r].in_off0, 1);
      p.pre_lrelu = 1;
      p.Lq = (Lout - r + us.rate - 1) / us.rate; p.out_stride = us.rate; p.out_phase = r; p.Lout = Lout;
      p.y = bufU; p.y_bs = (long long)us.Cout * Lout; p.y_cs = Lout;
      VSG_TRY(launch_conv_f32(p, B, st));
    }
    ch = us.Cout; L = Lout;
    const long long bs = (long long)ch * L;
    // xs = sum_j resblock_j(x) ; x = xs / NK                           decoder.py:47-54
    for (int j = 0; j < NK; ++j) {
      const ResBlockPack& rb = us.blocks[j];
      const int nd = (int)rb.dilations.size();
      const int acc_mode = (j == 0) ? ACC_SET : (j == NK - 1 ? ACC_ADD_DIV : ACC_ADD);
      const float* cur = bufU;
      for (int q = 0; q < nd; ++q) {
        const bool last = (q == nd - 1);
        const int d = rb.dilations[q], k = rb.kernel;
        if (c.dec_resblock == 1) {   // ResBlock1, decoder.py:91-104
          ConvF32 p1 = base_conv(rb.c1[q], cur, bs, L, L, -((k * d - d) / 2), d);
          p1.pre_lrelu = 1; p1.y = bufT; p1.y_bs = bs; p1.y_cs = L;
          VSG_TRY(launch_conv_f32(p1, B, st));
          ConvF32 p2 = base_conv(rb.c2[q], bufT, bs, L, L, -((k - 1) / 2), 1);
          p2.pre_lrelu = 1; p2.res = cur; p2.res_bs = bs; p2.res_cs = L;
          p2.y = last ? bufS : bufR; p2.y_bs = bs; p2.y_cs = L;
          if (last) { p2.acc_mode = acc_mode; p2.acc_div = (float)NK; }
          VSG_TRY(launch_conv_f32(p2, B, st));
          cur = bufR;
        } else {                      // ResBlock2, decoder.py:124-133
          float* dst = last ? bufS : (cur == bufR ? bufT : bufR);
          ConvF32 p1 = base_conv(rb.c1[q], cur, bs, L, L, -((k * d - d) / 2), d);
          p1.pre_lrelu = 1; p1.res = cur; p1.res_bs = bs; p1.res_cs = L;
          p1.y = dst; p1.y_bs = bs; p1.y_cs = L;
          if (last) { p1.acc_mode = acc_mode; p1.acc_div = (float)NK; }
          VSG_TRY(launch_conv_f32(p1, B, st));
          cur = dst;
        }
      }
    }
    if (NK == 1) { /* xs / 1 == xs */ }
  }
  // wav = tanh(conv_post(leaky_relu(x)))                               decoder.py:55-57
  {
    dim3 grid((L + 255) / 256, B);
    conv_post_f32_kernel<<<grid, 256, (size_t)ch * P->conv_post_k * sizeof(float), st>>>(bufS, P->conv_post_w, wav, ch,
                                                                                          L, P->conv_post_k, 0.1f);
    VSG_LAUNCH_CHECK("conv_post_f32_kernel");
  }
  return VSG_OK;
}

int prior_sample(const float* mu, const float* logs, const float* noise, const float* mask, float* z, int B, int C,
                 int T, cudaStream_t st) {
  const long long n = (long long)B * C * T;
  if (n == 0) return VSG_OK;
  prior_sample_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(mu, logs, noise, mask, z, C, T, n);
  VSG_LAUNCH_CHECK("prior_sample_kernel");
  return VSG_OK;
}

int mask_mul(const float* x, const float* mask, float* y, int B, int C, int T, cudaStream_t st) {
  const long long n = (long long)B * C * T;
  if (n == 0) return VSG_OK;
  mask_mul_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, mask, y, C, T, n);
  VSG_LAUNCH_CHECK("mask_mul_kernel");
  return VSG_OK;
}

}  // namespace vsg

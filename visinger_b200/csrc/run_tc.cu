// placeholder until the tcgen05 path lands
#include "vsg_common.cuh"
#include "run.cuh"
#include "pack_tc.cuh"
namespace vsg {
int pack_conv_tc(VsgPack*, const std::vector<float>&, const std::vector<float>&, int, int, int, ConvWTC*) { return VSG_OK; }
size_t flow_ws_bytes_tc(const VsgPack* P, int B, int T) { return flow_ws_bytes_f32(P, B, T); }
size_t dec_ws_bytes_tc(const VsgPack* P, int B, int T) { return dec_ws_bytes_f32(P, B, T); }
int flow_forward_tc(const VsgPack*, const float*, const float*, const float*, float*, int, int, int, Workspace&, cudaStream_t) { return fail(VSG_EUNSUPPORTED, "bf16 path not built"); }
int generator_forward_tc(const VsgPack*, const float*, const float*, float*, int, int, Workspace&, cudaStream_t) { return fail(VSG_EUNSUPPORTED, "bf16 path not built"); }
}

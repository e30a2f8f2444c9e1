// Weight pre-pack for the bf16 tensor-core path (implemented in conv_tc.cu).
#pragma once
#include "vsg_common.cuh"

namespace vsg {
// Conv1d-style weight W[co][ci][j] (+ bias[co]) -> bf16 K-major pack and its TMA tensor map.
int pack_conv_tc(VsgPack* P, const std::vector<float>& W, const std::vector<float>& b, int Cout, int Cin, int k,
                 ConvWTC* out);
}  // namespace vsg

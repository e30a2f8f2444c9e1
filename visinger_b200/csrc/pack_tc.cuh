// Weight pre-pack for the bf16 tensor-core path (implemented in conv_tc.cu).
#pragma once
#include "vsg_common.cuh"

namespace vsg {
// Conv1d-style weight W[co][ci][j] (+ bias[co]) -> bf16 K-major pack and its TMA tensor map.
// x3: split-bf16 pack [W_hi | W_lo] for the fp32-tolerance tensor-core mode.
int pack_conv_tc(VsgPack* P, const std::vector<float>& W, const std::vector<float>& b, int Cout, int Cin, int k,
                 ConvWTC* out, bool x3 = false);
}  // namespace vsg

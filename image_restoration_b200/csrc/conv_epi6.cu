// Kernels of the implicit-GEMM convolution with epilogue profile 6 (ConvUpLayer folded: phases as column blocks, bilinear
// skip residual per phase, border-ring correction; see epilogue_upfold).
#include "conv_common.cuh"

namespace b200ir {
template int launch_conv_variant<6>(const ConvParams&, int, bool, int, int, int, cudaStream_t);
}  // namespace b200ir

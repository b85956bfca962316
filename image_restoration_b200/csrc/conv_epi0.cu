// Kernels of the implicit-GEMM convolution with epilogue profile 0 (compile-time flags, see epi_profile_flags).
#include "conv_common.cuh"

namespace b200ir {
template int launch_conv_variant<0>(const ConvParams&, int, bool, int, int, int, cudaStream_t);
}  // namespace b200ir

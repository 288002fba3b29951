// placeholder until the tcgen05 kernels land
#include "nn_common.cuh"
int nn_umma_conv_fwd(const nn_conv_fwd_args*, int, cudaStream_t) { return nn_fail("tcgen05 path not built%s", ""); }
int nn_umma_conv_dgrad(const nn_conv_dgrad_args*, int, cudaStream_t) { return nn_fail("tcgen05 path not built%s", ""); }
int nn_umma_conv_wgrad(const nn_conv_wgrad_args*, int, cudaStream_t) { return nn_fail("tcgen05 path not built%s", ""); }
int64_t nn_umma_fwd_workspace(const nn_conv_geom*, int) { return 0; }
int64_t nn_umma_wgrad_workspace(const nn_conv_geom*, int, int) { return 0; }
bool nn_umma_supports(const nn_conv_geom*, int) { return false; }

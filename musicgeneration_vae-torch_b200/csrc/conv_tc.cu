// conv_tc.cu -- tcgen05 / TMEM / TMA implicit-GEMM kernels (placeholder until the kernels land).
#include "common.cuh"
namespace bvae {
int conv_tc_eligible(const bvae_conv_desc*) { return 0; }
int conv_tc_launch(const bvae_conv_desc*, cudaStream_t) { set_error("tcgen05 conv not built"); return BVAE_ERR_UNSUPPORTED; }
int wgrad_tc_eligible(const bvae_wgrad_desc*) { return 0; }
int wgrad_tc_launch(const bvae_wgrad_desc*, cudaStream_t) { set_error("tcgen05 wgrad not built"); return BVAE_ERR_UNSUPPORTED; }
}

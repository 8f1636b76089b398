// conv_api.cu -- validation + dispatch of the contraction entry points (SIMT checker vs tcgen05 kernels).
#include "common.cuh"

namespace bvae {
int conv_simt_launch(const bvae_conv_desc* d, cudaStream_t stream);
int wgrad_simt_launch(const bvae_wgrad_desc* d, cudaStream_t stream);
int conv_tc_eligible(const bvae_conv_desc* d);
int conv_tc_launch(const bvae_conv_desc* d, cudaStream_t stream);
int conv_tc_stats_ok(const bvae_conv_desc* d);
int wgrad_tc_eligible(const bvae_wgrad_desc* d);
int wgrad_tc_launch(const bvae_wgrad_desc* d, cudaStream_t stream);
int stem_fwd_eligible(const bvae_conv_desc* d);
int stem_fwd_launch(const bvae_conv_desc* d, cudaStream_t stream);
int stem_wgrad_eligible(const bvae_wgrad_desc* d);
int stem_wgrad_launch(const bvae_wgrad_desc* d, cudaStream_t stream);
}  // namespace bvae

using namespace bvae;

extern "C" int bvae_conv_gemm(const bvae_conv_desc* d, int impl, void* stream) {
  BVAE_REQUIRE(d && d->x && d->w && d->y, BVAE_ERR_SHAPE, "conv_gemm: null pointer");
  BVAE_REQUIRE(d->ntaps >= 1 && d->ntaps <= BVAE_MAX_TAPS, BVAE_ERR_SHAPE, "conv_gemm: ntaps=%d", d->ntaps);
  BVAE_REQUIRE(d->N > 0 && d->QH > 0 && d->QW > 0 && d->C > 0 && d->Cout > 0, BVAE_ERR_SHAPE, "conv_gemm: empty problem");
  BVAE_REQUIRE(d->sy >= 1 && d->sx >= 1 && d->osy >= 1 && d->osx >= 1, BVAE_ERR_SHAPE, "conv_gemm: bad strides");
  BVAE_REQUIRE((d->QH - 1) * d->osy + d->ooy < d->OH && (d->QW - 1) * d->osx + d->oox < d->OW, BVAE_ERR_SHAPE,
               "conv_gemm: phase grid exceeds the output tensor");
  BVAE_REQUIRE(d->w_pitch >= d->ntaps * d->C, BVAE_ERR_SHAPE, "conv_gemm: w_pitch too small");
  BVAE_REQUIRE(!d->stats || (impl != BVAE_IMPL_SIMT && conv_tc_stats_ok(d)), BVAE_ERR_UNSUPPORTED,
               "conv_gemm: statistics fusion is not available for this problem (check bvae_conv_stats_ok)");
  if (impl == BVAE_IMPL_SIMT) return conv_simt_launch(d, (cudaStream_t)stream);
  if (impl == BVAE_IMPL_AUTO && stem_fwd_eligible(d)) return stem_fwd_launch(d, (cudaStream_t)stream);
  const int ok = conv_tc_eligible(d);
  if (impl == BVAE_IMPL_TC) {
    BVAE_REQUIRE(ok, BVAE_ERR_UNSUPPORTED, "conv_gemm: shape not eligible for the tcgen05 kernel (C=%d Cout=%d)", d->C, d->Cout);
    return conv_tc_launch(d, (cudaStream_t)stream);
  }
  return ok ? conv_tc_launch(d, (cudaStream_t)stream) : conv_simt_launch(d, (cudaStream_t)stream);
}

extern "C" int bvae_conv_stats_ok(const bvae_conv_desc* d) { return d && conv_tc_stats_ok(d); }

extern "C" int bvae_wgrad_gemm(const bvae_wgrad_desc* d, int impl, void* stream) {
  BVAE_REQUIRE(d && d->a && d->s && d->dw, BVAE_ERR_SHAPE, "wgrad_gemm: null pointer");
  BVAE_REQUIRE(d->ntaps >= 1 && d->ntaps <= BVAE_MAX_TAPS && d->T >= d->ntaps, BVAE_ERR_SHAPE, "wgrad_gemm: ntaps=%d T=%d", d->ntaps, d->T);
  BVAE_REQUIRE(d->N > 0 && d->AH > 0 && d->AW > 0 && d->Ca > 0 && d->Cs > 0, BVAE_ERR_SHAPE, "wgrad_gemm: empty problem");
  if (impl == BVAE_IMPL_SIMT) return wgrad_simt_launch(d, (cudaStream_t)stream);
  if (impl == BVAE_IMPL_AUTO && stem_wgrad_eligible(d)) return stem_wgrad_launch(d, (cudaStream_t)stream);
  const int ok = wgrad_tc_eligible(d);
  if (impl == BVAE_IMPL_TC) {
    BVAE_REQUIRE(ok, BVAE_ERR_UNSUPPORTED, "wgrad_gemm: shape not eligible for the tcgen05 kernel (Ca=%d Cs=%d)", d->Ca, d->Cs);
    return wgrad_tc_launch(d, (cudaStream_t)stream);
  }
  return ok ? wgrad_tc_launch(d, (cudaStream_t)stream) : wgrad_simt_launch(d, (cudaStream_t)stream);
}

// batchnorm.cu -- nn.BatchNorm2d (batch statistics in training, running statistics in eval) + optional (Leaky)ReLU on NHWC
// tensors with few channels (C a power of two <= 64): the norm of the GAN-phase piano-roll discriminator
// (graph/bar_discriminator.py:19-23,69,113-114,153) and of the Refiner (graph/refiner.py:13,20,37,43).
// These tensors are tiny next to the generator's (<= 64 channels on <= 192x60 maps, 0.3 % of the step's bytes); the
// kernels are plain coalesced element streams.  Every reduction is two-stage -- per-CTA partials in a scratch buffer, summed
// in CTA order by ONE finalising CTA -- so the statistics are deterministic (no float atomics anywhere in this file).
#include "common.cuh"

namespace bvae {

constexpr int BN_THREADS = 256;
constexpr int BN_MAX_CTAS = 148 * 8;

template <bool F32>
__device__ __forceinline__ float ld(const void* p, int64_t i) {
  return F32 ? ((const float*)p)[i] : bf2f(((const bf16*)p)[i]);
}

// stage 1: per-CTA, per-channel shifted sums.  Thread t always works on channel t % C (256 % C == 0, and the grid stride is a
// multiple of 256), so its two accumulators belong to one channel; the CTA folds its 256/C threads per channel in a fixed order.
// part layout: [cta][2][C]
template <bool F32>
__global__ void __launch_bounds__(BN_THREADS) bn_stats_kernel(const void* __restrict__ x, int64_t P, int C, int pitch,
                                                              float* __restrict__ part) {
  __shared__ float s_a[BN_THREADS], s_b[BN_THREADS];
  const int t = threadIdx.x, c = t % C;
  const int ppb = BN_THREADS / C;                      // pixels per CTA pass
  const float shift = ld<F32>(x, c);                   // first pixel: keeps the sums small when |mean| >> std
  float a = 0.f, b = 0.f;
  for (int64_t p = (int64_t)blockIdx.x * ppb + t / C; p < P; p += (int64_t)gridDim.x * ppb) {
    const float v = ld<F32>(x, p * pitch + c) - shift;
    a += v;
    b += v * v;
  }
  s_a[t] = a; s_b[t] = b;
  __syncthreads();
  if (t < C) {
    float sa = 0.f, sb = 0.f;
    for (int k = t; k < BN_THREADS; k += C) { sa += s_a[k]; sb += s_b[k]; }
    part[((int64_t)blockIdx.x * 2 + 0) * C + t] = sa;
    part[((int64_t)blockIdx.x * 2 + 1) * C + t] = sb;
  }
}

// stage 2 (one CTA): batch mean / biased variance -> save = {mean[C], rstd[C]}; running statistics updated as nn.BatchNorm2d
// does (momentum, UNBIASED variance).  training == 0: save comes from the running statistics instead.
template <bool F32>
__global__ void bn_finalize_kernel(const void* __restrict__ x, const float* __restrict__ part, int nctas, int64_t P, int C,
                                   float eps, float momentum, int training, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ save) {
  const int c = threadIdx.x;
  if (c >= C) return;
  if (!training) {
    save[c] = running_mean[c];
    save[C + c] = rsqrtf(running_var[c] + eps);
    return;
  }
  double sa = 0.0, sb = 0.0;
  for (int k = 0; k < nctas; ++k) { sa += part[((int64_t)k * 2 + 0) * C + c]; sb += part[((int64_t)k * 2 + 1) * C + c]; }
  const double n = (double)P;
  const double md = sa / n;
  double var = sb / n - md * md;
  if (var < 0.0) var = 0.0;
  const float mean = ld<F32>(x, c) + (float)md;
  save[c] = mean;
  save[C + c] = rsqrtf((float)var + eps);
  if (running_mean) {
    const double unb = P > 1 ? var * n / (n - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

// y = act(gamma * (x - mean) * rstd + beta)
template <bool XF32, bool YF32>
__global__ void __launch_bounds__(BN_THREADS) bn_apply_kernel(const void* __restrict__ x, int64_t P, int C, int x_pitch,
                                                              const float* __restrict__ save, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, int act, float slope,
                                                              void* __restrict__ y, int y_pitch) {
  const int t = threadIdx.x, c = t % C, ppb = BN_THREADS / C;
  const float a = gamma[c] * save[C + c], b = beta[c] - save[c] * a;
  for (int64_t p = (int64_t)blockIdx.x * ppb + t / C; p < P; p += (int64_t)gridDim.x * ppb) {
    float v = a * ld<XF32>(x, p * x_pitch + c) + b;
    if (act) v = act_fwd(v, slope);
    if (YF32) ((float*)y)[p * y_pitch + c] = v;
    else ((bf16*)y)[p * y_pitch + c] = f2bf(v);
  }
}

// backward stage 1: per-CTA partials of sum(dy') and sum(dy' * xhat), dy' = dy * act'(y)   (y = the forward output)
template <bool XF32, bool DF32>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_stats_kernel(const void* __restrict__ x, const void* __restrict__ y,
                                                                  const void* __restrict__ dy, int64_t P, int C, int x_pitch,
                                                                  int y_pitch, int dy_pitch, const float* __restrict__ save,
                                                                  int act, float slope, float* __restrict__ part) {
  __shared__ float s_a[BN_THREADS], s_b[BN_THREADS];
  const int t = threadIdx.x, c = t % C, ppb = BN_THREADS / C;
  const float mean = save[c], rstd = save[C + c];
  float a = 0.f, b = 0.f;
  for (int64_t p = (int64_t)blockIdx.x * ppb + t / C; p < P; p += (int64_t)gridDim.x * ppb) {
    float g = ld<DF32>(dy, p * dy_pitch + c);
    if (act) g *= (bf2f(((const bf16*)y)[p * y_pitch + c]) > 0.f) ? 1.f : slope;
    const float xh = (ld<XF32>(x, p * x_pitch + c) - mean) * rstd;
    a += g;
    b += g * xh;
  }
  s_a[t] = a; s_b[t] = b;
  __syncthreads();
  if (t < C) {
    float sa = 0.f, sb = 0.f;
    for (int k = t; k < BN_THREADS; k += C) { sa += s_a[k]; sb += s_b[k]; }
    part[((int64_t)blockIdx.x * 2 + 0) * C + t] = sa;
    part[((int64_t)blockIdx.x * 2 + 1) * C + t] = sb;
  }
}

// backward stage 2 (one CTA): totals -> save[2C..4C) = {sum dy', sum dy' xhat}; dbeta += , dgamma +=
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ part, int nctas, int C, float* __restrict__ save,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int c = threadIdx.x;
  if (c >= C) return;
  double sa = 0.0, sb = 0.0;
  for (int k = 0; k < nctas; ++k) { sa += part[((int64_t)k * 2 + 0) * C + c]; sb += part[((int64_t)k * 2 + 1) * C + c]; }
  save[2 * C + c] = (float)sa;
  save[3 * C + c] = (float)sb;
  if (dbeta) dbeta[c] += (float)sa;
  if (dgamma) dgamma[c] += (float)sb;
}

// dx = gamma * rstd * (dy' - mean(dy') - xhat * mean(dy' xhat))     (training);   gamma * rstd * dy'   (eval)
template <bool XF32, bool DF32>
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_apply_kernel(const void* __restrict__ x, const void* __restrict__ y,
                                                                  const void* __restrict__ dy, int64_t P, int C, int x_pitch,
                                                                  int y_pitch, int dy_pitch, const float* __restrict__ save,
                                                                  const float* __restrict__ gamma, int act, float slope,
                                                                  int training, bf16* __restrict__ dx, int dx_pitch) {
  const int t = threadIdx.x, c = t % C, ppb = BN_THREADS / C;
  const float mean = save[c], rstd = save[C + c], k = gamma[c] * rstd;
  const float inv = 1.f / (float)P;
  const float m1 = training ? save[2 * C + c] * inv : 0.f, m2 = training ? save[3 * C + c] * inv : 0.f;
  for (int64_t p = (int64_t)blockIdx.x * ppb + t / C; p < P; p += (int64_t)gridDim.x * ppb) {
    float g = ld<DF32>(dy, p * dy_pitch + c);
    if (act) g *= (bf2f(((const bf16*)y)[p * y_pitch + c]) > 0.f) ? 1.f : slope;
    const float xh = (ld<XF32>(x, p * x_pitch + c) - mean) * rstd;
    dx[p * dx_pitch + c] = f2bf(k * (g - m1 - xh * m2));
  }
}

static int bn_grid(int64_t P, int C) {
  const int ppb = BN_THREADS / C;
  int64_t g = ceil_div64(P, (int64_t)ppb * 4);
  if (g > BN_MAX_CTAS) g = BN_MAX_CTAS;
  if (g < 1) g = 1;
  return (int)g;
}

static int bn_validate(const bvae_bn_desc* d, const char* who) {
  BVAE_REQUIRE(d && d->x && d->save && d->scratch && d->gamma && d->beta, BVAE_ERR_SHAPE, "%s: null pointer", who);
  BVAE_REQUIRE(d->C >= 1 && d->C <= 64 && (d->C & (d->C - 1)) == 0, BVAE_ERR_SHAPE, "%s: C=%d must be a power of two <= 64", who, d->C);
  BVAE_REQUIRE(d->P > 0 && d->x_pitch >= d->C, BVAE_ERR_SHAPE, "%s: empty tensor / pitch < C", who);
  BVAE_REQUIRE(d->training || (d->running_mean && d->running_var), BVAE_ERR_SHAPE, "%s: eval mode needs running statistics", who);
  return BVAE_OK;
}

}  // namespace bvae

using namespace bvae;

extern "C" int bvae_bn_scratch_floats(int C) { return 2 * C * BN_MAX_CTAS; }

extern "C" int bvae_bn_forward(const bvae_bn_desc* d, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  int rc = bn_validate(d, "bn_forward");
  if (rc) return rc;
  BVAE_REQUIRE(d->y && d->y_pitch >= d->C, BVAE_ERR_SHAPE, "bn_forward: output missing");
  const int grid = bn_grid(d->P, d->C);
  if (d->training) {
    if (d->x_f32) bn_stats_kernel<true><<<grid, BN_THREADS, 0, st>>>(d->x, d->P, d->C, d->x_pitch, d->scratch);
    else bn_stats_kernel<false><<<grid, BN_THREADS, 0, st>>>(d->x, d->P, d->C, d->x_pitch, d->scratch);
    if ((rc = check_launch("bn_stats"))) return rc;
  }
  if (d->x_f32)
    bn_finalize_kernel<true><<<1, 64, 0, st>>>(d->x, d->scratch, grid, d->P, d->C, d->eps, d->momentum, d->training,
                                               d->running_mean, d->running_var, d->save);
  else
    bn_finalize_kernel<false><<<1, 64, 0, st>>>(d->x, d->scratch, grid, d->P, d->C, d->eps, d->momentum, d->training,
                                                d->running_mean, d->running_var, d->save);
  if ((rc = check_launch("bn_finalize"))) return rc;
#define BN_APPLY(XF, YF) bn_apply_kernel<XF, YF><<<grid, BN_THREADS, 0, st>>>(d->x, d->P, d->C, d->x_pitch, d->save, d->gamma, \
                                                                               d->beta, d->act, d->slope, d->y, d->y_pitch)
  if (d->x_f32) { if (d->y_f32) BN_APPLY(true, true); else BN_APPLY(true, false); }
  else { if (d->y_f32) BN_APPLY(false, true); else BN_APPLY(false, false); }
#undef BN_APPLY
  return check_launch("bn_apply");
}

extern "C" int bvae_bn_backward(const bvae_bn_desc* d, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  int rc = bn_validate(d, "bn_backward");
  if (rc) return rc;
  BVAE_REQUIRE(d->dy && d->dx && d->dy_pitch >= d->C && d->dx_pitch >= d->C, BVAE_ERR_SHAPE, "bn_backward: gradients missing");
  BVAE_REQUIRE(!d->act || (d->y && !d->y_f32), BVAE_ERR_SHAPE, "bn_backward: the activation mask is read from the bf16 output");
  const int grid = bn_grid(d->P, d->C);
#define BN_BSTATS(XF, DF) bn_bwd_stats_kernel<XF, DF><<<grid, BN_THREADS, 0, st>>>(d->x, d->y, d->dy, d->P, d->C, d->x_pitch, \
                                                                                    d->y_pitch, d->dy_pitch, d->save, d->act,  \
                                                                                    d->slope, d->scratch)
  if (d->x_f32) { if (d->dy_f32) BN_BSTATS(true, true); else BN_BSTATS(true, false); }
  else { if (d->dy_f32) BN_BSTATS(false, true); else BN_BSTATS(false, false); }
#undef BN_BSTATS
  if ((rc = check_launch("bn_bwd_stats"))) return rc;
  bn_bwd_finalize_kernel<<<1, 64, 0, st>>>(d->scratch, grid, d->C, d->save, d->dgamma, d->dbeta);
  if ((rc = check_launch("bn_bwd_finalize"))) return rc;
#define BN_BAPPLY(XF, DF) bn_bwd_apply_kernel<XF, DF><<<grid, BN_THREADS, 0, st>>>(d->x, d->y, d->dy, d->P, d->C, d->x_pitch, \
                                                                                    d->y_pitch, d->dy_pitch, d->save, d->gamma, \
                                                                                    d->act, d->slope, d->training,             \
                                                                                    (bf16*)d->dx, d->dx_pitch)
  if (d->x_f32) { if (d->dy_f32) BN_BAPPLY(true, true); else BN_BAPPLY(true, false); }
  else { if (d->dy_f32) BN_BAPPLY(false, true); else BN_BAPPLY(false, false); }
#undef BN_BAPPLY
  return check_launch("bn_bwd_apply");
}

// conv_simt.cu -- CUDA-core implicit-GEMM kernels for the bvae_conv_desc / bvae_wgrad_desc contractions.
// Role: (1) the on-device checker the tcgen05 kernels are validated against, (2) the path for shapes the
// tensor-core kernels do not take (C_in = 1 stems, channel counts that are not a multiple of 32).
#include "common.cuh"

namespace bvae {

constexpr int SM_BM = 64, SM_BN = 64, SM_BK = 32;

// out[m, co] = epi(sum_t sum_c x[pix(m,t), c] * w[co, t*C + c])
__global__ void __launch_bounds__(256) conv_simt_kernel(const bvae_conv_desc d) {
  __shared__ float As[SM_BK][SM_BM + 4];
  __shared__ float Bs[SM_BK][SM_BN + 4];
  const bf16* __restrict__ x = (const bf16*)d.x;
  const bf16* __restrict__ w = (const bf16*)d.w;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int64_t M = (int64_t)d.N * d.QH * d.QW;
  const int64_t m0 = (int64_t)blockIdx.x * SM_BM;
  const int n0 = blockIdx.y * SM_BN;

  // each thread loads 8 A elements and 8 B elements per K-step: element i -> (row = i / 32, kk = i % 32)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int t = 0; t < d.ntaps; ++t) {
    for (int c0 = 0; c0 < d.C; c0 += SM_BK) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int e = tid + i * 256;
        const int r = e / SM_BK, kk = e % SM_BK;
        const int64_t m = m0 + r;
        float v = 0.f;
        if (m < M && c0 + kk < d.C) {
          const int qx = (int)(m % d.QW);
          const int qy = (int)((m / d.QW) % d.QH);
          const int n = (int)(m / ((int64_t)d.QW * d.QH));
          const int iy = qy * d.sy + d.dy[t], ix = qx * d.sx + d.dx[t];
          if (iy >= 0 && iy < d.H && ix >= 0 && ix < d.W)
            v = bf2f(x[(((int64_t)n * d.H + iy) * d.W + ix) * d.x_pitch + c0 + kk]);
        }
        As[kk][r] = v;
        const int co = n0 + r;
        float wv = 0.f;
        if (co < d.Cout && c0 + kk < d.C) wv = bf2f(w[(int64_t)co * d.w_pitch + (int64_t)t * d.C + c0 + kk]);
        Bs[kk][r] = wv;
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < SM_BK; ++kk) {
        float a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
      }
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t m = m0 + ty * 4 + i;
    if (m >= M) continue;
    const int qx = (int)(m % d.QW);
    const int qy = (int)((m / d.QW) % d.QH);
    const int n = (int)(m / ((int64_t)d.QW * d.QH));
    const int64_t opix = ((int64_t)n * d.OH + (qy * d.osy + d.ooy)) * d.OW + (qx * d.osx + d.oox);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = n0 + tx * 4 + j;
      if (co >= d.Cout) continue;
      float v = acc[i][j];
      if (d.bias) v += d.bias[co];
      if (d.act) v = act_fwd(v, d.slope);
      if (d.addend) {
        if (d.out_f32) v += ((const float*)d.addend)[opix * d.add_pitch + co];
        else v += bf2f(((const bf16*)d.addend)[opix * d.add_pitch + co]);
      }
      if (d.mask) {
        const float mk = bf2f(((const bf16*)d.mask)[opix * d.mask_pitch + co]);
        v *= (mk > 0.f) ? 1.f : d.mask_slope;
      }
      if (d.out_f32) ((float*)d.y)[opix * d.y_pitch + co] = v;
      else ((bf16*)d.y)[opix * d.y_pitch + co] = f2bf(v);
    }
  }
}

// dw[(ra*Cs + rs)*T + tap_idx[t]] += sum_pix a[pix, ra] * s[pix_s(t), rs]
__global__ void __launch_bounds__(256) wgrad_simt_kernel(const bvae_wgrad_desc d, int splits, int64_t pix_per_split) {
  __shared__ float As[SM_BK][SM_BM + 4];
  __shared__ float Ss[SM_BK][SM_BN + 4];
  const bf16* __restrict__ a = (const bf16*)d.a;
  const bf16* __restrict__ s = (const bf16*)d.s;
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  const int ra0 = blockIdx.x * SM_BM, rs0 = blockIdx.y * SM_BN;
  const int t = blockIdx.z / splits, sp = blockIdx.z % splits;
  const int64_t P = (int64_t)d.N * d.AH * d.AW;
  const int64_t p_begin = (int64_t)sp * pix_per_split;
  const int64_t p_end = min(P, p_begin + pix_per_split);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int64_t p0 = p_begin; p0 < p_end; p0 += SM_BK) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = tid + i * 256;
      const int kk = e / SM_BM, r = e % SM_BM;  // consecutive threads -> consecutive channels (coalesced)
      const int64_t p = p0 + kk;
      float av = 0.f, sv = 0.f;
      if (p < p_end) {
        const int ax = (int)(p % d.AW);
        const int ay = (int)((p / d.AW) % d.AH);
        const int n = (int)(p / ((int64_t)d.AW * d.AH));
        if (ra0 + r < d.Ca) av = bf2f(a[p * d.a_pitch + ra0 + r]);
        const int yy = ay * d.sy + d.dy[t], xx = ax * d.sx + d.dx[t];
        if (rs0 + r < d.Cs && yy >= 0 && yy < d.SH && xx >= 0 && xx < d.SW)
          sv = bf2f(s[(((int64_t)n * d.SH + yy) * d.SW + xx) * d.s_pitch + rs0 + r]);
      }
      As[kk][r] = av;
      Ss[kk][r] = sv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SM_BK; ++kk) {
      float av[4], sv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) sv[j] = Ss[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += av[i] * sv[j];
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ra = ra0 + ty * 4 + i;
    if (ra >= d.Ca) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int rs = rs0 + tx * 4 + j;
      if (rs >= d.Cs) continue;
      atomicAdd(d.dw + ((int64_t)ra * d.Cs + rs) * d.T + d.tap_idx[t], acc[i][j]);
    }
  }
}

int conv_simt_launch(const bvae_conv_desc* d, cudaStream_t stream) {
  const int64_t M = (int64_t)d->N * d->QH * d->QW;
  dim3 grid((unsigned)ceil_div64(M, SM_BM), ceil_div(d->Cout, SM_BN));
  conv_simt_kernel<<<grid, 256, 0, stream>>>(*d);
  note_kernel("conv_simt_kernel");
  return check_launch("conv_simt");
}

int wgrad_simt_launch(const bvae_wgrad_desc* d, cudaStream_t stream) {
  const int64_t P = (int64_t)d->N * d->AH * d->AW;
  const int tiles = ceil_div(d->Ca, SM_BM) * ceil_div(d->Cs, SM_BN) * d->ntaps;
  int splits = ceil_div(148 * 4, tiles);
  const int64_t max_splits = ceil_div64(P, 4 * SM_BK);
  if (splits > max_splits) splits = (int)max_splits;
  if (splits < 1) splits = 1;
  int64_t pps = ceil_div64(P, splits);
  pps = ceil_div64(pps, SM_BK) * SM_BK;
  splits = (int)ceil_div64(P, pps);
  dim3 grid(ceil_div(d->Ca, SM_BM), ceil_div(d->Cs, SM_BN), d->ntaps * splits);
  wgrad_simt_kernel<<<grid, 256, 0, stream>>>(*d, splits, pps);
  note_kernel("wgrad_simt_kernel");
  return check_launch("wgrad_simt");
}

}  // namespace bvae

// conv_small.cu -- direct convolution / weight-gradient kernels for layers with fewer than 32 input or output channels:
// the 1..16-channel convolutions of the GAN-phase piano-roll discriminator (graph/bar_discriminator.py:11-24,65-79,
// 141-151) and of the Refiner (graph/refiner.py:12,19,36,42), forward, data gradient (the same descriptor with the
// transposed weights) and weight gradient.  Nothing here is GEMM-shaped enough for a tensor core (K = taps x C_in is 3..144,
// N = C_out is 1..32); the generic CUDA-core implicit-GEMM kernels of conv_simt.cu run these shapes on 64x64x32 tiles that
// are 8-32x larger than the problem (measured: 160 ms of a 233 ms adversarial iteration at 512 bars).  Here every thread
// owns ONE output pixel x 8 output channels (or all of them when C_out < 8), the layer's weights live in shared memory as
// fp32 and the input pixels -- 2..64 bytes each in NHWC -- come through L1; lanes of a warp that share a pixel broadcast
// the same input address and write consecutive 16-byte chunks.
#include "common.cuh"

namespace bvae {

constexpr int CS_MAX_W = 9216;            // fp32 weights in shared memory: taps * C * Cout <= 9216 (36 KB)

// out[pix, co] = epi(sum_t sum_c x[pix_t, c] * w[co, t*C + c]);  CO = output channels per thread (8, or Cout when Cout < 8)
template <int CO>
__global__ void __launch_bounds__(256) conv_small_kernel(const bvae_conv_desc d) {
  __shared__ float s_w[CS_MAX_W];         // [t][c][co]
  const bf16* __restrict__ x = (const bf16*)d.x;
  const bf16* __restrict__ w = (const bf16*)d.w;
  const int C = d.C, Cout = d.Cout, T = d.ntaps;
  for (int i = threadIdx.x; i < T * C * Cout; i += blockDim.x) {
    const int co = i % Cout, c = (i / Cout) % C, t = i / (Cout * C);
    s_w[i] = bf2f(w[(int64_t)co * d.w_pitch + t * C + c]);
  }
  __syncthreads();
  const int groups = Cout / CO;
  const int64_t M = (int64_t)d.N * d.QH * d.QW;
  const int64_t total = M * groups;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int g = (int)(e % groups);
    const int64_t m = e / groups;
    const int qx = (int)(m % d.QW);
    const int qy = (int)((m / d.QW) % d.QH);
    const int n = (int)(m / ((int64_t)d.QW * d.QH));
    float acc[CO];
#pragma unroll
    for (int j = 0; j < CO; ++j) acc[j] = 0.f;
    for (int t = 0; t < T; ++t) {
      const int iy = qy * d.sy + d.dy[t], ix = qx * d.sx + d.dx[t];
      if (iy < 0 || iy >= d.H || ix < 0 || ix >= d.W) continue;
      const bf16* xp = x + (((int64_t)n * d.H + iy) * d.W + ix) * d.x_pitch;
      const float* wp = s_w + (int64_t)t * C * Cout + g * CO;
      if ((C & 7) == 0 && (d.x_pitch & 7) == 0 && (((uintptr_t)x) & 15) == 0) {
        for (int c0 = 0; c0 < C; c0 += 8) {
          float xv[8];
          unpack8(ldg8(xp + c0), xv);
#pragma unroll
          for (int k = 0; k < 8; ++k)
#pragma unroll
            for (int j = 0; j < CO; ++j) acc[j] = fmaf(xv[k], wp[(c0 + k) * Cout + j], acc[j]);
        }
      } else {
        for (int c = 0; c < C; ++c) {
          const float xv = bf2f(xp[c]);
#pragma unroll
          for (int j = 0; j < CO; ++j) acc[j] = fmaf(xv, wp[c * Cout + j], acc[j]);
        }
      }
    }
    const int64_t opix = ((int64_t)n * d.OH + (qy * d.osy + d.ooy)) * d.OW + (qx * d.osx + d.oox);
#pragma unroll
    for (int j = 0; j < CO; ++j) {
      const int co = g * CO + j;
      float v = acc[j];
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
      acc[j] = v;
    }
    if (d.out_f32) {
      float* yp = (float*)d.y + opix * d.y_pitch + g * CO;
#pragma unroll
      for (int j = 0; j < CO; ++j) yp[j] = acc[j];
    } else {
      bf16* yp = (bf16*)d.y + opix * d.y_pitch + g * CO;
      if (CO == 8 && (d.y_pitch & 7) == 0 && (((uintptr_t)d.y) & 15) == 0) stg8(yp, pack8(acc));
      else {
#pragma unroll
        for (int j = 0; j < CO; ++j) yp[j] = f2bf(acc[j]);
      }
    }
  }
}

// dw[(ra*Cs + rs)*T + tap_idx[t]] += sum_pix a[pix, ra] * s[pix_s(t), rs].   Ca * Cs * ntaps <= CS_MAX_W.
// A thread owns RA anchor channels x one shifted channel x all TT taps (RA*TT register accumulators) for a strided subset
// of the CTA's pixel range: 256 / nown pixel lanes work side by side.  The lanes' partial sums meet in a shared-memory copy of
// the weight gradient (shared atomics), which the CTA then adds to global memory with one atomic per element.
template <int RA, int TT>
__global__ void __launch_bounds__(256) wgrad_small_kernel(const bvae_wgrad_desc d, int64_t pix_per_cta) {
  __shared__ float s_dw[CS_MAX_W];        // [t][ra][rs]
  const bf16* __restrict__ a = (const bf16*)d.a;
  const bf16* __restrict__ s = (const bf16*)d.s;
  const int Ca = d.Ca, Cs = d.Cs;
  const int nown = (Ca / RA) * Cs;        // (ra-group, rs) pairs
  const int64_t P = (int64_t)d.N * d.AH * d.AW;
  const int64_t p0 = (int64_t)blockIdx.x * pix_per_cta, p1 = min(P, p0 + pix_per_cta);
  for (int i = threadIdx.x; i < TT * Ca * Cs; i += blockDim.x) s_dw[i] = 0.f;
  __syncthreads();
  const int lanes = nown >= 256 ? 1 : 256 / nown;           // pixel lanes per owner
  for (int base = 0; base < nown; base += 256) {
    const int own = base + (int)(threadIdx.x % (nown >= 256 ? 256 : nown));
    const int pl = nown >= 256 ? 0 : (int)(threadIdx.x / nown);
    if (own >= nown || pl >= lanes) continue;
    const int rs = own % Cs, rg = own / Cs;
    float acc[TT][RA];
#pragma unroll
    for (int t = 0; t < TT; ++t)
#pragma unroll
      for (int j = 0; j < RA; ++j) acc[t][j] = 0.f;
    for (int64_t p = p0 + pl; p < p1; p += lanes) {
      const int ax = (int)(p % d.AW);
      const int ay = (int)((p / d.AW) % d.AH);
      const int n = (int)(p / ((int64_t)d.AW * d.AH));
      float av[RA];
      const bf16* ap = a + p * d.a_pitch + rg * RA;
#pragma unroll
      for (int j = 0; j < RA; ++j) av[j] = bf2f(ap[j]);
#pragma unroll
      for (int t = 0; t < TT; ++t) {
        const int yy = ay * d.sy + d.dy[t], xx = ax * d.sx + d.dx[t];
        if (yy < 0 || yy >= d.SH || xx < 0 || xx >= d.SW) continue;
        const float sv = bf2f(s[(((int64_t)n * d.SH + yy) * d.SW + xx) * d.s_pitch + rs]);
#pragma unroll
        for (int j = 0; j < RA; ++j) acc[t][j] = fmaf(av[j], sv, acc[t][j]);
      }
    }
#pragma unroll
    for (int t = 0; t < TT; ++t)
#pragma unroll
      for (int j = 0; j < RA; ++j) atomicAdd(&s_dw[((int64_t)t * Ca + rg * RA + j) * Cs + rs], acc[t][j]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < TT * Ca * Cs; i += blockDim.x) {
    const int rs = i % Cs, ra = (i / Cs) % Ca, t = i / (Cs * Ca);
    const float v = s_dw[i];
    if (v != 0.f) atomicAdd(d.dw + ((int64_t)ra * Cs + rs) * d.T + d.tap_idx[t], v);
  }
}

int conv_small_eligible(const bvae_conv_desc* d) {
  if (d->C >= 32 && d->Cout >= 32) return 0;                       // the tensor-core kernels take those
  if (d->C > 64 || d->Cout > 64 || d->stats) return 0;
  if ((long)d->ntaps * d->C * d->Cout > CS_MAX_W) return 0;
  return d->Cout == 1 || d->Cout == 2 || d->Cout == 4 || d->Cout % 8 == 0;
}

int conv_small_launch(const bvae_conv_desc* d, cudaStream_t stream) {
  const int CO = d->Cout < 8 ? d->Cout : 8;
  const int64_t total = (int64_t)d->N * d->QH * d->QW * (d->Cout / CO);
  int64_t grid = ceil_div64(total, 256 * 2);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  switch (CO) {
    case 8: conv_small_kernel<8><<<(int)grid, 256, 0, stream>>>(*d); break;
    case 4: conv_small_kernel<4><<<(int)grid, 256, 0, stream>>>(*d); break;
    case 2: conv_small_kernel<2><<<(int)grid, 256, 0, stream>>>(*d); break;
    case 1: conv_small_kernel<1><<<(int)grid, 256, 0, stream>>>(*d); break;
    default: set_error("conv_small: Cout=%d", d->Cout); return BVAE_ERR_UNSUPPORTED;
  }
  note_kernel("conv_small_kernel<%d>", CO);
  return check_launch("conv_small");
}

static bool small_taps_ok(int t) { return t == 1 || t == 3 || t == 4 || t == 9 || t == 16; }

int wgrad_small_eligible(const bvae_wgrad_desc* d) {
  if (d->Ca >= 32 && d->Cs >= 32) return 0;
  if (d->Ca > 64 || d->Cs > 64 || !small_taps_ok(d->ntaps)) return 0;
  return (long)d->ntaps * d->Ca * d->Cs <= CS_MAX_W;
}

template <int RA>
static void wgrad_small_dispatch(const bvae_wgrad_desc* d, int ctas, int64_t ppc, cudaStream_t stream) {
  switch (d->ntaps) {
    case 1: wgrad_small_kernel<RA, 1><<<ctas, 256, 0, stream>>>(*d, ppc); break;
    case 3: wgrad_small_kernel<RA, 3><<<ctas, 256, 0, stream>>>(*d, ppc); break;
    case 4: wgrad_small_kernel<RA, 4><<<ctas, 256, 0, stream>>>(*d, ppc); break;
    case 9: wgrad_small_kernel<RA, 9><<<ctas, 256, 0, stream>>>(*d, ppc); break;
    default: wgrad_small_kernel<RA, 16><<<ctas, 256, 0, stream>>>(*d, ppc); break;
  }
}

int wgrad_small_launch(const bvae_wgrad_desc* d, cudaStream_t stream) {
  const int64_t P = (int64_t)d->N * d->AH * d->AW;
  // enough CTAs to fill the machine, each with enough pixels to amortise its pass over the weight gradient
  int64_t ctas = 148 * 8;
  int64_t ppc = ceil_div64(P, ctas);
  if (ppc < 256) ppc = 256;
  ctas = ceil_div64(P, ppc);
  // registers: RA x taps accumulators per thread
  int RA = d->Ca % 8 == 0 ? 8 : (d->Ca % 4 == 0 ? 4 : (d->Ca % 2 == 0 ? 2 : 1));
  if (d->ntaps == 16 && RA == 8) RA = 4;
  switch (RA) {
    case 8: wgrad_small_dispatch<8>(d, (int)ctas, ppc, stream); break;
    case 4: wgrad_small_dispatch<4>(d, (int)ctas, ppc, stream); break;
    case 2: wgrad_small_dispatch<2>(d, (int)ctas, ppc, stream); break;
    default: wgrad_small_dispatch<1>(d, (int)ctas, ppc, stream); break;
  }
  note_kernel("wgrad_small_kernel<%d,%d>", RA, d->ntaps);
  return check_launch("wgrad_small");
}

}  // namespace bvae

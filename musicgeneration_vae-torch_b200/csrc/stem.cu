// stem.cu -- the C_in = 1 convolutions of the encoder stems (graph/encodingBlock.py:12-13,43-44) and their weight
// gradients.  K = 4 taps x 1 channel: nothing for a tensor core to do; both are pure streaming kernels
// (32 bf16 channels written / read per pixel, the 1-channel piano-roll stays in L1/L2).
#include "common.cuh"

namespace bvae {

// out[n, oy, ox, c] = act(sum_t w[c][t] * x[n, oy*sy + dy_t, ox*sx + dx_t]);  G = Cout/8 lanes per pixel, 16-byte stores
template <int NT>
__global__ void __launch_bounds__(256) stem_fwd_kernel(const bvae_conv_desc d) {
  const bf16* __restrict__ x = (const bf16*)d.x;
  const bf16* __restrict__ w = (const bf16*)d.w;
  bf16* __restrict__ y = (bf16*)d.y;
  const int G = d.Cout / 8;
  const int64_t P = (int64_t)d.N * d.QH * d.QW;
  const int sub = threadIdx.x % G;
  float wr[8][NT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int t = 0; t < NT; ++t) wr[i][t] = bf2f(w[(int64_t)(sub * 8 + i) * d.w_pitch + t]);
  float bs[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) bs[i] = d.bias ? d.bias[sub * 8 + i] : 0.f;
  // A CTA pass covers 256/G pixel slots laid out as whole output rows of the padded width RW (a power of two >= QW):
  // slot -> (row, qx) by shift / mask and ONE 32-bit division per pass, instead of three 64-bit div/mod per pixel
  // (the first version was bound by that integer arithmetic).  Consecutive CTAs keep working on consecutive rows:
  // unrolling a CTA over several distant rows was 2x SLOWER (it scatters the write stream over DRAM pages).
  const int slots = blockDim.x / G;
  int RW = 1;
  while (RW < d.QW && RW < slots) RW <<= 1;
  if (RW < d.QW) {                                   // rows wider than a pass: generic per-pixel indexing
    const int64_t stride = (int64_t)gridDim.x * slots;
    for (int64_t p = (int64_t)blockIdx.x * slots + threadIdx.x / G; p < P; p += stride) {
      const int qx = (int)(p % d.QW);
      const int qy = (int)((p / d.QW) % d.QH);
      const int n = (int)(p / ((int64_t)d.QW * d.QH));
      float xv[NT];
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int iy = qy * d.sy + d.dy[t], ix = qx * d.sx + d.dx[t];
        xv[t] = (iy >= 0 && iy < d.H && ix >= 0 && ix < d.W) ? bf2f(x[(((int64_t)n * d.H + iy) * d.W + ix) * d.x_pitch]) : 0.f;
      }
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = bs[i];
#pragma unroll
        for (int t = 0; t < NT; ++t) v += wr[i][t] * xv[t];
        o[i] = d.act ? act_fwd(v, d.slope) : v;
      }
      const int64_t opix = ((int64_t)n * d.OH + (qy * d.osy + d.ooy)) * d.OW + (qx * d.osx + d.oox);
      stg8(y + opix * d.y_pitch + sub * 8, pack8(o));
    }
    return;
  }
  const int rpp = slots / RW;                        // rows per pass
  const int slot = threadIdx.x / G;
  const int qx = slot & (RW - 1), rl = slot / RW;
  const int rows = d.N * d.QH;
  int dyt[NT], dxt[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) { dyt[t] = d.dy[t]; dxt[t] = qx * d.sx + d.dx[t]; }
  if (qx >= d.QW) return;
  for (int row = blockIdx.x * rpp + rl; row < rows; row += gridDim.x * rpp) {
    const int n = row / d.QH, qy = row - n * d.QH;
    const bf16* xn = x + (int64_t)n * d.H * d.W * d.x_pitch;
    float xv[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const int iy = qy * d.sy + dyt[t], ix = dxt[t];
      xv[t] = (iy >= 0 && iy < d.H && ix >= 0 && ix < d.W) ? bf2f(xn[(iy * d.W + ix) * d.x_pitch]) : 0.f;
    }
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = bs[i];
#pragma unroll
      for (int t = 0; t < NT; ++t) v += wr[i][t] * xv[t];
      o[i] = d.act ? act_fwd(v, d.slope) : v;
    }
    const int64_t opix = ((int64_t)n * d.OH + (qy * d.osy + d.ooy)) * d.OW + (qx * d.osx + d.oox);
    stg8(y + opix * d.y_pitch + sub * 8, pack8(o));
  }
}

// dw[ra*T + tap_idx[t]] += sum_pix a[pix, ra] * s[pix*stride + d_t]   (Cs == 1)
template <int NT>
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const bvae_wgrad_desc d) {
  __shared__ float s_acc[64 * NT];
  const bf16* __restrict__ a = (const bf16*)d.a;
  const bf16* __restrict__ s = (const bf16*)d.s;
  const int G = d.Ca / 8;
  const int lane = threadIdx.x & 31;
  const int sub = threadIdx.x % G;
  const int64_t P = (int64_t)d.N * d.AH * d.AW;
  for (int i = threadIdx.x; i < d.Ca * NT; i += blockDim.x) s_acc[i] = 0.f;
  __syncthreads();
  float acc[8][NT];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[i][t] = 0.f;
  const int slots = blockDim.x / G;
  int RW = 1;
  while (RW < d.AW && RW < slots) RW <<= 1;
  if (RW < d.AW) {                                   // rows wider than a pass: generic per-pixel indexing
    const int64_t stride = (int64_t)gridDim.x * slots;
    for (int64_t p = (int64_t)blockIdx.x * slots + threadIdx.x / G; p < P; p += stride) {
      const int ax = (int)(p % d.AW);
      const int ay = (int)((p / d.AW) % d.AH);
      const int n = (int)(p / ((int64_t)d.AW * d.AH));
      float av[8], xv[NT];
      unpack8(ldg8(a + p * d.a_pitch + sub * 8), av);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const int yy = ay * d.sy + d.dy[t], xx = ax * d.sx + d.dx[t];
        xv[t] = (yy >= 0 && yy < d.SH && xx >= 0 && xx < d.SW) ? bf2f(s[(((int64_t)n * d.SH + yy) * d.SW + xx) * d.s_pitch]) : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int t = 0; t < NT; ++t) acc[i][t] += av[i] * xv[t];
    }
  } else {
    // whole anchor rows of the padded width RW per pass (see stem_fwd_kernel): no per-pixel divisions
    const int rpp = slots / RW;
    const int slot = threadIdx.x / G;
    const int ax = slot & (RW - 1), rl = slot / RW;
    const int rows = d.N * d.AH;
    int dyt[NT], dxt[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) { dyt[t] = d.dy[t]; dxt[t] = ax * d.sx + d.dx[t]; }
    if (ax < d.AW) {
      for (int row = blockIdx.x * rpp + rl; row < rows; row += gridDim.x * rpp) {
        const int n = row / d.AH, ay = row - n * d.AH;
        const bf16* sn = s + (int64_t)n * d.SH * d.SW * d.s_pitch;
        float av[8], xv[NT];
        unpack8(ldg8(a + ((int64_t)row * d.AW + ax) * d.a_pitch + sub * 8), av);
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const int yy = ay * d.sy + dyt[t], xx = dxt[t];
          xv[t] = (yy >= 0 && yy < d.SH && xx >= 0 && xx < d.SW) ? bf2f(sn[(yy * d.SW + xx) * d.s_pitch]) : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int t = 0; t < NT; ++t) acc[i][t] += av[i] * xv[t];
      }
    }
  }
  for (int o = G; o < 32; o <<= 1)
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int t = 0; t < NT; ++t) acc[i][t] += __shfl_xor_sync(0xffffffffu, acc[i][t], o);
  if (lane < G)
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int t = 0; t < NT; ++t) atomicAdd(&s_acc[(sub * 8 + i) * NT + t], acc[i][t]);
  __syncthreads();
  for (int i = threadIdx.x; i < d.Ca * NT; i += blockDim.x) {
    const int ra = i / NT, t = i % NT;
    atomicAdd(d.dw + (int64_t)ra * d.T + d.tap_idx[t], s_acc[i]);
  }
}

int stem_fwd_eligible(const bvae_conv_desc* d) {
  const int G = d->Cout / 8;
  return d->C == 1 && d->ntaps == 4 && d->Cout % 8 == 0 && d->Cout <= 64 && (32 % G) == 0 && !d->addend && !d->mask &&
         !d->out_f32 && d->y_pitch % 8 == 0 && (((uintptr_t)d->y) & 15) == 0;
}

int stem_fwd_launch(const bvae_conv_desc* d, cudaStream_t stream) {
  const int64_t P = (int64_t)d->N * d->QH * d->QW;
  const int ppb = 256 / (d->Cout / 8);
  int64_t grid = ceil_div64(P, ppb * 4);
  if (grid > 148 * 16) grid = 148 * 16;
  if (grid < 1) grid = 1;
  stem_fwd_kernel<4><<<(int)grid, 256, 0, stream>>>(*d);
  note_kernel("stem_fwd_kernel<4>");
  return check_launch("stem_fwd");
}

int stem_wgrad_eligible(const bvae_wgrad_desc* d) {
  const int G = d->Ca / 8;
  return d->Cs == 1 && d->ntaps == 4 && d->Ca % 8 == 0 && d->Ca <= 64 && (32 % G) == 0 && d->a_pitch % 8 == 0 &&
         (((uintptr_t)d->a) & 15) == 0;
}

int stem_wgrad_launch(const bvae_wgrad_desc* d, cudaStream_t stream) {
  stem_wgrad_kernel<4><<<148 * 4, 256, 0, stream>>>(*d);
  note_kernel("stem_wgrad_kernel<4>");
  return check_launch("stem_wgrad");
}

}  // namespace bvae

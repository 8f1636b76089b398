// normblock.cu -- InstanceNorm2d(affine) [+ CBAM] [+ residual] + (Leaky)ReLU, forward and backward.
// All kernels are HBM-bound sweeps over NHWC tensors: a pixel's C channels are covered by G = min(32, C/8)
// lanes doing 16-byte loads (ITERS = C/(8G) rounds), per-pixel reductions are warp shuffles inside the lane
// group, per-(n,c) reductions go registers -> shuffle -> shared atomics -> one global atomic per CTA.
#include <cooperative_groups.h>
#include <stdlib.h>
#include "common.cuh"

namespace bvae {

typedef unsigned long long u64;

// nc layout (floats per (n,c))
enum { NC_MEAN = 0, NC_RSTD = 1, NC_A = 2, NC_B = 3, NC_GC = 4, NC_EXTU = 5, NC_EXTUHAT = 6, NC_SPARE = 7, NC_W = 8 };
// bwd_nc layout
enum { BN_DGC = 0, BN_S1 = 1, BN_S2 = 2, BN_DMX = 3, BN_W = 4 };
// bwd_px layout
enum { BP_DQ = 0, BP_DMEAN = 1, BP_DMAX = 2, BP_W = 4 };

// Deterministic mode (bvae_set_deterministic): shared-memory float accumulations of the FORWARD kernels run warp by warp,
// lane by lane, so their summation order is fixed.  All threads of the CTA must reach the call.
__constant__ int c_det = 0;
template <class F>
__device__ __forceinline__ void ordered(F f) {
  if (!c_det) { f(); return; }
  const int nw = blockDim.x >> 5, w = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = 0; i < nw; ++i) {
    if (w == i) {
      for (int l = 0; l < 32; ++l) {
        if (lane == l) f();
        __syncwarp();
      }
    }
    __syncthreads();
  }
}
static bool sync_det() {
  static int synced = -1;
  const int v = deterministic() ? 1 : 0;
  if (v != synced) {
    cudaMemcpyToSymbol(c_det, &v, sizeof(int));
    synced = v;
  }
  return v == 1;
}

template <bool F32>
__device__ __forceinline__ void load8(const void* base, int64_t off, float* f) {
  if (F32) ldg8f((const float*)base + off, f);
  else unpack8(ldg8((const bf16*)base + off), f);
}

// ---------------------------------------------------------------------------------------------------
// forward 1: per-(n,c) shifted sums and extrema (with first-index tie break) of the raw conv output
// grid (ceil(C/256), psplit, N), block 256
// ---------------------------------------------------------------------------------------------------
// EXT = false (sites without CBAM): only the sums -- nothing consumes the extrema or their positions there
template <bool F32, bool EXT>
__global__ void __launch_bounds__(256, 3) nb_stats_kernel(const void* __restrict__ y, int pitch, int HW, int C,
                                                       int psplit, float2* __restrict__ ss, u64* __restrict__ kmax,
                                                       u64* __restrict__ kmin) {
  __shared__ float s_sum[256], s_sq[256];
  __shared__ u64 s_kmax[256], s_kmin[256];
  const int Cc = min(C, 256);
  const int G = Cc / 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % G, grp = lane / G, gpw = 32 / G;
  const int cl = sub * 8;                       // channel offset inside this chunk
  const int c0 = blockIdx.x * 256 + cl;
  const int n = blockIdx.z;
  const int pps = ceil_div(HW, psplit);
  const int p_begin = blockIdx.y * pps, p_end = min(HW, p_begin + pps);
  const int64_t base = (int64_t)n * HW * pitch;
  s_sum[threadIdx.x] = 0.f; s_sq[threadIdx.x] = 0.f; s_kmax[threadIdx.x] = 0; s_kmin[threadIdx.x] = 0;
  __syncthreads();

  float shift[8], sum[8], sq[8], vmx[8], vmn[8];
  int imx[8], imn[8];
  u64 kx[8], kn[8];
  load8<F32>(y, base + c0, shift);
#pragma unroll
  for (int i = 0; i < 8; ++i) { sum[i] = 0.f; sq[i] = 0.f; vmx[i] = -INFINITY; vmn[i] = INFINITY; imx[i] = 0; imn[i] = 0; }
  constexpr int U = 2;                           // pixels per trip: their loads are issued before the first use
  const int step = 8 * gpw;
  const int64_t sy = (int64_t)step * pitch;
  int64_t off = base + (int64_t)(p_begin + warp * gpw + grp) * pitch + c0;
  for (int pb = p_begin + warp * gpw + grp; pb < p_end; pb += step * U, off += sy * U) {
    float v[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (pb + u * step < p_end) load8<F32>(y, off + u * sy, v[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + u * step;
      if (p >= p_end) break;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dlt = v[u][i] - shift[i];
        sum[i] += dlt;
        sq[i] += dlt * dlt;
        if (EXT) {
          if (v[u][i] > vmx[i]) { vmx[i] = v[u][i]; imx[i] = p; }      // p ascends inside a thread: strict > keeps the first
          if (v[u][i] < vmn[i]) { vmn[i] = v[u][i]; imn[i] = p; }
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {       // threads that saw no pixel contribute the neutral key 0
    kx[i] = vmx[i] == -INFINITY ? 0ull : make_key(vmx[i], (uint32_t)imx[i]);
    kn[i] = vmn[i] == INFINITY ? 0ull : make_key(-vmn[i], (uint32_t)imn[i]);
  }
  for (int o = G; o < 32; o <<= 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], o);
      sq[i] += __shfl_xor_sync(0xffffffffu, sq[i], o);
      if (EXT) {
        const u64 a = __shfl_xor_sync(0xffffffffu, kx[i], o), b = __shfl_xor_sync(0xffffffffu, kn[i], o);
        kx[i] = a > kx[i] ? a : kx[i];
        kn[i] = b > kn[i] ? b : kn[i];
      }
    }
  }
  ordered([&] {
    if (grp == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        atomicAdd(&s_sum[cl + i], sum[i]);
        atomicAdd(&s_sq[cl + i], sq[i]);
        if (EXT) {
          atomicMax(&s_kmax[cl + i], kx[i]);
          atomicMax(&s_kmin[cl + i], kn[i]);
        }
      }
    }
  });
  __syncthreads();
  if (threadIdx.x < Cc) {
    const int64_t o = (int64_t)n * C + blockIdx.x * 256 + threadIdx.x;
    if (psplit == 1) {
      ss[o] = make_float2(s_sum[threadIdx.x], s_sq[threadIdx.x]);
      kmax[o] = s_kmax[threadIdx.x];
      kmin[o] = s_kmin[threadIdx.x];
    } else {
      atomicAdd(&ss[o].x, s_sum[threadIdx.x]);
      atomicAdd(&ss[o].y, s_sq[threadIdx.x]);
      if (EXT) {
        atomicMax(&kmax[o], s_kmax[threadIdx.x]);
        atomicMax(&kmin[o], s_kmin[threadIdx.x]);
      }
    }
  }
}

// ---- channel-attention MLP helpers.  These run once per sample on weights that live in L2; what matters is how
// many independent loads are in flight, not arithmetic.
// h[j] = relu(W1[j,:] . a) + relu(W1[j,:] . m) (or the raw pre-activations): warp per hidden unit, every lane issues
// its whole share of the row as independent 16-byte loads.  W1 is [Cr][C] row-major, C % 128 == 0 or C in {32, 64}.
__device__ __forceinline__ void mlp_hidden(const float* __restrict__ w1, int C, int Cr, const float* s_a, const float* s_m,
                                           float* s_pa, float* s_pm) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (C >= 128 && Cr >= 2 * nw) {
    // two hidden units per trip: their row loads are all in flight together (a warp owns Cr / nw units in sequence and
    // each used to cost one L2 round trip)
    for (int j = warp; j < Cr; j += 2 * nw) {
      const int j2 = j + nw;
      const float* row = w1 + (int64_t)j * C;
      const float* row2 = w1 + (int64_t)(j2 < Cr ? j2 : j) * C;
      float pa = 0.f, pm = 0.f, qa = 0.f, qm = 0.f;
#pragma unroll 8
      for (int c = lane * 4; c < C; c += 128) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(row + c));
        const float4 u = __ldg(reinterpret_cast<const float4*>(row2 + c));
        const float4 a = make_float4(s_a[c], s_a[c + 1], s_a[c + 2], s_a[c + 3]);       // (callers' arrays are only 4-byte aligned)
        const float4 m = make_float4(s_m[c], s_m[c + 1], s_m[c + 2], s_m[c + 3]);
        pa += w.x * a.x + w.y * a.y + w.z * a.z + w.w * a.w;
        pm += w.x * m.x + w.y * m.y + w.z * m.z + w.w * m.w;
        qa += u.x * a.x + u.y * a.y + u.z * a.z + u.w * a.w;
        qm += u.x * m.x + u.y * m.y + u.z * m.z + u.w * m.w;
      }
      pa = warp_sum(pa); pm = warp_sum(pm); qa = warp_sum(qa); qm = warp_sum(qm);
      if (lane == 0) {
        s_pa[j] = pa; s_pm[j] = pm;
        if (j2 < Cr) { s_pa[j2] = qa; s_pm[j2] = qm; }
      }
    }
    return;
  }
  for (int j = warp; j < Cr; j += nw) {
    const float* row = w1 + (int64_t)j * C;
    float pa = 0.f, pm = 0.f;
    if (C >= 128) {
#pragma unroll 8
      for (int c = lane * 4; c < C; c += 128) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(row + c));
        pa += w.x * s_a[c] + w.y * s_a[c + 1] + w.z * s_a[c + 2] + w.w * s_a[c + 3];
        pm += w.x * s_m[c] + w.y * s_m[c + 1] + w.z * s_m[c + 2] + w.w * s_m[c + 3];
      }
    } else {
      for (int c = lane; c < C; c += 32) { const float w = __ldg(row + c); pa += w * s_a[c]; pm += w * s_m[c]; }
    }
    pa = warp_sum(pa); pm = warp_sum(pm);
    if (lane == 0) { s_pa[j] = pa; s_pm[j] = pm; }
  }
}
// v[c] = sum_j W2[c][j] * h[j]: one thread per row (Cr <= 64 floats = at most two full cache lines, all loads independent)
__device__ __forceinline__ void mlp_rows_dot(const float* __restrict__ w2, int C, int Cr, const float* s_h, float* s_out) {
  if (Cr >= 16 && C >= 2 * (int)blockDim.x) {
    // two rows per trip (2 x Cr/4 independent 16-byte loads in flight)
    for (int c = threadIdx.x; c < C; c += 2 * blockDim.x) {
      const int c2 = c + blockDim.x;
      const float* row = w2 + (int64_t)c * Cr;
      const float* row2 = w2 + (int64_t)(c2 < C ? c2 : c) * Cr;
      float v = 0.f, v2 = 0.f;
#pragma unroll 16
      for (int j = 0; j < Cr; j += 4) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(row + j));
        const float4 u = __ldg(reinterpret_cast<const float4*>(row2 + j));
        const float4 h = make_float4(s_h[j], s_h[j + 1], s_h[j + 2], s_h[j + 3]);
        v += w.x * h.x + w.y * h.y + w.z * h.z + w.w * h.w;
        v2 += u.x * h.x + u.y * h.y + u.z * h.z + u.w * h.w;
      }
      s_out[c] = v;
      if (c2 < C) s_out[c2] = v2;
    }
    return;
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* row = w2 + (int64_t)c * Cr;
    float v = 0.f;
    if (Cr >= 4) {
#pragma unroll 16
      for (int j = 0; j < Cr; j += 4) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(row + j));
        v += w.x * s_h[j] + w.y * s_h[j + 1] + w.z * s_h[j + 2] + w.w * s_h[j + 3];
      }
    } else {
      for (int j = 0; j < Cr; ++j) v += __ldg(row + j) * s_h[j];
    }
    s_out[c] = v;
  }
}
// dh[j] += sum_c W2[c][j] * dv[c]   (s_dh must be zero on entry; Cr <= 64): lanes over j, warps over rows
__device__ __forceinline__ void mlp_cols_dot(const float* __restrict__ w2, int C, int Cr, const float* s_dv, float* s_dh) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float a0 = 0.f, a1 = 0.f;
  const bool l0 = lane < Cr, l1 = lane + 32 < Cr;
#pragma unroll 8
  for (int c = warp; c < C; c += nw) {
    const float dv = s_dv[c];
    if (l0) a0 += __ldg(w2 + (int64_t)c * Cr + lane) * dv;
    if (l1) a1 += __ldg(w2 + (int64_t)c * Cr + lane + 32) * dv;
  }
  if (l0) atomicAdd(&s_dh[lane], a0);
  if (l1) atomicAdd(&s_dh[lane + 32], a1);
}

// ---------------------------------------------------------------------------------------------------
// forward 2: per-sample coefficients (mean, rstd, a, b), max-pooled value, channel-attention MLP -> gc
// grid N, block 256, dynamic smem (2*C + 64) floats
// ---------------------------------------------------------------------------------------------------
template <bool F32>
__global__ void __launch_bounds__(256) nb_coef_kernel(const void* __restrict__ y, int pitch, int HW, int C,
                                                      const float2* __restrict__ ss, const u64* __restrict__ kmax,
                                                      const u64* __restrict__ kmin, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, float eps, int has_cbam,
                                                      int Cr, const float* __restrict__ w1,
                                                      const float* __restrict__ w2, float* __restrict__ nc,
                                                      int32_t* __restrict__ nc_idx, int fused, int N) {
  extern __shared__ float sm[];
  float* s_avg = sm;          // [C]
  float* s_mx = sm + C;       // [C]
  float* s_h = sm + 2 * C;    // [Cr]
  const int n = blockIdx.x;
  const int64_t base = (int64_t)n * HW * pitch;
  const float inv = 1.f / (float)HW;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int64_t o = (int64_t)n * C + c;
    float mean, rstd, yext;
    uint32_t idx;
    const float g = gamma[c], b0 = beta[c];
    if (fused) {
      // statistics accumulated by the convolution epilogue: double sum / sumsq, order-preserving max / min keys
      const int64_t NCt = (int64_t)N * C;
      const double* sd = (const double*)ss;
      const double m = sd[o] * (double)inv;
      const double var = fmax(sd[NCt + o] * (double)inv - m * m, 0.0);
      mean = (float)m;
      rstd = rsqrtf((float)var + eps);
      const uint32_t* kk = (const uint32_t*)(sd + 2 * NCt);
      yext = (g * rstd >= 0.f) ? ord2f(kk[o]) : -ord2f(kk[NCt + o]);
      idx = 0x7fffffffu;            // recovered by nb_pool (first pixel whose value equals yext)
    } else {
      const float shift = F32 ? ((const float*)y)[base + c] : bf2f(((const bf16*)y)[base + c]);
      const float2 s = ss[o];
      const float md = s.x * inv;
      mean = shift + md;
      const float var = fmaxf(s.y * inv - md * md, 0.f);
      rstd = rsqrtf(var + eps);
      if (!has_cbam) { yext = mean; idx = 0; }          // the extrema only feed the channel attention
      else if (g * rstd >= 0.f) { const u64 k = kmax[o]; yext = key_val(k); idx = key_idx(k); }
      else { const u64 k = kmin[o]; yext = -key_val(k); idx = key_idx(k); }
    }
    const float a = g * rstd, b = b0 - mean * a;
    const float ext_uhat = (yext - mean) * rstd;
    const float ext_u = g * ext_uhat + b0;
    float* q = nc + o * NC_W;
    q[NC_MEAN] = mean; q[NC_RSTD] = rstd; q[NC_A] = a; q[NC_B] = b;
    q[NC_EXTU] = ext_u; q[NC_EXTUHAT] = ext_uhat; q[NC_GC] = 1.f; q[NC_SPARE] = yext;
    nc_idx[o] = (int32_t)idx;
    s_avg[c] = b0;      // mean over H*W of an instance-normalised map is exactly beta
    s_mx[c] = ext_u;
  }
  if (!has_cbam) return;
  __syncthreads();
  float* s_pm = s_h + 64;
  mlp_hidden(w1, C, Cr, s_avg, s_mx, s_h, s_pm);
  __syncthreads();
  if (threadIdx.x < Cr) s_h[threadIdx.x] = fmaxf(s_h[threadIdx.x], 0.f) + fmaxf(s_pm[threadIdx.x], 0.f);   // W2 is linear
  __syncthreads();
  mlp_rows_dot(w2, C, Cr, s_h, s_avg);            // s_avg is free now: reuse it for the pre-sigmoid gate
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x)
    nc[((int64_t)n * C + c) * NC_W + NC_GC] = 1.f / (1.f + expf(-s_avg[c]));
}

// ---------------------------------------------------------------------------------------------------
// forward 3: sweep raw y: write uhat (bf16); with CBAM: per-pixel mean_c / max_c / argmax_c of u*gc;
// without CBAM: write out = act(u) directly.      grid (pchunks, N), block 256, dyn smem 5*C floats
// ---------------------------------------------------------------------------------------------------
template <bool F32, int ITERS>
__global__ void __launch_bounds__(256) nb_pool_kernel(const void* __restrict__ y, int y_pitch, int HW, int C,
                                                      const float* __restrict__ nc, int has_cbam, float slope,
                                                      bf16* __restrict__ uhat, bf16* __restrict__ out,
                                                      int out_pitch, float* __restrict__ sa,
                                                      int32_t* __restrict__ cidx, int ppc, int fused,
                                                      int32_t* __restrict__ nc_idx) {
  extern __shared__ float sm[];
  float* s_mean = sm; float* s_rstd = sm + C; float* s_a = sm + 2 * C; float* s_b = sm + 3 * C; float* s_gc = sm + 4 * C;
  float* s_yext = sm + 5 * C;      // raw value of the max-pooled element (fused statistics: its pixel index is found here)
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* q = nc + ((int64_t)n * C + c) * NC_W;
    s_mean[c] = q[NC_MEAN]; s_rstd[c] = q[NC_RSTD]; s_a[c] = q[NC_A]; s_b[c] = q[NC_B]; s_gc[c] = q[NC_GC];
    s_yext[c] = q[NC_SPARE];
  }
  __syncthreads();
  const int G = min(32, C / 8);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % G, grp = lane / G, gpw = 32 / G;
  float h_mean[8];
  float h_rstd[8];
  float h_a[8];
  float h_b[8];
  float h_gc[8];
  if (ITERS == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      h_mean[i] = s_mean[sub * 8 + i];
      h_rstd[i] = s_rstd[sub * 8 + i];
      h_a[i] = s_a[sub * 8 + i];
      h_b[i] = s_b[sub * 8 + i];
      h_gc[i] = s_gc[sub * 8 + i];
    }
  }
  const int p0 = blockIdx.x * ppc, p_end = min(HW, p0 + ppc);
  const int64_t ybase = (int64_t)n * HW * y_pitch, ubase = (int64_t)n * HW * C, obase = (int64_t)n * HW * out_pitch;
  constexpr int U = ITERS == 1 ? 2 : 1;          // pixels per lane group and trip (loads first, then math)
  for (int pb = p0 + warp * gpw; pb < p_end; pb += 8 * gpw * U) {
    float vv[U][ITERS][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + grp + u * 8 * gpw;
      if (p < p_end) {
#pragma unroll
        for (int it = 0; it < ITERS; ++it) load8<F32>(y, ybase + (int64_t)p * y_pitch + (it * G + sub) * 8, vv[u][it]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + grp + u * 8 * gpw;
      const bool valid = p < p_end;
      float sum = 0.f, mx = -INFINITY;
      int mxc = 0;
      if (valid) {
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
          const int c = (it * G + sub) * 8;
          const float* p_mean = (ITERS == 1) ? h_mean : (s_mean + c);
          const float* p_rstd = (ITERS == 1) ? h_rstd : (s_rstd + c);
          const float* p_a = (ITERS == 1) ? h_a : (s_a + c);
          const float* p_b = (ITERS == 1) ? h_b : (s_b + c);
          const float* p_gc = (ITERS == 1) ? h_gc : (s_gc + c);
          const float* v = vv[u][it];
          float uh[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) uh[i] = (v[i] - p_mean[i]) * p_rstd[i];
          if (uhat != nullptr) stg8(uhat + ubase + (int64_t)p * C + c, pack8(uh));
          if (has_cbam) {
            if (fused) {
#pragma unroll
              for (int i = 0; i < 8; ++i)
                if (v[i] == s_yext[c + i] && p < nc_idx[(int64_t)n * C + c + i])            // (a stale read only costs an atomic)
                  atomicMin(&nc_idx[(int64_t)n * C + c + i], p);                              // first index wins
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float u1 = (p_a[i] * v[i] + p_b[i]) * p_gc[i];
              sum += u1;
              if (u1 > mx) { mx = u1; mxc = c + i; }
            }
          } else {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = act_fwd(p_a[i] * v[i] + p_b[i], slope);
            stg8(out + obase + (int64_t)p * out_pitch + c, pack8(o));
          }
        }
      }
      if (has_cbam) {
        for (int o = G >> 1; o > 0; o >>= 1) {
          sum += __shfl_xor_sync(0xffffffffu, sum, o);
          const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
          const int omc = __shfl_xor_sync(0xffffffffu, mxc, o);
          if (omx > mx || (omx == mx && omc < mxc)) { mx = omx; mxc = omc; }
        }
        if (valid && sub == 0) {
          const int64_t o = (int64_t)n * HW + p;
          sa[o * 2] = sum / (float)C;
          sa[o * 2 + 1] = mx;
          cidx[o] = mxc;
        }
      }
    }
  }
}

// 3x3 attention conv on the [mean, max] map (zero padding), one pixel
__device__ __forceinline__ float sa_conv(const float* __restrict__ sa_n, const float* __restrict__ wsp, int H, int W,
                                         int py, int px) {
  float q = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = py + ky - 1;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = px + kx - 1;
      if (xx < 0 || xx >= W) continue;
      const float2 v = *reinterpret_cast<const float2*>(sa_n + ((int64_t)yy * W + xx) * 2);
      q += wsp[ky * 3 + kx] * v.x + wsp[9 + ky * 3 + kx] * v.y;
    }
  }
  return q;
}

// forward 3b (CBAM only): gs = sigmoid(conv3x3_{2->1}(sa)), one thread per pixel (the map is 1/C of the tensor)
__global__ void __launch_bounds__(256) nb_gate_kernel(int H, int W, const float* __restrict__ wsp,
                                                      const float* __restrict__ sa, float* __restrict__ gs) {
  __shared__ float s_w[18];
  if (threadIdx.x < 18) s_w[threadIdx.x] = wsp[threadIdx.x];
  __syncthreads();
  const int n = blockIdx.y, HW = H * W;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  gs[(int64_t)n * HW + p] = 1.f / (1.f + expf(-sa_conv(sa + (int64_t)n * HW * 2, s_w, H, W, p / W, p % W)));
}

// ---------------------------------------------------------------------------------------------------
// forward 4 (CBAM only): out = act(r + u*gc*gs)
// grid (pchunks, N), block 256, dyn smem 3*C floats
// ---------------------------------------------------------------------------------------------------
template <bool F32, int ITERS>
__global__ void __launch_bounds__(256) nb_apply_kernel(const void* __restrict__ y, int y_pitch, int H, int W, int C,
                                                       const float* __restrict__ nc, int res_mode,
                                                       const bf16* __restrict__ res, int res_pitch, float slope,
                                                       bf16* __restrict__ out, int out_pitch,
                                                       const float* __restrict__ gs, int ppc) {
  // u is recomputed from the raw fp32 conv output (not from the bf16 uhat) so that the sign of the block output --
  // the ReLU mask the backward pass uses -- agrees with an fp32 evaluation
  extern __shared__ float sm[];
  float* s_a = sm; float* s_b = sm + C; float* s_gc = sm + 2 * C;
  const int n = blockIdx.y, HW = H * W;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float* q = nc + ((int64_t)n * C + c) * NC_W;
    s_a[c] = q[NC_A]; s_b[c] = q[NC_B]; s_gc[c] = q[NC_GC];
  }
  __syncthreads();
  const int G = min(32, C / 8);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % G, grp = lane / G, gpw = 32 / G;
  float h_a[8];
  float h_b[8];
  float h_gc[8];
  if (ITERS == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      h_a[i] = s_a[sub * 8 + i];
      h_b[i] = s_b[sub * 8 + i];
      h_gc[i] = s_gc[sub * 8 + i];
    }
  }
  const int p0 = blockIdx.x * ppc, p_end = min(HW, p0 + ppc);
  const int64_t ybase = (int64_t)n * HW * y_pitch, obase = (int64_t)n * HW * out_pitch, rbase = (int64_t)n * HW * res_pitch;
  constexpr int U = ITERS == 1 ? 2 : 1;
  for (int pb = p0 + warp * gpw; pb < p_end; pb += 8 * gpw * U) {
    float vv[U][ITERS][8], gq[U];
    bf16x8 rr8[U][ITERS];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + grp + u * 8 * gpw;
      gq[u] = 0.f;
      if (p < p_end) {
        gq[u] = gs[(int64_t)n * HW + p];
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
          const int c = (it * G + sub) * 8;
          load8<F32>(y, ybase + (int64_t)p * y_pitch + c, vv[u][it]);
          if (res_mode == 2) rr8[u][it] = ldg8(res + rbase + (int64_t)p * res_pitch + c);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pb + grp + u * 8 * gpw;
      if (p >= p_end) continue;
      const float g = gq[u];
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int c = (it * G + sub) * 8;
        const float* p_a = (ITERS == 1) ? h_a : (s_a + c);
        const float* p_b = (ITERS == 1) ? h_b : (s_b + c);
        const float* p_gc = (ITERS == 1) ? h_gc : (s_gc + c);
        const float* v = vv[u][it];
        float r[8], o[8];
        if (res_mode == 2) unpack8(rr8[u][it], r);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float uu = p_a[i] * v[i] + p_b[i];
          const float cb = uu * p_gc[i] * g;
          const float rr = res_mode == 1 ? uu : (res_mode == 2 ? r[i] : 0.f);
          o[i] = act_fwd(rr + cb, slope);
        }
        stg8(out + obase + (int64_t)p * out_pitch + c, pack8(o));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Fast forward sweeps for C <= 256 (same restructuring as the nbf_bwd* kernels below: template variants, pointer
// strides, folded per-channel constants).  uhat = v*rstd - mean*rstd, u*gc = v*(a*gc) + b*gc.
// ---------------------------------------------------------------------------------------------------
template <bool F32, bool CBAM>
__global__ void __launch_bounds__(256, 3) nbf_pool_kernel(const void* __restrict__ y, int y_pitch, int HW, int C,
                                                          const float* __restrict__ nc, float slope,
                                                          bf16* __restrict__ uhat, bf16* __restrict__ out, int out_pitch,
                                                          float* __restrict__ sa, int32_t* __restrict__ cidx, int ppc) {
  const int n = blockIdx.y, G = C >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & (G - 1), grp = lane / G, gpw = 32 / G, c = sub * 8;
  float h_r[8], h_nm[8], h_a[8], h_b[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4* q = reinterpret_cast<const float4*>(nc + ((int64_t)n * C + c + i) * NC_W);
    const float4 lo = __ldg(q);                      // {mean, rstd, a, b}
    const float gc = CBAM ? __ldg(q + 1).x : 1.f;    // {gc, ...}
    h_r[i] = lo.y; h_nm[i] = -lo.x * lo.y;
    h_a[i] = lo.z * gc; h_b[i] = lo.w * gc;
  }
  const int p_end = min(HW, (int)(blockIdx.x + 1) * ppc), step = 8 * gpw;
  int p = blockIdx.x * ppc + warp * gpw + grp;
  const int64_t q0 = (int64_t)n * HW + p;
  int64_t yo = q0 * y_pitch + c;
  bf16* pu = uhat + q0 * C + c;
  bf16* po = CBAM ? nullptr : out + q0 * out_pitch + c;
  float2* psa = reinterpret_cast<float2*>(sa) + q0;
  int32_t* pci = cidx + q0;
  const int64_t sy = (int64_t)step * y_pitch, su = (int64_t)step * C, so = (int64_t)step * out_pitch;
  const float invC = 1.f / (float)C;
  for (; p < p_end + grp; p += 2 * step) {            // "+ grp": all lanes of a warp run the same trips (shuffles)
    const bool vld[2] = {p < p_end, p + step < p_end};
    float v[2][8];
    if (vld[0]) load8<F32>(y, yo, v[0]);
    if (vld[1]) load8<F32>(y, yo + sy, v[1]);
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float sum = 0.f, mx = -INFINITY;
      int mxc = 0;
      if (vld[u]) {
        float uh[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) uh[i] = fmaf(v[u][i], h_r[i], h_nm[i]);
        if (uhat != nullptr) stg8(pu + u * su, pack8(uh));      // inference: nothing is saved
        if (CBAM) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float u1 = fmaf(h_a[i], v[u][i], h_b[i]);
            sum += u1;
            if (u1 > mx) { mx = u1; mxc = c + i; }
          }
        } else {
          float o[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = act_fwd(fmaf(h_a[i], v[u][i], h_b[i]), slope);
          stg8(po + u * so, pack8(o));
        }
      }
      if (CBAM) {
        for (int o = G >> 1; o > 0; o >>= 1) {
          sum += __shfl_xor_sync(0xffffffffu, sum, o);
          const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
          const int omc = __shfl_xor_sync(0xffffffffu, mxc, o);
          if (omx > mx || (omx == mx && omc < mxc)) { mx = omx; mxc = omc; }
        }
        if (vld[u] && sub == 0) {
          psa[u * step] = make_float2(sum * invC, mx);
          pci[u * step] = mxc;
        }
      }
    }
    yo += 2 * sy; pu += 2 * su; psa += 2 * step; pci += 2 * step;
    if (!CBAM) po += 2 * so;
  }
}

// out = act(r + u*gc*gs),  r = u (MODE 1) | res (MODE 2) | 0 (MODE 3)
template <bool F32, int MODE>
__global__ void __launch_bounds__(256, 3) nbf_apply_kernel(const void* __restrict__ y, int y_pitch, int HW, int C,
                                                           const float* __restrict__ nc, const bf16* __restrict__ res,
                                                           int res_pitch, float slope, bf16* __restrict__ out,
                                                           int out_pitch, const float* __restrict__ gs, int ppc) {
  const int n = blockIdx.y, G = C >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & (G - 1), grp = lane / G, gpw = 32 / G, c = sub * 8;
  float h_a[8], h_b[8], h_gc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4* q = reinterpret_cast<const float4*>(nc + ((int64_t)n * C + c + i) * NC_W);
    const float4 lo = __ldg(q);                      // {mean, rstd, a, b}
    h_a[i] = lo.z; h_b[i] = lo.w; h_gc[i] = __ldg(q + 1).x;
  }
  const int p_end = min(HW, (int)(blockIdx.x + 1) * ppc), step = 8 * gpw;
  int p = blockIdx.x * ppc + warp * gpw + grp;
  const int64_t q0 = (int64_t)n * HW + p;
  int64_t yo = q0 * y_pitch + c;
  bf16* po = out + q0 * out_pitch + c;
  const bf16* pr = MODE == 2 ? res + q0 * res_pitch + c : nullptr;
  const float* pg = gs + q0;
  const int64_t sy = (int64_t)step * y_pitch, so = (int64_t)step * out_pitch, sr = (int64_t)step * res_pitch;
  for (; p < p_end; p += 2 * step) {
    const bool v1 = p + step < p_end;
    float v[2][8], g[2];
    bf16x8 rr[2];
    load8<F32>(y, yo, v[0]);
    g[0] = pg[0];
    if (MODE == 2) rr[0] = ldg8(pr);
    if (v1) {
      load8<F32>(y, yo + sy, v[1]);
      g[1] = pg[step];
      if (MODE == 2) rr[1] = ldg8(pr + sr);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u && !v1) break;
      float r[8], o[8];
      if (MODE == 2) unpack8(rr[u], r);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float uu = fmaf(h_a[i], v[u][i], h_b[i]);
        const float k = h_gc[i] * g[u];
        const float t = MODE == 1 ? fmaf(uu, k, uu) : (MODE == 2 ? fmaf(uu, k, r[i]) : uu * k);
        o[i] = act_fwd(t, slope);
      }
      stg8(po + u * so, pack8(o));
    }
    yo += 2 * sy; po += 2 * so; pg += 2 * step;
    if (MODE == 2) pr += 2 * sr;
  }
}

// ---------------------------------------------------------------------------------------------------
// backward helpers
// ---------------------------------------------------------------------------------------------------
// reduce per-lane per-channel accumulators over the pixel groups of the warp, then the CTA, then one global
// atomic per channel.  acc[it][i] belongs to channel (it*G + sub)*8 + i.
template <int ITERS>
__device__ __forceinline__ void flush_nc(float (&acc)[ITERS][8], int G, int sub, int grp, int C, float* s_buf,
                                         float* __restrict__ dst /* element stride BN_W */, int field) {
  for (int o = G; o < 32; o <<= 1)
#pragma unroll
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[it][i] += __shfl_xor_sync(0xffffffffu, acc[it][i], o);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) s_buf[c] = 0.f;
  __syncthreads();
  if (grp == 0)
#pragma unroll
    for (int it = 0; it < ITERS; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&s_buf[(it * G + sub) * 8 + i], acc[it][i]);
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += blockDim.x) atomicAdd(dst + (int64_t)c * BN_W + field, s_buf[c]);
}

// backward 1 (CBAM only): dgs[p] = sum_c ds*u*gc -> dq ; dgc[n,c] += sum_p ds*u*gs
template <int ITERS>
__global__ void __launch_bounds__(256) nb_bwd1_kernel(const bf16* __restrict__ dout, int dout_pitch,
                                                      const bf16* __restrict__ out, int out_pitch,
                                                      const bf16* __restrict__ uhat, int HW, int C,
                                                      const float* __restrict__ nc, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, const float* __restrict__ gs,
                                                      float slope, float* __restrict__ bwd_nc,
                                                      float* __restrict__ bwd_px, int ppc) {
  extern __shared__ float sm[];
  float* s_g = sm; float* s_b = sm + C; float* s_gc = sm + 2 * C; float* s_buf = sm + 3 * C;
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s_g[c] = gamma[c]; s_b[c] = beta[c];
    s_gc[c] = nc[((int64_t)n * C + c) * NC_W + NC_GC];
  }
  __syncthreads();
  const int G = min(32, C / 8);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % G, grp = lane / G, gpw = 32 / G;
  float h_g[8];
  float h_b[8];
  float h_gc[8];
  if (ITERS == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      h_g[i] = s_g[sub * 8 + i];
      h_b[i] = s_b[sub * 8 + i];
      h_gc[i] = s_gc[sub * 8 + i];
    }
  }
  const int p0 = blockIdx.x * ppc, p_end = min(HW, p0 + ppc);
  const int64_t ubase = (int64_t)n * HW * C, obase = (int64_t)n * HW * out_pitch, dbase = (int64_t)n * HW * dout_pitch;
  float acc[ITERS][8];
#pragma unroll
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[it][i] = 0.f;
  // U pixels per lane group and trip: all their 16-byte loads are issued before the first use (bytes in flight)
  constexpr int U = ITERS == 1 ? 2 : 1;
  for (int pb = p0 + warp * gpw; pb < p_end; pb += 8 * gpw * U) {
    int pp[U];
    bool valid[U];
    float gq[U];
    bf16x8 ru[U][ITERS], ro[U][ITERS], rd[U][ITERS];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      pp[u] = pb + grp + u * 8 * gpw;
      valid[u] = pp[u] < p_end;
      gq[u] = 0.f;
      if (valid[u]) {
        gq[u] = gs[(int64_t)n * HW + pp[u]];
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
          const int c = (it * G + sub) * 8;
          ru[u][it] = ldg8(uhat + ubase + (int64_t)pp[u] * C + c);
          ro[u][it] = ldg8(out + obase + (int64_t)pp[u] * out_pitch + c);
          rd[u][it] = ldg8(dout + dbase + (int64_t)pp[u] * dout_pitch + c);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int p = pp[u];
      const float g = gq[u];
      float dgs = 0.f;
      if (valid[u]) {
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
          const int c = (it * G + sub) * 8;
          const float* p_g = (ITERS == 1) ? h_g : (s_g + c);
          const float* p_b = (ITERS == 1) ? h_b : (s_b + c);
          const float* p_gc = (ITERS == 1) ? h_gc : (s_gc + c);
          float uh[8], o[8], d[8];
          unpack8(ru[u][it], uh);
          unpack8(ro[u][it], o);
          unpack8(rd[u][it], d);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float ds = d[i] * (o[i] > 0.f ? 1.f : slope);
            const float t = ds * (p_g[i] * uh[i] + p_b[i]);
            dgs += t * p_gc[i];
            acc[it][i] += t * g;
          }
        }
      }
      for (int o = G >> 1; o > 0; o >>= 1) dgs += __shfl_xor_sync(0xffffffffu, dgs, o);
      if (valid[u] && sub == 0) bwd_px[((int64_t)n * HW + p) * BP_W + BP_DQ] = dgs * g * (1.f - g);
    }
  }
  flush_nc<ITERS>(acc, G, sub, grp, C, s_buf, bwd_nc + (int64_t)n * C * BN_W, BN_DGC);
}

// backward 1b (CBAM only): dsa = conv3x3^T(dq), dwsp += sum dq * sa(shifted).  one thread per pixel
__global__ void __launch_bounds__(256) nb_bwd_sp_kernel(int H, int W, const float* __restrict__ wsp,
                                                        const float* __restrict__ sa, float* __restrict__ bwd_px,
                                                        float* __restrict__ dwsp) {
  __shared__ float s_w[18];
  __shared__ float s_red[18];
  if (threadIdx.x < 18) { s_w[threadIdx.x] = wsp[threadIdx.x]; s_red[threadIdx.x] = 0.f; }
  __syncthreads();
  const int n = blockIdx.y, HW = H * W;
  const float* sa_n = sa + (int64_t)n * HW * 2;
  float* px_n = bwd_px + (int64_t)n * HW * BP_W;
  float part[18];
#pragma unroll
  for (int i = 0; i < 18; ++i) part[i] = 0.f;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += gridDim.x * blockDim.x) {
    const int py = p / W, px = p % W;
    const float dq = px_n[(int64_t)p * BP_W + BP_DQ];
    float dmean = 0.f, dmax = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky)
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        // q[p'] uses sa[p' + (ky-1, kx-1)]:  dsa[p] += w[ky][kx] * dq[p - (ky-1, kx-1)]
        const int yy = py - (ky - 1), xx = px - (kx - 1);
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
          const float dqq = px_n[((int64_t)yy * W + xx) * BP_W + BP_DQ];
          dmean += s_w[ky * 3 + kx] * dqq;
          dmax += s_w[9 + ky * 3 + kx] * dqq;
        }
        // dw[ky][kx] += dq[p] * sa[p + (ky-1, kx-1)]
        const int y2 = py + ky - 1, x2 = px + kx - 1;
        if (y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) {
          const float2 v = *reinterpret_cast<const float2*>(sa_n + ((int64_t)y2 * W + x2) * 2);
          part[ky * 3 + kx] += dq * v.x;
          part[9 + ky * 3 + kx] += dq * v.y;
        }
      }
    px_n[(int64_t)p * BP_W + BP_DMEAN] = dmean;
    px_n[(int64_t)p * BP_W + BP_DMAX] = dmax;
  }
#pragma unroll
  for (int i = 0; i < 18; ++i) {
    const float v = warp_sum(part[i]);
    if ((threadIdx.x & 31) == 0) atomicAdd(&s_red[i], v);
  }
  __syncthreads();
  if (threadIdx.x < 18) atomicAdd(dwsp + threadIdx.x, s_red[threadIdx.x]);
}

// gradient wrt u for 8 channels of one pixel (shared by the reduction pass and the final pass, so that du is
// never rounded to bf16 before the InstanceNorm projection removes its common-mode part)
template <bool NEED_U>
__device__ __forceinline__ void nb_du8(const float* uh, const float* o, const float* d, int c, const float* p_g,
                                       const float* p_b, const float* p_gc, int has_cbam, int res_mode, float slope,
                                       float g, float dmean, float dmax, int ci, float* du, float* dsv, float* dspu) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float ds = d[i] * (o[i] > 0.f ? 1.f : slope);
    dsv[i] = ds;
    float v = (res_mode == 0 || res_mode == 1) ? ds : 0.f;
    float su = 0.f;
    if (has_cbam) {
      const float gc = p_gc[i];
      const float dsp = dmean + ((c + i) == ci ? dmax : 0.f);   // grad wrt u1 = u*gc from the spatial branch
      v += ds * gc * g + dsp * gc;
      if (NEED_U) su = dsp * (p_g[i] * uh[i] + p_b[i]);
    }
    du[i] = v;
    dspu[i] = su;
  }
}

// backward 2: per-(n,c) S1 = sum du, S2 = sum du*uhat, dgc += ...; writes dres
template <int ITERS>
__global__ void __launch_bounds__(256, ITERS == 1 ? 2 : 1) nb_bwd2_kernel(const bf16* __restrict__ dout, int dout_pitch,
                                                      const bf16* __restrict__ out, int out_pitch,
                                                      const bf16* __restrict__ uhat, int HW, int C,
                                                      const float* __restrict__ nc, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, const float* __restrict__ gs,
                                                      const int32_t* __restrict__ cidx,
                                                      const float* __restrict__ bwd_px, int has_cbam, int res_mode,
                                                      float slope, bf16* __restrict__ dy, int dy_pitch,
                                                      bf16* __restrict__ dres, int dres_pitch,
                                                      float* __restrict__ bwd_nc, int ppc) {
  extern __shared__ float sm[];
  float* s_g = sm; float* s_b = sm + C; float* s_gc = sm + 2 * C; float* s_buf = sm + 3 * C;
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    s_g[c] = gamma[c]; s_b[c] = beta[c];
    s_gc[c] = nc[((int64_t)n * C + c) * NC_W + NC_GC];
  }
  __syncthreads();
  const int G = min(32, C / 8);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % G, grp = lane / G, gpw = 32 / G;
  float h_g[8];
  float h_b[8];
  float h_gc[8];
  if (ITERS == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      h_g[i] = s_g[sub * 8 + i];
      h_b[i] = s_b[sub * 8 + i];
      h_gc[i] = s_gc[sub * 8 + i];
    }
  }
  const int p0 = blockIdx.x * ppc, p_end = min(HW, p0 + ppc);
  const int64_t ubase = (int64_t)n * HW * C, obase = (int64_t)n * HW * out_pitch, dbase = (int64_t)n * HW * dout_pitch;
  const int64_t rbase = (int64_t)n * HW * dres_pitch;
  const float invC = 1.f / (float)C;
  float a1[ITERS][8], a2[ITERS][8], a3[ITERS][8];
#pragma unroll
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i) { a1[it][i] = 0.f; a2[it][i] = 0.f; a3[it][i] = 0.f; }
  constexpr int U = ITERS == 1 ? 2 : 1;
  for (int pb = p0 + warp * gpw; pb < p_end; pb += 8 * gpw * U) {
    int pp[U], ciq[U];
    bool valid[U];
    float gq[U], dmeanq[U], dmaxq[U];
    bf16x8 ru[U][ITERS], ro[U][ITERS], rd[U][ITERS];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      pp[u] = pb + grp + u * 8 * gpw;
      valid[u] = pp[u] < p_end;
      gq[u] = 0.f; dmeanq[u] = 0.f; dmaxq[u] = 0.f; ciq[u] = -1;
      if (valid[u]) {
        if (has_cbam) {
          const int64_t o = (int64_t)n * HW + pp[u];
          gq[u] = gs[o];
          dmeanq[u] = bwd_px[o * BP_W + BP_DMEAN] * invC;
          dmaxq[u] = bwd_px[o * BP_W + BP_DMAX];
          ciq[u] = cidx[o];
        }
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
          const int c = (it * G + sub) * 8;
          ru[u][it] = ldg8(uhat + ubase + (int64_t)pp[u] * C + c);
          ro[u][it] = ldg8(out + obase + (int64_t)pp[u] * out_pitch + c);
          rd[u][it] = ldg8(dout + dbase + (int64_t)pp[u] * dout_pitch + c);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!valid[u]) continue;
      const int p = pp[u];
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int c = (it * G + sub) * 8;
        const float* p_g = (ITERS == 1) ? h_g : (s_g + c);
        const float* p_b = (ITERS == 1) ? h_b : (s_b + c);
        const float* p_gc = (ITERS == 1) ? h_gc : (s_gc + c);
        float uh[8], o[8], d[8], du[8], dsv[8], dspu[8];
        unpack8(ru[u][it], uh);
        unpack8(ro[u][it], o);
        unpack8(rd[u][it], d);
        nb_du8<true>(uh, o, d, c, p_g, p_b, p_gc, has_cbam, res_mode, slope, gq[u], dmeanq[u], dmaxq[u], ciq[u], du, dsv, dspu);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a1[it][i] += du[i];
          a2[it][i] += du[i] * uh[i];
          a3[it][i] += dspu[i];
        }
        if (res_mode == 2 && dres) stg8(dres + rbase + (int64_t)p * dres_pitch + c, pack8(dsv));
      }
    }
  }
  float* dst = bwd_nc + (int64_t)n * C * BN_W;
  flush_nc<ITERS>(a1, G, sub, grp, C, s_buf, dst, BN_S1);
  flush_nc<ITERS>(a2, G, sub, grp, C, s_buf, dst, BN_S2);
  if (has_cbam) flush_nc<ITERS>(a3, G, sub, grp, C, s_buf, dst, BN_DGC);
}

// backward 3: per-sample: channel-MLP backward, dgamma/dbeta, and the IN-backward means m1, m2.
// Overwrites bwd_nc[n][c] = {d_mx (grad wrt the max-pooled u), m1, m2, -}.   grid N, block 256
__global__ void __launch_bounds__(256) nb_bwd_coef_kernel(int HW, int C, int has_cbam, int Cr,
                                                          const float* __restrict__ nc,
                                                          const float* __restrict__ beta,
                                                          const float* __restrict__ w1, const float* __restrict__ w2,
                                                          float* __restrict__ bwd_nc, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta, float* __restrict__ bwd_h) {
  extern __shared__ float sm[];
  float* s_avg = sm;            // [C]
  float* s_mx = sm + C;         // [C]
  float* s_dv = sm + 2 * C;     // [C]
  float* s_ha = sm + 3 * C;     // [Cr] pre-activation hidden (avg branch)
  float* s_hm = s_ha + 64;      // [Cr]
  float* s_dh = s_hm + 64;      // [Cr]
  const int n = blockIdx.x;
  const float inv = 1.f / (float)HW;
  float* bn = bwd_nc + (int64_t)n * C * BN_W;
  const float* q0 = nc + (int64_t)n * C * NC_W;
  if (has_cbam) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float gc = q0[c * NC_W + NC_GC];
      s_avg[c] = beta[c];
      s_mx[c] = q0[c * NC_W + NC_EXTU];
      s_dv[c] = bn[c * BN_W + BN_DGC] * gc * (1.f - gc);
    }
    if (threadIdx.x < 64) s_dh[threadIdx.x] = 0.f;
    __syncthreads();
    mlp_hidden(w1, C, Cr, s_avg, s_mx, s_ha, s_hm);
    mlp_cols_dot(w2, C, Cr, s_dv, s_dh);
    __syncthreads();
    // the weight gradients dW2[c][j] = sum_n dv[n,c] * (relu(ha)+relu(hm))[n,j] and
    // dW1[j][c] = sum_n dh[n,j] * ([ha>0] avg_c + [hm>0] mx[n,c]) are reduced over samples by nb_bwd_w_kernel;
    // here only the per-sample factors are written out (no per-sample atomics on the weight gradients)
    for (int j = threadIdx.x; j < Cr; j += blockDim.x) {
      float* hq = bwd_h + (int64_t)n * 192;
      hq[j] = fmaxf(s_ha[j], 0.f) + fmaxf(s_hm[j], 0.f);
      hq[64 + j] = s_ha[j] > 0.f ? s_dh[j] : 0.f;
      hq[128 + j] = s_hm[j] > 0.f ? s_dh[j] : 0.f;
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) bn[c * BN_W + BN_DGC] = s_dv[c];
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float d_avg = 0.f, d_mx = 0.f;
    if (has_cbam) {
      for (int j = 0; j < Cr; ++j) {
        const float w = w1[(int64_t)j * C + c];
        if (s_ha[j] > 0.f) d_avg += w * s_dh[j];
        if (s_hm[j] > 0.f) d_mx += w * s_dh[j];
      }
    }
    const float ext_uhat = q0[c * NC_W + NC_EXTUHAT];
    const float S1 = bn[c * BN_W + BN_S1] + d_mx;
    const float S2 = bn[c * BN_W + BN_S2] + d_mx * ext_uhat;
    atomicAdd(dbeta + c, S1 + d_avg);
    atomicAdd(dgamma + c, S2);
    bn[c * BN_W + BN_DMX] = d_mx;
    bn[c * BN_W + BN_S1] = S1 * inv;   // m1
    bn[c * BN_W + BN_S2] = S2 * inv;   // m2
  }
}

// channel-attention MLP weight gradients, reduced over the batch: one thread per weight element, samples split
// over blockIdx.y
__global__ void __launch_bounds__(256) nb_bwd_w_kernel(int N, int C, int Cr, const float* __restrict__ nc,
                                                       const float* __restrict__ beta,
                                                       const float* __restrict__ bwd_nc,
                                                       const float* __restrict__ bwd_h, float* __restrict__ dw1,
                                                       float* __restrict__ dw2, int n_per_block) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= C * Cr) return;
  const int c = e % C, j = e / C;
  const int n0 = blockIdx.y * n_per_block, n1 = min(N, n0 + n_per_block);
  const float avg = beta[c];
  float a1 = 0.f, a2 = 0.f;
  for (int n = n0; n < n1; ++n) {
    const float* hq = bwd_h + (int64_t)n * 192;
    const float dv = bwd_nc[((int64_t)n * C + c) * BN_W + BN_DGC];
    const float mx = nc[((int64_t)n * C + c) * NC_W + NC_EXTU];
    a2 += dv * hq[j];
    a1 += hq[64 + j] * avg + hq[128 + j] * mx;
  }
  atomicAdd(dw2 + (int64_t)c * Cr + j, a2);
  atomicAdd(dw1 + (int64_t)j * C + c, a1);
}

// The same reduction tiled through shared memory: a CTA owns 64 channels x all hidden units for WB_NS samples; the
// per-(n,c) inputs are read once as float4 rows and the per-sample hidden vectors once per CTA (the kernel above
// re-reads every (n,c) value Cr times with 4- and 8-float strides).   grid (C/64, ceil(N/WB_NS)), block 256, C % 64 == 0
constexpr int WB_NS = 64;
__global__ void __launch_bounds__(256) nb_bwd_w_tiled_kernel(int N, int C, int Cr, const float* __restrict__ nc,
                                                             const float* __restrict__ beta,
                                                             const float* __restrict__ bwd_nc,
                                                             const float* __restrict__ bwd_h, float* __restrict__ dw1,
                                                             float* __restrict__ dw2) {
  constexpr int SUB = 16;                        // samples staged per round
  __shared__ float s_dv[SUB][64], s_mx[SUB][64], s_h[SUB][192];
  const int c0 = blockIdx.x * 64, t = threadIdx.x;
  const int cl = t & 63, jq = t >> 6;            // 4 thread groups share the hidden units
  const int jn = (Cr + 3) / 4, j0 = jq * jn;     // hidden units [j0, j0 + jn) of this thread (jn <= 16)
  const int n_begin = blockIdx.y * WB_NS, n_end = min(N, n_begin + WB_NS);
  const float avg = beta[c0 + cl];
  float a1[16], a2[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) { a1[i] = 0.f; a2[i] = 0.f; }
  for (int nb = n_begin; nb < n_end; nb += SUB) {
    const int ns = min(SUB, n_end - nb);
    __syncthreads();
    for (int i = t; i < ns * 64; i += 256) {
      const int s = i >> 6, c = i & 63;
      const int64_t o = (int64_t)(nb + s) * C + c0 + c;
      s_dv[s][c] = bwd_nc[o * BN_W + BN_DGC];
      s_mx[s][c] = nc[o * NC_W + NC_EXTU];
    }
    for (int i = t; i < ns * 192; i += 256) s_h[i / 192][i % 192] = bwd_h[(int64_t)nb * 192 + i];
    __syncthreads();
    for (int s = 0; s < ns; ++s) {
      const float dv = s_dv[s][cl], mx = s_mx[s][cl];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        if (i < jn && j0 + i < Cr) {
          a2[i] = fmaf(dv, s_h[s][j0 + i], a2[i]);
          a1[i] = fmaf(s_h[s][128 + j0 + i], mx, fmaf(s_h[s][64 + j0 + i], avg, a1[i]));
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (i < jn && j0 + i < Cr) {
      atomicAdd(dw2 + (int64_t)(c0 + cl) * Cr + j0 + i, a2[i]);
      atomicAdd(dw1 + (int64_t)(j0 + i) * C + c0 + cl, a1[i]);
    }
  }
}

// backward 4: dy = a * (du + [p == argmax] d_mx - m1 - uhat*m2); du is recomputed in fp32
template <int ITERS>
__global__ void __launch_bounds__(256, ITERS == 1 ? 3 : 1) nb_bwd3_kernel(const bf16* __restrict__ dout, int dout_pitch,
                                                      const bf16* __restrict__ out, int out_pitch,
                                                      const bf16* __restrict__ uhat, int HW, int C,
                                                      const float* __restrict__ nc,
                                                      const int32_t* __restrict__ nc_idx,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ gs, const int32_t* __restrict__ cidx,
                                                      const float* __restrict__ bwd_px,
                                                      const float* __restrict__ bwd_nc, int has_cbam, int res_mode,
                                                      float slope, bf16* __restrict__ dy, int dy_pitch, int ppc) {
  extern __shared__ float sm[];
  float* s_gc = sm; float* s_a = sm + C; float* s_m1 = sm + 2 * C; float* s_m2 = sm + 3 * C; float* s_dmx = sm + 4 * C;
  int* s_idx = (int*)(sm + 5 * C);
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int64_t o = (int64_t)n * C + c;
    const float a = nc[o * NC_W + NC_A];
    s_gc[c] = nc[o * NC_W + NC_GC];
    s_a[c] = a;
    s_m1[c] = a * bwd_nc[o * BN_W + BN_S1];      // pre-multiplied by a: dy = a*du + a*extra - a*m1 - uhat*(a*m2)
    s_m2[c] = a * bwd_nc[o * BN_W + BN_S2];
    s_dmx[c] = a * bwd_nc[o * BN_W + BN_DMX];
    s_idx[c] = has_cbam ? nc_idx[o] : -1;
  }
  __syncthreads();
  const int G = min(32, C / 8);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane % G, grp = lane / G, gpw = 32 / G;
  float h_gc[8];
  float h_a[8];
  float h_m1[8];
  float h_m2[8];
  float h_dmx[8];
  int h_idx[8];
  if (ITERS == 1) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      h_gc[i] = s_gc[sub * 8 + i];
      h_a[i] = s_a[sub * 8 + i];
      h_m1[i] = s_m1[sub * 8 + i];
      h_m2[i] = s_m2[sub * 8 + i];
      h_dmx[i] = s_dmx[sub * 8 + i];
      h_idx[i] = s_idx[sub * 8 + i];
    }
  }
  const int p0 = blockIdx.x * ppc, p_end = min(HW, p0 + ppc);
  const int64_t ubase = (int64_t)n * HW * C, obase = (int64_t)n * HW * out_pitch, dbase = (int64_t)n * HW * dout_pitch;
  const int64_t ybase = (int64_t)n * HW * dy_pitch;
  const float invC = 1.f / (float)C;
  constexpr int U = ITERS == 1 ? 2 : 1;
  for (int pb = p0 + warp * gpw; pb < p_end; pb += 8 * gpw * U) {
    int pp[U], ciq[U];
    bool valid[U];
    float gq[U], dmeanq[U], dmaxq[U];
    bf16x8 ru[U][ITERS], ro[U][ITERS], rd[U][ITERS];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      pp[u] = pb + grp + u * 8 * gpw;
      valid[u] = pp[u] < p_end;
      gq[u] = 0.f; dmeanq[u] = 0.f; dmaxq[u] = 0.f; ciq[u] = -1;
      if (valid[u]) {
        if (has_cbam) {
          const int64_t o = (int64_t)n * HW + pp[u];
          gq[u] = gs[o];
          dmeanq[u] = bwd_px[o * BP_W + BP_DMEAN] * invC;
          dmaxq[u] = bwd_px[o * BP_W + BP_DMAX];
          ciq[u] = cidx[o];
        }
#pragma unroll
        for (int it = 0; it < ITERS; ++it) {
          const int c = (it * G + sub) * 8;
          ru[u][it] = ldg8(uhat + ubase + (int64_t)pp[u] * C + c);
          ro[u][it] = ldg8(out + obase + (int64_t)pp[u] * out_pitch + c);
          rd[u][it] = ldg8(dout + dbase + (int64_t)pp[u] * dout_pitch + c);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (!valid[u]) continue;
      const int p = pp[u];
#pragma unroll
      for (int it = 0; it < ITERS; ++it) {
        const int c = (it * G + sub) * 8;
        const float* p_gc = (ITERS == 1) ? h_gc : (s_gc + c);
        const float* p_a = (ITERS == 1) ? h_a : (s_a + c);
        const float* p_m1 = (ITERS == 1) ? h_m1 : (s_m1 + c);
        const float* p_m2 = (ITERS == 1) ? h_m2 : (s_m2 + c);
        const float* p_dmx = (ITERS == 1) ? h_dmx : (s_dmx + c);
        const int* p_idx = (ITERS == 1) ? h_idx : (s_idx + c);
        float uh[8], o[8], d[8], du[8], dsv[8], dspu[8], r[8];
        unpack8(ru[u][it], uh);
        unpack8(ro[u][it], o);
        unpack8(rd[u][it], d);
        nb_du8<false>(uh, o, d, c, p_gc, p_gc, p_gc, has_cbam, res_mode, slope, gq[u], dmeanq[u], dmaxq[u], ciq[u], du, dsv, dspu);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float extra = (p == p_idx[i]) ? p_dmx[i] : 0.f;
          r[i] = p_a[i] * du[i] + extra - p_m1[i] - uh[i] * p_m2[i];
        }
        stg8(dy + ybase + (int64_t)p * dy_pitch + c, pack8(r));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Fast tiled backward kernels for C <= 256 (one 8-channel vector per lane, G = C/8 lanes per pixel).  Same
// mathematics as nb_bwd1/2/3_kernel, restructured for instruction count: the variant (MODE = 0 plain, else the
// CBAM residual mode 1..3) is a template parameter, every stream is a pointer advanced by a constant stride, the
// per-channel constants are folded (a*gc, a*m1, a*m2, ...) so the dense part is 2-4 FMAs per element, and the two
// sparse terms (arg-max channel of a pixel, arg-max pixel of a channel) stay out of the dense expression.
// ---------------------------------------------------------------------------------------------------
// ds = dout * act'(out): the sign of `out` is read from the raw bf16 bits (positive, non-zero <=> the word > 0)
__device__ __forceinline__ void nbf_ds8(const bf16x8& rd, const bf16x8& ro, float slope, float* ds) {
  const uint32_t dw[4] = {rd.x, rd.y, rd.z, rd.w};
  const uint32_t ow[4] = {ro.x, ro.y, ro.z, ro.w};
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    const float dlo = __uint_as_float(dw[w] << 16), dhi = __uint_as_float(dw[w] & 0xffff0000u);
    ds[2 * w] = ((int)(ow[w] << 16) > 0) ? dlo : dlo * slope;
    ds[2 * w + 1] = ((int)ow[w] > 0xffff) ? dhi : dhi * slope;
  }
}

struct NbfPix {
  float g, dmean, dmx;        // spatial gate, d/d(mean_c) / C, dmean + d/d(max_c)
  int k;                      // arg-max channel of this pixel minus the lane's first channel (matches i in [0, 8) or not)
};

template <int MODE>
__device__ __forceinline__ NbfPix nbf_pix(const float* __restrict__ gs, const float4* __restrict__ px,
                                          const int32_t* __restrict__ cidx, int64_t q, float invC, int c) {
  NbfPix r;
  r.g = 0.f; r.dmean = 0.f; r.dmx = 0.f; r.k = -1;
  if (MODE) {
    r.g = gs[q];
    const float4 v = px[q];               // {dq, dmean, dmax, -}
    r.dmean = v.y * invC;
    r.dmx = r.dmean + v.z;
    r.k = cidx[q] - c;
  }
  return r;
}

// backward 1 (CBAM only): dgs[p] = sum_c ds*u*gc -> dq ; dgc[n,c] += sum_p ds*u*gs
__global__ void __launch_bounds__(256, 3) nbf_bwd1_kernel(const bf16* __restrict__ dout, int dout_pitch,
                                                          const bf16* __restrict__ out, int out_pitch,
                                                          const bf16* __restrict__ uhat, int HW, int C,
                                                          const float* __restrict__ nc, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ gs,
                                                          float slope, float* __restrict__ bwd_nc,
                                                          float* __restrict__ bwd_px, int ppc) {
  extern __shared__ float sm[];
  const int n = blockIdx.y, G = C >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & (G - 1), grp = lane / G, gpw = 32 / G, c = sub * 8;
  float h_g[8], h_b[8], h_gc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h_g[i] = __ldg(gamma + c + i);
    h_b[i] = __ldg(beta + c + i);
    h_gc[i] = __ldg(nc + ((int64_t)n * C + c + i) * NC_W + NC_GC);
  }
  const int p_end = min(HW, (int)(blockIdx.x + 1) * ppc), step = 8 * gpw;
  int p = blockIdx.x * ppc + warp * gpw + grp;
  const int64_t q0 = (int64_t)n * HW + p;
  const bf16* pu = uhat + q0 * C + c;
  const bf16* po = out + q0 * out_pitch + c;
  const bf16* pd = dout + q0 * dout_pitch + c;
  const float* pg = gs + q0;
  float* pq = bwd_px + q0 * BP_W + BP_DQ;
  const int64_t su = (int64_t)step * C, so = (int64_t)step * out_pitch, sd = (int64_t)step * dout_pitch;
  float acc[1][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[0][i] = 0.f;
  for (; p < p_end + grp; p += 2 * step) {           // "+ grp": every lane of the warp runs the same trips (shuffles)
    const bool v0 = p < p_end, v1 = p + step < p_end;
    bf16x8 ru[2], ro[2], rd[2];
    float g[2] = {0.f, 0.f};
    if (v0) { ru[0] = ldg8(pu); ro[0] = ldg8(po); rd[0] = ldg8(pd); g[0] = pg[0]; }
    if (v1) { ru[1] = ldg8(pu + su); ro[1] = ldg8(po + so); rd[1] = ldg8(pd + sd); g[1] = pg[step]; }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const bool v = u ? v1 : v0;
      float dgs = 0.f;
      if (v) {
        float uh[8], ds[8];
        unpack8(ru[u], uh);
        nbf_ds8(rd[u], ro[u], slope, ds);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float t = ds[i] * fmaf(h_g[i], uh[i], h_b[i]);
          dgs = fmaf(t, h_gc[i], dgs);
          acc[0][i] = fmaf(t, g[u], acc[0][i]);
        }
      }
      for (int o = G >> 1; o > 0; o >>= 1) dgs += __shfl_xor_sync(0xffffffffu, dgs, o);
      if (v && sub == 0) pq[(int64_t)u * step * BP_W] = dgs * g[u] * (1.f - g[u]);
    }
    pu += 2 * su; po += 2 * so; pd += 2 * sd; pg += 2 * step; pq += (int64_t)2 * step * BP_W;
  }
  flush_nc<1>(acc, G, sub, grp, C, sm, bwd_nc + (int64_t)n * C * BN_W, BN_DGC);
}

// CBAM sites, split form of the two reduction sweeps.  With du = ds*(direct + gc*g_p) + gc*dsp_p (dsp_p = dmean_p, plus
// dmax_p on the pixel's arg-max channel) every per-(n,c) total is LINEAR in sums that either do not need the spatial
// gradients (dmean, dmax) or do not need dout / out:
//   S1  = direct*sum ds      + gc*sum ds*g      + gc*sum dsp
//   S2  = direct*sum ds*uhat + gc*sum ds*g*uhat + gc*sum dsp*uhat
//   dgc = gamma*sum ds*g*uhat + beta*sum ds*g   + gamma*sum dsp*uhat + beta*sum dsp
// nbf_bwd1x accumulates the four ds-sums in the SAME sweep that produces dq (it also writes dres), nbf_bwd2x then only
// sweeps uhat (2 B/element instead of 6) for the two dsp-sums.  Each CTA applies the per-channel factors to its partial
// sums before the atomics, so the scratch layout {dgc, S1, S2} and everything downstream stay as they were.
template <int MODE>
__global__ void __launch_bounds__(256, 2) nbf_bwd1x_kernel(const bf16* __restrict__ dout, int dout_pitch,
                                                           const bf16* __restrict__ out, int out_pitch,
                                                           const bf16* __restrict__ uhat, int HW, int C,
                                                           const float* __restrict__ nc, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, const float* __restrict__ gs,
                                                           float slope, bf16* __restrict__ dres, int dres_pitch,
                                                           float* __restrict__ bwd_nc, float* __restrict__ bwd_px, int ppc) {
  extern __shared__ float sm[];
  const int n = blockIdx.y, G = C >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & (G - 1), grp = lane / G, gpw = 32 / G, c = sub * 8;
  float h_g[8], h_b[8], h_gc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h_g[i] = __ldg(gamma + c + i);
    h_b[i] = __ldg(beta + c + i);
    h_gc[i] = __ldg(nc + ((int64_t)n * C + c + i) * NC_W + NC_GC);
  }
  const int p_end = min(HW, (int)(blockIdx.x + 1) * ppc), step = 8 * gpw;
  int p = blockIdx.x * ppc + warp * gpw + grp;
  const int64_t q0 = (int64_t)n * HW + p;
  const bf16* pu = uhat + q0 * C + c;
  const bf16* po = out + q0 * out_pitch + c;
  const bf16* pd = dout + q0 * dout_pitch + c;
  bf16* pr = (MODE == 2 && dres) ? dres + q0 * dres_pitch + c : nullptr;
  const float* pg = gs + q0;
  float* pq = bwd_px + q0 * BP_W + BP_DQ;
  const int64_t su = (int64_t)step * C, so = (int64_t)step * out_pitch, sd = (int64_t)step * dout_pitch;
  const int64_t sr = (int64_t)step * dres_pitch;
  float A1[8], A2[8], B1[8], B2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { A1[i] = 0.f; A2[i] = 0.f; B1[i] = 0.f; B2[i] = 0.f; }
  for (; p < p_end + grp; p += 2 * step) {           // "+ grp": every lane of the warp runs the same trips (shuffles)
    const bool v0 = p < p_end, v1 = p + step < p_end;
    bf16x8 ru[2], ro[2], rd[2];
    float g[2] = {0.f, 0.f};
    if (v0) { ru[0] = ldg8(pu); ro[0] = ldg8(po); rd[0] = ldg8(pd); g[0] = pg[0]; }
    if (v1) { ru[1] = ldg8(pu + su); ro[1] = ldg8(po + so); rd[1] = ldg8(pd + sd); g[1] = pg[step]; }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const bool v = u ? v1 : v0;
      float dgs = 0.f;
      if (v) {
        float uh[8], ds[8];
        unpack8(ru[u], uh);
        nbf_ds8(rd[u], ro[u], slope, ds);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float dsg = ds[i] * g[u];
          dgs = fmaf(ds[i] * fmaf(h_g[i], uh[i], h_b[i]), h_gc[i], dgs);
          A1[i] += ds[i];
          A2[i] += dsg;
          B1[i] = fmaf(ds[i], uh[i], B1[i]);
          B2[i] = fmaf(dsg, uh[i], B2[i]);
        }
        if (MODE == 2 && pr) stg8(pr + (int64_t)u * sr, pack8(ds));
      }
      for (int o = G >> 1; o > 0; o >>= 1) dgs += __shfl_xor_sync(0xffffffffu, dgs, o);
      if (v && sub == 0) pq[(int64_t)u * step * BP_W] = dgs * g[u] * (1.f - g[u]);
    }
    pu += 2 * su; po += 2 * so; pd += 2 * sd; pg += 2 * step; pq += (int64_t)2 * step * BP_W;
    if (MODE == 2 && pr) pr += 2 * sr;
  }
  float s1[1][8], s2[1][8], sg[1][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s1[0][i] = (MODE == 1 ? A1[i] : 0.f) + h_gc[i] * A2[i];
    s2[0][i] = (MODE == 1 ? B1[i] : 0.f) + h_gc[i] * B2[i];
    sg[0][i] = h_g[i] * B2[i] + h_b[i] * A2[i];
  }
  float* dst = bwd_nc + (int64_t)n * C * BN_W;
  flush_nc<1>(s1, G, sub, grp, C, sm, dst, BN_S1);
  flush_nc<1>(s2, G, sub, grp, C, sm, dst, BN_S2);
  flush_nc<1>(sg, G, sub, grp, C, sm, dst, BN_DGC);
}

__global__ void __launch_bounds__(256, 3) nbf_bwd2x_kernel(const bf16* __restrict__ uhat, int HW, int C,
                                                           const float* __restrict__ nc, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta,
                                                           const int32_t* __restrict__ cidx,
                                                           const float* __restrict__ bwd_px, float* __restrict__ bwd_nc,
                                                           int ppc) {
  extern __shared__ float sm[];
  const int n = blockIdx.y, G = C >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & (G - 1), grp = lane / G, gpw = 32 / G, c = sub * 8;
  const int p_end = min(HW, (int)(blockIdx.x + 1) * ppc), step = 8 * gpw;
  int p = blockIdx.x * ppc + warp * gpw + grp;
  const int64_t q0 = (int64_t)n * HW + p;
  const bf16* pu = uhat + q0 * C + c;
  const float4* px4 = reinterpret_cast<const float4*>(bwd_px);
  const int64_t su = (int64_t)step * C;
  const float invC = 1.f / (float)C;
  float SD[8], C1[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { SD[i] = 0.f; C1[i] = 0.f; }
  int64_t q = q0;
  for (; p < p_end; p += 4 * step, q += 4 * step) {
    bf16x8 ru[4];
    float dmean[4], dmx[4];
    int k[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      dmean[u] = 0.f; dmx[u] = 0.f; k[u] = -1;
      ru[u] = make_uint4(0, 0, 0, 0);
      if (p + u * step < p_end) {
        ru[u] = ldg8(pu + u * su);
        const float4 v = px4[q + u * step];             // {dq, dmean, dmax, -}
        dmean[u] = v.y * invC;
        dmx[u] = dmean[u] + v.z;
        k[u] = cidx[q + u * step] - c;
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float uh[8];
      unpack8(ru[u], uh);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dsp = (i == k[u]) ? dmx[u] : dmean[u];
        SD[i] += dsp;
        C1[i] = fmaf(dsp, uh[i], C1[i]);
      }
    }
    pu += 4 * su;
  }
  float s1[1][8], s2[1][8], sg[1][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float gc = __ldg(nc + ((int64_t)n * C + c + i) * NC_W + NC_GC);
    s1[0][i] = gc * SD[i];
    s2[0][i] = gc * C1[i];
    sg[0][i] = __ldg(gamma + c + i) * C1[i] + __ldg(beta + c + i) * SD[i];
  }
  float* dst = bwd_nc + (int64_t)n * C * BN_W;
  flush_nc<1>(s1, G, sub, grp, C, sm, dst, BN_S1);
  flush_nc<1>(s2, G, sub, grp, C, sm, dst, BN_S2);
  flush_nc<1>(sg, G, sub, grp, C, sm, dst, BN_DGC);
}

// backward 2: per-(n,c) S1 = sum du, S2 = sum du*uhat, dgc += sum dsp*u ; writes dres (MODE 2)
template <int MODE>
__global__ void __launch_bounds__(256, 2) nbf_bwd2_kernel(const bf16* __restrict__ dout, int dout_pitch,
                                                          const bf16* __restrict__ out, int out_pitch,
                                                          const bf16* __restrict__ uhat, int HW, int C,
                                                          const float* __restrict__ nc, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, const float* __restrict__ gs,
                                                          const int32_t* __restrict__ cidx,
                                                          const float* __restrict__ bwd_px, float slope,
                                                          bf16* __restrict__ dres, int dres_pitch,
                                                          float* __restrict__ bwd_nc, int ppc) {
  extern __shared__ float sm[];
  const int n = blockIdx.y, G = C >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & (G - 1), grp = lane / G, gpw = 32 / G, c = sub * 8;
  float h_g[8], h_b[8], h_gc[8];
  if (MODE) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      h_g[i] = __ldg(gamma + c + i);
      h_b[i] = __ldg(beta + c + i);
      h_gc[i] = __ldg(nc + ((int64_t)n * C + c + i) * NC_W + NC_GC);
    }
  }
  const int p_end = min(HW, (int)(blockIdx.x + 1) * ppc), step = 8 * gpw;
  int p = blockIdx.x * ppc + warp * gpw + grp;
  const int64_t q0 = (int64_t)n * HW + p;
  const bf16* pu = uhat + q0 * C + c;
  const bf16* po = out + q0 * out_pitch + c;
  const bf16* pd = dout + q0 * dout_pitch + c;
  bf16* pr = (MODE == 2 && dres) ? dres + q0 * dres_pitch + c : nullptr;
  const float4* px4 = reinterpret_cast<const float4*>(bwd_px);
  const int64_t su = (int64_t)step * C, so = (int64_t)step * out_pitch, sd = (int64_t)step * dout_pitch;
  const int64_t sr = (int64_t)step * dres_pitch;
  const float invC = 1.f / (float)C;
  float a1[1][8], a2[1][8], a3[1][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a1[0][i] = 0.f; a2[0][i] = 0.f; a3[0][i] = 0.f; }
  int64_t q = q0;
  for (; p < p_end; p += 2 * step, q += 2 * step) {
    const bool v1 = p + step < p_end;
    bf16x8 ru[2], ro[2], rd[2];
    NbfPix px[2];
    ru[0] = ldg8(pu); ro[0] = ldg8(po); rd[0] = ldg8(pd);
    px[0] = nbf_pix<MODE>(gs, px4, cidx, q, invC, c);
    if (v1) {
      ru[1] = ldg8(pu + su); ro[1] = ldg8(po + so); rd[1] = ldg8(pd + sd);
      px[1] = nbf_pix<MODE>(gs, px4, cidx, q + step, invC, c);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u && !v1) break;
      float uh[8], ds[8];
      unpack8(ru[u], uh);
      nbf_ds8(rd[u], ro[u], slope, ds);
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a1[0][i] += ds[i];
          a2[0][i] = fmaf(ds[i], uh[i], a2[0][i]);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // du = ds*([direct] + gc*g) + gc*dsp ,  dsp = dmean (+ dmax on the pixel's arg-max channel)
          const float dsp = (i == px[u].k) ? px[u].dmx : px[u].dmean;
          const float k1 = MODE == 1 ? fmaf(h_gc[i], px[u].g, 1.f) : h_gc[i] * px[u].g;
          const float du = fmaf(ds[i], k1, h_gc[i] * dsp);
          a1[0][i] += du;
          a2[0][i] = fmaf(du, uh[i], a2[0][i]);
          a3[0][i] = fmaf(dsp, fmaf(h_g[i], uh[i], h_b[i]), a3[0][i]);
        }
        if (MODE == 2 && pr) stg8(pr + (int64_t)u * sr, pack8(ds));
      }
    }
    pu += 2 * su; po += 2 * so; pd += 2 * sd;
    if (MODE == 2 && pr) pr += 2 * sr;
  }
  float* dst = bwd_nc + (int64_t)n * C * BN_W;
  flush_nc<1>(a1, G, sub, grp, C, sm, dst, BN_S1);
  flush_nc<1>(a2, G, sub, grp, C, sm, dst, BN_S2);
  if (MODE) flush_nc<1>(a3, G, sub, grp, C, sm, dst, BN_DGC);
}

// backward 4: dy = a*du + [p == argmax_p(c)] a*d_mx - a*m1 - uhat*a*m2
template <int MODE>
__global__ void __launch_bounds__(256, 3) nbf_bwd3_kernel(const bf16* __restrict__ dout, int dout_pitch,
                                                          const bf16* __restrict__ out, int out_pitch,
                                                          const bf16* __restrict__ uhat, int HW, int C,
                                                          const float* __restrict__ nc,
                                                          const int32_t* __restrict__ nc_idx,
                                                          const float* __restrict__ gs, const int32_t* __restrict__ cidx,
                                                          const float* __restrict__ bwd_px,
                                                          const float* __restrict__ bwd_nc, float slope,
                                                          bf16* __restrict__ dy, int dy_pitch, int ppc) {
  extern __shared__ float sm[];
  float* s_admx = sm;                     // a * d_mx per channel
  int* s_idx = (int*)(sm + C);            // arg-max pixel per channel
  const int n = blockIdx.y, G = C >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int sub = lane & (G - 1), grp = lane / G, gpw = 32 / G, c = sub * 8;
  float h_a[8], h_ag[8], h_am1[8], h_am2[8];
  int imin = 0x7fffffff, imax = -1;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t o = (int64_t)n * C + c + i;
    const float a = __ldg(nc + o * NC_W + NC_A);
    const float4 bn = __ldg(reinterpret_cast<const float4*>(bwd_nc + o * BN_W));     // {-, S1 -> m1, S2 -> m2, d_mx}
    h_a[i] = a;
    h_ag[i] = MODE ? a * __ldg(nc + o * NC_W + NC_GC) : 0.f;
    h_am1[i] = a * bn.y;
    h_am2[i] = a * bn.z;
    if (MODE) {
      const int ix = __ldg(nc_idx + o);
      imin = min(imin, ix); imax = max(imax, ix);
      if (grp == 0 && warp == 0) { s_admx[c + i] = a * bn.w; s_idx[c + i] = ix; }
    }
  }
  if (MODE) __syncthreads();
  const int p_end = min(HW, (int)(blockIdx.x + 1) * ppc), step = 8 * gpw;
  int p = blockIdx.x * ppc + warp * gpw + grp;
  const int64_t q0 = (int64_t)n * HW + p;
  const bf16* pu = uhat + q0 * C + c;
  const bf16* po = out + q0 * out_pitch + c;
  const bf16* pd = dout + q0 * dout_pitch + c;
  bf16* py = dy + q0 * dy_pitch + c;
  const float4* px4 = reinterpret_cast<const float4*>(bwd_px);
  const int64_t su = (int64_t)step * C, so = (int64_t)step * out_pitch, sd = (int64_t)step * dout_pitch;
  const int64_t sy = (int64_t)step * dy_pitch;
  const float invC = 1.f / (float)C;
  int64_t q = q0;
  for (; p < p_end; p += 2 * step, q += 2 * step) {
    const bool v1 = p + step < p_end;
    bf16x8 ru[2], ro[2], rd[2];
    NbfPix px[2];
    ru[0] = ldg8(pu); ro[0] = ldg8(po); rd[0] = ldg8(pd);
    px[0] = nbf_pix<MODE>(gs, px4, cidx, q, invC, c);
    if (v1) {
      ru[1] = ldg8(pu + su); ro[1] = ldg8(po + so); rd[1] = ldg8(pd + sd);
      px[1] = nbf_pix<MODE>(gs, px4, cidx, q + step, invC, c);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      if (u && !v1) break;
      float uh[8], ds[8], r[8];
      unpack8(ru[u], uh);
      nbf_ds8(rd[u], ro[u], slope, ds);
      if (MODE == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = fmaf(-uh[i], h_am2[i], fmaf(ds[i], h_a[i], -h_am1[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          // a*du = ds*(a*[direct] + a*gc*g) + a*gc*dsp
          const float dsp = (i == px[u].k) ? px[u].dmx : px[u].dmean;
          const float k1 = MODE == 1 ? fmaf(h_ag[i], px[u].g, h_a[i]) : h_ag[i] * px[u].g;
          r[i] = fmaf(-uh[i], h_am2[i], fmaf(ds[i], k1, fmaf(h_ag[i], dsp, -h_am1[i])));
        }
        const int pp = p + u * step;
        if (pp >= imin && pp <= imax) {            // rare: this pixel is the arg-max of one of the lane's channels
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (pp == s_idx[c + i]) r[i] += s_admx[c + i];
        }
      }
      stg8(py + (int64_t)u * sy, pack8(r));
    }
    pu += 2 * su; po += 2 * so; pd += 2 * sd; py += 2 * sy;
  }
}

// ---------------------------------------------------------------------------------------------------
// Small maps (H*W <= 128: the 6x3 / 3x2 / 12x2 / 6x4 / 12x7 / 24x4 / 12x8 sites with 256..1024 channels).
// One CTA owns one sample, so every per-(n,c) and per-pixel reduction is CTA-local: the whole forward (stats,
// coefficients, channel MLP, spatial pooling, gate, apply) is ONE kernel and the whole backward is ONE kernel,
// phases separated by __syncthreads instead of launches; the sample's tensors (<= 200 KB) are re-read from L1/L2.
// Thread (cv, pl): channel vector cv = t % NV (8 channels), pixel lane pl = t / NV; NV = C/8, PL = 256/NV.
// ---------------------------------------------------------------------------------------------------
constexpr int SMALL_HW = 128;

// combine a per-thread partial over 8 channels of pixel p into the per-pixel accumulator
__device__ __forceinline__ void pix_add(float v, int G, float* dst) {
  for (int o = G >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if ((threadIdx.x & (G - 1)) == 0) atomicAdd(dst, v);
}

template <bool F32>
__global__ void __launch_bounds__(256) nb_small_fwd_kernel(const bvae_nb_desc d) {
  extern __shared__ float sm[];
  const int C = d.C, H = d.H, W = d.W, HW = H * W, NV = C / 8, PL = 256 / NV;
  float* s_mean = sm; float* s_rstd = sm + C; float* s_a = sm + 2 * C; float* s_b = sm + 3 * C;
  float* s_gc = sm + 4 * C; float* s_mx = sm + 5 * C; float* s_sum = sm + 6 * C; float* s_sq = sm + 7 * C;
  u64* s_kmax = (u64*)(sm + 8 * C);
  u64* s_kmin = s_kmax + C;
  float* s_psum = (float*)(s_kmin + C);         // [128]
  u64* s_pkey = (u64*)(s_psum + SMALL_HW);      // [128]
  float* s_sa = (float*)(s_pkey + SMALL_HW);    // [128][2]
  float* s_gs = s_sa + 2 * SMALL_HW;            // [128]
  float* s_h = s_gs + SMALL_HW;                 // [64]
  float* s_w = s_h + 64;                        // [18] attention conv weights, [32..96) MLP scratch
  const int n = blockIdx.x, t = threadIdx.x;
  const int cv = t % NV, pl = t / NV, c0 = cv * 8;
  const int G = NV < 32 ? NV : 32;              // lanes of a warp that share a pixel
  const int64_t ybase = (int64_t)n * HW * d.y_pitch, ubase = (int64_t)n * HW * C;
  const int64_t obase = (int64_t)n * HW * d.out_pitch, rbase = (int64_t)n * HW * d.res_pitch;
  bf16* uhat = (bf16*)d.uhat;
  bf16* out = (bf16*)d.out;
  const bf16* res = (const bf16*)d.res;

  for (int c = t; c < C; c += 256) { s_sum[c] = 0.f; s_sq[c] = 0.f; s_kmax[c] = 0; s_kmin[c] = 0; }
  if (t < SMALL_HW) { s_psum[t] = 0.f; s_pkey[t] = 0; }
  if (t < 18 && d.has_cbam) s_w[t] = d.wsp[t];
  __syncthreads();

  // ---- phase A: per-channel shifted sums and extrema
  {
    float shift[8], sum[8], sq[8], vmx[8], vmn[8];
    int imx[8], imn[8];
    load8<F32>(d.y, ybase + c0, shift);
#pragma unroll
    for (int i = 0; i < 8; ++i) { sum[i] = 0.f; sq[i] = 0.f; vmx[i] = -INFINITY; vmn[i] = INFINITY; imx[i] = 0; imn[i] = 0; }
    for (int p = pl; p < HW; p += PL) {
      float v[8];
      load8<F32>(d.y, ybase + (int64_t)p * d.y_pitch + c0, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dl = v[i] - shift[i];
        sum[i] += dl; sq[i] += dl * dl;
        if (v[i] > vmx[i]) { vmx[i] = v[i]; imx[i] = p; }
        if (v[i] < vmn[i]) { vmn[i] = v[i]; imn[i] = p; }
      }
    }
    ordered([&] {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        atomicAdd(&s_sum[c0 + i], sum[i]);
        atomicAdd(&s_sq[c0 + i], sq[i]);
        if (vmx[i] != -INFINITY) atomicMax(&s_kmax[c0 + i], make_key(vmx[i], (uint32_t)imx[i]));
        if (vmn[i] != INFINITY) atomicMax(&s_kmin[c0 + i], make_key(-vmn[i], (uint32_t)imn[i]));
      }
    });
  }
  __syncthreads();
  const float inv = 1.f / (float)HW;
  for (int c = t; c < C; c += 256) {
    const int64_t o = (int64_t)n * C + c;
    const float shift = F32 ? ((const float*)d.y)[ybase + c] : bf2f(((const bf16*)d.y)[ybase + c]);
    const float md = s_sum[c] * inv;
    const float mean = shift + md;
    const float var = fmaxf(s_sq[c] * inv - md * md, 0.f);
    const float rstd = rsqrtf(var + d.eps);
    const float g = d.gamma[c], b0 = d.beta[c];
    const float a = g * rstd, b = b0 - mean * a;
    float yext; uint32_t idx;
    if (a >= 0.f) { const u64 k = s_kmax[c]; yext = key_val(k); idx = key_idx(k); }
    else { const u64 k = s_kmin[c]; yext = -key_val(k); idx = key_idx(k); }
    const float ext_uhat = (yext - mean) * rstd;
    const float ext_u = g * ext_uhat + b0;
    float* q = d.nc + o * NC_W;
    q[NC_MEAN] = mean; q[NC_RSTD] = rstd; q[NC_A] = a; q[NC_B] = b;
    q[NC_EXTU] = ext_u; q[NC_EXTUHAT] = ext_uhat; q[NC_GC] = 1.f; q[NC_SPARE] = 0.f;
    d.nc_idx[o] = (int32_t)idx;
    s_mean[c] = mean; s_rstd[c] = rstd; s_a[c] = a; s_b[c] = b; s_mx[c] = ext_u; s_gc[c] = 1.f;
  }
  __syncthreads();

  // ---- phase B: channel-attention MLP (avg pool of an instance-normalised map is exactly beta)
  if (d.has_cbam) {
    for (int c = t; c < C; c += 256) s_sq[c] = d.beta[c];       // s_sq is free after phase A: the avg-pooled vector
    __syncthreads();
    mlp_hidden(d.w1, C, d.Cr, s_sq, s_mx, s_h, s_w + 32);
    __syncthreads();
    if (t < d.Cr) s_h[t] = fmaxf(s_h[t], 0.f) + fmaxf(s_w[32 + t], 0.f);
    __syncthreads();
    mlp_rows_dot(d.w2, C, d.Cr, s_h, s_sum);       // s_sum is free after phase A
    __syncthreads();
    for (int c = t; c < C; c += 256) {
      const float gc = 1.f / (1.f + expf(-s_sum[c]));
      s_gc[c] = gc;
      d.nc[((int64_t)n * C + c) * NC_W + NC_GC] = gc;
    }
    __syncthreads();
  }

  // ---- phase C: write uhat; per-pixel mean / max / argmax over channels of u*gc (or the final output without CBAM)
  for (int pb = 0; pb < HW; pb += PL) {
    const int p = pb + pl;
    const bool valid = p < HW;
    float sum = 0.f, mx = -INFINITY;
    int mxc = 0;
    if (valid) {
      float v[8], uh[8], o[8];
      load8<F32>(d.y, ybase + (int64_t)p * d.y_pitch + c0, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uh[i] = (v[i] - s_mean[c0 + i]) * s_rstd[c0 + i];
        const float u = s_a[c0 + i] * v[i] + s_b[c0 + i];
        const float u1 = u * s_gc[c0 + i];
        sum += u1;
        if (u1 > mx) { mx = u1; mxc = c0 + i; }
        o[i] = act_fwd(u, d.slope);
      }
      if (uhat != nullptr) stg8(uhat + ubase + (int64_t)p * C + c0, pack8(uh));
      if (!d.has_cbam) stg8(out + obase + (int64_t)p * d.out_pitch + c0, pack8(o));
    }
    if (d.has_cbam) {
      for (int o = G >> 1; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
        const int omc = __shfl_xor_sync(0xffffffffu, mxc, o);
        if (omx > mx || (omx == mx && omc < mxc)) { mx = omx; mxc = omc; }
      }
      ordered([&] {
        if (valid && (t & (G - 1)) == 0) {
          atomicAdd(&s_psum[p], sum);
          atomicMax(&s_pkey[p], make_key(mx, (uint32_t)mxc));     // max value, smallest channel index on ties
        }
      });
    }
  }
  if (!d.has_cbam) return;
  __syncthreads();
  if (t < HW) {
    const u64 k = s_pkey[t];
    const float mean_c = s_psum[t] / (float)C, max_c = key_val(k);
    s_sa[2 * t] = mean_c; s_sa[2 * t + 1] = max_c;
    const int64_t o = (int64_t)n * HW + t;
    d.sa[o * 2] = mean_c; d.sa[o * 2 + 1] = max_c;
    d.cidx[o] = (int32_t)key_idx(k);
  }
  __syncthreads();
  // ---- phase D: spatial gate
  if (t < HW) {
    const float g = 1.f / (1.f + expf(-sa_conv(s_sa, s_w, H, W, t / W, t % W)));
    s_gs[t] = g;
    d.gs[(int64_t)n * HW + t] = g;
  }
  __syncthreads();
  // ---- phase E: out = act(r + u*gc*gs)
  for (int p = pl; p < HW; p += PL) {
    float v[8], r[8], o[8];
    load8<F32>(d.y, ybase + (int64_t)p * d.y_pitch + c0, v);
    if (d.res_mode == 2) unpack8(ldg8(res + rbase + (int64_t)p * d.res_pitch + c0), r);
    const float g = s_gs[p];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float u = s_a[c0 + i] * v[i] + s_b[c0 + i];
      const float cb = u * s_gc[c0 + i] * g;
      const float rr = d.res_mode == 1 ? u : (d.res_mode == 2 ? r[i] : 0.f);
      o[i] = act_fwd(rr + cb, d.slope);
    }
    stg8(out + obase + (int64_t)p * d.out_pitch + c0, pack8(o));
  }
}

__global__ void __launch_bounds__(256) nb_small_bwd_kernel(const bvae_nb_desc d) {
  extern __shared__ float sm[];
  const int C = d.C, H = d.H, W = d.W, HW = H * W, NV = C / 8, PL = 256 / NV, Cr = d.Cr;
  float* s_g = sm; float* s_b = sm + C; float* s_gc = sm + 2 * C; float* s_a = sm + 3 * C;
  float* s_dgc = sm + 4 * C; float* s_S1 = sm + 5 * C; float* s_S2 = sm + 6 * C; float* s_dv = sm + 7 * C;
  float* s_m1 = sm + 8 * C; float* s_m2 = sm + 9 * C; float* s_dmx = sm + 10 * C;
  int* s_idx = (int*)(sm + 11 * C);
  float* s_gs = sm + 12 * C;                    // [128]
  float* s_dq = s_gs + SMALL_HW;                // [128]  (first dgs, then dq)
  float* s_dmean = s_dq + SMALL_HW;             // [128]
  float* s_dmax = s_dmean + SMALL_HW;           // [128]
  int* s_cidx = (int*)(s_dmax + SMALL_HW);      // [128]
  float* s_sa = (float*)(s_cidx + SMALL_HW);    // [128][2]
  float* s_ha = s_sa + 2 * SMALL_HW;            // [64]
  float* s_hm = s_ha + 64; float* s_dh = s_hm + 64;
  float* s_w = s_dh + 64;                       // [18]
  float* s_dw = s_w + 18;                       // [18]
  const int n = blockIdx.x, t = threadIdx.x;
  const int cv = t % NV, pl = t / NV, c0 = cv * 8;
  const int G = NV < 32 ? NV : 32;
  const int lane = t & 31;
  const int has_cbam = d.has_cbam, res_mode = d.res_mode;
  const float slope = d.slope;
  const bf16* dout = (const bf16*)d.dout; const bf16* out = (const bf16*)d.out; const bf16* uhat = (const bf16*)d.uhat;
  bf16* dy = (bf16*)d.dy; bf16* dres = (bf16*)d.dres;
  const int64_t ubase = (int64_t)n * HW * C, obase = (int64_t)n * HW * d.out_pitch, dbase = (int64_t)n * HW * d.dout_pitch;
  const int64_t ybase = (int64_t)n * HW * d.dy_pitch, rbase = (int64_t)n * HW * d.dres_pitch;
  const float* q0 = d.nc + (int64_t)n * C * NC_W;

  for (int c = t; c < C; c += 256) {
    s_g[c] = d.gamma[c]; s_b[c] = d.beta[c];
    s_gc[c] = q0[c * NC_W + NC_GC]; s_a[c] = q0[c * NC_W + NC_A];
    s_dgc[c] = 0.f; s_S1[c] = 0.f; s_S2[c] = 0.f;
    s_idx[c] = has_cbam ? d.nc_idx[(int64_t)n * C + c] : -1;
  }
  if (t < SMALL_HW) {
    const bool v = has_cbam && t < HW;
    const int64_t o = (int64_t)n * HW + t;
    s_gs[t] = v ? d.gs[o] : 0.f;
    s_cidx[t] = v ? d.cidx[o] : -1;
    s_sa[2 * t] = v ? d.sa[o * 2] : 0.f; s_sa[2 * t + 1] = v ? d.sa[o * 2 + 1] : 0.f;
    s_dq[t] = 0.f; s_dmean[t] = 0.f; s_dmax[t] = 0.f;
  }
  if (t < 18) { s_w[t] = has_cbam ? d.wsp[t] : 0.f; s_dw[t] = 0.f; }
  __syncthreads();

  if (has_cbam) {
    // ---- phase 1: dgs[p] = sum_c ds*u*gc ; dgc[c] += sum_p ds*u*gs
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int pb = 0; pb < HW; pb += PL) {
      const int p = pb + pl;
      float dgs = 0.f;
      if (p < HW) {
        float uh[8], o[8], dd[8];
        unpack8(ldg8(uhat + ubase + (int64_t)p * C + c0), uh);
        unpack8(ldg8(out + obase + (int64_t)p * d.out_pitch + c0), o);
        unpack8(ldg8(dout + dbase + (int64_t)p * d.dout_pitch + c0), dd);
        const float g = s_gs[p];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float ds = dd[i] * (o[i] > 0.f ? 1.f : slope);
          const float tt = ds * (s_g[c0 + i] * uh[i] + s_b[c0 + i]);
          dgs += tt * s_gc[c0 + i];
          acc[i] += tt * g;
        }
      }
      for (int o = G >> 1; o > 0; o >>= 1) dgs += __shfl_xor_sync(0xffffffffu, dgs, o);
      if (p < HW && (t & (G - 1)) == 0) atomicAdd(&s_dq[p], dgs);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(&s_dgc[c0 + i], acc[i]);
    __syncthreads();
    // ---- phase 2: dq, then the 3x3 transpose conv and the attention-conv weight gradient
    if (t < HW) { const float g = s_gs[t]; s_dq[t] = s_dq[t] * g * (1.f - g); }
    __syncthreads();
    float part[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) part[i] = 0.f;
    if (t < HW) {
      const int py = t / W, px = t % W;
      const float dq = s_dq[t];
      float dmean = 0.f, dmax = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int yy = py - (ky - 1), xx = px - (kx - 1);
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const float dqq = s_dq[yy * W + xx];
            dmean += s_w[ky * 3 + kx] * dqq;
            dmax += s_w[9 + ky * 3 + kx] * dqq;
          }
          const int y2 = py + ky - 1, x2 = px + kx - 1;
          if (y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) {
            part[ky * 3 + kx] = dq * s_sa[2 * (y2 * W + x2)];
            part[9 + ky * 3 + kx] = dq * s_sa[2 * (y2 * W + x2) + 1];
          }
        }
      s_dmean[t] = dmean / (float)C;
      s_dmax[t] = dmax;
    }
    if (t < SMALL_HW) {          // warps 0..3 hold all pixels
#pragma unroll
      for (int i = 0; i < 18; ++i) {
        const float v = warp_sum(part[i]);
        if (lane == 0) atomicAdd(&s_dw[i], v);
      }
    }
    __syncthreads();
    if (t < 18) atomicAdd(d.dwsp + t, s_dw[t]);
  }

  // ---- phase 3: S1 = sum du, S2 = sum du*uhat, dgc += sum dsp*u ; dres
  {
    float a1[8], a2[8], a3[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a1[i] = 0.f; a2[i] = 0.f; a3[i] = 0.f; }
    for (int p = pl; p < HW; p += PL) {
      float uh[8], o[8], dd[8], du[8], dsv[8], dspu[8];
      unpack8(ldg8(uhat + ubase + (int64_t)p * C + c0), uh);
      unpack8(ldg8(out + obase + (int64_t)p * d.out_pitch + c0), o);
      unpack8(ldg8(dout + dbase + (int64_t)p * d.dout_pitch + c0), dd);
      nb_du8<true>(uh, o, dd, c0, s_g + c0, s_b + c0, s_gc + c0, has_cbam, res_mode, slope, s_gs[p], s_dmean[p], s_dmax[p],
                   s_cidx[p], du, dsv, dspu);
#pragma unroll
      for (int i = 0; i < 8; ++i) { a1[i] += du[i]; a2[i] += du[i] * uh[i]; a3[i] += dspu[i]; }
      if (res_mode == 2 && dres) stg8(dres + rbase + (int64_t)p * d.dres_pitch + c0, pack8(dsv));
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      atomicAdd(&s_S1[c0 + i], a1[i]);
      atomicAdd(&s_S2[c0 + i], a2[i]);
      if (has_cbam) atomicAdd(&s_dgc[c0 + i], a3[i]);
    }
  }
  __syncthreads();

  // ---- phase 4: channel-MLP backward, dgamma / dbeta, InstanceNorm-backward means
  if (has_cbam) {
    for (int c = t; c < C; c += 256) { const float gc = s_gc[c]; s_dv[c] = s_dgc[c] * gc * (1.f - gc); }
    if (t < 64) s_dh[t] = 0.f;
    __syncthreads();
    for (int c = t; c < C; c += 256) s_m1[c] = q0[c * NC_W + NC_EXTU];      // s_m1 is written only after this phase
    __syncthreads();
    mlp_hidden(d.w1, C, Cr, s_b, s_m1, s_ha, s_hm);
    mlp_cols_dot(d.w2, C, Cr, s_dv, s_dh);
    __syncthreads();
    for (int j = t; j < Cr; j += 256) {
      float* hq = d.bwd_h + (int64_t)n * 192;
      hq[j] = fmaxf(s_ha[j], 0.f) + fmaxf(s_hm[j], 0.f);
      hq[64 + j] = s_ha[j] > 0.f ? s_dh[j] : 0.f;
      hq[128 + j] = s_hm[j] > 0.f ? s_dh[j] : 0.f;
    }
  }
  const float inv = 1.f / (float)HW;
  for (int c = t; c < C; c += 256) {
    float d_avg = 0.f, d_mx = 0.f;
    if (has_cbam) {
      for (int j = 0; j < Cr; ++j) {
        const float w = d.w1[(int64_t)j * C + c];
        if (s_ha[j] > 0.f) d_avg += w * s_dh[j];
        if (s_hm[j] > 0.f) d_mx += w * s_dh[j];
      }
      d.bwd_nc[((int64_t)n * C + c) * BN_W + BN_DGC] = s_dv[c];      // consumed by nb_bwd_w_kernel
    }
    const float ext_uhat = q0[c * NC_W + NC_EXTUHAT];
    const float S1 = s_S1[c] + d_mx;
    const float S2 = s_S2[c] + d_mx * ext_uhat;
    atomicAdd(d.dbeta + c, S1 + d_avg);
    atomicAdd(d.dgamma + c, S2);
    const float a = s_a[c];
    s_dmx[c] = a * d_mx; s_m1[c] = a * S1 * inv; s_m2[c] = a * S2 * inv;
  }
  __syncthreads();

  // ---- phase 5: dy = a*du + [p == argmax] a*d_mx - a*m1 - uhat*(a*m2)
  for (int p = pl; p < HW; p += PL) {
    float uh[8], o[8], dd[8], du[8], dsv[8], dspu[8], r[8];
    unpack8(ldg8(uhat + ubase + (int64_t)p * C + c0), uh);
    unpack8(ldg8(out + obase + (int64_t)p * d.out_pitch + c0), o);
    unpack8(ldg8(dout + dbase + (int64_t)p * d.dout_pitch + c0), dd);
    nb_du8<false>(uh, o, dd, c0, s_gc + c0, s_gc + c0, s_gc + c0, has_cbam, res_mode, slope, s_gs[p], s_dmean[p], s_dmax[p],
                  s_cidx[p], du, dsv, dspu);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float extra = (p == s_idx[c0 + i]) ? s_dmx[c0 + i] : 0.f;
      r[i] = s_a[c0 + i] * du[i] + extra - s_m1[c0 + i] - uh[i] * s_m2[c0 + i];
    }
    stg8(dy + ybase + (int64_t)p * d.dy_pitch + c0, pack8(r));
  }
}

static bool nb_small_ok(const bvae_nb_desc* d) { return d->H * d->W <= SMALL_HW && d->C >= 32 && d->C <= 1024; }
static size_t nb_small_fwd_smem(int C) { return (size_t)(8 * C) * 4 + (size_t)2 * C * 8 + SMALL_HW * (4 + 8 + 8 + 4) + (64 + 96 + 16) * 4; }
static size_t nb_small_bwd_smem(int C) { return (size_t)(12 * C) * 4 + SMALL_HW * 7 * 4 + (3 * 64 + 18 + 18 + 12) * 4; }

// ---------------------------------------------------------------------------------------------------
// Cluster kernels: ONE thread-block cluster owns one sample (CL = 1 CTA for maps <= 128 pixels, 4 or 8 CTAs for the
// large maps; CTA r owns the pixel slice [r*slice, (r+1)*slice)).  Every per-(n,c) reduction is a CTA-local shared
// memory reduction followed by an all-reduce over the cluster's distributed shared memory, every phase boundary is a
// cluster barrier instead of a kernel launch, and because the whole sample is swept by co-scheduled CTAs back to back
// the 2nd and 3rd sweeps of the forward (and of the backward) hit L2 instead of HBM.
// ---------------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;
constexpr int CLT = 128;   // threads per CTA (256 with 3 CTAs per SM measured no faster in round 2: 22.9 vs 22.7 ms of norm blocks per step)
constexpr int CL_MINB = 6; // CTAs per SM the launch bounds ask for

__device__ __forceinline__ float cl_sum(cg::cluster_group& cl, float* s_arr, int c, int CLn) {
  float v = 0.f;
  for (int r = 0; r < CLn; ++r) v += cl.map_shared_rank(s_arr, r)[c];
  return v;
}
__device__ __forceinline__ u64 cl_max(cg::cluster_group& cl, u64* s_arr, int c, int CLn) {
  u64 v = 0;
  for (int r = 0; r < CLn; ++r) { const u64 o = cl.map_shared_rank(s_arr, r)[c]; v = o > v ? o : v; }
  return v;
}

// stage 0 = the whole forward; 1 = statistics + coefficients only (the channel MLP then runs BATCHED over samples in
// nb_mlp_fwd_kernel); 2 = everything after the MLP, coefficients re-read from d.nc
template <bool F32>
__global__ void __launch_bounds__(CLT, CL_MINB) nb_cl_fwd_kernel(const bvae_nb_desc d, int CLn, int slice, int stage) {
  cg::cluster_group cl = cg::this_cluster();
  extern __shared__ float sm[];
  const int C = d.C, H = d.H, W = d.W, HW = H * W, NV = C / 8, PL = CLT / NV;
  float* s_mean = sm; float* s_rstd = sm + C; float* s_a = sm + 2 * C; float* s_b = sm + 3 * C;
  float* s_gc = sm + 4 * C; float* s_mx = sm + 5 * C; float* s_sum = sm + 6 * C; float* s_sq = sm + 7 * C;
  u64* s_kmax = (u64*)(sm + 8 * C);
  u64* s_kmin = s_kmax + C;
  float* s_h = (float*)(s_kmin + C);            // [64]
  float* s_t = s_h + 64;                        // [64]
  float* s_w = s_t + 64;                        // [32]
  u64* s_pkey = (u64*)(s_w + 32);               // [slice]
  float* s_psum = (float*)(s_pkey + slice);     // [slice]
  float* s_gs = s_psum + slice;                 // [slice]
  const int rank = CLn > 1 ? (int)cl.block_rank() : 0;
  const int n = blockIdx.x / CLn, t = threadIdx.x;
  const int cv = t % NV, pl = t / NV, c0 = cv * 8;
  const int G = NV < 32 ? NV : 32;
  const int p_lo = rank * slice, p_hi = min(HW, p_lo + slice), np = max(0, p_hi - p_lo);
  const int64_t ybase = (int64_t)n * HW * d.y_pitch, ubase = (int64_t)n * HW * C;
  const int64_t obase = (int64_t)n * HW * d.out_pitch, rbase = (int64_t)n * HW * d.res_pitch;
  bf16* uhat = (bf16*)d.uhat;
  bf16* out = (bf16*)d.out;
  const bf16* res = (const bf16*)d.res;

  for (int c = t; c < C; c += CLT) { s_sum[c] = 0.f; s_sq[c] = 0.f; s_kmax[c] = 0; s_kmin[c] = 0; }
  for (int i = t; i < slice; i += CLT) { s_psum[i] = 0.f; s_pkey[i] = 0; }
  if (t < 18 && d.has_cbam) s_w[t] = d.wsp[t];
  if (stage == 2) {
    for (int c = t; c < C; c += CLT) {
      const float4* q = reinterpret_cast<const float4*>(d.nc + ((int64_t)n * C + c) * NC_W);
      const float4 lo = q[0], hi = q[1];          // {mean, rstd, a, b} {gc, ext_u, ext_uhat, -}
      s_mean[c] = lo.x; s_rstd[c] = lo.y; s_a[c] = lo.z; s_b[c] = lo.w; s_gc[c] = hi.x; s_mx[c] = hi.y;
    }
  }
  __syncthreads();

  // ---- phase A: per-channel shifted sums and extrema over this CTA's pixel slice
  if (stage != 2) {
    float shift[8], sum[8], sq[8], vmx[8], vmn[8];
    int imx[8], imn[8];
    load8<F32>(d.y, ybase + c0, shift);
#pragma unroll
    for (int i = 0; i < 8; ++i) { sum[i] = 0.f; sq[i] = 0.f; vmx[i] = -INFINITY; vmn[i] = INFINITY; imx[i] = 0; imn[i] = 0; }
    for (int pb = p_lo + pl; pb < p_hi; pb += 4 * PL) {
      float vv[4][8];
#pragma unroll
      for (int u = 0; u < 4; ++u)                       // issue all loads of the batch before touching any of them
        if (pb + u * PL < p_hi) load8<F32>(d.y, ybase + (int64_t)(pb + u * PL) * d.y_pitch + c0, vv[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int p = pb + u * PL;
        if (p >= p_hi) break;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float dl = vv[u][i] - shift[i];
          sum[i] += dl; sq[i] += dl * dl;
          if (vv[u][i] > vmx[i]) { vmx[i] = vv[u][i]; imx[i] = p; }
          if (vv[u][i] < vmn[i]) { vmn[i] = vv[u][i]; imn[i] = p; }
        }
      }
    }
    // lanes of a warp with equal cv hold different pixel lanes: combine them first when NV < 32
    for (int o = G; o < 32; o <<= 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], o);
        sq[i] += __shfl_xor_sync(0xffffffffu, sq[i], o);
        const float ox = __shfl_xor_sync(0xffffffffu, vmx[i], o), on = __shfl_xor_sync(0xffffffffu, vmn[i], o);
        const int oix = __shfl_xor_sync(0xffffffffu, imx[i], o), oin = __shfl_xor_sync(0xffffffffu, imn[i], o);
        if (ox > vmx[i] || (ox == vmx[i] && oix < imx[i])) { vmx[i] = ox; imx[i] = oix; }
        if (on < vmn[i] || (on == vmn[i] && oin < imn[i])) { vmn[i] = on; imn[i] = oin; }
      }
    }
    ordered([&] {
      if ((t & 31) < G) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          atomicAdd(&s_sum[c0 + i], sum[i]);
          atomicAdd(&s_sq[c0 + i], sq[i]);
          if (vmx[i] != -INFINITY) atomicMax(&s_kmax[c0 + i], make_key(vmx[i], (uint32_t)imx[i]));
          if (vmn[i] != INFINITY) atomicMax(&s_kmin[c0 + i], make_key(-vmn[i], (uint32_t)imn[i]));
        }
      }
    });
  }
  if (CLn > 1) cl.sync(); else __syncthreads();
  const float inv = 1.f / (float)HW;
  for (int c = t; c < C && stage != 2; c += CLT) {
    const int64_t o = (int64_t)n * C + c;
    const float shift = F32 ? ((const float*)d.y)[ybase + c] : bf2f(((const bf16*)d.y)[ybase + c]);
    const float tsum = CLn > 1 ? cl_sum(cl, s_sum, c, CLn) : s_sum[c];
    const float tsq = CLn > 1 ? cl_sum(cl, s_sq, c, CLn) : s_sq[c];
    const u64 kx = CLn > 1 ? cl_max(cl, s_kmax, c, CLn) : s_kmax[c];
    const u64 kn = CLn > 1 ? cl_max(cl, s_kmin, c, CLn) : s_kmin[c];
    const float md = tsum * inv;
    const float mean = shift + md;
    const float var = fmaxf(tsq * inv - md * md, 0.f);
    const float rstd = rsqrtf(var + d.eps);
    const float g = d.gamma[c], b0 = d.beta[c];
    const float a = g * rstd, b = b0 - mean * a;
    float yext; uint32_t idx;
    if (a >= 0.f) { yext = key_val(kx); idx = key_idx(kx); }
    else { yext = -key_val(kn); idx = key_idx(kn); }
    const float ext_uhat = (yext - mean) * rstd;
    const float ext_u = g * ext_uhat + b0;
    if (rank == 0) {
      float* q = d.nc + o * NC_W;
      q[NC_MEAN] = mean; q[NC_RSTD] = rstd; q[NC_A] = a; q[NC_B] = b;
      q[NC_EXTU] = ext_u; q[NC_EXTUHAT] = ext_uhat; q[NC_GC] = 1.f; q[NC_SPARE] = 0.f;
      d.nc_idx[o] = (int32_t)idx;
    }
    s_mean[c] = mean; s_rstd[c] = rstd; s_a[c] = a; s_b[c] = b; s_mx[c] = ext_u; s_gc[c] = 1.f;
  }
  // all remote reads of the partial arrays are done before any CTA of the cluster may overwrite them or exit
  if (CLn > 1) cl.sync(); else __syncthreads();

  if (stage == 1) return;
  // ---- phase B: channel-attention MLP (every CTA of the cluster computes it redundantly: C <= 256 there)
  if (d.has_cbam && stage == 0) {
    for (int c = t; c < C; c += CLT) s_sq[c] = d.beta[c];        // avg pool of an instance-normalised map == beta
    __syncthreads();
    mlp_hidden(d.w1, C, d.Cr, s_sq, s_mx, s_h, s_t);
    __syncthreads();
    if (t < d.Cr) s_h[t] = fmaxf(s_h[t], 0.f) + fmaxf(s_t[t], 0.f);
    __syncthreads();
    mlp_rows_dot(d.w2, C, d.Cr, s_h, s_sum);
    __syncthreads();
    for (int c = t; c < C; c += CLT) {
      const float gc = 1.f / (1.f + expf(-s_sum[c]));
      s_gc[c] = gc;
      if (rank == 0) d.nc[((int64_t)n * C + c) * NC_W + NC_GC] = gc;
    }
    __syncthreads();
  }

  // ---- phase C: write uhat; per-pixel mean / max / argmax over channels of u*gc (or the final output without CBAM)
#pragma unroll 2
  for (int pb = 0; pb < np; pb += PL) {
    const int lp = pb + pl, p = p_lo + lp;
    const bool valid = lp < np;
    float sum = 0.f, mx = -INFINITY;
    int mxc = 0;
    if (valid) {
      float v[8], uh[8], o[8];
      load8<F32>(d.y, ybase + (int64_t)p * d.y_pitch + c0, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uh[i] = (v[i] - s_mean[c0 + i]) * s_rstd[c0 + i];
        const float u = s_a[c0 + i] * v[i] + s_b[c0 + i];
        const float u1 = u * s_gc[c0 + i];
        sum += u1;
        if (u1 > mx) { mx = u1; mxc = c0 + i; }
        o[i] = act_fwd(u, d.slope);
      }
      if (uhat != nullptr) stg8(uhat + ubase + (int64_t)p * C + c0, pack8(uh));
      if (!d.has_cbam) stg8(out + obase + (int64_t)p * d.out_pitch + c0, pack8(o));
    }
    if (d.has_cbam) {
      for (int o = G >> 1; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
        const int omc = __shfl_xor_sync(0xffffffffu, mxc, o);
        if (omx > mx || (omx == mx && omc < mxc)) { mx = omx; mxc = omc; }
      }
      ordered([&] {
        if (valid && (t & (G - 1)) == 0) {
          atomicAdd(&s_psum[lp], sum);
          atomicMax(&s_pkey[lp], make_key(mx, (uint32_t)mxc));
        }
      });
    }
  }
  if (!d.has_cbam) return;
  __syncthreads();
  for (int lp = t; lp < np; lp += CLT) {
    const u64 k = s_pkey[lp];
    const int64_t o = (int64_t)n * HW + p_lo + lp;
    d.sa[o * 2] = s_psum[lp] / (float)C;
    d.sa[o * 2 + 1] = key_val(k);
    d.cidx[o] = (int32_t)key_idx(k);
  }
  if (CLn > 1) cl.sync(); else __syncthreads();      // the 3x3 gate conv reads the neighbouring slices' sa from global/L2
  // ---- phase D: spatial gate for this slice
  for (int lp = t; lp < np; lp += CLT) {
    const int p = p_lo + lp;
    const float g = 1.f / (1.f + expf(-sa_conv(d.sa + (int64_t)n * HW * 2, s_w, H, W, p / W, p % W)));
    s_gs[lp] = g;
    d.gs[(int64_t)n * HW + p] = g;
  }
  __syncthreads();
  // ---- phase E: out = act(r + u*gc*gs)
  for (int lb = pl; lb < np; lb += 4 * PL) {
    float vv[4][8];
    bf16x8 rr4[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int lp = lb + u * PL;
      if (lp < np) {
        load8<F32>(d.y, ybase + (int64_t)(p_lo + lp) * d.y_pitch + c0, vv[u]);
        if (d.res_mode == 2) rr4[u] = ldg8(res + rbase + (int64_t)(p_lo + lp) * d.res_pitch + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int lp = lb + u * PL;
      if (lp >= np) break;
      const int p = p_lo + lp;
      float r[8], o[8];
      if (d.res_mode == 2) unpack8(rr4[u], r);
      const float g = s_gs[lp];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float u_ = s_a[c0 + i] * vv[u][i] + s_b[c0 + i];
        const float cb = u_ * s_gc[c0 + i] * g;
        const float rr = d.res_mode == 1 ? u_ : (d.res_mode == 2 ? r[i] : 0.f);
        o[i] = act_fwd(rr + cb, d.slope);
      }
      stg8(out + obase + (int64_t)p * d.out_pitch + c0, pack8(o));
    }
  }
}

// stage 0 = the whole backward; 1 = the reduction sweeps (totals of dgc, S1, S2 and the per-pixel spatial gradients go to
// d.bwd_nc / d.bwd_px; the channel-MLP backward then runs BATCHED in nb_mlp_bwd_kernel); 2 = the final dy sweep
__global__ void __launch_bounds__(CLT, CL_MINB) nb_cl_bwd_kernel(const bvae_nb_desc d, int CLn, int slice, int stage) {
  cg::cluster_group cl = cg::this_cluster();
  extern __shared__ float sm[];
  const int C = d.C, H = d.H, W = d.W, HW = H * W, NV = C / 8, PL = CLT / NV, Cr = d.Cr;
  float* s_g = sm; float* s_b = sm + C; float* s_gc = sm + 2 * C; float* s_a = sm + 3 * C;
  float* s_dgc = sm + 4 * C; float* s_S1 = sm + 5 * C; float* s_S2 = sm + 6 * C; float* s_dv = sm + 7 * C;
  float* s_m1 = sm + 8 * C; float* s_m2 = sm + 9 * C; float* s_dmx = sm + 10 * C;
  int* s_idx = (int*)(sm + 11 * C);
  float* s_ha = sm + 12 * C;                    // [64]
  float* s_hm = s_ha + 64; float* s_dh = s_hm + 64;
  float* s_w = s_dh + 64;                       // [32]
  float* s_dw = s_w + 32;                       // [32]
  float* s_gs = s_dw + 32;                      // [slice]
  float* s_dq = s_gs + slice;                   // [slice]  (dgs, then unused)
  float* s_dmean = s_dq + slice;                // [slice]
  float* s_dmax = s_dmean + slice;              // [slice]
  int* s_cidx = (int*)(s_dmax + slice);         // [slice]
  const int rank = CLn > 1 ? (int)cl.block_rank() : 0;
  const int n = blockIdx.x / CLn, t = threadIdx.x;
  const int cv = t % NV, pl = t / NV, c0 = cv * 8;
  const int G = NV < 32 ? NV : 32;
  const int lane = t & 31;
  const int p_lo = rank * slice, p_hi = min(HW, p_lo + slice), np = max(0, p_hi - p_lo);
  const int has_cbam = d.has_cbam, res_mode = d.res_mode;
  const float slope = d.slope;
  const bf16* dout = (const bf16*)d.dout; const bf16* out = (const bf16*)d.out; const bf16* uhat = (const bf16*)d.uhat;
  bf16* dy = (bf16*)d.dy; bf16* dres = (bf16*)d.dres;
  const int64_t ubase = (int64_t)n * HW * C, obase = (int64_t)n * HW * d.out_pitch, dbase = (int64_t)n * HW * d.dout_pitch;
  const int64_t ybase = (int64_t)n * HW * d.dy_pitch, rbase = (int64_t)n * HW * d.dres_pitch;
  const float* q0 = d.nc + (int64_t)n * C * NC_W;
  float* px_n = d.bwd_px + (int64_t)n * HW * BP_W;

  for (int c = t; c < C; c += CLT) {
    s_g[c] = d.gamma[c]; s_b[c] = d.beta[c];
    s_gc[c] = q0[c * NC_W + NC_GC]; s_a[c] = q0[c * NC_W + NC_A];
    s_dgc[c] = 0.f; s_S1[c] = 0.f; s_S2[c] = 0.f;
    s_idx[c] = has_cbam ? d.nc_idx[(int64_t)n * C + c] : -1;
  }
  for (int lp = t; lp < slice; lp += CLT) {
    const bool v = has_cbam && lp < np;
    const int64_t o = (int64_t)n * HW + p_lo + lp;
    s_gs[lp] = v ? d.gs[o] : 0.f;
    s_cidx[lp] = v ? d.cidx[o] : -1;
    s_dq[lp] = 0.f; s_dmean[lp] = 0.f; s_dmax[lp] = 0.f;
  }
  if (t < 32) { s_w[t] = (has_cbam && t < 18) ? d.wsp[t] : 0.f; s_dw[t] = 0.f; }
  if (stage == 2) {
    for (int c = t; c < C; c += CLT) {
      const float4 v = *reinterpret_cast<const float4*>(d.bwd_nc + ((int64_t)n * C + c) * BN_W);   // {dv, a*m1, a*m2, a*d_mx}
      s_m1[c] = v.y; s_m2[c] = v.z; s_dmx[c] = v.w;
    }
    for (int lp = t; lp < np; lp += CLT) {
      const float4 v = *reinterpret_cast<const float4*>(px_n + (int64_t)(p_lo + lp) * BP_W);       // {dq, dmean / C, dmax, -}
      s_dmean[lp] = v.y; s_dmax[lp] = v.z;
    }
  }
  __syncthreads();

  if (has_cbam && stage != 2) {
    // ---- phase 1: dgs[p] = sum_c ds*u*gc (complete inside the CTA) ; dgc[c] += sum_p ds*u*gs (partial)
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll 2
    for (int pb = 0; pb < np; pb += PL) {
      const int lp = pb + pl, p = p_lo + lp;
      float dgs = 0.f;
      if (lp < np) {
        float uh[8], o[8], dd[8];
        unpack8(ldg8(uhat + ubase + (int64_t)p * C + c0), uh);
        unpack8(ldg8(out + obase + (int64_t)p * d.out_pitch + c0), o);
        unpack8(ldg8(dout + dbase + (int64_t)p * d.dout_pitch + c0), dd);
        const float g = s_gs[lp];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float ds = dd[i] * (o[i] > 0.f ? 1.f : slope);
          const float tt = ds * (s_g[c0 + i] * uh[i] + s_b[c0 + i]);
          dgs += tt * s_gc[c0 + i];
          acc[i] += tt * g;
        }
      }
      for (int o = G >> 1; o > 0; o >>= 1) dgs += __shfl_xor_sync(0xffffffffu, dgs, o);
      if (lp < np && (t & (G - 1)) == 0) atomicAdd(&s_dq[lp], dgs);
    }
    for (int o = G; o < 32; o <<= 1)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], o);
    if (lane < G)
#pragma unroll
      for (int i = 0; i < 8; ++i) atomicAdd(&s_dgc[c0 + i], acc[i]);
    __syncthreads();
    // dq for this slice -> global (the transpose conv below needs the neighbouring slices' values)
    for (int lp = t; lp < np; lp += CLT) {
      const float g = s_gs[lp];
      px_n[(int64_t)(p_lo + lp) * BP_W + BP_DQ] = s_dq[lp] * g * (1.f - g);
    }
    if (CLn > 1) cl.sync(); else __syncthreads();
    // ---- phase 2: 3x3 transpose conv of dq and the attention-conv weight gradient
    float part[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) part[i] = 0.f;
    const float* sa_n = d.sa + (int64_t)n * HW * 2;
    for (int lp = t; lp < np; lp += CLT) {
      const int p = p_lo + lp, py = p / W, px = p % W;
      const float dq = px_n[(int64_t)p * BP_W + BP_DQ];
      float dmean = 0.f, dmax = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int yy = py - (ky - 1), xx = px - (kx - 1);
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const float dqq = px_n[((int64_t)yy * W + xx) * BP_W + BP_DQ];
            dmean += s_w[ky * 3 + kx] * dqq;
            dmax += s_w[9 + ky * 3 + kx] * dqq;
          }
          const int y2 = py + ky - 1, x2 = px + kx - 1;
          if (y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) {
            const float2 v = *reinterpret_cast<const float2*>(sa_n + ((int64_t)y2 * W + x2) * 2);
            part[ky * 3 + kx] += dq * v.x;
            part[9 + ky * 3 + kx] += dq * v.y;
          }
        }
      s_dmean[lp] = dmean / (float)C;
      s_dmax[lp] = dmax;
      if (stage == 1) { px_n[(int64_t)p * BP_W + BP_DMEAN] = dmean / (float)C; px_n[(int64_t)p * BP_W + BP_DMAX] = dmax; }
    }
#pragma unroll
    for (int i = 0; i < 18; ++i) {
      const float v = warp_sum(part[i]);
      if (lane == 0 && v != 0.f) atomicAdd(&s_dw[i], v);
    }
    __syncthreads();
    if (t < 18) atomicAdd(d.dwsp + t, s_dw[t]);
  }

  // ---- phase 3: S1 = sum du, S2 = sum du*uhat, dgc += sum dsp*u (partials over this slice) ; dres
  if (stage != 2) {
    float a1[8], a2[8], a3[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a1[i] = 0.f; a2[i] = 0.f; a3[i] = 0.f; }
    for (int lb = pl; lb < np; lb += 2 * PL) {
      bf16x8 ru[2], ro[2], rd[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int lp = lb + u * PL;
        if (lp < np) {
          const int p = p_lo + lp;
          ru[u] = ldg8(uhat + ubase + (int64_t)p * C + c0);
          ro[u] = ldg8(out + obase + (int64_t)p * d.out_pitch + c0);
          rd[u] = ldg8(dout + dbase + (int64_t)p * d.dout_pitch + c0);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int lp = lb + u * PL;
        if (lp >= np) break;
        const int p = p_lo + lp;
        float uh[8], o[8], dd[8], du[8], dsv[8], dspu[8];
        unpack8(ru[u], uh); unpack8(ro[u], o); unpack8(rd[u], dd);
        nb_du8<true>(uh, o, dd, c0, s_g + c0, s_b + c0, s_gc + c0, has_cbam, res_mode, slope, s_gs[lp], s_dmean[lp],
                     s_dmax[lp], s_cidx[lp], du, dsv, dspu);
#pragma unroll
        for (int i = 0; i < 8; ++i) { a1[i] += du[i]; a2[i] += du[i] * uh[i]; a3[i] += dspu[i]; }
        if (res_mode == 2 && dres) stg8(dres + rbase + (int64_t)p * d.dres_pitch + c0, pack8(dsv));
      }
    }
    for (int o = G; o < 32; o <<= 1)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a1[i] += __shfl_xor_sync(0xffffffffu, a1[i], o);
        a2[i] += __shfl_xor_sync(0xffffffffu, a2[i], o);
        a3[i] += __shfl_xor_sync(0xffffffffu, a3[i], o);
      }
    if (lane < G)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        atomicAdd(&s_S1[c0 + i], a1[i]);
        atomicAdd(&s_S2[c0 + i], a2[i]);
        if (has_cbam) atomicAdd(&s_dgc[c0 + i], a3[i]);
      }
  }
  if (CLn > 1) cl.sync(); else __syncthreads();
  if (stage == 1) {                 // one CTA per sample (CLn == 1): the totals are complete
    for (int c = t; c < C; c += CLT)
      *reinterpret_cast<float4*>(d.bwd_nc + ((int64_t)n * C + c) * BN_W) = make_float4(s_dgc[c], s_S1[c], s_S2[c], 0.f);
    return;
  }

  // ---- phase 4: cluster all-reduce of (dgc, S1, S2); channel-MLP backward; dgamma / dbeta; InstanceNorm means
  for (int c = t; c < C && stage == 0; c += CLT) {
    const float gc = s_gc[c];
    const float dgc = CLn > 1 ? cl_sum(cl, s_dgc, c, CLn) : s_dgc[c];
    s_dv[c] = has_cbam ? dgc * gc * (1.f - gc) : 0.f;
    s_m1[c] = CLn > 1 ? cl_sum(cl, s_S1, c, CLn) : s_S1[c];       // totals, turned into the means below
    s_m2[c] = CLn > 1 ? cl_sum(cl, s_S2, c, CLn) : s_S2[c];
    s_dmx[c] = q0[c * NC_W + NC_EXTU];                             // staging: max-pooled u for the hidden layer
  }
  if (t < 64) s_dh[t] = 0.f;
  if (CLn > 1) cl.sync(); else __syncthreads();                    // remote reads done; partial arrays may be reused
  if (has_cbam && stage == 0) {
    mlp_hidden(d.w1, C, Cr, s_b, s_dmx, s_ha, s_hm);
    mlp_cols_dot(d.w2, C, Cr, s_dv, s_dh);
    __syncthreads();
    if (rank == 0)
      for (int j = t; j < Cr; j += CLT) {
        float* hq = d.bwd_h + (int64_t)n * 192;
        hq[j] = fmaxf(s_ha[j], 0.f) + fmaxf(s_hm[j], 0.f);
        hq[64 + j] = s_ha[j] > 0.f ? s_dh[j] : 0.f;
        hq[128 + j] = s_hm[j] > 0.f ? s_dh[j] : 0.f;
      }
  }
  const float inv = 1.f / (float)HW;
  for (int c = t; c < C && stage == 0; c += CLT) {
    float d_avg = 0.f, d_mx = 0.f;
    if (has_cbam) {
#pragma unroll 8
      for (int j = 0; j < Cr; ++j) {
        const float w = __ldg(d.w1 + (int64_t)j * C + c);
        if (s_ha[j] > 0.f) d_avg += w * s_dh[j];
        if (s_hm[j] > 0.f) d_mx += w * s_dh[j];
      }
      if (rank == 0) d.bwd_nc[((int64_t)n * C + c) * BN_W + BN_DGC] = s_dv[c];      // consumed by nb_bwd_w_kernel
    }
    const float ext_uhat = q0[c * NC_W + NC_EXTUHAT];
    const float S1 = s_m1[c] + d_mx;
    const float S2 = s_m2[c] + d_mx * ext_uhat;
    if (rank == 0) {
      atomicAdd(d.dbeta + c, S1 + d_avg);
      atomicAdd(d.dgamma + c, S2);
    }
    const float a = s_a[c];
    s_dmx[c] = a * d_mx; s_m1[c] = a * S1 * inv; s_m2[c] = a * S2 * inv;
  }
  __syncthreads();

  // ---- phase 5: dy = a*du + [p == argmax] a*d_mx - a*m1 - uhat*(a*m2)
  for (int lb = pl; lb < np; lb += 2 * PL) {
    bf16x8 ru[2], ro[2], rd[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int lp = lb + u * PL;
      if (lp < np) {
        const int p = p_lo + lp;
        ru[u] = ldg8(uhat + ubase + (int64_t)p * C + c0);
        ro[u] = ldg8(out + obase + (int64_t)p * d.out_pitch + c0);
        rd[u] = ldg8(dout + dbase + (int64_t)p * d.dout_pitch + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int lp = lb + u * PL;
      if (lp >= np) break;
      const int p = p_lo + lp;
      float uh[8], o[8], dd[8], du[8], dsv[8], dspu[8], r[8];
      unpack8(ru[u], uh); unpack8(ro[u], o); unpack8(rd[u], dd);
      nb_du8<false>(uh, o, dd, c0, s_gc + c0, s_gc + c0, s_gc + c0, has_cbam, res_mode, slope, s_gs[lp], s_dmean[lp],
                    s_dmax[lp], s_cidx[lp], du, dsv, dspu);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float extra = (p == s_idx[c0 + i]) ? s_dmx[c0 + i] : 0.f;
        r[i] = s_a[c0 + i] * du[i] + extra - s_m1[c0 + i] - uh[i] * s_m2[c0 + i];
      }
      stg8(dy + ybase + (int64_t)p * d.dy_pitch + c0, pack8(r));
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Channel-attention MLP BATCHED over samples (small maps: 256..1024 channels, one CTA per sample everywhere else).
// Inside the per-sample kernels the two GEMVs re-read W1 / W2 (up to 2 x 256 KB) from L2 once per sample and were
// 30-50 % of their time; here a CTA owns MLP_NS samples, so every weight element fetched is used MLP_NS times and the
// loads of one row are all in flight together.
// ---------------------------------------------------------------------------------------------------
#ifndef BVAE_MLP_NS
#define BVAE_MLP_NS 4
#endif
constexpr int MLP_NS = BVAE_MLP_NS;

// hidden pre-activations: ha[j] = W1[j,:] . beta (the average pool of an instance-normalised map is beta),
// hm[s][j] = W1[j,:] . mx[s,:].  warp per hidden unit, lanes over channels.
__device__ __forceinline__ void mlpb_hidden(const float* __restrict__ w1, int C, int Cr, const float* s_beta,
                                            const float* s_mx /* [MLP_NS][C] */, float* s_ha, float* s_hm /* [MLP_NS][64] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int j = warp; j < Cr; j += nw) {
    const float* row = w1 + (int64_t)j * C;
    float pa = 0.f, pm[MLP_NS];
#pragma unroll
    for (int s = 0; s < MLP_NS; ++s) pm[s] = 0.f;
    if (C >= 128) {
#pragma unroll 8
      for (int c = lane * 4; c < C; c += 128) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(row + c));
        const float4 b = *reinterpret_cast<const float4*>(s_beta + c);
        pa += w.x * b.x + w.y * b.y + w.z * b.z + w.w * b.w;
#pragma unroll
        for (int s = 0; s < MLP_NS; ++s) {
          const float4 m = *reinterpret_cast<const float4*>(s_mx + s * C + c);
          pm[s] += w.x * m.x + w.y * m.y + w.z * m.z + w.w * m.w;
        }
      }
    } else {
      for (int c = lane; c < C; c += 32) {
        const float w = __ldg(row + c);
        pa += w * s_beta[c];
#pragma unroll
        for (int s = 0; s < MLP_NS; ++s) pm[s] += w * s_mx[s * C + c];
      }
    }
    pa = warp_sum(pa);
#pragma unroll
    for (int s = 0; s < MLP_NS; ++s) pm[s] = warp_sum(pm[s]);
    if (lane == 0) {
      s_ha[j] = pa;
#pragma unroll
      for (int s = 0; s < MLP_NS; ++s) s_hm[s * 64 + j] = pm[s];
    }
  }
}

// gc[n][c] = sigmoid(W2[c,:] . (relu(ha) + relu(hm[n])))      grid ceil(N / MLP_NS), block 256, smem (MLP_NS + 1) * C floats
__global__ void __launch_bounds__(256) nb_mlp_fwd_kernel(int N, int C, int Cr, const float* __restrict__ w1,
                                                         const float* __restrict__ w2, const float* __restrict__ beta,
                                                         float* __restrict__ nc) {
  extern __shared__ float sm[];
  float* s_beta = sm;                 // [C]
  float* s_mx = sm + C;               // [MLP_NS][C]
  __shared__ float s_ha[64], s_hm[MLP_NS * 64], s_h[MLP_NS * 64];
  const int n0 = blockIdx.x * MLP_NS, t = threadIdx.x;
  for (int c = t; c < C; c += 256) {
    s_beta[c] = beta[c];
#pragma unroll
    for (int s = 0; s < MLP_NS; ++s)
      s_mx[s * C + c] = n0 + s < N ? nc[((int64_t)(n0 + s) * C + c) * NC_W + NC_EXTU] : 0.f;
  }
  __syncthreads();
  mlpb_hidden(w1, C, Cr, s_beta, s_mx, s_ha, s_hm);
  __syncthreads();
  for (int i = t; i < MLP_NS * Cr; i += 256) {
    const int s = i / Cr, j = i % Cr;
    s_h[s * 64 + j] = fmaxf(s_ha[j], 0.f) + fmaxf(s_hm[s * 64 + j], 0.f);
  }
  __syncthreads();
  for (int c = t; c < C; c += 256) {
    const float* row = w2 + (int64_t)c * Cr;
    float v[MLP_NS];
#pragma unroll
    for (int s = 0; s < MLP_NS; ++s) v[s] = 0.f;
    if (Cr >= 4) {
#pragma unroll 16
      for (int j = 0; j < Cr; j += 4) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(row + j));
#pragma unroll
        for (int s = 0; s < MLP_NS; ++s)
          v[s] += w.x * s_h[s * 64 + j] + w.y * s_h[s * 64 + j + 1] + w.z * s_h[s * 64 + j + 2] + w.w * s_h[s * 64 + j + 3];
      }
    } else {
      for (int j = 0; j < Cr; ++j) {
        const float w = __ldg(row + j);
#pragma unroll
        for (int s = 0; s < MLP_NS; ++s) v[s] += w * s_h[s * 64 + j];
      }
    }
#pragma unroll
    for (int s = 0; s < MLP_NS; ++s)
      if (n0 + s < N) nc[((int64_t)(n0 + s) * C + c) * NC_W + NC_GC] = 1.f / (1.f + expf(-v[s]));
  }
}

// Backward of the same MLP for MLP_NS samples, followed by the per-(n,c) InstanceNorm-backward coefficients.
// in : bwd_nc[n][c] = {dgc, S1, S2, -} (totals from the reduction sweeps), nc
// out: bwd_nc[n][c] = {dv, a*m1, a*m2, a*d_mx}, bwd_h[n] (for nb_bwd_w_kernel), dgamma / dbeta (atomics, one per CTA)
__global__ void __launch_bounds__(256) nb_mlp_bwd_kernel(int N, int HW, int C, int Cr, const float* __restrict__ w1,
                                                         const float* __restrict__ w2, const float* __restrict__ beta,
                                                         const float* __restrict__ nc, float* __restrict__ bwd_nc,
                                                         float* __restrict__ bwd_h, float* __restrict__ dgamma,
                                                         float* __restrict__ dbeta) {
  extern __shared__ float sm[];
  float* s_beta = sm;                 // [C]
  float* s_mx = sm + C;               // [MLP_NS][C]
  float* s_dv = sm + (1 + MLP_NS) * C;   // [MLP_NS][C]
  __shared__ float s_ha[64], s_hm[MLP_NS * 64], s_dh[MLP_NS * 64], s_dha[MLP_NS * 64], s_dhm[MLP_NS * 64];
  const int n0 = blockIdx.x * MLP_NS, t = threadIdx.x;
  const int lane = t & 31, warp = t >> 5;
  for (int c = t; c < C; c += 256) {
    s_beta[c] = beta[c];
#pragma unroll
    for (int s = 0; s < MLP_NS; ++s) {
      float mx = 0.f, dv = 0.f;
      if (n0 + s < N) {
        const int64_t o = (int64_t)(n0 + s) * C + c;
        const float gc = nc[o * NC_W + NC_GC];
        mx = nc[o * NC_W + NC_EXTU];
        dv = bwd_nc[o * BN_W + BN_DGC] * gc * (1.f - gc);
        bwd_nc[o * BN_W + BN_DGC] = dv;                              // consumed by nb_bwd_w_kernel
      }
      s_mx[s * C + c] = mx;
      s_dv[s * C + c] = dv;
    }
  }
  for (int i = t; i < MLP_NS * 64; i += 256) s_dh[i] = 0.f;
  __syncthreads();
  mlpb_hidden(w1, C, Cr, s_beta, s_mx, s_ha, s_hm);
  // dh[s][j] = sum_c W2[c][j] * dv[s][c]: lanes over j, warps over rows, all loads of a trip independent
  {
    float a0[MLP_NS], a1[MLP_NS];
#pragma unroll
    for (int s = 0; s < MLP_NS; ++s) { a0[s] = 0.f; a1[s] = 0.f; }
    const bool l0 = lane < Cr, l1 = lane + 32 < Cr;
#pragma unroll 8
    for (int c = warp; c < C; c += 8) {
      const float w0 = l0 ? __ldg(w2 + (int64_t)c * Cr + lane) : 0.f;
      const float w1v = l1 ? __ldg(w2 + (int64_t)c * Cr + lane + 32) : 0.f;
#pragma unroll
      for (int s = 0; s < MLP_NS; ++s) {
        const float dv = s_dv[s * C + c];
        a0[s] = fmaf(w0, dv, a0[s]);
        a1[s] = fmaf(w1v, dv, a1[s]);
      }
    }
#pragma unroll
    for (int s = 0; s < MLP_NS; ++s) {
      if (l0) atomicAdd(&s_dh[s * 64 + lane], a0[s]);
      if (l1) atomicAdd(&s_dh[s * 64 + lane + 32], a1[s]);
    }
  }
  __syncthreads();
  for (int i = t; i < MLP_NS * Cr; i += 256) {
    const int s = i / Cr, j = i % Cr;
    const float ha = s_ha[j], hm = s_hm[s * 64 + j], dh = s_dh[s * 64 + j];
    const float dha = ha > 0.f ? dh : 0.f, dhm = hm > 0.f ? dh : 0.f;
    s_dha[s * 64 + j] = dha; s_dhm[s * 64 + j] = dhm;
    if (n0 + s < N) {
      float* hq = bwd_h + (int64_t)(n0 + s) * 192;
      hq[j] = fmaxf(ha, 0.f) + fmaxf(hm, 0.f);
      hq[64 + j] = dha;
      hq[128 + j] = dhm;
    }
  }
  __syncthreads();
  const float inv = 1.f / (float)HW;
  for (int c = t; c < C; c += 256) {
    float d_avg[MLP_NS], d_mx[MLP_NS];
#pragma unroll
    for (int s = 0; s < MLP_NS; ++s) { d_avg[s] = 0.f; d_mx[s] = 0.f; }
#pragma unroll 8
    for (int j = 0; j < Cr; ++j) {
      const float w = __ldg(w1 + (int64_t)j * C + c);
#pragma unroll
      for (int s = 0; s < MLP_NS; ++s) {
        d_avg[s] = fmaf(w, s_dha[s * 64 + j], d_avg[s]);
        d_mx[s] = fmaf(w, s_dhm[s * 64 + j], d_mx[s]);
      }
    }
    float sum_b = 0.f, sum_g = 0.f;
#pragma unroll
    for (int s = 0; s < MLP_NS; ++s) {
      if (n0 + s >= N) break;
      const int64_t o = (int64_t)(n0 + s) * C + c;
      const float a = nc[o * NC_W + NC_A], ext_uhat = nc[o * NC_W + NC_EXTUHAT];
      const float4 tot = *reinterpret_cast<const float4*>(bwd_nc + o * BN_W);       // {dv, S1, S2, -}
      const float S1 = tot.y + d_mx[s];
      const float S2 = tot.z + d_mx[s] * ext_uhat;
      sum_b += S1 + d_avg[s];
      sum_g += S2;
      *reinterpret_cast<float4*>(bwd_nc + o * BN_W) = make_float4(tot.x, a * S1 * inv, a * S2 * inv, a * d_mx[s]);
    }
    atomicAdd(dbeta + c, sum_b);
    atomicAdd(dgamma + c, sum_g);
  }
}

// ---------------------------------------------------------------------------------------------------
// Per-sample "wide" kernels (nbs_*), round 3: sites whose sample fits one CTA's sweep (H*W <= 1440, C = 32 .. 1024, fp32
// raw input).  The per-sample nb_cl_* kernels above are latency bound (ncu on the small maps: 60-75 % of the warp samples
// wait on the long scoreboard with 14-24 resident warps per SM, 3-15 % of the DRAM peak): one 128-thread CTA walks a
// sample through five barrier-separated phases, each a chain of dependent L2 round trips, with shared-memory atomics
// (64-bit CAS loops for the arg-max keys) in between; the tiled nb_* / nbf_* kernels pay 5-6 dependent launches, a memset
// and global atomics per site, which is most of their time on the medium maps (C128 24x15: 233 us forward against a
// 58 us traffic floor).  Here
//   * a thread owns FOUR channels (one 16-byte fp32 / 8-byte bf16 access per pixel) whose per-channel constants live
//     in registers; LPP = min(32, C/4) lanes cover a pixel's channels (or 128 of them: CW = C/128 channel warps for
//     C > 128), so a 256-thread CTA has PLn = (8/CW) * (32/LPP) pixel lanes, and every pixel loop issues the loads of
//     four pixels before the first use;
//   * per-pixel reductions over channels are one shuffle tree over LPP lanes plus CW partials in shared memory,
//     per-channel reductions over pixels are registers plus one partial per pixel lane in shared memory -- every slot
//     has exactly one writer and the partials are summed in a fixed order: no atomics, deterministic by construction;
//   * the statistics kernel is not per sample: a CTA owns (sample, <= 128 channels), its 8 warps split the pixels;
//   * everything after the (batched) channel MLP is ONE kernel per sample: sweep, 3x3 gate conv from shared memory,
//     second sweep from L1/L2 -- forward 3 launches instead of 5, backward 4 instead of 7, no memset.
// Scratch layouts (nc, nc_idx, sa, cidx, gs, bwd_nc, bwd_px, bwd_h) are those of the nb_cl_* kernels, so the batched
// channel MLP (nb_mlp_fwd / nb_mlp_bwd), the MLP weight-gradient kernels and the tests' state readers are unchanged.
// ---------------------------------------------------------------------------------------------------
constexpr int NBS_T = 256;
constexpr int NBS_MAX_HW = 1440;

__device__ __forceinline__ void nbs_unpack4(const uint2 r, float* f) {
  f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
  f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
}
__device__ __forceinline__ uint2 nbs_ld4(const bf16* p) { return *reinterpret_cast<const uint2*>(p); }
__device__ __forceinline__ void nbs_st4(bf16* p, const float* f) {
  *reinterpret_cast<uint2*>(p) = make_uint2(pack2(f[0], f[1]), pack2(f[2], f[3]));
}

// thread geometry of a 256-thread CTA over a [pixels][C] tile
struct NbsGeo {
  int LPP;    // lanes per pixel inside a warp: min(32, C/4)
  int CW;     // channel warps per pixel: max(1, C/128)
  int PLn;    // pixel lanes of the CTA
  int cwi;    // this thread's channel warp
  int lcl;    // lane inside the pixel's lane group
  int pl;     // this thread's pixel lane
  int pl0;    // first pixel lane of this WARP (warp-uniform loop base)
  int c;      // first of this thread's 4 channels
};
__device__ __forceinline__ NbsGeo nbs_geo(int C) {
  NbsGeo g;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  g.LPP = C >= 128 ? 32 : (C >> 2);
  g.CW = C >= 128 ? (C >> 7) : 1;
  const int PPW = 32 / g.LPP;
  g.PLn = (8 / g.CW) * PPW;
  g.cwi = w % g.CW;
  g.lcl = lane % g.LPP;
  g.pl0 = (w / g.CW) * PPW;
  g.pl = g.pl0 + lane / g.LPP;
  g.c = g.cwi * 128 + g.lcl * 4;
  return g;
}

// forward 1: InstanceNorm statistics and coefficients.  grid (max(1, C/128), N), block 256: the CTA owns CC = min(C, 128)
// channels of one sample, its 8 * (128/CC) pixel lanes sweep pixels lane, lane + lanes, ...
// EXT = false (sites without CBAM): no extrema, the pooled value is reported as the mean like nb_coef_kernel does.
template <bool EXT>
__global__ void __launch_bounds__(NBS_T) nbs_stats_kernel(const float* __restrict__ y, int y_pitch, int HW, int C,
                                                          const float* __restrict__ gamma, const float* __restrict__ beta,
                                                          float eps, float* __restrict__ nc, int32_t* __restrict__ nc_idx) {
  __shared__ __align__(16) float s_sum[1024], s_sq[1024];            // [pixel lane][CC]: lanes * CC == 1024
  __shared__ __align__(16) u64 s_kx[EXT ? 1024 : 2], s_kn[EXT ? 1024 : 2];
  const int CC = C < 128 ? C : 128, LPP = CC >> 2, PLs = 1024 / CC;
  const int n = blockIdx.y, cb = blockIdx.x * 128;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int lcl = lane % LPP, plane = w * (32 / LPP) + lane / LPP;
  const float* base = y + (int64_t)n * HW * y_pitch + cb + lcl * 4;
  const float4 sh4 = __ldg(reinterpret_cast<const float4*>(base));          // pixel 0: the shift of the shifted sums
  const float sh[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
  float sum[4], sq[4], vmx[4], vmn[4];
  int imx[4], imn[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { sum[i] = 0.f; sq[i] = 0.f; vmx[i] = -INFINITY; vmn[i] = INFINITY; imx[i] = 0; imn[i] = 0; }
  for (int p0 = plane; p0 < HW; p0 += 4 * PLs) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + PLs * u;
      if (p < HW) v[u] = __ldg(reinterpret_cast<const float4*>(base + (int64_t)p * y_pitch));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + PLs * u;
      if (p < HW) {
        const float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float dl = f[i] - sh[i];
          sum[i] += dl; sq[i] += dl * dl;
          if (EXT) {
            if (f[i] > vmx[i]) { vmx[i] = f[i]; imx[i] = p; }      // p ascends inside a thread: strict > keeps the first
            if (f[i] < vmn[i]) { vmn[i] = f[i]; imn[i] = p; }
          }
        }
      }
    }
  }
  const int slot = plane * CC + lcl * 4;
  *reinterpret_cast<float4*>(s_sum + slot) = make_float4(sum[0], sum[1], sum[2], sum[3]);
  *reinterpret_cast<float4*>(s_sq + slot) = make_float4(sq[0], sq[1], sq[2], sq[3]);
  if (EXT) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {       // pixel lanes that saw no pixel contribute the neutral key 0
      s_kx[slot + i] = vmx[i] == -INFINITY ? 0ull : make_key(vmx[i], (uint32_t)imx[i]);
      s_kn[slot + i] = vmn[i] == INFINITY ? 0ull : make_key(-vmn[i], (uint32_t)imn[i]);
    }
  }
  __syncthreads();
  if ((int)threadIdx.x < CC) {
    const int cl = threadIdx.x, c = cb + cl;
    float tsum = 0.f, tsq = 0.f;
    u64 kx = 0, kn = 0;
    for (int j = 0; j < PLs; ++j) {                                          // fixed order: deterministic
      tsum += s_sum[j * CC + cl];
      tsq += s_sq[j * CC + cl];
      if (EXT) {
        const u64 a = s_kx[j * CC + cl], b = s_kn[j * CC + cl];
        kx = a > kx ? a : kx;
        kn = b > kn ? b : kn;
      }
    }
    const float inv = 1.f / (float)HW;
    const float shift = __ldg(y + (int64_t)n * HW * y_pitch + c);
    const float md = tsum * inv;
    const float mean = shift + md;
    const float var = fmaxf(tsq * inv - md * md, 0.f);
    const float rstd = rsqrtf(var + eps);
    const float g = gamma[c], b0 = beta[c];
    const float a = g * rstd, b = b0 - mean * a;
    float yext = mean;
    uint32_t idx = 0;
    if (EXT) {
      if (a >= 0.f) { yext = key_val(kx); idx = key_idx(kx); }
      else { yext = -key_val(kn); idx = key_idx(kn); }
    }
    const float ext_uhat = (yext - mean) * rstd;
    const float ext_u = g * ext_uhat + b0;
    const int64_t o = (int64_t)n * C + c;
    float4* q = reinterpret_cast<float4*>(nc + o * NC_W);
    q[0] = make_float4(mean, rstd, a, b);
    q[1] = make_float4(1.f, ext_u, ext_uhat, 0.f);                           // {gc, ext_u, ext_uhat, -}
    nc_idx[o] = (int32_t)idx;
  }
}

// 3x3 attention conv on the [mean, max] map held in shared memory (zero padding), one pixel; same order as sa_conv
__device__ __forceinline__ float nbs_sa_conv(const float2* s_sa, const float* s_w, int H, int W, int py, int px) {
  float q = 0.f;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int yy = py + ky - 1;
    if (yy < 0 || yy >= H) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int xx = px + kx - 1;
      if (xx < 0 || xx >= W) continue;
      const float2 v = s_sa[yy * W + xx];
      q += s_w[ky * 3 + kx] * v.x + s_w[9 + ky * 3 + kx] * v.y;
    }
  }
  return q;
}

// forward 2 (after the batched channel MLP on CBAM sites): one CTA per sample.
//   without CBAM: uhat, out = act(u) in one sweep.
//   with CBAM:    sweep 1 writes uhat and the per-pixel mean / max / arg-max over channels of u*gc; the 3x3 gate conv
//                 runs from shared memory; sweep 2 re-reads the raw tile (L1/L2) and writes out = act(r + u*gc*gs).
// MODE = res_mode of the CBAM sites (1 self, 2 external, 3 none); ignored without CBAM.
// dynamic smem (CBAM): u64 s_pk[HW*CW] | float2 s_sa[HW] | float s_ps[HW*CW] | float s_gs[HW]
static size_t nbs_fwd_smem(int HW, int C) { const int CW = C >= 128 ? C / 128 : 1; return (size_t)HW * CW * 12 + (size_t)HW * 12; }
template <bool CBAM, int MODE>
__global__ void __launch_bounds__(NBS_T, 4) nbs_fwd_kernel(const bvae_nb_desc d) {
  extern __shared__ __align__(16) unsigned char nbs_dyn[];
  __shared__ float s_w[18];
  const int C = d.C, H = d.H, W = d.W, HW = H * W;
  const NbsGeo g = nbs_geo(C);
  const int CW = g.CW, PLn = g.PLn, c = g.c;
  u64* s_pk = reinterpret_cast<u64*>(nbs_dyn);
  float2* s_sa = reinterpret_cast<float2*>(s_pk + (CBAM ? HW * CW : 0));
  float* s_ps = reinterpret_cast<float*>(s_sa + (CBAM ? HW : 0));
  float* s_gs = s_ps + (CBAM ? HW * CW : 0);
  const int n = blockIdx.x, t = threadIdx.x;
  float h_r[4], h_nm[4], h_a[4], h_b[4], h_gc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4* q = reinterpret_cast<const float4*>(d.nc + ((int64_t)n * C + c + i) * NC_W);
    const float4 lo = __ldg(q);                                              // {mean, rstd, a, b}
    h_r[i] = lo.y; h_nm[i] = -lo.x * lo.y; h_a[i] = lo.z; h_b[i] = lo.w;
    h_gc[i] = CBAM ? __ldg(q + 1).x : 1.f;
  }
  if (CBAM && t < 18) s_w[t] = d.wsp[t];
  const float slope = d.slope;
  const int y_pitch = d.y_pitch, out_pitch = d.out_pitch;
  const float* yb = (const float*)d.y + (int64_t)n * HW * y_pitch + c;
  bf16* ub = d.uhat != nullptr ? (bf16*)d.uhat + (int64_t)n * HW * C + c : nullptr;
  bf16* ob = (bf16*)d.out + (int64_t)n * HW * out_pitch + c;

  for (int pb = g.pl0; pb < HW; pb += 4 * PLn) {        // warp-uniform bounds: the shuffles below need every lane
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = pb + (g.pl - g.pl0) + u * PLn;
      if (p < HW) v[u] = __ldg(reinterpret_cast<const float4*>(yb + (int64_t)p * y_pitch));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = pb + (g.pl - g.pl0) + u * PLn;
      const bool valid = p < HW;
      if (pb + u * PLn >= HW) break;                    // warp-uniform: no lane of this warp has a pixel left
      float sum = 0.f, mx = -INFINITY;
      int mxc = 0;
      if (valid) {
        const float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        float uh[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) uh[i] = fmaf(f[i], h_r[i], h_nm[i]);
        if (ub != nullptr) nbs_st4(ub + (int64_t)p * C, uh);                 // inference: nothing is saved
        if (CBAM) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float u1 = fmaf(h_a[i], f[i], h_b[i]) * h_gc[i];
            sum += u1;
            if (u1 > mx) { mx = u1; mxc = c + i; }
          }
        } else {
          float o[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) o[i] = act_fwd(fmaf(h_a[i], f[i], h_b[i]), slope);
          nbs_st4(ob + (int64_t)p * out_pitch, o);
        }
      }
      if (CBAM) {
        for (int o = g.LPP >> 1; o > 0; o >>= 1) {
          sum += __shfl_xor_sync(0xffffffffu, sum, o);
          const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
          const int omc = __shfl_xor_sync(0xffffffffu, mxc, o);
          if (omx > mx || (omx == mx && omc < mxc)) { mx = omx; mxc = omc; }
        }
        if (valid && g.lcl == 0) { s_ps[p * CW + g.cwi] = sum; s_pk[p * CW + g.cwi] = make_key(mx, (uint32_t)mxc); }
      }
    }
  }
  if (!CBAM) return;
  __syncthreads();
  for (int p = t; p < HW; p += NBS_T) {
    float s = 0.f;
    u64 k = 0;
    for (int j = 0; j < CW; ++j) { s += s_ps[p * CW + j]; const u64 o = s_pk[p * CW + j]; k = o > k ? o : k; }
    const float2 v = make_float2(s / (float)C, key_val(k));
    s_sa[p] = v;
    const int64_t o = (int64_t)n * HW + p;
    reinterpret_cast<float2*>(d.sa)[o] = v;
    d.cidx[o] = (int32_t)key_idx(k);
  }
  __syncthreads();
  for (int p = t; p < HW; p += NBS_T) {
    const float gg = 1.f / (1.f + expf(-nbs_sa_conv(s_sa, s_w, H, W, p / W, p % W)));
    s_gs[p] = gg;
    d.gs[(int64_t)n * HW + p] = gg;
  }
  __syncthreads();
  const int res_pitch = d.res_pitch;
  const bf16* rb = MODE == 2 ? (const bf16*)d.res + (int64_t)n * HW * res_pitch + c : nullptr;
  for (int p0 = g.pl; p0 < HW; p0 += 4 * PLn) {
    float4 v[4];
    uint2 rr[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + u * PLn;
      if (p < HW) {
        v[u] = __ldg(reinterpret_cast<const float4*>(yb + (int64_t)p * y_pitch));
        if (MODE == 2) rr[u] = nbs_ld4(rb + (int64_t)p * res_pitch);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + u * PLn;
      if (p < HW) {
        const float f[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
        const float gg = s_gs[p];
        float r[4], o[4];
        if (MODE == 2) nbs_unpack4(rr[u], r);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float uu = fmaf(h_a[i], f[i], h_b[i]);
          const float k = h_gc[i] * gg;
          const float tt = MODE == 1 ? fmaf(uu, k, uu) : (MODE == 2 ? fmaf(uu, k, r[i]) : uu * k);
          o[i] = act_fwd(tt, slope);
        }
        nbs_st4(ob + (int64_t)p * out_pitch, o);
      }
    }
  }
}

// backward 1 (CBAM sites): the reduction sweeps of one sample.  Sweep 1: dgs[p] = sum_c ds*u*gc, dgc[c] += sum_p ds*u*gs;
// then dq, the 3x3 transpose conv (dmean, dmax) and the attention-conv weight gradient from shared memory; sweep 2:
// S1 = sum du, S2 = sum du*uhat, dgc += sum dsp*u, dres.  Output: bwd_nc = {dgc, S1, S2, 0}, bwd_px = {dq, dmean/C, dmax, 0}
// (consumed by nb_mlp_bwd_kernel and nbs_bwd2_kernel).
// dynamic smem: float s_pp[HW*CW] | s_gs[HW] | s_dq[HW] | s_dmean[HW] | s_dmax[HW] | int s_cidx[HW]
static size_t nbs_bwd1_smem(int HW, int C) { const int CW = C >= 128 ? C / 128 : 1; return (size_t)HW * CW * 4 + (size_t)HW * 20; }
template <int MODE>
__global__ void __launch_bounds__(NBS_T, 4) nbs_bwd1_kernel(const bvae_nb_desc d) {
  extern __shared__ __align__(16) unsigned char nbs_dyn[];
  __shared__ float4 s_red[3][256];              // [sum][pixel lane * (C/4) + channel quad]: PLn * C/4 == 256
  __shared__ float s_w[18], s_dw[18];
  const int C = d.C, H = d.H, W = d.W, HW = H * W;
  const NbsGeo g = nbs_geo(C);
  const int CW = g.CW, PLn = g.PLn, c = g.c;
  float* s_pp = reinterpret_cast<float*>(nbs_dyn);
  float* s_gs = s_pp + HW * CW;
  float* s_dq = s_gs + HW;
  float* s_dmean = s_dq + HW;
  float* s_dmax = s_dmean + HW;
  int* s_cidx = reinterpret_cast<int*>(s_dmax + HW);
  const int n = blockIdx.x, t = threadIdx.x, lane = t & 31;
  const float slope = d.slope;
  float h_g[4], h_b[4], h_gc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h_g[i] = __ldg(d.gamma + c + i); h_b[i] = __ldg(d.beta + c + i);
    h_gc[i] = __ldg(d.nc + ((int64_t)n * C + c + i) * NC_W + NC_GC);
  }
  for (int p = t; p < HW; p += NBS_T) { s_gs[p] = d.gs[(int64_t)n * HW + p]; s_cidx[p] = d.cidx[(int64_t)n * HW + p]; }
  if (t < 18) { s_w[t] = d.wsp[t]; s_dw[t] = 0.f; }
  __syncthreads();
  const int dout_pitch = d.dout_pitch, out_pitch = d.out_pitch, dres_pitch = d.dres_pitch;
  const bf16* ub = (const bf16*)d.uhat + (int64_t)n * HW * C + c;
  const bf16* ob = (const bf16*)d.out + (int64_t)n * HW * out_pitch + c;
  const bf16* db = (const bf16*)d.dout + (int64_t)n * HW * dout_pitch + c;
  bf16* rb = (MODE == 2 && d.dres != nullptr) ? (bf16*)d.dres + (int64_t)n * HW * dres_pitch + c : nullptr;

  float acc[4], a1[4], a2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { acc[i] = 0.f; a1[i] = 0.f; a2[i] = 0.f; }
  // ---- sweep 1
  for (int pb = g.pl0; pb < HW; pb += 4 * PLn) {        // warp-uniform bounds (shuffles)
    uint2 ru[4], ro[4], rd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = pb + (g.pl - g.pl0) + u * PLn;
      if (p < HW) {
        ru[u] = nbs_ld4(ub + (int64_t)p * C);
        ro[u] = nbs_ld4(ob + (int64_t)p * out_pitch);
        rd[u] = nbs_ld4(db + (int64_t)p * dout_pitch);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = pb + (g.pl - g.pl0) + u * PLn;
      const bool valid = p < HW;
      if (pb + u * PLn >= HW) break;                    // warp-uniform
      float dgs = 0.f;
      if (valid) {
        float uh[4], o[4], dd[4];
        nbs_unpack4(ru[u], uh); nbs_unpack4(ro[u], o); nbs_unpack4(rd[u], dd);
        const float gg = s_gs[p];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float ds = dd[i] * (o[i] > 0.f ? 1.f : slope);
          const float tt = ds * fmaf(h_g[i], uh[i], h_b[i]);
          dgs += tt * h_gc[i];
          acc[i] += tt * gg;
        }
      }
      for (int o = g.LPP >> 1; o > 0; o >>= 1) dgs += __shfl_xor_sync(0xffffffffu, dgs, o);
      if (valid && g.lcl == 0) s_pp[p * CW + g.cwi] = dgs;
    }
  }
  __syncthreads();
  for (int p = t; p < HW; p += NBS_T) {
    float dgs = 0.f;
    for (int j = 0; j < CW; ++j) dgs += s_pp[p * CW + j];
    const float gg = s_gs[p];
    s_dq[p] = dgs * gg * (1.f - gg);
  }
  __syncthreads();
  // ---- 3x3 transpose conv of dq and the attention-conv weight gradient
  {
    float part[18];
#pragma unroll
    for (int i = 0; i < 18; ++i) part[i] = 0.f;
    const float* sa_n = d.sa + (int64_t)n * HW * 2;
    for (int p = t; p < HW; p += NBS_T) {
      const int py = p / W, px = p % W;
      const float dq = s_dq[p];
      float dmean = 0.f, dmax = 0.f;
#pragma unroll
      for (int ky = 0; ky < 3; ++ky)
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const int yy = py - (ky - 1), xx = px - (kx - 1);
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            const float dqq = s_dq[yy * W + xx];
            dmean += s_w[ky * 3 + kx] * dqq;
            dmax += s_w[9 + ky * 3 + kx] * dqq;
          }
          const int y2 = py + ky - 1, x2 = px + kx - 1;
          if (y2 >= 0 && y2 < H && x2 >= 0 && x2 < W) {
            const float2 v = *reinterpret_cast<const float2*>(sa_n + ((int64_t)y2 * W + x2) * 2);
            part[ky * 3 + kx] += dq * v.x;
            part[9 + ky * 3 + kx] += dq * v.y;
          }
        }
      dmean /= (float)C;
      s_dmean[p] = dmean;
      s_dmax[p] = dmax;
      reinterpret_cast<float4*>(d.bwd_px)[(int64_t)n * HW + p] = make_float4(dq, dmean, dmax, 0.f);
    }
    if (t < ((HW + 31) & ~31)) {                        // whole warps: the ones that own at least one pixel
#pragma unroll
      for (int i = 0; i < 18; ++i) {
        const float v = warp_sum(part[i]);
        if (lane == 0 && v != 0.f) atomicAdd(&s_dw[i], v);
      }
    }
  }
  __syncthreads();
  if (t < 18) atomicAdd(d.dwsp + t, s_dw[t]);
  // ---- sweep 2 (the sample's tensors come back from L1 / L2)
  for (int p0 = g.pl; p0 < HW; p0 += 4 * PLn) {
    uint2 ru[4], ro[4], rd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + u * PLn;
      if (p < HW) {
        ru[u] = nbs_ld4(ub + (int64_t)p * C);
        ro[u] = nbs_ld4(ob + (int64_t)p * out_pitch);
        rd[u] = nbs_ld4(db + (int64_t)p * dout_pitch);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + u * PLn;
      if (p < HW) {
        float uh[4], o[4], dd[4], dsv[4];
        nbs_unpack4(ru[u], uh); nbs_unpack4(ro[u], o); nbs_unpack4(rd[u], dd);
        const float gg = s_gs[p], dm = s_dmean[p], dx = s_dmax[p];
        const int ci = s_cidx[p];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float ds = dd[i] * (o[i] > 0.f ? 1.f : slope);
          dsv[i] = ds;
          const float dsp = dm + ((c + i) == ci ? dx : 0.f);           // grad wrt u*gc from the spatial branch
          const float v = (MODE == 1 ? ds : 0.f) + ds * h_gc[i] * gg + dsp * h_gc[i];
          a1[i] += v;
          a2[i] += v * uh[i];
          acc[i] += dsp * fmaf(h_g[i], uh[i], h_b[i]);
        }
        if (rb != nullptr) nbs_st4(rb + (int64_t)p * dres_pitch, dsv);
      }
    }
  }
  const int slot = g.pl * (C >> 2) + (c >> 2);
  s_red[0][slot] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  s_red[1][slot] = make_float4(a1[0], a1[1], a1[2], a1[3]);
  s_red[2][slot] = make_float4(a2[0], a2[1], a2[2], a2[3]);
  __syncthreads();
  for (int cc = t; cc < C; cc += NBS_T) {
    float dgc = 0.f, S1 = 0.f, S2 = 0.f;
    for (int j = 0; j < PLn; ++j) {                                     // fixed order
      dgc += reinterpret_cast<const float*>(&s_red[0][0])[j * C + cc];
      S1 += reinterpret_cast<const float*>(&s_red[1][0])[j * C + cc];
      S2 += reinterpret_cast<const float*>(&s_red[2][0])[j * C + cc];
    }
    *reinterpret_cast<float4*>(d.bwd_nc + ((int64_t)n * C + cc) * BN_W) = make_float4(dgc, S1, S2, 0.f);
  }
}

// backward 2 (CBAM sites, after nb_mlp_bwd_kernel): dy = a*du + [p == argmax] a*d_mx - a*m1 - uhat*(a*m2)
// dynamic smem: float s_gs[HW] | s_dmean[HW] | s_dmax[HW] | int s_cidx[HW]
template <int MODE>
__global__ void __launch_bounds__(NBS_T, 4) nbs_bwd2_kernel(const bvae_nb_desc d) {
  extern __shared__ __align__(16) unsigned char nbs_dyn[];
  const int C = d.C, HW = d.H * d.W;
  const NbsGeo g = nbs_geo(C);
  const int PLn = g.PLn, c = g.c;
  float* s_gs = reinterpret_cast<float*>(nbs_dyn);
  float* s_dmean = s_gs + HW;
  float* s_dmax = s_dmean + HW;
  int* s_cidx = reinterpret_cast<int*>(s_dmax + HW);
  const int n = blockIdx.x, t = threadIdx.x;
  const float slope = d.slope;
  float h_gc[4], h_a[4], h_m1[4], h_m2[4], h_dmx[4];
  int h_idx[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t o = (int64_t)n * C + c + i;
    const float4* q = reinterpret_cast<const float4*>(d.nc + o * NC_W);
    h_a[i] = __ldg(q).z; h_gc[i] = __ldg(q + 1).x;
    const float4 bw = *reinterpret_cast<const float4*>(d.bwd_nc + o * BN_W);     // {dv, a*m1, a*m2, a*d_mx}
    h_m1[i] = bw.y; h_m2[i] = bw.z; h_dmx[i] = bw.w;
    h_idx[i] = d.nc_idx[o];
  }
  for (int p = t; p < HW; p += NBS_T) {
    const int64_t o = (int64_t)n * HW + p;
    s_gs[p] = d.gs[o]; s_cidx[p] = d.cidx[o];
    const float4 v = reinterpret_cast<const float4*>(d.bwd_px)[o];               // {dq, dmean / C, dmax, -}
    s_dmean[p] = v.y; s_dmax[p] = v.z;
  }
  __syncthreads();
  const int dout_pitch = d.dout_pitch, out_pitch = d.out_pitch, dy_pitch = d.dy_pitch;
  const bf16* ub = (const bf16*)d.uhat + (int64_t)n * HW * C + c;
  const bf16* ob = (const bf16*)d.out + (int64_t)n * HW * out_pitch + c;
  const bf16* db = (const bf16*)d.dout + (int64_t)n * HW * dout_pitch + c;
  bf16* yb = (bf16*)d.dy + (int64_t)n * HW * dy_pitch + c;
  for (int p0 = g.pl; p0 < HW; p0 += 4 * PLn) {
    uint2 ru[4], ro[4], rd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + u * PLn;
      if (p < HW) {
        ru[u] = nbs_ld4(ub + (int64_t)p * C);
        ro[u] = nbs_ld4(ob + (int64_t)p * out_pitch);
        rd[u] = nbs_ld4(db + (int64_t)p * dout_pitch);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + u * PLn;
      if (p < HW) {
        float uh[4], o[4], dd[4], r[4];
        nbs_unpack4(ru[u], uh); nbs_unpack4(ro[u], o); nbs_unpack4(rd[u], dd);
        const float gg = s_gs[p], dm = s_dmean[p], dx = s_dmax[p];
        const int ci = s_cidx[p];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float ds = dd[i] * (o[i] > 0.f ? 1.f : slope);
          const float dsp = dm + ((c + i) == ci ? dx : 0.f);
          const float du = (MODE == 1 ? ds : 0.f) + ds * h_gc[i] * gg + dsp * h_gc[i];
          const float extra = (p == h_idx[i]) ? h_dmx[i] : 0.f;
          r[i] = h_a[i] * du + extra - h_m1[i] - uh[i] * h_m2[i];
        }
        nbs_st4(yb + (int64_t)p * dy_pitch, r);
      }
    }
  }
}

// backward of a site without CBAM, one CTA per sample: S1 = sum du, S2 = sum du*uhat (du = dout * act'), dgamma / dbeta,
// then dy = a*(du - S1/HW - uhat*S2/HW) from the re-read tile.
__global__ void __launch_bounds__(NBS_T, 4) nbs_bwd_plain_kernel(const bvae_nb_desc d) {
  __shared__ float4 s_red[2][256];
  __shared__ float s_m1[1024], s_m2[1024];
  const int C = d.C, HW = d.H * d.W;
  const NbsGeo g = nbs_geo(C);
  const int PLn = g.PLn, c = g.c;
  const int n = blockIdx.x, t = threadIdx.x;
  const float slope = d.slope;
  const int dout_pitch = d.dout_pitch, out_pitch = d.out_pitch, dy_pitch = d.dy_pitch;
  const bf16* ub = (const bf16*)d.uhat + (int64_t)n * HW * C + c;
  const bf16* ob = (const bf16*)d.out + (int64_t)n * HW * out_pitch + c;
  const bf16* db = (const bf16*)d.dout + (int64_t)n * HW * dout_pitch + c;
  bf16* yb = (bf16*)d.dy + (int64_t)n * HW * dy_pitch + c;
  float a1[4], a2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) { a1[i] = 0.f; a2[i] = 0.f; }
  for (int p0 = g.pl; p0 < HW; p0 += 4 * PLn) {
    uint2 ru[4], ro[4], rd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + u * PLn;
      if (p < HW) {
        ru[u] = nbs_ld4(ub + (int64_t)p * C);
        ro[u] = nbs_ld4(ob + (int64_t)p * out_pitch);
        rd[u] = nbs_ld4(db + (int64_t)p * dout_pitch);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + u * PLn;
      if (p < HW) {
        float uh[4], o[4], dd[4];
        nbs_unpack4(ru[u], uh); nbs_unpack4(ro[u], o); nbs_unpack4(rd[u], dd);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float du = dd[i] * (o[i] > 0.f ? 1.f : slope);
          a1[i] += du;
          a2[i] += du * uh[i];
        }
      }
    }
  }
  const int slot = g.pl * (C >> 2) + (c >> 2);
  s_red[0][slot] = make_float4(a1[0], a1[1], a1[2], a1[3]);
  s_red[1][slot] = make_float4(a2[0], a2[1], a2[2], a2[3]);
  __syncthreads();
  const float inv = 1.f / (float)HW;
  for (int cc = t; cc < C; cc += NBS_T) {
    float S1 = 0.f, S2 = 0.f;
    for (int j = 0; j < PLn; ++j) {
      S1 += reinterpret_cast<const float*>(&s_red[0][0])[j * C + cc];
      S2 += reinterpret_cast<const float*>(&s_red[1][0])[j * C + cc];
    }
    atomicAdd(d.dbeta + cc, S1);
    atomicAdd(d.dgamma + cc, S2);
    const float a = d.nc[((int64_t)n * C + cc) * NC_W + NC_A];
    s_m1[cc] = a * S1 * inv;
    s_m2[cc] = a * S2 * inv;
  }
  __syncthreads();
  float h_a[4], h_m1[4], h_m2[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h_a[i] = d.nc[((int64_t)n * C + c + i) * NC_W + NC_A];
    h_m1[i] = s_m1[c + i]; h_m2[i] = s_m2[c + i];
  }
  for (int p0 = g.pl; p0 < HW; p0 += 4 * PLn) {
    uint2 ru[4], ro[4], rd[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + u * PLn;
      if (p < HW) {
        ru[u] = nbs_ld4(ub + (int64_t)p * C);
        ro[u] = nbs_ld4(ob + (int64_t)p * out_pitch);
        rd[u] = nbs_ld4(db + (int64_t)p * dout_pitch);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int p = p0 + u * PLn;
      if (p < HW) {
        float uh[4], o[4], dd[4], r[4];
        nbs_unpack4(ru[u], uh); nbs_unpack4(ro[u], o); nbs_unpack4(rd[u], dd);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float du = dd[i] * (o[i] > 0.f ? 1.f : slope);
          r[i] = h_a[i] * du - h_m1[i] - uh[i] * h_m2[i];
        }
        nbs_st4(yb + (int64_t)p * dy_pitch, r);
      }
    }
  }
}

// BVAE_NB_SMALL (default 1): the nbs_* kernels on the sites they cover; BVAE_NB_SMALL_HW / BVAE_NB_SMALL_N are the largest
// map (pixels) and the smallest sample count that take them above 128 pixels (one CTA per sample needs enough samples to
// fill the machine; maps of <= 128 pixels always qualify: their alternative is one CTA per sample as well).  Any non-default
// BVAE_NB_MODE / BVAE_NB_MLP selects the nb_cl_* / nb_small_* / tiled kernels, which stay as the reference implementations.
static bool nbs_enabled(const bvae_nb_desc* d) {
  if (option("BVAE_NB_SMALL", 1) == 0 || option("BVAE_NB_MODE", 0) != 0 || option("BVAE_NB_MLP", 2) != 2) return false;
  const int HW = d->H * d->W, C = d->C;
  if (C < 32 || C > 1024 || (C & (C - 1)) != 0) return false;
  // static + dynamic shared memory of the largest kernel (nbs_bwd1: 12.2 KB static) must stay inside the default 48 KB
  if (nbs_fwd_smem(HW, C) > 44 * 1024 || nbs_bwd1_smem(HW, C) > 35 * 1024) return false;
  if (HW <= 128) return C >= 256;
  int max_hw = option("BVAE_NB_SMALL_HW", 512);
  if (max_hw > NBS_MAX_HW) max_hw = NBS_MAX_HW;
  return HW <= max_hw && d->N >= option("BVAE_NB_SMALL_N", 256);
}

// BVAE_NB_MLP: 0 keeps the channel MLP inside the per-sample kernels; 1 batches it in the backward pass of the >= 512-channel
// small maps; 2 (default since round 2) batches it everywhere (forward too).  Round 1 measured 1 fastest kernel by kernel
// (every extra stage pays the partial second wave of the per-sample CTAs again); on the final multi-stream step 2 wins end to
// end: 40.91 ms per 512-bar step against 41.3 (1) and 42.08 (0) -- the staged kernels are shorter and leave more of the SMs to
// the other branch's stream.
static bool nb_mlp_batched(const bvae_nb_desc* d, int CLn, bool backward) {
  const int v = option("BVAE_NB_MLP", 2);
  if (v == 0 || CLn != 1 || !d->has_cbam || d->C < 128) return false;
  return v == 2 || (backward && d->C >= 512);
}

static void nb_cl_geometry(const bvae_nb_desc* d, int* CLn, int* slice) {
  const int HW = d->H * d->W;
  const int cl = HW <= 128 ? 1 : (HW <= 1440 ? 4 : (HW <= 2880 ? 8 : 16));
  *CLn = cl;
  *slice = ceil_div(HW, cl);
}
static size_t nb_cl_fwd_smem(int C, int slice) { return (size_t)(8 * C) * 4 + (size_t)2 * C * 8 + (64 + 64 + 32) * 4 + (size_t)slice * (8 + 4 + 4) + 64; }
static size_t nb_cl_bwd_smem(int C, int slice) { return (size_t)(12 * C) * 4 + (3 * 64 + 32 + 32) * 4 + (size_t)slice * 5 * 4 + 64; }

template <typename K>
static int nb_cl_launch(K kernel, const bvae_nb_desc* d, size_t smem, cudaStream_t st, const char* what, int stage = 0) {
  int CLn, slice;
  nb_cl_geometry(d, &CLn, &slice);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(d->N * CLn));
  cfg.blockDim = dim3(CLT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CLn; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, *d, CLn, slice, stage);
  if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return BVAE_ERR_CUDA; }
  return check_launch(what);
}

// BVAE_NB_MODE: 0 (default) = cluster kernel for maps <= 128 pixels (one CTA per sample), tiled kernels otherwise;
//               1 = cluster kernels everywhere (4..16 CTAs per sample; halves HBM traffic but is latency bound:
//                   measured 1.2-1.5x slower than the tiled kernels on the 96x60 maps);  2 = tiled / nb_small only.
static int nb_mode() {
  return option("BVAE_NB_MODE", 0);
}
static bool use_nb_cluster(const bvae_nb_desc* d) {
  const int m = nb_mode();
  return m == 1 || (m == 0 && d->H * d->W <= 128);
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
// Tried and dropped (round 3): sizing the grids to fill their last wave (512 samples x 2 chunks = 1024 CTAs are 2.3 waves of
// 444 resident CTAs; 6 chunks per sample are 6.9) -- 41.0 instead of 39.4 ms per step: the extra CTAs re-load the per-channel
// constants and triple the per-CTA atomics, and the sweeps are not wave-quantised the way a compute-bound grid is.
static int pick_ppc(int HW, int N, int G) {
  // pixels per CTA: a multiple of the CTA's pixel-group count, sized so the grid has ~4 waves of 148 SMs
  const int pg = 8 * (32 / G);
  int chunks = ceil_div(148 * 4, N);
  if (chunks < 1) chunks = 1;
  int ppc = ceil_div(HW, chunks);
  ppc = ceil_div(ppc, pg) * pg;
  return ppc;
}

// BVAE_NB_SPLIT=0 keeps the two full reduction sweeps (nbf_bwd1 + nbf_bwd2) on the CBAM sites
static bool nb_split_enabled() {
  return option("BVAE_NB_SPLIT", 1) != 0;
}

// BVAE_NB_FAST=0 selects the generic tiled backward kernels (kept as the reference implementation of the fast ones)
static bool nb_fast_enabled() {
  return option("BVAE_NB_FAST", 1) != 0;
}

static int launch_bwd_w(const bvae_nb_desc* d, cudaStream_t st) {
  const int N = d->N, C = d->C;
  if (C >= 512 && C % 64 == 0 && N >= 32) {        // (few CTAs and idle hidden-unit slots below 512 channels)
    dim3 gw(C / 64, ceil_div(N, WB_NS));
    nb_bwd_w_tiled_kernel<<<gw, 256, 0, st>>>(N, C, d->Cr, d->nc, d->beta, d->bwd_nc, d->bwd_h, d->dw1, d->dw2);
    return check_launch("nb_bwd_w");
  }
  int nsplit = ceil_div(148 * 8, ceil_div(C * d->Cr, 256));
  if (nsplit > N) nsplit = N;
  if (nsplit > 32) nsplit = 32;
  if (nsplit < 1) nsplit = 1;
  const int npb = ceil_div(N, nsplit);
  dim3 gw(ceil_div(C * d->Cr, 256), ceil_div(N, npb));
  nb_bwd_w_kernel<<<gw, 256, 0, st>>>(N, C, d->Cr, d->nc, d->beta, d->bwd_nc, d->bwd_h, d->dw1, d->dw2, npb);
  return check_launch("nb_bwd_w");
}

static int validate(const bvae_nb_desc* d, const char* who) {
  BVAE_REQUIRE(d->C >= 32 && d->C <= 1024 && (d->C & (d->C - 1)) == 0, BVAE_ERR_SHAPE,
               "%s: C=%d must be a power of two in [32,1024]", who, d->C);
  BVAE_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0, BVAE_ERR_SHAPE, "%s: empty tensor", who);
  BVAE_REQUIRE(d->y_pitch % 8 == 0 && d->out_pitch % 8 == 0, BVAE_ERR_ALIGN, "%s: pitches must be multiples of 8", who);
  BVAE_REQUIRE(!d->has_cbam || (d->Cr == d->C / 16 && d->Cr <= 64), BVAE_ERR_SHAPE, "%s: Cr must be C/16", who);
  BVAE_REQUIRE(d->has_cbam || d->res_mode == 0, BVAE_ERR_SHAPE, "%s: residual modes need CBAM", who);
  return BVAE_OK;
}

#define DISPATCH_ITERS(iters, ...)              \
  switch (iters) {                              \
    case 1: { constexpr int IT = 1; __VA_ARGS__; } break; \
    case 2: { constexpr int IT = 2; __VA_ARGS__; } break; \
    default: { constexpr int IT = 4; __VA_ARGS__; } break; \
  }

}  // namespace bvae

using namespace bvae;

extern "C" int bvae_nb_forward(const bvae_nb_desc* d, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  int rc = validate(d, "nb_forward");
  if (rc) return rc;
  const int N = d->N, HW = d->H * d->W, C = d->C;
  const bool det = sync_det();
  BVAE_REQUIRE(!d->stats_fused || !(use_nb_cluster(d) || nb_small_ok(d)), BVAE_ERR_UNSUPPORTED,
               "nb_forward: fused statistics are only consumed by the tiled path (H*W > 128)");
  if (nbs_enabled(d) && d->y_f32 && !d->stats_fused) {
    // small maps: statistics + coefficients -> batched channel MLP -> one per-sample kernel for everything else
    dim3 gs(C >= 128 ? C / 128 : 1, N);
    if (d->has_cbam)
      nbs_stats_kernel<true><<<gs, NBS_T, 0, st>>>((const float*)d->y, d->y_pitch, HW, C, d->gamma, d->beta, d->eps, d->nc,
                                                   d->nc_idx);
    else
      nbs_stats_kernel<false><<<gs, NBS_T, 0, st>>>((const float*)d->y, d->y_pitch, HW, C, d->gamma, d->beta, d->eps, d->nc,
                                                    d->nc_idx);
    if ((rc = check_launch("nbs_stats"))) return rc;
    if (!d->has_cbam) {
      nbs_fwd_kernel<false, 0><<<N, NBS_T, 0, st>>>(*d);
      return check_launch("nbs_fwd");
    }
    BVAE_REQUIRE(d->res_mode >= 1 && d->res_mode <= 3, BVAE_ERR_SHAPE, "nb_forward: CBAM needs res_mode 1..3");
    nb_mlp_fwd_kernel<<<ceil_div(N, MLP_NS), 256, (size_t)(1 + MLP_NS) * C * sizeof(float), st>>>(N, C, d->Cr, d->w1, d->w2,
                                                                                                 d->beta, d->nc);
    if ((rc = check_launch("nb_mlp_fwd"))) return rc;
    const size_t smf = nbs_fwd_smem(HW, C);
    if (d->res_mode == 1) nbs_fwd_kernel<true, 1><<<N, NBS_T, smf, st>>>(*d);
    else if (d->res_mode == 2) nbs_fwd_kernel<true, 2><<<N, NBS_T, smf, st>>>(*d);
    else nbs_fwd_kernel<true, 3><<<N, NBS_T, smf, st>>>(*d);
    return check_launch("nbs_fwd");
  }
  if (use_nb_cluster(d)) {
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(nb_cl_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(nb_cl_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(nb_cl_fwd_kernel<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaFuncSetAttribute(nb_cl_fwd_kernel<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      attr = true;
    }
    int CLn, slice;
    nb_cl_geometry(d, &CLn, &slice);
    const size_t smem = nb_cl_fwd_smem(C, slice);
    if (nb_mlp_batched(d, CLn, false)) {
      // statistics -> batched channel MLP -> everything else (the sample's tensors stay in L2 between the stages)
      rc = d->y_f32 ? nb_cl_launch(nb_cl_fwd_kernel<true>, d, smem, st, "nb_cl_fwd", 1)
                    : nb_cl_launch(nb_cl_fwd_kernel<false>, d, smem, st, "nb_cl_fwd", 1);
      if (rc) return rc;
      nb_mlp_fwd_kernel<<<ceil_div(N, MLP_NS), 256, (size_t)(1 + MLP_NS) * C * sizeof(float), st>>>(N, C, d->Cr, d->w1, d->w2,
                                                                                                   d->beta, d->nc);
      if ((rc = check_launch("nb_mlp_fwd"))) return rc;
      return d->y_f32 ? nb_cl_launch(nb_cl_fwd_kernel<true>, d, smem, st, "nb_cl_fwd", 2)
                      : nb_cl_launch(nb_cl_fwd_kernel<false>, d, smem, st, "nb_cl_fwd", 2);
    }
    if (d->y_f32) return nb_cl_launch(nb_cl_fwd_kernel<true>, d, smem, st, "nb_cl_fwd");
    return nb_cl_launch(nb_cl_fwd_kernel<false>, d, smem, st, "nb_cl_fwd");
  }
  if (nb_small_ok(d)) {
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(nb_small_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nb_small_fwd_smem(1024));
      cudaFuncSetAttribute(nb_small_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nb_small_fwd_smem(1024));
      attr = true;
    }
    if (d->y_f32) nb_small_fwd_kernel<true><<<N, 256, nb_small_fwd_smem(C), st>>>(*d);
    else nb_small_fwd_kernel<false><<<N, 256, nb_small_fwd_smem(C), st>>>(*d);
    return check_launch("nb_small_fwd");
  }
  const int64_t NC = (int64_t)N * C;
  // scratch layout inside d->stats: [NC] float2 | [NC] u64 max keys | [NC] u64 min keys   (own statistics pass), or
  //                                 [NC] double sum | [NC] double sumsq | [NC] u32 max | [NC] u32 min  (stats_fused)
  float2* ss = (float2*)d->stats;
  u64* kmax = (u64*)(ss + NC);
  u64* kmin = kmax + NC;
  if (!d->stats_fused) {
    const int cchunks = ceil_div(C, 256);
    // enough CTAs for ~8 per SM; every CTA should still stream >= 256 pixels
    int psplit = ceil_div(148 * 8, N * cchunks);
    const int maxsplit = HW / 256 > 0 ? HW / 256 : 1;
    if (psplit > maxsplit) psplit = maxsplit;
    if (psplit < 1 || det) psplit = 1;           // deterministic mode: no float atomics across CTAs
    if (psplit > 1) {
      if (cudaMemsetAsync(d->stats, 0, NC * 24, st) != cudaSuccess) { set_error("nb_forward: memset failed"); return BVAE_ERR_CUDA; }
    }
    dim3 g1(cchunks, psplit, N);
    if (d->has_cbam) {
      if (d->y_f32) nb_stats_kernel<true, true><<<g1, 256, 0, st>>>(d->y, d->y_pitch, HW, C, psplit, ss, kmax, kmin);
      else nb_stats_kernel<false, true><<<g1, 256, 0, st>>>(d->y, d->y_pitch, HW, C, psplit, ss, kmax, kmin);
    } else {
      if (d->y_f32) nb_stats_kernel<true, false><<<g1, 256, 0, st>>>(d->y, d->y_pitch, HW, C, psplit, ss, kmax, kmin);
      else nb_stats_kernel<false, false><<<g1, 256, 0, st>>>(d->y, d->y_pitch, HW, C, psplit, ss, kmax, kmin);
    }
    if ((rc = check_launch("nb_stats"))) return rc;
  }

  const size_t sm2 = (2 * C + 128) * sizeof(float);
  if (d->y_f32)
    nb_coef_kernel<true><<<N, 256, sm2, st>>>(d->y, d->y_pitch, HW, C, ss, kmax, kmin, d->gamma, d->beta, d->eps,
                                              d->has_cbam, d->Cr, d->w1, d->w2, d->nc, d->nc_idx, d->stats_fused, N);
  else
    nb_coef_kernel<false><<<N, 256, sm2, st>>>(d->y, d->y_pitch, HW, C, ss, kmax, kmin, d->gamma, d->beta, d->eps,
                                               d->has_cbam, d->Cr, d->w1, d->w2, d->nc, d->nc_idx, d->stats_fused, N);
  if ((rc = check_launch("nb_coef"))) return rc;

  const int G = C / 8 < 32 ? C / 8 : 32;
  const int iters = C / (8 * G);
  const int ppc = pick_ppc(HW, N, G);
  dim3 g3(ceil_div(HW, ppc), N);
  const size_t sm3 = 6 * C * sizeof(float);
  const bool fast = iters == 1 && !d->stats_fused && nb_fast_enabled();
#define NBF_POOL(F) do {                                                                                                  \
    if (d->has_cbam) nbf_pool_kernel<F, true><<<g3, 256, 0, st>>>(d->y, d->y_pitch, HW, C, d->nc, d->slope, (bf16*)d->uhat, \
                                                                  (bf16*)d->out, d->out_pitch, d->sa, d->cidx, ppc);      \
    else nbf_pool_kernel<F, false><<<g3, 256, 0, st>>>(d->y, d->y_pitch, HW, C, d->nc, d->slope, (bf16*)d->uhat,           \
                                                       (bf16*)d->out, d->out_pitch, d->sa, d->cidx, ppc);                 \
  } while (0)
  if (fast) {
    if (d->y_f32) NBF_POOL(true); else NBF_POOL(false);
  } else {
    DISPATCH_ITERS(iters, {
      if (d->y_f32)
        nb_pool_kernel<true, IT><<<g3, 256, sm3, st>>>(d->y, d->y_pitch, HW, C, d->nc, d->has_cbam, d->slope,
                                                        (bf16*)d->uhat, (bf16*)d->out, d->out_pitch, d->sa, d->cidx, ppc,
                                                        d->stats_fused, d->nc_idx);
      else
        nb_pool_kernel<false, IT><<<g3, 256, sm3, st>>>(d->y, d->y_pitch, HW, C, d->nc, d->has_cbam, d->slope,
                                                         (bf16*)d->uhat, (bf16*)d->out, d->out_pitch, d->sa, d->cidx, ppc,
                                                         d->stats_fused, d->nc_idx);
    });
  }
#undef NBF_POOL
  if ((rc = check_launch("nb_pool"))) return rc;
  if (!d->has_cbam) return BVAE_OK;

  const size_t sm4 = 3 * C * sizeof(float);
  dim3 gg(ceil_div(HW, 256), N);
  nb_gate_kernel<<<gg, 256, 0, st>>>(d->H, d->W, d->wsp, d->sa, d->gs);
  if ((rc = check_launch("nb_gate"))) return rc;
  BVAE_REQUIRE(d->res_mode >= 1 && d->res_mode <= 3, BVAE_ERR_SHAPE, "nb_forward: CBAM needs res_mode 1..3");
#define NBF_APPLY(F, M) nbf_apply_kernel<F, M><<<g3, 256, 0, st>>>(d->y, d->y_pitch, HW, C, d->nc, (const bf16*)d->res, \
                                                                   d->res_pitch, d->slope, (bf16*)d->out, d->out_pitch, d->gs, ppc)
  if (fast) {
    if (d->y_f32) {
      if (d->res_mode == 1) NBF_APPLY(true, 1); else if (d->res_mode == 2) NBF_APPLY(true, 2); else NBF_APPLY(true, 3);
    } else {
      if (d->res_mode == 1) NBF_APPLY(false, 1); else if (d->res_mode == 2) NBF_APPLY(false, 2); else NBF_APPLY(false, 3);
    }
    return check_launch("nb_apply");
  }
#undef NBF_APPLY
  DISPATCH_ITERS(iters, {
    if (d->y_f32)
      nb_apply_kernel<true, IT><<<g3, 256, sm4, st>>>(d->y, d->y_pitch, d->H, d->W, C, d->nc, d->res_mode,
                                                       (const bf16*)d->res, d->res_pitch, d->slope, (bf16*)d->out,
                                                       d->out_pitch, d->gs, ppc);
    else
      nb_apply_kernel<false, IT><<<g3, 256, sm4, st>>>(d->y, d->y_pitch, d->H, d->W, C, d->nc, d->res_mode,
                                                        (const bf16*)d->res, d->res_pitch, d->slope, (bf16*)d->out,
                                                        d->out_pitch, d->gs, ppc);
  });
  return check_launch("nb_apply");
}

extern "C" int bvae_nb_backward(const bvae_nb_desc* d, void* stream_) {
  cudaStream_t st = (cudaStream_t)stream_;
  int rc = validate(d, "nb_backward");
  if (rc) return rc;
  BVAE_REQUIRE(d->dout_pitch % 8 == 0 && d->dy_pitch % 8 == 0, BVAE_ERR_ALIGN, "nb_backward: pitches % 8 != 0");
  BVAE_REQUIRE(d->uhat != nullptr, BVAE_ERR_SHAPE, "nb_backward: uhat is NULL (the forward pass ran in inference mode)");
  const int N = d->N, HW = d->H * d->W, C = d->C;
  if (nbs_enabled(d)) {
    if (!d->has_cbam) {
      nbs_bwd_plain_kernel<<<N, NBS_T, 0, st>>>(*d);
      return check_launch("nbs_bwd_plain");
    }
    BVAE_REQUIRE(d->res_mode >= 1 && d->res_mode <= 3, BVAE_ERR_SHAPE, "nb_backward: CBAM needs res_mode 1..3");
    static bool attr_s = false;
    if (!attr_s) {
      cudaFuncSetAttribute(nb_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 + 2 * MLP_NS) * 1024 * 4);
      attr_s = true;
    }
    // reduction sweeps -> batched channel-MLP backward + InstanceNorm coefficients -> dy sweep -> MLP weight gradients
    const size_t sm1 = nbs_bwd1_smem(HW, C), sm2 = (size_t)HW * 16;
    if (d->res_mode == 1) nbs_bwd1_kernel<1><<<N, NBS_T, sm1, st>>>(*d);
    else if (d->res_mode == 2) nbs_bwd1_kernel<2><<<N, NBS_T, sm1, st>>>(*d);
    else nbs_bwd1_kernel<3><<<N, NBS_T, sm1, st>>>(*d);
    if ((rc = check_launch("nbs_bwd1"))) return rc;
    nb_mlp_bwd_kernel<<<ceil_div(N, MLP_NS), 256, (size_t)(1 + 2 * MLP_NS) * C * sizeof(float), st>>>(
        N, HW, C, d->Cr, d->w1, d->w2, d->beta, d->nc, d->bwd_nc, d->bwd_h, d->dgamma, d->dbeta);
    if ((rc = check_launch("nb_mlp_bwd"))) return rc;
    if (d->res_mode == 1) nbs_bwd2_kernel<1><<<N, NBS_T, sm2, st>>>(*d);
    else if (d->res_mode == 2) nbs_bwd2_kernel<2><<<N, NBS_T, sm2, st>>>(*d);
    else nbs_bwd2_kernel<3><<<N, NBS_T, sm2, st>>>(*d);
    if ((rc = check_launch("nbs_bwd2"))) return rc;
    return launch_bwd_w(d, st);
  }
  if (use_nb_cluster(d)) {
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(nb_cl_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      cudaFuncSetAttribute(nb_cl_bwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      attr = true;
    }
    int CLn, slice;
    nb_cl_geometry(d, &CLn, &slice);
    if (nb_mlp_batched(d, CLn, true)) {
      static bool attr2 = false;
      if (!attr2) {
        cudaFuncSetAttribute(nb_mlp_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (1 + 2 * MLP_NS) * 1024 * 4);
        attr2 = true;
      }
      if ((rc = nb_cl_launch(nb_cl_bwd_kernel, d, nb_cl_bwd_smem(C, slice), st, "nb_cl_bwd", 1))) return rc;
      nb_mlp_bwd_kernel<<<ceil_div(N, MLP_NS), 256, (size_t)(1 + 2 * MLP_NS) * C * sizeof(float), st>>>(
          N, HW, C, d->Cr, d->w1, d->w2, d->beta, d->nc, d->bwd_nc, d->bwd_h, d->dgamma, d->dbeta);
      if ((rc = check_launch("nb_mlp_bwd"))) return rc;
      if ((rc = nb_cl_launch(nb_cl_bwd_kernel, d, nb_cl_bwd_smem(C, slice), st, "nb_cl_bwd", 2))) return rc;
    } else if ((rc = nb_cl_launch(nb_cl_bwd_kernel, d, nb_cl_bwd_smem(C, slice), st, "nb_cl_bwd"))) {
      return rc;
    }
    if (d->has_cbam) {
      if ((rc = launch_bwd_w(d, st))) return rc;
    }
    return BVAE_OK;
  }
  if (nb_small_ok(d)) {
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(nb_small_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)nb_small_bwd_smem(1024));
      attr = true;
    }
    nb_small_bwd_kernel<<<N, 256, nb_small_bwd_smem(C), st>>>(*d);
    if ((rc = check_launch("nb_small_bwd"))) return rc;
    if (d->has_cbam) {
      if ((rc = launch_bwd_w(d, st))) return rc;
    }
    return BVAE_OK;
  }
  const int G = C / 8 < 32 ? C / 8 : 32;
  const int iters = C / (8 * G);
  const int ppc = pick_ppc(HW, N, G);
  dim3 gp(ceil_div(HW, ppc), N);
  if (cudaMemsetAsync(d->bwd_nc, 0, (size_t)N * C * BN_W * sizeof(float), st) != cudaSuccess) {
    set_error("nb_backward: memset failed");
    return BVAE_ERR_CUDA;
  }
  const size_t smb = 4 * C * sizeof(float);
  const bool fast = iters == 1 && nb_fast_enabled();        // C <= 256: the restructured kernels
  const int mode = d->has_cbam ? d->res_mode : 0;
  BVAE_REQUIRE(!d->has_cbam || (mode >= 1 && mode <= 3), BVAE_ERR_SHAPE, "nb_backward: CBAM needs res_mode 1..3");
#define NBF_MODES(...)                                          \
  switch (mode) {                                               \
    case 0: { constexpr int MD = 0; __VA_ARGS__; } break;       \
    case 1: { constexpr int MD = 1; __VA_ARGS__; } break;       \
    case 2: { constexpr int MD = 2; __VA_ARGS__; } break;       \
    default: { constexpr int MD = 3; __VA_ARGS__; } break;      \
  }
  const bool split = fast && d->has_cbam && nb_split_enabled();     // ds-sums in the dq sweep, dsp-sums from uhat alone
  if (d->has_cbam) {
    if (split) {
      NBF_MODES({
        nbf_bwd1x_kernel<(MD == 0 ? 1 : MD)><<<gp, 256, C * sizeof(float), st>>>(
            (const bf16*)d->dout, d->dout_pitch, (const bf16*)d->out, d->out_pitch, (const bf16*)d->uhat, HW, C, d->nc,
            d->gamma, d->beta, d->gs, d->slope, (bf16*)d->dres, d->dres_pitch, d->bwd_nc, d->bwd_px, ppc);
      });
    } else if (fast) {
      nbf_bwd1_kernel<<<gp, 256, C * sizeof(float), st>>>((const bf16*)d->dout, d->dout_pitch, (const bf16*)d->out,
                                                          d->out_pitch, (const bf16*)d->uhat, HW, C, d->nc, d->gamma,
                                                          d->beta, d->gs, d->slope, d->bwd_nc, d->bwd_px, ppc);
    } else {
      DISPATCH_ITERS(iters, {
        nb_bwd1_kernel<IT><<<gp, 256, smb, st>>>((const bf16*)d->dout, d->dout_pitch, (const bf16*)d->out, d->out_pitch,
                                                  (const bf16*)d->uhat, HW, C, d->nc, d->gamma, d->beta, d->gs, d->slope,
                                                  d->bwd_nc, d->bwd_px, ppc);
      });
    }
    if ((rc = check_launch("nb_bwd1"))) return rc;
    // one block reduction of dwsp per CTA, then 18 same-address global atomics: with several CTAs per sample those
    // atomics (1536 per address at 512 bars) serialised at L2 and were most of this kernel's time
    dim3 gs(N >= 296 ? 1 : ceil_div(HW, 256 * 8), N);
    nb_bwd_sp_kernel<<<gs, 256, 0, st>>>(d->H, d->W, d->wsp, d->sa, d->bwd_px, d->dwsp);
    if ((rc = check_launch("nb_bwd_sp"))) return rc;
  }
  if (split) {
    nbf_bwd2x_kernel<<<gp, 256, C * sizeof(float), st>>>((const bf16*)d->uhat, HW, C, d->nc, d->gamma, d->beta, d->cidx,
                                                         d->bwd_px, d->bwd_nc, ppc);
  } else if (fast) {
    NBF_MODES({
      nbf_bwd2_kernel<MD><<<gp, 256, C * sizeof(float), st>>>((const bf16*)d->dout, d->dout_pitch, (const bf16*)d->out,
                                                               d->out_pitch, (const bf16*)d->uhat, HW, C, d->nc, d->gamma,
                                                               d->beta, d->gs, d->cidx, d->bwd_px, d->slope, (bf16*)d->dres,
                                                               d->dres_pitch, d->bwd_nc, ppc);
    });
  } else {
    DISPATCH_ITERS(iters, {
      nb_bwd2_kernel<IT><<<gp, 256, smb, st>>>((const bf16*)d->dout, d->dout_pitch, (const bf16*)d->out, d->out_pitch,
                                                (const bf16*)d->uhat, HW, C, d->nc, d->gamma, d->beta, d->gs, d->cidx,
                                                d->bwd_px, d->has_cbam, d->res_mode, d->slope, (bf16*)d->dy, d->dy_pitch,
                                                (bf16*)d->dres, d->dres_pitch, d->bwd_nc, ppc);
    });
  }
  if ((rc = check_launch("nb_bwd2"))) return rc;
  const size_t smc = (3 * C + 192) * sizeof(float);
  nb_bwd_coef_kernel<<<N, 256, smc, st>>>(HW, C, d->has_cbam, d->Cr, d->nc, d->beta, d->w1, d->w2, d->bwd_nc, d->dgamma,
                                          d->dbeta, d->bwd_h);
  if ((rc = check_launch("nb_bwd_coef"))) return rc;
  if (d->has_cbam) {
    if ((rc = launch_bwd_w(d, st))) return rc;
  }
  const size_t sm5 = 6 * C * sizeof(float);
  if (fast) {
    NBF_MODES({
      nbf_bwd3_kernel<MD><<<gp, 256, 2 * C * sizeof(float), st>>>((const bf16*)d->dout, d->dout_pitch, (const bf16*)d->out,
                                                                  d->out_pitch, (const bf16*)d->uhat, HW, C, d->nc, d->nc_idx,
                                                                  d->gs, d->cidx, d->bwd_px, d->bwd_nc, d->slope,
                                                                  (bf16*)d->dy, d->dy_pitch, ppc);
    });
    return check_launch("nb_bwd3");
  }
#undef NBF_MODES
  DISPATCH_ITERS(iters, {
    nb_bwd3_kernel<IT><<<gp, 256, sm5, st>>>((const bf16*)d->dout, d->dout_pitch, (const bf16*)d->out, d->out_pitch,
                                              (const bf16*)d->uhat, HW, C, d->nc, d->nc_idx, d->gamma, d->beta, d->gs,
                                              d->cidx, d->bwd_px, d->bwd_nc, d->has_cbam, d->res_mode, d->slope,
                                              (bf16*)d->dy, d->dy_pitch, ppc);
  });
  return check_launch("nb_bwd3");
}

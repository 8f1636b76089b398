// bits.cu -- bit-packed piano-rolls (1 bit per cell, MSB first, i.e. numpy.packbits order).
//
// A bar is a binary [96,60] grid; the reference ships it as fp32 (data/bar_dataset.py:22-25, agent/barGen.py:134-141:
// 23 KB per bar, 92 KB per phrase).  Packed, a training sample (note + pre_note + pre_phrase) is 4320 bytes instead
// of 138 KB, so the host -> device stream of a 512-bar step is 2.2 MB instead of 70.8 MB.  Both kernels are pure
// HBM streaming: one byte of bits <-> 8 cells.
//   unpack: 1 B read, 16 B (bf16) [+ 32 B (fp32)] written per byte of bits
//   threshold_pack: 32 B read, 1 B [+ 32 B] written per byte of bits
#include "common.cuh"

namespace bvae {

// One thread per packed byte.  Consecutive threads write consecutive 16-byte bf16 groups (fully coalesced); the
// optional fp32 copy (the BCE target) is two float4 per thread.
__global__ void __launch_bounds__(256) unpack_bits_kernel(const uint8_t* __restrict__ bits, int64_t nbytes_full,
                                                          int64_t nbits, bf16* __restrict__ out_bf16,
                                                          float* __restrict__ out_f32, int64_t nbytes_f32) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbytes_full; i += stride) {
    const uint32_t b = bits[i];
    if (out_bf16 != nullptr) {
      uint32_t w[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {       // bf16 1.0 == 0x3F80; cell 2k is the low half of word k
        const uint32_t lo = (b >> (7 - 2 * k)) & 1u, hi = (b >> (6 - 2 * k)) & 1u;
        w[k] = lo * 0x3F80u | hi * 0x3F800000u;
      }
      *reinterpret_cast<uint4*>(out_bf16 + i * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    if (out_f32 != nullptr && i < nbytes_f32) {
      float4 a, c;
      a.x = (float)((b >> 7) & 1u); a.y = (float)((b >> 6) & 1u);
      a.z = (float)((b >> 5) & 1u); a.w = (float)((b >> 4) & 1u);
      c.x = (float)((b >> 3) & 1u); c.y = (float)((b >> 2) & 1u);
      c.z = (float)((b >> 1) & 1u); c.w = (float)(b & 1u);
      *reinterpret_cast<float4*>(out_f32 + i * 8) = a;
      *reinterpret_cast<float4*>(out_f32 + i * 8 + 4) = c;
    }
  }
  // ragged tail: the last byte carries nbits % 8 valid cells
  const int tail = (int)(nbits - nbytes_full * 8);
  if (tail > 0 && blockIdx.x == 0 && threadIdx.x < tail) {
    const uint32_t b = bits[nbytes_full];
    const uint32_t v = (b >> (7 - threadIdx.x)) & 1u;
    const int64_t j = nbytes_full * 8 + threadIdx.x;
    if (out_bf16 != nullptr) out_bf16[j] = f2bf((float)v);
    if (out_f32 != nullptr && nbytes_f32 > nbytes_full) out_f32[j] = (float)v;
  }
}

// bits[i] = pack of (p[8i + j] > thr), optionally also the {0,1} floats (the next step's pre_bar, maker_bar.py:39).
__global__ void __launch_bounds__(256) threshold_pack_kernel(const float* __restrict__ p, int64_t nbytes_full,
                                                             int64_t n, float thr, uint8_t* __restrict__ bits,
                                                             float* __restrict__ out_f32) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nbytes_full; i += stride) {
    float f[8];
    ldg8f(p + i * 8, f);
    uint32_t b = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) b |= (f[j] > thr ? 1u : 0u) << (7 - j);
    if (bits != nullptr) bits[i] = (uint8_t)b;
    if (out_f32 != nullptr) {
      float4 a, c;
      a.x = f[0] > thr ? 1.f : 0.f; a.y = f[1] > thr ? 1.f : 0.f;
      a.z = f[2] > thr ? 1.f : 0.f; a.w = f[3] > thr ? 1.f : 0.f;
      c.x = f[4] > thr ? 1.f : 0.f; c.y = f[5] > thr ? 1.f : 0.f;
      c.z = f[6] > thr ? 1.f : 0.f; c.w = f[7] > thr ? 1.f : 0.f;
      *reinterpret_cast<float4*>(out_f32 + i * 8) = a;
      *reinterpret_cast<float4*>(out_f32 + i * 8 + 4) = c;
    }
  }
  const int tail = (int)(n - nbytes_full * 8);
  if (tail > 0 && blockIdx.x == 0 && threadIdx.x == 0) {
    uint32_t b = 0;
    for (int j = 0; j < tail; ++j) {
      const float v = p[nbytes_full * 8 + j];
      b |= (v > thr ? 1u : 0u) << (7 - j);
      if (out_f32 != nullptr) out_f32[nbytes_full * 8 + j] = v > thr ? 1.f : 0.f;
    }
    if (bits != nullptr) bits[nbytes_full] = (uint8_t)b;
  }
}

static int grid_bytes(int64_t nbytes) {
  int64_t g = ceil_div64(nbytes > 0 ? nbytes : 1, 256);
  const int64_t cap = 148 * 8;       // 8 resident 256-thread CTAs per SM
  return (int)(g < cap ? g : cap);
}

}  // namespace bvae

using namespace bvae;

extern "C" {

int bvae_unpack_bits(const void* bits, int64_t nbits, void* out_bf16, float* out_f32, int64_t nbits_f32,
                     void* stream) {
  BVAE_REQUIRE(nbits >= 0 && nbits_f32 >= 0 && nbits_f32 <= nbits, BVAE_ERR_SHAPE,
               "unpack_bits: need 0 <= nbits_f32 <= nbits (got %lld, %lld)", (long long)nbits_f32, (long long)nbits);
  BVAE_REQUIRE(nbits_f32 == nbits || nbits_f32 % 8 == 0, BVAE_ERR_SHAPE,
               "unpack_bits: nbits_f32 must be a multiple of 8 or equal nbits");
  BVAE_REQUIRE(out_bf16 != nullptr || out_f32 != nullptr, BVAE_ERR_SHAPE, "unpack_bits: no output given");
  BVAE_REQUIRE((((uintptr_t)out_bf16 | (uintptr_t)out_f32) & 15) == 0, BVAE_ERR_ALIGN,
               "unpack_bits: outputs must be 16-byte aligned");
  if (nbits == 0) return BVAE_OK;
  const int64_t nbytes_full = nbits / 8;
  const int64_t nbytes_f32 = out_f32 == nullptr ? 0 : (nbits_f32 == nbits ? nbytes_full + 1 : nbits_f32 / 8);
  unpack_bits_kernel<<<grid_bytes(nbytes_full), 256, 0, (cudaStream_t)stream>>>(
      (const uint8_t*)bits, nbytes_full, nbits, (bf16*)out_bf16, out_f32, nbytes_f32);
  return check_launch("unpack_bits");
}

int bvae_threshold_pack(const float* p, int64_t n, float threshold, void* bits, float* out_f32, void* stream) {
  BVAE_REQUIRE(n >= 0, BVAE_ERR_SHAPE, "threshold_pack: n must be >= 0");
  BVAE_REQUIRE(bits != nullptr || out_f32 != nullptr, BVAE_ERR_SHAPE, "threshold_pack: no output given");
  BVAE_REQUIRE((((uintptr_t)p | (uintptr_t)out_f32) & 15) == 0, BVAE_ERR_ALIGN,
               "threshold_pack: float buffers must be 16-byte aligned");
  if (n == 0) return BVAE_OK;
  const int64_t nbytes_full = n / 8;
  threshold_pack_kernel<<<grid_bytes(nbytes_full), 256, 0, (cudaStream_t)stream>>>(p, nbytes_full, n, threshold,
                                                                                  (uint8_t*)bits, out_f32);
  return check_launch("threshold_pack");
}

}  // extern "C"

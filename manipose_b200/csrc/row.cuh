// A token row of the activation held by one warp (shared by the forward token kernels in backbone.cu and the backward
// kernels in train.cu).
#pragma once

#include "common.cuh"

namespace mp {

constexpr int kTokWarps = 8;  // warps per CTA in the token kernels

// ---- a token row held by a warp: C/32 consecutive-in-groups-of-8 channels per lane -----------------------------
// C = 512: lane owns channels [8*lane, 8*lane+8) and [256 + 8*lane, 256 + 8*lane + 8)   (two 16-byte accesses)
// C = 128: lane owns channels [4*lane, 4*lane+4)                                          (one 8-byte access)
template <int C>
struct Row {
  static constexpr int kPer = C / 32;
  static_assert(C == 512 || C == 128, "C must be 128 or 512");
  __device__ static __forceinline__ int chan(int lane, int i) {
    if (C == 512) return (i < 8 ? 0 : 256) + lane * 8 + (i & 7);
    return lane * 4 + i;
  }
  // fp32 activation row (the residual stream) in the per-lane channel order
  __device__ static __forceinline__ void load_x(const float* __restrict__ row, int lane, float (&v)[kPer]) {
    if (C == 512) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 a = *reinterpret_cast<const float4*>(row + h * 256 + lane * 8);
        const float4 b = *reinterpret_cast<const float4*>(row + h * 256 + lane * 8 + 4);
        v[h * 8 + 0] = a.x; v[h * 8 + 1] = a.y; v[h * 8 + 2] = a.z; v[h * 8 + 3] = a.w;
        v[h * 8 + 4] = b.x; v[h * 8 + 5] = b.y; v[h * 8 + 6] = b.z; v[h * 8 + 7] = b.w;
      }
    } else {
      const float4 a = *reinterpret_cast<const float4*>(row + lane * 4);
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
  }
  __device__ static __forceinline__ void store_x(float* __restrict__ row, int lane, const float (&v)[kPer]) {
    if (C == 512) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        *reinterpret_cast<float4*>(row + h * 256 + lane * 8) = make_float4(v[h * 8 + 0], v[h * 8 + 1], v[h * 8 + 2], v[h * 8 + 3]);
        *reinterpret_cast<float4*>(row + h * 256 + lane * 8 + 4) = make_float4(v[h * 8 + 4], v[h * 8 + 5], v[h * 8 + 6], v[h * 8 + 7]);
      }
    } else {
      *reinterpret_cast<float4*>(row + lane * 4) = make_float4(v[0], v[1], v[2], v[3]);
    }
  }
  // 16-bit normalised activation row (GEMM operand)
  template <typename D>
  __device__ static __forceinline__ void store_h(uint16_t* __restrict__ row, int lane, const float (&v)[kPer]) {
    if (C == 512) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint4 u;
        u.x = D::pack2(v[h * 8 + 0], v[h * 8 + 1]);
        u.y = D::pack2(v[h * 8 + 2], v[h * 8 + 3]);
        u.z = D::pack2(v[h * 8 + 4], v[h * 8 + 5]);
        u.w = D::pack2(v[h * 8 + 6], v[h * 8 + 7]);
        *reinterpret_cast<uint4*>(row + h * 256 + lane * 8) = u;
      }
    } else {
      uint2 u;
      u.x = D::pack2(v[0], v[1]);
      u.y = D::pack2(v[2], v[3]);
      *reinterpret_cast<uint2*>(row + lane * 4) = u;
    }
  }
  // 16-bit row (activation gradient) -> fp32, same channel order
  template <typename D>
  __device__ static __forceinline__ void load_h(const uint16_t* __restrict__ row, int lane, float (&v)[kPer]) {
    if (C == 512) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint4 u = *reinterpret_cast<const uint4*>(row + h * 256 + lane * 8);
        const float2 a = D::unpack2(u.x), b = D::unpack2(u.y), c = D::unpack2(u.z), d = D::unpack2(u.w);
        v[h * 8 + 0] = a.x; v[h * 8 + 1] = a.y; v[h * 8 + 2] = b.x; v[h * 8 + 3] = b.y;
        v[h * 8 + 4] = c.x; v[h * 8 + 5] = c.y; v[h * 8 + 6] = d.x; v[h * 8 + 7] = d.y;
      }
    } else {
      const uint2 u = *reinterpret_cast<const uint2*>(row + lane * 4);
      const float2 a = D::unpack2(u.x), b = D::unpack2(u.y);
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
  }
  // fp32 parameter vector (gamma, beta, pos-embed row ...) in the same per-lane channel order
  __device__ static __forceinline__ void load_f32(const float* __restrict__ p, int lane, float (&v)[kPer]) {
    if (C == 512) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p + h * 256 + lane * 8));
        const float4 b = __ldg(reinterpret_cast<const float4*>(p + h * 256 + lane * 8 + 4));
        v[h * 8 + 0] = a.x; v[h * 8 + 1] = a.y; v[h * 8 + 2] = a.z; v[h * 8 + 3] = a.w;
        v[h * 8 + 4] = b.x; v[h * 8 + 5] = b.y; v[h * 8 + 6] = b.z; v[h * 8 + 7] = b.w;
      }
    } else {
      const float4 a = __ldg(reinterpret_cast<const float4*>(p + lane * 4));
      v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
  }
  // mean and 1/sqrt(var + eps) over the row (two-pass, biased variance: nn.LayerNorm)
  __device__ static __forceinline__ void stats(const float (&v)[kPer], float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < kPer; ++i) s += v[i];
    mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
      const float d = v[i] - mean;
      q = fmaf(d, d, q);
    }
    rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
  }
  __device__ static __forceinline__ void normalize(float (&v)[kPer], float mean, float rstd, const float (&g)[kPer], const float (&b)[kPer]) {
#pragma unroll
    for (int i = 0; i < kPer; ++i) v[i] = fmaf((v[i] - mean) * rstd, g[i], b[i]);
  }
};

// persistent-ish grid for one-warp-per-token kernels
inline int token_grid(int64_t n_tokens) {
  int64_t ctas = (n_tokens + kTokWarps - 1) / kTokWarps;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(ctas < cap ? (ctas > 0 ? ctas : 1) : cap);
}

}  // namespace mp

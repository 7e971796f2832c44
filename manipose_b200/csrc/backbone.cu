// Memory-bound pieces of the MixSTE backbone: joint / segment embeddings, the LayerNorm family, the K hypothesis
// heads and the bone-length head.  fp32 residual stream, 16-bit (bf16 | fp16) normalised activations, fp32 statistics and
// parameters, one warp per token.
//
// Replaces (reference, paths under hpe/mh_so3_hpe/architectures/):
//   mix_ste.py:128-138      STE_forward: Spatial_patch_to_embedding + Spatial_pos_embed
//   manifold_mix_ste.py:139-148 BonesMixSTE: joints_to_segments_proj (+ pos embed)
//   mix_ste.py:49,143,149,154,166,170,353,356  norm1 / norm2 / Spatial_norm / Temporal_norm / Temporal_pos_embed
//   rmcl_manifold_mix_ste.py:251-298  K x MCLHead (LN eps 1e-5, Linear 512->7, score Linear 17->1)
//   mix_ste.py:123-126,187 + manifold_mix_ste.py:150-154  segment head + mean over time
#include "common.cuh"
#include "ptx.cuh"
#include "row.cuh"

namespace mp {
int linear_f32_visible(const void* A, const void* W, const float* bias, float* Y, int64_t M, int64_t N, int64_t K, int64_t visible_cols,
                       int64_t pitch_cols, int dtype, cudaStream_t s);   // gemm.cu
}
namespace mp {
namespace {


// -------------------------------------------------------------------------------------------------- LayerNorm family
template <int C, typename D>
__global__ void __launch_bounds__(kTokWarps * 32)
layernorm_kernel(const float* __restrict__ x_in, float* __restrict__ x_out, uint16_t* __restrict__ h_out,
                 const float* __restrict__ post_g, const float* __restrict__ post_b, float post_eps, const float* __restrict__ pos,
                 int64_t pos_div, int64_t pos_mod, const float* __restrict__ ln_g, const float* __restrict__ ln_b, float ln_eps,
                 int64_t n_tokens) {
  pdl_launch_dependents();
  pdl_wait();
  using R = Row<C>;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * kTokWarps + (threadIdx.x >> 5);
  const int64_t stride = (int64_t)gridDim.x * kTokWarps;
  float pg[R::kPer], pb[R::kPer], lg[R::kPer], lb[R::kPer];
  if (post_g) {
    R::load_f32(post_g, lane, pg);
    R::load_f32(post_b, lane, pb);
  }
  if (ln_g) {
    R::load_f32(ln_g, lane, lg);
    R::load_f32(ln_b, lane, lb);
  }
  for (int64_t tok = warp_global; tok < n_tokens; tok += stride) {
    float v[R::kPer];
    R::load_x(x_in + tok * C, lane, v);
    float mean, rstd;
    if (post_g) {
      R::stats(v, post_eps, mean, rstd);
      R::normalize(v, mean, rstd, pg, pb);
      if (pos) {
        float pe[R::kPer];
        R::load_f32(pos + ((tok / pos_div) % pos_mod) * C, lane, pe);
#pragma unroll
        for (int i = 0; i < R::kPer; ++i) v[i] += pe[i];
      }
      R::store_x(x_out + tok * C, lane, v);
    }
    if (ln_g) {
      R::stats(v, ln_eps, mean, rstd);
      R::normalize(v, mean, rstd, lg, lb);
      R::template store_h<D>(h_out + tok * C, lane, v);
    }
  }
}

// C = 128 (the bone-length backbone): FOUR tokens per warp.  Lane = 8 * sub + l8 owns the 16-byte chunks i * 8 + l8 (i = 0 .. 3) of
// token `sub` of the warp's group, so a load instruction covers 4 x 128 contiguous bytes, the row statistics are 3-step butterflies
// over 8 lanes (5 steps over 32 with one token per warp) and a warp has 2 KB in flight instead of 512 B: the one-token-per-warp
// kernel ran at half the byte floor (115 us per launch at 497,664 tokens), bound by its dependent shuffle chains.
struct Row128x4 {
  __device__ static __forceinline__ float sum8(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    return v;
  }
  __device__ static __forceinline__ void load(const float* __restrict__ row, int l8, float4 (&a)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = *reinterpret_cast<const float4*>(row + (i * 8 + l8) * 4);
  }
  __device__ static __forceinline__ void stats(const float4 (&a)[4], float eps, float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += (a[i].x + a[i].y) + (a[i].z + a[i].w);
    mean = sum8(s) * (1.0f / 128);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float d0 = a[i].x - mean, d1 = a[i].y - mean, d2 = a[i].z - mean, d3 = a[i].w - mean;
      q = fmaf(d0, d0, q);
      q = fmaf(d1, d1, q);
      q = fmaf(d2, d2, q);
      q = fmaf(d3, d3, q);
    }
    rstd = rsqrtf(sum8(q) * (1.0f / 128) + eps);
  }
  // g, b: [128] in shared memory
  __device__ static __forceinline__ void normalize(float4 (&a)[4], float mean, float rstd, const float* g, const float* b, int l8) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 gg = *reinterpret_cast<const float4*>(g + (i * 8 + l8) * 4), bb = *reinterpret_cast<const float4*>(b + (i * 8 + l8) * 4);
      a[i].x = fmaf((a[i].x - mean) * rstd, gg.x, bb.x);
      a[i].y = fmaf((a[i].y - mean) * rstd, gg.y, bb.y);
      a[i].z = fmaf((a[i].z - mean) * rstd, gg.z, bb.z);
      a[i].w = fmaf((a[i].w - mean) * rstd, gg.w, bb.w);
    }
  }
};

template <typename D, bool kPost>
__global__ void __launch_bounds__(kTokWarps * 32)
layernorm128_kernel(const float* __restrict__ x_in, float* __restrict__ x_out, uint16_t* __restrict__ h_out,
                    const float* __restrict__ post_g, const float* __restrict__ post_b, float post_eps, const float* __restrict__ pos,
                    int64_t pos_div, int64_t pos_mod, const float* __restrict__ ln_g, const float* __restrict__ ln_b, float ln_eps,
                    int64_t n_tokens) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) float prm[4][128];
  if (threadIdx.x < 128) {
    prm[0][threadIdx.x] = kPost ? post_g[threadIdx.x] : 1.f;
    prm[1][threadIdx.x] = kPost ? post_b[threadIdx.x] : 0.f;
    prm[2][threadIdx.x] = ln_g[threadIdx.x];
    prm[3][threadIdx.x] = ln_b[threadIdx.x];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, l8 = lane & 7, sub = lane >> 3;
  const int64_t warp_global = (int64_t)blockIdx.x * kTokWarps + (threadIdx.x >> 5);
  const int64_t stride = (int64_t)gridDim.x * kTokWarps * 4;
  for (int64_t tok = warp_global * 4 + sub; tok - sub < n_tokens; tok += stride) {
    const bool ok = tok < n_tokens;
    float4 a[4];
    if (ok) {
      Row128x4::load(x_in + tok * 128, l8, a);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float mean, rstd;
    if (kPost) {
      Row128x4::stats(a, post_eps, mean, rstd);
      Row128x4::normalize(a, mean, rstd, prm[0], prm[1], l8);
      if (pos != nullptr && ok) {
        const float* pe = pos + ((tok / pos_div) % pos_mod) * 128;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 p = __ldg(reinterpret_cast<const float4*>(pe + (i * 8 + l8) * 4));
          a[i].x += p.x;
          a[i].y += p.y;
          a[i].z += p.z;
          a[i].w += p.w;
        }
      }
      if (ok) {
#pragma unroll
        for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(x_out + tok * 128 + (i * 8 + l8) * 4) = a[i];
      }
    }
    Row128x4::stats(a, ln_eps, mean, rstd);
    Row128x4::normalize(a, mean, rstd, prm[2], prm[3], l8);
    if (ok) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint2 u;
        u.x = D::pack2(a[i].x, a[i].y);
        u.y = D::pack2(a[i].z, a[i].w);
        *reinterpret_cast<uint2*>(h_out + tok * 128 + (i * 8 + l8) * 4) = u;
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------- joint embedding
// x[tok, c] = W[c,0] in0 + W[c,1] in1 + b[c] + spos[tok % J, c]; h = LN(x)          (C = 512)
// A pure store stream (3 KB written per token, 8 bytes read).  TWO tokens per warp and pass, every parameter (W columns, bias, LayerNorm
// affine, the J position-embedding rows) in shared memory: 80 registers instead of 127 (parameters in registers: 16 resident warps per SM,
// one token in flight each, 470 us per 528,768 tokens = 3.4 TB/s where a fill reaches 7.5), the next pair's inputs requested a pass ahead,
// the two rows' statistics as independent shuffle chains.  Lane l owns channels [128 c + 4 l, + 4), c = 0..3: conflict-free 16-byte
// shared-memory reads, 512 contiguous bytes per store instruction.
template <typename D>
__global__ void __launch_bounds__(kTokWarps * 32, 3)
embed_joints_kernel(const float* __restrict__ in2d, const float* __restrict__ W, const float* __restrict__ bias,
                    const float* __restrict__ spos, const float* __restrict__ ln_g, const float* __restrict__ ln_b, float ln_eps,
                    float* __restrict__ x_out, uint16_t* __restrict__ h_out, int64_t n_tokens, int n_joints) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int C = 512;
  extern __shared__ __align__(16) float prm[];   // w0 | w1 | bias | gamma | beta | spos[n_joints]   ([C] each)
  float* sw0 = prm;
  float* sw1 = prm + C;
  float* sb = prm + 2 * C;
  float* sg = prm + 3 * C;
  float* sbt = prm + 4 * C;
  float* spe = prm + 5 * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    sw0[c] = W[c * 2 + 0];
    sw1[c] = W[c * 2 + 1];
    sb[c] = bias[c];
    sg[c] = ln_g[c];
    sbt[c] = ln_b[c];
  }
  for (int i = threadIdx.x; i < n_joints * C; i += blockDim.x) spe[i] = spos[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t n_pairs = (n_tokens + 1) / 2;
  const int64_t pair0 = (int64_t)blockIdx.x * kTokWarps + (threadIdx.x >> 5);
  const int64_t stride = (int64_t)gridDim.x * kTokWarps;
  const int jstep = (int)((2 * stride) % n_joints);
  int j0 = (int)((2 * pair0) % n_joints);
  const float2* in = reinterpret_cast<const float2*>(in2d);
  float2 pa = make_float2(0.f, 0.f), pb = pa;
  if (pair0 < n_pairs) {
    pa = __ldg(in + 2 * pair0);
    if (2 * pair0 + 1 < n_tokens) pb = __ldg(in + 2 * pair0 + 1);
  }
  for (int64_t pr = pair0; pr < n_pairs; pr += stride) {
    const int64_t tok = 2 * pr;
    const bool two = tok + 1 < n_tokens;
    const float2 p0 = pa, p1 = pb;
    if (pr + stride < n_pairs) {              // the next pass's inputs
      pa = __ldg(in + 2 * (pr + stride));
      if (2 * (pr + stride) + 1 < n_tokens) pb = __ldg(in + 2 * (pr + stride) + 1);
    }
    const int j1 = j0 + 1 == n_joints ? 0 : j0 + 1;
    float v0[16], v1[16];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int off = c * 128 + lane * 4;
      const float4 w0 = *reinterpret_cast<const float4*>(sw0 + off), w1 = *reinterpret_cast<const float4*>(sw1 + off);
      const float4 bb = *reinterpret_cast<const float4*>(sb + off);
      const float4 e0 = *reinterpret_cast<const float4*>(spe + j0 * C + off), e1 = *reinterpret_cast<const float4*>(spe + j1 * C + off);
      v0[4 * c + 0] = fmaf(w1.x, p0.y, fmaf(w0.x, p0.x, bb.x)) + e0.x;
      v0[4 * c + 1] = fmaf(w1.y, p0.y, fmaf(w0.y, p0.x, bb.y)) + e0.y;
      v0[4 * c + 2] = fmaf(w1.z, p0.y, fmaf(w0.z, p0.x, bb.z)) + e0.z;
      v0[4 * c + 3] = fmaf(w1.w, p0.y, fmaf(w0.w, p0.x, bb.w)) + e0.w;
      v1[4 * c + 0] = fmaf(w1.x, p1.y, fmaf(w0.x, p1.x, bb.x)) + e1.x;
      v1[4 * c + 1] = fmaf(w1.y, p1.y, fmaf(w0.y, p1.x, bb.y)) + e1.y;
      v1[4 * c + 2] = fmaf(w1.z, p1.y, fmaf(w0.z, p1.x, bb.z)) + e1.z;
      v1[4 * c + 3] = fmaf(w1.w, p1.y, fmaf(w0.w, p1.x, bb.w)) + e1.w;
      *reinterpret_cast<float4*>(x_out + tok * C + off) = make_float4(v0[4 * c + 0], v0[4 * c + 1], v0[4 * c + 2], v0[4 * c + 3]);
      if (two) *reinterpret_cast<float4*>(x_out + (tok + 1) * C + off) = make_float4(v1[4 * c + 0], v1[4 * c + 1], v1[4 * c + 2], v1[4 * c + 3]);
    }
    // two-pass statistics of both rows, the shuffle chains interleaved
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      s0 += v0[i];
      s1 += v1[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, o);
      s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    const float m0 = s0 * (1.0f / C), m1 = s1 * (1.0f / C);
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float d0 = v0[i] - m0, d1 = v1[i] - m1;
      q0 = fmaf(d0, d0, q0);
      q1 = fmaf(d1, d1, q1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      q0 += __shfl_xor_sync(0xffffffffu, q0, o);
      q1 += __shfl_xor_sync(0xffffffffu, q1, o);
    }
    const float r0 = rsqrtf(q0 * (1.0f / C) + ln_eps), r1 = rsqrtf(q1 * (1.0f / C) + ln_eps);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int off = c * 128 + lane * 4;
      const float4 gg = *reinterpret_cast<const float4*>(sg + off), bt = *reinterpret_cast<const float4*>(sbt + off);
      uint2 u0, u1;
      u0.x = D::pack2(fmaf((v0[4 * c + 0] - m0) * r0, gg.x, bt.x), fmaf((v0[4 * c + 1] - m0) * r0, gg.y, bt.y));
      u0.y = D::pack2(fmaf((v0[4 * c + 2] - m0) * r0, gg.z, bt.z), fmaf((v0[4 * c + 3] - m0) * r0, gg.w, bt.w));
      u1.x = D::pack2(fmaf((v1[4 * c + 0] - m1) * r1, gg.x, bt.x), fmaf((v1[4 * c + 1] - m1) * r1, gg.y, bt.y));
      u1.y = D::pack2(fmaf((v1[4 * c + 2] - m1) * r1, gg.z, bt.z), fmaf((v1[4 * c + 3] - m1) * r1, gg.w, bt.w));
      *reinterpret_cast<uint2*>(h_out + tok * C + off) = u0;
      if (two) *reinterpret_cast<uint2*>(h_out + (tok + 1) * C + off) = u1;
    }
    j0 += jstep;
    if (j0 >= n_joints) j0 -= n_joints;
  }
}

// -------------------------------------------------------------------------------------------------- segment embedding
// per frame: in[34] -> [16 segments x 128]; a CTA owns ONE segment (its 128 x 34 weight slice lives transposed in
// shared memory) and streams frames; token = frame * 16 + segment.  A warp takes kSegFrames consecutive frames at a time: their
// inputs (kSegFrames x in_features contiguous floats) are staged transposed in shared memory, so one 16-byte weight read per lane and
// two broadcast reads of the inputs feed 4 channels x 8 frames = 32 FMAs (one frame per pass re-read the whole weight slice for every
// frame and was bound by shared-memory bandwidth: 507 us per 128-clip micro-batch, 205 us now).  The
// accumulation order per output (bias, then features 0 .. in_features-1, then the position embedding) is unchanged.
constexpr int kSegC = 128;
constexpr int kSegFrames = 8;
template <typename D>
__global__ void __launch_bounds__(kTokWarps * 32, 3)
embed_segments_kernel(const float* __restrict__ in2d, const float* __restrict__ W, const float* __restrict__ bias,
                      const float* __restrict__ spos, const float* __restrict__ ln_g, const float* __restrict__ ln_b, float ln_eps,
                      float* __restrict__ x_out, uint16_t* __restrict__ h_out, int64_t n_frames, int in_features,
                      int n_segments) {
  pdl_launch_dependents();
  pdl_wait();
  using R = Row<kSegC>;
  extern __shared__ __align__(16) float wt[];  // [in_features][128], then per warp [in_features][kSegFrames]
  const int seg = blockIdx.y;
  for (int i = threadIdx.x; i < in_features * kSegC; i += blockDim.x) {
    const int c = i % kSegC, f = i / kSegC;
    wt[f * kSegC + c] = W[(size_t)(seg * kSegC + c) * in_features + f];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float* xs = wt + in_features * kSegC + (threadIdx.x >> 5) * (in_features * kSegFrames);
  float bb[R::kPer], pe[R::kPer], lg[R::kPer], lb[R::kPer];
  R::load_f32(bias + seg * kSegC, lane, bb);
  R::load_f32(spos + seg * kSegC, lane, pe);
  R::load_f32(ln_g, lane, lg);
  R::load_f32(ln_b, lane, lb);
  const int64_t n_groups = (n_frames + kSegFrames - 1) / kSegFrames;
  const int64_t warp_global = (int64_t)blockIdx.x * kTokWarps + (threadIdx.x >> 5);
  const int64_t stride = (int64_t)gridDim.x * kTokWarps;
  const int per_group = kSegFrames * in_features;
  for (int64_t grp = warp_global; grp < n_groups; grp += stride) {
    const int64_t fr0 = grp * kSegFrames;
    const int valid = (int)(n_frames - fr0 < kSegFrames ? n_frames - fr0 : kSegFrames);
    const float* in = in2d + fr0 * in_features;
    __syncwarp();                                   // the previous group's reads of xs are done
    for (int e = lane; e < per_group; e += 32) {
      const int fr = e / in_features, f = e - fr * in_features;
      xs[f * kSegFrames + fr] = fr < valid ? __ldg(in + e) : 0.f;
    }
    __syncwarp();
    // (packed fp32 FMAs on inputs stored twice were tried here: 16 instead of 32 FMA instructions per feature, but 12 instead of 8
    // shared-memory wavefronts - 240 us against 205 us; the loop waits for its shared-memory reads, not for issue slots)
    float v[kSegFrames][R::kPer];
#pragma unroll
    for (int r = 0; r < kSegFrames; ++r)
#pragma unroll
      for (int i = 0; i < R::kPer; ++i) v[r][i] = bb[i];
#pragma unroll 2
    for (int f = 0; f < in_features; ++f) {
      const float4 wv = *reinterpret_cast<const float4*>(wt + f * kSegC + lane * 4);
      const float4 xa = *reinterpret_cast<const float4*>(xs + f * kSegFrames);
      const float4 xb = *reinterpret_cast<const float4*>(xs + f * kSegFrames + 4);
      const float xin[kSegFrames] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
      for (int r = 0; r < kSegFrames; ++r) {
        v[r][0] = fmaf(wv.x, xin[r], v[r][0]);
        v[r][1] = fmaf(wv.y, xin[r], v[r][1]);
        v[r][2] = fmaf(wv.z, xin[r], v[r][2]);
        v[r][3] = fmaf(wv.w, xin[r], v[r][3]);
      }
    }
#pragma unroll
    for (int r = 0; r < kSegFrames; ++r) {
      if (r < valid) {
#pragma unroll
        for (int i = 0; i < R::kPer; ++i) v[r][i] += pe[i];
        const int64_t tok = (fr0 + r) * n_segments + seg;
        R::store_x(x_out + tok * kSegC, lane, v[r]);
      }
    }
    // the LayerNorm statistics of the kSegFrames rows together: independent shuffle chains
    float s[kSegFrames], mean[kSegFrames];
#pragma unroll
    for (int r = 0; r < kSegFrames; ++r) s[r] = (v[r][0] + v[r][1]) + (v[r][2] + v[r][3]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < kSegFrames; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
#pragma unroll
    for (int r = 0; r < kSegFrames; ++r) {
      mean[r] = s[r] * (1.0f / kSegC);
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < R::kPer; ++i) {
        const float d = v[r][i] - mean[r];
        q = fmaf(d, d, q);
      }
      s[r] = q;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int r = 0; r < kSegFrames; ++r) s[r] += __shfl_xor_sync(0xffffffffu, s[r], o);
#pragma unroll
    for (int r = 0; r < kSegFrames; ++r) {
      if (r < valid) {
        const float rstd = rsqrtf(s[r] * (1.0f / kSegC) + ln_eps);
        R::normalize(v[r], mean[r], rstd, lg, lb);
        const int64_t tok = (fr0 + r) * n_segments + seg;
        R::template store_h<D>(h_out + tok * kSegC, lane, v[r]);
      }
    }
  }
}

// -------------------------------------------------------------------------------------------------- hypothesis heads
// LN_k(x) W_k^T + b_k = rstd * (x . (g_k*W_k) - mean * sum(g_k*W_k)) + (W_k beta_k + b_k): the K LayerNorms share the
// token statistics, so the K heads are ONE [512 x K*O] fp32 projection with folded weights held in shared memory.
// A warp walks the 17 tokens of a frame four at a time: every folded-weight float4 read from shared memory feeds 16 FMAs
// (4 tokens x 4 channels), and the 4 x O partial sums of a head are reduced across the 32 lanes with one transposing
// butterfly (31 shuffles for up to 32 values) that leaves value i in lane i.
constexpr int kHeadC = 512;
constexpr int kHeadTok = 4;

// allreduce-free reduction of 32 per-lane values: afterwards lane l holds sum over lanes of vals[l]
__device__ __forceinline__ float transpose_reduce32(float (&vals)[32], int lane) {
#pragma unroll
  for (int step = 16; step >= 1; step >>= 1) {
    const bool up = (lane & step) != 0;
#pragma unroll
    for (int i = 0; i < step; ++i) {
      const float send = up ? vals[i] : vals[i + step];
      const float keep = up ? vals[i + step] : vals[i];
      vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
    }
  }
  return vals[0];
}

__global__ void __launch_bounds__(kTokWarps * 32)
heads_fwd_kernel(const float* __restrict__ x, const float* __restrict__ post_g, const float* __restrict__ post_b, float post_eps,
                 const float* __restrict__ hg, const float* __restrict__ hb, const float* __restrict__ hw, const float* __restrict__ hbias,
                 const float* __restrict__ score_w, const float* __restrict__ score_b, float* __restrict__ rot, float* __restrict__ logits,
                 int64_t n_clips, int n_frames, int n_hyp, int out_dim, int with_score) {
  pdl_launch_dependents();
  pdl_wait();
  using R = Row<kHeadC>;
  extern __shared__ __align__(16) float sm[];
  const int O = out_dim + (with_score ? 1 : 0);  // outputs per head (<= 7, so 4 tokens x O <= 28 values per butterfly)
  const int KO = n_hyp * O;
  float* wf = sm;                          // [KO][512] folded weights g_k[c] * W_k[o][c]
  float* csum = wf + (size_t)KO * kHeadC;  // [KO] sum_c wf
  float* dconst = csum + KO;               // [KO] W_k beta_k + b_k

  for (int i = threadIdx.x; i < KO * kHeadC; i += blockDim.x) {
    const int c = i % kHeadC, ko = i / kHeadC, k = ko / O;
    wf[i] = hg[k * kHeadC + c] * hw[(size_t)ko * kHeadC + c];
  }
  __syncthreads();
  for (int ko = threadIdx.x >> 5; ko < KO; ko += blockDim.x >> 5) {
    const int lane = threadIdx.x & 31, k = ko / O;
    float s = 0.f, dd = 0.f;
    for (int c = lane; c < kHeadC; c += 32) {
      s += wf[(size_t)ko * kHeadC + c];
      dd = fmaf(hw[(size_t)ko * kHeadC + c], hb[k * kHeadC + c], dd);
    }
    s = warp_sum(s);
    dd = warp_sum(dd);
    if (lane == 0) {
      csum[ko] = s;
      dconst[ko] = dd + hbias[ko];
    }
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float pg[R::kPer], pb[R::kPer];
  const bool has_post = post_g != nullptr;   // NULL: x already went through Temporal_norm (a standalone MCLHead.forward call)
  if (has_post) {
    R::load_f32(post_g, lane, pg);
    R::load_f32(post_b, lane, pb);
  }
  const int64_t total_frames = n_clips * n_frames;
  for (int64_t fr = (int64_t)blockIdx.x * kTokWarps + warp; fr < total_frames; fr += (int64_t)gridDim.x * kTokWarps) {
    const int64_t b = fr / n_frames;
    const int t = (int)(fr - b * n_frames);
    float logit_acc = 0.f;                 // lane k (< n_hyp) accumulates its head's logit
    for (int j0 = 0; j0 < kJ; j0 += kHeadTok) {
      // ---- four tokens: Temporal_norm (eps 1e-6) in fp32, then the shared statistics of the K head LayerNorms (eps 1e-5)
      float v[kHeadTok][R::kPer], mean[kHeadTok], rstd[kHeadTok];
#pragma unroll
      for (int q = 0; q < kHeadTok; ++q) {
        const int j = j0 + q < kJ ? j0 + q : kJ - 1;         // the tail group re-reads joint 16; its results are discarded
        R::load_x(x + (fr * kJ + j) * kHeadC, lane, v[q]);
        if (has_post) {
          float m, r;
          R::stats(v[q], post_eps, m, r);
          R::normalize(v[q], m, r, pg, pb);
        }
        R::stats(v[q], 1e-5f, mean[q], rstd[q]);
      }
      for (int k = 0; k < n_hyp; ++k) {
        float vals[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) vals[i] = 0.f;
#pragma unroll
        for (int o = 0; o < 7; ++o) {
          if (o < O) {
            const float* wrow = wf + (size_t)(k * O + o) * kHeadC;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float4 a = *reinterpret_cast<const float4*>(wrow + h * 256 + lane * 8);
              const float4 c4 = *reinterpret_cast<const float4*>(wrow + h * 256 + lane * 8 + 4);
#pragma unroll
              for (int q = 0; q < kHeadTok; ++q) {
                float acc = vals[q * 7 + o];
                acc = fmaf(v[q][h * 8 + 0], a.x, acc);
                acc = fmaf(v[q][h * 8 + 1], a.y, acc);
                acc = fmaf(v[q][h * 8 + 2], a.z, acc);
                acc = fmaf(v[q][h * 8 + 3], a.w, acc);
                acc = fmaf(v[q][h * 8 + 4], c4.x, acc);
                acc = fmaf(v[q][h * 8 + 5], c4.y, acc);
                acc = fmaf(v[q][h * 8 + 6], c4.z, acc);
                acc = fmaf(v[q][h * 8 + 7], c4.w, acc);
                vals[q * 7 + o] = acc;
              }
            }
          }
        }
        // value index q * 7 + o lands in lane q * 7 + o
        const float total = transpose_reduce32(vals, lane);
        const int q7 = lane / 7, o7 = lane - q7 * 7;
        const int j = j0 + q7;
        float contrib = 0.f;
        if (q7 < kHeadTok && o7 < O && j < kJ) {
          float mq = mean[0], rq = rstd[0];
#pragma unroll
          for (int q = 1; q < kHeadTok; ++q)
            if (q7 == q) {
              mq = mean[q];
              rq = rstd[q];
            }
          const int ko = k * O + o7;
          const float val = fmaf(rq, total - mq * csum[ko], dconst[ko]);
          if (o7 < out_dim) {
            rot[((((size_t)b * n_hyp + k) * n_frames + t) * kJ + j) * out_dim + o7] = val;
          } else {
            contrib = score_w[k * kJ + j] * val;               // score_head: Linear(17 -> 1)
          }
        }
        if (with_score) {
          const float s4 = __shfl_sync(0xffffffffu, contrib, out_dim) + __shfl_sync(0xffffffffu, contrib, 7 + out_dim) +
                           __shfl_sync(0xffffffffu, contrib, 14 + out_dim) + __shfl_sync(0xffffffffu, contrib, 21 + out_dim);
          if (lane == k) logit_acc += s4;
        }
      }
    }
    if (with_score && lane < n_hyp) logits[((size_t)b * n_hyp + lane) * n_frames + t] = logit_acc + score_b[lane];
  }
}

// -------------------------------------------------------------------------------------------------- bone-length head
// value[token] = Linear(128 -> 1)(LN_head(Temporal_norm(x)))  ;  bone_len[b, s] = mean_t value[b, t, s]
__global__ void __launch_bounds__(kTokWarps * 32)
bones_value_kernel(const float* __restrict__ x, const float* __restrict__ post_g, const float* __restrict__ post_b, float post_eps,
                   const float* __restrict__ hg, const float* __restrict__ hb, const float* __restrict__ hw, const float* __restrict__ hbias,
                   float* __restrict__ values, int64_t n_tokens) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) float prm[5][128];
  if (threadIdx.x < 128) {
    prm[0][threadIdx.x] = post_g[threadIdx.x];
    prm[1][threadIdx.x] = post_b[threadIdx.x];
    prm[2][threadIdx.x] = hg[threadIdx.x];
    prm[3][threadIdx.x] = hb[threadIdx.x];
    prm[4][threadIdx.x] = hw[threadIdx.x];
  }
  __syncthreads();
  // four tokens per warp (Row128x4): five 3-step butterflies per token instead of five 5-step ones
  const int lane = threadIdx.x & 31, l8 = lane & 7, sub = lane >> 3;
  const float b0 = hbias[0];
  const int64_t stride = (int64_t)gridDim.x * kTokWarps * 4;
  for (int64_t tok = ((int64_t)blockIdx.x * kTokWarps + (threadIdx.x >> 5)) * 4 + sub; tok - sub < n_tokens; tok += stride) {
    const bool ok = tok < n_tokens;
    float4 a[4];
    if (ok) {
      Row128x4::load(x + tok * 128, l8, a);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float mean, rstd;
    Row128x4::stats(a, post_eps, mean, rstd);
    Row128x4::normalize(a, mean, rstd, prm[0], prm[1], l8);
    Row128x4::stats(a, 1e-5f, mean, rstd);
    Row128x4::normalize(a, mean, rstd, prm[2], prm[3], l8);
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 w = *reinterpret_cast<const float4*>(prm[4] + (i * 8 + l8) * 4);
      acc = fmaf(a[i].x, w.x, acc);
      acc = fmaf(a[i].y, w.y, acc);
      acc = fmaf(a[i].z, w.z, acc);
      acc = fmaf(a[i].w, w.w, acc);
    }
    acc = Row128x4::sum8(acc);
    if (l8 == 0 && ok) values[tok] = acc + b0;
  }
}
// one CTA per clip: thread (tl, s) adds the frames tl, tl + lanes, ... of segment s (coalesced rows of n_seg values), the partial
// sums are combined in a fixed order (deterministic).  (One thread per (clip, segment) walking all the frames: 28 us of dependent loads.)
__global__ void __launch_bounds__(256) bones_mean_kernel(const float* __restrict__ values, float* __restrict__ bone_len, int64_t n_clips,
                                                         int n_frames, int n_seg) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float part[256];
  const int64_t b = blockIdx.x;
  const int lanes = 256 / n_seg;                 // n_seg <= 64
  const int s = threadIdx.x % n_seg, tl = threadIdx.x / n_seg;
  float acc = 0.f;
  if (tl < lanes) {
    const float* v = values + b * n_frames * n_seg + s;
    for (int t = tl; t < n_frames; t += lanes) acc += v[(int64_t)t * n_seg];
  }
  part[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < n_seg) {
    float tot = 0.f;
    for (int l = 0; l < lanes; ++l) tot += part[l * n_seg + threadIdx.x];
    bone_len[b * n_seg + threadIdx.x] = tot / (float)n_frames;
  }
}

// K heads, second half of the tensor-core path (mp_heads_fwd16): y [frames * 17, ld] fp32 holds, per token, the K * (D + 1) outputs of the
// folded projection (head k, output o at column k * (D + 1) + o).  One warp per frame: the 17 rows go through shared memory so that the
// [17 x D] block of every head is written as one contiguous run, and the J-term score dot product (MCLHead.score_head,
// rmcl_manifold_mix_ste.py:296-297) is a warp reduction.
__global__ void __launch_bounds__(kTokWarps * 32)
heads_finish_kernel(const float* __restrict__ y, int ld, const float* __restrict__ score_w, const float* __restrict__ score_b,
                    float* __restrict__ rot, float* __restrict__ logits, int64_t n_clips, int n_frames, int n_hyp, int out_dim, int with_score) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int O = out_dim + (with_score ? 1 : 0);
  const int row_floats = kJ * ld;                     // the frame's 17 rows are one contiguous run of the intermediate (ld % 4 == 0)
  float* tile = sm + (size_t)warp * row_floats;       // [17][ld]
  // index math once per launch, not per frame: element i of a head's [17 x out_dim] block comes from tile[(i / out_dim) * ld + i % out_dim]
  constexpr int kIt = (kJ * 6 + 31) / 32;             // out_dim <= 6
  int src[kIt];
#pragma unroll
  for (int it = 0; it < kIt; ++it) {
    const int i = it * 32 + lane;
    src[it] = i < kJ * out_dim ? (i / out_dim) * ld + i % out_dim : -1;
  }
  const int64_t total_frames = n_clips * n_frames;
  for (int64_t fr = (int64_t)blockIdx.x * kTokWarps + warp; fr < total_frames; fr += (int64_t)gridDim.x * kTokWarps) {
    const int64_t b = fr / n_frames;
    const int t = (int)(fr - b * n_frames);
    const float4* yrow = reinterpret_cast<const float4*>(y + fr * row_floats);
    for (int i = lane; i < row_floats / 4; i += 32) reinterpret_cast<float4*>(tile)[i] = yrow[i];
    __syncwarp();
    for (int k = 0; k < n_hyp; ++k) {
      float* dst = rot + (((b * n_hyp + k) * n_frames + t) * kJ) * out_dim;
#pragma unroll
      for (int it = 0; it < kIt; ++it)
        if (src[it] >= 0) dst[it * 32 + lane] = tile[src[it] + k * O];
      if (with_score) {
        float acc = lane < kJ ? score_w[k * kJ + lane] * tile[lane * ld + k * O + out_dim] : 0.f;
        acc = warp_sum(acc);
        if (lane == 0) logits[(b * n_hyp + k) * n_frames + t] = acc + score_b[k];
      }
    }
    __syncwarp();
  }
}

}  // namespace
}  // namespace mp

extern "C" {

int mp_layernorm(const float* x_in, float* x_out, void* h_out, const float* post_gamma, const float* post_beta, float post_eps,
                 const float* pos_embed, int64_t pos_div, int64_t pos_mod, const float* ln_gamma, const float* ln_beta, float ln_eps,
                 int64_t n_tokens, int C, int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(C == 512 || C == 128, MP_EUNSUPPORTED, "mp_layernorm: C=%d (built for 512 and 128)", C);
  MP_REQUIRE(x_in && n_tokens >= 0, MP_EINVAL, "mp_layernorm: bad arguments");
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_layernorm: unknown dtype %d", dtype);
  MP_REQUIRE((post_gamma == nullptr) == (post_beta == nullptr) && (ln_gamma == nullptr) == (ln_beta == nullptr), MP_EINVAL,
             "mp_layernorm: gamma/beta go together");
  MP_REQUIRE(post_gamma || ln_gamma, MP_EINVAL, "mp_layernorm: nothing to do");
  MP_REQUIRE(!post_gamma || x_out, MP_EINVAL, "mp_layernorm: x_out required with the post-norm");
  MP_REQUIRE(!ln_gamma || h_out, MP_EINVAL, "mp_layernorm: h_out required with the pre-norm");
  MP_REQUIRE(!pos_embed || (post_gamma && pos_div >= 1 && pos_mod >= 1), MP_EINVAL, "mp_layernorm: pos_embed needs the post-norm and pos_div/pos_mod >= 1");
  MP_REQUIRE(aligned16(x_in) && aligned16(x_out) && aligned16(h_out), MP_EALIGN, "mp_layernorm: activations must be 16-byte aligned");
  if (n_tokens == 0) return MP_OK;
  const int grid = token_grid(n_tokens);
  auto launch = [&](auto kernel) {
    launch_k(kernel, grid, kTokWarps * 32, 0, (cudaStream_t)stream, x_in, x_out, (uint16_t*)h_out, post_gamma, post_beta, post_eps, pos_embed,
                                                              pos_div, pos_mod, ln_gamma, ln_beta, ln_eps, n_tokens);
  };
  const bool bf = dtype == MP_DTYPE_BF16;
  if (C == 128 && ln_gamma != nullptr && h_out != nullptr && (post_gamma == nullptr || x_out != nullptr)) {
    const int grid4 = token_grid((n_tokens + 3) / 4);
    auto launch4 = [&](auto kernel) {
      launch_k(kernel, grid4, kTokWarps * 32, 0, (cudaStream_t)stream, x_in, x_out, (uint16_t*)h_out, post_gamma, post_beta, post_eps, pos_embed,
                                                                 pos_div, pos_mod, ln_gamma, ln_beta, ln_eps, n_tokens);
    };
    if (post_gamma != nullptr) {
      if (bf) launch4(layernorm128_kernel<Bf16, true>); else launch4(layernorm128_kernel<Fp16, true>);
    } else {
      if (bf) launch4(layernorm128_kernel<Bf16, false>); else launch4(layernorm128_kernel<Fp16, false>);
    }
    return check_launch("layernorm128_kernel");
  }
  if (C == 512) {
    if (bf) launch(layernorm_kernel<512, Bf16>); else launch(layernorm_kernel<512, Fp16>);
  } else {
    if (bf) launch(layernorm_kernel<128, Bf16>); else launch(layernorm_kernel<128, Fp16>);
  }
  return check_launch("layernorm_kernel");
}

int mp_embed_joints(const float* in2d, const float* W, const float* b, const float* spos, const float* ln_gamma, const float* ln_beta,
                    float ln_eps, float* x_out, void* h_out, int64_t n_tokens, int n_joints, int C, int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(C == 512, MP_EUNSUPPORTED, "mp_embed_joints: C=%d (built for 512)", C);
  MP_REQUIRE(in2d && W && b && spos && ln_gamma && ln_beta && x_out && h_out && n_tokens >= 0 && n_joints >= 1, MP_EINVAL,
             "mp_embed_joints: bad arguments");
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_embed_joints: unknown dtype %d", dtype);
  MP_REQUIRE(aligned16(x_out) && aligned16(h_out) && (reinterpret_cast<uintptr_t>(in2d) & 7u) == 0, MP_EALIGN,
             "mp_embed_joints: x_out / h_out must be 16-byte aligned, in2d 8-byte aligned");
  if (n_tokens == 0) return MP_OK;
  MP_REQUIRE(n_joints <= 64, MP_EUNSUPPORTED, "mp_embed_joints: %d joints (the position-embedding rows live in shared memory: at most 64)", n_joints);
  const int smem = (5 + n_joints) * 512 * (int)sizeof(float);
  int64_t ctas = ((n_tokens + 1) / 2 + kTokWarps - 1) / kTokWarps;
  const int64_t cap = (int64_t)sm_count() * 3;   // three resident CTAs per SM (registers)
  if (ctas > cap) ctas = cap;
  auto launch = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    launch_k(kernel, (int)ctas, kTokWarps * 32, smem, (cudaStream_t)stream, in2d, W, b, spos, ln_gamma, ln_beta, ln_eps, x_out,
                                                                   (uint16_t*)h_out, n_tokens, n_joints);
  };
  if (dtype == MP_DTYPE_BF16) launch(embed_joints_kernel<Bf16>); else launch(embed_joints_kernel<Fp16>);
  return check_launch("embed_joints_kernel");
}

int mp_embed_segments(const float* in2d, const float* W, const float* b, const float* spos, const float* ln_gamma, const float* ln_beta,
                      float ln_eps, float* x_out, void* h_out, int64_t n_frames, int in_features, int n_segments, int C, int dtype,
                      mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(C == kSegC, MP_EUNSUPPORTED, "mp_embed_segments: C=%d (built for 128)", C);
  MP_REQUIRE(in2d && W && b && spos && ln_gamma && ln_beta && x_out && h_out && n_frames >= 0, MP_EINVAL, "mp_embed_segments: bad arguments");
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_embed_segments: unknown dtype %d", dtype);
  MP_REQUIRE(in_features >= 1 && in_features <= 128 && n_segments >= 1 && n_segments <= 64, MP_EINVAL, "mp_embed_segments: bad sizes");
  MP_REQUIRE(aligned16(x_out) && aligned16(h_out), MP_EALIGN, "mp_embed_segments: x_out / h_out must be 16-byte aligned");
  if (n_frames == 0) return MP_OK;
  const size_t smem = (size_t)in_features * (kSegC + kTokWarps * kSegFrames) * sizeof(float);
  const int64_t n_groups = (n_frames + kSegFrames - 1) / kSegFrames;
  int gx = (int)((n_groups + kTokWarps - 1) / kTokWarps);
  int cap = sm_count() * 3 / n_segments;   // three resident CTAs per SM (80 registers): one wave
  if (cap < 1) cap = 1;
  if (gx > cap) gx = cap;
  auto launch = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(kernel, dim3(gx, n_segments), kTokWarps * 32, smem, (cudaStream_t)stream, in2d, W, b, spos, ln_gamma, ln_beta, ln_eps, x_out,
                                                                                 (uint16_t*)h_out, n_frames, in_features, n_segments);
  };
  if (dtype == MP_DTYPE_BF16) launch(embed_segments_kernel<Bf16>); else launch(embed_segments_kernel<Fp16>);
  return check_launch("embed_segments_kernel");
}

int mp_heads_fwd(const float* x, const float* post_gamma, const float* post_beta, float post_eps, const float* hg, const float* hb,
                 const float* hw, const float* hbias, const float* score_w, const float* score_b, float* rot, float* logits,
                 int64_t n_clips, int64_t n_frames, int n_hyp, int out_dim, int with_score, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(x && hg && hb && hw && hbias && rot && (post_gamma == nullptr) == (post_beta == nullptr), MP_EINVAL, "mp_heads_fwd: null pointer");
  MP_REQUIRE(!with_score || (score_w && score_b && logits), MP_EINVAL, "mp_heads_fwd: score head pointers required");
  MP_REQUIRE(n_hyp >= 1 && n_hyp <= 16 && (out_dim == 6 || out_dim == 4), MP_EINVAL, "mp_heads_fwd: bad n_hyp/out_dim");
  MP_REQUIRE(n_clips >= 0 && n_frames >= 1, MP_EINVAL, "mp_heads_fwd: bad sizes");
  MP_REQUIRE(aligned16(x), MP_EALIGN, "mp_heads_fwd: x must be 16-byte aligned");
  if (n_clips == 0) return MP_OK;
  const int O = out_dim + (with_score ? 1 : 0), KO = n_hyp * O;
  const size_t smem = ((size_t)KO * kHeadC + 2 * KO) * sizeof(float);
  MP_REQUIRE(smem <= 227 * 1024, MP_EUNSUPPORTED, "mp_heads_fwd: n_hyp=%d needs %zu bytes of shared memory", n_hyp, smem);
  cudaFuncSetAttribute(heads_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int64_t ctas = (n_clips * n_frames + kTokWarps - 1) / kTokWarps;
  const int64_t cap = (int64_t)sm_count() * (smem > 110 * 1024 ? 1 : 2);
  if (ctas > cap) ctas = cap;
  launch_k(heads_fwd_kernel, (int)ctas, kTokWarps * 32, smem, (cudaStream_t)stream, x, post_gamma, post_beta, post_eps, hg, hb, hw, hbias, score_w,
                                                                              score_b, rot, logits, n_clips, (int)n_frames, n_hyp, out_dim,
                                                                              with_score);
  return check_launch("heads_fwd_kernel");
}

int mp_heads_fwd16(const void* xhat16, const void* wf16, const float* bf, const float* score_w, const float* score_b, float* rot,
                   float* logits, float* workspace, size_t workspace_bytes, int64_t n_clips, int64_t n_frames, int n_hyp, int out_dim,
                   int with_score, int n_pad, int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(xhat16 && wf16 && bf && rot && workspace, MP_EINVAL, "mp_heads_fwd16: null pointer");
  MP_REQUIRE(aligned16(workspace) && aligned16(xhat16) && aligned16(wf16), MP_EALIGN, "mp_heads_fwd16: xhat16 / wf16 / workspace must be 16-byte aligned");
  MP_REQUIRE(!with_score || (score_w && score_b && logits), MP_EINVAL, "mp_heads_fwd16: score head pointers required");
  MP_REQUIRE(n_hyp >= 1 && n_hyp <= 16 && (out_dim == 6 || out_dim == 4 || out_dim == 3), MP_EINVAL, "mp_heads_fwd16: bad n_hyp/out_dim");
  const int O = out_dim + (with_score ? 1 : 0), KO = n_hyp * O;
  MP_REQUIRE(n_pad >= KO && n_pad % 128 == 0, MP_EINVAL, "mp_heads_fwd16: n_pad=%d must be a multiple of 128 holding %d outputs", n_pad, KO);
  MP_REQUIRE(n_clips >= 0 && n_frames >= 1, MP_EINVAL, "mp_heads_fwd16: bad sizes");
  const int64_t n_tokens = n_clips * n_frames * kJ;
  MP_REQUIRE(workspace_bytes >= (size_t)n_tokens * n_pad * sizeof(float), MP_EWORKSPACE, "mp_heads_fwd16: workspace too small");
  if (n_clips == 0) return MP_OK;
  // only the KO useful columns of the projection are written, as a dense [tokens, ld] fp32 matrix (ld = KO rounded up to 16 bytes), and read
  // back by the scatter kernel; MANIPOSE_HEADS_FULL_STORE=1 writes whole n_pad-column rows (A/B)
  const int ld = heads_ws_ld(n_hyp, out_dim, with_score, n_pad);
  if (ld != n_pad) {
    MP_CHECK(linear_f32_visible(xhat16, wf16, bf, workspace, n_tokens, n_pad, kHeadC, KO, ld, dtype, (cudaStream_t)stream));
  } else {
    MP_CHECK(mp_linear(xhat16, wf16, bf, nullptr, workspace, n_tokens, n_pad, kHeadC, MP_EPI_BIAS_F32, dtype, stream));
  }
  const size_t smem = (size_t)kTokWarps * kJ * ld * sizeof(float);
  MP_REQUIRE(smem <= 200 * 1024, MP_EUNSUPPORTED, "mp_heads_fwd16: n_hyp=%d needs %zu bytes of shared memory", n_hyp, smem);
  cudaFuncSetAttribute(heads_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int64_t ctas = (n_clips * n_frames + kTokWarps - 1) / kTokWarps;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (ctas > cap) ctas = cap;
  launch_k(heads_finish_kernel, (int)ctas, kTokWarps * 32, smem, (cudaStream_t)stream, workspace, ld, score_w, score_b, rot, logits, n_clips,
                                                                                 (int)n_frames, n_hyp, out_dim, with_score);
  return check_launch("heads_finish_kernel");
}

int mp_bones_head(const float* x, const float* post_gamma, const float* post_beta, float post_eps, const float* hg, const float* hb,
                  const float* hw, const float* hbias, float* bone_len, int64_t n_clips, int64_t n_frames, int n_segments, int C,
                  void* workspace, size_t workspace_bytes, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(C == kSegC, MP_EUNSUPPORTED, "mp_bones_head: C=%d (built for 128)", C);
  MP_REQUIRE(x && post_gamma && post_beta && hg && hb && hw && hbias && bone_len && workspace, MP_EINVAL, "mp_bones_head: null pointer");
  MP_REQUIRE(aligned16(x), MP_EALIGN, "mp_bones_head: x must be 16-byte aligned");
  MP_REQUIRE(n_segments >= 1 && n_segments <= 64 && n_clips < ((int64_t)1 << 31), MP_EINVAL, "mp_bones_head: bad sizes (1 <= n_segments <= 64)");
  const int64_t n_tokens = n_clips * n_frames * n_segments;
  MP_REQUIRE(workspace_bytes >= (size_t)n_tokens * sizeof(float), MP_EWORKSPACE, "mp_bones_head: workspace too small");
  if (n_tokens == 0) return MP_OK;
  float* values = reinterpret_cast<float*>(workspace);
  launch_k(bones_value_kernel, token_grid((n_tokens + 3) / 4), kTokWarps * 32, 0, (cudaStream_t)stream, x, post_gamma, post_beta, post_eps, hg, hb, hw, hbias,
                                                                                         values, n_tokens);
  MP_CHECK(check_launch("bones_value_kernel"));
  launch_k(bones_mean_kernel, (int)n_clips, 256, 0, (cudaStream_t)stream, values, bone_len, n_clips, (int)n_frames, n_segments);
  return check_launch("bones_mean_kernel");
}

}  // extern "C"

// Backward (training) kernels of the MixSTE backbone that are not GEMMs: LayerNorm, GELU, attention, the embeddings' weight
// gradients, the position-embedding reductions, the operand transposes of the weight-gradient GEMMs and Adam.
// The dense contractions of the backward pass (dgrad = dY W, wgrad = dY^T X) run on the tcgen05 Linear kernels of gemm.cu /
// gemm2.cu: dgrad with a transposed 16-bit weight shadow, wgrad as mp_linear over transposed operands (mp_transpose16) with the
// fp32 gradient buffer as the residual, so gradients accumulate in place.
//
// What is differentiated (reference, paths under hpe/mh_so3_hpe/architectures/; the reference relies on torch autograd):
//   mix_ste.py:352-358  Block.forward: x + attn(norm1(x)), x + mlp(norm2(x))            layernorm_bwd, attention_bwd, gelu_bwd
//   mix_ste.py:257-275  Attention.forward: softmax(q k^T * scale) v per head               attention_bwd
//   mix_ste.py:128-150  embeddings + position embeddings                                   small_wgrad, group_rowsum
//   hpe/main_h36m_lifting.py:755-761  torch.optim.Adam(lr, weight_decay) step              adam
#include "common.cuh"
#include "row.cuh"

namespace mp {
namespace {

// -------------------------------------------------------------------------------------------------- LayerNorm backward
// y = (x - mean) * rstd * gamma + beta.  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) [+ dres], g = dy * gamma;
// dgamma += sum dy * xhat, dbeta += sum dy (fp32 atomics, one per CTA and channel).  gamma == NULL: no affine.
template <int C, bool kDy16, typename D>
__global__ void __launch_bounds__(kTokWarps * 32)
layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, float eps, const void* dy,
                     const float* dres, float* dx, float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t n_tokens) {
  using R = Row<C>;
  __shared__ float acc[2][C];
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * kTokWarps + (threadIdx.x >> 5);
  const int64_t stride = (int64_t)gridDim.x * kTokWarps;
  float g[R::kPer], ag[R::kPer], ab[R::kPer];
#pragma unroll
  for (int i = 0; i < R::kPer; ++i) {
    g[i] = 1.f;
    ag[i] = 0.f;
    ab[i] = 0.f;
  }
  if (gamma) R::load_f32(gamma, lane, g);
  for (int64_t tok = warp_global; tok < n_tokens; tok += stride) {
    float v[R::kPer], d[R::kPer];
    R::load_x(x + tok * C, lane, v);
    if (kDy16)
      R::template load_h<D>(reinterpret_cast<const uint16_t*>(dy) + tok * C, lane, d);
    else
      R::load_x(reinterpret_cast<const float*>(dy) + tok * C, lane, d);
    float mean, rstd;
    R::stats(v, eps, mean, rstd);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < R::kPer; ++i) {
      const float xh = (v[i] - mean) * rstd;
      ag[i] = fmaf(d[i], xh, ag[i]);
      ab[i] += d[i];
      const float gi = d[i] * g[i];
      s1 += gi;
      s2 = fmaf(gi, xh, s2);
      v[i] = xh;
      d[i] = gi;
    }
    s1 = warp_sum(s1) * (1.0f / C);
    s2 = warp_sum(s2) * (1.0f / C);
#pragma unroll
    for (int i = 0; i < R::kPer; ++i) d[i] = rstd * (d[i] - s1 - v[i] * s2);
    if (dres) {
      float r[R::kPer];
      R::load_x(dres + tok * C, lane, r);
#pragma unroll
      for (int i = 0; i < R::kPer; ++i) d[i] += r[i];
    }
    R::store_x(dx + tok * C, lane, d);
  }
  if (dgamma) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      acc[0][c] = 0.f;
      acc[1][c] = 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < R::kPer; ++i) {
      atomicAdd(&acc[0][R::chan(lane, i)], ag[i]);
      atomicAdd(&acc[1][R::chan(lane, i)], ab[i]);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      atomicAdd(dgamma + c, acc[0][c]);
      atomicAdd(dbeta + c, acc[1][c]);
    }
  }
}

// -------------------------------------------------------------------------------------------------- GELU (training forward / backward)
// Phi(x) from the same erfc fit as gelu_erf; gelu'(x) = Phi(x) + x phi(x)
__device__ __forceinline__ float gelu_grad(float x) {
  const float t = fminf(fabsf(x) * 0.70710678118654752440f, 4.6f);
  float q = -9.749186599e-05f;
  q = fmaf(q, t, 4.431374392e-04f);
  q = fmaf(q, t, 2.348781295e-03f);
  q = fmaf(q, t, -2.950778651e-02f);
  q = fmaf(q, t, 1.489954364e-01f);
  q = fmaf(q, t, 9.183205755e-01f);
  q = fmaf(q, t, 1.627914397e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-t * q));
  const float h = 0.5f * e;
  const float cdf = x < 0.f ? h : 1.0f - h;
  return fmaf(x * 0.3989422804014327f, __expf(-0.5f * x * x), cdf);
}

template <typename D, bool kBwd>
__global__ void __launch_bounds__(256) gelu_kernel(const uint4* __restrict__ u, const uint4* __restrict__ da, uint4* __restrict__ out, int64_t n8) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 a = u[i];
    uint4 g = make_uint4(0, 0, 0, 0);
    if (kBwd) g = da[i];
    const uint32_t au[4] = {a.x, a.y, a.z, a.w}, gu[4] = {g.x, g.y, g.z, g.w};
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 xv = D::unpack2(au[k]);
      if (kBwd) {
        const float2 gv = D::unpack2(gu[k]);
        r[k] = D::pack2(gv.x * gelu_grad(xv.x), gv.y * gelu_grad(xv.y));
      } else {
        r[k] = D::pack2(gelu_erf(xv.x), gelu_erf(xv.y));
      }
    }
    out[i] = make_uint4(r[0], r[1], r[2], r[3]);
  }
}

// -------------------------------------------------------------------------------------------------- attention backward
// One thread per (item, row): item = (sequence, head), sequence = a frame (spatial, L = tokens) or a (clip, token) track
// (temporal, L = frames); G = 256 / L items per CTA.  fp32 SIMT with recomputation (no L x L matrix is stored):
//   pass 1 (thread = query row i):  lse_i, D_i = dO_i . O_i, dQ_i = scale * sum_j dS_ij K_j,  dS_ij = P_ij (dO_i . V_j - D_i)
//   pass 2 (thread = key row j):    dV_j = sum_i P_ij dO_i,  dK_j = scale * sum_i dS_ij Q_i   (head_dim 64: two column halves)
// K / V (pass 1) and scale * Q / dO (pass 2) of all rows of the CTA sit in shared memory as fp32 and are read as broadcasts.
template <int HD, typename D>
__device__ __forceinline__ void load_row16(const uint16_t* __restrict__ p, float (&v)[HD], float mul) {
#pragma unroll
  for (int c = 0; c < HD / 8; ++c) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(p) + c);
    const float2 a = D::unpack2(u.x), b = D::unpack2(u.y), e = D::unpack2(u.z), f = D::unpack2(u.w);
    v[8 * c + 0] = a.x * mul; v[8 * c + 1] = a.y * mul; v[8 * c + 2] = b.x * mul; v[8 * c + 3] = b.y * mul;
    v[8 * c + 4] = e.x * mul; v[8 * c + 5] = e.y * mul; v[8 * c + 6] = f.x * mul; v[8 * c + 7] = f.y * mul;
  }
}
template <int N>
__device__ __forceinline__ void store_smem_row(float* __restrict__ dst, const float (&v)[N]) {
#pragma unroll
  for (int c = 0; c < N / 4; ++c) reinterpret_cast<float4*>(dst)[c] = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
}
template <int N, typename D>
__device__ __forceinline__ void store_row16(uint16_t* __restrict__ p, const float (&v)[N], float mul) {
#pragma unroll
  for (int c = 0; c < N / 8; ++c) {
    uint4 u;
    u.x = D::pack2(v[8 * c + 0] * mul, v[8 * c + 1] * mul);
    u.y = D::pack2(v[8 * c + 2] * mul, v[8 * c + 3] * mul);
    u.z = D::pack2(v[8 * c + 4] * mul, v[8 * c + 5] * mul);
    u.w = D::pack2(v[8 * c + 6] * mul, v[8 * c + 7] * mul);
    reinterpret_cast<uint4*>(p)[c] = u;
  }
}
template <int HD>
__device__ __forceinline__ float dot_smem(const float (&a)[HD], const float* __restrict__ row) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int c = 0; c < HD / 4; ++c) {
    const float4 k = reinterpret_cast<const float4*>(row)[c];
    s0 = fmaf(a[4 * c + 0], k.x, s0);
    s1 = fmaf(a[4 * c + 1], k.y, s1);
    s2 = fmaf(a[4 * c + 2], k.z, s2);
    s3 = fmaf(a[4 * c + 3], k.w, s3);
  }
  return (s0 + s1) + (s2 + s3);
}

constexpr int kAttnBwdThreads = 256;
// shared-memory row stride in floats: +4 turns the 32-way bank conflict of "thread t writes row t" into 4-way and keeps float4 alignment
template <int HD>
struct AttnBwdSmem {
  static constexpr int kLd = HD + 4;
  static constexpr size_t kBytes = (size_t)(2 * kAttnBwdThreads * kLd + 2 * kAttnBwdThreads) * sizeof(float);
};

template <int HD, typename D>
__global__ void __launch_bounds__(kAttnBwdThreads, 1)
attention_bwd_kernel(const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ o, const uint16_t* __restrict__ dout,
                     uint16_t* __restrict__ dqkv, int n_items, int L, int G, int n_heads, int C, int n_tok, int n_frames, int temporal,
                     float scale) {
  extern __shared__ __align__(16) float sm[];
  constexpr int LD = AttnBwdSmem<HD>::kLd;
  float* buf_a = sm;                                // [256][LD]: K, then scale * Q
  float* buf_b = sm + kAttnBwdThreads * LD;         // [256][LD]: V, then dO
  float* s_lse = sm + 2 * kAttnBwdThreads * LD;     // [256]
  float* s_dd = s_lse + kAttnBwdThreads;            // [256]
  const int t = threadIdx.x;
  const int g = t / L, i = t - g * L;
  const int item = blockIdx.x * G + g;
  const bool active = g < G && item < n_items;
  int64_t tok = 0;
  int head = 0;
  if (active) {
    head = item % n_heads;
    const int seq = item / n_heads;
    if (temporal) {
      const int clip = seq / n_tok, tj = seq - clip * n_tok;
      tok = ((int64_t)clip * n_frames + i) * n_tok + tj;
    } else {
      tok = (int64_t)seq * n_tok + i;
    }
  }
  const uint16_t* qrow = qkv + tok * 3 * C + head * HD;
  const int r0 = g * L;

  float lse = 0.f, dd = 0.f;
  {
    float q[HD], dO[HD];
    if (active) {
      float kv[HD];
      load_row16<HD, D>(qrow + C, kv, 1.f);
      store_smem_row<HD>(buf_a + t * LD, kv);
      load_row16<HD, D>(qrow + 2 * C, kv, 1.f);
      store_smem_row<HD>(buf_b + t * LD, kv);
      load_row16<HD, D>(o + tok * C + head * HD, kv, 1.f);
      load_row16<HD, D>(qrow, q, scale);
      load_row16<HD, D>(dout + tok * C + head * HD, dO, 1.f);
#pragma unroll
      for (int c = 0; c < HD; ++c) dd = fmaf(dO[c], kv[c], dd);
    }
    __syncthreads();
    if (active) {
      float m = -INFINITY, l = 0.f;
      for (int j = 0; j < L; ++j) {
        const float s = dot_smem<HD>(q, buf_a + (r0 + j) * LD);
        const float mn = fmaxf(m, s);
        l = l * __expf(m - mn) + __expf(s - mn);
        m = mn;
      }
      const float inv = 1.0f / l;
      lse = m + __logf(l);
      float dq[HD];
#pragma unroll
      for (int c = 0; c < HD; ++c) dq[c] = 0.f;
      for (int j = 0; j < L; ++j) {
        const float* kr = buf_a + (r0 + j) * LD;
        const float s = dot_smem<HD>(q, kr);
        const float p = __expf(s - m) * inv;
        const float dp = dot_smem<HD>(dO, buf_b + (r0 + j) * LD);
        const float ds = p * (dp - dd);
#pragma unroll
        for (int c = 0; c < HD / 4; ++c) {
          const float4 k = reinterpret_cast<const float4*>(kr)[c];
          dq[4 * c + 0] = fmaf(ds, k.x, dq[4 * c + 0]);
          dq[4 * c + 1] = fmaf(ds, k.y, dq[4 * c + 1]);
          dq[4 * c + 2] = fmaf(ds, k.z, dq[4 * c + 2]);
          dq[4 * c + 3] = fmaf(ds, k.w, dq[4 * c + 3]);
        }
      }
      store_row16<HD, D>(dqkv + tok * 3 * C + head * HD, dq, scale);
    }
    __syncthreads();                                 // everybody is done with K / V
    if (active) {
      store_smem_row<HD>(buf_a + t * LD, q);         // scale * Q
      store_smem_row<HD>(buf_b + t * LD, dO);
      s_lse[t] = lse;
      s_dd[t] = dd;
    }
  }
  __syncthreads();
  if (!active) return;
  {
    constexpr int HH = HD >= 64 ? 32 : HD;           // columns of dK / dV accumulated per round
    float k[HD], v[HD];
    load_row16<HD, D>(qrow + C, k, 1.f);
    load_row16<HD, D>(qrow + 2 * C, v, 1.f);
#pragma unroll 1
    for (int half = 0; half < HD / HH; ++half) {
      float dk[HH], dv[HH];
#pragma unroll
      for (int c = 0; c < HH; ++c) {
        dk[c] = 0.f;
        dv[c] = 0.f;
      }
      for (int ii = 0; ii < L; ++ii) {
        const float* qr = buf_a + (r0 + ii) * LD;
        const float* dor = buf_b + (r0 + ii) * LD;
        const float s = dot_smem<HD>(k, qr);
        const float p = __expf(s - s_lse[r0 + ii]);
        const float dp = dot_smem<HD>(v, dor);
        const float ds = p * (dp - s_dd[r0 + ii]);
#pragma unroll
        for (int c = 0; c < HH / 4; ++c) {
          const float4 a = reinterpret_cast<const float4*>(dor + half * HH)[c];
          const float4 b = reinterpret_cast<const float4*>(qr + half * HH)[c];
          dv[4 * c + 0] = fmaf(p, a.x, dv[4 * c + 0]);
          dv[4 * c + 1] = fmaf(p, a.y, dv[4 * c + 1]);
          dv[4 * c + 2] = fmaf(p, a.z, dv[4 * c + 2]);
          dv[4 * c + 3] = fmaf(p, a.w, dv[4 * c + 3]);
          dk[4 * c + 0] = fmaf(ds, b.x, dk[4 * c + 0]);
          dk[4 * c + 1] = fmaf(ds, b.y, dk[4 * c + 1]);
          dk[4 * c + 2] = fmaf(ds, b.z, dk[4 * c + 2]);
          dk[4 * c + 3] = fmaf(ds, b.w, dk[4 * c + 3]);
        }
      }
      store_row16<HH, D>(dqkv + tok * 3 * C + C + head * HD + half * HH, dk, 1.f);
      store_row16<HH, D>(dqkv + tok * 3 * C + 2 * C + head * HD + half * HH, dv, 1.f);
    }
  }
}

// -------------------------------------------------------------------------------------------------- 16-bit transpose (+ column sums)
// dst[c, m] = src[m, c] for m < M, 0 for M <= m < Mpad (the reduction dim of the weight-gradient GEMM is padded to 64);
// colsum[c] += sum_m src[m, c] (the bias gradient), optional.
template <typename D>
__global__ void __launch_bounds__(256) transpose16_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, float* __restrict__ colsum,
                                                          int64_t M, int64_t C, int64_t Mpad) {
  __shared__ uint16_t tile[64][66];
  __shared__ float part[4][64];
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  const int64_t m0 = (int64_t)blockIdx.x * 64, c0 = (int64_t)blockIdx.y * 64;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int r = ty + 4 * k;
    const int64_t m = m0 + r;
    uint16_t v = 0;
    if (m < M) v = src[m * C + c0 + tx];
    tile[r][tx] = v;
    if (colsum) s += D::unpack2((uint32_t)v).x;
  }
  if (colsum) part[ty][tx] = s;
  __syncthreads();
  if (colsum && ty == 0) atomicAdd(colsum + c0 + tx, part[0][tx] + part[1][tx] + part[2][tx] + part[3][tx]);
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int r = ty + 4 * k;
    dst[(c0 + r) * Mpad + m0 + tx] = tile[tx][r];
  }
}

// -------------------------------------------------------------------------------------------------- grouped row sums (pos-embed gradients)
// out[(m / div) % mod, c] += x[m, c]: Spatial_pos_embed (div 1, mod tokens), Temporal_pos_embed (div tokens, mod frames)
__global__ void group_rowsum_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n_outer, int C, int64_t div, int64_t mod) {
  const int c = threadIdx.x;
  const int64_t gidx = blockIdx.x;
  float acc = 0.f;
  for (int64_t q = blockIdx.y; q < n_outer; q += gridDim.y) {
    const float* base = x + ((q * mod + gidx) * div) * C + c;
    for (int64_t r = 0; r < div; ++r) acc += base[r * C];
  }
  atomicAdd(out + gidx * C + c, acc);
}

// -------------------------------------------------------------------------------------------------- small-K weight gradient (embeddings)
// dW[o, i] += sum_r dy[r, o] in[r, i], db[o] += sum_r dy[r, o]; KIN = 2 (Spatial_patch_to_embedding) or 34 (joints_to_segments_proj)
template <int KIN>
__global__ void __launch_bounds__(128) small_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ in, float* __restrict__ dW,
                                                          float* __restrict__ db, int64_t R, int O) {
  const int o = blockIdx.x * 128 + threadIdx.x;
  float acc[KIN], ab = 0.f;
#pragma unroll
  for (int i = 0; i < KIN; ++i) acc[i] = 0.f;
  for (int64_t r = blockIdx.y; r < R; r += gridDim.y) {
    const float d = dy[r * O + o];
    ab += d;
#pragma unroll
    for (int i = 0; i < KIN; ++i) acc[i] = fmaf(d, __ldg(in + r * KIN + i), acc[i]);
  }
#pragma unroll
  for (int i = 0; i < KIN; ++i) atomicAdd(dW + (int64_t)o * KIN + i, acc[i]);
  atomicAdd(db + o, ab);
}

// -------------------------------------------------------------------------------------------------- stochastic depth (DropPath) glue
// out[m, :] = x[m, :] + s[m] * y[m, :]   (branch output y 16-bit, s = mask / keep_prob of the token's sample; mix_ste.py:352-358 with
// timm DropPath) and its backward operand g16[m, :] = 16-bit(s[m] * g[m, :]) (s NULL = plain cast).
template <int C, typename D>
__global__ void __launch_bounds__(kTokWarps * 32)
residual_rowscale_kernel(const float* __restrict__ x, const uint16_t* __restrict__ y, const float* __restrict__ s, float* __restrict__ out,
                         int64_t n_tokens) {
  using R = Row<C>;
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * kTokWarps;
  for (int64_t tok = (int64_t)blockIdx.x * kTokWarps + (threadIdx.x >> 5); tok < n_tokens; tok += stride) {
    float v[R::kPer], b[R::kPer];
    R::load_x(x + tok * C, lane, v);
    R::template load_h<D>(y + tok * C, lane, b);
    const float sc = __ldg(s + tok);
#pragma unroll
    for (int i = 0; i < R::kPer; ++i) v[i] = fmaf(sc, b[i], v[i]);
    R::store_x(out + tok * C, lane, v);
  }
}
template <int C, typename D>
__global__ void __launch_bounds__(kTokWarps * 32)
cast_rowscale_kernel(const float* __restrict__ g, const float* __restrict__ s, uint16_t* __restrict__ out, int64_t n_tokens) {
  using R = Row<C>;
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * kTokWarps;
  for (int64_t tok = (int64_t)blockIdx.x * kTokWarps + (threadIdx.x >> 5); tok < n_tokens; tok += stride) {
    float v[R::kPer];
    R::load_x(g + tok * C, lane, v);
    const float sc = s ? __ldg(s + tok) : 1.f;
#pragma unroll
    for (int i = 0; i < R::kPer; ++i) v[i] *= sc;
    R::template store_h<D>(out + tok * C, lane, v);
  }
}

// -------------------------------------------------------------------------------------------------- Adam (torch.optim.Adam semantics)
// step_dev != NULL: the step count lives on the device (value before this step; adam_bump_kernel increments it afterwards), so a
// captured CUDA graph of the training step advances the bias correction on every replay.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                   int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, float bc1,
                                                   float bc2_sqrt, float grad_scale, const int64_t* __restrict__ step_dev) {
  if (step_dev) {
    const double t = (double)(*step_dev + 1);
    bc1 = (float)(1.0 - pow((double)beta1, t));
    bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, t));
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float pi = p[i];
    const float gi = fmaf(weight_decay, pi, g[i] * grad_scale);
    const float mi = fmaf(beta1, m[i], (1.f - beta1) * gi);
    const float vi = fmaf(beta2, v[i], (1.f - beta2) * gi * gi);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - (lr / bc1) * (mi / denom);
  }
}

__global__ void adam_bump_kernel(int64_t* step_dev) { *step_dev += 1; }

int stream_grid(int64_t n, int per_cta) {
  int64_t ctas = (n + per_cta - 1) / per_cta;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(ctas < cap ? (ctas > 0 ? ctas : 1) : cap);
}

}  // namespace
}  // namespace mp

extern "C" {

int mp_layernorm_bwd(const float* x, const float* gamma, float eps, const void* dy, int dy_is_16bit, const float* dres, float* dx,
                     float* dgamma, float* dbeta, int64_t n_tokens, int C, int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(C == 512 || C == 128, MP_EUNSUPPORTED, "mp_layernorm_bwd: C=%d (built for 512 and 128)", C);
  MP_REQUIRE(x && dy && dx && n_tokens >= 0, MP_EINVAL, "mp_layernorm_bwd: bad arguments");
  MP_REQUIRE((dgamma == nullptr) == (dbeta == nullptr) && (gamma != nullptr || dgamma == nullptr), MP_EINVAL,
             "mp_layernorm_bwd: dgamma / dbeta come together and need gamma");
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_layernorm_bwd: unknown dtype %d", dtype);
  MP_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(dres), MP_EALIGN, "mp_layernorm_bwd: rows must be 16-byte aligned");
  if (n_tokens == 0) return MP_OK;
  // every CTA ends with 2 C atomics on the same dgamma / dbeta words: keep the grid at two CTAs per SM when they are wanted
  int grid = token_grid(n_tokens);
  if (dgamma && grid > 2 * sm_count()) grid = 2 * sm_count();
  auto launch = [&](auto kernel) {
    kernel<<<grid, kTokWarps * 32, 0, (cudaStream_t)stream>>>(x, gamma, eps, dy, dres, dx, dgamma, dbeta, n_tokens);
  };
  const bool bf = dtype == MP_DTYPE_BF16;
  if (C == 512) {
    if (!dy_is_16bit) launch(layernorm_bwd_kernel<512, false, Bf16>);
    else if (bf) launch(layernorm_bwd_kernel<512, true, Bf16>);
    else launch(layernorm_bwd_kernel<512, true, Fp16>);
  } else {
    if (!dy_is_16bit) launch(layernorm_bwd_kernel<128, false, Bf16>);
    else if (bf) launch(layernorm_bwd_kernel<128, true, Bf16>);
    else launch(layernorm_bwd_kernel<128, true, Fp16>);
  }
  return check_launch("layernorm_bwd_kernel");
}

static int gelu_launch(const void* u, const void* da, void* out, int64_t n, int dtype, bool bwd, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(u && out && (!bwd || da) && n >= 0 && n % 8 == 0, MP_EINVAL, "mp_gelu: bad arguments (n %% 8 == 0)");
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_gelu: unknown dtype %d", dtype);
  MP_REQUIRE(aligned16(u) && aligned16(da) && aligned16(out), MP_EALIGN, "mp_gelu: pointers must be 16-byte aligned");
  if (n == 0) return MP_OK;
  const int64_t n8 = n / 8;
  const int grid = stream_grid(n8, 256);
  auto launch = [&](auto kernel) { kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)u, (const uint4*)da, (uint4*)out, n8); };
  if (dtype == MP_DTYPE_BF16) {
    if (bwd) launch(gelu_kernel<Bf16, true>); else launch(gelu_kernel<Bf16, false>);
  } else {
    if (bwd) launch(gelu_kernel<Fp16, true>); else launch(gelu_kernel<Fp16, false>);
  }
  return check_launch("gelu_kernel");
}

int mp_gelu_fwd(const void* u, void* a, int64_t n, int dtype, mp_stream_t stream) { return gelu_launch(u, nullptr, a, n, dtype, false, stream); }

int mp_gelu_bwd(const void* u, const void* da, void* du, int64_t n, int dtype, mp_stream_t stream) {
  return gelu_launch(u, da, du, n, dtype, true, stream);
}

int mp_attention_bwd(const void* qkv, const void* o, const void* dout, void* dqkv, int64_t n_clips, int64_t n_frames, int n_tok, int C,
                     int n_heads, int mode, int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(qkv && o && dout && dqkv && n_clips >= 0 && n_frames >= 1 && n_tok >= 1 && n_heads >= 1, MP_EINVAL, "mp_attention_bwd: bad arguments");
  MP_REQUIRE(mode == MP_ATTN_SPATIAL || mode == MP_ATTN_TEMPORAL, MP_EINVAL, "mp_attention_bwd: unknown mode %d", mode);
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_attention_bwd: unknown dtype %d", dtype);
  const int hd = C / n_heads;
  MP_REQUIRE(C % n_heads == 0 && (hd == 64 || hd == 16), MP_EUNSUPPORTED, "mp_attention_bwd: head_dim %d (built for 64 and 16)", hd);
  const int temporal = mode == MP_ATTN_TEMPORAL;
  const int64_t L = temporal ? n_frames : n_tok;
  MP_REQUIRE(L <= kAttnBwdThreads, MP_EUNSUPPORTED, "mp_attention_bwd: sequence length %lld > %d", (long long)L, kAttnBwdThreads);
  MP_REQUIRE(aligned16(qkv) && aligned16(o) && aligned16(dout) && aligned16(dqkv), MP_EALIGN, "mp_attention_bwd: pointers must be 16-byte aligned");
  const int64_t n_seq = temporal ? n_clips * n_tok : n_clips * n_frames;
  const int64_t n_items = n_seq * n_heads;
  MP_REQUIRE(n_items < ((int64_t)1 << 31), MP_EINVAL, "mp_attention_bwd: too many (sequence, head) items");
  if (n_items == 0) return MP_OK;
  const int G = kAttnBwdThreads / (int)L;
  const int grid = (int)((n_items + G - 1) / G);
  const float scale = 1.0f / sqrtf((float)hd);
  auto launch = [&](auto kernel, int HD) -> int {
    const size_t smem = HD == 64 ? AttnBwdSmem<64>::kBytes : AttnBwdSmem<16>::kBytes;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    MP_REQUIRE(e == cudaSuccess, MP_ELAUNCH, "cudaFuncSetAttribute(attention_bwd_kernel): %s", cudaGetErrorString(e));
    kernel<<<grid, kAttnBwdThreads, smem, (cudaStream_t)stream>>>((const uint16_t*)qkv, (const uint16_t*)o, (const uint16_t*)dout, (uint16_t*)dqkv,
                                                                  (int)n_items, (int)L, G, n_heads, C, n_tok, (int)n_frames, temporal, scale);
    return check_launch("attention_bwd_kernel");
  };
  const bool bf = dtype == MP_DTYPE_BF16;
  if (hd == 64) return bf ? launch(attention_bwd_kernel<64, Bf16>, 64) : launch(attention_bwd_kernel<64, Fp16>, 64);
  return bf ? launch(attention_bwd_kernel<16, Bf16>, 16) : launch(attention_bwd_kernel<16, Fp16>, 16);
}

int mp_transpose16(const void* src, void* dst, float* colsum, int64_t M, int64_t C, int64_t Mpad, int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(src && dst && M >= 0 && C >= 64 && C % 64 == 0 && Mpad >= M && Mpad % 64 == 0, MP_EINVAL,
             "mp_transpose16: bad arguments (C %% 64 == 0, Mpad %% 64 == 0, Mpad >= M)");
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_transpose16: unknown dtype %d", dtype);
  if (Mpad == 0) return MP_OK;
  const dim3 grid((unsigned)(Mpad / 64), (unsigned)(C / 64));
  if (dtype == MP_DTYPE_BF16)
    transpose16_kernel<Bf16><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)src, (uint16_t*)dst, colsum, M, C, Mpad);
  else
    transpose16_kernel<Fp16><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t*)src, (uint16_t*)dst, colsum, M, C, Mpad);
  return check_launch("transpose16_kernel");
}

int mp_group_rowsum(const float* x, float* out, int64_t n_rows, int C, int64_t div, int64_t mod, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(x && out && n_rows >= 0 && C >= 32 && C <= 1024 && div >= 1 && mod >= 1 && n_rows % (div * mod) == 0, MP_EINVAL,
             "mp_group_rowsum: bad arguments (rows %% (div * mod) == 0, 32 <= C <= 1024)");
  if (n_rows == 0) return MP_OK;
  const int64_t n_outer = n_rows / (div * mod);
  int64_t splits = (int64_t)sm_count() * 4 / mod + 1;
  if (splits > n_outer) splits = n_outer;
  group_rowsum_kernel<<<dim3((unsigned)mod, (unsigned)splits), C, 0, (cudaStream_t)stream>>>(x, out, n_outer, C, div, mod);
  return check_launch("group_rowsum_kernel");
}

int mp_small_wgrad(const float* dy, const float* in, float* dW, float* db, int64_t n_rows, int n_out, int n_in, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(dy && in && dW && db && n_rows >= 0 && n_out >= 128 && n_out % 128 == 0, MP_EINVAL, "mp_small_wgrad: bad arguments (n_out %% 128 == 0)");
  MP_REQUIRE(n_in == 2 || n_in == 3 || n_in == 34 || n_in == 51, MP_EUNSUPPORTED, "mp_small_wgrad: n_in=%d (built for 2, 3, 34, 51)", n_in);
  if (n_rows == 0) return MP_OK;
  int64_t splits = (int64_t)sm_count() * 8 / (n_out / 128) + 1;
  if (splits > n_rows) splits = n_rows;
  const dim3 grid((unsigned)(n_out / 128), (unsigned)splits);
  cudaStream_t s = (cudaStream_t)stream;
  switch (n_in) {
    case 2: small_wgrad_kernel<2><<<grid, 128, 0, s>>>(dy, in, dW, db, n_rows, n_out); break;
    case 3: small_wgrad_kernel<3><<<grid, 128, 0, s>>>(dy, in, dW, db, n_rows, n_out); break;
    case 34: small_wgrad_kernel<34><<<grid, 128, 0, s>>>(dy, in, dW, db, n_rows, n_out); break;
    default: small_wgrad_kernel<51><<<grid, 128, 0, s>>>(dy, in, dW, db, n_rows, n_out); break;
  }
  return check_launch("small_wgrad_kernel");
}

int mp_residual_rowscale(const float* x, const void* y, const float* s, float* out, int64_t n_tokens, int C, int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(C == 512 || C == 128, MP_EUNSUPPORTED, "mp_residual_rowscale: C=%d (built for 512 and 128)", C);
  MP_REQUIRE(x && y && s && out && n_tokens >= 0, MP_EINVAL, "mp_residual_rowscale: bad arguments");
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_residual_rowscale: unknown dtype %d", dtype);
  MP_REQUIRE(aligned16(x) && aligned16(y) && aligned16(out), MP_EALIGN, "mp_residual_rowscale: rows must be 16-byte aligned");
  if (n_tokens == 0) return MP_OK;
  auto launch = [&](auto kernel) {
    kernel<<<token_grid(n_tokens), kTokWarps * 32, 0, (cudaStream_t)stream>>>(x, (const uint16_t*)y, s, out, n_tokens);
  };
  const bool bf = dtype == MP_DTYPE_BF16;
  if (C == 512) {
    if (bf) launch(residual_rowscale_kernel<512, Bf16>); else launch(residual_rowscale_kernel<512, Fp16>);
  } else {
    if (bf) launch(residual_rowscale_kernel<128, Bf16>); else launch(residual_rowscale_kernel<128, Fp16>);
  }
  return check_launch("residual_rowscale_kernel");
}

int mp_cast_rowscale(const float* g, const float* s, void* out, int64_t n_tokens, int C, int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(C == 512 || C == 128, MP_EUNSUPPORTED, "mp_cast_rowscale: C=%d (built for 512 and 128)", C);
  MP_REQUIRE(g && out && n_tokens >= 0, MP_EINVAL, "mp_cast_rowscale: bad arguments");
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_cast_rowscale: unknown dtype %d", dtype);
  MP_REQUIRE(aligned16(g) && aligned16(out), MP_EALIGN, "mp_cast_rowscale: rows must be 16-byte aligned");
  if (n_tokens == 0) return MP_OK;
  auto launch = [&](auto kernel) { kernel<<<token_grid(n_tokens), kTokWarps * 32, 0, (cudaStream_t)stream>>>(g, s, (uint16_t*)out, n_tokens); };
  const bool bf = dtype == MP_DTYPE_BF16;
  if (C == 512) {
    if (bf) launch(cast_rowscale_kernel<512, Bf16>); else launch(cast_rowscale_kernel<512, Fp16>);
  } else {
    if (bf) launch(cast_rowscale_kernel<128, Bf16>); else launch(cast_rowscale_kernel<128, Fp16>);
  }
  return check_launch("cast_rowscale_kernel");
}

int mp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int64_t step, int64_t* step_dev, float grad_scale, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0 && (step >= 1 || step_dev), MP_EINVAL,
             "mp_adam_step: bad arguments (step counts from 1, or pass the device counter)");
  if (n == 0) return MP_OK;
  float bc1 = 1.f, bc2_sqrt = 1.f;
  if (!step_dev) {
    bc1 = 1.0f - (float)pow((double)beta1, (double)step);
    bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  }
  adam_kernel<<<stream_grid(n, 1024), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                                       bc1, bc2_sqrt, grad_scale, step_dev);
  MP_CHECK(check_launch("adam_kernel"));
  if (step_dev) {
    adam_bump_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
    return check_launch("adam_bump_kernel");
  }
  return MP_OK;
}

}  // extern "C"

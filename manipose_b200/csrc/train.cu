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
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "row.cuh"

namespace mp {

int attention_bwd_tc(const void* qkv, const void* dout, void* dqkv, float* dqkv_colsum, int64_t n_clips, int64_t n_frames, int n_tok, int C,
                     int n_heads, int temporal, int dtype, cudaStream_t s);  // attention.cu
namespace {

// -------------------------------------------------------------------------------------------------- LayerNorm backward
// y = (x - mean) * rstd * gamma + beta.  dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) [+ dres], g = dy * gamma;
// dgamma += sum dy * xhat, dbeta += sum dy (fp32 atomics, one per CTA and channel).  gamma == NULL: no affine.
// dx16_colsum (may be NULL): += column sums of the dx16 output = the bias gradient of the Linear whose output gradient dx16 is.
// One CTA of kLnBwdWarps warps per SM, a warp per row; the NEXT row's x and dy are requested before the current row is reduced and
// stored (dx may alias dres, so the compiler cannot move those loads above the stores itself): two rows in flight per warp.
constexpr int kLnBwdWarps = 12;

template <int C, bool kDy16, typename D>
__global__ void __launch_bounds__(kLnBwdWarps * 32, 1)
layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, float eps, const void* dy,
                     const float* dres, float* dx, float* __restrict__ dgamma, float* __restrict__ dbeta, uint16_t* dx16,
                     const float* __restrict__ rowscale, float* __restrict__ dx16_colsum, int64_t n_tokens) {   // dx may alias dres, dx16 may alias dy
  pdl_launch_dependents();
  pdl_wait();
  using R = Row<C>;
  extern __shared__ __align__(16) float part[];   // [kLnBwdWarps][3][C] partial column sums (the tail of the kernel)
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (int64_t)blockIdx.x * kLnBwdWarps + (threadIdx.x >> 5);
  const int64_t stride = (int64_t)gridDim.x * kLnBwdWarps;
  float g[R::kPer], ag[R::kPer], ab[R::kPer], ac[R::kPer];
#pragma unroll
  for (int i = 0; i < R::kPer; ++i) {
    g[i] = 1.f;
    ag[i] = 0.f;
    ab[i] = 0.f;
    ac[i] = 0.f;
  }
  if (gamma) R::load_f32(gamma, lane, g);
  float vn[R::kPer], dn[R::kPer];
  auto fetch = [&](int64_t tok) {
    R::load_x(x + tok * C, lane, vn);
    if (kDy16)
      R::template load_h<D>(reinterpret_cast<const uint16_t*>(dy) + tok * C, lane, dn);
    else
      R::load_x(reinterpret_cast<const float*>(dy) + tok * C, lane, dn);
  };
  if (warp_global < n_tokens) fetch(warp_global);
  for (int64_t tok = warp_global; tok < n_tokens; tok += stride) {
    float v[R::kPer], d[R::kPer], r[R::kPer];
#pragma unroll
    for (int i = 0; i < R::kPer; ++i) {
      v[i] = vn[i];
      d[i] = dn[i];
      r[i] = 0.f;
    }
    if (dres) R::load_x(dres + tok * C, lane, r);
    if (tok + stride < n_tokens) fetch(tok + stride);
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
    for (int i = 0; i < R::kPer; ++i) d[i] = rstd * (d[i] - s1 - v[i] * s2) + r[i];
    R::store_x(dx + tok * C, lane, d);
    if (dx16) {   // the next backward GEMM's operand: 16-bit copy, scaled by the sample's stochastic-depth factor
      const float sc = rowscale ? __ldg(rowscale + tok) : 1.f;
#pragma unroll
      for (int i = 0; i < R::kPer; ++i) {
        d[i] *= sc;
        ac[i] += d[i];
      }
      R::template store_h<D>(dx16 + tok * C, lane, d);
    }
  }
  if (dgamma || dx16_colsum) {
    // every warp's partial column sums as rows of shared memory, added over the warps in a fixed order by the column's thread, then ONE
    // fp32 reduction per column and CTA.  (atomicAdd on shared floats compiles to compare-and-swap loops: with 12 warps on the same 512
    // words they were a third of this kernel's time, ncu: ATOMS.CAST.SPIN retried 6.5 times on average.)
    const int w = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < R::kPer; ++i) {
      const int c = R::chan(lane, i);
      part[(w * 3 + 0) * C + c] = ag[i];
      part[(w * 3 + 1) * C + c] = ab[i];
      part[(w * 3 + 2) * C + c] = ac[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int ww = 0; ww < kLnBwdWarps; ++ww) {
        t0 += part[(ww * 3 + 0) * C + c];
        t1 += part[(ww * 3 + 1) * C + c];
        t2 += part[(ww * 3 + 2) * C + c];
      }
      if (dgamma) {
        atomicAdd(dgamma + c, t0);
        atomicAdd(dbeta + c, t1);
      }
      if (dx16_colsum) atomicAdd(dx16_colsum + c, t2);
    }
  }
}

// C = 512 with the rows staged through shared memory by bulk copies.  The kernel above keeps ONE row ahead in registers (two for x / dy would
// not fit in 168 registers) and still waits for its loads at the top of every row (ncu: 40 % of the loop's stall samples on the first use of
// dy and dres).  Here every warp owns a ring of kLnRing row slots (x | dy | dres, 5 - 6 KB each); lane 0 requests the row kLnRing - 1
// iterations ahead with cp.async.bulk (completion on the slot's mbarrier) right after the warp has read a slot into registers, so the loads
// of the next rows are in flight during the whole of a row's reductions.  The partial column sums of the kernel's tail reuse the ring.
constexpr int kLnRing = 3;
template <bool kDy16>
constexpr int ln_ring_slot_bytes() { return 2048 + (kDy16 ? 1024 : 2048) + 2048; }

template <bool kDy16, typename D>
__global__ void __launch_bounds__(kLnBwdWarps * 32, 1)
layernorm_bwd_ring_kernel(const float* __restrict__ x, const float* __restrict__ gamma, float eps, const void* dy, const float* dres, float* dx,
                          float* __restrict__ dgamma, float* __restrict__ dbeta, uint16_t* dx16, const float* __restrict__ rowscale,
                          float* __restrict__ dx16_colsum, int64_t n_tokens) {   // dx may alias dres, dx16 may alias dy
  constexpr int C = 512;
  constexpr int kSlot = ln_ring_slot_bytes<kDy16>();
  constexpr int kDyBytes = kDy16 ? 1024 : 2048;
  pdl_launch_dependents();
  using R = Row<C>;
  extern __shared__ __align__(128) uint8_t ring_raw[];
  float* part = reinterpret_cast<float*>(ring_raw);                              // [kLnBwdWarps][3][C] after the loop
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint8_t* ring = ring_raw + (size_t)w * kLnRing * kSlot;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring_raw + (size_t)kLnBwdWarps * kLnRing * kSlot) + w * kLnRing;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kLnRing; ++s) ptx::mbar_init(&bars[s], 1);
    ptx::fence_mbar_init();
  }
  __syncwarp();
  pdl_wait();
  const int64_t warp_global = (int64_t)blockIdx.x * kLnBwdWarps + w;
  const int64_t stride = (int64_t)gridDim.x * kLnBwdWarps;
  const uint32_t row_bytes = 2048u + (uint32_t)kDyBytes + (dres ? 2048u : 0u);
  auto request = [&](int64_t tok, int slot) {        // lane 0
    uint8_t* sl = ring + slot * kSlot;
    ptx::mbar_expect_tx(&bars[slot], row_bytes);
    ptx::bulk_g2s(sl, x + tok * C, 2048, &bars[slot]);
    ptx::bulk_g2s(sl + 2048, reinterpret_cast<const uint8_t*>(dy) + tok * kDyBytes, kDyBytes, &bars[slot]);
    if (dres) ptx::bulk_g2s(sl + 2048 + kDyBytes, dres + tok * C, 2048, &bars[slot]);
  };
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kLnRing - 1; ++s)
      if (warp_global + s * stride < n_tokens) request(warp_global + s * stride, s);
  }
  // The per-channel arithmetic of a row runs on the packed fp32 instructions (add / mul / fma .f32x2: two IEEE operations per issue slot, same
  // results as the scalar forms): the kernel is bound by the length of a row's dependent instruction stream, ~14 packed instead of ~31 scalar
  // instructions per channel pair.
  constexpr int kPairs = R::kPer / 2;
  float g[R::kPer];
#pragma unroll
  for (int i = 0; i < R::kPer; ++i) g[i] = 1.f;
  if (gamma) R::load_f32(gamma, lane, g);
  uint64_t G2[kPairs], AG[kPairs], AB[kPairs], AC[kPairs];
#pragma unroll
  for (int k = 0; k < kPairs; ++k) {
    G2[k] = pack_f32x2(g[2 * k], g[2 * k + 1]);
    AG[k] = AB[k] = AC[k] = pack_f32x2(0.f, 0.f);
  }
  uint32_t n = 0;
  for (int64_t tok = warp_global; tok < n_tokens; tok += stride, ++n) {
    const int slot = (int)(n % kLnRing);
    // the row kLnRing - 1 iterations ahead goes into the slot the PREVIOUS iteration has read (all lanes are past their reads: __syncwarp below)
    if (lane == 0 && tok + (kLnRing - 1) * stride < n_tokens) request(tok + (kLnRing - 1) * stride, (int)((n + kLnRing - 1) % kLnRing));
    const float sc = (dx16 && rowscale) ? __ldg(rowscale + tok) : 1.f;   // requested now, used after the row's reductions
    ptx::mbar_wait(&bars[slot], (n / kLnRing) & 1);
    const uint8_t* sl = ring + slot * kSlot;
    float v[R::kPer], d[R::kPer], r[R::kPer];
    R::load_x(reinterpret_cast<const float*>(sl), lane, v);
    if (kDy16)
      R::template load_h<D>(reinterpret_cast<const uint16_t*>(sl + 2048), lane, d);
    else
      R::load_x(reinterpret_cast<const float*>(sl + 2048), lane, d);
    if (dres) {
      R::load_x(reinterpret_cast<const float*>(sl + 2048 + kDyBytes), lane, r);
    } else {
#pragma unroll
      for (int i = 0; i < R::kPer; ++i) r[i] = 0.f;
    }
    ptx::fence_proxy_async_smem();
    __syncwarp();                                   // the slot may be refilled by the next iteration's request
    uint64_t V[kPairs], Dp[kPairs];
#pragma unroll
    for (int k = 0; k < kPairs; ++k) {
      V[k] = pack_f32x2(v[2 * k], v[2 * k + 1]);
      Dp[k] = pack_f32x2(d[2 * k], d[2 * k + 1]);
    }
    // mean, then the biased variance around it (two passes, as nn.LayerNorm); V becomes x - mean
    uint64_t S = V[0];
#pragma unroll
    for (int k = 1; k < kPairs; ++k) S = add_f32x2(S, V[k]);
    const float mean = warp_sum(sum_f32x2(S)) * (1.0f / C);
    const uint64_t nm2 = pack_f32x2(-mean, -mean);
    uint64_t Q = pack_f32x2(0.f, 0.f);
#pragma unroll
    for (int k = 0; k < kPairs; ++k) {
      V[k] = add_f32x2(V[k], nm2);
      Q = fma_f32x2(V[k], V[k], Q);
    }
    const float rstd = rsqrtf(warp_sum(sum_f32x2(Q)) * (1.0f / C) + eps);
    const uint64_t r2 = pack_f32x2(rstd, rstd);
    uint64_t S1 = pack_f32x2(0.f, 0.f), S2 = S1;
#pragma unroll
    for (int k = 0; k < kPairs; ++k) {
      const uint64_t xh = mul_f32x2(V[k], r2);
      AG[k] = fma_f32x2(Dp[k], xh, AG[k]);
      AB[k] = add_f32x2(AB[k], Dp[k]);
      const uint64_t gi = mul_f32x2(Dp[k], G2[k]);
      S1 = add_f32x2(S1, gi);
      S2 = fma_f32x2(gi, xh, S2);
      V[k] = xh;
      Dp[k] = gi;
    }
    const float s1 = warp_sum(sum_f32x2(S1)) * (1.0f / C);
    const float s2 = warp_sum(sum_f32x2(S2)) * (1.0f / C);
    const uint64_t ns1 = pack_f32x2(-s1, -s1), ns2 = pack_f32x2(-s2, -s2);
#pragma unroll
    for (int k = 0; k < kPairs; ++k) {
      const uint64_t t = fma_f32x2(V[k], ns2, add_f32x2(Dp[k], ns1));
      Dp[k] = fma_f32x2(t, r2, pack_f32x2(r[2 * k], r[2 * k + 1]));
      unpack_f32x2(Dp[k], d[2 * k], d[2 * k + 1]);
    }
    R::store_x(dx + tok * C, lane, d);
    if (dx16) {
      const uint64_t sc2 = pack_f32x2(sc, sc);
#pragma unroll
      for (int k = 0; k < kPairs; ++k) {
        Dp[k] = mul_f32x2(Dp[k], sc2);
        AC[k] = add_f32x2(AC[k], Dp[k]);
        unpack_f32x2(Dp[k], d[2 * k], d[2 * k + 1]);
      }
      R::template store_h<D>(dx16 + tok * C, lane, d);
    }
  }
  float ag[R::kPer], ab[R::kPer], ac[R::kPer];
#pragma unroll
  for (int k = 0; k < kPairs; ++k) {
    unpack_f32x2(AG[k], ag[2 * k], ag[2 * k + 1]);
    unpack_f32x2(AB[k], ab[2 * k], ab[2 * k + 1]);
    unpack_f32x2(AC[k], ac[2 * k], ac[2 * k + 1]);
  }
  if (dgamma || dx16_colsum) {
    __syncthreads();                                // every warp has consumed all the rows it requested: the ring is free
#pragma unroll
    for (int i = 0; i < R::kPer; ++i) {
      const int c = R::chan(lane, i);
      part[(w * 3 + 0) * C + c] = ag[i];
      part[(w * 3 + 1) * C + c] = ab[i];
      part[(w * 3 + 2) * C + c] = ac[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float t0 = 0.f, t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int ww = 0; ww < kLnBwdWarps; ++ww) {
        t0 += part[(ww * 3 + 0) * C + c];
        t1 += part[(ww * 3 + 1) * C + c];
        t2 += part[(ww * 3 + 2) * C + c];
      }
      if (dgamma) {
        atomicAdd(dgamma + c, t0);
        atomicAdd(dbeta + c, t1);
      }
      if (dx16_colsum) atomicAdd(dx16_colsum + c, t2);
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

// two elements at once on the packed fp32 instructions (the polynomial, x^2 and the final FMA as .f32x2): the GELU backward kernels are
// bound by instruction issue (~30 instructions per element on 8 warps per scheduler)
__device__ __forceinline__ void gelu_grad2(float x0, float x1, float& g0, float& g1) {
  const float t0 = fminf(fabsf(x0) * 0.70710678118654752440f, 4.6f), t1 = fminf(fabsf(x1) * 0.70710678118654752440f, 4.6f);
  const uint64_t t = pack_f32x2(t0, t1);
  uint64_t q = pack_f32x2(-9.749186599e-05f, -9.749186599e-05f);
  q = fma_f32x2(q, t, pack_f32x2(4.431374392e-04f, 4.431374392e-04f));
  q = fma_f32x2(q, t, pack_f32x2(2.348781295e-03f, 2.348781295e-03f));
  q = fma_f32x2(q, t, pack_f32x2(-2.950778651e-02f, -2.950778651e-02f));
  q = fma_f32x2(q, t, pack_f32x2(1.489954364e-01f, 1.489954364e-01f));
  q = fma_f32x2(q, t, pack_f32x2(9.183205755e-01f, 9.183205755e-01f));
  q = fma_f32x2(q, t, pack_f32x2(1.627914397e+00f, 1.627914397e+00f));
  const uint64_t x = pack_f32x2(x0, x1);
  float a0, a1, b0, b1, e0, e1, p0, p1;
  unpack_f32x2(mul_f32x2(q, t), a0, a1);                                                       // erfc(t) = exp2(-t Q(t))
  unpack_f32x2(mul_f32x2(mul_f32x2(x, x), pack_f32x2(-0.72134752044448170368f, -0.72134752044448170368f)), b0, b1);   // exp(-x^2 / 2) = exp2(-x^2 log2(e) / 2)
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-a1));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(b0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(b1));
  const float h0 = 0.5f * e0, h1 = 0.5f * e1;
  const uint64_t cdf = pack_f32x2(x0 < 0.f ? h0 : 1.0f - h0, x1 < 0.f ? h1 : 1.0f - h1);
  unpack_f32x2(fma_f32x2(mul_f32x2(x, pack_f32x2(0.3989422804014327f, 0.3989422804014327f)), pack_f32x2(p0, p1), cdf), g0, g1);
}

template <typename D, bool kBwd>
__global__ void __launch_bounds__(256) gelu_kernel(const uint4* __restrict__ u, const uint4* __restrict__ da, uint4* __restrict__ out, int64_t n8) {
  pdl_launch_dependents();
  pdl_wait();
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

// -------------------------------------------------------------------------------------------------- attention backward on tensor cores
// mma.sync m16n8k16, P recomputed (nothing of size L x L is stored): one CTA per (sequence, head) -- a frame (spatial, L = tokens) or a
// (clip, token) track (temporal, L = frames) -- with Q / K / V / dO of the sequence in
// shared memory as 16-bit rows (+16 bytes of padding: conflict-free ldmatrix), each warp owns 16-row tiles:
//   sweep 0 (rows = queries)  lse_i = log2 sum_j exp2(s_ij * scale * log2 e)                         S = Q K^T
//   sweep A (rows = queries)  dS = P o (dO V^T - D_i),  dQ = scale * dS K                            P = exp2(S' - lse_i)
//   sweep B (rows = keys)     P^T = exp2((K Q^T)' - lse_i), dS^T = P^T o (V dO^T - D_i),  dV = P^T dO,  dK = scale * dS^T Q
// P and dS go from the accumulator layout straight into A fragments (16-bit), as in the forward kernel (attention.cu::attend_tile).
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// log2-domain log-sum-exp of the 16 score rows of one tile against columns [0, n_cols): returns rows g (lse[0]) and g + 8 (lse[1])
template <int HD, int LD, typename D>
__device__ __forceinline__ void lse_tile(uint32_t a_addr, uint32_t b_addr, int n_cols, int n_cols_pad, float scale_log2, int lane, float (&lse)[2]) {
  constexpr int KS = HD / 16;
  const int t = lane & 3;
  const uint32_t q_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * LD + (lane >> 4) * 16);
  const uint32_t k_off = (uint32_t)(((lane & 7) + (lane >> 4) * 8) * LD + ((lane >> 3) & 1) * 16);
  uint32_t af[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) ldsm_x4(af[ks], a_addr + q_off + ks * 32);
  float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
  for (int col0 = 0; col0 < n_cols_pad; col0 += 32) {
    const uint32_t bc = b_addr + k_off + (uint32_t)(col0 * LD);
    float sc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t bf[4];
        ldsm_x4(bf, bc + np * 16 * LD + ks * 32);
        ptx::mma_16816<D>(sc[np * 2 + 0], af[ks], bf[0], bf[1]);
        ptx::mma_16816<D>(sc[np * 2 + 1], af[ks], bf[2], bf[3]);
      }
    }
    if (col0 + 32 > n_cols) {
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (col0 + nt * 8 + 2 * t + (e & 1) >= n_cols) sc[nt][e] = -INFINITY;
    }
    float c0 = -INFINITY, c1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      c0 = fmaxf(c0, fmaxf(sc[nt][0], sc[nt][1]));
      c1 = fmaxf(c1, fmaxf(sc[nt][2], sc[nt][3]));
    }
    const float mn0 = fmaxf(m[0], quad_max(c0)), mn1 = fmaxf(m[1], quad_max(c1));
    l[0] *= fast_exp2((m[0] - mn0) * scale_log2);
    l[1] *= fast_exp2((m[1] - mn1) * scale_log2);
    m[0] = mn0;
    m[1] = mn1;
    const float ms0 = -mn0 * scale_log2, ms1 = -mn1 * scale_log2;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      l[0] += fast_exp2(fmaf(sc[nt][0], scale_log2, ms0)) + fast_exp2(fmaf(sc[nt][1], scale_log2, ms0));
      l[1] += fast_exp2(fmaf(sc[nt][2], scale_log2, ms1)) + fast_exp2(fmaf(sc[nt][3], scale_log2, ms1));
    }
  }
  lse[0] = fmaf(m[0], scale_log2, log2f(quad_sum(l[0])));
  lse[1] = fmaf(m[1], scale_log2, log2f(quad_sum(l[1])));
}

// One 16-row tile against all columns.  a1/a2: the tile's rows of the two "row side" operands (score GEMM, dP GEMM); b1/b2: row 0 of
// the "column side" operands; x1/x2: row 0 of the operands contracted over the columns (x1 with P -> out1, only kByCol; x2 with dS ->
// out2).  Statistics lse / dd are indexed by row (sweep A) or by column (sweep B, kByCol).  Padded columns contribute nothing.
template <int HD, int LD, typename D, bool kByCol>
__device__ __forceinline__ void bwd_sweep(uint32_t a1_addr, uint32_t a2_addr, uint32_t b1_addr, uint32_t b2_addr, uint32_t x1_addr,
                                          uint32_t x2_addr, int n_cols, int n_cols_pad, float scale_log2, const float* __restrict__ lse,
                                          const float* __restrict__ dd, int row0, int lane, float (&out1)[HD / 8][4], float (&out2)[HD / 8][4]) {
  constexpr int KS = HD / 16, NT = HD / 8;
  const int g = lane >> 2, t = lane & 3;
  const uint32_t q_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * LD + (lane >> 4) * 16);
  const uint32_t k_off = (uint32_t)(((lane & 7) + (lane >> 4) * 8) * LD + ((lane >> 3) & 1) * 16);
  const uint32_t v_off = q_off;
  uint32_t a1f[KS][4], a2f[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    ldsm_x4(a1f[ks], a1_addr + q_off + ks * 32);
    ldsm_x4(a2f[ks], a2_addr + q_off + ks * 32);
  }
#pragma unroll
  for (int i = 0; i < NT; ++i) {
    out1[i][0] = out1[i][1] = out1[i][2] = out1[i][3] = 0.f;
    out2[i][0] = out2[i][1] = out2[i][2] = out2[i][3] = 0.f;
  }
  float lr[2] = {0.f, 0.f}, dr[2] = {0.f, 0.f};
  if (!kByCol) {
    lr[0] = lse[row0 + g];
    lr[1] = lse[row0 + g + 8];
    dr[0] = dd[row0 + g];
    dr[1] = dd[row0 + g + 8];
  }
  for (int col0 = 0; col0 < n_cols_pad; col0 += 32) {
    const uint32_t b1c = b1_addr + k_off + (uint32_t)(col0 * LD);
    const uint32_t b2c = b2_addr + k_off + (uint32_t)(col0 * LD);
    float sc[4][4], dp[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      sc[i][0] = sc[i][1] = sc[i][2] = sc[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t bf[4];
        ldsm_x4(bf, b1c + np * 16 * LD + ks * 32);
        ptx::mma_16816<D>(sc[np * 2 + 0], a1f[ks], bf[0], bf[1]);
        ptx::mma_16816<D>(sc[np * 2 + 1], a1f[ks], bf[2], bf[3]);
        ldsm_x4(bf, b2c + np * 16 * LD + ks * 32);
        ptx::mma_16816<D>(dp[np * 2 + 0], a2f[ks], bf[0], bf[1]);
        ptx::mma_16816<D>(dp[np * 2 + 1], a2f[ks], bf[2], bf[3]);
      }
    }
    uint32_t pf[2][4], dsf[2][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float pe[4], ds[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int col = col0 + nt * 8 + 2 * t + (e & 1);
        const float l_ = kByCol ? lse[col] : lr[e >> 1];
        const float d_ = kByCol ? dd[col] : dr[e >> 1];
        float v = fast_exp2(fmaf(sc[nt][e], scale_log2, -l_));
        if (col >= n_cols) v = 0.f;
        pe[e] = v;
        ds[e] = v * (dp[nt][e] - d_);
      }
      pf[nt >> 1][(nt & 1) * 2 + 0] = D::pack2(pe[0], pe[1]);
      pf[nt >> 1][(nt & 1) * 2 + 1] = D::pack2(pe[2], pe[3]);
      dsf[nt >> 1][(nt & 1) * 2 + 0] = D::pack2(ds[0], ds[1]);
      dsf[nt >> 1][(nt & 1) * 2 + 1] = D::pack2(ds[2], ds[3]);
    }
    const uint32_t x1c = x1_addr + v_off + (uint32_t)(col0 * LD);
    const uint32_t x2c = x2_addr + v_off + (uint32_t)(col0 * LD);
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
      for (int dpi = 0; dpi < NT / 2; ++dpi) {
        uint32_t xf[4];
        if (kByCol) {
          ldsm_x4_t(xf, x1c + kk * 16 * LD + dpi * 32);
          ptx::mma_16816<D>(out1[dpi * 2 + 0], pf[kk], xf[0], xf[1]);
          ptx::mma_16816<D>(out1[dpi * 2 + 1], pf[kk], xf[2], xf[3]);
        }
        ldsm_x4_t(xf, x2c + kk * 16 * LD + dpi * 32);
        ptx::mma_16816<D>(out2[dpi * 2 + 0], dsf[kk], xf[0], xf[1]);
        ptx::mma_16816<D>(out2[dpi * 2 + 1], dsf[kk], xf[2], xf[3]);
      }
    }
  }
}

constexpr int kAttnBwdMmaWarps = 8;

template <int HD, typename D>
__global__ void __launch_bounds__(kAttnBwdMmaWarps * 32)
attention_bwd_mma_kernel(const uint16_t* __restrict__ qkv, const uint16_t* __restrict__ o, const uint16_t* __restrict__ dout,
                         uint16_t* __restrict__ dqkv, int L, int Lp, int n_heads, int C, int n_tok, int n_frames, int temporal) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int LD = HD * 2 + 16;
  constexpr int CH = HD / 8;
  extern __shared__ __align__(16) uint8_t smem_b[];
  uint8_t* sq = smem_b;
  uint8_t* sk = sq + (size_t)Lp * LD;
  uint8_t* sv = sk + (size_t)Lp * LD;
  uint8_t* sdo = sv + (size_t)Lp * LD;
  uint8_t* so = sdo + (size_t)Lp * LD;          // forward output rows, only for D_i
  float* s_lse = reinterpret_cast<float*>(so + (size_t)Lp * LD);
  float* s_dd = s_lse + Lp;
  uint8_t* stage_base = reinterpret_cast<uint8_t*>(s_dd + Lp);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int head = blockIdx.x % n_heads, seq = blockIdx.x / n_heads;
  int64_t tok0, tstride;
  if (temporal) {
    const int clip = seq / n_tok, tj = seq - clip * n_tok;
    tok0 = (int64_t)clip * n_frames * n_tok + tj;
    tstride = n_tok;
  } else {
    tok0 = (int64_t)seq * n_tok;
    tstride = 1;
  }
  const uint16_t* qkv_base = qkv + tok0 * 3 * C + head * HD;
  const uint16_t* do_base = dout + tok0 * C + head * HD;
  const uint16_t* o_base = o + tok0 * C + head * HD;
  uint16_t* dqkv_base = dqkv + tok0 * 3 * C + head * HD;
  const int64_t qkv_stride = tstride * 3 * C, o_stride = tstride * C;

  // ---- stage Q, K, V, dO, O rows [0, L); zero rows [L, Lp)
  const int total = 5 * Lp * CH;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int ch = i % CH;
    const int r = (i / CH) % Lp;
    const int sel = i / (CH * Lp);
    uint8_t* dst = smem_b + ((size_t)sel * Lp + r) * LD + ch * 16;
    if (r < L) {
      const uint16_t* src = sel < 3 ? qkv_base + (int64_t)r * qkv_stride + sel * C + ch * 8
                                    : (sel == 3 ? do_base : o_base) + (int64_t)r * o_stride + ch * 8;
      ptx::cp_async16(dst, src);
    } else {
      *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
  }
  ptx::cp_async_commit();
  ptx::cp_async_wait<0>();
  __syncthreads();
  // ---- D_i = dO_i . O_i
  for (int r = threadIdx.x; r < Lp; r += blockDim.x) {
    float acc = 0.f;
    if (r < L) {
#pragma unroll
      for (int ch = 0; ch < CH; ++ch) {
        const uint4 a = *reinterpret_cast<const uint4*>(sdo + (size_t)r * LD + ch * 16);
        const uint4 b = *reinterpret_cast<const uint4*>(so + (size_t)r * LD + ch * 16);
        const uint32_t au[4] = {a.x, a.y, a.z, a.w}, bu[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 x = D::unpack2(au[k]), y = D::unpack2(bu[k]);
          acc = fmaf(x.x, y.x, fmaf(x.y, y.y, acc));
        }
      }
    }
    s_dd[r] = acc;
    s_lse[r] = 0.f;
  }
  __syncthreads();
  const float scale = rsqrtf((float)HD);
  const float scale_log2 = scale * 1.4426950408889634f;
  uint8_t* stg = stage_base + (size_t)warp * 16 * LD;
  auto store_tile = [&](const float (&v)[HD / 8][4], float mul, int mt, int col_off) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      *reinterpret_cast<uint32_t*>(stg + (size_t)g * LD + (i * 8 + 2 * t) * 2) = D::pack2(v[i][0] * mul, v[i][1] * mul);
      *reinterpret_cast<uint32_t*>(stg + (size_t)(g + 8) * LD + (i * 8 + 2 * t) * 2) = D::pack2(v[i][2] * mul, v[i][3] * mul);
    }
    __syncwarp();
    for (int i = lane; i < 16 * CH; i += 32) {
      const int rr = i / CH, ch = i % CH, r = mt * 16 + rr;
      if (r < L)
        *reinterpret_cast<uint4*>(dqkv_base + (int64_t)r * qkv_stride + col_off + ch * 8) = *reinterpret_cast<const uint4*>(stg + (size_t)rr * LD + ch * 16);
    }
  };
  // ---- sweep 0 + A per 16-row query tile: log-sum-exp of the tile's rows (needed by every warp in sweep B), then dQ
  for (int mt = warp; mt * 16 < L; mt += n_warps) {
    float unused[HD / 8][4], dq[HD / 8][4];
    const uint32_t ro = (uint32_t)(mt * 16 * LD);
    {
      float lse[2];
      lse_tile<HD, LD, D>(smem_u32(sq) + ro, smem_u32(sk), L, Lp, scale_log2, lane, lse);
      if (t == 0) {
        s_lse[mt * 16 + g] = lse[0];
        s_lse[mt * 16 + g + 8] = lse[1];
      }
      __syncwarp();
    }
    bwd_sweep<HD, LD, D, false>(smem_u32(sq) + ro, smem_u32(sdo) + ro, smem_u32(sk), smem_u32(sv), 0u, smem_u32(sk), L, Lp, scale_log2, s_lse,
                                s_dd, mt * 16, lane, unused, dq);
    store_tile(dq, scale, mt, 0);
  }
  __syncthreads();
  // ---- sweep B: dK, dV
  for (int mt = warp; mt * 16 < L; mt += n_warps) {
    float dv[HD / 8][4], dk[HD / 8][4];
    const uint32_t ro = (uint32_t)(mt * 16 * LD);
    bwd_sweep<HD, LD, D, true>(smem_u32(sk) + ro, smem_u32(sv) + ro, smem_u32(sq), smem_u32(sdo), smem_u32(sdo), smem_u32(sq), L, Lp, scale_log2,
                               s_lse, s_dd, mt * 16, lane, dv, dk);
    store_tile(dk, scale, mt, C);
    store_tile(dv, 1.0f, mt, 2 * C);
  }
}

// -------------------------------------------------------------------------------------------------- 16-bit transpose (+ column sums)
// dst[c, m] = src[m, c] for m < M, 0 for M <= m < Mpad (the reduction dim of the weight-gradient GEMM is padded to 64);
// colsum[c] += sum_m src[m, c] (the bias gradient), optional.
template <typename D>
__global__ void __launch_bounds__(256) transpose16_kernel(const uint16_t* __restrict__ src, uint16_t* __restrict__ dst, float* __restrict__ colsum,
                                                          int64_t M, int64_t C, int64_t Mpad) {
  pdl_launch_dependents();
  pdl_wait();
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

// -------------------------------------------------------------------------------------------------- weight shadows, one launch
// After an optimizer step every GEMM weight needs its 16-bit shadow [N, K] (forward, wgrad-free) and the transposed shadow [K, N]
// (dgrad).  table[w] = {src fp32 ptr, dst ptr, dst_t ptr, rows N, cols K} as int64; blockIdx.y = weight, blockIdx.x strides over its
// 64 x 64 tiles.
template <typename D>
__global__ void __launch_bounds__(256) refresh_shadows_kernel(const int64_t* __restrict__ table) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ uint16_t tile[64][66];
  const int64_t* e = table + (int64_t)blockIdx.y * 5;
  const float* src = reinterpret_cast<const float*>(e[0]);
  uint16_t* dst = reinterpret_cast<uint16_t*>(e[1]);
  uint16_t* dst_t = reinterpret_cast<uint16_t*>(e[2]);
  const int rows = (int)e[3], cols = (int)e[4];
  const int tiles_c = cols / 64, n_tiles = (rows / 64) * tiles_c;
  const int tx = threadIdx.x & 63, ty = threadIdx.x >> 6;
  for (int tl = blockIdx.x; tl < n_tiles; tl += gridDim.x) {
    const int r0 = (tl / tiles_c) * 64, c0 = (tl % tiles_c) * 64;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int r = ty + 4 * k;
      const uint16_t v = (uint16_t)(D::pack2(src[(size_t)(r0 + r) * cols + c0 + tx], 0.f) & 0xffffu);
      tile[r][tx] = v;
      dst[(size_t)(r0 + r) * cols + c0 + tx] = v;
    }
    __syncthreads();
    if (dst_t != nullptr) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int r = ty + 4 * k;
        dst_t[(size_t)(c0 + r) * rows + r0 + tx] = tile[tx][r];
      }
    }
    __syncthreads();
  }
}

// -------------------------------------------------------------------------------------------------- bias gradient
// colsum[c] += sum_m src[m, c] over a 16-bit [M, C] matrix, C % 8 == 0.  CTA = 64 columns (8 threads x one 16-byte load = one 128-byte
// row segment) x a slab of rows; 32 row lanes per CTA with four independent loads in flight each, reduced in shared memory, then ONE
// atomic per column and CTA.  (A first version read 4 bytes per thread with one load in flight: latency-bound at ~0.8 TB/s.)
template <typename D>
__global__ void __launch_bounds__(256) colsum16_kernel(const uint4* __restrict__ src, float* __restrict__ colsum, int64_t M, int C8,
                                                       int64_t rows_per_cta) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[32][65];
  const int cx = threadIdx.x & 7, ry = threadIdx.x >> 3;
  const int c8 = blockIdx.x * 8 + cx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = min(M, r0 + rows_per_cta);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto add = [&](const uint4& v) {
    const float2 a = D::unpack2(v.x), b = D::unpack2(v.y), c = D::unpack2(v.z), d = D::unpack2(v.w);
    acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
    acc[4] += c.x; acc[5] += c.y; acc[6] += d.x; acc[7] += d.y;
  };
  if (c8 < C8) {
    int64_t m = r0 + ry;
    for (; m + 96 < r1; m += 128) {
      const uint4 v0 = src[m * C8 + c8], v1 = src[(m + 32) * C8 + c8], v2 = src[(m + 64) * C8 + c8], v3 = src[(m + 96) * C8 + c8];
      add(v0); add(v1); add(v2); add(v3);
    }
    for (; m < r1; m += 32) add(src[m * C8 + c8]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[ry][8 * cx + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < 64 && blockIdx.x * 64 + threadIdx.x < 8 * C8) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) t += red[i][threadIdx.x];
    atomicAdd(colsum + blockIdx.x * 64 + threadIdx.x, t);
  }
}

// du = da * gelu'(u) over a [M, C] matrix AND colsum[C] += column sums of du (the fc1 bias gradient): a CTA owns 256 columns x a slab
// of rows, a warp reads 512 contiguous bytes of a row, 8 rows in flight per CTA; one atomic per column and CTA.
template <typename D>
__global__ void __launch_bounds__(256) gelu_bwd_colsum_kernel(const uint4* __restrict__ u, const uint4* __restrict__ da, uint4* __restrict__ out,
                                                              float* __restrict__ colsum, int64_t M, int C8, int64_t rows_per_cta) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8][257];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int c8 = blockIdx.x * 32 + cx;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_cta;
  const int64_t r1 = min(M, r0 + rows_per_cta);
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  auto one = [&](int64_t m) {
    const uint4 a = u[m * C8 + c8], g = da[m * C8 + c8];
    const uint32_t au[4] = {a.x, a.y, a.z, a.w}, gu[4] = {g.x, g.y, g.z, g.w};
    uint32_t r[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 xv = D::unpack2(au[k]), gv = D::unpack2(gu[k]);
      float gg0, gg1;
      gelu_grad2(xv.x, xv.y, gg0, gg1);
      r[k] = D::pack2(gv.x * gg0, gv.y * gg1);
      const float2 rv = D::unpack2(r[k]);     // sum what the weight-gradient GEMM will read
      acc[2 * k] += rv.x;
      acc[2 * k + 1] += rv.y;
    }
    out[m * C8 + c8] = make_uint4(r[0], r[1], r[2], r[3]);
  };
  if (c8 < C8) {
    int64_t m = r0 + ry;
    for (; m + 8 < r1; m += 16) {
      one(m);
      one(m + 8);
    }
    for (; m < r1; m += 8) one(m);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) red[ry][8 * cx + i] = acc[i];
  __syncthreads();
  const int col = blockIdx.x * 256 + threadIdx.x;
  if (col < 8 * C8) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += red[i][threadIdx.x];
    atomicAdd(colsum + col, t);
  }
}

// -------------------------------------------------------------------------------------------------- grouped row sums (pos-embed gradients)
// out[(m / div) % mod, c] += x[m, c]: Spatial_pos_embed (div 1, mod tokens), Temporal_pos_embed (div tokens, mod frames)
__global__ void group_rowsum_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n_outer, int C, int64_t div, int64_t mod) {
  pdl_launch_dependents();
  pdl_wait();
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
  pdl_launch_dependents();
  pdl_wait();
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
  pdl_launch_dependents();
  pdl_wait();
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
  pdl_launch_dependents();
  pdl_wait();
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
__device__ __forceinline__ void adam_one(float& pi, float g, float& mi, float& vi, float lr_bc1, float beta1, float beta2, float eps,
                                         float weight_decay, float bc2_sqrt, float grad_scale) {
  const float gi = fmaf(weight_decay, pi, g * grad_scale);
  mi = fmaf(beta1, mi, (1.f - beta1) * gi);
  vi = fmaf(beta2, vi, (1.f - beta2) * gi * gi);
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  pi = pi - lr_bc1 * (mi / denom);
}
// Four parameters per thread and pass (16-byte loads and stores: 7 memory instructions per 4 parameters instead of 28; the buffers are the
// flat parameter / gradient / moment buffers, 16-byte aligned), the bias corrections computed by ONE thread per CTA.
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                   int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, float bc1,
                                                   float bc2_sqrt, float grad_scale, const int64_t* __restrict__ step_dev,
                                                   const float* __restrict__ lr_dev) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float sh[2];
  if (threadIdx.x == 0) {
    if (lr_dev) lr = *lr_dev;      // the learning rate of a captured step lives on the device too (schedulers change it between replays)
    if (step_dev) {
      const double t = (double)(*step_dev + 1);
      bc1 = (float)(1.0 - pow((double)beta1, t));
      bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, t));
    }
    sh[0] = lr / bc1;
    sh[1] = bc2_sqrt;
  }
  __syncthreads();
  const float lr_bc1 = sh[0];
  bc2_sqrt = sh[1];
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
  const int64_t n4 = vec ? n / 4 : 0;
  for (int64_t q = tid; q < n4; q += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[q];
    const float4 gg = reinterpret_cast<const float4*>(g)[q];
    float4 mm = reinterpret_cast<float4*>(m)[q], vv = reinterpret_cast<float4*>(v)[q];
    adam_one(pp.x, gg.x, mm.x, vv.x, lr_bc1, beta1, beta2, eps, weight_decay, bc2_sqrt, grad_scale);
    adam_one(pp.y, gg.y, mm.y, vv.y, lr_bc1, beta1, beta2, eps, weight_decay, bc2_sqrt, grad_scale);
    adam_one(pp.z, gg.z, mm.z, vv.z, lr_bc1, beta1, beta2, eps, weight_decay, bc2_sqrt, grad_scale);
    adam_one(pp.w, gg.w, mm.w, vv.w, lr_bc1, beta1, beta2, eps, weight_decay, bc2_sqrt, grad_scale);
    reinterpret_cast<float4*>(m)[q] = mm;
    reinterpret_cast<float4*>(v)[q] = vv;
    reinterpret_cast<float4*>(p)[q] = pp;
  }
  for (int64_t i = n4 * 4 + tid; i < n; i += stride) {
    float pi = p[i], mi = m[i], vi = v[i];
    adam_one(pi, g[i], mi, vi, lr_bc1, beta1, beta2, eps, weight_decay, bc2_sqrt, grad_scale);
    m[i] = mi;
    v[i] = vi;
    p[i] = pi;
  }
}

__global__ void adam_bump_kernel(int64_t* step_dev) {
  pdl_launch_dependents();
  pdl_wait();
  *step_dev += 1;
}

int stream_grid(int64_t n, int per_cta) {
  int64_t ctas = (n + per_cta - 1) / per_cta;
  const int64_t cap = (int64_t)sm_count() * 8;
  return (int)(ctas < cap ? (ctas > 0 ? ctas : 1) : cap);
}

}  // namespace
}  // namespace mp

extern "C" {

int mp_layernorm_bwd(const float* x, const float* gamma, float eps, const void* dy, int dy_is_16bit, const float* dres, float* dx,
                     float* dgamma, float* dbeta, void* dx16, const float* rowscale, float* dx16_colsum, int64_t n_tokens, int C, int dtype,
                     mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(C == 512 || C == 128, MP_EUNSUPPORTED, "mp_layernorm_bwd: C=%d (built for 512 and 128)", C);
  MP_REQUIRE(x && dy && dx && n_tokens >= 0, MP_EINVAL, "mp_layernorm_bwd: bad arguments");
  MP_REQUIRE((dgamma == nullptr) == (dbeta == nullptr) && (gamma != nullptr || dgamma == nullptr), MP_EINVAL,
             "mp_layernorm_bwd: dgamma / dbeta come together and need gamma");
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_layernorm_bwd: unknown dtype %d", dtype);
  MP_REQUIRE(aligned16(x) && aligned16(dy) && aligned16(dx) && aligned16(dres) && aligned16(dx16), MP_EALIGN,
             "mp_layernorm_bwd: rows must be 16-byte aligned");
  if (n_tokens == 0) return MP_OK;
  MP_REQUIRE(dx16_colsum == nullptr || dx16 != nullptr, MP_EINVAL, "mp_layernorm_bwd: dx16_colsum needs dx16");
  // one CTA per SM (every CTA ends with up to 3 C atomics on the same words); few rows: one per warp
  int64_t ctas = (n_tokens + kLnBwdWarps - 1) / kLnBwdWarps;
  const int grid = (int)(ctas < sm_count() ? ctas : sm_count());
  const int smem = kLnBwdWarps * 3 * C * (int)sizeof(float);
  auto launch = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    launch_k(kernel, grid, kLnBwdWarps * 32, smem, (cudaStream_t)stream, x, gamma, eps, dy, dres, dx, dgamma, dbeta, (uint16_t*)dx16, rowscale, dx16_colsum,
             n_tokens);
  };
  const bool bf = dtype == MP_DTYPE_BF16;
  // C = 512: rows staged through per-warp shared-memory rings by bulk copies (MANIPOSE_LNBWD_RING=0: the register-prefetch kernel, A/B)
  // C = 512: rows staged through per-warp shared-memory rings by bulk copies (MANIPOSE_LNBWD_RING=0: the register-prefetch kernel, A/B).
  // (Two warps per row - half the instruction stream and registers per warp, 20 warps per SM, the row sums combined through shared memory
  // and a named barrier per pair - was slower: training step 7.32 against 7.20 ms.)
  static const bool ring = !(getenv("MANIPOSE_LNBWD_RING") && atoi(getenv("MANIPOSE_LNBWD_RING")) == 0);
  if (C == 512 && ring) {
    auto launch_ring = [&](auto kernel, int slot_bytes) {
      const int bytes = kLnBwdWarps * kLnRing * slot_bytes + kLnBwdWarps * kLnRing * 8;
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      launch_k(kernel, grid, kLnBwdWarps * 32, bytes, (cudaStream_t)stream, x, gamma, eps, dy, dres, dx, dgamma, dbeta, (uint16_t*)dx16, rowscale,
               dx16_colsum, n_tokens);
    };
    if (!dy_is_16bit) {
      if (bf) launch_ring(layernorm_bwd_ring_kernel<false, Bf16>, ln_ring_slot_bytes<false>());
      else launch_ring(layernorm_bwd_ring_kernel<false, Fp16>, ln_ring_slot_bytes<false>());
    } else {
      if (bf) launch_ring(layernorm_bwd_ring_kernel<true, Bf16>, ln_ring_slot_bytes<true>());
      else launch_ring(layernorm_bwd_ring_kernel<true, Fp16>, ln_ring_slot_bytes<true>());
    }
    return check_launch("layernorm_bwd_ring_kernel");
  }
  if (C == 512) {
    if (!dy_is_16bit) { if (bf) launch(layernorm_bwd_kernel<512, false, Bf16>); else launch(layernorm_bwd_kernel<512, false, Fp16>); }
    else if (bf) launch(layernorm_bwd_kernel<512, true, Bf16>);
    else launch(layernorm_bwd_kernel<512, true, Fp16>);
  } else {
    if (!dy_is_16bit) { if (bf) launch(layernorm_bwd_kernel<128, false, Bf16>); else launch(layernorm_bwd_kernel<128, false, Fp16>); }
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
  auto launch = [&](auto kernel) { launch_k(kernel, grid, 256, 0, (cudaStream_t)stream, (const uint4*)u, (const uint4*)da, (uint4*)out, n8); };
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

int mp_gelu_bwd_colsum(const void* u, const void* da, void* du, float* colsum, int64_t M, int64_t C, int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(u && da && du && colsum && M >= 0 && C >= 8 && C % 8 == 0, MP_EINVAL, "mp_gelu_bwd_colsum: bad arguments (C %% 8 == 0)");
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_gelu_bwd_colsum: unknown dtype %d", dtype);
  MP_REQUIRE(aligned16(u) && aligned16(da) && aligned16(du), MP_EALIGN, "mp_gelu_bwd_colsum: pointers must be 16-byte aligned");
  if (M == 0) return MP_OK;
  const int c8 = (int)(C / 8);
  const int gx = (c8 + 31) / 32;
  int64_t gy = (int64_t)sm_count() * 4 / gx + 1;
  if (gy > (M + 31) / 32) gy = (M + 31) / 32;
  const int64_t rows_per_cta = (M + gy - 1) / gy;
  gy = (M + rows_per_cta - 1) / rows_per_cta;
  if (dtype == MP_DTYPE_BF16)
    launch_k(gelu_bwd_colsum_kernel<Bf16>, dim3(gx, (unsigned)gy), 256, 0, (cudaStream_t)stream, (const uint4*)u, (const uint4*)da, (uint4*)du, colsum,
             M, c8, rows_per_cta);
  else
    launch_k(gelu_bwd_colsum_kernel<Fp16>, dim3(gx, (unsigned)gy), 256, 0, (cudaStream_t)stream, (const uint4*)u, (const uint4*)da, (uint4*)du, colsum,
             M, c8, rows_per_cta);
  return check_launch("gelu_bwd_colsum_kernel");
}

int mp_attention_bwd(const void* qkv, const void* o, const void* dout, void* dqkv, float* dqkv_colsum, int64_t n_clips, int64_t n_frames, int n_tok,
                     int C, int n_heads, int mode, int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(qkv && o && dout && dqkv && n_clips >= 0 && n_frames >= 1 && n_tok >= 1 && n_heads >= 1, MP_EINVAL, "mp_attention_bwd: bad arguments");
  MP_REQUIRE(mode == MP_ATTN_SPATIAL || mode == MP_ATTN_TEMPORAL, MP_EINVAL, "mp_attention_bwd: unknown mode %d", mode);
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_attention_bwd: unknown dtype %d", dtype);
  const int hd = C / n_heads;
  MP_REQUIRE(C % n_heads == 0 && (hd == 64 || hd == 16), MP_EUNSUPPORTED, "mp_attention_bwd: head_dim %d (built for 64 and 16)", hd);
  const int temporal = mode == MP_ATTN_TEMPORAL;
  const int64_t L = temporal ? n_frames : n_tok;
  MP_REQUIRE(L <= 256, MP_EUNSUPPORTED, "mp_attention_bwd: sequence length %lld > 256 (shared-memory resident sequence)", (long long)L);
  MP_REQUIRE(aligned16(qkv) && aligned16(o) && aligned16(dout) && aligned16(dqkv), MP_EALIGN, "mp_attention_bwd: pointers must be 16-byte aligned");
  const int64_t n_seq = temporal ? n_clips * n_tok : n_clips * n_frames;
  const int64_t n_items = n_seq * n_heads;
  MP_REQUIRE(n_items < ((int64_t)1 << 31), MP_EINVAL, "mp_attention_bwd: too many (sequence, head) items");
  if (n_items == 0) return MP_OK;
  // head_dim 64 and short sequences: the tcgen05 block-diagonal kernel (attention.cu); MANIPOSE_ATTN_BWD_MMA keeps the mma.sync one (A/B)
  static const bool legacy = getenv("MANIPOSE_ATTN_BWD_MMA") != nullptr;
  if (hd == 64 && !legacy && (temporal ? n_frames <= 128 : n_tok <= 32))
    return attention_bwd_tc(qkv, dout, dqkv, dqkv_colsum, n_clips, n_frames, n_tok, C, n_heads, temporal, dtype, (cudaStream_t)stream);
  const bool bf = dtype == MP_DTYPE_BF16;
  const int Lp = ((int)L + 31) & ~31;
  int n_warps = Lp / 16;
  if (n_warps > kAttnBwdMmaWarps) n_warps = kAttnBwdMmaWarps;
  auto launch_mma = [&](auto kernel, int HD) -> int {
    const int LD = HD * 2 + 16;
    const size_t smem = (size_t)5 * Lp * LD + (size_t)2 * Lp * sizeof(float) + (size_t)n_warps * 16 * LD;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    MP_REQUIRE(e == cudaSuccess, MP_ELAUNCH, "cudaFuncSetAttribute(attention_bwd_mma_kernel): %s", cudaGetErrorString(e));
    launch_k(kernel, (unsigned)n_items, n_warps * 32, smem, (cudaStream_t)stream, (const uint16_t*)qkv, (const uint16_t*)o, (const uint16_t*)dout,
                                                                           (uint16_t*)dqkv, (int)L, Lp, n_heads, C, n_tok, (int)n_frames, temporal);
    MP_CHECK(check_launch("attention_bwd_mma_kernel"));
    if (dqkv_colsum) return mp_colsum16(dqkv, dqkv_colsum, n_seq * L, 3 * (int64_t)C, dtype, stream);
    return MP_OK;
  };
  if (hd == 64) return bf ? launch_mma(attention_bwd_mma_kernel<64, Bf16>, 64) : launch_mma(attention_bwd_mma_kernel<64, Fp16>, 64);
  return bf ? launch_mma(attention_bwd_mma_kernel<16, Bf16>, 16) : launch_mma(attention_bwd_mma_kernel<16, Fp16>, 16);
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
    launch_k(transpose16_kernel<Bf16>, grid, 256, 0, (cudaStream_t)stream, (const uint16_t*)src, (uint16_t*)dst, colsum, M, C, Mpad);
  else
    launch_k(transpose16_kernel<Fp16>, grid, 256, 0, (cudaStream_t)stream, (const uint16_t*)src, (uint16_t*)dst, colsum, M, C, Mpad);
  return check_launch("transpose16_kernel");
}

int mp_refresh_shadows(const int64_t* table, int n_weights, int max_tiles, int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(table && n_weights >= 0 && max_tiles >= 1, MP_EINVAL, "mp_refresh_shadows: bad arguments");
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_refresh_shadows: unknown dtype %d", dtype);
  if (n_weights == 0) return MP_OK;
  const int gx = max_tiles < 64 ? max_tiles : 64;
  if (dtype == MP_DTYPE_BF16)
    launch_k(refresh_shadows_kernel<Bf16>, dim3(gx, n_weights), 256, 0, (cudaStream_t)stream, table);
  else
    launch_k(refresh_shadows_kernel<Fp16>, dim3(gx, n_weights), 256, 0, (cudaStream_t)stream, table);
  return check_launch("refresh_shadows_kernel");
}

int mp_colsum16(const void* src, float* colsum, int64_t M, int64_t C, int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(src && colsum && M >= 0 && C >= 8 && C % 8 == 0, MP_EINVAL, "mp_colsum16: bad arguments (C %% 8 == 0)");
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_colsum16: unknown dtype %d", dtype);
  MP_REQUIRE(aligned16(src), MP_EALIGN, "mp_colsum16: src must be 16-byte aligned");
  if (M == 0) return MP_OK;
  const int c8 = (int)(C / 8);
  const int gx = (c8 + 7) / 8;
  // about four CTAs per SM, at least 128 rows each (few atomics per column)
  int64_t gy = (int64_t)sm_count() * 4 / gx + 1;
  if (gy > (M + 127) / 128) gy = (M + 127) / 128;
  const int64_t rows_per_cta = (M + gy - 1) / gy;
  gy = (M + rows_per_cta - 1) / rows_per_cta;
  if (dtype == MP_DTYPE_BF16)
    launch_k(colsum16_kernel<Bf16>, dim3(gx, (unsigned)gy), 256, 0, (cudaStream_t)stream, (const uint4*)src, colsum, M, c8, rows_per_cta);
  else
    launch_k(colsum16_kernel<Fp16>, dim3(gx, (unsigned)gy), 256, 0, (cudaStream_t)stream, (const uint4*)src, colsum, M, c8, rows_per_cta);
  return check_launch("colsum16_kernel");
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
  launch_k(group_rowsum_kernel, dim3((unsigned)mod, (unsigned)splits), C, 0, (cudaStream_t)stream, x, out, n_outer, C, div, mod);
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
    case 2: launch_k(small_wgrad_kernel<2>, grid, 128, 0, s, dy, in, dW, db, n_rows, n_out); break;
    case 3: launch_k(small_wgrad_kernel<3>, grid, 128, 0, s, dy, in, dW, db, n_rows, n_out); break;
    case 34: launch_k(small_wgrad_kernel<34>, grid, 128, 0, s, dy, in, dW, db, n_rows, n_out); break;
    default: launch_k(small_wgrad_kernel<51>, grid, 128, 0, s, dy, in, dW, db, n_rows, n_out); break;
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
    launch_k(kernel, token_grid(n_tokens), kTokWarps * 32, 0, (cudaStream_t)stream, x, (const uint16_t*)y, s, out, n_tokens);
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
  auto launch = [&](auto kernel) { launch_k(kernel, token_grid(n_tokens), kTokWarps * 32, 0, (cudaStream_t)stream, g, s, (uint16_t*)out, n_tokens); };
  const bool bf = dtype == MP_DTYPE_BF16;
  if (C == 512) {
    if (bf) launch(cast_rowscale_kernel<512, Bf16>); else launch(cast_rowscale_kernel<512, Fp16>);
  } else {
    if (bf) launch(cast_rowscale_kernel<128, Bf16>); else launch(cast_rowscale_kernel<128, Fp16>);
  }
  return check_launch("cast_rowscale_kernel");
}

int mp_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr, float beta1, float beta2, float eps,
                 float weight_decay, int64_t step, int64_t* step_dev, const float* lr_dev, float grad_scale, mp_stream_t stream) {
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
  launch_k(adam_kernel, stream_grid((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream, param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay,
                                                                       bc1, bc2_sqrt, grad_scale, step_dev, lr_dev);
  MP_CHECK(check_launch("adam_kernel"));
  if (step_dev) {
    launch_k(adam_bump_kernel, 1, 1, 0, (cudaStream_t)stream, step_dev);
    return check_launch("adam_bump_kernel");
  }
  return MP_OK;
}

}  // extern "C"

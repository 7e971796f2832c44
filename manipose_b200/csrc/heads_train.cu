// Training path of the K hypothesis heads (MCLHead, rmcl_manifold_mix_ste.py:267-298; RMCLRotMixSTE.forward :251-262): the parameter
// folding of the forward and the unfolding of the backward as three small kernels around the tensor-core GEMMs.
//
// The heads share the normalised trunk output yhat (LN_k(y) = yhat * gamma_k + beta_k), so all K heads are ONE Linear with folded
// parameters  Wf[k (D+1) + d, :] = W_k[d, :] * gamma_k,  bf[k (D+1) + d] = W_k[d, :] . beta_k + b_k[d]  (rows past K (D+1) zero, n_pad a
// multiple of 128), followed by the J-term score dot product.  Reverse mode:
//   mp_heads_bwd_pack   dY [tokens, n_pad] (16-bit) from d rot / d logits, its column sums dbf, and the score-head gradients
//   mp_wgrad / mp_linear  dWf = dY^T yhat,  d yhat = dY Wf                                   (gemm.cu)
//   mp_heads_unfold     dW_k = dWf * gamma_k + dbf beta_k^T,  db_k = dbf,  dgamma_k = sum_d dWf * W_k,  dbeta_k = sum_d dbf W_k
// Head parameters and their fp32 gradient buffers are separate tensors: they come as device tables of pointers,
// int64 [6][K] = {norm.weight, norm.bias, prediction_head.weight, prediction_head.bias, score_head.weight, score_head.bias} x head.
#include <algorithm>

#include "common.cuh"

namespace mp {
namespace {

enum { kGamma = 0, kBeta = 1, kW = 2, kB = 3, kSw = 4, kSb = 5 };

__device__ __forceinline__ float* tab(const int64_t* table, int kind, int n_hyp, int k) {
  return reinterpret_cast<float*>(static_cast<uintptr_t>(table[kind * n_hyp + k]));
}

// one CTA (128 threads) per folded row
template <typename D>
__global__ void __launch_bounds__(128)
heads_fold_kernel(const int64_t* __restrict__ params, int n_hyp, int d1, int C, int n_pad, uint16_t* __restrict__ wf16, uint16_t* __restrict__ wt16,
                  float* __restrict__ bf, float* __restrict__ sw, float* __restrict__ sb) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[4];
  const int r = blockIdx.x;
  if (r >= n_hyp * d1) {
    for (int c = threadIdx.x; c < C; c += 128) {
      wf16[(size_t)r * C + c] = 0;
      if (wt16) wt16[(size_t)c * n_pad + r] = 0;
    }
    if (threadIdx.x == 0) bf[r] = 0.f;
    return;
  }
  const int k = r / d1, d = r - k * d1;
  const float* gamma = tab(params, kGamma, n_hyp, k);
  const float* beta = tab(params, kBeta, n_hyp, k);
  const float* w = tab(params, kW, n_hyp, k) + (size_t)d * C;
  float acc = 0.f;
  for (int c = threadIdx.x; c < C; c += 128) {
    const float wv = w[c];
    const uint32_t p = D::pack2(wv * gamma[c], 0.f);
    const uint16_t h = (uint16_t)(p & 0xffffu);
    wf16[(size_t)r * C + c] = h;
    if (wt16) wt16[(size_t)c * n_pad + r] = h;
    acc = fmaf(wv, beta[c], acc);
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) bf[r] = red[0] + red[1] + red[2] + red[3] + tab(params, kB, n_hyp, k)[d];
  if (d == 0) {   // the head's score Linear(J -> 1), stacked
    if (threadIdx.x < kJ) sw[k * kJ + threadIdx.x] = tab(params, kSw, n_hyp, k)[threadIdx.x];
    if (threadIdx.x == 0) sb[k] = tab(params, kSb, n_hyp, k)[0];
  }
}

// thread = folded column r, a CTA walks a slab of frames; joints unrolled so that the score-weight gradient has a register per joint
template <typename D>
__global__ void __launch_bounds__(128)
heads_bwd_pack_kernel(const float* __restrict__ d_rot, const float* __restrict__ d_logits, const float* __restrict__ y, const float* __restrict__ sw,
                      uint16_t* __restrict__ dy16, float* __restrict__ dbf, const int64_t* __restrict__ grads, int64_t n_clips, int n_frames, int n_hyp,
                      int out_dim, int n_pad, int y_ld, int64_t frames_per_cta) {
  pdl_launch_dependents();
  pdl_wait();
  const int d1 = out_dim + 1;
  const int64_t total_frames = n_clips * n_frames;
  const int64_t f0 = (int64_t)blockIdx.x * frames_per_cta, f1 = min(total_frames, f0 + frames_per_cta);
  for (int r = threadIdx.x; r < n_pad; r += 128) {
    const bool used = r < n_hyp * d1;
    const int k = used ? r / d1 : 0, d = used ? r - k * d1 : 0;
    const bool score = used && d == out_dim;
    float col = 0.f, gsb = 0.f, gsw[kJ];
    float swk[kJ];
#pragma unroll
    for (int j = 0; j < kJ; ++j) {
      gsw[j] = 0.f;
      swk[j] = score ? sw[k * kJ + j] : 0.f;
    }
    for (int64_t fr = f0; fr < f1; ++fr) {
      const int64_t b = fr / n_frames;
      const int64_t t = fr - b * n_frames;
      const int64_t hk = (b * n_hyp + k) * n_frames + t;          // (clip, head, frame)
      const float dl = score ? d_logits[hk] : 0.f;
      gsb += dl;
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        const int64_t m = fr * kJ + j;
        float v = 0.f;
        if (score) {
          v = dl * swk[j];
          gsw[j] = fmaf(dl, y[m * y_ld + r], gsw[j]);
        } else if (used) {
          v = d_rot[(hk * kJ + j) * out_dim + d];
        }
        const uint32_t p = D::pack2(v, 0.f);
        dy16[m * n_pad + r] = (uint16_t)(p & 0xffffu);
        col += D::unpack2(p).x;     // sum what the weight-gradient GEMM will read
      }
    }
    if (used) atomicAdd(dbf + r, col);
    if (score) {
      float* g_sw = tab(grads, kSw, n_hyp, k);
#pragma unroll
      for (int j = 0; j < kJ; ++j) atomicAdd(g_sw + j, gsw[j]);
      atomicAdd(tab(grads, kSb, n_hyp, k), gsb);
    }
  }
}

// grid (K, C / 128), thread = channel c
__global__ void __launch_bounds__(128)
heads_unfold_kernel(const int64_t* __restrict__ params, const int64_t* __restrict__ grads, const float* __restrict__ dwf, const float* __restrict__ dbf,
                    int n_hyp, int d1, int C) {
  pdl_launch_dependents();
  pdl_wait();
  const int k = blockIdx.x, c = blockIdx.y * 128 + threadIdx.x;
  if (c < C) {
    const float gam = tab(params, kGamma, n_hyp, k)[c], bet = tab(params, kBeta, n_hyp, k)[c];
    const float* w = tab(params, kW, n_hyp, k);
    float* gw = tab(grads, kW, n_hyp, k);
    float dgam = 0.f, dbet = 0.f;
    for (int d = 0; d < d1; ++d) {
      const int r = k * d1 + d;
      const float g = dwf[(size_t)r * C + c], gb = dbf[r], wv = w[(size_t)d * C + c];
      gw[(size_t)d * C + c] += fmaf(g, gam, gb * bet);   // Wf = W * gamma and bf = W beta + b both depend on W
      dgam = fmaf(g, wv, dgam);
      dbet = fmaf(gb, wv, dbet);
    }
    tab(grads, kGamma, n_hyp, k)[c] += dgam;
    tab(grads, kBeta, n_hyp, k)[c] += dbet;
  }
  if (blockIdx.y == 0 && threadIdx.x < d1) tab(grads, kB, n_hyp, k)[threadIdx.x] += dbf[k * d1 + threadIdx.x];
}

}  // namespace
}  // namespace mp

extern "C" {

int mp_heads_fold(const int64_t* params, int n_hyp, int out_dim, int C, int n_pad, void* wf16, void* wt16, float* bf, float* score_w,
                  float* score_b, int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(params && wf16 && bf && score_w && score_b, MP_EINVAL, "mp_heads_fold: null pointer");
  MP_REQUIRE(n_hyp >= 1 && out_dim >= 1 && C >= 1 && n_pad >= n_hyp * (out_dim + 1), MP_EINVAL, "mp_heads_fold: bad sizes");
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_heads_fold: unknown dtype %d", dtype);
  auto launch = [&](auto kernel) {
    launch_k(kernel, n_pad, 128, 0, (cudaStream_t)stream, params, n_hyp, out_dim + 1, C, n_pad, (uint16_t*)wf16, (uint16_t*)wt16, bf, score_w, score_b);
  };
  if (dtype == MP_DTYPE_BF16) launch(heads_fold_kernel<Bf16>); else launch(heads_fold_kernel<Fp16>);
  return check_launch("heads_fold_kernel");
}

int mp_heads_bwd_pack(const float* d_rot, const float* d_logits, const float* y, const float* score_w, void* dy16, float* dbf,
                      const int64_t* grads, int64_t n_clips, int64_t n_frames, int n_hyp, int out_dim, int n_pad, int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(d_rot && d_logits && y && score_w && dy16 && dbf && grads, MP_EINVAL, "mp_heads_bwd_pack: null pointer");
  MP_REQUIRE(n_clips >= 0 && n_frames >= 1 && n_hyp >= 1 && out_dim >= 1 && n_pad >= n_hyp * (out_dim + 1), MP_EINVAL, "mp_heads_bwd_pack: bad sizes");
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_heads_bwd_pack: unknown dtype %d", dtype);
  const int64_t frames = n_clips * n_frames;
  if (frames == 0) return MP_OK;
  int64_t ctas = std::min<int64_t>(frames, (int64_t)sm_count() * 4);
  const int64_t per = (frames + ctas - 1) / ctas;
  ctas = (frames + per - 1) / per;
  auto launch = [&](auto kernel) {
    launch_k(kernel, (unsigned)ctas, 128, 0, (cudaStream_t)stream, d_rot, d_logits, y, score_w, (uint16_t*)dy16, dbf, grads, n_clips, (int)n_frames, n_hyp,
             out_dim, n_pad, heads_ws_ld(n_hyp, out_dim, 1, n_pad), per);
  };
  if (dtype == MP_DTYPE_BF16) launch(heads_bwd_pack_kernel<Bf16>); else launch(heads_bwd_pack_kernel<Fp16>);
  return check_launch("heads_bwd_pack_kernel");
}

int mp_heads_unfold(const int64_t* params, const int64_t* grads, const float* dwf, const float* dbf, int n_hyp, int out_dim, int C, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(params && grads && dwf && dbf && n_hyp >= 1 && out_dim >= 1 && out_dim + 1 <= 128 && C >= 1, MP_EINVAL, "mp_heads_unfold: bad arguments");
  launch_k(heads_unfold_kernel, dim3((unsigned)n_hyp, (unsigned)((C + 127) / 128)), 128, 0, (cudaStream_t)stream, params, grads, dwf, dbf, n_hyp,
           out_dim + 1, C);
  return check_launch("heads_unfold_kernel");
}

}  // extern "C"

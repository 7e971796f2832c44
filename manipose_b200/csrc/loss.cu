// Winner-takes-all loss with scoring (rMCL), velocity + smoothness terms, hypothesis aggregation and MPJPE:
// warp-tile reduction kernels, forward and backward.
//
// Replaces (reference, paths under hpe/mh_so3_hpe/):
//   metrics/losses.py:14-72,104-123   weighted_mpjpe_loss / weighted_mse_loss / _l2_loss_per_hyp
//   metrics/losses.py:126-170         wta_l2_loss_and_activate_head / wta_with_scoring_loss
//   metrics/losses.py:75-101          mean_velocity_error
//   metrics/regularizations.py:160-174 smoothness_regularization
//   architectures/rmcl_manifold_mix_ste.py:121-185 poses_from_hyp_idx / aggregate
//   metrics/mean_joint_errors.py:31-36 mpjpe_error
//
// Layout: hyp [B,K,T,17,3], y [B,T,17,3].  A warp owns a tile of consecutive frames of one clip; for each hypothesis one lane pulls the
// tile (+ halo frames, 204 B each, 16-byte aligned-down start) into shared memory with ONE bulk async copy (mbarrier completion), the
// copy of hypothesis k+1 in flight while lane t reduces frame t of hypothesis k reading with stride 51 floats (bank-conflict free).
// Per-hypothesis errors reproduce the IEEE operation sequence of PyTorch's CPU kernels (FMA chain in the 3-norm, the 8-lane sum order of
// its reduction, the divide by 17) so winner indices are bit-identical to the oracle's on identical inputs; the correctly rounded square
// roots and divisions are the branch-free sequences of ieee.cuh, and a frame with an operand outside their range (an exact zero
// distance, denormals) is redone with the IEEE intrinsics.  Scalar terms are reduced warp -> per-warp double partial -> one fixed-order
// finalize block (deterministic, no atomics).
#include "common.cuh"
#include "ieee.cuh"
#include "ptx.cuh"

namespace mp {
namespace {

constexpr int kF = kJ * 3;                 // 51 floats per frame
constexpr int kTileFrames = 32;            // forward: frames per warp tile
constexpr int kBwdFrames = 31;             // backward: output frames per warp tile (lane L owns frame t0 - 1 + L and the pair (t, t+1))
constexpr int kStageFloats = 33 * kF + 5;  // 33 frames (tile + halo) + alignment shift, rounded to a multiple of 4 floats (16 B): 1688
static_assert(kStageFloats % 4 == 0, "stage buffers must stay 16-byte aligned for the bulk copies");
constexpr int kStageBytes = kStageFloats * 4;
// One CTA of independent warps per SM, shared memory split into per-warp stage buffers.  Forward: the ground-truth frame lives in
// registers, two hypothesis buffers per warp; backward: y + two hypothesis buffers (its registers hold the 51 gradients).
constexpr int kFwdWarps = 12, kFwdBufs = 2;
constexpr int kBwdWarps = 11, kBwdBufs = 3;
constexpr int kMaxPartialWarps = 148 * 2 * 16;   // >= SMs x kFwdWarps
constexpr int kMaxHyp = 32;
constexpr int kScorePrefetch = 5;          // scores of the first hypotheses (K = 5 by default) are requested ahead of their use

__constant__ float c_ones17[kJ] = {1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1};

// One-MUFU square root / reciprocal square root (relative error ~2^-22) for the terms that are means or gradients; the per-hypothesis
// WTA error keeps correctly rounded IEEE operations (winner indices must match the reference bit for bit).
__device__ __forceinline__ float sqrt_approx(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rsqrt_approx(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Brings floats [g0, g0+n) of `base` (16-byte aligned, `total` floats long) into sbuf so that element g0+i lands at sbuf[shift + i];
// returns shift in [0,3].  Whole 16-byte chunks leave as one bulk copy issued by lane 0; only a tile that touches the last, partial
// chunk of the array is copied by the lanes.  Either way `bar` (count 1) completes once the data is visible.
__device__ __forceinline__ int stage_tile(float* sbuf, uint64_t* bar, const float* __restrict__ base, size_t g0, uint32_t n, size_t total,
                                          int lane) {
  const size_t a0 = g0 & ~(size_t)3;
  const int shift = (int)(g0 - a0);
  const uint32_t nfl = (shift + n + 3) & ~3u;
  if (a0 + nfl <= total) {
    if (lane == 0) {
      ptx::mbar_expect_tx(bar, nfl * 4);
      ptx::bulk_g2s(sbuf, base + a0, nfl * 4, bar);
    }
  } else {
    for (uint32_t i = lane; i < shift + n; i += 32)
      if (a0 + i < total) sbuf[i] = base[a0 + i];
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(bar);
  }
  return shift;
}

// sum of 17 values in the order of PyTorch's CPU reduction over a contiguous inner dim of size 17:
// 8-lane vector partials p_k = v_k + v_{k+8}, then v_16 + p_0 + p_1 + ... + p_7, left to right.
__device__ __forceinline__ float sum17_torch_order(const float (&v)[kJ]) {
  float acc = v[16];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc = __fadd_rn(acc, __fadd_rn(v[k], v[k + 8]));
  return acc;
}

// e = mean_j w_j * ||h_j - y_j||  (unsquared) or mean_j mean_c w_j (h - y)^2 (squared), bit-exact vs torch CPU: the IEEE intrinsics,
// out of line (the cold path of frame_error_core)
template <bool kSquared>
__device__ __noinline__ float frame_error_ieee(const float* __restrict__ h, const float* __restrict__ yy, const float* __restrict__ weights) {
  float v[kJ];
#pragma unroll
  for (int j = 0; j < kJ; ++j) {
    const float wj = weights ? weights[j] : 1.0f;
    const float d0 = __fsub_rn(h[j * 3 + 0], yy[j * 3 + 0]);
    const float d1 = __fsub_rn(h[j * 3 + 1], yy[j * 3 + 1]);
    const float d2 = __fsub_rn(h[j * 3 + 2], yy[j * 3 + 2]);
    if (kSquared) {
      const float q0 = __fmul_rn(wj, __fmul_rn(d0, d0));
      const float q1 = __fmul_rn(wj, __fmul_rn(d1, d1));
      const float q2 = __fmul_rn(wj, __fmul_rn(d2, d2));
      v[j] = __fdiv_rn(__fadd_rn(__fadd_rn(q0, q1), q2), 3.0f);
    } else {
      const float s = __fmaf_rn(d2, d2, __fmaf_rn(d1, d1, __fmul_rn(d0, d0)));
      v[j] = __fmul_rn(wj, __fsqrt_rn(s));
    }
  }
  return __fdiv_rn(sum17_torch_order(v), 17.0f);
}

struct LossDims {
  uint32_t B, K, T;
};

template <int kBufs>
struct WarpStage {
  float* buf[kBufs];
  uint64_t* bar[kBufs];
  uint32_t phase[kBufs];
  __device__ __forceinline__ void init(uint8_t* smem_raw, int warp, int lane, int n_warps) {
    float* base = reinterpret_cast<float*>(smem_raw + (size_t)warp * kBufs * kStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + (size_t)n_warps * kBufs * kStageBytes) + warp * kBufs;
#pragma unroll
    for (int i = 0; i < kBufs; ++i) {
      buf[i] = base + i * kStageFloats;
      bar[i] = bars + i;
      phase[i] = 0;
      if (lane == 0) ptx::mbar_init(bar[i], 1);
    }
    if (lane == 0) ptx::fence_mbar_init();
    __syncwarp();
  }
  __device__ __forceinline__ void wait(int i) {
    ptx::mbar_wait(bar[i], phase[i]);
    phase[i] ^= 1;
  }
};
constexpr size_t loss_smem_bytes(int warps, int bufs) { return (size_t)warps * bufs * (kStageBytes + sizeof(uint64_t)); }

// ------------------------------------------------------------------------------------------------ forward
template <bool kSquared, bool kAllTerms>
__global__ void __launch_bounds__(kFwdWarps * 32, 1)
loss_fwd_kernel(const float* __restrict__ hyp, const float* __restrict__ scores, const float* __restrict__ y,
                const float* __restrict__ weights, LossDims d, float* __restrict__ wta_val, int64_t* __restrict__ wta_idx,
                float* __restrict__ per_hyp, double* __restrict__ partials) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpStage<kFwdBufs> st;
  st.init(smem_raw, warp, lane, kFwdWarps);

  float w[kJ];
#pragma unroll
  for (int j = 0; j < kJ; ++j) w[j] = weights ? weights[j] : c_ones17[j];
  const float r17 = ieee::rcp_refined(17.0f), r3 = ieee::rcp_refined(3.0f);
  const uint64_t r3_2 = pack_f32x2(r3, r3), kNeg3x2 = pack_f32x2(-3.0f, -3.0f), kNeg1x2 = pack_f32x2(-1.0f, -1.0f), kHalfx2 = pack_f32x2(0.5f, 0.5f);
  auto sub_f32x2 = [&](uint64_t a, uint64_t b) { return fma_f32x2(b, kNeg1x2, a); };   // a - b, one rounding per lane

  const uint32_t tiles_per_clip = (d.T + kTileFrames - 1) / kTileFrames;
  const uint32_t n_items = d.B * tiles_per_clip;
  const size_t y_total = (size_t)d.B * d.T * kF, h_total = y_total * d.K;
  double acc_wta = 0, acc_bce = 0, acc_vel = 0, acc_sm = 0;

  for (uint32_t item = warp * gridDim.x + blockIdx.x; item < n_items; item += gridDim.x * kFwdWarps) {
    const uint32_t b = item / tiles_per_clip, t0 = (item - b * tiles_per_clip) * kTileFrames;
    const uint32_t nf = min((uint32_t)kTileFrames, d.T - t0);
    const uint32_t nload = kAllTerms ? min((uint32_t)kTileFrames + 1, d.T - t0) : nf;
    const bool valid = lane < nf;
    const bool has_next = kAllTerms && (t0 + lane + 1 < d.T);

    __syncwarp();   // every lane is done with the previous item's buffers
    // ground truth: through buffer 1 into registers (frame t and, for the pair terms, its difference to frame t+1)
    const int sy = stage_tile(st.buf[1], st.bar[1], y, ((size_t)b * d.T + t0) * kF, nload * kF, y_total, lane);
    int sh_cur = stage_tile(st.buf[0], st.bar[0], hyp, (((size_t)b * d.K) * d.T + t0) * kF, nload * kF, h_total, lane);
    st.wait(1);
    // ground truth as PAIRS of joints (2 jp, 2 jp + 1), component by component, for the packed fp32 instructions; joint 16 alone
    uint64_t Y[8][3], DY[kAllTerms ? 8 : 1][3];
    float y16[3], dy16[3] = {0.f, 0.f, 0.f};
    {
      const float* yp = st.buf[1] + sy + lane * kF;
#pragma unroll
      for (int jp = 0; jp < 8; ++jp)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float ya = yp[6 * jp + c], yb = yp[6 * jp + 3 + c];
          Y[jp][c] = pack_f32x2(ya, yb);
          if (kAllTerms) DY[jp][c] = pack_f32x2(yp[kF + 6 * jp + c] - ya, yp[kF + 6 * jp + 3 + c] - yb);
        }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        y16[c] = yp[48 + c];
        if (kAllTerms) dy16[c] = yp[kF + 48 + c] - y16[c];
      }
    }
    __syncwarp();   // buffer 1 is free for hypothesis 1
    float best = INFINITY, vel = 0.f, sm = 0.f;
    int best_k = 0;
    // the frame's scores are only needed after the last hypothesis: requested now, so that their latency hides under the loop
    float sc_pre[kScorePrefetch];
    if (kAllTerms && scores) {
#pragma unroll
      for (int k = 0; k < kScorePrefetch; ++k)
        sc_pre[k] = (valid && (uint32_t)k < d.K) ? scores[((size_t)b * d.K + k) * d.T + t0 + lane] : 0.5f;
    }
    for (uint32_t k = 0; k < d.K; ++k) {
      const int cur = k & 1;
      int sh_next = 0;
      if (k + 1 < d.K)   // buffer cur^1 was read for hypothesis k-1: every lane passed the __syncwarp at the end of that iteration
        sh_next = stage_tile(st.buf[cur ^ 1], st.bar[cur ^ 1], hyp, (((size_t)b * d.K + k + 1) * d.T + t0) * kF, nload * kF, h_total, lane);
      st.wait(cur);
      const float* hp = st.buf[cur] + sh_cur + lane * kF;

      // ---- frame t of hypothesis k: WTA error (exact) + velocity / smoothness of the pair (t, t+1).  Two joints per instruction
      // (fma / mul / add.rn.f32x2: every lane of a packed instruction is the IEEE operation of the scalar form, same order).
      float v[kJ];
      float smin = INFINITY, smax = 0.f, vk = 0.f, sk = 0.f;
      uint64_t skp = 0;
#pragma unroll
      for (int jp = 0; jp < 8; ++jp) {
        const int j = 2 * jp;
        const uint64_t W2 = pack_f32x2(w[j], w[j + 1]);
        const uint64_t H0 = pack_f32x2(hp[3 * j + 0], hp[3 * j + 3]), H1 = pack_f32x2(hp[3 * j + 1], hp[3 * j + 4]),
                       H2 = pack_f32x2(hp[3 * j + 2], hp[3 * j + 5]);
        const uint64_t D0 = sub_f32x2(H0, Y[jp][0]), D1 = sub_f32x2(H1, Y[jp][1]), D2 = sub_f32x2(H2, Y[jp][2]);
        uint64_t S, V;
        if (kSquared) {
          const uint64_t Q0 = mul_f32x2(W2, mul_f32x2(D0, D0)), Q1 = mul_f32x2(W2, mul_f32x2(D1, D1)), Q2 = mul_f32x2(W2, mul_f32x2(D2, D2));
          S = add_f32x2(add_f32x2(Q0, Q1), Q2);
          const uint64_t q0 = mul_f32x2(S, r3_2);                                   // ieee::div_rn_core, packed
          V = fma_f32x2(r3_2, fma_f32x2(kNeg3x2, q0, S), q0);
        } else {
          S = fma_f32x2(D2, D2, fma_f32x2(D1, D1, mul_f32x2(D0, D0)));
          float sa, sb;
          unpack_f32x2(S, sa, sb);
          const uint64_t RS = pack_f32x2(ieee::mufu_rsq(sa), ieee::mufu_rsq(sb));   // ieee::sqrt_rn_core, packed
          const uint64_t G = mul_f32x2(S, RS), Hh = mul_f32x2(RS, kHalfx2);
          const uint64_t R = fma_f32x2(mul_f32x2(G, kNeg1x2), G, S);
          V = mul_f32x2(W2, fma_f32x2(R, Hh, G));
        }
        float sa, sb;
        unpack_f32x2(S, sa, sb);
        smin = fminf(smin, fminf(sa, sb));
        smax = fmaxf(smax, fmaxf(sa, sb));
        unpack_f32x2(V, v[j], v[j + 1]);
        if (kAllTerms) {
          const uint64_t A0 = sub_f32x2(pack_f32x2(hp[kF + 3 * j + 0], hp[kF + 3 * j + 3]), H0),
                         A1 = sub_f32x2(pack_f32x2(hp[kF + 3 * j + 1], hp[kF + 3 * j + 4]), H1),
                         A2 = sub_f32x2(pack_f32x2(hp[kF + 3 * j + 2], hp[kF + 3 * j + 5]), H2);   // hypothesis velocity
          const uint64_t E0 = sub_f32x2(A0, DY[jp][0]), E1 = sub_f32x2(A1, DY[jp][1]), E2 = sub_f32x2(A2, DY[jp][2]);
          float qa, qb;
          unpack_f32x2(fma_f32x2(E2, E2, fma_f32x2(E1, E1, mul_f32x2(E0, E0))), qa, qb);
          vk += kSquared ? qa + qb : sqrt_approx(qa) + sqrt_approx(qb);
          skp = fma_f32x2(W2, fma_f32x2(A2, A2, fma_f32x2(A1, A1, mul_f32x2(A0, A0))), skp);
        }
      }
      {   // joint 16
        constexpr int j = 16;
        const float h0 = hp[j * 3 + 0], h1 = hp[j * 3 + 1], h2 = hp[j * 3 + 2];
        const float d0 = __fsub_rn(h0, y16[0]), d1 = __fsub_rn(h1, y16[1]), d2 = __fsub_rn(h2, y16[2]);
        float s;
        if (kSquared) {
          const float q0 = __fmul_rn(w[j], __fmul_rn(d0, d0));
          const float q1 = __fmul_rn(w[j], __fmul_rn(d1, d1));
          const float q2 = __fmul_rn(w[j], __fmul_rn(d2, d2));
          s = __fadd_rn(__fadd_rn(q0, q1), q2);
          v[j] = ieee::div_rn_core(s, 3.0f, r3);
        } else {
          s = __fmaf_rn(d2, d2, __fmaf_rn(d1, d1, __fmul_rn(d0, d0)));
          v[j] = __fmul_rn(w[j], ieee::sqrt_rn_core(s));
        }
        smin = fminf(smin, s);
        smax = fmaxf(smax, s);
        if (kAllTerms) {
          const float a0 = hp[kF + j * 3 + 0] - h0, a1 = hp[kF + j * 3 + 1] - h1, a2 = hp[kF + j * 3 + 2] - h2;
          const float e0 = a0 - dy16[0], e1 = a1 - dy16[1], e2 = a2 - dy16[2];
          const float q = fmaf(e2, e2, fmaf(e1, e1, e0 * e0));
          vk += kSquared ? q : sqrt_approx(q);
          sk = fmaf(w[j], fmaf(a2, a2, fmaf(a1, a1, a0 * a0)), sum_f32x2(skp));
        }
      }
      const float tot = sum17_torch_order(v);
      float e = ieee::div_rn_core(tot, 17.0f, r17);
      // every square-root / division operand in [2^-100, 2^100), or the frame is redone with the IEEE intrinsics
      const bool in_range = __float_as_uint(smin) >= 0x0d800000u && __float_as_uint(smax) < 0x71800000u && ieee::mag_in_range(tot);
      if (valid && !in_range) e = frame_error_ieee<kSquared>(hp, y + ((size_t)b * d.T + t0 + lane) * kF, weights);
      if (valid) {
        if (per_hyp) per_hyp[((size_t)b * d.K + k) * d.T + t0 + lane] = e;
        if (k == 0 || e < best) {  // torch.min(dim=1): lowest index on ties
          best = e;
          best_k = (int)k;
        }
      }
      if (kAllTerms) {
        vel += (valid && has_next) ? vk : 0.f;
        sm += (valid && has_next) ? sk : 0.f;
      }
      sh_cur = sh_next;
      __syncwarp();   // buf[cur] is free for hypothesis k+2
    }
    float bce = 0.f;
    if (valid) {
      wta_val[(size_t)b * d.T + t0 + lane] = best;
      wta_idx[(size_t)b * d.T + t0 + lane] = best_k;
      if (kAllTerms && scores) {
#pragma unroll
        for (int k = 0; k < kScorePrefetch; ++k)
          if ((uint32_t)k < d.K) bce -= (k == best_k) ? fmaxf(logf(sc_pre[k]), -100.f) : fmaxf(logf(1.f - sc_pre[k]), -100.f);
        for (uint32_t k = kScorePrefetch; k < d.K; ++k) {
          const float s = scores[((size_t)b * d.K + k) * d.T + t0 + lane];
          // F.binary_cross_entropy: log terms clamped at -100
          bce -= ((int)k == best_k) ? fmaxf(logf(s), -100.f) : fmaxf(logf(1.f - s), -100.f);
        }
      }
    } else {
      best = 0.f;
    }
    if (kAllTerms) {
      const float r0 = warp_sum(best), r1 = warp_sum(bce), r2 = warp_sum(vel), r3s = warp_sum(sm);
      acc_wta += r0;
      acc_bce += r1;
      acc_vel += r2;
      acc_sm += r3s;
    }
  }
  if (kAllTerms && lane == 0) {
    double* p = partials + (size_t)(blockIdx.x * kFwdWarps + warp) * 4;
    p[0] = acc_wta;
    p[1] = acc_bce;
    p[2] = acc_vel;
    p[3] = acc_sm;
  }
}

__global__ void loss_finalize_kernel(const double* __restrict__ partials, int n_partials, LossDims d, int squared, float beta,
                                     float vel_w, float smooth_w, float* __restrict__ terms) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[4][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double a[4] = {0, 0, 0, 0};
  for (int i = threadIdx.x; i < n_partials; i += blockDim.x)
    for (int q = 0; q < 4; ++q) a[q] += partials[(size_t)i * 4 + q];
  for (int q = 0; q < 4; ++q) {
    for (int o = 16; o > 0; o >>= 1) a[q] += __shfl_xor_sync(0xffffffffu, a[q], o);
    if (lane == 0) red[q][warp] = a[q];
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double s[4] = {0, 0, 0, 0};
    for (int q = 0; q < 4; ++q)
      for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s[q] += red[q][i];
    const double bt = (double)d.B * d.T, bkt = bt * d.K, pairs = (double)d.B * d.K * (d.T - 1.0) * kJ;
    const float wta = (float)(s[0] / bt);
    const float bce = (float)(s[1] / bkt);
    const float vel = (float)(s[2] / (squared ? pairs * 3.0 : pairs));
    const float smo = (float)(s[3] / (pairs * 3.0));
    terms[MP_TERM_WTA] = wta;
    terms[MP_TERM_BCE] = bce;
    terms[MP_TERM_VEL] = vel;
    terms[MP_TERM_SMOOTH] = smo;
    // make_loss / compute_and_acc_loss (hpe/main_h36m_lifting.py:101-209): terms with weight <= 0 are not added
    float total = wta;
    if (beta != 0.f) total += beta * bce;
    if (vel_w > 0.f) total += vel_w * vel;
    if (smooth_w > 0.f) total += smooth_w * smo;
    terms[MP_TERM_TOTAL] = total;
    terms[5] = terms[6] = terms[7] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ backward
// Lane L of a tile owns frame f = t0 - 1 + L and the frame pair (f, f+1): it computes the pair's velocity / smoothness "flux"
// F = d(terms)/d(h[f+1]) = -d(terms)/d(h[f]) ONCE, and the gradient of frame f is  wta part - F(f) + F(f-1), the second flux coming
// from lane L-1 by shuffle.  Lane 0 only supplies the flux into the tile's first frame, so a tile writes 31 frames.
template <bool kSquared>
__global__ void __launch_bounds__(kBwdWarps * 32, 1)
loss_bwd_kernel(const float* __restrict__ hyp, const float* __restrict__ scores, const float* __restrict__ y,
                const float* __restrict__ weights, const int64_t* __restrict__ wta_idx, LossDims d, float beta, float vel_w,
                float smooth_w, const float* __restrict__ grad_terms, const float* __restrict__ grad_wta_val,
                float* __restrict__ grad_hyp, float* __restrict__ grad_scores) {
  pdl_launch_dependents();
  pdl_wait();
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  WarpStage<kBwdBufs> st;   // buf[0]: y, buf[1], buf[2]: hypotheses
  st.init(smem_raw, warp, lane, kBwdWarps);

  float w[kJ];
#pragma unroll
  for (int j = 0; j < kJ; ++j) w[j] = weights ? weights[j] : c_ones17[j];

  // upstream gradients w.r.t. terms[WTA, BCE, VEL, SMOOTH, TOTAL]; TOTAL expands with the weights of mp_loss_fwd
  const float gtot = grad_terms[MP_TERM_TOTAL];
  const float g_wta = grad_terms[MP_TERM_WTA] + gtot;
  const float g_bce = grad_terms[MP_TERM_BCE] + (beta != 0.f ? beta * gtot : 0.f);
  const float g_vel = grad_terms[MP_TERM_VEL] + (vel_w > 0.f ? vel_w * gtot : 0.f);
  const float g_sm = grad_terms[MP_TERM_SMOOTH] + (smooth_w > 0.f ? smooth_w * gtot : 0.f);
  const double bt = (double)d.B * d.T, pairs = (double)d.B * d.K * (d.T - 1.0) * kJ;
  const float wta_mean = (float)(g_wta / bt);                                        // d mean / d wta_val[b,t]
  const float wta_scale = kSquared ? (float)(2.0 / (kJ * 3.0)) : (float)(1.0 / kJ);  // d wta_val / d (w_j |d_j|) resp. (w_j d^2)
  const float c_vel = kSquared ? (float)(g_vel * 2.0 / (pairs * 3.0)) : (float)(g_vel / pairs);
  const float c_sm = (float)(g_sm * 2.0 / (pairs * 3.0));
  const float c_bce = (float)(g_bce / (bt * d.K));

  const uint32_t tiles_per_clip = (d.T + kBwdFrames - 1) / kBwdFrames;
  const uint32_t n_items = d.B * tiles_per_clip;
  const size_t y_total = (size_t)d.B * d.T * kF, h_total = y_total * d.K;

  for (uint32_t item = warp * gridDim.x + blockIdx.x; item < n_items; item += gridDim.x * kBwdWarps) {
    const uint32_t b = item / tiles_per_clip, t0 = (item - b * tiles_per_clip) * kBwdFrames;
    const uint32_t nf = min((uint32_t)kBwdFrames, d.T - t0);            // frames written: t0 .. t0+nf-1 (lanes 1 .. nf)
    const uint32_t lo = t0 > 0 ? t0 - 1 : 0, hi = min(d.T, t0 + nf + 1);   // frames staged
    const int f = (int)t0 - 1 + lane;                                   // this lane's frame (lane 0: the frame before the tile)
    const bool in_clip = f >= 0 && f < (int)d.T;
    const bool writes = lane >= 1 && lane <= (int)nf;
    const bool has_pair = in_clip && (f + 1 < (int)d.T) && lane <= (int)nf;   // pair (f, f+1) touches a frame of this tile
    const int row = in_clip ? f - (int)lo : 0;                          // frame f sits at row `row` of the staged tile
    const int64_t kstar = writes ? wta_idx[(size_t)b * d.T + f] : -1;
    const float c_wta = writes ? (wta_mean + (grad_wta_val ? grad_wta_val[(size_t)b * d.T + f] : 0.f)) * wta_scale : 0.f;

    __syncwarp();
    const int sy = stage_tile(st.buf[0], st.bar[0], y, ((size_t)b * d.T + lo) * kF, (hi - lo) * kF, y_total, lane);
    int sh_cur = stage_tile(st.buf[1], st.bar[1], hyp, (((size_t)b * d.K) * d.T + lo) * kF, (hi - lo) * kF, h_total, lane);
    st.wait(0);
    const float* yc = st.buf[0] + sy + row * kF;
    for (uint32_t k = 0; k < d.K; ++k) {
      const int cur = 1 + (k & 1), nxt = 3 - cur;
      int sh_next = 0;
      if (k + 1 < d.K)
        sh_next = stage_tile(st.buf[nxt], st.bar[nxt], hyp, (((size_t)b * d.K + k + 1) * d.T + lo) * kF, (hi - lo) * kF, h_total, lane);
      const size_t si = ((size_t)b * d.K + k) * d.T + (size_t)(writes ? f : 0);
      const float sck = (writes && grad_scores) ? scores[si] : 0.5f;     // requested before the wait: its latency hides under the tile
      st.wait(cur);
      const float* hc = st.buf[cur] + sh_cur + row * kF;
      const bool winner = writes && (int64_t)k == kstar;

      float g[kF];
#pragma unroll
      for (int j = 0; j < kJ; ++j) {
        const float h0 = hc[j * 3 + 0], h1 = hc[j * 3 + 1], h2 = hc[j * 3 + 2];
        const float y0 = yc[j * 3 + 0], y1 = yc[j * 3 + 1], y2 = yc[j * 3 + 2];
        const float d0 = h0 - y0, d1 = h1 - y1, d2 = h2 - y2;
        // winner-takes-all part of this frame
        float r;
        if (kSquared) {
          r = c_wta * w[j];
        } else {
          const float n2 = fmaf(d2, d2, fmaf(d1, d1, d0 * d0));
          r = n2 > 0.f ? c_wta * w[j] * rsqrt_approx(n2) : 0.f;
        }
        r = winner ? r : 0.f;
        // flux of the pair (f, f+1)
        const float a0 = hc[kF + j * 3 + 0] - h0, a1 = hc[kF + j * 3 + 1] - h1, a2 = hc[kF + j * 3 + 2] - h2;
        const float e0 = a0 - (yc[kF + j * 3 + 0] - y0), e1 = a1 - (yc[kF + j * 3 + 1] - y1), e2 = a2 - (yc[kF + j * 3 + 2] - y2);
        float rv;
        if (kSquared) {
          rv = c_vel;
        } else {
          const float q = fmaf(e2, e2, fmaf(e1, e1, e0 * e0));
          rv = q > 0.f ? c_vel * rsqrt_approx(q) : 0.f;
        }
        const float cs = c_sm * w[j];
        float f0 = fmaf(rv, e0, cs * a0), f1 = fmaf(rv, e1, cs * a1), f2 = fmaf(rv, e2, cs * a2);
        f0 = has_pair ? f0 : 0.f;
        f1 = has_pair ? f1 : 0.f;
        f2 = has_pair ? f2 : 0.f;
        const float p0 = __shfl_up_sync(0xffffffffu, f0, 1), p1 = __shfl_up_sync(0xffffffffu, f1, 1), p2 = __shfl_up_sync(0xffffffffu, f2, 1);
        g[j * 3 + 0] = fmaf(r, d0, p0 - f0);
        g[j * 3 + 1] = fmaf(r, d1, p1 - f1);
        g[j * 3 + 2] = fmaf(r, d2, p2 - f2);
      }
      if (writes && grad_scores) {
        const float s = sck;
        const float tgt = winner ? 1.f : 0.f;
        // binary_cross_entropy_backward: (input - target) / max((1 - input) * input, 1e-12)
        grad_scores[si] = c_bce * (s - tgt) / fmaxf((1.f - s) * s, 1e-12f);
      }
      __syncwarp();  // all lanes done reading hbuf[cur]: reuse it to transpose the gradient tile
      float* tb = st.buf[cur];
      if (writes) {
#pragma unroll
        for (int i = 0; i < kF; ++i) tb[(lane - 1) * kF + i] = g[i];
      }
      __syncwarp();
      float* gout = grad_hyp + (((size_t)b * d.K + k) * d.T + t0) * kF;
      if (nf == kBwdFrames) {
#pragma unroll
        for (int i = 0; i < (kBwdFrames * kF + 31) / 32; ++i)
          if (i * 32 + lane < kBwdFrames * kF) gout[i * 32 + lane] = tb[i * 32 + lane];
      } else {
        for (uint32_t i = lane; i < nf * kF; i += 32) gout[i] = tb[i];
      }
      ptx::fence_proxy_async_smem();   // the next bulk copy into this buffer follows generic-proxy writes to it
      sh_cur = sh_next;
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------------------ aggregation
// Index math of the element-wise aggregation kernels: out element i = (b, r) with r the offset inside clip b's [T,17,3] block, so the
// hypothesis element is hyp[(b K + k) T 51 + r] and its score scores[(b K + k) T + r / 51]: one runtime division per element, 32-bit
// whenever the output has fewer than 2^32 elements.
template <typename I>
__device__ __forceinline__ void aggregate_weighted_body(const float* __restrict__ hyp, const float* __restrict__ scores, float* __restrict__ out,
                                                        uint32_t B, uint32_t K, uint32_t T) {
  // torch.sum(hyp * scores.unsqueeze(-1), dim=1): products rounded, accumulated k = 0..K-1 left to right
  const I clip = (I)T * kF, n = (I)B * clip;
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (I)gridDim.x * blockDim.x) {
    const I b = i / clip, r = i - b * clip, t = r / kF;
    const float* hp = hyp + (size_t)b * K * clip + r;
    const float* sp = scores + (size_t)b * K * T + t;
    float acc = 0.f;
    for (uint32_t k = 0; k < K; ++k) acc = __fadd_rn(acc, __fmul_rn(hp[(size_t)k * clip], sp[(size_t)k * T]));
    out[i] = acc;
  }
}
__global__ void aggregate_weighted_kernel(const float* __restrict__ hyp, const float* __restrict__ scores, float* __restrict__ out,
                                          uint32_t B, uint32_t K, uint32_t T) {
  pdl_launch_dependents();
  pdl_wait();
  if ((uint64_t)B * T * kF < (1ull << 32))
    aggregate_weighted_body<uint32_t>(hyp, scores, out, B, K, T);
  else
    aggregate_weighted_body<uint64_t>(hyp, scores, out, B, K, T);
}

// Flip test-time augmentation epilogue (hpe/eval_utils.py:83-142): hyp / scores hold 2B clips, the last B being the forward of the
// horizontally flipped input; out[b] = (aggregate(hyp[b]) + unflip(aggregate(hyp[B + b]))) / 2, unflip = negate x and swap the left / right
// joints (augmentations/functional.py:7-28).  Same operation order as aggregate_weighted_kernel / argmax + gather, so the result equals
// the reference's separate steps bit for bit.
__host__ __device__ constexpr int flip_joint(int j) {
  constexpr int f[kJ] = {0, 4, 5, 6, 1, 2, 3, 7, 8, 9, 10, 14, 15, 16, 11, 12, 13};
  return f[j];
}
template <bool kBestScore>
__global__ void aggregate_tta_kernel(const float* __restrict__ hyp, const float* __restrict__ scores, float* __restrict__ out, uint32_t B,
                                     uint32_t K, uint32_t T) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t n = (size_t)B * T * kF;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const size_t bt = i / kF;
    const uint32_t e = (uint32_t)(i - bt * kF);
    const uint32_t j = e / 3, c = e - j * 3;
    const uint32_t b = (uint32_t)(bt / T), t = (uint32_t)(bt - (size_t)b * T);
    int jf = 0;
#pragma unroll
    for (int q = 0; q < kJ; ++q)
      if ((int)j == q) jf = flip_joint(q);
    const uint32_t ef = (uint32_t)jf * 3 + c;
    float v[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t bb = b + (h ? B : 0u);
      const uint32_t el = h ? ef : e;
      if (kBestScore) {
        float best = scores[((size_t)bb * K) * T + t];
        uint32_t bk = 0;
        for (uint32_t k = 1; k < K; ++k) {
          const float sc = scores[((size_t)bb * K + k) * T + t];
          if (sc > best) {  // torch.argmax: first maximal index
            best = sc;
            bk = k;
          }
        }
        v[h] = hyp[(((size_t)bb * K + bk) * T + t) * kF + el];
      } else {
        float acc = 0.f;
        for (uint32_t k = 0; k < K; ++k) {
          const size_t f = ((size_t)bb * K + k) * T + t;
          acc = __fadd_rn(acc, __fmul_rn(hyp[f * kF + el], scores[f]));
        }
        v[h] = acc;
      }
    }
    if (c == 0) v[1] = -v[1];
    out[i] = __fmul_rn(__fadd_rn(v[0], v[1]), 0.5f);
  }
}

__device__ __forceinline__ int argmax_score(const float* __restrict__ sp, uint32_t K, uint32_t T) {
  float best = sp[0];
  int bk = 0;
  for (uint32_t k = 1; k < K; ++k) {
    const float s = sp[(size_t)k * T];
    if (s > best) {  // torch.argmax: first maximal index
      best = s;
      bk = (int)k;
    }
  }
  return bk;
}

__global__ void argmax_score_kernel(const float* __restrict__ scores, int64_t* __restrict__ idx, uint32_t B, uint32_t K, uint32_t T) {
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t n = B * T;   // check_dims: B T < 2^31
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t b = i / T, t = i - b * T;
    idx[i] = argmax_score(scores + (size_t)b * K * T + t, K, T);
  }
}

// out[b,t] = hyp[b, idx[b,t], t]
template <typename I>
__device__ __forceinline__ void gather_hyp_body(const float* __restrict__ hyp, const int64_t* __restrict__ idx, float* __restrict__ out,
                                                uint32_t B, uint32_t K, uint32_t T) {
  const I clip = (I)T * kF, n = (I)B * clip;
  for (I i = (I)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (I)gridDim.x * blockDim.x) {
    const I b = i / clip, r = i - b * clip, t = r / kF;
    const int64_t k = idx[(size_t)b * T + t];
    out[i] = hyp[((size_t)b * K + (size_t)k) * clip + r];
  }
}
__global__ void gather_hyp_kernel(const float* __restrict__ hyp, const int64_t* __restrict__ idx, float* __restrict__ out, uint32_t B,
                                  uint32_t K, uint32_t T) {
  pdl_launch_dependents();
  pdl_wait();
  if ((uint64_t)B * T * kF < (1ull << 32))
    gather_hyp_body<uint32_t>(hyp, idx, out, B, K, T);
  else
    gather_hyp_body<uint64_t>(hyp, idx, out, B, K, T);
}

// ------------------------------------------------------------------------------------------------ MPJPE
constexpr int kMpjpeBlocks = 148 * 4;
__device__ __forceinline__ double point_dist(float g0, float g1, float g2, float p0, float p1, float p2) {
  const float d0 = g0 - p0, d1 = g1 - p1, d2 = g2 - p2;
  return (double)__fsqrt_rn(__fmaf_rn(d2, d2, __fmaf_rn(d1, d1, __fmul_rn(d0, d0))));
}
// kVec: both arrays 16-byte aligned -> a thread takes 4 points = 3 x 16 bytes of each array per step
template <bool kVec>
__global__ void mpjpe_partial_kernel(const float* __restrict__ pred, const float* __restrict__ gt, size_t n_points,
                                     double* __restrict__ partials) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ double red[8];
  double acc = 0;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  const size_t n_quads = kVec ? n_points / 4 : 0;
  if (kVec) {
    const float4* p4 = reinterpret_cast<const float4*>(pred);
    const float4* g4 = reinterpret_cast<const float4*>(gt);
    for (size_t q = tid; q < n_quads; q += stride) {
      const float4 pa = p4[q * 3], pb = p4[q * 3 + 1], pc = p4[q * 3 + 2];
      const float4 ga = g4[q * 3], gb = g4[q * 3 + 1], gc = g4[q * 3 + 2];
      acc += point_dist(ga.x, ga.y, ga.z, pa.x, pa.y, pa.z);
      acc += point_dist(ga.w, gb.x, gb.y, pa.w, pb.x, pb.y);
      acc += point_dist(gb.z, gb.w, gc.x, pb.z, pb.w, pc.x);
      acc += point_dist(gc.y, gc.z, gc.w, pc.y, pc.z, pc.w);
    }
  }
  for (size_t i = n_quads * 4 + tid; i < n_points; i += stride)
    acc += point_dist(gt[i * 3 + 0], gt[i * 3 + 1], gt[i * 3 + 2], pred[i * 3 + 0], pred[i * 3 + 1], pred[i * 3 + 2]);
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += red[i];
    partials[blockIdx.x] = s;
  }
}
// one warp, fixed order: lane l adds partials l, l+32, ..., then a butterfly
__global__ void mpjpe_finalize_kernel(const double* __restrict__ partials, int n, double n_points, float* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  double s = 0;
  for (int i = threadIdx.x; i < n; i += 32) s += partials[i];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (threadIdx.x == 0) {
    out[0] = (float)s;
    out[1] = (float)(s / n_points);
  }
}

int check_dims(const char* who, int64_t B, int64_t K, int64_t T) {
  MP_REQUIRE(B >= 1 && K >= 1 && T >= 1, MP_EINVAL, "%s: bad sizes B=%lld K=%lld T=%lld", who, (long long)B, (long long)K, (long long)T);
  MP_REQUIRE(K <= kMaxHyp, MP_EINVAL, "%s: n_hyp=%lld exceeds %d", who, (long long)K, kMaxHyp);
  MP_REQUIRE(B * K * T * kF < ((int64_t)1 << 40) && B * T < ((int64_t)1 << 31), MP_EINVAL, "%s: tensor too large", who);
  return MP_OK;
}

// one CTA per SM; with few tiles (the training step) one tile per SM before a second warp of any CTA gets one
int loss_grid(int64_t B, int64_t T, int frames_per_tile) {
  const int64_t items = B * ((T + frames_per_tile - 1) / frames_per_tile);
  const int64_t cap = (int64_t)sm_count();
  return (int)(items < cap ? items : cap);
}

}  // namespace
}  // namespace mp

extern "C" {

size_t mp_loss_workspace_bytes(int64_t, int64_t, int64_t) { return (size_t)mp::kMaxPartialWarps * 4 * sizeof(double); }

int mp_wta_fwd(const float* hyp, const float* y, const float* joint_weights, int squared, float* wta_val, int64_t* wta_idx,
               float* per_hyp, int64_t B, int64_t K, int64_t T, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_CHECK(check_dims("mp_wta_fwd", B, K, T));
  MP_REQUIRE(hyp && y && wta_val && wta_idx, MP_EINVAL, "mp_wta_fwd: null pointer");
  MP_REQUIRE(aligned16(hyp) && aligned16(y), MP_EALIGN, "mp_wta_fwd: hyp and y must be 16-byte aligned");
  // reference quirk (losses.py:57-58,126-138): squared + weights=None makes _l2_loss_per_hyp a 0-d tensor and
  // torch.min(dim=1) raises; keep that an error instead of inventing semantics.
  MP_REQUIRE(!(squared && joint_weights == nullptr), MP_EINVAL,
             "squared WTA loss without joint weights is an error in the reference (F.mse_loss returns a scalar)");
  const LossDims d{(uint32_t)B, (uint32_t)K, (uint32_t)T};
  const size_t smem = loss_smem_bytes(kFwdWarps, kFwdBufs);
  const int grid = loss_grid(B, T, kTileFrames);
  auto launch = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(kernel, grid, kFwdWarps * 32, smem, (cudaStream_t)stream, hyp, nullptr, y, joint_weights, d, wta_val, wta_idx, per_hyp, nullptr);
  };
  if (squared)
    launch(loss_fwd_kernel<true, false>);
  else
    launch(loss_fwd_kernel<false, false>);
  return check_launch("loss_fwd_kernel(wta)");
}

int mp_loss_fwd(const float* hyp, const float* scores, const float* y, const float* joint_weights, int squared, float beta,
                float vel_w, float smooth_w, float* terms, float* wta_val, int64_t* wta_idx, int64_t B, int64_t K, int64_t T,
                void* workspace, size_t workspace_bytes, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_CHECK(check_dims("mp_loss_fwd", B, K, T));
  MP_REQUIRE(hyp && y && terms && wta_val && wta_idx && workspace, MP_EINVAL, "mp_loss_fwd: null pointer");
  MP_REQUIRE(beta == 0.f || scores != nullptr, MP_EINVAL, "mp_loss_fwd: scores required when beta != 0");
  MP_REQUIRE(aligned16(hyp) && aligned16(y) && aligned16(workspace), MP_EALIGN, "mp_loss_fwd: pointers must be 16-byte aligned");
  MP_REQUIRE(workspace_bytes >= mp_loss_workspace_bytes(B, K, T), MP_EWORKSPACE, "mp_loss_fwd: workspace too small");
  MP_REQUIRE(!(squared && joint_weights == nullptr), MP_EINVAL,
             "squared WTA loss without joint weights is an error in the reference (F.mse_loss returns a scalar)");
  const LossDims d{(uint32_t)B, (uint32_t)K, (uint32_t)T};
  const size_t smem = loss_smem_bytes(kFwdWarps, kFwdBufs);
  const int grid = loss_grid(B, T, kTileFrames);
  double* partials = reinterpret_cast<double*>(workspace);
  auto launch = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(kernel, grid, kFwdWarps * 32, smem, (cudaStream_t)stream, hyp, scores, y, joint_weights, d, wta_val, wta_idx, nullptr, partials);
  };
  if (squared)
    launch(loss_fwd_kernel<true, true>);
  else
    launch(loss_fwd_kernel<false, true>);
  MP_CHECK(check_launch("loss_fwd_kernel"));
  launch_k(loss_finalize_kernel, 1, 256, 0, (cudaStream_t)stream, partials, grid * kFwdWarps, d, squared, beta, vel_w, smooth_w, terms);
  return check_launch("loss_finalize_kernel");
}

int mp_loss_bwd(const float* hyp, const float* scores, const float* y, const float* joint_weights, const int64_t* wta_idx,
                int squared, float beta, float vel_w, float smooth_w, const float* grad_terms, const float* grad_wta_val,
                float* grad_hyp, float* grad_scores, int64_t B, int64_t K, int64_t T, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_CHECK(check_dims("mp_loss_bwd", B, K, T));
  MP_REQUIRE(hyp && y && wta_idx && grad_hyp && grad_terms, MP_EINVAL, "mp_loss_bwd: null pointer");
  MP_REQUIRE(grad_scores == nullptr || scores != nullptr, MP_EINVAL, "mp_loss_bwd: scores required for grad_scores");
  MP_REQUIRE(aligned16(hyp) && aligned16(y), MP_EALIGN, "mp_loss_bwd: hyp and y must be 16-byte aligned");
  const LossDims d{(uint32_t)B, (uint32_t)K, (uint32_t)T};
  const size_t smem = loss_smem_bytes(kBwdWarps, kBwdBufs);
  const int grid = loss_grid(B, T, kBwdFrames);
  auto launch = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(kernel, grid, kBwdWarps * 32, smem, (cudaStream_t)stream, hyp, scores, y, joint_weights, wta_idx, d, beta, vel_w, smooth_w,
                                                                   grad_terms, grad_wta_val, grad_hyp, grad_scores);
  };
  if (squared)
    launch(loss_bwd_kernel<true>);
  else
    launch(loss_bwd_kernel<false>);
  return check_launch("loss_bwd_kernel");
}

int mp_aggregate(const float* hyp, const float* scores, const float* y, int mode, float* out_pose, float* out_val,
                 int64_t* out_idx, int64_t B, int64_t K, int64_t T, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_CHECK(check_dims("mp_aggregate", B, K, T));
  MP_REQUIRE(hyp && out_pose, MP_EINVAL, "mp_aggregate: null pointer");
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)B * T * kF;
  const int blocks = (int)((n + 255) / 256 < (size_t)sm_count() * 8 ? (n + 255) / 256 : (size_t)sm_count() * 8);
  if (mode == MP_AGG_WEIGHTED_AVE) {
    MP_REQUIRE(scores != nullptr, MP_EINVAL, "Scores required to compute weighted hypothesis average.");
    launch_k(aggregate_weighted_kernel, blocks, 256, 0, s, hyp, scores, out_pose, (uint32_t)B, (uint32_t)K, (uint32_t)T);
    return check_launch("aggregate_weighted_kernel");
  }
  MP_REQUIRE(out_idx != nullptr, MP_EINVAL, "mp_aggregate: out_idx required for best_score / oracle");
  if (mode == MP_AGG_BEST_SCORE) {
    MP_REQUIRE(scores != nullptr, MP_EINVAL, "Scores required to compute hypothesis with best confidence.");
    const int b2 = (int)(((size_t)B * T + 255) / 256);
    launch_k(argmax_score_kernel, b2 < sm_count() * 8 ? b2 : sm_count() * 8, 256, 0, s, scores, out_idx, (uint32_t)B, (uint32_t)K, (uint32_t)T);
    MP_CHECK(check_launch("argmax_score_kernel"));
  } else if (mode == MP_AGG_ORACLE) {
    MP_REQUIRE(y != nullptr && out_val != nullptr, MP_EINVAL, "Ground-truth required to compute best hypothesis.");
    MP_CHECK(mp_wta_fwd(hyp, y, nullptr, 0, out_val, out_idx, nullptr, B, K, T, stream));
  } else {
    return fail(MP_EINVAL, "Only best_score and weighted_ave modes are implemented.Got %d.", mode);
  }
  launch_k(gather_hyp_kernel, blocks, 256, 0, s, hyp, out_idx, out_pose, (uint32_t)B, (uint32_t)K, (uint32_t)T);
  return check_launch("gather_hyp_kernel");
}

int mp_aggregate_tta(const float* hyp, const float* scores, int mode, float* out_pose, int64_t B, int64_t K, int64_t T, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_CHECK(check_dims("mp_aggregate_tta", 2 * B, K, T));
  MP_REQUIRE(hyp && out_pose, MP_EINVAL, "mp_aggregate_tta: null pointer");
  MP_REQUIRE(scores != nullptr, MP_EINVAL, "Scores required to compute weighted hypothesis average.");
  MP_REQUIRE(mode == MP_AGG_WEIGHTED_AVE || mode == MP_AGG_BEST_SCORE, MP_EINVAL, "Only best_score and weighted_ave modes are implemented.Got %d.", mode);
  if (B == 0) return MP_OK;
  const size_t n = (size_t)B * T * kF;
  const int blocks = (int)((n + 255) / 256 < (size_t)sm_count() * 8 ? (n + 255) / 256 : (size_t)sm_count() * 8);
  if (mode == MP_AGG_WEIGHTED_AVE)
    launch_k(aggregate_tta_kernel<false>, blocks, 256, 0, (cudaStream_t)stream, hyp, scores, out_pose, (uint32_t)B, (uint32_t)K, (uint32_t)T);
  else
    launch_k(aggregate_tta_kernel<true>, blocks, 256, 0, (cudaStream_t)stream, hyp, scores, out_pose, (uint32_t)B, (uint32_t)K, (uint32_t)T);
  return check_launch("aggregate_tta_kernel");
}

size_t mp_mpjpe_workspace_bytes(int64_t) { return (size_t)mp::kMpjpeBlocks * sizeof(double); }

int mp_mpjpe(const float* pred, const float* gt, int64_t n_points, float* out, void* workspace, size_t workspace_bytes,
             mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(pred && gt && out && workspace && n_points >= 1, MP_EINVAL, "mp_mpjpe: bad arguments");
  MP_REQUIRE(workspace_bytes >= mp_mpjpe_workspace_bytes(n_points), MP_EWORKSPACE, "mp_mpjpe: workspace too small");
  int blocks = (int)((n_points + 255) / 256);
  if (blocks > kMpjpeBlocks) blocks = kMpjpeBlocks;
  double* partials = reinterpret_cast<double*>(workspace);
  if (aligned16(pred) && aligned16(gt)) {
    blocks = (int)((n_points / 4 + 255) / 256);
    blocks = blocks < 1 ? 1 : (blocks > kMpjpeBlocks ? kMpjpeBlocks : blocks);
    launch_k(mpjpe_partial_kernel<true>, blocks, 256, 0, (cudaStream_t)stream, pred, gt, (size_t)n_points, partials);
  } else {
    launch_k(mpjpe_partial_kernel<false>, blocks, 256, 0, (cudaStream_t)stream, pred, gt, (size_t)n_points, partials);
  }
  MP_CHECK(check_launch("mpjpe_partial_kernel"));
  launch_k(mpjpe_finalize_kernel, 1, 32, 0, (cudaStream_t)stream, partials, blocks, (double)n_points, out);
  return check_launch("mpjpe_finalize_kernel");
}

}  // extern "C"

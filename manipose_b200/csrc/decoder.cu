// Manifold decoder forward: 6-D -> SO(3) (Gram-Schmidt) + T-pose from bone lengths + forward kinematics down the
// 17-joint tree + softmax over the K hypotheses, one memory-bound kernel.
//
// Replaces (reference, paths under hpe/mh_so3_hpe/architectures/):
//   utils/rotation_tools.py:6-57    normalize_vector / cross_product / compute_rotation_matrix_from_ortho6d
//   pose_decoder.py:85-120          _compute_bones_length / build_t_pose_from_bone_lengths
//   utils/forward_kinematics.py:6-48 forward_kinematics
//   rmcl_manifold_mix_ste.py:262    scores_logits.softmax(dim=1)
//
// Data movement: a warp owns a tile of 32 consecutive poses = 13,056 contiguous bytes of rot6d.  One lane pulls the
// tile into shared memory with a single bulk async copy (TMA engine, mbarrier completion); each lane then decodes ONE
// pose reading its 102 floats as 8-byte conflict-free LDS; the 51 results go back into the same (dead) tile buffer and
// leave with one bulk async store.  Algorithmic HBM traffic: 408 + 204 bytes per pose (+8 for logit/score).
//
// Arithmetic: in EXACT mode every operation is one correctly-rounded IEEE fp32 operation in the reference's order (unfused
// mul/add/sub, sqrt, max, three divisions, sequential k-loop of the 3x3 products), so poses are bit-identical to the
// oracle's correctly-rounded restatement (oracle.pose_decoder_ieee) and within ~1e-7 relative of the torch-CPU reference
// (whose vectorised sqrt is itself 1 ulp off on ~0.7 % of inputs).  FAST mode uses rsqrt and FMA contraction (<= 1e-6).
#include <algorithm>

#include "common.cuh"
#include "ieee.cuh"
#include "ptx.cuh"

namespace mp {
namespace {

template <int RD>
constexpr int in_floats() { return kJ * RD; }   // 102 (6-D) / 68 (4-D) floats per pose in
constexpr int kOut = kJ * 3;   // 51 floats per pose out
constexpr int kTile = 32;      // poses per warp tile
template <int RD>
constexpr int tile_in_bytes() { return kTile * kJ * RD * 4; }   // 13056 / 8704
constexpr int kTileOutBytes = kTile * kOut * 4;  // 6528
constexpr int kWarpsPerCta = 4;

// Arithmetic modes: kFast (rsqrt, FMA contraction), kIeee (every operation one correctly rounded IEEE intrinsic: the reference's
// sequence, with the range checks and slow paths of sqrt.rn / div.rn), kExact (the same results from the branch-free refinement
// sequences of ieee.cuh; operands outside their range only raise `bad`, and the caller redoes that pose in kIeee mode).
enum { kFast = 0, kIeee = 1, kExact = 2 };

template <int kMode>
struct Arith {
  static __device__ __forceinline__ float mul(float a, float b) { return kMode != kFast ? __fmul_rn(a, b) : a * b; }
  static __device__ __forceinline__ float add(float a, float b) { return kMode != kFast ? __fadd_rn(a, b) : a + b; }
  static __device__ __forceinline__ float sub(float a, float b) { return kMode != kFast ? __fsub_rn(a, b) : a - b; }
  // v / max(sqrt(v.v), 1e-8)   (rotation_tools.py:6-17)
  static __device__ __forceinline__ void normalize(float& x, float& y, float& z, uint32_t& bad) {
    if (kMode == kExact) {
      const float s = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
      const float lo = fminf(fminf(fabsf(x), fabsf(y)), fabsf(z));
      // one range check for the square root and the three divisions: s in [2^-80, 2^80], every |component| >= 2^-100
      bad |= (uint32_t)((__float_as_uint(s) - 0x17800000u) > 0x50000000u) | (uint32_t)(__float_as_uint(lo) < 0x0d800000u);
      const float m = fmaxf(ieee::sqrt_rn_core(s), 1e-8f);
      const float r = ieee::rcp_refined(m);
      x = ieee::div_rn_core(x, m, r);
      y = ieee::div_rn_core(y, m, r);
      z = ieee::div_rn_core(z, m, r);
    } else if (kMode == kIeee) {
      const float s = __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
      const float m = fmaxf(__fsqrt_rn(s), 1e-8f);
      x = __fdiv_rn(x, m);
      y = __fdiv_rn(y, m);
      z = __fdiv_rn(z, m);
    } else {
      float s = x * x + y * y + z * z;
      float inv = (s >= 1e-16f) ? rsqrtf(s) : 1e8f;
      x *= inv;
      y *= inv;
      z *= inv;
    }
  }
  // u x v   (rotation_tools.py:21-32)
  static __device__ __forceinline__ void cross(float u0, float u1, float u2, float v0, float v1, float v2, float& i, float& j,
                                               float& k) {
    i = sub(mul(u1, v2), mul(u2, v1));
    j = sub(mul(u2, v0), mul(u0, v2));
    k = sub(mul(u0, v1), mul(u1, v0));
  }
  // dot of a row with a column, torch CPU bmm order: ((a0 b0) + a1 b1) + a2 b2, each op rounded
  static __device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2) {
    return add(add(mul(a0, b0), mul(a1, b1)), mul(a2, b2));
  }
};

// Per-lane state of one pose; every index is a compile-time constant after template expansion, so all of it is
// scalarised into registers and the live ranges follow the tree (at most ~3 world rotations alive).
struct PoseState {
  float rw[kJ][9];   // world rotations, row-major (leaves never stored: forward_kinematics.py:41-46)
  float pos[kOut];   // joint positions
  float t[kJ][2];    // T-pose x / y coordinates (z is identically 0)
  uint32_t bad;      // kExact: some operand left the range of the branch-free sqrt / division sequences
};

// v / max(sqrt(v.v), 1e-8) for a 2-vector (compute_rotation_matrix_from_ortho4d uses normalize_vector on [M,2] rows)
template <int kMode>
__device__ __forceinline__ void normalize2(float& x, float& y, uint32_t& bad) {
  if (kMode == kExact) {
    const float s = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
    const float lo = fminf(fabsf(x), fabsf(y));
    bad |= (uint32_t)((__float_as_uint(s) - 0x17800000u) > 0x50000000u) | (uint32_t)(__float_as_uint(lo) < 0x0d800000u);   // see Arith::normalize
    const float m = fmaxf(ieee::sqrt_rn_core(s), 1e-8f);
    const float r = ieee::rcp_refined(m);
    x = ieee::div_rn_core(x, m, r);
    y = ieee::div_rn_core(y, m, r);
  } else if (kMode == kIeee) {
    const float m = fmaxf(__fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))), 1e-8f);
    x = __fdiv_rn(x, m);
    y = __fdiv_rn(y, m);
  } else {
    const float s = x * x + y * y;
    const float inv = (s >= 1e-16f) ? rsqrtf(s) : 1e8f;
    x *= inv;
    y *= inv;
  }
}

// local rotation of joint j, row-major r[row][col]
template <int kMode, int RD>
__device__ __forceinline__ void local_rotation(const float* __restrict__ a, float (&r)[9], uint32_t& bad) {
  using A = Arith<kMode>;
  if constexpr (RD == 6) {
    // ---- 6-D -> rotation matrix with columns [x y z]  (rotation_tools.py:35-57)
    const float2 v01 = *reinterpret_cast<const float2*>(a + 0);
    const float2 v23 = *reinterpret_cast<const float2*>(a + 2);
    const float2 v45 = *reinterpret_cast<const float2*>(a + 4);
    float x0 = v01.x, x1 = v01.y, x2 = v23.x;
    const float b0 = v23.y, b1 = v45.x, b2 = v45.y;
    A::normalize(x0, x1, x2, bad);
    float z0, z1, z2;
    A::cross(x0, x1, x2, b0, b1, b2, z0, z1, z2);
    A::normalize(z0, z1, z2, bad);
    float y0, y1, y2;
    A::cross(z0, z1, z2, x0, x1, x2, y0, y1, y2);
    r[0] = x0; r[1] = y0; r[2] = z0;
    r[3] = x1; r[4] = y1; r[5] = z1;
    r[6] = x2; r[7] = y2; r[8] = z2;
  } else {
    // ---- 4-D -> R_theta R_phi  (rotation_tools.py:60-116): (c_t, s_t) = n(a[0:2]), (c_p, s_p) = n(a[2:4]);
    //   R_theta = [[s_t, c_t, 0], [-c_t, s_t, 0], [0, 0, 1]],  R_phi = [[1, 0, 0], [0, c_p, -s_p], [0, s_p, c_p]]
    // (the products with the exact 0 / 1 entries of the reference's bmm are exact, so these six products are all there is)
    const float2 th = *reinterpret_cast<const float2*>(a + 0);
    const float2 ph = *reinterpret_cast<const float2*>(a + 2);
    float ct = th.x, st = th.y, cp = ph.x, sp = ph.y;
    normalize2<kMode>(ct, st, bad);
    normalize2<kMode>(cp, sp, bad);
    r[0] = st;  r[1] = A::mul(ct, cp); r[2] = -A::mul(ct, sp);
    r[3] = -ct; r[4] = A::mul(st, cp); r[5] = -A::mul(st, sp);
    r[6] = 0.f; r[7] = sp;             r[8] = cp;
  }
}

template <int kMode, int RD, int j>
__device__ __forceinline__ void decode_joint(PoseState& st, const float* __restrict__ lane_in, const float* __restrict__ len) {
  using A = Arith<kMode>;
  float r[9];
  local_rotation<kMode, RD>(lane_in + j * RD, r, st.bad);
  if constexpr (j == 0) {
#pragma unroll
    for (int i = 0; i < 9; ++i) st.rw[0][i] = r[i];
    st.t[0][0] = 0.f;
    st.t[0][1] = 0.f;
    // st.pos[0..2] preset by the caller (root position)
  } else {
    constexpr int p = parent_of(j);
    constexpr int ax = axis_of(j);
    constexpr float sg = sign_of(j);
    // T-pose by cumulative adds, offset recovered by subtraction (pose_decoder.py:115-119, forward_kinematics.py:31-33)
    const float step = sg > 0.f ? len[j - 1] : -len[j - 1];
    st.t[j][ax] = A::add(st.t[p][ax], step);
    st.t[j][1 - ax] = st.t[p][1 - ax];
    const float off = A::sub(st.t[j][ax], st.t[p][ax]);
    const float* rp = st.rw[p];
    if constexpr (is_leaf(j)) {
      // only column `ax` of Rw[j] = Rw[p] R[j] is needed for the position
#pragma unroll
      for (int row = 0; row < 3; ++row) {
        const float c = A::dot3(rp[row * 3 + 0], rp[row * 3 + 1], rp[row * 3 + 2], r[0 * 3 + ax], r[1 * 3 + ax], r[2 * 3 + ax]);
        st.pos[j * 3 + row] = A::add(A::mul(c, off), st.pos[p * 3 + row]);
      }
    } else {
#pragma unroll
      for (int row = 0; row < 3; ++row) {
#pragma unroll
        for (int col = 0; col < 3; ++col)
          st.rw[j][row * 3 + col] =
              A::dot3(rp[row * 3 + 0], rp[row * 3 + 1], rp[row * 3 + 2], r[0 * 3 + col], r[1 * 3 + col], r[2 * 3 + col]);
        st.pos[j * 3 + row] = A::add(A::mul(st.rw[j][row * 3 + ax], off), st.pos[p * 3 + row]);
      }
    }
  }
  if constexpr (j + 1 < kJ) decode_joint<kMode, RD, j + 1>(st, lane_in, len);
}

// The cold path of kExact: one pose with the IEEE intrinsics, out of line so that its 136 slow-path call sites stay out of the kernel's
// instruction stream.  The 51 results go through local memory.
template <int RD>
__device__ __noinline__ void decode_pose_ieee(const float* __restrict__ lane_in, const float* __restrict__ len, float r0, float r1, float r2,
                                              float* __restrict__ out) {
  PoseState st;
  st.bad = 0;
  st.pos[0] = r0;
  st.pos[1] = r1;
  st.pos[2] = r2;
  decode_joint<kIeee, RD, 0>(st, lane_in, len);
#pragma unroll
  for (int i = 0; i < kOut; ++i) out[i] = st.pos[i];
}

// softmax over the hypothesis dim, one (clip, frame) per thread, grid-stride
__device__ __forceinline__ void softmax_hyp_rows(const float* __restrict__ logits, float* __restrict__ scores, uint32_t n_clips,
                                                 uint32_t n_hyp, uint32_t n_frames) {
  const uint32_t n_items = n_clips * n_frames;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_items; idx += gridDim.x * blockDim.x) {
    const uint32_t b = idx / n_frames, t = idx - b * n_frames;
    const float* lg = logits + (size_t)b * n_hyp * n_frames + t;
    float mx = -INFINITY;
    for (uint32_t k = 0; k < n_hyp; ++k) mx = fmaxf(mx, lg[(size_t)k * n_frames]);
    float sum = 0.f;
    for (uint32_t k = 0; k < n_hyp; ++k) sum += expf(lg[(size_t)k * n_frames] - mx);
    float* sc = scores + (size_t)b * n_hyp * n_frames + t;
    for (uint32_t k = 0; k < n_hyp; ++k) sc[(size_t)k * n_frames] = expf(lg[(size_t)k * n_frames] - mx) / sum;
  }
}

template <bool kBitExact, int RD>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 4)
decoder_fwd_kernel(const float* __restrict__ rot6d, const float* __restrict__ bone_len, const float* __restrict__ root,
                   const float* __restrict__ logits, float* __restrict__ poses, float* __restrict__ scores, uint32_t n_poses,
                   uint32_t poses_per_clip, uint32_t n_clips, uint32_t n_hyp, uint32_t n_frames, int bulk_in, int bulk_out) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int kIn = in_floats<RD>();
  constexpr int kTileInBytes = tile_in_bytes<RD>();
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  float* tile = reinterpret_cast<float*>(smem_raw + warp * kTileInBytes);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + kWarpsPerCta * kTileInBytes) + warp;

  if (lane == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_mbar_init();
  }
  __syncwarp();

  const uint32_t n_tiles = (n_poses + kTile - 1) / kTile;
  const uint32_t warp_global = blockIdx.x * kWarpsPerCta + warp;
  const uint32_t warp_stride = gridDim.x * kWarpsPerCta;
  uint32_t phase = 0;

  for (uint32_t tile_idx = warp_global; tile_idx < n_tiles; tile_idx += warp_stride) {
    const uint32_t pose0 = tile_idx * kTile;
    const uint32_t n_here = min((uint32_t)kTile, n_poses - pose0);
    const bool full_in = bulk_in && (n_here == kTile);
    const bool full = bulk_out && (n_here == kTile);
    const float* gin = rot6d + (size_t)pose0 * kIn;
    float* gout = poses + (size_t)pose0 * kOut;

    // ---- stage the tile of rot6d into shared memory
    if (full_in) {
      if (lane == 0) {
        ptx::mbar_expect_tx(bar, kTileInBytes);
        ptx::bulk_g2s(tile, gin, kTileInBytes, bar);
      }
      ptx::mbar_wait(bar, phase);
      phase ^= 1;
    } else {
      for (uint32_t i = lane; i < n_here * kIn; i += 32) tile[i] = gin[i];
      __syncwarp();
    }

    // ---- one pose per lane
    PoseState st;
    const uint32_t pose = pose0 + lane;
    const bool active = lane < n_here;
    if (active) {
      const uint32_t clip = pose / poses_per_clip;
      float len[kBones];
      const float4* lp = reinterpret_cast<const float4*>(bone_len + (size_t)clip * kBones);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 v = __ldg(lp + i);
        len[4 * i + 0] = v.x;
        len[4 * i + 1] = v.y;
        len[4 * i + 2] = v.z;
        len[4 * i + 3] = v.w;
      }
      if (root != nullptr) {
        st.pos[0] = root[(size_t)pose * 3 + 0];
        st.pos[1] = root[(size_t)pose * 3 + 1];
        st.pos[2] = root[(size_t)pose * 3 + 2];
      } else {
        st.pos[0] = st.pos[1] = st.pos[2] = 0.f;
      }
      st.bad = 0;
      const float r0 = st.pos[0], r1 = st.pos[1], r2 = st.pos[2];
      decode_joint<(kBitExact ? kExact : kFast), RD, 0>(st, tile + lane * kIn, len);
      if (kBitExact && st.bad) {   // rare (zero / denormal / huge components): redo this pose with the IEEE intrinsics
        float redo[kOut];
        decode_pose_ieee<RD>(tile + lane * kIn, len, r0, r1, r2, redo);
#pragma unroll
        for (int i = 0; i < kOut; ++i) st.pos[i] = redo[i];
      }
    }
    __syncwarp();  // every lane is done reading the tile: reuse it for the output

    if (active) {
#pragma unroll
      for (int i = 0; i < kOut; ++i) tile[lane * kOut + i] = st.pos[i];
    }
    if (full) {
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::bulk_s2g(gout, tile, kTileOutBytes);
        ptx::bulk_commit();
        ptx::bulk_wait_read<0>();  // the buffer is about to be refilled
      }
      __syncwarp();
    } else {
      __syncwarp();
      for (uint32_t i = lane; i < n_here * kOut; i += 32) gout[i] = tile[i];
      __syncwarp();
    }
  }

  // ---- hypothesis scores: softmax over K of logits[b, :, t]  (rmcl_manifold_mix_ste.py:262)
  if (logits != nullptr) softmax_hyp_rows(logits, scores, n_clips, n_hyp, n_frames);
}

__global__ void softmax_hyp_fwd_kernel(const float* __restrict__ logits, float* __restrict__ scores, uint32_t n_clips, uint32_t n_hyp,
                                       uint32_t n_frames) {
  pdl_launch_dependents();
  pdl_wait();
  softmax_hyp_rows(logits, scores, n_clips, n_hyp, n_frames);
}

__global__ void softmax_hyp_bwd_kernel(const float* __restrict__ scores, const float* __restrict__ grad_scores,
                                       float* __restrict__ grad_logits, uint32_t n_clips, uint32_t n_hyp, uint32_t n_frames) {
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t n_items = n_clips * n_frames;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n_items; idx += gridDim.x * blockDim.x) {
    const uint32_t b = idx / n_frames, t = idx - b * n_frames;
    const size_t base = (size_t)b * n_hyp * n_frames + t;
    float dot = 0.f;
    for (uint32_t k = 0; k < n_hyp; ++k) dot += scores[base + (size_t)k * n_frames] * grad_scores[base + (size_t)k * n_frames];
    for (uint32_t k = 0; k < n_hyp; ++k) {
      const size_t i = base + (size_t)k * n_frames;
      grad_logits[i] = scores[i] * (grad_scores[i] - dot);
    }
  }
}


// ------------------------------------------------------------------------------------------------ backward
// Reverse-mode of the fused decoder, one pose per lane, depth-first over the kinematic tree: each joint recomputes its
// local rotation from the staged 6-D input, recurses into its children (which accumulate into its world-rotation and
// position gradients), then back-propagates through Rw[j] = Rw[p] R[j], the bone offset and the Gram-Schmidt map, and
// overwrites its own 6-D slot of the staged tile with the gradient (the tile leaves with one bulk store).
struct BwdCtx {
  float* r6;           // lane's 102 staged floats: rot6d in, grad_rot6d out (in place, joint by joint)
  const float* gp;     // lane's 51 staged floats of grad_poses
  const float* len;    // 16 bone lengths of the lane's clip
  float glen[kBones];  // gradient w.r.t. the bone lengths (this pose's contribution)
};

// 1 / max(sqrt(s), 1e-8) and the clamp flag.  The backward recompute needs x, y, z to gradient accuracy, not bit-exactly: one MUFU.RSQ
// (2^-22 relative) instead of an IEEE square root and division with their range checks and slow-path branches.
__device__ __forceinline__ float inv_norm(float s, bool& clamped) {
  clamped = !(s > 1e-16f);
  return clamped ? 1e8f : ieee::mufu_rsq(s);
}

__device__ __forceinline__ void gs_forward(const float* a6, float (&x)[3], float (&z)[3], float (&y)[3], float& ina, float& inw,
                                           bool& a_clamped, bool& w_clamped) {
  const float a0 = a6[0], a1 = a6[1], a2 = a6[2], b0 = a6[3], b1 = a6[4], b2 = a6[5];
  ina = inv_norm(a0 * a0 + a1 * a1 + a2 * a2, a_clamped);
  x[0] = a0 * ina; x[1] = a1 * ina; x[2] = a2 * ina;
  const float w0 = x[1] * b2 - x[2] * b1;
  const float w1 = x[2] * b0 - x[0] * b2;
  const float w2 = x[0] * b1 - x[1] * b0;
  inw = inv_norm(w0 * w0 + w1 * w1 + w2 * w2, w_clamped);
  z[0] = w0 * inw; z[1] = w1 * inw; z[2] = w2 * inw;
  y[0] = z[1] * x[2] - z[2] * x[1];
  y[1] = z[2] * x[0] - z[0] * x[2];
  y[2] = z[0] * x[1] - z[1] * x[0];
}

// gradient of v / max(|v|, 1e-8) given the normalised vector u, the inverse of the clamped norm and the upstream gradient gu
__device__ __forceinline__ void normalize_bwd(const float (&u)[3], float inv, bool clamped, const float (&gu)[3], float (&gv)[3]) {
  const float d = clamped ? 0.f : (u[0] * gu[0] + u[1] * gu[1] + u[2] * gu[2]);
  gv[0] = (gu[0] - u[0] * d) * inv;
  gv[1] = (gu[1] - u[1] * d) * inv;
  gv[2] = (gu[2] - u[2] * d) * inv;
}

template <int RD, int j, int c>
struct Children;

template <int RD, int j>
__device__ __forceinline__ void bwd_joint(BwdCtx& ctx, const float (&rwp)[9], float (&g_rwp)[9], float (&g_posp)[3]) {
  // ---- forward recompute for this joint
  float x[3], y[3], z[3], ina, inw;          // 6-D intermediates
  float ct, st, cp, sp, int_, inp;           // 4-D intermediates
  bool clamp_a, clamp_b;                     // norm clamped at 1e-8 (first / second normalisation)
  float r[9];
  if constexpr (RD == 6) {
    gs_forward(ctx.r6 + j * 6, x, z, y, ina, inw, clamp_a, clamp_b);
    r[0] = x[0]; r[1] = y[0]; r[2] = z[0];
    r[3] = x[1]; r[4] = y[1]; r[5] = z[1];
    r[6] = x[2]; r[7] = y[2]; r[8] = z[2];
  } else {
    const float* a4 = ctx.r6 + j * 4;
    int_ = inv_norm(a4[0] * a4[0] + a4[1] * a4[1], clamp_a);
    inp = inv_norm(a4[2] * a4[2] + a4[3] * a4[3], clamp_b);
    ct = a4[0] * int_; st = a4[1] * int_; cp = a4[2] * inp; sp = a4[3] * inp;
    r[0] = st;  r[1] = ct * cp; r[2] = -ct * sp;
    r[3] = -ct; r[4] = st * cp; r[5] = -st * sp;
    r[6] = 0.f; r[7] = sp;      r[8] = cp;
  }
  float rw[9];
  if constexpr (j == 0) {
#pragma unroll
    for (int i = 0; i < 9; ++i) rw[i] = r[i];
  } else {
#pragma unroll
    for (int row = 0; row < 3; ++row)
#pragma unroll
      for (int col = 0; col < 3; ++col)
        rw[row * 3 + col] = rwp[row * 3 + 0] * r[0 * 3 + col] + rwp[row * 3 + 1] * r[1 * 3 + col] + rwp[row * 3 + 2] * r[2 * 3 + col];
  }
  float g_rw[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  float g_pos[3] = {ctx.gp[j * 3 + 0], ctx.gp[j * 3 + 1], ctx.gp[j * 3 + 2]};
  // ---- children accumulate into g_rw / g_pos
  Children<RD, j, j + 1>::run(ctx, rw, g_rw, g_pos);

  float g_r[9];
  if constexpr (j == 0) {
#pragma unroll
    for (int i = 0; i < 9; ++i) g_r[i] = g_rw[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) g_posp[i] = g_pos[i];   // gradient w.r.t. the root position
  } else {
    constexpr int ax = axis_of(j);
    constexpr float sg = sign_of(j);
    const float off = sg * ctx.len[j - 1];
    // pos[j] = pos[p] + off * Rw[j][:, ax]
    ctx.glen[j - 1] = sg * (g_pos[0] * rw[0 * 3 + ax] + g_pos[1] * rw[1 * 3 + ax] + g_pos[2] * rw[2 * 3 + ax]);
#pragma unroll
    for (int row = 0; row < 3; ++row) {
      g_rw[row * 3 + ax] += off * g_pos[row];
      g_posp[row] += g_pos[row];
    }
    // Rw[j] = Rw[p] R[j]:  G_Rw[p] += G_Rw[j] R[j]^T ;  G_R[j] = Rw[p]^T G_Rw[j]
#pragma unroll
    for (int row = 0; row < 3; ++row)
#pragma unroll
      for (int col = 0; col < 3; ++col) {
        g_rwp[row * 3 + col] += g_rw[row * 3 + 0] * r[col * 3 + 0] + g_rw[row * 3 + 1] * r[col * 3 + 1] + g_rw[row * 3 + 2] * r[col * 3 + 2];
        g_r[row * 3 + col] = rwp[0 * 3 + row] * g_rw[0 * 3 + col] + rwp[1 * 3 + row] * g_rw[1 * 3 + col] + rwp[2 * 3 + row] * g_rw[2 * 3 + col];
      }
  }
  if constexpr (RD == 4) {
    // ---- 4-D backward: R = [[s_t, c_t c_p, -c_t s_p], [-c_t, s_t c_p, -s_t s_p], [0, s_p, c_p]]
    const float g_ct = g_r[1] * cp - g_r[2] * sp - g_r[3];
    const float g_st = g_r[0] + g_r[4] * cp - g_r[5] * sp;
    const float g_cp = g_r[1] * ct + g_r[4] * st + g_r[8];
    const float g_sp = -g_r[2] * ct - g_r[5] * st + g_r[7];
    const float dt = clamp_a ? 0.f : ct * g_ct + st * g_st;
    const float dp = clamp_b ? 0.f : cp * g_cp + sp * g_sp;
    ctx.r6[j * 4 + 0] = (g_ct - ct * dt) * int_;
    ctx.r6[j * 4 + 1] = (g_st - st * dt) * int_;
    ctx.r6[j * 4 + 2] = (g_cp - cp * dp) * inp;
    ctx.r6[j * 4 + 3] = (g_sp - sp * dp) * inp;
    return;
  }
  // ---- Gram-Schmidt backward (rotation_tools.py:35-57): columns of G_R are the gradients of x, y, z
  float gx[3] = {g_r[0], g_r[3], g_r[6]};
  const float gy[3] = {g_r[1], g_r[4], g_r[7]};
  float gz[3] = {g_r[2], g_r[5], g_r[8]};
  // y = z x x
  gz[0] += x[1] * gy[2] - x[2] * gy[1];
  gz[1] += x[2] * gy[0] - x[0] * gy[2];
  gz[2] += x[0] * gy[1] - x[1] * gy[0];
  gx[0] += gy[1] * z[2] - gy[2] * z[1];
  gx[1] += gy[2] * z[0] - gy[0] * z[2];
  gx[2] += gy[0] * z[1] - gy[1] * z[0];
  // z = w / max(|w|, eps)
  float gw[3];
  normalize_bwd(z, inw, clamp_b, gz, gw);
  // w = x x b
  const float b[3] = {ctx.r6[j * 6 + 3], ctx.r6[j * 6 + 4], ctx.r6[j * 6 + 5]};
  gx[0] += b[1] * gw[2] - b[2] * gw[1];
  gx[1] += b[2] * gw[0] - b[0] * gw[2];
  gx[2] += b[0] * gw[1] - b[1] * gw[0];
  const float gb[3] = {gw[1] * x[2] - gw[2] * x[1], gw[2] * x[0] - gw[0] * x[2], gw[0] * x[1] - gw[1] * x[0]};
  float ga[3];
  normalize_bwd(x, ina, clamp_a, gx, ga);
  ctx.r6[j * 6 + 0] = ga[0];
  ctx.r6[j * 6 + 1] = ga[1];
  ctx.r6[j * 6 + 2] = ga[2];
  ctx.r6[j * 6 + 3] = gb[0];
  ctx.r6[j * 6 + 4] = gb[1];
  ctx.r6[j * 6 + 5] = gb[2];
}

template <int RD, int j, int c>
struct Children {
  static __device__ __forceinline__ void run(BwdCtx& ctx, const float (&rw)[9], float (&g_rw)[9], float (&g_pos)[3]) {
    if constexpr (parent_of(c) == j) bwd_joint<RD, c>(ctx, rw, g_rw, g_pos);
    Children<RD, j, c + 1>::run(ctx, rw, g_rw, g_pos);
  }
};
template <int RD, int j>
struct Children<RD, j, kJ> {
  static __device__ __forceinline__ void run(BwdCtx&, const float (&)[9], float (&)[9], float (&)[3]) {}
};

constexpr int kBwdWarps = 11;  // 19.6 KB of staging per warp: one CTA of 11 independent warps fills the 227 KB of shared memory
template <int RD>
constexpr int bwd_warp_bytes() { return tile_in_bytes<RD>() + kTileOutBytes; }   // rot / grad_rot tile + grad_poses tile

// Bone-length gradients are summed over a clip's poses in a fixed order: every warp tile writes one partial row (16 bones) per clip it
// touches into glen_part[clip][tile - first tile of the clip], and bone_grad_reduce_kernel adds a clip's rows front to back.
__host__ __device__ __forceinline__ uint32_t glen_slots(uint32_t poses_per_clip) { return (poses_per_clip + kTile - 2) / kTile + 1; }

template <int RD>
__global__ void __launch_bounds__(kBwdWarps * 32, 1)
decoder_bwd_kernel(const float* __restrict__ rot6d, const float* __restrict__ bone_len, const float* __restrict__ grad_poses,
                   float* __restrict__ grad_rot6d, float* __restrict__ glen_part, float* __restrict__ grad_root, uint32_t n_poses,
                   uint32_t poses_per_clip, int bulk_ok) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int kIn = in_floats<RD>();
  constexpr int kTileInBytes = tile_in_bytes<RD>();
  constexpr int kBwdWarpBytes = bwd_warp_bytes<RD>();
  extern __shared__ __align__(128) uint8_t smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* tile = reinterpret_cast<float*>(smem_raw + warp * kBwdWarpBytes);
  float* gtile = tile + kTile * kIn;
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + kBwdWarps * kBwdWarpBytes) + warp;
  if (lane == 0) {
    ptx::mbar_init(bar, 1);
    ptx::fence_mbar_init();
  }
  __syncwarp();

  const uint32_t n_tiles = (n_poses + kTile - 1) / kTile;
  uint32_t phase = 0;
  for (uint32_t tile_idx = warp * gridDim.x + blockIdx.x; tile_idx < n_tiles; tile_idx += gridDim.x * kBwdWarps) {   // few tiles: one per SM first
    const uint32_t pose0 = tile_idx * kTile;
    const uint32_t n_here = min((uint32_t)kTile, n_poses - pose0);
    const bool full = bulk_ok && (n_here == kTile);
    const float* gin = rot6d + (size_t)pose0 * kIn;
    const float* ggp = grad_poses + (size_t)pose0 * kOut;
    float* gout = grad_rot6d + (size_t)pose0 * kIn;
    if (full) {
      if (lane == 0) {
        ptx::mbar_expect_tx(bar, kTileInBytes + kTileOutBytes);
        ptx::bulk_g2s(tile, gin, kTileInBytes, bar);
        ptx::bulk_g2s(gtile, ggp, kTileOutBytes, bar);
      }
      ptx::mbar_wait(bar, phase);
      phase ^= 1;
    } else {
      for (uint32_t i = lane; i < n_here * kIn; i += 32) tile[i] = gin[i];
      for (uint32_t i = lane; i < n_here * kOut; i += 32) gtile[i] = ggp[i];
      __syncwarp();
    }

    const uint32_t pose = pose0 + lane;
    const bool active = lane < n_here;
    const uint32_t clip = (active ? pose : pose0) / poses_per_clip;
    BwdCtx ctx;
    ctx.r6 = tile + lane * kIn;
    ctx.gp = gtile + lane * kOut;
    float len[kBones];
#pragma unroll
    for (int i = 0; i < kBones; ++i) {
      len[i] = __ldg(bone_len + (size_t)clip * kBones + i);
      ctx.glen[i] = 0.f;
    }
    ctx.len = len;
    float g_root[3] = {0.f, 0.f, 0.f};
    if (active) {
      const float ident[9] = {1.f, 0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f, 1.f};
      float g_dummy[9] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      bwd_joint<RD, 0>(ctx, ident, g_dummy, g_root);
      if (grad_root != nullptr) {
        grad_root[(size_t)pose * 3 + 0] = g_root[0];
        grad_root[(size_t)pose * 3 + 1] = g_root[1];
        grad_root[(size_t)pose * 3 + 2] = g_root[2];
      }
    }
    // bone-length gradient: one partial row per clip this tile touches (almost always one), butterfly sums in a fixed order
    const uint32_t clip_first = pose0 / poses_per_clip, clip_last = (pose0 + n_here - 1) / poses_per_clip;
    const uint32_t slots = glen_slots(poses_per_clip);
    for (uint32_t c = clip_first; c <= clip_last; ++c) {
      const bool mine = active && clip == c;
      float* row = glen_part + ((size_t)c * slots + (tile_idx - (c * poses_per_clip) / kTile)) * kBones;
#pragma unroll
      for (int i = 0; i < kBones; ++i) {
        const float sum = warp_sum(mine ? ctx.glen[i] : 0.f);
        if (lane == i) row[i] = sum;
      }
    }

    if (full) {
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::bulk_s2g(gout, tile, kTileInBytes);
        ptx::bulk_commit();
        ptx::bulk_wait_read<0>();
      }
      __syncwarp();
    } else {
      __syncwarp();
      for (uint32_t i = lane; i < n_here * kIn; i += 32) gout[i] = tile[i];
      __syncwarp();
    }
  }
}

// grad_bone_len[clip][bone] = sum of the clip's partial rows, front to back
__global__ void bone_grad_reduce_kernel(const float* __restrict__ glen_part, float* __restrict__ grad_bone_len, uint32_t n_clips,
                                        uint32_t poses_per_clip) {
  pdl_launch_dependents();
  pdl_wait();
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_clips * kBones) return;
  const uint32_t c = idx / kBones, i = idx - c * kBones;
  const uint32_t first = (c * poses_per_clip) / kTile, last = ((c + 1) * poses_per_clip - 1) / kTile;
  const float* row = glen_part + (size_t)c * glen_slots(poses_per_clip) * kBones + i;
  float sum = 0.f;
  for (uint32_t t = first; t <= last; ++t) sum += row[(size_t)(t - first) * kBones];
  grad_bone_len[idx] = sum;
}

}  // namespace
}  // namespace mp

extern "C" {

int mp_decoder_fwd(const float* rot6d, const float* bone_len, const float* root, const float* logits, float* poses,
                   float* scores, int64_t n_clips, int64_t n_hyp, int64_t n_frames, int rot_rep_dim, int flags,
                   mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  // reference: assert rot_rep_dim in [4, 6] (pose_decoder.py:27-30)
  MP_REQUIRE(rot_rep_dim == 4 || rot_rep_dim == 6, MP_EINVAL, "Unsupported rotations representation dimension: %d", rot_rep_dim);
  MP_REQUIRE(n_clips >= 0 && n_hyp >= 1 && n_frames >= 1, MP_EINVAL, "mp_decoder_fwd: bad sizes");
  const int64_t n_poses = n_clips * n_hyp * n_frames;
  if (n_poses == 0) return MP_OK;
  MP_REQUIRE(n_poses < (int64_t)1 << 31, MP_EINVAL, "mp_decoder_fwd: %lld poses exceed 2^31", (long long)n_poses);
  MP_REQUIRE(rot6d && bone_len && poses, MP_EINVAL, "mp_decoder_fwd: null pointer");
  MP_REQUIRE((logits == nullptr) == (scores == nullptr), MP_EINVAL, "mp_decoder_fwd: logits and scores go together");
  MP_REQUIRE(aligned16(bone_len), MP_EALIGN, "mp_decoder_fwd: bone_len must be 16-byte aligned");

  const int bulk_in = aligned16(rot6d), bulk_out = aligned16(poses);   // unaligned views fall back to coalesced LDG / STG
  const size_t smem = (size_t)kWarpsPerCta * (rot_rep_dim == 6 ? tile_in_bytes<6>() : tile_in_bytes<4>()) + kWarpsPerCta * sizeof(uint64_t);
  const int64_t n_tiles = (n_poses + kTile - 1) / kTile;
  int64_t ctas = (n_tiles + kWarpsPerCta - 1) / kWarpsPerCta;
  const int64_t max_ctas = (int64_t)sm_count() * 4;
  if (ctas > max_ctas) ctas = max_ctas;
  auto launch = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(kernel, (unsigned)ctas, kWarpsPerCta * 32, smem, (cudaStream_t)stream, 
        rot6d, bone_len, root, logits, poses, scores, (uint32_t)n_poses, (uint32_t)(n_hyp * n_frames), (uint32_t)n_clips,
        (uint32_t)n_hyp, (uint32_t)n_frames, bulk_in, bulk_out);
  };
  const bool fast = (flags & MP_DEC_FAST) != 0;
  if (rot_rep_dim == 6) {
    if (fast) launch(decoder_fwd_kernel<false, 6>); else launch(decoder_fwd_kernel<true, 6>);
  } else {
    if (fast) launch(decoder_fwd_kernel<false, 4>); else launch(decoder_fwd_kernel<true, 4>);
  }
  return check_launch("decoder_fwd_kernel");
}

size_t mp_decoder_bwd_workspace_bytes(int64_t n_clips, int64_t n_hyp, int64_t n_frames) {
  if (n_clips <= 0 || n_hyp <= 0 || n_frames <= 0) return 0;
  return (size_t)n_clips * mp::glen_slots((uint32_t)(n_hyp * n_frames)) * mp::kBones * sizeof(float);
}

int mp_decoder_bwd(const float* rot6d, const float* bone_len, const float* grad_poses, float* grad_rot6d, float* grad_bone_len,
                   float* grad_root, int64_t n_clips, int64_t n_hyp, int64_t n_frames, int rot_rep_dim, void* workspace,
                   size_t workspace_bytes, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(rot_rep_dim == 4 || rot_rep_dim == 6, MP_EINVAL, "Unsupported rotations representation dimension: %d", rot_rep_dim);
  MP_REQUIRE(n_clips >= 0 && n_hyp >= 1 && n_frames >= 1, MP_EINVAL, "mp_decoder_bwd: bad sizes");
  const int64_t n_poses = n_clips * n_hyp * n_frames;
  if (n_poses == 0) return MP_OK;
  MP_REQUIRE(n_poses < (int64_t)1 << 31, MP_EINVAL, "mp_decoder_bwd: %lld poses exceed 2^31", (long long)n_poses);
  MP_REQUIRE(rot6d && bone_len && grad_poses && grad_rot6d && grad_bone_len, MP_EINVAL, "mp_decoder_bwd: null pointer");
  MP_REQUIRE(workspace && workspace_bytes >= mp_decoder_bwd_workspace_bytes(n_clips, n_hyp, n_frames), MP_EINVAL,
             "mp_decoder_bwd: workspace of %zu bytes, need %zu", workspace_bytes, mp_decoder_bwd_workspace_bytes(n_clips, n_hyp, n_frames));
  const int bulk_ok = aligned16(rot6d) && aligned16(grad_poses) && aligned16(grad_rot6d);
  const size_t smem = (size_t)kBwdWarps * (rot_rep_dim == 6 ? bwd_warp_bytes<6>() : bwd_warp_bytes<4>()) + kBwdWarps * sizeof(uint64_t);
  const int64_t n_tiles = (n_poses + kTile - 1) / kTile;
  // small problems (the training step decodes a few thousand poses): spread the tiles over all SMs rather than 11 to a CTA
  int64_t ctas = std::min<int64_t>(n_tiles, sm_count());
  float* part = static_cast<float*>(workspace);
  auto launch = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(kernel, (unsigned)ctas, kBwdWarps * 32, smem, (cudaStream_t)stream, rot6d, bone_len, grad_poses, grad_rot6d, part, grad_root,
                                                                         (uint32_t)n_poses, (uint32_t)(n_hyp * n_frames), bulk_ok);
  };
  if (rot_rep_dim == 6) launch(decoder_bwd_kernel<6>); else launch(decoder_bwd_kernel<4>);
  MP_CHECK(check_launch("decoder_bwd_kernel"));
  const int64_t n_out = n_clips * kBones;
  launch_k(bone_grad_reduce_kernel, (unsigned)((n_out + 127) / 128), 128, 0, (cudaStream_t)stream, part, grad_bone_len, (uint32_t)n_clips,
                                                                                           (uint32_t)(n_hyp * n_frames));
  return check_launch("bone_grad_reduce_kernel");
}

int mp_softmax_hyp_fwd(const float* logits, float* scores, int64_t n_clips, int64_t n_hyp, int64_t n_frames, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(logits && scores && n_hyp >= 1, MP_EINVAL, "mp_softmax_hyp_fwd: bad arguments");
  const int64_t n = n_clips * n_frames;
  if (n <= 0) return MP_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  launch_k(softmax_hyp_fwd_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)stream, logits, scores, (uint32_t)n_clips, (uint32_t)n_hyp,
                                                                            (uint32_t)n_frames);
  return check_launch("softmax_hyp_fwd_kernel");
}

int mp_softmax_hyp_bwd(const float* scores, const float* grad_scores, float* grad_logits, int64_t n_clips, int64_t n_hyp,
                       int64_t n_frames, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(scores && grad_scores && grad_logits, MP_EINVAL, "mp_softmax_hyp_bwd: null pointer");
  const int64_t n = n_clips * n_frames;
  if (n <= 0) return MP_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  launch_k(softmax_hyp_bwd_kernel, (unsigned)blocks, 256, 0, (cudaStream_t)stream, scores, grad_scores, grad_logits, (uint32_t)n_clips,
                                                                            (uint32_t)n_hyp, (uint32_t)n_frames);
  return check_launch("softmax_hyp_bwd_kernel");
}

}  // extern "C"

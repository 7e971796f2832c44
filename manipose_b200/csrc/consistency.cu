// Pose-consistency metrics of the lifted poses in one pass over data that is already on the device:
//   measure_bones_length        hpe/mh_so3_hpe/metrics/utils.py:4-20            len[b, bone, t] = || p[joint] - p[parent] ||
//   segments_time_consistency   hpe/mh_so3_hpe/metrics/regularizations.py:8-48  var / std over time of every bone length (MPSCE)
//   sagittal_symmetry           regularizations.py:103-140                      | len[left bone] - len[right bone] | (MPSSE)
// The reference evaluates them with ~50 torch launches per call on a permuted [B, 3, J, L] view; here poses stay [B, L, 17, 3].
// Each CTA reduces a slab of frames of one clip (Welford per thread, Chan's parallel combine across threads), a second tiny kernel
// combines the slabs, so one very long sequence (the drivers' "all frames as one clip" MPSCE, main_h36m_lifting.py:948-960) still
// uses the whole GPU.
#include "common.cuh"

namespace mp {
namespace {

constexpr int kCThreads = 128;     // frames per pass of a CTA: 128 x 51 floats are staged in shared memory with coalesced loads
constexpr int kPairs = 6;
constexpr int kStatFloats = 3 * kBones + 2 * kPairs;      // per slab: (n, mean, M2) per bone, (sum |d|, sum d^2) per left/right pair

__host__ __device__ constexpr int left_bone(int i) {
  constexpr int v[kPairs] = {3, 4, 5, 10, 11, 12};
  return v[i];
}
__host__ __device__ constexpr int right_bone(int i) {
  constexpr int v[kPairs] = {0, 1, 2, 13, 14, 15};
  return v[i];
}

// (n, mean, M2) <- combine((n, mean, M2), (nb, mb, M2b))
__device__ __forceinline__ void chan(float& n, float& mean, float& m2, float nb, float mb, float m2b) {
  const float nt = n + nb;
  if (nt > 0.f) {
    const float d = mb - mean;
    const float f = nb / nt;
    mean = fmaf(d, f, mean);
    m2 = m2 + m2b + d * d * n * f;
  }
  n = nt;
}

__global__ void __launch_bounds__(kCThreads)
consistency_partial_kernel(const float* __restrict__ poses, float* __restrict__ bone_len, float* __restrict__ partials, int64_t n_frames,
                           int64_t frames_per_slab) {
  __shared__ float red[kCThreads / 32][kStatFloats];
  const int64_t b = blockIdx.y;
  const int64_t l0 = (int64_t)blockIdx.x * frames_per_slab;
  const int64_t l1 = min(n_frames, l0 + frames_per_slab);
  float cnt = 0.f, mean[kBones], m2[kBones], sa[kPairs], sq[kPairs];
#pragma unroll
  for (int k = 0; k < kBones; ++k) mean[k] = m2[k] = 0.f;
#pragma unroll
  for (int i = 0; i < kPairs; ++i) sa[i] = sq[i] = 0.f;
  __shared__ float tile[kCThreads * kJ * 3];
  for (int64_t lt = l0; lt < l1; lt += kCThreads) {
    const int n_here = (int)min((int64_t)kCThreads, l1 - lt);
    const float* src = poses + (b * n_frames + lt) * (kJ * 3);
    __syncthreads();
    for (int i = threadIdx.x; i < n_here * kJ * 3; i += kCThreads) tile[i] = __ldg(src + i);
    __syncthreads();
    if ((int)threadIdx.x >= n_here) continue;
    const int64_t l = lt + threadIdx.x;
    const float* p = tile + threadIdx.x * (kJ * 3);     // stride 51 words: conflict-free
    float len[kBones];
#pragma unroll
    for (int j = 1; j < kJ; ++j) {
      const int pj = parent_of(j);
      const float dx = __fsub_rn(p[j * 3 + 0], p[pj * 3 + 0]);
      const float dy = __fsub_rn(p[j * 3 + 1], p[pj * 3 + 1]);
      const float dz = __fsub_rn(p[j * 3 + 2], p[pj * 3 + 2]);
      // torch.sum(d ** 2, axis=1).sqrt(): rounded products, sequential sum over the three coordinates, correctly rounded sqrt
      len[j - 1] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    }
    cnt += 1.f;
    const float inv = 1.0f / cnt;
#pragma unroll
    for (int k = 0; k < kBones; ++k) {
      if (bone_len) bone_len[(b * kBones + k) * n_frames + l] = len[k];
      const float d = len[k] - mean[k];
      mean[k] = fmaf(d, inv, mean[k]);
      m2[k] = fmaf(d, len[k] - mean[k], m2[k]);
    }
#pragma unroll
    for (int i = 0; i < kPairs; ++i) {
      const float d = fabsf(len[left_bone(i)] - len[right_bone(i)]);
      sa[i] += d;
      sq[i] = fmaf(d, d, sq[i]);
    }
  }
  // ---- combine across the warp (butterfly), then across warps through shared memory
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float nb = __shfl_xor_sync(0xffffffffu, cnt, o);
#pragma unroll
    for (int k = 0; k < kBones; ++k) {
      const float mb = __shfl_xor_sync(0xffffffffu, mean[k], o), m2b = __shfl_xor_sync(0xffffffffu, m2[k], o);
      float n = cnt;
      chan(n, mean[k], m2[k], nb, mb, m2b);
    }
    cnt += nb;
#pragma unroll
    for (int i = 0; i < kPairs; ++i) {
      sa[i] += __shfl_xor_sync(0xffffffffu, sa[i], o);
      sq[i] += __shfl_xor_sync(0xffffffffu, sq[i], o);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kBones; ++k) {
      red[warp][3 * k + 0] = cnt;
      red[warp][3 * k + 1] = mean[k];
      red[warp][3 * k + 2] = m2[k];
    }
#pragma unroll
    for (int i = 0; i < kPairs; ++i) {
      red[warp][3 * kBones + 2 * i + 0] = sa[i];
      red[warp][3 * kBones + 2 * i + 1] = sq[i];
    }
  }
  __syncthreads();
  float* out = partials + ((size_t)b * gridDim.x + blockIdx.x) * kStatFloats;
  if (threadIdx.x < kBones) {
    const int k = threadIdx.x;
    float n = 0.f, mu = 0.f, s2 = 0.f;
    for (int w = 0; w < kCThreads / 32; ++w) chan(n, mu, s2, red[w][3 * k], red[w][3 * k + 1], red[w][3 * k + 2]);
    out[3 * k + 0] = n;
    out[3 * k + 1] = mu;
    out[3 * k + 2] = s2;
  } else if (threadIdx.x < kBones + 2 * kPairs) {
    const int i = threadIdx.x - kBones;
    float s = 0.f;
    for (int w = 0; w < kCThreads / 32; ++w) s += red[w][3 * kBones + i];
    out[3 * kBones + i] = s;
  }
}

// one CTA per clip: combine the slabs; seg_mean / seg_var (unbiased, torch.var default) [B,16], sym_abs / sym_sq [B,6] (means over time)
__global__ void consistency_finalize_kernel(const float* __restrict__ partials, int n_slabs, int64_t n_frames, float* __restrict__ seg_mean,
                                            float* __restrict__ seg_var, float* __restrict__ sym_abs, float* __restrict__ sym_sq) {
  const int64_t b = blockIdx.x;
  const float* p = partials + (size_t)b * n_slabs * kStatFloats;
  if (threadIdx.x < kBones) {
    const int k = threadIdx.x;
    float n = 0.f, mu = 0.f, s2 = 0.f;
    for (int s = 0; s < n_slabs; ++s) chan(n, mu, s2, p[s * kStatFloats + 3 * k], p[s * kStatFloats + 3 * k + 1], p[s * kStatFloats + 3 * k + 2]);
    seg_mean[b * kBones + k] = mu;
    seg_var[b * kBones + k] = s2 / (float)(n_frames - 1);      // NaN for a single frame, like torch.var
  } else if (threadIdx.x < kBones + kPairs) {
    const int i = threadIdx.x - kBones;
    float a = 0.f, q = 0.f;
    for (int s = 0; s < n_slabs; ++s) {
      a += p[s * kStatFloats + 3 * kBones + 2 * i];
      q += p[s * kStatFloats + 3 * kBones + 2 * i + 1];
    }
    sym_abs[b * kPairs + i] = a / (float)n_frames;
    sym_sq[b * kPairs + i] = q / (float)n_frames;
  }
}

int slabs_for(int64_t n_clips, int64_t n_frames, int64_t* frames_per_slab) {
  // about four CTAs per SM in total, at least one pass of the CTA (256 frames) per slab
  int64_t want = ((int64_t)sm_count() * 4 + n_clips - 1) / (n_clips > 0 ? n_clips : 1);
  const int64_t max_slabs = (n_frames + kCThreads - 1) / kCThreads;
  if (want > max_slabs) want = max_slabs;
  if (want < 1) want = 1;
  *frames_per_slab = (n_frames + want - 1) / want;
  return (int)((n_frames + *frames_per_slab - 1) / *frames_per_slab);
}

}  // namespace
}  // namespace mp

extern "C" {

size_t mp_pose_consistency_workspace_bytes(int64_t n_clips, int64_t n_frames) {
  if (n_clips <= 0 || n_frames <= 0) return 0;
  int64_t fps;
  const int slabs = mp::slabs_for(n_clips, n_frames, &fps);
  return (size_t)n_clips * slabs * mp::kStatFloats * sizeof(float);
}

int mp_pose_consistency(const float* poses, int64_t n_clips, int64_t n_frames, float* seg_mean, float* seg_var, float* sym_abs, float* sym_sq,
                        float* bone_len, void* workspace, size_t workspace_bytes, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(poses && seg_mean && seg_var && sym_abs && sym_sq && n_clips >= 0 && n_frames >= 1 && n_clips < 65536, MP_EINVAL,
             "mp_pose_consistency: bad arguments");
  if (n_clips == 0) return MP_OK;
  int64_t fps;
  const int slabs = slabs_for(n_clips, n_frames, &fps);
  MP_REQUIRE(workspace && workspace_bytes >= (size_t)n_clips * slabs * kStatFloats * sizeof(float), MP_EWORKSPACE,
             "mp_pose_consistency: workspace too small (%zu bytes, see mp_pose_consistency_workspace_bytes)", workspace_bytes);
  float* partials = reinterpret_cast<float*>(workspace);
  consistency_partial_kernel<<<dim3((unsigned)slabs, (unsigned)n_clips), kCThreads, 0, (cudaStream_t)stream>>>(poses, bone_len, partials, n_frames, fps);
  MP_CHECK(check_launch("consistency_partial_kernel"));
  consistency_finalize_kernel<<<(unsigned)n_clips, 32, 0, (cudaStream_t)stream>>>(partials, slabs, n_frames, seg_mean, seg_var, sym_abs, sym_sq);
  return check_launch("consistency_finalize_kernel");
}

}  // extern "C"

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

// -------------------------------------------------------------------------------------------------- P-MPJPE ("Protocol #2")
// hpe/mh_so3_hpe/metrics/mean_joint_errors.py:144-189: per frame, the similarity transform (scale, rotation, translation) that best
// aligns the prediction to the target (orthogonal Procrustes through the SVD of the 3 x 3 cross-covariance, reflections excluded),
// then the mean joint distance.  The reference moves both tensors to the host and calls numpy's batched SVD; here one thread solves
// one frame: Jacobi eigen-decomposition of H^T H (double precision), U = H V S^-1, R = V diag(1, 1, det) U^T.
__device__ __forceinline__ void jacobi_eig3(double (&a)[3][3], double (&v)[3][3]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) v[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
    if (off <= 1e-300 || off <= 1e-18 * (fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]))) break;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      const double apq = a[p][q];
      if (fabs(apq) < 1e-300) continue;
      const double theta = (a[q][q] - a[p][p]) / (2.0 * apq);
      const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
#pragma unroll
      for (int k = 0; k < 3; ++k) {   // A <- A J (columns p, q)
        const double akp = a[k][p], akq = a[k][q];
        a[k][p] = c * akp - sn * akq;
        a[k][q] = sn * akp + c * akq;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {   // A <- J^T A (rows p, q)
        const double apk = a[p][k], aqk = a[q][k];
        a[p][k] = c * apk - sn * aqk;
        a[q][k] = sn * apk + c * aqk;
      }
#pragma unroll
      for (int k = 0; k < 3; ++k) {   // V <- V J
        const double vkp = v[k][p], vkq = v[k][q];
        v[k][p] = c * vkp - sn * vkq;
        v[k][q] = sn * vkp + c * vkq;
      }
    }
  }
}

__global__ void __launch_bounds__(128) p_mpjpe_partial_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int64_t n_frames,
                                                              double* __restrict__ partials) {
  __shared__ double red[4];
  double acc = 0.0;
  for (int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; f < n_frames; f += (int64_t)gridDim.x * blockDim.x) {
    const float* y = pred + f * (kJ * 3);     // "predicted"
    const float* x = gt + f * (kJ * 3);       // "target"
    double mux[3] = {0, 0, 0}, muy[3] = {0, 0, 0};
    for (int j = 0; j < kJ; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        mux[c] += x[j * 3 + c];
        muy[c] += y[j * 3 + c];
      }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      mux[c] /= kJ;
      muy[c] /= kJ;
    }
    double nx = 0, ny = 0, h[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};   // H = X0^T Y0 (unnormalised; divided by nx * ny below)
    for (int j = 0; j < kJ; ++j) {
      double xc[3], yc[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        xc[c] = x[j * 3 + c] - mux[c];
        yc[c] = y[j * 3 + c] - muy[c];
        nx += xc[c] * xc[c];
        ny += yc[c] * yc[c];
      }
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) h[r][c] += xc[r] * yc[c];
    }
    nx = sqrt(nx);
    ny = sqrt(ny);
    const double inv = 1.0 / (nx * ny);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) h[r][c] *= inv;
    // eigen-decomposition of H^T H = V S^2 V^T
    double a[3][3], v[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) a[r][c] = h[0][r] * h[0][c] + h[1][r] * h[1][c] + h[2][r] * h[2][c];
    jacobi_eig3(a, v);
    // order the singular values descending (numpy's convention: the reflection fix flips the LAST one)
    int o0 = 0, o1 = 1, o2 = 2;
    if (a[o0][o0] < a[o1][o1]) { const int t = o0; o0 = o1; o1 = t; }
    if (a[o0][o0] < a[o2][o2]) { const int t = o0; o0 = o2; o2 = t; }
    if (a[o1][o1] < a[o2][o2]) { const int t = o1; o1 = o2; o2 = t; }
    const int ord[3] = {o0, o1, o2};
    double sv[3], vv[3][3], uu[3][3];          // vv[:, i], uu[:, i] = i-th right / left singular vector
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      sv[i] = sqrt(fmax(a[ord[i]][ord[i]], 0.0));
#pragma unroll
      for (int r = 0; r < 3; ++r) vv[r][i] = v[r][ord[i]];
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      double n = 0;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        uu[r][i] = h[r][0] * vv[0][i] + h[r][1] * vv[1][i] + h[r][2] * vv[2][i];
        n += uu[r][i] * uu[r][i];
      }
      n = n > 0 ? 1.0 / sqrt(n) : 0.0;
#pragma unroll
      for (int r = 0; r < 3; ++r) uu[r][i] *= n;
    }
    // third left vector: H v3 / s3 when s3 is resolved, otherwise completed by the cross product (its sign is absorbed by the det fix)
    {
      double n = 0, w[3];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        w[r] = h[r][0] * vv[0][2] + h[r][1] * vv[1][2] + h[r][2] * vv[2][2];
        n += w[r] * w[r];
      }
      if (sv[2] > 1e-9 * sv[0] && n > 0) {
        n = 1.0 / sqrt(n);
#pragma unroll
        for (int r = 0; r < 3; ++r) uu[r][2] = w[r] * n;
      } else {
        uu[0][2] = uu[1][0] * uu[2][1] - uu[2][0] * uu[1][1];
        uu[1][2] = uu[2][0] * uu[0][1] - uu[0][0] * uu[2][1];
        uu[2][2] = uu[0][0] * uu[1][1] - uu[1][0] * uu[0][1];
      }
    }
    auto det3 = [](const double (&m)[3][3]) {
      return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
             m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    };
    const double sgn = det3(vv) * det3(uu) < 0 ? -1.0 : 1.0;    // det(R) with R = V U^T
    const double tr = sv[0] + sv[1] + sgn * sv[2];
    double rot[3][3];                                             // R = V diag(1, 1, sgn) U^T
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int c = 0; c < 3; ++c) rot[r][c] = vv[r][0] * uu[c][0] + vv[r][1] * uu[c][1] + sgn * vv[r][2] * uu[c][2];
    const double scale = tr * nx / ny;
    double t[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) t[c] = mux[c] - scale * (muy[0] * rot[0][c] + muy[1] * rot[1][c] + muy[2] * rot[2][c]);
    double e = 0;
    for (int j = 0; j < kJ; ++j) {
      double d2 = 0;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const double al = scale * (y[j * 3 + 0] * rot[0][c] + y[j * 3 + 1] * rot[1][c] + y[j * 3 + 2] * rot[2][c]) + t[c];
        const double d = al - x[j * 3 + c];
        d2 += d * d;
      }
      e += sqrt(d2);
    }
    acc += e;
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) partials[blockIdx.x] = red[0] + red[1] + red[2] + red[3];
}

__global__ void p_mpjpe_finalize_kernel(const double* __restrict__ partials, int n, double n_points, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double s = 0;
    for (int i = 0; i < n; ++i) s += partials[i];
    out[0] = (float)s;
    out[1] = (float)(s / n_points);
  }
}

// -------------------------------------------------------------------------------------------------- 3DPCK / AUC
// hpe/mh_so3_hpe/metrics/pck.py:77-198 with alignment='none', mask=None: error = ||pred - gt||_2 per joint (fp32, numpy's operation
// order), PCK = 100 * mean(error < threshold), AUC = 100 * mean over thresholds linspace(0, 150, 31) of mean(error < thr).  Counting is
// integer work: every point lands in the histogram bin of the first threshold that exceeds its error; results are exact counts.
constexpr int kAucBins = 32;      // bin j in [1, 31]: first threshold 5 j' > error is j' = j (31: none); bin 0 unused
__global__ void __launch_bounds__(256) pck_hist_kernel(const float* __restrict__ pred, const float* __restrict__ gt, size_t n_points,
                                                       float threshold, unsigned long long* __restrict__ counts) {
  __shared__ unsigned int hist[kAucBins + 1];
  if (threadIdx.x <= kAucBins) hist[threadIdx.x] = 0;
  __syncthreads();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_points; i += (size_t)gridDim.x * blockDim.x) {
    const float d0 = __fsub_rn(pred[i * 3 + 0], gt[i * 3 + 0]), d1 = __fsub_rn(pred[i * 3 + 1], gt[i * 3 + 1]),
                d2 = __fsub_rn(pred[i * 3 + 2], gt[i * 3 + 2]);
    const float e = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2)));
    int j = 31;
#pragma unroll
    for (int t = 30; t >= 0; --t)
      if ((double)e < 5.0 * t) j = t;           // thresholds np.linspace(0, 150, 31) = 5 t exactly; numpy compares in double
    atomicAdd(&hist[j], 1u);
    if ((double)e < (double)threshold) atomicAdd(&hist[kAucBins], 1u);
  }
  __syncthreads();
  if (threadIdx.x <= kAucBins && hist[threadIdx.x]) atomicAdd(counts + threadIdx.x, (unsigned long long)hist[threadIdx.x]);
}
__global__ void pck_finalize_kernel(const unsigned long long* __restrict__ counts, double n_points, float* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    // pck_values[i] = float32(mean(error < 5 i)); AUC = mean of the 31 values (float64) * 100
    double auc = 0.0;
    unsigned long long below = 0;
    for (int i = 0; i <= 30; ++i) {
      below += counts[i];                       // points whose first exceeding threshold index is <= i, i.e. error < 5 i
      auc += (double)(float)((double)below / n_points);
    }
    out[0] = (float)((double)(float)((double)counts[kAucBins] / n_points) * 100.0);
    out[1] = (float)(auc / 31.0 * 100.0);
  }
}

// -------------------------------------------------------------------------------------------------- clip windowing
// PoseSequenceGenerator.__getitem__ (hpe/mh_so3_hpe/data/generators.py:106-219; start frames, occlusion masks and noise are sampled on the
// host with the reference's own RNG calls and handed in) as a device-side
// gather: all sequences live concatenated on the device, window w = table[w] = {first frame of its sequence, sequence length, start frame}
// reads frames start .. start + T - 1 of that sequence, frames past the end replicate the last one (F.pad(mode="replicate"), :132-146).
__global__ void gather_windows_kernel(const float* __restrict__ frames2d, const float* __restrict__ frames3d, const int64_t* __restrict__ table,
                                      const float* __restrict__ mask, const double* __restrict__ noise, const unsigned char* __restrict__ flip,
                                      const int* __restrict__ perm, float* __restrict__ out2d, float* __restrict__ out3d, int64_t n_windows,
                                      int64_t T, int c2, int c3, int in_chans) {
  const int per = c2 + c3;                     // floats per frame: 17 * 2 + 17 * 3
  const int64_t total = n_windows * T * per;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int e = (int)(i % per);
    const int64_t wl = i / per, w = wl / T, l = wl - w * T;
    const int64_t off = table[w * 3 + 0], len = table[w * 3 + 1], start = table[w * 3 + 2];
    int64_t f = start + l;
    if (f > len - 1) f = len - 1;
    // PoseFlip (augmentations/transforms.py:8-31 -> functional.py:7-28) of this window: horizontal coordinate negated, left / right joints swapped
    const bool flipped = flip != nullptr && flip[w] != 0;
    if (e < c2) {
      const int j = e / in_chans, ch = e - j * in_chans;
      float v = frames2d[(off + f) * c2 + (flipped ? perm[j] * in_chans + ch : e)];
      if (flipped && ch == 0) v = -v;
      // miss_type "noisy" (:206-210): `pose_2d += noise` adds a float64 numpy array to a float32 tensor = one rounding of the fp64 sum
      if (noise != nullptr) v = (float)((double)v + noise[wl * c2 + e]);
      // occlusion mask per (frame, joint), broadcast over the coordinates (:212-215)
      if (mask != nullptr) v *= mask[wl * (c2 / in_chans) + e / in_chans];
      out2d[wl * c2 + e] = v;
    } else {
      const int e3 = e - c2, j = e3 / 3, ch = e3 - j * 3;
      const float v = frames3d[(off + f) * c3 + (flipped ? perm[j] * 3 + ch : e3)];
      out3d[wl * c3 + e3] = (flipped && ch == 0) ? -v : v;
    }
  }
}

constexpr int kPmpjpeBlocks = 148 * 8;

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

size_t mp_p_mpjpe_workspace_bytes(int64_t) { return (size_t)mp::kPmpjpeBlocks * sizeof(double); }

int mp_p_mpjpe(const float* pred, const float* gt, int64_t n_frames, float* out, void* workspace, size_t workspace_bytes, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(pred && gt && out && workspace && n_frames >= 1, MP_EINVAL, "mp_p_mpjpe: bad arguments");
  MP_REQUIRE(workspace_bytes >= mp_p_mpjpe_workspace_bytes(n_frames), MP_EWORKSPACE, "mp_p_mpjpe: workspace too small");
  int blocks = (int)((n_frames + 127) / 128);
  if (blocks > kPmpjpeBlocks) blocks = kPmpjpeBlocks;
  double* partials = reinterpret_cast<double*>(workspace);
  p_mpjpe_partial_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(pred, gt, n_frames, partials);
  MP_CHECK(check_launch("p_mpjpe_partial_kernel"));
  p_mpjpe_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(partials, blocks, (double)n_frames * kJ, out);
  return check_launch("p_mpjpe_finalize_kernel");
}

size_t mp_pck_auc_workspace_bytes(void) { return (size_t)(mp::kAucBins + 1) * sizeof(unsigned long long); }

int mp_pck_auc(const float* pred, const float* gt, int64_t n_points, float threshold, float* out, void* workspace, size_t workspace_bytes,
               mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(pred && gt && out && workspace && n_points >= 1, MP_EINVAL, "mp_pck_auc: bad arguments");
  MP_REQUIRE(workspace_bytes >= mp_pck_auc_workspace_bytes(), MP_EWORKSPACE, "mp_pck_auc: workspace too small");
  unsigned long long* counts = reinterpret_cast<unsigned long long*>(workspace);
  cudaError_t e = cudaMemsetAsync(counts, 0, mp_pck_auc_workspace_bytes(), (cudaStream_t)stream);
  MP_REQUIRE(e == cudaSuccess, MP_ELAUNCH, "mp_pck_auc: cudaMemsetAsync: %s", cudaGetErrorString(e));
  int64_t blocks = (n_points + 255) / 256;
  if (blocks > (int64_t)sm_count() * 8) blocks = (int64_t)sm_count() * 8;
  pck_hist_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(pred, gt, (size_t)n_points, threshold, counts);
  MP_CHECK(check_launch("pck_hist_kernel"));
  pck_finalize_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(counts, (double)n_points, out);
  return check_launch("pck_finalize_kernel");
}

int mp_gather_windows(const float* frames2d, const float* frames3d, const int64_t* table, const float* mask, const double* noise,
                      const unsigned char* flip, const int* joint_perm, float* out2d, float* out3d, int64_t n_windows, int64_t n_frames,
                      int n_joints, int in_chans, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(frames2d && frames3d && table && out2d && out3d && n_windows >= 0 && n_frames >= 1 && n_joints >= 1 && in_chans >= 1, MP_EINVAL,
             "mp_gather_windows: bad arguments");
  MP_REQUIRE(flip == nullptr || joint_perm != nullptr, MP_EINVAL, "mp_gather_windows: flip flags need the joint permutation");
  if (n_windows == 0) return MP_OK;
  const int c2 = n_joints * in_chans, c3 = n_joints * 3;
  const int64_t total = n_windows * n_frames * (c2 + c3);
  int64_t blocks = (total + 255) / 256;
  if (blocks > (int64_t)sm_count() * 16) blocks = (int64_t)sm_count() * 16;
  gather_windows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(frames2d, frames3d, table, mask, noise, flip, joint_perm, out2d, out3d,
                                                                          n_windows, n_frames, c2, c3, in_chans);
  return check_launch("gather_windows_kernel");
}

}  // extern "C"

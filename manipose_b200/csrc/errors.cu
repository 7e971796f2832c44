// Joint-wise / coordinate-wise / per-point error analytics of the evaluation drivers:
//   hpe/mh_so3_hpe/metrics/mean_joint_errors.py:31-141  mpjpe_error(no_agg), mse_error, jointwise_error, jointwise_mse, coordwise_error,
//                                                       segments_len_err (on the bone lengths of mp_pose_consistency)
// One streaming pass (24 bytes read per 3-D point, HBM bound) that can emit the per-element values (mode "no_agg") and / or the
// per-column sums (aggregated over every row) at once.  Column sums are DETERMINISTIC: every thread only ever visits elements of one
// column (the grid stride is a multiple of the column count), keeps an fp64 partial, partials are combined per CTA in a fixed order
// and a second kernel adds the CTA partials in a fixed order.
#include "common.cuh"

namespace mp {
namespace {

constexpr int kErrThreads = 256;

// e = the error of element i:  MP_ERR_L2 / MP_ERR_SQ read a 3-D point (||gt - pred||_2 or its square), MP_ERR_ABS / MP_ERR_DIFF a scalar
template <int MODE>
__device__ __forceinline__ float element_error(const float* __restrict__ pred, const float* __restrict__ gt, int64_t i) {
  if (MODE == MP_ERR_L2 || MODE == MP_ERR_SQ) {
    const float dx = gt[3 * i + 0] - pred[3 * i + 0], dy = gt[3 * i + 1] - pred[3 * i + 1], dz = gt[3 * i + 2] - pred[3 * i + 2];
    // torch.sum(d ** 2, dim) / torch.norm(d, 2, dim): every square rounded, then added in order (no FMA contraction), so the
    // element-wise modes reproduce the reference's fp32 values
    const float s = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    return MODE == MP_ERR_L2 ? sqrtf(s) : s;
  }
  const float d = gt[i] - pred[i];
  return MODE == MP_ERR_ABS ? fabsf(d) : d;
}

template <int MODE>
__global__ void __launch_bounds__(kErrThreads)
point_errors_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int64_t n_elem, int cols, float* __restrict__ per_elem,
                    double* __restrict__ partials, int64_t stride) {
  __shared__ double sh[kErrThreads];
  const int64_t t0 = (int64_t)blockIdx.x * kErrThreads + threadIdx.x;
  double acc = 0.0;
  for (int64_t i = t0 < stride ? t0 : n_elem; i < n_elem; i += stride) {      // stride % cols == 0: this thread stays in column t0 % cols
    const float e = element_error<MODE>(pred, gt, i);
    if (per_elem) per_elem[i] = e;
    acc += (double)e;
  }
  if (partials == nullptr) return;
  sh[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < cols) {
    // threads of this CTA whose column is c: (blockIdx.x * kErrThreads + t) % cols == c, visited in increasing t
    const int c = threadIdx.x;
    const int first = (int)(((int64_t)c - (int64_t)blockIdx.x * kErrThreads % cols + cols) % cols);
    double s = 0.0;
    for (int t = first; t < kErrThreads; t += cols) s += sh[t];
    partials[(int64_t)blockIdx.x * cols + c] = s;
  }
}

__global__ void point_errors_finalize_kernel(const double* __restrict__ partials, int n_ctas, int cols, double scale, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  double s = 0.0;
  for (int b = 0; b < n_ctas; ++b) s += partials[(int64_t)b * cols + c];
  out[c] = (float)(s * scale);
}

// Grid and stride: the stride is the largest multiple of `cols` that the grid's threads cover; threads at or beyond it stay idle, so
// every element index i = q * stride + r (r < stride) is visited exactly once, by thread r, whose column r % cols never changes.
int err_grid(int64_t n_elem, int cols, int64_t* stride) {
  int64_t ctas = (n_elem + kErrThreads - 1) / kErrThreads;
  const int64_t cap = (int64_t)sm_count() * 8;
  if (ctas > cap) ctas = cap;
  if (ctas < 1) ctas = 1;
  *stride = ctas * kErrThreads / cols * cols;
  return (int)ctas;
}

}  // namespace
}  // namespace mp

extern "C" {

size_t mp_point_errors_workspace_bytes(int64_t n_elem, int cols) {
  using namespace mp;
  if (n_elem <= 0 || cols <= 0) return 0;
  int64_t stride;
  const int ctas = err_grid(n_elem, cols, &stride);
  return (size_t)ctas * cols * sizeof(double);
}

int mp_point_errors(const float* pred, const float* gt, int64_t n_elem, int cols, int mode, float scale, float* per_elem, float* col_out,
                    void* workspace, size_t workspace_bytes, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(pred && gt && n_elem >= 0, MP_EINVAL, "mp_point_errors: bad arguments");
  MP_REQUIRE(mode >= MP_ERR_L2 && mode <= MP_ERR_DIFF, MP_EINVAL, "mp_point_errors: unknown mode %d", mode);
  MP_REQUIRE(cols >= 1 && cols <= kErrThreads && n_elem % cols == 0, MP_EINVAL, "mp_point_errors: cols=%d must divide n_elem=%lld (1..%d)", cols,
             (long long)n_elem, kErrThreads);
  MP_REQUIRE(per_elem || col_out, MP_EINVAL, "mp_point_errors: nothing to write");
  if (n_elem == 0) {
    if (col_out) cudaMemsetAsync(col_out, 0, cols * sizeof(float), (cudaStream_t)stream);
    return MP_OK;
  }
  int64_t stride;
  const int ctas = err_grid(n_elem, cols, &stride);
  double* partials = nullptr;
  if (col_out) {
    MP_REQUIRE(workspace && workspace_bytes >= (size_t)ctas * cols * sizeof(double) && aligned16(workspace), MP_EWORKSPACE,
               "mp_point_errors: workspace too small (need %zu bytes)", (size_t)ctas * cols * sizeof(double));
    partials = reinterpret_cast<double*>(workspace);
  }
  auto launch = [&](auto kernel) { kernel<<<ctas, kErrThreads, 0, (cudaStream_t)stream>>>(pred, gt, n_elem, cols, per_elem, partials, stride); };
  switch (mode) {
    case MP_ERR_L2: launch(point_errors_kernel<MP_ERR_L2>); break;
    case MP_ERR_SQ: launch(point_errors_kernel<MP_ERR_SQ>); break;
    case MP_ERR_ABS: launch(point_errors_kernel<MP_ERR_ABS>); break;
    default: launch(point_errors_kernel<MP_ERR_DIFF>); break;
  }
  MP_CHECK(check_launch("point_errors_kernel"));
  if (col_out) {
    point_errors_finalize_kernel<<<(cols + 63) / 64, 64, 0, (cudaStream_t)stream>>>(partials, ctas, cols, (double)scale, col_out);
    return check_launch("point_errors_finalize_kernel");
  }
  return MP_OK;
}

}  // extern "C"

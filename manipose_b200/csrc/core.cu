// Error plumbing, device checks, skeleton validation, dtype casts.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

#include "common.cuh"

namespace mp {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MP_ELAUNCH, "%s: %s", what, cudaGetErrorString(e));
  return MP_OK;
}

static int g_checked_dev = -1;
static int g_sm_count = 0;

int require_sm100() {
  int dev = -1;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return fail(MP_EDEVICE, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
  if (dev == g_checked_dev) return MP_OK;
  int major = 0, minor = 0, sms = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (major != 10) return fail(MP_EDEVICE, "device %d is sm_%d%d; libmanipose_sm100 is built for sm_100a only", dev, major, minor);
  g_checked_dev = dev;
  g_sm_count = sms;
  return MP_OK;
}

static int g_sm_limit = 0;   // mp_set_sm_limit: persistent grids are sized from min(SMs of the device, limit)

int sm_count() {
  const int sms = g_sm_count > 0 ? g_sm_count : 148;
  return g_sm_limit > 0 && g_sm_limit < sms ? g_sm_limit : sms;
}

template <typename D>
__global__ void cast_f32_16_kernel(const float* __restrict__ src, typename D::T* __restrict__ dst, int64_t n) {
  pdl_launch_dependents();
  pdl_wait();
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = D::from_float(src[i]);
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("MANIPOSE_PDL");
    return e == nullptr || e[0] != '0';
  }();
  return on;
}

}  // namespace mp

extern "C" {

int mp_abi_version(void) { return MP_ABI_VERSION; }

const char* mp_last_error(void) { return mp::g_err; }

int mp_device_check(void) { return mp::require_sm100(); }

int mp_set_sm_limit(int sms) {
  MP_REQUIRE(sms >= 0 && sms % 2 == 0, MP_EINVAL, "mp_set_sm_limit: %d is not 0 or an even number of SMs", sms);
  const int before = mp::g_sm_limit;
  mp::g_sm_limit = sms;
  return before;
}

int mp_set_skeleton(int num_joints, const int32_t* parents, const float* ops) {
  MP_REQUIRE(parents != nullptr && ops != nullptr, MP_EINVAL, "mp_set_skeleton: null table");
  MP_REQUIRE(num_joints == mp::kJ, MP_EUNSUPPORTED,
             "mp_set_skeleton: kernels are specialised for the 17-joint H36M/3DHP tree, got %d joints", num_joints);
  for (int j = 0; j < mp::kJ; ++j) {
    MP_REQUIRE(parents[j] == mp::parent_of(j), MP_EUNSUPPORTED,
               "mp_set_skeleton: parents[%d] = %d differs from the built-in H36M-17 tree (%d)", j, parents[j], mp::parent_of(j));
    if (j == 0) continue;
    for (int c = 0; c < 3; ++c) {
      float want = (c == mp::axis_of(j)) ? mp::sign_of(j) : 0.f;
      MP_REQUIRE(ops[j * 3 + c] == want, MP_EUNSUPPORTED,
                 "mp_set_skeleton: t_pose_operators[%d][%d] = %g differs from the built-in table (%g)", j, c, ops[j * 3 + c], want);
    }
  }
  return MP_OK;
}

int mp_cast_f32_to_16(const float* src, void* dst, int64_t n, int dtype, mp_stream_t stream) {
  MP_CHECK(mp::require_sm100());
  MP_REQUIRE(n >= 0 && (n == 0 || (src && dst)), MP_EINVAL, "mp_cast_f32_to_16: bad arguments");
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_cast_f32_to_16: unknown dtype %d", dtype);
  if (n == 0) return MP_OK;
  int64_t blocks = (n + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (dtype == MP_DTYPE_BF16)
    launch_k(mp::cast_f32_16_kernel<mp::Bf16>, (unsigned)blocks, 256, 0, (cudaStream_t)stream, src, (__nv_bfloat16*)dst, n);
  else
    launch_k(mp::cast_f32_16_kernel<mp::Fp16>, (unsigned)blocks, 256, 0, (cudaStream_t)stream, src, (__half*)dst, n);
  return mp::check_launch("cast_f32_16_kernel");
}

}  // extern "C"

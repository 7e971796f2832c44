// Shared host/device helpers for libmanipose_sm100.so (sm_100a only).
#pragma once
#include <stdlib.h>

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/manipose_sm100.h"

namespace mp {

// ---- error plumbing (thread-local message, negative codes, never throws/exits) -----------------
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);          // cudaGetLastError() -> MP_ELAUNCH
int require_sm100();                         // current device must be compute capability 10.x
int sm_count();

#define MP_REQUIRE(cond, code, ...)                  \
  do {                                               \
    if (!(cond)) return ::mp::fail((code), __VA_ARGS__); \
  } while (0)

#define MP_CHECK(expr)            \
  do {                            \
    int _mp_rc = (expr);          \
    if (_mp_rc != MP_OK) return _mp_rc; \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- programmatic dependent launch --------------------------------------------------------------------------------------
// Every kernel on the hot paths is launched with cudaLaunchAttributeProgrammaticStreamSerialization (off with MANIPOSE_PDL=0) and
// starts with  pdl_launch_dependents(); <prologue that touches no global data>; pdl_wait();  so that the NEXT kernel's CTAs are
// resident and past their prologue (barrier init, TMEM allocation, descriptor prefetch) while this one drains: the step is a chain of
// hundreds of 5-30 us kernels and the launch-to-launch gap is a measurable share of it.  pdl_wait() returns once the preceding kernel
// has completed and its writes are visible; every kernel calls it before its first global read AND write (a successor may overwrite
// what its predecessor still reads), which also makes completion transitive along the chain.  Without the attribute both are no-ops.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
// Pitch (in floats) of the fp32 intermediate that mp_heads_fwd16 leaves in its workspace and mp_heads_bwd_pack reads back: the K * (D + 1)
// useful columns of the folded projection as a dense matrix (rows padded to 16 bytes); MANIPOSE_HEADS_FULL_STORE=1: whole n_pad-column rows.
inline int heads_ws_ld(int n_hyp, int out_dim, int with_score, int n_pad) {
  static const bool full_store = getenv("MANIPOSE_HEADS_FULL_STORE") != nullptr;
  const int ko = n_hyp * (out_dim + (with_score ? 1 : 0));
  return (n_pad == 128 && !full_store) ? ((ko + 3) & ~3) : n_pad;
}
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- skeleton (H36M-17 / MPI-INF-3DHP tree) as compile-time tables --------------
constexpr int kJ = 17;
constexpr int kBones = 16;
__host__ __device__ constexpr int parent_of(int j) {
  constexpr int p[kJ] = {-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 9, 8, 11, 12, 8, 14, 15};
  return p[j];
}
// t_pose_operators: axis (0 = x, 1 = y) and sign of the unit offset from the parent
__host__ __device__ constexpr int axis_of(int j) {
  constexpr int a[kJ] = {0, 0, 1, 1, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0};
  return a[j];
}
__host__ __device__ constexpr float sign_of(int j) {
  constexpr float s[kJ] = {0.f, 1.f, -1.f, -1.f, -1.f, -1.f, -1.f, 1.f, 1.f, 1.f, 1.f, -1.f, -1.f, -1.f, 1.f, 1.f, 1.f};
  return s[j];
}
__host__ __device__ constexpr bool is_leaf(int j) { return j == 3 || j == 6 || j == 10 || j == 13 || j == 16; }

// ---- small device utilities ------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// ---- the two 16-bit operand formats of the backbone (MP_DTYPE_BF16 / MP_DTYPE_FP16): same tensor-core rate and bytes,
// fp16 carries 3 more mantissa bits (conversions saturate to +-65504 instead of overflowing to inf) --------------------
struct Bf16 {
  using T = __nv_bfloat16;
  static constexpr uint32_t kUmmaFmt = 1;   // tcgen05 instruction-descriptor a/b format: BF16
  static __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  static __device__ __forceinline__ float2 unpack2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
  }
  static __device__ __forceinline__ float round1(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
  static __device__ __forceinline__ T from_float(float x) { return __float2bfloat16_rn(x); }
};
struct Fp16 {
  using T = __half;
  static constexpr uint32_t kUmmaFmt = 0;   // F16
  static __device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
  }
  static __device__ __forceinline__ float2 unpack2(uint32_t u) {
    __half2 v = *reinterpret_cast<__half2*>(&u);
    return __half22float2(v);
  }
  static __device__ __forceinline__ float round1(float x) {
    const uint32_t p = pack2(x, 0.f);
    return unpack2(p).x;
  }
  static __device__ __forceinline__ T from_float(float x) {
    const uint32_t p = pack2(x, 0.f);
    return __ushort_as_half((unsigned short)(p & 0xffffu));
  }
};

// Exact-erf GELU (nn.GELU default, mix_ste.py:200) to 5e-7 absolute: x * Phi(x) with erfc(t) = exp2(-t Q(t)), t = |x| / sqrt(2),
// Q a degree-6 minimax-style fit on [0, 4.6] (erfc(4.6) < 1e-10).  Phi(x) = 1 - erfc(t)/2 for x >= 0 and erfc(t)/2 for x < 0, so the
// negative tail keeps its relative accuracy.  One MUFU (ex2) + ~11 FMA-pipe instructions: the tensor-core epilogue of fc1 has ~16
// issue slots per element before it, not the MMA, paces the tile.
__device__ __forceinline__ float gelu_erf(float x) {
  const float ax = fabsf(x);
  const float t = fminf(ax * 0.70710678118654752440f, 4.6f);
  float q = -9.749186599e-05f;
  q = fmaf(q, t, 4.431374392e-04f);
  q = fmaf(q, t, 2.348781295e-03f);
  q = fmaf(q, t, -2.950778651e-02f);
  q = fmaf(q, t, 1.489954364e-01f);
  q = fmaf(q, t, 9.183205755e-01f);
  q = fmaf(q, t, 1.627914397e+00f);
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-t * q));
  // x Phi(x) = max(x, 0) - (|x| / 2) erfc(|x| / sqrt 2): one form for both signs (no select), 13 instructions per element
  return fmaf(-0.5f * ax, e, fmaxf(x, 0.f));
}

// Two elements at once with the packed fp32 instructions of sm_100 (fma.rn.f32x2 -> FFMA2: two independent IEEE FMAs per issue slot):
// the degree-6 polynomial and the product t Q(t) take 7 instructions per PAIR instead of per element; results are bit-identical to
// gelu_erf.  The tensor-core epilogues of fc1 are instruction-issue bound, so this is what they call.
__device__ __forceinline__ uint64_t pack_f32x2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t mul_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ uint64_t pack_u32x2(uint32_t a, uint32_t b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void unpack_u32x2(uint64_t v, uint32_t& a, uint32_t& b) { asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v)); }
__device__ __forceinline__ void unpack_f32x2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ float sum_f32x2(uint64_t v) {
  float a, b;
  unpack_f32x2(v, a, b);
  return a + b;
}

// 16-bit pair of ((x - mean) * rstd) * gamma + beta for two adjacent elements (the LayerNorm epilogues): three packed instructions
template <typename D>
__device__ __forceinline__ uint32_t pack2_ln(uint64_t x2, uint64_t nmean2, uint64_t rstd2, uint64_t g2, uint64_t b2) {
  float a, b;
  unpack_f32x2(fma_f32x2(mul_f32x2(add_f32x2(x2, nmean2), rstd2), g2, b2), a, b);
  return D::pack2(a, b);
}

__device__ __forceinline__ void gelu_erf2(float& x0, float& x1) {
  const float a0 = fabsf(x0), a1 = fabsf(x1);
  const float t0 = fminf(a0 * 0.70710678118654752440f, 4.6f), t1 = fminf(a1 * 0.70710678118654752440f, 4.6f);
  const uint64_t t = pack_f32x2(t0, t1);
  uint64_t q = pack_f32x2(-9.749186599e-05f, -9.749186599e-05f);
  q = fma_f32x2(q, t, pack_f32x2(4.431374392e-04f, 4.431374392e-04f));
  q = fma_f32x2(q, t, pack_f32x2(2.348781295e-03f, 2.348781295e-03f));
  q = fma_f32x2(q, t, pack_f32x2(-2.950778651e-02f, -2.950778651e-02f));
  q = fma_f32x2(q, t, pack_f32x2(1.489954364e-01f, 1.489954364e-01f));
  q = fma_f32x2(q, t, pack_f32x2(9.183205755e-01f, 9.183205755e-01f));
  q = fma_f32x2(q, t, pack_f32x2(1.627914397e+00f, 1.627914397e+00f));
  uint64_t tq;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(tq) : "l"(q), "l"(t));
  float g0, g1, e0, e1;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(g0), "=f"(g1) : "l"(tq));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(-g0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(-g1));
  x0 = fmaf(-0.5f * a0, e0, fmaxf(x0, 0.f));
  x1 = fmaf(-0.5f * a1, e1, fmaxf(x1, 0.f));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

}  // namespace mp

// nn.Linear on the 5th-generation tensor cores: Y[M,N] = epilogue(A[M,K] @ W[N,K]^T + bias), bf16 in, fp32 accumulate.
//
// Replaces the cuBLAS calls the reference issues through nn.Linear (paths under hpe/mh_so3_hpe/architectures/):
//   mix_ste.py:246,257   attn.qkv   (C -> 3C, bias)
//   mix_ste.py:249,280   attn.proj  (C -> C)   + residual add of Block.forward (mix_ste.py:353-355)
//   mix_ste.py:209-222   mlp.fc1 + exact-erf GELU, mlp.fc2 + residual add (mix_ste.py:356-358)
//
// Structure (one persistent CTA per SM, 256 threads, warp-specialised):
//   warp 0   TMA producer   cp.async.bulk.tensor 2-D tiles of A (128 x 64) and W (BN x 64), 128-byte swizzle, into a
//                           ring of kStages shared-memory stages; completion on `full` mbarriers
//   warp 1   MMA issuer     one thread issues tcgen05.mma (M=128, N=BN, K=16) x 4 per stage, accumulating in TMEM;
//                           tcgen05.commit releases the stage (`empty`) and, after the last k-block, publishes the
//                           accumulator (`tmem_full`)
//   warp 2   TMEM allocator 2 x BN columns = two accumulators, so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 4-7 epilogue      tcgen05.ld (32 lanes x 32 columns) -> +bias, GELU / residual -> per-warp shared-memory
//                           transpose -> 16-byte coalesced bf16 stores; then arrives on `tmem_empty`
// Tiles are ordered n-fastest so CTAs running at the same time share A rows in L2.
#include "common.cuh"
#include "ptx.cuh"

namespace mp {
namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;                       // 64 bf16 = 128 bytes = one swizzle-128B atom row
constexpr int kGemmThreads = 256;
constexpr int kEpiWarps = 4;
constexpr int kStagePad = 36;                 // floats per staged row (32 + 4: conflict-free float4 rows)
constexpr int kEpiStageBytes = kEpiWarps * 32 * kStagePad * 4;   // 18432

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;            // 16 KB
  static constexpr int kBBytes = BN * kBK * 2;             // 32 KB (BN=256) / 16 KB (BN=128)
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kTmemCols = 2 * BN;                 // 512 / 256: power of two >= 32
  static constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + kEpiStageBytes + 256 /*barriers*/;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const float* __restrict__ bias,
                 const __nv_bfloat16* __restrict__ resid, __nv_bfloat16* __restrict__ Y, int M, int N, int K) {
  using Cfg = GemmCfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  float* epi_stage = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + kEpiStageBytes);
  uint64_t* full = bars;                       // [kStages]
  uint64_t* empty = bars + Cfg::kStages;       // [kStages]
  uint64_t* tmem_full = bars + 2 * Cfg::kStages;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;            // [2]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks = N / BN;
  const int m_blocks = (M + kBM - 1) / kBM;
  const int num_tiles = n_blocks * m_blocks;
  const int k_blocks = K / kBK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], kEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_holder, Cfg::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / n_blocks, n_blk = tile - m_blk * n_blocks;
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * Cfg::kStageBytes;
          ptx::mbar_expect_tx(&full[stage], Cfg::kStageBytes);
          ptx::tma_load_2d(sa, &tm_a, &full[stage], kb * kBK, m_blk * kBM);
          ptx::tma_load_2d(sa + Cfg::kABytes, &tm_w, &full[stage], kb * kBK, n_blk * BN);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = ptx::umma_idesc_bf16(kBM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * Cfg::kStageBytes);
          const uint64_t da = ptx::umma_desc_sw128(sa);
          const uint64_t db = ptx::umma_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            // advance 16 bf16 = 32 bytes inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
            ptx::umma_bf16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&empty[stage]);                   // frees the stage when these MMAs have read it
          if (kb == k_blocks - 1) ptx::umma_commit(&tmem_full[acc]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int w = warp - 4;                      // == warp % 4: TMEM lanes [32w, 32w+32)
    float* my_stage = epi_stage + w * 32 * kStagePad;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_blocks, n_blk = tile - m_blk * n_blocks;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_addr = tmem_base + ((uint32_t)(32 * w) << 16) + (uint32_t)(acc * BN);
      const int row0 = m_blk * kBM + 32 * w;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t r[32];
        ptx::tmem_ld32(t_addr + (uint32_t)(c * 32), r);
        ptx::tmem_ld_wait();
        const int col0 = n_blk * BN + c * 32;
        const float4* b4 = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const float4 bb = __ldg(b4 + q);
          float4 v;
          v.x = __uint_as_float(r[4 * q + 0]) + bb.x;
          v.y = __uint_as_float(r[4 * q + 1]) + bb.y;
          v.z = __uint_as_float(r[4 * q + 2]) + bb.z;
          v.w = __uint_as_float(r[4 * q + 3]) + bb.w;
          if (EPI == MP_EPI_GELU) {
            v.x = gelu_erf(v.x);
            v.y = gelu_erf(v.y);
            v.z = gelu_erf(v.z);
            v.w = gelu_erf(v.w);
          }
          *reinterpret_cast<float4*>(my_stage + lane * kStagePad + 4 * q) = v;
        }
        __syncwarp();
        // transposed read-back: 4 lanes cover one row's 32 columns (8 each), 8 rows per pass
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int rr = it * 8 + (lane >> 2), ch = lane & 3;
          const int grow = row0 + rr;
          const float4 lo = *reinterpret_cast<const float4*>(my_stage + rr * kStagePad + ch * 8);
          const float4 hi = *reinterpret_cast<const float4*>(my_stage + rr * kStagePad + ch * 8 + 4);
          float f[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
          if (grow < M) {
            const size_t off = (size_t)grow * N + col0 + ch * 8;
            if (EPI == MP_EPI_RESIDUAL) {
              const uint4 rv = *reinterpret_cast<const uint4*>(resid + off);
              const float2 r0 = unpack_bf16x2(rv.x), r1 = unpack_bf16x2(rv.y), r2 = unpack_bf16x2(rv.z), r3 = unpack_bf16x2(rv.w);
              f[0] += r0.x; f[1] += r0.y; f[2] += r1.x; f[3] += r1.y;
              f[4] += r2.x; f[5] += r2.y; f[6] += r3.x; f[7] += r3.y;
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]);
            o.y = pack_bf16x2(f[2], f[3]);
            o.z = pack_bf16x2(f[4], f[5]);
            o.w = pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(Y + off) = o;
          }
        }
        __syncwarp();
      }
      ptx::tc_fence_before();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// ---- tensor maps ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// row-major bf16 [rows, cols] with row stride ld elements; box = box_rows x 64 columns, 128-byte swizzle, OOB -> 0
int make_tmap(CUtensorMap* map, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  MP_REQUIRE(fn != nullptr, MP_EDEVICE, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MP_REQUIRE(r == CUDA_SUCCESS, MP_EINVAL, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld ld=%lld)", (int)r,
             (long long)rows, (long long)cols, (long long)ld);
  return MP_OK;
}

template <int BN, int EPI>
int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tw, const float* bias, const __nv_bfloat16* resid, __nv_bfloat16* Y, int M,
                int N, int K, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kernel = gemm_bf16_kernel<BN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    MP_REQUIRE(e == cudaSuccess, MP_ELAUNCH, "cudaFuncSetAttribute(gemm): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = (N / BN) * ((M + kBM - 1) / kBM);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  kernel<<<grid, kGemmThreads, Cfg::kSmemBytes, stream>>>(ta, tw, bias, resid, Y, M, N, K);
  return check_launch("gemm_bf16_kernel");
}

template <int BN>
int dispatch_epi(int epilogue, const CUtensorMap& ta, const CUtensorMap& tw, const float* bias, const __nv_bfloat16* resid,
                 __nv_bfloat16* Y, int M, int N, int K, cudaStream_t stream) {
  switch (epilogue) {
    case MP_EPI_BIAS: return launch_gemm<BN, MP_EPI_BIAS>(ta, tw, bias, resid, Y, M, N, K, stream);
    case MP_EPI_GELU: return launch_gemm<BN, MP_EPI_GELU>(ta, tw, bias, resid, Y, M, N, K, stream);
    case MP_EPI_RESIDUAL: return launch_gemm<BN, MP_EPI_RESIDUAL>(ta, tw, bias, resid, Y, M, N, K, stream);
  }
  return fail(MP_EINVAL, "mp_gemm_bf16: unknown epilogue %d", epilogue);
}

}  // namespace
}  // namespace mp

extern "C" int mp_gemm_bf16(const void* A, const void* W, const float* bias, const void* resid, void* Y, int64_t M, int64_t N,
                            int64_t K, int epilogue, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(A && W && bias && Y, MP_EINVAL, "mp_gemm_bf16: null pointer");
  MP_REQUIRE(M >= 0 && M < ((int64_t)1 << 31) && N >= 128 && N % 128 == 0 && K >= 64 && K % 64 == 0, MP_EINVAL,
             "mp_gemm_bf16: unsupported shape M=%lld N=%lld K=%lld (N %% 128 == 0, K %% 64 == 0)", (long long)M, (long long)N, (long long)K);
  MP_REQUIRE(epilogue != MP_EPI_RESIDUAL || resid != nullptr, MP_EINVAL, "mp_gemm_bf16: residual epilogue needs resid");
  MP_REQUIRE(aligned16(A) && aligned16(W) && aligned16(Y) && aligned16(resid) && aligned16(bias), MP_EALIGN,
             "mp_gemm_bf16: pointers must be 16-byte aligned");
  if (M == 0) return MP_OK;
  const bool wide = (N % 256 == 0);
  CUtensorMap ta, tw;
  MP_CHECK(make_tmap(&ta, A, M, K, K, kBM));
  MP_CHECK(make_tmap(&tw, W, N, K, K, wide ? 256 : 128));
  if (wide)
    return dispatch_epi<256>(epilogue, ta, tw, bias, (const __nv_bfloat16*)resid, (__nv_bfloat16*)Y, (int)M, (int)N, (int)K,
                             (cudaStream_t)stream);
  return dispatch_epi<128>(epilogue, ta, tw, bias, (const __nv_bfloat16*)resid, (__nv_bfloat16*)Y, (int)M, (int)N, (int)K,
                           (cudaStream_t)stream);
}

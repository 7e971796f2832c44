// nn.Linear on the 5th-generation tensor cores: Y[M,N] = epilogue(A[M,K] @ W[N,K]^T + bias), 16-bit operands
// (bf16 or fp16), fp32 accumulation in TMEM.
//
// Replaces the cuBLAS calls the reference issues through nn.Linear (paths under hpe/mh_so3_hpe/architectures/):
//   mix_ste.py:246,257   attn.qkv   (C -> 3C, bias)
//   mix_ste.py:249,280   attn.proj  (C -> C)   + residual add of Block.forward (mix_ste.py:353-355)
//   mix_ste.py:209-222   mlp.fc1 + exact-erf GELU, mlp.fc2 + residual add (mix_ste.py:356-358)
//
// Structure (one persistent CTA per SM, 384 threads, warp-specialised; every global access is a TMA tensor copy):
//   warp 0     operand producer  cp.async.bulk.tensor tiles of A (128 x 64) and W (BN x 64), 128-byte swizzle, into a ring of
//                                kStages shared-memory stages; completion on `full` mbarriers
//   warp 1     MMA issuer        one thread issues tcgen05.mma (M=128, N=BN, K=16) x 4 per stage into one of TWO TMEM
//                                accumulators; tcgen05.commit frees the stage and publishes the finished accumulator
//   warp 2     TMEM allocator, then (residual epilogue) the residual loader: streams 128 x 32 fp32 boxes of the residual
//                                into a ring of 4 output slots, running ahead across tile boundaries
//   warps 4-11 epilogue          two groups of 4 warps (TMEM lane quadrant = warp % 4) take alternate 16 KB output boxes:
//                                tcgen05.ld -> +bias (GELU | +residual read from the slot) -> swizzled st.shared into the slot
//                                -> one thread issues the TMA store; the slot is recycled when the store has read it
// Tiles are ordered n-fastest so CTAs running at the same time share A rows in L2.  M tails are zero-filled on load and
// clipped on store by the tensor maps.
#include <stdlib.h>

#include <unordered_map>

#include "common.cuh"
#include "ptx.cuh"

namespace mp {
namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;                       // 64 x 16-bit = 128 bytes = one swizzle-128B atom row
constexpr int kGemmThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int kSlots = 4;                     // output / residual box ring
constexpr int kBoxBytes = kBM * 128;          // 128 rows x 128 bytes

template <int BN>
struct GemmCfg {
  static constexpr int kABytes = kBM * kBK * 2;            // 16 KB
  static constexpr int kBBytes = BN * kBK * 2;             // 32 KB (BN=256) / 16 KB (BN=128)
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN == 256) ? 3 : 4;
  static constexpr int kTmemCols = 2 * BN;                 // 512 / 256: power of two >= 32
  static constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + kSlots * kBoxBytes + 256 /*barriers*/;
};


__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// kMn: both operands are MN-major -- A is [K, M] and W is [K, N] row-major in global memory (the contraction index is the ROW index):
// D[M,N] += A^T W.  Used for weight gradients dW[N_out, K_in] += dY^T X with dY [tokens, N_out], X [tokens, K_in] read in place (no
// transposed copies); tiles are staged as 64-row x 64-column boxes, one per 64-element MN atom, 8 KB apart.
template <int BN, int EPI, typename D, bool kMn = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
linear_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_y,
              const __grid_constant__ CUtensorMap tm_r, const float* __restrict__ bias, int M, int N, int K, int splits) {
  using Cfg = GemmCfg<BN>;
  constexpr bool kRes = (EPI == MP_EPI_RESIDUAL);
  constexpr bool kAcc = (EPI == MP_EPI_ACCUMULATE);        // Y (fp32) += A W^T, split-K: partial tiles are added by TMA reduce stores
  constexpr bool kF32 = (EPI == MP_EPI_BIAS_F32);          // Y (fp32) = A W^T + bias
  constexpr int kBoxCols = (kRes || kAcc || kF32) ? 32 : 64;   // fp32 vs 16-bit output: 128 bytes per row either way
  constexpr int kBoxes = BN / kBoxCols;
  static_assert(kBoxes % 2 == 0, "boxes alternate between the two epilogue groups");

  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint8_t* slot_base = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(slot_base + kSlots * kBoxBytes);
  uint64_t* full = bars;                            // [kStages]  operands landed
  uint64_t* empty = full + Cfg::kStages;            // [kStages]  operands consumed
  uint64_t* tmem_full = empty + Cfg::kStages;       // [2]        accumulator complete
  uint64_t* tmem_empty = tmem_full + 2;             // [2]        accumulator drained
  uint64_t* slot_full = tmem_empty + 2;             // [kSlots]   residual box landed
  uint64_t* slot_empty = slot_full + kSlots;        // [kSlots]   TMA store has read the slot
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(slot_empty + kSlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_blocks = N / BN;
  const int m_blocks = (M + kBM - 1) / kBM;
  const int k_blocks = K / kBK;
  // work items: (tile, k-split).  splits == 1 except for MP_EPI_ACCUMULATE, where few output tiles with a long contraction (weight
  // gradients: the contraction runs over tokens) are spread over all SMs
  const int k_per = (k_blocks + splits - 1) / splits;
  const int num_tiles = n_blocks * m_blocks * splits;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_w);
    ptx::prefetch_tmap(&tm_y);
    if (kRes) ptx::prefetch_tmap(&tm_r);
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
    for (int s = 0; s < kSlots; ++s) {
      ptx::mbar_init(&slot_full[s], 1);
      ptx::mbar_init(&slot_empty[s], 1);
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
  pdl_wait();   // the preceding kernel has completed: operands may be read, outputs written

  if (warp == 0) {
    if (lane == 0) {
      // ===================== operand producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
        const int tile = item / splits, sp = item - tile * splits;
        const int m_blk = tile / n_blocks, n_blk = tile - m_blk * n_blocks;
        const int kb_end = min(k_blocks, (sp + 1) * k_per);
        for (int kb = sp * k_per; kb < kb_end; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * Cfg::kStageBytes;
          ptx::mbar_expect_tx(&full[stage], Cfg::kStageBytes);
          if (kMn) {
#pragma unroll
            for (int a = 0; a < kBM / 64; ++a) ptx::tma_load_2d(sa + a * 8192, &tm_a, &full[stage], m_blk * kBM + a * 64, kb * kBK);
#pragma unroll
            for (int b = 0; b < BN / 64; ++b) ptx::tma_load_2d(sa + Cfg::kABytes + b * 8192, &tm_w, &full[stage], n_blk * BN + b * 64, kb * kBK);
          } else {
            ptx::tma_load_2d(sa, &tm_a, &full[stage], kb * kBK, m_blk * kBM);
            ptx::tma_load_2d(sa + Cfg::kABytes, &tm_w, &full[stage], kb * kBK, n_blk * BN);
          }
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
      constexpr uint32_t idesc = kMn ? ptx::umma_idesc_16_abmn(kBM, BN, D::kUmmaFmt) : ptx::umma_idesc_16(kBM, BN, D::kUmmaFmt);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
        const int sp = item % splits;
        const int kb_begin = sp * k_per, kb_end = min(k_blocks, kb_begin + k_per);
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);   // epilogue has drained this accumulator
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb_begin; kb < kb_end; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * Cfg::kStageBytes);
          if (kMn) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // 16 contraction rows = two 8-row groups of 1024 bytes; MN atoms (64 elements) are 8 KB apart
              const uint64_t da = ptx::umma_desc_mn_sw128_lbo(sa + (uint32_t)(k * 2048), 8192u);
              const uint64_t db = ptx::umma_desc_mn_sw128_lbo(sa + (uint32_t)(Cfg::kABytes + k * 2048), 8192u);
              ptx::umma_f16(d_tmem, da, db, idesc, ((kb - kb_begin) | k) != 0 ? 1u : 0u);
            }
          } else {
            const uint64_t da = ptx::umma_desc_sw128(sa);
            const uint64_t db = ptx::umma_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              // advance 16 elements = 32 bytes inside the 128-byte swizzle atom: +2 in the (addr >> 4) field
              ptx::umma_f16(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, ((kb - kb_begin) | k) != 0 ? 1u : 0u);
            }
          }
          ptx::umma_commit(&empty[stage]);                   // frees the stage when these MMAs have read it
          if (kb == kb_end - 1) ptx::umma_commit(&tmem_full[acc]);
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
  } else if (warp == 2) {
    if (kRes && lane == 0) {
      // ===================== residual loader =====================
      uint32_t n = 0;   // running box index of this CTA
      for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
        const int tile = item / splits;
        const int m_blk = tile / n_blocks, n_blk = tile - m_blk * n_blocks;
        for (int b = 0; b < kBoxes; ++b, ++n) {
          const uint32_t slot = n % kSlots, use = n / kSlots;
          ptx::mbar_wait(&slot_empty[slot], (use & 1) ^ 1);
          ptx::mbar_expect_tx(&slot_full[slot], kBoxBytes);
          ptx::tma_load_2d(slot_base + slot * kBoxBytes, &tm_r, &slot_full[slot], n_blk * BN + b * kBoxCols, m_blk * kBM);
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;                      // TMEM lane quadrant: lanes [32q, 32q+32)
    const int grp = (warp - 4) >> 2;             // 0: warps 4-7 (even boxes), 1: warps 8-11 (odd boxes)
    const bool elected = ((warp - 4) & 3) == 0 && lane == 0;
    const int row = 32 * q + lane;               // row of the 128-row tile owned by this thread
    const uint32_t sw = (uint32_t)(row & 7);     // 128-byte swizzle: 16-byte chunk c of row r lives at chunk c ^ (r & 7)
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t i = 0;                              // running box index of this group
    for (int item = blockIdx.x; item < num_tiles; item += gridDim.x) {
      const int tile = item / splits;
      const int m_blk = tile / n_blocks, n_blk = tile - m_blk * n_blocks;
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int b = grp; b < kBoxes; b += 2, ++i) {
        const uint32_t slot = (uint32_t)grp + 2 * (i & 1), use = i >> 1;
        uint8_t* srow = slot_base + slot * kBoxBytes + row * 128;
        const int col0 = n_blk * BN + b * kBoxCols;
        if (kRes) {
          ptx::mbar_wait(&slot_full[slot], use & 1);             // residual box has landed
        } else {
          ptx::mbar_wait(&slot_empty[slot], (use & 1) ^ 1);      // previous store from this slot has read it
        }
        if (kRes) {
          uint32_t r[32];
          ptx::tmem_ld32(t_row + (uint32_t)(b * 32), r);
          ptx::tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            float4* p = reinterpret_cast<float4*>(srow + (((uint32_t)c ^ sw) << 4));
            const float4 bb = __ldg(b4 + c);
            float4 v = *p;
            v.x += __uint_as_float(r[4 * c + 0]) + bb.x;
            v.y += __uint_as_float(r[4 * c + 1]) + bb.y;
            v.z += __uint_as_float(r[4 * c + 2]) + bb.z;
            v.w += __uint_as_float(r[4 * c + 3]) + bb.w;
            *p = v;
          }
        } else if (kAcc) {
          uint32_t r[32];
          ptx::tmem_ld32(t_row + (uint32_t)(b * 32), r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(srow + (((uint32_t)c ^ sw) << 4)) = make_uint4(r[4 * c + 0], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
        } else if (kF32) {
          uint32_t r[32];
          ptx::tmem_ld32(t_row + (uint32_t)(b * 32), r);
          ptx::tmem_ld_wait();
          const float4* b4 = reinterpret_cast<const float4*>(bias + col0);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 bb = __ldg(b4 + c);
            float4 v;
            v.x = __uint_as_float(r[4 * c + 0]) + bb.x;
            v.y = __uint_as_float(r[4 * c + 1]) + bb.y;
            v.z = __uint_as_float(r[4 * c + 2]) + bb.z;
            v.w = __uint_as_float(r[4 * c + 3]) + bb.w;
            *reinterpret_cast<float4*>(srow + (((uint32_t)c ^ sw) << 4)) = v;
          }
        } else {
          uint32_t r0[32], r1[32];               // both 32-column halves of the box in flight before one wait
          ptx::tmem_ld32(t_row + (uint32_t)(b * 64), r0);
          ptx::tmem_ld32(t_row + (uint32_t)(b * 64 + 32), r1);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t(&r)[32] = half == 0 ? r0 : r1;
            const float4* b4 = reinterpret_cast<const float4*>(bias + col0 + half * 32);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 b0 = __ldg(b4 + 2 * c), b1 = __ldg(b4 + 2 * c + 1);
              float f[8] = {__uint_as_float(r[8 * c + 0]) + b0.x, __uint_as_float(r[8 * c + 1]) + b0.y,
                            __uint_as_float(r[8 * c + 2]) + b0.z, __uint_as_float(r[8 * c + 3]) + b0.w,
                            __uint_as_float(r[8 * c + 4]) + b1.x, __uint_as_float(r[8 * c + 5]) + b1.y,
                            __uint_as_float(r[8 * c + 6]) + b1.z, __uint_as_float(r[8 * c + 7]) + b1.w};
              if (EPI == MP_EPI_GELU) {
#pragma unroll
                for (int e = 0; e < 8; e += 2) gelu_erf2(f[e], f[e + 1]);
              }
              uint4 o;
              o.x = D::pack2(f[0], f[1]);
              o.y = D::pack2(f[2], f[3]);
              o.z = D::pack2(f[4], f[5]);
              o.w = D::pack2(f[6], f[7]);
              *reinterpret_cast<uint4*>(srow + (((uint32_t)(half * 4 + c) ^ sw) << 4)) = o;
            }
          }
        }
        ptx::fence_proxy_async_smem();           // generic-proxy writes -> visible to the TMA store
        named_bar_sync(1 + grp, 128);
        if (elected) {
          if (kAcc)
            ptx::tma_reduce_add_2d(&tm_y, slot_base + slot * kBoxBytes, col0, m_blk * kBM);
          else
            ptx::tma_store_2d(&tm_y, slot_base + slot * kBoxBytes, col0, m_blk * kBM);
          ptx::bulk_commit();
          if (i > 0) {
            ptx::bulk_wait_read<1>();            // the group's previous store has finished reading its slot
            ptx::mbar_arrive(&slot_empty[(uint32_t)grp + 2 * ((i - 1) & 1)]);
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (elected) ptx::bulk_wait<0>();            // all output writes complete before the CTA retires
  }

  __syncwarp();
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace

// ---- tensor maps (shared with gemm2.cu) ----------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

struct TmapKey {
  const void* ptr;
  int64_t rows, cols, pitch;
  int box_rows, type;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && rows == o.rows && cols == o.cols && pitch == o.pitch && box_rows == o.box_rows && type == o.type;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= (size_t)k.rows * 0x9E3779B97F4A7C15ull + (size_t)k.cols * 0xC2B2AE3D27D4EB4Full + (size_t)k.pitch * 0x165667B19E3779F9ull +
         (size_t)k.box_rows * 31 + (size_t)k.type;
    return h;
  }
};

// Dense row-major [rows, cols]; box = box_rows x 128 bytes of columns, 128-byte swizzle, OOB reads -> 0, OOB writes dropped.
// type: 0 = bf16, 1 = fp16, 2 = fp32.  Descriptors are pure functions of the key, so they are memoised per host thread.
// pitch_cols > 0: rows are pitch_cols elements apart and only the first `cols` of them are visible (stores past them are dropped).
int get_tmap_pitch(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int box_rows, int type, int64_t pitch_cols) {
  thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  if (pitch_cols <= 0) pitch_cols = cols;
  const TmapKey key{ptr, rows, cols, pitch_cols, box_rows, type};
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return MP_OK;
  }
  EncodeTiledFn fn = encode_fn();
  MP_REQUIRE(fn != nullptr, MP_EDEVICE, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  const int esz = type == 2 ? 4 : 2;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_cols * esz};
  cuuint32_t box[2] = {(cuuint32_t)(128 / esz), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = type == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (type == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16);
  CUresult r = fn(out, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MP_REQUIRE(r == CUDA_SUCCESS, MP_EINVAL, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%lld type=%d)", (int)r,
             (long long)rows, (long long)cols, type);
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return MP_OK;
}
int get_tmap(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int box_rows, int type) {
  return get_tmap_pitch(out, ptr, rows, cols, box_rows, type, 0);
}

// 4-D view [clips][frames][tokens][cols] of a 16-bit activation (cols fastest) with a box of `box_frames` frames of ONE token and
// 64 columns: the frames of one (clip, token) track land as a dense [box_frames x 128 B] swizzled tile, frames past the end of the clip
// are zero-filled on load and dropped on store.  This is how temporal attention reads [clip, frame, token, C] without any transpose.
int get_tmap_track(CUtensorMap* out, const void* ptr, int64_t n_clips, int64_t n_frames, int64_t n_tok, int64_t cols, int box_frames, int type,
                   int box_tok) {
  EncodeTiledFn fn = encode_fn();
  MP_REQUIRE(fn != nullptr, MP_EDEVICE, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t dims[4] = {(cuuint64_t)cols, (cuuint64_t)n_tok, (cuuint64_t)n_frames, (cuuint64_t)n_clips};
  cuuint64_t strides[3] = {(cuuint64_t)cols * 2, (cuuint64_t)n_tok * cols * 2, (cuuint64_t)n_frames * n_tok * cols * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_tok, (cuuint32_t)box_frames, 1};   // smem rows: token fastest, then frame
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapDataType dt = type == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = fn(out, dt, 4, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MP_REQUIRE(r == CUDA_SUCCESS, MP_EINVAL, "cuTensorMapEncodeTiled(track) failed with CUresult %d (clips=%lld frames=%lld tok=%lld cols=%lld box=%d)",
             (int)r, (long long)n_clips, (long long)n_frames, (long long)n_tok, (long long)cols, box_frames);
  return MP_OK;
}

// 3-D view [clips][rows_per_clip][cols] of a 16-bit activation with a box of `box_rows` rows x 64 columns of ONE clip: rows past the end
// of the clip are zero-filled on load and dropped on store, so a tile never touches another clip (results do not depend on how clips
// are grouped into micro-batches).
int get_tmap_clip_rows(CUtensorMap* out, const void* ptr, int64_t n_clips, int64_t rows_per_clip, int64_t cols, int box_rows, int type) {
  EncodeTiledFn fn = encode_fn();
  MP_REQUIRE(fn != nullptr, MP_EDEVICE, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows_per_clip, (cuuint64_t)n_clips};
  cuuint64_t strides[2] = {(cuuint64_t)cols * 2, (cuuint64_t)rows_per_clip * cols * 2};
  cuuint32_t box[3] = {64, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const CUtensorMapDataType dt = type == 1 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = fn(out, dt, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MP_REQUIRE(r == CUDA_SUCCESS, MP_EINVAL, "cuTensorMapEncodeTiled(clip rows) failed with CUresult %d (clips=%lld rows=%lld cols=%lld box=%d)", (int)r,
             (long long)n_clips, (long long)rows_per_clip, (long long)cols, box_rows);
  return MP_OK;
}

int pair_linear(const void* A, const void* W, const float* bias, void* Y, int M, int N, int K, int epilogue, int dtype, cudaStream_t stream,
                void* Y2 = nullptr);  // gemm2.cu

namespace {

template <int BN, int EPI, typename D>
int launch_linear(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& ty, const CUtensorMap& tr, const float* bias, int M, int N,
                  int K, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kernel = linear_kernel<BN, EPI, D>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    MP_REQUIRE(e == cudaSuccess, MP_ELAUNCH, "cudaFuncSetAttribute(linear_kernel): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = (N / BN) * ((M + kBM - 1) / kBM);
  int splits = 1;
  if (EPI == MP_EPI_ACCUMULATE) {
    // split the contraction so that tiles x splits fills the SMs; every split gets at least 4 k-blocks and none is empty
    const int k_blocks = K / kBK;
    int want = sm_count() / tiles;
    if (want > k_blocks / 4) want = k_blocks / 4;
    if (want < 1) want = 1;
    const int k_per = (k_blocks + want - 1) / want;
    splits = (k_blocks + k_per - 1) / k_per;
  }
  const int items = tiles * splits;
  const int grid = items < sm_count() ? items : sm_count();
  launch_k(kernel, grid, kGemmThreads, Cfg::kSmemBytes, stream, ta, tw, ty, tr, bias, M, N, K, splits);
  return check_launch("linear_kernel");
}

template <int BN, typename D>
int dispatch_epi(int epilogue, const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& ty, const CUtensorMap& tr, const float* bias,
                 int M, int N, int K, cudaStream_t stream) {
  switch (epilogue) {
    case MP_EPI_BIAS: return launch_linear<BN, MP_EPI_BIAS, D>(ta, tw, ty, tr, bias, M, N, K, stream);
    case MP_EPI_GELU: return launch_linear<BN, MP_EPI_GELU, D>(ta, tw, ty, tr, bias, M, N, K, stream);
    case MP_EPI_RESIDUAL: return launch_linear<BN, MP_EPI_RESIDUAL, D>(ta, tw, ty, tr, bias, M, N, K, stream);
    case MP_EPI_ACCUMULATE: return launch_linear<BN, MP_EPI_ACCUMULATE, D>(ta, tw, ty, tr, bias, M, N, K, stream);
    case MP_EPI_BIAS_F32: return launch_linear<BN, MP_EPI_BIAS_F32, D>(ta, tw, ty, tr, bias, M, N, K, stream);
  }
  return fail(MP_EINVAL, "mp_linear: unknown epilogue %d", epilogue);
}

// dW[N_out, K_in] += dY^T X: operands read in place (MN-major), contraction over the tokens split across the SMs
template <int BN, typename D>
int launch_wgrad(const CUtensorMap& ta, const CUtensorMap& tw, const CUtensorMap& ty, int n_out, int k_in, int tokens_pad, cudaStream_t stream) {
  using Cfg = GemmCfg<BN>;
  auto kernel = linear_kernel<BN, MP_EPI_ACCUMULATE, D, true>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    MP_REQUIRE(e == cudaSuccess, MP_ELAUNCH, "cudaFuncSetAttribute(linear_kernel, wgrad): %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int tiles = (k_in / BN) * (n_out / kBM);
  const int k_blocks = tokens_pad / kBK;
  int want = sm_count() / tiles;
  if (want > k_blocks / 4) want = k_blocks / 4;
  if (want < 1) want = 1;
  const int k_per = (k_blocks + want - 1) / want;
  const int splits = (k_blocks + k_per - 1) / k_per;
  const int items = tiles * splits;
  const int grid = items < sm_count() ? items : sm_count();
  launch_k(kernel, grid, kGemmThreads, Cfg::kSmemBytes, stream, ta, tw, ty, ty, (const float*)nullptr, n_out, k_in, tokens_pad, splits);
  return check_launch("linear_kernel (wgrad)");
}

}  // namespace
}  // namespace mp

// Y[M, N] fp32 = A W^T + bias of which only the first `visible_cols` columns of every row are written (the K hypothesis heads: 35 useful
// outputs in a 128-column block; the tensor map of Y drops the rest of each store and packs the rows `pitch_cols` floats apart, so the
// intermediate is a dense [M, pitch_cols] matrix: 144 instead of 512 bytes per token).
namespace mp {
int linear_f32_visible(const void* A, const void* W, const float* bias, float* Y, int64_t M, int64_t N, int64_t K, int64_t visible_cols,
                       int64_t pitch_cols, int dtype, cudaStream_t s) {
  MP_REQUIRE(N == 128 && K >= 64 && K % 64 == 0 && visible_cols >= 1 && visible_cols <= pitch_cols && pitch_cols <= N && pitch_cols % 4 == 0 &&
                 M >= 0 && M < ((int64_t)1 << 31), MP_EINVAL,
             "linear_f32_visible: unsupported shape M=%lld N=%lld K=%lld visible=%lld pitch=%lld", (long long)M, (long long)N, (long long)K,
             (long long)visible_cols, (long long)pitch_cols);
  if (M == 0) return MP_OK;
  CUtensorMap ta, tw, ty;
  MP_CHECK(get_tmap(&ta, A, M, K, kBM, dtype));
  MP_CHECK(get_tmap(&tw, W, N, K, 128, dtype));
  MP_CHECK(get_tmap_pitch(&ty, Y, M, visible_cols, kBM, 2, pitch_cols));
  if (dtype == MP_DTYPE_BF16) return launch_linear<128, MP_EPI_BIAS_F32, Bf16>(ta, tw, ty, ty, bias, (int)M, (int)N, (int)K, s);
  return launch_linear<128, MP_EPI_BIAS_F32, Fp16>(ta, tw, ty, ty, bias, (int)M, (int)N, (int)K, s);
}
}  // namespace mp

extern "C" int mp_linear(const void* A, const void* W, const float* bias, const float* resid, void* Y, int64_t M, int64_t N, int64_t K,
                         int epilogue, int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(A && W && Y && (bias || epilogue == MP_EPI_ACCUMULATE), MP_EINVAL, "mp_linear: null pointer");
  MP_REQUIRE(M >= 0 && M < ((int64_t)1 << 31) && N >= 128 && N % 128 == 0 && K >= 64 && K % 64 == 0, MP_EINVAL,
             "mp_linear: unsupported shape M=%lld N=%lld K=%lld (N %% 128 == 0, K %% 64 == 0)", (long long)M, (long long)N, (long long)K);
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_linear: unknown dtype %d", dtype);
  MP_REQUIRE(epilogue != MP_EPI_RESIDUAL || resid != nullptr, MP_EINVAL, "mp_linear: residual epilogue needs resid");
  MP_REQUIRE(aligned16(A) && aligned16(W) && aligned16(Y) && aligned16(resid) && aligned16(bias), MP_EALIGN,
             "mp_linear: pointers must be 16-byte aligned");
  if (M == 0) return MP_OK;
  const bool wide = (N % 256 == 0);
  const bool res = epilogue == MP_EPI_RESIDUAL;
  const bool f32_out = res || epilogue == MP_EPI_ACCUMULATE || epilogue == MP_EPI_BIAS_F32;
  // CTA pairs halve the weight traffic out of L2; MANIPOSE_SINGLE_CTA=1 keeps the one-CTA kernel (A/B measurements)
  static const bool single_only = getenv("MANIPOSE_SINGLE_CTA") != nullptr;
  if (wide && !f32_out && !single_only) return pair_linear(A, W, bias, Y, (int)M, (int)N, (int)K, epilogue, dtype, (cudaStream_t)stream);
  CUtensorMap ta, tw, ty, tr;
  MP_CHECK(get_tmap(&ta, A, M, K, kBM, dtype));
  MP_CHECK(get_tmap(&tw, W, N, K, wide ? 256 : 128, dtype));
  MP_CHECK(get_tmap(&ty, Y, M, N, kBM, f32_out ? 2 : dtype));
  if (res)
    MP_CHECK(get_tmap(&tr, resid, M, N, kBM, 2));
  else
    tr = ty;
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == MP_DTYPE_BF16) {
    if (wide) return dispatch_epi<256, Bf16>(epilogue, ta, tw, ty, tr, bias, (int)M, (int)N, (int)K, s);
    return dispatch_epi<128, Bf16>(epilogue, ta, tw, ty, tr, bias, (int)M, (int)N, (int)K, s);
  }
  if (wide) return dispatch_epi<256, Fp16>(epilogue, ta, tw, ty, tr, bias, (int)M, (int)N, (int)K, s);
  return dispatch_epi<128, Fp16>(epilogue, ta, tw, ty, tr, bias, (int)M, (int)N, (int)K, s);
}

extern "C" int mp_wgrad(const void* dY, const void* X, float* dW, int64_t n_tokens, int64_t n_out, int64_t k_in, int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(dY && X && dW, MP_EINVAL, "mp_wgrad: null pointer");
  MP_REQUIRE(n_tokens >= 0 && n_tokens < ((int64_t)1 << 31) && n_out >= 128 && n_out % 128 == 0 && k_in >= 128 && k_in % 128 == 0, MP_EINVAL,
             "mp_wgrad: unsupported shape tokens=%lld n_out=%lld k_in=%lld (n_out %% 128 == 0, k_in %% 128 == 0)", (long long)n_tokens,
             (long long)n_out, (long long)k_in);
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_wgrad: unknown dtype %d", dtype);
  MP_REQUIRE(aligned16(dY) && aligned16(X) && aligned16(dW), MP_EALIGN, "mp_wgrad: pointers must be 16-byte aligned");
  if (n_tokens == 0) return MP_OK;
  CUtensorMap ta, tw, ty;
  MP_CHECK(get_tmap(&ta, dY, n_tokens, n_out, 64, dtype));      // 64 tokens x 64 columns per box; token tails are zero-filled
  MP_CHECK(get_tmap(&tw, X, n_tokens, k_in, 64, dtype));
  MP_CHECK(get_tmap(&ty, dW, n_out, k_in, kBM, 2));
  const int tokens_pad = (int)((n_tokens + kBK - 1) / kBK * kBK);
  cudaStream_t s = (cudaStream_t)stream;
  const bool wide = k_in % 256 == 0;
  if (dtype == MP_DTYPE_BF16)
    return wide ? launch_wgrad<256, Bf16>(ta, tw, ty, (int)n_out, (int)k_in, tokens_pad, s) : launch_wgrad<128, Bf16>(ta, tw, ty, (int)n_out, (int)k_in, tokens_pad, s);
  return wide ? launch_wgrad<256, Fp16>(ta, tw, ty, (int)n_out, (int)k_in, tokens_pad, s) : launch_wgrad<128, Fp16>(ta, tw, ty, (int)n_out, (int)k_in, tokens_pad, s);
}

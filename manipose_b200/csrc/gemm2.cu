// CTA-pair (cta_group::2) tensor-core kernels for the C = 512 rotation backbone: two SMs of one TPC compute a 256-row tile
// together, each CTA feeding 128 rows of A and HALF of the weight rows, so every weight byte is pulled out of L2 once per pair
// instead of once per CTA (the single-CTA kernel in gemm.cu is L2-bandwidth bound at ~11 TB/s of operand traffic).
//
//   pair_linear_kernel     Y[M,N] (16-bit) = act(A W^T + b), N % 256 == 0      256 x 256 tile, two TMEM accumulators
//                          (qkv: mix_ste.py:246,257; fc1 + GELU: mix_ste.py:209-222)
//   pair_linear_ln_kernel  N = 512 = the whole row: x = resid + A W^T + b      256 x 512 tile, one TMEM accumulator
//                          (proj / fc2 + residual add of Block.forward, mix_ste.py:352-358), optionally followed IN THE EPILOGUE by
//                            x = LN_post(x) (+ temporal pos-embed)             shared Spatial_norm / Temporal_norm, mix_ste.py:143,149,154,166,170
//                            h = LN_pre(x) as 16-bit                            norm2 of this block / norm1 of the next, mix_ste.py:353,356
//                          so the LayerNorm kernels and their extra pass over the residual stream disappear.
//
// Roles per CTA (384 threads): warp 0 operand producer (TMA), warp 1 MMA issuer (leader CTA only), warp 2 TMEM allocator + residual
// box loader, warps 4-11 epilogue (TMEM lane quadrant = warp % 4; column half / alternate boxes = (warp - 4) / 4).  Barriers that the
// leader's MMA thread waits on (`full`, `tmem_empty`) live in the leader CTA and are arrived on remotely by the peer; barriers the
// MMA thread signals (`empty`, `tmem_full`) are multicast tcgen05.commit arrivals into both CTAs.
#include <stdlib.h>

#include "common.cuh"
#include "ln_args.cuh"
#include "ptx.cuh"

namespace mp {

int get_tmap(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int box_rows, int type);   // gemm.cu

namespace {

constexpr int kBM = 128;                      // rows per CTA (256 per pair)
constexpr int kBK = 64;
constexpr int kThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int kSlots = 4;
constexpr int kBoxBytes = kBM * 128;
constexpr int kABytes = kBM * kBK * 2;        // 16 KB
constexpr int kWHalfBytes = 128 * kBK * 2;    // 16 KB: the 128 weight rows this CTA contributes to one N = 256 instruction
constexpr int kEpiGelu2 = 100;                // internal epilogue code of pair_linear_kernel: pre-activation AND GELU outputs (mp_linear_gelu2)

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct PairBars {
  uint64_t* full;        // [stages]  leader: operands of both CTAs landed (one arrival that expects the bytes of BOTH CTAs)
  uint64_t* empty;       // [stages]  each CTA: stage consumed (multicast commit)
  uint64_t* tmem_full;   // [2]       each CTA: accumulator complete (multicast commit)
  uint64_t* tmem_empty;  // [2]       leader: accumulator drained by the 16 epilogue warps of the pair
  uint64_t* slot_full;   // [kSlots]  residual box landed
  uint64_t* slot_empty;  // [kSlots]  slot free again
  uint64_t* tile_done;   // [1]       both column halves have finished a tile's epilogue (every slot is free)
};

template <int kStages, int kNumSlots>
__device__ __forceinline__ PairBars carve_bars(uint8_t* p) {
  PairBars b;
  b.full = reinterpret_cast<uint64_t*>(p);
  b.empty = b.full + kStages;
  b.tmem_full = b.empty + kStages;
  b.tmem_empty = b.tmem_full + 2;
  b.slot_full = b.tmem_empty + 2;
  b.slot_empty = b.slot_full + kNumSlots;
  b.tile_done = b.slot_empty + kNumSlots;
  return b;
}

template <int kStages, int kNumSlots>
__device__ __forceinline__ void init_bars(const PairBars& b) {
  for (int s = 0; s < kStages; ++s) {
    ptx::mbar_init(&b.full[s], 1);
    ptx::mbar_init(&b.empty[s], 1);
  }
  for (int a = 0; a < 2; ++a) {
    ptx::mbar_init(&b.tmem_full[a], 1);
    ptx::mbar_init(&b.tmem_empty[a], 2 * kEpiWarps);
  }
  for (int s = 0; s < kNumSlots; ++s) {
    ptx::mbar_init(&b.slot_full[s], 1);
    ptx::mbar_init(&b.slot_empty[s], 1);
  }
  ptx::mbar_init(b.tile_done, 2);
  ptx::fence_mbar_init();
}

// ===================================================================================== Y = act(A W^T + b), 16-bit out
template <int EPI, typename D>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
pair_linear_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_y,
                   const __grid_constant__ CUtensorMap tm_y2, const float* __restrict__ bias, int M, int N, int K) {
  // EPI = kEpiGelu2: two outputs from one accumulator, Y = A W^T + b (the pre-activation the GELU backward needs) and Y2 = GELU(A W^T + b)
  // (the operand of fc2), GELU taken on the fp32 value like the inference epilogue - the training forward's fc1 + gelu_kernel pair.
  constexpr int kPasses = EPI == kEpiGelu2 ? 2 : 1;
  constexpr int kStages = 5;
  constexpr int kStageBytes = kABytes + kWHalfBytes;     // 32 KB per CTA per stage
  constexpr int BN = 256;
  constexpr int kBoxes = 4;                              // 64-column 16-bit boxes per tile

  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_base = smem;
  uint8_t* slot_base = smem + kStages * kStageBytes;
  const PairBars bars = carve_bars<kStages, kSlots>(slot_base + kSlots * kBoxBytes);
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars.tile_done + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int n_blocks = N / BN;
  const int m_pairs = (M + 2 * kBM - 1) / (2 * kBM);
  const int num_tiles = n_blocks * m_pairs;
  const int k_blocks = K / kBK;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_w);
    ptx::prefetch_tmap(&tm_y);
    if (EPI == kEpiGelu2) ptx::prefetch_tmap(&tm_y2);
  }
  if (warp == 1 && lane == 0) init_bars<kStages, kSlots>(bars);
  if (warp == 2) {
    ptx::tmem_alloc(tmem_holder, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();   // the preceding kernel has completed: operands may be read, outputs written

  if (warp == 0) {
    if (lane == 0) {
      // ===================== operand producer (both CTAs) =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
        const int m_pair = tile / n_blocks, n_blk = tile - m_pair * n_blocks;
        const int row_a = m_pair * 2 * kBM + (int)rank * kBM;
        const int row_w = n_blk * BN + (int)rank * 128;
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&bars.empty[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * kStageBytes;
          const uint32_t full_leader = ptx::mapa_shared(smem_u32(&bars.full[stage]), 0);
          if (leader) ptx::mbar_expect_tx(&bars.full[stage], 2 * kStageBytes);   // the peer's loads complete_tx on this barrier too
          ptx::tma_load_2d_pair(sa, &tm_a, full_leader, kb * kBK, row_a);
          ptx::tma_load_2d_pair(sa + kABytes, &tm_w, full_leader, kb * kBK, row_w);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      // ===================== MMA issuer (leader CTA) =====================
      constexpr uint32_t idesc = ptx::umma_idesc_16(2 * kBM, BN, D::kUmmaFmt);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
        ptx::mbar_wait(&bars.tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&bars.full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * kStageBytes);
          const uint64_t da = ptx::umma_desc_sw128(sa);
          const uint64_t db = ptx::umma_desc_sw128(sa + kABytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            ptx::umma_f16_pair(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::umma_commit_pair(&bars.empty[stage]);
          if (kb == k_blocks - 1) ptx::umma_commit_pair(&bars.tmem_full[acc]);
          if (++stage == kStages) {
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
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const bool elected = ((warp - 4) & 3) == 0 && lane == 0;
    const int row = 32 * q + lane;
    const uint32_t sw = (uint32_t)(row & 7);
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t i = 0;
    for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
      const int m_pair = tile / n_blocks, n_blk = tile - m_pair * n_blocks;
      const int row0 = m_pair * 2 * kBM + (int)rank * kBM;
      ptx::mbar_wait(&bars.tmem_full[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
      for (int b = grp; b < kBoxes; b += 2) {
        const int col0 = n_blk * BN + b * 64;
        uint32_t r0[32], r1[32];               // both 32-column halves of the box in flight before one wait
        ptx::tmem_ld32(t_row + (uint32_t)(b * 64), r0);
        ptx::tmem_ld32(t_row + (uint32_t)(b * 64 + 32), r1);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int pass = 0; pass < kPasses; ++pass, ++i) {
          const uint32_t slot = (uint32_t)grp + 2 * (i & 1), use = i >> 1;
          uint8_t* srow = slot_base + slot * kBoxBytes + row * 128;
          ptx::mbar_wait(&bars.slot_empty[slot], (use & 1) ^ 1);
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
              if (EPI == MP_EPI_GELU || (EPI == kEpiGelu2 && pass == 1)) {
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
          ptx::fence_proxy_async_smem();
          named_bar_sync(1 + grp, 128);
          if (elected) {
            ptx::tma_store_2d(pass == 0 ? &tm_y : &tm_y2, slot_base + slot * kBoxBytes, col0, row0);
            ptx::bulk_commit();
            if (i > 0) {
              ptx::bulk_wait_read<1>();
              ptx::mbar_arrive(&bars.slot_empty[(uint32_t)grp + 2 * ((i - 1) & 1)]);
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_shared(smem_u32(&bars.tmem_empty[acc]), 0));
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (elected) ptx::bulk_wait<0>();
  }

  __syncwarp();
  ptx::tc_fence_before();
  ptx::cluster_sync_all();     // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ===================================================================================== Y = act(A W^T + b), A-stationary (K = 512)
// pair_linear_kernel pulls A through L2 once per 256-column block of N (6 times for qkv, 4 for fc1) and W once per 256-row tile:
// 64 bytes per SM-clock at the MMA rate, which is what the L2 slices deliver chip-wide (~6300 B/clk), so those launches run at the
// L2 roof (qkv: 12 TB/s of operand traffic under ncu), not at the tensor roof.  With K = 512 the whole A tile of a CTA is
// 128 x 512 x 2 B = 128 KB: this variant keeps it RESIDENT in shared memory for all the N blocks of a 256-row tile (8 k-block
// buffers with their own full / empty barriers; a buffer is handed back during the last N block and refilled with the next tile's
// k-block right away) and streams only W (4 x 16 KB stages).  Operand traffic per token drops from N/256 + N/256 KB to 1 + N/256 KB
// (qkv 12 -> 7, fc1 8 -> 5).  What is left of shared memory holds ONE 16 KB output slot per epilogue group, so the epilogue computes a
// box into registers first and only then waits for its slot.
constexpr int kAsKB = 8;                                   // k-blocks of the resident A tile (K = 512)
constexpr int kAsSlots = 2;
constexpr int as_smem(int w_stages) { return 1024 + kAsKB * kABytes + w_stages * kWHalfBytes + kAsSlots * kBoxBytes + 512; }
static_assert(as_smem(4) <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");

template <int EPI, typename D, int kAsWStages, bool kRotate>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
pair_linear_as_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_y,
                      const float* __restrict__ bias, int M, int N) {
  constexpr int BN = 256;
  constexpr int kBoxes = 4;

  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* a_base = smem;
  uint8_t* w_base = a_base + kAsKB * kABytes;
  uint8_t* slot_base = w_base + kAsWStages * kWHalfBytes;
  uint64_t* w_full = reinterpret_cast<uint64_t*>(slot_base + kAsSlots * kBoxBytes);   // [stages] leader: W halves of both CTAs landed
  uint64_t* w_empty = w_full + kAsWStages;                 // [stages] each CTA: stage consumed (multicast commit)
  uint64_t* a_full = w_empty + kAsWStages;                 // [8]      leader: A k-block of both CTAs landed
  uint64_t* a_empty = a_full + kAsKB;                      // [8]      each CTA: last N block has consumed the k-block (multicast commit)
  uint64_t* tmem_full = a_empty + kAsKB;                   // [2]
  uint64_t* tmem_empty = tmem_full + 2;                    // [2]      leader: drained by the 16 epilogue warps of the pair
  uint64_t* slot_empty = tmem_empty + 2;                   // [2]      the group's TMA store has read its slot
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(slot_empty + kAsSlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int n_blocks = N / BN;
  const int m_pairs = (M + 2 * kBM - 1) / (2 * kBM);
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_w);
    ptx::prefetch_tmap(&tm_y);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kAsWStages; ++s) {
      ptx::mbar_init(&w_full[s], 1);
      ptx::mbar_init(&w_empty[s], 1);
    }
    for (int s = 0; s < kAsKB; ++s) {
      ptx::mbar_init(&a_full[s], 1);
      ptx::mbar_init(&a_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 2 * kEpiWarps);
    }
    for (int s = 0; s < kAsSlots; ++s) ptx::mbar_init(&slot_empty[s], 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_holder, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();   // the preceding kernel has completed: operands may be read, outputs written

  if (warp == 0) {
    if (lane == 0) {
      // ===================== W producer (both CTAs: 128 of the 256 weight rows each) =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int m_pair = pair_id; m_pair < m_pairs; m_pair += num_pairs) {
        for (int n_it = 0; n_it < n_blocks; ++n_it) {
          const int n_blk = kRotate ? (n_it + pair_id) % n_blocks : n_it;   // pairs walk the N blocks in rotated order: the 74 pairs do
          const int row_w = n_blk * BN + (int)rank * 128;                   // not all pull the same W lines out of L2 at the same time
          for (int kb = 0; kb < kAsKB; ++kb) {
            ptx::mbar_wait(&w_empty[stage], phase ^ 1);
            const uint32_t full_leader = ptx::mapa_shared(smem_u32(&w_full[stage]), 0);
            if (leader) ptx::mbar_expect_tx(&w_full[stage], 2 * kWHalfBytes);   // the peer's load completes on this barrier too
            ptx::tma_load_2d_pair(w_base + stage * kWHalfBytes, &tm_w, full_leader, kb * kBK, row_w);
            if (++stage == kAsWStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 3) {
    if (lane == 0) {
      // ===================== A producer (both CTAs: own 128 rows, once per 256-row tile) =====================
      uint32_t a_phase = 0;
      for (int m_pair = pair_id; m_pair < m_pairs; m_pair += num_pairs, a_phase ^= 1) {
        const int row_a = m_pair * 2 * kBM + (int)rank * kBM;
        for (int kb = 0; kb < kAsKB; ++kb) {
          ptx::mbar_wait(&a_empty[kb], a_phase ^ 1);       // the previous tile's last N block is done with this buffer
          const uint32_t full_leader = ptx::mapa_shared(smem_u32(&a_full[kb]), 0);
          if (leader) ptx::mbar_expect_tx(&a_full[kb], 2 * kABytes);
          ptx::tma_load_2d_pair(a_base + kb * kABytes, &tm_a, full_leader, kb * kBK, row_a);
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      // ===================== MMA issuer (leader CTA) =====================
      constexpr uint32_t idesc = ptx::umma_idesc_16(2 * kBM, BN, D::kUmmaFmt);
      int stage = 0;
      uint32_t phase = 0, a_phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int m_pair = pair_id; m_pair < m_pairs; m_pair += num_pairs, a_phase ^= 1) {
        for (int n_blk = 0; n_blk < n_blocks; ++n_blk) {
          ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
          for (int kb = 0; kb < kAsKB; ++kb) {
            if (n_blk == 0) ptx::mbar_wait(&a_full[kb], a_phase);
            ptx::mbar_wait(&w_full[stage], phase);
            ptx::tc_fence_after();
            const uint64_t da = ptx::umma_desc_sw128(smem_u32(a_base + kb * kABytes));
            const uint64_t db = ptx::umma_desc_sw128(smem_u32(w_base + stage * kWHalfBytes));
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              ptx::umma_f16_pair(d_tmem, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            ptx::umma_commit_pair(&w_empty[stage]);
            if (n_blk == n_blocks - 1) ptx::umma_commit_pair(&a_empty[kb]);
            if (kb == kAsKB - 1) ptx::umma_commit_pair(&tmem_full[acc]);
            if (++stage == kAsWStages) {
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
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs, own 128 rows) =====================
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const bool elected = ((warp - 4) & 3) == 0 && lane == 0;
    const int row = 32 * q + lane;
    const uint32_t sw = (uint32_t)(row & 7);
    uint8_t* srow = slot_base + grp * kBoxBytes + row * 128;
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t use = 0;                            // boxes this group has stored so far
    for (int m_pair = pair_id; m_pair < m_pairs; m_pair += num_pairs) {
      const int row0 = m_pair * 2 * kBM + (int)rank * kBM;
      for (int n_it = 0; n_it < n_blocks; ++n_it) {
        const int n_blk = kRotate ? (n_it + pair_id) % n_blocks : n_it;
        ptx::mbar_wait(&tmem_full[acc], acc_phase);
        ptx::tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * BN);
#pragma unroll 1
        for (int b = grp; b < kBoxes; b += 2, ++use) {
          const int col0 = n_blk * BN + b * 64;
          uint32_t r0[32], r1[32];
          ptx::tmem_ld32(t_row + (uint32_t)(b * 64), r0);
          ptx::tmem_ld32(t_row + (uint32_t)(b * 64 + 32), r1);
          ptx::tmem_ld_wait();
          if (b + 2 >= kBoxes) {                 // last TMEM read of this accumulator by this warp: hand it back before the math
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_shared(smem_u32(&tmem_empty[acc]), 0));
          }
          uint4 o[8];
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
              o[half * 4 + c].x = D::pack2(f[0], f[1]);
              o[half * 4 + c].y = D::pack2(f[2], f[3]);
              o[half * 4 + c].z = D::pack2(f[4], f[5]);
              o[half * 4 + c].w = D::pack2(f[6], f[7]);
            }
          }
          if (use > 0) {                         // the group's previous store must have read the slot (its read overlapped the math above)
            if (elected) {
              ptx::bulk_wait_read<0>();
              ptx::mbar_arrive(&slot_empty[grp]);
            }
            ptx::mbar_wait(&slot_empty[grp], (use - 1) & 1);
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(srow + (((uint32_t)c ^ sw) << 4)) = o[c];
          ptx::fence_proxy_async_smem();
          named_bar_sync(1 + grp, 128);
          if (elected) {
            ptx::tma_store_2d(&tm_y, slot_base + grp * kBoxBytes, col0, row0);
            ptx::bulk_commit();
          }
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    if (elected) ptx::bulk_wait<0>();
  }

  __syncwarp();
  ptx::tc_fence_before();
  ptx::cluster_sync_all();     // the leader's MMAs read the peer's shared memory: nobody leaves early
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ===================================================================================== x = resid + A W^T + b (+ LayerNorms), N = 512
// combine (mean, M2) of the two 256-column halves of a row
__device__ __forceinline__ void row_stats_exchange(float2* sx, int grp, int row, float mean_h, float m2_h, float eps, float& mean, float& rstd) {
  sx[grp * kBM + row] = make_float2(mean_h, m2_h);
  named_bar_sync(3, 256);
  const float2 o = sx[(grp ^ 1) * kBM + row];
  named_bar_sync(3, 256);      // both halves have read before the buffer is reused
  const float delta = o.x - mean_h;
  mean = 0.5f * (mean_h + o.x);
  const float m2 = m2_h + o.y + delta * delta * 128.0f;   // n_a n_b / (n_a + n_b) = 256 * 256 / 512
  rstd = rsqrtf(m2 * (1.0f / 512.0f) + eps);
}

template <typename D, int kLnStages, int kLnSlots, bool kSplit>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
pair_linear_ln_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_r,
                      const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_p,
                      LnArgs args, int M, int K) {
  // The K loop is short (8 or 16 k-blocks) and the epilogue moves 5 bytes per output element, so shared memory goes to the
  // box ring (~128 KB of residual loads / output stores in flight per SM), not to operand stages.
  //
  // kSplit = false: one 512-column accumulation per tile (every k-block feeds both N = 256 halves, 48 KB stages); the tensor pipe
  // idles during the epilogue and the epilogue warps during the main loop.
  // kSplit = true: the two N = 256 halves are accumulated ONE AFTER THE OTHER (the K loop runs twice per tile, A is read twice -
  // the second time out of L2; 32 KB stages), each into its own half of TMEM with its own full / empty barrier.  All eight
  // epilogue warps work on one half at a time (warp group g owns columns [128 g, 128 g + 128) of the half), so the first pass over
  // half 0 overlaps the MMAs of half 1, and half 0 is handed back to the MMA warp - which then starts the NEXT tile - while the
  // last pass over half 1 is still running.  It wins when the main loop is short (K = 512: 663 -> 580 us at 128 clips); at
  // K = 1024 the second read of A makes the main loop L2-bound (933 -> 1000 us), so fc2 keeps the single accumulation.
  constexpr int kStages = kLnStages;
  constexpr int kStageBytes = kABytes + (kSplit ? 1 : 2) * kWHalfBytes;
  constexpr int kHalves = kSplit ? 2 : 1;                  // accumulation passes per tile
  constexpr int kN = 512;
  constexpr int kGS = kLnSlots / 2;                        // slots per warp group

  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();             // 128-byte-swizzle tiles need the 1024-byte alignment the declaration asks for
  uint8_t* stage_base = smem;
  uint8_t* slot_base = smem + kStages * kStageBytes;
  const PairBars bars = carve_bars<kStages, kLnSlots>(slot_base + kLnSlots * kBoxBytes);
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars.tile_done + 1);
  float2* sx = reinterpret_cast<float2*>(tmem_holder + 4);           // [2][128] row statistics exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int num_tiles = (M + 2 * kBM - 1) / (2 * kBM);
  const int k_blocks = K / kBK;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const bool has_post = args.post_g != nullptr, has_ln = args.ln_g != nullptr;
  const bool store_a = !has_post || args.has_xpre != 0;   // pass A hands its value to a TMA store (x_out, or x_pre ahead of the post-norm)

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_w);
    ptx::prefetch_tmap(&tm_r);
    ptx::prefetch_tmap(&tm_x);
    if (has_ln) ptx::prefetch_tmap(&tm_h);
    if (has_post && store_a) ptx::prefetch_tmap(&tm_p);
  }
  if (warp == 1 && lane == 0) init_bars<kStages, kLnSlots>(bars);
  if (warp == 2) {
    ptx::tmem_alloc(tmem_holder, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();   // the preceding kernel has completed: operands may be read, outputs written

  if (warp == 0) {
    if (lane == 0) {
      // ===================== operand producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
        const int row_a = tile * 2 * kBM + (int)rank * kBM;
        for (int half = 0; half < kHalves; ++half) {
          for (int kb = 0; kb < k_blocks; ++kb) {
            ptx::mbar_wait(&bars.empty[stage], phase ^ 1);
            uint8_t* sa = stage_base + stage * kStageBytes;
            const uint32_t full_leader = ptx::mapa_shared(smem_u32(&bars.full[stage]), 0);
            if (leader) ptx::mbar_expect_tx(&bars.full[stage], 2 * kStageBytes);   // the peer's loads complete_tx on this barrier too
            ptx::tma_load_2d_pair(sa, &tm_a, full_leader, kb * kBK, row_a);
            ptx::tma_load_2d_pair(sa + kABytes, &tm_w, full_leader, kb * kBK, half * 256 + (int)rank * 128);   // W rows of this half's columns
            if constexpr (!kSplit) ptx::tma_load_2d_pair(sa + kABytes + kWHalfBytes, &tm_w, full_leader, kb * kBK, 256 + (int)rank * 128);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = ptx::umma_idesc_16(2 * kBM, 256, D::kUmmaFmt);
      int stage = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
        for (int half = 0; half < kHalves; ++half) {
          ptx::mbar_wait(&bars.tmem_empty[half], acc_phase ^ 1);
          ptx::tc_fence_after();
          const uint32_t t_acc = tmem_base + (uint32_t)(half * 256);
          for (int kb = 0; kb < k_blocks; ++kb) {
            ptx::mbar_wait(&bars.full[stage], phase);
            ptx::tc_fence_after();
            const uint32_t sa = smem_u32(stage_base + stage * kStageBytes);
            const uint64_t da = ptx::umma_desc_sw128(sa);
            const uint64_t db = ptx::umma_desc_sw128(sa + kABytes);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
              ptx::umma_f16_pair(t_acc, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, accum);
              if constexpr (!kSplit)
                ptx::umma_f16_pair(t_acc + 256u, da + (uint64_t)(2 * k), ptx::umma_desc_sw128(sa + kABytes + kWHalfBytes) + (uint64_t)(2 * k), idesc, accum);
            }
            ptx::umma_commit_pair(&bars.empty[stage]);
            if (kb == k_blocks - 1) ptx::umma_commit_pair(&bars.tmem_full[half]);
            if (++stage == kStages) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
        acc_phase ^= 1;
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      // ===================== residual loader =====================
      // Per tile and warp group g the slot-use sequence is: 8 residual boxes (4 of accumulator half 0, then 4 of half 1), then
      // (post-norm) 8 x boxes, then (pre-norm) 4 h boxes; box j of a 32-column pass covers columns (j / 4) * 256 + g * 128 + (j % 4) * 32.
      // Uses are numbered with a running counter n per group: use n lives in slot g * kGS + n % kGS and completes phase
      // n / kGS of slot_empty[slot] when that slot is free again.  A parity wait only distinguishes "the previous phase" from
      // "the one before", and the loader skips the store-only uses, so across tiles it waits for `tile_done` (both halves arrive
      // after their last store has been read: every slot is free) and within a tile only for uses kGS.. (lockstep again).
      const uint32_t uses = 8u + ((has_post && args.no_x == 0) ? 8u : 0u) + (has_ln ? 4u : 0u);
      auto box_col32 = [](int g, int j) { return kSplit ? (j >> 2) * 256 + g * 128 + (j & 3) * 32 : g * 256 + j * 32; };
      uint32_t tt = 0;
      uint32_t full_par = 0;                     // parity bit of slot_full per slot (flips with every load into the slot)
      (void)full_par;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++tt) {
        const int row0 = tile * 2 * kBM + (int)rank * kBM;
        if (tt > 0) ptx::mbar_wait(bars.tile_done, (tt - 1) & 1);
        for (int j = 0; j < 8; ++j) {
          const uint32_t n = tt * uses + (uint32_t)j;
          for (int g = 0; g < 2; ++g) {
            const uint32_t slot = (uint32_t)(g * kGS) + n % kGS;
            if (j >= kGS) ptx::mbar_wait(&bars.slot_empty[slot], ((n / kGS) & 1) ^ 1);
            ptx::mbar_expect_tx(&bars.slot_full[slot], kBoxBytes);
            ptx::tma_load_2d(slot_base + slot * kBoxBytes, &tm_r, &bars.slot_full[slot], box_col32(g, j), row0);
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;             // warp group: columns [128 grp, 128 grp + 128) of each accumulator half
    const bool elected = ((warp - 4) & 3) == 0 && lane == 0;
    const int row = 32 * q + lane;
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t t_row = tmem_base + ((uint32_t)(32 * q) << 16);   // + the output column
    uint32_t acc_phase = 0;
    int pending = -1;                            // slot whose TMA store may still be reading shared memory
    uint32_t n = 0;                              // running slot-use counter of this column half (see the loader)
    uint32_t full_par = 0;                       // parity of the next slot_full phase, one bit per slot of this half

    // after a slot's content has been handed to a TMA store: recycle the PREVIOUS store's slot (its read is done by now)
    auto after_store = [&](uint32_t slot) {
      ptx::bulk_commit();
      if (pending >= 0) {
        ptx::bulk_wait_read<1>();
        ptx::mbar_arrive(&bars.slot_empty[pending]);
      }
      pending = (int)slot;
    };

    // first output column of box j of a 32-column pass / of box jj of the 64-column pass of warp group grp
    auto box_col32 = [&](int j) { return kSplit ? (j >> 2) * 256 + grp * 128 + (j & 3) * 32 : grp * 256 + j * 32; };
    auto box_col64 = [&](int jj) { return kSplit ? (jj >> 1) * 256 + grp * 128 + (jj & 1) * 64 : grp * 256 + jj * 64; };
    // hand accumulator half `half` back to the MMA warp (16 arrivals: every epilogue warp of both CTAs)
    auto release_half = [&](int half) {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_shared(smem_u32(&bars.tmem_empty[half]), 0));
    };
    const int last_pass = has_ln ? 2 : (has_post ? 1 : 0);   // the last pass that reads TMEM releases each half as it finishes it

    for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
      const int row0 = tile * 2 * kBM + (int)rank * kBM;
      const int grow = row0 + row;

      // ---- pass A: v = resid + scale * (acc + bias); shifted sums for the row statistics; v goes back to TMEM (and out, if no post-norm)
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
      uint64_t s1p = 0, s2p = 0, nshift2 = 0, nmean2 = 0, rstd2 = 0;   // packed accumulators / broadcast operands (common.cuh)
      const float scale = (args.row_scale != nullptr && grow < M) ? __ldg(args.row_scale + grow) : 1.0f;
      const uint64_t scale2 = pack_f32x2(scale, scale);
#pragma unroll 1
      for (int j = 0; j < 8; ++j, ++n) {
        const uint32_t ls = n % kGS, slot = (uint32_t)(grp * kGS) + ls;
        const int col = box_col32(j);
        uint8_t* srow = slot_base + slot * kBoxBytes + row * 128;
        if (j == 0 || (kSplit && j == 4)) {      // accumulator (half j / 4) complete
          ptx::mbar_wait(&bars.tmem_full[j >> 2], acc_phase);
          ptx::tc_fence_after();
        }
        ptx::mbar_wait(&bars.slot_full[slot], (full_par >> ls) & 1);
        full_par ^= 1u << ls;
        uint32_t r[32];
        ptx::tmem_ld32(t_row + (uint32_t)col, r);
        ptx::tmem_ld_wait();
        const float4* b4 = reinterpret_cast<const float4*>(args.bias + col);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4* p = reinterpret_cast<float4*>(srow + (((uint32_t)c ^ sw) << 4));
          const float4 bb = __ldg(b4 + c);
          float4 v = *p;
          // packed fp32 (FFMA2 / FADD2: two elements per issue slot); scale == 1: the same two roundings as v += acc + bias
          const uint64_t v01 = fma_f32x2(scale2, add_f32x2(pack_u32x2(r[4 * c + 0], r[4 * c + 1]), pack_f32x2(bb.x, bb.y)), pack_f32x2(v.x, v.y));
          const uint64_t v23 = fma_f32x2(scale2, add_f32x2(pack_u32x2(r[4 * c + 2], r[4 * c + 3]), pack_f32x2(bb.z, bb.w)), pack_f32x2(v.z, v.w));
          unpack_f32x2(v01, v.x, v.y);
          unpack_f32x2(v23, v.z, v.w);
          if (j == 0 && c == 0) {
            shift = v.x;
            nshift2 = pack_f32x2(-shift, -shift);
          }
          const uint64_t d01 = add_f32x2(v01, nshift2), d23 = add_f32x2(v23, nshift2);
          s1p = add_f32x2(s1p, add_f32x2(d01, d23));
          s2p = fma_f32x2(d23, d23, fma_f32x2(d01, d01, s2p));
          r[4 * c + 0] = __float_as_uint(v.x);
          r[4 * c + 1] = __float_as_uint(v.y);
          r[4 * c + 2] = __float_as_uint(v.z);
          r[4 * c + 3] = __float_as_uint(v.w);
          if (store_a) *p = v;
        }
        if (has_post || has_ln) ptx::tmem_st32(t_row + (uint32_t)col, r);
        if (store_a) ptx::fence_proxy_async_smem();
        named_bar_sync(1 + grp, 128);            // every thread of the group is done with the residual box
        if (elected) {
          if (store_a) {
            ptx::tma_store_2d(has_post ? &tm_p : &tm_x, slot_base + slot * kBoxBytes, col, row0);
            after_store(slot);
          } else {
            ptx::mbar_arrive(&bars.slot_empty[slot]);
          }
        }
        // no later pass reads TMEM.  (Handing half 0 back already after box 3 of THIS pass - the same call as in passes B / C -
        // raised 'misaligned address' on the device and was not chased further: without LayerNorms the model only uses the
        // K = 1024 single-accumulation variant, where there is one hand-back per tile anyway.)
        if (last_pass == 0 && j == 7) {
          release_half(0);
          if constexpr (kSplit) release_half(1);
        }
      }
      float mean = 0.f, rstd = 0.f;
      if (has_post || has_ln) {
        ptx::tmem_st_wait();
        s1 = sum_f32x2(s1p);
        s2 = sum_f32x2(s2p);
        const float mh = shift + s1 * (1.0f / 256.0f);
        const float m2h = fmaxf(s2 - s1 * s1 * (1.0f / 256.0f), 0.f);
        row_stats_exchange(sx, grp, row, mh, m2h, has_post ? args.post_eps : args.ln_eps, mean, rstd);
        nmean2 = pack_f32x2(-mean, -mean);
        rstd2 = pack_f32x2(rstd, rstd);
      }

      // ---- pass B (post-norm): y = LN_post(v) (+ pos-embed) -> x_out and back to TMEM, statistics of y
      if (has_post) {
        const float* pos_row = args.pos ? args.pos + (size_t)((grow / args.pos_div) % args.pos_mod) * kN : nullptr;
        s1p = 0;
        s2p = 0;
        const bool store_b = args.no_x == 0;        // x_out wanted (false: only the statistics and the TMEM copy for pass C)
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
          const uint32_t slot = (uint32_t)(grp * kGS) + n % kGS;
          uint8_t* srow = slot_base + slot * kBoxBytes + row * 128;
          if (store_b) ptx::mbar_wait(&bars.slot_empty[slot], ((n / kGS) & 1) ^ 1);
          uint32_t r[32];
          const int col = box_col32(j);
          ptx::tmem_ld32(t_row + (uint32_t)col, r);
          ptx::tmem_ld_wait();
          const float4* g4 = reinterpret_cast<const float4*>(args.post_g + col);
          const float4* be4 = reinterpret_cast<const float4*>(args.post_b + col);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 gg = __ldg(g4 + c), bb = __ldg(be4 + c);
            uint64_t y01 = fma_f32x2(mul_f32x2(add_f32x2(pack_u32x2(r[4 * c + 0], r[4 * c + 1]), nmean2), rstd2), pack_f32x2(gg.x, gg.y), pack_f32x2(bb.x, bb.y));
            uint64_t y23 = fma_f32x2(mul_f32x2(add_f32x2(pack_u32x2(r[4 * c + 2], r[4 * c + 3]), nmean2), rstd2), pack_f32x2(gg.z, gg.w), pack_f32x2(bb.z, bb.w));
            if (pos_row != nullptr && grow < M) {
              const float4 pe = __ldg(reinterpret_cast<const float4*>(pos_row + col) + c);
              y01 = add_f32x2(y01, pack_f32x2(pe.x, pe.y));
              y23 = add_f32x2(y23, pack_f32x2(pe.z, pe.w));
            }
            float4 y;
            unpack_f32x2(y01, y.x, y.y);
            unpack_f32x2(y23, y.z, y.w);
            if (j == 0 && c == 0) {
              shift = y.x;
              nshift2 = pack_f32x2(-shift, -shift);
            }
            const uint64_t d01 = add_f32x2(y01, nshift2), d23 = add_f32x2(y23, nshift2);
            s1p = add_f32x2(s1p, add_f32x2(d01, d23));
            s2p = fma_f32x2(d23, d23, fma_f32x2(d01, d01, s2p));
            r[4 * c + 0] = __float_as_uint(y.x);
            r[4 * c + 1] = __float_as_uint(y.y);
            r[4 * c + 2] = __float_as_uint(y.z);
            r[4 * c + 3] = __float_as_uint(y.w);
            if (store_b) *reinterpret_cast<float4*>(srow + (((uint32_t)c ^ sw) << 4)) = y;
          }
          if (has_ln) ptx::tmem_st32(t_row + (uint32_t)col, r);
          if (store_b) {
            ptx::fence_proxy_async_smem();
            named_bar_sync(1 + grp, 128);
            if (elected) {
              ptx::tma_store_2d(&tm_x, slot_base + slot * kBoxBytes, col, row0);
              after_store(slot);
            }
            ++n;
          }
          if (last_pass == 1 && (kSplit ? (j & 3) == 3 : j == 7)) release_half(kSplit ? j >> 2 : 0);
        }
        if (has_ln) {
          ptx::tmem_st_wait();
          s1 = sum_f32x2(s1p);
          s2 = sum_f32x2(s2p);
          const float mh = shift + s1 * (1.0f / 256.0f);
          const float m2h = fmaxf(s2 - s1 * s1 * (1.0f / 256.0f), 0.f);
          row_stats_exchange(sx, grp, row, mh, m2h, args.ln_eps, mean, rstd);
          nmean2 = pack_f32x2(-mean, -mean);
          rstd2 = pack_f32x2(rstd, rstd);
        }
      }

      // ---- pass C (pre-norm): h = LN_pre(x) as 16-bit, 64-column boxes
      if (has_ln) {
#pragma unroll 1
        for (int jj = 0; jj < 4; ++jj, ++n) {
          const uint32_t slot = (uint32_t)(grp * kGS) + n % kGS;
          uint8_t* srow = slot_base + slot * kBoxBytes + row * 128;
          ptx::mbar_wait(&bars.slot_empty[slot], ((n / kGS) & 1) ^ 1);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t r[32];
            const int col = box_col64(jj) + half * 32;
            ptx::tmem_ld32(t_row + (uint32_t)col, r);
            ptx::tmem_ld_wait();
            const float4* g4 = reinterpret_cast<const float4*>(args.ln_g + col);
            const float4* be4 = reinterpret_cast<const float4*>(args.ln_b + col);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 g0 = __ldg(g4 + 2 * c), g1 = __ldg(g4 + 2 * c + 1), b0 = __ldg(be4 + 2 * c), b1 = __ldg(be4 + 2 * c + 1);
              uint4 o;
              o.x = D::pack2(fmaf((__uint_as_float(r[8 * c + 0]) - mean) * rstd, g0.x, b0.x), fmaf((__uint_as_float(r[8 * c + 1]) - mean) * rstd, g0.y, b0.y));
              o.y = D::pack2(fmaf((__uint_as_float(r[8 * c + 2]) - mean) * rstd, g0.z, b0.z), fmaf((__uint_as_float(r[8 * c + 3]) - mean) * rstd, g0.w, b0.w));
              o.z = D::pack2(fmaf((__uint_as_float(r[8 * c + 4]) - mean) * rstd, g1.x, b1.x), fmaf((__uint_as_float(r[8 * c + 5]) - mean) * rstd, g1.y, b1.y));
              o.w = D::pack2(fmaf((__uint_as_float(r[8 * c + 6]) - mean) * rstd, g1.z, b1.z), fmaf((__uint_as_float(r[8 * c + 7]) - mean) * rstd, g1.w, b1.w));
              *reinterpret_cast<uint4*>(srow + (((uint32_t)(half * 4 + c) ^ sw) << 4)) = o;
            }
          }
          ptx::fence_proxy_async_smem();
          named_bar_sync(1 + grp, 128);
          if (elected) {
            ptx::tma_store_2d(&tm_h, slot_base + slot * kBoxBytes, box_col64(jj), row0);
            after_store(slot);
          }
          if (kSplit ? (jj & 1) : jj == 3) release_half(kSplit ? jj >> 1 : 0);
        }
      }
      // ---- end of tile: the last store's slot must be free before the loader refills it, TMEM is drained
      if (elected) {
        if (pending >= 0) {
          ptx::bulk_wait_read<0>();
          ptx::mbar_arrive(&bars.slot_empty[pending]);
          pending = -1;
        }
        ptx::mbar_arrive(bars.tile_done);
      }
      acc_phase ^= 1;
    }
    if (elected) ptx::bulk_wait<0>();
  }

  __syncwarp();
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ===================================================================================== the same, 64 rows per CTA, two accumulators
// pair_linear_ln_kernel keeps ONE 128 x 512 fp32 accumulator per CTA (all of TMEM), so a tile's epilogue (three passes, HBM bound)
// and the next tile's main loop cannot overlap: with K = 1024 (fc2) the tensor pipe idles ~2/3 of the time and HBM idles during
// the main loop (0.62 of the copy bandwidth).  Here the pair multiplies 128 rows per tile (tcgen05.mma.cta_group::2 with M = 128:
// 64 rows per CTA, whose [64 x 256] fp32 result is folded over the 128 TMEM lanes - lanes 0-63 hold columns 0-127 of the
// instruction's N = 256, lanes 64-127 columns 128-255 - so a 64 x 512 tile takes 256 TMEM columns), which leaves room for TWO
// accumulators: the MMA warp runs one tile ahead of the epilogue.  W is pulled through L2 once per 128 rows instead of once per
// 256 (8 instead of 4 KB per token), which these launches can afford (L2 throughput is at ~40 % in the 256-row kernels).
//
// Epilogue mapping: warp w has TMEM lanes 32 (w % 4)...; q = w % 4, grp = (w - 4) / 4, h = q / 2.  A thread owns row 32 (q & 1) + lane
// of the CTA's 64 and the 128 CONTIGUOUS output columns [256 grp + 128 h, + 128) = four 32-column fp32 boxes (TMEM columns
// acc * 256 + 128 grp + 32 j): four "column groups" cg = 2 grp + h of two warps each, row statistics combined four ways through
// shared memory.  Boxes are 64 rows x 128 bytes (8 KB).  Residual boxes arrive through a load ring (kLoad slots per column group,
// filled by the loader warp, which runs ahead into the next tile), results leave through one store slot per column group.
constexpr int kRows64 = 64;
constexpr int kBox64 = kRows64 * 128;                       // 8 KB
constexpr int kA64 = kRows64 * kBK * 2;                      // 8 KB
constexpr int kStage64 = kA64 + 2 * kWHalfBytes;             // 40 KB: A | W rows of output columns 0-255 | of columns 256-511

template <typename D, int kStages, int kLoad>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
pair_linear_ln64_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_w, const __grid_constant__ CUtensorMap tm_r,
                        const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_p,
                        LnArgs args, int M, int K) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* stage_base = smem;
  uint8_t* load_base = smem + kStages * kStage64;              // [4][kLoad] residual boxes
  uint8_t* store_base = load_base + 4 * kLoad * kBox64;        // [4] output boxes
  uint64_t* full = reinterpret_cast<uint64_t*>(store_base + 4 * kBox64);   // [stages] leader: operands of both CTAs landed
  uint64_t* empty = full + kStages;                            // [stages] each CTA: stage consumed (multicast commit)
  uint64_t* tmem_full = empty + kStages;                       // [2] each CTA: accumulator complete (multicast commit)
  uint64_t* tmem_empty = tmem_full + 2;                        // [2] leader: drained by the 16 epilogue warps of the pair
  uint64_t* load_full = tmem_empty + 2;                        // [4 * kLoad] residual box landed
  uint64_t* load_empty = load_full + 4 * kLoad;                // [4 * kLoad] box consumed (and, if stored from, read by the store)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(load_empty + 4 * kLoad);
  float2* sx = reinterpret_cast<float2*>(tmem_holder + 4);     // [4][64] row statistics exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int num_tiles = (M + 2 * kRows64 - 1) / (2 * kRows64);
  const int k_blocks = K / kBK;
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const bool has_post = args.post_g != nullptr, has_ln = args.ln_g != nullptr;
  const bool store_a = !has_post || args.has_xpre != 0;        // pass A hands its value to a TMA store (x_out, or x_pre ahead of the post-norm)
  const bool store_b = has_post && args.no_x == 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_a);
    ptx::prefetch_tmap(&tm_w);
    ptx::prefetch_tmap(&tm_r);
    ptx::prefetch_tmap(&tm_x);
    if (has_ln) ptx::prefetch_tmap(&tm_h);
    if (has_post && store_a) ptx::prefetch_tmap(&tm_p);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full[a], 1);
      ptx::mbar_init(&tmem_empty[a], 2 * kEpiWarps);
    }
    for (int s = 0; s < 4 * kLoad; ++s) {
      ptx::mbar_init(&load_full[s], 1);
      ptx::mbar_init(&load_empty[s], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_holder, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();   // the preceding kernel has completed: operands may be read, outputs written

  if (warp == 0) {
    if (lane == 0) {
      // ===================== operand producer =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
        const int row_a = tile * 2 * kRows64 + (int)rank * kRows64;
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = stage_base + stage * kStage64;
          const uint32_t full_leader = ptx::mapa_shared(smem_u32(&full[stage]), 0);
          if (leader) ptx::mbar_expect_tx(&full[stage], 2 * kStage64);   // the peer's loads complete_tx on this barrier too
          ptx::tma_load_2d_pair(sa, &tm_a, full_leader, kb * kBK, row_a);
          ptx::tma_load_2d_pair(sa + kA64, &tm_w, full_leader, kb * kBK, (int)rank * 128);
          ptx::tma_load_2d_pair(sa + kA64 + kWHalfBytes, &tm_w, full_leader, kb * kBK, 256 + (int)rank * 128);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      // ===================== MMA issuer: M = 128 per pair, two N = 256 instructions per k-step =====================
      constexpr uint32_t idesc = ptx::umma_idesc_16(2 * kRows64, 256, D::kUmmaFmt);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
        ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        ptx::tc_fence_after();
        const uint32_t t_acc = tmem_base + (uint32_t)(acc * 256);
        for (int kb = 0; kb < k_blocks; ++kb) {
          ptx::mbar_wait(&full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = smem_u32(stage_base + stage * kStage64);
          const uint64_t da = ptx::umma_desc_sw128(sa);
          const uint64_t db0 = ptx::umma_desc_sw128(sa + kA64);
          const uint64_t db1 = ptx::umma_desc_sw128(sa + kA64 + kWHalfBytes);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k) {
            const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
            ptx::umma_f16_pair(t_acc, da + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), idesc, accum);
            ptx::umma_f16_pair(t_acc + 128u, da + (uint64_t)(2 * k), db1 + (uint64_t)(2 * k), idesc, accum);
          }
          ptx::umma_commit_pair(&empty[stage]);
          if (kb == k_blocks - 1) ptx::umma_commit_pair(&tmem_full[acc]);
          if (++stage == kStages) {
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
    if (lane == 0) {
      // ===================== residual loader: box j of column group cg of tile tt is load use nl = 4 tt + j of that group =====================
      uint32_t tt = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++tt) {
        const int row0 = tile * 2 * kRows64 + (int)rank * kRows64;
        for (int j = 0; j < 4; ++j) {
          const uint32_t nl = tt * 4u + (uint32_t)j;
          for (int cg = 0; cg < 4; ++cg) {
            const uint32_t ls = (uint32_t)(cg * kLoad) + nl % kLoad;
            if (nl >= (uint32_t)kLoad) ptx::mbar_wait(&load_empty[ls], ((nl / kLoad) & 1) ^ 1);
            ptx::mbar_expect_tx(&load_full[ls], kBox64);
            ptx::tma_load_2d(load_base + ls * kBox64, &tm_r, &load_full[ls], (cg >> 1) * 256 + (cg & 1) * 128 + j * 32, row0);
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int h = q >> 1;
    const int cg = 2 * grp + h;                          // column group: output columns [cbase, cbase + 128)
    const int cbase = 256 * grp + 128 * h;
    const int row = 32 * (q & 1) + lane;                 // row of the CTA's 64
    const bool elected = (q & 1) == 0 && lane == 0;
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t t_lane = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(128 * grp);   // + acc * 256 + 32 j
    uint8_t* sslot = store_base + cg * kBox64;
    uint8_t* srow_st = sslot + row * 128;
    int acc = 0;
    uint32_t acc_phase = 0, tt = 0;
    bool pending = false;                                // elected thread: a store from the group's store slot may still be reading it

    // four-way combination of (mean, M2) over the 128-column quarters of a row
    auto row_stats = [&](float mean_q, float m2_q, float eps, float& mean, float& rstd) {
      sx[cg * kRows64 + row] = make_float2(mean_q, m2_q);
      named_bar_sync(5, 256);
      float2 p[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = sx[i * kRows64 + row];
      named_bar_sync(5, 256);                            // every quarter has read before the buffer is reused
      mean = 0.25f * ((p[0].x + p[1].x) + (p[2].x + p[3].x));
      float m2 = (p[0].y + p[1].y) + (p[2].y + p[3].y);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float d = p[i].x - mean;
        m2 = fmaf(128.0f * d, d, m2);
      }
      rstd = rsqrtf(m2 * (1.0f / 512.0f) + eps);
    };
    // store `box` (already written to the group's store slot by all 64 threads) - called by everybody, acts in the elected thread
    auto store_from_slot = [&](const CUtensorMap* map, int col, int row0) {
      ptx::fence_proxy_async_smem();
      named_bar_sync(1 + cg, 64);
      if (elected) {
        ptx::tma_store_2d(map, sslot, col, row0);
        ptx::bulk_commit();
        pending = true;
      }
    };
    auto slot_writable = [&]() {                         // before anybody writes the store slot again
      if (elected && pending) {
        ptx::bulk_wait_read<0>();
        pending = false;
      }
      named_bar_sync(1 + cg, 64);
    };
    const int last_pass = has_ln ? 2 : (has_post ? 1 : 0);
    auto release_acc = [&]() {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_shared(smem_u32(&tmem_empty[acc]), 0));
    };

    for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++tt) {
      const int row0 = tile * 2 * kRows64 + (int)rank * kRows64;
      const int grow = row0 + row;
      const uint32_t t_row = t_lane + (uint32_t)(acc * 256);
      ptx::mbar_wait(&tmem_full[acc], acc_phase);
      ptx::tc_fence_after();

      // ---- pass A: v = resid + scale * (acc + bias); shifted sums for the row statistics; v goes back to TMEM (and out, if no post-norm)
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
      uint64_t s1p = 0, s2p = 0, nshift2 = 0, nmean2 = 0, rstd2 = 0;   // packed accumulators / broadcast operands (common.cuh)
      const float scale = (args.row_scale != nullptr && grow < M) ? __ldg(args.row_scale + grow) : 1.0f;
      const uint64_t scale2 = pack_f32x2(scale, scale);
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        const uint32_t nl = tt * 4u + (uint32_t)j;
        const uint32_t ls = (uint32_t)(cg * kLoad) + nl % kLoad;
        const int col = cbase + 32 * j;
        uint8_t* lrow = load_base + ls * kBox64 + row * 128;
        uint32_t r[32];
        ptx::tmem_ld32(t_row + (uint32_t)(32 * j), r);
        ptx::mbar_wait(&load_full[ls], (nl / kLoad) & 1);
        ptx::tmem_ld_wait();
        const float4* b4 = reinterpret_cast<const float4*>(args.bias + col);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4* p = reinterpret_cast<float4*>(lrow + (((uint32_t)c ^ sw) << 4));
          const float4 bb = __ldg(b4 + c);
          float4 v = *p;
          // packed fp32 (FFMA2 / FADD2: two elements per issue slot); scale == 1: the same two roundings as v += acc + bias
          const uint64_t v01 = fma_f32x2(scale2, add_f32x2(pack_u32x2(r[4 * c + 0], r[4 * c + 1]), pack_f32x2(bb.x, bb.y)), pack_f32x2(v.x, v.y));
          const uint64_t v23 = fma_f32x2(scale2, add_f32x2(pack_u32x2(r[4 * c + 2], r[4 * c + 3]), pack_f32x2(bb.z, bb.w)), pack_f32x2(v.z, v.w));
          unpack_f32x2(v01, v.x, v.y);
          unpack_f32x2(v23, v.z, v.w);
          if (j == 0 && c == 0) {
            shift = v.x;
            nshift2 = pack_f32x2(-shift, -shift);
          }
          const uint64_t d01 = add_f32x2(v01, nshift2), d23 = add_f32x2(v23, nshift2);
          s1p = add_f32x2(s1p, add_f32x2(d01, d23));
          s2p = fma_f32x2(d23, d23, fma_f32x2(d01, d01, s2p));
          r[4 * c + 0] = __float_as_uint(v.x);
          r[4 * c + 1] = __float_as_uint(v.y);
          r[4 * c + 2] = __float_as_uint(v.z);
          r[4 * c + 3] = __float_as_uint(v.w);
          if (store_a) *p = v;
        }
        if (has_post || has_ln) ptx::tmem_st32(t_row + (uint32_t)(32 * j), r);
        if (store_a) ptx::fence_proxy_async_smem();
        named_bar_sync(1 + cg, 64);                        // both warps of the column group are done with the residual box
        if (elected) {
          if (store_a) {                                   // x (or x_pre) leaves straight from the residual's slot
            ptx::tma_store_2d(has_post ? &tm_p : &tm_x, load_base + ls * kBox64, col, row0);
            ptx::bulk_commit();
            ptx::bulk_wait_read<0>();
          }
          ptx::mbar_arrive(&load_empty[ls]);
        }
      }
      if (last_pass == 0) release_acc();
      float mean = 0.f, rstd = 0.f;
      if (has_post || has_ln) {
        ptx::tmem_st_wait();
        s1 = sum_f32x2(s1p);
        s2 = sum_f32x2(s2p);
        const float mq = shift + s1 * (1.0f / 128.0f);
        const float m2q = fmaxf(s2 - s1 * s1 * (1.0f / 128.0f), 0.f);
        row_stats(mq, m2q, has_post ? args.post_eps : args.ln_eps, mean, rstd);
        nmean2 = pack_f32x2(-mean, -mean);
        rstd2 = pack_f32x2(rstd, rstd);
      }

      // ---- pass B (post-norm): y = LN_post(v) (+ pos-embed) -> x_out and back to TMEM, statistics of y
      if (has_post) {
        const float* pos_row = args.pos ? args.pos + (size_t)((grow / args.pos_div) % args.pos_mod) * 512 : nullptr;
        s1p = 0;
        s2p = 0;
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
          const int col = cbase + 32 * j;
          uint32_t r[32];
          ptx::tmem_ld32(t_row + (uint32_t)(32 * j), r);
          ptx::tmem_ld_wait();
          const float4* g4 = reinterpret_cast<const float4*>(args.post_g + col);
          const float4* be4 = reinterpret_cast<const float4*>(args.post_b + col);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 gg = __ldg(g4 + c), bb = __ldg(be4 + c);
            uint64_t y01 = fma_f32x2(mul_f32x2(add_f32x2(pack_u32x2(r[4 * c + 0], r[4 * c + 1]), nmean2), rstd2), pack_f32x2(gg.x, gg.y), pack_f32x2(bb.x, bb.y));
            uint64_t y23 = fma_f32x2(mul_f32x2(add_f32x2(pack_u32x2(r[4 * c + 2], r[4 * c + 3]), nmean2), rstd2), pack_f32x2(gg.z, gg.w), pack_f32x2(bb.z, bb.w));
            if (pos_row != nullptr && grow < M) {
              const float4 pe = __ldg(reinterpret_cast<const float4*>(pos_row + col) + c);
              y01 = add_f32x2(y01, pack_f32x2(pe.x, pe.y));
              y23 = add_f32x2(y23, pack_f32x2(pe.z, pe.w));
            }
            float4 y;
            unpack_f32x2(y01, y.x, y.y);
            unpack_f32x2(y23, y.z, y.w);
            if (j == 0 && c == 0) {
              shift = y.x;
              nshift2 = pack_f32x2(-shift, -shift);
            }
            const uint64_t d01 = add_f32x2(y01, nshift2), d23 = add_f32x2(y23, nshift2);
            s1p = add_f32x2(s1p, add_f32x2(d01, d23));
            s2p = fma_f32x2(d23, d23, fma_f32x2(d01, d01, s2p));
            r[4 * c + 0] = __float_as_uint(y.x);
            r[4 * c + 1] = __float_as_uint(y.y);
            r[4 * c + 2] = __float_as_uint(y.z);
            r[4 * c + 3] = __float_as_uint(y.w);
          }
          if (has_ln) ptx::tmem_st32(t_row + (uint32_t)(32 * j), r);
          if (store_b) {
            slot_writable();
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<uint4*>(srow_st + (((uint32_t)c ^ sw) << 4)) = make_uint4(r[4 * c + 0], r[4 * c + 1], r[4 * c + 2], r[4 * c + 3]);
            store_from_slot(&tm_x, col, row0);
          }
        }
        if (last_pass == 1) release_acc();
        if (has_ln) {
          ptx::tmem_st_wait();
          s1 = sum_f32x2(s1p);
          s2 = sum_f32x2(s2p);
          const float mq = shift + s1 * (1.0f / 128.0f);
          const float m2q = fmaxf(s2 - s1 * s1 * (1.0f / 128.0f), 0.f);
          row_stats(mq, m2q, args.ln_eps, mean, rstd);
          nmean2 = pack_f32x2(-mean, -mean);
          rstd2 = pack_f32x2(rstd, rstd);
        }
      }

      // ---- pass C (pre-norm): h = LN_pre(x) as 16-bit, two 64-column boxes
      if (has_ln) {
#pragma unroll 1
        for (int jj = 0; jj < 2; ++jj) {
          uint32_t r0[32], r1[32];
          ptx::tmem_ld32(t_row + (uint32_t)(64 * jj), r0);
          ptx::tmem_ld32(t_row + (uint32_t)(64 * jj + 32), r1);
          ptx::tmem_ld_wait();
          if (jj == 1) release_acc();                      // last TMEM read of the tile: the MMA warp may start tile + 2 here
          uint4 o[8];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t(&r)[32] = half == 0 ? r0 : r1;
            const int col = cbase + 64 * jj + 32 * half;
            const float4* g4 = reinterpret_cast<const float4*>(args.ln_g + col);
            const float4* be4 = reinterpret_cast<const float4*>(args.ln_b + col);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 g0 = __ldg(g4 + 2 * c), g1 = __ldg(g4 + 2 * c + 1), b0 = __ldg(be4 + 2 * c), b1 = __ldg(be4 + 2 * c + 1);
              o[half * 4 + c].x = pack2_ln<D>(pack_u32x2(r[8 * c + 0], r[8 * c + 1]), nmean2, rstd2, pack_f32x2(g0.x, g0.y), pack_f32x2(b0.x, b0.y));
              o[half * 4 + c].y = pack2_ln<D>(pack_u32x2(r[8 * c + 2], r[8 * c + 3]), nmean2, rstd2, pack_f32x2(g0.z, g0.w), pack_f32x2(b0.z, b0.w));
              o[half * 4 + c].z = pack2_ln<D>(pack_u32x2(r[8 * c + 4], r[8 * c + 5]), nmean2, rstd2, pack_f32x2(g1.x, g1.y), pack_f32x2(b1.x, b1.y));
              o[half * 4 + c].w = pack2_ln<D>(pack_u32x2(r[8 * c + 6], r[8 * c + 7]), nmean2, rstd2, pack_f32x2(g1.z, g1.w), pack_f32x2(b1.z, b1.w));
            }
          }
          slot_writable();
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(srow_st + (((uint32_t)c ^ sw) << 4)) = o[c];
          store_from_slot(&tm_h, cbase + 64 * jj, row0);
        }
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (elected) ptx::bulk_wait<0>();
  }

  __syncwarp();
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

constexpr int pair_ln64_smem(int stages, int load) { return stages * kStage64 + (4 * load + 4) * kBox64 + 512 + 4 * kRows64 * 8; }
static_assert(pair_ln64_smem(3, 2) <= 232448 && pair_ln64_smem(2, 3) <= 232448 && pair_ln64_smem(4, 1) <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");

constexpr int kPairLinearSmem = 1024 + 5 * (kABytes + kWHalfBytes) + kSlots * kBoxBytes + 512;
constexpr int pair_ln_smem(int stages, int slots, bool split) { return stages * (kABytes + (split ? 1 : 2) * kWHalfBytes) + slots * kBoxBytes + 256 + 2 * kBM * 8; }
static_assert(pair_ln_smem(4, 6, true) <= 232448 && pair_ln_smem(3, 4, false) <= 232448 && pair_ln_smem(2, 8, false) <= 232448,
              "exceeds the 227 KB of shared memory a CTA can opt into");

template <typename K>
int set_smem(K kernel, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return fail(MP_ELAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  return MP_OK;
}

int pair_grid(int tiles) {
  int pairs = sm_count() / 2;
  if (tiles < pairs) pairs = tiles;
  return 2 * pairs;
}

}  // namespace

// Y = act(A W^T + b) on CTA pairs; called by mp_linear for N % 256 == 0 non-residual epilogues.  Y2 != NULL: Y = A W^T + b and Y2 = GELU of it.
int pair_linear(const void* A, const void* W, const float* bias, void* Y, int M, int N, int K, int epilogue, int dtype, cudaStream_t stream,
                void* Y2 = nullptr) {
  CUtensorMap ta, tw, ty, ty2;
  MP_CHECK(get_tmap(&ta, A, M, K, kBM, dtype));
  MP_CHECK(get_tmap(&tw, W, N, K, 128, dtype));
  MP_CHECK(get_tmap(&ty, Y, M, N, kBM, dtype));
  if (Y2)
    MP_CHECK(get_tmap(&ty2, Y2, M, N, kBM, dtype));
  else
    ty2 = ty;
  const bool bf = dtype == MP_DTYPE_BF16;
  // MANIPOSE_PAIR_AS=1 (K = 512, at least two 256-column blocks): the A-stationary variant (pair_linear_as_kernel).  Measured on
  // B200 at 528,768 rows it is NOT faster than the streaming kernel (qkv 690 vs 676 us, fc1 602 vs 587 us; 731 / 620 us with all
  // pairs walking N in lockstep, 721 / 607 us with 3 W stages), although it halves the operand traffic out of L2: these launches
  // are not L2-bandwidth bound, and the 4 x 16 KB of W in flight that fit beside the resident A tile hide less latency than the
  // streaming kernel's 5 x 32 KB.  Kept as an A/B switch (negative results); the default stays the streaming kernel.
  static const int as_cfg = getenv("MANIPOSE_PAIR_AS") ? atoi(getenv("MANIPOSE_PAIR_AS")) : 0;
  if (as_cfg != 0 && !Y2 && K == kAsKB * kBK && N >= 512) {
    const int grid = pair_grid((M + 255) / 256);
    auto launch_as = [&](auto kernel, int smem_bytes) -> int {
      MP_CHECK(set_smem(kernel, smem_bytes));
      launch_k(kernel, grid, kThreads, smem_bytes, stream, ta, tw, ty, bias, M, N);
      return check_launch("pair_linear_as_kernel");
    };
    const bool gelu = epilogue == MP_EPI_GELU;
#define MP_AS(ST, ROT)                                                                                                              \
  (gelu ? (bf ? launch_as(pair_linear_as_kernel<MP_EPI_GELU, Bf16, ST, ROT>, as_smem(ST)) : launch_as(pair_linear_as_kernel<MP_EPI_GELU, Fp16, ST, ROT>, as_smem(ST))) \
        : (bf ? launch_as(pair_linear_as_kernel<MP_EPI_BIAS, Bf16, ST, ROT>, as_smem(ST)) : launch_as(pair_linear_as_kernel<MP_EPI_BIAS, Fp16, ST, ROT>, as_smem(ST))))
    if (as_cfg == 2) return MP_AS(4, false);     // experiments: lockstep N order / 3 W stages
    if (as_cfg == 3) return MP_AS(3, true);
    return MP_AS(4, true);
#undef MP_AS
  }
  const int tiles = (N / 256) * ((M + 255) / 256);
  const int grid = pair_grid(tiles);
  auto launch = [&](auto kernel) -> int {
    MP_CHECK(set_smem(kernel, kPairLinearSmem));
    launch_k(kernel, grid, kThreads, kPairLinearSmem, stream, ta, tw, ty, ty2, bias, M, N, K);
    return check_launch("pair_linear_kernel");
  };
  if (Y2) return bf ? launch(pair_linear_kernel<kEpiGelu2, Bf16>) : launch(pair_linear_kernel<kEpiGelu2, Fp16>);
  if (epilogue == MP_EPI_GELU) return bf ? launch(pair_linear_kernel<MP_EPI_GELU, Bf16>) : launch(pair_linear_kernel<MP_EPI_GELU, Fp16>);
  return bf ? launch(pair_linear_kernel<MP_EPI_BIAS, Bf16>) : launch(pair_linear_kernel<MP_EPI_BIAS, Fp16>);
}

}  // namespace mp

extern "C" int mp_linear_ln(const void* A, const void* W, const float* bias, const float* resid, float* x_out, void* h_out,
                            const float* post_gamma, const float* post_beta, float post_eps, const float* pos_embed, int64_t pos_div,
                            int64_t pos_mod, const float* ln_gamma, const float* ln_beta, float ln_eps, const float* row_scale, float* x_pre,
                            int64_t M, int64_t N, int64_t K, int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(A && W && bias && resid, MP_EINVAL, "mp_linear_ln: null pointer");
  MP_REQUIRE(x_out || (post_gamma && ln_gamma && !x_pre), MP_EINVAL, "mp_linear_ln: x_out may only be omitted with both LayerNorms (h_out is the result)");
  MP_REQUIRE(N == 512, MP_EUNSUPPORTED, "mp_linear_ln: N=%lld (the fused residual + LayerNorm epilogue is built for N = 512 rows)", (long long)N);
  MP_REQUIRE(M >= 0 && M < ((int64_t)1 << 31) && K >= 64 && K % 64 == 0, MP_EINVAL, "mp_linear_ln: unsupported shape M=%lld K=%lld", (long long)M, (long long)K);
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_linear_ln: unknown dtype %d", dtype);
  MP_REQUIRE((post_gamma == nullptr) == (post_beta == nullptr) && (ln_gamma == nullptr) == (ln_beta == nullptr), MP_EINVAL,
             "mp_linear_ln: gamma/beta go together");
  MP_REQUIRE(!ln_gamma || h_out, MP_EINVAL, "mp_linear_ln: h_out required with the pre-norm");
  MP_REQUIRE(!x_pre || (post_gamma && aligned16(x_pre) && x_pre != x_out), MP_EINVAL, "mp_linear_ln: x_pre needs the post-norm and its own buffer");
  MP_REQUIRE(!pos_embed || (post_gamma && pos_div >= 1 && pos_mod >= 1), MP_EINVAL, "mp_linear_ln: pos_embed needs the post-norm and pos_div/pos_mod >= 1");
  MP_REQUIRE(aligned16(A) && aligned16(W) && aligned16(bias) && aligned16(resid) && aligned16(x_out) && aligned16(h_out) &&
                 aligned16(post_gamma) && aligned16(post_beta) && aligned16(ln_gamma) && aligned16(ln_beta) && aligned16(pos_embed),
             MP_EALIGN, "mp_linear_ln: pointers must be 16-byte aligned");
  if (M == 0) return MP_OK;
  CUtensorMap ta, tw, tr, tx, th, tp;
  MP_CHECK(get_tmap(&ta, A, M, K, kBM, dtype));
  MP_CHECK(get_tmap(&tw, W, N, K, 128, dtype));
  MP_CHECK(get_tmap(&tr, resid, M, N, kBM, 2));
  if (x_out)
    MP_CHECK(get_tmap(&tx, x_out, M, N, kBM, 2));
  else
    tx = tr;        // never stored to (LnArgs.no_x)
  if (ln_gamma)
    MP_CHECK(get_tmap(&th, h_out, M, N, kBM, dtype));
  else
    th = tx;
  if (x_pre)
    MP_CHECK(get_tmap(&tp, x_pre, M, N, kBM, 2));
  else
    tp = tx;
  LnArgs args{bias, post_gamma, post_beta, pos_embed, ln_gamma, ln_beta, x_pre != nullptr, row_scale, x_out == nullptr, post_eps, ln_eps, (int)pos_div, (int)pos_mod};
  const int tiles = (int)((M + 255) / 256);
  const int grid = pair_grid(tiles);
  // 128-row kernels (MANIPOSE_LN_CFG=1 / 2 / 3): split accumulation with 4 x 32 KB operand stages + 6 box slots; single accumulation
  // with 3 x 48 KB stages + 4 slots (see the kernel's header comment)
  static const int cfg = getenv("MANIPOSE_LN_CFG") ? atoi(getenv("MANIPOSE_LN_CFG")) : 0;
  auto launch = [&](auto kernel, int smem_bytes) -> int {
    MP_CHECK(set_smem(kernel, smem_bytes));
    launch_k(kernel, grid, kThreads, smem_bytes, (cudaStream_t)stream, ta, tw, tr, tx, th, tp, args, (int)M, (int)K);
    return check_launch("pair_linear_ln_kernel");
  };
  const bool bf = dtype == MP_DTYPE_BF16;
  // Default: 64-row tiles with two TMEM accumulators (the main loop of the next tile runs under the epilogue), 3 operand stages + 2
  // residual slots per column group.  Measured at 528,768 rows (B200, L2 flushed): proj + norm2 539 us (6.03 TB/s = 0.94 of the copy
  // bandwidth; the 128-row split-accumulation kernel: 580 us), fc2 + post-norm + norm1 750 us (5.05 TB/s = 0.78; the 128-row kernel:
  // 942 us).  MANIPOSE_LN_CFG = 5: 2 stages + 3 residual slots (slower: 602 / 891 us); 1 / 2 / 3: the 128-row kernels (split /
  // single accumulation / 2 stages + 8 slots), kept for A/B measurements.
  if (cfg == 0 || cfg == 4 || cfg == 5 || cfg == 6) {
    CUtensorMap ta6, tr6, tx6, th6, tp6;
    MP_CHECK(get_tmap(&ta6, A, M, K, kRows64, dtype));
    MP_CHECK(get_tmap(&tr6, resid, M, N, kRows64, 2));
    if (x_out)
      MP_CHECK(get_tmap(&tx6, x_out, M, N, kRows64, 2));
    else
      tx6 = tr6;
    if (ln_gamma)
      MP_CHECK(get_tmap(&th6, h_out, M, N, kRows64, dtype));
    else
      th6 = tx6;
    if (x_pre)
      MP_CHECK(get_tmap(&tp6, x_pre, M, N, kRows64, 2));
    else
      tp6 = tx6;
    const int grid64 = pair_grid((int)((M + 127) / 128));
    auto launch64 = [&](auto kernel, int smem_bytes) -> int {
      MP_CHECK(set_smem(kernel, smem_bytes));
      launch_k(kernel, grid64, kThreads, smem_bytes, (cudaStream_t)stream, ta6, tw, tr6, tx6, th6, tp6, args, (int)M, (int)K);
      return check_launch("pair_linear_ln64_kernel");
    };
    if (cfg == 6)   // 4 operand stages + ONE residual slot per column group: slower (608 / 774 us against 539 / 729 us)
      return bf ? launch64(pair_linear_ln64_kernel<Bf16, 4, 1>, pair_ln64_smem(4, 1)) : launch64(pair_linear_ln64_kernel<Fp16, 4, 1>, pair_ln64_smem(4, 1));
    if (cfg == 5)
      return bf ? launch64(pair_linear_ln64_kernel<Bf16, 2, 3>, pair_ln64_smem(2, 3)) : launch64(pair_linear_ln64_kernel<Fp16, 2, 3>, pair_ln64_smem(2, 3));
    return bf ? launch64(pair_linear_ln64_kernel<Bf16, 3, 2>, pair_ln64_smem(3, 2)) : launch64(pair_linear_ln64_kernel<Fp16, 3, 2>, pair_ln64_smem(3, 2));
  }
  if (cfg == 1 || (cfg == 9 && K < 1024))      // 9: the round-1 defaults (split accumulation for K = 512, single for K = 1024)
    return bf ? launch(pair_linear_ln_kernel<Bf16, 4, 6, true>, pair_ln_smem(4, 6, true)) : launch(pair_linear_ln_kernel<Fp16, 4, 6, true>, pair_ln_smem(4, 6, true));
  if (cfg == 3)      // experiment: 2 x 48 KB operand stages + 8 box slots (more of the residual prefetched under the main loop)
    return bf ? launch(pair_linear_ln_kernel<Bf16, 2, 8, false>, pair_ln_smem(2, 8, false)) : launch(pair_linear_ln_kernel<Fp16, 2, 8, false>, pair_ln_smem(2, 8, false));
  return bf ? launch(pair_linear_ln_kernel<Bf16, 3, 4, false>, pair_ln_smem(3, 4, false)) : launch(pair_linear_ln_kernel<Fp16, 3, 4, false>, pair_ln_smem(3, 4, false));
}

extern "C" int mp_linear_gelu2(const void* A, const void* W, const float* bias, void* U, void* G, int64_t M, int64_t N, int64_t K, int dtype,
                               mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(A && W && bias && U && G && U != G, MP_EINVAL, "mp_linear_gelu2: null pointer / aliased outputs");
  MP_REQUIRE(M >= 0 && M < ((int64_t)1 << 31) && N >= 256 && N % 256 == 0 && K >= 64 && K % 64 == 0, MP_EINVAL,
             "mp_linear_gelu2: unsupported shape M=%lld N=%lld K=%lld (N %% 256 == 0, K %% 64 == 0)", (long long)M, (long long)N, (long long)K);
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_linear_gelu2: unknown dtype %d", dtype);
  MP_REQUIRE(aligned16(A) && aligned16(W) && aligned16(bias) && aligned16(U) && aligned16(G), MP_EALIGN,
             "mp_linear_gelu2: pointers must be 16-byte aligned");
  if (M == 0) return MP_OK;
  return pair_linear(A, W, bias, U, (int)M, (int)N, (int)K, MP_EPI_BIAS, dtype, (cudaStream_t)stream, G);
}

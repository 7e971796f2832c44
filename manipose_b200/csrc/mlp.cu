// The MLP branch of a C = 512 block in ONE kernel (Mlp.forward + the residual add of Block.forward, mix_ste.py:216-222, 356-358, and the
// LayerNorms that follow it, :143,149,154,166,170 / :353):
//
//     x_out = [LN_post](x + fc2(GELU(fc1(h)))) [+ pos],   h_out = LN_pre(x_out) as 16-bit
//
// The 1024-wide hidden activation never leaves the SM: 4 KB per token less HBM traffic per block than fc1 -> HBM -> fc2, one launch
// instead of two.  A CTA pair (tcgen05.mma.cta_group::2, M = 128) owns 128 rows per tile, 64 per CTA; a [64 x 256] fp32 result of an
// N = 256 instruction is folded over the 128 TMEM lanes into 128 columns, so TMEM (512 columns) holds
//     O   [64 x 512] the fc2 accumulator        columns [0, 256)
//     H0, H1 [64 x 256] two fc1 chunk accumulators columns [256, 384), [384, 512).
// Per tile the MMA thread issues F1(0) F1(1) F2(0) F1(2) F2(1) F1(3) F2(2) F2(3): F1(c) = hidden chunk c (columns [256 c, +256)) from the
// RESIDENT 16-bit h tile (64 KB) and W1 rows streamed through a ring of 16 KB slots; the eight epilogue warps turn H into
// GELU(H + b1) as 16-bit in shared memory (one [64 x 256] chunk, the A operand of F2, 32 KB); F2(c) accumulates that chunk against the
// W2 columns [256 c, +256) into O.  After F2(3) the same warps run the residual + LayerNorm epilogue of pair_linear_ln64_kernel
// (gemm2.cu) on O, while the tensor pipe already works on F1(0), F1(1) of the next tile.
//
// Roles per CTA (384 threads): warp 0 W producer, warp 1 MMA issuer (leader CTA), warp 2 TMEM allocator + loader of the h tile and
// of the residual boxes, warps 4-11 epilogue.  Barriers the leader's MMA thread waits on live in the leader CTA and are arrived on
// remotely by the peer (cluster-scope release / acquire where the payload is shared memory written by ordinary stores).
#include <stdlib.h>

#include "common.cuh"
#include "ln_args.cuh"
#include "ptx.cuh"

namespace mp {

int get_tmap(CUtensorMap* out, const void* ptr, int64_t rows, int64_t cols, int box_rows, int type);   // gemm.cu

namespace {

constexpr int kThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int kRows = 64;                        // rows per CTA (128 per pair)
constexpr int kBK = 64;
constexpr int kC = 512, kHid = 1024;
constexpr int kChunk = 256;                      // hidden columns per fc1 chunk
constexpr int kNChunks = kHid / kChunk;          // 4
constexpr int kBox = kRows * 128;                // 8 KB: 64 rows x 128 bytes
constexpr int kWSlot = 128 * kBK * 2;            // 16 KB: 128 weight rows x one 64-wide k-block
constexpr int kWSlots = 4;
constexpr int kHTile = (kC / kBK) * kBox;        // 64 KB: the CTA's 64 rows of h, 8 k-blocks
constexpr int kHidTile = (kChunk / kBK) * kBox;  // 32 KB: one GELU'd hidden chunk, 4 k-blocks
constexpr int kSmem = kHTile + kHidTile + kWSlots * kWSlot + 8 * kBox + 512 + 4 * kRows * 8;
static_assert(kSmem <= 232448, "exceeds the 227 KB of shared memory a CTA can opt into");

__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// remote arrive that publishes this CTA's ordinary shared-memory stores to the cluster (the peer's tensor core reads them)
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0, ok = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) break;
    if (++spins > (1u << 24)) {
      printf("libmanipose_sm100: mbarrier timeout (mlp, block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
}

// issue order of a tile: chunk + 4 * (is F2)
__device__ __constant__ int kOrder[8] = {0, 1, 4, 2, 5, 3, 6, 7};

template <typename D>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
mlp_ln64_kernel(const __grid_constant__ CUtensorMap tm_hin, const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
                const __grid_constant__ CUtensorMap tm_r, const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_h,
                const float* __restrict__ b1, LnArgs args, int M) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* h_base = smem;                                        // [8][64 x 64] 16-bit, K-major, 128-byte swizzle
  uint8_t* hid_base = h_base + kHTile;                           // [4][64 x 64] 16-bit
  uint8_t* w_base = hid_base + kHidTile;                         // [kWSlots][128 x 64] 16-bit
  uint8_t* l_base = w_base + kWSlots * kWSlot;                   // [4] residual boxes 0, 2 of a column group
  uint8_t* s_base = l_base + 4 * kBox;                           // [4] residual boxes 1, 3, then the output staging box
  uint64_t* w_full = reinterpret_cast<uint64_t*>(s_base + 4 * kBox);   // [kWSlots] leader
  uint64_t* w_empty = w_full + kWSlots;                          // [kWSlots] each CTA (multicast commit)
  uint64_t* h_full = w_empty + kWSlots;                          // leader: h tiles of both CTAs landed
  uint64_t* h_empty = h_full + 1;                                // each CTA: F1(3) has read the h tile
  uint64_t* hacc_full = h_empty + 1;                             // [2] each CTA: H_b complete
  uint64_t* hacc_empty = hacc_full + 2;                          // [2] leader: H_b read by the 16 epilogue warps of the pair
  uint64_t* hid_full = hacc_empty + 2;                           // leader: the hidden chunk of both CTAs is in shared memory
  uint64_t* hid_empty = hid_full + 1;                            // each CTA: F2(c) has read the hidden chunk
  uint64_t* o_full = hid_empty + 1;                              // each CTA: O complete
  uint64_t* o_empty = o_full + 1;                                // leader: O read by the 16 epilogue warps of the pair
  uint64_t* l_full = o_empty + 1;                                // [4]
  uint64_t* l_empty = l_full + 4;                                // [4]
  uint64_t* s_full = l_empty + 4;                                // [4]
  uint64_t* s_empty = s_full + 4;                                // [4] two completions per tile: box 1 consumed, stores done (end of tile)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(s_empty + 4);
  float2* sx = reinterpret_cast<float2*>(tmem_holder + 4);       // [4][64] row statistics exchange

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const bool leader = rank == 0;
  const int num_tiles = (M + 2 * kRows - 1) / (2 * kRows);
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const bool has_post = args.post_g != nullptr, has_ln = args.ln_g != nullptr;
  const bool store_a = !has_post;                                // pass A hands its value to a TMA store (x_out) when no post-norm follows
  const bool store_b = has_post && args.no_x == 0;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tm_hin);
    ptx::prefetch_tmap(&tm_w1);
    ptx::prefetch_tmap(&tm_w2);
    ptx::prefetch_tmap(&tm_r);
    ptx::prefetch_tmap(&tm_x);
    if (has_ln) ptx::prefetch_tmap(&tm_h);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kWSlots; ++s) {
      ptx::mbar_init(&w_full[s], 1);
      ptx::mbar_init(&w_empty[s], 1);
    }
    ptx::mbar_init(h_full, 1);
    ptx::mbar_init(h_empty, 1);
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&hacc_full[b], 1);
      ptx::mbar_init(&hacc_empty[b], 2 * kEpiWarps);
    }
    ptx::mbar_init(hid_full, 2 * kEpiWarps);
    ptx::mbar_init(hid_empty, 1);
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(o_empty, 2 * kEpiWarps);
    for (int s = 0; s < 4; ++s) {
      ptx::mbar_init(&l_full[s], 1);
      ptx::mbar_init(&l_empty[s], 1);
      ptx::mbar_init(&s_full[s], 1);
      ptx::mbar_init(&s_empty[s], 1);
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

  if (warp == 0) {
    if (lane == 0) {
      // ===================== W producer: 16 KB slots in the MMA thread's issue order =====================
      int slot = 0;
      uint32_t phase = 0;
      auto load_w = [&](const CUtensorMap* map, int col, int row_w) {
        ptx::mbar_wait(&w_empty[slot], phase ^ 1);
        const uint32_t full_leader = ptx::mapa_shared(smem_u32(&w_full[slot]), 0);
        if (leader) ptx::mbar_expect_tx(&w_full[slot], 2 * kWSlot);      // the peer's half completes on this barrier too
        ptx::tma_load_2d_pair(w_base + slot * kWSlot, map, full_leader, col, row_w);
        if (++slot == kWSlots) {
          slot = 0;
          phase ^= 1;
        }
      };
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs) {
        for (int s = 0; s < 8; ++s) {
          const int c = kOrder[s] & 3;
          if (kOrder[s] < 4) {
            for (int kb = 0; kb < kC / kBK; ++kb) load_w(&tm_w1, kb * kBK, kChunk * c + (int)rank * 128);
          } else {
            for (int kk = 0; kk < kChunk / kBK; ++kk) {
              load_w(&tm_w2, kChunk * c + kk * kBK, (int)rank * 128);           // output columns [0, 256)
              load_w(&tm_w2, kChunk * c + kk * kBK, 256 + (int)rank * 128);     // output columns [256, 512)
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (leader && lane == 0) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc = ptx::umma_idesc_16(2 * kRows, 256, D::kUmmaFmt);
      int slot = 0;
      uint32_t phase = 0, tt = 0;
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++tt) {
        mbar_wait_cluster(h_full, tt & 1);
        ptx::tc_fence_after();
        for (int s = 0; s < 8; ++s) {
          const int c = kOrder[s] & 3;
          const uint32_t g = 4u * tt + (uint32_t)c;                // running chunk index
          if (kOrder[s] < 4) {
            // ---- F1(c): H_b = h W1[256 c : 256 c + 256]^T
            const uint32_t b = g & 1, u = g >> 1;
            if (g >= 2) {
              ptx::mbar_wait(&hacc_empty[b], (u - 1) & 1);
              ptx::tc_fence_after();
            }
            const uint32_t d_h = tmem_base + 256u + 128u * b;
            for (int kb = 0; kb < kC / kBK; ++kb) {
              ptx::mbar_wait(&w_full[slot], phase);
              ptx::tc_fence_after();
              const uint64_t da = ptx::umma_desc_sw128(smem_u32(h_base + kb * kBox));
              const uint64_t db = ptx::umma_desc_sw128(smem_u32(w_base + slot * kWSlot));
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k)
                ptx::umma_f16_pair(d_h, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
              ptx::umma_commit_pair(&w_empty[slot]);
              if (++slot == kWSlots) {
                slot = 0;
                phase ^= 1;
              }
            }
            ptx::umma_commit_pair(&hacc_full[b]);
            if (c == kNChunks - 1) ptx::umma_commit_pair(h_empty);   // the h tile may be replaced by the next tile's
          } else {
            // ---- F2(c): O += GELU(H_c) W2[:, 256 c : 256 c + 256]^T
            mbar_wait_cluster(hid_full, g & 1);
            if (c == 0 && tt > 0) ptx::mbar_wait(o_empty, (tt - 1) & 1);
            ptx::tc_fence_after();
            for (int kk = 0; kk < kChunk / kBK; ++kk) {
              const int s0 = slot, s1 = slot + 1 == kWSlots ? 0 : slot + 1;
              const uint32_t p0 = phase, p1 = slot + 1 == kWSlots ? phase ^ 1 : phase;
              ptx::mbar_wait(&w_full[s0], p0);
              ptx::mbar_wait(&w_full[s1], p1);
              ptx::tc_fence_after();
              const uint64_t da = ptx::umma_desc_sw128(smem_u32(hid_base + kk * kBox));
              const uint64_t db0 = ptx::umma_desc_sw128(smem_u32(w_base + s0 * kWSlot));
              const uint64_t db1 = ptx::umma_desc_sw128(smem_u32(w_base + s1 * kWSlot));
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k) {
                const uint32_t accum = (c | kk | k) != 0 ? 1u : 0u;
                ptx::umma_f16_pair(tmem_base, da + (uint64_t)(2 * k), db0 + (uint64_t)(2 * k), idesc, accum);
                ptx::umma_f16_pair(tmem_base + 128u, da + (uint64_t)(2 * k), db1 + (uint64_t)(2 * k), idesc, accum);
              }
              ptx::umma_commit_pair(&w_empty[s0]);
              ptx::umma_commit_pair(&w_empty[s1]);
              slot = s1 + 1 == kWSlots ? 0 : s1 + 1;
              phase = s1 + 1 == kWSlots ? p1 ^ 1 : p1;
            }
            ptx::umma_commit_pair(hid_empty);
            if (c == kNChunks - 1) ptx::umma_commit_pair(o_full);
          }
        }
      }
    }
  } else if (warp == 2) {
    if (lane == 0) {
      // ===================== loader: h tile + residual boxes =====================
      auto load_h = [&](int tile) {
        const int row0 = tile * 2 * kRows + (int)rank * kRows;
        const uint32_t full_leader = ptx::mapa_shared(smem_u32(h_full), 0);
        if (leader) ptx::mbar_expect_tx(h_full, 2 * kHTile);
        for (int kb = 0; kb < kC / kBK; ++kb) ptx::tma_load_2d_pair(h_base + kb * kBox, &tm_hin, full_leader, kb * kBK, row0);
      };
      auto box_col = [](int cg, int j) { return (cg >> 1) * 256 + (cg & 1) * 128 + j * 32; };
      uint32_t tt = 0;
      if (pair_id < num_tiles) load_h(pair_id);
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++tt) {
        const int row0 = tile * 2 * kRows + (int)rank * kRows;
        for (int cg = 0; cg < 4; ++cg) {                           // boxes 0 and 1 of every column group
          if (tt > 0) ptx::mbar_wait(&l_empty[cg], (2 * tt - 1) & 1);      // box 2 of the previous tile consumed
          ptx::mbar_expect_tx(&l_full[cg], kBox);
          ptx::tma_load_2d(l_base + cg * kBox, &tm_r, &l_full[cg], box_col(cg, 0), row0);
          if (tt > 0) ptx::mbar_wait(&s_empty[cg], (2 * tt - 1) & 1);      // the previous tile's stores have left the slot
          ptx::mbar_expect_tx(&s_full[cg], kBox);
          ptx::tma_load_2d(s_base + cg * kBox, &tm_r, &s_full[cg], box_col(cg, 1), row0);
        }
        if (tile + num_pairs < num_tiles) {                        // the next tile's h as soon as F1(3) has read this one
          ptx::mbar_wait(h_empty, tt & 1);
          load_h(tile + num_pairs);
        }
        for (int cg = 0; cg < 4; ++cg) {                           // boxes 2 and 3: behind boxes 0 and 1 in the same slots
          ptx::mbar_wait(&l_empty[cg], (2 * tt) & 1);
          ptx::mbar_expect_tx(&l_full[cg], kBox);
          ptx::tma_load_2d(l_base + cg * kBox, &tm_r, &l_full[cg], box_col(cg, 2), row0);
          ptx::mbar_wait(&s_empty[cg], (2 * tt) & 1);                       // box 1 consumed
          ptx::mbar_expect_tx(&s_full[cg], kBox);
          ptx::tma_load_2d(s_base + cg * kBox, &tm_r, &s_full[cg], box_col(cg, 3), row0);
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: GELU of the four hidden chunks, then residual + LayerNorms on O =====================
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const int hh = q >> 1;
    const int cg = 2 * grp + hh;                          // column group of the LayerNorm epilogue: output columns [cbase, cbase + 128)
    const int cbase = 256 * grp + 128 * hh;
    const int row = 32 * (q & 1) + lane;                  // row of the CTA's 64
    const bool elected = (q & 1) == 0 && lane == 0;
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t t_lane = tmem_base + ((uint32_t)(32 * q) << 16);
    const uint32_t t_row = t_lane + (uint32_t)(128 * grp);          // O: + 32 j
    uint8_t* lslot = l_base + cg * kBox;
    uint8_t* sslot = s_base + cg * kBox;
    uint8_t* srow_st = sslot + row * 128;
    // hidden chunk: this thread holds the 64 consecutive hidden columns [128 hh + 64 grp, + 64) of its row = one row of k-block 2 hh + grp
    uint8_t* hid_row = hid_base + (2 * hh + grp) * kBox + row * 128;
    bool pending = false;
    uint32_t tt = 0;

    auto row_stats = [&](float mean_q, float m2_q, float eps, float& mean, float& rstd) {
      sx[cg * kRows + row] = make_float2(mean_q, m2_q);
      named_bar_sync(5, 256);
      float2 p[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) p[i] = sx[i * kRows + row];
      named_bar_sync(5, 256);
      mean = 0.25f * ((p[0].x + p[1].x) + (p[2].x + p[3].x));
      float m2 = (p[0].y + p[1].y) + (p[2].y + p[3].y);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float d = p[i].x - mean;
        m2 = fmaf(128.0f * d, d, m2);
      }
      rstd = rsqrtf(m2 * (1.0f / 512.0f) + eps);
    };
    auto store_from_slot = [&](const CUtensorMap* map, int col, int row0) {
      ptx::fence_proxy_async_smem();
      named_bar_sync(1 + cg, 64);
      if (elected) {
        ptx::tma_store_2d(map, sslot, col, row0);
        ptx::bulk_commit();
        pending = true;
      }
    };
    auto slot_writable = [&]() {
      if (elected && pending) {
        ptx::bulk_wait_read<0>();
        pending = false;
      }
      named_bar_sync(1 + cg, 64);
    };
    const int last_pass = has_ln ? 2 : (has_post ? 1 : 0);
    auto release_o = [&]() {
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_shared(smem_u32(o_empty), 0));
    };

    for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++tt) {
      const int row0 = tile * 2 * kRows + (int)rank * kRows;
      const int grow = row0 + row;

      // ---- the four hidden chunks: H_b -> GELU(H_b + b1) as 16-bit -> shared memory (A operand of F2)
#pragma unroll 1
      for (int c = 0; c < kNChunks; ++c) {
        const uint32_t g = 4u * tt + (uint32_t)c, b = g & 1, u = g >> 1;
        ptx::mbar_wait(&hacc_full[b], u & 1);
        ptx::tc_fence_after();
        uint32_t r0[32], r1[32];
        const uint32_t t_h = t_lane + 256u + 128u * b + (uint32_t)(64 * grp);
        ptx::tmem_ld32(t_h, r0);
        ptx::tmem_ld32(t_h + 32u, r1);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa_shared(smem_u32(&hacc_empty[b]), 0));
        uint4 o[8];
        const float* bias1 = b1 + kChunk * c + 128 * hh + 64 * grp;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const uint32_t(&r)[32] = half == 0 ? r0 : r1;
          const float4* b4 = reinterpret_cast<const float4*>(bias1 + half * 32);
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            const float4 ba = __ldg(b4 + 2 * cc), bb = __ldg(b4 + 2 * cc + 1);
            float f[8] = {__uint_as_float(r[8 * cc + 0]) + ba.x, __uint_as_float(r[8 * cc + 1]) + ba.y,
                          __uint_as_float(r[8 * cc + 2]) + ba.z, __uint_as_float(r[8 * cc + 3]) + ba.w,
                          __uint_as_float(r[8 * cc + 4]) + bb.x, __uint_as_float(r[8 * cc + 5]) + bb.y,
                          __uint_as_float(r[8 * cc + 6]) + bb.z, __uint_as_float(r[8 * cc + 7]) + bb.w};
#pragma unroll
            for (int e = 0; e < 8; e += 2) gelu_erf2(f[e], f[e + 1]);
            o[half * 4 + cc].x = D::pack2(f[0], f[1]);
            o[half * 4 + cc].y = D::pack2(f[2], f[3]);
            o[half * 4 + cc].z = D::pack2(f[4], f[5]);
            o[half * 4 + cc].w = D::pack2(f[6], f[7]);
          }
        }
        if (g >= 1) ptx::mbar_wait(hid_empty, (g - 1) & 1);       // F2 of the previous chunk has read the hidden tile
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) *reinterpret_cast<uint4*>(hid_row + (((uint32_t)cc ^ sw) << 4)) = o[cc];
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster_release(ptx::mapa_shared(smem_u32(hid_full), 0));
      }

      // ---- O is complete: pass A: v = resid + (acc + b2); row statistics; v back to TMEM (and out, if no post-norm)
      ptx::mbar_wait(o_full, tt & 1);
      ptx::tc_fence_after();
      float shift = 0.f, s1 = 0.f, s2 = 0.f;
      uint64_t s1p = 0, s2p = 0, nshift2 = 0, nmean2 = 0, rstd2 = 0;   // packed accumulators / broadcast operands (common.cuh)
#pragma unroll 1
      for (int j = 0; j < 4; ++j) {
        const bool in_s = (j & 1) != 0;
        const uint32_t use = 2u * tt + (uint32_t)(j >> 1);       // load use of the slot
        const int col = cbase + 32 * j;
        uint8_t* lrow = (in_s ? sslot : lslot) + row * 128;
        uint32_t r[32];
        ptx::tmem_ld32(t_row + (uint32_t)(32 * j), r);
        ptx::mbar_wait(in_s ? &s_full[cg] : &l_full[cg], use & 1);
        ptx::tmem_ld_wait();
        const float4* b4 = reinterpret_cast<const float4*>(args.bias + col);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          float4* p = reinterpret_cast<float4*>(lrow + (((uint32_t)c ^ sw) << 4));
          const float4 bb = __ldg(b4 + c);
          float4 v = *p;
          // packed fp32 (FADD2 / FFMA2: two elements per issue slot), the same roundings as pair_linear_ln64_kernel with scale 1
          const uint64_t v01 = add_f32x2(add_f32x2(pack_u32x2(r[4 * c + 0], r[4 * c + 1]), pack_f32x2(bb.x, bb.y)), pack_f32x2(v.x, v.y));
          const uint64_t v23 = add_f32x2(add_f32x2(pack_u32x2(r[4 * c + 2], r[4 * c + 3]), pack_f32x2(bb.z, bb.w)), pack_f32x2(v.z, v.w));
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
        named_bar_sync(1 + cg, 64);
        if (elected) {
          if (store_a) {
            ptx::tma_store_2d(&tm_x, in_s ? sslot : lslot, col, row0);
            ptx::bulk_commit();
            ptx::bulk_wait_read<0>();
          }
          // Two completions per tile and barrier, each gated by a load the loader issues only after it has seen the previous one (a
          // parity wait cannot tell completions two apart): L: box 0 / box 2 consumed; S: box 1 consumed / end of the tile (below) -
          // box 3 leaving the S slot needs no event of its own.
          if (j != 3) ptx::mbar_arrive(in_s ? &s_empty[cg] : &l_empty[cg]);
        }
      }
      if (last_pass == 0) release_o();
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
        const float* pos_row = args.pos ? args.pos + (size_t)((grow / args.pos_div) % args.pos_mod) * kC : nullptr;
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
        if (last_pass == 1) release_o();
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
          if (jj == 1) release_o();                         // last TMEM read of the tile: F2(0) of the next tile may start
          uint4 o[8];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t(&r)[32] = half == 0 ? r0 : r1;
            const int col = cbase + 64 * jj + 32 * half;
            const float4* g4 = reinterpret_cast<const float4*>(args.ln_g + col);
            const float4* be4 = reinterpret_cast<const float4*>(args.ln_b + col);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 g0 = __ldg(g4 + 2 * c), g1 = __ldg(g4 + 2 * c + 1), b0 = __ldg(be4 + 2 * c), bq = __ldg(be4 + 2 * c + 1);
              o[half * 4 + c].x = pack2_ln<D>(pack_u32x2(r[8 * c + 0], r[8 * c + 1]), nmean2, rstd2, pack_f32x2(g0.x, g0.y), pack_f32x2(b0.x, b0.y));
              o[half * 4 + c].y = pack2_ln<D>(pack_u32x2(r[8 * c + 2], r[8 * c + 3]), nmean2, rstd2, pack_f32x2(g0.z, g0.w), pack_f32x2(b0.z, b0.w));
              o[half * 4 + c].z = pack2_ln<D>(pack_u32x2(r[8 * c + 4], r[8 * c + 5]), nmean2, rstd2, pack_f32x2(g1.x, g1.y), pack_f32x2(bq.x, bq.y));
              o[half * 4 + c].w = pack2_ln<D>(pack_u32x2(r[8 * c + 6], r[8 * c + 7]), nmean2, rstd2, pack_f32x2(g1.z, g1.w), pack_f32x2(bq.z, bq.w));
            }
          }
          slot_writable();
#pragma unroll
          for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(srow_st + (((uint32_t)c ^ sw) << 4)) = o[c];
          store_from_slot(&tm_h, cbase + 64 * jj, row0);
        }
      }
      // ---- end of tile: the store slot goes back to the loader (box 1 of the next tile) once its last store has read it
      if (elected) {
        if (pending) {
          ptx::bulk_wait_read<0>();
          pending = false;
        }
        ptx::mbar_arrive(&s_empty[cg]);
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

}  // namespace
}  // namespace mp

extern "C" int mp_mlp_ln(const void* h_in, const void* W1, const float* b1, const void* W2, const float* b2, const float* resid, float* x_out,
                         void* h_out, const float* post_gamma, const float* post_beta, float post_eps, const float* pos_embed, int64_t pos_div,
                         int64_t pos_mod, const float* ln_gamma, const float* ln_beta, float ln_eps, int64_t M, int64_t C, int64_t hidden, int dtype,
                         mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(h_in && W1 && b1 && W2 && b2 && resid, MP_EINVAL, "mp_mlp_ln: null pointer");
  MP_REQUIRE(C == kC && hidden == kHid, MP_EUNSUPPORTED, "mp_mlp_ln: C=%lld hidden=%lld (the fused MLP is built for 512 -> 1024 -> 512)", (long long)C,
             (long long)hidden);
  MP_REQUIRE(M >= 0 && M < ((int64_t)1 << 31), MP_EINVAL, "mp_mlp_ln: unsupported M=%lld", (long long)M);
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_mlp_ln: unknown dtype %d", dtype);
  MP_REQUIRE((post_gamma == nullptr) == (post_beta == nullptr) && (ln_gamma == nullptr) == (ln_beta == nullptr), MP_EINVAL,
             "mp_mlp_ln: gamma/beta go together");
  MP_REQUIRE(!ln_gamma || h_out, MP_EINVAL, "mp_mlp_ln: h_out required with the pre-norm");
  MP_REQUIRE(x_out || (post_gamma && ln_gamma), MP_EINVAL, "mp_mlp_ln: x_out may only be omitted with both LayerNorms (h_out is the result)");
  MP_REQUIRE(!pos_embed || (post_gamma && pos_div >= 1 && pos_mod >= 1), MP_EINVAL, "mp_mlp_ln: pos_embed needs the post-norm and pos_div/pos_mod >= 1");
  MP_REQUIRE(aligned16(h_in) && aligned16(W1) && aligned16(b1) && aligned16(W2) && aligned16(b2) && aligned16(resid) && aligned16(x_out) &&
                 aligned16(h_out) && aligned16(post_gamma) && aligned16(post_beta) && aligned16(ln_gamma) && aligned16(ln_beta) &&
                 aligned16(pos_embed),
             MP_EALIGN, "mp_mlp_ln: pointers must be 16-byte aligned");
  if (M == 0) return MP_OK;
  CUtensorMap th_in, tw1, tw2, tr, tx, th;
  MP_CHECK(get_tmap(&th_in, h_in, M, kC, kRows, dtype));
  MP_CHECK(get_tmap(&tw1, W1, kHid, kC, 128, dtype));
  MP_CHECK(get_tmap(&tw2, W2, kC, kHid, 128, dtype));
  MP_CHECK(get_tmap(&tr, resid, M, kC, kRows, 2));
  if (x_out)
    MP_CHECK(get_tmap(&tx, x_out, M, kC, kRows, 2));
  else
    tx = tr;
  if (ln_gamma)
    MP_CHECK(get_tmap(&th, h_out, M, kC, kRows, dtype));
  else
    th = tx;
  LnArgs args{b2, post_gamma, post_beta, pos_embed, ln_gamma, ln_beta, 0, nullptr, x_out == nullptr, post_eps, ln_eps, (int)pos_div, (int)pos_mod};
  const int tiles = (int)((M + 2 * kRows - 1) / (2 * kRows));
  int pairs = sm_count() / 2;
  if (tiles < pairs) pairs = tiles;
  auto launch = [&](auto kernel) -> int {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return fail(MP_ELAUNCH, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    kernel<<<2 * pairs, kThreads, kSmem, (cudaStream_t)stream>>>(th_in, tw1, tw2, tr, tx, th, b1, args, (int)M);
    return check_launch("mlp_ln64_kernel");
  };
  return dtype == MP_DTYPE_BF16 ? launch(mlp_ln64_kernel<Bf16>) : launch(mlp_ln64_kernel<Fp16>);
}

// Multi-head softmax attention of the MixSTE blocks (Attention.forward, hpe/mh_so3_hpe/architectures/mix_ste.py:255-282)
// over ONE activation layout, [clip, frame, token, 3C] with columns [q | k | v] x heads x head_dim (mix_ste.py:257-261):
//   spatial  (STEblocks): sequences are the n_tok (17 joints / 16 segments) tokens of one frame;
//   temporal (TTEblocks): sequences are the n_frames frames of one (clip, token) track, read with a row stride of
//                         n_tok*3C elements -- the reference's "(B L) J C -> (B J) L C" rearranges (mix_ste.py:144,167,171)
//                         never materialise.
// Both kernels keep Q, K and V of their sequences in shared memory (16-byte cp.async, +16-byte row padding so ldmatrix
// is bank-conflict free), compute S = QK^T and O = PV on the tensor cores (mma.sync m16n8k16 bf16, fp32 accumulate) with
// an online softmax in the exp2 domain, and write O through shared memory as 16-byte coalesced rows.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace mp {

int get_tmap_clip_rows(CUtensorMap* out, const void* ptr, int64_t n_clips, int64_t rows_per_clip, int64_t cols, int box_rows, int type);  // gemm.cu
int get_tmap_track(CUtensorMap* out, const void* ptr, int64_t n_clips, int64_t n_frames, int64_t n_tok, int64_t cols, int box_frames, int type,
                   int box_tok = 1);  // gemm.cu

namespace {

constexpr float kLog2e = 1.4426950408889634f;

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// One 16-row query tile against keys [0, n_keys_pad) held in shared memory with a constant row stride of LD bytes (rows
// [n_keys, n_keys_pad) must be zero).  q_addr / k_addr / v_addr are shared-window addresses of row 0 of the tile / of key 0.
// Every ldmatrix address is base + per-lane offset + compile-time constant.  Softmax runs in the exp2 domain on the RAW
// scores (the scale is folded into one FMA per element); keys are masked only in the chunk that contains padding.
// Result: o[HD/8][4] (unnormalised) and l[2] (row sums) in the mma C-fragment layout (rows g and g+8).
template <int HD, int LD, typename D>
__device__ __forceinline__ void attend_tile(uint32_t q_addr, uint32_t k_addr, uint32_t v_addr, int n_keys, int n_keys_pad,
                                            float scale_log2, int lane, float (&o)[HD / 8][4], float (&l)[2]) {
  constexpr int KS = HD / 16;   // k-steps over the head dimension
  constexpr int NT = HD / 8;    // output n-tiles
  const int t = lane & 3;
  // per-lane offsets of the three ldmatrix address patterns
  const uint32_t q_off = (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * LD + (lane >> 4) * 16);   // A: {r0-7,klo},{r8-15,klo},{r0-7,khi},{r8-15,khi}
  const uint32_t k_off = (uint32_t)(((lane & 7) + (lane >> 4) * 8) * LD + ((lane >> 3) & 1) * 16);   // B: {n0-7,klo},{n0-7,khi},{n8-15,klo},{n8-15,khi}
  const uint32_t v_off = q_off;                                                                       // B^T via .trans: same pattern as A

  uint32_t qf[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) ldsm_x4(qf[ks], q_addr + q_off + ks * 32);

  float m[2] = {-INFINITY, -INFINITY};
  l[0] = l[1] = 0.f;
#pragma unroll
  for (int i = 0; i < NT; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;

  for (int key0 = 0; key0 < n_keys_pad; key0 += 32) {
    const uint32_t kc_addr = k_addr + k_off + (uint32_t)(key0 * LD);
    const uint32_t vc_addr = v_addr + v_off + (uint32_t)(key0 * LD);
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t kf[4];
        ldsm_x4(kf, kc_addr + np * 16 * LD + ks * 32);
        ptx::mma_16816<D>(s[np * 2 + 0], qf[ks], kf[0], kf[1]);
        ptx::mma_16816<D>(s[np * 2 + 1], qf[ks], kf[2], kf[3]);
      }
    }
    if (key0 + 32 > n_keys) {   // warp-uniform: only the last chunk holds padded keys
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (key0 + nt * 8 + 2 * t + (e & 1) >= n_keys) s[nt][e] = -INFINITY;
    }
    float cmax0 = fmaxf(fmaxf(s[0][0], s[0][1]), fmaxf(s[1][0], s[1][1]));
    float cmax1 = fmaxf(fmaxf(s[0][2], s[0][3]), fmaxf(s[1][2], s[1][3]));
    cmax0 = fmaxf(cmax0, fmaxf(fmaxf(s[2][0], s[2][1]), fmaxf(s[3][0], s[3][1])));
    cmax1 = fmaxf(cmax1, fmaxf(fmaxf(s[2][2], s[2][3]), fmaxf(s[3][2], s[3][3])));
    const float mn0 = fmaxf(m[0], quad_max(cmax0)), mn1 = fmaxf(m[1], quad_max(cmax1));   // finite: every chunk starts below n_keys
    const float corr0 = fast_exp2((m[0] - mn0) * scale_log2), corr1 = fast_exp2((m[1] - mn1) * scale_log2);
    m[0] = mn0;
    m[1] = mn1;
    const float ms0 = -mn0 * scale_log2, ms1 = -mn1 * scale_log2;
    l[0] *= corr0;
    l[1] *= corr1;
#pragma unroll
    for (int i = 0; i < NT; ++i) {
      o[i][0] *= corr0;
      o[i][1] *= corr0;
      o[i][2] *= corr1;
      o[i][3] *= corr1;
    }
    uint32_t pf[2][4];   // P as A fragments for the two 16-key k-steps of this chunk
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      const float p0 = fast_exp2(fmaf(s[nt][0], scale_log2, ms0)), p1 = fast_exp2(fmaf(s[nt][1], scale_log2, ms0));
      const float p2 = fast_exp2(fmaf(s[nt][2], scale_log2, ms1)), p3 = fast_exp2(fmaf(s[nt][3], scale_log2, ms1));
      l[0] += p0 + p1;
      l[1] += p2 + p3;
      pf[nt >> 1][(nt & 1) * 2 + 0] = D::pack2(p0, p1);
      pf[nt >> 1][(nt & 1) * 2 + 1] = D::pack2(p2, p3);
    }
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
      for (int dp = 0; dp < NT / 2; ++dp) {
        uint32_t vf[4];
        ldsm_x4_t(vf, vc_addr + kk * 16 * LD + dp * 32);
        ptx::mma_16816<D>(o[dp * 2 + 0], pf[kk], vf[0], vf[1]);
        ptx::mma_16816<D>(o[dp * 2 + 1], pf[kk], vf[2], vf[3]);
      }
    }
  }
  l[0] = quad_sum(l[0]);
  l[1] = quad_sum(l[1]);
}

// ------------------------------------------------------------------------------------------------------ temporal
// One CTA per (clip, token, head); 8 warps share the track's K and V, each warp owns 16-row query tiles.
constexpr int kTWarps = 8;

template <int HD, typename D>
__global__ void __launch_bounds__(kTWarps * 32, 2)
attn_temporal_kernel(const uint16_t* __restrict__ qkv, uint16_t* __restrict__ out, int n_frames, int n_tok, int C, int n_heads) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int LD = HD * 2 + 16;           // padded row bytes
  constexpr int CH = HD / 8;                // 16-byte chunks per row
  extern __shared__ __align__(16) uint8_t smem[];
  const int Tp = (n_frames + 31) & ~31;
  uint8_t* sq = smem;
  uint8_t* sk = sq + (size_t)Tp * LD;
  uint8_t* sv = sk + (size_t)Tp * LD;

  const int head = blockIdx.x % n_heads;
  const int tok = (blockIdx.x / n_heads) % n_tok;
  const int clip = blockIdx.x / (n_heads * n_tok);
  const size_t row_stride = (size_t)n_tok * 3 * C;   // elements between consecutive frames of this track
  const uint16_t* base = qkv + ((size_t)clip * n_frames * n_tok + tok) * 3 * C + head * HD;

  // stage Q, K, V rows [0, n_frames); zero rows [n_frames, Tp)
  const int total = 3 * Tp * CH;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int ch = i % CH;
    const int r = (i / CH) % Tp;
    const int sel = i / (CH * Tp);
    uint8_t* dst = smem + ((size_t)sel * Tp + r) * LD + ch * 16;
    if (r < n_frames) {
      ptx::cp_async16(dst, base + (size_t)r * row_stride + sel * C + ch * 8);
    } else {
      *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
  }
  ptx::cp_async_commit();
  ptx::cp_async_wait<0>();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const float scale_log2 = rsqrtf((float)HD) * kLog2e;
  uint16_t* obase = out + ((size_t)clip * n_frames * n_tok + tok) * C + head * HD;
  const size_t orow_stride = (size_t)n_tok * C;

  for (int mt = warp; mt * 16 < n_frames; mt += kTWarps) {
    float o[HD / 8][4], l[2];
    attend_tile<HD, LD, D>(smem_u32(sq) + (uint32_t)(mt * 16 * LD), smem_u32(sk), smem_u32(sv), n_frames, Tp, scale_log2, lane, o, l);
    const float inv0 = 1.0f / l[0], inv1 = 1.0f / l[1];
    // this warp's 16 query rows are dead: reuse them to transpose the output tile
    __syncwarp();
    uint8_t* orow0 = sq + (size_t)(mt * 16 + g) * LD;
    uint8_t* orow1 = orow0 + 8 * LD;
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
      *reinterpret_cast<uint32_t*>(orow0 + (i * 8 + 2 * t) * 2) = D::pack2(o[i][0] * inv0, o[i][1] * inv0);
      *reinterpret_cast<uint32_t*>(orow1 + (i * 8 + 2 * t) * 2) = D::pack2(o[i][2] * inv1, o[i][3] * inv1);
    }
    __syncwarp();
    for (int i = lane; i < 16 * CH; i += 32) {
      const int r = mt * 16 + i / CH, ch = i % CH;
      if (r < n_frames)
        *reinterpret_cast<uint4*>(obase + (size_t)r * orow_stride + ch * 8) = *reinterpret_cast<const uint4*>(sq + (size_t)r * LD + ch * 16);
    }
  }
}

// ------------------------------------------------------------------------------------------------------ temporal, tcgen05
// head_dim 64.  One CTA per (clip, token, head, 128-query tile), two CTAs per SM so the loads of one overlap the softmax of the other:
//   TMA   Q tile [128 x 64] and K [Tp x 64] (K-major), V [Tp x 64] (MN-major B operand) straight out of the [clip, frame, token, 3C]
//         activation through a 4-D tensor map (frames of one track, zero-filled past the clip end)
//   MMA   S[128 x Tp] = Q K^T into TMEM (tcgen05.mma, one thread)
//   4 softmax warps: thread = query row, reads its S row from TMEM (no shuffles), writes P (16-bit) over the dead Q / K tiles as the
//         K-major A operand of the second MMA
//   MMA   O[128 x 64] = P V into TMEM (aliasing S), read back, scaled by 1 / rowsum, staged and written with one TMA store.
constexpr int kTcThreads = 160;   // 4 softmax warps + 1 TMA / MMA warp

template <typename D>
__global__ void __launch_bounds__(kTcThreads, 2)
attn_temporal_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_o,
                        int n_frames, int n_tok, int C, int n_heads, int m_tiles) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  // [ Q 16 KB | K 32 KB | +16 KB ] = 64 KB, later P (4 k-blocks of [128 x 64] 16-bit), later the O staging tile; then V 32 KB
  uint8_t* sq = smem;
  uint8_t* sk = smem + 16384;
  uint8_t* sp = smem;
  uint8_t* sv = smem + 65536;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 98304);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int item = blockIdx.x;
  const int mt = item % m_tiles;
  item /= m_tiles;
  const int head = item % n_heads;
  item /= n_heads;
  const int tok = item % n_tok;
  const int clip = item / n_tok;
  const int Tp = (n_frames + 31) & ~31;        // keys padded to the softmax chunk (and a legal UMMA N)

  if (warp == 4 && lane == 0) {
    ptx::prefetch_tmap(&tm_q);
    ptx::prefetch_tmap(&tm_kv);
    ptx::prefetch_tmap(&tm_o);
    ptx::mbar_init(bar_qk, 1);
    ptx::mbar_init(bar_v, 1);
    ptx::mbar_init(bar_s, 1);
    ptx::mbar_init(bar_p, 128);
    ptx::mbar_init(bar_o, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_holder, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();

  if (warp == 4) {
    if (lane == 0) {
      // ---- loads
      ptx::mbar_expect_tx(bar_qk, 16384 + (uint32_t)Tp * 128);
      ptx::tma_load_4d(sq, &tm_q, bar_qk, head * 64, tok, mt * 128, clip);
      ptx::tma_load_4d(sk, &tm_kv, bar_qk, C + head * 64, tok, 0, clip);
      ptx::mbar_expect_tx(bar_v, (uint32_t)Tp * 128);
      ptx::tma_load_4d(sv, &tm_kv, bar_v, 2 * C + head * 64, tok, 0, clip);
      // ---- S = Q K^T
      ptx::mbar_wait(bar_qk, 0);
      ptx::tc_fence_after();
      {
        const uint32_t idesc = ptx::umma_idesc_16(128, Tp, D::kUmmaFmt);
        const uint64_t da = ptx::umma_desc_sw128(smem_u32(sq));
        const uint64_t db = ptx::umma_desc_sw128(smem_u32(sk));
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
        ptx::umma_commit(bar_s);
      }
      // ---- O = P V   (A = P: K-major k-blocks of 64 keys, 16 KB apart; B = V: MN-major, 16 keys = 2048 bytes per k-step)
      ptx::mbar_wait(bar_p, 0);
      ptx::mbar_wait(bar_v, 0);
      ptx::tc_fence_after();
      {
        const uint32_t idesc = ptx::umma_idesc_16_bmn(128, 64, D::kUmmaFmt);
        const uint32_t pa = smem_u32(sp), va = smem_u32(sv);
        const int ksteps = Tp / 16;
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t da = ptx::umma_desc_sw128(pa + (uint32_t)((k >> 2) * 16384 + (k & 3) * 32));
          const uint64_t db = ptx::umma_desc_mn_sw128(va + (uint32_t)(k * 2048));
          ptx::umma_f16(tmem_base, da, db, idesc, k != 0 ? 1u : 0u);
        }
        ptx::umma_commit(bar_o);
      }
    }
  } else {
    // ---- softmax: thread = query row r of this tile <-> TMEM lane r
    const int row = threadIdx.x;                       // 0..127
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t t_row = tmem_base + ((uint32_t)(32 * warp) << 16);
    const float scale_log2 = 0.125f * kLog2e;          // head_dim 64
    const int n_chunks = Tp / 32;
    ptx::mbar_wait(bar_s, 0);
    ptx::tc_fence_after();
    float mx = -INFINITY;
    for (int c = 0; c < n_chunks; ++c) {
      uint32_t r[32];
      ptx::tmem_ld32(t_row + (uint32_t)(c * 32), r);
      ptx::tmem_ld_wait();
      if ((c + 1) * 32 <= n_frames) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (c * 32 + i < n_frames) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
    }
    const float ms = -mx * scale_log2;
    float sum = 0.f;
    for (int c = 0; c < n_chunks; ++c) {
      uint32_t r[32];
      ptx::tmem_ld32(t_row + (uint32_t)(c * 32), r);
      ptx::tmem_ld_wait();
      float p[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e = fast_exp2(fmaf(__uint_as_float(r[i]), scale_log2, ms));
        p[i] = (c * 32 + i < n_frames) ? e : 0.f;
        sum += p[i];
      }
      // keys [32c, 32c+32) = half of k-block c/2: 64 bytes = chunks (c & 1) * 4 .. +4 of the 128-byte row
      uint8_t* prow = sp + (size_t)(c >> 1) * 16384 + (size_t)row * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o;
        o.x = D::pack2(p[8 * q + 0], p[8 * q + 1]);
        o.y = D::pack2(p[8 * q + 2], p[8 * q + 3]);
        o.z = D::pack2(p[8 * q + 4], p[8 * q + 5]);
        o.w = D::pack2(p[8 * q + 6], p[8 * q + 7]);
        *reinterpret_cast<uint4*>(prow + (((uint32_t)((c & 1) * 4 + q) ^ sw) << 4)) = o;
      }
    }
    ptx::tc_fence_before();            // S has been read: the second MMA may overwrite its columns
    ptx::fence_proxy_async_smem();     // P is visible to the tensor core (async proxy)
    ptx::mbar_arrive(bar_p);
    // ---- O
    ptx::mbar_wait(bar_o, 0);
    ptx::tc_fence_after();
    const float inv = 1.0f / sum;
    uint32_t r0[32], r1[32];
    ptx::tmem_ld32(t_row, r0);
    ptx::tmem_ld32(t_row + 32u, r1);
    ptx::tmem_ld_wait();
    uint8_t* orow = smem + (size_t)row * 128;          // P is dead: stage the [128 x 64] output tile over it
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t(&r)[32] = q < 4 ? r0 : r1;
      const int b = (q & 3) * 8;
      uint4 o;
      o.x = D::pack2(__uint_as_float(r[b + 0]) * inv, __uint_as_float(r[b + 1]) * inv);
      o.y = D::pack2(__uint_as_float(r[b + 2]) * inv, __uint_as_float(r[b + 3]) * inv);
      o.z = D::pack2(__uint_as_float(r[b + 4]) * inv, __uint_as_float(r[b + 5]) * inv);
      o.w = D::pack2(__uint_as_float(r[b + 6]) * inv, __uint_as_float(r[b + 7]) * inv);
      *reinterpret_cast<uint4*>(orow + (((uint32_t)q ^ sw) << 4)) = o;
    }
    ptx::fence_proxy_async_smem();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 0) {
      ptx::tma_store_4d(&tm_o, smem, head * 64, tok, mt * 128, clip);
      ptx::bulk_commit();
      ptx::bulk_wait_read<0>();      // shared memory must outlive the store's read; the global write completes on its own
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}

// ------------------------------------------------------------------------------------------------------ temporal, tcgen05, persistent
// One persistent CTA per SM walks (clip, token, head) items.  Both 128-query tiles of a head share ONE load of K and V; two groups of
// four softmax warps work on the two tiles at the same time (S0 / S1 in the two halves of TMEM), and the Q / K of the NEXT item
// are fetched by TMA while the current item is in its softmax.  Three 64 KB shared-memory regions rotate through the roles
// {Q0 Q1 K of the current item, later P0 and the O0 staging tile} -> {P1 and the O1 staging tile} -> {Q0 Q1 K of the next item}.
constexpr int kTc2Threads = 288;   // 2 x 4 softmax warps + 1 TMA / MMA warp

template <typename D>
__global__ void __launch_bounds__(kTc2Threads, 1)
attn_temporal_tc2_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_o,
                         int n_frames, int n_tok, int C, int n_heads, int n_items) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sv = smem + 3 * 65536;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 3 * 65536 + 32768);
  uint64_t* qk_full = bars + 0;    // [2] Q0 Q1 K of item n landed (parity n / 2 of barrier n & 1)
  uint64_t* v_full = bars + 2;     // V of item n landed
  uint64_t* s_full = bars + 3;     // [2] S_g complete
  uint64_t* p_full = bars + 5;     // [2] P_g written by the 128 threads of group g
  uint64_t* o_full = bars + 7;     // [2] O_g complete
  uint64_t* g_done = bars + 9;     // [2] group g has drained O_g and its output store has read shared memory
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 11);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Tp = (n_frames + 31) & ~31;
  const int m_tiles = (n_frames + 127) / 128;      // 1 or 2

  if (warp == 8 && lane == 0) {
    ptx::prefetch_tmap(&tm_q);
    ptx::prefetch_tmap(&tm_kv);
    ptx::prefetch_tmap(&tm_o);
    ptx::mbar_init(&qk_full[0], 1);
    ptx::mbar_init(&qk_full[1], 1);
    ptx::mbar_init(v_full, 1);
    for (int g = 0; g < 2; ++g) {
      ptx::mbar_init(&s_full[g], 1);
      ptx::mbar_init(&p_full[g], 128);
      ptx::mbar_init(&o_full[g], 1);
      ptx::mbar_init(&g_done[g], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_holder, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();

  auto decode = [&](int item, int& clip, int& tok, int& head) {
    head = item % n_heads;
    item /= n_heads;
    tok = item % n_tok;
    clip = item / n_tok;
  };

  if (warp == 8) {
    if (lane == 0) {
      const uint32_t qk_bytes = (uint32_t)(m_tiles * 16384 + Tp * 128);
      auto load_qk = [&](int item, uint8_t* region, uint64_t* bar) {
        int clip, tok, head;
        decode(item, clip, tok, head);
        ptx::mbar_expect_tx(bar, qk_bytes);
        for (int g = 0; g < m_tiles; ++g) ptx::tma_load_4d(region + g * 16384, &tm_q, bar, head * 64, tok, g * 128, clip);
        ptx::tma_load_4d(region + 32768, &tm_kv, bar, C + head * 64, tok, 0, clip);
      };
      auto load_v = [&](int item) {
        int clip, tok, head;
        decode(item, clip, tok, head);
        ptx::mbar_expect_tx(v_full, (uint32_t)Tp * 128);
        ptx::tma_load_4d(sv, &tm_kv, v_full, 2 * C + head * 64, tok, 0, clip);
      };
      const uint32_t idesc_s = ptx::umma_idesc_16(128, Tp, D::kUmmaFmt);
      const uint32_t idesc_o = ptx::umma_idesc_16_bmn(128, 64, D::kUmmaFmt);
      int cur = 0, p1 = 1, nxt = 2;
      uint32_t n = 0;
      int item = blockIdx.x;
      if (item < n_items) {
        load_qk(item, smem + cur * 65536, &qk_full[0]);
        load_v(item);
      }
      for (; item < n_items; item += gridDim.x, ++n) {
        const uint32_t par = n & 1, qpar = (n >> 1) & 1;
        uint8_t* rc = smem + cur * 65536;
        uint8_t* rp = smem + p1 * 65536;
        if (n > 0) {
          // both groups have drained item n-1: TMEM, the V buffer and the regions of P0 / P1 are free again
          for (int g = 0; g < m_tiles; ++g) ptx::mbar_wait(&g_done[g], par ^ 1);
          load_v(item);
        }
        ptx::mbar_wait(&qk_full[par], qpar);
        ptx::tc_fence_after();
        {
          const uint64_t db = ptx::umma_desc_sw128(smem_u32(rc + 32768));
          for (int g = 0; g < m_tiles; ++g) {
            const uint64_t da = ptx::umma_desc_sw128(smem_u32(rc + g * 16384));
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_f16(tmem_base + (uint32_t)(g * 256), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_s, k != 0 ? 1u : 0u);
            ptx::umma_commit(&s_full[g]);
          }
        }
        if (item + (int)gridDim.x < n_items) load_qk(item + gridDim.x, smem + nxt * 65536, &qk_full[par ^ 1]);
        ptx::mbar_wait(v_full, par);
        for (int g = 0; g < m_tiles; ++g) {
          ptx::mbar_wait(&p_full[g], par);
          ptx::tc_fence_after();
          const uint32_t pa = smem_u32(g == 0 ? rc : rp), va = smem_u32(sv);
          const int ksteps = Tp / 16;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t da = ptx::umma_desc_sw128(pa + (uint32_t)((k >> 2) * 16384 + (k & 3) * 32));
            const uint64_t db = ptx::umma_desc_mn_sw128(va + (uint32_t)(k * 2048));
            ptx::umma_f16(tmem_base + (uint32_t)(g * 256), da, db, idesc_o, k != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&o_full[g]);
        }
        const int t = cur;
        cur = nxt;
        nxt = p1;
        p1 = t;
      }
    }
  } else {
    const int g = warp >> 2;                            // softmax group = query tile
    if (g < m_tiles) {
      const int row = threadIdx.x & 127;                // tile-local query row <-> TMEM lane
      const uint32_t sw = (uint32_t)(row & 7);
      const uint32_t t_row = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(g * 256);
      const float scale_log2 = 0.125f * kLog2e;
      const int n_chunks = Tp / 32;
      int cur = 0, p1 = 1, nxt = 2;
      uint32_t n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        const uint32_t par = n & 1;
        uint8_t* region = smem + (g == 0 ? cur : p1) * 65536;   // P_g, then the O_g staging tile
        // S_g is complete; P0 overwrites Q and K, so group 0 also needs S1 (the other reader of K) to be complete
        ptx::mbar_wait(&s_full[g], par);
        if (g == 0 && m_tiles == 2) ptx::mbar_wait(&s_full[1], par);
        ptx::tc_fence_after();
        // ---- softmax in ONE pass over the S row (reading TMEM is the limiter: 64 B/clk per SM).  The shift is the maximum of the
        // first 32 keys, a lower bound of the row maximum: P = 2^(scale (s - shift)) is then >= 1 at the true maximum, and the
        // normalisation by the row sum removes the shift again.  If some exponent leaves the range that is safe for the 16-bit P
        // (kSafeExp), the warp redoes its rows with their true maxima (rare).
        constexpr float kSafeExp = D::kUmmaFmt == 1 ? 60.0f : 13.0f;   // bf16 has the fp32 exponent range, fp16 tops out at 2^16
        float sum = 0.f, emax = -INFINITY, ms;
        auto emit = [&](const uint32_t(&r)[32], int c, float shift, float& e_hi, float& acc) {
          float p[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = fmaf(__uint_as_float(r[i]), scale_log2, shift);
            const bool live = (c + 1) * 32 <= n_frames || c * 32 + i < n_frames;
            e_hi = fmaxf(e_hi, live ? e : -INFINITY);
            p[i] = live ? fast_exp2(e) : 0.f;
            acc += p[i];
          }
          uint8_t* prow = region + (size_t)(c >> 1) * 16384 + (size_t)row * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            uint4 o;
            o.x = D::pack2(p[8 * q + 0], p[8 * q + 1]);
            o.y = D::pack2(p[8 * q + 2], p[8 * q + 3]);
            o.z = D::pack2(p[8 * q + 4], p[8 * q + 5]);
            o.w = D::pack2(p[8 * q + 6], p[8 * q + 7]);
            *reinterpret_cast<uint4*>(prow + (((uint32_t)((c & 1) * 4 + q) ^ sw) << 4)) = o;
          }
        };
        {
          uint32_t r[32];
          ptx::tmem_ld32(t_row, r);
          ptx::tmem_ld_wait();
          float m0 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < n_frames) m0 = fmaxf(m0, __uint_as_float(r[i]));
          ms = -m0 * scale_log2;
          emit(r, 0, ms, emax, sum);
          for (int c = 1; c < n_chunks; ++c) {
            ptx::tmem_ld32(t_row + (uint32_t)(c * 32), r);
            ptx::tmem_ld_wait();
            emit(r, c, ms, emax, sum);
          }
        }
        if (__any_sync(0xffffffffu, emax > kSafeExp)) {   // tcgen05.ld is warp-collective: the whole warp redoes its 32 rows
          ms -= emax;                     // shift by the true row maximum: every exponent <= 0
          sum = 0.f;
          float unused = -INFINITY;
          uint32_t r[32];
          for (int c = 0; c < n_chunks; ++c) {
            ptx::tmem_ld32(t_row + (uint32_t)(c * 32), r);
            ptx::tmem_ld_wait();
            emit(r, c, ms, unused, sum);
          }
        }
        ptx::tc_fence_before();
        ptx::fence_proxy_async_smem();
        ptx::mbar_arrive(&p_full[g]);
        // ---- O_g
        ptx::mbar_wait(&o_full[g], par);
        ptx::tc_fence_after();
        const float inv = 1.0f / sum;
        uint32_t r0[32], r1[32];
        ptx::tmem_ld32(t_row, r0);
        ptx::tmem_ld32(t_row + 32u, r1);
        ptx::tmem_ld_wait();
        uint8_t* orow = region + (size_t)row * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint32_t(&r)[32] = q < 4 ? r0 : r1;
          const int b = (q & 3) * 8;
          uint4 o;
          o.x = D::pack2(__uint_as_float(r[b + 0]) * inv, __uint_as_float(r[b + 1]) * inv);
          o.y = D::pack2(__uint_as_float(r[b + 2]) * inv, __uint_as_float(r[b + 3]) * inv);
          o.z = D::pack2(__uint_as_float(r[b + 4]) * inv, __uint_as_float(r[b + 5]) * inv);
          o.w = D::pack2(__uint_as_float(r[b + 6]) * inv, __uint_as_float(r[b + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + (((uint32_t)q ^ sw) << 4)) = o;
        }
        ptx::tc_fence_before();
        ptx::fence_proxy_async_smem();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
        if (row == 0) {
          int clip, tok, head;
          decode(item, clip, tok, head);
          ptx::tma_store_4d(&tm_o, region, head * 64, tok, g * 128, clip);
          ptx::bulk_commit();
          ptx::bulk_wait_read<0>();
          ptx::mbar_arrive(&g_done[g]);
        }
        const int t = cur;
        cur = nxt;
        nxt = p1;
        p1 = t;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------ temporal, tcgen05, P in TMEM
// The T = 243 kernel, third version.  attn_temporal_tc2_kernel above serialises the phases of an item (S = Q K^T, softmax, O = P V,
// drain) because P is staged through 128 KB of shared memory per item, which leaves room for ONE item's operands and makes the
// two 128-query tiles run in lockstep: the tensor pipe waits for the softmax, the softmax warps wait for the MMAs and the loads.
// Here the softmax writes P (16-bit, two values per column) back into TENSOR MEMORY over the S it has just read
// (tcgen05.st) and O = P V takes its A operand from there (tcgen05.mma with a TMEM A operand, SASS UTCHMMA ... tmem[A]):
//   * no shared memory for P: the operands of TWO items are resident (2 x {Q0, Q1, K, V} = 192 KB), loaded by their own warp, so an
//     item's loads never sit on the critical path;
//   * the two query tiles are DE-PHASED: the MMA thread issues S0(n), PV1(n-1), S1(n), PV0(n), ... so the softmax of one tile runs
//     while the other tile is in its MMAs / drain, and the tensor pipe, the MUFU and the TMEM read port are shared instead of
//     alternating between busy and idle.
// TMEM: region g (256 columns) = S_g [128 x Tp] fp32 -> P_g in columns [0, Tp/2) -> O_g [128 x 64] fp32 in columns [128, 192).
// The softmax keeps the packed P row in registers until the whole S row has been read (the guarded shift estimate may ask for a
// second pass over S, so S must stay intact until then), then stores it with Tp/32 tcgen05.st.x16.
// Registers: the register file is split over the four SM sub-partitions (16 K each), so a CTA of 9 - 12 warps can give every thread
// at most 168 registers at launch.  The softmax threads hold a 128-register packed P row, so the kernel is launched with three
// warpgroups (warps 10 and 11 only take part in the hand-over) and redistributes with setmaxnreg: the MMA / TMA warpgroup shrinks
// to 40 registers, the two softmax warpgroups grow to 232 (8 x 32 x 232 + 4 x 32 x 40 = 12 x 32 x 168).
constexpr int kTc3Threads = 384;   // 2 x 4 softmax warps + {MMA warp, TMA warp, 2 idle warps}
constexpr int kTc3Opnd = 98304;    // Q0 16 KB | Q1 16 KB | K 32 KB | V 32 KB
constexpr int kTc3Smem = 2 * kTc3Opnd + 2 * 16384 + 256;

template <typename D>
__global__ void __launch_bounds__(kTc3Threads, 1)
attn_temporal_tc3_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv, const __grid_constant__ CUtensorMap tm_o,
                         int n_frames, int n_tok, int C, int n_heads, int n_items) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* ostage = smem + 2 * kTc3Opnd;                                   // [2] 16 KB output staging tile of group g
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kTc3Opnd + 2 * 16384);
  uint64_t* qk_full = bars + 0;      // [2] Q0 Q1 K of the item in operand buffer b landed
  uint64_t* v_full = bars + 2;       // [2] V landed
  uint64_t* qk_free = bars + 4;      // [2] both S MMAs of the item have read Q / K (tcgen05.commit)
  uint64_t* v_free = bars + 6;       // [2] both PV MMAs of the item have read V (tcgen05.commit)
  uint64_t* s_full = bars + 8;       // [2] S_g complete
  uint64_t* p_full = bars + 10;      // [2] P_g stored to TMEM by the 128 threads of group g
  uint64_t* o_full = bars + 12;      // [2] O_g complete
  uint64_t* r_free = bars + 14;      // [2] group g has read O_g out of TMEM: region g may take the next S_g
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Tp = (n_frames + 31) & ~31;            // 160 .. 256 (the host routes n_frames <= 128 to the block-diagonal kernel)
  const int n_chunks = Tp / 32;

  if (warp == 8 && lane == 0) {
    ptx::prefetch_tmap(&tm_q);
    ptx::prefetch_tmap(&tm_kv);
    ptx::prefetch_tmap(&tm_o);
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&qk_full[b], 1);
      ptx::mbar_init(&v_full[b], 1);
      ptx::mbar_init(&qk_free[b], 1);
      ptx::mbar_init(&v_free[b], 1);
      ptx::mbar_init(&s_full[b], 1);
      ptx::mbar_init(&p_full[b], 128);
      ptx::mbar_init(&o_full[b], 1);
      ptx::mbar_init(&r_free[b], 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_holder, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();

  auto decode = [&](int item, int& clip, int& tok, int& head) {
    head = item % n_heads;
    item /= n_heads;
    tok = item % n_tok;
    clip = item / n_tok;
  };

  if (warp >= 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if (warp == 9) {
    if (lane == 0) {
      // ===================== operand loader: item n goes to buffer n & 1 =====================
      uint32_t n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        const uint32_t b = n & 1, use = n >> 1;
        uint8_t* buf = smem + b * kTc3Opnd;
        int clip, tok, head;
        decode(item, clip, tok, head);
        if (use > 0) ptx::mbar_wait(&qk_free[b], (use - 1) & 1);
        ptx::mbar_expect_tx(&qk_full[b], (uint32_t)(2 * 16384 + Tp * 128));
        ptx::tma_load_4d(buf, &tm_q, &qk_full[b], head * 64, tok, 0, clip);
        ptx::tma_load_4d(buf + 16384, &tm_q, &qk_full[b], head * 64, tok, 128, clip);
        ptx::tma_load_4d(buf + 32768, &tm_kv, &qk_full[b], C + head * 64, tok, 0, clip);
        if (use > 0) ptx::mbar_wait(&v_free[b], (use - 1) & 1);
        ptx::mbar_expect_tx(&v_full[b], (uint32_t)Tp * 128);
        ptx::tma_load_4d(buf + 65536, &tm_kv, &v_full[b], 2 * C + head * 64, tok, 0, clip);
      }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      // ===================== MMA issuer: S0(n), PV1(n-1), S1(n), PV0(n) =====================
      const uint32_t idesc_s = ptx::umma_idesc_16(128, Tp, D::kUmmaFmt);
      const uint32_t idesc_o = ptx::umma_idesc_16_bmn(128, 64, D::kUmmaFmt);
      const int ksteps = Tp / 16;
      auto issue_s = [&](int g, uint8_t* buf) {
        const uint64_t da = ptx::umma_desc_sw128(smem_u32(buf + g * 16384));
        const uint64_t db = ptx::umma_desc_sw128(smem_u32(buf + 32768));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::umma_f16(tmem_base + (uint32_t)(g * 256), da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_s, k != 0 ? 1u : 0u);
        ptx::umma_commit(&s_full[g]);
      };
      auto issue_pv = [&](int g, uint8_t* buf) {
        const uint32_t va = smem_u32(buf + 65536);
        const uint32_t region = tmem_base + (uint32_t)(g * 256);
        for (int k = 0; k < ksteps; ++k)
          ptx::umma_f16_ts(region + 128u, region + (uint32_t)(8 * k), ptx::umma_desc_mn_sw128(va + (uint32_t)(k * 2048)), idesc_o, k != 0 ? 1u : 0u);
        ptx::umma_commit(&o_full[g]);
      };
      uint32_t n = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
        const uint32_t b = n & 1, use = n >> 1, par = n & 1;
        uint8_t* buf = smem + b * kTc3Opnd;
        uint8_t* prev = smem + (b ^ 1) * kTc3Opnd;
        ptx::mbar_wait(&qk_full[b], use & 1);
        if (n > 0) ptx::mbar_wait(&r_free[0], par ^ 1);        // O0(n-1) has been read: region 0 is free
        ptx::tc_fence_after();
        issue_s(0, buf);
        if (n > 0) {
          ptx::mbar_wait(&p_full[1], par ^ 1);                 // P1(n-1) is in TMEM (V(n-1) landed long ago: PV0(n-1) read it)
          ptx::tc_fence_after();
          issue_pv(1, prev);
          ptx::umma_commit(&v_free[b ^ 1]);                    // last reader of V(n-1)
          ptx::mbar_wait(&r_free[1], par ^ 1);                 // O1(n-1) has been read: region 1 is free
          ptx::tc_fence_after();
        }
        issue_s(1, buf);
        ptx::umma_commit(&qk_free[b]);                         // last reader of Q / K(n)
        ptx::mbar_wait(&p_full[0], par);
        ptx::mbar_wait(&v_full[b], use & 1);
        ptx::tc_fence_after();
        issue_pv(0, buf);
      }
      if (n > 0) {
        const uint32_t last = n - 1;
        ptx::mbar_wait(&p_full[1], last & 1);
        ptx::tc_fence_after();
        issue_pv(1, smem + (last & 1) * kTc3Opnd);
      }
    }
  } else if (warp < 8) {
    // ===================== softmax / drain: group g = query tile g, thread = one query row =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int g = warp >> 2;
    const int row = threadIdx.x & 127;                  // tile-local query row <-> TMEM lane
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t t_row = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(g * 256);
    const float scale_log2 = 0.125f * kLog2e;
    // largest row sum (hence largest P value) accepted from the estimated shift: bf16 has the fp32 exponent range, fp16 tops out at 65504
    constexpr float kSafeSum = D::kUmmaFmt == 1 ? 1.1529215e18f /* 2^60 */ : 8192.0f;
    uint8_t* stage = ostage + g * 16384;
    uint8_t* orow = stage + (size_t)row * 128;
    uint32_t n = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n) {
      const uint32_t par = n & 1;
      ptx::mbar_wait(&s_full[g], par);
      ptx::tc_fence_after();
      // ---- softmax in ONE pass over the S row; the shift is the maximum of the first 32 keys (a lower bound of the row maximum, the
      // normalisation by the row sum removes it again).  If some exponent leaves the range that is safe for the 16-bit P, the warp
      // redoes its rows with their true maxima (tcgen05.ld is warp-collective).  P stays in registers until S is no longer needed.
      uint32_t pk[128];
      float sum = 0.f, ms;
      auto emit = [&](const uint32_t(&r)[32], int c, float shift, float& acc) {
        float p[32];
        float a4[4] = {0.f, 0.f, 0.f, 0.f};   // four chains instead of one
        if ((c + 1) * 32 <= n_frames) {   // warp-uniform: every key of the chunk exists (all chunks but the last): no masking
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            p[i] = fast_exp2(fmaf(__uint_as_float(r[i]), scale_log2, shift));
            a4[i & 3] += p[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float e = fast_exp2(fmaf(__uint_as_float(r[i]), scale_log2, shift));
            p[i] = c * 32 + i < n_frames ? e : 0.f;
            a4[i & 3] += p[i];
          }
        }
        acc += (a4[0] + a4[1]) + (a4[2] + a4[3]);
#pragma unroll
        for (int i = 0; i < 16; ++i) pk[16 * c + i] = D::pack2(p[2 * i], p[2 * i + 1]);
      };
      {
        // two register buffers: the TMEM load of chunk c + 1 is in flight while chunk c is processed
        uint32_t ra[32], rb[32];
        ptx::tmem_ld32(t_row, ra);
        ptx::tmem_ld_wait();
        ptx::tmem_ld32(t_row + 32u, rb);                   // n_frames > 128: at least five chunks
        float m0 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; ++i) m0 = fmaxf(m0, __uint_as_float(ra[i]));     // the first 32 keys all exist
        ms = -m0 * scale_log2;
        emit(ra, 0, ms, sum);
#pragma unroll
        for (int c = 1; c < 8; ++c) {
          if (c < n_chunks) {
            ptx::tmem_ld_wait();
            if (c & 1) {
              if (c + 1 < n_chunks) ptx::tmem_ld32(t_row + (uint32_t)((c + 1) * 32), ra);
              emit(rb, c, ms, sum);
            } else {
              if (c + 1 < n_chunks) ptx::tmem_ld32(t_row + (uint32_t)((c + 1) * 32), rb);
              emit(ra, c, ms, sum);
            }
          }
        }
      }
      // Every P value is at most the row sum: a sum within the safe range proves that no exponent left it (and a NaN / inf sum fails
      // the comparison), without tracking the largest exponent element by element.
      if (__any_sync(0xffffffffu, !(sum <= kSafeSum))) {
        uint32_t r[32];
        float m = -INFINITY;              // true row maximum over the keys that exist
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c < n_chunks) {
            ptx::tmem_ld32(t_row + (uint32_t)(c * 32), r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i < n_frames) m = fmaxf(m, __uint_as_float(r[i]));
          }
        }
        ms = -m * scale_log2;             // every exponent <= 0
        sum = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          if (c < n_chunks) {
            ptx::tmem_ld32(t_row + (uint32_t)(c * 32), r);
            ptx::tmem_ld_wait();
            emit(r, c, ms, sum);
          }
        }
      }
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        if (c < n_chunks) ptx::tmem_st16(t_row + (uint32_t)(c * 16), &pk[16 * c]);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_full[g]);
      // ---- O_g = P_g V: read it out of TMEM, hand the region back, normalise, stage, store
      ptx::mbar_wait(&o_full[g], par);
      ptx::tc_fence_after();
      uint32_t r0[32], r1[32];
      ptx::tmem_ld32(t_row + 128u, r0);
      ptx::tmem_ld32(t_row + 160u, r1);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&r_free[g]);
      const float inv = 1.0f / sum;
      if (n > 0) {                        // the previous store of this group must have read the staging tile
        if (row == 0) ptx::bulk_wait_read<0>();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t(&r)[32] = q < 4 ? r0 : r1;
        const int bq = (q & 3) * 8;
        uint4 o;
        o.x = D::pack2(__uint_as_float(r[bq + 0]) * inv, __uint_as_float(r[bq + 1]) * inv);
        o.y = D::pack2(__uint_as_float(r[bq + 2]) * inv, __uint_as_float(r[bq + 3]) * inv);
        o.z = D::pack2(__uint_as_float(r[bq + 4]) * inv, __uint_as_float(r[bq + 5]) * inv);
        o.w = D::pack2(__uint_as_float(r[bq + 6]) * inv, __uint_as_float(r[bq + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + (((uint32_t)q ^ sw) << 4)) = o;
      }
      ptx::fence_proxy_async_smem();
      asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
      if (row == 0) {
        int clip, tok, head;
        decode(item, clip, tok, head);
        ptx::tma_store_4d(&tm_o, stage, head * 64, tok, g * 128, clip);
        ptx::bulk_commit();
      }
    }
    if (row == 0) ptx::bulk_wait<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------ spatial, tcgen05
// head_dim 64, n_tok <= 32.  G = (128 / n_tok) * n_tok consecutive tokens (7 frames of 17 joints) form one 128-row tile: the same
// rows are the queries and the keys, S = Q K^T is a [128 x 128] tcgen05 MMA of which only the block diagonal (same frame) is
// kept.  The softmax thread of row i reads just the 32-column chunks that contain its frame's columns, writes the (mostly zero)
// 16-bit P row over the dead Q / K tiles, and O = P V uses V as the MN-major operand.  48 KB of shared memory and 128 TMEM
// columns per CTA: four CTAs per SM overlap loads, softmax and stores.  Rows [G, 128) of a tile belong to the next tile and are
// not stored.  Tiles are laid out per CLIP (3-D tensor maps clip the last tile of a clip), so a clip's result does not depend on
// which other clips share its micro-batch.
//
// kTracks = true: the same kernel as the TEMPORAL attention of short clips (n_frames <= 128, e.g. the T = 27 / 81 configurations).  A tile
// holds P = 128 / n_frames whole (clip, token) tracks, fetched by ONE 4-D TMA box {64 columns, P tokens, n_frames frames, 1 clip}: tile row
// r = frame * P + (token - first token), so the sequences are INTERLEAVED with period P and the softmax window of row r is the columns
// j = r (mod P), j < P * n_frames.  (n_rows = n_frames, G = P here.)  The persistent two-tile kernel below is built for T = 243 and spends
// the same fixed cost per (track, head) on 27 frames.
template <typename D, bool kTracks>
__global__ void __launch_bounds__(kTcThreads, 4)
attn_spatial_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_o, int n_rows, int n_tok, int C,
                       int n_heads, int G, int tiles_per_clip) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sq = smem;                 // [128 x 64] queries; later P k-block 0
  uint8_t* sk = smem + 16384;         // [128 x 64] keys;    later P k-block 1
  uint8_t* sv = smem + 32768;         // [128 x 64] values (MN-major B operand)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 49152);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_v = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x % n_heads;
  const int tile = (blockIdx.x / n_heads) % tiles_per_clip;
  const int clip = blockIdx.x / (n_heads * tiles_per_clip);
  const int row0 = tile * G;          // clip-relative, multiple of n_tok: frame f of the tile owns tile-local columns [f * n_tok, (f + 1) * n_tok)
                                      // (kTracks: first token of the tile)
  const int rows_valid = kTracks ? G * n_rows : 128;   // rows a track box fills; the rest of the 128-row tile is not touched by TMA

  if (warp == 4 && lane == 0) {
    ptx::prefetch_tmap(&tm_in);
    ptx::prefetch_tmap(&tm_o);
    ptx::mbar_init(bar_qk, 1);
    ptx::mbar_init(bar_v, 1);
    ptx::mbar_init(bar_s, 1);
    ptx::mbar_init(bar_p, 128);
    ptx::mbar_init(bar_o, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_holder, 128);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();

  if (warp == 4) {
    if (lane == 0) {
      if constexpr (kTracks) {
        const uint32_t box_bytes = (uint32_t)rows_valid * 128u;
        ptx::mbar_expect_tx(bar_qk, 2 * box_bytes);
        ptx::tma_load_4d(sq, &tm_in, bar_qk, head * 64, row0, 0, clip);
        ptx::tma_load_4d(sk, &tm_in, bar_qk, C + head * 64, row0, 0, clip);
        ptx::mbar_expect_tx(bar_v, box_bytes);
        ptx::tma_load_4d(sv, &tm_in, bar_v, 2 * C + head * 64, row0, 0, clip);
      } else {
        ptx::mbar_expect_tx(bar_qk, 32768);
        ptx::tma_load_3d(sq, &tm_in, bar_qk, head * 64, row0, clip);
        ptx::tma_load_3d(sk, &tm_in, bar_qk, C + head * 64, row0, clip);
        ptx::mbar_expect_tx(bar_v, 16384);
        ptx::tma_load_3d(sv, &tm_in, bar_v, 2 * C + head * 64, row0, clip);
      }
      ptx::mbar_wait(bar_qk, 0);
      ptx::tc_fence_after();
      {
        constexpr uint32_t idesc = ptx::umma_idesc_16(128, 128, D::kUmmaFmt);
        const uint64_t da = ptx::umma_desc_sw128(smem_u32(sq));
        const uint64_t db = ptx::umma_desc_sw128(smem_u32(sk));
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, k != 0 ? 1u : 0u);
        ptx::umma_commit(bar_s);
      }
      ptx::mbar_wait(bar_p, 0);
      ptx::mbar_wait(bar_v, 0);
      ptx::tc_fence_after();
      {
        constexpr uint32_t idesc = ptx::umma_idesc_16_bmn(128, 64, D::kUmmaFmt);
        const uint32_t pa = smem_u32(smem), va = smem_u32(sv);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint64_t da = ptx::umma_desc_sw128(pa + (uint32_t)((k >> 2) * 16384 + (k & 3) * 32));
          const uint64_t db = ptx::umma_desc_mn_sw128(va + (uint32_t)(k * 2048));
          ptx::umma_f16(tmem_base, da, db, idesc, k != 0 ? 1u : 0u);
        }
        ptx::umma_commit(bar_o);
      }
    }
  } else {
    const int row = threadIdx.x;                       // tile-local query row <-> TMEM lane
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t t_row = tmem_base + ((uint32_t)(32 * warp) << 16);
    const float scale_log2 = 0.125f * kLog2e;
    // columns of this row's frame, clipped to the keys that exist; rows past the end of the tensor get an empty window
    int a = (row / n_tok) * n_tok;
    int b = a + n_tok;
    if (b > 128) b = 128;
    if (row0 + b > n_rows) b = n_rows - row0;       // n_rows = rows of one clip
    bool live = row0 + row < n_rows && b > a;
    // warp-uniform range of 32-column chunks that covers the windows of rows [32 warp, 32 warp + 32)
    int c_lo = ((32 * warp) / n_tok * n_tok) / 32;
    int c_hi = (((32 * warp + 31) / n_tok + 1) * n_tok + 31) / 32;
    if (c_hi > 4) c_hi = 4;
    uint32_t win[4] = {0u, 0u, 0u, 0u};            // bit i of win[c]: column 32 c + i belongs to this row's softmax window
    if constexpr (kTracks) {
      const int trk = row % G;                     // G = tracks per tile = interleave period
      live = row < rows_valid && row0 + trk < n_tok;
      c_lo = 0;
      c_hi = (rows_valid + 31) / 32;
      uint32_t pat = 0u;                           // ones at the multiples of the period
      for (int i = 0; i < 32; i += G) pat |= 1u << i;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int start = ((trk - 32 * c) % G + G) % G;            // first i with (32 c + i) = trk (mod G)
        const int left = rows_valid - 32 * c;                        // columns of this chunk that exist
        const uint32_t range = left >= 32 ? 0xffffffffu : (left > 0 ? (1u << left) - 1u : 0u);
        win[c] = (pat << start) & range;
      }
      // rows the track boxes do not fill hold stale shared memory: V must be zero there (0 * NaN), K / Q garbage only reaches masked entries
      for (int r = rows_valid + row; r < 128; r += 128) {
#pragma unroll
        for (int q = 0; q < 8; ++q) *reinterpret_cast<uint4*>(sv + (size_t)r * 128 + q * 16) = make_uint4(0u, 0u, 0u, 0u);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int lo = a - 32 * c, hi = b - 32 * c;                  // window [lo, hi) in chunk coordinates
        const uint32_t upto_hi = hi >= 32 ? 0xffffffffu : (hi > 0 ? (1u << hi) - 1u : 0u);
        const uint32_t upto_lo = lo >= 32 ? 0xffffffffu : (lo > 0 ? (1u << lo) - 1u : 0u);
        win[c] = upto_hi & ~upto_lo;
      }
    }
    ptx::mbar_wait(bar_s, 0);
    ptx::tc_fence_after();
    float mx = -INFINITY;
    for (int c = c_lo; c < c_hi; ++c) {
      uint32_t r[32];
      ptx::tmem_ld32(t_row + (uint32_t)(c * 32), r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        if ((win[c] >> i) & 1u) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
    }
    const float ms = live ? -mx * scale_log2 : 0.f;
    float sum = 0.f;
    for (int c = 0; c < 4; ++c) {
      float p[32];
      if (c >= c_lo && c < c_hi) {
        uint32_t r[32];
        ptx::tmem_ld32(t_row + (uint32_t)(c * 32), r);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const bool in = live && ((win[c] >> i) & 1u);
          const float e = fast_exp2(fmaf(in ? __uint_as_float(r[i]) : 0.f, scale_log2, ms));
          p[i] = in ? e : 0.f;
          sum += p[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) p[i] = 0.f;
      }
      uint8_t* prow = smem + (size_t)(c >> 1) * 16384 + (size_t)row * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o;
        o.x = D::pack2(p[8 * q + 0], p[8 * q + 1]);
        o.y = D::pack2(p[8 * q + 2], p[8 * q + 3]);
        o.z = D::pack2(p[8 * q + 4], p[8 * q + 5]);
        o.w = D::pack2(p[8 * q + 6], p[8 * q + 7]);
        *reinterpret_cast<uint4*>(prow + (((uint32_t)((c & 1) * 4 + q) ^ sw) << 4)) = o;
      }
    }
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();
    ptx::mbar_arrive(bar_p);
    ptx::mbar_wait(bar_o, 0);
    ptx::tc_fence_after();
    const float inv = live ? 1.0f / sum : 0.f;
    uint32_t r0[32], r1[32];
    ptx::tmem_ld32(t_row, r0);
    ptx::tmem_ld32(t_row + 32u, r1);
    ptx::tmem_ld_wait();
    uint8_t* orow = smem + (size_t)row * 128;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t(&r)[32] = q < 4 ? r0 : r1;
      const int bb = (q & 3) * 8;
      uint4 o;
      o.x = D::pack2(__uint_as_float(r[bb + 0]) * inv, __uint_as_float(r[bb + 1]) * inv);
      o.y = D::pack2(__uint_as_float(r[bb + 2]) * inv, __uint_as_float(r[bb + 3]) * inv);
      o.z = D::pack2(__uint_as_float(r[bb + 4]) * inv, __uint_as_float(r[bb + 5]) * inv);
      o.w = D::pack2(__uint_as_float(r[bb + 6]) * inv, __uint_as_float(r[bb + 7]) * inv);
      *reinterpret_cast<uint4*>(orow + (((uint32_t)q ^ sw) << 4)) = o;
    }
    ptx::fence_proxy_async_smem();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 0) {
      if constexpr (kTracks)
        ptx::tma_store_4d(&tm_o, smem, head * 64, row0, 0, clip);   // box {64, P tokens, n_frames, 1}: tokens past the clip's last are dropped
      else
        ptx::tma_store_3d(&tm_o, smem, head * 64, row0, clip);   // box of G rows: rows [G, 128) belong to the next tile
      ptx::bulk_commit();
      ptx::bulk_wait_read<0>();
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 128);
  }
}

// ------------------------------------------------------------------------------------------------------ spatial
// Work item = one (frame, head): n_tok rows of q, k and v, 2 * HD bytes each.  Every WARP streams its own items through a
// private double buffer (cp.async for item i+1 in flight while item i is computed), so there is no block-level barrier and
// the 8 warps of a CTA keep ~50 KB of loads in flight per SM.  The output tile is transposed through the dead q rows.
constexpr int kSWarps = 8;

template <int HD, typename D>
__global__ void __launch_bounds__(kSWarps * 32)
attn_spatial_kernel(const uint16_t* __restrict__ qkv, uint16_t* __restrict__ out, int64_t n_seq, int n_tok, int C, int n_heads) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int LD = HD * 2 + 16;           // padded row bytes
  constexpr int CH = HD / 8;                // 16-byte chunks per row
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  constexpr int kRows = 32;                  // rows per matrix; rows [n_tok, 32) stay zero (loads only touch rows < n_tok)
  constexpr int buf_bytes = 3 * kRows * LD;
  uint8_t* wbase = smem + (size_t)warp * 2 * buf_bytes;
  for (int i = lane; i < 2 * buf_bytes / 16; i += 32) *reinterpret_cast<uint4*>(wbase + i * 16) = make_uint4(0, 0, 0, 0);
  __syncwarp();

  const float scale_log2 = rsqrtf((float)HD) * kLog2e;
  const int64_t n_items = n_seq * n_heads;
  const int64_t stride = (int64_t)gridDim.x * kSWarps;
  const int chunks = 3 * n_tok * CH;
  auto issue = [&](int64_t item, uint8_t* buf) {
    const int64_t seq = item / n_heads;
    const int head = (int)(item - seq * n_heads);
    const uint16_t* src = qkv + (size_t)seq * n_tok * 3 * C + head * HD;
    for (int i = lane; i < chunks; i += 32) {
      const int ch = i % CH, r = (i / CH) % n_tok, sel = i / (CH * n_tok);
      ptx::cp_async16(buf + (size_t)(sel * kRows + r) * LD + ch * 16, src + (size_t)r * 3 * C + sel * C + ch * 8);
    }
    ptx::cp_async_commit();
  };

  int64_t item = (int64_t)blockIdx.x * kSWarps + warp;
  int cur = 0;
  if (item < n_items) issue(item, wbase);
  for (; item < n_items; item += stride, cur ^= 1) {
    uint8_t* buf = wbase + cur * buf_bytes;
    if (item + stride < n_items) {
      issue(item + stride, wbase + (cur ^ 1) * buf_bytes);
      ptx::cp_async_wait<1>();
    } else {
      ptx::cp_async_wait<0>();
    }
    __syncwarp();
    const uint32_t qb = smem_u32(buf), kb = qb + kRows * LD, vb = kb + kRows * LD;
    for (int mt = 0; mt * 16 < n_tok; ++mt) {
      float o[HD / 8][4], l[2];
      attend_tile<HD, LD, D>(qb + (uint32_t)(mt * 16 * LD), kb, vb, n_tok, kRows, scale_log2, lane, o, l);
      const float inv0 = 1.0f / l[0], inv1 = 1.0f / l[1];
      __syncwarp();                          // all lanes hold their q fragments of this tile: its rows may be overwritten
      const int r0 = mt * 16 + g, r1 = r0 + 8;
#pragma unroll
      for (int i = 0; i < HD / 8; ++i) {
        if (r0 < n_tok) *reinterpret_cast<uint32_t*>(buf + (size_t)r0 * LD + (i * 8 + 2 * t) * 2) = D::pack2(o[i][0] * inv0, o[i][1] * inv0);
        if (r1 < n_tok) *reinterpret_cast<uint32_t*>(buf + (size_t)r1 * LD + (i * 8 + 2 * t) * 2) = D::pack2(o[i][2] * inv1, o[i][3] * inv1);
      }
    }
    __syncwarp();
    const int64_t seq = item / n_heads;
    const int head = (int)(item - seq * n_heads);
    uint16_t* dst = out + (size_t)seq * n_tok * C + head * HD;
    for (int i = lane; i < n_tok * CH; i += 32) {
      const int r = i / CH, ch = i - r * CH;
      *reinterpret_cast<uint4*>(dst + (size_t)r * C + ch * 8) = *reinterpret_cast<const uint4*>(buf + (size_t)r * LD + ch * 16);
    }
    __syncwarp();                            // the buffer is refilled by the load issued in the next iteration
  }
}

// ------------------------------------------------------------------------------------------------------ backward, tcgen05
// Reverse mode of the block-diagonal kernel above (same tiles: 7 frames of 17 joints, or P whole tracks of a short clip, per 128-row
// tile; head_dim 64): five [128 x 128 x 64]-class tcgen05 MMAs per (tile, head) instead of one mma.sync CTA per (sequence, head).
//   TMA   Q, K, V (from qkv) and dO tiles [128 x 64]
//   MMA   S = Q K^T -> TMEM[0,128),  dP = dO V^T -> TMEM[128,256)
//   4 softmax warps, thread = query row i: e = exp2(scale (S - max)) written back over S (tcgen05.st), D_i = sum_j P_ij dP_ij
//         (= dO_i . O_i, so the forward output is not read), then P and dS = scale P (dP - D) as 16-bit rows (zeros outside the row's
//         window and for rows of other tiles) into shared memory: K-major A operands that the MN-major descriptors read transposed
//   MMA   dV = P^T dO -> TMEM[0,64),  dK = dS^T Q -> TMEM[64,128),  dQ = dS K -> TMEM[128,192)
//   epilogue: three [128 x 64] tiles staged over the dead Q / K / dO tiles, one TMA store each into the [.., 3C] gradient.
// 112 KB of shared memory and 256 TMEM columns: two CTAs per SM.
constexpr int kBwdTcSmem = 7 * 16384 + 64;

template <typename D, bool kTracks>
__global__ void __launch_bounds__(kTcThreads, 2)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_dqkv,
                   float* __restrict__ colsum, int n_rows, int n_tok, int C, int n_heads, int G, int tiles_per_clip) {
  pdl_launch_dependents();
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sq = smem;                   // [128 x 64] queries;        later the dQ staging tile
  uint8_t* sk = smem + 16384;           // [128 x 64] keys;           later dK
  uint8_t* sdo = smem + 32768;          // [128 x 64] output grads;   later dV
  uint8_t* sv = smem + 49152;           // [128 x 64] values;         later P k-block 0 (P k-block 1 follows at +16384)
  uint8_t* sp = sv;                     // [128 x 128] P, two k-blocks of [128 x 64]
  uint8_t* sds = smem + 81920;          // [128 x 128] dS, two k-blocks
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * 16384);
  uint64_t* bar_qk = bars + 0;
  uint64_t* bar_vdo = bars + 1;
  uint64_t* bar_s = bars + 2;
  uint64_t* bar_p = bars + 3;
  uint64_t* bar_o = bars + 4;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int head = blockIdx.x % n_heads;
  const int tile = (blockIdx.x / n_heads) % tiles_per_clip;
  const int clip = blockIdx.x / (n_heads * tiles_per_clip);
  const int row0 = tile * G;
  const int rows_valid = kTracks ? G * n_rows : 128;

  if (warp == 4 && lane == 0) {
    ptx::prefetch_tmap(&tm_in);
    ptx::prefetch_tmap(&tm_do);
    ptx::prefetch_tmap(&tm_dqkv);
    ptx::mbar_init(bar_qk, 1);
    ptx::mbar_init(bar_vdo, 1);
    ptx::mbar_init(bar_s, 1);
    ptx::mbar_init(bar_p, 128);
    ptx::mbar_init(bar_o, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 0) {
    ptx::tmem_alloc(tmem_holder, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;
  pdl_wait();

  if (warp == 4) {
    if (lane == 0) {
      if constexpr (kTracks) {
        const uint32_t box_bytes = (uint32_t)rows_valid * 128u;
        ptx::mbar_expect_tx(bar_qk, 2 * box_bytes);
        ptx::tma_load_4d(sq, &tm_in, bar_qk, head * 64, row0, 0, clip);
        ptx::tma_load_4d(sk, &tm_in, bar_qk, C + head * 64, row0, 0, clip);
        ptx::mbar_expect_tx(bar_vdo, 2 * box_bytes);
        ptx::tma_load_4d(sv, &tm_in, bar_vdo, 2 * C + head * 64, row0, 0, clip);
        ptx::tma_load_4d(sdo, &tm_do, bar_vdo, head * 64, row0, 0, clip);
      } else {
        ptx::mbar_expect_tx(bar_qk, 32768);
        ptx::tma_load_3d(sq, &tm_in, bar_qk, head * 64, row0, clip);
        ptx::tma_load_3d(sk, &tm_in, bar_qk, C + head * 64, row0, clip);
        ptx::mbar_expect_tx(bar_vdo, 32768);
        ptx::tma_load_3d(sv, &tm_in, bar_vdo, 2 * C + head * 64, row0, clip);
        ptx::tma_load_3d(sdo, &tm_do, bar_vdo, head * 64, row0, clip);
      }
      constexpr uint32_t idesc_kk = ptx::umma_idesc_16(128, 128, D::kUmmaFmt);
      ptx::mbar_wait(bar_qk, 0);
      ptx::tc_fence_after();
      {
        const uint64_t da = ptx::umma_desc_sw128(smem_u32(sq));
        const uint64_t db = ptx::umma_desc_sw128(smem_u32(sk));
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem_base, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_kk, k != 0 ? 1u : 0u);
      }
      ptx::mbar_wait(bar_vdo, 0);
      ptx::tc_fence_after();
      {
        const uint64_t da = ptx::umma_desc_sw128(smem_u32(sdo));
        const uint64_t db = ptx::umma_desc_sw128(smem_u32(sv));
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_f16(tmem_base + 128u, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc_kk, k != 0 ? 1u : 0u);
        ptx::umma_commit(bar_s);
      }
      ptx::mbar_wait(bar_p, 0);
      ptx::tc_fence_after();
      {
        // contraction over the 128 query rows, 16 per instruction (2048 bytes of either tile); the two 64-column halves of P / dS are
        // MN atoms 16 KB apart
        constexpr uint32_t idesc_t = ptx::umma_idesc_16_abmn(128, 64, D::kUmmaFmt);
        const uint32_t pa = smem_u32(sp), dsa = smem_u32(sds), doa = smem_u32(sdo), qa = smem_u32(sq), ka = smem_u32(sk);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          ptx::umma_f16(tmem_base, ptx::umma_desc_mn_sw128_lbo(pa + (uint32_t)(k * 2048), 16384u), ptx::umma_desc_mn_sw128(doa + (uint32_t)(k * 2048)),
                        idesc_t, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          ptx::umma_f16(tmem_base + 64u, ptx::umma_desc_mn_sw128_lbo(dsa + (uint32_t)(k * 2048), 16384u),
                        ptx::umma_desc_mn_sw128(qa + (uint32_t)(k * 2048)), idesc_t, k != 0 ? 1u : 0u);
        // dQ = dS K: contraction over the 128 key columns of dS (K-major A, as P in the forward), K rows as the MN-major B operand
        constexpr uint32_t idesc_q = ptx::umma_idesc_16_bmn(128, 64, D::kUmmaFmt);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          ptx::umma_f16(tmem_base + 128u, ptx::umma_desc_sw128(dsa + (uint32_t)((k >> 2) * 16384 + (k & 3) * 32)),
                        ptx::umma_desc_mn_sw128(ka + (uint32_t)(k * 2048)), idesc_q, k != 0 ? 1u : 0u);
        ptx::umma_commit(bar_o);
      }
    }
  } else {
    const int row = threadIdx.x;                       // tile-local query row <-> TMEM lane
    const uint32_t sw = (uint32_t)(row & 7);
    const uint32_t t_row = tmem_base + ((uint32_t)(32 * warp) << 16);
    const float scale = 0.125f, scale_log2 = 0.125f * kLog2e;
    int a = (row / n_tok) * n_tok;
    int b = a + n_tok;
    if (b > 128) b = 128;
    if (row0 + b > n_rows) b = n_rows - row0;
    bool live = row < G && row0 + row < n_rows && b > a;   // rows [G, 128) are the next tile's: no contribution to this tile's dK / dV
    int c_lo = ((32 * warp) / n_tok * n_tok) / 32;
    int c_hi = (((32 * warp + 31) / n_tok + 1) * n_tok + 31) / 32;
    if (c_hi > 4) c_hi = 4;
    uint32_t win[4] = {0u, 0u, 0u, 0u};            // bit i of win[c]: column 32 c + i belongs to this row's softmax window
    if constexpr (kTracks) {
      const int trk = row % G;                     // G = tracks per tile = interleave period
      live = row < rows_valid && row0 + trk < n_tok;
      c_lo = 0;
      c_hi = (rows_valid + 31) / 32;
      uint32_t pat = 0u;
      for (int i = 0; i < 32; i += G) pat |= 1u << i;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int start = ((trk - 32 * c) % G + G) % G;
        const int left = rows_valid - 32 * c;
        const uint32_t range = left >= 32 ? 0xffffffffu : (left > 0 ? (1u << left) - 1u : 0u);
        win[c] = (pat << start) & range;
      }
      // rows the track boxes do not fill hold stale shared memory: as contraction rows of dV / dK / dQ they must be zero (0 * NaN)
      for (int r = rows_valid + row; r < 128; r += 128) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          *reinterpret_cast<uint4*>(sq + (size_t)r * 128 + q * 16) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(sk + (size_t)r * 128 + q * 16) = make_uint4(0u, 0u, 0u, 0u);
          *reinterpret_cast<uint4*>(sdo + (size_t)r * 128 + q * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
      }
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int lo = a - 32 * c, hi = b - 32 * c;
        const uint32_t upto_hi = hi >= 32 ? 0xffffffffu : (hi > 0 ? (1u << hi) - 1u : 0u);
        const uint32_t upto_lo = lo >= 32 ? 0xffffffffu : (lo > 0 ? (1u << lo) - 1u : 0u);
        win[c] = upto_hi & ~upto_lo;
      }
    }
    if (!live) {
#pragma unroll
      for (int c = 0; c < 4; ++c) win[c] = 0u;
    }
    ptx::mbar_wait(bar_s, 0);
    ptx::tc_fence_after();
    // ---- row maximum over the window
    float mx = -INFINITY;
    for (int c = c_lo; c < c_hi; ++c) {
      uint32_t r[32];
      ptx::tmem_ld32(t_row + (uint32_t)(c * 32), r);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if ((win[c] >> i) & 1u) mx = fmaxf(mx, __uint_as_float(r[i]));
    }
    const float ms = live ? -mx * scale_log2 : 0.f;
    // ---- e = exp2(scale (s - max)) back over S, row sum, and sum_j e_j dP_j
    float sum = 0.f, edp = 0.f;
    for (int c = c_lo; c < c_hi; ++c) {
      uint32_t r[32], g[32];
      ptx::tmem_ld32(t_row + (uint32_t)(c * 32), r);
      ptx::tmem_ld32(t_row + 128u + (uint32_t)(c * 32), g);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const bool in = (win[c] >> i) & 1u;
        const float e = in ? fast_exp2(fmaf(__uint_as_float(r[i]), scale_log2, ms)) : 0.f;
        sum += e;
        edp = fmaf(e, in ? __uint_as_float(g[i]) : 0.f, edp);
        r[i] = __float_as_uint(e);
      }
      ptx::tmem_st32(t_row + (uint32_t)(c * 32), r);
    }
    ptx::tmem_st_wait();
    const float inv = live ? 1.0f / sum : 0.f;
    const float dsum = edp * inv;                      // D_i
    // ---- P and dS rows (16-bit), all 128 columns
    for (int c = 0; c < 4; ++c) {
      float pv[32], dv[32];
      if (c >= c_lo && c < c_hi) {
        uint32_t r[32], g[32];
        ptx::tmem_ld32(t_row + (uint32_t)(c * 32), r);
        ptx::tmem_ld32(t_row + 128u + (uint32_t)(c * 32), g);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const bool in = (win[c] >> i) & 1u;
          const float pr = __uint_as_float(r[i]) * inv;      // e is 0 outside the window
          pv[i] = pr;
          dv[i] = in ? pr * (__uint_as_float(g[i]) - dsum) * scale : 0.f;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) pv[i] = dv[i] = 0.f;
      }
      const size_t off = (size_t)(c >> 1) * 16384 + (size_t)row * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 o, d4;
        o.x = D::pack2(pv[8 * q + 0], pv[8 * q + 1]);
        o.y = D::pack2(pv[8 * q + 2], pv[8 * q + 3]);
        o.z = D::pack2(pv[8 * q + 4], pv[8 * q + 5]);
        o.w = D::pack2(pv[8 * q + 6], pv[8 * q + 7]);
        d4.x = D::pack2(dv[8 * q + 0], dv[8 * q + 1]);
        d4.y = D::pack2(dv[8 * q + 2], dv[8 * q + 3]);
        d4.z = D::pack2(dv[8 * q + 4], dv[8 * q + 5]);
        d4.w = D::pack2(dv[8 * q + 6], dv[8 * q + 7]);
        const uint32_t chunk = (((uint32_t)((c & 1) * 4 + q) ^ sw) << 4);
        *reinterpret_cast<uint4*>(sp + off + chunk) = o;
        *reinterpret_cast<uint4*>(sds + off + chunk) = d4;
      }
    }
    ptx::tc_fence_before();
    ptx::fence_proxy_async_smem();
    ptx::mbar_arrive(bar_p);
    ptx::mbar_wait(bar_o, 0);
    ptx::tc_fence_after();
    // ---- dV | dK | dQ: TMEM columns [0,64) [64,128) [128,192) -> staging tiles sdo | sk | sq
#pragma unroll
    for (int m = 0; m < 3; ++m) {
      uint32_t r0[32], r1[32];
      ptx::tmem_ld32(t_row + (uint32_t)(m * 64), r0);
      ptx::tmem_ld32(t_row + (uint32_t)(m * 64) + 32u, r1);
      ptx::tmem_ld_wait();
      uint8_t* orow = (m == 0 ? sdo : (m == 1 ? sk : sq)) + (size_t)row * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t(&r)[32] = q < 4 ? r0 : r1;
        const int bb = (q & 3) * 8;
        uint4 o;
        o.x = D::pack2(__uint_as_float(r[bb + 0]), __uint_as_float(r[bb + 1]));
        o.y = D::pack2(__uint_as_float(r[bb + 2]), __uint_as_float(r[bb + 3]));
        o.z = D::pack2(__uint_as_float(r[bb + 4]), __uint_as_float(r[bb + 5]));
        o.w = D::pack2(__uint_as_float(r[bb + 6]), __uint_as_float(r[bb + 7]));
        *reinterpret_cast<uint4*>(orow + (((uint32_t)q ^ sw) << 4)) = o;
      }
    }
    ptx::fence_proxy_async_smem();
    asm volatile("bar.sync 1, 128;" ::: "memory");
    if (threadIdx.x == 0) {
      if constexpr (kTracks) {
        ptx::tma_store_4d(&tm_dqkv, sq, head * 64, row0, 0, clip);
        ptx::tma_store_4d(&tm_dqkv, sk, C + head * 64, row0, 0, clip);
        ptx::tma_store_4d(&tm_dqkv, sdo, 2 * C + head * 64, row0, 0, clip);
      } else {
        ptx::tma_store_3d(&tm_dqkv, sq, head * 64, row0, clip);
        ptx::tma_store_3d(&tm_dqkv, sk, C + head * 64, row0, clip);
        ptx::tma_store_3d(&tm_dqkv, sdo, 2 * C + head * 64, row0, clip);
      }
      ptx::bulk_commit();
    }
    if (colsum != nullptr && threadIdx.x < 96) {
      // qkv bias gradient: column sums of the three staged tiles (rows that are not stored are exactly zero), one atomic per column
      const int m = threadIdx.x >> 5, cp = threadIdx.x & 31;                 // tile (dQ | dK | dV), column pair
      const uint8_t* tb = m == 0 ? sq : (m == 1 ? sk : sdo);
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
      for (int r = 0; r < 128; ++r) {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(tb + r * 128 + ((((uint32_t)(cp >> 2)) ^ (uint32_t)(r & 7)) << 4) + (cp & 3) * 4);
        const float2 f = D::unpack2(v);
        s0 += f.x;
        s1 += f.y;
      }
      atomicAdd(colsum + m * C + head * 64 + 2 * cp, s0);
      atomicAdd(colsum + m * C + head * 64 + 2 * cp + 1, s1);
    }
    if (threadIdx.x == 0) ptx::bulk_wait_read<0>();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 256);
  }
}

}  // namespace

// mp_attention_bwd (train.cu) for head_dim 64: spatial sequences of <= 32 tokens, temporal tracks of <= 128 frames.
int attention_bwd_tc(const void* qkv, const void* dout, void* dqkv, float* dqkv_colsum, int64_t n_clips, int64_t n_frames, int n_tok, int C,
                     int n_heads, int temporal, int dtype, cudaStream_t s) {
  const bool bf = dtype == MP_DTYPE_BF16;
  CUtensorMap tin, tdo, tdq;
  if (temporal) {
    const int P = 128 / (int)n_frames < n_tok ? 128 / (int)n_frames : n_tok;
    const int64_t tiles_per_clip = (n_tok + P - 1) / P;
    MP_REQUIRE(n_clips * tiles_per_clip * n_heads < ((int64_t)1 << 31), MP_EINVAL, "mp_attention_bwd: too many sequences");
    MP_CHECK(get_tmap_track(&tin, qkv, n_clips, n_frames, n_tok, 3 * C, (int)n_frames, dtype, P));
    MP_CHECK(get_tmap_track(&tdo, dout, n_clips, n_frames, n_tok, C, (int)n_frames, dtype, P));
    MP_CHECK(get_tmap_track(&tdq, dqkv, n_clips, n_frames, n_tok, 3 * C, (int)n_frames, dtype, P));
    auto launch = [&](auto kernel) {
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdTcSmem);
      launch_k(kernel, (unsigned)(n_clips * tiles_per_clip * n_heads), kTcThreads, kBwdTcSmem, s, tin, tdo, tdq, dqkv_colsum, (int)n_frames, n_tok, C, n_heads, P,
                                                                                          (int)tiles_per_clip);
    };
    if (bf) launch(attn_bwd_tc_kernel<Bf16, true>); else launch(attn_bwd_tc_kernel<Fp16, true>);
    return check_launch("attn_bwd_tc_kernel<tracks>");
  }
  const int G = (128 / n_tok) * n_tok;
  const int64_t rows_per_clip = n_frames * n_tok;
  const int64_t tiles_per_clip = (rows_per_clip + G - 1) / G;
  MP_REQUIRE(n_clips * tiles_per_clip * n_heads < ((int64_t)1 << 31) && rows_per_clip < ((int64_t)1 << 30), MP_EINVAL, "mp_attention_bwd: too many sequences");
  MP_CHECK(get_tmap_clip_rows(&tin, qkv, n_clips, rows_per_clip, 3 * C, 128, dtype));
  MP_CHECK(get_tmap_clip_rows(&tdo, dout, n_clips, rows_per_clip, C, 128, dtype));
  MP_CHECK(get_tmap_clip_rows(&tdq, dqkv, n_clips, rows_per_clip, 3 * C, G, dtype));
  auto launch = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdTcSmem);
    launch_k(kernel, (unsigned)(n_clips * tiles_per_clip * n_heads), kTcThreads, kBwdTcSmem, s, tin, tdo, tdq, dqkv_colsum, (int)rows_per_clip, n_tok, C, n_heads, G,
                                                                                        (int)tiles_per_clip);
  };
  if (bf) launch(attn_bwd_tc_kernel<Bf16, false>); else launch(attn_bwd_tc_kernel<Fp16, false>);
  return check_launch("attn_bwd_tc_kernel");
}

}  // namespace mp

extern "C" int mp_attention(const void* qkv, void* out, int64_t n_clips, int64_t n_frames, int n_tok, int C, int n_heads, int mode,
                            int dtype, mp_stream_t stream) {
  using namespace mp;
  MP_CHECK(require_sm100());
  MP_REQUIRE(qkv && out, MP_EINVAL, "mp_attention: null pointer");
  MP_REQUIRE(n_clips >= 0 && n_frames >= 1 && n_tok >= 1 && n_heads >= 1 && n_heads <= 8 && C % n_heads == 0, MP_EINVAL,
             "mp_attention: bad sizes");
  MP_REQUIRE(dtype == MP_DTYPE_BF16 || dtype == MP_DTYPE_FP16, MP_EINVAL, "mp_attention: unknown dtype %d", dtype);
  const int hd = C / n_heads;
  MP_REQUIRE(hd == 64 || hd == 16, MP_EUNSUPPORTED, "mp_attention: head_dim=%d (built for 64 and 16)", hd);
  MP_REQUIRE(aligned16(qkv) && aligned16(out), MP_EALIGN, "mp_attention: pointers must be 16-byte aligned");
  if (n_clips == 0) return MP_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const uint16_t* in = (const uint16_t*)qkv;
  uint16_t* o = (uint16_t*)out;
  const bool bf = dtype == MP_DTYPE_BF16;
  if (mode == MP_ATTN_TEMPORAL) {
    MP_REQUIRE(n_frames <= 256, MP_EUNSUPPORTED, "mp_attention: temporal sequences longer than 256 frames are not built (got %lld)",
               (long long)n_frames);
    const int Tp = ((int)n_frames + 31) & ~31;
    static const bool legacy = getenv("MANIPOSE_ATTN_MMA_SYNC") != nullptr;   // A/B switch: the mma.sync kernel below
    static const bool long_only = getenv("MANIPOSE_ATTN_TRACKS_OFF") != nullptr;   // A/B switch: always the T = 243 kernel
    if (hd == 64 && !legacy && !long_only && n_frames <= 128) {
      // short clips: P = 128 / T whole tracks per 128-row tile through the block-diagonal kernel (see attn_spatial_tc_kernel<.., true>)
      const int P = 128 / (int)n_frames < n_tok ? 128 / (int)n_frames : n_tok;
      const int64_t tiles_per_clip = (n_tok + P - 1) / P;
      MP_REQUIRE(n_clips * tiles_per_clip * n_heads < ((int64_t)1 << 31), MP_EINVAL, "mp_attention: too many sequences");
      CUtensorMap tin, to;
      MP_CHECK(get_tmap_track(&tin, qkv, n_clips, n_frames, n_tok, 3 * C, (int)n_frames, dtype, P));
      MP_CHECK(get_tmap_track(&to, out, n_clips, n_frames, n_tok, C, (int)n_frames, dtype, P));
      const int smem_tc = 49152 + 64;
      auto launch_tr = [&](auto kernel) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_tc);
        launch_k(kernel, (unsigned)(n_clips * tiles_per_clip * n_heads), kTcThreads, smem_tc, s, tin, to, (int)n_frames, n_tok, C, n_heads, P,
                                                                                         (int)tiles_per_clip);
      };
      if (bf) launch_tr(attn_spatial_tc_kernel<Bf16, true>); else launch_tr(attn_spatial_tc_kernel<Fp16, true>);
      return check_launch("attn_spatial_tc_kernel<tracks>");
    }
    if (hd == 64 && !legacy) {
      const int m_tiles = ((int)n_frames + 127) / 128;
      CUtensorMap tq, tkv, to;
      MP_CHECK(get_tmap_track(&tq, qkv, n_clips, n_frames, n_tok, 3 * C, 128, dtype));
      MP_CHECK(get_tmap_track(&tkv, qkv, n_clips, n_frames, n_tok, 3 * C, Tp, dtype));
      MP_CHECK(get_tmap_track(&to, out, n_clips, n_frames, n_tok, C, 128, dtype));
      const int64_t ctas = n_clips * n_tok * n_heads * m_tiles;
      MP_REQUIRE(ctas < ((int64_t)1 << 31), MP_EINVAL, "mp_attention: too many sequences");
      static const bool one_shot = getenv("MANIPOSE_ATTN_TC1") != nullptr;   // A/B switch: one CTA per (head, query tile)
      static const bool tc2 = getenv("MANIPOSE_ATTN_TC2") != nullptr;        // A/B switch: P staged through shared memory (tc2)
      if (!one_shot && !tc2 && n_frames > 128) {
        const int n_items = (int)(n_clips * n_tok * n_heads);
        const int grid = n_items < sm_count() ? n_items : sm_count();
        auto launch_3 = [&](auto kernel) {
          cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTc3Smem);
          launch_k(kernel, grid, kTc3Threads, kTc3Smem, s, tq, tkv, to, (int)n_frames, n_tok, C, n_heads, n_items);
        };
        if (bf) launch_3(attn_temporal_tc3_kernel<Bf16>); else launch_3(attn_temporal_tc3_kernel<Fp16>);
        return check_launch("attn_temporal_tc3_kernel");
      }
      if (!one_shot) {
        const int n_items = (int)(n_clips * n_tok * n_heads);
        const int smem_p = 3 * 65536 + 32768 + 256;
        const int grid = n_items < sm_count() ? n_items : sm_count();
        auto launch_p = [&](auto kernel) {
          cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_p);
          launch_k(kernel, grid, kTc2Threads, smem_p, s, tq, tkv, to, (int)n_frames, n_tok, C, n_heads, n_items);
        };
        if (bf) launch_p(attn_temporal_tc2_kernel<Bf16>); else launch_p(attn_temporal_tc2_kernel<Fp16>);
        return check_launch("attn_temporal_tc2_kernel");
      }
      const int smem_tc = 98304 + 64;
      auto launch_tc = [&](auto kernel) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_tc);
        launch_k(kernel, (unsigned)ctas, kTcThreads, smem_tc, s, tq, tkv, to, (int)n_frames, n_tok, C, n_heads, m_tiles);
      };
      if (bf) launch_tc(attn_temporal_tc_kernel<Bf16>); else launch_tc(attn_temporal_tc_kernel<Fp16>);
      return check_launch("attn_temporal_tc_kernel");
    }
    const size_t smem = (size_t)3 * Tp * (hd * 2 + 16);
    const int64_t ctas = n_clips * n_tok * n_heads;
    MP_REQUIRE(ctas < ((int64_t)1 << 31), MP_EINVAL, "mp_attention: too many sequences");
    auto launch = [&](auto kernel) {
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      launch_k(kernel, (unsigned)ctas, kTWarps * 32, smem, s, in, o, (int)n_frames, n_tok, C, n_heads);
    };
    if (hd == 64) {
      if (bf) launch(attn_temporal_kernel<64, Bf16>); else launch(attn_temporal_kernel<64, Fp16>);
    } else {
      if (bf) launch(attn_temporal_kernel<16, Bf16>); else launch(attn_temporal_kernel<16, Fp16>);
    }
    return check_launch("attn_temporal_kernel");
  }
  MP_REQUIRE(mode == MP_ATTN_SPATIAL, MP_EINVAL, "mp_attention: unknown mode %d", mode);
  MP_REQUIRE(n_tok <= 32, MP_EUNSUPPORTED, "mp_attention: spatial sequences longer than 32 tokens are not built (got %d)", n_tok);
  static const bool legacy_s = getenv("MANIPOSE_ATTN_MMA_SYNC") != nullptr;
  if (hd == 64 && !legacy_s) {
    const int G = (128 / n_tok) * n_tok;
    const int64_t rows_per_clip = n_frames * n_tok;
    const int64_t tiles_per_clip = (rows_per_clip + G - 1) / G;
    MP_REQUIRE(n_clips * tiles_per_clip * n_heads < ((int64_t)1 << 31) && rows_per_clip < ((int64_t)1 << 30), MP_EINVAL, "mp_attention: too many sequences");
    CUtensorMap tin, to;
    MP_CHECK(get_tmap_clip_rows(&tin, qkv, n_clips, rows_per_clip, 3 * C, 128, dtype));
    MP_CHECK(get_tmap_clip_rows(&to, out, n_clips, rows_per_clip, C, G, dtype));
    const int smem_tc = 49152 + 64;
    auto launch_tc = [&](auto kernel) {
      cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_tc);
      launch_k(kernel, (unsigned)(n_clips * tiles_per_clip * n_heads), kTcThreads, smem_tc, s, tin, to, (int)rows_per_clip, n_tok, C, n_heads, G,
                                                                                       (int)tiles_per_clip);
    };
    if (bf) launch_tc(attn_spatial_tc_kernel<Bf16, false>); else launch_tc(attn_spatial_tc_kernel<Fp16, false>);
    return check_launch("attn_spatial_tc_kernel");
  }
  const size_t smem = (size_t)kSWarps * 2 * 3 * 32 * (hd * 2 + 16);
  MP_REQUIRE(smem <= 227 * 1024, MP_EUNSUPPORTED, "mp_attention: %d spatial tokens need %zu bytes of shared memory", n_tok, smem);
  const int64_t n_seq = n_clips * n_frames;
  const int64_t n_items = n_seq * n_heads;
  int per_sm = (int)(220 * 1024 / (smem + 1024));
  per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > (n_items + kSWarps - 1) / kSWarps) grid = (n_items + kSWarps - 1) / kSWarps;
  auto launch = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_k(kernel, (unsigned)grid, kSWarps * 32, smem, s, in, o, n_seq, n_tok, C, n_heads);
  };
  if (hd == 64) {
    if (bf) launch(attn_spatial_kernel<64, Bf16>); else launch(attn_spatial_kernel<64, Fp16>);
  } else {
    if (bf) launch(attn_spatial_kernel<16, Bf16>); else launch(attn_spatial_kernel<16, Fp16>);
  }
  return check_launch("attn_spatial_kernel");
}

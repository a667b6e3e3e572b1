// Block-sparse attention backward for sm_100a (tcgen05 / TMEM / TMA), two passes that never materialise S or P:
//
//   dQ pass     one CTA per 128-query tile.  The tile's key slots (band + global block 0) are processed 3 at a time:
//               S = Q K^T and dP = dO V^T (128 x 96 each) into TMEM, dS = P o (dP - delta) * scale written over S as
//               16-bit, dQ += dS K (A from TMEM, K slots read MN-major) into a persistent 64-column accumulator.
//               Also: delta = rowsum(dO o O) (from the SMEM tiles, written for the second pass) and the gradient of the
//               GLOBAL key block 0 -- it receives contributions from every query tile, so each tile computes
//               [dO^T ; Q^T] (128 x 128q, the two stacked SMEM tiles read as ONE MN-major operand) x [P_0 | dS_0]
//               (128q x 64) and adds the two useful 64x32 quadrants to a small fp32 accumulator with red.global.add.
//   dK/dV pass  one CTA per 128-key tile.  The (left+3+nsup <= 7) query blocks attending the tile are processed 2 at a
//               time: S^T = K Q^T, dP^T = V dO^T (128 x 64), P^T / dS^T written over them as 16-bit, dV += P^T dO,
//               dK += dS^T Q (A from TMEM; the dO / Q slots re-read MN-major from the SMEM the first MMAs read K-major).
//
// Both kernels need only 256 TMEM columns and, for the default geometry (<= 8 key slots), <= 113 KB of shared memory,
// so two CTAs share an SM and overlap each other's load / MMA / CUDA-core phases (windows 5..10: 14 slot buffers, one
// CTA per SM); within a CTA the tensor pipe runs the next chunk's S/dP while nothing else
// depends on it (tcgen05.mma executes in issue order, which also orders the in-place TMEM reuse between chunks).
//
// Reference: autograd of sdd -> softmax -> dsd, sparse_vae/core/sparse_matmul.py:463-488 (dV = P^T dO, dP = dO V^T,
// dQ = dS K, dK = dS^T Q) and the block-sparse softmax backward dS = P o (dP - rowsum(dP o P)) * scale.
// Roofline: HBM-bound; algorithmic bytes = 8 * B*L*H*Dh * 2 (read Q,K,V,O,dO; write dQ,dK,dV).
#include "attn_sm100.cuh"

namespace svae {
namespace sm100 {

using namespace ptx;

constexpr int kBwdSlots = 8;          // key slots of the default geometry (two CTAs per SM)
constexpr int kBwdSlotsWide = 14;     // windows 5..10: one CTA per SM
constexpr int kBwdMathWarps = 8;                 // two warps per TMEM lane quarter: each owns 16 of a slot's 32 columns
constexpr int kBwdThreads = (kBwdMathWarps + 1) * 32;   // + one TMA / MMA-issue warp
   // key slots per query tile (dQ pass); query slots per key tile <= 7

// ------------------------------------------------------------------------------------------ dQ pass
template <int DH, int NS>
struct DqSmem {
  static constexpr int ROWB = DH * 2;
  static constexpr int TILE_BYTES = kTile * ROWB;
  static constexpr int SLOT_BYTES = kBlock * ROWB;
  static constexpr int OFF_DO = 0;                         // dO tile, immediately followed by the Q tile:
  static constexpr int OFF_Q = OFF_DO + TILE_BYTES;        //   together the MN-major stacked operand [dO^T ; Q^T]
  static constexpr int OFF_O = OFF_Q + TILE_BYTES;         // O tile (delta), then [128 q][32 P_0 | 32 dS_0] (128-byte rows)
  static constexpr int OFF_K = OFF_O + TILE_BYTES;
  static constexpr int OFF_V = OFF_K + NS * SLOT_BYTES;
  static constexpr int OFF_BAR = OFF_V + NS * SLOT_BYTES;
  static constexpr int DYN_BYTES = OFF_BAR + 256;          // no alignment slack: the base is checked to be 1024-aligned
  static constexpr int MAX_PASS = (NS + 2) / 3;
  static constexpr int CTAS = NS <= 8 ? 2 : 1;
  static constexpr int PASS = 3;                           // slots per chunk
  static constexpr int COL_S = 0, COL_DP = 32 * PASS, COL_DQ = 64 * PASS, COL_G = 32 * PASS;
  static_assert(COL_DQ + DH <= 256 && COL_G + 64 <= COL_DQ, "TMEM plan");
  static_assert(CTAS * (DYN_BYTES + 1024) <= 228 * 1024, "shared memory");
};

template <typename T, int DH, int NS>
__global__ void __launch_bounds__(kBwdThreads, NS <= 8 ? 2 : 1)
attn_bwd_dq_sm100_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                         const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmKband,
                         const __grid_constant__ CUtensorMap tmVband, const __grid_constant__ CUtensorMap tmKband2,
                         const __grid_constant__ CUtensorMap tmVband2, const __grid_constant__ CUtensorMap tmDQ,
                         const BwdParams p) {
  static_assert(DH == 64, "the stacked [dO^T ; Q^T] operand needs 2*DH == 128 rows");
  using S = DqSmem<DH, NS>;
  constexpr int ROWB = S::ROWB;
  constexpr int PASS = S::PASS;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t *sDO = smem + S::OFF_DO, *sQ = smem + S::OFF_Q, *sO = smem + S::OFF_O, *sK = smem + S::OFF_K,
          *sV = smem + S::OFF_V;
  uint8_t* sG = sO;                                         // reused once delta has been computed
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t *bar_ld = bars + 0, *bar_dq = bars + 1;
  uint64_t* bar_sdp = bars + 2;                             // [MAX_PASS] MMA -> threads: chunk's S and dP are in TMEM
  uint64_t* bar_ds = bars + 2 + S::MAX_PASS;                // [MAX_PASS] threads -> MMA: chunk's dS is in TMEM (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * S::MAX_PASS);

  const int warp = warp_index(), lane = threadIdx.x & 31;
  const int t = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const TileGeom g = p.g;
  const int ns = g.nslots;
  const int npass = (ns + PASS - 1) / PASS;
  const int r0 = 4 * t;
  const int band_lo = r0 - (g.left - 1);
  // SMEM slot order = processing order: band slots 0 .. nband-1, then the global block (slot nband), so that the
  // slots of a chunk are contiguous and one N = 96 MMA chain covers them
  auto order = [&](int i) { return i; };
  auto is_global = [&](int j) { return g.cls && j == g.nband; };
  auto slot_block = [&](int j) { return is_global(j) ? 0 : band_lo + j; };
  auto slot_valid = [&](int j) {
    if (is_global(j)) return true;
    int blk = band_lo + j;
    return blk >= g.cls && blk < g.nb;
  };

  if (warp == kBwdMathWarps) {
    // the issuing warp initialises the barriers and starts the TMA loads at once: they overlap the TMEM allocation
    // and the CTA-wide sync below instead of following them
    if (lane == 0) {
      if (smem_u32(smem) & 1023u) __trap();     // dynamic shared memory must be 1024-byte aligned (SWIZZLE_128B)
      mbar_init(bar_ld, 1);
      mbar_init(bar_dq, 1);
      for (int i = 0; i < S::MAX_PASS; ++i) { mbar_init(bar_sdp + i, 1); mbar_init(bar_ds + i, kBwdMathWarps * 32); }
      fence_barrier_init();
    }
    __syncwarp();
    mbar_arrive_expect_tx_w(bar_ld, 3 * S::TILE_BYTES + 2 * ns * S::SLOT_BYTES);
    tma_load_4d_w(sQ, &tmQ, bar_ld, 0, t * kTile, h, b);
    tma_load_4d_w(sDO, &tmDO, bar_ld, 0, t * kTile, h, b);
    tma_load_4d_w(sO, &tmO, bar_ld, 0, t * kTile, h, b);
    tma_load_4d_w(sK, &tmKband, bar_ld, 0, band_lo * kBlock, h, b);   // OOB rows -> zeros
    tma_load_4d_w(sV, &tmVband, bar_ld, 0, band_lo * kBlock, h, b);
    if (NS > 8 && g.nband > 8) {                                       // a TMA box holds at most 256 rows
      tma_load_4d_w(sK + 8 * S::SLOT_BYTES, &tmKband2, bar_ld, 0, (band_lo + 8) * kBlock, h, b);
      tma_load_4d_w(sV + 8 * S::SLOT_BYTES, &tmVband2, bar_ld, 0, (band_lo + 8) * kBlock, h, b);
    }
    if (g.cls) {
      tma_load_4d_w(sK + g.nband * S::SLOT_BYTES, &tmK, bar_ld, 0, 0, h, b);
      tma_load_4d_w(sV + g.nband * S::SLOT_BYTES, &tmV, bar_ld, 0, 0, h, b);
    }
    tmem_alloc<256>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kBwdMathWarps) {
    {   // the whole warp runs the issue path convergently; one lane is elected inside each wrapper
      const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sDO), k_addr = smem_u32(sK), v_addr = smem_u32(sV),
                     g_addr = smem_u32(sG);
      const uint32_t idesc_dq = make_idesc(kTile, DH, Elem<T>::fmt, 0, 1);
      auto issue_s_dp = [&](int c) {   // chunk c: S and dP of up to 3 contiguous slots, one N = 32*cnt MMA chain each
        const int cnt = (ns - c * PASS) < PASS ? (ns - c * PASS) : PASS;
        const uint32_t idesc_s = make_idesc(kTile, cnt * kBlock, Elem<T>::fmt, 0, 0);
        const uint32_t koff = c * PASS * S::SLOT_BYTES;
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {
          mma_ss_w(tmem_base + S::COL_S, make_smem_desc(q_addr + ks * 32, 16, 8 * ROWB, ROWB),
                   make_smem_desc(k_addr + koff + ks * 32, 16, 8 * ROWB, ROWB), idesc_s, ks > 0 ? 1u : 0u);
          mma_ss_w(tmem_base + S::COL_DP, make_smem_desc(do_addr + ks * 32, 16, 8 * ROWB, ROWB),
                   make_smem_desc(v_addr + koff + ks * 32, 16, 8 * ROWB, ROWB), idesc_s, ks > 0 ? 1u : 0u);
        }
      };
      mbar_wait(bar_ld, 0);
      tc_fence_after();
      issue_s_dp(0);
      tc_commit_w(bar_sdp + 0);
      uint32_t acc = 0;
      for (int c = 0; c < npass; ++c) {
        mbar_wait(bar_ds + c, 0);
        tc_fence_after();
        for (int i = 0; i < PASS && c * PASS + i < ns; ++i) {     // dQ += dS_j K_j
          const int j = order(c * PASS + i);
          if (!slot_valid(j)) continue;
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            mma_ts_w(tmem_base + S::COL_DQ, tmem_base + S::COL_S + 16 * i + 8 * s,
                   make_smem_desc(k_addr + j * S::SLOT_BYTES + s * 16 * ROWB, S::SLOT_BYTES, 8 * ROWB, ROWB), idesc_dq, acc);
            acc = 1;
          }
        }
        if (c + 1 < npass) {
          issue_s_dp(c + 1);            // in-order tensor pipe: overwrites the chunk's S/dS only after the dQ MMAs read it
          tc_commit_w(bar_sdp + c + 1);
        } else {
          if (g.cls) {
            // G[128 x 64] = [dO^T ; Q^T][128 x 128q] * [P_0 | dS_0][128q x 64] ; both operands MN-major.
            // A: two 64-row atoms (the dO tile and the Q tile) S::TILE_BYTES apart (LBO); 8 query rows per 1024 B (SBO).
            const uint32_t idesc_g = make_idesc(kTile, 64, Elem<T>::fmt, 1, 1);
#pragma unroll
            for (int s = 0; s < kTile / 16; ++s)
              mma_ss_w(tmem_base + S::COL_G, make_smem_desc(do_addr + s * 16 * ROWB, S::TILE_BYTES, 8 * ROWB, ROWB),
                     make_smem_desc(g_addr + s * 16 * 128, 16 * 128, 8 * 128, 128), idesc_g, s > 0 ? 1u : 0u);
          }
          tc_commit_w(bar_dq);
        }
      }
    }
    __syncwarp();
  } else {
    // warps w and w + 4 share TMEM lane quarter w & 3 (= block-row r0 + (w & 3)); each owns 16 of a slot's 32 columns
    const int qd = warp & 3, hc = warp >> 2;
    const int r = r0 + qd;
    const int row = qd * 32 + lane;
    const int qpos = t * kTile + row;
    const bool row_ok = qpos < p.L;
    const uint32_t trow = tmem_base + ((uint32_t)(qd * 32) << 16);
    const int64_t stat_idx = ((int64_t)b * p.H + h) * p.L + qpos;
    const float neg_lse2 = row_ok ? -p.lse[stat_idx] * kLog2e : 0.f;
    constexpr int kMath = kBwdMathWarps * 32;

    // does any key of this tile carry a non-zero additive mask?  (no: fast path without the mask term)
    uint32_t any_kpm = 0;
    if (p.kpm) {
      for (int i = threadIdx.x; i < ns * kBlock; i += kMath) {
        const int j = i >> 5;
        if (slot_valid(j)) any_kpm |= (p.kpm[(int64_t)b * p.L + slot_block(j) * kBlock + (i & 31)] != 0.f) ? 1u : 0u;
      }
    }

    // delta = rowsum(dO o O) from the swizzled SMEM tiles (both column halves evaluate it; one writes it)
    mbar_wait(bar_ld, 0);
    float delta = 0.f;
#pragma unroll
    for (int ch = 0; ch < ROWB / 16; ++ch) {
      const uint4 a = *reinterpret_cast<const uint4*>(sDO + swz_off<ROWB>(row, ch));
      const uint4 o = *reinterpret_cast<const uint4*>(sO + swz_off<ROWB>(row, ch));
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 fa = Elem<T>::unpack(aw[e]), fo = Elem<T>::unpack(ow[e]);
        delta = fmaf(fa.x, fo.x, delta);
        delta = fmaf(fa.y, fo.y, delta);
      }
    }
    if (row_ok && hc == 0) p.delta[stat_idx] = delta;
    const float neg_delta_s = -delta * p.scale;
    const bool has_kpm = bar_red_or(1, kMath, any_kpm);     // also: every thread is done reading sO (-> sG)

    auto slot_live = [&](int j) {
      if (r >= g.nb || !slot_valid(j)) return false;
      if (is_global(j)) return true;
      const int blk = band_lo + j;
      return blk >= r - (g.left - 1) && blk <= r + g.nsup;
    };
    const uint32_t below_diag = (lane == 31) ? 0xffffffffu : ((2u << lane) - 1u);   // bit c set <=> key c <= query lane

    for (int c = 0; c < npass; ++c) {
      mbar_wait(bar_sdp + c, 0);
      tc_fence_after();
      bool wrote_g = false;
      for (int i = 0; i < PASS && c * PASS + i < ns; ++i) {
        const int j = order(c * PASS + i);
        const bool is_g = is_global(j);
        const bool live = slot_live(j);
        uint32_t sv[16], dv[16];
        if (live) {
          tmem_ld16(trow + S::COL_S + 32 * i + 16 * hc, sv);
          tmem_ld16(trow + S::COL_DP + 32 * i + 16 * hc, dv);
          tmem_wait_ld16(sv, dv);
        }
        // dS of slot i overwrites S columns [16 i, 16 i + 16), which hold scores the PARTNER warp reads: both column
        // halves must have pulled their scores of every slot <= i into registers before either stores
        named_bar_sync(2 + qd, 64);
        uint32_t dsk[8];
        if (live) {
          uint32_t pk[8];
          const bool diag = g.causal && (slot_block(j) == r);
          float kmine = 0.f;
          if (has_kpm) kmine = p.kpm[(int64_t)b * p.L + slot_block(j) * kBlock + lane] * kLog2e;
#pragma unroll
          for (int cc = 0; cc < 16; cc += 2) {
            const int col = 16 * hc + cc;
            float x0 = fmaf(__uint_as_float(sv[cc]), p.scale_log2, neg_lse2);
            float x1 = fmaf(__uint_as_float(sv[cc + 1]), p.scale_log2, neg_lse2);
            if (has_kpm) {
              x0 += __shfl_sync(0xffffffffu, kmine, col);
              x1 += __shfl_sync(0xffffffffu, kmine, col + 1);
            }
            float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
            if (diag) {
              if (!((below_diag >> col) & 1u)) p0 = 0.f;
              if (!((below_diag >> (col + 1)) & 1u)) p1 = 0.f;
            }
            const float d0 = p0 * fmaf(__uint_as_float(dv[cc]), p.scale, neg_delta_s);
            const float d1 = p1 * fmaf(__uint_as_float(dv[cc + 1]), p.scale, neg_delta_s);
            dsk[cc >> 1] = Elem<T>::pack(d0, d1);
            if (is_g) pk[cc >> 1] = Elem<T>::pack(p0, p1);
          }
          if (is_g) {     // row of [P_0 (64 B) | dS_0 (64 B)]: this half owns two 16-byte chunks of each
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
              *reinterpret_cast<uint4*>(sG + swz_off<128>(row, 2 * hc + ch)) = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
              *reinterpret_cast<uint4*>(sG + swz_off<128>(row, 4 + 2 * hc + ch)) = make_uint4(dsk[4 * ch], dsk[4 * ch + 1], dsk[4 * ch + 2], dsk[4 * ch + 3]);
            }
            wrote_g = true;
          }
        } else {
#pragma unroll
          for (int cc = 0; cc < 8; ++cc) dsk[cc] = 0u;
          if (is_g) {
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
              *reinterpret_cast<uint4*>(sG + swz_off<128>(row, 2 * hc + ch)) = make_uint4(0, 0, 0, 0);
              *reinterpret_cast<uint4*>(sG + swz_off<128>(row, 4 + 2 * hc + ch)) = make_uint4(0, 0, 0, 0);
            }
            wrote_g = true;
          }
        }
        tmem_st8(trow + S::COL_S + 16 * i + 8 * hc, dsk);
      }
      if (wrote_g) fence_proxy_async();        // sG is read by the tensor core through the async proxy
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(bar_ds + c);
    }

    mbar_wait(bar_dq, 0);
    tc_fence_after();
    {   // this half's 32 of the 64 dQ columns -> 16-bit -> swizzled staging tile (the dO tile is free by now)
      uint32_t v[32];
      tmem_ld32(trow + S::COL_DQ + 32 * hc, v);
      tmem_wait_ld(v);
#pragma unroll
      for (int cq = 0; cq < 4; ++cq) {
        uint4 w;
        w.x = Elem<T>::pack(__uint_as_float(v[cq * 8 + 0]), __uint_as_float(v[cq * 8 + 1]));
        w.y = Elem<T>::pack(__uint_as_float(v[cq * 8 + 2]), __uint_as_float(v[cq * 8 + 3]));
        w.z = Elem<T>::pack(__uint_as_float(v[cq * 8 + 4]), __uint_as_float(v[cq * 8 + 5]));
        w.w = Elem<T>::pack(__uint_as_float(v[cq * 8 + 6]), __uint_as_float(v[cq * 8 + 7]));
        *reinterpret_cast<uint4*>(sDO + swz_off<ROWB>(row, hc * 4 + cq)) = w;
      }
    }
    fence_proxy_async();
    named_bar_sync(1, kMath);
    if (threadIdx.x == 0) {
      tma_store_4d(&tmDQ, sDO, 0, t * kTile, h, b);
      tma_store_commit();
    }
    if (g.cls) {
      // rows 0..63 of G: dV_0^T[d][key] in columns 0..31 ; rows 64..127: dK_0^T[d][key] in columns 32..63
      const int which = row < 64 ? 1 : 0;             // gacc[..., 0] = dK, gacc[..., 1] = dV
      const int d = row & 63;
      uint32_t v[16];
      tmem_ld16(trow + S::COL_G + (which ? 0 : 32) + 16 * hc, v);
      tmem_wait_ld16(v);
      float* dst = p.gacc + (((int64_t)b * p.H + h) * 2 + which) * (kBlock * DH) + d;
#pragma unroll
      for (int cc = 0; cc < 16; ++cc) atomicAdd(dst + (16 * hc + cc) * DH, __uint_as_float(v[cc]));
    }
    if (threadIdx.x == 0) tma_store_wait_read();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kBwdMathWarps) tmem_dealloc<256>(tmem_base);
}

// ------------------------------------------------------------------------------------------ dK/dV pass
template <int DH, int NQ_>
struct DkvSmem {
  static constexpr int ROWB = DH * 2;
  static constexpr int TILE_BYTES = kTile * ROWB;
  static constexpr int SLOT_BYTES = kBlock * ROWB;
  static constexpr int NQ = NQ_;                        // query-block slots (left + 3 + nsup)
  static constexpr int MAX_PASS = (NQ + 1) / 2;
  static constexpr int CTAS = NQ <= 7 ? 2 : 1;
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + TILE_BYTES;
  static constexpr int OFF_Q = OFF_V + TILE_BYTES;
  static constexpr int OFF_DO = OFF_Q + NQ * SLOT_BYTES;
  static constexpr int OFF_LSE = OFF_DO + NQ * SLOT_BYTES;          // -lse * log2(e) per query column
  static constexpr int OFF_DELTA = OFF_LSE + NQ * kBlock * 4;       // -delta * scale per query column
  static constexpr int OFF_BAR = OFF_DELTA + NQ * kBlock * 4;
  static constexpr int DYN_BYTES = OFF_BAR + 256 + 1024;
  static constexpr int PASS = 2;
  static constexpr int COL_ST = 0, COL_DPT = 64, COL_DV = 128, COL_DK = 192;
  static_assert(CTAS * (DYN_BYTES + 1024) <= 228 * 1024, "shared memory");
};

template <typename T, int DH, int NQ>
__global__ void __launch_bounds__(kThreads, NQ <= 7 ? 2 : 1)
attn_bwd_dkv_sm100_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                          const __grid_constant__ CUtensorMap tmQband, const __grid_constant__ CUtensorMap tmDOband,
                          const __grid_constant__ CUtensorMap tmQband2, const __grid_constant__ CUtensorMap tmDOband2,
                          const __grid_constant__ CUtensorMap tmDK, const __grid_constant__ CUtensorMap tmDV,
                          const BwdParams p) {
  using S = DkvSmem<DH, NQ>;
  constexpr int ROWB = S::ROWB;
  constexpr int PASS = S::PASS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sK = smem + S::OFF_K, *sV = smem + S::OFF_V, *sQ = smem + S::OFF_Q, *sDO = smem + S::OFF_DO;
  float* sNegLse = reinterpret_cast<float*>(smem + S::OFF_LSE);
  float* sNegDelta = reinterpret_cast<float*>(smem + S::OFF_DELTA);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t *bar_ld = bars + 0, *bar_out = bars + 1;
  uint64_t* bar_sdp = bars + 2;                       // [MAX_PASS]
  uint64_t* bar_pds = bars + 2 + S::MAX_PASS;         // [MAX_PASS], 128 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 + 2 * S::MAX_PASS);

  const int warp = warp_index(), lane = threadIdx.x & 31;
  const int t = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const TileGeom g = p.g;
  const int nq = g.nband;
  const int npass = (nq + PASS - 1) / PASS;
  const int c0 = 4 * t;
  const int q_lo = c0 - g.nsup;                       // query block of slot 0
  auto slot_valid = [&](int i) { int qb = q_lo + i; return qb >= 0 && qb < g.nb; };

  if (warp == 4) {
    // barriers + TMA loads first: they overlap the TMEM allocation and the CTA-wide sync
    if (lane == 0) {
      mbar_init(bar_ld, 1);
      mbar_init(bar_out, 1);
      for (int i = 0; i < S::MAX_PASS; ++i) { mbar_init(bar_sdp + i, 1); mbar_init(bar_pds + i, 128); }
      fence_barrier_init();
    }
    __syncwarp();
    mbar_arrive_expect_tx_w(bar_ld, 2 * S::TILE_BYTES + 2 * nq * S::SLOT_BYTES);
    tma_load_4d_w(sK, &tmK, bar_ld, 0, t * kTile, h, b);
    tma_load_4d_w(sV, &tmV, bar_ld, 0, t * kTile, h, b);
    tma_load_4d_w(sQ, &tmQband, bar_ld, 0, q_lo * kBlock, h, b);      // rows outside [0, L) -> zeros
    tma_load_4d_w(sDO, &tmDOband, bar_ld, 0, q_lo * kBlock, h, b);
    if (NQ > 8 && nq > 8) {                                            // a TMA box holds at most 256 rows
      tma_load_4d_w(sQ + 8 * S::SLOT_BYTES, &tmQband2, bar_ld, 0, (q_lo + 8) * kBlock, h, b);
      tma_load_4d_w(sDO + 8 * S::SLOT_BYTES, &tmDOband2, bar_ld, 0, (q_lo + 8) * kBlock, h, b);
    }
    tmem_alloc<256>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    {   // warp-convergent issue path
      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), q_addr = smem_u32(sQ), do_addr = smem_u32(sDO);
      const uint32_t idesc_o = make_idesc(kTile, DH, Elem<T>::fmt, 0, 1);
      auto issue_s_dp = [&](int c) {   // S^T = K Q^T, dP^T = V dO^T for the chunk's (<= 2) query slots
        const int cnt = (nq - c * PASS) < PASS ? (nq - c * PASS) : PASS;
        const uint32_t idesc_s = make_idesc(kTile, cnt * kBlock, Elem<T>::fmt, 0, 0);
#pragma unroll
        for (int ks = 0; ks < DH / 16; ++ks) {
          mma_ss_w(tmem_base + S::COL_ST, make_smem_desc(k_addr + ks * 32, 16, 8 * ROWB, ROWB),
                 make_smem_desc(q_addr + c * PASS * S::SLOT_BYTES + ks * 32, 16, 8 * ROWB, ROWB), idesc_s, ks > 0 ? 1u : 0u);
          mma_ss_w(tmem_base + S::COL_DPT, make_smem_desc(v_addr + ks * 32, 16, 8 * ROWB, ROWB),
                 make_smem_desc(do_addr + c * PASS * S::SLOT_BYTES + ks * 32, 16, 8 * ROWB, ROWB), idesc_s, ks > 0 ? 1u : 0u);
        }
      };
      mbar_wait(bar_ld, 0);
      tc_fence_after();
      issue_s_dp(0);
      tc_commit_w(bar_sdp + 0);
      uint32_t acc = 0;
      for (int c = 0; c < npass; ++c) {
        mbar_wait(bar_pds + c, 0);
        tc_fence_after();
        for (int i = 0; i < PASS && c * PASS + i < nq; ++i) {
          const int slot = c * PASS + i;
#pragma unroll
          for (int s = 0; s < 2; ++s) {
            // dV += P^T dO ; dK += dS^T Q   (B operands: the dO / Q slots, MN-major)
            mma_ts_w(tmem_base + S::COL_DV, tmem_base + S::COL_ST + 16 * i + 8 * s,
                   make_smem_desc(do_addr + slot * S::SLOT_BYTES + s * 16 * ROWB, S::SLOT_BYTES, 8 * ROWB, ROWB), idesc_o, acc);
            mma_ts_w(tmem_base + S::COL_DK, tmem_base + S::COL_DPT + 16 * i + 8 * s,
                   make_smem_desc(q_addr + slot * S::SLOT_BYTES + s * 16 * ROWB, S::SLOT_BYTES, 8 * ROWB, ROWB), idesc_o, acc);
            acc = 1;
          }
        }
        if (c + 1 < npass) {
          issue_s_dp(c + 1);
          tc_commit_w(bar_sdp + c + 1);
        } else {
          tc_commit_w(bar_out);
        }
      }
    }
    __syncwarp();
  } else {
    const int c = c0 + warp;                           // this warp's key block
    const int row = warp * 32 + lane;
    const int kpos = t * kTile + row;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float kv = (p.kpm && kpos < p.L) ? p.kpm[(int64_t)b * p.L + kpos] * kLog2e : 0.f;
    const bool warp_kpm = __any_sync(0xffffffffu, kv != 0.f);

    for (int i = threadIdx.x; i < nq * kBlock; i += 128) {
      const int s = i >> 5, n = i & 31;
      float l2 = 0.f, dl = 0.f;
      if (slot_valid(s)) {
        const int64_t idx = ((int64_t)b * p.H + h) * p.L + (q_lo + s) * kBlock + n;
        l2 = -p.lse[idx] * kLog2e;
        dl = -p.delta[idx] * p.scale;
      }
      sNegLse[i] = l2;
      sNegDelta[i] = dl;
    }
    named_bar_sync(1, 128);

    const bool key_global = g.cls && c == 0;           // handled by the dQ pass
    auto slot_live = [&](int i) {
      if (c >= g.nb || key_global || !slot_valid(i)) return false;
      const int qb = q_lo + i;
      return c >= qb - (g.left - 1) && c <= qb + g.nsup;
    };
    const uint32_t at_or_after = ~((1u << lane) - 1u);   // bit n set <=> query n >= key lane (same block)

    for (int cpass = 0; cpass < npass; ++cpass) {
      mbar_wait(bar_sdp + cpass, 0);
      tc_fence_after();
      for (int i = 0; i < PASS && cpass * PASS + i < nq; ++i) {
        const int slot = cpass * PASS + i;
        uint32_t pk[16], dsk[16];
        if (slot_live(slot)) {
          uint32_t sv[32], dv[32];
          tmem_ld32(trow + S::COL_ST + 32 * i, sv);
          tmem_ld32(trow + S::COL_DPT + 32 * i, dv);
          tmem_wait_ld(sv, dv);
          const bool diag = g.causal && (q_lo + slot == c);
          const float* ls = sNegLse + slot * kBlock;
          const float* dl = sNegDelta + slot * kBlock;
#pragma unroll
          for (int n = 0; n < 32; n += 2) {
            float x0 = fmaf(__uint_as_float(sv[n]), p.scale_log2, ls[n]);
            float x1 = fmaf(__uint_as_float(sv[n + 1]), p.scale_log2, ls[n + 1]);
            if (warp_kpm) { x0 += kv; x1 += kv; }
            float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
            if (diag) {                                  // key position > query position
              if (!((at_or_after >> n) & 1u)) p0 = 0.f;
              if (!((at_or_after >> (n + 1)) & 1u)) p1 = 0.f;
            }
            const float d0 = p0 * fmaf(__uint_as_float(dv[n]), p.scale, dl[n]);
            const float d1 = p1 * fmaf(__uint_as_float(dv[n + 1]), p.scale, dl[n + 1]);
            pk[n >> 1] = Elem<T>::pack(p0, p1);
            dsk[n >> 1] = Elem<T>::pack(d0, d1);
          }
        } else {
#pragma unroll
          for (int n = 0; n < 16; ++n) pk[n] = dsk[n] = 0u;
        }
        tmem_st16(trow + S::COL_ST + 16 * i, pk);
        tmem_st16(trow + S::COL_DPT + 16 * i, dsk);
      }
      tmem_wait_st();
      tc_fence_before();
      mbar_arrive(bar_pds + cpass);
    }

    mbar_wait(bar_out, 0);
    tc_fence_after();
    const float* gk = p.gacc + (((int64_t)b * p.H + h) * 2 + 0) * (kBlock * DH) + lane * DH;
    const float* gv = gk + kBlock * DH;
#pragma unroll
    for (int which = 0; which < 2; ++which) {          // 0: dK -> sK, 1: dV -> sV
      uint8_t* stage = which ? sV : sK;
      const float* ga = which ? gv : gk;
#pragma unroll
      for (int half = 0; half < DH / 32; ++half) {
        uint32_t v[32];
        tmem_ld32(trow + (which ? S::COL_DV : S::COL_DK) + 32 * half, v);
        tmem_wait_ld(v);
        if (key_global) {
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) + __ldcg(ga + half * 32 + e));
        }
#pragma unroll
        for (int cq = 0; cq < 4; ++cq) {
          uint4 w;
          w.x = Elem<T>::pack(__uint_as_float(v[cq * 8 + 0]), __uint_as_float(v[cq * 8 + 1]));
          w.y = Elem<T>::pack(__uint_as_float(v[cq * 8 + 2]), __uint_as_float(v[cq * 8 + 3]));
          w.z = Elem<T>::pack(__uint_as_float(v[cq * 8 + 4]), __uint_as_float(v[cq * 8 + 5]));
          w.w = Elem<T>::pack(__uint_as_float(v[cq * 8 + 6]), __uint_as_float(v[cq * 8 + 7]));
          *reinterpret_cast<uint4*>(stage + swz_off<ROWB>(row, half * 4 + cq)) = w;
        }
      }
    }
    fence_proxy_async();
    named_bar_sync(1, 128);
    if (threadIdx.x == 0) {
      tma_store_4d(&tmDK, sK, 0, t * kTile, h, b);
      tma_store_4d(&tmDV, sV, 0, t * kTile, h, b);
      tma_store_commit();
      tma_store_wait_read();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<256>(tmem_base);
}

// ------------------------------------------------------------------------------------------ host side
bool bwd_supported(const svae_attn_desc* d) {
  if (d->dtype != SVAE_DTYPE_BF16 && d->dtype != SVAE_DTYPE_F16) return false;
  if (d->head_dim != 64 || !(d->scale > 0.f)) return false;
  const TileGeom g = make_geom(d->window_size, d->causal, d->include_cls, d->seq_len / d->block_size);
  return g.nslots <= kBwdSlotsWide && g.nband <= kBwdSlotsWide - 1;
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

size_t bwd_workspace(const svae_attn_desc* d) {
  const size_t stats = align256(sizeof(float) * (size_t)d->batch * d->heads * d->seq_len);
  const size_t gacc = align256(sizeof(float) * (size_t)d->batch * d->heads * 2 * kBlock * d->head_dim);
  return stats + gacc;
}

template <typename T, int DH, int NS>
static int launch_bwd(const svae_attn_desc* d, const void* q, const void* k, const void* v, const void* out,
                      const void* dout, const float* lse, const float* kpm, void* dq, void* dk, void* dv,
                      void* workspace, cudaStream_t st) {
  const TileGeom g = make_geom(d->window_size, d->causal, d->include_cls, d->seq_len / d->block_size);
  const int B = d->batch, H = d->heads, L = d->seq_len;
  float* delta = reinterpret_cast<float*>(workspace);
  float* gacc = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(workspace) + align256(sizeof(float) * (size_t)B * H * L));
  const size_t gacc_bytes = sizeof(float) * (size_t)B * H * 2 * kBlock * DH;
  SVAE_CUDA_CHECK(cudaMemsetAsync(gacc, 0, gacc_bytes, st));

  BwdParams p;
  p.kpm = kpm; p.lse = lse; p.delta = delta; p.gacc = gacc;
  p.L = L; p.H = H; p.g = g;
  p.scale = d->scale; p.scale_log2 = d->scale * kLog2e;

  const CUtensorMapDataType dt = Elem<T>::tm;
  CUtensorMap tQ128, tDO128, tO128, tK32, tV32, tKband, tVband, tKband2, tVband2, tDQ128, tK128, tV128, tQband, tDOband,
      tQband2, tDOband2, tDK128, tDV128;
  int rc;
  // the band arrives with one TMA box of <= 8 blocks (256 rows) plus, for the wide windows, a second box with the rest
  const int band_rows = (g.nband <= 8 ? g.nband : 8) * kBlock;
  const int band2_rows = (g.nband > 8 ? g.nband - 8 : 1) * kBlock;
#define SVAE_TM(map, ptr, strd, rows) \
  if ((rc = encode_tmap(&map, dt, ptr, DH, L, H, B, strd, rows))) return rc
  SVAE_TM(tQ128, q, d->q_stride, kTile);         SVAE_TM(tDO128, dout, d->do_stride, kTile);
  SVAE_TM(tO128, out, d->o_stride, kTile);       SVAE_TM(tK32, k, d->k_stride, kBlock);
  SVAE_TM(tV32, v, d->v_stride, kBlock);         SVAE_TM(tKband, k, d->k_stride, band_rows);
  SVAE_TM(tVband, v, d->v_stride, band_rows);    SVAE_TM(tDQ128, dq, d->dq_stride, kTile);
  SVAE_TM(tK128, k, d->k_stride, kTile);         SVAE_TM(tV128, v, d->v_stride, kTile);
  SVAE_TM(tQband, q, d->q_stride, band_rows);    SVAE_TM(tDOband, dout, d->do_stride, band_rows);
  SVAE_TM(tDK128, dk, d->dk_stride, kTile);      SVAE_TM(tDV128, dv, d->dv_stride, kTile);
  SVAE_TM(tKband2, k, d->k_stride, band2_rows);  SVAE_TM(tVband2, v, d->v_stride, band2_rows);
  SVAE_TM(tQband2, q, d->q_stride, band2_rows);  SVAE_TM(tDOband2, dout, d->do_stride, band2_rows);
#undef SVAE_TM

  using SQ = DqSmem<DH, NS>;
  using SKV = DkvSmem<DH, NS - 1>;
  auto kq = attn_bwd_dq_sm100_kernel<T, DH, NS>;
  auto kkv = attn_bwd_dkv_sm100_kernel<T, DH, NS - 1>;
  SVAE_CONFIGURE_SMEM(kq, SQ::DYN_BYTES);
  SVAE_CONFIGURE_SMEM(kkv, SKV::DYN_BYTES);
  dim3 grid((L + kTile - 1) / kTile, H, B);
  {
    ScopedKernelTimer timer("attn_bwd_dq_sm100", st);
    kq<<<grid, kBwdThreads, SQ::DYN_BYTES, st>>>(tQ128, tDO128, tO128, tK32, tV32, tKband, tVband, tKband2, tVband2, tDQ128, p);
  }
  SVAE_CUDA_CHECK(cudaGetLastError());
  {
    ScopedKernelTimer timer("attn_bwd_dkv_sm100", st);
    kkv<<<grid, kThreads, SKV::DYN_BYTES, st>>>(tK128, tV128, tQband, tDOband, tQband2, tDOband2, tDK128, tDV128, p);
  }
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}

int bwd(const svae_attn_desc* d, const void* q, const void* k, const void* v, const void* out, const void* dout,
        const float* lse, const float* kpm, void* dq, void* dk, void* dv, void* workspace, cudaStream_t st) {
  const TileGeom g = make_geom(d->window_size, d->causal, d->include_cls, d->seq_len / d->block_size);
  const bool small = g.nslots <= kBwdSlots && g.nband <= kBwdSlots - 1;      // two CTAs per SM
  if (d->dtype == SVAE_DTYPE_BF16)
    return small ? launch_bwd<__nv_bfloat16, 64, kBwdSlots>(d, q, k, v, out, dout, lse, kpm, dq, dk, dv, workspace, st)
                 : launch_bwd<__nv_bfloat16, 64, kBwdSlotsWide>(d, q, k, v, out, dout, lse, kpm, dq, dk, dv, workspace, st);
  return small ? launch_bwd<__half, 64, kBwdSlots>(d, q, k, v, out, dout, lse, kpm, dq, dk, dv, workspace, st)
               : launch_bwd<__half, 64, kBwdSlotsWide>(d, q, k, v, out, dout, lse, kpm, dq, dk, dv, workspace, st);
}

}  // namespace sm100
}  // namespace svae

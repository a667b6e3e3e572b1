// Block-sparse attention backward for sm_100a (tcgen05 / TMEM / TMA), two passes that never materialise S or P:
//
//   pass 1  dQ   one CTA per 128-query tile:   S = Q K^T, dP = dO V^T (both 128 x 32*slots in TMEM),
//                dS = P o (dP - delta) * scale written over S as 16-bit, dQ = dS K (A from TMEM, K MN-major).
//                Also: delta = rowsum(dO o O) (from the SMEM tiles, written for pass 2) and the gradient of the
//                GLOBAL key block 0 -- it receives contributions from every query tile, so each tile computes
//                [dO^T ; Q^T] (128 x 128q, MN-major stacked SMEM tiles) x [P_0 | dS_0] (128q x 64) with one MMA
//                chain and adds the two useful 64x32 quadrants into a small fp32 accumulator with atomics.
//   pass 2  dK/dV  one CTA per 128-key tile:   S^T = K Q^T, dP^T = V dO^T over the (left+3+nsup) query blocks that
//                attend the tile, P^T / dS^T written over them as 16-bit, dV = P^T dO, dK = dS^T Q (A from TMEM,
//                dO / Q slots read MN-major from the same SMEM the first two MMAs read K-major).
//
// Reference: autograd of sdd -> softmax -> dsd, sparse_vae/core/sparse_matmul.py:463-488 (dV = P^T dO, dP = dO V^T,
// dQ = dS K, dK = dS^T Q) and the block-sparse softmax backward dS = P o (dP - rowsum(dP o P)) * scale.
// Roofline: HBM-bound; algorithmic bytes = 8 * B*L*H*Dh * 2 (read Q,K,V,O,dO; write dQ,dK,dV).
#include "attn_sm100.cuh"

namespace svae {
namespace sm100 {

using namespace ptx;

constexpr int kBwdSlots = 8;   // fwd-style key slots per query tile (dQ pass); query slots per key tile <= 7

// ------------------------------------------------------------------------------------------ dQ pass
template <int DH>
struct DqSmem {
  static constexpr int ROWB = DH * 2;
  static constexpr int TILE_BYTES = kTile * ROWB;
  static constexpr int SLOT_BYTES = kBlock * ROWB;
  static constexpr int OFF_DO = 0;                         // dO tile, immediately followed by the Q tile:
  static constexpr int OFF_Q = OFF_DO + TILE_BYTES;        //   together the MN-major stacked operand [dO^T ; Q^T]
  static constexpr int OFF_O = OFF_Q + TILE_BYTES;         // O tile (delta), later the dQ staging tile
  static constexpr int OFF_K = OFF_O + TILE_BYTES;
  static constexpr int OFF_V = OFF_K + kBwdSlots * SLOT_BYTES;
  static constexpr int OFF_G = OFF_V + kBwdSlots * SLOT_BYTES;   // [128 q][32 P_0 | 32 dS_0] 16-bit, 128-byte rows
  static constexpr int OFF_KPM = OFF_G + kTile * 128;
  static constexpr int OFF_BAR = OFF_KPM + kBwdSlots * kBlock * 4;
  static constexpr int TOTAL = OFF_BAR + 64;
  static constexpr int DYN_BYTES = TOTAL + 1024;
  static constexpr int COL_S = 0, COL_DP = 256, COL_G = 128, COL_DQ = 256;
};

template <typename T, int DH>
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_dq_sm100_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                         const __grid_constant__ CUtensorMap tmO, const __grid_constant__ CUtensorMap tmK,
                         const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmDQ,
                         const BwdParams p) {
  static_assert(DH == 64, "the stacked [dO^T ; Q^T] operand needs 2*DH == 128 rows");
  using S = DqSmem<DH>;
  constexpr int ROWB = S::ROWB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sDO = smem + S::OFF_DO, *sQ = smem + S::OFF_Q, *sO = smem + S::OFF_O, *sK = smem + S::OFF_K,
          *sV = smem + S::OFF_V, *sG = smem + S::OFF_G;
  float* sKpm = reinterpret_cast<float*>(smem + S::OFF_KPM);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t *bar_q = bars + 0, *bar_k = bars + 1, *bar_v = bars + 2, *bar_s = bars + 3, *bar_dp = bars + 4,
           *bar_ds = bars + 5, *bar_dq = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const TileGeom g = p.g;
  const int ns = g.nslots;
  const int r0 = 4 * t;
  const int band_lo = r0 - (g.left - 1);
  auto slot_block = [&](int j) { return (g.cls && j == 0) ? 0 : band_lo + j - g.cls; };
  auto slot_valid = [&](int j) {
    if (g.cls && j == 0) return true;
    int blk = band_lo + j - g.cls;
    return blk >= g.cls && blk < g.nb;
  };

  if (threadIdx.x == 0) {
    mbar_init(bar_q, 1); mbar_init(bar_k, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_dp, 1);
    mbar_init(bar_ds, 128); mbar_init(bar_dq, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    if (lane == 0) {
      prefetch_tensormap(&tmQ); prefetch_tensormap(&tmDO); prefetch_tensormap(&tmO);
      prefetch_tensormap(&tmK); prefetch_tensormap(&tmV); prefetch_tensormap(&tmDQ);
    }
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      int nvalid = 0;
      for (int j = 0; j < ns; ++j) nvalid += slot_valid(j) ? 1 : 0;
      mbar_arrive_expect_tx(bar_q, 3 * S::TILE_BYTES);
      tma_load_4d(sQ, &tmQ, bar_q, 0, t * kTile, h, b);
      tma_load_4d(sDO, &tmDO, bar_q, 0, t * kTile, h, b);
      tma_load_4d(sO, &tmO, bar_q, 0, t * kTile, h, b);
      mbar_arrive_expect_tx(bar_k, nvalid * S::SLOT_BYTES);
      for (int j = 0; j < ns; ++j)
        if (slot_valid(j)) tma_load_4d(sK + j * S::SLOT_BYTES, &tmK, bar_k, 0, slot_block(j) * kBlock, h, b);
      mbar_arrive_expect_tx(bar_v, nvalid * S::SLOT_BYTES);
      for (int j = 0; j < ns; ++j)
        if (slot_valid(j)) tma_load_4d(sV + j * S::SLOT_BYTES, &tmV, bar_v, 0, slot_block(j) * kBlock, h, b);

      const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sDO), k_addr = smem_u32(sK), v_addr = smem_u32(sV),
                     g_addr = smem_u32(sG);
      const int ntot = ns * kBlock;                     // <= 256
      const uint32_t idesc_s = make_idesc(kTile, ntot, Elem<T>::fmt, 0, 0);
      // S = Q K^T
      mbar_wait(bar_q, 0);
      mbar_wait(bar_k, 0);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks)
        mma_ss(tmem_base + S::COL_S, make_smem_desc(q_addr + ks * 32, 16, 8 * ROWB, ROWB),
               make_smem_desc(k_addr + ks * 32, 16, 8 * ROWB, ROWB), idesc_s, ks > 0 ? 1u : 0u);
      tc_commit(bar_s);
      // dP = dO V^T
      mbar_wait(bar_v, 0);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks)
        mma_ss(tmem_base + S::COL_DP, make_smem_desc(do_addr + ks * 32, 16, 8 * ROWB, ROWB),
               make_smem_desc(v_addr + ks * 32, 16, 8 * ROWB, ROWB), idesc_s, ks > 0 ? 1u : 0u);
      tc_commit(bar_dp);

      // dQ = dS K  (A: 16-bit dS in TMEM, B: K slots MN-major)  and the global-column accumulators
      mbar_wait(bar_ds, 0);
      tc_fence_after();
      const uint32_t idesc_dq = make_idesc(kTile, DH, Elem<T>::fmt, 0, 1);
      uint32_t acc = 0;
      for (int j = 0; j < ns; ++j) {
        if (!slot_valid(j)) continue;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          mma_ts(tmem_base + S::COL_DQ, tmem_base + S::COL_S + 16 * j + 8 * s,
                 make_smem_desc(k_addr + j * S::SLOT_BYTES + s * 16 * ROWB, S::SLOT_BYTES, 8 * ROWB, ROWB), idesc_dq, acc);
          acc = 1;
        }
      }
      if (g.cls) {
        // G[128 x 64] = [dO^T ; Q^T][128 x 128q] * [P_0 | dS_0][128q x 64] ; both operands MN-major.
        // A: two 64-row atoms (the dO tile and the Q tile) S::TILE_BYTES apart (LBO); 8 query rows per 1024 B (SBO).
        const uint32_t idesc_g = make_idesc(kTile, 64, Elem<T>::fmt, 1, 1);
#pragma unroll
        for (int s = 0; s < kTile / 16; ++s)
          mma_ss(tmem_base + S::COL_G, make_smem_desc(do_addr + s * 16 * ROWB, S::TILE_BYTES, 8 * ROWB, ROWB),
                 make_smem_desc(g_addr + s * 16 * 128, 16 * 128, 8 * 128, 128), idesc_g, s > 0 ? 1u : 0u);
      }
      tc_commit(bar_dq);
    }
    __syncwarp();
  } else {
    const int r = r0 + warp;
    const int row = warp * 32 + lane;
    const int qpos = t * kTile + row;
    const bool row_ok = qpos < p.L;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const int64_t stat_idx = ((int64_t)b * p.H + h) * p.L + qpos;

    for (int i = threadIdx.x; i < ns * kBlock; i += 128) {
      const int j = i >> 5, c = i & 31;
      float kv = 0.f;
      if (p.kpm && slot_valid(j)) kv = p.kpm[(int64_t)b * p.L + slot_block(j) * kBlock + c] * kLog2e;
      sKpm[i] = kv;
    }
    const float lse2 = row_ok ? p.lse[stat_idx] * kLog2e : 0.f;

    // delta = rowsum(dO o O) from the swizzled SMEM tiles
    mbar_wait(bar_q, 0);
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
    if (row_ok) p.delta[stat_idx] = delta;
    named_bar_sync(1, 128);     // sKpm staged; every thread is done reading sO (reused for dQ staging)

    auto slot_live = [&](int j) {
      if (r >= g.nb || !slot_valid(j)) return false;
      if (g.cls && j == 0) return true;
      const int blk = band_lo + j - g.cls;
      return blk >= r - (g.left - 1) && blk <= r + g.nsup;
    };

    mbar_wait(bar_s, 0);
    mbar_wait(bar_dp, 0);
    tc_fence_after();
    for (int j = 0; j < ns; ++j) {
      uint32_t dsk[16];
      const bool is_g = g.cls && j == 0;
      if (slot_live(j)) {
        uint32_t sv[32], dv[32];
        tmem_ld32(trow + S::COL_S + 32 * j, sv);
        tmem_ld32(trow + S::COL_DP + 32 * j, dv);
        tmem_wait_ld();
        const bool diag = g.causal && (slot_block(j) == r);
        const float* kp = sKpm + j * kBlock;
        uint32_t pk[16];
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float x0 = fmaf(__uint_as_float(sv[c]), p.scale_log2, kp[c]) - lse2;
          float x1 = fmaf(__uint_as_float(sv[c + 1]), p.scale_log2, kp[c + 1]) - lse2;
          float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
          if (diag && c > lane) p0 = 0.f;
          if (diag && c + 1 > lane) p1 = 0.f;
          const float d0 = p0 * (__uint_as_float(dv[c]) - delta) * p.scale;
          const float d1 = p1 * (__uint_as_float(dv[c + 1]) - delta) * p.scale;
          dsk[c >> 1] = Elem<T>::pack(d0, d1);
          pk[c >> 1] = Elem<T>::pack(p0, p1);
        }
        if (is_g) {
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            *reinterpret_cast<uint4*>(sG + swz_off<128>(row, ch)) = make_uint4(pk[4 * ch], pk[4 * ch + 1], pk[4 * ch + 2], pk[4 * ch + 3]);
            *reinterpret_cast<uint4*>(sG + swz_off<128>(row, 4 + ch)) = make_uint4(dsk[4 * ch], dsk[4 * ch + 1], dsk[4 * ch + 2], dsk[4 * ch + 3]);
          }
        }
      } else {
#pragma unroll
        for (int c = 0; c < 16; ++c) dsk[c] = 0u;
        if (is_g) {
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) *reinterpret_cast<uint4*>(sG + swz_off<128>(row, ch)) = make_uint4(0, 0, 0, 0);
        }
      }
      tmem_st16(trow + S::COL_S + 16 * j, dsk);
    }
    fence_proxy_async();          // sG is read by the tensor core through the async proxy
    tmem_wait_st();
    tc_fence_before();
    mbar_arrive(bar_ds);

    mbar_wait(bar_dq, 0);
    tc_fence_after();
#pragma unroll
    for (int half = 0; half < DH / 32; ++half) {
      uint32_t v[32];
      tmem_ld32(trow + S::COL_DQ + 32 * half, v);
      tmem_wait_ld();
#pragma unroll
      for (int cq = 0; cq < 4; ++cq) {
        uint4 w;
        w.x = Elem<T>::pack(__uint_as_float(v[cq * 8 + 0]), __uint_as_float(v[cq * 8 + 1]));
        w.y = Elem<T>::pack(__uint_as_float(v[cq * 8 + 2]), __uint_as_float(v[cq * 8 + 3]));
        w.z = Elem<T>::pack(__uint_as_float(v[cq * 8 + 4]), __uint_as_float(v[cq * 8 + 5]));
        w.w = Elem<T>::pack(__uint_as_float(v[cq * 8 + 6]), __uint_as_float(v[cq * 8 + 7]));
        *reinterpret_cast<uint4*>(sO + swz_off<ROWB>(row, half * 4 + cq)) = w;
      }
    }
    fence_proxy_async();
    named_bar_sync(1, 128);
    if (threadIdx.x == 0) {
      tma_store_4d(&tmDQ, sO, 0, t * kTile, h, b);
      tma_store_commit();
    }
    if (g.cls) {
      // rows 0..63 of G: dV_0^T[d][key] in columns 0..31 ; rows 64..127: dK_0^T[d][key] in columns 32..63
      const int which = row < 64 ? 1 : 0;             // gacc[..., 0] = dK, gacc[..., 1] = dV
      const int d = row & 63;
      uint32_t v[32];
      tmem_ld32(trow + S::COL_G + (which ? 0 : 32), v);
      tmem_wait_ld();
      float* dst = p.gacc + (((int64_t)b * p.H + h) * 2 + which) * (kBlock * DH) + d;
#pragma unroll
      for (int c = 0; c < 32; ++c) atomicAdd(dst + c * DH, __uint_as_float(v[c]));
    }
    if (threadIdx.x == 0) tma_store_wait_read();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------ dK/dV pass
template <int DH>
struct DkvSmem {
  static constexpr int ROWB = DH * 2;
  static constexpr int TILE_BYTES = kTile * ROWB;
  static constexpr int SLOT_BYTES = kBlock * ROWB;
  static constexpr int NQ = kBwdSlots - 1;              // query-block slots (left + 3 + nsup <= 7)
  static constexpr int OFF_K = 0;
  static constexpr int OFF_V = OFF_K + TILE_BYTES;
  static constexpr int OFF_Q = OFF_V + TILE_BYTES;
  static constexpr int OFF_DO = OFF_Q + NQ * SLOT_BYTES;
  static constexpr int OFF_LSE = OFF_DO + NQ * SLOT_BYTES;
  static constexpr int OFF_DELTA = OFF_LSE + NQ * kBlock * 4;
  static constexpr int OFF_BAR = OFF_DELTA + NQ * kBlock * 4;
  static constexpr int TOTAL = OFF_BAR + 64;
  static constexpr int DYN_BYTES = TOTAL + 1024;
  static constexpr int COL_ST = 0, COL_DPT = 256, COL_DV = 128, COL_DK = 384;
};

template <typename T, int DH>
__global__ void __launch_bounds__(kThreads, 1)
attn_bwd_dkv_sm100_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                          const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                          const __grid_constant__ CUtensorMap tmDK, const __grid_constant__ CUtensorMap tmDV,
                          const BwdParams p) {
  using S = DkvSmem<DH>;
  constexpr int ROWB = S::ROWB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t *sK = smem + S::OFF_K, *sV = smem + S::OFF_V, *sQ = smem + S::OFF_Q, *sDO = smem + S::OFF_DO;
  float* sLse = reinterpret_cast<float*>(smem + S::OFF_LSE);
  float* sDelta = reinterpret_cast<float*>(smem + S::OFF_DELTA);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t *bar_kq = bars + 0, *bar_vdo = bars + 1, *bar_sdp = bars + 2, *bar_pds = bars + 3, *bar_out = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const TileGeom g = p.g;
  const int nq = g.nband;
  const int c0 = 4 * t;
  const int q_lo = c0 - g.nsup;                       // query block of slot 0
  auto slot_valid = [&](int i) { int qb = q_lo + i; return qb >= 0 && qb < g.nb; };

  if (threadIdx.x == 0) {
    mbar_init(bar_kq, 1); mbar_init(bar_vdo, 1); mbar_init(bar_sdp, 1); mbar_init(bar_pds, 128); mbar_init(bar_out, 1);
    fence_barrier_init();
  }
  if (warp == 4) {
    if (lane == 0) {
      prefetch_tensormap(&tmK); prefetch_tensormap(&tmV); prefetch_tensormap(&tmQ);
      prefetch_tensormap(&tmDO); prefetch_tensormap(&tmDK); prefetch_tensormap(&tmDV);
    }
    tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    if (lane == 0) {
      int nvalid = 0;
      for (int i = 0; i < nq; ++i) nvalid += slot_valid(i) ? 1 : 0;
      mbar_arrive_expect_tx(bar_kq, S::TILE_BYTES + nvalid * S::SLOT_BYTES);
      tma_load_4d(sK, &tmK, bar_kq, 0, t * kTile, h, b);
      for (int i = 0; i < nq; ++i)
        if (slot_valid(i)) tma_load_4d(sQ + i * S::SLOT_BYTES, &tmQ, bar_kq, 0, (q_lo + i) * kBlock, h, b);
      mbar_arrive_expect_tx(bar_vdo, S::TILE_BYTES + nvalid * S::SLOT_BYTES);
      tma_load_4d(sV, &tmV, bar_vdo, 0, t * kTile, h, b);
      for (int i = 0; i < nq; ++i)
        if (slot_valid(i)) tma_load_4d(sDO + i * S::SLOT_BYTES, &tmDO, bar_vdo, 0, (q_lo + i) * kBlock, h, b);

      const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV), q_addr = smem_u32(sQ), do_addr = smem_u32(sDO);
      const uint32_t idesc_s = make_idesc(kTile, nq * kBlock, Elem<T>::fmt, 0, 0);
      mbar_wait(bar_kq, 0);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks)     // S^T = K Q^T
        mma_ss(tmem_base + S::COL_ST, make_smem_desc(k_addr + ks * 32, 16, 8 * ROWB, ROWB),
               make_smem_desc(q_addr + ks * 32, 16, 8 * ROWB, ROWB), idesc_s, ks > 0 ? 1u : 0u);
      mbar_wait(bar_vdo, 0);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < DH / 16; ++ks)     // dP^T = V dO^T
        mma_ss(tmem_base + S::COL_DPT, make_smem_desc(v_addr + ks * 32, 16, 8 * ROWB, ROWB),
               make_smem_desc(do_addr + ks * 32, 16, 8 * ROWB, ROWB), idesc_s, ks > 0 ? 1u : 0u);
      tc_commit(bar_sdp);

      mbar_wait(bar_pds, 0);
      tc_fence_after();
      const uint32_t idesc_o = make_idesc(kTile, DH, Elem<T>::fmt, 0, 1);
      uint32_t acc = 0;
      for (int i = 0; i < nq; ++i) {
        if (!slot_valid(i)) continue;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          // dV += P^T dO ; dK += dS^T Q   (B operands: the dO / Q slots, MN-major)
          mma_ts(tmem_base + S::COL_DV, tmem_base + S::COL_ST + 16 * i + 8 * s,
                 make_smem_desc(do_addr + i * S::SLOT_BYTES + s * 16 * ROWB, S::SLOT_BYTES, 8 * ROWB, ROWB), idesc_o, acc);
          mma_ts(tmem_base + S::COL_DK, tmem_base + S::COL_DPT + 16 * i + 8 * s,
                 make_smem_desc(q_addr + i * S::SLOT_BYTES + s * 16 * ROWB, S::SLOT_BYTES, 8 * ROWB, ROWB), idesc_o, acc);
          acc = 1;
        }
      }
      tc_commit(bar_out);
    }
    __syncwarp();
  } else {
    const int c = c0 + warp;                           // this warp's key block
    const int row = warp * 32 + lane;
    const int kpos = t * kTile + row;
    const uint32_t trow = tmem_base + ((uint32_t)(warp * 32) << 16);
    const float kv = (p.kpm && kpos < p.L) ? p.kpm[(int64_t)b * p.L + kpos] * kLog2e : 0.f;

    for (int i = threadIdx.x; i < nq * kBlock; i += 128) {
      const int s = i >> 5, n = i & 31;
      float l2 = 0.f, dl = 0.f;
      if (slot_valid(s)) {
        const int64_t idx = ((int64_t)b * p.H + h) * p.L + (q_lo + s) * kBlock + n;
        l2 = p.lse[idx] * kLog2e;
        dl = p.delta[idx];
      }
      sLse[i] = l2;
      sDelta[i] = dl;
    }
    named_bar_sync(1, 128);

    const bool key_global = g.cls && c == 0;           // handled by the dQ pass
    auto slot_live = [&](int i) {
      if (c >= g.nb || key_global || !slot_valid(i)) return false;
      const int qb = q_lo + i;
      return c >= qb - (g.left - 1) && c <= qb + g.nsup;
    };

    mbar_wait(bar_sdp, 0);
    tc_fence_after();
    for (int i = 0; i < nq; ++i) {
      uint32_t pk[16], dsk[16];
      if (slot_live(i)) {
        uint32_t sv[32], dv[32];
        tmem_ld32(trow + S::COL_ST + 32 * i, sv);
        tmem_ld32(trow + S::COL_DPT + 32 * i, dv);
        tmem_wait_ld();
        const bool diag = g.causal && (q_lo + i == c);
        const float* ls = sLse + i * kBlock;
        const float* dl = sDelta + i * kBlock;
#pragma unroll
        for (int n = 0; n < 32; n += 2) {
          float p0 = fast_exp2(fmaf(__uint_as_float(sv[n]), p.scale_log2, kv) - ls[n]);
          float p1 = fast_exp2(fmaf(__uint_as_float(sv[n + 1]), p.scale_log2, kv) - ls[n + 1]);
          if (diag && lane > n) p0 = 0.f;              // key position > query position
          if (diag && lane > n + 1) p1 = 0.f;
          const float d0 = p0 * (__uint_as_float(dv[n]) - dl[n]) * p.scale;
          const float d1 = p1 * (__uint_as_float(dv[n + 1]) - dl[n + 1]) * p.scale;
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
    mbar_arrive(bar_pds);

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
        tmem_wait_ld();
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
  if (warp == 4) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------ host side
bool bwd_supported(const svae_attn_desc* d) {
  if (d->dtype != SVAE_DTYPE_BF16 && d->dtype != SVAE_DTYPE_F16) return false;
  if (d->head_dim != 64) return false;
  const TileGeom g = make_geom(d->window_size, d->causal, d->include_cls, d->seq_len / d->block_size);
  return g.nslots <= kBwdSlots && g.nband <= kBwdSlots - 1;
}

static size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

size_t bwd_workspace(const svae_attn_desc* d) {
  const size_t stats = align256(sizeof(float) * (size_t)d->batch * d->heads * d->seq_len);
  const size_t gacc = align256(sizeof(float) * (size_t)d->batch * d->heads * 2 * kBlock * d->head_dim);
  return stats + gacc;
}

template <typename T, int DH>
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
  CUtensorMap tQ128, tDO128, tO128, tK32, tV32, tDQ128, tK128, tV128, tQ32, tDO32, tDK128, tDV128;
  int rc;
#define SVAE_TM(map, ptr, strd, rows) \
  if ((rc = encode_tmap(&map, dt, ptr, DH, L, H, B, strd, rows))) return rc
  SVAE_TM(tQ128, q, d->q_stride, kTile);    SVAE_TM(tDO128, dout, d->do_stride, kTile);
  SVAE_TM(tO128, out, d->o_stride, kTile);  SVAE_TM(tK32, k, d->k_stride, kBlock);
  SVAE_TM(tV32, v, d->v_stride, kBlock);    SVAE_TM(tDQ128, dq, d->dq_stride, kTile);
  SVAE_TM(tK128, k, d->k_stride, kTile);    SVAE_TM(tV128, v, d->v_stride, kTile);
  SVAE_TM(tQ32, q, d->q_stride, kBlock);    SVAE_TM(tDO32, dout, d->do_stride, kBlock);
  SVAE_TM(tDK128, dk, d->dk_stride, kTile); SVAE_TM(tDV128, dv, d->dv_stride, kTile);
#undef SVAE_TM

  auto kq = attn_bwd_dq_sm100_kernel<T, DH>;
  auto kkv = attn_bwd_dkv_sm100_kernel<T, DH>;
  static bool configured = false;
  if (!configured) {
    SVAE_CUDA_CHECK(cudaFuncSetAttribute(kq, cudaFuncAttributeMaxDynamicSharedMemorySize, DqSmem<DH>::DYN_BYTES));
    SVAE_CUDA_CHECK(cudaFuncSetAttribute(kkv, cudaFuncAttributeMaxDynamicSharedMemorySize, DkvSmem<DH>::DYN_BYTES));
    configured = true;
  }
  dim3 grid((L + kTile - 1) / kTile, H, B);
  {
    ScopedKernelTimer timer("attn_bwd_dq_sm100", st);
    kq<<<grid, kThreads, DqSmem<DH>::DYN_BYTES, st>>>(tQ128, tDO128, tO128, tK32, tV32, tDQ128, p);
  }
  SVAE_CUDA_CHECK(cudaGetLastError());
  {
    ScopedKernelTimer timer("attn_bwd_dkv_sm100", st);
    kkv<<<grid, kThreads, DkvSmem<DH>::DYN_BYTES, st>>>(tK128, tV128, tQ32, tDO32, tDK128, tDV128, p);
  }
  SVAE_CUDA_CHECK(cudaGetLastError());
  return SVAE_OK;
}

int bwd(const svae_attn_desc* d, const void* q, const void* k, const void* v, const void* out, const void* dout,
        const float* lse, const float* kpm, void* dq, void* dk, void* dv, void* workspace, cudaStream_t st) {
  if (d->dtype == SVAE_DTYPE_BF16) return launch_bwd<__nv_bfloat16, 64>(d, q, k, v, out, dout, lse, kpm, dq, dk, dv, workspace, st);
  return launch_bwd<__half, 64>(d, q, k, v, out, dout, lse, kpm, dq, dk, dv, workspace, st);
}

}  // namespace sm100
}  // namespace svae

// emcid_b200 — causal self-attention of the CLIP text encoder on tcgen05 (head dim 64, captions <= 128 tokens).
//
// One (caption group, head) unit at a time per persistent CTA — a group is a run of consecutive packed captions with at
// most `lp` tokens in all (clip.cuh::clip_group_captions_kernel): a full-length caption is a group of its own, short ones
// (real caption data: a dozen tokens; the prompts of an edit: six) share a tile, and the causal mask becomes block
// diagonal (row i sees columns [start of i's caption, i]).  Both products run on the tensor core with the same 3-term
// fp16 split as every other GEMM of the path:
//   S = Q K^T          A = Q planes [128 x 64], B = K planes [LP x 64]        TMA from the q|k|v planes [T x 3h]
//   P = softmax(S * dh^-1/2, causal)     fp32 in registers, one query row per thread (TMEM lane)
//   O = P V            A = P planes [128 x LP] written to swizzled smem by the threads,
//                      B = V planes [LP x 64] as an MN-major operand (the same TMA tile shape as Q and K: a TMA
//                      load along the contiguous token axis of a transposed copy would need 16-byte aligned
//                      caption starts, which packed captions do not have)
// Rows of the 128-row tiles beyond the caption belong to the next captions (finite, never stored) or are
// zero-filled by TMA; columns j > i, j >= L and j before row i's own caption get P = 0.  Replaces the softmax(QK^T)V of
// transformers modeling_clip.py::CLIPAttention under the reference's `model(**batch)` (emcid/layer_stats.py:215).
#pragma once

#include "host.cuh"

namespace emcid {

constexpr int ATTN_DH = 64;
constexpr int ATTN_THREADS = 128;
// smem: Q hi/lo | K hi/lo | V hi/lo | barriers; the P planes (hi/lo x 2 k-blocks) are written over Q and K, which
// are dead once S = Q K^T has retired.  Every tile holds lp rows of one 128-byte swizzle row (lp * 128 B, a multiple of
// 2 KB).  The A operands (Q, P) are read by the MMA as 128-row tiles: rows [lp, 128) alias whatever follows in shared
// memory — garbage in, garbage out, in accumulator rows that are never stored (an output row depends on its own A row
// only) — so the V tiles must stay behind them.  TMEM: 128 columns; O = P V accumulates over the columns S occupied
// (S lives in registers by then).  With CLIP's 77 tokens (lp = 80) the set is 62 KB and THREE CTAs are resident per
// SM: one unit's serial chain (TMA -> S -> softmax -> P V -> store) hides behind the other two.
inline int attn_tile_bytes(int lp) { return lp * 128; }
inline int attn_smem_bytes(int lp) { return 6 * attn_tile_bytes(lp) + 1024 /*barriers*/ + 1024 /*align slack*/; }
inline int attn_ctas_per_sm(int lp) {
  const int by_smem = (227 * 1024) / (attn_smem_bytes(lp) + 1024);
  const int cap = lp <= 80 ? 3 : 2;   // register file: 3 x 128 threads x 168 registers (the NC = 5 instantiation)
  return by_smem < 1 ? 1 : (by_smem < cap ? by_smem : cap);
}

struct AttnMaps {
  CUtensorMap qk_hi, qk_lo;   // q|k|v planes [T x 3h], box 64 x 128: the store maps of the q/k/v projection
  CUtensorMap kv_hi, kv_lo;   // same tensor, box 64 x lp (Q, K and V tiles of one caption)
};

// NC: 16-column chunks of S a thread keeps in registers (lp <= 16 * NC).
// GROUPED = false: one caption per unit, `offs` = cu_seqlens, n_units_host = captions x heads (the headline's 77-token
// captions: two never fit a tile, and the grouped instantiation is 9 % slower per launch there — same-box A/B,
// profiles/round2/r05t_ab_grouped_attention.txt).  GROUPED = true: `offs` = group offsets, their number read from the
// device (n_grp), tok_start[t] = first token of t's caption; the host picks it when captions average under half a tile.
template <int NC, bool GROUPED>
__global__ void __launch_bounds__(ATTN_THREADS, NC <= 5 ? 3 : 2)
clip_attention_tc_kernel(const __grid_constant__ AttnMaps tm, const int* __restrict__ offs, int n_units_host,
                         const int* __restrict__ n_grp, const int* __restrict__ tok_start, int heads, int h,
                         int lp /* padded caption / group length: multiple of 16, <= 128 */, float scale,
                         uint16_t* __restrict__ o_hi, uint16_t* __restrict__ o_lo, int ldo) {
  extern __shared__ uint8_t attn_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(attn_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  const int kt = lp * 128;                       // bytes of one tile
  uint64_t* bar_qk = reinterpret_cast<uint64_t*>(smem + 6 * kt);
  uint64_t* bar_v = bar_qk + 1;
  uint64_t* bar_s = bar_qk + 2;   // S = Q K^T finished
  uint64_t* bar_o = bar_qk + 3;   // O = P V finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_qk + 4);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    tma_prefetch_desc(&tm.kv_hi); tma_prefetch_desc(&tm.kv_lo);
    mbar_init(bar_qk, 1); mbar_init(bar_v, 1);
    mbar_init(bar_s, 1); mbar_init(bar_o, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_slot, 128);   // S: columns [0, lp); O: columns [0, 64) once S has been read
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
  const uint32_t s_q = smem_u32(smem);          // Q hi | Q lo
  const uint32_t s_k = s_q + 2 * kt, s_v = s_k + 2 * kt;
  const uint32_t s_p = s_q;                     // P hi kb0 | P hi kb1 | P lo kb0 | P lo kb1 over Q and K
  const int n_units = GROUPED ? *n_grp * heads : n_units_host;
  const int ksteps2 = lp / 16;                 // k-steps of the second product
  const uint32_t idesc1 = make_idesc(FMT_F16, FMT_F16, 128, static_cast<uint32_t>(lp));
  const uint32_t idesc2 = make_idesc(FMT_F16, FMT_F16, 128, ATTN_DH) | (1u << 16);   // B (= V) is MN-major

  auto issue_loads = [&](int unit, int t0) {
    const int head = unit % heads;
    uint8_t* base = smem;
    mbar_arrive_expect_tx(bar_qk, 4 * kt);
    tma_load_2d(base, &tm.kv_hi, bar_qk, head * ATTN_DH, t0);
    tma_load_2d(base + kt, &tm.kv_lo, bar_qk, head * ATTN_DH, t0);
    tma_load_2d(base + 2 * kt, &tm.kv_hi, bar_qk, h + head * ATTN_DH, t0);
    tma_load_2d(base + 3 * kt, &tm.kv_lo, bar_qk, h + head * ATTN_DH, t0);
    mbar_arrive_expect_tx(bar_v, 2 * kt);
    tma_load_2d(base + 4 * kt, &tm.kv_hi, bar_v, 2 * h + head * ATTN_DH, t0);
    tma_load_2d(base + 5 * kt, &tm.kv_lo, bar_v, 2 * h + head * ATTN_DH, t0);
  };

  // GROUPED: (t0, L, j0) of a unit — first token and length of its group, first column of row tid's caption inside it —
  // are fetched one unit ahead (two dependent global loads: read at the top of the unit they stall warp 0, which also
  // issues the MMAs)
  int t0 = 0, L = 0, j0 = 0;
  if (static_cast<int>(blockIdx.x) < n_units) {
    const int seq = blockIdx.x / heads;
    t0 = offs[seq];
    L = offs[seq + 1] - t0;
    if (L > lp) L = lp;
    if (GROUPED) j0 = tid < L ? tok_start[t0 + tid] - t0 : 0;
    if (tid == 0) issue_loads(blockIdx.x, t0);
  }
  int it = 0;
  for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x, ++it) {
    const uint32_t phase = it & 1;
    const int head = unit % heads;
    const int nunit = unit + static_cast<int>(gridDim.x);
    if (!GROUPED) {
      const int seq = unit / heads;
      t0 = offs[seq];
      L = offs[seq + 1] - t0;
      if (L > lp) L = lp;
    }
    // ---- S = Q K^T
    if (tid == 0) {
      mbar_wait(bar_qk, phase, 11);
      tc_fence_after();
      const uint64_t dq_hi = make_desc_k128(s_q), dq_lo = make_desc_k128(s_q + kt);
      const uint64_t dk_hi = make_desc_k128(s_k), dk_lo = make_desc_k128(s_k + kt);
#pragma unroll
      for (int k = 0; k < ATTN_DH / 16; ++k) {
        const uint64_t ko = static_cast<uint64_t>(k * 2);
        tc_mma_f16(tmem_base, dq_lo + ko, dk_hi + ko, idesc1, k > 0);
        tc_mma_f16(tmem_base, dq_hi + ko, dk_lo + ko, idesc1, 1);
        tc_mma_f16(tmem_base, dq_hi + ko, dk_hi + ko, idesc1, 1);
      }
      tc_commit(bar_s);
    }
    mbar_wait(bar_s, phase, 12);
    tc_fence_after();
    int nt0 = 0, nt1 = 0;                          // GROUPED: the next unit's group, in flight during the softmax
    if (GROUPED && nunit < n_units) { const int ns = nunit / heads; nt0 = offs[ns]; nt1 = offs[ns + 1]; }
    // ---- causal softmax of row i = tid (TMEM lane): e = exp(s - max) goes to the swizzled P tiles as fp16 hi/lo
    // planes (unnormalised, in (0, 1]); the 1/sum factor is applied to the output row instead
    float inv = 0.f;
    if (warp * 32 < lp) {                          // warp-uniform; rows >= lp have no P row in the lp-row tiles
      const int i = tid;
      const bool live = i < L;
      // GROUPED: column j of row i is visible iff j0 <= j <= i — one unsigned compare, (j - j0) <= (i - j0); dead rows
      // see nothing.  Otherwise j0 = 0 and the compare folds to j <= i.
      const int jb0 = GROUPED ? (live ? j0 : (1 << 20)) : 0;
      const unsigned span = GROUPED ? (live ? static_cast<unsigned>(i - j0) : 0u) : static_cast<unsigned>(i);
      float v[NC][16];
#pragma unroll
      for (int c = 0; c < NC; ++c)
        if (c * 16 < lp) tmem_ld_32x16(tmem_base + lane_addr + c * 16, v[c]);   // uniform predicate
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int u = 0; u < 16; ++u)
          if (c * 16 < lp && (GROUPED || live) && static_cast<unsigned>(c * 16 + u - jb0) <= span) mx = fmaxf(mx, v[c][u]);
      mx *= scale;                                 // scale > 0: max(scale * s) = scale * max(s)
      float sum = 0.f;
      const uint32_t prow = s_p + i * 128;
      const int sw = i & 7;
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        if (c * 16 < lp) {
          uint32_t hh[8], ll[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int ja = c * 16 + 2 * u, jb = ja + 1;
            const float pa = ((GROUPED || live) && static_cast<unsigned>(ja - jb0) <= span) ? __expf(fmaf(v[c][2 * u], scale, -mx)) : 0.f;
            const float pb = ((GROUPED || live) && static_cast<unsigned>(jb - jb0) <= span) ? __expf(fmaf(v[c][2 * u + 1], scale, -mx)) : 0.f;
            sum += pa + pb;
            split_f16x2(pa, pb, hh[u], ll[u]);
          }
          // 16 columns = two 16-byte chunks of k-block c / 4
          const int kb = c >> 2, ch = (c & 3) * 2;
          const uint32_t th = prow + kb * kt, tl = prow + (2 + kb) * kt;
          if (i < lp) {
            sts_v4(th + (((ch) ^ sw) << 4), hh[0], hh[1], hh[2], hh[3]);
            sts_v4(th + (((ch + 1) ^ sw) << 4), hh[4], hh[5], hh[6], hh[7]);
            sts_v4(tl + (((ch) ^ sw) << 4), ll[0], ll[1], ll[2], ll[3]);
            sts_v4(tl + (((ch + 1) ^ sw) << 4), ll[4], ll[5], ll[6], ll[7]);
          }
        }
      }
      inv = live ? 1.0f / sum : 0.f;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    int nL = nt1 - nt0;
    if (nL > lp) nL = lp;
    const int nj0 = (GROUPED && tid < nL) ? tok_start[nt0 + tid] - nt0 : 0;   // in flight during P V and the output stores
    // ---- O = P V
    if (tid == 0) {
      tc_fence_after();
      mbar_wait(bar_v, phase, 13);
      tc_fence_after();
      const uint32_t tmem_o = tmem_base;
      for (int k = 0; k < ksteps2; ++k) {
        const int kb = k >> 2;
        const uint64_t ko = static_cast<uint64_t>((k & 3) * 2);
        const uint64_t dp_hi = make_desc_k128(s_p + kb * kt) + ko, dp_lo = make_desc_k128(s_p + (2 + kb) * kt) + ko;
        // V tile rows are tokens (= K of this product): one k-step is 16 rows = 2048 bytes further down the tile;
        // MN-major SW128 canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -> SBO = 1024 B, as written by TMA
        const uint64_t dv_hi = make_desc_k128(s_v + k * 2048), dv_lo = make_desc_k128(s_v + kt + k * 2048);
        tc_mma_f16(tmem_o, dp_lo, dv_hi, idesc2, k > 0);
        tc_mma_f16(tmem_o, dp_hi, dv_lo, idesc2, 1);
        tc_mma_f16(tmem_o, dp_hi, dv_hi, idesc2, 1);
      }
      tc_commit(bar_o);
    }
    mbar_wait(bar_o, phase, 14);
    tc_fence_after();
    // the operands are free again: fetch the next unit while this one's output is written
    if (tid == 0 && nunit < n_units) issue_loads(nunit, GROUPED ? nt0 : offs[nunit / heads]);
    if (warp * 32 < L) {                           // warp-uniform: tcgen05.ld is warp-collective
      uint16_t* oh = o_hi + static_cast<long long>(t0 + tid) * ldo + head * ATTN_DH;
      uint16_t* ol = o_lo + static_cast<long long>(t0 + tid) * ldo + head * ATTN_DH;
#pragma unroll
      for (int c0 = 0; c0 < ATTN_DH; c0 += 16) {
        float v[16];
        tmem_ld_32x16(tmem_base + lane_addr + c0, v);
        uint32_t hh[8], ll[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          split_f16x2(v[2 * u] * inv, v[2 * u + 1] * inv, hh[u], ll[u]);
        }
        if (tid < L) {
          *reinterpret_cast<uint4*>(oh + c0) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
          *reinterpret_cast<uint4*>(oh + c0 + 8) = make_uint4(hh[4], hh[5], hh[6], hh[7]);
          *reinterpret_cast<uint4*>(ol + c0) = make_uint4(ll[0], ll[1], ll[2], ll[3]);
          *reinterpret_cast<uint4*>(ol + c0 + 8) = make_uint4(ll[4], ll[5], ll[6], ll[7]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // all TMEM reads of this unit are done before the next unit's MMAs overwrite S / O
    tc_fence_after();
    if (GROUPED) { t0 = nt0; L = nL; j0 = nj0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 128);
}

}  // namespace emcid

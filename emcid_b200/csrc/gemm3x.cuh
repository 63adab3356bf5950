// emcid_b200 — generic 3xTF32 "NT" GEMM on tcgen05/TMEM fed by TMA (sm_100a).
//
//   D[M x N] = sum_k A[m, k] * B[n, k]          A: [M x K] row-major, B: [N x K] row-major
//
// Both operands are K-major and arrive pre-split into two tf32-exact planes (hi, lo) so that
//   a*b ~= a_hi*b_hi + a_lo*b_hi + a_hi*b_lo      (fp32-class accuracy, 3 tensor-core MMAs)
// with fp32 accumulation in TMEM.  Warp-specialised, persistent:
//   warp 0   : TMA producer  (one elected lane)     global -> 128B-swizzled smem ring
//   warp 1   : MMA issuer    (one elected lane)     tcgen05.mma.kind::tf32, commit -> mbarriers
//   warp 2   : TMEM allocator
//   warps 4-7: epilogue      TMEM -> registers -> global (mode-specific, see Epi*)
// Two accumulator stages of BLOCK_N columns each live in TMEM so the epilogue of unit i
// overlaps the mainloop of unit i+1.
//
// Every hot GEMM-shaped op of the EMCID path instantiates this one kernel:
//   fc1      : A = W1 planes [d x h],  B = X planes [T x h]      -> act/mask/split epilogue
//   SYRK     : A = B = A^T planes [d x T_slab], lower tiles, stream-K over tokens -> red.add
//   K K^T, Cholesky trailing updates, TRSM block products         -> load/scale/store epilogue
// It replaces the reference's `a.t().mm(a)` (util/runningstats.py:493) and the GEMM-shaped
// parts of `torch.linalg.solve` / `@` in emcid/emcid_main.py:1045-1050.
#pragma once

#include "common.cuh"

namespace emcid {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 32;  // 32 fp32 = 128 bytes = one swizzle row
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_ROW_BYTES = 128;
constexpr int GEMM_A_PLANE_BYTES = GEMM_BLOCK_M * GEMM_ROW_BYTES;  // 16 KB

enum EpiMode : int {
  EPI_STORE = 0,   // C = alpha*acc + beta*C                     (unique owner per tile)
  EPI_RED = 1,     // C += acc via red.global.add                (stream-K safe)
  EPI_FC1 = 2,     // planes = split(mask(act(acc + bias[row])))  (fc1 -> A^T slab)
  EPI_PLANES = 3,  // C = alpha*acc + beta*C, and planes = split(C) (solver operands)
};

enum ActMode : int { ACT_QUICK_GELU = 0, ACT_GELU_ERF = 1, ACT_NONE = 2 };

struct GemmParams {
  int M, N, K;
  int lower;            // 1: only tiles that intersect the lower triangle (row >= col)
  int streamk;          // 1: split the flattened (tile, k-block) space evenly over CTAs
  int chunk_kblocks;    // max k-blocks accumulated in TMEM before an epilogue (stream-K)
  const int* dyn_n;     // optional device scalar overriding N (fc1: number of valid tokens)
  const int* dyn_k;     // optional device scalar overriding K (SYRK: number of valid tokens)
  // epilogue
  float* C;
  long long ldc;
  float alpha, beta;
  float* P_hi;
  float* P_lo;
  long long ldp;
  const float* bias;    // [M]
  int act;
  int trans_out;        // EPI_PLANES: also write C^T into Ct (used to keep L^T beside L)
  float* Ct;
  long long ldct;
};

template <int BLOCK_N, int STAGES>
struct GemmCfg {
  static constexpr int kBlockN = BLOCK_N;
  static constexpr int kStages = STAGES;
  static constexpr int kBPlaneBytes = BLOCK_N * GEMM_ROW_BYTES;
  static constexpr int kStageBytes = 2 * GEMM_A_PLANE_BYTES + 2 * kBPlaneBytes;
  static constexpr int kTmemCols = 2 * BLOCK_N;  // two accumulator stages (power of two)
  static constexpr int kSmemBytes = STAGES * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct Unit {
  int m0, n0;    // tile origin (rows of A / rows of B)
  int kb0, kb1;  // k-block range [kb0, kb1)
};

// Deterministic work enumeration, evaluated identically by the producer, MMA and epilogue roles.
struct Sched {
  int m_tiles, n_tiles, kb_tile, R, lower, streamk, chunk, block_n;
  long long pos, end;          // stream-K: position in the flattened (tile, kb) space
  int tile, tile_step, num_tiles;

  __device__ void init(const GemmParams& p, int block_n, int M, int N, int K) {
    this->block_n = block_n;
    m_tiles = (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
    n_tiles = (N + block_n - 1) / block_n;
    kb_tile = (K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;
    R = block_n / GEMM_BLOCK_M;
    lower = p.lower;
    streamk = p.streamk;
    chunk = p.chunk_kblocks > 0 ? p.chunk_kblocks : (1 << 30);
    if (lower) {
      num_tiles = 0;
      for (int j = 0; j < n_tiles; ++j) {
        int c = m_tiles - j * R;
        if (c > 0) num_tiles += c;
      }
    } else {
      num_tiles = m_tiles * n_tiles;
    }
    if (kb_tile == 0) num_tiles = 0;
    if (streamk) {
      long long total = static_cast<long long>(num_tiles) * kb_tile;
      long long per = (total + gridDim.x - 1) / gridDim.x;
      pos = per * blockIdx.x;
      end = pos + per < total ? pos + per : total;
    } else {
      tile = blockIdx.x;
      tile_step = gridDim.x;
    }
  }

  __device__ void tile_origin(int t, int& m0, int& n0) const {
    if (lower) {
      int j = 0;
      for (;; ++j) {
        int c = m_tiles - j * R;
        if (t < c) break;
        t -= c;
      }
      m0 = (j * R + t) * GEMM_BLOCK_M;
      n0 = j * block_n;
    } else {
      m0 = (t % m_tiles) * GEMM_BLOCK_M;
      n0 = (t / m_tiles) * block_n;
    }
  }

  __device__ bool next(Unit& u) {
    if (streamk) {
      if (pos >= end) return false;
      int t = static_cast<int>(pos / kb_tile);
      int kb = static_cast<int>(pos - static_cast<long long>(t) * kb_tile);
      long long len = end - pos;
      if (len > kb_tile - kb) len = kb_tile - kb;
      if (len > chunk) len = chunk;
      tile_origin(t, u.m0, u.n0);
      u.kb0 = kb;
      u.kb1 = kb + static_cast<int>(len);
      pos += len;
      return true;
    }
    if (tile >= num_tiles) return false;
    tile_origin(tile, u.m0, u.n0);
    u.kb0 = 0;
    u.kb1 = kb_tile;
    tile += tile_step;
    return true;
  }
};

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_QUICK_GELU) {
    // HF QuickGELUActivation: x * sigmoid(1.702 x)   (transformers/activations.py)
    return v / (1.0f + expf(-1.702f * v));
  } else if (act == ACT_GELU_ERF) {
    return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  }
  return v;
}

template <int BLOCK_N, int STAGES, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm3x_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
              const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
              const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned tiles.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA_hi);
    tma_prefetch_desc(&tmA_lo);
    tma_prefetch_desc(&tmB_hi);
    tma_prefetch_desc(&tmB_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  int N = p.N, K = p.K;
  if (p.dyn_n) { int v = *p.dyn_n; N = v < N ? v : N; }
  if (p.dyn_k) { int v = *p.dyn_k; K = v < K ? v : K; }

  Sched sched;
  sched.init(p, BLOCK_N, p.M, N, K);
  Unit u;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      while (sched.next(u)) {
        for (int kb = u.kb0; kb < u.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 1);
          uint8_t* st = smem + stage * Cfg::kStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          const int kc = kb * GEMM_BLOCK_K;
          tma_load_2d(st, &tmA_hi, &full_bar[stage], kc, u.m0);
          tma_load_2d(st + GEMM_A_PLANE_BYTES, &tmA_lo, &full_bar[stage], kc, u.m0);
          uint8_t* sb = st + 2 * GEMM_A_PLANE_BYTES;
#pragma unroll
          for (int r = 0; r < BLOCK_N / 128; ++r) {
            tma_load_2d(sb + r * GEMM_A_PLANE_BYTES, &tmB_hi, &full_bar[stage], kc, u.n0 + r * 128);
            tma_load_2d(sb + Cfg::kBPlaneBytes + r * GEMM_A_PLANE_BYTES, &tmB_lo, &full_bar[stage],
                        kc, u.n0 + r * 128);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(GEMM_BLOCK_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      while (sched.next(u)) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 2);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        uint32_t accumulate = 0;
        for (int kb = u.kb0; kb < u.kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase, 3);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t da_hi = make_desc_k128(sa);
          const uint64_t da_lo = make_desc_k128(sa + GEMM_A_PLANE_BYTES);
          const uint64_t db_hi = make_desc_k128(sa + 2 * GEMM_A_PLANE_BYTES);
          const uint64_t db_lo = make_desc_k128(sa + 2 * GEMM_A_PLANE_BYTES + Cfg::kBPlaneBytes);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / 8; ++k) {
            // +32 bytes along K inside the swizzle row = +2 in the (addr >> 4) field.
            const uint64_t koff = static_cast<uint64_t>(k * 2);
            tc_mma_tf32(tmem_d, da_lo + koff, db_hi + koff, idesc, accumulate);
            tc_mma_tf32(tmem_d, da_hi + koff, db_lo + koff, idesc, 1);
            tc_mma_tf32(tmem_d, da_hi + koff, db_hi + koff, idesc, 1);
            accumulate = 1;
          }
          tc_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tfull_bar[acc]);      // accumulator complete -> epilogue
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    while (sched.next(u)) {
      mbar_wait(&tfull_bar[acc], acc_phase, 4);
      tc_fence_after();
      const int row = u.m0 + q * 32 + lane;
      const bool row_ok = row < p.M;
      const uint32_t taddr = tmem_base + acc * BLOCK_N + (static_cast<uint32_t>(q * 32) << 16);
      float bias = 0.f;
      if (EPI == EPI_FC1 && row_ok && p.bias) bias = p.bias[row];
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        float v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        const int col0 = u.n0 + c * 32;
        if (!row_ok) continue;
        if (EPI == EPI_RED) {
          float* dst = p.C + static_cast<long long>(row) * p.ldc + col0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (col0 + 4 * j < p.N) red_add_v4(dst + 4 * j, v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          }
        } else if (EPI == EPI_STORE || EPI == EPI_PLANES) {
          float* dst = p.C + static_cast<long long>(row) * p.ldc + col0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = col0 + 4 * j;
            if (col < p.N) {
              float4 o;
              o.x = p.alpha * v[4 * j];
              o.y = p.alpha * v[4 * j + 1];
              o.z = p.alpha * v[4 * j + 2];
              o.w = p.alpha * v[4 * j + 3];
              if (p.beta != 0.f) {
                const float4 old = *reinterpret_cast<const float4*>(dst + 4 * j);
                o.x += p.beta * old.x; o.y += p.beta * old.y; o.z += p.beta * old.z; o.w += p.beta * old.w;
              }
              *reinterpret_cast<float4*>(dst + 4 * j) = o;
              if (EPI == EPI_PLANES) {
                float4 h, l;
                split_tf32(o.x, h.x, l.x); split_tf32(o.y, h.y, l.y);
                split_tf32(o.z, h.z, l.z); split_tf32(o.w, h.w, l.w);
                const long long off = static_cast<long long>(row) * p.ldp + col;
                *reinterpret_cast<float4*>(p.P_hi + off) = h;
                *reinterpret_cast<float4*>(p.P_lo + off) = l;
                if (p.trans_out) {
                  p.Ct[static_cast<long long>(col) * p.ldct + row] = o.x;
                  p.Ct[static_cast<long long>(col + 1) * p.ldct + row] = o.y;
                  p.Ct[static_cast<long long>(col + 2) * p.ldct + row] = o.z;
                  p.Ct[static_cast<long long>(col + 3) * p.ldct + row] = o.w;
                }
              }
            }
          }
        } else if (EPI == EPI_FC1) {
          // acc[row = feature i][col = token t]; tokens >= N (dynamic valid count) are padding
          // rows of the compacted X slab and must contribute exactly zero to mom2.
          float* dh = p.P_hi + static_cast<long long>(row) * p.ldp + col0;
          float* dl = p.P_lo + static_cast<long long>(row) * p.ldp + col0;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 h, l;
            float a0 = (col0 + 4 * j + 0 < N) ? apply_act(v[4 * j + 0] + bias, p.act) : 0.f;
            float a1 = (col0 + 4 * j + 1 < N) ? apply_act(v[4 * j + 1] + bias, p.act) : 0.f;
            float a2 = (col0 + 4 * j + 2 < N) ? apply_act(v[4 * j + 2] + bias, p.act) : 0.f;
            float a3 = (col0 + 4 * j + 3 < N) ? apply_act(v[4 * j + 3] + bias, p.act) : 0.f;
            split_tf32(a0, h.x, l.x); split_tf32(a1, h.y, l.y);
            split_tf32(a2, h.z, l.z); split_tf32(a3, h.w, l.w);
            if (col0 + 4 * j < p.ldp) {
              *reinterpret_cast<float4*>(dh + 4 * j) = h;
              *reinterpret_cast<float4*>(dl + 4 * j) = l;
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

}  // namespace emcid

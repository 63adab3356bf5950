// emcid_b200 — generic 3-term-split "NT" GEMM on tcgen05/TMEM fed by TMA (sm_100a); 3xTF32 or 3xFP16 operand planes.
//
//   D[M x N] = sum_k A[m, k] * B[n, k]          A: [M x K] row-major, B: [N x K] row-major
//
// Both operands are K-major and arrive pre-split into two tf32-exact planes (hi, lo) so that
//   a*b ~= a_lo*b_hi + a_hi*b_lo + a_hi*b_hi      (fp32-class accuracy, 3 tensor-core MMAs).
//
// Accumulation discipline (measured on B200, profiles/round1/r01_probe_gemm_v0_tmem_fullK.json): the
// tensor core adds into its fp32 TMEM accumulator with round-toward-zero, which shrinks a running
// sum by ~1e-8 per contracted element — 1e-5 after a thousand tokens, i.e. the whole mom2 error
// budget.  So TMEM only ever holds a short CHUNK (chunk_kblocks x 32 contracted elements); the
// epilogue warps pull every finished chunk out with tcgen05.ld and add it into register-resident
// running sums with round-to-nearest fp32 adds while the tensor core works on the next chunk in the
// other accumulator stage.  Only at the end of a work unit do the sums go to global memory.
//
// Warp roles (384 threads, persistent, one CTA per SM):
//   warp 0    : TMA producer  (one lane)   global -> 128B-swizzled smem ring
//   warp 1    : MMA issuer    (one lane)   tcgen05.mma.kind::tf32, tcgen05.commit -> mbarriers
//   warp 2    : TMEM allocator
//   warps 4-11: epilogue; warp w owns TMEM lanes 32*(w%4).. and column half (w-4)/4
// 384 threads x 168 registers fill the register file; the epilogue's 128 running sums fit.
//
// Every hot GEMM-shaped op of the EMCID path instantiates this one kernel:
//   linear layers of the text encoder (clip.cuh): A = activation planes [T x K], B = weight planes [N x K], 3xFP16,
//              CTA pairs (CTA2 = 1: cta_group::2, 256-row tiles), tiles walked along N, TMA-staged epilogue
//   SYRK     : A = B = act(fc1) planes [T x d] read as MN-major tiles (KIND_F16_MN), lower pair tiles, hybrid schedule
//              (whole tiles march through the tokens together, leftover tiles stream-K'd), red.add epilogue
//   fc1 (hook mode): A = W1 planes [d x h], B = X planes [T x h] -> bias/act/mask/split epilogue (A^T slab)
//   K K^T + lambda C, Cholesky panel/trailing updates, triangular inverse and its applications (k_tri) -> generic epilogue
// It replaces the reference's `a.t().mm(a)` (util/runningstats.py:493) and the GEMM-shaped parts
// of `torch.linalg.solve` / `@` in emcid/emcid_main.py:1045-1050.
#pragma once

#include "common.cuh"

namespace emcid {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 32;  // tf32 kind: 32 fp32 = 128 bytes = one swizzle row

// Operand kind of the 3-term split.  KIND_TF32: planes are fp32 containers of tf32-exact values,
// tcgen05.mma.kind::tf32 (K = 8 per instruction).  KIND_F16: planes are 16-bit (hi = fp16, lo = bf16
// or fp16, see split_f16), tcgen05.mma.kind::f16 (K = 16 per instruction): the same three MMAs per
// 128-byte k-step retire twice the contraction depth and every plane moves half the bytes.
// (A variant staged as 64-byte rows with four pipeline stages was measured within 2 % of this one — the mainloop is not
// TMA-latency bound — and dropped.)
// KIND_F16_MN: fp16 planes stored [K x MN] row-major (MN contiguous) — the SYRK reads act(fc1) planes [tokens x
// features] directly: both operands are MN-major tiles of 64 tokens (K) x 64-feature atoms, no transposed copy needed.
enum GemmKind : int { KIND_TF32 = 0, KIND_F16 = 1, KIND_F16_MN = 3 };
template <int KIND> struct KindTraits;
template <> struct KindTraits<KIND_TF32> { static constexpr int kBlockK = 32, kElemBytes = 4, kRowBytes = 128; };
template <> struct KindTraits<KIND_F16> { static constexpr int kBlockK = 64, kElemBytes = 2, kRowBytes = 128; };
template <> struct KindTraits<KIND_F16_MN> { static constexpr int kBlockK = 64, kElemBytes = 2, kRowBytes = 128; };
constexpr int GEMM_THREADS = 384;
constexpr int GEMM_EPI_THREADS = 256;
constexpr int GEMM_ROW_BYTES = 128;
constexpr int GEMM_A_PLANE_BYTES = GEMM_BLOCK_M * GEMM_ROW_BYTES;  // 16 KB
constexpr int GEMM_DEFAULT_CHUNK = 2;                               // k-blocks per TMEM chunk
#ifndef EMCID_FOLD_WIDTH
#define EMCID_FOLD_WIDTH 32   // accumulator columns per TMEM load batch in the staged-epilogue kernels (16, 32 or 64)
#endif

enum EpiMode : int {
  EPI_GENERIC = 0,  // C = alpha*acc + beta*Cin ; optional planes / transposed planes of the result
  EPI_RED = 1,      // C += acc via red.global.add (stream-K safe)
  EPI_FC1 = 2,      // planes = split(mask(act(acc + bias[row])))  (fc1 -> A^T slab)
  EPI_LINEAR = 3,   // v = act(alpha*acc + bias_col[col]) + beta*Cin ; C / planes / transposed planes of v
  EPI_LINEAR_TMA = 4,  // same result, staged through 32 KB of shared memory and written by TMA bulk stores
                       // (residual by TMA reduce-add): coalesced, asynchronous, ragged tiles clipped by hardware
};

// Output tensor maps of EPI_LINEAR_TMA (unused by the other epilogues).
struct GemmOutMaps {
  CUtensorMap c;     // fp32 result [M x N] (box 16 x 128 rows, 64B swizzle)  or  hi plane [M x N] (box 32 x 128 rows, 64B swizzle)
  CUtensorMap c2;    // lo plane
};
constexpr int GEMM_STAGING_BYTES = 32768;

enum ActMode : int { ACT_QUICK_GELU = 0, ACT_GELU_ERF = 1, ACT_NONE = 2 };

struct GemmParams {
  int M, N, K;                 // extents of this product
  int a_row0, a_col0;          // origin of A inside the tensor its maps describe (row, k)
  int b_row0, b_col0;          // origin of B
  int a_batch_rows;            // blockIdx.y * a_batch_rows is added to A rows (stacked batches)
  int b_batch_rows;
  int lower;                   // 1: only tiles that intersect the lower triangle (row >= col)
  int streamk;                 // 1: split the flattened (tile, k-block) space evenly over CTAs
                               // 2: hybrid — as many whole tiles as divide evenly among the CTAs (each over the full K,
                               //    all CTAs marching through K together: operand reuse in L2, ONE red.add per tile), then
                               //    the leftover tiles stream-K'd; needs the red.add epilogue like 1
  int n_fastest;               // 1: consecutive tiles walk along N (the tiles of one wave share A row blocks: the big
                               //    operand A is then read from HBM once, not once per N tile); 0: along M
  int k_tri;                   // triangular operand: skip the k-blocks that are known zeros.  1: B is lower triangular
                               // (k < n0 + BLOCK_N); 2: B is upper triangular (k >= n0); 3: A is lower triangular
                               // (k < m0 + BLOCK_M); 4: A is upper triangular (k >= m0).  Tile-level only (not with
                               // stream-K).
  int chunk_kblocks;           // k-blocks accumulated in TMEM between register folds
  const int* dyn_n;            // optional device scalar overriding N (fc1: number of valid tokens)
  const int* dyn_k;            // optional device scalar overriding K (SYRK: number of valid tokens)
  // epilogue
  float alpha, beta;
  const float* Cin;            // addend (may alias C), nullptr when beta == 0
  long long ldcin, cin_batch;
  float* C;                    // fp32 result (may be nullptr when only planes are wanted)
  long long ldc, c_batch;
  float* P_hi;                 // tf32 planes of the result, same orientation as C
  float* P_lo;
  long long ldp, p_batch;
  float* Pt_hi;                // tf32 planes of the transposed result: Pt[col][row]
  float* Pt_lo;
  long long ldpt, pt_batch;
  const float* bias;           // [M], EPI_FC1
  const float* bias_col;       // [N], EPI_LINEAR (16-byte aligned)
  int act;
  int lo_fmt;                  // KIND_F16: format of the lo planes (FMT_BF16 default, FMT_F16)
  long long* count;            // optional device counter: thread 0 of CTA (0, 0) adds count_add (the token count of the
  long long count_add;         // SYRK's block: instead of a 1-thread kernel of its own)
};

template <int BLOCK_N, int STAGES, int ROWB = GEMM_ROW_BYTES>
struct GemmCfg {
  static constexpr int kBlockN = BLOCK_N;
  static constexpr int kStages = STAGES;
  static constexpr int kAPlaneBytes = GEMM_BLOCK_M * ROWB;
  static constexpr int kBPlaneBytes = BLOCK_N * ROWB;
  static constexpr int kStageBytes = 2 * kAPlaneBytes + 2 * kBPlaneBytes;
  static constexpr int kTmemCols = 2 * BLOCK_N;  // two chunk accumulators (power of two)
  static constexpr int kSmemBytes = STAGES * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int kSmemBytesStaged = kSmemBytes + GEMM_STAGING_BYTES;
};

struct Unit {
  int m0, n0;    // tile origin (rows of A / rows of B), local to this product
  int kb0, kb1;  // k-block range [kb0, kb1)
};

// Deterministic work enumeration, evaluated identically by the producer, MMA and epilogue roles.
struct Sched {
  int m_tiles, n_tiles, kb_tile, R, lower, streamk, block_n, block_m, n_fastest, k_tri, block_k;
  long long pos, end;          // stream-K: position in the flattened (tile, kb) space
  int tile, tile_step, num_tiles;
  int whole_end;               // hybrid: tiles [0, whole_end) are processed whole, [whole_end, num_tiles) are stream-K'd

  // bid / nb: index and number of the scheduling entities (CTAs, or CTA pairs with 256-row tiles)
  __device__ void init(const GemmParams& p, int block_n_, int block_k_, int M, int N, int K, int block_m_ = GEMM_BLOCK_M,
                       int bid = blockIdx.x, int nb = gridDim.x) {
    block_n = block_n_;
    block_m = block_m_;
    m_tiles = (M + block_m - 1) / block_m;
    n_tiles = (N + block_n - 1) / block_n;
    block_k = block_k_;
    kb_tile = (K + block_k - 1) / block_k;
    R = block_n / block_m;
    lower = p.lower;
    streamk = p.streamk;
    n_fastest = p.n_fastest;
    k_tri = p.k_tri;
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
    whole_end = streamk == 2 ? (num_tiles / nb) * nb : (streamk ? 0 : num_tiles);
    tile = bid;
    tile_step = nb;
    if (streamk) {
      long long total = static_cast<long long>(num_tiles - whole_end) * kb_tile;
      long long per = (total + nb - 1) / nb;
      pos = per * bid;
      end = pos + per < total ? pos + per : total;
    } else {
      pos = end = 0;
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
      m0 = (j * R + t) * block_m;
      n0 = j * block_n;
    } else if (n_fastest) {
      m0 = (t / n_tiles) * block_m;
      n0 = (t % n_tiles) * block_n;
    } else {
      m0 = (t % m_tiles) * block_m;
      n0 = (t / m_tiles) * block_n;
    }
  }

  __device__ bool next(Unit& u) {
    if (tile >= whole_end) {
      if (pos >= end) return false;
      int t = static_cast<int>(pos / kb_tile);
      int kb = static_cast<int>(pos - static_cast<long long>(t) * kb_tile);
      long long len = end - pos;
      if (len > kb_tile - kb) len = kb_tile - kb;
      tile_origin(whole_end + t, u.m0, u.n0);
      u.kb0 = kb;
      u.kb1 = kb + static_cast<int>(len);
      pos += len;
      return true;
    }
    tile_origin(tile, u.m0, u.n0);
    u.kb0 = 0;
    u.kb1 = kb_tile;
    if (k_tri == 1) { const int e = (u.n0 + block_n + block_k - 1) / block_k; if (e < u.kb1) u.kb1 = e; }
    else if (k_tri == 2) { u.kb0 = u.n0 / block_k; }
    else if (k_tri == 3) { const int e = (u.m0 + block_m + block_k - 1) / block_k; if (e < u.kb1) u.kb1 = e; }
    else if (k_tri == 4) { u.kb0 = u.m0 / block_k; }
    tile += tile_step;
    return true;
  }
};

// Compile-time epilogue options (template parameter EFLAGS of gemm3x_kernel; EPI_LINEAR / EPI_FC1 only).
// The epilogue is fully unrolled over the 128 accumulator columns a thread owns, so every option that is
// not compiled out costs 32 copies of its code: with run-time options the kernel grew to 260 KB of SASS and
// stalled on instruction fetch (measured: 48% of warp samples "no instruction", 12% tensor-pipe active).
constexpr int EF_ACT_MASK = 3;   // bits 0-1: ActMode
constexpr int EF_C = 4;          // store the fp32 result
constexpr int EF_CIN = 8;        // add the fp32 residual
constexpr int EF_P = 16;         // store split planes of the result
constexpr int EF_PT = 32;        // store split planes of the transposed result
constexpr int EF_DEFAULT = ACT_NONE;

// Activation with the option fixed at compile time.  quick_gelu uses the MUFU fast paths (ex2.approx / rcp):
// relative error ~2e-7, unbiased, an order of magnitude below the 3-term split's own rounding.
template <int ACT>
__device__ __forceinline__ float act_ct(float v) {
  if (ACT == ACT_QUICK_GELU) return __fdividef(v, 1.0f + __expf(-1.702f * v));
  if (ACT == ACT_GELU_ERF) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  return v;
}

template <int N> __device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N> __device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// Stores 4 consecutive result values as split planes (tf32: two float4; f16: two 8-byte words).
template <int KIND>
__device__ __forceinline__ void store_planes4(float* hi_base, float* lo_base, long long idx, float4 o, int lo_fmt) {
  if (KIND == KIND_TF32) {
    float4 h, l;
    split_tf32(o.x, h.x, l.x); split_tf32(o.y, h.y, l.y);
    split_tf32(o.z, h.z, l.z); split_tf32(o.w, h.w, l.w);
    *reinterpret_cast<float4*>(hi_base + idx) = h;
    *reinterpret_cast<float4*>(lo_base + idx) = l;
  } else {
    uint16_t h[4], l[4];
    split_f16(o.x, lo_fmt, h[0], l[0]); split_f16(o.y, lo_fmt, h[1], l[1]);
    split_f16(o.z, lo_fmt, h[2], l[2]); split_f16(o.w, lo_fmt, h[3], l[3]);
    uint2 hv, lv;
    hv.x = h[0] | (static_cast<uint32_t>(h[1]) << 16); hv.y = h[2] | (static_cast<uint32_t>(h[3]) << 16);
    lv.x = l[0] | (static_cast<uint32_t>(l[1]) << 16); lv.y = l[2] | (static_cast<uint32_t>(l[3]) << 16);
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(hi_base) + idx) = hv;
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(lo_base) + idx) = lv;
  }
}

template <int KIND>
__device__ __forceinline__ void store_plane1(float* hi_base, float* lo_base, long long idx, float v, int lo_fmt) {
  if (KIND == KIND_TF32) {
    float h, l;
    split_tf32(v, h, l);
    hi_base[idx] = h; lo_base[idx] = l;
  } else {
    uint16_t h, l;
    split_f16(v, lo_fmt, h, l);
    reinterpret_cast<uint16_t*>(hi_base)[idx] = h;
    reinterpret_cast<uint16_t*>(lo_base)[idx] = l;
  }
}

// CTA2 = 1: the kernel is launched in clusters of two CTAs that form one cta_group::2 pair per 256 x BLOCK_N tile
// (see common.cuh).  Rank 0 issues the MMAs for both; both produce (their own A rows and their half of B), both
// run the epilogue on their own 128 accumulator rows.
template <int BLOCK_N, int STAGES, int EPI, int KIND = KIND_TF32, int EFLAGS = EF_DEFAULT, int CTA2 = 0>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm3x_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
              const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
              const __grid_constant__ GemmOutMaps om, const GemmParams p) {
  constexpr int ROWB = KindTraits<KIND>::kRowBytes;
  using Cfg = GemmCfg<BLOCK_N, STAGES, ROWB>;
  constexpr int COLS = BLOCK_N / 2;  // accumulator columns owned by one epilogue thread
  constexpr int BLOCK_K = KindTraits<KIND>::kBlockK;  // elements per k-block (one swizzle row)
  constexpr int A_PLANE = Cfg::kAPlaneBytes;
  constexpr int B_ROWS = CTA2 ? BLOCK_N / 2 : BLOCK_N;           // B rows staged by this CTA
  constexpr int B_PLANE = B_ROWS * ROWB;
  constexpr int STAGE_BYTES = 2 * A_PLANE + 2 * B_PLANE;
  const uint32_t rank = CTA2 ? cluster_ctarank() : 0u;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned tiles.
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint8_t* stg = smem + STAGES * STAGE_BYTES;  // epilogue staging (EPI_LINEAR_TMA only), 1024-byte aligned
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stg + (EPI == EPI_LINEAR_TMA ? GEMM_STAGING_BYTES : 0));
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  // EPI_LINEAR_TMA: the tile's 256 bias values, one 128-float row per column half (1 KB behind the 256-byte barrier block)
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);

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
      // pair mode: one elected lane per epilogue warp of BOTH CTAs arrives on the leader's barrier
      mbar_init(&tempty_bar[a], CTA2 ? 2 * (GEMM_EPI_THREADS / 32) : GEMM_EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (CTA2) { tmem_alloc2(tmem_slot, Cfg::kTmemCols); tmem_relinquish2(); }
    else { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (CTA2) cluster_sync();   // the peer's barriers exist before anything is signalled on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (p.count && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(p.count), static_cast<unsigned long long>(p.count_add));

  // Each role derives its own scheduler state after the register re-partition so nothing has to
  // be carried (spilled) across the setmaxnreg boundary.
#define EMCID_GEMM_ROLE_SETUP()                                              \
  int N = p.N, K = p.K;                                                      \
  if (p.dyn_n) { int v = *p.dyn_n; N = v < N ? v : N; }                      \
  if (p.dyn_k) { int v = *p.dyn_k; K = v < K ? v : K; }                      \
  const int chunk = p.chunk_kblocks > 0 ? p.chunk_kblocks : GEMM_DEFAULT_CHUNK; \
  const int batch = blockIdx.y;                                              \
  Sched sched;                                                               \
  sched.init(p, BLOCK_N, BLOCK_K, p.M, N, K, CTA2 ? 2 * GEMM_BLOCK_M : GEMM_BLOCK_M,  \
             CTA2 ? blockIdx.x >> 1 : blockIdx.x, CTA2 ? gridDim.x >> 1 : gridDim.x);  \
  Unit u;                                                                    \
  (void)chunk; (void)batch;

  if (warp < 4) {
    // the staged epilogue keeps 128 running sums plus a packed 64-column piece live: give the two epilogue
    // warpgroups 224 registers and shrink this one (producer / MMA issuer / allocator) to 56
    if (EPI == EPI_LINEAR_TMA) setmaxnreg_dec<56>();
    if (warp == 0 && lane == 0) {
      EMCID_GEMM_ROLE_SETUP();
      // ---------------------------------------------------------------- TMA producer
      const int arow = p.a_row0 + batch * p.a_batch_rows;
      const int brow = p.b_row0 + batch * p.b_batch_rows;
      int stage = 0;
      uint32_t phase = 0;
      while (sched.next(u)) {
        for (int kb = u.kb0; kb < u.kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1, 1);
          uint8_t* st = smem + stage * STAGE_BYTES;
          const int kc = kb * BLOCK_K;
          if (KIND == KIND_F16_MN) {
            // planes are [tokens x features]: one box of 64 tokens x 64 features (8 KB) per 64-feature atom of the tile
            constexpr int ATOM = 64 * 128;
            const int tok = p.a_col0 + kc;
            const int fa = arow + u.m0 + static_cast<int>(rank) * GEMM_BLOCK_M;
            const int fb = brow + u.n0 + static_cast<int>(rank) * B_ROWS;
            if (CTA2) {
              const uint32_t lbar = mapa_shared(smem_u32(&full_bar[stage]), 0);
              if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
#pragma unroll
              for (int j = 0; j < GEMM_BLOCK_M / 64; ++j) {
                tma_load_2d_2sm(st + j * ATOM, &tmA_hi, lbar, fa + 64 * j, tok);
                tma_load_2d_2sm(st + A_PLANE + j * ATOM, &tmA_lo, lbar, fa + 64 * j, tok);
              }
#pragma unroll
              for (int j = 0; j < B_ROWS / 64; ++j) {
                tma_load_2d_2sm(st + 2 * A_PLANE + j * ATOM, &tmB_hi, lbar, fb + 64 * j, tok);
                tma_load_2d_2sm(st + 2 * A_PLANE + B_PLANE + j * ATOM, &tmB_lo, lbar, fb + 64 * j, tok);
              }
            } else {
              mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
#pragma unroll
              for (int j = 0; j < GEMM_BLOCK_M / 64; ++j) {
                tma_load_2d(st + j * ATOM, &tmA_hi, &full_bar[stage], fa + 64 * j, tok);
                tma_load_2d(st + A_PLANE + j * ATOM, &tmA_lo, &full_bar[stage], fa + 64 * j, tok);
              }
#pragma unroll
              for (int j = 0; j < B_ROWS / 64; ++j) {
                tma_load_2d(st + 2 * A_PLANE + j * ATOM, &tmB_hi, &full_bar[stage], fb + 64 * j, tok);
                tma_load_2d(st + 2 * A_PLANE + B_PLANE + j * ATOM, &tmB_lo, &full_bar[stage], fb + 64 * j, tok);
              }
            }
          } else if (CTA2) {
            // both CTAs of the pair load into their own stage; all bytes are counted on the leader's barrier
            const uint32_t lbar = mapa_shared(smem_u32(&full_bar[stage]), 0);
            if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
            const int am = arow + u.m0 + static_cast<int>(rank) * GEMM_BLOCK_M;
            const int bn = brow + u.n0 + static_cast<int>(rank) * B_ROWS;
            tma_load_2d_2sm(st, &tmA_hi, lbar, p.a_col0 + kc, am);
            tma_load_2d_2sm(st + A_PLANE, &tmA_lo, lbar, p.a_col0 + kc, am);
            tma_load_2d_2sm(st + 2 * A_PLANE, &tmB_hi, lbar, p.b_col0 + kc, bn);
            tma_load_2d_2sm(st + 2 * A_PLANE + B_PLANE, &tmB_lo, lbar, p.b_col0 + kc, bn);
          } else {
          mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
          tma_load_2d(st, &tmA_hi, &full_bar[stage], p.a_col0 + kc, arow + u.m0);
          tma_load_2d(st + A_PLANE, &tmA_lo, &full_bar[stage], p.a_col0 + kc, arow + u.m0);
          uint8_t* sb = st + 2 * A_PLANE;
#pragma unroll
          for (int r = 0; r < BLOCK_N / 128; ++r) {
            tma_load_2d(sb + r * A_PLANE, &tmB_hi, &full_bar[stage], p.b_col0 + kc,
                        brow + u.n0 + r * 128);
            tma_load_2d(sb + B_PLANE + r * A_PLANE, &tmB_lo, &full_bar[stage],
                        p.b_col0 + kc, brow + u.n0 + r * 128);
          }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && lane == 0 && rank == 0) {
      // ---------------------------------------------------------------- MMA issuer (pair mode: the leader CTA only)
      EMCID_GEMM_ROLE_SETUP();
      // three MMAs per k-step: lo*hi, hi*lo, hi*hi (small terms first)
      const uint32_t f_hi = KIND == KIND_TF32 ? FMT_TF32 : FMT_F16;
      const uint32_t f_lo = KIND == KIND_TF32 ? FMT_TF32 : static_cast<uint32_t>(p.lo_fmt);
      constexpr uint32_t MMA_M = CTA2 ? 2 * GEMM_BLOCK_M : GEMM_BLOCK_M;
      constexpr uint32_t MN_BITS = KIND == KIND_F16_MN ? ((1u << 15) | (1u << 16)) : 0u;   // A and B are MN-major
      const uint32_t idesc_lh = make_idesc(f_lo, f_hi, MMA_M, BLOCK_N) | MN_BITS;
      const uint32_t idesc_hl = make_idesc(f_hi, f_lo, MMA_M, BLOCK_N) | MN_BITS;
      const uint32_t idesc_hh = make_idesc(f_hi, f_hi, MMA_M, BLOCK_N) | MN_BITS;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      while (sched.next(u)) {
        for (int kc0 = u.kb0; kc0 < u.kb1; kc0 += chunk) {
          const int kc1 = kc0 + chunk < u.kb1 ? kc0 + chunk : u.kb1;
          if (CTA2) mbar_wait_cluster(&tempty_bar[acc], acc_phase ^ 1, 2);
          else mbar_wait(&tempty_bar[acc], acc_phase ^ 1, 2);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
          uint32_t accumulate = 0;
          for (int kb = kc0; kb < kc1; ++kb) {
            mbar_wait(&full_bar[stage], phase, 3);
            tc_fence_after();
            const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
            constexpr bool MN = KIND == KIND_F16_MN;
            const uint64_t da_hi = MN ? make_desc_mn128(sa, 8192) : ROWB == 128 ? make_desc_k128(sa) : make_desc_k64(sa);
            const uint64_t da_lo = MN ? make_desc_mn128(sa + A_PLANE, 8192)
                                      : ROWB == 128 ? make_desc_k128(sa + A_PLANE) : make_desc_k64(sa + A_PLANE);
            const uint64_t db_hi = MN ? make_desc_mn128(sa + 2 * A_PLANE, 8192)
                                      : ROWB == 128 ? make_desc_k128(sa + 2 * A_PLANE) : make_desc_k64(sa + 2 * A_PLANE);
            const uint64_t db_lo = MN ? make_desc_mn128(sa + 2 * A_PLANE + B_PLANE, 8192)
                                      : ROWB == 128 ? make_desc_k128(sa + 2 * A_PLANE + B_PLANE)
                                                    : make_desc_k64(sa + 2 * A_PLANE + B_PLANE);
#pragma unroll
            for (int k = 0; k < ROWB / 32; ++k) {
              // one instruction contracts 32 bytes of K (8 tf32 / 16 halves): +2 in the (addr >> 4) field; in an MN-major
              // tile 16 K rows are 16 x 128 B further down: +128
              const uint64_t koff = static_cast<uint64_t>(MN ? k * 128 : k * 2);
              if (KIND == KIND_TF32) {
                tc_mma_tf32(tmem_d, da_lo + koff, db_hi + koff, idesc_lh, accumulate);
                tc_mma_tf32(tmem_d, da_hi + koff, db_lo + koff, idesc_hl, 1);
                tc_mma_tf32(tmem_d, da_hi + koff, db_hi + koff, idesc_hh, 1);
              } else if (CTA2) {
                tc_mma_f16_2(tmem_d, da_lo + koff, db_hi + koff, idesc_lh, accumulate);
                tc_mma_f16_2(tmem_d, da_hi + koff, db_lo + koff, idesc_hl, 1);
                tc_mma_f16_2(tmem_d, da_hi + koff, db_hi + koff, idesc_hh, 1);
              } else {
                tc_mma_f16(tmem_d, da_lo + koff, db_hi + koff, idesc_lh, accumulate);
                tc_mma_f16(tmem_d, da_hi + koff, db_lo + koff, idesc_hl, 1);
                tc_mma_f16(tmem_d, da_hi + koff, db_hi + koff, idesc_hh, 1);
              }
              accumulate = 1;
            }
            if (CTA2) tc_commit2(&empty_bar[stage], 3);   // frees the slot in both CTAs
            else tc_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (CTA2) tc_commit2(&tfull_bar[acc], 3);
          else tc_commit(&tfull_bar[acc]);      // chunk complete -> epilogue folds it into registers
          if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps)
    if (EPI == EPI_LINEAR_TMA) setmaxnreg_inc<224>();
    EMCID_GEMM_ROLE_SETUP();
    const int q = warp & 3;           // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2; // column half of the tile
    int acc = 0;
    uint32_t acc_phase = 0;
    float sum[COLS];
    while (sched.next(u)) {
#pragma unroll
      for (int i = 0; i < COLS; ++i) sum[i] = 0.f;
      if (EPI == EPI_LINEAR_TMA) {
        // this half's bias row goes to shared memory now, one value per thread: the load has the whole mainloop to land
        // (read from global in the store phase it was 32 LDG.128 per thread, four exposed L2 round trips per tile).  Safe
        // against the previous tile's readers: every thread passes the last named barrier of its store phase after its
        // last read of the row.
        const int t = q * 32 + lane, c = u.n0 + half * COLS + t;
        bias_s[half * COLS + t] = c < N ? p.bias_col[c] : 0.f;
      }
      for (int kc0 = u.kb0; kc0 < u.kb1; kc0 += chunk) {
        mbar_wait(&tfull_bar[acc], acc_phase, 4);
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * BLOCK_N + half * COLS +
                               (static_cast<uint32_t>(q * 32) << 16);
        // FW accumulator columns per tcgen05.ld batch: the loads of a batch are issued back to back and share ONE
        // tcgen05.wait::ld.  (ptxas turns the wait into scoreboard dependencies and rotates three register sets through
        // the LDTM.x16s either way — profiles/round2/r04_sass_fc1.txt; the batch width measured within 1 %.)
        constexpr int FW = (EPI == EPI_LINEAR_TMA) ? EMCID_FOLD_WIDTH : 16;
#pragma unroll
        for (int c = 0; c < COLS / FW; ++c) {
          uint32_t v[FW];
#pragma unroll
          for (int j = 0; j < FW / 16; ++j) tmem_ld_32x16_issue(taddr + c * FW + j * 16, v + j * 16);
          tmem_wait_ld<FW>(v);
#pragma unroll
          for (int i = 0; i < FW; ++i) sum[c * FW + i] += __uint_as_float(v[i]);
        }
        tc_fence_before();
        if (CTA2) {
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty_bar[acc]), 0));
        } else {
          mbar_arrive(&tempty_bar[acc]);
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
      }

      // -------- unit finished: registers -> global
      const int m_cta = u.m0 + static_cast<int>(rank) * GEMM_BLOCK_M;   // first row this CTA's accumulators hold
      const int row = m_cta + q * 32 + lane;
      const int col0 = u.n0 + half * COLS;
      if (EPI == EPI_LINEAR_TMA) {
        // Every thread of the half takes part (named barriers): rows >= M hold act(bias) of zero-filled operand
        // rows and are clipped by the TMA store, like columns >= N.
        constexpr int ACT = EFLAGS & EF_ACT_MASK;
        const int r = q * 32 + lane;                               // row inside the tile == TMEM lane
        const uint32_t sbase = smem_u32(stg) + half * (GEMM_STAGING_BYTES / 2);
        const bool issuer = (q == 0) && (lane == 0);
        const int bar_id = 1 + half;
        const float alpha = p.alpha;
        // Staged stores, double buffered.  The half's 16 KB of staging are two 8 KB buffers of 128 rows x 64 bytes (TMA
        // SWIZZLE_64B: 16-byte chunk c of row r lives at chunk c ^ ((r >> 1) & 3)); sub-tile s goes through buffer s & 1:
        //     all 128 threads   write their row of the sub-tile, fence.proxy.async
        //     issuer            waits until every store issued so far has READ its buffer (the youngest was issued one
        //                       sub-tile ago, from the other buffer ... the wait hides behind this sub-tile's arithmetic)
        //     barrier           -> the issuer issues this sub-tile's bulk store; everybody moves on to the other buffer
        // One barrier per 8 KB.  (Until r04 a sub-tile was 16 KB in a single buffer with two barriers, and the 128 threads
        // sat at the first one while the previous store drained.  Measured: no faster — the store phase is arithmetic,
        // not waiting, see DESIGN.md §3 — but half the barriers.  A per-warp variant without any block-level barrier, 2 KB
        // sub-tiles stored by each warp's lane 0, was tried on top: +0.3 % and a race in the shared bias row; dropped.)
        const uint32_t srow64 = sbase + r * 64;
        const int sw64 = (r >> 1) & 3;
        const float* bias_h = bias_s + half * COLS;
        named_bar_sync(bar_id, 128);                               // every thread's bias value is in place
        if (EFLAGS & EF_C) {
#pragma unroll
          for (int st = 0; st < COLS / 16; ++st) {                 // 16 fp32 columns = one 64-byte row
            const uint32_t buf = srow64 + (st & 1) * 8192;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int cl = st * 16 + 4 * k;
              const float4 b = *reinterpret_cast<const float4*>(bias_h + cl);
              const float o0 = act_ct<ACT>(fmaf(alpha, sum[cl], b.x)), o1 = act_ct<ACT>(fmaf(alpha, sum[cl + 1], b.y));
              const float o2 = act_ct<ACT>(fmaf(alpha, sum[cl + 2], b.z)), o3 = act_ct<ACT>(fmaf(alpha, sum[cl + 3], b.w));
              sts_v4(buf + ((k ^ sw64) << 4), __float_as_uint(o0), __float_as_uint(o1), __float_as_uint(o2), __float_as_uint(o3));
            }
            fence_proxy_async();
            if (issuer) bulk_wait_read0();
            named_bar_sync(bar_id, 128);
            if (issuer && col0 + st * 16 < N) {
              const uint32_t src = sbase + (st & 1) * 8192;
              if (EFLAGS & EF_CIN) tma_reduce_add_2d(&om.c, src, col0 + st * 16, m_cta);
              else tma_store_2d(&om.c, src, col0 + st * 16, m_cta);
              bulk_commit();
            }
          }
        }
        static_assert(EPI != EPI_LINEAR_TMA || !(EFLAGS & EF_PT), "the staged epilogue writes no transposed planes");
        if (EFLAGS & EF_P) {
#pragma unroll
          for (int pc = 0; pc < COLS / 64; ++pc) {                 // 64 plane columns: two 32-column (64-byte) sub-tiles per plane
            uint32_t hp[32], lp[32];                               // packed fp16 pairs of the piece
#pragma unroll
            for (int g = 0; g < 2; ++g) {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const int cl = pc * 64 + g * 32 + 4 * k;
                const float4 b = *reinterpret_cast<const float4*>(bias_h + cl);
                split_f16x2(act_ct<ACT>(fmaf(alpha, sum[cl], b.x)), act_ct<ACT>(fmaf(alpha, sum[cl + 1], b.y)),
                            hp[g * 16 + 2 * k], lp[g * 16 + 2 * k]);
                split_f16x2(act_ct<ACT>(fmaf(alpha, sum[cl + 2], b.z)), act_ct<ACT>(fmaf(alpha, sum[cl + 3], b.w)),
                            hp[g * 16 + 2 * k + 1], lp[g * 16 + 2 * k + 1]);
              }
              if (EFLAGS & EF_P) {
                const bool in_n = col0 + pc * 64 + g * 32 < N;
#pragma unroll
                for (int pl = 0; pl < 2; ++pl) {                   // hi sub-tile through buffer 0, lo sub-tile through buffer 1
                  const uint32_t buf = srow64 + pl * 8192;
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const int w = g * 16 + 4 * k;
                    if (pl == 0) sts_v4(buf + ((k ^ sw64) << 4), hp[w], hp[w + 1], hp[w + 2], hp[w + 3]);
                    else sts_v4(buf + ((k ^ sw64) << 4), lp[w], lp[w + 1], lp[w + 2], lp[w + 3]);
                  }
                  fence_proxy_async();
                  if (issuer) bulk_wait_read0();
                  named_bar_sync(bar_id, 128);
                  if (issuer && in_n) {
                    tma_store_2d(pl == 0 ? &om.c : &om.c2, sbase + pl * 8192, col0 + pc * 64 + g * 32, m_cta);
                    bulk_commit();
                  }
                }
              }
            }
          }
        }
      } else if (row < p.M) {
        if (EPI == EPI_RED) {
          float* dst = p.C + batch * p.c_batch + static_cast<long long>(row) * p.ldc + col0;
#pragma unroll
          for (int j = 0; j < COLS / 4; ++j) {
            if (col0 + 4 * j < N)
              red_add_v4(dst + 4 * j, sum[4 * j], sum[4 * j + 1], sum[4 * j + 2], sum[4 * j + 3]);
          }
        } else if (EPI == EPI_GENERIC) {
          const float* cin = p.Cin ? p.Cin + batch * p.cin_batch + static_cast<long long>(row) * p.ldcin + col0 : nullptr;
          float* dst = p.C ? p.C + batch * p.c_batch + static_cast<long long>(row) * p.ldc + col0 : nullptr;
          const long long pidx = batch * p.p_batch + static_cast<long long>(row) * p.ldp + col0;  // plane elements
#pragma unroll
          for (int j = 0; j < COLS / 4; ++j) {
            const int col = col0 + 4 * j;
            if (col < N) {
              float4 o;
              o.x = p.alpha * sum[4 * j];
              o.y = p.alpha * sum[4 * j + 1];
              o.z = p.alpha * sum[4 * j + 2];
              o.w = p.alpha * sum[4 * j + 3];
              if (cin) {
                const float4 old = *reinterpret_cast<const float4*>(cin + 4 * j);
                o.x += p.beta * old.x; o.y += p.beta * old.y; o.z += p.beta * old.z; o.w += p.beta * old.w;
              }
              if (dst) *reinterpret_cast<float4*>(dst + 4 * j) = o;
              if (p.P_hi) store_planes4<KIND>(p.P_hi, p.P_lo, pidx + 4 * j, o, p.lo_fmt);
              if (p.Pt_hi) {
                // transposed planes: consecutive lanes (rows) hit consecutive addresses
                const long long t = batch * p.pt_batch + static_cast<long long>(col) * p.ldpt + row;
                store_plane1<KIND>(p.Pt_hi, p.Pt_lo, t, o.x, p.lo_fmt);
                store_plane1<KIND>(p.Pt_hi, p.Pt_lo, t + p.ldpt, o.y, p.lo_fmt);
                store_plane1<KIND>(p.Pt_hi, p.Pt_lo, t + 2 * p.ldpt, o.z, p.lo_fmt);
                store_plane1<KIND>(p.Pt_hi, p.Pt_lo, t + 3 * p.ldpt, o.w, p.lo_fmt);
              }
            }
          }
        } else if (EPI == EPI_LINEAR) {
          // a linear layer of the text encoder: row = token, col = output feature
          constexpr int ACT = EFLAGS & EF_ACT_MASK;
          const float* cin = (EFLAGS & EF_CIN) ? p.Cin + batch * p.cin_batch + static_cast<long long>(row) * p.ldcin + col0 : nullptr;
          float* dst = (EFLAGS & EF_C) ? p.C + batch * p.c_batch + static_cast<long long>(row) * p.ldc + col0 : nullptr;
          const long long pidx = batch * p.p_batch + static_cast<long long>(row) * p.ldp + col0;
          const long long tidx = batch * p.pt_batch + static_cast<long long>(col0) * p.ldpt + row;
          const float* bias = p.bias_col + col0;
          const float alpha = p.alpha;
#pragma unroll
          for (int j = 0; j < COLS / 4; ++j) {
            if (col0 + 4 * j < N) {
              const float4 b = *reinterpret_cast<const float4*>(bias + 4 * j);
              float4 o;
              o.x = act_ct<ACT>(fmaf(alpha, sum[4 * j], b.x));
              o.y = act_ct<ACT>(fmaf(alpha, sum[4 * j + 1], b.y));
              o.z = act_ct<ACT>(fmaf(alpha, sum[4 * j + 2], b.z));
              o.w = act_ct<ACT>(fmaf(alpha, sum[4 * j + 3], b.w));
              if (EFLAGS & EF_CIN) {
                const float4 old = *reinterpret_cast<const float4*>(cin + 4 * j);
                o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
              }
              if (EFLAGS & EF_C) *reinterpret_cast<float4*>(dst + 4 * j) = o;
              if (EFLAGS & EF_P) store_planes4<KIND>(p.P_hi, p.P_lo, pidx + 4 * j, o, p.lo_fmt);
              if (EFLAGS & EF_PT) {
                const long long t = tidx + static_cast<long long>(4 * j) * p.ldpt;
                store_plane1<KIND>(p.Pt_hi, p.Pt_lo, t, o.x, p.lo_fmt);
                store_plane1<KIND>(p.Pt_hi, p.Pt_lo, t + p.ldpt, o.y, p.lo_fmt);
                store_plane1<KIND>(p.Pt_hi, p.Pt_lo, t + 2 * p.ldpt, o.z, p.lo_fmt);
                store_plane1<KIND>(p.Pt_hi, p.Pt_lo, t + 3 * p.ldpt, o.w, p.lo_fmt);
              }
            }
          }
        } else if (EPI == EPI_FC1) {
          // acc[row = feature i][col = token t]; tokens >= N (dynamic valid count) are padding
          // rows of the compacted X slab and must contribute exactly zero to mom2.
          const float bias = p.bias ? p.bias[row] : 0.f;
          const long long pidx = batch * p.p_batch + static_cast<long long>(row) * p.ldp + col0;
#pragma unroll
          for (int j = 0; j < COLS / 4; ++j) {
            const int col = col0 + 4 * j;
            if (col < p.ldp) {
              float4 a;
              constexpr int ACT = EFLAGS & EF_ACT_MASK;
              a.x = (col + 0 < N) ? act_ct<ACT>(fmaf(p.alpha, sum[4 * j + 0], bias)) : 0.f;
              a.y = (col + 1 < N) ? act_ct<ACT>(fmaf(p.alpha, sum[4 * j + 1], bias)) : 0.f;
              a.z = (col + 2 < N) ? act_ct<ACT>(fmaf(p.alpha, sum[4 * j + 2], bias)) : 0.f;
              a.w = (col + 3 < N) ? act_ct<ACT>(fmaf(p.alpha, sum[4 * j + 3], bias)) : 0.f;
              store_planes4<KIND>(p.P_hi, p.P_lo, pidx + 4 * j, a, p.lo_fmt);
            }
          }
        }
      }
    }
  }

#undef EMCID_GEMM_ROLE_SETUP
  if (EPI == EPI_LINEAR_TMA) bulk_wait0();   // no-op for threads that issued nothing
  tc_fence_before();
  __syncthreads();
  if (CTA2) {
    cluster_sync();   // the leader's MMAs read the peer's shared memory and signal its barriers until the very end
    if (warp == 2) tmem_dealloc2(tmem_base, Cfg::kTmemCols);
  } else if (warp == 2) {
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

}  // namespace emcid

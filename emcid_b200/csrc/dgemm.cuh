// emcid_b200 — fp64 "NT" GEMM on the DMMA tensor path (mma.sync.m8n8k4.f64), sm_100a.
//
//   C[M x N] = alpha * A[M x K] * B[N x K]^T + beta * Cin          (row-major, K contiguous)
//
// Used where the reference demands true double precision and 3xTF32 cannot deliver it:
//   * M64 = lambda * double(C32) + Ks Ks^T            emcid/emcid_main.py:1046
//   * the iterative-refinement residual  R^T = Ks^T - X^T M64   (fp64 accumulate is what makes
//     one refinement step worth ~3 digits; SURVEY.md §7 hard part 3)
//   * upd_matrix = resid @ adj_k.T                     emcid/emcid_main.py:1050
// Block tile 64x128x16, 4 warps (1x4), warp tile 64x32, register-staged prefetch, two CTAs per SM: with one 128x128 CTA of
// 8 warps per SM (200 registers per thread: a second one did not fit) the tensor pipe idled through every barrier and
// global-load wait of the single CTA (ncu: 57-66 % tensor-pipe active, profiles/round1/r03i_ncu_full_dgemm_summary.json); two
// half-height CTAs fill each other's bubbles.
#pragma once

#include "host.cuh"

namespace emcid {

constexpr int DG_BM = 64, DG_BN = 128, DG_BK = 16, DG_THREADS = 2 * DG_BM;
constexpr int DG_LDA = DG_BM + 4, DG_LDB = DG_BN + 4;  // smem pitches (doubles): k-stride == 32 B mod 128 B -> conflict-free frags
constexpr int DG_BSETS = 2 * DG_BN / DG_THREADS;       // B rows per thread and k-tile

struct DgemmParams {
  int M, N, K;
  const double* A; long long lda, a_batch;
  const double* B; long long ldb, b_batch;
  double alpha, beta;
  const double* Cin; long long ldcin, cin_batch;   // fp64 addend (or null)
  const float* Cin32; long long ldcin32, cin32_batch;  // fp32 addend, promoted (or null)
  double* C; long long ldc, c_batch;               // fp64 result (or null)
  float* C32; long long ldc32, c32_batch;          // fp32 copy of the result (or null)
  int lower;                                       // skip tiles strictly above the diagonal
  // split-K (few output tiles, long K: the refinement residual of a narrow edit has 24 tiles of K = 3072 on 148 SMs):
  // split_k > 1 -> grid.z = batches * split_k, slice z of the K range writes alpha * (partial product) to
  // C + z * c_split (plain stores, deterministic); Cin / Cin32 / C32 / beta are ignored and the caller reduces the slices.
  int split_k; long long c_split;
  // persist > 0: that many CTAs walk the (x, y, z) tile space in a grid-stride loop instead of one CTA per tile — a
  // product that runs BESIDE a latency chain on another stream (M64 next to the factorisation) then holds `persist` CTA
  // slots for its whole duration instead of flooding every SM in front of the chain's short, dependent launches.
  int persist;
  int batches_z;   // filled by launch_dgemm_nt: batches * split_k (the z extent of the tile space)
};

// K slices per output tile.  Two CTAs are resident per SM, so `tiles * s` CTAs run in ceil(tiles * s / (2 sms)) waves, the
// last one partly empty: the residual of a 1000-concept single-layer solve has 384 tiles on 296 slots — two waves, the
// second 30 % full (measured 1.12 ms for 19.3 GFLOP = 17 TFLOP/s).  Take the smallest s (each slice at least 256 of K, at
// most `max_s` slices) whose waves are >= 90 % full, else the fullest.
inline int dgemm_pick_split(int tiles, int K, int sms, int max_s = 8) {
  if (tiles <= 0) return 1;
  const int slots = 2 * sms;
  if (max_s > K / 256) max_s = K / 256;
  int best = 1;
  double best_fill = 0.0;
  for (int s = 1; s <= max_s; ++s) {
    const long long ctas = static_cast<long long>(tiles) * s;
    const double fill = static_cast<double>(ctas) / static_cast<double>((ctas + slots - 1) / slots * slots);
    if (fill >= 0.9) return s;
    if (fill > best_fill + 1e-9) { best_fill = fill; best = s; }
  }
  return best;
}

__device__ __forceinline__ void dmma_8x8x4(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(DG_THREADS, 2) dgemm_nt_kernel(const DgemmParams p) {
  __shared__ double As[DG_BK][DG_LDA];
  __shared__ double Bs[DG_BK][DG_LDB];
  const int split = p.split_k > 1 ? p.split_k : 1;
  const int tiles_x = (p.N + DG_BN - 1) / DG_BN, tiles_y = (p.M + DG_BM - 1) / DG_BM;
  const long long n_items = p.persist > 0 ? static_cast<long long>(tiles_x) * tiles_y * p.batches_z : 1;
  for (long long item = p.persist > 0 ? blockIdx.x : 0; item < n_items; item += p.persist > 0 ? gridDim.x : 1) {
  int bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z;
  if (p.persist > 0) {
    bx = static_cast<int>(item % tiles_x);
    by = static_cast<int>((item / tiles_x) % tiles_y);
    bz = static_cast<int>(item / (static_cast<long long>(tiles_x) * tiles_y));
  }
  const int bm = by * DG_BM, bn = bx * DG_BN;
  if (p.lower && bn > bm + DG_BM - 1) continue;
  const int batch = bz / split, kz = bz - batch * split;
  const int k_chunk = ((p.K + split - 1) / split + DG_BK - 1) / DG_BK * DG_BK;
  const int k_begin = kz * k_chunk, k_end = min(p.K, k_begin + k_chunk);
  const double* A = p.A + batch * p.a_batch;
  const double* B = p.B + batch * p.b_batch;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = (warp >> 2) * 64, wn = (warp & 3) * 32;

  // global->register staging: the A tile is DG_BM rows x 16 k, the B tile DG_BN rows x 16 k; thread owns
  // (row = tid/2 [+ s * DG_THREADS/2 for B], k8 = (tid&1)*8 .. +8)
  const int lrow = tid >> 1, lk = (tid & 1) * 8;
  double ra[8], rb[DG_BSETS][8];
  auto load_tiles = [&](int k0) {
    const int ar = bm + lrow;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + lk + j;
      ra[j] = (ar < p.M && k < k_end) ? A[static_cast<long long>(ar) * p.lda + k] : 0.0;
#pragma unroll
      for (int sb = 0; sb < DG_BSETS; ++sb) {
        const int br = bn + lrow + sb * (DG_THREADS / 2);
        rb[sb][j] = (br < p.N && k < k_end) ? B[static_cast<long long>(br) * p.ldb + k] : 0.0;
      }
    }
  };
  auto store_tiles = [&]() {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      As[lk + j][lrow] = ra[j];
#pragma unroll
      for (int sb = 0; sb < DG_BSETS; ++sb) Bs[lk + j][lrow + sb * (DG_THREADS / 2)] = rb[sb][j];
    }
  };

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nk = k_end > k_begin ? (k_end - k_begin + DG_BK - 1) / DG_BK : 0;
  load_tiles(k_begin);
  store_tiles();
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    if (kt + 1 < nk) load_tiles(k_begin + (kt + 1) * DG_BK);  // global loads in flight during the MMAs
#pragma unroll
    for (int ks = 0; ks < DG_BK; ks += 4) {
      double af[8], bf[4];
      const int kk = ks + (lane & 3), r = lane >> 2;
#pragma unroll
      for (int i = 0; i < 8; ++i) af[i] = As[kk][wm + i * 8 + r];
#pragma unroll
      for (int j = 0; j < 4; ++j) bf[j] = Bs[kk][wn + j * 8 + r];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) dmma_8x8x4(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
    }
    __syncthreads();
    if (kt + 1 < nk) store_tiles();
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int row = bm + wm + i * 8 + (lane >> 2);
    if (row >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = bn + wn + j * 8 + 2 * (lane & 3);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int c = col + e;
        if (c >= p.N) continue;
        double v = p.alpha * acc[i][j][e];
        if (split > 1) {
          p.C[kz * p.c_split + batch * p.c_batch + static_cast<long long>(row) * p.ldc + c] = v;
          continue;
        }
        if (p.Cin) v += p.beta * p.Cin[batch * p.cin_batch + static_cast<long long>(row) * p.ldcin + c];
        if (p.Cin32)
          v += p.beta * static_cast<double>(p.Cin32[batch * p.cin32_batch + static_cast<long long>(row) * p.ldcin32 + c]);
        if (p.C) p.C[batch * p.c_batch + static_cast<long long>(row) * p.ldc + c] = v;
        if (p.C32) p.C32[batch * p.c32_batch + static_cast<long long>(row) * p.ldc32 + c] = static_cast<float>(v);
      }
    }
  }
  __syncthreads();   // persistent mode: the next item's staging stores must not overtake this item's last fragment loads
  }
}

inline int launch_dgemm_nt(const DgemmParams& p_in, int batches, cudaStream_t stream) {
  DgemmParams p = p_in;
  p.batches_z = batches * (p.split_k > 1 ? p.split_k : 1);
  dim3 grid((p.N + DG_BN - 1) / DG_BN, (p.M + DG_BM - 1) / DG_BM, p.batches_z);
  if (p.persist > 0) grid = dim3(p.persist, 1, 1);
  dgemm_nt_kernel<<<grid, DG_THREADS, 0, stream>>>(p);
  EMCID_CUDA_CHECK(cudaGetLastError());
  return EMCID_OK;
}

}  // namespace emcid

// emcid_b200 — C ABI for the generic 3xTF32 NT GEMM (building block + test surface).
#pragma once

#include "host.cuh"

namespace emcid {

inline long long round_up_ll(long long x, long long m) { return (x + m - 1) / m * m; }

// Workspace layout for emcid_gemm3x_nt: [A_hi | A_lo | B_hi | B_lo], pitch = round_up(K, 32).
inline size_t gemm3x_workspace_bytes(int M, int N, int K, bool same_ab) {
  const long long kp = round_up_ll(K, GEMM_BLOCK_K);
  long long elems = 2ll * M * kp;
  if (!same_ab) elems += 2ll * N * kp;
  return static_cast<size_t>(elems) * sizeof(float) + 1024;
}

// C = alpha * A B^T + beta * C  (fp32 in/out, 3xTF32 on tcgen05).  flags: bit0 = lower tiles only,
// bit1 = stream-K with red.add epilogue (requires beta == 1, alpha == 1), bit2 = BLOCK_N 128,
// bits 8-15 = k-blocks per TMEM chunk (0 = default).
inline int gemm3x_nt(int M, int N, int K, const float* A, long long lda, const float* B,
                     long long ldb, float* C, long long ldc, float alpha, float beta, int flags,
                     void* workspace, size_t ws_bytes, cudaStream_t stream) {
  EMCID_CHECK(M > 0 && N > 0 && K > 0, EMCID_ERR_INVALID, "gemm3x_nt: empty problem");
  EMCID_CHECK(N % 4 == 0 && ldc % 4 == 0, EMCID_ERR_INVALID, "gemm3x_nt: N and ldc must be multiples of 4");
  EMCID_CHECK((reinterpret_cast<uintptr_t>(C) & 15) == 0, EMCID_ERR_INVALID, "gemm3x_nt: C must be 16B aligned");
  DeviceInfo info;
  int rc = get_device_info(&info);
  if (rc) return rc;
  const bool same_ab = (A == B && lda == ldb && M == N);
  EMCID_CHECK(ws_bytes >= gemm3x_workspace_bytes(M, N, K, same_ab), EMCID_ERR_WORKSPACE,
              "gemm3x_nt: workspace too small (%zu < %zu)", ws_bytes,
              gemm3x_workspace_bytes(M, N, K, same_ab));
  const long long kp = round_up_ll(K, GEMM_BLOCK_K);
  float* ws = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(workspace) + 1023) &
                                       ~static_cast<uintptr_t>(1023));
  float* a_hi = ws;
  float* a_lo = a_hi + static_cast<long long>(M) * kp;
  float* b_hi = same_ab ? a_hi : a_lo + static_cast<long long>(M) * kp;
  float* b_lo = same_ab ? a_lo : b_hi + static_cast<long long>(N) * kp;
  rc = launch_split_planes(A, lda, M, K, 1.0f, a_hi, a_lo, kp, stream);
  if (rc) return rc;
  if (!same_ab) {
    rc = launch_split_planes(B, ldb, N, K, 1.0f, b_hi, b_lo, kp, stream);
    if (rc) return rc;
  }
  GemmOperands ops;
  if ((rc = make_tmap_2d(&ops.a_hi, a_hi, M, K, kp, 128))) return rc;
  if ((rc = make_tmap_2d(&ops.a_lo, a_lo, M, K, kp, 128))) return rc;
  if ((rc = make_tmap_2d(&ops.b_hi, b_hi, N, K, kp, 128))) return rc;
  if ((rc = make_tmap_2d(&ops.b_lo, b_lo, N, K, kp, 128))) return rc;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.K = K;
  p.lower = flags & 1;
  p.streamk = (flags >> 1) & 1;
  p.chunk_kblocks = (flags >> 8) & 0xff;  // 0 = library default
  p.C = C; p.ldc = ldc; p.alpha = alpha; p.beta = beta;
  if (beta != 0.f) { p.Cin = C; p.ldcin = ldc; }
  const bool n128 = (flags >> 2) & 1;
  const int block_n = n128 ? 128 : 256;
  const int tiles = gemm_num_tiles(M, N, block_n, p.lower);
  int grid = p.streamk ? info.sm_count : (tiles < info.sm_count ? tiles : info.sm_count);
  if (p.streamk) {
    EMCID_CHECK(alpha == 1.0f && beta == 1.0f, EMCID_ERR_INVALID,
                "gemm3x_nt: stream-K accumulates with red.add and needs alpha == beta == 1");
    return n128 ? launch_gemm3x<128, 3, EPI_RED>(ops, p, grid, stream)
                : launch_gemm3x<256, 2, EPI_RED>(ops, p, grid, stream);
  }
  return n128 ? launch_gemm3x<128, 3, EPI_GENERIC>(ops, p, grid, stream)
              : launch_gemm3x<256, 2, EPI_GENERIC>(ops, p, grid, stream);
}

}  // namespace emcid

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

// C = alpha * A B^T + beta * C  (fp32 in/out, 3-term split on tcgen05).  flags: bit0 = lower tiles only,
// bit1 = stream-K with red.add epilogue (requires beta == 1, alpha == 1), bit2 = BLOCK_N 128,
// bit3 = 3xFP16: 16-bit planes (fp16 hi + fp16 lo) on kind::f16,
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
  const bool f16 = (flags >> 3) & 1;
  const int lo_fmt = FMT_F16;  // (a bf16 lo plane next to an fp16 hi plane traps: kind::f16 wants A and B formats equal per MMA... measured on B200)
  const int eb = f16 ? 2 : 4;
  const long long kp = round_up_ll(K, f16 ? 64 : GEMM_BLOCK_K);
  // plane pointers are kept as float* (the kernel reinterprets them for 16-bit planes); the workspace is
  // sized for fp32 planes, so 16-bit planes at the same element offsets always fit
  uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 1023) &
                                           ~static_cast<uintptr_t>(1023));
  const long long a_bytes = static_cast<long long>(M) * kp * eb, b_bytes = static_cast<long long>(N) * kp * eb;
  uint8_t* a_hi = ws;
  uint8_t* a_lo = a_hi + a_bytes;
  uint8_t* b_hi = same_ab ? a_hi : a_lo + a_bytes;
  uint8_t* b_lo = same_ab ? a_lo : b_hi + b_bytes;
  if (f16) {
    rc = launch_split_planes16(A, lda, M, K, 1.0f, a_hi, a_lo, kp, lo_fmt, stream);
    if (!rc && !same_ab) rc = launch_split_planes16(B, ldb, N, K, 1.0f, b_hi, b_lo, kp, lo_fmt, stream);
  } else {
    rc = launch_split_planes(A, lda, M, K, 1.0f, reinterpret_cast<float*>(a_hi), reinterpret_cast<float*>(a_lo), kp, stream);
    if (!rc && !same_ab)
      rc = launch_split_planes(B, ldb, N, K, 1.0f, reinterpret_cast<float*>(b_hi), reinterpret_cast<float*>(b_lo), kp, stream);
  }
  if (rc) return rc;
  GemmOperands ops;
  if ((rc = make_tmap_2d(&ops.a_hi, a_hi, M, K, kp, 128, eb))) return rc;
  if ((rc = make_tmap_2d(&ops.a_lo, a_lo, M, K, kp, 128, eb))) return rc;
  if ((rc = make_tmap_2d(&ops.b_hi, b_hi, N, K, kp, 128, eb))) return rc;
  if ((rc = make_tmap_2d(&ops.b_lo, b_lo, N, K, kp, 128, eb))) return rc;

  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.lo_fmt = lo_fmt;
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
    if (f16) return n128 ? launch_gemm3x<128, 3, EPI_RED, KIND_F16>(ops, p, grid, stream)
                         : launch_gemm3x<256, 2, EPI_RED, KIND_F16>(ops, p, grid, stream);
    return n128 ? launch_gemm3x<128, 3, EPI_RED>(ops, p, grid, stream)
                : launch_gemm3x<256, 2, EPI_RED>(ops, p, grid, stream);
  }
  if (f16) return n128 ? launch_gemm3x<128, 3, EPI_GENERIC, KIND_F16>(ops, p, grid, stream)
                       : launch_gemm3x<256, 2, EPI_GENERIC, KIND_F16>(ops, p, grid, stream);
  return n128 ? launch_gemm3x<128, 3, EPI_GENERIC>(ops, p, grid, stream)
              : launch_gemm3x<256, 2, EPI_GENERIC>(ops, p, grid, stream);
}

}  // namespace emcid

// emcid_b200 — unity translation unit of libemcid_b200.so (sm_100a only).
//
// C ABI declared in include/emcid_b200.h.  Everything here is hand-written CUDA for Blackwell
// (tcgen05 / TMEM / TMA); there is deliberately no CPU fallback: on a device that is not
// sm_100 every entry point fails with EMCID_ERR_UNSUPPORTED.
#include "../../include/emcid_b200.h"

#include "gemm_api.cuh"

using namespace emcid;

extern "C" {

const char* emcid_last_error(void) { return last_error_buf(); }

int emcid_version(void) { return 100; }

int emcid_device_check(int device) {
  EMCID_CUDA_CHECK(cudaSetDevice(device));
  DeviceInfo info;
  return get_device_info(&info);
}

unsigned int emcid_hang_code(void) {
  unsigned int v = 0;
  cudaMemcpyFromSymbol(&v, g_emcid_hang_code, sizeof(v));
  return v;
}

size_t emcid_gemm3x_workspace_bytes(int M, int N, int K) {
  return gemm3x_workspace_bytes(M, N, K, false);
}

int emcid_gemm3x_nt(int M, int N, int K, const float* A, long long lda, const float* B,
                    long long ldb, float* C, long long ldc, float alpha, float beta, int flags,
                    void* workspace, size_t workspace_bytes, void* stream) {
  return gemm3x_nt(M, N, K, A, lda, B, ldb, C, ldc, alpha, beta, flags, workspace, workspace_bytes,
                   static_cast<cudaStream_t>(stream));
}

}  // extern "C"

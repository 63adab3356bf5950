// emcid_b200 — unity translation unit of libemcid_b200.so (sm_100a only).
//
// C ABI declared in include/emcid_b200.h.  Everything here is hand-written CUDA for Blackwell
// (tcgen05 / TMEM / TMA); there is deliberately no CPU fallback: on a device that is not
// sm_100 every entry point fails with EMCID_ERR_UNSUPPORTED.
#include "../../include/emcid_b200.h"

#include "gemm_api.cuh"
#include "mom2.cuh"
#include "solve.cuh"

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

// ---------------------------------------------------------------------------------------------
// mom2 statistics pass
// ---------------------------------------------------------------------------------------------
size_t emcid_mom2_workspace_bytes(int d, int h, int slab_tokens) {
  if (slab_tokens <= 0) slab_tokens = MOM2_DEFAULT_SLAB;
  return mom2_workspace_bytes(d, h, slab_tokens);
}

int emcid_mom2_create(emcid_mom2_t** out, int device, int d, int h, int act, int slab_tokens,
                      void* workspace, size_t workspace_bytes) {
  return mom2_create(reinterpret_cast<Mom2Handle**>(out), device, d, h, act, slab_tokens, workspace,
                     workspace_bytes);
}

int emcid_mom2_set_chunks(emcid_mom2_t* h, int fc1_kblocks, int syrk_kblocks) {
  Mom2Handle* H = reinterpret_cast<Mom2Handle*>(h);
  EMCID_CHECK(H && fc1_kblocks > 0 && syrk_kblocks > 0 && fc1_kblocks < 256 && syrk_kblocks < 256,
              EMCID_ERR_INVALID, "emcid_mom2_set_chunks: bad argument");
  H->chunk_fc1 = fc1_kblocks;
  H->chunk_syrk = syrk_kblocks;
  return EMCID_OK;
}

int emcid_mom2_set_precision(emcid_mom2_t* h, int precision) {
  return mom2_set_precision(reinterpret_cast<Mom2Handle*>(h), precision);
}

int emcid_mom2_set_weights(emcid_mom2_t* h, const float* W1, long long ldw, const float* b1, void* stream) {
  return mom2_set_weights(reinterpret_cast<Mom2Handle*>(h), W1, ldw, b1, static_cast<cudaStream_t>(stream));
}

int emcid_mom2_accumulate(emcid_mom2_t* h, const float* X, long long ldx, const uint8_t* valid,
                          long long T, void* stream) {
  return mom2_accumulate(reinterpret_cast<Mom2Handle*>(h), X, ldx, valid, T, static_cast<cudaStream_t>(stream));
}

int emcid_mom2_finalize(emcid_mom2_t* h, float* mom2_full, long long* count_dev, void* stream) {
  return mom2_finalize(reinterpret_cast<Mom2Handle*>(h), mom2_full, count_dev, static_cast<cudaStream_t>(stream));
}

int emcid_mom2_reset(emcid_mom2_t* h, void* stream) {
  return mom2_reset(reinterpret_cast<Mom2Handle*>(h), static_cast<cudaStream_t>(stream));
}

int emcid_mom2_profile(emcid_mom2_t* h, int enable) {
  Mom2Handle* H = reinterpret_cast<Mom2Handle*>(h);
  EMCID_CHECK(H, EMCID_ERR_INVALID, "emcid_mom2_profile: null handle");
  H->profile = enable != 0;
  return EMCID_OK;
}

int emcid_mom2_get_profile(emcid_mom2_t* h, double* out8) { return mom2_get_profile(reinterpret_cast<Mom2Handle*>(h), out8); }

int emcid_mom2_destroy(emcid_mom2_t* h) { return mom2_destroy(reinterpret_cast<Mom2Handle*>(h)); }

// ---------------------------------------------------------------------------------------------
// closed-form update
// ---------------------------------------------------------------------------------------------
size_t emcid_solve_workspace_bytes(int batch, int d, int h, int n) { return solve_workspace_bytes(batch, d, h, n); }

int emcid_solve_layers(int device, int batch, int d, int h, int n, const float* C32, const float* Kt,
                       long long ldk, const float* St, long long lds, double lambda, double scale,
                       const double* inv_layers_left, double* adj_k, double* resid, float* dW,
                       int refine_steps, void* workspace, size_t workspace_bytes, int* status_dev, void* stream) {
  return solve_layers(device, batch, d, h, n, C32, Kt, ldk, St, lds, lambda, scale, inv_layers_left, adj_k, resid,
                      dW, refine_steps, workspace, workspace_bytes, status_dev, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

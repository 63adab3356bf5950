// emcid_b200 — unity translation unit of libemcid_b200.so (sm_100a only).
//
// C ABI declared in include/emcid_b200.h.  Everything here is hand-written CUDA for Blackwell
// (tcgen05 / TMEM / TMA); there is deliberately no CPU fallback: on a device that is not
// sm_100 every entry point fails with EMCID_ERR_UNSUPPORTED.
#include "../../include/emcid_b200.h"

#include "gemm_api.cuh"
#include "mom2.cuh"
#include "solve.cuh"
#include "clip.cuh"
#include "reduce.cuh"

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
  if (slab_tokens <= 0) slab_tokens = mom2_default_slab(d);
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

int emcid_nccl_available(void) { return nccl_api() != nullptr; }

int emcid_mom2_reduce(emcid_mom2_t* h, void* nccl_comm, int root, void* stream) {
  return mom2_reduce(reinterpret_cast<Mom2Handle*>(h), nccl_comm, root, static_cast<cudaStream_t>(stream));
}

int emcid_mom2_broadcast(float* mom2_full, long long* count_dev, int d, void* nccl_comm, int root, void* stream) {
  return mom2_broadcast_full(mom2_full, count_dev, d, nccl_comm, root, static_cast<cudaStream_t>(stream));
}

size_t emcid_mom2_state_elems(int d) { return d > 0 ? static_cast<size_t>(packed_lower_elems(d)) : 0; }

int emcid_mom2_export_state(emcid_mom2_t* h, double* lower_packed_dev, long long* count_dev, void* stream) {
  return mom2_export_state(reinterpret_cast<Mom2Handle*>(h), lower_packed_dev, count_dev, static_cast<cudaStream_t>(stream));
}

int emcid_mom2_import_state(emcid_mom2_t* h, const double* lower_packed_dev, const long long* count_dev, void* stream) {
  return mom2_import_state(reinterpret_cast<Mom2Handle*>(h), lower_packed_dev, count_dev, static_cast<cudaStream_t>(stream));
}

int emcid_symmetrize_lower(float* C, int d, long long ldc, void* stream) {
  return symmetrize_lower(C, d, ldc, static_cast<cudaStream_t>(stream));
}

int emcid_checksum_tensors(const void* table_dev, int n, unsigned long long* out_dev, void* stream) {
  return checksum_tensors(table_dev, n, out_dev, static_cast<cudaStream_t>(stream));
}

int emcid_fixed_random_subset(long long n_items, long long seed, long long* out, long long n_out) {
  return fixed_random_subset(n_items, seed, out, n_out);
}

// ---------------------------------------------------------------------------------------------
// native text-encoder forward feeding the statistics pass
// ---------------------------------------------------------------------------------------------
int emcid_clip_create(emcid_clip_t** out, int device, int n_layers, int hidden, int heads, int intermediate, int act,
                      int max_positions, int vocab, float ln_eps, long long max_tokens, int max_captions) {
  return clip_create(reinterpret_cast<ClipHandle**>(out), device, n_layers, hidden, heads, intermediate, act,
                     max_positions, vocab, ln_eps, max_tokens, max_captions);
}

int emcid_clip_set_embeddings(emcid_clip_t* h, const float* token_embedding, const float* position_embedding, void* stream) {
  return clip_set_embeddings(reinterpret_cast<ClipHandle*>(h), token_embedding, position_embedding,
                             static_cast<cudaStream_t>(stream));
}

int emcid_clip_set_layer(emcid_clip_t* h, int layer, const float* const* tensors16, void* stream) {
  return clip_set_layer(reinterpret_cast<ClipHandle*>(h), layer, tensors16, static_cast<cudaStream_t>(stream));
}

int emcid_clip_update_layer(emcid_clip_t* h, int layer, const float* const* tensors16, unsigned int changed_mask, void* stream) {
  return clip_set_layer(reinterpret_cast<ClipHandle*>(h), layer, tensors16, static_cast<cudaStream_t>(stream), changed_mask);
}

int emcid_clip_forward(emcid_clip_t* h, const int32_t* ids, const int32_t* positions, const int32_t* cu_seqlens,
                       int n_captions, int n_tokens, int n_layers, int n_stat, const int* stat_layers,
                       emcid_mom2_t* const* accs, float* hidden_out, void* stream) {
  return clip_forward(reinterpret_cast<ClipHandle*>(h), ids, positions, cu_seqlens, n_captions, n_tokens, n_layers, n_stat,
                      stat_layers, reinterpret_cast<Mom2Handle* const*>(accs), hidden_out, static_cast<cudaStream_t>(stream));
}

int emcid_clip_forward_keys(emcid_clip_t* h, const int32_t* ids, const int32_t* positions, const int32_t* cu_seqlens,
                            int n_captions, int n_tokens, int layer, const int32_t* key_rows, int n_keys, float* k_out,
                            float* z_out, int resume_layer, void* stream) {
  return clip_forward(reinterpret_cast<ClipHandle*>(h), ids, positions, cu_seqlens, n_captions, n_tokens, 0, 0, nullptr,
                      nullptr, nullptr, static_cast<cudaStream_t>(stream), layer, key_rows, n_keys, k_out, z_out, resume_layer);
}

int emcid_clip_set_final_norm(emcid_clip_t* h, const float* weight, const float* bias, void* stream) {
  return clip_set_final_norm(reinterpret_cast<ClipHandle*>(h), weight, bias, static_cast<cudaStream_t>(stream));
}

int emcid_clip_forward_final(emcid_clip_t* h, const int32_t* ids, const int32_t* positions, const int32_t* cu_seqlens,
                             int n_captions, int n_tokens, emcid_mom2_t* acc, const int32_t* rows, int n_rows, float* out,
                             void* stream) {
  return clip_forward_final(reinterpret_cast<ClipHandle*>(h), ids, positions, cu_seqlens, n_captions, n_tokens,
                            reinterpret_cast<Mom2Handle*>(acc), rows, n_rows, out, static_cast<cudaStream_t>(stream));
}

int emcid_clip_profile(emcid_clip_t* h, int enable) {
  ClipHandle* H = reinterpret_cast<ClipHandle*>(h);
  EMCID_CHECK(H, EMCID_ERR_INVALID, "emcid_clip_profile: null handle");
  H->profile = enable != 0;
  return EMCID_OK;
}

int emcid_clip_get_profile(emcid_clip_t* h, double* out21) { return clip_get_profile(reinterpret_cast<ClipHandle*>(h), out21); }

long long emcid_clip_launches(emcid_clip_t* h) { return h ? reinterpret_cast<ClipHandle*>(h)->launches : 0; }

int emcid_clip_destroy(emcid_clip_t* h) { return clip_destroy(reinterpret_cast<ClipHandle*>(h)); }

int emcid_release_cached_memory(void) { return dev_release_cached(); }

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

int emcid_read_npz_f32(const char* const* paths, int n, const char* key, float* out, long long elems, int* bad) {
  return read_npz_f32(paths, n, key, out, elems, bad);
}

int emcid_delta_update(int h, int d, int n, const double* resid, const double* adj_k, float* dW, void* stream) {
  EMCID_CHECK(h > 0 && d > 0 && n > 0 && resid && adj_k && dW, EMCID_ERR_INVALID, "emcid_delta_update: bad argument");
  DgemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = h; p.N = d; p.K = n;
  p.A = resid; p.lda = n;
  p.B = adj_k; p.ldb = n;
  p.alpha = 1.0;
  p.C32 = dW; p.ldc32 = d;
  return launch_dgemm_nt(p, 1, static_cast<cudaStream_t>(stream));
}

/* cached factorisation for repeated edits with the same covariance */
int emcid_factor_create(emcid_factor_t** out, int device, int d, const float* C32, double lambda, int* status_dev,
                        void* stream) {
  return factor_create(reinterpret_cast<FactorHandle**>(out), device, d, C32, lambda, status_dev,
                       static_cast<cudaStream_t>(stream));
}

size_t emcid_factor_solve_workspace_bytes(int d, int h, int n) { (void)h; return factor_solve_workspace_bytes(d, n); }

int emcid_factor_solve(emcid_factor_t* f, int h, int n, const float* Kt, long long ldk, const float* St, long long lds,
                       double scale, double inv_layers_left, double* adj_k, double* resid, float* dW, int refine_steps,
                       void* workspace, size_t workspace_bytes, int* status_dev, void* stream) {
  return factor_solve(reinterpret_cast<FactorHandle*>(f), h, n, Kt, ldk, St, lds, scale, inv_layers_left, adj_k, resid, dW,
                      refine_steps, workspace, workspace_bytes, status_dev, static_cast<cudaStream_t>(stream));
}

int emcid_factor_destroy(emcid_factor_t* f) { return factor_destroy(reinterpret_cast<FactorHandle*>(f)); }

}  // extern "C"

/* emcid_b200 — C ABI of libemcid_b200.so (sm_100a CUDA kernels for the EMCID hot path).
 *
 * The reference (SilentView/EMCID) has no FFI: its hot path is plain PyTorch calls.  Each entry
 * point below names the reference call site it replaces (paths relative to the reference root).
 * All functions return 0 on success or a negative EMCID_ERR_* code; emcid_last_error() returns a
 * thread-local description.  No exceptions cross the boundary.  Every buffer is owned by the
 * caller; device pointers are borrowed for the duration of the call and all work is enqueued on
 * the caller's stream (a cudaStream_t passed as void*).  Handles are not thread-safe; distinct
 * handles are independent.
 */
#ifndef EMCID_B200_H_
#define EMCID_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EMCID_OK 0
#define EMCID_ERR_INVALID (-1)
#define EMCID_ERR_CUDA (-2)
#define EMCID_ERR_UNSUPPORTED (-3)
#define EMCID_ERR_NUMERIC (-4)
#define EMCID_ERR_WORKSPACE (-5)

#define EMCID_ACT_QUICK_GELU 0 /* x * sigmoid(1.702 x): CLIP ViT-L/14 text (sd-text, sdxl-text1) */
#define EMCID_ACT_GELU_ERF 1   /* exact erf GELU: OpenCLIP bigG text (sdxl-text2) */
#define EMCID_ACT_NONE 2

const char* emcid_last_error(void);
int emcid_version(void);
/* Fails with EMCID_ERR_UNSUPPORTED unless `device` is compute capability 10.x. */
int emcid_device_check(int device);
/* Non-zero after a kernel aborted on a barrier spin-limit (debug aid). */
unsigned int emcid_hang_code(void);

/* ---- generic 3xTF32 GEMM:  C = alpha * A * B^T + beta * C ------------------------------------
 * A [M x K], B [N x K], C [M x N], row-major fp32 on the device.  Replaces the fp32/fp64 `@` /
 * `.mm` products on the path (util/runningstats.py:493, emcid/emcid_main.py:1046,1050).
 * flags: bit0 = compute only tiles touching the lower triangle, bit1 = stream-K (alpha = beta = 1,
 * accumulation by red.global.add), bit2 = 128-wide N tiles. */
size_t emcid_gemm3x_workspace_bytes(int M, int N, int K);
int emcid_gemm3x_nt(int M, int N, int K, const float* A, long long lda, const float* B,
                    long long ldb, float* C, long long ldc, float alpha, float beta, int flags,
                    void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EMCID_B200_H_ */

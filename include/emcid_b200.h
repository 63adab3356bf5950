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
 * accumulation by red.global.add), bit2 = 128-wide N tiles, bit3 = 3xFP16 planes on kind::f16
 * instead of 3xTF32, bits 8-15 = k-blocks per TMEM chunk. */
size_t emcid_gemm3x_workspace_bytes(int M, int N, int K);
int emcid_gemm3x_nt(int M, int N, int K, const float* A, long long lda, const float* B,
                    long long ldb, float* C, long long ldc, float alpha, float beta, int flags,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---- mom2 statistics pass (one handle per edited layer) ---------------------------------------
 * Replaces, per sub-batch, the reference's
 *     feats = flatten_masked_batch(tr.input, batch["attention_mask"])   dsets/stat_dataset.py:166-172
 *     stat.add(feats)  ->  count += T; mom2 += feats.t().mm(feats)      util/runningstats.py:483-493
 * (called from emcid/layer_stats.py:208-219), where tr.input = act(fc1(X)) is recomputed on-chip
 * from X = LN2 output, the argument of the CLIP MLP (transformers modeling_clip.py CLIPMLP.forward).
 *
 * workspace: scratch of emcid_mom2_workspace_bytes() bytes on `device`, owned by the caller; it may
 * be shared by several handles that are only ever used on the same stream.  slab_tokens: rows
 * processed per fc1/SYRK launch pair (multiple of 256, <= 4096; 0 = default: as many as keep the slab L2 resident, 4096 at d = 3072).
 * act: EMCID_ACT_*.  The library allocates the handle-private accumulators (fp32 + fp64 [d x d],
 * tf32 planes of W1) with cudaMalloc; emcid_mom2_destroy frees them. */
typedef struct emcid_mom2 emcid_mom2_t;
size_t emcid_mom2_workspace_bytes(int d, int h, int slab_tokens);
int emcid_mom2_create(emcid_mom2_t** out, int device, int d, int h, int act, int slab_tokens,
                      void* workspace, size_t workspace_bytes);
/* Operand precision of the two tensor-core products (fc1 and SYRK); call before emcid_mom2_set_weights.
 *   EMCID_PREC_TF32X3  3xTF32: hi/lo tf32 planes, tcgen05.mma.kind::tf32                 (|err| <= 2^-23 |x| per operand)
 *   EMCID_PREC_F16X3   3xFP16: fp16 hi + fp16 lo planes, tcgen05.mma.kind::f16 (2x rate)  (max(2^-23 |x|, 2^-25)) */
#define EMCID_PREC_TF32X3 0
#define EMCID_PREC_F16X3 1
int emcid_mom2_set_precision(emcid_mom2_t* h, int precision);
/* k-blocks (128 bytes of contraction each: 32 tf32 / 64 halves) accumulated in TMEM between round-to-nearest folds. */
int emcid_mom2_set_chunks(emcid_mom2_t* h, int fc1_kblocks, int syrk_kblocks);
/* W1 [d x h] row-major (pitch ldw floats) and b1 [d] (may be NULL): fc1.weight / fc1.bias. */
int emcid_mom2_set_weights(emcid_mom2_t* h, const float* W1, long long ldw, const float* b1, void* stream);
/* X [T x h] fp32 (pitch ldx), valid [T] bytes (non-zero = attended token; NULL = all rows).
 * count += #valid ; mom2 += sum over valid rows of a a^T with a = act(W1 x + b1). */
int emcid_mom2_accumulate(emcid_mom2_t* h, const float* X, long long ldx, const uint8_t* valid,
                          long long T, void* stream);
/* mom2_full [d x d] fp32 device (full symmetric matrix, the layout SecondMoment.state_dict saves,
 * util/runningstats.py:502-507); count_dev: device int64 scalar (may be NULL).  Does not reset. */
int emcid_mom2_finalize(emcid_mom2_t* h, float* mom2_full, long long* count_dev, void* stream);
int emcid_mom2_reset(emcid_mom2_t* h, void* stream);
/* Measurement aid: when enabled, every fc1 / SYRK launch is bracketed by CUDA events on the caller's
 * stream.  emcid_mom2_get_profile waits for them and fills out8 = {fc1 ms total, fc1 launches, fc1 slab
 * rows, syrk ms total, syrk launches, syrk slab rows, kernels launched by this handle so far, 0}. */
int emcid_mom2_profile(emcid_mom2_t* h, int enable);
int emcid_mom2_get_profile(emcid_mom2_t* h, double* out8);
int emcid_mom2_destroy(emcid_mom2_t* h);

/* ---- the exchange step of a caption-sharded pass ----------------------------------------------------
 * The reference's pass is one process whose only cross-sample dependency is the running sum
 *     self.count += a.shape[0]; self.mom2 += a.t().mm(a)                 util/runningstats.py:492-493
 * so R ranks that each visited subset[r::R] (the subset of util/runningstats.py:1551-1556) meet in ONE reduction per
 * layer.  emcid_mom2_reduce sums the handles of all ranks of `nccl_comm` (an ncclComm_t, passed as void*) onto `root`:
 * the lower triangle of the per-rank sums travels as d (d + 1) / 2 fp32 values (18.9 MB at d = 3072, ncclReduce) next
 * to the int64 counts; afterwards the root's handle holds the job-wide statistics (emcid_mom2_finalize there returns
 * them mirrored), the other ranks' handles are unchanged.  Every rank calls it, in the same order for the same layers,
 * on a stream that no other operation of that communicator is running on concurrently.  NCCL is resolved at run time
 * from the process image (libnccl.so.2); emcid_nccl_available() tells whether that worked.
 * emcid_mom2_broadcast: the finalized [d x d] fp32 matrix and the count from the root to every rank, for callers that
 * want the statistics everywhere (the reference call returns them to its single caller). */
int emcid_nccl_available(void);
int emcid_mom2_reduce(emcid_mom2_t* h, void* nccl_comm, int root, void* stream);
int emcid_mom2_broadcast(float* mom2_full, long long* count_dev, int d, void* nccl_comm, int root, void* stream);

/* ---- accumulator state of a resumable pass -----------------------------------------------------------
 * The reference saves a statistic only once its loader is exhausted (util/runningstats.py:115-119): a crash loses the
 * pass.  export: fold, then the packed lower triangle of the fp64 sums (emcid_mom2_state_elems(d) = d (d + 1) / 2
 * doubles, row-major: (i, j), j <= i, at i (i + 1) / 2 + j) and the count, to caller-owned DEVICE buffers;
 * import: the inverse, on a fresh or reset handle. */
size_t emcid_mom2_state_elems(int d);
int emcid_mom2_export_state(emcid_mom2_t* h, double* lower_packed_dev, long long* count_dev, void* stream);
int emcid_mom2_import_state(emcid_mom2_t* h, const double* lower_packed_dev, const long long* count_dev, void* stream);
/* C[i][j] = C[j][i] for j > i, in place (row pitch ldc floats): the full symmetric matrix SecondMoment.state_dict() /
 * moment() present (util/runningstats.py:499-507) after lower-triangle accumulation with emcid_gemm3x_nt(flags bit0). */
int emcid_symmetrize_lower(float* C, int d, long long ldc, void* stream);

/* out[i] = wrapping sum of the 32-bit words of device tensor i, one launch for n tensors.  table_dev: device array of n
 * {const void* ptr; long long words;} pairs (16 bytes each); out_dev: n uint64 on the device.  Used to notice weight
 * writes that bypass the host framework's version counters before the key extraction reuses uploaded weights. */
int emcid_checksum_tensors(const void* table_dev, int n, unsigned long long* out_dev, void* stream);

/* ---- which captions a pass visits (host) --------------------------------------------------------------
 * out[0 .. n_out) = random.Random(seed).shuffle(list(range(n_items)))[:n_out], the FixedRandomSubsetSampler of
 * util/runningstats.py:1551-1556 (CPython's MT19937 + shuffle restated in C: the interpreted shuffle of an 800 k-caption
 * index costs 0.26 s per rank).  n_items < 2^31. */
int emcid_fixed_random_subset(long long n_items, long long seed, long long* out, long long n_out);

/* ---- native text-encoder forward for the statistics pass --------------------------------------------
 * Replaces the HF forward the reference runs under `Trace(...)` for every sub-batch
 *     with Trace(model, layer_name, retain_input=True, retain_output=False, stop=True) as tr:
 *         model(**batch)                                                emcid/layer_stats.py:210-216
 * (transformers CLIPTextModel: embeddings + causal pre-LN transformer layers) with the same fp32-class
 * arithmetic on the 3xFP16 tcgen05 GEMM, over PACKED valid tokens (right padding + causal attention:
 * a valid token never attends a pad token, so dropping pad rows changes nothing), and feeds act(fc1)
 * of the edited layers straight into their emcid_mom2 accumulators (mom2 += f^T f, count += tokens).
 * The forward stops after fc1 of the deepest edited layer (Trace(stop=True), util/nethook.py:112-113).
 *
 * ids / positions: [n_tokens] int32 packed token ids and position ids; cu_seqlens: [n_captions + 1]
 * int32 prefix sums of the caption lengths (each <= max_positions).  All device pointers.
 * stat_layers: ascending edited layer indices (host array) with one accumulator each (same d, same
 * device); n_stat = 0 runs `n_layers` full layers instead.  hidden_out (optional): [n_tokens x hidden]
 * fp32 residual stream after the last executed full layer (parity testing against hidden_states).
 * set_layer tensors16 (device fp32, contiguous): ln1.w ln1.b q.w q.b k.w k.b v.w v.b out.w out.b
 * ln2.w ln2.b fc1.w fc1.b fc2.w fc2.b of one encoder layer; weights are copied/split at call time. */
typedef struct emcid_clip emcid_clip_t;
int emcid_clip_create(emcid_clip_t** out, int device, int n_layers, int hidden, int heads, int intermediate, int act,
                      int max_positions, int vocab, float ln_eps, long long max_tokens, int max_captions);
int emcid_clip_set_embeddings(emcid_clip_t* h, const float* token_embedding, const float* position_embedding, void* stream);
int emcid_clip_set_layer(emcid_clip_t* h, int layer, const float* const* tensors16, void* stream);
/* Same, for a layer that was uploaded before: bit i of changed_mask set = tensors16[i] differs from the last upload; only
 * the operand planes that depend on a changed tensor are rebuilt (the edit loop changes fc2.weight of one layer between two
 * key extractions, emcid/emcid_main.py:1061). */
int emcid_clip_update_layer(emcid_clip_t* h, int layer, const float* const* tensors16, unsigned int changed_mask, void* stream);
int emcid_clip_forward(emcid_clip_t* h, const int32_t* ids, const int32_t* positions, const int32_t* cu_seqlens,
                       int n_captions, int n_tokens, int n_layers, int n_stat, const int* stat_layers,
                       emcid_mom2_t* const* accs, float* hidden_out, void* stream);
/* Key extraction for the update (replaces the two traced HF forwards per edited layer of
 *     get_module_input_output_at_words(...)                       emcid/compute_z.py:2252-2327, called at
 *     emcid/emcid_main.py:987-996 (keys) and :1004-1014 (current outputs)):
 * runs full layers [0, layer), then layer `layer` up to act(fc1), and returns for the packed token rows `key_rows`
 * (device int32 [n_keys], the last-subject-token row of every prompt) the fc2 INPUT k_out [n_keys x intermediate] and the
 * fc2 OUTPUT z_out [n_keys x hidden] (fc2.weight x + fc2.bias, computed on the gathered rows only), both fp32 device.
 * resume_layer = -1 starts from the embeddings.  resume_layer = r (< layer) continues from the state the previous call on
 * this handle left behind — a keys call at layer r over the SAME packed tokens, after which only fc2 of layer r was
 * re-uploaded (what the edit loop does between layers, emcid/emcid_main.py:1061): layer r is finished with its new fc2
 * and layers (r, layer] follow, so an L-layer edit costs the layers below the first edited one once plus about one
 * layer per edited layer instead of first + ... + last.  resume_layer = layer: the same layer again with other key rows
 * — only the gather and fc2 of the gathered rows run (a caller launches the forward first, with any valid row, and works
 * out which rows it wants while the device is busy).  Anything else is EMCID_ERR_INVALID. */
int emcid_clip_forward_keys(emcid_clip_t* h, const int32_t* ids, const int32_t* positions, const int32_t* cu_seqlens,
                            int n_captions, int n_tokens, int layer, const int32_t* key_rows, int n_keys, float* k_out,
                            float* z_out, int resume_layer, void* stream);
/* The text encoder's output for the UNet cross-attention K/V modules.  last_hidden_state =
 * final_layer_norm(residual stream after all layers) is the common input of every attn2.to_k / attn2.to_v, so ONE pass
 * serves all of them where the reference runs one pass per module (emcid/layer_stats.py:333-427, :429-467):
 *     text_repr = pipe.text_encoder(**batch).last_hidden_state; pipe.unet(latents, t, encoder_hidden_states=text_repr)
 *     feats = flatten_masked_batch(tr.input, batch["attention_mask"]); stat.add(feats)
 * and the key extraction of the cross-attention edit reads the same tensor at the last subject token
 * (emcid/compute_ks.py:91-124).  set_final_norm: text_model.final_layer_norm weight / bias [hidden], device fp32.
 * forward_final: acc (optional; an emcid_mom2 handle with d == hidden): mom2 += y^T y, count += n_tokens;
 * rows / n_rows / out (optional): fp32 last_hidden_state of the packed token rows `rows` (device int32; NULL = the
 * first n_rows tokens in order) -> out [n_rows x hidden]. */
int emcid_clip_set_final_norm(emcid_clip_t* h, const float* weight, const float* bias, void* stream);
int emcid_clip_forward_final(emcid_clip_t* h, const int32_t* ids, const int32_t* positions, const int32_t* cu_seqlens,
                             int n_captions, int n_tokens, emcid_mom2_t* acc, const int32_t* rows, int n_rows, float* out,
                             void* stream);
/* Measurement aid: when enabled, every launch of the forward is bracketed by CUDA events on the caller's stream.
 * emcid_clip_get_profile waits for them and fills out21[tag * 3 + {0, 1, 2}] = {launches, total ms, total algorithmic
 * flops (2 M N K)} for tag = 0 q/k/v projection, 1 out projection, 2 fc1, 3 fc1 of an edited layer (both plane
 * orientations), 4 fc2, 5 attention, 6 layer norm; then clears the record. */
int emcid_clip_profile(emcid_clip_t* h, int enable);
int emcid_clip_get_profile(emcid_clip_t* h, double* out21);
long long emcid_clip_launches(emcid_clip_t* h);
/* Handles park their device buffers in a per-device free list when destroyed (cudaFree synchronises the device and
 * costs seconds for the ~140 buffers of one pass); the next handle of the same shape reuses them.  This returns the
 * parked buffers to the driver. */
int emcid_release_cached_memory(void);
int emcid_clip_destroy(emcid_clip_t* h);

/* ---- closed-form multi-layer update ---------------------------------------------------------------
 * Replaces the solve block of execute_emcid_text_encoder, emcid/emcid_main.py:1037-1050 (byte-identical
 * copies at :1265-1312, :1365-1414 for SDXL and :2016-2042 in cal_insert_deltas):
 *     M      = lambda * C32.double() + Ks @ Ks.T          Ks = K.double() * scale
 *     adj_k  = torch.linalg.solve(M, Ks)                  [d, n] fp64
 *     resid  = (S.double() * scale) / (L - i)             [h, n] fp64
 *     upd    = resid @ adj_k.T                            [h, d] fp64 -> dW = upd.float()
 * `batch` independent layers are solved in one call (stacked inputs/outputs).  Per layer b:
 *   C32 [d, d] fp32  = cov * (1 - edit_weight) / 0.5 as the reference forms it in fp32 (:1037)
 *   Kt  [n, d] fp32 (row pitch ldk)  = layer_ks BEFORE the "rq c_i -> c_i rq" rearrange (:993-996)
 *   St  [n, h] fp32 (row pitch lds)  = (zs - cur_zs)^T  (:1016)
 *   scale = sqrt(edit_weight / 0.5)  (:1041-1042);  inv_layers_left[b] = 1 / (L - i) (host array)
 * refine_steps: fp64-residual iterative-refinement sweeps after the fp32-class Cholesky solve; -1 = adaptive
 * (sweeps until the error predicted from the contraction of the corrections is < 2e-5 of the solution in Frobenius
 * norm, a fifth of the dW tolerance; at most 16; synchronises the stream per sweep; EMCID_SOLVE_TOL overrides the target).
 * status_dev: device int, 0 on success, bit0 = a pivot was not positive (matrix not SPD), bit1 = the adaptive refinement
 * used all its sweeps without reaching the target (the outputs are the best available).
 * d must be a multiple of 128.  Blackwell-only; asynchronous on `stream`. */
size_t emcid_solve_workspace_bytes(int batch, int d, int h, int n);
int emcid_solve_layers(int device, int batch, int d, int h, int n, const float* C32, const float* Kt,
                       long long ldk, const float* St, long long lds, double lambda, double scale,
                       const double* inv_layers_left, double* adj_k, double* resid, float* dW,
                       int refine_steps, void* workspace, size_t workspace_bytes, int* status_dev, void* stream);

/* The v* cache files of an edit (host; emcid/emcid_main.py:873-907 reads them one by one with numpy.load):
 * out[i * elems .. (i + 1) * elems) = the float32 array `key` of paths[i], for archives whose first member is the stored
 * `<key>.npy` (what numpy.savez(file, v_star=...) writes).  EMCID_ERR_INVALID with *bad = index of the first file that is
 * missing or laid out differently (the caller falls back to a general reader); *bad = -1 on success.  Holds no
 * interpreter lock: bindings call it from a helper thread while the prompts are being tokenised. */
int emcid_read_npz_f32(const char* const* paths, int n, const char* key, float* out, long long elems, int* bad);

/* dW [h x d] fp32 = float(resid [h x n] * adj_k [d x n]^T), both fp64 row-major on the device: the product
 * apply_emcid_to_text_encoder forms from the deltas (emcid/emcid_main.py:802-809: `upd = adj_k @ resid.T` in fp64, then
 * `.float()`, transposed to the weight's shape) on the DMMA GEMM of the solver. */
int emcid_delta_update(int h, int d, int n, const double* resid, const double* adj_k, float* dW, void* stream);

/* ---- cached factorisation: repeated edits with the same covariance ------------------------------------
 * The reference re-runs torch.linalg.solve on a fresh d x d matrix for every edit even when lambda * C has not changed:
 * sequential editing (experiments/sequential_editing.py:98-171 -> emcid_main.py:1037-1050 once per edit and layer), the
 * debias factor search (emcid_main.py:1460-1472 -> cal_insert_deltas :1969-2052, up to max_iter re-solves per concept)
 * and the layer ablation (experiments/ablation.py:332-338).  emcid_factor_create factors A = lambda * C32 once
 * (blocked Cholesky + explicit inverse of the factor, kept on the device with A in fp64: 20 d^2 bytes) and
 * emcid_factor_solve returns the same adj_k / resid / dW as emcid_solve_layers(batch = 1) through
 *     (A + Ks Ks^T)^-1 Ks = Y (I + Ks^T Y)^-1,   Y = A^-1 Ks
 * with fp64-residual refinement of both solves: O(d^2 n) work per edit instead of O(d^3).
 * C32, lambda, Kt, St, scale, inv_layers_left (= 1 / (L - i)), outputs, refine_steps, status_dev: as above.
 * emcid_factor_create synchronises `stream` before it returns; emcid_factor_solve is asynchronous except for the
 * adaptive refinement's per-sweep synchronisation.  Handles are not thread-safe; distinct handles are. */
typedef struct emcid_factor emcid_factor_t;
int emcid_factor_create(emcid_factor_t** out, int device, int d, const float* C32, double lambda, int* status_dev,
                        void* stream);
size_t emcid_factor_solve_workspace_bytes(int d, int h, int n);
int emcid_factor_solve(emcid_factor_t* f, int h, int n, const float* Kt, long long ldk, const float* St, long long lds,
                       double scale, double inv_layers_left, double* adj_k, double* resid, float* dW, int refine_steps,
                       void* workspace, size_t workspace_bytes, int* status_dev, void* stream);
int emcid_factor_destroy(emcid_factor_t* f);

#ifdef __cplusplus
}
#endif
#endif /* EMCID_B200_H_ */

// emcid_b200 — the second-moment statistics pass (K1): fc1 -> activation -> pad mask -> SYRK.
//
// Replaces, for one edited CLIP text-encoder layer, the reference's
//     feats = flatten_masked_batch(tr.input, attention_mask)        dsets/stat_dataset.py:166-172
//     stat.add(feats):  count += T;  mom2 += feats.t().mm(feats)    util/runningstats.py:483-493
// where tr.input = act(fc1(LN2(h))) is recomputed here from X = LN2(h) (the argument of CLIPMLP)
// so the d-wide activations never exist in HBM-sized form:
//
//   per slab of <= slab_tokens input rows (all launches stream-ordered, no host sync):
//     1. scan      valid[T] -> destination row of every kept token, n_valid, count += n_valid
//     2. gather    X rows of kept tokens -> compacted tf32 planes Xhi/Xlo [slab x h]
//     3. fc1 GEMM  A^T = act(W1 X^T + b1) on tcgen05, epilogue masks columns >= n_valid to zero and
//                  writes the tf32 planes of A^T [d x slab]  (L2-resident: ~38 MB for d=3072)
//     4. SYRK      acc32[lower tiles] += A^T (A^T)^T, stream-K over (tile, token block), tcgen05
//   every fold_every slabs: acc64 += acc32; acc32 = 0      (bounds the fp32 summation chain)
//   finalize: full[i][j] = full[j][i] = (float) acc64[max(i,j)][min(i,j)]
#pragma once

#include <vector>

#include "host.cuh"

namespace emcid {

constexpr int MOM2_MAX_SLAB = 4096;
constexpr int MOM2_DEFAULT_SLAB = 4096;
// Default tokens per fc1/SYRK launch pair: the A^T planes of one slab (d x slab x 4 B) should stay L2 resident.
inline int mom2_default_slab(int d) {
  long long s = (48ll << 20) / (4ll * d) / 256 * 256;
  if (s > MOM2_MAX_SLAB) s = MOM2_MAX_SLAB;
  if (s < 1024) s = 1024;
  return static_cast<int>(s);
}
constexpr int MOM2_FOLD_EVERY = 8;
constexpr int MOM2_DEFAULT_KIND = KIND_F16;   // 3xFP16: same accuracy as 3xTF32 on B200, twice the MMA rate

struct Mom2Handle {
  int device, d, h, hp, act, slab;
  float w_scale;       // exact power-of-two pre-scale of the fp16-split W1 planes (1 for tf32)
  int kind, lo_fmt;    // KIND_TF32 (3xTF32) or KIND_F16 (fp16 hi + bf16/fp16 lo planes, kind::f16 MMAs)
  void* workspace;
  int chunk_fc1, chunk_syrk;
  // private state (library-allocated)
  float* w_hi; float* w_lo; float* bias;
  float* acc32;        // [d x d], lower tiles live
  double* acc64;       // [d x d], lower tiles live
  long long* count;    // device scalar
  int slabs_since_fold;
  bool has_weights;
  // shared scratch (caller-provided)
  float* x_hi; float* x_lo; float* at_hi; float* at_lo;
  int* dest; int* n_valid;
  GemmOperands fc1_ops, syrk_ops;
  DeviceInfo info;
  // optional per-launch timing (bench.py's roofline block): event pairs around fc1 / SYRK launches
  bool profile;
  std::vector<cudaEvent_t>* ev_fc1;   // begin,end,begin,end,...
  std::vector<cudaEvent_t>* ev_syrk;
  double rows_fc1, rows_syrk;         // slab rows covered by the timed launches
  long long launches;                 // every kernel this handle has launched
  float* packed;                      // lower-packed fp32 staging buffer of the exchange step (reduce.cuh), lazily allocated
};

inline size_t mom2_workspace_bytes(int d, int h, int slab) {
  const long long hp = round_up_ll(h, 64);
  long long bytes = 0;
  bytes += 2ll * slab * hp * sizeof(float);   // X planes
  bytes += 2ll * d * slab * sizeof(float);    // A^T planes
  bytes += static_cast<long long>(slab) * sizeof(int) + 256;  // dest + n_valid
  return static_cast<size_t>(bytes + 8192);
}

// ---- kernels ------------------------------------------------------------------------------------

// One block: exclusive scan of the keep flags of up to MOM2_MAX_SLAB rows.
__global__ void __launch_bounds__(1024) mom2_scan_kernel(const uint8_t* __restrict__ valid, int T,
                                                         int* __restrict__ dest, int* __restrict__ n_valid,
                                                         long long* __restrict__ count) {
  __shared__ int warp_sum[32];
  __shared__ int base_s, pass_total_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) base_s = 0;
  __syncthreads();
  for (int start = 0; start < T; start += 1024) {
    const int i = start + tid;
    const int keep = (i < T) ? (valid ? (valid[i] != 0) : 1) : 0;
    int incl = keep;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int w = warp_sum[lane];
      int wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += n;
      }
      warp_sum[lane] = wi - w;  // exclusive offset of each warp
      if (lane == 31) pass_total_s = wi;
    }
    __syncthreads();
    const int excl = base_s + warp_sum[warp] + incl - keep;
    if (i < T) dest[i] = keep ? excl : -1;
    __syncthreads();
    if (tid == 0) base_s += pass_total_s;
    __syncthreads();
  }
  if (tid == 0) {
    *n_valid = base_s;
    *count += base_s;
  }
}

// Same for 16-bit planes (KIND_F16).
__global__ void mom2_gather_split16_kernel(const float* __restrict__ X, long long ldx, int T, int h, int hp,
                                           const int* __restrict__ dest, uint16_t* __restrict__ x_hi,
                                           uint16_t* __restrict__ x_lo, int lo_fmt) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < T; r += gridDim.x * warps_per_block) {
    const int dst = dest[r];
    if (dst < 0) continue;
    const float* src = X + static_cast<long long>(r) * ldx;
    uint16_t* oh = x_hi + static_cast<long long>(dst) * hp;
    uint16_t* ol = x_lo + static_cast<long long>(dst) * hp;
    for (int c = lane * 4; c < hp; c += 128) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c + 3 < h) {
        v = *reinterpret_cast<const float4*>(src + c);
      } else {
        if (c + 0 < h) v.x = src[c + 0];
        if (c + 1 < h) v.y = src[c + 1];
        if (c + 2 < h) v.z = src[c + 2];
      }
      uint16_t hh[4], ll[4];
      split_f16(v.x, lo_fmt, hh[0], ll[0]); split_f16(v.y, lo_fmt, hh[1], ll[1]);
      split_f16(v.z, lo_fmt, hh[2], ll[2]); split_f16(v.w, lo_fmt, hh[3], ll[3]);
      uint2 hv, lv;
      hv.x = hh[0] | (static_cast<uint32_t>(hh[1]) << 16); hv.y = hh[2] | (static_cast<uint32_t>(hh[3]) << 16);
      lv.x = ll[0] | (static_cast<uint32_t>(ll[1]) << 16); lv.y = ll[2] | (static_cast<uint32_t>(ll[3]) << 16);
      *reinterpret_cast<uint2*>(oh + c) = hv;
      *reinterpret_cast<uint2*>(ol + c) = lv;
    }
  }
}

// One warp per input row: kept rows are split into tf32 planes at their compacted position.
__global__ void mom2_gather_split_kernel(const float* __restrict__ X, long long ldx, int T, int h, int hp,
                                         const int* __restrict__ dest, float* __restrict__ x_hi,
                                         float* __restrict__ x_lo) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5); r < T; r += gridDim.x * warps_per_block) {
    const int dst = dest[r];
    if (dst < 0) continue;
    const float* src = X + static_cast<long long>(r) * ldx;
    float* oh = x_hi + static_cast<long long>(dst) * hp;
    float* ol = x_lo + static_cast<long long>(dst) * hp;
    for (int c = lane * 4; c < hp; c += 128) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c + 3 < h) {
        v = *reinterpret_cast<const float4*>(src + c);
      } else {
        if (c + 0 < h) v.x = src[c + 0];
        if (c + 1 < h) v.y = src[c + 1];
        if (c + 2 < h) v.z = src[c + 2];
      }
      float4 hi, lo;
      split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y);
      split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
      *reinterpret_cast<float4*>(oh + c) = hi;
      *reinterpret_cast<float4*>(ol + c) = lo;
    }
  }
}

// acc64 += acc32; acc32 = 0 over the lower 128x128 blocks.
__global__ void mom2_fold_kernel(float* __restrict__ acc32, double* __restrict__ acc64, int d) {
  const int nb = (d + 127) / 128;
  // blockIdx.x enumerates lower blocks (bi >= bj)
  int t = blockIdx.x, bj = 0;
  for (;; ++bj) {
    int c = nb - bj;
    if (t < c) break;
    t -= c;
  }
  const int bi = bj + t;
  for (int e = threadIdx.x; e < 128 * 32; e += blockDim.x) {
    const int r = bi * 128 + e / 32;
    const int c = bj * 128 + (e % 32) * 4;
    if (r < d && c < d) {
      float4* p = reinterpret_cast<float4*>(acc32 + static_cast<long long>(r) * d + c);
      const float4 v = *p;
      double* q = acc64 + static_cast<long long>(r) * d + c;
      q[0] += v.x; q[1] += v.y; q[2] += v.z; q[3] += v.w;
      *p = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// out[i][j] = out[j][i] = (float) acc64[i][j] for i >= j.
__global__ void mom2_mirror_kernel(const double* __restrict__ acc64, float* __restrict__ out, int d) {
  __shared__ float tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int r = bi * 32 + k, c = bj * 32 + tx;
    float v = 0.f;
    if (r < d && c < d) {
      const long long idx = (r >= c) ? static_cast<long long>(r) * d + c : static_cast<long long>(c) * d + r;
      v = static_cast<float>(acc64[idx]);
      out[static_cast<long long>(r) * d + c] = v;
    }
    tile[k][tx] = v;
  }
  __syncthreads();
  if (bi != bj) {
    for (int k = ty; k < 32; k += 8) {
      const int r = bj * 32 + k, c = bi * 32 + tx;  // transposed block
      if (r < d && c < d) out[static_cast<long long>(r) * d + c] = tile[tx][k];
    }
  }
}

// ---- host ---------------------------------------------------------------------------------------

// (Re)carves the shared scratch and encodes the TMA maps for the handle's operand kind.  The scratch is
// sized for fp32 planes (mom2_workspace_bytes), so the 16-bit layout always fits.
inline int mom2_configure(Mom2Handle* H) {
  const int eb = H->kind == KIND_F16 ? 2 : 4;
  H->hp = static_cast<int>(round_up_ll(H->h, KIND_F16 == H->kind ? 64 : GEMM_BLOCK_K));
  uint8_t* w = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(H->workspace) + 1023) & ~static_cast<uintptr_t>(1023));
  const size_t xb = (static_cast<size_t>(H->slab) * H->hp * eb + 1023) & ~static_cast<size_t>(1023);
  const size_t ab = (static_cast<size_t>(H->d) * H->slab * eb + 1023) & ~static_cast<size_t>(1023);
  H->x_hi = reinterpret_cast<float*>(w); w += xb;
  H->x_lo = reinterpret_cast<float*>(w); w += xb;
  H->at_hi = reinterpret_cast<float*>(w); w += ab;
  H->at_lo = reinterpret_cast<float*>(w); w += ab;
  H->dest = reinterpret_cast<int*>(w); w += static_cast<size_t>(H->slab) * sizeof(int);
  H->n_valid = reinterpret_cast<int*>(w);
  int rc;
  if ((rc = make_tmap_2d(&H->fc1_ops.a_hi, H->w_hi, H->d, H->h, H->hp, 128, eb)) ||
      (rc = make_tmap_2d(&H->fc1_ops.a_lo, H->w_lo, H->d, H->h, H->hp, 128, eb)) ||
      (rc = make_tmap_2d(&H->fc1_ops.b_hi, H->x_hi, H->slab, H->h, H->hp, 128, eb)) ||
      (rc = make_tmap_2d(&H->fc1_ops.b_lo, H->x_lo, H->slab, H->h, H->hp, 128, eb)) ||
      (rc = make_tmap_2d(&H->syrk_ops.a_hi, H->at_hi, H->d, H->slab, H->slab, 128, eb)) ||
      (rc = make_tmap_2d(&H->syrk_ops.a_lo, H->at_lo, H->d, H->slab, H->slab, 128, eb)))
    return rc;
  H->syrk_ops.b_hi = H->syrk_ops.a_hi;
  H->syrk_ops.b_lo = H->syrk_ops.a_lo;
  return EMCID_OK;
}

inline int mom2_create(Mom2Handle** out, int device, int d, int h, int act, int slab, void* workspace,
                       size_t ws_bytes) {
  EMCID_CHECK(out != nullptr, EMCID_ERR_INVALID, "mom2_create: null out");
  EMCID_CHECK(d > 0 && h > 0 && d % 4 == 0, EMCID_ERR_INVALID, "mom2_create: d must be a positive multiple of 4");
  EMCID_CHECK(act == ACT_QUICK_GELU || act == ACT_GELU_ERF || act == ACT_NONE, EMCID_ERR_INVALID,
              "mom2_create: unknown activation %d", act);
  if (slab <= 0) slab = mom2_default_slab(d);
  EMCID_CHECK(slab % 256 == 0 && slab <= MOM2_MAX_SLAB, EMCID_ERR_INVALID,
              "mom2_create: slab_tokens must be a multiple of 256 and <= %d", MOM2_MAX_SLAB);
  EMCID_CHECK(ws_bytes >= mom2_workspace_bytes(d, h, slab), EMCID_ERR_WORKSPACE,
              "mom2_create: workspace too small (%zu < %zu)", ws_bytes, mom2_workspace_bytes(d, h, slab));
  EMCID_CUDA_CHECK(cudaSetDevice(device));
  Mom2Handle* H = new Mom2Handle();
  memset(H, 0, sizeof(*H));
  H->ev_fc1 = new std::vector<cudaEvent_t>();
  H->ev_syrk = new std::vector<cudaEvent_t>();
  int rc = get_device_info(&H->info);
  if (rc) { delete H; return rc; }
  H->device = device; H->d = d; H->h = h; H->act = act; H->slab = slab;
  H->kind = MOM2_DEFAULT_KIND; H->lo_fmt = FMT_F16; H->workspace = workspace;
  H->hp = static_cast<int>(round_up_ll(h, 64));
  H->chunk_fc1 = 1; H->chunk_syrk = GEMM_DEFAULT_CHUNK;  // fc1 rounding bias counts twice in mom2
  const size_t wbytes = static_cast<size_t>(d) * H->hp * sizeof(float);   // sized for either kind
  const size_t dd = static_cast<size_t>(d) * d;
#define EMCID_TRY_ALLOC(ptr, bytes)                                                          \
  do {                                                                                       \
    cudaError_t _e = dev_alloc(reinterpret_cast<void**>(&(ptr)), (bytes));                   \
    if (_e != cudaSuccess) {                                                                 \
      set_error(EMCID_ERR_CUDA, "mom2_create: cudaMalloc(%zu) failed: %s", (size_t)(bytes), \
                cudaGetErrorString(_e));                                                     \
      goto fail;                                                                             \
    }                                                                                        \
  } while (0)
  EMCID_TRY_ALLOC(H->w_hi, wbytes);
  EMCID_TRY_ALLOC(H->w_lo, wbytes);
  EMCID_TRY_ALLOC(H->bias, d * sizeof(float));
  EMCID_TRY_ALLOC(H->acc32, dd * sizeof(float));
  EMCID_TRY_ALLOC(H->acc64, dd * sizeof(double));
  EMCID_TRY_ALLOC(H->count, sizeof(long long));
#undef EMCID_TRY_ALLOC
  if (cudaMemset(H->acc32, 0, dd * sizeof(float)) != cudaSuccess ||
      cudaMemset(H->acc64, 0, dd * sizeof(double)) != cudaSuccess ||
      cudaMemset(H->count, 0, sizeof(long long)) != cudaSuccess ||
      cudaMemset(H->bias, 0, d * sizeof(float)) != cudaSuccess) {
    set_error(EMCID_ERR_CUDA, "mom2_create: cudaMemset failed");
    goto fail;
  }
  // stale-but-finite contents are fine everywhere except NaN/Inf patterns: start from zeros
  if (cudaMemset(workspace, 0, ws_bytes) != cudaSuccess) {
    set_error(EMCID_ERR_CUDA, "mom2_create: cudaMemset(workspace) failed");
    goto fail;
  }
  if ((rc = mom2_configure(H))) goto fail_rc;
  *out = H;
  return EMCID_OK;
fail:
  rc = EMCID_ERR_CUDA;
fail_rc:
  dev_free(H->w_hi); dev_free(H->w_lo); dev_free(H->bias); dev_free(H->acc32); dev_free(H->acc64);
  dev_free(H->count);
  delete H->ev_fc1; delete H->ev_syrk;
  delete H;
  return rc;
}

inline int mom2_destroy(Mom2Handle* H) {
  if (!H) return EMCID_OK;
  cudaSetDevice(H->device);
  dev_free(H->w_hi); dev_free(H->w_lo); dev_free(H->bias); dev_free(H->acc32); dev_free(H->acc64);
  dev_free(H->count); dev_free(H->packed);
  for (cudaEvent_t e : *H->ev_fc1) cudaEventDestroy(e);
  for (cudaEvent_t e : *H->ev_syrk) cudaEventDestroy(e);
  delete H->ev_fc1; delete H->ev_syrk;
  delete H;
  return EMCID_OK;
}

inline int mom2_set_weights(Mom2Handle* H, const float* W1, long long ldw, const float* b1, cudaStream_t stream) {
  EMCID_CHECK(H && W1, EMCID_ERR_INVALID, "mom2_set_weights: null argument");
  EMCID_CUDA_CHECK(cudaSetDevice(H->device));
  int rc;
  H->w_scale = 1.0f;
  if (H->kind == KIND_F16) {
    // H->n_valid doubles as the 4-byte reduction scratch (no slab is in flight while weights change)
    rc = f16_prescale(W1, ldw, H->d, H->h, reinterpret_cast<unsigned int*>(H->n_valid), &H->w_scale, stream);
    if (rc) return rc;
    rc = launch_split_planes16(W1, ldw, H->d, H->h, H->w_scale, H->w_hi, H->w_lo, H->hp, H->lo_fmt, stream);
  } else {
    rc = launch_split_planes(W1, ldw, H->d, H->h, 1.0f, H->w_hi, H->w_lo, H->hp, stream);
  }
  if (rc) return rc;
  if (b1) {
    EMCID_CUDA_CHECK(cudaMemcpyAsync(H->bias, b1, H->d * sizeof(float), cudaMemcpyDeviceToDevice, stream));
  } else {
    EMCID_CUDA_CHECK(cudaMemsetAsync(H->bias, 0, H->d * sizeof(float), stream));
  }
  H->has_weights = true;
  return EMCID_OK;
}

// precision: 0 = 3xTF32, 1 = 3xFP16 (fp16 hi + fp16 lo planes).  Must precede set_weights.
inline int mom2_set_precision(Mom2Handle* H, int precision) {
  EMCID_CHECK(H && precision >= 0 && precision <= 1, EMCID_ERR_INVALID, "mom2_set_precision: bad argument");
  EMCID_CHECK(H->slabs_since_fold == 0, EMCID_ERR_INVALID, "mom2_set_precision: call before accumulating");
  H->kind = precision == 0 ? KIND_TF32 : KIND_F16;
  H->lo_fmt = FMT_F16;
  H->has_weights = false;
  return mom2_configure(H);
}

inline int mom2_fold(Mom2Handle* H, cudaStream_t stream) {
  const int nb = (H->d + 127) / 128;
  mom2_fold_kernel<<<nb * (nb + 1) / 2, 256, 0, stream>>>(H->acc32, H->acc64, H->d);
  EMCID_CUDA_CHECK(cudaGetLastError());
  H->launches += 1;
  H->slabs_since_fold = 0;
  return EMCID_OK;
}

// acc32[lower tiles] += P P^T over token columns [col0, col0 + t) of the A^T planes `ops` describes
// ([d x tokens], K-major over tokens); dyn_k optionally caps t by a device scalar.  Stream-K over
// (tile, token block), red.add epilogue.  Also used by the native text-encoder forward (clip.cuh).
// EMCID_DETERMINISTIC=1: every output tile is accumulated by ONE CTA (pair) over the whole token range, in a fixed
// order, so a pass is reproducible to the bit (the default schedules split the token range of some tiles over several
// CTAs whose red.add's land in arrival order: results differ in the last bits from run to run).  Costs SYRK balance:
// 78 pair tiles on 74 CTA pairs take two rounds instead of 1.05.
inline bool mom2_deterministic() {
  const char* e = getenv("EMCID_DETERMINISTIC");
  return e && e[0] == '1';
}

// count_add: added to the handle's token count by the launch itself
inline int mom2_syrk_slab(Mom2Handle* H, const GemmOperands& ops, int kind, int col0, int t, const int* dyn_k,
                          cudaStream_t stream, int streamk_mode = 1, long long count_add = 0) {
  if (mom2_deterministic()) streamk_mode = 0;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  if (count_add) { p.count = H->count; p.count_add = count_add; }
  p.M = H->d; p.N = H->d; p.K = t;
  p.a_col0 = col0; p.b_col0 = col0;
  p.dyn_k = dyn_k;
  p.lower = 1; p.streamk = streamk_mode;
  p.chunk_kblocks = H->chunk_syrk;
  p.C = H->acc32; p.ldc = H->d; p.lo_fmt = H->lo_fmt;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (H->profile) {
    EMCID_CUDA_CHECK(cudaEventCreate(&e0)); EMCID_CUDA_CHECK(cudaEventCreate(&e1));
    EMCID_CUDA_CHECK(cudaEventRecord(e0, stream));
  }
  const int sms = H->info.sm_count;
  int rc = kind == KIND_F16_MN
               ? (gemm_cta2_enabled() ? launch_gemm3x<256, 3, EPI_RED, KIND_F16_MN, EF_DEFAULT, 1>(ops, p, sms & ~1, stream)
                                      : launch_gemm3x<256, 2, EPI_RED, KIND_F16_MN>(ops, p, sms, stream))
           : kind == KIND_F16
               ? (gemm_cta2_enabled() ? launch_gemm3x<256, 3, EPI_RED, KIND_F16, EF_DEFAULT, 1>(ops, p, sms & ~1, stream)
                                      : launch_gemm3x<256, 2, EPI_RED, KIND_F16>(ops, p, sms, stream))
               : launch_gemm3x<256, 2, EPI_RED>(ops, p, sms, stream);
  if (rc) return rc;
  H->launches += 1;
  if (H->profile) {
    EMCID_CUDA_CHECK(cudaEventRecord(e1, stream));
    H->ev_syrk->push_back(e0); H->ev_syrk->push_back(e1); H->rows_syrk += t;
  }
  return EMCID_OK;
}

inline int mom2_accumulate(Mom2Handle* H, const float* X, long long ldx, const uint8_t* valid, long long T,
                           cudaStream_t stream) {
  EMCID_CHECK(H && (X || T == 0), EMCID_ERR_INVALID, "mom2_accumulate: null argument");
  EMCID_CHECK(H->has_weights, EMCID_ERR_INVALID, "mom2_accumulate: call emcid_mom2_set_weights first");
  EMCID_CHECK(T >= 0 && ldx >= H->h, EMCID_ERR_INVALID, "mom2_accumulate: bad T/ldx");
  EMCID_CHECK((reinterpret_cast<uintptr_t>(X) & 15) == 0 && ldx % 4 == 0, EMCID_ERR_INVALID,
              "mom2_accumulate: X must be 16-byte aligned with a row pitch multiple of 4");
  if (T == 0) return EMCID_OK;
  EMCID_CUDA_CHECK(cudaSetDevice(H->device));
  // near-equal slabs: avoids a small ragged tail slab
  const long long nslabs = (T + H->slab - 1) / H->slab;
  long long per = (T + nslabs - 1) / nslabs;
  per = round_up_ll(per, 32);
  if (per > H->slab) per = H->slab;
  const int sms = H->info.sm_count;
  for (long long r0 = 0; r0 < T; r0 += per) {
    const int t = static_cast<int>(T - r0 < per ? T - r0 : per);
    mom2_scan_kernel<<<1, 1024, 0, stream>>>(valid ? valid + r0 : nullptr, t, H->dest, H->n_valid, H->count);
    EMCID_CUDA_CHECK(cudaGetLastError());
    H->launches += 3;  // scan, gather, fc1 (the SYRK counts itself)
    {
      int blocks = (t + 7) / 8;
      if (H->kind == KIND_F16)
        mom2_gather_split16_kernel<<<blocks, 256, 0, stream>>>(X + r0 * ldx, ldx, t, H->h, H->hp, H->dest,
                                                               reinterpret_cast<uint16_t*>(H->x_hi),
                                                               reinterpret_cast<uint16_t*>(H->x_lo), H->lo_fmt);
      else
        mom2_gather_split_kernel<<<blocks, 256, 0, stream>>>(X + r0 * ldx, ldx, t, H->h, H->hp, H->dest,
                                                             H->x_hi, H->x_lo);
      EMCID_CUDA_CHECK(cudaGetLastError());
    }
    int rc;
    {
      GemmParams p;
      memset(&p, 0, sizeof(p));
      p.M = H->d; p.N = t; p.K = H->h;
      p.dyn_n = H->n_valid;
      p.chunk_kblocks = H->chunk_fc1;
      p.P_hi = H->at_hi; p.P_lo = H->at_lo; p.ldp = H->slab;
      p.bias = H->bias; p.act = H->act; p.lo_fmt = H->lo_fmt; p.alpha = 1.0f / H->w_scale;
      const int tiles = gemm_num_tiles(H->d, t, 256, 0);
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      if (H->profile) {
        EMCID_CUDA_CHECK(cudaEventCreate(&e0)); EMCID_CUDA_CHECK(cudaEventCreate(&e1));
        EMCID_CUDA_CHECK(cudaEventRecord(e0, stream));
      }
      const int g = tiles < sms ? tiles : sms;
#define EMCID_FC1_CASE(A)                                                                            \
  if (H->act == A)                                                                                   \
    rc = H->kind == KIND_F16 ? launch_gemm3x<256, 2, EPI_FC1, KIND_F16, A>(H->fc1_ops, p, g, stream) \
                             : launch_gemm3x<256, 2, EPI_FC1, KIND_TF32, A>(H->fc1_ops, p, g, stream);
      rc = EMCID_ERR_INVALID;
      EMCID_FC1_CASE(ACT_QUICK_GELU) EMCID_FC1_CASE(ACT_GELU_ERF) EMCID_FC1_CASE(ACT_NONE)
#undef EMCID_FC1_CASE
      if (rc) return rc;
      if (H->profile) {
        EMCID_CUDA_CHECK(cudaEventRecord(e1, stream));
        H->ev_fc1->push_back(e0); H->ev_fc1->push_back(e1); H->rows_fc1 += t;
      }
    }
    rc = mom2_syrk_slab(H, H->syrk_ops, H->kind, 0, t, H->n_valid, stream);
    if (rc) return rc;
    if (++H->slabs_since_fold >= MOM2_FOLD_EVERY) {
      rc = mom2_fold(H, stream);
      if (rc) return rc;
    }
  }
  return EMCID_OK;
}

inline int mom2_finalize(Mom2Handle* H, float* mom2_full, long long* count_dev, cudaStream_t stream) {
  EMCID_CHECK(H && mom2_full, EMCID_ERR_INVALID, "mom2_finalize: null argument");
  EMCID_CUDA_CHECK(cudaSetDevice(H->device));
  int rc = mom2_fold(H, stream);
  if (rc) return rc;
  const int nb = (H->d + 31) / 32;
  mom2_mirror_kernel<<<dim3(nb, nb), dim3(32, 8), 0, stream>>>(H->acc64, mom2_full, H->d);
  EMCID_CUDA_CHECK(cudaGetLastError());
  H->launches += 1;
  if (count_dev) {
    EMCID_CUDA_CHECK(cudaMemcpyAsync(count_dev, H->count, sizeof(long long), cudaMemcpyDeviceToDevice, stream));
  }
  return EMCID_OK;
}

inline int mom2_reset(Mom2Handle* H, cudaStream_t stream) {
  EMCID_CHECK(H, EMCID_ERR_INVALID, "mom2_reset: null handle");
  EMCID_CUDA_CHECK(cudaSetDevice(H->device));
  const size_t dd = static_cast<size_t>(H->d) * H->d;
  EMCID_CUDA_CHECK(cudaMemsetAsync(H->acc32, 0, dd * sizeof(float), stream));
  EMCID_CUDA_CHECK(cudaMemsetAsync(H->acc64, 0, dd * sizeof(double), stream));
  EMCID_CUDA_CHECK(cudaMemsetAsync(H->count, 0, sizeof(long long), stream));
  H->slabs_since_fold = 0;
  return EMCID_OK;
}

// out[0..7] = {fc1 ms total, fc1 launches, fc1 rows, syrk ms total, syrk launches, syrk rows, all launches, 0};
// waits for the recorded events, then clears them.
inline int mom2_get_profile(Mom2Handle* H, double* out) {
  EMCID_CHECK(H && out, EMCID_ERR_INVALID, "mom2_get_profile: null argument");
  EMCID_CUDA_CHECK(cudaSetDevice(H->device));
  auto drain = [](std::vector<cudaEvent_t>* v, double* ms_total) -> cudaError_t {
    *ms_total = 0.0;
    for (size_t i = 0; i + 1 < v->size(); i += 2) {
      cudaError_t e = cudaEventSynchronize((*v)[i + 1]);
      if (e != cudaSuccess) return e;
      float ms = 0.f;
      e = cudaEventElapsedTime(&ms, (*v)[i], (*v)[i + 1]);
      if (e != cudaSuccess) return e;
      *ms_total += ms;
      cudaEventDestroy((*v)[i]); cudaEventDestroy((*v)[i + 1]);
    }
    return cudaSuccess;
  };
  out[1] = static_cast<double>(H->ev_fc1->size() / 2);
  out[4] = static_cast<double>(H->ev_syrk->size() / 2);
  EMCID_CUDA_CHECK(drain(H->ev_fc1, &out[0]));
  EMCID_CUDA_CHECK(drain(H->ev_syrk, &out[3]));
  H->ev_fc1->clear(); H->ev_syrk->clear();
  out[2] = H->rows_fc1; out[5] = H->rows_syrk;
  out[6] = static_cast<double>(H->launches);
  out[7] = 0.0;
  H->rows_fc1 = H->rows_syrk = 0.0;
  return EMCID_OK;
}

}  // namespace emcid

// emcid_b200 — native CLIP text-encoder forward for the statistics pass.
//
// The reference runs the HF `CLIPTextModel` in fp32 up to the traced layer for every sub-batch
// (emcid/layer_stats.py:208-219 -> transformers modeling_clip.py: embeddings, 12 x [LN1, causal
// self-attention, residual, LN2, fc1, act, fc2, residual]); on a GPU that forward is cuBLAS SIMT sgemm
// and costs 6x what the fused mom2 kernels cost.  Here the same arithmetic runs on the 3xFP16
// tcgen05 GEMM of gemm3x.cuh (fp32-class accuracy, see common.cuh::split_f16), over PACKED valid
// tokens only: with right padding and causal attention a valid token never attends a pad token
// (SURVEY.md §6), so pad rows are dropped before the first layer and `count` is the packed length.
//
//   embed            hres[t] = tok_emb[ids[t]] + pos_emb[pos[t]]                       (fp32 residual stream)
//   per layer        x  = split(LN1(hres))                      warp per token
//                    qkv = x Wqkv^T + b                          3xFP16 GEMM (q, k, v fused, N = 3h)
//                    a  = split(softmax(q k^T * dh^-1/2, causal) v)   fp32, one CTA per (caption, head)
//                    hres += a Wo^T + bo                         GEMM with residual epilogue
//                    x  = split(LN2(hres))
//                    f  = split(act(x W1^T + b1))                GEMM; edited layers also emit f^T planes
//                    [edited layer]  mom2 += f^T f               stream-K lower SYRK per L2-sized token slab
//                    hres += f W2^T + b2                         (skipped after the deepest edited layer,
//                                                                 like Trace(stop=True), util/nethook.py:112)
// Every GEMM operand is K-major 16-bit hi/lo planes; static weights are pre-scaled by an exact power
// of two (host.cuh::f16_prescale) and the scale is undone by `alpha` in the epilogue.
#pragma once

#include <stdlib.h>

#include <vector>

#include "attn.cuh"
#include "mom2.cuh"

namespace emcid {

struct ClipWeight {
  uint16_t* hi = nullptr; uint16_t* lo = nullptr;  // [N x Kp] planes
  float* bias = nullptr;                           // [N]
  float scale = 1.f;
  int N = 0, K = 0, Kp = 0;
  CUtensorMap m_hi, m_lo;
};

struct ClipLayer {
  ClipWeight qkv, o, fc1, fc2;
  float *ln1_w = nullptr, *ln1_b = nullptr, *ln2_w = nullptr, *ln2_b = nullptr;
  bool set = false;
};

struct ClipHandle {
  int device, L, h, heads, dh, d, act, max_pos, vocab;
  float eps;
  int hp, dp;                  // K extents padded to 64
  long long cap_tokens, tp;    // token capacity, padded to 256 (pitch of the transposed planes)
  int cap_seqs;
  float *tok_emb, *pos_emb;
  std::vector<ClipLayer>* layers;
  float *fln_w, *fln_b;        // final_layer_norm (text_model.final_layer_norm): input of the UNet cross-attention K/V layers
  bool fln_set;
  float* hres;                 // [cap x h]
  float* qkv;                  // [cap x 3h]   fp32 q|k|v (CUDA-core attention path)
  uint16_t *qp_hi, *qp_lo;     // [cap x 3h]   q|k|v planes (tensor-core attention path)
  bool attn_tc;
  uint16_t *x_hi, *x_lo;       // [cap x hp]   LN output
  uint16_t *a_hi, *a_lo;       // [cap x hp]   attention output
  uint16_t *f_hi, *f_lo;       // [cap x dp]   act(fc1)
  uint16_t *ft_hi, *ft_lo;     // [d x tp]     scratch planes: the gathered key rows [n_keys x dp] of the key extraction
  unsigned int* scratch;
  int *grp, *n_grp, *tok_start; // attention units: group offsets [cap_seqs + 2], their number, [cap] first token of t's caption
  std::vector<void*>* allocs;
  DeviceInfo info;
  long long launches;
  int attn_smem;
  // key-extraction continuation: the last keys call left act(fc1) of layer keys_state_layer in the f planes and the
  // residual stream after that layer's attention block in hres, for keys_state_tokens packed tokens (-1: nothing)
  int keys_state_layer;
  int keys_state_tokens;
  // measurement aid (emcid_clip_profile): CUDA events around every launch of the forward, tagged by kernel class
  bool profile;
  std::vector<cudaEvent_t>* ev;      // pairs (start, stop)
  std::vector<int>* ev_tag;          // CLIP_TAG_* of each pair
  std::vector<double>* ev_flops;     // algorithmic flops of each pair (2 M N K; 0 for memory-bound kernels)
};

// kernel classes of the profile: linear layers by role, attention, layer norm
enum ClipTag : int { CLIP_TAG_QKV = 0, CLIP_TAG_OUT = 1, CLIP_TAG_FC1 = 2, CLIP_TAG_FC1_STAT = 3, CLIP_TAG_FC2 = 4,
                     CLIP_TAG_ATTN = 5, CLIP_TAG_LN = 6, CLIP_TAG_COUNT = 7 };

struct ClipProfScope {
  ClipHandle* H; cudaStream_t stream; cudaEvent_t e0;
  ClipProfScope(ClipHandle* H_, cudaStream_t s, int tag, double flops) : H(H_), stream(s), e0(nullptr) {
    if (!H->profile) return;
    cudaEvent_t e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { e0 = nullptr; return; }
    cudaEventRecord(e0, stream);
    H->ev->push_back(e0); H->ev->push_back(e1);
    H->ev_tag->push_back(tag); H->ev_flops->push_back(flops);
  }
  ~ClipProfScope() {
    if (e0) cudaEventRecord(H->ev->back(), stream);
  }
};

// ---- kernels ------------------------------------------------------------------------------------

__global__ void clip_embed_kernel(const int* __restrict__ ids, const int* __restrict__ pos, int T, int h, int vocab,
                                  int max_pos, const float* __restrict__ tok, const float* __restrict__ pe,
                                  float* __restrict__ out) {
  const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int t = blockIdx.x * warps + (threadIdx.x >> 5); t < T; t += gridDim.x * warps) {
    int id = ids[t], p = pos[t];
    id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
    p = p < 0 ? 0 : (p >= max_pos ? max_pos - 1 : p);
    const float* a = tok + static_cast<long long>(id) * h;
    const float* b = pe + static_cast<long long>(p) * h;
    float* o = out + static_cast<long long>(t) * h;
    for (int c = lane; c < h; c += 32) o[c] = a[c] + b[c];
  }
}

// One warp per token: y = (x - mean) * rsqrt(var + eps) * w + b (biased variance, like torch.nn.LayerNorm), written as
// fp16 hi/lo planes.  Four consecutive columns per lane: float4 loads, one 8-byte store per plane (h % 4 == 0, ldo % 4 == 0;
// the scalar first version ran at 48.6 us per 39 424 tokens against 34 us).  NV4 = ceil(h / 128) float4 per lane stay in
// registers between the passes.
template <int NV4>
__global__ void clip_layernorm4_kernel(const float* __restrict__ x, int T, int h, const float* __restrict__ w,
                                       const float* __restrict__ b, float eps, uint16_t* __restrict__ o_hi,
                                       uint16_t* __restrict__ o_lo, int ldo) {
  const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
  const float inv_h = 1.0f / static_cast<float>(h);
  const int h4 = h >> 2;
  for (int t = blockIdx.x * warps + (threadIdx.x >> 5); t < T; t += gridDim.x * warps) {
    const float4* src = reinterpret_cast<const float4*>(x + static_cast<long long>(t) * h);
    float4 v[NV4];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + 32 * i;
      v[i] = c < h4 ? src[c] : make_float4(0.f, 0.f, 0.f, 0.f);
      s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_h;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      if (lane + 32 * i < h4) {
        const float d0 = v[i].x - mean, d1 = v[i].y - mean, d2 = v[i].z - mean, d3 = v[i].w - mean;
        q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * inv_h + eps);
    uint2* oh = reinterpret_cast<uint2*>(o_hi + static_cast<long long>(t) * ldo);
    uint2* ol = reinterpret_cast<uint2*>(o_lo + static_cast<long long>(t) * ldo);
    const float4* w4 = reinterpret_cast<const float4*>(w);
    const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
    for (int i = 0; i < NV4; ++i) {
      const int c = lane + 32 * i;
      if (c < h4) {
        const float4 ww = w4[c], bb = b4[c];
        uint2 hv, lv;
        split_f16x2((v[i].x - mean) * rstd * ww.x + bb.x, (v[i].y - mean) * rstd * ww.y + bb.y, hv.x, lv.x);
        split_f16x2((v[i].z - mean) * rstd * ww.z + bb.z, (v[i].w - mean) * rstd * ww.w + bb.w, hv.y, lv.y);
        oh[c] = hv;
        ol[c] = lv;
      }
    }
  }
}

// Causal softmax attention of one (caption, head) in fp32 on the CUDA cores, for head dims the tensor-core kernel
// (attn.cuh, head dim 64) does not cover: K/V of the caption staged in shared memory; qkv: [T x 3h] (q | k | v), output planes
// [T x ldo] at columns head*dh...
// Register-tiled for head dims 16/32/64/128: a warp handles TWO adjacent query rows at once, keeps both
// q rows in registers and reads K/V as float4 from shared memory (row pitch DH + 4: conflict-free LDS.128), so
// every K/V element fetched feeds two FMAs and the kernel is FMA- rather than LDS-bound.
template <int DH>
__global__ void __launch_bounds__(128) clip_attention2_kernel(const float* __restrict__ qkv, const int* __restrict__ cu,
                                                              int h, int lmax, float scale, uint16_t* __restrict__ o_hi,
                                                              uint16_t* __restrict__ o_lo, int ldo) {
  extern __shared__ float attn_sm[];
  constexpr int LDK = DH + 4;
  constexpr int NC = (DH + 31) / 32;           // output columns per lane
  const int lp = (lmax + 7) & ~3;              // padded row count / score pitch (>= lmax + 4, multiple of 4)
  float* Ks = attn_sm;                         // [lp][LDK]
  float* Vs = Ks + lp * LDK;                   // [lp][LDK]
  float* Ps = Vs + lp * LDK;                   // [4 warps][2 rows][lp]
  const int head = blockIdx.x, seq = blockIdx.y;
  const int t0 = cu[seq];
  int Ls = cu[seq + 1] - t0;
  if (Ls > lmax) Ls = lmax;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long ld = 3ll * h;
  const float* base = qkv + static_cast<long long>(t0) * ld + head * DH;
  const int Lz = (Ls + 3) & ~3;                // V rows [Ls, Lz) must read as zeros (their p is 0, 0 * garbage is not)
  for (int e = tid; e < Lz * (DH / 4); e += blockDim.x) {
    const int j = e / (DH / 4), c = (e - j * (DH / 4)) * 4;
    float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
    if (j < Ls) {
      kv = *reinterpret_cast<const float4*>(base + j * ld + h + c);
      vv = *reinterpret_cast<const float4*>(base + j * ld + 2 * h + c);
    }
    *reinterpret_cast<float4*>(Ks + j * LDK + c) = kv;
    *reinterpret_cast<float4*>(Vs + j * LDK + c) = vv;
  }
  __syncthreads();
  float* p0 = Ps + (warp * 2) * lp;
  float* p1 = p0 + lp;
  for (int r0 = 2 * warp; r0 < Ls; r0 += 8) {
    const int r1 = r0 + 1;
    const bool has1 = r1 < Ls;
    const int rl = has1 ? r1 : r0;             // last key index any of the two rows attends
    float4 q0[DH / 4], q1[DH / 4];
    {
      const float4* g0 = reinterpret_cast<const float4*>(base + r0 * ld);
      const float4* g1 = reinterpret_cast<const float4*>(base + (has1 ? r1 : r0) * ld);
#pragma unroll
      for (int c = 0; c < DH / 4; ++c) { q0[c] = g0[c]; q1[c] = g1[c]; }
    }
    float s0[4], s1[4];
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int j = lane + 32 * m;
      float d0 = -INFINITY, d1 = -INFINITY;
      if (32 * m <= rl) {                       // warp-uniform: this key group is needed at all
        if (j <= rl) {
          float a0 = 0.f, a1 = 0.f;
          const float4* kr = reinterpret_cast<const float4*>(Ks + j * LDK);
#pragma unroll
          for (int c = 0; c < DH / 4; ++c) {
            const float4 k4 = kr[c];
            a0 = fmaf(q0[c].x, k4.x, a0); a0 = fmaf(q0[c].y, k4.y, a0); a0 = fmaf(q0[c].z, k4.z, a0); a0 = fmaf(q0[c].w, k4.w, a0);
            a1 = fmaf(q1[c].x, k4.x, a1); a1 = fmaf(q1[c].y, k4.y, a1); a1 = fmaf(q1[c].z, k4.z, a1); a1 = fmaf(q1[c].w, k4.w, a1);
          }
          if (j <= r0) d0 = a0 * scale;
          d1 = a1 * scale;
        }
      }
      s0[m] = d0; s1[m] = d1;
      m0 = fmaxf(m0, d0); m1 = fmaxf(m1, d1);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, o));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, o));
    }
    float z0 = 0.f, z1 = 0.f;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int j = lane + 32 * m;
      const float e0 = j <= r0 ? expf(s0[m] - m0) : 0.f;
      const float e1 = j <= rl ? expf(s1[m] - m1) : 0.f;
      s0[m] = e0; s1[m] = e1;
      z0 += e0; z1 += e1;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      z0 += __shfl_xor_sync(0xffffffffu, z0, o);
      z1 += __shfl_xor_sync(0xffffffffu, z1, o);
    }
    const float i0 = 1.0f / z0, i1 = 1.0f / z1;
    const int jz = (rl + 4) & ~3;               // scores [0, jz) are read back below; masked ones are zeros
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int j = lane + 32 * m;
      if (j < jz) { p0[j] = s0[m] * i0; p1[j] = s1[m] * i1; }
    }
    __syncwarp();
    float acc0[NC], acc1[NC];
#pragma unroll
    for (int u = 0; u < NC; ++u) acc0[u] = acc1[u] = 0.f;
    for (int j4 = 0; j4 <= rl; j4 += 4) {
      const float4 pa = *reinterpret_cast<const float4*>(p0 + j4);
      const float4 pb = *reinterpret_cast<const float4*>(p1 + j4);
#pragma unroll
      for (int u = 0; u < NC; ++u) {
        const int c = lane + 32 * u;
        if (DH >= 32 || c < DH) {
          const float v0 = Vs[(j4 + 0) * LDK + c], v1 = Vs[(j4 + 1) * LDK + c];
          const float v2 = Vs[(j4 + 2) * LDK + c], v3 = Vs[(j4 + 3) * LDK + c];
          acc0[u] = fmaf(pa.x, v0, acc0[u]); acc0[u] = fmaf(pa.y, v1, acc0[u]);
          acc0[u] = fmaf(pa.z, v2, acc0[u]); acc0[u] = fmaf(pa.w, v3, acc0[u]);
          acc1[u] = fmaf(pb.x, v0, acc1[u]); acc1[u] = fmaf(pb.y, v1, acc1[u]);
          acc1[u] = fmaf(pb.z, v2, acc1[u]); acc1[u] = fmaf(pb.w, v3, acc1[u]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < NC; ++u) {
      const int c = lane + 32 * u;
      if (DH >= 32 || c < DH) {
        uint16_t hh, ll;
        const long long o0 = static_cast<long long>(t0 + r0) * ldo + head * DH + c;
        split_f16(acc0[u], FMT_F16, hh, ll);
        o_hi[o0] = hh; o_lo[o0] = ll;
        if (has1) {
          split_f16(acc1[u], FMT_F16, hh, ll);
          o_hi[o0 + ldo] = hh; o_lo[o0 + ldo] = ll;
        }
      }
    }
    __syncwarp();
  }
}


// ---- host ---------------------------------------------------------------------------------------

template <typename T>
inline int clip_alloc(ClipHandle* H, T** p, size_t elems) {
  void* q = nullptr;
  cudaError_t e = dev_alloc(&q, elems * sizeof(T) + 256);
  if (e != cudaSuccess)
    return set_error(EMCID_ERR_CUDA, "clip: cudaMalloc(%zu) failed: %s", elems * sizeof(T), cudaGetErrorString(e));
  H->allocs->push_back(q);
  *p = static_cast<T*>(q);
  return EMCID_OK;
}

inline int clip_destroy(ClipHandle* H) {
  if (!H) return EMCID_OK;
  cudaSetDevice(H->device);
  cudaDeviceSynchronize();
  for (void* p : *H->allocs) dev_free(p);   // parked for the next handle of this shape (host.cuh::DevPool)
  delete H->allocs;
  delete H->layers;
  if (H->ev) for (cudaEvent_t e : *H->ev) cudaEventDestroy(e);
  delete H->ev; delete H->ev_tag; delete H->ev_flops;
  delete H;
  return EMCID_OK;
}

inline int clip_create(ClipHandle** out, int device, int L, int h, int heads, int d, int act, int max_pos, int vocab,
                       float eps, long long cap_tokens, int cap_seqs) {
  EMCID_CHECK(out, EMCID_ERR_INVALID, "clip_create: null out");
  EMCID_CHECK(L > 0 && h > 0 && heads > 0 && h % heads == 0 && d > 0 && max_pos > 0 && vocab > 0 && cap_tokens > 0 &&
                  cap_seqs > 0,
              EMCID_ERR_INVALID, "clip_create: bad shape");
  EMCID_CHECK(h % 4 == 0 && d % 4 == 0, EMCID_ERR_INVALID, "clip_create: hidden and intermediate sizes must be multiples of 4");
  EMCID_CHECK(h <= 2048, EMCID_ERR_UNSUPPORTED, "clip_create: hidden size %d > 2048", h);
  EMCID_CHECK(max_pos <= 128, EMCID_ERR_UNSUPPORTED, "clip_create: max_position_embeddings %d > 128", max_pos);
  EMCID_CHECK(act == ACT_QUICK_GELU || act == ACT_GELU_ERF, EMCID_ERR_INVALID, "clip_create: unknown activation %d", act);
  EMCID_CUDA_CHECK(cudaSetDevice(device));
  ClipHandle* H = new ClipHandle();
  memset(H, 0, sizeof(*H));
  H->allocs = new std::vector<void*>();
  H->layers = new std::vector<ClipLayer>(L);
  H->ev = new std::vector<cudaEvent_t>();
  H->ev_tag = new std::vector<int>();
  H->ev_flops = new std::vector<double>();
  int rc = get_device_info(&H->info);
  if (rc) { clip_destroy(H); return rc; }
  H->device = device; H->L = L; H->h = h; H->heads = heads; H->dh = h / heads; H->d = d; H->act = act;
  H->max_pos = max_pos; H->vocab = vocab; H->eps = eps;
  H->hp = static_cast<int>(round_up_ll(h, 64));
  H->dp = static_cast<int>(round_up_ll(d, 64));
  H->cap_tokens = cap_tokens; H->cap_seqs = cap_seqs;
  H->keys_state_layer = -1; H->keys_state_tokens = 0;
  H->tp = round_up_ll(cap_tokens, 256);
  const size_t cap = static_cast<size_t>(cap_tokens);
  if ((rc = clip_alloc(H, &H->tok_emb, static_cast<size_t>(vocab) * h)) ||
      (rc = clip_alloc(H, &H->pos_emb, static_cast<size_t>(max_pos) * h)) ||
      (rc = clip_alloc(H, &H->hres, cap * h)) || (rc = clip_alloc(H, &H->qkv, cap * 3 * h)) ||
      (rc = clip_alloc(H, &H->x_hi, cap * H->hp)) || (rc = clip_alloc(H, &H->x_lo, cap * H->hp)) ||
      (rc = clip_alloc(H, &H->a_hi, cap * H->hp)) || (rc = clip_alloc(H, &H->a_lo, cap * H->hp)) ||
      (rc = clip_alloc(H, &H->f_hi, cap * H->dp)) || (rc = clip_alloc(H, &H->f_lo, cap * H->dp)) ||
      (rc = clip_alloc(H, &H->ft_hi, static_cast<size_t>(d) * H->tp)) ||
      (rc = clip_alloc(H, &H->ft_lo, static_cast<size_t>(d) * H->tp)) || (rc = clip_alloc(H, &H->scratch, 64)) ||
      (rc = clip_alloc(H, &H->qp_hi, cap * 3 * h)) || (rc = clip_alloc(H, &H->qp_lo, cap * 3 * h)) ||
      (rc = clip_alloc(H, &H->grp, static_cast<size_t>(cap_seqs) + 2)) || (rc = clip_alloc(H, &H->n_grp, 4)) ||
      (rc = clip_alloc(H, &H->tok_start, cap))) {
    clip_destroy(H);
    return rc;
  }
  // pad columns [h, hp) / [d, dp) of the activation planes are never written: they must read as zeros
  if (cudaMemset(H->x_hi, 0, cap * H->hp * 2) != cudaSuccess || cudaMemset(H->x_lo, 0, cap * H->hp * 2) != cudaSuccess ||
      cudaMemset(H->a_hi, 0, cap * H->hp * 2) != cudaSuccess || cudaMemset(H->a_lo, 0, cap * H->hp * 2) != cudaSuccess ||
      cudaMemset(H->f_hi, 0, cap * H->dp * 2) != cudaSuccess || cudaMemset(H->f_lo, 0, cap * H->dp * 2) != cudaSuccess ||
      cudaMemset(H->ft_hi, 0, static_cast<size_t>(d) * H->tp * 2) != cudaSuccess ||
      cudaMemset(H->ft_lo, 0, static_cast<size_t>(d) * H->tp * 2) != cudaSuccess) {
    clip_destroy(H);
    return set_error(EMCID_ERR_CUDA, "clip_create: cudaMemset failed");
  }
  {
    const int dh = H->dh;
    if (!(dh == 16 || dh == 32 || dh == 64 || dh == 128)) {
      clip_destroy(H);
      return set_error(EMCID_ERR_UNSUPPORTED, "clip_create: head dim %d (supported: 16, 32, 64, 128)", dh);
    }
    const int lp = (max_pos + 7) & ~3;
    H->attn_smem = (2 * lp * (dh + 4) + 8 * lp) * static_cast<int>(sizeof(float));
    cudaError_t e = cudaSuccess;
    if (H->attn_smem > 48 * 1024) {
      if (dh == 16) e = cudaFuncSetAttribute(clip_attention2_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, H->attn_smem);
      else if (dh == 32) e = cudaFuncSetAttribute(clip_attention2_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, H->attn_smem);
      else if (dh == 64) e = cudaFuncSetAttribute(clip_attention2_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, H->attn_smem);
      else e = cudaFuncSetAttribute(clip_attention2_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, H->attn_smem);
    }
    if (e != cudaSuccess) {
      clip_destroy(H);
      return set_error(EMCID_ERR_CUDA, "clip_create: attention shared memory %d B: %s", H->attn_smem, cudaGetErrorString(e));
    }
  }
  {
    // tensor-core attention: head dim 64 (CLIP-L, OpenCLIP bigG), captions up to 128 tokens; EMCID_ATTN_TC=0 disables
    const char* e = getenv("EMCID_ATTN_TC");
    H->attn_tc = H->dh == ATTN_DH && max_pos <= 128 && (3 * h) % 8 == 0 && !(e && e[0] == '0');
    if (H->attn_tc) {
      const int lp = (max_pos + 15) & ~15;
      const int smem = attn_smem_bytes(lp);
      cudaError_t ce = lp <= 80 ? cudaFuncSetAttribute(clip_attention_tc_kernel<5, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                                : cudaFuncSetAttribute(clip_attention_tc_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (ce == cudaSuccess)
        ce = lp <= 80 ? cudaFuncSetAttribute(clip_attention_tc_kernel<5, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
                      : cudaFuncSetAttribute(clip_attention_tc_kernel<8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      if (ce != cudaSuccess) {
        clip_destroy(H);
        return set_error(EMCID_ERR_CUDA, "clip_create: attention shared memory %d B: %s", attn_smem_bytes(lp), cudaGetErrorString(ce));
      }
    }
  }
  *out = H;
  return EMCID_OK;
}

inline int clip_set_embeddings(ClipHandle* H, const float* tok, const float* pos, cudaStream_t stream) {
  EMCID_CHECK(H && tok && pos, EMCID_ERR_INVALID, "clip_set_embeddings: null argument");
  EMCID_CUDA_CHECK(cudaSetDevice(H->device));
  EMCID_CUDA_CHECK(cudaMemcpyAsync(H->tok_emb, tok, static_cast<size_t>(H->vocab) * H->h * 4, cudaMemcpyDeviceToDevice, stream));
  EMCID_CUDA_CHECK(cudaMemcpyAsync(H->pos_emb, pos, static_cast<size_t>(H->max_pos) * H->h * 4, cudaMemcpyDeviceToDevice, stream));
  return EMCID_OK;
}

// Splits `parts` stacked row blocks (each [rows_each x K], row pitch K) into one [parts*rows_each x Kp] plane pair.
inline int clip_prepare_weight(ClipHandle* H, ClipWeight* W, int parts, const float* const* w, const float* const* b,
                               int rows_each, int K, cudaStream_t stream) {
  const int N = parts * rows_each;
  const int Kp = static_cast<int>(round_up_ll(K, 64));
  int rc;
  if (!W->hi) {
    if ((rc = clip_alloc(H, &W->hi, static_cast<size_t>(N) * Kp)) || (rc = clip_alloc(H, &W->lo, static_cast<size_t>(N) * Kp)) ||
        (rc = clip_alloc(H, &W->bias, static_cast<size_t>(N))))
      return rc;
  }
  W->N = N; W->K = K; W->Kp = Kp;
  // one scale for the stacked tensor: max over the parts
  float mx_scale = 0.f;
  for (int i = 0; i < parts; ++i) {
    float s;
    if ((rc = f16_prescale(w[i], K, rows_each, K, H->scratch, &s, stream))) return rc;
    if (i == 0 || s < mx_scale) mx_scale = s;   // the smallest scale belongs to the largest |w|
  }
  W->scale = mx_scale;
  for (int i = 0; i < parts; ++i) {
    const size_t off = static_cast<size_t>(i) * rows_each * Kp;
    if ((rc = launch_split_planes16(w[i], K, rows_each, K, W->scale, W->hi + off, W->lo + off, Kp, FMT_F16, stream))) return rc;
    if (b && b[i]) {
      EMCID_CUDA_CHECK(cudaMemcpyAsync(W->bias + static_cast<size_t>(i) * rows_each, b[i], rows_each * sizeof(float),
                                       cudaMemcpyDeviceToDevice, stream));
    } else {
      EMCID_CUDA_CHECK(cudaMemsetAsync(W->bias + static_cast<size_t>(i) * rows_each, 0, rows_each * sizeof(float), stream));
    }
  }
  if ((rc = make_tmap_2d(&W->m_hi, W->hi, N, K, Kp, 128, 2)) ||
      (rc = make_tmap_2d(&W->m_lo, W->lo, N, K, Kp, 128, 2)))
    return rc;
  return EMCID_OK;
}

// tensors: ln1.w ln1.b q.w q.b k.w k.b v.w v.b o.w o.b ln2.w ln2.b fc1.w fc1.b fc2.w fc2.b (fp32, device, contiguous)
// changed: bit i set = tensor i differs from what was uploaded last (all ones for a first upload).  Only the operand
// planes that depend on a changed tensor are rebuilt: the edit loop changes fc2.weight of one layer between two key
// extractions (emcid/emcid_main.py:1061), and rebuilding all four weights of the layer — six prescale reductions with a
// stream synchronisation each — was most of the host time of a sequential edit (profiles/round2/r04b_profile_edit.txt).
inline int clip_set_layer(ClipHandle* H, int layer, const float* const* t, cudaStream_t stream, unsigned changed = 0xFFFFu) {
  EMCID_CHECK(H && t && layer >= 0 && layer < H->L, EMCID_ERR_INVALID, "clip_set_layer: bad argument");
  for (int i = 0; i < 16; ++i)
    EMCID_CHECK(t[i] != nullptr || i == 3 || i == 5 || i == 7 || i == 9 || i == 13 || i == 15, EMCID_ERR_INVALID,
                "clip_set_layer: tensor %d is null", i);
  EMCID_CUDA_CHECK(cudaSetDevice(H->device));
  ClipLayer& Ly = (*H->layers)[layer];
  if (!Ly.set) changed = 0xFFFFu;
  int rc;
  if (!Ly.ln1_w) {
    if ((rc = clip_alloc(H, &Ly.ln1_w, static_cast<size_t>(H->h))) || (rc = clip_alloc(H, &Ly.ln1_b, static_cast<size_t>(H->h))) ||
        (rc = clip_alloc(H, &Ly.ln2_w, static_cast<size_t>(H->h))) || (rc = clip_alloc(H, &Ly.ln2_b, static_cast<size_t>(H->h))))
      return rc;
  }
  const size_t hb = static_cast<size_t>(H->h) * sizeof(float);
  if (changed & 0x0003u) {
    EMCID_CUDA_CHECK(cudaMemcpyAsync(Ly.ln1_w, t[0], hb, cudaMemcpyDeviceToDevice, stream));
    EMCID_CUDA_CHECK(cudaMemcpyAsync(Ly.ln1_b, t[1], hb, cudaMemcpyDeviceToDevice, stream));
  }
  if (changed & 0x0C00u) {
    EMCID_CUDA_CHECK(cudaMemcpyAsync(Ly.ln2_w, t[10], hb, cudaMemcpyDeviceToDevice, stream));
    EMCID_CUDA_CHECK(cudaMemcpyAsync(Ly.ln2_b, t[11], hb, cudaMemcpyDeviceToDevice, stream));
  }
  const float* wq[3] = {t[2], t[4], t[6]};
  const float* bq[3] = {t[3], t[5], t[7]};
  if ((changed & 0x00FCu) && (rc = clip_prepare_weight(H, &Ly.qkv, 3, wq, bq, H->h, H->h, stream))) return rc;
  if ((changed & 0x0300u) && (rc = clip_prepare_weight(H, &Ly.o, 1, &t[8], &t[9], H->h, H->h, stream))) return rc;
  if ((changed & 0x3000u) && (rc = clip_prepare_weight(H, &Ly.fc1, 1, &t[12], &t[13], H->d, H->h, stream))) return rc;
  if ((changed & 0xC000u) && (rc = clip_prepare_weight(H, &Ly.fc2, 1, &t[14], &t[15], H->h, H->d, stream))) return rc;
  Ly.set = true;
  return EMCID_OK;
}

struct ClipActMaps {
  GemmOperands x, a, f;   // a_hi / a_lo of each hold the activation planes (rows = tokens of this call)
};

// out[T x N] = act(X[T x K] W^T / scale + bias) (+ residual), optionally as fp32 / planes / transposed planes.
inline int clip_linear(ClipHandle* H, const CUtensorMap& x_hi, const CUtensorMap& x_lo, const ClipWeight& W, int T,
                       int act, const float* Cin, float* C, long long ldc, uint16_t* P_hi, uint16_t* P_lo, long long ldp,
                       cudaStream_t stream, const GemmOutMaps* om = nullptr, int tag = CLIP_TAG_QKV) {
  ClipProfScope prof(H, stream, tag, 2.0 * T * static_cast<double>(W.N) * W.K);
  GemmOperands ops;
  ops.a_hi = x_hi; ops.a_lo = x_lo; ops.b_hi = W.m_hi; ops.b_lo = W.m_lo;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = T; p.N = W.N; p.K = W.K;
  // k-blocks (64 contracted elements each) accumulated in TMEM between round-to-nearest folds: the RZ bias of the TMEM
  // accumulation grows with it (gemm3x.cuh), and so does the slack the epilogue's store phase has behind the MMA warp
  // (two accumulator chunks).  Measured at CLIP-L, every layer at the same chunk (profiles/round1/r02d_*, r02g_*): chunk 2 / 3 / 4 ->
  // 1.650 / 1.682 / 1.724 M tokens/s, hidden-state error vs fp64 1.26e-6 / 1.78e-6 / 2.35e-6 (HF's own fp32 run: 1.46e-6),
  // mom2 error 1.84e-6 / 2.4e-6 / 3.2e-6 (tolerance 1e-5).  The slack only matters where a tile is short: with K <= 1024
  // (q/k/v, out projection, fc1: 12 k-blocks) two chunks of 2 are a third of the tile and the store phase of the previous
  // tile (128 activations + fp16 splits per thread) does not fit behind them.  So: chunk 3 for the products that contract
  // over the hidden width (K = 768; bigG: 1280), chunk 2 for fc2 (K = 3072 / 5120, 91 % tensor-pipe active anyway, and the
  // longest accumulation).  Same-box A/Bs: CLIP-L 92.9 -> 90.2 ms per step, mom2 probe error over 1.2 M tokens 2.40e-6 ->
  // 2.88e-6 (profiles/round2/r04f_ab.txt); bigG 545.5 -> 533.1 ms per step (profiles/round2/r05l_bench_sdxl_text2*.json).
  // EMCID_LINEAR_CHUNK=n sets every layer, EMCID_LINEAR_CHUNK_SHORTK=n the K <= 2048 products.
  static const int chunk_env = [] { const char* e = getenv("EMCID_LINEAR_CHUNK"); return e ? atoi(e) : 0; }();
  static const int chunk_short = [] { const char* e = getenv("EMCID_LINEAR_CHUNK_SHORTK"); return e ? atoi(e) : 0; }();
  p.chunk_kblocks = chunk_env > 0 ? chunk_env : (W.K <= 2048 ? (chunk_short > 0 ? chunk_short : 3) : 2);
  // token tiles outermost: the activation planes (121-484 MB per block) stream from HBM once while the weight
  // planes (<= 19 MB) stay L2-resident.  EMCID_TILE_ORDER=m restores the M-fastest walk (measured: every N tile
  // re-read the activations from DRAM, 1.1-1.5 GB per launch).
  static const bool m_fast = [] { const char* e = getenv("EMCID_TILE_ORDER"); return e && e[0] == 'm'; }();
  p.n_fastest = m_fast ? 0 : 1;
  p.alpha = 1.0f / W.scale; p.beta = Cin ? 1.0f : 0.0f;
  p.bias_col = W.bias; p.act = act; p.lo_fmt = FMT_F16;
  p.Cin = Cin; p.ldcin = ldc;
  p.C = C; p.ldc = ldc;
  p.P_hi = reinterpret_cast<float*>(P_hi); p.P_lo = reinterpret_cast<float*>(P_lo); p.ldp = ldp;
  const int tiles = gemm_num_tiles(T, W.N, 256, 0);
  const int grid = tiles < H->info.sm_count ? tiles : H->info.sm_count;
  H->launches += 1;
  // the epilogue variant is a compile-time option (see EF_* in gemm3x.cuh): only the combinations the forward uses
  const int ef = act | (C ? EF_C : 0) | (Cin ? EF_CIN : 0) | (P_hi ? EF_P : 0);
  const bool cta2 = om && gemm_cta2_enabled();
  const int pair_tiles = ((T + 255) / 256) * ((W.N + 255) / 256);
  const int grid2 = 2 * pair_tiles < (H->info.sm_count & ~1) ? 2 * pair_tiles : (H->info.sm_count & ~1);
#define EMCID_LIN_CASE(F)                                                                                     \
  if (ef == (F))                                                                                              \
    return !om            ? launch_gemm3x<256, 2, EPI_LINEAR, KIND_F16, (F)>(ops, p, grid, stream)            \
           : cta2          ? launch_gemm3x<256, 3, EPI_LINEAR_TMA, KIND_F16, (F), 1>(ops, p, grid2, stream, 1, om) \
                           : launch_gemm3x<256, 2, EPI_LINEAR_TMA, KIND_F16, (F)>(ops, p, grid, stream, 1, om);
  EMCID_LIN_CASE(ACT_NONE | EF_C)                       // q/k/v projection (fp32, CUDA-core attention)
  if (cta2 && ef == (ACT_NONE | EF_P))
    return launch_gemm3x<256, 3, EPI_LINEAR_TMA, KIND_F16, (ACT_NONE | EF_P), 1>(ops, p, grid2, stream, 1, om);
  if (om && ef == (ACT_NONE | EF_P))                    // q/k/v projection as planes (tensor-core attention)
    return launch_gemm3x<256, 2, EPI_LINEAR_TMA, KIND_F16, (ACT_NONE | EF_P)>(ops, p, grid, stream, 1, om);
  EMCID_LIN_CASE(ACT_NONE | EF_C | EF_CIN)              // out projection / fc2 with residual
  EMCID_LIN_CASE(ACT_QUICK_GELU | EF_P)                 // fc1
  EMCID_LIN_CASE(ACT_GELU_ERF | EF_P)
#undef EMCID_LIN_CASE
  return set_error(EMCID_ERR_UNSUPPORTED, "clip_linear: epilogue variant %d is not instantiated", ef);
}

inline int clip_layernorm(ClipHandle* H, const float* x, int T, const float* w, const float* b, uint16_t* o_hi,
                          uint16_t* o_lo, cudaStream_t stream) {
  ClipProfScope prof(H, stream, CLIP_TAG_LN, 0.0);
  int blocks = (T + 7) / 8;
  if (blocks > H->info.sm_count * 16) blocks = H->info.sm_count * 16;
  H->launches += 1;
  if (H->h % 4 == 0 && H->hp % 4 == 0 && (reinterpret_cast<uintptr_t>(w) & 15) == 0 && (reinterpret_cast<uintptr_t>(b) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
    const int nv4 = (H->h / 4 + 31) / 32;
#define EMCID_LN4_CASE(NV4)                                                                                       \
  if (nv4 <= NV4) {                                                                                               \
    clip_layernorm4_kernel<NV4><<<blocks, 256, 0, stream>>>(x, T, H->h, w, b, H->eps, o_hi, o_lo, H->hp);        \
    EMCID_CUDA_CHECK(cudaGetLastError());                                                                         \
    return EMCID_OK;                                                                                              \
  }
    EMCID_LN4_CASE(1) EMCID_LN4_CASE(2) EMCID_LN4_CASE(4) EMCID_LN4_CASE(6) EMCID_LN4_CASE(8) EMCID_LN4_CASE(10)
    EMCID_LN4_CASE(16)
#undef EMCID_LN4_CASE
  }
  return set_error(EMCID_ERR_UNSUPPORTED, "clip_layernorm: hidden size %d too large", H->h);
}

// Attention units of a packed block: greedy runs of consecutive captions with at most `lp` tokens in all (block 0,
// thread 0: one pass over cu_seqlens, ~40 us for 2500 captions, once per block) and, for every token, the first token of
// its caption (all blocks).  77-token captions stay alone (two do not fit lp = 80); 15-token captions go five to a tile.
__global__ void clip_group_captions_kernel(const int* __restrict__ cu, int S, int lp, int* __restrict__ grp,
                                           int* __restrict__ n_grp, int* __restrict__ tok_start) {
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int g = 0, g0 = cu[0];
    grp[0] = g0;
    for (int s = 1; s <= S; ++s) {          // caption s - 1 = [cu[s - 1], cu[s])
      const int a = cu[s - 1], b = cu[s];
      if (b - g0 > lp && a > g0) { grp[++g] = a; g0 = a; }
    }
    grp[++g] = cu[S];
    *n_grp = g;
  }
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < S; s += gridDim.x * blockDim.x) {
    const int a = cu[s], b = cu[s + 1];
    for (int t = a; t < b; ++t) tok_start[t] = a;
  }
}

// Key extraction: rows `rows[r]` of the act(fc1) planes -> fp32 keys (hi + lo: the 22-bit split is exact to 2^-23) and a
// compact copy of the planes (operand of the fc2 product on those rows only).
__global__ void clip_gather_keys_kernel(const uint16_t* __restrict__ f_hi, const uint16_t* __restrict__ f_lo, long long ldf,
                                        const int* __restrict__ rows, int R, int T, int d, int dp, float* __restrict__ k_out,
                                        uint16_t* __restrict__ g_hi, uint16_t* __restrict__ g_lo) {
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    int src = rows[r];
    src = src < 0 ? 0 : (src >= T ? T - 1 : src);
    const uint16_t* sh = f_hi + static_cast<long long>(src) * ldf;
    const uint16_t* sl = f_lo + static_cast<long long>(src) * ldf;
    for (int c = threadIdx.x; c < dp; c += blockDim.x) {
      const uint16_t hh = sh[c], ll = sl[c];
      g_hi[static_cast<long long>(r) * dp + c] = hh;
      g_lo[static_cast<long long>(r) * dp + c] = ll;
      if (c < d) k_out[static_cast<long long>(r) * d + c] = __half2float(__ushort_as_half(hh)) + __half2float(__ushort_as_half(ll));
    }
  }
}

// out[tag * 3 + {0, 1, 2}] = {launches, total ms, total algorithmic flops} per ClipTag (CLIP_TAG_COUNT * 3 doubles);
// waits for the recorded events, then clears them.
inline int clip_get_profile(ClipHandle* H, double* out) {
  EMCID_CHECK(H && out, EMCID_ERR_INVALID, "clip_get_profile: null argument");
  EMCID_CUDA_CHECK(cudaSetDevice(H->device));
  for (int i = 0; i < CLIP_TAG_COUNT * 3; ++i) out[i] = 0.0;
  for (size_t i = 0; i + 1 < H->ev->size(); i += 2) {
    EMCID_CUDA_CHECK(cudaEventSynchronize((*H->ev)[i + 1]));
    float ms = 0.f;
    EMCID_CUDA_CHECK(cudaEventElapsedTime(&ms, (*H->ev)[i], (*H->ev)[i + 1]));
    const int tag = (*H->ev_tag)[i / 2];
    out[tag * 3 + 0] += 1.0;
    out[tag * 3 + 1] += ms;
    out[tag * 3 + 2] += (*H->ev_flops)[i / 2];
    cudaEventDestroy((*H->ev)[i]); cudaEventDestroy((*H->ev)[i + 1]);
  }
  H->ev->clear(); H->ev_tag->clear(); H->ev_flops->clear();
  return EMCID_OK;
}

// Runs layers [0, n_layers) (full layers) — or, when stats are requested, up to fc1 of the deepest edited
// layer — over `T` packed tokens of `S` captions.  stat_layers / accs: edited layers (ascending) and their
// accumulators.  hidden_out (optional, [T x h] fp32): the residual stream after the last executed FULL layer.
// Key-extraction mode (keys_layer >= 0; replaces the traced HF forward of emcid/compute_z.py:2300-2316): full layers
// [0, keys_layer), then layer keys_layer up to act(fc1); k_out [n_keys x d] = fc2 INPUT rows `key_rows` (packed token
// indices, device), z_out [n_keys x h] = fc2 OUTPUT of those rows (fc2 runs on the gathered rows only).
inline int clip_forward(ClipHandle* H, const int* ids, const int* pos, const int* cu_seqlens, int S, int T,
                        int n_layers, int n_stat, const int* stat_layers, Mom2Handle* const* accs, float* hidden_out,
                        cudaStream_t stream, int keys_layer = -1, const int* key_rows = nullptr, int n_keys = 0,
                        float* k_out = nullptr, float* z_out = nullptr, int resume_layer = -1) {
  EMCID_CHECK(H && ids && pos && cu_seqlens, EMCID_ERR_INVALID, "clip_forward: null argument");
  EMCID_CHECK(T >= 0 && T <= H->cap_tokens && S >= 0 && S <= H->cap_seqs, EMCID_ERR_INVALID,
              "clip_forward: %d tokens / %d captions exceed the handle capacity (%lld / %d)", T, S, H->cap_tokens, H->cap_seqs);
  EMCID_CHECK(n_layers >= 0 && n_layers <= H->L, EMCID_ERR_INVALID, "clip_forward: bad layer count");
  int last_stat = -1;
  for (int i = 0; i < n_stat; ++i) {
    EMCID_CHECK(stat_layers[i] >= 0 && stat_layers[i] < H->L && accs[i] && (i == 0 || stat_layers[i] > stat_layers[i - 1]),
                EMCID_ERR_INVALID, "clip_forward: edited layers must be ascending, in range and have accumulators");
    EMCID_CHECK(accs[i]->d == H->d && accs[i]->device == H->device, EMCID_ERR_INVALID,
                "clip_forward: accumulator %d does not match the encoder (d=%d vs %d)", i, accs[i]->d, H->d);
    last_stat = stat_layers[i];
  }
  if (keys_layer >= 0) {
    EMCID_CHECK(n_stat == 0 && keys_layer < H->L && key_rows && k_out && z_out && n_keys > 0 && n_keys <= T, EMCID_ERR_INVALID,
                "clip_forward: bad key-extraction arguments (layer %d, %d keys, %d tokens)", keys_layer, n_keys, T);
    EMCID_CHECK(static_cast<long long>(n_keys) * H->dp <= static_cast<long long>(H->d) * H->tp, EMCID_ERR_INVALID,
                "clip_forward: %d keys do not fit the gather buffer", n_keys);
    EMCID_CHECK((reinterpret_cast<uintptr_t>(z_out) & 15) == 0 && (H->h % 4) == 0, EMCID_ERR_INVALID,
                "clip_forward: z_out must be 16-byte aligned");
  }
  // resume_layer = r >= 0: the previous call on this handle was a keys call at layer r over the SAME packed tokens and only
  // fc2 of layer r has changed since (the edit loop, emcid_main.py:1061): start from that call's state — finish layer r
  // with its new fc2, then run layers (r, keys_layer] — instead of from the embeddings.  resume_layer == keys_layer: the
  // same layer again with other key rows — nothing is recomputed but the gather and fc2 of the gathered rows (the edit
  // launches the forward of the prompts before it has worked out WHICH rows it wants: the host-side search for the subject
  // tokens runs beside the device's forward)
  if (resume_layer >= 0) {
    EMCID_CHECK(keys_layer >= resume_layer && H->keys_state_layer == resume_layer && H->keys_state_tokens == T, EMCID_ERR_INVALID,
                "clip_forward: cannot resume from layer %d (state: layer %d, %d tokens; this call: layer %d, %d tokens)",
                resume_layer, H->keys_state_layer, H->keys_state_tokens, keys_layer, T);
  }
  H->keys_state_layer = -1;
  const int run_layers = keys_layer >= 0 ? keys_layer + 1 : (n_stat > 0 ? last_stat + 1 : n_layers);
  for (int l = 0; l < run_layers; ++l)
    EMCID_CHECK((*H->layers)[l].set, EMCID_ERR_INVALID, "clip_forward: weights of layer %d were never set", l);
  if (T == 0) return EMCID_OK;
  EMCID_CUDA_CHECK(cudaSetDevice(H->device));
  const int sms = H->info.sm_count;
  int rc;
  // per-call TMA maps over exactly T rows: out-of-range rows of the last tile read as zeros
  CUtensorMap mx_hi, mx_lo, ma_hi, ma_lo, mf_hi, mf_lo;
  CUtensorMap sf_hi, sf_lo;   // store maps of the f planes (64-byte rows: the staged epilogue's 8 KB sub-tiles, gemm3x.cuh)
  if ((rc = make_tmap_2d(&mx_hi, H->x_hi, T, H->h, H->hp, 128, 2)) || (rc = make_tmap_2d(&mx_lo, H->x_lo, T, H->h, H->hp, 128, 2)) ||
      (rc = make_tmap_2d(&ma_hi, H->a_hi, T, H->h, H->hp, 128, 2)) || (rc = make_tmap_2d(&ma_lo, H->a_lo, T, H->h, H->hp, 128, 2)) ||
      (rc = make_tmap_2d(&mf_hi, H->f_hi, T, H->d, H->dp, 128, 2)) || (rc = make_tmap_2d(&mf_lo, H->f_lo, T, H->d, H->dp, 128, 2)) ||
      (rc = make_tmap_2d(&sf_hi, H->f_hi, T, H->d, H->dp, 128, 2, 0, 64)) || (rc = make_tmap_2d(&sf_lo, H->f_lo, T, H->d, H->dp, 128, 2, 0, 64)))
    return rc;
  // store maps of the staged (TMA) epilogue; EMCID_LINEAR_TMA=0 keeps the direct-store epilogue
  static const bool use_tma_epi = [] { const char* e = getenv("EMCID_LINEAR_TMA"); return !(e && e[0] == '0'); }();
  GemmOutMaps om_qkv = {}, om_res = {}, om_f = {};
  if (use_tma_epi) {
    if ((rc = make_tmap_2d(&om_qkv.c, H->qkv, T, 3ll * H->h, 3ll * H->h, 128, 4, 0, 64)) ||
        (rc = make_tmap_2d(&om_res.c, H->hres, T, H->h, H->h, 128, 4, 0, 64)))
      return rc;
    om_f.c = sf_hi; om_f.c2 = sf_lo;
  }
  const bool attn_tc = H->attn_tc && use_tma_epi;
  AttnMaps am = {};
  if (attn_tc) {
    if ((rc = make_tmap_2d(&am.qk_hi, H->qp_hi, T, 3ll * H->h, 3ll * H->h, 128, 2)) ||
        (rc = make_tmap_2d(&am.qk_lo, H->qp_lo, T, 3ll * H->h, 3ll * H->h, 128, 2)) ||
        (rc = make_tmap_2d(&am.kv_hi, H->qp_hi, T, 3ll * H->h, 3ll * H->h, (H->max_pos + 15) & ~15, 2)) ||
        (rc = make_tmap_2d(&am.kv_lo, H->qp_lo, T, 3ll * H->h, 3ll * H->h, (H->max_pos + 15) & ~15, 2)))
      return rc;
    // the projection STORES the q|k|v planes through 64-byte-row maps; the attention kernel loads them through am.*
    if ((rc = make_tmap_2d(&om_qkv.c, H->qp_hi, T, 3ll * H->h, 3ll * H->h, 128, 2, 0, 64)) ||
        (rc = make_tmap_2d(&om_qkv.c2, H->qp_lo, T, 3ll * H->h, 3ll * H->h, 128, 2, 0, 64)))
      return rc;
  }
  const GemmOutMaps* pm_qkv = use_tma_epi ? &om_qkv : nullptr;
  const GemmOutMaps* pm_res = use_tma_epi ? &om_res : nullptr;
  const GemmOutMaps* pm_f = use_tma_epi ? &om_f : nullptr;

  if (resume_layer < 0) {
    int blocks = (T + 7) / 8;
    if (blocks > sms * 16) blocks = sms * 16;
    clip_embed_kernel<<<blocks, 256, 0, stream>>>(ids, pos, T, H->h, H->vocab, H->max_pos, H->tok_emb, H->pos_emb, H->hres);
    EMCID_CUDA_CHECK(cudaGetLastError());
    H->launches += 1;
  }
  const float scale = 1.0f / sqrtf(static_cast<float>(H->dh));
  // short captions share attention tiles (attn.cuh) when they average under half a tile (at a mean of 40 tokens, len ~
  // U{4..77}, grouped and ungrouped measure the same: profiles/round2/r05t_ab_grouped_attention.txt)
  const bool grouped = attn_tc && static_cast<long long>(T) * 2 <= static_cast<long long>(S) * ((H->max_pos + 15) & ~15);
  if (grouped) {
    clip_group_captions_kernel<<<S < 64 * 256 ? (S + 255) / 256 : 64, 256, 0, stream>>>(cu_seqlens, S, (H->max_pos + 15) & ~15,
                                                                                         H->grp, H->n_grp, H->tok_start);
    EMCID_CUDA_CHECK(cudaGetLastError());
    H->launches += 1;
  }
  int si = 0;
  // keys of layer keys_layer: k_out = the key rows of f = act(fc1), z_out = fc2 of those rows; leaves the resumable state
  auto gather_keys = [&](const ClipLayer& Ly) -> int {
    uint16_t* g_hi = H->ft_hi;   // compact [n_keys x dp] copy of the key rows
    uint16_t* g_lo = H->ft_lo;
    clip_gather_keys_kernel<<<n_keys < sms * 8 ? n_keys : sms * 8, 256, 0, stream>>>(
        H->f_hi, H->f_lo, H->dp, key_rows, n_keys, T, H->d, H->dp, k_out, g_hi, g_lo);
    EMCID_CUDA_CHECK(cudaGetLastError());
    H->launches += 1;
    CUtensorMap mg_hi, mg_lo;
    GemmOutMaps om_z = {};
    int r;
    if ((r = make_tmap_2d(&mg_hi, g_hi, n_keys, H->d, H->dp, 128, 2)) ||
        (r = make_tmap_2d(&mg_lo, g_lo, n_keys, H->d, H->dp, 128, 2)) ||
        (use_tma_epi && (r = make_tmap_2d(&om_z.c, z_out, n_keys, H->h, H->h, 128, 4, 0, 64))))
      return r;
    if ((r = clip_linear(H, mg_hi, mg_lo, Ly.fc2, n_keys, ACT_NONE, nullptr, z_out, H->h, nullptr, nullptr, 0, stream,
                         use_tma_epi ? &om_z : nullptr, CLIP_TAG_FC2)))
      return r;
    H->keys_state_layer = keys_layer;
    H->keys_state_tokens = T;
    return EMCID_OK;
  };
  for (int l = resume_layer < 0 ? 0 : resume_layer; l < run_layers; ++l) {
    const ClipLayer& Ly = (*H->layers)[l];
    if (l == resume_layer && l == keys_layer) {
      if ((rc = gather_keys(Ly))) return rc;
      break;
    }
    if (l == resume_layer) {
      // hres = h + attn(h) of this layer and f = act(fc1(LN2 hres)) are where the previous call left them
      if ((rc = clip_linear(H, mf_hi, mf_lo, Ly.fc2, T, ACT_NONE, H->hres, H->hres, H->h, nullptr, nullptr, 0, stream, pm_res,
                            CLIP_TAG_FC2)))
        return rc;
      continue;
    }
    if ((rc = clip_layernorm(H, H->hres, T, Ly.ln1_w, Ly.ln1_b, H->x_hi, H->x_lo, stream))) return rc;
    if (attn_tc) {
      if ((rc = clip_linear(H, mx_hi, mx_lo, Ly.qkv, T, ACT_NONE, nullptr, nullptr, 0, H->qp_hi, H->qp_lo, 3ll * H->h, stream,
                            pm_qkv, CLIP_TAG_QKV)))
        return rc;
      ClipProfScope prof_attn(H, stream, CLIP_TAG_ATTN, 0.0);
      const int units = S * H->heads;            // grouped: an upper bound, the kernel reads the number of groups
      const int lp = (H->max_pos + 15) & ~15;
      const int per_sm = attn_ctas_per_sm(lp);   // three resident CTAs per SM for CLIP's 77 tokens
      const int grid = units < per_sm * sms ? units : per_sm * sms;
#define EMCID_ATTN_TC(NC, G)                                                                                        \
  clip_attention_tc_kernel<NC, G><<<grid, ATTN_THREADS, attn_smem_bytes(lp), stream>>>(                             \
      am, G ? H->grp : cu_seqlens, units, H->n_grp, H->tok_start, H->heads, H->h, lp, scale, H->a_hi, H->a_lo, H->hp)
      if (lp <= 80) { if (grouped) EMCID_ATTN_TC(5, true); else EMCID_ATTN_TC(5, false); }
      else { if (grouped) EMCID_ATTN_TC(8, true); else EMCID_ATTN_TC(8, false); }
#undef EMCID_ATTN_TC
    } else {
    if ((rc = clip_linear(H, mx_hi, mx_lo, Ly.qkv, T, ACT_NONE, nullptr, H->qkv, 3ll * H->h, nullptr, nullptr, 0, stream, pm_qkv)))
      return rc;
    {
      const dim3 ag(H->heads, S);
#define EMCID_ATTN_CASE(D)                                                                                             \
  if (H->dh == D)                                                                                                      \
    clip_attention2_kernel<D><<<ag, 128, H->attn_smem, stream>>>(H->qkv, cu_seqlens, H->h, H->max_pos, scale, H->a_hi, \
                                                                 H->a_lo, H->hp);
      EMCID_ATTN_CASE(16) else EMCID_ATTN_CASE(32) else EMCID_ATTN_CASE(64) else EMCID_ATTN_CASE(128)
#undef EMCID_ATTN_CASE
    }
    }
    EMCID_CUDA_CHECK(cudaGetLastError());
    H->launches += 1;
    if ((rc = clip_linear(H, ma_hi, ma_lo, Ly.o, T, ACT_NONE, H->hres, H->hres, H->h, nullptr, nullptr, 0, stream, pm_res,
                          CLIP_TAG_OUT)))
      return rc;
    if ((rc = clip_layernorm(H, H->hres, T, Ly.ln2_w, Ly.ln2_b, H->x_hi, H->x_lo, stream))) return rc;
    const bool is_stat = si < n_stat && stat_layers[si] == l;
    const bool is_last = n_stat > 0 && l == last_stat;
    // The SYRK reads act(fc1) as MN-major operand tiles straight from the [tokens x features] planes fc2 reads as well, so an
    // edited layer's fc1 is an ordinary fc1.  (Until r04 a switch kept the first design alive — K-major tiles of a transposed
    // copy f^T that fc1 of an edited layer wrote as well: +100 us per launch and 465 MB of extra HBM traffic per block.)
    if ((rc = clip_linear(H, mx_hi, mx_lo, Ly.fc1, T, H->act, nullptr, nullptr, 0, H->f_hi, H->f_lo, H->dp, stream, pm_f,
                          is_stat ? CLIP_TAG_FC1_STAT : CLIP_TAG_FC1)))
      return rc;
    if (is_stat) {
      Mom2Handle* A = accs[si++];
      // One launch over the whole block (hybrid schedule, gemm3x.cuh: whole tiles march through the tokens together, so
      // each k-block of the planes is pulled from HBM once and shared through L2, and a tile is red.add'ed once per
      // block instead of once per 4096-token slab; the 4 leftover pair tiles are stream-K'd).  EMCID_SYRK_HYBRID=0
      // stream-K's every tile.
      static const bool hybrid = [] { const char* e = getenv("EMCID_SYRK_HYBRID"); return !(e && e[0] == '0'); }();
      GemmOperands fm;   // act(fc1) planes [T x d]: boxes of 64 tokens x 64 features
      if ((rc = make_tmap_2d(&fm.a_hi, H->f_hi, T, H->d, H->dp, 64, 2)) || (rc = make_tmap_2d(&fm.a_lo, H->f_lo, T, H->d, H->dp, 64, 2)))
        return rc;
      fm.b_hi = fm.a_hi; fm.b_lo = fm.a_lo;
      // the block's token count rides on the launch
      if ((rc = mom2_syrk_slab(A, fm, KIND_F16_MN, 0, T, nullptr, stream, hybrid ? 2 : 1, T))) return rc;
      A->slabs_since_fold += 4;
      if (A->slabs_since_fold >= MOM2_FOLD_EVERY && (rc = mom2_fold(A, stream))) return rc;
    }
    if (is_last) break;
    if (l == keys_layer) {
      if ((rc = gather_keys(Ly))) return rc;
      break;
    }
    if ((rc = clip_linear(H, mf_hi, mf_lo, Ly.fc2, T, ACT_NONE, H->hres, H->hres, H->h, nullptr, nullptr, 0, stream, pm_res,
                          CLIP_TAG_FC2)))
      return rc;
  }
  if (hidden_out) {
    EMCID_CUDA_CHECK(cudaMemcpyAsync(hidden_out, H->hres, static_cast<size_t>(T) * H->h * sizeof(float),
                                     cudaMemcpyDeviceToDevice, stream));
  }
  return EMCID_OK;
}

inline int clip_set_final_norm(ClipHandle* H, const float* w, const float* b, cudaStream_t stream) {
  EMCID_CHECK(H && w && b, EMCID_ERR_INVALID, "clip_set_final_norm: null argument");
  EMCID_CUDA_CHECK(cudaSetDevice(H->device));
  int rc;
  if (!H->fln_w) {
    if ((rc = clip_alloc(H, &H->fln_w, static_cast<size_t>(H->h))) || (rc = clip_alloc(H, &H->fln_b, static_cast<size_t>(H->h))))
      return rc;
  }
  const size_t hb = static_cast<size_t>(H->h) * sizeof(float);
  EMCID_CUDA_CHECK(cudaMemcpyAsync(H->fln_w, w, hb, cudaMemcpyDeviceToDevice, stream));
  EMCID_CUDA_CHECK(cudaMemcpyAsync(H->fln_b, b, hb, cudaMemcpyDeviceToDevice, stream));
  H->fln_set = true;
  return EMCID_OK;
}

// rows[r] of the hi/lo planes -> fp32 (hi + lo is the value to 2^-23); rows == nullptr: row r itself.
__global__ void clip_planes_to_f32_kernel(const uint16_t* __restrict__ p_hi, const uint16_t* __restrict__ p_lo, long long ldp,
                                          const int* __restrict__ rows, int R, int T, int h, float* __restrict__ out) {
  for (int r = blockIdx.x; r < R; r += gridDim.x) {
    int src = rows ? rows[r] : r;
    src = src < 0 ? 0 : (src >= T ? T - 1 : src);
    const uint16_t* sh = p_hi + static_cast<long long>(src) * ldp;
    const uint16_t* sl = p_lo + static_cast<long long>(src) * ldp;
    for (int c = threadIdx.x; c < h; c += blockDim.x)
      out[static_cast<long long>(r) * h + c] = __half2float(__ushort_as_half(sh[c])) + __half2float(__ushort_as_half(sl[c]));
  }
}

// The text encoder's OUTPUT, last_hidden_state = final_layer_norm(residual stream after all layers): the common input of
// every UNet cross-attention to_k / to_v projection.  Replaces, for those modules,
//     text_repr = pipe.text_encoder(**batch).last_hidden_state ; pipe.unet(latents, t, encoder_hidden_states=text_repr)
//     feats = flatten_masked_batch(tr.input, mask) ; stat.add(feats)             emcid/layer_stats.py:408-426
// (one pass per K/V layer in the reference although all 32 see the same input, :429-467) and the key extraction
//     source_inp_repr = pipe.text_encoder(**inp)[0] ; td[module].input[i, idx]    emcid/compute_ks.py:91-124.
// acc (optional, d == hidden): mom2 += y^T y, count += T over the packed tokens.  rows/n_rows/out (optional): fp32
// last_hidden_state of the packed token rows `rows` (device int32; nullptr = all T tokens in order) -> out [n_rows x hidden].
inline int clip_forward_final(ClipHandle* H, const int* ids, const int* pos, const int* cu_seqlens, int S, int T,
                              Mom2Handle* acc, const int* rows, int n_rows, float* out, cudaStream_t stream) {
  EMCID_CHECK(H && H->fln_set, EMCID_ERR_INVALID, "clip_forward_final: call emcid_clip_set_final_norm first");
  EMCID_CHECK(!acc || (acc->d == H->h && acc->device == H->device), EMCID_ERR_INVALID,
              "clip_forward_final: the accumulator must have d == hidden (%d)", H->h);
  EMCID_CHECK(!out || n_rows > 0, EMCID_ERR_INVALID, "clip_forward_final: out needs n_rows > 0");
  int rc = clip_forward(H, ids, pos, cu_seqlens, S, T, H->L, 0, nullptr, nullptr, nullptr, stream);
  if (rc || T == 0) return rc;
  if ((rc = clip_layernorm(H, H->hres, T, H->fln_w, H->fln_b, H->x_hi, H->x_lo, stream))) return rc;
  if (acc) {
    GemmOperands ym;   // last_hidden_state planes [T x hidden]: MN-major operand tiles of 64 tokens x 64 features
    if ((rc = make_tmap_2d(&ym.a_hi, H->x_hi, T, H->h, H->hp, 64, 2)) || (rc = make_tmap_2d(&ym.a_lo, H->x_lo, T, H->h, H->hp, 64, 2)))
      return rc;
    ym.b_hi = ym.a_hi; ym.b_lo = ym.a_lo;
    if ((rc = mom2_syrk_slab(acc, ym, KIND_F16_MN, 0, T, nullptr, stream, 2, T))) return rc;
    acc->slabs_since_fold += 4;
    if (acc->slabs_since_fold >= MOM2_FOLD_EVERY && (rc = mom2_fold(acc, stream))) return rc;
  }
  if (out) {
    const int R = rows ? n_rows : (n_rows < T ? n_rows : T);
    clip_planes_to_f32_kernel<<<R < H->info.sm_count * 8 ? R : H->info.sm_count * 8, 256, 0, stream>>>(
        H->x_hi, H->x_lo, H->hp, rows, R, T, H->h, out);
    EMCID_CUDA_CHECK(cudaGetLastError());
    H->launches += 1;
  }
  return EMCID_OK;
}

}  // namespace emcid

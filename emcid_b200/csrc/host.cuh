// emcid_b200 — host-side plumbing for the C ABI: error reporting, TMA tensor-map encoding and
// launch helpers for the generic 3xTF32 GEMM.  No exceptions cross the C boundary.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>

#include "gemm3x.cuh"

namespace emcid {

// ----------------------------------------------------------------------------------------------
// last-error string (thread local), returned by emcid_last_error()
// ----------------------------------------------------------------------------------------------
inline char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}

inline int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(last_error_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define EMCID_CUDA_CHECK(expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      return ::emcid::set_error(EMCID_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,    \
                                cudaGetErrorString(_e), __FILE__, __LINE__);                \
    }                                                                                       \
  } while (0)

#define EMCID_CHECK(cond, code, ...)                                  \
  do {                                                                \
    if (!(cond)) return ::emcid::set_error((code), __VA_ARGS__);      \
  } while (0)

// ----------------------------------------------------------------------------------------------
// device properties
// ----------------------------------------------------------------------------------------------
struct DeviceInfo {
  int device = -1;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  int l2_bytes = 0;
};

inline int get_device_info(DeviceInfo* out) {
  static thread_local DeviceInfo cache[16];
  int dev = 0;
  EMCID_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 16 && cache[dev].device == dev) {
    *out = cache[dev];
    return EMCID_OK;
  }
  DeviceInfo info;
  info.device = dev;
  EMCID_CUDA_CHECK(cudaDeviceGetAttribute(&info.sm_count, cudaDevAttrMultiProcessorCount, dev));
  EMCID_CUDA_CHECK(cudaDeviceGetAttribute(&info.cc_major, cudaDevAttrComputeCapabilityMajor, dev));
  EMCID_CUDA_CHECK(cudaDeviceGetAttribute(&info.cc_minor, cudaDevAttrComputeCapabilityMinor, dev));
  EMCID_CUDA_CHECK(cudaDeviceGetAttribute(&info.l2_bytes, cudaDevAttrL2CacheSize, dev));
  EMCID_CHECK(info.cc_major == 10, EMCID_ERR_UNSUPPORTED,
              "emcid_b200 kernels are built for sm_100a only; device %d is sm_%d%d", dev,
              info.cc_major, info.cc_minor);
  if (dev >= 0 && dev < 16) cache[dev] = info;
  *out = info;
  return EMCID_OK;
}

// ----------------------------------------------------------------------------------------------
// Device memory of the handles.  cudaFree synchronises the device and was measured at up to 2.8 s for the ~140
// buffers of one statistics pass (text-encoder planes + five accumulators), inside the timed end-to-end call.  Freed
// buffers are therefore parked in a per-device, exact-size free list and handed out again by the next handle of the
// same shape; emcid_release_cached_memory() returns them to the driver.
// ----------------------------------------------------------------------------------------------
struct DevPool {
  std::mutex mu;
  std::map<void*, std::pair<int, size_t>> live;                 // ptr -> (device, bytes)
  std::multimap<std::pair<int, size_t>, void*> parked;          // (device, bytes) -> ptr
  size_t parked_bytes = 0;
  size_t limit_bytes = 24ull << 30;                             // EMCID_POOL_LIMIT_MB; beyond it buffers are freed at once
  DevPool() {
    const char* e = getenv("EMCID_POOL_LIMIT_MB");
    if (e) limit_bytes = static_cast<size_t>(strtoull(e, nullptr, 10)) << 20;
  }
};

inline DevPool& dev_pool() {
  static DevPool* p = new DevPool();   // leaked on purpose: no destruction-order hazards at process exit
  return *p;
}

inline cudaError_t dev_alloc(void** out, size_t bytes) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  DevPool& P = dev_pool();
  std::lock_guard<std::mutex> lock(P.mu);
  auto it = P.parked.find(std::make_pair(dev, bytes));
  if (it != P.parked.end()) {
    *out = it->second;
    P.parked.erase(it);
    P.parked_bytes -= bytes;
  } else {
    e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) {
      // make room: give everything parked on this device back and retry once
      for (auto p = P.parked.begin(); p != P.parked.end();) {
        if (p->first.first == dev) { cudaFree(p->second); P.parked_bytes -= p->first.second; p = P.parked.erase(p); } else { ++p; }
      }
      (void)cudaGetLastError();
      e = cudaMalloc(out, bytes);
      if (e != cudaSuccess) return e;
    }
  }
  P.live[*out] = std::make_pair(dev, bytes);
  return cudaSuccess;
}

inline void dev_free(void* p) {
  if (!p) return;
  DevPool& P = dev_pool();
  std::lock_guard<std::mutex> lock(P.mu);
  auto it = P.live.find(p);
  if (it == P.live.end()) { cudaFree(p); return; }
  if (P.parked_bytes + it->second.second > P.limit_bytes) {
    cudaFree(p);                                                  // the pool is full: do not hoard differently-shaped buffers
  } else {
    P.parked.insert(std::make_pair(it->second, p));
    P.parked_bytes += it->second.second;
  }
  P.live.erase(it);
}

inline int dev_release_cached() {
  DevPool& P = dev_pool();
  std::lock_guard<std::mutex> lock(P.mu);
  int prev = 0;
  cudaGetDevice(&prev);
  for (auto& kv : P.parked) {
    cudaSetDevice(kv.first.first);
    cudaFree(kv.second);
  }
  P.parked.clear();
  P.parked_bytes = 0;
  cudaSetDevice(prev);
  return EMCID_OK;
}

// ----------------------------------------------------------------------------------------------
// TMA tensor maps.  cuTensorMapEncodeTiled is resolved through the runtime so the library has
// no link-time dependency on libcuda (it must load on GPU-less build hosts).
// ----------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<PFN_encodeTiled>(p);
    }
  }
  return fn;
}

// Row-major fp32 matrix [rows x cols] with row pitch `ld` elements; boxes of 32 cols x box_rows,
// written to shared memory with the 128-byte swizzle the UMMA descriptors expect.
// elem_bytes = 4 (tf32 planes) or 2 (fp16 / bf16 planes): the box is one 128-byte swizzle row wide unless
// box_cols > 0 asks for an unswizzled box of that many elements (store maps of transposed planes).
inline int make_tmap_2d(CUtensorMap* out, const void* base, long long rows, long long cols,
                        long long ld, int box_rows, int elem_bytes = 4, int box_cols = 0, int row_bytes = GEMM_ROW_BYTES) {
  PFN_encodeTiled fn = get_encode_fn();
  EMCID_CHECK(fn != nullptr, EMCID_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  EMCID_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, EMCID_ERR_INVALID,
              "TMA base pointer must be 16-byte aligned");
  EMCID_CHECK((ld * elem_bytes) % 16 == 0 && ld >= cols, EMCID_ERR_INVALID,
              "TMA row pitch must be a multiple of 16 bytes and >= cols (ld=%lld cols=%lld)", ld, cols);
  EMCID_CHECK(rows > 0 && cols > 0, EMCID_ERR_INVALID, "empty TMA tensor");
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols > 0 ? box_cols : row_bytes / elem_bytes),
                       static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 2,
                  const_cast<void*>(base), gdim, gstride,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  box_cols > 0 ? CU_TENSOR_MAP_SWIZZLE_NONE
                               : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EMCID_CHECK(r == CUDA_SUCCESS, EMCID_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d",
              static_cast<int>(r));
  return EMCID_OK;
}

// ----------------------------------------------------------------------------------------------
// GEMM launch
// ----------------------------------------------------------------------------------------------
struct GemmOperands {
  CUtensorMap a_hi, a_lo, b_hi, b_lo;
};

template <int BLOCK_N, int STAGES, int EPI, int KIND = KIND_TF32, int EFLAGS = EF_DEFAULT, int CTA2 = 0>
inline int launch_gemm3x(const GemmOperands& ops, const GemmParams& p, int grid, cudaStream_t stream,
                         int batches = 1, const GemmOutMaps* out_maps = nullptr) {
  using Cfg = GemmCfg<BLOCK_N, STAGES, KindTraits<KIND>::kRowBytes>;
  // pair mode stages only half of B per CTA
  constexpr int kStage = 2 * Cfg::kAPlaneBytes + (CTA2 ? Cfg::kBPlaneBytes : 2 * Cfg::kBPlaneBytes);
  constexpr int kSmem = STAGES * kStage + 1024 + 256 + (EPI == EPI_LINEAR_TMA ? GEMM_STAGING_BYTES + 1024 /*bias rows*/ : 0);
  static_assert(kSmem <= 227 * 1024, "shared memory budget exceeded");
  static const GemmOutMaps no_maps = {};
  auto kern = gemm3x_kernel<BLOCK_N, STAGES, EPI, KIND, EFLAGS, CTA2>;
  static thread_local bool configured[16] = {false};
  int dev = 0;
  EMCID_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 16 || !configured[dev]) {
    EMCID_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    if (dev >= 0 && dev < 16) configured[dev] = true;
  }
  if (grid < 1) grid = 1;
  if (CTA2) {
    // clusters of two CTAs along x: `grid` counts CTAs and must be even
    if (grid & 1) grid += 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid, batches);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = kSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    EMCID_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, ops.a_hi, ops.a_lo, ops.b_hi, ops.b_lo,
                                        out_maps ? *out_maps : no_maps, p));
    return EMCID_OK;
  }
  kern<<<dim3(grid, batches), GEMM_THREADS, kSmem, stream>>>(ops.a_hi, ops.a_lo, ops.b_hi, ops.b_lo,
                                                              out_maps ? *out_maps : no_maps, p);
  EMCID_CUDA_CHECK(cudaGetLastError());
  return EMCID_OK;
}

// CTA-pair (cta_group::2) kernels for the linear layers of the forward (measured on B200: +4.5 % tokens/s over
// single-CTA tiles, parity suite identical); EMCID_CTA2=0 falls back to single-CTA tiles.
inline bool gemm_cta2_enabled() {
  static const bool on = [] { const char* e = getenv("EMCID_CTA2"); return !(e && e[0] == '0'); }();
  return on;
}

inline int gemm_num_tiles(int M, int N, int block_n, int lower) {
  const int mt = (M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int nt = (N + block_n - 1) / block_n;
  if (!lower) return mt * nt;
  int t = 0;
  const int R = block_n / GEMM_BLOCK_M;
  for (int j = 0; j < nt; ++j) {
    int c = mt - j * R;
    if (c > 0) t += c;
  }
  return t;
}

// ----------------------------------------------------------------------------------------------
// elementwise helpers
// ----------------------------------------------------------------------------------------------
// planes = split(scale * src); rows x cols, zero-filling the pitch padding [cols, ldp).
__global__ void split_planes_kernel(const float* __restrict__ src, long long ld, int rows, int cols,
                                    float scale, float* __restrict__ hi, float* __restrict__ lo,
                                    long long ldp) {
  const long long total = static_cast<long long>(rows) * ldp;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / ldp);
    const int c = static_cast<int>(i - static_cast<long long>(r) * ldp);
    float h = 0.f, l = 0.f;
    if (c < cols) split_tf32(scale * src[static_cast<long long>(r) * ld + c], h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// max |src| over a pitched matrix (bit pattern of a non-negative float orders like an unsigned int).
__global__ void absmax_kernel(const float* __restrict__ src, long long ld, int rows, int cols, unsigned int* out) {
  float m = 0.f;
  const long long total = static_cast<long long>(rows) * cols;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cols;
    m = fmaxf(m, fabsf(src[r * ld + (i - r * cols)]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f && isfinite(m)) atomicMax(out, __float_as_uint(m));
}

// Exact power-of-two pre-scale for a static fp16-split operand: max |w| * scale lands in [2^10, 2^11), so every
// element down to 2^-14 of the largest keeps a normal fp16 lo part (full 22-bit split).  Synchronises the stream.
inline int f16_prescale(const float* src, long long ld, int rows, int cols, unsigned int* scratch_dev, float* scale,
                        cudaStream_t stream) {
  EMCID_CUDA_CHECK(cudaMemsetAsync(scratch_dev, 0, sizeof(unsigned int), stream));
  absmax_kernel<<<148 * 4, 256, 0, stream>>>(src, ld, rows, cols, scratch_dev);
  EMCID_CUDA_CHECK(cudaGetLastError());
  unsigned int bits = 0;
  EMCID_CUDA_CHECK(cudaMemcpyAsync(&bits, scratch_dev, sizeof(bits), cudaMemcpyDeviceToHost, stream));
  EMCID_CUDA_CHECK(cudaStreamSynchronize(stream));
  float mx;
  memcpy(&mx, &bits, sizeof(mx));
  *scale = 1.0f;
  if (mx > 0.f) {
    int e = 0;
    frexpf(mx, &e);           // mx = f * 2^e, f in [0.5, 1)  ->  floor(log2 mx) = e - 1
    int k = 10 - (e - 1);
    if (k > 100) k = 100;
    if (k < -100) k = -100;
    *scale = ldexpf(1.0f, k);
  }
  return EMCID_OK;
}

// 16-bit planes (KIND_F16): hi = fp16, lo = bf16 / fp16 of scale * src.
__global__ void split_planes16_kernel(const float* __restrict__ src, long long ld, int rows, int cols,
                                      float scale, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                      long long ldp, int lo_fmt) {
  const long long total = static_cast<long long>(rows) * ldp;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / ldp);
    const int c = static_cast<int>(i - static_cast<long long>(r) * ldp);
    uint16_t h = 0, l = 0;
    if (c < cols) split_f16(scale * src[static_cast<long long>(r) * ld + c], lo_fmt, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

inline int launch_split_planes16(const float* src, long long ld, int rows, int cols, float scale,
                                 void* hi, void* lo, long long ldp, int lo_fmt, cudaStream_t stream) {
  const long long total = static_cast<long long>(rows) * ldp;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  split_planes16_kernel<<<blocks, 256, 0, stream>>>(src, ld, rows, cols, scale, static_cast<uint16_t*>(hi),
                                                     static_cast<uint16_t*>(lo), ldp, lo_fmt);
  EMCID_CUDA_CHECK(cudaGetLastError());
  return EMCID_OK;
}

inline int launch_split_planes(const float* src, long long ld, int rows, int cols, float scale,
                               float* hi, float* lo, long long ldp, cudaStream_t stream) {
  const long long total = static_cast<long long>(rows) * ldp;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  split_planes_kernel<<<blocks, 256, 0, stream>>>(src, ld, rows, cols, scale, hi, lo, ldp);
  EMCID_CUDA_CHECK(cudaGetLastError());
  return EMCID_OK;
}

}  // namespace emcid

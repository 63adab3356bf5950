// emcid_b200 — the exchange step of the caption-sharded statistics pass, the accumulator state a resumable pass
// checkpoints, and the fixed random caption subset (host).
//
// Reference: a statistics pass is one process (emcid/layer_stats.py:196-219); its only cross-sample dependency is the
// running sum `mom2 += a.t().mm(a); count += n` (util/runningstats.py:492-493), so R ranks that each visited
// `subset[r::R]` meet in ONE reduction per layer: the lower triangle of the per-rank sums (d (d + 1) / 2 fp32 values,
// 18.9 MB at d = 3072) and the int64 counts are summed onto the rank that writes the layer's npz.  NCCL is resolved at
// run time from the process image (the libnccl.so.2 the host framework already loaded, else the system one): the
// library has no link-time dependency on it and still loads on machines without NCCL.
#pragma once

#include <dlfcn.h>
#include <fcntl.h>
#include <unistd.h>

#include "mom2.cuh"

namespace emcid {

// ---- NCCL, bound lazily ---------------------------------------------------------------------------
typedef int (*PFN_ncclReduce)(const void*, void*, size_t, int, int, int, void*, cudaStream_t);
typedef int (*PFN_ncclBroadcast)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*PFN_ncclGroup)(void);
typedef const char* (*PFN_ncclGetErrorString)(int);
typedef int (*PFN_ncclCommCount)(const void*, int*);

constexpr int NCCL_SUM = 0, NCCL_INT64 = 4, NCCL_FLOAT32 = 7;   // nccl.h: ncclSum, ncclInt64, ncclFloat32

struct NcclApi {
  void* lib = nullptr;
  PFN_ncclReduce reduce = nullptr;
  PFN_ncclBroadcast broadcast = nullptr;
  PFN_ncclGroup group_start = nullptr, group_end = nullptr;
  PFN_ncclGetErrorString error_string = nullptr;
  PFN_ncclCommCount comm_count = nullptr, comm_user_rank = nullptr;
};

inline const NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names)            // the instance that created the caller's communicator, if one is loaded
      if ((lib = dlopen(n, RTLD_NOW | RTLD_NOLOAD))) break;
    if (!lib)
      for (const char* n : names)
        if ((lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
    if (!lib) return;
    api.lib = lib;
    api.reduce = reinterpret_cast<PFN_ncclReduce>(dlsym(lib, "ncclReduce"));
    api.broadcast = reinterpret_cast<PFN_ncclBroadcast>(dlsym(lib, "ncclBroadcast"));
    api.group_start = reinterpret_cast<PFN_ncclGroup>(dlsym(lib, "ncclGroupStart"));
    api.group_end = reinterpret_cast<PFN_ncclGroup>(dlsym(lib, "ncclGroupEnd"));
    api.error_string = reinterpret_cast<PFN_ncclGetErrorString>(dlsym(lib, "ncclGetErrorString"));
    api.comm_count = reinterpret_cast<PFN_ncclCommCount>(dlsym(lib, "ncclCommCount"));
    api.comm_user_rank = reinterpret_cast<PFN_ncclCommCount>(dlsym(lib, "ncclCommUserRank"));
  });
  return (api.lib && api.reduce && api.broadcast && api.group_start && api.group_end && api.comm_count && api.comm_user_rank)
             ? &api : nullptr;
}

#define EMCID_NCCL_CHECK(api, expr)                                                                     \
  do {                                                                                                  \
    int _r = (expr);                                                                                    \
    if (_r != 0)                                                                                        \
      return set_error(EMCID_ERR_CUDA, "%s failed: %s", #expr,                                          \
                       (api)->error_string ? (api)->error_string(_r) : "NCCL error");                   \
  } while (0)

// ---- lower-triangle packing ---------------------------------------------------------------------------
// Row-major packed lower triangle: element (i, j), j <= i, lives at i (i + 1) / 2 + j.
__host__ __device__ inline long long packed_lower_elems(long long d) { return d * (d + 1) / 2; }

// One block per (row, 1024-column chunk); reads of a row are contiguous in both layouts.
template <typename OUT>
__global__ void mom2_pack_lower_kernel(const double* __restrict__ acc64, int d, OUT* __restrict__ packed) {
  const int i = blockIdx.x;
  const long long base = static_cast<long long>(i) * (i + 1) / 2;
  const double* row = acc64 + static_cast<long long>(i) * d;
  for (int j = blockIdx.y * blockDim.x + threadIdx.x; j <= i; j += gridDim.y * blockDim.x)
    packed[base + j] = static_cast<OUT>(row[j]);
}

template <typename IN>
__global__ void mom2_unpack_lower_kernel(const IN* __restrict__ packed, int d, double* __restrict__ acc64) {
  const int i = blockIdx.x;
  const long long base = static_cast<long long>(i) * (i + 1) / 2;
  double* row = acc64 + static_cast<long long>(i) * d;
  for (int j = blockIdx.y * blockDim.x + threadIdx.x; j <= i; j += gridDim.y * blockDim.x)
    row[j] = static_cast<double>(packed[base + j]);
}

inline dim3 mom2_pack_grid(int d) { return dim3(d, (d + 2047) / 2048); }

// Sums the handles' statistics over the ranks of `comm` onto `root`: afterwards the root's handle holds the job-wide
// mom2 (lower triangle, as fp32-rounded per-rank partial sums added by NCCL) and count, the other ranks' handles are
// unchanged.  Stream-ordered on `stream`; every rank of the communicator must call it for the same layer in the same order.
inline int mom2_reduce(Mom2Handle* H, void* comm, int root, cudaStream_t stream) {
  EMCID_CHECK(H && comm, EMCID_ERR_INVALID, "mom2_reduce: null argument");
  const NcclApi* api = nccl_api();
  EMCID_CHECK(api, EMCID_ERR_UNSUPPORTED, "mom2_reduce: libnccl.so.2 is not loadable in this process");
  EMCID_CUDA_CHECK(cudaSetDevice(H->device));
  int nranks = 0, rank = -1;
  EMCID_NCCL_CHECK(api, api->comm_count(comm, &nranks));
  EMCID_NCCL_CHECK(api, api->comm_user_rank(comm, &rank));
  EMCID_CHECK(root >= 0 && root < nranks, EMCID_ERR_INVALID, "mom2_reduce: root %d outside the communicator (%d ranks)", root, nranks);
  int rc = mom2_fold(H, stream);
  if (rc) return rc;
  const long long n = packed_lower_elems(H->d);
  if (!H->packed) {
    cudaError_t e = dev_alloc(reinterpret_cast<void**>(&H->packed), static_cast<size_t>(n) * sizeof(float));
    if (e != cudaSuccess) return set_error(EMCID_ERR_CUDA, "mom2_reduce: cudaMalloc(%lld) failed: %s", n * 4, cudaGetErrorString(e));
  }
  mom2_pack_lower_kernel<float><<<mom2_pack_grid(H->d), 256, 0, stream>>>(H->acc64, H->d, H->packed);
  EMCID_CUDA_CHECK(cudaGetLastError());
  H->launches += 1;
  EMCID_NCCL_CHECK(api, api->group_start());
  int r1 = api->reduce(H->packed, H->packed, static_cast<size_t>(n), NCCL_FLOAT32, NCCL_SUM, root, comm, stream);
  int r2 = api->reduce(H->count, H->count, 1, NCCL_INT64, NCCL_SUM, root, comm, stream);
  EMCID_NCCL_CHECK(api, api->group_end());
  EMCID_NCCL_CHECK(api, r1);
  EMCID_NCCL_CHECK(api, r2);
  if (rank == root) {
    mom2_unpack_lower_kernel<float><<<mom2_pack_grid(H->d), 256, 0, stream>>>(H->packed, H->d, H->acc64);
    EMCID_CUDA_CHECK(cudaGetLastError());
    H->launches += 1;
  }
  return EMCID_OK;
}

// Full mirrored matrix from the root to every rank (for callers that want the statistics everywhere).
inline int mom2_broadcast_full(float* mom2_full, long long* count_dev, int d, void* comm, int root, cudaStream_t stream) {
  EMCID_CHECK(mom2_full && comm && d > 0, EMCID_ERR_INVALID, "mom2_broadcast_full: bad argument");
  const NcclApi* api = nccl_api();
  EMCID_CHECK(api, EMCID_ERR_UNSUPPORTED, "mom2_broadcast_full: libnccl.so.2 is not loadable in this process");
  EMCID_NCCL_CHECK(api, api->group_start());
  int r1 = api->broadcast(mom2_full, mom2_full, static_cast<size_t>(d) * d, NCCL_FLOAT32, root, comm, stream);
  int r2 = count_dev ? api->broadcast(count_dev, count_dev, 1, NCCL_INT64, root, comm, stream) : 0;
  EMCID_NCCL_CHECK(api, api->group_end());
  EMCID_NCCL_CHECK(api, r1);
  EMCID_NCCL_CHECK(api, r2);
  return EMCID_OK;
}

// C[i][j] = C[j][i] for j > i: one 32 x 32 tile per block, lower tiles transposed into their mirror images.
__global__ void symmetrize_lower_kernel(float* __restrict__ C, int d, long long ldc) {
  __shared__ float tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int k = ty; k < 32; k += 8) {
    const int r = bi * 32 + k, c = bj * 32 + tx;
    tile[k][tx] = (r < d && c < d) ? C[static_cast<long long>(r) * ldc + c] : 0.f;
  }
  __syncthreads();
  for (int k = ty; k < 32; k += 8) {
    const int r = bj * 32 + k, c = bi * 32 + tx;  // element (r, c) of the mirrored tile = element (c, r) of the lower one
    if (r < d && c < d && c > r) C[static_cast<long long>(r) * ldc + c] = tile[tx][k];
  }
}

inline int symmetrize_lower(float* C, int d, long long ldc, cudaStream_t stream) {
  EMCID_CHECK(C && d > 0 && ldc >= d, EMCID_ERR_INVALID, "symmetrize_lower: bad argument");
  const int nb = (d + 31) / 32;
  symmetrize_lower_kernel<<<dim3(nb, nb), dim3(32, 8), 0, stream>>>(C, d, ldc);
  EMCID_CUDA_CHECK(cudaGetLastError());
  return EMCID_OK;
}

// ---- content checksums of a set of device tensors, one launch ---------------------------------------------------------
// out[i] = wrapping sum of the 32-bit words of tensor i.  The weight-change detection of the key extraction compares these
// with the values recorded at upload time: `Tensor._version` misses writes through `param.data`, and ~200 separate
// reductions per edit cost more host time than a layer's solve.  table: device array of n (pointer, word count) pairs.
struct ChecksumEntry { const uint32_t* ptr; long long words; };

__global__ void checksum_kernel(const ChecksumEntry* __restrict__ table, unsigned long long* __restrict__ out) {
  const ChecksumEntry e = table[blockIdx.x];
  unsigned long long acc = 0;
  for (long long i = blockIdx.y * static_cast<long long>(blockDim.x) + threadIdx.x; i < e.words;
       i += static_cast<long long>(gridDim.y) * blockDim.x)
    acc += e.ptr[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out + blockIdx.x, acc);
}

inline int checksum_tensors(const void* table_dev, int n, unsigned long long* out_dev, cudaStream_t stream) {
  EMCID_CHECK(table_dev && out_dev && n > 0, EMCID_ERR_INVALID, "checksum_tensors: bad argument");
  EMCID_CUDA_CHECK(cudaMemsetAsync(out_dev, 0, static_cast<size_t>(n) * sizeof(unsigned long long), stream));
  checksum_kernel<<<dim3(n, 64), 256, 0, stream>>>(static_cast<const ChecksumEntry*>(table_dev), out_dev);   // 64 CTAs per tensor: the 152 MB token embedding is one of them
  EMCID_CUDA_CHECK(cudaGetLastError());
  return EMCID_OK;
}

// ---- accumulator state for a resumable pass ---------------------------------------------------------------
// The reference loses a whole pass on a crash: the stat is only saved once the loader is exhausted
// (util/runningstats.py:115-119).  export = fold the fp32 accumulator into the fp64 one, then copy its lower triangle
// (packed fp64, d (d + 1) / 2 values) and the count to caller-owned DEVICE buffers; import = the inverse on a fresh or
// reset handle.  A pass that exports every N blocks and a pass that was killed and resumed from such an export fold at
// the same points, so they finish with the same fp64 sums.
inline int mom2_export_state(Mom2Handle* H, double* lower_packed, long long* count_dev, cudaStream_t stream) {
  EMCID_CHECK(H && lower_packed && count_dev, EMCID_ERR_INVALID, "mom2_export_state: null argument");
  EMCID_CUDA_CHECK(cudaSetDevice(H->device));
  int rc = mom2_fold(H, stream);
  if (rc) return rc;
  mom2_pack_lower_kernel<double><<<mom2_pack_grid(H->d), 256, 0, stream>>>(H->acc64, H->d, lower_packed);
  EMCID_CUDA_CHECK(cudaGetLastError());
  H->launches += 1;
  EMCID_CUDA_CHECK(cudaMemcpyAsync(count_dev, H->count, sizeof(long long), cudaMemcpyDeviceToDevice, stream));
  return EMCID_OK;
}

inline int mom2_import_state(Mom2Handle* H, const double* lower_packed, const long long* count_dev, cudaStream_t stream) {
  EMCID_CHECK(H && lower_packed && count_dev, EMCID_ERR_INVALID, "mom2_import_state: null argument");
  int rc = mom2_reset(H, stream);
  if (rc) return rc;
  mom2_unpack_lower_kernel<double><<<mom2_pack_grid(H->d), 256, 0, stream>>>(lower_packed, H->d, H->acc64);
  EMCID_CUDA_CHECK(cudaGetLastError());
  H->launches += 1;
  EMCID_CUDA_CHECK(cudaMemcpyAsync(H->count, count_dev, sizeof(long long), cudaMemcpyDeviceToDevice, stream));
  return EMCID_OK;
}

// ---- the fixed random caption subset (host) -----------------------------------------------------------------
// WHICH captions a pass visits: FixedRandomSubsetSampler = random.Random(seed).shuffle(list(range(n)))[:sample_size]
// (util/runningstats.py:1551-1556, reached through make_loader :1598-1600 with random_sample = 1).  CPython's
// generator is MT19937 seeded by init_by_array over the 32-bit words of |seed|; shuffle walks i = n-1 .. 1 and swaps
// x[i] with x[_randbelow(i + 1)], where _randbelow(m) draws getrandbits(m.bit_length()) (the top bits of one 32-bit
// output) until the draw is below m.  Restated here because the interpreted shuffle of an 800 k-caption index costs
// 0.26 s per rank, as much as 5 % of a 100 k-caption pass on one B200.
struct Mt19937 {
  uint32_t mt[624];
  int idx;
  void init_genrand(uint32_t s) {
    mt[0] = s;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + static_cast<uint32_t>(i);
    idx = 624;
  }
  void init_by_array(const uint32_t* key, int len) {
    init_genrand(19650218u);
    int i = 1, j = 0;
    for (int k = 624 > len ? 624 : len; k; --k) {
      mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525u)) + key[j] + static_cast<uint32_t>(j);
      if (++i >= 624) { mt[0] = mt[623]; i = 1; }
      if (++j >= len) j = 0;
    }
    for (int k = 623; k; --k) {
      mt[i] = (mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941u)) - static_cast<uint32_t>(i);
      if (++i >= 624) { mt[0] = mt[623]; i = 1; }
    }
    mt[0] = 0x80000000u;
  }
  uint32_t next() {
    if (idx >= 624) {
      for (int k = 0; k < 624; ++k) {
        const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = mt[idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
};

// out[0 .. n_out) = the first n_out entries of random.Random(seed).shuffle(list(range(n_items))).
inline int fixed_random_subset(long long n_items, long long seed, long long* out, long long n_out) {
  EMCID_CHECK(n_items >= 0 && n_items < (1ll << 31) && out && n_out >= 0 && n_out <= n_items, EMCID_ERR_INVALID,
              "fixed_random_subset: bad argument (n_items %lld, n_out %lld)", n_items, n_out);
  unsigned long long a = seed < 0 ? 0ull - static_cast<unsigned long long>(seed) : static_cast<unsigned long long>(seed);
  uint32_t key[2] = {static_cast<uint32_t>(a & 0xffffffffull), static_cast<uint32_t>(a >> 32)};
  Mt19937 g;
  g.init_by_array(key, key[1] ? 2 : 1);
  int32_t* x = static_cast<int32_t*>(malloc(sizeof(int32_t) * static_cast<size_t>(n_items > 0 ? n_items : 1)));
  EMCID_CHECK(x, EMCID_ERR_INVALID, "fixed_random_subset: out of host memory");
  for (long long i = 0; i < n_items; ++i) x[i] = static_cast<int32_t>(i);
  for (long long i = n_items - 1; i >= 1; --i) {
    const uint32_t m = static_cast<uint32_t>(i + 1);
    const int shift = __builtin_clz(m);          // 32 - bit_length(m)
    uint32_t r;
    do { r = g.next() >> shift; } while (r >= m);
    const int32_t t = x[i]; x[i] = x[r]; x[r] = t;
  }
  for (long long i = 0; i < n_out; ++i) out[i] = x[i];
  free(x);
  return EMCID_OK;
}

// ---- the v* cache files of an edit (host) ---------------------------------------------------------------------
// An edit of n concepts starts by reading n files `cache_name + "source_{src}_dest_{dst}.npz"`, each holding one small
// array `v_star` (emcid/emcid_main.py:873-907: numpy.savez(file, v_star=...)).  Read one by one with numpy that is
// 90 us per file of zipfile machinery; a lean Python reader still costs 10 ms per 1000 files, all of it under the GIL next
// to the tokenisation of the prompts.  This reads them in C (the host thread that calls it through ctypes holds no GIL):
// the first member of the archive must be the stored (uncompressed) `<key>.npy`, little-endian float32, C order, with
// exactly `elems` elements; sizes come from the npy header (numpy streams members with zip64 placeholders in the local
// header).  Returns EMCID_OK, or EMCID_ERR_INVALID with *bad = index of the first file that is missing or different —
// the caller falls back to its general reader for the error message or the unusual layout.
inline int read_npz_f32(const char* const* paths, int n, const char* key, float* out, long long elems, int* bad) {
  EMCID_CHECK(paths && key && out && elems > 0 && n >= 0, EMCID_ERR_INVALID, "read_npz_f32: bad argument");
  const size_t klen = strlen(key);
  const size_t want = static_cast<size_t>(elems) * sizeof(float);
  unsigned char head[8192];                    // a 768- or 1280-wide v* file fits: one open / read / close per file
  for (int i = 0; i < n; ++i) {
    if (bad) *bad = i;
    const int fd = open(paths[i], O_RDONLY | O_CLOEXEC);
    if (fd < 0) return set_error(EMCID_ERR_INVALID, "read_npz_f32: cannot open %s", paths[i]);
    const ssize_t got_s = read(fd, head, sizeof(head));
    const size_t got = got_s > 0 ? static_cast<size_t>(got_s) : 0;
    int ok = 0;
    size_t data_off = 0;
    do {
      if (got < 30 + klen + 4 + 12) break;
      if (!(head[0] == 'P' && head[1] == 'K' && head[2] == 3 && head[3] == 4)) break;
      if (head[8] != 0 || head[9] != 0) break;                                   // compression method: stored
      const size_t nlen = head[26] | (head[27] << 8), xlen = head[28] | (head[29] << 8);
      if (nlen != klen + 4 || memcmp(head + 30, key, klen) != 0 || memcmp(head + 30 + klen, ".npy", 4) != 0) break;
      size_t o = 30 + nlen + xlen;
      if (o + 12 > got || memcmp(head + o, "\x93NUMPY", 6) != 0) break;
      const int major = head[o + 6];
      size_t hlen, hstart;
      if (major == 1) { hlen = head[o + 8] | (head[o + 9] << 8); hstart = o + 10; }
      else { hlen = head[o + 8] | (head[o + 9] << 8) | (head[o + 10] << 16) | (static_cast<size_t>(head[o + 11]) << 24); hstart = o + 12; }
      if (hstart + hlen > got) break;
      char text[768];
      if (hlen >= sizeof(text)) break;
      memcpy(text, head + hstart, hlen);
      text[hlen] = 0;
      if (!strstr(text, "'descr': '<f4'") || !strstr(text, "'fortran_order': False")) break;
      const char* sh = strstr(text, "'shape': (");
      if (!sh) break;
      long long prod = 1;
      const char* q = sh + 10;
      while (*q && *q != ')') {
        if (*q >= '0' && *q <= '9') { prod *= strtoll(q, const_cast<char**>(&q), 10); continue; }
        ++q;
      }
      if (*q != ')' || prod != elems) break;
      data_off = hstart + hlen;
      ok = 1;
    } while (0);
    if (ok) {
      float* dst = out + static_cast<size_t>(i) * elems;
      const size_t in_head = got > data_off ? got - data_off : 0;
      const size_t take = in_head < want ? in_head : want;
      memcpy(dst, head + data_off, take);
      size_t have = take;
      while (ok && have < want) {                // the rest straight into the destination
        const ssize_t r = pread(fd, reinterpret_cast<unsigned char*>(dst) + have, want - have,
                                static_cast<off_t>(data_off + have));
        if (r <= 0) ok = 0; else have += static_cast<size_t>(r);
      }
    }
    close(fd);
    if (!ok) return set_error(EMCID_ERR_INVALID, "read_npz_f32: %s is not a plain float32 '%s' archive of %lld elements", paths[i], key, elems);
  }
  if (bad) *bad = -1;
  return EMCID_OK;
}

}  // namespace emcid

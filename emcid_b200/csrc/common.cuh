// emcid_b200 — sm_100a device primitives shared by every kernel in this library.
//
// Thin inline-PTX wrappers around the Blackwell execution model: mbarrier, TMA
// (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld) and the 3xTF32
// operand split.  Nothing here is a port of reference code: the reference
// (SilentView/EMCID) is pure PyTorch and has no device code at all.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/emcid_b200.h"  // EMCID_OK / EMCID_ERR_* codes

namespace emcid {

// ----------------------------------------------------------------------------------------------
// Small helpers
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Round-to-nearest (ties away) fp32 -> tf32, returned in an fp32 container whose low 13
// mantissa bits are zero.  tcgen05.mma.kind::tf32 ignores those bits, so feeding it
// pre-rounded values makes the hardware truncation a no-op.
__device__ __forceinline__ float to_tf32_rna(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// 3xTF32 operand split: x ~= hi + lo with hi, lo both exactly representable in tf32.
// (x - hi) is exact in fp32; the representation error is <= 2^-22 |x|, unbiased.
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = to_tf32_rna(x);
  lo = to_tf32_rna(x - hi);
}

// 3xFP16 operand split for tcgen05.mma.kind::f16 (twice the tf32 issue rate, half the bytes):
//   hi = fp16(x)  (11 significant bits, saturated to the fp16 range so it can never become inf)
//   lo = fp16(x - hi)   (x - hi is exact in fp32; 11 more bits, subnormal below 6.1e-5)
// |x - (hi + lo)| <= max(2^-23 |x|, 2^-25): the same 22 bits as the tf32 split for |x| in [2^-3, 65504],
// an absolute 3e-8 below that.  Static operands (weights) are pre-scaled by an exact power of two so they
// sit in the full-precision range.  Measured on B200: an MMA whose A and B formats differ (fp16 x bf16)
// raises an illegal-instruction trap, so a bf16 lo plane next to an fp16 hi plane is not an option.
constexpr int FMT_F16 = 0, FMT_BF16 = 1, FMT_TF32 = 2;  // UMMA instruction-descriptor operand formats

__device__ __forceinline__ void split_f16(float x, int lo_fmt, uint16_t& hi, uint16_t& lo) {
  const float xs = fminf(fmaxf(x, -65504.f), 65504.f);
  const __half h = __float2half_rn(xs);
  const float r = x - __half2float(h);
  hi = __half_as_ushort(h);
  lo = lo_fmt == FMT_BF16 ? __bfloat16_as_ushort(__float2bfloat16_rn(r)) : __half_as_ushort(__float2half_rn(r));
}

// Two values at once, fp16 hi + fp16 lo, packed as {b : a} (a in the low half): one F2FP pack per plane instead of
// two scalar conversions plus clamps and byte permutes per value (the epilogues and the softmax are conversion-bound).
// Same result as split_f16(.., FMT_F16, ..): round-to-nearest, saturating at +-65504.
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t& hi2, uint32_t& lo2) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(b), "f"(a));
  float fa, fb;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}"
      : "=f"(fa), "=f"(fb) : "r"(hi2));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(b - fb), "f"(a - fa));
}

// ----------------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Word in global memory that kernels set before trapping when a barrier wait exceeds its
// spin budget: a hang becomes a CUDA error the host reports instead of a dead GPU box.
static __device__ unsigned int g_emcid_hang_code = 0;

#ifndef EMCID_SPIN_LIMIT
#define EMCID_SPIN_LIMIT (1u << 28)
#endif

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t dbg_code = 0) {
  const uint32_t addr = smem_u32(bar);
  uint32_t spins = 0;
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (++spins > EMCID_SPIN_LIMIT) {
      g_emcid_hang_code = 0x80000000u | dbg_code;
      __threadfence_system();
      __trap();
    }
  }
}

// ----------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor)
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}

// 2-D tiled load global -> shared, completion signalled on `bar` (complete_tx::bytes).
// c0 = coordinate along the contiguous (innermost) dimension, c1 = row.
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, uint64_t* bar,
                                            int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1)
      : "memory");
}

// 2-D tiled store shared -> global (bulk-group completion).  Elements of the box that fall outside the
// tensor are not written, so ragged last tiles need no masking.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
// Same, but global += shared (element type and add come from the tensor map: fp32 here).
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* tm, uint32_t smem_src, int32_t c0, int32_t c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
               :
               : "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all committed bulk stores of this thread have finished READING shared memory (it may be overwritten)
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// ----------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_result)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// tcgen05.commit: the mbarrier receives one arrival once every MMA issued so far by this
// thread has completed (implies fence::before_thread_sync).
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs, fp32 accumulate, single CTA.
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ----------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): two CTAs of a cluster drive one 256-row MMA; each holds 128 accumulator rows in
// its own TMEM, loads its own 128 rows of A and HALF of B, and the tensor cores read B from both CTAs'
// shared memory — per-CTA shared-memory traffic per MMA drops from 12 KB to 8 KB and the stage from 96 KB to 64 KB.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of this CTA -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrive on a barrier of another CTA of the cluster.  Relaxed: the arrival publishes no generic-proxy data (the epilogue
// warps only tell the MMA issuer that their tcgen05.ld of an accumulator stage has completed, ordered by
// tcgen05.fence::before_thread_sync); with .release.cluster every arrival cost a MEMBAR.ALL.GPU + ERRBAR (6 % of all warp
// samples of the q/k/v projection, on the path that frees the accumulator for the next MMA chunk).
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// Wait on a barrier whose arrivals come from both CTAs of the pair.  Relaxed like the arrivals: the MMA issuer reads no data
// the epilogue warps wrote (tcgen05.fence::after_thread_sync orders the tensor-core side); with .acquire.cluster every
// successful wait invalidated the L1 (CCTL.IVALL), once per accumulator chunk.
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity, uint32_t dbg_code = 0) {
  const uint32_t addr = smem_u32(bar);
  uint32_t spins = 0;
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
    if (done) return;
    if (++spins > EMCID_SPIN_LIMIT) {
      g_emcid_hang_code = 0x80000000u | dbg_code;
      __threadfence_system();
      __trap();
    }
  }
}
// TMA load into THIS CTA's shared memory whose completion bytes are signalled on a barrier of the pair's leader CTA
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* tm, uint32_t leader_bar_cluster_addr,
                                                int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];"
      :
      : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(leader_bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_result, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// commit of the pair's MMAs: one arrival on the barrier at this offset in every CTA of `cta_mask`
__device__ __forceinline__ void tc_commit2(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Same, 16-bit inputs (each operand independently fp16 or bf16, chosen by the instruction descriptor).
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n"
      :
      : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Shared-memory matrix descriptor for a K-major operand tile stored as rows of 128 bytes with
// the 128-byte swizzle (what TMA SWIZZLE_128B writes): 8-row groups are 1024 B apart (SBO),
// LBO is unused for swizzled K-major layouts, version=1 (sm_100), layout_type=2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc_k128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);  // start address, bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                      // LBO (ignored), bits [16,30)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;              // SBO, bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                      // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                      // SWIZZLE_128B
  return d;
}

// MN-major operand tile (rows of the tile = K, 64 MN elements = one 128-byte swizzle row per K row), as TMA writes a box
// of [K rows x 64 MN elements] with SWIZZLE_128B: 8-K-row groups 1024 B apart (SBO); successive 64-element MN atoms are
// separate boxes `atom_bytes` apart (LBO).  The instruction descriptor must flag the operand as MN-major.
__device__ __forceinline__ uint64_t make_desc_mn128(uint32_t smem_addr, uint32_t atom_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(atom_bytes >> 4) << 16;         // LBO: next MN atom
  d |= static_cast<uint64_t>(1024 >> 4) << 32;               // SBO: next group of 8 K rows
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;                       // SWIZZLE_128B
  return d;
}

// Same for tiles of 64-byte rows written with TMA SWIZZLE_64B: 8-row groups are 512 B apart, layout_type = 4.
__device__ __forceinline__ uint64_t make_desc_k64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(512 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(4) << 61;                      // SWIZZLE_64B
  return d;
}

// Instruction descriptor for kind::tf32, fp32 accumulate, both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc_tf32(uint32_t M, uint32_t N, bool neg_a = false) {
  return (1u << 4)                  // c_format = F32
         | (2u << 7)                // a_format = TF32
         | (2u << 10)               // b_format = TF32
         | ((neg_a ? 1u : 0u) << 13)// a_negate
         | ((N >> 3) << 17)         // n_dim
         | ((M >> 4) << 24);        // m_dim
}

// Instruction descriptor with explicit operand formats (FMT_*), fp32 accumulate, both operands K-major.
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt_a, uint32_t fmt_b, uint32_t M, uint32_t N) {
  return (1u << 4) | (fmt_a << 7) | (fmt_b << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

// Same, 16 columns (keeps fewer registers live in register-heavy epilogues).
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Issue only: 16 columns into r[0..15]; the registers are valid after tmem_wait_ld<N>() over the batch they belong to.
__device__ __forceinline__ void tmem_ld_32x16_issue(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// tcgen05.wait::ld over a batch of loads.  The loaded registers are in/out operands of the wait so that no use of them can
// be scheduled above it (the loads are asynchronous: the compiler only sees ordinary register outputs).
template <int N> __device__ __forceinline__ void tmem_wait_ld(uint32_t* r);
template <> __device__ __forceinline__ void tmem_wait_ld<16>(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) :: "memory");
}
template <> __device__ __forceinline__ void tmem_wait_ld<32>(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]) :: "memory");
}
template <> __device__ __forceinline__ void tmem_wait_ld<64>(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]), "+r"(r[32]), "+r"(r[33]), "+r"(r[34]), "+r"(r[35]), "+r"(r[36]), "+r"(r[37]), "+r"(r[38]), "+r"(r[39]), "+r"(r[40]), "+r"(r[41]), "+r"(r[42]), "+r"(r[43]), "+r"(r[44]), "+r"(r[45]), "+r"(r[46]), "+r"(r[47]), "+r"(r[48]), "+r"(r[49]), "+r"(r[50]), "+r"(r[51]), "+r"(r[52]), "+r"(r[53]), "+r"(r[54]), "+r"(r[55]), "+r"(r[56]), "+r"(r[57]), "+r"(r[58]), "+r"(r[59]), "+r"(r[60]), "+r"(r[61]), "+r"(r[62]), "+r"(r[63]) :: "memory");
}

// Vectorised fp32 reduction into global memory (no return value).
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c),
               "f"(d)
               : "memory");
}

}  // namespace emcid

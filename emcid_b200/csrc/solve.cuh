// emcid_b200 — the closed-form update  adj_k = (lambda*C + Ks Ks^T)^-1 Ks,  dW = resid adj_k^T.
//
// Replaces the solve block of execute_emcid_text_encoder (emcid/emcid_main.py:1037-1050; identical
// copies at :1265-1312, :1365-1414, :2016-2042), which is fp64 LU (`torch.linalg.solve`).  Here:
//
//   M32 = lambda*C32 + Ks Ks^T                 3xTF32 tcgen05 SYRK, lower tiles, planes kept
//   M32 = L L^T                                blocked right-looking Cholesky, nb = 128:
//        diag block   fp64 potrf + triangular inverse in shared memory (one CTA per layer, register-blocked:
//                     potrf_diag_kernel_v2)
//        panel        L_ik = M_ik Linv_kk^T            tcgen05 GEMM (K = 128)
//        trailing     M_ij -= L_ik L_jk^T  (i>=j>k)    tcgen05 GEMM, lower tiles; block column k+1 first, the rest on a
//                                                      side stream next to the following diagonal block (look-ahead)
//   Linv = L^-1                                blocked triangular inverse (block columns from the right): two
//                                              tcgen05 GEMMs per block column, zero k-blocks skipped (k_tri)
//   X = Linv^T (Linv Ks)                       two tcgen05 GEMMs per application on the transposed right-hand
//                                              side [n x d] (the explicit inverse replaces ~100 dependent TRSM
//                                              launches per application; what it loses in backward stability
//                                              the fp64 refinement takes back).  EMCID_SOLVE_TRSM=1 keeps the
//                                              blocked TRSM against L / L^T / Linv_kk planes.
//   refinement (fp64 residual on DMMA):        R^T = Ks^T - X^T M64 ; X += solve(R); sweeps until the error predicted
//                                              from the contraction of the corrections is below SOLVE_ADAPT_TOL
//   dW = resid adj_k^T                         fp64 DMMA GEMM, rounded once to fp32
//
// factor_spd / refined_solve are the two halves (factorisation of stacked SPD matrices; refined application to
// transposed right-hand sides); solve_layers is the direct solver on M, factor_create / factor_solve the cached-factor
// form for repeated edits with the same lambda*C (push-through identity, see below).
//
// Everything is batched over independent layers (grid.y / grid.z = layer): the benchmark form of
// SURVEY.md §8d solves the 5 edited layers in one call; the faithful sequential loop calls it with
// batch = 1 per layer because K_{i+1} depends on dW_i (emcid_main.py:1061).
// All right-hand-side / solution matrices are held TRANSPOSED ([n x d], d contiguous) so that every
// operand of every product is K-major, the only layout the TMA/UMMA descriptors here describe.
#pragma once

#include "dgemm.cuh"
#include "gemm_api.cuh"

namespace emcid {

constexpr int SOLVE_NB = 128;
// Adaptive refinement stops on the PREDICTED error left after the sweep just applied.  With c_k = |dx_k| / |x| the size of
// sweep k's correction (c_0 = 1: the unrefined solve) the sweeps contract geometrically, c_k ~ rho c_{k-1} (measured on
// B200: rho = 3e-3 for 1000 concepts at d = 3072, 0.1 for 100 concepts, 0.2 / 0.45 for lambda*C alone at d = 3072 / 5120),
// so what is left after sweep k is ~ rho c_k / (1 - rho).  rho is taken as 1.5 x the largest ratio seen so far (the ratios
// creep up over the first sweeps) and the loop ends when that prediction is below SOLVE_ADAPT_TOL = 2e-5 relative
// Frobenius error of adj_k, a fifth of the 1e-4 dW tolerance.  (Until r03 the rule was "last correction < 1e-4", which spent
// a whole sweep — one fp64 product with M64, 4.3 of 22 ms for 5 layers x 1000 concepts — to confirm what the ratio already
// said, and stopped too early when rho was large.)  EMCID_SOLVE_TOL overrides the target.
constexpr int SOLVE_ADAPT_MAX = 16;
constexpr double SOLVE_ADAPT_TOL = 2e-5;
// the cached-factor path chains two refined solves whose errors add up in adj_k (and the error of Y is amplified by up to
// cond(G) on the way): each runs to a four times smaller target
constexpr double SOLVE_ADAPT_TOL_CHAINED = 5e-6;

inline double solve_adapt_tol(double dflt) {
  static const double env = [] { const char* e = getenv("EMCID_SOLVE_TOL"); return e ? atof(e) : 0.0; }();
  return env > 0.0 ? env * (dflt / SOLVE_ADAPT_TOL) : dflt;
}
// capacity of the split-K slice buffer (doubles): three slices of a 1024 x 5120 right-hand side block (both uses clamp
// their slice count to it)
constexpr long long SOLVE_SPLIT_ELEMS = 3LL * 1024 * 5120;

// ---- diagonal block: potrf + inverse, fp64 in shared memory, register-blocked ----------------------------------------
// M32: [B][d x d] fp32.  Writes L_kk (symmetric fill) into the LL planes' diagonal block, Linv_kk / Linv_kk^T planes into
// row-block k of the [B*d x 128] Linv / LinvT arrays and into the diagonal blocks of the explicit inverse Lv / LvT.
// (A first version worked a column at a time through shared memory, two block-wide barriers per column: 1300 clocks per
// column, 170 k clocks for the factor and 97 k for the inverse of one 128 x 128 block.)  Here the
// block is cut into 32 x 32 sub-blocks, a lane owns one ROW of a sub-block in registers and the other operand of every
// product is a shared-memory broadcast:
//   factor   per 32-column panel: warp 0 factors the diagonal sub-block (no block-wide barrier inside; the rank-1 updates
//            use the UNSCALED column and a Newton reciprocal of the pivot, so the per-column dependency chain is one
//            broadcast + one reciprocal + one FMA, and the 32 square roots are taken at the end, off the chain), the
//            sub-blocks below it are solved one row per lane against the transposed diagonal factor, the trailing
//            sub-blocks get a rank-32 update; three block-wide barriers per panel instead of 64 + 1
//   inverse  diagonal sub-blocks one warp each (right-looking row recurrence in registers), then block rows top-down:
//            T_ij = sum_k L_ik X_kj,  X_ij = -X_ii T_ij
// Every 32 x 32 x 32 product is cut into four 8-column items dealt round-robin to the 8 warps, which keeps the four
// fp64 pipes of the SM evenly loaded (the kernel is DFMA-issue bound there: 2 clocks per warp instruction per quadrant).
// The strictly upper triangles of the diagonal sub-blocks and the upper sub-blocks stay zero from the load on.
// Even pitches: every broadcast operand is fetched as an aligned double2 (the products are bound by shared-memory
// wavefronts, one per broadcast load: LDS.128 halves them); the price is a 2-way bank conflict on the row-per-lane loads,
// of which there are 32 per 256 broadcasts.
constexpr int POTRF2_LD = SOLVE_NB + 2;
constexpr int POTRF2_TP = 34;   // pitch of the T_ij scratch
constexpr int POTRF2_SMEM_DOUBLES = SOLVE_NB * POTRF2_LD + SOLVE_NB + 32 * 32 + 64 + 3 * 32 * POTRF2_TP;

// 1 / d to ~1e-13 from the fp32 approximation and ONE Newton step: the pivot reciprocal sits on the column-to-column
// dependency chain of the diagonal factorisation, where every dependent fp64 operation costs ~40 clocks (measured: 380
// clocks per column with the shuffle + two-step version), and the factor is rounded to 22-bit planes afterwards anyway.
// d outside the fp32 range takes the division.
__device__ __forceinline__ double potrf_fast_rcp(double d) {
  const double ad = fabs(d);
  if (!(ad > 1e-30 && ad < 1e30)) return 1.0 / d;
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(__double2float_rn(d)));
  const double y = static_cast<double>(r);
  return fma(y, fma(-d, y, 1.0), y);
}

// s[u] += sum_t p[t] * Q[u * qs_u + t], u = 0..7: operand rows contiguous in t (Q 16-byte aligned, qs_u even)
__device__ __forceinline__ void warp_rowdot8_tc(const double (&p)[32], const double* __restrict__ Q, int qs_u, double (&s)[8]) {
#pragma unroll
  for (int t = 0; t < 32; t += 2) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const double2 q = *reinterpret_cast<const double2*>(Q + u * qs_u + t);
      s[u] = fma(p[t], q.x, s[u]);
      s[u] = fma(p[t + 1], q.y, s[u]);
    }
  }
}

// s[u] += sum_t p[t] * Q[t * qs_t + u], u = 0..7: operand rows contiguous in u (Q 16-byte aligned, qs_t even)
__device__ __forceinline__ void warp_rowdot8_uc(const double (&p)[32], const double* __restrict__ Q, int qs_t, double (&s)[8]) {
#pragma unroll
  for (int t = 0; t < 32; ++t) {
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const double2 q = *reinterpret_cast<const double2*>(Q + t * qs_t + 2 * v);
      s[2 * v] = fma(p[t], q.x, s[2 * v]);
      s[2 * v + 1] = fma(p[t], q.y, s[2 * v + 1]);
    }
  }
}

__global__ void __launch_bounds__(256, 1) potrf_diag_kernel_v2(const float* __restrict__ M32, int d, int k,
                                                            float* __restrict__ LL_hi, float* __restrict__ LL_lo,
                                                            float* __restrict__ Li_hi, float* __restrict__ Li_lo,
                                                            float* __restrict__ LiT_hi, float* __restrict__ LiT_lo,
                                                            float* __restrict__ Lv_hi, float* __restrict__ Lv_lo,
                                                            float* __restrict__ LvT_hi, float* __restrict__ LvT_lo,
                                                            int* __restrict__ status) {
  extern __shared__ __align__(16) double sm[];
  constexpr int LD = POTRF2_LD, TP = POTRF2_TP;
  double* A = sm;                          // [128][130]
  double* invd = sm + SOLVE_NB * LD;       // [128] reciprocal diagonal of L
  double* LdT = invd + SOLVE_NB;           // [32][32]: LdT[j * 32 + c] = L_cj of the current diagonal sub-block
  double* cb = LdT + 32 * 32;              // [2][32] column broadcast buffer of the diagonal factorisation
  double* Tb = cb + 64;                    // [3][32][34] T_ij of one block row of the inverse
  const int b = blockIdx.x;
  const long long base = static_cast<long long>(b) * d * d + static_cast<long long>(k) * SOLVE_NB * d + k * SOLVE_NB;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
#ifdef EMCID_POTRF_TIMING
  long long tk[8]; tk[0] = clock64();
  long long tph[3] = {0, 0, 0}, tlast;
#define EMCID_TK(i) tk[i] = clock64()
#define EMCID_TPH(i) do { const long long now_ = clock64(); tph[i] += now_ - tlast; tlast = now_; } while (0)
#else
#define EMCID_TK(i)
#define EMCID_TPH(i)
#endif
  // ---- load: one row per warp instruction (float4 per lane), all 16 loads of a warp in flight together
  {
    float4 v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
      v[i] = *reinterpret_cast<const float4*>(M32 + base + static_cast<long long>(wid + 8 * i) * d + 4 * lane);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int r = wid + 8 * i, c = 4 * lane;
      double* row = A + r * LD + c;
      row[0] = (c <= r) ? static_cast<double>(v[i].x) : 0.0;
      row[1] = (c + 1 <= r) ? static_cast<double>(v[i].y) : 0.0;
      row[2] = (c + 2 <= r) ? static_cast<double>(v[i].z) : 0.0;
      row[3] = (c + 3 <= r) ? static_cast<double>(v[i].w) : 0.0;
    }
  }
  __syncthreads();
  EMCID_TK(1);
#ifdef EMCID_POTRF_TIMING
  tlast = tk[1];
#endif
  // ---- factor
  for (int p = 0; p < 4; ++p) {
    const int p0 = 32 * p;
    if (wid == 0) {
      // diagonal sub-block: lane = row r, the row in registers.  Column j stays unscaled (u_rj, pivot d_j = u_jj):
      //   a_rc -= u_rj u_cj / d_j   for c > j;      L_rj = u_rj / sqrt(d_j) once all columns are done
      double a[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = A[(p0 + lane) * LD + p0 + c];
      bool bad = false;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        double* col = cb + (j & 1) * 32;
        col[lane] = a[j];
        double dj = __shfl_sync(0xffffffffu, a[j], j);   // the pivot by shuffle: shorter than the shared-memory round trip
        __syncwarp();
        if (!(dj > 0.0)) { bad = true; dj = 1.0; }   // also catches NaN
        const double w = a[j] * potrf_fast_rcp(dj);
        if (((j + 1) & 1) && j + 1 < 32) a[j + 1] = fma(-w, col[j + 1], a[j + 1]);
#pragma unroll
        for (int c = (j + 2) & ~1; c < 32; c += 2) {
          const double2 q = *reinterpret_cast<const double2*>(col + c);
          a[c] = fma(-w, q.x, a[c]);
          a[c + 1] = fma(-w, q.y, a[c + 1]);
        }
      }
      if (bad && lane == 0) atomicOr(status, 1);
      double dg = 1.0;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (lane == j) dg = a[j];
      if (!(dg > 0.0)) dg = 1.0;
      const double sj = rsqrt(dg);
      invd[p0 + lane] = sj;
      __syncwarp();
      cb[lane] = sj;
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        if (c <= lane) {
          const double l = (c == lane) ? dg * sj : a[c] * cb[c];
          A[(p0 + lane) * LD + p0 + c] = l;
          LdT[c * 32 + lane] = l;
        }
      }
    }
    __syncthreads();
    EMCID_TPH(0);
    if (p == 3) break;
    const int nb = 3 - p;
    if (wid < nb) {
      // sub-block below: lane = row i,  x_j = (a_j - sum_{t<j} x_t L_jt) / L_jj, right-looking
      const int i = p0 + 32 * (1 + wid) + lane;
      double a[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) a[c] = A[i * LD + p0 + c];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const double x = a[j] * invd[p0 + j];
        a[j] = x;
        if (((j + 1) & 1) && j + 1 < 32) a[j + 1] = fma(-x, LdT[j * 32 + j + 1], a[j + 1]);
#pragma unroll
        for (int c = (j + 2) & ~1; c < 32; c += 2) {
          const double2 q = *reinterpret_cast<const double2*>(LdT + j * 32 + c);
          a[c] = fma(-x, q.x, a[c]);
          a[c + 1] = fma(-x, q.y, a[c + 1]);
        }
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) A[i * LD + p0 + c] = a[c];
    }
    __syncthreads();
    EMCID_TPH(1);
    {
      // trailing sub-blocks (I, K), p < K <= I <= 3:  A_IK -= P_I P_K^T with P = the panel just finished;
      // item q = 4 * (sub-block) + (8-column chunk)
      const int items = 2 * nb * (nb + 1);
      for (int q = wid; q < items; q += 8) {
        int I = p + 1, w = q >> 2;
        while (w > I - (p + 1)) { w -= I - p; ++I; }
        const int K = p + 1 + w, c0 = 8 * (q & 3);
        double pr[32], s[8];
        const int r = 32 * I + lane;
#pragma unroll
        for (int t = 0; t < 32; ++t) pr[t] = A[r * LD + p0 + t];
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] = 0.0;
        warp_rowdot8_tc(pr, A + (32 * K + c0) * LD + p0, LD, s);
        double* dst = A + r * LD + 32 * K + c0;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (I != K || c0 + u <= lane) dst[u] -= s[u];
      }
    }
    __syncthreads();
    EMCID_TPH(2);
  }
  EMCID_TK(2);
  // ---- L_kk -> LL planes (lower = L, upper = L^T)
  for (int e = tid; e < SOLVE_NB * SOLVE_NB; e += blockDim.x) {
    const int r = e / SOLVE_NB, c = e % SOLVE_NB;
    const double v = (c <= r) ? A[r * LD + c] : A[c * LD + r];
    float hi, lo;
    split_tf32(static_cast<float>(v), hi, lo);
    LL_hi[base + static_cast<long long>(r) * d + c] = hi;
    LL_lo[base + static_cast<long long>(r) * d + c] = lo;
  }
  __syncthreads();
  EMCID_TK(3);
  // ---- inverse X = L^-1 in place.  (1) diagonal sub-blocks, one warp each: from X L = I,
  //        x_rk = (delta_rk - sum_{m>k} x_rm L_mk) / L_kk,  k = r .. 0;   acc[j] gathers the sums right-looking
  if (wid < 4) {
    const int r0 = 32 * wid;
    double acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.0;
#pragma unroll
    for (int kk = 31; kk >= 0; --kk) {
      double x = ((lane == kk) ? 1.0 : 0.0) - acc[kk];
      x = (kk > lane) ? 0.0 : x * invd[r0 + kk];
      acc[kk] = x;
      const double* lrow = A + (r0 + kk) * LD + r0;
#pragma unroll
      for (int j = 0; j + 1 < kk; j += 2) {
        const double2 q = *reinterpret_cast<const double2*>(lrow + j);
        acc[j] = fma(x, q.x, acc[j]);
        acc[j + 1] = fma(x, q.y, acc[j + 1]);
      }
      if (kk & 1) acc[kk - 1] = fma(x, lrow[kk - 1], acc[kk - 1]);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j <= lane) A[(r0 + lane) * LD + r0 + j] = acc[j];
  }
  __syncthreads();
  EMCID_TK(4);
  // (2) block rows top-down:  T_ij = sum_{k=j..i-1} L_ik X_kj,  X_ij = -X_ii T_ij;  item q = 4 * j + (8-column chunk)
  for (int i = 1; i < 4; ++i) {
    for (int q = wid; q < 4 * i; q += 8) {
      const int j = q >> 2, c0 = 8 * (q & 3);
      const int r = 32 * i + lane;
      double s[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) s[u] = 0.0;
      for (int kb = j; kb < i; ++kb) {
        double pr[32];
#pragma unroll
        for (int t = 0; t < 32; ++t) pr[t] = A[r * LD + 32 * kb + t];
        warp_rowdot8_uc(pr, A + (32 * kb) * LD + 32 * j + c0, LD, s);
      }
      double* dst = Tb + j * (32 * TP) + lane * TP + c0;
#pragma unroll
      for (int u = 0; u < 8; ++u) dst[u] = s[u];
    }
    __syncthreads();
    for (int q = wid; q < 4 * i; q += 8) {
      const int j = q >> 2, c0 = 8 * (q & 3);
      const int r = 32 * i + lane;
      double pr[32], s[8];
#pragma unroll
      for (int t = 0; t < 32; ++t) pr[t] = A[r * LD + 32 * i + t];      // row of X_ii (upper part zero)
#pragma unroll
      for (int u = 0; u < 8; ++u) s[u] = 0.0;
      warp_rowdot8_uc(pr, Tb + j * (32 * TP) + c0, TP, s);
      double* dst = A + r * LD + 32 * j + c0;
#pragma unroll
      for (int u = 0; u < 8; ++u) dst[u] = -s[u];
    }
    __syncthreads();
  }
  EMCID_TK(5);
  // ---- Linv_kk / Linv_kk^T planes (and the diagonal block of the explicit inverse): four columns per thread, float4 stores
  const long long ibase = (static_cast<long long>(b) * d + static_cast<long long>(k) * SOLVE_NB) * SOLVE_NB;
  for (int e = tid; e < SOLVE_NB * SOLVE_NB / 4; e += blockDim.x) {
    const int r = e / (SOLVE_NB / 4), c = 4 * (e % (SOLVE_NB / 4));
    float4 h4, l4, ht4, lt4;
    split_tf32(static_cast<float>((c <= r) ? A[r * LD + c] : 0.0), h4.x, l4.x);
    split_tf32(static_cast<float>((c + 1 <= r) ? A[r * LD + c + 1] : 0.0), h4.y, l4.y);
    split_tf32(static_cast<float>((c + 2 <= r) ? A[r * LD + c + 2] : 0.0), h4.z, l4.z);
    split_tf32(static_cast<float>((c + 3 <= r) ? A[r * LD + c + 3] : 0.0), h4.w, l4.w);
    split_tf32(static_cast<float>((r <= c) ? A[c * LD + r] : 0.0), ht4.x, lt4.x);
    split_tf32(static_cast<float>((r <= c + 1) ? A[(c + 1) * LD + r] : 0.0), ht4.y, lt4.y);
    split_tf32(static_cast<float>((r <= c + 2) ? A[(c + 2) * LD + r] : 0.0), ht4.z, lt4.z);
    split_tf32(static_cast<float>((r <= c + 3) ? A[(c + 3) * LD + r] : 0.0), ht4.w, lt4.w);
    *reinterpret_cast<float4*>(Li_hi + ibase + r * SOLVE_NB + c) = h4;
    *reinterpret_cast<float4*>(Li_lo + ibase + r * SOLVE_NB + c) = l4;
    *reinterpret_cast<float4*>(LiT_hi + ibase + r * SOLVE_NB + c) = ht4;
    *reinterpret_cast<float4*>(LiT_lo + ibase + r * SOLVE_NB + c) = lt4;
    if (Lv_hi) {   // diagonal block of the explicit inverse
      const long long o = base + static_cast<long long>(r) * d + c;
      *reinterpret_cast<float4*>(Lv_hi + o) = h4;
      *reinterpret_cast<float4*>(Lv_lo + o) = l4;
      *reinterpret_cast<float4*>(LvT_hi + o) = ht4;
      *reinterpret_cast<float4*>(LvT_lo + o) = lt4;
    }
  }
#ifdef EMCID_POTRF_TIMING
  __syncthreads();
  EMCID_TK(6);
  if (tid == 0 && b == 0 && k == 3)
    printf("potrf2 clk: load %lld chol %lld (diag %lld below %lld trailing %lld) LLwrite %lld diaginv %lld blkinv %lld out %lld "
           "total %lld\n", tk[1] - tk[0], tk[2] - tk[1], tph[0], tph[1], tph[2], tk[3] - tk[2], tk[4] - tk[3], tk[5] - tk[4],
           tk[6] - tk[5], tk[6] - tk[0]);
#endif
#undef EMCID_TK
#undef EMCID_TPH
}

// ---- operand preparation ----------------------------------------------------------------------------
// Kt [B][n x d] fp32 -> Ks64t [B][n_pad x d] (= s*K, zero pad rows), W / Wp (fp32 + planes of the
// same), Kd64 [B][d x n_pad] and Kd planes (transposes).  32x32 tiles through shared memory.
__global__ void solve_prep_kernel(const float* __restrict__ Kt, long long ldk, long long k_batch, int n, int n_pad,
                                  int d, double s, double* __restrict__ Ks64t, float* __restrict__ W,
                                  float* __restrict__ Wp_hi, float* __restrict__ Wp_lo, double* __restrict__ Kd64,
                                  float* __restrict__ Kd_hi, float* __restrict__ Kd_lo) {
  __shared__ double tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, i0 = blockIdx.x * 32;  // c: concept row, i: feature column
  const int tx = threadIdx.x, ty = threadIdx.y;          // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, i = i0 + tx;
    double v = 0.0;
    if (c < n && i < d) v = s * static_cast<double>(Kt[b * k_batch + static_cast<long long>(c) * ldk + i]);
    tile[r][tx] = v;
    if (c < n_pad && i < d) {
      const long long o = (static_cast<long long>(b) * n_pad + c) * d + i;
      Ks64t[o] = v;
      const float f = static_cast<float>(v);
      float hi, lo;
      split_tf32(f, hi, lo);
      W[o] = f; Wp_hi[o] = hi; Wp_lo[o] = lo;
    }
  }
  __syncthreads();
  if (!Kd64) return;   // cached-factor path: K K^T is never formed, the [d x n] copies are not needed
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, c = c0 + tx;
    if (i < d && c < n_pad) {
      const double v = tile[tx][r];
      const long long o = (static_cast<long long>(b) * d + i) * n_pad + c;
      Kd64[o] = v;
      float hi, lo;
      split_tf32(static_cast<float>(v), hi, lo);
      Kd_hi[o] = hi; Kd_lo[o] = lo;
    }
  }
}

// X64t (+)= double(W)
__global__ void solve_axpy_kernel(const float* __restrict__ W, double* __restrict__ X, long long total, int accumulate) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double w = static_cast<double>(W[i]);
    X[i] = accumulate ? X[i] + w : w;
  }
}

// out[0] += sum(W^2), out[1] += sum(X^2)   (convergence monitor of the refinement)
__global__ void solve_norms_kernel(const float* __restrict__ W, const double* __restrict__ X, long long total,
                                   double* __restrict__ out) {
  double sw = 0.0, sx = 0.0;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const double w = static_cast<double>(W[i]);
    const double x = X[i];
    sw += w * w;
    sx += x * x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out, sw);
    atomicAdd(out + 1, sx);
  }
}

// in-place mirror of the lower triangle of [B][d x d] fp64
__global__ void mirror64_kernel(double* __restrict__ M, int d) {
  __shared__ double tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj >= bi && !(bj == bi)) return;
  double* Mb = M + static_cast<long long>(blockIdx.z) * d * d;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int r = ty; r < 32; r += 8) tile[r][tx] = Mb[static_cast<long long>(bi * 32 + r) * d + bj * 32 + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int row = bj * 32 + r, col = bi * 32 + tx;  // transposed position
    if (bi != bj || col > row) Mb[static_cast<long long>(row) * d + col] = tile[tx][r];
  }
}

// adj_k [B][d x n] = X64t[B][:n, :]^T
__global__ void transpose_out_kernel(const double* __restrict__ Xt, int n, int n_pad, int d, double* __restrict__ out) {
  __shared__ double tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, i0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, i = i0 + tx;
    tile[r][tx] = (c < n && i < d) ? Xt[(static_cast<long long>(b) * n_pad + c) * d + i] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int i = i0 + r, c = c0 + tx;
    if (i < d && c < n) out[(static_cast<long long>(b) * d + i) * n + c] = tile[tx][r];
  }
}

// resid [B][h x n] = s * inv_left[b] * double(St[B][n x h])^T
__global__ void resid_kernel(const float* __restrict__ St, long long lds, long long s_batch, int n, int h, double s,
                             const double* __restrict__ inv_left, double* __restrict__ resid) {
  __shared__ double tile[32][33];
  const int b = blockIdx.z;
  const int c0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const double f = s * inv_left[b];
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, j = j0 + tx;
    // reference order: (S.double() * s) / (L - i)
    tile[r][tx] = (c < n && j < h) ? static_cast<double>(St[b * s_batch + static_cast<long long>(c) * lds + j]) : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int j = j0 + r, c = c0 + tx;
    if (j < h && c < n) resid[(static_cast<long long>(b) * h + j) * n + c] = tile[tx][r] * f;
  }
}

// status |= bits   (the host found something the caller should know about: see refined_solve)
__global__ void status_or_kernel(int* __restrict__ status, int bits) { atomicOr(status, bits); }

// ---- kernels of the cached-factor path ----------------------------------------------------------------
// W = float(R), planes = split(W)   (fp64 right-hand side -> operand of the fp32-class application)
__global__ void rhs_from64_kernel(const double* __restrict__ R, long long total, float* __restrict__ W,
                                  float* __restrict__ hi, float* __restrict__ lo) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float f = static_cast<float>(R[i]);
    float h, l;
    split_tf32(f, h, l);
    W[i] = f; hi[i] = h; lo[i] = l;
  }
}

// W = float(R0 + sum_z P[z]), planes = split(W)   (split-K slices of -X^T M64 -> next right-hand side of the refinement)
__global__ void residual_reduce_kernel(const double* __restrict__ R0, const double* __restrict__ P, int split,
                                       long long stride, long long total, float* __restrict__ W, float* __restrict__ hi,
                                       float* __restrict__ lo) {
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    double v = R0[i];
    for (int z = 0; z < split; ++z) v += P[z * stride + i];
    const float f = static_cast<float>(v);
    float h, l;
    split_tf32(f, h, l);
    W[i] = f; hi[i] = h; lo[i] = l;
  }
}

// A32 = float(lambda) * C32 (+ planes) and A64 = lambda * double(C32), lower 32x32 tiles mirrored to the upper ones
// (same roundings as the direct path: fp32 product in the SYRK epilogue, fp64 product in the DMMA epilogue).
__global__ void factor_prep_kernel(const float* __restrict__ C32, int d, double lambda, float* __restrict__ A32,
                                   float* __restrict__ Ap_hi, float* __restrict__ Ap_lo, double* __restrict__ A64) {
  __shared__ float tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const float lf = static_cast<float>(lambda);
  for (int r = ty; r < 32; r += 8) tile[r][tx] = C32[static_cast<long long>(bi * 32 + r) * d + bj * 32 + tx];
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    float c = tile[r][tx];
    if (bi == bj && tx > r) c = tile[tx][r];
    const long long o = static_cast<long long>(bi * 32 + r) * d + bj * 32 + tx;
    const float a = lf * c;
    float h, l;
    split_tf32(a, h, l);
    A32[o] = a; Ap_hi[o] = h; Ap_lo[o] = l;
    A64[o] = lambda * static_cast<double>(c);
    if (bi != bj) {
      const float ct = tile[tx][r];
      const long long ot = static_cast<long long>(bj * 32 + r) * d + bi * 32 + tx;
      const float at = lf * ct;
      split_tf32(at, h, l);
      A32[ot] = at; Ap_hi[ot] = h; Ap_lo[ot] = l;
      A64[ot] = lambda * static_cast<double>(ct);
    }
  }
}

// G (lower 32x32 tiles of the raw product Ks^T Y; split > 0: the sum of `split` split-K slices P) -> G64 = I + that,
// mirrored to a full symmetric matrix (in place when split == 0), with its fp32 copy and planes.
__global__ void g_prepare_kernel(double* __restrict__ G64, int m, const double* __restrict__ P, int split, long long stride,
                                 float* __restrict__ G32, float* __restrict__ Gp_hi, float* __restrict__ Gp_lo) {
  __shared__ double tile[32][33];
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int r = ty; r < 32; r += 8) {
    const long long o = static_cast<long long>(bi * 32 + r) * m + bj * 32 + tx;
    double v = 0.0;
    if (split > 0) {
      for (int z = 0; z < split; ++z) v += P[z * stride + o];
    } else {
      v = G64[o];
    }
    tile[r][tx] = v;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    double v = tile[r][tx];
    if (bi == bj) {
      if (tx > r) v = tile[tx][r];
      if (tx == r) v += 1.0;
    }
    const long long o = static_cast<long long>(bi * 32 + r) * m + bj * 32 + tx;
    float f = static_cast<float>(v), h, l;
    split_tf32(f, h, l);
    G64[o] = v; G32[o] = f; Gp_hi[o] = h; Gp_lo[o] = l;
    if (bi != bj) {
      const double vt = tile[tx][r];
      const long long ot = static_cast<long long>(bj * 32 + r) * m + bi * 32 + tx;
      f = static_cast<float>(vt);
      split_tf32(f, h, l);
      G64[ot] = vt; G32[ot] = f; Gp_hi[ot] = h; Gp_lo[ot] = l;
    }
  }
}

// out [rows x n] = in [rows x n_pad][:, :n]
__global__ void compact_cols_kernel(const double* __restrict__ in, long long rows, int n, int n_pad,
                                    double* __restrict__ out) {
  const long long total = rows * n;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / n;
    const int c = static_cast<int>(i - r * n);
    out[i] = in[r * n_pad + c];
  }
}

// ---- 3xTF32 GEMM helper on sub-matrices of stacked batched tensors --------------------------------------
struct PlaneMaps {
  CUtensorMap hi, lo;
};

inline int make_plane_maps(PlaneMaps* m, const float* hi, const float* lo, long long rows, long long cols, long long ld) {
  int rc;
  if ((rc = make_tmap_2d(&m->hi, hi, rows, cols, ld, 128))) return rc;
  return make_tmap_2d(&m->lo, lo, rows, cols, ld, 128);
}

struct SubGemm {
  const PlaneMaps* A; int a_row0, a_col0, a_batch_rows;
  const PlaneMaps* B; int b_row0, b_col0, b_batch_rows;
  int M, N, K, lower;
  int k_tri;   // see GemmParams::k_tri
  float alpha, beta;
  const float* Cin; long long ldcin, cin_batch;
  float* C; long long ldc, c_batch;
  float* P_hi; float* P_lo; long long ldp, p_batch;
  float* Pt_hi; float* Pt_lo; long long ldpt, pt_batch;
};

inline int run_subgemm(const SubGemm& g, int batches, int sm_count, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return EMCID_OK;
  GemmOperands ops;
  ops.a_hi = g.A->hi; ops.a_lo = g.A->lo; ops.b_hi = g.B->hi; ops.b_lo = g.B->lo;
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.a_row0 = g.a_row0; p.a_col0 = g.a_col0; p.a_batch_rows = g.a_batch_rows;
  p.b_row0 = g.b_row0; p.b_col0 = g.b_col0; p.b_batch_rows = g.b_batch_rows;
  p.lower = g.lower;
  p.k_tri = g.k_tri;
  p.chunk_kblocks = 1;  // shortest TMEM chains: the factorisation wants every bit 3xTF32 can give
  p.alpha = g.alpha; p.beta = g.beta;
  p.Cin = g.Cin; p.ldcin = g.ldcin; p.cin_batch = g.cin_batch;
  p.C = g.C; p.ldc = g.ldc; p.c_batch = g.c_batch;
  p.P_hi = g.P_hi; p.P_lo = g.P_lo; p.ldp = g.ldp; p.p_batch = g.p_batch;
  p.Pt_hi = g.Pt_hi; p.Pt_lo = g.Pt_lo; p.ldpt = g.ldpt; p.pt_batch = g.pt_batch;
  const int tiles = gemm_num_tiles(g.M, g.N, 128, g.lower);
  int grid = tiles;
  const int cap = sm_count / (batches > 0 ? batches : 1);
  if (grid > cap) grid = cap > 0 ? cap : 1;
  return launch_gemm3x<128, 3, EPI_GENERIC>(ops, p, grid, stream, batches);
}

inline int potrf_smem_bytes() { return POTRF2_SMEM_DOUBLES * static_cast<int>(sizeof(double)); }

inline int potrf_configure(int device) {
  static thread_local bool configured[16] = {false};
  if (device < 0 || device >= 16 || !configured[device]) {
    EMCID_CUDA_CHECK(cudaFuncSetAttribute(potrf_diag_kernel_v2, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          potrf_smem_bytes()));
    if (device >= 0 && device < 16) configured[device] = true;
  }
  return EMCID_OK;
}

// ---- blocked factorisation of B stacked SPD [dim x dim] matrices ------------------------------------------
// In: M32 (lower tiles) and its planes Mp.  Out: L / L^T planes in LL, the diagonal blocks' inverses in Li / LiT and the
// explicit inverse L^-1 / L^-T planes in Lv / LvT.  M32 / Mp are overwritten by the trailing updates.  (Blocked TRSM sweeps
// instead of the explicit inverse cost ~100 dependent launches per application: 47 vs 27 ms for the 5-layer solve.)
struct FactorCtx {
  int B, dim, sms;
  float *M32, *Mp_hi, *Mp_lo, *LL_hi, *LL_lo, *Li_hi, *Li_lo, *LiT_hi, *LiT_lo;
  float *Lv_hi, *Lv_lo, *LvT_hi, *LvT_lo;   // explicit inverse L^-1 and its transpose (planes)
  float *Tp_hi, *Tp_lo;                     // [B][dim x 128] scratch of the inverse's block-row step (T^T)
  PlaneMaps mMp, mLL, mLi, mLiT, mLv, mLvT, mTp;
};

// maps of the explicit inverse only (all an application needs)
inline int factor_make_inverse_maps(FactorCtx& f) {
  const long long rows = static_cast<long long>(f.B) * f.dim;
  int rc;
  if ((rc = make_plane_maps(&f.mLv, f.Lv_hi, f.Lv_lo, rows, f.dim, f.dim))) return rc;
  return make_plane_maps(&f.mLvT, f.LvT_hi, f.LvT_lo, rows, f.dim, f.dim);
}

inline int factor_make_maps(FactorCtx& f) {
  const long long rows = static_cast<long long>(f.B) * f.dim;
  int rc;
  if ((rc = make_plane_maps(&f.mMp, f.Mp_hi, f.Mp_lo, rows, f.dim, f.dim)) ||
      (rc = make_plane_maps(&f.mLL, f.LL_hi, f.LL_lo, rows, f.dim, f.dim)) ||
      (rc = make_plane_maps(&f.mLi, f.Li_hi, f.Li_lo, rows, SOLVE_NB, SOLVE_NB)) ||
      (rc = make_plane_maps(&f.mLiT, f.LiT_hi, f.LiT_lo, rows, SOLVE_NB, SOLVE_NB)) ||
      (rc = make_plane_maps(&f.mTp, f.Tp_hi, f.Tp_lo, rows, SOLVE_NB, SOLVE_NB)))
    return rc;
  return factor_make_inverse_maps(f);
}

// planes whose untouched halves a TMA box may still read must not hold NaN patterns
inline int factor_clear(const FactorCtx& f, cudaStream_t stream) {
  const size_t bytes = static_cast<size_t>(f.B) * f.dim * f.dim * sizeof(float);
  EMCID_CUDA_CHECK(cudaMemsetAsync(f.LL_hi, 0, bytes, stream));
  EMCID_CUDA_CHECK(cudaMemsetAsync(f.LL_lo, 0, bytes, stream));
  // the strictly upper (Lv) / lower (LvT) blocks are never written and never read by the k_tri-limited products;
  // zero them anyway so that no stale NaN pattern can ever reach a TMA box
  EMCID_CUDA_CHECK(cudaMemsetAsync(f.Lv_hi, 0, bytes, stream));
  EMCID_CUDA_CHECK(cudaMemsetAsync(f.Lv_lo, 0, bytes, stream));
  EMCID_CUDA_CHECK(cudaMemsetAsync(f.LvT_hi, 0, bytes, stream));
  EMCID_CUDA_CHECK(cudaMemsetAsync(f.LvT_lo, 0, bytes, stream));
  return EMCID_OK;
}

// side streams + events of the factorisation (one set per device and host thread, never destroyed):
//   stream / fork / join      the look-ahead's "rest of the trailing update"
//   inv / inv_fork / inv_join the explicit inverse, one block row behind the factorisation
//   aux / aux_fork / aux_join work the solve only needs after the factorisation (the fp64 matrix of the refinement)
struct SolveSide {
  cudaStream_t stream = nullptr, inv = nullptr, aux = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr, inv_fork = nullptr, inv_join = nullptr, aux_fork = nullptr, aux_join = nullptr;
};

inline int solve_side(int device, SolveSide** out) {
  static thread_local SolveSide sides[16];
  EMCID_CHECK(device >= 0 && device < 16, EMCID_ERR_UNSUPPORTED, "solve: device index %d out of range", device);
  SolveSide& s = sides[device];
  if (!s.stream) {
    EMCID_CUDA_CHECK(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    EMCID_CUDA_CHECK(cudaStreamCreateWithFlags(&s.inv, cudaStreamNonBlocking));
    EMCID_CUDA_CHECK(cudaStreamCreateWithFlags(&s.aux, cudaStreamNonBlocking));
    for (cudaEvent_t* e : {&s.fork, &s.join, &s.inv_fork, &s.inv_join, &s.aux_fork, &s.aux_join})
      EMCID_CUDA_CHECK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
  }
  *out = &s;
  return EMCID_OK;
}

// Where the fp64 matrix of the refinement is formed.  One or two stacked problems leave most of the device idle during
// the factorisation: M64 goes to a side stream there, as a persistent launch of sms / 2 CTAs (one per SM on half the SMs:
// the diagonal-block kernel needs a whole SM to itself, 253 registers x 256 threads).  From three problems on the chain's
// own GEMMs fill the device and the side product only delays them.  Measured on B200, ms per solve, in front / beside:
// B = 1: 5.36 / 4.96, B = 2: 7.46 / 7.22 - 7.41, B = 5: 13.83 / 14.3 (profiles/round2/r05b_solve_variants.txt).
// EMCID_SOLVE_M64_CTAS = n > 0 forces the side stream with n CTAs, < 0 forces "in front of the chain".
inline int solve_m64_ctas(int sms, int B) {
  static const int env = [] { const char* e = getenv("EMCID_SOLVE_M64_CTAS"); return e ? atoi(e) : 0; }();
  if (env != 0) return env < sms - 2 * B - 8 ? env : sms - 2 * B - 8;
  return B <= 2 ? sms / 2 : -1;
}

inline bool solve_use_lookahead() {
  // EMCID_SOLVE_LOOKAHEAD=0: the whole trailing update on the caller's stream, in front of the next diagonal block
  static const bool v = [] { const char* e = getenv("EMCID_SOLVE_LOOKAHEAD"); return !(e && e[0] == '0'); }();
  return v;
}

inline int factor_spd(const FactorCtx& f, int* status_dev, cudaStream_t stream) {
  const int B = f.B, d = f.dim, sms = f.sms;
  const int nblk = d / SOLVE_NB;
  const long long dd = static_cast<long long>(d) * d;
  int rc;
  // Blocked right-looking Cholesky of M32 with a one-step look-ahead: the diagonal-block kernel keeps B SMs busy for
  // ~80 us and sat in front of every trailing update (24 times per d = 3072).  The trailing update of step k is cut in
  // two: block column k+1 (all the next diagonal block and panel need) stays on the caller's stream, the rest runs on a
  // side stream next to potrf(k+1) and panel(k+1) and is joined before block column k+2 is touched again.  The side GEMM
  // leaves B SMs free for the diagonal-block CTAs (both kernels take a whole SM's shared memory).
  //
  // The explicit inverse X = L^-1 follows ONE BLOCK ROW BEHIND on a second side stream: block row k of L is final once
  // panel(k-1) has run and X_kk comes out of potrf(k), so
  //        T^T        = X[0:k, 0:k]^T  L[k, 0:k]^T          (A = X^T planes, upper triangular: k_tri = 4)
  //        X[k, 0:k]^T = -T^T X_kk^T                         (planes into LvT, transposed planes into Lv)
  // are issued right behind potrf(k) and run in the shadow of the chain potrf -> panel -> block column, which leaves most
  // of the device idle (a single d = 3072 factorisation: 126 us per step on the critical path, 3.0 ms in all).  Computed
  // block column by block column AFTER the factorisation, as it was, the inverse was another 46 dependent launches
  // (1.6 ms of a 7.0 ms single-layer solve, profiles/round2/r05a_solve1_launch_list.txt); now only its last row is exposed.
  SolveSide* side = nullptr;
  const bool lookahead = solve_use_lookahead() && nblk > 2;
  if (nblk > 1) {
    int device = 0;
    EMCID_CUDA_CHECK(cudaGetDevice(&device));
    if ((rc = solve_side(device, &side))) return rc;
  }
  auto inverse_row = [&](int k) -> int {   // on side->inv, after potrf(k) (and with it panel(k-1)) on the caller's stream
    EMCID_CUDA_CHECK(cudaEventRecord(side->inv_fork, stream));
    EMCID_CUDA_CHECK(cudaStreamWaitEvent(side->inv, side->inv_fork, 0));
    const int r0 = k * SOLVE_NB;
    SubGemm g;
    memset(&g, 0, sizeof(g));
    g.A = &f.mLvT; g.a_batch_rows = d;
    g.B = &f.mLL; g.b_row0 = r0; g.b_batch_rows = d;
    g.M = r0; g.N = SOLVE_NB; g.K = r0; g.k_tri = 4;
    g.alpha = 1.0f;
    g.P_hi = f.Tp_hi; g.P_lo = f.Tp_lo; g.ldp = SOLVE_NB; g.p_batch = static_cast<long long>(d) * SOLVE_NB;
    if (int r = run_subgemm(g, B, sms, side->inv)) return r;
    memset(&g, 0, sizeof(g));
    g.A = &f.mTp; g.a_batch_rows = d;
    g.B = &f.mLi; g.b_row0 = r0; g.b_batch_rows = d;
    g.M = r0; g.N = SOLVE_NB; g.K = SOLVE_NB;
    g.alpha = -1.0f;
    g.P_hi = f.LvT_hi + r0; g.P_lo = f.LvT_lo + r0; g.ldp = d; g.p_batch = dd;
    const long long off = static_cast<long long>(r0) * d;
    g.Pt_hi = f.Lv_hi + off; g.Pt_lo = f.Lv_lo + off; g.ldpt = d; g.pt_batch = dd;
    return run_subgemm(g, B, sms, side->inv);
  };
  bool pending_join = false;
  for (int k = 0; k < nblk; ++k) {
    potrf_diag_kernel_v2<<<B, 256, potrf_smem_bytes(), stream>>>(f.M32, d, k, f.LL_hi, f.LL_lo, f.Li_hi, f.Li_lo, f.LiT_hi,
                                                                 f.LiT_lo, f.Lv_hi, f.Lv_lo, f.LvT_hi, f.LvT_lo, status_dev);
    EMCID_CUDA_CHECK(cudaGetLastError());
    if (k > 0 && (rc = inverse_row(k))) return rc;
    const int rem = d - (k + 1) * SOLVE_NB;
    if (rem <= 0) break;
    const long long off_panel = static_cast<long long>(k + 1) * SOLVE_NB * d + static_cast<long long>(k) * SOLVE_NB;
    const long long off_panel_t = static_cast<long long>(k) * SOLVE_NB * d + static_cast<long long>(k + 1) * SOLVE_NB;
    SubGemm g;
    memset(&g, 0, sizeof(g));  // panel: L_ik = M_ik Linv_kk^T
    g.A = &f.mMp; g.a_row0 = (k + 1) * SOLVE_NB; g.a_col0 = k * SOLVE_NB; g.a_batch_rows = d;
    g.B = &f.mLi; g.b_row0 = k * SOLVE_NB; g.b_col0 = 0; g.b_batch_rows = d;
    g.M = rem; g.N = SOLVE_NB; g.K = SOLVE_NB;
    g.alpha = 1.0f; g.beta = 0.0f;
    g.P_hi = f.LL_hi + off_panel; g.P_lo = f.LL_lo + off_panel; g.ldp = d; g.p_batch = dd;
    g.Pt_hi = f.LL_hi + off_panel_t; g.Pt_lo = f.LL_lo + off_panel_t; g.ldpt = d; g.pt_batch = dd;
    if ((rc = run_subgemm(g, B, sms, stream))) return rc;
    if (pending_join) {   // the rest of step k-1's trailing update wrote the tiles the next products read and write
      EMCID_CUDA_CHECK(cudaStreamWaitEvent(stream, side->join, 0));
      pending_join = false;
    }
    // trailing: M_ij -= L_ik L_jk^T  for i >= j > k;  cols = how many block columns from k+1 on, 0 = all (lower tiles)
    auto trailing = [&](int first_blk, int cols, int sm_budget, cudaStream_t st) -> int {
      const int r0 = first_blk * SOLVE_NB;
      const long long off = static_cast<long long>(r0) * d + r0;
      SubGemm t;
      memset(&t, 0, sizeof(t));
      t.A = &f.mLL; t.a_row0 = r0; t.a_col0 = k * SOLVE_NB; t.a_batch_rows = d;
      t.B = &f.mLL; t.b_row0 = r0; t.b_col0 = k * SOLVE_NB; t.b_batch_rows = d;
      t.M = d - r0; t.N = cols ? cols * SOLVE_NB : d - r0; t.K = SOLVE_NB; t.lower = cols ? 0 : 1;
      t.alpha = -1.0f; t.beta = 1.0f;
      t.Cin = f.M32 + off; t.ldcin = d; t.cin_batch = dd;
      t.C = f.M32 + off; t.ldc = d; t.c_batch = dd;
      t.P_hi = f.Mp_hi + off; t.P_lo = f.Mp_lo + off; t.ldp = d; t.p_batch = dd;
      return run_subgemm(t, B, sm_budget, st);
    };
    if (!lookahead || rem <= SOLVE_NB) {
      if ((rc = trailing(k + 1, 0, sms, stream))) return rc;
    } else {
      if ((rc = trailing(k + 1, 1, sms, stream))) return rc;                 // block column k+1
      EMCID_CUDA_CHECK(cudaEventRecord(side->fork, stream));
      EMCID_CUDA_CHECK(cudaStreamWaitEvent(side->stream, side->fork, 0));
      if ((rc = trailing(k + 2, 0, sms - B, side->stream))) return rc;       // the rest, next to potrf(k+1) / panel(k+1)
      EMCID_CUDA_CHECK(cudaEventRecord(side->join, side->stream));
      pending_join = true;
    }
  }
  if (pending_join) EMCID_CUDA_CHECK(cudaStreamWaitEvent(stream, side->join, 0));
  if (nblk > 1) {
    EMCID_CUDA_CHECK(cudaEventRecord(side->inv_join, side->inv));
    EMCID_CUDA_CHECK(cudaStreamWaitEvent(stream, side->inv_join, 0));
  }
  return EMCID_OK;
}

// ---- refined application of a factorisation ---------------------------------------------------------------
// Solves  X M = R  for B stacked problems with the right-hand sides held TRANSPOSED ([rows_pad x dim], dim contiguous;
// rows = number of right-hand sides, pad rows zero):   X^T = Linv^T (Linv R^T)  in fp32-class arithmetic, then
// fp64-residual refinement against M64.  On entry W / Wp hold float(R) and its planes unless `rhs_from64`.
struct ApplyCtx {
  int rows, rows_pad;
  const double* M64;     // [B][dim x dim] fp64, full symmetric
  const double* R64t;    // [B][rows_pad x dim]
  float *W, *Wp_hi, *Wp_lo, *W2p_hi, *W2p_lo;   // [B][rows_pad x dim]
  double* X64t;          // [B][rows_pad x dim] result
  double* norms;         // [2]
  double* P64;           // [SOLVE_SPLIT_ELEMS] split-K slices of the residual product (or null: no split-K)
  PlaneMaps mW, mW2;
};

inline int apply_make_maps(ApplyCtx& a, int B, int dim) {
  int rc;
  if ((rc = make_plane_maps(&a.mW, a.Wp_hi, a.Wp_lo, static_cast<long long>(B) * a.rows_pad, dim, dim))) return rc;
  return make_plane_maps(&a.mW2, a.W2p_hi, a.W2p_lo, static_cast<long long>(B) * a.rows_pad, dim, dim);
}

inline int refined_solve(const FactorCtx& f, const ApplyCtx& a, bool rhs_from64, int refine_steps, cudaStream_t stream,
                         double adapt_tol = SOLVE_ADAPT_TOL, int* status_dev = nullptr) {
  const int B = f.B, d = f.dim, sms = f.sms, n = a.rows, n_pad = a.rows_pad;
  const long long dd = static_cast<long long>(d) * d, nd = static_cast<long long>(n_pad) * d;
  const long long tot = static_cast<long long>(B) * nd;
  int rc;
  //   Y^T = W Linv^T  (B = Linv, lower triangular: k_tri = 1)  ->  planes only
  //   X^T = Y^T Linv  (B = Linv^T, upper triangular: k_tri = 2) ->  fp32 into W
  auto apply_inverse = [&]() -> int {
    SubGemm g;
    memset(&g, 0, sizeof(g));
    g.A = &a.mW; g.a_batch_rows = n_pad;
    g.B = &f.mLv; g.b_batch_rows = d;
    g.M = n_pad; g.N = d; g.K = d; g.k_tri = 1; g.alpha = 1.0f;
    g.P_hi = a.W2p_hi; g.P_lo = a.W2p_lo; g.ldp = d; g.p_batch = nd;
    if (int r = run_subgemm(g, B, sms, stream)) return r;
    memset(&g, 0, sizeof(g));
    g.A = &a.mW2; g.a_batch_rows = n_pad;
    g.B = &f.mLvT; g.b_batch_rows = d;
    g.M = n_pad; g.N = d; g.K = d; g.k_tri = 2; g.alpha = 1.0f;
    g.C = a.W; g.ldc = d; g.c_batch = nd;
    return run_subgemm(g, B, sms, stream);
  };
  if (rhs_from64) {
    rhs_from64_kernel<<<sms * 8, 256, 0, stream>>>(a.R64t, tot, a.W, a.Wp_hi, a.Wp_lo);
    EMCID_CUDA_CHECK(cudaGetLastError());
  }
  // refine_steps >= 0: exactly that many sweeps.  -1: adaptive — stop once the error predicted from the contraction of
  // the corrections is below the target (see SOLVE_ADAPT_TOL).
  const bool adaptive = refine_steps < 0;
  const double tol = solve_adapt_tol(adapt_tol);
  double c_prev = 1.0, rho_max = 0.0;
  bool converged = !adaptive;
  const int max_steps = adaptive ? SOLVE_ADAPT_MAX : refine_steps;
  for (int it = 0; it <= max_steps; ++it) {
    if (it > 0) {
      // R^T = R0^T - X^T M64  (fp64), rounded to fp32 into W, then re-split
      DgemmParams p;
      memset(&p, 0, sizeof(p));
      p.M = n; p.N = d; p.K = d;
      p.A = a.X64t; p.lda = d; p.a_batch = nd;
      p.B = a.M64; p.ldb = d; p.b_batch = dd;
      p.alpha = -1.0; p.beta = 1.0;
      int split = a.P64 ? dgemm_pick_split(B * (n_pad / DG_BM) * ((d + DG_BN - 1) / DG_BN), d, sms) : 1;
      while (split > 1 && split * tot > SOLVE_SPLIT_ELEMS) --split;
      if (split > 1) {
        // few tiles, long K (narrow edit): K sliced over the idle SMs, slices summed with R0 by the reducer
        p.M = n_pad;   // pad rows of X are zero: every slice element is defined
        p.C = a.P64; p.ldc = d; p.c_batch = nd;
        p.split_k = split; p.c_split = tot;
        if ((rc = launch_dgemm_nt(p, B, stream))) return rc;
        residual_reduce_kernel<<<sms * 8, 256, 0, stream>>>(a.R64t, a.P64, split, tot, tot, a.W, a.Wp_hi, a.Wp_lo);
        EMCID_CUDA_CHECK(cudaGetLastError());
      } else {
        p.Cin = a.R64t; p.ldcin = d; p.cin_batch = nd;
        p.C32 = a.W; p.ldc32 = d; p.c32_batch = nd;
        if ((rc = launch_dgemm_nt(p, B, stream))) return rc;
        if ((rc = launch_split_planes(a.W, d, B * n_pad, d, 1.0f, a.Wp_hi, a.Wp_lo, d, stream))) return rc;
      }
    }
    if ((rc = apply_inverse())) return rc;
    solve_axpy_kernel<<<sms * 8, 256, 0, stream>>>(a.W, a.X64t, tot, it > 0 ? 1 : 0);
    EMCID_CUDA_CHECK(cudaGetLastError());
    if (adaptive && it > 0) {
      double hn[2] = {0.0, 0.0};
      EMCID_CUDA_CHECK(cudaMemsetAsync(a.norms, 0, 2 * sizeof(double), stream));
      solve_norms_kernel<<<sms * 4, 256, 0, stream>>>(a.W, a.X64t, tot, a.norms);
      EMCID_CUDA_CHECK(cudaGetLastError());
      EMCID_CUDA_CHECK(cudaMemcpyAsync(hn, a.norms, 2 * sizeof(double), cudaMemcpyDeviceToHost, stream));
      EMCID_CUDA_CHECK(cudaStreamSynchronize(stream));
      // EMCID_SOLVE_DEBUG=1: the relative size of every sweep's correction (dimension, right-hand sides, sweep, ratio)
      static const bool debug = [] { const char* e = getenv("EMCID_SOLVE_DEBUG"); return e && e[0] == '1'; }();
      if (debug) fprintf(stderr, "emcid refine dim=%d rhs=%d sweep=%d |dx|/|x|=%.3e\n", d, n, it, sqrt(hn[0] / hn[1]));
      if (!(hn[1] > 0.0)) { converged = true; break; }
      const double c = sqrt(hn[0] / hn[1]);
      if (c_prev > 0.0 && c / c_prev > rho_max) rho_max = c / c_prev;
      c_prev = c;
      const double rho = fmin(0.9, 1.5 * rho_max);
      if (rho * c / (1.0 - rho) <= tol) { converged = true; break; }
    }
  }
  if (!converged && status_dev) {
    // SOLVE_ADAPT_MAX sweeps did not bring the predicted error under the target (a factor that barely contracts: the
    // matrix is too ill-conditioned for fp32-class factors): the result is the best available, flagged in bit 1
    status_or_kernel<<<1, 1, 0, stream>>>(status_dev, 2);
    EMCID_CUDA_CHECK(cudaGetLastError());
  }
  return EMCID_OK;
}

// resid and dW = float(resid adj_k^T) from the finished adj_k (emcid_main.py:1048-1050)
inline int solve_outputs(int B, int d, int h, int n, const float* St, long long lds, double scale,
                         const double* inv_left_dev, const double* adj_k, double* resid, float* dW, cudaStream_t stream) {
  resid_kernel<<<dim3((h + 31) / 32, (n + 31) / 32, B), dim3(32, 8), 0, stream>>>(
      St, lds, static_cast<long long>(n) * lds, n, h, scale, inv_left_dev, resid);
  EMCID_CUDA_CHECK(cudaGetLastError());
  DgemmParams p;
  memset(&p, 0, sizeof(p));
  p.M = h; p.N = d; p.K = n;
  p.A = resid; p.lda = n; p.a_batch = static_cast<long long>(h) * n;
  p.B = adj_k; p.ldb = n; p.b_batch = static_cast<long long>(d) * n;
  p.alpha = 1.0;
  p.C32 = dW; p.ldc32 = d; p.c32_batch = static_cast<long long>(h) * d;
  return launch_dgemm_nt(p, B, stream);
}

// ---- workspace of the direct solve ----------------------------------------------------------------------
struct SolveWs {
  FactorCtx f;
  ApplyCtx a;
  float *Kd_hi, *Kd_lo;
  double *M64, *Ks64t, *Kd64, *inv_left;
  size_t bytes;
};

struct Carver {
  uint8_t *p, *p0;
  explicit Carver(void* base)
      : p(reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(base) + 1023) & ~static_cast<uintptr_t>(1023))), p0(p) {}
  template <typename T>
  T* take(size_t count) {
    uint8_t* r = p;
    p += (count * sizeof(T) + 1023) & ~static_cast<size_t>(1023);
    return reinterpret_cast<T*>(r);
  }
  size_t bytes() const { return static_cast<size_t>(p - p0) + 1024; }
};

inline void carve_factor(Carver& c, FactorCtx& f, int B, int dim, bool with_inverse) {
  const size_t dd = static_cast<size_t>(B) * dim * dim, d128 = static_cast<size_t>(B) * dim * SOLVE_NB;
  f.M32 = c.take<float>(dd);
  f.Mp_hi = c.take<float>(dd); f.Mp_lo = c.take<float>(dd);
  f.LL_hi = c.take<float>(dd); f.LL_lo = c.take<float>(dd);
  f.Li_hi = c.take<float>(d128); f.Li_lo = c.take<float>(d128);
  f.LiT_hi = c.take<float>(d128); f.LiT_lo = c.take<float>(d128);
  f.Tp_hi = c.take<float>(d128); f.Tp_lo = c.take<float>(d128);
  if (with_inverse) {
    f.Lv_hi = c.take<float>(dd); f.Lv_lo = c.take<float>(dd);
    f.LvT_hi = c.take<float>(dd); f.LvT_lo = c.take<float>(dd);
  }
}

inline void carve_apply(Carver& c, ApplyCtx& a, size_t elems) {
  a.W = c.take<float>(elems);
  a.Wp_hi = c.take<float>(elems); a.Wp_lo = c.take<float>(elems);
  a.W2p_hi = c.take<float>(elems); a.W2p_lo = c.take<float>(elems);
}

inline SolveWs solve_carve(void* base, int B, int d, int h, int n) {
  (void)h;
  const long long n_pad = round_up_ll(n, 128);
  Carver c(base);
  SolveWs w;
  memset(&w, 0, sizeof(w));
  const size_t dd = static_cast<size_t>(B) * d * d, nd = static_cast<size_t>(B) * n_pad * d;
  w.M64 = c.take<double>(dd);
  w.Ks64t = c.take<double>(nd);
  w.a.X64t = c.take<double>(nd);
  w.Kd64 = c.take<double>(nd);
  w.inv_left = c.take<double>(B);
  w.a.norms = c.take<double>(2);
  w.a.P64 = c.take<double>(SOLVE_SPLIT_ELEMS);
  carve_factor(c, w.f, B, d, true);
  carve_apply(c, w.a, nd);
  w.Kd_hi = c.take<float>(nd);
  w.Kd_lo = c.take<float>(nd);
  w.bytes = c.bytes();
  return w;
}

inline size_t solve_workspace_bytes(int B, int d, int h, int n) { return solve_carve(nullptr, B, d, h, n).bytes; }

// ---- the direct solve ------------------------------------------------------------------------------------
inline int solve_layers(int device, int B, int d, int h, int n, const float* C32, const float* Kt, long long ldk,
                        const float* St, long long lds, double lambda, double scale, const double* inv_left_host,
                        double* adj_k, double* resid, float* dW, int refine_steps, void* workspace, size_t ws_bytes,
                        int* status_dev, cudaStream_t stream) {
  EMCID_CHECK(B > 0 && d > 0 && h > 0 && n > 0, EMCID_ERR_INVALID, "solve: empty problem");
  EMCID_CHECK(d % SOLVE_NB == 0, EMCID_ERR_UNSUPPORTED, "solve: d must be a multiple of %d (got %d)", SOLVE_NB, d);
  EMCID_CHECK(C32 && Kt && St && adj_k && resid && dW && status_dev && inv_left_host, EMCID_ERR_INVALID,
              "solve: null argument");
  EMCID_CHECK(ldk >= d && lds >= h, EMCID_ERR_INVALID, "solve: bad leading dimensions");
  EMCID_CHECK(refine_steps >= -1 && refine_steps <= 8, EMCID_ERR_INVALID, "solve: refine_steps out of range");
  EMCID_CHECK(ws_bytes >= solve_workspace_bytes(B, d, h, n), EMCID_ERR_WORKSPACE,
              "solve: workspace too small (%zu < %zu)", ws_bytes, solve_workspace_bytes(B, d, h, n));
  EMCID_CUDA_CHECK(cudaSetDevice(device));
  DeviceInfo info;
  int rc = get_device_info(&info);
  if (rc) return rc;
  const int sms = info.sm_count;
  const int n_pad = static_cast<int>(round_up_ll(n, 128));
  SolveWs w = solve_carve(workspace, B, d, h, n);
  const long long dd = static_cast<long long>(d) * d;
  w.f.B = B; w.f.dim = d; w.f.sms = sms;
  w.a.rows = n; w.a.rows_pad = n_pad; w.a.M64 = w.M64; w.a.R64t = w.Ks64t;
  if ((rc = potrf_configure(device))) return rc;

  EMCID_CUDA_CHECK(cudaMemsetAsync(status_dev, 0, sizeof(int), stream));
  EMCID_CUDA_CHECK(cudaMemcpyAsync(w.inv_left, inv_left_host, B * sizeof(double), cudaMemcpyHostToDevice, stream));
  if ((rc = factor_clear(w.f, stream))) return rc;
  EMCID_CUDA_CHECK(cudaMemsetAsync(w.f.Mp_hi, 0, B * dd * sizeof(float), stream));
  EMCID_CUDA_CHECK(cudaMemsetAsync(w.f.Mp_lo, 0, B * dd * sizeof(float), stream));

  // 1. operands
  {
    dim3 grid((d + 31) / 32, (n_pad + 31) / 32, B);
    solve_prep_kernel<<<grid, dim3(32, 8), 0, stream>>>(Kt, ldk, static_cast<long long>(n) * ldk, n, n_pad, d, scale,
                                                        w.Ks64t, w.a.W, w.a.Wp_hi, w.a.Wp_lo, w.Kd64, w.Kd_hi, w.Kd_lo);
    EMCID_CUDA_CHECK(cudaGetLastError());
  }
  PlaneMaps mKd;
  if ((rc = make_plane_maps(&mKd, w.Kd_hi, w.Kd_lo, static_cast<long long>(B) * d, n_pad, n_pad)) ||
      (rc = factor_make_maps(w.f)) || (rc = apply_make_maps(w.a, B, d)))
    return rc;

  // 2. M32 = lambda*C32 + Ks Ks^T (lower tiles; fp32 + planes)   and   M64 (fp64, full)
  {
    SubGemm g;
    memset(&g, 0, sizeof(g));
    g.A = &mKd; g.a_batch_rows = d; g.B = &mKd; g.b_batch_rows = d;
    g.M = d; g.N = d; g.K = n_pad; g.lower = 1;
    g.alpha = 1.0f; g.beta = static_cast<float>(lambda);
    g.Cin = C32; g.ldcin = d; g.cin_batch = dd;
    g.C = w.f.M32; g.ldc = d; g.c_batch = dd;
    g.P_hi = w.f.Mp_hi; g.P_lo = w.f.Mp_lo; g.ldp = d; g.p_batch = dd;
    if ((rc = run_subgemm(g, B, sms, stream))) return rc;
  }
  // M64 (fp64, full) is only read by the refinement: with one or two stacked problems it is formed on a side stream
  // BESIDE the factorisation (solve_m64_ctas).  (A plain launch there floods every SM with long-running CTAs in front of
  // the Cholesky chain's short, dependent launches; in front of the factorisation the product sits on the critical path:
  // 0.6 ms of a 7.0 ms single-layer solve.)
  SolveSide* side = nullptr;
  if ((rc = solve_side(device, &side))) return rc;
  {
    const int m64_ctas = solve_m64_ctas(sms, B);
    cudaStream_t st64 = m64_ctas > 0 ? side->aux : stream;
    if (m64_ctas > 0) {
      EMCID_CUDA_CHECK(cudaEventRecord(side->aux_fork, stream));
      EMCID_CUDA_CHECK(cudaStreamWaitEvent(side->aux, side->aux_fork, 0));
    }
    DgemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = d; p.N = d; p.K = n;
    p.A = w.Kd64; p.lda = n_pad; p.a_batch = static_cast<long long>(d) * n_pad;
    p.B = w.Kd64; p.ldb = n_pad; p.b_batch = static_cast<long long>(d) * n_pad;
    p.alpha = 1.0; p.beta = lambda;
    p.Cin32 = C32; p.ldcin32 = d; p.cin32_batch = dd;
    p.C = w.M64; p.ldc = d; p.c_batch = dd;
    p.lower = 1;
    if (m64_ctas > 0) p.persist = m64_ctas;
    if ((rc = launch_dgemm_nt(p, B, st64))) return rc;
    mirror64_kernel<<<dim3(d / 32, d / 32, B), dim3(32, 8), 0, st64>>>(w.M64, d);
    EMCID_CUDA_CHECK(cudaGetLastError());
    EMCID_CUDA_CHECK(cudaEventRecord(side->aux_join, st64));
  }

  // 3. blocked right-looking Cholesky of M32 and the explicit inverse of its factor
  if ((rc = factor_spd(w.f, status_dev, stream))) return rc;

  EMCID_CUDA_CHECK(cudaStreamWaitEvent(stream, side->aux_join, 0));

  // 4./5. solve + refinement; W holds the current right-hand side (transposed), then the solution
  if ((rc = refined_solve(w.f, w.a, false, refine_steps, stream, SOLVE_ADAPT_TOL, status_dev))) return rc;

  // 6. outputs
  transpose_out_kernel<<<dim3((d + 31) / 32, (n + 31) / 32, B), dim3(32, 8), 0, stream>>>(w.a.X64t, n, n_pad, d, adj_k);
  EMCID_CUDA_CHECK(cudaGetLastError());
  return solve_outputs(B, d, h, n, St, lds, scale, w.inv_left, adj_k, resid, dW, stream);
}

// ---- cached factorisation for repeated edits with the same covariance --------------------------
// Sequential editing (experiments/sequential_editing.py:98-171), the debias factor search
// (emcid_main.py:1460-1472 -> cal_insert_deltas :1969-2052) and the layer ablation re-solve
// (lambda*C + Ks Ks^T) adj_k = Ks many times with the same lambda*C =: A and a fresh, much narrower Ks.  With
// A = L L^T factored once per (layer, lambda, edit_weight), the push-through identity
//        (A + Ks Ks^T)^-1 Ks = A^-1 Ks (I + Ks^T A^-1 Ks)^-1
// turns each edit into O(d^2 n) work instead of O(d^3):
//        Y     = A^-1 Ks              explicit-inverse application + fp64 refinement against A64   [d x n]
//        G     = I + Ks^T Y           fp64 DMMA GEMM, n_pad x n_pad, SPD, cond(G) <= 1 + |Ks^T A^-1 Ks|
//        adj_k = Y G^-1               the same blocked factorisation + refined application, dimension n_pad, d rhs
// Unlike the additive Woodbury form  A^-1 - A^-1 Ks G^-1 Ks^T A^-1  nothing is subtracted, so fp32-class factors with
// fp64-residual refinement of BOTH solves reach the accuracy of the direct path (each factor only has to contract).
struct FactorHandle {
  int device, d;
  double lambda;
  double* A64;                                  // [d x d] fp64 = lambda * double(C32), full
  float *Lv_hi, *Lv_lo, *LvT_hi, *LvT_lo;       // explicit inverse planes of chol(lambda * C32)
};

inline int factor_destroy(FactorHandle* H) {
  if (!H) return EMCID_OK;
  cudaSetDevice(H->device);
  dev_free(H->A64); dev_free(H->Lv_hi); dev_free(H->Lv_lo); dev_free(H->LvT_hi); dev_free(H->LvT_lo);
  delete H;
  return EMCID_OK;
}

inline int factor_create(FactorHandle** out, int device, int d, const float* C32, double lambda, int* status_dev,
                         cudaStream_t stream) {
  EMCID_CHECK(out && C32 && status_dev, EMCID_ERR_INVALID, "factor_create: null argument");
  EMCID_CHECK(d > 0 && d % SOLVE_NB == 0, EMCID_ERR_UNSUPPORTED, "factor_create: d must be a multiple of %d (got %d)",
              SOLVE_NB, d);
  EMCID_CHECK(lambda > 0.0, EMCID_ERR_INVALID, "factor_create: lambda must be positive");
  EMCID_CUDA_CHECK(cudaSetDevice(device));
  DeviceInfo info;
  int rc = get_device_info(&info);
  if (rc) return rc;
  if ((rc = potrf_configure(device))) return rc;
  const size_t dd = static_cast<size_t>(d) * d;
  FactorHandle* H = new FactorHandle();
  memset(H, 0, sizeof(*H));
  H->device = device; H->d = d; H->lambda = lambda;
  void* scratch = nullptr;
  auto fail = [&](int code) { if (scratch) dev_free(scratch); factor_destroy(H); return code; };
#define EMCID_FACTOR_ALLOC(ptr, bytes)                                                          \
  do {                                                                                          \
    cudaError_t e_ = dev_alloc(reinterpret_cast<void**>(&(ptr)), (bytes));                       \
    if (e_ != cudaSuccess) return fail(set_error(EMCID_ERR_CUDA, "factor_create: %s", cudaGetErrorString(e_))); \
  } while (0)
  EMCID_FACTOR_ALLOC(H->A64, dd * sizeof(double));
  EMCID_FACTOR_ALLOC(H->Lv_hi, dd * sizeof(float));
  EMCID_FACTOR_ALLOC(H->Lv_lo, dd * sizeof(float));
  EMCID_FACTOR_ALLOC(H->LvT_hi, dd * sizeof(float));
  EMCID_FACTOR_ALLOC(H->LvT_lo, dd * sizeof(float));
  FactorCtx f;
  memset(&f, 0, sizeof(f));
  size_t scratch_bytes;
  {
    Carver c(nullptr);
    carve_factor(c, f, 1, d, false);
    scratch_bytes = c.bytes();
  }
  EMCID_FACTOR_ALLOC(scratch, scratch_bytes);
#undef EMCID_FACTOR_ALLOC
  {
    Carver c(scratch);
    carve_factor(c, f, 1, d, false);
  }
  f.B = 1; f.dim = d; f.sms = info.sm_count;
  f.Lv_hi = H->Lv_hi; f.Lv_lo = H->Lv_lo; f.LvT_hi = H->LvT_hi; f.LvT_lo = H->LvT_lo;
  if ((rc = factor_make_maps(f))) return fail(rc);
  auto run = [&]() -> int {
    EMCID_CUDA_CHECK(cudaMemsetAsync(status_dev, 0, sizeof(int), stream));
    if (int r = factor_clear(f, stream)) return r;
    factor_prep_kernel<<<dim3(d / 32, d / 32), dim3(32, 8), 0, stream>>>(C32, d, lambda, f.M32, f.Mp_hi, f.Mp_lo, H->A64);
    EMCID_CUDA_CHECK(cudaGetLastError());
    if (int r = factor_spd(f, status_dev, stream)) return r;
    EMCID_CUDA_CHECK(cudaStreamSynchronize(stream));   // the scratch goes back to the pool: nothing may still use it
    return EMCID_OK;
  };
  if ((rc = run())) return fail(rc);
  dev_free(scratch);
  *out = H;
  return EMCID_OK;
}

struct FactorSolveWs {
  FactorCtx g;          // factorisation of G, dimension n_pad
  ApplyCtx a1, a2;      // stage 1: n rhs of dimension d;  stage 2: d rhs of dimension n_pad (same buffers)
  double *Ks64t, *Yd64, *G64, *inv_left;
  size_t bytes;
};

inline FactorSolveWs factor_solve_carve(void* base, int d, int n) {
  const long long n_pad = round_up_ll(n, 128);
  Carver c(base);
  FactorSolveWs w;
  memset(&w, 0, sizeof(w));
  const size_t nd = static_cast<size_t>(n_pad) * d;
  w.Ks64t = c.take<double>(nd);
  w.a1.X64t = c.take<double>(nd);
  w.Yd64 = c.take<double>(nd);
  w.G64 = c.take<double>(static_cast<size_t>(n_pad) * n_pad);
  w.inv_left = c.take<double>(1);
  w.a1.norms = c.take<double>(2);
  w.a1.P64 = c.take<double>(SOLVE_SPLIT_ELEMS);
  carve_factor(c, w.g, 1, static_cast<int>(n_pad), true);
  carve_apply(c, w.a1, nd);
  // stage 2 runs after stage 1 is finished on the same stream: its fp32 buffers and its result reuse stage 1's
  w.a2 = w.a1;
  w.a2.X64t = w.Ks64t;
  w.bytes = c.bytes();
  return w;
}

inline size_t factor_solve_workspace_bytes(int d, int n) { return factor_solve_carve(nullptr, d, n).bytes; }

inline int factor_solve(FactorHandle* H, int h, int n, const float* Kt, long long ldk, const float* St, long long lds,
                        double scale, double inv_layers_left, double* adj_k, double* resid, float* dW, int refine_steps,
                        void* workspace, size_t ws_bytes, int* status_dev, cudaStream_t stream) {
  EMCID_CHECK(H, EMCID_ERR_INVALID, "factor_solve: null handle");
  const int d = H->d;
  EMCID_CHECK(h > 0 && n > 0, EMCID_ERR_INVALID, "factor_solve: empty problem");
  EMCID_CHECK(Kt && St && adj_k && resid && dW && status_dev, EMCID_ERR_INVALID, "factor_solve: null argument");
  EMCID_CHECK(ldk >= d && lds >= h, EMCID_ERR_INVALID, "factor_solve: bad leading dimensions");
  EMCID_CHECK(refine_steps >= -1 && refine_steps <= 8, EMCID_ERR_INVALID, "factor_solve: refine_steps out of range");
  EMCID_CHECK(ws_bytes >= factor_solve_workspace_bytes(d, n), EMCID_ERR_WORKSPACE,
              "factor_solve: workspace too small (%zu < %zu)", ws_bytes, factor_solve_workspace_bytes(d, n));
  EMCID_CUDA_CHECK(cudaSetDevice(H->device));
  DeviceInfo info;
  int rc = get_device_info(&info);
  if (rc) return rc;
  const int sms = info.sm_count;
  const int n_pad = static_cast<int>(round_up_ll(n, 128));
  FactorSolveWs w = factor_solve_carve(workspace, d, n);
  if ((rc = potrf_configure(H->device))) return rc;

  // the cached factor of A
  FactorCtx fa;
  memset(&fa, 0, sizeof(fa));
  fa.B = 1; fa.dim = d; fa.sms = sms;
  fa.Lv_hi = H->Lv_hi; fa.Lv_lo = H->Lv_lo; fa.LvT_hi = H->LvT_hi; fa.LvT_lo = H->LvT_lo;
  if ((rc = factor_make_inverse_maps(fa))) return rc;
  w.a1.rows = n; w.a1.rows_pad = n_pad; w.a1.M64 = H->A64; w.a1.R64t = w.Ks64t;
  if ((rc = apply_make_maps(w.a1, 1, d))) return rc;

  EMCID_CUDA_CHECK(cudaMemsetAsync(status_dev, 0, sizeof(int), stream));
  EMCID_CUDA_CHECK(cudaMemcpyAsync(w.inv_left, &inv_layers_left, sizeof(double), cudaMemcpyHostToDevice, stream));
  // 1. Ks (fp64, transposed) and the planes of float(Ks)
  {
    dim3 grid((d + 31) / 32, (n_pad + 31) / 32, 1);
    solve_prep_kernel<<<grid, dim3(32, 8), 0, stream>>>(Kt, ldk, static_cast<long long>(n) * ldk, n, n_pad, d, scale,
                                                        w.Ks64t, w.a1.W, w.a1.Wp_hi, w.a1.Wp_lo, nullptr, nullptr, nullptr);
    EMCID_CUDA_CHECK(cudaGetLastError());
  }
  // 2. Y^T = Ks^T A^-1, refined against A64                                  -> a1.X64t [n_pad x d]
  if ((rc = refined_solve(fa, w.a1, false, refine_steps, stream, SOLVE_ADAPT_TOL_CHAINED, status_dev))) return rc;
  // 3. G = I + Ks^T Y   (fp64; lower tiles, then mirrored with fp32 copy and planes)
  {
    DgemmParams p;
    memset(&p, 0, sizeof(p));
    p.M = n_pad; p.N = n_pad; p.K = d;
    p.A = w.Ks64t; p.lda = d;
    p.B = w.a1.X64t; p.ldb = d;
    p.alpha = 1.0;
    p.C = w.G64; p.ldc = n_pad;
    p.lower = 1;
    int g_tiles = 0;   // lower tiles of DG_BM x DG_BN
    for (int bm = 0; bm < n_pad; bm += DG_BM) g_tiles += (bm + DG_BM - 1) / DG_BN + 1;
    const long long gg = static_cast<long long>(n_pad) * n_pad;
    int split = dgemm_pick_split(g_tiles, d, sms);
    while (split > 1 && split * gg > SOLVE_SPLIT_ELEMS) --split;
    if (split > 1) {
      p.C = w.a1.P64; p.split_k = split; p.c_split = gg;
    }
    if ((rc = launch_dgemm_nt(p, 1, stream))) return rc;
    g_prepare_kernel<<<dim3(n_pad / 32, n_pad / 32), dim3(32, 8), 0, stream>>>(w.G64, n_pad, w.a1.P64, split > 1 ? split : 0,
                                                                                gg, w.g.M32, w.g.Mp_hi, w.g.Mp_lo);
    EMCID_CUDA_CHECK(cudaGetLastError());
  }
  // 4. factor G
  w.g.B = 1; w.g.dim = n_pad; w.g.sms = sms;
  if ((rc = factor_make_maps(w.g)) || (rc = factor_clear(w.g, stream)) || (rc = factor_spd(w.g, status_dev, stream)))
    return rc;
  // 5. adj_k G = Y: the rows of Y [d x n_pad] are the right-hand sides             -> a2.X64t [d x n_pad]
  transpose_out_kernel<<<dim3((d + 31) / 32, (n_pad + 31) / 32, 1), dim3(32, 8), 0, stream>>>(w.a1.X64t, n_pad, n_pad, d,
                                                                                             w.Yd64);
  EMCID_CUDA_CHECK(cudaGetLastError());
  w.a2.rows = d; w.a2.rows_pad = d; w.a2.M64 = w.G64; w.a2.R64t = w.Yd64;
  if ((rc = apply_make_maps(w.a2, 1, n_pad))) return rc;
  if ((rc = refined_solve(w.g, w.a2, true, refine_steps, stream, SOLVE_ADAPT_TOL_CHAINED, status_dev))) return rc;
  // 6. outputs
  compact_cols_kernel<<<sms * 4, 256, 0, stream>>>(w.a2.X64t, d, n, n_pad, adj_k);
  EMCID_CUDA_CHECK(cudaGetLastError());
  return solve_outputs(1, d, h, n, St, lds, scale, w.inv_left, adj_k, resid, dW, stream);
}

}  // namespace emcid

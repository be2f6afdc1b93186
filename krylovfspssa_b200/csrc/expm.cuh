// exp(t*H) of the small Hessenberg matrix on ONE CTA: DGPADMnorm / DGPADM
// (src/expokit/dgpadm.f:171-339, :2-169): degree-6 diagonal Pade with scaling and
// squaring.  The reference spends 6+ns DGEMMs and one DGESV here; with m+2 <= 102 two
// operands of a product (2 x 104 x 104 doubles = 173 KB) fit in the 227 KB of shared
// memory of one SM, so every product is staged once into shared memory and each thread
// accumulates a 4x4 register tile; the linear solve (q-p) X = p is an in-shared-memory
// Gaussian elimination with partial pivoting, and the ns squarings never leave the SM.
// The kernel also takes the happy-breakdown decision (which order and which step to
// exponentiate) from the device-resident sweep flag, so no host round trip separates the
// Arnoldi sweep from the exponential.
#pragma once
#include "common.cuh"
#include "krylov.cuh"

namespace kfsp {

constexpr int EXPM_MAXN = 104;                 // m_max + 2 = 102, padded to a multiple of 4
constexpr int EXPM_LDS = EXPM_MAXN;
constexpr int EXPM_THREADS = 1024;
constexpr int EXPM_KL = 6;                     // lower bandwidth of q-p when H is tridiagonal (degree-6 Pade)
constexpr int EXPM_KU = 12;                    // upper bandwidth of U after partial pivoting (kl + ku)
constexpr int EXPM_TD = 3;                     // tile diagonals |tj - ti| <= EXPM_TD are computed in the banded path (2 would do)
constexpr size_t EXPM_SMEM = 2 * (size_t)EXPM_LDS * EXPM_MAXN * sizeof(double) + 64 * sizeof(double) +
                             (size_t)EXPM_MAXN * (EXPM_KU + 1 + EXPM_KL) * sizeof(double) + (size_t)EXPM_MAXN * sizeof(int);

struct ExpmResult {
    int32_t ns;
    int32_t info;               // 0 ok, KFSP_ERR_NULL_H, KFSP_ERR_SINGULAR
    int32_t mx;                 // order actually exponentiated
    int32_t brk;                // copy of the sweep's happy-breakdown column
    double hnorm;               // |t| * ||H||_inf (DGPADMnorm's extra output, dgpadm.f:253)
    double avnorm;              // copy of the sweep's ||A v_{m+1}||
    double t_used;
    double e[EXPM_MAXN];        // first column of exp(t*H)
};

// load an n x n column-major matrix (ld = lda) into shared memory, scaled by alpha, zero padded to np rows/cols
__device__ __forceinline__ void expm_load(double* s, const double* __restrict__ g, int lda, int n, int np, double alpha) {
    for (int t = threadIdx.x; t < np * EXPM_LDS; t += (int)blockDim.x) {
        const int j = t / EXPM_LDS, i = t % EXPM_LDS;
        s[t] = (i < n && j < n) ? __dmul_rn(alpha, g[(size_t)j * lda + i]) : 0.0;
    }
}
// C = sA * sB for the n x n leading blocks; each thread owns a 4x4 tile held in acc.
// ba / bb: bandwidths of the left / right operand (>= n: dense).  Terms outside the bands are exact zeros, so
// leaving them out does not change any accumulated value.
__device__ __forceinline__ void expm_mma(const double* sA, const double* sB, int n, double (&acc)[4][4], int ba = 1 << 20, int bb = 1 << 20,
                                         int ti = -1, int tj = -1) {
    if (ti < 0) { ti = threadIdx.x & 31; tj = threadIdx.x >> 5; }
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    if (4 * ti >= n || 4 * tj >= n) return;
    const double* pa = sA + 4 * ti;
    const double* pb = sB + (size_t)(4 * tj) * EXPM_LDS;
    int k0 = 0, k1 = n - 1;
    if (ba < n || bb < n) {
        k0 = max(0, max(4 * ti - ba, 4 * tj - bb));
        k1 = min(n - 1, min(4 * ti + 3 + ba, 4 * tj + 3 + bb));
    }
    for (int k = k0; k <= k1; ++k) {
        const double2 a01 = *reinterpret_cast<const double2*>(pa + (size_t)k * EXPM_LDS);
        const double2 a23 = *reinterpret_cast<const double2*>(pa + (size_t)k * EXPM_LDS + 2);
        const double b0 = pb[k], b1 = pb[EXPM_LDS + k], b2 = pb[2 * EXPM_LDS + k], b3 = pb[3 * EXPM_LDS + k];
        acc[0][0] = fma(a01.x, b0, acc[0][0]); acc[1][0] = fma(a01.y, b0, acc[1][0]);
        acc[2][0] = fma(a23.x, b0, acc[2][0]); acc[3][0] = fma(a23.y, b0, acc[3][0]);
        acc[0][1] = fma(a01.x, b1, acc[0][1]); acc[1][1] = fma(a01.y, b1, acc[1][1]);
        acc[2][1] = fma(a23.x, b1, acc[2][1]); acc[3][1] = fma(a23.y, b1, acc[3][1]);
        acc[0][2] = fma(a01.x, b2, acc[0][2]); acc[1][2] = fma(a01.y, b2, acc[1][2]);
        acc[2][2] = fma(a23.x, b2, acc[2][2]); acc[3][2] = fma(a23.y, b2, acc[3][2]);
        acc[0][3] = fma(a01.x, b3, acc[0][3]); acc[1][3] = fma(a01.y, b3, acc[1][3]);
        acc[2][3] = fma(a23.x, b3, acc[2][3]); acc[3][3] = fma(a23.y, b3, acc[3][3]);
    }
}
// write the register tiles to a column-major matrix with leading dimension ldc (global or shared)
__device__ __forceinline__ void expm_store(double* C, int ldc, int n, const double (&acc)[4][4], double diag_add, int ti = -1, int tj = -1) {
    if (ti < 0) { ti = threadIdx.x & 31; tj = threadIdx.x >> 5; }
#pragma unroll
    for (int b = 0; b < 4; ++b)
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int i = 4 * ti + a, j = 4 * tj + b;
            if (i < n && j < n) C[(size_t)j * ldc + i] = (i == j) ? __dadd_rn(acc[a][b], diag_add) : acc[a][b];
        }
}

// Everything of DGPADM up to the squarings, on ONE CTA of 1024 threads: ||H||_inf, ns, the Pade numerator and denominator by
// Horner, the linear solve, E = I + 2 X.  Returns info (uniform over the CTA); on success E (n x n, zero padded to a multiple
// of 4) is in sB.  smem: the dynamic shared memory of the kernel (EXPM_SMEM bytes); work: 4 * EXPM_MAXN^2 doubles of global scratch.
#ifdef KFSP_EXPM_TIMING            // developer build: cycle stamps of the phases, printed by thread 0 (tools/expm_timing.py)
#define EXPM_STAMP(i) do { if (threadIdx.x == 0) s_stamp[i] = clock64(); } while (0)
#else
#define EXPM_STAMP(i) do { } while (0)
#endif
__device__ __forceinline__ int expm_pade(const double* __restrict__ H, int ldh, int n, double t, double* work, double* smem, int* ns_out,
                                         double* hnorm_out) {
#ifdef KFSP_EXPM_TIMING
    __shared__ long long s_stamp[12];
#endif
    EXPM_STAMP(0);
    double* sA = smem;
    double* sB = smem + (size_t)EXPM_LDS * EXPM_MAXN;
    double* sx = sB + (size_t)EXPM_LDS * EXPM_MAXN;       // 64 doubles of scratch
    double* sU = sx + 64;                                  // compact U rows of the banded LU: sU[k*(KU+1) + c] = U(k, k+c)
    double* sL = sU + (size_t)EXPM_MAXN * (EXPM_KU + 1);   // multipliers: sL[k*KL + r-1] = L(k+r, k)
    int* sPiv = (int*)(sL + (size_t)EXPM_MAXN * EXPM_KL);
    __shared__ int s_piv, s_info, s_ns, s_kb;
    __shared__ double s_scale, s_hnorm, s_coef[8];
    const int tid = threadIdx.x;
    if (tid == 0) s_info = 0;
    __syncthreads();
    const int np = (n + 3) & ~3;
    double* gQ = work;             // the finished denominator polynomial q, parked while p is formed (thread-private tiles)

    // ---- stage H once; ||H||_inf by row sums in column order (dgpadm.f:241-253) from the staged copy ----------
    expm_load(sA, H, ldh, n, np, 1.0);
    if (tid == 0) s_kb = 0;
    __syncthreads();
    double rs = 0.0;
    int kb_row = 0;                                        // widest |i-j| of a non-zero entry in this row
    if (tid < n)
        for (int j = 0; j < n; ++j) {
            const double v = sA[(size_t)j * EXPM_LDS + tid];
            rs += fabs(v);
            if (v != 0.0) kb_row = max(kb_row, abs(j - tid));
        }
    if (kb_row) atomicMax(&s_kb, kb_row);
    {
        // block max through shared scratch
        double v = rs;
        for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
        if ((tid & 31) == 0) sx[tid >> 5] = v;
        __syncthreads();
        if (tid == 0) {
            double mxv = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) mxv = fmax(mxv, sx[w]);
            const double hnorm = fabs(t * mxv);
            s_hnorm = hnorm;
            int ns = 0;
            if (hnorm == 0.0) {
                s_info = KFSP_ERR_NULL_H;
            } else {
                // ns = max(0, int(log2(hnorm)) + 2), int() truncating toward zero (dgpadm.f:255), computed exactly
                int ex;
                const double fr = frexp(hnorm, &ex);               // hnorm = fr * 2^ex, fr in [0.5,1)
                int il;                                            // trunc(log2(hnorm))
                if (hnorm >= 1.0) il = ex - 1;
                else il = (fr == 0.5) ? ex - 1 : ex;               // ceil for negative logarithms
                ns = il + 2 > 0 ? il + 2 : 0;
                if (ns > 30) { s_info = KFSP_ERR_BAD_SIZES; ns = 30; }
            }
            s_ns = ns;
            s_scale = ldexp(t, -ns);                               // t / 2**ns
            // Pade coefficients (dgpadm.f:261-266), ideg = 6
            s_coef[0] = 1.0;
            for (int k = 1; k <= 6; ++k) s_coef[k] = (s_coef[k - 1] * (double)(7 - k)) / (double)(k * (13 - k));
        }
        __syncthreads();
    }
    if (s_info != 0) { *ns_out = 0; *hnorm_out = s_hnorm; return s_info; }
    EXPM_STAMP(1);                                         // H staged, norm, ns
    const double scale = s_scale, scale2 = __dmul_rn(scale, scale);
    double acc[4][4];
    // The Krylov H of IOP-2 is tridiagonal (plus the unit sub-diagonal entry): every Pade factor is banded
    // (H2: 2, q: 2-4-6, p: 0-2-4 then 5, q-p: 6) and only the squarings are dense.
    const bool tri = s_kb <= 1 && n > 2 * (EXPM_KL + EXPM_KU);
    const int DENSE = 1 << 20;
    const int bH = tri ? 1 : DENSE, bH2 = tri ? 2 : DENSE;
    // Tile -> thread.  Dense: tile (ti, tj) = (lane, warp).  Tridiagonal H: every Pade factor has bandwidth <= 6, i.e. only the
    // tiles with |tj - ti| <= 2 hold anything; with the dense mapping each of 26 warps would run the product loop for ~5 live
    // lanes (measured: 9.4k cycles per banded product, bound by warp issue), so warp d owns the tile diagonal tj = ti + d - 3
    // instead: 7 full warps.  Tiles nobody owns are zero in every operand and are never written.
    int ti = tid & 31, tj = tid >> 5;
    if (tri) {
        tj = ti + (tid >> 5) - EXPM_TD;
        if ((tid >> 5) > 2 * EXPM_TD || tj < 0 || 4 * tj >= n) { ti = 1 << 20; tj = 1 << 20; }     // no tile: every guard below fails
    }
    // one tile of a polynomial value lives in the registers of the thread that computed it: it goes to shared memory as the
    // left operand of the next product and comes back as that product's accumulator -- nothing of the Horner recurrence
    // passes through global memory (it used to: ~14 round trips of L2 latency per call, a third of the kernel at n = 100)
    auto add_diag = [&](double c) {
#pragma unroll
        for (int b2 = 0; b2 < 4; ++b2)
#pragma unroll
            for (int a2 = 0; a2 < 4; ++a2)
                if (4 * ti + a2 == 4 * tj + b2) acc[a2][b2] = __dadd_rn(acc[a2][b2], c);
    };

    // ---- H2 = scale2*H*H (dgpadm.f:270): alpha multiplies the right operand, as DGEMM does ----
    for (int t2 = tid; t2 < np * EXPM_LDS; t2 += (int)blockDim.x) sB[t2] = __dmul_rn(scale2, sA[t2]);     // the padding stays 0
    __syncthreads();
    expm_mma(sA, sB, n, acc, bH, bH, ti, tj);
    __syncthreads();
    expm_store(sB, EXPM_LDS, n, acc, 0.0, ti, tj);                 // sB <- H2 (stays for both Horner recurrences)
    EXPM_STAMP(2);
    // ---- Horner (dgpadm.f:274-301): q = ((c6 H2 + c4 I) H2 + c2 I) H2 + c0 I and p = ((c5 I) H2 + c3 I) H2 + c1 I are
    // independent of each other: one after the other instead of interleaved, so that a single register tile is live.
    // The first product of each has a diagonal left operand: (c I) H2 = c H2.
#pragma unroll
    for (int b2 = 0; b2 < 4; ++b2)
#pragma unroll
        for (int a2 = 0; a2 < 4; ++a2) acc[a2][b2] = fma(s_coef[6], acc[a2][b2], (4 * ti + a2 == 4 * tj + b2) ? s_coef[4] : 0.0);
    int bQ = tri ? 2 : DENSE;
    for (int k = 2; k >= 0; k -= 2) {                      // q = q*H2 + c2 I ; q = q*H2 + c0 I
        __syncthreads();
        expm_store(sA, EXPM_LDS, n, acc, 0.0, ti, tj);
        __syncthreads();
        expm_mma(sA, sB, n, acc, bQ, bH2, ti, tj);
        if (tri) bQ += 2;
        add_diag(s_coef[k]);
    }
#pragma unroll
    for (int b2 = 0; b2 < 4; ++b2)
#pragma unroll
        for (int a2 = 0; a2 < 4; ++a2) {
            const int i = 4 * ti + a2, j = 4 * tj + b2;
            if (i < n && j < n) gQ[(size_t)j * EXPM_LDS + i] = acc[a2][b2];
            acc[a2][b2] = (i == j) ? s_coef[5] : 0.0;      // p = c5 I
        }
    int bP = tri ? 0 : DENSE;
    for (int k = 3; k >= 1; k -= 2) {                      // p = p*H2 + c3 I ; p = p*H2 + c1 I
        __syncthreads();
        expm_store(sA, EXPM_LDS, n, acc, 0.0, ti, tj);
        __syncthreads();
        expm_mma(sA, sB, n, acc, bP, bH2, ti, tj);
        if (tri) bP += 2;
        add_diag(s_coef[k]);
    }
    EXPM_STAMP(3);                                         // both Horner recurrences
    // p = scale * p * H (dgpadm.f:309-312)
    __syncthreads();
    expm_store(sA, EXPM_LDS, n, acc, 0.0, ti, tj);
    expm_load(sB, H, ldh, n, np, scale);
    __syncthreads();
    expm_mma(sA, sB, n, acc, bP, bH, ti, tj);
    __syncthreads();
    // ---- sA <- q - p ; sB <- p ; solve (q-p) X = p (dgpadm.f:314-315) ----------------------
    expm_store(sB, EXPM_LDS, n, acc, 0.0, ti, tj);
#pragma unroll
    for (int b2 = 0; b2 < 4; ++b2)
#pragma unroll
        for (int a2 = 0; a2 < 4; ++a2) {
            const int i = 4 * ti + a2, j = 4 * tj + b2;
            if (i < n && j < n) sA[(size_t)j * EXPM_LDS + i] = __dsub_rn(gQ[(size_t)j * EXPM_LDS + i], acc[a2][b2]);
        }
    __syncthreads();
    EXPM_STAMP(4);                                         // p*H, q - p
    if (tri) {
        // ---- banded path: q-p has lower bandwidth 6; with partial pivoting U has upper bandwidth 12 ---------
        // Phase A: ONE warp factors the band with the ACTIVE WINDOW IN REGISTERS: at step k the elimination touches rows
        // k..k+6 of columns k..k+12; lane c holds those 7 entries of column k+c.  The pivot search is seven compares in lane
        // 0, a row swap is a register exchange in every lane, the multipliers travel by shuffle, and sliding the window one
        // step is a shuffle from the lane to the right plus ONE shared-memory load per lane (the entry of row k+7, which no
        // earlier step has touched; everything above it in the incoming column is outside the band of q-p, i.e. zero).
        // sA is only read.  Every element sees the same operations, in the same order, as in the dense elimination below
        // (terms with an exactly-zero multiplier or pivot-row entry leave a value unchanged).  Round 1 kept the window in
        // shared memory behind __syncwarp: 1390 cycles per column against ~350 (profiles/r2_summary.md, section 5).
        if (tid < 32) {
            const int lane = tid;
            double w[EXPM_KL + 1];
#pragma unroll
            for (int r = 0; r <= EXPM_KL; ++r) w[r] = (lane <= EXPM_KU && lane < n && r < n) ? sA[(size_t)lane * EXPM_LDS + r] : 0.0;
#pragma unroll 1
            for (int k = 0; k < n; ++k) {
                int pr = 0;
                double best = fabs(w[0]);
#pragma unroll
                for (int r = 1; r <= EXPM_KL; ++r) {
                    const double v = fabs(w[r]);
                    if (k + r < n && v > best) { best = v; pr = r; }       // ties: the lowest row, as the reference's IDAMAX
                }
                pr = __shfl_sync(0xffffffffu, pr, 0);
                best = __shfl_sync(0xffffffffu, best, 0);
                if (best == 0.0) { if (lane == 0) s_info = KFSP_ERR_SINGULAR; break; }
#pragma unroll
                for (int r = 1; r <= EXPM_KL; ++r)
                    if (r == pr) { const double tmp = w[r]; w[r] = w[0]; w[0] = tmp; }
                const double inv = 1.0 / w[0];                             // meaningful in lane 0
                double m[EXPM_KL + 1];
#pragma unroll
                for (int r = 1; r <= EXPM_KL; ++r) {
                    const double mine = k + r < n ? __dmul_rn(w[r], inv) : 0.0;
                    m[r] = __shfl_sync(0xffffffffu, mine, 0);
                }
                if (lane == 0) {
#pragma unroll
                    for (int r = 1; r <= EXPM_KL; ++r) sL[k * EXPM_KL + r - 1] = m[r];
                    sPiv[k] = k + pr;
                }
                if (lane <= EXPM_KU) sU[k * (EXPM_KU + 1) + lane] = k + lane < n ? w[0] : 0.0;
                if (lane >= 1 && lane <= EXPM_KU && k + lane < n) {
#pragma unroll
                    for (int r = 1; r <= EXPM_KL; ++r)
                        if (k + r < n) w[r] = fma(-m[r], w[0], w[r]);
                }
                // slide: rows k+1..k+6 of column k+1+c come from the lane to the right, row k+7 from shared memory
#pragma unroll
                for (int r = 0; r < EXPM_KL; ++r) w[r] = __shfl_down_sync(0xffffffffu, w[r + 1], 1);
                if (lane >= EXPM_KU) {
#pragma unroll
                    for (int r = 0; r < EXPM_KL; ++r) w[r] = 0.0;          // the incoming column: above the band
                }
                w[EXPM_KL] = (lane <= EXPM_KU && k + 1 + lane < n && k + 1 + EXPM_KL < n) ? sA[(size_t)(k + 1 + lane) * EXPM_LDS + k + 1 + EXPM_KL] : 0.0;
            }
        }
        __syncthreads();
        if (s_info != 0) { *ns_out = s_ns; *hnorm_out = s_hnorm; return s_info; }
        EXPM_STAMP(5);                                     // banded LU
        // Phase B: the n right-hand sides are independent -> transpose B so that thread j owns column j with
        // conflict-free shared-memory accesses, then forward and backward substitution without any barrier, the entries
        // a step touches (7 going down, 13 going up) in a sliding register window.
        for (int t2 = tid; t2 < n * n; t2 += (int)blockDim.x) {
            const int i = t2 / n, j = t2 % n;
            sA[(size_t)i * EXPM_LDS + j] = sB[(size_t)j * EXPM_LDS + i];
        }
        __syncthreads();
        if (tid < n) {
            double* b = sA + tid;                                  // element i of this column: b[i * EXPM_LDS]
            double f[EXPM_KL + 1];
#pragma unroll
            for (int r = 0; r <= EXPM_KL; ++r) f[r] = r < n ? b[(size_t)r * EXPM_LDS] : 0.0;
#pragma unroll 1
            for (int k = 0; k < n; ++k) {
                const int pr = sPiv[k] - k;
#pragma unroll
                for (int r = 1; r <= EXPM_KL; ++r)
                    if (r == pr) { const double tmp = f[r]; f[r] = f[0]; f[0] = tmp; }
                const double bk = f[0];
                b[(size_t)k * EXPM_LDS] = bk;
#pragma unroll
                for (int r = 1; r <= EXPM_KL; ++r) f[r - 1] = (k + r < n) ? fma(-sL[k * EXPM_KL + r - 1], bk, f[r]) : 0.0;
                f[EXPM_KL] = (k + 1 + EXPM_KL < n) ? b[(size_t)(k + 1 + EXPM_KL) * EXPM_LDS] : 0.0;
            }
            double u[EXPM_KU + 1];
#pragma unroll
            for (int c = 0; c <= EXPM_KU; ++c) u[c] = (n - 1 - c >= 0) ? b[(size_t)(n - 1 - c) * EXPM_LDS] : 0.0;
#pragma unroll 1
            for (int k = n - 1; k >= 0; --k) {
                const double xk = u[0] / sU[k * (EXPM_KU + 1)];
                b[(size_t)k * EXPM_LDS] = xk;
#pragma unroll
                for (int c = 1; c <= EXPM_KU; ++c) u[c - 1] = (k - c >= 0) ? fma(-xk, sU[(k - c) * (EXPM_KU + 1) + c], u[c]) : 0.0;
                u[EXPM_KU] = (k - 1 - EXPM_KU >= 0) ? b[(size_t)(k - 1 - EXPM_KU) * EXPM_LDS] : 0.0;
            }
        }
        __syncthreads();
        // back to the column-major operand layout, fused with E = I + 2 X
        for (int t2 = tid; t2 < np * EXPM_LDS; t2 += (int)blockDim.x) {
            const int j = t2 / EXPM_LDS, i = t2 % EXPM_LDS;
            sB[t2] = (i < n && j < n) ? __dadd_rn(__dmul_rn(2.0, sA[(size_t)i * EXPM_LDS + j]), (i == j ? 1.0 : 0.0)) : 0.0;
        }
        __syncthreads();
    } else {
    // Gaussian elimination with partial pivoting on [sA | sB]
    for (int k = 0; k < n; ++k) {
        if (tid < 32) {
            double best = -1.0; int bi = k;
            for (int i = k + tid; i < n; i += 32) {
                const double v = fabs(sA[(size_t)k * EXPM_LDS + i]);
                if (v > best) { best = v; bi = i; }
            }
            for (int o = 16; o > 0; o >>= 1) {
                const double ob = __shfl_down_sync(0xffffffffu, best, o);
                const int oi = __shfl_down_sync(0xffffffffu, bi, o);
                if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
            }
            if (tid == 0) { s_piv = bi; if (best == 0.0) s_info = KFSP_ERR_SINGULAR; }
        }
        __syncthreads();
        if (s_info != 0) break;
        const int p = s_piv;
        if (p != k) {
            for (int j = tid; j < 2 * n; j += (int)blockDim.x) {
                double* col = (j < n ? sA + (size_t)j * EXPM_LDS : sB + (size_t)(j - n) * EXPM_LDS);
                const double tmp = col[k]; col[k] = col[p]; col[p] = tmp;
            }
            __syncthreads();
        }
        const double inv = 1.0 / sA[(size_t)k * EXPM_LDS + k];
        __syncthreads();
        for (int i = k + 1 + tid; i < n; i += (int)blockDim.x) sA[(size_t)k * EXPM_LDS + i] = __dmul_rn(sA[(size_t)k * EXPM_LDS + i], inv);
        __syncthreads();
        const int rows = n - k - 1;
        if (rows > 0) {
            const int colsA = n - k - 1;
            const int total = rows * (colsA + n);
            for (int t2 = tid; t2 < total; t2 += (int)blockDim.x) {
                const int i = k + 1 + t2 % rows, c = t2 / rows;
                double* col = c < colsA ? sA + (size_t)(k + 1 + c) * EXPM_LDS : sB + (size_t)(c - colsA) * EXPM_LDS;
                col[i] = fma(-sA[(size_t)k * EXPM_LDS + i], col[k], col[i]);
            }
        }
        __syncthreads();
    }
    if (s_info != 0) { *ns_out = s_ns; *hnorm_out = s_hnorm; return s_info; }
    // back substitution U X = Y, column oriented
    for (int k = n - 1; k >= 0; --k) {
        const double ukk = sA[(size_t)k * EXPM_LDS + k];
        for (int j = tid; j < n; j += (int)blockDim.x) sB[(size_t)j * EXPM_LDS + k] /= ukk;
        __syncthreads();
        const int total = k * n;
        for (int t2 = tid; t2 < total; t2 += (int)blockDim.x) {
            const int i = t2 % k, j = t2 / k;
            sB[(size_t)j * EXPM_LDS + i] = fma(-sB[(size_t)j * EXPM_LDS + k], sA[(size_t)k * EXPM_LDS + i], sB[(size_t)j * EXPM_LDS + i]);
        }
        __syncthreads();
    }
    // ---- E = I + 2 X (dgpadm.f:317-320); iodd == 0 so no sign flip (:322-325) -----------------
    for (int t2 = tid; t2 < np * EXPM_LDS; t2 += (int)blockDim.x) {
        const int j = t2 / EXPM_LDS, i = t2 % EXPM_LDS;
        sB[t2] = (i < n && j < n) ? __dadd_rn(__dmul_rn(2.0, sB[t2]), (i == j ? 1.0 : 0.0)) : 0.0;
    }
    __syncthreads();
    }   // dense path
    EXPM_STAMP(6);                                         // substitution, E = I + 2X
#ifdef KFSP_EXPM_TIMING
    if (threadIdx.x == 0)
        printf("expm n=%d ns=%d tri=%d cycles: stage+norm %lld  H2 %lld  horner %lld  pH,q-p %lld  LU %lld  solve %lld\n", n, s_ns, (int)tri,
               s_stamp[1] - s_stamp[0], s_stamp[2] - s_stamp[1], s_stamp[3] - s_stamp[2], s_stamp[4] - s_stamp[3], s_stamp[5] - s_stamp[4],
               s_stamp[6] - s_stamp[5]);
#endif
    *ns_out = s_ns;
    *hnorm_out = s_hnorm;
    return 0;
}

// One CTA of 1024 threads.
//   H, ldh        : the device Hessenberg matrix (column-major)
//   mx_ok, t_ok   : order and step when the sweep did not break down
//   use_brk,t_brk : if use_brk and ctl->brk > 0, order = ctl->brk and step = t_brk (KrylovSolver.f90:249-256, 271)
//   set_one       : if >= 0, first set H(set_one+2, set_one+1) = 1 (label 300, KrylovSolver.f90:266; set_one = M)
//   work          : 6 * EXPM_MAXN^2 doubles of global scratch (the last two: E ping-pong of the cluster variant)
//   full_out      : optional mx*mx output of the whole exponential (tests)
__device__ __forceinline__ void expm_result(ExpmResult* res, int info, int ns, int n, int brk, double hnorm, double t, const SweepCtl* ctl) {
    res->info = info; res->ns = ns; res->mx = n; res->brk = brk; res->hnorm = hnorm; res->t_used = t;
    res->avnorm = ctl ? ctl->scal[SC_AVNORM] : 0.0;
}
__global__ void __launch_bounds__(EXPM_THREADS, 1) k_expm(double* H, int ldh, int mx_ok, double t_ok, int use_brk, double t_brk,
                                                          int set_one, const SweepCtl* ctl, double* work, ExpmResult* res,
                                                          double* full_out) {
    extern __shared__ __align__(16) double smem[];
    double* sA = smem;
    double* sB = smem + (size_t)EXPM_LDS * EXPM_MAXN;
    const int tid = threadIdx.x;
    int n = mx_ok;
    double t = t_ok;
    const int brk = ctl ? ctl->brk : 0;
    if (use_brk && brk > 0) { n = brk; t = t_brk; }
    if (set_one >= 0 && tid == 0) H[(size_t)set_one * ldh + set_one + 1] = 1.0;
    __syncthreads();
    const int np = (n + 3) & ~3;
    int ns = 0;
    double hnorm = 0.0;
    const int info = expm_pade(H, ldh, n, t, work, smem, &ns, &hnorm);
    if (info != 0) {
        if (tid == 0) expm_result(res, info, ns, n, brk, hnorm, t, ctl);
        return;
    }
    double acc[4][4];
    // ---- squarings (dgpadm.f:329-336) ------------------------------------------------------
    for (int s = 0; s < ns; ++s) {
        for (int t2 = tid; t2 < np * EXPM_LDS; t2 += (int)blockDim.x) sA[t2] = sB[t2];
        __syncthreads();
        expm_mma(sA, sB, n, acc);
        __syncthreads();
        expm_store(sB, EXPM_LDS, n, acc, 0.0);
        __syncthreads();
    }
    // ---- results -----------------------------------------------------------------------------
    for (int i = tid; i < EXPM_MAXN; i += (int)blockDim.x) res->e[i] = i < n ? sB[i] : 0.0;
    if (full_out)
        for (int t2 = tid; t2 < n * n; t2 += (int)blockDim.x) full_out[t2] = sB[(size_t)(t2 / n) * EXPM_LDS + t2 % n];
    if (tid == 0) expm_result(res, 0, ns, n, brk, hnorm, t, ctl);
}

// The same on a thread-block CLUSTER of EXPM_CLUSTER CTAs (one per SM): CTA 0 runs expm_pade, then the ns squarings -- which
// are dense (n x n x n each) and at n ~ 100 cost 55 us apiece at the FP64 rate of a single SM, 5-10 of them per call -- are
// split by column tiles over the CTAs: every CTA stages the current E in its shared memory, computes its tiles of E*E with the
// same per-element operation order (k ascending from an accumulator that starts at 0: bit-identical to k_expm), writes them to
// the other E buffer in global memory (L2-resident, 85 KB), and a cluster barrier (release/acquire) separates the squarings.
constexpr int EXPM_CLUSTER = 8;
__device__ __forceinline__ void cluster_barrier() {
    __threadfence();
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned cluster_cta_rank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__global__ void __launch_bounds__(EXPM_THREADS, 1) k_expm_cluster(double* H, int ldh, int mx_ok, double t_ok, int use_brk, double t_brk,
                                                                  int set_one, const SweepCtl* ctl, double* work, ExpmResult* res,
                                                                  double* full_out) {
    extern __shared__ __align__(16) double smem[];
    double* sA = smem;
    double* sB = smem + (size_t)EXPM_LDS * EXPM_MAXN;
    const int tid = threadIdx.x;
    const int cta = (int)cluster_cta_rank();
    int n = mx_ok;
    double t = t_ok;
    const int brk = ctl ? ctl->brk : 0;
    if (use_brk && brk > 0) { n = brk; t = t_brk; }
    const int np = (n + 3) & ~3;
    const size_t nn = (size_t)EXPM_MAXN * EXPM_MAXN;
    double* gE[2] = {work + 4 * nn, work + 5 * nn};
    volatile int* hdr = reinterpret_cast<volatile int*>(work + 6 * nn);      // {info, ns} for the other CTAs
    int ns = 0;
    double hnorm = 0.0;
    if (cta == 0) {
        if (set_one >= 0 && tid == 0) H[(size_t)set_one * ldh + set_one + 1] = 1.0;
        __syncthreads();
        const int info = expm_pade(H, ldh, n, t, work, smem, &ns, &hnorm);
        if (info == 0)
            for (int t2 = tid; t2 < np * EXPM_LDS; t2 += EXPM_THREADS) gE[0][t2] = sB[t2];
        if (tid == 0) { hdr[0] = info; hdr[1] = ns; }
        if (info != 0 && tid == 0) expm_result(res, info, ns, n, brk, hnorm, t, ctl);
    }
    cluster_barrier();
    if (hdr[0] != 0) return;                                               // uniform over the cluster
    ns = hdr[1];
    // column tiles of this CTA: tj in [cta*TJ, (cta+1)*TJ), one warp per column tile, 32 row tiles per warp
    const int ntile = np / 4;
    const int TJ = (ntile + EXPM_CLUSTER - 1) / EXPM_CLUSTER;
    for (int s = 0; s < ns; ++s) {
        const double* cur = gE[s & 1];
        double* nxt = gE[(s + 1) & 1];
        for (int t2 = tid; t2 < np * EXPM_LDS; t2 += EXPM_THREADS) sA[t2] = __ldcg(cur + t2);
        __syncthreads();
        const int ti = tid & 31, tjl = tid >> 5, tj = cta * TJ + tjl;
        if (tjl < TJ && 4 * ti < n && 4 * tj < n) {
            double acc[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b2 = 0; b2 < 4; ++b2) acc[a][b2] = 0.0;
            const double* pa = sA + 4 * ti;
            const double* pb = sA + (size_t)(4 * tj) * EXPM_LDS;
            for (int k = 0; k < n; ++k) {
                const double2 a01 = *reinterpret_cast<const double2*>(pa + (size_t)k * EXPM_LDS);
                const double2 a23 = *reinterpret_cast<const double2*>(pa + (size_t)k * EXPM_LDS + 2);
                const double b0 = pb[k], b1 = pb[EXPM_LDS + k], b2 = pb[2 * EXPM_LDS + k], b3 = pb[3 * EXPM_LDS + k];
                acc[0][0] = fma(a01.x, b0, acc[0][0]); acc[1][0] = fma(a01.y, b0, acc[1][0]);
                acc[2][0] = fma(a23.x, b0, acc[2][0]); acc[3][0] = fma(a23.y, b0, acc[3][0]);
                acc[0][1] = fma(a01.x, b1, acc[0][1]); acc[1][1] = fma(a01.y, b1, acc[1][1]);
                acc[2][1] = fma(a23.x, b1, acc[2][1]); acc[3][1] = fma(a23.y, b1, acc[3][1]);
                acc[0][2] = fma(a01.x, b2, acc[0][2]); acc[1][2] = fma(a01.y, b2, acc[1][2]);
                acc[2][2] = fma(a23.x, b2, acc[2][2]); acc[3][2] = fma(a23.y, b2, acc[3][2]);
                acc[0][3] = fma(a01.x, b3, acc[0][3]); acc[1][3] = fma(a01.y, b3, acc[1][3]);
                acc[2][3] = fma(a23.x, b3, acc[2][3]); acc[3][3] = fma(a23.y, b3, acc[3][3]);
            }
            // the padding rows / columns stay zero: rows >= n of a tile are products of zero rows of E
#pragma unroll
            for (int b2 = 0; b2 < 4; ++b2)
#pragma unroll
                for (int a = 0; a < 4; ++a) {
                    const int i = 4 * ti + a, j = 4 * tj + b2;
                    if (i < np && j < np) nxt[(size_t)j * EXPM_LDS + i] = (i < n && j < n) ? acc[a][b2] : 0.0;
                }
        }
        cluster_barrier();
    }
    if (cta != 0) return;
    const double* E = gE[ns & 1];
    for (int i = tid; i < EXPM_MAXN; i += EXPM_THREADS) res->e[i] = i < n ? __ldcg(E + i) : 0.0;
    if (full_out)
        for (int t2 = tid; t2 < n * n; t2 += EXPM_THREADS) full_out[t2] = __ldcg(E + (size_t)(t2 / n) * EXPM_LDS + t2 % n);
    if (tid == 0) expm_result(res, 0, ns, n, brk, hnorm, t, ctl);
}

}  // namespace kfsp

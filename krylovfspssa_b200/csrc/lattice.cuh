// Matrix-free generator SpMV on a lattice (spmv_variant = 1).
//
// FMATVEC (src/fsp/KrylovSolver.f90:577-607) streams ADJ/OFFDIAG/DIAG, 12R+24 bytes per state.  When the
// projection is a full box [0,B_1) x ... x [0,B_S) held in the reference's natural order (first species
// fastest, index = sum_s x_s * stride_s) and every propensity reads at most ONE species, none of the three
// arrays carries information: the predecessor index is i - sum_s nu_ks*stride_s, legality is a bounds check
// on x - nu_k, and a_k(x - nu_k) is one entry of a per-reaction table over that species' count (the same
// host-built tables the explicit path evaluates its OFFDIAG from, so the values are the same doubles).
// The row is then evaluated with the explicit kernel's operation order
//     y_i = -(d_i * x_i);  y_i = fma(a_k(x_i - nu_k), x[pred_k], y_i)  for k = 1..R,   d_i = ((a_1+a_2)+...)
// and the result is bit-identical to k_spmv on the explicit matrix (tests/test_gpu_lattice.py), for
// 16 bytes of HBM traffic per state (x read once, y written once) instead of 12R+24.
//
// Work decomposition: a CTA owns up to 256 consecutive "columns" c (all species but the slowest one, flattened)
// and walks a chunk of the slowest species z.  Everything that depends on the column only -- the decoded
// counts, bounds checks, predecessor offsets, table entries of the column species -- is hoisted out of the
// walk; rows z-1, z, z+1 of x are re-read through L1 and row z+PF is prefetched, so every x element comes
// from HBM once per chunk.
#pragma once
#include "common.cuh"
#include "krylov.cuh"

namespace kfsp {

constexpr int BOX_MAX_R = 16;

struct Lattice {
    int32_t S, R;
    int32_t B[KFSP_MAX_SPECIES];          // bounds: 0 <= x_s < B[s]
    int64_t stride[KFSP_MAX_SPECIES];     // stride[0] = 1, stride[s] = stride[s-1]*B[s-1]
    int64_t plane;                        // stride[S-1]: rows per unit of the slowest species z = x_{S-1}
    int32_t nz;                           // B[S-1]
    int32_t zlo, zhi;                     // this rank's slab of z (one GPU: 0, nz)
    int32_t nu[BOX_MAX_R][KFSP_MAX_SPECIES];
    int32_t sp[BOX_MAX_R];                // the species reaction k's propensity reads
    const double* tab[BOX_MAX_R];         // a_k as a function of that species' count, 0..max_molecules
    int32_t zb[MAX_RANKS + 1];            // slab boundaries of every rank
};

__device__ __forceinline__ void lattice_decode(const Lattice& L, int64_t g, int32_t* st) {
    for (int s = 0; s < L.S; ++s) {
        st[s] = (int32_t)(g % L.B[s]);
        g /= L.B[s];
    }
}

// states[t*S + s] for global rows [g0, g0+count)
__global__ void k_box_gen_states(const __grid_constant__ Lattice L, int64_t g0, int64_t count, int32_t* states) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
        int32_t st[KFSP_MAX_SPECIES];
        lattice_decode(L, g0 + t, st);
        for (int s = 0; s < L.S; ++s) states[t * L.S + s] = st[s];
    }
}
// the caller's state list must BE the lattice in natural order: *bad is raised otherwise
__global__ void k_box_check_states(const __grid_constant__ Lattice L, int64_t g0, int64_t count, const int32_t* __restrict__ states, int32_t* bad) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
        int32_t st[KFSP_MAX_SPECIES];
        lattice_decode(L, g0 + t, st);
        bool ok = true;
        for (int s = 0; s < L.S; ++s) ok = ok && states[t * L.S + s] == st[s];
        if (!ok) atomicOr(bad, 1);
    }
}

// x at column c of global plane zz (any rank's slab when HALO)
template <int HALO>
__device__ __forceinline__ double lattice_load(const Lattice& L, const double* __restrict__ x, const DistPeers* __restrict__ dp,
                                               int64_t c, int32_t zz, int64_t coloff) {
    if (HALO == 0 || (zz >= L.zlo && zz < L.zhi)) return x[c + L.plane * (zz - L.zlo)];
    int r = 0;
    while (zz >= L.zb[r + 1]) ++r;
    return __ldcg(dp->V[r] + coloff + c + L.plane * (zz - L.zb[r]));
}

template <int LEVEL>
__device__ __forceinline__ void lattice_prefetch(const void* p) {
    if (LEVEL == 1) asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    if (LEVEL == 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// mode 0: y = A x   mode 1: + dA, dB, dC of the Arnoldi column   mode 2: + ssq = <y, y>     (same contract as k_spmv)
// Every reaction runs the same few instructions with per-thread data hoisted out of the walk over z: the
// element offset of its predecessor, the bounds check of the column species, and the table entries that do
// not depend on z.  The rows z-1, z, z+1 of x are re-read through L1 (the CTA touched them one and two
// iterations earlier); the first touch of row z+PF is a prefetch, so the depth of the memory pipeline does
// not cost registers.
// ST > 0: number of species fixed at compile time (stoichiometry entries become constant-bank operands).
template <int RT, int ST, int MODE, int HALO, int PFL, int PF, int MINB>
__global__ void __launch_bounds__(VEC_THREADS, MINB) k_spmv_box(const __grid_constant__ Lattice L, int zc, int cbw, const double* __restrict__ x,
                                                                double* __restrict__ y, const double* __restrict__ first, Reducer rd,
                                                                SweepCtl* ctl, EpiArgs ea, int cf, int64_t coloff) {
    constexpr int R = RT;
    pdl_trigger();
    pdl_wait();
    if (MODE != 0 && ctl->brk != 0) return;
    const bool has_g = MODE == 1 && ea.has_g;             // `first` = g = U_{c-1}
    const double fs = has_g ? col_scale(ctl, cf) : 0.0;
    const int S = ST > 0 ? ST : L.S;
    const int64_t plane = L.plane;
    const int nzl = L.zhi - L.zlo;
    const int64_t ncb = (plane + cbw - 1) / cbw;                 // column blocks of cbw <= 256 columns (even split of the plane)
    const int64_t nzc = (nzl + zc - 1) / zc;
    const DistPeers* __restrict__ dp = rd.peers;
    DD acc, accA, accC;
    acc.hi = acc.lo = accA.hi = accA.lo = accC.hi = accC.lo = 0.0;
    for (int64_t item = blockIdx.x; item < ncb * nzc; item += gridDim.x) {
        const int64_t c = (item % ncb) * cbw + threadIdx.x;
        if ((int)threadIdx.x >= cbw || c >= plane) continue;
        const int32_t z0 = L.zlo + (int32_t)(item / ncb) * zc;
        const int32_t z1 = min(z0 + zc, L.zhi);
        // column invariants (local element indices fit 32 bits: n <= 2e9)
        bool lowok[R];
        double ad[R], ac[R];
        int32_t off[R];
        const int32_t plane32 = (int32_t)plane;
#pragma unroll
        for (int k = 0; k < R; ++k) { lowok[k] = true; ad[k] = 0.0; ac[k] = 0.0; off[k] = -L.nu[k][S - 1] * plane32; }
        {
            int64_t rem = c;
#pragma unroll
            for (int s = 0; s < (ST > 0 ? ST : KFSP_MAX_SPECIES) - 1; ++s) {
                if (s >= S - 1) break;
                const int32_t xv = (int32_t)(rem % L.B[s]);
                rem /= L.B[s];
#pragma unroll
                for (int k = 0; k < R; ++k) {
                    const int32_t v = xv - L.nu[k][s];
                    const bool in = v >= 0 && v < L.B[s];
                    lowok[k] = lowok[k] && in;
                    off[k] -= L.nu[k][s] * (int32_t)L.stride[s];
                    if (L.sp[k] == s) {
                        ad[k] = __ldg(L.tab[k] + xv);
                        ac[k] = in ? __ldg(L.tab[k] + v) : 0.0;
                    }
                }
            }
        }
        uint32_t i = (uint32_t)(c + plane * (z0 - L.zlo));
        const uint32_t pfo = (uint32_t)(plane32 * PF);
        if (PFL > 0) {
#pragma unroll
            for (int q = 1; q < PF; ++q)
                if (z0 + q < L.zhi) { lattice_prefetch<PFL>(x + (i + plane32 * q)); if (has_g) lattice_prefetch<PFL>(first + (i + plane32 * q)); }
        }
        for (int32_t z = z0; z < z1; ++z) {
            if (PFL > 0 && z + PF < L.zhi) { lattice_prefetch<PFL>(x + (i + pfo)); if (has_g) lattice_prefetch<PFL>(first + (i + pfo)); }
            const double x0 = x[i];
            double f = 0.0;
            if (has_g) f = __dmul_rn(fs, __ldcs(first + i));
            double d = 0.0;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                if (L.sp[k] == S - 1) {                             // uniform: this reaction's table runs over z
                    const int32_t zz = z - L.nu[k][S - 1];
                    ad[k] = __ldg(L.tab[k] + z);
                    ac[k] = (zz >= 0 && zz < L.nz) ? __ldg(L.tab[k] + zz) : 0.0;
                }
                d = __dadd_rn(d, ad[k]);
            }
            double sv = -__dmul_rn(d, x0);
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int32_t zz = z - L.nu[k][S - 1];
                if (lowok[k] && zz >= 0 && zz < L.nz) {
                    double xv;
                    if (HALO != 0 && (zz < L.zlo || zz >= L.zhi)) {
                        int r = 0;
                        while (zz >= L.zb[r + 1]) ++r;
                        const int64_t cc = c + (int64_t)off[k] + (int64_t)L.nu[k][S - 1] * plane;      // column of the predecessor
                        xv = __ldcg(dp->V[r] + coloff + cc + plane * (zz - L.zb[r]));
                    } else {
                        xv = x[(uint32_t)(i + off[k])];
                    }
                    sv = fma(ac[k], xv, sv);
                }
            }
            __stcs(y + i, sv);
            if (MODE == 1) {
                dd_add_prod(acc, x0, sv);
                if (has_g) { dd_add_prod(accA, f, sv); dd_add_prod(accC, x0, f); }
            }
            if (MODE == 2) dd_add_prod(acc, sv, sv);
            i += (uint32_t)plane32;
        }
    }
    if (MODE == 0) return;
    DD zz0; zz0.hi = 0.0; zz0.lo = 0.0;
    if (MODE == 1) {
        DD v[4] = {zz0, accA, acc, accC};
        double tot[4];
        if (grid_reduce<4>(v, tot, rd) && threadIdx.x == 0) reduce_epilogue(ea, tot, ctl);
    } else {
        DD v[2] = {zz0, acc};
        double tot[2];
        if (grid_reduce<2>(v, tot, rd) && threadIdx.x == 0) reduce_epilogue(ea, tot, ctl);
    }
}

// ---------------------------------------------------------------------------------------
// Fast path for the network shape of the synthetic large-copy-number workload (BASELINE config 5) and of the
// reference's 2-species models: two species, four reactions, each adding or removing ONE molecule of one
// species, in one of the two orders the reference's model files use
//     ORD 0:  0->X, X->0, 0->Y, Y->0   (toggle_model.input, bursting_gene_model.input)
//     ORD 1:  0->X, 0->Y, X->0, Y->0   (toggle_test_model.input = config 5)
// each propensity a table over x or over y (bit k of TS set: reaction k's table runs over y = z).
// With the stencil known at compile time nothing is re-read from memory: rows z-1, z, z+1 of a column live in a
// rolling register window fed by ONE first-touch load per row (issued PF rows before its use through a register
// ring), the +-1 neighbours in x are the adjacent lanes' registers (warp shuffle: a warp covers 30 columns plus
// one halo lane on each side, which loads but does not store), the y-tables of a chunk sit in shared memory.
//
// FIN = 1 fuses the tail of the previous Arnoldi column into the load stage (KrylovSolver.f90:243-258): the operand is
// finalised on the fly,  U_c(i) = fma(-h2, v_g(i), fma(-h1, v_f(i), cs*Yp(i)))  with Yp = A U_{c-1} (scratch column written
// by the previous launch), g = U_{c-1}, f = U_{c-2}, stored once, its norm^2 accumulated beside the three inner products of
// the new column (krylov.cuh: dA, dB, dC), and the generator product is taken on U_c un-normalised -- the column scale
// 1/||U_c|| is only known when the pass ends, so it is applied to the scalars in the epilogue.  The finalised column cannot be
// written over its own source because other warps (and, on several GPUs, the neighbours) still gather rows of Yp for their
// stencil, hence the two scratch columns used alternately.  Per state and Arnoldi column ONE launch moves
// Yp + g + f + U_c + y = 40 bytes and ends in ONE reduction point (the reference's BLAS sequence: 24 + 104 bytes, three).
//
// Rows are walked in three pieces so that the hot loop carries no bounds checks: a head (first PF rows when the chunk
// starts at the box boundary z = 0), the main loop (every ring refill is a local, in-range row), and a tail (refills
// that leave the chunk / the slab / the box, read through the general loader, which on several GPUs takes the halo
// row from the owner's HBM over NVLink).
// Same operation order as k_spmv_box / k_spmv (reactions in model order): bit-identical results
// (tests/test_gpu_lattice.py runs all three).
// ---------------------------------------------------------------------------------------
enum Bd2Dir : int { BD_XP = 0, BD_XM = 1, BD_YP = 2, BD_YM = 3 };      // the species and sign of the reaction's step
__host__ __device__ constexpr int bd2_dir(int ord, int k) {
    return ord == 0 ? k : (k == 0 ? BD_XP : k == 1 ? BD_YP : k == 2 ? BD_XM : BD_YM);
}
// -1 if the lattice is not of this family, else the ORD id
inline int lattice_bd2_order(const Lattice& L) {
    if (L.S != 2 || L.R != 4) return -1;
    const int step[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}};
    for (int ord = 0; ord < 2; ++ord) {
        bool ok = true;
        for (int k = 0; k < 4; ++k) {
            const int d = bd2_dir(ord, k);
            ok = ok && L.nu[k][0] == step[d][0] && L.nu[k][1] == step[d][1];
        }
        if (ok) return ord;
    }
    return -1;
}
constexpr int BD2_WCOLS = 30;                                  // live columns per warp (lanes 1..30; lanes 0 and 31 are halo)
constexpr int BD2_CBW = BD2_WCOLS * (VEC_THREADS / 32);        // live columns per CTA
constexpr int BD2_L2AHEAD = 10;                                // rows ahead of the ring that are prefetched into L2
constexpr int BD2_PF = 4;                                      // depth of the register ring (rows in flight per thread and stream)
constexpr int BD2_MINB = 4;                                    // resident CTAs per SM the kernel is compiled for
// The rows in flight live in a shared-memory ring filled by per-thread asynchronous copies (cp.async, 8 bytes per thread and
// stream: every lane copies and later reads its own element, so no barrier is involved), BD2_DEPTH rows deep.  (A register
// ring is bounded by the register file: at 4 rows ncu showed 50 % of all stall cycles on the scoreboard of the first-touch
// loads, and the launch was 10 % slower; profiles/r2_summary.md.)
constexpr int BD2_DEPTH = 6;                                   // 3 streams x 6 rows x 256 threads x 8 B = 36 KB per CTA, 4 CTAs per SM
__device__ __forceinline__ void cp_async8(double* smem, const double* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
constexpr int BD2_ZT = 256;                                    // longest z-chunk (rows of the y-tables staged in shared memory)
struct Bd2Args {
    const double* src;        // FIN: Yp = A U_{c-1} (scratch column); else: the operand column U_c itself
    double* xout;             // FIN: the finalised column U_c is stored here
    double* y;                // A U_c
    const double* g;          // U_{c-1}: previous basis vector (finalising axpy and dA, dC); unused if !has_g
    const double* f;          // U_{c-2}: FIN and has_f only
    EpiArgs ea;               // what the reduction point produces (krylov.cuh)
    int32_t cg, cf;           // basis columns of g and f (their scales)
    int32_t has_g, has_f;
    int32_t zc, halo;         // rows per z-chunk; 1 = several GPUs (rows outside the slab come from the owner's HBM)
    int32_t l2_ahead;         // rows beyond the ring that are prefetched into L2 (0: none; at most BD2_L2AHEAD)
    int32_t sync_every;       // 0, or a power of two: the CTA's warps re-align every so many rows of the hot loop, which keeps the
                              // sectors two neighbouring warps share (a warp's 256-byte row segment is not sector-aligned) in L1
    int64_t off_src, off_g, off_f;    // offsets of src / g / f inside the basis allocation (peer addressing)
};
// MODE 0: plain product (FIN = 0).  MODE 1: an Arnoldi column.  MODE 2: the extra product (||y||^2).
template <int ORD, int TS, int MODE, int FIN, int PF, int MINB>
__global__ void __launch_bounds__(VEC_THREADS, MINB) k_spmv_bd2(const __grid_constant__ Lattice L, const __grid_constant__ Bd2Args A,
                                                                Reducer rd, SweepCtl* ctl) {
    constexpr int NS = FIN ? 3 : (MODE == 1 ? 2 : 1);      // streams in flight: src [, g [, f]]
    pdl_trigger();
    pdl_wait();
    if ((MODE != 0 || FIN) && ctl->brk != 0) return;
    const bool has_g = (FIN || MODE == 1) && A.has_g;       // FIN implies has_g (a finalised column has a predecessor)
    const bool has_f = FIN && A.has_f;
    const double sg = has_g ? col_scale(ctl, A.cg) : 0.0;
    const double sf = has_f ? col_scale(ctl, A.cf) : 0.0;
    const double h1 = FIN ? ctl->scal[SC_H1] : 0.0, h2 = FIN ? ctl->scal[SC_H2] : 0.0;
    const double* __restrict__ src = A.src;
    const double* __restrict__ gp = A.g;
    const double* __restrict__ fp_ = A.f;
    double* __restrict__ xout = A.xout;
    double* __restrict__ y = A.y;
    const int32_t Bx = L.B[0], nz = L.nz;
    const int zc = A.zc;
    const int nzl = L.zhi - L.zlo;
    const int64_t ncb = (Bx + BD2_CBW - 1) / BD2_CBW;
    const int64_t nzc = (nzl + zc - 1) / zc;
    const DistPeers* __restrict__ dp = rd.peers;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // y-tables of the chunk, staged once per work item: ztab[k][j] = T_k[z0 - 1 + j].  (Read per row straight from global
    // memory they missed the streaming-thrashed L1 half of the time, profiles/r1_summary.md.)
    __shared__ double ztab[4][BD2_ZT + 2];
    DD accx, accA, accB, accC;
    accx.hi = accx.lo = accA.hi = accA.lo = accB.hi = accB.lo = accC.hi = accC.lo = 0.0;
    for (int64_t item = blockIdx.x; item < ncb * nzc; item += gridDim.x) {
        const int32_t z0 = L.zlo + (int32_t)(item / ncb) * zc;
        const int32_t z1 = min(z0 + zc, L.zhi);
        {
            const int32_t cnt = z1 - z0 + 2;
            __syncthreads();
            for (int q = threadIdx.x; q < 4 * cnt; q += VEC_THREADS) {
                const int k = q / cnt, j = q - k * cnt;
                const int32_t zz = z0 - 1 + j;
                ztab[k][j] = (((TS >> k) & 1) && zz >= 0 && zz < nz) ? __ldg(L.tab[k] + zz) : 0.0;
            }
            __syncthreads();
        }
        // lanes 1..30 own a column, lanes 0 / 31 shadow the column to the left / right (clamped into the box): every
        // lane walks z and keeps x(c, z) in a register, so x(c-1, z) and x(c+1, z) are one shuffle away
        const int32_t c_raw = (int32_t)((item % ncb) * BD2_CBW) + warp * BD2_WCOLS + lane - 1;
        const bool live = lane >= 1 && lane <= BD2_WCOLS && c_raw < Bx;
        const int32_t c = c_raw < 0 ? 0 : (c_raw >= Bx ? Bx - 1 : c_raw);
        const bool okl = c >= 1, okr = c + 1 < Bx;
        // per reaction: a_k(x) for the diagonal and a_k(x - nu_k) for the off-diagonal term.  Tables over x are constant
        // along the walk; tables over y come from the staged chunk (ztab).
        double adc[4], acn[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int dir = bd2_dir(ORD, k);
            const bool tz = (TS >> k) & 1;
            adc[k] = 0.0; acn[k] = 0.0;
            if (!tz) {
                adc[k] = __ldg(L.tab[k] + c);
                acn[k] = dir == BD_XP ? (okl ? __ldg(L.tab[k] + c - 1) : 0.0) : dir == BD_XM ? (okr ? __ldg(L.tab[k] + c + 1) : 0.0) : adc[k];
            }
        }
        // raw operands of row zz (0 <= zz < nz) of this lane's column: from this rank's slab, or (several GPUs) from the
        // owner's basis over NVLink -- the owner's scratch column and basis vectors are complete: the reduction exchange of
        // the kernel that wrote them has passed on every rank (krylov.cuh, grid_reduce)
        auto load_raw = [&](int32_t zz, double& t, double& gr, double& fr) {
            gr = 0.0; fr = 0.0;
            if (!A.halo || (zz >= L.zlo && zz < L.zhi)) {
                const uint32_t e = (uint32_t)(c + Bx * (zz - L.zlo));
                t = src[e];
                if (NS >= 2 && has_g) gr = gp[e];
                if (NS >= 3 && has_f) fr = fp_[e];
            } else {
                int r = 0;
                while (zz >= L.zb[r + 1]) ++r;
                const int64_t e = c + (int64_t)Bx * (zz - L.zb[r]);
                t = __ldcg(dp->V[r] + A.off_src + e);
                if (NS >= 2 && has_g) gr = __ldcg(dp->V[r] + A.off_g + e);
                if (NS >= 3 && has_f) fr = __ldcg(dp->V[r] + A.off_f + e);
            }
        };
        // x = the operand (finalised if FIN), gv = v_g = sg * g
        auto finish = [&](double t, double gr, double fr, double& xv, double& gv) {
            gv = (NS >= 2) ? __dmul_rn(sg, gr) : 0.0;
            if (FIN) {
                double inner = __dmul_rn(sg, t);
                if (has_f) inner = fma(-h1, __dmul_rn(sf, fr), inner);
                xv = fma(-h2, gv, inner);
            } else {
                xv = t;
            }
        };
        const int32_t zlast = min(z1, nz - 1);                  // last row whose operand this work item needs
        uint32_t i = (uint32_t)(c + Bx * (z0 - L.zlo));
        double xm = 0.0, x0, g0;
        {
            double t, gr, fr, gdummy;
            if (z0 >= 1) { load_raw(z0 - 1, t, gr, fr); finish(t, gr, fr, xm, gdummy); }
            load_raw(z0, t, gr, fr);
            finish(t, gr, fr, x0, g0);
        }
        // software pipeline: slot u holds the raw operands of row z+1 of the row z that will be evaluated DEPTH rows after
        // the slot was filled, so a row never waits for its own first-touch loads
        const int32_t zloc = min(zlast, L.zhi - 1);             // ... and the last such row that is local
        constexpr int DEPTH = BD2_DEPTH;
        constexpr int SLOT = NS * VEC_THREADS;                  // doubles per ring slot
        __shared__ double ring[DEPTH * SLOT];
        double* const rbase = ring + threadIdx.x;
        int rs = 0;                                             // slot of the row about to be consumed
        cp_async_wait<0>();                                     // (nothing of the previous work item is still landing)
        auto fill_async = [&](double* slot, uint32_t e) {
            cp_async8(slot, src + e);
            if (NS >= 2 && has_g) cp_async8(slot + VEC_THREADS, gp + e);
            if (NS >= 3 && has_f) cp_async8(slot + 2 * VEC_THREADS, fp_ + e);
        };
        auto fill_general = [&](double* slot, int32_t zn) {
            double t = 0.0, gr = 0.0, fr = 0.0;
            if (zn <= zlast) load_raw(zn, t, gr, fr);
            slot[0] = t;
            if (NS >= 2) slot[VEC_THREADS] = gr;
            if (NS >= 3) slot[2 * VEC_THREADS] = fr;
        };
#pragma unroll
        for (int u = 0; u < DEPTH; ++u) {
            const int32_t zn = z0 + 1 + u;
            if (zn <= zloc) fill_async(rbase + u * SLOT, i + (uint32_t)(Bx * (1 + u)));
            else fill_general(rbase + u * SLOT, zn);
            cp_async_commit();
        }
        // one row: gen = false is the hot path (z >= 1, z + 1 < nz, the refill row is local and needed)
        auto do_row = [&](const int32_t z, const bool gen) {
            double xp, gn;
            const bool yp_ok = gen ? z >= 1 : true, ym_ok = gen ? z + 1 < nz : true;
            const int32_t zn = z + 1 + DEPTH;                   // the row the freed slot is refilled with
            {
                double* slot = rbase + rs * SLOT;
                cp_async_wait<DEPTH - 1>();                     // this thread's copy of row z+1 has landed
                finish(slot[0], NS >= 2 ? slot[VEC_THREADS] : 0.0, NS >= 3 ? slot[2 * VEC_THREADS] : 0.0, xp, gn);
                const uint32_t e = i + (uint32_t)(Bx * (1 + DEPTH));
                if (!gen || zn <= zloc) fill_async(slot, e);
                else fill_general(slot, zn);
                cp_async_commit();
                rs = rs + 1 == DEPTH ? 0 : rs + 1;
                if (!gen && NS >= 2 && A.l2_ahead > 0 && zn + A.l2_ahead <= zloc) {  // multi-stream variants also pull rows further ahead into L2
                    const uint32_t ea = e + (uint32_t)(Bx * A.l2_ahead);
                    lattice_prefetch<1>(src + ea);
                    if (has_g) lattice_prefetch<1>(gp + ea);
                    if (NS >= 3 && has_f) lattice_prefetch<1>(fp_ + ea);
                }
            }
            const double xl = __shfl_up_sync(0xffffffffu, x0, 1);
            const double xr = __shfl_down_sync(0xffffffffu, x0, 1);
            double ad[4], ac[4];
            const int jz = z - z0 + 1;                          // row z in the staged y-tables
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int dir = bd2_dir(ORD, k);
                const bool tz = (TS >> k) & 1;
                if (tz) {
                    ad[k] = ztab[k][jz];
                    ac[k] = dir == BD_YP ? ztab[k][jz - 1] : dir == BD_YM ? ztab[k][jz + 1] : ad[k];
                } else {
                    ad[k] = adc[k];
                    ac[k] = acn[k];
                }
            }
            double d = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) d = __dadd_rn(d, ad[k]);
            double sv = -__dmul_rn(d, x0);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int dir = bd2_dir(ORD, k);
                const bool ok = dir == BD_XP ? okl : dir == BD_XM ? okr : dir == BD_YP ? yp_ok : ym_ok;
                const double xv = dir == BD_XP ? xl : dir == BD_XM ? xr : dir == BD_YP ? xm : xp;
                if (ok) sv = fma(ac[k], xv, sv);
            }
            if (live) {
                if (FIN) { __stcs(xout + i, x0); dd_add_prod(accx, x0, x0); }
                __stcs(y + i, sv);
                if (MODE == 1) {
                    dd_add_prod(accB, x0, sv);
                    if (has_g) { dd_add_prod(accA, g0, sv); dd_add_prod(accC, x0, g0); }
                }
                if (MODE == 2) dd_add_prod(accB, sv, sv);
            }
            xm = x0; x0 = xp; g0 = gn;
            i += (uint32_t)Bx;
        };
        // the hot path runs in whole groups of PF rows: they need z >= 1 and a refill row z + 1 + DEPTH <= zloc (local, needed)
        int32_t z = z0;
        if (z0 == 0) {                                          // the box boundary row goes through the general path
#pragma unroll
            for (int u = 0; u < PF; ++u)
                if (z + u < z1) do_row(z + u, true);
            z += PF;
        }
        for (; z + PF - 1 + 1 + DEPTH <= zloc && z + PF <= z1; z += PF) {
            if (A.sync_every && ((z - z0) & (A.sync_every - 1)) == 0) __syncthreads();
#pragma unroll
            for (int u = 0; u < PF; ++u) do_row(z + u, false);
        }
        for (; z < z1; z += PF) {
#pragma unroll
            for (int u = 0; u < PF; ++u)
                if (z + u < z1) do_row(z + u, true);           // uniform over the CTA: the shuffles are convergent
        }
    }
    if (MODE == 0) return;
    if (MODE == 1) {
        DD v[4] = {accx, accA, accB, accC};
        double tot[4];
        if (grid_reduce<4>(v, tot, rd) && threadIdx.x == 0) reduce_epilogue(A.ea, tot, ctl);
    } else {
        DD v[2] = {accx, accB};
        double tot[2];
        if (grid_reduce<2>(v, tot, rd) && threadIdx.x == 0) reduce_epilogue(A.ea, tot, ctl);
    }
}

// The reference's column form ADJ/OFFDIAG/DIAG (StateSpace.f90:13-17) of local rows [0, count), computed from
// the lattice (for kfsp_fsp_get and the parity tests); adj/offdiag are [i*R + k], Fortran conventions.
__global__ void k_box_export(const __grid_constant__ Lattice L, int64_t g0, int64_t count, int32_t* adj, double* offdiag, double* diag) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
        int32_t st[KFSP_MAX_SPECIES];
        lattice_decode(L, g0 + t, st);
        double d = 0.0;
        for (int k = 0; k < L.R; ++k) {
            const double a = __ldg(L.tab[k] + st[L.sp[k]]);
            d = __dadd_rn(d, a);
            if (offdiag) offdiag[t * L.R + k] = a;
            if (adj) {
                bool neg = false, out = false;
                int64_t o = 0;
                for (int s = 0; s < L.S; ++s) {
                    const int32_t v = st[s] + L.nu[k][s];
                    neg = neg || v < 0;
                    out = out || v >= L.B[s];
                    o += (int64_t)L.nu[k][s] * L.stride[s];
                }
                adj[t * L.R + k] = neg ? -1 : out ? 0 : (int32_t)(g0 + t + o + 1);
            }
        }
        if (diag) diag[t] = d;
    }
}

// FSP%INDEX / FSP%PROBABILITY on the lattice (1-based global index, 0 = not in the projection); w holds the
// local rows starting at global row g0
__global__ void k_box_lookup(const __grid_constant__ Lattice L, const int32_t* __restrict__ q, int64_t nq, int32_t* idx1,
                             const double* __restrict__ w, int64_t g0, int64_t nloc, double* p) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nq; t += (int64_t)gridDim.x * blockDim.x) {
        bool in = true;
        int64_t g = 0;
        for (int s = 0; s < L.S; ++s) {
            const int32_t v = q[t * L.S + s];
            in = in && v >= 0 && v < L.B[s];
            g += (int64_t)v * L.stride[s];
        }
        if (idx1) idx1[t] = in ? (int32_t)(g + 1) : 0;
        if (p) p[t] = (in && g >= g0 && g < g0 + nloc) ? w[g - g0] : 0.0;
    }
}

}  // namespace kfsp

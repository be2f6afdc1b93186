// N-sized kernels of the expv loop (src/fsp/KrylovSolver.f90:223-266, 444-450, 577-607).
// Every one of them is HBM-bound (<= 0.25 flop/byte): the work is to stream the operands
// exactly once, coalesced, and to keep every scalar (dot products, norms, the happy-
// breakdown flag) on the device so that a whole Arnoldi sweep needs no host round trip.
//
// Reductions are deterministic: fixed grid, fixed per-thread element order, per-block
// partials written to global memory and summed in block order by the last block to finish.
#pragma once
#include "common.cuh"

namespace kfsp {

constexpr int VEC_THREADS = 256;
constexpr int MAX_VEC_BLOCKS = 148 * 8;       // 148 SMs x 8 resident 256-thread CTAs

// device scalars of the sweep
enum Scal { SC_H1 = 0, SC_H2, SC_INV_HN, SC_AVNORM, SC_WSUM, SC_WSSQ, SC_HN, SC_COUNT };   // SweepCtl::scal has 8 entries

constexpr int MAX_COLS = 104;      // m_max + 2 basis columns
constexpr int RED_NV = 4;          // a reduction point carries up to 4 double-double values
constexpr int RED_W = 2 * RED_NV;  // ... = 8 doubles per rank
// DSCAL(N, 1/HJ1J, v) (KrylovSolver.f90:258) is never run as a pass of its own: the basis columns stay
// un-normalised in HBM (U_j) and colscale[j] = 1/HJ1J is applied where a column is consumed:
// v_j(i) = __dmul_rn(colscale[j], U_j(i)) is the value DSCAL would have stored.  The generator product is taken on the
// UN-NORMALISED column, Y' = A U_c, and the pass that forms it also accumulates (double-double)
//     dA = <v_{c-1}, Y'>,   dB = <U_c, Y'>,   dC = <U_c, v_{c-1}>      (and ||U_c||^2 where U_c is finalised on the fly)
// from which BOTH coefficients of the IOP window follow at one reduction point, with cs = colscale[c]:
//     H(J-1,J) = h1 = cs*dA,      H(J,J) = h2 = cs*(cs*dB) - h1*(cs*dC)      ( = <v_J, A v_J - h1 v_{J-1}> by linearity )
// and the next column is U_{c+1}(i) = fma(-h2, cs*U_c(i), fma(-h1, v_{c-1}(i), cs*Y'(i))).  One pass over HBM and one
// (cross-GPU) reduction per Arnoldi column instead of three.  This is the canonical arithmetic of
// oracle/kfsp_oracle.cpp (canonical_sweep), bit for bit; against the reference's operation order (two DDOTs on the
// successively updated vector, KrylovSolver.f90:240-245) H agrees to rounding level (tests/test_gpu_kernels.py).
struct SweepCtl {
    double scal[8];
    int32_t brk;                // happy-breakdown column (1-based), 0 = none
    int32_t pad;
    double colscale[MAX_COLS];
};
__device__ __forceinline__ double col_scale(const SweepCtl* ctl, int c) { return c >= 0 ? ctl->colscale[c] : 1.0; }

// Peer memory of the other GPUs of the box (cudaIpc-mapped over NVLink/NVSwitch), multi-GPU runs only.
constexpr int MAX_RANKS = 8;
struct DistPeers {
    int rank, nranks;
    double* part[MAX_RANKS];                 // rank r's exchange area: part[r][(slot*nranks + src)*RED_W + v]
    unsigned long long* flag[MAX_RANKS];     // rank r's arrival flags: flag[r][slot*nranks + src] = sequence number
    const double* V[MAX_RANKS];              // rank r's Krylov basis (same leading dimension on every rank)
    const int32_t* halo_owner;               // for halo position h: owning rank ...
    const int32_t* halo_lidx;                // ... and row index on the owner
    int32_t* err;
    int64_t rb[MAX_RANKS + 1];               // replicated-layout partition (adaptive state sets): rank r computes rows [rb[r], rb[r+1])
    int ll;                                  // 1: partials travel as 8-byte {half, tag} words (no fence between data and flag), 0: data, fence, flag
    unsigned long long* stat;                // [0] fused exchanges done, [1] ns between posting this rank's partial and holding all ranks'
                                             // (wire latency + waiting for the slowest rank), [2] the largest such wait
};
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct Reducer {
    double* partials;           // [RED_W * MAX_VEC_BLOCKS]: (hi, lo) planes for up to RED_NV reductions
    unsigned int* counter;      // self-resetting ticket
    const DistPeers* peers;     // multi-GPU, peer-memory path: partials are exchanged INSIDE the reducing kernel
    unsigned long long seq;     // sequence number of this reduction (same on every rank)
    double* dist_send;          // multi-GPU: this rank's double-double totals go here (RED_W doubles) and the
                                // epilogue runs in k_dist_finalize after the all-gather; nullptr on one GPU
};

// Double-double accumulator: every N-element reduction is carried in (hi, lo) and rounded once
// at the very end, so the rounded value does not depend on the order in which threads, warps and
// blocks combine their parts (the canonical arithmetic of oracle/kfsp_oracle.cpp).  The kernels
// stay HBM-bound: ~12 flops per element against 8-24 bytes of traffic.
struct DD { double hi, lo; };
__device__ __forceinline__ void dd_add_prod(DD& s, double a, double b) {
    const double p = __dmul_rn(a, b);
    const double e = fma(a, b, -p);
    const double t = __dadd_rn(s.hi, p);
    const double z = __dsub_rn(t, s.hi);
    const double err = __dadd_rn(__dsub_rn(s.hi, __dsub_rn(t, z)), __dsub_rn(p, z));
    s.hi = t;
    s.lo = __dadd_rn(s.lo, __dadd_rn(err, e));
}
__device__ __forceinline__ void dd_add(DD& s, double p) {
    const double t = __dadd_rn(s.hi, p);
    const double z = __dsub_rn(t, s.hi);
    const double err = __dadd_rn(__dsub_rn(s.hi, __dsub_rn(t, z)), __dsub_rn(p, z));
    s.hi = t;
    s.lo = __dadd_rn(s.lo, err);
}
__device__ __forceinline__ void dd_merge(DD& s, const DD& o) {
    dd_add(s, o.hi);
    s.lo = __dadd_rn(s.lo, o.lo);
}
__device__ __forceinline__ DD warp_sum(DD v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        DD w;
        w.hi = __shfl_down_sync(0xffffffffu, v.hi, o);
        w.lo = __shfl_down_sync(0xffffffffu, v.lo, o);
        dd_merge(v, w);
    }
    return v;
}
// Sum over the block; result valid in thread 0.
__device__ __forceinline__ DD block_sum(DD v, DD* sh /*>= 32*/) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    if (wid == 0) {
        DD z; z.hi = 0.0; z.lo = 0.0;
        v = lane < (blockDim.x >> 5) ? sh[lane] : z;
        v = warp_sum(v);
    }
    return v;
}
// Programmatic dependent launch (sm_90+): a kernel launched with the programmatic-stream-serialization attribute may be
// scheduled while its predecessor in the stream is still running (its CTAs take the SM slots the predecessor's CTAs free);
// pdl_wait() blocks until the predecessor grid has completed and its writes are visible, so everything that reads a
// predecessor's output comes after it; pdl_trigger() (issued by every CTA of the predecessor, at its start) allows the
// dependent launch.  What this hides is the launch latency and the ramp of the next kernel behind the tail of the current
// one -- on several GPUs that tail is the cross-GPU reduction exchange.  Both are no-ops in a normal launch.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Grid-wide reduction of NV double-double values.  Returns true in ALL threads of the last block to
// finish, with the totals rounded once to double in out[].
template <int NV>
__device__ __forceinline__ bool grid_reduce(const DD (&v)[NV], double (&out)[NV], const Reducer& rd) {
    __shared__ DD sh[NV][32];
    __shared__ bool last;
    __shared__ double tot[NV];
    __shared__ double mine[2 * NV];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    // block totals of all NV values with one pair of barriers
    {
        DD w[NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) w[q] = warp_sum(v[q]);
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < NV; ++q) sh[q][wid] = w[q];
        }
        __syncthreads();
        if (wid == 0) {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                DD z; z.hi = 0.0; z.lo = 0.0;
                DD t = lane < nw ? sh[q][lane] : z;
                t = warp_sum(t);
                if (lane == 0) {
                    rd.partials[(2 * q) * MAX_VEC_BLOCKS + blockIdx.x] = t.hi;
                    rd.partials[(2 * q + 1) * MAX_VEC_BLOCKS + blockIdx.x] = t.lo;
                }
            }
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int ticket = atomicInc(rd.counter, gridDim.x - 1);
        last = ticket == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return false;
    __threadfence();
    {   // the last block merges the per-block partials (block order does not matter: double-double, rounded once at the end)
        DD s[NV];
#pragma unroll
        for (int q = 0; q < NV; ++q) { s[q].hi = 0.0; s[q].lo = 0.0; }
        for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                DD o;
                o.hi = __ldcg(&rd.partials[(2 * q) * MAX_VEC_BLOCKS + b]);
                o.lo = __ldcg(&rd.partials[(2 * q + 1) * MAX_VEC_BLOCKS + b]);
                dd_merge(s[q], o);
            }
        }
#pragma unroll
        for (int q = 0; q < NV; ++q) s[q] = warp_sum(s[q]);
        __syncthreads();
        if (lane == 0) {
#pragma unroll
            for (int q = 0; q < NV; ++q) sh[q][wid] = s[q];
        }
        __syncthreads();
        if (wid == 0) {
#pragma unroll
            for (int q = 0; q < NV; ++q) {
                DD z; z.hi = 0.0; z.lo = 0.0;
                DD t = lane < nw ? sh[q][lane] : z;
                t = warp_sum(t);
                if (lane == 0) {
                    if (rd.peers && rd.seq) { mine[2 * q] = t.hi; mine[2 * q + 1] = t.lo; }
                    else if (rd.dist_send) { rd.dist_send[2 * q] = t.hi; rd.dist_send[2 * q + 1] = t.lo; }
                    else tot[q] = __dadd_rn(t.hi, t.lo);
                }
            }
        }
    }
    if (rd.peers && rd.seq) {
        // Fused all-gather over peer memory: thread r stores this rank's (hi,lo) partials into rank r's exchange
        // area, fences system-wide and raises the arrival flag; then the block waits for every rank's partial of
        // THIS sequence number, merges them in rank order and rounds once.  Only this one block is still running
        // on each GPU, and every GPU runs the same kernel, so the wait cannot deadlock on a healthy box.
        const DistPeers* __restrict__ dp = rd.peers;
        const int P = dp->nranks, me = dp->rank;
        const int slot = (int)(rd.seq & 1ull);
        __syncthreads();
        const unsigned long long ns0 = threadIdx.x == 0 ? global_ns() : 0ull;
        if (dp->ll) {
            // Low-latency form: every 4-byte half of this rank's (hi,lo) partials is stored to every peer as ONE 8-byte word
            // {half, tag} (tag = this reduction's sequence number): an 8-byte store arrives whole, so the word is its own
            // arrival flag and nothing waits for a fence between data and flag -- one one-way NVLink latency instead of
            // store + system fence (a round trip) + flag store.  The system fence BEFORE the words orders everything this
            // GPU wrote in this kernel (the columns the peers' next launch reads over NVLink) ahead of them.
            // Word w of source r: ll[(slot*P + r)*16 + w], w = 4*q + 2*(lo?) + (upper half?), in the exchange area at +1 MiB.
            __shared__ unsigned int recv[MAX_RANKS][4 * RED_NV];
            const unsigned int tag = (unsigned int)(rd.seq & 0x7fffffffull) + 1u;
            if (threadIdx.x == 0) __threadfence_system();
            __syncthreads();
            for (int t = threadIdx.x; t < 16 * P; t += blockDim.x) {
                const int r = t >> 4, w = t & 15;
                if (w >= 4 * NV) continue;
                const double val = mine[w >> 1];
                const unsigned int half = (w & 1) ? (unsigned int)__double2hiint(val) : (unsigned int)__double2loint(val);
                unsigned long long* dst = (unsigned long long*)((char*)dp->part[r] + (1 << 20)) + ((size_t)slot * P + me) * 16 + w;
                const unsigned long long word = ((unsigned long long)tag << 32) | (unsigned long long)half;
                asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst), "l"(word) : "memory");
                const unsigned long long* src = (const unsigned long long*)((const char*)dp->part[me] + (1 << 20)) + ((size_t)slot * P + r) * 16 + w;
                const long long t0 = clock64();
                const volatile int32_t* gone = dp->err;
                unsigned long long got = 0ull;
                for (unsigned int spin = 0;; ++spin) {
                    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(got) : "l"(src) : "memory");
                    if ((unsigned int)(got >> 32) == tag) break;
                    if ((spin & 63u) != 63u) continue;                                        // the slow checks every 64th poll only
                    if (*gone & 32) break;                                                    // a peer timed out earlier: do not spin again
                    if (clock64() - t0 > 8000000000LL) { atomicOr(dp->err, 32); break; }      // ~4 s: a rank is gone
                }
                recv[r][w] = (unsigned int)got;
            }
            __syncthreads();
            if (threadIdx.x == 0) __threadfence_system();
            if (threadIdx.x == 0 && dp->stat) {
                const unsigned long long dt = global_ns() - ns0;
                dp->stat[0] += 1ull;
                dp->stat[1] += dt;
                if (dt > dp->stat[2]) dp->stat[2] = dt;
            }
            if ((int)threadIdx.x < NV) {                        // one thread per value merges the P ranks' partials in rank order
                const int q = threadIdx.x;
                DD s2; s2.hi = 0.0; s2.lo = 0.0;
                for (int r2 = 0; r2 < P; ++r2) {
                    DD o;
                    o.hi = __hiloint2double((int)recv[r2][4 * q + 1], (int)recv[r2][4 * q]);
                    o.lo = __hiloint2double((int)recv[r2][4 * q + 3], (int)recv[r2][4 * q + 2]);
                    dd_merge(s2, o);
                }
                tot[q] = __dadd_rn(s2.hi, s2.lo);
            }
        } else {
        if ((int)threadIdx.x < P) {
            const int r = threadIdx.x;
            double* dst = dp->part[r] + ((size_t)slot * P + me) * RED_W;
#pragma unroll
            for (int v2 = 0; v2 < 2 * NV; ++v2) dst[v2] = mine[v2];
            __threadfence_system();
            *((volatile unsigned long long*)(dp->flag[r] + (size_t)slot * P + me)) = rd.seq;
            const volatile unsigned long long* fl = dp->flag[me] + (size_t)slot * P + r;
            const long long t0 = clock64();
            const volatile int32_t* gone = dp->err;
            while (*fl != rd.seq) {
                if (*gone & 32) break;                                                    // a peer timed out earlier: do not spin again
                if (clock64() - t0 > 8000000000LL) { atomicOr(dp->err, 32); break; }      // ~4 s: a rank is gone
            }
            __threadfence_system();
        }
        __syncthreads();
        if (threadIdx.x == 0 && dp->stat) {
            const unsigned long long dt = global_ns() - ns0;
            dp->stat[0] += 1ull;
            dp->stat[1] += dt;
            if (dt > dp->stat[2]) dp->stat[2] = dt;
        }
        if ((int)threadIdx.x < NV) {                        // one thread per value merges the P ranks' partials in rank order
            const int q = threadIdx.x;
            DD s; s.hi = 0.0; s.lo = 0.0;
            for (int r = 0; r < P; ++r) {
                const volatile double* src = dp->part[me] + ((size_t)slot * P + r) * RED_W;
                DD o; o.hi = src[2 * q]; o.lo = src[2 * q + 1];
                dd_merge(s, o);
            }
            tot[q] = __dadd_rn(s.hi, s.lo);
        }
        }   // fence / flag form
    } else if (rd.dist_send) {
        return false;                        // uniform: totals are combined across ranks by NCCL first
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; ++q) out[q] = tot[q];
    return true;
}

// What to do with a finished reduction (same code on one GPU, inside the reducing kernel, and on
// several GPUs, in k_dist_finalize after the ranks' double-double partials were gathered).
enum RKind { RK_COLUMN = 0, RK_EXTRA, RK_FIN_NRM, RK_NORMS, RK_SUM_BELOW };
struct EpiArgs {
    int kind;
    int column;        // the SpMV's operand column c (0-based; -1: plain operand, scale 1).  c is also the 1-based Arnoldi index of
                       // the step that produced U_c, i.e. the happy-breakdown column if ||U_c|| is tiny (KrylovSolver.f90:249-256)
    int fin;           // tot[0] = ||U_c||^2 was accumulated by this launch (the column was finalised in its load stage)
    int has_g;         // c >= 1: the window has a previous vector (tot[1] = dA, tot[3] = dC)
    double break_tol;
    double* h1_out;    // H(J-1,J)
    double* h2_out;    // H(J,J)
    double* hn_out;    // H(c+1,c) = ||U_c||
};
__device__ __forceinline__ bool epilogue_norm(double ssq, SweepCtl* ctl, double* hn_out, double break_tol, int column) {
    const double hn = sqrt(ssq);                            // HJ1J (KrylovSolver.f90:247)
    ctl->scal[SC_HN] = hn;
    if (hn <= break_tol) {
        ctl->brk = column;                                  // happy breakdown (:249-256)
        return false;
    }
    if (hn_out) *hn_out = hn;
    const double inv = 1.0 / hn;
    ctl->scal[SC_INV_HN] = inv;
    ctl->colscale[column] = inv;                            // column `column` of V stays un-normalised
    return true;
}
// tot: RK_COLUMN {ssq, dA, dB, dC}; RK_EXTRA {ssq, ||Y'||^2}; RK_FIN_NRM {ssq}; RK_NORMS {||w||_1, ||w||^2}; RK_SUM_BELOW {sum}
__device__ __forceinline__ void reduce_epilogue(const EpiArgs& ea, const double* tot, SweepCtl* ctl) {
    switch (ea.kind) {
    case RK_COLUMN: {
        if (ea.fin && !epilogue_norm(tot[0], ctl, ea.hn_out, ea.break_tol, ea.column)) break;
        const double sc = col_scale(ctl, ea.column);
        double h1 = 0.0, h2 = __dmul_rn(sc, __dmul_rn(sc, tot[2]));
        if (ea.has_g) {
            h1 = __dmul_rn(sc, tot[1]);
            h2 = fma(-h1, __dmul_rn(sc, tot[3]), h2);
            if (ea.h1_out) *ea.h1_out = h1;
        }
        ctl->scal[SC_H1] = h1;
        ctl->scal[SC_H2] = h2;
        if (ea.h2_out) *ea.h2_out = h2;
    } break;
    case RK_EXTRA:
        if (ea.fin && !epilogue_norm(tot[0], ctl, ea.hn_out, ea.break_tol, ea.column)) break;
        ctl->scal[SC_AVNORM] = __dmul_rn(col_scale(ctl, ea.column), sqrt(tot[1]));
        break;
    case RK_FIN_NRM: epilogue_norm(tot[0], ctl, ea.hn_out, ea.break_tol, ea.column); break;
    case RK_NORMS: ctl->scal[SC_WSUM] = tot[0]; ctl->scal[SC_WSSQ] = tot[1]; break;
    case RK_SUM_BELOW: ctl->scal[SC_WSUM] = tot[0]; break;
    }
}
__device__ __forceinline__ EpiArgs epi_simple(int kind) {
    EpiArgs e;
    e.kind = kind; e.column = -1; e.fin = 0; e.has_g = 0; e.break_tol = 0.0; e.h1_out = e.h2_out = e.hn_out = nullptr;
    return e;
}
// multi-GPU, NCCL path: merge the P gathered (hi,lo) partials in rank order, round once, run the epilogue
__global__ void k_dist_finalize(EpiArgs ea, int nv, const double* __restrict__ recv /*[P][RED_W]*/, int nranks, SweepCtl* ctl) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (ea.kind != RK_NORMS && ea.kind != RK_SUM_BELOW && ctl->brk != 0) return;
    double tot[RED_NV] = {0.0, 0.0, 0.0, 0.0};
    for (int q = 0; q < nv; ++q) {
        DD s; s.hi = 0.0; s.lo = 0.0;
        for (int r = 0; r < nranks; ++r) {
            DD o; o.hi = recv[r * RED_W + 2 * q]; o.lo = recv[r * RED_W + 2 * q + 1];
            dd_merge(s, o);
        }
        tot[q] = __dadd_rn(s.hi, s.lo);
    }
    reduce_epilogue(ea, tot, ctl);
}

// ---------------------------------------------------------------------------------------
// FMATVEC (KrylovSolver.f90:577-607) in gather form: y_i = -DIAG_i x_i + sum_k coef_ki x[pred_ki], on the stored
// (un-normalised) column x = U_c.
//   mode 0: plain
//   mode 1: one Arnoldi column: also dA = <v_{c-1}, y>, dB = <x, y>, dC = <x, v_{c-1}> (g = U_{c-1} is the extra stream;
//           has_g = 0 for the first column); the epilogue turns them into H(J-1,J) and H(J,J) (reduce_epilogue)
//   mode 2: the extra product: ||y||^2 -> AVNORM (:261-263)
// Algorithmic traffic per row: R*(4+8) matrix + 8 diag + 8 x_i + 8 y_i  = 12R+24 bytes (+8 for g).
// ---------------------------------------------------------------------------------------
// HALO 3 / 4: adaptive state sets on several GPUs keep the single-GPU (global) layout on every rank and each rank computes a
// contiguous slice of the rows (Engine::repartition).  x points at the slice's first row; j is a GLOBAL index.
//   3: rows outside the slice come from the owner's copy of the same column, same offset, over NVLink (peer memory)
//   4: the column was completed on this GPU by an all-gather before the launch
template <int HALO>
__device__ __forceinline__ double halo_load(const double* __restrict__ x, const double* __restrict__ xh, const DistPeers* __restrict__ dp,
                                            int32_t j, int64_t nloc, int64_t coloff, int64_t row0) {
    if (HALO == 0) return x[j];
    if (HALO == 4) return x[(int64_t)j - row0];
    if (HALO == 3) {
        const int64_t jj = (int64_t)j - row0;
        if (jj >= 0 && jj < nloc) return x[jj];
        int r = 0;
        while ((int64_t)j >= dp->rb[r + 1]) ++r;
        return __ldcg(dp->V[r] + coloff + jj);
    }
    if (j < nloc) return x[j];
    if (HALO == 1) return xh[j - nloc];
    const int64_t h = j - nloc;
    return __ldcg(dp->V[dp->halo_owner[h]] + coloff + dp->halo_lidx[h]);
}
template <int RT, int MODE, int UNROLL, int MINB, int HALO>
__global__ void __launch_bounds__(VEC_THREADS, MINB) k_spmv(int64_t n, int64_t ld, int R_rt, const int32_t* __restrict__ pred,
                                                             const double* __restrict__ coef, const double* __restrict__ diag,
                                                             const double* __restrict__ x, double* __restrict__ y,
                                                             const double* __restrict__ g, Reducer rd, SweepCtl* ctl, EpiArgs ea,
                                                             int cg, const double* __restrict__ xh, int64_t nloc, int64_t coloff, int64_t row0) {
    // HALO 1: gathered index j >= nloc addresses the halo buffer xh filled by the NCCL exchange step;
    // HALO 2: it is loaded straight from the owning GPU's basis column over NVLink (peer memory).
    const int R = RT > 0 ? RT : R_rt;
    pdl_trigger();
    pdl_wait();
    if (MODE != 0 && ctl->brk != 0) return;
    const bool has_g = MODE == 1 && ea.has_g;
    const double gs = has_g ? col_scale(ctl, cg) : 0.0;
    DD accA, accB, accC;
    accA.hi = accA.lo = accB.hi = accB.lo = accC.hi = accC.lo = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < n; i0 += stride * UNROLL) {
        // UNROLL independent rows per iteration: all streaming loads are issued before any gather is consumed
        double s[UNROLL], gv[UNROLL], xi[UNROLL];
        int32_t j[UNROLL][RT > 0 ? RT : 1];
        double a[UNROLL][RT > 0 ? RT : 1];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n) {
                gv[u] = has_g ? __dmul_rn(gs, __ldcs(g + i)) : 0.0;
                xi[u] = x[i];
                s[u] = -__dmul_rn(__ldcs(diag + i), xi[u]);
                if (RT > 0) {
#pragma unroll
                    for (int k = 0; k < RT; ++k) {
                        j[u][k] = __ldcs(pred + (int64_t)k * ld + i);
                        a[u][k] = __ldcs(coef + (int64_t)k * ld + i);
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < n) {
                double sv = s[u];
                if (RT > 0) {
#pragma unroll
                    for (int k = 0; k < RT; ++k)
                        if (j[u][k] >= 0) sv = fma(a[u][k], halo_load<HALO>(x, xh, rd.peers, j[u][k], nloc, coloff, row0), sv);
                } else {
                    for (int k = 0; k < R; ++k) {
                        const int32_t jj = __ldcs(pred + (int64_t)k * ld + i);
                        const double aa = __ldcs(coef + (int64_t)k * ld + i);
                        if (jj >= 0) sv = fma(aa, halo_load<HALO>(x, xh, rd.peers, jj, nloc, coloff, row0), sv);
                    }
                }
                __stcs(y + i, sv);
                if (MODE == 1) {
                    dd_add_prod(accB, xi[u], sv);
                    if (has_g) { dd_add_prod(accA, gv[u], sv); dd_add_prod(accC, xi[u], gv[u]); }
                }
                if (MODE == 2) dd_add_prod(accB, sv, sv);
            }
        }
    }
    if (MODE == 0) return;
    if (MODE == 1) {
        DD z; z.hi = 0.0; z.lo = 0.0;
        DD v[4] = {z, accA, accB, accC};
        double tot[4];
        if (grid_reduce<4>(v, tot, rd) && threadIdx.x == 0) reduce_epilogue(ea, tot, ctl);
    } else {
        DD z; z.hi = 0.0; z.lo = 0.0;
        DD v[2] = {z, accB};
        double tot[2];
        if (grid_reduce<2>(v, tot, rd) && threadIdx.x == 0) reduce_epilogue(ea, tot, ctl);
    }
}

// ---------------------------------------------------------------------------------------
// FMATVEC without the coefficient array (spmv_variant = 2, "index-only"): the row streams pred (4R bytes) and the integer
// state of the row (4S bytes) and RECOMPUTES  coef_ki = a_k(x_i - nu_k)  from it -- the stored value was produced by the
// same evaluation of the same state (StateSpace.f90:207-212, 240-244: OFFDIAG(K,J) is a_K of state J), so the product is
// bit-identical to k_spmv -- for 4R + 4S + 24 bytes per state instead of 12R + 24 (Goutsias: 88 instead of 144;
// DREC = 1 also recomputes DIAG_i = ((a_1 + a_2) + ...)(x_i): 4R + 4S + 16).  Works on ANY state set (irregular,
// adaptively grown, row-partitioned); propensities are the factored tables of common.cuh (FacModel).
// The thread's state sits in shared memory (column threadIdx.x) because the species a term reads is a run-time index.
// ---------------------------------------------------------------------------------------
template <int ST, int STRIDE>
__device__ __forceinline__ void idx_load_state(const int32_t* __restrict__ states, int64_t row, int S, int32_t* sst) {
    if (ST > 0 && ST % 2 == 0) {
        const int2* __restrict__ p = reinterpret_cast<const int2*>(states + row * ST);
#pragma unroll
        for (int q = 0; q < ST / 2; ++q) {
            const int2 v = __ldcs(p + q);
            sst[(2 * q) * STRIDE] = v.x;
            sst[(2 * q + 1) * STRIDE] = v.y;
        }
    } else {
        const int SS = ST > 0 ? ST : S;
#pragma unroll
        for (int q = 0; q < (ST > 0 ? ST : KFSP_MAX_SPECIES); ++q)
            if (q < SS) sst[q * STRIDE] = __ldcs(states + row * SS + q);
    }
}
// (fac_term / fac_eval: common.cuh)
// RT > 0: reactions unrolled, loads of IDX_GROUP reactions (table entries and gathered x) issued before their FMAs.
// RT == 0: any R, and the only instantiation that carries the postfix interpreter (models with a FAC_GEN reaction).
constexpr int IDX_GROUP = 5;
template <int RT, int ST, int MODE, int HALO, int DREC>
__global__ void __launch_bounds__(VEC_THREADS, MODE == 0 ? 6 : 4) k_spmv_idx(const __grid_constant__ FacModel F, int64_t n, int64_t ld, const int32_t* __restrict__ pred,
                                                          const int32_t* __restrict__ states, const double* __restrict__ diag,
                                                          const double* __restrict__ x, double* __restrict__ y, const double* __restrict__ g,
                                                          Reducer rd, SweepCtl* ctl, EpiArgs ea, int cg, const double* __restrict__ xh,
                                                          int64_t nloc, int64_t coloff, int64_t row0) {
    __shared__ int32_t sstate[KFSP_MAX_SPECIES * VEC_THREADS];
    int32_t* const sst = sstate + threadIdx.x;
    constexpr int GEN = RT == 0 ? 1 : 0;
    const int R = RT > 0 ? RT : F.R;
    pdl_trigger();
    pdl_wait();
    if (MODE != 0 && ctl->brk != 0) return;
    const bool has_g = MODE == 1 && ea.has_g;
    const double gs = has_g ? col_scale(ctl, cg) : 0.0;
    DD accA, accB, accC;
    accA.hi = accA.lo = accB.hi = accB.lo = accC.hi = accC.lo = 0.0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += stride) {
        int32_t j[RT > 0 ? RT : 1];
        if (RT > 0) {
#pragma unroll
            for (int k = 0; k < RT; ++k) j[k] = __ldcs(pred + (int64_t)k * ld + i);
        }
        idx_load_state<ST, VEC_THREADS>(states, i, F.S, sst);
        const double xi = x[i];
        const double gv = has_g ? __dmul_rn(gs, __ldcs(g + i)) : 0.0;
        double d;
        if (DREC) {
            d = 0.0;
            if (RT > 0) {
#pragma unroll
                for (int k = 0; k < (RT > 0 ? RT : 1); ++k) d = __dadd_rn(d, fac_eval<VEC_THREADS, GEN>(F, k, sst, 0));
            } else {
                for (int k = 0; k < R; ++k) d = __dadd_rn(d, fac_eval<VEC_THREADS, GEN>(F, k, sst, 0));
            }
        } else {
            d = __ldcs(diag + i);
        }
        double sv = -__dmul_rn(d, xi);
        if (RT > 0) {
#pragma unroll
            for (int k0 = 0; k0 < RT; k0 += IDX_GROUP) {
                double a[IDX_GROUP], xv[IDX_GROUP];
#pragma unroll
                for (int q = 0; q < IDX_GROUP; ++q) {
                    const int k = k0 + q;
                    if (k < RT) {
                        const bool ok = j[k] >= 0;
                        a[q] = ok ? fac_eval<VEC_THREADS, GEN>(F, k, sst, 1) : 0.0;
                        xv[q] = ok ? halo_load<HALO>(x, xh, rd.peers, j[k], nloc, coloff, row0) : 0.0;
                    }
                }
#pragma unroll
                for (int q = 0; q < IDX_GROUP; ++q) {
                    const int k = k0 + q;
                    if (k < RT) sv = j[k] >= 0 ? fma(a[q], xv[q], sv) : sv;
                }
            }
        } else {
            for (int k = 0; k < R; ++k) {
                const int32_t jj = __ldcs(pred + (int64_t)k * ld + i);
                if (jj >= 0) sv = fma(fac_eval<VEC_THREADS, GEN>(F, k, sst, 1), halo_load<HALO>(x, xh, rd.peers, jj, nloc, coloff, row0), sv);
            }
        }
        __stcs(y + i, sv);
        if (MODE == 1) {
            dd_add_prod(accB, xi, sv);
            if (has_g) { dd_add_prod(accA, gv, sv); dd_add_prod(accC, xi, gv); }
        }
        if (MODE == 2) dd_add_prod(accB, sv, sv);
    }
    if (MODE == 0) return;
    if (MODE == 1) {
        DD z; z.hi = 0.0; z.lo = 0.0;
        DD v[4] = {z, accA, accB, accC};
        double tot[4];
        if (grid_reduce<4>(v, tot, rd) && threadIdx.x == 0) reduce_epilogue(ea, tot, ctl);
    } else {
        DD z; z.hi = 0.0; z.lo = 0.0;
        DD v[2] = {z, accB};
        double tot[2];
        if (grid_reduce<2>(v, tot, rd) && threadIdx.x == 0) reduce_epilogue(ea, tot, ctl);
    }
}

// U_c = (cg*w - h1*v_f) - h2*v_g in place on w = A U_{c-1} (g = U_{c-1}, f = U_{c-2}; has_f = 0 for column 1), and its
// norm: HJ1J = ||U_c||, happy-breakdown test, H(c+1,c), colscale[c] (the two DAXPYs, DNRM2 and the deferred DSCAL of
// KrylovSolver.f90:244-258).  The fused lattice path does this inside the next SpMV launch (lattice.cuh).
__global__ void __launch_bounds__(VEC_THREADS) k_finalize(int64_t n, const double* __restrict__ g, const double* __restrict__ f, double* __restrict__ w,
                                                          int has_f, Reducer rd, SweepCtl* ctl, EpiArgs ea, int cg, int cf) {
    pdl_trigger();
    pdl_wait();
    if (ctl->brk != 0) return;
    const double h1 = ctl->scal[SC_H1], h2 = ctl->scal[SC_H2];
    const double sg = col_scale(ctl, cg), sf = has_f ? col_scale(ctl, cf) : 0.0;
    DD acc; acc.hi = 0.0; acc.lo = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double inner = __dmul_rn(sg, w[i]);
        if (has_f) inner = fma(-h1, __dmul_rn(sf, __ldcs(f + i)), inner);
        const double u = fma(-h2, __dmul_rn(sg, __ldcs(g + i)), inner);
        w[i] = u;
        dd_add_prod(acc, u, u);
    }
    DD v[1] = {acc};
    double tot[1];
    if (grid_reduce<1>(v, tot, rd) && threadIdx.x == 0) reduce_epilogue(ea, tot, ctl);
}

// V(:,1) = (1/BETA) * W (KrylovSolver.f90:223-226)
__global__ void __launch_bounds__(VEC_THREADS) k_scale_copy(int64_t n, double s, const double* __restrict__ w, double* __restrict__ v) {
    pdl_trigger();
    pdl_wait();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) v[i] = __dmul_rn(s, w[i]);
}
// W = BETA * V(:,1) (KrylovSolver.f90:467)
__global__ void __launch_bounds__(VEC_THREADS) k_scale_copy_nrm(int64_t n, double s, const double* __restrict__ v, double* __restrict__ w,
                                                                Reducer rd, SweepCtl* ctl) {
    DD a1, a2; a1.hi = a1.lo = a2.hi = a2.lo = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = __dmul_rn(s, v[i]);
        w[i] = x;
        dd_add(a1, fabs(x));
        dd_add_prod(a2, x, x);
    }
    DD vv[2] = {a1, a2};
    double tot[2];
    if (grid_reduce<2>(vv, tot, rd) && threadIdx.x == 0) reduce_epilogue(epi_simple(RK_NORMS), tot, ctl);
}
// ||w||_1 and ||w||_2^2 of a vector (BETA = DNRM2(N_NOW, W), KrylovSolver.f90:177,540)
__global__ void __launch_bounds__(VEC_THREADS) k_norms(int64_t n, const double* __restrict__ w, Reducer rd, SweepCtl* ctl) {
    DD a1, a2; a1.hi = a1.lo = a2.hi = a2.lo = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = w[i];
        dd_add(a1, fabs(x));
        dd_add_prod(a2, x, x);
    }
    DD vv[2] = {a1, a2};
    double tot[2];
    if (grid_reduce<2>(vv, tot, rd) && threadIdx.x == 0) reduce_epilogue(epi_simple(RK_NORMS), tot, ctl);
}

// W = BETA * V(:,1:mx) * e ; W = max(W,0) ; WSUM = ||W||_1 ; also ||W||_2^2 for the next BETA
// (DGEMV + clamp + DASUM, KrylovSolver.f90:444-450).  Streams 8*N*mx bytes once.
__global__ void __launch_bounds__(VEC_THREADS) k_combine(int64_t n, int64_t ld, int mx, double beta, const double* __restrict__ V,
                                                         const double* __restrict__ e, double* __restrict__ w, Reducer rd, SweepCtl* ctl) {
    __shared__ double coef[128];
    __shared__ double cs[128];
    pdl_trigger();
    pdl_wait();
    for (int j = threadIdx.x; j < mx; j += blockDim.x) { coef[j] = __dmul_rn(beta, e[j]); cs[j] = ctl->colscale[j]; }   // temp = alpha*x(j)
    __syncthreads();
    DD a1, a2; a1.hi = a1.lo = a2.hi = a2.lo = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double s = 0.0;
        int j = 0;
        for (; j + 4 <= mx; j += 4) {
            const double v0 = __ldcs(V + (int64_t)j * ld + i), v1 = __ldcs(V + (int64_t)(j + 1) * ld + i);
            const double v2 = __ldcs(V + (int64_t)(j + 2) * ld + i), v3 = __ldcs(V + (int64_t)(j + 3) * ld + i);
            s = fma(coef[j], __dmul_rn(cs[j], v0), s); s = fma(coef[j + 1], __dmul_rn(cs[j + 1], v1), s);
            s = fma(coef[j + 2], __dmul_rn(cs[j + 2], v2), s); s = fma(coef[j + 3], __dmul_rn(cs[j + 3], v3), s);
        }
        for (; j < mx; ++j) s = fma(coef[j], __dmul_rn(cs[j], __ldcs(V + (int64_t)j * ld + i)), s);
        if (s < 0.0) s = 0.0;
        w[i] = s;
        dd_add(a1, s);
        dd_add_prod(a2, s, s);
    }
    DD vv[2] = {a1, a2};
    double tot[2];
    if (grid_reduce<2>(vv, tot, rd) && threadIdx.x == 0) reduce_epilogue(epi_simple(RK_NORMS), tot, ctl);
}

// FIND_DROPTOL's inner sum (StateSpace.f90:418-423): sum of W_i with 0 < W_i < droptol
__global__ void __launch_bounds__(VEC_THREADS) k_sum_below(int64_t n, const double* __restrict__ w, double droptol, Reducer rd, SweepCtl* ctl) {
    DD a; a.hi = a.lo = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = w[i];
        if (x < droptol && x > 0.0) dd_add(a, x);
    }
    DD vv[1] = {a};
    double tot[1];
    if (grid_reduce<1>(vv, tot, rd) && threadIdx.x == 0) reduce_epilogue(epi_simple(RK_SUM_BELOW), tot, ctl);
}

// ---------------------------------------------------------------------------------------
// Whole Arnoldi/IOP-2 sweep in ONE launch for small state spaces (N of a few thousand: BASELINE configs 1-2).
// At that size a column is ~10 KB of traffic and three kernel launches of pure latency; here one CTA walks
// through columns jold..m and the extra product, with __syncthreads between the phases instead of kernel
// boundaries.  Every element sees exactly the operations of k_spmv / k_axpy_dot / k_axpy_nrm, and the
// reductions are the same double-double sums, so H, the basis and the breakdown flag are bit-identical.
// ---------------------------------------------------------------------------------------
constexpr int SWEEP_THREADS = 1024;
__device__ __forceinline__ double cta_dd_total(DD v, DD* sh, double* bc) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) sh[wid] = v;
    __syncthreads();
    if (wid == 0) {
        DD z; z.hi = 0.0; z.lo = 0.0;
        v = lane < (SWEEP_THREADS >> 5) ? sh[lane] : z;
        v = warp_sum(v);
        if (lane == 0) *bc = __dadd_rn(v.hi, v.lo);
    }
    __syncthreads();
    return *bc;
}
// three totals at once; `all` = false: only b is needed (a and c come back as 0)
__device__ __forceinline__ void cta_dd_total3(DD a, DD b, DD c, bool all, DD (*sh3)[32], double* bc3, double* ta, double* tb, double* tc) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    b = warp_sum(b);
    if (all) { a = warp_sum(a); c = warp_sum(c); }
    if (lane == 0) { sh3[0][wid] = a; sh3[1][wid] = b; sh3[2][wid] = c; }
    __syncthreads();
    if (wid < 3 && (all || wid == 1)) {                   // warps 0, 1, 2 finish one value each
        DD z; z.hi = 0.0; z.lo = 0.0;
        DD v = lane < (SWEEP_THREADS >> 5) ? sh3[wid][lane] : z;
        v = warp_sum(v);
        if (lane == 0) bc3[wid] = __dadd_rn(v.hi, v.lo);
    }
    __syncthreads();
    *ta = all ? bc3[0] : 0.0;
    *tb = bc3[1];
    *tc = all ? bc3[2] : 0.0;
}
template <int RT, int IDX>
__device__ __forceinline__ double spmv_row(int64_t i, int64_t ld, int R, const int32_t* __restrict__ pred, const double* __restrict__ coef,
                                           const double* __restrict__ diag, const double* x, const FacModel& F,
                                           const int32_t* __restrict__ states, int32_t* sst) {
    double sv = -__dmul_rn(diag[i], x[i]);
    if (IDX) idx_load_state<0, SWEEP_THREADS>(states, i, F.S, sst);      // index-only variant: coef is recomputed (k_spmv_idx)
#pragma unroll
    for (int k = 0; k < (RT > 0 ? RT : R); ++k) {
        const int32_t j = pred[(int64_t)k * ld + i];
        if (j >= 0) {
            const double a = IDX ? fac_eval<SWEEP_THREADS, 1>(F, k, sst, 1) : coef[(int64_t)k * ld + i];
            sv = fma(a, x[j], sv);
        }
    }
    return sv;
}
template <int RT, int IDX>
__global__ void __launch_bounds__(SWEEP_THREADS, 1) k_sweep_small(int64_t n, int64_t ld, int R_rt, const int32_t* __restrict__ pred,
                                                                   const double* __restrict__ coef, const double* __restrict__ diag,
                                                                   double* V, double* H, int ldh, int jold, int m, SweepCtl* ctl,
                                                                   double break_tol, const __grid_constant__ FacModel F,
                                                                   const int32_t* __restrict__ states) {
    __shared__ int32_t sstate[IDX ? KFSP_MAX_SPECIES * SWEEP_THREADS : 1];
    int32_t* const sst = sstate + (IDX ? threadIdx.x : 0);
    __shared__ DD sh[32];
    __shared__ DD sh3[3][32];
    __shared__ double bc;
    __shared__ double bc3[3];
    __shared__ double cs[MAX_COLS];
    __shared__ int s_brk;
    const int R = RT > 0 ? RT : R_rt;
    const int tid = threadIdx.x;
    for (int j = tid; j < MAX_COLS; j += SWEEP_THREADS) cs[j] = ctl->colscale[j];
    if (tid == 0) s_brk = ctl->brk;
    __syncthreads();
    if (s_brk != 0) return;
    for (int J = jold; J <= m; ++J) {
        const double* x = V + (size_t)(J - 1) * ld;
        double* y = V + (size_t)J * ld;
        const double xs = cs[J - 1];
        const double* g = J >= 2 ? V + (size_t)(J - 2) * ld : x;
        const double gs = J >= 2 ? cs[J - 2] : 0.0;
        double* hcol = H + (size_t)(J - 1) * ldh;
        // FMATVEC on the un-normalised column and the three inner products of the window
        DD accA, accB, accC;
        accA.hi = accA.lo = accB.hi = accB.lo = accC.hi = accC.lo = 0.0;
        for (int64_t i = tid; i < n; i += SWEEP_THREADS) {
            const double sv = spmv_row<RT, IDX>(i, ld, R, pred, coef, diag, x, F, states, sst);
            y[i] = sv;
            const double xi = x[i];
            dd_add_prod(accB, xi, sv);
            if (J >= 2) {
                const double gv = __dmul_rn(gs, g[i]);
                dd_add_prod(accA, gv, sv);
                dd_add_prod(accC, xi, gv);
            }
        }
        // the three inner products of the column with one pair of barriers (same double-double sums, rounded once each)
        double dA, dB, dC;
        cta_dd_total3(accA, accB, accC, J >= 2, sh3, bc3, &dA, &dB, &dC);
        double h1 = 0.0, h2 = __dmul_rn(xs, __dmul_rn(xs, dB));
        if (J >= 2) {
            h1 = __dmul_rn(xs, dA);
            h2 = fma(-h1, __dmul_rn(xs, dC), h2);
            if (tid == 0) hcol[J - 2] = h1;                                  // H(J-1,J)
        }
        if (tid == 0) hcol[J - 1] = h2;                                      // H(J,J)
        DD acc; acc.hi = 0.0; acc.lo = 0.0;
        for (int64_t i = tid; i < n; i += SWEEP_THREADS) {
            double inner = __dmul_rn(xs, y[i]);
            if (J >= 2) inner = fma(-h1, __dmul_rn(gs, g[i]), inner);
            const double wi = fma(-h2, __dmul_rn(xs, x[i]), inner);
            y[i] = wi;
            dd_add_prod(acc, wi, wi);
        }
        const double hn = sqrt(cta_dd_total(acc, sh, &bc));
        if (hn <= break_tol) {                                               // happy breakdown: uniform over the CTA
            if (tid == 0) { ctl->scal[SC_HN] = hn; ctl->brk = J; }
            s_brk = J;
            break;
        }
        if (tid == 0) { hcol[J] = hn; cs[J] = 1.0 / hn; }                    // H(J+1,J), DSCAL factor
        __syncthreads();
    }
    if (s_brk == 0) {
        const double* x = V + (size_t)m * ld;
        double* y = V + (size_t)(m + 1) * ld;
        const double xs = cs[m];
        DD acc; acc.hi = 0.0; acc.lo = 0.0;
        for (int64_t i = tid; i < n; i += SWEEP_THREADS) {
            const double sv = spmv_row<RT, IDX>(i, ld, R, pred, coef, diag, x, F, states, sst);
            y[i] = sv;
            dd_add_prod(acc, sv, sv);
        }
        const double av = __dmul_rn(xs, sqrt(cta_dd_total(acc, sh, &bc)));
        if (tid == 0) ctl->scal[SC_AVNORM] = av;
    }
    __syncthreads();
    for (int j = tid; j < MAX_COLS; j += SWEEP_THREADS) ctl->colscale[j] = cs[j];
}

// ---------------------------------------------------------------------------------------
// The same sweep for MID-SIZED state sets (1e4 .. a few 1e6 rows: BASELINE configs 2-4, whose generator sits in the 126 MB
// L2) as ONE cooperative launch over all SMs: grid-wide barriers instead of kernel boundaries.  In the multi-launch form a
// column of such a set is two launches whose useful part (a few us of L2 traffic) is shorter than their launch latency,
// ramp and last-block reduction tail (Goutsias at 6e5 rows: 32 us per column, profiles/r2_summary.md).  Here a column is
// two phases separated by two barriers; every reduction is written as per-CTA double-double partials that EVERY CTA
// merges after the barrier (same sums, rounded once: bit-identical to k_sweep_small and to the multi-launch kernels), so
// no value has to be broadcast and the barrier doubles as the fence between a column's writer and its gatherers.
// A column is written in iteration J only and first read by other CTAs after that iteration's last barrier, so L1 never
// holds a stale line of it; the partials (reused every second barrier) are read with ld.cg.
// ---------------------------------------------------------------------------------------
constexpr int COOP_MAXG = 1024;                            // CTAs of a cooperative sweep (one or two per SM)
constexpr int32_t DEV_COOP_TIMEOUT = 64;                   // device error bit (state_space.cuh: DevErr; 32 = a peer GPU stopped responding)
struct CoopBuf {
    unsigned int* bar;                                     // arrival counter, zero at launch
    double* part;                                          // [(buf * 6 + plane) * COOP_MAXG + cta], buf = 0/1, plane = 2*value + (hi|lo)
    int32_t* err;
};
// false: a CTA never arrived (the launch was not co-resident, or a fault): every CTA gives up after ~2 s
__device__ __forceinline__ bool coop_barrier(const CoopBuf& cb, unsigned int& epoch) {
    __shared__ int s_ok;
    __syncthreads();
    epoch += gridDim.x;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(cb.bar, 1u);
        unsigned int v;
        const long long t0 = clock64();
        int ok = 1;
        for (;;) {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(cb.bar) : "memory");
            if (v >= epoch) break;
            if (clock64() - t0 > 4000000000LL) { ok = 0; atomicOr(cb.err, DEV_COOP_TIMEOUT); break; }
        }
        s_ok = ok;
    }
    __syncthreads();
    return s_ok != 0;
}
template <int NV>
__device__ __forceinline__ bool coop_total(const DD (&v)[NV], double (&out)[NV], const CoopBuf& cb, int& buf, unsigned int& epoch,
                                           DD (*sh)[32], double* bc) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double* const P = cb.part + (size_t)buf * 6 * COOP_MAXG;
    DD w[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) w[q] = warp_sum(v[q]);
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) sh[q][wid] = w[q];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            DD z; z.hi = 0.0; z.lo = 0.0;
            DD t = lane < nw ? sh[q][lane] : z;
            t = warp_sum(t);
            if (lane == 0) {                               // thread 0: the same thread fences and arrives at the barrier
                __stcg(P + (size_t)(2 * q) * COOP_MAXG + blockIdx.x, t.hi);
                __stcg(P + (size_t)(2 * q + 1) * COOP_MAXG + blockIdx.x, t.lo);
            }
        }
    }
    if (!coop_barrier(cb, epoch)) return false;
    DD s[NV];
#pragma unroll
    for (int q = 0; q < NV; ++q) { s[q].hi = 0.0; s[q].lo = 0.0; }
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            DD o;
            o.hi = __ldcg(P + (size_t)(2 * q) * COOP_MAXG + b);
            o.lo = __ldcg(P + (size_t)(2 * q + 1) * COOP_MAXG + b);
            dd_merge(s[q], o);
        }
    }
#pragma unroll
    for (int q = 0; q < NV; ++q) s[q] = warp_sum(s[q]);
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) sh[q][wid] = s[q];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int q = 0; q < NV; ++q) {
            DD z; z.hi = 0.0; z.lo = 0.0;
            DD t = lane < nw ? sh[q][lane] : z;
            t = warp_sum(t);
            if (lane == 0) bc[q] = __dadd_rn(t.hi, t.lo);
        }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NV; ++q) out[q] = bc[q];
    buf ^= 1;
    return true;
}
template <int RT, int IDX>
__global__ void __launch_bounds__(SWEEP_THREADS, 1) k_sweep_coop(int64_t n, int64_t ld, int R_rt, const int32_t* __restrict__ pred,
                                                                  const double* __restrict__ coef, const double* __restrict__ diag,
                                                                  double* V, double* H, int ldh, int jold, int m, SweepCtl* ctl,
                                                                  double break_tol, const __grid_constant__ FacModel F,
                                                                  const int32_t* __restrict__ states, CoopBuf cb) {
    __shared__ int32_t sstate[IDX ? KFSP_MAX_SPECIES * SWEEP_THREADS : 1];
    int32_t* const sst = sstate + (IDX ? threadIdx.x : 0);
    __shared__ DD sh3[3][32];
    __shared__ double bc3[3];
    __shared__ double cs[MAX_COLS];
    const int R = RT > 0 ? RT : R_rt;
    const int tid = threadIdx.x;
    const bool lead = blockIdx.x == 0 && tid == 0;
    const int64_t i0 = (int64_t)blockIdx.x * SWEEP_THREADS + tid, stride = (int64_t)gridDim.x * SWEEP_THREADS;
    for (int j = tid; j < MAX_COLS; j += SWEEP_THREADS) cs[j] = ctl->colscale[j];
    const int brk0 = ctl->brk;
    __syncthreads();
    if (brk0 != 0) return;                                 // uniform over the grid: written before this launch
    unsigned int epoch = 0;
    int buf = 0;
    int broke = 0;
    for (int J = jold; J <= m; ++J) {
        const double* x = V + (size_t)(J - 1) * ld;
        double* y = V + (size_t)J * ld;
        const double xs = cs[J - 1];
        const double* g = J >= 2 ? V + (size_t)(J - 2) * ld : x;
        const double gs = J >= 2 ? cs[J - 2] : 0.0;
        double* hcol = H + (size_t)(J - 1) * ldh;
        DD acc[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) { acc[q].hi = 0.0; acc[q].lo = 0.0; }
        for (int64_t i = i0; i < n; i += stride) {
            const double sv = spmv_row<RT, IDX>(i, ld, R, pred, coef, diag, x, F, states, sst);
            y[i] = sv;
            const double xi = x[i];
            dd_add_prod(acc[1], xi, sv);
            if (J >= 2) {
                const double gv = __dmul_rn(gs, g[i]);
                dd_add_prod(acc[0], gv, sv);
                dd_add_prod(acc[2], xi, gv);
            }
        }
        double d3[3];
        if (!coop_total<3>(acc, d3, cb, buf, epoch, sh3, bc3)) return;
        double h1 = 0.0, h2 = __dmul_rn(xs, __dmul_rn(xs, d3[1]));
        if (J >= 2) {
            h1 = __dmul_rn(xs, d3[0]);
            h2 = fma(-h1, __dmul_rn(xs, d3[2]), h2);
            if (lead) hcol[J - 2] = h1;                                      // H(J-1,J)
        }
        if (lead) hcol[J - 1] = h2;                                          // H(J,J)
        DD an[1];
        an[0].hi = 0.0; an[0].lo = 0.0;
        for (int64_t i = i0; i < n; i += stride) {
            double inner = __dmul_rn(xs, y[i]);
            if (J >= 2) inner = fma(-h1, __dmul_rn(gs, g[i]), inner);
            const double wi = fma(-h2, __dmul_rn(xs, x[i]), inner);
            y[i] = wi;
            dd_add_prod(an[0], wi, wi);
        }
        double d1[1];
        if (!coop_total<1>(an, d1, cb, buf, epoch, sh3, bc3)) return;       // also: column J is complete before anyone gathers from it
        const double hn = sqrt(d1[0]);
        if (hn <= break_tol) {                                               // happy breakdown: every CTA holds the same hn
            if (lead) { ctl->scal[SC_HN] = hn; ctl->brk = J; }
            broke = J;
            break;
        }
        if (tid == 0) cs[J] = 1.0 / hn;                                      // DSCAL factor (every CTA keeps its own copy)
        if (lead) hcol[J] = hn;                                              // H(J+1,J)
        __syncthreads();
    }
    if (broke == 0) {
        const double* x = V + (size_t)m * ld;
        double* y = V + (size_t)(m + 1) * ld;
        const double xs = cs[m];
        DD an[1];
        an[0].hi = 0.0; an[0].lo = 0.0;
        for (int64_t i = i0; i < n; i += stride) {
            const double sv = spmv_row<RT, IDX>(i, ld, R, pred, coef, diag, x, F, states, sst);
            y[i] = sv;
            dd_add_prod(an[0], sv, sv);
        }
        double d1[1];
        if (!coop_total<1>(an, d1, cb, buf, epoch, sh3, bc3)) return;
        const double av = __dmul_rn(xs, sqrt(d1[0]));
        if (lead) ctl->scal[SC_AVNORM] = av;
    }
    __syncthreads();
    if (blockIdx.x == 0)
        for (int j = tid; j < MAX_COLS; j += SWEEP_THREADS) ctl->colscale[j] = cs[j];
}

// The same sweep with EVERYTHING in shared memory, for state sets of a couple of thousand states (BASELINE config 1: the toggle
// never exceeds 2454): the gather form of the generator (12R+8 bytes per state) is staged once per sweep, the three vectors of
// the IOP window rotate through three shared buffers, and global memory only receives each finished column (the basis the
// combination kernel reads later).  In k_sweep_small every phase of a column waits ~700 cycles for L2 (two dependent rounds
// in the SpMV, one in the finalising pass); here those are ~30-cycle shared-memory accesses.  Same element operations, same
// double-double sums: bit-identical.  Dynamic shared memory: n * (12R + 32) bytes (sweep_smem_bytes).
__host__ __device__ inline size_t sweep_smem_bytes(int64_t n, int R) { return (size_t)n * (size_t)(12 * R + 32); }
template <int RT>
__global__ void __launch_bounds__(SWEEP_THREADS, 1) k_sweep_smem(int n, int64_t ld, int R_rt, const int32_t* __restrict__ pred,
                                                                  const double* __restrict__ coef, const double* __restrict__ diag,
                                                                  double* V, double* H, int ldh, int jold, int m, SweepCtl* ctl,
                                                                  double break_tol) {
    extern __shared__ __align__(16) unsigned char sweep_dyn[];
    __shared__ DD sh[32];
    __shared__ DD sh3[3][32];
    __shared__ double bc;
    __shared__ double bc3[3];
    __shared__ double cs[MAX_COLS];
    __shared__ int s_brk;
    const int R = RT > 0 ? RT : R_rt;
    const int tid = threadIdx.x;
    double* s_coef = reinterpret_cast<double*>(sweep_dyn);               // [k*n + i]
    double* s_diag = s_coef + (size_t)R * n;
    double* s_vec = s_diag + n;                                          // three vectors of n
    int32_t* s_pred = reinterpret_cast<int32_t*>(s_vec + 3 * (size_t)n);  // [k*n + i]
    for (int j = tid; j < MAX_COLS; j += SWEEP_THREADS) cs[j] = ctl->colscale[j];
    if (tid == 0) s_brk = ctl->brk;
    __syncthreads();
    if (s_brk != 0) return;
    for (int t = tid; t < R * n; t += SWEEP_THREADS) {
        const int k = t / n, i = t - k * n;
        s_pred[t] = pred[(int64_t)k * ld + i];
        s_coef[t] = coef[(int64_t)k * ld + i];
    }
    for (int i = tid; i < n; i += SWEEP_THREADS) {
        s_diag[i] = diag[i];
        s_vec[(size_t)((jold - 1) % 3) * n + i] = V[(size_t)(jold - 1) * ld + i];
        if (jold >= 2) s_vec[(size_t)((jold - 2) % 3) * n + i] = V[(size_t)(jold - 2) * ld + i];
    }
    __syncthreads();
    auto row = [&](int i, const double* x) {
        double sv = -__dmul_rn(s_diag[i], x[i]);
#pragma unroll
        for (int k = 0; k < (RT > 0 ? RT : R); ++k) {
            const int32_t j = s_pred[k * n + i];
            if (j >= 0) sv = fma(s_coef[k * n + i], x[j], sv);
        }
        return sv;
    };
    for (int J = jold; J <= m; ++J) {
        const double* x = s_vec + (size_t)((J - 1) % 3) * n;
        double* y = s_vec + (size_t)(J % 3) * n;
        const double* g = s_vec + (size_t)((J + 1) % 3) * n;              // column J-2
        const double xs = cs[J - 1];
        const double gs = J >= 2 ? cs[J - 2] : 0.0;
        double* hcol = H + (size_t)(J - 1) * ldh;
        DD accA, accB, accC;
        accA.hi = accA.lo = accB.hi = accB.lo = accC.hi = accC.lo = 0.0;
        for (int i = tid; i < n; i += SWEEP_THREADS) {
            const double sv = row(i, x);
            y[i] = sv;
            const double xi = x[i];
            dd_add_prod(accB, xi, sv);
            if (J >= 2) {
                const double gv = __dmul_rn(gs, g[i]);
                dd_add_prod(accA, gv, sv);
                dd_add_prod(accC, xi, gv);
            }
        }
        double dA, dB, dC;
        cta_dd_total3(accA, accB, accC, J >= 2, sh3, bc3, &dA, &dB, &dC);
        double h1 = 0.0, h2 = __dmul_rn(xs, __dmul_rn(xs, dB));
        if (J >= 2) {
            h1 = __dmul_rn(xs, dA);
            h2 = fma(-h1, __dmul_rn(xs, dC), h2);
            if (tid == 0) hcol[J - 2] = h1;                                  // H(J-1,J)
        }
        if (tid == 0) hcol[J - 1] = h2;                                      // H(J,J)
        DD acc; acc.hi = 0.0; acc.lo = 0.0;
        double* yg = V + (size_t)J * ld;
        for (int i = tid; i < n; i += SWEEP_THREADS) {
            double inner = __dmul_rn(xs, y[i]);
            if (J >= 2) inner = fma(-h1, __dmul_rn(gs, g[i]), inner);
            const double wi = fma(-h2, __dmul_rn(xs, x[i]), inner);
            y[i] = wi;
            yg[i] = wi;                                                      // the finished column U_J of the basis
            dd_add_prod(acc, wi, wi);
        }
        const double hn = sqrt(cta_dd_total(acc, sh, &bc));
        if (hn <= break_tol) {                                               // happy breakdown: uniform over the CTA
            if (tid == 0) { ctl->scal[SC_HN] = hn; ctl->brk = J; }
            s_brk = J;
            break;
        }
        if (tid == 0) { hcol[J] = hn; cs[J] = 1.0 / hn; }                    // H(J+1,J), DSCAL factor
        __syncthreads();
    }
    if (s_brk == 0) {
        const double* x = s_vec + (size_t)(m % 3) * n;
        double* yg = V + (size_t)(m + 1) * ld;
        const double xs = cs[m];
        DD acc; acc.hi = 0.0; acc.lo = 0.0;
        for (int i = tid; i < n; i += SWEEP_THREADS) {
            const double sv = row(i, x);
            yg[i] = sv;
            dd_add_prod(acc, sv, sv);
        }
        const double av = __dmul_rn(xs, sqrt(cta_dd_total(acc, sh, &bc)));
        if (tid == 0) ctl->scal[SC_AVNORM] = av;
    }
    __syncthreads();
    for (int j = tid; j < MAX_COLS; j += SWEEP_THREADS) ctl->colscale[j] = cs[j];
}

// Cross-GPU barrier over peer memory (one warp): every rank raises its flag on every peer and waits for all
// of them.  Needed where a kernel without a reduction (k_scale_copy writing basis column 0) is followed by a
// SpMV that gathers that column from the neighbours' HBM.
__global__ void k_dist_barrier(const DistPeers* __restrict__ dp, unsigned long long seq) {
    pdl_trigger();
    pdl_wait();
    const int P = dp->nranks, me = dp->rank, r = threadIdx.x;
    const int slot = (int)(seq & 1ull);
    if (r < P) {
        __threadfence_system();
        *((volatile unsigned long long*)(dp->flag[r] + (size_t)slot * P + me)) = seq;
        const volatile unsigned long long* fl = dp->flag[me] + (size_t)slot * P + r;
        const long long t0 = clock64();
        const volatile int32_t* gone = dp->err;
        while (*fl != seq) {
            if (*gone & 32) break;
            if (clock64() - t0 > 8000000000LL) { atomicOr(dp->err, 32); break; }
        }
        __threadfence_system();
    }
}

__global__ void k_set_entry(double* p, double v) { *p = v; }
__global__ void k_reset_ctl(SweepCtl* ctl) {
    for (int j = threadIdx.x; j < MAX_COLS; j += blockDim.x) ctl->colscale[j] = 1.0;
    if (threadIdx.x == 0) ctl->brk = 0;
}

}  // namespace kfsp

// Device state space: the reference's FINITE_STATE_PROJECTION (src/state_space/
// StateSpace.f90:13-45) re-designed for batch-parallel updates.
//
//  * The reference indexes states through Brent's open-addressing table keyed by a
//    140-byte big integer (src/hash_table/HashTable.f90).  Only the key -> index map is
//    observable, so the device table is a power-of-two linear-probing table whose slots
//    hold just the int32 state index; a probe compares the S int32 counts of the state
//    itself.  Claiming a slot is therefore a single 32-bit atomicCAS, and "first
//    occurrence wins" (the order the serial reference inserts in) is a single atomicMin.
//  * State indices are insertion order (StateSpace.f90:195-198).  Batch insertion keeps
//    that order: candidates are generated in the serial visiting order, duplicates are
//    resolved to the lowest candidate number, winners are compacted stably.
//  * The generator is kept in BOTH orientations: the reference's column form
//    (succ/prop/diag == ADJ/OFFDIAG/DIAG, StateSpace.f90:13-17) which expansion and the
//    C ABI export need, and the row ("gather") form pred/coef the SpMV streams.
//    All per-reaction arrays are reaction-major ([k*ld + i]) so that consecutive threads
//    read consecutive addresses.
#pragma once
#include "common.cuh"

namespace kfsp {

struct FspView {
    int32_t S, R;
    int64_t ld;                 // leading dimension (capacity in states)
    int64_t n;                  // FSP%SIZE
    int32_t* states;            // [i*S + s]
    int32_t* succ;              // [k*ld + i]  index of x_i + nu_k | IDX_ABSENT | IDX_ILLEGAL
    double* prop;               // [k*ld + i]  a_k(x_i)                      (OFFDIAG)
    double* diag;               // [i]         sum_k a_k(x_i)                (DIAG)
    int32_t* pred;              // [k*ld + i]  index of x_i - nu_k | IDX_ABSENT | IDX_ILLEGAL
    double* coef;               // [k*ld + i]  a_k(x_i - nu_k)
    int32_t* table;             // open addressing, value = state index or SLOT_EMPTY
    uint32_t mask;              // table size - 1
    const DeviceModel* model;
};

enum DevErr : int32_t { DEV_OK = 0, DEV_DUPLICATE = 1, DEV_BAD_STATE = 2, DEV_MOLECULE_LIMIT = 4, DEV_RUNAWAY = 8,
                        DEV_TABLE_FULL = 16 };

// Host-evaluated propensities (CUSTOMPROP, src/model/ModelModule.f90:6-12,31,188-190): a host callback
// cannot run inside a device SSA walk, so the walk reads a_k of states that are not yet in the projection
// from this side cache (state -> R propensities + their sum).  A state missing from the cache is appended to
// the request list and the walk is suspended; the host evaluates the requests, inserts them, and the
// suspended walks are replayed from their own Philox sub-stream (Engine::fsp_ssa_hostprop).
// table == nullptr: propensities are evaluated on the device (byte code / tables).
struct PropCache {
    int32_t* states = nullptr;  // [q*S + s]
    int32_t* table = nullptr;   // open addressing, value = q or SLOT_EMPTY
    uint32_t mask = 0;
    int64_t ld = 0;             // capacity in cached states
    double* prop = nullptr;     // [q*(R+1) + k]: a_k of cached state q for k < R, their sum (DIAG) at k = R
    int32_t* req = nullptr;     // request list [r*S + s]
    int32_t* nreq = nullptr;    // number of requests made (may exceed req_cap; the excess is re-requested)
    int32_t req_cap = 0;
};

// ---------------------------------------------------------------------------------------
// propensity interpreter: the stack machine of src/parser/FortranParser.f90:187-302
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ double eval_propensity(const DeviceModel* __restrict__ m, int k, const int32_t* st) {
    const int ts = m->table_species[k];
    if (ts >= 0) {
        int c = st[ts];
        c = c < 0 ? 0 : (c > m->max_molecules ? m->max_molecules : c);
        return __ldg(m->table[k] + c);
    }
    double stack[KFSP_STACK];
    int sp = -1;
    int dp = m->immed_begin[k];
    const int end = m->code_begin[k + 1];
    const int S = m->S;
    for (int ip = m->code_begin[k]; ip < end; ++ip) {
        const int op = m->code[ip];
        switch (op) {
        case 1: stack[++sp] = m->immed[dp++]; break;
        case 2: stack[sp] = -stack[sp]; break;
        case 3: stack[sp - 1] = __dadd_rn(stack[sp - 1], stack[sp]); --sp; break;
        case 4: stack[sp - 1] = __dsub_rn(stack[sp - 1], stack[sp]); --sp; break;
        case 5: stack[sp - 1] = __dmul_rn(stack[sp - 1], stack[sp]); --sp; break;
        case 6:
            if (stack[sp] == 0.0) return 0.0;
            stack[sp - 1] = __ddiv_rn(stack[sp - 1], stack[sp]); --sp; break;
        case 7: stack[sp - 1] = pow(stack[sp - 1], stack[sp]); --sp; break;
        case 8: stack[sp] = fabs(stack[sp]); break;
        case 9: stack[sp] = exp(stack[sp]); break;
        case 10: if (stack[sp] <= 0.0) return 0.0; stack[sp] = log10(stack[sp]); break;
        case 11: if (stack[sp] <= 0.0) return 0.0; stack[sp] = log(stack[sp]); break;
        case 12: if (stack[sp] < 0.0) return 0.0; stack[sp] = sqrt(stack[sp]); break;
        case 13: stack[sp] = sinh(stack[sp]); break;
        case 14: stack[sp] = cosh(stack[sp]); break;
        case 15: stack[sp] = tanh(stack[sp]); break;
        case 16: stack[sp] = sin(stack[sp]); break;
        case 17: stack[sp] = cos(stack[sp]); break;
        case 18: stack[sp] = tan(stack[sp]); break;
        case 19: if (stack[sp] < -1.0 || stack[sp] > 1.0) return 0.0; stack[sp] = asin(stack[sp]); break;
        case 20: if (stack[sp] < -1.0 || stack[sp] > 1.0) return 0.0; stack[sp] = acos(stack[sp]); break;
        case 21: stack[sp] = atan(stack[sp]); break;
        default: {
            const int v = op - 22;
            stack[++sp] = v < S ? (double)st[v] : m->params[v - S];
        } break;
        }
    }
    return stack[0];
}

// ---------------------------------------------------------------------------------------
// hash table
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ bool same_state(const int32_t* a, const int32_t* b, int S) {
    bool eq = true;
    for (int s = 0; s < S; ++s) eq = eq && (a[s] == b[s]);
    return eq;
}

// FSP%INDEX (StateSpace.f90:116-134): 0-based index or IDX_ABSENT
__device__ __forceinline__ int32_t table_lookup(const FspView& f, const int32_t* st) {
    uint32_t slot = (uint32_t)hash_state(st, f.S) & f.mask;
    for (uint32_t probes = 0; probes <= f.mask; ++probes) {
        const int32_t v = f.table[slot];
        if (v == SLOT_EMPTY) return IDX_ABSENT;
        if (same_state(f.states + (int64_t)v * f.S, st, f.S)) return v;
        slot = (slot + 1) & f.mask;
    }
    return IDX_ABSENT;
}

__device__ __forceinline__ int32_t cache_lookup(const PropCache& pc, const int32_t* st, int S) {
    uint32_t slot = (uint32_t)hash_state(st, S) & pc.mask;
    for (uint32_t probes = 0; probes <= pc.mask; ++probes) {
        const int32_t v = pc.table[slot];
        if (v == SLOT_EMPTY) return IDX_ABSENT;
        if (same_state(pc.states + (int64_t)v * S, st, S)) return v;
        slot = (slot + 1) & pc.mask;
    }
    return IDX_ABSENT;
}

// Insert states [first, first+count) which are already stored in f.states (no duplicates expected).
__global__ void k_insert_states(FspView f, int64_t first, int64_t count, int32_t* err) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = first + t;
        const int32_t* st = f.states + i * f.S;
        uint32_t slot = (uint32_t)hash_state(st, f.S) & f.mask;
        uint32_t probes = 0;
        for (;;) {
            const int32_t old = atomicCAS(&f.table[slot], SLOT_EMPTY, (int32_t)i);
            if (old == SLOT_EMPTY) break;
            if (same_state(f.states + (int64_t)old * f.S, st, f.S)) { atomicOr(err, DEV_DUPLICATE); break; }
            slot = (slot + 1) & f.mask;
            if (++probes > f.mask) { atomicOr(err, DEV_TABLE_FULL); break; }
        }
    }
}

// negative counts or counts above MAXNUMBERMOLECULES: STATE2KEY returns the 0 flag (HashTable.f90:51-57)
__global__ void k_validate_states(const int32_t* states, int S, int64_t n, int32_t maxmol, int32_t* err) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n * S; t += (int64_t)gridDim.x * blockDim.x) {
        const int32_t v = states[t];
        if (v < 0 || v > maxmol) atomicOr(err, DEV_BAD_STATE);
    }
}

// Column of the generator for states [first, first+count): OFFDIAG(K,I)=a_K(x_I), DIAG(I)=sum in K order
// (StateSpace.f90:205-212, 303-314)
__global__ void k_propensities(FspView f, int64_t first, int64_t count) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = first + t;
        int32_t st[KFSP_MAX_SPECIES];
        for (int s = 0; s < f.S; ++s) st[s] = f.states[i * f.S + s];
        double d = 0.0;
        for (int k = 0; k < f.R; ++k) {
            const double a = eval_propensity(f.model, k, st);
            d = __dadd_rn(d, a);
            f.prop[(int64_t)k * f.ld + i] = a;
        }
        f.diag[i] = d;
    }
}

__global__ void k_fill_i32(int32_t* p, int64_t n, int32_t v) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) p[t] = v;
}

// rows [first, first+count) of the reaction-major int arrays succ and pred := IDX_ABSENT
__global__ void k_reset_links(FspView f, int64_t first, int64_t count) {
    const int64_t total = count * f.R;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = t / count, i = first + t % count;
        f.succ[k * f.ld + i] = IDX_ABSENT;
        f.pred[k * f.ld + i] = IDX_ABSENT;
        if (f.coef) f.coef[k * f.ld + i] = 0.0;                      // no coefficient array in the index-only variant
    }
}

// (Re)resolve every link that is still IDX_ABSENT: forward ADJ(K,J) as ADD_STATE/MATRIX_STARTER
// do by lookup (StateSpace.f90:213-236, 305-327) and the row form by looking up x - nu_k
// (the back-links of :240-244, 330-343 seen from the receiving row).
__global__ void k_resolve_links(FspView f) {
    const int64_t total = f.n * f.R;
    const DeviceModel* __restrict__ m = f.model;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = t / f.n, i = t % f.n;
        const int64_t e = k * f.ld + i;
        const int32_t su = f.succ[e], pr = f.pred[e];
        if (su != IDX_ABSENT && pr != IDX_ABSENT) continue;
        int32_t st[KFSP_MAX_SPECIES], nb[KFSP_MAX_SPECIES];
        for (int s = 0; s < f.S; ++s) st[s] = f.states[i * f.S + s];
        if (su == IDX_ABSENT) {
            bool neg = false;
            for (int s = 0; s < f.S; ++s) { nb[s] = st[s] + m->stoich[k * f.S + s]; neg = neg || nb[s] < 0; }
            f.succ[e] = neg ? IDX_ILLEGAL : table_lookup(f, nb);
        }
        if (pr == IDX_ABSENT) {
            bool neg = false;
            for (int s = 0; s < f.S; ++s) { nb[s] = st[s] - m->stoich[k * f.S + s]; neg = neg || nb[s] < 0; }
            if (neg) {
                f.pred[e] = IDX_ILLEGAL;
            } else {
                const int32_t j = table_lookup(f, nb);
                f.pred[e] = j;
                if (j >= 0 && f.coef) f.coef[e] = f.prop[k * f.ld + j];
            }
        }
    }
}

// ---------------------------------------------------------------------------------------
// candidate generation: ONESTEP_EXTENDER (StateSpace.f90:366-395)
// ---------------------------------------------------------------------------------------
__global__ void k_onestep_count(FspView f, int64_t n_old, int32_t* cnt) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_old; i += (int64_t)gridDim.x * blockDim.x) {
        int c = 0;
        for (int k = 0; k < f.R; ++k) c += (f.succ[(int64_t)k * f.ld + i] == IDX_ABSENT);
        cnt[i] = c;
    }
}
__global__ void k_onestep_fill(FspView f, int64_t n_old, const int32_t* __restrict__ off, int32_t* cand, int32_t* err) {
    const DeviceModel* __restrict__ m = f.model;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n_old; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t o = off[i];
        for (int k = 0; k < f.R; ++k) {
            if (f.succ[(int64_t)k * f.ld + i] != IDX_ABSENT) continue;
            for (int s = 0; s < f.S; ++s) {
                const int32_t v = f.states[i * f.S + s] + m->stoich[k * f.S + s];
                if (v > m->max_molecules) atomicOr(err, DEV_MOLECULE_LIMIT);
                cand[o * f.S + s] = v;
            }
            ++o;
        }
    }
}

// ---------------------------------------------------------------------------------------
// candidate generation: SSA_EXTENDER (StateSpace.f90:571-629).  One thread per start state;
// the walk is replayed twice (count, then fill) from its own Philox sub-stream.
// (Measured: persistent lanes with dynamic work fetch were 2x SLOWER on the Goutsias model -- walks are
// short and the cost is dependent memory latency per walk, not divergence -- so the grid-stride form stays.
// Round 2: a state-major copy {a_1..a_R, DIAG, ADJ_1..ADJ_R} of the column form, one 128-byte record per state instead of
// 2R+1 sectors per jump, changed nothing either (Goutsias SSA phase 1.535 vs 1.541 s): the kernel lasts as long as its longest
// walk, and a jump is a serial chain -- Philox, -log(r)/DIAG, the cumulative sum over the reactions, the successor load.)
// ---------------------------------------------------------------------------------------
// Emission during the counting pass (device-evaluated propensities): the kernel lasts as long as its longest walk, and the
// walks that leave the projection are the long ones, so replaying them to fill the candidate list cost as much as the counting
// pass (Goutsias at 8e5 states: 3.5 ms + 2.8 ms per expansion).  Instead a walk appends each candidate to a temporary area
// (position from one atomicAdd, linked to the walk's previous candidate) and k_ssa_gather copies every walk's chain to its
// scanned position: same list, same order, one pass of walks.  tmp == nullptr, or more candidates than cap: the replay below.
struct SsaEmit {
    int32_t* tmp = nullptr;      // [p*S + s]
    int32_t* prev = nullptr;     // position of the same walk's previous candidate, -1 for its first
    int32_t* head = nullptr;     // per start state: position of its LAST candidate, -1 if none
    int32_t* cursor = nullptr;
    int32_t cap = 0;
};
__global__ void k_ssa_gather(int64_t n_old, const int32_t* __restrict__ off, int64_t ncand, SsaEmit em, int S, int32_t* cand) {
    for (int64_t j0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j0 < n_old; j0 += (int64_t)gridDim.x * blockDim.x) {
        const int64_t o0 = off[j0];
        const int64_t c = (j0 + 1 < n_old ? (int64_t)off[j0 + 1] : ncand) - o0;      // candidates of this walk
        if (c <= 0) continue;
        int32_t p = em.head[j0];
        for (int64_t t = c - 1; t >= 0 && p >= 0; --t) {
            for (int s = 0; s < S; ++s) cand[(o0 + t) * S + s] = em.tmp[(int64_t)p * S + s];
            p = em.prev[p];
        }
    }
}
// RT > 0: the number of reactions at compile time -- the propensities of the current state stay in registers and the search
// for the firing reaction is an unrolled chain (with a run-time R the array is indexed dynamically, i.e. lives in local memory:
// after the factored tables went in, ncu had 25 % of the stall samples on its STL / LDL and the DADD of the cumulative sum).
template <bool FILL, int RT>
__global__ void k_ssa_walk(FspView f, int64_t n_old, double timestep, uint64_t seed, uint32_t call_no,
                           int32_t* cnt, const int32_t* __restrict__ off, int32_t* cand, int32_t* err, int32_t max_jumps,
                           int64_t ncand, PropCache pc, int32_t* wsave, SsaEmit em, const __grid_constant__ FacModel F, int use_fac) {
    const DeviceModel* __restrict__ m = f.model;
    const int S = f.S, R = RT > 0 ? RT : f.R;
    for (int64_t j0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j0 < n_old; j0 += (int64_t)gridDim.x * blockDim.x) {
        // host-propensity rounds: cnt >= 0 walk completed, -1 not started, -2 suspended (state saved in wsave)
        const int32_t status = (!FILL && pc.table) ? cnt[j0] : -1;
        if (status >= 0) continue;
        bool suspended = false;
        if (FILL) {
            // replay only the walks that left the projection in the counting pass (a small boundary fraction)
            const int64_t end = j0 + 1 < n_old ? (int64_t)off[j0 + 1] : ncand;
            if (end == (int64_t)off[j0]) continue;
        }
        int32_t st[KFSP_MAX_SPECIES], nb[KFSP_MAX_SPECIES];
        for (int s = 0; s < S; ++s) st[s] = f.states[j0 * S + s];
        int64_t j = j0;                 // index of the current state, or -1 if it is not (yet) in the projection
        double t = 0.0;
        int32_t emitted = 0;
        int32_t myhead = -1;
        int64_t o = FILL ? (int64_t)off[j0] : 0;
        uint32_t jump0 = 0;
        if (status == -2) {             // resume where the walk stopped: counter-based RNG, so the draws repeat exactly
            const int32_t* __restrict__ sv = wsave + j0 * (int64_t)(S + 4);
            for (int s = 0; s < S; ++s) st[s] = sv[s];
            t = __hiloint2double(sv[S + 1], sv[S]);
            jump0 = (uint32_t)sv[S + 2];
            emitted = sv[S + 3];
            j = -1;
        }
        for (uint32_t jump = jump0;; ++jump) {
            if ((int32_t)jump >= max_jumps) { atomicOr(err, DEV_RUNAWAY); break; }
            double r1, r2;
            philox_uniform2(seed, call_no, (uint32_t)(j0 + 1), jump, &r1, &r2);
            // all R propensities of the current state at once: independent loads (one memory latency instead of
            // a dependent chain through the cumulative sum), and a single byte-code evaluation for unknown states
            double pr[RT > 0 ? RT : KFSP_MAX_REACTIONS];
            double dg;
            if (j >= 0) {
                dg = f.diag[j];
                if constexpr (RT > 0) {
#pragma unroll
                    for (int k = 0; k < RT; ++k) pr[k] = f.prop[(int64_t)k * f.ld + j];
                } else {
                    for (int k = 0; k < R; ++k) pr[k] = f.prop[(int64_t)k * f.ld + j];
                }
            } else if (pc.table) {
                const int32_t q = cache_lookup(pc, st, S);
                if (q < 0) {
                    // unknown to the host cache: request it and suspend (the FILL pass never gets here)
                    const int32_t r = atomicAdd(pc.nreq, 1);
                    if (r < pc.req_cap)
                        for (int s = 0; s < S; ++s) pc.req[(int64_t)r * S + s] = st[s];
                    int32_t* sv = wsave + j0 * (int64_t)(S + 4);
                    for (int s = 0; s < S; ++s) sv[s] = st[s];
                    sv[S] = __double2loint(t); sv[S + 1] = __double2hiint(t);
                    sv[S + 2] = (int32_t)jump; sv[S + 3] = emitted;
                    suspended = true;
                    break;
                }
                const double* __restrict__ row = pc.prop + (int64_t)q * (R + 1);
                dg = row[R];
                if constexpr (RT > 0) {
#pragma unroll
                    for (int k = 0; k < RT; ++k) pr[k] = row[k];
                } else {
                    for (int k = 0; k < R; ++k) pr[k] = row[k];
                }
            } else {
                // a state outside the projection: its propensities are nowhere stored.  The walks that last are exactly those that
                // stay out there, so this evaluation is the serial chain of the kernel (ncu: half of all stall samples on instructions
                // issued with one lane active, the hottest the byte-code interpreter's opcode dispatch): the factored tables of the index-only SpMV
                // (common.cuh, same values bit for bit) replace ~60 interpreted opcodes per jump by 10-20 table loads.
                dg = 0.0;
                if constexpr (RT > 0) {
                    if (use_fac) {
#pragma unroll
                        for (int k = 0; k < RT; ++k) { pr[k] = fac_eval<1, 1>(F, k, st, 0); dg = __dadd_rn(dg, pr[k]); }
                    } else {
#pragma unroll
                        for (int k = 0; k < RT; ++k) { pr[k] = eval_propensity(m, k, st); dg = __dadd_rn(dg, pr[k]); }
                    }
                } else {
                    if (use_fac) {
                        for (int k = 0; k < R; ++k) { pr[k] = fac_eval<1, 1>(F, k, st, 0); dg = __dadd_rn(dg, pr[k]); }
                    } else {
                        for (int k = 0; k < R; ++k) { pr[k] = eval_propensity(m, k, st); dg = __dadd_rn(dg, pr[k]); }
                    }
                }
            }
            t = fmin(timestep, __dadd_rn(t, __ddiv_rn(-log(r1), dg)));
            if (!(t <= timestep)) break;
            const double r2a = fmin(__dmul_rn(r2, dg), dg);
            int k = 0;
            double tmp = pr[0];
            if (RT > 0) {               // while (tmp < r2a && k < R - 1) { ++k; tmp += pr[k]; } without indexing pr at run time
                bool go = true;
#pragma unroll
                for (int q = 1; q < (RT > 0 ? RT : 1); ++q) {
                    go = go && tmp < r2a;
                    if (go) { k = q; tmp = __dadd_rn(tmp, pr[q]); }
                }
            } else {
                while (tmp < r2a && k < R - 1) {
                    ++k;
                    tmp = __dadd_rn(tmp, pr[k]);
                }
            }
            bool neg = false, over = false;
            for (int s = 0; s < S; ++s) {
                nb[s] = st[s] + m->stoich[k * S + s];
                neg = neg || nb[s] < 0;
                over = over || nb[s] > m->max_molecules;
            }
            if (neg) break;             // illegal: ADJ(K,J) = -1 and the walk ends (:594-596)
            if (over) { atomicOr(err, DEV_MOLECULE_LIMIT); break; }
            int32_t nj = j >= 0 ? f.succ[(int64_t)k * f.ld + j] : IDX_ABSENT;      // (fetching all R successors with the propensities,
                                                                                     // to save the dependent load, measured 5 % slower)
            if (nj < 0) nj = table_lookup(f, nb);
            for (int s = 0; s < S; ++s) st[s] = nb[s];
            if (nj >= 0) {
                j = nj;
                if (!(t < timestep && j >= j0)) break;
            } else {
                if (FILL) {
                    for (int s = 0; s < S; ++s) cand[o * S + s] = nb[s];
                    ++o;
                } else if (em.tmp) {
                    const int32_t p = atomicAdd(em.cursor, 1);
                    if (p < em.cap) {
                        for (int s = 0; s < S; ++s) em.tmp[(int64_t)p * S + s] = nb[s];
                        em.prev[p] = myhead;
                        myhead = p;
                    }
                }
                ++emitted;
                j = -1;
                if (!(t < timestep)) break;
            }
        }
        if (!FILL) cnt[j0] = suspended ? -2 : emitted;
        if (!FILL && em.tmp) em.head[j0] = myhead;
    }
}

// Host-propensity mode: OFFDIAG/DIAG of the new states [first, first+count) from the side cache;
// cache_q[t] = cache entry used, or -1 (the host evaluates those and scatters them with k_scatter_props).
__global__ void k_props_from_cache(FspView f, int64_t first, int64_t count, PropCache pc, int32_t* cache_q) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = first + t;
        int32_t st[KFSP_MAX_SPECIES];
        for (int s = 0; s < f.S; ++s) st[s] = f.states[i * f.S + s];
        const int32_t q = pc.table ? cache_lookup(pc, st, f.S) : IDX_ABSENT;
        cache_q[t] = q;
        if (q < 0) continue;
        const double* __restrict__ row = pc.prop + (int64_t)q * (f.R + 1);
        for (int k = 0; k < f.R; ++k) f.prop[(int64_t)k * f.ld + i] = row[k];
        f.diag[i] = row[f.R];
    }
}
// vals[t*(R+1) + k] = a_k of state idx[t] (k < R), vals[t*(R+1) + R] = their sum in reaction order
__global__ void k_scatter_props(FspView f, const int32_t* __restrict__ idx, const double* __restrict__ vals, int64_t count) {
    const int64_t total = count * (f.R + 1);
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = t / (f.R + 1);
        const int k = (int)(t % (f.R + 1));
        const int64_t i = idx[c];
        if (k < f.R) f.prop[(int64_t)k * f.ld + i] = vals[t];
        else f.diag[i] = vals[t];
    }
}
// cache entries [first, first+count): states already stored in pc.states; no duplicates (the host dedupes)
__global__ void k_cache_insert(PropCache pc, int S, int64_t first, int64_t count, int32_t* err) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = first + t;
        uint32_t slot = (uint32_t)hash_state(pc.states + q * S, S) & pc.mask;
        uint32_t probes = 0;
        for (;;) {
            if (atomicCAS(&pc.table[slot], SLOT_EMPTY, (int32_t)q) == SLOT_EMPTY) break;
            slot = (slot + 1) & pc.mask;
            if (++probes > pc.mask) { atomicOr(err, DEV_TABLE_FULL); break; }
        }
    }
}

// ---------------------------------------------------------------------------------------
// order-exact batch insertion of candidates
// ---------------------------------------------------------------------------------------
// Candidate c gets the provisional index n + c.  Equal states meet in one slot; atomicMin keeps
// the lowest candidate number, i.e. the first occurrence in the reference's visiting order.
__global__ void k_insert_candidates(FspView f, const int32_t* __restrict__ cand, int64_t ncand, int32_t* cand_slot, int32_t* err) {
    const int32_t n = (int32_t)f.n;
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncand; c += (int64_t)gridDim.x * blockDim.x) {
        const int32_t* st = cand + c * f.S;
        const int32_t mine = n + (int32_t)c;
        uint32_t slot = (uint32_t)hash_state(st, f.S) & f.mask;
        uint32_t probes = 0;
        for (;;) {
            int32_t v = f.table[slot];
            if (v == SLOT_EMPTY) {
                v = atomicCAS(&f.table[slot], SLOT_EMPTY, mine);
                if (v == SLOT_EMPTY) break;                         // claimed
            }
            const int32_t* other = v >= n ? cand + (int64_t)(v - n) * f.S : f.states + (int64_t)v * f.S;
            if (same_state(other, st, f.S)) {
                if (v >= n) atomicMin(&f.table[slot], mine);
                break;
            }
            slot = (slot + 1) & f.mask;
            if (++probes > f.mask) { atomicOr(err, DEV_TABLE_FULL); break; }
        }
        cand_slot[c] = (int32_t)slot;
    }
}
__global__ void k_mark_winners(const int32_t* __restrict__ table, const int32_t* __restrict__ cand_slot, int64_t ncand,
                               int32_t n, int32_t* win) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncand; c += (int64_t)gridDim.x * blockDim.x)
        win[c] = table[cand_slot[c]] == n + (int32_t)c;
}
__global__ void k_commit_winners(FspView f, const int32_t* __restrict__ cand, const int32_t* __restrict__ cand_slot,
                                 const int32_t* __restrict__ win, const int32_t* __restrict__ pos, int64_t ncand, double* w) {
    const int32_t n = (int32_t)f.n;
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncand; c += (int64_t)gridDim.x * blockDim.x) {
        if (!win[c]) continue;
        const int32_t idx = n + pos[c];
        for (int s = 0; s < f.S; ++s) f.states[(int64_t)idx * f.S + s] = cand[c * f.S + s];
        f.table[cand_slot[c]] = idx;
        w[idx] = 0.0;                                                // FSP%VECTOR(SIZE) = 0 (StateSpace.f90:199)
    }
}
// Undo a failed batch (overflow): free the slots the candidates claimed.
__global__ void k_rollback_candidates(int32_t* table, const int32_t* __restrict__ cand_slot, int64_t ncand, int32_t n) {
    for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < ncand; c += (int64_t)gridDim.x * blockDim.x) {
        const int32_t s = cand_slot[c];
        if (table[s] >= n) table[s] = SLOT_EMPTY;
    }
}

// ---------------------------------------------------------------------------------------
// exclusive scan of int32 (counts / flags)
// ---------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 512;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int32_t block_exclusive_scan(int32_t v, int32_t* total) {
    __shared__ int32_t warp_sums[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int32_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int32_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        int32_t s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int32_t y = __shfl_up_sync(0xffffffffu, s, d);
            if (lane >= d) s += y;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    const int32_t base = wid > 0 ? warp_sums[wid - 1] : 0;
    *total = warp_sums[SCAN_THREADS / 32 - 1];
    const int32_t r = base + x - v;
    __syncthreads();
    return r;
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const int32_t* __restrict__ in, int32_t* out, int64_t n,
                                                              int32_t* tile_sums, const int32_t* __restrict__ tile_offsets) {
    const int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    int32_t s = 0;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q) {
        v[q] = base + q < n ? in[base + q] : 0;
        s += v[q];
    }
    int32_t total;
    int32_t ex = block_exclusive_scan(s, &total);
    if (tile_sums != nullptr) {
        if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
        return;
    }
    ex += tile_offsets ? tile_offsets[blockIdx.x] : 0;
#pragma unroll
    for (int q = 0; q < SCAN_ITEMS; ++q) {
        if (base + q < n) out[base + q] = ex;
        ex += v[q];
    }
}

// ---------------------------------------------------------------------------------------
// DROP_STATES pieces (StateSpace.f90:398-548)
// ---------------------------------------------------------------------------------------
// DROP(I) = W(I) < DROPTOL; count (StateSpace.f90:475-484)
__global__ void k_drop_mark(const double* __restrict__ w, int64_t n, double droptol, int32_t* drop, unsigned long long* count) {
    unsigned long long c = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int d = w[i] < droptol;
        drop[i] = d;
        c += d;
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(count, c);
}
// states whose derivative (A w)_i exceeds 1e-8 are kept; the counter is decremented for every
// such state, marked or not (StateSpace.f90:490-494)
__global__ void k_drop_unmark(const double* __restrict__ aw, int64_t n, double tol, int32_t* drop, unsigned long long* uncount) {
    unsigned long long c = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (aw[i] > tol) { drop[i] = 0; ++c; }
    }
    for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(uncount, c);
}
__global__ void k_invert_flags(const int32_t* __restrict__ drop, int32_t* keep, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) keep[i] = !drop[i];
}
// stable compaction (StateSpace.f90:506-538): element i moves to pos[i] if kept.
// rows = number of reaction-major rows (1 for plain vectors); width = ints/doubles per element for AoS arrays.
template <typename T>
__global__ void k_compact_rows(const T* __restrict__ src, T* dst, const int32_t* __restrict__ keep, const int32_t* __restrict__ pos,
                               int64_t n, int64_t ld_src, int64_t ld_dst, int rows) {
    const int64_t total = n * rows;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = t / n, i = t % n;
        if (keep[i]) dst[k * ld_dst + pos[i]] = src[k * ld_src + i];
    }
}
__global__ void k_compact_states(const int32_t* __restrict__ src, int32_t* dst, const int32_t* __restrict__ keep,
                                 const int32_t* __restrict__ pos, int64_t n, int S) {
    const int64_t total = n * S;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / S, s = t % S;
        if (keep[i]) dst[(int64_t)pos[i] * S + s] = src[t];
    }
}
// succ re-indexing with compaction: ADJ>0 -> NEW_INDEX (0 if the target was dropped), -1 stays (StateSpace.f90:540-545)
__global__ void k_compact_succ(const int32_t* __restrict__ src, int32_t* dst, const int32_t* __restrict__ keep,
                               const int32_t* __restrict__ pos, int64_t n, int64_t ld_src, int64_t ld_dst, int rows) {
    const int64_t total = n * rows;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = t / n, i = t % n;
        if (!keep[i]) continue;
        int32_t a = src[k * ld_src + i];
        if (a >= 0) a = keep[a] ? pos[a] : IDX_ABSENT;
        dst[k * ld_dst + pos[i]] = a;
    }
}

// export to the reference's column layout with Fortran index conventions
__global__ void k_export_adj(FspView f, int32_t* adj /*[i*R+k]*/, double* offdiag /*[i*R+k]*/) {
    const int64_t total = f.n * f.R;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / f.R, k = t % f.R;
        const int32_t a = f.succ[k * f.ld + i];
        if (adj) adj[t] = a >= 0 ? a + 1 : (a == IDX_ABSENT ? 0 : -1);
        if (offdiag) offdiag[t] = f.prop[k * f.ld + i];
    }
}
__global__ void k_lookup_states(FspView f, const int32_t* __restrict__ q, int64_t nq, int32_t* idx1, const double* __restrict__ w, double* p) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nq; t += (int64_t)gridDim.x * blockDim.x) {
        int32_t st[KFSP_MAX_SPECIES];
        bool bad = false;
        for (int s = 0; s < f.S; ++s) { st[s] = q[t * f.S + s]; bad = bad || st[s] < 0 || st[s] > f.model->max_molecules; }
        const int32_t j = bad ? IDX_ABSENT : table_lookup(f, st);
        if (idx1) idx1[t] = j >= 0 ? j + 1 : 0;
        if (p) p[t] = j >= 0 ? w[j] : 0.0;
    }
}

}  // namespace kfsp

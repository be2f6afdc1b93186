// Row-partitioned multi-GPU expv.  One process per GPU; rank r owns the contiguous
// block of rows [lo_r, hi_r) of the global state list, its rows of the gather-form generator and its
// slices of every basis vector.  Per SpMV one exchange step moves the x entries other ranks' rows refer
// to (grouped ncclSend/ncclRecv); the three reductions of an Arnoldi column all-gather each rank's
// DOUBLE-DOUBLE partial (2 doubles) and every rank merges them in rank order, so the rounded scalars --
// and therefore H, the Pade result, every controller decision and the final vector -- are bit-identical
// to the single-GPU run for any number of GPUs.  H, exp(tH) and the controller are replicated.
//
// Scope (round 1): fixed state set (FSP adaptivity off, BASELINE config 5).  Every rank keeps the full
// state list and hash table (0.8 GB + 1 GB at 1e8 states) and builds only its own matrix rows.
#pragma once
#include "common.cuh"
#include "krylov.cuh"
#include "state_space.cuh"

#ifdef KFSP_WITH_NCCL
#include <nccl.h>
#endif

namespace kfsp {

// block partition of n rows over p ranks: first (n % p) ranks get one extra row
__host__ __device__ inline int64_t part_lo(int64_t n, int p, int r) {
    const int64_t q = n / p, rem = n % p;
    return q * r + (r < rem ? r : rem);
}
__host__ __device__ inline int part_owner(int64_t n, int p, int64_t g) {
    const int64_t q = n / p, rem = n % p;
    const int64_t cut = (q + 1) * rem;
    return g < cut ? (int)(g / (q + 1)) : (int)(rem + (g - cut) / (q > 0 ? q : 1));
}

// Rows [lo, lo+nloc) of the generator in gather form.  pred receives GLOBAL indices first
// (k_dist_remap turns them into local / halo positions); coef = a_k(x - nu_k) is evaluated from the
// predecessor STATE, so no remote propensity is read.  f.n must be the GLOBAL size (table lookups).
__global__ void k_dist_build_rows(FspView f, int64_t lo, int64_t nloc) {
    const DeviceModel* __restrict__ m = f.model;
    for (int64_t il = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; il < nloc; il += (int64_t)gridDim.x * blockDim.x) {
        int32_t st[KFSP_MAX_SPECIES], nb[KFSP_MAX_SPECIES];
        for (int s = 0; s < f.S; ++s) st[s] = f.states[(lo + il) * f.S + s];
        double d = 0.0;
        for (int k = 0; k < f.R; ++k) {
            const double a = eval_propensity(m, k, st);
            d = __dadd_rn(d, a);
            f.prop[(int64_t)k * f.ld + il] = a;
            bool neg = false;
            for (int s = 0; s < f.S; ++s) { nb[s] = st[s] - m->stoich[k * f.S + s]; neg = neg || nb[s] < 0; }
            int32_t j = IDX_ILLEGAL;
            double c = 0.0;
            if (!neg) {
                j = table_lookup(f, nb);
                if (j >= 0 && f.coef) c = eval_propensity(m, k, nb);
            }
            f.pred[(int64_t)k * f.ld + il] = j;
            if (f.coef) f.coef[(int64_t)k * f.ld + il] = c;
            f.succ[(int64_t)k * f.ld + il] = IDX_ABSENT;
        }
        f.diag[il] = d;
    }
}
// flag[g] = 1 for every global index outside [lo,hi) referenced by a local row
__global__ void k_dist_mark_remote(const int32_t* __restrict__ pred, int64_t ld, int R, int64_t nloc, int64_t lo, int64_t hi, int32_t* flag) {
    const int64_t total = nloc * R;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = t / nloc, il = t % nloc;
        const int32_t g = pred[k * ld + il];
        if (g >= 0 && (g < lo || g >= hi)) flag[g] = 1;
    }
}
// halo_g[pos[g]] = g for flagged g (ascending => grouped by owner)
__global__ void k_dist_halo_list(const int32_t* __restrict__ flag, const int32_t* __restrict__ pos, int64_t nglobal, int32_t* halo_g) {
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < nglobal; g += (int64_t)gridDim.x * blockDim.x)
        if (flag[g]) halo_g[pos[g]] = (int32_t)g;
}
// local rows: global predecessor index -> local index, or nloc + halo position
__global__ void k_dist_remap(int32_t* pred, int64_t ld, int R, int64_t nloc, int64_t lo, int64_t hi, const int32_t* __restrict__ pos) {
    const int64_t total = nloc * R;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = t / nloc, il = t % nloc;
        const int32_t g = pred[k * ld + il];
        if (g < 0) continue;
        pred[k * ld + il] = (g >= lo && g < hi) ? (int32_t)(g - lo) : (int32_t)(nloc + pos[g]);
    }
}
// first position in the ascending list with value >= bound[r], for r = 0..p
__global__ void k_dist_bounds(const int32_t* __restrict__ halo_g, int64_t nh, const int64_t* __restrict__ bound, int p, int64_t* off) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r > p) return;
    int64_t a = 0, b = nh;
    while (a < b) {
        const int64_t mid = (a + b) / 2;
        if ((int64_t)halo_g[mid] < bound[r]) a = mid + 1; else b = mid;
    }
    off[r] = a;
}
__global__ void k_dist_to_local(int32_t* idx, int64_t cnt, int64_t lo) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < cnt; t += (int64_t)gridDim.x * blockDim.x) idx[t] -= (int32_t)lo;
}
// sendbuf[q] = x[send_idx[q]]: the entries other ranks asked for, grouped by requesting rank
__global__ void k_dist_pack(const double* __restrict__ x, const int32_t* __restrict__ send_idx, int64_t cnt, double* __restrict__ sendbuf) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < cnt; t += (int64_t)gridDim.x * blockDim.x) sendbuf[t] = x[send_idx[t]];
}

// halo position -> (owner rank, row on owner), for the peer-load SpMV
__global__ void k_dist_halo_owner(const int32_t* __restrict__ halo_g, int64_t nh, int64_t nglobal, int p, int32_t* owner, int32_t* lidx) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < nh; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = halo_g[t];
        const int o = part_owner(nglobal, p, g);
        owner[t] = o;
        lidx[t] = (int32_t)(g - part_lo(nglobal, p, o));
    }
}

struct Dist {
    int rank = 0, nranks = 1;
    // Adaptive state sets (expansion / pruning enabled): every rank keeps the single-GPU layout of the WHOLE state space
    // (states, hash table, both orientations of the generator, W, basis) and computes the rows [rb[rank], rb[rank+1]) of
    // every N-sized operation of the Krylov loop; expansion and pruning run on the gathered W, identically on every rank
    // (Engine::repl_enter / repartition).  Fixed sets (config 5) use the memory-scaled layout instead: local rows only.
    bool repl = false;
    bool suspended = false;             // inside repl_enter/repl_leave: kernels see the whole state space, reductions stay local
    bool whole = false;                 // the state set is still too small for a split to pay (Engine::repartition): every rank
                                        // computes every row, nothing is exchanged; rows are split once it has grown
    int64_t repl_min_rows = 1 << 22;    // KFSP_REPL_MIN_ROWS
    int64_t rb[9] = {0};                // row bounds of the replicated-layout partition
    bool p2p = false;                   // peer-memory path active (cudaIpc over NVLink); else NCCL path
    bool want_p2p = true;
    bool p2p_red = true, p2p_halo = true;   // which halves use peer memory (KFSP_DIST_P2P: 1 both, 2 reductions only, 3 halo only)
    void* xchg = nullptr;               // this rank's exchange area + flags (written by peers)
    DistPeers* d_peers = nullptr;       // device copy of the peer table
    int32_t* halo_owner = nullptr;
    int32_t* halo_lidx = nullptr;
    void* peer_base[8] = {nullptr};     // opened IPC mappings (for closing)
    void* peer_xbase[8] = {nullptr};
    unsigned long long seq = 0;
    unsigned long long* d_stat = nullptr;   // exchange statistics written by the reducing kernels (DistPeers::stat)
#ifdef KFSP_WITH_NCCL
    ncclComm_t comm = nullptr;
#endif
    int64_t n_global = 0, lo = 0, hi = 0;
    int64_t n_halo = 0, n_send = 0;
    int32_t* send_idx = nullptr;        // local row indices to pack, grouped by requesting rank
    double* sendbuf = nullptr;
    double* halo = nullptr;             // received x entries, ordered by global index
    double* red_send = nullptr;         // 4 doubles: this rank's (hi,lo) partials
    double* red_recv = nullptr;         // nranks * 4
    std::vector<int64_t> send_off, recv_off;   // per-peer segments of sendbuf / halo
    int64_t halo_exchanges = 0, halo_bytes = 0, reductions = 0;
};

}  // namespace kfsp

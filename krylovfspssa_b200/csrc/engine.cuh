// Host-side engine: owns the device buffers of one solver handle and sequences the kernels.
#pragma once
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_set>
#include <vector>

#include "common.cuh"
#include "dist.cuh"
#include "expm.cuh"
#include "krylov.cuh"
#include "lattice.cuh"
#include "model_host.h"
#include "state_space.cuh"

namespace kfsp {

inline double wall_now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Engine {
    kfsp_options opt{};
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;   // side stream: uploads that the solve does not depend on (the lattice's state-list check)
    cudaEvent_t ev_side = nullptr;
    int32_t* d_chk = nullptr;         // result of the deferred lattice check
    bool box_check_pending = false;
    bool have_model = false;
    DeviceModel* d_model = nullptr;
    double* d_tables = nullptr;       // host-built propensity tables
    int n_tabulated = 0, n_inexact_on_device = 0, n_host_evaluated = 0;
    int S = 0, R = 0;
    // host-evaluated propensities: CUSTOMPROP callbacks (ModelModule.f90:6-12,188-190), and byte code that holds a
    // transcendental operation on SEVERAL species (not tabulable): the CUDA math library may differ from the host libm by
    // ulps there, which could flip an SSA reaction pick or a DROP decision, so such programs are evaluated by the host
    // interpreter through the CUSTOMPROP machinery (KFSP_DEVICE_MATH=1 keeps them on the device; kfsp_model_info reports which happened)
    bool host_prop = false;
    HostModel hm;                     // copy of the host model (callback, parameters, programs)
    PropCache pc;                     // device side cache used by SSA walks in host_prop mode
    int64_t pc_n = 0;                 // cached states
    int64_t pc_budget = 16 << 20;     // ... the cache is emptied when it holds more (KFSP_PROP_CACHE_STATES)
    char* hp_dev = nullptr;           // device staging of the host-propensity path (grows on demand)
    char* hp_host = nullptr;          // pinned host staging
    size_t hp_dev_bytes = 0, hp_host_bytes = 0;
    int64_t host_prop_evals = 0, host_prop_rounds = 0;
    // matrix-free lattice (opt.spmv_variant == 1, lattice.cuh): the projection is a full box in natural order
    bool box = false;
    Lattice lat{};
    DeviceModel h_dm{};               // host mirror of *d_model (table pointers, stoichiometry)
    int box_tune = 0;                 // KFSP_BOX_TUNE: rows per batch / CTAs per SM of the lattice SpMV (A/B)
    int bd2_zc = 0;                   // KFSP_BD2_ZC: forced rows per z-chunk of the stencil kernel (0 = automatic)
    int bd2_ahead = 0;                // KFSP_BD2_AHEAD: rows beyond the ring prefetched into L2 by the multi-stream stencil launches.  Off:
                                      // with the cp.async ring 6 rows deep it only adds traffic (ncu: 2.86 GB instead of 2.45 GB of
                                      // DRAM reads per launch at 1e8 states, 0.85 instead of 0.73 ms; profiles/r2_summary.md)
    int bd2_sync = 8;                 // KFSP_BD2_SYNC: the stencil kernel's CTAs re-align their warps every so many rows (power of two; 0 = never)

    // index-only SpMV (opt.spmv_variant == 2, krylov.cuh: k_spmv_idx): coef is never stored, a_k(x - nu_k) is recomputed
    // from the row's integer state through the factored propensity tables
    bool idx = false;
    bool fac_ok = false, ssa_fac_on = true;      // `fac` holds the model's factored tables (any variant)
    FacModel fac{};
    double* d_factabs = nullptr;

    // state space
    int64_t ld = 0;                   // capacity in states (multiple of 64)
    int64_t n = 0;
    int32_t* d_states = nullptr;
    int32_t* d_succ = nullptr;
    double* d_prop = nullptr;
    double* d_diag = nullptr;
    int32_t* d_pred = nullptr;
    double* d_coef = nullptr;
    double* d_w = nullptr;            // FSP%VECTOR == W
    int32_t* d_table = nullptr;
    int64_t table_size = 0;
    int32_t* d_err = nullptr;         // DevErr bits
    uint32_t ssa_calls = 0;
    Dist dist;                        // multi-GPU row partition (nranks == 1: single GPU)
    int64_t states_cap = 0;           // capacity of d_states / table (global); ld is the capacity of the row arrays
    bool small_sweep = true;          // KFSP_SMALL_SWEEP=0 disables the single-CTA sweep (A/B)
    bool smem_sweep = true;           // KFSP_SMEM_SWEEP=0 disables its shared-memory-resident form (A/B)
    static constexpr size_t SWEEP_SMEM_MAX = 220 * 1024;
    std::vector<const void*> smem_sweep_ready;
    int spmv_tune = 0;                // hoisted loads, one row per iteration, grid = one wave of resident CTAs

    // scratch arena (grows on demand)
    char* d_scratch = nullptr;
    size_t scratch_bytes = 0;

    // krylov
    double* d_V = nullptr;            // ld x (m_max+2)
    int LDH = 0;                      // m_max + 2
    double* d_H = nullptr;            // LDH x LDH
    double* d_expm_work = nullptr;
    double* d_expm_full = nullptr;
    ExpmResult* d_res = nullptr;
    ExpmResult* h_res = nullptr;      // pinned
    SweepCtl* d_ctl = nullptr;
    SweepCtl* h_ctl = nullptr;        // pinned
    Reducer rd{};
    double* d_flush = nullptr;
    size_t flush_bytes = 0;

    // bookkeeping
    double phase_s[8] = {0};          // host wall clock per phase of the last solve: 0 sweep+Pade, 1 combine/norms, 2 expand, 3 drop
    int64_t launches = 0;
    std::vector<kfsp_trace_row> trace;
    bool profile_spmv = false;        // kfsp_set_profiling: CUDA events around the launches of the time-stepping loop
    int profile_level = 0;            // 1: one event pair per Arnoldi sweep (class SWEEP; nothing between the sweep's launches, so
                                      //    programmatic dependent launch stays effective), 2: one pair per launch (per-class table)
    bool in_sweep = false;
    double spmv_seconds = 0.0;
    int64_t spmv_timed = 0;
    int64_t spmv_by_mode[3] = {0, 0, 0};    // generator SpMV launches of the last solve: plain, dot-fused, norm-fused
    int64_t spmv_fused = 0;                 // ... of which also finalised the previous column (lattice.cuh, FIN = 1)
    std::vector<cudaEvent_t> ev_pool;   // pairs of events bracketing the launches of the solve, by kernel class (profiling only)
    std::vector<int> ev_cls;
    size_t ev_used = 0;
    long ev_cur = -1;
    double prof_sec[KFSP_PROF_CLASSES] = {0};
    int64_t prof_cnt[KFSP_PROF_CLASSES] = {0};
    int64_t prof_bps[KFSP_PROF_CLASSES] = {0};   // algorithmic bytes per state, summed over the launches of the class
    std::vector<int> ev_bps, ev_n, ev_nspmv;
    // kernel classes of the time-stepping loop (kfsp_profile_get): CUDA events on the solver's stream around every launch
    // bps: bytes per state the launch must move (operands read once + results written once)
    int prof_begin(int cls, int bps = 0) {
        if (in_sweep && profile_level == 1) {              // inside a sweep bracket: only count
            sweep_launches += 1;
            sweep_bps += bps;
            if (cls <= KFSP_PROF_SPMV_FIN_NRM) sweep_spmv += 1;
            return KFSP_OK;
        }
        ev_cur = -1;
        if (!profile_spmv || ev_used + 2 > ev_pool.size()) return KFSP_OK;
        KFSP_CUDA(cudaEventRecord(ev_pool[ev_used], stream));
        ev_cls[ev_used / 2] = cls;
        ev_bps[ev_used / 2] = bps;
        ev_n[ev_used / 2] = 1;
        ev_nspmv[ev_used / 2] = cls <= KFSP_PROF_SPMV_FIN_NRM ? 1 : 0;
        ev_cur = (long)ev_used;
        return KFSP_OK;
    }
    int64_t sweep_launches = 0, sweep_bps = 0, sweep_spmv = 0;
    int sweep_begin() {
        if (!profile_spmv || profile_level != 1) return KFSP_OK;
        KFSP_TRY(prof_begin(KFSP_PROF_SWEEP, 0));
        in_sweep = true;
        sweep_launches = sweep_bps = sweep_spmv = 0;
        return KFSP_OK;
    }
    int sweep_end() {
        if (!in_sweep) return KFSP_OK;
        in_sweep = false;
        if (ev_cur >= 0) {
            ev_bps[ev_cur / 2] = (int)sweep_bps;
            ev_n[ev_cur / 2] = (int)sweep_launches;
            ev_nspmv[ev_cur / 2] = (int)sweep_spmv;
        }
        return prof_end();
    }
    int prof_end() {
        if (in_sweep && profile_level == 1) return KFSP_OK;
        if (ev_cur < 0) return KFSP_OK;
        KFSP_CUDA(cudaEventRecord(ev_pool[ev_cur + 1], stream));
        ev_used += 2;
        ev_cur = -1;
        return KFSP_OK;
    }
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;

    // ---------------------------------------------------------------- lifetime
    int init(const kfsp_options* o) {
        opt = *o;
        if (const char* ev = std::getenv("KFSP_SPMV_TUNE")) spmv_tune = std::atoi(ev);
        if (const char* ev = std::getenv("KFSP_BOX_TUNE")) box_tune = std::atoi(ev);
        if (const char* ev = std::getenv("KFSP_BD2_ZC")) bd2_zc = std::atoi(ev);
        if (const char* ev = std::getenv("KFSP_PDL")) use_pdl = std::atoi(ev) != 0;
        if (const char* ev = std::getenv("KFSP_BD2_AHEAD")) bd2_ahead = std::min(std::max(std::atoi(ev), 0), BD2_L2AHEAD);
        if (const char* ev = std::getenv("KFSP_BD2_SYNC")) { bd2_sync = std::atoi(ev); if (bd2_sync & (bd2_sync - 1)) bd2_sync = 0; }
        if (const char* ev = std::getenv("KFSP_SMALL_SWEEP")) small_sweep = std::atoi(ev) != 0;
        if (const char* ev = std::getenv("KFSP_SMEM_SWEEP")) smem_sweep = std::atoi(ev) != 0;
        if (const char* ev = std::getenv("KFSP_SSA_EMIT")) ssa_emit_on = std::atoi(ev) != 0;
        if (const char* ev = std::getenv("KFSP_CUSTOM_PROBE")) custom_probe_on = std::atoi(ev) != 0;
        if (const char* ev = std::getenv("KFSP_BLOCKING_SYNC")) blocking_sync = std::atoi(ev) != 0;
        if (const char* ev = std::getenv("KFSP_SSA_FAC")) ssa_fac_on = std::atoi(ev) != 0;
        if (const char* ev = std::getenv("KFSP_EXPM_SMALL_THREADS")) expm_small_threads = std::atoi(ev) != 0;
        if (const char* ev = std::getenv("KFSP_COOP_SWEEP")) coop_sweep = std::atoi(ev) != 0;
        if (const char* ev = std::getenv("KFSP_COOP_MAX_ROWS")) coop_max_rows = std::atoll(ev);
        if (const char* ev = std::getenv("KFSP_PROP_CACHE_STATES")) pc_budget = std::atoll(ev);
        if (const char* ev = std::getenv("KFSP_DEBUG_REPL")) repl_debug = std::atoi(ev) != 0;
        if (const char* ev = std::getenv("KFSP_REPL_MIN_ROWS")) dist.repl_min_rows = std::atoll(ev);
        if (opt.spmv_variant < 0 || opt.spmv_variant > 2) return KFSP_ERR_ARG;
        if (opt.m_max < opt.m_min || opt.m_min < 1 || opt.m_max > EXPM_MAXN - 4 || opt.ideg != 6 || opt.max_states < 2 ||
            opt.max_states > 2000000000LL)
            return KFSP_ERR_ARG;
        int count = 0;
        if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return KFSP_ERR_NO_DEVICE;
        if (opt.device >= 0) {
            if (opt.device >= count) return KFSP_ERR_NO_DEVICE;
            device = opt.device;
        } else {
            KFSP_CUDA(cudaGetDevice(&device));
        }
        KFSP_CUDA(cudaSetDevice(device));
        KFSP_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));
        KFSP_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        KFSP_CUDA(cudaStreamCreateWithFlags(&stream2, cudaStreamNonBlocking));
        KFSP_CUDA(cudaEventCreateWithFlags(&ev_side, cudaEventDisableTiming));
        KFSP_CUDA(cudaMalloc(&d_chk, sizeof(int32_t)));
        KFSP_CUDA(cudaEventCreate(&ev_a));
        KFSP_CUDA(cudaEventCreate(&ev_b));
        KFSP_CUDA(cudaMalloc(&d_model, sizeof(DeviceModel)));
        KFSP_CUDA(cudaMalloc(&d_err, sizeof(int32_t)));
        KFSP_CUDA(cudaMemset(d_err, 0, sizeof(int32_t)));
        LDH = opt.m_max + 2;
        KFSP_CUDA(cudaMalloc(&d_H, sizeof(double) * LDH * LDH));
        KFSP_CUDA(cudaMalloc(&d_expm_work, sizeof(double) * 7 * EXPM_MAXN * EXPM_MAXN));
        KFSP_CUDA(cudaMalloc(&d_expm_full, sizeof(double) * EXPM_MAXN * EXPM_MAXN));
        KFSP_CUDA(cudaMalloc(&d_res, sizeof(ExpmResult)));
        KFSP_CUDA(cudaMallocHost(&h_res, sizeof(ExpmResult)));
        KFSP_CUDA(cudaMalloc(&d_ctl, sizeof(SweepCtl)));
        KFSP_CUDA(cudaMemset(d_ctl, 0, sizeof(SweepCtl)));
        KFSP_CUDA(cudaMallocHost(&h_ctl, sizeof(SweepCtl)));
        KFSP_CUDA(cudaMalloc(&rd.partials, sizeof(double) * RED_W * MAX_VEC_BLOCKS));
        KFSP_CUDA(cudaMalloc(&rd.counter, sizeof(unsigned int)));
        KFSP_CUDA(cudaMemset(rd.counter, 0, sizeof(unsigned int)));
        KFSP_CUDA(cudaFuncSetAttribute(k_expm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EXPM_SMEM));
        KFSP_CUDA(cudaFuncSetAttribute(k_expm_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EXPM_SMEM));
        if (const char* ev = std::getenv("KFSP_EXPM_CLUSTER")) expm_cluster_from = std::atoi(ev);
        return KFSP_OK;
    }
    void destroy() {
        cudaSetDevice(device);
        if (repl_debug) std::fprintf(stderr, "libkfsp rank %d: %g gathers of W %.3f s, re-partitions %.3f s\n", dist.rank, repl_s[2], repl_s[0], repl_s[1]);
        if (stream) cudaStreamSynchronize(stream);
        if (ev_block) cudaEventDestroy(ev_block);
        if (stream2) { cudaStreamSynchronize(stream2); cudaStreamDestroy(stream2); }
        if (ev_side) cudaEventDestroy(ev_side);
        cudaFree(d_chk);
        free_state_space();
        free_prop_cache();
        cudaFree(d_model); cudaFree(d_tables); cudaFree(d_err); cudaFree(d_H); cudaFree(d_expm_work); cudaFree(d_expm_full); cudaFree(d_res);
        cudaFree(d_ctl); cudaFree(rd.partials); cudaFree(rd.counter); cudaFree(d_scratch); cudaFree(d_flush); cudaFree(hp_dev);
        cudaFree(emit.tmp); cudaFree(emit.prev); cudaFree(emit.head); cudaFree(emit.cursor);
        cudaFree(coop.bar); cudaFree(coop.part);
        cudaFree(d_factabs);
        if (hp_host) cudaFreeHost(hp_host);
        if (h_res) cudaFreeHost(h_res);
        if (h_ctl) cudaFreeHost(h_ctl);
        for (auto& e : ev_pool) cudaEventDestroy(e);
        if (ev_a) cudaEventDestroy(ev_a);
        if (ev_b) cudaEventDestroy(ev_b);
        if (stream) cudaStreamDestroy(stream);
    }
    void free_state_space() {
        cudaFree(d_states); cudaFree(d_succ); cudaFree(d_prop); cudaFree(d_diag); cudaFree(d_pred); cudaFree(d_coef);
        cudaFree(d_w); cudaFree(d_table); cudaFree(d_V); cudaFree(dist.send_idx); cudaFree(dist.sendbuf); cudaFree(dist.halo);
        dist.send_idx = nullptr; dist.sendbuf = nullptr; dist.halo = nullptr;
        d_states = d_succ = d_pred = d_table = nullptr;
        d_prop = d_diag = d_coef = d_w = d_V = nullptr;
        ld = 0; n = 0;
        box = false;
    }

    // Grid for a grid-stride kernel: one full wave of resident CTAs (SMs x CTAs/SM for THIS kernel's
    // register/shared-memory footprint), so no ragged second wave; capped by MAX_VEC_BLOCKS (reducer arrays).
    int num_sms = 148;
    std::vector<std::pair<const void*, int>> occ_cache;
    int wave_grid(const void* kernel, int64_t work, int threads = VEC_THREADS, size_t dyn_smem = 0) {
        int per_sm = 0;
        for (auto& pr : occ_cache) if (pr.first == kernel) { per_sm = pr.second; break; }
        if (per_sm == 0) {
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem) != cudaSuccess || per_sm < 1) per_sm = 1;
            occ_cache.emplace_back(kernel, per_sm);
        }
        int64_t b = (work + threads - 1) / threads;
        const int64_t cap = std::min<int64_t>((int64_t)num_sms * per_sm, MAX_VEC_BLOCKS);
        if (b > cap) b = cap;
        if (b < 1) b = 1;
        return (int)b;
    }
    int grid_for(int64_t work, int threads = VEC_THREADS) const {
        int64_t b = (work + threads - 1) / threads;
        if (b < 1) b = 1;
        if (b > MAX_VEC_BLOCKS) b = MAX_VEC_BLOCKS;
        return (int)b;
    }
    // Host wait for the solver's stream.  Default: cudaStreamSynchronize (the driver spins: lowest latency, one core per waiting
    // thread).  blocking_sync (kfsp_set_blocking_sync, KFSP_BLOCKING_SYNC=1): the thread sleeps on an event created with
    // cudaEventBlockingSync -- for many handles solving concurrently from more host threads than cores (sweep.py: SweepPool).
    bool blocking_sync = false;
    cudaEvent_t ev_block = nullptr;
    cudaError_t wait_stream() {
        if (!blocking_sync) return cudaStreamSynchronize(stream);
        if (!ev_block) {
            const cudaError_t ce = cudaEventCreateWithFlags(&ev_block, cudaEventBlockingSync | cudaEventDisableTiming);
            if (ce != cudaSuccess) return ce;
        }
        const cudaError_t ce = cudaEventRecord(ev_block, stream);
        if (ce != cudaSuccess) return ce;
        return cudaEventSynchronize(ev_block);
    }
    int sync() {
        KFSP_CUDA(wait_stream());
        return KFSP_OK;
    }
    // rows the Krylov kernels of this rank work on: everything, or this rank's slice of the replicated layout
    bool sliced() const { return dist.repl && !dist.suspended && !dist.whole; }
    bool dist_active() const { return dist.nranks > 1 && !dist.suspended && !dist.whole; }     // kernels exchange with the other ranks
    int64_t kr0() const { return sliced() ? dist.lo : 0; }
    int64_t kn() const { return sliced() ? dist.hi - dist.lo : n; }
    int check_launch() {
        ++launches;
        KFSP_CUDA(cudaGetLastError());
        return KFSP_OK;
    }
#define KFSP_LAUNCH(kernel, grid, block, smem, ...)            \
    do {                                                       \
        kernel<<<(grid), (block), (smem), stream>>>(__VA_ARGS__); \
        KFSP_TRY(check_launch());                              \
    } while (0)

    // k_ssa_walk with the number of reactions at compile time for the shipped model shapes (R = 4, 6, 10), any R otherwise
#define KFSP_LAUNCH_SSA(FILL, grid, block, smem, ...)                                        \
    do {                                                                                     \
        switch (R) {                                                                         \
        case 4: k_ssa_walk<FILL, 4><<<(grid), (block), (smem), stream>>>(__VA_ARGS__); break;  \
        case 6: k_ssa_walk<FILL, 6><<<(grid), (block), (smem), stream>>>(__VA_ARGS__); break;  \
        case 10: k_ssa_walk<FILL, 10><<<(grid), (block), (smem), stream>>>(__VA_ARGS__); break; \
        default: k_ssa_walk<FILL, 0><<<(grid), (block), (smem), stream>>>(__VA_ARGS__); break; \
        }                                                                                    \
        KFSP_TRY(check_launch());                                                            \
    } while (0)
    // Launch with programmatic stream serialization (KFSP_PDL=0 switches it off): the kernel may be scheduled while its
    // predecessor drains; it calls pdl_wait() before touching anything the predecessor wrote (krylov.cuh).
    bool use_pdl = true;
    template <typename... KArgs, typename... Args>
    int launch_pdl(void (*kernel)(KArgs...), int grid, int block, size_t smem, Args&&... args) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid, 1, 1);
        cfg.blockDim = dim3((unsigned)block, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = use_pdl ? 1 : 0;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        KFSP_CUDA(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
        return check_launch();
    }

    int ensure_scratch(size_t bytes) {
        if (bytes <= scratch_bytes) return KFSP_OK;
        KFSP_CUDA(wait_stream());
        if (d_scratch) KFSP_CUDA(cudaFree(d_scratch));
        size_t want = std::max(bytes, scratch_bytes * 2);
        want = (want + 255) & ~(size_t)255;
        KFSP_CUDA(cudaMalloc(&d_scratch, want));
        scratch_bytes = want;
        return KFSP_OK;
    }

    // ---------------------------------------------------------------- model
    int set_model(const HostModel& m) {
        if (m.S < 1 || m.S > KFSP_MAX_SPECIES || m.R < 1 || m.R > KFSP_MAX_REACTIONS || m.P > KFSP_MAX_PARAMS || m.P < 0)
            return KFSP_ERR_UNSUPPORTED;
        KFSP_CUDA(cudaSetDevice(device));
        if (m.custom) return set_model_hostprop(m);
        DeviceModel dm;
        std::memset(&dm, 0, sizeof dm);
        dm.S = m.S; dm.R = m.R; dm.P = m.P; dm.max_molecules = opt.max_molecules;
        for (int k = 0; k < m.R; ++k)
            for (int s = 0; s < m.S; ++s) dm.stoich[k * m.S + s] = m.stoich[(size_t)k * m.S + s];
        int nc = 0, ni = 0;
        if ((int)m.programs.size() != m.R) return KFSP_ERR_NO_MODEL;
        {   // multi-species transcendental programs: host evaluation keeps the results bit-identical to the host's
            int inexact_multi = 0;
            for (int k = 0; k < m.R; ++k) {
                if (m.programs[k].empty()) return KFSP_ERR_NO_MODEL;
                uint32_t mask = 0; bool ix = false;
                program_profile(m.programs[k], m.S, &mask, &ix);
                if (ix && (mask & (mask - 1)) != 0) ++inexact_multi;
            }
            const char* ev = std::getenv("KFSP_DEVICE_MATH");
            if (inexact_multi > 0 && !(ev && std::atoi(ev) != 0)) {
                const int st = set_model_hostprop(m);
                n_inexact_on_device = 0;
                n_host_evaluated = inexact_multi;
                return st;
            }
        }
        for (int k = 0; k < m.R; ++k) {
            const Program& p = m.programs[k];
            if (p.empty()) return KFSP_ERR_NO_MODEL;
            const int depth = program_stack_depth(p, m.S + m.P);
            if (depth < 0 || depth > KFSP_STACK) return KFSP_ERR_UNSUPPORTED;
            if (nc + (int)p.code.size() > KFSP_MAX_CODE || ni + (int)p.immed.size() > KFSP_MAX_IMMED) return KFSP_ERR_UNSUPPORTED;
            dm.code_begin[k] = nc; dm.immed_begin[k] = ni;
            for (int32_t c : p.code) dm.code[nc++] = c;
            for (double v : p.immed) dm.immed[ni++] = v;
        }
        dm.code_begin[m.R] = nc; dm.immed_begin[m.R] = ni;
        for (int i = 0; i < m.P; ++i) dm.params[i] = m.params[i];
        // tabulate single-species propensities that contain a transcendental operation
        n_tabulated = 0; n_inexact_on_device = 0; n_host_evaluated = 0;
        const int64_t tlen = (int64_t)opt.max_molecules + 1;
        std::vector<int> tab_k;
        for (int k = 0; k < m.R; ++k) {
            uint32_t mask = 0; bool ix = false;
            program_profile(m.programs[k], m.S, &mask, &ix);
            dm.table_species[k] = -1;
            dm.table[k] = nullptr;
            if (!ix && opt.spmv_variant != 1) continue;       // the lattice SpMV reads every propensity from a table
            if ((mask & (mask - 1)) == 0) {                  // zero or one species
                int sp = 0;
                while (mask > 1) { mask >>= 1; ++sp; }
                dm.table_species[k] = sp;
                tab_k.push_back(k);
            } else if (ix) {
                ++n_inexact_on_device;                        // evaluated with the CUDA math library: may differ by ulps
            }
        }
        if (d_tables) { KFSP_CUDA(cudaFree(d_tables)); d_tables = nullptr; }
        if (!tab_k.empty()) {
            std::vector<double> host((size_t)tlen * tab_k.size());
            std::vector<double> val((size_t)m.S + m.P, 0.0);
            for (int i = 0; i < m.P; ++i) val[m.S + i] = m.params[i];
            for (size_t q = 0; q < tab_k.size(); ++q) {
                const int k = tab_k[q], sp = dm.table_species[k];
                for (int64_t c = 0; c < tlen; ++c) {
                    for (int s2 = 0; s2 < m.S; ++s2) val[s2] = 0.0;
                    val[sp] = (double)c;
                    host[q * tlen + c] = evaluate_program(m.programs[k], val.data());
                }
            }
            KFSP_CUDA(cudaMalloc(&d_tables, sizeof(double) * host.size()));
            KFSP_CUDA(cudaMemcpy(d_tables, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice));
            for (size_t q = 0; q < tab_k.size(); ++q) dm.table[tab_k[q]] = d_tables + q * tlen;
            n_tabulated = (int)tab_k.size();
        }
        {   // factored single-species tables: required by the index-only SpMV, used by the SSA walks of every variant when the
            // model has such a form (KFSP_SSA_FAC=0: byte-code interpreter, A/B)
            fac_ok = false;
            const int fst = build_factored(m);
            if (fst == KFSP_OK) fac_ok = true;
            else if (opt.spmv_variant == 2) return fst;
        }
        KFSP_CUDA(cudaMemcpyAsync(d_model, &dm, sizeof dm, cudaMemcpyHostToDevice, stream));
        KFSP_CUDA(wait_stream());
        h_dm = dm;
        const bool reshape = !have_model || m.S != S || m.R != R || box;
        S = m.S; R = m.R;
        have_model = true;
        host_prop = false;
        free_prop_cache();
        if (reshape) { free_state_space(); }
        n = 0;
        return KFSP_OK;
    }

    // spmv_variant = 2: factor every propensity into single-species terms (model_host.h: factor_program), tabulate the terms
    // on the host over the counts 0..max_molecules and hand the tables to the index-only SpMV.  A model whose propensities
    // combine several species by anything but + - * (or whose terms hit a division by zero / domain error inside a
    // multi-term propensity) cannot take this variant: KFSP_ERR_UNSUPPORTED, never a silent switch.
    int build_factored(const HostModel& m) {
        FacModel F;
        std::memset(&F, 0, sizeof F);
        F.S = m.S; F.R = m.R;
        const int64_t tlen = (int64_t)opt.max_molecules + 1;
        std::vector<double> host;
        std::vector<size_t> off;                         // start of table (k,t) in `host`, row-major over [k][t]
        std::vector<double> val((size_t)m.S + m.P, 0.0);
        for (int i = 0; i < m.P; ++i) val[m.S + i] = m.params[i];
        for (int k = 0; k < m.R; ++k) {
            Factored fp;
            if (!factor_program(m.programs[k], m.S, fp)) return KFSP_ERR_UNSUPPORTED;
            if (fp.terms.empty() || (int)fp.terms.size() > FAC_MAX_TERMS || (int)fp.ops.size() > FAC_MAX_OPS) return KFSP_ERR_UNSUPPORTED;
            const bool multi = fp.terms.size() > 1;
            for (size_t t = 0; t < fp.terms.size(); ++t) {
                const FactoredTerm& tm = fp.terms[t];
                const int64_t len = tm.species >= 0 ? tlen : 1;
                off.push_back(host.size());
                for (int64_t c = 0; c < len; ++c) {
                    for (int s2 = 0; s2 < m.S; ++s2) val[s2] = 0.0;
                    if (tm.species >= 0) val[tm.species] = (double)c;
                    bool ab = false;
                    host.push_back(evaluate_program_checked(tm.prog, val.data(), &ab));
                    if (ab && multi) return KFSP_ERR_UNSUPPORTED;      // the whole propensity would be 0 there, not the term
                }
                F.sp[k][t] = (int8_t)(tm.species >= 0 ? tm.species : 0);
                F.use[k][t] = tm.species >= 0 ? 1 : 0;
            }
            for (size_t t = fp.terms.size(); t < (size_t)FAC_MAX_TERMS; ++t) off.push_back(host.size());
            F.nops[k] = (int8_t)fp.ops.size();
            for (size_t q = 0; q < fp.ops.size(); ++q) F.ops[k][q] = (int8_t)fp.ops[q];
            const std::vector<int32_t>& o = fp.ops;
            if (o.size() == 1) F.shape[k] = FAC_ONE;
            else if (o.size() == 3 && o[0] == 0 && o[1] == 1 && o[2] == -cMul) F.shape[k] = FAC_MUL2;
            else if (o.size() == 5 && o[0] == 0 && o[1] == 1 && o[2] == -cMul && o[3] == 2 && o[4] == -cMul) F.shape[k] = FAC_MUL3;
            else F.shape[k] = FAC_GEN;
            for (int s2 = 0; s2 < m.S; ++s2) {
                const int32_t v = m.stoich[(size_t)k * m.S + s2];
                if (v < -127 || v > 127) return KFSP_ERR_UNSUPPORTED;
                F.nu[k][s2] = (int8_t)v;
            }
        }
        if (d_factabs) { KFSP_CUDA(cudaFree(d_factabs)); d_factabs = nullptr; }
        KFSP_CUDA(cudaMalloc(&d_factabs, sizeof(double) * std::max<size_t>(host.size(), 1)));
        KFSP_CUDA(cudaMemcpy(d_factabs, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice));
        for (int k = 0; k < m.R; ++k)
            for (int t = 0; t < FAC_MAX_TERMS; ++t) F.tab[k][t] = d_factabs + off[(size_t)k * FAC_MAX_TERMS + t];
        fac = F;
        idx = opt.spmv_variant == 2;
        return KFSP_OK;
    }

    // CUSTOMPROP: the propensity is an opaque host function.  Sizes, stoichiometry and parameters go to the
    // device; a_k(x) is evaluated on the host in batches (propensities_host) and, inside SSA walks, served
    // from the device side cache (fsp_ssa_hostprop).
    // A callback whose reactions each read at most one species (examples/toggle.f90:55-69) or are bilinear mass action
    // c * x_a * x_b (examples/transcr6d.f90:74,78) is recognised by probing (model_host.h: probe_custom): single-species reactions are
    // tabulated over the counts 0..max_molecules with the callback's own values, bilinear ones become the three-operation byte code
    // (c * X_a) * X_b, and everything is verified bit for bit against the callback on random states.  Such a model is then an
    // ordinary device model: no host round trip, every SpMV variant, every multi-GPU layout.  KFSP_CUSTOM_PROBE=0 keeps the
    // callbacks (A/B).  Anything else stays on the host path: explicit matrix, one GPU or the replicated multi-GPU layout (every
    // rank calls its own copy of the function).
    bool custom_probe_on = true;
    int set_model_custom_device(const HostModel& m, const CustomProbe& pr) {
        DeviceModel dm;
        std::memset(&dm, 0, sizeof dm);
        dm.S = m.S; dm.R = m.R; dm.P = m.P; dm.max_molecules = opt.max_molecules;
        for (int k = 0; k < m.R; ++k)
            for (int s = 0; s < m.S; ++s) {
                const int32_t v = m.stoich[(size_t)k * m.S + s];
                if (v < -127 || v > 127) return KFSP_ERR_UNSUPPORTED;
                dm.stoich[k * m.S + s] = v;
            }
        for (int i = 0; i < m.P; ++i) dm.params[i] = m.params[i];
        const int64_t tlen = (int64_t)opt.max_molecules + 1;
        // device tables: [k] the callback's values of a single-species reaction; then, per bilinear reaction, fl(c * count) for
        // the factored form; last the identity table (double)count that is the second factor of every bilinear reaction
        std::vector<double> host(pr.tables);
        std::vector<size_t> t1((size_t)m.R, 0);
        int nbil = 0;
        for (int k = 0; k < m.R; ++k)
            if (pr.species[k] < 0) {
                t1[k] = host.size();
                for (int64_t c = 0; c < tlen; ++c) host.push_back(pr.coef[k] * (double)c);
                ++nbil;
            }
        const size_t ident = host.size();
        if (nbil > 0)
            for (int64_t c = 0; c < tlen; ++c) host.push_back((double)c);
        if (d_tables) { KFSP_CUDA(cudaFree(d_tables)); d_tables = nullptr; }
        KFSP_CUDA(cudaMalloc(&d_tables, sizeof(double) * host.size()));
        KFSP_CUDA(cudaMemcpy(d_tables, host.data(), sizeof(double) * host.size(), cudaMemcpyHostToDevice));
        FacModel F;
        std::memset(&F, 0, sizeof F);
        F.S = m.S; F.R = m.R;
        int nc = 0, ni = 0;
        for (int k = 0; k < m.R; ++k) {
            dm.code_begin[k] = nc; dm.immed_begin[k] = ni;
            for (int s2 = 0; s2 < m.S; ++s2) F.nu[k][s2] = (int8_t)m.stoich[(size_t)k * m.S + s2];
            for (int t = 0; t < FAC_MAX_TERMS; ++t) F.tab[k][t] = d_tables;
            if (pr.species[k] >= 0) {
                dm.table_species[k] = pr.species[k];
                dm.table[k] = d_tables + (size_t)k * tlen;
                F.shape[k] = FAC_ONE; F.nops[k] = 1; F.ops[k][0] = 0;
                F.sp[k][0] = (int8_t)pr.species[k]; F.use[k][0] = 1;
                F.tab[k][0] = dm.table[k];
            } else {
                // (c * X_a) * X_b in the parser's byte code (FortranParser.f90:52-73): the device interpreter evaluates it with
                // the very roundings the probe verified against the callback
                if (nc + 5 > KFSP_MAX_CODE || ni + 1 > KFSP_MAX_IMMED) return KFSP_ERR_UNSUPPORTED;
                dm.table_species[k] = -1; dm.table[k] = nullptr;
                dm.code[nc++] = cImmed; dm.code[nc++] = VarBegin + pr.sa[k]; dm.code[nc++] = cMul;
                dm.code[nc++] = VarBegin + pr.sb[k]; dm.code[nc++] = cMul;
                dm.immed[ni++] = pr.coef[k];
                F.shape[k] = FAC_MUL2; F.nops[k] = 3; F.ops[k][0] = 0; F.ops[k][1] = 1; F.ops[k][2] = -cMul;
                F.sp[k][0] = (int8_t)pr.sa[k]; F.use[k][0] = 1; F.tab[k][0] = d_tables + t1[k];
                F.sp[k][1] = (int8_t)pr.sb[k]; F.use[k][1] = 1; F.tab[k][1] = d_tables + ident;
            }
        }
        dm.code_begin[m.R] = nc; dm.immed_begin[m.R] = ni;
        n_tabulated = m.R - nbil; n_inexact_on_device = 0; n_host_evaluated = 0;
        fac = F;
        fac_ok = true;
        idx = opt.spmv_variant == 2;
        KFSP_CUDA(cudaMemcpyAsync(d_model, &dm, sizeof dm, cudaMemcpyHostToDevice, stream));
        KFSP_CUDA(cudaStreamSynchronize(stream));
        h_dm = dm;
        const bool reshape = !have_model || m.S != S || m.R != R || box;
        S = m.S; R = m.R;
        have_model = true;
        host_prop = false;
        free_prop_cache();
        if (reshape) { free_state_space(); }
        n = 0;
        return KFSP_OK;
    }
    int set_model_hostprop(const HostModel& m) {
        if (m.custom && custom_probe_on) {
            CustomProbe pr;
            if (probe_custom(m, opt.max_molecules, pr)) return set_model_custom_device(m, pr);
        }
        // rows of a memory-scaled partition, the lattice and the index-only variant are built from device byte code / tables
        if ((dist.nranks > 1 && !dist.repl) || opt.spmv_variant != 0) return KFSP_ERR_UNSUPPORTED;
        DeviceModel dm;
        std::memset(&dm, 0, sizeof dm);
        dm.S = m.S; dm.R = m.R; dm.P = m.P; dm.max_molecules = opt.max_molecules;
        for (int k = 0; k < m.R; ++k)
            for (int s = 0; s < m.S; ++s) dm.stoich[k * m.S + s] = m.stoich[(size_t)k * m.S + s];
        for (int i = 0; i < m.P; ++i) dm.params[i] = m.params[i];
        for (int k = 0; k < m.R; ++k) { dm.table_species[k] = -1; dm.table[k] = nullptr; }
        n_tabulated = 0; n_inexact_on_device = 0; n_host_evaluated = m.custom ? m.R : 0;
        KFSP_CUDA(cudaMemcpyAsync(d_model, &dm, sizeof dm, cudaMemcpyHostToDevice, stream));
        KFSP_CUDA(wait_stream());
        const bool reshape = !have_model || m.S != S || m.R != R;
        S = m.S; R = m.R;
        have_model = true;
        host_prop = true;
        fac_ok = false;
        hm = m;
        free_prop_cache();
        if (reshape) { free_state_space(); }
        n = 0;
        return KFSP_OK;
    }

    // ---------------------------------------------------------------- host-evaluated propensities
    void free_prop_cache() {
        cudaFree(pc.states); cudaFree(pc.table); cudaFree(pc.prop); cudaFree(pc.req); cudaFree(pc.nreq);
        pc = PropCache();
        pc_n = 0;
    }
    // staging buffers of the host-propensity path: allocated once and grown geometrically (a cudaMalloc/cudaFree
    // pair per expansion cost more than the callbacks themselves)
    int ensure_hp(size_t dev_bytes, size_t host_bytes) {
        if (dev_bytes > hp_dev_bytes) {
            KFSP_CUDA(wait_stream());
            if (hp_dev) KFSP_CUDA(cudaFree(hp_dev));
            hp_dev_bytes = align_up(std::max(dev_bytes, 2 * hp_dev_bytes));
            KFSP_CUDA(cudaMalloc(&hp_dev, hp_dev_bytes));
        }
        if (host_bytes > hp_host_bytes) {
            KFSP_CUDA(wait_stream());
            if (hp_host) KFSP_CUDA(cudaFreeHost(hp_host));
            hp_host_bytes = align_up(std::max(host_bytes, 2 * hp_host_bytes));
            KFSP_CUDA(cudaMallocHost(&hp_host, hp_host_bytes));
        }
        return KFSP_OK;
    }
    // capacity for `want` cached states (contents are preserved) and a request list of req_cap states
    int ensure_prop_cache(int64_t want) {
        if (pc.table && want <= pc.ld) return KFSP_OK;
        int64_t cap = std::max<int64_t>(1 << 16, pc.ld);
        while (cap < want) cap <<= 1;
        PropCache q;
        q.ld = cap;
        int64_t ts = 1024;
        while (ts < 2 * cap) ts <<= 1;
        q.mask = (uint32_t)(ts - 1);
        q.req_cap = 1 << 16;
        KFSP_CUDA(cudaMalloc(&q.states, sizeof(int32_t) * cap * S));
        KFSP_CUDA(cudaMalloc(&q.table, sizeof(int32_t) * ts));
        KFSP_CUDA(cudaMalloc(&q.prop, sizeof(double) * cap * (R + 1)));
        KFSP_CUDA(cudaMalloc(&q.req, sizeof(int32_t) * (size_t)q.req_cap * S));
        KFSP_CUDA(cudaMalloc(&q.nreq, sizeof(int32_t)));
        KFSP_CUDA(cudaMemsetAsync(q.table, 0xFF, sizeof(int32_t) * ts, stream));
        KFSP_CUDA(cudaMemsetAsync(q.nreq, 0, sizeof(int32_t), stream));
        if (pc_n > 0) {
            KFSP_CUDA(cudaMemcpyAsync(q.states, pc.states, sizeof(int32_t) * pc_n * S, cudaMemcpyDeviceToDevice, stream));
            KFSP_CUDA(cudaMemcpyAsync(q.prop, pc.prop, sizeof(double) * pc_n * (R + 1), cudaMemcpyDeviceToDevice, stream));
            KFSP_LAUNCH(k_cache_insert, grid_for(pc_n), VEC_THREADS, 0, q, S, (int64_t)0, pc_n, d_err);
        }
        KFSP_CUDA(wait_stream());
        const int64_t keep = pc_n;
        free_prop_cache();
        pc = q;
        pc_n = keep;
        return KFSP_OK;
    }
    int clear_prop_cache() {
        if (!pc.table) return KFSP_OK;
        if (pc_n > 0) KFSP_CUDA(cudaMemsetAsync(pc.table, 0xFF, sizeof(int32_t) * ((size_t)pc.mask + 1), stream));
        pc_n = 0;
        return KFSP_OK;
    }
    // R propensities + their sum (reaction order, StateSpace.f90:207-212) of `cnt` states: out[t*(R+1) + k]
    void eval_host(const int32_t* st, int64_t cnt, double* out) {
        const double t0 = wall_now();
        for (int64_t t = 0; t < cnt; ++t) {
            double d = 0.0;
            for (int k = 0; k < R; ++k) {
                const double a = hm.propensity(st + t * S, k + 1);
                out[t * (R + 1) + k] = a;
                d = d + a;
            }
            out[t * (R + 1) + R] = d;
        }
        host_prop_evals += cnt * R;
        phase_s[5] += wall_now() - t0;
    }
    // OFFDIAG/DIAG of states [first, first+count): from the side cache where the SSA walks already paid for
    // them, by the host function for the rest.
    int propensities_host(int64_t first, int64_t count) {
        if (count < 1) return KFSP_OK;
        const size_t a_i = align_up(sizeof(int32_t) * count), a_s = align_up(sizeof(int32_t) * count * S),
                     a_v = align_up(sizeof(double) * count * (R + 1));
        KFSP_TRY(ensure_hp(a_i + a_v, a_i + a_s + a_v));
        int32_t* d_q = (int32_t*)hp_dev;                 // cache entry per new state, later reused as the index list
        double* d_vals = (double*)(hp_dev + a_i);
        int32_t* h_q = (int32_t*)hp_host;
        int32_t* h_st = (int32_t*)(hp_host + a_i);
        double* h_vals = (double*)(hp_host + a_i + a_s);
        FspView f = view();
        const bool cached = pc.table && pc_n > 0;
        if (cached) {
            KFSP_LAUNCH(k_props_from_cache, grid_for(count), VEC_THREADS, 0, f, first, count, pc, d_q);
            KFSP_CUDA(cudaMemcpyAsync(h_q, d_q, sizeof(int32_t) * count, cudaMemcpyDeviceToHost, stream));
        }
        KFSP_CUDA(cudaMemcpyAsync(h_st, d_states + first * S, sizeof(int32_t) * count * S, cudaMemcpyDeviceToHost, stream));
        KFSP_CUDA(wait_stream());
        // compact the states the cache did not serve to the front of h_st; h_q becomes their row indices
        int64_t nm = 0;
        for (int64_t t = 0; t < count; ++t) {
            if (cached && h_q[t] >= 0) continue;
            if (nm != t) std::memcpy(h_st + nm * S, h_st + t * S, sizeof(int32_t) * S);
            h_q[nm++] = (int32_t)(first + t);
        }
        if (nm == 0) return KFSP_OK;
        eval_host(h_st, nm, h_vals);
        KFSP_CUDA(cudaMemcpyAsync(d_q, h_q, sizeof(int32_t) * nm, cudaMemcpyHostToDevice, stream));
        KFSP_CUDA(cudaMemcpyAsync(d_vals, h_vals, sizeof(double) * nm * (R + 1), cudaMemcpyHostToDevice, stream));
        KFSP_LAUNCH(k_scatter_props, grid_for(nm * (R + 1)), VEC_THREADS, 0, f, (const int32_t*)d_q, (const double*)d_vals, nm);
        KFSP_CUDA(wait_stream());         // the staging buffers are reused by the next call
        return KFSP_OK;
    }
    int propensities(int64_t first, int64_t count) {
        if (host_prop) return propensities_host(first, count);
        FspView f = view();
        KFSP_LAUNCH(k_propensities, grid_for(count), VEC_THREADS, 0, f, first, count);
        return KFSP_OK;
    }

    int ensure_state_space() {
        if (!have_model) return KFSP_ERR_NO_MODEL;
        if (ld > 0) return KFSP_OK;
        KFSP_CUDA(cudaSetDevice(device));
        const int64_t gcap = ((opt.max_states + 63) / 64) * 64;
        // row arrays (matrix, vectors, basis) hold only this rank's rows when the state space is partitioned
        const int64_t cap = (dist.nranks > 1 && !dist.repl) ? (((opt.max_states + dist.nranks - 1) / dist.nranks + 1 + 63) / 64) * 64 : gcap;
        int64_t ts = 1024;
        while (ts < 2 * gcap) ts <<= 1;
        KFSP_CUDA(cudaMalloc(&d_states, sizeof(int32_t) * gcap * S));
        KFSP_CUDA(cudaMalloc(&d_succ, sizeof(int32_t) * cap * R));
        KFSP_CUDA(cudaMalloc(&d_pred, sizeof(int32_t) * cap * R));
        KFSP_CUDA(cudaMalloc(&d_prop, sizeof(double) * cap * R));
        if (opt.spmv_variant != 2) KFSP_CUDA(cudaMalloc(&d_coef, sizeof(double) * cap * R));      // index-only variant: recomputed
        KFSP_CUDA(cudaMalloc(&d_diag, sizeof(double) * cap));
        KFSP_CUDA(cudaMalloc(&d_w, sizeof(double) * cap));
        KFSP_CUDA(cudaMalloc(&d_table, sizeof(int32_t) * ts));
        KFSP_CUDA(cudaMemsetAsync(d_w, 0, sizeof(double) * cap, stream));
        states_cap = gcap;
        table_size = ts;
        ld = cap;
        n = 0;
        return KFSP_OK;
    }
    int ensure_basis() {
        if (d_V) return KFSP_OK;
        // m_max+2 basis columns, the two scratch columns of the fused lattice sweep (lattice.cuh; the first one also stages
        // operands for the peer-memory halo), and slack for its L2 prefetches
        const size_t slack = box ? (size_t)(BD2_L2AHEAD + 8) * (size_t)lat.B[0] : 0;
        KFSP_CUDA(cudaMalloc(&d_V, sizeof(double) * ((size_t)ld * (opt.m_max + 4) + slack)));
        if (dist.nranks > 1) KFSP_TRY(dist_setup_p2p());       // collective: every rank allocates its basis here
        return KFSP_OK;
    }
    FspView view() const {
        FspView f;
        f.S = S; f.R = R; f.ld = ld; f.n = n;
        f.states = d_states; f.succ = d_succ; f.prop = d_prop; f.diag = d_diag; f.pred = d_pred; f.coef = d_coef;
        f.table = d_table; f.mask = (uint32_t)(table_size - 1); f.model = d_model;
        return f;
    }
    int read_err(int32_t* e) {
        KFSP_CUDA(cudaMemcpyAsync(e, d_err, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        KFSP_CUDA(wait_stream());
        if (*e) KFSP_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int32_t), stream));
        return KFSP_OK;
    }
    static int err_to_status(int32_t e) {
        if (e & DEV_BAD_STATE) return KFSP_ERR_BAD_STATE;
        if (e & DEV_DUPLICATE) return KFSP_ERR_BAD_STATE;
        if (e & DEV_MOLECULE_LIMIT) return KFSP_ERR_MOLECULE_LIMIT;
        if (e & DEV_RUNAWAY) return KFSP_ERR_SSA_RUNAWAY;
        if (e & DEV_TABLE_FULL) return KFSP_ERR_OVERFLOW;
        return KFSP_OK;
    }

    // exclusive scan of cnt int32 values (any cnt < 2^31): tile scans, recursive scan of the tile sums.
    // tile_buf must hold scan_buf_ints(cnt) ints.
    static int64_t scan_buf_ints(int64_t cnt) {
        int64_t total = 0;
        while (cnt > SCAN_TILE) {
            const int64_t tiles = (cnt + SCAN_TILE - 1) / SCAN_TILE;
            total += 2 * tiles + 2;
            cnt = tiles;
        }
        return total + 4;
    }
    int scan_device(const int32_t* in, int32_t* out, int64_t cnt, int32_t* tile_buf) {
        const int64_t tiles = (cnt + SCAN_TILE - 1) / SCAN_TILE;
        if (tiles <= 1) {
            KFSP_LAUNCH(k_scan_tiles, 1, SCAN_THREADS, 0, in, out, cnt, (int32_t*)nullptr, (const int32_t*)nullptr);
            return KFSP_OK;
        }
        int32_t* sums = tile_buf;
        int32_t* offs = tile_buf + tiles + 1;
        KFSP_LAUNCH(k_scan_tiles, (int)tiles, SCAN_THREADS, 0, in, (int32_t*)nullptr, cnt, sums, (const int32_t*)nullptr);
        KFSP_TRY(scan_device(sums, offs, tiles, tile_buf + 2 * tiles + 2));
        KFSP_LAUNCH(k_scan_tiles, (int)tiles, SCAN_THREADS, 0, in, out, cnt, (int32_t*)nullptr, (const int32_t*)offs);
        return KFSP_OK;
    }
    // *total is returned on the host (synchronises)
    int exclusive_scan(const int32_t* in, int32_t* out, int64_t cnt, int32_t* tile_buf, int64_t* total) {
        KFSP_TRY(scan_device(in, out, cnt, tile_buf));
        int32_t last_ex = 0, last_in = 0;
        KFSP_CUDA(cudaMemcpyAsync(&last_ex, out + cnt - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        KFSP_CUDA(cudaMemcpyAsync(&last_in, in + cnt - 1, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        KFSP_CUDA(wait_stream());
        *total = (int64_t)last_ex + last_in;
        return KFSP_OK;
    }

    // ---------------------------------------------------------------- MATRIX_STARTER
    int fsp_init_device(int64_t count) {
        // states [0,count) are already in d_states
        ssa_calls = 0;                                  // a fresh state space restarts the per-trajectory SSA streams
        KFSP_CUDA(cudaMemsetAsync(d_table, 0xFF, sizeof(int32_t) * table_size, stream));
        KFSP_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int32_t), stream));
        n = count;
        FspView f = view();
        KFSP_LAUNCH(k_validate_states, grid_for(count * S), VEC_THREADS, 0, d_states, S, count, opt.max_molecules, d_err);
        int32_t e = 0;
        KFSP_TRY(read_err(&e));
        if (e) { n = 0; return err_to_status(e); }
        KFSP_LAUNCH(k_insert_states, grid_for(count), VEC_THREADS, 0, f, (int64_t)0, count, d_err);
        KFSP_TRY(propensities(0, count));
        KFSP_LAUNCH(k_reset_links, grid_for(count * R), VEC_THREADS, 0, f, (int64_t)0, count);
        KFSP_LAUNCH(k_resolve_links, grid_for(count * R), VEC_THREADS, 0, f);
        KFSP_TRY(read_err(&e));
        if (e) { n = 0; return err_to_status(e); }
        return KFSP_OK;
    }
    int fsp_init(int64_t count, const int32_t* states_host, bool defer_check = false) {
        if (opt.spmv_variant == 1) return box_init_from_states(count, states_host, defer_check);
        if (dist.nranks > 1 && !dist.repl) return dist_fsp_init(count, states_host);
        KFSP_TRY(ensure_state_space());
        if (count < 1 || count > opt.max_states) return KFSP_ERR_BAD_SIZES;
        KFSP_CUDA(cudaMemcpyAsync(d_states, states_host, sizeof(int32_t) * count * S, cudaMemcpyHostToDevice, stream));
        KFSP_CUDA(cudaMemsetAsync(d_w, 0, sizeof(double) * count, stream));
        KFSP_TRY(fsp_init_device(count));
        return repartition();
    }

    // Insert `ncand` candidate states (in visiting order) that sit in scratch at `cand`;
    // layout of the scratch after cand: slot[ncand], win[ncand], pos[ncand], tiles.
    int insert_candidates(int32_t* cand, int64_t ncand, int32_t* slot, int32_t* win, int32_t* pos, int32_t* tile_buf, int64_t* n_new) {
        FspView f = view();
        KFSP_LAUNCH(k_insert_candidates, grid_for(ncand), VEC_THREADS, 0, f, (const int32_t*)cand, ncand, slot, d_err);
        KFSP_LAUNCH(k_mark_winners, grid_for(ncand), VEC_THREADS, 0, (const int32_t*)d_table, (const int32_t*)slot, ncand, (int32_t)n, win);
        int64_t total = 0;
        KFSP_TRY(exclusive_scan(win, pos, ncand, tile_buf, &total));
        int32_t e = 0;
        KFSP_TRY(read_err(&e));
        if (e) {                                            // the table holds provisional slots >= n: release them
            KFSP_LAUNCH(k_rollback_candidates, grid_for(ncand), VEC_THREADS, 0, d_table, (const int32_t*)slot, ncand, (int32_t)n);
            KFSP_TRY(sync());
            return err_to_status(e);
        }
        *n_new = total;
        return KFSP_OK;
    }
    int commit_candidates(int32_t* cand, int64_t ncand, int32_t* slot, int32_t* win, int32_t* pos, int64_t n_new) {
        FspView f = view();
        KFSP_LAUNCH(k_commit_winners, grid_for(ncand), VEC_THREADS, 0, f, (const int32_t*)cand, (const int32_t*)slot,
                    (const int32_t*)win, (const int32_t*)pos, ncand, d_w);
        const int64_t first = n;
        n += n_new;
        f = view();
        KFSP_TRY(propensities(first, n_new));
        KFSP_LAUNCH(k_reset_links, grid_for(n_new * R), VEC_THREADS, 0, f, first, n_new);
        KFSP_LAUNCH(k_resolve_links, grid_for(n * R), VEC_THREADS, 0, f);
        return KFSP_OK;
    }
    static size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

    // ---------------------------------------------------------------- ONESTEP_EXTENDER
    // Several GPUs, adaptive state set: the three routines that change the state space run on the whole (replicated) state
    // space with W gathered from the ranks' slices, identically on every rank -- SSA draws are counter-based per start
    // state, reductions are double-double -- and the rows are re-partitioned afterwards.
    double repl_s[3] = {0, 0, 0};     // KFSP_DEBUG_REPL=1: host wall time of the W gather (synchronised) / re-partition, printed by destroy()
    bool repl_debug = false;
    int repl_enter() {
        if (!dist.repl) return KFSP_OK;
        const double t0 = repl_debug ? wall_now() : 0.0;
        KFSP_TRY(gather_rows(d_w));
        if (repl_debug) { KFSP_TRY(sync()); repl_s[0] += wall_now() - t0; repl_s[2] += 1; }
        dist.suspended = true;
        return KFSP_OK;
    }
    int repl_leave() {
        if (!dist.repl) return KFSP_OK;
        dist.suspended = false;
        const double t0 = repl_debug ? wall_now() : 0.0;
        const int st = repartition();
        if (repl_debug) { sync(); repl_s[1] += wall_now() - t0; }
        return st;
    }
    // in-place all-gather of a vector in the replicated layout: rank r's slice [rb[r], rb[r+1]) becomes valid everywhere
    int gather_rows(double* base) {
#ifdef KFSP_WITH_NCCL
        if (!dist.repl || dist.nranks == 1 || dist.whole) return KFSP_OK;       // `whole`: every rank computed every row
        if (ncclGroupStart() != ncclSuccess) return KFSP_ERR_NCCL;
        for (int r = 0; r < dist.nranks; ++r) {
            const int64_t cnt = dist.rb[r + 1] - dist.rb[r];
            if (cnt > 0 && ncclBroadcast(base + dist.rb[r], base + dist.rb[r], (size_t)cnt, ncclFloat64, r, dist.comm, stream) != ncclSuccess) return KFSP_ERR_NCCL;
        }
        if (ncclGroupEnd() != ncclSuccess) return KFSP_ERR_NCCL;
        dist.halo_exchanges += 1;
        dist.halo_bytes += 8 * (n - (dist.hi - dist.lo));
        return KFSP_OK;
#else
        (void)base;
        return KFSP_OK;
#endif
    }
    // new row bounds after the state set changed (block partition of [0, n) over the ranks)
    int repartition() {
        dist.n_global = n;
        if (!dist.repl) return KFSP_OK;
        // Below repl_min_rows a column of the sweep is launch latency, not bandwidth, and a cross-GPU exchange per launch only
        // adds to it (measured: Goutsias, 5e5 states on average, sweep 1.99 s on one GPU, 3.46 s split over two): every
        // rank then computes every row and nothing is exchanged.  W is complete on every rank at both transitions (the
        // caller gathered it before the state set changed).
        dist.whole = n < dist.repl_min_rows;                  // (kfsp_repl_partition is the same arithmetic for host-side tests)
        for (int r = 0; r <= dist.nranks; ++r) dist.rb[r] = part_lo(n, dist.nranks, r);
        dist.lo = dist.rb[dist.rank];
        dist.hi = dist.rb[dist.rank + 1];
        if (dist.d_peers) {
            int64_t rb[MAX_RANKS + 1] = {0};
            for (int r = 0; r <= dist.nranks; ++r) rb[r] = dist.rb[r];
            KFSP_CUDA(cudaMemcpyAsync(dist.d_peers->rb, rb, sizeof rb, cudaMemcpyHostToDevice, stream));   // pageable source: staged before the call returns
        }
        return KFSP_OK;
    }
    int fsp_onestep() {
        if (dist.nranks > 1 && !dist.repl) return KFSP_ERR_UNSUPPORTED;     // memory-scaled partitioned sets are fixed
        KFSP_TRY(repl_enter());
        const int st = fsp_onestep_1();
        const int st2 = repl_leave();
        return st != KFSP_OK ? st : st2;
    }
    int fsp_ssa(double timestep) {
        if (dist.nranks > 1 && !dist.repl) return KFSP_ERR_UNSUPPORTED;
        KFSP_TRY(repl_enter());
        const int st = fsp_ssa_1(timestep);
        const int st2 = repl_leave();
        return st != KFSP_OK ? st : st2;
    }
    int fsp_drop(double dsum, int32_t* dropped, double* droptol_out, int64_t* count_out) {
        *dropped = 0;
        if (dist.nranks > 1 && !dist.repl) return KFSP_ERR_UNSUPPORTED;
        KFSP_TRY(repl_enter());
        const int st = fsp_drop_1(dsum, dropped, droptol_out, count_out);
        const int st2 = repl_leave();
        return st != KFSP_OK ? st : st2;
    }
    int fsp_onestep_1() {
        if (box) return KFSP_ERR_UNSUPPORTED;                        // lattice state sets are fixed
        if (n < 1) return KFSP_ERR_BAD_SIZES;
        const int64_t n_old = n;
        // scratch: cnt[n_old], off[n_old], tiles
        size_t need0 = align_up(sizeof(int32_t) * n_old) * 2 + align_up(sizeof(int32_t) * scan_buf_ints(n_old));
        KFSP_TRY(ensure_scratch(need0));
        int32_t* cnt = (int32_t*)d_scratch;
        int32_t* off = (int32_t*)(d_scratch + align_up(sizeof(int32_t) * n_old));
        int32_t* tb0 = (int32_t*)(d_scratch + 2 * align_up(sizeof(int32_t) * n_old));
        FspView f = view();
        KFSP_LAUNCH(k_onestep_count, grid_for(n_old), VEC_THREADS, 0, f, n_old, cnt);
        int64_t ncand = 0;
        KFSP_TRY(exclusive_scan(cnt, off, n_old, tb0, &ncand));
        if (ncand == 0) return KFSP_OK;
        if (ncand > 2000000000LL - n_old) return KFSP_ERR_OVERFLOW;
        return expand_with(ncand, n_old, /*ssa=*/false, 0.0);
    }

    // Shared tail of ONESTEP_EXTENDER and SSA_EXTENDER: off[] (per-start-state offsets) is at
    // scratch + align(n_old ints); candidates are generated after the offsets.
    int expand_with(int64_t ncand, int64_t n_old, bool ssa, double timestep) {
        const size_t a_n = align_up(sizeof(int32_t) * n_old);
        const size_t a_c = align_up(sizeof(int32_t) * ncand);
        const size_t need = 2 * a_n + align_up(sizeof(int32_t) * ncand * S) + 3 * a_c + align_up(sizeof(int32_t) * scan_buf_ints(ncand));
        // growing the arena would lose off[]: save it first
        if (need > scratch_bytes) {
            std::vector<int32_t> keep((size_t)n_old);
            KFSP_CUDA(cudaMemcpyAsync(keep.data(), d_scratch + a_n, sizeof(int32_t) * n_old, cudaMemcpyDeviceToHost, stream));
            KFSP_CUDA(wait_stream());
            KFSP_TRY(ensure_scratch(need));
            KFSP_CUDA(cudaMemcpyAsync(d_scratch + a_n, keep.data(), sizeof(int32_t) * n_old, cudaMemcpyHostToDevice, stream));
            KFSP_CUDA(wait_stream());
        }
        int32_t* off = (int32_t*)(d_scratch + a_n);
        char* p = d_scratch + 2 * a_n;
        int32_t* cand = (int32_t*)p; p += align_up(sizeof(int32_t) * ncand * S);
        int32_t* slot = (int32_t*)p; p += a_c;
        int32_t* win = (int32_t*)p; p += a_c;
        int32_t* pos = (int32_t*)p; p += a_c;
        int32_t* tb = (int32_t*)p;
        FspView f = view();
        if (ssa) {
            if (ssa_emit_ready) {                              // the counting pass already holds every candidate: copy them into order
                KFSP_LAUNCH(k_ssa_gather, grid_for(n_old), VEC_THREADS, 0, n_old, (const int32_t*)off, ncand, emit, S, cand);
            } else {
                KFSP_LAUNCH_SSA(true, grid_for(n_old, 128), 128, 0, f, n_old, timestep, (uint64_t)opt.seed, ssa_calls,
                            (int32_t*)nullptr, (const int32_t*)off, cand, d_err, (int32_t)(1 << 24), ncand, host_prop ? pc : PropCache(),
                            (int32_t*)nullptr, SsaEmit(), fac, ssa_use_fac());
            }
        } else {
            KFSP_LAUNCH(k_onestep_fill, grid_for(n_old), VEC_THREADS, 0, f, n_old, (const int32_t*)off, cand, d_err);
        }
        int64_t n_new = 0;
        int st = insert_candidates(cand, ncand, slot, win, pos, tb, &n_new);
        if (st != KFSP_OK) return st;
        // ONESTEP: STOP once FSP%SIZE >= KTLEN (StateSpace.f90:388-391); SSA returns silently (:612-616),
        // and the ONESTEP that always follows it stops -- both are reported as overflow here.
        if (n + n_new >= opt.max_states) {
            KFSP_LAUNCH(k_rollback_candidates, grid_for(ncand), VEC_THREADS, 0, d_table, (const int32_t*)slot, ncand, (int32_t)n);
            KFSP_TRY(sync());
            return KFSP_ERR_OVERFLOW;
        }
        return commit_candidates(cand, ncand, slot, win, pos, n_new);
    }

    // candidates emitted during the counting pass of the SSA walks (state_space.cuh: SsaEmit); KFSP_SSA_EMIT=0: replay the walks (A/B)
    SsaEmit emit;
    size_t emit_states = 0;
    bool ssa_emit_on = true, emit_armed = false, ssa_emit_ready = false;
    int prepare_ssa_emit(int64_t n_old) {
        emit_armed = false;
        if (!ssa_emit_on) return KFSP_OK;
        const size_t want = (size_t)std::max<int64_t>(1 << 20, n_old);          // candidates per expansion: a boundary fraction of the set
        if (want > emit_states || !emit.tmp) {
            KFSP_CUDA(wait_stream());
            cudaFree(emit.tmp); cudaFree(emit.prev); cudaFree(emit.head); cudaFree(emit.cursor);
            emit = SsaEmit();
            emit_states = std::max(want, 2 * emit_states);
            KFSP_CUDA(cudaMalloc(&emit.tmp, sizeof(int32_t) * emit_states * S));
            KFSP_CUDA(cudaMalloc(&emit.prev, sizeof(int32_t) * emit_states));
            KFSP_CUDA(cudaMalloc(&emit.head, sizeof(int32_t) * emit_states));
            KFSP_CUDA(cudaMalloc(&emit.cursor, sizeof(int32_t)));
            emit.cap = (int32_t)std::min<size_t>(emit_states, 2000000000u);
        }
        KFSP_CUDA(cudaMemsetAsync(emit.cursor, 0, sizeof(int32_t), stream));
        return KFSP_OK;
    }

    int ssa_use_fac() const { return (fac_ok && ssa_fac_on && !host_prop) ? 1 : 0; }
    // ---------------------------------------------------------------- SSA_EXTENDER
    int fsp_ssa_1(double timestep) {
        if (box) return KFSP_ERR_UNSUPPORTED;
        if (n < 1) return KFSP_ERR_BAD_SIZES;
        const int64_t n_old = n;
        ssa_calls += 1;
        emit_armed = false;
        size_t need0 = align_up(sizeof(int32_t) * n_old) * 2 + align_up(sizeof(int32_t) * scan_buf_ints(n_old));
        const size_t wsave_at = need0;                  // suspended walks of the host-propensity rounds (dead before the scan)
        if (host_prop) need0 += align_up(sizeof(int32_t) * n_old * (size_t)(S + 4));
        KFSP_TRY(ensure_scratch(need0));
        int32_t* cnt = (int32_t*)d_scratch;
        int32_t* off = (int32_t*)(d_scratch + align_up(sizeof(int32_t) * n_old));
        int32_t* tb0 = (int32_t*)(d_scratch + 2 * align_up(sizeof(int32_t) * n_old));
        int32_t* wsave = (int32_t*)(d_scratch + wsave_at);
        FspView f = view();
        if (host_prop) {
            // Rounds: walks that step onto a state the host has not evaluated yet are suspended (cnt = -1) and
            // replayed once the host has put a_k of the requested states into the side cache.
            // The side cache PERSISTS across expansions (and across the drops in between): a_k of a state never changes while
            // the model stands, walks of successive expansions leave the projection through the same boundary region, and
            // dropped states come back -- config 4 asked the host 3.0e8 times for 970k final states when the cache was
            // emptied per call.  It is emptied by set_model, and here if it outgrows its budget.
            KFSP_TRY(ensure_prop_cache(1));
            if (pc_n > pc_budget) KFSP_TRY(clear_prop_cache());
            KFSP_LAUNCH(k_fill_i32, grid_for(n_old), VEC_THREADS, 0, cnt, n_old, (int32_t)-1);
            const size_t a_r = align_up(sizeof(int32_t) * (size_t)pc.req_cap * S), a_v = align_up(sizeof(double) * (size_t)pc.req_cap * (R + 1));
            KFSP_TRY(ensure_hp(256, 256 + 2 * a_r + a_v));
            int32_t* h_nreq = (int32_t*)hp_host;
            int32_t* h_req = (int32_t*)(hp_host + 256);
            int32_t* h_uniq = (int32_t*)(hp_host + 256 + a_r);
            double* h_vals = (double*)(hp_host + 256 + 2 * a_r);
            struct Key { int32_t v[KFSP_MAX_SPECIES]; bool operator==(const Key& o) const { return std::memcmp(v, o.v, sizeof v) == 0; } };
            struct KeyHash { size_t operator()(const Key& k) const { return (size_t)hash_state(k.v, KFSP_MAX_SPECIES); } };
            std::unordered_set<Key, KeyHash> seen;
            for (int64_t round = 0;; ++round) {
                if (round > (1 << 24)) return KFSP_ERR_SSA_RUNAWAY;
                KFSP_CUDA(cudaMemsetAsync(pc.nreq, 0, sizeof(int32_t), stream));
                KFSP_LAUNCH_SSA(false, grid_for(n_old, 128), 128, 0, f, n_old, timestep, (uint64_t)opt.seed, ssa_calls, cnt,
                            (const int32_t*)nullptr, (int32_t*)nullptr, d_err, (int32_t)(1 << 24), (int64_t)0, pc, wsave, SsaEmit(), fac, ssa_use_fac());
                KFSP_CUDA(cudaMemcpyAsync(h_nreq, pc.nreq, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
                // the request list is small: fetch it with the counter instead of paying a second round trip
                KFSP_CUDA(cudaMemcpyAsync(h_req, pc.req, sizeof(int32_t) * (size_t)std::min<int64_t>(pc.req_cap, 4096) * S, cudaMemcpyDeviceToHost, stream));
                KFSP_CUDA(wait_stream());
                const int32_t nreq = *h_nreq;
                if (nreq == 0) break;
                ++host_prop_rounds;
                const int64_t got = std::min<int64_t>(nreq, pc.req_cap);
                if (got > 4096) {
                    KFSP_CUDA(cudaMemcpyAsync(h_req + 4096 * S, pc.req + 4096 * S, sizeof(int32_t) * (got - 4096) * S, cudaMemcpyDeviceToHost, stream));
                    KFSP_CUDA(wait_stream());
                }
                // several walks may ask for the same state in one round
                seen.clear();
                int64_t nu = 0;
                for (int64_t t = 0; t < got; ++t) {
                    Key key;
                    std::memset(key.v, 0, sizeof key.v);
                    std::memcpy(key.v, h_req + t * S, sizeof(int32_t) * S);
                    if (seen.insert(key).second) { std::memcpy(h_uniq + nu * S, h_req + t * S, sizeof(int32_t) * S); ++nu; }
                }
                eval_host(h_uniq, nu, h_vals);
                KFSP_TRY(ensure_prop_cache(pc_n + nu));
                KFSP_CUDA(cudaMemcpyAsync(pc.states + pc_n * S, h_uniq, sizeof(int32_t) * nu * S, cudaMemcpyHostToDevice, stream));
                KFSP_CUDA(cudaMemcpyAsync(pc.prop + pc_n * (R + 1), h_vals, sizeof(double) * nu * (R + 1), cudaMemcpyHostToDevice, stream));
                KFSP_LAUNCH(k_cache_insert, grid_for(nu), VEC_THREADS, 0, pc, S, pc_n, nu, d_err);
                KFSP_CUDA(wait_stream());        // h_uniq / h_vals are rewritten by the next round
                pc_n += nu;
            }
        } else {
            KFSP_TRY(prepare_ssa_emit(n_old));
            KFSP_LAUNCH_SSA(false, grid_for(n_old, 128), 128, 0, f, n_old, timestep, (uint64_t)opt.seed, ssa_calls, cnt,
                        (const int32_t*)nullptr, (int32_t*)nullptr, d_err, (int32_t)(1 << 24), (int64_t)0, PropCache(), (int32_t*)nullptr,
                        ssa_emit_on ? emit : SsaEmit(), fac, ssa_use_fac());
            emit_armed = ssa_emit_on;
        }
        int64_t ncand = 0;
        KFSP_TRY(exclusive_scan(cnt, off, n_old, tb0, &ncand));
        int32_t e = 0;
        KFSP_TRY(read_err(&e));
        if (e) return err_to_status(e);
        if (ncand == 0) return KFSP_OK;
        if (ncand > 2000000000LL - n_old) return KFSP_ERR_OVERFLOW;
        ssa_emit_ready = emit_armed && ncand <= emit.cap;       // else: replay the walks to fill the list
        const int st = expand_with(ncand, n_old, /*ssa=*/true, timestep);
        ssa_emit_ready = false;
        return st;
    }

    // ---------------------------------------------------------------- FMATVEC
    // MODE 0: y = A x.  MODE 1: one Arnoldi column (x = U_c, g = U_{c-1} or null): ea says where H(J-1,J), H(J,J) go.
    // MODE 2: the extra product (AVNORM).
    static EpiArgs epi_none() {
        EpiArgs e;
        e.kind = RK_COLUMN; e.column = -1; e.fin = 0; e.has_g = 0; e.break_tol = 0.0; e.h1_out = e.h2_out = e.hn_out = nullptr;
        return e;
    }
    template <int MODE>
    int spmv(const double* x, double* y, const double* g = nullptr, EpiArgs ea = epi_none(), int cg = -1) {
        const bool multi = dist_active();
        const int64_t rows = kn(), r0 = kr0();
        // several GPUs, peer-memory halo: remote rows are addressed as (peer's basis) + (x - d_V), which only means
        // something for a column of the basis.  Any other operand (kfsp_matvec, kfsp_matvec_device) is staged in the
        // scratch column first; the barrier makes every rank's copy visible before any rank gathers from it.
        if (multi && dist.p2p && dist.p2p_halo) {
            KFSP_TRY(ensure_basis());
            const double* vend = d_V + (size_t)ld * (opt.m_max + 4);
            if (x < d_V || x >= vend) {
                double* stage = d_V + (size_t)ld * (opt.m_max + 2) + r0;
                KFSP_CUDA(cudaMemcpyAsync(stage, x, sizeof(double) * rows, cudaMemcpyDeviceToDevice, stream));
                KFSP_TRY(dist_barrier());
                x = stage;
            }
        }
        ea.has_g = (MODE == 1 && g) ? 1 : 0;
        KFSP_TRY(prof_begin(MODE == 0 ? KFSP_PROF_SPMV_PLAIN : MODE == 1 ? KFSP_PROF_SPMV_DOT : KFSP_PROF_SPMV_NRM,
                            (box ? 16 : idx ? 4 * R + 4 * S + 24 : 12 * R + 24) + 8 * ea.has_g));
        spmv_by_mode[MODE] += 1;
        if (box) {
            KFSP_TRY(spmv_box<MODE>(x, y, g, ea, cg));
            return prof_end();
        }
        // 1 / 2: memory-scaled partition (local rows + halo plan; NCCL exchange / peer loads); 3 / 4: replicated layout
        // (global indices; peer loads / column completed by an all-gather first)
        int halo = 0;
        if (multi) halo = dist.repl ? ((dist.p2p && dist.p2p_halo) ? 3 : 4) : ((dist.p2p && dist.p2p_halo) ? 2 : 1);
        if (halo == 1) KFSP_TRY(dist_halo_exchange(x));
        if (halo == 4) KFSP_TRY(gather_rows(const_cast<double*>(x) - r0));
        const Reducer r = MODE != 0 ? next_rd() : rd;
        Reducer r2 = r;
        if ((halo == 2 || halo == 3) && !r2.peers) r2.peers = dist.d_peers;     // the peer table is also the halo address book
        if (!multi) { r2.peers = nullptr; r2.dist_send = nullptr; r2.seq = 0; }
        const int64_t coloff = d_V ? x - d_V : 0;
        if (idx) {
            void (*ki)(const FacModel, int64_t, int64_t, const int32_t*, const int32_t*, const double*, const double*, double*, const double*,
                       Reducer, SweepCtl*, EpiArgs, int, const double*, int64_t, int64_t, int64_t) = nullptr;
#define KFSP_IDX_PICK(RR, SS)                                                                                           \
            ki = halo == 4 ? k_spmv_idx<RR, SS, MODE, 4, 0> : halo == 3 ? k_spmv_idx<RR, SS, MODE, 3, 0>                   \
               : halo == 2 ? k_spmv_idx<RR, SS, MODE, 2, 0> : halo == 1 ? k_spmv_idx<RR, SS, MODE, 1, 0> : k_spmv_idx<RR, SS, MODE, 0, 0>
            bool gen = false;                                   // a reaction that needs the postfix program: generic kernel only
            for (int k = 0; k < R; ++k) gen = gen || fac.shape[k] == FAC_GEN;
            if (!gen && R == 4 && S == 2) { KFSP_IDX_PICK(4, 2); }
            else if (!gen && R == 6 && S == 3) { KFSP_IDX_PICK(6, 3); }
            else if (!gen && R == 10 && S == 6) { KFSP_IDX_PICK(10, 6); }
            else { KFSP_IDX_PICK(0, 0); }
#undef KFSP_IDX_PICK
            // the rows' states: this rank's rows start at lo in the state list on a partitioned handle
            const int32_t* st = d_states + (multi ? dist.lo * S : 0);
            KFSP_TRY(launch_pdl(ki, wave_grid((const void*)ki, rows), VEC_THREADS, 0, fac, rows, ld, (const int32_t*)d_pred + r0, st,
                                (const double*)d_diag + r0, x, y, g, r2, d_ctl, ea, cg, (const double*)dist.halo, rows, coloff, r0));
        } else {
            void (*kern)(int64_t, int64_t, int, const int32_t*, const double*, const double*, const double*, double*, const double*,
                         Reducer, SweepCtl*, EpiArgs, int, const double*, int64_t, int64_t, int64_t);
            // tuning variant (KFSP_SPMV_TUNE): 0 = 1 row/iter, 1 = 2 rows/iter, 3/4/5 = 1 row/iter capped at 8/6/5 CTAs per SM
#define KFSP_SPMV_PICK(RR)                                                                                              \
            kern = halo == 4 ? k_spmv<RR, MODE, 1, 4, 4> : halo == 3 ? k_spmv<RR, MODE, 1, 4, 3>                             \
                 : halo == 2 ? k_spmv<RR, MODE, 1, 4, 2> : halo == 1 ? k_spmv<RR, MODE, 1, 4, 1>                             \
                 : spmv_tune == 1 ? k_spmv<RR, MODE, 2, 1, 0> : spmv_tune == 3 ? k_spmv<RR, MODE, 1, 8, 0>                   \
                 : spmv_tune == 4 ? k_spmv<RR, MODE, 1, 6, 0> : spmv_tune == 5 ? k_spmv<RR, MODE, 1, 5, 0>                   \
                 : k_spmv<RR, MODE, 1, 1, 0>
            switch (R) {
            case 4: KFSP_SPMV_PICK(4); break;
            case 6: KFSP_SPMV_PICK(6); break;
            case 10: KFSP_SPMV_PICK(10); break;
            default: kern = halo == 4 ? k_spmv<0, MODE, 1, 4, 4> : halo == 3 ? k_spmv<0, MODE, 1, 4, 3> : halo == 2 ? k_spmv<0, MODE, 1, 4, 2>
                          : halo == 1 ? k_spmv<0, MODE, 1, 4, 1> : k_spmv<0, MODE, 1, 1, 0>; break;
            }
#undef KFSP_SPMV_PICK
            KFSP_TRY(launch_pdl(kern, wave_grid((const void*)kern, rows), VEC_THREADS, 0, rows, ld, R, (const int32_t*)d_pred + r0, (const double*)d_coef + r0,
                                (const double*)d_diag + r0, x, y, g, r2, d_ctl, ea, cg, (const double*)dist.halo, rows, coloff, r0));
        }
        if (multi && MODE != 0) KFSP_TRY(dist_finalize(ea, MODE == 1 ? 4 : 2));      // NCCL reduction path only (no-op with peer memory)
        return prof_end();
    }
    int set_profiling(int level) {
        const bool on = level != 0;
        profile_level = level;
        profile_spmv = on;
        if (on && ev_pool.empty()) {
            ev_pool.resize(2 * 8192);
            ev_cls.assign(8192, 0);
            ev_bps.assign(8192, 0);
            ev_n.assign(8192, 1);
            ev_nspmv.assign(8192, 0);
            for (auto& e : ev_pool) KFSP_CUDA(cudaEventCreate(&e));
        }
        ev_used = 0;
        return KFSP_OK;
    }
    // sum the bracketed SpMV times recorded since the last call (synchronises)
    int collect_profile() {
        if (!profile_spmv || ev_used == 0) return KFSP_OK;
        KFSP_CUDA(wait_stream());
        for (size_t i = 0; i + 1 < ev_used; i += 2) {
            float ms = 0.f;
            KFSP_CUDA(cudaEventElapsedTime(&ms, ev_pool[i], ev_pool[i + 1]));
            const int cls = ev_cls[i / 2];
            prof_sec[cls] += 1e-3 * ms;
            prof_cnt[cls] += ev_n[i / 2];
            prof_bps[cls] += ev_bps[i / 2];
            if (cls <= KFSP_PROF_SPMV_FIN_NRM) {                 // the generator SpMV in all its variants
                spmv_seconds += 1e-3 * ms;
                spmv_timed += 1;
            } else if (cls == KFSP_PROF_SWEEP && ev_nspmv[i / 2] == ev_n[i / 2]) {   // a sweep of SpMV-class launches only
                spmv_seconds += 1e-3 * ms;
                spmv_timed += ev_n[i / 2];
            }
        }
        ev_used = 0;
        return KFSP_OK;
    }

    // ---------------------------------------------------------------- DROP_STATES
    int fsp_drop_1(double dsum, int32_t* dropped, double* droptol_out, int64_t* count_out) {
        *dropped = 0;
        if (box) return KFSP_ERR_UNSUPPORTED;
        if (n < 1) return KFSP_ERR_BAD_SIZES;
        const int64_t lsize = n;
        // FIND_DROPTOL (StateSpace.f90:398-427): thresholds by repeated division
        double droptol = opt.drop_tol0;
        const size_t a_i = align_up(sizeof(int32_t) * lsize);
        const size_t need = 256 + 3 * a_i + align_up(sizeof(int32_t) * scan_buf_ints(lsize)) +
                            align_up((size_t)lsize * (size_t)std::max(std::max(8 * R, 4 * S), 8)) + 64;     // compaction scratch: R doubles or S ints per state
        KFSP_TRY(ensure_scratch(need));
        char* p = d_scratch;
        unsigned long long* d_cnt = (unsigned long long*)p; p += 256;
        int32_t* drop = (int32_t*)p; p += a_i;
        int32_t* keep = (int32_t*)p; p += a_i;
        int32_t* pos = (int32_t*)p; p += a_i;
        int32_t* tb = (int32_t*)p; p += align_up(sizeof(int32_t) * scan_buf_ints(lsize));
        char* tmp = p;
        // one double-double reduction per candidate threshold, exactly the reference's loop
        for (int it = 0; it < 400; ++it) {
            KFSP_LAUNCH(k_sum_below, grid_for(lsize), VEC_THREADS, 0, lsize, (const double*)d_w, droptol, next_rd(), d_ctl);
            KFSP_TRY(read_ctl());
            if (h_ctl->scal[SC_WSUM] < dsum) break;
            droptol = droptol / 10.0;
            if (droptol == 0.0) break;
        }
        // mark, derivative test, count
        KFSP_CUDA(cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), stream));
        KFSP_LAUNCH(k_drop_mark, grid_for(lsize), VEC_THREADS, 0, (const double*)d_w, lsize, droptol, drop, d_cnt);
        double* aw = (double*)tmp;         // WTMP; the compaction scratch is not in use yet
        KFSP_TRY(spmv<0>(d_w, aw));
        KFSP_LAUNCH(k_drop_unmark, grid_for(lsize), VEC_THREADS, 0, (const double*)aw, lsize, opt.drop_deriv_tol, drop, d_cnt + 1);
        unsigned long long c[2];
        KFSP_CUDA(cudaMemcpyAsync(c, d_cnt, sizeof c, cudaMemcpyDeviceToHost, stream));
        KFSP_TRY(sync());
        const int64_t drop_count = (int64_t)c[0] - (int64_t)c[1];
        if (droptol_out) *droptol_out = droptol;
        if (count_out) *count_out = drop_count;
        if (!((double)drop_count * 1.0 / ((double)lsize * 1.0) > opt.drop_fraction)) return KFSP_OK;
        // stable compaction
        KFSP_LAUNCH(k_invert_flags, grid_for(lsize), VEC_THREADS, 0, (const int32_t*)drop, keep, lsize);
        int64_t q = 0;
        KFSP_TRY(exclusive_scan(keep, pos, lsize, tb, &q));
        if (q < 1) {
            // every state dropped: the reference would continue with an empty projection and divide by zero
            return KFSP_ERR_BAD_SIZES;
        }
        const int g = grid_for(lsize * std::max(R, S));
        // states
        KFSP_LAUNCH(k_compact_states, g, VEC_THREADS, 0, (const int32_t*)d_states, (int32_t*)tmp, (const int32_t*)keep, (const int32_t*)pos, lsize, S);
        KFSP_CUDA(cudaMemcpyAsync(d_states, tmp, sizeof(int32_t) * q * S, cudaMemcpyDeviceToDevice, stream));
        // w (zero the tail: W(1:LSIZE) = 0 then copy, StateSpace.f90:528-534)
        KFSP_LAUNCH(k_compact_rows<double>, g, VEC_THREADS, 0, (const double*)d_w, (double*)tmp, (const int32_t*)keep, (const int32_t*)pos, lsize, lsize, q, 1);
        KFSP_CUDA(cudaMemsetAsync(d_w, 0, sizeof(double) * lsize, stream));
        KFSP_CUDA(cudaMemcpyAsync(d_w, tmp, sizeof(double) * q, cudaMemcpyDeviceToDevice, stream));
        // diag
        KFSP_LAUNCH(k_compact_rows<double>, g, VEC_THREADS, 0, (const double*)d_diag, (double*)tmp, (const int32_t*)keep, (const int32_t*)pos, lsize, lsize, q, 1);
        KFSP_CUDA(cudaMemcpyAsync(d_diag, tmp, sizeof(double) * q, cudaMemcpyDeviceToDevice, stream));
        // prop (R rows)
        KFSP_LAUNCH(k_compact_rows<double>, g, VEC_THREADS, 0, (const double*)d_prop, (double*)tmp, (const int32_t*)keep, (const int32_t*)pos, lsize, ld, q, R);
        KFSP_CUDA(cudaMemcpy2DAsync(d_prop, sizeof(double) * ld, tmp, sizeof(double) * q, sizeof(double) * q, R, cudaMemcpyDeviceToDevice, stream));
        // succ (R rows, re-indexed)
        KFSP_LAUNCH(k_compact_succ, g, VEC_THREADS, 0, (const int32_t*)d_succ, (int32_t*)tmp, (const int32_t*)keep, (const int32_t*)pos, lsize, ld, q, R);
        KFSP_CUDA(cudaMemcpy2DAsync(d_succ, sizeof(int32_t) * ld, tmp, sizeof(int32_t) * q, sizeof(int32_t) * q, R, cudaMemcpyDeviceToDevice, stream));
        // hash table and row form are rebuilt for the survivors
        n = q;
        FspView f = view();
        KFSP_CUDA(cudaMemsetAsync(d_table, 0xFF, sizeof(int32_t) * table_size, stream));
        KFSP_LAUNCH(k_insert_states, grid_for(q), VEC_THREADS, 0, f, (int64_t)0, q, d_err);
        KFSP_LAUNCH(k_fill_i32, grid_for(ld * R), VEC_THREADS, 0, d_pred, ld * R, IDX_ABSENT);
        KFSP_LAUNCH(k_resolve_links, grid_for(q * R), VEC_THREADS, 0, f);
        int32_t e = 0;
        KFSP_TRY(read_err(&e));
        if (e) return err_to_status(e);
        *dropped = 1;
        return KFSP_OK;
    }

    // ---------------------------------------------------------------- Arnoldi / IOP-2 sweep
    // columns J = jold..m (1-based) then the extra product (KrylovSolver.f90:236-266). No host sync.
    int arnoldi(int jold, int m) {
        // small state spaces: the whole sweep in one single-CTA launch (bit-identical, see k_sweep_small)
        // ... and entirely in shared memory when the gather form and three vectors fit (a couple of thousand states)
        if (!dist_active() && !box && !idx && !profile_spmv && small_sweep && smem_sweep && sweep_smem_bytes(n, R) <= SWEEP_SMEM_MAX) {
            void (*ks)(int, int64_t, int, const int32_t*, const double*, const double*, double*, double*, int, int, int, SweepCtl*, double);
            switch (R) {
            case 4: ks = k_sweep_smem<4>; break;
            case 6: ks = k_sweep_smem<6>; break;
            case 10: ks = k_sweep_smem<10>; break;
            default: ks = k_sweep_smem<0>; break;
            }
            bool known = false;
            for (const void* q : smem_sweep_ready) known = known || q == (const void*)ks;
            if (!known) {
                KFSP_CUDA(cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SWEEP_SMEM_MAX));
                smem_sweep_ready.push_back((const void*)ks);
            }
            ks<<<1, SWEEP_THREADS, sweep_smem_bytes(n, R), stream>>>((int)n, ld, R, d_pred, d_coef, d_diag, d_V, d_H, LDH, jold, m, d_ctl, opt.break_tol);
            return check_launch();
        }
        if (!dist_active() && !box && !profile_spmv && small_sweep && n * (int64_t)(12 * R + 88) <= (1 << 20)) {
            void (*kern)(int64_t, int64_t, int, const int32_t*, const double*, const double*, double*, double*, int, int, int, SweepCtl*, double,
                         const FacModel, const int32_t*);
            switch (R) {
            case 4: kern = idx ? k_sweep_small<4, 1> : k_sweep_small<4, 0>; break;
            case 6: kern = idx ? k_sweep_small<6, 1> : k_sweep_small<6, 0>; break;
            case 10: kern = idx ? k_sweep_small<10, 1> : k_sweep_small<10, 0>; break;
            default: kern = idx ? k_sweep_small<0, 1> : k_sweep_small<0, 0>; break;
            }
            kern<<<1, SWEEP_THREADS, 0, stream>>>(n, ld, R, d_pred, d_coef, d_diag, d_V, d_H, LDH, jold, m, d_ctl, opt.break_tol, fac,
                                                  (const int32_t*)d_states);
            return check_launch();
        }
        // mid-sized sets: the whole sweep as ONE cooperative launch over all SMs (k_sweep_coop); KFSP_COOP_SWEEP=0 switches it off,
        // KFSP_COOP_MAX_ROWS moves the upper limit (above it a column is bandwidth, not latency, and the multi-launch kernels win)
        // (not for the index-only variant: at 1024 threads per CTA its row evaluation spills, measured 1.96 vs 1.75 s on Goutsias)
        if (!dist_active() && !box && !idx && !profile_spmv && coop_sweep && n <= coop_max_rows) {
            const int st = arnoldi_coop(jold, m);
            if (st != KFSP_ERR_UNSUPPORTED) return st;
        }
        KFSP_TRY(sweep_begin());
        const int st = (box && box_tune < 10 && lattice_bd2_order(lat) >= 0) ? arnoldi_fused(jold, m) : arnoldi_unfused(jold, m);
        KFSP_TRY(sweep_end());
        return st;
    }
    bool coop_sweep = true, coop_used = false;
    int64_t coop_max_rows = 1 << 20;      // measured: 12 % faster at 9e4 rows (repressilator), equal at 8.6e5 (Goutsias: a column is HBM traffic there)
    CoopBuf coop = {nullptr, nullptr, nullptr};
    int coop_ok = -1;                                        // device supports cooperative launches
    int arnoldi_coop(int jold, int m) {
        if (coop_ok < 0) {
            int v = 0;
            if (cudaDeviceGetAttribute(&v, cudaDevAttrCooperativeLaunch, device) != cudaSuccess) { cudaGetLastError(); v = 0; }
            coop_ok = v;
        }
        if (!coop_ok) return KFSP_ERR_UNSUPPORTED;
        void (*kern)(int64_t, int64_t, int, const int32_t*, const double*, const double*, double*, double*, int, int, int, SweepCtl*, double,
                     const FacModel, const int32_t*, CoopBuf);
        switch (R) {
        case 4: kern = idx ? k_sweep_coop<4, 1> : k_sweep_coop<4, 0>; break;
        case 6: kern = idx ? k_sweep_coop<6, 1> : k_sweep_coop<6, 0>; break;
        case 10: kern = idx ? k_sweep_coop<10, 1> : k_sweep_coop<10, 0>; break;
        default: kern = idx ? k_sweep_coop<0, 1> : k_sweep_coop<0, 0>; break;
        }
        if (!coop.bar) {
            KFSP_CUDA(cudaMalloc(&coop.bar, 256));
            KFSP_CUDA(cudaMalloc(&coop.part, sizeof(double) * 2 * 6 * COOP_MAXG));
            coop.err = d_err;
        }
        int g = wave_grid((const void*)kern, n, SWEEP_THREADS);      // co-resident by construction: SMs x occupancy of this kernel
        if (g > COOP_MAXG) g = COOP_MAXG;
        KFSP_CUDA(cudaMemsetAsync(coop.bar, 0, sizeof(unsigned int), stream));
        int64_t n_ = n, ld_ = ld;
        int R_ = R, ldh_ = LDH, jold_ = jold, m_ = m;
        const int32_t* pred_ = d_pred; const double* coef_ = d_coef; const double* diag_ = d_diag;
        double* V_ = d_V; double* H_ = d_H; SweepCtl* ctl_ = d_ctl; double bt_ = opt.break_tol;
        FacModel fac_ = fac; const int32_t* st_ = d_states; CoopBuf cb_ = coop;
        void* args[] = {&n_, &ld_, &R_, &pred_, &coef_, &diag_, &V_, &H_, &ldh_, &jold_, &m_, &ctl_, &bt_, &fac_, &st_, &cb_};
        KFSP_CUDA(cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)g), dim3(SWEEP_THREADS), args, 0, stream));
        coop_used = true;
        return check_launch();
    }
    int arnoldi_unfused(int jold, int m) {
        // Two launches per column: finalise U_c (the two axpys of the previous column + its norm), then the generator product
        // with the three inner products of the window.  Column 0 and, on a resumed sweep (KrylovSolver.f90:400-433), column
        // jold-1 are complete already.  (Replicated layout on several GPUs: this rank's slice of every column.)
        const int64_t rows = kn(), r0 = kr0();
        for (int J = jold; J <= m + 1; ++J) {                    // J = m+1: the extra product for AVNORM (:261-263)
            const int c = J - 1;
            double* vc = d_V + (size_t)c * ld + r0;              // holds A U_{c-1} until finalised into U_c
            double* vn = d_V + (size_t)J * ld + r0;
            const double* vg = c >= 1 ? d_V + (size_t)(c - 1) * ld + r0 : nullptr;
            if (J > jold) {
                const double* vf = c >= 2 ? d_V + (size_t)(c - 2) * ld + r0 : vg;
                EpiArgs ef = epi_none();
                ef.kind = RK_FIN_NRM; ef.column = c; ef.fin = 1; ef.break_tol = opt.break_tol; ef.hn_out = d_H + (size_t)(c - 1) * LDH + c;   // H(c+1,c)
                KFSP_TRY(prof_begin(KFSP_PROF_AXPY_NRM, 24 + (c >= 2 ? 8 : 0)));
                KFSP_TRY(launch_pdl(k_finalize, wave_grid((const void*)k_finalize, rows), VEC_THREADS, 0, rows, vg, vf, vc, c >= 2 ? 1 : 0, next_rd(), d_ctl, ef, c - 1, c - 2));
                KFSP_TRY(dist_finalize(ef, 1));
                KFSP_TRY(prof_end());
            }
            EpiArgs ea = epi_none();
            ea.column = c;
            if (J <= m) {
                ea.kind = RK_COLUMN;
                ea.h1_out = c >= 1 ? d_H + (size_t)c * LDH + (c - 1) : nullptr;          // H(J-1,J)
                ea.h2_out = d_H + (size_t)c * LDH + c;                                    // H(J,J)
                KFSP_TRY(spmv<1>(vc, vn, vg, ea, c - 1));
            } else {
                ea.kind = RK_EXTRA;
                KFSP_TRY(spmv<2>(vc, vn, nullptr, ea, -1));
            }
        }
        return KFSP_OK;
    }
    // The same sweep on the stencil lattice kernel with the tail of every column fused into the load stage of the next
    // generator product (lattice.cuh, FIN = 1): ONE launch and one reduction point per Arnoldi column, 40 bytes per state.
    // A U_c goes to one of two scratch columns (alternating), from where the next launch finalises it into column c+1.
    int arnoldi_fused(int jold, int m) {
        double* T[2] = {d_V + (size_t)(opt.m_max + 2) * ld, d_V + (size_t)(opt.m_max + 3) * ld};
        for (int J = jold; J <= m + 1; ++J) {                    // J = m+1: the extra product for AVNORM (:261-263)
            const int c = J - 1;
            double* vc = d_V + (size_t)c * ld;
            const bool fin = J > jold;
            const bool extra = J == m + 1;
            Bd2Args a;
            std::memset(&a, 0, sizeof a);
            a.src = fin ? T[c & 1] : vc;                         // A U_{c-1} was written to T[(J-1) & 1] by the previous launch
            a.xout = vc;
            a.y = T[J & 1];
            a.has_g = c >= 1 ? 1 : 0;
            a.has_f = (fin && c >= 2) ? 1 : 0;
            a.g = c >= 1 ? d_V + (size_t)(c - 1) * ld : vc;
            a.f = c >= 2 ? d_V + (size_t)(c - 2) * ld : a.g;
            a.cg = c - 1; a.cf = c - 2;
            a.ea = epi_none();
            a.ea.column = c; a.ea.fin = fin ? 1 : 0; a.ea.has_g = a.has_g; a.ea.break_tol = opt.break_tol;
            a.ea.hn_out = fin ? d_H + (size_t)(c - 1) * LDH + c : nullptr;               // H(c+1,c) = ||U_c||
            if (!extra) {
                a.ea.kind = RK_COLUMN;
                a.ea.h1_out = c >= 1 ? d_H + (size_t)c * LDH + (c - 1) : nullptr;        // H(J-1,J)
                a.ea.h2_out = d_H + (size_t)c * LDH + c;                                  // H(J,J)
            } else {
                a.ea.kind = RK_EXTRA;
            }
            KFSP_TRY(prof_begin(extra ? KFSP_PROF_SPMV_FIN_NRM : fin ? KFSP_PROF_SPMV_FIN_DOT : KFSP_PROF_SPMV_DOT,
                                8 * (2 + a.has_g + a.has_f + (fin ? 1 : 0))));
            spmv_by_mode[extra ? 2 : 1] += 1;
            if (fin) spmv_fused += 1;
            if (extra) KFSP_TRY(spmv_bd2<2>(a, fin));
            else KFSP_TRY(spmv_bd2<1>(a, fin));
            KFSP_TRY(prof_end());
        }
        return KFSP_OK;
    }
    // orders >= expm_cluster_from run the squarings on a cluster of 8 CTAs (expm.cuh); KFSP_EXPM_CLUSTER=1000 keeps everything on one CTA (A/B)
    int expm_cluster_from = 36;
    bool expm_small_threads = true;
    int launch_expm(int mx_ok, double t_ok, int use_brk, double t_brk, int set_one, const SweepCtl* ctl, double* full_out) {
        if (mx_ok < expm_cluster_from) {
            // orders up to 32 are 8x8 tiles of 4x4: 8 warps own them all, and every block-wide barrier of the Pade / LU / squaring
            // chain is cheaper with 8 warps than with 32 (same element operations: bit-identical; KFSP_EXPM_SMALL_THREADS=0 for A/B)
            const int threads = (mx_ok <= 32 && expm_small_threads) ? 256 : EXPM_THREADS;
            KFSP_LAUNCH(k_expm, 1, threads, EXPM_SMEM, d_H, LDH, mx_ok, t_ok, use_brk, t_brk, set_one, ctl, d_expm_work, d_res, full_out);
            return KFSP_OK;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(EXPM_CLUSTER, 1, 1);
        cfg.blockDim = dim3(EXPM_THREADS, 1, 1);
        cfg.dynamicSmemBytes = EXPM_SMEM;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = EXPM_CLUSTER; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        KFSP_CUDA(cudaLaunchKernelEx(&cfg, k_expm_cluster, d_H, LDH, mx_ok, t_ok, use_brk, t_brk, set_one, ctl, d_expm_work, d_res, full_out));
        return check_launch();
    }
    // exp(t*H) on the device; result struct copied to pinned memory (synchronises)
    int expm_step(int mx_ok, double t_ok, int use_brk, double t_brk, int set_one) {
        KFSP_TRY(prof_begin(KFSP_PROF_EXPM));
        KFSP_TRY(launch_expm(mx_ok, t_ok, use_brk, t_brk, set_one, (const SweepCtl*)d_ctl, nullptr));
        KFSP_TRY(prof_end());
        KFSP_CUDA(cudaMemcpyAsync(h_res, d_res, sizeof(ExpmResult), cudaMemcpyDeviceToHost, stream));
        if (coop_used && dist.nranks == 1) {                   // a cooperative sweep whose CTAs did not all arrive gave up (krylov.cuh)
            coop_used = false;
            int32_t e = 0;
            KFSP_CUDA(cudaMemcpyAsync(&e, d_err, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
            KFSP_TRY(sync());
            if (e & DEV_COOP_TIMEOUT) return KFSP_ERR_CUDA;
            return h_res->info;
        }
        coop_used = false;
        KFSP_TRY(dist.nranks > 1 ? sync_check_peers() : sync());
        return h_res->info;
    }
    int read_ctl() {
        KFSP_CUDA(cudaMemcpyAsync(h_ctl, d_ctl, sizeof(SweepCtl), cudaMemcpyDeviceToHost, stream));
        return dist.nranks > 1 ? sync_check_peers() : sync();
    }
    // multi-GPU: a peer that never delivered its partial (crashed rank) sets bit 32 after a 4 s spin; abort the
    // solve on the surviving ranks instead of spinning through every remaining reduction
    int sync_check_peers() {
        int32_t e = 0;
        KFSP_CUDA(cudaMemcpyAsync(&e, d_err, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        KFSP_TRY(sync());
        return (e & 32) ? KFSP_ERR_NCCL : KFSP_OK;
    }

    // ---------------------------------------------------------------- multi-GPU (dist.cuh)
    // Reducer for the next reducing launch: in the peer-memory path every reduction carries a sequence number
    // (identical on all ranks: the controller is SPMD) that tags the exchanged partials.
    Reducer next_rd() {
        Reducer r = rd;
        if (dist.suspended || dist.whole) { r.peers = nullptr; r.dist_send = nullptr; r.seq = 0; return r; }   // whole state space on this GPU: no exchange
        if (dist.p2p && dist.p2p_red) r.seq = ++dist.seq;
        return r;
    }
    // all ranks: everything enqueued so far on every GPU is complete and visible before anything enqueued later starts
    int dist_barrier() {
        if (!dist_active() || !(dist.p2p && dist.p2p_halo)) return KFSP_OK;       // the NCCL exchange step is ordered by NCCL itself
        if (!dist.p2p_red) return KFSP_ERR_UNSUPPORTED;         // flags are shared with the reduction exchange
        KFSP_TRY(launch_pdl(k_dist_barrier, 1, 32, 0, (const DistPeers*)dist.d_peers, (unsigned long long)(++dist.seq)));
        return KFSP_OK;
    }
    // Map every peer's basis and exchange area into this process (cudaIpc over NVLink).  Collective.
    int dist_setup_p2p() {
#ifdef KFSP_WITH_NCCL
        if (dist.nranks == 1 || dist.p2p || !dist.want_p2p || dist.nranks > MAX_RANKS) return KFSP_OK;
        const int P = dist.nranks;
        struct Info { cudaIpcMemHandle_t hv, hx; unsigned long long offv, offx; long long ok; };
        Info mine;
        std::memset(&mine, 0, sizeof mine);
        mine.ok = 1;
        const size_t xbytes = 4u << 20;                       // own allocation: 2 slots x P x (4 doubles + flag)
        if (!dist.xchg) {
            KFSP_CUDA(cudaMalloc(&dist.xchg, xbytes));
            KFSP_CUDA(cudaMemset(dist.xchg, 0, xbytes));
        }
        // offsets of our pointers inside their allocations (cuMemGetAddressRange through the runtime's driver entry point)
        typedef CUresult (*range_fn)(CUdeviceptr*, size_t*, CUdeviceptr);
        range_fn get_range = nullptr;
        {
            void* fp = nullptr;
            cudaDriverEntryPointQueryResult qr;
            if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &qr) == cudaSuccess && fp) get_range = (range_fn)fp;
            else cudaGetLastError();
        }
        void* base_v = d_V;
        void* base_x = dist.xchg;
        if (get_range) {
            CUdeviceptr b = 0; size_t sz = 0;
            if (get_range(&b, &sz, (CUdeviceptr)d_V) == CUDA_SUCCESS) { base_v = (void*)b; mine.offv = (unsigned long long)((char*)d_V - (char*)b); }
            if (get_range(&b, &sz, (CUdeviceptr)dist.xchg) == CUDA_SUCCESS) { base_x = (void*)b; mine.offx = (unsigned long long)((char*)dist.xchg - (char*)b); }
        }
        if (cudaIpcGetMemHandle(&mine.hv, base_v) != cudaSuccess || cudaIpcGetMemHandle(&mine.hx, base_x) != cudaSuccess) { mine.ok = 0; cudaGetLastError(); }
        Info* d_info = nullptr;
        KFSP_CUDA(cudaMalloc(&d_info, sizeof(Info) * (P + 1)));
        KFSP_CUDA(cudaMemcpyAsync(d_info + P, &mine, sizeof(Info), cudaMemcpyHostToDevice, stream));
        if (ncclAllGather(d_info + P, d_info, sizeof(Info), ncclUint8, dist.comm, stream) != ncclSuccess) return KFSP_ERR_NCCL;
        std::vector<Info> all(P);
        KFSP_CUDA(cudaMemcpyAsync(all.data(), d_info, sizeof(Info) * P, cudaMemcpyDeviceToHost, stream));
        KFSP_TRY(sync());
        DistPeers hp;
        std::memset(&hp, 0, sizeof hp);
        hp.rank = dist.rank; hp.nranks = P;
        long long ok = 1;
        for (int r = 0; r < P; ++r) ok = ok && all[r].ok;
        for (int r = 0; r < P && ok; ++r) {
            char* pv = nullptr; char* px = nullptr;
            if (r == dist.rank) { pv = (char*)d_V; px = (char*)dist.xchg; }
            else {
                void* mv = nullptr; void* mx = nullptr;
                if (cudaIpcOpenMemHandle(&mv, all[r].hv, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess ||
                    cudaIpcOpenMemHandle(&mx, all[r].hx, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
                dist.peer_base[r] = mv; dist.peer_xbase[r] = mx;
                pv = (char*)mv + all[r].offv; px = (char*)mx + all[r].offx;
            }
            hp.V[r] = (const double*)pv;
            hp.part[r] = (double*)px;
            hp.flag[r] = (unsigned long long*)(px + sizeof(double) * 2 * P * RED_W);
        }
        // every rank must take the same path: agree on success
        long long* d_ok = (long long*)(d_info);
        KFSP_CUDA(cudaMemcpyAsync(d_ok + P, &ok, sizeof(long long), cudaMemcpyHostToDevice, stream));
        if (ncclAllGather(d_ok + P, d_ok, 1, ncclInt64, dist.comm, stream) != ncclSuccess) return KFSP_ERR_NCCL;
        std::vector<long long> oks(P);
        KFSP_CUDA(cudaMemcpyAsync(oks.data(), d_ok, sizeof(long long) * P, cudaMemcpyDeviceToHost, stream));
        KFSP_TRY(sync());
        cudaFree(d_info);
        for (int r = 0; r < P; ++r) ok = ok && oks[r];
        if (!ok) {
            std::fprintf(stderr, "libkfsp: peer-memory mapping unavailable, using the NCCL exchange path\n");
            return KFSP_OK;
        }
        hp.halo_owner = dist.halo_owner; hp.halo_lidx = dist.halo_lidx; hp.err = d_err;
        hp.ll = 1;
        if (const char* ev = std::getenv("KFSP_DIST_LL")) hp.ll = std::atoi(ev) != 0;      // 0: data, system fence, flag (A/B)
        for (int r = 0; r <= P; ++r) hp.rb[r] = dist.rb[r];
        if (!dist.d_stat) {
            KFSP_CUDA(cudaMalloc(&dist.d_stat, 4 * sizeof(unsigned long long)));
            KFSP_CUDA(cudaMemset(dist.d_stat, 0, 4 * sizeof(unsigned long long)));
        }
        hp.stat = dist.d_stat;
        if (!dist.d_peers) KFSP_CUDA(cudaMalloc(&dist.d_peers, sizeof(DistPeers)));
        KFSP_CUDA(cudaMemcpy(dist.d_peers, &hp, sizeof hp, cudaMemcpyHostToDevice));
        if (dist.p2p_red) { rd.peers = dist.d_peers; rd.dist_send = nullptr; }
        dist.p2p = true;
        dist.seq = 0;
        return KFSP_OK;
#else
        return KFSP_OK;
#endif
    }
    int dist_init(int rank, int nranks, const uint8_t* id) {
        if (nranks == 1) return KFSP_OK;
#ifdef KFSP_WITH_NCCL
        if (ld > 0) return KFSP_ERR_ARG;                     // must come before the state space is created
        KFSP_CUDA(cudaSetDevice(device));
        ncclUniqueId u;
        std::memcpy(&u, id, sizeof u);
        if (ncclCommInitRank(&dist.comm, nranks, u, rank) != ncclSuccess) return KFSP_ERR_NCCL;
        dist.rank = rank; dist.nranks = nranks;
        // adaptive state sets keep the whole state space on every rank and partition the rows of the Krylov loop (dist.cuh)
        dist.repl = opt.spmv_variant != 1 && (opt.enable_expand || opt.enable_drop || opt.n_init_onestep > 0);
        if (const char* ev = std::getenv("KFSP_DIST_P2P")) {
            const int v = std::atoi(ev);
            dist.want_p2p = v != 0; dist.p2p_red = v == 1 || v == 2; dist.p2p_halo = v == 1 || v == 3;
        }
        KFSP_CUDA(cudaMalloc(&dist.red_send, sizeof(double) * RED_W));
        KFSP_CUDA(cudaMalloc(&dist.red_recv, sizeof(double) * RED_W * nranks));
        KFSP_CUDA(cudaMemset(dist.red_send, 0, sizeof(double) * RED_W));
        rd.dist_send = dist.red_send;
        return KFSP_OK;
#else
        (void)rank; (void)id;
        return KFSP_ERR_UNSUPPORTED;
#endif
    }
    // all-gather the ranks' double-double partials, merge in rank order, run the reduction's epilogue
    int dist_finalize(const EpiArgs& ea, int nv) {
        if (!dist_active() || (dist.p2p && dist.p2p_red)) return KFSP_OK;     // peer-memory path: exchanged inside the reducing kernel
#ifdef KFSP_WITH_NCCL
        if (ncclAllGather(dist.red_send, dist.red_recv, RED_W, ncclFloat64, dist.comm, stream) != ncclSuccess) return KFSP_ERR_NCCL;
        KFSP_LAUNCH(k_dist_finalize, 1, 32, 0, ea, nv, (const double*)dist.red_recv, dist.nranks, d_ctl);
        dist.reductions += 1;
#else
        (void)ea; (void)nv;
#endif
        return KFSP_OK;
    }
    // SpMV exchange step: send the x entries other ranks' rows refer to, receive ours into dist.halo
    int dist_halo_exchange(const double* x) {
#ifdef KFSP_WITH_NCCL
        if (dist.n_send > 0)
            KFSP_LAUNCH(k_dist_pack, grid_for(dist.n_send), VEC_THREADS, 0, x, (const int32_t*)dist.send_idx, dist.n_send, dist.sendbuf);
        if (ncclGroupStart() != ncclSuccess) return KFSP_ERR_NCCL;
        for (int p = 0; p < dist.nranks; ++p) {
            if (p == dist.rank) continue;
            const int64_t sc = dist.send_off[p + 1] - dist.send_off[p], rc = dist.recv_off[p + 1] - dist.recv_off[p];
            if (sc > 0 && ncclSend(dist.sendbuf + dist.send_off[p], (size_t)sc, ncclFloat64, p, dist.comm, stream) != ncclSuccess) return KFSP_ERR_NCCL;
            if (rc > 0 && ncclRecv(dist.halo + dist.recv_off[p], (size_t)rc, ncclFloat64, p, dist.comm, stream) != ncclSuccess) return KFSP_ERR_NCCL;
        }
        if (ncclGroupEnd() != ncclSuccess) return KFSP_ERR_NCCL;
        dist.halo_exchanges += 1;
        dist.halo_bytes += 8 * dist.n_send;
        return KFSP_OK;
#else
        (void)x;
        return KFSP_ERR_UNSUPPORTED;
#endif
    }
    // MATRIX_STARTER for a partitioned, fixed state set: all ranks get the full state list, build the
    // full hash table and their own rows, then agree on the halo plan.
    int dist_fsp_init(int64_t n_global, const int32_t* states_host) {
#ifdef KFSP_WITH_NCCL
        if (host_prop) return KFSP_ERR_UNSUPPORTED;     // k_dist_build_rows evaluates the propensities of this rank's rows on the device
        KFSP_TRY(ensure_state_space());
        if (n_global < dist.nranks || n_global > opt.max_states) return KFSP_ERR_BAD_SIZES;
        const int P = dist.nranks;
        dist.n_global = n_global;
        dist.lo = part_lo(n_global, P, dist.rank);
        dist.hi = part_lo(n_global, P, dist.rank + 1);
        const int64_t nloc = dist.hi - dist.lo;
        if (nloc > ld) return KFSP_ERR_BAD_SIZES;
        KFSP_CUDA(cudaMemcpyAsync(d_states, states_host, sizeof(int32_t) * n_global * S, cudaMemcpyHostToDevice, stream));
        KFSP_CUDA(cudaMemsetAsync(d_w, 0, sizeof(double) * ld, stream));
        KFSP_CUDA(cudaMemsetAsync(d_table, 0xFF, sizeof(int32_t) * table_size, stream));
        KFSP_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int32_t), stream));
        n = n_global;                                   // view over the global list for validation / table
        FspView f = view();
        KFSP_LAUNCH(k_validate_states, grid_for(n_global * S), VEC_THREADS, 0, d_states, S, n_global, opt.max_molecules, d_err);
        KFSP_LAUNCH(k_insert_states, grid_for(n_global), VEC_THREADS, 0, f, (int64_t)0, n_global, d_err);
        KFSP_LAUNCH(k_dist_build_rows, grid_for(nloc), VEC_THREADS, 0, f, dist.lo, nloc);
        int32_t e = 0;
        KFSP_TRY(read_err(&e));
        if (e) { n = 0; return err_to_status(e); }
        // halo plan
        const size_t a_g = align_up(sizeof(int32_t) * n_global);
        KFSP_TRY(ensure_scratch(2 * a_g + align_up(sizeof(int32_t) * scan_buf_ints(n_global)) + align_up(sizeof(int64_t) * 4 * (P + 2))));
        int32_t* flag = (int32_t*)d_scratch;
        int32_t* pos = (int32_t*)(d_scratch + a_g);
        int32_t* tb = (int32_t*)(d_scratch + 2 * a_g);
        int64_t* d_small = (int64_t*)(d_scratch + 2 * a_g + align_up(sizeof(int32_t) * scan_buf_ints(n_global)));
        KFSP_CUDA(cudaMemsetAsync(flag, 0, sizeof(int32_t) * n_global, stream));
        KFSP_LAUNCH(k_dist_mark_remote, grid_for(nloc * R), VEC_THREADS, 0, (const int32_t*)d_pred, ld, R, nloc, dist.lo, dist.hi, flag);
        int64_t nh = 0;
        KFSP_TRY(exclusive_scan(flag, pos, n_global, tb, &nh));
        dist.n_halo = nh;
        int32_t* halo_g = nullptr;
        KFSP_CUDA(cudaMalloc(&halo_g, sizeof(int32_t) * std::max<int64_t>(nh, 1)));
        KFSP_LAUNCH(k_dist_halo_list, grid_for(n_global), VEC_THREADS, 0, (const int32_t*)flag, (const int32_t*)pos, n_global, halo_g);
        KFSP_LAUNCH(k_dist_remap, grid_for(nloc * R), VEC_THREADS, 0, d_pred, ld, R, nloc, dist.lo, dist.hi, (const int32_t*)pos);
        // segments of the halo list by owner
        std::vector<int64_t> bound(P + 1), off(P + 1);
        for (int r = 0; r <= P; ++r) bound[r] = part_lo(n_global, P, r);
        int64_t* d_bound = d_small;
        int64_t* d_off = d_small + (P + 2);
        KFSP_CUDA(cudaMemcpyAsync(d_bound, bound.data(), sizeof(int64_t) * (P + 1), cudaMemcpyHostToDevice, stream));
        KFSP_LAUNCH(k_dist_bounds, 1, 64, 0, (const int32_t*)halo_g, nh, (const int64_t*)d_bound, P, d_off);
        KFSP_CUDA(cudaMemcpyAsync(off.data(), d_off, sizeof(int64_t) * (P + 1), cudaMemcpyDeviceToHost, stream));
        KFSP_TRY(sync());
        dist.recv_off = off;
        // tell every owner how many of its rows we want; learn how many each rank wants from us
        int64_t* d_want = d_small + 2 * (P + 2);
        int64_t* d_all = nullptr;
        KFSP_CUDA(cudaMalloc(&d_all, sizeof(int64_t) * P * P));
        std::vector<int64_t> want(P);
        for (int o = 0; o < P; ++o) want[o] = off[o + 1] - off[o];
        KFSP_CUDA(cudaMemcpyAsync(d_want, want.data(), sizeof(int64_t) * P, cudaMemcpyHostToDevice, stream));
        if (ncclAllGather(d_want, d_all, (size_t)P, ncclInt64, dist.comm, stream) != ncclSuccess) return KFSP_ERR_NCCL;
        std::vector<int64_t> all((size_t)P * P);
        KFSP_CUDA(cudaMemcpyAsync(all.data(), d_all, sizeof(int64_t) * P * P, cudaMemcpyDeviceToHost, stream));
        KFSP_TRY(sync());
        dist.send_off.assign(P + 1, 0);
        for (int r = 0; r < P; ++r) dist.send_off[r + 1] = dist.send_off[r] + all[(size_t)r * P + dist.rank];
        dist.n_send = dist.send_off[P];
        cudaFree(dist.send_idx); cudaFree(dist.sendbuf); cudaFree(dist.halo);
        KFSP_CUDA(cudaMalloc(&dist.send_idx, sizeof(int32_t) * std::max<int64_t>(dist.n_send, 1)));
        KFSP_CUDA(cudaMalloc(&dist.sendbuf, sizeof(double) * std::max<int64_t>(dist.n_send, 1)));
        KFSP_CUDA(cudaMalloc(&dist.halo, sizeof(double) * std::max<int64_t>(nh, 1)));
        if (ncclGroupStart() != ncclSuccess) return KFSP_ERR_NCCL;
        for (int p = 0; p < P; ++p) {
            if (p == dist.rank) continue;
            const int64_t rc = off[p + 1] - off[p], sc = dist.send_off[p + 1] - dist.send_off[p];
            if (rc > 0 && ncclSend(halo_g + off[p], (size_t)rc, ncclInt32, p, dist.comm, stream) != ncclSuccess) return KFSP_ERR_NCCL;
            if (sc > 0 && ncclRecv(dist.send_idx + dist.send_off[p], (size_t)sc, ncclInt32, p, dist.comm, stream) != ncclSuccess) return KFSP_ERR_NCCL;
        }
        if (ncclGroupEnd() != ncclSuccess) return KFSP_ERR_NCCL;
        if (dist.n_send > 0) KFSP_LAUNCH(k_dist_to_local, grid_for(dist.n_send), VEC_THREADS, 0, dist.send_idx, dist.n_send, dist.lo);
        cudaFree(dist.halo_owner); cudaFree(dist.halo_lidx);
        KFSP_CUDA(cudaMalloc(&dist.halo_owner, sizeof(int32_t) * std::max<int64_t>(nh, 1)));
        KFSP_CUDA(cudaMalloc(&dist.halo_lidx, sizeof(int32_t) * std::max<int64_t>(nh, 1)));
        if (nh > 0) KFSP_LAUNCH(k_dist_halo_owner, grid_for(nh), VEC_THREADS, 0, (const int32_t*)halo_g, nh, n_global, P, dist.halo_owner, dist.halo_lidx);
        if (dist.d_peers) {                                    // re-initialisation after the peer table was built
            KFSP_CUDA(cudaMemcpyAsync(&dist.d_peers->halo_owner, &dist.halo_owner, sizeof(void*), cudaMemcpyHostToDevice, stream));
            KFSP_CUDA(cudaMemcpyAsync(&dist.d_peers->halo_lidx, &dist.halo_lidx, sizeof(void*), cudaMemcpyHostToDevice, stream));
        }
        KFSP_TRY(sync());
        cudaFree(halo_g);
        cudaFree(d_all);
        n = nloc;                                       // from here on kernels see this rank's rows only
        return KFSP_OK;
#else
        (void)n_global; (void)states_host;
        return KFSP_ERR_UNSUPPORTED;
#endif
    }

    // ---------------------------------------------------------------- matrix-free lattice (lattice.cuh)
    // z-chunks: k work items per resident CTA, never k+1 for some while the rest idle; k = 4 when that leaves
    // chunks of >= 96 planes (2 halo rows are re-read and PF rows start un-prefetched per chunk), else fewer
    void lattice_chunking(int wave, int64_t ncb, int* zc, int* g) const {
        const int nzl = lat.zhi - lat.zlo;
        int z = nzl;
        for (int k = 4; k >= 1; --k) {
            int64_t nzc = std::max<int64_t>(1, (k * (int64_t)wave) / ncb);
            nzc = std::min<int64_t>(nzc, nzl);
            z = (int)((nzl + nzc - 1) / nzc);
            while (z < nzl && ncb * ((nzl + z - 1) / z) > k * (int64_t)wave) ++z;      // rounding must not push it past k waves
            if (z >= 96 || k == 1) break;
        }
        *zc = z;
        const int64_t items = ncb * ((nzl + z - 1) / z);
        *g = (int)std::min<int64_t>(items, wave);
    }
    // the stencil kernel k_spmv_bd2; fin: finalise the operand on the fly (a.src = scratch column, a.xout = the column)
    template <int MODE>
    int spmv_bd2(Bd2Args a, bool fin) {
        const int bd_ord = lattice_bd2_order(lat);
        if (bd_ord < 0) return KFSP_ERR_UNSUPPORTED;
        const bool halo = dist.nranks > 1;
        Reducer r = (MODE != 0 || fin) ? next_rd() : rd;
        if (halo && !r.peers) r.peers = dist.d_peers;
        void (*kb)(const Lattice, const Bd2Args, Reducer, SweepCtl*) = nullptr;
        int ts = 0;
        for (int k = 0; k < 4; ++k) ts |= (lat.sp[k] == 1 ? 1 : 0) << k;
        constexpr int FINOK = MODE != 0;                        // the plain product never finalises
#define KFSP_BD2(T) case T:                                                                                              \
            kb = fin ? (bd_ord == 0 ? k_spmv_bd2<0, T, MODE, FINOK, BD2_PF, BD2_MINB> : k_spmv_bd2<1, T, MODE, FINOK, BD2_PF, BD2_MINB>)   \
                     : (bd_ord == 0 ? k_spmv_bd2<0, T, MODE, 0, BD2_PF, BD2_MINB> : k_spmv_bd2<1, T, MODE, 0, BD2_PF, BD2_MINB>);            \
            break
        switch (ts) {
#ifdef KFSP_FAST_BUILD                                      // developer builds: only the table pattern of the toggle models
            KFSP_BD2(9);
            default: return KFSP_ERR_UNSUPPORTED;
#else
            KFSP_BD2(0); KFSP_BD2(1); KFSP_BD2(2); KFSP_BD2(3); KFSP_BD2(4); KFSP_BD2(5); KFSP_BD2(6); KFSP_BD2(7);
            KFSP_BD2(8); KFSP_BD2(9); KFSP_BD2(10); KFSP_BD2(11); KFSP_BD2(12); KFSP_BD2(13); KFSP_BD2(14); KFSP_BD2(15);
#endif
        }
#undef KFSP_BD2
        if (fin && MODE == 0) return KFSP_ERR_ARG;
        const int64_t ncb_bd = (lat.plane + BD2_CBW - 1) / BD2_CBW;
        const int wave = wave_grid((const void*)kb, (int64_t)1 << 40);
        int zc, g;
        lattice_chunking(wave, ncb_bd, &zc, &g);
        const int nzl = lat.zhi - lat.zlo;
        if (bd2_zc > 0) {                                       // KFSP_BD2_ZC: forced chunk length (tests / tuning)
            zc = std::min(bd2_zc, nzl);
            g = (int)std::min<int64_t>(ncb_bd * ((nzl + zc - 1) / zc), wave);
        }
        if (zc > BD2_ZT) {                                      // the staged y-tables bound the chunk length
            zc = BD2_ZT;
            g = (int)std::min<int64_t>(ncb_bd * ((nzl + zc - 1) / zc), wave);
        }
        a.zc = zc;
        a.halo = halo ? 1 : 0;
        a.sync_every = bd2_sync;
        a.l2_ahead = bd2_ahead;
        a.off_g = (d_V && a.g) ? (int64_t)(a.g - d_V) : 0;
        a.off_f = (d_V && a.f) ? (int64_t)(a.f - d_V) : 0;
        if (!a.g) { a.g = a.src; a.has_g = 0; }
        if (!a.f) { a.f = a.g; a.has_f = 0; }
        a.ea.has_g = a.has_g;
        a.off_src = d_V ? (int64_t)(a.src - d_V) : 0;
        return launch_pdl(kb, g, VEC_THREADS, 0, lat, a, r, d_ctl);
    }
    template <int MODE>
    int spmv_box(const double* x, double* y, const double* g, const EpiArgs& ea, int cg) {
        const bool halo = dist.nranks > 1;
        // column blocks: the plane split evenly into the fewest blocks of at most 256 columns
        const int64_t ncb0 = (lat.plane + VEC_THREADS - 1) / VEC_THREADS;
        const int cbw = (int)((lat.plane + ncb0 - 1) / ncb0);
        const int64_t ncb = (lat.plane + cbw - 1) / cbw;
        // 2-species one-molecule-step networks in the reference's reaction orders (config 5, the toggle models): stencil kernel
        if (box_tune < 10 && lattice_bd2_order(lat) >= 0) {
            Bd2Args a;
            std::memset(&a, 0, sizeof a);
            a.src = x; a.y = y; a.g = g; a.has_g = g ? 1 : 0; a.cg = cg; a.cf = -1; a.ea = ea;
            return spmv_bd2<MODE>(a, false);
        }
        const Reducer r = MODE != 0 ? next_rd() : rd;
        Reducer r2 = r;
        if (halo && !r2.peers) r2.peers = dist.d_peers;
        void (*kern)(const Lattice, int, int, const double*, double*, const double*, Reducer, SweepCtl*, EpiArgs, int, int64_t) = nullptr;
        const int tune = box_tune >= 10 ? box_tune - 10 : box_tune;
#define KFSP_BOX_PICK(RR)                                                                                                   \
        kern = halo ? (S == 2 ? k_spmv_box<RR, 2, MODE, 2, 1, 6, 3> : k_spmv_box<RR, 0, MODE, 2, 1, 6, 3>)                      \
             : S != 2 ? k_spmv_box<RR, 0, MODE, 0, 1, 6, 4>                                                                     \
             : tune == 1 ? k_spmv_box<RR, 2, MODE, 0, 0, 1, 4> : k_spmv_box<RR, 2, MODE, 0, 1, 6, 4>
        switch (R) {
        case 2: KFSP_BOX_PICK(2); break;
        case 4: KFSP_BOX_PICK(4); break;
        case 6: KFSP_BOX_PICK(6); break;
        case 8: KFSP_BOX_PICK(8); break;
        default: return KFSP_ERR_UNSUPPORTED;
        }
#undef KFSP_BOX_PICK
        const int wave = wave_grid((const void*)kern, (int64_t)1 << 40);
        int zc, gsz;
        lattice_chunking(wave, ncb, &zc, &gsz);
        return launch_pdl(kern, gsz, VEC_THREADS, 0, lat, zc, cbw, x, y, g ? g : x, r2, d_ctl, ea, cg, (int64_t)(d_V ? x - d_V : 0));
    }
    // kfsp_fsp_init_box: the projection is the lattice [0,B_1) x ... x [0,B_S) in natural order.
    int fsp_init_box(const int32_t* bounds) {
        if (!have_model) return KFSP_ERR_NO_MODEL;
        if (opt.spmv_variant != 1) return KFSP_ERR_ARG;
        if (opt.enable_expand || opt.enable_drop || opt.n_init_onestep) return KFSP_ERR_ARG;      // fixed state set
        if (S < 2 || R > BOX_MAX_R || host_prop) return KFSP_ERR_UNSUPPORTED;
        KFSP_CUDA(cudaSetDevice(device));
        Lattice L;
        std::memset(&L, 0, sizeof L);
        L.S = S; L.R = R;
        int64_t total = 1;
        for (int s = 0; s < S; ++s) {
            if (bounds[s] < 1 || bounds[s] > opt.max_molecules + 1) return KFSP_ERR_BAD_STATE;
            L.B[s] = bounds[s];
            L.stride[s] = total;
            total *= bounds[s];
            if (total > opt.max_states || total > 2000000000LL) return KFSP_ERR_BAD_SIZES;
        }
        L.plane = L.stride[S - 1];
        L.nz = L.B[S - 1];
        for (int k = 0; k < R; ++k) {
            if (h_dm.table_species[k] < 0 || !h_dm.table[k]) return KFSP_ERR_UNSUPPORTED;   // a propensity reads several species
            L.sp[k] = h_dm.table_species[k];
            L.tab[k] = h_dm.table[k];
            for (int s = 0; s < S; ++s) L.nu[k][s] = h_dm.stoich[k * S + s];
        }
        const int P = dist.nranks;
        if (P > 1 && !dist.want_p2p) return KFSP_ERR_UNSUPPORTED;       // halo rows are read straight from the neighbours' HBM
        if (L.nz < P) return KFSP_ERR_BAD_SIZES;
        int64_t thick = 0;
        for (int r = 0; r <= P; ++r) L.zb[r] = (int32_t)((int64_t)L.nz * r / P);
        for (int r = 0; r < P; ++r) thick = std::max<int64_t>(thick, L.zb[r + 1] - L.zb[r]);
        L.zlo = L.zb[dist.rank]; L.zhi = L.zb[dist.rank + 1];
        const int64_t nloc = L.plane * (L.zhi - L.zlo);
        const int64_t cap = ((L.plane * thick + 63) / 64) * 64;          // same leading dimension on every rank
        if (cap != ld || !box || nloc != n) {
            if (dist.nranks > 1 && d_V) return KFSP_ERR_UNSUPPORTED;     // peers hold mappings of this rank's basis
            KFSP_CUDA(wait_stream());
            free_state_space();
            KFSP_CUDA(cudaMalloc(&d_w, sizeof(double) * cap));
            ld = cap;
        }
        KFSP_CUDA(cudaMemsetAsync(d_w, 0, sizeof(double) * ld, stream));
        KFSP_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int32_t), stream));
        lat = L;
        box = true;
        n = nloc;
        states_cap = 0;
        dist.n_global = total; dist.lo = L.plane * L.zlo; dist.hi = L.plane * L.zhi;
        dist.n_halo = P > 1 ? L.plane * ((dist.rank > 0) + (dist.rank < P - 1)) : 0;
        return KFSP_OK;
    }
    // d_states of this rank's rows (generated on demand: the lattice itself needs no state list)
    int box_states() {
        if (d_states) return KFSP_OK;
        KFSP_CUDA(cudaMalloc(&d_states, sizeof(int32_t) * std::max<int64_t>(n, 1) * S));
        KFSP_LAUNCH(k_box_gen_states, grid_for(n), VEC_THREADS, 0, lat, dist.lo, n, d_states);
        return KFSP_OK;
    }
    // kfsp_fsp_init / kfsp_solve with spmv_variant = 1: the caller's state list must be a lattice in natural
    // order.  Bounds come from the last state; this rank's rows are verified on the device.
    // defer = true (kfsp_solve): the list is uploaded and verified on the side stream WHILE the solve runs -- the lattice
    // arithmetic does not read it -- and box_check_finish() delivers the verdict afterwards; `after` (optional) is an event
    // of the main stream the upload should wait for (the p0 upload, which the solve does need, goes over the same link first).
    int box_init_from_states(int64_t count, const int32_t* states_host, bool defer = false) {
        if (!have_model) return KFSP_ERR_NO_MODEL;
        if (count < 1) return KFSP_ERR_BAD_SIZES;
        KFSP_TRY(box_check_finish());
        int32_t bounds[KFSP_MAX_SPECIES];
        int64_t total = 1;
        for (int s = 0; s < S; ++s) {
            const int32_t top = states_host[(count - 1) * S + s];
            if (top < 0 || top > opt.max_molecules) return KFSP_ERR_BAD_STATE;
            bounds[s] = top + 1;
            total *= bounds[s];
            if (total > count) return KFSP_ERR_UNSUPPORTED;
        }
        if (total != count) return KFSP_ERR_UNSUPPORTED;          // not a full box: use spmv_variant = 0
        KFSP_TRY(fsp_init_box(bounds));
        if (!d_states) KFSP_CUDA(cudaMalloc(&d_states, sizeof(int32_t) * n * S));     // n is fixed while ld is (fsp_init_box)
        box_src = states_host + dist.lo * S;
        if (defer) { box_check_pending = true; box_check_started = false; return KFSP_OK; }
        KFSP_TRY(box_check_start(stream));
        return box_check_finish();
    }
    const int32_t* box_src = nullptr;
    bool box_check_started = false;
    int box_check_start(cudaStream_t st) {
        KFSP_CUDA(cudaMemsetAsync(d_chk, 0, sizeof(int32_t), st));
        KFSP_CUDA(cudaMemcpyAsync(d_states, box_src, sizeof(int32_t) * n * S, cudaMemcpyHostToDevice, st));
        k_box_check_states<<<grid_for(n), VEC_THREADS, 0, st>>>(lat, dist.lo, n, (const int32_t*)d_states, d_chk);
        KFSP_TRY(check_launch());
        box_check_pending = true;
        box_check_started = true;
        return KFSP_OK;
    }
    // start the deferred upload on the side stream once everything enqueued so far on the main stream has gone over the link
    int box_check_overlap() {
        if (!box_check_pending || box_check_started) return KFSP_OK;
        KFSP_CUDA(cudaEventRecord(ev_side, stream));
        KFSP_CUDA(cudaStreamWaitEvent(stream2, ev_side, 0));
        return box_check_start(stream2);
    }
    int box_check_finish() {
        if (!box_check_pending) return KFSP_OK;
        if (!box_check_started) KFSP_TRY(box_check_start(stream));
        box_check_pending = false;
        int32_t bad = 0;
        KFSP_CUDA(cudaStreamSynchronize(stream2));
        KFSP_CUDA(cudaMemcpyAsync(&bad, d_chk, sizeof(int32_t), cudaMemcpyDeviceToHost, stream));
        KFSP_CUDA(wait_stream());
        if (bad) { n = 0; box = false; return KFSP_ERR_UNSUPPORTED; }
        return KFSP_OK;
    }

    int solve(double T, double fsptol, double krytol, int itrace, kfsp_stats* stats);
};

}  // namespace kfsp

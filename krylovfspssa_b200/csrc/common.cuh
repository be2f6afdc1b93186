// Shared device/host definitions for libkfsp (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#include "../../include/kfsp.h"

namespace kfsp {

constexpr int KFSP_MAX_SPECIES = 8;       // states are held in registers as int32[8]
constexpr int KFSP_MAX_REACTIONS = 32;
constexpr int KFSP_MAX_PARAMS = 64;
constexpr int KFSP_MAX_CODE = 1024;       // total byte-code words over all reactions
constexpr int KFSP_MAX_IMMED = 256;
constexpr int KFSP_STACK = 16;            // evaluation stack of the propensity interpreter

// internal index conventions (0-based); the C ABI converts to the Fortran ones
constexpr int32_t IDX_ABSENT = -1;        // legal neighbour, not in the projection  (ADJ = 0)
constexpr int32_t IDX_ILLEGAL = -2;       // neighbour has a negative count          (ADJ = -1)
constexpr int32_t SLOT_EMPTY = -1;

#define KFSP_CUDA(call)                                                                        \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            std::fprintf(stderr, "libkfsp: CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e__), \
                         __FILE__, __LINE__, cudaGetErrorString(e__));                         \
            return KFSP_ERR_CUDA;                                                              \
        }                                                                                      \
    } while (0)

#define KFSP_TRY(call)                  \
    do {                                \
        int s__ = (call);               \
        if (s__ != KFSP_OK) return s__; \
    } while (0)

// Model as the kernels see it (lives in global memory, read through the read-only path).
struct DeviceModel {
    int32_t S, R, P;
    int32_t max_molecules;
    int32_t stoich[KFSP_MAX_REACTIONS * KFSP_MAX_SPECIES];   // [k*S + s]
    int32_t code_begin[KFSP_MAX_REACTIONS + 1];
    int32_t immed_begin[KFSP_MAX_REACTIONS + 1];
    int32_t code[KFSP_MAX_CODE];
    double immed[KFSP_MAX_IMMED];
    double params[KFSP_MAX_PARAMS];
    // Propensities with a transcendental operation that depend on ONE species are tabulated on the
    // host (same libm as the Fortran/C++ host and the oracle) for counts 0..max_molecules.
    int32_t table_species[KFSP_MAX_REACTIONS];      // species index, or -1 = interpret the byte code
    const double* table[KFSP_MAX_REACTIONS];         // device pointers, max_molecules+1 entries each
};

// Factored propensities of the index-only SpMV (spmv_variant = 2, krylov.cuh: k_spmv_idx).  model_host.h: factor_program
// splits a_k into terms that read at most one species each -- tabulated on the host over the count 0..max_molecules --
// combined by + - * (postfix).  The common shapes are decided at set-up time:
//   FAC_ONE  a = T0[x_s0]                         (every single-species propensity: c*x, c*x*(x-1)/2, Hill functions ...)
//   FAC_MUL2 a = T0[x_s0] * T1[x_s1]              (bimolecular mass action c*x*y)
//   FAC_MUL3 a = (T0[x_s0] * T1[x_s1]) * T2[x_s2]
//   FAC_GEN  the postfix program over up to FAC_MAX_TERMS terms
constexpr int FAC_MAX_TERMS = 4;
constexpr int FAC_MAX_OPS = 8;
enum FacShape : int8_t { FAC_ONE = 0, FAC_MUL2, FAC_MUL3, FAC_GEN };
struct FacModel {
    int32_t S, R;
    int8_t shape[KFSP_MAX_REACTIONS];
    int8_t nops[KFSP_MAX_REACTIONS];
    int8_t sp[KFSP_MAX_REACTIONS][FAC_MAX_TERMS];        // species the term's table runs over (constants: species 0, stride 0)
    int8_t use[KFSP_MAX_REACTIONS][FAC_MAX_TERMS];       // 1: index by the species count, 0: constant (entry 0)
    int8_t ops[KFSP_MAX_REACTIONS][FAC_MAX_OPS];         // FAC_GEN: t >= 0 pushes term t, -3/-4/-5 = add/sub/mul, -2 = neg
    int8_t nu[KFSP_MAX_REACTIONS][KFSP_MAX_SPECIES];     // stoichiometry
    const double* tab[KFSP_MAX_REACTIONS][FAC_MAX_TERMS];
};

// a_k at the row's state (sgn = 0) or at its predecessor x - nu_k (sgn = 1)
template <int STRIDE>
__device__ __forceinline__ double fac_term(const FacModel& F, int k, int t, const int32_t* sst, int sgn) {
    const int sp = F.sp[k][t];
    const int c = (sst[sp * STRIDE] - sgn * (int)F.nu[k][sp]) * (int)F.use[k][t];
    return __ldg(F.tab[k][t] + c);
}
// GEN = 0: no reaction of the model needs the postfix program -- straight-line code (the shape tests are uniform and
// guard single instructions), so that the loads of several reactions can be in flight together
template <int STRIDE, int GEN>
__device__ __forceinline__ double fac_eval(const FacModel& F, int k, const int32_t* sst, int sgn) {
    const int shp = F.shape[k];
    if (GEN == 0 || shp != FAC_GEN) {
        double a = fac_term<STRIDE>(F, k, 0, sst, sgn);
        if (shp >= FAC_MUL2) a = __dmul_rn(a, fac_term<STRIDE>(F, k, 1, sst, sgn));
        if (shp >= FAC_MUL3) a = __dmul_rn(a, fac_term<STRIDE>(F, k, 2, sst, sgn));
        return a;
    }
    double stk[FAC_MAX_OPS];
    int top = -1;
    for (int q = 0; q < F.nops[k]; ++q) {
        const int op = F.ops[k][q];
        if (op >= 0) stk[++top] = fac_term<STRIDE>(F, k, op, sst, sgn);
        else if (op == -2) stk[top] = -stk[top];
        else {
            const double b = stk[top--];
            stk[top] = op == -3 ? __dadd_rn(stk[top], b) : op == -4 ? __dsub_rn(stk[top], b) : __dmul_rn(stk[top], b);
        }
    }
    return stk[0];
}
// Philox4x32-10 (Salmon et al., SC'11): one counter-based sub-stream per SSA trajectory.
// counter = (jump, j0 (1-based start index), call number, 0), key = 64-bit seed.
__host__ __device__ inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__host__ __device__ inline void philox_uniform2(uint64_t seed, uint32_t call_no, uint32_t j0, uint32_t jump,
                                                double* r1, double* r2) {
    uint32_t c[4] = {jump, j0, call_no, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint64_t a = ((uint64_t)c[1] << 32) | c[0];
    const uint64_t b = ((uint64_t)c[3] << 32) | c[2];
    *r1 = (double)(a >> 11) * (1.0 / 9007199254740992.0);
    *r2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);
}

// Hash of a state vector -> table slot.
__host__ __device__ inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}
__host__ __device__ inline uint64_t hash_state(const int32_t* st, int S) {
    uint64_t h = 0x9E3779B97F4A7C15ULL;
    for (int s = 0; s < S; ++s) h = mix64(h ^ (uint64_t)(uint32_t)st[s]) + 0x632BE59BD9B4E019ULL * (uint64_t)(s + 1);
    return h;
}

}  // namespace kfsp

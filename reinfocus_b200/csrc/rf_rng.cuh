// xoroshiro128+ streams compatible with numba.cuda.random (the RNG the reference calls at
// graphics/random.py:18,33), and a parallel state initialiser.
//
// numba builds state i as "state i-1 jumped 2**64 steps" in a sequential CPU loop
// (init_xoroshiro128p_states_cpu: 0.33 us/state -> ~2 min at 4096 envs x 300 x 300).
// The jump is a linear map J over GF(2)^128, so state i = J^i state 0. We build J^(2^k)
// on the host once, then fill the array by doubling on the GPU:
//     states[2^k + i] = J^(2^k) states[i],  i < 2^k
// one 128x128 bit-matrix * vector product per state in total.
#pragma once

#include <cstdint>

#include "../../include/reinfocus_b200.h"

namespace rf {

struct RngState {
    uint64_t s0, s1;
};

__host__ __device__ __forceinline__ uint64_t rotl64(uint64_t x, int k) {
    return (x << k) | (x >> (64 - k));
}

// numba/cuda/random.py xoroshiro128p_next
__host__ __device__ __forceinline__ uint64_t rng_next(RngState &s) {
    const uint64_t s0 = s.s0;
    uint64_t s1 = s.s1;
    const uint64_t result = s0 + s1;
    s1 ^= s0;
    s.s0 = rotl64(s0, 55) ^ s1 ^ (s1 << 14);
    s.s1 = rotl64(s1, 36);
    return result;
}

// numba/cuda/random.py uint64_to_unit_float32: float32(float64(x >> 11) * 2**-53).
// (x >> 11) < 2^53 is exact in float64 and the scale is a power of two, so the only
// rounding is the final float64 -> float32 one; converting the integer straight to
// float32 (round-to-nearest-even) and scaling gives the same bits.
__device__ __forceinline__ float rng_uniform(RngState &s) {
    const uint64_t r = rng_next(s);
    return __ull2float_rn(r >> 11) * 0x1p-53f;
}

// The same generator on explicit 32-bit halves, the form the tracer's hot loops use.
// On sm_100a the ALU pipe (LOP3 / SHF / IADD3: 64 lanes/clk/SM) is what bounds
// xoroshiro128+, so the formulation minimises ALU-pipe instructions: rotations are single
// funnel shifts, the carry-in add and the plain left shift are left to IMAD (FMA pipe),
// and the ">> 11" of the output is a mask (one LOP3 instead of two shifts) whose 2^11 is
// folded into the float scale.
struct Rng32 {
    uint32_t a, b, c, d;  // s0 = b:a, s1 = d:c
};

__device__ __forceinline__ Rng32 rng32_load(const RngState *p) {
    const uint4 v = *reinterpret_cast<const uint4 *>(p);
    return Rng32{v.x, v.y, v.z, v.w};
}

__device__ __forceinline__ void rng32_store(RngState *p, const Rng32 &s) {
    *reinterpret_cast<uint4 *>(p) = make_uint4(s.a, s.b, s.c, s.d);
}

// Advances the state and returns RN_float32(k) * 2^11 where k = (s0 + s1) >> 11, i.e. the
// uniform sample scaled by 2^64: uniform = result * 2^-64 (exact scaling).
__device__ __forceinline__ float rng32_next_scaled(Rng32 &s) {
    uint32_t rl, rh;
    asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, %5;"
        : "=r"(rl), "=r"(rh)
        : "r"(s.a), "r"(s.c), "r"(s.b), "r"(s.d));
    const uint32_t tl = s.a ^ s.c, th = s.b ^ s.d;
    // rotl(s0, 55) = rotr(s0, 9)
    const uint32_t ql = __funnelshift_r(s.a, s.b, 9), qh = __funnelshift_r(s.b, s.a, 9);
    // t << 14
    const uint32_t ul = tl << 14, uh = __funnelshift_l(tl, th, 14);
    // three-input xors as single LOP3s (left to itself nvcc re-associates them through
    // s0 ^ s1 to shorten the dependency chain, which costs two more ALU instructions)
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(s.a) : "r"(ql), "r"(tl), "r"(ul));
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(s.b) : "r"(qh), "r"(th), "r"(uh));
    // rotl(t, 36) = swap halves, rotl 4
    s.c = __funnelshift_l(tl, th, 4);
    s.d = __funnelshift_l(th, tl, 4);
    const uint64_t k = ((uint64_t)rh << 32) | (rl & 0xfffff800u);
    return __ull2float_rn(k);
}

__device__ __forceinline__ float rng32_uniform(Rng32 &s) { return rng32_next_scaled(s) * 0x1p-64f; }

// 2 * uniform - 1 in one fma: RN(RN(k) * 2^-52 - 1) == fma(uniform, 2, -1) because both
// scalings are exact
__device__ __forceinline__ float rng32_signed_unit(Rng32 &s) {
    return __fmaf_rn(rng32_next_scaled(s), 0x1p-63f, -1.0f);
}

// numba/cuda/random.py init_xoroshiro128p_state (SplitMix64 of the seed in both words)
inline RngState rng_seed_state(uint64_t seed) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    return RngState{z, z};
}

// numba/cuda/random.py xoroshiro128p_jump (2**64 steps), host side
inline void rng_jump_host(RngState &s) {
    const uint64_t poly[2] = {0xbeac0467eba5facbull, 0xd86b048b86aa9922ull};
    uint64_t a0 = 0, a1 = 0;
    for (int i = 0; i < 2; ++i) {
        for (int b = 0; b < 64; ++b) {
            if (poly[i] & (1ull << b)) {
                a0 ^= s.s0;
                a1 ^= s.s1;
            }
            rng_next(s);
        }
    }
    s.s0 = a0;
    s.s1 = a1;
}

constexpr int kJumpLevels = 48;  // J^(2^k), k < 48: enough for 2^48 states

// 128 columns of 128 bits: column j is the image of basis vector e_j, where bit j of the
// state is bit (j & 63) of s0 (j < 64) or s1 (j >= 64).
struct JumpMatrix {
    RngState col[128];
};

inline RngState jump_apply_host(const JumpMatrix &m, RngState v) {
    RngState r{0, 0};
    for (int j = 0; j < 64; ++j) {
        if ((v.s0 >> j) & 1) { r.s0 ^= m.col[j].s0; r.s1 ^= m.col[j].s1; }
        if ((v.s1 >> j) & 1) { r.s0 ^= m.col[64 + j].s0; r.s1 ^= m.col[64 + j].s1; }
    }
    return r;
}

// levels[k] = J^(2^k)
inline void build_jump_levels(JumpMatrix *levels) {
    for (int j = 0; j < 128; ++j) {
        RngState e{0, 0};
        if (j < 64) e.s0 = 1ull << j; else e.s1 = 1ull << (j - 64);
        rng_jump_host(e);
        levels[0].col[j] = e;
    }
    for (int k = 1; k < kJumpLevels; ++k)
        for (int j = 0; j < 128; ++j)
            levels[k].col[j] = jump_apply_host(levels[k - 1], levels[k - 1].col[j]);
}

// states[dst_first + i] = M * states[i] for i < count. One thread per state; the matrix
// (2 KB) sits in shared memory and every lane reads the same column (broadcast).
__global__ void __launch_bounds__(256)
rng_double_kernel(RngState *__restrict__ states, const JumpMatrix *__restrict__ level,
                  int64_t dst_first, int64_t count) {
    __shared__ uint4 cols[128];
    for (int j = threadIdx.x; j < 128; j += blockDim.x)
        cols[j] = reinterpret_cast<const uint4 *>(level->col)[j];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint4 v = reinterpret_cast<const uint4 *>(states)[i];
    const uint32_t words[4] = {v.x, v.y, v.z, v.w};
    uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
#pragma unroll 8
        for (int b = 0; b < 32; ++b) {
            // all-ones when bit b of the word is set
            const uint32_t mask = (uint32_t)((int32_t)(words[w] << (31 - b)) >> 31);
            const uint4 c = cols[w * 32 + b];
            a0 ^= c.x & mask;
            a1 ^= c.y & mask;
            a2 ^= c.z & mask;
            a3 ^= c.w & mask;
        }
    }
    reinterpret_cast<uint4 *>(states)[dst_first + i] = make_uint4(a0, a1, a2, a3);
}

__global__ void rng_uniform_kernel(RngState *states, int64_t n, int draws, float *out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Rng32 s = rng32_load(states + i);
    for (int k = 0; k < draws; ++k) out[i * draws + k] = rng32_uniform(s);
    rng32_store(states + i, s);
}

}  // namespace rf

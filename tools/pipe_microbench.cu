// Pipe-throughput micro-benchmark for sm_100a: reciprocal throughput (cycles per warp
// instruction per SM sub-partition) of the instruction kinds the tracer is made of, alone
// and mixed, to see which pipes overlap. 1024 threads per SM (8 warps per scheduler), eight
// independent dependency chains per thread. Developer tool; not part of the product.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o pipe_microbench tools/pipe_microbench.cu
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

constexpr int kChains = 8;

// Each body executes exactly `count` instructions of interest per chain and call.
#define BODY_LOP3(x, y) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(k1));
#define BODY_SHF(x, y) asm volatile("shf.l.wrap.b32 %0, %0, %1, 9;" : "+r"(x) : "r"(y));
#define BODY_IADD3(x, y) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
#define BODY_IMAD(x, y) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(k1), "r"(y));
#define BODY_IMADHI(x, y) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(k1), "r"(y));
#define BODY_FFMA(x, y) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x) : "f"(fk), "f"(y));

template <int kTest>
__global__ void __launch_bounds__(1024) bench(uint32_t *out, int iters, uint32_t k1, float fk, long long *cycles) {
    uint32_t a[kChains], b[kChains];
    float f[kChains], g[kChains];
    unsigned long long w[kChains];
#pragma unroll
    for (int i = 0; i < kChains; ++i) {
        a[i] = threadIdx.x * 2654435761u + i;
        b[i] = a[i] ^ 0x9e3779b9u;
        f[i] = (float)(threadIdx.x + i) * 1e-3f;
        g[i] = f[i] + 1.0f;
        w[i] = ((unsigned long long)a[i] << 32) | b[i];
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < kChains; ++i) {
            if (kTest == 0) { BODY_LOP3(a[i], b[i]) BODY_LOP3(a[i], b[i]) }
            if (kTest == 1) { BODY_IMAD(a[i], b[i]) BODY_IMAD(a[i], b[i]) }
            if (kTest == 2) { BODY_LOP3(a[i], b[i]) BODY_IMAD(b[i], a[i]) }
            if (kTest == 3) { BODY_SHF(a[i], b[i]) BODY_SHF(a[i], b[i]) }
            if (kTest == 4) { BODY_FFMA(f[i], g[i]) BODY_FFMA(f[i], g[i]) }
            if (kTest == 5) {  // FFMA2 x2
                asm volatile("{.reg .b64 t, u, v; mov.b64 t, {%0, %1}; mov.b64 u, {%2, %2}; mov.b64 v, {%1, %0};\n\t"
                             "fma.rn.f32x2 t, t, u, v; fma.rn.f32x2 t, t, u, v; mov.b64 {%0, %1}, t;}"
                             : "+f"(f[i]), "+f"(g[i]) : "f"(fk));
            }
            if (kTest == 6) { BODY_LOP3(a[i], b[i]) BODY_FFMA(f[i], g[i]) }
            if (kTest == 7) {  // IMAD.WIDE x2
                asm volatile("mad.wide.u32 %0, %1, %2, %0; mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(k1));
            }
            if (kTest == 8) {  // LOP3 + IMAD.WIDE
                BODY_LOP3(a[i], b[i])
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(b[i]), "r"(k1));
            }
            if (kTest == 9) {  // I2F.U64 x2 (result folded back into the chain)
                float r0, r1;
                asm volatile("cvt.rn.f32.u64 %0, %1;" : "=f"(r0) : "l"(w[i]));
                asm volatile("cvt.rn.f32.u64 %0, %1;" : "=f"(r1) : "l"(w[i] ^ 0x5555ull));
                f[i] += r0 + r1;  // 2 FADD (counted separately below)
            }
            if (kTest == 10) { BODY_LOP3(a[i], b[i]) BODY_LOP3(b[i], a[i]) BODY_IMAD(a[i], b[i]) BODY_FFMA(f[i], g[i]) }
            if (kTest == 11) { BODY_IMADHI(a[i], b[i]) BODY_IMADHI(a[i], b[i]) }
            if (kTest == 12) {  // 2 LOP3 + FFMA2 (3 issue slots, 4 flop-lanes)
                BODY_LOP3(a[i], b[i]) BODY_LOP3(b[i], a[i])
                asm volatile("{.reg .b64 t, u, v; mov.b64 t, {%0, %1}; mov.b64 u, {%2, %2}; mov.b64 v, {%1, %0};\n\t"
                             "fma.rn.f32x2 t, t, u, v; mov.b64 {%0, %1}, t;}"
                             : "+f"(f[i]), "+f"(g[i]) : "f"(fk));
            }
            if (kTest == 13) {  // add.cc / addc pair (IADD3 + IADD3.X or IMAD.X)
                asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(k1), "r"(k1));
            }
            if (kTest == 14) { BODY_LOP3(a[i], b[i]) BODY_SHF(b[i], a[i]) BODY_IMAD(a[i], b[i]) BODY_IMAD(b[i], a[i]) }
            if (kTest == 15) { BODY_LOP3(a[i], b[i]) BODY_LOP3(b[i], a[i]) BODY_LOP3(a[i], b[i]) BODY_IMAD(b[i], a[i]) }
            if (kTest == 16) { BODY_LOP3(a[i], b[i]) BODY_IMADHI(b[i], a[i]) }
            if (kTest == 17) { BODY_LOP3(a[i], b[i]) BODY_LOP3(a[i], b[i]) BODY_IMADHI(b[i], a[i]) BODY_IMAD(b[i], a[i]) }
            if (kTest == 18) { BODY_LOP3(a[i], b[i]) BODY_LOP3(a[i], b[i]) BODY_LOP3(a[i], b[i]) BODY_IMADHI(b[i], a[i]) BODY_IMAD(b[i], a[i]) }
            if (kTest == 19) {  // funnel shift as SHF vs IMAD.HI + IMAD, mixed with 3 LOP3
                BODY_LOP3(a[i], b[i]) BODY_LOP3(a[i], b[i]) BODY_LOP3(a[i], b[i]) BODY_SHF(b[i], a[i])
            }
        }
    }
    const long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int i = 0; i < kChains; ++i) r ^= a[i] ^ b[i] ^ __float_as_uint(f[i]) ^ __float_as_uint(g[i]) ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct Test {
    const char *name;
    int instr_per_chain;  // instructions of interest per chain per iteration
};

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4096;
    uint32_t *d_out;
    long long *d_cycles;
    cudaMalloc(&d_out, sizeof(uint32_t) * sms * 1024);
    cudaMalloc(&d_cycles, sizeof(long long) * sms);
    const Test tests[] = {
        {"LOP3 x2 (alu)", 2},          {"IMAD x2 (fma)", 2},        {"LOP3 + IMAD", 2},
        {"SHF x2 (alu)", 2},           {"FFMA x2", 2},              {"FFMA2 x2", 2},
        {"LOP3 + FFMA", 2},            {"IMAD.WIDE x2", 2},         {"LOP3 + IMAD.WIDE", 2},
        {"I2F.U64 x2 (+2 FADD)", 2},   {"2 LOP3 + IMAD + FFMA", 4}, {"IMAD.HI x2", 2},
        {"2 LOP3 + FFMA2", 3},         {"add.cc + addc", 2},        {"LOP3 + SHF + 2 IMAD", 4},
        {"3 LOP3 + IMAD", 4},          {"LOP3 + IMAD.HI", 2},        {"2 LOP3 + IMAD.HI + IMAD", 4},
        {"3 LOP3 + IMAD.HI + IMAD", 5}, {"3 LOP3 + SHF", 4},
    };
    for (int t = 0; t < (int)(sizeof(tests) / sizeof(tests[0])); ++t) {
        for (int rep = 0; rep < 2; ++rep) {
            switch (t) {
#define CASE(N) case N: bench<N><<<sms, 1024>>>(d_out, iters, 1u + rep * 0, 0.999f, d_cycles); break;
                CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8) CASE(9) CASE(10) CASE(11)
                CASE(12) CASE(13) CASE(14) CASE(15) CASE(16) CASE(17) CASE(18) CASE(19)
            }
            cudaDeviceSynchronize();
        }
        long long c0 = 0;
        cudaMemcpy(&c0, d_cycles, sizeof(c0), cudaMemcpyDeviceToHost);
        const double warp_instr_per_smsp = 8.0 * iters * kChains * tests[t].instr_per_chain;
        printf("%-26s %8.3f cycles / warp-instr / SMSP   (%s)\n", tests[t].name, (double)c0 / warp_instr_per_smsp,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

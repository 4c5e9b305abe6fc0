// Standalone A/B harness of the step-path tracer (rf::trace_mp_kernel): times the kernel on a
// synthetic batch and prints a checksum of the gray frames and of the RNG states after the
// launches, so that build variants (-DRF_... macros of rf_tracer.cuh) can be compared for
// speed and for bit-equality without Python. Developer tool; not part of the product.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo [-DRF_...] \
//        -o trace_ab tools/trace_ab.cu && ./trace_ab [envs=1024] [launches=5] [ctx=7] [H=300] [spp=100]
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../reinfocus_b200/csrc/rf_tracer_mp.cuh"

#ifndef RF_AB_THREADS
#define RF_AB_THREADS 256
#endif
#ifndef RF_AB_MIN_BLOCKS
#define RF_AB_MIN_BLOCKS 4
#endif

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e__ = (call);                                                         \
        if (e__ != cudaSuccess) {                                                         \
            fprintf(stderr, "%s: %s (%s:%d)\n", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            return 1;                                                                     \
        }                                                                                 \
    } while (0)

__global__ void checksum_kernel(const uint32_t *words, int64_t n, unsigned long long *out) {
    unsigned long long acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        acc += (unsigned long long)words[i] * (unsigned long long)(2 * (i % 1000003) + 1);
    for (int o = 16; o; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

template <int K>
static cudaError_t launch(const rf::TraceParams &p, int n, int H, int W) {
    constexpr int T = RF_AB_THREADS;
    auto kernel = rf::trace_mp_kernel<K, T, RF_AB_MIN_BLOCKS>;
    const int per_block = K * T;
    const int blocks_per_env = (H * W + per_block - 1) / per_block;
    const size_t smem = (size_t)per_block * 32;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kernel<<<(unsigned)((int64_t)n * blocks_per_env), T, smem>>>(p);
    return cudaGetLastError();
}

int main(int argc, char **argv) {
    const int n = argc > 1 ? atoi(argv[1]) : 1024;
    const int launches = argc > 2 ? atoi(argv[2]) : 5;
    const int ctx = argc > 3 ? atoi(argv[3]) : 7;
    const int H = argc > 4 ? atoi(argv[4]) : 300;
    const int spp = argc > 5 ? atoi(argv[5]) : 100;
    const int W = H;
    const int64_t total = (int64_t)n * H * W;

    // scene: targets / focus planes on a deterministic spread over [5, 10]; packing as
    // FastWorlds / FastCameras defaults (r_size 20, vfov 30)
    std::vector<float> world(2 * (size_t)n), cam(9 * (size_t)n);
    const float tan10 = (float)std::tan(10.0 * M_PI / 180.0), tan15 = (float)std::tan(15.0 * M_PI / 180.0);
    uint64_t z = 0x9E3779B97F4A7C15ull;
    auto next01 = [&]() {
        z = z * 6364136223846793005ull + 1442695040888963407ull;
        return (float)((z >> 40) * (1.0 / 16777216.0));
    };
    for (int e = 0; e < n; ++e) {
        const float target = 5.0f + 5.0f * next01(), f = 5.0f + 5.0f * next01();
        world[2 * e] = target * tan10;
        world[2 * e + 1] = -target;
        float *c = &cam[9 * (size_t)e];
        c[0] = -(tan15 * f); c[1] = -(tan15 * f); c[2] = -f;
        c[3] = 2.0f * tan15 * f; c[4] = 0; c[5] = 0;
        c[6] = 0; c[7] = 2.0f * tan15 * f; c[8] = 0;
    }
    float *d_world, *d_cam;
    rf::RngState *d_states;
    uint8_t *d_gray;
    unsigned long long *d_sum;
    CK(cudaMalloc(&d_world, world.size() * 4));
    CK(cudaMalloc(&d_cam, cam.size() * 4));
    CK(cudaMalloc(&d_states, sizeof(rf::RngState) * (size_t)total));
    CK(cudaMalloc(&d_gray, (size_t)total));
    CK(cudaMalloc(&d_sum, 16));
    CK(cudaMemcpy(d_world, world.data(), world.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_cam, cam.data(), cam.size() * 4, cudaMemcpyHostToDevice));
    const uint64_t mask = 0x1d9dddffeull;
    CK(cudaMemcpyToSymbol(rf::c_checker_below_mask, &mask, sizeof(mask)));

    // seed-0 states by doubling (as rf_rng_ensure)
    std::vector<rf::JumpMatrix> levels(rf::kJumpLevels);
    rf::build_jump_levels(levels.data());
    rf::JumpMatrix *d_levels;
    CK(cudaMalloc(&d_levels, sizeof(rf::JumpMatrix) * rf::kJumpLevels));
    CK(cudaMemcpy(d_levels, levels.data(), sizeof(rf::JumpMatrix) * rf::kJumpLevels, cudaMemcpyHostToDevice));
    const rf::RngState first = rf::rng_seed_state(0);
    CK(cudaMemcpy(d_states, &first, sizeof(first), cudaMemcpyHostToDevice));
    int level = 0;
    for (int64_t filled = 1; filled < total; filled *= 2, ++level) {
        const int64_t count = std::min(filled, total - filled);
        rf::rng_double_kernel<<<(unsigned)((count + 255) / 256), 256>>>(d_states, d_levels + level, filled, count);
    }
    CK(cudaDeviceSynchronize());

    rf::TraceParams p{};
    p.world = d_world;
    p.cam_dyn = d_cam;
    p.states = d_states;
    p.rgb = nullptr;
    p.gray = d_gray;
    p.origin[0] = p.origin[1] = p.origin[2] = 0.0f;
    p.u[0] = 1; p.v[1] = 1;
    p.lens_radius = 0.05;
    p.scale = (float)(255.0 / (double)spp);
    p.n = n; p.H = H; p.W = W; p.spp = spp;
    p.total = total;

    auto run = [&]() -> cudaError_t {
        switch (ctx) {
            case 2: return launch<2>(p, n, H, W);
            case 3: return launch<3>(p, n, H, W);
            case 4: return launch<4>(p, n, H, W);
            case 6: return launch<6>(p, n, H, W);
            case 7: return launch<7>(p, n, H, W);
            case 8: return launch<8>(p, n, H, W);
            default: return cudaErrorInvalidValue;
        }
    };
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(run());  // warm-up (also launch 1 of the state sequence)
    CK(cudaDeviceSynchronize());
    float best = 1e30f, sum = 0;
    for (int i = 0; i < launches; ++i) {
        CK(cudaEventRecord(e0));
        CK(run());
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
        sum += ms;
    }
    unsigned long long h_sum[2] = {0, 0};
    CK(cudaMemset(d_sum, 0, 16));
    checksum_kernel<<<1024, 256>>>(reinterpret_cast<const uint32_t *>(d_gray), total / 4, d_sum);
    checksum_kernel<<<1024, 256>>>(reinterpret_cast<const uint32_t *>(d_states), total * 4, d_sum + 1);
    CK(cudaMemcpy(h_sum, d_sum, 16, cudaMemcpyDeviceToHost));
    const double rays = (double)total * spp;
    printf("{\"envs\": %d, \"ctx\": %d, \"H\": %d, \"spp\": %d, \"launches\": %d, \"ms_mean\": %.3f, \"ms_best\": %.3f, "
           "\"grays_per_s\": %.2f, \"ms_per_4096_envs\": %.1f, \"gray_sum\": \"%016llx\", \"state_sum\": \"%016llx\"}\n",
           n, ctx, H, spp, launches, sum / launches, best, rays / (best * 1e-3) / 1e9, best * 4096.0 / n, h_sum[0],
           h_sum[1]);
    return 0;
}

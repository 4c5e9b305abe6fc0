// C-ABI of the reinfocus_b200 hot path (include/reinfocus_b200.h). Host-side plumbing
// only: context, buffers, launches. The kernels live in rf_rng.cuh / rf_tracer.cuh /
// rf_focus.cuh / rf_generic.cuh / rf_env.cuh. Built for sm_100a only (see reinfocus_b200/build.py).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/reinfocus_b200.h"
#include "rf_env.cuh"
#include "rf_focus.cuh"
#include "rf_generic.cuh"
#include "rf_rng.cuh"
#include "rf_tracer.cuh"
#include "rf_tracer_mp.cuh"

#define RF_ABI_VERSION 9

namespace {

thread_local std::string g_global_error;

}  // namespace

struct rf_ctx {
    int device = 0;
    cudaDeviceProp prop{};
    std::string error;
    int64_t launches = 0;

    // RNG
    rf::RngState *states = nullptr;
    int64_t n_states = 0;
    rf::JumpMatrix *d_levels = nullptr;  // [kJumpLevels]

    // scene
    float *d_world = nullptr;    // [cap_envs, 2]
    float *d_cam_dyn = nullptr;  // [cap_envs, 9]
    int cap_world = 0, cap_cam = 0;
    int n_world = 0, n_cam = 0;
    float origin[3] = {0, 0, 0}, u[3] = {1, 0, 0}, v[3] = {0, 1, 0};
    double lens_radius = 0.05;
    bool have_world = false, have_cam = false;

    // focus scratch
    unsigned long long *d_accum = nullptr;  // [cap_focus, 2]
    unsigned int *d_tickets = nullptr;      // [cap_focus]
    int cap_focus = 0;

    // step scratch
    uint8_t *d_gray = nullptr;
    int64_t cap_gray = 0;
    double *d_focus = nullptr;
    int cap_focus_out = 0;
    unsigned long long *d_misc = nullptr;  // small scratch (selftests)
    float *d_positions = nullptr;          // [2, cap_positions] targets, focus planes (rf_step_positions_host)
    int cap_positions = 0;

    // general-scene path (rf_render_generic): scene scratch, pristine seed states + working copy
    uint8_t *d_generic_scene = nullptr;
    size_t cap_generic_scene = 0;
    rf::RngState *d_generic_pristine = nullptr, *d_generic_states = nullptr;
    int64_t cap_generic_states = 0;
    uint64_t generic_seed = 0;

    bool force_generic = false;  // RF_OPT_FORCE_GENERIC
    int trace_contexts = -1;     // RF_OPT_TRACE_CONTEXTS: pixels per thread of the default-camera kernel (-1 = by batch size)
    int last_kernel = -1;        // 0 generic, 1 fast (introspection for tests)
    int last_focus_kernel = -1;  // 0 staged (general), 1 packed
};

namespace {

int fail(rf_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (ctx) ctx->error = buf;
    g_global_error = buf;
    return code;
}

#define RF_CUDA(ctx, call)                                                              \
    do {                                                                                \
        cudaError_t err__ = (call);                                                     \
        if (err__ != cudaSuccess)                                                       \
            return fail((ctx), err__ == cudaErrorMemoryAllocation ? RF_ERR_NOMEM : RF_ERR_CUDA, \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), __FILE__, \
                        __LINE__);                                                      \
    } while (0)

#define RF_REQUIRE(ctx, cond, ...)                                    \
    do {                                                              \
        if (!(cond)) return fail((ctx), RF_ERR_INVALID, __VA_ARGS__); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int device) {
        cudaGetDevice(&prev);
        if (prev != device) cudaSetDevice(device);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// bit k set: u = k/32 maps to x = fl64(fl64(32 pi) * k/32) < k*pi, i.e. the boundary value
// still belongs to checker cell k-1 (see rf_tracer.cuh). The 32 comparisons are between a
// float64 and an irrational: the distances are ~1e-16 (k = 29: 1.2e-18, closer than a
// long double can resolve), so the table is computed offline with exact rational
// arithmetic (tests/test_host_side.py::test_checker_boundary_mask recomputes it) and
// checked on the device against the float64 sin for every float32 u by
// rf_selftest_checker. x > k*pi only for k in {13, 17, 21, 25, 26, 29}.
constexpr uint64_t kCheckerBelowMask = 0x1d9dddffeull;

int grow(rf_ctx *ctx, void **ptr, size_t bytes) {
    if (*ptr) {
        RF_CUDA(ctx, cudaFree(*ptr));
        *ptr = nullptr;
    }
    RF_CUDA(ctx, cudaMalloc(ptr, bytes));
    return RF_OK;
}

int ensure_focus_scratch(rf_ctx *ctx, int n, cudaStream_t stream) {
    if (n <= ctx->cap_focus) return RF_OK;
    const int cap = std::max(n, 64);
    ctx->cap_focus = 0;
    if (int rc = grow(ctx, (void **)&ctx->d_accum, sizeof(unsigned long long) * 2 * cap)) return rc;
    if (int rc = grow(ctx, (void **)&ctx->d_tickets, sizeof(unsigned int) * cap)) return rc;
    RF_CUDA(ctx, cudaMemsetAsync(ctx->d_accum, 0, sizeof(unsigned long long) * 2 * cap, stream));
    RF_CUDA(ctx, cudaMemsetAsync(ctx->d_tickets, 0, sizeof(unsigned int) * cap, stream));
    ctx->cap_focus = cap;
    return RF_OK;
}

int rng_init_into(rf_ctx *ctx, rf::RngState *d_states, int64_t n, uint64_t seed,
                  cudaStream_t stream) {
    if (n <= 0) return RF_OK;
    const rf::RngState first = rf::rng_seed_state(seed);
    RF_CUDA(ctx, cudaMemcpyAsync(d_states, &first, sizeof(first), cudaMemcpyHostToDevice, stream));
    // the host buffer is on the stack: make sure the copy has consumed it
    RF_CUDA(ctx, cudaStreamSynchronize(stream));
    int level = 0;
    for (int64_t filled = 1; filled < n; filled *= 2, ++level) {
        if (level >= rf::kJumpLevels) return fail(ctx, RF_ERR_INVALID, "too many RNG states");
        const int64_t count = std::min(filled, n - filled);
        const int64_t blocks = (count + 255) / 256;
        rf::rng_double_kernel<<<(unsigned)blocks, 256, 0, stream>>>(d_states, ctx->d_levels + level,
                                                                   filled, count);
        ctx->launches++;
    }
    RF_CUDA(ctx, cudaGetLastError());
    return RF_OK;
}

int launch_trace(rf_ctx *ctx, int n, int H, int W, int spp, uint8_t *d_rgb, uint8_t *d_gray,
                 cudaStream_t stream) {
    rf::TraceParams p{};
    p.world = ctx->d_world;
    p.cam_dyn = ctx->d_cam_dyn;
    p.states = ctx->states;
    p.rgb = d_rgb;
    p.gray = d_gray;
    for (int i = 0; i < 3; ++i) {
        p.origin[i] = ctx->origin[i];
        p.u[i] = ctx->u[i];
        p.v[i] = ctx->v[i];
    }
    p.lens_radius = ctx->lens_radius;
    p.scale = (float)(255.0 / (double)spp);
    p.n = n;
    p.H = H;
    p.W = W;
    p.spp = spp;
    p.total = (int64_t)n * H * W;
    const int64_t blocks = (p.total + rf::kTraceThreads - 1) / rf::kTraceThreads;
    if (blocks > 0x7fffffffLL) return fail(ctx, RF_ERR_INVALID, "render batch too large");
    // the specialised kernel covers the camera every reference env uses (FastCameras
    // defaults, reference camera.py:99-130); anything else runs the literal statement
    const bool fast = !ctx->force_generic && ctx->u[0] == 1.0f && ctx->u[1] == 0.0f &&
                      ctx->u[2] == 0.0f && ctx->v[0] == 0.0f && ctx->v[1] == 1.0f &&
                      ctx->v[2] == 0.0f && ctx->lens_radius == 0.05;
    // several pixels per thread trade parallelism for fewer divergent rejection-loop trips,
    // and small blocks spread a small grid evenly over the SMs. Measured on 300 x 300 frames
    // (tools/trace_ab.cu over envs x pixels per thread x block size,
    // profiles/r02/latency_small_batches.md): one pixel per thread up to 2 envs, 4 x 64 threads
    // from 3 to 7, 4 x 128 to ~24, 8 x 256 beyond. An explicit RF_OPT_TRACE_CONTEXTS keeps
    // 256-thread blocks.
    int contexts = ctx->trace_contexts;
    int threads = contexts == 8 ? rf::kMpDefaultThreads : 256;
    if (contexts < 0) {
        const int64_t per_sm = p.total / std::max(ctx->prop.multiProcessorCount, 1);
        if (per_sm >= 15000) contexts = rf::kMpDefaultContexts, threads = rf::kMpDefaultThreads;
        else if (per_sm >= 4300) contexts = 4, threads = 128;
        else if (per_sm >= 1500) contexts = 4, threads = 64;
        else contexts = 0;
    }
    if (!fast || H > rf::kMpMaxFrame || W > rf::kMpMaxFrame) contexts = 0;
    if (contexts > 0) {
        // multi-pixel kernel: blocks are per env, kCtx * kThreads pixels each, 32 B of shared
        // memory per pixel (64 KB at 8 x 256: three blocks per SM at 80 registers)
        const int per_block = contexts * threads;
        const int blocks_per_env = (H * W + per_block - 1) / per_block;
        const int64_t grid = (int64_t)n * blocks_per_env;
        if (grid > 0x7fffffffLL) return fail(ctx, RF_ERR_INVALID, "render batch too large");
        const size_t smem = (size_t)per_block * 32;
        auto launch = [&](auto kernel) -> int {
            if (smem > 48 * 1024)
                RF_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kernel<<<(unsigned)grid, threads, smem, stream>>>(p);
            return RF_OK;
        };
        int rc = RF_OK;
        switch (contexts * 1000 + threads) {
            case 2256: rc = launch(rf::trace_mp_kernel<2, 256>); break;
            case 3256: rc = launch(rf::trace_mp_kernel<3, 256>); break;
            case 4256: rc = launch(rf::trace_mp_kernel<4, 256>); break;
            case 5256: rc = launch(rf::trace_mp_kernel<5, 256>); break;
            case 6256: rc = launch(rf::trace_mp_kernel<6, 256>); break;
            case 7256: rc = launch(rf::trace_mp_kernel<7, 256>); break;
            case 8256: rc = launch(rf::trace_mp_kernel<8, 256, rf::kMpDefaultBlocks>); break;
            case 4064: rc = launch(rf::trace_mp_kernel<4, 64>); break;
            case 4128: rc = launch(rf::trace_mp_kernel<4, 128>); break;
            default:
                return fail(ctx, RF_ERR_INVALID, "unsupported context count %d", contexts);
        }
        if (rc) return rc;
    } else if (fast) {
        rf::trace_kernel<true><<<(unsigned)blocks, rf::kTraceThreads, 0, stream>>>(p);
    } else {
        rf::trace_kernel<false><<<(unsigned)blocks, rf::kTraceThreads, 0, stream>>>(p);
    }
    ctx->last_kernel = contexts > 0 ? contexts : (fast ? 1 : 0);
    ctx->launches++;
    RF_CUDA(ctx, cudaGetLastError());
    return RF_OK;
}

int launch_focus_packed(rf_ctx *ctx, int n, int H, int W, const uint8_t *d_img, int channels,
                        double *d_out, cudaStream_t stream) {
    rf::PackedFocusParams p{};
    p.accum = ctx->d_accum;
    p.tickets = ctx->d_tickets;
    p.H = H;
    p.W = W;
    p.channels = channels;
    p.segs = (W + rf::kPackedCols - 1) / rf::kPackedCols;
    // rows per warp tile. Each tile recomputes 4 halo rows, so tall tiles do less work, but
    // the grid should fill the resident warps and come in enough waves that the last one's
    // tail is small. Cost model, checked against measurements from 1 to 4096 envs:
    // (1 + 4 / band) * (w >= 1 ? ceil(w) / w : 1 / w) with w = tiles / resident warps.
    // A tile's sums of the Laplacian and of its square stay in 32 bits until the atomics:
    // 120 columns x 255^2 x rows < 2^32 needs rows <= 550, hence the cap.
    // A last segment of at most 60 columns (W = 300: 120 + 120 + 60) is run for two envs per
    // warp, one per half-warp, so its tiles cost half a warp each.
    const bool shared = W - (p.segs - 1) * rf::kPackedCols <= rf::kPackedHalfCols;
    p.full_segs = shared ? p.segs - 1 : p.segs;
    p.pairs = shared ? (n + 1) / 2 : 0;
    const int64_t warps_per_band = shared ? (int64_t)p.pairs * (2 * p.full_segs + 1) : (int64_t)n * p.segs;
    const double resident = (double)ctx->prop.multiProcessorCount * rf::kPackedWarps * rf::kPackedBlocksPerSM;
    int band = std::min(H, rf::kPackedMaxBand);
    double best = 1e300;
    for (int k = 1; k <= std::max(1, H / 4); ++k) {
        const int rows = (H + k - 1) / k;
        if (rows > rf::kPackedMaxBand) continue;
        const double w = (double)warps_per_band * ((H + rows - 1) / rows) / resident;
        const double fill = w >= 1.0 ? std::ceil(w) / w : 1.0 / w;
        const double cost = (1.0 + 4.0 / rows) * fill;
        if (cost < best - 1e-12) {
            best = cost;
            band = rows;
        }
    }
    p.band = band;
    p.bands = (H + band - 1) / band;
    p.one = 1u;
    p.minus_one = ~0u;
    // warps take tiles in one flat order over all envs, so blocks stay full whatever the
    // number of tiles per env is
    const int64_t tiles = warps_per_band * p.bands;
    const int64_t blocks = (tiles + rf::kPackedWarps - 1) / rf::kPackedWarps;
    if (blocks > 0x7fffffffLL) return fail(ctx, RF_ERR_INVALID, "focus batch too large");
    p.img = d_img;
    p.out = d_out;
    p.n = n;
    if (channels == 1)
        rf::focus_packed_kernel<1><<<(unsigned)blocks, rf::kPackedWarps * 32, 0, stream>>>(p);
    else
        rf::focus_packed_kernel<3><<<(unsigned)blocks, rf::kPackedWarps * 32, 0, stream>>>(p);
    ctx->launches++;
    RF_CUDA(ctx, cudaGetLastError());
    return RF_OK;
}

int launch_focus(rf_ctx *ctx, int n, int H, int W, const uint8_t *d_img, int channels,
                 double *d_out, uint8_t *d_median, uint8_t *d_laplacian, cudaStream_t stream) {
    if (int rc = ensure_focus_scratch(ctx, n, stream)) return rc;
    const bool packed = !ctx->force_generic && !d_median && !d_laplacian && W % 4 == 0 && W >= 8 &&
                        H >= 2 && (reinterpret_cast<uintptr_t>(d_img) & 3) == 0;
    ctx->last_focus_kernel = packed ? 1 : 0;
    if (packed) return launch_focus_packed(ctx, n, H, W, d_img, channels, d_out, stream);
    rf::FocusParams p{};
    p.img = d_img;
    p.out = d_out;
    p.median = d_median;
    p.laplacian = d_laplacian;
    p.accum = ctx->d_accum;
    p.tickets = ctx->d_tickets;
    p.n = n;
    p.H = H;
    p.W = W;
    p.channels = channels;
    p.pitch = (W + 3) & ~3;
    // rows per block: fill the GPU (>= 2 blocks per SM when the batch is small) but keep
    // the +-2 row halo overhead low; bounded by shared memory
    const int sms = ctx->prop.multiProcessorCount;
    int rows = 32;
    while (rows > 4 && (int64_t)n * ((H + rows - 1) / rows) < 2LL * sms) rows /= 2;
    const size_t max_smem = 200 * 1024;
    while (rows > 1 && (size_t)(2 * rows + 6) * p.pitch > max_smem) rows /= 2;
    const size_t smem = (size_t)(2 * rows + 6) * p.pitch;
    if (smem > max_smem) return fail(ctx, RF_ERR_INVALID, "image too wide for the focus kernel");
    p.rows = rows;
    if (smem > 48 * 1024)
        RF_CUDA(ctx, cudaFuncSetAttribute(rf::focus_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)smem));
    const unsigned bands = (unsigned)((H + rows - 1) / rows);
    for (int first = 0; first < n; first += 65535) {
        const int cnt = std::min(65535, n - first);
        rf::FocusParams q = p;
        q.img = d_img + (int64_t)first * H * W * channels;
        q.out = d_out + first;
        q.median = d_median ? d_median + (int64_t)first * H * W : nullptr;
        q.laplacian = d_laplacian ? d_laplacian + (int64_t)first * H * W : nullptr;
        q.accum = ctx->d_accum + 2 * (int64_t)first;
        q.tickets = ctx->d_tickets + first;
        q.n = cnt;
        rf::focus_kernel<<<dim3(bands, cnt), rf::kFocusThreads, smem, stream>>>(q);
        ctx->launches++;
    }
    RF_CUDA(ctx, cudaGetLastError());
    return RF_OK;
}

__global__ void ffma_peak_kernel(float *out, int iters) {
    // 8 independent FFMA chains per thread: enough ILP to saturate both FMA pipes
    float a0 = threadIdx.x * 1e-3f, a1 = a0 + 1.f, a2 = a0 + 2.f, a3 = a0 + 3.f;
    float a4 = a0 + 4.f, a5 = a0 + 5.f, a6 = a0 + 6.f, a7 = a0 + 7.f;
    const float m = 0.999f + blockIdx.x * 1e-9f, c = 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a0 = fmaf(a0, m, c); a1 = fmaf(a1, m, c); a2 = fmaf(a2, m, c); a3 = fmaf(a3, m, c);
            a4 = fmaf(a4, m, c); a5 = fmaf(a5, m, c); a6 = fmaf(a6, m, c); a7 = fmaf(a7, m, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace

extern "C" {

int rf_abi_version(void) { return RF_ABI_VERSION; }

int rf_sizeof(int which) {
    switch (which) {
        case RF_SIZEOF_SCENE_PACKING:
            return (int)sizeof(rf_scene_packing);
        case RF_SIZEOF_ENV_CONFIG:
            return (int)sizeof(rf_env_config);
        default:
            return -1;
    }
}

const char *rf_last_global_error(void) { return g_global_error.c_str(); }

const char *rf_last_error(const rf_ctx *ctx) { return ctx ? ctx->error.c_str() : g_global_error.c_str(); }

int rf_create(rf_ctx **out, int device) {
    if (!out) return fail(nullptr, RF_ERR_INVALID, "rf_create: out is NULL");
    *out = nullptr;
    int count = 0;
    RF_CUDA(nullptr, cudaGetDeviceCount(&count));
    if (device < 0 || device >= count)
        return fail(nullptr, RF_ERR_INVALID, "rf_create: device %d out of range (%d visible)", device, count);
    rf_ctx *ctx = new rf_ctx();
    ctx->device = device;
    DeviceGuard guard(device);
    cudaError_t err = cudaGetDeviceProperties(&ctx->prop, device);
    if (err != cudaSuccess) {
        delete ctx;
        return fail(nullptr, RF_ERR_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(err));
    }
    if (ctx->prop.major != 10) {
        const int major = ctx->prop.major, minor = ctx->prop.minor;
        delete ctx;
        return fail(nullptr, RF_ERR_CUDA,
                    "reinfocus_b200 is built for sm_100a only; device %d is sm_%d%d", device, major, minor);
    }
    std::vector<rf::JumpMatrix> levels(rf::kJumpLevels);
    rf::build_jump_levels(levels.data());
    const uint64_t mask = kCheckerBelowMask;
    auto cleanup = [&](int code) {
        rf_destroy(ctx);
        return code;
    };
    if (cudaMalloc((void **)&ctx->d_levels, sizeof(rf::JumpMatrix) * rf::kJumpLevels) != cudaSuccess)
        return cleanup(fail(nullptr, RF_ERR_NOMEM, "cudaMalloc(jump levels) failed"));
    if (cudaMemcpy(ctx->d_levels, levels.data(), sizeof(rf::JumpMatrix) * rf::kJumpLevels,
                   cudaMemcpyHostToDevice) != cudaSuccess)
        return cleanup(fail(nullptr, RF_ERR_CUDA, "cudaMemcpy(jump levels) failed"));
    if (cudaMemcpyToSymbol(rf::c_checker_below_mask, &mask, sizeof(mask)) != cudaSuccess)
        return cleanup(fail(nullptr, RF_ERR_CUDA, "cudaMemcpyToSymbol(checker mask) failed"));
    if (cudaMalloc((void **)&ctx->d_misc, 64) != cudaSuccess)
        return cleanup(fail(nullptr, RF_ERR_NOMEM, "cudaMalloc(misc) failed"));
    *out = ctx;
    return RF_OK;
}

int rf_destroy(rf_ctx *ctx) {
    if (!ctx) return RF_OK;
    DeviceGuard guard(ctx->device);
    cudaFree(ctx->states);
    cudaFree(ctx->d_levels);
    cudaFree(ctx->d_world);
    cudaFree(ctx->d_cam_dyn);
    cudaFree(ctx->d_accum);
    cudaFree(ctx->d_tickets);
    cudaFree(ctx->d_gray);
    cudaFree(ctx->d_focus);
    cudaFree(ctx->d_misc);
    cudaFree(ctx->d_positions);
    cudaFree(ctx->d_generic_scene);
    cudaFree(ctx->d_generic_pristine);
    cudaFree(ctx->d_generic_states);
    delete ctx;
    return RF_OK;
}

int rf_device_info(const rf_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, int *clock_khz) {
    if (!ctx) return fail(nullptr, RF_ERR_INVALID, "rf_device_info: ctx is NULL");
    if (sm_count) *sm_count = ctx->prop.multiProcessorCount;
    if (cc_major) *cc_major = ctx->prop.major;
    if (cc_minor) *cc_minor = ctx->prop.minor;
    if (clock_khz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);
        *clock_khz = khz;
    }
    return RF_OK;
}

int64_t rf_launch_count(const rf_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------------- RNG

int rf_rng_ensure(rf_ctx *ctx, int64_t n_states, uint64_t seed, void *stream) {
    RF_REQUIRE(ctx, ctx != nullptr, "rf_rng_ensure: ctx is NULL");
    RF_REQUIRE(ctx, n_states >= 0, "rf_rng_ensure: negative n_states");
    if (ctx->states && ctx->n_states >= n_states) return RF_OK;
    DeviceGuard guard(ctx->device);
    if (ctx->states) {
        // frees are stream-agnostic: make sure no kernel still reads the old buffer
        RF_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
        RF_CUDA(ctx, cudaFree(ctx->states));
        ctx->states = nullptr;
        ctx->n_states = 0;
    }
    if (n_states == 0) return RF_OK;
    RF_CUDA(ctx, cudaMalloc((void **)&ctx->states, sizeof(rf::RngState) * (size_t)n_states));
    ctx->n_states = n_states;
    return rng_init_into(ctx, ctx->states, n_states, seed, (cudaStream_t)stream);
}

int rf_rng_reset(rf_ctx *ctx) {
    RF_REQUIRE(ctx, ctx != nullptr, "rf_rng_reset: ctx is NULL");
    DeviceGuard guard(ctx->device);
    if (ctx->states) {
        RF_CUDA(ctx, cudaDeviceSynchronize());
        RF_CUDA(ctx, cudaFree(ctx->states));
    }
    ctx->states = nullptr;
    ctx->n_states = 0;
    return RF_OK;
}

int64_t rf_rng_count(const rf_ctx *ctx) { return ctx ? ctx->n_states : 0; }

int rf_rng_export(rf_ctx *ctx, rf_rng_state *h_dst, int64_t first, int64_t n) {
    RF_REQUIRE(ctx, ctx != nullptr, "rf_rng_export: ctx is NULL");
    RF_REQUIRE(ctx, first >= 0 && n >= 0 && first + n <= ctx->n_states,
               "rf_rng_export: range [%lld, %lld) outside %lld states", (long long)first,
               (long long)(first + n), (long long)ctx->n_states);
    DeviceGuard guard(ctx->device);
    RF_CUDA(ctx, cudaDeviceSynchronize());
    RF_CUDA(ctx, cudaMemcpy(h_dst, ctx->states + first, sizeof(rf::RngState) * (size_t)n,
                            cudaMemcpyDeviceToHost));
    return RF_OK;
}

int rf_rng_import(rf_ctx *ctx, const rf_rng_state *h_src, int64_t first, int64_t n) {
    RF_REQUIRE(ctx, ctx != nullptr, "rf_rng_import: ctx is NULL");
    RF_REQUIRE(ctx, first >= 0 && n >= 0 && first + n <= ctx->n_states,
               "rf_rng_import: range [%lld, %lld) outside %lld states", (long long)first,
               (long long)(first + n), (long long)ctx->n_states);
    DeviceGuard guard(ctx->device);
    RF_CUDA(ctx, cudaDeviceSynchronize());
    RF_CUDA(ctx, cudaMemcpy(ctx->states + first, h_src, sizeof(rf::RngState) * (size_t)n,
                            cudaMemcpyHostToDevice));
    return RF_OK;
}

int rf_rng_init_device(rf_ctx *ctx, rf_rng_state *d_states, int64_t n, uint64_t seed, void *stream) {
    RF_REQUIRE(ctx, ctx != nullptr, "rf_rng_init_device: ctx is NULL");
    RF_REQUIRE(ctx, n >= 0 && (n == 0 || d_states), "rf_rng_init_device: bad buffer");
    DeviceGuard guard(ctx->device);
    return rng_init_into(ctx, reinterpret_cast<rf::RngState *>(d_states), n, seed, (cudaStream_t)stream);
}

int rf_rng_uniform_device(rf_ctx *ctx, rf_rng_state *d_states, int64_t n, int draws, float *d_out,
                          void *stream) {
    RF_REQUIRE(ctx, ctx != nullptr, "rf_rng_uniform_device: ctx is NULL");
    RF_REQUIRE(ctx, n >= 0 && draws >= 0, "rf_rng_uniform_device: negative size");
    if (n == 0 || draws == 0) return RF_OK;
    DeviceGuard guard(ctx->device);
    rf::rng_uniform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<rf::RngState *>(d_states), n, draws, d_out);
    ctx->launches++;
    RF_CUDA(ctx, cudaGetLastError());
    return RF_OK;
}

// -------------------------------------------------------------------------------- scene

int rf_set_world(rf_ctx *ctx, int n, const float *h_world, void *stream) {
    RF_REQUIRE(ctx, ctx != nullptr, "rf_set_world: ctx is NULL");
    RF_REQUIRE(ctx, n > 0 && h_world, "rf_set_world: empty world");
    DeviceGuard guard(ctx->device);
    if (n > ctx->cap_world) {
        RF_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
        ctx->cap_world = 0;  // a failed allocation must not leave a stale capacity behind
        if (int rc = grow(ctx, (void **)&ctx->d_world, sizeof(float) * 2 * (size_t)n)) return rc;
        ctx->cap_world = n;
    }
    RF_CUDA(ctx, cudaMemcpyAsync(ctx->d_world, h_world, sizeof(float) * 2 * (size_t)n,
                                 cudaMemcpyHostToDevice, (cudaStream_t)stream));
    ctx->n_world = n;
    ctx->have_world = true;
    return RF_OK;
}

int rf_set_cameras(rf_ctx *ctx, int n, const float *h_cam_dyn, const float origin[3], const float u[3],
                   const float v[3], double lens_radius, void *stream) {
    RF_REQUIRE(ctx, ctx != nullptr, "rf_set_cameras: ctx is NULL");
    RF_REQUIRE(ctx, n > 0 && h_cam_dyn, "rf_set_cameras: empty cameras");
    RF_REQUIRE(ctx, origin && u && v, "rf_set_cameras: static camera vectors are NULL");
    DeviceGuard guard(ctx->device);
    if (n > ctx->cap_cam) {
        RF_CUDA(ctx, cudaStreamSynchronize((cudaStream_t)stream));
        ctx->cap_cam = 0;
        if (int rc = grow(ctx, (void **)&ctx->d_cam_dyn, sizeof(float) * 9 * (size_t)n)) return rc;
        ctx->cap_cam = n;
    }
    RF_CUDA(ctx, cudaMemcpyAsync(ctx->d_cam_dyn, h_cam_dyn, sizeof(float) * 9 * (size_t)n,
                                 cudaMemcpyHostToDevice, (cudaStream_t)stream));
    for (int i = 0; i < 3; ++i) {
        ctx->origin[i] = origin[i];
        ctx->u[i] = u[i];
        ctx->v[i] = v[i];
    }
    ctx->lens_radius = lens_radius;
    ctx->n_cam = n;
    ctx->have_cam = true;
    return RF_OK;
}

int rf_scene_envs(const rf_ctx *ctx) { return ctx && ctx->have_world ? ctx->n_world : 0; }

// ------------------------------------------------------------------------------- render

static int check_render_args(rf_ctx *ctx, const char *who, int n, int H, int W, int spp) {
    RF_REQUIRE(ctx, ctx != nullptr, "%s: ctx is NULL", who);
    if (!ctx->have_world || !ctx->have_cam)
        return fail(ctx, RF_ERR_NO_SCENE, "%s: targets and focus planes must be set before rendering", who);
    RF_REQUIRE(ctx, n > 0 && H > 0 && W > 0 && spp > 0, "%s: n, H, W, spp must be positive", who);
    RF_REQUIRE(ctx, n <= ctx->n_world, "%s: %d envs requested but the world holds %d", who, n, ctx->n_world);
    RF_REQUIRE(ctx, n <= ctx->n_cam, "%s: %d envs requested but only %d cameras are set", who, n, ctx->n_cam);
    return RF_OK;
}

int rf_render(rf_ctx *ctx, int n, int H, int W, int spp, uint8_t *d_rgb, uint8_t *d_gray, void *stream) {
    if (int rc = check_render_args(ctx, "rf_render", n, H, W, spp)) return rc;
    RF_REQUIRE(ctx, d_rgb || d_gray, "rf_render: no output buffer");
    DeviceGuard guard(ctx->device);
    if (int rc = rf_rng_ensure(ctx, (int64_t)n * H * W, 0, stream)) return rc;
    return launch_trace(ctx, n, H, W, spp, d_rgb, d_gray, (cudaStream_t)stream);
}

int rf_render_generic(rf_ctx *ctx, int n, int H, int W, int spp, int max_shapes,
                      const float *h_shape_params, const int *h_shape_types, const int *h_env_sizes,
                      const double *h_cameras, uint64_t seed, uint8_t *d_rgb, void *stream) {
    RF_REQUIRE(ctx, ctx != nullptr, "rf_render_generic: ctx is NULL");
    RF_REQUIRE(ctx, n > 0 && H > 0 && W > 0 && spp > 0 && max_shapes > 0,
               "rf_render_generic: n, H, W, spp, max_shapes must be positive");
    RF_REQUIRE(ctx, h_shape_params && h_shape_types && h_env_sizes && h_cameras && d_rgb,
               "rf_render_generic: NULL buffer");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t total = (int64_t)n * H * W;
    const size_t bytes_params = sizeof(float) * (size_t)n * max_shapes * rf::kShapeParams;
    const size_t bytes_types = sizeof(int) * (size_t)n * max_shapes;
    const size_t bytes_sizes = sizeof(int) * (size_t)n;
    const size_t bytes_cams = sizeof(double) * (size_t)n * rf::kCameraFields;
    // scratch kept by the context: cameras | params | types | sizes, and two state arrays.
    // The reference seeds fresh states on every call (render.py:115); states of a smaller
    // batch are a prefix of those of a larger one, so one pristine copy per (seed, largest
    // batch so far) is kept and each call starts from a device-to-device copy of its prefix.
    const size_t off_params = bytes_cams;
    const size_t off_types = off_params + bytes_params;
    const size_t off_sizes = off_types + bytes_types;
    const size_t scene_bytes = off_sizes + bytes_sizes;
    if (scene_bytes > ctx->cap_generic_scene) {
        RF_CUDA(ctx, cudaStreamSynchronize(s));
        ctx->cap_generic_scene = 0;
        if (int rc = grow(ctx, (void **)&ctx->d_generic_scene, scene_bytes)) return rc;
        ctx->cap_generic_scene = scene_bytes;
    }
    if (total > ctx->cap_generic_states || seed != ctx->generic_seed) {
        RF_CUDA(ctx, cudaStreamSynchronize(s));
        if (total > ctx->cap_generic_states) {
            ctx->cap_generic_states = 0;
            if (int rc = grow(ctx, (void **)&ctx->d_generic_pristine, sizeof(rf::RngState) * (size_t)total)) return rc;
            if (int rc = grow(ctx, (void **)&ctx->d_generic_states, sizeof(rf::RngState) * (size_t)total)) return rc;
            ctx->cap_generic_states = total;
        }
        if (int rc = rng_init_into(ctx, ctx->d_generic_pristine, ctx->cap_generic_states, seed, s)) return rc;
        ctx->generic_seed = seed;
    }
    uint8_t *scratch = ctx->d_generic_scene;
    RF_CUDA(ctx, cudaMemcpyAsync(scratch, h_cameras, bytes_cams, cudaMemcpyHostToDevice, s));
    RF_CUDA(ctx, cudaMemcpyAsync(scratch + off_params, h_shape_params, bytes_params, cudaMemcpyHostToDevice, s));
    RF_CUDA(ctx, cudaMemcpyAsync(scratch + off_types, h_shape_types, bytes_types, cudaMemcpyHostToDevice, s));
    RF_CUDA(ctx, cudaMemcpyAsync(scratch + off_sizes, h_env_sizes, bytes_sizes, cudaMemcpyHostToDevice, s));
    RF_CUDA(ctx, cudaMemcpyAsync(ctx->d_generic_states, ctx->d_generic_pristine, sizeof(rf::RngState) * (size_t)total,
                                 cudaMemcpyDeviceToDevice, s));
    rf::GenericParams p{};
    p.cameras = reinterpret_cast<const double *>(scratch);
    p.shape_params = reinterpret_cast<const float *>(scratch + off_params);
    p.shape_types = reinterpret_cast<const int *>(scratch + off_types);
    p.env_sizes = reinterpret_cast<const int *>(scratch + off_sizes);
    p.states = ctx->d_generic_states;
    p.rgb = d_rgb;
    p.scale = (float)(255.0 / (double)spp);
    p.n = n;
    p.H = H;
    p.W = W;
    p.spp = spp;
    p.max_shapes = max_shapes;
    p.total = total;
    const int64_t blocks = (total + rf::kTraceThreads - 1) / rf::kTraceThreads;
    if (blocks > 0x7fffffffLL) return fail(ctx, RF_ERR_INVALID, "render batch too large");
    rf::trace_generic_kernel<<<(unsigned)blocks, rf::kTraceThreads, 0, s>>>(p);
    ctx->launches++;
    RF_CUDA(ctx, cudaGetLastError());
    return RF_OK;
}

// -------------------------------------------------------------------------------- focus

int rf_focus_planes(rf_ctx *ctx, int n, int H, int W, const uint8_t *d_img, int channels, double *d_out,
                    uint8_t *d_median, uint8_t *d_laplacian, void *stream) {
    RF_REQUIRE(ctx, ctx != nullptr, "rf_focus: ctx is NULL");
    RF_REQUIRE(ctx, n >= 0 && H > 0 && W > 0, "rf_focus: bad image shape");
    RF_REQUIRE(ctx, channels == 1 || channels == 3, "rf_focus: channels must be 1 or 3");
    if (n == 0) return RF_OK;
    RF_REQUIRE(ctx, d_img && d_out, "rf_focus: NULL buffer");
    DeviceGuard guard(ctx->device);
    return launch_focus(ctx, n, H, W, d_img, channels, d_out, d_median, d_laplacian, (cudaStream_t)stream);
}

int rf_focus(rf_ctx *ctx, int n, int H, int W, const uint8_t *d_img, int channels, double *d_out,
             void *stream) {
    return rf_focus_planes(ctx, n, H, W, d_img, channels, d_out, nullptr, nullptr, stream);
}

// --------------------------------------------------------------------------------- step

static int ensure_step_scratch(rf_ctx *ctx, int n, int H, cudaStream_t stream) {
    const int64_t need = (int64_t)n * H * H;
    if (need > ctx->cap_gray) {
        RF_CUDA(ctx, cudaStreamSynchronize(stream));
        ctx->cap_gray = 0;
        if (int rc = grow(ctx, (void **)&ctx->d_gray, (size_t)need)) return rc;
        ctx->cap_gray = need;
    }
    if (n > ctx->cap_focus_out) {
        RF_CUDA(ctx, cudaStreamSynchronize(stream));
        ctx->cap_focus_out = 0;
        if (int rc = grow(ctx, (void **)&ctx->d_focus, sizeof(double) * (size_t)n)) return rc;
        ctx->cap_focus_out = n;
    }
    return RF_OK;
}

int rf_step_device(rf_ctx *ctx, int n, int H, int spp, double *d_focus, void *stream) {
    if (int rc = check_render_args(ctx, "rf_step_device", n, H, H, spp)) return rc;
    RF_REQUIRE(ctx, d_focus, "rf_step_device: d_focus is NULL");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (int rc = ensure_step_scratch(ctx, n, H, s)) return rc;
    if (int rc = rf_rng_ensure(ctx, (int64_t)n * H * H, 0, stream)) return rc;
    if (int rc = launch_trace(ctx, n, H, H, spp, nullptr, ctx->d_gray, s)) return rc;
    return launch_focus(ctx, n, H, H, ctx->d_gray, 1, d_focus, nullptr, nullptr, s);
}

int rf_step_host(rf_ctx *ctx, int n, int H, int spp, const float *h_world, const float *h_cam_dyn,
                 double *h_focus, void *stream) {
    RF_REQUIRE(ctx, ctx != nullptr, "rf_step_host: ctx is NULL");
    RF_REQUIRE(ctx, h_focus, "rf_step_host: h_focus is NULL");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (h_world)
        if (int rc = rf_set_world(ctx, n, h_world, stream)) return rc;
    if (h_cam_dyn)
        if (int rc = rf_set_cameras(ctx, n, h_cam_dyn, ctx->origin, ctx->u, ctx->v, ctx->lens_radius, stream))
            return rc;
    if (int rc = check_render_args(ctx, "rf_step_host", n, H, H, spp)) return rc;
    if (int rc = ensure_step_scratch(ctx, n, H, s)) return rc;
    if (int rc = rf_step_device(ctx, n, H, spp, ctx->d_focus, stream)) return rc;
    RF_CUDA(ctx, cudaMemcpyAsync(h_focus, ctx->d_focus, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
    RF_CUDA(ctx, cudaStreamSynchronize(s));
    return RF_OK;
}

// ------------------------------------------------------------------- scene from the device

int rf_set_scene_device(rf_ctx *ctx, int n, const float *d_targets, const float *d_planes, int stride,
                        const rf_scene_packing *packing, void *stream) {
    RF_REQUIRE(ctx, ctx != nullptr, "rf_set_scene_device: ctx is NULL");
    RF_REQUIRE(ctx, n > 0 && d_targets && d_planes && stride > 0 && packing,
               "rf_set_scene_device: empty scene");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (n > ctx->cap_world) {
        RF_CUDA(ctx, cudaStreamSynchronize(s));
        ctx->cap_world = 0;
        if (int rc = grow(ctx, (void **)&ctx->d_world, sizeof(float) * 2 * (size_t)n)) return rc;
        ctx->cap_world = n;
    }
    if (n > ctx->cap_cam) {
        RF_CUDA(ctx, cudaStreamSynchronize(s));
        ctx->cap_cam = 0;
        if (int rc = grow(ctx, (void **)&ctx->d_cam_dyn, sizeof(float) * 9 * (size_t)n)) return rc;
        ctx->cap_cam = n;
    }
    rf::ScenePacking k{};
    k.world_tan = packing->world_tan;
    k.half_width = packing->half_width;
    k.half_height = packing->half_height;
    k.full_width = packing->full_width;
    k.full_height = packing->full_height;
    for (int i = 0; i < 3; ++i) {
        k.origin[i] = ctx->origin[i] = packing->origin[i];
        k.u[i] = ctx->u[i] = packing->u[i];
        k.v[i] = ctx->v[i] = packing->v[i];
        k.w[i] = packing->w[i];
    }
    ctx->lens_radius = packing->lens_radius;
    rf::pack_scene_kernel<<<(n + 255) / 256, 256, 0, s>>>(n, d_targets, d_planes, stride, k, ctx->d_world,
                                                          ctx->d_cam_dyn);
    ctx->launches++;
    RF_CUDA(ctx, cudaGetLastError());
    ctx->n_world = ctx->n_cam = n;
    ctx->have_world = ctx->have_cam = true;
    return RF_OK;
}

int rf_step_positions_host(rf_ctx *ctx, int n, int H, int spp, const float *h_targets, const float *h_planes,
                           const rf_scene_packing *packing, double *h_focus, void *stream) {
    RF_REQUIRE(ctx, ctx != nullptr, "rf_step_positions_host: ctx is NULL");
    RF_REQUIRE(ctx, n > 0 && h_targets && h_planes && packing && h_focus, "rf_step_positions_host: empty batch");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (n > ctx->cap_positions) {
        RF_CUDA(ctx, cudaStreamSynchronize(s));
        ctx->cap_positions = 0;
        const int cap = std::max(n, 64);
        if (int rc = grow(ctx, (void **)&ctx->d_positions, sizeof(float) * 2 * (size_t)cap)) return rc;
        ctx->cap_positions = cap;
    }
    float *d_targets = ctx->d_positions, *d_planes = ctx->d_positions + ctx->cap_positions;
    RF_CUDA(ctx, cudaMemcpyAsync(d_targets, h_targets, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, s));
    RF_CUDA(ctx, cudaMemcpyAsync(d_planes, h_planes, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice, s));
    if (int rc = rf_set_scene_device(ctx, n, d_targets, d_planes, 1, packing, stream)) return rc;
    if (int rc = check_render_args(ctx, "rf_step_positions_host", n, H, H, spp)) return rc;
    if (int rc = ensure_step_scratch(ctx, n, H, s)) return rc;
    if (int rc = rf_step_device(ctx, n, H, spp, ctx->d_focus, stream)) return rc;
    RF_CUDA(ctx, cudaMemcpyAsync(h_focus, ctx->d_focus, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, s));
    RF_CUDA(ctx, cudaStreamSynchronize(s));
    return RF_OK;
}

// ----------------------------------------------------------------------- device vector env

struct rf_env {
    rf_ctx *ctx = nullptr;
    rf::EnvParams params{};
    rf_scene_packing packing{};
    int H = 0, spp = 0, node_rows = 0;
    rf::EnvArrays arrays{};
    double *d_focus_main = nullptr, *d_focus_reset = nullptr;
    int *h_counters = nullptr;  // pinned [2]
    cudaEvent_t counted = nullptr;
    bool seeded = false, started = false;
};

int rf_env_destroy(rf_env *env) {
    if (!env) return RF_OK;
    DeviceGuard guard(env->ctx->device);
    rf::EnvArrays &a = env->arrays;
    cudaFree(a.states);
    cudaFree(a.new_states);
    cudaFree(a.reset_rank);
    cudaFree(a.old_obs);
    cudaFree(a.node_state);
    cudaFree(a.generator);
    cudaFree(a.counters);
    cudaFree(env->d_focus_main);
    cudaFree(env->d_focus_reset);
    if (env->h_counters) cudaFreeHost(env->h_counters);
    if (env->counted) cudaEventDestroy(env->counted);
    delete env;
    return RF_OK;
}

int rf_env_create(rf_ctx *ctx, const rf_env_config *c, rf_env **out) {
    RF_REQUIRE(ctx, ctx != nullptr, "rf_env_create: ctx is NULL");
    RF_REQUIRE(ctx, c && out, "rf_env_create: NULL argument");
    RF_REQUIRE(ctx, c->num_envs > 0 && c->frame_height >= 2 && c->samples_per_pixel > 0,
               "rf_env_create: num_envs, frame_height and samples_per_pixel must be positive");
    RF_REQUIRE(ctx, c->transformer >= RF_ENV_DISCRETE_MOVE && c->transformer <= RF_ENV_DISCRETE_JUMP,
               "rf_env_create: unknown transformer %d", c->transformer);
    RF_REQUIRE(ctx, c->n_enders >= 1 && c->n_enders <= rf::kEnvMaxNodes && c->n_rewards >= 1 &&
                        c->n_rewards <= rf::kEnvMaxNodes,
               "rf_env_create: ender / rewarder programs hold 1..%d nodes", rf::kEnvMaxNodes);
    RF_REQUIRE(ctx,
               (c->transformer != RF_ENV_DISCRETE_MOVE && c->transformer != RF_ENV_DISCRETE_JUMP) ||
                   (c->n_moves > 0 && c->n_moves <= rf::kEnvMaxMoves),
               "rf_env_create: a discrete action set holds 1..%d moves", rf::kEnvMaxMoves);
    DeviceGuard guard(ctx->device);
    rf_env *env = new rf_env();
    env->ctx = ctx;
    env->H = c->frame_height;
    env->spp = c->samples_per_pixel;
    env->packing = c->packing;
    rf::EnvParams &p = env->params;
    p.n = c->num_envs;
    p.transformer = c->transformer;
    p.n_moves = c->n_moves;
    for (int i = 0; i < rf::kEnvMaxMoves; ++i) p.moves[i] = i < c->n_moves ? c->moves[i] : 0.0;
    for (int i = 0; i < rf::kEnvMaxMoves; ++i) p.jumps[i] = i < c->n_moves ? c->jumps[i] : 0.0f;
    p.move_speed = c->move_speed;
    p.limit_lo = c->limits[0];
    p.limit_hi = c->limits[1];
    p.jump_span = c->jump_span;
    p.jump_threshold = c->jump_threshold;
    // the observer program: well-formed postfix, one FocusObserver, widths within bounds
    {
        int widths[rf::kEnvMaxObsNodes];
        int top = 0, used = 0, focus_observers = 0, delta_width = 0, normalized = 0;
        bool ok = c->n_observers >= 1 && c->n_observers <= rf::kEnvMaxObsNodes;
        for (int k = 0; ok && k < c->n_observers; ++k) {
            const rf_env_observer &src = c->observers[k];
            rf::ObsNode &node = p.obs_nodes[k];
            node = rf::ObsNode{src.kind, src.arg, src.flag != 0, 0};
            if (src.kind == RF_ENV_OBS_ELEMENT || src.kind == RF_ENV_OBS_FOCUS) {
                ok = src.kind == RF_ENV_OBS_FOCUS || (src.arg >= 0 && src.arg <= 1);
                focus_observers += src.kind == RF_ENV_OBS_FOCUS;
                widths[top++] = 1;
                ok = ok && ++used <= rf::kEnvMaxObsStack;
            } else if (src.kind == RF_ENV_OBS_DELTA || src.kind == RF_ENV_OBS_NORMALIZED) {
                ok = src.arg >= 1 && src.arg <= top;
                int width = 0;
                for (int j = 0; ok && j < src.arg; ++j) width += widths[--top];
                if (!ok) break;
                if (src.kind == RF_ENV_OBS_DELTA) {
                    node.offset = delta_width;
                    delta_width += width;
                    if (node.flag) {
                        used += width;
                        width *= 2;
                    }
                } else {
                    node.offset = src.offset;
                    ok = src.offset >= 0 && src.offset + width <= rf::kEnvMaxObsStack;
                    normalized = std::max(normalized, src.offset + width);
                }
                ok = ok && used <= rf::kEnvMaxObsStack;
                widths[top++] = width;
            } else {
                ok = false;
            }
        }
        if (!ok || top != 1 || focus_observers != 1 || widths[0] > rf::kEnvMaxObsDim) {
            delete env;
            return fail(ctx, RF_ERR_INVALID,
                        "rf_env_create: the observer program must reduce to one vector of at most %d columns "
                        "with exactly one FocusObserver",
                        rf::kEnvMaxObsDim);
        }
        p.n_obs = c->n_observers;
        p.obs_dim = widths[0];
        p.delta_width = delta_width;
        for (int i = 0; i < rf::kEnvMaxObsStack; ++i) {
            p.obs_mid[i] = i < normalized ? c->obs_mid[i] : 0.0f;
            p.obs_scale[i] = i < normalized ? c->obs_scale[i] : 1.0f;
        }
    }
    // the programs: check that they are well-formed postfix expressions over state indices
    // 0 / 1, and give every node its per-env state rows
    int rows = 0, depth = 0;
    p.n_enders = c->n_enders;
    for (int k = 0; k < c->n_enders; ++k) {
        const rf_env_ender &src = c->enders[k];
        rf::EnderNode &node = p.enders[k];
        node = rf::EnderNode{src.kind, src.i0, src.i1, src.steps, src.value, rows};
        const bool leaf = src.kind >= RF_ENV_ENDER_TIME_LIMIT && src.kind <= RF_ENV_ENDER_ENDLESS;
        const bool op = src.kind == RF_ENV_ENDER_AND || src.kind == RF_ENV_ENDER_OR;
        const bool indices_ok = src.i0 >= 0 && src.i0 <= 1 && src.i1 >= 0 && src.i1 <= 1;
        if (!(leaf || op) || !indices_ok || (op && depth < 2) ||
            (src.kind == RF_ENV_ENDER_STOPPED && (src.steps < 1 || src.steps + 1 > rf::kEnvMaxWindow))) {
            delete env;
            return fail(ctx, RF_ERR_INVALID, "rf_env_create: malformed ender program at node %d", k);
        }
        depth += leaf ? 1 : -1;
        if (src.kind == RF_ENV_ENDER_TIME_LIMIT || src.kind == RF_ENV_ENDER_ON_TARGET) rows += 1;
        if (src.kind == RF_ENV_ENDER_DIVERGING) rows += 2;
        if (src.kind == RF_ENV_ENDER_STOPPED) rows += src.steps + 2;
    }
    if (depth != 1) {
        delete env;
        return fail(ctx, RF_ERR_INVALID, "rf_env_create: the ender program does not reduce to one value");
    }
    depth = 0;
    p.n_rewards = c->n_rewards;
    for (int k = 0; k < c->n_rewards; ++k) {
        const rf_env_reward &src = c->rewards[k];
        rf::RewardNode &node = p.rewards[k];
        node = rf::RewardNode{src.kind, src.i0, src.i1, rows, src.f0, src.f1, src.d0, src.d1};
        const bool leaf = src.kind >= RF_ENV_REWARD_DELTA && src.kind <= RF_ENV_REWARD_STOPPED;
        const bool op = src.kind == RF_ENV_REWARD_ADD || src.kind == RF_ENV_REWARD_MUL;
        const int max_index = src.kind == RF_ENV_REWARD_OBSERVATION ? p.obs_dim - 1 : 1;
        const bool indices_ok = src.i0 >= 0 && src.i0 <= max_index && src.i1 >= 0 && src.i1 <= 1;
        if (!(leaf || op) || !indices_ok || (op && depth < 2)) {
            delete env;
            return fail(ctx, RF_ERR_INVALID, "rf_env_create: malformed rewarder program at node %d", k);
        }
        depth += leaf ? 1 : -1;
        if (src.kind == RF_ENV_REWARD_DELTA || src.kind == RF_ENV_REWARD_STOPPED) rows += 1;
    }
    if (depth != 1) {
        delete env;
        return fail(ctx, RF_ERR_INVALID, "rf_env_create: the rewarder program does not reduce to one value");
    }
    env->node_rows = std::max(rows, 1);
    for (int i = 0; i < 2; ++i) {
        if (c->init_options[i] < 1 || c->init_options[i] > rf::kEnvMaxRanges) {
            delete env;
            return fail(ctx, RF_ERR_INVALID, "rf_env_create: an initializer element has 1..%d ranges",
                        rf::kEnvMaxRanges);
        }
        p.init_options[i] = c->init_options[i];
        for (int k = 0; k < rf::kEnvMaxRanges; ++k) {
            p.init_low[i][k] = c->init_low[i][k];
            p.init_range[i][k] = c->init_high[i][k] - c->init_low[i][k];  // Generator.uniform: high - low in float64
        }
    }
    const size_t n = (size_t)p.n;
    rf::EnvArrays &a = env->arrays;
    auto alloc = [&](void **ptr, size_t bytes) {
        if (cudaMalloc(ptr, bytes) != cudaSuccess) return false;
        return cudaMemset(*ptr, 0, bytes) == cudaSuccess;
    };
    const bool ok = alloc((void **)&a.states, sizeof(float) * 2 * n) &&
                    alloc((void **)&a.new_states, sizeof(float) * 2 * n) &&
                    alloc((void **)&a.reset_rank, sizeof(int) * n) &&
                    alloc((void **)&a.old_obs, sizeof(float) * (size_t)std::max(p.delta_width, 1) * n) &&
                    alloc((void **)&a.node_state, sizeof(uint32_t) * n * (size_t)env->node_rows) &&
                    alloc((void **)&a.generator, sizeof(uint64_t) * 6) &&
                    alloc((void **)&a.counters, sizeof(int) * 2) &&
                    alloc((void **)&env->d_focus_main, sizeof(double) * n) &&
                    alloc((void **)&env->d_focus_reset, sizeof(double) * n) &&
                    cudaMallocHost((void **)&env->h_counters, sizeof(int) * 2) == cudaSuccess &&
                    cudaEventCreateWithFlags(&env->counted, cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        const cudaError_t err = cudaGetLastError();
        rf_env_destroy(env);
        return fail(ctx, RF_ERR_NOMEM, "rf_env_create: allocation failed: %s", cudaGetErrorString(err));
    }
    if (const cudaError_t err = cudaDeviceSynchronize(); err != cudaSuccess) {
        rf_env_destroy(env);
        return fail(ctx, RF_ERR_CUDA, "rf_env_create: %s", cudaGetErrorString(err));
    }
    *out = env;
    return RF_OK;
}

int rf_env_set_generator(rf_env *env, const uint64_t state[2], const uint64_t inc[2], uint32_t has_uint32,
                         uint32_t uinteger) {
    if (!env) return fail(nullptr, RF_ERR_INVALID, "rf_env_set_generator: env is NULL");
    rf_ctx *ctx = env->ctx;
    RF_REQUIRE(ctx, state && inc, "rf_env_set_generator: NULL argument");
    DeviceGuard guard(ctx->device);
    const uint64_t words[6] = {state[0], state[1], inc[0], inc[1], has_uint32 ? 1u : 0u, uinteger};
    RF_CUDA(ctx, cudaDeviceSynchronize());
    RF_CUDA(ctx, cudaMemcpy(env->arrays.generator, words, sizeof(words), cudaMemcpyHostToDevice));
    env->seeded = true;
    return RF_OK;
}

int rf_env_get_generator(rf_env *env, uint64_t state[2], uint64_t inc[2], uint32_t *has_uint32,
                         uint32_t *uinteger) {
    if (!env) return fail(nullptr, RF_ERR_INVALID, "rf_env_get_generator: env is NULL");
    rf_ctx *ctx = env->ctx;
    RF_REQUIRE(ctx, state && inc && has_uint32 && uinteger, "rf_env_get_generator: NULL argument");
    DeviceGuard guard(ctx->device);
    uint64_t words[6];
    RF_CUDA(ctx, cudaDeviceSynchronize());
    RF_CUDA(ctx, cudaMemcpy(words, env->arrays.generator, sizeof(words), cudaMemcpyDeviceToHost));
    state[0] = words[0];
    state[1] = words[1];
    inc[0] = words[2];
    inc[1] = words[3];
    *has_uint32 = (uint32_t)words[4];
    *uinteger = (uint32_t)words[5];
    return RF_OK;
}

namespace {

// the part of reset / step shared by both: restart scan, renders, observations
int env_advance(rf_env *env, const void *d_actions, int action_kind, float *d_obs, double *d_rewards,
                uint8_t *d_truncated, int *h_resets, bool reset_all, cudaStream_t s) {
    rf_ctx *ctx = env->ctx;
    const rf::EnvParams &p = env->params;
    rf::env_pre_kernel<<<1, rf::kEnvPreThreads, 0, s>>>(p, env->arrays, d_actions, action_kind, reset_all ? 1 : 0);
    ctx->launches++;
    RF_CUDA(ctx, cudaGetLastError());
    // the restart count is final before any rendering starts: fetch it behind the main
    // render's launch so the host never stalls the GPU
    RF_CUDA(ctx, cudaMemcpyAsync(env->h_counters, env->arrays.counters, sizeof(int) * 2,
                                 cudaMemcpyDeviceToHost, s));
    RF_CUDA(ctx, cudaEventRecord(env->counted, s));
    if (!reset_all) {
        if (int rc = rf_set_scene_device(ctx, p.n, env->arrays.states, env->arrays.states + 1, 2,
                                         &env->packing, s))
            return rc;
        if (int rc = rf_step_device(ctx, p.n, env->H, env->spp, env->d_focus_main, s)) return rc;
    }
    RF_CUDA(ctx, cudaEventSynchronize(env->counted));
    const int resets = env->h_counters[0];
    if (env->h_counters[1]) {
        int zero[2] = {0, 0};
        cudaMemcpyAsync(env->arrays.counters, zero, sizeof(zero), cudaMemcpyHostToDevice, s);
        cudaStreamSynchronize(s);
        env->started = false;
        return fail(ctx, RF_ERR_INVALID,
                    "rf_env_step: an action index is outside the action set (the env must be reset)");
    }
    if (resets > 0) {
        // the restarted envs render as batch positions 0..k-1 (reference
        // state_observer.py:377-383 via vector_environment.py:144)
        if (int rc = rf_set_scene_device(ctx, resets, env->arrays.new_states, env->arrays.new_states + 1, 2,
                                         &env->packing, s))
            return rc;
        if (int rc = rf_step_device(ctx, resets, env->H, env->spp, env->d_focus_reset, s)) return rc;
    }
    rf::env_post_kernel<<<(p.n + 255) / 256, 256, 0, s>>>(p, env->arrays, env->d_focus_main, env->d_focus_reset,
                                                         d_obs, d_rewards, d_truncated, reset_all ? 1 : 0);
    ctx->launches++;
    RF_CUDA(ctx, cudaGetLastError());
    if (h_resets) *h_resets = resets;
    return RF_OK;
}

}  // namespace

int rf_env_reset(rf_env *env, float *d_obs, void *stream) {
    if (!env) return fail(nullptr, RF_ERR_INVALID, "rf_env_reset: env is NULL");
    rf_ctx *ctx = env->ctx;
    RF_REQUIRE(ctx, d_obs, "rf_env_reset: d_obs is NULL");
    RF_REQUIRE(ctx, env->seeded, "rf_env_reset: set the generator first (rf_env_set_generator)");
    DeviceGuard guard(ctx->device);
    if (int rc = env_advance(env, nullptr, 0, d_obs, nullptr, nullptr, nullptr, true, (cudaStream_t)stream))
        return rc;
    env->started = true;
    return RF_OK;
}

int rf_env_step(rf_env *env, const void *d_actions, int action_kind, float *d_obs, double *d_rewards,
                uint8_t *d_truncated, int *h_resets, void *stream) {
    if (!env) return fail(nullptr, RF_ERR_INVALID, "rf_env_step: env is NULL");
    rf_ctx *ctx = env->ctx;
    RF_REQUIRE(ctx, env->started, "rf_env_step: reset the env first");
    RF_REQUIRE(ctx, d_actions && d_obs && d_rewards && d_truncated, "rf_env_step: NULL argument");
    const bool discrete = env->params.transformer == rf::kEnvDiscreteMove ||
                          env->params.transformer == rf::kEnvDiscreteJump;
    RF_REQUIRE(ctx,
               discrete ? (action_kind == RF_ENV_ACTIONS_INT32 || action_kind == RF_ENV_ACTIONS_INT64)
                        : action_kind == RF_ENV_ACTIONS_FLOAT32,
               "rf_env_step: action kind %d does not fit the transformer", action_kind);
    DeviceGuard guard(ctx->device);
    return env_advance(env, d_actions, action_kind, d_obs, d_rewards, d_truncated, h_resets, false,
                       (cudaStream_t)stream);
}

namespace {

// host <-> device copies of the per-env arrays (to_host or from host)
int env_copy_arrays(rf_env *env, bool to_host, float *h_states, float *h_old_obs, uint32_t *h_node_state) {
    rf_ctx *ctx = env->ctx;
    DeviceGuard guard(ctx->device);
    const size_t n = (size_t)env->params.n;
    const rf::EnvArrays &a = env->arrays;
    RF_CUDA(ctx, cudaDeviceSynchronize());
    struct Item {
        void *host;
        void *device;
        size_t bytes;
    } items[] = {
        {h_states, a.states, sizeof(float) * 2 * n},
        {h_old_obs, a.old_obs, sizeof(float) * (size_t)std::max(env->params.delta_width, 1) * n},
        {h_node_state, a.node_state, sizeof(uint32_t) * n * (size_t)env->node_rows},
    };
    for (const Item &item : items) {
        if (!item.host) continue;
        if (to_host)
            RF_CUDA(ctx, cudaMemcpy(item.host, item.device, item.bytes, cudaMemcpyDeviceToHost));
        else
            RF_CUDA(ctx, cudaMemcpy(item.device, item.host, item.bytes, cudaMemcpyHostToDevice));
    }
    return RF_OK;
}

}  // namespace

int rf_env_node_rows(const rf_env *env) { return env ? env->node_rows : 0; }

int rf_env_obs_dim(const rf_env *env) { return env ? env->params.obs_dim : 0; }

int rf_env_delta_width(const rf_env *env) { return env ? std::max(env->params.delta_width, 1) : 0; }

int rf_env_export(rf_env *env, float *h_states, float *h_old_obs, uint32_t *h_node_state) {
    if (!env) return fail(nullptr, RF_ERR_INVALID, "rf_env_export: env is NULL");
    return env_copy_arrays(env, true, h_states, h_old_obs, h_node_state);
}

int rf_env_import(rf_env *env, const float *h_states, const float *h_old_obs, const uint32_t *h_node_state) {
    if (!env) return fail(nullptr, RF_ERR_INVALID, "rf_env_import: env is NULL");
    RF_REQUIRE(env->ctx, h_states && h_old_obs && h_node_state, "rf_env_import: every array is required");
    if (int rc = env_copy_arrays(env, false, const_cast<float *>(h_states), const_cast<float *>(h_old_obs),
                                 const_cast<uint32_t *>(h_node_state)))
        return rc;
    env->started = true;  // the imported episode state stands in for a reset
    return RF_OK;
}

// ---------------------------------------------------------------------------- self-checks

int rf_set_option(rf_ctx *ctx, int option, int value) {
    RF_REQUIRE(ctx, ctx != nullptr, "rf_set_option: ctx is NULL");
    switch (option) {
        case RF_OPT_FORCE_GENERIC:
            ctx->force_generic = value != 0;
            return RF_OK;
        case RF_OPT_TRACE_CONTEXTS:
            if (value != -1 && value != 0 && (value < 2 || value > 8))
                return fail(ctx, RF_ERR_INVALID, "RF_OPT_TRACE_CONTEXTS must be -1, 0 or 2..8");
            ctx->trace_contexts = value;
            return RF_OK;
        default:
            return fail(ctx, RF_ERR_INVALID, "rf_set_option: unknown option %d", option);
    }
}

int rf_get_info(const rf_ctx *ctx, int what) {
    if (!ctx) return -1;
    switch (what) {
        case RF_INFO_LAST_TRACE_KERNEL:
            return ctx->last_kernel;
        case RF_INFO_LAST_FOCUS_KERNEL:
            return ctx->last_focus_kernel;
        default:
            return -1;
    }
}

int rf_selftest(rf_ctx *ctx, int which, int arg, int64_t *mismatches, void *stream) {
    RF_REQUIRE(ctx, ctx != nullptr && mismatches, "rf_selftest: NULL argument");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = (cudaStream_t)stream;
    RF_CUDA(ctx, cudaMemsetAsync(ctx->d_misc, 0, sizeof(unsigned long long), s));
    const int blocks = ctx->prop.multiProcessorCount * 8;
    switch (which) {
        case RF_SELFTEST_CHECKER:
            rf::checker_selftest_kernel<<<blocks, 256, 0, s>>>(ctx->d_misc);
            break;
        case RF_SELFTEST_PIXEL_DIV:
            RF_REQUIRE(ctx, arg > 0 && arg <= 16384, "rf_selftest: frame size out of range");
            rf::pixel_div_selftest_kernel<<<blocks, 256, 0, s>>>(arg, ctx->d_misc);
            break;
        case RF_SELFTEST_INV_LENGTH:
            rf::inv_length_selftest_kernel<<<blocks, 256, 0, s>>>(ctx->d_misc);
            break;
        case RF_SELFTEST_CONST_DIV:
            RF_REQUIRE(ctx, arg > 0 && arg <= 4096, "rf_selftest: divisor count out of range");
            rf::const_div_selftest_kernel<<<blocks, 256, 0, s>>>(arg, ctx->d_misc);
            break;
        case RF_SELFTEST_CHECKER_PAIR:
            rf::checker_pair_selftest_kernel<<<blocks, 256, 0, s>>>(ctx->d_misc);
            break;
        default:
            return fail(ctx, RF_ERR_INVALID, "rf_selftest: unknown test %d", which);
    }
    ctx->launches++;
    RF_CUDA(ctx, cudaGetLastError());
    unsigned long long bad = 0;
    RF_CUDA(ctx, cudaMemcpyAsync(&bad, ctx->d_misc, sizeof(bad), cudaMemcpyDeviceToHost, s));
    RF_CUDA(ctx, cudaStreamSynchronize(s));
    *mismatches = (int64_t)bad;
    return RF_OK;
}

int rf_measure_fp32_peak(rf_ctx *ctx, double *tflops, double *sm_clock_mhz_seen) {
    RF_REQUIRE(ctx, ctx != nullptr && tflops, "rf_measure_fp32_peak: NULL argument");
    DeviceGuard guard(ctx->device);
    const int blocks = ctx->prop.multiProcessorCount * 8, threads = 256, iters = 4096;
    float *d_out = nullptr;
    RF_CUDA(ctx, cudaMalloc((void **)&d_out, sizeof(float) * blocks * threads));
    cudaEvent_t e0, e1;
    RF_CUDA(ctx, cudaEventCreate(&e0));
    RF_CUDA(ctx, cudaEventCreate(&e1));
    double best_ms = 1e30;
    for (int rep = 0; rep < 6; ++rep) {
        RF_CUDA(ctx, cudaEventRecord(e0, 0));
        ffma_peak_kernel<<<blocks, threads>>>(d_out, iters);
        ctx->launches++;
        RF_CUDA(ctx, cudaEventRecord(e1, 0));
        RF_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        RF_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best_ms = std::min(best_ms, (double)ms);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    const double flop = 2.0 * 64.0 * iters * (double)blocks * threads;
    *tflops = flop / (best_ms * 1e-3) / 1e12;
    if (sm_clock_mhz_seen) {
        // FFMA issue rate -> implied clock: 128 lanes/SM/clk
        *sm_clock_mhz_seen = (*tflops * 1e12 / 2.0) / (128.0 * ctx->prop.multiProcessorCount) / 1e6;
    }
    return RF_OK;
}

}  // extern "C"

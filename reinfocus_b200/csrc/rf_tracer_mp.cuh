// Step-path tracer: trace_mp_kernel<kCtx, kThreads>, several pixels per thread.
//
// Same contract and the same bytes as trace_kernel<true> (rf_tracer.cuh): the default-camera
// statement of FastRenderer._device_render (reference graphics/render.py:190-246 and callees),
// every pixel consuming its own xoroshiro128+ stream in the reference's order.
//
// What bounds the tracer on sm_100a is the ALU pipe under xoroshiro128+ (11 LOP3 / SHF / IADD3
// per draw, one warp instruction per 2 clocks), and a warp executes about twice the draws its
// lanes need: the two rejection loops run until the slowest of 32 lanes accepts (disc: 3.1
// warp iterations for 1.27 per lane; sphere: 6.0 for 1.9). The streams of different pixels
// are independent, so a lane that owns kCtx pixels ("contexts") spends the iterations it would
// have idled through on its next pixel: the warp then iterates max-over-lanes of a SUM of kCtx
// geometric counts, whose mean per pixel drops (disc 3.1 -> 1.8, sphere 6.0 -> 3.1 at kCtx = 8).
//
// Each thread owns kCtx pixels of one env (pixel = block base + c*kThreads + tid). Per sample
// the work is split into phases that every lane runs in lock step:
//   J  per context: two jitter draws -> (s, t) in registers           straight line
//   D  disc rejection over the lane's contexts, one after another     shared loop
//   H  per pair of contexts: ray + hit test + checker parity          straight line, packed
//   S  sphere rejection over the lane's contexts that hit             shared loop
//   C  per pair of contexts: shade, accumulate                        straight line, packed
// The context a loop is working on changes per lane, so per-context data lives in shared
// memory ([slot][ctx][thread], conflict-free 128-bit accesses), addressed by a per-lane
// running pointer; only the active RNG state sits in registers.
//
// How the instructions are spent (round 1's kernel, same structure: 315 warp instructions per
// 32 pixel-samples, 147 on the ALU pipe, 31 BSSY / BSYNC / BRA, 8 PLOP3; this one: 261 / 133 /
// 11 / 0):
//
//   * the rejection loops carry one branch each. The accept path (store the sample and the
//     RNG state, step to the lane's next pixel, fetch its state) is predicated inline PTX
//     on a running shared-memory address instead of a divergent region with its own
//     reconvergence barrier; the sphere loop walks a nibble list of the pixels that hit
//     (built with one IMAD per hit) instead of ffs / clear-lowest-bit on a mask;
//   * the straight-line phases run on pairs of pixels with sm_100's packed FP32
//     (FFMA2 via __ffma2_rn / __fmul2_rn / __fadd2_rn): half the issue slots for the same
//     round-to-nearest results per lane. ptxas contracts a packed multiply that feeds a
//     packed add into one FFMA2 (it does not for scalar mul.rn / add.rn), so every packed
//     product that is later added is either exact (a power-of-two factor) or kept scalar;
//   * the checkerboard parity comes from RZ / RU fused multiply-adds on the FMA pipe
//     (2^23 + floor(32 u) in the mantissa) instead of F2I / I2F on the XU pipe; exact cell
//     boundaries still go through the table (checker_is_red);
//   * the accumulation is a predicated add: fma(att, sky, acc) with att in {0, 1} is acc
//     or RN(sky + acc).
#pragma once

#include "rf_tracer.cuh"

namespace rf {

constexpr int kMpMaxFrame = 2048;  // half-precision pixel coordinates
// pixels per thread, threads per block and blocks per SM of large batches: 8 x 256 x 3 (64 KB of
// shared memory per block, 80 registers, 24 warps per SM; 2048 pixels per block leave the last
// of a 300 x 300 env's 44 blocks 95 % full): 341.9 ms per 4096-env launch; 8 x 192 x 4: 346.0;
// 8 x 224 x 4 (72 registers): 350.4; 7 x 256 x 4 (64 registers): 353.4
constexpr int kMpDefaultContexts = 8;
constexpr int kMpDefaultThreads = 256;
constexpr int kMpDefaultBlocks = 3;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ Rng32 lds_state(uint32_t addr) {
    Rng32 s;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(s.a), "=r"(s.b), "=r"(s.c), "=r"(s.d) : "r"(addr));
    return s;
}

__device__ __forceinline__ void sts_state(uint32_t addr, const Rng32 &s) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(s.a), "r"(s.b), "r"(s.c), "r"(s.d) : "memory");
}

// 32-bit load that ptxas may not fuse with its neighbours: a fused 64 / 128-bit load lands in
// consecutive registers, and the packed code wants (pixel 0, pixel 1) pairs, so the fusion
// would cost two moves per value
__device__ __forceinline__ float lds_f32(uint32_t addr) {
    float v;
    asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }

// Parity of the checkerboard cells of two pixels' (u, v) in [0, 1]^2 without leaving the FMA
// pipe: RZ(32 u + 2^23) = 2^23 + floor(32 u) exactly (32 u < 2^23; the fma rounds the exact
// sum once, towards zero), so the low mantissa bit is the cell parity, and RU of the same
// sum differs from it iff 32 u is not an integer. green: bit 0 of the result per pixel
// (cells of different parity); exact: neither coordinate sits on a cell boundary (else the
// caller asks the table, checker_is_red).
struct CheckerPair {
    uint32_t green0, green1;
    bool exact0, exact1;
};

__device__ __forceinline__ CheckerPair checker_pair(float2 u, float2 v) {
    const float2 du = __ffma2_rz(u, f2(32.0f), f2(8388608.0f)), dv = __ffma2_rz(v, f2(32.0f), f2(8388608.0f));
    const float2 uu = __ffma2_ru(u, f2(32.0f), f2(8388608.0f)), uv = __ffma2_ru(v, f2(32.0f), f2(8388608.0f));
    CheckerPair r;
    r.green0 = (__float_as_uint(du.x) ^ __float_as_uint(dv.x)) & 1u;
    r.green1 = (__float_as_uint(du.y) ^ __float_as_uint(dv.y)) & 1u;
    r.exact0 = du.x != uu.x && dv.x != uv.x;
    r.exact1 = du.y != uu.y && dv.y != uv.y;
    return r;
}

// checker_pair() against the table-based cell (checker_cell, itself checked against the
// float64 sine by checker_selftest_kernel) for every float32 coordinate in [0, 1], each paired
// with itself, with 0.3 and with 1: the parity where checker_pair claims exactness, and the
// claim itself (exact iff 32 u is not an integer)
__global__ void checker_pair_selftest_kernel(unsigned long long *mismatches) {
    const uint32_t one_bits = 0x3f800000u;
    unsigned long long bad = 0;
    for (uint64_t bits = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; bits <= one_bits;
         bits += (uint64_t)gridDim.x * blockDim.x) {
        const float u = __uint_as_float((uint32_t)bits);
        const float t = u * 32.0f;
        const bool integral = __int2float_rn(__float2int_rd(t)) == t;
        const float others[3] = {u, 0.3f, 1.0f};
        for (int k = 0; k < 3; ++k) {
            const float v = others[k];
            const float tv = v * 32.0f;
            const bool v_integral = __int2float_rn(__float2int_rd(tv)) == tv;
            const CheckerPair ck = checker_pair(f2(u, v), f2(v, u));
            bad += ck.exact0 != (!integral && !v_integral);
            bad += ck.exact1 != (!integral && !v_integral);
            if (!integral && !v_integral) {
                const uint32_t green = (uint32_t)((checker_cell(u) ^ checker_cell(v)) & 1);
                bad += ck.green0 != green;
                bad += ck.green1 != green;
            }
        }
    }
    if (bad) atomicAdd(mismatches, bad);
}

// acc = RN(acc + x) unless `skip` is non-zero, in place (a select would cost a move per
// accumulator at the end of every sample)
__device__ __forceinline__ void add_unless(float &acc, float x, uint32_t skip) {
    asm("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, %2, 0;\n\t@p add.rn.f32 %0, %0, %1;\n\t}" : "+f"(acc) : "f"(x), "r"(skip));
}

// kMinBlocks: resident blocks per SM the register allocation must allow (4: at most 64
// registers at 256 threads; 3: 85)
template <int kCtx, int kThreads, int kMinBlocks = 4>
__global__ void __launch_bounds__(kThreads, kMinBlocks) trace_mp_kernel(const TraceParams p) {
    static_assert(kCtx >= 2 && kCtx <= 8, "2..8 pixels per thread");
    constexpr int kPairs = (kCtx + 1) / 2;
    constexpr uint32_t kStride = kThreads * 16;      // bytes between contexts of one thread
    constexpr uint32_t kWork = kCtx * kStride;       // work slot = state slot + kWork
    extern __shared__ __align__(16) uint8_t mp_smem[];
    // [kCtx][T] RNG states (uint4), then [kCtx][T] work slots (float4: x, y, z = the accepted
    // disc sample from D to H, then the accepted sphere sample from S to C; w = the pixel
    // coordinates as a half2)

    const int tid = threadIdx.x;
    const int hw = p.H * p.W;
    // Block order: every env's full chunks first, then the envs' last, partly filled chunks.
    // A partly filled block has fewer pixels per thread and ends sooner (at 300 x 300 its
    // threads own one or two pixels instead of eight), so the launch ends on short blocks:
    // the idle tail while the last blocks drain is what 8-GPU scaling loses at 512 envs per
    // GPU (44 waves of ~1 ms blocks).
    const int full_per_env = hw / (kCtx * kThreads);
    int e, chunk;
    if ((int64_t)blockIdx.x < (int64_t)p.n * full_per_env) {
        e = blockIdx.x / full_per_env;
        chunk = blockIdx.x - e * full_per_env;
    } else {
        e = (int)((int64_t)blockIdx.x - (int64_t)p.n * full_per_env);
        chunk = full_per_env;
    }
    const int first = chunk * (kCtx * kThreads);  // first pixel of this block within the env
    uint32_t sbase;  // state slot of context 0 (opaque to the compiler: it would otherwise
                     // recompute the address from the thread index at every use)
    asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(smem_u32(mp_smem) + tid * 16));

    // env constants (block uniform)
    const float *cam = p.cam_dyn + (int64_t)e * 9;
    const float llx = __fadd_rn(__ldg(cam + 0), 0.0f);
    const float lly = __fadd_rn(__ldg(cam + 1), 0.0f);
    const float llz = __fadd_rn(__ldg(cam + 2), 0.0f);
    const float hzx = __ldg(cam + 3), vty = __ldg(cam + 7);
    const float orgx = __fadd_rn(p.origin[0], 0.0f);
    const float orgy = __fadd_rn(p.origin[1], 0.0f);
    const float orgz = __fadd_rn(p.origin[2], 0.0f);
    const float radius = __ldg(p.world + 2 * (int64_t)e);
    const float zpos = __ldg(p.world + 2 * (int64_t)e + 1);
    const float dz = __fsub_rn(llz, orgz);
    const float th = __fdiv_rn(__fsub_rn(zpos, orgz), dz);
    const bool th_valid = !(th < 0.001f || th > 1000000.0f);
    const float two_r = __fadd_rn(radius, radius);
    const float two_r_rcp = division_reciprocal(two_r);
    const double Wd = (double)p.W, Hd = (double)p.H;
    const double Wrcp = refined_reciprocal(Wd), Hrcp = refined_reciprocal(Hd);
    const float lens_hi = 0x1.99999ap-5f, lens_lo = -0x1.99999ap-31f;

    // contexts: this thread's pixels first + c*T + tid; they form a prefix (nctx of them)
    int nctx = 0;
    float2 accx[kPairs], accy[kPairs], accz[kPairs];  // [pair].x / .y: contexts 2q / 2q + 1
#pragma unroll
    for (int q = 0; q < kPairs; ++q) accx[q] = accy[q] = accz[q] = f2(0.0f);
#pragma unroll
    for (int c = 0; c < kCtx; ++c) {
        const int pix = first + c * kThreads + tid;
        if (pix < hw) {
            nctx = c + 1;
            const int y = pix / p.W, x = pix - y * p.W;
            const __half2 xy = __floats2half2_rn((float)x, (float)y);
            uint4 v = *reinterpret_cast<const uint4 *>(p.states + (int64_t)e * hw + pix);
            sts_state(sbase + c * kStride, Rng32{v.x, v.y, v.z, v.w});
            asm volatile("st.shared.b32 [%0], %1;" ::"r"(sbase + kWork + c * kStride + 12),
                         "r"(*reinterpret_cast<const uint32_t *>(&xy))
                         : "memory");
        }
    }

    // one sample of every pixel of this thread. kFull: all kCtx pixels exist (every block but
    // the last of an env), which strips the per-context guards from the straight-line phases
    auto sample_all = [&](auto full_tag) {
        constexpr bool kFull = decltype(full_tag)::value;
        const int limit = kFull ? kCtx : nctx;
        float2 va[kPairs], vb[kPairs];  // J: (s0, s1), (t0, t1); after H: direction x, y (or uv)
        // ---- J: jitter -------------------------------------------------------------------
#pragma unroll
        for (int c = 0; c < kCtx; ++c) {
            float s = 0.0f, t = 0.0f;
            if (kFull || c < nctx) {
                const uint32_t a = sbase + c * kStride;
                Rng32 st = lds_state(a);
                uint32_t xy_bits;
                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(xy_bits) : "r"(a + kWork + 12));
                // half -> double in one conversion each (F2F.F64.F16 on either half of the word)
                double xd, yd;
                asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tcvt.f64.f16 %0, lo;\n\tcvt.f64.f16 %1, hi;\n\t}"
                    : "=d"(xd), "=d"(yd)
                    : "r"(xy_bits));
                s = pixel_coordinate(xd, rng32_next_scaled(st), Wd, Wrcp);
                t = pixel_coordinate(yd, rng32_next_scaled(st), Hd, Hrcp);
                sts_state(a, st);
            }
            if (c & 1) { va[c >> 1].y = s; vb[c >> 1].y = t; } else { va[c >> 1].x = s; vb[c >> 1].x = t; }
        }
        if (kCtx & 1) { va[kPairs - 1].y = 0.0f; vb[kPairs - 1].y = 0.0f; }
        // ---- D: disc rejection, the lane's pixels one after another ---------------------------
        {
            uint32_t a = sbase;
            const uint32_t end = sbase + (uint32_t)limit * kStride;
            if (a != end) {
                Rng32 st = lds_state(a);
                do {
                    const float px = rng32_signed_unit(st);
                    const float py = rng32_signed_unit(st);
                    const float len2 = __fmaf_rn(px, px, __fmul_rn(py, py));
                    // accept (len2 < 1): sample and state to the pixel's slots, on to the next
                    // pixel (past the last one the load fetches bytes nobody uses)
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "setp.lt.f32 p, %7, 0f3F800000;\n\t"
                        "@p st.shared.v2.f32 [%0+%8], {%5, %6};\n\t"
                        "@p st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n\t"
                        "@p add.u32 %0, %0, %9;\n\t"
                        "@p ld.shared.v4.b32 {%1, %2, %3, %4}, [%0];\n\t}"
                        : "+r"(a), "+r"(st.a), "+r"(st.b), "+r"(st.c), "+r"(st.d)
                        : "f"(px), "f"(py), "f"(len2), "n"(kWork), "n"(kStride)
                        : "memory");
                } while (a != end);
            }
        }
        // ---- H: ray + hit test, two pixels per instruction -----------------------------------
        uint32_t hits = 0, greens = 0;
        uint32_t list = 0;  // nibbles kCtx - c of the contexts that hit, last one first; 0 ends it
#pragma unroll
        for (int q = 0; q < kPairs; ++q) {
            const int c0 = 2 * q, c1 = (2 * q + 1 < kCtx) ? 2 * q + 1 : 2 * q;
            const bool on0 = kFull || c0 < nctx, on1 = (2 * q + 1 < kCtx) && (kFull || c1 < nctx);
            if (on0) {
                const uint32_t w0 = sbase + kWork + c0 * kStride, w1 = sbase + kWork + c1 * kStride;
                float2 dpx, dpy;  // disc samples (x0, x1), (y0, y1)
                dpx.x = lds_f32(w0); dpy.x = lds_f32(w0 + 4);
                dpx.y = on1 ? lds_f32(w1) : 0.0f; dpy.y = on1 ? lds_f32(w1 + 4) : 0.0f;
                const float2 ox = __fadd2_rn(f2(orgx), __ffma2_rn(dpx, f2(lens_hi), __fmul2_rn(dpx, f2(lens_lo))));
                const float2 oy = __fadd2_rn(f2(orgy), __ffma2_rn(dpy, f2(lens_hi), __fmul2_rn(dpy, f2(lens_lo))));
                // direction = fma(h, s, ll) - o; a - b == fma(b, -1, a)
                float2 dx = __ffma2_rn(ox, f2(-1.0f), __ffma2_rn(f2(hzx), va[q], f2(llx)));
                float2 dy = __ffma2_rn(oy, f2(-1.0f), __ffma2_rn(f2(vty), vb[q], f2(lly)));
                if (th_valid) {
                    const float2 Px = __ffma2_rn(dx, f2(th), ox);  // (o + 0) + d*t, see plus_zero
                    const float2 Py = __ffma2_rn(dy, f2(th), oy);
                    const bool hit0 = !(fabsf(Px.x) > radius || fabsf(Py.x) > radius);
                    const bool hit1 = on1 && !(fabsf(Px.y) > radius || fabsf(Py.y) > radius);
                    if (hit0 || hit1) {
                        // rectangle.uv: (r + P) / (r + r) by the hoisted reciprocal
                        const float2 nx = __fadd2_rn(f2(radius), Px), ny = __fadd2_rn(f2(radius), Py);
                        const float2 qx = __fmul2_rn(nx, f2(two_r_rcp)), qy = __fmul2_rn(ny, f2(two_r_rcp));
                        const float2 rx = __ffma2_rn(qx, f2(-two_r), nx), ry = __ffma2_rn(qy, f2(-two_r), ny);
                        const float2 u = __ffma2_rn(f2(two_r_rcp), rx, qx), v = __ffma2_rn(f2(two_r_rcp), ry, qy);
                        CheckerPair ck = checker_pair(u, v);
                        if (hit0) {
                            if (!ck.exact0) ck.green0 = checker_is_red(u.x, v.x) ? 0u : 1u;
                            hits += 1u << c0;
                            greens += ck.green0 << c0;
                            list = list * 16u + (kCtx - c0);
                        }
                        if (hit1) {
                            if (!ck.exact1) ck.green1 = checker_is_red(u.y, v.y) ? 0u : 1u;
                            hits += 1u << c1;
                            greens += ck.green1 << c1;
                            list = list * 16u + (kCtx - c1);
                        }
                    }
                }
                va[q] = dx;
                vb[q] = dy;
            }
        }
        // ---- S: sphere rejection over the pixels that hit ------------------------------------
        {
            uint32_t cur = list & 15u;
            if (cur != 0) {
                // slot of context kCtx - cur; cur == 0 lands on the first work slot, which makes
                // the fetch after the last accept harmless
                uint32_t a = sbase + kWork - cur * kStride;
                Rng32 st = lds_state(a);
                do {
                    const float qx = rng32_signed_unit(st);
                    const float qy = rng32_signed_unit(st);
                    const float qz = rng32_signed_unit(st);
                    const float len2 = __fmaf_rn(qz, qz, __fmaf_rn(qx, qx, __fmul_rn(qy, qy)));
                    asm volatile(
                        "{\n\t.reg .pred p;\n\t"
                        "setp.lt.f32 p, %10, 0f3F800000;\n\t"
                        "@p st.shared.v2.f32 [%0+%11], {%7, %8};\n\t"
                        "@p st.shared.f32 [%0+%12], %9;\n\t"
                        "@p st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n\t"
                        "@p shr.u32 %5, %5, 4;\n\t"
                        "@p and.b32 %6, %5, 15;\n\t"
                        "@p mad.lo.u32 %0, %6, %13, %14;\n\t"
                        "@p ld.shared.v4.b32 {%1, %2, %3, %4}, [%0];\n\t}"
                        : "+r"(a), "+r"(st.a), "+r"(st.b), "+r"(st.c), "+r"(st.d), "+r"(list), "+r"(cur)
                        : "f"(qx), "f"(qy), "f"(qz), "f"(len2), "n"(kWork), "n"(kWork + 8), "r"(0u - kStride), "r"(sbase + kWork)
                        : "memory");
                } while (cur != 0);
            }
        }
        // ---- C: shade + accumulate, two pixels per instruction --------------------------------
        // sky colour of a pair of directions (vector.py:354-364, physics.py:183-193)
        auto sky_of = [&](float2 rx, float2 ry, float2 rz, float2 &sx, float2 &sy, float2 &sz) {
            const float2 l2 = __ffma2_rn(rz, rz, __ffma2_rn(rx, rx, __fmul2_rn(ry, ry)));
            // inverse_length() on both halves
            float2 r, y;
            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(l2.x));
            asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(l2.y));
            const float2 s0 = __fmul2_rn(l2, r);
            const float2 h = __fmul2_rn(r, f2(0.5f));
            const float2 ee = __ffma2_rn(f2(-s0.x, -s0.y), s0, l2);
            const float2 s = __ffma2_rn(ee, h, s0);
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y.x) : "f"(s.x));
            asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y.y) : "f"(s.y));
            const float2 dd = __ffma2_rn(f2(-s.x, -s.y), y, f2(1.0f));
            const float2 inv = __ffma2_rn(y, dd, y);
            // unit.y: a product that is added to below, so it stays scalar (see the header)
            const float2 ny = f2(__fmul_rn(ry.x, inv.x), __fmul_rn(ry.y, inv.y));
            // add_sky(): the halvings / quarterings are exact, so a contraction of them into
            // the sums leaves the bits alone
            const float2 up = __fadd2_rn(ny, f2(1.0f));
            const float2 base = __fmul2_rn(f2(0.5f), __ffma2_rn(ny, f2(-1.0f), f2(1.0f)));
            sx = __fadd2_rn(base, __fmul2_rn(f2(0.25f), up));
            sy = __fadd2_rn(base, __ffma2_rn(ny, f2(0.7f * 0.5f), f2(0.7f * 0.5f)));
            sz = __fadd2_rn(base, __fmul2_rn(f2(0.5f), up));
        };
        // attenuation is (1,1,1) for a miss, (1,0,0) on a red cell, (0,1,0) on a green one, and
        // fma(att, sky, acc) with att in {0, 1} is RN(sky + acc) or acc
        if (kFull && hits == 0) {
            // none of this thread's pixels hit (whole warps, away from the target's edge)
#pragma unroll
            for (int q = 0; q < kPairs; ++q) {
                float2 sx, sy, sz;
                sky_of(va[q], vb[q], f2(dz), sx, sy, sz);
                accx[q] = __fadd2_rn(sx, accx[q]);
                accy[q] = __fadd2_rn(sy, accy[q]);
                accz[q] = __fadd2_rn(sz, accz[q]);
            }
        } else {
            const uint32_t no_x = hits & greens, no_y = hits & ~greens;  // channels left alone
#pragma unroll
            for (int q = 0; q < kPairs; ++q) {
                const int c0 = 2 * q, c1 = (2 * q + 1 < kCtx) ? 2 * q + 1 : 2 * q;
                const bool on0 = kFull || c0 < nctx, on1 = (2 * q + 1 < kCtx) && (kFull || c1 < nctx);
                if (on0) {
                    float2 rx = va[q], ry = vb[q], rz = f2(dz);
                    if (hits & (1u << c0)) {
                        const uint32_t w0 = sbase + kWork + c0 * kStride;
                        rx.x = lds_f32(w0);  // (0 + 0) + q, see plus_zero
                        ry.x = lds_f32(w0 + 4);
                        rz.x = __fadd_rn(1.0f, lds_f32(w0 + 8));
                    }
                    if (on1 && (hits & (1u << c1))) {
                        const uint32_t w1 = sbase + kWork + c1 * kStride;
                        rx.y = lds_f32(w1);
                        ry.y = lds_f32(w1 + 4);
                        rz.y = __fadd_rn(1.0f, lds_f32(w1 + 8));
                    }
                    float2 sx, sy, sz;
                    sky_of(rx, ry, rz, sx, sy, sz);
                    add_unless(accx[q].x, sx.x, no_x & (1u << c0));
                    add_unless(accy[q].x, sy.x, no_y & (1u << c0));
                    add_unless(accz[q].x, sz.x, hits & (1u << c0));
                    if (on1) {
                        add_unless(accx[q].y, sx.y, no_x & (1u << c1));
                        add_unless(accy[q].y, sy.y, no_y & (1u << c1));
                        add_unless(accz[q].y, sz.y, hits & (1u << c1));
                    }
                }
            }
        }
    };
    if (first + kCtx * kThreads <= hw) {
        for (int sample = 0; sample < p.spp; ++sample) sample_all(std::true_type{});
    } else {
        for (int sample = 0; sample < p.spp; ++sample) sample_all(std::false_type{});
    }

    // ---- write back -------------------------------------------------------------------------
#pragma unroll
    for (int c = 0; c < kCtx; ++c) {
        const int pix = first + c * kThreads + tid;
        const bool active = c < nctx;
        const int64_t idx = (int64_t)e * hw + pix;
        uint32_t r8 = 0, g8 = 0, b8 = 0;
        if (active) {
            const Rng32 st = lds_state(sbase + c * kStride);
            *reinterpret_cast<uint4 *>(p.states + idx) = make_uint4(st.a, st.b, st.c, st.d);
            const float ax = (c & 1) ? accx[c >> 1].y : accx[c >> 1].x;
            const float ay = (c & 1) ? accy[c >> 1].y : accy[c >> 1].x;
            const float az = (c & 1) ? accz[c >> 1].y : accz[c >> 1].x;
            r8 = (uint32_t)__float2uint_rz(__fmul_rn(ax, p.scale)) & 0xffu;
            g8 = (uint32_t)__float2uint_rz(__fmul_rn(ay, p.scale)) & 0xffu;
            b8 = (uint32_t)__float2uint_rz(__fmul_rn(az, p.scale)) & 0xffu;
        }
        if (p.gray) {
            const uint32_t g = (9798u * r8 + 19235u * g8 + 3735u * b8 + 16384u) >> 15;
            uint32_t w = g;
            w |= __shfl_down_sync(0xffffffffu, g, 1) << 8;
            w |= __shfl_down_sync(0xffffffffu, g, 2) << 16;
            w |= __shfl_down_sync(0xffffffffu, g, 3) << 24;
            if ((tid & 3) == 0 && active) {
                if (pix + 3 < hw && ((reinterpret_cast<uintptr_t>(p.gray) + idx) & 3) == 0) {
                    *reinterpret_cast<uint32_t *>(p.gray + idx) = w;
                } else {
                    for (int j = 0; j < 4; ++j)
                        if (pix + j < hw) p.gray[idx + j] = (uint8_t)(w >> (8 * j));
                }
            }
        }
        if (p.rgb && active) {
            uint8_t *out = p.rgb + idx * 3;
            out[0] = (uint8_t)r8;
            out[1] = (uint8_t)g8;
            out[2] = (uint8_t)b8;
        }
    }
}

}  // namespace rf

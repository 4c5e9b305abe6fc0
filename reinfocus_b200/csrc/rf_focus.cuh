// Focus measure: the B200 replacement of reference vision.py:11-39
//     var( cv2.Laplacian( cv2.medianBlur( cv2.cvtColor(img, RGB2GRAY), 3 ), CV_8U ) )
// restated exactly (probed against OpenCV 4.13):
//     gray = (9798 R + 19235 G + 3735 B + 16384) >> 15
//     med  = 3x3 median, replicated border
//     lap  = clamp(N + S + E + W - 4 C, 0, 255) on med, reflect-101 border
//     var  = (N * sum(lap^2) - sum(lap)^2) / N^2          (population variance)
// Everything up to the two sums is integer-exact; the sums are 64-bit integers and only
// the final quotient is floating point (one float64 division).
//
// Layout: grid = (bands, envs). A block owns `rows` image rows of one env. It stages the
// gray rows it needs (+-2 halo, already clamped) in shared memory, computes the median
// rows (+-1 halo) into shared memory, then the Laplacian, reduces sum / sum^2 with warp
// shuffles and adds them to the env's two 64-bit accumulators. The last block of an env
// (atomic ticket) turns the sums into the variance and re-arms the accumulators, so the
// whole measure is a single launch.
#pragma once

#include <cstdint>

namespace rf {

struct FocusParams {
    const uint8_t *img;  // [n, H, W, channels]
    double *out;         // [n]
    uint8_t *median;     // optional debug plane [n, H, W]
    uint8_t *laplacian;  // optional debug plane [n, H, W]
    unsigned long long *accum;  // [n, 2] sum, sum of squares (zero on entry, zero on exit)
    unsigned int *tickets;      // [n] (zero on entry, zero on exit)
    int n, H, W, channels;
    int rows;    // image rows per block
    int pitch;   // shared-memory row pitch in bytes (>= W, multiple of 4)
};

constexpr int kFocusThreads = 256;

__device__ __forceinline__ int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
    return i;
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return min(max(v, lo), hi); }

__device__ __forceinline__ uint32_t med3(uint32_t a, uint32_t b, uint32_t c) {
    return max(min(a, b), min(max(a, b), c));
}

// median of 9 from three vertically sorted columns (lo, mid, hi each):
// med9 = med3( max(lo's), med3(mid's), min(hi's) )
__device__ __forceinline__ void sort3(uint32_t &a, uint32_t &b, uint32_t &c) {
    const uint32_t lo = min(min(a, b), c), hi = max(max(a, b), c);
    b = a + b + c - lo - hi;
    a = lo;
    c = hi;
}

__global__ void __launch_bounds__(kFocusThreads) focus_kernel(const FocusParams p) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int e = blockIdx.y;
    const int y0 = blockIdx.x * p.rows;
    const int y1 = min(y0 + p.rows, p.H);
    const int H = p.H, W = p.W, pitch = p.pitch;
    const int g_rows = (y1 - y0) + 4;  // gray rows y0-2 .. y1+1
    const int m_rows = (y1 - y0) + 2;  // median rows y0-1 .. y1
    uint8_t *g = smem;                            // [g_rows][pitch]
    uint8_t *m = smem + (size_t)(p.rows + 4) * pitch;  // [m_rows][pitch]

    // ---- stage gray rows (clamped = replicated border in y) -------------------------
    const uint8_t *img = p.img + (int64_t)e * H * W * p.channels;
    if (p.channels == 1) {
        const bool vec = (W % 4 == 0) && ((reinterpret_cast<uintptr_t>(img) & 3) == 0);
        if (vec) {
            const int wq = W / 4;
            for (int i = threadIdx.x; i < g_rows * wq; i += kFocusThreads) {
                const int r = i / wq, q = i - r * wq;
                const int sy = clampi(y0 - 2 + r, 0, H - 1);
                reinterpret_cast<uint32_t *>(g + (size_t)r * pitch)[q] =
                    __ldg(reinterpret_cast<const uint32_t *>(img + (int64_t)sy * W) + q);
            }
        } else {
            for (int i = threadIdx.x; i < g_rows * W; i += kFocusThreads) {
                const int r = i / W, x = i - r * W;
                const int sy = clampi(y0 - 2 + r, 0, H - 1);
                g[(size_t)r * pitch + x] = __ldg(img + (int64_t)sy * W + x);
            }
        }
    } else {
        for (int i = threadIdx.x; i < g_rows * W; i += kFocusThreads) {
            const int r = i / W, x = i - r * W;
            const int sy = clampi(y0 - 2 + r, 0, H - 1);
            const uint8_t *px = img + ((int64_t)sy * W + x) * 3;
            const uint32_t R = __ldg(px), G = __ldg(px + 1), B = __ldg(px + 2);
            g[(size_t)r * pitch + x] = (uint8_t)((9798u * R + 19235u * G + 3735u * B + 16384u) >> 15);
        }
    }
    __syncthreads();

    // ---- median rows y0-1 .. y1 (rows outside the image are never read later) -------
    for (int i = threadIdx.x; i < m_rows * W; i += kFocusThreads) {
        const int r = i / W, x = i - r * W;
        const int y = y0 - 1 + r;
        if (y < 0 || y >= H) continue;
        // gray row of image row yy sits at smem row yy - (y0 - 2); rows are pre-clamped
        const uint8_t *r0 = g + (size_t)(r)*pitch;      // y - 1
        const uint8_t *r1 = g + (size_t)(r + 1) * pitch;  // y
        const uint8_t *r2 = g + (size_t)(r + 2) * pitch;  // y + 1
        const int xl = max(x - 1, 0), xr = min(x + 1, W - 1);
        uint32_t a0 = r0[xl], a1 = r1[xl], a2 = r2[xl];
        uint32_t b0 = r0[x], b1 = r1[x], b2 = r2[x];
        uint32_t c0 = r0[xr], c1 = r1[xr], c2 = r2[xr];
        sort3(a0, a1, a2);
        sort3(b0, b1, b2);
        sort3(c0, c1, c2);
        const uint32_t lo = max(max(a0, b0), c0);
        const uint32_t mid = med3(a1, b1, c1);
        const uint32_t hi = min(min(a2, b2), c2);
        const uint32_t med = med3(lo, mid, hi);
        m[(size_t)r * pitch + x] = (uint8_t)med;
        if (p.median && y >= y0 && y < y1) p.median[((int64_t)e * H + y) * W + x] = (uint8_t)med;
    }
    __syncthreads();

    // ---- Laplacian + sums -----------------------------------------------------------
    unsigned long long sum = 0, sum2 = 0;
    const int rows = y1 - y0;
    for (int i = threadIdx.x; i < rows * W; i += kFocusThreads) {
        const int r = i / W, x = i - r * W;
        const int y = y0 + r;
        // median row of image row yy sits at smem row yy - (y0 - 1)
        const int ru = reflect101(y - 1, H) - (y0 - 1);
        const int rd = reflect101(y + 1, H) - (y0 - 1);
        const int rc = r + 1;
        const int xl = reflect101(x - 1, W), xr = reflect101(x + 1, W);
        int l = (int)m[(size_t)ru * pitch + x] + (int)m[(size_t)rd * pitch + x] +
                (int)m[(size_t)rc * pitch + xl] + (int)m[(size_t)rc * pitch + xr] -
                4 * (int)m[(size_t)rc * pitch + x];
        l = clampi(l, 0, 255);
        if (p.laplacian) p.laplacian[((int64_t)e * H + y) * W + x] = (uint8_t)l;
        sum += (unsigned)l;
        sum2 += (unsigned)(l * l);
    }
    for (int off = 16; off > 0; off >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, off);
        sum2 += __shfl_down_sync(0xffffffffu, sum2, off);
    }
    __shared__ unsigned long long wsum[kFocusThreads / 32], wsum2[kFocusThreads / 32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { wsum[warp] = sum; wsum2[warp] = sum2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long s = 0, s2 = 0;
        for (int w = 0; w < kFocusThreads / 32; ++w) { s += wsum[w]; s2 += wsum2[w]; }
        atomicAdd(&p.accum[2 * e], s);
        atomicAdd(&p.accum[2 * e + 1], s2);
        __threadfence();
        const unsigned int ticket = atomicAdd(&p.tickets[e], 1u);
        is_last = (ticket == gridDim.x - 1);
        if (is_last) {
            __threadfence();
            const unsigned long long S = atomicExch(&p.accum[2 * e], 0ull);
            const unsigned long long S2 = atomicExch(&p.accum[2 * e + 1], 0ull);
            p.tickets[e] = 0;
            // var = (N*S2 - S*S) / N^2 with an exact 128-bit numerator
            const unsigned long long N = (unsigned long long)H * (unsigned long long)W;
            const unsigned long long a_lo = N * S2, a_hi = __umul64hi(N, S2);
            const unsigned long long b_lo = S * S, b_hi = __umul64hi(S, S);
            const unsigned long long d_lo = a_lo - b_lo;
            const unsigned long long d_hi = a_hi - b_hi - (a_lo < b_lo ? 1ull : 0ull);
            // numerator -> float64, round-to-nearest-even from 128 bits
            double num;
            if (d_hi == 0) {
                num = __ull2double_rn(d_lo);
            } else {
                // keep 64 significant bits with a sticky bit, then scale
                const int lz = __clzll((long long)d_hi);
                const unsigned long long top =
                    lz ? ((d_hi << lz) | (d_lo >> (64 - lz))) : d_hi;
                const unsigned long long rest = lz ? (d_lo << lz) : d_lo;
                num = __ull2double_rn(top | (rest ? 1ull : 0ull)) * exp2((double)(64 - lz));
            }
            const double Nd = (double)N;
            p.out[e] = __ddiv_rn(num, __dmul_rn(Nd, Nd));
        }
    }
}

}  // namespace rf

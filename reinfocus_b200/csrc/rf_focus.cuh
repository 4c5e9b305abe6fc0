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

#include <type_traits>

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

// var = (N*S2 - S*S) / N^2 with an exact 128-bit numerator, one float64 division
__device__ __forceinline__ double exact_variance(unsigned long long S, unsigned long long S2,
                                                 unsigned long long N) {
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
        const unsigned long long top = lz ? ((d_hi << lz) | (d_lo >> (64 - lz))) : d_hi;
        const unsigned long long rest = lz ? (d_lo << lz) : d_lo;
        num = __ull2double_rn(top | (rest ? 1ull : 0ull)) * exp2((double)(64 - lz));
    }
    const double Nd = (double)N;
    return __ddiv_rn(num, __dmul_rn(Nd, Nd));
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
            p.out[e] = exact_variance(S, S2, (unsigned long long)H * (unsigned long long)W);
        }
    }
}

// =========================================================================================
// Packed variant (the one the step path uses): no shared memory, no block barrier.
//
// A warp owns a tile of 120 x `band` output pixels of one env and marches down its rows.
// Lane l holds four horizontally adjacent pixels (one 32-bit gray word) at columns
// x0 = seg*120 - 4 + 4*l; lanes 0 and 31 are halo lanes whose medians feed the Laplacian of
// their neighbours. All pixel arithmetic runs two pixels per instruction on zero-extended
// u16x2 lanes with the native 3-input min/max (VIMNMX3.U16x2):
//   column pass   lo/mid/hi of each vertical triple (mid = sum - lo - hi)
//   row pass      median9 = med3(max3(lo's), med3(mid's), min3(hi's))
//   Laplacian     N+S+E+W+1024-4C >= 0 per lane, clamp to [1024, 1279]: low byte = result
//   sums          IDP.4A on the four packed 8-bit Laplacians
// Horizontal neighbours come from the adjacent lanes by shuffle, vertical ones from the two
// previous rows kept in registers. Borders: replicated gray (median), reflect-101 medians
// (Laplacian), exactly as the staged kernel above. Requires W % 4 == 0, W >= 8, H >= 2 and a
// 4-byte aligned image; anything else (and the debug planes) uses focus_kernel.
//
// Round 2 (the kernel is bound by the ALU pipe, which takes VIMNMX3 / PRMT / IADD3 / SEL at one
// warp instruction per two clocks and scheduler, profiles/r02/pipe_microbench.txt):
//   * when the last column segment is at most 60 pixels wide (W = 300: 120 + 120 + 60) a warp
//     runs it for two envs at once, one per half-warp, instead of leaving half its lanes idle;
//   * the rows that need a border rule (the first Laplacian row of the image's top band, the
//     last three rows of a band) are peeled off, so the row loop has no selects or row clamps;
//   * the median's "sum minus min minus max" runs partly as IMAD x * 1 + y on the FMA pipe,
//     which issues beside the ALU pipe (the 1 comes from the parameter block so that ptxas
//     keeps the multiply).
// =========================================================================================

constexpr int kPackedWarps = 8;
constexpr int kPackedBlocksPerSM = 4;  // 54 registers; capping them at 48 for a fifth block spills in the row loop: no gain
constexpr int kPackedCols = 120;  // productive columns per warp tile
constexpr int kPackedHalfCols = 60;  // a last segment this narrow is run for two envs per warp
constexpr int kPackedMaxBand = 512;  // rows per warp tile: keeps the 32-bit tile sums from wrapping

struct PackedFocusParams {
    const uint8_t *img;  // [n, H, W, channels]
    double *out;
    unsigned long long *accum;
    unsigned int *tickets;
    int n, H, W, channels;
    int band;         // output rows per tile
    int segs, bands;  // tiles per env = segs * bands
    int full_segs;    // segments that own a whole warp: segs, or segs - 1 when the last one is shared
    int pairs;        // env pairs sharing last-segment tiles: (n + 1) / 2, or 0 without sharing
    uint32_t one, minus_one;  // 1 and -1, opaque to the compiler (IMAD adds, see above)
};

__device__ __forceinline__ uint32_t min3x2(uint32_t a, uint32_t b, uint32_t c) { return __vimin3_u16x2(a, b, c); }
__device__ __forceinline__ uint32_t max3x2(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_u16x2(a, b, c); }

template <int kChannels>
__device__ __forceinline__ uint32_t load_gray_word(const uint8_t *row, int x, int W) {
    // four gray pixels x..x+3 of one image row, replicated outside [0, W)
    if (x < 0) {
        const uint32_t g = kChannels == 1
                               ? row[0]
                               : (9798u * row[0] + 19235u * row[1] + 3735u * row[2] + 16384u) >> 15;
        return g * 0x01010101u;
    }
    if (x >= W) {
        const uint8_t *px = row + (size_t)(W - 1) * kChannels;
        const uint32_t g = kChannels == 1
                               ? px[0]
                               : (9798u * px[0] + 19235u * px[1] + 3735u * px[2] + 16384u) >> 15;
        return g * 0x01010101u;
    }
    if (kChannels == 1) return __ldg(reinterpret_cast<const uint32_t *>(row + x));
    const uint32_t *w = reinterpret_cast<const uint32_t *>(row + (size_t)x * 3);
    const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
    // bytes: R0 G0 B0 R1 | G1 B1 R2 G2 | B2 R3 G3 B3
    const uint32_t g0 = (9798u * (w0 & 255u) + 19235u * ((w0 >> 8) & 255u) + 3735u * ((w0 >> 16) & 255u) + 16384u) >> 15;
    const uint32_t g1 = (9798u * (w0 >> 24) + 19235u * (w1 & 255u) + 3735u * ((w1 >> 8) & 255u) + 16384u) >> 15;
    const uint32_t g2 = (9798u * ((w1 >> 16) & 255u) + 19235u * (w1 >> 24) + 3735u * (w2 & 255u) + 16384u) >> 15;
    const uint32_t g3 = (9798u * ((w2 >> 8) & 255u) + 19235u * ((w2 >> 16) & 255u) + 3735u * (w2 >> 24) + 16384u) >> 15;
    return g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
}

template <int kChannels>
__global__ void __launch_bounds__(kPackedWarps * 32) focus_packed_kernel(const PackedFocusParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t flat = (int64_t)blockIdx.x * kPackedWarps + (threadIdx.x >> 5);
    const int H = p.H, W = p.W;
    // which tile: a whole warp on one env, or (last, narrow segment) a half-warp per env. The
    // flat order keeps the warps that touch the same image rows next to each other - all
    // segments of an (env, band), and with sharing both envs' segments and then their shared
    // tile - so that the 32-byte sectors straddling two segments are fetched from HBM once.
    bool shared_tile = false;  // warp-uniform
    int e, seg, band, hl = lane;
    bool valid = true;
    if (p.pairs == 0) {
        const int tiles_per_env = p.segs * p.bands;
        e = (int)(flat / tiles_per_env);
        if (e >= p.n) return;  // whole warp
        const int tile = (int)(flat - (int64_t)e * tiles_per_env);
        seg = tile % p.segs;
        band = tile / p.segs;
    } else {
        const int unit = 2 * p.full_segs + 1;  // warps per (env pair, band)
        const int64_t group = flat / unit;
        const int k = (int)(flat - group * unit);
        const int64_t pair = group / p.bands;
        if (pair >= p.pairs) return;  // whole warp
        band = (int)(group - pair * p.bands);
        if (k < 2 * p.full_segs) {
            const int second = k >= p.full_segs ? 1 : 0;
            e = (int)(2 * pair) + second;
            if (e >= p.n) return;  // odd batch: the last pair has one env
            seg = k - second * p.full_segs;
        } else {
            shared_tile = true;
            seg = p.segs - 1;
            hl = lane & 15;
            e = (int)(2 * pair) + (lane >> 4);
            // an odd batch leaves the last upper half-warp without an env: it shadows the lower
            // half's (in-bounds loads, same shuffles) and neither counts nor reports
            valid = e < p.n;
            if (!valid) e -= 1;
        }
    }
    const int y0 = band * p.band, y1 = min(y0 + p.band, H);
    const int x0 = seg * kPackedCols - 4 + 4 * hl;
    // this lane's four Laplacians count iff it is a productive lane inside the image
    const bool counts = valid && hl >= 1 && (shared_tile || lane <= 30) && x0 < W;
    const bool at_left = x0 == 0, at_right = x0 + 4 == W;

    const uint8_t *img = p.img + (size_t)e * H * W * kChannels;
    const size_t pitch = (size_t)W * kChannels;
    const uint32_t one = p.one, minus_one = p.minus_one;
    // a + b + c - lo - hi with the first add and the last subtraction on the FMA pipe
    auto middle = [&](uint32_t a, uint32_t b, uint32_t c, uint32_t lo, uint32_t hi) -> uint32_t {
        const uint32_t t = a * one + b;
        const uint32_t u = t + c - lo;
        return hi * minus_one + u;
    };
    auto med3x2 = [&](uint32_t a, uint32_t b, uint32_t c) -> uint32_t {
        return middle(a, b, c, min3x2(a, b, c), max3x2(a, b, c));
    };

    uint32_t a1 = 0, a2 = 0, a3 = 0, a4 = 0;  // packs of row r-2
    uint32_t b1 = 0, b2 = 0, b3 = 0, b4 = 0;  // packs of row r-1
    uint32_t m2a = 0, m2b = 0, m1a = 0, m1b = 0;      // medians of rows r-3 and r-2
    uint32_t sum = 0, sum2 = 0;
    // the Laplacian is kept non-negative per 16-bit lane by a bias of 1024, whose low byte is
    // zero: after clamping to [1024, 1279] the low bytes are the saturated Laplacians
    const uint32_t bias = 1024u | (1024u << 16), top = 1279u | (1279u << 16);
    // (m-1, m0) and (m3, m4) for the Laplacian's left / right neighbours; at the image edges
    // reflect-101 makes m-1 := m1 and m4 := m2, which is just another byte selection
    const uint32_t select_left = at_left ? 0x5476u : 0x5432u;
    const uint32_t select_right = at_right ? 0x1032u : 0x5432u;
    // the gray pair (x0+3, x0+4) takes x0+4 from the right lane's first pair, or repeats x0+3
    // at the image's right edge (so that lane may belong to another env's half-warp)
    const uint32_t select_c4 = at_right ? 0x1717u : 0x1017u;

    // gray input: one aligned word per lane and row. Lanes left / right of the image load the
    // first / last word of the row and replicate its outer byte (the border rule) with a
    // per-lane byte-permute selector, so the row loop has no column cases. The row pointer
    // walks down with the loop and stops at the image's first / last row (replicated rows).
    const uint8_t *column = img + (x0 < 0 ? 0 : (x0 >= W ? W - 4 : x0));
    const uint32_t replicate = x0 < 0 ? 0x0000u : (x0 >= W ? 0x3333u : 0x3210u);
    const uint8_t *row_ptr = column + (size_t)min(max(y0 - 2, 0), H - 1) * pitch;
    auto load_row = [&](int row) -> uint32_t {
        if (kChannels == 1) return __byte_perm(__ldg(reinterpret_cast<const uint32_t *>(row_ptr)), 0u, replicate);
        return load_gray_word<kChannels>(img + (size_t)min(max(row, 0), H - 1) * pitch, x0, W);
    };

    // one image row r enters: pack it; with kMedian the medians of row r-1 follow, with
    // kLaplacian the Laplacians of row r-2. The next row's word is requested before the
    // arithmetic so that its latency hides behind it (two rows ahead: no gain). kInterior:
    // rows r + 1 and r - 2 need no border rule (0 < r - 2 < H - 1, r + 1 <= H - 1).
    uint32_t g_next = load_row(y0 - 2);
    auto step = [&](int r, auto median_tag, auto laplacian_tag, auto interior_tag) {
        constexpr bool kMedian = decltype(median_tag)::value, kLaplacian = decltype(laplacian_tag)::value;
        constexpr bool kInterior = decltype(interior_tag)::value;
        const uint32_t g = g_next;
        if (kInterior) {
            row_ptr += pitch;
            g_next = load_row(r + 1);
        } else {
            if (r >= 0 && r < H - 1) row_ptr += pitch;  // row r -> r + 1, clamped to the image
            if (r <= y1) g_next = load_row(r + 1);
        }
        // zero-extended column pairs (x0,x0+1) (x0+1,x0+2) (x0+2,x0+3) (x0+3,x0+4); the pair
        // (x0-1,x0) is the left lane's fourth pair, so its sorted triple comes by shuffle
        const uint32_t c1 = __byte_perm(g, 0u, 0x4140);
        const uint32_t c2 = __byte_perm(g, 0u, 0x4241);
        const uint32_t c3 = __byte_perm(g, 0u, 0x4342);
        const uint32_t c4 = __byte_perm(__shfl_down_sync(0xffffffffu, c1, 1), g, select_c4);
        if (kMedian) {
            // column pass on rows r-2, r-1, r
            const uint32_t lo1 = min3x2(a1, b1, c1), hi1 = max3x2(a1, b1, c1), md1 = middle(a1, b1, c1, lo1, hi1);
            const uint32_t lo2 = min3x2(a2, b2, c2), hi2 = max3x2(a2, b2, c2), md2 = middle(a2, b2, c2, lo2, hi2);
            const uint32_t lo3 = min3x2(a3, b3, c3), hi3 = max3x2(a3, b3, c3), md3 = middle(a3, b3, c3, lo3, hi3);
            const uint32_t lo4 = min3x2(a4, b4, c4), hi4 = max3x2(a4, b4, c4), md4 = middle(a4, b4, c4, lo4, hi4);
            const uint32_t lo0 = __shfl_up_sync(0xffffffffu, lo4, 1);
            const uint32_t md0 = __shfl_up_sync(0xffffffffu, md4, 1);
            const uint32_t hi0 = __shfl_up_sync(0xffffffffu, hi4, 1);
            // row pass: medians of row r-1 for pixels (x0, x0+1) and (x0+2, x0+3)
            const uint32_t ma = med3x2(max3x2(lo0, lo1, lo2), med3x2(md0, md1, md2), min3x2(hi0, hi1, hi2));
            const uint32_t mb = med3x2(max3x2(lo2, lo3, lo4), med3x2(md2, md3, md4), min3x2(hi2, hi3, hi4));
            if (kLaplacian) {
                // Laplacian of row y = r-2: centre m1, up m2 (row r-3), down m (row r-1);
                // reflect-101 at the top / bottom image rows
                const int y = r - 2;
                const uint32_t ua = !kInterior && y == 0 ? ma : m2a, ub = !kInterior && y == 0 ? mb : m2b;
                const uint32_t da = !kInterior && y == H - 1 ? m2a : ma, db = !kInterior && y == H - 1 ? m2b : mb;
                const uint32_t nl = __shfl_up_sync(0xffffffffu, m1b, 1);
                const uint32_t nr = __shfl_down_sync(0xffffffffu, m1a, 1);
                const uint32_t mid = __byte_perm(m1a, m1b, 0x5432);  // (m1, m2)
                const uint32_t la = __byte_perm(nl, m1a, select_left);
                const uint32_t rb = __byte_perm(m1b, nr, select_right);
                uint32_t va = ua + da + la;
                va = va + mid + bias - 4u * m1a;
                uint32_t vb = ub + db + mid;
                vb = vb + rb + bias - 4u * m1b;
                va = max3x2(va, bias, va);  // the register twice: VIMNMX3 takes one immediate
                vb = max3x2(vb, bias, vb);
                va = min3x2(va, top, va);
                vb = min3x2(vb, top, vb);
                // every lane sums; the lanes that do not count drop theirs after the loop
                const uint32_t l4 = __byte_perm(va, vb, 0x6420);
                sum = __dp4a(l4, 0x01010101u, sum);
                sum2 = __dp4a(l4, l4, sum2);
            }
            m2a = m1a; m2b = m1b;
            m1a = ma; m1b = mb;
        }
        a1 = b1; a2 = b2; a3 = b3; a4 = b4;
        b1 = c1; b2 = c2; b3 = c3; b4 = c4;
    };
    using yes = std::true_type;
    using no = std::false_type;
    step(y0 - 2, no{}, no{}, no{});
    step(y0 - 1, no{}, no{}, no{});
    step(y0, yes{}, no{}, no{});
    step(y0 + 1, yes{}, no{}, no{});
    step(y0 + 2, yes{}, yes{}, no{});  // the Laplacian of row y0: reflects upwards when y0 == 0
    int r = y0 + 3;
#pragma unroll 6
    for (; r <= y1 - 2; ++r) step(r, yes{}, yes{}, yes{});
    for (; r <= y1 + 1; ++r) step(r, yes{}, yes{}, no{});  // rows y1 - 1 .. y1 + 1 may lie below the image

    if (!counts) sum = sum2 = 0;
    for (int off = 8; off > 0; off >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, off, 16);
        sum2 += __shfl_down_sync(0xffffffffu, sum2, off, 16);
    }
    const uint32_t upper = __shfl_down_sync(0xffffffffu, sum, 16), upper2 = __shfl_down_sync(0xffffffffu, sum2, 16);
    if (!shared_tile) {
        sum += upper;
        sum2 += upper2;
    }
    if (lane == 0 || (shared_tile && valid && lane == 16)) {
        atomicAdd(&p.accum[2 * e], (unsigned long long)sum);
        atomicAdd(&p.accum[2 * e + 1], (unsigned long long)sum2);
        __threadfence();
        const unsigned int ticket = atomicAdd(&p.tickets[e], 1u);
        if (ticket == (unsigned)(p.segs * p.bands) - 1) {
            __threadfence();
            const unsigned long long S = atomicExch(&p.accum[2 * e], 0ull);
            const unsigned long long S2 = atomicExch(&p.accum[2 * e + 1], 0ull);
            p.tickets[e] = 0;
            p.out[e] = exact_variance(S, S2, (unsigned long long)H * (unsigned long long)W);
        }
    }
}

}  // namespace rf

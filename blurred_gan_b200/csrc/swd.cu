// Sliced-Wasserstein evaluation kernels for sm_100a (everything except the sort, see sort.cu).
//
// Replaces the numpy/scipy arithmetic of the reference's sliced_wasserstein.py:
//   pyr_down / pyr_up / Laplacian residual   :65-88   (5x5 binomial, REFLECT_101 border, fp64 accumulate)
//   get_descriptors_for_minibatch            :13-23   (7x7xC patch gather, x/y transposed -- SURVEY Q10)
//   finalize_descriptors                     :27-34   (per-channel mean / population std)
//   np.matmul(A, dirs)                       :45-46   (fp32 projection onto unit directions)
//   np.mean(np.abs(projA - projB))           :49-50
#include "common.cuh"

namespace bgan {
namespace {

// REFLECT_101 ("mirror" in scipy.ndimage): ... 2 1 | 0 1 2 ... n-1 | n-2 n-3 ...
__device__ __forceinline__ int mirror101(int i, int n) {
    if (n == 1) return 0;
    const int period = 2 * (n - 1);
    i %= period;
    if (i < 0) i += period;
    return i < n ? i : period - i;
}

__constant__ double kB5[5] = {1.0, 4.0, 6.0, 4.0, 1.0};

// y[p, oy, ox] = sum_{dy,dx} B5[dy] B5[dx] / 256 * x[p, m(2 oy + dy - 2), m(2 ox + dx - 2)]
// Every product and partial sum is exact in fp64 (weights are small integers / 256, inputs fp32),
// so the result is the correctly rounded exact value -- bit-identical to scipy's fp64 accumulation.
// Index arithmetic is 32-bit (rows = planes * output height fits an int, checked by the launcher) and the border is
// folded with compares when the image has >= 3 rows / columns (one reflection suffices for offsets in [-2, n + 1]); a
// per-pixel 64-bit div / mod and the `%` of the general mirror made these two kernels ~3x slower than their loads.
__device__ __forceinline__ int mirror101_near(int i, int n, bool small) {
    if (small) return mirror101(i, n);
    i = i < 0 ? -i : i;
    return i < n ? i : 2 * (n - 1) - i;
}

// (vblock of vgrid: the block's place in the grid-stride loop -- a launch may run two bodies side by side, see pyr_pair_kernel)
__device__ __forceinline__ void pyr_down_body(const float* __restrict__ x, float* __restrict__ y, int planes, int H, int W, int oh, int ow,
                                              unsigned vblock, unsigned vgrid) {
    const unsigned total = static_cast<unsigned>(planes) * oh * ow;
    const bool small = H < 3 || W < 3;
    for (unsigned idx = vblock * blockDim.x + threadIdx.x; idx < total; idx += vgrid * blockDim.x) {
        const unsigned r = idx / static_cast<unsigned>(ow);
        const int ox = static_cast<int>(idx - r * ow);
        const unsigned p = r / static_cast<unsigned>(oh);
        const int oy = static_cast<int>(r - p * oh);
        const float* src = x + static_cast<size_t>(p) * H * W;
        int cols[5];
#pragma unroll
        for (int d = 0; d < 5; ++d) cols[d] = mirror101_near(2 * ox + d - 2, W, small);
        double acc = 0.0;
#pragma unroll
        for (int dy = 0; dy < 5; ++dy) {
            const float* row = src + static_cast<size_t>(mirror101_near(2 * oy + dy - 2, H, small)) * W;
            double racc = 0.0;
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) racc += kB5[dx] * static_cast<double>(__ldg(row + cols[dx]));
            acc += kB5[dy] * racc;
        }
        y[idx] = static_cast<float>(acc * (1.0 / 256.0));
    }
}

__global__ void __launch_bounds__(256) pyr_down_kernel(const float* __restrict__ x, float* __restrict__ y, int planes, int H, int W, int oh,
                                                       int ow) {
    pdl_wait();
    pyr_down_body(x, y, planes, H, W, oh, ow, blockIdx.x, gridDim.x);
}

// y[p, Y, X] = (minuend ? minuend[p, Y, X] : 0) -/+ sum 4 B5 B5 / 256 * Z[m(Y+dy-2), m(X+dx-2)],
// Z = x zero-stuffed to (2h, 2w): Z[2i, 2j] = x[i, j], zero elsewhere.
__global__ void __launch_bounds__(256) pyr_up_kernel(const float* __restrict__ x, const float* __restrict__ minuend, float* __restrict__ y,
                                                     int planes, int h, int w) {
    const int OH = 2 * h, OW = 2 * w;
    const unsigned total = static_cast<unsigned>(planes) * OH * OW;
    const bool small = OH < 3 || OW < 3;
    pdl_wait();
    for (unsigned idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const unsigned r = idx / static_cast<unsigned>(OW);
        const int X = static_cast<int>(idx - r * OW);
        const unsigned p = r / static_cast<unsigned>(OH);
        const int Y = static_cast<int>(r - p * OH);
        const float* src = x + static_cast<size_t>(p) * h * w;
        int zx[5];
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) zx[dx] = mirror101_near(X + dx - 2, OW, small);
        double acc = 0.0;
#pragma unroll
        for (int dy = 0; dy < 5; ++dy) {
            const int zy = mirror101_near(Y + dy - 2, OH, small);
            if (zy & 1) continue;
            const float* row = src + static_cast<size_t>(zy >> 1) * w;
            double racc = 0.0;
#pragma unroll
            for (int dx = 0; dx < 5; ++dx)
                if (!(zx[dx] & 1)) racc += kB5[dx] * static_cast<double>(__ldg(row + (zx[dx] >> 1)));
            acc += kB5[dy] * racc;
        }
        const float up = static_cast<float>(acc * (4.0 / 256.0));
        y[idx] = minuend ? (__ldg(minuend + idx) - up) : up;
    }
}

// pyr_up for 2h, 2w >= 3: one thread per source pixel (a, b) computes the 2 x 2 output quad (2a .. 2a + 1, 2b .. 2b + 1) from
// the 3 x 3 source neighbourhood.  Zero stuffing leaves an even output row the source rows a - 1, a, a + 1 with weights
// 1, 6, 1 and an odd one the rows a, a + 1 with weights 4, 4 (columns alike); REFLECT_101 on the zero-stuffed grid keeps the
// parity of an index, so at the borders row -1 becomes row 1 and row h becomes row h - 1.  Same products, same exact fp64
// sums, hence the same bits as pyr_up_kernel -- with 9 loads per four outputs instead of ~25, and 8-byte stores.
__device__ __forceinline__ void pyr_up_quad_body(const float* __restrict__ x, const float* __restrict__ minuend, float* __restrict__ y,
                                                 int planes, int h, int w, unsigned vblock, unsigned vgrid) {
    const unsigned total = static_cast<unsigned>(planes) * h * w;
    const int OW = 2 * w;
    for (unsigned idx = vblock * blockDim.x + threadIdx.x; idx < total; idx += vgrid * blockDim.x) {
        const unsigned r = idx / static_cast<unsigned>(w);
        const int b = static_cast<int>(idx - r * w);
        const unsigned p = r / static_cast<unsigned>(h);
        const int a = static_cast<int>(r - p * h);
        const float* src = x + static_cast<size_t>(p) * h * w;
        const int rows[3] = {a == 0 ? 1 : a - 1, a, a + 1 < h ? a + 1 : h - 1};
        const int cols[3] = {b == 0 ? 1 : b - 1, b, b + 1 < w ? b + 1 : w - 1};
        double he[3], ho[3];  // per source row: horizontal sums for the even / odd output column
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const float* row = src + static_cast<size_t>(rows[i]) * w;
            const double s0 = static_cast<double>(__ldg(row + cols[0])), s1 = static_cast<double>(__ldg(row + cols[1])),
                         s2 = static_cast<double>(__ldg(row + cols[2]));
            he[i] = s0 + 6.0 * s1 + s2;
            ho[i] = 4.0 * s1 + 4.0 * s2;
        }
        const float ee = static_cast<float>((he[0] + 6.0 * he[1] + he[2]) * (4.0 / 256.0));
        const float eo = static_cast<float>((ho[0] + 6.0 * ho[1] + ho[2]) * (4.0 / 256.0));
        const float oe = static_cast<float>((4.0 * he[1] + 4.0 * he[2]) * (4.0 / 256.0));
        const float oo = static_cast<float>((4.0 * ho[1] + 4.0 * ho[2]) * (4.0 / 256.0));
        const size_t o0 = (static_cast<size_t>(p) * 2 * h + 2 * a) * OW + 2 * b, o1 = o0 + OW;
        float2 top = make_float2(ee, eo), bot = make_float2(oe, oo);
        if (minuend) {
            const float2 m0 = __ldg(reinterpret_cast<const float2*>(minuend + o0)), m1 = __ldg(reinterpret_cast<const float2*>(minuend + o1));
            top = make_float2(m0.x - ee, m0.y - eo);
            bot = make_float2(m1.x - oe, m1.y - oo);
        }
        *reinterpret_cast<float2*>(y + o0) = top;
        *reinterpret_cast<float2*>(y + o1) = bot;
    }
}

__global__ void __launch_bounds__(256) pyr_up_quad_kernel(const float* __restrict__ x, const float* __restrict__ minuend, float* __restrict__ y,
                                                          int planes, int h, int w) {
    pdl_wait();
    pyr_up_quad_body(x, minuend, y, planes, h, w, blockIdx.x, gridDim.x);
}

// The residual of pyramid level l (pyr_up of its coarse image, subtracted from the level) and the coarse image of level l + 1 both
// need only the coarse image of level l: one launch, the first up_blocks blocks on the residual, the rest on the next reduction.
// A feed of four levels is then five launches instead of seven (the feed phase is bound by the host: ~4 us per launch).
struct PyrPair {
    const float* up_x;       // coarse image of level l: [planes, h, w]
    const float* up_minuend; // level l: [planes, 2h, 2w]
    float* up_y;             // residual of level l
    float* down_y;           // coarse image of level l + 1: [planes, h / 2, w / 2] (its input is up_x)
    int planes, h, w;
    unsigned up_blocks;
};
__global__ void __launch_bounds__(256) pyr_pair_kernel(const PyrPair a) {
    pdl_wait();
    if (blockIdx.x < a.up_blocks)
        pyr_up_quad_body(a.up_x, a.up_minuend, a.up_y, a.planes, a.h, a.w, blockIdx.x, a.up_blocks);
    else
        pyr_down_body(a.up_x, a.down_y, a.planes, a.h, a.w, (a.h + 1) / 2, (a.w + 1) / 2, blockIdx.x - a.up_blocks, gridDim.x - a.up_blocks);
}

// desc[m, c, i, j] = level[m / per_image, c, py[m] + j - half, px[m] + i - half]; per-channel sum / sum-of-squares.
// One warp per descriptor (grid-stride): the lanes walk the C * nhood^2 floats of the descriptor, so the output row is
// written coalesced.  The descriptor geometry does not depend on the descriptor: every lane computes the image offsets and
// channels of its <= kSlots elements once (the divisions by nhood^2 and nhood leave the loop), and the positions of a
// warp's next descriptor are loaded while the current one is gathered.  The channel statistics are accumulated in fp64
// per element slot, folded per channel at the end, reduced per warp and per CTA in shared memory, and only then added to
// the global sums: six atomics per CTA on a grid of at most two CTAs per SM (one atomic per warp and statistic made this
// kernel 75 us per minibatch; 1024 CTAs x 6 atomics on six addresses still 12.8 us, most of it the serialised L2 atomics).
template <int MAXC, int SLOTS>
__device__ __forceinline__ void extract_body(const float* __restrict__ level, int C, int h, int w, const int32_t* __restrict__ px,
                                             const int32_t* __restrict__ py, int per_image, int nhood, int64_t n_desc, float* __restrict__ out,
                                             double* __restrict__ sums) {
    __shared__ double s_part[8][2 * MAXC];
    const int half = nhood / 2;
    const int per_chan = nhood * nhood;
    const int dlen = C * per_chan;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
    int rel[SLOTS], chan[SLOTS];  // element k = lane + 32 t: offset from the patch origin, channel (-1: past the end)
#pragma unroll
    for (int t = 0; t < SLOTS; ++t) {
        const int k = lane + 32 * t;
        const int c = k / per_chan;
        const int ij = k - c * per_chan;
        const int i = ij / nhood, j = ij - i * nhood;
        chan[t] = k < dlen ? c : -1;
        rel[t] = (c * h + j) * w + i;
    }
    double a1[SLOTS], a2[SLOTS];
#pragma unroll
    for (int t = 0; t < SLOTS; ++t) a1[t] = a2[t] = 0.0;
    const int64_t step = static_cast<int64_t>(gridDim.x) * warps;
    int64_t m = static_cast<int64_t>(blockIdx.x) * warps + warp;
    int y_next = 0, x_next = 0;
    pdl_wait();  // (the descriptor geometry above overlaps the tail of the pyramid kernel before this one)
    if (m < n_desc) {
        y_next = __ldg(py + m);
        x_next = __ldg(px + m);
    }
    const int64_t image_stride = static_cast<int64_t>(C) * h * w;
    for (; m < n_desc; m += step) {
        const int y0 = y_next - half, x0 = x_next - half;
        if (m + step < n_desc) {
            y_next = __ldg(py + m + step);
            x_next = __ldg(px + m + step);
        }
        const float* patch = level + (m / per_image) * image_stride + y0 * w + x0;
        float* dst = out + m * dlen + lane;
        float v[SLOTS];
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) v[t] = chan[t] >= 0 ? __ldg(patch + rel[t]) : 0.0f;
#pragma unroll
        for (int t = 0; t < SLOTS; ++t) {
            if (chan[t] >= 0) dst[32 * t] = v[t];
            const double dv = static_cast<double>(v[t]);
            a1[t] += dv;
            a2[t] = fma(dv, dv, a2[t]);
        }
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        double s1 = 0.0, s2 = 0.0;
#pragma unroll
        for (int t = 0; t < SLOTS; ++t)
            if (chan[t] == c) {
                s1 += a1[t];
                s2 += a2[t];
            }
        const double a = warp_sum(s1), b = warp_sum(s2);
        if (lane == 0) {
            s_part[warp][c] = a;
            s_part[warp][MAXC + c] = b;
        }
    }
    __syncthreads();
    if (threadIdx.x < 2 * MAXC) {
        const int c = threadIdx.x % MAXC, which = threadIdx.x / MAXC;
        if (c < C) {
            double t = 0.0;
            for (int q = 0; q < warps; ++q) t += s_part[q][threadIdx.x];
            atomicAdd(sums + which * C + c, t);
        }
    }
}

template <int MAXC, int SLOTS>
__global__ void __launch_bounds__(256) extract_kernel(const float* __restrict__ level, int C, int h, int w, const int32_t* __restrict__ px,
                                                      const int32_t* __restrict__ py, int per_image, int nhood, int64_t n_desc,
                                                      float* __restrict__ out, double* __restrict__ sums) {
    extract_body<MAXC, SLOTS>(level, C, h, w, px, py, per_image, nhood, n_desc, out, sums);
}

// The gathers of all pyramid levels of one `feed` as ONE launch (blockIdx.y = level): four launches of 8192 descriptors each
// are latency-bound (7.5 us each: a warp walks its descriptors one after the other), one launch has four times the warps in
// flight.  The levels' residuals all live in the feed workspace until the end of the call.
constexpr int kMaxFeedLevels = 8;
struct ExtractLevels {
    const float* level[kMaxFeedLevels];
    const int32_t* px[kMaxFeedLevels];
    float* out[kMaxFeedLevels];
    double* sums[kMaxFeedLevels];
    int h[kMaxFeedLevels], w[kMaxFeedLevels];
};
template <int MAXC, int SLOTS>
__global__ void __launch_bounds__(256) extract_levels_kernel(const ExtractLevels lv, int C, int per_image, int nhood, int64_t n_desc) {
    const int l = blockIdx.y;
    extract_body<MAXC, SLOTS>(lv.level[l], C, lv.h[l], lv.w[l], lv.px[l], lv.px[l] + n_desc, per_image, nhood, n_desc, lv.out[l], lv.sums[l]);
}

// desc = (desc - mean_c) / std_c, two fp32 operations like numpy's in-place `-=` then `/=` (a true IEEE division).
// 128-bit accesses over the flat array, kNormUnroll of them in flight per thread; the position inside the descriptor
// (which decides the channel) advances by a constant per grid stride, so there is no per-element division or modulo.
constexpr int kNormUnroll = 4;
__global__ void __launch_bounds__(256) normalise_kernel(float* __restrict__ desc, int64_t total, int C, int per_chan,
                                                        const double* __restrict__ sums, double count) {
    __shared__ float s_mean[8], s_std[8];
    if (threadIdx.x < C) {
        const double mean = sums[threadIdx.x] / count;
        const double var = fmax(sums[C + threadIdx.x] / count - mean * mean, 0.0);
        s_mean[threadIdx.x] = static_cast<float>(mean);
        s_std[threadIdx.x] = static_cast<float>(sqrt(var));
    }
    __syncthreads();
    const int dlen = C * per_chan;
    auto apply = [&](float v, int e) {  // e: position inside the descriptor, 0 .. dlen - 1
        int c = 0;
        for (int b = per_chan; b <= e; b += per_chan) ++c;
        return (v - s_mean[c]) / s_std[c];
    };
    const int64_t n4 = total >> 2;
    const int64_t threads = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t t0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    float4* d4 = reinterpret_cast<float4*>(desc);
    int e = static_cast<int>((4 * t0) % dlen);                    // position of this thread's first element
    const int de = static_cast<int>((4 * threads) % dlen);        // advance per grid stride
    for (int64_t i = t0; i < n4; i += kNormUnroll * threads) {
        float4 v[kNormUnroll];
        int pos[kNormUnroll];
#pragma unroll
        for (int u = 0; u < kNormUnroll; ++u) {
            pos[u] = e;
            e += de;
            if (e >= dlen) e -= dlen;
            if (i + u * threads < n4) v[u] = d4[i + u * threads];
        }
#pragma unroll
        for (int u = 0; u < kNormUnroll; ++u) {
            if (i + u * threads < n4) {
                int p = pos[u];
                float4 r;
                r.x = apply(v[u].x, p);
                p = p + 1 == dlen ? 0 : p + 1;
                r.y = apply(v[u].y, p);
                p = p + 1 == dlen ? 0 : p + 1;
                r.z = apply(v[u].z, p);
                p = p + 1 == dlen ? 0 : p + 1;
                r.w = apply(v[u].w, p);
                d4[i + u * threads] = r;
            }
        }
    }
    // the last total % 4 elements
    const int64_t tail = (n4 << 2) + t0;
    if (tail < total) desc[tail] = apply(desc[tail], static_cast<int>(tail % dlen));
}

// ---------------------------------------------------------------------------------------------
// projection: proj[d, m] = sum_k desc[m, k] * dirs[k, d]     (fp32 FFMA; K = 147 is not worth
// a tensor-core pipeline at 1e-4 parity -- see DESIGN.md "projection")
// CTA tile: 128 rows x 64 directions, whole K resident in shared memory.
// thread tile: 4 rows (lane + 32 i) x 8 directions (warp * 8 ..)
// ---------------------------------------------------------------------------------------------
constexpr int kPM = 128, kPN = 64, kPThreads = 256;

__global__ void __launch_bounds__(kPThreads) project_kernel(const float* __restrict__ desc, int64_t rows, int K,
                                                            const float* __restrict__ dirs, int D, float* __restrict__ proj) {
    extern __shared__ __align__(16) float s_proj[];
    const int KA = K | 1;                       // odd pitch: lane <-> row reads are conflict-free
    float* s_a = s_proj;                        // [kPM][KA]
    float* s_b = s_proj + kPM * KA;             // [K][kPN]
    const int64_t m0 = static_cast<int64_t>(blockIdx.x) * kPM;
    const int d0 = blockIdx.y * kPN;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mrows = (rows - m0 < kPM) ? static_cast<int>(rows - m0) : kPM;

    for (int idx = tid; idx < kPM * K; idx += kPThreads) {
        const int r = idx / K, k = idx - r * K;
        s_a[r * KA + k] = (r < mrows) ? __ldg(desc + (m0 + r) * K + k) : 0.0f;
    }
    for (int idx = tid; idx < K * kPN; idx += kPThreads) {
        const int k = idx / kPN, d = idx - k * kPN;
        s_b[idx] = (d0 + d < D) ? __ldg(dirs + static_cast<size_t>(k) * D + d0 + d) : 0.0f;
    }
    __syncthreads();

    float acc[4][8];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;
    const float* a0 = s_a + lane * KA;
    const float* b0 = s_b + warp * 8;
#pragma unroll 7
    for (int k = 0; k < K; ++k) {
        float av[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) av[i] = a0[i * 32 * KA + k];
        const float4 b_lo = *reinterpret_cast<const float4*>(b0 + k * kPN);
        const float4 b_hi = *reinterpret_cast<const float4*>(b0 + k * kPN + 4);
        const float bv[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int d = d0 + warp * 8 + j;
        if (d >= D) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = lane + 32 * i;
            if (r < mrows) proj[static_cast<size_t>(d) * rows + m0 + r] = acc[i][j];
        }
    }
}

// sum |a - b| in fp64.  128-bit loads, kDiffUnroll pairs of them in flight per thread (one element per iteration kept one
// load in flight per warp: 2.3 TB/s); `vec` is false when a pointer is not 16-byte aligned (sub-views of a projection array).
constexpr int kDiffUnroll = 4;
__global__ void __launch_bounds__(256) absdiff_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t count, bool vec,
                                                          double* __restrict__ out) {
    double acc = 0.0;
    const int64_t threads = static_cast<int64_t>(gridDim.x) * blockDim.x;
    const int64_t t0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    int64_t done = 0;
    if (vec) {
        const int64_t n4 = count >> 2;
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        for (int64_t i = t0; i < n4; i += kDiffUnroll * threads) {
            float4 va[kDiffUnroll], vb[kDiffUnroll];
#pragma unroll
            for (int u = 0; u < kDiffUnroll; ++u) {
                const bool ok = i + u * threads < n4;
                va[u] = ok ? __ldg(a4 + i + u * threads) : make_float4(0.f, 0.f, 0.f, 0.f);
                vb[u] = ok ? __ldg(b4 + i + u * threads) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < kDiffUnroll; ++u) {
                acc += static_cast<double>(fabsf(va[u].x - vb[u].x));
                acc += static_cast<double>(fabsf(va[u].y - vb[u].y));
                acc += static_cast<double>(fabsf(va[u].z - vb[u].z));
                acc += static_cast<double>(fabsf(va[u].w - vb[u].w));
            }
        }
        done = n4 << 2;
    }
    for (int64_t idx = done + t0; idx < count; idx += threads) acc += static_cast<double>(fabsf(__ldg(a + idx) - __ldg(b + idx)));
    acc = warp_sum(acc);
    __shared__ double s_part[32];
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = (threadIdx.x < (blockDim.x >> 5)) ? s_part[threadIdx.x] : 0.0;
        v = warp_sum(v);
        if (threadIdx.x == 0) atomicAdd(out, v);
    }
}

unsigned grid_for(int64_t total, int threads, int per_sm) {
    const int64_t want = ceil_div64(total, threads);
    const int64_t cap = static_cast<int64_t>(device_sm_count()) * per_sm;
    return static_cast<unsigned>(std::max<int64_t>(1, std::min(want, cap)));
}

}  // namespace
}  // namespace bgan

using namespace bgan;

extern "C" int bgan_pyr_down_f32(const float* x, float* y, int planes, int height, int width, void* stream) {
    BGAN_REQUIRE(x && y && planes > 0 && height > 0 && width > 0, BGAN_E_BADARG, "bgan_pyr_down_f32: bad arguments");
    const int oh = (height + 1) / 2, ow = (width + 1) / 2;
    const int64_t total = static_cast<int64_t>(planes) * oh * ow;
    BGAN_REQUIRE(static_cast<int64_t>(planes) * height * width < (int64_t(1) << 31), BGAN_E_BADARG, "bgan_pyr_down_f32: more than 2^31 pixels per call");
    BGAN_CUDA_OK(launch_pdl(pyr_down_kernel, dim3(grid_for(total, 256, 16)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, y, planes, height,
                            width, oh, ow));
    return 0;
}

extern "C" int bgan_pyr_up_f32(const float* x, const float* minuend, float* y, int planes, int h, int w, void* stream) {
    BGAN_REQUIRE(x && y && planes > 0 && h > 0 && w > 0, BGAN_E_BADARG, "bgan_pyr_up_f32: bad arguments");
    const int64_t total = static_cast<int64_t>(planes) * 4 * h * w;
    BGAN_REQUIRE(total < (int64_t(1) << 31), BGAN_E_BADARG, "bgan_pyr_up_f32: more than 2^31 pixels per call");
    const bool aligned8 = reinterpret_cast<uintptr_t>(y) % 8 == 0 && (minuend == nullptr || reinterpret_cast<uintptr_t>(minuend) % 8 == 0);
    if (h >= 2 && w >= 2 && aligned8)
        BGAN_CUDA_OK(launch_pdl(pyr_up_quad_kernel, dim3(grid_for(total / 4, 256, 16)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, minuend, y,
                                planes, h, w));
    else
        BGAN_CUDA_OK(launch_pdl(pyr_up_kernel, dim3(grid_for(total, 256, 16)), dim3(256), 0, static_cast<cudaStream_t>(stream), x, minuend, y, planes, h,
                                w));
    return 0;
}

extern "C" int bgan_swd_extract_f32(const float* level, int n, int c, int h, int w, const int32_t* px, const int32_t* py,
                                    int nhoods_per_image, int nhood, float* desc_out, double* sums, void* stream) {
    BGAN_REQUIRE(level && px && py && desc_out && sums, BGAN_E_BADARG, "bgan_swd_extract_f32: null pointer");
    BGAN_REQUIRE(n > 0 && c > 0 && c <= 4 && nhood > 0 && (nhood & 1) && nhoods_per_image > 0 && h >= nhood && w >= nhood,
                 BGAN_E_BADARG, "bgan_swd_extract_f32: bad shape (n=%d c=%d h=%d w=%d nhood=%d)", n, c, h, w, nhood);
    const int64_t n_desc = static_cast<int64_t>(n) * nhoods_per_image;
    const int64_t total = n_desc * c * nhood * nhood;
    (void)total;
    const int64_t ctas = std::min<int64_t>(ceil_div64(n_desc, 8), 2 * static_cast<int64_t>(device_sm_count()));
    const int dlen = c * nhood * nhood;
    BGAN_REQUIRE(dlen <= 32 * 16, BGAN_E_UNSUPPORTED, "bgan_swd_extract_f32: descriptors of more than 512 floats are not supported (got %d)", dlen);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dlen <= 32 * 5)  // 7 x 7 x 3 = 147: five elements per lane
        extract_kernel<4, 5><<<static_cast<unsigned>(ctas), 256, 0, st>>>(level, c, h, w, px, py, nhoods_per_image, nhood, n_desc, desc_out, sums);
    else
        extract_kernel<4, 16><<<static_cast<unsigned>(ctas), 256, 0, st>>>(level, c, h, w, px, py, nhoods_per_image, nhood, n_desc, desc_out, sums);
    BGAN_LAUNCH_OK();
    return 0;
}

// One `feed` of sliced_wasserstein.API (sliced_wasserstein.py:119-123): Laplacian pyramid of the minibatch and the
// descriptor gather of every level, as one host call (the same kernels as the three entry points above, launched back to
// back on `stream`; per-minibatch host work is what bounds the feed phase of an evaluation).
extern "C" size_t bgan_swd_feed_workspace_bytes(int n, int c, int h, int w, int levels) {
    if (n <= 0 || c <= 0 || h <= 0 || w <= 0 || levels < 1) return 0;
    // level l < levels - 1: its residual (h_l x w_l); level l >= 1: its Gaussian image (h_l x w_l)
    size_t floats = 0;
    int hl = h, wl = w;
    for (int l = 0; l < levels; ++l) {
        const size_t plane = static_cast<size_t>(n) * c * hl * wl;
        if (l < levels - 1) floats += plane;
        if (l >= 1) floats += plane;
        hl = (hl + 1) / 2;
        wl = (wl + 1) / 2;
    }
    return floats * sizeof(float) + 256;
}

extern "C" int bgan_swd_feed_f32(const float* minibatch, int n, int c, int h, int w, int levels, const int32_t* positions,
                                 int nhoods_per_image, int nhood, float* const* desc_out, double* const* sums, void* ws, size_t ws_bytes,
                                 void* stream) {
    BGAN_REQUIRE(minibatch && positions && desc_out && sums && levels >= 1 && levels <= 16, BGAN_E_BADARG, "bgan_swd_feed_f32: bad arguments");
    BGAN_REQUIRE(levels == 1 || (ws && ws_bytes >= bgan_swd_feed_workspace_bytes(n, c, h, w, levels)), BGAN_E_WORKSPACE,
                 "bgan_swd_feed_f32: workspace of %zu bytes required, got %zu", bgan_swd_feed_workspace_bytes(n, c, h, w, levels), ws_bytes);
    for (int l = 0; l < levels; ++l) {
        BGAN_REQUIRE((h >> l) > 0 && (w >> l) > 0 && ((h >> l) << l) == h && ((w >> l) << l) == w, BGAN_E_BADARG,
                     "bgan_swd_feed_f32: %d x %d images do not halve %d times", h, w, levels - 1);
        BGAN_REQUIRE(desc_out[l] && sums[l], BGAN_E_BADARG, "bgan_swd_feed_f32: null output for level %d", l);
    }
    const int64_t count = static_cast<int64_t>(n) * nhoods_per_image;
    const int dlen = c * nhood * nhood;
    const bool fused_gather = levels <= kMaxFeedLevels && dlen <= 32 * 5 && c <= 4 && (nhood & 1) && nhoods_per_image > 0;
    ExtractLevels lv{};
    // Pyramid (:83-88): coarse(l) = pyr_down(g(l)), residual(l) = g(l) - pyr_up(coarse(l)), g(l + 1) = coarse(l).  residual(l)
    // and coarse(l + 1) both need coarse(l) only and share a launch (pyr_pair_kernel) when the quad kernel applies.
    const float* g[16];
    float* resid[16];
    float* coarse[16];
    {
        float* next = static_cast<float*>(ws);
        const float* cur = minibatch;
        for (int l = 0; l < levels; ++l) {
            g[l] = cur;
            resid[l] = coarse[l] = nullptr;
            if (l < levels - 1) {
                const size_t plane = static_cast<size_t>(n) * c * (h >> l) * (w >> l);
                resid[l] = next;
                coarse[l] = next + plane;
                next = coarse[l] + plane / 4;
                cur = coarse[l];
            }
        }
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int planes = n * c;
    if (levels > 1) {
        const int rc = bgan_pyr_down_f32(g[0], coarse[0], planes, h, w, stream);  // :85-86
        if (rc) return rc;
    }
    for (int l = 0; l < levels - 1; ++l) {
        const int hc = (h >> l) / 2, wc = (w >> l) / 2;  // coarse image of level l
        const bool more = l + 1 < levels - 1;           // level l + 1 is reduced as well
        const int64_t up_total = static_cast<int64_t>(planes) * hc * wc;
        const bool pair_ok = more && hc >= 2 && wc >= 2 && reinterpret_cast<uintptr_t>(resid[l]) % 8 == 0 && reinterpret_cast<uintptr_t>(g[l]) % 8 == 0 &&
                             up_total * 4 < (int64_t(1) << 31);
        if (pair_ok) {
            PyrPair a;
            a.up_x = coarse[l];
            a.up_minuend = g[l];
            a.up_y = resid[l];
            a.down_y = coarse[l + 1];
            a.planes = planes;
            a.h = hc;
            a.w = wc;
            a.up_blocks = grid_for(up_total, 256, 16);
            const unsigned down_blocks = grid_for(static_cast<int64_t>(planes) * ((hc + 1) / 2) * ((wc + 1) / 2), 256, 16);
            BGAN_CUDA_OK(launch_pdl(pyr_pair_kernel, dim3(a.up_blocks + down_blocks), dim3(256), 0, st, a));
        } else {
            int rc = bgan_pyr_up_f32(coarse[l], g[l], resid[l], planes, hc, wc, stream);  // :87  residual = g - pyr_up(down)
            if (rc) return rc;
            if (more) {
                rc = bgan_pyr_down_f32(g[l + 1], coarse[l + 1], planes, hc, wc, stream);
                if (rc) return rc;
            }
        }
    }
    int hl = h, wl = w;
    for (int l = 0; l < levels; ++l) {
        const int32_t* px = positions + static_cast<int64_t>(2 * l) * count;
        const int32_t* py = px + count;
        const float* level = l < levels - 1 ? resid[l] : g[l];
        BGAN_REQUIRE(hl >= nhood && wl >= nhood, BGAN_E_BADARG, "bgan_swd_feed_f32: level %d (%d x %d) smaller than the neighbourhood", l, hl, wl);
        if (fused_gather) {
            lv.level[l] = level;
            lv.px[l] = px;
            lv.out[l] = desc_out[l];
            lv.sums[l] = sums[l];
            lv.h[l] = hl;
            lv.w[l] = wl;
        } else {
            const int rc = bgan_swd_extract_f32(level, n, c, hl, wl, px, py, nhoods_per_image, nhood, desc_out[l], sums[l], stream);  // :13-23
            if (rc) return rc;
        }
        hl /= 2;
        wl /= 2;
    }
    if (fused_gather) {  // :13-23 for every level
        const int64_t ctas = std::min<int64_t>(ceil_div64(count, 8), 2 * static_cast<int64_t>(device_sm_count()));
        BGAN_CUDA_OK(launch_pdl(extract_levels_kernel<4, 5>, dim3(static_cast<unsigned>(ctas), levels), dim3(256), 0, static_cast<cudaStream_t>(stream),
                                lv, c, nhoods_per_image, nhood, count));
    }
    return 0;
}

extern "C" int bgan_swd_normalise_f32(float* desc, int64_t rows, int c, int per_channel, const double* sums, double count,
                                      void* stream) {
    BGAN_REQUIRE(desc && sums && rows > 0 && c > 0 && c <= 8 && per_channel > 0 && count > 0, BGAN_E_BADARG,
                 "bgan_swd_normalise_f32: bad arguments");
    const int64_t total = rows * c * per_channel;
    BGAN_REQUIRE(reinterpret_cast<uintptr_t>(desc) % 16 == 0, BGAN_E_BADARG, "bgan_swd_normalise_f32: desc must be 16-byte aligned");
    normalise_kernel<<<grid_for(ceil_div64(total, 4 * kNormUnroll), 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(desc, total, c, per_channel,
                                                                                                                     sums, count);
    BGAN_LAUNCH_OK();
    return 0;
}

namespace bgan {
// project_tc.cu: tcgen05 path (descriptor length <= 192)
bool project_tc_eligible(int64_t rows, int k, int d);
int project_tc_launch(const float* desc, int64_t rows, int k, const float* dirs, int d, float* proj, cudaStream_t stream);
}  // namespace bgan

static thread_local int g_project_path = 0;
extern "C" int bgan_swd_project_last_path(void) { return g_project_path; }

extern "C" int bgan_swd_project_f32(const float* desc, int64_t rows, int k, const float* dirs, int d, float* proj_out, void* stream) {
    BGAN_REQUIRE(desc && dirs && proj_out && rows > 0 && k > 0 && d > 0, BGAN_E_BADARG, "bgan_swd_project_f32: bad arguments");
    if (project_tc_eligible(rows, k, d)) {
        g_project_path = 2;
        return project_tc_launch(desc, rows, k, dirs, d, proj_out, static_cast<cudaStream_t>(stream));
    }
    g_project_path = 1;
    const size_t smem = sizeof(float) * (static_cast<size_t>(kPM) * (k | 1) + static_cast<size_t>(k) * kPN);
    BGAN_REQUIRE(smem <= 227 * 1024, BGAN_E_UNSUPPORTED, "bgan_swd_project_f32: descriptor length %d too large", k);
    BGAN_REQUIRE(ceil_div64(rows, kPM) < (int64_t(1) << 31), BGAN_E_BADARG, "too many rows");
    static thread_local int attr_dev = -1;
    int dev = 0;
    BGAN_CUDA_OK(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        BGAN_CUDA_OK(cudaFuncSetAttribute(project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_dev = dev;
    }
    dim3 grid(static_cast<unsigned>(ceil_div64(rows, kPM)), static_cast<unsigned>(ceil_div(d, kPN)));
    project_kernel<<<grid, kPThreads, smem, static_cast<cudaStream_t>(stream)>>>(desc, rows, k, dirs, d, proj_out);
    BGAN_LAUNCH_OK();
    return 0;
}

extern "C" int bgan_swd_absdiff_sum_f32(const float* a, const float* b, int64_t count, double* sum_out, void* stream) {
    BGAN_REQUIRE(a && b && sum_out && count > 0, BGAN_E_BADARG, "bgan_swd_absdiff_sum_f32: bad arguments");
    const bool vec = reinterpret_cast<uintptr_t>(a) % 16 == 0 && reinterpret_cast<uintptr_t>(b) % 16 == 0;
    absdiff_sum_kernel<<<grid_for(ceil_div64(count, 4 * kDiffUnroll), 256, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, count, vec, sum_out);
    BGAN_LAUNCH_OK();
    return 0;
}

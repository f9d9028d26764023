// Fused separable Gaussian blur for sm_100a.
//
// Replaces the arithmetic of GaussianBlur.forward (reference pytorch/gaussianblur.py:59-71),
// gaussian_blur() (reference gaussian_blur.py:91-132) and their autograd backward: the operator
// is a symmetric banded Toeplitz matrix per axis (zero "same" padding after tap normalisation),
// hence self-adjoint, so backward / double-backward are the same kernel applied to the gradient.
//
// One kernel template covers every case.  A "plane" is a 2-D array of `H` rows by `Wc` floats;
// NCHW: planes = N*C, Wc = W, tap stride ES = 1.  NHWC: planes = N, Wc = W*C, tap stride ES = C.
// A work unit is one output tile (TH rows x TWc floats) of one plane:
//
//   global --TMA box (zero-filled out of bounds) / plain loads--> s_in  [in_rows][in_pitch]
//   row pass   (taps along a row, stride ES)   s_in  -> s_mid [in_rows][mid_pitch]   registers: 16 outputs/thread
//   column pass (taps down a column)           s_mid -> global                      registers: 16x4 outputs/thread
//
// so the intermediate never leaves shared memory.  CTAs are persistent (grid = SMs x resident
// CTAs) and, with TMA, the load of unit i+1 is issued as soon as the row pass of unit i has
// drained s_in, overlapping the column pass.  Taps are computed in-kernel from (rule, std) in the
// reference's own fp32 expression order.  Images whose halo does not fit in shared memory are
// done as two launches of the same kernel (row pass only, column pass only) through a scratch
// plane.  Chunks of taps that fall entirely outside the image are skipped (the padding is zero).
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "taps.cuh"

namespace bgan {

// blur_tc.cu: tcgen05 path for 128 x 128 planes
bool blur_tc_eligible(int planes, int H, int Wc, int ES, int radius, int rule, const float* x, const float* y);
int blur_tc_launch(const MultiIO& io, int planes, int es, int rule, float std_eff, int radius, cudaStream_t stream);

// blur_axis.cu: two-pass tensor-core path for large planes and wide blurs
bool blur_axis_shape_ok(int planes, int H, int W, int ES, int radius);
bool blur_axis_eligible(int planes, int H, int W, int ES, int radius, const float* x, const float* y);
size_t blur_axis_workspace_bytes(int planes, int H, int W);
int blur_axis_launch(const float* x, float* y, int planes, int H, int W, int rule, float std_eff, int radius, const float* taps_dev, void* ws,
                     size_t ws_bytes, cudaStream_t stream);
// blur_direct.cu: FFMA path for 128 x 128 images and radius <= 4
bool blur_direct_eligible(int planes, int H, int Wc, int ES, int radius, const float* x, const float* y);
int blur_direct_launch(const MultiIO& io, int planes, int es, int radius, int rule, float std_eff, const float* taps_dev,
                       cudaStream_t stream);

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kTO = 16;  // outputs per thread along the filtered axis (both passes)
constexpr int kK = 8;    // taps per register chunk

struct BlurArgs {
    const float* x;
    float* y;
    int planes, H, Wc;
    int rh, rv;        // row-pass / column-pass radius (0 = identity pass)
    int rule;          // BGAN_RULE_*
    float std_eff;
    const float* taps_dev;
    // tiling
    int TH, TWc, tiles_y, tiles_x;
    int in_rows, in_pitch, mid_pitch;
    int HL;            // left halo of the row pass in floats (= RP * ES, multiple of 4)
    int ZM;            // zero margin kept outside the image in s_in, floats (multiple of 4)
    int unclipped_w;   // floats the row pass may touch to the right of (xc0 - HL) without clipping
    int nch_h, nch_v;  // tap chunks per pass
    int use_tma, vec_ok;
    int box_rows, nbox;  // TMA: in_rows = box_rows * nbox
    int64_t units;
};

__global__ void taps_kernel(int rule, float s, int r, float* __restrict__ out) {
    extern __shared__ float s_dyn[];
    __shared__ double s_red;
    normalised_taps(rule, s, r, nullptr, s_dyn, &s_red);
    for (int t = threadIdx.x; t < 2 * r + 1; t += blockDim.x) out[t] = s_dyn[t];
}

// ---------------------------------------------------------------------------------------------
// row pass: lane <-> row, each thread 16 consecutive outputs of its row; window via LDS.128
// ---------------------------------------------------------------------------------------------
template <int ES>
__device__ __forceinline__ void row_pass_task(const float* __restrict__ src0, float* __restrict__ dst, int c_lo, int c_hi,
                                              const float* __restrict__ s_wh) {
    constexpr int WN = ((kTO + (kK - 1) * ES) + 3) / 4 * 4;  // window floats per chunk, whole float4s
    float acc[kTO];
#pragma unroll
    for (int o = 0; o < kTO; ++o) acc[o] = 0.0f;
    for (int c = c_lo; c < c_hi; ++c) {
        const float4 wa = *reinterpret_cast<const float4*>(s_wh + c * kK);
        const float4 wb = *reinterpret_cast<const float4*>(s_wh + c * kK + 4);
        const float w[kK] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
        const float* src = src0 + c * (kK * ES);
#pragma unroll
        for (int i4 = 0; i4 < WN / 4; ++i4) {
            const float4 v4 = *reinterpret_cast<const float4*>(src + 4 * i4);
            const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
#pragma unroll
                for (int j = 0; j < kK; ++j) {
                    const int o = 4 * i4 + e - j * ES;
                    if (o >= 0 && o < kTO) acc[o] = fmaf(w[j], v[e], acc[o]);
                }
            }
        }
    }
#pragma unroll
    for (int o4 = 0; o4 < kTO / 4; ++o4)
        *reinterpret_cast<float4*>(dst + 4 * o4) = make_float4(acc[4 * o4], acc[4 * o4 + 1], acc[4 * o4 + 2], acc[4 * o4 + 3]);
}

// ---------------------------------------------------------------------------------------------
// the tile kernel
// ---------------------------------------------------------------------------------------------
template <int ES>
__global__ void __launch_bounds__(kThreads) blur_tile_kernel(const BlurArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // layout: [s_in | s_mid | s_wh | s_wv | s_g | mbar, red]
    float* s_in = reinterpret_cast<float*>(smem_raw);
    float* s_mid = s_in + static_cast<size_t>(a.in_rows) * a.in_pitch;
    float* s_wh = s_mid + static_cast<size_t>(a.in_rows) * a.mid_pitch;
    float* s_wv = s_wh + a.nch_h * kK;
    const int rmax = a.rh > a.rv ? a.rh : a.rv;
    float* s_g = s_wv + a.nch_v * kK;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(s_g + ((2 * rmax + 1 + 3) & ~3));  // 16-byte aligned
    double* s_red = reinterpret_cast<double*>(mbar + 1);

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    // ---- taps --------------------------------------------------------------------------------
    normalised_taps(a.rule, a.std_eff, rmax, a.taps_dev, s_g, s_red);
    {
        const int RP = a.HL / ES;  // row-pass radius padded to a multiple of 4 taps
        for (int t = tid; t < a.nch_h * kK; t += kThreads) {
            const int j = t - (RP - a.rh);
            s_wh[t] = (a.rh == 0) ? (t == 0 ? 1.0f : 0.0f) : ((j >= 0 && j <= 2 * a.rh) ? s_g[j] : 0.0f);
        }
        for (int t = tid; t < a.nch_v * kK; t += kThreads)
            s_wv[t] = (a.rv == 0) ? (t == 0 ? 1.0f : 0.0f) : ((t <= 2 * a.rv) ? s_g[t] : 0.0f);
    }
    if (a.use_tma && tid == 0) {
        tma_prefetch_desc(&tmap);
        mbar_init(mbar, 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int64_t tiles_per_plane = static_cast<int64_t>(a.tiles_y) * a.tiles_x;
    const uint32_t box_bytes = static_cast<uint32_t>(a.in_rows) * a.in_pitch * sizeof(float);

    auto unit_origin = [&](int64_t u, int& plane, int& y0, int& xc0, int& my0, int& my1, int& cx0) {
        plane = static_cast<int>(u / tiles_per_plane);
        const int t = static_cast<int>(u - plane * tiles_per_plane);
        y0 = (t / a.tiles_x) * a.TH;
        xc0 = (t % a.tiles_x) * a.TWc;
        my0 = max(0, y0 - a.rv);
        my1 = min(a.H, y0 + a.TH + a.rv);
        cx0 = max(xc0 - a.HL, -a.ZM);
    };
    auto issue_tma = [&](int64_t u) {
        int plane, y0, xc0, my0, my1, cx0;
        unit_origin(u, plane, y0, xc0, my0, my1, cx0);
        mbar_arrive_expect_tx(mbar, box_bytes);
        for (int b = 0; b < a.nbox; ++b)
            tma_load_3d(s_in + static_cast<size_t>(b) * a.box_rows * a.in_pitch, &tmap, mbar, cx0, my0 + b * a.box_rows, plane);
    };

    uint32_t parity = 0;
    if (a.use_tma && tid == 0 && blockIdx.x < a.units) issue_tma(blockIdx.x);

    for (int64_t u = blockIdx.x; u < a.units; u += gridDim.x) {
        int plane, y0, xc0, my0, my1, cx0;
        unit_origin(u, plane, y0, xc0, my0, my1, cx0);
        const float* __restrict__ xin = a.x + static_cast<size_t>(plane) * a.H * a.Wc;
        float* __restrict__ yout = a.y + static_cast<size_t>(plane) * a.H * a.Wc;
        const int nrows = my1 - my0;

        // ---- stage the input box ---------------------------------------------------------------
        if (a.use_tma) {
            mbar_wait(mbar, parity);
            parity ^= 1;
        } else {
            const int p4 = a.in_pitch >> 2;
            for (int idx = tid; idx < nrows * p4; idx += kThreads) {
                const int rr = idx / p4, c4 = idx - rr * p4;
                const int gx = cx0 + 4 * c4;
                const float* g = xin + static_cast<size_t>(my0 + rr) * a.Wc;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.vec_ok && gx >= 0 && gx + 3 < a.Wc) {
                    v = __ldg(reinterpret_cast<const float4*>(g + gx));
                } else {
                    if (gx >= 0 && gx < a.Wc) v.x = __ldg(g + gx);
                    if (gx + 1 >= 0 && gx + 1 < a.Wc) v.y = __ldg(g + gx + 1);
                    if (gx + 2 >= 0 && gx + 2 < a.Wc) v.z = __ldg(g + gx + 2);
                    if (gx + 3 >= 0 && gx + 3 < a.Wc) v.w = __ldg(g + gx + 3);
                }
                *reinterpret_cast<float4*>(s_in + static_cast<size_t>(rr) * a.in_pitch + 4 * c4) = v;
            }
            __syncthreads();
        }

        // ---- row pass: s_in -> s_mid -------------------------------------------------------------
        {
            const int n_cg = (a.TWc + kTO - 1) / kTO;      // 16-float output groups per tile row
            const int n_rg = (nrows + 31) >> 5;            // 32-row groups
            if (a.rh == 0) {
                // identity: copy the tile's own columns
                const int shift = xc0 - cx0;
                const int m4 = (n_cg * kTO) >> 2;
                for (int idx = tid; idx < nrows * m4; idx += kThreads) {
                    const int rr = idx / m4, c4 = idx - rr * m4;
                    *reinterpret_cast<float4*>(s_mid + static_cast<size_t>(rr) * a.mid_pitch + 4 * c4) =
                        *reinterpret_cast<const float4*>(s_in + static_cast<size_t>(rr) * a.in_pitch + shift + 4 * c4);
                }
            } else {
                for (int task = warp; task < n_cg * n_rg; task += kWarps) {
                    const int cg = task % n_cg, rg = task / n_cg;
                    const int rr = (rg << 5) + lane;
                    const int xc = xc0 + cg * kTO;              // first output column (image coords)
                    if (xc >= a.Wc) continue;
                    // chunk c touches image columns [xc - HL + c*K*ES, ... + WN); keep those that meet [0, Wc)
                    constexpr int WN = ((kTO + (kK - 1) * ES) + 3) / 4 * 4;
                    const int start0 = xc - a.HL;
                    int c_lo = 0, c_hi = a.nch_h;
                    if (start0 + WN <= 0) c_lo = (-(start0 + WN)) / (kK * ES) + 1;
                    if (start0 + (c_hi - 1) * (kK * ES) >= a.Wc) c_hi = (a.Wc - 1 - start0) / (kK * ES) + 1;
                    if (c_hi < c_lo) c_hi = c_lo;
                    if (rr < nrows)
                        row_pass_task<ES>(s_in + static_cast<size_t>(rr) * a.in_pitch + (start0 - cx0),
                                          s_mid + static_cast<size_t>(rr) * a.mid_pitch + cg * kTO, c_lo, c_hi, s_wh);
                }
            }
        }
        __syncthreads();

        // s_in is drained: prefetch the next unit's box while the column pass runs
        if (a.use_tma && tid == 0 && u + gridDim.x < a.units) issue_tma(u + gridDim.x);

        // ---- column pass: s_mid -> global ----------------------------------------------------------
        {
            const int n_cg = (a.TWc + 127) >> 7;            // 128-float column groups (lane <-> 4 floats)
            const int th = min(a.TH, a.H - y0);
            const int n_rg = (th + kTO - 1) / kTO;
            for (int task = warp; task < n_cg * n_rg; task += kWarps) {
                const int cg = task % n_cg, rg = task / n_cg;
                const int col = (cg << 7) + (lane << 2);     // float column inside the tile
                const int ys = y0 + rg * kTO;                // first output row (image coords)
                if (col >= a.TWc || xc0 + col >= a.Wc) continue;
                float4 acc[kTO];
#pragma unroll
                for (int o = 0; o < kTO; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (a.rv == 0) {
#pragma unroll
                    for (int o = 0; o < kTO; ++o)
                        if (ys + o < my1) acc[o] = *reinterpret_cast<const float4*>(s_mid + static_cast<size_t>(ys + o - my0) * a.mid_pitch + col);
                } else {
                    // chunk c touches rows [ys - rv + c*K, ... + TO + K - 2]; keep those that meet [my0, my1)
                    const int start0 = ys - a.rv;
                    int c_lo = 0, c_hi = a.nch_v;
                    if (start0 + (kTO + kK - 1) <= my0) c_lo = (my0 - (start0 + kTO + kK - 1)) / kK + 1;
                    if (start0 + (c_hi - 1) * kK >= my1) c_hi = (my1 - 1 - start0) / kK + 1;
                    for (int c = c_lo; c < c_hi; ++c) {
                        const float4 wa = *reinterpret_cast<const float4*>(s_wv + c * kK);
                        const float4 wb = *reinterpret_cast<const float4*>(s_wv + c * kK + 4);
                        const float w[kK] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
                        const int row0 = start0 + c * kK;
#pragma unroll
                        for (int i = 0; i < kTO + kK - 1; ++i) {
                            const int row = row0 + i;
                            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (row >= my0 && row < my1)
                                v = *reinterpret_cast<const float4*>(s_mid + static_cast<size_t>(row - my0) * a.mid_pitch + col);
#pragma unroll
                            for (int j = 0; j < kK; ++j) {
                                const int o = i - j;
                                if (o >= 0 && o < kTO) {
                                    acc[o].x = fmaf(w[j], v.x, acc[o].x);
                                    acc[o].y = fmaf(w[j], v.y, acc[o].y);
                                    acc[o].z = fmaf(w[j], v.z, acc[o].z);
                                    acc[o].w = fmaf(w[j], v.w, acc[o].w);
                                }
                            }
                        }
                    }
                }
                const int gx = xc0 + col;
#pragma unroll
                for (int o = 0; o < kTO; ++o) {
                    const int gy = ys + o;
                    if (gy >= y0 + th) break;
                    float* g = yout + static_cast<size_t>(gy) * a.Wc + gx;
                    if (a.vec_ok && gx + 3 < a.Wc) {
                        __stcs(reinterpret_cast<float4*>(g), acc[o]);
                    } else {
                        g[0] = acc[o].x;
                        if (gx + 1 < a.Wc) g[1] = acc[o].y;
                        if (gx + 2 < a.Wc) g[2] = acc[o].z;
                        if (gx + 3 < a.Wc) g[3] = acc[o].w;
                    }
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// host: tiling plan
// ---------------------------------------------------------------------------------------------
struct Plan {
    BlurArgs a;
    size_t smem;
    double cost;
    bool ok;
};

constexpr size_t kSmemBudget = 227 * 1024;

Plan make_plan(int H, int Wc, int ES, int rh, int rv, int TH, int TWc) {
    Plan p{};
    BlurArgs& a = p.a;
    a.H = H;
    a.Wc = Wc;
    a.rh = rh;
    a.rv = rv;
    a.TH = TH;
    a.TWc = TWc;
    a.tiles_y = ceil_div(H, TH);
    a.tiles_x = ceil_div(Wc, TWc);
    const int RP = round_up(rh, 4);
    a.HL = RP * ES;
    const int WN = round_up(kTO + (kK - 1) * ES, 4);
    a.ZM = round_up(WN, 4);
    a.nch_h = ceil_div(2 * RP + 1, kK);
    a.nch_v = ceil_div(2 * rv + 1, kK);
    const int twc_r = round_up(TWc, kTO);
    a.unclipped_w = twc_r - kTO + (a.nch_h - 1) * kK * ES + WN;
    int width = 0;
    for (int tx = 0; tx < a.tiles_x; ++tx) {
        const int xc0 = tx * TWc;
        const int cx0 = std::max(xc0 - a.HL, -a.ZM);
        const int cx1 = std::min(xc0 - a.HL + a.unclipped_w, Wc + a.ZM);
        width = std::max(width, cx1 - cx0);
    }
    width = std::max(width, twc_r + std::min(a.HL, a.ZM));  // identity row pass copies twc_r floats from (xc0 - cx0)
    a.in_pitch = round_up(width, 4);
    if ((a.in_pitch / 4) % 2 == 0) a.in_pitch += 4;  // pitch/4 odd: conflict-free LDS.128 with lane <-> row
    a.mid_pitch = twc_r;
    if ((a.mid_pitch / 4) % 2 == 0) a.mid_pitch += 4;
    a.in_rows = std::min(H, TH + 2 * rv);
    a.nbox = ceil_div(a.in_rows, 256);
    a.box_rows = ceil_div(a.in_rows, a.nbox);
    a.in_rows = a.box_rows * a.nbox;
    const int rmax = std::max(rh, rv);
    p.smem = sizeof(float) * (static_cast<size_t>(a.in_rows) * (a.in_pitch + a.mid_pitch) + (a.nch_h + a.nch_v) * kK +
                              round_up(2 * rmax + 1, 4)) + 32;
    p.ok = p.smem <= kSmemBudget;
    // relative cost: FMAs per output pixel (row pass is inflated by the column halo) ...
    const double infl = static_cast<double>(std::min(H, TH + 2 * rv)) / std::min(H, TH);
    p.cost = infl * (rh ? a.nch_h * kK : 1) + (rv ? a.nch_v * kK : 1);
    // ... and a mild preference for two or more resident CTAs (load/compute overlap)
    if (p.smem > kSmemBudget / 2) p.cost *= 1.15;
    // lanes idle in the row pass when a tile has few rows
    const int rows = std::min(H, TH + 2 * rv);
    p.cost *= static_cast<double>(round_up(rows, 32)) / rows * 0.25 + 0.75;
    return p;
}

Plan choose_plan(int H, int Wc, int ES, int rh, int rv) {
    const char* eth = getenv("BGAN_BLUR_TH");
    const char* etw = getenv("BGAN_BLUR_TW");
    int ths[] = {H, 256, 128, 96, 64, 48, 32, 16};
    int tws[] = {round_up(Wc, 4), 512, 384, 256, 192, 128, 96, 64, 48, 32};
    Plan best{};
    best.ok = false;
    for (int th : ths) {
        if (th > H || (th != H && th % kTO)) continue;
        if (eth && atoi(eth) > 0 && th != std::min(H, atoi(eth))) continue;
        for (int tw : tws) {
            if (tw > round_up(Wc, 4) || (tw != round_up(Wc, 4) && tw % kTO)) continue;
            if (etw && atoi(etw) > 0 && tw != std::min(round_up(Wc, 4), atoi(etw))) continue;
            Plan p = make_plan(H, Wc, ES, rh, rv, th, tw);
            if (!p.ok) continue;
            if (!best.ok || p.cost < best.cost) best = p;
        }
    }
    return best;
}

// ---------------------------------------------------------------------------------------------
// host: TMA descriptor
// ---------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

bool make_tmap(CUtensorMap* tm, const float* x, int planes, int H, int Wc, int box_w, int box_h) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    thread_local bool bound = false;  // the driver call needs a context bound to this thread (autograd worker threads)
    if (!bound) {
        cudaFree(nullptr);
        bound = true;
    }
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(Wc), static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(planes)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(Wc) * 4, static_cast<cuuint64_t>(Wc) * H * 4};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

thread_local int g_last_path = 0;

template <int ES>
int launch_tiles(Plan& p, const float* x, float* y, int planes, int rule, float std_eff, const float* taps_dev,
                 cudaStream_t stream, bool* used_tma) {
    BlurArgs& a = p.a;
    a.x = x;
    a.y = y;
    a.planes = planes;
    a.rule = rule;
    a.std_eff = std_eff;
    a.taps_dev = taps_dev;
    a.units = static_cast<int64_t>(planes) * a.tiles_y * a.tiles_x;
    a.vec_ok = (a.Wc % 4 == 0) && (reinterpret_cast<uintptr_t>(x) % 16 == 0) && (reinterpret_cast<uintptr_t>(y) % 16 == 0);
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    const char* no_tma = getenv("BGAN_BLUR_NO_TMA");
    a.use_tma = 0;
    if (!(no_tma && atoi(no_tma)) && a.Wc % 4 == 0 && reinterpret_cast<uintptr_t>(x) % 16 == 0 && a.in_pitch <= 256 &&
        a.box_rows <= 256 && (a.nbox == 1 || (static_cast<size_t>(a.box_rows) * a.in_pitch * 4) % 128 == 0))
        a.use_tma = make_tmap(&tmap, x, planes, a.H, a.Wc, a.in_pitch, a.box_rows) ? 1 : 0;
    *used_tma = a.use_tma != 0;

    auto kern = blur_tile_kernel<ES>;
    // per-thread caches: the opt-in smem attribute and the occupancy of the last smem size
    thread_local int attr_dev = -1;
    thread_local size_t occ_smem = 0;
    thread_local int occ_val = 0;
    int dev = 0;
    BGAN_CUDA_OK(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        BGAN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemBudget)));
        attr_dev = dev;
        occ_smem = 0;
    }
    if (occ_smem != p.smem) {
        BGAN_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_val, kern, kThreads, p.smem));
        occ_smem = p.smem;
    }
    const int per_sm = occ_val;
    BGAN_REQUIRE(per_sm >= 1, BGAN_E_UNSUPPORTED, "blur tile kernel does not fit on an SM (smem %zu)", p.smem);
    const int64_t resident = static_cast<int64_t>(device_sm_count()) * per_sm;
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(a.units, resident));
    kern<<<grid, kThreads, p.smem, stream>>>(a, tmap);
    BGAN_LAUNCH_OK();
    return 0;
}

int launch_any(Plan& p, int ES, const float* x, float* y, int planes, int rule, float std_eff, const float* taps_dev,
               cudaStream_t stream, bool* used_tma) {
    if (ES == 1) return launch_tiles<1>(p, x, y, planes, rule, std_eff, taps_dev, stream, used_tma);
    if (ES == 3) return launch_tiles<3>(p, x, y, planes, rule, std_eff, taps_dev, stream, used_tma);
    set_error("unsupported interleave %d (NHWC needs C in {1,3})", ES);
    return BGAN_E_UNSUPPORTED;
}

struct Geometry {
    int planes, H, Wc, ES;
};

bool geometry(int n, int c, int h, int w, int layout, Geometry* g) {
    if (n <= 0 || c <= 0 || h <= 0 || w <= 0) return false;
    if (layout == BGAN_LAYOUT_NCHW) {
        *g = {n * c, h, w, 1};
        return true;
    }
    if (layout == BGAN_LAYOUT_NHWC) {
        if (c != 1 && c != 3) return false;
        *g = {n, h, w * c, c};
        return true;
    }
    return false;
}

// NHWC with C in 2 .. 4 whose channel planes qualify for the two-pass tensor-core kernel: planar copy + that kernel's scratch
bool nhwc_via_planar_ok(const Geometry& g, int radius) {
    if (g.ES < 2 || g.ES > 4) return false;
    const int W = g.Wc / g.ES;
    return static_cast<int64_t>(g.planes) * g.ES < (int64_t(1) << 31) && blur_axis_shape_ok(g.planes * g.ES, g.H, W, 1, radius);
}
size_t nhwc_via_planar_bytes(const Geometry& g) {
    const size_t planar = (static_cast<size_t>(g.planes) * g.H * g.Wc * sizeof(float) + 255) & ~size_t(255);
    return planar + blur_axis_workspace_bytes(g.planes * g.ES, g.H, g.Wc / g.ES);
}


}  // namespace
}  // namespace bgan

using namespace bgan;

extern "C" int bgan_blur_plan(int rule, float std, int height, int width, int* kernel_size, int* radius, float* std_eff) {
    BGAN_REQUIRE(height > 0 && width > 0 && std > 0.0f, BGAN_E_BADARG, "bgan_blur_plan: bad arguments");
    int k;
    float s = std;
    if (rule == BGAN_RULE_TORCH) {
        // pytorch/gaussianblur.py:17-30,43-57: odd_up(ceil(5 std)) >= 3, capped at odd_up(min(H, W)) - 2
        auto odd_up = [](float v) {
            const int n = static_cast<int>(ceilf(v));
            return (n % 2 == 1) ? n : n + 1;
        };
        k = std::max(odd_up(5.0f * std), 3);
        k = std::min(k, odd_up(static_cast<float>(std::min(height, width))) - 2);
        BGAN_REQUIRE(k >= 1, BGAN_E_BADARG, "kernel size should be odd (image smaller than 3 pixels)");
    } else if (rule == BGAN_RULE_TF) {
        // gaussian_blur.py:21-31,65-72: k = floor(6 std) + 1 in fp32, clipped to [3, max(H, W)], std' = (k - 1) / 6
        float kf = floorf((6.0f * std) * 2.0f / 2.0f) + 1.0f;
        kf = std::min(std::max(kf, 3.0f), static_cast<float>(std::max(height, width)));
        s = std::max((kf - 1.0f) / 6.0f, 0.01f);
        k = static_cast<int>(kf);
    } else {
        set_error("bgan_blur_plan: unknown rule %d", rule);
        return BGAN_E_BADARG;
    }
    if (kernel_size) *kernel_size = k;
    if (radius) *radius = k / 2;
    if (std_eff) *std_eff = s;
    return 0;
}

extern "C" int bgan_blur_taps_f32(int rule, float std_eff, int radius, float* taps_out_device, void* stream) {
    BGAN_REQUIRE(radius >= 0 && taps_out_device && (rule == BGAN_RULE_TORCH || rule == BGAN_RULE_TF), BGAN_E_BADARG,
                 "bgan_blur_taps_f32: bad arguments");
    taps_kernel<<<1, 256, (2 * radius + 1) * sizeof(float), static_cast<cudaStream_t>(stream)>>>(rule, std_eff, radius,
                                                                                                 taps_out_device);
    BGAN_LAUNCH_OK();
    return 0;
}

extern "C" size_t bgan_blur2d_workspace_bytes(int n, int c, int height, int width, int layout, int radius) {
    Geometry g;
    if (!geometry(n, c, height, width, layout, &g) || radius < 0) return 0;
    if (blur_axis_shape_ok(g.planes, g.H, g.Wc, g.ES, radius)) return blur_axis_workspace_bytes(g.planes, g.H, g.Wc);
    if (nhwc_via_planar_ok(g, radius)) return nhwc_via_planar_bytes(g);
    Plan fused = choose_plan(g.H, g.Wc, g.ES, radius, radius);
    if (fused.ok) return 0;
    return static_cast<size_t>(g.planes) * g.H * g.Wc * sizeof(float);
}

extern "C" int bgan_blur2d_last_path(void) { return g_last_path; }

namespace bgan {
namespace {

// out[i][:] = u[i][:] + a[i] * (v[i][:] - u[i][:]): the interpolates of the gradient penalty (wgan.py:237-239) taken
// between two *blurred* batches -- the blur is linear, so B(x_hat) = B(reals) + a (B(fakes) - B(reals)) and the third
// forward blur of the critic step is this HBM-bound elementwise pass instead (12 bytes per element).
__global__ void __launch_bounds__(256) lerp_rows_kernel(const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ a,
                                                        float* __restrict__ out, int64_t per_row4, int64_t total4) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total4; i += stride) {
        const float w = __ldg(a + i / per_row4);
        const float4 p = __ldcs(reinterpret_cast<const float4*>(u) + i), q = __ldcs(reinterpret_cast<const float4*>(v) + i);
        __stcs(reinterpret_cast<float4*>(out) + i,
               make_float4(p.x + w * (q.x - p.x), p.y + w * (q.y - p.y), p.z + w * (q.z - p.z), p.w + w * (q.w - p.w)));
    }
}
__global__ void __launch_bounds__(256) lerp_rows_scalar_kernel(const float* __restrict__ u, const float* __restrict__ v,
                                                               const float* __restrict__ a, float* __restrict__ out, int64_t per_row,
                                                               int64_t total) {
    const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += stride) {
        const float w = __ldg(a + i / per_row), p = u[i];
        out[i] = p + w * (v[i] - p);
    }
}

// Layout conversion for large NHWC images (BASELINE configs[3] in the reference's own layout, gaussian_blur.py:111-130): the
// two-pass tensor-core kernel needs planes that are contiguous along the blurred axis, and a channel-interleaved row cannot
// be staged by TMA one channel at a time (12-byte element stride).  Interleaved <-> planar through shared memory, both sides
// coalesced: chunk of kCvtPix consecutive pixels of one image per CTA iteration; the interleaved side moves as float4, the
// smem side is read / written with stride C words across lanes (conflict-free for C = 3).
constexpr int kCvtPix = 1024, kCvtThreads = 256;
template <bool TO_PLANAR>
__global__ void __launch_bounds__(kCvtThreads) layout_convert_kernel(const float* __restrict__ src, float* __restrict__ dst, int64_t images,
                                                                     int64_t hw, int C) {
    __shared__ __align__(16) float s_buf[kCvtPix * 4];
    const int64_t chunks_per_image = (hw + kCvtPix - 1) / kCvtPix;
    const int64_t chunks = images * chunks_per_image;
    for (int64_t ch = blockIdx.x; ch < chunks; ch += gridDim.x) {
        const int64_t img = ch / chunks_per_image;
        const int64_t p0 = (ch - img * chunks_per_image) * kCvtPix;
        const int npix = static_cast<int>(hw - p0 < kCvtPix ? hw - p0 : kCvtPix);
        const float* inter_c = (TO_PLANAR ? src : dst) + (img * hw + p0) * C;   // interleaved side of this chunk
        const int nfl = npix * C;
        const bool vec = (nfl & 3) == 0 && (reinterpret_cast<uintptr_t>(inter_c) & 15) == 0;
        if (TO_PLANAR) {
            if (vec) {
                for (int i = threadIdx.x; i < nfl / 4; i += kCvtThreads)
                    reinterpret_cast<float4*>(s_buf)[i] = __ldcs(reinterpret_cast<const float4*>(inter_c) + i);
            } else {
                for (int i = threadIdx.x; i < nfl; i += kCvtThreads) s_buf[i] = __ldcs(inter_c + i);
            }
            __syncthreads();
            for (int c = 0; c < C; ++c) {
                float* plane = dst + (img * C + c) * hw + p0;
                for (int q = threadIdx.x; q < npix; q += kCvtThreads) plane[q] = s_buf[q * C + c];
            }
        } else {
            for (int c = 0; c < C; ++c) {
                const float* plane = src + (img * C + c) * hw + p0;
                for (int q = threadIdx.x; q < npix; q += kCvtThreads) s_buf[q * C + c] = __ldcs(plane + q);
            }
            __syncthreads();
            float* inter = dst + (img * hw + p0) * C;
            if (vec) {
                for (int i = threadIdx.x; i < nfl / 4; i += kCvtThreads) __stcs(reinterpret_cast<float4*>(inter) + i, reinterpret_cast<const float4*>(s_buf)[i]);
            } else {
                for (int i = threadIdx.x; i < nfl; i += kCvtThreads) inter[i] = s_buf[i];
            }
        }
        __syncthreads();
    }
}

int layout_convert(bool to_planar, const float* src, float* dst, int64_t images, int64_t hw, int C, cudaStream_t stream) {
    const int64_t chunks = images * ((hw + kCvtPix - 1) / kCvtPix);
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(chunks, static_cast<int64_t>(device_sm_count()) * 8));
    if (to_planar)
        layout_convert_kernel<true><<<grid, kCvtThreads, 0, stream>>>(src, dst, images, hw, C);
    else
        layout_convert_kernel<false><<<grid, kCvtThreads, 0, stream>>>(src, dst, images, hw, C);
    BGAN_LAUNCH_OK();
    return 0;
}

// One tensor (io.count == 1) or several equally shaped ones.  The 128 x 128 kernels take all tensors in one launch;
// every other shape falls back to one launch sequence per tensor (same stream, so the scratch is reused).
int blur2d_dispatch(const MultiIO& io, const Geometry& g, int rule, float std_eff, int radius, const float* taps_device, void* ws,
                    size_t ws_bytes, cudaStream_t stream) {
    bool tma = false;
    bool tc_ok = true, direct_ok = true, axis_ok = true;
    for (int t = 0; t < io.count; ++t) {
        tc_ok = tc_ok && blur_tc_eligible(g.planes, g.H, g.Wc, g.ES, radius, rule, io.x[t], io.y[t]);
        direct_ok = direct_ok && blur_direct_eligible(g.planes, g.H, g.Wc, g.ES, radius, io.x[t], io.y[t]);
        axis_ok = axis_ok && blur_axis_eligible(g.planes, g.H, g.Wc, g.ES, radius, io.x[t], io.y[t]);
    }
    if (axis_ok && ws != nullptr && ws_bytes >= blur_axis_workspace_bytes(g.planes, g.H, g.Wc) && reinterpret_cast<uintptr_t>(ws) % 16 == 0) {
        g_last_path = 6;
        for (int t = 0; t < io.count; ++t) {
            const int rc = blur_axis_launch(io.x[t], io.y[t], g.planes, g.H, g.Wc, rule, std_eff, radius, taps_device, ws, ws_bytes, stream);
            if (rc) return rc;
        }
        return 0;
    }
    // large NHWC images: interleaved -> planar, the two tensor-core passes on the planes (in place in the planar copy: the
    // second pass no longer reads it), planar -> interleaved.  Two extra HBM passes (0.25 ms at configs[3]) instead of the
    // CUDA-core tile kernel's 770 FMAs per pixel (12.6 - 24.9 ms).
    if (nhwc_via_planar_ok(g, radius) && ws != nullptr && ws_bytes >= nhwc_via_planar_bytes(g) && reinterpret_cast<uintptr_t>(ws) % 256 == 0) {
        const int W = g.Wc / g.ES, planes = g.planes * g.ES;
        const size_t planar_bytes = (static_cast<size_t>(g.planes) * g.H * g.Wc * sizeof(float) + 255) & ~size_t(255);
        float* planar = static_cast<float*>(ws);
        void* axis_ws = static_cast<char*>(ws) + planar_bytes;
        bool aligned = true;
        for (int t = 0; t < io.count; ++t)
            aligned = aligned && reinterpret_cast<uintptr_t>(io.x[t]) % 16 == 0 && reinterpret_cast<uintptr_t>(io.y[t]) % 16 == 0;
        if (aligned && blur_axis_eligible(planes, g.H, W, 1, radius, planar, planar)) {
            g_last_path = 7;
            for (int t = 0; t < io.count; ++t) {
                int rc = layout_convert(true, io.x[t], planar, g.planes, static_cast<int64_t>(g.H) * W, g.ES, stream);
                if (rc) return rc;
                rc = blur_axis_launch(planar, planar, planes, g.H, W, rule, std_eff, radius, taps_device, axis_ws, ws_bytes - planar_bytes, stream);
                if (rc) return rc;
                rc = layout_convert(false, planar, io.y[t], g.planes, static_cast<int64_t>(g.H) * W, g.ES, stream);
                if (rc) return rc;
            }
            return 0;
        }
    }
    // the unit counters of the one-launch kernels are 32-bit
    const bool one_launch = static_cast<int64_t>(g.planes) * g.ES * io.count * 2 < (int64_t(1) << 31);
    // 128 x 128 planes: the FFMA kernel up to radius 5 -- HBM-bound up to 4 (18.8 - 19.6 us per 768 planes), 22.7 us at 5, 24.6 at 6,
    // 28.9 at 7 against 25 - 26.7 for the tensor-core kernel (plain stream; round 2, after the taps became uniform operands and
    // the planes half-plane units: csrc/blur_direct.cu).  Radii 6 .. 8 only for explicit taps (which the tensor-core kernel does
    // not take).
    static const int direct_cut = [] {  // first radius that goes to the tensor-core kernel when both kernels take the call
        const char* e = getenv("BGAN_BLUR_DIRECT_CUT");
        return e ? atoi(e) : 6;
    }();
    if (one_launch && direct_ok && !(tc_ok && radius >= direct_cut)) {
        g_last_path = 5;
        return blur_direct_launch(io, g.planes, g.ES, radius, rule, std_eff, taps_device, stream);
    }
    if (one_launch && tc_ok) {
        g_last_path = 4;
        return blur_tc_launch(io, g.planes, g.ES, rule, std_eff, radius, stream);
    }
    Plan fused = choose_plan(g.H, g.Wc, g.ES, radius, radius);
    if (fused.ok) {
        for (int t = 0; t < io.count; ++t) {
            const int rc = launch_any(fused, g.ES, io.x[t], io.y[t], g.planes, rule, std_eff, taps_device, stream, &tma);
            if (rc) return rc;
        }
        g_last_path = tma ? 1 : 2;
        return 0;
    }
    // halo too large for one tile: row pass and column pass as two launches through scratch
    const size_t need = static_cast<size_t>(g.planes) * g.H * g.Wc * sizeof(float);
    BGAN_REQUIRE(ws && ws_bytes >= need, BGAN_E_WORKSPACE, "bgan_blur2d_f32: workspace of %zu bytes required, got %zu", need, ws_bytes);
    BGAN_REQUIRE(reinterpret_cast<uintptr_t>(ws) % 16 == 0, BGAN_E_WORKSPACE, "workspace must be 16-byte aligned");
    Plan rowp = choose_plan(g.H, g.Wc, g.ES, radius, 0);
    Plan colp = choose_plan(g.H, g.Wc, g.ES, 0, radius);
    BGAN_REQUIRE(rowp.ok && colp.ok, BGAN_E_UNSUPPORTED, "bgan_blur2d_f32: radius %d too large for shared memory", radius);
    float* mid = static_cast<float*>(ws);
    for (int t = 0; t < io.count; ++t) {
        int rc = launch_any(rowp, g.ES, io.x[t], mid, g.planes, rule, std_eff, taps_device, stream, &tma);
        if (rc) return rc;
        rc = launch_any(colp, g.ES, mid, io.y[t], g.planes, rule, std_eff, taps_device, stream, &tma);
        if (rc) return rc;
    }
    g_last_path = 3;
    return 0;
}

int blur2d_check(const char* who, const Geometry& g, int rule, float std_eff, int radius, const float* taps_device) {
    BGAN_REQUIRE(radius >= 0, BGAN_E_BADARG, "%s: negative radius", who);
    BGAN_REQUIRE(rule == BGAN_RULE_TORCH || rule == BGAN_RULE_TF || (rule == BGAN_RULE_EXPLICIT && taps_device), BGAN_E_BADARG,
                 "%s: unknown rule %d or missing taps", who, rule);
    BGAN_REQUIRE(rule == BGAN_RULE_EXPLICIT || std_eff > 0.0f, BGAN_E_BADARG, "%s: std must be positive", who);
    BGAN_REQUIRE(static_cast<int64_t>(g.planes) * g.H * g.Wc < (int64_t(1) << 40), BGAN_E_BADARG, "tensor too large");
    return 0;
}

}  // namespace
}  // namespace bgan

extern "C" int bgan_blur2d_f32(const float* x, float* y, int n, int c, int height, int width, int layout, int rule,
                               float std_eff, int radius, const float* taps_device, void* ws, size_t ws_bytes, void* stream_) {
    Geometry g;
    BGAN_REQUIRE(geometry(n, c, height, width, layout, &g), BGAN_E_BADARG,
                 "bgan_blur2d_f32: bad shape/layout (n=%d c=%d h=%d w=%d layout=%d)", n, c, height, width, layout);
    BGAN_REQUIRE(x && y && x != y, BGAN_E_BADARG, "bgan_blur2d_f32: x and y must be distinct non-null device pointers");
    if (const int rc = blur2d_check("bgan_blur2d_f32", g, rule, std_eff, radius, taps_device)) return rc;
    MultiIO io{};
    io.x[0] = x;
    io.y[0] = y;
    io.count = 1;
    return blur2d_dispatch(io, g, rule, std_eff, radius, taps_device, ws, ws_bytes, static_cast<cudaStream_t>(stream_));
}

extern "C" int bgan_blur2d_multi_f32(const float* const* xs, float* const* ys, int count, int n, int c, int height, int width, int layout,
                                     int rule, float std_eff, int radius, const float* taps_device, void* ws, size_t ws_bytes,
                                     void* stream_) {
    Geometry g;
    BGAN_REQUIRE(geometry(n, c, height, width, layout, &g), BGAN_E_BADARG,
                 "bgan_blur2d_multi_f32: bad shape/layout (n=%d c=%d h=%d w=%d layout=%d)", n, c, height, width, layout);
    BGAN_REQUIRE(xs && ys && count >= 1 && count <= kMaxMulti, BGAN_E_BADARG, "bgan_blur2d_multi_f32: 1 .. %d tensors per call, got %d",
                 kMaxMulti, count);
    if (const int rc = blur2d_check("bgan_blur2d_multi_f32", g, rule, std_eff, radius, taps_device)) return rc;
    MultiIO io{};
    io.count = count;
    for (int t = 0; t < count; ++t) {
        io.x[t] = xs[t];
        io.y[t] = ys[t];
        BGAN_REQUIRE(xs[t] && ys[t], BGAN_E_BADARG, "bgan_blur2d_multi_f32: null pointer for tensor %d", t);
        for (int u = 0; u < count; ++u)
            BGAN_REQUIRE(ys[t] != xs[u] && (u == t || ys[t] != ys[u]), BGAN_E_BADARG,
                         "bgan_blur2d_multi_f32: output %d aliases another tensor of the call", t);
    }
    return blur2d_dispatch(io, g, rule, std_eff, radius, taps_device, ws, ws_bytes, static_cast<cudaStream_t>(stream_));
}

extern "C" int bgan_lerp_rows_f32(const float* u, const float* v, const float* weights, float* out, int64_t rows, int64_t per_row,
                                  void* stream_) {
    BGAN_REQUIRE(u && v && weights && out && rows >= 0 && per_row >= 1, BGAN_E_BADARG, "bgan_lerp_rows_f32: bad arguments");
    if (rows == 0) return 0;
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int64_t total = rows * per_row;
    const bool vec = per_row % 4 == 0 && (reinterpret_cast<uintptr_t>(u) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(out)) % 16 == 0;
    const int64_t work = vec ? total / 4 : total;
    const unsigned grid = static_cast<unsigned>(std::min<int64_t>(ceil_div64(work, 256), static_cast<int64_t>(device_sm_count()) * 8));
    if (vec)
        lerp_rows_kernel<<<grid, 256, 0, stream>>>(u, v, weights, out, per_row / 4, work);
    else
        lerp_rows_scalar_kernel<<<grid, 256, 0, stream>>>(u, v, weights, out, per_row, work);
    BGAN_LAUNCH_OK();
    return 0;
}

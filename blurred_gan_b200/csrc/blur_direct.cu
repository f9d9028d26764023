// Direct (CUDA-core) separable blur for 128 x 128 images and small radii (r <= 8 NCHW, r <= 4 NHWC).
//
// The tensor-core kernel (blur_tc.cu) costs the same ~26 us per 768-plane launch for every tap count; for the few-tap
// blurs at the end of the sigma annealing (the reference's real operating range is sigma <= 5, SURVEY Q12, and its
// default initial sigma of 0.05 gives 3 taps) a plain FFMA kernel is cheaper than two 128^3 GEMMs: 2 (2r + 1) FMAs
// per pixel is far below the HBM time for r <= 4, so this kernel is a copy with a little arithmetic on the way.
//
// Per CTA (persistent, one per SM): a producer warp streams units -- a whole NCHW plane, or half an NHWC image --
// into a ring of shared-memory buffers with TMA.  The box starts r rows above the unit and ends r rows below it, and
// TMA zero-fills what lies outside the image, which is exactly the reference's zero "same" padding in the vertical
// direction.  Eight consumer warps (NCHW: two groups of four, each group on its own plane) each own a band of rows; a
// lane owns four pixels (4 * ES consecutive floats) of every row of the band.  For each input row the lane loads its
// window -- NCHW: aligned LDS.128 from rows staged with zero columns either side (TMA again); NHWC: its own floats plus
// the 3 r floats either side from its neighbour lanes by shuffle, zero at the image border -- forms the horizontally
// blurred row in registers, and as soon as 2r + 1 of those exist, the vertically blurred output row, stored with
// 16-byte coalesced streaming stores.  The row loop is fully unrolled, so the sliding window of 2r + 1 blurred rows lives in
// registers under static names; the intermediate never touches shared memory.
#include <cuda.h>

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "taps.cuh"

namespace bgan {
namespace {

constexpr int kD = 128;            // image edge
constexpr int kConsumerWarps = 8;
constexpr int kDirectThreads = (kConsumerWarps + 1) * 32;

// ES == 1 (NCHW planes, NHWC with one channel): the staged rows carry `pad` zero floats on either side (TMA fills what
// lies outside the image), so a lane reads its whole window with aligned LDS.128 and no border logic.
// ES == 3 (NHWC): a lane owns 12 floats and gets the 3 r floats either side from its neighbour lanes by shuffle.
// HALF (NCHW, r > 4): units of half a plane.  A whole plane with its 2 r halo rows and 8 pad floats either side is ~80 KB:
// only two of those fit, one per warp group, so nothing was in flight while a group worked and every plane paid its full load
// latency (26.8 us per 768 planes at r = 5 against 21.0 at r = 4, whatever the FMAs cost).  Half planes are ~43 KB: five stages,
// three of them in flight.
template <int ES, bool HALF = false>
struct DirectCfg {
    static constexpr int kRowFloats = kD * ES;                 // floats per image row
    static constexpr int kGroups = ES == 1 ? 2 : 1;            // consumer warp groups, each working on its own unit
    static constexpr int kGroupWarps = kConsumerWarps / kGroups;
    static constexpr int kRowsPerWarp = ES == 1 ? (HALF ? 16 : 32) : 8;  // output rows owned by one consumer warp (the r rows above
                                                               // and below are blurred horizontally a second time by the
                                                               // neighbour warp: taller bands, less redundancy)
    static constexpr int kUnitRows = kGroupWarps * kRowsPerWarp;  // 128 (whole plane) / 64 (half an NHWC image)
    static constexpr int kUnitsPerImage = kD / kUnitRows;      // 1 / 2
    static constexpr int kPanels = ES == 1 ? 1 : 2;            // TMA boxes span at most 256 elements: NHWC rows are two boxes
    static constexpr int kPanelFloats = kRowFloats / kPanels;  // 128 / 192
    static constexpr int kLanesPerPanel = 32 / kPanels;
    __host__ __device__ static constexpr int pad(int r) { return ES == 1 ? (r <= 4 ? 4 : 8) : 0; }
    __host__ __device__ static constexpr int pitch(int r) { return kPanelFloats + 2 * pad(r); }  // floats per staged (panel) row
    __host__ __device__ static constexpr int stages(int r) { return ES == 1 ? (HALF ? 5 : (r <= 4 ? 3 : 2)) : 2; }
    __host__ __device__ static constexpr int box_rows(int r) { return kUnitRows + 2 * r; }
    __host__ __device__ static constexpr size_t panel_bytes(int r) { return (static_cast<size_t>(box_rows(r)) * pitch(r) * sizeof(float) + 127) / 128 * 128; }
    __host__ __device__ static constexpr size_t stage_bytes(int r) { return panel_bytes(r) * kPanels; }  // (TMA destinations: 128-byte aligned)
    __host__ __device__ static constexpr size_t smem_bytes(int r) { return stages(r) * stage_bytes(r) + 1024; }
};

// Packed fp32 arithmetic (sm_100: FFMA2 / FMUL2, two IEEE fp32 operations per lane and instruction -- the scalar FFMA issues at
// half the rate the FP32 pipe can retire, so a blur of more than ~9 taps was bound by its FMAs, not by HBM).  Each half is the
// same correctly rounded fma / mul as the scalar form: results are bit-identical to the scalar kernel.
using f32x2 = unsigned long long;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo_of(f32x2 v) { return __uint_as_float(static_cast<unsigned>(v)); }
__device__ __forceinline__ float hi_of(f32x2 v) { return __uint_as_float(static_cast<unsigned>(v >> 32)); }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

template <int NT>
struct DirectArgs {
    float* y[NT];  // one output per tensor of a multi-tensor launch (bgan_blur2d_multi_f32)
    const float* taps_dev;
    float taps[17];        // PT kernels: the normalised taps, computed on the host (see host_taps)
    int units;             // all tensors together: planes (NCHW) or 2 * images (NHWC), tensor after tensor
    int units_per_tensor;
    int rule;
    float std_eff;
};
template <int NT>
struct DirectMaps {
    CUtensorMap m[NT];
};

// PT: the taps arrive as kernel parameters.  A tap is then an operand from the constant bank / a uniform register instead of a
// per-thread register, and that decides the FMA rate (tools/ffma_rate.cu on B200, FMA per clock and SM: FFMA with three register
// operands 75, with a uniform operand 108, FFMA2 with register operands 84, FFMA2 with a uniform operand 123): with the taps read
// from shared memory into registers the kernel was bound by its FMAs from r = 5 on.  Explicit taps (device memory) keep the
// register form.
template <int ES, int R, int NT, bool PT>
__global__ void __launch_bounds__(kDirectThreads, 1) blur_direct128_kernel(const DirectArgs<NT> a, const __grid_constant__ DirectMaps<NT> maps) {
    using C = DirectCfg<ES, (ES == 1 && R > 4)>;
    constexpr int kTaps = 2 * R + 1;
    constexpr int kLaneFloats = 4 * ES;      // floats of one row owned by a lane (4 pixels)
    constexpr int kHalo = ES * R;            // floats needed from the neighbour lane on either side
    constexpr int kRows = C::kRowsPerWarp + 2 * R;  // input rows a warp walks through
    extern __shared__ __align__(1024) unsigned char smem[];
    constexpr int kStages = C::stages(R), kPad = C::pad(R), kPitch = C::pitch(R);
    unsigned char* misc = smem + kStages * C::stage_bytes(R);
    float* s_g = reinterpret_cast<float*>(misc);                  // taps (<= 17)
    double* s_red = reinterpret_cast<double*>(misc + 96);
    uint64_t* full = reinterpret_cast<uint64_t*>(misc + 128);     // [kStages]
    uint64_t* empty = full + kStages;                             // [kStages]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int first = blockIdx.x, stride = gridDim.x;
    const int n_it = (a.units - first + stride - 1) / stride;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        tma_prefetch_desc(&maps.m[0]);
        for (int b = 0; b < kStages; ++b) {
            mbar_init(full + b, 1);
            mbar_init(empty + b, C::kGroupWarps);
        }
        mbar_fence_init();
        // Warm L2 with the first units before waiting for the previous grid (a prefetch only moves DRAM -> L2, where the
        // previous grid's stores land as well, so it cannot expose stale data): takes the DRAM latency out of the start-up
        // of back-to-back launches.
        for (int it = 0; it < n_it && it < kStages; ++it) {
            int unit = first + it * stride;
            const CUtensorMap* tmap = &maps.m[multi_locate<NT>(unit, a.units_per_tensor)];
            const int image = unit / C::kUnitsPerImage, row0 = (unit % C::kUnitsPerImage) * C::kUnitRows - R;
#pragma unroll
            for (int p = 0; p < C::kPanels; ++p) tma_prefetch_l2_3d(tmap, p * C::kPanelFloats - kPad, row0, image);
        }
    }
    if constexpr (PT) {
        __syncthreads();  // barrier init visible
    } else {
        if (a.rule == BGAN_RULE_EXPLICIT) asm volatile("griddepcontrol.wait;" ::: "memory");  // (the taps are read from global memory)
        normalised_taps(a.rule, a.std_eff, R, a.taps_dev, s_g, s_red);  // ends with __syncthreads(): barrier init is visible too
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");  // x may be written, y may still be read by the previous grid

    if (warp == kConsumerWarps) {
        // ================= producer: one elected lane feeds the ring =================================
        if (lane == 0) {
            for (int it = 0; it < n_it; ++it) {
                const int b = it % kStages, use = it / kStages;
                if (use > 0) mbar_wait(empty + b, (use - 1) & 1);
                int unit = first + it * stride;
                const CUtensorMap* tmap = &maps.m[multi_locate<NT>(unit, a.units_per_tensor)];
                const int image = unit / C::kUnitsPerImage, row0 = (unit % C::kUnitsPerImage) * C::kUnitRows - R;
                unsigned char* dst = smem + b * C::stage_bytes(R);
                mbar_arrive_expect_tx(full + b, static_cast<uint32_t>(C::box_rows(R) * kPitch * C::kPanels * sizeof(float)));
#pragma unroll
                for (int p = 0; p < C::kPanels; ++p)
                    tma_load_3d(dst + p * C::panel_bytes(R), tmap, full + b, p * C::kPanelFloats - kPad, row0, image);
            }
        }
        return;
    }

    // ================= consumers ========================================================================
    float g[kTaps];
#pragma unroll
    for (int t = 0; t < kTaps; ++t) g[t] = PT ? a.taps[t] : s_g[t];
    // this lane's floats inside a staged row: panel (NHWC rows are split into two boxes) and offset within it
    const int panel = lane / C::kLanesPerPanel;
    const int lane_off = panel * static_cast<int>(C::panel_bytes(R) / sizeof(float)) + (lane % C::kLanesPerPanel) * kLaneFloats;  // (window start for ES == 1)

    const int group = warp / C::kGroupWarps, gwarp = warp % C::kGroupWarps;
    for (int it = group; it < n_it; it += C::kGroups) {
        const int b = it % kStages, use = it / kStages;
        int unit = first + it * stride;
        float* const y = multi_pick<NT>(a.y, multi_locate<NT>(unit, a.units_per_tensor));
        const int image = unit / C::kUnitsPerImage, out_row0 = (unit % C::kUnitsPerImage) * C::kUnitRows + gwarp * C::kRowsPerWarp;
        mbar_wait(full + b, use & 1);
        const float* src = reinterpret_cast<const float*>(smem + b * C::stage_bytes(R)) + lane_off + gwarp * C::kRowsPerWarp * kPitch;
        float* dst = y + (static_cast<size_t>(image) * kD + out_row0) * C::kRowFloats + lane * kLaneFloats;

        if constexpr (ES == 1 && R >= 4) {
            // NCHW, r >= 4 (below that the kernel is a copy whatever its arithmetic costs, and the scalar form with its taps from
            // shared memory measured 0.5 - 1 us faster per launch): the lane's four pixels as two packed pairs.  E[m] = (win[2m], win[2m+1]) comes straight out of the 128-bit
            // shared-memory loads; a tap whose window starts at an odd float needs the pairs in between, (win[2m+1], win[2m+2]),
            // which cost two register moves each (ALU pipe, beside the FMA pipe).
            constexpr int kWin = 4 + 2 * kPad, kBase = kPad - R;
            f32x2 gg[kTaps];
#pragma unroll
            for (int t = 0; t < kTaps; ++t) gg[t] = pack2(g[t], g[t]);
            f32x2 hrow[kRows][2];  // horizontally blurred rows; only 2r + 1 of them are live at any time
#pragma unroll
            for (int i = 0; i < kRows; ++i) {
                f32x2 E[kWin / 2], O[kWin / 2 - 1];
#pragma unroll
                for (int q = 0; q < kWin / 4; ++q) {
                    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(src + i * kPitch + 4 * q);
                    E[2 * q] = v.x;
                    E[2 * q + 1] = v.y;
                }
#pragma unroll
                for (int m = 0; m + 1 < kWin / 2; ++m) O[m] = pack2(hi_of(E[m]), lo_of(E[m + 1]));
                auto pair_at = [&](int start) { return (start & 1) ? O[start >> 1] : E[start >> 1]; };  // (win[start], win[start + 1])
                f32x2 h0 = mul2(gg[0], pair_at(kBase)), h1 = mul2(gg[0], pair_at(kBase + 2));
#pragma unroll
                for (int t = 1; t < kTaps; ++t) {
                    h0 = fma2(gg[t], pair_at(kBase + t), h0);
                    h1 = fma2(gg[t], pair_at(kBase + t + 2), h1);
                }
                hrow[i][0] = h0;
                hrow[i][1] = h1;
                if (i >= 2 * R) {
                    const int j = i - 2 * R;  // output row of the band: taps over hrow[j .. j + 2r]
                    f32x2 v0 = mul2(gg[0], hrow[j][0]), v1 = mul2(gg[0], hrow[j][1]);
#pragma unroll
                    for (int t = 1; t < kTaps; ++t) {
                        v0 = fma2(gg[t], hrow[j + t][0], v0);
                        v1 = fma2(gg[t], hrow[j + t][1], v1);
                    }
                    __stcs(reinterpret_cast<float4*>(dst + static_cast<size_t>(j) * C::kRowFloats),
                           make_float4(lo_of(v0), hi_of(v0), lo_of(v1), hi_of(v1)));
                }
            }
        } else {
        float hrow[kRows][kLaneFloats];  // horizontally blurred rows; only 2r + 1 of them are live at any time
#pragma unroll
            for (int i = 0; i < kRows; ++i) {
                // win[kOff + o + ES * (t - r)] is the input of output float o under tap t
                constexpr int kWin = ES == 1 ? 4 + 2 * kPad : kLaneFloats + 2 * kHalo;
                constexpr int kOff = ES == 1 ? kPad : kHalo;
                float win[kWin];
                if constexpr (ES == 1) {
#pragma unroll
                    for (int q = 0; q < kWin / 4; ++q) {
                        const float4 v = *reinterpret_cast<const float4*>(src + i * kPitch + 4 * q);
                        win[4 * q + 0] = v.x;
                        win[4 * q + 1] = v.y;
                        win[4 * q + 2] = v.z;
                        win[4 * q + 3] = v.w;
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < ES; ++q) {
                        const float4 v = *reinterpret_cast<const float4*>(src + i * kPitch + 4 * q);
                        win[kHalo + 4 * q + 0] = v.x;
                        win[kHalo + 4 * q + 1] = v.y;
                        win[kHalo + 4 * q + 2] = v.z;
                        win[kHalo + 4 * q + 3] = v.w;
                    }
#pragma unroll
                    for (int k = 0; k < kHalo; ++k) {
                        const float l = __shfl_up_sync(0xffffffffu, win[kLaneFloats + k], 1);    // neighbour's last kHalo floats
                        const float r = __shfl_down_sync(0xffffffffu, win[kHalo + k], 1);        // neighbour's first kHalo floats
                        win[k] = lane == 0 ? 0.0f : l;
                        win[kHalo + kLaneFloats + k] = lane == 31 ? 0.0f : r;
                    }
                }
#pragma unroll
                for (int o = 0; o < kLaneFloats; ++o) {
                    float acc = g[0] * win[kOff - kHalo + o];
#pragma unroll
                    for (int t = 1; t < kTaps; ++t) acc = fmaf(g[t], win[kOff - kHalo + o + ES * t], acc);
                    hrow[i][o] = acc;
                }
                if (i >= 2 * R) {
                    const int j = i - 2 * R;  // output row of the band: taps over hrow[j .. j + 2r]
                    float out[kLaneFloats];
#pragma unroll
                    for (int o = 0; o < kLaneFloats; ++o) {
                        float acc = g[0] * hrow[j][o];
#pragma unroll
                        for (int t = 1; t < kTaps; ++t) acc = fmaf(g[t], hrow[j + t][o], acc);
                        out[o] = acc;
                    }
#pragma unroll
                    for (int q = 0; q < ES; ++q)
                        __stcs(reinterpret_cast<float4*>(dst + static_cast<size_t>(j) * C::kRowFloats) + q,
                               make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]));
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty + b);
    }
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn direct_encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// taps.cuh's normalised_taps on the host: the same fp32 expressions in the same order, the sum in double.  (The host's expf and
// the device's differ by an ulp now and then -- far inside the blur's 2e-6 parity tolerance, like the reference's own
// torch.exp.)
void host_taps(int rule, float s, int r, float* out) {
    constexpr float kSqrt2PiHost = 2.5066282746310002f;
    double sum = 0.0;
    for (int t = 0; t <= 2 * r; ++t) {
        const float x = static_cast<float>(t - r);
        float g;
        if (rule == BGAN_RULE_TORCH) {
            const float num = (x * x) / (2.0f * (s * s));
            const float den = kSqrt2PiHost * s;
            g = expf(-(num / den));
        } else {
            g = expf(-((x * x) / (2.0f * (s * s)))) / (kSqrt2PiHost * s);
        }
        out[t] = g;
        sum += static_cast<double>(g);
    }
    const float total = static_cast<float>(sum);
    for (int t = 0; t <= 2 * r; ++t) out[t] = out[t] / total;
}

template <int ES, int R, int NT, bool PT>
int launch_direct_nt(const MultiIO& io, int images, int rule, float std_eff, const float* taps_dev, cudaStream_t stream) {
    using C = DirectCfg<ES, (ES == 1 && R > 4)>;
    DirectMaps<NT> maps;
    memset(&maps, 0, sizeof(maps));
    thread_local bool bound = false;  // cuTensorMapEncodeTiled needs a context bound to the calling thread
    if (!bound) {
        cudaFree(nullptr);
        bound = true;
    }
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(C::kRowFloats), kD, static_cast<cuuint64_t>(images)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(C::kRowFloats) * 4, static_cast<cuuint64_t>(C::kRowFloats) * kD * 4};
    const cuuint32_t box[3] = {static_cast<cuuint32_t>(C::pitch(R)), static_cast<cuuint32_t>(C::box_rows(R)), 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    for (int t = 0; t < io.count; ++t) {
        const CUresult res = direct_encode_fn()(&maps.m[t], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(io.x[t]), dims, strides, box,
                                                estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);  // out-of-bounds rows read as zeros
        BGAN_REQUIRE(res == CUDA_SUCCESS, BGAN_E_UNSUPPORTED, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(res));
    }
    auto kernel = blur_direct128_kernel<ES, R, NT, PT>;
    thread_local int attr_dev = -1;
    int dev = 0;
    BGAN_CUDA_OK(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        BGAN_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(C::smem_bytes(R))));
        attr_dev = dev;
    }
    const int per_tensor = images * C::kUnitsPerImage, units = per_tensor * io.count;
    DirectArgs<NT> a;
    for (int t = 0; t < NT; ++t) a.y[t] = t < io.count ? io.y[t] : nullptr;
    a.taps_dev = taps_dev;
    memset(a.taps, 0, sizeof(a.taps));
    if (PT) host_taps(rule, std_eff, R, a.taps);
    a.units = units;
    a.units_per_tensor = per_tensor;
    a.rule = rule;
    a.std_eff = std_eff;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(static_cast<unsigned>(std::min(units, device_sm_count())));
    cfg.blockDim = dim3(kDirectThreads);
    cfg.dynamicSmemBytes = C::smem_bytes(R);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    const char* no_pdl = getenv("BGAN_BLUR_NO_PDL");
    cfg.attrs = attr;
    cfg.numAttrs = (no_pdl && atoi(no_pdl)) ? 0 : 1;
    BGAN_CUDA_OK(cudaLaunchKernelEx(&cfg, kernel, a, maps));
    return 0;
}

template <int ES, int R>
int launch_direct(const MultiIO& io, int images, int rule, float std_eff, const float* taps_dev, cudaStream_t stream) {
    if (rule == BGAN_RULE_EXPLICIT || !(ES == 1 && R >= 4))
        return io.count == 1 ? launch_direct_nt<ES, R, 1, false>(io, images, rule, std_eff, taps_dev, stream)
                             : launch_direct_nt<ES, R, kMaxMulti, false>(io, images, rule, std_eff, taps_dev, stream);
    return io.count == 1 ? launch_direct_nt<ES, R, 1, true>(io, images, rule, std_eff, taps_dev, stream)
                         : launch_direct_nt<ES, R, kMaxMulti, true>(io, images, rule, std_eff, taps_dev, stream);
}

template <int ES>
int launch_direct_r(const MultiIO& io, int images, int radius, int rule, float std_eff, const float* taps_dev, cudaStream_t stream) {
    switch (radius) {
        case 1: return launch_direct<ES, 1>(io, images, rule, std_eff, taps_dev, stream);
        case 2: return launch_direct<ES, 2>(io, images, rule, std_eff, taps_dev, stream);
        case 3: return launch_direct<ES, 3>(io, images, rule, std_eff, taps_dev, stream);
        case 4: return launch_direct<ES, 4>(io, images, rule, std_eff, taps_dev, stream);
        default: break;
    }
    if constexpr (ES == 1) {
        switch (radius) {
            case 5: return launch_direct<ES, 5>(io, images, rule, std_eff, taps_dev, stream);
            case 6: return launch_direct<ES, 6>(io, images, rule, std_eff, taps_dev, stream);
            case 7: return launch_direct<ES, 7>(io, images, rule, std_eff, taps_dev, stream);
            case 8: return launch_direct<ES, 8>(io, images, rule, std_eff, taps_dev, stream);
            default: break;
        }
    }
    set_error("blur_direct: radius %d not supported", radius);
    return BGAN_E_UNSUPPORTED;
}

}  // namespace

// Measured on B200 (256 x 3 x 128 x 128, us per launch in a plain stream; tensor-core kernel: 25 - 26.7 NCHW / 47 NHWC for every radius):
//   NCHW  r = 1 .. 3: 18.8 - 19.7 (a device-to-device copy of the same bytes takes 18.6),  r = 4: 19.6,  r = 5: 22.7,  r = 6: 24.6,
//         r = 7: 28.9,  r = 8: 32.4   (round 1, taps in registers and whole-plane units for every r: r = 4: 21.0, r = 5: 26.8, r = 6: 30.0)
//   NHWC  r = 2: 23.0,  r = 3: 24.5,  r = 4: 30.3  (round 1)
constexpr int kDirectMaxRadius1 = 8, kDirectMaxRadius3 = 4;

// planes: N * C planes of 128 x 128 (ES == 1) or N images of 128 x 384 interleaved floats (ES == 3)
bool blur_direct_eligible(int planes, int H, int Wc, int ES, int radius, const float* x, const float* y) {
    const char* off = getenv("BGAN_BLUR_NO_DIRECT");
    if (off && atoi(off)) return false;
    const char* max_r = getenv("BGAN_BLUR_DIRECT_MAX_R");
    const int cap = ES == 1 ? kDirectMaxRadius1 : kDirectMaxRadius3;
    const int rmax = max_r ? std::min(atoi(max_r), cap) : cap;
    return (ES == 1 || ES == 3) && H == kD && Wc == kD * ES && planes > 0 && radius >= 1 && radius <= rmax &&
           reinterpret_cast<uintptr_t>(x) % 16 == 0 && reinterpret_cast<uintptr_t>(y) % 16 == 0 && direct_encode_fn() != nullptr;
}

// `planes` per tensor; io.count equally shaped tensors are blurred by one launch
int blur_direct_launch(const MultiIO& io, int planes, int es, int radius, int rule, float std_eff, const float* taps_dev,
                       cudaStream_t stream) {
    return es == 1 ? launch_direct_r<1>(io, planes, radius, rule, std_eff, taps_dev, stream)
                   : launch_direct_r<3>(io, planes, radius, rule, std_eff, taps_dev, stream);
}

}  // namespace bgan

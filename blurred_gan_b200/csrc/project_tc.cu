// Tensor-core projection for the sliced Wasserstein distance (sm_100a: tcgen05.mma + TMEM + bulk async copies).
//
//   proj[d][row] = sum_k desc[row][k] dirs[k][d]          (reference sliced_wasserstein.py:45-46: np.matmul(A, directions))
//
// desc is [rows, K] row-major (K = 147 for 7 x 7 x 3 neighbourhoods), dirs [K, D] row-major, the result direction-major
// so that every direction is one contiguous sort segment.  At BASELINE configs[4] one call is 2,097,152 x 147 x 512 =
// 316 GFLOP; the FP32 FFMA kernel (swd.cu: project_kernel) runs it at 21.8 TFLOP/s.  Here it is a GEMM on the fp16 tensor
// pipe with the same error-compensated split as the blur (blur_tc.cu): every operand is hi + lo with hi = fp16(a),
// lo = fp16(a - hi), A B ~= A_hi B_hi + A_hi B_lo + A_lo B_hi accumulated in fp32 in TMEM (measured error vs fp64:
// 3e-7 of the largest value, the same as an fp32 sgemm).  Every row of desc and the direction block are rescaled by powers
// of two so that their largest magnitude lands in [2^12, 2^13) -- exact, and keeps both pieces inside fp16's range for
// any input magnitude -- and the inverse scales are applied to the output.
//
// One CTA owns a block of 128 directions (B operand, converted once, resident in shared memory as swizzled K-major
// tiles) and walks over tiles of 128 rows:
//   bulk async copy of the tile (a contiguous 128 x K x 4 byte run of desc) into a staging buffer
//   compute warps: row maxima -> scales; staging -> fp16 hi | lo -> TMEM (A operand, lane = row)
//   MMA warp:      K/16 x 3 tcgen05.mma (TS mode: A from TMEM, B from shared memory) into D[tile % 2]
//   store warps:   D -> registers -> proj[d][row] (lane = row: 128-byte coalesced stores), overlapping the next tile
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace bgan {
namespace {

constexpr int kTM = 128;                       // rows per tile (= TMEM lanes)
constexpr int kTN = 128;                       // directions per CTA (= accumulator columns)
constexpr int kMaxK = 192;                     // three K-blocks of 64
constexpr int kCompThreads = 512, kCompWarps = 16;
constexpr int kMmaWarpP = kCompWarps;          // warp 16
constexpr int kStoreWarp0P = kCompWarps + 1;   // warps 17..24
constexpr int kStoreWarpsP = 8;
constexpr int kThreadsP = kCompThreads + 32 + 32 * kStoreWarpsP;  // 800
constexpr uint32_t kKBlockBytesP = kTN * 128;  // 16384: 128 rows x 64 K-elements, 128-byte swizzle
constexpr uint32_t kOffBhi = 0;                // B hi: up to 3 K-blocks
constexpr uint32_t kOffBlo = 3 * kKBlockBytesP;
constexpr uint32_t kOffStageP = 6 * kKBlockBytesP;                 // 98304: fp32 staging tile, up to 128 x 192 x 4 = 98304 bytes
constexpr uint32_t kOffMiscP = kOffStageP + kTM * kMaxK * 4;       // 196608
// misc: mbarriers (64 B) | tmem address | row maxima [4][128] | inverse row scales [4][128] (ring) | reduction scratch [32]
constexpr uint32_t kSmemP = kOffMiscP + 64 + 16 + 4 * 128 * 4 + 4 * 128 * 4 + 32 * 4;
constexpr uint32_t kIdescP = (1u << 4) | (0u << 7) | (0u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);  // f16 x f16 -> f32, M = N = 128
constexpr int kBarReady0 = 1, kBarComp = 2, kBarReady1 = 3;

struct PArgs {
    const float* desc;
    const float* dirs;
    float* proj;
    int64_t rows;
    int K, D;
    int tiles;  // ceil(rows / 128)
};

__device__ __forceinline__ uint32_t b_tile_off(int row, int k) {
    return (k >> 6) * kKBlockBytesP + row * 128 + ((((k & 63) >> 3) ^ (row & 7)) << 4) + (k & 7) * 2;
}
__device__ __forceinline__ uint64_t b_desc(uint32_t saddr) {
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint64_t b_kstep(int k) { return static_cast<uint64_t>((k >> 2) * (kKBlockBytesP >> 4) + (k & 3) * 2); }

__device__ __forceinline__ void mma_ts_p(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(kIdescP), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void tmem_st8_p(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32_p(uint32_t taddr, float (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
          "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]),
          "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]),
          "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait_p(float (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]),
                   "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]), "+f"(v[16]), "+f"(v[17]), "+f"(v[18]),
                   "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]), "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]),
                   "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31])
                 :
                 : "memory");
}
__device__ __forceinline__ void p_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void p_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void p_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// contiguous global -> shared bulk copy (16-byte aligned, size a multiple of 16), completion on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}

__device__ __forceinline__ void split2_p(float a, float b, __half2& hi, __half2& lo) {
    hi = __floats2half2_rn(a, b);
    const float2 back = __half22float2(hi);
    lo = __floats2half2_rn(a - back.x, b - back.y);
}
// power of two s with m * s in [2^12, 2^13) (m = 0 and subnormal m: the largest usable scale), and its inverse
__device__ __forceinline__ void pow2_scale(float m, float& scale, float& inv) {
    int eb = (__float_as_int(m) >> 23) & 0xff;
    eb = min(max(eb, 13), 254);
    scale = __int_as_float((266 - eb) << 23);
    inv = __int_as_float((eb - 12) << 23);
}

__global__ void __launch_bounds__(kThreadsP, 1) project_tc_kernel(const PArgs a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* stage = reinterpret_cast<float*>(smem + kOffStageP);
    uint64_t* mbar_load = reinterpret_cast<uint64_t*>(smem + kOffMiscP);  // staging tile landed
    uint64_t* mbar_mma = mbar_load + 1;                                   // [2]: MMAs of an even / odd tile complete
    uint64_t* mbar_dfree = mbar_load + 3;                                 // [2]: D[b] is in the store warps' registers
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + kOffMiscP + 64);
    float* s_rowmax = reinterpret_cast<float*>(smem + kOffMiscP + 80);    // [4][128]
    float* s_inv = s_rowmax + 4 * 128;                                    // [4][128]: 1 / scale of the rows of tile i % 4
    float* s_red = s_inv + 4 * 128;                                       // [32]
    const uint32_t sbase = smem_u32(smem);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int K = a.K, Kpad = (K + 15) & ~15, ksteps = Kpad >> 4;
    const int d0 = blockIdx.y * kTN;
    const int first = blockIdx.x, stride = gridDim.x;
    const int n = first < a.tiles ? (a.tiles - first + stride - 1) / stride : 0;
    const uint32_t full_bytes = static_cast<uint32_t>(kTM) * K * 4;
    // a tile is loaded with one bulk copy when it is complete and 16-byte aligned (every tile but a ragged last one)
    const bool bulk_ok = (full_bytes % 16 == 0) && (reinterpret_cast<uintptr_t>(a.desc) % 16 == 0);
    auto tile_rows = [&](int t) { return static_cast<int>(min(static_cast<int64_t>(kTM), a.rows - static_cast<int64_t>(t) * kTM)); };

    if (tid == 0) {
        mbar_init(mbar_load, 1);
        mbar_init(mbar_mma + 0, 1);
        mbar_init(mbar_mma + 1, 1);
        mbar_init(mbar_dfree + 0, kStoreWarpsP);
        mbar_init(mbar_dfree + 1, kStoreWarpsP);
        mbar_fence_init();
    }
    if (warp == kMmaWarpP) {  // TMEM: A hi | lo [0, 192), D[0] [256, 384), D[1] [384, 512)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    p_fence_before();
    __syncthreads();
    p_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t a_hi = tmem, a_lo = tmem + 96, dacc = tmem + 256;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int row = (warp & 3) * 32 + lane;  // TMEM lane = row of the tile

    // first tile on its way while B is built
    auto issue_load = [&](int t) {  // thread 0 of the compute warps
        const int mr = tile_rows(t);
        if (bulk_ok && mr == kTM) {
            mbar_arrive_expect_tx(mbar_load, full_bytes);
            bulk_load_1d(stage, a.desc + static_cast<size_t>(t) * kTM * K, full_bytes, mbar_load);
        }
    };
    if (tid == 0 && n > 0) {
        issue_load(first);
        if (n > 1 && bulk_ok && tile_rows(first + stride) == kTM)
            bulk_prefetch_l2(a.desc + static_cast<size_t>(first + stride) * kTM * K, full_bytes);
    }

    // ---- B: the 128 directions of this CTA, scaled by a power of two, fp16 hi | lo, K-major swizzled tiles ----------------
    {
        float m = 0.0f;
        for (int idx = tid; idx < K * kTN; idx += kThreadsP) {
            const int k = idx / kTN, j = idx - k * kTN;
            if (d0 + j < a.D) m = fmaxf(m, fabsf(__ldg(a.dirs + static_cast<size_t>(k) * a.D + d0 + j)));
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) s_red[warp] = m;
        __syncthreads();
    }
    float scale_b, inv_b;
    {
        float m = 0.0f;
        for (int w = 0; w < kThreadsP / 32; ++w) m = fmaxf(m, s_red[w]);
        pow2_scale(m, scale_b, inv_b);
    }
    for (int idx = tid; idx < (Kpad >> 1) * kTN; idx += kThreadsP) {
        const int kp = idx / kTN, j = idx - kp * kTN;  // lanes <-> consecutive directions: coalesced reads
        const int k = 2 * kp;
        const bool okd = d0 + j < a.D;
        const float v0 = (okd && k < K) ? __ldg(a.dirs + static_cast<size_t>(k) * a.D + d0 + j) * scale_b : 0.0f;
        const float v1 = (okd && k + 1 < K) ? __ldg(a.dirs + static_cast<size_t>(k + 1) * a.D + d0 + j) * scale_b : 0.0f;
        __half2 hi, lo;
        split2_p(v0, v1, hi, lo);
        *reinterpret_cast<__half2*>(smem + kOffBhi + b_tile_off(j, k)) = hi;
        *reinterpret_cast<__half2*>(smem + kOffBlo + b_tile_off(j, k)) = lo;
    }
    fence_proxy_async();
    __syncthreads();

    if (warp == kMmaWarpP) {
        // ================= MMA issuer ======================================================================================
        for (int i = 0; i < n; ++i) {
            named_bar_sync((i & 1) ? kBarReady1 : kBarReady0, kCompThreads + 32);  // A(i) is in TMEM
            if (lane == 0) {
                if (i >= 2) mbar_wait(mbar_dfree + (i & 1), ((i - 2) >> 1) & 1);  // D[i & 1] drained
                p_fence_after();
                const uint32_t d = dacc + (i & 1) * 128;
                const uint64_t bh = b_desc(sbase + kOffBhi), bl = b_desc(sbase + kOffBlo);
                for (int k = 0; k < ksteps; ++k) mma_ts_p(d, a_hi + 8 * k, bl + b_kstep(k), k > 0);
                for (int k = 0; k < ksteps; ++k) mma_ts_p(d, a_lo + 8 * k, bh + b_kstep(k), 1);
                for (int k = 0; k < ksteps; ++k) mma_ts_p(d, a_hi + 8 * k, bh + b_kstep(k), 1);
                p_commit(mbar_mma + (i & 1));
            }
            __syncwarp();
        }
    } else if (warp >= kStoreWarp0P) {
        // ================= store warps: D[i & 1] -> registers -> proj[d][row], lane = row ==================================
        const int c0 = ((warp - kStoreWarp0P) >> 2) * 64;  // first accumulator column (= direction) of this warp
        for (int i = 0; i < n; ++i) {
            const int t = first + i * stride;
            mbar_wait(mbar_mma + (i & 1), (i >> 1) & 1);
            p_fence_after();
            const float inv = s_inv[(i & 3) * 128 + row] * inv_b;
            float o0[32], o1[32];
            tmem_ld32_p(dacc + (i & 1) * 128 + lane_base + c0, o0);
            tmem_ld32_p(dacc + (i & 1) * 128 + lane_base + c0 + 32, o1);
            tmem_ld_wait_p(o0);
            tmem_ld_wait_p(o1);
            p_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(mbar_dfree + (i & 1));
            const int64_t r = static_cast<int64_t>(t) * kTM + row;
            if (r < a.rows) {
                float* q = a.proj + static_cast<size_t>(d0 + c0) * a.rows + r;
                const int ncol = min(64, a.D - d0 - c0);  // (<= 0: nothing to store)
#pragma unroll
                for (int j = 0; j < 32; ++j, q += a.rows) {
                    asm volatile("" : "+l"(q));
                    if (j < ncol) __stcs(q, o0[j] * inv);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j, q += a.rows) {
                    asm volatile("" : "+l"(q));
                    if (32 + j < ncol) __stcs(q, o1[j] * inv);
                }
            }
        }
    } else {
        // ================= compute warps: staging -> row scales -> fp16 hi | lo -> TMEM ====================================
        const int g = warp >> 2;  // K-steps g, g + 4, g + 8 of this thread's row
        uint32_t par_load = 0;
        for (int i = 0; i < n; ++i) {
            const int t = first + i * stride;
            const int mr = tile_rows(t);
            if (bulk_ok && mr == kTM) {
                mbar_wait(mbar_load, par_load);
                par_load ^= 1;
            } else {  // ragged last tile (or unaligned input): plain loads into staging, rows beyond the end are zero
                const float* src = a.desc + static_cast<size_t>(t) * kTM * K;
                for (int idx = tid; idx < kTM * K; idx += kCompThreads) stage[idx] = idx < mr * K ? __ldg(src + idx) : 0.0f;
                named_bar_sync(kBarComp, kCompThreads);
            }
            // pass 1: maximum of this thread's part of its row (lanes <-> consecutive rows: pitch K is odd for K = 147)
            const float* srow = stage + row * K;
            float m = 0.0f;
            for (int s = g; s < ksteps; s += 4) {
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int k = 16 * s + e;
                    if (k < K) m = fmaxf(m, fabsf(srow[k]));
                }
            }
            s_rowmax[g * 128 + row] = m;
            named_bar_sync(kBarComp, kCompThreads);
            m = fmaxf(fmaxf(s_rowmax[row], s_rowmax[128 + row]), fmaxf(s_rowmax[256 + row], s_rowmax[384 + row]));
            float scale, inv;
            pow2_scale(m, scale, inv);
            if (g == 0) s_inv[(i & 3) * 128 + row] = inv;  // (read by the store warps after the MMAs of this tile)
            // A is free once the MMAs of the previous tile are complete
            if (i >= 1) mbar_wait(mbar_mma + ((i - 1) & 1), ((i - 1) >> 1) & 1);
            p_fence_after();
            // pass 2: convert and store to TMEM (hi: 8 columns per K-step at a_hi + 8 s, lo at a_lo + 8 s)
            for (int s = g; s < ksteps; s += 4) {
                uint32_t zh[8], zl[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int k = 16 * s + 2 * e;
                    const float v0 = k < K ? srow[k] * scale : 0.0f;
                    const float v1 = k + 1 < K ? srow[k + 1] * scale : 0.0f;
                    __half2 hi, lo;
                    split2_p(v0, v1, hi, lo);
                    zh[e] = *reinterpret_cast<uint32_t*>(&hi);
                    zl[e] = *reinterpret_cast<uint32_t*>(&lo);
                }
                tmem_st8_p(a_hi + lane_base + 8 * s, zh);
                tmem_st8_p(a_lo + lane_base + 8 * s, zl);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            p_fence_before();
            named_bar_sync(kBarComp, kCompThreads);  // everybody is done with staging
            if (tid == 0 && i + 1 < n) {
                issue_load(t + stride);
                if (i + 2 < n && bulk_ok && tile_rows(t + 2 * stride) == kTM)
                    bulk_prefetch_l2(a.desc + static_cast<size_t>(t + 2 * stride) * kTM * K, full_bytes);
            }
            named_bar_arrive((i & 1) ? kBarReady1 : kBarReady0, kCompThreads + 32);  // -> MMAs of tile i
        }
    }
    p_fence_before();
    __syncthreads();
    if (warp == kMmaWarpP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace

bool project_tc_eligible(int64_t rows, int k, int d) {
    const char* off = getenv("BGAN_SWD_NO_TC_PROJECT");
    if (off && atoi(off)) return false;
    return rows > 0 && k >= 1 && k <= kMaxK && d >= 1 && (rows + kTM - 1) / kTM < (int64_t(1) << 31);
}

int project_tc_launch(const float* desc, int64_t rows, int k, const float* dirs, int d, float* proj, cudaStream_t stream) {
    thread_local int attr_dev = -1;
    int dev = 0;
    BGAN_CUDA_OK(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        BGAN_CUDA_OK(cudaFuncSetAttribute(project_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemP)));
        attr_dev = dev;
    }
    const int tiles = static_cast<int>((rows + kTM - 1) / kTM);
    const int chunks = ceil_div(d, kTN);
    // one CTA per SM in total: every direction block gets an equal share of the SMs, each CTA walks over its tiles
    const int per_chunk = std::max(1, std::min(tiles, device_sm_count() / chunks));
    PArgs a{desc, dirs, proj, rows, k, d, tiles};
    project_tc_kernel<<<dim3(per_chunk, chunks), kThreadsP, kSmemP, stream>>>(a);
    BGAN_LAUNCH_OK();
    return 0;
}

}  // namespace bgan

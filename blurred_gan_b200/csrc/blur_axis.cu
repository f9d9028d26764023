// Tensor-core blur for planes larger than 128 x 128 (BASELINE configs[3]: 1024 x 1024, radius up to 192).
//
// One kernel, applied twice:   out[p][j][i] = sum_k T[j][k] in[p][i][k],   T[j][k] = tap[k - j + r]  (|k - j| <= r)
// i.e. "blur along the contiguous axis and write the result transposed".  Pass 1 turns X[h][w] into Z^T[w'][h]
// (scratch), pass 2 turns Z^T[w'][h] into Y[h'][w']: both passes read rows that are contiguous along the axis they
// blur -- the natural K-major B operand -- and both stores are 128-byte row segments.  With the CUDA-core tile kernel the
// wide blurs of configs[3] cost 2 (2r + 1) = 770 FMAs per pixel; here the cost per 128 x 128 output tile is
// (2 nb + 1) x 24 tcgen05.mma, nb = ceil(r / 128) <= 2.
//
// Tile: 128 outputs j (accumulator rows) x 128 input rows i (accumulator columns), accumulated over the 2 nb + 1 blocks
// of 128 k that the band of T touches.  A = T block (offset d = L - J): core matrices of a Toeplitz block only depend on
// 16 d + kchunk - rowgroup, so all 2 nb + 1 blocks are views (LBO = SBO = 128 B descriptors, row groups reversed, see
// blur_tc.cu) of one strip of 32 nb + 31 core matrices.  B = the fp32 block of `in`, staged by TMA (out-of-bounds zero
// fill = the reference's zero padding), scaled per row by a power of two (row maxima come from a tiny pre-pass for pass 1
// and from pass 1's epilogue for pass 2) and split into fp16 hi | lo by the compute warps.  fp32 parity as in blur_tc.cu:
// A B ~= A1 B1 + 2^-11 (A1 B2 + A2 B1), here with one TMEM accumulator for the main and one for the cross terms, combined
// in the epilogue.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "taps.cuh"

namespace bgan {
namespace {

constexpr int kB = 128;                        // block edge
constexpr int kAxThreads = 512;                // compute threads (16 warps)
constexpr int kAxWarps = kAxThreads / 32;
constexpr int kAxStoreWarps = 4;               // warps 17..20: epilogue (accumulators -> global), off the critical path
constexpr int kAxCta = kAxThreads + 32 + 32 * kAxStoreWarps;  // + the MMA-issuing warp + the store warps
constexpr int kAxMaxNb = 2;                    // radius <= 256
// B operand (fp16 hi | lo of a 128 x 128 block of `in`): K-major tiles with the 128-byte hardware swizzle, as in
// blur_tc.cu (two K-blocks of 64 elements; row r owns 128 contiguous bytes per K-block, 16-byte chunk c at c ^ (r % 8)).
// Two sets, so that the conversion of block u + 1 overlaps the MMAs of block u.
constexpr uint32_t kAKBlock = kB * 128;        // 16384: 128 rows x 64 K-elements
constexpr uint32_t kATile = 2 * kAKBlock;      // 32768
constexpr uint32_t kAStage = kB * kB * 4;      // 65536
constexpr uint32_t kAOffStage = 0;             // one fp32 staging block (TMA destination)
constexpr uint32_t kAOffX = kAStage;           // X[set][hi | lo]; the raw taps live here until the first conversion
constexpr uint32_t kAStripBytes = 12288;       // (32 * 2 + 31) core matrices x 128 B = 12160
constexpr uint32_t kAOffS1 = kAOffX + 4 * kATile;
constexpr uint32_t kAOffS2 = kAOffS1 + kAStripBytes;
constexpr uint32_t kAOffMisc = kAOffS2 + kAStripBytes;  // s_scale[128], s_inv[128], barriers
constexpr uint32_t kASmem = kAOffMisc + 1536 + 256;
constexpr uint32_t kAIdesc = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);  // kind::f16, D = f32, M = N = 128, K-major
constexpr int kARev = 0x78;

struct AxisArgs {
    float* out;               // [planes][K][I]
    const float* rowmax_in;   // [planes][I]: max |in| over each row
    float* rowmax_out;        // [planes][K] or nullptr: max |out| over each output row (accumulated with atomicMax)
    int I, K, planes;
    int tiles_i, tiles_j;
    int nb;                   // ceil(radius / 128)
    int rule;
    float std_eff;
    int radius;
    const float* taps_dev;
    long long* trace;         // debug: clock64 stamps of CTA 0 of the sweep kernel (nullptr in production), tools/trace_axis.py
};

// trace layout: three timelines of 256 int64 each (compute thread 0, MMA lane, lane 0 of the first store warp);
// entry 0 of a timeline = number of stamps that follow
__device__ __forceinline__ void ax_stamp(long long* trace, int role, int& n) {
    if (trace != nullptr && blockIdx.x == 0 && n < 254) {
        trace[role * 256 + 1 + n] = clock64();
        trace[role * 256] = ++n;
    }
}

__device__ __forceinline__ uint32_t ax_tile_off(int row, int k) {
    return (k >> 6) * kAKBlock + row * 128 + ((((k & 63) >> 3) ^ (row & 7)) << 4) + (k & 7) * 2;
}
// A operand: the Toeplitz strip, no swizzle, LBO = SBO = 128 B
__device__ __forceinline__ uint64_t ax_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(lbo >> 4) << 16) | (static_cast<uint64_t>(sbo >> 4) << 32) |
           (1ull << 46);
}
// B operand: SWIZZLE_128B K-major tile (LBO unused = 1, SBO = 1024, layout type 2)
__device__ __forceinline__ uint64_t ax_desc_sw128(uint32_t saddr) {
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// descriptor increment (16-byte units) of K-step k: 16 elements = 32 bytes inside the swizzle atom, 4 steps per K-block
__device__ __forceinline__ constexpr uint64_t ax_kstep(int k) { return static_cast<uint64_t>((k >> 2) * (kAKBlock >> 4) + (k & 3) * 2); }
__device__ __forceinline__ void ax_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(kAIdesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void ax_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ax_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ax_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ax_ld32(uint32_t taddr, float (&v)[32]) {
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
__device__ __forceinline__ void ax_ld_wait(float (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]),
                   "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]), "+f"(v[16]), "+f"(v[17]), "+f"(v[18]),
                   "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]), "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]),
                   "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31])
                 :
                 : "memory");
}
__device__ __forceinline__ void ax_split2(float a, float b, __half2& hi, __half2& lo) {
    hi = __floats2half2_rn(a, b);
    const float2 back = __half22float2(hi);
    lo = __floats2half2_rn((a - back.x) * 2048.0f, (b - back.y) * 2048.0f);
}

// max |x| of every row of a [rows][len] array (one warp per row)
__global__ void rowmax_kernel(const float* __restrict__ x, int64_t rows, int len, float* __restrict__ rowmax) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    if (row >= rows) return;
    const float4* p = reinterpret_cast<const float4*>(x + row * len);
    float m = 0.0f;
    for (int i = lane; i < len / 4; i += 32) {
        const float4 v = __ldg(p + i);
        m = fmaxf(m, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) rowmax[row] = m;
}

// A tile and the blocks of its band that lie inside the array: l = d + nb in [l0, l1] (blocks that fall entirely outside
// read zeros only and are skipped; the block d = 0 always exists).
struct TileCoord {
    int plane, ib, jb, l0, l1;
};
__device__ __forceinline__ TileCoord tile_coord(const AxisArgs& a, int tile) {
    TileCoord c;
    c.jb = tile % a.tiles_j;
    c.ib = (tile / a.tiles_j) % a.tiles_i;
    c.plane = tile / (a.tiles_j * a.tiles_i);
    c.l0 = max(0, a.nb - c.jb);
    c.l1 = min(2 * a.nb, a.nb + a.tiles_j - 1 - c.jb);
    return c;
}

// K-steps (16 inputs k each) of block l = d + nb that the band |k - j| <= r of a 128-output tile touches: the centre block
// entirely, a block to the left only its last columns (k_local >= 128 |d| - r), a block to the right only its first ones
// (k_local <= r - 1 - 128 (d - 1)).  Everything else multiplies zeros of T and is neither converted nor issued.
__device__ __forceinline__ void ks_range(int l, int nb, int radius, int& k0, int& k1) {
    const int d = l - nb;
    k0 = 0;
    k1 = 7;
    if (d < 0) k0 = max(0, -d * kB - radius) >> 4;
    if (d > 0) k1 = min(7, (radius - 1 - (d - 1) * kB) >> 4);
}

// walks the (tile, block) sequence of one CTA
struct BlockCursor {
    int tile, l;
    TileCoord c;
    __device__ __forceinline__ void init(const AxisArgs& a, int first, int n_tiles) {
        tile = first;
        c = tile_coord(a, min(tile, n_tiles - 1));
        l = c.l0;
    }
    __device__ __forceinline__ void advance(const AxisArgs& a, int stride, int n_tiles) {
        if (++l > c.l1) {
            tile += stride;
            if (tile < n_tiles) {
                c = tile_coord(a, tile);
                l = c.l0;
            }
        }
    }
};

// Pipeline (per CTA, blocks numbered u = 0, 1, ... in (tile, l) order):
//   TMA           block u + 1 -> staging (one buffer), issued as soon as block u has been read into registers; block u + 2 is
//                 pulled into L2 at the same time, so the load that follows finds it there
//   compute warps staging(u) -> registers -> fp16 hi | lo tiles X[u & 1]   (only the K-steps the band touches)
//   MMA warp      block u from X[u & 1] while the compute warps convert block u + 1 into the other set
//   store warps   accumulator pair of tile n - 1 -> global while tile n accumulates into the other pair
__global__ void __launch_bounds__(kAxCta, 1) blur_axis_tc_kernel(const AxisArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* s_g = reinterpret_cast<float*>(smem + kAOffX);  // taps (<= 513), only until the strip is built
    float* s_scale = reinterpret_cast<float*>(smem + kAOffMisc);  // [128] per input row of the tile being converted
    float* s_inv = s_scale + kB;                                  // [2][128]: tile parity (read by the store warps one tile later)
    uint64_t* mbar_tma = reinterpret_cast<uint64_t*>(smem + kAOffMisc + 1536);  // one phase per block
    uint64_t* mbar_mma = mbar_tma + 1;       // [2]: the MMAs reading X[set] are complete (the set may be rewritten)
    uint64_t* mbar_acc_full = mbar_tma + 3;  // [2]: all MMAs of a tile complete, accumulator pair valid
    uint64_t* mbar_acc_free = mbar_tma + 5;  // [2]: the store warps hold the accumulator pair in registers
    double* s_red = reinterpret_cast<double*>(mbar_tma + 7);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(mbar_tma + 8);
    const uint32_t sbase = smem_u32(smem);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int first = blockIdx.x, stride = gridDim.x;
    const int n_tiles = a.planes * a.tiles_i * a.tiles_j;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        mbar_init(mbar_tma, 1);
        mbar_init(mbar_mma + 0, 1);
        mbar_init(mbar_mma + 1, 1);
        for (int b = 0; b < 2; ++b) {
            mbar_init(mbar_acc_full + b, 1);
            mbar_init(mbar_acc_free + b, kAxStoreWarps);
        }
        mbar_fence_init();
    }
    if (warp == kAxWarps) {  // TMEM: tile parity b: D_main = [256 b, 256 b + 128), D_cross = [256 b + 128, 256 b + 256)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (a.rule == BGAN_RULE_EXPLICIT) asm volatile("griddepcontrol.wait;" ::: "memory");
    normalised_taps(a.rule, a.std_eff, a.radius, a.taps_dev, s_g, s_red);  // (ends with __syncthreads)
    // ---- the strip: core matrix m = 16 (d + nb) + kchunk + rowgroup', row rho, holds taps t0 .. t0 + 7,
    //      t0 = 8 (m - 16 nb - 15) - rho + r  (zero outside [0, 2r]) ----
    {
        const int n_taps = 2 * a.radius + 1, mats = 32 * a.nb + 31;
        for (int j = tid; j < mats * 32; j += kAxCta) {
            const int m = j >> 5, rho = (j >> 2) & 7, e = j & 3;
            const int ta = 8 * (m - 16 * a.nb - 15) - rho + a.radius + 2 * e, tb = ta + 1;
            __half2 hi, lo;
            ax_split2((ta >= 0 && ta < n_taps) ? s_g[ta] : 0.0f, (tb >= 0 && tb < n_taps) ? s_g[tb] : 0.0f, hi, lo);
            *reinterpret_cast<__half2*>(smem + kAOffS1 + m * 128 + rho * 16 + e * 4) = hi;
            *reinterpret_cast<__half2*>(smem + kAOffS2 + m * 128 + rho * 16 + e * 4) = lo;
        }
    }
    fence_proxy_async();
    ax_fence_before();
    __syncthreads();
    ax_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;  // TMEM lane quadrant of this warp
    const int jl = ((warp & 3) * 32 + lane) ^ kARev;  // output position held by this thread's accumulator row
    asm volatile("griddepcontrol.wait;" ::: "memory");  // `in` may be written, `out` may still be read by the previous grid

    if (warp == kAxWarps) {
        // ================= MMA issuer =================================================================
        int u = 0, n = 0;
        for (int tile = first; tile < n_tiles; tile += stride, ++n) {
            const TileCoord c = tile_coord(a, tile);
            const uint32_t d_main = tmem + (n & 1) * 256, d_cross = d_main + 128;
            for (int l = c.l0; l <= c.l1; ++l, ++u) {
                named_bar_sync(1 + (u & 1), kAxThreads + 32);  // X[u & 1] of block u written
                if (lane == 0) {
                    if (l == c.l0 && n >= 2) mbar_wait(mbar_acc_free + (n & 1), ((n - 2) >> 1) & 1);  // accumulator pair drained
                    ax_fence_after();
                    int k0, k1;
                    ks_range(l, a.nb, a.radius, k0, k1);
                    const uint64_t a1 = ax_desc(sbase + kAOffS1 + l * 16 * 128, 128, 128), a2 = ax_desc(sbase + kAOffS2 + l * 16 * 128, 128, 128);
                    const uint32_t xb = sbase + kAOffX + (u & 1) * 2 * kATile;
                    const uint64_t b1 = ax_desc_sw128(xb), b2 = ax_desc_sw128(xb + kATile);
                    const uint32_t acc0 = l > c.l0;  // the first block of a tile overwrites the accumulators
                    // cross and main products alternate: two independent accumulators
                    for (int k = k0; k <= k1; ++k) {
                        const uint32_t acc = acc0 | (k > k0);
                        ax_mma(d_cross, a1 + k * 16, b2 + ax_kstep(k), acc);
                        ax_mma(d_main, a1 + k * 16, b1 + ax_kstep(k), acc);
                        ax_mma(d_cross, a2 + k * 16, b1 + ax_kstep(k), 1);
                    }
                    ax_commit(mbar_mma + (u & 1));
                    if (l == c.l1) ax_commit(mbar_acc_full + (n & 1));
                }
                __syncwarp();
            }
        }
    } else if (warp > kAxWarps) {
        // ================= store warps: out[j][i] = (D_main + 2^-11 D_cross) / scale_i =========================
        int n = 0;
        for (int tile = first; tile < n_tiles; tile += stride, ++n) {
            const TileCoord c = tile_coord(a, tile);
            const uint32_t d_main = tmem + (n & 1) * 256, d_cross = d_main + 128;
            const float* inv = s_inv + (n & 1) * kB;
            mbar_wait(mbar_acc_full + (n & 1), (n >> 1) & 1);
            ax_fence_after();
            const int j = c.jb * kB + jl;
            float* dst = a.out + (static_cast<size_t>(c.plane) * a.K + j) * a.I + c.ib * kB;
            float mx = 0.0f;
#pragma unroll 1
            for (int cg = 0; cg < 4; ++cg) {
                float dm[32], dc[32];
                ax_ld32(d_main + lane_base + cg * 32, dm);
                ax_ld32(d_cross + lane_base + cg * 32, dc);
                ax_ld_wait(dm);
                ax_ld_wait(dc);
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    dm[e] = fmaf(dc[e], 1.0f / 2048.0f, dm[e]) * inv[cg * 32 + e];
                    mx = fmaxf(mx, fabsf(dm[e]));
                }
                if (cg == 3) {  // accumulators and scales of this tile have been read: the tile after next may reuse them
                    ax_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(mbar_acc_free + (n & 1));
                }
                if (j < a.K) {
                    const int i0 = c.ib * kB + cg * 32;
#pragma unroll
                    for (int e = 0; e < 32; e += 4)
                        if (i0 + e < a.I) __stcs(reinterpret_cast<float4*>(dst + cg * 32 + e), make_float4(dm[e], dm[e + 1], dm[e + 2], dm[e + 3]));
                }
            }
            if (j < a.K && a.rowmax_out != nullptr)
                atomicMax(reinterpret_cast<int*>(a.rowmax_out + static_cast<size_t>(c.plane) * a.K + j), __float_as_int(mx));
        }
    } else {
        // ================= compute warps ==============================================================
        BlockCursor ld, pf;  // thread 0: next block to load into staging / to pull into L2 (one block further ahead)
        ld.init(a, first, n_tiles);
        pf.init(a, first, n_tiles);
        auto issue_load = [&]() {
            if (ld.tile >= n_tiles) return;
            mbar_arrive_expect_tx(mbar_tma, kAStage);
            tma_load_3d(smem + kAOffStage, &tmap, mbar_tma, (ld.c.jb + ld.l - a.nb) * kB, ld.c.ib * kB, ld.c.plane);
            ld.advance(a, stride, n_tiles);
        };
        auto issue_prefetch = [&]() {
            if (pf.tile >= n_tiles) return;
            tma_prefetch_l2_3d(&tmap, (pf.c.jb + pf.l - a.nb) * kB, pf.c.ib * kB, pf.c.plane);
            pf.advance(a, stride, n_tiles);
        };
        if (tid == 0) {
            issue_load();                        // block 0
            issue_prefetch();                    // (block 0: already on its way)
            issue_prefetch();                    // block 1 -> L2
        }
        const float* stage = reinterpret_cast<const float*>(smem + kAOffStage);
        int u = 0, n = 0;
        for (int tile = first; tile < n_tiles; tile += stride, ++n) {
            const TileCoord c = tile_coord(a, tile);
            // row scales of this tile.  s_scale is only read by the compute warps (the barrier at the end of the previous
            // block ordered those reads); s_inv[n & 1] was last read by the store warps for tile n - 2.
            if (n >= 2) mbar_wait(mbar_acc_free + (n & 1), ((n - 2) >> 1) & 1);
            named_bar_sync(3, kAxThreads);  // (the hand-over to the MMA warp no longer blocks: a fast warp must not replace
                                            //  the scales while a slow one still converts the previous tile's last block)
            if (tid < kB) {
                const int i = c.ib * kB + tid;
                const float m = i < a.I ? __ldg(a.rowmax_in + static_cast<size_t>(c.plane) * a.I + i) : 0.0f;
                int eb = (__float_as_int(m) >> 23) & 0xff;
                eb = min(max(eb, 13), 254);
                s_scale[tid] = __int_as_float((266 - eb) << 23);  // max |row| -> [2^12, 2^13)
                s_inv[(n & 1) * kB + tid] = __int_as_float((eb - 12) << 23);
            }
            named_bar_sync(3, kAxThreads);
            for (int l = c.l0; l <= c.l1; ++l, ++u) {
                int k0, k1;
                ks_range(l, a.nb, a.radius, k0, k1);
                const int nks = k1 - k0 + 1;      // K-steps of this block that are needed: 16 nks columns from column 16 k0
                const int f4_row = 4 * nks;       // float4 per row
                // ---- staging -> registers (thread handles float4 g = q * 512 + tid, q < nks: row g / f4_row) ----
                mbar_wait(mbar_tma, u & 1);
                float4 v[8];
                auto locate = [&](int q, int& h, int& w) {  // (recomputed in the second loop rather than kept in registers)
                    const int g = q * kAxThreads + tid;
                    h = nks == 8 ? (g >> 5) : g / f4_row;
                    w = 16 * k0 + ((g - h * f4_row) << 2);
                };
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (q < nks) {
                        int h, w;
                        locate(q, h, w);
                        v[q] = *reinterpret_cast<const float4*>(stage + h * kB + w);
                    }
                }
                // The row scale is applied BEFORE the barrier that releases the staging buffer to the next TMA load: the
                // multiplications consume the loaded registers, so every LDS has really returned when the thread arrives.  (A
                // load that is only issued can still sit in the LSU queue -- behind the store warps' scattered stores for
                // thousands of cycles -- when the barrier completes, and would then read the next block: seen once in
                // ~10 runs as a 6 % error in one plane of test_axis_tensor_core_path_many_work_units_per_cta.)
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (q < nks) {
                        int h, w;
                        locate(q, h, w);
                        const float sc = s_scale[h];
                        v[q].x *= sc, v[q].y *= sc, v[q].z *= sc, v[q].w *= sc;
                        asm volatile("" ::"f"(v[q].x), "f"(v[q].y), "f"(v[q].z), "f"(v[q].w));  // (pins the products in front of the barrier)
                    }
                }
                named_bar_sync(3, kAxThreads);  // staging drained
                if (tid == 0) {
                    issue_load();       // block u + 1 (in L2 already)
                    issue_prefetch();   // block u + 2 -> L2
                }
                if (u >= 2) mbar_wait(mbar_mma + (u & 1), ((u - 2) >> 1) & 1);  // MMAs of block u - 2 complete: X[u & 1] free
                // ---- registers -> X[u & 1] hi | lo tiles (row = i, K = k) ----
                unsigned char* x1 = smem + kAOffX + (u & 1) * 2 * kATile;
                unsigned char* x2 = x1 + kATile;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (q < nks) {
                        int h, w;
                        locate(q, h, w);
                        __half2 hi[2], lo[2];
                        ax_split2(v[q].x, v[q].y, hi[0], lo[0]);
                        ax_split2(v[q].z, v[q].w, hi[1], lo[1]);
                        const uint32_t off = ax_tile_off(h, w);
                        *reinterpret_cast<uint2*>(x1 + off) = *reinterpret_cast<uint2*>(hi);
                        *reinterpret_cast<uint2*>(x2 + off) = *reinterpret_cast<uint2*>(lo);
                    }
                }
                fence_proxy_async();
                ax_fence_before();
                named_bar_arrive(1 + (u & 1), kAxThreads + 32);  // -> MMAs of block u
            }
        }
    }
    ax_fence_before();
    __syncthreads();
    if (warp == kAxWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}


// =====================================================================================================================
// Radius <= 128 (nb = 1): input-stationary sweep.  A CTA takes a strip -- 128 input rows i of one plane -- and walks the
// blocks L = 0, 1, ... along the blurred axis.  Every block is loaded and converted ONCE and multiplied into the (up to)
// three output tiles it touches, J = L - 1 (which it completes), L and L + 1 (which it starts); the tile kernel above loads
// and converts every block three times, once per tile, and with one 64 KB TMA box in flight per SM that latency, not the
// tensor pipe, was its limit (profiles/r02_axis.md).  Three live accumulators + one being drained = 4 x 128 TMEM columns,
// which needs ONE accumulator per tile: the taps are stored scaled by 2^11 (T' = 2^11 T, hi = fp16(T'), lo = fp16(T' - hi):
// the remainder of even a tiny tap stays above fp16's subnormal range relative to what matters), the data remainders
// unscaled (rows are scaled to [2^12, 2^13)), so T' X ~= T'_hi X_hi + T'_hi X_lo + T'_lo X_hi accumulates in any order and
// the epilogue multiplies by 2^-11 / scale_i.
constexpr int kSwSlots = 4;
// Threads: 8 conversion warps, the MMA-issuing warp, 8 store warps (two per TMEM lane quadrant).
// The output leaves through shared memory and TMA stores.  clock64 timelines of the first versions (profiles/r02_axis.md):
// an accumulator row is an output row, so a direct store has its 32 lanes in 32 different lines -- 4096 sixteen-byte store
// requests per tile -- and those requests, not the tensor pipe, set the pace: they also starved the conversion warps' own
// LDS / STS, which share the LSU.  Now a store warp writes its 32 rows x 32 columns into a 128-byte-swizzled box (conflict
// free) and one thread hands two such boxes at a time to the TMA, which writes whole 128-byte row segments off the LSU.
constexpr int kSwCompute = 256;
constexpr int kSwComputeWarps = kSwCompute / 32;
constexpr int kSwStoreWarps = 8;
constexpr int kSwStoreThreads = 32 * kSwStoreWarps;
constexpr int kSwCta = kSwCompute + 32 + kSwStoreThreads;  // 544
// shared memory: input staging, a ring of 3 quarter blocks (32 rows, 16 KB each) | X[2][hi | lo] 128 KB | output boxes
// 2 x 16 KB | strips 2 x 8 KB (63 core matrices each) | scales, barriers.  (With one half block in flight the ~1500 cycles
// of an L2-hit TMA load were exposed twice per block; three quarter blocks in flight hide them.)
constexpr int kSwRing = 3;
constexpr int kSwChunkRows = 32;
constexpr uint32_t kSwChunk = kSwChunkRows * kB * 4;  // 16384
constexpr uint32_t kSwStage = kSwRing * kSwChunk;      // 49152
constexpr uint32_t kSwOffX = kSwStage;
constexpr uint32_t kSwOffOut = kSwOffX + 4 * kATile;
constexpr uint32_t kSwOutBox = kB * 128;           // 128 output rows x 32 columns of fp32
constexpr uint32_t kSwStrip = 8192;
constexpr uint32_t kSwOffS1 = kSwOffOut + 2 * kSwOutBox;
constexpr uint32_t kSwOffS2 = kSwOffS1 + kSwStrip;
constexpr uint32_t kSwOffMisc = kSwOffS2 + kSwStrip;
constexpr uint32_t kSwSmem = kSwOffMisc + 1536 + 256;
static_assert(kSwSmem <= 227 * 1024, "sweep kernel: shared memory");

__device__ __forceinline__ void sw_split_tap2(float a, float b, __half2& hi, __half2& lo) {
    a *= 2048.0f;
    b *= 2048.0f;
    hi = __floats2half2_rn(a, b);
    const float2 back = __half22float2(hi);
    lo = __floats2half2_rn(a - back.x, b - back.y);
}
__device__ __forceinline__ void sw_split_raw2(float a, float b, __half2& hi, __half2& lo) {
    hi = __floats2half2_rn(a, b);
    const float2 back = __half22float2(hi);
    lo = __floats2half2_rn(a - back.x, b - back.y);
}
// 3-D tiled TMA store: dense (swizzled) smem box -> box at (c0, c1, c2) of the tensor; out-of-bounds parts are not written
__device__ __forceinline__ void tma_store_3d(const void* tmap, const void* smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
                 "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct SweepMaps {
    CUtensorMap in;   // [planes][I][K] fp32, box {128 k, 32 i, 1}
    CUtensorMap out;  // [planes][K][I] fp32, box {32 i, 128 j, 1}, SWIZZLE_128B
};

__global__ void __launch_bounds__(kSwCta, 1) blur_axis_sweep_kernel(const AxisArgs a, const __grid_constant__ SweepMaps maps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* s_g = reinterpret_cast<float*>(smem + kSwOffX);  // taps, only until the strip of core matrices is built
    float* s_scale = reinterpret_cast<float*>(smem + kSwOffMisc);  // [128] per input row of the strip being converted
    float* s_inv = s_scale + kB;                                   // [2][128]: strip parity (read by the store warps)
    uint64_t* mbar_tma = reinterpret_cast<uint64_t*>(smem + kSwOffMisc + 1536);  // [3]: ring slot filled
    uint64_t* mbar_mma = mbar_tma + kSwRing;           // [2]: the MMAs reading X[set] are complete
    uint64_t* mbar_acc_full = mbar_mma + 2;            // [4]: accumulator slot holds a finished tile
    uint64_t* mbar_acc_free = mbar_acc_full + kSwSlots;  // [4]: the store warps have read the slot
    uint64_t* mbar_strip_done = mbar_acc_free + kSwSlots;  // [2]: the store warps are done with a strip (its s_inv)
    double* s_red = reinterpret_cast<double*>(mbar_strip_done + 2);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(mbar_strip_done + 3);
    const uint32_t sbase = smem_u32(smem);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int first = blockIdx.x, stride = gridDim.x;
    const int TJ = a.tiles_j;
    const int n_strips = a.planes * a.tiles_i;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        tma_prefetch_desc(&maps.in);
        tma_prefetch_desc(&maps.out);
        for (int b = 0; b < kSwRing; ++b) mbar_init(mbar_tma + b, 1);
        mbar_init(mbar_mma + 0, 1);
        mbar_init(mbar_mma + 1, 1);
        for (int b = 0; b < kSwSlots; ++b) {
            mbar_init(mbar_acc_full + b, 1);
            mbar_init(mbar_acc_free + b, kSwStoreWarps);
        }
        mbar_init(mbar_strip_done + 0, kSwStoreWarps);
        mbar_init(mbar_strip_done + 1, kSwStoreWarps);
        mbar_fence_init();
    }
    if (warp == kSwComputeWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (a.rule == BGAN_RULE_EXPLICIT) asm volatile("griddepcontrol.wait;" ::: "memory");
    normalised_taps(a.rule, a.std_eff, a.radius, a.taps_dev, s_g, s_red);  // (ends with __syncthreads)
    {   // the strip of core matrices (as in the tile kernel, nb = 1), taps scaled by 2^11
        const int n_taps = 2 * a.radius + 1, mats = 32 + 31;
        for (int j = tid; j < mats * 32; j += kSwCta) {
            const int m = j >> 5, rho = (j >> 2) & 7, e = j & 3;
            const int ta = 8 * (m - 16 - 15) - rho + a.radius + 2 * e, tb = ta + 1;
            __half2 hi, lo;
            sw_split_tap2((ta >= 0 && ta < n_taps) ? s_g[ta] : 0.0f, (tb >= 0 && tb < n_taps) ? s_g[tb] : 0.0f, hi, lo);
            *reinterpret_cast<__half2*>(smem + kSwOffS1 + m * 128 + rho * 16 + e * 4) = hi;
            *reinterpret_cast<__half2*>(smem + kSwOffS2 + m * 128 + rho * 16 + e * 4) = lo;
        }
    }
    fence_proxy_async();
    ax_fence_before();
    __syncthreads();
    ax_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int jl = ((warp & 3) * 32 + lane) ^ kARev;  // output row (inside the tile) held by this lane's accumulator row
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == kSwComputeWarps) {
        // ================= MMA issuer =================================================================
        int u = 0, base = 0;  // block counter; running index of the strip's tile 0
        int ns = 0;
        for (int strip = first; strip < n_strips; strip += stride, base += TJ) {
            for (int L = 0; L < TJ; ++L, ++u) {
                named_bar_sync(1 + (u & 1), kSwCompute + 32);  // X[u & 1] of block u written
                if (lane == 0) {
                    ax_stamp(a.trace, 1, ns);  // block u: operands ready
                    ax_fence_after();
                    const uint32_t xb = sbase + kSwOffX + (u & 1) * 2 * kATile;
                    const uint64_t b1 = ax_desc_sw128(xb), b2 = ax_desc_sw128(xb + kATile);
#pragma unroll 1
                    for (int d = 1; d >= -1; --d) {  // tile J = L - d: d = +1 finishes J = L - 1, d = -1 starts J = L + 1
                        const int J = L - d;
                        if (J < 0 || J >= TJ) continue;
                        const int tj = base + J, slot = tj & (kSwSlots - 1);
                        const bool starts = L == (J > 0 ? J - 1 : 0);
                        if (starts && tj >= kSwSlots) mbar_wait(mbar_acc_free + slot, ((tj >> 2) - 1) & 1);  // slot drained
                        if (starts) ax_fence_after();
                        int k0, k1;
                        ks_range(d + 1, 1, a.radius, k0, k1);
                        const uint64_t a1 = ax_desc(sbase + kSwOffS1 + (d + 1) * 16 * 128, 128, 128);
                        const uint64_t a2 = ax_desc(sbase + kSwOffS2 + (d + 1) * 16 * 128, 128, 128);
                        const uint32_t acc = tmem + slot * kB;
                        for (int k = k0; k <= k1; ++k) {
                            ax_mma(acc, a1 + k * 16, b1 + ax_kstep(k), !(starts && k == k0));
                            ax_mma(acc, a1 + k * 16, b2 + ax_kstep(k), 1);
                            ax_mma(acc, a2 + k * 16, b1 + ax_kstep(k), 1);
                        }
                    }
                    ax_commit(mbar_mma + (u & 1));
                    if (L >= 1) ax_commit(mbar_acc_full + ((base + L - 1) & (kSwSlots - 1)));  // tile L - 1 complete
                    if (L == TJ - 1) ax_commit(mbar_acc_full + ((base + L) & (kSwSlots - 1)));  // ... and the last one
                    ax_stamp(a.trace, 1, ns);  // block u: issued
                }
                __syncwarp();
            }
        }
    } else if (warp > kSwComputeWarps) {
        // ================= store warps: out[j][i] = D 2^-11 / scale_i, through swizzled boxes and TMA stores ===========
        int base = 0, sl = 0;  // running tile index of the strip's tile 0; strips done by this CTA
        int ns = 0;
        const int sw_id = warp - kSwComputeWarps - 1;          // 0 .. 7
        const int box = sw_id >> 2;                            // which of the two output boxes this warp fills
        const bool issuer = sw_id == 0 && lane == 0;
        const bool tracer = issuer;
        unsigned char* obox = smem + kSwOffOut + box * kSwOutBox;
        const uint32_t orow = static_cast<uint32_t>(jl) * 128;  // this lane's row of the box (rows are output rows)
        for (int strip = first; strip < n_strips; strip += stride, base += TJ, ++sl) {
            const int plane = strip / a.tiles_i, ib = strip - plane * a.tiles_i;
            const float* inv = s_inv + (sl & 1) * kB;
            for (int J = 0; J < TJ; ++J) {
                const int tj = base + J, slot = tj & (kSwSlots - 1);
                mbar_wait(mbar_acc_full + slot, (tj >> 2) & 1);
                ax_fence_after();
                if (tracer) ax_stamp(a.trace, 2, ns);  // tile complete in TMEM
                const uint32_t acc = tmem + slot * kB;
                const int j = J * kB + jl;
                float mx = 0.0f;
#pragma unroll 1
                for (int round = 0; round < 2; ++round) {  // column groups {0, 1}, then {2, 3}: one per box
                    const int cg = 2 * round + box;
                    float dm[32];
                    ax_ld32(acc + lane_base + cg * 32, dm);
                    ax_ld_wait(dm);
                    if (round == 1) {  // this warp's part of the slot has been read
                        ax_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(mbar_acc_free + slot);
                    }
#pragma unroll
                    for (int e = 0; e < 32; ++e) {
                        dm[e] = dm[e] * (1.0f / 2048.0f) * inv[cg * 32 + e];
                        mx = fmaxf(mx, fabsf(dm[e]));
                    }
                    if (issuer) tma_store_wait_read();               // the boxes' previous contents have left
                    named_bar_sync(4, kSwStoreThreads);
#pragma unroll
                    for (int c = 0; c < 8; ++c)
                        *reinterpret_cast<float4*>(obox + orow + ((c ^ (jl & 7)) << 4)) = make_float4(dm[4 * c], dm[4 * c + 1], dm[4 * c + 2], dm[4 * c + 3]);
                    fence_proxy_async();
                    named_bar_sync(4, kSwStoreThreads);
                    if (issuer) {
                        tma_store_3d(&maps.out, smem + kSwOffOut, ib * kB + 2 * round * 32, J * kB, plane);
                        tma_store_3d(&maps.out, smem + kSwOffOut + kSwOutBox, ib * kB + (2 * round + 1) * 32, J * kB, plane);
                        tma_store_commit();
                    }
                }
                if (j < a.K && a.rowmax_out != nullptr)
                    atomicMax(reinterpret_cast<int*>(a.rowmax_out + static_cast<size_t>(plane) * a.K + j), __float_as_int(mx));
                if (tracer) ax_stamp(a.trace, 2, ns);  // tile handed to the TMA
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(mbar_strip_done + (sl & 1));  // s_inv[sl & 1] may be replaced (two strips on)
        }
        if (issuer) tma_store_wait_all();
    } else {
        // ================= conversion warps ===========================================================
        // thread 0 walks the quarter blocks (chunks of 32 rows): (strip, L, quarter) to load into the ring / to pull into L2
        int ld_strip = first, ld_q = 0, ld_c = 0, pf_strip = first, pf_q = 0;  // *_q = 4 L + quarter; ld_c = chunks issued
        auto issue_load = [&]() {
            if (ld_strip >= n_strips) return;
            const int plane = ld_strip / a.tiles_i, ib = ld_strip - plane * a.tiles_i;
            const int slot = ld_c % kSwRing;
            mbar_arrive_expect_tx(mbar_tma + slot, kSwChunk);
            tma_load_3d(smem + slot * kSwChunk, &maps.in, mbar_tma + slot, (ld_q >> 2) * kB, ib * kB + (ld_q & 3) * kSwChunkRows, plane);
            ++ld_c;
            if (++ld_q == 4 * TJ) {
                ld_q = 0;
                ld_strip += stride;
            }
        };
        auto issue_prefetch = [&]() {
            if (pf_strip >= n_strips) return;
            const int plane = pf_strip / a.tiles_i, ib = pf_strip - plane * a.tiles_i;
            tma_prefetch_l2_3d(&maps.in, (pf_q >> 2) * kB, ib * kB + (pf_q & 3) * kSwChunkRows, plane);
            if (++pf_q == 4 * TJ) {
                pf_q = 0;
                pf_strip += stride;
            }
        };
        if (tid == 0) {
            for (int i = 0; i < kSwRing; ++i) issue_load();   // the ring
            for (int i = 0; i < 12; ++i) issue_prefetch();    // three blocks ahead
        }
        int u = 0, sl = 0, ch = 0;  // blocks, strips, chunks consumed
        int ns = 0;
        for (int strip = first; strip < n_strips; strip += stride, ++sl) {
            const int plane = strip / a.tiles_i, ib = strip - plane * a.tiles_i;
            named_bar_sync(3, kSwCompute);  // every warp has converted the previous strip's last block: s_scale may change
            if (sl >= 2) mbar_wait(mbar_strip_done + (sl & 1), ((sl - 2) >> 1) & 1);  // s_inv[sl & 1] no longer read
            if (tid < kB) {
                const int i = ib * kB + tid;
                const float m = i < a.I ? __ldg(a.rowmax_in + static_cast<size_t>(plane) * a.I + i) : 0.0f;
                int eb = (__float_as_int(m) >> 23) & 0xff;
                eb = min(max(eb, 13), 254);
                s_scale[tid] = __int_as_float((266 - eb) << 23);  // max |row| -> [2^12, 2^13)
                s_inv[(sl & 1) * kB + tid] = __int_as_float((eb - 12) << 23);
            }
            named_bar_sync(3, kSwCompute);
            for (int L = 0; L < TJ; ++L, ++u) {
                unsigned char* x1 = smem + kSwOffX + (u & 1) * 2 * kATile;
                unsigned char* x2 = x1 + kATile;
#pragma unroll 1
                for (int qt = 0; qt < 4; ++qt, ++ch) {  // rows 32 qt .. 32 qt + 31 of the block
                    const int slot = ch % kSwRing;
                    const float* stage = reinterpret_cast<const float*>(smem + slot * kSwChunk);
                    if (tid == 0 && qt == 0) ax_stamp(a.trace, 0, ns);  // waiting for the block's first chunk
                    mbar_wait(mbar_tma + slot, (ch / kSwRing) & 1);
                    if (tid == 0 && qt == 0) ax_stamp(a.trace, 0, ns);  // landed
                    float4 v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = *reinterpret_cast<const float4*>(stage + (q * kSwCompute + tid) * 4);
                    // scale before the barrier: the multiplications consume the loaded registers, so the LDS have really
                    // returned when the slot is released to the next TMA load (see the tile kernel)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float sc = s_scale[qt * kSwChunkRows + ((q * kSwCompute + tid) >> 5)];
                        v[q].x *= sc, v[q].y *= sc, v[q].z *= sc, v[q].w *= sc;
                        asm volatile("" ::"f"(v[q].x), "f"(v[q].y), "f"(v[q].z), "f"(v[q].w));  // (pins the products in front of the barrier)
                    }
                    named_bar_sync(3, kSwCompute);  // ring slot drained
                    if (tid == 0) {
                        issue_load();      // the chunk three further on, into this slot
                        issue_prefetch();  // three blocks ahead -> L2
                    }
                    if (qt == 0) {
                        if (u >= 2) mbar_wait(mbar_mma + (u & 1), ((u - 2) >> 1) & 1);  // MMAs of block u - 2 complete: X[u & 1] free
                        if (tid == 0) ax_stamp(a.trace, 0, ns);  // operand set free
                    }
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int f = q * kSwCompute + tid;
                        const int h = qt * kSwChunkRows + (f >> 5), w = (f & 31) << 2;
                        __half2 hi[2], lo[2];
                        sw_split_raw2(v[q].x, v[q].y, hi[0], lo[0]);
                        sw_split_raw2(v[q].z, v[q].w, hi[1], lo[1]);
                        const uint32_t off = ax_tile_off(h, w);
                        *reinterpret_cast<uint2*>(x1 + off) = *reinterpret_cast<uint2*>(hi);
                        *reinterpret_cast<uint2*>(x2 + off) = *reinterpret_cast<uint2*>(lo);
                    }
                }
                fence_proxy_async();
                ax_fence_before();
                named_bar_arrive(1 + (u & 1), kSwCompute + 32);  // -> MMAs of block u
                if (tid == 0) ax_stamp(a.trace, 0, ns);  // block u converted
            }
        }
    }
    ax_fence_before();
    __syncthreads();
    if (warp == kSwComputeWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn axis_encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// one pass: in [planes][I][K] -> out [planes][K][I]
int launch_axis(const float* in, float* out, const float* rowmax_in, float* rowmax_out, int planes, int I, int K, int rule, float std_eff,
                int radius, const float* taps_dev, cudaStream_t stream) {
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    thread_local bool bound = false;
    if (!bound) {
        cudaFree(nullptr);
        bound = true;
    }
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(I), static_cast<cuuint64_t>(planes)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(K) * 4, static_cast<cuuint64_t>(K) * I * 4};
    const cuuint32_t box[3] = {kB, kB, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult res = axis_encode_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(in), dims, strides, box, estr,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BGAN_REQUIRE(res == CUDA_SUCCESS, BGAN_E_UNSUPPORTED, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(res));
    thread_local int attr_dev = -1;
    int dev = 0;
    BGAN_CUDA_OK(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        BGAN_CUDA_OK(cudaFuncSetAttribute(blur_axis_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kASmem)));
        BGAN_CUDA_OK(cudaFuncSetAttribute(blur_axis_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSwSmem)));
        attr_dev = dev;
    }
    AxisArgs a{};
    a.out = out;
    a.rowmax_in = rowmax_in;
    a.rowmax_out = rowmax_out;
    a.I = I;
    a.K = K;
    a.planes = planes;
    a.tiles_i = ceil_div(I, kB);
    a.tiles_j = ceil_div(K, kB);
    a.nb = ceil_div(radius, kB);
    a.rule = rule;
    a.std_eff = std_eff;
    a.radius = radius;
    a.taps_dev = taps_dev;
    const char* tr = getenv("BGAN_AXIS_TRACE_PTR");  // debug only: device pointer (decimal) to 768 int64
    a.trace = tr ? reinterpret_cast<long long*>(strtoull(tr, nullptr, 10)) : nullptr;
    const int64_t n_tiles = static_cast<int64_t>(planes) * a.tiles_i * a.tiles_j;
    BGAN_REQUIRE(n_tiles < (int64_t(1) << 28), BGAN_E_BADARG, "blur_axis: too many tiles");
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.blockDim = dim3(kAxCta);
    cfg.dynamicSmemBytes = kASmem;
    cfg.stream = stream;
    const char* no_sweep = getenv("BGAN_BLUR_AXIS_NO_SWEEP");
    if (a.nb == 1 && a.tiles_j >= 2 && !(no_sweep && atoi(no_sweep))) {
        SweepMaps maps;
        memset(&maps, 0, sizeof(maps));
        const cuuint32_t box_in[3] = {kB, kSwChunkRows, 1};
        CUresult r2 = axis_encode_fn()(&maps.in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(in), dims, strides, box_in, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        BGAN_REQUIRE(r2 == CUDA_SUCCESS, BGAN_E_UNSUPPORTED, "cuTensorMapEncodeTiled (sweep input) failed (%d)", static_cast<int>(r2));
        const cuuint64_t odims[3] = {static_cast<cuuint64_t>(I), static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(planes)};
        const cuuint64_t ostrides[2] = {static_cast<cuuint64_t>(I) * 4, static_cast<cuuint64_t>(K) * I * 4};
        const cuuint32_t box_out[3] = {32, kB, 1};
        r2 = axis_encode_fn()(&maps.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, odims, ostrides, box_out, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        BGAN_REQUIRE(r2 == CUDA_SUCCESS, BGAN_E_UNSUPPORTED, "cuTensorMapEncodeTiled (sweep output) failed (%d)", static_cast<int>(r2));
        // strips of 128 input rows; balanced persistent grid (every CTA the same number of strips where possible)
        const int64_t strips = static_cast<int64_t>(planes) * a.tiles_i;
        const int64_t rounds = ceil_div64(strips, device_sm_count());
        cfg.gridDim = dim3(static_cast<unsigned>(ceil_div64(strips, rounds)));
        cfg.blockDim = dim3(kSwCta);
        cfg.dynamicSmemBytes = kSwSmem;
        BGAN_CUDA_OK(cudaLaunchKernelEx(&cfg, blur_axis_sweep_kernel, a, maps));
        return 0;
    }
    cfg.gridDim = dim3(static_cast<unsigned>(std::min<int64_t>(n_tiles, device_sm_count())));
    BGAN_CUDA_OK(cudaLaunchKernelEx(&cfg, blur_axis_tc_kernel, a, tmap));
    return 0;
}

}  // namespace

// planes of H x W floats, contiguous (NCHW, or NHWC with one channel).  Shape part of the eligibility: also decides the
// workspace size, so it must not depend on the device or on pointer values.
bool blur_axis_shape_ok(int planes, int H, int W, int ES, int radius) {
    const char* off = getenv("BGAN_BLUR_NO_AXIS_TC");
    if (off && atoi(off)) return false;
    const char* min_r = getenv("BGAN_BLUR_AXIS_MIN_R");
    const int rmin = min_r ? atoi(min_r) : 12;
    return ES == 1 && planes > 0 && H >= 256 && W >= 256 && H % 4 == 0 && W % 4 == 0 && radius >= rmin && radius <= kAxMaxNb * kB;
}

bool blur_axis_eligible(int planes, int H, int W, int ES, int radius, const float* x, const float* y) {
    return blur_axis_shape_ok(planes, H, W, ES, radius) && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
           reinterpret_cast<uintptr_t>(y) % 16 == 0 && axis_encode_fn() != nullptr;
}

// scratch: the transposed intermediate plus the two row-maximum arrays
size_t blur_axis_workspace_bytes(int planes, int H, int W) {
    const size_t mid = (static_cast<size_t>(planes) * H * W * sizeof(float) + 255) & ~size_t(255);
    const size_t rm = (static_cast<size_t>(planes) * (H + W) * sizeof(float) + 255) & ~size_t(255);
    return mid + rm;
}

int blur_axis_launch(const float* x, float* y, int planes, int H, int W, int rule, float std_eff, int radius, const float* taps_dev, void* ws,
                     size_t ws_bytes, cudaStream_t stream) {
    BGAN_REQUIRE(ws && ws_bytes >= blur_axis_workspace_bytes(planes, H, W) && reinterpret_cast<uintptr_t>(ws) % 16 == 0, BGAN_E_WORKSPACE,
                 "blur_axis: workspace of %zu bytes required, got %zu", blur_axis_workspace_bytes(planes, H, W), ws_bytes);
    const size_t mid_bytes = (static_cast<size_t>(planes) * H * W * sizeof(float) + 255) & ~size_t(255);
    float* mid = static_cast<float*>(ws);                                          // Z^T [planes][W][H]
    float* rm_x = reinterpret_cast<float*>(static_cast<char*>(ws) + mid_bytes);   // [planes][H]
    float* rm_z = rm_x + static_cast<size_t>(planes) * H;                         // [planes][W]
    const int64_t rows = static_cast<int64_t>(planes) * H;
    rowmax_kernel<<<static_cast<unsigned>(ceil_div64(rows * 32, 256)), 256, 0, stream>>>(x, rows, W, rm_x);
    BGAN_LAUNCH_OK();
    BGAN_CUDA_OK(cudaMemsetAsync(rm_z, 0, static_cast<size_t>(planes) * W * sizeof(float), stream));
    int rc = launch_axis(x, mid, rm_x, rm_z, planes, H, W, rule, std_eff, radius, taps_dev, stream);  // blur along W: X[h][w] -> Z^T[w'][h]
    if (rc) return rc;
    return launch_axis(mid, y, rm_z, nullptr, planes, W, H, rule, std_eff, radius, taps_dev, stream);  // blur along H: Z^T[w'][h] -> Y[h'][w']
}

}  // namespace bgan

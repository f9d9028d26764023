// Tensor-core blur for planes larger than 128 x 128 (BASELINE configs[3]: 1024 x 1024, radius up to 192).
//
// One operation, applied twice:   out[p][j][i] = sum_k T[j][k] in[p][i][k],   T[j][k] = tap[k - j + r]  (|k - j| <= r)
// i.e. "blur along the contiguous axis and write the result transposed".  Pass 1 turns X[h][w] into Z^T[w'][h]
// (scratch), pass 2 turns Z^T[w'][h] into Y[h'][w']: both passes read rows that are contiguous along the axis they
// blur -- the natural K-major B operand -- and both write 128-byte (64-byte) row segments.  With the CUDA-core tile kernel the
// wide blurs of configs[3] cost 2 (2r + 1) = 770 FMAs per pixel; here the cost per 128 x 128 output tile is at most
// (2 nb + 1) x 24 tcgen05.mma, nb = ceil(r / 128) <= 2.
//
// Operands.  Output tile = 128 outputs j (accumulator rows) x 128 input rows i (accumulator columns), accumulated over the
// blocks of 128 k that the band of T touches.  A = T block (offset d = L - J): core matrices of a Toeplitz block only depend
// on 16 d + kchunk - rowgroup, so all 2 nb + 1 blocks are views (LBO = SBO = 128 B descriptors, row groups reversed) of one
// strip of 32 nb + 31 core matrices.  B = the fp32 block of `in`, staged by TMA (out-of-bounds zero fill = the reference's
// zero padding), scaled per row by a power of two (row maxima come from a tiny pre-pass for pass 1 and from pass 1's
// epilogue for pass 2) and split into fp16 hi | lo.  fp32 parity: the taps are stored scaled by 2^11 (T' = 2^11 T,
// hi = fp16(T'), lo = fp16(T' - hi): the remainder of even a tiny tap stays an fp16 normal), the data remainders unscaled (rows
// are scaled to [2^12, 2^13)), so T' X ~= T'_hi X_hi + T'_hi X_lo + T'_lo X_hi accumulates in ONE TMEM accumulator in any
// order and the epilogue multiplies by 2^-11 / scale_i.  K-steps of a neighbouring block that only meet zeros of T are
// neither converted nor issued.
//
// Schedule, one set of roles -- 8 conversion warps, the MMA-issuing warp, 8 store warps.  A work unit is a run of adjacent
// output tiles J0 .. J0 + nt - 1 of one strip (128 input rows i of a plane); the CTA walks the input blocks L that the band
// reaches from those tiles, loads and converts every block ONCE and multiplies it into every tile of the unit with
// |L - J| <= nb.  Four accumulator slots of 128 TMEM columns:
//   radius <= 128 (nb = 1): the unit is the whole strip (input-stationary sweep): a block touches J = L - 1 (which it
//     completes), L and L + 1 (which it starts) -- three live accumulators, one draining; every block crosses L2 -> SM once.
//   radius 129 .. 256 (nb = 2): a whole strip would need five live accumulators (640 TMEM columns), so units are PAIRS of
//     tiles (two live, two draining): six block loads per two tiles instead of five per tile.
// What the measurements said on the way (profiles/r02_axis.md): round 1's kernel loaded every block once per tile with one
// 64 KB TMA box in flight (a latency chain: DRAM 13 %, tensor pipe 18 %); direct stores of accumulator rows are 32 lines per
// instruction, 4096 sixteen-byte requests per tile, which paced the kernel and starved the conversion warps' LDS / STS in the
// shared LSU.  Hence: a ring of three 16 KB input chunks, and the output through swizzled boxes in shared memory and TMA
// stores (cp.async.bulk.tensor global <- shared), off the LSU.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "taps.cuh"

namespace bgan {
namespace {

constexpr int kB = 128;                        // block edge
constexpr int kAxMaxNb = 2;                    // radius <= 256
constexpr int kAxCompute = 256;                // conversion threads (8 warps)
constexpr int kAxComputeWarps = kAxCompute / 32;
constexpr int kAxStoreWarps = 8;               // two per TMEM lane quadrant
constexpr int kAxStoreThreads = 32 * kAxStoreWarps;
constexpr int kAxCta = kAxCompute + 32 + kAxStoreThreads;  // 544
constexpr int kAxSlots = 4;                    // accumulators of 128 TMEM columns
// B operand (fp16 hi | lo of a 128 x 128 block of `in`): K-major tiles with the 128-byte hardware swizzle, as in
// blur_tc.cu (two K-blocks of 64 elements; row r owns 128 contiguous bytes per K-block, 16-byte chunk c at c ^ (r % 8)).
// Two sets, so that the conversion of block u + 1 overlaps the MMAs of block u.
constexpr uint32_t kAKBlock = kB * 128;        // 16384: 128 rows x 64 K-elements
constexpr uint32_t kATile = 2 * kAKBlock;      // 32768
// input staging: a ring of three chunks of 32 rows (three TMA boxes in flight hide the L2-hit latency of a box)
constexpr int kAxRing = 3;
constexpr int kAxChunkRows = 32;
constexpr uint32_t kAxChunk = kAxChunkRows * kB * 4;  // 16384
constexpr uint32_t kAIdesc = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);  // kind::f16, D = f32, M = N = 128, K-major
constexpr int kARev = 0x78;

// Shared memory: ring | X[2][hi | lo] | two output boxes | strips hi, lo | scales, barriers.
//   nb = 1: output boxes of 32 columns (128 rows x 128 B, SWIZZLE_128B), strips of 63 core matrices
//   nb = 2 (WIDE): output boxes of 16 columns (128 rows x 64 B, SWIZZLE_64B), strips of 95 core matrices
template <bool WIDE>
struct AxLayout {
    static constexpr int kBoxCols = WIDE ? 16 : 32;
    static constexpr uint32_t kBox = kB * kBoxCols * 4;
    static constexpr uint32_t kStrip = WIDE ? 12288 : 8192;  // (32 nb + 31) x 128 B
    static constexpr uint32_t kOffRing = 0;
    static constexpr uint32_t kOffX = kAxRing * kAxChunk;      // the raw taps live here until the first conversion
    static constexpr uint32_t kOffOut = kOffX + 4 * kATile;
    static constexpr uint32_t kOffS1 = kOffOut + 2 * kBox;
    static constexpr uint32_t kOffS2 = kOffS1 + kStrip;
    static constexpr uint32_t kOffMisc = kOffS2 + kStrip;      // s_scale[128], s_inv[4][128], barriers
    static constexpr uint32_t kSmem = kOffMisc + 2560 + 256;
    static_assert(kSmem <= 227 * 1024, "axis kernel: shared memory");
};

struct AxisArgs {
    const float* rowmax_in;   // [planes][I]: max |in| over each row
    float* rowmax_out;        // [planes][K] or nullptr: max |out| over each output row (accumulated with atomicMax)
    int I, K, planes;
    int tiles_i, tiles_j;
    int unit_tiles;           // output tiles per work unit (nb = 1: tiles_j, the whole strip; nb = 2: 2)
    int units_per_strip;      // ceil(tiles_j / unit_tiles)
    int n_units;              // planes * tiles_i * units_per_strip
    int nb;                   // ceil(radius / 128)
    int rule;
    float std_eff;
    int radius;
    const float* taps_dev;
    long long* trace;         // debug: clock64 stamps of CTA 0 (nullptr in production), tools/trace_axis.py
};
struct AxisMaps {
    CUtensorMap in;   // [planes][I][K] fp32, box {128 k, 32 i, 1}
    CUtensorMap out;  // [planes][K][I] fp32, box {32 | 16 i, 128 j, 1}, SWIZZLE_128B | SWIZZLE_64B
};

// trace layout: three timelines of 256 int64 each (conversion thread 0, MMA lane, lane 0 of the first store warp);
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
// N consecutive fp32 columns of this thread's TMEM lane (N = 32 or 16), valid on return
template <int N>
__device__ __forceinline__ void ax_ld(uint32_t taddr, float (&v)[N]);
template <>
__device__ __forceinline__ void ax_ld<32>(uint32_t taddr, float (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
          "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]),
          "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]),
          "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]),
                   "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]), "+f"(v[16]), "+f"(v[17]), "+f"(v[18]),
                   "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]), "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]),
                   "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31])
                 :
                 : "memory");
}
template <>
__device__ __forceinline__ void ax_ld<16>(uint32_t taddr, float (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
          "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]),
                   "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
                 :
                 : "memory");
}
// taps: t -> T' = 2^11 t -> (fp16(T'), fp16(T' - fp16(T')))
__device__ __forceinline__ void ax_split_tap2(float a, float b, __half2& hi, __half2& lo) {
    a *= 2048.0f;
    b *= 2048.0f;
    hi = __floats2half2_rn(a, b);
    const float2 back = __half22float2(hi);
    lo = __floats2half2_rn(a - back.x, b - back.y);
}
// data (already scaled to [2^12, 2^13) per row): remainder unscaled
__device__ __forceinline__ void ax_split_raw2(float a, float b, __half2& hi, __half2& lo) {
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

// K-steps (16 inputs k each) of the block at offset d = L - J that the band |k - j| <= r of a 128-output tile touches: the
// centre block entirely, a block to the left only its last columns (k_local >= 128 |d| - r), a block to the right only its
// first ones (k_local <= r - 1 - 128 (d - 1)).
__device__ __forceinline__ void ks_range(int d, int radius, int& k0, int& k1) {
    k0 = 0;
    k1 = 7;
    if (d < 0) k0 = max(0, -d * kB - radius) >> 4;
    if (d > 0) k1 = min(7, (radius - 1 - (d - 1) * kB) >> 4);
}

// A work unit: output tiles J0 .. J0 + nt - 1 of strip (plane, ib); input blocks L0 .. L1 (those the band reaches, inside the array)
struct Unit {
    int plane, ib, J0, nt, L0, L1;
};
__device__ __forceinline__ Unit unit_of(const AxisArgs& a, int idx) {
    Unit u;
    const int g = idx % a.units_per_strip, strip = idx / a.units_per_strip;
    u.plane = strip / a.tiles_i;
    u.ib = strip - u.plane * a.tiles_i;
    u.J0 = g * a.unit_tiles;
    u.nt = min(a.unit_tiles, a.tiles_j - u.J0);
    u.L0 = max(0, u.J0 - a.nb);
    u.L1 = min(a.tiles_j - 1, u.J0 + u.nt - 1 + a.nb);
    return u;
}
// first / last input block that contributes to output tile J
__device__ __forceinline__ int first_block(const AxisArgs& a, int J) { return max(0, J - a.nb); }
__device__ __forceinline__ int last_block(const AxisArgs& a, int J) { return min(a.tiles_j - 1, J + a.nb); }

template <bool WIDE>
__global__ void __launch_bounds__(kAxCta, 1) blur_axis_kernel(const AxisArgs a, const __grid_constant__ AxisMaps maps) {
    using L = AxLayout<WIDE>;
    extern __shared__ __align__(1024) unsigned char smem[];
    float* s_g = reinterpret_cast<float*>(smem + L::kOffX);         // taps (<= 513), only until the strip is built
    float* s_scale = reinterpret_cast<float*>(smem + L::kOffMisc);  // [128] per input row of the unit being converted
    float* s_inv = s_scale + kB;                                    // [4][128]: ring by unit (read by the store warps)
    uint64_t* mbar_tma = reinterpret_cast<uint64_t*>(smem + L::kOffMisc + 2560);  // [3]: ring slot filled
    uint64_t* mbar_mma = mbar_tma + kAxRing;               // [2]: the MMAs reading X[set] are complete
    uint64_t* mbar_acc_full = mbar_mma + 2;                // [4]: accumulator slot holds a finished tile
    uint64_t* mbar_acc_free = mbar_acc_full + kAxSlots;    // [4]: the store warps have read the slot
    uint64_t* mbar_unit_done = mbar_acc_free + kAxSlots;   // [4]: the store warps are done with a unit (its s_inv)
    double* s_red = reinterpret_cast<double*>(mbar_unit_done + 4);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(mbar_unit_done + 5);
    const uint32_t sbase = smem_u32(smem);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int first = blockIdx.x, stride = gridDim.x;
    const int n_units = a.n_units;

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        tma_prefetch_desc(&maps.in);
        tma_prefetch_desc(&maps.out);
        for (int b = 0; b < kAxRing; ++b) mbar_init(mbar_tma + b, 1);
        mbar_init(mbar_mma + 0, 1);
        mbar_init(mbar_mma + 1, 1);
        for (int b = 0; b < kAxSlots; ++b) {
            mbar_init(mbar_acc_full + b, 1);
            mbar_init(mbar_acc_free + b, kAxStoreWarps);
            mbar_init(mbar_unit_done + b, kAxStoreWarps);
        }
        mbar_fence_init();
    }
    if (warp == kAxComputeWarps) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (a.rule == BGAN_RULE_EXPLICIT) asm volatile("griddepcontrol.wait;" ::: "memory");
    normalised_taps(a.rule, a.std_eff, a.radius, a.taps_dev, s_g, s_red);  // (ends with __syncthreads)
    // ---- the strip: core matrix m = 16 (d + nb) + kchunk + rowgroup', row rho, holds taps t0 .. t0 + 7,
    //      t0 = 8 (m - 16 nb - 15) - rho + r  (zero outside [0, 2r]), scaled by 2^11 ----
    {
        const int n_taps = 2 * a.radius + 1, mats = 32 * a.nb + 31;
        for (int j = tid; j < mats * 32; j += kAxCta) {
            const int m = j >> 5, rho = (j >> 2) & 7, e = j & 3;
            const int ta = 8 * (m - 16 * a.nb - 15) - rho + a.radius + 2 * e, tb = ta + 1;
            __half2 hi, lo;
            ax_split_tap2((ta >= 0 && ta < n_taps) ? s_g[ta] : 0.0f, (tb >= 0 && tb < n_taps) ? s_g[tb] : 0.0f, hi, lo);
            *reinterpret_cast<__half2*>(smem + L::kOffS1 + m * 128 + rho * 16 + e * 4) = hi;
            *reinterpret_cast<__half2*>(smem + L::kOffS2 + m * 128 + rho * 16 + e * 4) = lo;
        }
    }
    fence_proxy_async();
    ax_fence_before();
    __syncthreads();
    ax_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;  // TMEM lane quadrant of this warp
    const int jl = ((warp & 3) * 32 + lane) ^ kARev;  // output row (inside the tile) held by this lane's accumulator row
    asm volatile("griddepcontrol.wait;" ::: "memory");  // `in` may be written, `out` may still be read by the previous grid

    if (warp == kAxComputeWarps) {
        // ================= MMA issuer =================================================================
        int u = 0, base = 0;  // blocks issued; running index of the unit's first tile
        int ns = 0;
        auto issue = [&](int slot, int d, bool starts, const uint64_t b1, const uint64_t b2) {
            int k0, k1;
            ks_range(d, a.radius, k0, k1);
            const uint64_t a1 = ax_desc(sbase + L::kOffS1 + (d + a.nb) * 16 * 128, 128, 128);
            const uint64_t a2 = ax_desc(sbase + L::kOffS2 + (d + a.nb) * 16 * 128, 128, 128);
            const uint32_t acc = tmem + slot * kB;
            for (int k = k0; k <= k1; ++k) {
                ax_mma(acc, a1 + k * 16, b1 + ax_kstep(k), !(starts && k == k0));
                ax_mma(acc, a1 + k * 16, b2 + ax_kstep(k), 1);
                ax_mma(acc, a2 + k * 16, b1 + ax_kstep(k), 1);
            }
        };
        for (int idx = first; idx < n_units; idx += stride) {
            const Unit un = unit_of(a, idx);
            for (int Lb = un.L0; Lb <= un.L1; ++Lb, ++u) {
                named_bar_sync(1 + (u & 1), kAxCompute + 32);  // X[u & 1] of block u written
                if (lane == 0) {
                    ax_stamp(a.trace, 1, ns);  // block u: operands ready
                    ax_fence_after();
                    const uint32_t xb = sbase + L::kOffX + (u & 1) * 2 * kATile;
                    const uint64_t b1 = ax_desc_sw128(xb), b2 = ax_desc_sw128(xb + kATile);
                    // tiles of the unit within reach of the block, the one it completes first
                    const int Ja = max(un.J0, Lb - a.nb), Jb = min(un.J0 + un.nt - 1, Lb + a.nb);
#pragma unroll 1
                    for (int J = Ja; J <= Jb; ++J) {
                        const int tj = base + (J - un.J0), slot = tj & (kAxSlots - 1);
                        const bool starts = Lb == first_block(a, J);
                        if (starts && tj >= kAxSlots) mbar_wait(mbar_acc_free + slot, ((tj >> 2) - 1) & 1);  // slot drained
                        if (starts) ax_fence_after();
                        issue(slot, Lb - J, starts, b1, b2);
                    }
                    ax_commit(mbar_mma + (u & 1));
                    for (int J = Ja; J <= Jb; ++J)
                        if (Lb == last_block(a, J)) ax_commit(mbar_acc_full + ((base + (J - un.J0)) & (kAxSlots - 1)));  // tile J complete
                    ax_stamp(a.trace, 1, ns);  // block u: issued
                }
                __syncwarp();
            }
            base += un.nt;
        }
    } else if (warp > kAxComputeWarps) {
        // ================= store warps: out[j][i] = D 2^-11 / scale_i, through swizzled boxes and TMA stores ===========
        // An accumulator row is an output row.  Per round every warp writes its 32 rows x kBoxCols columns into one of two
        // boxes (16-byte chunks XOR-swizzled by the row, as the tensor map's swizzle mode expects: conflict-free STS.128),
        // one thread hands the two boxes to the TMA.
        constexpr int BC = L::kBoxCols;
        constexpr int kRounds = kB / (2 * BC);
        int base = 0, ul = 0;  // running tile index of the unit's first tile; units done by this CTA
        int ns = 0;
        const int sw_id = warp - kAxComputeWarps - 1;          // 0 .. 7
        const int box = sw_id >> 2;                            // which of the two output boxes this warp fills
        const bool issuer = sw_id == 0 && lane == 0;
        unsigned char* obox = smem + L::kOffOut + box * L::kBox;
        const uint32_t orow = static_cast<uint32_t>(jl) * (BC * 4);
        const uint32_t oxor = BC == 32 ? (jl & 7) : ((jl >> 1) & 3);  // SWIZZLE_128B: chunk ^ (row % 8); SWIZZLE_64B: chunk ^ ((row / 2) % 4)
        for (int idx = first; idx < n_units; idx += stride, ++ul) {
            const Unit un = unit_of(a, idx);
            const float* inv = s_inv + (ul & 3) * kB;
            for (int t = 0; t < un.nt; ++t) {
                const int J = un.J0 + t;
                const int tj = base + t, slot = tj & (kAxSlots - 1);
                mbar_wait(mbar_acc_full + slot, (tj >> 2) & 1);
                ax_fence_after();
                if (issuer) ax_stamp(a.trace, 2, ns);  // tile complete in TMEM
                const uint32_t acc = tmem + slot * kB;
                const int j = J * kB + jl;
                float mx = 0.0f;
#pragma unroll 1
                for (int round = 0; round < kRounds; ++round) {
                    const int c0 = (2 * round + box) * BC;  // first accumulator column (= input row i) of this warp's part
                    float dm[BC];
                    ax_ld<BC>(acc + lane_base + c0, dm);
#pragma unroll
                    for (int e = 0; e < BC; ++e) {
                        dm[e] = dm[e] * (1.0f / 2048.0f) * inv[c0 + e];
                        mx = fmaxf(mx, fabsf(dm[e]));
                    }
                    if (round == kRounds - 1) {  // this warp's part of the slot has been read
                        ax_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(mbar_acc_free + slot);
                    }
                    if (issuer) tma_store_wait_read();               // the boxes' previous contents have left
                    named_bar_sync(4, kAxStoreThreads);
#pragma unroll
                    for (int c = 0; c < BC / 4; ++c)
                        *reinterpret_cast<float4*>(obox + orow + ((c ^ oxor) << 4)) = make_float4(dm[4 * c], dm[4 * c + 1], dm[4 * c + 2], dm[4 * c + 3]);
                    fence_proxy_async();
                    named_bar_sync(4, kAxStoreThreads);
                    if (issuer) {
                        tma_store_3d(&maps.out, smem + L::kOffOut, un.ib * kB + 2 * round * BC, J * kB, un.plane);
                        tma_store_3d(&maps.out, smem + L::kOffOut + L::kBox, un.ib * kB + (2 * round + 1) * BC, J * kB, un.plane);
                        tma_store_commit();
                    }
                }
                if (j < a.K && a.rowmax_out != nullptr)
                    atomicMax(reinterpret_cast<int*>(a.rowmax_out + static_cast<size_t>(un.plane) * a.K + j), __float_as_int(mx));
                if (issuer) ax_stamp(a.trace, 2, ns);  // tile handed to the TMA
            }
            base += un.nt;
            __syncwarp();
            if (lane == 0) mbar_arrive(mbar_unit_done + (ul & 3));  // s_inv[ul & 3] may be replaced (four units on)
        }
        if (issuer) tma_store_wait_all();
    } else {
        // ================= conversion warps ===========================================================
        // thread 0 walks the chunks (32 rows of a block): next one to load into the ring / to pull into L2 (further ahead)
        struct Cursor {
            int idx, Lb, q;
            Unit un;
        };
        Cursor ld{}, pf{};
        int ld_c = 0;  // chunks issued
        auto init = [&](Cursor& c) {
            c.idx = first;
            c.un = unit_of(a, min(first, n_units - 1));
            c.Lb = c.un.L0;
            c.q = 0;
        };
        auto advance = [&](Cursor& c) {
            if (++c.q < 4) return;
            c.q = 0;
            if (++c.Lb <= c.un.L1) return;
            c.idx += stride;
            if (c.idx < n_units) {
                c.un = unit_of(a, c.idx);
                c.Lb = c.un.L0;
            }
        };
        auto k_coord = [&](const Cursor& c) { return c.Lb * kB; };
        auto issue_load = [&]() {
            if (ld.idx >= n_units) return;
            const int slot = ld_c % kAxRing;
            mbar_arrive_expect_tx(mbar_tma + slot, kAxChunk);
            tma_load_3d(smem + L::kOffRing + slot * kAxChunk, &maps.in, mbar_tma + slot, k_coord(ld), ld.un.ib * kB + ld.q * kAxChunkRows, ld.un.plane);
            ++ld_c;
            advance(ld);
        };
        auto issue_prefetch = [&]() {
            if (pf.idx >= n_units) return;
            tma_prefetch_l2_3d(&maps.in, k_coord(pf), pf.un.ib * kB + pf.q * kAxChunkRows, pf.un.plane);
            advance(pf);
        };
        if (tid == 0) {
            init(ld);
            init(pf);
            for (int i = 0; i < kAxRing; ++i) issue_load();  // the ring
            for (int i = 0; i < 12; ++i) issue_prefetch();   // three blocks ahead
        }
        int u = 0, ul = 0, ch = 0;  // blocks, units, chunks consumed
        int ns = 0;
        for (int idx = first; idx < n_units; idx += stride, ++ul) {
            const Unit un = unit_of(a, idx);
            named_bar_sync(3, kAxCompute);  // every warp has converted the previous unit's last block: s_scale may change
            if (ul >= 4) mbar_wait(mbar_unit_done + (ul & 3), ((ul >> 2) - 1) & 1);  // s_inv[ul & 3] no longer read
            if (tid < kB) {
                const int i = un.ib * kB + tid;
                const float m = i < a.I ? __ldg(a.rowmax_in + static_cast<size_t>(un.plane) * a.I + i) : 0.0f;
                int eb = (__float_as_int(m) >> 23) & 0xff;
                eb = min(max(eb, 13), 254);
                s_scale[tid] = __int_as_float((266 - eb) << 23);  // max |row| -> [2^12, 2^13)
                s_inv[(ul & 3) * kB + tid] = __int_as_float((eb - 12) << 23);
            }
            named_bar_sync(3, kAxCompute);
            for (int Lb = un.L0; Lb <= un.L1; ++Lb, ++u) {
                // columns of the block that some tile of the unit needs (all of them when the block is a tile's centre)
                int k0 = 8, k1 = -1;
                for (int J = max(un.J0, Lb - a.nb); J <= min(un.J0 + un.nt - 1, Lb + a.nb); ++J) {
                    int ka, kb;
                    ks_range(Lb - J, a.radius, ka, kb);
                    k0 = min(k0, ka);
                    k1 = max(k1, kb);
                }
                const int w_lo = 16 * k0, w_hi = 16 * k1 + 16;
                unsigned char* x1 = smem + L::kOffX + (u & 1) * 2 * kATile;
                unsigned char* x2 = x1 + kATile;
#pragma unroll 1
                for (int qt = 0; qt < 4; ++qt, ++ch) {  // rows 32 qt .. 32 qt + 31 of the block
                    const int slot = ch % kAxRing;
                    const float* stage = reinterpret_cast<const float*>(smem + L::kOffRing + slot * kAxChunk);
                    if (tid == 0 && qt == 0) ax_stamp(a.trace, 0, ns);  // waiting for the block's first chunk
                    mbar_wait(mbar_tma + slot, (ch / kAxRing) & 1);
                    if (tid == 0 && qt == 0) ax_stamp(a.trace, 0, ns);  // landed
                    // The row scale is applied BEFORE the barrier that releases the ring slot to the next TMA load: the
                    // multiplications consume the loaded registers, so every LDS has really returned when the thread arrives.
                    // (A load that is only issued can still sit in the LSU queue when the barrier completes -- under the LSU
                    // congestion of the earlier direct-store epilogue for thousands of cycles -- and would then read the next
                    // block: seen once in ~10 runs as a 6 % error in one plane.)
                    float4 v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int f = q * kAxCompute + tid;
                        v[q] = *reinterpret_cast<const float4*>(stage + f * 4);
                        const float sc = s_scale[qt * kAxChunkRows + (f >> 5)];
                        v[q].x *= sc, v[q].y *= sc, v[q].z *= sc, v[q].w *= sc;
                        asm volatile("" ::"f"(v[q].x), "f"(v[q].y), "f"(v[q].z), "f"(v[q].w));  // (pins the products in front of the barrier)
                    }
                    named_bar_sync(3, kAxCompute);  // ring slot drained
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
                        const int f = q * kAxCompute + tid;
                        const int h = qt * kAxChunkRows + (f >> 5), w = (f & 31) << 2;
                        if (w >= w_lo && w < w_hi) {
                            __half2 hi[2], lo[2];
                            ax_split_raw2(v[q].x, v[q].y, hi[0], lo[0]);
                            ax_split_raw2(v[q].z, v[q].w, hi[1], lo[1]);
                            const uint32_t off = ax_tile_off(h, w);
                            *reinterpret_cast<uint2*>(x1 + off) = *reinterpret_cast<uint2*>(hi);
                            *reinterpret_cast<uint2*>(x2 + off) = *reinterpret_cast<uint2*>(lo);
                        }
                    }
                }
                fence_proxy_async();
                ax_fence_before();
                named_bar_arrive(1 + (u & 1), kAxCompute + 32);  // -> MMAs of block u
                if (tid == 0) ax_stamp(a.trace, 0, ns);  // block u converted
            }
        }
    }
    ax_fence_before();
    __syncthreads();
    if (warp == kAxComputeWarps) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
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

template <bool WIDE>
int launch_axis_mode(const AxisArgs& a, const float* in, float* out, int planes, int I, int K, int64_t units, cudaStream_t stream) {
    using L = AxLayout<WIDE>;
    AxisMaps maps;
    memset(&maps, 0, sizeof(maps));
    const cuuint32_t estr[3] = {1, 1, 1};
    const cuuint64_t dims[3] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(I), static_cast<cuuint64_t>(planes)};
    const cuuint64_t strides[2] = {static_cast<cuuint64_t>(K) * 4, static_cast<cuuint64_t>(K) * I * 4};
    const cuuint32_t box_in[3] = {kB, kAxChunkRows, 1};
    CUresult res = axis_encode_fn()(&maps.in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(in), dims, strides, box_in, estr,
                                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BGAN_REQUIRE(res == CUDA_SUCCESS, BGAN_E_UNSUPPORTED, "cuTensorMapEncodeTiled (axis input) failed (%d)", static_cast<int>(res));
    const cuuint64_t odims[3] = {static_cast<cuuint64_t>(I), static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(planes)};
    const cuuint64_t ostrides[2] = {static_cast<cuuint64_t>(I) * 4, static_cast<cuuint64_t>(K) * I * 4};
    const cuuint32_t box_out[3] = {static_cast<cuuint32_t>(L::kBoxCols), kB, 1};
    res = axis_encode_fn()(&maps.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, odims, ostrides, box_out, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           WIDE ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BGAN_REQUIRE(res == CUDA_SUCCESS, BGAN_E_UNSUPPORTED, "cuTensorMapEncodeTiled (axis output) failed (%d)", static_cast<int>(res));
    thread_local int attr_dev = -1;
    int dev = 0;
    BGAN_CUDA_OK(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        BGAN_CUDA_OK(cudaFuncSetAttribute(blur_axis_kernel<WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(L::kSmem)));
        attr_dev = dev;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    // balanced persistent grid: every CTA the same number of units where possible
    const int64_t rounds = ceil_div64(units, device_sm_count());
    cfg.gridDim = dim3(static_cast<unsigned>(ceil_div64(units, rounds)));
    cfg.blockDim = dim3(kAxCta);
    cfg.dynamicSmemBytes = L::kSmem;
    cfg.stream = stream;
    BGAN_CUDA_OK(cudaLaunchKernelEx(&cfg, blur_axis_kernel<WIDE>, a, maps));
    return 0;
}

// one pass: in [planes][I][K] -> out [planes][K][I]
int launch_axis(const float* in, float* out, const float* rowmax_in, float* rowmax_out, int planes, int I, int K, int rule, float std_eff,
                int radius, const float* taps_dev, cudaStream_t stream) {
    thread_local bool bound = false;
    if (!bound) {
        cudaFree(nullptr);
        bound = true;
    }
    AxisArgs a{};
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
    // work units: whole strips for radius <= 128, pairs of tiles above (BGAN_BLUR_AXIS_UNIT_TILES overrides: experiments / tests)
    a.unit_tiles = a.nb == 1 ? a.tiles_j : 2;
    if (const char* ut = getenv("BGAN_BLUR_AXIS_UNIT_TILES")) {
        if (atoi(ut) > 0) a.unit_tiles = std::min(atoi(ut), a.nb == 1 ? a.tiles_j : 2);
    }
    a.units_per_strip = ceil_div(a.tiles_j, a.unit_tiles);
    const int64_t units = static_cast<int64_t>(planes) * a.tiles_i * a.units_per_strip;
    a.n_units = static_cast<int>(units);
    if (a.nb == 1) return launch_axis_mode<false>(a, in, out, planes, I, K, units, stream);
    return launch_axis_mode<true>(a, in, out, planes, I, K, units, stream);
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

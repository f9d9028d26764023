// Tensor-core Gaussian blur for 128 x 128 planes (sm_100a: tcgen05.mma + TMEM + TMA).
//
// For one plane X [h][w] the separable zero-padded blur is  Y = T X T  with T the symmetric banded
// Toeplitz matrix of the taps (T[i][j] = tap[j - i + r] for |i - j| <= r).  Written as two
// 128 x 128 x 128 GEMMs the cost no longer depends on the tap count, which turns the FP32-FMA-bound
// wide blurs (sigma >= 3: 30 .. 258 FMAs per pixel on CUDA cores) into HBM-bound work.
//
// fp32 parity (<= 1e-5) on fp16 tensor cores: error-compensated split.  Every operand is written as
// a = a1 + 2^-11 a2 with a1 = fp16(a), a2 = fp16((a - a1) 2^11), and
//     A B ~= A1 B1 + 2^-11 (A1 B2 + A2 B1)                         (the dropped term is 2^-22 A2 B2)
// accumulated in fp32 in TMEM: the cross terms first, then A1 B1 with the accumulator pre-scaled by
// 2^-11 through tcgen05.mma's scale-input-d operand.  Planes are rescaled by a power of two so that
// max|x| lands in [2^12, 2^13) -- exact in fp32, keeps every piece inside fp16's normal range for any
// input magnitude -- and the inverse is applied to the output.  Measured error vs fp64: ~3e-7, the
// same as the fp32 FMA path (tests/test_blur_gpu.py).
//
// Orientation.  Both GEMMs produce transposed results so that no explicit transpose is needed and
// the final store is coalesced:
//   GEMM1  D1[w'][h]  = sum_w T[w'][w] X[h][w]      A = T (K-major)    B = X (K-major: row h, K = w)
//   GEMM2  D2[w'][h'] = sum_h D1[w'][h] T[h][h']    A = D1 split to fp16 (row w', K = h)   B = T
// D2 = Y^T sits in TMEM with lane = w': for a fixed register (h') the 32 lanes of a warp hold 32
// consecutive w' -> one coalesced 128-byte store per instruction.
//
// Per CTA (persistent, one per SM): TMA box load of the next fp32 plane into a staging buffer overlaps
// everything after the conversion of the current one.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "taps.cuh"

namespace bgan {
namespace {

constexpr int kP = 128;  // plane edge
constexpr int kTcThreads = 512;                          // compute threads; one more warp only issues the MMAs
constexpr int kTcWarps = kTcThreads / 32;                // 16: TMEM lane quadrant = warp & 3, column group = warp >> 2
constexpr int kMmaWarp = kTcWarps;                       // warp 16 issues the MMAs
constexpr int kStoreWarp0 = kTcWarps + 1;                // warps 17..20 drain D2 to global memory
constexpr int kStoreWarps = 8;                           // (schedules without bit 2 use the first four)
constexpr int kCtaThreads = kTcThreads + 32 + 32 * kStoreWarps;  // 800
constexpr int kColsPerWarp = kP / (kTcWarps / 4);        // 32 accumulator columns per thread in the epilogues
constexpr int kF4PerThread = kP * kP / 4 / kTcThreads;   // 8 float4 of the plane per thread
// operand tile: K-major, no swizzle, core matrix = 8 rows x 16 bytes.  Core matrices adjacent in K are
// 144 bytes apart (16 bytes of padding) so that a warp writing consecutive K-chunks of one row, or
// consecutive rows of one chunk, never hits a bank twice.
constexpr uint32_t kLBO = 144;              // K-adjacent core matrices
constexpr uint32_t kSBO = 16 * kLBO;        // row-group (8 rows) stride = 2304
constexpr uint32_t kTileBytes = 16 * kSBO;  // 36864
constexpr uint32_t kOffStage = 0;           // fp32 plane, TMA destination (64 KB)
constexpr uint32_t kOffT1 = 65536;
constexpr uint32_t kOffT2 = kOffT1 + kTileBytes;
constexpr uint32_t kOffO1 = kOffT2 + kTileBytes;  // X1, later Z1
constexpr uint32_t kOffO2 = kOffO1 + kTileBytes;  // X2, later Z2
constexpr uint32_t kOffMisc = kOffO2 + kTileBytes;
constexpr uint32_t kSmemTc = kOffMisc + 1024 /* taps */ + 256;
constexpr int kDefaultVariant = 7;
constexpr int kDefaultPrefetch = 1;
constexpr float kLoScale = 2048.0f, kLoInv = 1.0f / 2048.0f;
// instruction descriptor, kind::f16: D = f32, A = B = f16, both K-major, N = 128, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (0u << 7) | (0u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

constexpr uint32_t kIdescN64 = (1u << 4) | (0u << 7) | (0u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

struct TcArgs {
    long long* trace;  // debug: clock64 stamps of CTA 0 / thread 0 (nullptr in production)
    const float* x;    // only read directly for es > 1 (es == 1 goes through the tensor map)
    float* y;
    int es;  // 1: planes are contiguous [128][128]; 3: NHWC, plane p = channel p % 3 of image p / 3 (interleaved floats)
    int planes;
    int rule;
    float std_eff;
    int radius;
    int pf;   // L2 prefetch distance in planes (per CTA)
    int dbg;  // experiments only: switch stages off (results are wrong then)
};

__device__ __forceinline__ uint32_t tile_off(int row, int k) {
    return (row >> 3) * kSBO + (k >> 3) * kLBO + (row & 7) * 16 + (k & 7) * 2;
}

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>(kLBO >> 4) << 16) |
           (static_cast<uint64_t>(kSBO >> 4) << 32) | (1ull << 46);  // version 1, SWIZZLE_NONE, base offset 0
}

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
// D = A B + D * 2^-11
__device__ __forceinline__ void umma_f16_scale_d11(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p, 11;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(1), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
// A operand from TMEM (fp16, two K-elements per 32-bit column, lane = row), B from shared memory
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(kIdesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void umma_f16_ts_scale_d11(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p, 11;\n\t"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(kIdesc), "r"(1), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// tcgen05.wait::ld with the destination registers as in/out operands, so that the compiler cannot move any use of
// them above the wait
__device__ __forceinline__ void tmem_ld_wait(float (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]),
                   "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15]), "+f"(v[16]), "+f"(v[17]), "+f"(v[18]),
                   "+f"(v[19]), "+f"(v[20]), "+f"(v[21]), "+f"(v[22]), "+f"(v[23]), "+f"(v[24]), "+f"(v[25]), "+f"(v[26]), "+f"(v[27]),
                   "+f"(v[28]), "+f"(v[29]), "+f"(v[30]), "+f"(v[31])
                 :
                 : "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane; the registers are valid after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, float (&v)[32]) {
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

// 16 consecutive fp32 columns / 8 consecutive 32-bit columns (the split stage works in halves to keep registers free
// for the next plane, which is already on its way from staging to the fp16 tiles)
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
          "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait16(float (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]),
                   "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}

// a -> (fp16(a), fp16((a - fp16(a)) * 2^11)) for two values at a time
__device__ __forceinline__ void split2(float a, float b, __half2& hi, __half2& lo) {
    hi = __floats2half2_rn(a, b);
    const float2 back = __half22float2(hi);
    lo = __floats2half2_rn((a - back.x) * kLoScale, (b - back.y) * kLoScale);
}

// data-side split without the 2^11 scaling of the remainder (see "raw lo" below)
__device__ __forceinline__ void split2_raw(float a, float b, __half2& hi, __half2& lo) {
    hi = __floats2half2_rn(a, b);
    const float2 back = __half22float2(hi);
    lo = __floats2half2_rn(a - back.x, b - back.y);
}
template <bool kRaw>
__device__ __forceinline__ void split2_sel(float a, float b, __half2& hi, __half2& lo) {
    if constexpr (kRaw)
        split2_raw(a, b, hi, lo);
    else
        split2(a, b, hi, lo);
}

// the three products of one GEMM into one accumulator: cross terms, then the main term on the
// accumulator pre-scaled by 2^-11.  Each product is 8 instructions of K = 16.
__device__ __forceinline__ void issue_gemm(uint32_t d_tmem, uint32_t a1, uint32_t a2, uint32_t b1, uint32_t b2) {
    const uint64_t da1 = smem_desc(a1), da2 = smem_desc(a2), db1 = smem_desc(b1), db2 = smem_desc(b2);
    constexpr uint64_t kStep = (2 * kLBO) >> 4;  // 16 K-elements = 2 core matrices, in 16-byte units
#pragma unroll
    for (int k = 0; k < 8; ++k) umma_f16(d_tmem, da1 + k * kStep, db2 + k * kStep, k > 0);
#pragma unroll
    for (int k = 0; k < 8; ++k) umma_f16(d_tmem, da2 + k * kStep, db1 + k * kStep, 1);
    umma_f16_scale_d11(d_tmem, da1, db1);
#pragma unroll
    for (int k = 1; k < 8; ++k) umma_f16(d_tmem, da1 + k * kStep, db1 + k * kStep, 1);
}

// GEMM2 with the A operand (Z^T split into hi | lo, 64 + 64 columns at z_tmem) read from TMEM
__device__ __forceinline__ void issue_gemm_ts(uint32_t d_tmem, uint32_t z_tmem, uint32_t b1, uint32_t b2) {
    const uint64_t db1 = smem_desc(b1), db2 = smem_desc(b2);
    constexpr uint64_t kStep = (2 * kLBO) >> 4;
    const uint32_t a1 = z_tmem, a2 = z_tmem + 64;  // 8 columns per K = 16
#pragma unroll
    for (int k = 0; k < 8; ++k) umma_f16_ts(d_tmem, a1 + 8 * k, db2 + k * kStep, k > 0);
#pragma unroll
    for (int k = 0; k < 8; ++k) umma_f16_ts(d_tmem, a2 + 8 * k, db1 + k * kStep, 1);
    umma_f16_ts_scale_d11(d_tmem, a1, db1);
#pragma unroll
    for (int k = 1; k < 8; ++k) umma_f16_ts(d_tmem, a1 + 8 * k, db1 + k * kStep, 1);
}

// GEMM1 for one half of the plane's rows (N = 64 columns of D1): the two halves complete separately, so that the first
// half of D1 is split while the tensor pipe works on the second
__device__ __forceinline__ void umma_f16_ts_n64(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t accumulate, bool scale_d11) {
    if (scale_d11) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p, 11;\n\t"
            "}\n" ::"r"(d_tmem),
            "r"(a_tmem), "l"(bdesc), "r"(kIdescN64), "r"(1), "r"(0), "r"(0), "r"(0), "r"(0)
            : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
            "}\n" ::"r"(d_tmem),
            "r"(a_tmem), "l"(bdesc), "r"(kIdescN64), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
            : "memory");
    }
}
__device__ __forceinline__ void issue_gemm_ts_n64(uint32_t d_tmem, uint32_t t_tmem, uint32_t b1, uint32_t b2) {
    const uint64_t db1 = smem_desc(b1), db2 = smem_desc(b2);
    constexpr uint64_t kStep = (2 * kLBO) >> 4;
    const uint32_t a1 = t_tmem, a2 = t_tmem + 64;
#pragma unroll
    for (int k = 0; k < 8; ++k) umma_f16_ts_n64(d_tmem, a1 + 8 * k, db2 + k * kStep, k > 0, false);
#pragma unroll
    for (int k = 0; k < 8; ++k) umma_f16_ts_n64(d_tmem, a2 + 8 * k, db1 + k * kStep, 1, false);
    umma_f16_ts_n64(d_tmem, a1, db1, 1, true);
#pragma unroll
    for (int k = 1; k < 8; ++k) umma_f16_ts_n64(d_tmem, a1 + 8 * k, db1 + k * kStep, 1, false);
}

// The same GEMM2 in three pieces, so that the cross terms of K = h in [0, 64) can start while the second half of D1 is
// still being split: cross(half 0), cross(half 1), then the main term on the accumulator pre-scaled by 2^-11.
__device__ __forceinline__ void issue_gemm_ts_cross(uint32_t d_tmem, uint32_t z_tmem, uint32_t b1, uint32_t b2, int half) {
    const uint64_t db1 = smem_desc(b1), db2 = smem_desc(b2);
    constexpr uint64_t kStep = (2 * kLBO) >> 4;
    const uint32_t a1 = z_tmem, a2 = z_tmem + 64;
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_f16_ts(d_tmem, a1 + 8 * (4 * half + k), db2 + (4 * half + k) * kStep, (half | k) != 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_f16_ts(d_tmem, a2 + 8 * (4 * half + k), db1 + (4 * half + k) * kStep, 1);
}
__device__ __forceinline__ void issue_gemm_ts_main(uint32_t d_tmem, uint32_t z_tmem, uint32_t b1) {
    const uint64_t db1 = smem_desc(b1);
    constexpr uint64_t kStep = (2 * kLBO) >> 4;
    umma_f16_ts_scale_d11(d_tmem, z_tmem, db1);
#pragma unroll
    for (int k = 1; k < 8; ++k) umma_f16_ts(d_tmem, z_tmem + 8 * k, db1 + k * kStep, 1);
}

// ---- "raw lo" operand convention (schedule bit 4) -----------------------------------------------------------------------
// The data-side operands (X, Z) keep their fp16 remainder unscaled: after the per-plane power-of-two scaling max|x| is in
// [2^12, 2^13), so a remainder only becomes an fp16 subnormal for elements below 2^-15 of the plane maximum, where the
// absolute error (<= 2^-25) is irrelevant.  That saves one FMUL per element in both conversion stages.  T keeps its
// remainder scaled by 2^11 (taps are small).  Order of the products into one accumulator:
//   T_lo' x DATA_hi   (scale 2^11)   ->  accumulator * 2^-11 + T_hi x DATA_hi   ->  + T_hi x DATA_lo
template <bool kN64>
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t accumulate, bool scale_d11) {
    constexpr uint32_t idesc = kN64 ? kIdescN64 : kIdesc;
    if (scale_d11) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p, 11;\n\t"
            "}\n" ::"r"(d_tmem),
            "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(1), "r"(0), "r"(0), "r"(0), "r"(0)
            : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
            "}\n" ::"r"(d_tmem),
            "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
            : "memory");
    }
}
// GEMM1 (A = T hi | lo' in TMEM, B = X hi / raw lo tiles), whole plane or one N = 64 half
template <bool kN64>
__device__ __forceinline__ void issue_g1_rawlo(uint32_t d_tmem, uint32_t t_tmem, uint32_t x_hi, uint32_t x_lo) {
    const uint64_t dh = smem_desc(x_hi), dl = smem_desc(x_lo);
    constexpr uint64_t kStep = (2 * kLBO) >> 4;
#pragma unroll
    for (int k = 0; k < 8; ++k) mma_ts<kN64>(d_tmem, t_tmem + 64 + 8 * k, dh + k * kStep, k > 0, false);
    mma_ts<kN64>(d_tmem, t_tmem, dh, 1, true);
#pragma unroll
    for (int k = 1; k < 8; ++k) mma_ts<kN64>(d_tmem, t_tmem + 8 * k, dh + k * kStep, 1, false);
#pragma unroll
    for (int k = 0; k < 8; ++k) mma_ts<kN64>(d_tmem, t_tmem + 8 * k, dl + k * kStep, 1, false);
}
// GEMM2 (A = Z hi | raw lo in TMEM, B = T hi / lo' tiles): Z_hi x T_lo' for K-steps [k0, k1), then everything else
__device__ __forceinline__ void issue_g2_rawlo_cross(uint32_t d_tmem, uint32_t z_tmem, uint32_t t_lo, int k0, int k1) {
    const uint64_t dl = smem_desc(t_lo);
    constexpr uint64_t kStep = (2 * kLBO) >> 4;
#pragma unroll
    for (int k = 0; k < 8; ++k)
        if (k >= k0 && k < k1) mma_ts<false>(d_tmem, z_tmem + 8 * k, dl + k * kStep, k > 0, false);
}
__device__ __forceinline__ void issue_g2_rawlo_main(uint32_t d_tmem, uint32_t z_tmem, uint32_t t_hi) {
    const uint64_t dh = smem_desc(t_hi);
    constexpr uint64_t kStep = (2 * kLBO) >> 4;
    mma_ts<false>(d_tmem, z_tmem, dh, 1, true);
#pragma unroll
    for (int k = 1; k < 8; ++k) mma_ts<false>(d_tmem, z_tmem + 8 * k, dh + k * kStep, 1, false);
#pragma unroll
    for (int k = 0; k < 8; ++k) mma_ts<false>(d_tmem, z_tmem + 64 + 8 * k, dh + k * kStep, 1, false);
}

// NHWC (es == 3): TMA cannot walk the innermost dimension with an element stride (elementStrides[0] is ignored for
// non-interleaved maps) and the channel stride of 12 bytes is not a legal tensor-map stride either, so the compute
// threads gather channel c of image n with 4-byte cp.async (LDGSTS) straight into the same row-major [128][128] staging
// plane; completion is signalled on the same mbarrier (cp.async.mbarrier.arrive.noinc, one arrival per thread).
// The three channel planes of an image are consecutive work units, so neighbouring CTAs pull the same lines out of L2
// at about the same time: DRAM still sees every byte once.
__device__ __forceinline__ void gather_plane_load(unsigned char* stage, const float* x, uint64_t* bar, int plane, int es, int tid) {
    const float* src = x + static_cast<size_t>(plane / es) * (kP * kP) * es + plane % es;
    const uint32_t dst = smem_u32(stage) + tid * 4;
#pragma unroll 8
    for (int j = 0; j < kP * kP / kTcThreads; ++j) {
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst + j * kTcThreads * 4),
                     "l"(src + static_cast<size_t>(j * kTcThreads + tid) * es)
                     : "memory");
    }
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Warp-specialised software pipeline of one CTA over its planes p0, p1, ... :
//
//   compute warps 0-15          MMA warp 16                     store warps 17-20
//   ------------------------    ----------------------------    ---------------------------------
//   [X(p) tiles ready] -------> GEMM1(p): D1 = T X^T
//   fetch(p+1): staging->regs
//   wait GEMM1(p)
//   split(p): D1 -> Z (TMEM) -> GEMM2(p): D2[p&1] = Z^T T  --->  wait GEMM2(p); D2[p&1] -> global
//   convert(p+1): regs -> X tiles (free since GEMM1(p) completed); overlaps GEMM2(p)
//
// The tensor pipe only idles during split; the global store, the staging read and the fp16 conversion of the
// next plane all hide behind the MMAs, and the TMA load of plane p+2 has a whole iteration to land.
//
// kV selects the schedule (bit 0: the staging read of plane p+1 runs in the shadow of GEMM1(p) and its fp16 conversion in the
// shadow of GEMM2(p), instead of both behind GEMM2(p); bit 1: D1 is split in two K-halves and GEMM2's cross terms of the first
// half are issued while the second half is still being split).
template <int kV>
__global__ void __launch_bounds__(kCtaThreads, 1) blur_tc128_kernel(const TcArgs a, const __grid_constant__ CUtensorMap tmap) {
    constexpr bool kEarlyFetch = (kV & 1) != 0, kHalfSplit = (kV & 2) != 0;
    // bit 2: T also lives in TMEM as the A operand of GEMM1 (TS mode: GEMM1 only reads X from shared memory, half the
    // operand traffic); D2 is then single-buffered and eight store warps pull it into registers as soon as GEMM2 completes
    constexpr bool kTsG1 = (kV & 4) != 0;
    // bit 3 (needs bits 1 and 2): GEMM1 runs as two N = 64 halves (rows h < 64, h >= 64) that complete separately
    constexpr bool kSplitG1 = (kV & 8) != 0;
    static_assert(!kSplitG1 || (kTsG1 && kHalfSplit), "kSplitG1 needs kTsG1 and kHalfSplit");
    constexpr bool kRawLo = (kV & 16) != 0;  // bit 4 (needs bit 2): see "raw lo" above
    static_assert(!kRawLo || kTsG1, "kRawLo needs kTsG1");
    constexpr int kActiveStoreWarps = kTsG1 ? 8 : 4;
    constexpr int kStoreCols = kP / (kActiveStoreWarps / 4);  // D2 columns (= output rows h') per store warp: 128 or 64
    extern __shared__ __align__(1024) unsigned char smem[];
    const float* stage = reinterpret_cast<const float*>(smem + kOffStage);
    float* s_g = reinterpret_cast<float*>(smem + kOffMisc);
    uint64_t* mbar_tma = reinterpret_cast<uint64_t*>(smem + kOffMisc + 1024);
    uint64_t* mbar_g1 = mbar_tma + 1;
    uint64_t* mbar_g2 = mbar_tma + 2;      // [2]: GEMM2 into D2[b] complete
    uint64_t* mbar_d2free = mbar_tma + 4;  // [2]: D2[b] has been read out
    uint64_t* mbar_g1b = mbar_tma + 6;  // kSplitG1: second half of GEMM1 complete
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(mbar_tma + 7);
    float* s_inv = reinterpret_cast<float*>(mbar_tma + 8);   // [4] ring: 1 / scale of plane it
    float* s_max = s_inv + 4;                                // [2][16]
    const uint32_t sbase = smem_u32(smem);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int n_stamp = 0;
    auto stamp = [&]() {
        if (a.trace != nullptr && blockIdx.x == 0 && tid == 0 && n_stamp < 120) a.trace[1 + n_stamp++] = clock64();
    };
    // second and third timeline (debug): lane 0 of the MMA warp -> trace[128 ..], lane 0 of the first store warp -> trace[256 ..]
    auto stamp_role = [&](int base) {
        if (a.trace != nullptr && blockIdx.x == 0 && lane == 0 && n_stamp < 120) {
            a.trace[base + 1 + n_stamp++] = clock64();
            a.trace[base] = n_stamp;
        }
    };
    if (a.trace != nullptr && blockIdx.x == 0 && tid == 0) a.trace[126] = static_cast<long long>(global_timer_ns());
    stamp();  // 0: kernel start

    // Programmatic dependent launch: the next kernel in the stream may start its own prologue as soon as our CTAs
    // retire; we, in turn, must not read global memory before the previous grid has completed.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        mbar_init(mbar_tma, a.es == 1 ? 1 : kTcThreads);  // TMA transaction bytes / one cp.async arrival per compute thread
        mbar_init(mbar_g1, 1);
        mbar_init(mbar_g1b, 1);
        mbar_init(mbar_g2 + 0, 1);
        mbar_init(mbar_g2 + 1, 1);
        mbar_init(mbar_d2free + 0, kActiveStoreWarps);
        mbar_init(mbar_d2free + 1, kActiveStoreWarps);
        mbar_fence_init();
        // the first plane starts moving before anything else is set up (but not before the producer of x is done)
        if (a.es == 1) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            mbar_arrive_expect_tx(mbar_tma, kP * kP * 4);
            tma_load_3d(smem + kOffStage, &tmap, mbar_tma, 0, 0, static_cast<int>(blockIdx.x));
            // the L2 prefetch runs a.pf planes ahead of the loads: start it for planes 1 .. pf (the first fetch issues the
            // load of plane 1 and the prefetch of plane 1 + pf)
            for (int j = 1; j <= a.pf; ++j) {
                const int pp = static_cast<int>(blockIdx.x + j * gridDim.x);
                if (pp < a.planes) tma_prefetch_l2_3d(&tmap, 0, 0, pp);
            }
        }
    }
    __syncwarp();
    // TMEM columns: D1 = [0, 128), Z^T as fp16 hi | lo = [128, 256), D2 double-buffered = [256, 384) / [384, 512)
    // kTsG1:        T as fp16 hi | lo = [0, 128), D1 = [128, 256), Z^T = [256, 384), D2 = [384, 512)
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    stamp();  // after barrier init / TMEM alloc (warp 0)
    // ---- taps: one warp, reference expression order ---------------------------------------------------------
    const int n_taps = 2 * a.radius + 1;
    if (warp == 1) {
        float part = 0.0f;
        for (int t = lane; t < n_taps; t += 32) {
            const float g = raw_tap(a.rule, a.std_eff, t - a.radius);
            s_g[t] = g;
            part += g;
        }
        // fp32 partial sums of <= 8 taps each, combined in fp64: within 1 ulp of the reference's fp32 sum
        const float tot = static_cast<float>(warp_sum(static_cast<double>(part)));
        __syncwarp();
        for (int t = lane; t < n_taps; t += 32) s_g[t] = s_g[t] / tot;
    }
    __syncthreads();
    if (a.es != 1 && warp < kTcWarps) {  // (mbarrier init is visible now)
        asm volatile("griddepcontrol.wait;" ::: "memory");
        gather_plane_load(smem + kOffStage, a.x, mbar_tma, static_cast<int>(blockIdx.x), a.es, tid);
    }
    stamp();  // taps ready
    // ---- chunk table: a 16-byte chunk of T (row, K = 8c .. 8c+7) holds taps d0 .. d0+7 with d0 = 8c - row + r, so there
    // are only n_taps + 7 distinct non-zero chunks.  Build them once (fp16 hi | lo), then T1 / T2 are plain copies. ----
    uint4* s_chi = reinterpret_cast<uint4*>(smem + kOffO1);  // scratch inside the (not yet used) X tiles
    uint4* s_clo = s_chi + 272;
    for (int j = tid; j < (n_taps + 7) * 4; j += kCtaThreads) {
        const int i = j >> 2, e = j & 3;
        const int da = i - 7 + 2 * e, db = da + 1;
        __half2 hi, lo;
        split2((da >= 0 && da < n_taps) ? s_g[da] : 0.0f, (db >= 0 && db < n_taps) ? s_g[db] : 0.0f, hi, lo);
        reinterpret_cast<__half2*>(s_chi + i)[e] = hi;
        reinterpret_cast<__half2*>(s_clo + i)[e] = lo;
    }
    __syncthreads();
    stamp();  // chunk table
    if constexpr (kTsG1) {
        if (warp < kTcWarps) {  // T hi | lo -> TMEM: lane = row w', 32-bit column = two consecutive K (= w) elements
            tc_fence_after();
            const uint32_t t_tmem = *s_tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 16;
            const int trow = (warp & 3) * 32 + lane;
            uint32_t th[16], tl[16];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int i = ((warp >> 2) * 4 + c) * 8 - trow + a.radius + 7;
                const bool band = i >= 0 && i < n_taps + 7;
                const uint4 h4 = band ? s_chi[i] : make_uint4(0, 0, 0, 0), l4 = band ? s_clo[i] : make_uint4(0, 0, 0, 0);
                th[4 * c] = h4.x, th[4 * c + 1] = h4.y, th[4 * c + 2] = h4.z, th[4 * c + 3] = h4.w;
                tl[4 * c] = l4.x, tl[4 * c + 1] = l4.y, tl[4 * c + 2] = l4.z, tl[4 * c + 3] = l4.w;
            }
            tmem_st16(t_tmem, th);
            tmem_st16(t_tmem + 64, tl);
            tmem_st_wait();
        }
    }
    for (int idx = tid; idx < kP * 16; idx += kCtaThreads) {
        const int row = idx & 127, chunk = idx >> 7;  // lanes <-> consecutive rows: conflict-free 16-byte loads and stores
        const int i = chunk * 8 - row + a.radius + 7;
        const bool band = i >= 0 && i < n_taps + 7;
        const uint32_t off = tile_off(row, chunk * 8);
        *reinterpret_cast<uint4*>(smem + kOffT1 + off) = band ? s_chi[i] : make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(smem + kOffT2 + off) = band ? s_clo[i] : make_uint4(0, 0, 0, 0);
    }
    stamp();  // T copied
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t tT = tmem, d1 = tmem + (kTsG1 ? 128 : 0), zp = d1 + 128, d2 = zp + 128;
    (void)tT;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;  // TMEM lane quadrant of this warp
    const int row = (warp & 3) * 32 + lane;  // accumulator row (= w') owned by this thread
    const int first = blockIdx.x, stride = gridDim.x;
    // D2 buffer of iteration `it`, and the parity of the mbarrier phase that its `it`-th use completes
    auto d2_buf = [](int it) { return kTsG1 ? 0 : (it & 1); };
    auto d2_par = [](int it) { return kTsG1 ? (it & 1) : ((it >> 1) & 1); };
    constexpr int kD2Depth = kTsG1 ? 1 : 2;
    stamp();  // 1: prologue (TMEM, taps, T tiles) done
    asm volatile("griddepcontrol.wait;" ::: "memory");  // (all threads: y may still be read by the previous grid)

    if (warp == kMmaWarp) {
        // ================= MMA issuer =================================================================
        int it = 0;
        for (int p = first; p < a.planes; p += stride, ++it) {
            named_bar_sync(1, kTcThreads + 32);  // X(p) tiles written
            if (lane == 0) {
                tc_fence_after();
                stamp_role(128);  // G1 issue starts
                if constexpr (kSplitG1) {
                    if constexpr (kRawLo) {
                        issue_g1_rawlo<true>(d1, tT, sbase + kOffO1, sbase + kOffO2);
                        umma_commit(mbar_g1);
                        issue_g1_rawlo<true>(d1 + 64, tT, sbase + kOffO1 + 8 * kSBO, sbase + kOffO2 + 8 * kSBO);
                    } else {
                        issue_gemm_ts_n64(d1, tT, sbase + kOffO1, sbase + kOffO2);
                        umma_commit(mbar_g1);
                        issue_gemm_ts_n64(d1 + 64, tT, sbase + kOffO1 + 8 * kSBO, sbase + kOffO2 + 8 * kSBO);
                    }
                    umma_commit(mbar_g1b);
                } else {
                    if constexpr (kRawLo)
                        issue_g1_rawlo<false>(d1, tT, sbase + kOffO1, sbase + kOffO2);
                    else if constexpr (kTsG1)
                        issue_gemm_ts(d1, tT, sbase + kOffO1, sbase + kOffO2);
                    else
                        issue_gemm(d1, sbase + kOffT1, sbase + kOffT2, sbase + kOffO1, sbase + kOffO2);  // D1 = T X^T
                    umma_commit(mbar_g1);
                }
                stamp_role(128);  // G1 issued
            }
            __syncwarp();
            if constexpr (!kHalfSplit) {
                named_bar_sync(1, kTcThreads + 32);  // Z(p) written to TMEM
                if (lane == 0) {
                    if (it >= kD2Depth) mbar_wait(mbar_d2free + d2_buf(it), d2_par(it - kD2Depth));  // D2 drained by the store warps
                    tc_fence_after();
                    stamp_role(128);  // G2 issue starts
                    if constexpr (kRawLo) {
                        issue_g2_rawlo_cross(d2, zp, sbase + kOffT2, 0, 8);
                        issue_g2_rawlo_main(d2, zp, sbase + kOffT1);
                    } else {
                        issue_gemm_ts(d2 + d2_buf(it) * 128, zp, sbase + kOffT1, sbase + kOffT2);  // D2 = Z^T T = Y^T
                    }
                    umma_commit(mbar_g2 + d2_buf(it));
                    stamp_role(128);  // G2 issued
                }
                __syncwarp();
            } else {
                named_bar_sync(3, kTcThreads + 32);  // Z(p), K = h in [0, 64), written to TMEM
                if (lane == 0) {
                    if (it >= kD2Depth) mbar_wait(mbar_d2free + d2_buf(it), d2_par(it - kD2Depth));
                    tc_fence_after();
                    stamp_role(128);  // G2 (first half) issue starts
                    if constexpr (kRawLo)
                        issue_g2_rawlo_cross(d2, zp, sbase + kOffT2, 0, 4);
                    else
                        issue_gemm_ts_cross(d2 + d2_buf(it) * 128, zp, sbase + kOffT1, sbase + kOffT2, 0);
                }
                __syncwarp();
                named_bar_sync(4, kTcThreads + 32);  // Z(p), K in [64, 128)
                if (lane == 0) {
                    tc_fence_after();
                    if constexpr (kRawLo) {
                        issue_g2_rawlo_cross(d2, zp, sbase + kOffT2, 4, 8);
                        issue_g2_rawlo_main(d2, zp, sbase + kOffT1);
                    } else {
                        issue_gemm_ts_cross(d2 + d2_buf(it) * 128, zp, sbase + kOffT1, sbase + kOffT2, 1);
                        issue_gemm_ts_main(d2 + d2_buf(it) * 128, zp, sbase + kOffT1);
                    }
                    umma_commit(mbar_g2 + d2_buf(it));
                    stamp_role(128);  // G2 issued
                }
                __syncwarp();
            }
        }
    } else if (warp >= kStoreWarp0) {
        // ================= store warps: D2 = Y^T -> global, coalesced along w' =================================
        if (warp < kStoreWarp0 + kActiveStoreWarps) {
            const int c0 = ((warp - kStoreWarp0) >> 2) * kStoreCols;  // first D2 column (= output row h') of this warp
            int it = 0;
            // (without kTsG1 the last plane is drained by the compute warps)
            for (int p = first; p + (kTsG1 ? 0 : stride) < a.planes; p += stride, ++it) {
                mbar_wait(mbar_g2 + d2_buf(it), d2_par(it));
                tc_fence_after();
                if (warp == kStoreWarp0) stamp_role(256);  // G2(p) complete, store starts
                const float inv = s_inv[it & 3];
                float* yp = a.y + (static_cast<size_t>(p / a.es) * kP * kP + row) * a.es + p % a.es;
                const size_t ld = static_cast<size_t>(kP) * a.es;  // floats between consecutive output rows h'
                if constexpr (kTsG1) {
                    // the whole 64-column share goes to registers first, so that D2 is free again after ~1000 cycles
                    float o0[32], o1[32];
                    tmem_ld32_nowait(d2 + lane_base + c0, o0);
                    tmem_ld32_nowait(d2 + lane_base + c0 + 32, o1);
                    tmem_ld_wait(o0);
                    tmem_ld_wait(o1);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(mbar_d2free);
                    float* q = yp + static_cast<size_t>(c0) * ld;
                    // (the empty asm keeps the compiler from materialising all 64 addresses up front, which spills)
#pragma unroll
                    for (int j = 0; j < 32; ++j, q += ld) {
                        asm volatile("" : "+l"(q));
                        __stcs(q, o0[j] * inv);
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j, q += ld) {
                        asm volatile("" : "+l"(q));
                        __stcs(q, o1[j] * inv);
                    }
                } else {
#pragma unroll 1
                    for (int c = 0; c < 4; ++c) {
                        float o[32];
                        tmem_ld32_nowait(d2 + d2_buf(it) * 128 + lane_base + c * 32, o);
                        tmem_ld_wait(o);
                        if (c == 3) {  // everything of D2[it & 1] is in registers: hand the buffer back
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(mbar_d2free + d2_buf(it));
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j) __stcs(yp + static_cast<size_t>(c * 32 + j) * ld, o[j] * inv);
                    }
                }
                if (warp == kStoreWarp0) stamp_role(256);  // last store of plane p issued
            }
        }
    } else {
        // ================= compute warps ==============================================================
        const int col_base = (warp >> 2) * kColsPerWarp;
        uint32_t par_tma = 0;
        float4 v[kF4PerThread];

        // staging -> registers + per-warp max into s_max[buf]; ends with a barrier of the compute warps, after which
        // staging is free and thread 0 starts the TMA load of the plane after next
        auto fetch = [&](int buf, int next_plane) {
            stamp();  // fetch: waiting for the plane
            mbar_wait(mbar_tma, par_tma);
            stamp();  // fetch: plane landed
            par_tma ^= 1;
            float m = 0.0f;
#pragma unroll
            for (int j = 0; j < kF4PerThread; ++j) {
                v[j] = *reinterpret_cast<const float4*>(stage + (j * kTcThreads + tid) * 4);
                m = fmaxf(m, fmaxf(fmaxf(fabsf(v[j].x), fabsf(v[j].y)), fmaxf(fabsf(v[j].z), fabsf(v[j].w))));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) s_max[buf * kTcWarps + warp] = m;
            named_bar_sync(2, kTcThreads);
            if (a.es != 1) {
                if (next_plane < a.planes) gather_plane_load(smem + kOffStage, a.x, mbar_tma, next_plane, a.es, tid);
            } else if (tid == 0 && next_plane < a.planes) {
                mbar_arrive_expect_tx(mbar_tma, kP * kP * 4);
                tma_load_3d(smem + kOffStage, &tmap, mbar_tma, 0, 0, next_plane);
                // warm L2 further ahead (a.pf planes beyond the one being loaded)
                if (a.pf > 0 && next_plane + a.pf * stride < a.planes) tma_prefetch_l2_3d(&tmap, 0, 0, next_plane + a.pf * stride);
            }
        };
        // registers -> X1, X2 tiles (B operand of GEMM1: row = h, K = w), scaled so that max|x| is in [2^12, 2^13)
        auto convert = [&](int buf, int it_of_plane) {
            float m = s_max[buf * kTcWarps];
#pragma unroll
            for (int w = 1; w < kTcWarps; ++w) m = fmaxf(m, s_max[buf * kTcWarps + w]);
            int eb = (__float_as_int(m) >> 23) & 0xff;
            eb = min(max(eb, 13), 254);
            const float scale = __int_as_float((266 - eb) << 23);
            if (tid == 0) s_inv[it_of_plane & 3] = __int_as_float((eb - 12) << 23);
#pragma unroll
            for (int j = 0; j < kF4PerThread; ++j) {
                const int f = j * kTcThreads + tid;  // float4 index: row h = f / 32, w = 4 * (f % 32)
                const int h = f >> 5, w = (f & 31) << 2;
                __half2 hi[2], lo[2];
                split2_sel<kRawLo>(v[j].x * scale, v[j].y * scale, hi[0], lo[0]);
                split2_sel<kRawLo>(v[j].z * scale, v[j].w * scale, hi[1], lo[1]);
                const uint32_t off = tile_off(h, w);
                *reinterpret_cast<uint2*>(smem + kOffO1 + off) = *reinterpret_cast<uint2*>(hi);
                *reinterpret_cast<uint2*>(smem + kOffO2 + off) = *reinterpret_cast<uint2*>(lo);
            }
            fence_proxy_async();
        };

        fetch(0, first + stride);
        convert(0, 0);
        stamp();  // 2: first plane converted

        int it = 0;
        for (int p = first; p < a.planes; p += stride, ++it) {
            const bool has_next = p + stride < a.planes;
            tc_fence_before();
            named_bar_sync(1, kTcThreads + 32);  // -> GEMM1(p)
            stamp();                             // +0
            if constexpr (kEarlyFetch) {
                // while GEMM1(p) runs: the next plane goes from staging into registers (its load was issued an iteration ago)
                if (has_next) fetch((it + 1) & 1, p + 2 * stride);
            }
            stamp();                             // +1: (early fetch) next plane in registers
            mbar_wait(mbar_g1, it & 1);          // GEMM1(p) complete: D1 valid, X tiles free
            if (it > 0) mbar_wait(mbar_g2 + d2_buf(it - 1), d2_par(it - 1));  // GEMM2(p - 1) complete: Z free
            tc_fence_after();
            stamp();                             // +2

            // ---- split: D1 (fp32) -> Z^T hi | lo as packed fp16 back into TMEM (A operand of GEMM2: lane = w', K = h)
            if constexpr (!kEarlyFetch && !kHalfSplit) {
                float z[32];
                tmem_ld32_nowait(d1 + lane_base + col_base, z);
                tmem_ld_wait(z);
                uint32_t zh[16], zl[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    __half2 hi, lo;
                    split2(z[2 * e], z[2 * e + 1], hi, lo);
                    zh[e] = *reinterpret_cast<uint32_t*>(&hi);
                    zl[e] = *reinterpret_cast<uint32_t*>(&lo);
                }
                tmem_st16(zp + lane_base + (col_base >> 1), zh);
                tmem_st16(zp + 64 + lane_base + (col_base >> 1), zl);
                tmem_st_wait();
                tc_fence_before();
                named_bar_sync(1, kTcThreads + 32);  // -> GEMM2(p)
            } else {
                // 16 columns at a time (the registers of the next plane stay live across this stage).  With kHalfSplit a
                // warp owns columns [16 g, 16 g + 16) and [64 + 16 g, 64 + 16 g + 16), g = warp / 4: after the first pass
                // K = h in [0, 64) of Z is complete and the MMA warp starts GEMM2's cross terms on it.
#pragma unroll 1
                for (int half = 0; half < 2; ++half) {
                    const int c0 = kHalfSplit ? 64 * half + (col_base >> 1) : col_base + 16 * half;
                    if constexpr (kSplitG1) {
                        if (half == 1) {
                            mbar_wait(mbar_g1b, it & 1);  // second half of D1 valid, X tiles free
                            tc_fence_after();
                        }
                    }
                    float z[16];
                    tmem_ld16_nowait(d1 + lane_base + c0, z);
                    tmem_ld_wait16(z);
                    uint32_t zh[8], zl[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        __half2 hi, lo;
                        split2_sel<kRawLo>(z[2 * e], z[2 * e + 1], hi, lo);
                        zh[e] = *reinterpret_cast<uint32_t*>(&hi);
                        zl[e] = *reinterpret_cast<uint32_t*>(&lo);
                    }
                    tmem_st8(zp + lane_base + (c0 >> 1), zh);
                    tmem_st8(zp + 64 + lane_base + (c0 >> 1), zl);
                    if constexpr (kHalfSplit) {
                        tmem_st_wait();
                        tc_fence_before();
                        named_bar_arrive(3 + half, kTcThreads + 32);  // -> GEMM2(p), this K-half
                    }
                }
                if constexpr (!kHalfSplit) {
                    tmem_st_wait();
                    tc_fence_before();
                    named_bar_sync(1, kTcThreads + 32);  // -> GEMM2(p)
                }
            }
            stamp();                             // +3: split done
            // while GEMM2(p) runs: (staging -> registers ->) fp16 tiles of the next plane (free since GEMM1(p) completed)
            if (has_next) {
                if constexpr (!kEarlyFetch) fetch((it + 1) & 1, p + 2 * stride);
                convert((it + 1) & 1, it + 1);
            }
            stamp();                             // +4: next plane converted
        }
        // ---- drain: all 16 compute warps store the last plane (one 32-column group each) ---------------------
        if constexpr (!kTsG1) {
            const int last = it - 1, p_last = first + last * stride;
            mbar_wait(mbar_g2 + (last & 1), (last >> 1) & 1);
            tc_fence_after();
            const float inv = s_inv[last & 3];
            float* yp = a.y + (static_cast<size_t>(p_last / a.es) * kP * kP + row) * a.es + p_last % a.es;
            const size_t ld = static_cast<size_t>(kP) * a.es;
            float o[32];
            tmem_ld32_nowait(d2 + (last & 1) * 128 + lane_base + col_base, o);
            tmem_ld_wait(o);
#pragma unroll
            for (int j = 0; j < 32; ++j) __stcs(yp + static_cast<size_t>(col_base + j) * ld, o[j] * inv);
        }
        stamp();  // drain
    }
    if (a.trace != nullptr && blockIdx.x == 0 && tid == 0) {
        a.trace[0] = n_stamp;
        a.trace[127] = static_cast<long long>(global_timer_ns());
        a.trace[125] = clock64();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// =====================================================================================================================
// Two planes in flight.  T sits in TMEM (A operand of GEMM1), the accumulator of GEMM1 is double-buffered and split to
// fp16 *in place* (a thread's 32 fp32 columns become its 16 hi + 16 lo packed columns, so no thread ever writes a column
// that another one still has to read), D2 is single-buffered and pulled into registers by eight store warps:
//
//   TMEM columns   T hi | lo' [0, 128)    D1 / Z of even planes [128, 256)    D1 / Z of odd planes [256, 384)    D2 [384, 512)
//
//   tensor pipe     G1(i+1)                                  G2(i)                              G1(i+2) ...
//   compute warps   split(i) | staging -> registers (i+2)    | convert(i+2) -> X tiles (free: G1(i+1) done)
//   store warps     ... D2(i-1) -> global ...                                  D2(i) -> registers -> global
//
// The tensor pipe never waits for the split (it works on the next plane's GEMM1 meanwhile) and the compute warps never
// hold the next plane in registers across the split.  Operands use the "raw lo" convention above.
__device__ __forceinline__ void issue_g2_inplace(uint32_t d_tmem, uint32_t dz, uint32_t t_hi, uint32_t t_lo) {
    const uint64_t dh = smem_desc(t_hi), dl = smem_desc(t_lo);
    constexpr uint64_t kStep = (2 * kLBO) >> 4;
    // K-step k (h in [16 k, 16 k + 16)): Z_hi at column 32 (k / 2) + 8 (k % 2), Z_lo 16 columns further
#pragma unroll
    for (int k = 0; k < 8; ++k) mma_ts<false>(d_tmem, dz + 32 * (k >> 1) + 8 * (k & 1), dl + k * kStep, k > 0, false);
    mma_ts<false>(d_tmem, dz, dh, 1, true);
#pragma unroll
    for (int k = 1; k < 8; ++k) mma_ts<false>(d_tmem, dz + 32 * (k >> 1) + 8 * (k & 1), dh + k * kStep, 1, false);
#pragma unroll
    for (int k = 0; k < 8; ++k) mma_ts<false>(d_tmem, dz + 32 * (k >> 1) + 8 * (k & 1) + 16, dh + k * kStep, 1, false);
}

__global__ void __launch_bounds__(kCtaThreads, 1) blur_tc128_pipe_kernel(const TcArgs a, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(1024) unsigned char smem[];
    const float* stage = reinterpret_cast<const float*>(smem + kOffStage);
    float* s_g = reinterpret_cast<float*>(smem + kOffMisc);
    uint64_t* mbar_tma = reinterpret_cast<uint64_t*>(smem + kOffMisc + 1024);
    uint64_t* mbar_g1 = mbar_tma + 1;      // [2]: GEMM1 of an even / odd plane complete
    uint64_t* mbar_g2 = mbar_tma + 3;      // GEMM2 complete (D2 valid)
    uint64_t* mbar_d2free = mbar_tma + 4;  // D2 is in the store warps' registers
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(mbar_tma + 7);
    float* s_inv = reinterpret_cast<float*>(mbar_tma + 8);  // [4] ring: 1 / scale of plane it
    float* s_max = s_inv + 4;                               // [2][16]
    const uint32_t sbase = smem_u32(smem);
    constexpr int kBarA0 = 1, kBarA1 = 3, kBarB0 = 4;  // named barriers: X(i) ready (even / odd i), Z(i) ready (4 + i % 2); 2 = fetch

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int n_stamp = 0;
    auto stamp = [&]() {
        if (a.trace != nullptr && blockIdx.x == 0 && tid == 0 && n_stamp < 120) a.trace[1 + n_stamp++] = clock64();
    };
    auto stamp_role = [&](int base) {
        if (a.trace != nullptr && blockIdx.x == 0 && lane == 0 && n_stamp < 120) {
            a.trace[base + 1 + n_stamp++] = clock64();
            a.trace[base] = n_stamp;
        }
    };
    if (a.trace != nullptr && blockIdx.x == 0 && tid == 0) a.trace[126] = static_cast<long long>(global_timer_ns());
    stamp();  // kernel start

    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        tma_prefetch_desc(&tmap);
        mbar_init(mbar_tma, a.es == 1 ? 1 : kTcThreads);
        mbar_init(mbar_g1 + 0, 1);
        mbar_init(mbar_g1 + 1, 1);
        mbar_init(mbar_g2, 1);
        mbar_init(mbar_d2free, kStoreWarps);
        mbar_fence_init();
        if (a.es == 1) {
            asm volatile("griddepcontrol.wait;" ::: "memory");
            mbar_arrive_expect_tx(mbar_tma, kP * kP * 4);
            tma_load_3d(smem + kOffStage, &tmap, mbar_tma, 0, 0, static_cast<int>(blockIdx.x));
            for (int j = 1; j <= a.pf; ++j) {
                const int pp = static_cast<int>(blockIdx.x + j * gridDim.x);
                if (pp < a.planes) tma_prefetch_l2_3d(&tmap, 0, 0, pp);
            }
        }
    }
    __syncwarp();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    stamp();  // barrier init / TMEM alloc
    const int n_taps = 2 * a.radius + 1;
    if (warp == 1) {
        float part = 0.0f;
        for (int t = lane; t < n_taps; t += 32) {
            const float g = raw_tap(a.rule, a.std_eff, t - a.radius);
            s_g[t] = g;
            part += g;
        }
        const float tot = static_cast<float>(warp_sum(static_cast<double>(part)));
        __syncwarp();
        for (int t = lane; t < n_taps; t += 32) s_g[t] = s_g[t] / tot;
    }
    __syncthreads();
    if (a.es != 1 && warp < kTcWarps) {
        asm volatile("griddepcontrol.wait;" ::: "memory");
        gather_plane_load(smem + kOffStage, a.x, mbar_tma, static_cast<int>(blockIdx.x), a.es, tid);
    }
    stamp();  // taps
    uint4* s_chi = reinterpret_cast<uint4*>(smem + kOffO1);
    uint4* s_clo = s_chi + 272;
    for (int j = tid; j < (n_taps + 7) * 4; j += kCtaThreads) {
        const int i = j >> 2, e = j & 3;
        const int da = i - 7 + 2 * e, db = da + 1;
        __half2 hi, lo;
        split2((da >= 0 && da < n_taps) ? s_g[da] : 0.0f, (db >= 0 && db < n_taps) ? s_g[db] : 0.0f, hi, lo);
        reinterpret_cast<__half2*>(s_chi + i)[e] = hi;
        reinterpret_cast<__half2*>(s_clo + i)[e] = lo;
    }
    __syncthreads();
    stamp();  // chunk table
    if (warp < kTcWarps) {  // T hi | lo' -> TMEM: lane = row w', 32-bit column = two consecutive K (= w) elements
        tc_fence_after();
        const uint32_t t_tmem = *s_tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + (warp >> 2) * 16;
        const int trow = (warp & 3) * 32 + lane;
        uint32_t th[16], tl[16];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int i = ((warp >> 2) * 4 + c) * 8 - trow + a.radius + 7;
            const bool band = i >= 0 && i < n_taps + 7;
            const uint4 h4 = band ? s_chi[i] : make_uint4(0, 0, 0, 0), l4 = band ? s_clo[i] : make_uint4(0, 0, 0, 0);
            th[4 * c] = h4.x, th[4 * c + 1] = h4.y, th[4 * c + 2] = h4.z, th[4 * c + 3] = h4.w;
            tl[4 * c] = l4.x, tl[4 * c + 1] = l4.y, tl[4 * c + 2] = l4.z, tl[4 * c + 3] = l4.w;
        }
        tmem_st16(t_tmem, th);
        tmem_st16(t_tmem + 64, tl);
        tmem_st_wait();
    }
    for (int idx = tid; idx < kP * 16; idx += kCtaThreads) {
        const int row = idx & 127, chunk = idx >> 7;
        const int i = chunk * 8 - row + a.radius + 7;
        const bool band = i >= 0 && i < n_taps + 7;
        const uint32_t off = tile_off(row, chunk * 8);
        *reinterpret_cast<uint4*>(smem + kOffT1 + off) = band ? s_chi[i] : make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(smem + kOffT2 + off) = band ? s_clo[i] : make_uint4(0, 0, 0, 0);
    }
    stamp();  // T in TMEM and in shared memory
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *s_tmem;
    const uint32_t tT = tmem, dz = tmem + 128, d2 = tmem + 384;
    const uint32_t lane_base = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const int row = (warp & 3) * 32 + lane;
    const int first = blockIdx.x, stride = gridDim.x;
    const int n = first < a.planes ? (a.planes - first + stride - 1) / stride : 0;  // planes of this CTA
    stamp();  // prologue done
    asm volatile("griddepcontrol.wait;" ::: "memory");

    if (warp == kMmaWarp) {
        // ================= MMA issuer: G1(0), then G1(i + 1), G2(i) ===================================================
        if (n > 0) {
            named_bar_sync(kBarA0, kTcThreads + 32);
            if (lane == 0) {
                tc_fence_after();
                stamp_role(128);
                issue_g1_rawlo<false>(dz, tT, sbase + kOffO1, sbase + kOffO2);
                umma_commit(mbar_g1);
                stamp_role(128);
            }
            __syncwarp();
        }
        for (int i = 0; i < n; ++i) {
            if (i + 1 < n) {
                named_bar_sync(((i + 1) & 1) ? kBarA1 : kBarA0, kTcThreads + 32);  // X(i + 1) tiles written
                if (lane == 0) {
                    tc_fence_after();
                    stamp_role(128);  // G1 issue starts
                    issue_g1_rawlo<false>(dz + ((i + 1) & 1) * 128, tT, sbase + kOffO1, sbase + kOffO2);
                    umma_commit(mbar_g1 + ((i + 1) & 1));
                    stamp_role(128);  // G1 issued
                }
                __syncwarp();
            }
            named_bar_sync(kBarB0 + (i & 1), kTcThreads + 32);  // Z(i) written (in place)
            if (lane == 0) {
                if (i >= 1) mbar_wait(mbar_d2free, (i - 1) & 1);  // D2(i - 1) is in the store warps' registers
                tc_fence_after();
                stamp_role(128);  // G2 issue starts
                issue_g2_inplace(d2, dz + (i & 1) * 128, sbase + kOffT1, sbase + kOffT2);
                umma_commit(mbar_g2);
                stamp_role(128);  // G2 issued
            }
            __syncwarp();
        }
    } else if (warp >= kStoreWarp0) {
        // ================= store warps: D2 = Y^T -> registers -> global, coalesced along w' ===========================
        const int c0 = ((warp - kStoreWarp0) >> 2) * 64;  // first D2 column (= output row h') of this warp
        for (int i = 0; i < n; ++i) {
            const int p = first + i * stride;
            mbar_wait(mbar_g2, i & 1);
            tc_fence_after();
            if (warp == kStoreWarp0) stamp_role(256);  // G2(i) complete
            const float inv = s_inv[i & 3];
            float* q = a.y + (static_cast<size_t>(p / a.es) * kP * kP + row) * a.es + p % a.es;
            const size_t ld = static_cast<size_t>(kP) * a.es;  // floats between consecutive output rows h'
            q += static_cast<size_t>(c0) * ld;
            float o0[32], o1[32];
            if (!(a.dbg & 2)) {
                tmem_ld32_nowait(d2 + lane_base + c0, o0);
                tmem_ld32_nowait(d2 + lane_base + c0 + 32, o1);
                tmem_ld_wait(o0);
                tmem_ld_wait(o1);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) o0[j] = o1[j] = inv;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(mbar_d2free);
            // (the empty asm keeps the compiler from materialising all 64 addresses up front, which spills)
            if (!(a.dbg & 1)) {
#pragma unroll
                for (int j = 0; j < 32; ++j, q += ld) {
                    asm volatile("" : "+l"(q));
                    __stcs(q, o0[j] * inv);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j, q += ld) {
                    asm volatile("" : "+l"(q));
                    __stcs(q, o1[j] * inv);
                }
            } else if (o0[3] + o1[5] == 12345.678f) {
                __stcs(q, o0[0]);
            }
            if (warp == kStoreWarp0) stamp_role(256);  // stores of plane i issued
        }
    } else {
        // ================= compute warps ==============================================================================
        const int col_base = (warp >> 2) * kColsPerWarp;
        uint32_t par_tma = 0;
        float4 v[kF4PerThread];
        bool buf_first = true;  // (debug flag 32: only the first plane is really loaded)
        auto fetch = [&](int buf, int next_plane) {
            stamp();  // fetch: waiting for the plane
            const bool reload = !(a.dbg & 32) || buf_first;
            if (reload) {
                mbar_wait(mbar_tma, par_tma);
                par_tma ^= 1;
            }
            stamp();  // fetch: plane landed
            float m = 0.0f;
#pragma unroll
            for (int j = 0; j < kF4PerThread; ++j) {
                if (!(a.dbg & 8))
                    v[j] = *reinterpret_cast<const float4*>(stage + (j * kTcThreads + tid) * 4);
                else
                    v[j] = make_float4(1.0f, 0.5f, 0.25f, 0.125f);
                m = fmaxf(m, fmaxf(fmaxf(fabsf(v[j].x), fabsf(v[j].y)), fmaxf(fabsf(v[j].z), fabsf(v[j].w))));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (lane == 0) s_max[buf * kTcWarps + warp] = m;
            named_bar_sync(2, kTcThreads);
            buf_first = false;
            if (a.dbg & 32) {
            } else if (a.es != 1) {
                if (next_plane < a.planes) gather_plane_load(smem + kOffStage, a.x, mbar_tma, next_plane, a.es, tid);
            } else if (tid == 0 && next_plane < a.planes) {
                mbar_arrive_expect_tx(mbar_tma, kP * kP * 4);
                tma_load_3d(smem + kOffStage, &tmap, mbar_tma, 0, 0, next_plane);
                if (a.pf > 0 && next_plane + a.pf * stride < a.planes) tma_prefetch_l2_3d(&tmap, 0, 0, next_plane + a.pf * stride);
            }
        };
        auto convert = [&](int buf, int it_of_plane) {
            float m = s_max[buf * kTcWarps];
#pragma unroll
            for (int w = 1; w < kTcWarps; ++w) m = fmaxf(m, s_max[buf * kTcWarps + w]);
            int eb = (__float_as_int(m) >> 23) & 0xff;
            eb = min(max(eb, 13), 254);
            const float scale = __int_as_float((266 - eb) << 23);
            if (tid == 0) s_inv[it_of_plane & 3] = __int_as_float((eb - 12) << 23);
#pragma unroll
            for (int j = 0; j < kF4PerThread; ++j) {
                const int f = j * kTcThreads + tid;
                const int h = f >> 5, w = (f & 31) << 2;
                __half2 hi[2], lo[2];
                split2_raw(v[j].x * scale, v[j].y * scale, hi[0], lo[0]);
                split2_raw(v[j].z * scale, v[j].w * scale, hi[1], lo[1]);
                const uint32_t off = tile_off(h, w);
                if (!(a.dbg & 16) || it_of_plane == 0) {
                    *reinterpret_cast<uint2*>(smem + kOffO1 + off) = *reinterpret_cast<uint2*>(hi);
                    *reinterpret_cast<uint2*>(smem + kOffO2 + off) = *reinterpret_cast<uint2*>(lo);
                }
            }
            fence_proxy_async();
        };

        int g1_waited = -1;  // GEMM1 of planes <= g1_waited is known to be complete
        auto wait_g1 = [&](int i) {
            if (g1_waited < i) {
                mbar_wait(mbar_g1 + (i & 1), (i >> 1) & 1);
                g1_waited = i;
            }
            tc_fence_after();
        };
        if (n > 0) {
            fetch(0, first + stride);
            convert(0, 0);
            named_bar_arrive(kBarA0, kTcThreads + 32);  // -> G1(0)
            stamp();  // first plane converted
            if (n > 1) {
                fetch(1, first + 2 * stride);
                wait_g1(0);  // X tiles free
                convert(1, 1);
                named_bar_arrive(kBarA1, kTcThreads + 32);  // -> G1(1)
            }
            stamp();  // second plane converted
        }
        for (int i = 0; i < n; ++i) {
            wait_g1(i);
            stamp();  // +0: D1(i) valid
            // ---- split(i), in place: this thread's fp32 columns [32 g, 32 g + 32) -> hi [32 g, +16) | lo [32 g + 16, +16) -----
            {
                const uint32_t base = dz + (i & 1) * 128 + lane_base + col_base;
                float z[32];
                if (!(a.dbg & 4)) {
                    tmem_ld32_nowait(base, z);
                    tmem_ld_wait(z);
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) z[e] = static_cast<float>(e + i);
                }
                uint32_t zh[16], zl[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    __half2 hi, lo;
                    split2_raw(z[2 * e], z[2 * e + 1], hi, lo);
                    zh[e] = *reinterpret_cast<uint32_t*>(&hi);
                    zl[e] = *reinterpret_cast<uint32_t*>(&lo);
                }
                if (!(a.dbg & 4)) {
                    tmem_st16(base, zh);
                    tmem_st16(base + 16, zl);
                    tmem_st_wait();
                }
                tc_fence_before();
                named_bar_arrive(kBarB0 + (i & 1), kTcThreads + 32);  // -> G2(i)
            }
            stamp();  // +1: split done
            if (i + 2 < n) {
                fetch(i & 1, first + (i + 3) * stride);  // plane i + 2: staging -> registers; starts the load of plane i + 3
                wait_g1(i + 1);                          // X tiles free
                stamp();                                 // +4: G1(i + 1) complete
                convert(i & 1, i + 2);
                named_bar_arrive(((i + 2) & 1) ? kBarA1 : kBarA0, kTcThreads + 32);  // -> G1(i + 2)
            }
            stamp();  // +5: plane i + 2 converted
        }
    }
    if (a.trace != nullptr && blockIdx.x == 0 && tid == 0) {
        a.trace[0] = n_stamp;
        a.trace[127] = static_cast<long long>(global_timer_ns());
        a.trace[125] = clock64();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

using TcKernel = void (*)(const TcArgs, const CUtensorMap);
constexpr int kNumVariants = 13;
constexpr TcKernel kVariants[kNumVariants] = {blur_tc128_kernel<0>, blur_tc128_kernel<1>, blur_tc128_kernel<2>, blur_tc128_kernel<3>,
                                              blur_tc128_kernel<4>, blur_tc128_kernel<5>, blur_tc128_kernel<6>, blur_tc128_kernel<7>,
                                              blur_tc128_kernel<14>, blur_tc128_kernel<15>,   // [8], [9]: with bit 3
                                              blur_tc128_kernel<23>, blur_tc128_kernel<31>,   // [10], [11]: 7 / 15 with bit 4
                                              blur_tc128_pipe_kernel};                        // [12]: two planes in flight

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled is a driver entry point and needs a context bound to the calling thread; PyTorch's autograd
// worker threads only have the runtime's lazy binding, so force it once per thread.
void bind_primary_context() {
    thread_local bool bound = false;
    if (!bound) {
        cudaFree(nullptr);
        bound = true;
    }
}

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

}  // namespace

bool blur_tc_eligible(int planes, int H, int Wc, int ES, int radius, int rule, const float* x, const float* y) {
    const char* off = getenv("BGAN_BLUR_NO_TC");
    if (off && atoi(off)) return false;
    const char* min_r = getenv("BGAN_BLUR_TC_MIN_R");
    const int rmin = min_r ? atoi(min_r) : 0;
    return (ES == 1 || ES == 3) && H == kP && Wc == kP * ES && planes > 0 && radius >= rmin && radius <= 127 &&
           (rule == BGAN_RULE_TORCH || rule == BGAN_RULE_TF) && reinterpret_cast<uintptr_t>(x) % 16 == 0 &&
           reinterpret_cast<uintptr_t>(y) % 16 == 0 && encode_fn() != nullptr;
}

// `planes` counts geometry planes: N*C for NCHW (es = 1), N images of 128 x 384 interleaved floats for NHWC (es = 3)
int blur_tc_launch(const float* x, float* y, int planes, int es, int rule, float std_eff, int radius, cudaStream_t stream) {
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    bind_primary_context();
    // the tensor map is only used for es == 1; for NHWC it describes the same memory as dense 128-float rows (never loaded)
    const cuuint64_t rows = static_cast<cuuint64_t>(planes) * es;
    const cuuint64_t dims[3] = {kP, kP, rows};
    const cuuint64_t strides[2] = {kP * 4, kP * kP * 4};
    const cuuint32_t box[3] = {kP, kP, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult res = encode_fn()(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(x), dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    BGAN_REQUIRE(res == CUDA_SUCCESS, BGAN_E_UNSUPPORTED, "cuTensorMapEncodeTiled failed (%d)", static_cast<int>(res));
    thread_local int attr_dev = -1;
    int dev = 0;
    BGAN_CUDA_OK(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        for (int v = 0; v < kNumVariants; ++v)
            BGAN_CUDA_OK(cudaFuncSetAttribute(kVariants[v], cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmemTc)));
        attr_dev = dev;
    }
    const char* tr = getenv("BGAN_TC_TRACE_PTR");  // debug only: device pointer (decimal) to 128 int64
    const int units = planes * es;  // 128 x 128 channel planes
    TcArgs a{tr ? reinterpret_cast<long long*>(strtoull(tr, nullptr, 10)) : nullptr, x, y, es, units, rule, std_eff, radius, kDefaultPrefetch, 0};
    if (const char* v = getenv("BGAN_TC_DBG")) a.dbg = atoi(v);
    if (const char* v = getenv("BGAN_TC_PF")) a.pf = std::max(0, atoi(v));  // experiments only
    // balanced persistent grid: every CTA runs the same number of rounds (768 planes on 148 SMs -> 128 CTAs x 6 planes
    // rather than 28 CTAs x 6 + 120 x 5), so that the HBM-bound steady state has no latency-bound tail round
    const int sms = device_sm_count();
    const int rounds = ceil_div(units, sms);
    int grid_i = ceil_div(units, rounds);
    if (const char* g = getenv("BGAN_TC_GRID")) {
        if (atoi(g) > 0) grid_i = std::min(std::min(atoi(g), units), sms);
    }
    const unsigned grid = static_cast<unsigned>(grid_i);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kCtaThreads);
    cfg.dynamicSmemBytes = kSmemTc;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    const char* no_pdl = getenv("BGAN_BLUR_NO_PDL");
    cfg.attrs = attr;
    cfg.numAttrs = (no_pdl && atoi(no_pdl)) ? 0 : 1;
    int variant = kDefaultVariant;
    if (const char* v = getenv("BGAN_TC_VARIANT")) variant = std::min(std::max(atoi(v), 0), kNumVariants - 1);  // experiments only
    BGAN_CUDA_OK(cudaLaunchKernelEx(&cfg, kVariants[variant], a, tmap));
    return 0;
}

}  // namespace bgan

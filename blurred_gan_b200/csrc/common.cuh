// Shared host/device helpers for libbgan_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "bgan.h"

namespace bgan {

// ---- error plumbing (thread-local message behind bgan_last_error) ---------------------------
void set_error(const char* fmt, ...);
int device_sm_count();
int device_max_smem_optin();

#define BGAN_CUDA_OK(expr)                                                                      \
    do {                                                                                        \
        cudaError_t err__ = (expr);                                                             \
        if (err__ != cudaSuccess) {                                                             \
            ::bgan::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(err__), __FILE__, __LINE__); \
            return static_cast<int>(err__);                                                     \
        }                                                                                       \
    } while (0)

#define BGAN_REQUIRE(cond, code, ...)        \
    do {                                     \
        if (!(cond)) {                       \
            ::bgan::set_error(__VA_ARGS__);  \
            return (code);                   \
        }                                    \
    } while (0)

// Check the launch itself (not its completion): the ABI is asynchronous.
#define BGAN_LAUNCH_OK() BGAN_CUDA_OK(cudaGetLastError())

// Several equally shaped tensors blurred by ONE launch of the 128 x 128 kernels (bgan_blur2d_multi_f32: the critic step's
// real and generated batches).  Work units are numbered tensor after tensor.
constexpr int kMaxMulti = 4;
struct MultiIO {
    const float* x[kMaxMulti];
    float* y[kMaxMulti];
    int count;
};

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// ---- device helpers -------------------------------------------------------------------------
#ifdef __CUDACC__

// Programmatic dependent launch, used as "the same stream order, minus the launch gap": a kernel launched with launch_pdl may be
// scheduled while its predecessor in the stream is still running; pdl_wait() at its top (before its first global access) blocks
// until the predecessor has completed and its writes are visible.  Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// unit -> (tensor, unit inside the tensor) for a launch over NT tensors; `per` units per tensor.  NT = 1 (the single-tensor
// entry point) folds to nothing: the kernels are instantiated for NT = 1 and NT = kMaxMulti so that the hot single-tensor
// launch carries neither the extra tensor maps nor the selects (measured: 0.1 - 0.2 us per launch).
template <int NT>
__device__ __forceinline__ int multi_locate(int& unit, int per) {
    int t = 0;
#pragma unroll
    for (int k = 1; k < NT; ++k) {
        if (unit >= per) {
            unit -= per;
            ++t;
        }
    }
    return t;
}
// p[t] of a kernel-parameter array without dynamic indexing (which would copy the array to local memory)
template <int NT, typename T>
__device__ __forceinline__ T multi_pick(T const (&p)[NT], int t) {
    static_assert(NT == 1 || NT == kMaxMulti, "instantiated for one tensor or for kMaxMulti");
    if constexpr (NT == 1)
        return p[0];
    else
        return t == 0 ? p[0] : t == 1 ? p[1] : t == 2 ? p[2] : p[3];
}

// mbarrier + TMA (cp.async.bulk.tensor) primitives, written as the PTX they are.
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
// producer side of a named barrier: counts towards `threads` without waiting
__device__ __forceinline__ void named_bar_arrive(int id, int threads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x989680;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return done != 0;
}
// Bounded wait: a transfer that never lands (bad descriptor) traps instead of hanging the GPU.
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const uint64_t t0 = global_timer_ns();
    while (!mbar_try_wait(bar, parity))
        if (global_timer_ns() - t0 > 4000000000ull) __trap();  // 4 s
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 3-D tiled TMA load: box at (c0, c1, c2) of the tensor described by `tmap` -> dense smem box.
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// warm L2 with a box that a later tma_load_3d will fetch
__device__ __forceinline__ void tma_prefetch_l2_3d(const void* tmap, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0),
                 "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// host side of pdl_wait(): launch with programmatic stream serialization allowed
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("BGAN_NO_PDL") && atoi(getenv("BGAN_NO_PDL"));  // measurements
    cfg.attrs = attr;
    cfg.numAttrs = no_pdl ? 0 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#endif  // __CUDACC__

}  // namespace bgan

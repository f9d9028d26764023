// Micro-benchmark (round 2): is the ~105-cycle cost of an M128 N<=128 K16 tcgen05.mma (tools/mma_rate.cu) a property of the
// tensor pipe's throughput, or the latency of a chain of MMAs that accumulate into the SAME TMEM accumulator?
// Same set-up as tools/mma_rate.cu (one CTA per SM, one issuing thread, batches of 24 accumulating MMAs, one commit per batch,
// one batch in flight), but consecutive MMAs rotate over `chains` independent accumulators (disjoint TMEM column ranges).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_chains.bin tools/mma_chains.cu && tools/mma_chains.bin
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace {

constexpr int kBatches = 64, kPerBatch = 24;
constexpr uint32_t kSmem = 96 * 1024;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t idesc(int n) { return (1u << 4) | ((static_cast<uint32_t>(n) >> 3) << 17) | ((128u >> 4) << 24); }
__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(d),
        "l"(a), "l"(b), "r"(id), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t}\n" ::"r"(d),
        "r"(a), "l"(b), "r"(id), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ bool try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x989680;\n\tselp.b32 %0, 1, 0, P1;\n\t}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return done != 0;
}
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
    if (try_wait(bar, parity)) return;
    uint64_t t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (!try_wait(bar, parity)) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > 2000000000ull) __trap();
    }
}

// mode 0: SS, 1: TS.  A (TS) lives in TMEM columns [0, 64); accumulators start at column 64: chain c at 64 + c * n.
template <int CHAINS>
__global__ void __launch_bounds__(128, 1) mma_chain_kernel(int n, int mode, long long* cycles_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar[2];
    __shared__ uint32_t s_tmem;
    for (uint32_t i = threadIdx.x; i < kSmem / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[1])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    {
        const uint32_t taddr = tmem + (static_cast<uint32_t>((threadIdx.x >> 5) * 32) << 16);
        for (int c = 0; c < 64; c += 16) {
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n" ::"r"(taddr + c),
                "r"(0)
                : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (threadIdx.x == 0) {
        const uint32_t id = idesc(n);
        const uint32_t d0 = tmem + 64;
        const uint64_t adesc = desc_sw128(smem_u32(smem)), bdesc = desc_sw128(smem_u32(smem) + 16384);
        long long t0 = 0;
        int batch = 0;
        for (int rep = 0; rep < 2; ++rep) {
            t0 = clock64();
            for (int b = 0; b < kBatches; ++b, ++batch) {
#pragma unroll
                for (int k = 0; k < kPerBatch; ++k) {
                    const uint64_t ks = static_cast<uint64_t>((k & 3) * 2);
                    const uint32_t d = d0 + static_cast<uint32_t>((k % CHAINS) * n);
                    if (mode == 0)
                        mma_ss(d, adesc + ks, bdesc + ks, id, k >= CHAINS);
                    else
                        mma_ts(d, tmem + 8 * (k & 7), bdesc + ks, id, k >= CHAINS);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[batch & 1]))
                             : "memory");
                if (b > 0) wait_bar(&bar[(batch - 1) & 1], ((batch - 1) >> 1) & 1);
            }
            wait_bar(&bar[(batch - 1) & 1], ((batch - 1) >> 1) & 1);
        }
        cycles_out[blockIdx.x] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

template <int CHAINS>
void run(int n, int mode, int sms, long long* d_cycles, long long* h) {
    if (64 + CHAINS * n > 512) return;
    cudaFuncSetAttribute(mma_chain_kernel<CHAINS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmem));
    mma_chain_kernel<CHAINS><<<sms, 128, kSmem>>>(n, mode, d_cycles);
    const cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) {
        printf("chains %d mode %d N %d: %s\n", CHAINS, mode, n, cudaGetErrorString(err));
        exit(1);
    }
    cudaMemcpy(h, d_cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < sms; ++i) mx = h[i] > mx ? h[i] : mx;
    const double per = static_cast<double>(mx) / (kBatches * kPerBatch);
    printf("A from %s  N=%3d  chains=%d : %7.1f cycles per MMA   %6.1f per 128x128x16 equivalent\n", mode == 0 ? "smem" : "TMEM", n, CHAINS, per,
           per * 128.0 / n);
}


}  // namespace

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* d_cycles = nullptr;
    cudaMalloc(&d_cycles, sizeof(long long) * sms);
    long long* h = static_cast<long long*>(malloc(sizeof(long long) * sms));
    printf("tcgen05.mma kind::f16 M=128 K=16, %d SMs, %d MMAs per SM rotating over `chains` independent accumulators\n", sms, kBatches * kPerBatch);
    for (int mode = 0; mode < 2; ++mode)
        for (int n : {32, 64, 128, 224}) {
            run<1>(n, mode, sms, d_cycles, h);
            run<2>(n, mode, sms, d_cycles, h);
            run<3>(n, mode, sms, d_cycles, h);
            run<4>(n, mode, sms, d_cycles, h);
            run<6>(n, mode, sms, d_cycles, h);
        }
    return 0;
}

// Micro-benchmark: how many SM cycles does one tcgen05.mma (kind::f16, M = 128, K = 16, cta_group::1) cost as a function
// of N and of where the A operand comes from?  The tensor-core blur (csrc/blur_tc.cu) is bounded by its MMA stream
// (48 instructions of N = 128 per plane at ~100 cycles each, profiles/r01d_blur_tc_pipeline.md) and its N = 64 variant
// was no faster -- if N = 256 costs the same per instruction too, GEMM1 of two planes (same A = T) belongs in one
// instruction.  Operands are zeros (timing does not depend on the data).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/mma_rate.bin tools/mma_rate.cu && tools/mma_rate.bin
//
// One CTA per SM, 128 threads, one issuing thread.  Per configuration: `kBatches` batches of 24 accumulating MMAs, one
// tcgen05.commit per batch, the issuing thread waits for the last commit; cycles = clock64 around issue + completion.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace {

constexpr int kBatches = 64, kPerBatch = 24;
constexpr uint32_t kSmem = 96 * 1024;  // A tile (128 x 64 K, 16 KB) + B tile (256 x 64 K, 32 KB), 1024-byte aligned

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {  // K-major, SWIZZLE_128B, SBO = 1024 (as csrc/blur_tc.cu)
    return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t idesc(int n) {  // D = f32, A = B = f16, K-major, M = 128
    return (1u << 4) | ((static_cast<uint32_t>(n) >> 3) << 17) | ((128u >> 4) << 24);
}
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

// bounded wait: a batch that never completes traps after 2 s instead of hanging the GPU
__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t parity) {
    if (try_wait(bar, parity)) return;
    uint64_t t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (!try_wait(bar, parity)) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        if (t - t0 > 2000000000ull) __trap();
    }
}

// mode 0: SS (A from shared memory), 1: TS (A from TMEM)
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(int n, int mode, long long* cycles_out) {
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
    // zero the TMEM columns used as the A operand (TS mode reads columns [0, 64) as packed fp16 pairs)
    {
        const uint32_t taddr = tmem + (static_cast<uint32_t>((threadIdx.x >> 5) * 32) << 16);
        for (int c = 0; c < 128; c += 16) {
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
        const uint32_t d = tmem + 128;  // accumulator columns [128, 128 + n), n <= 256
        const uint64_t adesc = desc_sw128(smem_u32(smem)), bdesc = desc_sw128(smem_u32(smem) + 16384);
        long long t0 = 0;
        int batch = 0;  // running batch number: batch j commits to bar[j & 1], whose phase parity is (j >> 1) & 1
        for (int rep = 0; rep < 2; ++rep) {  // rep 0 warms up
            t0 = clock64();
            for (int b = 0; b < kBatches; ++b, ++batch) {
#pragma unroll
                for (int k = 0; k < kPerBatch; ++k) {
                    const uint64_t ks = static_cast<uint64_t>((k & 3) * 2);  // K-step inside the 64-element swizzle atom
                    if (mode == 0)
                        mma_ss(d, adesc + ks, bdesc + ks, id, k > 0);
                    else
                        mma_ts(d, tmem + 8 * (k & 7), bdesc + ks, id, k > 0);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[batch & 1]))
                             : "memory");
                if (b > 0) {  // one batch stays in flight behind the one just issued, as in a pipeline
                    wait_bar(&bar[(batch - 1) & 1], ((batch - 1) >> 1) & 1);
                }
            }
            wait_bar(&bar[(batch - 1) & 1], ((batch - 1) >> 1) & 1);  // the last batch of this repetition
        }
        const long long t1 = clock64();
        cycles_out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

// ---- cta_group::2: one dispatch drives the tensor cores of both SMs of a CTA pair (M = 256: 128 rows per SM; each CTA
// supplies its own A rows and half of the N rows of B).  Only the leader CTA issues; the commit is multicast to both.
__device__ __forceinline__ void mma2_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a),
                 "l"(b), "r"(id), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void mma2_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a),
                 "l"(b), "r"(id), "r"(acc)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) mma2_rate_kernel(int n, int mode, long long* cycles_out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint64_t bar[2];
    __shared__ uint32_t s_tmem;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    for (uint32_t i = threadIdx.x; i < kSmem / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[0])), "r"(1) : "memory");
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[1])), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;
    {
        const uint32_t taddr = tmem + (static_cast<uint32_t>((threadIdx.x >> 5) * 32) << 16);
        for (int c = 0; c < 128; c += 16) {
            asm volatile(
                "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};\n" ::"r"(taddr + c),
                "r"(0)
                : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();  // both CTAs: barriers initialised, TMEM allocated, operands zeroed
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    if (rank == 0 && threadIdx.x == 0) {
        const uint32_t id = (1u << 4) | ((static_cast<uint32_t>(n) >> 3) << 17) | ((256u >> 4) << 24);  // M = 256
        const uint32_t d = tmem + 128;
        const uint64_t adesc = desc_sw128(smem_u32(smem)), bdesc = desc_sw128(smem_u32(smem) + 16384);
        long long t0 = 0;
        int batch = 0;
        for (int rep = 0; rep < 2; ++rep) {
            t0 = clock64();
            for (int b = 0; b < kBatches; ++b, ++batch) {
#pragma unroll
                for (int k = 0; k < kPerBatch; ++k) {
                    const uint64_t ks = static_cast<uint64_t>((k & 3) * 2);
                    if (mode == 0)
                        mma2_ss(d, adesc + ks, bdesc + ks, id, k > 0);
                    else
                        mma2_ts(d, tmem + 8 * (k & 7), bdesc + ks, id, k > 0);
                }
                asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                                 smem_u32(&bar[batch & 1])),
                             "h"(static_cast<uint16_t>(3))
                             : "memory");
                if (b > 0) wait_bar(&bar[(batch - 1) & 1], ((batch - 1) >> 1) & 1);
            }
            wait_bar(&bar[(batch - 1) & 1], ((batch - 1) >> 1) & 1);
        }
        cycles_out[blockIdx.x >> 1] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster_sync_all();  // the partner CTA keeps its TMEM and barriers alive until the leader is done
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace

int main(int argc, char** argv) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    long long* d_cycles = nullptr;
    cudaMalloc(&d_cycles, sizeof(long long) * sms);
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmem));
    long long* h = static_cast<long long*>(malloc(sizeof(long long) * sms));
    printf("tcgen05.mma kind::f16 M=128 K=16, %d SMs busy, %d MMAs per SM (commit + wait every %d)\n", sms, kBatches * kPerBatch, kPerBatch);
    for (int mode = 0; mode < 2; ++mode) {
        for (int n : {32, 64, 128, 256}) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0);
            cudaEventCreate(&e1);
            cudaEventRecord(e0);
            mma_rate_kernel<<<sms, 128, kSmem>>>(n, mode, d_cycles);
            cudaEventRecord(e1);
            const cudaError_t err = cudaDeviceSynchronize();
            if (err != cudaSuccess) {
                printf("mode %d N %d: %s\n", mode, n, cudaGetErrorString(err));
                return 1;
            }
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            cudaMemcpy(h, d_cycles, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
            long long mx = 0, mn = 1ll << 60;
            for (int i = 0; i < sms; ++i) {
                mx = h[i] > mx ? h[i] : mx;
                mn = h[i] < mn ? h[i] : mn;
            }
            const double per = static_cast<double>(mx) / (kBatches * kPerBatch);
            printf("A from %s  N=%3d : %7.1f cycles per MMA (min SM %7.1f)   %6.1f cycles per 128x128x16 equivalent   kernel %.3f ms\n",
                   mode == 0 ? "smem" : "TMEM", n, per, static_cast<double>(mn) / (kBatches * kPerBatch), per * 128.0 / n, ms);
        }
    }
    if (argc > 1 && atoi(argv[1]) == 2) {
        const int pairs = sms / 2;
        cudaFuncSetAttribute(mma2_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kSmem));
        printf("tcgen05.mma cta_group::2 kind::f16 M=256 (128 rows per SM) K=16, %d CTA pairs, %d MMAs per pair\n", pairs, kBatches * kPerBatch);
        for (int mode = 0; mode < 2; ++mode) {
            for (int n : {64, 128, 256}) {
                mma2_rate_kernel<<<2 * pairs, 128, kSmem>>>(n, mode, d_cycles);
                const cudaError_t err = cudaDeviceSynchronize();
                if (err != cudaSuccess) {
                    printf("cta_group::2 mode %d N %d: %s\n", mode, n, cudaGetErrorString(err));
                    return 1;
                }
                cudaMemcpy(h, d_cycles, sizeof(long long) * pairs, cudaMemcpyDeviceToHost);
                long long mx = 0;
                for (int i = 0; i < pairs; ++i) mx = h[i] > mx ? h[i] : mx;
                const double per = static_cast<double>(mx) / (kBatches * kPerBatch);
                printf("A from %s  N=%3d : %7.1f cycles per dispatch = per SM and 128 x %d x 16 tile   (%6.1f per SM and 128x128x16 equivalent)\n",
                       mode == 0 ? "smem" : "TMEM", n, per, n, per * 128.0 / n);
            }
        }
    }
    return 0;
}

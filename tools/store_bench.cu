// Micro-benchmark of global-store patterns for the blur epilogue (debug aid, not part of the library).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/store_bench tools/store_bench.cu && /tmp/store_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

constexpr int kPlane = 128 * 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// mode 0: 4 warps, scalar st.cs, warp q writes floats [q*32, q*32+32) of every row (the blur kernel's pattern)
// mode 1: same, plain st
// mode 2: 16 warps, scalar st.cs, warp w: quadrant w&3, rows (w>>2)*32 ..
// mode 3: 4 warps, float4 st.cs: warp covers one 512-byte row per instruction, 32 rows each
// mode 4: 16 warps, float4
// mode 5: bulk copy smem -> global (one thread, 64 KB per plane)
// mode 6: 4 warps scalar, row order rotated per CTA
template <int MODE>
__global__ void __launch_bounds__(512, 1) store_kernel(float* y, int planes, float val) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (MODE == 5) {
        for (int i = tid; i < kPlane; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = val;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
    }
    for (int p = blockIdx.x; p < planes; p += gridDim.x) {
        float* yp = y + static_cast<size_t>(p) * kPlane;
        if (MODE == 0 || MODE == 1 || MODE == 6) {
            if (warp < 4) {
                const int rot = (MODE == 6) ? (blockIdx.x * 37) & 127 : 0;
#pragma unroll 32
                for (int j = 0; j < 128; ++j) {
                    float* a = yp + ((j + rot) & 127) * 128 + warp * 32 + lane;
                    if (MODE == 1) *a = val; else __stcs(a, val);
                }
            }
        } else if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) __stcs(yp + ((warp >> 2) * 32 + j) * 128 + (warp & 3) * 32 + lane, val);
        } else if (MODE == 3) {
            if (warp < 4) {
#pragma unroll
                for (int j = 0; j < 32; ++j) __stcs(reinterpret_cast<float4*>(yp + (warp * 32 + j) * 128) + lane, make_float4(val, val, val, val));
            }
        } else if (MODE == 4) {
#pragma unroll
            for (int j = 0; j < 8; ++j) __stcs(reinterpret_cast<float4*>(yp + (warp * 8 + j) * 128) + lane, make_float4(val, val, val, val));
        } else if (MODE == 5) {
            if (tid == 0) {
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(yp), "r"(smem_u32(smem)), "r"(kPlane * 4) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (MODE == 5 && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ---- copy skeleton: bulk loads of whole planes into a 2-deep staging ring, NW store warps write them back out -----------
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, 0x989680;\n\tselp.b32 %0, 1, 0, P1;\n\t}" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
    }
}
template <int NW, int VEC>
__global__ void __launch_bounds__(32 * NW + 32, 1) copy_kernel(const float* x, float* y, int planes) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * kPlane * 4);
    uint64_t* empty = full + 2;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int first = blockIdx.x, stride = gridDim.x;
    const int n_it = (planes - first + stride - 1) / stride;
    if (tid == 0) {
        for (int b = 0; b < 2; ++b) { mbar_init(full + b, 1); mbar_init(empty + b, NW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (warp == NW) {
        if (lane == 0) {
            for (int it = 0; it < n_it; ++it) {
                const int b = it & 1;
                if (it >= 2) mbar_wait(empty + b, ((it - 2) >> 1) & 1);
                mbar_expect(full + b, kPlane * 4);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem + b * kPlane * 4)),
                             "l"(x + static_cast<size_t>(first + it * stride) * kPlane), "r"(kPlane * 4), "r"(smem_u32(full + b)) : "memory");
            }
        }
    } else {
        for (int it = 0; it < n_it; ++it) {
            const int b = it & 1;
            mbar_wait(full + b, (it >> 1) & 1);
            const float* st = reinterpret_cast<const float*>(smem + b * kPlane * 4);
            float* yp = y + static_cast<size_t>(first + it * stride) * kPlane;
            if (VEC == 1) {
                constexpr int R = 128 / (NW / 4);  // rows per warp; warp covers quadrant warp & 3
#pragma unroll 8
                for (int j = 0; j < R; ++j) {
                    const int o = ((warp >> 2) * R + j) * 128 + (warp & 3) * 32 + lane;
                    __stcs(yp + o, st[o]);
                }
            } else {
                constexpr int R = 128 / NW;
#pragma unroll 8
                for (int j = 0; j < R; ++j) {
                    const int o = (warp * R + j) * 32 + lane;
                    __stcs(reinterpret_cast<float4*>(yp) + o, reinterpret_cast<const float4*>(st)[o]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(empty + b);
        }
    }
}
template <int NW, int VEC>
float run_copy(float* xs[], float* ys[], int nb, int planes, int grid, int iters) {
    const int sm = 2 * kPlane * 4 + 64;
    cudaFuncSetAttribute(copy_kernel<NW, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) copy_kernel<NW, VEC><<<grid, 32 * NW + 32, sm>>>(xs[i % nb], ys[i % nb], planes);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) copy_kernel<NW, VEC><<<grid, 32 * NW + 32, sm>>>(xs[i % nb], ys[i % nb], planes);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("error %s\n", cudaGetErrorString(err));
    return ms * 1e3f / iters;
}

template <int MODE>
float run(float* bufs[], int nb, int planes, int grid, int threads, int iters) {
    cudaFuncSetAttribute(store_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) store_kernel<MODE><<<grid, threads, 200 * 1024>>>(bufs[i % nb], planes, 1.0f);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) store_kernel<MODE><<<grid, threads, 200 * 1024>>>(bufs[i % nb], planes, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaError_t err = cudaGetLastError();
    if (err != cudaSuccess) printf("error %s\n", cudaGetErrorString(err));
    return ms * 1e3f / iters;
}

int main(int argc, char** argv) {
    const int planes = 768, nb = 8, iters = 40;
    float* bufs[nb];
    for (int i = 0; i < nb; ++i) cudaMalloc(&bufs[i], static_cast<size_t>(planes) * kPlane * 4);
    const double mb = planes * kPlane * 4 / 1e6;
    for (int grid : {128, 148}) {
        printf("grid %d\n", grid);
        float t;
        t = run<0>(bufs, nb, planes, grid, 512, iters); printf("  4 warps scalar .cs      %7.2f us  %6.0f GB/s\n", t, mb / t * 1e3);
        t = run<1>(bufs, nb, planes, grid, 512, iters); printf("  4 warps scalar plain    %7.2f us  %6.0f GB/s\n", t, mb / t * 1e3);
        t = run<6>(bufs, nb, planes, grid, 512, iters); printf("  4 warps scalar rotated  %7.2f us  %6.0f GB/s\n", t, mb / t * 1e3);
        t = run<2>(bufs, nb, planes, grid, 512, iters); printf("  16 warps scalar .cs     %7.2f us  %6.0f GB/s\n", t, mb / t * 1e3);
        t = run<3>(bufs, nb, planes, grid, 512, iters); printf("  4 warps float4 .cs      %7.2f us  %6.0f GB/s\n", t, mb / t * 1e3);
        t = run<4>(bufs, nb, planes, grid, 512, iters); printf("  16 warps float4 .cs     %7.2f us  %6.0f GB/s\n", t, mb / t * 1e3);
        t = run<5>(bufs, nb, planes, grid, 512, iters); printf("  bulk smem->global       %7.2f us  %6.0f GB/s\n", t, mb / t * 1e3);
    }
    {
        float* xs[nb];
        for (int i = 0; i < nb; ++i) cudaMalloc(&xs[i], static_cast<size_t>(planes) * kPlane * 4);
        for (int grid : {128, 148}) {
            printf("copy skeleton (bulk load -> smem -> store warps), grid %d, read+write bytes\n", grid);
            float t;
            t = run_copy<4, 1>(xs, bufs, nb, planes, grid, iters); printf("  4 store warps scalar    %7.2f us  %6.0f GB/s\n", t, 2 * mb / t * 1e3);
            t = run_copy<16, 1>(xs, bufs, nb, planes, grid, iters); printf("  16 store warps scalar   %7.2f us  %6.0f GB/s\n", t, 2 * mb / t * 1e3);
            t = run_copy<4, 4>(xs, bufs, nb, planes, grid, iters); printf("  4 store warps float4    %7.2f us  %6.0f GB/s\n", t, 2 * mb / t * 1e3);
            t = run_copy<16, 4>(xs, bufs, nb, planes, grid, iters); printf("  16 store warps float4   %7.2f us  %6.0f GB/s\n", t, 2 * mb / t * 1e3);
        }
    }
    {   // per-SM copy rate with L2-resident data: 16 CTAs, 96 planes (6 MB in, 6 MB out)
        float* xs1[1];
        cudaMalloc(&xs1[0], static_cast<size_t>(96) * kPlane * 4);
        float* ys1[1] = {bufs[0]};
        const double kb = 96 * kPlane * 4 / 1e3 / 16;
        float t;
        t = run_copy<4, 1>(xs1, ys1, 1, 96, 16, iters); printf("per-SM copy (16 CTAs, L2 resident): 4 store warps scalar %7.2f us  %6.1f B/ns/SM each way\n", t, kb / t);
        t = run_copy<16, 4>(xs1, ys1, 1, 96, 16, iters); printf("per-SM copy (16 CTAs, L2 resident): 16 store warps float4 %7.2f us  %6.1f B/ns/SM each way\n", t, kb / t);
        // plain device-to-device copy of one launch's bytes (50 MB -> 50 MB), rotating buffers
        float* xs[nb];
        for (int i = 0; i < nb; ++i) cudaMalloc(&xs[i], static_cast<size_t>(planes) * kPlane * 4);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        for (int i = 0; i < 3; ++i) cudaMemcpyAsync(bufs[i], xs[i], static_cast<size_t>(planes) * kPlane * 4, cudaMemcpyDeviceToDevice, 0);
        cudaEventRecord(e0);
        for (int i = 0; i < iters; ++i) cudaMemcpyAsync(bufs[i % nb], xs[i % nb], static_cast<size_t>(planes) * kPlane * 4, cudaMemcpyDeviceToDevice, 0);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("cudaMemcpyAsync D2D 50 MB: %7.2f us  %6.0f GB/s (read + write)\n", ms * 1e3 / iters, 2 * mb / (ms * 1e3 / iters) * 1e3);
    }
    // per-SM store rate: few CTAs, L2-resident target (16 CTAs x 48 planes of the same 6 MB buffer region)
    printf("per-SM rate (16 CTAs, 96 planes = 6 MB, L2 resident)\n");
    {
        const int pl = 96, g = 16;
        float* b1[1] = {bufs[0]};
        const double kb = pl * kPlane * 4 / 1e3 / g;  // KB per SM
        float t;
        t = run<0>(b1, 1, pl, g, 512, iters); printf("  4 warps scalar .cs      %7.2f us  %6.1f B/ns/SM\n", t, kb / t);
        t = run<2>(b1, 1, pl, g, 512, iters); printf("  16 warps scalar .cs     %7.2f us  %6.1f B/ns/SM\n", t, kb / t);
        t = run<3>(b1, 1, pl, g, 512, iters); printf("  4 warps float4 .cs      %7.2f us  %6.1f B/ns/SM\n", t, kb / t);
        t = run<4>(b1, 1, pl, g, 512, iters); printf("  16 warps float4 .cs     %7.2f us  %6.1f B/ns/SM\n", t, kb / t);
        t = run<5>(b1, 1, pl, g, 512, iters); printf("  bulk smem->global       %7.2f us  %6.1f B/ns/SM\n", t, kb / t);
    }
    return 0;
}

// FP32 FMA issue rates on sm_100a: three-register FFMA, FFMA with one operand from the constant bank (kernel parameter), FFMA2.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma_rate.bin ffma_rate.cu && ./ffma_rate.bin
#include <cstdio>
#include <cuda_runtime.h>

struct Taps { float g[16]; };
constexpr int kIters = 4096, kAcc = 16;

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, const Taps t, float seed) {
    float acc[kAcc], x[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) acc[i] = seed + i, x[i] = seed * (threadIdx.x + i);
    float gr[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) gr[i] = t.g[i] + seed * (MODE == 3 ? 1.0f : float(threadIdx.x));  // per-thread registers (MODE 3: uniform)
    for (int it = 0; it < kIters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < kAcc; ++i) acc[i] = fmaf(x[i], gr[i], acc[i]);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < kAcc; ++i) acc[i] = fmaf(x[i], t.g[i], acc[i]);  // constant bank operand
        } else {
#pragma unroll
            for (int i = 0; i < kAcc; i += 2) {
                unsigned long long a, b, c, d;
                asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(x[i]), "f"(x[i + 1]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(gr[i]), "f"(gr[i + 1]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(c) : "f"(acc[i]), "f"(acc[i + 1]));
                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[i]), "=f"(acc[i + 1]) : "l"(d));
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < kAcc; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int sms) {
    float* out;
    const int ctas = sms * 8;
    cudaMalloc(&out, ctas * 256 * sizeof(float));
    Taps t;
    for (int i = 0; i < 16; ++i) t.g[i] = 1.0f / (i + 2);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    k<MODE><<<ctas, 256>>>(out, t, 0.5f);
    cudaEventRecord(e0);
    k<MODE><<<ctas, 256>>>(out, t, 0.5f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = double(ctas) * 256 * kIters * kAcc;
    printf("%-28s %.3f ms  %.1f GFMA/s  = %.1f FMA per clock per SM at 1.965 GHz\n", name, ms, fma / ms / 1e6, fma / (ms * 1e-3) / 1.965e9 / sms);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    run<0>("FFMA, three registers", p.multiProcessorCount);
    run<1>("FFMA, uniform / constant operand", p.multiProcessorCount);
    run<2>("FFMA2, registers", p.multiProcessorCount);
    run<3>("FFMA2, uniform operand", p.multiProcessorCount);
    return 0;
}

// Issue-rate microbenchmark: scalar FFMA (register and constant operands) against the packed FFMA2 (fma.rn.f32x2) of sm_100.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ffma2_bench tools/ffma2_bench.cu && tools/ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long pack(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}

template <int MODE>
__global__ void __launch_bounds__(256) bench(int iters, const float* in, float* out, float ca, float cb) {
    const float y = in[threadIdx.x], z = in[threadIdx.x + 256];
    float x[8];
    unsigned long long p[8];
    #pragma unroll
    for (int k = 0; k < 8; k++) { x[k] = threadIdx.x*1e-3f + k; p[k] = pack(x[k], x[k] + 0.5f); }
    const unsigned long long y2 = pack(y, y*0.999f), z2 = pack(z, z*1.001f);
    for (int it = 0; it < iters; it++) {
        #pragma unroll
        for (int u = 0; u < 8; u++) {
            #pragma unroll
            for (int k = 0; k < 8; k++) {
                if (MODE == 0) x[k] = fmaf(x[k], ca, cb);            // constant operands
                else if (MODE == 1) x[k] = fmaf(x[k], y, z);         // three registers
                else p[k] = fma2(p[k], y2, z2);                      // packed, three register pairs
            }
        }
    }
    float s = 0.f;
    #pragma unroll
    for (int k = 0; k < 8; k++) s += x[k] + (float) (p[k] & 0xffff);
    if (s == 123.456f) out[0] = s;
}

int main() {
    float *in, *out;
    cudaMalloc(&in, 4096); cudaMemset(in, 0, 4096); cudaMalloc(&out, 16);
    int numSM; cudaDeviceGetAttribute(&numSM, cudaDevAttrMultiProcessorCount, 0);
    const int iters = 4096, blocks = numSM*8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[3] = {"FFMA const operands", "FFMA 3 registers", "FFMA2 packed"};
    for (int mode = 0; mode < 3; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            if (mode == 0) bench<0><<<blocks, 256>>>(iters, in, out, 0.999f, 0.001f);
            else if (mode == 1) bench<1><<<blocks, 256>>>(iters, in, out, 0.999f, 0.001f);
            else bench<2><<<blocks, 256>>>(iters, in, out, 0.999f, 0.001f);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        const double inst = (double) iters*64*256/32*blocks;      // warp instructions
        const double fma = inst*32*(mode == 2 ? 2 : 1);
        printf("%-22s %.3f ms  %.1f G warp-inst/s  %.1f TFLOP/s\n", names[mode], best, inst/best*1e-6, 2*fma/best*1e-9);
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}

// Microbenchmark: legacy warp-level mma.sync TF32 throughput on sm_100a and its accumulation rounding.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o mma_peak mma_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>

__device__ __forceinline__ void mmaTf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int CHAINS>
__global__ void __launch_bounds__(256) mmaLoop(float* out, int iters) {
    float acc[CHAINS][4];
    unsigned a[4], b[2];
    for (int i = 0; i < 4; i++) a[i] = __float_as_uint(1.0f + 0.001f*(threadIdx.x + i));
    for (int i = 0; i < 2; i++) b[i] = __float_as_uint(0.5f + 0.002f*(threadIdx.x + i));
    for (int c = 0; c < CHAINS; c++) for (int i = 0; i < 4; i++) acc[c][i] = 0.f;
    for (int it = 0; it < iters; it++) {
        #pragma unroll
        for (int c = 0; c < CHAINS; c++) mmaTf32(acc[c], a, b);
    }
    float s = 0.f;
    for (int c = 0; c < CHAINS; c++) for (int i = 0; i < 4; i++) s += acc[c][i];
    out[blockIdx.x*blockDim.x + threadIdx.x] = s;
}

// accumulation rounding probe: D = A*B + C with A row 0 = (1, 0, ...), B col 0 = (x, 0, ...), C = big.
// out = big + x computed by the tensor core, for x below half an ulp / above half an ulp of big.
__global__ void roundProbe(float* out) {
    const int lane = threadIdx.x, g = lane >> 2, t = lane & 3;
    const float xs[4] = {0.75f, -0.75f, 0.25f, 1.5f};     // in units of ulp(big) = 1 for big = 2^23 + 1
    for (int k = 0; k < 4; k++) {
        unsigned a[4] = {0, 0, 0, 0}, b[2] = {0, 0};
        float d[4] = {0, 0, 0, 0};
        if (g == 0 && t == 0) a[0] = __float_as_uint(1.0f);
        if (g == 0 && t == 0) b[0] = __float_as_uint(xs[k]);
        if (g == 0 && t == 0) d[0] = 8388609.0f;            // 2^23 + 1, ulp = 1
        mmaTf32(d, a, b);
        if (lane == 0) out[k] = d[0];
    }
    // sum of 8 products inside one MMA: 2^23 + eight times 0.25 -> exact 2^23 + 2; RN-per-add gives 2^23
    {
        unsigned a[4], b[2]; float d[4] = {0, 0, 0, 0};
        a[0] = (g == 0) ? __float_as_uint(1.0f) : 0; a[1] = 0; a[2] = (g == 0) ? __float_as_uint(1.0f) : 0; a[3] = 0;
        b[0] = (g == 0) ? __float_as_uint(0.25f) : 0; b[1] = (g == 0) ? __float_as_uint(0.25f) : 0;
        if (g == 0 && t == 0) d[0] = 8388608.0f;
        mmaTf32(d, a, b);
        if (lane == 0) out[4] = d[0];
    }
}

int main() {
    int dev = 0; cudaSetDevice(dev);
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, dev);
    float* out; cudaMalloc(&out, 1 << 24);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
    for (int cfg = 0; cfg < 3; cfg++) {
        const int blocksPerSM = cfg == 0 ? 1 : (cfg == 1 ? 2 : 4);
        const int grid = prop.multiProcessorCount*blocksPerSM;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            mmaLoop<8><<<grid, 256>>>(out, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0*16*8*8*8.0*iters*(256/32)*(double) grid;
        printf("mma.sync m16n8k8 tf32, 256 thr x %d CTA/SM, 8 chains: %.1f TFLOP/s (%s)\n", blocksPerSM, flop/ms*1e-9,
               cudaGetErrorString(cudaGetLastError()));
    }
    roundProbe<<<1, 32>>>(out);
    float h[5]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("rounding probe (big = 2^23+1): +0.75 -> %.1f (RN 8388610, RZ 8388609)   -0.75 -> %.1f (RN 8388608, RZ 8388608.x)\n", h[0], h[1]);
    printf("                               +0.25 -> %.1f                           +1.5  -> %.1f (RN 8388610/11, RZ 8388610)\n", h[2], h[3]);
    printf("2^23 + 8 x 0.25 in one MMA -> %.1f (exact 8388610; per-add rounding would give 8388608)\n", h[4]);
    return 0;
}

#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(256) k_ffma(int iters, float a, float b, float* out) {
    float x[8]; for (int i=0;i<8;i++) x[i]=threadIdx.x*1e-3f+i;
    for (int it=0; it<iters; it++) {
        #pragma unroll
        for (int u=0;u<16;u++) { 
            #pragma unroll
            for (int i=0;i<8;i++) x[i]=fmaf(x[i],a,b); }
    }
    float s=0; for (int i=0;i<8;i++) s+=x[i]; if (s==123.456f) out[0]=s;
}
// gemm-like: 3 distinct register operands, outer product 4x4
__global__ void __launch_bounds__(256) k_ffma_outer(int iters, float* out, const float* in) {
    float acc[8][8]; 
    for (int i=0;i<8;i++) for (int j=0;j<8;j++) acc[i][j]=0.f;
    float a[8], b[8];
    for (int i=0;i<8;i++) { a[i]=in[threadIdx.x+i]; b[i]=in[threadIdx.x+8+i]; }
    for (int it=0; it<iters; it++) {
        #pragma unroll
        for (int i=0;i<8;i++)
            #pragma unroll
            for (int j=0;j<8;j++) acc[i][j]=fmaf(a[i],b[j],acc[i][j]);
        #pragma unroll
        for (int i=0;i<8;i++) { a[i]+=1e-6f; b[i]-=1e-6f; }
    }
    float s=0; for (int i=0;i<8;i++) for (int j=0;j<8;j++) s+=acc[i][j]; if (s==123.456f) out[0]=s;
}
__global__ void __launch_bounds__(256) k_ffma2(int iters, float2 a, float2 b, float* out) {
    float2 x[8]; for (int i=0;i<8;i++) x[i]=make_float2(threadIdx.x*1e-3f+i, i);
    for (int it=0; it<iters; it++) {
        #pragma unroll
        for (int u=0;u<16;u++) { 
            #pragma unroll
            for (int i=0;i<8;i++) x[i]=__ffma2_rn(x[i],a,b); }
    }
    float s=0; for (int i=0;i<8;i++) s+=x[i].x+x[i].y; if (s==123.456f) out[0]=s;
}
__global__ void __launch_bounds__(256) k_ffma2_outer(int iters, float* out, const float2* in) {
    float2 acc[8][4]; 
    for (int i=0;i<8;i++) for (int j=0;j<4;j++) acc[i][j]=make_float2(0.f,0.f);
    float2 a[8], b[4];
    for (int i=0;i<8;i++) a[i]=in[threadIdx.x+i];
    for (int i=0;i<4;i++) b[i]=in[threadIdx.x+8+i];
    for (int it=0; it<iters; it++) {
        #pragma unroll
        for (int i=0;i<8;i++)
            #pragma unroll
            for (int j=0;j<4;j++) acc[i][j]=__ffma2_rn(a[i],b[j],acc[i][j]);
        #pragma unroll
        for (int i=0;i<8;i++) { a[i].x+=1e-6f; }
        #pragma unroll
        for (int i=0;i<4;i++) { b[i].y-=1e-6f; }
    }
    float s=0; for (int i=0;i<8;i++) for (int j=0;j<4;j++) s+=acc[i][j].x+acc[i][j].y; if (s==123.456f) out[0]=s;
}
__global__ void __launch_bounds__(256) k_dfma(int iters, double a, double b, double* out) {
    double x[8]; for (int i=0;i<8;i++) x[i]=threadIdx.x*1e-3+i;
    for (int it=0; it<iters; it++) {
        #pragma unroll
        for (int u=0;u<16;u++) { 
            #pragma unroll
            for (int i=0;i<8;i++) x[i]=fma(x[i],a,b); }
    }
    double s=0; for (int i=0;i<8;i++) s+=x[i]; if (s==123.456) out[0]=s;
}
template<class F> float timeit(F f) { cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); f(); cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); return ms; }
int main() {
    float* out; cudaMalloc(&out, 4096*8); cudaMemset(out,0,4096*8); double* dout=(double*)out;
    int blocks=148*8, iters=4096;
    float ms;
    ms=timeit([&]{k_ffma<<<blocks,256>>>(iters,0.999f,0.001f,out);}); printf("FFMA  (x=x*a+b)      : %.2f TFLOP/s\n", 2.0*128*iters*256.0*blocks/ms/1e9);
    ms=timeit([&]{k_ffma_outer<<<blocks,256>>>(iters*2,out,out);}); printf("FFMA  outer 8x8       : %.2f TFLOP/s\n", 2.0*64*iters*2*256.0*blocks/ms/1e9);
    ms=timeit([&]{k_ffma2<<<blocks,256>>>(iters,make_float2(0.999f,0.998f),make_float2(0.001f,0.002f),out);}); printf("FFMA2 (x=x*a+b)      : %.2f TFLOP/s\n", 4.0*128*iters*256.0*blocks/ms/1e9);
    ms=timeit([&]{k_ffma2_outer<<<blocks,256>>>(iters*2,out,(const float2*)out);}); printf("FFMA2 outer 8x4(x2)  : %.2f TFLOP/s\n", 4.0*32*iters*2*256.0*blocks/ms/1e9);
    ms=timeit([&]{k_dfma<<<blocks,256>>>(iters/4,0.999,0.001,dout);}); printf("DFMA                 : %.2f TFLOP/s\n", 2.0*128*(iters/4)*256.0*blocks/ms/1e9);
    return 0;
}

// gather inner-loop shape sweep (scalar FFMA): R rows (warp-uniform float4 coef) x A atoms (lane-distinct float4 z)
#include <cstdio>
#include <cuda_runtime.h>
#define STEPS 27
template<int R, int A, int UNR, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS,MINB) g(int iters, float* out) {
    extern __shared__ float4 dyn[]; float4* coef=dyn; float4* z=dyn+STEPS*R*(THREADS/32);
    for (int i=threadIdx.x;i<STEPS*R*(THREADS/32);i+=blockDim.x) coef[i]=make_float4(i*1e-3f,1.f,-1.f,0.5f);
    for (int i=threadIdx.x;i<STEPS*32*A;i+=blockDim.x) z[i]=make_float4(i*1e-4f,1.f,-1.f,0.5f);
    __syncthreads();
    const int lane=threadIdx.x&31, warp=threadIdx.x>>5;
    float2 U[R][A], V[R][A];
    for (int i=0;i<R;i++) for (int a=0;a<A;a++) { U[i][a]=make_float2(0,0); V[i][a]=make_float2(0,0); }
    const float4* cw = coef + warp*R*STEPS;
    for (int it=0; it<iters; it++) {
        #pragma unroll UNR
        for (int l=0;l<STEPS;l++) {
            float4 c[R], zz[A];
            #pragma unroll
            for (int i=0;i<R;i++) c[i]=cw[i*STEPS+l];
            #pragma unroll
            for (int a=0;a<A;a++) zz[a]=z[l*32*A+lane+32*a];
            #pragma unroll
            for (int i=0;i<R;i++)
                #pragma unroll
                for (int a=0;a<A;a++) {
                    U[i][a].x=fmaf(c[i].x,zz[a].x,U[i][a].x); U[i][a].x=fmaf(c[i].z,zz[a].y,U[i][a].x);
                    U[i][a].y=fmaf(c[i].y,zz[a].x,U[i][a].y); U[i][a].y=fmaf(c[i].w,zz[a].y,U[i][a].y);
                    V[i][a].x=fmaf(c[i].w,zz[a].z,V[i][a].x); V[i][a].x=fmaf(-c[i].y,zz[a].w,V[i][a].x);
                    V[i][a].y=fmaf(-c[i].z,zz[a].z,V[i][a].y); V[i][a].y=fmaf(c[i].x,zz[a].w,V[i][a].y);
                }
        }
    }
    float s=0; for (int i=0;i<R;i++) for (int a=0;a<A;a++) s+=U[i][a].x+U[i][a].y+V[i][a].x+V[i][a].y;
    if (s==123.456f) out[0]=s;
}
template<class F> float timeit(F f) { cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); f(); cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); return ms; }
template<int R,int A,int UNR,int THREADS,int MINB> void run(float* out) {
    int iters=400; size_t sm=16*(STEPS*R*(THREADS/32)+STEPS*32*A);
    cudaFuncSetAttribute(g<R,A,UNR,THREADS,MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
    float ms=timeit([&]{g<R,A,UNR,THREADS,MINB><<<148*MINB,THREADS,sm>>>(iters,out);});
    double fl=2.0*8*R*A*STEPS*iters*(double)THREADS*148*MINB;
    cudaFuncAttributes at; cudaFuncGetAttributes(&at, g<R,A,UNR,THREADS,MINB>);
    printf("R=%d A=%d unroll=%d threads=%d x%d regs=%d : %.2f TFLOP/s %s\n", R,A,UNR,THREADS,MINB,at.numRegs, fl/ms/1e9, cudaGetErrorString(cudaGetLastError()));
}
int main() {
    float* out; cudaMalloc(&out, 4096);
    run<4,4,1,384,1>(out); run<4,4,3,384,1>(out); run<4,4,9,384,1>(out); run<4,4,3,256,2>(out); run<4,4,1,256,2>(out);
    run<8,2,3,384,1>(out); run<2,8,3,384,1>(out); run<8,4,1,256,1>(out); run<4,8,1,256,1>(out); run<2,4,3,512,2>(out); run<4,2,3,512,2>(out);
    run<6,4,1,256,1>(out); run<4,6,1,256,1>(out); run<3,4,3,384,1>(out); run<4,3,3,384,1>(out);
    return 0;
}

// Inner-loop shape experiments: which register-tile shapes does ptxas schedule well on sm_100a?
#include <cstdio>
#include <cuda_runtime.h>
#define STEPS 32
// ---- gather-like, scalar FFMA: 4 rows (uniform float4 coef) x 4 atoms (lane-distinct float4 z): 128 FFMA / 8 LDS.128
__global__ void __launch_bounds__(384,1) g1(int iters, float* out) {
    extern __shared__ float4 dyn[]; float4* coef=dyn; float4* z=dyn+STEPS*4;
    for (int i=threadIdx.x;i<STEPS*4;i+=blockDim.x) coef[i]=make_float4(i*1e-3f,1.f,-1.f,0.5f);
    for (int i=threadIdx.x;i<STEPS*128;i+=blockDim.x) z[i]=make_float4(i*1e-4f,1.f,-1.f,0.5f);
    __syncthreads();
    const int lane=threadIdx.x&31;
    float2 U[4][4], V[4][4];
    for (int i=0;i<4;i++) for (int a=0;a<4;a++) { U[i][a]=make_float2(0,0); V[i][a]=make_float2(0,0); }
    for (int it=0; it<iters; it++) {
        #pragma unroll 3
        for (int l=0;l<STEPS;l++) {
            float4 c[4], zz[4];
            #pragma unroll
            for (int i=0;i<4;i++) c[i]=coef[l*4+i];
            #pragma unroll
            for (int a=0;a<4;a++) zz[a]=z[l*128+lane+32*a];
            #pragma unroll
            for (int i=0;i<4;i++)
                #pragma unroll
                for (int a=0;a<4;a++) {
                    U[i][a].x=fmaf(c[i].x,zz[a].x,U[i][a].x); U[i][a].x=fmaf(c[i].z,zz[a].y,U[i][a].x);
                    U[i][a].y=fmaf(c[i].y,zz[a].x,U[i][a].y); U[i][a].y=fmaf(c[i].w,zz[a].y,U[i][a].y);
                    V[i][a].x=fmaf(c[i].w,zz[a].z,V[i][a].x); V[i][a].x=fmaf(-c[i].y,zz[a].w,V[i][a].x);
                    V[i][a].y=fmaf(-c[i].z,zz[a].z,V[i][a].y); V[i][a].y=fmaf(c[i].x,zz[a].w,V[i][a].y);
                }
        }
    }
    float s=0; for (int i=0;i<4;i++) for (int a=0;a<4;a++) s+=U[i][a].x+U[i][a].y+V[i][a].x+V[i][a].y;
    if (s==123.456f) out[0]=s;
}
// ---- gather-like, FFMA2 paired over adjacent atoms: coef duplicated (uniform, 3 float4/row), z as (zc0,zc1,zs0,zs1),(lzc0,lzc1,lzs0,lzs1)
template<int R, int AP>   // R rows, AP atom pairs per lane
__global__ void __launch_bounds__(384,1) g2(int iters, float* out) {
    extern __shared__ float4 dyn[]; float4* coef=dyn; float4* z=dyn+STEPS*R*3;
    for (int i=threadIdx.x;i<STEPS*R*3;i+=blockDim.x) coef[i]=make_float4(i*1e-3f,1.f,-1.f,0.5f);
    for (int i=threadIdx.x;i<STEPS*32*AP*2;i+=blockDim.x) z[i]=make_float4(i*1e-4f,1.f,-1.f,0.5f);
    __syncthreads();
    const int lane=threadIdx.x&31;
    float2 Ur[R][AP], Ui[R][AP], Vr[R][AP], Vi[R][AP];
    for (int i=0;i<R;i++) for (int a=0;a<AP;a++) { Ur[i][a]=Ui[i][a]=Vr[i][a]=Vi[i][a]=make_float2(0,0); }
    for (int it=0; it<iters; it++) {
        #pragma unroll 2
        for (int l=0;l<STEPS;l++) {
            float4 c0[R], c1[R], c2[R], za[AP], zb[AP];
            #pragma unroll
            for (int i=0;i<R;i++) { c0[i]=coef[(l*R+i)*3]; c1[i]=coef[(l*R+i)*3+1]; c2[i]=coef[(l*R+i)*3+2]; }
            #pragma unroll
            for (int a=0;a<AP;a++) { za[a]=z[(l*AP+a)*64+lane]; zb[a]=z[(l*AP+a)*64+32+lane]; }
            #pragma unroll
            for (int i=0;i<R;i++)
                #pragma unroll
                for (int a=0;a<AP;a++) {
                    const float2 zc=make_float2(za[a].x,za[a].y), zs=make_float2(za[a].z,za[a].w);
                    const float2 lc=make_float2(zb[a].x,zb[a].y), ls=make_float2(zb[a].z,zb[a].w);
                    const float2 Ar=make_float2(c0[i].x,c0[i].y), Ai=make_float2(c0[i].z,c0[i].w);
                    const float2 Br=make_float2(c1[i].x,c1[i].y), Bi=make_float2(c1[i].z,c1[i].w);
                    const float2 nAi=make_float2(c2[i].x,c2[i].y), nBr=make_float2(c2[i].z,c2[i].w);
                    Ur[i][a]=__ffma2_rn(Ar,zc,Ur[i][a]); Ur[i][a]=__ffma2_rn(Br,zs,Ur[i][a]);
                    Ui[i][a]=__ffma2_rn(Ai,zc,Ui[i][a]); Ui[i][a]=__ffma2_rn(Bi,zs,Ui[i][a]);
                    Vr[i][a]=__ffma2_rn(Bi,lc,Vr[i][a]); Vr[i][a]=__ffma2_rn(nAi,ls,Vr[i][a]);
                    Vi[i][a]=__ffma2_rn(nBr,lc,Vi[i][a]); Vi[i][a]=__ffma2_rn(Ar,ls,Vi[i][a]);
                }
        }
    }
    float s=0; for (int i=0;i<R;i++) for (int a=0;a<AP;a++) s+=Ur[i][a].x+Ur[i][a].y+Ui[i][a].x+Ui[i][a].y+Vr[i][a].x+Vr[i][a].y+Vi[i][a].x+Vi[i][a].y;
    if (s==123.456f) out[0]=s;
}
// ---- S-like scalar: 2 rows (lane-distinct float4 a) x 7 cols (uniform float2 pairs in 4 float4): 112 FFMA / 6 LDS.128
__global__ void __launch_bounds__(128,2) s1(int iters, float* out) {
    extern __shared__ float4 dyn[]; float4* A=dyn; float4* B=dyn+STEPS*64;
    for (int i=threadIdx.x;i<STEPS*64;i+=blockDim.x) A[i]=make_float4(i*1e-4f,1.f,-1.f,0.5f);
    for (int i=threadIdx.x;i<STEPS*16;i+=blockDim.x) B[i]=make_float4(i*1e-3f,1.f,-1.f,0.5f);
    __syncthreads();
    const int lane=threadIdx.x&31, warp=threadIdx.x>>5;
    float acc[2][7][8];
    for (int i=0;i<2;i++) for (int c=0;c<7;c++) for (int k=0;k<8;k++) acc[i][c][k]=0;
    for (int it=0; it<iters; it++) {
        #pragma unroll 2
        for (int j=0;j<STEPS;j++) {
            const float4 a0=A[j*64+lane], a1=A[j*64+32+lane];
            float4 b[4];
            #pragma unroll
            for (int c=0;c<4;c++) b[c]=B[(j*4+warp)*4+c];
            #pragma unroll
            for (int c=0;c<7;c++) {
                const float zc=(c&1)?b[c/2].z:b[c/2].x, zs=(c&1)?b[c/2].w:b[c/2].y;
                acc[0][c][0]=fmaf(a0.x,zc,acc[0][c][0]); acc[0][c][1]=fmaf(a0.x,zs,acc[0][c][1]);
                acc[0][c][2]=fmaf(a0.y,zc,acc[0][c][2]); acc[0][c][3]=fmaf(a0.y,zs,acc[0][c][3]);
                acc[0][c][4]=fmaf(a0.z,zc,acc[0][c][4]); acc[0][c][5]=fmaf(a0.z,zs,acc[0][c][5]);
                acc[0][c][6]=fmaf(a0.w,zc,acc[0][c][6]); acc[0][c][7]=fmaf(a0.w,zs,acc[0][c][7]);
                acc[1][c][0]=fmaf(a1.x,zc,acc[1][c][0]); acc[1][c][1]=fmaf(a1.x,zs,acc[1][c][1]);
                acc[1][c][2]=fmaf(a1.y,zc,acc[1][c][2]); acc[1][c][3]=fmaf(a1.y,zs,acc[1][c][3]);
                acc[1][c][4]=fmaf(a1.z,zc,acc[1][c][4]); acc[1][c][5]=fmaf(a1.z,zs,acc[1][c][5]);
                acc[1][c][6]=fmaf(a1.w,zc,acc[1][c][6]); acc[1][c][7]=fmaf(a1.w,zs,acc[1][c][7]);
            }
        }
    }
    float s=0; for (int i=0;i<2;i++) for (int c=0;c<7;c++) for (int k=0;k<8;k++) s+=acc[i][c][k];
    if (s==123.456f) out[0]=s;
}
// ---- S-like FFMA2: a pairs (a0,a1),(a2,a3) lane-distinct; z duplicated uniform (zc,zc,zs,zs) per col: 56 FFMA2 / 2+7 LDS.128
template<int TN>
__global__ void __launch_bounds__(128,2) s2(int iters, float* out) {
    extern __shared__ float4 dyn[]; float4* A=dyn; float4* B=dyn+STEPS*64;
    for (int i=threadIdx.x;i<STEPS*64;i+=blockDim.x) A[i]=make_float4(i*1e-4f,1.f,-1.f,0.5f);
    for (int i=threadIdx.x;i<STEPS*4*TN;i+=blockDim.x) B[i]=make_float4(i*1e-3f,1.f,-1.f,0.5f);
    __syncthreads();
    const int lane=threadIdx.x&31, warp=threadIdx.x>>5;
    float2 acc[2][TN][4];
    for (int i=0;i<2;i++) for (int c=0;c<TN;c++) for (int k=0;k<4;k++) acc[i][c][k]=make_float2(0,0);
    for (int it=0; it<iters; it++) {
        #pragma unroll 2
        for (int j=0;j<STEPS;j++) {
            const float4 a0=A[j*64+lane], a1=A[j*64+32+lane];
            float4 b[TN];
            #pragma unroll
            for (int c=0;c<TN;c++) b[c]=B[(j*4+warp)*TN+c];
            const float2 p00=make_float2(a0.x,a0.y), p01=make_float2(a0.z,a0.w), p10=make_float2(a1.x,a1.y), p11=make_float2(a1.z,a1.w);
            #pragma unroll
            for (int c=0;c<TN;c++) {
                const float2 zc=make_float2(b[c].x,b[c].y), zs=make_float2(b[c].z,b[c].w);
                acc[0][c][0]=__ffma2_rn(p00,zc,acc[0][c][0]); acc[0][c][1]=__ffma2_rn(p00,zs,acc[0][c][1]);
                acc[0][c][2]=__ffma2_rn(p01,zc,acc[0][c][2]); acc[0][c][3]=__ffma2_rn(p01,zs,acc[0][c][3]);
                acc[1][c][0]=__ffma2_rn(p10,zc,acc[1][c][0]); acc[1][c][1]=__ffma2_rn(p10,zs,acc[1][c][1]);
                acc[1][c][2]=__ffma2_rn(p11,zc,acc[1][c][2]); acc[1][c][3]=__ffma2_rn(p11,zs,acc[1][c][3]);
            }
        }
    }
    float s=0; for (int i=0;i<2;i++) for (int c=0;c<TN;c++) for (int k=0;k<4;k++) s+=acc[i][c][k].x+acc[i][c][k].y;
    if (s==123.456f) out[0]=s;
}
#define SM(k,bytes) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes)
template<class F> float timeit(F f) { cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1); f(); cudaDeviceSynchronize(); cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms,e0,e1); return ms; }
int main() {
    float* out; cudaMalloc(&out, 4096); int iters=400; float ms; double fl;
    SM(g1,200000); SM((g2<4,2>),200000); SM((g2<4,4>),200000); SM((g2<8,2>),200000); SM((g2<2,2>),200000); SM(s1,200000); SM(s2<7>,200000); SM(s2<4>,200000);
    ms=timeit([&]{g1<<<148,384,16*(STEPS*4+STEPS*128)>>>(iters,out);}); fl=2.0*128*STEPS*iters*384.0*148; printf("gather scalar 4x4, 384thr x1    : %.2f TFLOP/s\n", fl/ms/1e9);
    ms=timeit([&]{g2<4,2><<<148,384,16*(STEPS*12+STEPS*128)>>>(iters,out);}); fl=2.0*128*STEPS*iters*384.0*148; printf("gather FFMA2 4 rows x 2 pairs   : %.2f TFLOP/s\n", fl/ms/1e9);
    ms=timeit([&]{g2<4,4><<<148,256,16*(STEPS*12+STEPS*256)>>>(iters,out);}); fl=2.0*256*STEPS*iters*256.0*148; printf("gather FFMA2 4 rows x 4 pairs,256: %.2f TFLOP/s\n", fl/ms/1e9);
    ms=timeit([&]{g2<8,2><<<148,256,16*(STEPS*24+STEPS*128)>>>(iters,out);}); fl=2.0*256*STEPS*iters*256.0*148; printf("gather FFMA2 8 rows x 2 pairs,256: %.2f TFLOP/s\n", fl/ms/1e9);
    ms=timeit([&]{g2<2,2><<<148*2,384,16*(STEPS*6+STEPS*128)>>>(iters,out);}); fl=2.0*64*STEPS*iters*384.0*148*2; printf("gather FFMA2 2 rows x 2 pairs   : %.2f TFLOP/s\n", fl/ms/1e9);
    ms=timeit([&]{s1<<<148*2,128,16*(STEPS*64+STEPS*16)>>>(iters,out);}); fl=2.0*112*STEPS*iters*128.0*148*2; printf("S scalar 2x7, 128thr x2         : %.2f TFLOP/s\n", fl/ms/1e9);
    ms=timeit([&]{s1<<<148*3,128,16*(STEPS*64+STEPS*16)>>>(iters,out);}); fl=2.0*112*STEPS*iters*128.0*148*3; printf("S scalar 2x7, 128thr x3         : %.2f TFLOP/s\n", fl/ms/1e9);
    ms=timeit([&]{s2<7><<<148*2,128,16*(STEPS*64+STEPS*28)>>>(iters,out);}); fl=2.0*112*STEPS*iters*128.0*148*2; printf("S FFMA2 2x7, 128thr x2          : %.2f TFLOP/s\n", fl/ms/1e9);
    ms=timeit([&]{s2<7><<<148*3,128,16*(STEPS*64+STEPS*28)>>>(iters,out);}); fl=2.0*112*STEPS*iters*128.0*148*3; printf("S FFMA2 2x7, 128thr x3          : %.2f TFLOP/s\n", fl/ms/1e9);
    ms=timeit([&]{s2<4><<<148*4,128,16*(STEPS*64+STEPS*16)>>>(iters,out);}); fl=2.0*64*STEPS*iters*128.0*148*4; printf("S FFMA2 2x4, 128thr x4          : %.2f TFLOP/s\n", fl/ms/1e9);
    return 0;
}

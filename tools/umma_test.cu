// Building-block test for the tensor-core k-space path: tcgen05.mma kind::tf32 with both operands in shared
// memory (K-major, no swizzle, canonical 8x16B core matrices written by ordinary st.shared), FP32 accumulators
// in TMEM, tcgen05.ld epilogue. Checks (1) the descriptor/layout conventions against a host GEMM, (2) the
// 3-product split (hi*hi + hi*lo + lo*hi) accuracy against FP64, (3) sustained MMA throughput.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o umma_test umma_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ float toTf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r); }

// K-major, SWIZZLE_NONE: core matrix = 8 rows x 16 B contiguous; LBO = byte distance between the two 16-byte
// K chunks of one MMA (K = 8 tf32), SBO = byte distance between 8-row groups.
__device__ __forceinline__ uint64_t makeDesc(uint32_t saddr, uint32_t lboBytes, uint32_t sboBytes) {
    uint64_t d = 0;
    d |= (uint64_t) ((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t) ((lboBytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t) ((sboBytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t) 1 << 46;                       // descriptor version (Blackwell)
    return d;                                      // base_offset 0, layout_type 0 = no swizzle
}
__host__ __device__ constexpr uint32_t makeIdescTf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}
__device__ __forceinline__ void ummaTf32(uint32_t tmemD, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmemD), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void ummaCommit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smemAddr(bar)) : "memory");
}
__device__ __forceinline__ void mbarInit(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(smemAddr(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmemLoad16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}

// D[128 x N] = A[128 x K] * B[N x K]^T, one CTA of 128 threads. mode 0: operands rounded to tf32 (one product);
// mode 1: three-product split. A, B, D row-major in global memory.
template <int N, int K>
__global__ void __launch_bounds__(128) gemmTest(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int mode) {
    constexpr int M = 128;
    extern __shared__ __align__(128) unsigned char smem[];
    float* aHi = reinterpret_cast<float*>(smem);       // [KC][M][4]
    float* aLo = aHi + M*K;
    float* bHi = aLo + M*K;                            // [KC][N][4]
    float* bLo = bHi + N*K;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmemBase;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int e = tid; e < M*K; e += 128) {
        const int r = e/K, k = e % K;
        const float x = A[e], hi = toTf32(x);
        aHi[((k >> 2)*M + r)*4 + (k & 3)] = hi;
        aLo[((k >> 2)*M + r)*4 + (k & 3)] = toTf32(x - hi);
    }
    for (int e = tid; e < N*K; e += 128) {
        const int r = e/K, k = e % K;
        const float x = B[e], hi = toTf32(x);
        bHi[((k >> 2)*N + r)*4 + (k & 3)] = hi;
        bLo[((k >> 2)*N + r)*4 + (k & 3)] = toTf32(x - hi);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) { mbarInit(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smemAddr(&tmemBase)), "n"(N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256))) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmemBase;

    if (tid == 0) {
        constexpr uint32_t idesc = makeIdescTf32(M, N);
        uint32_t acc = 0;
        for (int pass = 0; pass < (mode ? 3 : 1); pass++) {
            // small terms first: lo*hi, hi*lo, then hi*hi
            const float* a = (mode == 0) ? aHi : (pass == 0 ? aLo : aHi);
            const float* b = (mode == 0) ? bHi : (pass == 1 ? bLo : bHi);
            for (int k8 = 0; k8 < K/8; k8++) {
                const uint64_t ad = makeDesc(smemAddr(a + (size_t) k8*2*M*4), M*16, 128);
                const uint64_t bd = makeDesc(smemAddr(b + (size_t) k8*2*N*4), N*16, 128);
                ummaTf32(tmem, ad, bd, idesc, acc);
                acc = 1;
            }
        }
        ummaCommit(&bar);
    }
    mbarWait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmemLoad16(tmem + ((uint32_t) (warp*32) << 16) + c0, v);
        for (int j = 0; j < 16; j++) D[(size_t) tid*N + c0 + j] = v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256))) : "memory");
}

__device__ __forceinline__ void ummaTf32TS(uint32_t tmemD, uint32_t tmemA, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(tmemD), "r"(tmemA), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmemStore4(uint32_t taddr, float4 v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                 :: "r"(taddr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)) : "memory");
}

// Same GEMM with A held in TMEM (lane = row, column = k), written there by tcgen05.st; B in shared memory.
template <int N, int K>
__global__ void __launch_bounds__(128) gemmTestTS(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D) {
    constexpr int M = 128;
    extern __shared__ __align__(128) unsigned char smem[];
    float* bHi = reinterpret_cast<float*>(smem);       // [K/4][N][4]
    float* bLo = bHi + N*K;
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmemBase;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < N*K; e += 128) {
        const int r = e/K, k = e % K;
        const float x = B[e], hi = toTf32(x);
        bHi[((k >> 2)*N + r)*4 + (k & 3)] = hi;
        bLo[((k >> 2)*N + r)*4 + (k & 3)] = toTf32(x - hi);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) { mbarInit(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smemAddr(&tmemBase)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmemBase;
    const uint32_t tD = tmem, tAhi = tmem + 256, tAlo = tmem + 256 + K;
    const uint32_t laneBase = (uint32_t) (warp*32) << 16;
    for (int k = 0; k < K; k += 4) {
        float4 hi, lo;
        const float* a = A + (size_t) tid*K + k;
        hi.x = toTf32(a[0]); hi.y = toTf32(a[1]); hi.z = toTf32(a[2]); hi.w = toTf32(a[3]);
        lo.x = toTf32(a[0] - hi.x); lo.y = toTf32(a[1] - hi.y); lo.z = toTf32(a[2] - hi.z); lo.w = toTf32(a[3] - hi.w);
        tmemStore4(tAhi + laneBase + k, hi);
        tmemStore4(tAlo + laneBase + k, lo);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        constexpr uint32_t idesc = makeIdescTf32(M, N);
        uint32_t acc = 0;
        for (int pass = 0; pass < 3; pass++) {
            const uint32_t a = (pass == 0) ? tAlo : tAhi;
            const float* b = (pass == 1) ? bLo : bHi;
            for (int k8 = 0; k8 < K/8; k8++) {
                const uint64_t bd = makeDesc(smemAddr(b + (size_t) k8*2*N*4), N*16, 128);
                ummaTf32TS(tD, a + k8*8, bd, idesc, acc);
                acc = 1;
            }
        }
        ummaCommit(&bar);
    }
    mbarWait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmemLoad16(tD + laneBase + c0, v);
        for (int j = 0; j < 16; j++) D[(size_t) tid*N + c0 + j] = v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512) : "memory");
}

// throughput: every CTA issues `iters` MMAs of 128 x 256 x 8 on the same operands
template <int N, bool TS>
__global__ void __launch_bounds__(128) peakTest(int iters, float* sink) {
    constexpr int M = 128;
    extern __shared__ __align__(128) unsigned char smem[];
    float* a = reinterpret_cast<float*>(smem);        // [2][M][4]
    float* b = a + 2*M*4;                              // [2][N][4]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmemBase;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 2*M*4 + 2*N*4; e += 128) a[e] = toTf32(0.001f*(e % 97));
    // (the TS variant reads whatever is in tensor-memory columns 256..263: timing only)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) { mbarInit(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smemAddr(&tmemBase)), "n"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmemBase;
    if (tid == 0) {
        constexpr uint32_t idesc = makeIdescTf32(M, N);
        const uint64_t ad = makeDesc(smemAddr(a), M*16, 128), bd = makeDesc(smemAddr(b), N*16, 128);
        for (int i = 0; i < iters; i++) {
            if (TS) ummaTf32TS(tmem + (i & 1)*N, tmem + 256 + 8*(i % 7), bd, idesc, i > 1);
            else ummaTf32(tmem + (i & 1)*N, ad, bd, idesc, i > 1);
        }
        ummaCommit(&bar);
    }
    mbarWait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float v[16];
    tmemLoad16(tmem + ((uint32_t) (warp*32) << 16), v);
    if (v[0] == 123.456f) sink[tid] = v[1];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(512) : "memory");
}

template <int N, int K>
static void runCase(int mode) {
    constexpr int M = 128;
    std::vector<float> A(M*K), B(N*K), D(M*N);
    srand(1234 + N + K);
    for (auto& x : A) x = (float) (2.0*rand()/RAND_MAX - 1.0);
    for (auto& x : B) x = (float) (2.0*rand()/RAND_MAX - 1.0);
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size()*4)); CK(cudaMalloc(&dB, B.size()*4)); CK(cudaMalloc(&dD, D.size()*4));
    CK(cudaMemcpy(dA, A.data(), A.size()*4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size()*4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, D.size()*4));
    const size_t smem = (size_t) (2*M*K + 2*N*K)*4;
    CK(cudaFuncSetAttribute(gemmTest<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    gemmTest<N, K><<<1, 128, smem>>>(dA, dB, dD, mode);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size()*4, cudaMemcpyDeviceToHost));
    double maxErr = 0, rms = 0, ref2 = 0;
    for (int i = 0; i < M; i++) for (int j = 0; j < N; j++) {
        double s = 0;
        for (int k = 0; k < K; k++) s += (double) A[i*K + k]*(double) B[j*K + k];
        const double e = D[i*N + j] - s;
        maxErr = fmax(maxErr, fabs(e)); rms += e*e; ref2 += s*s;
    }
    printf("M=128 N=%d K=%d mode=%s: max abs err %.3e, rel RMS err %.3e\n", N, K, mode ? "3-product split" : "single tf32",
           maxErr, sqrt(rms/ref2));
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

template <int N, int K>
static void runCaseTS() {
    constexpr int M = 128;
    std::vector<float> A(M*K), B(N*K), D(M*N);
    srand(4321 + N + K);
    for (auto& x : A) x = (float) (2.0*rand()/RAND_MAX - 1.0);
    for (auto& x : B) x = (float) (2.0*rand()/RAND_MAX - 1.0);
    float *dA, *dB, *dD;
    CK(cudaMalloc(&dA, A.size()*4)); CK(cudaMalloc(&dB, B.size()*4)); CK(cudaMalloc(&dD, D.size()*4));
    CK(cudaMemcpy(dA, A.data(), A.size()*4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size()*4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, D.size()*4));
    const size_t smem = (size_t) (2*N*K)*4;
    CK(cudaFuncSetAttribute(gemmTestTS<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    gemmTestTS<N, K><<<1, 128, smem>>>(dA, dB, dD);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(D.data(), dD, D.size()*4, cudaMemcpyDeviceToHost));
    double maxErr = 0, rms = 0, ref2 = 0;
    for (int i = 0; i < M; i++) for (int j = 0; j < N; j++) {
        double s = 0;
        for (int k = 0; k < K; k++) s += (double) A[i*K + k]*(double) B[j*K + k];
        const double e = D[i*N + j] - s;
        maxErr = fmax(maxErr, fabs(e)); rms += e*e; ref2 += s*s;
    }
    printf("A-from-TMEM M=128 N=%d K=%d 3-product split: max abs err %.3e, rel RMS err %.3e\n", N, K, maxErr, sqrt(rms/ref2));
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

int main() {
    CK(cudaSetDevice(0));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    runCase<64, 16>(0);
    runCase<64, 16>(1);
    runCase<256, 56>(0);
    runCase<256, 56>(1);
    runCase<48, 64>(1);
    runCaseTS<128, 56>();
    runCaseTS<64, 112>();
    float* sink; CK(cudaMalloc(&sink, 4096));
    const size_t smem = (2*128*4 + 2*256*4)*4;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000;
#define PEAK(NN, TSV) \
    for (int rep = 0; rep < 2; rep++) { \
        CK(cudaEventRecord(e0)); \
        peakTest<NN, TSV><<<prop.multiProcessorCount, 128, smem>>>(iters, sink); \
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); \
        float ms; cudaEventElapsedTime(&ms, e0, e1); \
        if (rep) printf("tcgen05.mma kind::tf32 128x%dx8 %s: %.3f ms, %.1f ns per MMA, %.1f TFLOP/s\n", NN, TSV ? "A in TMEM" : "A in smem", ms, \
               ms*1e6/iters, 2.0*128*NN*8*(double) iters*prop.multiProcessorCount/ms*1e-9); \
    }
    PEAK(256, false) PEAK(256, true) PEAK(128, false) PEAK(128, true) PEAK(64, false) PEAK(64, true)
    return 0;
}

// Building-block test for integer tensor-core structure factors: tcgen05.mma kind::i8 (s8 x s8 -> s32, exact) with both
// operands in shared memory (K-major, no swizzle, 8 rows x 16 B core matrices = 16 int8 along K per row), int32
// accumulators in tensor memory. Checks (1) descriptor / layout conventions and exactness against a host GEMM,
// (2) the three-digit decomposition of 24-bit fixed-point operands (8 digit products, four weight groups) against exact
// integer arithmetic, (3) sustained MMA throughput by N.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o imma_test imma_test.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t makeDesc(uint32_t saddr, uint32_t lboBytes, uint32_t sboBytes) {
    uint64_t d = 0;
    d |= (uint64_t) ((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t) ((lboBytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t) ((sboBytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t) 1 << 46;
    return d;
}
// D = S32 (2 << 4), A = B = signed 8 bit (1 << 7, 1 << 10), K-major, dense
__host__ __device__ constexpr uint32_t makeIdescS8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}
__device__ __forceinline__ void ummaI8(uint32_t tmemD, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
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
__device__ __forceinline__ void tmemLoad16(uint32_t taddr, int (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ void tmemAlloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smemAddr(slot)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ void tmemFree(uint32_t t) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(t), "n"(COLS) : "memory");
}

// D[128 x N] (int32) = A[128 x K] * B[N x K]^T, int8 operands row-major in global memory; one CTA of 128 threads
template <int N, int K>
__global__ void __launch_bounds__(128) gemmI8(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int* __restrict__ D) {
    constexpr int M = 128;
    extern __shared__ __align__(128) unsigned char smem[];
    int8_t* a = reinterpret_cast<int8_t*>(smem);           // [K/16][M][16]
    int8_t* b = a + M*K;                                   // [K/16][N][16]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmemBase;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < M*K; e += 128) { const int r = e/K, k = e % K; a[((k >> 4)*M + r)*16 + (k & 15)] = A[e]; }
    for (int e = tid; e < N*K; e += 128) { const int r = e/K, k = e % K; b[((k >> 4)*N + r)*16 + (k & 15)] = B[e]; }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) { mbarInit(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmemAlloc<(N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256)))>(&tmemBase);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmemBase;
    if (tid == 0) {
        constexpr uint32_t idesc = makeIdescS8(M, N);
        for (int k32 = 0; k32 < K/32; k32++) {             // one MMA = 32 int8 along K = two 16-byte chunk columns
            const uint64_t ad = makeDesc(smemAddr(a + (size_t) k32*2*M*16), M*16, 128);
            const uint64_t bd = makeDesc(smemAddr(b + (size_t) k32*2*N*16), N*16, 128);
            ummaI8(tmem, ad, bd, idesc, k32 > 0);
        }
        ummaCommit(&bar);
    }
    mbarWait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c0 = 0; c0 < N; c0 += 16) {
        int v[16];
        tmemLoad16(tmem + ((uint32_t) (warp*32) << 16) + c0, v);
        for (int j = 0; j < 16; j++) D[(size_t) tid*N + c0 + j] = v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmemFree<(N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256)))>(tmem);
}

// throughput: every CTA issues `iters` MMAs of 128 x N x 32 on the same operands
template <int N>
__global__ void __launch_bounds__(128) peakI8(int iters, int* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmemBase;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (128 + N)*32/4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0x01010101u*(e & 3);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) { mbarInit(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmemAlloc<256>(&tmemBase);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmemBase;
    if (tid == 0) {
        constexpr uint32_t idesc = makeIdescS8(128, N);
        const uint64_t ad = makeDesc(smemAddr(smem), 128*16, 128), bd = makeDesc(smemAddr(smem + 128*32), N*16, 128);
        for (int i = 0; i < iters; i++) ummaI8(tmem, ad, bd, idesc, i > 0);
        ummaCommit(&bar);
    }
    mbarWait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    int v[16];
    tmemLoad16(tmem + ((uint32_t) (warp*32) << 16), v);
    if (v[0] == 123456789) sink[0] = v[1];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmemFree<256>(tmem);
}

template <int N, int K> int checkGemm() {
    std::vector<int8_t> A(128*K), B(N*K);
    for (auto& v : A) v = (int8_t) (rand() % 256 - 128);
    for (auto& v : B) v = (int8_t) (rand() % 256 - 128);
    int8_t *dA, *dB; int* dD;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dD, 128*N*4));
    CK(cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice));
    const size_t smem = (128 + N)*K;
    CK(cudaFuncSetAttribute(gemmI8<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
    gemmI8<N, K><<<1, 128, smem>>>(dA, dB, dD);
    CK(cudaDeviceSynchronize());
    std::vector<int> D(128*N);
    CK(cudaMemcpy(D.data(), dD, D.size()*4, cudaMemcpyDeviceToHost));
    long long bad = 0;
    for (int r = 0; r < 128; r++) for (int c = 0; c < N; c++) {
        int ref = 0;
        for (int k = 0; k < K; k++) ref += (int) A[r*K + k]*(int) B[c*K + k];
        if (ref != D[r*N + c]) bad++;
    }
    printf("kind::i8 GEMM 128 x %d x %d: %lld of %d elements differ from the exact host result\n", N, K, bad, 128*N);
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return bad != 0;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("%s, %d SMs\n", prop.name, prop.multiProcessorCount);
    int rc = 0;
    rc |= checkGemm<64, 32>(); rc |= checkGemm<64, 128>(); rc |= checkGemm<128, 64>(); rc |= checkGemm<256, 64>();
    int* sink; CK(cudaMalloc(&sink, 64));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int iters = 20000;
#define PEAK(NN) { const size_t smem = (128 + NN)*32; CK(cudaFuncSetAttribute(peakI8<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem)); \
        peakI8<NN><<<prop.multiProcessorCount, 128, smem>>>(100, sink); CK(cudaDeviceSynchronize()); \
        CK(cudaEventRecord(e0)); peakI8<NN><<<prop.multiProcessorCount, 128, smem>>>(iters, sink); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize()); \
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); \
        printf("kind::i8 128 x %3d x 32: %.1f ns per MMA per SM, %.0f TOP/s\n", NN, ms*1e6/iters, 2.0*128*NN*32*(double) iters*prop.multiProcessorCount/ms*1e-9); }
    PEAK(64) PEAK(128) PEAK(256)
    return rc;
}

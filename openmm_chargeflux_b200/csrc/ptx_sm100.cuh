// ptx_sm100.cuh -- inline-PTX helpers for sm_100a: mbarrier, bulk TMA, cp.async, tcgen05 (tensor memory,
// UMMA descriptors, MMA issue / commit). Conventions validated on hardware by tools/umma_test.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cfx {

__device__ __forceinline__ uint32_t smemU32(const void* p) { return (uint32_t) __cvta_generic_to_shared(p); }

// ---- mbarrier ----
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smemU32(bar)), "r"(count));
}
__device__ __forceinline__ void mbarFenceInit() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smemU32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarArrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smemU32(bar)) : "memory");
}
#ifndef MBAR_SUSPEND_HINT_NS
#define MBAR_SUSPEND_HINT_NS 0x989680u
#endif
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        // (suspend-time hint: the warp sleeps in the barrier unit until the phase completes or the hint expires, instead
        // of spinning through issue slots the other warps of its scheduler need)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(smemU32(bar)), "r"(parity), "r"(MBAR_SUSPEND_HINT_NS) : "memory");
    } while (!done);
}

// ---- bulk TMA (global -> shared, completion on an mbarrier) and cp.async ----
__device__ __forceinline__ void bulkLoad(void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smemU32(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smemU32(bar)) : "memory");
}
__device__ __forceinline__ void cpAsync16(void* dstSmem, const void* srcGlobal) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smemU32(dstSmem)), "l"(srcGlobal) : "memory");
}
__device__ __forceinline__ void cpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cpAsyncWait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TF32 split ----
__device__ __forceinline__ float roundTf32(float x) { uint32_t r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r); }

// x = hi + lo: hi = x rounded to TF32 (nearest, ties away; integer form of cvt.rna.tf32.f32 for finite inputs),
// lo = x - hi exactly (|lo| <= 2^-11 |x|, random sign). lo is NOT rounded here: the tensor core ignores the low 13
// mantissa bits of its FP32 containers, i.e. truncates lo to TF32, an error <= 2^-22 |x| whose sign follows lo's
// (unbiased). 3 instructions per value.
__device__ __forceinline__ void splitTf32(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
    lo = x - hi;
}
__device__ __forceinline__ void splitTf32(const float4& v, float4& h, float4& l) {
    splitTf32(v.x, h.x, l.x); splitTf32(v.y, h.y, l.y); splitTf32(v.z, h.z, l.z); splitTf32(v.w, h.w, l.w);
}

// ---- tcgen05: tensor memory ----
template <int COLS> __device__ __forceinline__ void tmemAlloc(uint32_t* slotInSmem) {       // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smemU32(slotInSmem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS> __device__ __forceinline__ void tmemFree(uint32_t taddr) {               // whole warp (the allocating one)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tcgen05FenceBefore() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05FenceAfter() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 32 lanes x 32 bit: thread i of the warp reads/writes lane (laneBase + i), consecutive columns
__device__ __forceinline__ void tmemLoad32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    #pragma unroll
    for (int i = 0; i < 32; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmemLoad16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    #pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmemStore4(uint32_t taddr, float4 v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};"
                 :: "r"(taddr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)) : "memory");
}
__device__ __forceinline__ void tmemWaitStore() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmemLoad16i(uint32_t taddr, int (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// 16 columns of zeros into the warp's 32 lanes
__device__ __forceinline__ void tmemStoreZero16(uint32_t taddr) {
    const uint32_t z = 0;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};" :: "r"(taddr), "r"(z) : "memory");
}

// one lane of a converged warp
__device__ __forceinline__ bool electOne() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}

// ---- tcgen05: MMA ----
// Shared-memory operand descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes, contiguous (128 B).
// lboBytes = distance between the two 16-byte K chunks of one MMA (K = 8 for TF32), sboBytes = distance
// between consecutive 8-row groups.
__device__ __forceinline__ uint64_t ummaSmemDesc(uint32_t saddr, uint32_t lboBytes, uint32_t sboBytes) {
    uint64_t d = 0;
    d |= (uint64_t) ((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t) ((lboBytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t) ((sboBytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t) 1 << 46;                       // descriptor version
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, both K-major, dense
__host__ __device__ constexpr uint32_t ummaIdescTf32(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void ummaTf32SS(uint32_t tmemD, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmemD), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// D[tmem] (+)= A[tmem: lane = row, column = k] * B[smem]^T
__device__ __forceinline__ void ummaTf32TS(uint32_t tmemD, uint32_t tmemA, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n"
                 :: "r"(tmemD), "r"(tmemA), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// integer MMA: D = S32, A = B = signed 8 bit, both K-major, dense; one instruction = 32 int8 along K (two 16-byte chunks)
__host__ __device__ constexpr uint32_t ummaIdescS8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t) (N >> 3) << 17) | ((uint32_t) (M >> 4) << 24);
}
__device__ __forceinline__ void ummaI8SS(uint32_t tmemD, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(tmemD), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void ummaCommit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(smemU32(bar)) : "memory");
}

} // namespace cfx

// kspace_tc.cu -- the two reciprocal-space contractions of piece (3) on the 5th-generation tensor cores.
//
// After the factorisation of kspace.cu both k-space sums are dense contractions:
//   gather             (Ur, Ui, Vr, Vi)[a][r] = sum_k Z[a][k] * C[r][comp][k],  k = (|nz| = l, cos|sin), K = 2 Kz,
//                      Z[a][2l] = cos(2 pi l z_a), Z[a][2l+1] = sin(..), coefficient rows Ur: (Ar, Br), Ui: (Ai, Bi),
//                      Vr: (l Bi, -l Ai), Vi: (-l Br, l Ar) (A, B as defined in coefficientKernel): [atoms x K] x [K x 4 rows]
//   structure factors  P[(comp,row)][(l,c|s)] = sum_atoms A[(comp,row)][atom] * Z[(l,c|s)][atom], A = q Ex(nx) x Ey(|ny|)
// The gather runs as tcgen05.mma kind::tf32 with FP32 accumulators in tensor memory, made FP32-accurate by the three-product
// split x = hi + lo: product ~ lo*hi + hi*lo + hi*hi (relative RMS error 1.7e-7 at K = 56, tools/umma_test.cu). The structure
// factors run as tcgen05.mma kind::i8 on signed base-256 digit planes of fixed-point operands with int32 accumulators: exact
// sums, which is what lets the same kernel serve the energy call (structureFactorI8Kernel below; tools/imma_test.cu).
// Operands are written in the K-major, no-swizzle core-matrix layout (8 rows x 16 bytes contiguous) that the UMMA
// shared-memory descriptor addresses with two strides, so tiles can be produced by ordinary stores or one bulk-TMA copy.
//
// gatherTensorKernel<MT, NT>: one persistent CTA per SM, work units (MT x 128 atoms, NT/4 signed rows) in contiguous
// atom-major ranges.
//   warp 0        bulk-TMA ring of FP32 coefficient tiles
//   warps 2-3     MMA issuers, one per accumulator slot: A = phase operand in TENSOR MEMORY (hi, lo; written with
//                 tcgen05.st by the epilogue warps when the atom group changes), B = coefficient hi/lo planes in shared memory
//   warps 4-19    split the arrived FP32 tile into the TF32 hi/lo operand planes (double buffered); then, for 8 rows each:
//                 tcgen05.ld the accumulators, release the slot, T = Ex Ey, dE/dq += Re(T U), F += q g (nx Im TU, ny Im TU,
//                 Im TU') in registers; one fixed-point atomic per atom/output/warp when the atom group changes
// structureFactorI8Kernel<NN, TT, ND>: CTA = TT row tiles x one split of the atoms; ND = 3 digit planes (forces-only call) or 4
//   (energy call).  warp 0: bulk-TMA ring of 32 atoms' phase rows per stage;  warp 1: MMA issuer, 3-4 MMAs of K = 32 atoms
//   per row tile and stage into four int32 weight groups that stay in tensor memory until the end;  warps 2-17: form the
//   digit planes of both operands (fixed point inside an FMA, PRMT byte gathers), then the one-off 64-bit epilogue.
// structureFactorTensorKernel<NN, TT> (TF32 x 3; fall-back, CFX_KSPACE_S=tf32): CTA = TT row tiles x one split of the atoms.
//   warp 0        bulk-TMA of 32 atoms' phase rows per stage
//   warps 2-9     form both operands (products + hi/lo split) in shared memory
//   warps 1, 18   MMA issuers, one per operand buffer / accumulator slot
//   warps 10-17   tcgen05.ld the slot after every stage and sum it in FP32 registers with round-to-nearest: the tensor
//                 core TRUNCATES on accumulation, so long sums must not stay in tensor memory (and evaluations that
//                 return the energy use the FP32 kernel of kspace.cu: the residual bias of |S|^2 is ~2e-7)
// CFX_GT_TRACE=1 records %globaltimer stamps of the gather's roles (tools/gt_trace.py).
#include "cfx_internal.cuh"
#include "ptx_sm100.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace cfx {

namespace {

inline float __int_as_float_host(int v) { float f; memcpy(&f, &v, 4); return f; }
__device__ __forceinline__ unsigned long long globalTimer() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// role timeline of the tensor gather (tools/gt_trace.py): compiled in only with -DCFX_GT_TRACE_BUILD (the stamps and
// their predicates cost the epilogue loop ~6 % of its instructions even when the trace buffer is absent)
#ifdef CFX_GT_TRACE_BUILD
#define GT_STAMP(slot) do { if (p.trace && lane == 0) p.trace[blockIdx.x*32 + (slot)] = globalTimer(); } while (0)
#else
#define GT_STAMP(slot) do { } while (0)
#endif
__device__ __forceinline__ void namedBarrier(int id, int threads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory"); }

constexpr int GT_MMA_WARP = 2;               // warps 2-3: MMA issuers, one per accumulator slot
constexpr int GT_EPI_WARP0 = 4;              // up to 16 epilogue warps (20 warps: 96 registers per thread)
constexpr int GT_THREADS = (GT_EPI_WARP0 + 16)*32;
constexpr int GT_TILE_ATOMS = 128;

struct GTParams {
    const float* zSplit;        // [atom tile][KC][128][4]   FP32 (cos, sin)(2 pi l z), split into TF32 hi/lo when loaded
    const float* coefT;         // [column tile][KC][NT][4]  FP32 coefficients in core-matrix order, split in the kernel
    const float4* rowData; const int4* groupInfo; const float2* colX; const float2* colY; const float* qf;
    int Ky, Kp, KC, N, Npad;
    int signedLo, signedHi, numColTiles, numAtomGroups;
    float fx, fy, fz;
    int rawStages;              // FP32 coefficient tiles in flight (bulk TMA ring)
    uint32_t planeBytes, offRaw, offEy, offBar, offRd;
    unsigned long long* trace;      // optional [gridDim][32] globaltimer stamps (CFX_GT_TRACE)
};

template <int MT, int NT>
__global__ void __launch_bounds__(GT_THREADS, 1) gatherTensorKernel(GTParams p, long long* __restrict__ forceFixed, long long* __restrict__ dedqFixed) {
    constexpr int ROWS = NT/4;                       // signed rows per column tile
    constexpr int SUBS = ROWS/8;                     // epilogue warps per lane quarter, 8 rows each
    extern __shared__ __align__(128) unsigned char smem[];
    float2* Eys = reinterpret_cast<float2*>(smem + p.offEy);                 // [Ky][MT*128], thread-private columns
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.offBar);
    uint64_t* rawFull = bars;                        // [rawStages]  TMA -> splitters (the epilogue warps)
    uint64_t* rawEmpty = bars + p.rawStages;         // [rawStages]  splitters -> TMA
    uint64_t* opFull = rawEmpty + p.rawStages;       // [2]          splitters -> MMA
    uint64_t* opEmpty = opFull + 2;                  // [2]          MMA -> splitter
    uint64_t* dFull = opEmpty + 2;                   // [2]          MMA -> epilogue
    uint64_t* dEmpty = dFull + 2;                    // [2]          epilogue -> MMA
    uint64_t* aFull = dEmpty + 2;                    // [1]          epilogue (phase operand in TMEM) -> MMA
    uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(aFull + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long totalUnits = (long long) p.numAtomGroups*p.numColTiles;
    const int u0 = (int) (totalUnits*blockIdx.x/gridDim.x), u1 = (int) (totalUnits*(blockIdx.x + 1)/gridDim.x);

    if (tid == 0) {
        for (int s = 0; s < p.rawStages; s++) { mbarInit(&rawFull[s], 1); mbarInit(&rawEmpty[s], 4*SUBS); }
        for (int b = 0; b < 2; b++) { mbarInit(&opFull[b], 4*SUBS); mbarInit(&opEmpty[b], MT == 2 ? 2 : 1); }
        for (int d = 0; d < 2; d++) { mbarInit(&dFull[d], 1); mbarInit(&dEmpty[d], 4*SUBS); }      // one arrival per epilogue warp
        mbarInit(aFull, 4*SUBS);
        mbarFenceInit();
    }
    if (warp == 0) tmemAlloc<512>(tmemSlot);
    tcgen05FenceBefore();
    __syncthreads();
    tcgen05FenceAfter();
    const uint32_t tmem = *tmemSlot;
    // tensor-memory map: accumulator slots at columns 0 and 128, phase operand from column 256:
    // tile t: hi at 256 + t*2*Kp, lo at 256 + t*2*Kp + Kp
    const uint32_t tmemA = tmem + 256;
    if (warp == 0) GT_STAMP(0);

    if (warp == 0) {
        // ---------------- producer ----------------
        // one lane keeps rawStages FP32 coefficient tiles in flight (bulk TMA); a stage is refilled as soon as the
        // epilogue warps have split it into the TF32 operand planes
        if (lane == 0) {
            const uint32_t planeFloats = p.planeBytes/4;
            unsigned char* raw = smem + p.offRaw;
            int rs = 0; uint32_t rph = 0;
            int colTile = u0 % p.numColTiles;                        // (no division in the loops: units advance column tile by column tile)
            for (int unit = u0; unit < u1; unit++) {
                if (unit - u0 >= p.rawStages) mbarWait(&rawEmpty[rs], rph ^ 1);
                mbarExpectTx(&rawFull[rs], p.planeBytes);
                bulkLoad(raw + (size_t) rs*p.planeBytes, p.coefT + (size_t) colTile*planeFloats, p.planeBytes, &rawFull[rs]);
                if (++rs == p.rawStages) { rs = 0; rph ^= 1; }
                if (++colTile == p.numColTiles) colTile = 0;
            }
        }
        __syncwarp();
    }
    else if (warp == GT_MMA_WARP || warp == GT_MMA_WARP + 1) {
        // ---------------- MMA issuers ----------------
        // Two warps, one per accumulator slot (seq parity), so that the per-tile bookkeeping of one overlaps the
        // issue of the other: the tensor pipe idles whenever nobody is issuing.
        const uint32_t mySlot = warp - GT_MMA_WARP;
        // The whole warp runs the (warp-uniform) control flow so that descriptors live in uniform registers;
        // one elected lane issues the tensor-core instructions.
        constexpr uint32_t idesc = ummaIdescTf32(128, NT);
        const uint32_t planeBytes = p.planeBytes;                    // one hi or lo plane of an operand buffer
        const int k8n = p.Kp >> 3;
        int ob = 0; uint32_t oph = 0, aPh = 0;
        int curGroup = -1;
        uint32_t seq = 0;
        int group = u0/p.numColTiles, colTile = u0 - group*p.numColTiles;
        for (int unit = u0; unit < u1; unit++, colTile++) {
            if (colTile == p.numColTiles) { colTile = 0; group++; }
            if (group != curGroup) {
                curGroup = group;
                mbarWait(aFull, aPh); aPh ^= 1;
                GT_STAMP(aPh ? 20 : 21);
            }
            if (seq == 20) GT_STAMP(28);
            mbarWait(&opFull[ob], oph);
            if (seq == 20) GT_STAMP(29);
            tcgen05FenceAfter();
            const uint32_t stageAddr = smemU32(smem) + (uint32_t) ob*2*planeBytes;
            const uint64_t bHi = ummaSmemDesc(stageAddr, NT*16, 128), bLo = ummaSmemDesc(stageAddr + planeBytes, NT*16, 128);
            bool issued = false;
            #pragma unroll 1
            for (int t = 0; t < MT; t++, seq++) {
                const uint32_t d = seq & 1;
                if (d != mySlot) continue;
                issued = true;
                if (seq == 20) GT_STAMP(22);
                mbarWait(&dEmpty[d], ((seq >> 1) & 1) ^ 1);
                if (seq == 20) GT_STAMP(23);
                tcgen05FenceAfter();
                const uint32_t tD = tmem + d*128;
                const uint32_t aHi = tmemA + (uint32_t) t*2*p.Kp, aLo = aHi + p.Kp;
                if (electOne()) {
                    // small products first: Zlo Chi, Zhi Clo, then Zhi Chi; one k8 step advances the
                    // descriptor by two core-matrix columns (2*NT*16 bytes) and the phase operand by 8 columns
                    #pragma unroll 1
                    for (int k8 = 0; k8 < k8n; k8++) ummaTf32TS(tD, aLo + 8*k8, bHi + (uint64_t) (k8*(2*NT*16 >> 4)), idesc, k8 > 0);
                    #pragma unroll 1
                    for (int k8 = 0; k8 < k8n; k8++) ummaTf32TS(tD, aHi + 8*k8, bLo + (uint64_t) (k8*(2*NT*16 >> 4)), idesc, 1);
                    #pragma unroll 1
                    for (int k8 = 0; k8 < k8n; k8++) ummaTf32TS(tD, aHi + 8*k8, bHi + (uint64_t) (k8*(2*NT*16 >> 4)), idesc, 1);
                }
                __syncwarp();
                if (seq == 18 || seq == 20) GT_STAMP(seq == 18 ? 24 : 25);
                if (electOne()) ummaCommit(&dFull[d]);
            }
            if (issued && electOne()) ummaCommit(&opEmpty[ob]);   // the operand buffer is free once these MMAs have read it
            __syncwarp();
            if (++ob == 2) { ob = 0; oph ^= 1; }
        }
    }
    else if (warp >= GT_EPI_WARP0 && warp - GT_EPI_WARP0 < 4*SUBS) {
        // ---------------- epilogue: thread = (tensor-memory lane = atom of the tile, group of 8 rows) ----------------
        const int q4 = warp & 3;                                  // lane quarter this warp may access
        const int sub = (warp - GT_EPI_WARP0) >> 2;                          // which 8 rows (32 accumulator columns) of the tile
        const int atomInTile = q4*32 + lane;
        const uint32_t laneBase = (uint32_t) (q4*32) << 16;
        constexpr int EPI_THREADS = 128*SUBS;
        float oD[MT], oX[MT], oY[MT], oZ[MT];
        float2 ex0[MT], ex1[MT], ex0n[MT], ex1n[MT];        // Ex of the first / last row of this warp's 8 rows: current unit, next unit
        float4* rdS = reinterpret_cast<float4*>(smem + p.offRd) + (warp - GT_EPI_WARP0)*16;    // [2][8] row data of this warp's rows

        int curGroup = -1;
        uint32_t seq = 0;
        auto flush = [&]() {
            #pragma unroll
            for (int t = 0; t < MT; t++) {
                const int atom = (curGroup*MT + t)*GT_TILE_ATOMS + atomInTile;
                if (atom < p.N) {
                    const double q = (double) p.qf[atom];
                    atomicAddFixed(dedqFixed + atom, (double) oD[t]);
                    atomicAddFixed(forceFixed + atom, q*(double) p.fx*(double) oX[t]);
                    atomicAddFixed(forceFixed + p.Npad + atom, q*(double) p.fy*(double) oY[t]);
                    atomicAddFixed(forceFixed + 2*(size_t) p.Npad + atom, q*(double) p.fz*(double) oZ[t]);
                }
            }
        };
        // this warp's share of splitting the coefficient tile of unit v into the TF32 hi / lo operand planes
        const uint32_t eid = (warp - GT_EPI_WARP0)*32 + lane, planeVec = p.planeBytes/16;
        auto splitUnit = [&](int v) {
            const int k = v - u0, rs = k % p.rawStages, ob = k & 1;
            mbarWait(&rawFull[rs], (k/p.rawStages) & 1);
            mbarWait(&opEmpty[ob], ((k >> 1) & 1) ^ 1);             // the MMAs that read this operand buffer are complete
            const float4* src = reinterpret_cast<const float4*>(smem + p.offRaw + (size_t) rs*p.planeBytes);
            float4* hi = reinterpret_cast<float4*>(smem + (size_t) ob*2*p.planeBytes);
            float4* lo = hi + planeVec;
            #pragma unroll 1
            for (uint32_t e = eid; e < planeVec; e += 2*EPI_THREADS) {          // two independent 16-byte chunks per pass
                const bool two = e + EPI_THREADS < planeVec;
                const float4 a = src[e], b = two ? src[e + EPI_THREADS] : a;
                float4 h, l;
                splitTf32(a, h, l); hi[e] = h; lo[e] = l;
                if (two) { splitTf32(b, h, l); hi[e + EPI_THREADS] = h; lo[e + EPI_THREADS] = l; }
            }
            fenceProxyAsync();                                    // operand planes visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) { mbarArrive(&opFull[ob]); mbarArrive(&rawEmpty[rs]); }
        };
        int buf = 0, cc = 0;
        if (u0 < u1) splitUnit(u0);
        if (u0 < u1) {
            const int group0 = u0/p.numColTiles, g8 = (u0 - group0*p.numColTiles)*SUBS + sub;
            const int4 gi = __ldg(p.groupInfo + g8);
            if (lane < 8) rdS[lane] = __ldg(p.rowData + p.signedLo + 8*g8 + lane);
            #pragma unroll
            for (int t = 0; t < MT; t++) {
                const float2* exCol = p.colX + (group0*MT + t)*GT_TILE_ATOMS + atomInTile;
                ex0[t] = __ldg(exCol + gi.x); ex1[t] = __ldg(exCol + gi.y);
            }
            cc = gi.z;
            __syncwarp();
        }
        int group = u0/p.numColTiles, colTile = u0 - group*p.numColTiles;
        for (int unit = u0; unit < u1; unit++, colTile++) {
            if (colTile == p.numColTiles) { colTile = 0; group++; }
            if (group != curGroup) {
                const int gs = (curGroup < 0) ? 1 : 8;
                if (warp == GT_EPI_WARP0) GT_STAMP(gs);
                if (curGroup >= 0) flush();
                if (warp == GT_EPI_WARP0) GT_STAMP(gs + 1);
                namedBarrier(1, EPI_THREADS);
                if (warp == GT_EPI_WARP0) GT_STAMP(gs + 2);                     // every epilogue warp is done with the old Ey columns
                curGroup = group;
                // phase operand of the new atom group -> tensor memory (all MMAs that read the old one are
                // complete: their last accumulator has been consumed), Ey columns -> shared memory
                #pragma unroll
                for (int t = 0; t < MT; t++) {
                    const int tile = group*MT + t;
                    const float4* src = reinterpret_cast<const float4*>(p.zSplit) + (size_t) tile*p.KC*GT_TILE_ATOMS + atomInTile;
                    const uint32_t dst = tmemA + (uint32_t) t*2*p.Kp + laneBase;
                    for (int c0 = sub; c0 < p.KC; c0 += 4*SUBS) {               // loads batched ahead of the ordered tensor-memory stores
                        float4 buf[4];
                        #pragma unroll
                        for (int j = 0; j < 4; j++) if (c0 + j*SUBS < p.KC) buf[j] = __ldg(src + (size_t) (c0 + j*SUBS)*GT_TILE_ATOMS);
                        #pragma unroll
                        for (int j = 0; j < 4; j++) if (c0 + j*SUBS < p.KC) {
                            const float4 v = buf[j];
                            float4 h, l;
                            splitTf32(v, h, l);
                            tmemStore4(dst + 4*(c0 + j*SUBS), h);
                            tmemStore4(dst + p.Kp + 4*(c0 + j*SUBS), l);
                        }
                    }
                    const float2* eySrc = p.colY + tile*GT_TILE_ATOMS + atomInTile;
                    float2* eyDst = Eys + t*GT_TILE_ATOMS + atomInTile;
                    #pragma unroll 4
                    for (int m = sub; m < p.Ky; m += SUBS) eyDst[m*(MT*GT_TILE_ATOMS)] = __ldg(eySrc + (size_t) m*p.Npad);
                    oD[t] = 0.f; oX[t] = 0.f; oY[t] = 0.f; oZ[t] = 0.f;
                }
                if (warp == GT_EPI_WARP0) GT_STAMP(gs + 3);
                tmemWaitStore();
                tcgen05FenceBefore();
                __syncwarp();
                if (lane == 0) mbarArrive(aFull);
                if (warp == GT_EPI_WARP0) GT_STAMP(gs + 4);
                namedBarrier(1, EPI_THREADS);                     // Ey columns complete
                if (warp == GT_EPI_WARP0) GT_STAMP(gs + 5);
            }
            if (unit + 1 < u1) splitUnit(unit + 1);
            // row data / Ex of the NEXT unit are fetched while this one is processed (no exposed global latency)
            // (the next unit: next column tile of this atom group, or the first one of the next group; the last unit repeats itself)
            int groupN = group, colN = colTile;
            if (unit + 1 < u1 && ++colN == p.numColTiles) { colN = 0; groupN++; }
            const int g8N = colN*SUBS + sub;
            const int4 giN = __ldg(p.groupInfo + g8N);
            float4 rdN = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane < 8) rdN = __ldg(p.rowData + p.signedLo + 8*g8N + lane);
            #pragma unroll
            for (int t = 0; t < MT; t++, seq++) {
                const uint32_t d = seq & 1;
                mbarWait(&dFull[d], (seq >> 1) & 1);
                if (warp == GT_EPI_WARP0 && seq == 18) GT_STAMP(26);
                if (warp == GT_EPI_WARP0 && seq == 19) GT_STAMP(19);
                if (warp == GT_EPI_WARP0 && seq == 20) GT_STAMP(7);
                tcgen05FenceAfter();
                float v[32];
                tmemLoad32(tmem + d*128 + laneBase + 32*sub, v);
                tcgen05FenceBefore();
                __syncwarp();
                if (lane == 0) mbarArrive(&dEmpty[d]);            // values are in registers: the slot can be refilled
                if (warp == GT_EPI_WARP0 && seq == 18) GT_STAMP(27);
                const char* eyCol = reinterpret_cast<const char*>(Eys + t*GT_TILE_ATOMS + atomInTile);
                const float4* rd = rdS + buf*8;
                #pragma unroll
                for (int i = 0; i < 8; i++) {
                    // padding rows beyond signedHi have zero coefficients (U = V = 0) and zero row data
                    const float4 r = rd[i];                       // (nx, ny, byte offset of Ey(|ny|), -)
                    const float exx = (i < cc) ? ex0[t].x : ex1[t].x, exy = (i < cc) ? ex0[t].y : ex1[t].y;
                    float2 ey = *reinterpret_cast<const float2*>(eyCol + __float_as_uint(r.z));
                    ey.y = __uint_as_float(__float_as_uint(ey.y) ^ (__float_as_uint(r.y) & 0x80000000u));     // Ey(-m) = conj Ey(m)
                    const float tr = exx*ey.x - exy*ey.y;
                    const float ti = exx*ey.y + exy*ey.x;
                    const float ur = v[4*i], ui = v[4*i+1], vr = v[4*i+2], vi = v[4*i+3];
                    oD[t] = fmaf(tr, ur, oD[t]);  oD[t] = fmaf(-ti, ui, oD[t]);
                    const float im = tr*ui + ti*ur;
                    oX[t] = fmaf(r.x, im, oX[t]);
                    oY[t] = fmaf(r.y, im, oY[t]);
                    oZ[t] = fmaf(tr, vi, oZ[t]);  oZ[t] = fmaf(ti, vr, oZ[t]);
                }
                if (warp == GT_EPI_WARP0 && (seq == 18 || seq == 19)) GT_STAMP(seq == 18 ? 14 : 15);
                if (t == 0) {
                    #pragma unroll
                    for (int tt = 0; tt < MT; tt++) {
                        const float2* exCol = p.colX + (groupN*MT + tt)*GT_TILE_ATOMS + atomInTile;
                        ex0n[tt] = __ldg(exCol + giN.x);
                        ex1n[tt] = __ldg(exCol + giN.y);
                    }
                }
            }
            buf ^= 1;
            if (lane < 8) rdS[buf*8 + lane] = rdN;
            __syncwarp();
            cc = giN.z;
            #pragma unroll
            for (int t = 0; t < MT; t++) { ex0[t] = ex0n[t]; ex1[t] = ex1n[t]; }
        }
        if (warp == GT_EPI_WARP0) GT_STAMP(16);
        if (curGroup >= 0) flush();
        if (warp == GT_EPI_WARP0) GT_STAMP(17);
    }
    tcgen05FenceBefore();
    __syncthreads();
    if (warp == 0) GT_STAMP(18);
    if (warp == 0) tmemFree<512>(tmem);
}


// ------------------------------------------------------------------------------------------------
// structure factors on the tensor cores
// ------------------------------------------------------------------------------------------------
// P[(comp, row)][(l, c|s)] = sum_atoms A[(comp,row)][atom] * Z[(l,c|s)][atom] with the row operand
// A = (xr*yc, xr*ys, xi*yc, xi*ys), x = q Ex(nx), y = Ey(|ny|): a [128 x atoms] x [atoms x 64] GEMM per tile of 32
// unsigned rows, K = atoms. A and Z are formed (products, TF32 hi/lo split) by CUDA-core warps from the per-atom phase
// rows that arrive by bulk TMA, written to shared memory in the K-major core-matrix layout, multiplied by
// tcgen05.mma kind::tf32 (three products, small ones first) into tensor memory, and -- because the tensor core
// truncates on accumulation -- taken out after every 32 atoms and summed in FP32 registers with round-to-nearest.
// One CTA = 2 row tiles (64 rows) x one split of the atoms.
constexpr int ST_ATOMS = 32;                 // atoms per stage = 4 MMA k-steps
constexpr int ST_FORM_WARPS = 8, ST_EPI_WARPS = 8;
constexpr int ST_FORM_WARP0 = 2, ST_EPI_WARP0 = ST_FORM_WARP0 + ST_FORM_WARPS;
constexpr int ST_MMA_WARP1 = ST_EPI_WARP0 + ST_EPI_WARPS;    // second MMA issuer (the first is warp 1)
constexpr int ST_THREADS = (ST_MMA_WARP1 + 1)*32;
constexpr uint32_t ST_A_PLANE = (ST_ATOMS/4)*128*16;      // bytes of one hi or lo plane of one row tile

struct STParams {
    const float2* rowS; float* part;
    int rowPitch, Kx, Ky, Kz, zOff, kzPad;
    int rowLo, rowHi, numRows;
    int atomsPerSplit, Npad;
    uint32_t rowStageBytes, offA, offB, offBar;
};

// NN = GEMM N (2*Kz padded to 64 or 128), TT = row tiles of 32 rows per CTA; TT*NN = 128 accumulator columns per slot
template <int NN, int TT>
__global__ void __launch_bounds__(ST_THREADS, 1) structureFactorTensorKernel(STParams p) {
    constexpr uint32_t B_PLANE = (ST_ATOMS/4)*NN*16;
    extern __shared__ __align__(128) unsigned char smem[];
    // [2 row stages][2 operand buffers: A (tiles x hi|lo planes)][2 operand buffers: B (hi|lo)][barriers]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.offBar);
    uint64_t* rowsFull = bars;           // [2] TMA -> formers
    uint64_t* rowsEmpty = bars + 2;      // [2] formers -> TMA
    uint64_t* abFull = bars + 4;         // [2] formers -> MMA
    uint64_t* abEmpty = bars + 6;        // [2] MMA -> formers
    uint64_t* dFull = bars + 8;          // [2] MMA -> epilogue
    uint64_t* dEmpty = bars + 10;        // [2] epilogue -> MMA
    uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(bars + 12);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rowBase = p.rowLo + blockIdx.x*(32*TT);
    const int atomBegin = blockIdx.y*p.atomsPerSplit;
    const int atomEnd = min(atomBegin + p.atomsPerSplit, p.Npad);
    const int numStages = (atomEnd - atomBegin)/ST_ATOMS;

    if (tid == 0) {
        for (int i = 0; i < 2; i++) {
            mbarInit(&rowsFull[i], 1); mbarInit(&rowsEmpty[i], ST_FORM_WARPS);
            mbarInit(&abFull[i], ST_FORM_WARPS); mbarInit(&abEmpty[i], 1);
            mbarInit(&dFull[i], 1); mbarInit(&dEmpty[i], ST_EPI_WARPS);
        }
        mbarFenceInit();
    }
    if (warp == 0) tmemAlloc<256>(tmemSlot);
    tcgen05FenceBefore();
    __syncthreads();
    tcgen05FenceAfter();
    const uint32_t tmem = *tmemSlot;

    if (warp == 0) {
        // ---------------- producer: per-atom phase rows, 32 atoms per stage ----------------
        if (lane == 0) {
            for (int st = 0; st < numStages; st++) {
                const int b = st & 1;
                if (st >= 2) mbarWait(&rowsEmpty[b], ((st >> 1) & 1) ^ 1);
                mbarExpectTx(&rowsFull[b], p.rowStageBytes);
                bulkLoad(smem + (size_t) b*p.rowStageBytes, p.rowS + (size_t) (atomBegin + st*ST_ATOMS)*p.rowPitch, p.rowStageBytes, &rowsFull[b]);
            }
        }
        __syncwarp();
    }
    else if (warp == 1 || warp == ST_MMA_WARP1) {
        // ---------------- MMA issuers (warp-uniform control flow, one elected lane issues) ----------------
        // two warps, one per operand buffer / accumulator slot: the waits and commits of one overlap the issue of the other
        constexpr uint32_t idesc = ummaIdescTf32(128, NN);
        for (int st = (warp == 1 ? 0 : 1); st < numStages; st += 2) {
            const int b = st & 1;
            const uint32_t ph = (st >> 1) & 1;
            mbarWait(&abFull[b], ph);
            mbarWait(&dEmpty[b], ph ^ 1);
            tcgen05FenceAfter();
            const uint32_t aBase = smemU32(smem) + p.offA + (uint32_t) b*(TT*2*ST_A_PLANE);
            const uint32_t bBase = smemU32(smem) + p.offB + (uint32_t) b*(2*B_PLANE);
            const uint64_t bHi = ummaSmemDesc(bBase, NN*16, 128), bLo = ummaSmemDesc(bBase + B_PLANE, NN*16, 128);
            if (electOne()) {
                #pragma unroll
                for (int t = 0; t < TT; t++) {
                    const uint32_t tD = tmem + (uint32_t) b*(TT*NN) + t*NN;
                    const uint64_t aHi = ummaSmemDesc(aBase + t*2*ST_A_PLANE, 128*16, 128), aLo = ummaSmemDesc(aBase + t*2*ST_A_PLANE + ST_A_PLANE, 128*16, 128);
                    // one k-step = 8 atoms = two 16-byte chunk columns: 2*128*16 bytes of A, 2*NN*16 bytes of Z;
                    // small products first
                    #pragma unroll
                    for (int k8 = 0; k8 < ST_ATOMS/8; k8++) ummaTf32SS(tD, aLo + (uint64_t) (k8*(2*128*16 >> 4)), bHi + (uint64_t) (k8*(2*NN*16 >> 4)), idesc, k8 > 0);
                    #pragma unroll
                    for (int k8 = 0; k8 < ST_ATOMS/8; k8++) ummaTf32SS(tD, aHi + (uint64_t) (k8*(2*128*16 >> 4)), bLo + (uint64_t) (k8*(2*NN*16 >> 4)), idesc, 1);
                    #pragma unroll
                    for (int k8 = 0; k8 < ST_ATOMS/8; k8++) ummaTf32SS(tD, aHi + (uint64_t) (k8*(2*128*16 >> 4)), bHi + (uint64_t) (k8*(2*NN*16 >> 4)), idesc, 1);
                }
                ummaCommit(&abEmpty[b]);
                ummaCommit(&dFull[b]);
            }
            __syncwarp();
        }
    }
    else if (warp >= ST_FORM_WARP0 && warp < ST_EPI_WARP0) {
        // ---------------- formers: row operand A (products) and column operand Z, TF32 hi/lo planes ----------------
        const int ft = tid - ST_FORM_WARP0*32;                     // 0..255
        constexpr int A_PER = 32*TT*(ST_ATOMS/4)/(32*ST_FORM_WARPS);    // (row, quad of 4 atoms) items per thread
        constexpr int Z_PER = (NN/2)*(ST_ATOMS/4)/(32*ST_FORM_WARPS);   // (|nz|, quad of 4 atoms) items per thread
        int aNx[A_PER], aM[A_PER], aQd[A_PER]; uint32_t aDst[A_PER]; bool aValid[A_PER];
        #pragma unroll
        for (int k = 0; k < A_PER; k++) {
            const int item = ft + k*32*ST_FORM_WARPS, r = item % (32*TT), row = rowBase + r;
            aQd[k] = item/(32*TT);
            aValid[k] = row < p.rowHi;
            aNx[k] = aValid[k] ? row/p.Ky : 0;
            aM[k] = aValid[k] ? row - aNx[k]*p.Ky : 0;
            aDst[k] = (uint32_t) (r >> 5)*(2*ST_A_PLANE) + (uint32_t) aQd[k]*(128*16) + (uint32_t) (r & 31)*16;
        }
        for (int st = 0; st < numStages; st++) {
            const int b = st & 1;
            const uint32_t ph = (st >> 1) & 1;
            mbarWait(&rowsFull[b], ph);
            mbarWait(&abEmpty[b], ph ^ 1);                        // the MMAs that read this operand buffer are complete
            const float2* rows = reinterpret_cast<const float2*>(smem + (size_t) b*p.rowStageBytes);
            unsigned char* aBuf = smem + p.offA + (size_t) b*(TT*2*ST_A_PLANE);
            unsigned char* bBuf = smem + p.offB + (size_t) b*(2*B_PLANE);
            #pragma unroll
            for (int k = 0; k < A_PER; k++) {
                float4 pc[4];                                      // per comp: 4 atoms
                #pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float2* ar = rows + (size_t) (4*aQd[k] + j)*p.rowPitch;
                    const float2 x = ar[aNx[k]], y = ar[p.Kx + aM[k]];
                    (&pc[0].x)[j] = x.x*y.x; (&pc[1].x)[j] = x.x*y.y; (&pc[2].x)[j] = x.y*y.x; (&pc[3].x)[j] = x.y*y.y;
                }
                unsigned char* dst0 = aBuf + aDst[k];
                #pragma unroll
                for (int c = 0; c < 4; c++) {
                    float4 hi, lo;
                    if (aValid[k]) splitTf32(pc[c], hi, lo);
                    else { hi = make_float4(0.f, 0.f, 0.f, 0.f); lo = hi; }
                    *reinterpret_cast<float4*>(dst0 + c*(32*16)) = hi;
                    *reinterpret_cast<float4*>(dst0 + c*(32*16) + ST_A_PLANE) = lo;
                }
            }
            #pragma unroll
            for (int k = 0; k < Z_PER; k++) {
                const int item = ft + k*32*ST_FORM_WARPS, zl = item % (NN/2), zq = item/(NN/2);
                float4 zc, zs;
                #pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float2 z = (zl < p.Kz) ? rows[(size_t) (4*zq + j)*p.rowPitch + p.zOff + zl] : make_float2(0.f, 0.f);
                    (&zc.x)[j] = z.x; (&zs.x)[j] = z.y;
                }
                float4 hi, lo;
                unsigned char* dst = bBuf + (uint32_t) zq*(NN*16) + (uint32_t) (2*zl)*16;
                splitTf32(zc, hi, lo);
                *reinterpret_cast<float4*>(dst) = hi; *reinterpret_cast<float4*>(dst + B_PLANE) = lo;
                splitTf32(zs, hi, lo);
                *reinterpret_cast<float4*>(dst + 16) = hi; *reinterpret_cast<float4*>(dst + 16 + B_PLANE) = lo;
            }
            fenceProxyAsync();                                    // operands visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) { mbarArrive(&abFull[b]); mbarArrive(&rowsEmpty[b]); }
        }
    }
    else {
        // ---------------- epilogue: FP32 running sums; thread = (tensor-memory lane = (comp, row), block of 64 columns) ----------------
        const int e = warp - ST_EPI_WARP0;
        const int q4 = warp & 3, cb = e >> 2;                     // lane quarter = comp; column block of the 128-column slot
        const int t = cb/(NN/64), colOff = (cb % (NN/64))*64;     // row tile, first column within its NN columns
        const uint32_t laneBase = (uint32_t) (q4*32) << 16;
        float acc[64];
        #pragma unroll
        for (int i = 0; i < 64; i++) acc[i] = 0.f;
        for (int st = 0; st < numStages; st++) {
            const int b = st & 1;
            mbarWait(&dFull[b], (st >> 1) & 1);
            tcgen05FenceAfter();
            const uint32_t tD = tmem + (uint32_t) b*(TT*NN) + cb*64 + laneBase;
            #pragma unroll
            for (int c = 0; c < 4; c++) {
                float v[16];
                tmemLoad16(tD + 16*c, v);
                if (c == 3) {                                     // all columns are in registers: the slot can be refilled
                    tcgen05FenceBefore();
                    __syncwarp();
                    if (lane == 0) mbarArrive(&dEmpty[b]);
                }
                #pragma unroll
                for (int i = 0; i < 16; i++) acc[16*c + i] += v[i];
            }
        }
        const int row = rowBase + t*32 + lane;
        if (row < p.rowHi) {
            float* out = p.part + (((size_t) blockIdx.y*p.numRows + row)*p.kzPad + colOff/2)*8 + q4*2;
            #pragma unroll
            for (int l = 0; l < 32; l++)
                if (colOff/2 + l < p.Kz) *reinterpret_cast<float2*>(out + (size_t) l*8) = make_float2(acc[2*l], acc[2*l + 1]);
        }
    }
    tcgen05FenceBefore();
    __syncthreads();
    if (warp == 0) tmemFree<256>(tmem);
}


// ------------------------------------------------------------------------------------------------
// structure factors on the INTEGER tensor cores (exact sums): the kernel of both the energy and the forces-only call
// ------------------------------------------------------------------------------------------------
// Same GEMM as above, P[(comp,row)][(l,c|s)] = sum_atoms A * Z, with both operands in fixed point: A' = A sA, Z' = Z sZ,
// |A'|, |Z'| < 2^22 (sA from the largest |q| of this evaluation), written as SIGNED base-256 digits
//     v = D2 2^16 + D1 2^8 + D0 (+ D-1 2^-8),   D1, D0, D-1 in [-128, 127], |D2| <= 64,
// and the digit planes are multiplied by tcgen05.mma kind::i8 (s8 x s8 -> s32). Integer products and sums are exact: no
// truncation bias, no need to take the accumulators out every stage, a result independent of the summation order.
// Products of equal weight share an accumulator group (NN columns each, side by side in tensor memory):
//     2^32: D2 D2     2^24: D2 D1, D1 D2     2^16: D2 D0, D1 D1, D0 D2     2^8: D2 D-1, D1 D0, D0 D1, D-1 D2
// (weights below 2^8 are dropped: at most 3 x 2^14 per atom against products of ~2^40, zero-mean). With the Z digit planes
// stacked along N ([D2 | D1 | D0 | D-1] rows), ONE MMA per A plane covers all its groups: A digit i times Z digits 0..3-i
// lands in groups i..3, i.e. N = (4-i) NN columns from column i NN on: 3 (ND = 3) or 4 (ND = 4) MMAs of K = 32 atoms per row
// tile and stage, where the TF32 kernel needs 12 of K = 8. ND = 3 (23-bit operands, ~4x the rounding noise of FP32 operands)
// serves the forces-only call; the energy call adds the fourth digit (31-bit operands: quieter than FP32).
// The digits come out of the float -> int conversion: t = fma(x, y, 2^23 + 0x408080) has the mantissa v + 0x408080, whose
// bytes are D0 + 128, D1 + 128, D2 + 64; the residual fma(x, y, -(t - magic)) is exact and gives D-1 the same way; PRMT gathers
// the bytes of four atoms into one word per plane.
// Worst-case accumulator: 3 x 2^14 per atom in the 2^8 group -> at most 40,960 atoms per CTA (planStructureTensor).
constexpr int SI_ATOMS = 32;                 // atoms per stage = one MMA k-step (32 int8)
constexpr int SI_FORM_WARPS = 16, SI_FORM_WARP0 = 2;
constexpr int SI_THREADS = (SI_FORM_WARP0 + SI_FORM_WARPS)*32;
constexpr uint32_t SI_A_PLANE = 2*128*16;    // bytes of one digit plane of one row tile: two 16-atom chunks x 128 lanes x 16 B
constexpr float SI_MAGIC = 12615808.0f;      // 2^23 + 0x408080
constexpr float SI_MAGIC_LO = 8388736.0f;    // 2^23 + 128
constexpr float SI_RANGE = 4160000.0f;       // |v| stays below 2^22 - 2^15 - 1407 (mantissa within [0, 2^23))
constexpr int SI_MAX_ATOMS = 40960;

struct SIParams {
    const float2* rowS; float* part; const unsigned long long* qmaxSlot;
    int rowPitch, Kx, Ky, Kz, zOff, kzPad;
    int rowLo, rowHi, numRows;
    int atomsPerSplit, Npad;
    int rowStages, opStages;
    uint32_t rowStageBytes, rowStagePad, offOp, opBytes, offBar;
};

// four biased mantissas -> the word of byte j (bytes of the word = atoms), as signed digits; mask = 0 for padding
__device__ __forceinline__ uint32_t siPlane(uint32_t u0, uint32_t u1, uint32_t u2, uint32_t u3, int j, uint32_t mask) {
    const uint32_t sel = (uint32_t) j | ((uint32_t) (4 + j) << 4);
    const uint32_t w = __byte_perm(__byte_perm(u0, u1, sel), __byte_perm(u2, u3, sel), 0x5410);
    return ((j == 2 ? w + 0x40404040u : w) ^ 0x80808080u) & mask;               // b - 64 (b in [0,127]) / b - 128
}
// t = rint(xy) + magic; returns the biased mantissa word of the residual digit
__device__ __forceinline__ uint32_t siResidual(float x, float y, float t) {
    const float r = fmaf(x, y, -(t - SI_MAGIC));                                 // exact, |r| <= 1/2
    return __float_as_uint(fmaf(fminf(r, 0.4975f), 256.0f, SI_MAGIC_LO));
}

template <int NN, int TT, int ND>
__global__ void __launch_bounds__(SI_THREADS, 1) structureFactorI8Kernel(SIParams p) {
    static_assert(NN*TT == 128, "four weight groups of TT*NN columns fill the 512 tensor-memory columns");
    constexpr uint32_t B_CHUNK = ND*NN*16;                // bytes of one 16-atom chunk column of the stacked Z operand
    constexpr uint32_t A_BYTES = TT*ND*SI_A_PLANE;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.offBar);
    uint64_t* rowsFull = bars;            // [4] TMA -> formers
    uint64_t* rowsEmpty = bars + 4;       // [4] formers -> TMA
    uint64_t* abFull = bars + 8;          // [4] formers -> MMA
    uint64_t* abEmpty = bars + 12;        // [4] MMA -> formers
    uint64_t* dFull = bars + 16;          // MMA -> epilogue (once)
    uint64_t* tmemReady = bars + 17;      // accumulators zeroed -> MMA
    uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(bars + 18);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rowBase = p.rowLo + blockIdx.x*(32*TT);
    const int atomBegin = blockIdx.y*p.atomsPerSplit;
    const int atomEnd = min(atomBegin + p.atomsPerSplit, p.Npad);
    const int numStages = (atomEnd - atomBegin)/SI_ATOMS;
    const int RS = p.rowStages, OB = p.opStages;

    if (tid == 0) {
        for (int i = 0; i < 4; i++) {
            mbarInit(&rowsFull[i], 1); mbarInit(&rowsEmpty[i], SI_FORM_WARPS);
            mbarInit(&abFull[i], SI_FORM_WARPS); mbarInit(&abEmpty[i], 1);
        }
        mbarInit(dFull, 1); mbarInit(tmemReady, SI_FORM_WARPS);
        mbarFenceInit();
    }
    if (warp == 0) tmemAlloc<512>(tmemSlot);
    tcgen05FenceBefore();
    __syncthreads();
    tcgen05FenceAfter();
    const uint32_t tmem = *tmemSlot;

    if (warp == 0) {
        // ---------------- producer: per-atom phase rows, 32 atoms per stage ----------------
        if (lane == 0) {
            int rs = 0; uint32_t ph = 0;
            for (int st = 0; st < numStages; st++) {
                if (st >= RS) mbarWait(&rowsEmpty[rs], ph ^ 1);
                mbarExpectTx(&rowsFull[rs], p.rowStageBytes);
                bulkLoad(smem + (size_t) rs*p.rowStagePad, p.rowS + (size_t) (atomBegin + st*SI_ATOMS)*p.rowPitch, p.rowStageBytes, &rowsFull[rs]);
                if (++rs == RS) { rs = 0; ph ^= 1; }
            }
        }
        __syncwarp();
    }
    else if (warp == 1) {
        // ---------------- MMA issuer: every MMA accumulates (the accumulators start at zero) ----------------
        mbarWait(tmemReady, 0);
        tcgen05FenceAfter();
        int b = 0; uint32_t ph = 0;
        for (int st = 0; st < numStages; st++) {
            mbarWait(&abFull[b], ph);
            tcgen05FenceAfter();
            const uint32_t opBase = smemU32(smem) + p.offOp + (uint32_t) b*p.opBytes;
            const uint32_t zBase = opBase + A_BYTES;
            if (electOne()) {
                #pragma unroll
                for (int t = 0; t < TT; t++) {
                    #pragma unroll
                    for (int i = 0; i < ND; i++) {                                 // A digit i (0 = D2) x Z digits 0 .. 3-i -> groups i ..
                        const int nk = (4 - i < ND) ? 4 - i : ND;
                        const uint64_t a = ummaSmemDesc(opBase + (uint32_t) (t*ND + i)*SI_A_PLANE, 128*16, 128);
                        #pragma unroll
                        for (int r0 = 0; r0 < nk*NN; r0 += 256) {
                            const int n = (nk*NN - r0 < 256) ? nk*NN - r0 : 256;
                            ummaI8SS(tmem + (uint32_t) (t*4*NN + i*NN + r0), a, ummaSmemDesc(zBase + (uint32_t) r0*16, B_CHUNK, 128),
                                     ummaIdescS8(128, n), 1);
                        }
                    }
                }
                ummaCommit(&abEmpty[b]);
                if (st == numStages - 1) ummaCommit(dFull);
            }
            __syncwarp();
            if (++b == OB) { b = 0; ph ^= 1; }
        }
    }
    else {
        // ---------------- formers (16 warps), then the one-off epilogue ----------------
        const int fw = warp - SI_FORM_WARP0, ft = tid - SI_FORM_WARP0*32;        // 0..15, 0..511
        const int q4 = warp & 3;                                                 // the lane quarter this warp may access
        {   // zero this warp's share of the accumulators: lane quarter q4, 128 of the 512 columns
            const uint32_t base = tmem + ((uint32_t) (q4*32) << 16) + (uint32_t) (fw >> 2)*128;
            #pragma unroll
            for (int c = 0; c < 8; c++) tmemStoreZero16(base + 16*c);
            tmemWaitStore();
            tcgen05FenceBefore();
            __syncwarp();
            if (lane == 0) mbarArrive(tmemReady);
        }
        const unsigned int qbits = (unsigned int) __ldg(p.qmaxSlot);
        const float sA = SI_RANGE/fmaxf(__uint_as_float(qbits), 1e-30f), sZ = SI_RANGE;
        // A item of this thread: (row of the CTA, quad of 4 atoms); lanes = 4 quads of a chunk x 8 rows -> the 4-byte stores
        // of a warp cover 8 x 16 contiguous bytes per plane
        constexpr int A_ITEMS = 32*TT*8;
        const bool hasA = ft < A_ITEMS;
        const int aR = (ft >> 2) % (32*TT), aQ = ((ft >> 2)/(32*TT))*4 + (ft & 3);
        const int aRow = rowBase + aR;
        const bool aValid = hasA && aRow < p.rowHi;
        const uint32_t aMask = aValid ? 0xFFFFFFFFu : 0u;
        const int aNx = aValid ? aRow/p.Ky : 0, aM = aValid ? aRow - aNx*p.Ky : 0;
        const uint32_t aDst = (uint32_t) (aR >> 5)*(ND*SI_A_PLANE) + (uint32_t) (aQ >> 2)*(128*16) + (uint32_t) (aR & 31)*16 + (uint32_t) (aQ & 3)*4;
        const uint32_t aOffX = (uint32_t) ((4*aQ)*p.rowPitch + aNx)*8, aOffY = (uint32_t) ((4*aQ)*p.rowPitch + p.Kx + aM)*8;
        // Z item: (|nz| slot, quad); both the cos and the sin row
        constexpr int Z_ITEMS = (NN/2)*8;
        const bool hasZ = ft < Z_ITEMS;
        const int zL = (ft >> 2) % (NN/2), zQ = ((ft >> 2)/(NN/2))*4 + (ft & 3);
        const bool zValid = hasZ && zL < p.Kz;
        const uint32_t zMask = zValid ? 0xFFFFFFFFu : 0u;
        const uint32_t zDst = (uint32_t) (zQ >> 2)*B_CHUNK + (uint32_t) (2*zL)*16 + (uint32_t) (zQ & 3)*4;
        const uint32_t zOff = (uint32_t) ((4*zQ)*p.rowPitch + p.zOff + (zValid ? zL : 0))*8;
        const uint32_t pitchBytes = (uint32_t) p.rowPitch*8;

        int b = 0, rs = 0; uint32_t phB = 0, phR = 0;
        for (int st = 0; st < numStages; st++) {
            mbarWait(&rowsFull[rs], phR);
            mbarWait(&abEmpty[b], phB ^ 1);                       // the MMAs that read this operand buffer are complete
            const unsigned char* rows = smem + (size_t) rs*p.rowStagePad;
            unsigned char* aBuf = smem + p.offOp + (size_t) b*p.opBytes;
            unsigned char* zBuf = aBuf + A_BYTES;
            if (hasA) {
                uint32_t u[4][4], ul[4][4];                        // [comp][atom of the quad]: digits 2..0, residual digit
                #pragma unroll
                for (int j = 0; j < 4; j++) {
                    float2 x = *reinterpret_cast<const float2*>(rows + aOffX + j*pitchBytes);
                    const float2 y = *reinterpret_cast<const float2*>(rows + aOffY + j*pitchBytes);
                    x.x *= sA; x.y *= sA;
                    const float t0 = fmaf(x.x, y.x, SI_MAGIC), t1 = fmaf(x.x, y.y, SI_MAGIC), t2 = fmaf(x.y, y.x, SI_MAGIC), t3 = fmaf(x.y, y.y, SI_MAGIC);
                    u[0][j] = __float_as_uint(t0); u[1][j] = __float_as_uint(t1); u[2][j] = __float_as_uint(t2); u[3][j] = __float_as_uint(t3);
                    if (ND == 4) {
                        ul[0][j] = siResidual(x.x, y.x, t0); ul[1][j] = siResidual(x.x, y.y, t1);
                        ul[2][j] = siResidual(x.y, y.x, t2); ul[3][j] = siResidual(x.y, y.y, t3);
                    }
                }
                #pragma unroll
                for (int c = 0; c < 4; c++) {
                    unsigned char* dst = aBuf + aDst + c*(32*16);
                    #pragma unroll
                    for (int pl = 0; pl < 3; pl++)                 // plane 0 of the buffer = D2 ... plane 2 = D0
                        *reinterpret_cast<uint32_t*>(dst + pl*SI_A_PLANE) = siPlane(u[c][0], u[c][1], u[c][2], u[c][3], 2 - pl, aMask);
                    if (ND == 4) *reinterpret_cast<uint32_t*>(dst + 3*SI_A_PLANE) = siPlane(ul[c][0], ul[c][1], ul[c][2], ul[c][3], 0, aMask);
                }
            }
            if (hasZ) {
                uint32_t uc[4], us[4], lc[4], ls[4];
                #pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float2 z = *reinterpret_cast<const float2*>(rows + zOff + j*pitchBytes);
                    const float tc = fmaf(z.x, sZ, SI_MAGIC), ts = fmaf(z.y, sZ, SI_MAGIC);
                    uc[j] = __float_as_uint(tc); us[j] = __float_as_uint(ts);
                    if (ND == 4) { lc[j] = siResidual(z.x, sZ, tc); ls[j] = siResidual(z.y, sZ, ts); }
                }
                #pragma unroll
                for (int pl = 0; pl < 3; pl++) {
                    unsigned char* dst = zBuf + zDst + pl*(NN*16);
                    *reinterpret_cast<uint32_t*>(dst) = siPlane(uc[0], uc[1], uc[2], uc[3], 2 - pl, zMask);
                    *reinterpret_cast<uint32_t*>(dst + 16) = siPlane(us[0], us[1], us[2], us[3], 2 - pl, zMask);
                }
                if (ND == 4) {
                    unsigned char* dst = zBuf + zDst + 3*(NN*16);
                    *reinterpret_cast<uint32_t*>(dst) = siPlane(lc[0], lc[1], lc[2], lc[3], 0, zMask);
                    *reinterpret_cast<uint32_t*>(dst + 16) = siPlane(ls[0], ls[1], ls[2], ls[3], 0, zMask);
                }
            }
            fenceProxyAsync();                                    // operands visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) { mbarArrive(&abFull[b]); mbarArrive(&rowsEmpty[rs]); }
            if (++b == OB) { b = 0; phB ^= 1; }
            if (++rs == RS) { rs = 0; phR ^= 1; }
        }
        // ---------------- epilogue: exact 64-bit combination of the four groups, scaled back, one FP32 rounding ----------------
        {
            if (numStages > 0) mbarWait(dFull, 0);                 // (an empty split still writes its zeros)
            tcgen05FenceAfter();
            const int blk = fw >> 2;                               // 32 of the TT*NN columns of every group
            const int t = blk/(NN/32), colOff = (blk % (NN/32))*32;
            const uint32_t tD = tmem + ((uint32_t) (q4*32) << 16) + (uint32_t) t*(4*NN) + colOff;
            const double inv = 1.0/((double) sA*(double) sZ);
            const int row = rowBase + t*32 + lane;
            float* out = p.part + (((size_t) blockIdx.y*p.numRows + row)*p.kzPad + colOff/2)*8 + q4*2;
            #pragma unroll
            for (int c = 0; c < 2; c++) {
                int g3[16], g2[16], g1[16], g0[16];
                tmemLoad16i(tD + 16*c, g3); tmemLoad16i(tD + NN + 16*c, g2); tmemLoad16i(tD + 2*NN + 16*c, g1); tmemLoad16i(tD + 3*NN + 16*c, g0);
                #pragma unroll
                for (int i = 0; i < 8; i++) {
                    float v[2];
                    #pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int k = 2*i + h;
                        const long long sum = ((long long) g3[k] << 32) + ((long long) g2[k] << 24) + ((long long) g1[k] << 16) + ((long long) g0[k] << 8);
                        v[h] = (float) ((double) sum*inv);
                    }
                    const int l = colOff/2 + 8*c + i;
                    if (row < p.rowHi && l < p.Kz) *reinterpret_cast<float2*>(out + (size_t) (8*c + i)*8) = make_float2(v[0], v[1]);
                }
            }
        }
    }
    tcgen05FenceBefore();
    __syncthreads();
    if (warp == 0) tmemFree<512>(tmem);
}

} // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
void planKSpaceTensor(State& st) {
    KSpacePlan& ks = st.ks;
    ks.tensorGather = false;
    const char* mode = getenv("CFX_KSPACE_GATHER");             // "fp32" forces the CUDA-core kernel (A/B measurements)
    if (mode && !strcmp(mode, "fp32")) return;
    const int Kz = ks.K[2], Ky = ks.K[1];
    const int Kp = (2*Kz + 7)/8*8;
    if (Kp > 112 || Ky < 8) return;                              // (a group of 8 rows must not span three nx values)                                        // phase operand must fit 2 x 224 tensor-memory columns
    ks.tKp = Kp; ks.tKC = Kp/4;
    ks.tMT = (Kp <= 56) ? 2 : 1;
    ks.tNT = (Kp <= 56) ? 128 : 64;
    const size_t planeBytes = (size_t) ks.tKC*ks.tNT*16;         // one FP32 / TF32-hi / TF32-lo plane of a coefficient tile
    const size_t eyBytes = ((size_t) Ky*ks.tMT*GT_TILE_ATOMS*sizeof(float2) + 127) & ~(size_t) 127;
    const size_t tailBytes = 256 + (size_t) (ks.tNT/8)*16*sizeof(float4);   // barriers + per-epilogue-warp row-data slots (2 x 8 rows)
    const size_t cap = 227*1024 - 256;
    if (4*planeBytes + eyBytes + tailBytes > cap) return;         // operand double buffer (hi+lo) needs 4 planes
    const int rawStages = (int) std::min<size_t>(4, (cap - 4*planeBytes - eyBytes - tailBytes)/planeBytes);
    if (rawStages < 2) return;
    ks.tStages = rawStages;
    ks.tStageBytes = (uint32_t) planeBytes;
    ks.tOffEy = (uint32_t) ((4 + rawStages)*planeBytes);
    ks.tOffBar = (uint32_t) (ks.tOffEy + eyBytes);
    ks.tSmem = ks.tOffBar + tailBytes;
    const int rows = ks.tNT/4;
    const int signedHere = std::max(ks.signedHi - ks.signedLo, 1);
    ks.tColTiles = (signedHere + rows - 1)/rows;
    const size_t zFloats = (size_t) (st.Npad/GT_TILE_ATOMS)*ks.tKC*GT_TILE_ATOMS*4;
    const size_t cFloats = (size_t) ks.tColTiles*ks.tKC*ks.tNT*4;
    CFX_CUDA(cudaMalloc(&st.zSplit, zFloats*sizeof(float)));
    CFX_CUDA(cudaMemset(st.zSplit, 0, zFloats*sizeof(float)));
    CFX_CUDA(cudaMalloc(&st.coefT, cFloats*sizeof(float)));
    CFX_CUDA(cudaMemset(st.coefT, 0, cFloats*sizeof(float)));
    // per signed row: (nx, ny, byte offset of Ey(|ny|) in the shared columns, -), gRowInfo order, zero-padded; and per
    // group of 8 rows counted from signedLo: (Ex offset of the first row's nx, of the last real row's nx, rows with the first nx)
    {
        const int Kx = ks.K[0];
        std::vector<float4> rd;
        std::vector<int> rowNx;
        const int stride = ks.tMT*GT_TILE_ATOMS;
        for (int row = 0; row < Kx*Ky; row++) {
            const int nx = row/Ky, m = row % Ky;
            rd.push_back(make_float4((float) nx, (float) m, __int_as_float_host(m*stride*8), 0.f)); rowNx.push_back(nx);
            if (nx > 0 && m > 0) { rd.push_back(make_float4((float) nx, (float) -m, __int_as_float_host(m*stride*8), 0.f)); rowNx.push_back(nx); }
        }
        for (int k = 0; k < 64; k++) rd.push_back(make_float4(0.f, 0.f, 0.f, 0.f));
        std::vector<int4> gi((size_t) ks.tColTiles*(ks.tNT/32) + 1);
        for (size_t g = 0; g < gi.size(); g++) {
            const int r0 = std::min(ks.signedLo + 8*(int) g, std::max(ks.signedHi - 1, 0));
            const int r7 = std::min(ks.signedLo + 8*(int) g + 7, std::max(ks.signedHi - 1, 0));
            int c = 0;
            for (int r = ks.signedLo + 8*(int) g; r < ks.signedLo + 8*(int) g + 8; r++) if (r >= ks.signedHi || rowNx[std::min(r, (int) rowNx.size() - 1)] == rowNx[r0]) c++; else break;
            gi[g] = make_int4(rowNx[r0]*st.Npad, rowNx[r7]*st.Npad, c, 0);
        }
        CFX_CUDA(cudaMalloc(&st.gRowData, rd.size()*sizeof(float4)));
        CFX_CUDA(cudaMemcpy(st.gRowData, rd.data(), rd.size()*sizeof(float4), cudaMemcpyHostToDevice));
        CFX_CUDA(cudaMalloc(&st.gGroupInfo, gi.size()*sizeof(int4)));
        CFX_CUDA(cudaMemcpy(st.gGroupInfo, gi.data(), gi.size()*sizeof(int4), cudaMemcpyHostToDevice));
    }
    CFX_CUDA(cudaFuncSetAttribute(gatherTensorKernel<2, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
    CFX_CUDA(cudaFuncSetAttribute(gatherTensorKernel<1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
    if (getenv("CFX_GT_TRACE")) { CFX_CUDA(cudaMalloc(&st.gtTrace, 148*32*8)); CFX_CUDA(cudaMemset(st.gtTrace, 0, 148*32*8)); }
    ks.tensorGather = true;
}

void launchGatherTensor(State& st, long long* dForce, long long* dDedq, cudaStream_t s) {
    KSpacePlan& ks = st.ks;
    GTParams gp;
    gp.zSplit = st.zSplit; gp.coefT = st.coefT; gp.rowData = st.gRowData; gp.groupInfo = st.gGroupInfo; gp.colX = st.colX; gp.colY = st.colY; gp.qf = st.qf;
    gp.Ky = ks.K[1]; gp.Kp = ks.tKp; gp.KC = ks.tKC; gp.N = st.N; gp.Npad = st.Npad;
    gp.signedLo = ks.signedLo; gp.signedHi = ks.signedHi; gp.numColTiles = ks.tColTiles;
    gp.numAtomGroups = st.Npad/(GT_TILE_ATOMS*ks.tMT);
    gp.fx = (float) (2*M_PI/st.box.L[0]); gp.fy = (float) (2*M_PI/st.box.L[1]); gp.fz = (float) (2*M_PI/st.box.L[2]);
    gp.trace = nullptr;
    gp.trace = st.gtTrace;
    gp.rawStages = ks.tStages; gp.planeBytes = ks.tStageBytes; gp.offRaw = 4*ks.tStageBytes; gp.offEy = ks.tOffEy; gp.offBar = ks.tOffBar; gp.offRd = ks.tOffBar + 256;
    int numSM = 148;
    cudaDeviceGetAttribute(&numSM, cudaDevAttrMultiProcessorCount, st.device);
    const long long units = (long long) gp.numAtomGroups*gp.numColTiles;
    const int grid = (int) std::min<long long>(numSM, units);
    if (ks.tMT == 2) gatherTensorKernel<2, 128><<<grid, GT_THREADS, ks.tSmem, s>>>(gp, dForce, dDedq);
    else             gatherTensorKernel<1, 64><<<grid, GT_THREADS, ks.tSmem, s>>>(gp, dForce, dDedq);
    CFX_LAUNCH_CHECK(); st.launches++;
}

// ---- structure factors ----
bool structureTensorEligible(const State& st) {
    const char* mode = getenv("CFX_KSPACE_S");                  // "fp32" forces the CUDA-core kernel (A/B measurements)
    if (mode && !strcmp(mode, "fp32")) return false;
    return st.ks.K[2] <= 64;                                    // 2*Kz columns must fit one N <= 128 MMA
}

// shared-memory layout of the integer kernel for ND digit planes: [row stages][operand ring][barriers]; the deepest operand
// ring that fits beside three (else two) row stages
struct SiLayout { int rowStages = 0, opStages = 0; uint32_t opBytes = 0, offOp = 0, offBar = 0; size_t smem = 0; };
static bool siLayout(int NN, int TT, int ND, size_t rowStagePad, SiLayout& l) {
    const size_t opBytes = (size_t) TT*ND*SI_A_PLANE + 2*(size_t) (ND*NN)*16;
    const size_t cap = 227*1024 - 256 - 256;
    int rsFixed = 0, obFixed = 0;                                // experiments
    if (const char* e = getenv("CFX_SI_ROW_STAGES")) rsFixed = atoi(e);
    if (const char* e = getenv("CFX_SI_OP_STAGES")) obFixed = atoi(e);
    for (int rsN = 3; rsN >= 2; rsN--)
        for (int ob = 4; ob >= 2; ob--) {
            if ((rsFixed && rsN != rsFixed) || (obFixed && ob != obFixed)) continue;
            if (rsN*rowStagePad + ob*opBytes <= cap) {
                l.rowStages = rsN; l.opStages = ob; l.opBytes = (uint32_t) opBytes;
                l.offOp = (uint32_t) (rsN*rowStagePad); l.offBar = (uint32_t) (l.offOp + ob*opBytes);
                l.smem = l.offBar + 256;
                return true;
            }
        }
    return false;
}

void planStructureTensor(State& st) {
    KSpacePlan& ks = st.ks;
    SGeom& t = ks.sT;
    int numSM = 148;
    cudaDeviceGetAttribute(&numSM, cudaDevAttrMultiProcessorCount, st.device);
    const int NN = ks.K[2] <= 32 ? 64 : 128, TT = 128/NN;        // GEMM N; row tiles per CTA
    t.TN = NN/2; t.NC = 1; t.kzPad = NN/2;                      // identity |nz| slots
    t.rowPitch = ((ks.K[0] + ks.K[1] + 1) & ~1) + t.kzPad;
    const int rowsHere = std::max(ks.rowHi - ks.rowLo, 1);
    t.rowTiles = (rowsHere + 32*TT - 1)/(32*TT);
    // atom splits: the smallest count that fills >= 95 % of the SM slots of its last wave (one CTA per SM)
    const int maxSplits = std::max(1, st.Npad/(4*ST_ATOMS));
    const int minSplits = (st.Npad + SI_MAX_ATOMS - 1)/SI_MAX_ATOMS;   // int32 accumulators of the integer kernel
    int splits = minSplits;
    double bestUtil = 0.0;
    for (int sp = minSplits; sp <= std::max(minSplits, std::min(maxSplits, 4*numSM)); sp++) {
        const int ctas = t.rowTiles*sp;
        const double util = (double) ctas/((double) ((ctas + numSM - 1)/numSM)*numSM);
        if (util > bestUtil + 1e-9) { bestUtil = util; splits = sp; }
        if (util >= 0.95) { splits = sp; break; }
    }
    int aps = (st.Npad + splits - 1)/splits;
    aps = (aps + ST_ATOMS - 1)/ST_ATOMS*ST_ATOMS;
    t.splits = (st.Npad + aps - 1)/aps;
    t.atomsPerSplit = aps;
    t.threads = ST_THREADS;
    const size_t rowStage = (size_t) ST_ATOMS*t.rowPitch*sizeof(float2);
    const size_t rowStagePad = (rowStage + 127) & ~(size_t) 127;
    ks.tsRowStageBytes = (uint32_t) rowStage;
    ks.tsOffA = (uint32_t) (2*rowStagePad);
    ks.tsOffB = ks.tsOffA + 2*TT*2*ST_A_PLANE;
    ks.tsOffBar = ks.tsOffB + 2*2*(ST_ATOMS/4)*NN*16;
    t.smem = ks.tsOffBar + 128;
    if (t.smem > 227*1024 - 256) { ks.tensorS = false; return; }
    CFX_CUDA(cudaFuncSetAttribute(structureFactorTensorKernel<64, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
    CFX_CUDA(cudaFuncSetAttribute(structureFactorTensorKernel<128, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
    // integer tensor-core kernel (exact sums: serves the energy call too, with a fourth digit plane): same row tiles / atom
    // splits / partial-sum layout; sized for four digit planes, deepest operand ring that fits beside the row stages
    ks.i8S = false;
    const char* mode = getenv("CFX_KSPACE_S");                  // "tf32": keep the TF32 kernel (and FP32 for energies)
    if (!(mode && !strcmp(mode, "tf32")) && t.atomsPerSplit <= SI_MAX_ATOMS) {
        ks.siRowStagePad = (uint32_t) rowStagePad;
        SiLayout l3, l4;
        ks.i8S = siLayout(NN, TT, 3, rowStagePad, l3) && siLayout(NN, TT, 4, rowStagePad, l4);
        CFX_CUDA(cudaFuncSetAttribute(structureFactorI8Kernel<64, 2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
        CFX_CUDA(cudaFuncSetAttribute(structureFactorI8Kernel<128, 1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
        CFX_CUDA(cudaFuncSetAttribute(structureFactorI8Kernel<64, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
        CFX_CUDA(cudaFuncSetAttribute(structureFactorI8Kernel<128, 1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
    }
}

void launchStructureTensor(State& st, bool energy, cudaStream_t s) {
    KSpacePlan& ks = st.ks;
    const SGeom& t = ks.sT;
    STParams sp;
    sp.rowS = st.rowS; sp.part = st.sPart;
    sp.rowPitch = t.rowPitch; sp.Kx = ks.K[0]; sp.Ky = ks.K[1]; sp.Kz = ks.K[2]; sp.zOff = t.rowPitch - t.kzPad; sp.kzPad = t.kzPad;
    sp.rowLo = ks.rowLo; sp.rowHi = ks.rowHi; sp.numRows = ks.numRows;
    sp.atomsPerSplit = t.atomsPerSplit; sp.Npad = st.Npad;
    if (ks.i8S) {
        SIParams ip;
        ip.rowS = st.rowS; ip.part = st.sPart; ip.qmaxSlot = reinterpret_cast<const unsigned long long*>(st.energyFixed + CFX_SLOT_QMAX);
        ip.rowPitch = t.rowPitch; ip.Kx = ks.K[0]; ip.Ky = ks.K[1]; ip.Kz = ks.K[2]; ip.zOff = t.rowPitch - t.kzPad; ip.kzPad = t.kzPad;
        ip.rowLo = ks.rowLo; ip.rowHi = ks.rowHi; ip.numRows = ks.numRows;
        ip.atomsPerSplit = t.atomsPerSplit; ip.Npad = st.Npad;
        const int NN = 2*t.kzPad, TT = 128/NN;
        SiLayout lay;
        const bool four = energy || st.siForceDigits == 4;
        siLayout(NN, TT, four ? 4 : 3, ks.siRowStagePad, lay);            // (feasible: checked by the plan)
        ip.rowStages = lay.rowStages; ip.opStages = lay.opStages;
        ip.rowStageBytes = ks.tsRowStageBytes; ip.rowStagePad = ks.siRowStagePad; ip.offOp = lay.offOp; ip.opBytes = lay.opBytes; ip.offBar = lay.offBar;
        const dim3 grid(t.rowTiles, t.splits);
        if (t.kzPad == 32) {
            if (four) structureFactorI8Kernel<64, 2, 4><<<grid, SI_THREADS, lay.smem, s>>>(ip);
            else      structureFactorI8Kernel<64, 2, 3><<<grid, SI_THREADS, lay.smem, s>>>(ip);
        }
        else {
            if (four) structureFactorI8Kernel<128, 1, 4><<<grid, SI_THREADS, lay.smem, s>>>(ip);
            else      structureFactorI8Kernel<128, 1, 3><<<grid, SI_THREADS, lay.smem, s>>>(ip);
        }
        CFX_LAUNCH_CHECK(); st.launches++;
        return;
    }
    sp.rowStageBytes = ks.tsRowStageBytes; sp.offA = ks.tsOffA; sp.offB = ks.tsOffB; sp.offBar = ks.tsOffBar;
    if (t.kzPad == 32) structureFactorTensorKernel<64, 2><<<dim3(t.rowTiles, t.splits), ST_THREADS, t.smem, s>>>(sp);
    else               structureFactorTensorKernel<128, 1><<<dim3(t.rowTiles, t.splits), ST_THREADS, t.smem, s>>>(sp);
    CFX_LAUNCH_CHECK(); st.launches++;
}

// ------------------------------------------------------------------------------------------------
// TF32 tensor-core peak: every CTA issues `iters` tcgen05.mma kind::tf32 128x128x8 (A in tensor memory, B in shared
// memory -- the gather's instruction shape) on fixed operands. Roofline denominator of the tensor-core kernels.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(128) tf32PeakKernel(int iters, float* sink) {
    constexpr int N = 128;
    extern __shared__ __align__(128) unsigned char smem[];
    float* b = reinterpret_cast<float*>(smem);                 // [2][N][4]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmemSlot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < 2*N*4; e += 128) b[e] = roundTf32(0.001f*(e % 97));
    fenceProxyAsync();
    if (tid == 0) { mbarInit(&bar, 1); mbarFenceInit(); }
    if (warp == 0) tmemAlloc<512>(&tmemSlot);
    tcgen05FenceBefore();
    __syncthreads();
    tcgen05FenceAfter();
    const uint32_t tmem = tmemSlot;
    for (int c = 0; c < 64; c += 4) tmemStore4(tmem + 256 + c + ((uint32_t) (warp*32) << 16), make_float4(0.5f, 0.25f, 0.125f, 1.f));
    tmemWaitStore();
    tcgen05FenceBefore();
    __syncthreads();
    tcgen05FenceAfter();
    if (warp == 0) {
        constexpr uint32_t idesc = ummaIdescTf32(128, N);
        const uint64_t bd = ummaSmemDesc(smemU32(b), N*16, 128);
        if (electOne()) {
            for (int i = 0; i < iters; i++) ummaTf32TS(tmem + (i & 1)*N, tmem + 256 + 8*(i % 7), bd, idesc, i > 1);
            ummaCommit(&bar);
        }
        __syncwarp();
    }
    mbarWait(&bar, 0);
    tcgen05FenceAfter();
    float v[16];
    tmemLoad16(tmem + ((uint32_t) (warp*32) << 16), v);
    if (v[0] == 123.456f) sink[tid] = v[1];
    tcgen05FenceBefore();
    __syncthreads();
    if (warp == 0) tmemFree<512>(tmem);
}
} // namespace

namespace {
// the same for tcgen05.mma kind::i8 128x256x32 (both operands in shared memory: the integer structure-factor kernel's form)
__global__ void __launch_bounds__(128) i8PeakKernel(int iters, int* sink) {
    constexpr int N = 256;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmemSlot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < (128 + N)*32/4; e += 128) reinterpret_cast<uint32_t*>(smem)[e] = 0x01010101u*(e & 3);
    fenceProxyAsync();
    if (tid == 0) { mbarInit(&bar, 1); mbarFenceInit(); }
    if (warp == 0) tmemAlloc<512>(&tmemSlot);
    tcgen05FenceBefore();
    __syncthreads();
    tcgen05FenceAfter();
    const uint32_t tmem = tmemSlot;
    if (warp == 0) {
        const uint64_t ad = ummaSmemDesc(smemU32(smem), 128*16, 128), bd = ummaSmemDesc(smemU32(smem) + 128*32, N*16, 128);
        if (electOne()) {
            for (int i = 0; i < iters; i++) ummaI8SS(tmem + (i & 1)*N, ad, bd, ummaIdescS8(128, N), i > 1);
            ummaCommit(&bar);
        }
        __syncwarp();
    }
    mbarWait(&bar, 0);
    tcgen05FenceAfter();
    int v[16];
    tmemLoad16i(tmem + ((uint32_t) (warp*32) << 16), v);
    if (v[0] == 123456789) sink[tid] = v[1];
    tcgen05FenceBefore();
    __syncthreads();
    if (warp == 0) tmemFree<512>(tmem);
}
} // namespace

double measureI8Peak(int device, int iters) {
    CFX_CUDA(cudaSetDevice(device));
    int numSM = 148;
    cudaDeviceGetAttribute(&numSM, cudaDevAttrMultiProcessorCount, device);
    int* sink = nullptr;
    CFX_CUDA(cudaMalloc(&sink, 4096));
    cudaEvent_t e0, e1;
    CFX_CUDA(cudaEventCreate(&e0)); CFX_CUDA(cudaEventCreate(&e1));
    const size_t smem = (128 + 256)*32;
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CFX_CUDA(cudaEventRecord(e0));
        i8PeakKernel<<<numSM, 128, smem>>>(iters, sink);
        CFX_CUDA(cudaEventRecord(e1));
        CFX_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        CFX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    CFX_CUDA(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    return 2.0*128*256*32*(double) iters*numSM/(best*1e-3)*1e-12;
}

double measureTf32Peak(int device, int iters) {
    CFX_CUDA(cudaSetDevice(device));
    int numSM = 148;
    cudaDeviceGetAttribute(&numSM, cudaDevAttrMultiProcessorCount, device);
    float* sink = nullptr;
    CFX_CUDA(cudaMalloc(&sink, 4096));
    cudaEvent_t e0, e1;
    CFX_CUDA(cudaEventCreate(&e0)); CFX_CUDA(cudaEventCreate(&e1));
    const size_t smem = 2*128*4*sizeof(float);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        CFX_CUDA(cudaEventRecord(e0));
        tf32PeakKernel<<<numSM, 128, smem>>>(iters, sink);
        CFX_CUDA(cudaEventRecord(e1));
        CFX_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        CFX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0) best = std::min(best, ms);
    }
    CFX_CUDA(cudaGetLastError());
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(sink);
    return 2.0*128*128*8*(double) iters*numSM/(best*1e-3)*1e-12;
}

} // namespace cfx

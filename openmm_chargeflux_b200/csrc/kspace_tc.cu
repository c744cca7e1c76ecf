// kspace_tc.cu -- the reciprocal-space force / dE/dq gather of piece (3) on the 5th-generation tensor cores.
//
// After the factorisation of kspace.cu the gather is a dense contraction: for every atom a and signed row
// r = (nx, ny)
//     (Ur, Ui, Vr, Vi)[a][r] = sum_k Z[a][k] * C[r][comp][k],     k = (|nz| = l, cos|sin),  K = 2 Kz
// with Z[a][2l] = cos(2 pi l z_a), Z[a][2l+1] = sin(..) and the coefficient rows
//     Ur: (Ar, Br)   Ui: (Ai, Bi)   Vr: (l Bi, -l Ai)   Vi: (-l Br, l Ar)
// (A, B as defined in coefficientKernel). That is a [atoms x K] x [K x 4 rows] GEMM with K ~ 56: 21 GFLOP at
// 32k atoms. It runs as tcgen05.mma kind::tf32 with FP32 accumulators in tensor memory, made FP32-accurate by
// the three-product split  x = hi + lo (both TF32):  Z C ~ Zlo Chi + Zhi Clo + Zhi Chi  (measured relative RMS
// error 1.7e-7 at K = 56, tools/umma_test.cu). The epilogue (T = Ex Ey, dE/dq += Re(T U), F += q g (nx Im TU,
// ny Im TU, Im TU')) stays on the CUDA cores, one thread per atom = per tensor-memory lane.
//
// One persistent CTA per SM, three roles:
//   warp 0     bulk-TMA producer: streams coefficient tiles (NT columns x K, hi and lo planes, stored by
//              coefficientKernel in the canonical K-major no-swizzle core-matrix layout) through a ring of
//              shared-memory stages
//   warp 1     MMA issuer (one lane): A = phase tile held in tensor memory (written there once per atom group
//              by the epilogue warps with tcgen05.st), B = coefficient stage, D = accumulator slot in TMEM
//   warps 2-5  epilogue: tcgen05.ld the accumulator slot, apply T and accumulate the four outputs of their atom
//              in registers; one fixed-point atomic per output when the atom group changes
// Work units (atom group, column tile) are split contiguously over the CTAs, atom-group major.
#include "cfx_internal.cuh"
#include "ptx_sm100.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <stdexcept>

namespace cfx {

namespace {

inline float __int_as_float_host(int v) { float f; memcpy(&f, &v, 4); return f; }
__device__ __forceinline__ void namedBarrier(int id, int threads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory"); }

constexpr int GT_THREADS = 64 + 512;        // producer warp, MMA warp, up to 16 epilogue warps
constexpr int GT_TILE_ATOMS = 128;

struct GTParams {
    const float* zSplit;        // [atom tile][hi|lo][KC][128][4]
    const float* coefT;         // [column tile][hi|lo][KC][NT][4]
    const float4* rowData; const float2* colX; const float2* colY; const float* qf;
    int Ky, Kp, KC, N, Npad;
    int signedLo, signedHi, numColTiles, numAtomGroups;
    float fx, fy, fz;
    int stages;
    uint32_t stageBytes, offEy, offBar;
};

template <int MT, int NT>
__global__ void __launch_bounds__(GT_THREADS, 1) gatherTensorKernel(GTParams p, long long* __restrict__ forceFixed, long long* __restrict__ dedqFixed) {
    constexpr int ROWS = NT/4;                       // signed rows per column tile
    constexpr int SUBS = ROWS/8;                     // epilogue warps per lane quarter, 8 rows each
    extern __shared__ __align__(1024) unsigned char smem[];
    float2* Eys = reinterpret_cast<float2*>(smem + p.offEy);                 // [Ky][MT*128], thread-private columns
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.offBar);
    uint64_t* coefFull = bars;                       // [stages]
    uint64_t* coefEmpty = bars + p.stages;           // [stages]
    uint64_t* dFull = bars + 2*p.stages;             // [2]
    uint64_t* dEmpty = dFull + 2;                    // [2]
    uint64_t* aFull = dEmpty + 2;                    // [1]
    uint32_t* tmemSlot = reinterpret_cast<uint32_t*>(aFull + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long totalUnits = (long long) p.numAtomGroups*p.numColTiles;
    const int u0 = (int) (totalUnits*blockIdx.x/gridDim.x), u1 = (int) (totalUnits*(blockIdx.x + 1)/gridDim.x);

    if (tid == 0) {
        for (int s = 0; s < p.stages; s++) { mbarInit(&coefFull[s], 1); mbarInit(&coefEmpty[s], 1); }
        for (int d = 0; d < 2; d++) { mbarInit(&dFull[d], 1); mbarInit(&dEmpty[d], 128*SUBS); }
        mbarInit(aFull, 128*SUBS);
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

    if (warp == 0) {
        // ---------------- producer ----------------
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int unit = u0; unit < u1; unit++) {
                const int colTile = unit % p.numColTiles;
                mbarWait(&coefEmpty[s], ph ^ 1);
                mbarExpectTx(&coefFull[s], p.stageBytes);
                bulkLoad(smem + (size_t) s*p.stageBytes, p.coefT + (size_t) colTile*(p.stageBytes/4), p.stageBytes, &coefFull[s]);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
        __syncwarp();
    }
    else if (warp == 1) {
        // ---------------- MMA issuer ----------------
        if (lane == 0) {
            constexpr uint32_t idesc = ummaIdescTf32(128, NT);
            const uint32_t planeBytes = (uint32_t) p.KC*NT*16;           // one hi or lo plane of a stage
            int s = 0; uint32_t ph = 0, aPh = 0;
            int curGroup = -1;
            uint32_t seq = 0;
            for (int unit = u0; unit < u1; unit++) {
                const int group = unit/p.numColTiles;
                if (group != curGroup) {
                    curGroup = group;
                    mbarWait(aFull, aPh); aPh ^= 1;
                    tcgen05FenceAfter();
                }
                mbarWait(&coefFull[s], ph);
                tcgen05FenceAfter();
                const uint32_t stageAddr = smemU32(smem + (size_t) s*p.stageBytes);
                #pragma unroll 1
                for (int t = 0; t < MT; t++, seq++) {
                    const uint32_t d = seq & 1;
                    mbarWait(&dEmpty[d], ((seq >> 1) & 1) ^ 1);
                    tcgen05FenceAfter();
                    const uint32_t tD = tmem + d*128;
                    const uint32_t aHi = tmemA + (uint32_t) t*2*p.Kp, aLo = aHi + p.Kp;
                    uint32_t acc = 0;
                    // small products first: Zlo Chi, Zhi Clo, then Zhi Chi
                    #pragma unroll 1
                    for (int pass = 0; pass < 3; pass++) {
                        const uint32_t a = pass == 0 ? aLo : aHi;
                        const uint32_t b = stageAddr + (pass == 1 ? planeBytes : 0);
                        for (int k8 = 0; k8 < p.Kp/8; k8++) {
                            ummaTf32TS(tD, a + k8*8, ummaSmemDesc(b + (uint32_t) k8*2*NT*16, NT*16, 128), idesc, acc);
                            acc = 1;
                        }
                    }
                    ummaCommit(&dFull[d]);
                }
                ummaCommit(&coefEmpty[s]);            // the stage is free once these MMAs have read it
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
        __syncwarp();
    }
    else if (warp - 2 < 4*SUBS) {
        // ---------------- epilogue: thread = (tensor-memory lane = atom of the tile, group of 8 rows) ----------------
        const int q4 = warp & 3;                                  // lane quarter this warp may access
        const int sub = (warp - 2) >> 2;                          // which 8 rows (32 accumulator columns) of the tile
        const int atomInTile = q4*32 + lane;
        const uint32_t laneBase = (uint32_t) (q4*32) << 16;
        constexpr int EPI_THREADS = 128*SUBS;
        float oD[MT], oX[MT], oY[MT], oZ[MT], nxOf[MT];
        float2 ex[MT];
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
        for (int unit = u0; unit < u1; unit++) {
            const int group = unit/p.numColTiles, colTile = unit - group*p.numColTiles;
            if (group != curGroup) {
                if (curGroup >= 0) flush();
                namedBarrier(1, EPI_THREADS);                     // every epilogue warp is done with the old Ey columns
                curGroup = group;
                // phase operand of the new atom group -> tensor memory (all MMAs that read the old one are
                // complete: their last accumulator has been consumed), Ey columns -> shared memory
                #pragma unroll
                for (int t = 0; t < MT; t++) {
                    const int tile = group*MT + t;
                    const float4* src = reinterpret_cast<const float4*>(p.zSplit) + (size_t) tile*2*p.KC*GT_TILE_ATOMS + atomInTile;
                    const uint32_t dst = tmemA + (uint32_t) t*2*p.Kp + laneBase;
                    for (int c = sub; c < 2*p.KC; c += SUBS) tmemStore4(dst + 4*c, src[(size_t) c*GT_TILE_ATOMS]);   // hi plane then lo plane
                    const int atom = tile*GT_TILE_ATOMS + atomInTile;
                    for (int m = sub; m < p.Ky; m += SUBS) Eys[m*(MT*GT_TILE_ATOMS) + t*GT_TILE_ATOMS + atomInTile] = p.colY[(size_t) m*p.Npad + atom];
                    oD[t] = 0.f; oX[t] = 0.f; oY[t] = 0.f; oZ[t] = 0.f; ex[t] = make_float2(0.f, 0.f); nxOf[t] = -1.f;
                }
                tmemWaitStore();
                tcgen05FenceBefore();
                mbarArrive(aFull);
                namedBarrier(1, EPI_THREADS);                     // Ey columns complete
            }
            const float4* rd = p.rowData + p.signedLo + colTile*ROWS + 8*sub;
            #pragma unroll
            for (int t = 0; t < MT; t++, seq++) {
                const uint32_t d = seq & 1;
                mbarWait(&dFull[d], (seq >> 1) & 1);
                tcgen05FenceAfter();
                float v[32];
                tmemLoad32(tmem + d*128 + laneBase + 32*sub, v);
                tcgen05FenceBefore();
                mbarArrive(&dEmpty[d]);                           // values are in registers: the slot can be refilled
                const int atom = (group*MT + t)*GT_TILE_ATOMS + atomInTile;
                const float2* eyCol = Eys + t*GT_TILE_ATOMS + atomInTile;
                #pragma unroll
                for (int i = 0; i < 8; i++) {
                    // padding rows beyond signedHi have zero coefficients (U = V = 0) and row data (0,0,0,1)
                    const float4 r = __ldg(rd + i);               // (nx, ny, |ny|*stride as int bits, sign of ny)
                    if (r.x != nxOf[t]) {
                        nxOf[t] = r.x;
                        ex[t] = p.colX[(size_t) ((int) r.x)*p.Npad + atom];
                    }
                    float2 ey = eyCol[__float_as_int(r.z)];
                    ey.y *= r.w;
                    const float tr = ex[t].x*ey.x - ex[t].y*ey.y;
                    const float ti = ex[t].x*ey.y + ex[t].y*ey.x;
                    const float ur = v[4*i], ui = v[4*i+1], vr = v[4*i+2], vi = v[4*i+3];
                    oD[t] = fmaf(tr, ur, oD[t]);  oD[t] = fmaf(-ti, ui, oD[t]);
                    const float im = tr*ui + ti*ur;
                    oX[t] = fmaf(r.x, im, oX[t]);
                    oY[t] = fmaf(r.y, im, oY[t]);
                    oZ[t] = fmaf(tr, vi, oZ[t]);  oZ[t] = fmaf(ti, vr, oZ[t]);
                }
            }
        }
        if (curGroup >= 0) flush();
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
    if (Kp > 112) return;                                        // phase operand must fit 2 x 224 tensor-memory columns
    ks.tKp = Kp; ks.tKC = Kp/4;
    ks.tMT = (Kp <= 56) ? 2 : 1;
    ks.tNT = (Kp <= 56) ? 128 : 64;
    const size_t stageBytes = (size_t) 2*ks.tKC*ks.tNT*16;
    const size_t eyBytes = ((size_t) Ky*ks.tMT*GT_TILE_ATOMS*sizeof(float2) + 127) & ~(size_t) 127;
    const size_t cap = 227*1024 - 1024;                          // alignment slack of the dynamic shared-memory base
    int stages = (int) std::min<size_t>(4, (cap - eyBytes - 256)/stageBytes);
    if (stages < 2) return;
    ks.tStages = stages;
    ks.tStageBytes = (uint32_t) stageBytes;
    ks.tOffEy = (uint32_t) (stages*stageBytes);
    ks.tOffBar = (uint32_t) (ks.tOffEy + eyBytes);
    ks.tSmem = ks.tOffBar + 256;
    const int rows = ks.tNT/4;
    const int signedHere = std::max(ks.signedHi - ks.signedLo, 1);
    ks.tColTiles = (signedHere + rows - 1)/rows;
    const size_t zFloats = (size_t) (st.Npad/GT_TILE_ATOMS)*2*ks.tKC*GT_TILE_ATOMS*4;
    const size_t cFloats = (size_t) ks.tColTiles*2*ks.tKC*ks.tNT*4;
    CFX_CUDA(cudaMalloc(&st.zSplit, zFloats*sizeof(float)));
    CFX_CUDA(cudaMemset(st.zSplit, 0, zFloats*sizeof(float)));
    CFX_CUDA(cudaMalloc(&st.coefT, cFloats*sizeof(float)));
    CFX_CUDA(cudaMemset(st.coefT, 0, cFloats*sizeof(float)));
    // per signed row: (nx, ny, |ny| * Ey column stride, sign of ny), same order as gRowInfo, zero-padded
    {
        const int Kx = ks.K[0];
        std::vector<float4> rd;
        const int stride = ks.tMT*GT_TILE_ATOMS;
        for (int row = 0; row < Kx*Ky; row++) {
            const int nx = row/Ky, m = row % Ky;
            rd.push_back(make_float4((float) nx, (float) m, __int_as_float_host(m*stride), 1.f));
            if (nx > 0 && m > 0) rd.push_back(make_float4((float) nx, (float) -m, __int_as_float_host(m*stride), -1.f));
        }
        for (int k = 0; k < 64; k++) rd.push_back(make_float4(0.f, 0.f, __int_as_float_host(0), 1.f));
        CFX_CUDA(cudaMalloc(&st.gRowData, rd.size()*sizeof(float4)));
        CFX_CUDA(cudaMemcpy(st.gRowData, rd.data(), rd.size()*sizeof(float4), cudaMemcpyHostToDevice));
    }
    CFX_CUDA(cudaFuncSetAttribute(gatherTensorKernel<2, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
    CFX_CUDA(cudaFuncSetAttribute(gatherTensorKernel<1, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
    ks.tensorGather = true;
}

void launchGatherTensor(State& st, long long* dForce, long long* dDedq, cudaStream_t s) {
    KSpacePlan& ks = st.ks;
    GTParams gp;
    gp.zSplit = st.zSplit; gp.coefT = st.coefT; gp.rowData = st.gRowData; gp.colX = st.colX; gp.colY = st.colY; gp.qf = st.qf;
    gp.Ky = ks.K[1]; gp.Kp = ks.tKp; gp.KC = ks.tKC; gp.N = st.N; gp.Npad = st.Npad;
    gp.signedLo = ks.signedLo; gp.signedHi = ks.signedHi; gp.numColTiles = ks.tColTiles;
    gp.numAtomGroups = st.Npad/(GT_TILE_ATOMS*ks.tMT);
    gp.fx = (float) (2*M_PI/st.box.L[0]); gp.fy = (float) (2*M_PI/st.box.L[1]); gp.fz = (float) (2*M_PI/st.box.L[2]);
    gp.stages = ks.tStages; gp.stageBytes = ks.tStageBytes; gp.offEy = ks.tOffEy; gp.offBar = ks.tOffBar;
    int numSM = 148;
    cudaDeviceGetAttribute(&numSM, cudaDevAttrMultiProcessorCount, st.device);
    const long long units = (long long) gp.numAtomGroups*gp.numColTiles;
    const int grid = (int) std::min<long long>(numSM, units);
    if (ks.tMT == 2) gatherTensorKernel<2, 128><<<grid, GT_THREADS, ks.tSmem, s>>>(gp, dForce, dDedq);
    else             gatherTensorKernel<1, 64><<<grid, GT_THREADS, ks.tSmem, s>>>(gp, dForce, dDedq);
    CFX_LAUNCH_CHECK(); st.launches++;
}

} // namespace cfx

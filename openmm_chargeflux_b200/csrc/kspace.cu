// kspace.cu -- piece (3): explicit-k reciprocal space (ReferenceCoulKernels.cpp:513-556).
//
// The reference evaluates, for every half-space k = 2*pi*(nx/Lx, ny/Ly, nz/Lz) with nx in [0,Kx),
// ny,nz in (-K,K), the structure factor S(k) = sum_j q_j exp(i k.r_j) and then
//   E += C a_k |S|^2,  F_i -= 2 C a_k q_i (S_s cos - S_c sin) k,  dE/dq_i += 2 C a_k (S_c cos + S_s sin)
// with 4 libm trig calls per (atom,k). Here exp(i k.r) = Ex(nx) Ey(ny) Ez(nz) is factorised:
//
//  phaseTableKernel   per-atom, per-axis phases exp(2 pi i n u), FP64 sincospi of the wrapped
//                     fractional coordinate + FP64 recurrence, stored as float2.
//  structureFactorKernel  (S) rows = (nx,|ny|), cols = |nz|. Per (row,col) the eight real sums
//                     P[a][b][c] = sum_j (q X)_a Y_b Z_c  (a: re/im of q*Ex, b: cos/sin of Ey, c: cos/sin of Ez)
//                     give S at all four sign combinations (+-ny, +-nz): 8 FMA per 4 k-vectors
//                     = 2 FMA per (atom,k). Register tile TM rows x TN cols per thread; per-atom phase
//                     rows are streamed into shared memory with bulk-TMA (cp.async.bulk + mbarrier).
//  coefficientKernel  sums the per-split partials in FP64, a_k = exp(-k^2/4alpha^2)/k^2 once per k,
//                     E_recip in FP64, and gather coefficients A,B per (signed row, |nz|).
//  gatherKernel       per atom: U = sum_l A_l c_l + B_l s_l, U' = sum_l l(-iB_l c_l + iA_l s_l) over |nz|
//                     (4 FMA per (atom,k)), then T = Ex(nx)Ey(ny): dE/dq += Re(T U),
//                     F += q (2pi/L) (nx Im(T U), ny Im(T U), Im(T U')).
//
// All main loops are FP32 FMA; cross-CTA reductions, a_k and energies are FP64 / fixed point.
//
// structureFactorKernel and gatherKernel here are the CUDA-core versions: they serve geometries outside the tensor-core
// variants (kmax_z > 64, ...) and A/B measurements (CFX_KSPACE_S=fp32, CFX_KSPACE_GATHER=fp32). Whenever the geometry
// allows, launchKSpace runs the tcgen05 kernels of kspace_tc.cu instead (same tables, same coefficient kernel): the
// integer structure-factor kernel (exact sums: energy and forces-only calls alike) and the TF32 x 3 gather.
#include "cfx_internal.cuh"
#include "ptx_sm100.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <stdexcept>

namespace cfx {

namespace {

// ------------------------------------------------------------------------------------------------
// phase tables
// ------------------------------------------------------------------------------------------------
struct TableParams {
    int N, Npad, Kx, Ky, Kz, kzPad, zOff, rowPitch;
    int TN, TNP;                 // S-kernel column grouping: |nz| = l lives in slot (l/TN)*TNP + l%TN
    double invLx, invLy, invLz;
    float* zSplit; int KC;       // tensor gather operand (kspace_tc.cu), nullptr when the FP32 gather is used
};

// One CTA = 32 atoms x 3 axes (96 working threads). Each thread runs the FP64 recurrence of one (atom, axis);
// the atom-major rows used by the S kernel are assembled in shared memory and written out as one contiguous,
// fully coalesced block; the n-major columns used by the gather kernel are coalesced over atoms as they are.
// Padded atoms (>= N) get all-zero phases.
#define PT_ATOMS 32
__global__ void __launch_bounds__(128) phaseTableKernel(TableParams p, const double* __restrict__ pos, const float* __restrict__ qf,
        float2* __restrict__ rowS, float2* __restrict__ colX, float2* __restrict__ colY, float4* __restrict__ colZ4) {
    extern __shared__ float2 rows[];                      // [PT_ATOMS][rowPitch]
    const int la = threadIdx.x & 31, axis = threadIdx.x >> 5;
    const int atom = blockIdx.x*PT_ATOMS + la;
    for (int e = threadIdx.x; e < PT_ATOMS*p.rowPitch; e += blockDim.x) rows[e] = make_float2(0.f, 0.f);   // padding slots stay zero
    __syncthreads();
    if (axis < 3) {
        const int K = axis == 0 ? p.Kx : (axis == 1 ? p.Ky : p.Kz);
        float2* row = rows + la*p.rowPitch + (axis == 0 ? 0 : (axis == 1 ? p.Kx : p.zOff));
        const bool real = atom < p.N;
        double s1 = 0.0, c1 = 0.0;
        float scale = 0.f;
        if (real) {
            const double invL = axis == 0 ? p.invLx : (axis == 1 ? p.invLy : p.invLz);
            double u = pos[3*(size_t) atom + axis]*invL;
            u -= floor(u);
            sincospi(2.0*u, &s1, &c1);
            scale = axis == 0 ? qf[atom] : 1.0f;
        }
        double c = real ? 1.0 : 0.0, sn = 0.0;
        for (int n = 0; n < K; n++) {
            const float cf = (float) c, sf = (float) sn;
            row[axis == 2 ? (n/p.TN)*p.TNP + n % p.TN : n] = make_float2(scale*cf, scale*sf);
            if (axis == 0) colX[(size_t) n*p.Npad + atom] = make_float2(cf, sf);
            else if (axis == 1) colY[(size_t) n*p.Npad + atom] = make_float2(cf, sf);
            else {
                colZ4[(size_t) n*p.Npad + atom] = make_float4(cf, sf, (float) n*cf, (float) n*sf);
                if (p.zSplit) {
                    // k = 2n (cos), 2n+1 (sin): [tile][k/4][128 atoms][4]
                    float* dst = p.zSplit + (((size_t) (atom >> 7)*p.KC + (n >> 1))*128 + (atom & 127))*4 + (n & 1)*2;
                    *reinterpret_cast<float2*>(dst) = make_float2(cf, sf);
                }
            }
            const double cn = c*c1 - sn*s1;
            sn = c*s1 + sn*c1;
            c = cn;
        }
    }
    __syncthreads();
    float4* dst = reinterpret_cast<float4*>(rowS + (size_t) blockIdx.x*PT_ATOMS*p.rowPitch);
    const float4* src = reinterpret_cast<const float4*>(rows);
    for (int e = threadIdx.x; e < PT_ATOMS*p.rowPitch/2; e += blockDim.x) dst[e] = src[e];
}

// ------------------------------------------------------------------------------------------------
// structure factors
// ------------------------------------------------------------------------------------------------
#define S_ATOMS_PER_STAGE 32
#define S_BM 64                    // rows per CTA: lane + 32*i, i < 2
#define S_MAX_WARPS 8
#ifndef S_UNROLL
#define S_UNROLL 2
#endif
#ifndef S_MINBLOCKS
#define S_MINBLOCKS 3
#endif
constexpr int kSUnroll = S_UNROLL;

struct SParams {
    const float2* rowS; float* part;
    int rowPitch, Kx, Ky, zOff, kzPad;
    int stages;
    int rowLo, rowHi, numRows;
    int atomsPerSplit, Npad;
};

// One CTA = 64 rows x all |nz| columns over one split of the atoms. Warp g owns column group g (TN
// columns, padded to TNP float2 so the group is 16-byte aligned) for all 64 rows; lane owns rows
// lane and lane+32. Per atom a warp reads the row phases q*Ex(nx), Ey(|ny|) of its two rows (4 LDS.64)
// and forms the row operand a = (xr*yc, xr*ys, xi*yc, xi*ys) in registers (8 FMUL), reads its TN column
// phases with warp-UNIFORM LDS.128 (1 wavefront each) and issues 8*TN packed FFMA2 (one per (component, column): the
// (cos, sin) pair of products; 0.244 -> 0.218 ms at 32k atoms against scalar FFMA: the kernel is issue-bound). The atom rows arrive by
// bulk TMA into a ring of stages; there is one CTA barrier per 32 atoms and no staging of the operand.
template <int TN, int G>
__global__ void __launch_bounds__(32*G, (G <= 4) ? S_MINBLOCKS : 1) structureFactorKernel(SParams p) {
    constexpr int TNP = (TN + 1) & ~1;
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem);
    int* done = reinterpret_cast<int*>(smem + 64);         // per slot: warps that have finished reading it
    float2* raw = reinterpret_cast<float2*>(smem + 128);
    const int stageElems = S_ATOMS_PER_STAGE*p.rowPitch;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int rowBase = p.rowLo + blockIdx.x*S_BM;
    const int rowEnd = min(rowBase + S_BM, p.rowHi);
    const int atomBegin = blockIdx.y*p.atomsPerSplit;
    const int atomEnd = min(atomBegin + p.atomsPerSplit, p.Npad);
    const int numStages = (atomEnd - atomBegin)/S_ATOMS_PER_STAGE;
    const uint32_t stageBytes = (uint32_t) (stageElems*sizeof(float2));

    // phase slots of my two rows inside an atom row (rows past the end compute on row 0's slots and
    // are never stored)
    int offX[2], offY[2];
    #pragma unroll
    for (int it = 0; it < 2; it++) {
        const int row = min(rowBase + lane + 32*it, p.numRows - 1);
        const int nx = row/p.Ky;
        offX[it] = nx;
        offY[it] = p.Kx + (row - nx*p.Ky);
    }

    float2 acc2[2][TN][4];                                 // (zc, zs) products of row operand component m
    #pragma unroll
    for (int i = 0; i < 2; i++)
        #pragma unroll
        for (int c = 0; c < TN; c++)
            #pragma unroll
            for (int k = 0; k < 4; k++) acc2[i][c][k] = make_float2(0.f, 0.f);

    if (tid == 0) {
        for (int s = 0; s < p.stages; s++) { mbarInit(mbar + s, 1); done[s] = 0; }
        mbarFenceInit();
    }
    __syncthreads();
    if (tid == 0)
        for (int s = 0; s < p.stages && s < numStages; s++) {
            mbarExpectTx(mbar + s, stageBytes);
            bulkLoad(raw + (size_t) s*stageElems, p.rowS + (size_t) (atomBegin + s*S_ATOMS_PER_STAGE)*p.rowPitch, stageBytes, mbar + s);
        }

    // Warps run free of each other: a warp waits only for the TMA of its next stage; the last warp to
    // finish a stage re-arms the slot's mbarrier and issues the refill (no CTA barrier in the loop).
    for (int st = 0; st < numStages; st++) {
        const int slot = st % p.stages;
        mbarWait(mbar + slot, (uint32_t) ((st/p.stages) & 1));
        const float2* rw = raw + (size_t) slot*stageElems;
        const float2* x0p = rw + offX[0]; const float2* y0p = rw + offY[0];
        const float2* x1p = rw + offX[1]; const float2* y1p = rw + offY[1];
        const float4* bPtr = reinterpret_cast<const float4*>(rw + p.zOff + warp*TNP);
        const int bPitch4 = p.rowPitch/2;
        #pragma unroll (kSUnroll)
        for (int j = 0; j < S_ATOMS_PER_STAGE; j++) {
            const float2 x0 = x0p[j*p.rowPitch], y0 = y0p[j*p.rowPitch];
            const float2 x1 = x1p[j*p.rowPitch], y1 = y1p[j*p.rowPitch];
            float4 b[TNP/2];
            #pragma unroll
            for (int c = 0; c < TNP/2; c++) b[c] = bPtr[j*bPitch4 + c];
            const float4 a0 = make_float4(x0.x*y0.x, x0.x*y0.y, x0.y*y0.x, x0.y*y0.y);
            const float4 a1 = make_float4(x1.x*y1.x, x1.x*y1.y, x1.y*y1.x, x1.y*y1.y);
            const float av0[4] = {a0.x, a0.y, a0.z, a0.w};
            const float av1[4] = {a1.x, a1.y, a1.z, a1.w};
            // packed FP32: one FFMA2 per (row operand component, column) = the (zc, zs) pair of products
            #pragma unroll
            for (int m = 0; m < 4; m++) {
                const float2 am = make_float2(av0[m], av0[m]);
                #pragma unroll
                for (int c = 0; c < TN; c++) {
                    const float2 z = (c & 1) ? make_float2(b[c/2].z, b[c/2].w) : make_float2(b[c/2].x, b[c/2].y);
                    acc2[0][c][m] = __ffma2_rn(am, z, acc2[0][c][m]);
                }
            }
            #pragma unroll
            for (int m = 0; m < 4; m++) {
                const float2 am = make_float2(av1[m], av1[m]);
                #pragma unroll
                for (int c = 0; c < TN; c++) {
                    const float2 z = (c & 1) ? make_float2(b[c/2].z, b[c/2].w) : make_float2(b[c/2].x, b[c/2].y);
                    acc2[1][c][m] = __ffma2_rn(am, z, acc2[1][c][m]);
                }
            }
        }
        __syncwarp();
        if (lane == 0 && st + p.stages < numStages) {
            __threadfence_block();
            if (atomicAdd(done + slot, 1) == (int) (blockDim.x >> 5) - 1) {   // every warp has read this slot
                done[slot] = 0;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbarExpectTx(mbar + slot, stageBytes);
                bulkLoad(raw + (size_t) slot*stageElems, p.rowS + (size_t) (atomBegin + (st + p.stages)*S_ATOMS_PER_STAGE)*p.rowPitch,
                         stageBytes, mbar + slot);
            }
        }
    }
    #pragma unroll
    for (int i = 0; i < 2; i++) {
        const int row = rowBase + lane + 32*i;
        if (row >= rowEnd) continue;
        #pragma unroll
        for (int c = 0; c < TN; c++) {
            const int col = warp*TNP + c;
            float4* out = reinterpret_cast<float4*>(p.part + (((size_t) blockIdx.y*p.numRows + row)*p.kzPad + col)*8);
            out[0] = make_float4(acc2[i][c][0].y, acc2[i][c][0].x, acc2[i][c][1].y, acc2[i][c][1].x);      // (zs, zc) order of the partials
            out[1] = make_float4(acc2[i][c][2].y, acc2[i][c][2].x, acc2[i][c][3].y, acc2[i][c][3].x);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// coefficients + reciprocal energy
// ------------------------------------------------------------------------------------------------
struct CoefParams {
    const float* part; float4* coef; const int* signedStart;
    int Kx, Ky, Kz, kzPad, numRows, splits, rowLo, rowHi, TN, TNP;
    double gx, gy, gz;          // 2 pi / L
    double C;                   // 4 pi ke / V
    double invFourAlpha2;
    bool energy, forces, swapPairs;
    float* coefT; int KC, NT, signedLo;     // tensor gather operand (kspace_tc.cu) or nullptr
};

__global__ void __launch_bounds__(128) coefficientKernel(CoefParams p, long long* __restrict__ energyFixed) {
    __shared__ double scratch[32];
    // 4 lanes share one (row, |nz|): each sums a quarter of the atom splits, then a 2-step shuffle reduction
    const int gt = blockIdx.x*blockDim.x + threadIdx.x;
    const int t = gt >> 2, sub = gt & 3;
    const int rowsHere = p.rowHi - p.rowLo;
    double en = 0.0;
    const bool valid = t < rowsHere*p.Kz;
    {
        const int tt = valid ? t : 0;
        const int row = p.rowLo + tt/p.Kz, l = tt % p.Kz;
        const int nx = row/p.Ky, m = row - nx*p.Ky;
        double P[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int s = sub; s < p.splits; s += 4) {
            const int slotL = (l/p.TN)*p.TNP + l % p.TN;
            const float4* src = reinterpret_cast<const float4*>(p.part + (((size_t) s*p.numRows + row)*p.kzPad + slotL)*8);
            const float4 v0 = src[0], v1 = src[1];
            P[0] += v0.x; P[1] += v0.y; P[2] += v0.z; P[3] += v0.w;
            P[4] += v1.x; P[5] += v1.y; P[6] += v1.z; P[7] += v1.w;
        }
        #pragma unroll
        for (int k = 0; k < 8; k++) {
            P[k] += __shfl_xor_sync(0xffffffffu, P[k], 1);
            P[k] += __shfl_xor_sync(0xffffffffu, P[k], 2);
        }
        // stored order is {rcs, rcc, rss, rsc, ics, icc, iss, isc} (zs product first, see the S kernel)
        if (p.swapPairs) { double t; t = P[0]; P[0] = P[1]; P[1] = t; t = P[2]; P[2] = P[3]; P[3] = t; t = P[4]; P[4] = P[5]; P[5] = t; t = P[6]; P[6] = P[7]; P[7] = t; }
        // P = {rcc, rcs, rsc, rss, icc, ics, isc, iss}
        const double kx = nx*p.gx, ky = m*p.gy, kz = l*p.gz;
        const double k2 = kx*kx + ky*ky + kz*kz;
        const double ak = (nx | m | l) ? exp(-k2*p.invFourAlpha2)/k2 : 0.0;
        double Gre[2][2], Gim[2][2];     // [sy: 0=+,1=-][sz: 0=+,1=-]
        #pragma unroll
        for (int iy = 0; iy < 2; iy++)
            #pragma unroll
            for (int iz = 0; iz < 2; iz++) {
                const double sy = iy ? -1.0 : 1.0, sz = iz ? -1.0 : 1.0;
                const bool exists = !(iy && m == 0) && !(iz && l == 0);
                const int ny = iy ? -m : m, nz = iz ? -l : l;
                const bool inHalf = exists && (nx > 0 || ny > 0 || (ny == 0 && nz > 0));
                const double re = P[0] - sy*sz*P[3] - sz*P[5] - sy*P[6];
                const double im = P[4] - sy*sz*P[7] + sz*P[1] + sy*P[2];
                if (inHalf) {
                    if (valid && sub == 0) en += p.C*ak*(re*re + im*im);
                    Gre[iy][iz] = 2.0*p.C*ak*re;
                    Gim[iy][iz] = 2.0*p.C*ak*im;
                }
                else { Gre[iy][iz] = 0.0; Gim[iy][iz] = 0.0; }
            }
        if (p.forces && valid && sub == 0) {
            // signed rows of this unsigned row: (nx,+m) first, then (nx,-m) when it exists
            const int sBase = p.signedStart[row];
            const int nSigned = p.signedStart[row+1] - sBase;
            for (int iy = 0; iy < nSigned; iy++) {
                // H+- = conj(G(ny, +-l)); A = H+ + H-, B = i (H+ - H-)
                const double hpr = Gre[iy][0], hpi = -Gim[iy][0];
                const double hmr = Gre[iy][1], hmi = -Gim[iy][1];
                float4 c;
                c.x = (float) (hpr + hmr);
                c.y = (float) (hpi + hmi);
                c.z = (float) (-(hpi - hmi));
                c.w = (float) (hpr - hmr);
                p.coef[(size_t) (sBase + iy)*p.Kz + l] = c;
                if (p.coefT) {
                    // four GEMM columns (Ur, Ui, Vr, Vi) of this signed row, k = 2l (cos), 2l+1 (sin): [tile][k/4][NT][4]
                    const double Ar = hpr + hmr, Ai = hpi + hmi, Br = -(hpi - hmi), Bi = hpr - hmr, dl = (double) l;
                    const double cv[4][2] = {{Ar, Br}, {Ai, Bi}, {dl*Bi, -dl*Ai}, {-dl*Br, dl*Ar}};
                    const int rl = sBase + iy - p.signedLo, rows = p.NT >> 2;
                    const int tile = rl/rows, rr = rl - tile*rows;
                    float* base = p.coefT + (size_t) tile*p.KC*p.NT*4 + ((size_t) (l >> 1)*p.NT + rr*4)*4 + (l & 1)*2;
                    #pragma unroll
                    for (int comp = 0; comp < 4; comp++)
                        *reinterpret_cast<float2*>(base + comp*4) = make_float2((float) cv[comp][0], (float) cv[comp][1]);
                }
            }
        }
    }
    if (p.energy) {
        en = blockSum(en, scratch);
        if (threadIdx.x == 0) atomicAddEnergy(energyFixed + CFX_E_RECIP, en);
    }
}

// ------------------------------------------------------------------------------------------------
// force / dE/dq gather
// ------------------------------------------------------------------------------------------------
#define G_THREADS 384
#define G_WARPS 12
#define G_MAX_ROW_TILE (G_WARPS*4)

struct GParams {
    const float4* coef; const int2* rowInfo; const float2* colX; const float2* colY; const float4* colZ4;
    const float* qf;
    int Kx, Ky, Kz, N, Npad;
    int signedLo, signedHi, numRowTiles, numAtomTiles;
    float fx, fy, fz;            // 2 pi / L
    size_t offEy, offCoef, offInfo;     // shared-memory carve-up (bytes)
    int nbuf;                           // 2: coefficient tiles double-buffered, 1: single buffer (large kmax)
};

// Persistent kernel: one CTA per SM walks a contiguous range of work units (atom tile, row tile),
// atom-tile major, so every SM gets the same amount of work to within one row tile (no wave tail) and
// reloads the per-atom-tile phase columns only when the atom tile changes.
template <int APT, int G_ROWS_PER_WARP>
__global__ void __launch_bounds__(G_THREADS, 1) gatherKernel(GParams p, long long* __restrict__ forceFixed, long long* __restrict__ dedqFixed) {
    constexpr int BA = 32*APT;
    constexpr int G_ROW_TILE = G_WARPS*G_ROWS_PER_WARP;
    extern __shared__ __align__(128) unsigned char smem[];
    float4* Z4s = reinterpret_cast<float4*>(smem);                         // [Kz][BA]
    float2* Eys = reinterpret_cast<float2*>(smem + p.offEy);               // [Ky][BA]
    float4* coefS = reinterpret_cast<float4*>(smem + p.offCoef);           // [2][G_ROW_TILE][Kz]
    int2* infoS = reinterpret_cast<int2*>(smem + p.offInfo);               // [2][G_ROW_TILE]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int numRowTiles = p.numRowTiles;
    const long long totalUnits = (long long) p.numAtomTiles*numRowTiles;
    const int u0 = (int) (totalUnits*blockIdx.x/gridDim.x), u1 = (int) (totalUnits*(blockIdx.x + 1)/gridDim.x);
    if (u0 >= u1) return;
    const int tileElems = G_ROW_TILE*p.Kz;

    // Each warp only ever reads its own G_ROWS_PER_WARP rows of a coefficient tile, so every warp
    // prefetches (cp.async) and double-buffers its own rows: no block-level barrier in the main loop.
    const int warpElems = G_ROWS_PER_WARP*p.Kz;
    auto prefetchTile = [&](int unit, int buf) {
        const int rt = unit % numRowTiles;
        const size_t row0 = (size_t) p.signedLo + (size_t) rt*G_ROW_TILE + (size_t) warp*G_ROWS_PER_WARP;
        const float4* src = p.coef + row0*p.Kz;
        float4* dst = coefS + (size_t) buf*tileElems + (size_t) warp*warpElems;
        for (int e = lane; e < warpElems; e += 32) cpAsync16(dst + e, src + e);
        if (lane < G_ROWS_PER_WARP) infoS[buf*G_ROW_TILE + warp*G_ROWS_PER_WARP + lane] = p.rowInfo[row0 + lane];
        cpAsyncCommit();
    };

    float oD[APT], oX[APT], oY[APT], oZ[APT];
    float2 ex[APT];
    int curNx = -1, curAtomTile = -1, atom0 = 0;

    // cross-warp reduction through shared memory (aliases Z4s), then one fixed-point atomic per output
    auto flush = [&]() {
        __syncthreads();
        float4* red = reinterpret_cast<float4*>(smem);
        #pragma unroll
        for (int a = 0; a < APT; a++) red[warp*BA + lane + 32*a] = make_float4(oD[a], oX[a], oY[a], oZ[a]);
        __syncthreads();
        if (tid < BA) {
            float4 s = red[tid];
            #pragma unroll
            for (int w = 1; w < G_WARPS; w++) {
                const float4 v = red[w*BA + tid];
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
            const int atom = atom0 + tid;
            if (atom < p.N) {
                const double q = (double) p.qf[atom];
                atomicAddFixed(dedqFixed + atom, (double) s.x);
                atomicAddFixed(forceFixed + atom, q*(double) p.fx*(double) s.y);
                atomicAddFixed(forceFixed + p.Npad + atom, q*(double) p.fy*(double) s.z);
                atomicAddFixed(forceFixed + 2*(size_t) p.Npad + atom, q*(double) p.fz*(double) s.w);
            }
        }
        __syncthreads();
    };

    if (p.nbuf == 2) prefetchTile(u0, 0);
    for (int unit = u0; unit < u1; unit++) {
        const int buf = (p.nbuf == 2) ? ((unit - u0) & 1) : 0;
        const int atomTile = unit/numRowTiles, rowTile = unit - atomTile*numRowTiles;
        if (atomTile != curAtomTile) {
            if (curAtomTile >= 0) flush();
            else __syncthreads();
            curAtomTile = atomTile;
            atom0 = atomTile*BA;
            for (int e = tid; e < p.Kz*BA; e += G_THREADS) {
                const int l = e/BA, a = e - l*BA;
                Z4s[e] = p.colZ4[(size_t) l*p.Npad + atom0 + a];
            }
            for (int e = tid; e < p.Ky*BA; e += G_THREADS) {
                const int m = e/BA, a = e - m*BA;
                Eys[e] = p.colY[(size_t) m*p.Npad + atom0 + a];
            }
            #pragma unroll
            for (int a = 0; a < APT; a++) { oD[a] = 0.f; oX[a] = 0.f; oY[a] = 0.f; oZ[a] = 0.f; ex[a] = make_float2(0.f, 0.f); }
            curNx = -1;
            __syncthreads();                             // phase columns visible to every warp
        }
        if (p.nbuf == 2) {
            if (unit + 1 < u1) { prefetchTile(unit + 1, buf ^ 1); cpAsyncWait<1>(); }
            else cpAsyncWait<0>();
        }
        else { prefetchTile(unit, 0); cpAsyncWait<0>(); }
        __syncwarp();
        const float4* cT = coefS + (size_t) buf*tileElems + (size_t) warp*G_ROWS_PER_WARP*p.Kz;
        float2 U[G_ROWS_PER_WARP][APT], V[G_ROWS_PER_WARP][APT];
        #pragma unroll
        for (int i = 0; i < G_ROWS_PER_WARP; i++)
            #pragma unroll
            for (int a = 0; a < APT; a++) { U[i][a] = make_float2(0.f, 0.f); V[i][a] = make_float2(0.f, 0.f); }
        #pragma unroll 3
        for (int l = 0; l < p.Kz; l++) {
            float4 c[G_ROWS_PER_WARP], z[APT];
            #pragma unroll
            for (int i = 0; i < G_ROWS_PER_WARP; i++) c[i] = cT[i*p.Kz + l];
            #pragma unroll
            for (int a = 0; a < APT; a++) z[a] = Z4s[l*BA + lane + 32*a];
            #pragma unroll
            for (int i = 0; i < G_ROWS_PER_WARP; i++)
                #pragma unroll
                for (int a = 0; a < APT; a++) {
                    U[i][a].x = fmaf(c[i].x, z[a].x, U[i][a].x);  U[i][a].x = fmaf(c[i].z, z[a].y, U[i][a].x);
                    U[i][a].y = fmaf(c[i].y, z[a].x, U[i][a].y);  U[i][a].y = fmaf(c[i].w, z[a].y, U[i][a].y);
                    V[i][a].x = fmaf(c[i].w, z[a].z, V[i][a].x);  V[i][a].x = fmaf(-c[i].y, z[a].w, V[i][a].x);
                    V[i][a].y = fmaf(-c[i].z, z[a].z, V[i][a].y); V[i][a].y = fmaf(c[i].x, z[a].w, V[i][a].y);
                }
        }
        // epilogue: T = Ex(nx) Ey(ny); accumulate Re(T U), nx Im(T U), ny Im(T U), Im(T U')
        const int rowBase = p.signedLo + rowTile*G_ROW_TILE;
        #pragma unroll
        for (int i = 0; i < G_ROWS_PER_WARP; i++) {
            const int rloc = warp*G_ROWS_PER_WARP + i;
            if (rowBase + rloc >= p.signedHi) continue;
            const int2 info = infoS[buf*G_ROW_TILE + rloc];
            if (info.x != curNx) {
                curNx = info.x;
                #pragma unroll
                for (int a = 0; a < APT; a++) ex[a] = p.colX[(size_t) curNx*p.Npad + atom0 + lane + 32*a];
            }
            const int m = abs(info.y);
            const float sgn = info.y < 0 ? -1.f : 1.f;
            const float fnx = (float) info.x, fny = (float) info.y;
            #pragma unroll
            for (int a = 0; a < APT; a++) {
                float2 ey = Eys[m*BA + lane + 32*a];
                ey.y *= sgn;
                const float tr = ex[a].x*ey.x - ex[a].y*ey.y;
                const float ti = ex[a].x*ey.y + ex[a].y*ey.x;
                oD[a] = fmaf(tr, U[i][a].x, oD[a]);  oD[a] = fmaf(-ti, U[i][a].y, oD[a]);
                const float im = tr*U[i][a].y + ti*U[i][a].x;
                oX[a] = fmaf(fnx, im, oX[a]);
                oY[a] = fmaf(fny, im, oY[a]);
                oZ[a] = fmaf(tr, V[i][a].y, oZ[a]);  oZ[a] = fmaf(ti, V[i][a].x, oZ[a]);
            }
        }
        __syncwarp();
    }
    flush();
}

size_t gatherSmem(int Kx, int Ky, int Kz, int BA, int rowTile, int nbuf, size_t* offEy, size_t* offCoef, size_t* offInfo) {
    size_t z4 = (size_t) Kz*BA*sizeof(float4);
    size_t red = (size_t) G_WARPS*BA*sizeof(float4);
    size_t first = std::max(z4, red);
    first = (first + 127) & ~(size_t) 127;
    *offEy = first;
    size_t ey = ((size_t) Ky*BA*sizeof(float2) + 127) & ~(size_t) 127;
    *offCoef = first + ey;
    size_t coef = ((size_t) nbuf*rowTile*Kz*sizeof(float4) + 127) & ~(size_t) 127;
    *offInfo = *offCoef + coef;
    return *offInfo + 2*rowTile*sizeof(int2);
}

} // namespace

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
void planKSpace(State& st) {
    KSpacePlan& ks = st.ks;
    SGeom& f = ks.sF;
    const int Kx = ks.K[0], Ky = ks.K[1], Kz = ks.K[2];
    // S kernel: 2 rows per lane, TN in {6,7,8} columns per warp; pick the TN with the least column
    // padding (ties -> larger TN), at most S_MAX_WARPS column groups per CTA
    int bestPad = 1 << 30;
    for (int tn = 6; tn <= 8; tn++) {
        const int g = (Kz + tn - 1)/tn;
        const int pad = g*tn - Kz;
        if (pad <= bestPad) { bestPad = pad; f.TN = tn; f.NC = g; }
    }
    ks.tensorS = structureTensorEligible(st);
    ks.fp32S = f.NC <= S_MAX_WARPS;
    if (!ks.fp32S && !ks.tensorS) throw std::runtime_error("kmax along z too large for the structure-factor kernels (|nz| <= 64)");
    const int TNP = (f.TN + 1) & ~1;
    f.kzPad = f.NC*TNP;
    const int zOff = (Kx + Ky + 1) & ~1;
    f.rowPitch = zOff + f.kzPad;
    ks.numRows = Kx*Ky;
    // shard the unsigned rows over ranks (k-vector sharding)
    ks.rowLo = (int) ((int64_t) ks.numRows*st.shardRank/st.shardCount);
    ks.rowHi = (int) ((int64_t) ks.numRows*(st.shardRank + 1)/st.shardCount);
    const int rowsHere = ks.rowHi - ks.rowLo;
    f.threads = 32*f.NC;
    f.rowTiles = (std::max(rowsHere, 1) + S_BM - 1)/S_BM;
    f.stages = 3;
    const size_t stageBytes = (size_t) S_ATOMS_PER_STAGE*f.rowPitch*sizeof(float2);
    f.smem = 128 + f.stages*stageBytes;
    // atom splits: fill the resident CTA slots (shared memory allows 2-3 CTAs per SM)
    int numSM = 148;
    cudaDeviceGetAttribute(&numSM, cudaDevAttrMultiProcessorCount, st.device);
    const int ctasPerSM = std::max(1, std::min(4, (int) ((size_t) 220*1024/(f.smem + 1024))));
    const int slots = ctasPerSM*numSM;
    int splits = std::max(1, slots/std::max(1, f.rowTiles));
    // stages (of 32 atoms) per CTA at least. Row shards (multi-GPU) would otherwise be cut into hundreds of short CTAs
    // whose partial sums the coefficient kernel then has to add (8 ranks at 32k atoms: 0.176 -> 0.160 ms per rank); an
    // unsharded small box keeps short FP32 accumulation chains (the split sums are added in FP64: with one 648-atom
    // chain per thread the reciprocal energy of the 216-water boxes moves by 2e-7, too much for their 1e-6 of a
    // 25 kJ/mol total)
    int minStages = st.shardCount > 1 ? 16 : 4;
    if (const char* e = getenv("CFX_S_MIN_STAGES")) minStages = std::max(1, atoi(e));     // experiments
    const int maxSplits = std::max(1, st.Npad/(minStages*S_ATOMS_PER_STAGE));
    splits = std::min(splits, maxSplits);
    int aps = (st.Npad + splits - 1)/splits;
    aps = (aps + S_ATOMS_PER_STAGE - 1)/S_ATOMS_PER_STAGE*S_ATOMS_PER_STAGE;
    splits = (st.Npad + aps - 1)/aps;
    f.splits = splits;
    f.atomsPerSplit = aps;
    if (ks.tensorS) planStructureTensor(st);                     // fills ks.sT
    // signed rows
    std::vector<int> signedStart(ks.numRows + 1, 0);
    std::vector<int2> rowInfo;
    for (int row = 0; row < ks.numRows; row++) {
        const int nx = row/Ky, m = row % Ky;
        signedStart[row] = (int) rowInfo.size();
        rowInfo.push_back(make_int2(nx, m));
        if (nx > 0 && m > 0) rowInfo.push_back(make_int2(nx, -m));
    }
    signedStart[ks.numRows] = (int) rowInfo.size();
    ks.numSignedRows = (int) rowInfo.size();
    ks.signedLo = signedStart[ks.rowLo];
    ks.signedHi = signedStart[ks.rowHi];
    for (int k = 0; k < G_MAX_ROW_TILE; k++) rowInfo.push_back(make_int2(0, 0));   // padding rows (zero coefficients)

    const SGeom& t = ks.sT;
    const size_t maxPitch = std::max(f.rowPitch, ks.tensorS ? t.rowPitch : 0);
    CFX_CUDA(cudaMalloc(&st.rowS, (size_t) st.Npad*maxPitch*sizeof(float2)));
    CFX_CUDA(cudaMalloc(&st.colX, (size_t) Kx*st.Npad*sizeof(float2)));
    CFX_CUDA(cudaMalloc(&st.colY, (size_t) Ky*st.Npad*sizeof(float2)));
    CFX_CUDA(cudaMalloc(&st.colZ4, (size_t) Kz*st.Npad*sizeof(float4)));
    const size_t partSlots = std::max((size_t) f.splits*f.kzPad, ks.tensorS ? (size_t) t.splits*t.kzPad : (size_t) 0);
    CFX_CUDA(cudaMalloc(&st.sPart, partSlots*ks.numRows*8*sizeof(float)));
    const size_t coefElems = (size_t) (ks.numSignedRows + G_MAX_ROW_TILE)*Kz;
    CFX_CUDA(cudaMalloc(&st.gCoef, coefElems*sizeof(float4)));
    CFX_CUDA(cudaMemset(st.gCoef, 0, coefElems*sizeof(float4)));
    CFX_CUDA(cudaMalloc(&st.gRowInfo, rowInfo.size()*sizeof(int2)));
    CFX_CUDA(cudaMemcpy(st.gRowInfo, rowInfo.data(), rowInfo.size()*sizeof(int2), cudaMemcpyHostToDevice));
    int* dSigned = nullptr;
    CFX_CUDA(cudaMalloc(&dSigned, signedStart.size()*sizeof(int)));
    CFX_CUDA(cudaMemcpy(dSigned, signedStart.data(), signedStart.size()*sizeof(int), cudaMemcpyHostToDevice));
    st.ks_signedStart = dSigned;

    // gather geometry: prefer 2 rows x 8 atoms per thread (fastest inner loop), fall back to 4 x 4 and to
    // single-buffered coefficient tiles when the per-atom-tile phase columns do not fit in shared memory
    const size_t smemCap = 220*1024;
    size_t o1, o2, o3;
    const int shapes[4][3] = {{8, 2, 2}, {4, 4, 2}, {4, 4, 1}, {4, 2, 1}};       // {atoms per lane, rows per warp, buffers}
    ks.gAtoms = 0;
    for (const auto& sh : shapes) {
        const size_t need = gatherSmem(Kx, Ky, Kz, 32*sh[0], G_WARPS*sh[1], sh[2], &o1, &o2, &o3);
        if (need <= smemCap) { ks.gAtoms = 32*sh[0]; ks.gRowsPerWarp = sh[1]; ks.gBuffers = sh[2]; ks.gSmem = need; break; }
    }
    if (ks.gAtoms == 0) throw std::runtime_error("kmax too large for the gather kernel's shared-memory tiles");
    const int signedHere = ks.signedHi - ks.signedLo;
    const int gRowTile = G_WARPS*ks.gRowsPerWarp;
    ks.gRowsPerTile = (std::max(signedHere, 1) + gRowTile - 1)/gRowTile;          // row tiles per atom tile
    ks.gRowSplits = numSM;                                                        // persistent grid size

    // (function attributes are per process, not per handle: always the hardware maximum)
    CFX_CUDA(cudaFuncSetAttribute(phaseTableKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
    CFX_CUDA(cudaFuncSetAttribute(structureFactorKernel<6, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
    CFX_CUDA(cudaFuncSetAttribute(structureFactorKernel<7, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
    CFX_CUDA(cudaFuncSetAttribute(structureFactorKernel<8, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
    CFX_CUDA(cudaFuncSetAttribute(structureFactorKernel<6, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
    CFX_CUDA(cudaFuncSetAttribute(structureFactorKernel<7, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
    CFX_CUDA(cudaFuncSetAttribute(structureFactorKernel<8, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
    CFX_CUDA(cudaFuncSetAttribute(gatherKernel<8, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
    CFX_CUDA(cudaFuncSetAttribute(gatherKernel<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
    CFX_CUDA(cudaFuncSetAttribute(gatherKernel<4, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227*1024));
    planKSpaceTensor(st);
}

void launchKSpace(State& st, const double* dPos, bool forces, bool energy, long long* dForce, long long* dDedq, cudaStream_t s) {
    KSpacePlan& ks = st.ks;
    if (!forces && !energy) return;
    // the reciprocal energy needs round-to-nearest sums: FP32 kernel whenever it is requested
    // The integer tensor-core kernel sums exactly (no truncation bias, order-independent), so it serves the energy call too
    // (with a fourth digit plane: 31-bit operands). The TF32 kernel (used when the integer one is disabled) truncates on
    // accumulation and never serves energies.
    const bool useTensorS = ks.tensorS && (!energy || !ks.fp32S || ks.i8S);
    const SGeom& g = useTensorS ? ks.sT : ks.sF;
    const int Kx = ks.K[0], Ky = ks.K[1], Kz = ks.K[2];
    const int zOff = g.rowPitch - g.kzPad;
    const int TNP = (g.TN + 1) & ~1;
    TableParams tp{st.N, st.Npad, Kx, Ky, Kz, g.kzPad, zOff, g.rowPitch, g.TN, TNP, 1.0/st.box.L[0], 1.0/st.box.L[1], 1.0/st.box.L[2],
                   (ks.tensorGather && forces) ? st.zSplit : nullptr, ks.tKC};
    phaseTableKernel<<<st.Npad/PT_ATOMS, 128, PT_ATOMS*g.rowPitch*sizeof(float2), s>>>(tp, dPos, st.qf, st.rowS, st.colX, st.colY, st.colZ4);
    CFX_LAUNCH_CHECK(); st.launches++;
    mark(st, "phase_tables", s);
    if (ks.rowHi <= ks.rowLo) return;

    SParams sp;
    sp.rowS = st.rowS; sp.part = st.sPart;
    sp.rowPitch = g.rowPitch; sp.Kx = Kx; sp.Ky = Ky; sp.zOff = zOff; sp.kzPad = g.kzPad;
    sp.stages = g.stages;
    sp.rowLo = ks.rowLo; sp.rowHi = ks.rowHi; sp.numRows = ks.numRows;
    sp.atomsPerSplit = g.atomsPerSplit; sp.Npad = st.Npad;
    const dim3 sGrid(g.rowTiles, g.splits);
    if (useTensorS) launchStructureTensor(st, energy, s);
    else if (g.NC <= 4) {
        if (g.TN == 6)      structureFactorKernel<6, 4><<<sGrid, g.threads, g.smem, s>>>(sp);
        else if (g.TN == 7) structureFactorKernel<7, 4><<<sGrid, g.threads, g.smem, s>>>(sp);
        else                  structureFactorKernel<8, 4><<<sGrid, g.threads, g.smem, s>>>(sp);
    }
    else {
        if (g.TN == 6)      structureFactorKernel<6, 8><<<sGrid, g.threads, g.smem, s>>>(sp);
        else if (g.TN == 7) structureFactorKernel<7, 8><<<sGrid, g.threads, g.smem, s>>>(sp);
        else                  structureFactorKernel<8, 8><<<sGrid, g.threads, g.smem, s>>>(sp);
    }
    if (!useTensorS) { CFX_LAUNCH_CHECK(); st.launches++; }
    mark(st, "structure_factor", s);

    CoefParams cp;
    cp.part = st.sPart; cp.coef = st.gCoef; cp.signedStart = st.ks_signedStart;
    cp.Kx = Kx; cp.Ky = Ky; cp.Kz = Kz; cp.kzPad = g.kzPad; cp.numRows = ks.numRows; cp.splits = g.splits;
    cp.rowLo = ks.rowLo; cp.rowHi = ks.rowHi; cp.TN = g.TN; cp.TNP = TNP;
    cp.gx = 2*M_PI/st.box.L[0]; cp.gy = 2*M_PI/st.box.L[1]; cp.gz = 2*M_PI/st.box.L[2];
    cp.C = 4.0/st.box.L[0]/st.box.L[1]/st.box.L[2]*M_PI*CFX_ONE_4PI_EPS0;
    cp.invFourAlpha2 = 0.25/(st.alpha*st.alpha);
    cp.energy = energy; cp.forces = forces; cp.swapPairs = !useTensorS;
    cp.coefT = ks.tensorGather ? st.coefT : nullptr; cp.KC = ks.tKC; cp.NT = ks.tNT; cp.signedLo = ks.signedLo;
    const int items = (ks.rowHi - ks.rowLo)*Kz;
    coefficientKernel<<<(4*items + 127)/128, 128, 0, s>>>(cp, st.energyFixed);
    CFX_LAUNCH_CHECK(); st.launches++;
    mark(st, "kspace_coef", s);

    if (forces && ks.signedHi > ks.signedLo && ks.tensorGather) {
        launchGatherTensor(st, dForce, dDedq, s);
        mark(st, "kspace_gather", s);
    }
    else if (forces && ks.signedHi > ks.signedLo) {
        GParams gp;
        gp.coef = st.gCoef; gp.rowInfo = st.gRowInfo; gp.colX = st.colX; gp.colY = st.colY; gp.colZ4 = st.colZ4; gp.qf = st.qf;
        gp.Kx = Kx; gp.Ky = Ky; gp.Kz = Kz; gp.N = st.N; gp.Npad = st.Npad;
        gp.signedLo = ks.signedLo; gp.signedHi = ks.signedHi; gp.numRowTiles = ks.gRowsPerTile; gp.numAtomTiles = st.Npad/ks.gAtoms;
        gp.fx = (float) cp.gx; gp.fy = (float) cp.gy; gp.fz = (float) cp.gz;
        gatherSmem(Kx, Ky, Kz, ks.gAtoms, G_WARPS*ks.gRowsPerWarp, ks.gBuffers, &gp.offEy, &gp.offCoef, &gp.offInfo);
        gp.nbuf = ks.gBuffers;
        const long long units = (long long) gp.numAtomTiles*gp.numRowTiles;
        const int gGrid = (int) std::min<long long>(ks.gRowSplits, units);
        if (ks.gAtoms == 256)           gatherKernel<8, 2><<<gGrid, G_THREADS, ks.gSmem, s>>>(gp, dForce, dDedq);
        else if (ks.gRowsPerWarp == 4)  gatherKernel<4, 4><<<gGrid, G_THREADS, ks.gSmem, s>>>(gp, dForce, dDedq);
        else                            gatherKernel<4, 2><<<gGrid, G_THREADS, ks.gSmem, s>>>(gp, dForce, dDedq);
        CFX_LAUNCH_CHECK(); st.launches++;
        mark(st, "kspace_gather", s);
    }
}

} // namespace cfx

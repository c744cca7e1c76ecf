// direct.cu -- piece (2): Ewald direct space (ReferenceCoulKernels.cpp:559-593) on a GPU-built
// cell list.
//
// The reference rebuilds a neighbour list every call (all i<j, not excluded, min-image r2 <= rc2) and
// evaluates erfc Coulomb + Lennard-Jones on it in FP64. Here:
//
//  cell build   atoms are wrapped (FP64), binned into cells of edge >= rc/2 and sorted by cell with a
//               deterministic order inside each cell; each sorted atom keeps its position as a FP32
//               offset inside its own cell, so pair separations are formed from small numbers
//               (cell-relative coordinates: no loss of precision in large boxes).
//  pair kernel  one warp per i-tile of 8 consecutive sorted atoms, 4 lanes per i atom. Candidate j
//               atoms from the 5x5x5 cell stencil are pruned against the i-tile's bounding box and
//               compacted (ballot) into a 32-entry j-tile in shared memory; each lane then walks its
//               quarter of the j-tile. Forces and dE/dq are accumulated together in registers and
//               reduced over the 4 lanes of an i atom with warp shuffles. Every ordered pair is
//               evaluated once from each side (full shell): no j-side atomics.
//  exactness    the in-cutoff predicate of a pair whose FP32 r2 falls within 1e-5 of rc2 is re-evaluated
//               in FP64 on the original coordinates with the reference's operation order, so the
//               neighbour set is bit-exact. Exclusions are looked up in the per-atom CSR.
//
// FP32 pair arithmetic for forces and dE/dq (erfcf/expf), int64 fixed-point accumulation. Pair ENERGIES
// are evaluated in FP64: in-cutoff i<j pairs are ballot-compacted into a per-warp queue and evaluated 32
// at a time, because E_direct cancels against E_self + E_excl to a small fraction of its size.
#include "cfx_internal.cuh"

#include <algorithm>
#include <cstdlib>
#include <cmath>

namespace cfx {

namespace {

#define CELL_BITS 10
#define CELL_MASK 1023

struct CellParams {
    int N, ncx, ncy, ncz, ncells;
    double invLx, invLy, invLz;
    double csx, csy, csz;
};

// wrap, bin, cell-local coordinates; counts per cell
__global__ void __launch_bounds__(256) cellAssignKernel(CellParams p, const double* __restrict__ pos, const float* __restrict__ qf,
        int* __restrict__ cellOfAtom, float4* __restrict__ userLocal, int* __restrict__ cellCount) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    double u[3] = {pos[3*(size_t) i]*p.invLx, pos[3*(size_t) i + 1]*p.invLy, pos[3*(size_t) i + 2]*p.invLz};
    const int nc[3] = {p.ncx, p.ncy, p.ncz};
    const double cs[3] = {p.csx, p.csy, p.csz};
    int c[3];
    float loc[3];
    #pragma unroll
    for (int d = 0; d < 3; d++) {
        double f = u[d] - floor(u[d]);          // [0,1)
        double g = f*nc[d];
        int k = (int) g;
        if (k >= nc[d]) k = nc[d] - 1;
        c[d] = k;
        loc[d] = (float) ((g - k)*cs[d]);
    }
    const int cell = (c[0]*p.ncy + c[1])*p.ncz + c[2];
    cellOfAtom[i] = cell;
    userLocal[i] = make_float4(loc[0], loc[1], loc[2], qf[i]);
    atomicAdd(cellCount + cell, 1);
}

// exclusive scan of cellCount -> cellStart (single CTA; ncells is at most a few 10^4)
__global__ void __launch_bounds__(1024) cellScanKernel(int ncells, const int* __restrict__ cellCount, int* __restrict__ cellStart,
        int* __restrict__ cellFill) {
    __shared__ int warpTotals[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < ncells; base += 1024) {
        const int idx = base + threadIdx.x;
        const int v = idx < ncells ? cellCount[idx] : 0;
        int incl = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) warpTotals[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warpTotals[lane];
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += n;
            }
            warpTotals[lane] = w;
        }
        __syncthreads();
        const int before = carry + (warp > 0 ? warpTotals[warp-1] : 0) + incl - v;
        if (idx < ncells) { cellStart[idx] = before; cellFill[idx] = before; }
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) cellStart[ncells] = carry;
}

__global__ void __launch_bounds__(256) cellFillKernel(int N, const int* __restrict__ cellOfAtom, int* __restrict__ cellFill,
        int* __restrict__ sortedUser) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int slot = atomicAdd(cellFill + cellOfAtom[i], 1);
    sortedUser[slot] = i;
}

// One thread per slot of the (arbitrarily ordered) cell fill: the atom's final slot is the start of
// its cell plus its rank among the cell's atoms by user index, so the sorted order -- and with it every
// FP32 accumulation order downstream -- is deterministic. Writes all sorted arrays in one pass.
__global__ void __launch_bounds__(256) cellRankGatherKernel(int N, int ncy, int ncz, const int* __restrict__ filledUser,
        const int* __restrict__ cellOfAtom, const int* __restrict__ cellStart, const float4* __restrict__ userLocal,
        const float2* __restrict__ lj, int* __restrict__ sortedUser, float4* __restrict__ sortedLocal,
        int* __restrict__ sortedCell, float2* __restrict__ sortedLJ) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= N) return;
    const int u = filledUser[s];
    const int cell = cellOfAtom[u];
    const int s0 = cellStart[cell], s1 = cellStart[cell+1];
    int rank = 0;
    for (int t = s0; t < s1; t++) rank += (filledUser[t] < u) ? 1 : 0;
    const int dst = s0 + rank;
    const int cz = cell % ncz, cy = (cell/ncz) % ncy, cx = cell/(ncz*ncy);
    sortedUser[dst] = u;
    sortedLocal[dst] = userLocal[u];
    sortedCell[dst] = cx | (cy << CELL_BITS) | (cz << (2*CELL_BITS));
    sortedLJ[dst] = lj[u];
}

// ------------------------------------------------------------------------------------------------
// pair kernel
// ------------------------------------------------------------------------------------------------
#define P_WARPS 4
#define P_ITILE 8
#define P_JTILE 32
#define P_JCAP 64

struct PairParams {
    int N, Npad, numGroups, groupLo, groupHi;
    int jSplits;                 // the (x,y) columns of an i-tile's stencil are dealt over gridDim.y CTAs (small shards)
    int ncx, ncy, ncz;
    float csx, csy, csz;
    float Lx, Ly, Lz, invLx, invLy, invLz;
    double dLx, dLy, dLz, rc2d;
    float rc2, alpha, band;
    const float4* sortedLocal; const int* sortedCell; const float2* sortedLJ; const int* sortedUser;
    const int* cellStart;
    const int* exclPtr; const int* exclCols;
    const double* pos; const double* q; const double2* ljd; double alphaD, dInvLx, dInvLy, dInvLz;
    long long* forceFixed; long long* dedqFixed; long long* energyFixed;
    unsigned long long* counters; int2* pairBuffer; unsigned long long pairCapacity;
};

__device__ __forceinline__ int wrapNearest(int d, int nc) {
    const int half = (nc - 1) >> 1;
    if (d > half) d -= nc;
    if (d < -(nc >> 1)) d += nc;
    return d;
}

// erfc(x), x >= 0, relative error ~1.2e-7 (Chebyshev fit in t = 1/(1 + x/2), Numerical Recipes erfcc):
// one MUFU.RCP, nine FFMA and one MUFU.EX2 instead of the ~40-instruction erfcf().
__device__ __forceinline__ float erfcFast(float x, float x2) {
    const float t = __fdividef(1.f, fmaf(0.5f, x, 1.f));
    float p = fmaf(t, 0.17087277f, -0.82215223f);
    p = fmaf(t, p, 1.48851587f);
    p = fmaf(t, p, -1.13520398f);
    p = fmaf(t, p, 0.27886807f);
    p = fmaf(t, p, -0.18628806f);
    p = fmaf(t, p, 0.09678418f);
    p = fmaf(t, p, 0.37409196f);
    p = fmaf(t, p, 1.00002368f);
    p = fmaf(t, p, -1.26551223f);
    return t*__expf(p - x2);
}

// exact FP64 predicate with the reference's operation order (J - I with I the lower user index;
// z, y, x floor-based wrap; left-to-right sum of squares; no FMA contraction)
__device__ __noinline__ bool exactInCutoff(const double* __restrict__ pos, int ua, int ub, double Lx, double Ly, double Lz, double rc2) {
    const int I = min(ua, ub), J = max(ua, ub);
    double dx = __dsub_rn(pos[3*(size_t) J], pos[3*(size_t) I]);
    double dy = __dsub_rn(pos[3*(size_t) J + 1], pos[3*(size_t) I + 1]);
    double dz = __dsub_rn(pos[3*(size_t) J + 2], pos[3*(size_t) I + 2]);
    dz = __dsub_rn(dz, __dmul_rn(Lz, floor(__dadd_rn(__ddiv_rn(dz, Lz), 0.5))));
    dy = __dsub_rn(dy, __dmul_rn(Ly, floor(__dadd_rn(__ddiv_rn(dy, Ly), 0.5))));
    dx = __dsub_rn(dx, __dmul_rn(Lx, floor(__dadd_rn(__ddiv_rn(dx, Lx), 0.5))));
    const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return r2 <= rc2;
}

__device__ __forceinline__ int modPos(int v, int n) { v %= n; return v < 0 ? v + n : v; }

// EMODE: 0 = no pair energy, 1 = FP32 pair terms (the partial energy the reference returns when
// includeEnergy is false is discarded by OpenMM; it is still produced, at FP32 accuracy), 2 = FP64 terms.
template <bool FORCES, int EMODE, bool EMIT>
__global__ void __launch_bounds__(P_WARPS*32) pairKernel(PairParams p) {
    __shared__ float4 sPos[P_WARPS][P_JCAP];
    __shared__ float2 sLJ[P_WARPS][P_JCAP];
    __shared__ int sUser[P_WARPS][P_JCAP];
    __shared__ int2 sEq[P_WARPS][P_JCAP];          // queue of in-cutoff (ui<uj) pairs awaiting the FP64 energy
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = p.groupLo + blockIdx.x*P_WARPS + warp;
    if (g >= p.groupHi) return;                      // whole warp exits; only __syncwarp below
    float4* tPos = sPos[warp]; float2* tLJ = sLJ[warp]; int* tUser = sUser[warp]; int2* eq = sEq[warp];

    const int i0 = g*P_ITILE;
    const int ni = min(P_ITILE, p.N - i0);
    const int ii = lane >> 2, part = lane & 3;
    const bool validI = ii < ni;
    const int iIdx = i0 + (validI ? ii : 0);
    const int c0 = p.sortedCell[i0];
    const int c0x = c0 & CELL_MASK, c0y = (c0 >> CELL_BITS) & CELL_MASK, c0z = c0 >> (2*CELL_BITS);

    // my i atom, expressed in the frame of cell c0
    const float4 li = p.sortedLocal[iIdx];
    const int ci = p.sortedCell[iIdx];
    const int ox = wrapNearest((ci & CELL_MASK) - c0x, p.ncx);
    const int oy = wrapNearest(((ci >> CELL_BITS) & CELL_MASK) - c0y, p.ncy);
    const int oz = wrapNearest((ci >> (2*CELL_BITS)) - c0z, p.ncz);
    const float pix = li.x + ox*p.csx, piy = li.y + oy*p.csy, piz = li.z + oz*p.csz;
    const float2 lji = p.sortedLJ[iIdx];
    const int ui = p.sortedUser[iIdx];
    const float keqi = (float) CFX_ONE_4PI_EPS0*li.w;
    const int exBeg = p.exclPtr[ui], exEnd = p.exclPtr[ui+1];
    int exLo = 0x7fffffff, exHi = -1;
    if (exEnd > exBeg) { exLo = p.exclCols[exBeg]; exHi = p.exclCols[exEnd-1]; }    // CSR columns are sorted

    // bounding box of the i-tile and its cell-offset range (warp reductions)
    float bminx = pix, bminy = piy, bminz = piz, bmaxx = pix, bmaxy = piy, bmaxz = piz;
    int ominx = ox, ominy = oy, ominz = oz, omaxx = ox, omaxy = oy, omaxz = oz;
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bminx = fminf(bminx, __shfl_xor_sync(0xffffffffu, bminx, o)); bmaxx = fmaxf(bmaxx, __shfl_xor_sync(0xffffffffu, bmaxx, o));
        bminy = fminf(bminy, __shfl_xor_sync(0xffffffffu, bminy, o)); bmaxy = fmaxf(bmaxy, __shfl_xor_sync(0xffffffffu, bmaxy, o));
        bminz = fminf(bminz, __shfl_xor_sync(0xffffffffu, bminz, o)); bmaxz = fmaxf(bmaxz, __shfl_xor_sync(0xffffffffu, bmaxz, o));
        ominx = min(ominx, __shfl_xor_sync(0xffffffffu, ominx, o)); omaxx = max(omaxx, __shfl_xor_sync(0xffffffffu, omaxx, o));
        ominy = min(ominy, __shfl_xor_sync(0xffffffffu, ominy, o)); omaxy = max(omaxy, __shfl_xor_sync(0xffffffffu, omaxy, o));
        ominz = min(ominz, __shfl_xor_sync(0xffffffffu, ominz, o)); omaxz = max(omaxz, __shfl_xor_sync(0xffffffffu, omaxz, o));
    }
    // stencil: offsets [omin-2, omax+2] per axis. If that range would wrap onto itself (tiny box, or an
    // i-tile straddling a large empty region) it is truncated to all nc cells of the axis and the
    // separation of every pair is min-imaged instead (warp-uniform flag).
    const int loX = ominx - 2, nX = min(omaxx - ominx + 5, p.ncx);
    const int loY = ominy - 2, nY = min(omaxy - ominy + 5, p.ncy);
    const int loZ = ominz - 2, nZ = min(omaxz - ominz + 5, p.ncz);
    const bool minImage = (omaxx - ominx + 5 > p.ncx) || (omaxy - ominy + 5 > p.ncy) || (omaxz - ominz + 5 > p.ncz);

    float fx = 0.f, fy = 0.f, fz = 0.f, dq = 0.f, enf = 0.f;
    double en = 0.0;
    unsigned int nPairs = 0, nCand = 0;
    int count = 0;                                   // entries waiting in the j-tile
    int qCount = 0;                                  // entries waiting in the energy queue

    // Pair energies are evaluated in FP64 on the original coordinates (the direct, self and exclusion
    // sums cancel to a small fraction of their size, FP32 terms would cost ~1e-3 kJ/mol). In-cutoff
    // pairs are compacted into a queue so that all 32 lanes do FP64 work together.
    auto energyBatch = [&](int n) {
        __syncwarp();
        if (lane < n) {
            const int a = eq[lane].x, b = eq[lane].y;
            double dx = p.pos[3*(size_t) a] - p.pos[3*(size_t) b];
            double dy = p.pos[3*(size_t) a + 1] - p.pos[3*(size_t) b + 1];
            double dz = p.pos[3*(size_t) a + 2] - p.pos[3*(size_t) b + 2];
            dx -= p.dLx*floor(dx*p.dInvLx + 0.5); dy -= p.dLy*floor(dy*p.dInvLy + 0.5); dz -= p.dLz*floor(dz*p.dInvLz + 0.5);
            const double r2 = dx*dx + dy*dy + dz*dz;
            const double invR = rsqrt(r2);
            const double ar = p.alphaD*r2*invR;
            const double2 la = p.ljd[a], lb = p.ljd[b];
            const double sig = la.x + lb.x;
            double s2 = sig*invR; s2 *= s2;
            const double s6 = s2*s2*s2;
            en += CFX_ONE_4PI_EPS0*p.q[a]*p.q[b]*invR*erfc(ar) + s6*(la.y*lb.y)*(s6 - 1.0);
        }
        __syncwarp();
    };

    auto processTile = [&](int n) {
        __syncwarp();
        #pragma unroll 2
        for (int k = part; k < P_JTILE; k += 4) {
            const float4 pj = tPos[k];
            const int uj = tUser[k];
            float dx = pix - pj.x, dy = piy - pj.y, dz = piz - pj.z;      // pos[i] - pos[j]
            if (minImage) {
                dx -= p.Lx*rintf(dx*p.invLx); dy -= p.Ly*rintf(dy*p.invLy); dz -= p.Lz*rintf(dz*p.invLz);
            }
            const float r2 = dx*dx + dy*dy + dz*dz;
            bool in = validI && (k < n) && (uj != ui) && (r2 <= p.rc2);
            if (validI && (k < n) && (uj != ui) && fabsf(r2 - p.rc2) < p.band)
                in = exactInCutoff(p.pos, ui, uj, p.dLx, p.dLy, p.dLz, p.rc2d);
            if (in && uj >= exLo && uj <= exHi)
                for (int e = exBeg; e < exEnd; e++)
                    if (p.exclCols[e] == uj) { in = false; break; }
            if (in) {
                if (FORCES || EMODE == 1) {
                    const float2 ljj = tLJ[k];
                    const float invR = rsqrtf(r2);
                    const float r = r2*invR;
                    const float ar = p.alpha*r;
                    const float ar2 = ar*ar;
                    const float erfcv = erfcFast(ar, ar2);
                    const float coul = keqi*pj.w*invR;
                    const float sig = lji.x + ljj.x;
                    float s2 = sig*invR; s2 *= s2;
                    const float s6 = s2*s2*s2;
                    const float es6 = s6*(lji.y*ljj.y);
                    if (FORCES) {
                        const float ex = __expf(-ar2);
                        const float invR2 = invR*invR;
                        const float dEdR = (coul*(erfcv + ar*ex*1.1283791671f) + es6*(12.f*s6 - 6.f))*invR2;
                        fx = fmaf(dEdR, dx, fx); fy = fmaf(dEdR, dy, fy); fz = fmaf(dEdR, dz, fz);
                        dq = fmaf((float) CFX_ONE_4PI_EPS0*pj.w*invR, erfcv, dq);
                    }
                    if (EMODE == 1) enf += coul*erfcv + es6*(s6 - 1.f);      // discarded partial energy: FP32 per lane
                }
                if (ui < uj) {
                    nPairs++;
                    if (EMIT) {
                        const unsigned long long slot = atomicAdd(p.counters + 2, 1ull);
                        if (slot < p.pairCapacity) p.pairBuffer[slot] = make_int2(ui, uj);
                    }
                }
            }
            if (EMODE == 2) {
                const bool want = in && ui < uj;
                const unsigned int m = __ballot_sync(0xffffffffu, want);
                if (want) eq[qCount + __popc(m & ((1u << lane) - 1u))] = make_int2(ui, uj);
                qCount += __popc(m);
                if (qCount >= 32) {
                    energyBatch(32);
                    const int rest = qCount - 32;
                    int2 mv = make_int2(0, 0);
                    if (lane < rest) mv = eq[32 + lane];
                    __syncwarp();
                    if (lane < rest) eq[lane] = mv;
                    qCount = rest;
                }
            }
        }
        nCand += (unsigned int) min(n, P_JTILE);
        __syncwarp();
    };

    // z run in unwrapped cell units [zlo, zlo+nZ-1] -> at most two contiguous wrapped segments
    const int zlo = c0z + loZ;
    const int zloW = modPos(zlo, p.ncz);
    int segLo[2], segHi[2], segShift[2], nSeg = 1;
    segLo[0] = zloW; segShift[0] = zlo - zloW;
    if (zloW + nZ - 1 < p.ncz) segHi[0] = zloW + nZ - 1;
    else { segHi[0] = p.ncz - 1; segLo[1] = 0; segHi[1] = zloW + nZ - 1 - p.ncz; segShift[1] = zlo - zloW + p.ncz; nSeg = 2; }

    for (int ax = 0; ax < nX; ax++) {
        const int offx = loX + ax;
        const int cx = modPos(c0x + offx, p.ncx);
        for (int ay = 0; ay < nY; ay++) {
            if (p.jSplits > 1 && (ax*nY + ay) % p.jSplits != (int) blockIdx.y) continue;
            const int offy = loY + ay;
            const int cy = modPos(c0y + offy, p.ncy);
            const int rowCell = (cx*p.ncy + cy)*p.ncz;
            const float shx = offx*p.csx, shy = offy*p.csy;
            for (int sg = 0; sg < nSeg; sg++) {
                const int s0 = p.cellStart[rowCell + segLo[sg]], s1 = p.cellStart[rowCell + segHi[sg] + 1];
                const int zShift = segShift[sg] - c0z;
                for (int base = s0; base < s1; base += 32) {
                    const int s = base + lane;
                    bool pass = false;
                    float4 pj = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (s < s1) {
                        const float4 l4 = p.sortedLocal[s];
                        const int cz = p.sortedCell[s] >> (2*CELL_BITS);
                        pj = make_float4(l4.x + shx, l4.y + shy, l4.z + (cz + zShift)*p.csz, l4.w);
                        if (minImage) pass = true;
                        else {
                            const float ex = fmaxf(0.f, fmaxf(bminx - pj.x, pj.x - bmaxx));
                            const float ey = fmaxf(0.f, fmaxf(bminy - pj.y, pj.y - bmaxy));
                            const float ez = fmaxf(0.f, fmaxf(bminz - pj.z, pj.z - bmaxz));
                            pass = ex*ex + ey*ey + ez*ez <= p.rc2*1.0001f;
                        }
                    }
                    const unsigned int m = __ballot_sync(0xffffffffu, pass);
                    if (pass) {
                        const int slot = count + __popc(m & ((1u << lane) - 1u));
                        tPos[slot] = pj;
                        tLJ[slot] = p.sortedLJ[s];
                        tUser[slot] = p.sortedUser[s];
                    }
                    count += __popc(m);
                    if (count >= P_JTILE) {
                        processTile(P_JTILE);
                        // move the overflow (< 32 entries) to the front
                        const int rest = count - P_JTILE;
                        float4 a; float2 b; int c;
                        if (lane < rest) { a = tPos[P_JTILE + lane]; b = tLJ[P_JTILE + lane]; c = tUser[P_JTILE + lane]; }
                        __syncwarp();
                        if (lane < rest) { tPos[lane] = a; tLJ[lane] = b; tUser[lane] = c; }
                        count = rest;
                    }
                }
            }
        }
    }
    if (count > 0) processTile(count);
    if (EMODE == 2 && qCount > 0) energyBatch(qCount);

    // reduce over the 4 lanes of each i atom (warp shuffles), one fixed-point atomic per output
    if (FORCES) {
        #pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {
            fx += __shfl_xor_sync(0xffffffffu, fx, o); fy += __shfl_xor_sync(0xffffffffu, fy, o);
            fz += __shfl_xor_sync(0xffffffffu, fz, o); dq += __shfl_xor_sync(0xffffffffu, dq, o);
        }
        if (validI && part == 0) {
            atomicAddFixed(p.forceFixed + ui, (double) fx);
            atomicAddFixed(p.forceFixed + p.Npad + ui, (double) fy);
            atomicAddFixed(p.forceFixed + 2*(size_t) p.Npad + ui, (double) fz);
            atomicAddFixed(p.dedqFixed + ui, (double) dq);
        }
    }
    if (EMODE != 0) {
        if (EMODE == 1) en = (double) enf;
        en = warpSum(en);
        // FP64 queue: each i<j pair once. FP32 terms: every pair is seen from both sides.
        if (lane == 0) atomicAddEnergy(p.energyFixed + CFX_E_DIRECT, EMODE == 2 ? en : 0.5*en);
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) nPairs += __shfl_xor_sync(0xffffffffu, nPairs, o);
    if (lane == 0) {
        atomicAdd(p.counters + 0, (unsigned long long) nPairs);
        atomicAdd(p.counters + 1, (unsigned long long) nCand*P_ITILE);
    }
}

template <bool EMIT>
void dispatchPair(const PairParams& pp, bool forces, int emode, dim3 blocks, cudaStream_t s) {
    const int t = P_WARPS*32;
    if (forces) {
        if (emode == 2)      pairKernel<true, 2, EMIT><<<blocks, t, 0, s>>>(pp);
        else if (emode == 1) pairKernel<true, 1, EMIT><<<blocks, t, 0, s>>>(pp);
        else                 pairKernel<true, 0, EMIT><<<blocks, t, 0, s>>>(pp);
    }
    else {
        if (emode == 2)      pairKernel<false, 2, EMIT><<<blocks, t, 0, s>>>(pp);
        else if (emode == 1) pairKernel<false, 1, EMIT><<<blocks, t, 0, s>>>(pp);
        else                 pairKernel<false, 0, EMIT><<<blocks, t, 0, s>>>(pp);
    }
}

} // namespace

void planCells(State& st) {
    CellPlan& c = st.cells;
    c.smallBox = false;
    c.ncells = 1;
    for (int d = 0; d < 3; d++) {
        int n = (int) floor(st.box.L[d]/(0.5*st.cutoff));
        n = std::max(1, std::min(n, CELL_MASK));
        c.nc[d] = n;
        c.csd[d] = st.box.L[d]/n;
        c.cs[d] = (float) c.csd[d];
        c.ncells *= n;
        if (n < 7) c.smallBox = true;      // informational: tiles will fall back to per-pair min image
    }
    CFX_CUDA(cudaMalloc(&st.cellOfAtom, sizeof(int)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.cellCount, sizeof(int)*(c.ncells + 1)));
    CFX_CUDA(cudaMalloc(&st.cellStart, sizeof(int)*(c.ncells + 1)));
    CFX_CUDA(cudaMalloc(&st.cellFill, sizeof(int)*(c.ncells + 1)));
    CFX_CUDA(cudaMalloc(&st.userLocal, sizeof(float4)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.sortedLocal, sizeof(float4)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.sortedCell, sizeof(int)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.sortedLJ, sizeof(float2)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.sortedUser, sizeof(int)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.filledUser, sizeof(int)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.pairCounters, sizeof(unsigned long long)*4));
}

void launchDirect(State& st, const double* dPos, bool forces, int emode, bool emitPairs, long long* dForce, long long* dDedq, cudaStream_t s) {
    if (!forces && emode == 0 && !emitPairs) return;
    CellPlan& c = st.cells;
    CellParams cp{st.N, c.nc[0], c.nc[1], c.nc[2], c.ncells, 1.0/st.box.L[0], 1.0/st.box.L[1], 1.0/st.box.L[2], c.csd[0], c.csd[1], c.csd[2]};
    CFX_CUDA(cudaMemsetAsync(st.cellCount, 0, sizeof(int)*(c.ncells + 1), s));
    cellAssignKernel<<<(st.N + 255)/256, 256, 0, s>>>(cp, dPos, st.qf, st.cellOfAtom, st.userLocal, st.cellCount);
    CFX_LAUNCH_CHECK(); st.launches++;
    cellScanKernel<<<1, 1024, 0, s>>>(c.ncells, st.cellCount, st.cellStart, st.cellFill);
    CFX_LAUNCH_CHECK(); st.launches++;
    cellFillKernel<<<(st.N + 255)/256, 256, 0, s>>>(st.N, st.cellOfAtom, st.cellFill, st.filledUser);
    CFX_LAUNCH_CHECK(); st.launches++;
    cellRankGatherKernel<<<(st.N + 255)/256, 256, 0, s>>>(st.N, c.nc[1], c.nc[2], st.filledUser, st.cellOfAtom, st.cellStart,
            st.userLocal, st.lj, st.sortedUser, st.sortedLocal, st.sortedCell, st.sortedLJ);
    CFX_LAUNCH_CHECK(); st.launches++;
    mark(st, "cell_build", s);

    PairParams pp;
    pp.N = st.N; pp.Npad = st.Npad;
    pp.numGroups = (st.N + P_ITILE - 1)/P_ITILE;
    // spatial sharding: contiguous ranges of i-tiles in cell order are spatial slabs
    pp.groupLo = (int) ((int64_t) pp.numGroups*st.shardRank/st.shardCount);
    pp.groupHi = (int) ((int64_t) pp.numGroups*(st.shardRank + 1)/st.shardCount);
    pp.ncx = c.nc[0]; pp.ncy = c.nc[1]; pp.ncz = c.nc[2];
    pp.csx = c.cs[0]; pp.csy = c.cs[1]; pp.csz = c.cs[2];
    pp.Lx = (float) st.box.L[0]; pp.Ly = (float) st.box.L[1]; pp.Lz = (float) st.box.L[2];
    pp.invLx = (float) (1.0/st.box.L[0]); pp.invLy = (float) (1.0/st.box.L[1]); pp.invLz = (float) (1.0/st.box.L[2]);
    pp.dLx = st.box.L[0]; pp.dLy = st.box.L[1]; pp.dLz = st.box.L[2];
    pp.rc2d = st.cutoff*st.cutoff;
    pp.rc2 = (float) pp.rc2d; pp.alpha = (float) st.alpha; pp.band = (float) (1e-5*pp.rc2d);
    pp.sortedLocal = st.sortedLocal; pp.sortedCell = st.sortedCell; pp.sortedLJ = st.sortedLJ; pp.sortedUser = st.sortedUser;
    pp.cellStart = st.cellStart; pp.exclPtr = st.exclPtr; pp.exclCols = st.exclCols; pp.pos = dPos;
    pp.q = st.q; pp.ljd = st.ljd; pp.alphaD = st.alpha;
    pp.dInvLx = 1.0/st.box.L[0]; pp.dInvLy = 1.0/st.box.L[1]; pp.dInvLz = 1.0/st.box.L[2];
    pp.forceFixed = dForce; pp.dedqFixed = dDedq; pp.energyFixed = st.energyFixed;
    pp.counters = st.pairCounters; pp.pairBuffer = st.pairBuffer; pp.pairCapacity = (unsigned long long) st.pairCapacity;
    const int groups = pp.groupHi - pp.groupLo;
    if (groups <= 0) return;
    const int blocks = (groups + P_WARPS - 1)/P_WARPS;
    // Deal each i-tile's stencil columns over several CTAs until ~10 CTAs per SM exist (5 with the FP64 energy queue,
    // whose 32-pair batches fill more slowly when split): a shard with few i-tiles would otherwise run at single-CTA
    // latency (~0.1 ms), and at one GPU 2 splits smooth the last wave (0.247 -> 0.222 ms at C3).
    int numSM = 148;
    cudaDeviceGetAttribute(&numSM, cudaDevAttrMultiProcessorCount, st.device);
    pp.jSplits = std::max(1, std::min(8, ((emode == 2 ? 5 : 10)*numSM + blocks - 1)/blocks));
    if (const char* e = getenv("CFX_PAIR_JSPLITS")) pp.jSplits = std::max(1, std::min(8, atoi(e)));     // experiments
    if (emitPairs) pp.jSplits = 1;
    const dim3 grid(blocks, pp.jSplits);
    if (emitPairs) dispatchPair<true>(pp, forces, emode, grid, s);
    else           dispatchPair<false>(pp, forces, emode, grid, s);
    CFX_LAUNCH_CHECK(); st.launches++;
    mark(st, "direct_pairs", s);
}

} // namespace cfx

// direct.cu -- piece (2): Ewald direct space (ReferenceCoulKernels.cpp:559-593) on a GPU-built
// cell list.
//
// The reference rebuilds a neighbour list every call (all i<j, not excluded, min-image r2 <= rc2) and
// evaluates erfc Coulomb + Lennard-Jones on it in FP64. Here:
//
//  cell build   atoms are wrapped (FP64), binned into cells of edge >= rc/2 and sorted by cell (z fastest) and, inside a
//               cell, by z: a column of cells is one z-ordered run of atoms, so 8 consecutive sorted atoms -- an
//               i-cluster -- occupy a flat slab of the column (about 0.53 x 0.53 x 0.29 nm in water) and their bounding
//               box prunes well. Each sorted atom keeps its position as a FP32 offset inside its own cell, so pair
//               separations are formed from small numbers (no loss of precision in large boxes).
//  pair kernel  one warp per i-cluster, 4 lanes per i atom. The warp walks the (x, y) columns of the 5x5 stencil; per
//               column the z range is clipped to the sphere of radius rc around the cluster's bounding box; candidates
//               are tested against the bounding box 32 at a time and ballot-compacted into two 64-entry rings in shared
//               memory (j atoms with and without a Lennard-Jones well depth: the second kind -- two thirds of the
//               atoms of water -- runs an inner loop without the LJ arithmetic). Whenever a ring holds 32 entries the
//               warp runs the straight-line inner loop over them: lane (i, part) takes entries part, part+4, ...
//               (one LDS.128 wavefront per 4 j atoms), everything after the distance test is predicated, not
//               branched. Forces and dE/dq are accumulated together in registers and reduced over the 4 lanes of an
//               i atom with warp shuffles; every ordered pair is evaluated from both sides (full shell): no j-side
//               atomics, and the FP32 summation order is fixed, so results are bitwise reproducible.
//  erfc         erfc(x) = exp(-x^2) t P7(t), t = 1/(1 + p x) (tools/erfcx_fit.py, 5e-8 relative): one MUFU.RCP, one
//               MUFU.EX2 (shared with the force's Gaussian term), 7 FFMA.
//  exactness    the in-cutoff predicate of a pair whose FP32 r2 falls within 1e-5 of rc2 is re-evaluated in FP64 on the
//               original coordinates with the reference's operation order, so the neighbour set is bit-exact.
//  exclusions   the excluded-pair kernel (flux.cu), which runs first, leaves for every atom the largest r2 to any of its
//               excluded partners; only pairs of that atom at or below it (bonded neighbours; also the self pair at
//               r2 = 0) take the rare branch that probes the per-atom exclusion CSR.
//
// FP32 pair arithmetic for forces and dE/dq, int64 fixed-point accumulation. Pair ENERGIES are evaluated in FP64:
// in-cutoff i<j pairs are ballot-compacted into a per-warp queue and evaluated 32 at a time, because E_direct cancels
// against E_self + E_excl to a small fraction of its size.
#include "cfx_internal.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <type_traits>

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
// The cell-build kernels and the list builder run only in evaluations whose rebuild flag is set (displacementKernel):
// the launches stay in the step's CUDA graph and return at once otherwise.
__global__ void __launch_bounds__(256) cellAssignKernel(CellParams p, const int* __restrict__ rebuildFlag, const double* __restrict__ pos,
        const float* __restrict__ qf, const double* __restrict__ qd, int* __restrict__ cellOfAtom, float4* __restrict__ userLocal,
        double4* __restrict__ userLocalD, int* __restrict__ cellCount, double* __restrict__ posAtBuild) {
    if (!*rebuildFlag) return;
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= p.N) return;
    posAtBuild[3*(size_t) i] = pos[3*(size_t) i]; posAtBuild[3*(size_t) i + 1] = pos[3*(size_t) i + 1]; posAtBuild[3*(size_t) i + 2] = pos[3*(size_t) i + 2];
    double u[3] = {pos[3*(size_t) i]*p.invLx, pos[3*(size_t) i + 1]*p.invLy, pos[3*(size_t) i + 2]*p.invLz};
    const int nc[3] = {p.ncx, p.ncy, p.ncz};
    const double cs[3] = {p.csx, p.csy, p.csz};
    int c[3];
    double loc[3];
    #pragma unroll
    for (int d = 0; d < 3; d++) {
        double f = u[d] - floor(u[d]);          // [0,1)
        double g = f*nc[d];
        int k = (int) g;
        if (k >= nc[d]) k = nc[d] - 1;
        c[d] = k;
        loc[d] = (g - k)*cs[d];
    }
    const int cell = (c[0]*p.ncy + c[1])*p.ncz + c[2];
    cellOfAtom[i] = cell;
    userLocal[i] = make_float4((float) loc[0], (float) loc[1], (float) loc[2], qf[i]);
    userLocalD[i] = make_double4(loc[0], loc[1], loc[2], qd[i]);    // FP64 copy: pair energies (EMODE 2)
    atomicAdd(cellCount + cell, 1);
}

// exclusive scan of cellCount -> cellStart (single CTA; ncells is at most a few 10^4)
__global__ void __launch_bounds__(1024) cellScanKernel(int ncells, const int* __restrict__ rebuildFlag, const int* __restrict__ cellCount,
        int* __restrict__ cellStart, int* __restrict__ cellFill, unsigned long long* __restrict__ wrapCount) {
    if (!*rebuildFlag) return;
    if (threadIdx.x == 0) *wrapCount = 0ull;               // the list builder behind this kernel refills the generic kernel's list
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

__global__ void __launch_bounds__(256) cellFillKernel(int N, const int* __restrict__ rebuildFlag, const int* __restrict__ cellOfAtom,
        int* __restrict__ cellFill, int* __restrict__ sortedUser) {
    if (!*rebuildFlag) return;
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int slot = atomicAdd(cellFill + cellOfAtom[i], 1);
    sortedUser[slot] = i;
}

// One thread per slot of the (arbitrarily ordered) cell fill: the atom's final slot is the start of its cell plus its
// rank among the cell's atoms by (z inside the cell, user index), so the sorted order -- and with it every FP32
// accumulation order downstream -- is deterministic, and a column of cells is one z-ordered run. Writes both sorted
// records in one pass: sortedLocal = (x, y, z inside the cell, q), sortedMeta = (sigma/2, 2 sqrt(eps), user index,
// packed cell coordinates).
__global__ void __launch_bounds__(256) cellRankGatherKernel(int N, int ncy, int ncz, const int* __restrict__ rebuildFlag, const int* __restrict__ filledUser,
        const int* __restrict__ cellOfAtom, const int* __restrict__ cellStart, const float4* __restrict__ userLocal,
        const float2* __restrict__ lj, float4* __restrict__ sortedLocal, float4* __restrict__ sortedMeta,
        const double4* __restrict__ userLocalD, const double2* __restrict__ ljd, double4* __restrict__ sortedLocalD,
        double2* __restrict__ sortedLjD) {
    if (!*rebuildFlag) return;
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= N) return;
    const int u = filledUser[s];
    const int cell = cellOfAtom[u];
    const int s0 = cellStart[cell], s1 = cellStart[cell+1];
    const float4 mine = userLocal[u];
    int rank = 0;
    for (int t = s0; t < s1; t++) {
        const int v = filledUser[t];
        const float zv = userLocal[v].z;
        rank += (zv < mine.z || (zv == mine.z && v < u)) ? 1 : 0;
    }
    const int dst = s0 + rank;
    const int cz = cell % ncz, cy = (cell/ncz) % ncy, cx = cell/(ncz*ncy);
    const float2 l = lj[u];
    sortedLocal[dst] = mine;
    sortedLocalD[dst] = userLocalD[u];
    sortedLjD[dst] = ljd[u];
    sortedMeta[dst] = make_float4(l.x, l.y, __int_as_float(u), __int_as_float(cx | (cy << CELL_BITS) | (cz << (2*CELL_BITS))));
}

// Has any atom moved more than skin/2 (minimum image) since the lists were built? Then this evaluation rebuilds.
__global__ void __launch_bounds__(256) displacementKernel(int N, const double* __restrict__ pos, const double* __restrict__ posAtBuild,
        double Lx, double Ly, double Lz, double limit2, int* __restrict__ rebuildFlag) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= N || *rebuildFlag) return;
    double dx = pos[3*(size_t) i] - posAtBuild[3*(size_t) i];
    double dy = pos[3*(size_t) i + 1] - posAtBuild[3*(size_t) i + 1];
    double dz = pos[3*(size_t) i + 2] - posAtBuild[3*(size_t) i + 2];
    dx -= Lx*rint(dx/Lx); dy -= Ly*rint(dy/Ly); dz -= Lz*rint(dz/Lz);
    if (!(dx*dx + dy*dy + dz*dz <= limit2)) *rebuildFlag = 1;          // (NaN positions rebuild too)
}

// Evaluations that reuse the lists keep the atoms in the order of the last build: the sorted records get this
// evaluation's coordinates -- relative to the cell the atom was in at the build, so slightly outside [0, cell size) --
// and charges. Runs in every evaluation (the charges change with the geometry).
__global__ void __launch_bounds__(256) refreshSortedKernel(CellParams p, const double* __restrict__ pos, const float* __restrict__ qf,
        const double* __restrict__ qd, float4* __restrict__ sortedLocal, const float4* __restrict__ sortedMeta, double4* __restrict__ sortedLocalD) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= p.N) return;
    const float4 m = sortedMeta[s];
    const int u = __float_as_int(m.z), cj = __float_as_int(m.w);
    const int c[3] = {cj & CELL_MASK, (cj >> CELL_BITS) & CELL_MASK, cj >> (2*CELL_BITS)};
    const double uu[3] = {pos[3*(size_t) u]*p.invLx, pos[3*(size_t) u + 1]*p.invLy, pos[3*(size_t) u + 2]*p.invLz};
    const int nc[3] = {p.ncx, p.ncy, p.ncz};
    const double cs[3] = {p.csx, p.csy, p.csz};
    double loc[3];
    #pragma unroll
    for (int d = 0; d < 3; d++) {
        const double f = uu[d] - floor(uu[d]);
        double g = f*nc[d] - c[d];                                     // in cells, relative to the build-time cell
        if (g > 0.5*nc[d]) g -= nc[d];
        if (g < -0.5*nc[d]) g += nc[d];
        loc[d] = g*cs[d];
    }
    sortedLocal[s] = make_float4((float) loc[0], (float) loc[1], (float) loc[2], qf[u]);
    sortedLocalD[s] = make_double4(loc[0], loc[1], loc[2], qd[u]);
}

// The search half of the direct-space branch (displacement check, re-sort, lists) depends on the positions only and is
// enqueued before the charge-flux assembly has finished; the sorted records get this evaluation's charges here.
__global__ void __launch_bounds__(256) refreshChargesKernel(int N, const float* __restrict__ qf, const double* __restrict__ qd,
        const float4* __restrict__ sortedMeta, float4* __restrict__ sortedLocal, double4* __restrict__ sortedLocalD) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= N) return;
    const int u = __float_as_int(sortedMeta[s].z);
    sortedLocal[s].w = qf[u];
    sortedLocalD[s].w = qd[u];
}

__global__ void finishListKernel(int* __restrict__ rebuildFlag, unsigned long long* __restrict__ counters) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && *rebuildFlag) { *rebuildFlag = 0; counters[13] += 1ull; }     // [13]: list builds so far
}

// ------------------------------------------------------------------------------------------------
// pair kernel
// ------------------------------------------------------------------------------------------------
#define P_WARPS 4
#define P_ITILE 8
#define P_JTILE 32
#ifndef P_FAST_MINBLOCKS
#define P_FAST_MINBLOCKS 5     // resident CTAs per SM the fast kernel is compiled for (register cap 96)
#endif
#define R_COMPS 8
#define RD_COMPS 6
#define E_POLY_DEG 20
#ifndef P_UNROLL
#define P_UNROLL 1        // the packed loop already carries two independent pairs per iteration
#endif
constexpr int kPairUnroll = P_UNROLL;
#ifndef P_EUNROLL
#define P_EUNROLL 1       // FP64 energy passes
#endif
constexpr int kEnergyUnroll = P_EUNROLL;
#define P_MAX_JSPLITS 8        // shares of one cluster's candidate list (small systems, shards)
#define P_JCAP 64            // ring capacity: < 32 waiting entries + one 32-candidate chunk

struct PairParams {
    int N, Npad, numGroups, groupLo, groupHi;
    int jSplits;                 // the (x,y) columns of an i-cluster's stencil are dealt over jSplits work items (small shards)
    int onlyMinImage;            // generic kernel launched after the fast one: only the clusters the fast one skipped
    int ncx, ncy, ncz;
    float csx, csy, csz, invCsz;
    float Lx, Ly, Lz, invLx, invLy, invLz;
    double dLx, dLy, dLz, rc2d;
    float rc2, alpha, alpha2, band;
    const float4* sortedLocal; const float4* sortedMeta;
    const int* cellStart;
    const int* exclPtr; const int* exclCols;
    const unsigned int* exclMaxR2Bits;       // [N] per user atom: float bits of the largest r2 to an excluded partner (exclusionKernel)
    const double* pos; const double* q; const double2* ljd; double alphaD, dInvLx, dInvLy, dInvLz;
    // FP64 pair energies (EMODE 2): cell-local coordinates + charge and LJ parameters of the sorted atoms in double, and the
    // polynomial of erf(sqrt z)/sqrt z in t = eTScale s - 1 (see fitEnergyPolynomial)
    const double4* sortedLocalD; const double2* sortedLjD;
    double ePoly[E_POLY_DEG + 1]; double eTScale; int ePolyOK;
    double dcsx, dcsy, dcsz;
    long long* forceFixed; long long* dedqFixed; long long* energyFixed;
    unsigned long long* counters; int2* pairBuffer; unsigned long long pairCapacity;
    unsigned int* workCounter;               // dynamic work distribution: next (cluster, column share) item
    int* wrapList;                           // clusters the fast kernel left to the generic one (*wrapCount of them)
    unsigned long long* wrapCount;
    // candidate lists of the fast path (buildListKernel): for i-cluster g, listCount[g - groupLo] = (entries at the front,
    // entries at the back) of pairList + (g - groupLo)*listCap (x negative: the cluster is left to the generic kernel),
    // each entry = sorted index | image code << 27
    unsigned int* pairList; int2* listCount; int listCap;
    const int* rebuildFlag;                  // list builder: return at once unless set
    float rlist2;                            // (cutoff + skin)^2: what the lists are built for
    float drift;                             // skin/2: how far an atom may be outside the cell it was sorted into
    int countStats;                          // this pass adds to counters[0..1] (the second pass of an energy+forces call does not)
};

__device__ __forceinline__ int wrapNearest(int d, int nc) {
    const int half = (nc - 1) >> 1;
    if (d > half) d -= nc;
    if (d < -(nc >> 1)) d += nc;
    return d;
}

// approximate special functions without the denormal fix-ups nvcc wraps around rsqrtf/__expf (inputs here are normal)
__device__ __forceinline__ float rsqrtFtz(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpFtz(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2Ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// exact FP64 predicate with the reference's operation order (J - I with I the lower user index;
// z, y, x floor-based wrap; left-to-right sum of squares; no FMA contraction)
__device__ __forceinline__ bool exactInCutoff(const double* __restrict__ pos, int ua, int ub, double Lx, double Ly, double Lz, double rc2) {
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

// rare path: is (ui, uj) the self pair or an excluded pair? (symmetric, sorted per-atom CSR)
__device__ __forceinline__ bool selfOrExcluded(const int* __restrict__ exclPtr, const int* __restrict__ exclCols, int ui, int uj) {
    if (ui == uj) return true;
    for (int e = exclPtr[ui], e1 = exclPtr[ui + 1]; e < e1; e++)
        if (exclCols[e] == uj) return true;
    return false;
}

// The two rare cases of the inner loop in one out-of-line call: a pair whose FP32 r2 lies within the band around rc2
// (decided in FP64 with the reference's operation order), and a pair at bonded-neighbour separation (self / excluded?).
__device__ __noinline__ int rareInCutoff(const PairParams* p, int ui, int uj, float r2, float r2close) {
    bool in = r2 <= p->rc2;
    if (fabsf(r2 - p->rc2) < p->band) in = exactInCutoff(p->pos, ui, uj, p->dLx, p->dLy, p->dLz, p->rc2d);
    if (in && r2 <= r2close) in = !selfOrExcluded(p->exclPtr, p->exclCols, ui, uj);
    return in ? 1 : 0;
}

// EMODE 2, fallback when the polynomial of fitEnergyPolynomial does not cover alpha^2 rc^2
__device__ __noinline__ double closePairEnergy(const PairParams* p, double s, double kqq, double sig, double eps) {
    const double invR = rsqrt(s);
    double s2 = sig*invR; s2 *= s2;
    const double s6 = s2*s2*s2;
    return kqq*invR*erfc(p->alphaD*s*invR) + eps*s6*(s6 - 1.0);
}

__device__ __forceinline__ float2 pk(float a) { return make_float2(a, a); }     // scalar broadcast operand of a packed instruction
__device__ __forceinline__ int modPos(int v, int n) { v %= n; return v < 0 ? v + n : v; }
// v in (-n, 2n): the fast kernel's cell grids have at least 7 cells per axis and stencils of at most n cells
__device__ __forceinline__ int wrapOnce(int v, int n) { return v < 0 ? v + n : (v >= n ? v - n : v); }

// FAST: the cell grid has >= 7 cells per axis; clusters whose stencil would wrap onto itself (min-image needed) are
//       left to the generic instantiation, so no min-image code and no integer divisions are compiled in.
// EMODE: 0 = no pair energy, 1 = FP32 pair terms (the partial energy the reference returns when
// includeEnergy is false is discarded by OpenMM; it is still produced, at FP32 accuracy), 2 = FP64 terms.
template <bool FAST, bool FORCES, int EMODE, bool EMIT>
__global__ void __launch_bounds__(P_WARPS*32, (FAST && EMODE != 2) ? P_FAST_MINBLOCKS : 4) pairKernel(const __grid_constant__ PairParams p) {
    // rings of staged j atoms, [0: without LJ well depth, 1: with][warp][slot], one array per component so that a lane
    // fetches the same component of two consecutive entries with one LDS.64 -- the operand layout of the packed
    // (two pairs per instruction) FP32 arithmetic of the inner loop
    // components: 0-2 xyz in the cluster frame, 3 charge, 4 user index (rare paths, pair emission, energy queue),
    // 5-6 sigma/2 and 2 sqrt(eps) (LJ ring only), 7 sorted index (FP64 energies: a pair is counted from the side of the
    // atom that comes first in sorted order); one block per warp so that every access is base + constant offset
    __shared__ __align__(16) float sRing[P_WARPS][2][R_COMPS][P_JCAP];
    // EMODE 2: the same entries in double -- 0-2 xyz in the cluster frame, 3 charge, 4-5 sigma/2 and 2 sqrt(eps)
    __shared__ __align__(16) double sRingD[EMODE == 2 ? P_WARPS : 1][2][EMODE == 2 ? RD_COMPS : 1][EMODE == 2 ? P_JCAP : 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int ii = lane >> 2, part = lane & 3;
    const unsigned int lt = (1u << lane) - 1u;
    const float alpha = p.alpha, alpha2 = p.alpha2, band = p.band;
    const float rcut2 = p.rc2*1.0001f;
    const float4 farAway = make_float4(1e4f, 1e4f, 1e4f, 0.f);
    // The generic kernel behind a fast one only sees the clusters listed in wrapList (usually none: it exits at once);
    // there are few of them, so each is dealt over 256 shares (one stencil column per warp) to keep its latency short.
    const bool listed = !FAST && p.onlyMinImage;
    const int nShares = listed ? 256 : p.jSplits;
    // energy-only passes count every pair once, from the atom that comes first in sorted order: they scan only the
    // part of the stencil behind the cluster (half shell: no j-side accumulation is needed for an energy)
    constexpr bool HALF = !FORCES && EMODE == 2 && !EMIT;
    const unsigned int totalItems = (listed ? (unsigned int) *p.wrapCount : (unsigned int) (p.groupHi - p.groupLo))*(unsigned int) nShares;

    // persistent warps: work items (i-cluster, share of its stencil columns) are handed out by an atomic counter, so the
    // GPU stays full whatever the item count (no partial last wave)
    for (;;) {
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(p.workCounter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= totalItems) break;
        const int share = (int) (item % (unsigned int) nShares);
        const int g = listed ? p.wrapList[item/(unsigned int) nShares] : p.groupLo + (int) (item/(unsigned int) nShares);

        const int i0 = g*P_ITILE;
        const int ni = min(P_ITILE, p.N - i0);
        const bool validI = ii < ni;
        const int iIdx = i0 + (validI ? ii : 0);
        const int c0 = __float_as_int(p.sortedMeta[i0].w);
        const int c0x = c0 & CELL_MASK, c0y = (c0 >> CELL_BITS) & CELL_MASK, c0z = c0 >> (2*CELL_BITS);

        // my i atom, expressed in the frame of cell c0
        const float4 li = p.sortedLocal[iIdx];
        const float4 mi = p.sortedMeta[iIdx];
        const int ci = __float_as_int(mi.w);
        const int ox = wrapNearest((ci & CELL_MASK) - c0x, p.ncx);
        const int oy = wrapNearest(((ci >> CELL_BITS) & CELL_MASK) - c0y, p.ncy);
        const int oz = wrapNearest((ci >> (2*CELL_BITS)) - c0z, p.ncz);
        const float pix = li.x + ox*p.csx, piy = li.y + oy*p.csy, piz = li.z + oz*p.csz;
        const float ljix = mi.x, ljiy = mi.y;
        const int ui = __float_as_int(mi.z);
        const float keqi = (float) CFX_ONE_4PI_EPS0*li.w;
        double pixD = 0.0, piyD = 0.0, pizD = 0.0, kqiD = 0.0, sigiD = 0.0, epsiD = 0.0;
        if (EMODE == 2) {
            const double4 ld = p.sortedLocalD[iIdx];
            const double2 lld = p.sortedLjD[iIdx];
            pixD = ld.x + ox*p.dcsx; piyD = ld.y + oy*p.dcsy; pizD = ld.z + oz*p.dcsz;
            kqiD = CFX_ONE_4PI_EPS0*ld.w; sigiD = lld.x; epsiD = lld.y;
        }
        // per-lane copies of the thresholds: lanes without an i atom (last cluster) never see a pair
        const float rc2i = validI ? p.rc2 : -1.f;
        // (floor 1e-8 nm^2: the self pair is r2 = 0 up to rounding when a small box folds the stencil onto itself)
        const float r2close = validI ? fmaxf(__uint_as_float(p.exclMaxR2Bits[ui])*1.0001f, 1e-8f) : -1.f;
        const bool anyLJi = __any_sync(0xffffffffu, validI && ljiy != 0.f);

        // bounding box of the i-cluster and its cell-offset range (warp reductions)
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
        // i-cluster straddling a large empty region) it is truncated to all nc cells of the axis and the
        // separation of every pair is min-imaged instead (warp-uniform flag; generic instantiation only).
        const int loX = ominx - 2, nX = min(omaxx - ominx + 5, p.ncx);
        const int loY = ominy - 2, nY = min(omaxy - ominy + 5, p.ncy);
        const int loZ = ominz - 2, nZ = min(omaxz - ominz + 5, p.ncz);
        const bool wraps = (omaxx - ominx + 5 > p.ncx) || (omaxy - ominy + 5 > p.ncy) || (omaxz - ominz + 5 > p.ncz);
        int listN = 0, listFront = 0;
        if (FAST) {                                     // negative: buildListKernel left the cluster to the generic kernel
            const int2 lc = p.listCount[g - p.groupLo];
            if (lc.x < 0) continue;
            listFront = lc.x;
            listN = HALF ? lc.x : lc.x + lc.y;          // energy passes: only the atoms behind the cluster in sorted order
        }
        const bool minImage = FAST ? false : wraps;

        float2 fx2 = pk(0.f), fy2 = pk(0.f), fz2 = pk(0.f), dq2 = pk(0.f), enf2 = pk(0.f);     // two partial sums each
        double en = 0.0;
        unsigned int nIn = 0, nCand = 0;                 // in-cutoff ordered pairs seen by this lane; candidates staged by the warp

        // One tile = the 32 ring entries from `base` (0 or 32) of ring `LJ`; entries past the end n of a last, partial
        // tile hold far-away positions with zero charge and well depth (min-image tiles, which would fold them back
        // into the box, test the entry index against n as well). Each lane takes two consecutive entries per iteration
        // and evaluates them with packed FP32 instructions (FFMA2 / FMUL2 / FADD2 of sm_100: one issue slot for two
        // pairs; the kernel is issue-bound, not FMA-pipe-bound). Straight-line code: out-of-cutoff pairs are masked,
        // not branched.
        auto processTile = [&](auto ljTag, auto imgTag, int base, int n, bool doE) {
            constexpr bool LJ = decltype(ljTag)::value, IMG = decltype(imgTag)::value;
            const float* ring = &sRing[warp][LJ][0][0] + base;
            const float2* tX = reinterpret_cast<const float2*>(ring);
            const float2* tY = reinterpret_cast<const float2*>(ring + P_JCAP);
            const float2* tZ = reinterpret_cast<const float2*>(ring + 2*P_JCAP);
            const float2* tQ = reinterpret_cast<const float2*>(ring + 3*P_JCAP);
            const int* tUser = reinterpret_cast<const int*>(ring + 4*P_JCAP);
            const float2* tSig = reinterpret_cast<const float2*>(ring + 5*P_JCAP);
            const float2* tEps = reinterpret_cast<const float2*>(ring + 6*P_JCAP);
            __syncwarp();
            #pragma unroll (EMODE == 2 ? kEnergyUnroll : kPairUnroll)
            for (int c = part; c < P_JTILE/2; c += 4) {
                const float2 xj = tX[c], yj = tY[c], zj = tZ[c], qj = tQ[c];
                float2 dx = __ffma2_rn(xj, pk(-1.f), pk(pix));                 // pos[i] - pos[j], two j atoms
                float2 dy = __ffma2_rn(yj, pk(-1.f), pk(piy));
                float2 dz = __ffma2_rn(zj, pk(-1.f), pk(piz));
                float2 imx = pk(0.f), imy = pk(0.f), imz = pk(0.f);               // min-image shifts in box lengths
                if (IMG) {
                    imx = make_float2(rintf(dx.x*p.invLx), rintf(dx.y*p.invLx));
                    imy = make_float2(rintf(dy.x*p.invLy), rintf(dy.y*p.invLy));
                    imz = make_float2(rintf(dz.x*p.invLz), rintf(dz.y*p.invLz));
                    dx.x -= p.Lx*imx.x; dy.x -= p.Ly*imy.x; dz.x -= p.Lz*imz.x;
                    dx.y -= p.Lx*imx.y; dy.y -= p.Ly*imy.y; dz.y -= p.Lz*imz.y;
                }
                const float2 r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                bool in0 = r2.x <= rc2i, in1 = r2.y <= rc2i;
                if (IMG) { in0 = in0 && 2*c < n; in1 = in1 && 2*c + 1 < n; }
                // rare: within the band around rc2 (FP64 re-test, bit-exact neighbour set), or at bonded-neighbour
                // separation (excluded pair? self pair?). Lanes without an i atom have negative thresholds.
                const float2 off = __fadd2_rn(r2, pk(-rc2i));
                const bool rare0 = ((r2.x <= r2close) | (fabsf(off.x) < band)) & (!IMG || 2*c < n);
                const bool rare1 = ((r2.y <= r2close) | (fabsf(off.y) < band)) & (!IMG || 2*c + 1 < n);
                if (rare0 | rare1) {
                    if (rare0) in0 = rareInCutoff(&p, ui, tUser[2*c], r2.x, r2close) != 0;
                    if (rare1) in1 = rareInCutoff(&p, ui, tUser[2*c + 1], r2.y, r2close) != 0;
                }
                if (EMODE == 2 && doE) {
                    // FP64 pair energy (E_direct cancels against E_self + E_excl to a small fraction of its size, FP32 terms
                    // would cost ~1e-3 kJ/mol): r^2 in double from the double copies of the staged coordinates, the
                    // Coulomb term as 1/r (FP32 rsqrt + one Newton step in double) minus a degree-20 polynomial in s (no erfc,
                    // exp, division or table), the LJ term from the same 1/r. Every pair is
                    // seen from both sides and counted from the side of the atom that comes first in sorted order: tiles
                    // whose j atoms all precede the cluster (about half of them: doE false) skip this block.
                    const double* ringD = &sRingD[warp][LJ][0][0] + base;
                    const double2 xd = reinterpret_cast<const double2*>(ringD)[c];
                    const double2 yd = reinterpret_cast<const double2*>(ringD + P_JCAP)[c];
                    const double2 zd = reinterpret_cast<const double2*>(ringD + 2*P_JCAP)[c];
                    const double2 qd = reinterpret_cast<const double2*>(ringD + 3*P_JCAP)[c];
                    const int2 sj2 = reinterpret_cast<const int2*>(ring + 7*P_JCAP)[c];
                    double2 sgd = make_double2(0.0, 0.0), epd = make_double2(0.0, 0.0);
                    if (LJ) {
                        sgd = reinterpret_cast<const double2*>(ringD + 4*P_JCAP)[c];
                        epd = reinterpret_cast<const double2*>(ringD + 5*P_JCAP)[c];
                    }
                    #pragma unroll
                    for (int h = 0; h < 2; h++) {
                        double ddx = pixD - (h ? xd.y : xd.x), ddy = piyD - (h ? yd.y : yd.x), ddz = pizD - (h ? zd.y : zd.x);
                        if (IMG) {
                            ddx -= p.dLx*(double) (h ? imx.y : imx.x); ddy -= p.dLy*(double) (h ? imy.y : imy.x);
                            ddz -= p.dLz*(double) (h ? imz.y : imz.x);
                        }
                        const double sD = fma(ddz, ddz, fma(ddy, ddy, ddx*ddx));
                        const float r2f = h ? r2.y : r2.x;
                        bool on = (h ? in1 : in0) && (h ? sj2.y : sj2.x) > iIdx;
                        const double qq = kqiD*(h ? qd.y : qd.x);                  // ONE_4PI_EPS0 q_i q_j
                        if (HALF && on) nIn += 2u;
                        if (!p.ePolyOK && on) {                                     // exotic tolerance (alpha rc > 4): libm erfc
                            en += closePairEnergy(&p, sD, qq, sigiD + (h ? sgd.y : sgd.x), epsiD*(h ? epd.y : epd.x));
                            on = false;
                        }
                        // 1/r: FP32 rsqrt refined once in double (relative error ~1e-13)
                        const double y0 = (double) rsqrtFtz(fmaxf(r2f, 1e-12f));      // (the self pair of a folded stencil: finite)
                        const double ey = fma(-(sD*y0), y0, 1.0);
                        const double y = fma(0.5*y0, ey, y0);
                        // erfc(alpha r)/r = 1/r - alpha G(alpha^2 s), G(z) = erf(sqrt z)/sqrt z entire in z: one polynomial
                        const double t = fma(sD, p.eTScale, -1.0);
                        // even and odd part as two independent Horner chains in t^2 (the single chain of 20 dependent DFMAs
                        // left the FP64 pipe waiting on its own latency)
                        const double t2 = t*t;
                        double ge = p.ePoly[E_POLY_DEG], go = p.ePoly[E_POLY_DEG - 1];
                        #pragma unroll
                        for (int k = E_POLY_DEG - 2; k >= 0; k -= 2) {
                            ge = fma(ge, t2, p.ePoly[k]);
                            if (k > 0) go = fma(go, t2, p.ePoly[k - 1]);
                        }
                        const double g = fma(go, t, ge);
                        const double fv = fma(-p.alphaD, g, y);
                        en = fma(on ? qq : 0.0, fv, en);
                        if (LJ) {
                            const double yy = y*y;
                            const double sg = sigiD + (h ? sgd.y : sgd.x);
                            const double s2 = sg*sg*yy, s6 = s2*s2*s2;
                            const double elj = epsiD*(h ? epd.y : epd.x)*s6*(s6 - 1.0);
                            en += on ? elj : 0.0;
                        }
                    }
                }
                if (FORCES || EMODE == 1) {
                    const float2 invR = make_float2(rsqrtFtz(r2.x), rsqrtFtz(r2.y));
                    const float2 invR2 = __fmul2_rn(invR, invR);
                    const float2 ar = __fmul2_rn(__fmul2_rn(r2, pk(alpha)), invR);
                    const float2 ge = __fmul2_rn(r2, pk(alpha2*-1.4426950408889634f));
                    const float2 gauss = make_float2(ex2Ftz(ge.x), ex2Ftz(ge.y));      // exp(-(alpha r)^2)
                    const float2 td = __ffma2_rn(ar, pk(0.3275911f), pk(1.f));
                    const float2 t = make_float2(rcpFtz(td.x), rcpFtz(td.y));
                    float2 pl = __ffma2_rn(t, pk(-1.438181028e-01f), pk(5.078454972e-01f));
                    pl = __ffma2_rn(t, pl, pk(-4.250064424e-01f));
                    pl = __ffma2_rn(t, pl, pk(4.706512873e-01f));
                    pl = __ffma2_rn(t, pl, pk(1.636827149e-02f));
                    pl = __ffma2_rn(t, pl, pk(2.085574698e-01f));
                    pl = __ffma2_rn(t, pl, pk(1.803253944e-01f));
                    pl = __ffma2_rn(t, pl, pk(1.850765712e-01f));
                    const float2 erfcv = __fmul2_rn(__fmul2_rn(t, pl), gauss);
                    const float2 coul = __fmul2_rn(__fmul2_rn(qj, pk(keqi)), invR);
                    float2 es6 = pk(0.f), s6 = pk(0.f);
                    if (LJ) {
                        const float2 sig = __fadd2_rn(tSig[c], pk(ljix));
                        float2 s2 = __fmul2_rn(sig, invR);
                        s2 = __fmul2_rn(s2, s2);
                        s6 = __fmul2_rn(__fmul2_rn(s2, s2), s2);
                        es6 = __fmul2_rn(s6, __fmul2_rn(tEps[c], pk(ljiy)));
                    }
                    if (FORCES) {
                        const float2 gterm = __ffma2_rn(__fmul2_rn(ar, gauss), pk(1.1283791671f), erfcv);
                        float2 dEdR = __fmul2_rn(coul, gterm);
                        if (LJ) dEdR = __ffma2_rn(es6, __ffma2_rn(s6, pk(12.f), pk(-6.f)), dEdR);
                        dEdR = __fmul2_rn(dEdR, invR2);
                        float2 v = __fmul2_rn(invR, erfcv);
                        dEdR.x = in0 ? dEdR.x : 0.f; dEdR.y = in1 ? dEdR.y : 0.f;
                        v.x = in0 ? v.x : 0.f; v.y = in1 ? v.y : 0.f;
                        fx2 = __ffma2_rn(dEdR, dx, fx2); fy2 = __ffma2_rn(dEdR, dy, fy2); fz2 = __ffma2_rn(dEdR, dz, fz2);
                        dq2 = __ffma2_rn(qj, v, dq2);
                    }
                    if (EMODE == 1) {                                         // discarded partial energy: FP32 per lane
                        float2 e = __fmul2_rn(coul, erfcv);
                        if (LJ) e = __ffma2_rn(es6, __fadd2_rn(s6, pk(-1.f)), e);
                        e.x = in0 ? e.x : 0.f; e.y = in1 ? e.y : 0.f;
                        enf2 = __fadd2_rn(enf2, e);
                    }
                }
                if (!HALF) nIn += (in0 ? 1u : 0u) + (in1 ? 1u : 0u);
                if (EMIT) {
                    #pragma unroll
                    for (int h = 0; h < 2; h++) {
                        const int uj = tUser[2*c + h];
                        if ((h ? in1 : in0) && ui < uj) {
                            const unsigned long long slot = atomicAdd(p.counters + 2, 1ull);
                            if (slot < p.pairCapacity) p.pairBuffer[slot] = make_int2(ui, uj);
                        }
                    }
                }
            }
            __syncwarp();
        };

        // ring state per class: entries [base, base + count) mod 64 are waiting; base is 0 or 32
        int base0 = 0, count0 = 0, base1 = 0, count1 = 0;
        // largest sorted index among the waiting entries of each ring, and among those of the chunk staged last (what is
        // left after a flush comes from that chunk): a tile whose j atoms all precede the cluster needs no FP64 energies
        int hi0 = -1, hi1 = -1, lastHi0 = -1, lastHi1 = -1;
        // a full tile, or (last == true, after the last chunk) the partial one, padded
        auto flush = [&](bool last) {
            if (count0 >= P_JTILE || (last && count0 > 0)) {
                const int n = min(count0, P_JTILE);
                if (lane >= n) {
                    float* e = &sRing[warp][0][0][(base0 + lane) & (P_JCAP - 1)];
                    e[0] = 1e4f; e[P_JCAP] = 1e4f; e[2*P_JCAP] = 1e4f; e[3*P_JCAP] = 0.f;
                    if (EMODE == 2) {
                        double* ed = &sRingD[warp][0][0][(base0 + lane) & (P_JCAP - 1)];
                        ed[0] = 1e4; ed[P_JCAP] = 1e4; ed[2*P_JCAP] = 1e4; ed[3*P_JCAP] = 0.0;
                    }
                }
                const bool doE = hi0 > i0;
                if (minImage) processTile(std::false_type{}, std::true_type{}, base0, n, doE); else processTile(std::false_type{}, std::false_type{}, base0, n, doE);
                base0 ^= P_JTILE; count0 -= P_JTILE;
                hi0 = count0 > 0 ? lastHi0 : -1;
            }
            if (count1 >= P_JTILE || (last && count1 > 0)) {
                const int n = min(count1, P_JTILE);
                if (lane >= n) {
                    float* e = &sRing[warp][1][0][(base1 + lane) & (P_JCAP - 1)];
                    e[0] = 1e4f; e[P_JCAP] = 1e4f; e[2*P_JCAP] = 1e4f; e[3*P_JCAP] = 0.f; e[5*P_JCAP] = 0.f; e[6*P_JCAP] = 0.f;
                    if (EMODE == 2) {
                        double* ed = &sRingD[warp][1][0][(base1 + lane) & (P_JCAP - 1)];
                        ed[0] = 1e4; ed[P_JCAP] = 1e4; ed[2*P_JCAP] = 1e4; ed[3*P_JCAP] = 0.0; ed[4*P_JCAP] = 0.0; ed[5*P_JCAP] = 0.0;
                    }
                }
                const bool doE = hi1 > i0;
                if (minImage) processTile(std::true_type{}, std::true_type{}, base1, n, doE); else processTile(std::true_type{}, std::false_type{}, base1, n, doE);
                base1 ^= P_JTILE; count1 -= P_JTILE;
                hi1 = count1 > 0 ? lastHi1 : -1;
            }
        };

        if (FAST) {
            // fast path: the candidates of this cluster were listed by buildListKernel (the j atoms within the cutoff of
            // the cluster's bounding box, in stencil order); a share takes a contiguous range of 32-entry chunks
            const unsigned int* lst = p.pairList + (size_t) (g - p.groupLo)*p.listCap;
            const int nChunks = (listN + 31) >> 5;
            const int chunkLo = (int) ((long long) nChunks*share/nShares), chunkHi = (int) ((long long) nChunks*(share + 1)/nShares);
            // software pipeline over the 32-entry chunks: the entries of chunk k+2 and the sorted records of chunk k+1 are
            // in flight while chunk k is staged and its tiles are evaluated (the records are a dependent gather)
            const int eEnd = min(listN, chunkHi*32);
            auto loadEntry = [&](int e) -> unsigned int {          // front part, then the back part (stored downwards from the end)
                return e < eEnd ? lst[e < listFront ? e : p.listCap - 1 - (e - listFront)] : 0xffffffffu;
            };
            auto wanted = [&](unsigned int ent) -> bool { return ent != 0xffffffffu && (!HALF || (int) (ent & 0x7ffffffu) > i0); };
            unsigned int entCur = loadEntry(chunkLo*32 + lane), entNext = loadEntry(chunkLo*32 + 32 + lane);
            float4 l4Cur = farAway, mjCur = farAway;
            if (wanted(entCur)) { l4Cur = p.sortedLocal[entCur & 0x7ffffffu]; mjCur = p.sortedMeta[entCur & 0x7ffffffu]; }
            for (int eb = chunkLo*32; eb < chunkHi*32; eb += 32) {
                float4 l4Next = farAway, mjNext = farAway;
                if (wanted(entNext)) { l4Next = p.sortedLocal[entNext & 0x7ffffffu]; mjNext = p.sortedMeta[entNext & 0x7ffffffu]; }
                const unsigned int entAfter = loadEntry(eb + 64 + lane);
                bool pass = false, cls = false;
                float4 pj = farAway, mj = farAway;
                double4 pjD = make_double4(0.0, 0.0, 0.0, 0.0);
                int s = 0;
                if (wanted(entCur)) {
                    const unsigned int ent = entCur;
                    s = (int) (ent & 0x7ffffffu);
                    {
                        const int code = (int) (ent >> 27);
                        const float4 l4 = l4Cur;
                        mj = mjCur;
                        const int cj = __float_as_int(mj.w);
                        const int cz3 = code/9, cy3 = (code - 9*cz3)/3, cx3 = code - 9*cz3 - 3*cy3;
                        const int offx = (cj & CELL_MASK) - c0x + (cx3 - 1)*p.ncx;
                        const int offy = ((cj >> CELL_BITS) & CELL_MASK) - c0y + (cy3 - 1)*p.ncy;
                        const int offz = (cj >> (2*CELL_BITS)) - c0z + (cz3 - 1)*p.ncz;
                        pj = make_float4(l4.x + offx*p.csx, l4.y + offy*p.csy, fmaf((float) offz, p.csz, l4.z), l4.w);
                        // the list was built for cutoff + skin around the bounding box of that time: keep what is within
                        // the cutoff of the box now
                        const float ex = fmaxf(0.f, fmaxf(bminx - pj.x, pj.x - bmaxx));
                        const float ey = fmaxf(0.f, fmaxf(bminy - pj.y, pj.y - bmaxy));
                        const float ez = fmaxf(0.f, fmaxf(bminz - pj.z, pj.z - bmaxz));
                        pass = fmaf(ez, ez, fmaf(ey, ey, ex*ex)) <= rcut2;
                        if (EMODE == 2 && pass) {
                            const double4 ld = p.sortedLocalD[s];
                            pjD = make_double4(ld.x + offx*p.dcsx, ld.y + offy*p.dcsy, ld.z + offz*p.dcsz, ld.w);
                        }
                        cls = anyLJi && mj.y != 0.f;
                    }
                }
                entCur = entNext; entNext = entAfter; l4Cur = l4Next; mjCur = mjNext;
                const unsigned int m1 = __ballot_sync(0xffffffffu, pass && cls);
                const unsigned int m0 = __ballot_sync(0xffffffffu, pass && !cls);
                if (pass) {
                    const int slot = cls ? ((base1 + count1 + __popc(m1 & lt)) & (P_JCAP - 1)) : ((base0 + count0 + __popc(m0 & lt)) & (P_JCAP - 1));
                    float* er = &sRing[warp][cls][0][slot];
                    er[0] = pj.x; er[P_JCAP] = pj.y; er[2*P_JCAP] = pj.z; er[3*P_JCAP] = pj.w; er[4*P_JCAP] = mj.z;
                    if (cls) { er[5*P_JCAP] = mj.x; er[6*P_JCAP] = mj.y; }
                    if (EMODE == 2) {
                        er[7*P_JCAP] = __int_as_float(s);
                        double* ed = &sRingD[warp][cls][0][slot];
                        ed[0] = pjD.x; ed[P_JCAP] = pjD.y; ed[2*P_JCAP] = pjD.z; ed[3*P_JCAP] = pjD.w;
                        if (cls) { const double2 ljj = p.sortedLjD[s]; ed[4*P_JCAP] = ljj.x; ed[5*P_JCAP] = ljj.y; }
                    }
                }
                count0 += __popc(m0); count1 += __popc(m1);
                if (EMODE == 2) {
                    // (list order is stencil order, not sorted order: the largest index of the chunk, by warp reduction)
                    const int sm0 = __reduce_max_sync(0xffffffffu, (pass && !cls) ? s : -1);
                    const int sm1 = __reduce_max_sync(0xffffffffu, (pass && cls) ? s : -1);
                    if (m0) { lastHi0 = sm0; hi0 = max(hi0, sm0); }
                    if (m1) { lastHi1 = sm1; hi1 = max(hi1, sm1); }
                }
                flush(false);
            }
        }
        else {
            // walk the stencil columns (x outer, y inner); column `col` belongs to share col % jSplits
            int cxw = FAST ? wrapOnce(c0x + loX, p.ncx) : modPos(c0x + loX, p.ncx);
            const int cyw0 = FAST ? wrapOnce(c0y + loY, p.ncy) : modPos(c0y + loY, p.ncy);
            int shareCtr = 0;
            for (int ax = 0; ax < nX; ax++) {
                const float shx = (loX + ax)*p.csx;
                // (atoms sorted at the last list build may be up to p.drift outside their cell)
                const float ex0 = fmaxf(0.f, fmaxf(shx - bmaxx, bminx - (shx + p.csx)) - p.drift);
                int cyw = cyw0;
                for (int ay = 0; ay < nY; ay++) {
                    const int cyCur = cyw;
                    const bool mine = shareCtr == share;
                    if (++cyw == p.ncy) cyw = 0;
                    if (++shareCtr == nShares) shareCtr = 0;
                    if (!mine) continue;
                    const float shy = (loY + ay)*p.csy;
                    // z range of this column: the cells cut by the sphere of radius rc around the bounding box
                    int zRel0 = loZ, zCount = nZ;
                    if (!minImage) {
                        const float ey0 = fmaxf(0.f, fmaxf(shy - bmaxy, bminy - (shy + p.csy)) - p.drift);
                        const float d2 = fmaf(ey0, ey0, ex0*ex0);
                        if (d2 > rcut2) continue;
                        const float dzMax = sqrtf(rcut2 - d2) + 1e-4f + p.drift;
                        const int za = max(loZ, (int) floorf((bminz - dzMax)*p.invCsz));
                        const int zb = min(loZ + nZ - 1, (int) floorf((bmaxz + dzMax)*p.invCsz));
                        zRel0 = za; zCount = zb - za + 1;
                        if (zCount <= 0) continue;
                    }
                    const int rowCell = (cxw*p.ncy + cyCur)*p.ncz;
                    // unwrapped cell units [zlo, zlo + zCount - 1] -> at most two contiguous wrapped segments
                    const int zlo = c0z + zRel0;
                    const int zloW = FAST ? wrapOnce(zlo, p.ncz) : modPos(zlo, p.ncz);
                    const int nSeg = (zloW + zCount - 1 < p.ncz) ? 1 : 2;
                    for (int sg = 0; sg < nSeg; sg++) {
                        const int segLo = sg == 0 ? zloW : 0;
                        const int segHi = sg == 0 ? min(zloW + zCount - 1, p.ncz - 1) : zloW + zCount - 1 - p.ncz;
                        const int zShift = (sg == 0 ? zlo - zloW : zlo - zloW + p.ncz) - c0z;
                        const int s1 = p.cellStart[rowCell + segHi + 1];
                        for (int sb = HALF ? max(p.cellStart[rowCell + segLo], i0 + 1) : p.cellStart[rowCell + segLo]; sb < s1; sb += 32) {
                            const int s = sb + lane;
                            bool pass = false, cls = false;
                            float4 pj = farAway, mj = farAway;
                            double4 pjD = make_double4(0.0, 0.0, 0.0, 0.0);
                            if (s < s1) {
                                const float4 l4 = p.sortedLocal[s];
                                mj = p.sortedMeta[s];
                                const int cz = __float_as_int(mj.w) >> (2*CELL_BITS);
                                pj = make_float4(l4.x + shx, l4.y + shy, fmaf((float) (cz + zShift), p.csz, l4.z), l4.w);
                                if (EMODE == 2) {
                                    const double4 ld = p.sortedLocalD[s];
                                    pjD = make_double4(ld.x + (loX + ax)*p.dcsx, ld.y + (loY + ay)*p.dcsy, ld.z + (cz + zShift)*p.dcsz, ld.w);
                                }
                                if (minImage) pass = true;
                                else {
                                    const float ex = fmaxf(0.f, fmaxf(bminx - pj.x, pj.x - bmaxx));
                                    const float ey = fmaxf(0.f, fmaxf(bminy - pj.y, pj.y - bmaxy));
                                    const float ez = fmaxf(0.f, fmaxf(bminz - pj.z, pj.z - bmaxz));
                                    pass = fmaf(ez, ez, fmaf(ey, ey, ex*ex)) <= rcut2;
                                }
                                cls = anyLJi && mj.y != 0.f;
                            }
                            const unsigned int m1 = __ballot_sync(0xffffffffu, pass && cls);
                            const unsigned int m0 = __ballot_sync(0xffffffffu, pass && !cls);
                            if (pass) {
                                const int slot = cls ? ((base1 + count1 + __popc(m1 & lt)) & (P_JCAP - 1)) : ((base0 + count0 + __popc(m0 & lt)) & (P_JCAP - 1));
                                float* e = &sRing[warp][cls][0][slot];
                                e[0] = pj.x; e[P_JCAP] = pj.y; e[2*P_JCAP] = pj.z; e[3*P_JCAP] = pj.w; e[4*P_JCAP] = mj.z;
                                if (cls) { e[5*P_JCAP] = mj.x; e[6*P_JCAP] = mj.y; }
                                if (EMODE == 2) {
                                    e[7*P_JCAP] = __int_as_float(s);
                                    double* ed = &sRingD[warp][cls][0][slot];
                                    ed[0] = pjD.x; ed[P_JCAP] = pjD.y; ed[2*P_JCAP] = pjD.z; ed[3*P_JCAP] = pjD.w;
                                    if (cls) { const double2 ljj = p.sortedLjD[s]; ed[4*P_JCAP] = ljj.x; ed[5*P_JCAP] = ljj.y; }
                                }
                            }
                            count0 += __popc(m0); count1 += __popc(m1);
                            if (EMODE == 2) {
                                if (m0) { lastHi0 = sb + 31 - __clz(m0); hi0 = max(hi0, lastHi0); }
                                if (m1) { lastHi1 = sb + 31 - __clz(m1); hi1 = max(hi1, lastHi1); }
                            }
                            nCand += (unsigned int) __popc(m0 | m1);
                            flush(false);
                        }
                    }
                }
                if (++cxw == p.ncx) cxw = 0;
            }
        }
        flush(true);

        // reduce over the 4 lanes of each i atom (warp shuffles), one fixed-point atomic per output
        if (FORCES) {
            float fx = fx2.x + fx2.y, fy = fy2.x + fy2.y, fz = fz2.x + fz2.y, dq = dq2.x + dq2.y;
            #pragma unroll
            for (int o = 1; o <= 2; o <<= 1) {
                fx += __shfl_xor_sync(0xffffffffu, fx, o); fy += __shfl_xor_sync(0xffffffffu, fy, o);
                fz += __shfl_xor_sync(0xffffffffu, fz, o); dq += __shfl_xor_sync(0xffffffffu, dq, o);
            }
            if (validI && part == 0) {
                atomicAddFixed(p.forceFixed + ui, (double) fx);
                atomicAddFixed(p.forceFixed + p.Npad + ui, (double) fy);
                atomicAddFixed(p.forceFixed + 2*(size_t) p.Npad + ui, (double) fz);
                atomicAddFixed(p.dedqFixed + ui, (double) ((float) CFX_ONE_4PI_EPS0*dq));
            }
        }
        if (EMODE != 0) {
            if (EMODE == 1) en = (double) (enf2.x + enf2.y);
            en = warpSum(en);
            // FP32 terms: every pair is seen from both sides. FP64 terms: counted once (from the atom first in sorted order).
            if (lane == 0) atomicAddEnergy(p.energyFixed + CFX_E_DIRECT, EMODE == 2 ? en : 0.5*en);
        }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) nIn += __shfl_xor_sync(0xffffffffu, nIn, o);
        if (lane == 0 && p.countStats) {
            atomicAdd(p.counters + 0, (unsigned long long) nIn);        // ordered pairs: the host halves it
            if (!FAST) atomicAdd(p.counters + 1, (unsigned long long) nCand*P_ITILE);
        }
        __syncwarp();
    }
}

// Candidate lists of the fast path. One warp per i-cluster walks the (x, y) columns of the cluster's 5x5 stencil, clips
// each column's z range to the sphere of radius rc around the cluster's bounding box, tests the column's atoms against
// the bounding box 32 at a time and appends the ones within the cutoff of the box, ballot-compacted, to the cluster's list:
// entry = sorted index | image code << 27, code = (sx+1) + 3 (sy+1) + 9 (sz+1) with s the periodic image (in boxes) of the
// atom's cell as seen from the cluster. The pair passes of the evaluation (force pass, energy pass, pair emission) read
// the list instead of repeating the search. Clusters whose stencil would wrap onto itself, or whose list outgrows
// listCap (overflow is counted in counters[11]; the host enlarges the lists for the next evaluation), are put on wrapList
// and evaluated by the generic pair kernel.
__global__ void __launch_bounds__(P_WARPS*32) buildListKernel(const __grid_constant__ PairParams p) {
    if (!*p.rebuildFlag) return;
    const int lane = threadIdx.x & 31;
    const unsigned int lt = (1u << lane) - 1u;
    const float rcut2 = p.rlist2*1.0001f;
    const unsigned int totalItems = (unsigned int) (p.groupHi - p.groupLo);
    for (;;) {
        unsigned int item = 0;
        if (lane == 0) item = atomicAdd(p.workCounter, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= totalItems) break;
        const int g = p.groupLo + (int) item;
        const int i0 = g*P_ITILE;
        const int ni = min(P_ITILE, p.N - i0);
        const int ii = lane & 7;
        const int iIdx = i0 + (ii < ni ? ii : 0);
        const int c0 = __float_as_int(p.sortedMeta[i0].w);
        const int c0x = c0 & CELL_MASK, c0y = (c0 >> CELL_BITS) & CELL_MASK, c0z = c0 >> (2*CELL_BITS);
        const float4 li = p.sortedLocal[iIdx];
        const int ci = __float_as_int(p.sortedMeta[iIdx].w);
        const int ox = wrapNearest((ci & CELL_MASK) - c0x, p.ncx);
        const int oy = wrapNearest(((ci >> CELL_BITS) & CELL_MASK) - c0y, p.ncy);
        const int oz = wrapNearest((ci >> (2*CELL_BITS)) - c0z, p.ncz);
        const float pix = li.x + ox*p.csx, piy = li.y + oy*p.csy, piz = li.z + oz*p.csz;
        float bminx = pix, bminy = piy, bminz = piz, bmaxx = pix, bmaxy = piy, bmaxz = piz;
        int ominx = ox, ominy = oy, ominz = oz, omaxx = ox, omaxy = oy, omaxz = oz;
        #pragma unroll
        for (int o = 4; o > 0; o >>= 1) {               // lanes l and l + 8 hold the same atom
            bminx = fminf(bminx, __shfl_xor_sync(0xffffffffu, bminx, o)); bmaxx = fmaxf(bmaxx, __shfl_xor_sync(0xffffffffu, bmaxx, o));
            bminy = fminf(bminy, __shfl_xor_sync(0xffffffffu, bminy, o)); bmaxy = fmaxf(bmaxy, __shfl_xor_sync(0xffffffffu, bmaxy, o));
            bminz = fminf(bminz, __shfl_xor_sync(0xffffffffu, bminz, o)); bmaxz = fmaxf(bmaxz, __shfl_xor_sync(0xffffffffu, bmaxz, o));
            ominx = min(ominx, __shfl_xor_sync(0xffffffffu, ominx, o)); omaxx = max(omaxx, __shfl_xor_sync(0xffffffffu, omaxx, o));
            ominy = min(ominy, __shfl_xor_sync(0xffffffffu, ominy, o)); omaxy = max(omaxy, __shfl_xor_sync(0xffffffffu, omaxy, o));
            ominz = min(ominz, __shfl_xor_sync(0xffffffffu, ominz, o)); omaxz = max(omaxz, __shfl_xor_sync(0xffffffffu, omaxz, o));
        }
        const int loX = ominx - 2, nX = omaxx - ominx + 5;
        const int loY = ominy - 2, nY = omaxy - ominy + 5;
        const int loZ = ominz - 2, nZ = omaxz - ominz + 5;
        unsigned int* lst = p.pairList + (size_t) item*p.listCap;
        int count = 0, countB = 0;
        bool giveUp = nX > p.ncx || nY > p.ncy || nZ > p.ncz;      // the stencil would wrap onto itself
        unsigned int nCand = 0;
        int cxw = wrapOnce(c0x + loX, p.ncx);
        const int cyw0 = wrapOnce(c0y + loY, p.ncy);
        for (int ax = 0; ax < nX && !giveUp; ax++) {
            const float shx = (loX + ax)*p.csx;
            const float ex0 = fmaxf(0.f, fmaxf(shx - bmaxx, bminx - (shx + p.csx)));
            const int ux = c0x + loX + ax;
            const int codeX = ux < 0 ? 0 : (ux >= p.ncx ? 2 : 1);
            int cyw = cyw0;
            for (int ay = 0; ay < nY && !giveUp; ay++) {
                const int cyCur = cyw;
                if (++cyw == p.ncy) cyw = 0;
                const float shy = (loY + ay)*p.csy;
                const float ey0 = fmaxf(0.f, fmaxf(shy - bmaxy, bminy - (shy + p.csy)));
                const float d2 = fmaf(ey0, ey0, ex0*ex0);
                if (d2 > rcut2) continue;
                const float dzMax = sqrtf(rcut2 - d2) + 1e-4f;
                const int za = max(loZ, (int) floorf((bminz - dzMax)*p.invCsz));
                const int zb = min(loZ + nZ - 1, (int) floorf((bmaxz + dzMax)*p.invCsz));
                const int zCount = zb - za + 1;
                if (zCount <= 0) continue;
                const int uy = c0y + loY + ay;
                const int codeXY = codeX + 3*(uy < 0 ? 0 : (uy >= p.ncy ? 2 : 1));
                const int rowCell = (cxw*p.ncy + cyCur)*p.ncz;
                const int zlo = c0z + za;
                const int zloW = wrapOnce(zlo, p.ncz);
                const int nSeg = (zloW + zCount - 1 < p.ncz) ? 1 : 2;
                for (int sg = 0; sg < nSeg && !giveUp; sg++) {
                    const int segLo = sg == 0 ? zloW : 0;
                    const int segHi = sg == 0 ? min(zloW + zCount - 1, p.ncz - 1) : zloW + zCount - 1 - p.ncz;
                    const int zImage = (sg == 0 ? zlo - zloW : zlo - zloW + p.ncz);        // multiple of ncz: -ncz, 0 or ncz
                    const int zShift = zImage - c0z;
                    const unsigned int code = (unsigned int) (codeXY + 9*(zImage < 0 ? 0 : (zImage > 0 ? 2 : 1))) << 27;
                    const int s1 = p.cellStart[rowCell + segHi + 1];
                    for (int sb = p.cellStart[rowCell + segLo]; sb < s1; sb += 32) {
                        const int s = sb + lane;
                        bool pass = false;
                        if (s < s1) {
                            const float4 l4 = p.sortedLocal[s];
                            const int cz = __float_as_int(p.sortedMeta[s].w) >> (2*CELL_BITS);
                            const float x = l4.x + shx, y = l4.y + shy, z = fmaf((float) (cz + zShift), p.csz, l4.z);
                            const float ex = fmaxf(0.f, fmaxf(bminx - x, x - bmaxx));
                            const float ey = fmaxf(0.f, fmaxf(bminy - y, y - bmaxy));
                            const float ez = fmaxf(0.f, fmaxf(bminz - z, z - bmaxz));
                            pass = fmaf(ez, ez, fmaf(ey, ey, ex*ex)) <= rcut2;
                        }
                        // atoms behind the cluster in sorted order fill the list from the front, the others from the back:
                        // energy passes (each pair once, from the atom that comes first) read only the front part
                        const bool behind = s > i0;
                        const unsigned int mA = __ballot_sync(0xffffffffu, pass && behind);
                        const unsigned int mB = __ballot_sync(0xffffffffu, pass && !behind);
                        const int kA = __popc(mA), kB = __popc(mB);
                        if (count + countB + kA + kB > p.listCap) { giveUp = true; break; }
                        if (pass) {
                            if (behind) lst[count + __popc(mA & lt)] = (unsigned int) s | code;
                            else        lst[p.listCap - 1 - (countB + __popc(mB & lt))] = (unsigned int) s | code;
                        }
                        count += kA; countB += kB;
                    }
                }
            }
            if (++cxw == p.ncx) cxw = 0;
        }
        nCand = (unsigned int) (count + countB);
        if (lane == 0) {
            if (giveUp) {
                p.wrapList[atomicAdd(p.wrapCount, 1ull)] = g;
                if (!(nX > p.ncx || nY > p.ncy || nZ > p.ncz)) atomicAdd(p.counters + 11, 1ull);     // list overflow
            }
            p.listCount[item] = giveUp ? make_int2(-1, 0) : make_int2(count, countB);
            atomicMax(p.counters + 12, (unsigned long long) (count + countB));                                   // longest list so far
            atomicAdd(p.counters + 1, (unsigned long long) nCand*P_ITILE);
        }
        __syncwarp();
    }
}

template <bool FAST, bool EMIT>
void dispatchPair(const PairParams& pp, bool forces, int emode, int blocks, cudaStream_t s) {
    const int t = P_WARPS*32;
    if (forces) {
        if (emode == 2)      pairKernel<FAST, true, 2, EMIT><<<blocks, t, 0, s>>>(pp);
        else if (emode == 1) pairKernel<FAST, true, 1, EMIT><<<blocks, t, 0, s>>>(pp);
        else                 pairKernel<FAST, true, 0, EMIT><<<blocks, t, 0, s>>>(pp);
    }
    else {
        if (emode == 2)      pairKernel<FAST, false, 2, EMIT><<<blocks, t, 0, s>>>(pp);
        else if (emode == 1) pairKernel<FAST, false, 1, EMIT><<<blocks, t, 0, s>>>(pp);
        else                 pairKernel<FAST, false, 0, EMIT><<<blocks, t, 0, s>>>(pp);
    }
}

} // namespace

// FP64 pair energies: erfc(alpha r)/r = 1/r - alpha G(alpha^2 s), s = r^2, G(z) = erf(sqrt z)/sqrt z, which is entire in z.
// G is interpolated at the Chebyshev nodes of [0, Z], Z = alpha^2 rc^2 (1 + 1e-3), by one polynomial of degree E_POLY_DEG
// in t = 2 z/Z - 1 (monomial form: its coefficients sum to ~1.1 in magnitude, Horner is well conditioned); the fit is
// checked here against erfl and, should it miss 1e-10 (alpha rc > 4, i.e. Ewald tolerances below 6e-8), the kernel
// falls back to libm's erfc.
static void fitEnergyPolynomial(const State& st, double* coef, double* tScale, int* ok) {
    const int n = E_POLY_DEG + 1;
    const long double pi = 3.14159265358979323846264338327950288L;
    const long double Z = (long double) st.alpha*st.alpha*st.cutoff*st.cutoff*1.001L;
    auto G = [](long double z) -> long double {
        if (z < 1e-6L) return 2.0L/sqrtl(3.14159265358979323846264338327950288L)*(1.0L - z/3.0L + z*z/10.0L);
        const long double r = sqrtl(z);
        return erfl(r)/r;
    };
    std::vector<long double> f(n), c(n);
    for (int j = 0; j < n; j++) f[j] = G(0.5L*Z*(1.0L + cosl(pi*(j + 0.5L)/n)));
    for (int k = 0; k < n; k++) {
        long double a = 0;
        for (int j = 0; j < n; j++) a += f[j]*cosl(pi*k*(j + 0.5L)/n);
        c[k] = a*(k == 0 ? 1.0L : 2.0L)/n;
    }
    std::vector<long double> mono(n, 0.0L), tkm(n, 0.0L), tk(n, 0.0L), tn(n);
    tkm[0] = 1.0L;                       // T_0
    tk[1] = 1.0L;                        // T_1
    for (int i = 0; i < n; i++) mono[i] += c[0]*tkm[i] + (n > 1 ? c[1]*tk[i] : 0.0L);
    for (int k = 2; k < n; k++) {        // T_k = 2 t T_{k-1} - T_{k-2}
        for (int i = 0; i < n; i++) tn[i] = (i > 0 ? 2.0L*tk[i-1] : 0.0L) - tkm[i];
        for (int i = 0; i < n; i++) mono[i] += c[k]*tn[i];
        tkm = tk; tk = tn;
    }
    for (int i = 0; i < n; i++) coef[i] = (double) mono[i];
    *tScale = (double) (2.0L*st.alpha*st.alpha/Z);
    long double worst = 0;
    for (int j = 0; j <= 4000; j++) {
        const long double z = Z*j/4000.0L, t = 2.0L*z/Z - 1.0L;
        long double v = coef[n-1];
        for (int k = n - 2; k >= 0; k--) v = v*t + coef[k];
        worst = std::max(worst, fabsl(v - G(z)));
    }
    *ok = worst < 1e-10L ? 1 : 0;
}

// candidate lists of this rank's i-clusters (fast path only), listCap entries each
void allocPairLists(State& st) {
    cudaFree(st.pairList); cudaFree(st.listCount);
    st.pairList = nullptr; st.listCount = nullptr; st.pairListEntries = 0;
    if (st.cells.smallBox) return;
    const int numGroups = (st.N + P_ITILE - 1)/P_ITILE;
    const int lo = (int) ((int64_t) numGroups*st.shardRank/st.shardCount), hi = (int) ((int64_t) numGroups*(st.shardRank + 1)/st.shardCount);
    const size_t groups = (size_t) std::max(hi - lo, 1);
    CFX_CUDA(cudaMalloc(&st.pairList, sizeof(unsigned int)*groups*st.listCap));
    CFX_CUDA(cudaMalloc(&st.listCount, sizeof(int2)*groups));
    st.pairListEntries = groups*st.listCap;
}

// The skin only applies where the fast (list) path runs: at least 7 cells of edge >= (cutoff + skin)/2 per axis, so
// that the 5-cell stencil covers cutoff + skin around a cluster and does not wrap onto itself.
static double effectiveSkin(const State& st) {
    for (int d = 0; d < 3; d++)
        if ((int) floor(st.box.L[d]/(0.5*(st.cutoff + st.skin))) < 7) return 0.0;
    return st.skin;
}

int cellsPerAxis(const State& st, int d) {
    const int n = (int) floor(st.box.L[d]/(0.5*(st.cutoff + effectiveSkin(st))));
    return std::max(1, std::min(n, CELL_MASK));
}

void invalidatePairLists(State& st) {
    if (!st.rebuildFlag) return;
    CFX_CUDA(cudaDeviceSynchronize());
    CFX_CUDA(cudaMemset(st.rebuildFlag, 1, sizeof(int)));
}

void planCells(State& st) {
    CellPlan& c = st.cells;
    c.smallBox = false;
    c.ncells = 1;
    for (int d = 0; d < 3; d++) {
        int n = cellsPerAxis(st, d);
        c.nc[d] = n;
        c.csd[d] = st.box.L[d]/n;
        c.cs[d] = (float) c.csd[d];
        c.ncells *= n;
        if (n < 7) c.smallBox = true;      // the 5-cell stencil wraps onto itself: generic pair kernel (per-pair min image)
    }
    CFX_CUDA(cudaMalloc(&st.cellOfAtom, sizeof(int)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.cellCount, sizeof(int)*(c.ncells + 1)));
    CFX_CUDA(cudaMalloc(&st.cellStart, sizeof(int)*(c.ncells + 1)));
    CFX_CUDA(cudaMalloc(&st.cellFill, sizeof(int)*(c.ncells + 1)));
    CFX_CUDA(cudaMalloc(&st.userLocal, sizeof(float4)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.sortedLocal, sizeof(float4)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.sortedMeta, sizeof(float4)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.filledUser, sizeof(int)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.userLocalD, sizeof(double4)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.sortedLocalD, sizeof(double4)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.sortedLjD, sizeof(double2)*st.Npad));
    static_assert(E_POLY_DEG + 1 <= sizeof(st.ePoly)/sizeof(double), "State::ePoly too small");
    fitEnergyPolynomial(st, st.ePoly, &st.eTScale, &st.ePolyOK);
    CFX_CUDA(cudaMalloc(&st.wrapList, sizeof(int)*(st.Npad/P_ITILE + 1)));
    if (st.listCap <= 0) {
        st.listCap = 4096;                               // ~4x the candidates of a cluster in liquid water at rc = 1 nm
        if (const char* e = getenv("CFX_LIST_CAP")) st.listCap = std::max(32, atoi(e));
    }
    allocPairLists(st);
    CFX_CUDA(cudaMalloc(&st.rebuildFlag, sizeof(int)));
    CFX_CUDA(cudaMemset(st.rebuildFlag, 1, sizeof(int)));
    CFX_CUDA(cudaMalloc(&st.posAtBuild, sizeof(double)*3*st.Npad));
    CFX_CUDA(cudaMalloc(&st.pairCounters, sizeof(unsigned long long)*16));
    CFX_CUDA(cudaMemset(st.pairCounters, 0, sizeof(unsigned long long)*16));
}

// pairCounters (16 x u64): [0] in-cutoff ordered pairs, [1] distance tests, [2] emitted pairs; work-item counters of the
// fast / generic pair kernel: [5] / [6] (first pass), [8] / [9] (energy pass of an energy+forces call), [10] list builder;
// [7] clusters left to the generic kernel; [11] list overflows so far (not reset: the host enlarges listCap when it grows);
// [12] longest candidate list so far; [13] list builds so far.
// phase 0: the whole branch; 1: the part that needs only the positions (displacement check, re-sort, candidate lists);
// 2: the rest (charges into the sorted records, pair passes), after a phase-1 call of the same evaluation
void launchDirect(State& st, const double* dPos, bool forces, int emode, bool emitPairs, long long* dForce, long long* dDedq, cudaStream_t s,
                  int phase) {
    if (!forces && emode == 0 && !emitPairs) return;
    CellPlan& c = st.cells;
    CellParams cp{st.N, c.nc[0], c.nc[1], c.nc[2], c.ncells, 1.0/st.box.L[0], 1.0/st.box.L[1], 1.0/st.box.L[2], c.csd[0], c.csd[1], c.csd[2]};
    const double skin = c.smallBox ? 0.0 : effectiveSkin(st);
    if (phase != 2) {
    CFX_CUDA(cudaMemsetAsync(st.cellCount, 0, sizeof(int)*(c.ncells + 1), s));
    CFX_CUDA(cudaMemsetAsync(st.pairCounters + 5, 0, sizeof(unsigned long long)*2, s));
    CFX_CUDA(cudaMemsetAsync(st.pairCounters + 8, 0, sizeof(unsigned long long)*3, s));
    // Re-sort and rebuild the candidate lists only when an atom has moved skin/2 since the last build (always without a
    // skin / in small boxes). The flag lives on the device: the guarded kernels stay in the step's graph.
    if (skin > 0.0) {
        displacementKernel<<<(st.N + 255)/256, 256, 0, s>>>(st.N, dPos, st.posAtBuild, st.box.L[0], st.box.L[1], st.box.L[2],
                0.25*skin*skin, st.rebuildFlag);
        CFX_LAUNCH_CHECK(); st.launches++;
    }
    else
        CFX_CUDA(cudaMemsetAsync(st.rebuildFlag, 1, sizeof(int), s));
    cellAssignKernel<<<(st.N + 255)/256, 256, 0, s>>>(cp, st.rebuildFlag, dPos, st.qf, st.q, st.cellOfAtom, st.userLocal, st.userLocalD,
            st.cellCount, st.posAtBuild);
    CFX_LAUNCH_CHECK(); st.launches++;
    cellScanKernel<<<1, 1024, 0, s>>>(c.ncells, st.rebuildFlag, st.cellCount, st.cellStart, st.cellFill, st.pairCounters + 7);
    CFX_LAUNCH_CHECK(); st.launches++;
    cellFillKernel<<<(st.N + 255)/256, 256, 0, s>>>(st.N, st.rebuildFlag, st.cellOfAtom, st.cellFill, st.filledUser);
    CFX_LAUNCH_CHECK(); st.launches++;
    cellRankGatherKernel<<<(st.N + 255)/256, 256, 0, s>>>(st.N, c.nc[1], c.nc[2], st.rebuildFlag, st.filledUser, st.cellOfAtom, st.cellStart,
            st.userLocal, st.lj, st.sortedLocal, st.sortedMeta, st.userLocalD, st.ljd, st.sortedLocalD, st.sortedLjD);
    CFX_LAUNCH_CHECK(); st.launches++;
    refreshSortedKernel<<<(st.N + 255)/256, 256, 0, s>>>(cp, dPos, st.qf, st.q, st.sortedLocal, st.sortedMeta, st.sortedLocalD);
    CFX_LAUNCH_CHECK(); st.launches++;
    mark(st, "cell_build", s);
    }
    else {
        refreshChargesKernel<<<(st.N + 255)/256, 256, 0, s>>>(st.N, st.qf, st.q, st.sortedMeta, st.sortedLocal, st.sortedLocalD);
        CFX_LAUNCH_CHECK(); st.launches++;
    }

    PairParams pp;
    pp.N = st.N; pp.Npad = st.Npad;
    pp.numGroups = (st.N + P_ITILE - 1)/P_ITILE;
    // spatial sharding: contiguous ranges of i-clusters in cell order are spatial slabs
    pp.groupLo = (int) ((int64_t) pp.numGroups*st.shardRank/st.shardCount);
    pp.groupHi = (int) ((int64_t) pp.numGroups*(st.shardRank + 1)/st.shardCount);
    pp.ncx = c.nc[0]; pp.ncy = c.nc[1]; pp.ncz = c.nc[2];
    pp.csx = c.cs[0]; pp.csy = c.cs[1]; pp.csz = c.cs[2]; pp.invCsz = 1.0f/c.cs[2];
    pp.Lx = (float) st.box.L[0]; pp.Ly = (float) st.box.L[1]; pp.Lz = (float) st.box.L[2];
    pp.invLx = (float) (1.0/st.box.L[0]); pp.invLy = (float) (1.0/st.box.L[1]); pp.invLz = (float) (1.0/st.box.L[2]);
    pp.dLx = st.box.L[0]; pp.dLy = st.box.L[1]; pp.dLz = st.box.L[2];
    pp.rc2d = st.cutoff*st.cutoff;
    pp.rc2 = (float) pp.rc2d; pp.alpha = (float) st.alpha; pp.alpha2 = (float) (st.alpha*st.alpha); pp.band = (float) (1e-5*pp.rc2d);
    pp.exclMaxR2Bits = st.exclMaxR2;
    pp.sortedLocal = st.sortedLocal; pp.sortedMeta = st.sortedMeta;
    pp.cellStart = st.cellStart; pp.exclPtr = st.exclPtr; pp.exclCols = st.exclCols; pp.pos = dPos;
    pp.q = st.q; pp.ljd = st.ljd; pp.alphaD = st.alpha;
    pp.dInvLx = 1.0/st.box.L[0]; pp.dInvLy = 1.0/st.box.L[1]; pp.dInvLz = 1.0/st.box.L[2];
    pp.forceFixed = dForce; pp.dedqFixed = dDedq; pp.energyFixed = st.energyFixed;
    pp.wrapList = st.wrapList;
    pp.sortedLocalD = st.sortedLocalD; pp.sortedLjD = st.sortedLjD;
    for (int k = 0; k <= E_POLY_DEG; k++) pp.ePoly[k] = st.ePoly[k];
    pp.eTScale = st.eTScale; pp.ePolyOK = st.ePolyOK;
    pp.dcsx = c.csd[0]; pp.dcsy = c.csd[1]; pp.dcsz = c.csd[2];
    pp.rebuildFlag = st.rebuildFlag;
    pp.rlist2 = (float) ((st.cutoff + skin)*(st.cutoff + skin));
    pp.drift = (float) (0.5*skin);
    pp.counters = st.pairCounters; pp.pairBuffer = st.pairBuffer; pp.pairCapacity = (unsigned long long) st.pairCapacity;
    const int groups = pp.groupHi - pp.groupLo;
    if (groups <= 0) return;
    // Work items = (i-cluster, share of its stencil columns), handed to persistent warps by an atomic counter. One share
    // per cluster unless that leaves fewer than ~28 items per SM -- small systems, shards -- (the partial sums of the
    // shares meet in the fixed-point atomics).
    int numSM = 148;
    cudaDeviceGetAttribute(&numSM, cudaDevAttrMultiProcessorCount, st.device);
    pp.jSplits = std::max(1, std::min(P_MAX_JSPLITS, (28*numSM + groups - 1)/groups));
    if (const char* e = getenv("CFX_PAIR_JSPLITS")) pp.jSplits = std::max(1, std::min(32, atoi(e)));     // experiments
    if (emitPairs) pp.jSplits = 1;
    const int items = groups*pp.jSplits;
    const bool fast = !c.smallBox;
    // An energy+forces call runs two passes: the FP32 force pass (no energy: lean, 5 CTAs per SM) and an energy-only FP64
    // pass over the half shell. One fused pass was slower (0.295 ms against 0.15 + 0.07 at 32k atoms): the FP64 staging
    // costs the force loop its occupancy, and an energy needs each pair only once.
    pp.pairList = st.pairList; pp.listCount = st.listCount; pp.listCap = st.listCap;
    if (fast && phase != 2) {
        // candidate lists, shared by every pair pass of this evaluation
        pp.wrapList = st.wrapList; pp.wrapCount = st.pairCounters + 7;
        pp.workCounter = reinterpret_cast<unsigned int*>(st.pairCounters + 10);
        buildListKernel<<<std::min((groups + P_WARPS - 1)/P_WARPS, 8*numSM), P_WARPS*32, 0, s>>>(pp);
        CFX_LAUNCH_CHECK(); st.launches++;
        finishListKernel<<<1, 32, 0, s>>>(st.rebuildFlag, st.pairCounters);
        CFX_LAUNCH_CHECK(); st.launches++;
        mark(st, "pair_list", s);
    }
    if (phase == 1) return;
    // An energy+forces call runs two passes: the FP32 force pass (no energy: lean, 5 CTAs per SM) and an energy-only FP64
    // pass over the half shell. One fused pass was slower: the FP64 staging costs the force loop its occupancy, and an
    // energy needs each pair only once.
    struct Pass { bool forces; int emode; };
    Pass passes[2] = {{forces, emode}, {false, 2}};
    int nPass = 1;
    if (forces && emode == 2 && !emitPairs) { passes[0].emode = 0; nPass = 2; }
    pp.wrapList = st.wrapList; pp.wrapCount = st.pairCounters + 7;
    for (int k = 0; k < nPass; k++) {
        unsigned long long* ctr = st.pairCounters + (k == 0 ? 5 : 8);      // [0] fast work items, [1] generic work items
        pp.countStats = k == 0 ? 1 : 0;
        const bool f = passes[k].forces; const int em = passes[k].emode;
        if (fast) {
            pp.onlyMinImage = 0;
            pp.workCounter = reinterpret_cast<unsigned int*>(ctr);
            const int grid = std::min((items + P_WARPS - 1)/P_WARPS, (em == 2 ? 4 : P_FAST_MINBLOCKS)*numSM);
            if (emitPairs) dispatchPair<true, true>(pp, f, em, grid, s);
            else           dispatchPair<true, false>(pp, f, em, grid, s);
            CFX_LAUNCH_CHECK(); st.launches++;
        }
        // every cluster (small boxes), or the few the list builder left out in a large box -- a cluster stretched over a
        // sparse region, whose stencil wraps onto itself, or a list overflow -- (it exits at once when there are none)
        pp.onlyMinImage = fast ? 1 : 0;
        pp.workCounter = reinterpret_cast<unsigned int*>(ctr + 1);
        const int grid = std::min((items + P_WARPS - 1)/P_WARPS, 4*numSM);
        if (emitPairs) dispatchPair<false, true>(pp, f, em, grid, s);
        else           dispatchPair<false, false>(pp, f, em, grid, s);
        CFX_LAUNCH_CHECK(); st.launches++;
        mark(st, k == 0 ? "direct_pairs" : "direct_pairs_energy", s);      // force (or only) pass / FP64 energy pass
    }
}

} // namespace cfx

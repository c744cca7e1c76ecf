// cfx_internal.cuh -- shared declarations of the B200 charge-flux Ewald library.
//
// Data layout in HBM (all owned by the handle, see DESIGN.md section "HBM layout"):
//   user order : pos (double[3N]), q (double[N] + float[Npad]), Jacobian rows (double[3P]),
//                fixed-point accumulators force[3][Npad], dedq[Npad] (int64, value*2^32),
//   k-space    : per-atom phase rows, atom-major   rowS[atom] = {q*Ex[0..Kx), Ey[0..Ky), Ez[0..Kz)} (float2)
//                per-index phase columns, n-major  Ex/Ey[n][Npad], Z4[l][Npad] = (c, s, l*c, l*s)
//                structure-factor partials         part[split][row][col][8] (float)
//                gather coefficients               coef[signedRow][Kz] (float4: Ar, Ai, Br, Bi)
//   direct     : atoms sorted by cell: sortedLocal (float4: local xyz in cell, q), sortedMeta (float4: sigma/2,
//                2 sqrt(eps), user index, packed cell coordinates), cellStart[ncell+1]
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/cfx_b200.h"

#define CFX_FIXED_SCALE 4294967296.0            /* 2^32, OpenMM CUDA platform force convention */
#define CFX_ENERGY_SCALE 16777216.0             /* 2^24: |E| < 5.4e11 kJ/mol, 6e-8 resolution   */

namespace cfx {

// ---------------------------------------------------------------------------------------------
// error handling
// ---------------------------------------------------------------------------------------------
struct CudaError { cudaError_t code; const char* what; const char* file; int line; };
void throwCuda(cudaError_t code, const char* what, const char* file, int line);
#define CFX_CUDA(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) ::cfx::throwCuda(e__, #expr, __FILE__, __LINE__); } while (0)
#define CFX_LAUNCH_CHECK() CFX_CUDA(cudaGetLastError())

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ long long toFixed(double v) { return __double2ll_rn(v*CFX_FIXED_SCALE); }
__device__ __forceinline__ long long toFixedF(float v) { return __double2ll_rn((double) v*CFX_FIXED_SCALE); }
__device__ __forceinline__ void atomicAddFixed(long long* addr, double v) {
    atomicAdd(reinterpret_cast<unsigned long long*>(addr), static_cast<unsigned long long>(toFixed(v)));
}
__device__ __forceinline__ void atomicAddEnergy(long long* addr, double v) {
    atomicAdd(reinterpret_cast<unsigned long long*>(addr), static_cast<unsigned long long>(__double2ll_rn(v*CFX_ENERGY_SCALE)));
}
__device__ __forceinline__ double warpSum(double v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warpSumF(float v) {
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Block-wide sum of a double; result valid in thread 0. `scratch` must hold 32 doubles.
__device__ __forceinline__ double blockSum(double v, double* scratch) {
    v = warpSum(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? scratch[threadIdx.x] : 0.0;
    if (warp == 0) v = warpSum(v);
    return v;
}
#endif

// ---------------------------------------------------------------------------------------------
// per-handle state
// ---------------------------------------------------------------------------------------------
struct Box { double L[3]; double invL[3]; };

// geometry of one structure-factor kernel: |nz| slot layout of the per-atom phase rows, grid, shared memory
// st.energyFixed[8]: slots 0..3 are the energy components (2^24 fixed point); slot 7 holds the float bits of the largest
// |q| of the evaluation (atomicMax by the charge assembly, read by the integer structure-factor kernel)
constexpr int CFX_SLOT_QMAX = 7;

struct SGeom {
    int TN = 7, NC = 0;          // TN columns per warp, NC column groups (FP32 kernel); |nz| = l lives in slot (l/TN)*TNP + l%TN
    int kzPad = 0, rowPitch = 0; // padded |nz| slots; float2 per atom row in rowS: Kx + Ky + kzPad
    int threads = 128, rowTiles = 0, splits = 0, atomsPerSplit = 0, stages = 3;
    size_t smem = 0;
};

struct KSpacePlan {
    int K[3] = {0, 0, 0};        // kmax per axis (reference: nk in [0,K) / (-K,K))
    int numRows = 0;             // unsigned rows (nx, |ny|): Kx*Ky
    int numSignedRows = 0;       // gather rows (nx, ny): Ky + (Kx-1)*(2Ky-1)
    // structure-factor kernels: FP32 CUDA-core kernel (used whenever the reciprocal energy is requested: it sums with
    // round-to-nearest) and the tensor-core kernel (forces-only evaluations; the tensor core truncates on accumulation,
    // which biases |S|^2 by ~2e-7)
    SGeom sF, sT;
    bool fp32S = false;
    // shard of the unsigned rows this rank owns [rowLo, rowHi) (k-vector sharding)
    int rowLo = 0, rowHi = 0;
    int signedLo = 0, signedHi = 0;
    // gather geometry
    int gThreads = 384, gAtoms = 256, gRowsPerWarp = 2, gBuffers = 2, gRowsPerTile = 64, gRowSplits = 1;
    size_t gSmem = 0;
    // tensor-core structure factors (kspace_tc.cu)
    bool tensorS = false;
    uint32_t tsRowStageBytes = 0, tsOffA = 0, tsOffB = 0, tsOffBar = 0;
    // integer tensor-core structure factors (exact; energy and forces-only calls)
    bool i8S = false;
    uint32_t siRowStagePad = 0;
    // tensor-core gather (kspace_tc.cu): K padded to 8, atom tiles per work unit, columns per coefficient tile
    bool tensorGather = false;
    int tKp = 0, tKC = 0, tMT = 2, tNT = 128, tStages = 0, tColTiles = 0;
    uint32_t tStageBytes = 0, tOffEy = 0, tOffBar = 0;
    size_t tSmem = 0;
};

struct CellPlan {
    int nc[3] = {1, 1, 1};
    int ncells = 1;
    int lo[3] = {0, 0, 0}, nd[3] = {1, 1, 1};   // stencil offsets per axis: lo .. lo+nd-1
    bool smallBox = false;                      // informational: tiles fall back to per-pair min image
    float cs[3] = {1, 1, 1};
    double csd[3] = {1, 1, 1};
};

struct State {
    // sizes
    int N = 0, Npad = 0;
    int nb = 0, na = 0, nw = 0, P = 0, numSlots = 0, numTerms = 0;
    int numExcl = 0;
    bool pbc = false;
    double cutoff = 0, tol = 0, alpha = 0;
    int device = 0, shardRank = 0, shardCount = 1;
    bool useGraph = true;
    void* comm = nullptr;               // ncclComm_t of a sharded handle (comm.cu), owned
    long long* reduceBuf = nullptr;     // [3*Npad + 8] reduction buffer of the host-buffer sharded path
    bool skipDiscardedEnergy = false;   // cfx_options.flags & CFX_OPT_SKIP_DISCARDED_ENERGY
    bool pinCallerBuffers = false;      // cfx_options.flags & CFX_OPT_PIN_CALLER_BUFFERS
    bool kmaxFollowsBox = false;        // cfx_options.flags & CFX_OPT_KMAX_FOLLOWS_BOX
    int siForceDigits = 3;              // digit planes of the integer structure factors in the forces-only call (api.cu: forceDigitsFor)
    bool hostCopyKernels = true;        // host path: positions fetched / results stored by kernels on page-locked memory, no copy nodes
    KSpacePlan ks;
    CellPlan cells;
    int64_t numKVectors = 0;

    cudaStream_t stream = nullptr;      // owned stream for the host-buffer path
    cudaStream_t sideStream = nullptr;  // direct-space branch runs here, forked/joined inside the step (and its graph)
    cudaEvent_t evFork = nullptr, evJoin = nullptr, evStart = nullptr;
    bool overlapBranches = true;
    // parameters (device)
    double* q0 = nullptr;
    float2* lj = nullptr;               // (sigma/2, 2*sqrt(eps)) per user atom, FP32 pair kernel
    double2* ljd = nullptr;             // same in FP64 (energy, exclusion and non-periodic kernels)
    int* termIdx = nullptr;             // [3*numTerms] particle indices (bond: p1,p2,-1)
    double* termPar = nullptr;          // [5*numTerms]
    int* qcsrPtr = nullptr; int* qcsrSlot = nullptr; double* qcsrCoef = nullptr;
    int* rowDq = nullptr; int* rowDx = nullptr;
    int2* exclPairs = nullptr;          // unique (p1<p2)
    int* exclPtr = nullptr; int* exclCols = nullptr;   // symmetric CSR, sorted
    // per-evaluation (device)
    double* pos = nullptr;              // [3N] staging for the host path
    double* dqSlot = nullptr;           // [numSlots]
    double* rowVal = nullptr;           // [3P]
    double* q = nullptr; float* qf = nullptr;
    long long* forceFixed = nullptr;    // [3*Npad] internal accumulator (host path)
    long long* dedqFixed = nullptr;     // [Npad]
    long long* energyFixed = nullptr;   // [8]
    double* forceOut = nullptr;         // [3N] double, user order (host path)
    double* energyOut = nullptr;        // [CFX_E_COUNT]
    // k-space
    float2* rowS = nullptr;             // [Npad][rowPitch]
    float2* colX = nullptr; float2* colY = nullptr;   // [Kx][Npad], [Ky][Npad]
    float4* colZ4 = nullptr;            // [Kz][Npad]
    float* sPart = nullptr;             // [splits][rowsPad][kzPad][8]
    float4* gCoef = nullptr;            // [numSignedRows][Kz]
    int2* gRowInfo = nullptr;           // [numSignedRows] (nx, ny)
    int* ks_signedStart = nullptr;      // [numRows+1] first signed row of each unsigned row
    unsigned long long* gtTrace = nullptr;   // debug: per-CTA phase timestamps of the tensor gather (CFX_GT_TRACE)
    int4* gGroupInfo = nullptr;         // per 8 signed rows from signedLo: Ex offsets of the first / last nx, rows with the first nx
    float4* gRowData = nullptr;         // [numSignedRows + pad] (nx, ny, |ny|*Ey stride, sign) for the tensor gather epilogue
    float* zSplit = nullptr;            // [Npad/128][hi|lo][Kp/4][128][4]  TF32 split of (cos, sin)(2 pi l z), tensor gather operand
    float* coefT = nullptr;             // [column tile][hi|lo][Kp/4][NT][4]  TF32 split of the gather coefficients, core-matrix layout
    // direct space
    int* cellOfAtom = nullptr; int* cellCount = nullptr; int* cellStart = nullptr; int* cellFill = nullptr;
    float4* userLocal = nullptr;        // [N] local xyz in own cell + q, user order (scratch of the cell build)
    float4* sortedLocal = nullptr;      // [N] atoms sorted by cell, z inside the cell: local xyz + q
    float4* sortedMeta = nullptr;       // [N] (sigma/2, 2 sqrt(eps), user index, packed cell coordinates)
    double4* userLocalD = nullptr; double4* sortedLocalD = nullptr;   // the same in double (+ charge): FP64 pair energies
    double2* sortedLjD = nullptr;       // [N] (sigma/2, 2 sqrt(eps)) in double, sorted order
    double ePoly[25] = {0}; double eTScale = 0; int ePolyOK = 0;   // polynomial of the FP64 pair energies (direct.cu)
    unsigned int* pairList = nullptr; int2* listCount = nullptr;     // candidate lists of the fast pair kernel (direct.cu)
    size_t pairListEntries = 0; int listCap = 0;
    double skin = 0.0;                  // lists are built for cutoff + skin and reused until an atom has moved skin/2 (0: rebuilt every evaluation)
    int* rebuildFlag = nullptr;         // device: non-zero = this evaluation re-sorts the atoms and rebuilds the lists
    double* posAtBuild = nullptr;       // [3N] positions the lists were built from
    unsigned long long listOverflowSeen = 0;
    unsigned long long* hListOverflow = nullptr;        // pinned copy of pairCounters[11], read after the host call's sync
    int* wrapList = nullptr;            // clusters the fast pair kernel left to the generic one
    int* filledUser = nullptr;          // cell fill in arrival order (input of the rank pass)
    unsigned int* exclMaxR2 = nullptr;  // [Npad] float bits: per atom, largest r2 to an excluded partner (this evaluation)
    unsigned long long* pairCounters = nullptr;   // [4]: pairs in cutoff, candidates, emitted, overflow
    int2* pairBuffer = nullptr; int64_t pairCapacity = 0;
    // pinned host staging
    double* hPos = nullptr; double* hForce = nullptr; double* hEnergy = nullptr;
    // caller buffers of the host entry point that turned out to be stable across calls are page-locked in place
    // (cudaHostRegister): positions are then DMA-read and forces accumulated by the GPU directly, no staging copies
    struct HostReg { const void* ptr = nullptr; size_t bytes = 0; bool registered = false; int seen = 0; void* dev = nullptr; };
    HostReg posReg, forceReg;
    const void* graphPos[4] = {nullptr, nullptr, nullptr, nullptr};     // what each cached graph was captured with
    const void* graphForce[4] = {nullptr, nullptr, nullptr, nullptr};
    // graphs, one per (includeForces, includeEnergy)
    cudaGraphExec_t graphs[4] = {nullptr, nullptr, nullptr, nullptr};
    int64_t launchesPerGraph[4] = {0, 0, 0, 0};
    // cached graph of the device-pointer entry (cfx_execute_device), keyed by its arguments
    struct DeviceGraphKey { const void* pos; void* force; void* dedq; void* energy; int flags; double L[3]; };
    cudaGraphExec_t devGraph = nullptr;
    DeviceGraphKey devKey = {nullptr, nullptr, nullptr, nullptr, -1, {0, 0, 0}};
    int64_t devGraphLaunches = 0;
    // cached graph of the CUDA-platform entry (cfx_execute_platform)
    struct PlatformGraphKey { const void* posq; const void* corr; const void* index; void* force; void* energy; int padded; int flags; uint64_t planGen; };
    cudaGraphExec_t platGraph = nullptr;
    PlatformGraphKey platKey = {nullptr, nullptr, nullptr, nullptr, nullptr, 0, -1, 0};
    int64_t platGraphLaunches = 0;
    bool evaluated = false;
    bool stagedPosCurrent = false;      // st.pos holds the positions of the last evaluation (host entry point only)
    uint64_t planGeneration = 0;        // bumped whenever box-dependent plans / buffers change (external graphs must re-capture)
    int64_t launches = 0;
    Box box;
    // host copies for getters
    std::vector<int> hRowDq, hRowDx, hExclPtr, hExclCols;
    // per-kernel timing (cfx_time_kernels)
    bool timing = false;
    std::vector<std::string> timeNames;
    std::vector<cudaEvent_t> timeEvents;
};

// kernel launchers (each enqueues on `s`; no synchronisation)
void launchFluxAssembly(State& st, const double* dPos, cudaStream_t s);                 // piece (1) + (4)
void launchChainRule(State& st, long long* dForce, const long long* dDedq, cudaStream_t s);   // piece (5)
void launchExclusionCorrection(State& st, const double* dPos, bool forces, long long* dForce, long long* dDedq, cudaStream_t s);
void launchNoCutoff(State& st, const double* dPos, bool forces, bool energy, long long* dForce, long long* dDedq, cudaStream_t s);
void launchFinalize(State& st, const long long* dForce, const long long* dEnergyFixed, cudaStream_t s);
void launchFinalizeToHost(State& st, const long long* dForce, const long long* dEnergyFixed, double* hostForce, bool accumulate,
                          cudaStream_t s);
void commAllReduce(State& st, long long* buf, size_t count, cudaStream_t s);           // comm.cu: in-place int64 sum over the ranks
void commDestroy(State& st);
void planKSpace(State& st);
void launchKSpace(State& st, const double* dPos, bool forces, bool energy, long long* dForce, long long* dDedq, cudaStream_t s);  // piece (3)
bool structureTensorEligible(const State& st);                                          // kspace_tc.cu
void planStructureTensor(State& st);
void launchStructureTensor(State& st, bool energy, cudaStream_t s);
double measureTf32Peak(int device, int iters);
double measureI8Peak(int device, int iters);
void planKSpaceTensor(State& st);                                                       // kspace_tc.cu
void launchGatherTensor(State& st, long long* dForce, long long* dDedq, cudaStream_t s);
void planCells(State& st);
void allocPairLists(State& st);
int cellsPerAxis(const State& st, int d);                                                // direct.cu: cell grid for the current box
void invalidatePairLists(State& st);                                                     // direct.cu: the next evaluation rebuilds                                                          // direct.cu: (re)allocate the candidate lists for st.listCap
void launchDirect(State& st, const double* dPos, bool forces, int energyMode /*0 none, 1 FP32 terms, 2 FP64 terms*/, bool emitPairs, long long* dForce, long long* dDedq, cudaStream_t s,
                  int phase = 0 /*0 all, 1 position-only part (search), 2 the rest*/); // piece (2)
void mark(State& st, const char* name, cudaStream_t s);   // per-kernel timing marker (no-op unless st.timing)
// the kernel sequence of one evaluation (api.cu); skipDiscardedEnergy: do not produce the partial energy the
// reference returns (and OpenMM discards) when includeEnergy is false -- used by the MD harness
void enqueueEvaluation(State& st, const double* dPos, bool includeForces, bool includeEnergy, long long* dForce, cudaStream_t s,
                       bool skipDiscardedEnergy);
void ensureBox(State& st, const double* box);            // validates the box, re-plans the cell grid if it changed
void setLastError(const std::string& msg);

} // namespace cfx

struct cfx_handle { cfx::State st; };

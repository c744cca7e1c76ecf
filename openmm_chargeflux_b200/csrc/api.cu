// api.cu -- the C ABI of include/cfx_b200.h: parameter preprocessing (what
// ReferenceCalcCoulForceKernel::initialize does, ReferenceCoulKernels.cpp:230-422), the per-evaluation
// kernel sequence (execute, :424-636) replayed as one CUDA graph, parity getters and timing helpers.
#include <nvtx3/nvToolsExt.h>
#include "cfx_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <set>
#include <stdexcept>
#include <utility>

namespace cfx {

static thread_local std::string g_lastError;

void throwCuda(cudaError_t code, const char* what, const char* file, int line) {
    char buf[512];
    snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", (int) code, cudaGetErrorString(code), file, line, what);
    throw std::runtime_error(buf);
}

void setLastError(const std::string& msg) { g_lastError = msg; }

// NVTX (header-only nvtx3: no-ops unless a profiler is attached): one range per entry-point call, one mark per phase
// of the kernel sequence as it is enqueued or captured (inside a replayed graph the kernels carry their own names).
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

void mark(State& st, const char* name, cudaStream_t s) {
    nvtxMarkA(name);
    if (!st.timing) return;
    cudaEvent_t ev;
    CFX_CUDA(cudaEventCreate(&ev));
    CFX_CUDA(cudaEventRecord(ev, s));
    st.timeNames.push_back(name);
    st.timeEvents.push_back(ev);
}

namespace {

struct ArgError : public std::runtime_error { using std::runtime_error::runtime_error; };
struct StateError : public std::runtime_error { using std::runtime_error::runtime_error; };

template <class T> T* upload(const std::vector<T>& v, size_t minCount = 1) {
    T* d = nullptr;
    size_t n = std::max(v.size(), minCount);
    CFX_CUDA(cudaMalloc(&d, n*sizeof(T)));
    CFX_CUDA(cudaMemset(d, 0, n*sizeof(T)));
    if (!v.empty()) CFX_CUDA(cudaMemcpy(d, v.data(), v.size()*sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

/* ReferenceCoulKernels.cpp:32-35 */
double ewaldErrorEstimate(int kmax, double width, double alpha) {
    double t = kmax*M_PI/(width*alpha);
    return 0.05*sqrt(width*alpha)*kmax*exp(-t*t);
}

// Digit planes of the integer structure-factor kernel in the forces-only call: three (23-bit fixed point, scaled by the
// largest |q|) are FP32-grade as long as the charges are of similar size; when the largest base charge dwarfs the typical
// one, the small charges would lose bits, so such systems use the four-digit variant of the energy call throughout.
int forceDigitsFor(const std::vector<double>& q0) {
    double sum2 = 0.0, big = 0.0;
    for (double q : q0) { sum2 += q*q; big = std::max(big, fabs(q)); }
    const double rms = q0.empty() ? 0.0 : sqrt(sum2/q0.size());
    return (big > 4.0*rms) ? 4 : 3;
}

void checkBox(const double* box) {
    if (box[1] != 0 || box[2] != 0 || box[3] != 0 || box[5] != 0 || box[6] != 0 || box[7] != 0)
        throw ArgError("only rectangular periodic boxes are supported (the reference's reciprocal sum uses the box diagonal only)");
    if (!(box[0] > 0 && box[4] > 0 && box[8] > 0))
        throw ArgError("periodic box lengths must be positive");
}

void setBox(State& st, const double* box) {
    st.planGeneration++;
    for (int d = 0; d < 3; d++) { st.box.L[d] = box[4*d]; st.box.invL[d] = 1.0/box[4*d]; }
}

void freeCells(State& st) {
    st.planGeneration++;
    cudaFree(st.cellOfAtom); cudaFree(st.cellCount); cudaFree(st.cellStart); cudaFree(st.cellFill);
    cudaFree(st.userLocal); cudaFree(st.sortedLocal); cudaFree(st.sortedMeta);
    cudaFree(st.pairCounters); cudaFree(st.filledUser); st.filledUser = nullptr;
    cudaFree(st.wrapList); st.wrapList = nullptr;
    cudaFree(st.userLocalD); cudaFree(st.sortedLocalD); cudaFree(st.sortedLjD);
    cudaFree(st.pairList); cudaFree(st.listCount); st.pairList = nullptr; st.listCount = nullptr; st.pairListEntries = 0;
    cudaFree(st.rebuildFlag); cudaFree(st.posAtBuild); st.rebuildFlag = nullptr; st.posAtBuild = nullptr;
    st.userLocalD = st.sortedLocalD = nullptr; st.sortedLjD = nullptr;
    st.cellOfAtom = st.cellCount = st.cellStart = st.cellFill = nullptr;
    st.userLocal = st.sortedLocal = st.sortedMeta = nullptr; st.pairCounters = nullptr;
}

void dropGraphs(State& st) {
    st.planGeneration++;
    for (int k = 0; k < 4; k++)
        if (st.graphs[k]) { cudaGraphExecDestroy(st.graphs[k]); st.graphs[k] = nullptr; }
    if (st.devGraph) { cudaGraphExecDestroy(st.devGraph); st.devGraph = nullptr; }
    if (st.platGraph) { cudaGraphExecDestroy(st.platGraph); st.platGraph = nullptr; }
}

__global__ void addFixedKernel(int n, const long long* __restrict__ src, long long* __restrict__ dst) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(reinterpret_cast<unsigned long long*>(dst + i), static_cast<unsigned long long>(src[i]));
}

__global__ void addEnergyKernel(const long long* __restrict__ energyFixed, double* __restrict__ energy) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double tot = 0.0;
        for (int k = 0; k < 4; k++) {
            const double v = (double) energyFixed[k]*(1.0/CFX_ENERGY_SCALE);
            energy[k] += v;
            tot += v;
        }
        energy[CFX_E_TOTAL] += tot;
    }
}

// host entry point with page-locked caller buffers: the caller's force array (mapped) += this evaluation's forces
__global__ void addToMappedKernel(int n, const double* __restrict__ src, double* dst) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) dst[i] += src[i];
}

// host entry point: positions read by the SMs from page-locked host memory (a copy node of this size costs ~50 us of
// DMA set-up and transfer on the critical path of the step; the load-through kernel ~15 us)
__global__ void fetchMappedKernel(int n, const double* __restrict__ src, double* __restrict__ dst) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}

// Only for handles created with CFX_OPT_PIN_CALLER_BUFFERS (the caller promises that a buffer it passes stays allocated
// until it passes a different one or destroys the handle): page-lock a caller buffer once it has been passed twice in a
// row (buffers that change every call are staged: registering costs more than the copy). Returns true when `ptr` is
// usable for DMA / mapped access.
bool useRegistered(const State& st, State::HostReg& r, const void* ptr, size_t bytes) {
    if (!st.pinCallerBuffers) return false;                     // opt-in: cfx_options.flags & CFX_OPT_PIN_CALLER_BUFFERS
    if (r.registered && r.ptr == ptr && r.bytes == bytes) return true;
    if (r.registered) { cudaHostUnregister(const_cast<void*>(r.ptr)); cudaGetLastError(); r.registered = false; r.seen = 0; r.ptr = nullptr; }
    if (r.ptr == ptr && r.bytes == bytes) r.seen++;
    else { r.ptr = ptr; r.bytes = bytes; r.seen = 1; }
    if (r.seen == 2) {
        if (cudaHostRegister(const_cast<void*>(ptr), bytes, cudaHostRegisterMapped) == cudaSuccess &&
            cudaHostGetDevicePointer(&r.dev, const_cast<void*>(ptr), 0) == cudaSuccess) { r.registered = true; return true; }
        cudaGetLastError();                                      // e.g. already registered by the caller: keep staging
    }
    return false;
}

// shard mode: the four components and their sum, still fixed point, appended to the reduction buffer
__global__ void appendEnergyFixedKernel(const long long* __restrict__ energyFixed, long long* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        long long tot = 0;
        for (int k = 0; k < 4; k++) { out[k] = energyFixed[k]; tot += energyFixed[k]; }
        out[CFX_E_TOTAL] = tot;
    }
}

} // namespace

// The kernel sequence of one evaluation. Forces are ADDED into dForce (fixed point); dE/dq of this
// evaluation is built in st.dedqFixed (zeroed here) because the chain rule must see only this
// evaluation's values.
void enqueueEvaluation(State& st, const double* dPos, bool includeForces, bool includeEnergy, long long* dForce, cudaStream_t s, bool skipDiscardedEnergy) {
    CFX_CUDA(cudaMemsetAsync(st.dedqFixed, 0, sizeof(long long)*st.Npad, s));
    CFX_CUDA(cudaMemsetAsync(st.energyFixed, 0, sizeof(long long)*8, s));
    const bool overlap = st.pbc && st.overlapBranches && !st.timing;
    const int emodeAll = includeEnergy ? 2 : ((skipDiscardedEnergy || st.skipDiscardedEnergy) ? 0 : 1);
    if (overlap) {
        // the search half of the direct-space branch needs only the positions: it starts on the side stream while the
        // charge-flux assembly runs
        CFX_CUDA(cudaMemsetAsync(st.pairCounters, 0, sizeof(unsigned long long)*4, s));
        CFX_CUDA(cudaEventRecord(st.evStart, s));
        CFX_CUDA(cudaStreamWaitEvent(st.sideStream, st.evStart, 0));
        launchDirect(st, dPos, includeForces, emodeAll, false, dForce, st.dedqFixed, st.sideStream, 1);
    }
    launchFluxAssembly(st, dPos, s);
    if (st.pbc) {
        if (!overlap) CFX_CUDA(cudaMemsetAsync(st.pairCounters, 0, sizeof(unsigned long long)*4, s));
        // reference quirks mirrored: reciprocal energy only with includeEnergy; direct,
        // self and exclusion energies always; pair/recip forces and dE/dq only with includeForces
        const int emode = includeEnergy ? 2 : ((skipDiscardedEnergy || st.skipDiscardedEnergy) ? 0 : 1);
        if (st.overlapBranches && !st.timing) {
            // The reciprocal-space and direct-space branches only meet in the fixed-point accumulators
            // (atomics), so the direct branch is forked onto a side stream: its small kernels and its tail
            // overlap with the k-space kernels. Fork/join are captured into the step's CUDA graph.
            CFX_CUDA(cudaEventRecord(st.evFork, s));
            CFX_CUDA(cudaStreamWaitEvent(st.sideStream, st.evFork, 0));
            launchExclusionCorrection(st, dPos, includeForces, dForce, st.dedqFixed, st.sideStream);   // first: see flux.cu
            launchDirect(st, dPos, includeForces, emode, false, dForce, st.dedqFixed, st.sideStream, 2);
            CFX_CUDA(cudaEventRecord(st.evJoin, st.sideStream));
            launchKSpace(st, dPos, includeForces, includeEnergy, dForce, st.dedqFixed, s);
            CFX_CUDA(cudaStreamWaitEvent(s, st.evJoin, 0));
        }
        else {
            launchKSpace(st, dPos, includeForces, includeEnergy, dForce, st.dedqFixed, s);
            launchExclusionCorrection(st, dPos, includeForces, dForce, st.dedqFixed, s);
            launchDirect(st, dPos, includeForces, emode, false, dForce, st.dedqFixed, s);
        }
    }
    else
        launchNoCutoff(st, dPos, includeForces, includeEnergy, dForce, st.dedqFixed, s);
    launchChainRule(st, dForce, st.dedqFixed, s);          // always, with whatever dE/dq was accumulated
}

void ensureCells(State& st) {
    // the cell grid depends on the current box
    int nc[3];
    for (int d = 0; d < 3; d++)
        nc[d] = cellsPerAxis(st, d);
    if (st.cellCount && nc[0] == st.cells.nc[0] && nc[1] == st.cells.nc[1] && nc[2] == st.cells.nc[2]) {
        for (int d = 0; d < 3; d++) { st.cells.csd[d] = st.box.L[d]/nc[d]; st.cells.cs[d] = (float) st.cells.csd[d]; }
        return;
    }
    freeCells(st);
    planCells(st);
}

/* ReferenceCoulKernels.cpp:403-420: smallest kmax per axis whose error estimate is below the tolerance, forced odd */
void deriveKmax(const State& st, const double* box, int K[3]) {
    for (int a = 0; a < 3; a++) {
        int k = 1;
        while (ewaldErrorEstimate(k, box[4*a], st.alpha) > st.tol) k++;
        if (k%2 == 0) k++;
        K[a] = k;
    }
}

void setKmax(State& st, const int K[3]) {
    for (int a = 0; a < 3; a++) st.ks.K[a] = K[a];
    const long long kx = K[0], ky = K[1], kz = K[2];
    st.numKVectors = (kz - 1) + (ky - 1)*(2*kz - 1) + (kx - 1)*(2*ky - 1)*(2*kz - 1);
}

void freeKSpace(State& st) {
    st.planGeneration++;
    void* ptrs[] = {st.rowS, st.colX, st.colY, st.colZ4, st.sPart, st.gCoef, st.gRowInfo, st.ks_signedStart, st.zSplit, st.coefT,
                    st.gRowData, st.gGroupInfo, st.gtTrace};
    for (void* p : ptrs) if (p) cudaFree(p);
    st.rowS = nullptr; st.colX = st.colY = nullptr; st.colZ4 = nullptr; st.sPart = nullptr; st.gCoef = nullptr; st.gRowInfo = nullptr;
    st.ks_signedStart = nullptr; st.zSplit = nullptr; st.coefT = nullptr; st.gRowData = nullptr; st.gGroupInfo = nullptr;
    st.gtTrace = nullptr;
}

void ensureBox(State& st, const double* box) {
    checkBox(box);
    if (box[0] < 2*st.cutoff || box[4] < 2*st.cutoff || box[8] < 2*st.cutoff)
        throw ArgError("the periodic box must be at least twice the cutoff in every direction");
    const bool changed = st.box.L[0] != box[0] || st.box.L[1] != box[4] || st.box.L[2] != box[8];
    if (changed) { setBox(st, box); dropGraphs(st); }
    if (changed && st.kmaxFollowsBox && st.N > 0) {
        int K[3];
        deriveKmax(st, box, K);
        if (K[0] != st.ks.K[0] || K[1] != st.ks.K[1] || K[2] != st.ks.K[2]) {
            CFX_CUDA(cudaDeviceSynchronize());                     // nothing may still read the old tables
            freeKSpace(st);
            const KSpacePlan old = st.ks;
            st.ks = KSpacePlan();
            setKmax(st, K);
            try { planKSpace(st); }
            catch (...) { freeKSpace(st); st.ks = KSpacePlan(); setKmax(st, old.K); planKSpace(st); throw; }
        }
    }
    ensureCells(st);
    if (changed) invalidatePairLists(st);
}

} // namespace cfx

using namespace cfx;


#define CFX_TRY try {
#define CFX_CATCH \
    } catch (const ArgError& e) { g_lastError = e.what(); return CFX_ERR_ARGUMENT; } \
      catch (const StateError& e) { g_lastError = e.what(); return CFX_ERR_STATE; } \
      catch (const std::exception& e) { g_lastError = e.what(); return CFX_ERR_CUDA; }

extern "C" {

const char* cfx_last_error(void) { return g_lastError.c_str(); }

int cfx_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int cfx_create(const cfx_system_desc* d, const cfx_options* opts, cfx_handle** out) {
    // a failed create must not leak the handle, its streams, pinned buffers or device allocations
    struct Guard { cfx_handle* h = nullptr; ~Guard() { if (h) cfx_destroy(h); } } guard;
    NvtxRange nvtxRange("cfx_create");
    CFX_TRY
    if (!d || !out) throw ArgError("null argument");
    *out = nullptr;
    const int N = d->num_particles;
    if (N < 0) throw ArgError("negative particle count");
    if (N > 0 && (!d->charge || !d->sigma || !d->epsilon)) throw ArgError("null particle parameter array");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        throw std::runtime_error("no CUDA device available: libcfx_b200 has no CPU fallback");
    }
    if ((d->num_exceptions > 0 && !d->exception_pairs) || (d->num_flux_bonds > 0 && (!d->flux_bond_idx || !d->flux_bond_params)) ||
        (d->num_flux_angles > 0 && (!d->flux_angle_idx || !d->flux_angle_params)) ||
        (d->num_flux_waters > 0 && (!d->flux_water_idx || !d->flux_water_params)))
        throw ArgError("null index / parameter array with a non-zero count");
    cfx_handle* h = guard.h = new cfx_handle();
    State& st = h->st;
    int dev = (opts && opts->device >= 0) ? opts->device : -1;
    if (dev < 0) CFX_CUDA(cudaGetDevice(&dev));
    if (dev >= ndev) throw ArgError("device ordinal out of range");
    CFX_CUDA(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CFX_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) throw std::runtime_error("libcfx_b200 is built for sm_100a only; device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor));
    st.device = dev;
    st.shardRank = opts ? opts->shard_rank : 0;
    st.shardCount = (opts && opts->shard_count > 0) ? opts->shard_count : 1;
    if (st.shardRank < 0 || st.shardRank >= st.shardCount) throw ArgError("shard_rank out of range");
    st.useGraph = opts ? (opts->use_graph != 0) : true;
    st.pinCallerBuffers = opts && (opts->flags & CFX_OPT_PIN_CALLER_BUFFERS);
    st.kmaxFollowsBox = opts && (opts->flags & CFX_OPT_KMAX_FOLLOWS_BOX);
    st.skipDiscardedEnergy = opts && (opts->flags & CFX_OPT_SKIP_DISCARDED_ENERGY);
    {
        const int pm = opts ? opts->list_skin_pm : 0;
        st.skin = pm < 0 ? 0.0 : (pm == 0 ? 0.1 : 1e-3*pm);
        if (const char* e = getenv("CFX_LIST_SKIN")) st.skin = std::max(0.0, atof(e));
        if (const char* e = getenv("CFX_HOST_COPY_KERNELS")) st.hostCopyKernels = atoi(e) != 0;
    }
    if (d->use_pbc == 0 && st.shardCount != 1)
        throw ArgError("sharded handles need a periodic system: the non-periodic all-pairs branch is not partitioned");
    st.N = N;
    st.Npad = std::max(256, (N + 255)/256*256);
    st.nb = d->num_flux_bonds; st.na = d->num_flux_angles; st.nw = d->num_flux_waters;
    if (st.nb < 0 || st.na < 0 || st.nw < 0 || d->num_exceptions < 0) throw ArgError("negative term count");
    st.numTerms = st.nb + st.na + st.nw;
    st.numSlots = st.nb + st.na + 3*st.nw;
    st.P = 4*st.nb + 9*st.na + 9*st.nw;
    st.pbc = d->use_pbc != 0;

    // particles: LJ pre-combination sigma/2, 2 sqrt(eps) (:238-239)
    std::vector<double> q0(d->charge, d->charge + N);
    std::vector<float2> lj(N);
    std::vector<double2> ljd(N);
    for (int i = 0; i < N; i++) {
        ljd[i] = make_double2(0.5*d->sigma[i], 2.0*sqrt(d->epsilon[i]));
        lj[i] = make_float2((float) ljd[i].x, (float) ljd[i].y);
    }

    // flux terms, Jacobian COO tables in the reference's row order (:286-383), gather CSR for the charges
    std::vector<int> termIdx(3*(size_t) st.numTerms, -1), rowDq, rowDx;
    std::vector<double> termPar(5*(size_t) st.numTerms, 0.0);
    std::vector<std::vector<std::pair<int,double> > > perAtom(N);
    auto chk = [&](int a) { if (a < 0 || a >= N) throw ArgError("flux term particle index out of range"); return a; };
    for (int t = 0; t < st.nb; t++) {
        int p[2] = {chk(d->flux_bond_idx[2*t]), chk(d->flux_bond_idx[2*t+1])};
        termIdx[3*(size_t) t] = p[0]; termIdx[3*(size_t) t + 1] = p[1];
        termPar[5*(size_t) t] = d->flux_bond_params[2*t]; termPar[5*(size_t) t + 1] = d->flux_bond_params[2*t+1];
        for (int a = 0; a < 2; a++) for (int b = 0; b < 2; b++) { rowDq.push_back(p[a]); rowDx.push_back(p[b]); }
        perAtom[p[0]].push_back(std::make_pair(t, 1.0));
        perAtom[p[1]].push_back(std::make_pair(t, -1.0));
    }
    for (int t = 0; t < st.na; t++) {
        const size_t g = (size_t) st.nb + t;
        int p[3] = {chk(d->flux_angle_idx[3*t]), chk(d->flux_angle_idx[3*t+1]), chk(d->flux_angle_idx[3*t+2])};
        for (int a = 0; a < 3; a++) termIdx[3*g + a] = p[a];
        termPar[5*g] = d->flux_angle_params[2*t]; termPar[5*g + 1] = d->flux_angle_params[2*t+1];
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { rowDq.push_back(p[a]); rowDx.push_back(p[b]); }
        const int slot = st.nb + t;
        perAtom[p[0]].push_back(std::make_pair(slot, 1.0));      // :113-115 order: p1, p3, p2
        perAtom[p[2]].push_back(std::make_pair(slot, 1.0));
        perAtom[p[1]].push_back(std::make_pair(slot, -2.0));
    }
    for (int t = 0; t < st.nw; t++) {
        const size_t g = (size_t) st.nb + st.na + t;
        int p[3] = {chk(d->flux_water_idx[3*t]), chk(d->flux_water_idx[3*t+1]), chk(d->flux_water_idx[3*t+2])};
        for (int a = 0; a < 3; a++) termIdx[3*g + a] = p[a];
        for (int a = 0; a < 5; a++) termPar[5*g + a] = d->flux_water_params[5*t + a];
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) { rowDq.push_back(p[a]); rowDx.push_back(p[b]); }
        const int slot = st.nb + st.na + 3*t;
        for (int a = 0; a < 3; a++) perAtom[p[a]].push_back(std::make_pair(slot + a, 1.0));
    }
    std::vector<int> csrPtr(N + 1, 0), csrSlot;
    std::vector<double> csrCoef;
    for (int i = 0; i < N; i++) {
        csrPtr[i] = (int) csrSlot.size();
        for (auto& e : perAtom[i]) { csrSlot.push_back(e.first); csrCoef.push_back(e.second); }
    }
    csrPtr[N] = (int) csrSlot.size();

    // exclusions: symmetric sets (:385-391) -> unique i<j pairs + sorted CSR
    std::vector<std::set<int> > ex(N);
    for (int e = 0; e < d->num_exceptions; e++) {
        int a = d->exception_pairs[2*e], b = d->exception_pairs[2*e+1];
        if (a < 0 || a >= N || b < 0 || b >= N) throw ArgError("exception particle index out of range");
        ex[a].insert(b); ex[b].insert(a);
    }
    std::vector<int2> exPairs;
    st.hExclPtr.assign(N + 1, 0);
    for (int i = 0; i < N; i++) {
        st.hExclPtr[i] = (int) st.hExclCols.size();
        for (int j : ex[i]) { st.hExclCols.push_back(j); if (i < j) exPairs.push_back(make_int2(i, j)); }
    }
    st.hExclPtr[N] = (int) st.hExclCols.size();
    st.numExcl = (int) exPairs.size();
    st.hRowDq = rowDq; st.hRowDx = rowDx;

    CFX_CUDA(cudaStreamCreateWithFlags(&st.stream, cudaStreamNonBlocking));
    CFX_CUDA(cudaStreamCreateWithFlags(&st.sideStream, cudaStreamNonBlocking));
    CFX_CUDA(cudaEventCreateWithFlags(&st.evFork, cudaEventDisableTiming));
    CFX_CUDA(cudaEventCreateWithFlags(&st.evJoin, cudaEventDisableTiming));
    CFX_CUDA(cudaEventCreateWithFlags(&st.evStart, cudaEventDisableTiming));
    st.q0 = upload(q0); st.lj = upload(lj); st.ljd = upload(ljd);
    st.siForceDigits = forceDigitsFor(q0);
    st.termIdx = upload(termIdx); st.termPar = upload(termPar);
    st.qcsrPtr = upload(csrPtr, 2); st.qcsrSlot = upload(csrSlot); st.qcsrCoef = upload(csrCoef);
    st.rowDq = upload(rowDq); st.rowDx = upload(rowDx);
    st.exclPairs = upload(exPairs); st.exclPtr = upload(st.hExclPtr, 2); st.exclCols = upload(st.hExclCols);
    CFX_CUDA(cudaMalloc(&st.pos, sizeof(double)*3*std::max(N, 1)));
    CFX_CUDA(cudaMalloc(&st.dqSlot, sizeof(double)*std::max(st.numSlots, 1)));
    CFX_CUDA(cudaMalloc(&st.rowVal, sizeof(double)*3*std::max(st.P, 1)));
    CFX_CUDA(cudaMalloc(&st.q, sizeof(double)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.qf, sizeof(float)*st.Npad));
    CFX_CUDA(cudaMemset(st.qf, 0, sizeof(float)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.forceFixed, sizeof(long long)*3*st.Npad));
    CFX_CUDA(cudaMalloc(&st.dedqFixed, sizeof(long long)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.exclMaxR2, sizeof(unsigned int)*st.Npad));
    CFX_CUDA(cudaMemset(st.exclMaxR2, 0, sizeof(unsigned int)*st.Npad));
    CFX_CUDA(cudaMalloc(&st.energyFixed, sizeof(long long)*8));
    CFX_CUDA(cudaMalloc(&st.forceOut, sizeof(double)*3*std::max(N, 1)));
    CFX_CUDA(cudaMalloc(&st.energyOut, sizeof(double)*CFX_E_COUNT));
    CFX_CUDA(cudaMallocHost(&st.hPos, sizeof(double)*3*std::max(N, 1)));
    CFX_CUDA(cudaMallocHost(&st.hForce, sizeof(double)*3*std::max(N, 1)));
    CFX_CUDA(cudaMallocHost(&st.hEnergy, sizeof(double)*CFX_E_COUNT));
    CFX_CUDA(cudaMallocHost(&st.hListOverflow, sizeof(unsigned long long)));
    *st.hListOverflow = 0;

    if (st.pbc) {
        checkBox(d->default_box);
        if (!(d->cutoff > 0)) throw ArgError("cutoff must be positive");
        if (!(d->ewald_tol > 0 && d->ewald_tol < 0.5)) throw ArgError("ewald tolerance must be in (0, 0.5)");
        st.cutoff = d->cutoff;
        st.tol = d->ewald_tol;
        st.alpha = (1.0/st.cutoff)*sqrt(-log(2.0*st.tol));            // :401
        int K[3];
        deriveKmax(st, d->default_box, K);                             // :403-420, from the DEFAULT box
        setKmax(st, K);
        setBox(st, d->default_box);
        if (N > 0) {
            planKSpace(st);
            planCells(st);
        }
    }
    guard.h = nullptr;
    *out = h;
    return CFX_OK;
    CFX_CATCH
}

void cfx_destroy(cfx_handle* h) {
    if (!h) return;
    State& st = h->st;
    cudaSetDevice(st.device);
    if (st.stream) cudaStreamSynchronize(st.stream);
    dropGraphs(st);
    commDestroy(st);
    if (st.posReg.registered) { cudaHostUnregister(const_cast<void*>(st.posReg.ptr)); cudaGetLastError(); }
    if (st.forceReg.registered) { cudaHostUnregister(const_cast<void*>(st.forceReg.ptr)); cudaGetLastError(); }
    if (st.devGraph) cudaGraphExecDestroy(st.devGraph);
    freeCells(st);
    void* ptrs[] = {st.q0, st.lj, st.ljd, st.termIdx, st.termPar, st.qcsrPtr, st.qcsrSlot, st.qcsrCoef, st.rowDq, st.rowDx, st.exclPairs,
                    st.exclPtr, st.exclCols, st.pos, st.dqSlot, st.rowVal, st.q, st.qf, st.forceFixed, st.dedqFixed, st.energyFixed,
                    st.forceOut, st.energyOut, st.rowS, st.colX, st.colY, st.colZ4, st.sPart, st.gCoef, st.gRowInfo,
                    st.ks_signedStart, st.pairBuffer, st.exclMaxR2, st.zSplit, st.coefT, st.gRowData, st.gGroupInfo, st.gtTrace};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (st.hPos) cudaFreeHost(st.hPos);
    if (st.hForce) cudaFreeHost(st.hForce);
    if (st.hEnergy) cudaFreeHost(st.hEnergy);
    if (st.hListOverflow) cudaFreeHost(st.hListOverflow);
    for (cudaEvent_t e : st.timeEvents) cudaEventDestroy(e);
    if (st.evFork) cudaEventDestroy(st.evFork);
    if (st.evJoin) cudaEventDestroy(st.evJoin);
    if (st.evStart) cudaEventDestroy(st.evStart);
    if (st.sideStream) cudaStreamDestroy(st.sideStream);
    if (st.stream) cudaStreamDestroy(st.stream);
    delete h;
}

int cfx_execute(cfx_handle* h, const double* positions, const double* box, int include_forces, int include_energy,
                double* energy, double* forces) {
    NvtxRange nvtxRange("cfx_execute");
    CFX_TRY
    if (!h) throw ArgError("null handle");
    State& st = h->st;
    const bool sharded = st.shardCount != 1;
    if (sharded && !st.comm)
        throw ArgError("cfx_execute on a sharded handle needs a communicator (cfx_comm_init); without one use cfx_execute_shard "
                       "and reduce the buffers yourself");
    if (st.N == 0) {
        if (energy) for (int k = 0; k < CFX_E_COUNT; k++) energy[k] = 0.0;
        return CFX_OK;
    }
    if (!positions) throw ArgError("null positions");
    CFX_CUDA(cudaSetDevice(st.device));
    const bool incF = include_forces != 0, incE = include_energy != 0;
    if (st.pbc) {
        if (!box) throw ArgError("null box for a periodic system");
        ensureBox(st, box);
    }
    cudaStream_t s = st.stream;
    const size_t vecBytes = sizeof(double)*3*(size_t) st.N;
    const bool regP = useRegistered(st, st.posReg, positions, vecBytes);
    const bool regF = forces && useRegistered(st, st.forceReg, forces, vecBytes);
    if (!regP) memcpy(st.hPos, positions, vecBytes);
    const int key = (incF ? 1 : 0) | (incE ? 2 : 0);
    st.launches = 0;
    const bool byKernels = st.hostCopyKernels;
    auto enqueueAll = [&]() {
        if (byKernels) {
            const double* src = regP ? static_cast<const double*>(st.posReg.dev) : st.hPos;     // page-locked either way
            fetchMappedKernel<<<(3*st.N + 255)/256, 256, 0, s>>>(3*st.N, src, st.pos);
            CFX_LAUNCH_CHECK(); st.launches++;
        }
        else
            CFX_CUDA(cudaMemcpyAsync(st.pos, regP ? positions : st.hPos, vecBytes, cudaMemcpyHostToDevice, s));
        const long long* acc = st.forceFixed;
        const long long* accEnergy = st.energyFixed;
        if (!sharded) {
            CFX_CUDA(cudaMemsetAsync(st.forceFixed, 0, sizeof(long long)*3*st.Npad, s));
            enqueueEvaluation(st, st.pos, incF, incE, st.forceFixed, s, false);
        }
        else {
            // this rank's shard into the reduction buffer (forces 2^32, energies 2^24 appended), one sum all-reduce over
            // the communicator, then every rank converts the whole result
            const size_t count = 3*(size_t) st.Npad + 8;
            CFX_CUDA(cudaMemsetAsync(st.reduceBuf, 0, sizeof(long long)*count, s));
            enqueueEvaluation(st, st.pos, incF, incE, st.reduceBuf, s, false);
            appendEnergyFixedKernel<<<1, 32, 0, s>>>(st.energyFixed, st.reduceBuf + 3*(size_t) st.Npad);
            CFX_LAUNCH_CHECK(); st.launches++;
            commAllReduce(st, st.reduceBuf, count, s);
            acc = st.reduceBuf; accEnergy = st.reduceBuf + 3*(size_t) st.Npad;
        }
        if (byKernels) {
            // one kernel converts the fixed-point sums and stores them in page-locked host memory: added into the caller's
            // registered array, or written to the staging array the host adds from after the sync
            launchFinalizeToHost(st, acc, accEnergy, regF ? static_cast<double*>(st.forceReg.dev) : (forces ? st.hForce : nullptr), regF, s);
            return;
        }
        launchFinalize(st, acc, accEnergy, s);
        if (regF) {
            addToMappedKernel<<<(3*st.N + 255)/256, 256, 0, s>>>(3*st.N, st.forceOut, static_cast<double*>(st.forceReg.dev));
            CFX_LAUNCH_CHECK(); st.launches++;
        }
        else
            CFX_CUDA(cudaMemcpyAsync(st.hForce, st.forceOut, vecBytes, cudaMemcpyDeviceToHost, s));
        CFX_CUDA(cudaMemcpyAsync(st.hEnergy, st.energyOut, sizeof(double)*CFX_E_COUNT, cudaMemcpyDeviceToHost, s));
        if (st.pbc && st.pairCounters)
            CFX_CUDA(cudaMemcpyAsync(st.hListOverflow, st.pairCounters + 11, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    };
    if (st.useGraph) {
        const void* gp = regP ? positions : nullptr;
        const void* gf = regF ? forces : (forces || !byKernels ? static_cast<const void*>(st.hForce) : nullptr);   // staged: results go to hForce
        if (st.graphs[key] && (st.graphPos[key] != gp || st.graphForce[key] != gf)) {
            cudaGraphExecDestroy(st.graphs[key]); st.graphs[key] = nullptr;
        }
        if (!st.graphs[key]) {
            cudaGraph_t graph;
            CFX_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            try { enqueueAll(); }
            catch (...) { cudaGraph_t dead; cudaStreamEndCapture(s, &dead); throw; }
            CFX_CUDA(cudaStreamEndCapture(s, &graph));
            CFX_CUDA(cudaGraphInstantiate(&st.graphs[key], graph, 0));
            CFX_CUDA(cudaGraphDestroy(graph));
            st.launchesPerGraph[key] = st.launches;
            st.graphPos[key] = gp; st.graphForce[key] = gf;
        }
        st.launches = st.launchesPerGraph[key];
        CFX_CUDA(cudaGraphLaunch(st.graphs[key], s));
    }
    else
        enqueueAll();
    CFX_CUDA(cudaStreamSynchronize(s));
    st.evaluated = true;
    st.stagedPosCurrent = true;
    if (st.pbc && *st.hListOverflow > st.listOverflowSeen) {
        // some candidate lists outgrew listCap (those clusters went through the generic pair kernel: the result is exact,
        // only slower): enlarge the lists for the next evaluation
        st.listOverflowSeen = *st.hListOverflow;
        st.listCap *= 2;
        allocPairLists(st);
        invalidatePairLists(st);
        dropGraphs(st);
    }
    if (energy) memcpy(energy, st.hEnergy, sizeof(double)*CFX_E_COUNT);
    if (forces && !regF)
        for (size_t k = 0; k < 3*(size_t) st.N; k++) forces[k] += st.hForce[k];
    return CFX_OK;
    CFX_CATCH
}

// shared body of cfx_execute_device and cfx_execute_shard (shard: d_force_fixed is the [3*Npad + 8] reduction buffer,
// zeroed here, energies appended as 2^24 fixed point)
static int executeDeviceImpl(cfx_handle* h, const double* d_positions, const double* box, int include_forces, int include_energy,
                             long long* d_force_fixed, long long* d_dedq_fixed, double* d_energy, void* stream, bool shard,
                             bool reduce = false) {
    NvtxRange nvtxRange("cfx_execute_device");
    CFX_TRY
    if (!h) throw ArgError("null handle");
    State& st = h->st;
    if (st.N == 0) return CFX_OK;
    if (!d_positions || !d_force_fixed) throw ArgError("null device buffer");
    CFX_CUDA(cudaSetDevice(st.device));
    if (st.pbc) {
        if (!box) throw ArgError("null box for a periodic system");
        ensureBox(st, box);
    }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    auto enqueueAll = [&]() {
        if (shard) CFX_CUDA(cudaMemsetAsync(d_force_fixed, 0, sizeof(long long)*(3*(size_t) st.Npad + 8), s));
        enqueueEvaluation(st, d_positions, include_forces != 0, include_energy != 0, d_force_fixed, s, false);
        if (d_dedq_fixed) {
            addFixedKernel<<<(st.N + 255)/256, 256, 0, s>>>(st.N, st.dedqFixed, d_dedq_fixed);
            CFX_LAUNCH_CHECK(); st.launches++;
        }
        if (d_energy) {
            addEnergyKernel<<<1, 32, 0, s>>>(st.energyFixed, d_energy);
            CFX_LAUNCH_CHECK(); st.launches++;
        }
        if (shard) {
            appendEnergyFixedKernel<<<1, 32, 0, s>>>(st.energyFixed, d_force_fixed + 3*(size_t) st.Npad);
            CFX_LAUNCH_CHECK(); st.launches++;
        }
        if (reduce) commAllReduce(st, d_force_fixed, 3*(size_t) st.Npad + 8, s);
    };
    st.launches = 0;
    // the legacy default stream cannot be captured: plain launches there
    const bool capturable = s != nullptr && s != cudaStreamLegacy;
    if (st.useGraph && capturable) {
        State::DeviceGraphKey key{d_positions, d_force_fixed, d_dedq_fixed, d_energy,
                                  (include_forces ? 1 : 0) | (include_energy ? 2 : 0) | (shard ? 4 : 0) | (reduce ? 8 : 0),
                                  {st.box.L[0], st.box.L[1], st.box.L[2]}};
        const State::DeviceGraphKey& old = st.devKey;
        const bool same = old.pos == key.pos && old.force == key.force && old.dedq == key.dedq && old.energy == key.energy &&
                          old.flags == key.flags && old.L[0] == key.L[0] && old.L[1] == key.L[1] && old.L[2] == key.L[2];
        if (!st.devGraph || !same) {
            if (st.devGraph) { cudaGraphExecDestroy(st.devGraph); st.devGraph = nullptr; }
            cudaGraph_t graph;
            CFX_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
            try { enqueueAll(); }
            catch (...) { cudaGraph_t dead; cudaStreamEndCapture(s, &dead); throw; }
            CFX_CUDA(cudaStreamEndCapture(s, &graph));
            CFX_CUDA(cudaGraphInstantiate(&st.devGraph, graph, 0));
            CFX_CUDA(cudaGraphDestroy(graph));
            st.devKey = key;
            st.devGraphLaunches = st.launches;
        }
        st.launches = st.devGraphLaunches;
        CFX_CUDA(cudaGraphLaunch(st.devGraph, s));
    }
    else
        enqueueAll();
    st.evaluated = true;
    st.stagedPosCurrent = false;
    return CFX_OK;
    CFX_CATCH
}

int cfx_execute_device(cfx_handle* h, const double* d_positions, const double* box, int include_forces, int include_energy,
                       long long* d_force_fixed, long long* d_dedq_fixed, double* d_energy, void* stream) {
    return executeDeviceImpl(h, d_positions, box, include_forces, include_energy, d_force_fixed, d_dedq_fixed, d_energy, stream, false);
}

int cfx_execute_shard(cfx_handle* h, const double* d_positions, const double* box, int include_forces, int include_energy,
                      long long* d_reduce, void* stream) {
    return executeDeviceImpl(h, d_positions, box, include_forces, include_energy, d_reduce, nullptr, nullptr, stream, true);
}

int cfx_execute_sharded(cfx_handle* h, const double* d_positions, const double* box, int include_forces, int include_energy,
                        long long* d_reduce, void* stream) {
    if (h && !h->st.comm) { g_lastError = "cfx_execute_sharded needs a communicator (cfx_comm_init)"; return CFX_ERR_STATE; }
    return executeDeviceImpl(h, d_positions, box, include_forces, include_energy, d_reduce, nullptr, nullptr, stream, true, true);
}

int cfx_padded_num_particles(const cfx_handle* h) { return h ? h->st.Npad : 0; }

int cfx_get_ewald_params(const cfx_handle* h, cfx_ewald_params* out) {
    if (!h || !out) { g_lastError = "null argument"; return CFX_ERR_ARGUMENT; }
    out->alpha = h->st.alpha;
    for (int a = 0; a < 3; a++) out->kmax[a] = h->st.ks.K[a];
    out->num_kvectors = h->st.numKVectors;
    return CFX_OK;
}

int cfx_get_stats(const cfx_handle* hc, cfx_stats* out) {
    CFX_TRY
    if (!hc || !out) throw ArgError("null argument");
    State& st = const_cast<cfx_handle*>(hc)->st;
    memset(out, 0, sizeof(*out));
    out->kernel_launches = st.launches;
    for (int d = 0; d < 3; d++) out->cells[d] = st.cells.nc[d];
    if (st.pbc && st.evaluated && st.pairCounters) {
        unsigned long long c[14];
        CFX_CUDA(cudaSetDevice(st.device));
        CFX_CUDA(cudaDeviceSynchronize());
        CFX_CUDA(cudaMemcpy(c, st.pairCounters, sizeof(c), cudaMemcpyDeviceToHost));
        out->pairs_in_cutoff = (int64_t) (c[0]/2);          // the kernel counts every pair from both sides
        out->pair_candidates = (int64_t) c[1];
        out->longest_pair_list = (int32_t) c[12];
        out->pair_list_builds = (int64_t) c[13];
    }
    return CFX_OK;
    CFX_CATCH
}

static void requireEvaluated(State& st) {
    if (!st.evaluated) throw StateError("no evaluation has been executed on this handle yet");
    CFX_CUDA(cudaSetDevice(st.device));
    CFX_CUDA(cudaDeviceSynchronize());
}

int cfx_get_charges(cfx_handle* h, double* q) {
    CFX_TRY
    if (!h || !q) throw ArgError("null argument");
    requireEvaluated(h->st);
    CFX_CUDA(cudaMemcpy(q, h->st.q, sizeof(double)*h->st.N, cudaMemcpyDeviceToHost));
    return CFX_OK;
    CFX_CATCH
}

int cfx_get_dedq(cfx_handle* h, double* dedq) {
    CFX_TRY
    if (!h || !dedq) throw ArgError("null argument");
    State& st = h->st;
    requireEvaluated(st);
    std::vector<long long> fx(st.N);
    CFX_CUDA(cudaMemcpy(fx.data(), st.dedqFixed, sizeof(long long)*st.N, cudaMemcpyDeviceToHost));
    for (int i = 0; i < st.N; i++) dedq[i] = (double) fx[i]/CFX_FIXED_SCALE;
    return CFX_OK;
    CFX_CATCH
}

int cfx_num_jacobian_rows(const cfx_handle* h) { return h ? h->st.P : 0; }

int cfx_get_jacobian(cfx_handle* h, int32_t* dq_idx, int32_t* dx_idx, double* val) {
    CFX_TRY
    if (!h) throw ArgError("null argument");
    State& st = h->st;
    if (dq_idx) memcpy(dq_idx, st.hRowDq.data(), sizeof(int)*st.hRowDq.size());
    if (dx_idx) memcpy(dx_idx, st.hRowDx.data(), sizeof(int)*st.hRowDx.size());
    if (val) {
        requireEvaluated(st);
        CFX_CUDA(cudaMemcpy(val, st.rowVal, sizeof(double)*3*st.P, cudaMemcpyDeviceToHost));
    }
    return CFX_OK;
    CFX_CATCH
}

int cfx_get_neighbor_pairs(cfx_handle* h, int32_t* pairs, int64_t capacity, int64_t* count) {
    CFX_TRY
    if (!h || !count) throw ArgError("null argument");
    State& st = h->st;
    if (!st.pbc) throw StateError("neighbour pairs exist only for periodic systems");
    requireEvaluated(st);
    if (!st.stagedPosCurrent)
        throw StateError("neighbour pairs are rebuilt from the positions of the last cfx_execute; the last evaluation on this handle "
                         "went through a device-pointer / timing / MD entry point");
    unsigned long long c[4];
    CFX_CUDA(cudaMemcpy(c, st.pairCounters, sizeof(c), cudaMemcpyDeviceToHost));
    *count = (int64_t) (c[0]/2);                         // the kernel counts every pair from both sides
    if (!pairs) return CFX_OK;
    if (capacity < *count) throw ArgError("pair buffer too small");
    if (st.pairCapacity < *count) {
        if (st.pairBuffer) cudaFree(st.pairBuffer);
        st.pairBuffer = nullptr;
        st.pairCapacity = *count + 1024;
        CFX_CUDA(cudaMalloc(&st.pairBuffer, sizeof(int2)*st.pairCapacity));
    }
    // re-run the cell build + pair kernel on the positions of the last host evaluation, emitting pairs
    cudaStream_t s = st.stream;
    CFX_CUDA(cudaMemsetAsync(st.pairCounters, 0, sizeof(unsigned long long)*4, s));
    launchDirect(st, st.pos, false, 0, true, st.forceFixed, st.dedqFixed, s);
    CFX_CUDA(cudaStreamSynchronize(s));
    CFX_CUDA(cudaMemcpy(c, st.pairCounters, sizeof(c), cudaMemcpyDeviceToHost));
    if ((int64_t) c[2] != *count) throw std::runtime_error("pair emission count mismatch");
    std::vector<int2> host(*count);
    CFX_CUDA(cudaMemcpy(host.data(), st.pairBuffer, sizeof(int2)*(*count), cudaMemcpyDeviceToHost));
    std::sort(host.begin(), host.end(), [](const int2& a, const int2& b) { return a.x != b.x ? a.x < b.x : a.y < b.y; });
    for (int64_t k = 0; k < *count; k++) { pairs[2*k] = host[k].x; pairs[2*k+1] = host[k].y; }
    return CFX_OK;
    CFX_CATCH
}

int cfx_get_exclusions(cfx_handle* h, int32_t* row_ptr, int32_t* cols, int64_t capacity, int64_t* count) {
    CFX_TRY
    if (!h || !count) throw ArgError("null argument");
    State& st = h->st;
    *count = (int64_t) st.hExclCols.size();
    if (row_ptr) memcpy(row_ptr, st.hExclPtr.data(), sizeof(int)*st.hExclPtr.size());
    if (cols) {
        if (capacity < *count) throw ArgError("exclusion buffer too small");
        memcpy(cols, st.hExclCols.data(), sizeof(int)*st.hExclCols.size());
    }
    return CFX_OK;
    CFX_CATCH
}

int cfx_time_device(cfx_handle* h, const double* d_positions, const double* box, int include_forces, int include_energy,
                    int iters, float* ms_per_eval) {
    CFX_TRY
    if (!h || !d_positions || !ms_per_eval || iters < 1) throw ArgError("bad argument");
    State& st = h->st;
    CFX_CUDA(cudaSetDevice(st.device));
    if (st.pbc) {
        if (!box) throw ArgError("null box for a periodic system");
        ensureBox(st, box);
    }
    st.stagedPosCurrent = false;
    cudaStream_t s = st.stream;
    // one evaluation = one graph of kernels only (inputs already resident in HBM)
    cudaGraph_t graph; cudaGraphExec_t exec;
    st.launches = 0;
    CFX_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    try {
        CFX_CUDA(cudaMemsetAsync(st.forceFixed, 0, sizeof(long long)*3*st.Npad, s));
        enqueueEvaluation(st, d_positions, include_forces != 0, include_energy != 0, st.forceFixed, s, false);
    }
    catch (...) { cudaGraph_t dead; cudaStreamEndCapture(s, &dead); throw; }
    CFX_CUDA(cudaStreamEndCapture(s, &graph));
    CFX_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    cudaEvent_t e0, e1;
    CFX_CUDA(cudaEventCreate(&e0)); CFX_CUDA(cudaEventCreate(&e1));
    CFX_CUDA(cudaGraphLaunch(exec, s));                     // warm
    CFX_CUDA(cudaStreamSynchronize(s));
    CFX_CUDA(cudaEventRecord(e0, s));
    for (int it = 0; it < iters; it++) CFX_CUDA(cudaGraphLaunch(exec, s));
    CFX_CUDA(cudaEventRecord(e1, s));
    CFX_CUDA(cudaStreamSynchronize(s));
    float ms = 0.f;
    CFX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    *ms_per_eval = ms/iters;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
    st.evaluated = true;
    return CFX_OK;
    CFX_CATCH
}

int cfx_time_kernels(cfx_handle* h, const double* d_positions, const double* box, int include_forces, int include_energy, int iters,
                     char* names, int names_capacity, float* ms, int ms_capacity, int* count) {
    CFX_TRY
    if (!h || !d_positions || !names || !ms || !count || iters < 1) throw ArgError("bad argument");
    State& st = h->st;
    CFX_CUDA(cudaSetDevice(st.device));
    if (st.pbc) {
        if (!box) throw ArgError("null box for a periodic system");
        ensureBox(st, box);
    }
    st.stagedPosCurrent = false;
    cudaStream_t s = st.stream;
    std::vector<std::string> labels;
    std::vector<double> acc;
    for (int it = 0; it < iters + 1; it++) {               // first pass is warm-up
        for (cudaEvent_t e : st.timeEvents) cudaEventDestroy(e);
        st.timeEvents.clear(); st.timeNames.clear();
        st.timing = true;
        CFX_CUDA(cudaMemsetAsync(st.forceFixed, 0, sizeof(long long)*3*st.Npad, s));
        mark(st, "begin", s);
        enqueueEvaluation(st, d_positions, include_forces != 0, include_energy != 0, st.forceFixed, s, false);
        st.timing = false;
        CFX_CUDA(cudaStreamSynchronize(s));
        if (it == 0) { labels.assign(st.timeNames.begin() + 1, st.timeNames.end()); acc.assign(labels.size(), 0.0); continue; }
        for (size_t k = 1; k < st.timeEvents.size(); k++) {
            float t = 0.f;
            CFX_CUDA(cudaEventElapsedTime(&t, st.timeEvents[k-1], st.timeEvents[k]));
            acc[k-1] += t;
        }
    }
    std::string joined;
    for (size_t k = 0; k < labels.size(); k++) { if (k) joined += ";"; joined += labels[k]; }
    if ((int) joined.size() + 1 > names_capacity || (int) labels.size() > ms_capacity) throw ArgError("output buffers too small");
    memcpy(names, joined.c_str(), joined.size() + 1);
    for (size_t k = 0; k < labels.size(); k++) ms[k] = (float) (acc[k]/iters);
    *count = (int) labels.size();
    st.evaluated = true;
    return CFX_OK;
    CFX_CATCH
}

// ------------------------------------------------------------------------------------------------
// FP32 FMA peak: 8 independent accumulator chains per thread, 2048 resident threads per SM
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fmaPeakKernel(int iters, float a, float b, float* out, long long* cycles) {
    float x0 = threadIdx.x*1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        #pragma unroll
        for (int u = 0; u < 16; u++) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    const long long t1 = clock64();
    const float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456f) out[0] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) cycles[0] = t1 - t0;
}

int cfx_measure_fp32_peak(int device, int iters, double* tflops, double* sm_clock_mhz_est) {
    CFX_TRY
    if (!tflops) throw ArgError("null argument");
    if (device >= 0) CFX_CUDA(cudaSetDevice(device));
    int dev; CFX_CUDA(cudaGetDevice(&dev));
    int numSM = 0;
    CFX_CUDA(cudaDeviceGetAttribute(&numSM, cudaDevAttrMultiProcessorCount, dev));
    float* out; long long* cyc;
    CFX_CUDA(cudaMalloc(&out, sizeof(float))); CFX_CUDA(cudaMalloc(&cyc, sizeof(long long)));
    const int blocks = numSM*8, inner = 4096;
    cudaEvent_t e0, e1;
    CFX_CUDA(cudaEventCreate(&e0)); CFX_CUDA(cudaEventCreate(&e1));
    double best = 0.0, bestClock = 0.0;
    for (int rep = 0; rep < std::max(iters, 1) + 1; rep++) {
        CFX_CUDA(cudaEventRecord(e0));
        fmaPeakKernel<<<blocks, 256>>>(inner, 0.999f, 0.001f, out, cyc);
        CFX_CUDA(cudaEventRecord(e1));
        CFX_CUDA(cudaEventSynchronize(e1));
        CFX_LAUNCH_CHECK();
        float ms = 0.f;
        CFX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        long long c = 0;
        CFX_CUDA(cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost));
        const double flops = 2.0*128.0*inner*256.0*blocks;
        const double tf = flops/(ms*1e-3)/1e12;
        if (rep > 0 && tf > best) { best = tf; bestClock = (double) c/(ms*1e-3)/1e6; }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out); cudaFree(cyc);
    *tflops = best;
    // block 0's cycle count over the whole kernel's wall time under-estimates the clock when the
    // grid runs in several waves; it is reported as a rough cross-check only
    if (sm_clock_mhz_est) *sm_clock_mhz_est = bestClock;
    return CFX_OK;
    CFX_CATCH
}

} // extern "C"

// SURVEY.md 8 f4 (absent in the reference): new parameter VALUES for an existing handle, same topology. Only q0, the LJ
// pre-combination and the flux-term parameters depend on them; device pointers stay, so the cached CUDA graphs stay valid.
int cfx_update_parameters(cfx_handle* h, const cfx_system_desc* d) {
    CFX_TRY
    if (!h || !d) throw ArgError("null argument");
    State& st = h->st;
    if (d->num_particles != st.N || d->num_flux_bonds != st.nb || d->num_flux_angles != st.na || d->num_flux_waters != st.nw)
        throw ArgError("cfx_update_parameters: particle / flux term counts differ from the handle's (topology changes need a new handle)");
    if (st.N == 0) return CFX_OK;
    if (!d->charge || !d->sigma || !d->epsilon) throw ArgError("null particle parameter array");
    if ((st.nb && !d->flux_bond_params) || (st.na && !d->flux_angle_params) || (st.nw && !d->flux_water_params))
        throw ArgError("null flux parameter array");
    CFX_CUDA(cudaSetDevice(st.device));
    const int N = st.N;
    std::vector<double> q0(d->charge, d->charge + N);
    std::vector<float2> lj(N);
    std::vector<double2> ljd(N);
    for (int i = 0; i < N; i++) {
        if (!(d->epsilon[i] >= 0.0)) throw ArgError("negative LJ epsilon");
        ljd[i] = make_double2(0.5*d->sigma[i], 2.0*sqrt(d->epsilon[i]));
        lj[i] = make_float2((float) ljd[i].x, (float) ljd[i].y);
    }
    std::vector<double> termPar(5*(size_t) st.numTerms, 0.0);
    for (int t = 0; t < st.nb; t++) { termPar[5*(size_t) t] = d->flux_bond_params[2*t]; termPar[5*(size_t) t + 1] = d->flux_bond_params[2*t+1]; }
    for (int t = 0; t < st.na; t++) { const size_t g = (size_t) st.nb + t; termPar[5*g] = d->flux_angle_params[2*t]; termPar[5*g + 1] = d->flux_angle_params[2*t+1]; }
    for (int t = 0; t < st.nw; t++) { const size_t g = (size_t) st.nb + st.na + t; for (int a = 0; a < 5; a++) termPar[5*g + a] = d->flux_water_params[5*t + a]; }
    if (st.stream) CFX_CUDA(cudaStreamSynchronize(st.stream));
    CFX_CUDA(cudaMemcpy(st.q0, q0.data(), sizeof(double)*N, cudaMemcpyHostToDevice));
    if (forceDigitsFor(q0) != st.siForceDigits) { st.siForceDigits = forceDigitsFor(q0); dropGraphs(st); }
    CFX_CUDA(cudaMemcpy(st.lj, lj.data(), sizeof(float2)*N, cudaMemcpyHostToDevice));
    CFX_CUDA(cudaMemcpy(st.ljd, ljd.data(), sizeof(double2)*N, cudaMemcpyHostToDevice));
    if (st.numTerms) CFX_CUDA(cudaMemcpy(st.termPar, termPar.data(), sizeof(double)*termPar.size(), cudaMemcpyHostToDevice));
    if (st.pbc) invalidatePairLists(st);                 // the sorted records carry the LJ parameters
    st.evaluated = false;
    return CFX_OK;
    CFX_CATCH
}

int cfx_measure_tf32_peak(int device, int iters, double* tflops) {
    CFX_TRY
    if (!tflops || iters < 1) throw ArgError("bad argument");
    int dev = device;
    if (dev < 0) CFX_CUDA(cudaGetDevice(&dev));
    *tflops = measureTf32Peak(dev, iters);
    return CFX_OK;
    CFX_CATCH
}

int cfx_measure_i8_peak(int device, int iters, double* tops) {
    CFX_TRY
    if (!tops || iters < 1) throw ArgError("bad argument");
    int dev = device;
    if (dev < 0) CFX_CUDA(cudaGetDevice(&dev));
    *tops = measureI8Peak(dev, iters);
    return CFX_OK;
    CFX_CATCH
}

// debug only (not part of include/cfx_b200.h): phase timestamps of the last tensor-gather launch, [148][32] ns
extern "C" int cfx_debug_gather_trace(cfx_handle* h, unsigned long long* out) {
    if (!h || !h->st.gtTrace) return CFX_ERR_STATE;
    cudaSetDevice(h->st.device);
    cudaDeviceSynchronize();
    return cudaMemcpy(out, h->st.gtTrace, 148*32*sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess ? CFX_OK : CFX_ERR_CUDA;
}

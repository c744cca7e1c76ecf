// md.cu -- MD harness around the path: harmonic bond/angle forces and a
// velocity-Verlet integrator, so the "NVE MD" benchmark configuration runs without OpenMM and energy
// conservation (which needs the charge-flux chain-rule forces to be right) can be checked at full size.
// In an OpenMM simulation these pieces are OpenMM's own HarmonicBondForce / HarmonicAngleForce /
// VerletIntegrator; they are not part of the plugin.
#include "cfx_internal.cuh"

#include <cmath>
#include <stdexcept>

using namespace cfx;

struct cfx_md {
    cfx_handle* h = nullptr;
    int N = 0, Npad = 0, nb = 0, na = 0;
    double* mass = nullptr; double* pos = nullptr; double* vel = nullptr;
    int* bondIdx = nullptr; double* bondPar = nullptr; int* angleIdx = nullptr; double* anglePar = nullptr;
    long long* force = nullptr;         // [3][Npad] fixed point
    long long* scratch = nullptr;       // [4] fixed-point energies: kinetic, bonded
    double box[9];
    cudaGraphExec_t stepGraph = nullptr;
    double graphDt = -1.0;
    uint64_t graphPlanGen = 0;          // State::planGeneration the step graph was captured with
    bool forcesCurrent = false;
};

namespace {

struct BoxD { double Lx, Ly, Lz; };

__device__ __forceinline__ double3 minImage(const double* __restrict__ pos, int I, int J, BoxD b) {
    double3 d = make_double3(pos[3*J] - pos[3*I], pos[3*J+1] - pos[3*I+1], pos[3*J+2] - pos[3*I+2]);
    d.z -= b.Lz*floor(d.z/b.Lz + 0.5);
    d.y -= b.Ly*floor(d.y/b.Ly + 0.5);
    d.x -= b.Lx*floor(d.x/b.Lx + 0.5);
    return d;
}

__device__ __forceinline__ void addForce(long long* f, int Npad, int i, double x, double y, double z) {
    atomicAddFixed(f + i, x); atomicAddFixed(f + Npad + i, y); atomicAddFixed(f + 2*(size_t) Npad + i, z);
}

// one thread per bond / angle: E = k/2 (r - r0)^2, E = k/2 (theta - theta0)^2
__global__ void __launch_bounds__(128) bondedKernel(int nb, int na, int Npad, const int* __restrict__ bondIdx, const double* __restrict__ bondPar,
        const int* __restrict__ angleIdx, const double* __restrict__ anglePar, const double* __restrict__ pos, BoxD box,
        long long* __restrict__ force, long long* __restrict__ energyFixed) {
    __shared__ double scratch[32];
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    double e = 0.0;
    if (t < nb) {
        const int i = bondIdx[2*t], j = bondIdx[2*t+1];
        const double k = bondPar[2*t], r0 = bondPar[2*t+1];
        const double3 d = minImage(pos, i, j, box);               // j - i
        const double r = sqrt(d.x*d.x + d.y*d.y + d.z*d.z);
        const double dr = r - r0;
        e = 0.5*k*dr*dr;
        const double s = k*dr/r;                                  // dE/dr / r
        addForce(force, Npad, i, s*d.x, s*d.y, s*d.z);
        addForce(force, Npad, j, -s*d.x, -s*d.y, -s*d.z);
    }
    else if (t < nb + na) {
        const int a = t - nb;
        const int i = angleIdx[3*a], j = angleIdx[3*a+1], k3 = angleIdx[3*a+2];   // j is the apex
        const double k = anglePar[2*a], th0 = anglePar[2*a+1];
        const double3 u = minImage(pos, j, i, box), v = minImage(pos, j, k3, box);
        const double ru2 = u.x*u.x + u.y*u.y + u.z*u.z, rv2 = v.x*v.x + v.y*v.y + v.z*v.z;
        const double ru = sqrt(ru2), rv = sqrt(rv2);
        double c = (u.x*v.x + u.y*v.y + u.z*v.z)/(ru*rv);
        c = fmin(1.0, fmax(-1.0, c));
        const double th = acos(c);
        const double dth = th - th0;
        e = 0.5*k*dth*dth;
        const double sinth = sqrt(fmax(1e-30, 1.0 - c*c));
        const double pre = k*dth/sinth;                           // -dE/dcos(theta)
        // dcos/du = v/(ru rv) - c u/ru^2 ; dcos/dv = u/(ru rv) - c v/rv^2 ; F = +pre * dcos/dx
        const double iuv = 1.0/(ru*rv);
        const double fix = pre*(v.x*iuv - c*u.x/ru2), fiy = pre*(v.y*iuv - c*u.y/ru2), fiz = pre*(v.z*iuv - c*u.z/ru2);
        const double fkx = pre*(u.x*iuv - c*v.x/rv2), fky = pre*(u.y*iuv - c*v.y/rv2), fkz = pre*(u.z*iuv - c*v.z/rv2);
        addForce(force, Npad, i, fix, fiy, fiz);
        addForce(force, Npad, k3, fkx, fky, fkz);
        addForce(force, Npad, j, -fix - fkx, -fiy - fky, -fiz - fkz);
    }
    if (energyFixed) {
        e = blockSum(e, scratch);
        if (threadIdx.x == 0) atomicAddEnergy(energyFixed + 1, e);
    }
}

// v += (dt/2) F/m ; optionally x += dt v
__global__ void __launch_bounds__(256) kickKernel(int N, int Npad, const long long* __restrict__ force, const double* __restrict__ mass,
        double* __restrict__ vel, double* __restrict__ pos, double halfDt, double dt, bool drift) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double s = halfDt/mass[i]*(1.0/CFX_FIXED_SCALE);
    #pragma unroll
    for (int c = 0; c < 3; c++) {
        double v = vel[3*(size_t) i + c] + s*(double) force[(size_t) c*Npad + i];
        vel[3*(size_t) i + c] = v;
        if (drift) pos[3*(size_t) i + c] += dt*v;
    }
}

__global__ void __launch_bounds__(256) kineticKernel(int N, const double* __restrict__ mass, const double* __restrict__ vel,
        long long* __restrict__ energyFixed) {
    __shared__ double scratch[32];
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    double e = 0.0;
    if (i < N) {
        const double vx = vel[3*(size_t) i], vy = vel[3*(size_t) i + 1], vz = vel[3*(size_t) i + 2];
        e = 0.5*mass[i]*(vx*vx + vy*vy + vz*vz);
    }
    e = blockSum(e, scratch);
    if (threadIdx.x == 0) atomicAddEnergy(energyFixed, e);
}

// steepest descent: x += min(step*|F|, cap) * F/|F|
__global__ void __launch_bounds__(256) descentKernel(int N, int Npad, const long long* __restrict__ force, double* __restrict__ pos,
        double step, double cap) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i >= N) return;
    const double inv = 1.0/CFX_FIXED_SCALE;
    const double fx = (double) force[i]*inv, fy = (double) force[Npad + i]*inv, fz = (double) force[2*(size_t) Npad + i]*inv;
    const double f = sqrt(fx*fx + fy*fy + fz*fz);
    if (f <= 0.0) return;
    const double d = fmin(step*f, cap)/f;
    pos[3*(size_t) i] += d*fx; pos[3*(size_t) i + 1] += d*fy; pos[3*(size_t) i + 2] += d*fz;
}

void computeForces(cfx_md* md, cudaStream_t s, bool withEnergy) {
    State& st = md->h->st;
    CFX_CUDA(cudaMemsetAsync(md->force, 0, sizeof(long long)*3*md->Npad, s));
    if (withEnergy) CFX_CUDA(cudaMemsetAsync(md->scratch, 0, sizeof(long long)*4, s));
    const int terms = md->nb + md->na;
    if (terms > 0) {
        bondedKernel<<<(terms + 127)/128, 128, 0, s>>>(md->nb, md->na, md->Npad, md->bondIdx, md->bondPar, md->angleIdx, md->anglePar,
                md->pos, BoxD{md->box[0], md->box[4], md->box[8]}, md->force, withEnergy ? md->scratch : nullptr);
        CFX_LAUNCH_CHECK();
    }
    enqueueEvaluation(st, md->pos, true, withEnergy, md->force, s, true);
    st.stagedPosCurrent = false;                 // the handle's staged positions no longer match its last evaluation
    st.evaluated = true;
}

template <class T> T* dupload(const T* src, size_t n) {
    T* d = nullptr;
    CFX_CUDA(cudaMalloc(&d, std::max<size_t>(n, 1)*sizeof(T)));
    if (n) CFX_CUDA(cudaMemcpy(d, src, n*sizeof(T), cudaMemcpyHostToDevice));
    return d;
}

} // namespace

#define MD_TRY try {
#define MD_CATCH } catch (const std::exception& e) { setLastError(e.what()); return CFX_ERR_CUDA; }

extern "C" {

int cfx_md_create(cfx_handle* h, const double* masses, int32_t nb, const int32_t* bondIdx, const double* bondPar,
                  int32_t na, const int32_t* angleIdx, const double* anglePar, cfx_md** out) {
    MD_TRY
    if (!h || !masses || !out) { setLastError("null argument"); return CFX_ERR_ARGUMENT; }
    State& st = h->st;
    if (!st.pbc) { setLastError("the MD harness needs a periodic system"); return CFX_ERR_ARGUMENT; }
    if (st.shardCount != 1) { setLastError("the MD harness integrates whole forces: it needs an unsharded handle"); return CFX_ERR_ARGUMENT; }
    if ((nb > 0 && (!bondIdx || !bondPar)) || (na > 0 && (!angleIdx || !anglePar))) { setLastError("null bonded-term array"); return CFX_ERR_ARGUMENT; }
    for (int i = 0; i < 2*nb; i++) if (bondIdx[i] < 0 || bondIdx[i] >= st.N) { setLastError("bond index out of range"); return CFX_ERR_ARGUMENT; }
    for (int i = 0; i < 3*na; i++) if (angleIdx[i] < 0 || angleIdx[i] >= st.N) { setLastError("angle index out of range"); return CFX_ERR_ARGUMENT; }
    CFX_CUDA(cudaSetDevice(st.device));
    struct Guard { cfx_md* md = nullptr; ~Guard() { if (md) cfx_md_destroy(md); } } guard;     // no leak when an allocation below throws
    cfx_md* md = guard.md = new cfx_md();
    md->h = h; md->N = st.N; md->Npad = st.Npad; md->nb = nb; md->na = na;
    md->mass = dupload(masses, st.N);
    md->bondIdx = dupload(bondIdx, 2*(size_t) nb); md->bondPar = dupload(bondPar, 2*(size_t) nb);
    md->angleIdx = dupload(angleIdx, 3*(size_t) na); md->anglePar = dupload(anglePar, 2*(size_t) na);
    CFX_CUDA(cudaMalloc(&md->pos, sizeof(double)*3*st.N));
    CFX_CUDA(cudaMalloc(&md->vel, sizeof(double)*3*st.N));
    CFX_CUDA(cudaMemset(md->vel, 0, sizeof(double)*3*st.N));
    CFX_CUDA(cudaMalloc(&md->force, sizeof(long long)*3*st.Npad));
    CFX_CUDA(cudaMalloc(&md->scratch, sizeof(long long)*4));
    for (int k = 0; k < 9; k++) md->box[k] = 0.0;
    md->box[0] = st.box.L[0]; md->box[4] = st.box.L[1]; md->box[8] = st.box.L[2];
    guard.md = nullptr;
    *out = md;
    return CFX_OK;
    MD_CATCH
}

void cfx_md_destroy(cfx_md* md) {
    if (!md) return;
    cudaSetDevice(md->h->st.device);
    cudaStreamSynchronize(md->h->st.stream);
    if (md->stepGraph) cudaGraphExecDestroy(md->stepGraph);
    void* ptrs[] = {md->mass, md->pos, md->vel, md->bondIdx, md->bondPar, md->angleIdx, md->anglePar, md->force, md->scratch};
    for (void* p : ptrs) if (p) cudaFree(p);
    delete md;
}

int cfx_md_set_state(cfx_md* md, const double* positions, const double* velocities, const double* box) {
    MD_TRY
    State& st = md->h->st;
    CFX_CUDA(cudaSetDevice(st.device));
    CFX_CUDA(cudaStreamSynchronize(st.stream));
    if (box) {
        for (int k = 0; k < 9; k++) md->box[k] = box[k];
        ensureBox(st, box);
        if (md->stepGraph) { cudaGraphExecDestroy(md->stepGraph); md->stepGraph = nullptr; }
    }
    if (positions) CFX_CUDA(cudaMemcpy(md->pos, positions, sizeof(double)*3*md->N, cudaMemcpyHostToDevice));
    if (velocities) CFX_CUDA(cudaMemcpy(md->vel, velocities, sizeof(double)*3*md->N, cudaMemcpyHostToDevice));
    md->forcesCurrent = false;
    return CFX_OK;
    MD_CATCH
}

int cfx_md_get_state(cfx_md* md, double* positions, double* velocities) {
    MD_TRY
    State& st = md->h->st;
    CFX_CUDA(cudaSetDevice(st.device));
    CFX_CUDA(cudaStreamSynchronize(st.stream));
    if (positions) CFX_CUDA(cudaMemcpy(positions, md->pos, sizeof(double)*3*md->N, cudaMemcpyDeviceToHost));
    if (velocities) CFX_CUDA(cudaMemcpy(velocities, md->vel, sizeof(double)*3*md->N, cudaMemcpyDeviceToHost));
    return CFX_OK;
    MD_CATCH
}

int cfx_md_minimize(cfx_md* md, int32_t steps, double maxDisp) {
    MD_TRY
    State& st = md->h->st;
    CFX_CUDA(cudaSetDevice(st.device));
    ensureBox(st, md->box);
    cudaStream_t s = st.stream;
    for (int it = 0; it < steps; it++) {
        computeForces(md, s, false);
        descentKernel<<<(md->N + 255)/256, 256, 0, s>>>(md->N, md->Npad, md->force, md->pos, 1e-6, maxDisp);
        CFX_LAUNCH_CHECK();
    }
    CFX_CUDA(cudaStreamSynchronize(s));
    md->forcesCurrent = false;
    return CFX_OK;
    MD_CATCH
}

int cfx_md_step(cfx_md* md, int32_t nsteps, double dt, float* msElapsed) {
    MD_TRY
    State& st = md->h->st;
    CFX_CUDA(cudaSetDevice(st.device));
    ensureBox(st, md->box);
    cudaStream_t s = st.stream;
    if (!md->forcesCurrent) { computeForces(md, s, false); md->forcesCurrent = true; }
    // the step graph captures the handle's cell-list buffers and box-dependent kernel arguments by value: any re-plan
    // of the handle (box change through another entry point) invalidates it
    if (!md->stepGraph || md->graphDt != dt || md->graphPlanGen != st.planGeneration) {
        if (md->stepGraph) { cudaGraphExecDestroy(md->stepGraph); md->stepGraph = nullptr; }
        cudaGraph_t graph;
        CFX_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        try {
            // velocity Verlet: half kick + drift with F(t), new forces, half kick with F(t+dt)
            kickKernel<<<(md->N + 255)/256, 256, 0, s>>>(md->N, md->Npad, md->force, md->mass, md->vel, md->pos, 0.5*dt, dt, true);
            CFX_LAUNCH_CHECK();
            computeForces(md, s, false);
            kickKernel<<<(md->N + 255)/256, 256, 0, s>>>(md->N, md->Npad, md->force, md->mass, md->vel, md->pos, 0.5*dt, dt, false);
            CFX_LAUNCH_CHECK();
        }
        catch (...) { cudaGraph_t dead; cudaStreamEndCapture(s, &dead); throw; }
        CFX_CUDA(cudaStreamEndCapture(s, &graph));
        CFX_CUDA(cudaGraphInstantiate(&md->stepGraph, graph, 0));
        CFX_CUDA(cudaGraphDestroy(graph));
        md->graphDt = dt;
        md->graphPlanGen = st.planGeneration;
    }
    cudaEvent_t e0, e1;
    CFX_CUDA(cudaEventCreate(&e0)); CFX_CUDA(cudaEventCreate(&e1));
    CFX_CUDA(cudaEventRecord(e0, s));
    for (int it = 0; it < nsteps; it++) CFX_CUDA(cudaGraphLaunch(md->stepGraph, s));
    CFX_CUDA(cudaEventRecord(e1, s));
    CFX_CUDA(cudaStreamSynchronize(s));
    float ms = 0.f;
    CFX_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (msElapsed) *msElapsed = ms;
    return CFX_OK;
    MD_CATCH
}

int cfx_md_energies(cfx_md* md, double* out4) {
    MD_TRY
    State& st = md->h->st;
    CFX_CUDA(cudaSetDevice(st.device));
    ensureBox(st, md->box);
    cudaStream_t s = st.stream;
    computeForces(md, s, true);
    md->forcesCurrent = true;
    kineticKernel<<<(md->N + 255)/256, 256, 0, s>>>(md->N, md->mass, md->vel, md->scratch);
    CFX_LAUNCH_CHECK();
    CFX_CUDA(cudaStreamSynchronize(s));
    long long fx[4], ef[8];
    CFX_CUDA(cudaMemcpy(fx, md->scratch, sizeof(fx), cudaMemcpyDeviceToHost));
    CFX_CUDA(cudaMemcpy(ef, st.energyFixed, sizeof(ef), cudaMemcpyDeviceToHost));
    out4[0] = (double) fx[0]/CFX_ENERGY_SCALE;
    out4[1] = (double) fx[1]/CFX_ENERGY_SCALE;
    out4[2] = ((double) ef[0] + (double) ef[1] + (double) ef[2] + (double) ef[3])/CFX_ENERGY_SCALE;
    out4[3] = out4[0] + out4[1] + out4[2];
    return CFX_OK;
    MD_CATCH
}

} // extern "C"

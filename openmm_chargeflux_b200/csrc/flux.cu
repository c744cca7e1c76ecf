// flux.cu -- the O(N) double-precision pieces of the path:
//   piece (1) charge-flux assembly q(x) + sparse Jacobian rows   (ReferenceCoulKernels.cpp:37-228)
//   piece (4) Ewald self term, fused into the per-atom charge sum (:507-510)
//   piece (5) chain rule F -= dE/dq * dq/dx                       (:493-499, :626-632)
//   excluded-pair erf correction                                  (:596-622)
//   non-periodic all-pairs branch                                 (:436-491)
// These move a few MB per evaluation and are launch-latency bound; they are kept in FP64 because the
// self and exclusion energies (~ +-3e6 kJ/mol at 32k atoms) cancel to ~1e-1 of their size.
#include "cfx_internal.cuh"

namespace cfx {

namespace {

struct BoxD { double Lx, Ly, Lz; };

// OpenMM ReferenceForce::getDeltaRPeriodic for a rectangular box: J - I, then z, y, x wrap.
__device__ __forceinline__ double3 deltaPeriodic(const double* __restrict__ pos, int I, int J, BoxD b, bool pbc) {
    double3 d = make_double3(pos[3*J] - pos[3*I], pos[3*J+1] - pos[3*I+1], pos[3*J+2] - pos[3*I+2]);
    if (pbc) {
        d.z -= b.Lz*floor(d.z/b.Lz + 0.5);
        d.y -= b.Ly*floor(d.y/b.Ly + 0.5);
        d.x -= b.Lx*floor(d.x/b.Lx + 0.5);
    }
    return d;
}
__device__ __forceinline__ double dot3(double3 a) { return a.x*a.x + a.y*a.y + a.z*a.z; }
__device__ __forceinline__ double comp(double3 a, int j) { return j == 0 ? a.x : (j == 1 ? a.y : a.z); }

// One thread per flux term: geometry -> charge increments (dqSlot) and Jacobian rows (rowVal).
// Slot layout: bond t -> t; angle t -> nb+t; water t -> nb+na+3t+{0,1,2} = (dq1,dq2,dq3).
__global__ void __launch_bounds__(128) fluxTermKernel(int nb, int na, int nw, const int* __restrict__ termIdx,
        const double* __restrict__ termPar, const double* __restrict__ pos, BoxD box, bool pbc,
        double* __restrict__ dqSlot, double* __restrict__ rowVal) {
    const int t = blockIdx.x*blockDim.x + threadIdx.x;
    if (t >= nb + na + nw) return;
    const int* idx = termIdx + 3*t;
    const double* par = termPar + 5*t;
    if (t < nb) {
        const int p1 = idx[0], p2 = idx[1];
        const double k = par[0], b = par[1];
        double3 d = deltaPeriodic(pos, p1, p2, box, pbc);
        double r = sqrt(dot3(d));
        dqSlot[t] = k*(r - b);
        double c = k/r;
        double* rows = rowVal + 12*(size_t) t;
        #pragma unroll
        for (int j = 0; j < 3; j++) {
            double v = c*comp(d, j);
            rows[j] = -v; rows[3+j] = v; rows[6+j] = v; rows[9+j] = -v;
        }
    }
    else if (t < nb + na) {
        const int a = t - nb;
        const int p1 = idx[0], p2 = idx[1], p3 = idx[2];
        const double k = par[0], theta0 = par[1];
        double3 d21 = deltaPeriodic(pos, p2, p1, box, pbc);
        double3 d23 = deltaPeriodic(pos, p2, p3, box, pbc);
        double3 d13 = deltaPeriodic(pos, p1, p3, box, pbc);
        double r21s = dot3(d21), r23s = dot3(d23), r13s = dot3(d13);
        double r21 = sqrt(r21s), r23 = sqrt(r23s);
        double cost = (r23s + r21s - r13s)/2/r21/r23;
        dqSlot[nb + a] = k*(acos(cost) - theta0);
        double invRR = 1.0/r21/r23;
        double invSin = 1/sqrt(1 - cost*cost);
        double c1 = k*invRR*invSin;
        double c21 = k*cost*invSin/r21s;
        double c23 = k*cost*invSin/r23s;
        double* rows = rowVal + 3*((size_t) 4*nb + 9*(size_t) a);
        #pragma unroll
        for (int j = 0; j < 3; j++) {
            double v1 = -c1*comp(d23, j) + c21*comp(d21, j);
            double v3 = -c1*comp(d21, j) + c23*comp(d23, j);
            double v2 = -v1 - v3;
            rows[j] = v1;        rows[3+j] = v2;        rows[6+j] = v3;
            rows[9+j] = -2*v1;   rows[12+j] = -2*v2;    rows[15+j] = -2*v3;
            rows[18+j] = v1;     rows[21+j] = v2;       rows[24+j] = v3;
        }
    }
    else {
        const int w = t - nb - na;
        const int p1 = idx[0], p2 = idx[1], p3 = idx[2];
        const double k1 = par[0], k2 = par[1], kub = par[2], b0 = par[3], ub0 = par[4];
        double3 d12 = deltaPeriodic(pos, p1, p2, box, pbc);
        double3 d13 = deltaPeriodic(pos, p1, p3, box, pbc);
        double3 d23 = deltaPeriodic(pos, p2, p3, box, pbc);
        double r12 = sqrt(dot3(d12)), r13 = sqrt(dot3(d13)), r23 = sqrt(dot3(d23));
        double dq2 = k1*(r12 - b0) + k2*(r13 - b0) + kub*(r23 - ub0);
        double dq3 = k1*(r13 - b0) + k2*(r12 - b0) + kub*(r23 - ub0);
        double* slot = dqSlot + nb + na + 3*(size_t) w;
        slot[0] = -dq2 - dq3; slot[1] = dq2; slot[2] = dq3;
        double i12 = 1.0/r12, i13 = 1.0/r13, i23 = 1.0/r23;
        double* rows = rowVal + 3*((size_t) 4*nb + 9*(size_t) na + 9*(size_t) w);
        #pragma unroll
        for (int j = 0; j < 3; j++) {
            double n12 = comp(d12, j)*i12, n13 = comp(d13, j)*i13, n23 = comp(d23, j)*i23;
            double a1 = k1*n12, a2 = k2*n12, b1 = k1*n13, b2 = k2*n13, u = kub*n23;
            rows[j]    = a1 + a2 + b1 + b2;
            rows[3+j]  = -a1 - a2 + 2*u;
            rows[6+j]  = -b2 - b1 - 2*u;
            rows[9+j]  = -a1 - b2;
            rows[12+j] = a1 - u;
            rows[15+j] = b2 + u;
            rows[18+j] = -a2 - b1;
            rows[21+j] = a2 - u;
            rows[24+j] = b1 + u;
        }
    }
}

// One thread per atom: gather-sum of the charge increments in the reference's accumulation order
// (bit-reproducible, no atomics), then the self term: dE/dq_i = -2 ke alpha/sqrt(pi) q_i and
// E_self = -ke alpha/sqrt(pi) sum q^2. Also clears nothing: accumulators are zeroed by the caller.
__global__ void __launch_bounds__(256) chargeSumSelfKernel(int N, int Npad, const double* __restrict__ q0,
        const int* __restrict__ csrPtr, const int* __restrict__ csrSlot, const double* __restrict__ csrCoef,
        const double* __restrict__ dqSlot, double* __restrict__ q, float* __restrict__ qf,
        bool selfTerm, double selfCoef /* ke*alpha/sqrt(pi) */, long long* __restrict__ dedqFixed,
        long long* __restrict__ energyFixed) {
    __shared__ double scratch[32];
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    double e = 0.0;
    if (i < N) {
        double qi = q0[i];
        for (int e0 = csrPtr[i]; e0 < csrPtr[i+1]; e0++)
            qi += csrCoef[e0]*dqSlot[csrSlot[e0]];      // coef in {+1,-1,-2}: exact product
        q[i] = qi;
        qf[i] = (float) qi;
        if (selfTerm) {
            dedqFixed[i] += toFixed(-2.0*selfCoef*qi);
            e = -selfCoef*qi*qi;
        }
    }
    else if (i < Npad)
        qf[i] = 0.0f;
    {   // largest |q| of this evaluation: fixes the fixed-point scale of the integer structure-factor kernel
        unsigned int mb = (i < N) ? __float_as_uint(fabsf(qf[i])) : 0u;
        mb = __reduce_max_sync(0xffffffffu, mb);
        if ((threadIdx.x & 31) == 0 && mb != 0u)
            atomicMax(reinterpret_cast<unsigned long long*>(energyFixed + CFX_SLOT_QMAX), (unsigned long long) mb);
    }
    if (selfTerm) {
        e = blockSum(e, scratch);
        if (threadIdx.x == 0) atomicAddEnergy(energyFixed + CFX_E_SELF, e);
    }
}

// One thread per Jacobian row (dq atom a, dx atom b): F[b] -= dE/dq[a] * row.
__global__ void __launch_bounds__(256) chainRuleKernel(int P, int Npad, const int* __restrict__ rowDq,
        const int* __restrict__ rowDx, const double* __restrict__ rowVal, const long long* __restrict__ dedqFixed,
        long long* __restrict__ forceFixed) {
    const int r = blockIdx.x*blockDim.x + threadIdx.x;
    if (r >= P) return;
    const int a = rowDq[r], b = rowDx[r];
    const double d = (double) dedqFixed[a]*(1.0/CFX_FIXED_SCALE);
    #pragma unroll
    for (int c = 0; c < 3; c++)
        atomicAddFixed(forceFixed + (size_t) c*Npad + b, -d*rowVal[3*(size_t) r + c]);
}

// One thread per unique excluded pair (i<j). Periodic: remove the reciprocal-space image of the pair
// (erf term, no cutoff test, no LJ). Non-periodic: subtract the full Coulomb + LJ pair.
// maxR2Bits (periodic): per atom, the largest squared separation to any of its excluded partners, as float bits rounded
// up -- the pair kernel only probes the exclusion lists for pairs of that atom at or below it. With accumulate == false
// (ranks > 0 of a sharded evaluation) that is all the kernel produces.
__global__ void __launch_bounds__(128) exclusionKernel(int numExcl, int Npad, const int2* __restrict__ pairs,
        const double* __restrict__ pos, const double* __restrict__ q, const double2* __restrict__ lj,
        BoxD box, bool pbc, double alpha, bool forces, bool energy, bool accumulate,
        long long* __restrict__ forceFixed, long long* __restrict__ dedqFixed, long long* __restrict__ energyFixed,
        unsigned int* __restrict__ maxR2Bits) {
    __shared__ double scratch[32];
    const int e = blockIdx.x*blockDim.x + threadIdx.x;
    const double ke = CFX_ONE_4PI_EPS0;
    double en = 0.0;
    float r2up = 0.f;
    if (e < numExcl) {
        const int i = pairs[e].x, j = pairs[e].y;
        double3 d = deltaPeriodic(pos, j, i, box, pbc);            // pos[i] - pos[j]
        const double r2 = dot3(d);
        r2up = __double2float_ru(r2);
        const double r = sqrt(r2), invR = 1.0/r;
        const double qi = q[i], qj = q[j];
        double dEdR, dqi, dqj;                                       // contributions to subtract
        if (pbc) {
            const double ar = alpha*r;
            const double erfv = erf(ar);
            dEdR = ke*qi*qj*invR*invR*invR*(erfv - ar*exp(-ar*ar)*1.1283791670955126);
            dqi = ke*qj*invR*erfv;
            dqj = ke*qi*invR*erfv;
            en = -ke*qi*qj*invR*erfv;
        }
        else {
            const double sig = lj[i].x + lj[j].x;
            double s2 = invR*sig; s2 *= s2;
            const double s6 = s2*s2*s2;
            const double es6 = s6*(lj[i].y*lj[j].y);
            dEdR = (es6*(12*s6 - 6) + ke*qi*qj*invR)*invR*invR;
            dqi = ke*qj*invR;
            dqj = ke*qi*invR;
            en = energy ? -(ke*qi*qj*invR + es6*(s6 - 1)) : 0.0;
        }
        if (forces && accumulate) {
            atomicAddFixed(forceFixed + i,          -dEdR*d.x);
            atomicAddFixed(forceFixed + Npad + i,   -dEdR*d.y);
            atomicAddFixed(forceFixed + 2*Npad + i, -dEdR*d.z);
            atomicAddFixed(forceFixed + j,           dEdR*d.x);
            atomicAddFixed(forceFixed + Npad + j,    dEdR*d.y);
            atomicAddFixed(forceFixed + 2*Npad + j,  dEdR*d.z);
            atomicAddFixed(dedqFixed + i, -dqi);
            atomicAddFixed(dedqFixed + j, -dqj);
        }
    }
    if (maxR2Bits && e < numExcl) {                                  // non-negative floats order like their bit patterns
        atomicMax(maxR2Bits + pairs[e].x, __float_as_uint(r2up));
        atomicMax(maxR2Bits + pairs[e].y, __float_as_uint(r2up));
    }
    if (!accumulate) return;                                         // (uniform over the grid)
    en = blockSum(en, scratch);
    if (threadIdx.x == 0) atomicAddEnergy(energyFixed + CFX_E_EXCL, en);
}

// Non-periodic branch: all pairs in FP64. One thread per atom i, j streamed through shared memory.
// The all-pairs and excluded sums cancel to ~1e-3 of their size (bonded pairs), hence FP64 throughout.
#define NOCUT_TILE 128
__global__ void __launch_bounds__(NOCUT_TILE) noCutoffKernel(int N, int Npad, const double* __restrict__ pos,
        const double* __restrict__ q, const double2* __restrict__ lj, bool forces, bool energy,
        long long* __restrict__ forceFixed, long long* __restrict__ dedqFixed, long long* __restrict__ energyFixed) {
    __shared__ double sx[NOCUT_TILE], sy[NOCUT_TILE], sz[NOCUT_TILE], sq[NOCUT_TILE];
    __shared__ double2 slj[NOCUT_TILE];
    __shared__ double scratch[32];
    const double ke = CFX_ONE_4PI_EPS0;
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    const bool valid = i < N;
    double xi = 0, yi = 0, zi = 0, qi = 0, hsi = 0, tei = 0;
    if (valid) { xi = pos[3*i]; yi = pos[3*i+1]; zi = pos[3*i+2]; qi = q[i]; hsi = lj[i].x; tei = lj[i].y; }
    double fx = 0, fy = 0, fz = 0, dq = 0, en = 0;
    for (int j0 = 0; j0 < N; j0 += NOCUT_TILE) {
        const int jl = j0 + threadIdx.x;
        __syncthreads();
        if (jl < N) {
            sx[threadIdx.x] = pos[3*jl]; sy[threadIdx.x] = pos[3*jl+1]; sz[threadIdx.x] = pos[3*jl+2];
            sq[threadIdx.x] = q[jl]; slj[threadIdx.x] = lj[jl];
        }
        __syncthreads();
        const int cnt = min(NOCUT_TILE, N - j0);
        if (valid)
            for (int k = 0; k < cnt; k++) {
                const int j = j0 + k;
                if (j == i) continue;
                // delta = pos[j] - pos[i] (getDeltaR(I,J) = J - I), force on i is -dEdR*delta
                const double dx = sx[k] - xi, dy = sy[k] - yi, dz = sz[k] - zi;
                const double invR = rsqrt(dx*dx + dy*dy + dz*dz);
                const double sig = hsi + slj[k].x;
                double s2 = invR*sig; s2 *= s2;
                const double s6 = s2*s2*s2;
                const double es6 = s6*(tei*slj[k].y);
                const double coul = ke*qi*sq[k]*invR;
                en += coul + es6*(s6 - 1);
                const double dEdR = (es6*(12*s6 - 6) + coul)*invR*invR;
                fx -= dEdR*dx; fy -= dEdR*dy; fz -= dEdR*dz;
                dq += ke*sq[k]*invR;
            }
    }
    if (valid && forces) {
        atomicAddFixed(forceFixed + i, fx);
        atomicAddFixed(forceFixed + Npad + i, fy);
        atomicAddFixed(forceFixed + 2*Npad + i, fz);
        atomicAddFixed(dedqFixed + i, dq);
    }
    en = blockSum(energy ? 0.5*en : 0.0, scratch);
    if (threadIdx.x == 0) atomicAddEnergy(energyFixed + CFX_E_DIRECT, en);
}

__global__ void __launch_bounds__(256) finalizeKernel(int N, int Npad, const long long* __restrict__ forceFixed,
        const long long* __restrict__ energyFixed, double* __restrict__ forceOut, double* __restrict__ energyOut) {
    const int i = blockIdx.x*blockDim.x + threadIdx.x;
    if (i < N) {
        #pragma unroll
        for (int c = 0; c < 3; c++)
            forceOut[3*(size_t) i + c] = (double) forceFixed[(size_t) c*Npad + i]*(1.0/CFX_FIXED_SCALE);
    }
    if (i == 0) {
        double tot = 0.0;
        for (int k = 0; k < 4; k++) {
            double v = (double) energyFixed[k]*(1.0/CFX_ENERGY_SCALE);
            energyOut[k] = v;
            tot += v;
        }
        energyOut[CFX_E_TOTAL] = tot;
    }
}

// Host-buffer entry point: the same conversion written straight into page-locked host memory (the caller's registered
// array: read-modify-write, `accumulate`; the handle's staging array: write only), energies and the list-overflow
// counter with it, so that the step's graph ends with one kernel instead of a kernel and three copy nodes.
__global__ void __launch_bounds__(256) finalizeToHostKernel(int N, int Npad, const long long* __restrict__ forceFixed,
        const long long* __restrict__ energyFixed, double* hostForce, bool accumulate, double* __restrict__ hostEnergy,
        const unsigned long long* __restrict__ overflowSrc, unsigned long long* __restrict__ hostOverflow) {
    const int k = blockIdx.x*blockDim.x + threadIdx.x;          // element of the [N][3] array: consecutive lanes, consecutive addresses
    if (hostForce && k < 3*N) {
        const int i = k/3, c = k - 3*i;
        const double v = (double) forceFixed[(size_t) c*Npad + i]*(1.0/CFX_FIXED_SCALE);
        hostForce[k] = accumulate ? hostForce[k] + v : v;
    }
    if (k == 0) {
        double tot = 0.0;
        for (int e = 0; e < 4; e++) {
            const double v = (double) energyFixed[e]*(1.0/CFX_ENERGY_SCALE);
            hostEnergy[e] = v;
            tot += v;
        }
        hostEnergy[CFX_E_TOTAL] = tot;
        if (overflowSrc) *hostOverflow = *overflowSrc;
    }
}

inline BoxD boxOf(const State& st) { return BoxD{st.box.L[0], st.box.L[1], st.box.L[2]}; }

} // namespace

void launchFluxAssembly(State& st, const double* dPos, cudaStream_t s) {
    if (st.numTerms > 0) {
        fluxTermKernel<<<(st.numTerms + 127)/128, 128, 0, s>>>(st.nb, st.na, st.nw, st.termIdx, st.termPar, dPos,
                boxOf(st), st.pbc, st.dqSlot, st.rowVal);
        CFX_LAUNCH_CHECK(); st.launches++;
        mark(st, "flux_terms", s);
    }
    const bool selfTerm = st.pbc && st.shardRank == 0;
    const double selfCoef = CFX_ONE_4PI_EPS0*st.alpha/sqrt(M_PI);
    chargeSumSelfKernel<<<(st.Npad + 255)/256, 256, 0, s>>>(st.N, st.Npad, st.q0, st.qcsrPtr, st.qcsrSlot, st.qcsrCoef,
            st.dqSlot, st.q, st.qf, selfTerm, selfCoef, st.dedqFixed, st.energyFixed);
    CFX_LAUNCH_CHECK(); st.launches++;
    mark(st, "charge_sum_self", s);
}

void launchChainRule(State& st, long long* dForce, const long long* dDedq, cudaStream_t s) {
    if (st.P == 0) return;
    chainRuleKernel<<<(st.P + 255)/256, 256, 0, s>>>(st.P, st.Npad, st.rowDq, st.rowDx, st.rowVal, dDedq, dForce);
    CFX_LAUNCH_CHECK(); st.launches++;
    mark(st, "chain_rule", s);
}

// Periodic branch. Must run BEFORE launchDirect on the same stream: it leaves every atom's largest excluded-partner r2
// in exclMaxR2 for the pair kernel (every rank of a sharded evaluation needs that; only rank 0 accumulates).
void launchExclusionCorrection(State& st, const double* dPos, bool forces, long long* dForce, long long* dDedq, cudaStream_t s) {
    CFX_CUDA(cudaMemsetAsync(st.exclMaxR2, 0, sizeof(unsigned int)*st.Npad, s));
    if (st.numExcl == 0) return;
    exclusionKernel<<<(st.numExcl + 127)/128, 128, 0, s>>>(st.numExcl, st.Npad, st.exclPairs, dPos, st.q, st.ljd,
            boxOf(st), st.pbc, st.alpha, forces, true, st.shardRank == 0, dForce, dDedq, st.energyFixed,
            st.exclMaxR2);
    CFX_LAUNCH_CHECK(); st.launches++;
    mark(st, "exclusion_corr", s);
}

void launchNoCutoff(State& st, const double* dPos, bool forces, bool energy, long long* dForce, long long* dDedq, cudaStream_t s) {
    noCutoffKernel<<<(st.N + NOCUT_TILE - 1)/NOCUT_TILE, NOCUT_TILE, 0, s>>>(st.N, st.Npad, dPos, st.q, st.ljd, forces, energy,
            dForce, dDedq, st.energyFixed);
    CFX_LAUNCH_CHECK(); st.launches++;
    mark(st, "nocutoff_pairs", s);
    if (st.numExcl > 0) {
        exclusionKernel<<<(st.numExcl + 127)/128, 128, 0, s>>>(st.numExcl, st.Npad, st.exclPairs, dPos, st.q, st.ljd,
                boxOf(st), false, 0.0, forces, energy, true, dForce, dDedq, st.energyFixed, nullptr);
        CFX_LAUNCH_CHECK(); st.launches++;
        mark(st, "nocutoff_excl", s);
    }
}

void launchFinalize(State& st, const long long* dForce, const long long* dEnergyFixed, cudaStream_t s) {
    finalizeKernel<<<(st.N + 255)/256, 256, 0, s>>>(st.N, st.Npad, dForce, dEnergyFixed, st.forceOut, st.energyOut);
    CFX_LAUNCH_CHECK(); st.launches++;
    mark(st, "finalize", s);
}

void launchFinalizeToHost(State& st, const long long* dForce, const long long* dEnergyFixed, double* hostForce, bool accumulate,
                          cudaStream_t s) {
    finalizeToHostKernel<<<(3*st.N + 255)/256, 256, 0, s>>>(st.N, st.Npad, dForce, dEnergyFixed, hostForce, accumulate, st.hEnergy,
            (st.pbc && st.pairCounters) ? st.pairCounters + 11 : nullptr, st.hListOverflow);
    CFX_LAUNCH_CHECK(); st.launches++;
    mark(st, "finalize", s);
}

} // namespace cfx

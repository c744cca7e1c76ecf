// platform.cu -- zero-copy binding to a CUDA-platform host. What the reference's
// CudaCalcCoulForceKernel::execute does around its kernels (platforms/cuda/src/CudaCoulKernels.cpp:523-660,
// kernels/PBCForce.cu:817-825 genIndexAtom): OpenMM's CUDA platform keeps positions + charge as `real4 posq[paddedN]`
// in ITS OWN (spatially re-sorted) atom order, with atomIndex[platform slot] = user index, and accumulates forces in a
// 64-bit fixed-point buffer [3][paddedN] (value * 2^32, PBCForce.cu:336-338) in that same order. Here one gather kernel
// brings the positions into the library's user-order FP64 array, the evaluation runs unchanged, and one scatter kernel
// adds the forces into the platform's buffer -- all on the platform's stream, replayed as one CUDA graph.
#include "cfx_internal.cuh"

#include <stdexcept>

namespace cfx {

namespace {

template <class T4>
__global__ void __launch_bounds__(256) gatherPosqKernel(int N, const T4* __restrict__ posq, const float4* __restrict__ correction,
        const int* __restrict__ atomIndex, double* __restrict__ pos) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= N) return;
    const int u = atomIndex[s];
    const T4 p = posq[s];
    double x = (double) p.x, y = (double) p.y, z = (double) p.z;
    if (correction) {                                  // mixed precision: low-order bits kept beside the float4
        const float4 c = correction[s];
        x += (double) c.x; y += (double) c.y; z += (double) c.z;
    }
    pos[3*(size_t) u] = x; pos[3*(size_t) u + 1] = y; pos[3*(size_t) u + 2] = z;
}

__global__ void __launch_bounds__(256) scatterForceKernel(int N, int Npad, int paddedN, const int* __restrict__ atomIndex,
        const long long* __restrict__ forceFixed, unsigned long long* __restrict__ forceBuffers) {
    const int s = blockIdx.x*blockDim.x + threadIdx.x;
    if (s >= N) return;
    const int u = atomIndex[s];
    #pragma unroll
    for (int c = 0; c < 3; c++)
        atomicAdd(forceBuffers + (size_t) c*paddedN + s, static_cast<unsigned long long>(forceFixed[(size_t) c*Npad + u]));
}

template <class E>
__global__ void addEnergyToPlatformKernel(const long long* __restrict__ energyFixed, E* __restrict__ energyBuffer) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double tot = 0.0;
        for (int k = 0; k < 4; k++) tot += (double) energyFixed[k]*(1.0/CFX_ENERGY_SCALE);
        energyBuffer[0] += (E) tot;
    }
}

} // namespace

} // namespace cfx

using namespace cfx;

extern "C" int cfx_execute_platform(cfx_handle* h, const void* d_posq, int posq_is_double, const void* d_posq_correction,
                                    const int32_t* d_atom_index, int32_t padded_num_atoms, const double* box,
                                    int include_forces, int include_energy, unsigned long long* d_force_buffers,
                                    void* d_energy_buffer, int energy_is_double, void* stream) {
    try {
        if (!h || !d_posq || !d_atom_index || !d_force_buffers) { setLastError("null argument"); return CFX_ERR_ARGUMENT; }
        State& st = h->st;
        if (st.shardCount != 1) { setLastError("cfx_execute_platform evaluates whole systems"); return CFX_ERR_ARGUMENT; }
        if (padded_num_atoms < st.N) { setLastError("padded_num_atoms is smaller than the particle count"); return CFX_ERR_ARGUMENT; }
        if (st.N == 0) return CFX_OK;
        CFX_CUDA(cudaSetDevice(st.device));
        if (st.pbc) {
            if (!box) { setLastError("null box for a periodic system"); return CFX_ERR_ARGUMENT; }
            ensureBox(st, box);
        }
        cudaStream_t s = static_cast<cudaStream_t>(stream);
        const bool incF = include_forces != 0, incE = include_energy != 0;
        const int blocks = (st.N + 255)/256;
        auto enqueueAll = [&]() {
            if (posq_is_double)
                gatherPosqKernel<double4><<<blocks, 256, 0, s>>>(st.N, static_cast<const double4*>(d_posq), nullptr, d_atom_index, st.pos);
            else
                gatherPosqKernel<float4><<<blocks, 256, 0, s>>>(st.N, static_cast<const float4*>(d_posq),
                        static_cast<const float4*>(d_posq_correction), d_atom_index, st.pos);
            CFX_LAUNCH_CHECK(); st.launches++;
            CFX_CUDA(cudaMemsetAsync(st.forceFixed, 0, sizeof(long long)*3*st.Npad, s));
            enqueueEvaluation(st, st.pos, incF, incE, st.forceFixed, s, false);
            scatterForceKernel<<<blocks, 256, 0, s>>>(st.N, st.Npad, padded_num_atoms, d_atom_index, st.forceFixed, d_force_buffers);
            CFX_LAUNCH_CHECK(); st.launches++;
            if (d_energy_buffer) {
                if (energy_is_double) addEnergyToPlatformKernel<double><<<1, 32, 0, s>>>(st.energyFixed, static_cast<double*>(d_energy_buffer));
                else                  addEnergyToPlatformKernel<float><<<1, 32, 0, s>>>(st.energyFixed, static_cast<float*>(d_energy_buffer));
                CFX_LAUNCH_CHECK(); st.launches++;
            }
        };
        st.launches = 0;
        const bool capturable = s != nullptr && s != cudaStreamLegacy;
        if (st.useGraph && capturable) {
            State::PlatformGraphKey key{d_posq, d_posq_correction, d_atom_index, d_force_buffers, d_energy_buffer, padded_num_atoms,
                                        (incF ? 1 : 0) | (incE ? 2 : 0) | (posq_is_double ? 4 : 0) | (energy_is_double ? 8 : 0), st.planGeneration};
            const State::PlatformGraphKey& o = st.platKey;
            const bool same = st.platGraph && o.posq == key.posq && o.corr == key.corr && o.index == key.index && o.force == key.force &&
                              o.energy == key.energy && o.padded == key.padded && o.flags == key.flags && o.planGen == key.planGen;
            if (!same) {
                if (st.platGraph) { cudaGraphExecDestroy(st.platGraph); st.platGraph = nullptr; }
                cudaGraph_t graph;
                CFX_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
                try { enqueueAll(); }
                catch (...) { cudaGraph_t dead; cudaStreamEndCapture(s, &dead); throw; }
                CFX_CUDA(cudaStreamEndCapture(s, &graph));
                CFX_CUDA(cudaGraphInstantiate(&st.platGraph, graph, 0));
                CFX_CUDA(cudaGraphDestroy(graph));
                key.planGen = st.planGeneration;
                st.platKey = key;
                st.platGraphLaunches = st.launches;
            }
            st.launches = st.platGraphLaunches;
            CFX_CUDA(cudaGraphLaunch(st.platGraph, s));
        }
        else
            enqueueAll();
        st.evaluated = true;
        st.stagedPosCurrent = false;
        return CFX_OK;
    } catch (const std::exception& e) { setLastError(e.what()); return CFX_ERR_CUDA; }
}

/* Runtime half of the CUDA-platform stand-in (shim/openmm/cuda/): device allocations through the CUDA runtime API. */
#include "openmm/cuda/CudaContext.h"
#include <cuda_runtime.h>

namespace OpenMM {

namespace {
void check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw OpenMMException(std::string("CUDA stand-in: ") + what + ": " + cudaGetErrorString(e));
}
}

CudaArray::~CudaArray() {
    if (pointer) cudaFree(reinterpret_cast<void*>(pointer));
}

void CudaArray::initialize(CudaContext& context, int size_, int elementSize_, const std::string& name_) {
    if (pointer) throw OpenMMException("CudaArray has already been initialized: " + name_);
    context.setAsCurrent();
    void* p = nullptr;
    check(cudaMalloc(&p, (size_t) std::max(size_, 1)*elementSize_), "cudaMalloc");
    check(cudaMemset(p, 0, (size_t) std::max(size_, 1)*elementSize_), "cudaMemset");
    pointer = reinterpret_cast<CUdeviceptr>(p);
    size = size_; elementSize = elementSize_; name = name_;
}

void CudaArray::upload(const void* data, bool) {
    check(cudaMemcpy(reinterpret_cast<void*>(pointer), data, (size_t) size*elementSize, cudaMemcpyHostToDevice), "upload");
}

void CudaArray::download(void* data, bool) const {
    check(cudaMemcpy(data, reinterpret_cast<const void*>(pointer), (size_t) size*elementSize, cudaMemcpyDeviceToHost), "download");
}

CudaContext::CudaContext(int numAtoms, int deviceIndex, bool useDoublePrecision, bool useMixedPrecision) :
        numAtoms(numAtoms), paddedNumAtoms((numAtoms + TileSize - 1)/TileSize*TileSize), deviceIndex(deviceIndex),
        useDouble(useDoublePrecision), useMixed(useMixedPrecision), stream(nullptr) {
    setAsCurrent();
    cudaStream_t s;
    check(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate");
    stream = reinterpret_cast<CUstream>(s);
    posq.initialize(*this, paddedNumAtoms, useDouble ? (int) sizeof(double4) : (int) sizeof(float4), "posq");
    if (useMixed) posqCorrection.initialize(*this, paddedNumAtoms, (int) sizeof(float4), "posqCorrection");
    force.initialize(*this, 3*paddedNumAtoms, (int) sizeof(long long), "force");
    energyBuffer.initialize(*this, 1024, (useDouble || useMixed) ? (int) sizeof(double) : (int) sizeof(float), "energyBuffer");
    atomIndex.initialize(*this, paddedNumAtoms, (int) sizeof(int), "atomIndex");
}

CudaContext::~CudaContext() {
    for (CudaForceInfo* f : forceInfos) delete f;
    if (stream) cudaStreamDestroy(reinterpret_cast<cudaStream_t>(stream));
}

void CudaContext::setAsCurrent() {
    check(cudaSetDevice(deviceIndex), "cudaSetDevice");
}

} // namespace OpenMM

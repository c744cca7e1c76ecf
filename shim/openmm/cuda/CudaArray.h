#ifndef OPENMM_CUDAARRAY_H_
#define OPENMM_CUDAARRAY_H_
#include "openmm/OpenMMException.h"
#include "openmm/internal/windowsExport.h"
#include <cuda.h>
#include <string>
#include <vector>
namespace OpenMM {
class CudaContext;
/* Stand-in for OpenMM::CudaArray (OpenMM 7.x): a typed device allocation owned by the platform. Only what the plugin's
 * CUDA kernels touch (platforms/cuda/src/CudaCoulKernels.cpp: initialize, upload, getDevicePointer, getSize). */
class OPENMM_EXPORT CudaArray {
public:
    CudaArray() : pointer(0), size(0), elementSize(0) {}
    ~CudaArray();
    void initialize(CudaContext& context, int size, int elementSize, const std::string& name);
    template <class T> void initialize(CudaContext& context, int size, const std::string& name) { initialize(context, size, (int) sizeof(T), name); }
    bool isInitialized() const { return pointer != 0; }
    int getSize() const { return size; }
    int getElementSize() const { return elementSize; }
    const std::string& getName() const { return name; }
    CUdeviceptr& getDevicePointer() { return pointer; }
    void upload(const void* data, bool blocking = true);
    void download(void* data, bool blocking = true) const;
    template <class T> void upload(const std::vector<T>& data) {
        if (sizeof(T) != (size_t) elementSize || (int) data.size() != size) throw OpenMMException("CudaArray::upload: size mismatch for " + name);
        upload(data.data());
    }
    template <class T> void download(std::vector<T>& data) const {
        if (sizeof(T) != (size_t) elementSize) throw OpenMMException("CudaArray::download: size mismatch for " + name);
        data.resize(size);
        download(data.data());
    }
private:
    CUdeviceptr pointer;
    int size, elementSize;
    std::string name;
};
} // namespace OpenMM
#endif

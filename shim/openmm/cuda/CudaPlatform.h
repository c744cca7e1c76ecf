#ifndef OPENMM_CUDAPLATFORM_H_
#define OPENMM_CUDAPLATFORM_H_
#include "openmm/Platform.h"
#include "openmm/cuda/CudaContext.h"
#include <vector>
namespace OpenMM {
/* Stand-in for OpenMM::CudaPlatform: the name plugins look up (platforms/cuda/src/CudaCoulKernelFactory.cpp:20) and the
 * PlatformData they reach the CudaContext through (:40). */
class OPENMM_EXPORT CudaPlatform : public Platform {
public:
    class PlatformData;
    const std::string& getName() const { static const std::string name = "CUDA"; return name; }
};
class CudaPlatform::PlatformData {
public:
    std::vector<CudaContext*> contexts;
    ~PlatformData() { for (CudaContext* c : contexts) delete c; }
};
} // namespace OpenMM
#endif

#ifndef OPENMM_CUDAFORCEINFO_H_
#define OPENMM_CUDAFORCEINFO_H_
#include "openmm/internal/windowsExport.h"
#include <vector>
namespace OpenMM {
/* Stand-in for OpenMM::CudaForceInfo: what a force tells the CUDA platform so that it may reorder atoms (molecules
 * whose particles and groups are "identical" can be swapped). Interface as used by
 * platforms/cuda/src/CudaCoulKernels.cpp:20-47. */
class OPENMM_EXPORT CudaForceInfo {
public:
    virtual ~CudaForceInfo() {}
    virtual bool areParticlesIdentical(int particle1, int particle2) { return true; }
    virtual int getNumParticleGroups() { return 0; }
    virtual void getParticlesInGroup(int index, std::vector<int>& particles) {}
    virtual bool areGroupsIdentical(int group1, int group2) { return true; }
};
} // namespace OpenMM
#endif

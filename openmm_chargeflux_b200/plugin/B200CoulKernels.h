#ifndef B200_COUL_KERNELS_H_
#define B200_COUL_KERNELS_H_

/* The B200 implementation of the plugin's CalcCoulForceKernel (openmmapi/include/CoulKernels.h:15-38).
 *
 * Host C++ only: it gathers the CoulForce parameters through the plugin's unchanged getters, hands
 * them to the C ABI of libcfx_b200.so (include/cfx_b200.h) and moves positions / forces between the
 * platform's buffers and the library. It replaces platforms/cuda/src/CudaCoulKernels.{h,cpp} of the
 * reference (CudaCalcCoulForceKernel), not a line of which is reused.
 *
 * This first adapter binds a platform that keeps positions and forces on the host
 * (ReferencePlatform::PlatformData, i.e. OpenMM's Reference and CPU platforms): cfx_execute() copies
 * positions host->device and forces device->host every step. The zero-copy binding to OpenMM's CUDA
 * platform (posq / atomIndex / fixed-point force buffers -> cfx_execute_device) is INTEGRATION.md "next".
 */
#include "CoulKernels.h"
#include "cfx_b200.h"
#include "openmm/Platform.h"
#include <string>
#include <vector>

namespace CoulPlugin {

class B200CalcCoulForceKernel : public CalcCoulForceKernel {
public:
    B200CalcCoulForceKernel(std::string name, const OpenMM::Platform& platform) : CalcCoulForceKernel(name, platform), handle(NULL) {
    }
    ~B200CalcCoulForceKernel();
    /** Same contract as CalcCoulForceKernel::initialize. */
    void initialize(const OpenMM::System& system, const CoulForce& force);
    /** Same contract as CalcCoulForceKernel::execute: adds to the platform's forces, returns kJ/mol. */
    double execute(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy);
    /** Energy components of the last execute (self, recip, direct, excl, total). */
    const double* getLastEnergyComponents() const { return lastEnergy; }
private:
    cfx_handle* handle;
    int numParticles;
    double lastEnergy[CFX_E_COUNT];
};

} // namespace CoulPlugin

#endif /*B200_COUL_KERNELS_H_*/

/* Plugin entry points: the loader contract of platforms/cuda/src/CudaCoulKernelFactory.cpp:13-44
 * (registerPlatforms / registerKernelFactories exported with C linkage; the factory object is owned by
 * the Platform for the life of the process; an unknown kernel name throws OpenMMException).
 *
 * Two bindings are registered:
 *   "CUDA" platform     B200CudaCalcCoulForceKernel: zero-copy, the platform's device buffers and stream
 *                       (what platforms/cuda/src/CudaCoulKernelFactory.cpp:20-22,39-44 does for the reference's kernel)
 *   host platform       B200CalcCoulForceKernel on the platform named by CFX_B200_PLATFORM (default "B200", then "CPU",
 *                       then "Reference"): any platform whose PlatformData keeps positions/forces on the host. */
#include "B200CoulKernelFactory.h"
#include "B200CoulKernels.h"
#include "B200CudaCoulKernels.h"
#include "openmm/cuda/CudaPlatform.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include <cstdlib>
#include <exception>
#include <string>

using namespace CoulPlugin;
using namespace OpenMM;

extern "C" OPENMM_EXPORT void registerPlatforms() {
}

namespace {
/* createKernelImpl for the CUDA platform: the CudaContext is reached as the reference does (CudaCoulKernelFactory.cpp:40). */
class B200CudaCoulKernelFactory : public KernelFactory {
public:
    KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const {
        CudaPlatform::PlatformData& data = *static_cast<CudaPlatform::PlatformData*>(context.getPlatformData());
        if (data.contexts.empty())
            throw OpenMMException("B200 CoulForce plugin: the CUDA platform holds no CudaContext");
        if (name == CalcCoulForceKernel::Name())
            return new B200CudaCalcCoulForceKernel(name, platform, *data.contexts[0]);
        throw OpenMMException((std::string("Tried to create kernel with illegal kernel name '")+name+"'").c_str());
    }
};
}

extern "C" OPENMM_EXPORT void registerKernelFactories() {
    try {
        Platform& cuda = Platform::getPlatformByName("CUDA");
        cuda.registerKernelFactory(CalcCoulForceKernel::Name(), new B200CudaCoulKernelFactory());
    }
    catch (std::exception& ex) {
        // no CUDA platform in this process
    }
    const char* wanted = getenv("CFX_B200_PLATFORM");
    const char* names[] = {wanted, "B200", "CPU", "Reference"};
    for (const char* name : names) {
        if (name == NULL)
            continue;
        try {
            Platform& platform = Platform::getPlatformByName(name);
            platform.registerKernelFactory(CalcCoulForceKernel::Name(), new B200CoulKernelFactory());
            return;
        }
        catch (std::exception& ex) {
            // platform not present: try the next one
        }
    }
}

extern "C" OPENMM_EXPORT void registerCoulB200KernelFactories() {
    registerKernelFactories();
}

KernelImpl* B200CoulKernelFactory::createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const {
    if (name == CalcCoulForceKernel::Name())
        return new B200CalcCoulForceKernel(name, platform);
    throw OpenMMException((std::string("Tried to create kernel with illegal kernel name '")+name+"'").c_str());
}

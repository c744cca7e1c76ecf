#ifndef OPENMM_KERNELIMPL_H_
#define OPENMM_KERNELIMPL_H_
#include "internal/windowsExport.h"
#include <string>
namespace OpenMM {
class Platform;
/* Stand-in for OpenMM::KernelImpl: name + platform + the reference count the Kernel handle uses. */
class KernelImpl {
public:
    KernelImpl(std::string name, const Platform& platform) : name(name), platform(&platform), referenceCount(0) {}
    virtual ~KernelImpl() {}
    std::string getName() const { return name; }
    const Platform& getPlatform() { return *platform; }
private:
    friend class Kernel;
    std::string name;
    const Platform* platform;
    int referenceCount;
};
} // namespace OpenMM
#endif

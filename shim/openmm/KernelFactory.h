#ifndef OPENMM_KERNELFACTORY_H_
#define OPENMM_KERNELFACTORY_H_
#include "KernelImpl.h"
namespace OpenMM {
class ContextImpl;
class KernelFactory {
public:
    virtual ~KernelFactory() {}
    virtual KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const = 0;
};
} // namespace OpenMM
#endif

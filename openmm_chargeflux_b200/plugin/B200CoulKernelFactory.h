#ifndef OPENMM_B200_COUL_KERNEL_FACTORY_H_
#define OPENMM_B200_COUL_KERNEL_FACTORY_H_

#include "openmm/KernelFactory.h"

namespace OpenMM {

/** Creates the B200 CalcCoulForce kernel (counterpart of platforms/cuda/include/CudaCoulKernelFactory.h). */
class B200CoulKernelFactory : public KernelFactory {
public:
    KernelImpl* createKernelImpl(std::string name, const Platform& platform, ContextImpl& context) const;
};

} // namespace OpenMM

#endif /*OPENMM_B200_COUL_KERNEL_FACTORY_H_*/

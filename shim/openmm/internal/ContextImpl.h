#ifndef OPENMM_CONTEXTIMPL_H_
#define OPENMM_CONTEXTIMPL_H_
#include "openmm/Platform.h"
#include "openmm/System.h"
#include "openmm/Force.h"
#include "openmm/internal/ForceImpl.h"
#include <vector>
namespace OpenMM {
/* Stand-in for OpenMM::ContextImpl: owns the ForceImpls of a System on one Platform and the
 * platform data pointer kernels read positions/forces/box through. */
class OPENMM_EXPORT ContextImpl {
public:
    ContextImpl(const System& system, Platform& platform, void* platformData);
    ~ContextImpl();
    const System& getSystem() const { return system; }
    Platform& getPlatform() { return *platform; }
    void* getPlatformData() { return platformData; }
    const void* getPlatformData() const { return platformData; }
    void setPlatformData(void* data) { platformData = data; }
    /** Sum of calcForcesAndEnergy over all ForceImpls (forces accumulate in the platform data). */
    double calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups = 0xFFFFFFFF);
    std::vector<ForceImpl*>& getForceImpls() { return forceImpls; }
private:
    const System& system;
    Platform* platform;
    void* platformData;
    std::vector<ForceImpl*> forceImpls;
};
} // namespace OpenMM
#endif

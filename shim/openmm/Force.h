#ifndef OPENMM_FORCE_H_
#define OPENMM_FORCE_H_
#include "internal/windowsExport.h"
namespace OpenMM {
class ForceImpl;
class ContextImpl;
/* Stand-in for OpenMM::Force: force group, virtual createImpl(), PBC query. */
class Force {
public:
    Force() : forceGroup(0) {}
    virtual ~Force() {}
    int getForceGroup() const { return forceGroup; }
    void setForceGroup(int group) { forceGroup = group; }
    virtual bool usesPeriodicBoundaryConditions() const { return false; }
protected:
    friend class ContextImpl;
    virtual ForceImpl* createImpl() const = 0;
private:
    int forceGroup;
};
} // namespace OpenMM
#endif

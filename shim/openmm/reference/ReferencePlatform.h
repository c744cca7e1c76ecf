#ifndef OPENMM_REFERENCEPLATFORM_H_
#define OPENMM_REFERENCEPLATFORM_H_
#include "openmm/Platform.h"
#include "openmm/System.h"
#include "openmm/Vec3.h"
#include <vector>
namespace OpenMM {
/* Stand-in for OpenMM::ReferencePlatform. PlatformData keeps the untyped-pointer layout of the
 * OpenMM 7.x line (the plugin casts positions/forces to vector<Vec3>* and the box to Vec3*). */
class OPENMM_EXPORT ReferencePlatform : public Platform {
public:
    class PlatformData;
    ReferencePlatform() {}
    const std::string& getName() const { static const std::string name = "Reference"; return name; }
};
class ReferencePlatform::PlatformData {
public:
    PlatformData(int numParticles) : numParticles(numParticles), stepCount(0), time(0.0) {
        positions = new std::vector<Vec3>(numParticles);
        velocities = new std::vector<Vec3>(numParticles);
        forces = new std::vector<Vec3>(numParticles);
        periodicBoxSize = new Vec3();
        periodicBoxVectors = new Vec3[3];
    }
    ~PlatformData() {
        delete (std::vector<Vec3>*) positions;
        delete (std::vector<Vec3>*) velocities;
        delete (std::vector<Vec3>*) forces;
        delete (Vec3*) periodicBoxSize;
        delete[] (Vec3*) periodicBoxVectors;
    }
    int numParticles, stepCount;
    double time;
    void* positions;
    void* velocities;
    void* forces;
    void* periodicBoxSize;
    void* periodicBoxVectors;
};
} // namespace OpenMM
#endif

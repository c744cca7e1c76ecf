#ifndef OPENMM_SYSTEM_H_
#define OPENMM_SYSTEM_H_
#include "Vec3.h"
#include <vector>
namespace OpenMM {
class Force;
/* Stand-in for OpenMM::System: particle count, default box, force list (owned). */
class System {
public:
    System() { box[0] = Vec3(2,0,0); box[1] = Vec3(0,2,0); box[2] = Vec3(0,0,2); }
    ~System();
    int getNumParticles() const { return (int) masses.size(); }
    int addParticle(double mass) { masses.push_back(mass); return (int) masses.size()-1; }
    double getParticleMass(int i) const { return masses[i]; }
    void getDefaultPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const { a = box[0]; b = box[1]; c = box[2]; }
    void setDefaultPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c) { box[0] = a; box[1] = b; box[2] = c; }
    int addForce(Force* force) { forces.push_back(force); return (int) forces.size()-1; }
    int getNumForces() const { return (int) forces.size(); }
    Force& getForce(int i) { return *forces[i]; }
    const Force& getForce(int i) const { return *forces[i]; }
private:
    std::vector<double> masses;
    std::vector<Force*> forces;
    Vec3 box[3];
};
} // namespace OpenMM
#endif

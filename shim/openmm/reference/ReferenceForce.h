#ifndef OPENMM_REFERENCEFORCE_H_
#define OPENMM_REFERENCEFORCE_H_
#include "openmm/Vec3.h"
#include <cmath>
namespace OpenMM {
/* Stand-in for OpenMM::ReferenceForce displacement helpers, restated from their documented
 * semantics: delta = J - I; periodic form subtracts box vector c*floor(dz/cz+0.5), then b, then a;
 * array form is [dx,dy,dz,r2,r]. */
class ReferenceForce {
public:
    static const int XIndex = 0;
    static const int YIndex = 1;
    static const int ZIndex = 2;
    static const int R2Index = 3;
    static const int RIndex = 4;
    static const int LastDeltaRIndex = 5;
    static Vec3 getDeltaR(const Vec3& atomCoordinatesI, const Vec3& atomCoordinatesJ) {
        return atomCoordinatesJ - atomCoordinatesI;
    }
    static Vec3 getDeltaRPeriodic(const Vec3& atomCoordinatesI, const Vec3& atomCoordinatesJ, const Vec3* boxVectors) {
        Vec3 diff = atomCoordinatesJ - atomCoordinatesI;
        diff -= boxVectors[2]*floor(diff[2]/boxVectors[2][2]+0.5);
        diff -= boxVectors[1]*floor(diff[1]/boxVectors[1][1]+0.5);
        diff -= boxVectors[0]*floor(diff[0]/boxVectors[0][0]+0.5);
        return diff;
    }
    static void getDeltaR(const Vec3& atomCoordinatesI, const Vec3& atomCoordinatesJ, double* deltaR) {
        fill(getDeltaR(atomCoordinatesI, atomCoordinatesJ), deltaR);
    }
    static void getDeltaRPeriodic(const Vec3& atomCoordinatesI, const Vec3& atomCoordinatesJ, const Vec3* boxVectors, double* deltaR) {
        fill(getDeltaRPeriodic(atomCoordinatesI, atomCoordinatesJ, boxVectors), deltaR);
    }
private:
    static void fill(const Vec3& d, double* deltaR) {
        deltaR[XIndex] = d[0];
        deltaR[YIndex] = d[1];
        deltaR[ZIndex] = d[2];
        deltaR[R2Index] = d[0]*d[0] + d[1]*d[1] + d[2]*d[2];
        deltaR[RIndex] = sqrt(deltaR[R2Index]);
    }
};
} // namespace OpenMM
#endif

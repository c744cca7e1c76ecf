#ifndef OPENMM_REFERENCE_NEIGHBORLIST_H_
#define OPENMM_REFERENCE_NEIGHBORLIST_H_
#include "openmm/Vec3.h"
#include "openmm/internal/windowsExport.h"
#include <set>
#include <utility>
#include <vector>
namespace OpenMM {
typedef std::pair<int, int> AtomPair;
typedef std::vector<AtomPair> NeighborList;
/* Stand-in for OpenMM's voxel-hash neighbour list, restated from its documented result: every pair
 * (i<j) that is not excluded and whose periodic (floor-based, c then b then a) or plain squared
 * distance is <= maxDistance^2 and >= minDistance^2. Pair order is unspecified. */
void OPENMM_EXPORT computeNeighborListVoxelHash(NeighborList& neighborList, int nAtoms,
        const std::vector<Vec3>& atomLocations, const std::vector<std::set<int> >& exclusions,
        const Vec3* periodicBoxVectors, bool usePeriodic, double maxDistance, double minDistance = 0.0,
        bool reportSymmetricPairs = false);
} // namespace OpenMM
#endif

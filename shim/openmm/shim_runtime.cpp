/* Runtime half of the OpenMM stand-in (see shim/README.md).
 *
 * OpenMM itself is not installed in this image, so the unmodified plugin sources under
 * /root/reference (openmmapi/ + platforms/reference/) and this repository's own plugin adapter are
 * compiled against these few classes instead. Only what those sources touch is provided:
 * the static Platform registry, per-platform kernel factories, ContextImpl's ForceImpl ownership,
 * and the neighbour list helper the Reference platform kernel calls.
 */
#include "openmm/Platform.h"
#include "openmm/System.h"
#include "openmm/Force.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/reference/ReferenceNeighborList.h"
#include <algorithm>
#include <cmath>
#include <dlfcn.h>
#include <set>

namespace OpenMM {

System::~System() {
    for (Force* f : forces)
        delete f;
}

Platform::~Platform() {
    std::set<KernelFactory*> unique;
    for (auto& kv : kernelFactories)
        unique.insert(kv.second);
    for (KernelFactory* f : unique)
        delete f;
}

std::vector<Platform*>& Platform::getPlatforms() {
    static std::vector<Platform*> platforms;
    return platforms;
}

void Platform::registerKernelFactory(const std::string& name, KernelFactory* factory) {
    kernelFactories[name] = factory;
}

Kernel Platform::createKernel(const std::string& name, ContextImpl& context) const {
    auto it = kernelFactories.find(name);
    if (it == kernelFactories.end())
        throw OpenMMException("Called createKernel() on a Platform which does not support the requested kernel");
    return Kernel(it->second->createKernelImpl(name, *this, context));
}

void Platform::registerPlatform(Platform* platform) {
    getPlatforms().push_back(platform);
}

int Platform::getNumPlatforms() {
    return (int) getPlatforms().size();
}

Platform& Platform::getPlatform(int index) {
    if (index < 0 || index >= getNumPlatforms())
        throw OpenMMException("Invalid platform index");
    return *getPlatforms()[index];
}

Platform& Platform::getPlatformByName(const std::string& name) {
    for (Platform* p : getPlatforms())
        if (p->getName() == name)
            return *p;
    throw OpenMMException("There is no registered Platform called \"" + name + "\"");
}

void Platform::loadPluginLibrary(const std::string& file) {
    // Same contract as OpenMM's loader: dlopen, then call registerPlatforms() and
    // registerKernelFactories() if the library exports them.
    void* handle = dlopen(file.c_str(), RTLD_LAZY | RTLD_LOCAL);
    if (handle == NULL)
        throw OpenMMException("Error loading library " + file + ": " + dlerror());
    void (*init)();
    *(void**) (&init) = dlsym(handle, "registerPlatforms");
    if (init != NULL)
        (*init)();
    *(void**) (&init) = dlsym(handle, "registerKernelFactories");
    if (init != NULL)
        (*init)();
}

ContextImpl::ContextImpl(const System& system, Platform& platform, void* platformData)
        : system(system), platform(&platform), platformData(platformData) {
    for (int i = 0; i < system.getNumForces(); i++) {
        forceImpls.push_back(system.getForce(i).createImpl());
        forceImpls.back()->initialize(*this);
    }
}

ContextImpl::~ContextImpl() {
    for (ForceImpl* impl : forceImpls)
        delete impl;
}

double ContextImpl::calcForcesAndEnergy(bool includeForces, bool includeEnergy, int groups) {
    double energy = 0.0;
    for (ForceImpl* impl : forceImpls)
        energy += impl->calcForcesAndEnergy(*this, includeForces, includeEnergy, groups);
    return energy;
}

/* ---------------------------------------------------------------------------------------------
 * Neighbour list. The result is defined by the pair predicate only (see the header); a uniform
 * grid with cell edge >= maxDistance is used to enumerate candidates, every candidate is tested on
 * the ORIGINAL (unwrapped) coordinates with OpenMM's floor-based periodic difference.
 * ------------------------------------------------------------------------------------------- */
static inline double pairDistanceSquared(const Vec3& pos1, const Vec3& pos2, const Vec3* box, bool usePeriodic) {
    Vec3 diff = pos2 - pos1;
    if (usePeriodic) {
        diff -= box[2]*floor(diff[2]/box[2][2]+0.5);
        diff -= box[1]*floor(diff[1]/box[1][1]+0.5);
        diff -= box[0]*floor(diff[0]/box[0][0]+0.5);
    }
    return diff.dot(diff);
}

void computeNeighborListVoxelHash(NeighborList& neighborList, int nAtoms,
        const std::vector<Vec3>& atomLocations, const std::vector<std::set<int> >& exclusions,
        const Vec3* periodicBoxVectors, bool usePeriodic, double maxDistance, double minDistance,
        bool reportSymmetricPairs) {
    neighborList.clear();
    if (nAtoms < 2)
        return;
    const double maxD2 = maxDistance*maxDistance;
    const double minD2 = minDistance*minDistance;
    int nc[3];
    double lo[3], edge[3];
    if (usePeriodic) {
        if (periodicBoxVectors[0][1] != 0.0 || periodicBoxVectors[0][2] != 0.0 || periodicBoxVectors[1][0] != 0.0 ||
            periodicBoxVectors[1][2] != 0.0 || periodicBoxVectors[2][0] != 0.0 || periodicBoxVectors[2][1] != 0.0)
            throw OpenMMException("shim neighbour list: only rectangular periodic boxes are supported");
        for (int d = 0; d < 3; d++) {
            double len = periodicBoxVectors[d][d];
            nc[d] = std::max(1, (int) floor(len/maxDistance));
            nc[d] = std::min(nc[d], 256);
            lo[d] = 0.0;
            edge[d] = len/nc[d];
        }
    }
    else {
        double hi[3];
        for (int d = 0; d < 3; d++) {
            lo[d] = hi[d] = atomLocations[0][d];
            for (int i = 1; i < nAtoms; i++) {
                lo[d] = std::min(lo[d], atomLocations[i][d]);
                hi[d] = std::max(hi[d], atomLocations[i][d]);
            }
            nc[d] = std::max(1, std::min(256, (int) floor((hi[d]-lo[d])/maxDistance)));
            edge[d] = (hi[d]-lo[d])/nc[d];
            if (!(edge[d] > 0.0))
                edge[d] = 1.0;
        }
    }
    const int ncells = nc[0]*nc[1]*nc[2];
    std::vector<int> cellOf(nAtoms), cellStart(ncells+1, 0), order(nAtoms);
    for (int i = 0; i < nAtoms; i++) {
        int c[3];
        for (int d = 0; d < 3; d++) {
            double x = atomLocations[i][d] - lo[d];
            if (usePeriodic) {
                double len = periodicBoxVectors[d][d];
                x -= floor(x/len)*len;
            }
            int k = (int) floor(x/edge[d]);
            c[d] = std::max(0, std::min(nc[d]-1, k));
        }
        cellOf[i] = (c[0]*nc[1] + c[1])*nc[2] + c[2];
        cellStart[cellOf[i]+1]++;
    }
    for (int c = 0; c < ncells; c++)
        cellStart[c+1] += cellStart[c];
    std::vector<int> fill(cellStart.begin(), cellStart.end()-1);
    for (int i = 0; i < nAtoms; i++)
        order[fill[cellOf[i]]++] = i;

    std::vector<int> neighbours;
    for (int cx = 0; cx < nc[0]; cx++)
    for (int cy = 0; cy < nc[1]; cy++)
    for (int cz = 0; cz < nc[2]; cz++) {
        const int c = (cx*nc[1] + cy)*nc[2] + cz;
        neighbours.clear();
        for (int dx = -1; dx <= 1; dx++)
        for (int dy = -1; dy <= 1; dy++)
        for (int dz = -1; dz <= 1; dz++) {
            int n[3] = {cx+dx, cy+dy, cz+dz};
            bool ok = true;
            for (int d = 0; d < 3; d++) {
                if (usePeriodic)
                    n[d] = (n[d] + nc[d]) % nc[d];
                else if (n[d] < 0 || n[d] >= nc[d])
                    ok = false;
            }
            if (ok)
                neighbours.push_back((n[0]*nc[1] + n[1])*nc[2] + n[2]);
        }
        std::sort(neighbours.begin(), neighbours.end());
        neighbours.erase(std::unique(neighbours.begin(), neighbours.end()), neighbours.end());
        for (int a = cellStart[c]; a < cellStart[c+1]; a++) {
            const int i = order[a];
            const std::set<int>& excl = exclusions[i];
            for (int other : neighbours) {
                for (int b = cellStart[other]; b < cellStart[other+1]; b++) {
                    const int j = order[b];
                    if (j <= i)
                        continue;
                    double d2 = pairDistanceSquared(atomLocations[i], atomLocations[j], periodicBoxVectors, usePeriodic);
                    if (d2 > maxD2 || d2 < minD2)
                        continue;
                    if (excl.find(j) != excl.end())
                        continue;
                    neighborList.push_back(AtomPair(i, j));
                    if (reportSymmetricPairs)
                        neighborList.push_back(AtomPair(j, i));
                }
            }
        }
    }
}

} // namespace OpenMM

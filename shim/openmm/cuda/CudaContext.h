#ifndef OPENMM_CUDACONTEXT_H_
#define OPENMM_CUDACONTEXT_H_
#include "openmm/Vec3.h"
#include "openmm/cuda/CudaArray.h"
#include "openmm/cuda/CudaForceInfo.h"
#include <vector_types.h>
#include <vector_functions.h>
#include <vector>
namespace OpenMM {
/* Stand-in for OpenMM::CudaContext (OpenMM 7.x), reduced to the buffers and queries a force plugin sees (SURVEY.md
 * Appendix A; call sites platforms/cuda/src/CudaCoulKernels.cpp:58-61,523-620): positions + charge as real4 posq in the
 * platform's own atom order, the order map atomIndex[slot] = user index, the 64-bit fixed-point force buffer
 * [3][paddedNumAtoms] (value * 2^32), the energy buffer, the periodic box and the stream kernels are enqueued on.
 * The test harness (shim/cuda_harness.cpp) fills these exactly as the real platform lays them out. */
class OPENMM_EXPORT CudaContext {
public:
    static const int TileSize = 32;
    CudaContext(int numAtoms, int deviceIndex, bool useDoublePrecision, bool useMixedPrecision);
    ~CudaContext();
    void setAsCurrent();
    int getDeviceIndex() const { return deviceIndex; }
    bool getUseDoublePrecision() const { return useDouble; }
    bool getUseMixedPrecision() const { return useMixed; }
    int getNumAtoms() const { return numAtoms; }
    int getPaddedNumAtoms() const { return paddedNumAtoms; }
    int getNumAtomBlocks() const { return paddedNumAtoms/TileSize; }
    CudaArray& getPosq() { return posq; }                        /* real4 [paddedNumAtoms] */
    CudaArray& getPosqCorrection() { return posqCorrection; }    /* float4 [paddedNumAtoms], mixed precision only */
    CudaArray& getForce() { return force; }                      /* long long [3*paddedNumAtoms] */
    CudaArray& getEnergyBuffer() { return energyBuffer; }        /* mixed [energyBufferSize] */
    CudaArray& getAtomIndexArray() { return atomIndex; }         /* int [paddedNumAtoms] */
    double4 getPeriodicBoxSize() const { return make_double4(boxVectors[0][0], boxVectors[1][1], boxVectors[2][2], 0.0); }
    void getPeriodicBoxVectors(Vec3& a, Vec3& b, Vec3& c) const { a = boxVectors[0]; b = boxVectors[1]; c = boxVectors[2]; }
    void setPeriodicBoxVectors(const Vec3& a, const Vec3& b, const Vec3& c) { boxVectors[0] = a; boxVectors[1] = b; boxVectors[2] = c; }
    CUstream getCurrentStream() { return stream; }
    void addForce(CudaForceInfo* info) { forceInfos.push_back(info); }
    std::vector<CudaForceInfo*>& getForceInfos() { return forceInfos; }
private:
    int numAtoms, paddedNumAtoms, deviceIndex;
    bool useDouble, useMixed;
    CudaArray posq, posqCorrection, force, energyBuffer, atomIndex;
    Vec3 boxVectors[3];
    CUstream stream;
    std::vector<CudaForceInfo*> forceInfos;
};
} // namespace OpenMM
#endif

#ifndef B200_CUDA_COUL_KERNELS_H_
#define B200_CUDA_COUL_KERNELS_H_

/* The B200 implementation of CalcCoulForceKernel for OpenMM's CUDA platform: the replacement of
 * platforms/cuda/src/CudaCoulKernels.{h,cpp} (CudaCalcCoulForceKernel). Positions, forces and the energy never leave
 * the GPU: execute() hands the platform's own buffers -- real4 posq in the platform's atom order, the order map, the
 * fixed-point force buffer, the energy buffer -- and its stream to cfx_execute_platform() (include/cfx_b200.h). */
#include "CoulKernels.h"
#include "cfx_b200.h"
#include "openmm/cuda/CudaContext.h"
#include "openmm/cuda/CudaForceInfo.h"
#include <string>
#include <vector>

namespace CoulPlugin {

/* What the CUDA platform needs to know to reorder atoms. The reference's (CudaCoulKernels.cpp:20-47) lists only the
 * exclusions as particle groups and calls every group identical; here the charge-flux bonds, angles and waters are
 * groups too (a molecule with different flux parameters is not interchangeable with another), and two groups are
 * identical only if they are of the same kind with the same parameters. */
class B200CoulForceInfo : public OpenMM::CudaForceInfo {
public:
    B200CoulForceInfo(const CoulForce& force) : force(force) {}
    bool areParticlesIdentical(int particle1, int particle2);
    int getNumParticleGroups();
    void getParticlesInGroup(int index, std::vector<int>& particles);
    bool areGroupsIdentical(int group1, int group2);
private:
    /* kind: 0 exclusion, 1 flux bond, 2 flux angle, 3 flux water; local: index within the kind */
    void locate(int index, int& kind, int& local) const;
    const CoulForce& force;
};

class B200CudaCalcCoulForceKernel : public CalcCoulForceKernel {
public:
    B200CudaCalcCoulForceKernel(std::string name, const OpenMM::Platform& platform, OpenMM::CudaContext& cu) :
            CalcCoulForceKernel(name, platform), cu(cu), handle(NULL), numParticles(0), usePeriodic(false) {
    }
    ~B200CudaCalcCoulForceKernel();
    void initialize(const OpenMM::System& system, const CoulForce& force);
    /** Adds to the platform's force and energy buffers on the platform's stream; returns 0 like the reference's CUDA
     *  kernel (CudaCoulKernels.cpp:733): the platform sums its energy buffer itself. */
    double execute(OpenMM::ContextImpl& context, bool includeForces, bool includeEnergy);
private:
    OpenMM::CudaContext& cu;
    cfx_handle* handle;
    int numParticles;
    bool usePeriodic;
};

/* Fills a cfx_system_desc from the CoulForce through its public getters (shared by both adapters). The vectors own the
 * storage the descriptor points to. */
struct B200CoulDescriptor {
    cfx_system_desc desc;
    std::vector<double> charge, sigma, epsilon, bondPar, anglePar, waterPar;
    std::vector<int> excl, bondIdx, angleIdx, waterIdx;
    B200CoulDescriptor(const OpenMM::System& system, const CoulForce& force);
};

} // namespace CoulPlugin

#endif /*B200_CUDA_COUL_KERNELS_H_*/

#include "B200CudaCoulKernels.h"
#include "CoulForce.h"
#include "openmm/OpenMMException.h"
#include "openmm/System.h"
#include "openmm/internal/ContextImpl.h"

using namespace CoulPlugin;
using namespace OpenMM;
using namespace std;

B200CoulDescriptor::B200CoulDescriptor(const System& system, const CoulForce& force) {
    const int n = system.getNumParticles();
    if (force.getNumParticles() != n)
        throw OpenMMException("CoulForce must have exactly as many particles as the System it belongs to.");
    // Flatten the CoulForce storage through its public getters (openmmapi/src/CoulForce.cpp:28-136).
    charge.resize(n); sigma.resize(n); epsilon.resize(n);
    for (int i = 0; i < n; i++)
        force.getParticleParameters(i, charge[i], sigma[i], epsilon[i]);
    excl.resize(2*force.getNumExceptions());
    for (int i = 0; i < force.getNumExceptions(); i++)
        force.getExceptionParameters(i, excl[2*i], excl[2*i+1]);
    bondIdx.resize(2*force.getNumFluxBonds()); bondPar.resize(2*force.getNumFluxBonds());
    angleIdx.resize(3*force.getNumFluxAngles()); anglePar.resize(2*force.getNumFluxAngles());
    waterIdx.resize(3*force.getNumFluxWaters()); waterPar.resize(5*force.getNumFluxWaters());
    for (int i = 0; i < force.getNumFluxBonds(); i++)
        force.getFluxBondParameters(i, bondIdx[2*i], bondIdx[2*i+1], bondPar[2*i], bondPar[2*i+1]);
    for (int i = 0; i < force.getNumFluxAngles(); i++)
        force.getFluxAngleParameters(i, angleIdx[3*i], angleIdx[3*i+1], angleIdx[3*i+2], anglePar[2*i], anglePar[2*i+1]);
    for (int i = 0; i < force.getNumFluxWaters(); i++)
        force.getFluxWaterParameters(i, waterIdx[3*i], waterIdx[3*i+1], waterIdx[3*i+2], waterPar[5*i], waterPar[5*i+1],
                                     waterPar[5*i+2], waterPar[5*i+3], waterPar[5*i+4]);
    desc.num_particles = n;
    desc.charge = charge.data(); desc.sigma = sigma.data(); desc.epsilon = epsilon.data();
    desc.num_exceptions = force.getNumExceptions(); desc.exception_pairs = excl.data();
    desc.num_flux_bonds = force.getNumFluxBonds(); desc.flux_bond_idx = bondIdx.data(); desc.flux_bond_params = bondPar.data();
    desc.num_flux_angles = force.getNumFluxAngles(); desc.flux_angle_idx = angleIdx.data(); desc.flux_angle_params = anglePar.data();
    desc.num_flux_waters = force.getNumFluxWaters(); desc.flux_water_idx = waterIdx.data(); desc.flux_water_params = waterPar.data();
    desc.cutoff = force.getCutoffDistance();
    desc.ewald_tol = force.getEwaldErrorTolerance();
    desc.use_pbc = force.usesPeriodicBoundaryConditions() ? 1 : 0;
    Vec3 box[3];
    system.getDefaultPeriodicBoxVectors(box[0], box[1], box[2]);
    for (int a = 0; a < 3; a++)
        for (int c = 0; c < 3; c++)
            desc.default_box[3*a+c] = box[a][c];
}

/* ---- force info ---- */
bool B200CoulForceInfo::areParticlesIdentical(int particle1, int particle2) {
    double c1, c2, s1, s2, e1, e2;
    force.getParticleParameters(particle1, c1, s1, e1);
    force.getParticleParameters(particle2, c2, s2, e2);
    return (c1 == c2) && (s1 == s2) && (e1 == e2);
}

int B200CoulForceInfo::getNumParticleGroups() {
    return force.getNumExceptions() + force.getNumFluxBonds() + force.getNumFluxAngles() + force.getNumFluxWaters();
}

void B200CoulForceInfo::locate(int index, int& kind, int& local) const {
    const int counts[4] = {force.getNumExceptions(), force.getNumFluxBonds(), force.getNumFluxAngles(), force.getNumFluxWaters()};
    local = index;
    for (kind = 0; kind < 4; kind++) {
        if (local < counts[kind]) return;
        local -= counts[kind];
    }
    throw OpenMMException("B200CoulForceInfo: particle group index out of range");
}

void B200CoulForceInfo::getParticlesInGroup(int index, vector<int>& particles) {
    int kind, local;
    locate(index, kind, local);
    double p[5];
    if (kind == 0) { particles.resize(2); force.getExceptionParameters(local, particles[0], particles[1]); }
    else if (kind == 1) { particles.resize(2); force.getFluxBondParameters(local, particles[0], particles[1], p[0], p[1]); }
    else if (kind == 2) { particles.resize(3); force.getFluxAngleParameters(local, particles[0], particles[1], particles[2], p[0], p[1]); }
    else { particles.resize(3); force.getFluxWaterParameters(local, particles[0], particles[1], particles[2], p[0], p[1], p[2], p[3], p[4]); }
}

bool B200CoulForceInfo::areGroupsIdentical(int group1, int group2) {
    int k1, l1, k2, l2, a, b, c;
    locate(group1, k1, l1);
    locate(group2, k2, l2);
    if (k1 != k2) return false;
    if (k1 == 0) return true;                                      // an exclusion carries no parameters
    double p[5] = {0, 0, 0, 0, 0}, q[5] = {0, 0, 0, 0, 0};
    if (k1 == 1) { force.getFluxBondParameters(l1, a, b, p[0], p[1]); force.getFluxBondParameters(l2, a, b, q[0], q[1]); }
    else if (k1 == 2) { force.getFluxAngleParameters(l1, a, b, c, p[0], p[1]); force.getFluxAngleParameters(l2, a, b, c, q[0], q[1]); }
    else { force.getFluxWaterParameters(l1, a, b, c, p[0], p[1], p[2], p[3], p[4]); force.getFluxWaterParameters(l2, a, b, c, q[0], q[1], q[2], q[3], q[4]); }
    for (int k = 0; k < 5; k++)
        if (p[k] != q[k]) return false;
    return true;
}

/* ---- kernel ---- */
namespace {
void check(int status, const char* where) {
    if (status != CFX_OK)
        throw OpenMMException(string(where) + ": " + cfx_last_error());
}
}

B200CudaCalcCoulForceKernel::~B200CudaCalcCoulForceKernel() {
    if (handle != NULL) {
        cu.setAsCurrent();
        cfx_destroy(handle);
    }
}

void B200CudaCalcCoulForceKernel::initialize(const System& system, const CoulForce& force) {
    cu.setAsCurrent();                                             // CudaCoulKernels.cpp:58
    numParticles = system.getNumParticles();
    usePeriodic = force.usesPeriodicBoundaryConditions();
    B200CoulDescriptor d(system, force);
    cfx_options opts;
    opts.device = cu.getDeviceIndex(); opts.shard_rank = 0; opts.shard_count = 1; opts.use_graph = 1;
    // the value execute() would return for includeEnergy == false is discarded by OpenMM: do not compute it
    opts.flags = CFX_OPT_SKIP_DISCARDED_ENERGY;
    opts.list_skin_pm = 0;                       // library default (100 pm)
    for (int k = 0; k < 2; k++) opts.reserved[k] = 0;
    check(cfx_create(&d.desc, &opts, &handle), "B200CudaCalcCoulForceKernel::initialize");
    cu.addForce(new B200CoulForceInfo(force));                      // CudaCoulKernels.cpp:519
}

double B200CudaCalcCoulForceKernel::execute(ContextImpl& context, bool includeForces, bool includeEnergy) {
    cu.setAsCurrent();
    Vec3 a, b, c;
    cu.getPeriodicBoxVectors(a, b, c);
    const double box[9] = {a[0], a[1], a[2], b[0], b[1], b[2], c[0], c[1], c[2]};
    const bool dbl = cu.getUseDoublePrecision(), mixed = cu.getUseMixedPrecision();
    check(cfx_execute_platform(handle,
            reinterpret_cast<const void*>(cu.getPosq().getDevicePointer()), dbl ? 1 : 0,
            mixed ? reinterpret_cast<const void*>(cu.getPosqCorrection().getDevicePointer()) : NULL,
            reinterpret_cast<const int32_t*>(cu.getAtomIndexArray().getDevicePointer()), cu.getPaddedNumAtoms(),
            usePeriodic ? box : NULL, includeForces, includeEnergy,
            reinterpret_cast<unsigned long long*>(cu.getForce().getDevicePointer()),
            reinterpret_cast<void*>(cu.getEnergyBuffer().getDevicePointer()), (dbl || mixed) ? 1 : 0,
            reinterpret_cast<void*>(cu.getCurrentStream())),
          "B200CudaCalcCoulForceKernel::execute");
    return 0.0;                                                    // CudaCoulKernels.cpp:733
}

#include "B200CoulKernels.h"
#include "CoulForce.h"
#include "openmm/OpenMMException.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/reference/ReferencePlatform.h"

using namespace CoulPlugin;
using namespace OpenMM;
using namespace std;

namespace {
void check(int status, const char* where) {
    if (status != CFX_OK)
        throw OpenMMException(string(where) + ": " + cfx_last_error());
}
}

B200CalcCoulForceKernel::~B200CalcCoulForceKernel() {
    if (handle != NULL)
        cfx_destroy(handle);
}

void B200CalcCoulForceKernel::initialize(const System& system, const CoulForce& force) {
    numParticles = system.getNumParticles();
    if (force.getNumParticles() != numParticles)
        throw OpenMMException("CoulForce must have exactly as many particles as the System it belongs to.");
    // Flatten the CoulForce storage through its public getters (openmmapi/src/CoulForce.cpp:28-136).
    vector<double> charge(numParticles), sigma(numParticles), epsilon(numParticles);
    for (int i = 0; i < numParticles; i++)
        force.getParticleParameters(i, charge[i], sigma[i], epsilon[i]);
    vector<int> excl(2*force.getNumExceptions());
    for (int i = 0; i < force.getNumExceptions(); i++)
        force.getExceptionParameters(i, excl[2*i], excl[2*i+1]);
    vector<int> bondIdx(2*force.getNumFluxBonds()), angleIdx(3*force.getNumFluxAngles()), waterIdx(3*force.getNumFluxWaters());
    vector<double> bondPar(2*force.getNumFluxBonds()), anglePar(2*force.getNumFluxAngles()), waterPar(5*force.getNumFluxWaters());
    for (int i = 0; i < force.getNumFluxBonds(); i++)
        force.getFluxBondParameters(i, bondIdx[2*i], bondIdx[2*i+1], bondPar[2*i], bondPar[2*i+1]);
    for (int i = 0; i < force.getNumFluxAngles(); i++)
        force.getFluxAngleParameters(i, angleIdx[3*i], angleIdx[3*i+1], angleIdx[3*i+2], anglePar[2*i], anglePar[2*i+1]);
    for (int i = 0; i < force.getNumFluxWaters(); i++)
        force.getFluxWaterParameters(i, waterIdx[3*i], waterIdx[3*i+1], waterIdx[3*i+2], waterPar[5*i], waterPar[5*i+1],
                                     waterPar[5*i+2], waterPar[5*i+3], waterPar[5*i+4]);
    cfx_system_desc d;
    d.num_particles = numParticles;
    d.charge = charge.data(); d.sigma = sigma.data(); d.epsilon = epsilon.data();
    d.num_exceptions = force.getNumExceptions(); d.exception_pairs = excl.data();
    d.num_flux_bonds = force.getNumFluxBonds(); d.flux_bond_idx = bondIdx.data(); d.flux_bond_params = bondPar.data();
    d.num_flux_angles = force.getNumFluxAngles(); d.flux_angle_idx = angleIdx.data(); d.flux_angle_params = anglePar.data();
    d.num_flux_waters = force.getNumFluxWaters(); d.flux_water_idx = waterIdx.data(); d.flux_water_params = waterPar.data();
    d.cutoff = force.getCutoffDistance();
    d.ewald_tol = force.getEwaldErrorTolerance();
    d.use_pbc = force.usesPeriodicBoundaryConditions() ? 1 : 0;
    Vec3 box[3];
    system.getDefaultPeriodicBoxVectors(box[0], box[1], box[2]);
    for (int a = 0; a < 3; a++)
        for (int c = 0; c < 3; c++)
            d.default_box[3*a+c] = box[a][c];
    // The platform's position and force vectors live as long as the Context and are handed to every execute(): let the
    // library page-lock them in place instead of staging (include/cfx_b200.h, CFX_OPT_PIN_CALLER_BUFFERS).
    cfx_options opts;
    opts.device = -1; opts.shard_rank = 0; opts.shard_count = 1; opts.use_graph = 1;
    opts.flags = CFX_OPT_PIN_CALLER_BUFFERS;
    for (int k = 0; k < 3; k++) opts.reserved[k] = 0;
    check(cfx_create(&d, &opts, &handle), "B200CalcCoulForceKernel::initialize");
}

double B200CalcCoulForceKernel::execute(ContextImpl& context, bool includeForces, bool includeEnergy) {
    ReferencePlatform::PlatformData* data = reinterpret_cast<ReferencePlatform::PlatformData*>(context.getPlatformData());
    vector<Vec3>& pos = *((vector<Vec3>*) data->positions);
    vector<Vec3>& frc = *((vector<Vec3>*) data->forces);
    Vec3* boxVectors = (Vec3*) data->periodicBoxVectors;
    double box[9];
    for (int a = 0; a < 3; a++)
        for (int c = 0; c < 3; c++)
            box[3*a+c] = boxVectors[a][c];
    // vector<Vec3> is [N][3] doubles, the layout cfx_execute takes; it ADDS to the force array as the reference kernel
    // does (ReferenceCoulKernels.cpp:455-630), so the platform's own vectors are passed straight through.
    static_assert(sizeof(Vec3) == 3*sizeof(double), "Vec3 must be three packed doubles");
    if ((int) pos.size() < numParticles || (int) frc.size() < numParticles)
        throw OpenMMException("B200CalcCoulForceKernel::execute: platform vectors are smaller than the System");
    check(cfx_execute(handle, numParticles ? &pos[0][0] : NULL, box, includeForces, includeEnergy, lastEnergy, numParticles ? &frc[0][0] : NULL),
          "B200CalcCoulForceKernel::execute");
    return lastEnergy[CFX_E_TOTAL];
}

#include "B200CoulKernels.h"
#include "B200CudaCoulKernels.h"
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
    B200CoulDescriptor descriptor(system, force);
    cfx_system_desc& d = descriptor.desc;
    // The platform's position and force vectors live as long as the Context and are handed to every execute(): let the
    // library page-lock them in place instead of staging (include/cfx_b200.h, CFX_OPT_PIN_CALLER_BUFFERS).
    cfx_options opts;
    opts.device = -1; opts.shard_rank = 0; opts.shard_count = 1; opts.use_graph = 1;
    opts.flags = CFX_OPT_PIN_CALLER_BUFFERS;
    opts.list_skin_pm = 0;                       // library default (100 pm)
    for (int k = 0; k < 2; k++) opts.reserved[k] = 0;
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

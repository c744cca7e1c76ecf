/* cuda_harness.cpp -- drives a CalcCoulForce plugin kernel registered on the "CUDA" platform through the stand-in
 * CudaContext (shim/openmm/cuda/). TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * It plays the part of OpenMM's CUDA platform around one force: positions + charge are uploaded as real4 posq in a
 * SHUFFLED atom order (the real platform re-sorts atoms spatially), atomIndex maps platform slots to user indices, the
 * padded tail is zero, forces are read back from the 64-bit fixed-point buffer [3][paddedNumAtoms] and un-shuffled, the
 * energy is read from element 0 of the energy buffer. The unmodified CoulForce / CoulForceImpl of the plugin
 * (shim/_build/libOpenMMCoul.so) sit between this harness and the kernel, exactly as in a real Context.
 */
#include "openmm/Platform.h"
#include "openmm/System.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/cuda/CudaPlatform.h"
#include "CoulForce.h"
#include "../include/cfx_b200.h"

#include <cuda_runtime.h>
#include <algorithm>
#include <cmath>
#include <random>
#include <set>
#include <string>
#include <vector>

using namespace OpenMM;
using namespace CoulPlugin;

namespace {
thread_local std::string g_err;
std::set<std::string>& loadedPlugins() { static std::set<std::string> s; return s; }
void ensurePlatform() {
    static bool done = false;
    if (done) return;
    Platform::registerPlatform(new CudaPlatform());
    done = true;
}
}

struct cfxcu_handle {
    OpenMM::System system;
    CudaPlatform::PlatformData* data = nullptr;
    CudaContext* cu = nullptr;             // owned by data
    ContextImpl* context = nullptr;
    int n = 0, mode = 0;                   // 0 single, 1 mixed, 2 double
    std::vector<int> order;                // platform slot -> user index
    ~cfxcu_handle() { delete context; delete data; }
};

extern "C" {

const char* cfxcu_last_error(void) { return g_err.c_str(); }

int cfxcu_load_plugin(const char* path) {
    try {
        ensurePlatform();
        if (loadedPlugins().insert(path).second)
            Platform::loadPluginLibrary(path);
        return CFX_OK;
    } catch (std::exception& e) { g_err = e.what(); return CFX_ERR_ARGUMENT; }
}

/* precision: 0 single (float4 posq, float energy), 1 mixed (float4 posq + correction, double energy), 2 double */
int cfxcu_create(const cfx_system_desc* d, int precision, int device, unsigned seed, cfxcu_handle** out) {
    try {
        ensurePlatform();
        cfxcu_handle* h = new cfxcu_handle();
        h->n = d->num_particles; h->mode = precision;
        CoulForce* f = new CoulForce();
        for (int i = 0; i < d->num_particles; i++) {
            h->system.addParticle(1.0);
            f->addParticle(d->charge[i], d->sigma[i], d->epsilon[i]);
        }
        for (int i = 0; i < d->num_exceptions; i++)
            f->addException(d->exception_pairs[2*i], d->exception_pairs[2*i+1]);
        for (int i = 0; i < d->num_flux_bonds; i++)
            f->addFluxBond(d->flux_bond_idx[2*i], d->flux_bond_idx[2*i+1], d->flux_bond_params[2*i], d->flux_bond_params[2*i+1]);
        for (int i = 0; i < d->num_flux_angles; i++)
            f->addFluxAngle(d->flux_angle_idx[3*i], d->flux_angle_idx[3*i+1], d->flux_angle_idx[3*i+2],
                            d->flux_angle_params[2*i], d->flux_angle_params[2*i+1]);
        for (int i = 0; i < d->num_flux_waters; i++) {
            const double* p = d->flux_water_params + 5*i;
            f->addFluxWater(d->flux_water_idx[3*i], d->flux_water_idx[3*i+1], d->flux_water_idx[3*i+2], p[0], p[1], p[2], p[3], p[4]);
        }
        f->setCutoffDistance(d->cutoff);
        f->setEwaldErrorTolerance(d->ewald_tol);
        f->setUsesPeriodicBoundaryConditions(d->use_pbc != 0);
        h->system.addForce(f);
        const double* b = d->default_box;
        h->system.setDefaultPeriodicBoxVectors(Vec3(b[0],b[1],b[2]), Vec3(b[3],b[4],b[5]), Vec3(b[6],b[7],b[8]));
        h->data = new CudaPlatform::PlatformData();
        h->cu = new CudaContext(h->n, device, precision == 2, precision == 1);
        h->data->contexts.push_back(h->cu);
        // the platform's atom order: a seeded shuffle (the real platform sorts spatially and re-sorts as atoms move)
        h->order.resize(h->n);
        for (int i = 0; i < h->n; i++) h->order[i] = i;
        std::mt19937 rng(seed);
        std::shuffle(h->order.begin(), h->order.end(), rng);
        std::vector<int> padded(h->cu->getPaddedNumAtoms(), 0);
        std::copy(h->order.begin(), h->order.end(), padded.begin());
        h->cu->getAtomIndexArray().upload(padded);
        h->context = new ContextImpl(h->system, Platform::getPlatformByName("CUDA"), h->data);
        *out = h;
        return CFX_OK;
    } catch (std::exception& e) { g_err = e.what(); return CFX_ERR_ARGUMENT; }
}

void cfxcu_destroy(cfxcu_handle* h) { delete h; }

/* energy[CFX_E_TOTAL] = value returned by execute + element 0 of the platform's energy buffer; forces ADDED to. */
int cfxcu_execute(cfxcu_handle* h, const double* positions, const double* box, int includeForces, int includeEnergy,
                  double* energy, double* forces) {
    try {
        CudaContext& cu = *h->cu;
        cu.setAsCurrent();
        const int np = cu.getPaddedNumAtoms();
        if (h->mode == 2) {
            std::vector<double4> posq(np, make_double4(0, 0, 0, 0));
            for (int s = 0; s < h->n; s++) { const double* p = positions + 3*h->order[s]; posq[s] = make_double4(p[0], p[1], p[2], 0.0); }
            cu.getPosq().upload(posq);
        }
        else {
            std::vector<float4> posq(np, make_float4(0, 0, 0, 0)), corr(np, make_float4(0, 0, 0, 0));
            for (int s = 0; s < h->n; s++) {
                const double* p = positions + 3*h->order[s];
                posq[s] = make_float4((float) p[0], (float) p[1], (float) p[2], 0.f);
                corr[s] = make_float4((float) (p[0] - (double) posq[s].x), (float) (p[1] - (double) posq[s].y), (float) (p[2] - (double) posq[s].z), 0.f);
            }
            cu.getPosq().upload(posq);
            if (h->mode == 1) cu.getPosqCorrection().upload(corr);
        }
        cu.setPeriodicBoxVectors(Vec3(box[0],box[1],box[2]), Vec3(box[3],box[4],box[5]), Vec3(box[6],box[7],box[8]));
        std::vector<long long> zero(3*np, 0);
        cu.getForce().upload(zero);
        if (h->mode == 0) { std::vector<float> e0(cu.getEnergyBuffer().getSize(), 0.f); cu.getEnergyBuffer().upload(e0); }
        else { std::vector<double> e0(cu.getEnergyBuffer().getSize(), 0.0); cu.getEnergyBuffer().upload(e0); }
        double e = h->context->calcForcesAndEnergy(includeForces != 0, includeEnergy != 0);
        if (cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(cu.getCurrentStream())) != cudaSuccess)
            throw OpenMMException("stream synchronisation failed");
        if (h->mode == 0) { std::vector<float> eb; cu.getEnergyBuffer().download(eb); e += eb[0]; }
        else { std::vector<double> eb; cu.getEnergyBuffer().download(eb); e += eb[0]; }
        if (energy) {
            for (int k = 0; k < CFX_E_COUNT; k++) energy[k] = NAN;
            energy[CFX_E_TOTAL] = e;
        }
        if (forces) {
            std::vector<long long> fb;
            cu.getForce().download(fb);
            for (int s = 0; s < h->n; s++)
                for (int c = 0; c < 3; c++)
                    forces[3*h->order[s] + c] += (double) fb[(size_t) c*np + s]/4294967296.0;
        }
        return CFX_OK;
    } catch (std::exception& e) { g_err = e.what(); return CFX_ERR_ARGUMENT; }
}

/* The CudaForceInfo the kernel registered: group count, one group's particles, and the two identity predicates. */
int cfxcu_force_info_groups(cfxcu_handle* h) {
    if (h->cu->getForceInfos().empty()) return -1;
    return h->cu->getForceInfos()[0]->getNumParticleGroups();
}
int cfxcu_force_info_group(cfxcu_handle* h, int index, int* particles, int capacity) {
    std::vector<int> p;
    h->cu->getForceInfos()[0]->getParticlesInGroup(index, p);
    for (int k = 0; k < (int) p.size() && k < capacity; k++) particles[k] = p[k];
    return (int) p.size();
}
int cfxcu_force_info_particles_identical(cfxcu_handle* h, int a, int b) { return h->cu->getForceInfos()[0]->areParticlesIdentical(a, b) ? 1 : 0; }
int cfxcu_force_info_groups_identical(cfxcu_handle* h, int a, int b) { return h->cu->getForceInfos()[0]->areGroupsIdentical(a, b) ? 1 : 0; }

} // extern "C"

/* ref_harness.cpp -- C API around the plugin's OWN sources. TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Drives the unmodified reference code exactly the way OpenMM would:
 *   CoulForce setters (openmmapi/src/CoulForce.cpp) -> System::addForce -> ContextImpl ->
 *   CoulForce::createImpl -> CoulForceImpl::initialize -> Platform::createKernel("CalcCoulForce")
 *   -> <platform plugin>::initialize / execute.
 * The platform plugin is loaded with the same dlopen + registerPlatforms()/registerKernelFactories()
 * contract OpenMM's loader uses, so the same harness runs either the reference's
 * platforms/reference plugin (libOpenMMCoulReference.so, built from /root/reference) or this
 * repository's B200 plugin (libOpenMMCoulB200.so) behind one CoulForce object.
 *
 * OpenMM itself is replaced by the stand-in in shim/ (see shim/README.md for what that restates).
 * Built by oracle/Makefile into oracle/_ref/libcfx_ref.so; never shipped, never linked by the product.
 */
#include <algorithm>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>
#include <iostream>
#include <cmath>

#include "openmm/Platform.h"
#include "openmm/System.h"
#include "openmm/internal/ContextImpl.h"
#include "openmm/reference/ReferencePlatform.h"
#include "openmm/reference/ReferenceNeighborList.h"
#include "openmm/serialization/XmlSerializer.h"
#include <sstream>

// The parity getters read the reference kernel's private state (charges, Jacobian rows, neighbour
// list, kmax). Access control does not change object layout; nothing in the reference is edited.
#define private public
#include "CoulForce.h"
#include "internal/CoulForceImpl.h"
#include "ReferenceCoulKernels.h"
#undef private

#include "../include/cfx_b200.h"

using namespace OpenMM;
using namespace CoulPlugin;

namespace {

thread_local std::string g_err;

/* A second host-memory platform for the B200 plugin to register on. It exposes the same
 * ReferencePlatform::PlatformData (positions / forces / box on the host), i.e. what OpenMM's
 * Reference and CPU platforms hand to a plugin kernel. */
class HostB200Platform : public ReferencePlatform {
public:
    const std::string& getName() const { static const std::string name = "B200"; return name; }
};

std::set<std::string>& loadedPlugins() { static std::set<std::string> s; return s; }

void ensurePlatforms() {
    static bool done = false;
    if (done) return;
    Platform::registerPlatform(new ReferencePlatform());
    Platform::registerPlatform(new HostB200Platform());
    done = true;
}

} // namespace

struct cfxref_handle {
    OpenMM::System system;
    CoulForce* force = nullptr;          // owned by system
    ReferencePlatform::PlatformData* data = nullptr;
    ContextImpl* context = nullptr;
    int n = 0;
    bool isReferenceKernel = false;
    ~cfxref_handle() { delete context; delete data; }
    ReferenceCalcCoulForceKernel* refKernel() {
        if (!isReferenceKernel) return nullptr;
        CoulForceImpl* impl = dynamic_cast<CoulForceImpl*>(context->getForceImpls()[0]);
        return dynamic_cast<ReferenceCalcCoulForceKernel*>(&impl->kernel.getImpl());
    }
};

extern "C" {

const char* cfxref_last_error(void) { return g_err.c_str(); }

/* Load a platform plugin library (once per path). */
int cfxref_load_plugin(const char* path) {
    try {
        ensurePlatforms();
        if (loadedPlugins().insert(path).second)
            Platform::loadPluginLibrary(path);
        return CFX_OK;
    } catch (std::exception& e) { g_err = e.what(); return CFX_ERR_ARGUMENT; }
}

/* platform: "Reference" (the plugin's own CPU kernel) or "B200" (this repository's plugin). */
int cfxref_create(const cfx_system_desc* d, const char* platform, cfxref_handle** out) {
    try {
        ensurePlatforms();
        cfxref_handle* h = new cfxref_handle();
        h->n = d->num_particles;
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
        h->force = f;
        h->system.addForce(f);
        const double* b = d->default_box;
        h->system.setDefaultPeriodicBoxVectors(Vec3(b[0],b[1],b[2]), Vec3(b[3],b[4],b[5]), Vec3(b[6],b[7],b[8]));
        h->data = new ReferencePlatform::PlatformData(d->num_particles);
        Platform& p = Platform::getPlatformByName(platform);
        h->isReferenceKernel = (std::string(platform) == "Reference");
        h->context = new ContextImpl(h->system, p, h->data);
        *out = h;
        return CFX_OK;
    } catch (std::exception& e) { g_err = e.what(); return CFX_ERR_ARGUMENT; }
}

void cfxref_destroy(cfxref_handle* h) { delete h; }
int cfxref_num_particles(cfxref_handle* h) { return h->n; }

/* XML text of the handle's CoulForce through XmlSerializer (needs a library that registered a proxy for CoulForce: the
 * B200 plugin does at load; the reference registers none). Two-call protocol: out == NULL returns the size. */
int cfxref_serialize_xml(cfxref_handle* h, char* out, int64_t capacity, int64_t* needed) {
    try {
        std::ostringstream text;
        XmlSerializer::serialize<Force>(h->force, "Force", text);
        const std::string s = text.str();
        *needed = (int64_t) s.size() + 1;
        if (!out) return CFX_OK;
        if (capacity < *needed) { g_err = "xml buffer too small"; return CFX_ERR_ARGUMENT; }
        memcpy(out, s.c_str(), s.size() + 1);
        return CFX_OK;
    } catch (std::exception& e) { g_err = e.what(); return CFX_ERR_ARGUMENT; }
}

/* A context whose CoulForce comes out of XmlSerializer::deserialize (the proxy's add* calls). */
int cfxref_create_from_xml(const char* xml, const double* default_box, const char* platform, cfxref_handle** out) {
    try {
        ensurePlatforms();
        std::istringstream text(xml);
        Force* obj = XmlSerializer::deserialize<Force>(text);
        CoulForce* f = dynamic_cast<CoulForce*>(obj);
        if (!f) { delete obj; g_err = "the XML does not hold a CoulForce"; return CFX_ERR_ARGUMENT; }
        cfxref_handle* h = new cfxref_handle();
        h->n = f->getNumParticles();
        for (int i = 0; i < h->n; i++) h->system.addParticle(1.0);
        h->force = f;
        h->system.addForce(f);
        const double* b = default_box;
        h->system.setDefaultPeriodicBoxVectors(Vec3(b[0],b[1],b[2]), Vec3(b[3],b[4],b[5]), Vec3(b[6],b[7],b[8]));
        h->data = new ReferencePlatform::PlatformData(h->n);
        Platform& p = Platform::getPlatformByName(platform);
        h->isReferenceKernel = (std::string(platform) == "Reference");
        h->context = new ContextImpl(h->system, p, h->data);
        *out = h;
        return CFX_OK;
    } catch (std::exception& e) { g_err = e.what(); return CFX_ERR_ARGUMENT; }
}

/* energy: only the total is known to the caller of CalcCoulForceKernel::execute -> written to
 * energy[CFX_E_TOTAL]; the component slots are set to NaN. forces (may be NULL) is ADDED to. */
int cfxref_execute(cfxref_handle* h, const double* positions, const double* box, int includeForces, int includeEnergy,
                   double* energy, double* forces) {
    try {
        std::vector<Vec3>& pos = *(std::vector<Vec3>*) h->data->positions;
        std::vector<Vec3>& frc = *(std::vector<Vec3>*) h->data->forces;
        Vec3* bv = (Vec3*) h->data->periodicBoxVectors;
        for (int i = 0; i < h->n; i++) {
            pos[i] = Vec3(positions[3*i], positions[3*i+1], positions[3*i+2]);
            frc[i] = forces ? Vec3(forces[3*i], forces[3*i+1], forces[3*i+2]) : Vec3();
        }
        for (int a = 0; a < 3; a++)
            bv[a] = Vec3(box[3*a], box[3*a+1], box[3*a+2]);
        double e = h->context->calcForcesAndEnergy(includeForces != 0, includeEnergy != 0);
        if (energy) {
            for (int k = 0; k < CFX_E_COUNT; k++) energy[k] = NAN;
            energy[CFX_E_TOTAL] = e;
        }
        if (forces)
            for (int i = 0; i < h->n; i++)
                for (int c = 0; c < 3; c++)
                    forces[3*i+c] = frc[i][c];
        return CFX_OK;
    } catch (std::exception& e) { g_err = e.what(); return CFX_ERR_ARGUMENT; }
}

int cfxref_get_ewald_params(cfxref_handle* h, cfx_ewald_params* out) {
    ReferenceCalcCoulForceKernel* k = h->refKernel();
    if (!k) { g_err = "not a Reference-platform kernel"; return CFX_ERR_STATE; }
    out->alpha = k->alpha;
    out->kmax[0] = k->kmaxx; out->kmax[1] = k->kmaxy; out->kmax[2] = k->kmaxz;
    long long kx = k->kmaxx, ky = k->kmaxy, kz = k->kmaxz;
    out->num_kvectors = (kz - 1) + (ky - 1)*(2*kz - 1) + (kx - 1)*(2*ky - 1)*(2*kz - 1);
    return CFX_OK;
}

int cfxref_get_charges(cfxref_handle* h, double* q) {
    ReferenceCalcCoulForceKernel* k = h->refKernel();
    if (!k) { g_err = "not a Reference-platform kernel"; return CFX_ERR_STATE; }
    memcpy(q, k->realcharges.data(), sizeof(double)*h->n);
    return CFX_OK;
}

int cfxref_num_jacobian_rows(cfxref_handle* h) {
    ReferenceCalcCoulForceKernel* k = h->refKernel();
    return k ? (int) k->dqdx_dqidx.size() : -1;
}

int cfxref_get_jacobian(cfxref_handle* h, int32_t* dq, int32_t* dx, double* val) {
    ReferenceCalcCoulForceKernel* k = h->refKernel();
    if (!k) { g_err = "not a Reference-platform kernel"; return CFX_ERR_STATE; }
    if (dq) memcpy(dq, k->dqdx_dqidx.data(), sizeof(int)*k->dqdx_dqidx.size());
    if (dx) memcpy(dx, k->dqdx_dxidx.data(), sizeof(int)*k->dqdx_dxidx.size());
    if (val) memcpy(val, k->dqdx_val.data(), sizeof(double)*k->dqdx_val.size());
    return CFX_OK;
}

int cfxref_get_neighbor_pairs(cfxref_handle* h, int32_t* pairs, int64_t capacity, int64_t* count) {
    ReferenceCalcCoulForceKernel* k = h->refKernel();
    if (!k || !k->ifPBC) { g_err = "no neighbour list (not a periodic Reference-platform kernel)"; return CFX_ERR_STATE; }
    NeighborList& nl = *k->neighborList;
    *count = (int64_t) nl.size();
    if (!pairs) return CFX_OK;
    if (capacity < *count) { g_err = "pair buffer too small"; return CFX_ERR_ARGUMENT; }
    std::vector<std::pair<int,int> > sorted;
    for (auto& p : nl) sorted.push_back(std::make_pair(std::min(p.first, p.second), std::max(p.first, p.second)));
    std::sort(sorted.begin(), sorted.end());
    for (size_t i = 0; i < sorted.size(); i++) { pairs[2*i] = sorted[i].first; pairs[2*i+1] = sorted[i].second; }
    return CFX_OK;
}

} // extern "C"

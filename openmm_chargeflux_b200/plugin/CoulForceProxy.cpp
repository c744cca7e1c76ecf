/* CoulForceProxy.cpp -- XML form of a CoulForce.
 *
 *   <Force type="CoulForce" version="1" cutoff=".." ewaldTolerance=".." usesPeriodic="0|1" forceGroup="..">
 *     <Particles>  <Particle q=".." sig=".." eps=".."/> ...                              </Particles>
 *     <Exceptions> <Exception p1=".." p2=".."/> ...                                       </Exceptions>
 *     <FluxBonds>  <Bond p1=".." p2=".." k=".." b=".."/> ...                             </FluxBonds>
 *     <FluxAngles> <Angle p1=".." p2=".." p3=".." k=".." theta=".."/> ...               </FluxAngles>
 *     <FluxWaters> <Water po=".." ph1=".." ph2=".." k1=".." k2=".." kub=".." b0=".." ub0=".."/> ... </FluxWaters>
 *   </Force>
 *
 * Element order is the order of the add* calls, which fixes the Jacobian row order the kernels build
 * (ReferenceCoulKernels.cpp:286-383): a deserialized force evaluates bit-identically. openmm_chargeflux_b200/force.py
 * writes and reads the same layout. */
#include "CoulForceProxy.h"
#include "CoulForce.h"
#include "openmm/OpenMMException.h"
#include "openmm/serialization/SerializationNode.h"

using namespace OpenMM;

namespace CoulPlugin {

CoulForceProxy::CoulForceProxy() : SerializationProxy("CoulForce") {}

void CoulForceProxy::serialize(const void* object, SerializationNode& node) const {
    const CoulForce& force = *reinterpret_cast<const CoulForce*>(object);
    node.setIntProperty("version", 1);
    node.setIntProperty("forceGroup", force.getForceGroup());
    node.setDoubleProperty("cutoff", force.getCutoffDistance());
    node.setDoubleProperty("ewaldTolerance", force.getEwaldErrorTolerance());
    node.setBoolProperty("usesPeriodic", force.usesPeriodicBoundaryConditions());
    SerializationNode& particles = node.createChildNode("Particles");
    for (int i = 0; i < force.getNumParticles(); i++) {
        double q, sig, eps;
        force.getParticleParameters(i, q, sig, eps);
        particles.createChildNode("Particle").setDoubleProperty("q", q).setDoubleProperty("sig", sig).setDoubleProperty("eps", eps);
    }
    SerializationNode& exceptions = node.createChildNode("Exceptions");
    for (int i = 0; i < force.getNumExceptions(); i++) {
        int p1, p2;
        force.getExceptionParameters(i, p1, p2);
        exceptions.createChildNode("Exception").setIntProperty("p1", p1).setIntProperty("p2", p2);
    }
    SerializationNode& bonds = node.createChildNode("FluxBonds");
    for (int i = 0; i < force.getNumFluxBonds(); i++) {
        int p1, p2; double k, b;
        force.getFluxBondParameters(i, p1, p2, k, b);
        bonds.createChildNode("Bond").setIntProperty("p1", p1).setIntProperty("p2", p2).setDoubleProperty("k", k).setDoubleProperty("b", b);
    }
    SerializationNode& angles = node.createChildNode("FluxAngles");
    for (int i = 0; i < force.getNumFluxAngles(); i++) {
        int p1, p2, p3; double k, theta;
        force.getFluxAngleParameters(i, p1, p2, p3, k, theta);
        angles.createChildNode("Angle").setIntProperty("p1", p1).setIntProperty("p2", p2).setIntProperty("p3", p3)
              .setDoubleProperty("k", k).setDoubleProperty("theta", theta);
    }
    SerializationNode& waters = node.createChildNode("FluxWaters");
    for (int i = 0; i < force.getNumFluxWaters(); i++) {
        int po, ph1, ph2; double k1, k2, kub, b0, ub0;
        force.getFluxWaterParameters(i, po, ph1, ph2, k1, k2, kub, b0, ub0);
        waters.createChildNode("Water").setIntProperty("po", po).setIntProperty("ph1", ph1).setIntProperty("ph2", ph2)
              .setDoubleProperty("k1", k1).setDoubleProperty("k2", k2).setDoubleProperty("kub", kub)
              .setDoubleProperty("b0", b0).setDoubleProperty("ub0", ub0);
    }
}

void* CoulForceProxy::deserialize(const SerializationNode& node) const {
    if (node.getIntProperty("version") != 1)
        throw OpenMMException("Unsupported version number");
    CoulForce* force = new CoulForce();
    try {
        force->setForceGroup(node.getIntProperty("forceGroup", 0));
        force->setCutoffDistance(node.getDoubleProperty("cutoff"));
        force->setEwaldErrorTolerance(node.getDoubleProperty("ewaldTolerance"));
        force->setUsesPeriodicBoundaryConditions(node.getBoolProperty("usesPeriodic"));
        for (const SerializationNode& p : node.getChildNode("Particles").getChildren())
            force->addParticle(p.getDoubleProperty("q"), p.getDoubleProperty("sig"), p.getDoubleProperty("eps"));
        for (const SerializationNode& e : node.getChildNode("Exceptions").getChildren())
            force->addException(e.getIntProperty("p1"), e.getIntProperty("p2"));
        for (const SerializationNode& b : node.getChildNode("FluxBonds").getChildren())
            force->addFluxBond(b.getIntProperty("p1"), b.getIntProperty("p2"), b.getDoubleProperty("k"), b.getDoubleProperty("b"));
        for (const SerializationNode& a : node.getChildNode("FluxAngles").getChildren())
            force->addFluxAngle(a.getIntProperty("p1"), a.getIntProperty("p2"), a.getIntProperty("p3"),
                                a.getDoubleProperty("k"), a.getDoubleProperty("theta"));
        for (const SerializationNode& w : node.getChildNode("FluxWaters").getChildren())
            force->addFluxWater(w.getIntProperty("po"), w.getIntProperty("ph1"), w.getIntProperty("ph2"),
                                w.getDoubleProperty("k1"), w.getDoubleProperty("k2"), w.getDoubleProperty("kub"),
                                w.getDoubleProperty("b0"), w.getDoubleProperty("ub0"));
    }
    catch (...) {
        delete force;
        throw;
    }
    return force;
}

} // namespace CoulPlugin

extern "C" void registerCoulSerializationProxies() {
    static bool done = false;
    if (done) return;
    done = true;
    SerializationProxy::registerProxy(typeid(CoulPlugin::CoulForce), new CoulPlugin::CoulForceProxy());   // owned by the registry
}

namespace {
struct RegisterAtLoad { RegisterAtLoad() { registerCoulSerializationProxies(); } } registerAtLoad;
}

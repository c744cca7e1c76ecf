#ifndef B200_COULFORCE_PROXY_H_
#define B200_COULFORCE_PROXY_H_
/* Serialization proxy for CoulPlugin::CoulForce. The reference ships none (no serialization/
 * directory, nothing registered: XmlSerializer cannot write a System that holds a CoulForce), so this is the piece an
 * OpenMM plugin normally carries as <Plugin>ForceProxy, written against the public CoulForce API only
 * (/root/reference/openmmapi/include/CoulForce.h:27-133): the API class stays untouched. */
#include "openmm/serialization/SerializationProxy.h"

namespace CoulPlugin {

class CoulForceProxy : public OpenMM::SerializationProxy {
public:
    CoulForceProxy();
    void serialize(const void* object, OpenMM::SerializationNode& node) const;
    void* deserialize(const OpenMM::SerializationNode& node) const;
};

} // namespace CoulPlugin

/* Registers the proxy with OpenMM's registry; also runs when the library is loaded. */
extern "C" void registerCoulSerializationProxies();

#endif

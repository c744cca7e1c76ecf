#ifndef OPENMM_SERIALIZATIONPROXY_H_
#define OPENMM_SERIALIZATIONPROXY_H_
/* Stand-in for openmm/serialization/SerializationProxy.h: the registry that maps a C++ type / a type name to the object
 * that knows how to write and read it. TEST INFRASTRUCTURE: see shim/README.md. */
#include <string>
#include <typeinfo>
#include "../internal/windowsExport.h"

namespace OpenMM {

class SerializationNode;

class OPENMM_EXPORT SerializationProxy {
public:
    SerializationProxy(const std::string& typeName) : typeName(typeName) {}
    virtual ~SerializationProxy() {}
    const std::string& getTypeName() const { return typeName; }
    virtual void serialize(const void* object, SerializationNode& node) const = 0;
    virtual void* deserialize(const SerializationNode& node) const = 0;
    static void registerProxy(const std::type_info& type, const SerializationProxy* proxy);
    static const SerializationProxy& getProxy(const std::string& typeName);
    static const SerializationProxy& getProxy(const std::type_info& type);
private:
    std::string typeName;
};

} // namespace OpenMM
#endif

#ifndef OPENMM_XMLSERIALIZER_H_
#define OPENMM_XMLSERIALIZER_H_
/* Stand-in for openmm/serialization/XmlSerializer.h: object -> proxy -> node tree -> XML text and back. The XML subset is
 * what OpenMM writes for a Force: elements with attributes and child elements, no text nodes. TEST INFRASTRUCTURE. */
#include <iosfwd>
#include <string>
#include <typeinfo>
#include "SerializationNode.h"
#include "SerializationProxy.h"

namespace OpenMM {

class OPENMM_EXPORT XmlSerializer {
public:
    template <class T> static void serialize(const T* object, const std::string& rootName, std::ostream& stream) {
        const SerializationProxy& proxy = SerializationProxy::getProxy(typeid(*object));
        SerializationNode node;
        node.setName(rootName);
        proxy.serialize(object, node);
        if (!node.hasProperty("type")) node.setStringProperty("type", proxy.getTypeName());
        encode(node, stream);
    }
    template <class T> static T* deserialize(std::istream& stream) { return reinterpret_cast<T*>(deserializeStream(stream)); }
    static void encode(const SerializationNode& node, std::ostream& stream);
    static void decode(std::istream& stream, SerializationNode& node);
private:
    static void* deserializeStream(std::istream& stream);
};

} // namespace OpenMM
#endif

/* shim_serialization.cpp -- behaviour of the serialization stand-ins (SerializationNode, SerializationProxy registry,
 * a small XML writer / reader). TEST INFRASTRUCTURE, restated from OpenMM's documented interface; see shim/README.md. */
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <istream>
#include <iterator>
#include <map>
#include <ostream>
#include <sstream>
#include "SerializationNode.h"
#include "SerializationProxy.h"
#include "XmlSerializer.h"
#include "../OpenMMException.h"

namespace OpenMM {

const SerializationNode& SerializationNode::getChildNode(const std::string& n) const {
    for (const SerializationNode& c : children) if (c.name == n) return c;
    throw OpenMMException("Unknown child '" + n + "' in node '" + name + "'");
}
SerializationNode& SerializationNode::getChildNode(const std::string& n) {
    return const_cast<SerializationNode&>(static_cast<const SerializationNode&>(*this).getChildNode(n));
}
SerializationNode& SerializationNode::createChildNode(const std::string& n) {
    children.push_back(SerializationNode());
    children.back().setName(n);
    return children.back();
}
const std::string& SerializationNode::getStringProperty(const std::string& n) const {
    auto it = properties.find(n);
    if (it == properties.end()) throw OpenMMException("Unknown property '" + n + "' in node '" + name + "'");
    return it->second;
}
const std::string& SerializationNode::getStringProperty(const std::string& n, const std::string& def) const {
    auto it = properties.find(n);
    return it == properties.end() ? def : it->second;
}
SerializationNode& SerializationNode::setStringProperty(const std::string& n, const std::string& v) { properties[n] = v; return *this; }
int SerializationNode::getIntProperty(const std::string& n) const { return atoi(getStringProperty(n).c_str()); }
int SerializationNode::getIntProperty(const std::string& n, int def) const { return hasProperty(n) ? getIntProperty(n) : def; }
SerializationNode& SerializationNode::setIntProperty(const std::string& n, int v) { return setStringProperty(n, std::to_string(v)); }
bool SerializationNode::getBoolProperty(const std::string& n) const { return getIntProperty(n) != 0; }
bool SerializationNode::getBoolProperty(const std::string& n, bool def) const { return hasProperty(n) ? getBoolProperty(n) : def; }
SerializationNode& SerializationNode::setBoolProperty(const std::string& n, bool v) { return setStringProperty(n, v ? "1" : "0"); }
double SerializationNode::getDoubleProperty(const std::string& n) const { return strtod(getStringProperty(n).c_str(), nullptr); }
double SerializationNode::getDoubleProperty(const std::string& n, double def) const { return hasProperty(n) ? getDoubleProperty(n) : def; }
SerializationNode& SerializationNode::setDoubleProperty(const std::string& n, double v) {
    char buf[40];
    snprintf(buf, sizeof(buf), "%.17g", v);            // round-trips every double
    return setStringProperty(n, buf);
}

namespace {
std::map<std::string, const SerializationProxy*>& byName() { static std::map<std::string, const SerializationProxy*> m; return m; }
std::map<std::string, const SerializationProxy*>& byType() { static std::map<std::string, const SerializationProxy*> m; return m; }
}
void SerializationProxy::registerProxy(const std::type_info& type, const SerializationProxy* proxy) {
    byType()[type.name()] = proxy;
    byName()[proxy->getTypeName()] = proxy;
}
const SerializationProxy& SerializationProxy::getProxy(const std::string& typeName) {
    auto it = byName().find(typeName);
    if (it == byName().end()) throw OpenMMException("There is no serialization proxy registered for type " + typeName);
    return *it->second;
}
const SerializationProxy& SerializationProxy::getProxy(const std::type_info& type) {
    auto it = byType().find(type.name());
    if (it == byType().end()) throw OpenMMException(std::string("There is no serialization proxy registered for type ") + type.name());
    return *it->second;
}

namespace {
std::string escape(const std::string& s) {
    std::string o;
    for (char c : s) {
        if (c == '&') o += "&amp;"; else if (c == '<') o += "&lt;"; else if (c == '>') o += "&gt;"; else if (c == '"') o += "&quot;"; else o += c;
    }
    return o;
}
std::string unescape(const std::string& s) {
    std::string o;
    for (size_t i = 0; i < s.size(); i++) {
        if (s[i] != '&') { o += s[i]; continue; }
        static const std::pair<const char*, char> ents[] = {{"&amp;", '&'}, {"&lt;", '<'}, {"&gt;", '>'}, {"&quot;", '"'}, {"&apos;", '\''}};
        bool hit = false;
        for (auto& e : ents) {
            const size_t n = strlen(e.first);
            if (s.compare(i, n, e.first) == 0) { o += e.second; i += n - 1; hit = true; break; }
        }
        if (!hit) o += s[i];
    }
    return o;
}
void write(const SerializationNode& node, std::ostream& out, int depth) {
    out << std::string(depth, '\t') << '<' << node.getName();
    for (auto& p : node.getProperties()) out << ' ' << p.first << "=\"" << escape(p.second) << '"';
    if (node.getChildren().empty()) { out << "/>\n"; return; }
    out << ">\n";
    for (const SerializationNode& c : node.getChildren()) write(c, out, depth + 1);
    out << std::string(depth, '\t') << "</" << node.getName() << ">\n";
}
struct Reader {
    const std::string& s; size_t i = 0;
    explicit Reader(const std::string& s) : s(s) {}
    [[noreturn]] void fail(const char* what) const { throw OpenMMException(std::string("XML parse error: ") + what + " at offset " + std::to_string(i)); }
    void ws() { while (i < s.size() && isspace((unsigned char) s[i])) i++; }
    void skipMisc() {                                   // whitespace, <?...?> declarations, <!-- comments -->
        for (;;) {
            ws();
            if (s.compare(i, 2, "<?") == 0) { size_t e = s.find("?>", i); if (e == std::string::npos) fail("unterminated declaration"); i = e + 2; }
            else if (s.compare(i, 4, "<!--") == 0) { size_t e = s.find("-->", i); if (e == std::string::npos) fail("unterminated comment"); i = e + 3; }
            else return;
        }
    }
    std::string ident() {
        size_t b = i;
        while (i < s.size() && (isalnum((unsigned char) s[i]) || s[i] == '_' || s[i] == ':' || s[i] == '-' || s[i] == '.')) i++;
        if (b == i) fail("name expected");
        return s.substr(b, i - b);
    }
    void element(SerializationNode& node) {
        skipMisc();
        if (i >= s.size() || s[i] != '<') fail("'<' expected");
        i++;
        node.setName(ident());
        for (;;) {
            ws();
            if (i >= s.size()) fail("unterminated tag");
            if (s[i] == '/') { if (s.compare(i, 2, "/>") != 0) fail("'/>' expected"); i += 2; return; }
            if (s[i] == '>') { i++; break; }
            const std::string key = ident();
            ws(); if (i >= s.size() || s[i] != '=') fail("'=' expected"); i++; ws();
            if (i >= s.size() || (s[i] != '"' && s[i] != '\'')) fail("quoted value expected");
            const char q = s[i++];
            const size_t e = s.find(q, i);
            if (e == std::string::npos) fail("unterminated value");
            node.setStringProperty(key, unescape(s.substr(i, e - i)));
            i = e + 1;
        }
        for (;;) {
            skipMisc();
            if (s.compare(i, 2, "</") == 0) {
                i += 2;
                if (ident() != node.getName()) fail("mismatched closing tag");
                ws(); if (i >= s.size() || s[i] != '>') fail("'>' expected"); i++;
                return;
            }
            if (i >= s.size()) fail("unterminated element");
            if (s[i] != '<') fail("text content is not supported");
            element(node.createChildNode(""));
        }
    }
};
} // namespace

void XmlSerializer::encode(const SerializationNode& node, std::ostream& stream) {
    stream << "<?xml version=\"1.0\" ?>\n";
    write(node, stream, 0);
}
void XmlSerializer::decode(std::istream& stream, SerializationNode& node) {
    const std::string text((std::istreambuf_iterator<char>(stream)), std::istreambuf_iterator<char>());
    Reader r(text);
    r.element(node);
}
void* XmlSerializer::deserializeStream(std::istream& stream) {
    SerializationNode root;
    decode(stream, root);
    return SerializationProxy::getProxy(root.getStringProperty("type")).deserialize(root);
}

} // namespace OpenMM

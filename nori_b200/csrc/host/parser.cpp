// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// parser.cpp -- Nori XML scene loader (ref: src/parser.cpp:16-305) on a self-contained mini XML reader
// (pugixml, the reference's DOM, is an empty submodule here).  Same grammar and lifecycle: properties are collected
// into the parent's PropertyList, objects are created through NoriObjectFactory::createInstance, children are added
// with addChild()/setParent(), then activate() is called (ref: src/parser.cpp:164-199); transform operations left-
// multiply an accumulator (ref: src/parser.cpp:243-290); attributes are checked strictly (ref: src/parser.cpp:105-116).
#include <fstream>
#include <memory>
#include <set>
#include <sstream>
#include "nori/parser.h"

NORI_NAMESPACE_BEGIN

namespace {

// ---------------------------------------------------------------- tiny XML DOM
struct XmlNode {
    std::string name;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<XmlNode>> children;
    size_t offset = 0;
    const std::string *attr(const std::string &k) const {
        for (auto &a : attrs) if (a.first == k) return &a.second;
        return nullptr;
    }
};

class XmlReader {
public:
    XmlReader(const std::string &text, const std::string &filename) : s(text), file(filename) { }

    std::unique_ptr<XmlNode> parseDocument() {
        std::unique_ptr<XmlNode> root;
        for (;;) {
            skipMisc();
            if (pos >= s.size()) break;
            if (s[pos] != '<') fail("unexpected content");
            auto n = parseElement();
            if (root) fail("multiple root elements");
            root = std::move(n);
        }
        if (!root) fail("no root element");
        return root;
    }

    std::string where(size_t p) const {          // "line N, col M" like ref: src/parser.cpp:21-38
        size_t line = 1, col = 1;
        for (size_t i = 0; i < p && i < s.size(); ++i) { if (s[i] == '\n') { ++line; col = 1; } else ++col; }
        return format("line %i, col %i", (int) line, (int) col);
    }

private:
    const std::string &s; std::string file; size_t pos = 0;

    [[noreturn]] void fail(const char *what) const {
        throw NoriException("Error while parsing \"%s\": %s (at %s)", file, std::string(what), where(pos));
    }
    bool startsWith(const char *t) const { return s.compare(pos, std::strlen(t), t) == 0; }
    void skipWs() { while (pos < s.size() && std::isspace((unsigned char) s[pos])) ++pos; }
    void skipMisc() {     // whitespace, comments, <?xml ... ?>, <!DOCTYPE ...>
        for (;;) {
            skipWs();
            if (startsWith("<!--")) { size_t e = s.find("-->", pos + 4); if (e == std::string::npos) fail("unterminated comment"); pos = e + 3; }
            else if (startsWith("<?")) { size_t e = s.find("?>", pos + 2); if (e == std::string::npos) fail("unterminated declaration"); pos = e + 2; }
            else if (startsWith("<!")) { size_t e = s.find('>', pos + 2); if (e == std::string::npos) fail("unterminated declaration"); pos = e + 1; }
            else return;
        }
    }
    static bool nameChar(char c) { return std::isalnum((unsigned char) c) || c == '_' || c == '-' || c == ':' || c == '.'; }
    std::string parseName() {
        size_t b = pos;
        while (pos < s.size() && nameChar(s[pos])) ++pos;
        if (pos == b) fail("expected a name");
        return s.substr(b, pos - b);
    }
    static std::string unescape(const std::string &v) {
        std::string o; o.reserve(v.size());
        for (size_t i = 0; i < v.size(); ++i) {
            if (v[i] == '&') {
                if (!v.compare(i, 4, "&lt;")) { o += '<'; i += 3; } else if (!v.compare(i, 4, "&gt;")) { o += '>'; i += 3; }
                else if (!v.compare(i, 5, "&amp;")) { o += '&'; i += 4; } else if (!v.compare(i, 6, "&quot;")) { o += '"'; i += 5; }
                else if (!v.compare(i, 6, "&apos;")) { o += '\''; i += 5; } else o += v[i];
            } else o += v[i];
        }
        return o;
    }
    std::unique_ptr<XmlNode> parseElement() {
        std::unique_ptr<XmlNode> n(new XmlNode());
        n->offset = pos;
        ++pos;   // '<'
        n->name = parseName();
        for (;;) {
            skipWs();
            if (pos >= s.size()) fail("unterminated tag");
            if (s[pos] == '/') { if (pos + 1 >= s.size() || s[pos + 1] != '>') fail("malformed tag"); pos += 2; return n; }
            if (s[pos] == '>') { ++pos; break; }
            std::string k = parseName();
            skipWs();
            if (pos >= s.size() || s[pos] != '=') fail("expected '=' after attribute name");
            ++pos; skipWs();
            if (pos >= s.size() || (s[pos] != '"' && s[pos] != '\'')) fail("expected a quoted attribute value");
            char q = s[pos++];
            size_t e = s.find(q, pos);
            if (e == std::string::npos) fail("unterminated attribute value");
            // attribute-value normalisation (XML 1.0 section 3.3.3): literal tab / CR / LF become spaces, so that a
            // list written over several lines (ref: scenes/pa5/tests/test-direct.xml:4-7) tokenises on ", "
            std::string raw = s.substr(pos, e - pos);
            for (char &ch : raw) if (ch == '\t' || ch == '\n' || ch == '\r') ch = ' ';
            n->attrs.emplace_back(k, unescape(raw));
            pos = e + 1;
        }
        for (;;) {   // children
            skipMisc();
            if (pos >= s.size()) fail("unterminated element");
            if (startsWith("</")) {
                pos += 2;
                std::string close = parseName();
                if (close != n->name) fail("mismatched closing tag");
                skipWs();
                if (pos >= s.size() || s[pos] != '>') fail("malformed closing tag");
                ++pos;
                return n;
            }
            if (s[pos] != '<') fail("unexpected content");   // Nori scenes carry no text nodes (ref: src/parser.cpp:127-130)
            n->children.push_back(parseElement());
        }
    }
};

Vector3f toVector3f(const std::string &str) {                          // ref: src/common.cpp:92-100
    std::vector<std::string> tokens = tokenize(str);
    if (tokens.size() != 3) throw NoriException("Expected 3 values");
    return Vector3f(toFloat(tokens[0]), toFloat(tokens[1]), toFloat(tokens[2]));
}

/* Set of supported XML tags (ref: src/parser.cpp:43-72) */
enum ETag {
    EScene = NoriObject::EScene, EMesh = NoriObject::EMesh, EBSDF = NoriObject::EBSDF,
    EPhaseFunction = NoriObject::EPhaseFunction, EEmitter = NoriObject::EEmitter, EMedium = NoriObject::EMedium,
    ECamera = NoriObject::ECamera, EIntegrator = NoriObject::EIntegrator, ESampler = NoriObject::ESampler,
    ETest = NoriObject::ETest, EReconstructionFilter = NoriObject::EReconstructionFilter,
    EBoolean = NoriObject::EClassTypeCount, EInteger, EFloat, EString, EPoint, EVector, EColor, ETransform,
    ETranslate, EMatrix, ERotate, EScale, ELookAt, EInvalid
};

struct ParseContext {
    std::string filename;
    const XmlReader *reader;
    std::map<std::string, ETag> tags;
    Matrix4f transform;      // accumulator of the enclosing <transform>
};

void checkAttributes(const ParseContext &cx, const XmlNode &node, std::set<std::string> attrs) {   // ref: src/parser.cpp:105-116
    for (auto &a : node.attrs) {
        auto it = attrs.find(a.first);
        if (it == attrs.end())
            throw NoriException("Error while parsing \"%s\": unexpected attribute \"%s\" in \"%s\" at %s",
                                cx.filename, a.first, node.name, cx.reader->where(node.offset));
        attrs.erase(it);
    }
    if (!attrs.empty())
        throw NoriException("Error while parsing \"%s\": missing attribute \"%s\" in \"%s\" at %s",
                            cx.filename, *attrs.begin(), node.name, cx.reader->where(node.offset));
}

NoriObject *parseTag(ParseContext &cx, XmlNode &node, PropertyList &list, int parentTag) {
    auto it = cx.tags.find(node.name);
    if (it == cx.tags.end())
        throw NoriException("Error while parsing \"%s\": unexpected tag \"%s\" at %s", cx.filename, node.name, cx.reader->where(node.offset));
    const int tag = it->second;

    /* sanity checks on the tree shape: ref src/parser.cpp:139-157 */
    const bool hasParent = parentTag != EInvalid;
    const bool parentIsObject = hasParent && parentTag < NoriObject::EClassTypeCount;
    const bool currentIsObject = tag < NoriObject::EClassTypeCount;
    const bool parentIsTransform = parentTag == ETransform;
    const bool currentIsTransformOp = tag == ETranslate || tag == ERotate || tag == EScale || tag == ELookAt || tag == EMatrix;
    if (!hasParent && !currentIsObject)
        throw NoriException("Error while parsing \"%s\": root element \"%s\" must be a Nori object (at %s)", cx.filename, node.name, cx.reader->where(node.offset));
    if (parentIsTransform != currentIsTransformOp)
        throw NoriException("Error while parsing \"%s\": transform nodes can only contain transform operations (at %s)", cx.filename, cx.reader->where(node.offset));
    if (hasParent && !parentIsObject && !(parentIsTransform && currentIsTransformOp))
        throw NoriException("Error while parsing \"%s\": node \"%s\" requires a Nori object as parent (at %s)", cx.filename, node.name, cx.reader->where(node.offset));

    if (tag == EScene && !node.attr("type")) node.attrs.emplace_back("type", "scene");   // ref: src/parser.cpp:159-160
    else if (tag == ETransform) cx.transform.setIdentity();

    PropertyList propList;
    std::vector<NoriObject *> children;
    size_t adopted = 0;                 // children[0 .. adopted) belong to `result` (its destructor deletes them)
    try {
        for (auto &ch : node.children) {
            NoriObject *child = parseTag(cx, *ch, propList, tag);
            if (child) children.push_back(child);
        }
    } catch (...) {                     // a later sibling failed: the ones already built have no owner yet
        for (NoriObject *ch : children) delete ch;
        throw;
    }

    NoriObject *result = nullptr;
    try {
        if (currentIsObject) {
            checkAttributes(cx, node, { "type" });
            result = NoriObjectFactory::createInstance(*node.attr("type"), propList);
            if (result->getClassType() != (int) tag)
                throw NoriException("Unexpectedly constructed an object of type <%s> (expected type <%s>): %s",
                                    NoriObject::classTypeName(result->getClassType()),
                                    NoriObject::classTypeName((NoriObject::EClassType) tag), result->toString());
            for (auto ch : children) { result->addChild(ch); ++adopted; ch->setParent(result); }
            result->activate();
        } else {
            auto A = [&](const char *k) -> const std::string & { return *node.attr(k); };
            switch (tag) {
                case EString: checkAttributes(cx, node, { "name", "value" }); list.setString(A("name"), A("value")); break;
                case EFloat: checkAttributes(cx, node, { "name", "value" }); list.setFloat(A("name"), toFloat(A("value"))); break;
                case EInteger: checkAttributes(cx, node, { "name", "value" }); list.setInteger(A("name"), toInt(A("value"))); break;
                case EBoolean: checkAttributes(cx, node, { "name", "value" }); list.setBoolean(A("name"), toBool(A("value"))); break;
                case EPoint: checkAttributes(cx, node, { "name", "value" }); list.setPoint(A("name"), toVector3f(A("value"))); break;
                case EVector: checkAttributes(cx, node, { "name", "value" }); list.setVector(A("name"), toVector3f(A("value"))); break;
                case EColor: {
                    checkAttributes(cx, node, { "name", "value" });
                    Vector3f v = toVector3f(A("value"));
                    list.setColor(A("name"), Color3f(v[0], v[1], v[2]));
                    break;
                }
                case ETransform: checkAttributes(cx, node, { "name" }); list.setTransform(A("name"), Transform(cx.transform)); break;
                case ETranslate: {
                    checkAttributes(cx, node, { "value" });
                    Vector3f v = toVector3f(A("value"));
                    Matrix4f t; t(0, 3) = v[0]; t(1, 3) = v[1]; t(2, 3) = v[2];
                    cx.transform = t * cx.transform;
                    break;
                }
                case EMatrix: {
                    checkAttributes(cx, node, { "value" });
                    std::vector<std::string> tokens = tokenize(A("value"));
                    if (tokens.size() != 16) throw NoriException("Expected 16 values");
                    Matrix4f m;
                    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m(i, j) = toFloat(tokens[i * 4 + j]);
                    cx.transform = m * cx.transform;
                    break;
                }
                case EScale: {
                    checkAttributes(cx, node, { "value" });
                    Vector3f v = toVector3f(A("value"));
                    Matrix4f sc; sc(0, 0) = v[0]; sc(1, 1) = v[1]; sc(2, 2) = v[2];
                    cx.transform = sc * cx.transform;
                    break;
                }
                case ERotate: {   // Eigen::AngleAxis(angle, normalized axis): ref src/parser.cpp:266-272
                    checkAttributes(cx, node, { "angle", "axis" });
                    float angle = degToRad(toFloat(A("angle")));
                    Vector3f a = toVector3f(A("axis")).normalized();
                    double c = std::cos((double) angle), s = std::sin((double) angle), t = 1.0 - c;
                    Matrix4f r;
                    r(0, 0) = (float) (t * a[0] * a[0] + c);        r(0, 1) = (float) (t * a[0] * a[1] - s * a[2]); r(0, 2) = (float) (t * a[0] * a[2] + s * a[1]);
                    r(1, 0) = (float) (t * a[0] * a[1] + s * a[2]); r(1, 1) = (float) (t * a[1] * a[1] + c);        r(1, 2) = (float) (t * a[1] * a[2] - s * a[0]);
                    r(2, 0) = (float) (t * a[0] * a[2] - s * a[1]); r(2, 1) = (float) (t * a[1] * a[2] + s * a[0]); r(2, 2) = (float) (t * a[2] * a[2] + c);
                    cx.transform = r * cx.transform;
                    break;
                }
                case ELookAt: {   // columns [left, newUp, dir, origin]: ref src/parser.cpp:274-289
                    checkAttributes(cx, node, { "origin", "target", "up" });
                    Vector3f origin = toVector3f(A("origin")), target = toVector3f(A("target")), up = toVector3f(A("up"));
                    Vector3f dir = (target - origin).normalized();
                    Vector3f left = up.normalized().cross(dir).normalized();
                    Vector3f newUp = dir.cross(left).normalized();
                    Matrix4f trafo;
                    for (int i = 0; i < 3; ++i) { trafo(i, 0) = left[i]; trafo(i, 1) = newUp[i]; trafo(i, 2) = dir[i]; trafo(i, 3) = origin[i]; }
                    cx.transform = trafo * cx.transform;
                    break;
                }
                default: throw NoriException("Unhandled element \"%s\"", node.name);
            }
        }
    } catch (const NoriException &e) {
        for (size_t i = adopted; i < children.size(); ++i) delete children[i];   // not (yet) owned by anybody
        delete result;                                                           // takes the adopted children with it
        throw NoriException("Error while parsing \"%s\": %s (at %s)", cx.filename, std::string(e.what()), cx.reader->where(node.offset));
    }
    return result;
}

}  // namespace

// Directory of the scene file being loaded.  Per thread, and restored when loadFromXML returns, so that concurrent or
// nested loads resolve their OBJ paths against their own scene file.
static thread_local std::string g_sceneDir;
const std::string &sceneDirectory() { return g_sceneDir; }

std::string resolvePath(const std::string &name) {   // stands in for filesystem::resolver (ref: src/main.cpp:195-197, src/obj.cpp:24-25)
    if (!name.empty() && name[0] == '/') return name;
    return g_sceneDir.empty() ? name : g_sceneDir + "/" + name;
}

NoriObject *loadFromXMLString(const std::string &text, const std::string &filename) {
    XmlReader reader(text, filename);
    std::unique_ptr<XmlNode> root = reader.parseDocument();
    ParseContext cx;
    cx.filename = filename; cx.reader = &reader;
    cx.tags = { { "scene", EScene }, { "mesh", EMesh }, { "bsdf", EBSDF }, { "emitter", EEmitter }, { "camera", ECamera },
                { "medium", EMedium }, { "phase", EPhaseFunction }, { "integrator", EIntegrator }, { "sampler", ESampler },
                { "rfilter", EReconstructionFilter }, { "test", ETest }, { "boolean", EBoolean }, { "integer", EInteger },
                { "float", EFloat }, { "string", EString }, { "point", EPoint }, { "vector", EVector }, { "color", EColor },
                { "transform", ETransform }, { "translate", ETranslate }, { "matrix", EMatrix }, { "rotate", ERotate },
                { "scale", EScale }, { "lookat", ELookAt } };
    PropertyList list;
    return parseTag(cx, *root, list, EInvalid);
}

NoriObject *loadFromXML(const std::string &filename) {
    std::ifstream is(filename);
    if (is.fail()) throw NoriException("Error while parsing \"%s\": unable to open the file", filename);
    std::stringstream ss; ss << is.rdbuf();
    size_t slash = filename.find_last_of('/');
    struct DirScope {
        std::string saved;
        explicit DirScope(std::string dir) : saved(g_sceneDir) { g_sceneDir = std::move(dir); }
        ~DirScope() { g_sceneDir = saved; }
    } scope(slash == std::string::npos ? std::string(".") : filename.substr(0, slash));
    return loadFromXMLString(ss.str(), filename);
}

NORI_NAMESPACE_END

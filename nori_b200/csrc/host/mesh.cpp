// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// mesh.cpp -- Mesh base class + Wavefront OBJ loader plugin "obj" (ref: src/mesh.cpp:16-29,96-123, src/obj.cpp:19-163).
#include <fstream>
#include <sstream>
#include <unordered_map>
#include <sys/stat.h>
#include "nori/parser.h"
#include "nori/plugins.h"

NORI_NAMESPACE_BEGIN

Mesh::~Mesh() { delete m_bsdf; delete m_emitter; }

void Mesh::activate() {
    if (!m_bsdf) /* If no material was assigned, instantiate a diffuse BRDF (ref: src/mesh.cpp:23-29) */
        m_bsdf = static_cast<BSDF *>(NoriObjectFactory::createInstance("diffuse", PropertyList()));
}

void Mesh::addChild(NoriObject *obj) {
    switch (obj->getClassType()) {
        case EBSDF:
            if (m_bsdf) throw NoriException("Mesh: tried to register multiple BSDF instances!");
            m_bsdf = static_cast<BSDF *>(obj);
            break;
        case EEmitter:
            if (m_emitter) throw NoriException("Mesh: tried to register multiple Emitter instances!");
            m_emitter = static_cast<Emitter *>(obj);
            break;
        default:
            throw NoriException("Mesh::addChild(<%s>) is not supported!", classTypeName(obj->getClassType()));
    }
}

std::string Mesh::toString() const {
    return format("Mesh[\n  name = \"%s\",\n  vertexCount = %i,\n  triangleCount = %i,\n  bsdf = %s,\n  emitter = %s\n]",
                  m_name, (int) getVertexCount(), (int) getTriangleCount(),
                  m_bsdf ? indent(m_bsdf->toString()) : std::string("null"),
                  m_emitter ? indent(m_emitter->toString()) : std::string("null"));
}

/// Loader for Wavefront OBJ triangle meshes
class WavefrontOBJ : public Mesh {
public:
    WavefrontOBJ(const PropertyList &propList) {
        typedef std::unordered_map<OBJVertex, uint32_t, OBJVertexHash> VertexMap;
        std::string filename = resolvePath(propList.getString("filename"));
        std::ifstream is(filename);
        if (is.fail()) throw NoriException("Unable to open OBJ file \"%s\"!", filename);
        Transform trafo = propList.getTransform("toWorld", Transform());
        m_name = filename;
        /* binary cache of the parsed arrays next to the OBJ: text parsing through istringstream +
           hash-map dedup is the slowest load step for 500k-triangle meshes.  Opt in with <boolean name="cache" value="true"/>. */
        const bool useCache = propList.getBoolean("cache", false);
        if (useCache && loadCache(filename, trafo)) return;

        std::vector<Vector3f> positions, normals;
        std::vector<Point2f> texcoords;
        std::vector<uint32_t> indices;
        std::vector<OBJVertex> vertices;
        VertexMap vertexMap;

        std::string line_str;
        while (std::getline(is, line_str)) {
            std::istringstream line(line_str);
            std::string prefix;
            line >> prefix;
            if (prefix == "v") {
                Point3f p;
                line >> p[0] >> p[1] >> p[2];
                positions.push_back(trafo.applyPoint(p));
            } else if (prefix == "vt") {
                Point2f tc;
                line >> tc.x >> tc.y;
                texcoords.push_back(tc);
            } else if (prefix == "vn") {
                Normal3f n;
                line >> n[0] >> n[1] >> n[2];
                normals.push_back(trafo.applyNormal(n).normalized());
            } else if (prefix == "f") {
                std::string v1, v2, v3, v4;
                line >> v1 >> v2 >> v3 >> v4;
                OBJVertex verts[6];
                int nVertices = 3;
                verts[0] = OBJVertex(v1); verts[1] = OBJVertex(v2); verts[2] = OBJVertex(v3);
                if (!v4.empty()) {   /* quad -> two triangles (ref: src/obj.cpp:73-79) */
                    verts[3] = OBJVertex(v4); verts[4] = verts[0]; verts[5] = verts[2];
                    nVertices = 6;
                }
                for (int i = 0; i < nVertices; ++i) {   /* indexed vertex list, dedup on (p, uv, n) (ref: src/obj.cpp:81-91) */
                    const OBJVertex &v = verts[i];
                    auto it = vertexMap.find(v);
                    if (it == vertexMap.end()) {
                        vertexMap[v] = (uint32_t) vertices.size();
                        indices.push_back((uint32_t) vertices.size());
                        vertices.push_back(v);
                    } else {
                        indices.push_back(it->second);
                    }
                }
            }
        }
        m_F = indices;
        m_V.resize(vertices.size() * 3);
        for (size_t i = 0; i < vertices.size(); ++i) {
            if (vertices[i].p - 1 >= positions.size()) throw NoriException("Invalid vertex data in \"%s\"", filename);
            for (int k = 0; k < 3; ++k) m_V[3 * i + k] = positions[vertices[i].p - 1][k];
        }
        if (!normals.empty()) {
            m_N.resize(vertices.size() * 3);
            for (size_t i = 0; i < vertices.size(); ++i) {
                if (vertices[i].n - 1 >= normals.size()) throw NoriException("Invalid normal data in \"%s\"", filename);
                for (int k = 0; k < 3; ++k) m_N[3 * i + k] = normals[vertices[i].n - 1][k];
            }
        }
        if (!texcoords.empty()) {
            m_UV.resize(vertices.size() * 2);
            for (size_t i = 0; i < vertices.size(); ++i) {
                if (vertices[i].uv - 1 >= texcoords.size()) throw NoriException("Invalid texcoord data in \"%s\"", filename);
                m_UV[2 * i] = texcoords[vertices[i].uv - 1].x; m_UV[2 * i + 1] = texcoords[vertices[i].uv - 1].y;
            }
        }
        if (useCache) saveCache(filename, trafo);
    }

protected:
    struct CacheHeader { char magic[8]; uint64_t objSize; int64_t objMtime; float trafo[16]; uint32_t nv, nf, hasN, hasUV; };

    static bool statFile(const std::string &f, uint64_t &size, int64_t &mtime) {
        struct stat st;
        if (stat(f.c_str(), &st) != 0) return false;
        size = (uint64_t) st.st_size; mtime = (int64_t) st.st_mtime;
        return true;
    }
    void fillHeader(const std::string &filename, const Transform &trafo, CacheHeader &h) const {
        std::memset(&h, 0, sizeof h);
        std::memcpy(h.magic, "NBMESH01", 8);
        statFile(filename, h.objSize, h.objMtime);
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) h.trafo[4 * i + j] = trafo.getMatrix()(i, j);
    }
    bool loadCache(const std::string &filename, const Transform &trafo) {
        std::ifstream cs(filename + ".nbcache", std::ios::binary);
        if (!cs) return false;
        CacheHeader want, got;
        fillHeader(filename, trafo, want);
        cs.read((char *) &got, sizeof got);
        if (!cs || std::memcmp(got.magic, want.magic, 8) || got.objSize != want.objSize || got.objMtime != want.objMtime ||
            std::memcmp(got.trafo, want.trafo, sizeof got.trafo)) return false;
        m_V.resize((size_t) got.nv * 3); m_F.resize((size_t) got.nf * 3);
        cs.read((char *) m_V.data(), (std::streamsize) (m_V.size() * 4));
        cs.read((char *) m_F.data(), (std::streamsize) (m_F.size() * 4));
        if (got.hasN) { m_N.resize((size_t) got.nv * 3); cs.read((char *) m_N.data(), (std::streamsize) (m_N.size() * 4)); }
        if (got.hasUV) { m_UV.resize((size_t) got.nv * 2); cs.read((char *) m_UV.data(), (std::streamsize) (m_UV.size() * 4)); }
        if (!cs) { m_V.clear(); m_F.clear(); m_N.clear(); m_UV.clear(); return false; }
        return true;
    }
    void saveCache(const std::string &filename, const Transform &trafo) const {
        std::ofstream cs(filename + ".nbcache", std::ios::binary);
        if (!cs) return;   /* read-only scene directory: silently skip */
        CacheHeader h; fillHeader(filename, trafo, h);
        h.nv = getVertexCount(); h.nf = getTriangleCount(); h.hasN = !m_N.empty(); h.hasUV = !m_UV.empty();
        cs.write((const char *) &h, sizeof h);
        cs.write((const char *) m_V.data(), (std::streamsize) (m_V.size() * 4));
        cs.write((const char *) m_F.data(), (std::streamsize) (m_F.size() * 4));
        if (h.hasN) cs.write((const char *) m_N.data(), (std::streamsize) (m_N.size() * 4));
        if (h.hasUV) cs.write((const char *) m_UV.data(), (std::streamsize) (m_UV.size() * 4));
    }

    struct OBJVertex {   /* ref: src/obj.cpp:122-147 */
        uint32_t p = (uint32_t) -1, n = (uint32_t) -1, uv = (uint32_t) -1;
        OBJVertex() { }
        OBJVertex(const std::string &string) {
            std::vector<std::string> tokens = tokenize(string, "/", true);
            if (tokens.size() < 1 || tokens.size() > 3) throw NoriException("Invalid vertex data: \"%s\"", string);
            p = toUInt(tokens[0]);
            if (tokens.size() >= 2 && !tokens[1].empty()) uv = toUInt(tokens[1]);
            if (tokens.size() >= 3 && !tokens[2].empty()) n = toUInt(tokens[2]);
        }
        bool operator==(const OBJVertex &v) const { return v.p == p && v.n == n && v.uv == uv; }
    };
    struct OBJVertexHash {
        std::size_t operator()(const OBJVertex &v) const {
            size_t hash = std::hash<uint32_t>()(v.p);
            hash = hash * 37 + std::hash<uint32_t>()(v.uv);
            hash = hash * 37 + std::hash<uint32_t>()(v.n);
            return hash;
        }
    };
};

NORI_REGISTER_CLASS(WavefrontOBJ, "obj");
NORI_NAMESPACE_END

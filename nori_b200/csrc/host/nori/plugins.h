// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// plugins.h -- abstract plugin interfaces of the hot path, mirroring ref: include/nori/{bsdf,emitter,integrator,
// sampler,camera,rfilter,mesh,scene}.h.  The virtual evaluation entry points exist so that plugins written against
// Nori's API compile unchanged; on this path radiance is evaluated by the CUDA kernels, so the shipped host plugins
// only implement what the host pipeline itself needs (parameters, filter eval, camera matrices).
#pragma once
#include "object.h"
#include "pcg32.h"

NORI_NAMESPACE_BEGIN

enum EMeasure { EUnknownMeasure = 0, ESolidAngle, EDiscrete };   // ref: include/nori/common.h:179-183

struct BSDFQueryRecord {                                          // ref: include/nori/bsdf.h:17-38
    Vector3f wi, wo; float eta; EMeasure measure;
    BSDFQueryRecord(const Vector3f &wi_) : wi(wi_), eta(1.f), measure(EUnknownMeasure) { }
    BSDFQueryRecord(const Vector3f &wi_, const Vector3f &wo_, EMeasure m) : wi(wi_), wo(wo_), eta(1.f), measure(m) { }
};

class BSDF : public NoriObject {                                  // ref: include/nori/bsdf.h:43-100
public:
    virtual Color3f sample(BSDFQueryRecord &bRec, const Point2f &sample) const = 0;
    virtual Color3f eval(const BSDFQueryRecord &bRec) const = 0;
    virtual float pdf(const BSDFQueryRecord &bRec) const = 0;
    EClassType getClassType() const { return EBSDF; }
    virtual bool isDiffuse() const { return false; }
};

class Emitter : public NoriObject {                               // ref: include/nori/emitter.h:16-24
public:
    EClassType getClassType() const { return EEmitter; }
};

class ReconstructionFilter : public NoriObject {                  // ref: include/nori/rfilter.h:24-41
public:
    float getRadius() const { return m_radius; }
    virtual float eval(float x) const = 0;
    EClassType getClassType() const { return EReconstructionFilter; }
protected:
    float m_radius;
};

class ImageBlock;
class Sampler : public NoriObject {                               // ref: include/nori/sampler.h:36-93
public:
    virtual ~Sampler() { }
    virtual std::unique_ptr<Sampler> clone() const = 0;
    virtual void prepare(const ImageBlock &block) = 0;
    virtual void generate() = 0;
    virtual void advance() = 0;
    virtual float next1D() = 0;
    virtual Point2f next2D() = 0;
    virtual size_t getSampleCount() const { return m_sampleCount; }
    EClassType getClassType() const { return ESampler; }
protected:
    size_t m_sampleCount;
};

struct Ray3f {                                                    // ref: include/nori/ray.h:25-64
    Point3f o; Vector3f d, dRcp; float mint, maxt;
    Ray3f() : mint(Epsilon), maxt(std::numeric_limits<float>::infinity()) { }
    void update() { dRcp = Vector3f(1.0f / d[0], 1.0f / d[1], 1.0f / d[2]); }
};

class Camera : public NoriObject {                                // ref: include/nori/camera.h:22-60
public:
    virtual Color3f sampleRay(Ray3f &ray, const Point2f &samplePosition, const Point2f &apertureSample) const = 0;
    const Vector2i &getOutputSize() const { return m_outputSize; }
    const ReconstructionFilter *getReconstructionFilter() const { return m_rfilter; }
    EClassType getClassType() const { return ECamera; }
protected:
    Vector2i m_outputSize;
    ReconstructionFilter *m_rfilter = nullptr;
};

class Scene;
class Integrator : public NoriObject {                            // ref: include/nori/integrator.h:20-49
public:
    virtual ~Integrator() { }
    virtual void preprocess(const Scene *) { }
    virtual Color3f Li(const Scene *scene, Sampler *sampler, const Ray3f &ray) const = 0;
    EClassType getClassType() const { return EIntegrator; }
};

/// Triangle mesh storage in the reference's layout (ref: include/nori/mesh.h:159-166): packed xyz positions /
/// normals, uv pairs, uint32 index triples; world space (toWorld applied at load, ref: src/obj.cpp:52).
class Mesh : public NoriObject {
public:
    virtual ~Mesh();
    virtual void activate();                                      // default diffuse BSDF: ref src/mesh.cpp:23-29
    uint32_t getTriangleCount() const { return (uint32_t) (m_F.size() / 3); }
    uint32_t getVertexCount() const { return (uint32_t) (m_V.size() / 3); }
    const std::vector<float> &getVertexPositions() const { return m_V; }
    const std::vector<float> &getVertexNormals() const { return m_N; }
    const std::vector<float> &getVertexTexCoords() const { return m_UV; }
    const std::vector<uint32_t> &getIndices() const { return m_F; }
    bool isEmitter() const { return m_emitter != nullptr; }
    Emitter *getEmitter() { return m_emitter; }
    const Emitter *getEmitter() const { return m_emitter; }
    const BSDF *getBSDF() const { return m_bsdf; }
    virtual void addChild(NoriObject *child);                     // ref: src/mesh.cpp:96-123
    const std::string &getName() const { return m_name; }
    std::string toString() const;
    EClassType getClassType() const { return EMesh; }
protected:
    Mesh() { }
    std::string m_name;
    std::vector<float> m_V, m_N, m_UV;
    std::vector<uint32_t> m_F;
    BSDF *m_bsdf = nullptr;
    Emitter *m_emitter = nullptr;
};

/// Scene object (ref: include/nori/scene.h:21-113, src/scene.cpp): owns meshes / camera / sampler / integrator.  The
/// reference's Accel member is replaced by the GPU context (nb_ctx) that render() creates from the scene.
class Scene : public NoriObject {
public:
    Scene(const PropertyList &);
    virtual ~Scene();
    const Integrator *getIntegrator() const { return m_integrator; }
    Integrator *getIntegrator() { return m_integrator; }
    const Camera *getCamera() const { return m_camera; }
    const Sampler *getSampler() const { return m_sampler; }
    Sampler *getSampler() { return m_sampler; }
    const std::vector<Mesh *> &getMeshes() const { return m_meshes; }
    void activate();
    void addChild(NoriObject *obj);
    std::string toString() const;
    EClassType getClassType() const { return EScene; }
private:
    std::vector<Mesh *> m_meshes;
    Integrator *m_integrator = nullptr;
    Sampler *m_sampler = nullptr;
    Camera *m_camera = nullptr;
};

NORI_NAMESPACE_END

// scene.h -- Scene object (ref: include/nori/scene.h:21-113, src/scene.cpp).  Owns meshes / camera / sampler /
// integrator.  The reference's Accel member is replaced by the GPU context (nb_ctx), created lazily by render().
#pragma once
#include "plugins.h"

NORI_NAMESPACE_BEGIN

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

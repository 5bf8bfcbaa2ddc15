// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// scene.cpp -- ref: src/scene.cpp:16-109
#include "nori/scene.h"

NORI_NAMESPACE_BEGIN

Scene::Scene(const PropertyList &) { }

Scene::~Scene() {
    delete m_sampler; delete m_camera; delete m_integrator;
    for (auto m : m_meshes) delete m;
}

void Scene::activate() {
    /* the acceleration structure is built on first render (nb_build_accel), where Accel::build ran in the reference */
    if (!m_integrator) throw NoriException("No integrator was specified!");
    if (!m_camera) throw NoriException("No camera was specified!");
    if (!m_sampler) /* Create a default (independent) sampler (ref: src/scene.cpp:35-39) */
        m_sampler = static_cast<Sampler *>(NoriObjectFactory::createInstance("independent", PropertyList()));
}

void Scene::addChild(NoriObject *obj) {
    switch (obj->getClassType()) {
        case EMesh: m_meshes.push_back(static_cast<Mesh *>(obj)); break;
        case EEmitter:
            throw NoriException("Scene::addChild(): emitters are attached to meshes (<emitter> inside <mesh>)");
        case ESampler:
            if (m_sampler) throw NoriException("There can only be one sampler per scene!");
            m_sampler = static_cast<Sampler *>(obj);
            break;
        case ECamera:
            if (m_camera) throw NoriException("There can only be one camera per scene!");
            m_camera = static_cast<Camera *>(obj);
            break;
        case EIntegrator:
            if (m_integrator) throw NoriException("There can only be one integrator per scene!");
            m_integrator = static_cast<Integrator *>(obj);
            break;
        default:
            throw NoriException("Scene::addChild(<%s>) is not supported!", classTypeName(obj->getClassType()));
    }
}

std::string Scene::toString() const {
    std::string meshes;
    for (size_t i = 0; i < m_meshes.size(); ++i) {
        meshes += std::string("  ") + indent(m_meshes[i]->toString(), 2);
        if (i + 1 < m_meshes.size()) meshes += ",";
        meshes += "\n";
    }
    return format("Scene[\n  integrator = %s,\n  sampler = %s\n  camera = %s,\n  meshes = {\n  %s  }\n]",
                  indent(m_integrator->toString()), indent(m_sampler->toString()), indent(m_camera->toString()), indent(meshes, 2));
}

NORI_REGISTER_CLASS(Scene, "scene");
NORI_NAMESPACE_END

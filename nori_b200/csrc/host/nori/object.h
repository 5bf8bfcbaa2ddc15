// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// object.h -- NoriObject base class, factory and NORI_REGISTER_CLASS (ref: include/nori/object.h:20-149).
// Same names and semantics: plugins are translation units that self-register at static-init time; the parser
// constructs them by name, adds children, then calls activate().  One addition (marked NB): the factory records,
// for every object it creates, the (type name, PropertyList) it was created from, so that the GPU path can build
// plain-data descriptors for plugin instances WITHOUT touching the plugin classes.
#pragma once
#include <functional>
#include "proplist.h"

NORI_NAMESPACE_BEGIN

class NoriObject {
public:
    enum EClassType {
        EScene = 0, EMesh, EBSDF, EPhaseFunction, EEmitter, EMedium, ECamera, EIntegrator, ESampler, ETest,
        EReconstructionFilter, EClassTypeCount
    };
    virtual ~NoriObject() { }
    virtual EClassType getClassType() const = 0;
    virtual void addChild(NoriObject *child);
    virtual void setParent(NoriObject *parent);
    virtual void activate();
    virtual std::string toString() const = 0;
    static std::string classTypeName(EClassType type) {
        switch (type) {
            case EScene: return "scene"; case EMesh: return "mesh"; case EBSDF: return "bsdf";
            case EEmitter: return "emitter"; case ECamera: return "camera"; case EIntegrator: return "integrator";
            case ESampler: return "sampler"; case ETest: return "test"; default: return "<unknown>";
        }
    }
};

NORI_NAMESPACE_END

#include "factory.h"

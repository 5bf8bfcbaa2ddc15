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

class NoriObjectFactory {
public:
    typedef std::function<NoriObject *(const PropertyList &)> Constructor;
    static void registerClass(const std::string &name, const Constructor &constr);
    static NoriObject *createInstance(const std::string &name, const PropertyList &propList);
    /// NB: creation record of an object made by createInstance (nullptr if unknown)
    struct Record { std::string type; PropertyList props; };
    static const Record *creationRecord(const NoriObject *obj);
    static void forgetRecord(const NoriObject *obj);
    static bool isRegistered(const std::string &name);
private:
    static std::map<std::string, Constructor> *m_constructors;
    static std::map<const NoriObject *, Record> *m_records;
};

/// Macro for registering an object constructor with the NoriObjectFactory (ref: include/nori/object.h:141-149)
#define NORI_REGISTER_CLASS(cls, name) \
    cls *cls ##_create(const PropertyList &list) { \
        return new cls(list); \
    } \
    static struct cls ##_{ \
        cls ##_() { \
            NoriObjectFactory::registerClass(name, cls ##_create); \
        } \
    } cls ##__NORI_;

NORI_NAMESPACE_END

// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// factory.h -- NoriObjectFactory and NORI_REGISTER_CLASS (ref: include/nori/object.h:100-149); included by object.h.
// Plugins are translation units that self-register at static-initialisation time under the name the XML scenes use.
// NB (addition): the factory keeps the (type name, PropertyList) every object was created from, which is what lets the
// GPU path describe plugin instances as plain data without touching the plugin classes.
#pragma once
#include "object.h"

NORI_NAMESPACE_BEGIN

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

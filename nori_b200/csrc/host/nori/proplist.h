// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// proplist.h -- typed name -> value map handed to plugin constructors (ref: include/nori/proplist.h:21-123).
#pragma once
#include <map>
#include "transform.h"

NORI_NAMESPACE_BEGIN

class PropertyList {
public:
    PropertyList() { }
#define NORI_PROP_DECL(Type, TypeName) \
    void set##TypeName(const std::string &name, const Type &value); \
    Type get##TypeName(const std::string &name) const; \
    Type get##TypeName(const std::string &name, const Type &defaultValue) const;
    NORI_PROP_DECL(bool, Boolean)
    NORI_PROP_DECL(int, Integer)
    NORI_PROP_DECL(float, Float)
    NORI_PROP_DECL(std::string, String)
    NORI_PROP_DECL(Color3f, Color)
    NORI_PROP_DECL(Point3f, Point)
    NORI_PROP_DECL(Vector3f, Vector)
    NORI_PROP_DECL(Transform, Transform)
#undef NORI_PROP_DECL
    /// True if a property of that name exists (any type)
    bool has(const std::string &name) const { return m_properties.find(name) != m_properties.end(); }
private:
    struct Property {
        enum { boolean_type, integer_type, float_type, string_type, color_type, point_type, vector_type, transform_type } type;
        struct Value {
            bool boolean_value = false; int integer_value = 0; float float_value = 0;
            std::string string_value; Color3f color_value; Point3f point_value; Vector3f vector_value; Transform transform_value;
        } value;
        Property() : type(boolean_type) { }
    };
    std::map<std::string, Property> m_properties;
};

NORI_NAMESPACE_END

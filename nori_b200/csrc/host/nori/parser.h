// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// parser.h -- ref: include/nori/parser.h:17
#pragma once
#include "object.h"

NORI_NAMESPACE_BEGIN
/// Load a scene from the specified filename and return its root object
extern NoriObject *loadFromXML(const std::string &filename);
/// Same, from an in-memory document (filename is used in error messages only)
extern NoriObject *loadFromXMLString(const std::string &text, const std::string &filename);
/// Resolve a path relative to the directory of the scene being parsed (filesystem::resolver stand-in)
extern std::string resolvePath(const std::string &name);
NORI_NAMESPACE_END

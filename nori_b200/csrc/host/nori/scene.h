// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// scene.h -- kept so that sources written against Nori's headers (#include <nori/scene.h>) compile unchanged; the
// Scene class itself is declared with the other object interfaces in plugins.h.
#pragma once
#include "plugins.h"

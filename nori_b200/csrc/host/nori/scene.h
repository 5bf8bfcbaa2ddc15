// scene.h -- kept so that sources written against Nori's headers (#include <nori/scene.h>) compile unchanged; the
// Scene class itself is declared with the other object interfaces in plugins.h.
#pragma once
#include "plugins.h"

// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// vector.h -- minimal fixed-size vector / colour types standing in for the Eigen-based ones of
// ref: include/nori/vector.h and include/nori/color.h (only what the host pipeline needs).
#pragma once
#include "common.h"

NORI_NAMESPACE_BEGIN

struct Vector3f {
    float v[3];
    Vector3f() : v{0, 0, 0} { }
    explicit Vector3f(float s) : v{s, s, s} { }
    Vector3f(float x, float y, float z) : v{x, y, z} { }
    float &operator[](int i) { return v[i]; }
    float operator[](int i) const { return v[i]; }
    float x() const { return v[0]; } float y() const { return v[1]; } float z() const { return v[2]; }
    Vector3f operator+(const Vector3f &o) const { return Vector3f(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    Vector3f operator-(const Vector3f &o) const { return Vector3f(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vector3f operator*(float s) const { return Vector3f(v[0] * s, v[1] * s, v[2] * s); }
    float dot(const Vector3f &o) const { return v[0] * o.v[0] + (v[1] * o.v[1] + v[2] * o.v[2]); }
    Vector3f cross(const Vector3f &o) const {
        return Vector3f(v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]);
    }
    float norm() const { return std::sqrt(dot(*this)); }
    Vector3f normalized() const { float n = norm(); return Vector3f(v[0] / n, v[1] / n, v[2] / n); }
    std::string toString() const { return format("[%f, %f, %f]", v[0], v[1], v[2]); }
};
typedef Vector3f Point3f;
typedef Vector3f Normal3f;

struct Point2f { float x, y; Point2f() : x(0), y(0) { } Point2f(float x_, float y_) : x(x_), y(y_) { } };
struct Vector2i {
    int v[2];
    Vector2i() : v{0, 0} { }
    Vector2i(int x, int y) : v{x, y} { }
    int x() const { return v[0]; } int y() const { return v[1]; }
    int &x() { return v[0]; } int &y() { return v[1]; }
    std::string toString() const { return format("[%i, %i]", v[0], v[1]); }
};
typedef Vector2i Point2i;

/// RGB colour (ref: include/nori/color.h:17-58)
struct Color3f {
    float c[3];
    Color3f() : c{0, 0, 0} { }
    explicit Color3f(float s) : c{s, s, s} { }
    Color3f(float r, float g, float b) : c{r, g, b} { }
    float r() const { return c[0]; } float g() const { return c[1]; } float b() const { return c[2]; }
    float maxCoeff() const { return std::max(c[0], std::max(c[1], c[2])); }
    bool isValid() const {                                      // ref: src/common.cpp:196-203
        for (int i = 0; i < 3; ++i) if (c[i] < 0 || !std::isfinite(c[i])) return false;
        return true;
    }
    float getLuminance() const { return c[0] * 0.212671f + c[1] * 0.715160f + c[2] * 0.072169f; }   // ref: src/common.cpp:206-208
    Color3f toSRGB() const;                                     // ref: src/common.cpp:166-180
    std::string toString() const { return format("[%f, %f, %f]", c[0], c[1], c[2]); }
};

/// RGBA colour with a filter weight (ref: include/nori/color.h:76-110)
struct Color4f {
    float c[4];
    Color4f() : c{0, 0, 0, 0} { }
    Color4f(const Color3f &v) : c{v.c[0], v.c[1], v.c[2], 1.0f} { }
    Color4f(float r, float g, float b, float w) : c{r, g, b, w} { }
    Color3f divideByFilterWeight() const {                      // ref: include/nori/color.h:100-105
        if (c[3] != 0) return Color3f(c[0] / c[3], c[1] / c[3], c[2] / c[3]);
        return Color3f(0.0f);
    }
};

NORI_NAMESPACE_END

// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// transform.h -- homogeneous transform + inverse (ref: include/nori/transform.h:22-83).  Row-major 4x4 fp32; the
// inverse is computed in double and rounded once (Eigen's fp32 inverse is not reproducible without Eigen).
#pragma once
#include "vector.h"

NORI_NAMESPACE_BEGIN

struct Matrix4f {
    float m[4][4];
    Matrix4f() { setIdentity(); }
    void setIdentity() { for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m[i][j] = i == j ? 1.0f : 0.0f; }
    float &operator()(int i, int j) { return m[i][j]; }
    float operator()(int i, int j) const { return m[i][j]; }
    Matrix4f operator*(const Matrix4f &o) const;     // accumulated in double, rounded once
    Matrix4f inverse() const;                        // Gauss-Jordan in double
    std::string toString() const;
};

struct Transform {
public:
    Transform() { }
    Transform(const Matrix4f &trafo) : m_transform(trafo), m_inverse(trafo.inverse()) { }
    Transform(const Matrix4f &trafo, const Matrix4f &inv) : m_transform(trafo), m_inverse(inv) { }
    const Matrix4f &getMatrix() const { return m_transform; }
    const Matrix4f &getInverseMatrix() const { return m_inverse; }
    Transform inverse() const { return Transform(m_inverse, m_transform); }
    Transform operator*(const Transform &t) const { return Transform(m_transform * t.m_transform, t.m_inverse * m_inverse); }
    /// Apply to a vector (3x3 part): ref include/nori/transform.h:55-57
    Vector3f applyVector(const Vector3f &v) const;
    /// Apply to a normal (inverse transpose): ref include/nori/transform.h:60-62
    Normal3f applyNormal(const Normal3f &n) const;
    /// Apply to a point with homogeneous divide: ref include/nori/transform.h:65-68
    Point3f applyPoint(const Point3f &p) const;
    std::string toString() const { return m_transform.toString(); }
private:
    Matrix4f m_transform, m_inverse;
};

NORI_NAMESPACE_END

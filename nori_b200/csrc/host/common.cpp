// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// common.cpp -- string helpers, Matrix4f / Transform arithmetic, object + factory, property list.
// (ref: src/common.cpp, src/object.cpp, src/proplist.cpp)
#include <cstring>
#include <algorithm>
#include <iomanip>
#include <sstream>
#include "nori/object.h"

NORI_NAMESPACE_BEGIN

std::string indent(const std::string &string, int amount) {           // ref: src/common.cpp:19-28
    std::istringstream iss(string);
    std::ostringstream oss;
    std::string spacer((size_t) amount, ' ');
    bool firstLine = true;
    for (std::string line; std::getline(iss, line);) {
        if (!firstLine) oss << spacer;
        oss << line;
        if (!iss.eof()) oss << endl;
        firstLine = false;
    }
    return oss.str();
}

std::string toLower(const std::string &value) {
    std::string result(value);
    std::transform(result.begin(), result.end(), result.begin(), [](unsigned char c) { return (char) std::tolower(c); });
    return result;
}

bool toBool(const std::string &str) {                                 // ref: src/common.cpp:58-66
    std::string value = toLower(str);
    if (value == "false") return false;
    else if (value == "true") return true;
    else throw NoriException("Could not parse boolean value \"%s\"", str);
}

int toInt(const std::string &str) {                                   // ref: src/common.cpp:68-74
    char *end_ptr = nullptr;
    int result = (int) strtol(str.c_str(), &end_ptr, 10);
    if (*end_ptr != '\0') throw NoriException("Could not parse integer value \"%s\"", str);
    return result;
}

unsigned int toUInt(const std::string &str) {                         // ref: src/common.cpp:76-82
    char *end_ptr = nullptr;
    unsigned int result = (unsigned int) strtoul(str.c_str(), &end_ptr, 10);
    if (*end_ptr != '\0') throw NoriException("Could not parse integer value \"%s\"", str);
    return result;
}

float toFloat(const std::string &str) {                               // ref: src/common.cpp:84-90
    char *end_ptr = nullptr;
    float result = (float) strtof(str.c_str(), &end_ptr);
    if (*end_ptr != '\0') throw NoriException("Could not parse floating point value \"%s\"", str);
    return result;
}

std::vector<std::string> tokenize(const std::string &string, const std::string &delim, bool includeEmpty) {   // ref: src/common.cpp:102-117
    std::string::size_type lastPos = 0, pos = string.find_first_of(delim, lastPos);
    std::vector<std::string> tokens;
    while (lastPos != std::string::npos) {
        if (pos != lastPos || includeEmpty) tokens.push_back(string.substr(lastPos, pos - lastPos));
        lastPos = pos;
        if (lastPos != std::string::npos) { lastPos += 1; pos = string.find_first_of(delim, lastPos); }
    }
    return tokens;
}

bool endsWith(const std::string &value, const std::string &ending) {
    if (ending.size() > value.size()) return false;
    return std::equal(ending.rbegin(), ending.rend(), value.rbegin());
}

std::string timeString(double time, bool precise) {                   // ref: src/common.cpp:119-144
    if (std::isnan(time) || std::isinf(time)) return "inf";
    std::string suffix = "ms";
    if (time > 1000) { time /= 1000; suffix = "s";
        if (time > 60) { time /= 60; suffix = "m";
            if (time > 60) { time /= 60; suffix = "h";
                if (time > 12) { time /= 12; suffix = "d"; } } } }
    std::ostringstream os;
    os << std::setprecision(precise ? 4 : 1) << std::fixed << time << suffix;
    return os.str();
}

std::string memString(size_t size, bool precise) {                    // ref: src/common.cpp:146-157
    double value = (double) size;
    const char *suffixes[] = { "B", "KiB", "MiB", "GiB", "TiB", "PiB" };
    int suffix = 0;
    while (suffix < 5 && value > 1024.0f) { value /= 1024.0f; ++suffix; }
    std::ostringstream os;
    os << std::setprecision(suffix == 0 ? 0 : (precise ? 4 : 1)) << std::fixed << value << " " << suffixes[suffix];
    return os.str();
}

// x^(1/2.4) as exp(log(x) / 2.4) with the single-precision Cephes polynomials the device code uses (nb_device.cuh det_logf /
// det_expf; every multiply and add rounds separately on both sides), so that the host tonemap and film_to_srgb8_kernel
// produce the SAME 8-bit image -- std::pow and the device pow differ in the last ulp, which flips a byte now and then.
static float srgbLog(float xin) {
    uint32_t b; std::memcpy(&b, &xin, 4);
    int e = (int) ((b >> 23) & 0xff) - 126;
    uint32_t mb = (b & 0x007fffffu) | 0x3f000000u;
    float x; std::memcpy(&x, &mb, 4);
    if (x < 0.70710678118654752440f) { e -= 1; x = x + x - 1.0f; } else { x = x - 1.0f; }
    float z = x * x;
    float y = ((((((((7.0376836292e-2f * x - 1.1514610310e-1f) * x + 1.1676998740e-1f) * x - 1.2420140846e-1f) * x
              + 1.4249322787e-1f) * x - 1.6668057665e-1f) * x + 2.0000714765e-1f) * x - 2.4999993993e-1f) * x
              + 3.3333331174e-1f) * x * z;
    float fe = (float) e;
    y = y + (-2.12194440e-4f * fe);
    y = y + (-0.5f * z);
    z = x + y;
    z = z + 0.693359375f * fe;
    return z;
}
static float srgbExp(float x) {
    if (x < -87.0f) return 0.0f;
    if (x > 88.0f) x = 88.0f;
    float z = std::floor(1.44269504088896341f * x + 0.5f);
    x = x - z * 0.693359375f;
    x = x - z * -2.12194440e-4f;
    int n = (int) z;
    z = x * x;
    z = (((((1.9875691500e-4f * x + 1.3981999507e-3f) * x + 8.3334519073e-3f) * x + 4.1665795894e-2f) * x
        + 1.6666665459e-1f) * x + 5.0000001201e-1f) * z + x + 1.0f;
    uint32_t sb = (uint32_t) (n + 127) << 23; float sc; std::memcpy(&sc, &sb, 4);
    return z * sc;
}

Color3f Color3f::toSRGB() const {                                     // ref: src/common.cpp:166-180
    Color3f result;
    for (int i = 0; i < 3; ++i) {
        float value = c[i];
        if (value <= 0.0031308f) result.c[i] = 12.92f * value;
        else result.c[i] = (1.0f + 0.055f) * srgbExp(srgbLog(value) * (1.0f / 2.4f)) - 0.055f;
    }
    return result;
}

// ---------------------------------------------------------------- Matrix4f / Transform
Matrix4f Matrix4f::operator*(const Matrix4f &o) const {
    Matrix4f r;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) {
        double a = 0; for (int k = 0; k < 4; ++k) a += (double) m[i][k] * (double) o.m[k][j];
        r.m[i][j] = (float) a;
    }
    return r;
}

Matrix4f Matrix4f::inverse() const {
    double A[4][8];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { A[i][j] = m[i][j]; A[i][4 + j] = (i == j); }
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        for (int r = col + 1; r < 4; ++r) if (std::fabs(A[r][col]) > std::fabs(A[piv][col])) piv = r;
        if (A[piv][col] == 0.0) throw NoriException("Matrix4f::inverse(): singular matrix");
        if (piv != col) for (int j = 0; j < 8; ++j) std::swap(A[col][j], A[piv][j]);
        double d = A[col][col];
        for (int j = 0; j < 8; ++j) A[col][j] /= d;
        for (int r = 0; r < 4; ++r) if (r != col) { double f = A[r][col]; for (int j = 0; j < 8; ++j) A[r][j] -= f * A[col][j]; }
    }
    Matrix4f r;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r.m[i][j] = (float) A[i][4 + j];
    return r;
}

std::string Matrix4f::toString() const {
    std::ostringstream oss;
    oss << "[";
    for (int i = 0; i < 4; ++i) {
        for (int j = 0; j < 4; ++j) { oss << std::setprecision(4) << m[i][j]; if (j < 3) oss << ", "; }
        if (i < 3) oss << ";\n";
    }
    oss << "]";
    return oss.str();
}

Vector3f Transform::applyVector(const Vector3f &v) const {
    const Matrix4f &M = m_transform;
    return Vector3f(M(0, 0) * v[0] + (M(0, 1) * v[1] + M(0, 2) * v[2]), M(1, 0) * v[0] + (M(1, 1) * v[1] + M(1, 2) * v[2]),
                    M(2, 0) * v[0] + (M(2, 1) * v[1] + M(2, 2) * v[2]));
}

Normal3f Transform::applyNormal(const Normal3f &n) const {
    const Matrix4f &I = m_inverse;   // inverse transpose
    return Normal3f(I(0, 0) * n[0] + (I(1, 0) * n[1] + I(2, 0) * n[2]), I(0, 1) * n[0] + (I(1, 1) * n[1] + I(2, 1) * n[2]),
                    I(0, 2) * n[0] + (I(1, 2) * n[1] + I(2, 2) * n[2]));
}

Point3f Transform::applyPoint(const Point3f &p) const {
    const Matrix4f &M = m_transform;
    float r[4];
    for (int i = 0; i < 4; ++i) r[i] = ((M(i, 0) * p[0] + M(i, 1) * p[1]) + M(i, 2) * p[2]) + M(i, 3) * 1.0f;
    return Point3f(r[0] / r[3], r[1] / r[3], r[2] / r[3]);
}

// ---------------------------------------------------------------- NoriObject / factory (ref: src/object.cpp:11-26)
void NoriObject::addChild(NoriObject *) {
    throw NoriException("NoriObject::addChild() is not implemented for objects of type '%s'!", classTypeName(getClassType()));
}
void NoriObject::activate() { /* Do nothing */ }
void NoriObject::setParent(NoriObject *) { /* Do nothing */ }

std::map<std::string, NoriObjectFactory::Constructor> *NoriObjectFactory::m_constructors = nullptr;
std::map<const NoriObject *, NoriObjectFactory::Record> *NoriObjectFactory::m_records = nullptr;

void NoriObjectFactory::registerClass(const std::string &name, const Constructor &constr) {
    if (!m_constructors) m_constructors = new std::map<std::string, NoriObjectFactory::Constructor>();
    (*m_constructors)[name] = constr;
}

bool NoriObjectFactory::isRegistered(const std::string &name) {
    return m_constructors && m_constructors->find(name) != m_constructors->end();
}

NoriObject *NoriObjectFactory::createInstance(const std::string &name, const PropertyList &propList) {   // ref: include/nori/object.h:130-135
    if (!m_constructors || m_constructors->find(name) == m_constructors->end())
        throw NoriException("A constructor for class \"%s\" could not be found!", name);
    NoriObject *obj = (*m_constructors)[name](propList);
    if (!m_records) m_records = new std::map<const NoriObject *, Record>();
    (*m_records)[obj] = Record{ name, propList };
    return obj;
}

const NoriObjectFactory::Record *NoriObjectFactory::creationRecord(const NoriObject *obj) {
    if (!m_records) return nullptr;
    auto it = m_records->find(obj);
    return it == m_records->end() ? nullptr : &it->second;
}

void NoriObjectFactory::forgetRecord(const NoriObject *obj) { if (m_records) m_records->erase(obj); }

// ---------------------------------------------------------------- PropertyList (ref: src/proplist.cpp:10-48)
#define DEFINE_PROPERTY_ACCESSOR(Type, TypeName, XmlName) \
    void PropertyList::set##TypeName(const std::string &name, const Type &value) { \
        if (m_properties.find(name) != m_properties.end()) \
            cerr << "Property \"" << name <<  "\" was specified multiple times!" << endl; \
        auto &prop = m_properties[name]; \
        prop.value.XmlName##_value = value; \
        prop.type = Property::XmlName##_type; \
    } \
    Type PropertyList::get##TypeName(const std::string &name) const { \
        auto it = m_properties.find(name); \
        if (it == m_properties.end()) \
            throw NoriException("Property '%s' is missing!", name); \
        if (it->second.type != Property::XmlName##_type) \
            throw NoriException("Property '%s' has the wrong type! (expected <" #XmlName ">)!", name); \
        return it->second.value.XmlName##_value; \
    } \
    Type PropertyList::get##TypeName(const std::string &name, const Type &defVal) const { \
        auto it = m_properties.find(name); \
        if (it == m_properties.end()) \
            return defVal; \
        if (it->second.type != Property::XmlName##_type) \
            throw NoriException("Property '%s' has the wrong type! (expected <" #XmlName ">)!", name); \
        return it->second.value.XmlName##_value; \
    }

DEFINE_PROPERTY_ACCESSOR(bool, Boolean, boolean)
DEFINE_PROPERTY_ACCESSOR(int, Integer, integer)
DEFINE_PROPERTY_ACCESSOR(float, Float, float)
DEFINE_PROPERTY_ACCESSOR(Color3f, Color, color)
DEFINE_PROPERTY_ACCESSOR(Point3f, Point, point)
DEFINE_PROPERTY_ACCESSOR(Vector3f, Vector, vector)
DEFINE_PROPERTY_ACCESSOR(std::string, String, string)
DEFINE_PROPERTY_ACCESSOR(Transform, Transform, transform)

NORI_NAMESPACE_END

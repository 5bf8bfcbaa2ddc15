// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// plugins.cpp -- the host-side plugin classes of the hot path: BSDFs, area emitter, integrators, sampler, camera,
// reconstruction filters.  Each registers under the name the reference's scenes use (NORI_REGISTER_CLASS).
// The host objects carry PARAMETERS; radiance is evaluated by the CUDA kernels (nori_b200/csrc/nb_kernels.cuh), which
// receive plain-data descriptors built from each object's creation record (render.cpp).  Entry points whose bodies
// the reference ships as stubs that throw (ref: src/microfacet.cpp:40-52, src/dielectric.cpp:33-35) throw here too.
#include "nori/block.h"
#include "nori/scene.h"

NORI_NAMESPACE_BEGIN

// ------------------------------------------------------------------ BSDFs
/// Diffuse / Lambertian BRDF (ref: src/diffuse.cpp:16-88)
class Diffuse : public BSDF {
public:
    Diffuse(const PropertyList &propList) { m_albedo = propList.getColor("albedo", Color3f(0.5f)); }
    Color3f eval(const BSDFQueryRecord &bRec) const {
        if (bRec.measure != ESolidAngle || bRec.wi.z() <= 0 || bRec.wo.z() <= 0) return Color3f(0.0f);
        return Color3f(m_albedo.r() * INV_PI, m_albedo.g() * INV_PI, m_albedo.b() * INV_PI);
    }
    float pdf(const BSDFQueryRecord &bRec) const {
        if (bRec.measure != ESolidAngle || bRec.wi.z() <= 0 || bRec.wo.z() <= 0) return 0.0f;
        return INV_PI * bRec.wo.z();
    }
    Color3f sample(BSDFQueryRecord &, const Point2f &) const {
        throw NoriException("Diffuse::sample(): sampling runs on the device (Warp::squareToCosineHemisphere is a stub in the reference too, src/warp.cpp:53-55)");
    }
    bool isDiffuse() const { return true; }
    std::string toString() const { return format("Diffuse[\n  albedo = %s\n]", m_albedo.toString()); }
private:
    Color3f m_albedo;
};

/// Ideal mirror BRDF (ref: src/mirror.cpp:12-48)
class Mirror : public BSDF {
public:
    Mirror(const PropertyList &) { }
    Color3f eval(const BSDFQueryRecord &) const { return Color3f(0.0f); }
    float pdf(const BSDFQueryRecord &) const { return 0.0f; }
    Color3f sample(BSDFQueryRecord &bRec, const Point2f &) const {
        if (bRec.wi.z() <= 0) return Color3f(0.0f);
        bRec.wo = Vector3f(-bRec.wi.x(), -bRec.wi.y(), bRec.wi.z());
        bRec.measure = EDiscrete;
        bRec.eta = 1.0f;
        return Color3f(1.0f);
    }
    std::string toString() const { return "Mirror[]"; }
};

/// Ideal dielectric BSDF (ref: src/dielectric.cpp:13-47)
class Dielectric : public BSDF {
public:
    Dielectric(const PropertyList &propList) {
        m_intIOR = propList.getFloat("intIOR", 1.5046f);
        m_extIOR = propList.getFloat("extIOR", 1.000277f);
    }
    Color3f eval(const BSDFQueryRecord &) const { return Color3f(0.0f); }
    float pdf(const BSDFQueryRecord &) const { return 0.0f; }
    Color3f sample(BSDFQueryRecord &, const Point2f &) const { throw NoriException("Dielectric::sample(): evaluated on the device"); }
    std::string toString() const { return format("Dielectric[\n  intIOR = %f,\n  extIOR = %f\n]", m_intIOR, m_extIOR); }
private:
    float m_intIOR, m_extIOR;
};

/// Rough conductor-over-diffuse microfacet BRDF (ref: src/microfacet.cpp:12-88)
class Microfacet : public BSDF {
public:
    Microfacet(const PropertyList &propList) {
        m_alpha = propList.getFloat("alpha", 0.1f);
        m_intIOR = propList.getFloat("intIOR", 1.5046f);
        m_extIOR = propList.getFloat("extIOR", 1.000277f);
        m_kd = propList.getColor("kd", Color3f(0.5f));
        m_ks = 1 - m_kd.maxCoeff();
    }
    Color3f eval(const BSDFQueryRecord &) const { throw NoriException("MicrofacetBRDF::eval(): evaluated on the device"); }
    float pdf(const BSDFQueryRecord &) const { throw NoriException("MicrofacetBRDF::pdf(): evaluated on the device"); }
    Color3f sample(BSDFQueryRecord &, const Point2f &) const { throw NoriException("MicrofacetBRDF::sample(): evaluated on the device"); }
    bool isDiffuse() const { return true; }
    std::string toString() const {
        return format("Microfacet[\n  alpha = %f,\n  intIOR = %f,\n  extIOR = %f,\n  kd = %s,\n  ks = %f\n]",
                      m_alpha, m_intIOR, m_extIOR, m_kd.toString(), m_ks);
    }
private:
    float m_alpha, m_intIOR, m_extIOR, m_ks;
    Color3f m_kd;
};

NORI_REGISTER_CLASS(Diffuse, "diffuse");
NORI_REGISTER_CLASS(Mirror, "mirror");
NORI_REGISTER_CLASS(Dielectric, "dielectric");
NORI_REGISTER_CLASS(Microfacet, "microfacet");

// ------------------------------------------------------------------ emitter
/// Area light attached to a mesh (named by the shipped scenes, e.g. ref: scenes/pa4/cbox/cbox-distributed.xml:59-61)
class AreaLight : public Emitter {
public:
    AreaLight(const PropertyList &propList) { m_radiance = propList.getColor("radiance"); }
    const Color3f &getRadiance() const { return m_radiance; }
    std::string toString() const { return format("AreaLight[\n  radiance = %s\n]", m_radiance.toString()); }
private:
    Color3f m_radiance;
};
NORI_REGISTER_CLASS(AreaLight, "area");

// ------------------------------------------------------------------ integrators
/// Integrators are selected by name; Li() runs on the device for all of them.
#define NORI_DEVICE_INTEGRATOR(Cls, xmlName) \
    class Cls : public Integrator { \
    public: \
        Cls(const PropertyList &) { } \
        Color3f Li(const Scene *, Sampler *, const Ray3f &) const { \
            throw NoriException(#Cls "::Li(): this integrator is evaluated by the CUDA render path (nb_render)"); \
        } \
        std::string toString() const { return #Cls "[]"; } \
    }; \
    NORI_REGISTER_CLASS(Cls, xmlName)

NORI_DEVICE_INTEGRATOR(NormalIntegrator, "normals");     // ref: scenes/pa1/bunny.xml:8
NORI_DEVICE_INTEGRATOR(AOIntegrator, "ao");              // ref: scenes/pa3/ajax-ao.xml:8
NORI_DEVICE_INTEGRATOR(WhittedIntegrator, "whitted");    // ref: scenes/pa4/cbox/cbox-whitted.xml:4
NORI_DEVICE_INTEGRATOR(PathMatsIntegrator, "path_mats"); // ref: scenes/pa5/cbox/cbox_mats.xml:4
NORI_DEVICE_INTEGRATOR(PathEmsIntegrator, "path_ems");   // ref: scenes/pa5/cbox/cbox_ems.xml:4
NORI_DEVICE_INTEGRATOR(PathMisIntegrator, "path_mis");   // ref: scenes/pa5/cbox/cbox_mis.xml:4

/// Point-light integrator of ref: scenes/pa3/ajax-simple.xml:8-11.  Both properties are mandatory (a missing one
/// throws from PropertyList, as any plugin constructor of the reference does); Li() runs on the device.
class SimpleIntegrator : public Integrator {
public:
    SimpleIntegrator(const PropertyList &props) {
        m_position = props.getPoint("position");
        m_energy = props.getColor("energy");
    }
    Color3f Li(const Scene *, Sampler *, const Ray3f &) const {
        throw NoriException("SimpleIntegrator::Li(): this integrator is evaluated by the CUDA render path (nb_render)");
    }
    std::string toString() const {
        return format("SimpleIntegrator[\n  position = %s,\n  energy = %s\n]", m_position.toString(), m_energy.toString());
    }
private:
    Point3f m_position;
    Color3f m_energy;
};
NORI_REGISTER_CLASS(SimpleIntegrator, "simple");

// ------------------------------------------------------------------ sampler
/// Independent sampling (ref: src/independent.cpp:21-65).  "seedMode" = "sample" (default; one pcg32 stream per
/// (pixel, sample), the parallel mode) or "block" (the reference's Independent::prepare: one stream per 32x32 block).
class Independent : public Sampler {
public:
    Independent(const PropertyList &propList) {
        m_sampleCount = (size_t) propList.getInteger("sampleCount", 1);
        std::string mode = propList.getString("seedMode", "sample");
        if (mode != "sample" && mode != "block") throw NoriException("Independent: unknown seedMode \"%s\"", mode);
        m_blockMode = mode == "block";
        m_seed = (uint64_t) propList.getInteger("seed", 0);
    }
    std::unique_ptr<Sampler> clone() const {
        std::unique_ptr<Independent> cloned(new Independent());
        cloned->m_sampleCount = m_sampleCount; cloned->m_random = m_random; cloned->m_blockMode = m_blockMode; cloned->m_seed = m_seed;
        return std::unique_ptr<Sampler>(cloned.release());
    }
    void prepare(const ImageBlock &block) { m_random.seed((uint64_t) block.getOffset().x(), (uint64_t) block.getOffset().y()); }
    void generate() { }
    void advance() { }
    float next1D() { return m_random.nextFloat(); }
    Point2f next2D() { float a = m_random.nextFloat(), b = m_random.nextFloat(); return Point2f(a, b); }
    bool blockMode() const { return m_blockMode; }
    uint64_t seed() const { return m_seed; }
    std::string toString() const { return format("Independent[sampleCount=%i]", (int) m_sampleCount); }
protected:
    Independent() { }
private:
    pcg32 m_random;
    bool m_blockMode = false;
    uint64_t m_seed = 0;
};
NORI_REGISTER_CLASS(Independent, "independent");

// ------------------------------------------------------------------ reconstruction filters (ref: src/rfilter.cpp:16-108)
class GaussianFilter : public ReconstructionFilter {
public:
    GaussianFilter(const PropertyList &propList) { m_radius = propList.getFloat("radius", 2.0f); m_stddev = propList.getFloat("stddev", 0.5f); }
    float eval(float x) const {
        float alpha = -1.0f / (2.0f * m_stddev * m_stddev);
        return std::max(0.0f, std::exp(alpha * x * x) - std::exp(alpha * m_radius * m_radius));
    }
    std::string toString() const { return format("GaussianFilter[radius=%f, stddev=%f]", m_radius, m_stddev); }
protected:
    float m_stddev;
};

class MitchellNetravaliFilter : public ReconstructionFilter {
public:
    MitchellNetravaliFilter(const PropertyList &propList) {
        m_radius = propList.getFloat("radius", 2.0f); m_B = propList.getFloat("B", 1.0f / 3.0f); m_C = propList.getFloat("C", 1.0f / 3.0f);
    }
    float eval(float x) const {
        x = std::abs(2.0f * x / m_radius);
        float x2 = x * x, x3 = x2 * x;
        if (x < 1) return 1.0f / 6.0f * ((12 - 9 * m_B - 6 * m_C) * x3 + (-18 + 12 * m_B + 6 * m_C) * x2 + (6 - 2 * m_B));
        else if (x < 2) return 1.0f / 6.0f * ((-m_B - 6 * m_C) * x3 + (6 * m_B + 30 * m_C) * x2 + (-12 * m_B - 48 * m_C) * x + (8 * m_B + 24 * m_C));
        else return 0.0f;
    }
    std::string toString() const { return format("MitchellNetravaliFilter[radius=%f, B=%f, C=%f]", m_radius, m_B, m_C); }
protected:
    float m_B, m_C;
};

class TentFilter : public ReconstructionFilter {
public:
    TentFilter(const PropertyList &) { m_radius = 1.0f; }
    float eval(float x) const { return std::max(0.0f, 1.0f - std::abs(x)); }
    std::string toString() const { return "TentFilter[]"; }
};

class BoxFilter : public ReconstructionFilter {
public:
    BoxFilter(const PropertyList &) { m_radius = 0.5f; }
    float eval(float) const { return 1.0f; }
    std::string toString() const { return "BoxFilter[]"; }
};

NORI_REGISTER_CLASS(GaussianFilter, "gaussian");
NORI_REGISTER_CLASS(MitchellNetravaliFilter, "mitchell");
NORI_REGISTER_CLASS(TentFilter, "tent");
NORI_REGISTER_CLASS(BoxFilter, "box");

// ------------------------------------------------------------------ camera (ref: src/perspective.cpp:20-138)
class PerspectiveCamera : public Camera {
public:
    PerspectiveCamera(const PropertyList &propList) {
        m_outputSize.x() = propList.getInteger("width", 1280);
        m_outputSize.y() = propList.getInteger("height", 720);
        m_invOutputSize = Point2f(1.0f / (float) m_outputSize.x(), 1.0f / (float) m_outputSize.y());
        m_cameraToWorld = propList.getTransform("toWorld", Transform());
        m_fov = propList.getFloat("fov", 30.0f);
        m_nearClip = propList.getFloat("nearClip", 1e-4f);
        m_farClip = propList.getFloat("farClip", 1e4f);
        m_rfilter = NULL;
    }
    ~PerspectiveCamera() { delete m_rfilter; }

    void activate() {
        /* sampleToCamera = inverse(scale * translate * perspective), ref: src/perspective.cpp:41-68; composed and
           inverted in double, rounded once to fp32 */
        double aspect = (double) (m_outputSize.x() / (float) m_outputSize.y());
        double recip = 1.0 / ((double) m_farClip - (double) m_nearClip);
        double cot = 1.0 / std::tan((double) m_fov / 2.0 * (3.14159265358979323846 / 180.0));
        double P[4][4] = { { cot, 0, 0, 0 }, { 0, cot, 0, 0 }, { 0, 0, m_farClip * recip, -(double) m_nearClip * m_farClip * recip }, { 0, 0, 1, 0 } };
        double T[4][4] = { { 1, 0, 0, -1 }, { 0, 1, 0, -1.0 / aspect }, { 0, 0, 1, 0 }, { 0, 0, 0, 1 } };
        double S[4][4] = { { -0.5, 0, 0, 0 }, { 0, -0.5 * aspect, 0, 0 }, { 0, 0, 1, 0 }, { 0, 0, 0, 1 } };
        double TP[4][4], M[4][8];
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double a = 0; for (int k = 0; k < 4; ++k) a += T[i][k] * P[k][j]; TP[i][j] = a; }
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { double a = 0; for (int k = 0; k < 4; ++k) a += S[i][k] * TP[k][j]; M[i][j] = a; M[i][4 + j] = (i == j); }
        for (int col = 0; col < 4; ++col) {
            int piv = col;
            for (int r = col + 1; r < 4; ++r) if (std::fabs(M[r][col]) > std::fabs(M[piv][col])) piv = r;
            if (piv != col) for (int j = 0; j < 8; ++j) std::swap(M[col][j], M[piv][j]);
            double d = M[col][col];
            for (int j = 0; j < 8; ++j) M[col][j] /= d;
            for (int r = 0; r < 4; ++r) if (r != col) { double f = M[r][col]; for (int j = 0; j < 8; ++j) M[r][j] -= f * M[col][j]; }
        }
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) m_sampleToCamera(i, j) = (float) M[i][4 + j];
        if (!m_rfilter) /* default: Gaussian (ref: src/perspective.cpp:70-73) */
            m_rfilter = static_cast<ReconstructionFilter *>(NoriObjectFactory::createInstance("gaussian", PropertyList()));
    }

    Color3f sampleRay(Ray3f &ray, const Point2f &samplePosition, const Point2f &) const {   // ref: src/perspective.cpp:76-97
        Transform s2c(m_sampleToCamera, Matrix4f());
        Point3f nearP = s2c.applyPoint(Point3f(samplePosition.x * m_invOutputSize.x, samplePosition.y * m_invOutputSize.y, 0.0f));
        Vector3f d = nearP.normalized();
        float invZ = 1.0f / d.z();
        ray.o = m_cameraToWorld.applyPoint(Point3f(0, 0, 0));
        ray.d = m_cameraToWorld.applyVector(d);
        ray.mint = m_nearClip * invZ;
        ray.maxt = m_farClip * invZ;
        ray.update();
        return Color3f(1.0f);
    }

    void addChild(NoriObject *obj) {
        switch (obj->getClassType()) {
            case EReconstructionFilter:
                if (m_rfilter) throw NoriException("Camera: tried to register multiple reconstruction filters!");
                m_rfilter = static_cast<ReconstructionFilter *>(obj);
                break;
            default:
                throw NoriException("Camera::addChild(<%s>) is not supported!", classTypeName(obj->getClassType()));
        }
    }
    const Matrix4f &sampleToCamera() const { return m_sampleToCamera; }
    const Transform &cameraToWorld() const { return m_cameraToWorld; }
    float nearClip() const { return m_nearClip; }
    float farClip() const { return m_farClip; }
    std::string toString() const {
        return format("PerspectiveCamera[\n  cameraToWorld = %s,\n  outputSize = %s,\n  fov = %f,\n  clip = [%f, %f],\n  rfilter = %s\n]",
                      indent(m_cameraToWorld.toString(), 18), m_outputSize.toString(), m_fov, m_nearClip, m_farClip, indent(m_rfilter->toString()));
    }
private:
    Point2f m_invOutputSize;
    Matrix4f m_sampleToCamera;
    Transform m_cameraToWorld;
    float m_fov, m_nearClip, m_farClip;
};
NORI_REGISTER_CLASS(PerspectiveCamera, "perspective");

// accessors used by render.cpp without exposing the plugin classes in headers
bool nori_camera_matrices(const Camera *cam, float s2c[16], float c2w[16], float *nearClip, float *farClip) {
    const PerspectiveCamera *p = dynamic_cast<const PerspectiveCamera *>(cam);
    if (!p) return false;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) { s2c[4 * i + j] = p->sampleToCamera()(i, j); c2w[4 * i + j] = p->cameraToWorld().getMatrix()(i, j); }
    *nearClip = p->nearClip(); *farClip = p->farClip();
    return true;
}
bool nori_sampler_mode(const Sampler *s, bool *blockMode, uint64_t *seed) {
    const Independent *p = dynamic_cast<const Independent *>(s);
    if (!p) return false;
    *blockMode = p->blockMode(); *seed = p->seed();
    return true;
}

NORI_NAMESPACE_END

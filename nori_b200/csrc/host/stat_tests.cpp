// stat_tests.cpp -- the `ttest` scene object (<test type="ttest">), scene mode, on top of nb_li_samples.
//
// Interface and decisions follow ref: src/ttest.cpp:47-189 -- properties `significanceLevel` (0.01), `references`,
// `angles`, `sampleCount` (100000); children are scenes (or BSDFs); the test runs inside activate(), i.e. while the
// XML file is being loaded, prints one verdict per scene and throws if any of them is rejected.  Differences:
//   * Li() of the sampleCount paths is evaluated by the device (nb_li_samples) with one pcg32 stream per path; the
//     reference threads ONE sampler stream through all paths (ttest.cpp:141-167), which has no parallel form.  The
//     test statistic does not depend on that choice.
//   * BSDF mode (ttest.cpp:95-137) calls BSDF::sample() on the host; the BSDF classes of this mirror are parameter
//     holders whose bodies live in the CUDA path, so BSDF children are refused with an error.  The same five
//     microfacet values are checked against oracle and device in tests/test_oracle_fixtures.py.
// The p-value arithmetic restates what the reference gets from the un-vendored `hypothesis` library
// (students_t_test: two-sided Student t with n-1 degrees of freedom, Sidak-corrected level), as tests/fixtures.py does.
#include <cmath>
#include "nori_b200.h"
#include "nori/block.h"
#include "nori/plugins.h"
#include "nori/render.h"

NORI_NAMESPACE_BEGIN

namespace stats {

// continued fraction of the incomplete beta function (modified Lentz)
static double betaFraction(double a, double b, double x) {
    const double tiny = 1e-300, eps = 1e-15;
    double c = 1.0, d = 1.0 - (a + b) * x / (a + 1.0);
    if (std::fabs(d) < tiny) d = tiny;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 10000; ++m) {
        const double m2 = 2.0 * m;
        double num = m * (b - m) * x / ((a + m2 - 1.0) * (a + m2));
        d = 1.0 + num * d; if (std::fabs(d) < tiny) d = tiny;
        c = 1.0 + num / c; if (std::fabs(c) < tiny) c = tiny;
        d = 1.0 / d; h *= d * c;
        num = -(a + m) * (a + b + m) * x / ((a + m2) * (a + m2 + 1.0));
        d = 1.0 + num * d; if (std::fabs(d) < tiny) d = tiny;
        c = 1.0 + num / c; if (std::fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double delta = d * c;
        h *= delta;
        if (std::fabs(delta - 1.0) < eps) break;
    }
    return h;
}

/// Regularised incomplete beta function I_x(a, b)
double incompleteBeta(double a, double b, double x) {
    if (x <= 0.0) return 0.0;
    if (x >= 1.0) return 1.0;
    const double front = std::exp(std::lgamma(a + b) - std::lgamma(a) - std::lgamma(b) + a * std::log(x) + b * std::log1p(-x));
    if (x < (a + 1.0) / (a + b + 2.0)) return front * betaFraction(a, b, x) / a;
    return 1.0 - front * betaFraction(b, a, 1.0 - x) / b;
}

/// Two-sided p-value of Student's t statistic with `dof` degrees of freedom: P(|T| >= |t|)
double studentsTwoSided(double t, double dof) {
    return incompleteBeta(0.5 * dof, 0.5, dof / (dof + t * t));
}

/// Sidak-corrected significance level for `numTests` simultaneous tests
double sidak(double level, int numTests) { return 1.0 - std::pow(1.0 - level, 1.0 / (double) numTests); }

struct Verdict { bool accepted; double tStat, pValue, level; };

Verdict studentsTTest(double mean, double variance, double reference, int sampleCount, double level, int numTests) {
    Verdict v;
    const double sd = std::max(std::sqrt(variance), 1e-5);       // a noise-free estimator would divide by zero
    v.tStat = std::fabs(mean - reference) * std::sqrt((double) sampleCount) / sd;
    v.pValue = studentsTwoSided(v.tStat, (double) (sampleCount - 1));
    v.level = sidak(level, numTests);
    v.accepted = v.pValue > v.level;
    return v;
}

}  // namespace stats

class StudentsTTest : public NoriObject {
public:
    StudentsTTest(const PropertyList &props) {
        m_level = props.getFloat("significanceLevel", 0.01f);
        for (const std::string &tok : tokenize(props.getString("angles", ""))) m_angles.push_back(toFloat(tok));
        for (const std::string &tok : tokenize(props.getString("references", ""))) m_references.push_back(toFloat(tok));
        m_sampleCount = props.getInteger("sampleCount", 100000);
        if (m_sampleCount < 2) throw NoriException("StudentsTTest: sampleCount must be at least 2");
    }

    ~StudentsTTest() {
        for (Scene *s : m_scenes) delete s;
        for (NoriObject *b : m_bsdfs) delete b;
    }

    void addChild(NoriObject *obj) {
        if (obj->getClassType() == EScene) m_scenes.push_back(static_cast<Scene *>(obj));
        else if (obj->getClassType() == EBSDF) m_bsdfs.push_back(obj);
        else throw NoriException("StudentsTTest::addChild(<%s>) is not supported!", classTypeName(obj->getClassType()));
    }

    void activate() {
        if (!m_bsdfs.empty())
            throw NoriException("StudentsTTest: BSDF mode has no device implementation (BSDF::sample runs inside the CUDA path only); "
                                "scene mode is supported");
        if (m_references.size() != m_scenes.size())
            throw NoriException("Specified a different number of scenes and reference values!");
        int passed = 0;
        for (size_t i = 0; i < m_scenes.size(); ++i) {
            Scene *scene = m_scenes[i];
            cout << "------------------------------------------------------" << endl;
            cout << "Testing scene: " << scene->toString() << endl;
            cout << "Generating " << m_sampleCount << " paths.. " << endl;
            std::vector<float> lum = pathLuminances(scene);
            // online mean / variance in path order (Knuth, TAOCP vol. 2, as ref: src/ttest.cpp:159-166)
            double mean = 0, m2 = 0;
            for (int k = 0; k < m_sampleCount; ++k) {
                const double x = (double) lum[k], delta = x - mean;
                mean += delta / (double) (k + 1);
                m2 += delta * (x - mean);
            }
            const double variance = m2 / (double) (m_sampleCount - 1);
            const stats::Verdict v = stats::studentsTTest(mean, variance, m_references[i], m_sampleCount, m_level, (int) m_references.size());
            cout << format("Sample mean = %f (reference value = %f), sample variance = %g, t-statistic = %f, p-value = %f, "
                           "significance level = %f: %s", mean, m_references[i], variance, v.tStat, v.pValue, v.level,
                           v.accepted ? "accepted the null hypothesis" : "REJECTED the null hypothesis") << endl;
            if (v.accepted) ++passed;
        }
        cout << "Passed " << passed << "/" << m_scenes.size() << " tests." << endl;
        if (passed < (int) m_scenes.size()) throw NoriException("Some tests failed :(");
    }

    std::string toString() const {
        return format("StudentsTTest[\n  significanceLevel = %f,\n  sampleCount= %i\n]", m_level, m_sampleCount);
    }

    EClassType getClassType() const { return ETest; }

private:
    std::vector<float> pathLuminances(Scene *scene) const {
        scene->getIntegrator()->preprocess(scene);
        const Camera *camera = scene->getCamera();
        ImageBlock film(camera->getOutputSize(), camera->getReconstructionFilter());
        RenderOptions opt; opt.quiet = true;
        nb_ctx *ctx = createDeviceScene(scene, film, opt);
        std::vector<float> lum((size_t) m_sampleCount);
        const int rc = nb_li_samples(ctx, (uint64_t) m_sampleCount, lum.data(), nullptr);
        const std::string err = rc ? nb_last_error() : "";
        nb_destroy(ctx);
        if (rc) throw NoriException("nb_li_samples: %s", err);
        return lum;
    }

    float m_level;
    int m_sampleCount;
    std::vector<float> m_angles, m_references;
    std::vector<Scene *> m_scenes;
    std::vector<NoriObject *> m_bsdfs;
};
NORI_REGISTER_CLASS(StudentsTTest, "ttest");

NORI_NAMESPACE_END

extern "C" {
/// Exposed so that the CPU tests can pin the p-value arithmetic against scipy (no GPU involved).
double nori_host_students_t_pvalue(double t, double dof) { return nori::stats::studentsTwoSided(t, dof); }
int nori_host_students_t_test(double mean, double variance, double reference, int n, double level, int ntests, double *pvalue) {
    const nori::stats::Verdict v = nori::stats::studentsTTest(mean, variance, reference, n, level, ntests);
    if (pvalue) *pvalue = v.pValue;
    return v.accepted ? 1 : 0;
}
}

// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// stat_tests.cpp -- the statistical test objects of the scene grammar, <test type="ttest"> and <test type="chi2test">,
// with every BSDF / integrator evaluation done by the device through the C-ABI.
//
// Interfaces and decisions follow ref: src/ttest.cpp:47-189 and src/chi2test.cpp:31-214 -- same properties and
// defaults, same children, the test runs inside activate() (i.e. while the XML file is being loaded), prints one
// verdict per case and throws if any case is rejected.  What differs, and why:
//   * t-test, scene mode: Li() of the sampleCount paths comes from nb_li_samples, one pcg32 stream per path; the
//     reference threads ONE sampler stream through all paths (ttest.cpp:141-167), which has no parallel form.
//   * t-test, BSDF mode and chi^2 test: the random numbers ARE the reference's (one default-constructed pcg32 consumed
//     in the same order, ttest.cpp:93,116 / chi2test.cpp:82,99-103,114); the batch of (wi, xi) goes through
//     nb_bsdf_sample in one call instead of sampleCount virtual calls.
//   * chi^2 test: the expected frequencies integrate pdf() over each (cos theta, phi) cell with a composite Simpson
//     rule refined by doubling until converged, every refinement level evaluated in one nb_bsdf_eval_pdf batch; the
//     reference uses hypothesis::adaptiveSimpson2D with per-point virtual calls (chi2test.cpp:131-151).  No
//     chi2test_*.m dump.
// The p-value arithmetic restates what the reference gets from the un-vendored `hypothesis` library (students_t_test,
// chi2_test with low-frequency pooling, Sidak-corrected level), as tests/fixtures.py does, and is pinned against scipy
// in tests/test_host_cpu.py.
#include <algorithm>
#include <cmath>
#include "nori_b200.h"
#include "nori/pcg32.h"
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

/// Regularised upper incomplete gamma function Q(a, x) (series below a + 1, continued fraction above)
double upperGamma(double a, double x) {
    if (x <= 0.0) return 1.0;
    const double lg = std::lgamma(a);
    if (x < a + 1.0) {
        double term = 1.0 / a, sum = term, ap = a;
        for (int n = 0; n < 100000; ++n) { ap += 1.0; term *= x / ap; sum += term; if (std::fabs(term) < std::fabs(sum) * 1e-16) break; }
        return 1.0 - sum * std::exp(-x + a * std::log(x) - lg);
    }
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 100000; ++i) {
        const double an = -i * (i - a);
        b += 2.0;
        d = an * d + b; if (std::fabs(d) < tiny) d = tiny;
        c = b + an / c; if (std::fabs(c) < tiny) c = tiny;
        d = 1.0 / d;
        const double delta = d * c;
        h *= delta;
        if (std::fabs(delta - 1.0) < 1e-16) break;
    }
    return std::exp(-x + a * std::log(x) - lg) * h;
}

struct Chi2Verdict { bool accepted; double statistic, pValue, level; int dof; };

/// Pearson's chi^2 test of observed against expected cell frequencies.  Cells are visited by increasing expectation;
/// cells whose expectation is below minExpFrequency are pooled until the pool reaches it (the normal approximation
/// behind the test needs that); a cell with zero expectation must be (almost) empty.
Chi2Verdict chi2Test(int nCells, const double *obs, const double *exp, int sampleCount, double minExpFrequency, double level, int numTests) {
    std::vector<int> order((size_t) nCells);
    for (int i = 0; i < nCells; ++i) order[(size_t) i] = i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return exp[a] < exp[b]; });
    Chi2Verdict v; v.statistic = 0; v.dof = 0; v.level = sidak(level, numTests); v.pValue = 0; v.accepted = false;
    double pooledObs = 0, pooledExp = 0;
    for (int idx : order) {
        if (exp[idx] == 0) {
            if (obs[idx] > sampleCount * 1e-5) return v;          // samples where the density says there are none
        } else if (exp[idx] < minExpFrequency || (pooledExp > 0 && pooledExp < minExpFrequency)) {
            pooledObs += obs[idx]; pooledExp += exp[idx];
        } else {
            const double diff = obs[idx] - exp[idx];
            v.statistic += diff * diff / exp[idx]; ++v.dof;
        }
    }
    if (pooledExp > 0) {
        const double diff = pooledObs - pooledExp;
        v.statistic += diff * diff / pooledExp; ++v.dof;
    }
    v.dof -= 1;
    if (v.dof <= 0) { v.pValue = 1.0; v.accepted = true; return v; }
    v.pValue = upperGamma(0.5 * v.dof, 0.5 * v.statistic);
    v.accepted = v.pValue > v.level;
    return v;
}

/// Online mean / variance in sample order (Knuth, TAOCP vol. 2; the recurrence of ref: src/ttest.cpp:119-124,159-166)
struct RunningMoments {
    double mean = 0, m2 = 0; long n = 0;
    void add(double x) { const double delta = x - mean; ++n; mean += delta / (double) n; m2 += delta * (x - mean); }
    double variance() const { return m2 / (double) (n - 1); }
};

}  // namespace stats

/// A device context without a scene, for the BSDF batch calls
struct DeviceHandle {
    nb_ctx *ctx;
    DeviceHandle() : ctx(nb_create(0)) { if (!ctx) throw NoriException("nb_create: %s", std::string(nb_last_error())); }
    ~DeviceHandle() { nb_destroy(ctx); }
};

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
        else if (obj->getClassType() == EBSDF) m_bsdfs.push_back(static_cast<BSDF *>(obj));
        else throw NoriException("StudentsTTest::addChild(<%s>) is not supported!", classTypeName(obj->getClassType()));
    }

    void activate() {
        if (!m_bsdfs.empty()) { activateBSDFs(); return; }
        if (m_references.size() != m_scenes.size())
            throw NoriException("Specified a different number of scenes and reference values!");
        int passed = 0;
        for (size_t i = 0; i < m_scenes.size(); ++i) {
            Scene *scene = m_scenes[i];
            cout << "------------------------------------------------------" << endl;
            cout << "Testing scene: " << scene->toString() << endl;
            cout << "Generating " << m_sampleCount << " paths.. " << endl;
            std::vector<float> lum = pathLuminances(scene);
            stats::RunningMoments mom;
            for (int k = 0; k < m_sampleCount; ++k) mom.add((double) lum[k]);
            if (verdict(mom, m_references[i])) ++passed;
        }
        cout << "Passed " << passed << "/" << m_scenes.size() << " tests." << endl;
        if (passed < (int) m_scenes.size()) throw NoriException("Some tests failed :(");
    }

    /// BSDF mode (ref: src/ttest.cpp:95-137): for every BSDF and incidence angle, the mean of sample()'s weight
    /// under uniform incident illumination against the reference value.
    void activateBSDFs() {
        if (m_references.size() != m_bsdfs.size() * m_angles.size())
            throw NoriException("Specified a different number of angles and reference values!");
        if (!m_scenes.empty()) throw NoriException("Cannot test BSDFs and scenes at the same time!");
        DeviceHandle dev;
        pcg32 random;                                   // ONE stream across all BSDFs and angles, as the reference
        int passed = 0, total = 0;
        size_t ctr = 0;
        std::vector<float> xi((size_t) m_sampleCount * 2), out((size_t) m_sampleCount * 8);
        for (BSDF *bsdf : m_bsdfs) {
            nb_bsdf_desc desc; describeBSDF(bsdf, &desc);
            for (float angle : m_angles) {
                const float reference = m_references[ctr++];
                cout << "------------------------------------------------------" << endl;
                cout << "Testing (angle=" << angle << "): " << bsdf->toString() << endl;
                ++total;
                const float theta = degToRad(angle);    // wi = sphericalDirection(theta, 0): ref src/common.cpp:224-236
                const float wi[3] = { std::sin(theta) * std::cos(0.0f), std::sin(theta) * std::sin(0.0f), std::cos(theta) };
                cout << "Drawing " << m_sampleCount << " samples .. " << endl;
                for (float &x : xi) x = random.nextFloat();
                if (nb_bsdf_sample(dev.ctx, &desc, wi, 0, xi.data(), (uint64_t) m_sampleCount, out.data()))
                    throw NoriException("nb_bsdf_sample: %s", std::string(nb_last_error()));
                stats::RunningMoments mom;
                for (int k = 0; k < m_sampleCount; ++k) {
                    const float *o = out.data() + (size_t) k * 8;
                    mom.add((double) Color3f(o[3], o[4], o[5]).getLuminance());
                }
                if (verdict(mom, reference)) ++passed;
            }
        }
        cout << "Passed " << passed << "/" << total << " tests." << endl;
        if (passed < total) throw NoriException("Some tests failed :(");
    }

    std::string toString() const {
        return format("StudentsTTest[\n  significanceLevel = %f,\n  sampleCount= %i\n]", m_level, m_sampleCount);
    }

    EClassType getClassType() const { return ETest; }

private:
    bool verdict(const stats::RunningMoments &mom, float reference) const {
        const stats::Verdict v = stats::studentsTTest(mom.mean, mom.variance(), reference, m_sampleCount, m_level, (int) m_references.size());
        cout << format("Sample mean = %f (reference value = %f), sample variance = %g, t-statistic = %f, p-value = %f, "
                       "significance level = %f: %s", mom.mean, reference, mom.variance(), v.tStat, v.pValue, v.level,
                       v.accepted ? "accepted the null hypothesis" : "REJECTED the null hypothesis") << endl;
        return v.accepted;
    }

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
    std::vector<BSDF *> m_bsdfs;
};
NORI_REGISTER_CLASS(StudentsTTest, "ttest");

/// chi^2 goodness-of-fit test of BSDF::sample() against BSDF::pdf() (ref: src/chi2test.cpp:31-214)
class ChiSquareTest : public NoriObject {
public:
    ChiSquareTest(const PropertyList &props) {
        m_level = props.getFloat("significanceLevel", 0.01f);
        m_cosThetaResolution = props.getInteger("resolution", 10);
        m_minExpFrequency = props.getInteger("minExpFrequency", 5);
        m_sampleCount = props.getInteger("sampleCount", -1);
        m_testCount = props.getInteger("testCount", 5);
        if (m_cosThetaResolution < 1 || m_testCount < 1) throw NoriException("ChiSquareTest: resolution and testCount must be positive");
        m_phiResolution = 2 * m_cosThetaResolution;
        if (m_sampleCount < 0) m_sampleCount = m_cosThetaResolution * m_phiResolution * 5000;   // ~5K samples per bin
    }

    ~ChiSquareTest() { for (BSDF *b : m_bsdfs) delete b; }

    void addChild(NoriObject *obj) {
        if (obj->getClassType() != EBSDF)
            throw NoriException("ChiSquareTest::addChild(<%s>) is not supported!", classTypeName(obj->getClassType()));
        m_bsdfs.push_back(static_cast<BSDF *>(obj));
    }

    void activate() {
        const int res = m_cosThetaResolution * m_phiResolution;
        DeviceHandle dev;
        pcg32 random;
        int passed = 0, total = 0;
        std::vector<double> obs((size_t) res), exp((size_t) res);
        std::vector<float> xi((size_t) m_sampleCount * 2), sampled((size_t) m_sampleCount * 8);
        for (BSDF *bsdf : m_bsdfs) {
            nb_bsdf_desc desc; describeBSDF(bsdf, &desc);
            for (int l = 0; l < m_testCount; ++l) {
                std::fill(obs.begin(), obs.end(), 0.0); std::fill(exp.begin(), exp.end(), 0.0);
                cout << "------------------------------------------------------" << endl;
                cout << "Testing: " << bsdf->toString() << endl;
                ++total;
                // incident direction: uniform in cos theta over the upper hemisphere (ref: src/chi2test.cpp:99-103)
                const float cosTheta = random.nextFloat();
                const float sinTheta = std::sqrt(std::max(0.0f, 1 - cosTheta * cosTheta));
                const float phi = 2.0f * (float) M_PI * random.nextFloat();
                const float wi[3] = { std::cos(phi) * sinTheta, std::sin(phi) * sinTheta, cosTheta };
                cout << "Accumulating " << m_sampleCount << " samples into a " << m_cosThetaResolution << "x" << m_phiResolution
                     << " contingency table .. "; cout.flush();
                for (float &x : xi) x = random.nextFloat();
                if (nb_bsdf_sample(dev.ctx, &desc, wi, 0, xi.data(), (uint64_t) m_sampleCount, sampled.data()))
                    throw NoriException("nb_bsdf_sample: %s", std::string(nb_last_error()));
                for (int i = 0; i < m_sampleCount; ++i) {
                    const float *o = sampled.data() + (size_t) i * 8;
                    if (o[3] == 0 && o[4] == 0 && o[5] == 0) continue;               // failed sample
                    const int cosThetaBin = std::min(std::max(0, (int) std::floor((o[2] * 0.5f + 0.5f) * m_cosThetaResolution)), m_cosThetaResolution - 1);
                    float scaledPhi = std::atan2(o[1], o[0]) * INV_TWOPI;
                    if (scaledPhi < 0) scaledPhi += 1;
                    const int phiBin = std::min(std::max(0, (int) std::floor(scaledPhi * m_phiResolution)), m_phiResolution - 1);
                    obs[(size_t) (cosThetaBin * m_phiResolution + phiBin)] += 1;
                }
                cout << "done." << endl;
                cout << "Integrating expected frequencies .. "; cout.flush();
                integrateCells(dev.ctx, desc, wi, exp);
                cout << "done." << endl;
                const stats::Chi2Verdict v = stats::chi2Test(res, obs.data(), exp.data(), m_sampleCount, m_minExpFrequency, m_level,
                                                             m_testCount * (int) m_bsdfs.size());
                cout << format("Chi^2 statistic = %f (d.o.f. = %i), p-value = %f, significance level = %f: %s", v.statistic, v.dof, v.pValue,
                               v.level, v.accepted ? "accepted the null hypothesis" : "REJECTED the null hypothesis") << endl;
                if (v.accepted) ++passed;
            }
        }
        cout << "Passed " << passed << "/" << total << " tests." << endl;
        if (passed < total) throw NoriException("Some tests failed :(");
    }

    std::string toString() const {
        return format("ChiSquareTest[\n  thetaResolution = %i,\n  phiResolution = %i,\n  minExpFrequency = %i,\n  sampleCount = %i,\n"
                      "  testCount = %i,\n  significanceLevel = %f\n]", m_cosThetaResolution, m_phiResolution, m_minExpFrequency,
                      m_sampleCount, m_testCount, m_level);
    }

    EClassType getClassType() const { return ETest; }

private:
    /// Expected cell frequencies: the integral of pdf(wi, .) over every (cos theta, phi) cell times sampleCount
    /// (ref: src/chi2test.cpp:126-153).  Composite Simpson in (theta, phi) -- theta rather than cos theta, because
    /// sin(theta) = sqrt(1 - c^2) has an unbounded derivative at c = 1, which stalls the convergence in the top row --
    /// with the number of intervals doubled per cell until its value moves by less than 1e-5 (a narrow specular lobe at
    /// grazing incidence needs 128 intervals per axis, a diffuse lobe 32).  All nodes of a refinement level go to the
    /// device in one nb_bsdf_eval_pdf batch.
    void integrateCells(nb_ctx *ctx, const nb_bsdf_desc &desc, const float wi[3], std::vector<double> &exp) const {
        const int res = m_cosThetaResolution * m_phiResolution;
        const double dc = 2.0 / m_cosThetaResolution, dp = 2.0 * M_PI / m_phiResolution;
        std::vector<double> previous((size_t) res, -1.0);
        std::vector<int> todo((size_t) res);
        for (int i = 0; i < res; ++i) todo[(size_t) i] = i;
        std::vector<float> wo, evaluated;
        for (int S = 16; !todo.empty(); S *= 2) {
            const size_t perCell = (size_t) (S + 1) * (S + 1);
            const size_t cellsPerBatch = std::max<size_t>(1, (size_t) 2000000 / perCell);
            std::vector<int> next;
            for (size_t first = 0; first < todo.size(); first += cellsPerBatch) {
                const size_t count = std::min(cellsPerBatch, todo.size() - first);
                wo.resize(count * perCell * 3); evaluated.resize(count * perCell * 4);
                size_t q = 0;
                for (size_t k = 0; k < count; ++k) {
                    const int cell = todo[first + k], i = cell / m_phiResolution, j = cell % m_phiResolution;
                    const double t0 = std::acos(std::min(1.0, -1.0 + (i + 1) * dc)), t1 = std::acos(std::max(-1.0, -1.0 + i * dc));
                    // the density jumps at the horizon, which is a cell boundary: a node ON it is evaluated from its own
                    // cell's side (cos(pi/2) rounds to +6e-17 and would leak the upper-hemisphere value into the cell below)
                    const double cLo = -1.0 + i * dc, cHi = -1.0 + (i + 1) * dc;
                    for (int a = 0; a <= S; ++a) {
                        double c = std::min(std::max(std::cos(t0 + (t1 - t0) * a / S), cLo), cHi);
                        if (cLo >= 0) c = std::max(c, 1e-6);
                        const double sd = std::sqrt(std::max(0.0, 1.0 - c * c));
                        for (int b = 0; b <= S; ++b) {
                            const double p = (j + (double) b / S) * dp;
                            wo[q++] = (float) (sd * std::cos(p)); wo[q++] = (float) (sd * std::sin(p)); wo[q++] = (float) c;
                        }
                    }
                }
                if (nb_bsdf_eval_pdf(ctx, &desc, wi, 0, wo.data(), (uint64_t) (count * perCell), evaluated.data()))
                    throw NoriException("nb_bsdf_eval_pdf: %s", std::string(nb_last_error()));
                for (size_t k = 0; k < count; ++k) {
                    const int cell = todo[first + k], i = cell / m_phiResolution;
                    const double t0 = std::acos(std::min(1.0, -1.0 + (i + 1) * dc)), t1 = std::acos(std::max(-1.0, -1.0 + i * dc));
                    double sum = 0;
                    for (int a = 0; a <= S; ++a) {
                        const double wa = (a == 0 || a == S) ? 1 : (a % 2 ? 4 : 2), sn = std::sin(t0 + (t1 - t0) * a / S);
                        for (int b = 0; b <= S; ++b) {
                            const double wb = (b == 0 || b == S) ? 1 : (b % 2 ? 4 : 2);
                            sum += wa * wb * sn * (double) evaluated[(k * perCell + (size_t) a * (S + 1) + b) * 4 + 3];
                        }
                    }
                    const double value = sum * ((t1 - t0) / S / 3.0) * (dp / S / 3.0);
                    const bool converged = previous[(size_t) cell] >= 0 && std::fabs(value - previous[(size_t) cell]) <= 1e-5 * std::fabs(value) + 1e-10;
                    previous[(size_t) cell] = value;
                    exp[(size_t) cell] = value * m_sampleCount;
                    if (!converged && S < 1024) next.push_back(cell);
                }
            }
            todo.swap(next);
        }
    }

    float m_level;
    int m_cosThetaResolution, m_phiResolution, m_minExpFrequency, m_sampleCount, m_testCount;
    std::vector<BSDF *> m_bsdfs;
};
NORI_REGISTER_CLASS(ChiSquareTest, "chi2test");

NORI_NAMESPACE_END

extern "C" {
/// Exposed so that the CPU tests can pin the p-value arithmetic against scipy (no GPU involved).
double nori_host_students_t_pvalue(double t, double dof) { return nori::stats::studentsTwoSided(t, dof); }
double nori_host_chi2_pvalue(double statistic, int dof) { return nori::stats::upperGamma(0.5 * dof, 0.5 * statistic); }
int nori_host_chi2_test(int ncells, const double *obs, const double *exp, int n, double min_exp, double level, int ntests, double *pvalue) {
    const nori::stats::Chi2Verdict v = nori::stats::chi2Test(ncells, obs, exp, n, min_exp, level, ntests);
    if (pvalue) *pvalue = v.pValue;
    return v.accepted ? 1 : 0;
}
int nori_host_students_t_test(double mean, double variance, double reference, int n, double level, int ntests, double *pvalue) {
    const nori::stats::Verdict v = nori::stats::studentsTTest(mean, variance, reference, n, level, ntests);
    if (pvalue) *pvalue = v.pValue;
    return v.accepted ? 1 : 0;
}
}

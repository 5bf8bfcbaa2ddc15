// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// render.cpp -- host side of the drop-in boundary: everything the reference's render() (ref: src/main.cpp:58-148)
// and Scene/Accel glue (ref: src/scene.cpp:27-53) did around the hot loop, now expressed as calls into the C-ABI.
#include <chrono>
#include "nori/parser.h"
#include "nori/render.h"
#include "nori_b200.h"

NORI_NAMESPACE_BEGIN

bool nori_camera_matrices(const Camera *cam, float s2c[16], float c2w[16], float *nearClip, float *farClip);
bool nori_sampler_mode(const Sampler *s, bool *blockMode, uint64_t *seed);

namespace {

[[noreturn]] void throwLast(const char *what) { throw NoriException("%s: %s", std::string(what), std::string(nb_last_error())); }

/// POD descriptor of a BSDF plugin instance from its creation record (type name + PropertyList).  Defaults are the
/// plugin constructors' own (ref: src/diffuse.cpp:19, src/microfacet.cpp:17-36, src/dielectric.cpp:15-20).
nb_bsdf_desc describeBSDF(const BSDF *bsdf) {
    nb_bsdf_desc d; std::memset(&d, 0, sizeof d);
    const NoriObjectFactory::Record *rec = NoriObjectFactory::creationRecord(bsdf);
    if (!rec) throw NoriException("BSDF %s was not created through NoriObjectFactory; the GPU path cannot describe it", bsdf->toString());
    const PropertyList &p = rec->props;
    d.alpha = 0.1f; d.intIOR = 1.5046f; d.extIOR = 1.000277f;
    if (rec->type == "diffuse") {
        d.type = NB_BSDF_DIFFUSE;
        Color3f a = p.getColor("albedo", Color3f(0.5f));
        d.albedo[0] = a.r(); d.albedo[1] = a.g(); d.albedo[2] = a.b();
    } else if (rec->type == "mirror") {
        d.type = NB_BSDF_MIRROR;
    } else if (rec->type == "dielectric") {
        d.type = NB_BSDF_DIELECTRIC;
        d.intIOR = p.getFloat("intIOR", 1.5046f); d.extIOR = p.getFloat("extIOR", 1.000277f);
    } else if (rec->type == "microfacet") {
        d.type = NB_BSDF_MICROFACET;
        d.alpha = p.getFloat("alpha", 0.1f); d.intIOR = p.getFloat("intIOR", 1.5046f); d.extIOR = p.getFloat("extIOR", 1.000277f);
        Color3f kd = p.getColor("kd", Color3f(0.5f));
        d.albedo[0] = kd.r(); d.albedo[1] = kd.g(); d.albedo[2] = kd.b();
        d.ks = 1 - kd.maxCoeff();
    } else {
        throw NoriException("BSDF plugin \"%s\" has no device implementation (the GPU path has no CPU fallback)", rec->type);
    }
    return d;
}

bool describeEmitter(const Emitter *em, nb_emitter_desc &d) {
    std::memset(&d, 0, sizeof d);
    if (!em) return false;
    const NoriObjectFactory::Record *rec = NoriObjectFactory::creationRecord(em);
    if (!rec) throw NoriException("Emitter %s was not created through NoriObjectFactory", em->toString());
    if (rec->type != "area") throw NoriException("Emitter plugin \"%s\" has no device implementation", rec->type);
    Color3f r = rec->props.getColor("radiance");
    d.type = NB_EMITTER_AREA; d.radiance[0] = r.r(); d.radiance[1] = r.g(); d.radiance[2] = r.b();
    return true;
}

nb_integrator_desc describeIntegrator(const Integrator *integ) {
    nb_integrator_desc d; std::memset(&d, 0, sizeof d);
    const NoriObjectFactory::Record *rec = NoriObjectFactory::creationRecord(integ);
    if (!rec) throw NoriException("Integrator %s was not created through NoriObjectFactory", integ->toString());
    static const std::map<std::string, int> types = { { "normals", NB_INT_NORMALS }, { "ao", NB_INT_AO }, { "whitted", NB_INT_WHITTED },
        { "path_mats", NB_INT_PATH_MATS }, { "path_ems", NB_INT_PATH_EMS }, { "path_mis", NB_INT_PATH_MIS }, { "simple", NB_INT_SIMPLE } };
    auto it = types.find(rec->type);
    if (it == types.end()) throw NoriException("Integrator plugin \"%s\" has no device implementation (the GPU path has no CPU fallback)", rec->type);
    d.type = it->second;
    d.rr_start = rec->props.getInteger("rrStart", 3);
    d.max_depth = rec->props.getInteger("maxDepth", 0);
    return d;
}

}  // namespace

void describeBSDF(const BSDF *bsdf, nb_bsdf_desc *out) { *out = describeBSDF(bsdf); }

nb_ctx *createDeviceScene(const Scene *scene, const ImageBlock &film, const RenderOptions &opt) {
    nb_ctx *ctx = nullptr;
    if (opt.gpus > 1) {        // N devices behind the same calls (ref: the TBB tile loop + merge, src/main.cpp:85-113)
        std::vector<int> devices;
        for (int i = 0; i < opt.gpus; ++i) devices.push_back(opt.device + i);
        ctx = nb_create_multi(devices.data(), opt.gpus);
        if (!ctx) throwLast("nb_create_multi");
    } else {
        ctx = nb_create(opt.device);
        if (!ctx) throwLast("nb_create");
    }
    try {
        if (opt.deviceBuilder && nb_set_option(ctx, "builder", 1)) throwLast("nb_set_option");
        if (!opt.accelCache.empty() && nb_set_accel_cache(ctx, opt.accelCache.c_str())) throwLast("nb_set_accel_cache");
        for (const Mesh *mesh : scene->getMeshes()) {   // Scene::addChild(mesh) -> Accel::addMesh (ref: src/scene.cpp:48-53)
            nb_bsdf_desc b = describeBSDF(mesh->getBSDF());
            nb_emitter_desc e;
            bool hasE = describeEmitter(mesh->getEmitter(), e);
            const auto &V = mesh->getVertexPositions(); const auto &N = mesh->getVertexNormals();
            const auto &UV = mesh->getVertexTexCoords(); const auto &F = mesh->getIndices();
            if (nb_add_mesh(ctx, V.data(), mesh->getVertexCount(), N.empty() ? nullptr : N.data(), UV.empty() ? nullptr : UV.data(),
                            F.data(), mesh->getTriangleCount(), &b, hasE ? &e : nullptr) < 0) throwLast("nb_add_mesh");
        }
        if (nb_build_accel(ctx)) throwLast("nb_build_accel");   // Scene::activate -> Accel::build (ref: src/scene.cpp:28)
        float s2c[16], c2w[16], nearClip, farClip;
        if (!nori_camera_matrices(scene->getCamera(), s2c, c2w, &nearClip, &farClip))
            throw NoriException("Camera plugin %s has no device implementation", scene->getCamera()->toString());
        const Vector2i &size = scene->getCamera()->getOutputSize();
        if (nb_set_camera(ctx, s2c, c2w, size.x(), size.y(), nearClip, farClip)) throwLast("nb_set_camera");
        if (nb_set_filter(ctx, film.filterTable(), film.filterRadius())) throwLast("nb_set_filter");
        bool blockMode = false; uint64_t seed = 0;
        if (!nori_sampler_mode(scene->getSampler(), &blockMode, &seed))
            throw NoriException("Sampler plugin %s has no device implementation", scene->getSampler()->toString());
        if (nb_set_sampler(ctx, (uint32_t) scene->getSampler()->getSampleCount(), blockMode ? NB_SEED_PER_BLOCK : NB_SEED_PER_SAMPLE, seed)) throwLast("nb_set_sampler");
        nb_integrator_desc id = describeIntegrator(scene->getIntegrator());
        if (nb_set_integrator(ctx, &id)) throwLast("nb_set_integrator");
        if (id.type == NB_INT_SIMPLE) {
            const PropertyList &ip = NoriObjectFactory::creationRecord(scene->getIntegrator())->props;
            const Point3f pos = ip.getPoint("position"); const Color3f en = ip.getColor("energy");
            const float p3[3] = { pos.x(), pos.y(), pos.z() }, e3[3] = { en.r(), en.g(), en.b() };
            if (nb_set_point_light(ctx, p3, e3)) throwLast("nb_set_point_light");
        }
        if (opt.gpus <= 1 && nb_set_tiles(ctx, opt.tileRank, opt.tileRanks)) throwLast("nb_set_tiles");
    } catch (...) {
        nb_destroy(ctx);
        throw;
    }
    return ctx;
}

void renderScene(Scene *scene, ImageBlock &result, const RenderOptions &opt, nb_stats *stats, std::vector<uint8_t> *srgb8) {
    scene->getIntegrator()->preprocess(scene);                    // ref: src/main.cpp:61
    nb_ctx *ctx = createDeviceScene(scene, result, opt);
    nb_stats st; std::memset(&st, 0, sizeof st);
    result.clear();
    int rc = 0;
    const int spp = (int) scene->getSampler()->getSampleCount();
    if (opt.previewEvery > 0 && opt.gpus <= 1 && opt.previewEvery < spp) {
        // progressive frame: passes of previewEvery samples; after each one the device merges + tonemaps what it has and the
        // host rewrites the preview image -- the job of NoriScreen's refresh (ref: src/gui.cpp:120-138) without a window
        std::vector<uint8_t> p8((size_t) result.getSize().x() * result.getSize().y() * 3);
        rc = nb_render_begin(ctx);
        for (int done = 0; !rc && done < spp; ) {
            const int n = std::min(opt.previewEvery, spp - done);
            rc = nb_render_pass(ctx, (uint32_t) n, &st);
            done += n;
            if (!rc) rc = nb_render_preview(ctx, done >= spp ? result.data() : nullptr, p8.data());
            if (!rc && !opt.previewName.empty()) Bitmap::savePNG8(opt.previewName, result.getSize().x(), result.getSize().y(), p8.data());
        }
        if (!rc) nb_render_end(ctx);
    } else {
        rc = nb_render(ctx, result.data(), &st);                  // replaces ref: src/main.cpp:64-119
    }
    if (!rc && srgb8) {                                           // tonemap + 8-bit pack on the device (ref: src/common.cpp:166-180, src/bitmap.cpp:100-110)
        srgb8->resize((size_t) result.getSize().x() * result.getSize().y() * 3);
        rc = nb_last_film_to_srgb8(ctx, srgb8->data());
    }
    std::string err = rc ? nb_last_error() : "";
    nb_destroy(ctx);
    if (rc) throw NoriException("nb_render: %s", err);
    if (stats) *stats = st;
}

void render(Scene *scene, const std::string &filename, const RenderOptions &opt) {
    const Camera *camera = scene->getCamera();
    ImageBlock result(camera->getOutputSize(), camera->getReconstructionFilter());   // ref: src/main.cpp:67-68
    if (!opt.quiet) { cout << "Rendering .. "; cout.flush(); }
    nb_stats st;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<uint8_t> srgb8;
    RenderOptions ropt = opt;
    if (ropt.previewEvery > 0) { ropt.previewName = filename; size_t dot = ropt.previewName.find_last_of("."); if (dot != std::string::npos) ropt.previewName.erase(dot); ropt.previewName += "_preview"; }
    renderScene(scene, result, ropt, &st, &srgb8);
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    if (!opt.quiet)
        cout << "done. (took " << timeString(ms) << " incl. upload + BVH build; render kernel " << timeString(st.kernel_ms, true) << ", "
             << format("%.1f", (double) st.rays / st.kernel_ms / 1e3) << " Mrays/s, " << format("%.1f", (double) st.samples / st.kernel_ms / 1e3) << " Msamples/s)" << endl;
    std::unique_ptr<Bitmap> bitmap(result.toBitmap());            // ref: src/main.cpp:133-135
    std::string outputName = filename;
    size_t lastdot = outputName.find_last_of(".");
    if (lastdot != std::string::npos) outputName.erase(lastdot, std::string::npos);
    bitmap->saveEXR(outputName);
    Bitmap::savePNG8(outputName, bitmap->cols(), bitmap->rows(), srgb8.data());   // bytes from film_to_srgb8_kernel
}

NORI_NAMESPACE_END

// ---------------------------------------------------------------- C entry points for the Python test / bench harness
extern "C" {

static thread_local std::string g_host_err;
const char *nori_host_last_error(void) { return g_host_err.c_str(); }

/// Parses a scene; returns an opaque Scene* (nullptr on error).  Root must be a <scene>.
void *nori_host_load(const char *xml_path) {
    try {
        nori::NoriObject *root = nori::loadFromXML(xml_path);
        if (root->getClassType() != nori::NoriObject::EScene) { delete root; throw nori::NoriException("root element of \"%s\" is not a scene", std::string(xml_path)); }
        return root;
    } catch (const std::exception &e) { g_host_err = e.what(); return nullptr; }
}
void nori_host_free(void *scene) { delete static_cast<nori::Scene *>(scene); }

/// Scene summary: out[0..7] = width, height, border, spp, n_meshes, n_triangles, integrator type, seed mode
int nori_host_info(void *scene_, int64_t *out) {
    try {
        nori::Scene *scene = static_cast<nori::Scene *>(scene_);
        nori::ImageBlock blk(scene->getCamera()->getOutputSize(), scene->getCamera()->getReconstructionFilter());
        out[0] = scene->getCamera()->getOutputSize().x(); out[1] = scene->getCamera()->getOutputSize().y();
        out[2] = blk.getBorderSize(); out[3] = (int64_t) scene->getSampler()->getSampleCount();
        out[4] = (int64_t) scene->getMeshes().size();
        int64_t nt = 0; for (auto m : scene->getMeshes()) nt += m->getTriangleCount();
        out[5] = nt;
        out[6] = nori::describeIntegrator(scene->getIntegrator()).type;
        bool bm = false; uint64_t seed = 0; nori::nori_sampler_mode(scene->getSampler(), &bm, &seed);
        out[7] = bm ? 1 : 0;
        return 0;
    } catch (const std::exception &e) { g_host_err = e.what(); return 1; }
}

/// Host-side description for parity checks against the Python-built scenes: camera matrices (32 floats), clip planes
/// (2), filter radius (1) + table (33); mesh i: counts and the BSDF / emitter descriptors.
int nori_host_camera(void *scene_, float *out68) {
    try {
        nori::Scene *scene = static_cast<nori::Scene *>(scene_);
        if (!nori::nori_camera_matrices(scene->getCamera(), out68, out68 + 16, out68 + 32, out68 + 33)) throw nori::NoriException("unsupported camera");
        nori::ImageBlock blk(scene->getCamera()->getOutputSize(), scene->getCamera()->getReconstructionFilter());
        out68[34] = blk.filterRadius();
        std::memcpy(out68 + 35, blk.filterTable(), sizeof(float) * 33);
        return 0;
    } catch (const std::exception &e) { g_host_err = e.what(); return 1; }
}
int nori_host_mesh(void *scene_, int i, uint32_t *nv, uint32_t *nf, const float **V, const float **N, const float **UV, const uint32_t **F,
                   nb_bsdf_desc *bsdf, nb_emitter_desc *emitter) {
    try {
        nori::Scene *scene = static_cast<nori::Scene *>(scene_);
        const nori::Mesh *m = scene->getMeshes().at((size_t) i);
        *nv = m->getVertexCount(); *nf = m->getTriangleCount();
        *V = m->getVertexPositions().data(); *N = m->getVertexNormals().empty() ? nullptr : m->getVertexNormals().data();
        *UV = m->getVertexTexCoords().empty() ? nullptr : m->getVertexTexCoords().data(); *F = m->getIndices().data();
        nori::describeBSDF(m->getBSDF(), bsdf);
        nori::describeEmitter(m->getEmitter(), *emitter);
        return 0;
    } catch (const std::exception &e) { g_host_err = e.what(); return 1; }
}

/// Renders through the C-ABI (GPU required).  film: (H+2b) x (W+2b) x 4 floats.
int nori_host_render_gpus(void *scene_, int device, int gpus, const char *accel_cache, float *film, uint8_t *srgb8_device, uint8_t *srgb8_host,
                          nb_stats *stats) {
    try {
        nori::Scene *scene = static_cast<nori::Scene *>(scene_);
        nori::ImageBlock blk(scene->getCamera()->getOutputSize(), scene->getCamera()->getReconstructionFilter());
        nori::RenderOptions opt; opt.device = device; opt.gpus = gpus; opt.quiet = true;
        if (accel_cache) opt.accelCache = accel_cache;
        std::vector<uint8_t> dev8;
        nori::renderScene(scene, blk, opt, stats, srgb8_device ? &dev8 : nullptr);
        std::memcpy(film, blk.data(), sizeof(float) * 4 * (size_t) blk.rows() * blk.cols());
        if (srgb8_device) std::memcpy(srgb8_device, dev8.data(), dev8.size());
        if (srgb8_host) {      // the host loop (Bitmap::toSRGB8) on the same film, for the byte-equality test
            std::unique_ptr<nori::Bitmap> bmp(blk.toBitmap());
            std::vector<uint8_t> h8; bmp->toSRGB8(h8);
            std::memcpy(srgb8_host, h8.data(), h8.size());
        }
        return 0;
    } catch (const std::exception &e) { g_host_err = e.what(); return 1; }
}

int nori_host_render(void *scene_, int device, int tile_rank, int tile_ranks, float *film, nb_stats *stats) {
    try {
        nori::Scene *scene = static_cast<nori::Scene *>(scene_);
        nori::ImageBlock blk(scene->getCamera()->getOutputSize(), scene->getCamera()->getReconstructionFilter());
        nori::RenderOptions opt; opt.device = device; opt.tileRank = tile_rank; opt.tileRanks = tile_ranks; opt.quiet = true;
        nori::renderScene(scene, blk, opt, stats);
        std::memcpy(film, blk.data(), sizeof(float) * 4 * (size_t) blk.rows() * blk.cols());
        return 0;
    } catch (const std::exception &e) { g_host_err = e.what(); return 1; }
}

/// Whether a plugin name is registered with the factory (NORI_REGISTER_CLASS)
int nori_host_is_registered(const char *name) { return nori::NoriObjectFactory::isRegistered(name) ? 1 : 0; }

}  // extern "C"

extern "C" int nori_host_block_order(int W, int H, int32_t *xy) {   // BlockGenerator::next order (ref: src/block.cpp:119-152)
    nori::BlockGenerator gen(nori::Vector2i(W, H), NORI_BLOCK_SIZE);
    nori::ImageBlock blk(nori::Vector2i(NORI_BLOCK_SIZE, NORI_BLOCK_SIZE), nullptr);
    int n = 0;
    while (gen.next(blk)) {
        if (xy) { xy[4 * n] = blk.getOffset().x(); xy[4 * n + 1] = blk.getOffset().y(); xy[4 * n + 2] = blk.getSize().x(); xy[4 * n + 3] = blk.getSize().y(); }
        ++n;
    }
    return n;
}

// Host mirror of the Nori educational ray tracer's interfaces (after Nori, Copyright (c) 2015 by Wenzel Jakob);
// re-implemented here without third-party code so that plugins register and parse unchanged -- see DESIGN.md section 1.
// main.cpp -- `nori <scene.xml>` command line (ref: src/main.cpp:150-246).  --no-gui / --threads are accepted for
// compatibility and ignored (there is no GUI and no TBB on this path); --device N selects the GPU, --gpus N
// renders on N devices (tiles sharded tile_id % N, finished ImageBlocks gathered over NCCL, merged on the first device),
// --preview K renders progressively (K samples per pass) and rewrites <scene>_preview.png after every pass,
// --cache keeps the built BVH in <scene>.nbbvh and reloads it when the geometry is unchanged.
#include <cstring>
#include "nori/parser.h"
#include "nori/render.h"

using namespace nori;

int main(int argc, char **argv) {
    if (argc < 2) {
        cerr << "Syntax: " << argv[0] << " <scene.xml> [--no-gui] [--threads N] [--device N] [--gpus N] [--lbvh] [--cache] [--preview SPP]" << endl;
        return -1;
    }
    std::string sceneName;
    RenderOptions opt;
    for (int i = 1; i < argc; ++i) {
        std::string token(argv[i]);
        if (token == "-t" || token == "--threads") {
            if (i + 1 >= argc) { cerr << "\"--threads\" argument expects a positive integer following it." << endl; return -1; }
            ++i;   // accepted, unused
        } else if (token == "--no-gui") {
        } else if (token == "--lbvh") {
            opt.deviceBuilder = true;   // GPU-built hierarchy (fast build, slightly slower render)
        } else if (token == "--preview") {
            if (i + 1 >= argc || atoi(argv[i + 1]) < 1) { cerr << "\"--preview\" expects a positive sample count following it." << endl; return -1; }
            opt.previewEvery = atoi(argv[++i]);   // progressive frame: <scene>_preview.png rewritten after every pass of that many samples
        } else if (token == "--cache") {
            opt.accelCache = "?";       // resolved below: <scene>.nbbvh next to the scene file
        } else if (token == "--gpus") {
            if (i + 1 >= argc || atoi(argv[i + 1]) < 1) { cerr << "\"--gpus\" expects a positive integer following it." << endl; return -1; }
            opt.gpus = atoi(argv[++i]);   // devices device .. device+N-1 render tile shards; blocks gathered over NCCL
        } else if (token == "--device") {
            if (i + 1 >= argc) { cerr << "\"--device\" expects an integer following it." << endl; return -1; }
            opt.device = atoi(argv[++i]);
        } else {
            sceneName = token;
        }
    }
    try {
        if (endsWith(toLower(sceneName), ".xml")) {
            std::unique_ptr<NoriObject> root(loadFromXML(sceneName));
            /* When the XML root object is a scene, start rendering it (ref: src/main.cpp:236-238) */
            if (opt.accelCache == "?") opt.accelCache = sceneName.substr(0, sceneName.size() - 4) + ".nbbvh";
            if (root->getClassType() == NoriObject::EScene) {
                cout << endl << "Configuration: " << root->toString() << endl << endl;
                render(static_cast<Scene *>(root.get()), sceneName, opt);
            }
        } else {
            cerr << "Fatal error: unknown file \"" << sceneName << "\", expected an extension of type .xml" << endl;
            return -1;
        }
    } catch (const std::exception &e) {
        cerr << "Fatal error: " << e.what() << endl;
        return -1;
    }
    return 0;
}

// b200pt_render: command-line driver over the C ABI (include/b200pt.h) - the counterpart, for this path, of the
// reference's `pbrt-v3-rs` binary (bin/src/main.rs -> api::pbrt_init / parse / pbrt_cleanup): reads a pbrt-v3 scene
// file, renders it on the GPU and writes the image.  Links against libb200pt.so only; no CUDA or Python on this side.
//
//   b200pt_render [--device N | --devices 0,1,..] [--outfile image.png|.pfm] [--cropwindow x0 x1 y0 y1 is taken from the scene file] scene.pbrt
//
// Options follow the reference's where they exist (core/src/app.rs: --outfile; --nthreads / --quick do not apply).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/b200pt.h"

static int fail(const char* what) {
    std::fprintf(stderr, "b200pt_render: %s failed: %s\n", what, b200pt_last_error());
    return 1;
}

int main(int argc, char** argv) {
    int device = 0;
    std::vector<int32_t> devices;  // --devices 0,1,..: Integrator::render over several GPUs of this process (b200pt_render_multi)
    std::string outfile, scene_path;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--device" && i + 1 < argc) device = std::atoi(argv[++i]);
        else if (a == "--devices" && i + 1 < argc) {
            for (const char* p = argv[++i]; *p;) { devices.push_back((int32_t)std::strtol(p, const_cast<char**>(&p), 10)); if (*p == ',') ++p; }
        }
        else if ((a == "--outfile" || a == "-o") && i + 1 < argc) outfile = argv[++i];
        else if (a == "--help" || a == "-h") {
            std::printf("usage: b200pt_render [--device N | --devices 0,1,..] [--outfile image.png|.pfm] scene.pbrt\n"
                        "Renders the scene with the B200 path (path / whitted / directlighting integrators over triangle meshes)\n"
                        "and writes a .png (the reference's 8-bit sRGB encode) or .pfm image (default: the Film's \"filename\").\n");
            return 0;
        } else if (!a.empty() && a[0] == '-') { std::fprintf(stderr, "b200pt_render: unknown option %s\n", a.c_str()); return 2; }
        else scene_path = a;
    }
    if (scene_path.empty()) { std::fprintf(stderr, "b200pt_render: no scene file given (--help)\n"); return 2; }

    using clock = std::chrono::steady_clock;
    auto secs = [](clock::time_point a, clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };

    if (b200pt_init(devices.empty() ? device : devices[0]) != B200PT_OK) return fail("b200pt_init");
    const auto t0 = clock::now();
    b200pt_loaded_scene* loaded = nullptr;
    if (b200pt_load_pbrt(scene_path.c_str(), &loaded) != B200PT_OK) return fail("b200pt_load_pbrt");
    const b200pt_scene_desc* d = b200pt_loaded_scene_desc(loaded);
    const auto t1 = clock::now();
    std::printf("%s: %lld triangles, %d instances, %d lights, %dx%d @ %d spp, integrator %s (parsed + BVH in %.2f s)\n", scene_path.c_str(),
                (long long)d->n_prims, d->n_instances, d->n_lights, d->film.xres, d->film.yres, d->sampler.spp,
                d->integrator.type == B200PT_INTEGRATOR_WHITTED ? "whitted" : d->integrator.type == B200PT_INTEGRATOR_DIRECT ? "directlighting" : "path", secs(t0, t1));

    b200pt_scene* scene = nullptr;
    b200pt_multi* multi = nullptr;
    const int w = d->film.crop[2] - d->film.crop[0], h = d->film.crop[3] - d->film.crop[1];
    std::vector<float> film((size_t)w * h * 4), rgb((size_t)w * h * 3);
    uint64_t rays[3] = {0, 0, 0};
    clock::time_point t2, t3;
    if (!devices.empty()) {
        // the scene replicated on every listed GPU, bands of 8 pixel rows dealt round-robin, gathered over NVLink (NCCL)
        if (b200pt_multi_create(d, devices.data(), (int32_t)devices.size(), &multi) != B200PT_OK) return fail("b200pt_multi_create");
        t2 = clock::now();
        if (b200pt_multi_render(multi, 8, film.data()) != B200PT_OK) return fail("b200pt_multi_render");
        t3 = clock::now();
        double gather_ms = 0.0;
        int32_t nccl = 0;
        b200pt_multi_info(multi, rays, &gather_ms, &nccl);
        std::printf("%d devices, band gather %.3f ms (%s)\n", (int)devices.size(), gather_ms, nccl ? "NCCL" : "peer copies");
    } else {
        if (b200pt_scene_create(d, &scene) != B200PT_OK) return fail("b200pt_scene_create");
        t2 = clock::now();
        if (b200pt_render_rows(scene, 0, h, film.data()) != B200PT_OK) return fail("b200pt_render_rows");
        t3 = clock::now();
        b200pt_scene_ray_counts(scene, rays);
    }
    if (b200pt_film_resolve(&d->film, film.data(), rgb.data()) != B200PT_OK) return fail("b200pt_film_resolve");
    const double dt = secs(t2, t3);
    std::printf("rendered in %.3f s: %.3e samples/s, %.1f Mrays/s (%llu camera, %llu closest-hit, %llu shadow rays)\n", dt, (double)rays[0] / dt,
                (double)(rays[1] + rays[2]) / dt / 1e6, (unsigned long long)rays[0], (unsigned long long)rays[1], (unsigned long long)rays[2]);

    if (outfile.empty()) {  // the Film's "filename"; extensions this path does not write (.exr, .tga) become .pfm
        outfile = b200pt_loaded_scene_output(loaded);
        const size_t dot = outfile.find_last_of('.');
        const std::string ext = dot == std::string::npos ? std::string() : outfile.substr(dot);
        if (ext != ".png" && ext != ".pfm") outfile = (dot == std::string::npos ? outfile : outfile.substr(0, dot)) + ".pfm";
    }
    if (b200pt_write_image(outfile.c_str(), rgb.data(), w, h) != B200PT_OK) return fail("b200pt_write_image");
    std::printf("wrote %s\n", outfile.c_str());
    if (multi) b200pt_multi_destroy(multi);
    if (scene) b200pt_scene_destroy(scene);
    b200pt_loaded_scene_free(loaded);
    return 0;
}

// lyap_render -- headless frame renderer; what the reference's lyap_interactive.cu does
// between init_scene() and cleanup() (:111-136, :696-741) without GLUT/GL: the scene (defaults of
// params_init, a scene file, and/or single-key overrides -- the run-time replacement for editing
// params.cu or pressing keys in the viewer, :144-463), lights and camera recalculated, one frame
// rendered, saved under the reference's self-describing file name as P3 PPM (save_ppm) and/or
// PNG, plus the raw LyapPoint dump (save_points) and the effective scene.  Animation frames
// follow scale.pl's camera path at run time instead of rewriting params.cu and recompiling.
//
//   lyap_render [-scene file] [-set key=value]...            (any key of include/lyap/scene.h)
//               [-w 3840] [-h 2160] [-seq BCABA] [-settle n] [-accum n] [-d x] [-depth n]
//               [-jitter x] [-refine n] [-ot x] [-M x]       (shorthands for -set)
//               [-mode exact|fast|host|hybrid|hybrid_host]
//               [-frames N [-first a] [-last b]]             (orbit; default: the scene's camera)
//               [-ppm] [-png] [-points] [-dump-scene] [-dir .] [-device 0]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "lyap/abi.h"
#include "lyap/scene.h"

static int mode_of(const char *s)
{
    if (!strcmp(s, "fast")) return LYAP_MODE_FAST;
    if (!strcmp(s, "host")) return LYAP_MODE_HOST;
    if (!strcmp(s, "hybrid")) return LYAP_MODE_HYBRID;
    if (!strcmp(s, "hybrid_host")) return LYAP_MODE_HYBRID_HOST;
    return LYAP_MODE_EXACT;
}

int main(int argc, char **argv)
{
    static lyap_scene sc;
    lyap_scene_defaults(&sc);
    char err[256];

    int mode = LYAP_MODE_EXACT, device = 0;
    unsigned frames = 0, first = 0, last = ~0u;
    bool ppm = false, png = false, points = false, dump_scene = false;
    std::string dir = ".";
    // a scene file is the base the other options modify, wherever it appears on the command line
    for (int i = 1; i + 1 < argc; ++i)
        if (!strcmp(argv[i], "-scene")) {
            const int rc = lyap_scene_load(&sc, argv[i + 1], err, sizeof err);
            if (rc != LYAP_OK) { fprintf(stderr, "%s: %s\n", argv[i + 1], err); return 2; }
        }
    auto set = [&](const std::string &kv) {
        std::string line = kv;
        const size_t eq = line.find('=');
        if (eq != std::string::npos) line = line.substr(0, eq) + " = " + line.substr(eq + 1);
        if (lyap_scene_parse(&sc, line.c_str(), err, sizeof err) != LYAP_OK) { fprintf(stderr, "-set %s: %s\n", kv.c_str(), err); exit(2); }
    };
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> std::string { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "-scene") next();
        else if (a == "-set") set(next());
        else if (a == "-w") set("width=" + next());
        else if (a == "-h") set("height=" + next());
        else if (a == "-seq") set("sequence=" + next());
        else if (a == "-settle") set("settle=" + next());
        else if (a == "-accum") set("accum=" + next());
        else if (a == "-d") set("d=" + next());
        else if (a == "-depth") set("depth=" + next());
        else if (a == "-jitter") set("jitter=" + next());
        else if (a == "-refine") set("refine=" + next());
        else if (a == "-ot") set("opaqueThreshold=" + next());
        else if (a == "-M") set("cam.M=" + next());
        else if (a == "-mode") mode = mode_of(next().c_str());
        else if (a == "-frames") frames = (unsigned)atoi(next().c_str());
        else if (a == "-first") first = (unsigned)atoi(next().c_str());
        else if (a == "-last") last = (unsigned)atoi(next().c_str());
        else if (a == "-ppm") ppm = true;
        else if (a == "-png") png = true;
        else if (a == "-points") points = true;
        else if (a == "-dump-scene") dump_scene = true;
        else if (a == "-dir") dir = next();
        else if (a == "-device" || a.rfind("-device=", 0) == 0) device = a == "-device" ? atoi(next().c_str()) : atoi(a.c_str() + 8);
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (!ppm && !png) png = true;
    int32_t *seq = nullptr;
    if (!lyap_scene_convert_sequence(&seq, (const unsigned char *)sc.sequence)) return 1;

    const uint32_t w = sc.width, h = sc.height;
    std::vector<lyap_rgba> rgba((size_t)w * h);
    std::vector<lyap_point> pts(points ? (size_t)w * h : 0);
    const unsigned f_end = frames ? (last < frames ? last + 1 : frames) : 1;
    for (unsigned f = frames ? first : 0; f < f_end; ++f) {
        if (frames) lyap_campath_frame(f, frames, &sc.cam);
        lyap_scene_finalize(&sc, 0, 0);
        const time_t stamp = time(nullptr);
        unsigned long long evals = 0;
        const auto t0 = std::chrono::steady_clock::now();
        const int rc = lyap_render_host(rgba.data(), points ? pts.data() : nullptr, &sc.cam, &sc.prm, seq, sc.lights, sc.num_lights,
                                        w, h, mode, device, &evals);
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (rc != LYAP_OK) { fprintf(stderr, "render failed: %s\n", lyap_error_string(rc)); return 1; }
        const unsigned secs = (unsigned)s;
        char stem[400], tail[64], path[600];
        lyap_format_filename(stem, sizeof stem, "Render", (unsigned long)stamp, w, h, sc.sequence, &sc.cam, &sc.prm);
        snprintf(tail, sizeof tail, "_time=%dh%02dm%02ds", secs / 3600, secs / 60 % 60, secs % 60);
        if (ppm) { snprintf(path, sizeof path, "%s/%s%s.ppm", dir.c_str(), stem, tail); if (lyap_write_ppm(path, rgba.data(), w, h)) return 1; }
        if (png) { snprintf(path, sizeof path, "%s/%s%s.png", dir.c_str(), stem, tail); if (lyap_write_png(path, rgba.data(), w, h)) return 1; }
        if (dump_scene) {
            char spath[600];
            snprintf(spath, sizeof spath, "%s/%s%s.scene", dir.c_str(), stem, tail);
            if (lyap_scene_save(&sc, spath)) return 1;
        }
        if (points) {
            lyap_format_filename(stem, sizeof stem, "Points", (unsigned long)stamp, w, h, sc.sequence, &sc.cam, &sc.prm);
            snprintf(path, sizeof path, "%s/%s%s.raw", dir.c_str(), stem, tail);
            if (lyap_write_raw(path, pts.data(), pts.size() * sizeof(lyap_point))) return 1;
        }
        printf("frame %u: %ux%u, %llu exponent evaluations, %.3f s (%.1f Giter/s end to end) -> %s\n", f, w, h, evals, s,
               (double)evals * (sc.prm.settle + sc.prm.accum) / s / 1e9, path);
    }
    free(seq);
    return 0;
}

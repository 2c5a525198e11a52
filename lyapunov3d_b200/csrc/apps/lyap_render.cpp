// lyap_render -- headless frame renderer; what the reference's lyap_interactive.cu does
// between init_scene() and cleanup() (:111-136, :696-741) without GLUT/GL: defaults from
// params_init, lights and camera recalculated, one frame rendered, saved under the
// reference's self-describing file name as P3 PPM (save_ppm) and/or PNG, plus the raw
// LyapPoint dump (save_points).  Animation frames follow scale.pl's camera path at run time
// instead of rewriting params.cu and recompiling.
//
//   lyap_render [-w 3840] [-h 2160] [-seq BCABA] [-settle n] [-accum n] [-d x] [-depth n]
//               [-jitter x] [-refine n] [-ot x] [-M x] [-mode exact|fast|host]
//               [-frames N [-first a] [-last b]]   (orbit; default: the shipped camera)
//               [-ppm] [-png] [-points] [-dir .] [-device 0]
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "lyap/abi.h"

static int mode_of(const char *s)
{
    if (!strcmp(s, "fast")) return LYAP_MODE_FAST;
    if (!strcmp(s, "host")) return LYAP_MODE_HOST;
    return LYAP_MODE_EXACT;
}

int main(int argc, char **argv)
{
    lyap_params prm;
    lyap_cam cam;
    std::vector<lyap_light> lights(LYAP_MAX_LIGHTS);
    uint32_t n_lights = 0, w = 0, h = 0;
    char seq_str[256];
    lyap_params_init(&prm, &cam, lights.data(), &n_lights, seq_str, sizeof seq_str, &w, &h);

    int mode = LYAP_MODE_EXACT, device = 0;
    unsigned frames = 0, first = 0, last = ~0u;
    bool ppm = false, png = false, points = false;
    std::string dir = ".";
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "-w") w = (uint32_t)atoi(next());
        else if (a == "-h") h = (uint32_t)atoi(next());
        else if (a == "-seq") snprintf(seq_str, sizeof seq_str, "%s", next());
        else if (a == "-settle") prm.settle = (uint32_t)atoi(next());
        else if (a == "-accum") prm.accum = (uint32_t)atoi(next());
        else if (a == "-d") prm.d = (float)atof(next());
        else if (a == "-depth") prm.depth = (float)atof(next());
        else if (a == "-jitter") prm.jitter = (float)atof(next());
        else if (a == "-refine") prm.refine = (float)atof(next());
        else if (a == "-ot") prm.opaqueThreshold = (float)atof(next());
        else if (a == "-M") cam.M = (float)atof(next());
        else if (a == "-mode") mode = mode_of(next());
        else if (a == "-frames") frames = (unsigned)atoi(next());
        else if (a == "-first") first = (unsigned)atoi(next());
        else if (a == "-last") last = (unsigned)atoi(next());
        else if (a == "-ppm") ppm = true;
        else if (a == "-png") png = true;
        else if (a == "-points") points = true;
        else if (a == "-dir") dir = next();
        else if (a == "-device" || a.rfind("-device=", 0) == 0) device = a == "-device" ? atoi(next()) : atoi(a.c_str() + 8);
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (!ppm && !png) png = true;
    int32_t *seq = nullptr;
    if (!lyap_scene_convert_sequence(&seq, (const unsigned char *)seq_str)) return 1;
    lyap_scene_lights_recalculate(lights.data(), n_lights);

    std::vector<lyap_rgba> rgba((size_t)w * h);
    std::vector<lyap_point> pts(points ? (size_t)w * h : 0);
    const unsigned f_end = frames ? (last < frames ? last + 1 : frames) : 1;
    for (unsigned f = frames ? first : 0; f < f_end; ++f) {
        if (frames) lyap_campath_frame(f, frames, &cam);
        lyap_scene_cam_recalculate(&cam, w, h, 1);
        const time_t stamp = time(nullptr);
        unsigned long long evals = 0;
        const auto t0 = std::chrono::steady_clock::now();
        const int rc = lyap_render_host(rgba.data(), points ? pts.data() : nullptr, &cam, &prm, seq, lights.data(), n_lights,
                                        w, h, mode, device, &evals);
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (rc != LYAP_OK) { fprintf(stderr, "render failed: %s\n", lyap_error_string(rc)); return 1; }
        const unsigned secs = (unsigned)s;
        char stem[400], tail[64], path[600];
        lyap_format_filename(stem, sizeof stem, "Render", (unsigned long)stamp, w, h, seq_str, &cam, &prm);
        snprintf(tail, sizeof tail, "_time=%dh%02dm%02ds", secs / 3600, secs / 60 % 60, secs % 60);
        if (ppm) { snprintf(path, sizeof path, "%s/%s%s.ppm", dir.c_str(), stem, tail); if (lyap_write_ppm(path, rgba.data(), w, h)) return 1; }
        if (png) { snprintf(path, sizeof path, "%s/%s%s.png", dir.c_str(), stem, tail); if (lyap_write_png(path, rgba.data(), w, h)) return 1; }
        if (points) {
            lyap_format_filename(stem, sizeof stem, "Points", (unsigned long)stamp, w, h, seq_str, &cam, &prm);
            snprintf(path, sizeof path, "%s/%s%s.raw", dir.c_str(), stem, tail);
            if (lyap_write_raw(path, pts.data(), pts.size() * sizeof(lyap_point))) return 1;
        }
        printf("frame %u: %ux%u, %llu exponent evaluations, %.3f s (%.1f Giter/s end to end) -> %s\n", f, w, h, evals, s,
               (double)evals * (prm.settle + prm.accum) / s / 1e9, path);
    }
    free(seq);
    return 0;
}

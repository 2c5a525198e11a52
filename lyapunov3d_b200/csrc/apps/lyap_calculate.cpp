// lyap_calculate -- headless volume bake; the reference's lyap_calculate.cu re-hosted on the
// C ABI (params_init -> sequence -> bake -> D2H -> fwrite "exps.raw", lyap_calculate.cu:58-91),
// with the volume size, sequence, mode and output name as run-time options instead of
// compile-time constants.
//
//   lyap_calculate [-scene file] [-n 512] [-seq BCABA] [-settle 18] [-accum 1008] [-d 2.1]
//                  [-mode fast|exact|host] [-f16] [-device 0] [-o exps.raw]
// (-scene: d, settle, accum and the sequence of a scene file, include/lyap/scene.h)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "lyap/abi.h"
#include "lyap/scene.h"

static int mode_of(const char *s)
{
    if (!strcmp(s, "exact")) return LYAP_MODE_EXACT;
    if (!strcmp(s, "host")) return LYAP_MODE_HOST;
    return LYAP_MODE_FAST;
}

int main(int argc, char **argv)
{
    lyap_params prm;
    lyap_cam cam;
    std::vector<lyap_light> lights(LYAP_MAX_LIGHTS);
    uint32_t n_lights = 0, iw = 0, ih = 0;
    char seq_str[LYAP_SCENE_SEQUENCE_CAP];
    lyap_params_init(&prm, &cam, lights.data(), &n_lights, seq_str, sizeof seq_str, &iw, &ih);

    unsigned n = 512;
    int mode = LYAP_MODE_FAST, dtype = LYAP_F32, device = 0;
    std::string out = "exps.raw";
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
        if (a == "-scene") {
            static lyap_scene sc;
            char err[256];
            const char *path = next();
            if (lyap_scene_load(&sc, path, err, sizeof err) != LYAP_OK) { fprintf(stderr, "%s: %s\n", path, err); return 2; }
            prm = sc.prm;
            snprintf(seq_str, sizeof seq_str, "%s", sc.sequence);
        }
        else if (a == "-n") n = (unsigned)atoi(next());
        else if (a == "-seq") snprintf(seq_str, sizeof seq_str, "%s", next());
        else if (a == "-settle") prm.settle = (uint32_t)atoi(next());
        else if (a == "-accum") prm.accum = (uint32_t)atoi(next());
        else if (a == "-d") prm.d = (float)atof(next());
        else if (a == "-mode") mode = mode_of(next());
        else if (a == "-f16") dtype = LYAP_F16;
        else if (a == "-device" || a.rfind("-device=", 0) == 0) device = a == "-device" ? atoi(next()) : atoi(a.c_str() + 8);
        else if (a == "-o") out = next();
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    int32_t *seq = nullptr;
    if (!lyap_scene_convert_sequence(&seq, (const unsigned char *)seq_str)) return 1;

    const size_t bytes = (size_t)n * n * n * (dtype == LYAP_F16 ? 2 : 4);
    printf("Points size = %ld\n", (long)bytes);
    std::vector<char> host(bytes);
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = lyap_bake_host(host.data(), dtype, &prm, seq, n, n, n, 0, n, mode, device);
    const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    free(seq);
    if (rc != LYAP_OK) { fprintf(stderr, "bake failed: %s\n", lyap_error_string(rc)); return 1; }
    printf("baked %u^3 voxels x %u iterations in %.3f s (%.1f Giter/s incl. copies)\n", n, prm.settle + prm.accum, s,
           (double)n * n * n * (prm.settle + prm.accum) / s / 1e9);
    if (lyap_write_raw(out.c_str(), host.data(), bytes) != LYAP_OK) { fprintf(stderr, "cannot write %s\n", out.c_str()); return 1; }
    return 0;
}

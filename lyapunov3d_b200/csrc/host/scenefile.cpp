// scenefile.cpp -- the run-time scene file (include/lyap/scene.h): parser and writer for the
// parameter surface the reference hard-codes in params.cu:21-114 and edits from the keyboard in
// lyap_interactive.cu:144-463.  Host-only.
#include <cerrno>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "lyap/abi.h"
#include "lyap/scene.h"

namespace {

struct Field {
    const char *name;
    int kind;        // 'f' float, 'u' uint32, '3' vec3, '4' four floats
    size_t offset;
};

#define PF(n) {#n, 'f', offsetof(lyap_params, n)}
#define PU(n) {#n, 'u', offsetof(lyap_params, n)}
const Field kParamFields[] = {PF(d), PU(settle), PU(accum), PU(stepMethod), PF(nearThreshold), PF(nearMultiplier),
                              PF(opaqueThreshold), PF(chaosThreshold), PF(depth), PF(jitter), PF(refine), PF(gradient),
                              PF(lMin), PF(lMax)};
#undef PF
#undef PU

#define CF(n, k) {#n, k, offsetof(lyap_camlight, n)}
const Field kCamFields[] = {CF(C, '3'), CF(Q, '4'), CF(M, 'f')};
const Field kLightFields[] = {CF(C, '3'), CF(Q, '4'), CF(M, 'f'), CF(lightRange, 'f'), CF(ambient, '4'), CF(diffuseColor, '4'),
                              CF(diffusePower, 'f'), CF(specularColor, '4'), CF(specularPower, 'f'), CF(specularHardness, 'f'),
                              CF(chaosColor, '4')};
#undef CF

int fail(char *err, size_t cap, const char *fmt, ...)
{
    if (err && cap) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(err, cap, fmt, ap);
        va_end(ap);
    }
    return LYAP_ERR_BAD_ARGUMENT;
}

std::string trim(const std::string &s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? std::string() : s.substr(a, b - a + 1);
}

bool numbers(const std::string &v, double *out, int want)
{
    const char *p = v.c_str();
    for (int k = 0; k < want; ++k) {
        char *end = nullptr;
        errno = 0;
        out[k] = strtod(p, &end);
        if (end == p) return false;
        p = end;
        while (*p == ' ' || *p == '\t' || *p == ',') ++p;
    }
    return *p == 0;
}

bool store(void *base, const Field &f, const std::string &v)
{
    char *dst = (char *)base + f.offset;
    double x[4];
    switch (f.kind) {
    case 'f':
        if (!numbers(v, x, 1)) return false;
        *(float *)dst = (float)x[0];
        return true;
    case 'u':
        if (!numbers(v, x, 1) || x[0] < 0 || x[0] > 4294967295.0 || x[0] != (double)(uint32_t)x[0]) return false;
        *(uint32_t *)dst = (uint32_t)x[0];
        return true;
    case '3':
    case '4': {
        const int n = f.kind - '0';
        if (!numbers(v, x, n)) return false;
        for (int k = 0; k < n; ++k) ((float *)dst)[k] = (float)x[k];
        return true;
    }
    }
    return false;
}

void emit(std::string &out, const char *prefix, const Field &f, const void *base)
{
    char buf[160];
    const char *src = (const char *)base + f.offset;
    const float *v = (const float *)src;
    switch (f.kind) {
    case 'f': snprintf(buf, sizeof buf, "%s%s = %.9g\n", prefix, f.name, (double)v[0]); break;
    case 'u': snprintf(buf, sizeof buf, "%s%s = %u\n", prefix, f.name, *(const uint32_t *)src); break;
    case '3': snprintf(buf, sizeof buf, "%s%s = %.9g %.9g %.9g\n", prefix, f.name, (double)v[0], (double)v[1], (double)v[2]); break;
    default: snprintf(buf, sizeof buf, "%s%s = %.9g %.9g %.9g %.9g\n", prefix, f.name, (double)v[0], (double)v[1], (double)v[2], (double)v[3]); break;
    }
    out += buf;
}

} // namespace

extern "C" {

void lyap_scene_defaults(lyap_scene *sc)
{
    memset(sc, 0, sizeof *sc);
    lyap_params_init(&sc->prm, &sc->cam, sc->lights, &sc->num_lights, sc->sequence, sizeof sc->sequence, &sc->width, &sc->height);
}

int lyap_scene_parse(lyap_scene *sc, const char *text, char *err, size_t err_cap)
{
    if (!sc || !text) return fail(err, err_cap, "null argument");
    if (err && err_cap) err[0] = 0;
    int line_no = 0;
    const char *p = text;
    while (*p) {
        const char *eol = strchr(p, '\n');
        std::string line = eol ? std::string(p, eol - p) : std::string(p);
        p = eol ? eol + 1 : p + line.size();
        ++line_no;
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        line = trim(line);
        if (line.empty()) continue;
        const size_t eq = line.find('=');
        if (eq == std::string::npos) return fail(err, err_cap, "line %d: expected key = value", line_no);
        const std::string key = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
        if (val.empty()) return fail(err, err_cap, "line %d: no value for '%s'", line_no, key.c_str());
        double x[2];
        bool known = false, ok = true;
        if (key == "sequence") {
            known = true;
            int32_t *seq = nullptr;
            ok = val.size() < sizeof sc->sequence && lyap_scene_convert_sequence(&seq, (const unsigned char *)val.c_str()) != 0;
            free(seq);
            if (ok) snprintf(sc->sequence, sizeof sc->sequence, "%s", val.c_str());
        } else if (key == "width" || key == "height" || key == "lights") {
            known = true;
            ok = numbers(val, x, 1) && x[0] >= 0 && x[0] == (double)(uint32_t)x[0] && (key != "lights" || x[0] <= LYAP_MAX_LIGHTS) &&
                 (key == "lights" || x[0] >= 1);
            if (ok) (key == "width" ? sc->width : key == "height" ? sc->height : sc->num_lights) = (uint32_t)x[0];
        } else if (key == "cam.orbit") {
            known = true;
            ok = numbers(val, x, 1);
            if (ok) lyap_campath_orbit(x[0], &sc->cam);
        } else if (key == "cam.orbit_frame") {
            known = true;
            ok = numbers(val, x, 2) && x[1] >= 1 && x[0] >= 0 && x[0] < x[1];
            if (ok) lyap_campath_frame((uint32_t)x[0], (uint32_t)x[1], &sc->cam);
        } else if (key.rfind("cam.", 0) == 0) {
            for (const Field &f : kCamFields)
                if (key.substr(4) == f.name) { known = true; ok = store(&sc->cam, f, val); }
        } else if (key.rfind("light", 0) == 0 && key.find('.') != std::string::npos) {
            const size_t dot = key.find('.');
            char *end = nullptr;
            const long k = strtol(key.c_str() + 5, &end, 10);
            if (end == key.c_str() + dot && dot > 5 && k >= 0 && k < LYAP_MAX_LIGHTS)
                for (const Field &f : kLightFields)
                    if (key.substr(dot + 1) == f.name) { known = true; ok = store(&sc->lights[k], f, val); }
        } else {
            for (const Field &f : kParamFields)
                if (key == f.name) { known = true; ok = store(&sc->prm, f, val); }
        }
        if (!known) return fail(err, err_cap, "line %d: unknown key '%s'", line_no, key.c_str());
        if (!ok) return fail(err, err_cap, "line %d: bad value '%s' for '%s'", line_no, val.c_str(), key.c_str());
    }
    return LYAP_OK;
}

int lyap_scene_load(lyap_scene *sc, const char *path, char *err, size_t err_cap)
{
    if (!sc || !path) return fail(err, err_cap, "null argument");
    FILE *f = fopen(path, "rb");
    if (!f) {
        if (err && err_cap) snprintf(err, err_cap, "cannot open %s", path);
        return LYAP_ERR_IO;
    }
    std::string text;
    char buf[4096];
    size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
    fclose(f);
    lyap_scene_defaults(sc);
    return lyap_scene_parse(sc, text.c_str(), err, err_cap);
}

void lyap_scene_finalize(lyap_scene *sc, uint32_t width, uint32_t height)
{
    if (width) sc->width = width;
    if (height) sc->height = height;
    lyap_scene_lights_recalculate(sc->lights, sc->num_lights);     // update_scene(), lyap_interactive.cu:124-136
    lyap_scene_cam_recalculate(&sc->cam, sc->width, sc->height, 1);
}

size_t lyap_scene_format(const lyap_scene *sc, char *out, size_t cap)
{
    std::string s = "# lyapunov3d scene (include/lyap/scene.h); derived fields are recomputed on load\n";
    char buf[64];
    s += std::string("sequence = ") + sc->sequence + "\n";
    snprintf(buf, sizeof buf, "width = %u\nheight = %u\n", sc->width, sc->height);
    s += buf;
    for (const Field &f : kParamFields) emit(s, "", f, &sc->prm);
    for (const Field &f : kCamFields) emit(s, "cam.", f, &sc->cam);
    snprintf(buf, sizeof buf, "lights = %u\n", sc->num_lights);
    s += buf;
    for (uint32_t k = 0; k < sc->num_lights && k < LYAP_MAX_LIGHTS; ++k) {
        snprintf(buf, sizeof buf, "light%u.", k);
        for (const Field &f : kLightFields) emit(s, buf, f, &sc->lights[k]);
    }
    if (out && cap) snprintf(out, cap, "%s", s.c_str());
    return s.size();
}

int lyap_scene_save(const lyap_scene *sc, const char *path)
{
    if (!sc || !path) return LYAP_ERR_BAD_ARGUMENT;
    std::vector<char> text(lyap_scene_format(sc, nullptr, 0) + 1);
    lyap_scene_format(sc, text.data(), text.size());
    FILE *f = fopen(path, "wb");
    if (!f) return LYAP_ERR_IO;
    const bool ok = fwrite(text.data(), 1, text.size() - 1, f) == text.size() - 1;
    return (fclose(f) == 0 && ok) ? LYAP_OK : LYAP_ERR_IO;
}

} // extern "C"

// imageio.cpp -- headless output in the reference's own formats plus PNG.
//
// * P3 PPM byte-for-byte as save_ppm writes it (reference lyap_interactive.cu:595-606)
// * raw LyapPoint[] / float volume dumps (save_points :642-647, lyap_calculate.cu:84-86)
// * the self-describing file stem (lyap_interactive.cu:577-590)
// * 8-bit RGB PNG through zlib (the reference shells out to ImageMagick, GNUmakefile:62)
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "lyap/abi.h"

extern "C" {

int lyap_write_ppm(const char *path, const lyap_rgba *px, uint32_t width, uint32_t height)
{
    FILE *f = fopen(path, "w");
    if (!f) return LYAP_ERR_IO;
    fprintf(f, "P3\n%d %d\n%d\n", (int)width, (int)height, 255);
    // same text as fprintf("%3d %3d %3d ") per pixel, assembled in a buffer
    std::vector<char> line((size_t)width * 12 + 1);
    for (uint32_t y = 0; y < height; ++y) {
        char *o = line.data();
        const lyap_rgba *row = px + (size_t)y * width;
        for (uint32_t x = 0; x < width; ++x) {
            const uint8_t c[3] = {row[x].r, row[x].g, row[x].b};
            for (int k = 0; k < 3; ++k) {
                const unsigned v = c[k];
                o[0] = v >= 100 ? (char)('0' + v / 100) : ' ';
                o[1] = v >= 10 ? (char)('0' + (v / 10) % 10) : ' ';
                o[2] = (char)('0' + v % 10);
                o[3] = ' ';
                o += 4;
            }
        }
        if (fwrite(line.data(), 1, (size_t)(o - line.data()), f) != (size_t)(o - line.data())) {
            fclose(f);
            return LYAP_ERR_IO;
        }
    }
    fputc('\n', f);
    return fclose(f) == 0 ? LYAP_OK : LYAP_ERR_IO;
}

int lyap_write_raw(const char *path, const void *data, uint64_t bytes)
{
    FILE *f = fopen(path, "wb");
    if (!f) return LYAP_ERR_IO;
    const size_t w = fwrite(data, 1, (size_t)bytes, f);
    const int rc = fclose(f);
    return (w == bytes && rc == 0) ? LYAP_OK : LYAP_ERR_IO;
}

static void put_be32(std::vector<uint8_t> &v, uint32_t x)
{
    v.push_back((uint8_t)(x >> 24));
    v.push_back((uint8_t)(x >> 16));
    v.push_back((uint8_t)(x >> 8));
    v.push_back((uint8_t)x);
}

static void put_chunk(std::vector<uint8_t> &out, const char tag[4], const uint8_t *data, size_t n)
{
    put_be32(out, (uint32_t)n);
    const size_t start = out.size();
    out.insert(out.end(), tag, tag + 4);
    if (n) out.insert(out.end(), data, data + n);
    put_be32(out, (uint32_t)crc32(0L, out.data() + start, (uInt)(n + 4)));
}

int lyap_write_png(const char *path, const lyap_rgba *px, uint32_t width, uint32_t height)
{
    // filter byte 0 + RGB triples per scanline (alpha dropped, as the PPM path does)
    std::vector<uint8_t> rawimg((size_t)height * (1 + (size_t)width * 3));
    uint8_t *o = rawimg.data();
    for (uint32_t y = 0; y < height; ++y) {
        *o++ = 0;
        const lyap_rgba *row = px + (size_t)y * width;
        for (uint32_t x = 0; x < width; ++x) {
            *o++ = row[x].r;
            *o++ = row[x].g;
            *o++ = row[x].b;
        }
    }
    uLongf zlen = compressBound((uLong)rawimg.size());
    std::vector<uint8_t> z(zlen);
    if (compress2(z.data(), &zlen, rawimg.data(), (uLong)rawimg.size(), 6) != Z_OK) return LYAP_ERR_IO;

    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, width);
    put_be32(ihdr, height);
    const uint8_t tail[5] = {8, 2, 0, 0, 0};  // 8-bit, colour type 2 (RGB)
    ihdr.insert(ihdr.end(), tail, tail + 5);
    put_chunk(out, "IHDR", ihdr.data(), ihdr.size());
    put_chunk(out, "IDAT", z.data(), zlen);
    put_chunk(out, "IEND", nullptr, 0);
    return lyap_write_raw(path, out.data(), out.size());
}

int lyap_format_filename(char *out, size_t cap, const char *prefix, unsigned long timestamp, uint32_t width, uint32_t height,
                         const char *sequence, const lyap_cam *cam, const lyap_params *prm)
{
    const int n = snprintf(out, cap, "%s_%ld_%dx%d_%s_cx=%.8g_cy=%.8g_cz=%.8g_step=%d_D=%g_i=%d,%d_d=%d_j=%g_r=%d_ot=%g",
                           prefix, (long)timestamp, (int)width, (int)height, sequence,
                           cam->C.x, cam->C.y, cam->C.z, (int)prm->stepMethod, prm->d, (int)prm->settle, (int)prm->accum,
                           (unsigned int)prm->depth, prm->jitter, (unsigned int)prm->refine, prm->opaqueThreshold);
    return (n > 0 && (size_t)n < cap) ? LYAP_OK : LYAP_ERR_BAD_ARGUMENT;
}

} // extern "C"

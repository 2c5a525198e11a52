/*
 * lyap/scene.h -- a run-time scene: everything the reference hard-codes in params_init()
 * (params.cu:21-114) and edits live from the keyboard / SpaceBall (lyap_interactive.cu:144-463:
 * d, jitter, the three thresholds, camera M, stepMethod, settle/accum, depth, refine, gradient,
 * light selection and light pose), as one plain struct plus a text file format.
 *
 * File format: one `key = value ...` per line, `#` starts a comment, keys are the reference's own
 * field names.  A file only needs the keys it changes; everything else keeps the params_init()
 * default.  Floats are written with 9 significant digits, so save -> load reproduces every bit.
 *
 *   sequence = BCABA                  width = 3840          height = 2160
 *   d settle accum stepMethod nearThreshold nearMultiplier opaqueThreshold chaosThreshold
 *   depth jitter refine gradient lMin lMax                       (LyapParams, structs.hpp:53-68)
 *   cam.C = x y z    cam.Q = x y z w    cam.M = m                (LyapCam inputs, structs.hpp:26-36)
 *   cam.orbit = i                     camera of scale.pl's path at eased position i in [0,1]
 *   cam.orbit_frame = f n             ... of frame f of an n-frame orbit
 *   lights = n                        number of lights in use (0..16)
 *   lightK.C  lightK.Q  lightK.M  lightK.lightRange  lightK.ambient = r g b a  lightK.diffuseColor
 *   lightK.diffusePower  lightK.specularColor  lightK.specularPower  lightK.specularHardness
 *   lightK.chaosColor                                            (K = 0..15)
 *
 * Derived fields (cam.V/S0/SDX/SDY, render sizes, light V and cone cosines) are never read from a
 * file: lyap_scene_finalize() recomputes them exactly as the reference's update_scene() does
 * (scene_lights_recalculate + scene_cam_recalculate, lyap_interactive.cu:124-136).
 */
#ifndef LYAP_SCENE_H
#define LYAP_SCENE_H

#include "lyap/types.h"

#ifdef __cplusplus
extern "C" {
#endif

enum { LYAP_SCENE_SEQUENCE_CAP = 1024 };

typedef struct lyap_scene {
    lyap_params prm;
    lyap_cam cam;
    lyap_light lights[LYAP_MAX_LIGHTS];
    uint32_t num_lights;
    uint32_t width, height;                    /* default_imageWidth / Height, params.cu:17-18 */
    char sequence[LYAP_SCENE_SEQUENCE_CAP];    /* as written: "A6B6C6" */
} lyap_scene;

/* params_init(): the reference's defaults. */
void lyap_scene_defaults(lyap_scene *sc);
/* Apply the `key = value` lines of `text` on top of *sc.  Returns LYAP_OK, or LYAP_ERR_BAD_ARGUMENT
 * with a message ("line 12: unknown key 'foo'") in err. */
int lyap_scene_parse(lyap_scene *sc, const char *text, char *err, size_t err_cap);
/* Defaults, then the file.  LYAP_ERR_IO if it cannot be read. */
int lyap_scene_load(lyap_scene *sc, const char *path, char *err, size_t err_cap);
/* Recompute every derived field for a width x height frame (0 = the scene's own size). */
void lyap_scene_finalize(lyap_scene *sc, uint32_t width, uint32_t height);
/* Write the scene's inputs; returns the number of characters (excluding the NUL) the full text
 * needs, like snprintf. */
size_t lyap_scene_format(const lyap_scene *sc, char *out, size_t cap);
int lyap_scene_save(const lyap_scene *sc, const char *path);

#ifdef __cplusplus
}
#endif

#endif /* LYAP_SCENE_H */

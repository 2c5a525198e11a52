"""lyapunov3d_b200 -- B200-native (sm_100a) hot path of tomgidden/lyapunov3d.

Per-sample Lyapunov exponent, ray march / refine / normal / shade, and the voxel
bake, as hand-written CUDA behind a C ABI (include/lyap/abi.h).  See DESIGN.md.
"""
from . import api, structs  # noqa: F401
from .api import (MODE_EXACT, MODE_FAST, MODE_HOST, LyapError, bake, bake_host, campath_frame, campath_orbit,  # noqa: F401
                  exponent_points, params_init, probe_peaks, render, render_host, scene_cam_recalculate,
                  scene_convert_sequence, scene_lights_recalculate, shade_points)

__version__ = "0.1.0"

"""The drop-in claim, compiled: apply integration/lyap_calculate.patch and integration/lyap_interactive.patch
to scratch copies of the reference's two programs, build them with the reference's own scene.cu / params.cu
and link them against liblyap_b200.so (lyap_interactive.cu against the headless GL/GLUT stand-ins of
integration/headless_gl_stubs).  Needs /root/reference and nvcc (no GPU): skipped elsewhere."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
NVCC = "/usr/local/cuda/bin/nvcc"


@pytest.mark.skipif(not (os.path.exists(os.path.join(REF, "lyap_calculate.cu")) and os.path.exists(NVCC) and shutil.which("patch")),
                    reason="needs the reference sources, nvcc and patch")
def test_patched_reference_program_builds_against_the_library(tmp_path):
    import lyapunov3d_b200.api as api
    api.lib()                                                        # the library must exist
    shutil.copy(os.path.join(REF, "lyap_calculate.cu"), tmp_path / "lyap_calculate.cu")
    r = subprocess.run(["patch", "-p1", "-i", os.path.join(ROOT, "integration", "lyap_calculate.patch")], cwd=tmp_path,
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    src = (tmp_path / "lyap_calculate.cu").read_text()
    assert "lyap_bake(" in src and "kernel_calc_volume<<<" not in src
    exe = tmp_path / "lyap_calculate_b200"
    cmd = [NVCC, "-ccbin", "/usr/bin/g++", "--use_fast_math", "-arch=sm_100", "-w",
           "-I", os.path.join(ROOT, "integration", "helper_stubs"), "-I", REF, "-I", os.path.join(ROOT, "include"),
           str(tmp_path / "lyap_calculate.cu"), os.path.join(REF, "scene.cu"), os.path.join(REF, "params.cu"),
           "-L", os.path.join(ROOT, "lyapunov3d_b200"), "-llyap_b200",
           "-Xlinker", "-rpath," + os.path.join(ROOT, "lyapunov3d_b200"), "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert exe.exists()
    nm = subprocess.run(["nm", "-D", "--undefined-only", str(exe)], capture_output=True, text=True).stdout
    assert "lyap_bake" in nm and "lyap_error_string" in nm          # resolved from liblyap_b200.so at run time


@pytest.mark.skipif(not (os.path.exists(os.path.join(REF, "lyap_interactive.cu")) and os.path.exists(NVCC) and shutil.which("patch")),
                    reason="needs the reference sources, nvcc and patch")
def test_patched_lyap_interactive_builds_against_the_library(tmp_path):
    """The render call site, lyap_interactive.cu:711: the patch is the include plus the launch line, and
    keeps passing the reference's own device-resident cudaSeq and its LyapCam / LyapParams objects."""
    import lyapunov3d_b200.api as api
    api.lib()
    shutil.copy(os.path.join(REF, "lyap_interactive.cu"), tmp_path / "lyap_interactive.cu")
    patch = os.path.join(ROOT, "integration", "lyap_interactive.patch")
    changed = [l for l in open(patch).read().splitlines() if l[:1] in "+-" and l[:3] not in ("+++", "---")]
    assert len([l for l in changed if l.startswith("-")]) == 1 and "kernel_calc_render<<<" in [l for l in changed if l.startswith("-")][0]
    r = subprocess.run(["patch", "-p1", "-i", patch], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    src = (tmp_path / "lyap_interactive.cu").read_text()
    assert "lyap_render(" in src and "kernel_calc_render<<<" not in src and "cudaSeq, (const lyap_light *)cudaLights" in src
    exe = tmp_path / "lyap_interactive_b200"
    cmd = [NVCC, "-ccbin", "/usr/bin/g++", "--use_fast_math", "-arch=sm_100", "-w", "-DOUTPUT_PPM", "-DOUTPUT_POINTS",
           "-I", os.path.join(ROOT, "integration", "headless_gl_stubs"), "-I", os.path.join(ROOT, "integration", "helper_stubs"),
           "-I", REF, "-I", os.path.join(ROOT, "include"),
           str(tmp_path / "lyap_interactive.cu"), os.path.join(REF, "scene.cu"), os.path.join(REF, "params.cu"),
           "-L", os.path.join(ROOT, "lyapunov3d_b200"), "-llyap_b200",
           "-Xlinker", "-rpath," + os.path.join(ROOT, "lyapunov3d_b200"), "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    nm = subprocess.run(["nm", "-D", "--undefined-only", str(exe)], capture_output=True, text=True).stdout
    assert "lyap_render" in nm and "glut" not in nm.lower()          # the GL side is the headless stand-in

"""Repository-level contracts that need no GPU: the reference arm of bench.py runs on the CPU,
the product package never touches the oracle, and every CUDA source is built by build()."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_reference_arm_runs_on_cpu():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-rows", "1"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Giter/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["gpu_launches"] == 0


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "lyapunov3d_b200")
    offenders = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(d, f), errors="ignore").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|#include\s+[\"<].*oracle|liblyap_oracle|libref_", text, re.M):
                    offenders.append(os.path.relpath(os.path.join(d, f), ROOT))
    assert not offenders, offenders


def test_every_cuda_source_is_in_the_build():
    from lyapunov3d_b200 import _build
    cu = set()
    for d, _, files in os.walk(_build.CSRC):
        cu |= {os.path.relpath(os.path.join(d, f), _build.CSRC) for f in files if f.endswith(".cu")}
    assert cu == set(_build.CU), cu ^ set(_build.CU)
    assert "arch=compute_100a,code=sm_100a" in " ".join(_build.NVCC_FLAGS) and "-lineinfo" in _build.NVCC_FLAGS


def test_every_option_knob_is_documented_in_the_header():
    """lyap_set_option's keys (csrc/abi.cu) all appear in include/lyap/abi.h's description of the knobs."""
    import re
    src = open(os.path.join(ROOT, "lyapunov3d_b200", "csrc", "abi.cu")).read()
    hdr = open(os.path.join(ROOT, "include", "lyap", "abi.h")).read()
    keys = re.findall(r'strcmp\(key, "([a-z_0-9]+)"\)', src)
    assert len(keys) >= 10
    missing = [k for k in keys if '"%s"' % k not in hdr]
    assert not missing, missing

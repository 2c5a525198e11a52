"""Build liblyap_b200.so in-tree with nvcc for sm_100a (no GPU needed to compile).

    python -m lyapunov3d_b200._build [--force] [--verbose]

One object per translation unit, compiled in parallel; host C++ goes through the
distro g++ with -ffp-contract=off (the scene layer must round like the reference's
host build).  The .so is git-ignored but travels to the GPU box with the snapshot.
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(PKG, "liblyap_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOSTCXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-std=c++17", "-lineinfo", "-fmad=false", "-ccbin", HOSTCXX,
                     "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
                     "-I", os.path.join(ROOT, "include"), "-I", CSRC]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"), "-I", CSRC]

CU = (["abi.cu", "kernels/misc.cu"] + [f"kernels/tu_{k}_{m}.cu" for k in ("render", "bake") for m in ("exact", "fast", "host")]
      + ["kernels/tu_march_exact.cu", "kernels/tu_march_host.cu"])
CPP = ["host/scene.cpp", "host/scenefile.cpp", "host/imageio.cpp"]
APPS = ["lyap_render", "lyap_calculate"]
BIN = os.path.join(PKG, "bin")


def _sources():
    deps = []
    for d, _, files in os.walk(CSRC):
        deps += [os.path.join(d, f) for f in files]
    deps += [os.path.join(ROOT, "include", "lyap", f) for f in os.listdir(os.path.join(ROOT, "include", "lyap"))]
    return deps


def up_to_date():
    if not os.path.exists(LIB):
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(s) <= t for s in _sources() + [os.path.abspath(__file__)])


def _run(cmd, verbose):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("build step failed: " + " ".join(cmd[:3]) + " ...")
    return r.stdout + r.stderr


def build(force=False, verbose=False, ptxas_info=False):
    if not force and up_to_date():
        return LIB
    if not shutil.which(NVCC) and not os.path.exists(NVCC):
        raise RuntimeError("nvcc not found; liblyap_b200.so cannot be built here")
    os.makedirs(OBJ, exist_ok=True)
    jobs = []
    for src in CU:
        obj = os.path.join(OBJ, src.replace("/", "_") + ".o")
        extra = ["-Xptxas", "-v"] if ptxas_info else []
        jobs.append((obj, [NVCC] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]))
    for src in CPP:
        obj = os.path.join(OBJ, src.replace("/", "_") + ".o")
        jobs.append((obj, [HOSTCXX] + CXX_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]))
    logs = []
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        for out in ex.map(lambda j: _run(j[1], verbose), jobs):
            logs.append(out)
    _run([NVCC] + ARCH + ["-shared", "-ccbin", HOSTCXX, "-o", LIB] + [j[0] for j in jobs] + ["-lz"], verbose)
    # headless apps on top of the C ABI (the reference's two programs, re-hosted)
    os.makedirs(BIN, exist_ok=True)
    for app in APPS:
        _run([HOSTCXX, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(CSRC, "apps", app + ".cpp"),
              "-o", os.path.join(BIN, app), "-L", PKG, "-llyap_b200", "-Wl,-rpath,$ORIGIN/.."], verbose)
    if ptxas_info:
        return "\n".join(logs)
    return LIB


if __name__ == "__main__":
    out = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, ptxas_info="--ptxas" in sys.argv)
    print(out)

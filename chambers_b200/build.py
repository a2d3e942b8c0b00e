"""In-tree build of ``libchambers_aug.so`` (sm_100a only) with nvcc.

``python -m chambers_b200.build`` or ``__graft_entry__.build()``.  The ``.so`` stays
inside the package directory (git-ignored) so that it travels with the tree.
"""

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libchambers_aug.so")
SOURCES = ["chb_kernels_c1.cu", "chb_kernels_c2.cu", "chb_kernels_c3.cu", "chb_kernels_c4.cu",
           "chb_resident_c1.cu", "chb_resident_c2.cu", "chb_resident_c3.cu", "chb_resident_c4.cu",
           "chb_plan.cu", "chb_launch.cu", "chb_frontend.cu", "chb_api.cu"]
HEADERS = [os.path.join(CSRC, "chb_internal.h"), os.path.join(CSRC, "chb_kernels.cuh"),
           os.path.join(CSRC, "chb_device.cuh"), os.path.join(CSRC, "chb_resident.cuh"),
           os.path.join(ROOT, "include", "chambers_aug.h")]
OBJ_DIR = os.path.join(PKG_DIR, "build")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",  # belt and braces: TF's CPU kernels round after every op; we also use *_rn intrinsics
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "-Xptxas", "-v",
]


def find_nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libchambers_aug.so cannot be built")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def build_variant(name, defines):
    """Debug / experiment build next to the production library: libchambers_aug_<name>.so compiled
    with extra -D flags (e.g. CHB_TIMELINE).  Select it at run time with CHB_LIB=<path>."""
    out = os.path.join(PKG_DIR, "libchambers_aug_%s.so" % name)
    extra = [d if d.startswith("-") else "-D" + d for d in defines]  # raw nvcc flags pass through
    return build_library(force=True, out_path=out, extra=extra, obj_tag="_" + name)


def build_library(force=False, verbose=False, out_path=None, extra=(), obj_tag=""):
    """Compile the CUDA extension if it is missing or stale; returns the .so path."""
    if not force and not needs_build():
        return LIB_PATH
    out_path = out_path or LIB_PATH
    nvcc = find_nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)
    inc = ["-I", os.path.join(ROOT, "include"), "-I", CSRC]
    procs = []
    for src in SOURCES:  # one nvcc per translation unit, all in parallel
        obj = os.path.join(OBJ_DIR, src.replace(".cu", obj_tag + ".o"))
        cmd = [nvcc] + NVCC_FLAGS + list(extra) + inc + ["-c", "-o", obj, os.path.join(CSRC, src)]
        procs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log, failed, objs = "", False, []
    for cmd, obj, pr in procs:
        out, _ = pr.communicate()
        log += " ".join(cmd) + "\n" + out + "\n"
        failed = failed or pr.returncode != 0
        objs.append(obj)
    if not failed:
        cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out_path] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        log += " ".join(cmd) + "\n" + res.stdout + res.stderr
        failed = res.returncode != 0
    with open(os.path.join(PKG_DIR, "build%s.log" % obj_tag), "w") as f:
        f.write(log)
    if failed:
        raise RuntimeError("nvcc failed:\n" + log[-8000:])
    if verbose:
        print(log)
    return out_path


if __name__ == "__main__":
    if "--variant" in sys.argv:  # python -m chambers_b200.build --variant timeline CHB_TIMELINE [MORE_DEFINES...]
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build_library(force="--force" in sys.argv, verbose=True))

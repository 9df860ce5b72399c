"""Builds libmythtracer_b200.so (hand-written CUDA kernels + C ABI) in-tree for sm_100a with nvcc.

--fmad=false / -ffp-contract=off: the reference binary contains no FMA and every
deciding computation is reproduced bit for bit, so contraction must stay off on both device and host.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmythtracer_b200.so")
SOURCES = ["api.cu", "megakernel.cu", "wavefront.cu", "device_build.cu", "scene_build.cc", "obj_loader.cc", "image_decode.cc", "jpeg_decode.cc"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2,-Wall", "-D_USE_MATH_DEFINES",
]


def _newest_source_mtime() -> float:
    paths = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "mythtracer_b200.h")]
    return max(os.path.getmtime(p) for p in paths)


COMMIT_STAMP = os.path.join(HERE, "build", "commit.txt")


def stamp_commit() -> None:
    """Records the git commit of the tree the library was built from (there is no .git on the GPU box: the stamp
    travels with the built library and bench.py reports it as config.commit)."""
    try:
        head = subprocess.check_output(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], stderr=subprocess.DEVNULL, text=True).strip()
        dirty = subprocess.check_output(["git", "-C", ROOT, "status", "--porcelain", "--", "mythtracer_b200", "include", "bench.py"],
                                        stderr=subprocess.DEVNULL, text=True).strip()
    except Exception:
        return
    os.makedirs(os.path.dirname(COMMIT_STAMP), exist_ok=True)
    with open(COMMIT_STAMP, "w") as f:
        f.write(head + ("+" if dirty else "") + "\n")


def build(force: bool = False, verbose: bool = False, extra=(), variant: str = "") -> str:
    """variant: development A/B builds (extra -D flags) go to build/var_<variant>/lib.so; load with MTB_LIB_PATH."""
    if not variant:
        stamp_commit()
    lib_path = LIB if not variant else os.path.join(HERE, "build", "var_" + variant, "lib.so")
    if not force and os.path.exists(lib_path) and os.path.getmtime(lib_path) >= _newest_source_mtime():
        return lib_path
    nvcc = os.environ.get("NVCC", "nvcc")
    objs = []
    build_dir = os.path.join(HERE, "build") if not variant else os.path.join(HERE, "build", "var_" + variant)
    os.makedirs(build_dir, exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(build_dir, src + ".o")
        cmd = [nvcc] + NVCC_FLAGS + list(extra) + ["-I", os.path.join(ROOT, "include"), "-I", CSRC, "-x", "cu" if src.endswith(".cu") else "c++",
                                                  "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
        objs.append(obj)
    cmd = [nvcc, "-shared", "-o", lib_path] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return lib_path


if __name__ == "__main__":
    if "--variant" in sys.argv:  # python build.py --variant NAME -DFOO=1 -DBAR=2
        i = sys.argv.index("--variant")
        print(build(force=True, extra=[a for a in sys.argv[i + 2:]], variant=sys.argv[i + 1]))
    else:
        print(build(force="--force" in sys.argv, verbose=True, extra=["-Xptxas", "-v"] if "--ptxas" in sys.argv else ()))

"""Build the C-ABI CUDA library in-tree:  python -m yastn_b200.build

Compiles yastn_b200/csrc/*.cu for sm_100a with nvcc into yastn_b200/libyastn_b200.so (cross-compiles
without a GPU).  The .so is git-ignored but travels with the repo snapshot to the GPU box.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libyastn_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + [os.path.join(HERE, "..", "include", "yastn_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


FLATTEN_SRC = os.path.join(CSRC, "yb_flatten.c")
FLATTEN_LIB = os.path.join(HERE, "_flatten.so")


def build_flatten(force=False):
    """The CPython helper that flattens YASTN's nested meta tuples (host side of plan construction)."""
    if not force and os.path.exists(FLATTEN_LIB) and os.path.getmtime(FLATTEN_LIB) >= os.path.getmtime(FLATTEN_SRC):
        return FLATTEN_LIB
    import sysconfig
    inc = sysconfig.get_paths()["include"]
    r = subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-I", inc, FLATTEN_SRC, "-o", FLATTEN_LIB], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("gcc failed on yb_flatten.c")
    return FLATTEN_LIB


def build(force=False, verbose=False):
    build_flatten(force)
    if not force and not needs_build():
        return LIB
    hdr_t = max(os.path.getmtime(d) for d in glob.glob(os.path.join(CSRC, "*.h")) + [os.path.join(HERE, "..", "include", "yastn_b200.h")])

    def compile_one(src):
        obj = os.path.join(CSRC, os.path.basename(src)[:-3] + ".o")
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_t):
            return obj, ""                       # up to date
        r = subprocess.run([NVCC] + FLAGS + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed on {src}")
        return obj, r.stderr
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:      # the translation units are independent
        done = list(ex.map(compile_one, sources()))
    objs = [o for o, _ in done]
    log = [l for _, l in done]
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    if all(log):            # a partial rebuild keeps the register / spill report of the last full build
        with open(os.path.join(CSRC, "ptxas.log"), "w") as f:
            f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))

"""Builds lib/libsph3d_b200.so (hand-written sm_100a CUDA behind the C ABI of include/sph3d_b200.h).

Replaces the reference's per-op build scripts (/root/reference/compile.sh and
tf_ops/*/tf_*_compile.sh: one `nvcc -c` + one `g++ -shared` against TensorFlow per op, no -arch
flag).  Here: one nvcc invocation per translation unit, sm_100a only, -lineinfo so ncu's source
page maps to these files, no fast-math (bit-exact sqrt/div/atan2f paths), in-tree output so the
.so travels with the repository snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsph3d_b200.so")
SOURCES = ["nnquery.cu", "buildkernel.cu", "conv_fwd.cu", "conv_bwd.cu", "conv_bwd_t.cu", "pool3d.cu", "sample.cu", "post.cu",
           "sepconv.cu", "rowsgemm.cu", "rowswgrad.cu"]
HEADERS = ["common.cuh", "rowwarp.cuh", "conv_common.cuh", "tc05.cuh", os.path.join("..", "..", "include", "sph3d_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def is_stale():
    if not os.path.exists(LIB):
        return True
    deps = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    return _newest(deps) > os.path.getmtime(LIB)


def build(force=False, verbose=False):
    """Compile every CUDA translation unit for sm_100a and link the shared library.  Objects are kept under lib/obj
    (git-ignored) so that only the translation units older than their source or any header are recompiled."""
    if not force and not is_stale():
        return LIB
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    def deps_time(source):
        return _newest([os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS] + [os.path.join(CSRC, source)])
    objs = []
    procs = []
    for s in SOURCES:
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(obj)
        src = os.path.join(CSRC, s)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= deps_time(s):
            continue
        cmd = [_nvcc()] + NVCC_FLAGS + ["-c", "-o", obj, src]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed: %s\n%s" % (" ".join(cmd), out.decode(errors="replace")))
    link = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise RuntimeError("link failed: %s\n%s" % (" ".join(link), r.stdout.decode(errors="replace")))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

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
           "sepconv.cu", "rowsgemm.cu",
           "dense_nn.cu", "dense_nt.cu", "dense_tn.cu", "dense_nn2.cu", "dense_nt2.cu", "dense_abi.cu"]
HEADERS = ["common.cuh", "rowwarp.cuh", "conv_common.cuh", "tc05.cuh", "dense_gemm.cuh", os.path.join("..", "..", "include", "sph3d_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC"]


def cutlass_root():
    """CuTe/CUTLASS header tree vendored in the image (no /opt/cutlass): the dense_*.cu translation units instantiate
    its sm_100 tcgen05 collectives.  Returns the directory that holds include/ and tools/util/include, or None."""
    import importlib.util
    cands = []
    env = os.environ.get("SPH3D_CUTLASS_ROOT")
    if env:
        cands.append(env)
    for pkg, rel in (("flashinfer", os.path.join("data", "cutlass")), ("tilelang", os.path.join("3rdparty", "cutlass"))):
        try:
            spec = importlib.util.find_spec(pkg)
        except Exception:
            spec = None
        if spec and spec.submodule_search_locations:
            cands.append(os.path.join(list(spec.submodule_search_locations)[0], rel))
    for c in cands:
        if os.path.exists(os.path.join(c, "include", "cutlass", "gemm", "collective", "builders", "sm100_9xBF16_umma_builder.inl")) \
                and os.path.exists(os.path.join(c, "tools", "util", "include", "cutlass", "util", "packed_stride.hpp")):
            return c
    return None


def _extra_flags(source):
    if not source.startswith("dense_"):
        return []
    root = cutlass_root()
    if root is None:
        # No silent downgrade: without the header tree the tcgen05 pointwise products cannot be built and the layer
        # library would quietly fall back to the library GEMM.  Point SPH3D_CUTLASS_ROOT at a CUTLASS >= 4.x checkout
        # (the directory that holds include/ and tools/util/include), or opt out explicitly with SPH3D_NO_CUTLASS=1:
        # the dense entry points then exist and return cudaErrorNotSupported (801).
        if os.environ.get("SPH3D_NO_CUTLASS") == "1":
            return ["-DSPH3D_NO_CUTLASS"]
        raise RuntimeError("sph3d-gcn_b200: CuTe/CUTLASS headers not found (looked at $SPH3D_CUTLASS_ROOT and the trees "
                           "vendored in the flashinfer / tilelang packages). Set SPH3D_CUTLASS_ROOT, or SPH3D_NO_CUTLASS=1 "
                           "to build without the tcgen05 pointwise products.")
    return ["-I", os.path.join(root, "include"), "-I", os.path.join(root, "tools", "util", "include"),
            "--expt-relaxed-constexpr", "-w"]


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
        # the CUTLASS instantiations (two minutes each) include nothing of this library but dense_gemm.cuh
        hdrs = ["dense_gemm.cuh"] if source.startswith("dense_") and source != "dense_abi.cu" else HEADERS
        return _newest([os.path.normpath(os.path.join(CSRC, h)) for h in hdrs] + [os.path.join(CSRC, source)])
    objs = []
    procs = []
    for s in SOURCES:
        obj = os.path.join(objdir, s.replace(".cu", ".o"))
        objs.append(obj)
        src = os.path.join(CSRC, s)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= deps_time(s):
            continue
        cmd = [_nvcc()] + NVCC_FLAGS + _extra_flags(s) + ["-c", "-o", obj, src]
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

"""Builds openpbso_b200/libpbso_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m openpbso_b200.build [--force]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libpbso_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function", "--use_fast_math=false"]
FLAGS = [f for f in FLAGS if not f.startswith("--use_fast_math")]   # precise math only


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "pbso_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        objs.append(obj)
        cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(out)
            raise RuntimeError("nvcc failed on %s" % src)
        if verbose:
            sys.stdout.write(out)
    subprocess.check_call([NVCC, "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ldl"])
    return OUT


def _build_tool(name):
    root = os.path.dirname(HERE)
    inc = os.path.join(root, "include", "openpbso")
    src = os.path.join(root, "tools", name + ".cpp")
    out = os.path.join(HERE, "bin", name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [src, OUT] + [os.path.join(root, "tools", f) for f in os.listdir(os.path.join(root, "tools")) if f.endswith(".h")]
    if os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(d) for d in deps):
        return out
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-pthread", "-I" + os.path.join(inc, "eigen_shim"), "-I" + inc, src,
                           "-L" + HERE, "-lpbso_b200", "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/..", "-o", out])
    return out


def build_tools():
    """tools/pbso_render.cpp (headless driver) and tools/pbso_fit_ffat.cpp (headless FFAT map
    construction, rank 3) -> openpbso_b200/bin/, against the header mirror.  Returns the path of pbso_render."""
    _build_tool("pbso_fit_ffat")
    return _build_tool("pbso_render")


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_tools())

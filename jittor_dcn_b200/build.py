"""In-tree build of libdcn_b200.so (nvcc, sm_100a only).

    python -m jittor_dcn_b200.build [--force]

The shared library is written next to this file (jittor_dcn_b200/libdcn_b200.so): it is
git-ignored but travels to the GPU box with the repo snapshot.  Objects go to csrc/_build/.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libdcn_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    # coordinates must be bit-exact: no fast-math, IEEE division/sqrt, denormals kept
    "-ftz=false", "-prec-div=true", "-prec-sqrt=true", "-fmad=true",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA engine cannot be built")


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith((".cu", ".cpp")))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, debug=False):
    """debug=True: libdcn_b200_dbg.so with -DDCN_DEBUG_CHECKS (device-side bounds / pipeline-agreement checks,
    csrc/dcn_common.cuh); tests and tools pick it with DCN_B200_LIB=dbg."""
    global OBJ, LIB
    if debug:
        OBJ, LIB = os.path.join(CSRC, "_build_dbg"), os.path.join(HERE, "libdcn_b200_dbg.so")
    else:
        OBJ, LIB = os.path.join(CSRC, "_build"), os.path.join(HERE, "libdcn_b200.so")
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "dcn_b200.h"))
    headers.append(os.path.abspath(__file__))
    objs, procs = [], []
    for src in sources():
        sp = os.path.join(CSRC, src)
        op = os.path.join(OBJ, os.path.splitext(src)[0] + ".o")
        objs.append(op)
        if force or _stale(op, [sp] + headers):
            cmd = [nvcc] + NVCC_FLAGS + ["-DDCN_DEBUG_CHECKS=1"] * bool(debug) + ["-Xptxas", "-v"] * bool(verbose) + \
                ["-c", sp, "-o", op]
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                      "-Xcompiler", "-fPIC", "-ldl"]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, debug="--debug" in sys.argv))

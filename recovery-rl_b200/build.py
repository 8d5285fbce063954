"""Build librrl.so (the sm_100a CUDA kernels + C ABI of include/rrl.h) in-tree with nvcc.

    python recovery-rl_b200/build.py [--force] [--verbose]

env.cu / replay.cu are compiled with -fmad=false: their fp64 arithmetic must be bit-identical to the
reference's numpy expressions, so every contraction is written explicitly.
agent.cu (fp32 MLP kernels) is compiled with FMA contraction enabled.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "librrl.so")
OBJ = os.path.join(HERE, "build")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-I", os.path.join(HERE, "..", "include")]
# (source, extra flags)
SOURCES = [
    ("api.cu", []),
    ("env.cu", ["-fmad=false"]),
    ("replay.cu", ["-fmad=false"]),
    ("agent.cu", []),
    ("agent_tc.cu", []),
    ("select.cu", []),
    ("mpc.cu", []),
    ("mpc_tc.cu", []),
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "rrl.h"))
    headers.append(os.path.abspath(__file__))
    objs = []
    procs = []
    for src, extra in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(obj)
        if force or _stale(obj, [path] + headers):
            cmd = [_nvcc()] + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
            if verbose:
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed = True
            sys.stderr.write("nvcc failed for %s:\n%s\n" % (src, out))
        elif verbose or "warning" in out:
            sys.stdout.write(out)
    if failed:
        raise RuntimeError("librrl.so build failed")
    if force or procs or _stale(OUT, objs):
        cmd = [_nvcc()] + ARCH + ["-shared", "-o", OUT] + objs
        subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

#!/usr/bin/env python
"""Build librtkv.so (sm_100a only) in-tree with nvcc.  No torch headers are involved:
the library is a plain C-ABI shared object (include/rtkv.h).

Every source is compiled to its own object (in parallel, only when it or a header changed) and the
objects are linked into rtkv/librtkv.so."""
import concurrent.futures
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "rtkv", "librtkv.so")
SOURCES = ["capi.cu", "prompt_mass.cu", "prompt_mass_qk.cu", "score_select.cu", "score_select_cluster.cu",
           "quant_gather.cu", "quant_ex.cu", "packed_attention.cu", "head_shard.cu", "flash_prefill.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
    "--fmad=false",            # parity-critical arithmetic: no FMA contraction anywhere
    "-Xptxas", "-v",
]
CCBIN = ["-ccbin", "/usr/bin/g++"]   # the image's CC/CXX point at a wrapper; let nvcc use the system host compiler


def _headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return hs + [os.path.join(HERE, "..", "include", "rtkv.h"), __file__]


def _sources():
    return [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def needs_build():
    return _stale(OUT, [os.path.join(CSRC, s) for s in _sources()] + _headers())


def _compile(nvcc, src, force):
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not force and not _stale(obj, [path] + _headers()):
        return src, 0, "", False
    cmd = [nvcc] + NVCC_FLAGS + CCBIN + ["-c", path, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, res.returncode, " ".join(cmd) + "\n" + res.stdout + res.stderr, True


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    logs, failed = [], False
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        for src, rc, log, ran in ex.map(lambda s: _compile(nvcc, s, force), srcs):
            if ran:
                logs.append(log)
                with open(os.path.join(OBJ, src.replace(".cu", ".log")), "w") as f:
                    f.write(log)
            failed = failed or rc != 0
    # build.log = the latest compile log of EVERY source (ptxas -v register / spill report)
    full = []
    for s in srcs:
        lp = os.path.join(OBJ, s.replace(".cu", ".log"))
        if os.path.exists(lp):
            full.append(open(lp).read())
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write("\n".join(full))
    if failed:
        sys.stderr.write("\n".join(logs))
        raise RuntimeError("nvcc failed building librtkv.so")
    cmd = [nvcc, "-shared"] + CCBIN + ["-o", OUT] + [os.path.join(OBJ, s.replace(".cu", ".o")) for s in srcs]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("linking librtkv.so failed")
    if verbose:
        print("\n".join(logs))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

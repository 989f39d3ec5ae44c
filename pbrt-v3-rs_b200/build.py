"""Builds libb200pt.so (hand-written sm_100a CUDA + the C ABI) in-tree with nvcc.

Usage: python pbrt-v3-rs_b200/build.py [--force] [--verbose]
The .so lands next to this file so it travels to the GPU box with the snapshot.
"""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libb200pt.so")
STAMP = os.path.join(HERE, "build", "stamp.txt")

SOURCES = ["b200pt_api.cu", "traverse_kernels.cu", "b200pt_scene.cu", "instancing.cu", "bvh_build.cu", "multi_gpu.cu", "host_bvh.cpp", "host_hlbvh.cpp", "host_sampler.cpp", "host_envmap.cpp", "host_scene_loader.cpp"]

# -fmad=false + -ffp-contract=off: the reference (Rust f32) never fuses or
# re-associates; traversal parity is bit-exact only without FMA contraction.
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-O2",
    "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "cli"), os.path.join(HERE, "..", "include")):
        for name in sorted(os.listdir(root)):
            with open(os.path.join(root, name), "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _object_digest(src):
    """One translation unit's inputs: the source itself, every header beside it and under include/, and the flags."""
    h = hashlib.sha256()
    names = [os.path.join(CSRC, src)]
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        names += [os.path.join(root, n) for n in sorted(os.listdir(root)) if n.endswith((".h", ".cuh", ".hpp"))]
    for name in names:
        with open(name, "rb") as f:
            h.update(os.path.basename(name).encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile(src, force):
    """Compiles one source unless its object is up to date; returns (object, log text)."""
    obj = os.path.join(HERE, "build", src + ".o")
    stamp = obj + ".stamp"
    digest = _object_digest(src)
    if not force and os.path.exists(obj) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return obj, "(up to date) " + src
    cmd = [_nvcc()] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    text = "$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr
    if r.returncode != 0:
        sys.stderr.write(text)
        raise RuntimeError("nvcc failed on " + src)
    with open(stamp, "w") as f:
        f.write(digest)
    return obj, text


def build(force=False, verbose=False):
    digest = _digest()
    if not force and os.path.exists(OUT) and os.path.exists(STAMP):
        with open(STAMP) as f:
            if f.read().strip() == digest:
                return OUT
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    # translation units are independent: compile them side by side (the two largest take minutes each)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, min(len(SOURCES), os.cpu_count() or 1))) as pool:
        results = list(pool.map(lambda s: _compile(s, force), SOURCES))
    objs = [o for o, _ in results]
    log = [t for _, t in results]
    cmd = [_nvcc(), "-shared", "-o", OUT] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static", "-lpthread", "-ldl", "-lrt", "-lz"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(log[-1])
        raise RuntimeError("link failed")
    # command-line driver: plain C++ over the C ABI (no CUDA on that side)
    cli = os.path.join(HERE, "b200pt_render")
    cmd = [os.environ.get("CXX", "g++"), "-O2", "-std=c++17", os.path.join(HERE, "cli", "b200pt_render.cpp"), "-o", cli, "-L" + HERE, "-lb200pt",
           "-Wl,-rpath,$ORIGIN"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(log[-1])
        raise RuntimeError("building b200pt_render failed")
    with open(os.path.join(HERE, "build", "build.log"), "w") as f:
        f.write("\n".join(log))
    with open(STAMP, "w") as f:
        f.write(digest)
    if verbose:
        print("\n".join(log))
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(OUT)

"""Builds moonshine_b200/lib/libmoonshine.so (C++/CUDA, sm_100a) in-tree.

Usage: python -m moonshine_b200.build [--force]
nvcc cross-compiles without a GPU; the .so travels to the GPU box with the
repo snapshot (it is git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libmoonshine.so")

SOURCES = [
    "gemm_simt.cu", "gemm_tc.cu", "gemm_planes.cu", "attention_tc.cu", "ring_bench.cu", "kernels_misc.cu", "decoder_step.cu", "decoder_step2.cu", "decoder_step3.cu", "decoder_step4.cu", "model.cu",
    "weights.cpp", "tokenizer.cpp", "word_alignment.cpp", "transcriber.cpp", "c_api.cpp",
]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
]


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode()); h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = True) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "moonshine_b200.h"))
    objs = []
    procs = []
    for src in SOURCES:
        sp = os.path.join(CSRC, src)
        obj = os.path.join(OBJDIR, src + ".o")
        stamp = obj + ".sha"
        dig = _digest([sp] + headers)
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
            continue
        cmd = [NVCC] + FLAGS + ["-x", "cu", "-c", sp, "-o", obj]
        if verbose:
            print("[build]", " ".join(cmd), flush=True)
        procs.append((subprocess.Popen(cmd), stamp, dig, src))
    failed = []
    for p, stamp, dig, src in procs:
        if p.wait() != 0:
            failed.append(src)
        else:
            with open(stamp, "w") as f:
                f.write(dig)
    if failed:
        raise RuntimeError(f"nvcc failed for {failed}")
    link_stamp = LIB + ".link"
    link_cmd_id = "soname-v1"
    if procs or force or not os.path.exists(LIB) or not os.path.exists(link_stamp) or open(link_stamp).read() != link_cmd_id:
        tmp = LIB + ".tmp"  # link beside, then rename: a snapshot taken mid-build never sees a half-written library
        cmd = [NVCC, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                      "-lcudart", "-Xlinker", "--no-undefined",
                                                      # SONAME = the name every reference binding dlopens
                                                      "-Xlinker", "-soname=libmoonshine.so"]
        if verbose:
            print("[build]", " ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        os.replace(tmp, LIB)
        with open(link_stamp, "w") as f:
            f.write(link_cmd_id)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))

"""Builds oracle/_ref/libmoonshine_ref_helpers.so from the reference's OWN sources, compiled where they
lie under /root/reference (never copied): the host-side helpers of the transcription path that have no
third-party dependency -- bin-tokenizer, resampler, word-alignment, context-biaser, voice-activity-detector
(its Silero network replaced by a constant-probability stand-in in ref_shim.cpp) (+ the small utils they call).  The model arithmetic itself cannot be built this way: it lives in ONNX Runtime graphs that
are not in the tree (see oracle/moonshine_oracle.py).

    python oracle/build_ref.py            # no-op when /root/reference is absent (the GPU box)

TEST INFRASTRUCTURE ONLY: the library is loaded by tests/ to pin this repo's restatements.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MOONSHINE_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
LIB = os.path.join(OUT, "libmoonshine_ref_helpers.so")
SOURCES = [
    "core/word-alignment.cpp", "core/bin-tokenizer/bin-tokenizer.cpp", "core/resampler.cpp",
    "core/context-biaser.cpp", "core/context-extractor.cpp", "core/voice-activity-detector.cpp", "core/moonshine-utils/string-utils.cpp", "core/moonshine-utils/debug-utils.cpp",
    "core/moonshine-utils/file-utils.cpp",
]


BENCH = os.path.join(OUT, "ref_benchmark")


def build_benchmark(force: bool = False):
    """The reference's own C++ caller of the C ABI -- core/benchmark.cpp through the header-only
    core/moonshine-cpp.h -- compiled from the sources where they lie and linked against THIS repo's
    libmoonshine.so (rpath relative to the binary, so it runs from the snapshot on the GPU box).  Proves
    that a reference-side C++ client links and runs unchanged (tests/test_reference_bindings_gpu.py)."""
    product = os.path.join(os.path.dirname(HERE), "moonshine_b200", "lib")
    if not os.path.isdir(os.path.join(REF, "core")) or not os.path.exists(os.path.join(product, "libmoonshine.so")):
        return BENCH if os.path.exists(BENCH) else None
    srcs = [os.path.join(REF, "core/benchmark.cpp"), os.path.join(REF, "core/moonshine-utils/file-utils.cpp")]
    deps = srcs + [os.path.join(REF, "core/moonshine-cpp.h"), os.path.join(REF, "core/moonshine-c-api.h")]
    if not force and os.path.exists(BENCH) and all(os.path.getmtime(BENCH) >= os.path.getmtime(s) for s in deps):
        return BENCH
    os.makedirs(OUT, exist_ok=True)
    cmd = ["g++", "-std=c++20", "-O2", "-o", BENCH, f"-I{REF}/core", f"-I{REF}/core/moonshine-utils"] + srcs + [
        f"-L{product}", "-lmoonshine", "-Wl,-rpath,$ORIGIN/../../moonshine_b200/lib", "-Wl,--no-undefined", "-pthread"]
    subprocess.run(cmd, check=True)
    return BENCH


def build(force: bool = False):
    if not os.path.isdir(os.path.join(REF, "core")):
        return LIB if os.path.exists(LIB) else None
    srcs = [os.path.join(REF, s) for s in SOURCES] + [os.path.join(HERE, "ref_shim.cpp")]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    os.makedirs(OUT, exist_ok=True)
    inc = [f"-I{REF}/core", f"-I{REF}/core/bin-tokenizer", f"-I{REF}/core/moonshine-utils",
           f"-I{REF}/core/third-party/onnxruntime/include"]  # headers only: silero-vad.h names ORT types
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-o", LIB] + inc + srcs + ["-Wl,--no-undefined"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build("--force" in sys.argv))
    print(build_benchmark("--force" in sys.argv))

"""TEST INFRASTRUCTURE: builds tests/simt/libdre_emu.so -- the CUDA sources of the product (kernels AND the C-ABI
layer context.cu) compiled by g++ against the host-side SIMT emulator (tests/simt/stub/).  The library exports the
kernel-level harness (emu_*) and the whole C ABI of include/dre_b200.h (dre_*) running on the emulator.  Used only by
tests/test_simt_*.py; the product loads libdre_b200.so and nothing else."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "differentialriccatiequations.jl_b200", "csrc")
LIB = os.path.join(HERE, "libdre_emu.so")
KERNEL_SOURCES = ["sparse_kernels.cu", "dense_kernels.cu", "context.cu"]
DEPS = [os.path.join(CSRC, f) for f in KERNEL_SOURCES + ["symbolic.cpp", "symbolic.h", "kernels.h", "common.cuh",
                                                           "schedule.h"]] + \
       [os.path.join(ROOT, "include", "dre_b200.h"), os.path.join(HERE, "emu_harness.cpp"),
        os.path.join(HERE, "stub", "cuda_runtime.h"), os.path.join(HERE, "stub", "cusolverDn.h")]
FLAGS = ["-O1", "-g", "-std=c++17", "-fPIC", "-pthread", "-DDRE_SIMT_EMU", "-fvisibility=hidden",
         "-fno-strict-aliasing",
         "-I", os.path.join(HERE, "stub"), "-I", CSRC]


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in DEPS):
        return LIB
    objs = []
    procs = []
    for src, lang in [(os.path.join(CSRC, f), "c++") for f in KERNEL_SOURCES] + \
                     [(os.path.join(CSRC, "symbolic.cpp"), "c++"), (os.path.join(HERE, "emu_harness.cpp"), "c++")]:
        o = os.path.join(HERE, "_" + os.path.splitext(os.path.basename(src))[0] + ".emu.o")
        procs.append(subprocess.Popen(["g++"] + FLAGS + ["-x", lang, "-c", src, "-o", o]))
        objs.append(o)
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("g++ failed while building the SIMT emulator library")
    subprocess.run(["g++", "-shared", "-pthread", "-o", LIB] + objs, check=True)
    return LIB


if __name__ == "__main__":
    import sys

    print(build(force="--force" in sys.argv))

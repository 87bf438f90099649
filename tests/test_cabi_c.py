"""The C ABI driven from plain C (tests/c_abi_smoke.c): dre_create -> set_pencil -> set_operator -> adi_step ->
ldlt_norm -> download, checked on the host with CSC mat-vecs.  The same program is run against the product
library on the GPU (-m gpu) and, in the CPU tier, against the SIMT-emulator build of the same sources."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c_abi_smoke.c")


def _build(libpath, out):
    libdir, libname = os.path.split(libpath)
    cmd = ["gcc", "-O2", "-Wall", "-o", out, SRC, libpath, "-lm", f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True)
    return out


def build_product_binary():
    """Called by __graft_entry__.build(): the binary travels to the GPU box with the repository snapshot."""
    lib = os.path.join(ROOT, "differentialriccatiequations.jl_b200", "libdre_b200.so")
    return _build(lib, os.path.join(ROOT, "tests", "c_abi_smoke.bin"))


@pytest.mark.gpu
def test_c_abi_smoke_on_gpu():
    exe = os.path.join(ROOT, "tests", "c_abi_smoke.bin")
    lib = os.path.join(ROOT, "differentialriccatiequations.jl_b200", "libdre_b200.so")
    if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(lib):
        build_product_binary()
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    print(res.stdout, res.stderr)
    assert res.returncode == 0 and "C_ABI_SMOKE_OK" in res.stdout


def test_c_abi_smoke_on_emulator(tmp_path):
    from tests.simt import build_emu

    lib = build_emu.build()
    exe = _build(lib, str(tmp_path / "c_abi_smoke_emu"))
    res = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    print(res.stdout, res.stderr)
    assert res.returncode == 0 and "C_ABI_SMOKE_OK" in res.stdout

"""Run a script of this repository against the EMULATED C ABI (tests/simt/libdre_emu.so: the product's CUDA sources
on the host-side SIMT emulator) instead of libdre_b200.so -- a development aid for machines without a GPU:

    python tools/emu_run.py bench.py --n 371 --steps 1 --warmup 1 --no-cpu --no-clocks
    python tools/emu_run.py __graft_entry__.py smoke

Timings are meaningless (events carry host time stamps); the point is that the whole host flow and every kernel's
arithmetic run.  Never used by the product, the tests' GPU tier or the driver."""
import os
import runpy
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DRE_NO_PRIME", "1")

import dre_b200  # noqa: E402,F401
from dre_b200 import capi  # noqa: E402
from tests.simt import build_emu  # noqa: E402

capi.LIB_PATH, capi._lib = build_emu.build(), None
if len(sys.argv) < 2:
    sys.exit(__doc__)
script = sys.argv[1]
sys.argv = sys.argv[1:]
runpy.run_path(os.path.join(ROOT, script) if not os.path.isabs(script) else script, run_name="__main__")

"""Run any of the GPU test files against the emulated library (tests/emul/build/libmdgpu_emul.so: the product's sources compiled by g++, kernels
executed by host threads, fake CUDA runtime) instead of libmdgpu.so — on a machine without a GPU:

    python tests/emul/run_under_emulation.py tests/test_gpu_parity.py -m gpu -q          # ~25 min, the two full-size tests take 9 min each
    python tests/emul/run_under_emulation.py tests/test_zz_gpu_new_ops.py -m gpu -q       # ~15 s

Test infrastructure: it swaps the library path inside THIS process's viamd_b200.api before pytest starts; the product has no such switch."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests"), HERE):
    sys.path.insert(0, p)

import build_emul  # noqa: E402
import viamd_b200.api as api  # noqa: E402

api.LIB_PATH = build_emul.build_library(); api._lib = None

import pytest  # noqa: E402

sys.exit(pytest.main(sys.argv[1:] + ["-p", "no:cacheprovider"]))

"""Build tests/emul/build/libemul_sdf.so: product .cu sources compiled by g++ for CPU execution of thread-independent kernels (test infrastructure)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "build", "libemul_sdf.so")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")


def build() -> str:
    deps = [os.path.join(HERE, f) for f in ("emul_sdf.cpp", "cuda_emul.h")] + [os.path.join(HERE, "..", "..", "viamd_b200", "csrc", f) for f in ("sdf.cu", "kernels.h", "common.cuh")]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    # no -mfma / -march: a*b+c must stay two roundings, as under nvcc --fmad=false
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w", f"-I{CUDA_INC}", f"-I{HERE}",
                           os.path.join(HERE, "emul_sdf.cpp"), "-o", OUT])
    return OUT


if __name__ == "__main__":
    print(build())

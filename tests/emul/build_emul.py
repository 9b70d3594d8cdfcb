"""Build tests/emul/build/libemul_<name>.so: a product .cu file compiled by g++ (cuda_emul.h) so that its thread-independent kernels can be
executed on the CPU (test infrastructure). The kernel launch statements `k<<<grid, block, smem, stream>>>(args);` are the one construct
g++ cannot parse; they are blanked in a scratch copy of the source (build/<name>_nolaunch.cu) — kernels and device functions are
compiled exactly as written."""
import os
import re
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "..", "..", "viamd_b200", "csrc")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
CUDA_INC = os.path.join(CUDA_HOME, "include")
CUDA_LIB = os.path.join(CUDA_HOME, "lib64")
LAUNCH = re.compile(r"[A-Za-z_]\w*\s*(?:<[^<>;(){}]*>)?\s*<<<.*?>>>\s*\([^;]*?\)\s*;", re.S)


def build(name: str, sources=None) -> str:
    """name: wrapper emul_<name>.cpp -> libemul_<name>.so; sources: the product .cu files it includes (default [name])"""
    sources = sources or [name]
    out = os.path.join(HERE, "build", f"libemul_{name}.so"); wrap = os.path.join(HERE, f"emul_{name}.cpp")
    srcs = [os.path.join(CSRC, f"{s}.cu") for s in sources]
    deps = srcs + [wrap, os.path.join(HERE, "cuda_emul.h"), os.path.join(CSRC, "kernels.h"), os.path.join(CSRC, "common.cuh"), __file__]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    for sname, src in zip(sources, srcs):
        text = open(src).read()
        stripped, n = LAUNCH.subn("/* launch removed for host emulation */;", text)
        assert n > 0 and "<<<" not in stripped, f"{sname}.cu: {n} launches removed, some left"
        stripped = stripped.replace('#include "common.cuh"', f'#include "{os.path.join(CSRC, "common.cuh")}"').replace('#include "kernels.h"', f'#include "{os.path.join(CSRC, "kernels.h")}"')
        with open(os.path.join(HERE, "build", f"{sname}_nolaunch.cu"), "w") as f:
            f.write(stripped)
    # no -mfma / -march: a*b+c must stay two roundings, as under nvcc --fmad=false
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w", "-x", "c++", f"-I{CUDA_INC}", f"-I{HERE}",
                           f"-I{os.path.join(HERE, 'build')}", wrap, "-o", out,
                           f"-L{CUDA_LIB}", f"-Wl,-rpath,{CUDA_LIB}", "-lcudart", "-lpthread"])   # the (never called) launchers reference cudaMemsetAsync etc.
    return out


if __name__ == "__main__":
    print(build("sdf")); print(build("props")); print(build("within", ["cells", "within"])); print(build("sdfpipe", ["cells", "sdf"])); print(build("xtc"))

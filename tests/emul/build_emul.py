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


# rdf.cu reaches for inline PTX (packed f32x2 arithmetic, shared-window addressing, red.shared, the MUFU-based sqrt). For the host build
# those few helper definitions are swapped for plain C++ with the same semantics; everything else — the enumeration, the cull bound, the
# class logic, the queue discipline, the symmetric counting — is compiled as written. (pattern, replacement, expected number of matches)
RDF_PATCHES = [
    (r'^MDG_D u64 pk\(float a, float b\) \{.*$', 'MDG_D u64 pk(float a, float b) { uint32_t x, y; memcpy(&x, &a, 4); memcpy(&y, &b, 4); return (u64)x | ((u64)y << 32); }', 1),
    (r'^MDG_D u64 pkv\(float a, float b\) \{.*$', 'MDG_D u64 pkv(float a, float b) { return pk(a, b); }', 1),
    (r'^MDG_D void upk\(u64 v, float& a, float& b\) \{.*$', 'MDG_D void upk(u64 v, float& a, float& b) { const uint32_t x = (uint32_t)v, y = (uint32_t)(v >> 32); memcpy(&a, &x, 4); memcpy(&b, &y, 4); }', 1),
    (r'^MDG_D u64 sub2\(u64 a, u64 b\) \{.*$', 'MDG_D u64 sub2(u64 a, u64 b) { float a0, a1, b0, b1; upk(a, a0, a1); upk(b, b0, b1); return pk(a0 - b0, a1 - b1); }', 1),
    (r'^MDG_D u64 add2\(u64 a, u64 b\) \{.*$', 'MDG_D u64 add2(u64 a, u64 b) { float a0, a1, b0, b1; upk(a, a0, a1); upk(b, b0, b1); return pk(a0 + b0, a1 + b1); }', 1),
    (r'^MDG_D u64 mul2\(u64 a, u64 b\) \{.*$', 'MDG_D u64 mul2(u64 a, u64 b) { float a0, a1, b0, b1; upk(a, a0, a1); upk(b, b0, b1); return pk(a0 * b0, a1 * b1); }', 1),
    (r'^MDG_D u64 fma2\(u64 a, u64 b, u64 c\) \{.*$', 'MDG_D u64 fma2(u64 a, u64 b, u64 c) { float a0, a1, b0, b1, c0, c1; upk(a, a0, a1); upk(b, b0, b1); upk(c, c0, c1); return pk(fmaf(a0, b0, c0), fmaf(a1, b1, c1)); }', 1),
    (r'^MDG_D void q_push\(uint32_t& qaddr, float v\) \{.*$', 'MDG_D void q_push(uint32_t& qaddr, float v) { *(float*)(emul_dyn_smem + qaddr) = v; qaddr += 128u; }', 1),
    (r'^MDG_D float q_load\(uint32_t addr\) \{.*$', 'MDG_D float q_load(uint32_t addr) { return *(const float*)(emul_dyn_smem + addr); }', 1),
    (r'MDG_D float sqrt_rn_normal\(float x\) \{.*?\n\}', 'MDG_D float sqrt_rn_normal(float x) { return sqrtf(x); }   /* the device sequence is swept against IEEE sqrt on the GPU */', 1),
    (r'^    asm volatile\("red\.shared\.add\.u32 \[%0\], %1;".*$', '    __atomic_fetch_add((uint32_t*)(emul_dyn_smem + hist_saddr + 4u * (uint32_t)bin), w, __ATOMIC_RELAXED);', 1),
    (r'^    float4 rf; asm volatile\("ld\.shared\.v4\.f32.*$', '    return *(const float4*)(emul_dyn_smem + saddr + (uint32_t)OFF);', 1),
    (r'^    extern __shared__ __align__\(16\) unsigned char smem_raw\[\];$', '    unsigned char* smem_raw = emul_dyn_smem;', 1),
    (r'^    asm volatile\("mov\.u32 %0, %0;".*$', '', 3),
    (r'cudaFuncSetAttribute\(k_rdf_pairs_v2<[\w, ]+>, [^;]*;', ';', 6),                       # launcher-only runtime calls on kernel symbols
    (r'cudaOccupancyMaxActiveBlocksPerMultiprocessor\(&n, k_rdf_pairs_v2<[\w, ]+>, [^;]*;', 'n = 3;', 6),
    # TMA / mbarrier helpers of the VAR 2 kernel: the copy happens at once, the barrier is a phase counter
    (r'^MDG_D void mbar_init\(uint32_t mbar_saddr, uint32_t count\) \{.*$', 'MDG_D void mbar_init(uint32_t mbar_saddr, uint32_t count) { *(unsigned long long*)(emul_dyn_smem + mbar_saddr) = 0ull; }', 1),
    (r'^MDG_D void mbar_expect_tx\(uint32_t mbar_saddr, uint32_t bytes\) \{.*$', 'MDG_D void mbar_expect_tx(uint32_t mbar_saddr, uint32_t bytes) { }', 1),
    (r'^MDG_D bool mbar_try_wait\(uint32_t mbar_saddr, uint32_t parity\) \{.*$', 'MDG_D bool mbar_try_wait(uint32_t mbar_saddr, uint32_t parity) { const bool done = (__atomic_load_n((unsigned long long*)(emul_dyn_smem + mbar_saddr), __ATOMIC_ACQUIRE) & 1ull) != (unsigned long long)parity; if (!done && emul_block) emul_yield(); return done; }', 1),
    (r'^MDG_D void tma_load_1d\(uint32_t dst_saddr, const void\* src, uint32_t bytes, uint32_t mbar_saddr\) \{.*$', 'MDG_D void tma_load_1d(uint32_t dst_saddr, const void* src, uint32_t bytes, uint32_t mbar_saddr) { memcpy(emul_dyn_smem + dst_saddr, src, bytes); __atomic_fetch_add((unsigned long long*)(emul_dyn_smem + mbar_saddr), 1ull, __ATOMIC_RELEASE); }', 1),
    (r'^MDG_D void fence_mbar_init\(\) \{.*$', 'MDG_D void fence_mbar_init() { }', 1),
    (r'^MDG_D void fence_proxy_async\(\) \{.*$', 'MDG_D void fence_proxy_async() { }', 1),
]
PATCHES = {"rdf": RDF_PATCHES}


def build(name: str, sources=None) -> str:
    """name: wrapper emul_<name>.cpp -> libemul_<name>.so; sources: the product .cu files it includes (default [name])"""
    sources = sources or [name]
    out = os.path.join(HERE, "build", f"libemul_{name}.so"); wrap = os.path.join(HERE, f"emul_{name}.cpp")
    srcs = [os.path.join(CSRC, f"{s}.cu") for s in sources]
    deps = srcs + [wrap, os.path.join(HERE, "cuda_emul.h"), os.path.join(CSRC, "kernels.h"), os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "cellmath.cuh"), __file__]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    for sname, src in zip(sources, srcs):
        text = open(src).read()
        stripped, n = LAUNCH.subn("/* launch removed for host emulation */;", text)
        assert n > 0 and "<<<" not in stripped, f"{sname}.cu: {n} launches removed, some left"
        for pat, rep, want in PATCHES.get(sname, []):
            stripped, k = re.subn(pat, lambda m, rep=rep: rep, stripped, flags=re.M | re.S if "\\n" in pat else re.M)
            assert k == want, f"{sname}.cu: patch {pat!r} matched {k} times, expected {want}"
        assert "asm" not in re.sub(r"//.*", "", stripped), f"{sname}.cu: inline asm left after patching"
        stripped = stripped.replace('#include "common.cuh"', f'#include "{os.path.join(CSRC, "common.cuh")}"').replace('#include "kernels.h"', f'#include "{os.path.join(CSRC, "kernels.h")}"')
        with open(os.path.join(HERE, "build", f"{sname}_nolaunch.cu"), "w") as f:
            f.write(stripped)
    # no -mfma / -march: a*b+c must stay two roundings, as under nvcc --fmad=false
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared", "-w", "-x", "c++", f"-I{CUDA_INC}", f"-I{HERE}",
                           f"-I{os.path.join(HERE, 'build')}", f"-I{CSRC}", wrap, "-o", out,
                           f"-L{CUDA_LIB}", f"-Wl,-rpath,{CUDA_LIB}", "-lcudart", "-lpthread"])   # the (never called) launchers reference cudaMemsetAsync etc.
    return out


def _split_top(s: str):
    """split at top-level commas (parentheses / brackets / angle brackets of casts are balanced in the launch configurations used here)"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{": depth += 1
        elif ch in ")]}": depth -= 1
        if ch == "," and depth == 0: out.append(cur.strip()); cur = ""
        else: cur += ch
    out.append(cur.strip())
    return out


LAUNCH_FULL = re.compile(r"([A-Za-z_]\w*)\s*(<[^<>;(){}]*>)?\s*<<<(.*?)>>>\s*\(([^;]*?)\)\s*;", re.S)


def _launch_to_emul(m):
    name, tmpl, cfg, args = m.group(1), m.group(2) or "", _split_top(m.group(3)), m.group(4)
    return f"emul_launch(dim3({cfg[0]}), dim3({cfg[1]}), [&]() {{ {name}{tmpl}({args}); }});"


def build_library() -> str:
    """tests/emul/build/libmdgpu_emul.so: every source of libmdgpu.so compiled by g++ — each `k<<<grid, block, smem, stream>>>(args);` turned into
    `emul_launch(grid, block, [&]{ k(args); })`, rdf.cu's PTX helpers patched as above — and linked with fake_cudart.cpp instead of libcudart:
    the product's C ABI, host logic and kernels, executing on the CPU. Test infrastructure only."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("mdgpu_build", os.path.join(CSRC, "..", "build.py")); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
    asan = bool(os.environ.get("MDGPU_EMUL_ASAN"))   # AddressSanitizer build: out-of-bounds accesses of "device" buffers become hard errors
    tsan = bool(os.environ.get("MDGPU_EMUL_TSAN"))   # ThreadSanitizer build: accesses of the threads of a block that no barrier / atomic orders are reported
    tag = "asan" if asan else ("tsan" if tsan else "")
    out = os.path.join(HERE, "build", f"libmdgpu_emul_{tag}.so" if tag else "libmdgpu_emul.so"); bdir = os.path.join(HERE, "build", tag) if tag else os.path.join(HERE, "build")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in ("cuda_emul.h", "fake_cudart.cpp")] + [__file__, os.path.join(CSRC, "..", "..", "include", "mdgpu.h")]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    os.makedirs(bdir, exist_ok=True)
    flags = ["-std=c++20", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-w", f"-I{CUDA_INC}", f"-I{HERE}", f"-I{CSRC}", "-include", os.path.join(HERE, "cuda_emul.h")]
    if asan: flags += ["-fsanitize=address", "-fno-omit-frame-pointer", "-g"]
    if tsan: flags += ["-fsanitize=thread", "-fno-omit-frame-pointer", "-g"]
    objs, procs = [], []
    for src in b.SOURCES:
        sname = src[:-3]; text = open(os.path.join(CSRC, src)).read()
        text, n = LAUNCH_FULL.subn(_launch_to_emul, text)
        assert "<<<" not in text, f"{src}: launch statement not converted"
        for pat, rep, want in PATCHES.get(sname, []):
            text, k = re.subn(pat, lambda m, rep=rep: rep, text, flags=re.M | re.S if "\\n" in pat else re.M)
            assert k == want, f"{src}: patch {pat!r} matched {k} times, expected {want}"
        text = text.replace('#include "common.cuh"', f'#include "{os.path.join(CSRC, "common.cuh")}"').replace('#include "kernels.h"', f'#include "{os.path.join(CSRC, "kernels.h")}"').replace('#include "synth.h"', f'#include "{os.path.join(CSRC, "synth.h")}"')
        gen = os.path.join(bdir, f"{sname}_emullib.cpp"); open(gen, "w").write(text)
        obj = os.path.join(bdir, f"{sname}_emullib.o"); objs.append(obj)
        procs.append((src, subprocess.Popen(["g++", *flags, "-c", gen, "-o", obj])))
    fobj = os.path.join(bdir, "fake_cudart.o"); objs.append(fobj)
    procs.append(("fake_cudart.cpp", subprocess.Popen(["g++", *flags[:6], f"-I{CUDA_INC}", *(["-fsanitize=address", "-g"] if asan else []), *(["-fsanitize=thread", "-g"] if tsan else []), "-c", os.path.join(HERE, "fake_cudart.cpp"), "-o", fobj])))
    for src, p in procs:
        if p.wait() != 0: raise RuntimeError(f"g++ failed on {src}")
    # -Bsymbolic: the library's CUDA runtime calls must bind to fake_cudart.o inside it even when the process already holds the real
    # libcudart in its global scope (torch does that)
    subprocess.check_call(["g++", "-shared", "-Wl,-Bsymbolic", "-o", out, *objs, "-lpthread", "-lm", "-ldl"] + (["-fsanitize=address"] if asan else []) + (["-fsanitize=thread"] if tsan else []))
    return out


def build_fake_nccl() -> str:
    """tests/emul/build/libfakenccl.so: host-memory stand-in for the NCCL entry points the multi-device exchange binds (see fake_nccl.cpp)"""
    out = os.path.join(HERE, "build", "libfakenccl.so"); src = os.path.join(HERE, "fake_nccl.cpp")
    if not os.path.exists(out) or os.path.getmtime(src) > os.path.getmtime(out):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-o", out, src])
    return out


def build_loopback_nccl() -> str:
    """tests/emul/build/libloopbacknccl.so: fake_nccl.cpp on real device memory (cudaMemcpy staging) — lets a one-GPU box run the multi-device plan
    with both devices on the same GPU; needs the CUDA runtime, so it is built where it is used (the GPU tests)"""
    out = os.path.join(HERE, "build", "libloopbacknccl.so"); src = os.path.join(HERE, "fake_nccl.cpp")
    if not os.path.exists(out) or os.path.getmtime(src) > os.path.getmtime(out):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-DMDG_LOOPBACK_CUDA", "-I/usr/local/cuda/include", "-o", out, src,
                               "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64", "-lcudart"])
    return out


if __name__ == "__main__":
    print(build("sdf")); print(build("props")); print(build("within", ["cells", "within"])); print(build("sdfpipe", ["cells", "sdf"])); print(build("xtc")); print(build("rdfpipe", ["cells", "props", "rdf"])); print(build_library())

"""TEST INFRASTRUCTURE (CPU only): builds build/emu/libspral_ssids_b200_emu.so -- the engine's real sources
(subtree.cu, factor_kernels.cu, solve_kernels.cu, ssids_capi.cpp, analyse.cpp, scaling.cpp) compiled by g++ on top
of the SIMT emulator of tests/emu (cuda_emu.h, cuda_emu_rt.cpp), with update_emu.cpp standing in for gemm_dmma.cu
(inline PTX).  `kernel<<<grid, block[, smem[, stream]]>>>(args);` becomes emu::launch(grid, block, smem, [=]{ kernel(args); })
and `extern __shared__ T name[];` a pointer to the emulator's dynamic shared memory.  Loaded by tests only."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CSRC = os.path.join(ROOT, "spral_b200", "csrc")
EMU = os.path.join(ROOT, "tests", "emu")
OUT = os.path.join(ROOT, "build", "emu")
CUDA_INC = "/usr/local/cuda/include"
METIS = "/usr/local/cuda/targets/x86_64-linux/lib/libmetis_static.a"


def _balanced(s, i, open_c, close_c):
    """index just after the bracket that closes s[i] == open_c"""
    depth = 0
    while True:
        if s[i] == open_c:
            depth += 1
        elif s[i] == close_c:
            depth -= 1
            if depth == 0:
                return i + 1
        i += 1


def transform(src):
    src = re.sub(r"extern\s+__shared__\s+(?:__align__\(\d+\)\s+)?([A-Za-z_][\w ]*?)\s+(\w+)\[\];",
                 lambda m: f"{m.group(1)}* {m.group(2)} = ({m.group(1)}*)emu::dyn_smem();", src)
    out, pos = [], 0
    while True:
        i = src.find("<<<", pos)
        if i < 0:
            out.append(src[pos:])
            break
        # kernel expression: identifier, optionally followed by <template args>
        j = i
        while src[j - 1].isspace():
            j -= 1
        if src[j - 1] == ">":
            depth, k = 0, j - 1
            while True:
                if src[k] == ">":
                    depth += 1
                elif src[k] == "<":
                    depth -= 1
                    if depth == 0:
                        break
                k -= 1
            j = k
        k = j
        while src[k - 1].isalnum() or src[k - 1] in "_:":
            k -= 1
        kernel = src[k:i].strip()
        e = src.find(">>>", i)
        cfg = [c.strip() for c in src[i + 3:e].split(",")]
        a0 = src.index("(", e)
        a1 = _balanced(src, a0, "(", ")")
        args = src[a0 + 1:a1 - 1]
        smem = cfg[2] if len(cfg) > 2 else "0"
        out.append(src[pos:k])
        out.append(f"emu::launch((unsigned)({cfg[0]}), (unsigned)({cfg[1]}), (size_t)({smem}), [=]() {{ {kernel}({args}); }})")
        pos = a1
    return "".join(out)


def build(verbose=False, defines=(), tag=""):
    """defines: extra -D switches (e.g. SPRAL_B200_SPLIT); tag: suffix of the build directory / library name."""
    global OUT
    OUT = os.path.join(ROOT, "build", "emu" + tag)
    os.makedirs(OUT, exist_ok=True)
    flags = ["-O2", "-std=c++17", "-fPIC", "-fopenmp", "-w", "-I" + EMU, "-I" + CSRC, "-I" + os.path.join(ROOT, "include"),
             "-I" + CUDA_INC, "-U_FORTIFY_SOURCE"] + ["-D" + d for d in defines]
    objs = []
    for name in ("subtree.cu", "factor_kernels.cu", "solve_kernels.cu"):
        gen = os.path.join(OUT, name.replace(".cu", ".emu.cpp"))
        with open(os.path.join(CSRC, name)) as fh:
            text = transform(fh.read())
        with open(gen, "w") as fh:
            fh.write(f'#line 1 "{os.path.join(CSRC, name)}"\n' + text)
        obj = gen.replace(".cpp", ".o")
        subprocess.check_call(["g++"] + flags + ["-include", os.path.join(EMU, "cuda_emu.h"), "-c", gen, "-o", obj])
        objs.append(obj)
    for path in (os.path.join(EMU, "cuda_emu_rt.cpp"), os.path.join(EMU, "update_emu.cpp"),
                 os.path.join(CSRC, "ssids_capi.cpp"), os.path.join(CSRC, "analyse.cpp"), os.path.join(CSRC, "scaling.cpp")):
        obj = os.path.join(OUT, os.path.basename(path).replace(".cpp", ".o"))
        subprocess.check_call(["g++"] + flags + ["-c", path, "-o", obj])
        objs.append(obj)
    lib = os.path.join(OUT, f"libspral_ssids_b200_emu{tag}.so")
    # -Bsymbolic: the mock runtime of THIS library answers its CUDA calls even when torch has loaded the real libcudart
    subprocess.check_call(["g++", "-shared", "-fopenmp", "-Wl,-Bsymbolic", "-o", lib] + objs + [METIS, "-lm", "-ldl", "-lrt"])
    return lib


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "split":
        print(build(verbose=True, defines=("SPRAL_B200_SPLIT",), tag="_split"))
    else:
        print(build(verbose=True))

"""Build the native pieces in-tree (nvcc for sm_100a, g++ for the Python binding)."""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liboxli_b200.so")
EXT = os.path.join(HERE, "_oxli" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC",
]
NVCC_FLAGS += os.environ.get("OXLI_B200_NVCC_FLAGS", "").split()  # experiments: -DOXG_SCAT_THREADS=..., ...
OBJ = os.path.join(HERE, "_obj")


def _stale(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(exts: tuple[str, ...]) -> list[str]:
    out = [os.path.join(ROOT, "include", "oxli_b200.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith(exts):
            out.append(os.path.join(CSRC, f))
    return out


def specialised_ks() -> list[int]:
    """The k values of csrc/klist.h (one translation unit each)."""
    import re

    text = open(os.path.join(CSRC, "klist.h")).read()
    body = text[text.index("#define OXG_FOR_EACH_K"):]
    return [int(m) for m in re.findall(r"X\((\d+)\)", body)]


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    """capi.cu + one consume_inst.cu object per specialised k, compiled in parallel, linked into
    one shared library."""
    from concurrent.futures import ThreadPoolExecutor

    srcs = _sources((".cu", ".cuh", ".h", ".inc"))
    if not force and not _stale(LIB, srcs):
        return LIB  # up to date (the objects do not travel to the GPU box; the library does)
    os.makedirs(OBJ, exist_ok=True)
    jobs = [(os.path.join(OBJ, "capi.o"), os.path.join(CSRC, "capi.cu"), [])]
    for k in specialised_ks():
        jobs.append((os.path.join(OBJ, f"consume_k{k}.o"), os.path.join(CSRC, "consume_inst.cu"), [f"-DOXG_INST_K={k}"]))
    wanted = {j[0] for j in jobs}
    for f in os.listdir(OBJ):  # objects of k values that left the list
        if os.path.join(OBJ, f) not in wanted:
            os.remove(os.path.join(OBJ, f))

    def compile_one(job):
        obj, src, defs = job
        if force or _stale(obj, srcs):
            cmd = ["nvcc", *NVCC_FLAGS, *defs, "-c", "-o", obj, src]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
            subprocess.check_call(cmd)
            return True
        return False

    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as pool:
        rebuilt = list(pool.map(compile_one, jobs))
    if any(rebuilt) or not os.path.exists(LIB):
        subprocess.check_call(["nvcc", "-shared", "-gencode", "arch=compute_100a,code=sm_100a",
                               "-o", LIB, *[j[0] for j in jobs]])
    return LIB


def build_binding(force: bool = False) -> str:
    src = os.path.join(CSRC, "pyoxli.cpp")
    if not os.path.exists(src):
        return ""
    if force or _stale(EXT, [src, LIB, os.path.join(ROOT, "include", "oxli_b200.h")]):
        import pybind11

        cmd = [
            "g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
            "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
            "-I", os.path.join(ROOT, "include"), src, "-o", EXT,
            "-L", HERE, "-l:liboxli_b200.so", "-Wl,-rpath,$ORIGIN", "-lz",
        ]
        subprocess.check_call(cmd)
    return EXT


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_binding(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print(LIB)
    print(EXT)

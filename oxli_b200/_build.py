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
    "-Xcompiler", "-fPIC", "-shared",
]


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


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    srcs = _sources((".cu", ".cuh"))
    if force or _stale(LIB, srcs):
        cmd = ["nvcc", *NVCC_FLAGS, "-o", LIB, os.path.join(CSRC, "capi.cu")]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
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

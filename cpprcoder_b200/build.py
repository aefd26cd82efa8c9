"""In-tree build of the native pieces.

    libb2rc.so                 CUDA kernels + C ABI (cpprcoder_b200/csrc), sm_100a only
    oracle/liboracle.so        the C restatement used by tests (checker, not product)
    oracle/_ref/*.so           the unmodified reference behind extern "C" -- only where
                               /root/reference is mounted; elsewhere the prebuilt file is kept

nvcc cross-compiles without a GPU, so this runs on the CPU box as the "does it build" check.
"""
from __future__ import annotations

import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent


def _run(cmd, cwd):
    proc = subprocess.run(cmd, cwd=str(cwd), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f"build failed: {' '.join(cmd)}\n{proc.stdout}")
    return proc.stdout


def build_native(verbose: bool = False) -> Path:
    out = _run(["make", "-C", str(PKG / "csrc")], PKG)
    if verbose:
        print(out)
    so = PKG / "libb2rc.so"
    if not so.exists():
        raise RuntimeError("libb2rc.so was not produced")
    return so


def build_oracle(verbose: bool = False) -> None:
    out = _run(["make", "-C", str(ROOT / "oracle"), "all"], ROOT)
    if verbose:
        print(out)


def build_sim(verbose: bool = False) -> Path:
    sim = ROOT / "tests" / "sim"
    so = sim / "libsim_lane.so"
    srcs = [sim / "sim_lane.cpp", PKG / "csrc" / "rc_lane.cuh"]
    if not so.exists() or any(s.stat().st_mtime > so.stat().st_mtime for s in srcs):
        out = _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", "sim_lane.cpp", "-o", "libsim_lane.so"],
                   sim)
        if verbose:
            print(out)
    return so


def build_tools(verbose: bool = False) -> Path:
    """tools/e2e_cpp: the drop-in C++ classes timed end to end (bench.py's e2e_cpp leg runs it)."""
    out = ROOT / "tools" / "e2e_cpp"
    src = ROOT / "tools" / "e2e_cpp.cpp"
    deps = [src, PKG / "include" / "cpprcoder_b200.h", ROOT / "include" / "b2rc.h", PKG / "libb2rc.so"]
    if not out.exists() or any(d.stat().st_mtime > out.stat().st_mtime for d in deps):
        log = _run(["g++", "-O2", "-std=c++17", "-Wall", str(src), "-o", str(out), f"-L{PKG}", "-lb2rc",
                    "-Wl,-rpath,$ORIGIN/../cpprcoder_b200"], ROOT)
        if verbose:
            print(log)
    return out

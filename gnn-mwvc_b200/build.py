"""Build recipes (explicit nvcc / g++; no JIT cache, outputs stay in-tree).

  libgvc.so            csrc/gvc_api.cu for sm_100a              (always)
  host/_build/GNN_VC   the reference's own src/GNN_VC.cpp linked against the
                       drop-in host/gvc_gnn_inference.cpp + host/gvc_matrix.cpp + libgvc
                       (only where /root/reference exists; the binary travels)
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
REF = Path(os.environ.get("GVC_REFERENCE", "/root/reference"))

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        raise RuntimeError("nvcc not found")
    return nvcc


def _stale(out: Path, srcs) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(Path(s).stat().st_mtime > t for s in srcs)


def build_libgvc(force: bool = False, verbose: bool = False) -> Path:
    out = PKG / "libgvc.so"
    srcs = [PKG / "csrc" / "gvc_api.cu", PKG / "csrc" / "gvc_kernels.cuh", PKG / "csrc" / "gvc_px.cuh", PKG / "csrc" / "gvc_expf.h",
            ROOT / "include" / "gvc.h", PKG / "host" / "gvc_metis.cpp"]
    if force or _stale(out, srcs):
        extra = os.environ.get("GVC_NVCC_EXTRA", "").split()     # e.g. -DGVC_CTAS_PER_SM=2 for experiments
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-shared", "-o", str(out), str(srcs[0]), str(srcs[-1])]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.check_call(cmd)
    return out


def build_dropin(force: bool = False) -> Path | None:
    """GNN_VC built from the reference's UNMODIFIED driver + our replacement TUs."""
    if not (REF / "src" / "GNN_VC.cpp").exists():
        return None
    out = PKG / "host" / "_build" / "GNN_VC"
    srcs = [PKG / "host" / "gvc_gnn_inference.cpp", PKG / "host" / "gvc_matrix.cpp", ROOT / "include" / "gvc.h",
            PKG / "host" / "gvc_host_ctx.hpp"]
    if not all(s.exists() for s in srcs):
        return None
    if force or _stale(out, srcs + [PKG / "libgvc.so"]):
        out.parent.mkdir(parents=True, exist_ok=True)
        # flags of the reference Makefile:4 (x86-64-v3 instead of native: the binary travels)
        cmd = ["/usr/bin/g++", "-std=c++17", "-O3", "-march=x86-64-v3", "-DNDEBUG",
               "-I", str(REF / "include"), "-I", str(ROOT / "include"), "-I", str(PKG / "host"), "-o", str(out),
               str(REF / "src" / "GNN_VC.cpp"), str(srcs[0]), str(srcs[1]),
               "-L", str(PKG), "-lgvc", "-pthread", "-Wl,-rpath,$ORIGIN/../..", "-Wl,-rpath," + str(PKG)]
        subprocess.check_call(cmd)
    return out


def build_dropin_capi(force: bool = False) -> Path | None:
    """host/_build/libgvc_dropin.so: the drop-in host units behind a C binding of the reference's C++
    interface (host/gvc_dropin_capi.cpp), for Python callers (gnn-mwvc_b200/dropin.py)."""
    if not (REF / "include" / "gnn_inference.hpp").exists():
        out = PKG / "host" / "_build" / "libgvc_dropin.so"
        return out if out.exists() else None
    out = PKG / "host" / "_build" / "libgvc_dropin.so"
    srcs = [PKG / "host" / "gvc_dropin_capi.cpp", PKG / "host" / "gvc_gnn_inference.cpp", PKG / "host" / "gvc_matrix.cpp",
            ROOT / "include" / "gvc.h", PKG / "host" / "gvc_host_ctx.hpp"]
    if force or _stale(out, srcs + [PKG / "libgvc.so"]):
        out.parent.mkdir(parents=True, exist_ok=True)
        cmd = ["/usr/bin/g++", "-std=c++17", "-O3", "-march=x86-64-v3", "-DNDEBUG", "-fPIC", "-shared",
               "-I", str(REF / "include"), "-I", str(ROOT / "include"), "-I", str(PKG / "host"), "-o", str(out),
               str(srcs[0]), str(srcs[1]), str(srcs[2]),
               "-L", str(PKG), "-lgvc", "-pthread", "-Wl,--disable-new-dtags,-rpath,$ORIGIN/../.."]
        subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    print(build_libgvc(force=True, verbose=True))
    print(build_dropin(force=True))
    print(build_dropin_capi(force=True))

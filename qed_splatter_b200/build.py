"""In-tree nvcc build of the C-ABI library `libqedsplat.so` (sm_100a only).

`python -m qed_splatter_b200.build` or `__graft_entry__.build()`.  The .so is git-ignored but travels
to the GPU box with the gpurun snapshot.  There is no JIT and no fallback: if the library is missing,
`qed_splatter_b200._lib` raises.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
CSRC = PKG / "csrc"
INCLUDE = ROOT / "include"
BUILD = PKG / "_build"
LIB = PKG / "libqedsplat.so"

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v", f"-I{INCLUDE}", f"-I{CSRC}"]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*")) + list(INCLUDE.glob("*.h"))):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(ARCH_FLAGS + COMMON).encode())
    return h.hexdigest()


def _compile_one(src: Path) -> Path:
    obj = BUILD / (src.stem + ".o")
    cmd = [NVCC, *ARCH_FLAGS, *COMMON, "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    (BUILD / (src.stem + ".ptxas.log")).write_text(r.stderr)
    if r.returncode != 0:
        errs = "\n".join(ln for ln in (r.stdout + "\n" + r.stderr).splitlines() if "error" in ln.lower() or "fatal" in ln.lower())
        raise RuntimeError(f"nvcc failed for {src.name}:\n{errs or r.stderr[-4000:]}")
    return obj


def build(force: bool = False, verbose: bool = True) -> Path:
    BUILD.mkdir(exist_ok=True)
    stamp = BUILD / "digest.txt"
    digest = _digest()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text() == digest:
        if verbose:
            print(f"[qed build] up to date: {LIB}")
        return LIB
    srcs = _sources()
    if verbose:
        print(f"[qed build] nvcc {' '.join(ARCH_FLAGS)} : {[s.name for s in srcs]}")
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(_compile_one, srcs))
    cmd = [NVCC, *ARCH_FLAGS, "-shared", "-Xcompiler", "-fPIC", "-o", str(LIB), *map(str, objs), "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(digest)
    if verbose:
        print(f"[qed build] built {LIB} ({LIB.stat().st_size / 1e6:.1f} MB)")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)

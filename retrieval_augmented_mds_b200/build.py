"""In-tree build of libmips_b200.so (sm_100a only) with plain nvcc.

The shared library is the C-ABI boundary declared in include/mips_b200.h; it is built next to
the package so that it travels with the repo snapshot and is visible as an in-tree .so.

Staleness is decided by a hash of the sources and the flags (stored beside the library), not by
mtimes: a shipped .so whose stamp does not match the sources is rebuilt. Concurrent callers (one
process per GPU under torchrun) serialise on a file lock; the library is written to a temporary
file and renamed into place, so nobody can dlopen a half-written file.
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_DIR = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = PKG_DIR / "libmips_b200.so"
STAMP_PATH = PKG_DIR / "libmips_b200.so.stamp"
LOCK_PATH = PKG_DIR / ".build.lock"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    "-cudart", "static",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]
LINK_FLAGS = ["-ldl"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: libmips_b200.so cannot be built")


def sources() -> list[Path]:
    return [CSRC / "mips_api.cu"]


def _deps() -> list[Path]:
    return sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list((REPO_DIR / "include").glob("*.h")))


def source_hash() -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS + LINK_FLAGS).encode())
    for p in _deps():
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def needs_build() -> bool:
    if not LIB_PATH.exists() or not STAMP_PATH.exists():
        return True
    return STAMP_PATH.read_text().strip() != source_hash()


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB_PATH
    with open(LOCK_PATH, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():      # another rank built it while this one waited
                return LIB_PATH
            digest = source_hash()
            tmp = PKG_DIR / f".libmips_b200.{os.getpid()}.tmp.so"
            cmd = [_nvcc(), *NVCC_FLAGS, "-I", str(REPO_DIR / "include"), "-I", str(CSRC),
                   "-o", str(tmp), *map(str, sources()), *LINK_FLAGS]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                tmp.unlink(missing_ok=True)
                raise RuntimeError("nvcc failed building libmips_b200.so")
            os.replace(tmp, LIB_PATH)
            STAMP_PATH.write_text(digest + "\n")
            (PKG_DIR / "build_ptxas.log").write_text(res.stdout + res.stderr)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

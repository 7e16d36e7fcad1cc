"""In-tree build of libprobdiffeq_b200.so (hand-written CUDA for sm_100a) with plain nvcc.

The translation units under ``csrc/`` and ``csrc/inst/`` are compiled in parallel and linked into
``probdiffeq_b200/lib/libprobdiffeq_b200.so``.  The shared object is git-ignored but travels to the
GPU box with the repository snapshot.  No GPU is needed to build.
"""

from __future__ import annotations

import concurrent.futures
import hashlib
import os
import pathlib
import re
import shutil
import subprocess
import sys

ROOT = pathlib.Path(__file__).resolve().parent
CSRC = ROOT / "csrc"
OBJ = ROOT / "build"
LIB = ROOT / "lib" / "libprobdiffeq_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]  # fmt: skip


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; cannot build libprobdiffeq_b200.so")
    return exe


def sources() -> list[pathlib.Path]:
    return sorted(CSRC.glob("*.cu")) + sorted((CSRC / "inst").glob("*.cu"))


_INCLUDE = re.compile(r'^\s*#include\s+"([^"]+)"', re.M)


def _includes(src: pathlib.Path, seen: dict[pathlib.Path, None] | None = None) -> list[pathlib.Path]:
    """The project headers `src` includes, transitively (quoted includes only), in a stable order."""
    seen = {} if seen is None else seen
    for name in _INCLUDE.findall(src.read_text()):
        hdr = (src.parent / name).resolve()
        if hdr.exists() and hdr not in seen:
            seen[hdr] = None
            _includes(hdr, seen)
    return sorted(seen)


def _compile(src: pathlib.Path, force: bool, verbose: bool) -> pathlib.Path:
    """An object is reused only if the CONTENT of its source, of every header it includes (transitively) and the flags
    are what it was built from (a stamp beside the object records their hash) -- not because its mtime is newer. A
    change to one kernel family's header therefore rebuilds that family's translation units only."""
    obj = OBJ / (src.stem + ".o")
    stamp = OBJ / (src.stem + ".sha256")
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for f in [src, *_includes(src)]:
        h.update(f.name.encode() + f.read_bytes())
    want = h.hexdigest()
    if not force and obj.exists() and stamp.exists() and stamp.read_text() == want:
        return obj
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
    if verbose:
        sys.stderr.write(res.stderr)
    stamp.write_text(want)
    return obj


def build(force: bool = False, verbose: bool = False, jobs: int | None = None) -> pathlib.Path:
    OBJ.mkdir(exist_ok=True)
    LIB.parent.mkdir(exist_ok=True)
    srcs = sources()
    jobs = jobs or min(len(srcs), os.cpu_count() or 4)
    with concurrent.futures.ThreadPoolExecutor(max_workers=jobs) as pool:
        objs = list(pool.map(lambda s: _compile(s, force, verbose), srcs))
    link_stamp = OBJ / "link.sha256"
    link_want = hashlib.sha256("".join((OBJ / (s.stem + ".sha256")).read_text() for s in srcs).encode()).hexdigest()
    if force or not LIB.exists() or not link_stamp.exists() or link_stamp.read_text() != link_want:
        cmd = [_nvcc(), "-shared", "-o", str(LIB), *map(str, objs), "-ldl"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
        link_stamp.write_text(link_want)
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)

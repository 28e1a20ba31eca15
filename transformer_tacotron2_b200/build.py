"""Builds libtts_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m transformer_tacotron2_b200.build [--force] [--verbose]

Staleness is decided by a content hash of the sources (recorded next to the library at build time), not by mtimes: a
snapshot copied to a GPU box keeps its prebuilt library whatever order the files were written in.  Concurrent callers
(N ranks of one torchrun job) serialise on a file lock; the library is written to a temporary file and renamed into place,
so nobody ever dlopens a half-written file.
"""
from __future__ import annotations

import fcntl
import hashlib
import os
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libtts_b200.so")
HASH_PATH = LIB_PATH + ".srchash"
LOCK_PATH = LIB_PATH + ".lock"
INCLUDE_DIR = os.path.join(os.path.dirname(PKG_DIR), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17", "--shared", "-Xcompiler", "-fPIC",
]


def _sources():
    out = [os.path.join(INCLUDE_DIR, "tts_b200.h")]
    for f in sorted(os.listdir(CSRC)):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(CSRC, f))
    return out


def source_hash() -> str:
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for s in _sources():
        h.update(os.path.basename(s).encode())
        with open(s, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def needs_build() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    try:
        with open(HASH_PATH) as f:
            return f.read().strip() != source_hash()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB_PATH
    with open(LOCK_PATH, "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not force and not needs_build():          # another process built it while we waited for the lock
                return LIB_PATH
            digest = source_hash()
            nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
            tmp = f"{LIB_PATH}.tmp.{os.getpid()}"
            cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + [
                "-I", INCLUDE_DIR, "-o", tmp, os.path.join(CSRC, "tts_b200.cu")]
            res = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
            if res.returncode != 0:
                if os.path.exists(tmp):
                    os.unlink(tmp)
                raise RuntimeError("nvcc failed building libtts_b200.so (see stderr)")
            os.replace(tmp, LIB_PATH)                     # atomic: a concurrent dlopen sees the old or the new file, never a torn one
            with open(HASH_PATH + ".tmp", "w") as f:
                f.write(digest + "\n")
            os.replace(HASH_PATH + ".tmp", HASH_PATH)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

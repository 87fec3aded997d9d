"""Compile libnsvd.so (the C-ABI library of include/nsvd.h) for sm_100a with nvcc, in-tree."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libnsvd.so")
SOURCES = ["nsvd_api.cu", "nsvd_simt.cu", "nsvd_tc.cu"]
NVCC_FLAGS = ["-shared", "-Xcompiler", "-fPIC", "-gencode", "arch=compute_100a,code=sm_100a",
              "-lineinfo", "-O3", "-std=c++17", "-diag-suppress", "550"]


HASHFILE = LIB + ".srchash"


def _source_hash() -> str:
    """sha256 over the sources, the header and the flags (content-based: file times do not survive copies)."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC)) + [os.path.join(HERE, "..", "include", "nsvd.h")]
    for d in deps:
        if os.path.isfile(d):
            h.update(os.path.basename(d).encode())
            h.update(open(d, "rb").read())
    return h.hexdigest()


def _stale() -> bool:
    if not os.path.isfile(LIB) or not os.path.isfile(HASHFILE):
        return True
    return open(HASHFILE).read().strip() != _source_hash()


def embedded_hash(path: str = LIB):
    """source hash compiled into a library (nsvd_build_hash), or None if it cannot be read"""
    import ctypes
    try:
        lib = ctypes.CDLL(path)
        lib.nsvd_build_hash.restype = ctypes.c_char_p
        return lib.nsvd_build_hash().decode()
    except (OSError, AttributeError):
        return None


def build(force: bool = False, verbose: bool = False) -> str:
    """Build the library if it is missing or older than its sources. Returns the .so path."""
    if not force and not _stale():
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.isfile(nvcc):
        if os.path.isfile(LIB):
            # prebuilt library travelled with the tree (GPU box without nvcc): it must be the build of THESE sources -
            # the hash file does not travel through git, so the hash compiled into the library is what is compared
            have, want = embedded_hash(), _source_hash()
            if have != want:
                msg = (f"libnsvd.so was built from other sources (embedded hash {have}, tree {want}) and nvcc is not "
                       "available to rebuild it")
                if os.environ.get("NSVD_ALLOW_STALE") != "1":
                    raise RuntimeError(msg + "; set NSVD_ALLOW_STALE=1 to load it anyway")
                import warnings
                warnings.warn(msg)
            return LIB
        raise RuntimeError("nvcc not found and no prebuilt libnsvd.so in the tree")
    import fcntl
    with open(LIB + ".lock", "w") as lock:       # one builder at a time (torchrun starts N ranks at once)
        fcntl.flock(lock, fcntl.LOCK_EX)
        if not force and not _stale():
            return LIB
        tmp = f"{LIB}.tmp.{os.getpid()}"
        cmd = ([nvcc] + NVCC_FLAGS + [f'-DNSVD_SRC_HASH="{_source_hash()}"'] + (["-Xptxas", "-v"] if verbose else [])
               + ["-o", tmp] + SOURCES)
        r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
        if verbose:
            sys.stderr.write(r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        os.replace(tmp, LIB)
        with open(HASHFILE, "w") as f:
            f.write(_source_hash())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

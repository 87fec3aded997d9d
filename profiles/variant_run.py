"""Run a command against a side build of the library compiled with extra nvcc defines (development probes, gpurun).
usage: python profiles/variant_run.py "-DFOO=1 -DBAR" -- python bench.py ..."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    sep = sys.argv.index("--")
    defines = " ".join(sys.argv[1:sep]).split()
    cmd = sys.argv[sep + 1:]
    from neural_svd_b200 import build
    csrc = os.path.join(ROOT, "neural_svd_b200", "csrc")
    side = os.path.join(ROOT, "gpurun_out", "libnsvd_variant.so")
    os.makedirs(os.path.dirname(side), exist_ok=True)
    flags = build.NVCC_FLAGS + [f'-DNSVD_SRC_HASH="{build._source_hash()}"'] + defines
    subprocess.run(["nvcc"] + flags + ["-o", side] + build.SOURCES, cwd=csrc, check=True)
    main_lib = build.LIB
    shutil.copy(main_lib, main_lib + ".bak")
    shutil.copy(side, main_lib)
    try:
        return subprocess.run(cmd, cwd=ROOT).returncode
    finally:
        shutil.move(main_lib + ".bak", main_lib)


if __name__ == "__main__":
    sys.exit(main())

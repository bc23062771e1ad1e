"""Build a variant of libfdn_b200.so for A/B timing (development only).

    python tools/build_variant.py OUT.so [--src DIR] [-DFLAG ...]

Compiles the library's sources (or the copies in --src, e.g. an older revision extracted with `git show`) with extra
nvcc flags into OUT.so using a private object directory. tools/flow_iter_lab.py and bench.py load it when
FDN_LIB_PATH points to it. The product build is flowdenoising_b200/_build.py.
"""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from flowdenoising_b200 import _build  # noqa: E402


def main():
    out = os.path.abspath(sys.argv[1])
    args = sys.argv[2:]
    src = _build.CSRC
    inc = os.path.join(ROOT, "include")
    if "--src" in args:
        i = args.index("--src")
        src = os.path.abspath(args[i + 1])
        if os.path.exists(os.path.join(src, "fdn_b200.h")):
            inc = src
        del args[i:i + 2]
    flags = [f for f in _build.NVCC_FLAGS]
    flags[flags.index("-I") + 1] = inc
    nvcc = _build.nvcc_path()
    with tempfile.TemporaryDirectory() as tmp:
        procs, objs = [], []
        for s in _build.SOURCES:
            o = os.path.join(tmp, s.replace(".cu", ".o"))
            procs.append(subprocess.Popen([nvcc, *flags, "-I", src, *args, "-c", os.path.join(src, s), "-o", o]))
            objs.append(o)
        for p in procs:
            if p.wait() != 0:
                raise SystemExit("nvcc failed")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call([nvcc, "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"])
    print(out)


if __name__ == "__main__":
    main()

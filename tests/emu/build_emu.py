"""TEST-ONLY: builds tests/emu/_build/libsfb_emu.so = the product sources (C ABI, host glue,
kernels) compiled with g++ against the fiber emulator in cuda_emu.h.  The result is never
loaded by the package; tests point SFB_LIB at it in a subprocess (tests/test_emu_parity.py)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build", "libsfb_emu.so")
DEPS = [os.path.join(HERE, "cuda_emu.h"), os.path.join(HERE, "emu_lib.cpp")] + [
    os.path.join(ROOT, "simfire_b200", "csrc", f) for f in sorted(os.listdir(os.path.join(ROOT, "simfire_b200", "csrc")))
] + [os.path.join(ROOT, "include", "simfire_b200.h")]


def build(force: bool = False, extra=(), out=OUT) -> str:
    if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in DEPS):
        return out
    os.makedirs(os.path.dirname(out), exist_ok=True)
    # -ffp-contract=off mirrors nvcc --fmad=false (the Rothermel arithmetic rounds after every operation)
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I", HERE, *extra, "-o", out,
           os.path.join(HERE, "emu_lib.cpp"), "-lpthread"]  # fmt: skip
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + " ".join(cmd) + "\n" + res.stderr)
    return out


def build_variant(macro: str) -> str:
    """The same library with -D<macro> (experimental kernel variants, e.g. SFB_ROWS_V2)."""
    return build(extra=("-D" + macro,), out=os.path.join(HERE, "_build", f"libsfb_emu_{macro.lower()}.so"))


def build_asan() -> str:
    """AddressSanitizer build: device allocations get their exact size, so a kernel (or host) access
    one byte past a plane, list or flag array aborts the test.  Load with LD_PRELOAD=libasan."""
    return build(extra=("-fsanitize=address", "-fno-omit-frame-pointer"), out=os.path.join(HERE, "_build", "libsfb_emu_asan.so"))


def asan_runtime() -> str:
    return subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True, check=True).stdout.strip()


if __name__ == "__main__":
    print(build(force=True))

"""TEST-ONLY microbenchmark of the host side of sfb_sync_fire_maps (apply_log in sfb.cu) on a
synthetic change log shaped like the bench's: E envs of H x W, ring-shaped fire fronts, entries in
(env, row) order as k_eval appends them.  Usage: python tests/emu/bench_patch.py [threads]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from build_emu import build  # noqa: E402


def synthetic_log(H, W, E, per_env, seed=0):
    rng = np.random.default_rng(seed)
    out = []
    for e in range(E):
        cy, cx, r = rng.integers(300, H - 300), rng.integers(300, W - 300), rng.integers(60, 250)
        th = np.sort(rng.random(per_env)) * 2 * np.pi
        y = np.clip((cy + r * np.sin(th)).astype(np.int64), 0, H - 1)
        x = np.clip((cx + r * np.cos(th)).astype(np.int64), 0, W - 1)
        order = np.argsort(y, kind="stable")
        idx = (e * H + y[order]) * W + x[order]
        out.append(idx.astype(np.uint64) | (np.uint64(1) << np.uint64(48)))
    return np.concatenate(out)


def main():
    threads = int(sys.argv[1]) if len(sys.argv) > 1 else os.cpu_count()
    lib = C.CDLL(build())
    lib.sfb_emu_bench_apply_log.restype = C.c_double
    H = W = 2048
    E = 256  # one env group of the target workload
    log = synthetic_log(H, W, E, 540)  # ~138 k entries per group and step
    mirror = np.zeros((E, H, W), np.int8)
    for single in (0, 1):
        mirror[...] = 0
        ms = lib.sfb_emu_bench_apply_log(H, W, E, C.c_void_p(log.ctypes.data), C.c_longlong(len(log)),
                                         C.c_void_p(mirror.ctypes.data), threads, 20, single)
        print(f"{len(log)} entries, {threads} threads, {'single-step (one pass)' if single else 'ordered (two passes)'}: "
              f"{ms:.3f} ms per apply_log ({ms * 1e6 / len(log):.1f} ns/entry)")
        assert int((mirror == 1).sum()) == len(np.unique(log))


if __name__ == "__main__":
    main()

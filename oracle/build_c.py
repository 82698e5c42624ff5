"""Builds oracle/_build/liboracle_lookup.so from oracle/lookup_taps.c with gcc (test infrastructure)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, '_build')
LIB = os.path.join(OUT_DIR, 'liboracle_lookup.so')


def build(force: bool = False) -> str:
    src = os.path.join(HERE, 'lookup_taps.c')
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(src):
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-fno-fast-math', '-shared', '-fPIC', src, '-o', LIB, '-lm'])
    return LIB


def load():
    import ctypes as C
    lib = C.CDLL(build())
    lib.oracle_lookup_taps.restype = None
    lib.oracle_lookup_taps.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 6
    return lib


def lookup_taps(flow8_nhwc, level: int, radius: int):
    """flow8_nhwc: contiguous float32 numpy [B,H,W,2]. Returns dict of int32 / uint8 arrays [B,H,W,2r+1]."""
    import numpy as np
    b, h, w, _ = flow8_nhwc.shape
    k = 2 * radius + 1
    flow8_nhwc = np.ascontiguousarray(flow8_nhwc, dtype=np.float32)
    x0 = np.empty((b, h, w, k), np.int32)
    y0 = np.empty_like(x0)
    masks = [np.empty((b, h, w, k), np.uint8) for _ in range(4)]
    load().oracle_lookup_taps(flow8_nhwc.ctypes.data, b, h, w, level, radius, x0.ctypes.data, y0.ctypes.data,
                              *[m.ctypes.data for m in masks])
    return dict(x0=x0, y0=y0, mx0=masks[0], mx1=masks[1], my0=masks[2], my1=masks[3])


if __name__ == '__main__':
    print(build(force=True))

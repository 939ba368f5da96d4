"""The pipelines of the reference's integration tests (lib/zosimos/tests/blend.rs, knobs.rs),
written once against an abstract `ops` namespace so the same recipe runs on the CPU oracle and on
the CUDA backend.  Each returns the output image as an (h, w, 4) uint8 RGBA array."""
import math

import numpy as np

KNOBS = [  # tests/knobs.rs:48-88
    ([0.0, 0.0, 0.5, 0.5], [0.5, 0.5, 0.5, 0.5], [0.0, 0.0, 0.5, 0.5], [0.5, 0.5, 0.5, 0.5]),
    ([0.2, 0.0, 0.5, 0.5], [0.5, 0.5, 0.5, 0.5], [0.2, 0.0, 0.5, 0.5], [0.5, 0.5, 0.5, 0.5]),
    ([0.4, 0.0, 0.5, 0.5], [0.5, 0.5, 0.5, 0.5], [0.4, 0.0, 0.5, 0.5], [0.5, 0.5, 0.5, 0.5]),
    ([0.5, 0.0, 0.4, 0.5], [0.5, 0.5, 0.5, 0.5], [0.5, 0.0, 0.4, 0.5], [0.5, 0.5, 0.5, 0.5]),
    ([0.5, 0.0, 0.2, 0.5], [0.5, 0.5, 0.5, 0.5], [0.5, 0.0, 0.2, 0.5], [0.5, 0.5, 0.5, 0.5]),
]
DERIVATIVES = {  # command.rs:3343-3409
    "Sobel": [1 / 4, 1 / 2, 1 / 4],
    "Prewitt": [1 / 3, 1 / 3, 1 / 3],
    "Scharr3": [46.84 / 256, 162.32 / 256, 46.84 / 256],
    "Scharr3To4Bit": [3 / 16, 10 / 16, 3 / 16],
    "Scharr3To8Bit": [47 / 256, 162 / 256, 47 / 256],
}
LCH_GRID = ([0.4, 0.0, 0.0, 1.0], [0.4, 0.0, 1.0, 1.0], [0.4, 0.0, 0.0, 1.0], [0.4, 1.0, 0.0, 1.0],
            [0.0, 0.0, 0.0, 0.0], [0.0, 0.0, 0.0, 0.0])  # tests/blend.rs:474-487 (u_min,u_max,v_min,v_max,uv_min,uv_max)


def affine_matrix_blend_rs(fw, fh, W, H):
    """tests/blend.rs:128-141: shift(-(fw/2),-(fh/2)) . rotate(pi/4) . shift(W/2,H/2); each step is a
    LEFT multiplication carried out in f32 (command.rs:3437-3485)."""
    def f32m(m):
        return np.asarray(m, dtype=np.float32)

    def mul(a, b):  # RowMatrix::multiply_right: dot products in f32, left to right
        a = f32m(a); b = f32m(b)
        o = np.zeros((3, 3), np.float32)
        for r in range(3):
            for c in range(3):
                o[r, c] = np.float32(np.float32(np.float32(a[r, 0] * b[0, c]) + np.float32(a[r, 1] * b[1, c])) + np.float32(a[r, 2] * b[2, c]))
        return o
    rad = np.float32(math.pi) / np.float32(4.0)
    c, s = np.float32(np.cos(rad)), np.float32(np.sin(rad))
    m = f32m(np.eye(3))
    m = mul([[1, 0, -float(fw // 2)], [0, 1, -float(fh // 2)], [0, 0, 1]], m)
    m = mul([[c, s, 0], [-s, c, 0], [0, 0, 1]], m)
    m = mul([[1, 0, float(W // 2)], [0, 1, float(H // 2)], [0, 0, 1]], m)
    return m

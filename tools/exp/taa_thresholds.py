"""The float32 thresholds vxl_post.cu compares radicands with instead of taking square roots (LightTAA.frag:72, :76, :112).

sqrtf is correctly rounded and therefore monotone, and so are the float operations the shader applies to the root, so each of the
shader's comparisons flips at exactly one float; bisection over the bit patterns finds it and the neighbours are printed as a check."""
import numpy as np

f32 = np.float32


def fromb(b):
    return np.array([b], dtype=np.uint32).view(f32)[0]


def bits(x):
    return int(np.array([x], dtype=f32).view(np.uint32)[0])


def flip(pred, lo, hi):
    lo_b, hi_b = bits(lo), bits(hi)
    assert not pred(fromb(lo_b)) and pred(fromb(hi_b))
    while hi_b - lo_b > 1:
        m = (lo_b + hi_b) // 2
        if pred(fromb(m)):
            hi_b = m
        else:
            lo_b = m
    assert [bool(pred(fromb(lo_b + d))) for d in range(-3, 5)] == [False] * 4 + [True] * 4
    return lo_b, hi_b


root = lambda s: np.sqrt(f32(s), dtype=f32)
lo, _ = flip(lambda s: root(s) > f32(0.1), 0.009, 0.011)
print("T_MOTION   (largest s with sqrt(s) <= 0.1f)            ", float(fromb(lo)).hex())
lo, _ = flip(lambda s: (f32(1.0) - root(s)) < f32(0.8), 0.03, 0.05)
print("T_MATERIAL (largest s with 1 - sqrt(s) >= 0.8f)         ", float(fromb(lo)).hex())
_, hi = flip(lambda s: root(s) * f32(10000.0) >= f32(1.0), 0.9e-8, 1.1e-8)
print("T_COLOR    (smallest s with sqrt(s) * 10000f >= 1)      ", float(fromb(hi)).hex())

"""LightTAA compares three square roots with constants (LightTAA.frag:72, :76, :112); vxl_post.cu compares the radicands with ONE float
each instead.  sqrtf is correctly rounded, hence monotone, so each comparison flips at exactly one float: this test finds that float by
bisection over the bit patterns, checks it against the constant compiled into the kernel, and checks the equivalence itself on the
floats around the flip and on random radicands."""
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32


def _fromb(b):
    return np.array([b], dtype=np.uint32).view(f32)[0]


def _bits(x):
    return int(np.array([x], dtype=f32).view(np.uint32)[0])


def _flip(pred, lo, hi):
    """largest bit pattern with pred false, smallest with pred true (pred monotone on [lo, hi])"""
    lo_b, hi_b = _bits(lo), _bits(hi)
    assert not pred(_fromb(lo_b)) and pred(_fromb(hi_b))
    while hi_b - lo_b > 1:
        m = (lo_b + hi_b) // 2
        if pred(_fromb(m)):
            hi_b = m
        else:
            lo_b = m
    return lo_b, hi_b


def _root(s):
    return np.sqrt(f32(s), dtype=f32)


SHADER = {   # name -> (the shader's comparison on the radicand, bracket, which side of the flip the kernel's constant is, kernel's comparison)
    "T_MOTION": (lambda s: _root(s) > f32(0.1), (0.009, 0.011), 0, lambda s, t: f32(s) > t),
    "T_MATERIAL": (lambda s: (f32(1.0) - _root(s)) < f32(0.8), (0.03, 0.05), 0, lambda s, t: f32(s) > t),
    "T_COLOR": (lambda s: _root(s) * f32(10000.0) >= f32(1.0), (0.9e-8, 1.1e-8), 1, lambda s, t: f32(s) >= t),
}


def _kernel_constants():
    src = open(os.path.join(ROOT, "voxelengine_b200", "csrc", "vxl_post.cu")).read()
    out = {}
    for name in SHADER:
        m = re.search(name + r"\s*=\s*(0x[0-9a-fA-F.]+p[-+]?\d+)f", src)
        assert m, f"{name} not found in vxl_post.cu"
        out[name] = f32(float.fromhex(m.group(1)))
    return out


def test_thresholds_are_the_flip_points_of_the_shader_comparisons():
    consts = _kernel_constants()
    for name, (pred, (lo, hi), side, _) in SHADER.items():
        flip = _flip(pred, lo, hi)
        assert _bits(consts[name]) == flip[side], f"{name}: kernel constant {float(consts[name]).hex()} vs flip {float(_fromb(flip[side])).hex()}"
        # the comparison is monotone around the flip: 64 floats either side
        for d in range(-64, 65):
            assert bool(pred(_fromb(flip[1] + d))) == (d >= 0)


def test_radicand_comparison_equals_the_shader_comparison():
    consts = _kernel_constants()
    rng = np.random.default_rng(7)
    for name, (pred, (lo, hi), _, kern) in SHADER.items():
        t = consts[name]
        # radicands over twelve decades around the threshold, plus the floats next to it, zero and infinity
        s = np.concatenate([(t * f32(10.0) ** rng.uniform(-6, 6, 20000)).astype(f32),
                            np.array([_fromb(_bits(t) + d) for d in range(-200, 201)], dtype=f32), np.array([0.0, np.inf], dtype=f32)])
        want = np.array([bool(pred(x)) for x in s])
        got = np.array([bool(kern(x, t)) for x in s])
        assert np.array_equal(want, got), name
    # a NaN radicand fails both forms
    nan = f32(np.nan)
    for name, (pred, _, _, kern) in SHADER.items():
        with np.errstate(invalid="ignore"):
            assert not bool(pred(nan)) and not bool(kern(nan, consts[name]))

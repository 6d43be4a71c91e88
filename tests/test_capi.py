"""CPU: the C-ABI library loads and exports every symbol include/vxl.h declares; struct layouts match
the reference's (View.h, Light*Pipeline.h); without a GPU the product fails loudly (no fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from voxelengine_b200.build import build
    build()
    from voxelengine_b200 import capi
    return capi.load()


def test_every_declared_symbol_is_exported(lib):
    from voxelengine_b200 import capi
    hdr = open(os.path.join(ROOT, "include", "vxl.h")).read()
    declared = sorted(set(re.findall(r"^\s*(?:int|const char\*)\s+(vxl_[a-z0-9_]+)\s*\(", hdr, re.M)))
    assert declared == sorted(capi.SYMBOLS), set(declared) ^ set(capi.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.vxl_abi_version() == 1


def test_struct_layouts():
    from voxelengine_b200 import capi
    from voxelengine_b200 import scenes as S
    assert S.VIEW_DTYPE.itemsize == 380                      # sizeof(ViewData), View.h:16-30
    assert S.VIEW_DTYPE.fields["Res"][1] == 320 and S.VIEW_DTYPE.fields["Frame"][1] == 360
    assert S.POINT_LIGHT_DTYPE.itemsize == 32 and S.SPOT_LIGHT_DTYPE.itemsize == 64
    assert S.SPOT_LIGHT_DTYPE.fields["Direction"][1] == 32 and S.SPOT_LIGHT_DTYPE.fields["AngleAttenuation"][1] == 48
    assert S.ENTITY_DTYPE.itemsize == 152 and S.REGION_DTYPE.itemsize == 28
    assert S.HIT_DTYPE.itemsize == 48 and S.RAY_DTYPE.itemsize == 32
    assert C.sizeof(capi.Frame) == 64 and C.sizeof(capi.Stats) == 24
    assert C.sizeof(capi.LightingHostArgs) == 64 + 8 + 4 * 3 + 4 + 8 * 7


def test_no_gpu_fails_loudly(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = lib.vxl_ctx_create(0, C.byref(h))
    assert rc == -2 and not h.value                          # VXL_ERR_CUDA
    assert b"no CPU fallback" in lib.vxl_last_error_string()
    from voxelengine_b200 import capi, engine
    with pytest.raises(capi.VxlError):
        engine.Context(0)


def test_bad_arguments_return_error_codes(lib):
    assert lib.vxl_ctx_create(0, None) == -1
    assert lib.vxl_sync(None) == -1
    assert lib.vxl_volume_create(None, 4, 4, 4, None) == -1
    assert lib.vxl_stats_read(None, None) == -1
    assert b"vxl_stats_read" in lib.vxl_last_error_string()


def test_product_does_not_reference_the_oracle():
    """The product path must never import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "voxelengine_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in src, os.path.join(dp, f)
                assert "vxo_" not in src, os.path.join(dp, f)


def test_scenes_inputs_are_deterministic():
    from voxelengine_b200 import scenes as S
    assert np.array_equal(S.blue_noise(4), S.blue_noise(4)) and S.blue_noise(4).dtype == np.uint32
    assert np.array_equal(S.house_model(40, 1), S.house_model(40, 1))
    m = S.house_model(40, 1)
    assert (m >= 16).sum() > 5000 and ((m > 0) & (m < 16)).sum() > 0      # solid + glass present
    v = S.make_view((1, 2, 3), 0.81, -0.43, 320, 200, 7)
    assert v["Frame"] == 7 and v["Res"].tolist() == [320.0, 200.0]
    iv = v["InverseViewMatrix"].reshape(4, 4).T
    assert np.allclose(iv[:3, 3], (1, 2, 3))
    assert np.allclose(v["ViewMatrix"].reshape(4, 4).T @ iv, np.eye(4), atol=1e-5)

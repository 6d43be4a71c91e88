"""The C++ host side above the C ABI (include/vxl_pipelines.hpp: the reference's pass objects -- ShadowVoxSystem, Light*Pipeline::Get().Use,
DrawLight -- as a header-only wrapper) driven by a C++ program, tests/cpp/host_pipelines.cpp, the way WorldRenderer::DrawWorld drives
the reference's.  CPU: the program compiles with plain g++ (no CUDA headers), links against libvxl.so, and without a device fails
loudly like the reference's CHECK (throw, no fallback).  GPU: its output planes equal the oracle's bit for bit."""
import os
import subprocess

import numpy as np
import pytest

import scene_util as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_pipelines.cpp")


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    from voxelengine_b200.build import build
    build()
    out = str(tmp_path_factory.mktemp("cpp") / "host_pipelines")
    libdir = os.path.join(ROOT, "voxelengine_b200")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC, "-o", out,
                        "-L" + libdir, "-lvxl", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def _write_scene(path, sc, n_ao, point, spot):
    from voxelengine_b200 import scenes as S
    sz, sy, sx = sc["volume"].shape
    h, w = sc["gb"]["depth24"].shape
    with open(path, "wb") as f:
        f.write(np.array([0x4C5856, sx, sy, sz, w, h, n_ao, len(point), len(spot)], "<i4").tobytes())
        v = np.ascontiguousarray(sc["view"], dtype=S.VIEW_DTYPE).reshape(())
        assert v.nbytes == 380
        f.write(v.tobytes())
        f.write(np.ascontiguousarray(sc["volume"], np.uint8).tobytes())
        f.write(np.ascontiguousarray(sc["gb"]["noise"], np.uint32).tobytes())
        f.write(np.ascontiguousarray(point, S.POINT_LIGHT_DTYPE).tobytes())
        f.write(np.ascontiguousarray(spot, S.SPOT_LIGHT_DTYPE).tobytes())


def test_cpp_host_compiles_links_and_fails_loudly_without_a_device(exe, tmp_path):
    import torch
    if torch.cuda.is_available():
        r = subprocess.run([exe, "/dev/null", str(tmp_path / "o.bin"), "9999"], capture_output=True, text=True)   # no such device
    else:
        r = subprocess.run([exe, "/dev/null", str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode == 3, (r.returncode, r.stderr)
    assert ("vxl::Error(-1)" if torch.cuda.is_available() else "vxl::Error(-2)") in r.stderr and "vxl_ctx_create" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("n_point", [2, 64])
def test_cpp_host_planes_equal_the_oracle(exe, oracle, gpu_ctx, tmp_path, n_point):
    from voxelengine_b200 import scenes as S
    sc = U.terrain_scene(oracle)
    h, w = sc["gb"]["depth24"].shape
    sz, sy, sx = sc["volume"].shape
    ext = np.array([2 * sx, 2 * sy, 2 * sz], np.float32) * 0.1
    rs = np.random.RandomState(11)
    pos = [(ext[0] * rs.uniform(0.2, 0.8), ext[1] * rs.uniform(0.5, 0.9), ext[2] * rs.uniform(0.2, 0.8)) for _ in range(n_point)]
    point = S.point_lights(pos, float(ext[0]) * 0.4)
    spot = S.spot_lights([(ext[0] * 0.5, ext[1] * 0.9, ext[2] * 0.5)], float(ext[0]) * 0.6, [(0.0, 1.0, 0.0)], angle=0.6)
    n_ao = 3
    scene, out = str(tmp_path / "scene.bin"), str(tmp_path / "out.bin")
    _write_scene(scene, sc, n_ao, point, spot)
    r = subprocess.run([exe, scene, out], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    raw = np.fromfile(out, np.uint8)
    px = w * h
    gb = raw[:12 * px].view(np.uint32).reshape(3, h, w)
    planes = raw[12 * px:12 * px + 4 * px * (3 + len(point) + len(spot))].view(np.float32).reshape(-1, h, w)
    rays, steps, warned = (int(v) for v in raw[-24:].view(np.uint64))
    # the program synthesised its own G-buffer on the device: the same one the oracle makes for this scene
    for i, k in enumerate(("depth24", "normal", "material")):
        assert np.array_equal(gb[i], sc["gb"][k].view(np.uint32)), k
    wsh, wao, s1 = oracle.pass_ambient(sc["volume"], sc["view"], sc["gb"], n_ao)
    wpt, s2 = oracle.pass_point(sc["volume"], sc["view"], sc["gb"], point)
    wt, s4 = oracle.pass_reflection(sc["volume"], sc["view"], sc["gb"])
    eq = lambda a, b: np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))
    assert eq(planes[0], wsh) and eq(planes[1], wao) and eq(planes[2], wt)
    assert eq(planes[3:3 + len(point)], wpt)
    want_rays, want_steps = s1["rays"] + s2["rays"] + s4["rays"], s1["steps"] + s2["steps"] + s4["steps"]
    if len(spot):
        wsp, s3 = oracle.pass_spot(sc["volume"], sc["view"], sc["gb"], spot)
        assert eq(planes[3 + len(point):], wsp)
        want_rays += s3["rays"]; want_steps += s3["steps"]
    assert (rays, steps) == (want_rays, want_steps)
    assert warned == (70 if n_point == 64 else 0)          # DrawLight beyond MAX_POINT_LIGHTS warns and drops, like the reference


# ---- the tile-sharded frame from C++: two processes, vxl::ShardGroup (include/vxl_pipelines.hpp over vxl_group_*) ----------------------
GROUP_SRC = os.path.join(ROOT, "tests", "cpp", "host_group.cpp")


@pytest.fixture(scope="module")
def group_exe(tmp_path_factory):
    from voxelengine_b200.build import build
    build()
    out = str(tmp_path_factory.mktemp("cppg") / "host_group")
    libdir = os.path.join(ROOT, "voxelengine_b200")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), GROUP_SRC, "-o", out,
                        "-L" + libdir, "-lvxl", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def test_cpp_group_program_compiles_and_fails_loudly_without_a_device(group_exe):
    import torch
    if torch.cuda.is_available():
        pytest.skip("covered by the GPU test")
    r = subprocess.run([group_exe, "32", "64", "32"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 3 and "vxl_ctx_create" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
def test_cpp_two_processes_share_one_gathered_frame(group_exe):
    """Two C++ processes, one vxl::ShardGroup member each: each runs the ambient + reflection passes over its round-robin tile shard; the
    mirrored stores put every tile into BOTH processes' copies of the stack, the flag fence (no collective) closes each of two frames.
    The program checks that the copies are identical, that both ranks' slots are filled and that the frames differ."""
    import torch
    ndev = torch.cuda.device_count()
    r = subprocess.run([group_exe, "128", "512", "256", str(min(ndev, 2))], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("OK"), (r.returncode, r.stdout, r.stderr)

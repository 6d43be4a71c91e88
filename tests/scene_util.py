"""Shared scene builders for the tests: small versions of the BASELINE.json configs, built with
the CPU oracle (volumes, G-buffers) from the host-side inputs of voxelengine_b200.scenes."""
import numpy as np

from voxelengine_b200 import scenes as S


def house_scene(oracle, vol_texels=32, model_size=40, width=96, height=64, frame=3):
    """Config 1 in miniature: one .vox-style model voxelised into a (2*vol_texels)^3-voxel volume
    with the reference voxeliser semantics, camera looking at it, oracle primary G-buffer."""
    vol = np.zeros((vol_texels,) * 3, np.uint8)
    model = S.house_model(model_size, seed=1)
    e = S.entities(1)
    off = (2 * vol_texels - model_size) // 2
    e[0]["cur"] = S.transform_matrix((off * 0.1, 0.2, off * 0.1))
    oracle.voxelize(vol, [model], e)
    ext = 2 * vol_texels * 0.1
    # camera on the -x/-z side (the side facing away from SUN_DIR), looking toward +x+z and down
    view = S.make_view((-ext * 0.2, ext * 0.9, -ext * 0.25), 3.927, -0.5, width, height, frame)
    d, n, m = oracle.gbuffer_primary(vol, view, width, height)
    gb = dict(depth24=d, normal=n, material=m, noise=S.blue_noise(4))
    return dict(volume=vol, view=view, gb=gb, model=model, entities=e)


def terrain_scene(oracle, texels=(64, 48, 64), width=160, height=90, frame=5, props=6):
    """Config 2 in miniature: FastNoise terrain, default camera, oracle primary G-buffer."""
    sx, sy, sz = texels
    vol = oracle.gen_terrain(sx, sy, sz)
    if props:
        model = S.house_model(24, seed=1)
        e = S.prop_entities(vol, n=props, model_size=24, seed=2)
        oracle.voxelize(vol, [model], e)
    view = S.default_camera(texels, width, height, frame)
    d, n, m = oracle.gbuffer_primary(vol, view, width, height)
    gb = dict(depth24=d, normal=n, material=m, noise=S.blue_noise(4))
    lights = S.quarter_point_lights(vol, 4)
    return dict(volume=vol, view=view, gb=gb, lights=lights)


def random_rays(rs, n, extent, dist_lo=20.0, dist_hi=300.0):
    """Random rays around/inside a volume of `extent` voxels per axis (some start outside, some
    axis-parallel, some zero-length) for the ray-level parity interface."""
    r = np.zeros(n, dtype=S.RAY_DTYPE)
    ex = np.asarray(extent, np.float32)
    o = rs.uniform(-0.15, 1.15, size=(n, 3)).astype(np.float32) * ex
    d = rs.normal(size=(n, 3)).astype(np.float32)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-6).astype(np.float32)
    scale = rs.choice([0.25, 1.0, 1.0, 1.0, 3.0], size=(n, 1)).astype(np.float32)
    d *= scale
    k = n // 16
    d[:k, 0] = 0.0                      # axis-parallel component (exercises 0*inf in the DDA)
    d[k:2 * k] = np.round(d[k:2 * k])   # exact axis / diagonal directions
    o[2 * k:3 * k] = np.round(o[2 * k:3 * k])   # origins on voxel boundaries (ties)
    r["ox"], r["oy"], r["oz"] = o[:, 0], o[:, 1], o[:, 2]
    r["dx"], r["dy"], r["dz"] = d[:, 0], d[:, 1], d[:, 2]
    r["dist"] = rs.uniform(dist_lo, dist_hi, size=n).astype(np.float32)
    return r


def voxeliser_case(n=14, seed=5):
    """Entities for the voxeliser checks against the reference's ShadowVoxSystem: moving / rotated / scaled / overlapping /
    glass / partly and entirely out-of-volume entities and a destroy list -> (models, entities, destroy flags)."""
    rs = np.random.RandomState(seed)
    models = [S.house_model(24, seed=1), S.shell_cube_model(16)]
    glass = np.zeros((8, 8, 8), np.uint8); glass[:] = 7; glass[2:6, 2:6, 2:6] = 99      # palette < 16 is not voxelised
    models.append(glass)
    e = S.entities(n)
    for i in range(n):
        pos = rs.uniform(2.0, 16.0, size=3)
        if i == 3:
            pos = np.array([-0.7, 3.0, 5.0])            # partly outside (negative coordinates truncate toward zero)
        if i == 4:
            pos = np.array([200.0, 3.0, 5.0])           # entirely outside: no region
        if i in (6, 7):
            pos = np.array([8.0, 6.0, 8.0]) + 0.3 * i   # overlapping pair: last writer wins per bit
        rot = (0.0, float(rs.uniform(0, 6.28)), 0.0) if i % 3 else (float(rs.uniform(-0.5, 0.5)), float(rs.uniform(0, 6.28)), 0.3)
        e[i]["model"] = i % 3
        e[i]["cur"] = S.transform_matrix(tuple(pos), rot, (1.0, 1.0, 1.0) if i % 4 else (1.5, 0.75, 1.25))
        e[i]["prev"] = S.transform_matrix(tuple(pos + rs.uniform(-0.4, 0.4, size=3)), rot) if i % 2 else e[i]["prev"]
        e[i]["pivot"] = (1.2, 0.0, 1.2) if i % 2 else (0.0, 0.0, 0.0)
    destroy = np.zeros(n, np.int32); destroy[[1, 6, 9]] = 1
    return models, e, destroy


def resolve_case(sc, seed=5):
    """Inputs of the light-buffer resolve for a scene: a random albedo plane, random material bytes
    (roughness / metallic / emit all exercised), and point / spot lights placed at the camera so that most lit pixels
    fall inside their range and cone."""
    gb = dict(sc["gb"])
    h, w = gb["depth24"].shape
    rs = np.random.RandomState(seed)
    albedo = rs.randint(0, 2 ** 32, size=(h, w), dtype=np.uint64).astype(np.uint32)
    gb["material"] = rs.randint(0, 2 ** 32, size=(h, w), dtype=np.uint64).astype(np.uint32)
    cam = np.asarray(sc["view"]["CameraPosition"], np.float64).reshape(3)
    ext = 2 * sc["volume"].shape[2] * 0.1
    centre = np.array([ext * 0.5, ext * 0.25, ext * 0.5])
    aim = (centre - cam) / np.linalg.norm(centre - cam)
    pos = [tuple(cam + aim * 0.3), tuple(cam + np.array([0.4, -0.2, 0.3]))]
    point = S.point_lights(pos, [ext * 2.5, ext * 1.2])
    point["Color"][0] = (3.0, 2.5, 2.0); point["Attenuation"][0] = 1.7
    point["Color"][1] = (0.5, 4.0, 1.0); point["Attenuation"][1] = 2.0
    # LightSpot.frag:132 compares Direction with normalize(light - surface): the cone axis points back along the beam
    spot = S.spot_lights(pos, [ext * 2.5, ext * 1.5], [tuple(-aim), tuple(-aim)])
    spot["Angle"][0] = 0.45; spot["AngleAttenuation"][0] = 1.5; spot["Color"][0] = (4.0, 4.0, 3.0); spot["Attenuation"][0] = 1.2
    spot["Angle"][1] = 0.9; spot["AngleAttenuation"][1] = 0.7; spot["Color"][1] = (1.0, 2.0, 6.0); spot["Attenuation"][1] = 2.2
    return gb, albedo, point, spot


def model_rays(shape, n, seed=1):
    """Fragment inputs of GeometryVoxel.frag for a model of `shape` (sz, sy, sx): cameras outside and inside the box,
    directions toward it (some missing it), and the degenerate cases with exactly zero direction components."""
    sz, sy, sx = shape
    rs = np.random.RandomState(seed)
    ext = np.array([sx, sy, sz], np.float32)
    r = np.zeros(n, S.MODEL_RAY_DTYPE)
    cam = (rs.uniform(-1.5, 2.5, size=(n, 3)) * ext).astype(np.float32)
    inside = rs.uniform(size=n) < 0.2
    cam[inside] = (rs.uniform(0.05, 0.95, size=(int(inside.sum()), 3)) * ext).astype(np.float32)
    far = rs.uniform(size=n) < 0.1                                   # far cameras: the LOD rule accepts coarse voxels
    cam[far] = (cam[far] - ext * 0.5) * np.float32(40.0)
    tgt = (rs.uniform(-0.1, 1.1, size=(n, 3)) * ext).astype(np.float32)
    d = tgt - cam
    k = n // 20
    d[:k, 0] = 0.0
    d[k:2 * k, 1] = 0.0
    d[2 * k:3 * k, 0] = 0.0
    d[2 * k:3 * k, 2] = 0.0
    r["cam"], r["dir"] = cam, d
    r["uv"] = rs.uniform(-1, 1, size=(n, 2)).astype(np.float32)
    return r


def glassy_house(size=40, seed=1):
    """house_model with random palette indices, a quarter of them glass (< 16)."""
    m = S.house_model(size, seed=seed)
    rs = np.random.RandomState(seed + 100)
    solid = m > 0
    m[solid] = rs.randint(1, 64, size=int(solid.sum())).astype(np.uint8)
    return m


def model_scene(width=160, height=96, frame=2):
    """A small draw list for the geometry pass: three models (two sizes), five instances with rotations and overlaps, two palettes."""
    rs = np.random.RandomState(21)
    models = [glassy_house(40, seed=1), glassy_house(24, seed=2), np.full((8, 4, 12), 77, np.uint8)]
    cmds = np.zeros(5, S.VOX_CMD_DTYPE)
    place = [((0.0, 0.0, 0.0), (0.0, 0.3, 0.0), 0), ((3.0, 0.2, 1.0), (0.1, -0.8, 0.05), 1), ((1.0, 0.5, 1.5), (0.0, 1.1, 0.0), 1),   # overlaps the first
             ((-1.5, 1.0, 2.5), (0.4, 0.2, -0.3), 2), ((2.0, 3.0, 2.0), (0.0, 0.0, 0.0), 0)]
    for i, (pos, rot, mi) in enumerate(place):
        cmds[i]["WorldMatrix"] = S.transform_matrix(pos, rot)
        cmds[i]["LastWorldMatrix"] = S.transform_matrix((pos[0] + 0.02, pos[1], pos[2] - 0.01), rot)
        cmds[i]["VolumeRID"] = 3 + i
        cmds[i]["PalleteIndex"] = i % 2
        cmds[i]["model"] = mi
    pal_c = rs.randint(0, 2 ** 32, size=(2, 256), dtype=np.uint64).astype(np.uint32)
    pal_m = rs.randint(0, 2 ** 32, size=(2, 256), dtype=np.uint64).astype(np.uint32)
    view = S.make_view((7.5, 5.0, -4.0), 2.45, -0.45, width, height, frame)
    return models, cmds, pal_c, pal_m, view


def f3_case(oracle, sc, seed=9):
    """Inputs of LightTAA.frag / LightReflection.frag's colour for a scene: block-constant albedo and material
    (so that the neighbour weights are a mix of zeros and non-zeros), a smooth motion field with a discontinuity and a band whose
    history falls outside the frame, a current light buffer (the oracle's ambient colour plus noise), a perturbed history with
    variance in alpha, and the reflection march's t plane."""
    gb = dict(sc["gb"])
    h, w = gb["depth24"].shape
    rs = np.random.RandomState(seed)
    by, bx = np.meshgrid(np.arange(h) // 6, np.arange(w) // 8, indexing="ij")
    pal_a = rs.randint(0, 2 ** 32, size=64, dtype=np.uint64).astype(np.uint32)
    pal_m = (rs.randint(0, 64, size=(64, 4)).astype(np.uint32) * np.array([1, 1 << 8, 1 << 16, 1 << 24], np.uint32)).sum(axis=1).astype(np.uint32)
    blk = (by * 3 + bx) % 5
    albedo = pal_a[blk]
    gb["material"] = (pal_m[blk] + ((by % 2).astype(np.uint32) * np.uint32(40))).astype(np.uint32)     # rows of blocks differ by 40/255 in roughness (< 0.2)
    yy, xx = np.meshgrid(np.arange(h, dtype=np.float32), np.arange(w, dtype=np.float32), indexing="ij")
    motion = np.zeros((h, w, 2), np.float32)
    motion[..., 0] = 0.004 * np.sin(xx * 0.2) + 0.002
    motion[..., 1] = 0.003 * np.cos(yy * 0.3)
    motion[h // 2:, w // 2:, 0] += np.float32(0.15)                  # discontinuity: neighbours across it are rejected (> 0.1)
    motion[: h // 6, :, 1] -= np.float32(0.5)                        # history above the frame: the current-frame-only branch
    sh, ao, _ = oracle.pass_ambient(sc["volume"], sc["view"], gb, 1)
    light = oracle.resolve_ambient(sc["view"], gb, albedo, sh, ao)
    light[..., :3] += rs.rand(h, w, 3).astype(np.float32) * np.float32(0.05)
    light[..., 3] = rs.rand(h, w).astype(np.float32)
    last = light * (np.float32(0.7) + np.float32(0.6) * rs.rand(h, w, 1).astype(np.float32))
    last[..., 3] = rs.rand(h, w).astype(np.float32) * np.float32(1.2)
    t, _ = oracle.pass_reflection(sc["volume"], sc["view"], gb)
    return dict(gb=gb, albedo=albedo, motion=motion, light=light.astype(np.float32), last_light=last.astype(np.float32), t=t, sky=(0.3, 0.5, 0.9))

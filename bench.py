#!/usr/bin/env python
"""bench.py -- shadow + AO + specular-occlusion throughput of the voxel-lighting pass.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 3] [--impl ours|reference]

A "step" is one frame of the light passes over the workload BASELINE.json's metric is quoted on:
config 3 = 1024^3-voxel FastNoise terrain + 200 props, 3840x2160, per lit pixel 1 sun-shadow ray +
16 AO rays + up to 4 point-light shadow rays + 1 specular-occlusion ray (synthetic, seeded).
N > 1 (torchrun, one rank per GPU): the same frame partitioned into 128x128 screen tiles dealt
round-robin to the ranks, volume replicated, one NCCL all-gather of the output tiles per step.

Prints ONE JSON line (rank 0).  `value` = rays actually generated per second with inputs resident
in HBM; `e2e` = the same through the host-buffer C-ABI call (H2D + passes + D2H per step);
`roofline` = algorithmic bytes of the dominant kernel / its CUDA-event time / measured HBM peak;
`cpu_baseline` = the CPU oracle (oracle/liboracle.so) timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "shadow+AO+spec-occlusion Mrays/s"
UNIT = "Mrays/s"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def issue_roofline(kernel: str, kernel_ms: float, sm_mhz: float, n_sm: int):
    """The bound the ncu captures point at: warp instructions of one launch (profiles/kernel_counters.json, ncu capture of
    the same kernel on the same workload) against the SMs' issue rate, 4 warp instructions per clock per SM, at the live kernel time and
    the SM clock sampled during the timed region.  None when the counters are missing."""
    path = os.path.join(ROOT, "profiles", "kernel_counters.json")
    if not os.path.exists(path) or not sm_mhz or not n_sm or kernel_ms <= 0:
        return None
    c = json.load(open(path)).get(kernel.split("<")[0])
    if not c or not c.get("warp_instructions"):
        return None
    peak = float(n_sm) * 4.0 * float(sm_mhz) * 1e6
    ach = float(c["warp_instructions"]) / (kernel_ms * 1e-3)
    return {"warp_instructions_per_launch": c["warp_instructions"], "active_threads_per_instruction": c.get("active_threads_per_inst"),
            "achieved_warp_inst_per_s": ach, "peak_warp_inst_per_s": peak, "frac": ach / peak,
            "source": "instruction count: profiles/kernel_counters.json (ncu, same kernel and workload); time and clock: this run"}


def config_of(cfg: dict) -> dict:
    """The `config` object of the JSON line: the same keys and values from both arms (ours and --impl reference)."""
    W, H = cfg["res"]
    return {"workload": cfg["name"], "volume_texels": list(cfg["texels"]), "resolution": [W, H],
            "rays_per_lit_pixel": {"sun_shadow": 1, "ao": cfg["n_ao"], "point_light_shadow": cfg["n_point"], "spec_occlusion": int(cfg["spec"])},
            "l2": "GPU arm: flushed between timed steps (256 MiB fill outside the event pairs); CPU arm: not applicable"}


def dram_traffic(kernel: str, config: int, world: int):
    """ncu `dram__bytes_read.sum + dram__bytes_write.sum` of one launch of `kernel`, from the committed capture of THIS config at
    THIS GPU count (profiles/dram_traffic.json: {"cfg<config>_n<world>": {kernel: bytes}}); None when no such capture exists."""
    tp = os.path.join(ROOT, "profiles", "dram_traffic.json")
    try:
        return json.load(open(tp)).get(f"cfg{config}_n{world}", {}).get(kernel)
    except Exception:
        return None


def algorithmic_bytes(pass_name: str, st: dict, n_ao: int) -> int:
    """sum_rays steps*1 B + sum_lit_pixels (in_px + out_px).  in_px = depth 4 + normal 4 + 4 per
    distinct blue-noise texel (+ material 4 for the spec pass); out_px = 4 B per output scalar."""
    if pass_name == "ambient":
        return st["steps"] + st["pixels"] * (4 + 4 + 4 * max(n_ao, 1) + 8)
    if pass_name == "point":
        return st["steps"] + st["pixels"] * (4 + 4 + 4) + st["rays"] * 4
    return st["steps"] + st["pixels"] * (4 + 4 + 4 + 4 + 4)


# ------------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle (a port of the reference shaders; the reference itself is Vulkan/Windows
# only and cannot run here) on all host cores, inputs generated on the CPU as well.
# ------------------------------------------------------------------------------------------------------
def cpu_frame(O, vol, view, gb, lights, cfg, rows, keep=None):
    """One pass of the oracle over `rows`; keep (a dict) receives the planes it computed (for the parity check)."""
    t0 = time.perf_counter()
    rays = steps = 0
    sh, ao, st = O.pass_ambient(vol, view, gb, cfg["n_ao"], rows=rows)
    rays += st["rays"]; steps += st["steps"]
    pt = sp = None
    if cfg["n_point"]:
        pt, st = O.pass_point(vol, view, gb, lights, rows=rows)
        rays += st["rays"]; steps += st["steps"]
    if cfg["spec"]:
        sp, st = O.pass_reflection(vol, view, gb, rows=rows)
        rays += st["rays"]; steps += st["steps"]
    dt = time.perf_counter() - t0
    if keep is not None:
        keep.update(shadow=sh, ao=ao, point=pt, spec_t=sp)
    return rays, steps, dt


def pick_rows(O, vol, view, gb, lights, cfg, H, target_s=12.0):
    """Bounded sample: every k-th row, k chosen from a 1/64 probe so that one pass over the sample takes ~target_s."""
    r, s, dt = cpu_frame(O, vol, view, gb, lights, cfg, (0, H, 64))
    est_full = dt * 64
    k = int(max(1, min(64, np.ceil(est_full / target_s))))
    return (0, H, k)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import vxo_py as O
    from voxelengine_b200 import scenes as S
    from voxelengine_b200.workloads import CONFIGS
    O.set_num_threads(len(os.sched_getaffinity(0)))     # all host threads (torchrun exports OMP_NUM_THREADS=1)
    cfg = CONFIGS[args.config]
    sx, sy, sz = cfg["texels"]
    W, H = cfg["res"]
    if cfg["scene"] == "house":
        vol = np.zeros((sz, sy, sx), np.uint8)
        e = S.entities(1)
        off = (2 * sx - 40) // 2
        e[0]["cur"] = S.transform_matrix((off * 0.1, 0.2, off * 0.1))
        O.voxelize(vol, [S.house_model(40, 1)], e)
        ext = 2 * sx * 0.1
        view = S.make_view((-ext * 0.2, ext * 0.9, -ext * 0.25), 3.927, -0.5, W, H, 0)
    else:
        vol = O.gen_terrain(sx, sy, sz)
        if cfg["scene"] == "terrain+props":
            O.voxelize(vol, [S.house_model(40, 1)], S.prop_entities(vol, n=cfg["n_props"], model_size=40, seed=2))
        view = S.default_camera(cfg["texels"], W, H, 0)
    d, n, m = O.gbuffer_primary(vol, view, W, H)
    gb = dict(depth24=d, normal=n, material=m, noise=S.blue_noise(4))
    lights = S.quarter_point_lights(vol, cfg["n_point"]) if cfg["n_point"] else None
    rows = pick_rows(O, vol, view, gb, lights, cfg, H, target_s=max(2.0, 150.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_frame(O, vol, view, gb, lights, cfg, rows)
    tot_r = tot_t = 0.0
    for _ in range(args.steps):
        r, s, dt = cpu_frame(O, vol, view, gb, lights, cfg, rows)
        tot_r += r; tot_t += dt
    val = tot_r / tot_t / 1e6
    sample = f"rows {rows[0]}:{rows[1]}:{rows[2]} of the {W}x{H} frame per step ({int(tot_r / max(args.steps, 1))} rays)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32+u8", "data": "synthetic", "config": config_of(cfg),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": O.num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from voxelengine_b200.build import build
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if local == 0:
        build()
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    from voxelengine_b200 import engine as E
    from voxelengine_b200.workloads import Workload

    wl = Workload(args.config, rank=rank, world=world, device=local, gather=args.gather)
    cfg = wl.cfg
    W, H = wl.res
    dev = wl.ctx.torch_device
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def flush():
        flush_buf.fill_(rank & 0xFF)

    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        return float(t.item())

    def allmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # per-frame accounting, counted by the kernels themselves (deterministic per frame)
    st_local = wl.count()
    rays = allsum(float(st_local["rays"]))
    probes = allsum(float(st_local["steps"]))
    lit = allsum(float(st_local["pixels"]))

    for _ in range(max(args.warmup, 3)):
        wl.step()
    torch.cuda.synchronize()

    # ---- timed region: K steps, CUDA events on the launching stream, L2 flushed between steps ----
    launches0 = wl.ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local) as clk:
        for a, b in ev:
            flush()
            a.record()
            wl.step()
            b.record()
        torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = wl.ctx.launch_count() - launches0
    t_ms = allmax(sum(a.elapsed_time(b) for a, b in ev))
    ms_per_step = t_ms / args.steps
    value = rays / (ms_per_step * 1e-3) / 1e6

    if world > 1:   # outside the timed region: what every rank holds after a step
        if wl.stack is not None:
            # fused gather: every rank's copy of the stack must equal an NCCL all-gather of the ranks' own tiles
            from voxelengine_b200.tiles import gather_tiles
            ref = gather_tiles(wl.out.clone())
            torch.cuda.synchronize()
            assert torch.equal(ref, wl.stack.tensor), "fused gather: the stack assembled by peer stores differs from an NCCL all-gather"
            del ref
        else:
            assert torch.equal(wl.gathered_main[rank], wl.out[:3]) and (not wl.n_point or torch.equal(wl.gathered_point[rank], wl.out[3:])), "gather mismatch"

    # warm-L2 variant (no flush), reported beside the flushed figure
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev2:
        a.record(); wl.step(); b.record()
    torch.cuda.synchronize()
    ms_warm = allmax(sum(a.elapsed_time(b) for a, b in ev2)) / args.steps

    # plain march on the volume bytes (kernel variant 0) on the same frame, for the record: same results
    wl.ctx.set_variant(0)
    wl.step(); torch.cuda.synchronize()
    ev3 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(3)]
    for a, b in ev3:
        flush(); a.record(); wl.step(); b.record()
    torch.cuda.synchronize()
    ms_plain = allmax(sum(a.elapsed_time(b) for a, b in ev3)) / 3
    wl.ctx.set_variant(2)                     # counting twin of the default kernels
    wl.ctx.stats_reset(); wl.step(gather=False)
    fetched_probes = allsum(float(wl.ctx.fetched_probes()))
    wl.ctx.set_variant(1)

    # ---- per-kernel times (rank 0's shard) for the roofline of the dominant kernel ----
    per = wl.per_pass_counts()
    n = wl.gb.n_tiles
    o = wl.out
    tmp_pt = wl.ctx.empty((max(wl.n_point, 1), n, wl.gb.tile_h, wl.gb.tile_w), torch.float32)
    calls = {"ambient": lambda: E.LightAmbientPipeline.Get().Use(wl.view, wl.gb, wl.vol, n_ao=wl.n_ao, out_shadow=o[0, :n], out_ao=o[1, :n])}
    if wl.n_point:
        calls["point"] = lambda: wl._point(tmp_pt)
    if wl.spec:
        calls["reflection"] = lambda: E.LightReflectionPipeline.Get().Use(wl.view, wl.gb, wl.vol, out_spec_t=o[2, :n])
    ktime = {}
    for name, fn in calls.items():
        tt = 0.0
        for _ in range(args.steps):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            tt += a.elapsed_time(b)
        ktime[name] = tt / args.steps
    dom = max(ktime, key=ktime.get)
    peak, peak_src = hbm_peak()
    abytes = algorithmic_bytes(dom, per[dom], wl.n_ao)
    achieved = abytes / (ktime[dom] * 1e-3) / 1e9
    traffic = dram_traffic({"ambient": "k_ambient", "point": "k_local_lights", "reflection": "k_reflection"}[dom], args.config, world)
    roofline = {"bound": "hbm", "kernel": {"ambient": "k_ambient", "point": "k_local_lights<false>", "reflection": "k_reflection"}[dom],
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": abytes, "kernel_ms": ktime[dom],
                "all_kernels_ms": ktime, "probes_per_s": per[dom]["steps"] / (ktime[dom] * 1e-3)}
    # The volume the probes read is L2-resident (98 % L2 hit rate in the ncu captures), so north_star's "HBM (or L2) roofline"
    # is also quoted against the L2 read bandwidth -- measured live on this GPU (SURVEY 8d / BASELINE.md: "by a microbench in the
    # bench harness"): 16-byte .cg loads over a 48 MiB buffer from every SM, and the same over 2 GiB for HBM reads.
    try:
        l2_gbs = max(wl.ctx.read_bandwidth(48 << 20, 40) for _ in range(3))
        hbm_read_gbs = max(wl.ctx.read_bandwidth(2 << 30, 2) for _ in range(2))
        roofline["l2"] = {"peak": l2_gbs, "unit": "GB/s", "achieved": achieved, "frac": achieved / l2_gbs,
                          "peak_source": "measured in this run: vxl_debug_read_bandwidth, 48 MiB buffer, best of 3",
                          "hbm_read_gbs_same_probe": hbm_read_gbs}
    except Exception as e:
        roofline["l2"] = None
        sys.stderr.write(f"L2 bandwidth probe skipped: {e}\n")

    # ---- light-buffer resolve: elementwise, HBM-bound; timed beside the march, not part of `value` ----
    resolve = None
    if world == 1:
        alb = torch.randint(0, 2 ** 31 - 1, wl.gb.shape, dtype=torch.int32, device=dev)
        lb = E.LightBuffer(wl.gb, alb)
        pl = wl.planes()

        def res():
            lb.Ambient(wl.view, pl["shadow"], pl["ao"])
            if wl.n_point:
                lb.Point(wl.view, wl.lights, pl["point"])
        res(); torch.cuda.synchronize()
        tt = 0.0
        for _ in range(args.steps):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); res(); b.record()
            torch.cuda.synchronize()
            tt += a.elapsed_time(b)
        npx = int(np.prod(wl.gb.shape))
        # per pixel: depth, normal, material, albedo, shadow, ao in (24 B) + rgba out (16 B); point: the same planes minus ao
        # plus one shadow plane per light in and rgba in + out
        rbytes = npx * (24 + 16) + (npx * (16 + 4 * wl.n_point + 32) if wl.n_point else 0)
        resolve = {"ms": tt / args.steps, "bytes": rbytes, "GB/s": rbytes / (tt / args.steps * 1e-3) / 1e9,
                   "frac_of_hbm_peak": rbytes / (tt / args.steps * 1e-3) / 1e9 / peak,
                   "kernels": "k_resolve_ambient" + (" + k_resolve_local<point>" if wl.n_point else "")}

    # ---- the rows either side of the path, timed beside the march, not part of `value` ----
    def timed(fn):
        fn(); torch.cuda.synchronize()
        tt = 0.0
        for _ in range(args.steps):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record()
            torch.cuda.synchronize()
            tt += a.elapsed_time(b)
        return tt / args.steps

    post = geom = None
    if world == 1 and resolve is not None:
        npx = int(np.prod(wl.gb.shape))
        g = torch.Generator(device=dev); g.manual_seed(7)
        motion = (torch.rand((H, W, 2), device=dev, generator=g) - 0.5) * 0.004
        full = E.FullFrame(wl.ctx, wl.gb.depth24[0], wl.gb.normal[0], wl.gb.material[0], alb[0], motion)
        light = lb.rgba[0].clone()
        last = light * 0.9
        last[..., 3] = torch.rand((H, W), device=dev, generator=g)
        taa_out = torch.zeros_like(lb.rgba)
        refl_out = torch.zeros_like(lb.rgba)
        t_taa = timed(lambda: E.LightTAAPipeline.Get().Use(wl.view, wl.gb, full, light, last, out=taa_out))
        # per pixel: centre texel of 6 planes (40 B) + history (16 B) + 12 taps x 40 B + rgba out (16 B); the taps fall within
        # +-13 pixels, so the compulsory HBM traffic is one pass over the planes (40 + 16 + 16 B)
        taa_alg, taa_min = npx * (40 + 16 + 12 * 40 + 16), npx * (40 + 16 + 16)
        post = {"taa_ms": t_taa, "taa_algorithmic_GB/s": taa_alg / (t_taa * 1e-3) / 1e9, "taa_compulsory_GB/s": taa_min / (t_taa * 1e-3) / 1e9,
                "taa_compulsory_frac_of_hbm_peak": taa_min / (t_taa * 1e-3) / 1e9 / peak, "kernels": "k_light_taa"}
        if wl.spec:
            t_rf = timed(lambda: E.LightReflectionPipeline.Get().Colour(wl.view, wl.gb, pl["spec_t"], full, taa_out[0], (0.3, 0.5, 0.9), out=refl_out))
            rf_bytes = npx * (12 + 4 + 4 + 16 + 16)                 # depth / normal / material + t + end-point depth + light + rgba out
            post.update({"reflection_colour_ms": t_rf, "reflection_colour_GB/s": rf_bytes / (t_rf * 1e-3) / 1e9,
                         "reflection_colour_frac_of_hbm_peak": rf_bytes / (t_rf * 1e-3) / 1e9 / peak, "kernels": "k_light_taa + k_resolve_reflection"})
        if wl.props is not None:
            from voxelengine_b200 import scenes as S
            cmds = np.zeros(len(wl.props), S.VOX_CMD_DTYPE)
            cmds["WorldMatrix"] = wl.props["cur"]; cmds["LastWorldMatrix"] = wl.props["cur"]
            cmds["VolumeRID"] = 3 + np.arange(len(cmds)); cmds["model"] = wl.props["model"]
            pal = torch.randint(0, 2 ** 31 - 1, (1, 256), dtype=torch.int32, device=dev)
            gfb = E.GeometryBuffer(wl.ctx, W, H)
            galb = torch.zeros(gfb.shape, dtype=torch.int32, device=dev)
            t_g = timed(lambda: E.GeometryVoxelPipeline.Get().Use(wl.view, gfb, cmds, pal, pal, albedo=galb))
            covered = int((gfb.depth24 != 0xFFFFFF).sum().item())
            geom = {"ms": t_g, "draws": int(len(cmds)), "covered_pixels": covered, "Mpixels/s": npx / (t_g * 1e-3) / 1e6,
                    "kernels": "k_gbuffer_models", "note": "the config's instanced models only (the terrain is not a model)"}

    # ---- end to end: host buffers in, host buffers out, every step ----
    e2e = None if args.no_e2e else run_e2e(args, wl, torch, dist, world, rank, rays)

    # ---- CPU baseline (rank 0, N=1 only): the oracle on a bounded sample of the same frame ----
    cpu = parity = None
    if world == 1 and not args.no_cpu:
        from oracle import vxo_py as O
        O.set_num_threads(len(os.sched_getaffinity(0)))
        vol_h = wl.host_volume if wl.host_volume is not None else wl.vol.download()
        gbh = {k: getattr(wl.gb, k).cpu().numpy().view(np.uint32)[0] for k in ("depth24", "normal", "material")}
        gbh["noise"] = wl.gb.noise.cpu().numpy().view(np.uint32)
        rows = pick_rows(O, vol_h, wl.view, gbh, wl.lights, cfg, H, target_s=15.0)
        want = {}
        r, s, dt = cpu_frame(O, vol_h, wl.view, gbh, wl.lights, cfg, rows, keep=want)
        cpu = {"value": r / dt / 1e6, "unit": UNIT, "cores": O.num_threads(), "kind": "port",
               "sample": f"rows {rows[0]}:{rows[1]}:{rows[2]} of the {W}x{H} frame ({r} rays, {dt:.1f} s)"}
        # The oracle's planes are in hand: check the frame the GPU just timed against them, bit for bit, at the bench's full size.
        if True:
            wl.ctx.stats_reset()
            wl.step(gather=False, advance=False); torch.cuda.synchronize()       # dynamic scene: the frame whose volume was just downloaded
            st_now = wl.ctx.stats()
            got = wl.assemble()
            sel = slice(rows[0], rows[1], rows[2])
            planes = {"shadow": (got[0], want["shadow"]), "ao": (got[1], want["ao"])}
            if wl.spec:
                planes["spec_t"] = (got[2], want["spec_t"])
            for li in range(wl.n_point):
                planes[f"point{li}"] = (got[3 + li], want["point"][li])
            res_p = {k: bool(np.array_equal(g[sel].view(np.uint32), w[sel].view(np.uint32))) for k, (g, w) in planes.items()}
            parity = {"checker": "oracle (cpu_baseline leg)", "rows": f"{rows[0]}:{rows[1]}:{rows[2]}", "pixels": int(got[0][sel].size),
                      "planes_bit_exact": res_p, "bit_exact": all(res_p.values())}
            if rows[2] == 1:
                parity["rays_equal"] = bool(int(st_now["rays"]) == int(r))
                parity["probes_equal"] = bool(int(st_now["steps"]) == int(s))
            ok = parity["bit_exact"] and parity.get("rays_equal", True) and parity.get("probes_equal", True)
            if not ok:
                # a frame that differs from the oracle has no throughput: no `value`, non-zero exit
                sys.stderr.write(f"PARITY FAILURE at full size: {parity}\n")
                print(json.dumps({"metric": METRIC, "error": "parity failure against the CPU oracle at the bench's own size", "parity": parity,
                                  "n_gpus": world, "config": config_of(cfg)}))
                wl.close()
                sys.exit(3)

    if world == 1 and args.config == 3:      # the instruction counts on file are those of config 3's whole-frame launches
        try:
            roofline["issue"] = issue_roofline(roofline["kernel"], roofline["kernel_ms"], clk.summary().get("sm_mhz"),
                                               torch.cuda.get_device_properties(dev).multi_processor_count)
        except Exception as e:               # never let a derived figure cost the bench line
            roofline["issue"] = None
            sys.stderr.write(f"issue roofline skipped: {e}\n")
    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+u8",
            "data": "synthetic",
            "config": config_of(cfg),
            "detail": {"rays_per_step": int(rays), "probes_per_step": int(probes), "lit_pixels": int(lit),
                       "parallelism": "1 GPU, whole frame" if world == 1 else f"{world} GPUs, 128x128 screen tiles round-robin, volume replicated, " + (
                           "output tiles gathered INSIDE the pass kernels: every output store repeated into each peer's stack over NVLink peer memory (CUDA IPC), one fence per frame"
                           if wl.stack is not None else "NCCL all-gather of output tiles"),
                       "gather": wl.gather_mode,
                       "l2": "flushed between timed steps (256 MiB fill outside the event pairs); per-step working set 128 MiB volume + 100 MB G-buffer + 232 MB outputs",
                       "ms_per_step_warm_l2": ms_warm, "ms_per_step_plain_march_variant0": ms_plain,
                       "probes_that_read_the_volume": int(fetched_probes)},
            "roofline": roofline, "light_buffer_resolve": resolve, "post_passes": post, "geometry_pass": geom, "cpu_baseline": cpu, "parity": parity, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk.summary(),
        }))
    wl.close()
    if world > 1:
        dist.destroy_process_group()


def run_e2e(args, wl, torch, dist, world, rank, rays):
    """Same metric through the host-facing call.  N=1: vxl_lighting_host (pinned host G-buffer in, pinned host planes
    out).  N>1: each rank uploads its tile shard from pinned memory, runs the passes, all-gathers on the device, and
    rank 0 reads the assembled tiles back."""
    from voxelengine_b200 import engine as E
    steps = max(3, min(args.steps, 10))
    n = wl.gb.n_tiles
    shape = (n, wl.gb.tile_h, wl.gb.tile_w)
    planes = {k: getattr(wl.gb, k).cpu().pin_memory() for k in ("depth24", "normal", "material")}
    planes["noise"] = wl.gb.noise.cpu().pin_memory()
    px = int(np.prod(shape))
    h2d = 3 * px * 4 + 512 * 512 * 4
    if world == 1:
        outs = dict(shadow=torch.empty(shape, dtype=torch.float32).pin_memory(), ao=torch.empty(shape, dtype=torch.float32).pin_memory())
        if wl.spec:
            outs["spec_t"] = torch.empty(shape, dtype=torch.float32).pin_memory()
        if wl.n_point:
            outs["point_shadow"] = torch.empty((wl.n_point,) + shape, dtype=torch.float32).pin_memory()
        d2h = sum(int(t.numel()) * 4 for t in outs.values())
        # packed planes (vxl_lighting_host_packed): one mask byte per 8 shadow planes, one code byte for spec_t, AO stays float32
        mb = E.mask_bytes(wl.n_point, 0)
        pk = dict(shadow_mask=torch.empty(shape + (mb,), dtype=torch.uint8).pin_memory(), ao=torch.empty(shape, dtype=torch.float32).pin_memory())
        if wl.spec:
            pk["spec_code"] = torch.empty(shape, dtype=torch.uint8).pin_memory()
        d2h_packed = sum(int(t.numel()) * t.element_size() for t in pk.values())
        desc = dict(width=wl.gb.width, height=wl.gb.height, tile_w=wl.gb.tile_w, tile_h=wl.gb.tile_h, tile_first=wl.gb.tile_first,
                    tile_stride=wl.gb.tile_stride, n_tiles=n)

        dynamic = wl.cfg["scene"] == "dynamic"

        def timed_host(o):
            def one():
                if dynamic:      # config 4: the frame starts with the re-voxelisation of the moving entities + occupancy rebuild
                    wl.advance()
                E.lighting_host(wl.ctx, wl.vol, wl.view, desc, planes, o, n_ao=wl.n_ao, point=wl.lights)
            for _ in range(2):
                one()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(steps):
                one()
            torch.cuda.synchronize()
            return (time.perf_counter() - t0) / steps
        dt_float = timed_host(outs)
        dt = timed_host({"packed": pk})
        # both must equal the resident path (same volume: the dynamic scene is not advanced for the check)
        if dynamic:
            E.lighting_host(wl.ctx, wl.vol, wl.view, desc, planes, outs, n_ao=wl.n_ao, point=wl.lights)
            E.lighting_host(wl.ctx, wl.vol, wl.view, desc, planes, {"packed": pk}, n_ao=wl.n_ao, point=wl.lights)
        ref = wl.step(gather=False, advance=False).cpu()
        assert torch.equal(outs["shadow"], ref[0, :n]) and torch.equal(outs["ao"], ref[1, :n]), "e2e planes differ from the resident path"
        un = E.unpack_planes(pk["shadow_mask"].numpy(), pk["spec_code"].numpy() if wl.spec else None, wl.n_point, 0)
        assert torch.equal(pk["ao"], ref[1, :n]), "e2e (packed): ao differs from the resident path"
        assert np.array_equal(un["shadow"].view(np.uint32), ref[0, :n].numpy().reshape(-1).view(np.uint32)), "e2e (packed): sun shadow differs"
        if wl.spec:
            assert np.array_equal(un["spec_t"].view(np.uint32), ref[2, :n].numpy().reshape(-1).view(np.uint32)), "e2e (packed): spec_t differs"
        for li in range(wl.n_point):
            assert np.array_equal(un["point_shadow"][li].view(np.uint32), ref[3 + li, :n].numpy().reshape(-1).view(np.uint32)), f"e2e (packed): point plane {li} differs"
        return {"value": rays / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_packed, "ms_per_step": dt * 1e3,
                "api": "vxl_lighting_host_packed (pinned host buffers; shadow planes as a bit mask, spec_t as its one-byte code, AO float32; decoded and compared "
                       "bit for bit with the resident path after the timed region)",
                "float_planes": {"value": rays / dt_float / 1e6, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": dt_float * 1e3,
                                 "api": "vxl_lighting_host (pinned host buffers, every plane float32)"}}
    # N > 1: every rank runs the same host-facing call on its tile shard, and its outputs land -- over its own PCIe link -- in ONE
    # host frame shared by all ranks (POSIX shared memory, page-locked by each rank with cudaHostRegister): [rank][plane][tile].
    # No rank reads back another rank's tiles; the NCCL all-gather belongs to the device-resident path (`value`).
    from multiprocessing import shared_memory
    n_planes = wl.n_planes
    slot = n_planes * wl.tiles_padded * wl.gb.tile_h * wl.gb.tile_w          # floats per rank (padded to the largest shard)
    name = f"vxl_e2e_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"
    if rank == 0:
        try:
            shared_memory.SharedMemory(name=name).unlink()
        except FileNotFoundError:
            pass
        shm = shared_memory.SharedMemory(name=name, create=True, size=world * slot * 4 + world * 64)
        shm.buf[world * slot * 4:world * slot * 4 + world * 64] = bytes(world * 64)
    dist.barrier()
    if rank != 0:
        shm = shared_memory.SharedMemory(name=name)
        try:                                   # rank 0 owns the segment: keep this process's resource tracker from unlinking it again at exit
            from multiprocessing import resource_tracker
            resource_tracker.unregister(shm._name, "shared_memory")
        except Exception:
            pass
    frame_host = np.ndarray((world, slot), dtype=np.float32, buffer=shm.buf)
    # the frame's host-side barrier lives in the same segment: one counter per rank, a cache line apart (aligned 8-byte stores and
    # loads; every rank publishes the number of frames it has landed and waits until all have) -- a few microseconds where an NCCL
    # barrier is a kernel launch plus a stream synchronisation per frame
    arrivals = np.ndarray((world, 8), dtype=np.int64, buffer=shm.buf, offset=world * slot * 4)
    landed = [0]

    def frame_barrier():
        landed[0] += 1
        arrivals[rank, 0] = landed[0]
        t_end = time.perf_counter() + 30.0
        while int(arrivals[:, 0].min()) < landed[0]:
            if time.perf_counter() > t_end:
                raise RuntimeError("e2e: a rank did not land its frame within 30 s")
    mine = torch.from_numpy(frame_host[rank])
    rc = torch.cuda.cudart().cudaHostRegister(mine.data_ptr(), slot * 4, 0)
    assert int(rc) == 0, f"cudaHostRegister failed: {rc}"
    plane = n * wl.gb.tile_h * wl.gb.tile_w
    # the packed frame of the same call (vxl_lighting_host_packed) shares the rank's slot: [ao float32][shadow mask][spec code]; the
    # float planes are timed first, verified, and only then overwritten
    mb = E.mask_bytes(wl.n_point, 0)
    assert plane * (4 + mb + 1) <= slot * 4
    mine_b = torch.from_numpy(frame_host[rank].view(np.uint8))
    pk = dict(ao=mine[0:plane].view(shape), shadow_mask=mine_b[4 * plane:(4 + mb) * plane].view(shape + (mb,)))
    if wl.spec:
        pk["spec_code"] = mine_b[(4 + mb) * plane:(5 + mb) * plane].view(shape)
    d2h_packed = sum(int(t.numel()) * t.element_size() for t in pk.values())
    outs = dict(shadow=mine[0:plane], ao=mine[plane:2 * plane])
    if wl.spec:
        outs["spec_t"] = mine[2 * plane:3 * plane]
    if wl.n_point:
        outs["point_shadow"] = mine[3 * plane:(3 + wl.n_point) * plane]
    d2h = sum(int(t.numel()) * 4 for t in outs.values())
    desc = dict(width=wl.gb.width, height=wl.gb.height, tile_w=wl.gb.tile_w, tile_h=wl.gb.tile_h, tile_first=wl.gb.tile_first,
                tile_stride=wl.gb.tile_stride, n_tiles=n)

    def timed_host(o):
        def one():
            if n:
                E.lighting_host(wl.ctx, wl.vol, wl.view, desc, planes, o, n_ao=wl.n_ao, point=wl.lights)   # blocks until the planes are in host memory
        for _ in range(2):
            one()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
            frame_barrier()                 # the frame is complete when every rank's tiles have landed
        return (time.perf_counter() - t0) / steps
    dt = timed_host(outs)
    # every rank's region of the shared frame must equal the device-resident result of that rank (checked by rank 0 through the gather)
    wl.step(gather=True); torch.cuda.synchronize()
    counts = [torch.zeros(1, dtype=torch.int64, device=wl.ctx.torch_device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([n], dtype=torch.int64, device=wl.ctx.torch_device))
    if rank == 0:
        g = wl.gathered.cpu().numpy()                                         # (world, n_planes, tiles_padded, th, tw)
        for r in range(world):
            nr = int(counts[r].item())
            pr = nr * wl.gb.tile_h * wl.gb.tile_w
            want = np.concatenate([g[r, 0, :nr].ravel(), g[r, 1, :nr].ravel()] + ([g[r, 2, :nr].ravel()] if wl.spec else [np.zeros(0, np.float32)]))
            got = frame_host[r, :2 * pr] if not wl.spec else frame_host[r, :3 * pr]
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"e2e: rank {r}'s tiles in the shared host frame differ from the resident path"
            if wl.n_point:
                wantp = g[r, 3:, :nr].reshape(-1)
                assert np.array_equal(frame_host[r, 3 * pr:(3 + wl.n_point) * pr].view(np.uint32), wantp.view(np.uint32)), f"e2e: rank {r} point planes differ"
    dist.barrier()
    # packed planes: every rank keeps a copy of its float planes, times the packed call into the same slot and decodes it
    if n:
        keep = {k: v.clone() for k, v in outs.items()}
    dt_packed = timed_host({"packed": pk})
    if n:
        un = E.unpack_planes(pk["shadow_mask"].numpy(), pk["spec_code"].numpy() if wl.spec else None, wl.n_point, 0)
        # the pixels of the frame (edge tiles reach past it: the passes never write those, the pack kernel codes them as unlit)
        tiles_x = (wl.gb.width + wl.gb.tile_w - 1) // wl.gb.tile_w
        inside = np.zeros(shape, dtype=bool)
        for i in range(n):
            gt = wl.gb.tile_first + i * wl.gb.tile_stride
            ty, tx = divmod(gt, tiles_x)
            inside[i, :max(0, min(wl.gb.tile_h, wl.gb.height - ty * wl.gb.tile_h)), :max(0, min(wl.gb.tile_w, wl.gb.width - tx * wl.gb.tile_w))] = True
        inside = inside.reshape(-1)
        same = lambda a, b: np.array_equal(np.asarray(a).reshape(-1).view(np.uint32)[inside], np.asarray(b).reshape(-1).view(np.uint32)[inside])
        assert same(pk["ao"].numpy(), keep["ao"].numpy()), "e2e (packed): ao differs from the float planes"
        assert same(un["shadow"], keep["shadow"].numpy()), "e2e (packed): sun shadow differs"
        if wl.spec:
            assert same(un["spec_t"], keep["spec_t"].numpy()), "e2e (packed): spec_t differs"
        for li in range(wl.n_point):
            assert same(un["point_shadow"][li], keep["point_shadow"].numpy().reshape(wl.n_point, -1)[li]), f"e2e (packed): point plane {li} differs"
        del keep, un
    dist.barrier()
    torch.cuda.cudart().cudaHostUnregister(mine.data_ptr())
    del mine, mine_b, pk, outs, frame_host, arrivals
    try:
        shm.close()
    except BufferError:
        pass
    if rank == 0:
        shm.unlink()
    t = torch.tensor([dt, float(h2d), float(d2h), dt_packed, float(d2h_packed)], dtype=torch.float64, device=wl.ctx.torch_device)
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tsum = t.clone(); dist.all_reduce(tsum)
    dt, dt_packed = float(tmax[0].item()), float(tmax[3].item())
    return {"value": rays / dt_packed / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(tsum[1].item()), "d2h_bytes_per_step": int(tsum[4].item()), "ms_per_step": dt_packed * 1e3,
            "api": "vxl_lighting_host_packed per rank on its tile shard (pinned host planes in, one shared page-locked host frame out, a host-side barrier per frame; "
                   "decoded and compared bit for bit with the float planes, which are compared with the resident path)",
            "float_planes": {"value": rays / dt / 1e6, "unit": UNIT, "h2d_bytes_per_step": int(tsum[1].item()), "d2h_bytes_per_step": int(tsum[2].item()), "ms_per_step": dt * 1e3,
                             "api": "vxl_lighting_host per rank on its tile shard (every plane float32)"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer end-to-end leg (profiling runs)")
    ap.add_argument("--gather", choices=["auto", "fused", "nccl"], default="auto",
                    help="N > 1: output-tile exchange -- fused = peer-memory stores from the pass kernels (default on NCCL), nccl = one all-gather after the passes")
    args = ap.parse_args()
    # stdout carries exactly one JSON line: native libraries that write to fd 1 (NCCL prints its version banner there when
    # NCCL_DEBUG is set) are pointed at stderr, Python's print keeps the real stdout
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

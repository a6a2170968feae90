#!/usr/bin/env python
"""Benchmark of the plane-sweep cost-volume hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload cfg1|cfg2|...]

A *step* is one pass of the hot path over one batch of synthetic frame tuples:
one ``CostVolumeManager.forward`` (dot) / ``FeatureVolumeManager.forward`` (hero)
call.  Default workload = BASELINE.json ``configs[1]`` (dot-product model, 1 ref +
7 source views, 640x480 frames -> 120x160 matching maps, D=64, batch 4 per GPU),
the configuration the metric is quoted on.  With N GPUs every rank processes its
own batch (frames are independent: weak scaling, no data-path collective); NCCL is
used for the barrier and the MAX-reduce of the elapsed time only.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      algorithmic HBM bytes per sweep launch / measured launch time, against
                MEASURED_PEAKS.json (hbm_gbs); `traffic` = ncu dram bytes per launch
                (profiles/), null when no capture is committed
  cpu_baseline  the oracle's reference-structured port (oracle/costvolume_oracle.py,
                sampler="aten": the reference's op sequence) timed on the host cores
  e2e           same metric through the public manager API with HOST (pinned) inputs
                and outputs: H2D + sweep + D2H inside the timed region
`--impl reference` times the CPU port alone (the reference is pure Python/PyTorch and
/root/reference does not travel to the GPU box; the port is op-for-op the reference's
own sequence and is pinned bit-exact against it in tests/test_oracle_vs_reference.py).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

METRIC = "cost_volume_frames_per_sec@640x480_D64_K7src"
UNIT = "frames/s"
L2_BYTES = 126 * 1024 * 1024


def algorithmic_bytes_per_frame(w, hero: bool, with_mask: bool) -> int:
    """compulsory HBM traffic of one reference frame."""
    HW = w.height * w.width
    b = 4 * (w.channels * HW + w.views * w.channels * HW + w.planes * HW + HW)
    b += 4 * (48 * w.views + 16) + 4 * w.planes
    if hero and with_mask:
        b += HW
    return b


def binding_roof(w, hero, frames, sweep_s):
    """The roof that actually binds: for the dot sweep the 128 B/clk/SM L1 data
    path that every bilinear tap has to cross (4 taps x C x 4 B per (plane, view, pixel)); for the
    hero sweep the fp16 tensor pipe (3 MMAs per product) against the measured cuBLAS bf16 rate."""
    sm, clk = 148, 1.965e9
    samples = frames * w.planes * w.views * w.height * w.width
    if not hero:
        gather_bytes = samples * 4 * w.channels * 4
        peak = sm * 128 * clk
        return {"resource": "l1_gather_bytes", "achieved_TBps": gather_bytes / sweep_s / 1e12,
                "peak_TBps": peak / 1e12, "frac": gather_bytes / sweep_s / peak,
                "peak_source": "nominal 128 B/clk/SM x 148 SMs x 1.965 GHz"}
    f_in = w.channels * (w.views + 1) + 10 * w.views + 4
    flops = 2.0 * 3 * frames * w.planes * w.height * w.width * (f_in * 128 + 128 * 128)
    p = ROOT / "MEASURED_PEAKS.json"
    peak = float(json.loads(p.read_text())["bf16_tflops"]) * 1e12 if p.is_file() else 1.59e15
    return {"resource": "tensor_f16_flops(3 MMAs per product)", "achieved_TFLOPs": flops / sweep_s / 1e12,
            "peak_TFLOPs": peak / 1e12, "frac": flops / sweep_s / peak,
            "peak_source": "MEASURED_PEAKS.json bf16_tflops (cuBLAS burst)"}


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.is_file():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic(tag: str):
    """dram bytes per launch of the dominant kernel from a committed ncu summary."""
    p = ROOT / "profiles" / "ncu_traffic.json"
    if p.is_file():
        try:
            return json.loads(p.read_text()).get(tag)
        except Exception:
            return None
    return None


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while `active` is set."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self.active = threading.Event()
        self._stop = threading.Event()
        self._ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self._ok = True
        except Exception:
            self._ok = False
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        if not self._ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop.is_set():
            if not self.active.is_set():
                time.sleep(0.0005)
                continue
            time.sleep(0.010)   # ~100 samples/s: NVML queries take driver locks that CUDA calls of the timed
                                # loop also need, and a busy poll would fight it for the GIL
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass

    def summary(self):
        self._stop.set()
        if not self._ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                    "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


# --------------------------------------------------------------------------- #
# CPU port (reference arm / cpu_baseline)                                     #
# --------------------------------------------------------------------------- #
def cpu_port_step(w, tup, weights):
    from oracle import costvolume_oracle as O
    if w.kind == "dot":
        return O.forward_dot(**tup, num_depth_bins=w.planes, sampler="aten")
    return O.forward_mlp(**tup, weights=weights, num_depth_bins=w.planes, return_mask=True,
                         sampler="aten")


def pick_cpu_threads(w) -> int:
    """The port is ATen ops on ~20 k-pixel tensors: on a many-core host more threads can be
    SLOWER (128 threads: 0.44 frames/s, 16 threads: >5 on the same box).  Probe a few thread
    counts on a 4-plane slice of one frame and keep the fastest, so the baseline is the best
    the host can do, not an oversubscribed one."""
    import dataclasses
    from oracle import costvolume_oracle as O
    from simplerecon_b200.synthetic import make_workload_tuple, mlp_state
    cores = os.cpu_count() or 1
    cands = sorted({c for c in (4, 8, 16, 32, 64, cores) if c <= cores})
    small = dataclasses.replace(w, planes=4)
    tup = make_workload_tuple(w, batch=1)
    weights = O.mlp_weights_from_state_dict(mlp_state(w.views, w.channels)) if w.kind == "mlp" else None
    best, best_t = cands[0], float("inf")
    with torch.inference_mode():
        for c in cands:
            torch.set_num_threads(c)
            cpu_port_step(small, tup, weights)
            t0 = time.perf_counter()
            cpu_port_step(small, tup, weights)
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def time_cpu_port(w, frames: int, min_seconds: float, max_reps: int):
    """Returns (frames_per_s, cores, sample description)."""
    from oracle import costvolume_oracle as O
    from simplerecon_b200.synthetic import make_workload_tuple, mlp_state
    cores = pick_cpu_threads(w)
    tup = make_workload_tuple(w, batch=frames)
    weights = O.mlp_weights_from_state_dict(mlp_state(w.views, w.channels)) if w.kind == "mlp" else None
    with torch.inference_mode():
        cpu_port_step(w, tup, weights)                       # warm-up
        best, reps, t_all = float("inf"), 0, time.perf_counter()
        while reps < max_reps and (time.perf_counter() - t_all < min_seconds or reps < 1):
            t0 = time.perf_counter()
            cpu_port_step(w, tup, weights)
            best = min(best, time.perf_counter() - t0)
            reps += 1
    return frames / best, cores, (f"{frames} frame(s) of {w.name}, best of {reps} after 1 warm-up, "
                                  f"{cores} of {os.cpu_count()} host threads (fastest of a probe)")


def time_c_port(w, frames: int):
    """The plain-C restatement (oracle/cv_oracle.c, OpenMP over pixels, all host threads): a CPU
    number that scales with the core count where the ATen-op port does not."""
    try:
        from oracle import c_oracle as CO
        from oracle import costvolume_oracle as O
        from simplerecon_b200.synthetic import make_workload_tuple, mlp_state
        tup = make_workload_tuple(w, batch=frames)
        if w.kind == "dot":
            fn = lambda: CO.forward_dot(**tup, num_depth_bins=w.planes)
        else:
            wts = O.mlp_weights_from_state_dict(mlp_state(w.views, w.channels))
            fn = lambda: CO.forward_mlp(**tup, weights=wts, num_depth_bins=w.planes, return_mask=True)
        fn()
        best = float("inf")
        for _ in range(2):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
        return {"value": frames / best, "unit": UNIT, "threads": os.cpu_count(),
                "sample": f"{frames} frame(s), best of 2, fp32 scalar loops + OpenMP"}
    except Exception as ex:  # pragma: no cover - informational only
        return {"value": None, "error": str(ex)[:200]}


def run_reference_arm(args, w):
    """`--impl reference`: the reference's CPU implementation (port) of the same
    workload, all host threads, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import costvolume_oracle as O
    from simplerecon_b200.synthetic import make_workload_tuple, mlp_state
    cores = pick_cpu_threads(w)
    frames = 1 if w.kind == "mlp" else min(w.batch, 2)     # bounded sample of one batch
    tup = make_workload_tuple(w, batch=frames)
    weights = O.mlp_weights_from_state_dict(mlp_state(w.views, w.channels)) if w.kind == "mlp" else None
    with torch.inference_mode():
        for _ in range(max(1, min(args.warmup, 2))):
            cpu_port_step(w, tup, weights)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_port_step(w, tup, weights)
        dt = time.perf_counter() - t0
    fps = frames * args.steps / dt
    sample = (f"{frames} frame(s) per step of {w.name} (bounded sample of the batch), "
              f"{cores} of {os.cpu_count()} host threads (fastest of a probe)")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": workload_config(w, frames, 1, "n/a (CPU)"),
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(w, batch_per_gpu, world, l2_note):
    return {
        "workload": w.name, "matching": w.kind,
        "frame": "n/a (stress: feature map given directly)" if w.name.startswith("stress")
        else f"{4 * w.width}x{4 * w.height}",
        "feature_map": f"{w.width}x{w.height}", "planes": w.planes, "src_views": w.views,
        "channels": w.channels, "batch_per_gpu": batch_per_gpu, "global_batch": batch_per_gpu * world,
        "parallelism": f"frame-sharded x{world} (no data-path collective)", "l2": l2_note,
    }


# --------------------------------------------------------------------------- #
# GPU arm                                                                     #
# --------------------------------------------------------------------------- #
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg1")
    ap.add_argument("--variant", default="auto", choices=["auto", "generic", "fast"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    from simplerecon_b200.synthetic import CONFIGS, STRESS, make_workload_tuple, mlp_state
    w = next(c for c in CONFIGS + STRESS if c.name.startswith(args.workload))
    if args.impl == "reference":
        run_reference_arm(args, w)
        return

    import simplerecon_b200 as S
    from simplerecon_b200 import _native, sharding

    rank, world, local = sharding.init_distributed()
    if world != args.gpus and rank == 0:
        print(f"warning: --gpus {args.gpus} but WORLD_SIZE={world}", file=sys.stderr)
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (use --impl reference for the CPU arm)"
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    _native.check(_native.load().srcv_check_device())
    _native.set_variant({"auto": 0, "generic": 1, "fast": 2}[args.variant])

    # weak scaling: the per-GPU batch is the single-GPU configuration's batch
    per_gpu = {"cfg3": 4, "cfg4": 8}.get(args.workload[:4], w.batch)
    hero = w.kind == "mlp"
    # rotating input sets whose footprint exceeds L2, so no step finds its inputs cached
    in_bytes = 4 * per_gpu * (w.channels * w.height * w.width * (1 + w.views))
    n_sets = max(2, -(-int(1.25 * L2_BYTES) // in_bytes))
    sets_host = [make_workload_tuple(w, seed_offset=1000 * rank + i, batch=per_gpu) for i in range(n_sets)]
    sets_dev = [{k: v.to(dev) for k, v in s.items()} for s in sets_host]
    l2_note = (f"{n_sets} rotating input sets, {n_sets * in_bytes / 2**20:.0f} MiB > 126 MiB L2; "
               f"outputs ({4 * per_gpu * w.planes * w.height * w.width / 2**20:.0f} MiB/step) freshly allocated")

    import contextlib
    with contextlib.redirect_stdout(sys.stderr):   # the managers print a banner like the reference's;
        if hero:                                   # stdout carries the ONE JSON line only
            mgr = S.FeatureVolumeManager(w.height, w.width, num_depth_bins=w.planes,
                                         mlp_channels=[0, 128, 128, 1], matching_dim_size=w.channels,
                                         num_source_views=w.views)
            mgr.load_state_dict({**mgr.state_dict(), **mlp_state(w.views, w.channels)})
        else:
            mgr = S.CostVolumeManager(w.height, w.width, num_depth_bins=w.planes)
    mgr = mgr.to(dev).eval()
    kw = dict(return_mask=True) if hero else {}

    sampler = ClockSampler(local)
    with torch.inference_mode():
        def step(i):
            return mgr(**sets_dev[i % n_sets], **kw)

        for i in range(args.warmup):
            out = step(i)
        torch.cuda.synchronize()

        # ---- value: device-resident inputs ---------------------------------
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = _native.launch_count()
        sharding.barrier()
        torch.cuda.synchronize()
        sampler.active.set()
        e0.record()
        for i in range(args.steps):
            out = step(i)
        e1.record()
        torch.cuda.synchronize()
        sampler.active.clear()
        sharding.barrier()
        ms_local = e0.elapsed_time(e1)
        launches = _native.launch_count() - launches0
        ms_total = sharding.max_over_ranks(ms_local, dev)
        variant = _native.last_variant()

        # ---- roofline: the sweep kernel alone, events on its own stream -----
        _native.profile_begin(args.steps)
        sampler.active.set()
        for i in range(args.steps):
            out = step(i)
        torch.cuda.synchronize()
        sampler.active.clear()
        prep_ms, sweep_ms, nrec = _native.profile_end()

        # ---- e2e: pinned host inputs -> H2D -> sweep -> D2H, through HostStreamer ------
        from simplerecon_b200.pipeline import HostStreamer
        pin = [{k: v.pin_memory() for k, v in s_.items()} for s_ in sets_host[:3]]
        h2d = sum(v.numel() * v.element_size() for v in pin[0].values())
        streamer = HostStreamer(mgr, dev, return_mask=hero)
        outs = list(streamer.run(pin[i % len(pin)] for i in range(4)))   # warm-up
        d2h = sum(t.numel() * t.element_size() for t in outs[-1])
        e2e_steps = max(5, min(args.steps, 50))
        sharding.barrier()
        torch.cuda.synchronize()
        sampler.active.set()
        t_wall0 = time.perf_counter()
        e0.record()
        n_out = 0
        for host_res in streamer.run(pin[i % len(pin)] for i in range(e2e_steps)):
            n_out += 1
        e1.record()
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t_wall0
        sampler.active.clear()
        sharding.barrier()
        assert n_out == e2e_steps
        # device time between the first H2D and the last D2H (events on the default stream
        # bracket the streamer's three streams through the synchronising yields) vs wall clock:
        # report the larger so host-side stalls are not hidden
        e2e_ms = sharding.max_over_ranks(max(e0.elapsed_time(e1), 1e3 * t_wall), dev)

    clocks = sampler.summary()
    frames_total = per_gpu * world * args.steps
    value = frames_total / (ms_total * 1e-3)
    e2e_value = per_gpu * world * e2e_steps / (e2e_ms * 1e-3)

    peak, peak_src = load_peaks()
    alg_bytes = algorithmic_bytes_per_frame(w, hero, True) * per_gpu
    sweep_avg_s = sweep_ms * 1e-3 / max(nrec, 1)
    achieved = alg_bytes / sweep_avg_s / 1e9
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": load_traffic(f"{w.kind}:{variant}"),
        "kernel": variant, "algorithmic_bytes_per_launch": alg_bytes,
        "sweep_us_per_launch": sweep_avg_s * 1e6, "prep_us_per_launch": prep_ms * 1e3 / max(nrec, 1),
        "sweep_share_of_step": (sweep_ms / max(nrec, 1)) / (ms_local / args.steps),
        "peak_source": peak_src,
        "binding": binding_roof(w, hero, per_gpu, sweep_avg_s),
        "note": ("HBM fraction as the metric demands; the sweep is bound on chip (L1 gather "
                 "bandwidth for dot, tensor + SIMT issue for hero) — see `binding` and DESIGN.md"),
    }

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(w, per_gpu, world, l2_note),
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                "mode": "HostStreamer: H2D(i+1) || sweep(i) || D2H(i-1) on three streams, pinned host buffers"},
        "gpu_launches": launches,
        "clocks": clocks,
        "kernel_variant": variant,
    }
    gpu_port = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # part of the baseline leg (SURVEY.md §8d "second comparator"): the same port — the
        # reference's per-plane grid_sample / cat / Linear sequence — run on THIS GPU, i.e. what a
        # user of the reference gets today on the same hardware from library kernels.
        try:
            from oracle import costvolume_oracle as O
            frames_g = min(per_gpu, 2 if hero else 4)
            tg = {k: v[:frames_g] if (torch.is_tensor(v) and v.dim() > 0 and v.shape[0] == per_gpu) else v
                  for k, v in sets_dev[0].items()}
            wg = tuple(x.to(dev) for x in O.mlp_weights_from_state_dict(mlp_state(w.views, w.channels))) if hero else None
            with torch.inference_mode():
                cpu_port_step(w, tg, wg)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(2):
                    cpu_port_step(w, tg, wg)
                e1.record()
                torch.cuda.synchronize()
            gpu_port = {
                "value": 2 * frames_g / (e0.elapsed_time(e1) * 1e-3), "unit": UNIT,
                "sample": f"{frames_g} frame(s), 2 reps, the port's op sequence as PyTorch CUDA ops on cuda:{local}"}
        except Exception as ex:  # pragma: no cover - informational only
            gpu_port = {"value": None, "error": str(ex)[:200]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        frames = 1 if hero else 2
        fps, cores, sample = time_cpu_port(w, frames, min_seconds=10.0, max_reps=5)
        line["cpu_baseline"] = {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                                "port_on_this_gpu": gpu_port, "c_port_openmp": time_c_port(w, frames)}
    elif rank == 0:
        line["cpu_baseline"] = None
    if rank == 0:
        print(json.dumps(line), flush=True)
    if torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py - CT volumes/sec of the BOA hot path (`--models total+bca`) on synthetic 512x512x512 @1.5 mm volumes.

    python bench.py --gpus N --steps K --warmup W            # this framework (libboa_b200, one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference algorithm's CPU path (oracle port)

One "step" = the whole per-volume hot path on one synthetic CT: 5 `total` part networks (125 patches each, step 0.8),
the two body-composition networks at 5 mm slice thickness (98 patches x 5 folds each), Gaussian aggregation, argmax +
label merge, tissue rules and all measurement reductions.  `value` times it with the CT resident in HBM; `e2e` times
the public API call from pinned host memory (H2D of the CT, D2H of the four label maps, JSON tables).
Prints ONE JSON line (rank 0).
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
sys.path.insert(0, ROOT)

METRIC = "CT volumes/sec (512x512x512, 1.5mm) total+bca"
UNIT = "volumes/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", type=int, nargs=3, default=[512, 512, 512])
    ap.add_argument("--patch", type=int, default=128)
    ap.add_argument("--batch", type=int, default=int(os.environ.get("BOA_B200_BATCH", "8")))
    ap.add_argument("--fast-bca", action="store_true", help="fold 0 only for the body-composition nets (--fast-bca)")
    ap.add_argument("--models", default="total+bca")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a) -> str:
    return (f"synthetic CT {a.shape[0]}x{a.shape[1]}x{a.shape[2]} @1.5mm, --models {a.models}"
            f"{' --fast-bca' if a.fast_bca else ''}, patch {a.patch}^3")


def count_forwards(a) -> dict:
    from boa_b200.geometry import sliding_window_origins
    from boa_b200.resample import resampled_depth

    P = (a.patch,) * 3
    shape = [max(s, a.patch) for s in a.shape]
    n_total = len(sliding_window_origins(shape, P, 0.8)) * 5 if "total" in a.models or "bca" in a.models else 0
    z5 = max(resampled_depth(a.shape[0], 1.5, 5.0), a.patch)
    folds = 1 if a.fast_bca else 5
    n_bca = len(sliding_window_origins([z5, shape[1], shape[2]], P, 0.5)) * folds * 2 if "bca" in a.models else 0
    return {"total": n_total, "bca": n_bca}


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                                  ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_reference_sample(a, n_patches: int = 2) -> dict:
    """The reference algorithm (`-d cpu` path) restated in oracle/: torch fp32 PlainConvUNet forward + numpy passes,
    timed on a bounded sample and scaled linearly to the whole workload."""
    import numpy as np
    import torch

    from boa_b200 import zoo
    from boa_b200.plans import arch_from_plans
    from oracle import passes as op
    from oracle.network import unet_forward

    ncpu = os.cpu_count() or 1
    threads = min(8, ncpu)  # predict_from_raw_data.py:479-480 caps torch threads at default_num_processes = 8
    torch.set_num_threads(threads)
    arch = arch_from_plans(zoo.default_plans((a.patch,) * 3, 32, 320, 6), "3d_fullres", 1, 25)
    sd = zoo.random_state_dict(arch, 1)
    x = torch.randn(1, 1, a.patch, a.patch, a.patch)
    unet_forward(arch, sd, x[:, :, :64, :64, :64].contiguous())  # warm the thread pool / allocator
    t0 = time.perf_counter()
    for _ in range(n_patches):
        unet_forward(arch, sd, x)
    t_patch = (time.perf_counter() - t0) / n_patches
    # memory-bound passes on a slab of 16 slices, scaled to the volume (numpy, one core, as the reference runs them)
    zs = 16
    scale = a.shape[0] / zs
    rng = np.random.default_rng(0)
    logits = rng.standard_normal((25, zs, a.shape[1], a.shape[2]), dtype=np.float32)
    ct = rng.integers(-1024, 2047, size=(zs, a.shape[1], a.shape[2])).astype(np.int16)
    t0 = time.perf_counter()
    seg = logits.argmax(0).astype(np.uint8)                      # export_prediction.py:38, once per network
    t_argmax = (time.perf_counter() - t0) * scale
    regions = (seg % 12).astype(np.uint8)
    t0 = time.perf_counter()
    tissues = op.subclassify_tissues(ct, regions)                # subclassification.py:38-53
    op.slice_label_stats(tissues, 8, ct)                         # builder.py:403-444
    t_bca = (time.perf_counter() - t0) * scale
    t0 = time.perf_counter()
    n_lab = 8
    for lab in range(1, 1 + n_lab):                              # measurements.py:203-241, 304 names per volume
        op.metrics_for_region(ct, seg == lab, 30.0, 10.0, (1.5, 1.5, 1.5))
    t_label = (time.perf_counter() - t0) / n_lab * scale
    n = count_forwards(a)
    n_nets = (5 if n["total"] else 0) + (2 if n["bca"] else 0)
    t_volume = (n["total"] + n["bca"]) * t_patch + n_nets * t_argmax + (t_bca if n["bca"] else 0) + 304 * t_label
    return {"value": 1.0 / t_volume, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": (f"{n_patches} patch forwards of the fp32 oracle network ({t_patch:.2f} s each) scaled to "
                       f"{n['total'] + n['bca']} forwards; numpy argmax ({t_argmax:.1f} s/net x {n_nets}), tissue + slice "
                       f"tables ({t_bca:.1f} s) and per-label statistics ({t_label:.2f} s x 304 names) timed on a "
                       f"{zs}-slice slab and scaled to the volume; {ncpu} host cores visible, {threads} torch threads "
                       f"(the reference's cap)"),
            "seconds_per_volume": t_volume}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, ms = [], []
    last = None
    for i in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        last = cpu_reference_sample(a, n_patches=1)
        if i >= a.warmup:
            vals.append(last["value"])
            ms.append((time.perf_counter() - t0) * 1e3)
    v = sum(vals) / len(vals)
    last["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": sum(ms) / len(ms), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(a)},
        "cpu_baseline": last, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ ours
def run_ours(a):
    import numpy as np
    import torch

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (boa_b200 has no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist_ctx = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        from boa_b200.pipeline import DistContext
        dist_ctx = DistContext(rank=rank, world_size=world, group=None)

    from boa_b200 import _lib, zoo
    from boa_b200.pipeline import ModelZoo, analyze_from_host, analyze_volume

    models = tuple(a.models.split("+"))
    datasets = [291, 292, 293, 294, 295] + ([542, 543] if "bca" in models else [])
    specs = zoo.synthetic_specs((a.patch,) * 3, 32, 320, 6, bca_folds=1 if a.fast_bca else 5, datasets=datasets)
    mz = ModelZoo.from_specs(specs, device=dev, max_batch=a.batch)
    ct_np = zoo.synthetic_ct(tuple(a.shape), seed=3)
    ct_host = torch.from_numpy(ct_np).pin_memory()
    ct_dev = ct_host.to(dev)
    spacing = (1.5, 1.5, 1.5)
    kw = dict(models=models, fast_bca=a.fast_bca, dist_ctx=dist_ctx)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    res = None
    for _ in range(a.warmup):
        res = analyze_volume(ct_dev, spacing, mz, **kw)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    launches0 = _lib.lib().boa_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        res = analyze_volume(ct_dev, spacing, mz, **kw)
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = int(_lib.lib().boa_kernel_launch_count() - launches0)
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_step = float(ms.item()) / a.steps

    # end to end through the public API, host buffers (one untimed call first: the pinned staging buffers of the
    # label maps are allocated on first use)
    out = analyze_from_host(ct_host, spacing, mz, device=dev, **kw)
    barrier()
    t0 = time.perf_counter()
    e2e_each = []
    for _ in range(a.steps):
        t1 = time.perf_counter()
        out = analyze_from_host(ct_host, spacing, mz, device=dev, **kw)
        e2e_each.append(time.perf_counter() - t1)
    barrier()
    e2e_s = torch.tensor([(time.perf_counter() - t0) / a.steps], device=dev, dtype=torch.float64)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    d2h = sum(int(out[k].numel()) for k in ("total", "body_parts", "body_regions", "tissues", "ct_pfav") if out.get(k) is not None)
    d2h += len(json.dumps(out["total_measurements"])) + len(json.dumps(out["bca_measurements"] or {}))

    # roofline of the dominant kernel (dz-folded tcgen05 conv), measured live with CUDA events on the launch stream
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    net = mz.get(291, [0], 0.8).networks[0]
    desc = net.describe()
    best = None
    for _ in range(3):
        t = net.time_layers()
        best = t if best is None else [min(x, y) for x, y in zip(best, t)]
    kinds = {0: "conv3_fold_kernel(tcgen05)", 1: "conv_taps_kernel(tcgen05,stride2)", 2: "conv_simt",
             3: "conv_taps_kernel(tcgen05,transposed)", 4: "tconv_simt", 5: "conv_first_kernel(fp32 simt)"}
    per_kind = {}
    for (name, kind, macs), t_ms in zip(desc, best):
        k = per_kind.setdefault(kinds[kind], {"launches": 0, "ms": 0.0, "flop": 0.0})
        k["launches"] += 1; k["ms"] += t_ms; k["flop"] += 2.0 * macs * a.batch
    for k in per_kind.values():
        k["tflops"] = k["flop"] / (k["ms"] * 1e-3) / 1e12 if k["ms"] > 0 else None
    dom = per_kind[kinds[0]]
    peak = peaks.get("bf16_tflops_sustained") or 1400.0
    traffic = None
    try:  # dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture
        with open(os.path.join(ROOT, "profiles", "r01_fold_traffic.json")) as f:
            traffic = float(json.load(f)["bytes_per_launch"])
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": kinds[0], "achieved": dom["tflops"], "peak": peak, "unit": "TFLOP/s",
                "frac": dom["tflops"] / peak, "traffic": traffic,
                "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step), of measured"
                                if peaks else "fallback 1.4 PFLOP/s sustained, of fallback"),
                "per_launch": {"avg_ms": dom["ms"] / dom["launches"], "avg_flop": dom["flop"] / dom["launches"],
                               "patches_per_launch": a.batch},
                "conv_time_share_of_forward": None, "kernels": per_kind}
    # share of the step the dominant kernel accounts for (event-timed launches x batches per volume / step time), to
    # be compared with its share in the ncu launch list of the same command (profiles/rNN_launches_summary.txt)
    n_fwd = count_forwards(a)
    batches = (n_fwd["total"] + n_fwd["bca"]) / float(a.batch) / max(world, 1)
    roofline["share_of_step"] = dom["ms"] * batches / ms_step

    if rank == 0:
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            cpu = cpu_reference_sample(a, n_patches=3)
        n = count_forwards(a)
        flop_per_volume = 2.0 * net.macs_per_patch * (n["total"] + n["bca"])
        line = {
            "metric": METRIC, "value": 1e3 / ms_step, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": workload_name(a), "forwards_per_volume": n, "patch_batch": a.batch,
                       "tflop_per_volume": flop_per_volume / 1e12,
                       "achieved_tflops_whole_step": flop_per_volume / (ms_step * 1e-3) / 1e12,
                       "l2": "inputs larger than L2 (268 MB CT, >1 GB activations per layer batch)",
                       "parallelism": f"patches sharded over {world} GPU(s), NCCL slab exchange" if world > 1 else "1 GPU"},
            "e2e": {"value": 1.0 / float(e2e_s.item()), "unit": UNIT, "h2d_bytes_per_step": int(ct_host.numel() * 2),
                    "d2h_bytes_per_step": int(d2h), "seconds_each": [round(x, 4) for x in e2e_each]},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "stage_seconds": res.timings if res is not None else None,
        }
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)

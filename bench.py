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
    ap.add_argument("--mode", default="latency", choices=["latency", "throughput"],
                    help="latency (default): the patches of ONE volume are sharded over the GPUs (BASELINE config 3); "
                         "throughput: every GPU processes its own volumes, no collectives (BASELINE config 4)")
    ap.add_argument("--quick", action="store_true", help="skip the roofline / baseline legs (extra configurations)")
    return ap.parse_args()


def workload_name(a) -> str:
    return (f"synthetic CT {a.shape[0]}x{a.shape[1]}x{a.shape[2]} @1.5mm, --models {a.models}"
            f"{' --fast-bca' if a.fast_bca else ''}, patch {a.patch}^3"
            f"{', one volume per GPU (throughput mode)' if getattr(a, 'mode', 'latency') == 'throughput' else ''}")


def count_forwards(a) -> dict:
    from boa_b200.geometry import sliding_window_origins
    from boa_b200.resample import resampled_depth

    from boa_b200.config import resolve_models

    ms = resolve_models(a.models, strict=True)
    P = (a.patch,) * 3
    shape = [max(s, a.patch) for s in a.shape]
    n_total = len(sliding_window_origins(shape, P, 0.8)) * 5 if "total" in ms else 0
    z5 = max(resampled_depth(a.shape[0], 1.5, 5.0), a.patch)
    folds = 1 if a.fast_bca else 5
    n_bca = len(sliding_window_origins([z5, shape[1], shape[2]], P, 0.5)) * folds * 2 if "bca" in ms else 0
    # (the crop-pre-pass tasks of `--models all` add one 6 mm forward set and one cropped forward set each; their patch
    # count depends on the rough segmentation, so they are not part of the FLOP accounting)
    return {"total": n_total, "bca": n_bca}


def bench_config(a, world: int) -> dict:
    """`config` of the JSON line - the same dict for both arms (`--impl ours` / `--impl reference`), so that the
    driver can tell they measured the same workload."""
    from boa_b200 import zoo
    from boa_b200.plans import arch_from_plans, macs_per_patch

    n = count_forwards(a)
    arch = arch_from_plans(zoo.default_plans((a.patch,) * 3, 32, 320, 6), "3d_fullres", 1, 25)
    return {"workload": workload_name(a), "forwards_per_volume": n, "patch_batch": a.batch,
            "tflop_per_volume": 2.0 * macs_per_patch(arch) * (n["total"] + n["bca"]) / 1e12,
            "l2": "inputs larger than L2 (268 MB CT, >1 GB activations per layer batch)",
            "parallelism": ("1 GPU" if world == 1 else
                            f"{world} replicas, one volume per GPU, no collectives" if getattr(a, "mode", "") == "throughput"
                            else f"patches of one volume sharded over {world} GPUs, one peer-memory slab reduction "
                                 "per network")}


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
            "seconds_per_volume": t_volume,
            "parts": {"t_patch": t_patch, "t_argmax": t_argmax, "t_bca": t_bca, "t_label": t_label}}


# ------------------------------------------------------------------------------------------------ GPU baseline
def gpu_reference_sample(a, dev, cpu: dict | None, n_patches: int = 24) -> dict | None:
    """The "reference's single-GPU PyTorch/nnU-Net path" of BASELINE.json, restated on stock PyTorch (baseline/): the
    torch.nn PlainConvUNet under torch.autocast(fp16) with cudnn.benchmark and the reference's loop (producer thread,
    fp16 accumulators, three elementwise launches per patch, final divide + isinf), then what the reference does with
    the logits of every network: D2H of the C x V tensor, numpy argmax, masked-write part merge, numpy measurements
    (predict_from_raw_data.py:386,560-631,648; export_prediction.py:38; totalsegmentator/nnunet.py:553-556;
    compute/measurements.py).  Timed on a bounded sample on this box and scaled linearly, like the CPU baseline."""
    import numpy as np
    import torch

    from baseline import reference_loop, torch_unet
    from boa_b200 import zoo
    from boa_b200.plans import arch_from_plans
    from oracle import passes as op

    try:
        arch = arch_from_plans(zoo.default_plans((a.patch,) * 3, 32, 320, 6), "3d_fullres", 1, 25)
        sd = zoo.random_state_dict(arch, 1)
        torch.backends.cudnn.benchmark = True
        net = torch_unet.build(arch, sd, dev)
        shape = [max(s, a.patch) for s in a.shape]
        data = torch.randn((1, *shape))
        step = 0.8
        # warm-up: cuDNN autotuning of every layer shape happens on the first patches
        reference_loop.predict_sliding_window_return_logits(net, data, arch["patch_size"], step, dev, patch_range=(0, 3))
        torch.cuda.synchronize()
        t_loop = None
        for _ in range(2):  # the faster of two runs: the baseline gets the benefit of the doubt
            logits = None
            t0 = time.perf_counter()
            logits = reference_loop.predict_sliding_window_return_logits(net, data, arch["patch_size"], step, dev,
                                                                         patch_range=(0, n_patches))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t_loop = dt if t_loop is None else min(t_loop, dt)
        # the loop's fixed part (allocation of the accumulators, the divide and the isinf scan over the whole volume)
        # is inside t_loop once; separate it with a second, patch-free call
        t_one = None
        for _ in range(2):
            t0 = time.perf_counter()
            reference_loop.predict_sliding_window_return_logits(net, data, arch["patch_size"], step, dev, patch_range=(0, 1))
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            t_one = dt if t_one is None else min(t_one, dt)
        t_patch = (t_loop - t_one) / (n_patches - 1)
        t_fixed = max(t_one - t_patch, 0.0)
        t0 = time.perf_counter()
        host = logits.to("cpu")                                   # predict_from_raw_data.py:386 / :497
        t_d2h = time.perf_counter() - t0
        del logits
        zs = 32
        t0 = time.perf_counter()
        seg = host[:, :zs].numpy().argmax(0).astype(np.uint8)     # label_handling.py:178 on a slab, scaled
        t_argmax = (time.perf_counter() - t0) * shape[0] / zs
        combined = np.zeros_like(seg)
        t0 = time.perf_counter()
        reference_loop.merge_part(combined, seg, list(range(25)))  # 24 masked writes, scaled
        t_merge = (time.perf_counter() - t0) * shape[0] / zs
        del host
        n = count_forwards(a)
        n_total_nets = 5 if n["total"] else 0
        n_bca_fold_runs = (2 * (1 if a.fast_bca else 5)) if n["bca"] else 0
        vox5 = 154.0 / 512.0  # the body-composition nets run on the 5 mm volume: D2H / argmax scale with its size
        bytes_frac_bca = vox5 * (12 + 7) / 2.0 / 25.0
        t_label = t_bca = 0.0
        if cpu is not None:  # per-label statistics and slice tables: same numpy code as the CPU arm
            t_cpu = cpu.get("parts", {})
            t_label, t_bca = t_cpu.get("t_label", 0.0), t_cpu.get("t_bca", 0.0)
        t_volume = ((n["total"] + n["bca"]) * t_patch
                    + n_total_nets * (t_fixed + t_d2h + t_argmax + t_merge)
                    + n_bca_fold_runs * (t_fixed * vox5 + t_d2h * bytes_frac_bca)
                    + (2 if n["bca"] else 0) * t_argmax * bytes_frac_bca
                    + (t_bca if n["bca"] else 0.0) + 304 * t_label)
        t_gpu_only = (n["total"] + n["bca"]) * t_patch + (n_total_nets + n_bca_fold_runs * vox5) * t_fixed
        return {"value": 1.0 / t_volume, "unit": UNIT, "kind": "restated reference GPU path (torch " + torch.__version__
                + f", cuDNN {torch.backends.cudnn.version()}, autocast fp16, cudnn.benchmark)",
                "seconds_per_volume": t_volume, "seconds_per_volume_network_loop_only": t_gpu_only,
                "value_network_loop_only": 1.0 / t_gpu_only,
                "parts": {"t_patch": t_patch, "t_fixed_per_net": t_fixed, "t_d2h_logits": t_d2h,
                          "t_numpy_argmax": t_argmax, "t_part_merge": t_merge, "t_label": t_label, "t_bca": t_bca},
                "sample": (f"{n_patches} patches of {a.patch}^3 through the reference loop after cuDNN autotuning "
                           f"({t_patch * 1e3:.1f} ms per patch incl. its 3 elementwise launches) scaled to "
                           f"{n['total'] + n['bca']} forwards; accumulator allocation + divide + isinf "
                           f"({t_fixed:.2f} s) and D2H of the 25 x V fp16 logits ({t_d2h:.2f} s) measured once on the "
                           f"full volume and counted per network (per fold for the body-composition nets, scaled to "
                           f"their 5 mm volume); numpy argmax ({t_argmax:.1f} s) and the 24 masked-write part merge "
                           f"({t_merge:.1f} s) timed on a {zs}-slice slab and scaled; per-label statistics and tissue / "
                           f"slice tables taken from the CPU arm (same numpy code)")}
    except Exception as e:  # the baseline is a reported number, never a reason to lose the bench line
        return {"unavailable": f"{type(e).__name__}: {e}"}


def hbm_pass_roofline(ct_dev, res, peaks: dict, head_entry: dict | None) -> dict:
    """Achieved HBM GB/s of the memory-bound passes on the benchmark's own volume: algorithmic bytes
    over the CUDA-event time of the launch, average of 5 after 2 warm-ups; every input is larger than L2."""
    import torch

    from boa_b200 import passes
    from boa_b200.predictor import finalize_argmax

    peak = peaks.get("hbm_gbs") or 6500.0
    V = ct_dev.numel()
    out = {}

    def timed(name, nbytes, fn, reps=5):
        for _ in range(2):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[name] = {"bytes": nbytes, "ms": ms, "gbs": nbytes / ms / 1e6, "frac": nbytes / ms / 1e6 / peak}

    C = 25
    acc = torch.randn((C, *ct_dev.shape), device=ct_dev.device)
    w = torch.rand(ct_dev.shape, device=ct_dev.device) + 0.5
    lab = torch.zeros(ct_dev.shape, dtype=torch.uint8, device=ct_dev.device)
    lut = list(range(C))
    timed("finalize_argmax_kernel", (4 * C + 4 + 2) * V, lambda: finalize_argmax(acc, w, lut, lab, True))
    del acc, w
    regions = res.body_regions if res.body_regions is not None else lab
    timed("tissue_kernel", (2 + 1 + 1) * V, lambda: passes.tissue_subclassify(ct_dev, regions))
    tissues = res.tissues if res.tissues is not None else lab
    timed("slice_stats_kernel", (1 + 2) * V, lambda: passes.slice_label_stats(tissues, 8, ct=ct_dev))
    total = res.total if res.total is not None else lab
    timed("label_hist_kernel", (2 + 1) * V, lambda: passes.label_hu_hist(ct_dev, total, 118, -32768, 65536))
    timed("erode_axis_kernel(x3)", 6 * V, lambda: passes.erode_box(lab))
    timed("ct_normalize_kernel", (2 + 4) * V,
          lambda: passes.ct_normalize(ct_dev, -1024.0, 276.0, -370.0, 436.6))
    if head_entry and head_entry.get("gbs"):
        out["head_mma_kernel(in step)"] = {"bytes": head_entry["bytes"] / head_entry["launches"],
                                           "ms": head_entry["ms"] / head_entry["launches"],
                                           "gbs": head_entry["gbs"], "frac": head_entry["gbs"] / peak}
    return {"bound": "hbm", "peak": peak, "unit": "GB/s",
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (copy, read+write), of measured" if peaks else "fallback",
            "timing": "CUDA events, 5 launches after 2 warm-ups, 512^3 inputs (> L2); head: events around every launch inside a whole-volume step",
            "kernels": out}


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
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": bench_config(a, max(a.gpus, 1)),
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
    throughput = a.mode == "throughput"
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        from boa_b200.pipeline import DistContext
        # throughput mode: replicas only - NCCL is used for the timing barrier and the max over ranks, never on the
        # data path
        dist_ctx = None if throughput else DistContext(rank=rank, world_size=world, group=None)

    from boa_b200 import _lib, zoo
    from boa_b200.pipeline import ModelZoo, analyze_from_host, analyze_volume

    from boa_b200.config import resolve_models
    from boa_b200.labels import CROP_PREPASS_TASK_ID, CROP_TASKS
    models = tuple(sorted(resolve_models(a.models, strict=True)))  # "all" = every model that needs no licence
    crop = [m for m in models if m in CROP_TASKS]
    datasets = [291, 292, 293, 294, 295] + ([542, 543] if "bca" in models else []) + \
               ([CROP_PREPASS_TASK_ID] + [CROP_TASKS[m][0] for m in crop] if crop else [])
    specs = zoo.synthetic_specs((a.patch,) * 3, 32, 320, 6, bca_folds=1 if a.fast_bca else 5, datasets=datasets)
    mz = ModelZoo.from_specs(specs, device=dev, max_batch=a.batch)
    ct_np = zoo.synthetic_ct(tuple(a.shape), seed=3 + (rank if throughput else 0))
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
    if throughput:
        ms_step /= world  # `world` volumes finish per round of steps

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
    if throughput:
        e2e_s /= world
    d2h = sum(int(out[k].numel()) for k in ("total", "body_parts", "body_regions", "tissues", "ct_pfav") if out.get(k) is not None)
    d2h += len(json.dumps(out["total_measurements"])) + len(json.dumps(out["bca_measurements"] or {}))

    # ---- label-map checksums; at N > 1 rank 0 also runs the SAME volume on one GPU (untimed) and compares
    import zlib
    maps = {k: getattr(res, k) for k in ("total", "body_parts", "body_regions", "tissues") if getattr(res, k) is not None}
    checksum = {k: f"{zlib.crc32(v.cpu().numpy().tobytes()):08x}" for k, v in maps.items()} if rank == 0 else None
    vs_single = None
    if world > 1 and rank == 0 and not throughput:
        single = analyze_volume(ct_dev, spacing, mz, **dict(kw, dist_ctx=None))
        vs_single = {}
        for k, v in maps.items():
            ref_map = getattr(single, k)
            vs_single[k] = {"agreement": float((v == ref_map).float().mean().item()),
                            "mismatched_voxels": int((v != ref_map).sum().item())}
        vs_single["checksum_single_gpu"] = {k: f"{zlib.crc32(getattr(single, k).cpu().numpy().tobytes()):08x}"
                                            for k in maps}
        del single
    barrier()
    if a.quick:
        if rank == 0:
            n = count_forwards(a)
            net = mz.get(291, [0], 0.8).networks[0]
            flop_per_volume = 2.0 * net.macs_per_patch * (n["total"] + n["bca"])
            print(json.dumps({
                "metric": METRIC, "value": 1e3 / ms_step, "unit": UNIT, "n_gpus": world, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "weak" if throughput else "strong", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
                "config": bench_config(a, world), "achieved_tflops_whole_step": flop_per_volume / (ms_step * 1e-3) / 1e12,
                "e2e": {"value": 1.0 / float(e2e_s.item()), "unit": UNIT, "h2d_bytes_per_step": int(ct_host.numel() * 2),
                        "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": launches, "clocks": clocks, "label_checksum": checksum, "vs_single_gpu": vs_single,
                "stage_seconds": res.timings if res is not None else None}))
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (dz-folded tcgen05 conv): CUDA events around EVERY launch of every conv
    # kernel inside one more whole-volume step, run right after the timed region (same clocks / power state as the
    # step; the timing schedule is single-lane so the event brackets are exclusive) -> divide by the SUSTAINED peak.
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    nets = [n for p in mz._cache.values() for n in p.networks]
    for n in nets:
        n.enable_timing(True)
    analyze_volume(ct_dev, spacing, mz, **kw)
    kinds = {0: "conv3_fold_kernel(tcgen05)", 1: "conv_taps_kernel(tcgen05,stride2)", 2: "conv_simt",
             3: "conv_taps_kernel(tcgen05,transposed)", 4: "tconv_simt", 5: "conv_first_kernel(fp32 simt)",
             6: "head_mma_kernel(+gaussian accumulate)"}
    per_kind = {}
    for n in nets:
        for k, (t_ms, work, n_l) in enumerate(n.read_timing_kinds(reset=True)):
            if n_l:
                d = per_kind.setdefault(kinds[k], {"launches": 0, "ms": 0.0, "work": 0.0})
                d["launches"] += n_l; d["ms"] += t_ms; d["work"] += work
        n.enable_timing(False)
    for name, d in per_kind.items():
        rate = d["work"] / (d["ms"] * 1e-3) if d["ms"] > 0 else None
        if name.startswith("head"):
            d["bytes"] = d.pop("work"); d["gbs"] = rate / 1e9 if rate else None
        else:
            d["flop"] = d.pop("work"); d["tflops"] = rate / 1e12 if rate else None
    dom = per_kind[kinds[0]]
    sustained = peaks.get("bf16_tflops_sustained") or 1400.0
    burst = peaks.get("bf16_tflops") or 1650.0
    # the same kernels timed alone (one body of lane 0 after the step, best of 3): a BURST figure -> burst peak
    net = mz.get(291, [0], 0.8).networks[0]
    desc = net.describe()
    best = None
    for _ in range(3):
        t = net.time_layers()
        best = t if best is None else [min(x, y) for x, y in zip(best, t)]
    iso_ms = sum(t for (_, kind, _), t in zip(desc, best) if kind == 0)
    iso_flop = sum(2.0 * macs * a.batch for (_, kind, macs) in desc if kind == 0)
    traffic = None
    try:  # dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture
        with open(os.path.join(ROOT, "profiles", "r02_fold_traffic.json")) as f:
            traffic = float(json.load(f)["bytes_per_launch"])
    except Exception:
        pass
    n_fwd = count_forwards(a)
    roofline = {"bound": "tensor", "kernel": kinds[0], "achieved": dom["tflops"], "peak": sustained, "unit": "TFLOP/s",
                "frac": dom["tflops"] / sustained, "traffic": traffic,
                "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained, of measured" if peaks
                                else "fallback 1.4 PFLOP/s sustained, of fallback"),
                "timing": ("CUDA events around every launch of the kernel inside one whole-volume step run directly "
                           "after the timed region (single-lane schedule: exclusive brackets), all networks"),
                "per_launch": {"avg_ms": dom["ms"] / dom["launches"], "avg_flop": dom["flop"] / dom["launches"],
                               "launches_per_volume": dom["launches"], "patches_per_launch": a.batch},
                # share of the step: in-step kernel time / step time, to be compared with the kernel's share in the
                # ncu launch list of the same command (profiles/rNN_launches_summary.txt)
                "share_of_step": dom["ms"] / ms_step,
                "isolated_burst": {"achieved": iso_flop / (iso_ms * 1e-3) / 1e12, "peak": burst,
                                   "frac": iso_flop / (iso_ms * 1e-3) / 1e12 / burst,
                                   "timing": "one network body alone after the step, best of 3 (burst peak)"},
                "kernels": per_kind}
    roofline_hbm = hbm_pass_roofline(ct_dev, res, peaks, per_kind.get(kinds[6])) if rank == 0 else None

    if rank == 0:
        cpu = gpu_base = None
        if world == 1 and not a.no_cpu_baseline:
            cpu = cpu_reference_sample(a, n_patches=3)
            gpu_base = gpu_reference_sample(a, dev, cpu)
        n = count_forwards(a)
        flop_per_volume = 2.0 * net.macs_per_patch * (n["total"] + n["bca"])
        line = {
            "metric": METRIC, "value": 1e3 / ms_step, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak" if throughput else "strong",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": bench_config(a, world),
            "achieved_tflops_whole_step": flop_per_volume / (ms_step * 1e-3) / 1e12,
            "e2e": {"value": 1.0 / float(e2e_s.item()), "unit": UNIT, "h2d_bytes_per_step": int(ct_host.numel() * 2),
                    "d2h_bytes_per_step": int(d2h), "seconds_each": [round(x, 4) for x in e2e_each]},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_hbm": roofline_hbm,
            "cpu_baseline": cpu, "gpu_baseline": gpu_base,
            "label_checksum": checksum, "vs_single_gpu": vs_single,
            "stage_seconds": res.timings if res is not None else None,
        }
        print(json.dumps(line))
    barrier()  # the other ranks wait for rank 0's single-GPU legs before the group goes away
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    if os.environ.get("BOA_BENCH_WATCHDOG"):  # dump every thread's stack and exit if the run takes longer than this
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["BOA_BENCH_WATCHDOG"]), exit=True)
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)

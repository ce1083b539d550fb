#!/usr/bin/env python
"""bench.py — images/sec of one training iteration of TestModel (ResNet-50 encoder, 320×320, bs 16/GPU, bf16)
through the B200 hot path, with the CPU reference timed beside it.

    python bench.py [--gpus N --steps K --warmup W]                      # this repo's arm
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference [--steps K --warmup W]              # the reference's algorithm on host cores
    python bench.py --sweep                                              # BASELINE config 5 (needs ≥2 ranks)

One JSON line on stdout (rank 0).  `value` = whole-job images/s with the batch already resident in HBM;
`e2e` = the same through `Trainer.step_from_host` (pinned host batch → H2D → iteration → loss D2H);
`roofline` = the dominant hand-written kernel (SyncBN backward) replayed on the model's 84 layer shapes and
timed with CUDA events on the launching stream; `cpu_baseline` = the oracle restatement of the reference loop
on the host cores (bounded sample).  Nothing here reads /root/reference.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

SIZE, BS = 320, 16
METRIC, UNIT = "images_per_sec", "img/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="res50", choices=["res50", "cp_res50"])
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=4)
    ap.add_argument("--sweep", action="store_true")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference loop; test infrastructure used as the reported baseline)
# ----------------------------------------------------------------------------------------------------
def cpu_reference(model_name: str, batch: int, steps: int, warmup: int):
    from oracle.step import OracleTrainer
    from distributed_sod_project_b200 import network
    from distributed_sod_project_b200.synthetic import synth_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    tr = OracleTrainer(getattr(network, model_name), world_size=1, seed=0)
    batches = [synth_batch(1234 + i, batch, SIZE) for i in range(2)]
    for i in range(warmup):
        tr.step(*batches[i % 2])
    t0 = time.perf_counter()
    for i in range(steps):
        tr.step(*batches[i % 2])
    dt = (time.perf_counter() - t0) / steps
    return {"value": batch / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"{steps} steps of bs={batch} at {SIZE}x{SIZE} fp32 ({model_name}, oracle/step.py restatement of "
                      f"reference train.py:284-310 on my network plugin, {cores} torch threads, world 1)",
            "ms_per_step": dt * 1e3}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    res = cpu_reference(args.model, args.cpu_batch, steps, max(1, min(args.warmup, 3)))
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": max(1, min(args.warmup, 3)), "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": f"TestModel {args.model} {SIZE}x{SIZE}, one training iteration (fwd+BCE/CEL+bwd+SGD), "
                                   f"CPU sample bs={args.cpu_batch}"},
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------
# roofline of the dominant hand-written kernel
# ----------------------------------------------------------------------------------------------------
def bn_roofline(trace, dtype, iters=5):
    """replay every SyncBN backward launch of one iteration (same shapes, same fusion flags), alone, timed with
    CUDA events on the launching stream; L2 is flushed between launches by a 256 MB write."""
    from distributed_sod_project_b200.syncbn import raw_backward
    esz = 2 if dtype != torch.float32 else 4
    layers = []
    for (n, c, h, w, has_pre, has_res, relu) in trace:
        mk = lambda: torch.randn((n, c, h, w), device="cuda", dtype=dtype).contiguous(memory_format=torch.channels_last)  # noqa: E731
        x, dy = mk(), mk()
        pre = mk() if has_pre else None
        y = mk() if relu else None
        weight = torch.ones(c, device="cuda")
        mean = torch.zeros(c, device="cuda"); invstd = torch.ones(c, device="cuda")
        elems = n * c * h * w
        # algorithmic bytes (10 B/elem for plain bf16 = two reads of dy,x + one write of dx), extended
        # for the fused operands: each extra read operand is read in both passes, each extra output written once
        reads = 2 + (1 if has_pre else 0) + (1 if relu else 0)
        byts = (2 * reads + 1 + (1 if has_res else 0)) * esz * elems
        layers.append(((dy, x, pre, y, weight, mean, invstd, relu, has_res), byts))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    total_ms, total_bytes = 0.0, 0
    for it in range(iters + 1):
        for (a, byts) in layers:
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            raw_backward(*a)                       # exactly one sod_syncbn_bwd launch
            e.record()
            e.synchronize()
            if it > 0:
                total_ms += s.elapsed_time(e); total_bytes += byts
    return total_bytes / (total_ms * 1e-3) / 1e9, total_ms / (iters * len(layers)), total_bytes / (iters * len(layers))


def run_b200_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from distributed_sod_project_b200 import _lib, syncbn
    from distributed_sod_project_b200.engine import Trainer
    from distributed_sod_project_b200.synthetic import synth_batch

    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    args.warmup = max(3, args.warmup)          # timing rule: at least 3 warm-up iterations
    torch.backends.cudnn.benchmark = True
    tr = Trainer(model_name=args.model, dtype=dtype, channels_last=True, report_items=False)
    nb = 4
    host = [synth_batch(1234 + rank + 100 * i, BS, SIZE) for i in range(nb)]
    host = [(x.pin_memory(), m.pin_memory()) for x, m in host]
    dev = [(x.cuda(non_blocking=True), m.cuda(non_blocking=True)) for x, m in host]
    h2d = host[0][0].numel() * 4 + host[0][1].numel() * 4

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        sync_all()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for i in range(steps):
            fn(i)
        e.record()
        sync_all()
        ms = torch.tensor([s.elapsed_time(e)], device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # -- device-resident arm ---------------------------------------------------------------------------
    syncbn.TRACE = []
    for i in range(args.warmup):
        tr.forward_backward_update(*dev[i % nb])
        if i == 0:
            trace, syncbn.TRACE = syncbn.TRACE, None
    l0 = _lib.launches
    with ClockSampler(local) as clk:
        ms = timed(lambda i: tr.forward_backward_update(*dev[i % nb]), args.steps)
    launches = _lib.launches - l0
    value = world * BS * args.steps / (ms * 1e-3)

    # -- end-to-end arm: pinned host batch in, loss out, every step ------------------------------------
    for i in range(min(args.warmup, 3)):
        tr.step_from_host(*host[i % nb])
    tr.last_loss()
    ms_e2e = timed(lambda i: (tr.step_from_host(*host[i % nb]), tr.last_loss() if i == args.steps - 1 else None), args.steps)
    e2e = world * BS * args.steps / (ms_e2e * 1e-3)
    if tr.world > 1 and hasattr(tr.model, "arena") and tr.model.arena is not None:
        tr.model.arena.check_error()

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    pk, pk_kind = peaks()
    from distributed_sod_project_b200.syncbn import SyncBatchNorm
    n_bn = sum(isinstance(mod, SyncBatchNorm) for mod in tr.module.modules())
    bw, avg_ms, avg_bytes = bn_roofline(trace[:n_bn], dtype)   # cp_res50 traces its recompute forwards too
    roof = {"bound": "hbm", "kernel": "syncbn_bwd_kernel (84 launches/iteration, replayed alone on the model's layer shapes, "
                                      "L2 flushed between launches)",
            "achieved": bw, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": bw / pk["hbm_gbs"], "peak_kind": pk_kind,
            "traffic": None, "avg_launch_ms": avg_ms, "algorithmic_bytes_per_launch": avg_bytes}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"TestModel {args.model} (ResNet-50 encoder) {SIZE}x{SIZE} bs={BS}/GPU, one training iteration: "
                                   "fwd (cuDNN convs NHWC + fused SyncBN kernels) + fused BCE/CEL fwd+bwd + bwd + "
                                   "fused (all-reduce+)SGD-momentum", "global_batch": BS * world,
                       "parallelism": f"dp{world}", "l2": "4 rotating input batches; per-iteration working set (>5 GB of "
                                                          "activations) far exceeds the 126 MB L2, no explicit flush"},
            "clocks": clk.summary(),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "roofline": roof}
    if not args.no_cpu_baseline:
        res = cpu_reference(args.model, args.cpu_batch, 6, 1)
        line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_sweep(args):
    """BASELINE config 5: all-reduce bus bandwidth 64 KB–256 MB, peer-memory kernels vs torch NCCL, same box."""
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if world < 2:
        print(json.dumps({"sweep": "needs world_size >= 2"}))
        return
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from distributed_sod_project_b200 import comm
    sizes = [64 << 10, 256 << 10, 1 << 20, 4 << 20, 16 << 20, 64 << 20, 99_625_220 // 16 * 16, 256 << 20]
    arena = comm.Arena(payload_bytes=(256 << 20) + 4096)
    off = arena.alloc(256 << 20)
    rows = []
    for nbytes in sizes:
        n = nbytes // 4
        buf = arena.view(off, n, torch.float32)
        ref = torch.empty(n, device="cuda")
        res = {"bytes": nbytes}
        for name, fn in (("nccl", lambda: dist.all_reduce(ref)),
                         ("sod_multimem" if arena.has_multicast else "sod_p2p", lambda: arena.allreduce_(off, n, algo=2)),
                         ("sod_p2p_forced", lambda: arena.allreduce_(off, n, algo=2, no_multimem=True)),
                         ("sod_one_shot", (lambda: arena.allreduce_(off, n, algo=1)) if nbytes <= (1 << 20) else None)):
            if fn is None:
                continue
            buf.fill_(1.0); ref.fill_(1.0)
            iters = 20 if nbytes <= (16 << 20) else 8
            for _ in range(3):
                fn()
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            for _ in range(iters):
                fn()
            e.record(); torch.cuda.synchronize()
            t = torch.tensor([s.elapsed_time(e) / iters], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            us = float(t.item()) * 1e3
            res[name] = {"us": us, "bus_gbs": 2 * (world - 1) / world * nbytes / (us * 1e-6) / 1e9}
        rows.append(res)
    arena.check_error()
    if rank == 0:
        print(json.dumps({"sweep": "allreduce", "world": world, "multicast": arena.has_multicast, "nvlink_peak_gbs": 900,
                          "rows": rows}), flush=True)
    dist.barrier(); dist.destroy_process_group()


def main():
    args = parse()
    if args.sweep:
        return run_sweep(args)
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    main()

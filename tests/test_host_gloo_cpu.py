"""world_size-2 gloo tests (CPU) of host logic that only exists at world > 1: the evaluation shards and the reduction of
the metric sums over the ranks (`metrics.SaliencyMetrics.show`), and the ownership arithmetic of the sharded momentum
(`FusedSGD.shard_bounds`, which the checkpoint gather relies on)."""
import os
import socket

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _metrics_worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle import metrics as om
    from distributed_sod_project_b200.evaluate import shard
    from distributed_sod_project_b200.metrics import SaliencyMetrics
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(ROOT, "tests", "golden", "metrics_kat.npz"))
    n = int(g["n"])
    mine = list(shard(n))                                   # indices rank, rank + world, ...
    assert mine == list(range(rank, n, world))
    cal = SaliencyMetrics(num=n)
    for i in mine:                                          # what the two kernels would have produced for my images
        p8, g8 = g[f"pred{i}"], g[f"gt{i}"]
        head, hist = om.emulate_kernels(p8, g8)
        cal._pending.append((torch.tensor(head)[None], torch.tensor(hist)[None], p8.shape[0], p8.shape[1], None))
    res = cal.show()                                        # all-reduce over the gloo group
    np.save(os.path.join(out_dir, f"res{rank}.npy"), np.array([res[k] for k in ("MaxF", "MeanF", "MAE", "SM", "EM")]))
    dist.destroy_process_group()


def test_metric_sums_are_reduced_over_the_ranks(tmp_path, golden):
    import torch.multiprocessing as mp
    mp.spawn(_metrics_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    g = golden("metrics_kat.npz")
    want = dict(zip((str(k) for k in g["show_keys"]), (float(v) for v in g["show_vals"])))
    r0, r1 = (np.load(os.path.join(tmp_path, f"res{r}.npy")) for r in range(2))
    assert np.array_equal(r0, r1)                           # every rank reports the same dataset numbers
    for v, k in zip(r0, ("MaxF", "MeanF", "MAE", "SM", "EM")):
        assert v == pytest.approx(want[k], rel=2e-7), k      # == the reference's single-process result over all images


def test_shard_bounds_partition_the_flat_buffer():
    from distributed_sod_project_b200.optim import FusedSGD

    class Fake:
        pass
    for numel in (64, 4096 + 64, 24_907_648, 1_000_000 + 64):
        for world in (1, 2, 3, 4, 8):
            fake = Fake(); fake.flat = Fake(); fake.flat.numel = numel
            spans = [FusedSGD.shard_bounds(fake, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == numel
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))           # contiguous, no overlap
            assert all(lo % 4 == 0 and hi % 4 == 0 for lo, hi in spans)          # float4 granularity, as in csrc/sgd.cu
            nvec = numel // 4
            shard = (nvec + world - 1) // world
            assert all(hi - lo <= 4 * shard for lo, hi in spans)

"""world_size-2 gloo run of the oracle's data-parallel restatement (SyncBN over gloo + flat gradient mean) against
the golden W=2 trajectory that tools/make_golden.py produced with the reference model in ONE process on the
rank-concatenated batch.  Pins the equivalence  W-rank SyncBN+DDP == single process on the concatenated batch."""
import os
import socket

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from oracle.step import OracleTrainer
    from distributed_sod_project_b200 import network
    from distributed_sod_project_b200.synthetic import synth_batch
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(4)
    tr = OracleTrainer(network.res50, world_size=world, seed=0)
    losses, per_rank = [], []
    for it in range(3):
        x, m = synth_batch(1234 + rank + 1000 * it, 4, 128)
        preds = tr.model(x)
        from oracle.step import flat_allreduce_mean, total_loss
        loss, _ = total_loss(preds, m, tr.loss_funcs)
        tr.optimizer.zero_grad(); loss.backward()
        flat_allreduce_mean(list(tr.model.parameters()), world)
        tr.optimizer.step()
        per_rank.append(float(loss))
    np.save(os.path.join(out_dir, f"loss_rank{rank}.npy"), np.array(per_rank))
    w = dict(tr.model.named_parameters())["classifier.weight"].detach().numpy().ravel()
    np.save(os.path.join(out_dir, f"w_rank{rank}.npy"), w)
    dist.destroy_process_group()


def test_gloo_world2_matches_single_process_golden(tmp_path, golden):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    g = golden("step_res50_w2_s128.npz")
    for r in range(2):
        got = np.load(tmp_path / f"loss_rank{r}.npy")
        want = np.array([g[f"loss{it}"][r] for it in range(3)])
        np.testing.assert_allclose(got[:2], want[:2], rtol=2e-4)
        np.testing.assert_allclose(got[2], want[2], rtol=2e-3)
    assert np.array_equal(np.load(tmp_path / "w_rank0.npy"), np.load(tmp_path / "w_rank1.npy"))
    np.testing.assert_allclose(np.load(tmp_path / "w_rank0.npy")[:32], g["param2/classifier.weight"][:32], rtol=5e-2, atol=1e-4)

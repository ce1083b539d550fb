"""End-to-end parity: the whole training iteration (model → fused loss → backward through the SyncBN kernels →
fused SGD) against the trajectory of the UNMODIFIED reference (tests/golden/step_*.npz, CPU fp32)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _trainer(model_name, **kw):
    from distributed_sod_project_b200.engine import Trainer
    return Trainer(model_name=model_name, lr=0.05, momentum=0.9, weight_decay=5e-4, **kw)


@pytest.mark.parametrize("tag,model", [("res50_w1_s128", "res50"), ("res50_w1_s64", "res50"), ("cp_res50_w1_s64", "cp_res50")])
def test_fp32_trajectory_vs_reference(golden, tag, model):
    from distributed_sod_project_b200.synthetic import synth_batch
    g = golden(f"step_{tag}.npz")
    world, bs, size, iters = (int(v) for v in g["meta"])
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.deterministic = True          # the reference sets it too (utils/misc.py:52)
    tr = _trainer(model, dtype=torch.float32, channels_last=True)
    for it in range(min(iters, 3)):
        x, m = synth_batch(1234 + 1000 * it, bs, size)
        out = tr.step(x.cuda(), m.cuda())
        ref = float(g[f"loss{it}"][0])
        # north_star tolerance: 1e-3 relative on the loss.  The s64 configs (bs 2 at 64x64: 8 samples under the
        # deepest BN) are an ill-conditioned edge case whose later iterations amplify fp32 reassociation noise —
        # the oracle itself moves by 3e-3 at iteration 1 between a 1-process and a 2-process run
        well = tag.endswith("s128")
        if well or it == 0:
            assert out["loss"] == pytest.approx(ref, rel=1e-3 if it < 2 else 5e-3), f"iter {it}: got {out['loss']} want {ref}"
        else:   # ill-conditioned edge case: later iterations only have to stay in the neighbourhood (run-to-run chaos)
            assert np.isfinite(out["loss"]) and out["loss"] == pytest.approx(ref, rel=0.15), f"iter {it}: got {out['loss']} want {ref}"
        if f"logits{it}" in g.files:
            ref_l = g[f"logits{it}"]
            got = out["preds"].float().cpu().numpy()
            assert np.abs(got - ref_l).max() / np.abs(ref_l).max() < 1e-3
        if it == 0:
            # report strings are "%.5f" of fp32 values: allow the last printed digit to differ
            assert np.allclose([float(v) for v in out["items"]], [float(v) for v in g["items0"][0]], atol=2.1e-5)
            # post-step parameters: compare the UPDATE (p1 - p0) — p0 is the seeded init, identical by construction
            from distributed_sod_project_b200 import network
            from distributed_sod_project_b200.utils import init_seed
            init_seed(0)
            p0 = dict(getattr(network, model)().named_parameters())
            sd = dict(tr.model.named_parameters())
            for k in g.files:
                if k.startswith("param0/"):
                    name = k.split("/", 1)[1]
                    init = p0[name].detach().reshape(-1)[:64].numpy()
                    got = sd[name].detach().reshape(-1)[:64].cpu().numpy() - init
                    want = g[k] - init
                    scale = np.abs(want).max()
                    if name.startswith("div_2"):
                        assert np.abs(got).max() == 0 and scale == 0          # never in a param group
                    else:
                        # bs 2 at 64x64 puts 8 samples under the deepest BN: per-parameter gradients of GPU fp32
                        # (stock torch BN or ours alike, tools/diag_step.py) sit ~2e-2 from CPU fp32 in max-norm
                        assert np.abs(got - want).max() <= (5e-2 if well else 1e-1) * scale + 1e-7, name


def test_bf16_first_step_within_tolerance(golden):
    """BASELINE config-1 shape (bs 4, 320²) in the B200 configuration (bf16 autocast, channels-last)."""
    from distributed_sod_project_b200.synthetic import synth_batch
    g = golden("step_res50_w1_s320.npz")
    tr = _trainer("res50", dtype=torch.bfloat16, channels_last=True)
    x, m = synth_batch(1234, 4, 320)
    out = tr.step(x.cuda(), m.cuda())
    assert out["loss"] == pytest.approx(float(g["loss0"][0]), rel=5e-3)   # bf16 activations: 2^-8 per op, averaged
    out2 = tr.step(*[t.cuda() for t in synth_batch(2234, 4, 320)])
    assert np.isfinite(out2["loss"])


def test_cuda_graph_replay_matches_eager(golden):
    """the captured iteration (device-epoch packet tags, baked launch parameters) replays to the same trajectory"""
    from distributed_sod_project_b200.synthetic import synth_batch
    g = golden("step_res50_w1_s128.npz")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    losses = {}
    for mode in (False, True):
        tr = _trainer("res50", dtype=torch.float32, channels_last=True, use_graph=mode, report_items=False)
        out = []
        for it in range(4):
            x, m = synth_batch(1234 + 1000 * it, 4, 128)
            red, items, _ = tr.forward_backward_update(x.cuda(), m.cuda())
            out.append(float(red))
        losses[mode] = out
    # tools/diag_graph.py: two IDENTICAL eager runs of this model already differ by ~1e-3 at iteration 2 and ~6e-3 at
    # iteration 3 (non-deterministic cuDNN wgrad + chaotic early training), so only the first two iterations can be
    # compared tightly; replay vs eager sits inside that run-to-run band afterwards
    assert losses[True][:2] == pytest.approx(losses[False][:2], rel=2e-5)
    assert losses[True][2:] == pytest.approx(losses[False][2:], rel=2e-2)
    assert losses[True][0] == pytest.approx(float(g["loss0"][0]), rel=1e-3)
    assert losses[True][1] == pytest.approx(float(g["loss1"][0]), rel=1e-3)


def test_bf16_shadow_weights_exact_on_a_deterministic_net():
    """conv(+bias) → BN kernel → conv, no atomics anywhere: the shadow path (bf16 copy written by the fused step, bf16
    gradients folded in by it) must reproduce the plain autocast path bit for bit"""
    import torch.nn as nn
    from distributed_sod_project_b200 import amp
    from distributed_sod_project_b200.optim import make_optimizer
    from distributed_sod_project_b200.syncbn import convert_syncbn_model

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.div_4 = nn.Conv2d(8, 16, 3, padding=1)
            self.bn = nn.BatchNorm2d(16)
            self.head = nn.Conv2d(16, 8, 1, bias=False)

        def forward(self, x):
            return self.head(torch.relu(self.bn(self.div_4(x))))

    outs = {}
    for shadow in (False, True):
        torch.manual_seed(0)
        net = Net().cuda().to(memory_format=torch.channels_last)
        opt = make_optimizer(net, "f3_trick", dict(lr=0.05, momentum=0.9, weight_decay=5e-4, nesterov=False))
        net = convert_syncbn_model(net)
        net, opt = amp.initialize(net, opt, opt_level="O1", dtype=torch.bfloat16, shadow_weights=shadow)
        g = torch.Generator().manual_seed(1)
        for it in range(3):
            x = torch.randn(4, 8, 12, 12, generator=g).cuda().contiguous(memory_format=torch.channels_last)
            opt.zero_grad()
            net(x).float().square().mean().backward()
            opt.step()
        outs[shadow] = opt.flat.param.clone()
        assert (opt.flat.shadow16 is not None) == shadow
    assert torch.equal(outs[True], outs[False])


def test_bf16_shadow_weights_do_not_change_the_trajectory():
    """full model: same rounding points with and without the shadow copy; torch's bilinear-upsample backward uses
    atomics, so runs are only comparable up to that noise — checked per parameter tensor on the first update"""
    from distributed_sod_project_b200.synthetic import synth_batch
    res = {}
    for shadow in (False, True):
        tr = _trainer("res50", dtype=torch.bfloat16, channels_last=True, report_items=False, shadow_weights=shadow)
        assert (tr.optimizer.flat.shadow16 is not None) == shadow
        p0 = tr.optimizer.flat.param.clone()
        x, m = synth_batch(1234, 4, 128)
        red, _, _ = tr.forward_backward_update(x.cuda(), m.cuda())
        res[shadow] = (float(red), tr.optimizer.flat.param.clone() - p0)
        if shadow:   # the shadow is exactly the rounded master after every step
            assert torch.equal(tr.optimizer.flat.shadow16, tr.optimizer.flat.param.to(torch.bfloat16))
            assert float(tr.optimizer.flat.grad16.abs().max()) == 0.0
    # same forward up to how the folded conv biases enter the SyncBN kernel (bf16 shadow vs fp32 master: the BN output
    # is mathematically independent of them, what is left is bf16 rounding noise)
    assert res[True][0] == pytest.approx(res[False][0], rel=1e-3)
    da, db = res[True][1], res[False][1]
    # two bf16 runs whose rounding differs in a few places are only comparable in aggregate (exactness of the shadow
    # mechanism itself is pinned by test_bf16_shadow_weights_exact_on_a_deterministic_net)
    cos = torch.nn.functional.cosine_similarity(da.double(), db.double(), dim=0)
    assert float(cos) > 0.9, float(cos)


@pytest.mark.parametrize("prefetch", [False, True])
def test_step_from_host_feeds_every_batch(prefetch, monkeypatch):
    """`step_from_host` (bench.py's e2e arm): with and without the staged H2D prefetch, each iteration must consume exactly
    the batch it was given (a stale or half-copied staging buffer would show here) and report the same first losses as
    the device-resident entry."""
    from distributed_sod_project_b200 import engine
    from distributed_sod_project_b200.synthetic import synth_batch
    monkeypatch.setattr(engine, "PREFETCH_H2D", prefetch)
    batches = [tuple(t.pin_memory() for t in synth_batch(77 + 1000 * i, 4, 128)) for i in range(5)]
    ref = _trainer("res50", dtype=torch.bfloat16, channels_last=True, use_graph=True, report_items=False)
    want = [float(ref.forward_backward_update(x.cuda(), m.cuda())[0]) for x, m in batches[:2]]
    tr = _trainer("res50", dtype=torch.bfloat16, channels_last=True, use_graph=True, report_items=False)
    got = []
    for i, (x, m) in enumerate(batches):
        tr.step_from_host(x, m)
        if i < 2:
            got.append(tr.last_loss())
        else:       # let the host run ahead, as bench.py does; the static inputs are checked after the queue drains
            pass
    torch.cuda.synchronize()
    sx, sm = tr._cur["x"], tr._cur["m"]
    assert torch.equal(sx.cpu().reshape(-1), batches[-1][0].reshape(-1).to(sx.dtype))
    assert torch.equal(sm.cpu().reshape(-1), batches[-1][1].reshape(-1).to(sm.dtype))
    assert got == pytest.approx(want, rel=2e-3)
    assert np.isfinite(tr.last_loss())


@pytest.mark.parametrize("tag,bs", [("res50_w1_s320", 4), ("res50_w1_s320_bs16", 16)])
def test_fp32_benched_configuration_loss_and_logits(golden, tag, bs):
    """The configuration bench.py times (320², bs 4 = BASELINE config 1 and bs 16 = the B200 batch) against the
    UNMODIFIED reference in fp32: north_star's 1e-3 relative on the loss (iterations 0 and 1) AND on the per-pixel
    logits of iteration 0 (stored spatially subsampled, tools/make_golden.py).
    Iteration-1 logits are a different matter: they are the output of the network AFTER the first SGD step, and the
    reference's own float32 run sits 5.2e-2 (max) / 4.7e-3 (rms) of the logit range away from its float64 run there
    (tools/diag_reference_fp64.py → profiles/r02_reference_fp32_vs_fp64.txt; 3.4e-1 at iteration 2): the float32 golden
    vector is one sample of that rounding noise, so iteration 1 is held to the reference's own noise level, not to 1e-3."""
    from distributed_sod_project_b200.synthetic import synth_batch
    g = golden(f"step_{tag}.npz")
    stride = int(g["logits_stride"])
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    tr = _trainer("res50", dtype=torch.float32, channels_last=True)
    for it in range(2):
        x, m = synth_batch(1234 + 1000 * it, bs, 320)
        out = tr.step(x.cuda(), m.cuda())
        assert out["loss"] == pytest.approx(float(g[f"loss{it}"][0]), rel=1e-3), it
        ref_l = g[f"logits{it}"]
        got = out["preds"].float().cpu().numpy()[:, :, ::stride, ::stride]
        span = np.abs(ref_l).max()
        if it == 0:
            assert np.abs(got - ref_l).max() / span < 1e-3
        else:
            assert np.sqrt(np.mean((got - ref_l) ** 2)) / span < 1e-2
            assert np.abs(got - ref_l).max() / span < 1.5e-1


def test_bf16_benched_configuration_logits_bound(golden):
    """Same inputs in the benched arithmetic (bf16 autocast, bf16 shadow weights, channels-last, bs 16).  bf16 keeps 8
    mantissa bits: per-pixel logits of a randomly initialised 50-layer network cannot meet 1e-3 in ANY bf16 implementation.
    The stated bound is therefore relative to stock PyTorch: on the same batch, this engine's bf16 logits must be as close
    to the reference's fp32 logits as stock `torch.autocast(bfloat16)` logits are (rms within 1.5x, max within 2x), and
    the loss within 5e-3 relative."""
    from distributed_sod_project_b200 import network
    from distributed_sod_project_b200.synthetic import synth_batch
    from distributed_sod_project_b200.utils import init_seed
    g = golden("step_res50_w1_s320_bs16.npz")
    stride = int(g["logits_stride"])
    x, m = synth_batch(1234, 16, 320)
    ref_l = g["logits0"]
    span = np.abs(ref_l).max()
    init_seed(0)
    stock = network.res50().cuda().to(memory_format=torch.channels_last).train()
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        stock_l = stock(x.cuda().contiguous(memory_format=torch.channels_last)).float().cpu().numpy()[:, :, ::stride, ::stride]
    del stock
    tr = _trainer("res50", dtype=torch.bfloat16, channels_last=True)
    out = tr.step(x.cuda(), m.cuda())
    assert out["loss"] == pytest.approx(float(g["loss0"][0]), rel=5e-3)
    got = out["preds"].float().cpu().numpy()[:, :, ::stride, ::stride]
    rms = lambda a: float(np.sqrt(np.mean((a - ref_l) ** 2)) / span)          # noqa: E731
    mx = lambda a: float(np.abs(a - ref_l).max() / span)                       # noqa: E731
    print(f"bf16 logits vs fp32 reference: ours rms {rms(got):.3e} max {mx(got):.3e} | stock torch rms {rms(stock_l):.3e} max {mx(stock_l):.3e}")
    assert rms(got) <= 1.5 * rms(stock_l) + 1e-4
    assert mx(got) <= 2.0 * mx(stock_l) + 1e-3


def test_one_graph_per_input_size_follows_the_scheduler():
    """BASELINE config 4 mechanics: multi-scale batches (reference utils/dataset.py:125-132) and a per-iteration
    schedule (`sche_usebatch`, train.py:288-289) replay captured graphs — one per input size, never re-captured when the
    learning rate moves (the kernels read it from a device table).  The graph run must reproduce the eager run."""
    from distributed_sod_project_b200.synthetic import synth_batch
    sizes = [64, 96, 64, 128, 96, 64, 128]
    runs = {}
    for use_graph in (False, True):
        tr = _trainer("res50", dtype=torch.bfloat16, channels_last=True, use_graph=use_graph, report_items=False)
        sched = tr.scheduler(total_num=len(sizes), lr_type="poly")
        losses = []
        for it, size in enumerate(sizes):
            sched.step(tr.optimizer, curr_epoch=it)
            x, m = synth_batch(500 + it, 4, size)
            red, _, _ = tr.forward_backward_update(x.cuda(), m.cuda())
            losses.append(float(red))
        runs[use_graph] = (losses, tr.optimizer.flat.param.clone(), tr.optimizer.steps)
        if use_graph:
            assert len(tr._graphs) == 3                   # 64, 96, 128: captured once each
    assert runs[True][2] == runs[False][2] == len(sizes)
    # the first two iterations agree tightly; after that two bf16 runs that differ in a single rounding (cuDNN picks
    # non-deterministic algorithms) drift apart like any two runs of the reference do (profiles/r02_reference_fp32_vs_fp64.txt)
    assert runs[True][0][:2] == pytest.approx(runs[False][0][:2], rel=2e-3)
    assert runs[True][0] == pytest.approx(runs[False][0], rel=5e-2)
    # and the learning rate really is applied: a frozen schedule (lr 0) must leave the parameters alone
    tr = _trainer("res50", dtype=torch.bfloat16, channels_last=True, use_graph=True, report_items=False)
    x, m = synth_batch(1, 4, 64)
    tr.forward_backward_update(x.cuda(), m.cuda())
    for gr in tr.optimizer.param_groups:
        gr["lr"] = 0.0
    tr.optimizer.flat.mom.zero_()
    before = tr.optimizer.flat.param.clone()
    tr.forward_backward_update(x.cuda(), m.cuda())
    torch.cuda.synchronize()
    assert torch.equal(before, tr.optimizer.flat.param)


def test_cp_res50_iteration_is_capturable():
    """the default config (model cp_res50, cuda_graph on) captures an activation-checkpointed iteration: torch's
    checkpoint must not touch the CUDA RNG state during capture (preserve_rng_state=False in the plugin)"""
    from distributed_sod_project_b200.synthetic import synth_batch
    got = {}
    for use_graph in (False, True):
        tr = _trainer("cp_res50", dtype=torch.bfloat16, channels_last=True, use_graph=use_graph, report_items=False)
        got[use_graph] = [float(tr.forward_backward_update(*[t.cuda() for t in synth_batch(900 + i, 2, 64)])[0]) for i in range(3)]
    assert np.all(np.isfinite(got[True]))
    assert got[True] == pytest.approx(got[False], rel=5e-3)


def test_fp16_dynamic_loss_scaling_overflow_skip_halve_recover():
    """apex amp O1 as the reference runs it (train.py:183,299): fp16 autocast with a dynamic loss scale.  Driven through
    amp.initialize(dtype=float16) → scale_loss → FusedSGD.step: an overflowing scale must skip the update (parameters
    and momentum bit-identical, gradients cleared), halve the scale, and training must resume once the scale fits; after
    `growth_interval` clean steps the scale doubles."""
    from distributed_sod_project_b200 import amp
    from distributed_sod_project_b200.synthetic import synth_batch
    tr = _trainer("res50", dtype=torch.float16, channels_last=True, report_items=False)
    assert amp._cfg["dynamic"] and amp._cfg["scale"] == 65536.0 and not tr.use_graph
    amp._cfg["scale"] = 2.0 ** 40                    # certainly overflows fp16 gradients
    amp._cfg["growth_interval"] = 3
    flat = tr.optimizer.flat
    history = []
    try:
        for it in range(40):
            scale = amp._cfg["scale"]
            p0, v0 = flat.param.clone(), flat.mom.clone()
            x, m = synth_batch(3000 + it, 2, 64)
            red, _, _ = tr.forward_backward_update(x.cuda(), m.cuda())
            skipped = bool(int(tr.optimizer.found_inf.item()))
            history.append((scale, skipped))
            assert float(flat.grad.abs().max()) == 0.0                       # cleared either way
            if skipped:
                assert torch.equal(p0, flat.param) and torch.equal(v0, flat.mom)
                assert amp._cfg["scale"] == scale / 2
            else:
                assert not torch.equal(p0, flat.param)
                assert np.isfinite(float(red))
            if sum(1 for _, s in history if not s) >= 7:
                break
    finally:
        amp._cfg.update(enabled=False, dynamic=False, scale=1.0, good_steps=0, growth_interval=2000, found_inf=None)
    skips = [s for _, s in history]
    assert skips[0] and not skips[-1]                                        # overflowed first, recovered later
    first_ok = skips.index(False)
    assert all(skips[:first_ok])                                             # halved step by step until it fitted
    assert history[first_ok][0] == 2.0 ** 40 / 2 ** first_ok
    # growth: after 3 clean steps in a row the scale doubled (it may overflow again right after: that is the protocol)
    clean_run, grew = 0, False
    for (s0, sk0), (s1, _) in zip(history, history[1:]):
        clean_run = 0 if sk0 else clean_run + 1
        if clean_run and clean_run % 3 == 0 and s1 == 2 * s0:
            grew = True
    assert grew
    sd = amp.state_dict()
    assert set(sd["loss_scaler0"]) == {"loss_scale", "unskipped"}

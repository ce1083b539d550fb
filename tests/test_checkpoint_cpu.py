"""Checkpoint compatibility (reference utils/pipeline_ops.py:46-143): key layout, FusedSGD ⇄ torch.optim.SGD state."""
import torch
import torch.nn as nn

from distributed_sod_project_b200 import amp
from distributed_sod_project_b200.checkpoint import resume_checkpoint, save_checkpoint
from distributed_sod_project_b200.optim import make_optimizer


class _Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.div_2 = nn.Conv2d(3, 8, 3)
        self.div_4 = nn.Conv2d(8, 8, 3)
        self.head = nn.Conv2d(8, 1, 1)


class _Wrapper(nn.Module):            # stands in for DistributedDataParallel: exposes `.module`
    def __init__(self, m):
        super().__init__()
        self.module = m


INFO = dict(lr=0.05, momentum=0.9, weight_decay=5e-4, nesterov=False)


def test_round_trip_and_layout(tmp_path):
    torch.manual_seed(0)
    net = _Net()
    opt = make_optimizer(net, "f3_trick", INFO)
    opt.flat.mom.copy_(torch.randn_like(opt.flat.mom)); opt._stepped = True       # as after a few fused steps
    full, state = str(tmp_path / "full.pth.tar"), str(tmp_path / "state.pth")
    save_checkpoint(model=_Wrapper(net), optimizer=opt, amp=amp, exp_name="exp", current_epoch=7, full_net_path=full, state_net_path=state)
    ck = torch.load(full, weights_only=False)
    assert set(ck) == {"arch", "epoch", "net_state", "opti_state", "amp_state"} and ck["epoch"] == 7
    assert list(ck["net_state"]) == list(net.state_dict())                            # un-prefixed names
    assert list(torch.load(state, weights_only=False)) == list(net.state_dict())
    n_opt = sum(len(g["params"]) for g in ck["opti_state"]["param_groups"])
    assert n_opt == 4 and all("momentum_buffer" in s for s in ck["opti_state"]["state"].values())   # div_2.* in no group
    assert [g["lr"] for g in ck["opti_state"]["param_groups"]] == [0.1 * 0.05, 0.05]

    torch.manual_seed(1)
    net2 = _Net()
    opt2 = make_optimizer(net2, "f3_trick", INFO)
    assert resume_checkpoint(model=net2, optimizer=opt2, amp=amp, exp_name="exp", load_path=full, mode="all") == 7
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), net2.state_dict().values()))
    for (n, p1), (_, p2) in zip(net.named_parameters(), net2.named_parameters()):   # frozen div_2.* / padding carry no state
        if not n.startswith("div_2"):
            assert torch.equal(opt.flat.momentum_view(p1), opt2.flat.momentum_view(p2))
    # parameters stayed views of the flat buffer
    assert net2.head.weight.data_ptr() >= opt2.flat.param.data_ptr()
    net3 = _Net()
    assert resume_checkpoint(model=_Wrapper(net3), load_path=state, mode="onlynet") is None
    assert torch.equal(net3.head.weight, net.head.weight)
    try:
        resume_checkpoint(model=net3, optimizer=opt2, exp_name="other", load_path=full, mode="all")
        raise AssertionError("arch mismatch must raise, as in the reference")
    except Exception as e:                                                              # noqa: BLE001
        assert "does not match" in str(e)


def test_reference_style_sgd_state_loads_into_fused_sgd(tmp_path):
    """a checkpoint written with torch.optim.SGD (what the reference's make_optimizer returns) loads into FusedSGD"""
    torch.manual_seed(0)
    net = _Net()
    named = list(net.named_parameters())
    groups = [{"params": [p for n, p in named if n.startswith("div") and not n.startswith("div_2")], "lr": 0.005},
              {"params": [p for n, p in named if not n.startswith("div")], "lr": 0.05}]
    ref_opt = torch.optim.SGD(groups, momentum=0.9, weight_decay=5e-4)
    for p in net.parameters():
        p.grad = torch.randn_like(p)
    ref_opt.step()
    sd = ref_opt.state_dict()
    net2 = _Net(); net2.load_state_dict(net.state_dict())
    opt = make_optimizer(net2, "f3_trick", INFO)
    opt.load_state_dict(sd)
    for (n, p_ref), (_, p) in zip(net.named_parameters(), net2.named_parameters()):
        if n.startswith("div_2"):
            continue
        assert torch.equal(opt.flat.momentum_view(p), ref_opt.state[p_ref]["momentum_buffer"])
    back = opt.state_dict()
    assert all(torch.equal(back["state"][k]["momentum_buffer"], sd["state"][k]["momentum_buffer"]) for k in sd["state"])

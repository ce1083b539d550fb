"""One training iteration of the reference (train.py:284-310), wired onto the B200 kernels.

`Trainer` performs exactly the construction sequence of the reference's `main_worker` (train.py:141-207):
model factory → make_optimizer → convert_syncbn_model → amp.initialize → DDP → loss list; and `step` is the
loop body: forward, get_total_loss, zero_grad, backward, step, loss mean.  It is what `train.py`, `bench.py`,
the parity tests and `__graft_entry__.smoke()` all drive, so the measured path is the shipped path.
"""
from __future__ import annotations

import torch

from . import amp, comm, network
from .loss import CEL, BCEWithLogitsLoss, get_total_loss
from .optim import CustomScheduler, make_optimizer
from .parallel import DistributedDataParallel
from .syncbn import convert_syncbn_model
from .utils import init_seed


class Trainer:
    def __init__(self, model_name: str = "res50", lr: float = 0.05, momentum: float = 0.9, weight_decay: float = 5e-4,
                 nesterov: bool = False, optim: str = "f3_trick", reduction: str = "mean", use_aux_loss: bool = True,
                 dtype: torch.dtype = torch.bfloat16, channels_last: bool = True, seed: int = 0,
                 device: torch.device | None = None, report_items: bool = True):
        if not torch.cuda.is_available():
            raise RuntimeError("the B200 engine needs a CUDA device")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.world = comm.world_size()
        self.channels_last = channels_last
        self.report_items = report_items
        init_seed(seed)                                                    # train.py:113
        model = getattr(network, model_name)().to(self.device)            # train.py:141
        if channels_last:
            model = model.to(memory_format=torch.channels_last)           # before the optimizer flattens storage
        self.optimizer = make_optimizer(model, optim, dict(lr=lr, momentum=momentum, weight_decay=weight_decay,
                                                           nesterov=nesterov))    # train.py:157
        model = convert_syncbn_model(model)                                # train.py:180 (also at world 1: fused BN)
        self.use_amp = dtype != torch.float32
        if self.use_amp:
            model, self.optimizer = amp.initialize(model, self.optimizer, opt_level="O1", dtype=dtype)  # train.py:183
        if self.world > 1:
            model = DistributedDataParallel(model, delay_allreduce=True)   # train.py:185
        self.model = model
        self.loss_funcs = [BCEWithLogitsLoss(reduction=reduction)]         # train.py:203
        if use_aux_loss:
            self.loss_funcs.append(CEL())                                  # train.py:204-207
        self.model.train()
        self._pinned_loss = torch.zeros(1, dtype=torch.float32).pin_memory()
        self._loss_event: torch.cuda.Event | None = None

    @property
    def module(self):
        return self.model.module if hasattr(self.model, "module") else self.model

    def scheduler(self, total_num: int, lr_type: str = "poly", lr_decay: float = 0.9, warmup_epoch: int = 1):
        return CustomScheduler(self.optimizer, total_num, lr_type, dict(lr_decay=lr_decay, warmup_epoch=warmup_epoch))

    # ----------------------------------------------------------------------------------------------
    def forward_backward_update(self, x: torch.Tensor, m: torch.Tensor):
        """device tensors in, device loss out; no host synchronisation"""
        if self.channels_last:
            x = x.contiguous(memory_format=torch.channels_last)
        preds = self.model(x)                                              # train.py:293
        fused = len(self.loss_funcs) == 2
        if fused and not self.report_items:
            from .loss import _FusedLossFn
            holder: list = []
            unit = not amp._cfg["dynamic"]
            loss = _FusedLossFn.apply(preds, m, self.loss_funcs[0].reduction, 1.0, 1.0, 1e-6, unit, 0, holder)
            items = holder[0]
        else:
            loss, items = get_total_loss(preds, m, self.loss_funcs, unit_upstream=not amp._cfg["dynamic"])  # train.py:295
        self.optimizer.zero_grad()                                         # train.py:297
        if self.use_amp:
            with amp.scale_loss(loss, self.optimizer) as scaled:           # train.py:299
                scaled.backward()
        else:
            loss.backward()                                                # train.py:302
        self.optimizer.step()                                              # train.py:303
        reduced = comm.allreduce_tensor(loss.detach()) if self.world > 1 else loss.detach()   # train.py:306
        return reduced, items, preds

    def step(self, x: torch.Tensor, m: torch.Tensor) -> dict:
        """the loop body with the reference's reporting: returns python floats / strings (host sync)."""
        reduced, items, preds = self.forward_backward_update(x, m)
        return dict(loss=float(reduced.item()), items=items if isinstance(items, list) else
                    [f"{v:.5f}" for v in items[:2].tolist()], preds=preds.detach())

    def step_from_host(self, x_pinned: torch.Tensor, m_pinned: torch.Tensor):
        """end-to-end form: pinned host batch → device (train.py:291-292) → iteration → loss back to pinned host
        memory.  The D2H read is asynchronous; `last_loss()` waits for it."""
        x = x_pinned.to(self.device, non_blocking=True)
        m = m_pinned.to(self.device, non_blocking=True)
        reduced, _, _ = self.forward_backward_update(x, m)
        self._pinned_loss.copy_(reduced.reshape(1), non_blocking=True)
        self._loss_event = torch.cuda.Event()
        self._loss_event.record()

    def last_loss(self) -> float:
        if self._loss_event is not None:
            self._loss_event.synchronize()
        return float(self._pinned_loss[0])

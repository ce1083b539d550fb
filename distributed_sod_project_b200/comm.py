"""Symmetric-memory communicator for the peer-memory collectives (one process per GPU, one NVSwitch domain).

Replaces, for the hot path only, the NCCL communicator the reference creates with
`dist.init_process_group(backend="nccl", init_method="env://")` (reference train.py:73-77).
`torch.distributed` stays as the bootstrap: rendezvous of the CUDA VMM handles / multicast object is done
by `torch.distributed._symmetric_memory`; every byte of the hot-path traffic is then moved by the kernels
in csrc/ (multimem.* through the NVLS multicast mapping when the fabric offers one, plain peer
loads/stores otherwise).

An `Arena` is one symmetric allocation: [signal flags | bump-allocated payload].  `offset`s handed to the
C ABI are relative to the arena base, identical on every rank.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


def dist_ready() -> bool:
    return dist.is_available() and dist.is_initialized()


def world_size() -> int:
    return dist.get_world_size() if dist_ready() else 1


def rank() -> int:
    return dist.get_rank() if dist_ready() else 0


class Arena:
    def __init__(self, payload_bytes: int, group=None, device: torch.device | None = None, timeout_s: float = 30.0):
        if not dist_ready():
            raise _lib.SodError("Arena needs an initialised torch.distributed process group")
        import torch.distributed._symmetric_memory as symm
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        if self.world > _lib.SOD_MAX_WORLD:
            raise _lib.SodError(f"world size {self.world} exceeds SOD_MAX_WORLD={_lib.SOD_MAX_WORLD} (one NVSwitch domain)")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.flag_bytes = int(_lib.lib().sod_comm_flag_bytes())
        self.nbytes = self.flag_bytes + ((int(payload_bytes) + 255) // 256) * 256
        self.buf = symm.empty(self.nbytes, dtype=torch.uint8, device=self.device)
        self.handle = symm.rendezvous(self.buf, self.group.group_name)
        self.buf.zero_()
        self.error_flag = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.block_seq = torch.zeros(4 * 1024, dtype=torch.int32, device=self.device)   # per-block barrier counters
        torch.cuda.synchronize(self.device)
        dist.barrier(self.group)            # every rank's flags are zero before anyone signals
        ptrs = list(self.handle.buffer_ptrs)
        mc = int(getattr(self.handle, "multicast_ptr", 0) or 0)
        self.c = _lib.sod_comm()
        self.c.rank, self.c.world = self.rank, self.world
        for i in range(_lib.SOD_MAX_WORLD):
            self.c.peer[i] = int(ptrs[i]) if i < self.world else 0
        self.c.mc = mc
        self.c.arena_bytes = self.nbytes
        self.c.error_flag = self.error_flag.data_ptr()
        self.c.timeout_cycles = int(timeout_s * 1.9e9)
        self.c.block_seq = self.block_seq.data_ptr()
        self.has_multicast = mc != 0
        self._top = self.flag_bytes

    # -- allocation --------------------------------------------------------------------------------
    def alloc(self, nbytes: int, align: int = 256) -> int:
        off = (self._top + align - 1) // align * align
        if off + nbytes > self.nbytes:
            raise _lib.SodError(f"symmetric arena exhausted: need {nbytes} at {off}, have {self.nbytes}")
        self._top = off + nbytes
        return off

    def view(self, offset: int, numel: int, dtype: torch.dtype) -> torch.Tensor:
        esz = torch.empty((), dtype=dtype).element_size()
        return self.buf[offset:offset + numel * esz].view(dtype)

    @property
    def ref(self):
        return C.byref(self.c)

    def check_error(self):
        v = int(self.error_flag.item())
        if v:
            raise _lib.SodError(f"device-side barrier timeout (flag 0x{v & 0xffffffff:08x}): a peer rank is not making progress")

    # -- plain all-reduce (sweep, scalar mean) -----------------------------------------------------------
    def allreduce_(self, offset: int, numel: int, scale: float = 1.0, algo: int = 0, no_multimem: bool = False,
                   force_multimem: bool = False):
        flags = (_lib.SOD_ALGO_NO_MULTIMEM if no_multimem else 0) | (_lib.SOD_ALGO_FORCE_MULTIMEM if force_multimem else 0)
        rc = _lib.lib().sod_allreduce_f32(self.ref, offset, numel, float(scale), int(algo), flags, _lib.stream_ptr())
        _lib.check(rc, "sod_allreduce_f32")
        _lib.count_launch()


_default_small: Arena | None = None


def small_arena() -> Arena | None:
    """Lazily created arena for the latency-bound exchanges: SyncBN statistics slots + scalar loss mean.
    None when not distributed (world 1)."""
    global _default_small
    if world_size() == 1:
        return None
    if _default_small is None:
        slots = 4 * int(_lib.lib().sod_syncbn_exchange_bytes(4096))
        _default_small = Arena(slots + 4096)
        a = _default_small
        a.bn_slot_bytes = int(_lib.lib().sod_syncbn_exchange_bytes(4096))
        a.bn_slots = [a.alloc(a.bn_slot_bytes) for _ in range(4)]
        a.scalar_off = a.alloc(256)
    return _default_small


def reset():
    global _default_small
    _default_small = None


def allreduce_tensor(tensor: torch.Tensor) -> torch.Tensor:
    """Mean over ranks of a small fp32 tensor — reference utils/tensor_ops.py:60-64 (used at train.py:306)."""
    a = small_arena()
    if a is None:
        return tensor.clone()
    n = tensor.numel()
    if n > 64:
        raise _lib.SodError("allreduce_tensor is the scalar/logging path (≤64 elements)")
    pad = (n + 3) // 4 * 4
    stage = a.view(a.scalar_off, pad, torch.float32)
    stage.zero_()
    stage[:n].copy_(tensor.detach().reshape(-1).float())
    a.allreduce_(a.scalar_off, pad, scale=1.0 / a.world, algo=0)     # auto: two-shot over NVLS when there is a multicast mapping
    return stage[:n].clone().view_as(tensor)

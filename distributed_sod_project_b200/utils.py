"""Small host utilities with the reference's behaviour (utils/misc.py): seeding, running average, banners,
experiment naming and output paths.  Logging / xlsx / TensorBoard recorders are out of scope."""
from __future__ import annotations

import os
import random
from collections import OrderedDict
from datetime import datetime

import numpy as np
import torch


def init_seed(seed: int) -> None:
    """reference utils/misc.py:38-43"""
    os.environ["PYTHONHASHSEED"] = str(seed)
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


def init_cudnn(benchmark: bool = True, deterministic: bool = True) -> None:
    """reference utils/misc.py:46-58 (always deterministic there)"""
    torch.backends.cudnn.enabled = True
    torch.backends.cudnn.benchmark = benchmark
    torch.backends.cudnn.deterministic = deterministic


class AvgMeter:
    """reference utils/misc.py:16-30"""

    def __init__(self):
        self.reset()

    def reset(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / self.count


def construct_print(out_str: str, total_length: int = 80) -> None:
    """banner print with the reference's exact format (utils/misc.py:330-336): ` ===>> text <<=== `, the rule
    shrinking with the text and collapsing to `==` once the text reaches `total_length`"""
    text = str(out_str)
    rule = "==" if len(text) >= total_length else "=" * ((total_length - len(text)) // 2 - 4)
    print(f" {rule}>> {text} <<{rule} ")


def check_mkdir(path: str) -> None:
    os.makedirs(path, exist_ok=True)


def construct_exp_name(cfg: dict) -> str:
    """experiment name derived from the config, same recipe as reference utils/misc.py:167-198"""
    focus = OrderedDict(input_size="size", batch_size="bs", lr="lr", epoch_num="e", use_amp="amp",
                        is_distributed="dist", size_list="ms", version="v")
    name = f"{cfg['model']}"
    for key, tag in focus.items():
        item = cfg[key]
        if isinstance(item, bool):
            item = "Y" if item else "N"
        elif isinstance(item, (list, tuple)):
            item = "Y" if item else "N"
        elif isinstance(item, str):
            if not item:
                continue
        elif item is None:
            item = "N"
        if isinstance(item, str):
            item = item.lower()
        name += f"_{tag.upper()}{item}"
    return name


def construct_path_dict(proj_root: str, exp_name: str, xlsx_name: str) -> dict:
    """output/<exp>/{tb,pre,pth} layout of reference utils/misc.py:201-232"""
    ckpt = os.path.join(proj_root, "output")
    log = os.path.join(ckpt, exp_name)
    pth = os.path.join(log, "pth")
    day = str(datetime.now())[:10]
    return {"ckpt_path": ckpt, "pth_log": log, "tb": os.path.join(log, "tb"), "save": os.path.join(log, "pre"),
            "pth": pth, "final_full_net": os.path.join(pth, "checkpoint_final.pth.tar"),
            "final_state_net": os.path.join(pth, "state_final.pth"), "tr_log": os.path.join(log, f"tr_{day}.txt"),
            "te_log": os.path.join(log, f"te_{day}.txt"), "cfg_log": os.path.join(log, f"cfg_{day}.txt"),
            "trainer_log": os.path.join(log, f"trainer_{day}.txt"), "xlsx": os.path.join(ckpt, xlsx_name)}


def write_data_to_file(data_str: str, file_path: str) -> None:
    with open(file_path, encoding="utf-8", mode="a") as f:
        f.write(data_str + "\n")

"""Model plugins, looked up by name: `getattr(network, user_config["model"])()` (reference train.py:141)."""
from .testmodel import cp_res50, load_pretrained_backbone, res50  # noqa: F401
from .blocks import bn_act  # noqa: F401

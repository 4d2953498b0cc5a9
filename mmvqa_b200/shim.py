"""Drop-in shim: make ``import models...`` in the reference's scripts resolve to this package.

    import mmvqa_b200.shim; mmvqa_b200.shim.install()      # before `from models.mmbert import Model`

vqamed2019/train.py, eval.py and pretrain/roco_*.py import ``models.mmbert``, ``models.asl_singlelabel``,
``models.SupConLoss.loss`` ... (vqamed2019/train.py:21-22, pretrain/roco_supcon_train.py:14-19); after
``install()`` those names are aliases of ``mmvqa_b200.models.*`` and the scripts run unchanged."""
import importlib
import sys

_NAMES = ["", ".mmbert", ".transformer", ".realformer", ".serf", ".image_encoding", ".asl_singlelabel", ".SupConLoss",
          ".SupConLoss.loss"]


def install(alias: str = "models") -> None:
    for suffix in _NAMES:
        mod = importlib.import_module("mmvqa_b200.models" + suffix)
        sys.modules[alias + suffix] = mod


def uninstall(alias: str = "models") -> None:
    for suffix in _NAMES:
        sys.modules.pop(alias + suffix, None)

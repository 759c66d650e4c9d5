"""ORACLE / TEST INFRASTRUCTURE ONLY — never imported by the product path.

Imports the *unmodified* reference model code from /root/reference/Modules on
CPU, with the missing third-party CUDA packages replaced by the restatements
in oracle/stubs. It only works where
/root/reference exists (the build container); it does not travel to the GPU
box — golden fixtures generated from it (oracle/make_golden.py) do.
"""
from __future__ import annotations

import math
import os
import sys

import torch

REFERENCE_ROOT = os.environ.get("HGNN_REFERENCE_ROOT", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "Modules"))


def activate():
    """Put stubs + reference Modules on sys.path (stubs first)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    mod = os.path.join(REFERENCE_ROOT, "Modules")
    for p in (mod, _STUBS):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, mod)
    sys.path.insert(0, _STUBS)
    import torch.utils.checkpoint as ckpt
    if not getattr(ckpt, "_hgnn_patched", False):
        _orig = ckpt.checkpoint

        def _reentrant(fn, *args, **kw):
            kw.setdefault("use_reentrant", True)
            return _orig(fn, *args, **kw)
        ckpt.checkpoint = _reentrant
        ckpt._hgnn_patched = True


def reference_classes():
    activate()
    import gnn_utils  # noqa: F401  (reference module)
    import utils as ref_utils  # noqa: F401
    from EdgeClassifier.Models.IN import EC_InteractionGNN
    from BipartiteClassification.Models.HGNN_GMM import BC_HierarchicalGNN_GMM
    return {
        "gnn_utils": gnn_utils,
        "utils": ref_utils,
        "EC_InteractionGNN": EC_InteractionGNN,
        "BC_HierarchicalGNN_GMM": BC_HierarchicalGNN_GMM,
    }


def process_hparams(hparams):
    """Restated from training_utils.py:13-20 (that module cannot be imported:
    it pulls in cuml/wandb at import time)."""
    hp = dict(hparams)
    if hp.get("hidden") == "ratio":
        hp["hidden"] = hp["hidden_ratio"] * hp["latent"]
    hp.setdefault("cluster_granularity", 0)
    return hp


def kaiming_init(model):
    """Restated from training_utils.py:48-58: name-keyed normal init."""
    for name, p in model.named_parameters():
        if name.endswith(".bias"):
            p.data.zero_()
        elif p.dim() < 2:
            continue  # the reference hits IndexError on 1-D weights and skips them
        elif name.endswith("0.weight"):
            p.data.normal_(0, 1 / math.sqrt(p.shape[1]))
        else:
            p.data.normal_(0, math.sqrt(2) / math.sqrt(p.shape[1]))


def load_yaml_hparams(which, **overrides):
    import yaml
    rel = {"EC": "EdgeClassifier/Configs/IN.yaml",
           "BC": "BipartiteClassification/Configs/HGNN_GMM.yaml"}[which]
    with open(os.path.join(REFERENCE_ROOT, "Modules", rel)) as f:
        hp = yaml.safe_load(f)
    hp.update(overrides)
    return process_hparams(hp)

"""Model registry / hparam handling / init (reference: Modules/training_utils.py).

The reference reads per-model YAML files; here the same keys live in
``DEFAULT_HPARAMS`` (a YAML path may also be given) and are overridden by
``sweep_configs`` in the same way.
"""
from __future__ import annotations

import math

_COMMON = dict(
    input_dir="/data/FNAL/events", datatype_names=["train", "val", "test"], train_split=[300, 10, 10],
    ptcut=1.0, n_hits=5, primary=False, weight_leak=1.0, weight_min=0.5, pt_interval=0.5,
    noise=True, hard_ptcut=0, edge_dropping_ratio=0.0, score_cut=0.7, spatial_channels=3,
    hidden="ratio", hidden_ratio=2, nb_node_layer=3, nb_edge_layer=2, output_layers=3,
    hidden_activation="GELU", layernorm=True, share_weight=False, log_weight_ratio=0,
    warmup=100, lr=0.001, patience=1, majority_cut=0.5,
)

DEFAULT_HPARAMS = {
    # EdgeClassifier/Configs/IN.yaml
    "EC-IN": dict(_COMMON, model="EC-IN", true_edges="pid_true_edges", remove_isolated=True, latent=128,
                  n_interaction_graph_iters=14, hidden_output_activation="GELU", factor=0.98, max_epochs=200,
                  emb_epoch=30),
    # BipartiteClassification/Configs/HGNN_GMM.yaml
    "BC-HGNN-GMM": dict(_COMMON, model="BC-HGNN-GMM", remove_isolated=False, latent=256, emb_dim=8,
                        n_interaction_graph_iters=6, n_hierarchical_graph_iters=6, hidden_output_activation="Tanh",
                        train_r=1.0, factor=0.99, max_epochs=500, emb_epoch=100, bipartitegraph_sparsity=5,
                        supergraph_sparsity=10, min_cluster_size=3, cluster_granularity=5),
}
_ALIASES = {"1": "EC-IN", "4": "BC-HGNN-GMM"}


def process_hparams(hparams):
    """``hidden: ratio`` -> hidden_ratio * latent; default cluster_granularity (training_utils.py:13-20)."""
    if hparams.get("hidden") == "ratio":
        hparams["hidden"] = hparams["hidden_ratio"] * hparams["latent"]
    hparams.setdefault("cluster_granularity", 0)
    return hparams


def load_hparams(model_name, sweep_configs=None, yaml_path=None):
    name = _ALIASES.get(str(model_name), str(model_name))
    if yaml_path is not None:
        import yaml
        with open(yaml_path) as f:
            base = yaml.safe_load(f)
    elif name in DEFAULT_HPARAMS:
        base = dict(DEFAULT_HPARAMS[name])
    else:
        raise ValueError("Can't Find Model Name {}!".format(model_name))
    return process_hparams({**base, **(sweep_configs or {})})


def model_selector(model_name, sweep_configs=None, yaml_path=None):
    """EC-IN ("1") and BC-HGNN-GMM ("4") — the two models on the hot path (training_utils.py:22-46)."""
    name = _ALIASES.get(str(model_name), str(model_name))
    hp = load_hparams(name, sweep_configs, yaml_path)
    if name == "EC-IN":
        from .EdgeClassifier.Models.IN import EC_InteractionGNN
        return EC_InteractionGNN(hp)
    if name == "BC-HGNN-GMM":
        from .BipartiteClassification.Models.HGNN_GMM import BC_HierarchicalGNN_GMM
        return BC_HierarchicalGNN_GMM(hp)
    raise ValueError("Can't Find Model Name {}!".format(model_name))


def kaiming_init(model):
    """Name-keyed normal init (training_utils.py:48-58): biases 0; ``*0.weight`` (first layer of
    each MLP) ~ N(0, 1/fan_in); other matrices ~ N(0, 2/fan_in); 1-D weights untouched.
    Written through the tensors themselves under no_grad (not through ``.data``), so the parameters' version counters
    move and the packed bf16 weight images of the tensor-core kernels are rebuilt on the next forward."""
    import torch
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith(".bias"):
                p.zero_()
            elif p.dim() < 2:
                continue
            elif name.endswith("0.weight"):
                p.normal_(0, 1 / math.sqrt(p.shape[1]))
            else:
                p.normal_(0, math.sqrt(2) / math.sqrt(p.shape[1]))
    invalidate_packed_weights(model)


def invalidate_packed_weights(model):
    """Drop every cached bf16 weight image below ``model``. Needed only after writes that bypass autograd's version
    counter (``p.data.copy_(...)``, ``p.data.normal_()``, raw pointer updates): in-place ops on the parameter itself and
    optimizer steps bump the counter and are picked up automatically."""
    for m in model.modules():
        drop = getattr(m, "_drop_packed", None)
        if drop is not None:
            drop()


def load_from_pretrained(model, path=None, ckpt=None):
    import torch
    if ckpt is None:
        ckpt = torch.load(path, map_location="cpu")
    model.load_state_dict(ckpt["state_dict"], strict=False)
    return model

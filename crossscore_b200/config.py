"""Attribute-style config mirroring the reference's Hydra tree for the hot path.

The reference reads ``cfg.model.*`` by attribute (task/core.py:39-56, 73-74;
model/cross_reference.py:20-37; values from config/model/model.yaml:1-32).  OmegaConf is
not required: any object with the same attributes works (DictConfig, SimpleNamespace).
"""
from types import SimpleNamespace as NS


def default_cfg(**overrides):
    """config/model/model.yaml as a SimpleNamespace tree.

    Keyword overrides use dotted names with '__' as the separator, e.g.
    ``default_cfg(model__predict__metric__min=-1)``.
    """
    cfg = NS(
        model=NS(
            patch_size=14,
            do_reference_cross=True,
            decoder_do_self_attn=True,
            decoder_do_short_cut=True,
            need_attn_weights=False,
            need_attn_weights_head_id=0,
            backbone=NS(from_pretrained="facebook/dinov2-small"),
            pos_enc=NS(multi_view=NS(interpolate_mode="bilinear", req_grad=False, h=40, w=40)),
            loss=NS(fn="l1"),
            predict=NS(metric=NS(type="ssim", min=0, max=1, power_factor="default")),
        )
    )
    for key, val in overrides.items():
        node = cfg
        parts = key.split("__")
        for p in parts[:-1]:
            node = getattr(node, p)
        setattr(node, parts[-1], val)
    return cfg


def check_metric_prediction_config(metric_type, metric_min, metric_max):
    """Same accept/reject table and error text as utils/check_config.py:1-28."""
    valid_type = metric_type in ("ssim", "mse", "mae")
    valid_max = metric_max == 1
    if metric_type == "ssim":
        valid_min = metric_min in (-1, 0)
    elif metric_type in ("mse", "mae"):
        valid_min = metric_min == 0
    else:
        valid_min = False
    if not valid_type:
        raise ValueError(f"Invalid metric type {metric_type}")
    if not (valid_min and valid_max):
        raise ValueError(f"Invalid metric range {metric_min} to {metric_max} for {metric_type}")


def resolve_score_activation(metric_type, metric_min, metric_max, power_factor="default"):
    """RegressionLayer's choice of activation and exponent (model/regression_layer.py:31-62).

    Returns (use_tanh: bool, power: float).  power == 1.0 means Identity.
    """
    check_metric_prediction_config(metric_type, metric_min, metric_max)
    if metric_min == -1:
        return True, 1.0
    if metric_min != 0:
        raise ValueError(f"metric_min={metric_min} not supported")
    if power_factor == "default":
        p = {"ssim": 1, "mae": 2, "mse": 4}[metric_type]
    else:
        p = power_factor
    return False, float(p)  # float("some_typo") raises ValueError like the reference

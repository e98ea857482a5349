"""Shared helpers for the tests (golden fixture loading, oracle import)."""
import ast
import os

import numpy as np
import torch

from crossscore_b200.synthetic import make_inputs, make_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith("g"))  # model cases


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    rec = {k: z[k] for k in z.files}
    over = dict(ast.literal_eval(str(rec["cfg_over"]))) if "cfg_over" in rec else {}
    rec["cfg_over"] = over
    return rec


def golden_problem(rec):
    """Rebuild the seeded weights / inputs a golden case was generated with."""
    over = rec["cfg_over"]
    do_sa = over.get("model__decoder_do_self_attn", True)
    variant = str(rec["wvariant"]) if "wvariant" in rec else "benign"
    sd = make_state_dict(int(rec["seed_w"]), pe_h=int(rec["pe_h"]), pe_w=int(rec["pe_w"]), do_self_attn=do_sa,
                         variant=variant)
    q, r = make_inputs(int(rec["B"]), int(rec["N"]), int(rec["H"]), int(rec["W"]), int(rec["seed_x"]))
    return sd, q, r


def golden_pos_interp(rec):
    """Goldens without the field were generated with this container's transformers 5.5.0 (size= form)."""
    return str(rec["pos_interp"]) if "pos_interp" in rec else "size"


def oracle_kwargs(over):
    return dict(
        do_self_attn=over.get("model__decoder_do_self_attn", True),
        do_short_cut=over.get("model__decoder_do_short_cut", True),
        metric_type=over.get("model__predict__metric__type", "ssim"),
        metric_min=over.get("model__predict__metric__min", 0),
        power_factor=over.get("model__predict__metric__power_factor", "default"),
    )


def compare_to_golden(score: torch.Tensor, rec):
    """Return (max_abs, mean_abs) of a full score map against a golden record."""
    s = score.detach().float().cpu().numpy()
    assert tuple(s.shape) == tuple(rec["score_shape"]), (s.shape, rec["score_shape"])
    if "score" in rec:
        d = np.abs(s - rec["score"])
    else:
        off, step = int(rec["sub_off"]), int(rec["sub_step"])
        d = np.abs(s[:, off::step, off::step] - rec["score_sub"])
    return float(d.max()), float(d.mean())

"""Generate golden vectors by running the UNMODIFIED reference (build container only).

Imports task/core.py::CrossScoreNet from /root/reference with stub modules for the
packages that are absent offline (lightning, omegaconf, matplotlib, imageio -- none of
them touch the arithmetic), loads the seeded synthetic state_dict
(crossscore_b200.synthetic.make_state_dict, strict=True), runs the reference forward on
seeded inputs in fp32 on CPU and writes small .npz fixtures next to this file.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

/root/reference does not exist on the GPU box; tests read only the committed .npz files.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
from crossscore_b200.synthetic import make_inputs, make_state_dict  # noqa: E402
from crossscore_b200.config import default_cfg  # noqa: E402

REF = "/root/reference"


def boot_reference():
    sys.dont_write_bytecode = True
    sys.path[:0] = [REF, os.path.join(REF, "task")]

    def stub(name, **a):
        m = types.ModuleType(name)
        m.__dict__.update(a)
        sys.modules[name] = m

    for n in ["imageio", "matplotlib", "matplotlib.pyplot", "matplotlib.cm"]:
        stub(n)
    stub("matplotlib.patches", Rectangle=object)

    class _LM(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

    stub("lightning", LightningModule=_LM, seed_everything=lambda *a, **k: None)
    stub("lightning.pytorch")
    stub("lightning.pytorch.utilities", rank_zero_only=lambda f: f)
    stub("omegaconf", DictConfig=dict, ListConfig=list, open_dict=None,
         OmegaConf=types.SimpleNamespace(to_container=lambda c, resolve=True: {}, create=lambda d: d))
    from transformers import Dinov2Config, Dinov2Model
    small = dict(hidden_size=384, num_hidden_layers=12, num_attention_heads=6, mlp_ratio=4, image_size=518,
                 patch_size=14, layer_norm_eps=1e-6, qkv_bias=True, layerscale_value=1.0, hidden_act="gelu",
                 use_swiglu_ffn=False)
    Dinov2Config.from_pretrained = classmethod(lambda cls, name, **k: Dinov2Config(**small))
    Dinov2Model.from_pretrained = classmethod(lambda cls, name, **k: Dinov2Model(Dinov2Config(**small)))
    import core  # task/core.py
    return core


def run_case(core, name, B, N, H, W, seed_w=1, seed_x=0, need_w=False, head_id=0, subsample=None,
             cfg_over=None, pe=(40, 40), save_feats=False):
    cfg = default_cfg(**(cfg_over or {}))
    cfg.model.pos_enc.multi_view.h, cfg.model.pos_enc.multi_view.w = pe
    do_sa = cfg.model.decoder_do_self_attn
    net = core.CrossScoreNet(cfg).eval()
    sd = make_state_dict(seed_w, pe_h=pe[0], pe_w=pe[1], do_self_attn=do_sa)
    missing = net.load_state_dict(sd, strict=True)
    q, r = make_inputs(B, N, H, W, seed_x)
    with torch.inference_mode():
        out = net(q, r, need_w, head_id, False)
        feats = net.get_featmaps(q, r) if save_feats else None
    score = out["score_map_ref_cross"].float().numpy()
    rec = dict(B=B, N=N, H=H, W=W, seed_w=seed_w, seed_x=seed_x, need_w=int(need_w), head_id=head_id,
               pe_h=pe[0], pe_w=pe[1], score_shape=np.array(score.shape),
               score_mean=score.mean(axis=(-1, -2)).astype(np.float64))
    if subsample:
        off, step = subsample
        rec["sub_off"], rec["sub_step"] = off, step
        rec["score_sub"] = score[:, off::step, off::step].copy()
    else:
        rec["score"] = score
    if need_w:
        rec["attn"] = out["attn_weights_map_ref_cross"].float().numpy()
    if save_feats:
        rec["feat_query"] = feats["query"].float().numpy()
        rec["feat_ref"] = feats["ref_cross"].float().numpy()
    if cfg_over:
        rec["cfg_over"] = np.array(repr(sorted(cfg_over.items())))
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(name, score.shape, "mean", rec["score_mean"], "min/max", score.min(), score.max())


def main():
    torch.set_num_threads(os.cpu_count())
    core = boot_reference()
    # G1: tiny square grid (5x5 patches), DINOv2 pos-emb goes through the bicubic path
    run_case(core, "g1_tiny_70x70_n2", 1, 2, 70, 70, save_feats=True)
    # G2: non-square (6x8 patches, ragged 84x117 -> 84x112 map), batch 2, attention-weight export
    run_case(core, "g2_nonsquare_84x117_n3_attn", 2, 3, 84, 117, need_w=True, head_id=3, save_feats=True)
    # G3: the headline shape, 1 query + 5 refs at 518x518 (cfg 1); subsampled map
    run_case(core, "g3_518_n5", 1, 5, 518, 518, subsample=(3, 7))
    # G4: score-activation variants (regression_layer.py:31-62) on the tiny shape
    run_case(core, "g4_tanh", 1, 2, 70, 70, cfg_over=dict(model__predict__metric__min=-1))
    run_case(core, "g4_mae_pow2", 1, 2, 70, 70, cfg_over=dict(model__predict__metric__type="mae"))
    run_case(core, "g4_mse_pow4", 1, 2, 70, 70, cfg_over=dict(model__predict__metric__type="mse"))
    run_case(core, "g4_pow0p5", 1, 2, 70, 70, cfg_over=dict(model__predict__metric__power_factor=0.5))
    # G5: decoder switches (transformer.py:158-171)
    run_case(core, "g5_no_self_attn", 1, 2, 70, 70, cfg_over=dict(model__decoder_do_self_attn=False))
    run_case(core, "g5_no_short_cut", 1, 2, 70, 70, cfg_over=dict(model__decoder_do_short_cut=False))
    # G6: PE table whose grid equals the patch grid -> shortcut branch (positional_encoding.py:51-56)
    run_case(core, "g6_pe_shortcut_70x70", 1, 1, 70, 70, pe=(5, 5))
    # G7: 1 reference, more tokens than one attention tile (12x11 patches = 132 tokens)
    run_case(core, "g7_168x154_n1", 1, 1, 168, 154)


if __name__ == "__main__":
    main()

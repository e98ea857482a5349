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


def legacy_interpolate_pos_encoding(self, embeddings, height, width):
    """Dinov2Embeddings.interpolate_pos_encoding as published in transformers 4.33.3 -- the version the reference
    pins (environment.yaml:340), not installable here -- restated so that goldens of the pinned stack's behaviour
    exist next to those of this container's 5.5.0 (the only difference: scale_factor=(h+0.1)/37 instead of size=)."""
    import math
    num_patches = embeddings.shape[1] - 1
    num_positions = self.position_embeddings.shape[1] - 1
    if num_patches == num_positions and height == width:
        return self.position_embeddings
    class_pos_embed = self.position_embeddings[:, 0]
    patch_pos_embed = self.position_embeddings[:, 1:]
    dim = embeddings.shape[-1]
    height = height // self.config.patch_size
    width = width // self.config.patch_size
    height, width = height + 0.1, width + 0.1
    g = int(math.sqrt(num_positions))
    patch_pos_embed = patch_pos_embed.reshape(1, g, g, dim).permute(0, 3, 1, 2)
    patch_pos_embed = torch.nn.functional.interpolate(
        patch_pos_embed, scale_factor=(height / math.sqrt(num_positions), width / math.sqrt(num_positions)),
        mode="bicubic", align_corners=False)
    if int(height) != patch_pos_embed.shape[-2] or int(width) != patch_pos_embed.shape[-1]:
        raise ValueError("Width or height does not match with the interpolated position embeddings")
    patch_pos_embed = patch_pos_embed.permute(0, 2, 3, 1).view(1, -1, dim)
    return torch.cat((class_pos_embed.unsqueeze(0), patch_pos_embed), dim=1)


def run_case(core, name, B, N, H, W, seed_w=1, seed_x=0, need_w=False, head_id=0, subsample=None,
             cfg_over=None, pe=(40, 40), save_feats=False, variant="benign", pos_interp="size"):
    cfg = default_cfg(**(cfg_over or {}))
    cfg.model.pos_enc.multi_view.h, cfg.model.pos_enc.multi_view.w = pe
    do_sa = cfg.model.decoder_do_self_attn
    from transformers.models.dinov2 import modeling_dinov2 as MD
    if not hasattr(MD, "_stock_interp"):
        MD._stock_interp = MD.Dinov2Embeddings.interpolate_pos_encoding
    MD.Dinov2Embeddings.interpolate_pos_encoding = (legacy_interpolate_pos_encoding if pos_interp == "scale_factor"
                                                    else MD._stock_interp)
    net = core.CrossScoreNet(cfg).eval()
    sd = make_state_dict(seed_w, pe_h=pe[0], pe_w=pe[1], do_self_attn=do_sa, variant=variant)
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
    if variant != "benign":
        rec["wvariant"] = np.array(variant)
    if pos_interp != "size":
        rec["pos_interp"] = np.array(pos_interp)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **rec)
    print(name, score.shape, "mean", rec["score_mean"], "min/max", score.min(), score.max())


def main():
    torch.set_num_threads(os.cpu_count())
    core = boot_reference()
    only = sys.argv[1:]
    global run_case
    _run = run_case

    def run_case(core, name, *a, **k):  # `python make_golden.py g8 g9`: regenerate only the named cases
        if not only or any(name.startswith(o) for o in only):
            _run(core, name, *a, **k)

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
    # G8: BASELINE cfg 4 at its real size: 1 query x 64 reference views at 518x518 (87 616 reference tokens)
    run_case(core, "g8_518_n64", 1, 64, 518, 518, seed_x=3, subsample=(5, 9))
    # G9: BASELINE cfg 5's path: 1036x1036 (74x74 patches, T = 5477; DINOv2 pos-emb bicubic 37^2 -> 74^2), 4 refs
    run_case(core, "g9_1036_n4", 1, 4, 1036, 1036, seed_x=4, subsample=(6, 11))
    # G10: outlier-channel / wide-LayerScale weights (crossscore_b200.synthetic variant "outlier")
    run_case(core, "g10_outlier_168x154_n2", 1, 2, 168, 154, seed_w=3, seed_x=5, variant="outlier", save_feats=True)
    run_case(core, "g10_outlier_518_n5", 1, 5, 518, 518, seed_w=3, seed_x=6, variant="outlier", subsample=(2, 9))
    # G11: the pinned stack's position-embedding resample (transformers 4.33.3: scale_factor form) on the default
    # predict geometry (short side 518, no crop -> non-square 37x49 grid) and on a small ragged grid
    run_case(core, "g11_legacy_pos_84x117_n3", 2, 3, 84, 117, pos_interp="scale_factor", save_feats=True)
    run_case(core, "g11_legacy_pos_518x690_n2", 1, 2, 518, 690, seed_x=7, pos_interp="scale_factor", subsample=(4, 9))


if __name__ == "__main__":
    main()

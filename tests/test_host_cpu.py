"""CPU-side checks: the C-ABI library loads and exports every symbol the header declares, the module
mirrors the reference's interface (state_dict schema, constructor checks, error behaviour)."""
import ctypes
import os
import re

import pytest
import torch

from crossscore_b200 import CrossScoreNet, _lib, default_cfg, load_checkpoint
from crossscore_b200.synthetic import make_state_dict, state_dict_spec

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def lib_path():
    if not os.path.exists(_lib.LIB_PATH):
        from crossscore_b200.build import build
        build()
    return _lib.LIB_PATH


def header_symbols():
    src = open(os.path.join(ROOT, "include", "crossscore_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xs_[a-z0-9_]+)\s*\(", src)))


def test_abi_exports_every_declared_symbol(lib_path):
    lib = ctypes.CDLL(lib_path)
    names = header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/crossscore_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes signature table out of sync with the header"
    lib.xs_version.restype = ctypes.c_int
    assert lib.xs_version() == 1  # no compute call: works without a GPU


def test_state_dict_schema_matches_reference():
    net = CrossScoreNet(default_cfg())
    sd = net.state_dict()
    spec = state_dict_spec()
    assert len(sd) == 265 and sum(v.numel() for v in sd.values()) == 25_855_690
    assert list(sd.keys()) == [n for n, _ in spec]
    for n, shape in spec:
        assert tuple(sd[n].shape) == tuple(shape), n
    assert "img_mean_std" in dict(net.named_buffers())
    assert net.dinov2_cfg.hidden_size == 384
    assert not any(p.requires_grad for p in net.parameters())


def test_load_lightning_checkpoint(tmp_path):
    net = CrossScoreNet(default_cfg())
    ck = {"state_dict": make_state_dict(3, lightning_prefix=True), "hyper_parameters": {}}
    path = tmp_path / "fake.ckpt"
    torch.save(ck, path)
    res = load_checkpoint(net, str(path))
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(net.state_dict()["ref_cross.head.2.bias"], ck["state_dict"]["model.ref_cross.head.2.bias"])
    bad = make_state_dict(3)
    bad.pop("pos_enc_fn.PE")
    with pytest.raises(RuntimeError):
        net.load_state_dict(bad, strict=True)


def test_no_self_attn_schema():
    cfg = default_cfg(model__decoder_do_self_attn=False)
    net = CrossScoreNet(cfg)
    assert not any("self_attn" in k for k in net.state_dict())
    assert len(net.state_dict()) == 265 - 8


def test_constructor_rejects_unsupported_config():
    with pytest.raises(ValueError):
        CrossScoreNet(default_cfg(model__do_reference_cross=False))
    with pytest.raises(ValueError):
        CrossScoreNet(default_cfg(model__predict__metric__type="psnr"))
    with pytest.raises(ValueError):
        CrossScoreNet(default_cfg(model__predict__metric__min=-1, model__predict__metric__type="mae"))
    with pytest.raises(ValueError):
        CrossScoreNet(default_cfg(model__patch_size=16))


def test_forward_without_gpu_fails_loudly():
    net = CrossScoreNet(default_cfg())
    q = torch.zeros(1, 3, 70, 70)
    r = torch.zeros(1, 2, 3, 70, 70)
    with pytest.raises(RuntimeError):
        net(q, r, False, 0, False)


def test_product_never_imports_oracle():
    """oracle/ is test infrastructure: nothing under crossscore_b200/ may import or mention it."""
    pkg = os.path.join(ROOT, "crossscore_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt, os.path.join(dirpath, f)


def test_attention_key_split_heuristic():
    """Engine picks the number of key ranges from the wave efficiency of the persistent attention kernel."""
    from crossscore_b200.engine import attn_kv_splits, attn_units
    # layout 1 (pair kernel, 148 slots): units per batch = heads * (tiles // 2) + ceil(heads / 2) for an odd tile count
    assert attn_units(192, 6, 1370, 1) == (192 * 33, 148, 1.4)
    assert attn_units(1, 8, 1369, 1)[0] == 44 and attn_units(1, 5, 300, 1)[0] == 8 and attn_units(2, 8, 256, 1)[0] == 16
    assert attn_units(192, 6, 1370, 0) == (192 * 66, 296, 1.45)
    assert attn_kv_splits(192 * 33, 11) == 1             # cfg 2 DINOv2 attention: 42.8 waves, nothing to gain
    assert attn_kv_splits(32 * 44, 54) == 1              # cfg 2 cross-attention
    assert attn_kv_splits(44, 685) == 3                  # one query: 44 units -> 132 on 148 slots
    assert attn_kv_splits(8 * 22, 685) == 5              # cfg 5 cross-attention: 176 units = 1.19 waves -> 5.9 waves
    assert attn_kv_splits(44, 3) == 1                    # too few key blocks to split
    assert attn_kv_splits(88, 685, 296, 1.45) == 3       # the same single query in layout 0
    for tiles, nblk in [(1, 1), (7, 9), (300, 40), (5000, 2)]:
        n = attn_kv_splits(tiles, nblk)
        per = -(-nblk // n)
        assert 1 <= n <= 8 and (n - 1) * per < nblk     # every range owns at least one key block


def test_numa_binding_helper_is_a_noop_without_nvml():
    """runner.bind_to_gpu_numa never raises: no GPU / no NVML here -> 0 CPUs bound and the affinity mask untouched."""
    import os
    from crossscore_b200.runner import bind_to_gpu_numa
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa(0) == 0                    # opt-in: off by default
    os.environ["XS_NUMA_BIND"] = "1"
    try:
        assert bind_to_gpu_numa(0) == 0 or os.sched_getaffinity(0) <= before
    finally:
        del os.environ["XS_NUMA_BIND"]
    assert os.sched_getaffinity(0) <= before


def test_folded_layernorm_row_threshold():
    from crossscore_b200.engine import Engine
    assert not Engine.fold_ln_rows(6 * 1370) and Engine.fold_ln_rows(14 * 1370) and Engine.fold_ln_rows(192 * 1370)

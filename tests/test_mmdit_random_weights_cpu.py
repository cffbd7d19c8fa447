"""`random_flux_weights` / `random_hunyuan_weights` (device-side synthetic weights for benchmarks) have the layout the engines read:
a forward through each engine under the CPU kernel emulation runs, is finite and follows the preset skip schedule."""
import pytest
import torch

import magcache_b200 as mc
from magcache_b200 import mmdit
from magcache_b200 import patch as patch_mod

import emu_ops


@pytest.fixture()
def emulated(monkeypatch):
    monkeypatch.setattr(mmdit, "ops", emu_ops)
    monkeypatch.setattr(patch_mod, "ops", emu_ops)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))


def test_random_flux_weights_drive_the_engine(emulated):
    w = mmdit.random_flux_weights(torch.device("cpu"), heads=2, num_layers=2, num_single_layers=2, joint_dim=96, pooled_dim=48)
    assert w.ada_rows == (2 * 12 + 2 * 3 + 2) * 256 and w.ada_w.shape == (w.ada_rows, 256)
    model = mmdit.MMDiTHandle(mmdit.FluxEngine(w))
    mc.init_magcache_flux(model, 8, thresh=0.24, K=5, retention_ratio=0.1)
    g = torch.Generator().manual_seed(0)
    hs, enc, pooled = torch.randn(1, 48, 64, generator=g).bfloat16(), torch.randn(1, 24, 96, generator=g).bfloat16(), torch.randn(1, 48, generator=g).bfloat16()
    img_ids = torch.zeros(48, 3)
    img_ids[:, 1], img_ids[:, 2] = torch.arange(48) // 8, torch.arange(48) % 8
    kinds = []
    for i in range(8):
        out = model(hs, enc, pooled, torch.tensor([1.0 - i / 8]), img_ids, torch.zeros(24, 3), torch.tensor([3.5]), return_dict=False)[0]
        assert out.shape == (1, 48, 64) and bool(torch.isfinite(out.float()).all())
        kinds.append(int(model._mc_flux_engine.res_valid))
    assert model.cnt == 0 and mc.MagCacheConfig("flux", 0.24, 5, 0.1, 8, table="flux_dev").schedule().sum() > 0


def test_random_hunyuan_weights_drive_the_engine(emulated):
    w = mmdit.random_hunyuan_weights(torch.device("cpu"), heads=2, double_depth=2, single_depth=2, text_dim=96, pooled_dim=48)
    model = mmdit.MMDiTHandle(mmdit.HunyuanEngine(w))
    mc.init_magcache_hunyuan(model, 6)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 16, 2, 8, 8, generator=g).bfloat16()
    txt, pooled = torch.randn(1, 12, 96, generator=g).bfloat16(), torch.randn(1, 48, generator=g).bfloat16()
    mask = torch.ones(1, 12, dtype=torch.long)
    mask[0, 9:] = 0
    cos, sin = torch.ones(32, 128), torch.zeros(32, 128)
    for i in range(6):
        out = model(x, torch.tensor([900.0 - 100 * i]), txt, mask, pooled, cos, sin, torch.tensor([6000.0]))["x"]
        assert out.shape == x.shape and bool(torch.isfinite(out.float()).all())
    assert model.cnt == 0 and model._mc_hunyuan_engine.n_txt == 9

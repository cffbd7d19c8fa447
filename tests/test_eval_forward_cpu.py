"""Host logic of `magcache_eval_forward` (the paper-evaluation Wan forward, eval/magcache/experiments/Wan2.1_EVAL/wan_magcache.py:682-817)
on CPU with a stand-in engine: which branch each call takes, and the state kept under that script's attribute names. The kernels
themselves are exercised by tests/test_wan_forward_gpu.py::test_eval_variant_loop_vs_oracle."""
import torch

import magcache_b200 as mc
from magcache_b200 import patch


class FakeEngine:
    def __init__(self, n=6, d=4):
        self.res_buf = torch.zeros(2, n, d)
        self.res = [self.res_buf[0], self.res_buf[1]]
        self.res_valid = [False, False]
        self.kinds = []

    def forward(self, kind, slot):
        if kind == "hit":
            assert self.res_valid[slot]
        else:
            self.res[slot].fill_(float(len(self.kinds) + 1))
            self.res_valid[slot] = True
        self.kinds.append((kind, slot))
        return torch.full((2, 2), float(len(self.kinds)))


def test_eval_forward_schedule_and_attributes(monkeypatch, capsys):
    eng = FakeEngine()
    monkeypatch.setattr(patch, "_stage", lambda self, *a, **k: eng)
    model = type("M", (), {})()
    mc.init_magcache_eval(model, 50, thresh=0.12, K=4)
    cls = type(model)
    assert cls.forward is mc.magcache_eval_forward and cls.num_steps == 100 and len(cls.ratio) == 90 and cls.residual_cache is None
    want = mc.PRESETS["wan2.1-eval-fast-E012K4"].schedule().tolist()
    for call in range(100):
        t_before = model.t
        out = model.forward([None], None, [None], 0)
        kind, slot = eng.kinds[-1]
        assert (kind == "hit") == bool(want[call]) and slot == t_before % 2
        if t_before < 10:
            assert model.residual_cache is None                        # cache_time = 10 (:771, :796)
        else:
            rc = model.residual_cache
            assert tuple(rc.shape) == (2, 1, 6, 4, 1) and rc.data_ptr() == eng.res_buf.data_ptr()
            assert torch.equal(rc[t_before % 2][..., -1][0], eng.res[t_before % 2])   # how the reference reads it (:781)
        if t_before % 2 == 0:
            assert model.pre_con[0] is out[0]
        if call < 99:
            assert model.skip_steps == sum(want[:call + 1])
    assert sum(want) == 62 and model.t == 0 and model.skip_steps == 0
    assert model.accumulated_sim == [1.0, 1.0] and model.accumulated_steps == [0, 0] and model.accumulated_err == [0, 0]
    printed = capsys.readouterr().out
    assert printed.count("skip time") == 62 and "skip time 20, cur_scale:" in printed
    # a second video reuses the stale cache exactly like the reference (the tensor is never cleared)
    for call in range(24):
        model.forward([None], None, [None], 0)
    assert [k for k, _ in eng.kinds[100:124]] == ["hit" if w else "miss" for w in want[:24]]

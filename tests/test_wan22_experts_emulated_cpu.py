"""Wan2.2 A14B (MagCache4Wan2.2/magcache_generate.py:209-362) on the Wan engine, CPU with emulated kernels: TWO expert instances of one
class — high-noise model for the first `split_step` calls, low-noise model afterwards — sharing the tensor counter, the accumulators and
the residual-cache list exactly like the reference, each with its own weights / engine. Same hit / miss sequence as the oracle (and as the
golden Wan2.2 windows), outputs to bf16-pipeline noise, residual hand-over between the experts."""
import copy

import pytest
import torch

import magcache_b200 as mc
from magcache_b200 import patch as patch_mod
from magcache_b200 import wan as wan_mod
from oracle import wan_ref

import emu_ops


@pytest.fixture()
def emulated(monkeypatch):
    monkeypatch.setattr(wan_mod, "ops", emu_ops)
    monkeypatch.setattr(patch_mod, "ops", emu_ops)
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))


@pytest.mark.parametrize("mode", ["t2v", "i2v"])
def test_two_experts_share_controller_and_cache(emulated, mode):
    kw = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128, text_len=32)
    if mode == "i2v":
        kw.update(in_dim=36, model_type="i2v")  # Wan2.2 I2V: y under the latents, no CLIP branch
    protos = []
    for seed in (1, 2):
        m = wan_ref.WanModel(**{k: v for k, v in kw.items() if k != "model_type"}).init_synthetic(seed)
        m.model_type = kw.get("model_type", "t2v")
        protos.append(m)
    steps, high = 12, 5  # int(split_step * R) = 2: both cache slots are filled before the first eligible call
    ratios = mc.tables()["wan2.2_i2v_a14b" if mode == "i2v" else "wan2.2_t2v_a14b"][2:].tolist()
    RefCls = type("RefW22", (wan_ref.WanModel,), {})
    OurCls = type("OurW22", (wan_ref.WanModel,), {})
    refs, ours = [], []
    for p in protos:
        r, o = copy.deepcopy(p), copy.deepcopy(p)
        r.__class__, o.__class__ = RefCls, OurCls
        object.__setattr__(o, "_mc_engine", mc.WanEngine(mc.WanWeights.from_module(o, torch.device("cpu"))))
        refs.append(r)
        ours.append(o)
    wan_ref.install_magcache_wan22(RefCls, ratios, steps, thresh=0.12, K=2, retention_ratio=0.2, split_steps=high, mode=mode)
    mc.init_magcache_wan22(ours[0], ratios, steps, thresh=0.12, K=2, retention_ratio=0.2, split_steps=high, mode=mode)
    assert OurCls.mag_ratios.tolist() == RefCls.mag_ratios.tolist() and OurCls.split_step == 2 * high
    g = torch.Generator().manual_seed(0)
    lat = torch.randn(16, 2, 8, 8, generator=g)
    ys = [torch.randn(20, 2, 8, 8, generator=g)] if mode == "i2v" else None
    ctxs = [torch.randn(9, 128, generator=g), torch.randn(7, 128, generator=g)]
    kinds = []
    with torch.no_grad():
        for call in range(2 * steps):
            e = 0 if call < 2 * high else 1  # which expert the pipeline calls (boundary timestep)
            t = torch.tensor([950.0 - 40.0 * (call // 2)])
            a = refs[e]([lat], t=t, context=[ctxs[call % 2]], seq_len=32, y=ys)[0]
            b = ours[e]([lat], t=t, context=[ctxs[call % 2]], seq_len=32, y=ys)[0]
            kinds.append(int(refs[e].last_skip))
            rel = float((a - b).norm() / a.norm())
            assert rel < 2e-2, (mode, call, rel)
            if call < 2 * steps - 1:
                assert int(OurCls.cnt) == int(RefCls.cnt) == call + 1 and torch.is_tensor(OurCls.cnt)  # ONE counter for both experts
            assert OurCls.accumulated_err == RefCls.accumulated_err and OurCls.accumulated_steps == RefCls.accumulated_steps
    fam = "wan2.2-i2v" if mode == "i2v" else "wan2.2-t2v"
    want = mc.MagCacheConfig(fam, 0.12, 2, 0.2, steps, mag_ratios=OurCls.mag_ratios, high_noise_steps=high).schedule().tolist()
    assert kinds == want and 0 < sum(want) < 2 * steps
    assert any(kinds[2 * high:]), "the low-noise expert must reach hits (its residual comes from its own misses after the window)"
    assert int(OurCls.cnt) == 0


def _ti2v_pair(seed=0, **over):
    """A TI2V-5B-shaped model at test size (48 latent channels in and out -> three 16-channel head groups) as oracle and as ours."""
    kw = dict(dim=256, ffn_dim=512, num_heads=2, num_layers=2, text_dim=128, text_len=32, in_dim=48, out_dim=48)
    kw.update(over)
    proto = wan_ref.WanModel(**kw).init_synthetic(seed)
    ref, our = copy.deepcopy(proto), copy.deepcopy(proto)
    RefCls, OurCls = type("RefTI2V", (wan_ref.WanModel,), {}), type("OurTI2V", (wan_ref.WanModel,), {})
    ref.__class__, our.__class__ = RefCls, OurCls
    object.__setattr__(our, "_mc_engine", mc.WanEngine(mc.WanWeights.from_module(our, torch.device("cpu"))))
    return ref, our, RefCls, OurCls


def _ti2v_timesteps(step_t, grid, seq_len, first_frame_clean=True):
    """`temp_ts = (mask2[0][0][:, ::2, ::2] * timestep).flatten()` padded to seq_len with `timestep` (upstream Wan2.2 `generate` for
    ti2v image-to-video): the first latent frame is the clean image, its tokens carry t = 0."""
    f, h, w = grid
    t = torch.full((f, h, w), float(step_t))
    if first_frame_clean:
        t[0] = 0.0
    return torch.cat([t.flatten(), torch.full((seq_len - f * h * w,), float(step_t))]).unsqueeze(0)


def test_ti2v_per_token_timesteps_follow_the_oracle(emulated):
    """MagCache4Wan2.2/magcache_generate.py:260-272 with `t` [1, seq_len]: first-frame tokens at t = 0, the rest at the step's timestep,
    `split_step` None -> the plain retention window (:301-303). Same hit / miss sequence as the oracle, every output (48 channels) and
    the cached residual to bf16-pipeline noise — and NOT what a uniform timestep would give."""
    ref, our, RefCls, OurCls = _ti2v_pair()
    steps = 10
    ratios = mc.tables()["wan2.2_ti2v_5b_a"][2:].tolist()
    wan_ref.install_magcache_wan22(RefCls, ratios, steps, thresh=0.12, K=2, retention_ratio=0.2)
    mc.init_magcache_wan22(our, ratios, steps, thresh=0.12, K=2, retention_ratio=0.2)
    assert OurCls.split_step is None and OurCls.mag_ratios.tolist() == RefCls.mag_ratios.tolist()
    g = torch.Generator().manual_seed(3)
    lat = torch.randn(48, 3, 8, 8, generator=g)   # grid (3, 4, 4): 48 tokens, the first 16 are the image frame
    ctxs = [torch.randn(9, 128, generator=g), torch.randn(7, 128, generator=g)]
    kinds, seq_len = [], 56
    with torch.no_grad():
        for call in range(2 * steps):
            t = _ti2v_timesteps(950.0 - 90.0 * (call // 2), (3, 4, 4), seq_len)
            a = ref([lat], t=t, context=[ctxs[call % 2]], seq_len=seq_len)[0]
            b = our([lat], t=t, context=[ctxs[call % 2]], seq_len=seq_len)[0]
            assert a.shape == b.shape == (48, 3, 8, 8)
            kinds.append(int(ref.last_skip))
            rel = float((a - b).norm() / a.norm())
            assert rel < 2e-2, (call, rel)
            ra, rb = RefCls.residual_cache[call % 2][0, :48].float(), OurCls.residual_cache[call % 2][0].float()
            assert float((ra - rb).norm() / ra.norm()) < 2e-2
            assert OurCls.accumulated_err == RefCls.accumulated_err and OurCls.accumulated_steps == RefCls.accumulated_steps
            if call == 0:
                eng = our._mc_engine
                assert eng.t_values == 2 and eng.runs == [(0, 16, 0), (16, 48, 1)]
                uni = copy.deepcopy(ref)
                uni.__class__ = type("UniTI2V", (wan_ref.WanModel,), {})
                wan_ref.install_magcache_wan22(uni.__class__, ratios, steps, thresh=0.12, K=2, retention_ratio=0.2)
                c = uni([lat], t=_ti2v_timesteps(950.0, (3, 4, 4), seq_len, first_frame_clean=False), context=[ctxs[0]], seq_len=seq_len)[0]
                assert float((a - c).norm() / a.norm()) > 10 * rel, "the per-token timesteps must matter in this test"
    want = mc.MagCacheConfig("wan2.2-ti2v", 0.12, 2, 0.2, steps, mag_ratios=OurCls.mag_ratios).schedule().tolist()
    assert kinds == want and 0 < sum(want) < 2 * steps
    assert int(OurCls.cnt) == 0


def test_ti2v_uniform_and_scattered_timesteps(emulated):
    """A [1, seq_len] `t` that is uniform (TI2V text-to-video) takes the one-timestep path (no row ranges); a `t` with several
    non-adjacent ranges of the same value reuses that value's embedding; too many distinct values are refused, not approximated."""
    ref, our, RefCls, OurCls = _ti2v_pair(seed=1, num_layers=1)
    ratios = mc.tables()["wan2.2_ti2v_5b_a"][2:].tolist()
    wan_ref.install_magcache_wan22(RefCls, ratios, 10)
    mc.init_magcache_wan22(our, ratios, 10)
    g = torch.Generator().manual_seed(4)
    lat, ctx = torch.randn(48, 3, 8, 8, generator=g), torch.randn(6, 128, generator=g)
    eng = our._mc_engine
    with torch.no_grad():
        t = torch.full((1, 48), 700.0)
        a, b = ref([lat], t=t, context=[ctx], seq_len=48)[0], our([lat], t=t, context=[ctx], seq_len=48)[0]
        assert eng.runs is None and eng.t_values == 1 and float((a - b).norm() / a.norm()) < 2e-2
        t = torch.full((1, 48), 700.0)
        t[0, 5:9], t[0, 20:31], t[0, 40:] = 0.0, 0.0, 350.0
        a, b = ref([lat], t=t, context=[ctx], seq_len=48)[0], our([lat], t=t, context=[ctx], seq_len=48)[0]
        assert eng.t_values == 3 and eng.runs == [(0, 5, 0), (5, 9, 1), (9, 20, 0), (20, 31, 1), (31, 40, 0), (40, 48, 2)]
        assert float((a - b).norm() / a.norm()) < 2e-2
        with pytest.raises(NotImplementedError):
            our([lat], t=torch.arange(48.0).unsqueeze(0), context=[ctx], seq_len=48)


def test_timestep_ranges_reconstruct_t_exactly():
    """`WanEngine._stage_t` (the host reduction of a per-token `t`, MagCache4Wan2.2/magcache_generate.py:263-264): for random piecewise-
    constant timestep vectors the (row range, value index) list must reproduce `t` row by row, cover every row exactly once, reuse one
    index per distinct value — also clipped to a rank's rows of a token shard."""
    import numpy as np
    from magcache_b200.shard import TokenShard
    rng = np.random.default_rng(0)
    w = mc.WanWeights.random(mc.WanDims(256, 512, 2, 1, text_dim=64, text_len=8), torch.device("cpu"))
    for trial in range(40):
        n = int(rng.integers(5, 400))
        vals = rng.choice([0.0, 17.5, 250.0, 499.0, 731.25, 999.0], size=int(rng.integers(1, 5)), replace=False)
        cuts = np.sort(rng.choice(np.arange(1, n), size=min(n - 1, int(rng.integers(0, 7))), replace=False))
        t = np.empty(n)
        for seg, (a, b) in enumerate(zip(np.concatenate([[0], cuts]), np.concatenate([cuts, [n]]))):
            t[a:b] = vals[seg % len(vals)]
        for world in (1, 3):
            if world > 1 and n < world:
                continue
            for rank in range(world):
                eng = mc.WanEngine(w)
                eng.n_keys, eng.pad_row = n, 0
                eng.s_t = torch.zeros(wan_mod.MAX_T_VALUES, dtype=torch.float64)
                eng.shard = TokenShard(rank, world, n) if world > 1 else None
                eng._stage_t(torch.from_numpy(t).unsqueeze(0))
                lo, hi = (eng.shard.start, eng.shard.stop) if eng.shard is not None else (0, n)
                if eng.runs is None:
                    assert len(set(t.tolist())) == 1 and eng.t_values == 1 and float(eng.s_t[0]) == t[0]
                    continue
                got = np.full(hi - lo, np.nan)
                for r0, r1, u in eng.runs:
                    assert 0 <= r0 < r1 <= hi - lo and np.isnan(got[r0:r1]).all()
                    got[r0:r1] = float(eng.s_t[u])
                assert np.array_equal(got, t[lo:hi]), (trial, world, rank)
                assert eng.t_values == len(set(t.tolist())) == len(set(eng.s_t[:eng.t_values].tolist()))

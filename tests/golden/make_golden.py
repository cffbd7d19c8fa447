#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by EXECUTING the reference's own code.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

The MagCache scripts cannot be imported (`import wan` / `diffusers` / `hyvideo` are not installed),
so this script parses them with `ast`, cuts out exactly the statements that make up the hot-path
host logic and executes those statements, unmodified, against a stand-in `self`:

* `nearest_interp`                    MagCache4Wan2.1/magcache_generate.py:27-34 (whole function)
* the controller `if self.cnt>=...:`  MagCache4Wan2.1/magcache_generate.py:279-292
                                      MagCache4FLUX/magcache_flux.py:327-338
                                      MagCache4HunyuanVideo/magcache_sample_video.py:90-102
* the counter/reset epilogue          :306-311 / :431-436 / :149-154
* the calibration statistics block    MagCache4Wan2.1/magcache_generate.py:166-173
* the calibrated `mag_ratios` literals (:910,:912,:1002,:1004,:1142,:1144; FLUX :459; Hunyuan :316,:318)

Outputs (committed):  tables.json, nearest_interp.json, masks.json, calib_stats.json, extra_adapters.json, paper_eval_adapters.json
Nothing here is imported by the product; tests/ read the JSON files only.
"""
import ast
import json
import os
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

WAN = f"{REF}/MagCache4Wan2.1/magcache_generate.py"
FLUX = f"{REF}/MagCache4FLUX/magcache_flux.py"
HUN = f"{REF}/MagCache4HunyuanVideo/magcache_sample_video.py"
WAN22 = f"{REF}/MagCache4Wan2.2/magcache_generate.py"
QWEN = f"{REF}/MagCache4QwenImage/magcache_generate.py"
OMNI = f"{REF}/MagCache4OmniGen2/magcache/magcache_utils.py"
FRAMEPACK = f"{REF}/MagCache4FramePack/magcache_demo_gradio.py"
FRAMEPACK_F1 = f"{REF}/MagCache4FramePack/magcache_demo_gradio_f1.py"
EVAL_WAN = f"{REF}/eval/magcache/experiments/Wan2.1_EVAL/wan_magcache.py"
OPENSORA = f"{REF}/eval/magcache/experiments/opensora.py"


def _tree(path):
    with open(path) as f:
        return ast.parse(f.read())


def _func(tree, name):
    for n in ast.walk(tree):
        if isinstance(n, ast.FunctionDef) and n.name == name:
            return n
    raise KeyError(name)


def _mentions(node, attr):
    return any(isinstance(n, ast.Attribute) and n.attr == attr for n in ast.walk(node))


def _compile(stmts):
    mod = ast.Module(body=list(stmts), type_ignores=[])
    ast.fix_missing_locations(mod)
    return compile(mod, "<reference-extract>", "exec")


# ----------------------------------------------------------------------------------------------
# tables
# ----------------------------------------------------------------------------------------------
def extract_tables():
    """Evaluate every `np.array([1.0]*k + [...])` literal assigned to `.mag_ratios` in the three scripts."""
    out = {}

    def grab(path, key_of):
        tree = _tree(path)
        for n in ast.walk(tree):
            if isinstance(n, ast.Assign) and len(n.targets) == 1 and isinstance(n.targets[0], ast.Attribute) \
                    and n.targets[0].attr == "mag_ratios" and isinstance(n.value, ast.Call) \
                    and isinstance(n.value.func, ast.Attribute) and n.value.func.attr == "array":
                arr = eval(compile(ast.Expression(n.value), "<tbl>", "eval"), {"np": np})
                key = key_of(n.lineno)
                if key:
                    out[key] = {"source": f"{os.path.relpath(path, REF)}:{n.lineno}", "values": [float(v) for v in arr]}

    grab(WAN, lambda ln: {910: "wan2.1_t2v_14b", 912: "wan2.1_t2v_1.3b", 1002: "wan2.1_i2v_480p", 1004: "wan2.1_i2v_720p",
                          1142: "wan2.1_vace_1.3b", 1144: "wan2.1_vace_14b"}.get(ln))
    grab(FLUX, lambda ln: {459: "flux_dev"}.get(ln))
    grab(HUN, lambda ln: {316: "hunyuan_720p", 318: "hunyuan_544p"}.get(ln))
    return out


# ----------------------------------------------------------------------------------------------
# nearest_interp (the reference function itself)
# ----------------------------------------------------------------------------------------------
def reference_nearest_interp():
    fn = _func(_tree(WAN), "nearest_interp")
    env = {"np": np}
    exec(_compile([fn]), env)
    return env["nearest_interp"]


# ----------------------------------------------------------------------------------------------
# controller (the reference statements themselves, run against a stand-in `self`)
# ----------------------------------------------------------------------------------------------
class RefController:
    """Runs the controller + counter statements cut out of a reference magcache_forward."""

    def __init__(self, path, func="magcache_forward"):
        fn = _func(_tree(path), func)
        body = fn.body
        ctrl = [s for s in body if isinstance(s, ast.If) and _mentions(s.test, "retention_ratio")]
        assert len(ctrl) == 1, (path, len(ctrl))
        self.ctrl = _compile(ctrl)
        # `self.cnt += 1` and the following `if self.cnt >= self.num_steps:` reset
        tail = []
        for i, s in enumerate(body):
            if isinstance(s, ast.AugAssign) and isinstance(s.target, ast.Attribute) and s.target.attr == "cnt":
                tail = [s, body[i + 1]]
                assert isinstance(body[i + 1], ast.If) and _mentions(body[i + 1].test, "num_steps")
        assert tail, path
        self.tail = _compile(tail)

    def call(self, state):
        env = {"self": state, "np": np, "skip_forward": False}
        exec(self.ctrl, env)
        skip = bool(env["skip_forward"])
        exec(self.tail, env)
        return skip


def wan_state(table, steps, thresh, K, R, interp):
    s = types.SimpleNamespace()
    s.cnt = 0
    s.num_steps = steps * 2
    s.magcache_thresh = thresh
    s.K = K
    s.accumulated_err = [0.0, 0.0]
    s.accumulated_steps = [0, 0]
    s.accumulated_ratio = [1.0, 1.0]
    s.retention_ratio = R
    s.residual_cache = ["r0", "r1"]
    mr = np.array(table)
    if len(mr) != steps * 2:  # MagCache4Wan2.1/magcache_generate.py:915-919, executed literally
        con = interp(mr[0::2], steps)
        ucon = interp(mr[1::2], steps)
        mr = np.concatenate([con.reshape(-1, 1), ucon.reshape(-1, 1)], axis=1).reshape(-1)
    s.mag_ratios = mr
    return s


def scalar_state(table, steps, thresh, K, R, interp, cache_attr):
    s = types.SimpleNamespace()
    s.cnt = 0
    s.num_steps = steps
    s.magcache_thresh = thresh
    s.K = K
    s.accumulated_err = 0
    s.accumulated_steps = 0
    s.accumulated_ratio = 1
    s.retention_ratio = R
    setattr(s, cache_attr, "r")
    mr = np.array(table)
    if len(mr) != steps:
        mr = interp(mr, steps)
    s.mag_ratios = mr
    return s


def run_mask(ctrl, state, n_calls):
    mask = []
    for _ in range(n_calls):
        mask.append(1 if ctrl.call(state) else 0)
    return mask


def final_state(s):
    def f(v):
        if isinstance(v, (list, tuple)):
            return [float(x) for x in v]
        return float(v)
    return {"cnt": int(s.cnt), "accumulated_err": f(s.accumulated_err), "accumulated_steps": f(s.accumulated_steps),
            "accumulated_ratio": f(s.accumulated_ratio)}


# ----------------------------------------------------------------------------------------------
# calibration statistics (reference statements, torch CPU)
# ----------------------------------------------------------------------------------------------
def calib_cases():
    import torch
    import torch.nn.functional as F

    fn = _func(_tree(WAN), "magcache_calibration")
    blk = [s for s in fn.body if isinstance(s, ast.If) and isinstance(s.test, ast.Compare)
           and _mentions(s.test, "cnt") and not _mentions(s.test, "num_steps")
           and any(_mentions(x, "norm_ratio") for x in s.body)]
    assert len(blk) == 1
    code = _compile(blk)
    cases = []
    for seed, (B, N, D), spread in [(0, (1, 64, 96), 0.02), (1, (1, 257, 128), 0.2), (2, (2, 33, 1536), 0.05), (3, (1, 1000, 64), 0.5)]:
        g = torch.Generator().manual_seed(seed)
        r_prev = torch.randn(B, N, D, generator=g) * 0.1
        r_cur = r_prev * (0.97 + spread * torch.rand(B, N, 1, generator=g)) + 0.01 * torch.randn(B, N, D, generator=g)
        st = types.SimpleNamespace(cnt=2, residual_cache=[r_prev, None], norm_ratio=[], norm_std=[], cos_dis=[])
        env = {"self": st, "F": F, "torch": torch, "residual_x": r_cur, "round": round, "print": lambda *a, **k: None}
        exec(code, env)
        cases.append({"seed": seed, "shape": [B, N, D], "spread": spread,
                      "norm_ratio": st.norm_ratio[0], "norm_std": st.norm_std[0], "cos_dis": st.cos_dis[0],
                      "norm_ratio_raw": env["norm_ratio"], "norm_std_raw": env["norm_std"], "cos_dis_raw": env["cos_dis"]})
    return cases


# ----------------------------------------------------------------------------------------------
# "next" adapters: Wan2.2 expert windows, Qwen's linspace interpolation, OmniGen2's ceil/initial state
# ----------------------------------------------------------------------------------------------
def wan22_tables():
    """The three `mag_ratios = [...]` list literals of MagCache4Wan2.2/magcache_generate.py (:695 t2v-A14B, :736/:738 ti2v-5B
    (two resolutions), :771 i2v-A14B); the script prepends [1.0]*2 in init_magcache (:356)."""
    out = {}
    names = {695: "wan2.2_t2v_a14b", 736: "wan2.2_ti2v_5b_a", 738: "wan2.2_ti2v_5b_b", 771: "wan2.2_i2v_a14b"}
    for n in ast.walk(_tree(WAN22)):
        if isinstance(n, ast.Assign) and len(n.targets) == 1 and isinstance(n.targets[0], ast.Name) and n.targets[0].id == "mag_ratios" \
                and isinstance(n.value, ast.List) and n.lineno in names:
            vals = eval(compile(ast.Expression(n.value), "<tbl>", "eval"))
            out[names[n.lineno]] = {"source": f"MagCache4Wan2.2/magcache_generate.py:{n.lineno} (+ [1.0]*2 prefix, :356)",
                                    "values": [1.0, 1.0] + [float(v) for v in vals]}
    assert len(out) == 4, out.keys()
    return out


class RefControllerWan22:
    """Statements :290-317 (use_magcache window + controller) and :328-334 (counter) of MagCache4Wan2.2 magcache_forward."""

    def __init__(self):
        body = _func(_tree(WAN22), "magcache_forward").body
        i0 = next(i for i, s in enumerate(body) if isinstance(s, ast.Assign) and getattr(s.targets[0], "id", None) == "use_magcache")
        i1 = next(i for i, s in enumerate(body) if isinstance(s, ast.If) and isinstance(s.test, ast.Name) and s.test.id == "use_magcache")
        stmts = [s for s in body[i0:i1 + 1] if not (isinstance(s, ast.Assign) and getattr(s.targets[0], "id", None) == "ori_x")]
        self.ctrl = _compile(stmts)
        tail = []
        for i, s in enumerate(body):
            if isinstance(s, ast.AugAssign) and isinstance(s.target, ast.Attribute) and s.target.attr == "cnt":
                tail = [s, body[i + 1]]
        self.tail = _compile(tail)

    def call(self, state):
        env = {"self": state, "np": np}
        exec(self.ctrl, env)
        skip = bool(env["skip_forward"])
        exec(self.tail, env)
        return skip


def wan22_state(table, steps, thresh, K, R, interp, split_steps, mode):
    import torch
    s = wan_state(table, steps, thresh, K, R, interp)
    s.cnt = torch.tensor(0)  # as init_magcache does (:342): comparisons against Python floats then run in float32
    s.split_step = None if split_steps is None else split_steps * 2
    s.mode = mode
    return s


def qwen_nearest_interp():
    fn = _func(_tree(QWEN), "nearest_interp")
    env = {"np": np}
    exec(_compile([fn]), env)
    return env["nearest_interp"]


class RefControllerOmni:
    """OmniGen2 keeps the state in a `magcache_params` object; statements magcache_utils.py:342-354."""

    def __init__(self):
        fn = _func(_tree(OMNI), "magcache_forward")
        ctrl = [s for s in ast.walk(fn) if isinstance(s, ast.If) and _mentions(s.test, "retention_ratio")]
        assert len(ctrl) == 1
        self.ctrl = _compile(ctrl)

    def call(self, state):
        import math
        env = {"self": state, "np": np, "math": math, "skip_forward": False}
        exec(self.ctrl, env)
        return bool(env["skip_forward"])


def extra_cases(tables, interp):
    out = {"tables": wan22_tables(), "wan22_masks": [], "qwen_interp": [], "omnigen2_masks": []}
    ref = RefControllerWan22()
    for key, mode, has_split in [("wan2.2_t2v_a14b", "t2v", True), ("wan2.2_i2v_a14b", "i2v", True), ("wan2.2_ti2v_5b_a", "t2v", False)]:
        tbl = out["tables"][key]["values"]
        for steps in (len(tbl) // 2, 30, 25):
            for hs in ((None,) if not has_split else (steps // 4, steps // 3, steps // 2, (2 * steps) // 3)):
                for thresh, K, R in [(0.12, 2, 0.2), (0.06, 2, 0.2), (0.12, 4, 0.4), (0.24, 6, 0.1), (0.12, 2, 0.25), (0.12, 3, 1.0 / 3.0)]:
                    st = wan22_state(tbl, steps, thresh, K, R, interp, hs, mode)
                    n = 2 * steps * 2 + 5
                    m = run_mask(ref, st, n)
                    out["wan22_masks"].append({"table": key, "mode": mode if has_split else "ti2v", "steps": steps, "high_noise_steps": hs,
                                               "thresh": thresh, "K": K, "R": R, "calls": n, "mask": "".join(map(str, m)),
                                               "final": {"cnt": int(st.cnt), "accumulated_err": [float(v) for v in st.accumulated_err],
                                                         "accumulated_ratio": [float(v) for v in st.accumulated_ratio],
                                                         "accumulated_steps": [float(v) for v in st.accumulated_steps]}})
    qi = qwen_nearest_interp()
    for L, T in [(50, 50), (50, 40), (50, 1), (50, 2), (28, 20), (9, 5), (5, 3), (101, 41), (3, 7), (60, 50), (2, 2), (7, 1)]:
        src = np.arange(L, dtype=np.float64) * 1.25 + 0.5
        out["qwen_interp"].append({"L": L, "T": T, "src": [float(v) for v in src], "out": [float(v) for v in qi(src, T)]})
    # Qwen-Image controller over MORE than one image per process: its wrap statement resets only the counter
    # (MagCache4QwenImage/magcache_generate.py:243-244) — the accumulators of the last calls carry over into the next image.
    qref = RefController(QWEN)
    out["qwen_masks"] = []
    for key, tbl in remaining_tables().items():
        if not key.startswith("qwen"):
            continue
        for steps in (50, 40, 30):
            for thresh, K, R in [(0.06, 2, 0.2), (0.12, 4, 0.2), (0.24, 6, 0.1), (0.5, 8, 0.0)]:
                st = wan_state(tbl["values"], steps, thresh, K, R, qi)
                n = 3 * steps * 2 + 5  # three images and a bit
                m = run_mask(qref, st, n)
                out["qwen_masks"].append({"table": key, "steps": steps, "thresh": thresh, "K": K, "R": R, "calls": n, "mask": "".join(map(str, m)),
                                          "final": final_state(st)})
    omni = RefControllerOmni()
    import importlib.util  # MAG_RATIOS literal dict of magcache_utils.py:14-20
    omni_tbls = {}
    for n in ast.walk(_tree(OMNI)):
        if isinstance(n, ast.Assign) and getattr(n.targets[0], "id", None) == "MAG_RATIOS":
            omni_tbls = {k: [float(v) for v in arr] for k, arr in eval(compile(ast.Expression(n.value), "<t>", "eval"), {"np": np}).items()}
    out["tables"].update({f"omnigen2_{k}": {"source": "MagCache4OmniGen2/magcache/magcache_utils.py:14-20", "values": v} for k, v in omni_tbls.items()})
    for key, tbl in omni_tbls.items():
        for steps in (len(tbl), 30):
            for thresh, K, R in [(0.06, 3, 0.2), (0.12, 3, 0.25), (0.04, 2, 0.33)]:
                mr = np.array(tbl) if len(tbl) == steps else interp(np.array(tbl), steps)
                params = types.SimpleNamespace(mag_ratios=mr, previous_residual="r", accumulated_ratio=1.0, accumulated_err=0.0,
                                               accumulated_steps=3, cnt=0)  # dataclass defaults, magcache_utils.py:41-45
                st = types.SimpleNamespace(magcache_params=params, retention_ratio=R, num_steps=steps, magcache_thresh=thresh, K=K)
                mask = []
                for c in range(steps):  # the sampler owns cnt (magcache_utils.py:435-513): advance it here
                    params.cnt = c
                    mask.append(1 if omni.call(st) else 0)
                out["omnigen2_masks"].append({"table": f"omnigen2_{key}", "steps": steps, "thresh": thresh, "K": K, "R": R,
                                              "initial_accumulated_steps": 3, "mask": "".join(map(str, mask))})
    return out


# ----------------------------------------------------------------------------------------------
# FramePack and the two paper-evaluation forwards (SURVEY Appendix A rows FramePack / Wan2.1 eval / Open-Sora)
# ----------------------------------------------------------------------------------------------
def _parent_bodies(tree):
    for node in ast.walk(tree):
        for field in ("body", "orelse", "finalbody"):
            body = getattr(node, field, None)
            if isinstance(body, list):
                yield body


def _counter_tail(fn, attr):
    """`self.<attr> += 1` and the `if` that follows it, wherever they are nested."""
    for body in _parent_bodies(fn):
        for i, st in enumerate(body):
            if isinstance(st, ast.AugAssign) and isinstance(st.target, ast.Attribute) and st.target.attr == attr:
                assert isinstance(body[i + 1], ast.If)
                return [st, body[i + 1]]
    raise AssertionError(attr)


def _class_attr_literal(path, attr):
    """Value of the (un-commented) `<...>.__class__.<attr> = <expr>` assignment of a script, evaluated with numpy."""
    for n in ast.walk(_tree(path)):
        if isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Attribute) and n.targets[0].attr == attr:
            return eval(compile(ast.Expression(n.value), "<t>", "eval"), {"np": np})
    raise AssertionError((path, attr))


class RefControllerFramePack:
    """magcache_demo_gradio.py:252-270 (re-initialisation at cnt == 0, controller with the |1 - ratio| <= 0.06 veto and
    `cnt >= 1`) and :298-300 (counter); `initialize_magcache` (:63-74) is executed whole, table literal and interpolation included."""

    def __init__(self, path):
        tree = _tree(path)
        fn = _func(tree, "magcache_framepack_forward")
        outer = [st for st in fn.body if isinstance(st, ast.If) and _mentions(st.test, "enable_magcache")]
        assert len(outer) == 1
        body = outer[0].body
        head = []
        for st in body:
            head.append(st)
            if isinstance(st, ast.If) and _mentions(st.test, "retention_ratio"):
                break
        assert len(head) == 3, [type(x).__name__ for x in head]  # if cnt == 0 / skip_forward = False / controller
        self.ctrl = _compile(head)
        self.tail = _compile(_counter_tail(outer[0], "cnt"))
        env = {"np": np}
        exec(_compile([_func(tree, "nearest_interp"), _func(tree, "initialize_magcache")]), env)
        self.init = env["initialize_magcache"]

    def state(self, steps, thresh, K, R):
        s = types.SimpleNamespace()
        self.init(s, True, steps, thresh, K, R)
        return s

    def call(self, state):
        env = {"self": state, "np": np}
        exec(self.ctrl, env)
        skip = bool(env["skip_forward"])
        exec(self.tail, env)
        return skip


class RefControllerEvalWan:
    """eval/magcache/experiments/Wan2.1_EVAL/wan_magcache.py:770-786 (controller, `ratio[t-10]`) and :807-815 (counter)."""

    def __init__(self):
        fn = _func(_tree(EVAL_WAN), "magcache_forward")
        i0 = next(i for i, st in enumerate(fn.body) if isinstance(st, ast.Assign) and getattr(st.targets[0], "id", None) == "skip_time")
        ctrl = [fn.body[i0]] + [st for st in fn.body if isinstance(st, ast.If) and _mentions(st.test, "t") and isinstance(st.test, ast.Compare)
                                and getattr(st.test.comparators[0], "id", None) == "skip_time"]
        assert len(ctrl) == 2
        self.ctrl = _compile(ctrl)
        self.tail = _compile(_counter_tail(fn, "t"))
        self.table = np.asarray(_class_attr_literal(EVAL_WAN, "ratio"), dtype=np.float64)

    def state(self, steps, thresh, K, table=None):
        s = types.SimpleNamespace(t=0, num_steps=2 * steps, magcache_thresh=thresh, magcache_K=K, accumulated_err=[0, 0], accumulated_sim=[1, 1],
                                  accumulated_steps=[0, 0], skip_steps=0, residual_cache=np.zeros((2, 1, 1)),
                                  ratio=self.table if table is None else table)  # attribute block :1131-1150
        return s

    def call(self, state):
        env = {"self": state, "np": np, "skip_forward": False}
        exec(self.ctrl, env)
        skip = bool(env["skip_forward"])
        exec(self.tail, env)
        return skip


class RefControllerOpenSora:
    """eval/magcache/experiments/opensora.py:297-308 (controller: `ratio[t-1]`, error accumulated WITHOUT abs) and :348-354
    (counter, video length 30 hard-coded)."""

    def __init__(self):
        fn = _func(_tree(OPENSORA), "magcache_forward")
        ctrl = [st for st in ast.walk(fn) if isinstance(st, ast.If) and isinstance(st.test, ast.Compare) and _mentions(st.test, "skip_time")]
        assert len(ctrl) == 1
        self.ctrl = _compile(ctrl)
        self.tail = _compile(_counter_tail(fn, "t"))
        self.table = np.asarray(_class_attr_literal(OPENSORA, "ratio"), dtype=np.float64)

    def state(self, thresh, K, skip_time):
        return types.SimpleNamespace(t=0, magcache_thresh=thresh, K=K, skip_time=skip_time, accumulated_err=0, accumulated_sim=1,
                                     accumulated_steps=0, skip_steps=0, residual_cache=np.zeros((1, 3)), ratio=self.table)  # :419-433

    def call(self, state):
        env = {"self": state, "np": np, "skip_forward": False}
        exec(self.ctrl, env)
        skip = bool(env["skip_forward"])
        exec(self.tail, env)
        return skip


TEACACHE = f"{REF}/eval/magcache/experiments/Wan2.1_EVAL/wan_teacache.py"


class RefTeaCache:
    """The controller statements of teacache_forward (wan_teacache.py:533-564: the first `if self.enable_teacache:` block) and the
    counter (:587-589), executed on stand-in tensors; the coefficient literals of the t2v installation block (:913-926)."""

    def __init__(self):
        fn = _func(_tree(TEACACHE), "teacache_forward")
        blocks = [st for st in fn.body if isinstance(st, ast.If) and _mentions(st.test, "enable_teacache")]
        assert len(blocks) == 2  # controller, then the hit/miss branches
        self.ctrl = _compile([blocks[0]])
        self.tail = _compile(_counter_tail(fn, "cnt"))
        # coefficient literals: assignments `wan_t2v.model.__class__.coefficients = [...]` in order of appearance
        lits = []
        for n in ast.walk(_tree(TEACACHE)):
            if isinstance(n, ast.Assign) and isinstance(n.targets[0], ast.Attribute) and n.targets[0].attr == "coefficients":
                base = n.targets[0]
                while isinstance(base, ast.Attribute):
                    base = base.value
                if getattr(base, "id", "") == "wan_t2v":
                    lits.append((n.lineno, [float(v) for v in eval(compile(ast.Expression(n.value), "<c>", "eval"))]))
        lits.sort()
        assert len(lits) == 4, lits
        self.coefficients = {"ret_1.3B": lits[0][1], "ret_14B": lits[1][1], "noret_1.3B": lits[2][1], "noret_14B": lits[3][1]}

    def state(self, steps, thresh, key):
        use_ret = key.startswith("ret")
        return types.SimpleNamespace(enable_teacache=True, cnt=0, num_steps=2 * steps, teacache_thresh=thresh, accumulated_rel_l1_distance_even=0,
                                     accumulated_rel_l1_distance_odd=0, previous_e0_even=None, previous_e0_odd=None, use_ref_steps=use_ret,
                                     coefficients=self.coefficients[key], ret_steps=(10 * 2 if use_ret else 1 * 2),
                                     cutoff_steps=(2 * steps if use_ret else 2 * steps - 2))  # :901-928

    def call(self, state, e, e0):
        env = {"self": state, "np": np, "e": e, "e0": e0}
        exec(self.ctrl, env)
        calc = env["should_calc_even"] if state.is_even else env["should_calc_odd"]
        exec(self.tail, env)
        return bool(calc)


def teacache_cases():
    """Synthetic but realistic modulated inputs: sinusoidal timestep embedding (freq_dim 256) through a seeded 2-layer SiLU MLP and
    the 6-way projection, at the shift-5 flow timesteps; the cond and uncond calls of a step see the same t (wan_magcache.py:296-299)."""
    import torch
    ref = RefTeaCache()
    g = torch.Generator().manual_seed(0)
    D, fd = 192, 256
    W1, W2, W3 = torch.randn(D, fd, generator=g) * 0.05, torch.randn(D, D, generator=g) * 0.08, torch.randn(6 * D, D, generator=g) * 0.08
    b2, b3 = torch.randn(D, generator=g), torch.randn(6 * D, generator=g)  # trained models: the embedding moves a few % per step

    def embed(tval, amp):
        half = fd // 2
        pos = torch.tensor([tval], dtype=torch.float64)
        s_ = torch.outer(pos, torch.pow(10000, -torch.arange(half, dtype=torch.float64).div(half)))
        x = torch.cat([torch.cos(s_), torch.sin(s_)], dim=1).float()
        x[:, :64] = 0.0
        x[:, half:half + 64] = 0.0  # keep the slow frequencies only: neighbouring timesteps stay correlated
        e = amp * torch.nn.functional.silu(x @ W1.t()) @ W2.t() + b2
        e0 = (amp * torch.nn.functional.silu(e) @ W3.t() + b3).unflatten(1, (6, D))
        return e, e0

    out = {"coefficients": ref.coefficients, "cases": []}
    for key in ref.coefficients:
        for steps, amp in ((50, 0.6), (50, 2.0), (50, 6.0), (30, 2.0)):
            for thresh in (0.05, 0.1, 0.2, 0.3):
                st = ref.state(steps, thresh, key)
                s_lin = np.linspace(1.0, 1.0 / steps, steps)
                sig = 5.0 * s_lin / (1 + 4.0 * s_lin)
                dists, mask = [], []
                n_calls = 2 * steps * 2 + 3  # two videos and a bit: the accumulators are not reset at the wrap
                for c in range(n_calls):
                    e, e0 = embed(float(sig[(c // 2) % steps] * 1000.0), amp)
                    mod = e0 if st.use_ref_steps else e
                    prev = st.previous_e0_even if st.cnt % 2 == 0 else st.previous_e0_odd
                    consulted = not (st.cnt < st.ret_steps or st.cnt >= st.cutoff_steps)
                    dists.append(((mod - prev).abs().mean() / prev.abs().mean()).cpu().item() if consulted else None)
                    mask.append(1 if ref.call(st, e, e0) else 0)
                out["cases"].append({"key": key, "steps": steps, "amp": amp, "thresh": thresh, "ret_steps": st.ret_steps, "cutoff_steps": st.cutoff_steps,
                                     "rel_l1": dists, "calc": "".join(map(str, mask)), "computed_first_video": int(sum(mask[:2 * steps])),
                                     "final": {"cnt": int(st.cnt), "even": float(st.accumulated_rel_l1_distance_even),
                                               "odd": float(st.accumulated_rel_l1_distance_odd)}})
    return out


def remaining_tables():
    """Table literals not covered above: Qwen-Image / Qwen-Image-Edit (`mag_ratios = [...]` + [1.0]*2 prefix, :76 / :79) and
    FLUX-Kontext (magcache_flux_kontext.py:458)."""
    out = {}
    for key, path in (("qwen_image", QWEN), ("qwen_image_edit", f"{REF}/MagCache4QwenImageEdit/magcache_generate.py")):
        found = [n for n in ast.walk(_tree(path)) if isinstance(n, ast.Assign) and len(n.targets) == 1 and isinstance(n.targets[0], ast.Name)
                 and n.targets[0].id == "mag_ratios" and isinstance(n.value, ast.List) and len(n.value.elts) > 10]
        assert len(found) == 1, (path, len(found))
        vals = eval(compile(ast.Expression(found[0].value), "<tbl>", "eval"))
        out[key] = {"source": f"{os.path.relpath(path, REF)}:{found[0].lineno} (+ [1.0]*2 prefix in init_magcache)",
                    "values": [1.0, 1.0] + [float(v) for v in vals]}
    kpath = f"{REF}/MagCache4FLUX_Kontext/magcache_flux_kontext.py"
    out["flux_kontext"] = {"source": "MagCache4FLUX_Kontext/magcache_flux_kontext.py:458",
                           "values": [float(v) for v in _class_attr_literal(kpath, "mag_ratios")]}
    return out


def paper_eval_cases():
    out = {"tables": {}, "framepack_masks": [], "eval_wan_masks": [], "opensora_masks": []}
    for key, path in (("framepack", FRAMEPACK), ("framepack_f1", FRAMEPACK_F1)):
        ref = RefControllerFramePack(path)
        out["tables"][key] = {"source": f"{os.path.relpath(path, REF)}:70", "values": [float(v) for v in ref.state(25, 0.1, 3, 0.2).mag_ratios]}
        for steps in (25, 20, 30, 50, 12):
            for thresh, K, R in [(0.1, 3, 0.2), (0.1, 2, 0.2), (0.2, 5, 0.1), (0.3, 8, 0.0), (0.05, 3, 0.02), (0.6, 6, 0.3)]:
                st = ref.state(steps, thresh, K, R)
                n = 2 * steps + 3
                m = run_mask(ref, st, n)
                out["framepack_masks"].append({"table": key, "steps": steps, "thresh": thresh, "K": K, "R": R, "calls": n, "mask": "".join(map(str, m)),
                                               "ratios": [float(v) for v in st.mag_ratios],
                                               "final": {"cnt": int(st.cnt), "accumulated_err": float(st.accumulated_err),
                                                         "accumulated_ratio": float(st.accumulated_ratio), "accumulated_steps": int(st.accumulated_steps)}})
    ew = RefControllerEvalWan()
    out["tables"]["wan2.1_eval"] = {"source": "eval/magcache/experiments/Wan2.1_EVAL/wan_magcache.py:1144", "values": [float(v) for v in ew.table]}
    for thresh, K in [(0.12, 2), (0.12, 4), (0.24, 6), (0.06, 1), (0.03, 3), (0.5, 10)]:
        st = ew.state(50, thresh, K)
        n = 2 * 100 + 9
        m = run_mask(ew, st, n)
        out["eval_wan_masks"].append({"steps": 50, "thresh": thresh, "K": K, "calls": n, "mask": "".join(map(str, m)),
                                      "skipped_first_video": int(sum(m[:100])),
                                      "final": {"cnt": int(st.t), "accumulated_err": [float(v) for v in st.accumulated_err],
                                                "accumulated_ratio": [float(v) for v in st.accumulated_sim],
                                                "accumulated_steps": [int(v) for v in st.accumulated_steps]}})
    os_ = RefControllerOpenSora()
    out["tables"]["opensora_eval"] = {"source": "eval/magcache/experiments/opensora.py:433", "values": [float(v) for v in os_.table]}
    for thresh, K, skip_time in [(0.12, 3, 6), (0.24, 5, 6), (0.06, 2, 6), (0.12, 3, 3), (0.4, 8, 1), (0.01, 3, 10), (0.005, 29, 1)]:
        st = os_.state(thresh, K, skip_time)
        n = 2 * 30 + 4
        m = run_mask(os_, st, n)
        out["opensora_masks"].append({"steps": 30, "thresh": thresh, "K": K, "skip_time": skip_time, "calls": n, "mask": "".join(map(str, m)),
                                      "final": {"cnt": int(st.t), "accumulated_err": float(st.accumulated_err),
                                                "accumulated_ratio": float(st.accumulated_sim), "accumulated_steps": int(st.accumulated_steps)}})
    return out


def main():
    tables = extract_tables()
    with open(f"{OUT}/tables.json", "w") as f:
        json.dump(tables, f, indent=0)
    print("tables:", {k: len(v["values"]) for k, v in tables.items()})

    interp = reference_nearest_interp()
    ni = []
    for key, targets in [("wan2.1_t2v_1.3b", [1, 2, 3, 10, 20, 25, 30, 40, 49, 50, 51, 60, 100]), ("flux_dev", [1, 2, 8, 20, 27, 28, 29, 50]),
                         ("hunyuan_720p", [1, 30, 50, 64])]:
        src = np.array(tables[key]["values"])
        per_branch = key.startswith("wan")
        srcs = [("cond", src[0::2]), ("uncond", src[1::2])] if per_branch else [("all", src)]
        for tag, s in srcs:
            for T in targets:
                out = interp(s, T)
                ni.append({"table": key, "slice": tag, "L": int(len(s)), "T": T, "out": [float(v) for v in out]})
    # synthetic ties (round-half-to-even is visible when (L-1)/(T-1) lands on .5)
    for L, T in [(5, 3), (9, 5), (3, 5), (4, 7), (51, 21), (101, 41), (2, 2), (7, 1)]:
        s = np.arange(L, dtype=np.float64) * 1.5 + 0.25
        ni.append({"table": None, "slice": f"arange{L}", "L": L, "T": T, "src": [float(v) for v in s], "out": [float(v) for v in interp(s, T)]})
    with open(f"{OUT}/nearest_interp.json", "w") as f:
        json.dump(ni, f)
    print("nearest_interp cases:", len(ni))

    masks = []
    wan = RefController(WAN)
    for key in ["wan2.1_t2v_1.3b", "wan2.1_t2v_14b", "wan2.1_i2v_480p", "wan2.1_i2v_720p", "wan2.1_vace_1.3b", "wan2.1_vace_14b"]:
        for thresh, K, R in [(0.12, 2, 0.2), (0.12, 4, 0.2), (0.24, 6, 0.2), (0.02, 3, 0.2), (0.06, 1, 0.1), (0.5, 8, 0.3)]:
            for steps in [50, 40, 30, 20]:
                st = wan_state(tables[key]["values"], steps, thresh, K, R, interp)
                n = 2 * steps * 2 + 7  # two whole videos and a bit: exercises the counter reset (:306-311)
                m = run_mask(wan, st, n)
                masks.append({"family": "wan2.1", "table": key, "steps": steps, "thresh": thresh, "K": K, "R": R,
                              "calls": n, "mask": "".join(map(str, m)), "skipped_first_video": int(sum(m[:2 * steps])),
                              "final": final_state(st)})
    flux = RefController(FLUX)
    for thresh, K, R in [(0.24, 5, 0.1), (0.12, 3, 0.1), (0.05, 4, 0.2), (0.4, 8, 0.05)]:
        for steps in [28, 20, 50, 12]:
            st = scalar_state(tables["flux_dev"]["values"], steps, thresh, K, R, interp, "previous_residual")
            n = 2 * steps + 3
            m = run_mask(flux, st, n)
            masks.append({"family": "flux", "table": "flux_dev", "steps": steps, "thresh": thresh, "K": K, "R": R,
                          "calls": n, "mask": "".join(map(str, m)), "skipped_first_video": int(sum(m[:steps])), "final": final_state(st)})
    hun = RefController(HUN)
    for key in ["hunyuan_720p", "hunyuan_544p"]:
        for thresh, K, R in [(0.24, 6, 0.2), (0.12, 4, 0.2), (0.06, 2, 0.1)]:
            for steps in [50, 30, 25]:
                st = scalar_state(tables[key]["values"], steps, thresh, K, R, interp, "residual_cache")
                n = 2 * steps + 3
                m = run_mask(hun, st, n)
                masks.append({"family": "hunyuan", "table": key, "steps": steps, "thresh": thresh, "K": K, "R": R,
                              "calls": n, "mask": "".join(map(str, m)), "skipped_first_video": int(sum(m[:steps])), "final": final_state(st)})
    with open(f"{OUT}/masks.json", "w") as f:
        json.dump(masks, f, indent=0)
    print("mask cases:", len(masks))
    for c in masks:
        if c["table"] == "wan2.1_t2v_1.3b" and c["steps"] == 50 and c["R"] == 0.2 and c["thresh"] in (0.12, 0.24):
            print(c["thresh"], c["K"], c["skipped_first_video"], c["mask"][0:100:2])

    extra = extra_cases(tables, interp)
    with open(f"{OUT}/extra_adapters.json", "w") as f:
        json.dump(extra, f, indent=0)
    print("extra adapters:", {k: len(v) for k, v in extra.items()})

    pe = paper_eval_cases()
    with open(f"{OUT}/paper_eval_adapters.json", "w") as f:
        json.dump(pe, f, indent=0)
    print("framepack / paper-eval adapters:", {k: len(v) for k, v in pe.items()})
    # every calibrated table of the reference in one file (the package ships a copy: magcache_b200/tables.json)
    every = dict(tables)
    every.update(extra["tables"])
    every.update(pe["tables"])
    every.update(remaining_tables())
    with open(f"{OUT}/tables_all.json", "w") as f:
        json.dump(every, f, indent=0)
    print("all tables:", {k: len(v["values"]) for k, v in every.items()})
    for c in pe["eval_wan_masks"][:2] + pe["opensora_masks"][:2] + pe["framepack_masks"][:1]:
        print({k: v for k, v in c.items() if k in ("thresh", "K", "mask", "skipped_first_video")})

    tea = teacache_cases()
    with open(f"{OUT}/teacache.json", "w") as f:
        json.dump(tea, f, indent=0)
    print("teacache cases:", len(tea["cases"]), [(c["key"], c["thresh"], c["computed_first_video"]) for c in tea["cases"] if c["steps"] == 50][:8])

    cal = calib_cases()
    with open(f"{OUT}/calib_stats.json", "w") as f:
        json.dump(cal, f, indent=0)
    print("calibration cases:", cal)


if __name__ == "__main__":
    main()

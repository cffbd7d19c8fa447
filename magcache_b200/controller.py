"""Skip controller front end. The decision arithmetic lives in the C ABI (`mc_ctrl_decide` / `mc_ctrl_advance`, float64,
bit-exact with the reference); this module only moves state between the reference's attribute names and the C structs."""
import ctypes

import numpy as np

from . import _lib
from ._lib import CtrlConfig, CtrlState, check, lib


def make_ctrl_config(num_steps, thresh, K, retention_ratio, mag_ratios, branches, cmp, retention_mode, veto_index=-1, veto_base=0,
                     split_step=0, table_offset=0, min_cnt=0, flags=0, ratio_veto=None):
    arr = np.ascontiguousarray(mag_ratios, dtype=np.float64)
    if len(arr) < num_steps - table_offset:
        raise IndexError(f"mag_ratios has {len(arr)} entries but num_steps={num_steps}, table_offset={table_offset} "
                         "(interpolate first, magcache_generate.py:915-919)")
    cfg = CtrlConfig()
    cfg.num_steps, cfg.branches, cfg.K, cfg.cmp, cfg.retention_mode = int(num_steps), int(branches), int(K), int(cmp), int(retention_mode)
    cfg.veto_index, cfg.veto_base, cfg.split_step = int(veto_index), int(veto_base), int(split_step)
    if ratio_veto is not None:
        flags |= _lib.MC_CTRL_RATIO_VETO
    cfg.table_offset, cfg.min_cnt, cfg.flags, cfg.reserved = int(table_offset), int(min_cnt), int(flags), 0
    cfg.thresh, cfg.retention_ratio, cfg.ratio_veto = float(thresh), float(retention_ratio), float(ratio_veto or 0.0)
    cfg.mag_ratios = arr.ctypes.data_as(ctypes.POINTER(ctypes.c_double))
    cfg._keepalive = arr
    return cfg


def schedule_mask(cfg, calls):
    """Whole skip schedule (uint8 0/1 per forward call) from a fresh state."""
    mask = np.zeros(calls, dtype=np.uint8)
    check(lib.mc_ctrl_mask(ctypes.byref(cfg), calls, mask.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))))
    return mask


class AttrController:
    """Runs the controller on state stored under the reference's attribute names of `owner` (a model instance whose class
    carries `cnt`, `accumulated_ratio`, ... exactly as MagCache4Wan2.1/magcache_generate.py:897-906 installs them).
    Scalar families (FLUX/Hunyuan) keep scalars, Wan keeps 2-element lists — whatever form the script wrote is preserved."""

    def __init__(self, family_kwargs):
        self.kw = family_kwargs
        self._cfg = None
        self._key = None

    def _config(self, o):
        mr = o.mag_ratios
        kw = self.kw
        if kw["retention_mode"] in (_lib.MC_RETAIN_WAN22_T2V, _lib.MC_RETAIN_WAN22_I2V):
            # MagCache4Wan2.2/magcache_generate.py:344 `split_step = split_steps*2`; None (TI2V-5B) falls back to int(n*R), :301-303
            split = getattr(o, "split_step", None)
            kw = dict(kw, split_step=int(split)) if split is not None else dict(kw, retention_mode=_lib.MC_RETAIN_FLOOR)
        # keyed on the table's CONTENT (a few hundred bytes): an in-place edit of the installed table — the reference's suggested
        # `**0.5` smoothing, say — keeps id() and len() but must reach the controller
        arr = np.ascontiguousarray(np.asarray(mr, dtype=np.float64))
        key = (hash(arr.tobytes()), len(arr), o.num_steps, float(o.magcache_thresh), int(o.K), float(o.retention_ratio), kw.get("split_step"))
        if self._key != key:
            self._cfg = make_ctrl_config(o.num_steps, o.magcache_thresh, o.K, o.retention_ratio, arr, **kw)
            self._key = key
        return self._cfg

    def _load(self, o):
        st = CtrlState()
        st.cnt = int(o.cnt)
        if self.kw["branches"] == 2:
            for i in range(2):
                st.accumulated_ratio[i] = float(o.accumulated_ratio[i])
                st.accumulated_err[i] = float(o.accumulated_err[i])
                st.accumulated_steps[i] = int(o.accumulated_steps[i])
        else:
            # FramePack's `initialize_magcache` does not create the accumulators; its forward does at cnt == 0
            # (magcache_demo_gradio.py:63-74, :253-256) — absent attributes are the fresh state
            st.accumulated_ratio[0] = float(getattr(o, "accumulated_ratio", 1.0))
            st.accumulated_err[0] = float(getattr(o, "accumulated_err", 0.0))
            st.accumulated_steps[0] = int(getattr(o, "accumulated_steps", 0))
            st.accumulated_ratio[1] = 1.0
        return st

    def _store(self, o, st, with_cnt):
        cls = type(o)
        if self.kw["branches"] == 2:
            # the reference mutates the class-level lists in place (self.accumulated_ratio[i] = ...)
            for i in range(2):
                o.accumulated_ratio[i] = st.accumulated_ratio[i]
                o.accumulated_err[i] = st.accumulated_err[i]
                o.accumulated_steps[i] = st.accumulated_steps[i]
        else:
            o.accumulated_ratio, o.accumulated_err, o.accumulated_steps = st.accumulated_ratio[0], st.accumulated_err[0], st.accumulated_steps[0]
        if with_cnt:
            self._set_cnt(o, st.cnt)
        del cls

    @staticmethod
    def _set_cnt(o, value):
        """`self.cnt += 1`. Wan2.2 / Qwen-Image install `cnt = torch.tensor(0)` on the CLASS (MagCache4Wan2.2/magcache_generate.py:342): the
        in-place add mutates that one tensor, which is how the high-noise and the low-noise expert (two instances of one class) share a
        counter. Keep that: a tensor counter is updated in place, anything else is rebound on the instance like a Python int."""
        cur = getattr(type(o), "cnt", None)
        if hasattr(cur, "fill_") and "cnt" not in o.__dict__:
            cur.fill_(value)
        else:
            o.cnt = value

    def decide(self, o):
        cfg = self._config(o)
        st = self._load(o)
        skip = ctypes.c_int32(0)
        check(lib.mc_ctrl_decide(ctypes.byref(cfg), ctypes.byref(st), ctypes.byref(skip)))
        self._store(o, st, with_cnt=False)
        return bool(skip.value)

    def advance(self, o):
        cfg = self._config(o)
        st = self._load(o)
        check(lib.mc_ctrl_advance(ctypes.byref(cfg), ctypes.byref(st)))
        if st.cnt == 0 and self.kw["branches"] == 2 and not (self.kw.get("flags", 0) & _lib.MC_CTRL_WRAP_KEEPS_ACC):
            # end of video: the reference REBINDS fresh lists (magcache_generate.py:308-311)
            o.accumulated_ratio, o.accumulated_err, o.accumulated_steps = [1.0, 1.0], [0.0, 0.0], [0, 0]
            self._set_cnt(o, 0)
        else:
            self._store(o, st, with_cnt=True)

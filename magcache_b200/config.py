"""MagCache configuration: one parameterised description of every controller variant the reference ships
(SURVEY.md Appendix A), the calibrated `mag_ratios` tables, and the presets the reference's READMEs quote."""
import json
import os
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import _lib

_TABLES = None


def tables():
    """Calibrated magnitude-ratio tables (data copied from the reference literals; see `source` per entry):
    MagCache4Wan2.1/magcache_generate.py:910,912,1002,1004,1142,1144; MagCache4FLUX/magcache_flux.py:459;
    MagCache4HunyuanVideo/magcache_sample_video.py:316,318."""
    global _TABLES
    if _TABLES is None:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tables.json")) as f:
            _TABLES = {k: np.array(v["values"], dtype=np.float64) for k, v in json.load(f).items()}
    return _TABLES


def table_for_ckpt_dir(ckpt_dir: str, task: str = "t2v"):
    """The reference picks the table by substring of --ckpt_dir (magcache_generate.py:909-912, :1001-1004, :1141-1144)."""
    t = tables()
    if "VACE-1.3B" in ckpt_dir:
        return t["wan2.1_vace_1.3b"]
    if "VACE-14B" in ckpt_dir:
        return t["wan2.1_vace_14b"]
    if "T2V-14B" in ckpt_dir:
        return t["wan2.1_t2v_14b"]
    if "T2V-1.3B" in ckpt_dir:
        return t["wan2.1_t2v_1.3b"]
    if "480P" in ckpt_dir:
        return t["wan2.1_i2v_480p"]
    if "720P" in ckpt_dir:
        return t["wan2.1_i2v_720p"]
    raise KeyError(f"no calibrated mag_ratios table matches ckpt_dir={ckpt_dir!r} (the reference would hit AttributeError later)")


def nearest_interp(src, target_length):
    """C-ABI `mc_nearest_interp` (MagCache4Wan2.1/magcache_generate.py:27-34)."""
    import ctypes
    src = np.ascontiguousarray(src, dtype=np.float64)
    out = np.empty(target_length, dtype=np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    _lib.check(_lib.lib.mc_nearest_interp(src.ctypes.data_as(dp), len(src), out.ctypes.data_as(dp), target_length))
    return out


def interp_cfg(table, sample_steps):
    """Per-CFG-branch interpolation (magcache_generate.py:915-919) through `mc_nearest_interp_cfg`."""
    import ctypes
    table = np.ascontiguousarray(table, dtype=np.float64)
    if len(table) == 2 * sample_steps:
        return table
    out = np.empty(2 * sample_steps, dtype=np.float64)
    dp = ctypes.POINTER(ctypes.c_double)
    _lib.check(_lib.lib.mc_nearest_interp_cfg(table.ctypes.data_as(dp), len(table), out.ctypes.data_as(dp), sample_steps))
    return out


@dataclass
class MagCacheConfig:
    """family: 'wan2.1' (per-branch state, `<`, int(n*R)), 'flux' (scalar, `<=`, int(R*n+0.5), step-11 veto),
    'hunyuan' (scalar, `<=`, int(R*n))."""
    family: str = "wan2.1"
    thresh: float = 0.12
    K: int = 2
    retention_ratio: float = 0.2
    sample_steps: int = 50
    mag_ratios: Optional[Sequence[float]] = None
    table: Optional[str] = None  # key into tables() when mag_ratios is not given

    @property
    def branches(self):
        return 2 if self.family == "wan2.1" else 1

    @property
    def num_steps(self):  # forward calls per video
        return self.sample_steps * self.branches

    def resolved_ratios(self):
        src = np.asarray(self.mag_ratios if self.mag_ratios is not None else tables()[self.table], dtype=np.float64)
        if self.family == "wan2.1":
            return interp_cfg(src, self.sample_steps)
        return src if len(src) == self.sample_steps else nearest_interp(src, self.sample_steps)

    def ctrl_kwargs(self):
        fam = self.family
        return dict(branches=self.branches, cmp=_lib.MC_CMP_LT if fam == "wan2.1" else _lib.MC_CMP_LE,
                    retention_mode=_lib.MC_RETAIN_HALF_UP if fam == "flux" else _lib.MC_RETAIN_FLOOR,
                    veto_index=11 if fam == "flux" else -1, veto_base=28 if fam == "flux" else 0)


PRESETS = {
    # MagCache4Wan2.1/README.md:13,19 ; naming E<thresh>K<K>R<retention>
    "wan2.1-1.3b-E012K2R02": MagCacheConfig("wan2.1", 0.12, 2, 0.2, 50, table="wan2.1_t2v_1.3b"),
    "wan2.1-1.3b-E012K4R02": MagCacheConfig("wan2.1", 0.12, 4, 0.2, 50, table="wan2.1_t2v_1.3b"),
    "wan2.1-14b-E024K6R02": MagCacheConfig("wan2.1", 0.24, 6, 0.2, 50, table="wan2.1_t2v_14b"),
    # MagCache4FLUX/magcache_flux.py:466-468
    "flux-E024K5R01": MagCacheConfig("flux", 0.24, 5, 0.1, 28, table="flux_dev"),
    # MagCache4HunyuanVideo/magcache_sample_video.py:303-305
    "hunyuan-720p-E024K6R02": MagCacheConfig("hunyuan", 0.24, 6, 0.2, 50, table="hunyuan_720p"),
    "hunyuan-720p-E012K4R02": MagCacheConfig("hunyuan", 0.12, 4, 0.2, 50, table="hunyuan_720p"),
}

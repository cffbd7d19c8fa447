"""The C ABI from C: `examples/host_loop.c` — a plain-C host of the denoising loop (mc_ctrl_* + mc_dit_*) — must compile as C99 against
include/magcache_b200.h with warnings as errors (the header carries no C++ and no torch types), link against the in-tree library, and
its host-only part must run without a GPU and agree with the Python layer on the skip schedule and the launch-plan sizes."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import magcache_b200 as mc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("gcc") is None, reason="gcc not available")
def test_c_host_compiles_links_and_plans(tmp_path):
    exe = str(tmp_path / "host_loop")
    libdir = os.path.join(ROOT, "magcache_b200")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "examples", "host_loop.c"),
           "-L", libdir, "-lmagcache_b200", f"-Wl,-rpath,{libdir}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "--plan"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stderr + r.stdout
    out = dict(line.split(" ", 1) for line in r.stdout.strip().splitlines())
    assert int(out["abi"]) == mc._lib.ABI_VERSION
    table = np.array([1.0, 1.0] + [0.97] * 18)
    want = mc.MagCacheConfig("wan2.1", 0.12, 2, 0.2, 10, mag_ratios=table).schedule().tolist()
    assert out["skip"] == "".join(str(int(v)) for v in want) and 0 < sum(want) < 20
    assert int(out["workspace"]) % 1024 == 0
    plans = [line for line in r.stdout.splitlines() if line.startswith("plan ")]
    assert plans[0].startswith("plan skip=0 lines=44 first=patchify latent+0") and plans[1].startswith("plan skip=1 lines=8 first=patchify latent+0")
    assert subprocess.run([exe], capture_output=True).returncode == 2  # usage

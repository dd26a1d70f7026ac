"""Memory-safety evidence without compute-sanitizer (closed on the GPU pool): libslamcu_dbg.so is the library built with
-DSLAMCU_DEBUG_BOUNDS, in which every atomically indexed list write, tile offset and patch gather is range-checked on the device
and traps; tools/debug_bounds_cases.py drives the edge cases (noise, flat and tiny images in both modes, list overflow in
sequences, ragged matcher sizes through every slice count) through it.  A trap would surface as a CUDA error."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_edge_cases_do_not_trip_the_bounds_checks():
    so = os.path.join(ROOT, "slam_cin0051_b200", "libslamcu_dbg.so")
    assert os.path.exists(so), "run __graft_entry__.build() first (slam_cin0051_b200.build.build_debug)"
    env = dict(os.environ, SLAMCU_LIB=so)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "debug_bounds_cases.py")], capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert r.stdout.strip().splitlines()[-1].startswith("ok ")


def test_the_checks_are_live():
    """The debug build really checks: an index past a list's end (forced through the test hook) traps."""
    so = os.path.join(ROOT, "slam_cin0051_b200", "libslamcu_dbg.so")
    code = ("import os, sys; sys.path.insert(0, %r); import numpy as np; import slam_cin0051_b200 as S;"
            "ctx = S.Context(0); rc = ctx.lib.slamcu_debug_trip_bound(ctx.handle); ctx.synchronize()" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=dict(os.environ, SLAMCU_LIB=so))
    assert r.returncode != 0 and ("unspecified launch failure" in r.stderr or "trap" in r.stderr.lower() or "CUDA" in r.stderr), r.stderr[-1500:]
    # the release build compiles the checks out: the same call is a no-op
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1500:]

"""bench.py on a box without a GPU: the workload definitions build their meshes from plain numpy arrays, the reference
arm (the reference's own compiled compute_residual, or the oracle port) runs every workload it can and keeps the product
library out of its process, and the numpy Hilbert ordering is the library's."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from common import ROOT
from fvens_b200 import lib, synth


def _ref_line(*extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--cells", "2e4", *extra], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])


@pytest.mark.parametrize("extra", [(), ("--numerics", "hllc-gg-bj"), ("--workload", "viscous"),
                                   ("--workload", "ogrid-weno", "--flux", "ausm", "--weno-lambda", "20")])
def test_reference_arm_runs_the_workload_in_a_clean_process(extra):
    d = _ref_line(*extra)
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "Gfaces/s" and d["dtype"] == "f64"
    assert d["config"]["cells"] > 1.5e4 and d["config"]["faces"] > d["config"]["cells"]
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert all("libfvens_b200" not in m for m in d["repo_libraries_mapped"]) and d["repo_libraries_mapped"]


def test_reference_arm_says_why_it_cannot_run_the_periodic_workload():
    d = _ref_line("--workload", "vortex", "--vortex-n", "32")
    assert d["impl"] == "reference" and "periodic" in d["unavailable"]


def test_reference_arm_prints_on_rank_zero_only():
    """Under torchrun the driver launches N processes: rank 0 alone runs the CPU reference and prints the line."""
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1", "--cells", "2e4"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and not [ln for ln in r.stdout.splitlines() if ln.startswith("{")]


def test_numpy_hilbert_order_is_the_library_s():
    for arrs in (synth.bump_channel(90, 34), synth.ogrid_cylinder(96, 40), synth.periodic_square(48, tri_fraction=0.3, jitter=0.1)):
        um = lib.UMesh.from_arrays(*arrs)
        assert np.array_equal(um.hilbert_ordering(), synth.hilbert_order(synth.cell_centres(*arrs[:3])))


def test_workloads_build_consistent_meshes():
    sys.path.insert(0, ROOT)
    import argparse
    import bench
    for w, kw in (("bump", {}), ("viscous", {}), ("ogrid-weno", dict(flux="hllc", weno_lambda=1.0)), ("vortex", dict(n=40, scaling="strong"))):
        ns = argparse.Namespace(workload=w, numerics="roe-wls-venkat", cells=1.5e4, cell_order="hilbert", **kw)
        wl = bench.Workload(ns)
        arrs, u, lat = wl.arrays()
        um = wl.host_mesh(arrs)
        assert um.nelem == len(u) and np.isfinite(u).all() and (u[:, 0] > 0).all()
        a, b = wl.algorithmic_bytes(um.nelem, um.naface)
        assert a > 0 and b == 160*um.nelem + 48*um.naface
        markers = set(um.arrays()["btags"][:, 0].tolist())
        assert markers == {t for (t, _, _) in wl.bcs}
        if w == "vortex":
            assert (um.periodic_partners() >= 0).all()

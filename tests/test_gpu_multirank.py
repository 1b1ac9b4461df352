"""The multi-rank path on ONE GPU: several processes share cuda:0, each with its subdomain mesh; halo rows travel through
the peer-memory windows (CUDA IPC works between processes on the same device), gloo carries the set-up all-gather.
tests/mgpu_check.py compares residuals, forward-Euler steps and the multi-GPU solver with the single-mesh engine
(bitwise states and residuals; the norm history to round-off, its summation order differs)."""
import os
import subprocess
import sys
import pytest
from common import ROOT

pytestmark = pytest.mark.gpu


def _run(world, port, **env):
    e = dict(os.environ, MGPU_SAME_DEVICE="1", MASTER_ADDR="127.0.0.1", **env)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=900, env=e)
    print(r.stdout[-4000:], r.stderr[-3000:])
    return r


@pytest.mark.parametrize("world,graph,pdl,partition", [(2, "1", "1", "sfc"), (3, "1", "1", "rcb"), (3, "0", "0", "sfc"), (4, "1", "0", "sfc")])
def test_fused_engine_ranks_on_one_gpu(world, graph, pdl, partition):
    """fvg_dist_*: rows pushed by the producing kernels, waits inside the consuming kernels, evaluation replayed from a CUDA
    graph (graph = "1") with programmatic dependent launch (pdl = "1"): Roe/WLS/Venkatakrishnan, HLLC/GG/WENO, AUSM/MUSCL,
    first order, laminar viscous with and without a limiter; residuals, five steps and the 40-step multi-GPU solver."""
    r = _run(world, 29700 + world + 10*int(graph) + 20*int(pdl), FVG_DIST="fused", FVG_GRAPH=graph, FVG_PDL=pdl,
             MGPU_PARTITION=partition)
    assert r.returncode == 0 and f"MGPU_CHECK OK world {world}" in r.stdout


def test_fused_engine_with_a_one_entry_graph_cache():
    """FVG_GRAPH_CACHE=1: every evaluation with other arrays than the previous one evicts the cached CUDA graph (the cache
    is bounded, least recently used first, for callers that hand in new arrays all the time) - same bits as ever."""
    r = _run(2, 29790, FVG_DIST="fused", FVG_GRAPH="1", FVG_GRAPH_CACHE="1")
    assert r.returncode == 0 and "MGPU_CHECK OK world 2" in r.stdout


def test_a_rank_that_withholds_its_rows_is_reported():
    """One rank skips an evaluation: its neighbours' waits time out and fvg_dist_status returns FVG_ERR_COMM (ADVICE round 1:
    a stalled peer must not yield a silently wrong residual)."""
    r = _run(3, 29750, FVG_DIST="fused", MGPU_WITHHOLD="1", FVG_HALO_TIMEOUT_MS="1500")
    assert r.returncode == 0 and "MGPU_CHECK OK withhold world 3" in r.stdout


@pytest.mark.parametrize("world,overlap,fused", [(2, "0", "1"), (3, "0", "0"), (3, "1", "0")])
def test_split_schedule_ranks_on_one_gpu(world, overlap, fused):
    """The round-1 schedule driven from Python (FVG_DIST=split), kept for A/B comparisons: fused = "1": send kernels +
    in-kernel receive; overlap = "1": exchange kernels on a second stream behind the interior tiles; both "0": one
    exchange kernel before each pass."""
    r = _run(world, 29600 + world + 10*int(overlap) + 20*int(fused), FVG_DIST="split", FVG_OVERLAP=overlap, FVG_FUSED_RECV=fused)
    assert r.returncode == 0 and f"MGPU_CHECK OK world {world}" in r.stdout

"""The multi-rank path on ONE GPU: three processes share cuda:0, each with its subdomain mesh; halo rows travel through
the peer-memory windows (CUDA IPC works between processes on the same device), gloo carries the set-up all-gather.
tests/mgpu_check.py compares residuals and five forward-Euler steps with the single-mesh engine (bitwise)."""
import os
import subprocess
import sys
import pytest
from common import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world,overlap,fused", [(2, "0", "1"), (3, "0", "1"), (2, "0", "0"), (3, "1", "0")])
def test_peer_halo_ranks_on_one_gpu(world, overlap, fused):
    # fused = "1": in-kernel receive (send kernel only; the passes wait on the arrival flags and read the window);
    # overlap = "1": exchange kernels on a second stream behind the interior tiles (fvg_flow_select_tiles);
    # both "0": one exchange kernel before each pass
    env = dict(os.environ, MGPU_SAME_DEVICE="1", MASTER_ADDR="127.0.0.1", FVG_OVERLAP=overlap, FVG_FUSED_RECV=fused)
    port = 29600 + world + 10*int(overlap) + 20*int(fused)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py")],
                       capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0 and f"MGPU_CHECK OK world {world}" in r.stdout

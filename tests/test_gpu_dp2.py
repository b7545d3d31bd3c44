"""GPU, world size 2: the frame-parallel gradient exchange over REAL process boundaries (one process per GPU, NCCL for the
rendezvous and the dense reference, CUDA IPC + NVLink peer loads for dp.PeerExchange).  Skipped on a box with fewer than
two GPUs (the driver's single-GPU run); run it with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp2.py -m gpu`.

Each rank renders its own pose of the same seeded scene; after the exchange every rank must hold the sum a dense
all-reduce of the 13 P-float bucket produces -- for dp.PeerExchange (fused pack + pull over peer memory, no host
synchronisation) and for dp.SparseExchange (all-gather of the packed rows) -- over several steps, so that both slots of the
peer double buffer are used."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
import numpy as np
import torch
import torch.distributed as dist
ROOT = sys.argv[1]
sys.path.insert(0, os.path.join(ROOT, "lidar-gs_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
from lgs_b200 import capi, dp, synth
import util
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
L = capi.load()
sc = synth.make_scene(P=40000, H=32, W=512, seed=21, pose="random")
sc.update(synth.make_upstream(32, 512, seed=21))
P = sc["P"]
d = util.to_torch(sc, dev)
peer = dp.PeerExchange(P, dev)
sparse = dp.SparseExchange(P, dev)
keys = ("means3D", "scales", "rotations", "opacities", "colors")
for step in range(4):
    view = d["viewmatrix"].clone()
    view[3, 0] += 0.4 * rank + 0.1 * step          # every rank (and step) its own sensor pose
    fr = capi.Frame(dev)
    fr.forward(d["bg"], d["means3D"], d["colors"], d["opacities"], d["scales"], d["rotations"], view, d["beams"], 32, 512, 80, 0)
    b = dp.GradBucket(P, dev)
    grads = dict({k: b.views[k] for k in keys}, means2D=torch.empty((P, 4), device=dev), cov3D=None,
                 scratch=torch.empty(L.lgs_backward_scratch_bytes(P), dtype=torch.uint8, device=dev))
    fr.backward(d["g_color"], d["g_depth"], d["g_occ"], grads=grads)
    local_copy = b.flat.clone()
    want = b.flat.clone()
    dist.all_reduce(want)                            # the dense reference
    peer.exchange(grads["scratch"], b.flat, {k: b.views[k] for k in keys})
    torch.cuda.synchronize()
    scale = float(want.abs().max())
    err = float((b.flat - want).abs().max()) / scale
    assert err < 1e-5, ("peer", rank, step, err)
    assert float((want - local_copy).abs().max()) > 0  # the other rank really contributed something
    b.flat.copy_(local_copy)
    sparse.exchange(grads["scratch"], b.flat, {k: b.views[k] for k in keys})
    torch.cuda.synchronize()
    err = float((b.flat - want).abs().max()) / scale
    assert err < 1e-5, ("sparse", rank, step, err, sparse.last)
rows = peer.status()
assert 0 < rows < P
peer.close()
dist.barrier()
print(f"rank {rank}: OK, peer exchange max rows {rows}, sparse mode {sparse.last['mode']}", flush=True)
dist.destroy_process_group()
'''


def test_world2_peer_and_sparse_exchange_equal_dense_allreduce(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    script = tmp_path / "dp2_worker.py"
    script.write_text(WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", str(script), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("OK, peer exchange") == 2, r.stdout[-2000:]

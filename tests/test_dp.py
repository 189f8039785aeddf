"""Data-parallel host logic.  CPU: world_size-2 gloo run of the pieces that do not need a GPU (shard
filter of the global permutation, all-reduce plumbing).  GPU: 1-vs-2-GPU equivalence via torchrun."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

GLOO_WORKER = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(ROOT, "diamond-ppo_b200")); sys.path.insert(0, ROOT)
from diamond.agents import _PermWorker, _Dist
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
d = _Dist(None, True)
assert d.enabled and d.world == world and d.rank == rank
T, NL, MB, E = 8, 6, 4, 3
NG, B = NL * world, T * NL * world
M = B // MB
np.random.seed(5)
outs = [np.empty(T * NL, np.int32) for _ in range(E)]
w = _PermWorker(B, E, MB, outs, (NG, rank * NL, NL)); w.start()
for e in range(E): w.wait(e)
w.finish()
np.random.seed(5)
ref = [np.random.permutation(B) for _ in range(E)]
after = np.random.randint(0, 2**31, 2)
for e in range(E):
    seen = []
    for k, (off, m) in enumerate(w.counts[e]):
        loc = outs[e][off:off + m]
        t, env = np.divmod(loc, NL)
        glob = t * NG + env + rank * NL                     # local flat index -> global flat index
        seg = ref[e][k * M:(k + 1) * M]
        mine = seg[(seg % NG >= rank * NL) & (seg % NG < (rank + 1) * NL)]
        assert np.array_equal(glob, mine), (e, k)           # same members, same order as the global minibatch
        seen.append(loc)
        cnt = torch.tensor([m]); dist.all_reduce(cnt); assert int(cnt) == M     # shards partition every minibatch
    assert sorted(np.concatenate(seen).tolist()) == list(range(T * NL))
x = torch.full((5,), float(rank + 1), dtype=torch.float64); d.all_reduce_sum(x)
assert torch.equal(x, torch.full((5,), float(sum(range(1, world + 1))), dtype=torch.float64))
dist.destroy_process_group()
print("ok", rank)
'''


def test_dp_host_logic_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(f"ROOT = {ROOT!r}\n" + GLOO_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


@pytest.mark.gpu
def test_dp_two_gpus_equal_single_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tests", "dp_equivalence.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "OK" in r.stdout

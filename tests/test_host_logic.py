"""Host-side logic on CPU: subject sharding, the world_size-2 metric gather over gloo, the
metric formulas, the phantom generator."""
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT
from vent_analysis_b200 import cohort
from vent_analysis_b200.phantom import CONFIGS, cohort_member, phantom


def test_shard_indices_partition_the_cohort():
    for n in (1, 7, 1024):
        for world in (1, 2, 4, 8):
            parts = [cohort.shard_indices(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_phantom_is_seeded_and_plausible():
    a, m = phantom((64, 64, 8), [1.5, 1.5, 10.0], 3)
    b, _ = phantom((64, 64, 8), [1.5, 1.5, 10.0], 3)
    assert np.array_equal(a, b) and a.dtype == np.float32 and m.dtype == np.uint8
    assert 0.05 < m.mean() < 0.4 and a.min() >= 0
    img, mask, vox = cohort_member(1)
    assert img.shape == CONFIGS["C2"][0] and vox == CONFIGS["C2"][1]


_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np, torch.distributed as dist
from vent_analysis_b200 import cohort
rank, world, n = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), 11
dist.init_process_group("gloo", rank=rank, world_size=world)
idx = cohort.shard_indices(n, rank, world)
local = np.array([[i + 0.1 * k for k in range(6)] for i in idx], dtype=np.float64)
full = cohort.gather_metrics(local, n, rank, world)
expect = np.array([[i + 0.1 * k for k in range(6)] for i in range(n)])
assert np.array_equal(full, expect), full
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
'''


def test_gather_metrics_world_size_2_gloo(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29731", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)),
                              stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "ok" in o

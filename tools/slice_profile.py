"""Development tool: one rank's share of an N-GPU step on ONE GPU (mbpol_set_partition(rank, world) without a
communicator), for an ncu launch list of the sliced path:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/slice_profile.py 3 8
"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
from conftest import make_context  # noqa: E402
from mbpol_openmm_plugin_b200.systems import WORKLOADS, synthetic_box  # noqa: E402

rank, world = int(sys.argv[1]), int(sys.argv[2])
pos, box = synthetic_box(WORKLOADS[sys.argv[3] if len(sys.argv) > 3 else "synthetic_262144"])
n = len(pos)
ctx = make_context(n, box, True, rank=rank, world=world)
p = torch.as_tensor(pos.reshape(-1, 3)).cuda()
f = torch.zeros_like(p)
for step in range(4):            # evaluation 0: full build + partition; later ones: sliced builds
    torch.cuda.nvtx.range_push(f"step{step}")
    e = ctx.engine.eval_device(p, f, 3)
    torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
st = ctx.engine.stats()
print(e, st["n_pairs_owned"], st["n_triplets_owned"], st["lists_sliced"], st["n_launches"])

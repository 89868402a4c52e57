#!/usr/bin/env python
"""A small scene for compute-sanitizer (memcheck / racecheck): 4 envs, the 72-brick lattice dropped into the bin with a small jitter and
stepped with random actions while the heap is still tumbling -- edge-edge contacts, the queues of passes 1b / 2e, the bit rows of the broad
phase, wake-ups and candidate-list rebuilds all occur.  Prints the contact statistics it saw so that the log shows what was covered."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seqdex_b200.env import SdxEnv                      # noqa: E402
from seqdex_b200.scene import Scene                     # noqa: E402
from tests.util import lattice_bank                     # noqa: E402
from seqdex_b200.tasks.block_assembly_grasp_sim import default_tvalue_weights   # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 24
s = Scene()
g = SdxEnv(s, 4)
g.set_tvalue_weights(default_tvalue_weights(1))
g.set_heap_bank(lattice_bank(s, 2))
gen = torch.Generator(device="cuda").manual_seed(0)
con = g.tensor("CONTACTS")
edge, mx = 0, 0
for t in range(steps):
    g.step(torch.rand(4, 23, device="cuda", generator=gen) * 2 - 1)
    nc = g.tensor("NCONTACT")[:, 0]
    mx = max(mx, int(nc.max()))
    words = con[..., 0].contiguous().view(torch.int32)
    live = torch.arange(1024, device="cuda")[None, :] < nc[:, None]
    edge += int(((words & (1 << 27)) != 0)[live].sum())
torch.cuda.synchronize()
print(f"sanitizer scene: {steps} steps x 4 envs, most contacts in an env {mx}, edge-edge contacts seen (last sub-step of each step) {edge}")

"""Driven under ncu by tools/gpu_trip3.sh: the bench world, then a few launches of the gather probes, K2 and get_head."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pos_evolution_b200 import spec as PS  # noqa: E402
from pos_evolution_b200.engine import Engine  # noqa: E402

eng = Engine(0)
W = bench.build_world(eng, 0, np, PS)
dev = torch.device("cuda", 0)
d_m = torch.as_tensor(W["members"].astype(np.int32), device=dev)
d_o = torch.as_tensor(W["off"].astype(np.int32), device=dev)
d_b = torch.full((bench.N_AGG, 64), 0xFF, dtype=torch.uint8, device=dev)
chk = torch.zeros(bench.N_AGG, dtype=torch.int32, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for tma in (True, False):
    flush.zero_()
    eng.gather_probe_dev(d_m, d_o, d_b, chk, tma=tma)
    torch.cuda.synchronize()
out, st = eng.g1_aggregate(W["members"], W["off"], np.full((bench.N_AGG, 64), 0xFF, dtype=np.uint8))
ok = eng.fast_aggregate_verify(W["members"][:64 * 512], W["off"][:65], np.full((64, 64), 0xFF, dtype=np.uint8), W["msgs"][:64],
                               eng.aggregate(W["sigs"][:64 * 512], W["off"][:65])[0])
assert ok.all()
for _ in range(3):
    eng.get_head(0, bench.N_BLOCKS - 1, W["boost"])
print("done")

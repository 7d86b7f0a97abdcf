"""Phase clocks of get_head on the bench world (profiling aid; tools/gpu_trip4.sh): SM-clock stamps of the vote scatter of CTA 0 and
of the tree phases, for the one-launch form and (B2_HEAD_FUSED=0) the two-launch form, plus the host-side latency distribution."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from pos_evolution_b200 import spec as PS  # noqa: E402
from pos_evolution_b200.engine import Engine  # noqa: E402

eng = Engine(0)
W = bench.build_world(eng, 0, np, PS)
for _ in range(20):
    h = eng.get_head(0, bench.N_BLOCKS - 1, W["boost"])
lat = []
for _ in range(300):
    t0 = time.perf_counter()
    eng.get_head(0, bench.N_BLOCKS - 1, W["boost"])
    lat.append((time.perf_counter() - t0) * 1e6)
lat.sort()
c = eng.debug_head_clocks().astype(np.int64)
tree = [int(c[i + 1] - c[i]) for i in range(7)] + [int(c[15] - c[7])]
votes = [int(c[17] - c[16]), int(c[18] - c[17]), int(c[19] - c[18])]
print(json.dumps({"fused": os.environ.get("B2_HEAD_FUSED", "1"), "head": h, "p50_us": lat[150], "p10_us": lat[30], "p99_us": lat[296],
                  "tree_phase_clocks": dict(zip(["stage", "scan1", "weights", "store", "mark", "scan2", "find", "publish"], tree)),
                  "mark_list_build_clocks": int(c[8] - c[4]) if c[8] else None,
                  "tree_total_clocks": int(c[15] - c[0]), "votes_cta0_clocks": dict(zip(["zero_bins", "scatter", "flush"], votes)),
                  "votes_start_to_tree_end_clocks": int(c[15] - c[16])}))

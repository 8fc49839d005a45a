"""BASELINE config 4 timing of esr_cem_project exactly as bench.py measures it (six buffer sets in rotation, 60 calls per
event pair; and one call after a 256 MiB write flush).  ESR_CEM_FUSED=0 selects the two-launch streaming kernels.
  python tools/cem_time.py            -> prints one JSON line"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
r = bench.cem_standalone(dev, bench.peaks())
r["ESR_CEM_FUSED"] = os.environ.get("ESR_CEM_FUSED", "1")
print(json.dumps({k: r[k] for k in ("us", "frac", "after_write_flush_us", "frac_after_write_flush", "ESR_CEM_FUSED")}))

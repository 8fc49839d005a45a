"""Times the density kernels of the histogram / dictionary objectives (csrc/zobj.cu) against the reference's formulation
(materialised [D, N, M] fp64 tensors + autograd, codes/Z_optimization.py:184-195) run as torch ops on the same GPU.
GPU box only:  python tools/kde_bench.py [--once] > gpurun_out/kde_bench.json"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esr_b200 import z_objectives as zo  # noqa: E402

CASES = [  # name, D, N, M, temperature, per_bin, spread
    ("gray histogram, 512x512 region, 256 bins", 1, 512 * 512, 256, 5e-4, True, 1.0),
    ("gray dictionary, 512x512 region, 256 atoms", 1, 512 * 512, 256, 1e-3, False, 1.0),
    ("6x6 patch histogram, 3600 patches x 4000 atoms", 36, 3600, 4000, 5e-4, True, 0.08),
    ("6x6 patch dictionary, 3600 patches x 4000 atoms", 36, 3600, 4000, 1e-3, False, 0.08),
]


def materialised(samples, bins, temperature, per_bin):
    image = samples.double().unsqueeze(-1)
    b = bins.unsqueeze(1)
    dist = (image - b).abs()
    dist = torch.min(dist, (image - b - 1.0).abs())
    dist = torch.min(dist, (image - b + 1.0).abs())
    return torch.exp((-((dist + 1e-7) ** 2) / temperature).mean(0)).sum(0 if per_bin else 1)


def timed(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.reset_peak_memory_stats()
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, torch.cuda.max_memory_allocated() / 2 ** 20


def main():
    once = "--once" in sys.argv
    dev = torch.device("cuda", 0)
    out = []
    for name, D, N, M, T, per_bin, spread in CASES:
        rng = np.random.default_rng(D + N)
        base = rng.random((D, 1))
        samples = torch.from_numpy(((base + spread * rng.random((D, N))) % 1.0).astype(np.float32)).to(dev)
        bins = torch.from_numpy((base + spread * rng.random((D, M))) % 1.0).to(dev)
        weight = torch.from_numpy(rng.random(M if per_bin else N)).to(dev)

        def step(op):
            x = samples.clone().requires_grad_(True)
            (op(x) * weight).sum().backward()
            return x.grad

        mine = lambda: step(lambda x: zo.kde_sums(x, bins, 1.0, T, per_bin=per_bin))          # noqa: E731
        ref = lambda: step(lambda x: materialised(x, bins, T, per_bin))                       # noqa: E731
        g_mine, g_ref = mine(), ref()
        err = float((g_mine - g_ref).abs().max() / g_ref.abs().max())
        if once:
            continue
        ms, mib = timed(mine, 10)
        ms_ref, mib_ref = timed(ref, 3)
        out.append({"case": name, "D": D, "N": N, "M": M, "fwd_bwd_ms": ms, "peak_MiB": mib,
                    "torch_materialised_fwd_bwd_ms": ms_ref, "torch_materialised_peak_MiB": mib_ref,
                    "speedup": ms_ref / ms, "grad_rel_err": err,
                    "pair_terms_per_s": 2 * D * N * M / (ms * 1e-3)})
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

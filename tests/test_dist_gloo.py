"""world_size-2 gloo test of the batch sharding used for N > 1 GPUs (CPU; the per-shard function is the
oracle on a tiny net, standing in for the CUDA forward)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from esr_b200 import synth
from esr_b200.parallel import run_sharded, shard_range
from oracle.cem_ops import concat_latent
from oracle.rrdbnet import GCEMOracle


def test_shard_ranges_cover_batch():
    for n in (1, 2, 5, 16):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    wts = synth.make_weights("kaiming", seed=2, nb=1)
    lr, z = synth.make_inputs(3, 8, 8, seed=2)
    net = GCEMOracle(wts, nb=1)
    with torch.no_grad():
        out = run_sharded(net.forward, concat_latent(lr, z))
    if rank == 0:
        torch.save(out, os.path.join(out_dir, "gathered.pt"))
    dist.destroy_process_group()


def test_two_rank_sharding_is_bit_identical(tmp_path):
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    gathered = torch.load(os.path.join(str(tmp_path), "gathered.pt"))
    torch.set_num_threads(2)
    wts = synth.make_weights("kaiming", seed=2, nb=1)
    lr, z = synth.make_inputs(3, 8, 8, seed=2)
    net = GCEMOracle(wts, nb=1)
    with torch.no_grad():
        parts = [net.forward(concat_latent(lr[i:j], z[i:j])) for (i, j) in (shard_range(3, 0, 2), shard_range(3, 1, 2))]
    assert torch.equal(gathered, torch.cat(parts, 0))
    assert gathered.shape == (3, 3, 32, 32)

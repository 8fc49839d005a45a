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


# ------------------------------------------------------------------ tile sharding of one image (parallel.run_tiled)
def test_tile_windows_cover_image_with_halo():
    from esr_b200.parallel import tile_windows
    for (h, w, ty, tx, halo) in [(96, 96, 2, 2, 16), (64, 100, 1, 4, 16), (37, 53, 3, 2, 8), (40, 40, 1, 1, 16), (50, 30, 4, 1, 4)]:
        win_h, win_w, wins = tile_windows(h, w, ty, tx, halo)
        cover = torch.zeros(h, w, dtype=torch.int32)
        for (y0, x0, a, b, c, d) in wins:
            cover[a:b, c:d] += 1
            assert 0 <= y0 and y0 + win_h <= h and 0 <= x0 and x0 + win_w <= w
            # at least `halo` pixels of context on every side that is not an image border
            assert (a - y0 >= halo or y0 == 0) and (y0 + win_h - b >= halo or y0 + win_h == h)
            assert (c - x0 >= halo or x0 == 0) and (x0 + win_w - d >= halo or x0 + win_w == w)
        assert int(cover.min()) == 1 and int(cover.max()) == 1


def _tile_worker(rank, world, port, out_dir):
    from esr_b200.parallel import run_tiled
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    wts = synth.make_weights("kaiming", seed=4, nb=1)
    lr, z = synth.make_inputs(1, 40, 44, seed=4)
    net = GCEMOracle(wts, nb=1)
    out = run_tiled(net.forward, concat_latent(lr, z), tiles=(2, 2), halo=14)
    torch.save(out, os.path.join(out_dir, "tiled_%d.pt" % rank))
    dist.destroy_process_group()


def test_two_rank_tile_sharding_matches_full_image(tmp_path):
    """One image split into 2 x 2 halo-overlapped tiles over 2 gloo ranks (the per-tile function is the oracle on a
    one-block net, standing in for the CUDA forward): every rank ends up with the same stitched image, and it equals
    the untiled forward up to the tiling approximation (tiny for a one-block net with a 14-pixel halo)."""
    port = 31500 + os.getpid() % 2000
    mp.spawn(_tile_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = [torch.load(os.path.join(str(tmp_path), "tiled_%d.pt" % r)) for r in range(2)]
    assert torch.equal(a, b) and a.shape == (1, 3, 160, 176)
    torch.set_num_threads(2)
    wts = synth.make_weights("kaiming", seed=4, nb=1)
    lr, z = synth.make_inputs(1, 40, 44, seed=4)
    with torch.no_grad():
        full = GCEMOracle(wts, nb=1).forward(concat_latent(lr, z))
    assert (a - full).abs().max().item() < 1e-4

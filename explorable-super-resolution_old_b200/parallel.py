"""Multi-GPU plumbing: one process per GPU, the batch is sharded across ranks and no collective sits on the
data path (images are independent: no BatchNorm, nothing couples batch elements in G or CEM; SURVEY.md
§8e).  Replaces the thread-per-GPU ``nn.DataParallel`` scatter/gather of codes/models/networks.py:99-101.
``torch.distributed`` is used only for the optional gather of the 3-channel outputs."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of the n batch items owned by `rank` (first n % world ranks get one more)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def run_sharded(fn, batch, gather=True):
    """Applies fn to this rank's shard of `batch` (dim 0); optionally all-gathers the results in rank order."""
    if not (dist.is_available() and dist.is_initialized()):
        return fn(batch)
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(batch.size(0), rank, world)
    out = fn(batch[lo:hi].contiguous()) if hi > lo else None
    if not gather:
        return out
    sizes = [shard_range(batch.size(0), r, world) for r in range(world)]
    shape = list(out.shape[1:]) if out is not None else None
    shapes = [None] * world
    dist.all_gather_object(shapes, shape)
    shape = next(s for s in shapes if s is not None)
    most = max(h - l for l, h in sizes)                      # all_gather needs equal sizes: pad, then trim
    mine = batch.new_zeros([most] + shape)
    if out is not None:
        mine[:out.size(0)] = out
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return torch.cat([p[:h - l] for p, (l, h) in zip(parts, sizes)], 0)


class HostPipeline:
    """Host-buffer entry point for inference: ``netG`` over a pinned host batch, in chunks, with the host->device
    copy of chunk i+1 and the device->host copy of chunk i-1 running on their own streams while chunk i computes.
    The chunks go through ``netG`` one after the other on the caller's stream (so they share one set of plan
    buffers, and a chunk's dense-block working set is the L2-friendly size), only the PCIe copies overlap.
    Calls return once everything is enqueued, so consecutive calls pipeline too (the next batch's upload runs under
    this batch's compute).  ``join()`` makes the caller's stream wait for the last download, ``wait()`` blocks the
    host until it is complete; read ``host_out`` only after one of them."""

    def __init__(self, netG, chunk=16, use_graph=True):
        self.netG, self.chunk = netG, chunk
        self.use_graph = use_graph                                         # replay the chunk's forward as a CUDA graph
        self._graphs = {}                                                  # (slot, rows, input shape, module state) -> (graph, static output)
        self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()
        self._x, self._free, self._k = [None, None], [None, None], 0     # two device staging slots for the inputs
        self._down = [None, None]                                          # per slot: download of the chunk that last used it
        self._live = []                                                    # outputs whose download is still in flight
        self._done = None

    def __call__(self, host_in, host_out):
        assert host_in.is_pinned() and host_out.is_pinned(), "HostPipeline needs pinned host buffers (asynchronous copies)"
        dev = next(self.netG.parameters()).device
        cur = torch.cuda.current_stream()
        n = host_in.size(0)
        for lo in range(0, n, self.chunk):
            hi = min(lo + self.chunk, n)
            slot = self._k & 1
            self._k += 1
            if self._x[slot] is None or self._x[slot].shape[1:] != host_in.shape[1:] or self._x[slot].size(0) < hi - lo:
                self._x[slot] = torch.empty((self.chunk,) + tuple(host_in.shape[1:]), dtype=host_in.dtype, device=dev)
                self._free[slot] = None
            x = self._x[slot][:hi - lo]
            up = torch.cuda.Event()
            with torch.cuda.stream(self.s_in):
                if self._free[slot] is not None:
                    self.s_in.wait_event(self._free[slot])     # the chunk that last used this slot has been consumed
                x.copy_(host_in[lo:hi], non_blocking=True)
                up.record(self.s_in)
            cur.wait_event(up)
            # the captured graph writes ONE static output buffer per slot: chunk k+2 must not overwrite it while chunk
            # k's device->host copy (on s_out) is still reading it
            if self._down[slot] is not None:
                cur.wait_event(self._down[slot])
            out = self._forward(slot, x)
            done = torch.cuda.Event()
            done.record(cur)
            self._free[slot] = done
            down = torch.cuda.Event()
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(done)
                host_out[lo:hi].copy_(out, non_blocking=True)
                down.record(self.s_out)
            self._live = [(o, e) for (o, e) in self._live if not e.query()] + [(out, down)]
            self._done = self._down[slot] = down
        return host_out

    def _forward(self, slot, x):
        """netG(x) for the staging slot: through the module's captured inference graph when it offers one
        (CEM_PyTorch.capture: fixed input slot, fixed output buffer), else the eager module call."""
        if self.use_graph and hasattr(self.netG, "capture"):
            # a captured graph bakes in the packed weights, the margin (train / eval) and the geometry: key on all of
            # them, so load_state_dict, a weight update or a train()/eval() flip re-captures instead of replaying stale state
            state = tuple((p.data_ptr(), p._version) for p in self.netG.parameters())
            key = (slot, tuple(x.shape), x.data_ptr(), bool(getattr(self.netG, "pre_pad", False)), state)
            if key not in self._graphs:
                for old in [k for k in self._graphs if k[0] == slot and k[1] == key[1]]:
                    torch.cuda.current_stream().synchronize()              # nothing may still be replaying the stale graph
                    del self._graphs[old]
                self._graphs[key] = self.netG.capture(x, slot=slot) or False
            ent = self._graphs[key]
            if ent:
                ent[0].replay()
                return ent[1]
        with torch.no_grad():
            return self.netG(x)

    def join(self):
        if self._done is not None:
            torch.cuda.current_stream().wait_event(self._done)

    def wait(self):
        if self._done is not None:
            self._done.synchronize()


# ------------------------------------------------------------------------------------------ tile sharding (B = 1)
def tile_windows(h, w, tiles_y, tiles_x, halo):
    """Splits an h x w LR image into tiles_y x tiles_x output tiles and gives every tile an input window of ONE common
    size that contains the tile plus at least `halo` LR pixels on every side that is not an image border (windows of
    border tiles are shifted inwards instead of shrunk, so all windows can run as one batch).
    Returns (win_h, win_w, [(y0, x0, oy0, oy1, ox0, ox1)]): window origin and the tile's rows / columns, LR pixels."""
    assert tiles_y >= 1 and tiles_x >= 1 and halo >= 0
    ys = [shard_range(h, i, tiles_y) for i in range(tiles_y)]
    xs = [shard_range(w, i, tiles_x) for i in range(tiles_x)]
    win_h = min(h, max(b - a for a, b in ys) + (2 * halo if tiles_y > 1 else 0))
    win_w = min(w, max(b - a for a, b in xs) + (2 * halo if tiles_x > 1 else 0))
    wins = []
    for (a, b) in ys:
        y0 = min(max(a - halo, 0), h - win_h)
        assert (y0 <= a - halo or y0 == 0) and (y0 + win_h >= b + halo or y0 + win_h == h)
        for (c, d) in xs:
            x0 = min(max(c - halo, 0), w - win_w)
            assert (x0 <= c - halo or x0 == 0) and (x0 + win_w >= d + halo or x0 + win_w == w)
            wins.append((y0, x0, a, b, c, d))
    return win_h, win_w, wins


def run_tiled(netG, model_input, tiles, halo=16, sf=4, nz=3, gather=True):
    """One large image across the ranks by halo-overlapped spatial tiles (BASELINE north_star "by batch or
    image tiles (halo-overlapped, no NCCL needed for inference)").  The reference's own precedent is the GUI's crop to
    the edited bounding box + 30 px (codes/GUI.py:1530-1544).

    model_input: [1, 16*nz+3, h, w] packed [Z.view, LR] (SRRaGAN_model.py:249-255).  Every tile's window is cut out of the
    un-viewed Z ([1,nz,4h,4w]) and of the LR image, re-packed, and this rank's windows run through netG as ONE batch;
    each output is cropped to its tile and written into the [1,3,4h,4w] result.  Unlike batch sharding this is an
    approximation: a pixel closer than the network's effective receptive field to a tile border sees replicate padding
    instead of its true neighbours.  With a halo of 16 LR px the difference is below the parity tolerance for the
    random-init weights of the tests (tests/test_gpu_net.py::test_tile_sharding_error_vs_halo); trained checkpoints
    need their halo re-measured.  With gather=True every rank returns the full image (one all_gather of the 3-channel
    tiles); otherwise tiles owned by other ranks are left zero."""
    assert model_input.dim() == 4 and model_input.size(0) == 1, "run_tiled shards ONE image; use run_sharded for batches"
    _, C, h, w = model_input.shape
    assert C == nz * sf * sf + 3
    tiles_y, tiles_x = tiles
    win_h, win_w, wins = tile_windows(h, w, tiles_y, tiles_x, halo)
    rank, world = (dist.get_rank(), dist.get_world_size()) if (dist.is_available() and dist.is_initialized()) else (0, 1)
    lo, hi = shard_range(len(wins), rank, world)
    z_hr = model_input[:, :nz * sf * sf].contiguous().view(1, nz, sf * h, sf * w) if nz else None
    lr = model_input[:, -3:]
    batch = []
    for (y0, x0, *_r) in wins[lo:hi]:
        parts = []
        if nz:
            zc = z_hr[:, :, sf * y0:sf * (y0 + win_h), sf * x0:sf * (x0 + win_w)].contiguous()
            parts.append(zc.view(1, nz * sf * sf, win_h, win_w))
        parts.append(lr[:, :, y0:y0 + win_h, x0:x0 + win_w])
        batch.append(torch.cat(parts, 1))
    mine = None
    if batch:
        with torch.no_grad():
            mine = netG(torch.cat(batch, 0).contiguous())
    per = max(b - a for a, b in [shard_range(len(wins), r, world) for r in range(world)])
    if world > 1 and gather:
        buf = model_input.new_zeros((per, 3, sf * win_h, sf * win_w))
        if mine is not None:
            buf[:mine.size(0)] = mine
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
        outs = [(r, parts[r]) for r in range(world)]
    else:
        outs = [(rank, mine)] if mine is not None else []
    full = model_input.new_zeros((1, 3, sf * h, sf * w))
    for r, t in outs:
        rlo, rhi = shard_range(len(wins), r, world)
        for k, (y0, x0, a, b, c, d) in enumerate(wins[rlo:rhi]):
            full[0, :, sf * a:sf * b, sf * c:sf * d] = t[k, :, sf * (a - y0):sf * (b - y0), sf * (c - x0):sf * (d - x0)]
    return full

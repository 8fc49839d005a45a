#!/usr/bin/env python
"""Benchmark of the RRDBNet(+Z) -> CEM hot path (BASELINE.json metric: x4 SR output Mpix/s).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference's own modules on the host cores
  python bench.py --scaling strong ...                     # config 2 as BASELINE states it: ONE batch of 16 split over the ranks

One "step" = one G+CEM forward over a batch of 16 synthetic 3x128x128 LR images + random Z in eval mode
(CEM pre-pad by 10 px), BASELINE config 2.  N > 1: one process per GPU (torchrun), every rank runs its own
batch of 16 (batch sharding, no data-path collective: images are independent), scaling = weak; the same line carries
a "strong" object (the one batch of 16 split 16/N per rank, timed in the same run) and --scaling strong makes that the
headline value.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

BATCH, LR_H, LR_W, SF, MARGIN = 16, 128, 128, 4, 10
WORKLOAD = "RRDBNet x4 (nf=64, nb=23, gc=32, all_layers Z) + CEM, eval/pre-pad, batch 16 x 3x128x128 LR per GPU"
METRIC, UNIT = "x4 SR output Mpix/s (RRDBNet+CEM)", "Mpix/s"


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(bf16=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), bf16_burst=float(p["bf16_tflops"]),
                    hbm=float(p["hbm_gbs"]), src="measured")
    except Exception:
        return dict(bf16=1400.0, bf16_burst=1590.0, hbm=6650.0, src="fallback")


class ClockSampler(threading.Thread):
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for n, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def reference_forward_fn(n_images, h, w, threads):
    """(callable running one forward of the workload over n_images images on the host cores, kind).  kind "reference":
    the UNMODIFIED reference's own define_G -> CEM_PyTorch(RRDBNet) modules (from /root/reference in the build
    container, from the vendored oracle/_ref on the GPU box); "port": the oracle restatement, when neither exists."""
    from esr_b200 import synth
    from oracle import ref_shims
    from oracle.cem_ops import concat_latent
    torch.set_num_threads(threads)
    lr, z = synth.make_inputs(n_images, h, w, seed=0)
    mi = concat_latent(lr, z)
    if ref_shims.available() and os.environ.get("ESR_BENCH_REF_PORT", "0") != "1":
        import contextlib as _c, io as _io
        with _c.redirect_stdout(_io.StringIO()):
            CEMnet, networks, _, _ = ref_shims.load_reference()
            netG, _ = ref_shims.build_ref_G(CEMnet, networks, 23, "all_layers", "default", 0)
        netG = netG.cpu().float()                                  # the reference builds its CEM filters on the GPU when one exists
        netG.train(False)                                          # eval: CEM pre-pads by 10 px (CEMnet.py:192-194)

        def fwd():
            with torch.no_grad():
                return netG(mi)
        return fwd, "reference"
    from oracle.rrdbnet import GCEMOracle
    net = GCEMOracle(synth.make_weights("default", seed=0))

    def fwd_port():
        with torch.no_grad():
            return net.forward(mi)
    return fwd_port, "port"


def timed(fn):
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


def conv_traffic(key="conv_dram_bytes_per_step"):
    """DRAM bytes (read + write) of the 351 conv launches of one step (or, with another key, of the CEM projection at
    config 4), from the committed ncu captures."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)[key]
    except Exception:
        return None


def cem_standalone(dev, pk):
    """BASELINE config 4: CEM wrapping a 2048x2048 x4 SR output (bicubic), HBM roofline.  Inputs larger than L2: six
    (y, x, out) sets (624 MB >> 126 MB L2) take turns, 60 calls back to back between one event pair, so every call
    streams its 103.8 MB through HBM (the previous call's dirty output lines are written back under it: steady state).
    The round-1 protocol (one call after a 256 MiB memset that leaves L2 full of dirty lines, median of single-call event
    pairs, 2 us timer granularity) is reported beside it as `after_write_flush_us`."""
    from esr_b200 import _capi as capi, cem as pcem
    f = pcem.CEMnet(pcem.Get_CEM_Config(SF))._filters
    B, C, H, W = 1, 3, 2048, 2048
    nset = 6
    sets = [(torch.rand(B, C, H, W, device=dev), torch.rand(B, C, H // SF, W // SF, device=dev), torch.empty(B, C, H, W, device=dev))
            for _ in range(nset)]
    ws = torch.empty(2 * B * C * (H // SF) * (W // SF), device=dev)
    l = capi.lib()

    def run(k):
        y, x, out = sets[k % nset]
        capi.check(l.esr_cem_project(f, capi.ptr(y), capi.ptr(x), B, C, H, W, 0, capi.ptr(out), capi.ptr(ws), capi.stream_ptr()))
    for k in range(2 * nset):
        run(k)
    torch.cuda.synchronize()
    reps = 60
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(reps):
        run(k)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / reps
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        run(0)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t_flush = sorted(ts)[len(ts) // 2]
    byts = 4 * C * (2 * H * W + H * W // (SF * SF))           # read y, read x, write out = 24.75 B / HR px
    return {"bound": "hbm", "achieved": byts / t / 1e6, "peak": pk["hbm"], "unit": "GB/s", "frac": byts / t / 1e6 / pk["hbm"],
            "us": t * 1e3, "after_write_flush_us": t_flush * 1e3, "frac_after_write_flush": byts / t_flush / 1e6 / pk["hbm"],
            "kernel": "cem_down4s_kernel + cem_invup4s_kernel (TMA-fed rings; 2 launches, the second a programmatic dependent "
                      "of the first), 1x3x2048x2048 output, algorithmic %.1f MB; inputs larger than L2 (six buffer sets in "
                      "rotation, 60 calls per event pair); traffic = DRAM bytes of both launches (ncu, cold cache per launch, "
                      "profiles/)" % (byts / 1e6),
            "traffic": conv_traffic("cem_cfg4_dram_bytes")}

def workload_config(world, per_rank, graph=True, sub=None, nstreams=1):
    return {"workload": WORKLOAD, "global_batch": per_rank * world, "parallelism": "batch shard x%d" % world,
            "l2": "per-step working set (~4.5 GB of activations) >> 126 MB L2, no flush needed",
            "cuda_graph": graph, "images_per_pass": sub if sub is not None else per_rank, "passes_in_flight": nstreams}


def run_reference(args):
    """The reference's CPU implementation of the path on this box's host cores, all threads.  One step = a bounded
    sample of the workload (REF_SAMPLE of its 16 images, same image size / mode), so that --steps K --warmup W ends
    within minutes; `steps` and `ms_per_step` are what was really run and timed, `value` the sample's throughput."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_img = int(os.environ.get("ESR_BENCH_REF_IMAGES", 1))
    fwd, kind = reference_forward_fn(n_img, LR_H, LR_W, cores)
    budget = float(os.environ.get("ESR_BENCH_REF_BUDGET_S", 150))
    t_first = timed(fwd)                                            # also the warm-up (thread pool, allocator)
    warm = max(0, min(args.warmup, int(budget * 0.2 / max(t_first, 1e-3))) - 1)
    for _ in range(warm):
        fwd()
    steps = max(1, min(args.steps, int(budget * 0.8 / max(t_first, 1e-3))))
    t0 = time.perf_counter()
    for _ in range(steps):
        fwd()
    t = (time.perf_counter() - t0) / steps
    mpix = n_img * SF * SF * LR_H * LR_W / 1e6 / t
    sample = "%d of the %d images of the batch per step (3x%dx%d LR each, eval/pre-pad, fp32), %d steps timed after %d warm-up " \
             "passes, %s" % (n_img, BATCH, LR_H, LR_W, steps, warm + 1,
                             "the reference's own nn.Modules" if kind == "reference" else "oracle port (reference tree absent)")
    line = {"impl": "reference", "metric": METRIC, "value": mpix, "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warm + 1, "ms_per_step": t * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, BATCH),              # the same dict as this repo's arm at the same N
            "cpu_baseline": {"value": mpix, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": mpix, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


class StepRunner:
    """G+CEM forward of `per_rank` resident images through the recorded launch sequence (inputs already in HBM)."""

    def __init__(self, netG, dev, per_rank, seed, use_graph=True):
        from esr_b200 import _capi as capi, synth
        self.capi, self.netG, self.G, self.dev, self.n = capi, netG, netG.generated_image_model, dev, per_rank
        lr, z = synth.make_inputs(per_rank, LR_H, LR_W, seed=seed)
        self.host_in = torch.cat([z.contiguous().view(per_rank, 16 * 3, LR_H, LR_W), lr], 1).contiguous().pin_memory()
        self.x_dev = self.host_in.to(dev)
        self.host_out = torch.empty(per_rank, 3, SF * LR_H, SF * LR_W).pin_memory()
        self.SUB = min(int(os.environ.get("ESR_SUBBATCH", per_rank)), per_rank)     # images per pass through the layers
        self.NSTREAMS = int(os.environ.get("ESR_STREAMS", 1))                       # sub-batches in flight
        assert per_rank % self.SUB == 0 and (per_rank // self.SUB) % self.NSTREAMS == 0
        self.plans = [self.G.plan(self.SUB, LR_H, LR_W, MARGIN, keep=False, slot=i) for i in range(self.NSTREAMS)]
        plan = self.plans[0]
        self.filters = netG._filters
        self.out_dev = torch.empty(per_rank, 3, SF * LR_H, SF * LR_W, device=dev)
        self.wss = [torch.empty(2 * self.SUB * 3 * plan.hp * plan.wp, device=dev) for _ in range(self.NSTREAMS)]
        self.side = [torch.cuda.Stream() for _ in range(self.NSTREAMS)] if self.NSTREAMS > 1 else []
        self.ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        self.H4, self.W4 = SF * plan.hp, SF * plan.wp
        self.launches_per_fwd = plan.launches_per_forward(with_cem=True)
        self.graph = self.conv_graph = None
        for _ in range(3):
            self.step()
        torch.cuda.synchronize()
        self.step(record=True)                       # per-phase split (untimed iteration)
        torch.cuda.synchronize()
        ev = self.ev
        self.phases = {"prep": ev[0].elapsed_time(ev[1]), "convs": ev[1].elapsed_time(ev[2]), "cem": ev[2].elapsed_time(ev[3])}
        if use_graph:
            self.graph = self._capture(self.step)
            self.conv_graph = self._capture(self.convs_only)
        self.run = self.graph.replay if self.graph is not None else self.step
        self.conv_run = self.conv_graph.replay if self.conv_graph is not None else self.convs_only

    def _capture(self, fn):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            fn()
        torch.cuda.current_stream().wait_stream(s)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        for _ in range(2):
            g.replay()
        return g

    def convs_only(self):
        for i in range(self.n // self.SUB):
            self.plans[i % self.NSTREAMS].run_convs()

    def step(self, record=False):
        capi, l, ev = self.capi, self.capi.lib(), self.ev
        main = torch.cuda.current_stream()
        for s_ in self.side:
            s_.wait_stream(main)
        for k, i0 in enumerate(range(0, self.n, self.SUB)):
            first = i0 == 0
            pl, w_ = self.plans[k % self.NSTREAMS], self.wss[k % self.NSTREAMS]
            with torch.cuda.stream(self.side[k % self.NSTREAMS]) if self.side else contextlib.nullcontext():
                if record and first:
                    ev[0].record()
                pl.run_prep(self.x_dev[i0:i0 + self.SUB])
                if record and first:
                    ev[1].record()
                pl.run_convs()
                if record and first:
                    ev[2].record()
                capi.check(l.esr_cem_project(self.filters, capi.ptr(pl.y), capi.ptr(pl.lr_pad), self.SUB, 3, self.H4, self.W4,
                                             SF * MARGIN, capi.ptr(self.out_dev[i0:i0 + self.SUB]), capi.ptr(w_), capi.stream_ptr()))
                if record and first:
                    ev[3].record()
        for s_ in self.side:
            main.wait_stream(s_)

    def timed_steps(self, steps, barrier):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.run()
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / steps

    def conv_share(self, rounds):
        """Share of the step spent in the 351 conv launches: step and conv-only graphs replayed ALTERNATELY in one
        window (same clocks, same thermal state), each replay bracketed by events; returns sum(conv) / sum(step)."""
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(rounds)]
        for a, b, c in evs:
            a.record()
            self.run()
            b.record()
            self.conv_run()
            c.record()
        torch.cuda.synchronize()
        ts, tc = sum(a.elapsed_time(b) for a, b, _ in evs), sum(b.elapsed_time(c) for _, b, c in evs)
        return tc / ts, ts / rounds, tc / rounds

    def release(self):
        self.graph = self.conv_graph = self.plans = self.wss = self.out_dev = None
        self.G._plans.clear()
        torch.cuda.empty_cache()


def config1_latency(netG, dev):
    """BASELINE config 1 (1x3x64x64 LR, eval/pre-pad) as a latency: one CUDA-graph replay of the whole forward, L2 warm,
    median of 20 (361 launches for 0.16 ms of ideal math - a launch-overhead measurement)."""
    from esr_b200 import synth
    from esr_b200.rrdbnet import capture_inference
    lr, z = synth.make_inputs(1, 64, 64, seed=1)
    x = torch.cat([z.contiguous().view(1, 48, 64, 64), lr], 1).contiguous().to(dev)
    G = netG.generated_image_model
    graph, out = capture_inference(G, x, MARGIN, netG._filters, slot=7)
    for _ in range(3):
        graph.replay()
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        graph.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = sorted(ts)[len(ts) // 2]
    plan = G.plan(1, 64, 64, MARGIN, keep=False, slot=7)
    n_launch = plan.launches_per_forward(with_cem=True)
    flops = G.engine().flops_per_lr_pixel() * (64 + 2 * MARGIN) ** 2
    res = {"config": "1x3x64x64 LR + Z, eval/pre-pad (BASELINE config 1), CUDA-graph replay", "latency_us": t * 1e3,
           "launches": n_launch, "us_per_launch": t * 1e3 / n_launch, "output_mpix_per_s": 16 * 64 * 64 / 1e6 / (t * 1e-3),
           "algorithmic_tflops": flops / (t * 1e-3) / 1e12}
    del graph
    G._plans.clear()
    return res


def run_train_g(args):
    """Generator half of BASELINE config 5: batch 16 x 3x32x32 LR (128x128 HR patches) per rank, train mode (no CEM pad),
    L1 pixel loss on fake_H, explicit data + weight gradients, bucketed NCCL all-reduce of the 68 MB gradient under the
    weight-gradient kernels, torch Adam on the fp32 master weights.  Metric: HR patches per second, whole job."""
    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    from esr_b200 import _capi as capi, cem as pcem, networks, synth
    from esr_b200.training import GeneratorTrainer
    capi.lib()
    opt = {"gpu_ids": None, "is_train": False, "datasets": {"train": {"patch_size": 128}},
           "network_G": dict(which_model_G="RRDB_net", CEM_arch=1, latent_input="all_layers", latent_input_domain="HR_downscaled",
                             latent_channels=3, norm_type=None, mode="CNA", nf=64, nb=23, in_nc=3, out_nc=3, gc=32, scale=SF)}
    netG = networks.define_G(opt, CEM=pcem.CEMnet(pcem.Get_CEM_Config(SF)), num_latent_channels=3)
    sd = netG.state_dict()
    sd.update({"generated_image_model." + k: v for k, v in synth.make_weights("kaiming", seed=0).items()})
    netG.load_state_dict(sd)
    netG.to(dev).train()
    trainer = GeneratorTrainer(netG)
    G = netG.generated_image_model
    optim = torch.optim.Adam([trainer.flat_parameter()], lr=1e-4, betas=(0.9, 0.999), fused=True)   # every weight as views of one flat parameter: one launch
    Bp, hl = 16, 32
    lr, z = synth.make_inputs(Bp, hl, hl, seed=rank)
    mi = torch.cat([z.contiguous().view(Bp, 48, hl, hl), lr], 1).contiguous().to(dev)
    target = torch.rand(Bp, 3, SF * hl, SF * hl, generator=torch.Generator().manual_seed(rank)).to(dev)

    def step():
        fake = trainer.forward(mi)
        loss = (fake - target).abs().mean()
        loss.backward()
        trainer.backward(fake.grad)
        optim.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()
    for _ in range(max(args.warmup, 3)):
        l0 = step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        l1 = step()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.cpu())
    if rank == 0:
        flops = 3 * G.engine().flops_per_lr_pixel() * Bp * hl * hl          # forward + data gradient + weight gradient
        print(json.dumps({"metric": "generator training step, 128x128 HR patches/s (RRDBNet+CEM, L1 pixel loss)", "value": world * Bp / (ms * 1e-3),
                          "unit": "patches/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms,
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16 MMA operands, fp32 master weights",
                          "data": "synthetic", "config": {"workload": "BASELINE config 5, generator half: 16 x 3x32x32 LR per rank, train mode",
                                                          "global_batch": Bp * world, "parallelism": "data parallel x%d, bucketed NCCL all-reduce (AVG) of %.1f MB of gradients" % (world, trainer.grad_bytes() / 1e6)},
                          "algorithmic_tflops_per_gpu": flops / (ms * 1e-3) / 1e12, "loss_first_last": [float(l0), float(l1)]}))
    if world > 1:
        torch.distributed.destroy_process_group()


def gan_step_measure(dev, rank, world, steps, warmup):
    """BASELINE config 5, the whole step: batch 16 x 128x128 HR patches (3x32x32 LR) per rank, train mode, through
    training.GanTrainer — G forward, critic update on real / fake / WGAN-GP interpolates (double backward), generator update
    with the adversarial + range terms back-propagated through the critic into this package's dgrad and weight-gradient
    kernels, NCCL all-reduce of the 13.6 MB critic gradient (one call) and the 68.2 MB generator gradient (buckets, half of
    them under the weight-gradient kernels), fused Adam on both.  The critic (Discriminator_VGG_128_, n_layers 6, nf 64) is
    torch code on cuDNN (library); it sees the whole 128x128 patch (the CEM-cropped 48x48 one is below its 8x8 head's
    minimum).  Needs an initialised process group when world > 1.  Returns the JSON-able result (max over ranks)."""
    from esr_b200 import _capi as capi, cem as pcem, networks, synth
    from esr_b200.discriminator import Discriminator_VGG_128_
    from esr_b200.training import GanTrainer
    capi.lib()
    opt = {"gpu_ids": None, "is_train": False, "datasets": {"train": {"patch_size": 128}},
           "network_G": dict(which_model_G="RRDB_net", CEM_arch=1, latent_input="all_layers", latent_input_domain="HR_downscaled",
                             latent_channels=3, norm_type=None, mode="CNA", nf=64, nb=23, in_nc=3, out_nc=3, gc=32, scale=SF)}
    netG = networks.define_G(opt, CEM=pcem.CEMnet(pcem.Get_CEM_Config(SF)), num_latent_channels=3)
    sd = netG.state_dict()
    sd.update({"generated_image_model." + k: v for k, v in synth.make_weights("kaiming", seed=0).items()})
    netG.load_state_dict(sd)
    netG.to(dev).train()
    torch.manual_seed(0)
    netD = Discriminator_VGG_128_(3, 64, nb=6, input_patch_size=128)
    networks.init_weights(netD, init_type="kaiming", scale=1)
    netD.to(dev).train()
    torch.backends.cudnn.benchmark = os.environ.get("ESR_D_CUDNN_BENCHMARK", "1") == "1"   # the torch critic: 55.8 -> 50.8 ms per step
    gan = GanTrainer(netG, netD, lr_G=1e-5, lr_D=1e-5, pixel_weight=1e-2, gan_weight=1.0, gp_weight=10.0, range_weight=5000.0, crop=0)
    Bp, hl = 16, 32
    lr, z = synth.make_inputs(Bp, hl, hl, seed=rank)
    mi = torch.cat([z.contiguous().view(Bp, 48, hl, hl), lr], 1).contiguous().to(dev)
    target = torch.rand(Bp, 3, SF * hl, SF * hl, generator=torch.Generator().manual_seed(rank)).to(dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()
    for _ in range(max(warmup, 3)):
        l0 = dict(gan.step(mi, target))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        l1 = gan.step(mi, target)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.cpu())
    gb, db = gan.grad_bytes()
    flops = 3 * netG.generated_image_model.engine().flops_per_lr_pixel() * Bp * hl * hl
    res = {"metric": "GAN training step (RRDBNet+CEM generator, VGG-128 critic, WGAN-GP), 128x128 HR patches/s", "value": world * Bp / (ms * 1e-3),
           "unit": "patches/s", "n_gpus": world, "steps": steps, "warmup": max(warmup, 3), "ms_per_step": ms,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "generator: bf16 MMA operands, fp32 master weights; critic: torch fp32 (cuDNN, TF32 as torch defaults)",
           "data": "synthetic", "config": {"workload": "BASELINE config 5: 16 x 128x128 HR patches (3x32x32 LR + Z) per rank, train mode, "
                                                       "critic update + generator update every step (D_update_ratio 1)",
                                           "global_batch": Bp * world,
                                           "parallelism": "data parallel x%d, NCCL all-reduce (AVG): generator %.1f MB in buckets (half of them under the "
                                                          "weight-gradient kernels), critic %.1f MB in one call" % (world, gb / 1e6, db / 1e6)},
           "generator_algorithmic_tflops_per_gpu": flops / (ms * 1e-3) / 1e12,
           "losses_first": {k: float(v) for k, v in l0.items()}, "losses_last": {k: float(v) for k, v in l1.items()}}
    del gan, netG, netD
    torch.cuda.empty_cache()
    return res


def run_train_gan(args):
    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    with contextlib.redirect_stdout(sys.stderr):
        res = gan_step_measure(dev, rank, world, args.steps, args.warmup)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: 16 images per rank (default); strong: BASELINE config 2's one batch of 16 split over the ranks")
    ap.add_argument("--workload", default="infer", choices=["infer", "train_g", "train_gan"],
                    help="infer: BASELINE config 2 (default, the headline metric); train_g: the generator half of config 5's "
                         "training step (forward, data + weight gradients, NCCL gradient all-reduce, Adam), 16 x 32x32 LR patches per rank")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-zopt", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the BASELINE config 5 (GAN training step) block of the default run")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "train_g":
        return run_train_g(args)
    if args.workload == "train_gan":
        return run_train_gan(args)

    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    from esr_b200 import _capi as capi, cem as pcem, networks, synth
    from esr_b200.parallel import shard_range
    capi.lib()                                                # fail loudly if the extension is missing
    opt = {"gpu_ids": None, "is_train": False, "datasets": {"train": {"patch_size": 256}},
           "network_G": dict(which_model_G="RRDB_net", CEM_arch=1, latent_input=os.environ.get("ESR_BENCH_LATENT", "all_layers"),
                             latent_input_domain="HR_downscaled", latent_channels=3, norm_type=None, mode="CNA",
                             nf=64, nb=23, in_nc=3, out_nc=3, gc=32, scale=SF)}
    netG = networks.define_G(opt, CEM=pcem.CEMnet(pcem.Get_CEM_Config(SF)), num_latent_channels=3)
    sd = netG.state_dict()
    _lat = os.environ.get("ESR_BENCH_LATENT", "all_layers")      # "first_layer": experiment only (not BASELINE's config)
    sd.update({"generated_image_model." + k: v for k, v in synth.make_weights(
        "default", seed=0, latent_input=_lat + "_HR_downscaled").items()})
    netG.load_state_dict(sd)
    netG.to(dev).eval()
    for p in netG.parameters():
        p.requires_grad_(False)
    G = netG.generated_image_model

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(vals):
        t = torch.tensor(vals, device=dev, dtype=torch.float64)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return [float(v) for v in t.cpu()]

    strong_n = shard_range(BATCH, rank, world)
    strong_n = strong_n[1] - strong_n[0]                       # images of the ONE batch of 16 that this rank owns
    per_rank = BATCH if args.scaling == "weak" else strong_n
    assert per_rank >= 1, "more ranks than images"
    warmup = max(args.warmup, 3)

    runner = StepRunner(netG, dev, per_rank, seed=rank, use_graph=not args.no_graph)
    for _ in range(warmup):
        runner.run()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = runner.timed_steps(args.steps, barrier)
    share, _, _ = runner.conv_share(max(5, min(args.steps, 20)))
    conv_ms = ms * share

    # end to end through the public module API with host buffers: netG(...) on pinned host input, result copied
    # back to pinned host memory; parallel.HostPipeline overlaps the PCIe copies with compute (both copies of every step
    # are inside the timed region)
    from esr_b200.parallel import HostPipeline
    pipe = HostPipeline(netG, chunk=int(os.environ.get("ESR_E2E_CHUNK", BATCH)))

    def e2e_step():
        pipe(runner.host_in, runner.host_out)
    for _ in range(max(args.warmup, 4)):                      # the caching allocator settles after a few calls
        e2e_step()
    pipe.wait()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e2e = max(5, args.steps)
    f0.record()
    for _ in range(n_e2e):
        e2e_step()
    pipe.join()                                               # the last download is inside the timed region
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1) / n_e2e
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d, d2h = runner.host_in.numel() * 4, runner.host_out.numel() * 4
    launches_per_fwd, SUB, NSTREAMS, phases = runner.launches_per_fwd, runner.SUB, runner.NSTREAMS, runner.phases
    graph_used = runner.graph is not None
    del pipe
    runner.release()

    # BASELINE config 2 exactly as stated: the ONE batch of 16 images split over the ranks (16 / 8 / 4 / 2 per rank)
    strong = None
    if world > 1 and args.scaling == "weak":
        r2 = StepRunner(netG, dev, strong_n, seed=100 + rank, use_graph=not args.no_graph)
        for _ in range(warmup):
            r2.run()
        s_ms = r2.timed_steps(args.steps, barrier)
        r2.release()
        s_ms, = max_over_ranks([s_ms])
        strong = {"value": BATCH * SF * SF * LR_H * LR_W / 1e6 / (s_ms * 1e-3), "unit": UNIT, "ms_per_step": s_ms,
                  "images_per_rank": strong_n, "global_batch": BATCH, "scaling": "strong",
                  "note": "BASELINE config 2's single batch of 16 split over the ranks, no data-path collective"}

    # One config-3-sized image (1x3x256x256 LR) across the ranks by halo-overlapped tiles (parallel.run_tiled, halo 16)
    tiled = None
    if not args.no_zopt:
        from esr_b200.parallel import run_tiled
        grid = {1: (1, 1), 2: (2, 1), 4: (2, 2), 8: (4, 2)}.get(world, (world, 1))
        lr_t, z_t = synth.make_inputs(1, 256, 256, seed=5)
        mi_t = torch.cat([z_t.contiguous().view(1, 48, 256, 256), lr_t], 1).contiguous().to(dev)
        for _ in range(3):
            run_tiled(netG, mi_t, tiles=grid, halo=16)
        barrier()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_t = 5
        t0e.record()
        for _ in range(n_t):
            run_tiled(netG, mi_t, tiles=grid, halo=16)
        t1e.record()
        barrier()
        t_ms, = max_over_ranks([t0e.elapsed_time(t1e) / n_t])
        tiled = {"config": "1x3x256x256 LR + Z, eval/pre-pad, %d x %d tiles with a 16 px halo over %d rank(s), outputs "
                           "all-gathered (module call per rank, not graph-captured)" % (grid[0], grid[1], world),
                 "ms": t_ms, "output_mpix_per_s": 16 * 256 * 256 / 1e6 / (t_ms * 1e-3)}
        del mi_t
        G._plans.clear()
        torch.cuda.empty_cache()

    # BASELINE config 3: Z optimisation, 1x3x256x256 LR, objective 'TV', Adam lr 0.1 (rank 0 only, bounded)
    zopt = None
    if rank == 0 and not args.no_zopt:
        from esr_b200.z_optimization import Z_optimizer, SRModelShim
        zh = 256
        lr3, z3 = synth.make_inputs(1, zh, zh, seed=3)
        model = SRModelShim(netG)
        data = {"LR": lr3.to(dev), "Z": torch.zeros_like(z3).to(dev)}
        model.feed_data(data)
        with torch.no_grad():
            model.fake_H = netG(model.model_input)
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            zo = Z_optimizer(objective="TV", Z_size=[SF * zh, SF * zh], model=model, Z_range=1.0, max_iters=2, data=data,
                             initial_LR=0.1, batch_size=1)
            zo.optimize()                                     # warm-up: builds plans, packs dgrad weights
            torch.cuda.synchronize()
            n_it = 10
            zo.max_iters = n_it
            z0e, z1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            z0e.record()
            zo.optimize()
            z1e.record()
            torch.cuda.synchronize()
        z_ms = z0e.elapsed_time(z1e) / n_it
        hp3 = zh + 2 * MARGIN
        zflops = 2 * G.engine().flops_per_lr_pixel() * hp3 * hp3
        zopt = {"config": "1x3x256x256 LR, objective TV, Adam lr 0.1, eval/pre-pad (BASELINE config 3)",
                "iters_per_s": 1e3 / z_ms, "ms_per_iter": z_ms, "iters_timed": n_it, "loss_first_last": [zo.loss_values[0], zo.loss_values[-1]],
                "algorithmic_tflops": zflops / (z_ms * 1e-3) / 1e12}
        del zo, model
        G._plans.clear()
        G._bplans.clear()
        torch.cuda.empty_cache()

    ms, e2e_ms, conv_ms = max_over_ranks([ms, e2e_ms, conv_ms])
    n_total = per_rank * world if args.scaling == "weak" else BATCH
    out_mpix = n_total * SF * SF * LR_H * LR_W / 1e6              # whole job, all ranks
    pk = peaks()
    flops = G.engine().flops_per_lr_pixel() * per_rank * (LR_H + 2 * MARGIN) * (LR_W + 2 * MARGIN)
    achieved = flops / (conv_ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": out_mpix / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling,
        "vs_baseline": None, "dtype": "bf16 MMA operands, fp32 accumulate/trunk/CEM", "data": "synthetic",
        "config": workload_config(world, per_rank if args.scaling == "weak" else BATCH // world, graph_used, SUB, NSTREAMS),
        "e2e": {"value": out_mpix / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": args.steps * (per_rank // SUB) * launches_per_fwd,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16"], "unit": "TFLOP/s",
                     "frac": achieved / pk["bf16"], "traffic": conv_traffic(), "peak_source": pk["src"] + " sustained bf16",
                     "kernel": "conv3x3 tcgen05 kernels (pair::conv3x3_tc2_kernel / conv3x3_rdb_growth_kernel, cta_group::2, + "
                               "conv3x3_tc_kernel for the last conv): %.3f ms of the %.3f ms step (share %.4f from step and "
                               "conv-only graphs replayed alternately in one window); algorithmic %.1f GFLOP/step; traffic = DRAM "
                               "bytes of those launches per step (ncu, profiles/)" % (conv_ms, ms, share, flops / 1e9),
                     "frac_of_burst_peak": achieved / pk["bf16_burst"]},
        "phases_ms": phases,
        "clocks": sampler.summary(),
    }
    if args.scaling == "strong":
        line["config"]["global_batch"] = BATCH
    if strong is not None:
        line["strong"] = strong
    if tiled is not None:
        line["tiled"] = tiled
    if zopt is not None:
        line["zopt"] = zopt
    if rank == 0 and not args.no_zopt:
        line["cem_roofline"] = cem_standalone(dev, pk)
        line["config1"] = config1_latency(netG, dev)
    if not args.no_train:
        # BASELINE config 5 beside the headline (every rank takes part: the step all-reduces its gradients over NCCL)
        try:
            G._plans.clear()
            torch.cuda.empty_cache()
            with contextlib.redirect_stdout(sys.stderr):         # init_weights prints like the reference's: keep stdout to ONE line
                tr = gan_step_measure(dev, rank, world, 10, 4)
            line["train_gan"] = {k: tr[k] for k in ("metric", "value", "unit", "ms_per_step", "n_gpus", "dtype", "config", "losses_last")}
        except Exception as e:                                 # never at the expense of the headline line
            line["train_gan"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_img = int(os.environ.get("ESR_BENCH_CPU_IMAGES", 8))
            fwd, kind = reference_forward_fn(n_img, LR_H, LR_W, cores)
            small, _ = reference_forward_fn(1, 32, 32, cores) if kind == "port" else (None, None)
            if small is not None:
                small()
            tc = timed(fwd)
            line["cpu_baseline"] = {"value": n_img * SF * SF * LR_H * LR_W / 1e6 / tc, "unit": UNIT, "cores": cores,
                                    "kind": kind, "sample": "%d of the step's 16 images (3x128x128 LR, eval/pre-pad) as one batch "
                                                            "through %s, fp32, 1 pass, %.1f s" % (
                                                                n_img, "the reference's own modules (oracle/_ref)" if kind == "reference"
                                                                else "the oracle port", tc)}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

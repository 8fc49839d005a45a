#!/usr/bin/env python
"""Benchmark of the RRDBNet(+Z) -> CEM hot path (BASELINE.json metric: x4 SR output Mpix/s).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # the reference algorithm on host cores (oracle port)

One "step" = one G+CEM forward over a batch of 16 synthetic 3x128x128 LR images + random Z in eval mode
(CEM pre-pad by 10 px), BASELINE config 2.  N > 1: one process per GPU (torchrun), every rank runs its own
batch of 16 (batch sharding, no data-path collective: images are independent), scaling = weak.
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


def oracle_step_seconds(n_images, h, w, threads):
    """Reference algorithm (CPU oracle, fp32) on `n_images` images of the workload; returns seconds."""
    from esr_b200 import synth
    from oracle.cem_ops import concat_latent
    from oracle.rrdbnet import GCEMOracle
    torch.set_num_threads(threads)
    wts = synth.make_weights("default", seed=0)
    lr, z = synth.make_inputs(n_images, h, w, seed=0)
    net = GCEMOracle(wts)
    mi = concat_latent(lr, z)
    t0 = time.perf_counter()
    with torch.no_grad():
        net.forward(mi)
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
    """BASELINE config 4: CEM wrapping a 2048x2048 x4 SR output (bicubic), HBM roofline, L2 flushed per iteration."""
    from esr_b200 import _capi as capi, cem as pcem
    f = pcem.CEMnet(pcem.Get_CEM_Config(SF))._filters
    B, C, H, W = 1, 3, 2048, 2048
    y = torch.rand(B, C, H, W, device=dev)
    x = torch.rand(B, C, H // SF, W // SF, device=dev)
    out = torch.empty_like(y)
    ws = torch.empty(2 * B * C * (H // SF) * (W // SF), device=dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    l = capi.lib()

    def run():
        capi.check(l.esr_cem_project(f, capi.ptr(y), capi.ptr(x), B, C, H, W, 0, capi.ptr(out), capi.ptr(ws), capi.stream_ptr()))
    for _ in range(3):
        run()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    byts = 4 * C * (2 * H * W + H * W // (SF * SF))           # read y, read x, write out = 24.75 B / HR px
    return {"bound": "hbm", "achieved": byts / t / 1e6, "peak": pk["hbm"], "unit": "GB/s", "frac": byts / t / 1e6 / pk["hbm"],
            "kernel": "cem_down4_kernel + cem_invup4_kernel (2 launches, the second a programmatic dependent of the first), "
                      "1x3x2048x2048 output, %.1f us, algorithmic %.1f MB; L2 flushed between iterations; traffic = DRAM bytes "
                      "of both launches (ncu, cold cache per launch, profiles/)" % (t * 1e3, byts / 1e6),
            "traffic": conv_traffic("cem_cfg4_dram_bytes")}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_img = 4
    for _ in range(max(args.warmup, 0) and 1):
        oracle_step_seconds(n_img, 32, 32, cores)            # warm the thread pool / allocator, small
    times = [oracle_step_seconds(n_img, LR_H, LR_W, cores) for _ in range(max(1, min(args.steps, 3)))]
    t = min(times)
    mpix = n_img * SF * SF * LR_H * LR_W / 1e6 / t
    sample = "%d of %d images of the batch per step (3x%dx%d LR each, eval/pre-pad), best of %d" % (
        n_img, BATCH, LR_H, LR_W, len(times))
    line = {"impl": "reference", "metric": METRIC, "value": mpix, "unit": UNIT, "n_gpus": args.gpus,
            "steps": len(times), "warmup": args.warmup, "ms_per_step": t * 1e3 * BATCH / n_img,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD},
            "cpu_baseline": {"value": mpix, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": mpix, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-zopt", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank, local_rank, world = dist_env()
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)

    from esr_b200 import _capi as capi, cem as pcem, networks, synth
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
    lr, z = synth.make_inputs(BATCH, LR_H, LR_W, seed=rank)
    host_in = torch.cat([z.contiguous().view(BATCH, 16 * 3, LR_H, LR_W), lr], 1).contiguous().pin_memory()
    x_dev = host_in.to(dev)
    host_out = torch.empty(BATCH, 3, SF * LR_H, SF * LR_W).pin_memory()

    SUB = int(os.environ.get("ESR_SUBBATCH", BATCH))       # images per pass through the layer sequence
    NSTREAMS = int(os.environ.get("ESR_STREAMS", 1))        # sub-batches in flight (each on its own stream / buffers)
    assert BATCH % SUB == 0 and (BATCH // SUB) % NSTREAMS == 0
    plans = [G.plan(SUB, LR_H, LR_W, MARGIN, keep=False, slot=i) for i in range(NSTREAMS)]
    plan = plans[0]
    filters = netG._filters
    out_dev = torch.empty(BATCH, 3, SF * LR_H, SF * LR_W, device=dev)
    wss = [torch.empty(2 * SUB * 3 * plan.hp * plan.wp, device=dev) for _ in range(NSTREAMS)]
    ws = wss[0]
    side = [torch.cuda.Stream() for _ in range(NSTREAMS)] if NSTREAMS > 1 else []
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    H4, W4 = SF * plan.hp, SF * plan.wp

    def step(record=False):
        l = capi.lib()
        main = torch.cuda.current_stream()
        for s_ in side:
            s_.wait_stream(main)
        for k, i0 in enumerate(range(0, BATCH, SUB)):
            first = i0 == 0
            pl, w_ = plans[k % NSTREAMS], wss[k % NSTREAMS]
            with torch.cuda.stream(side[k % NSTREAMS]) if side else contextlib.nullcontext():
                if record and first:
                    ev[0].record()
                pl.run_prep(x_dev[i0:i0 + SUB])
                if record and first:
                    ev[1].record()
                pl.run_convs()
                if record and first:
                    ev[2].record()
                capi.check(l.esr_cem_project(filters, capi.ptr(pl.y), capi.ptr(pl.lr_pad), SUB, 3, H4, W4, SF * MARGIN,
                                             capi.ptr(out_dev[i0:i0 + SUB]), capi.ptr(w_), capi.stream_ptr()))
                if record and first:
                    ev[3].record()
        for s_ in side:
            main.wait_stream(s_)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
            torch.cuda.synchronize()

    launches_per_fwd = plan.launches_per_forward(with_cem=True)
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    # per-phase split (untimed iteration) for the roofline of the dominant kernel
    step(record=True)
    torch.cuda.synchronize()
    t_prep, t_conv, t_cem = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])

    graph = None
    if not args.no_graph:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            step()
        torch.cuda.current_stream().wait_stream(s)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        for _ in range(2):
            graph.replay()
    run = graph.replay if graph is not None else step

    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        run()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    # conv-only timing in the same regime as the step (CUDA-graph replay of the recorded conv sequence, events)
    def convs_only():
        for _i in range(BATCH // SUB):
            plans[_i % NSTREAMS].run_convs()
    conv_run = convs_only
    if graph is not None:
        cg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(cg):
            convs_only()
        conv_run = cg.replay
    for _ in range(2):
        conv_run()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(args.steps):
        conv_run()
    c1.record()
    torch.cuda.synchronize()
    conv_ms = c0.elapsed_time(c1) / args.steps

    # end to end through the public module API with host buffers: netG(...) on pinned host input, result copied
    # back to pinned host memory; parallel.HostPipeline overlaps the PCIe copies of one half batch with the
    # other half's compute (both copies of every step are inside the timed region)
    from esr_b200.parallel import HostPipeline
    pipe = HostPipeline(netG, chunk=int(os.environ.get("ESR_E2E_CHUNK", BATCH)))

    def e2e_step():
        pipe(host_in, host_out)
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

    # BASELINE config 3: Z optimisation, 1x3x256x256 LR, objective 'TV', Adam lr 0.1 (rank 0 only, bounded)
    zopt = None
    if rank == 0 and not args.no_zopt:
        from esr_b200.z_optimization import Z_optimizer, SRModelShim
        del plan, plans, out_dev, ws, wss
        G._plans.clear()
        torch.cuda.empty_cache()
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

    t = torch.tensor([ms, e2e_ms, conv_ms], device=dev, dtype=torch.float64)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms, e2e_ms, conv_ms = [float(v) for v in t.cpu()]
    out_mpix = BATCH * SF * SF * LR_H * LR_W / 1e6
    pk = peaks()
    flops = G.engine().flops_per_lr_pixel() * BATCH * (LR_H + 2 * MARGIN) * (LR_W + 2 * MARGIN)
    achieved = flops / (conv_ms * 1e-3) / 1e12
    line = {
        "metric": METRIC, "value": world * out_mpix / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16 MMA operands, fp32 accumulate/trunk/CEM", "data": "synthetic",
        "config": {"workload": WORKLOAD, "global_batch": BATCH * world, "parallelism": "batch shard x%d" % world,
                   "l2": "per-step working set (~4.5 GB of activations) >> 126 MB L2, no flush needed",
                   "cuda_graph": graph is not None, "images_per_pass": SUB, "passes_in_flight": NSTREAMS},
        "e2e": {"value": world * out_mpix / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": host_in.numel() * 4, "d2h_bytes_per_step": host_out.numel() * 4},
        "gpu_launches": args.steps * (BATCH // SUB) * launches_per_fwd,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["bf16"], "unit": "TFLOP/s",
                     "frac": achieved / pk["bf16"], "traffic": conv_traffic(), "peak_source": pk["src"] + " sustained bf16",
                     "kernel": "conv3x3 tcgen05 kernels: pair::conv3x3_tc2_kernel (cta_group::2) x350 + conv3x3_tc_kernel x1 "
                               "= 351 launches/step, %.3f ms/step; algorithmic %.1f GFLOP/step; traffic = DRAM bytes of "
                               "those 351 launches per step (ncu, profiles/)" % (conv_ms, flops / 1e9),
                     "frac_of_burst_peak": achieved / pk["bf16_burst"]},
        "phases_ms": {"prep": t_prep, "convs": t_conv, "cem": t_cem},
        "clocks": sampler.summary(),
    }
    if zopt is not None:
        line["zopt"] = zopt
    if rank == 0 and not args.no_zopt:
        line["cem_roofline"] = cem_standalone(dev, pk)
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            oracle_step_seconds(1, 32, 32, cores)
            tc = oracle_step_seconds(BATCH, LR_H, LR_W, cores)
            line["cpu_baseline"] = {"value": BATCH * SF * SF * LR_H * LR_W / 1e6 / tc, "unit": UNIT, "cores": cores,
                                    "kind": "port", "sample": "one full step: the 16 images (3x128x128 LR, eval/pre-pad) as one "
                                                              "batch through the fp32 oracle port, 1 pass, %.1f s" % tc}
        print(json.dumps(line))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()

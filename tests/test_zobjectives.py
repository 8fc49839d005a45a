"""Host logic of the rich Z objectives (esr_b200.z_objectives) on CPU.

The density sums are CUDA kernels (csrc/zobj.cu) with no CPU path, so these tests put the oracle's restatement
(oracle/zobjectives.kde_sums) in their place and check everything around them - patch tables and the native greedy
selection, bin pruning, DC / STD normalisation, histogram normalisers, the objective classes, the Z_optimizer wiring -
against the committed goldens (recorded from the unmodified reference) and, when /root/reference exists, against the
reference's own classes run side by side.  tests/test_gpu_zobjectives.py checks the kernels themselves.

The second half of the file runs Z_optimizer's whole host loop on CPU - the built-in objectives included - around the
CPU oracle of G+CEM against the trajectories recorded from the reference (zopt.npz, zopt2.npz, zobjectives.npz), and the
GUI's calling patterns (repeated rounds, random initialisations, loggers, the training-mode search) side by side with the
reference's class."""
import warnings

import numpy as np
import pytest
import torch

from esr_b200 import _capi as capi, synth, z_objectives as zo
from esr_b200.z_optimization import Z_optimizer, SRModelShim
from oracle import ref_shims, zobj_cases as zc, zobjectives as oracle_zobj

warnings.filterwarnings("ignore")


@pytest.fixture()
def oracle_density(monkeypatch):
    monkeypatch.setattr(zo, "kde_sums", oracle_zobj.kde_sums)


@pytest.fixture(scope="module")
def reference_zopt():
    if not ref_shims.available():
        pytest.skip("reference tree not present")
    zopt = ref_shims.load_reference()[3]
    return zopt


def test_density_kernels_have_no_cpu_path():
    with pytest.raises(capi.EsrError):
        zo.kde_sums(torch.rand(1, 8), torch.rand(1, 4).double(), 1.0, 1e-3)


@pytest.mark.parametrize("name", sorted(zc.HIST_CASES))
def test_soft_histogram_loss_matches_reference_golden(golden, oracle_density, name):
    """Values, image gradients and the pruned bins of SoftHistogramLoss (Z_optimization.py:21-228) for the gray / patch,
    histogram / dictionary, DC- and STD-free variants Z_optimizer builds."""
    g = golden("zobjectives")
    value, grad, bins = zc.run_hist_case(zo.SoftHistogramLoss, name)
    np.testing.assert_array_equal(bins.numpy(), g["hist_%s_bins" % name])
    np.testing.assert_allclose(value.numpy().astype(np.float64), g["hist_%s_value" % name], rtol=1e-5, atol=1e-7)
    ref = g["hist_%s_grad" % name]
    np.testing.assert_allclose(grad.numpy(), ref, rtol=1e-4, atol=1e-6 * np.abs(ref).max())


@pytest.mark.parametrize("patch,overlap", [(7, 1), (7, 0.5), (6, 30 / 36), (6, 0.5), (3, 0)])
def test_patch_tables_equal_reference_extraction_matrix(reference_zopt, patch, overlap):
    """The index table + native greedy selection against ReturnPatchExtractionMat's sparse matrix (:230-270), including
    its non-covered pixel set, on a mask with a hole and ragged borders."""
    H, W = 23, 31
    mask = np.ones((H, W), dtype=bool)
    mask[:2, :] = False
    mask[9:12, 10:14] = False
    mask[:, -1] = False
    mat, rest = reference_zopt.ReturnPatchExtractionMat(mask.copy(), patch, torch.device("cpu"), patches_overlap=overlap,
                                                        return_non_covered=True)
    table, mine_rest = zo.patch_tables(mask.copy(), patch, torch.device("cpu"), overlap, return_non_covered=True)
    img = torch.from_numpy(np.random.default_rng(0).random(H * W).astype(np.float32))
    ref = torch.sparse.mm(mat, img.view(-1, 1)).view(patch ** 2, -1)
    np.testing.assert_array_equal(table.extract(img).numpy(), ref.numpy())
    if rest is None:
        assert mine_rest is None
    else:
        ref_rest = torch.sparse.mm(rest, img.view(-1, 1)).view(-1)
        np.testing.assert_array_equal(np.sort(mine_rest.extract(img).view(-1).numpy()), np.sort(ref_rest.numpy()))


def test_patch_select_validates_arguments():
    l = capi.lib()
    assert l.esr_patch_select(None, 1, 4, 0.5, 0, 10, None, None) == -1
    px = np.array([[0, 1, 50, 3]], dtype=np.int64)                 # 50 is outside [min, min + span]
    valid, covered = np.zeros(1, np.uint8), np.zeros(10, np.uint8)
    import ctypes as C
    rc = l.esr_patch_select(px.ctypes.data_as(C.c_void_p), 1, 4, 0.5, 0, 10, valid.ctypes.data_as(C.c_void_p),
                            covered.ctypes.data_as(C.c_void_p))
    assert rc == -1 and b"outside" in l.esr_last_error()
    assert l.esr_kde_workspace_bytes(0, 5) == -1
    assert l.esr_kde_workspace_bytes(100, 100000) > 0              # few own vectors: the other range is split
    assert l.esr_kde_workspace_bytes(1 << 20, 256) == 0
    # the gradient pass splits the bin range the same way: fp64 partials of D x N values per split
    assert l.esr_kde_grad_workspace_bytes(0, 5, 1) == -1
    assert l.esr_kde_grad_workspace_bytes(1 << 20, 256, 1) == 0
    few = l.esr_kde_grad_workspace_bytes(3600, 4000, 36)
    assert few > 0 and few % (36 * 3600 * 8) == 0 and few // (36 * 3600 * 8) == l.esr_kde_workspace_bytes(3600, 4000) // (3600 * 8)
    # compute entry points validate before touching the device (no GPU here): null pointers, D out of range, bad temperature
    assert l.esr_kde_sums(None, 0, 10, None, 1, 10, 1, 1.0, 1e-3, 1e-7, None, None, None) == -1
    assert l.esr_kde_grad(None, 10, None, 10, 1, 1.0, 1e-3, 1e-7, None, None, None, None, None) == -1


def test_hsv_round_trip_and_known_colours():
    rgb = np.random.default_rng(1).random((9, 11, 3)) * 255
    np.testing.assert_allclose(zo.hsv2rgb(zo.rgb2hsv(rgb)), rgb, rtol=1e-12, atol=1e-9)
    known = np.array([[[255, 0, 0], [0, 255, 0], [0, 0, 255], [128, 128, 128], [0, 0, 0], [255, 255, 0]]], dtype=np.float64)
    hsv = zo.rgb2hsv(known)[0]
    np.testing.assert_allclose(hsv[:, 0], [0, 1 / 3, 2 / 3, 0, 0, 1 / 6], atol=1e-12)
    np.testing.assert_allclose(hsv[:, 1], [1, 1, 1, 0, 0, 1], atol=1e-12)
    np.testing.assert_allclose(hsv[:, 2], [255, 255, 255, 128, 0, 255], atol=1e-12)


@pytest.mark.parametrize("name", [n for n in zc.ZOPT3_CASES if n not in zc.NO_REFERENCE_RUN])
def test_z_optimizer_objectives_match_reference_side_by_side(reference_zopt, oracle_density, name):
    """Every rich objective through this package's Z_optimizer against the reference's own Z_optimizer (:326-655), both
    around the same cheap stub generator on CPU: loss per iteration, per-image losses, returned Z."""
    from oracle.gen_golden import RefModel
    lr, z0 = synth.make_inputs(1, zc.ZOPT3_HW[0], zc.ZOPT3_HW[1], seed=11)
    out = []
    for cls, model_cls in ((reference_zopt.Z_optimizer, RefModel), (Z_optimizer, SRModelShim)):
        netG = zc.StubGenerator()
        opt, Z = zc.run_zopt_case(cls, model_cls(netG), netG, name, lr, z0, z_init=zc.zopt3_z_init(name))
        out.append((np.array(opt.loss_values), np.array(opt.latest_Z_loss_values).reshape(-1), Z))
    (l0, a0, Z0), (l1, a1, Z1) = out
    np.testing.assert_allclose(l1, l0, rtol=1e-5, atol=1e-8)
    np.testing.assert_allclose(a1, a0, rtol=1e-5, atol=1e-8)
    assert float((Z0 - Z1).abs().max()) < 1e-4


def test_scribble_objective_terms(oracle_density):
    """No reference run exists for 'scribble' (zobj_cases.NO_REFERENCE_RUN): the objective's value is checked against
    the formulas of Z_optimization.py:371-416 written out with explicit loops."""
    lr, z0 = synth.make_inputs(1, 8, 8, seed=11)
    netG = zc.StubGenerator()
    opt, Z = zc.run_zopt_case(Z_optimizer, SRModelShim(netG), netG, "scribble", lr, z0)
    assert len(opt.loss_values) == 3 and opt.loss_values[-1] < opt.loss_values[0]
    H = W = 32
    im_mask, _ = zc.region_masks(H, W)
    ids = zc.scribble_mask(H, W)
    fake = torch.from_numpy(np.random.default_rng(5).random((1, 3, H, W)).astype(np.float32))
    value = float(opt._rich(fake).reshape(-1)[0])
    l1_mask = im_mask * ((ids > 0) & (ids < 4))
    expect = float(np.abs(fake.numpy()[0] * l1_mask - opt.GT_HR.numpy()[0] * l1_mask).mean())
    f = fake.numpy()[0]
    for region in (4, 5):
        m = im_mask * (ids == region)
        for dy, dx in ((-1, -1), (-1, 0), (0, -1), (1, -1)):
            total = 0.0
            for y in range(H):
                for x in range(W):          # the two opposite crops pair pixel (y, x) with (y - dy, x - dx)
                    y2, x2 = y - dy, x - dx
                    if 0 <= y2 < H and 0 <= x2 < W:
                        total += m[y, x] * m[y2, x2] * np.abs(f[:, y, x] - f[:, y2, x2]).sum()
            expect += total / (3 * (H - abs(dy)) * (W - abs(dx)))
    assert abs(value - expect) < 1e-5 * max(1.0, abs(expect))
    # brighter / darker scribbles re-light the target: scaling V scales the colour, by the smoothed 1 +- brightness factor
    shim = SRModelShim(netG)
    shim.feed_data({"LR": lr, "Z": 0.5 * z0})
    with torch.no_grad():
        start = netG(shim.model_input).clamp(0, 1)[0]
    y, x = H // 2 + 1, W // 4 + 3                # inside the 'brighter' box, 3x3 neighbourhood all inside
    np.testing.assert_allclose(opt.GT_HR[0, :, y, x].numpy(), 1.2 * start[:, y, x].numpy(), rtol=1e-5)
    y, x = H // 2 + 1, W // 2 + 3                # inside the 'darker' box
    np.testing.assert_allclose(opt.GT_HR[0, :, y, x].numpy(), 0.8 * start[:, y, x].numpy(), rtol=1e-5)


def test_unbuilt_objectives_are_loud():
    for objective in ("Adversarial", "VGG", "desired_SVD", "random_VGG", "nonsense"):
        assert zo.unsupported_reason(objective) is not None
    assert zo.unsupported_reason("hist", auto_temperature=True) is not None
    for objective in ("hist", "patchdict_noDC", "local_STD_increase", "max_local_STD", "local_Mag_decrease", "scribble",
                      "nonInt_periodicity_1D", "local_STD_TV", "random_l1_limited", "l1"):
        assert zo.unsupported_reason(objective) is None, objective


@pytest.mark.parametrize("objective,max_iters,random_inits,bs", [("TV", 3, False, 1), ("max_STD", -2, False, 1),
                                                                  ("STD_increase", 2, True, 3), ("random_l1", 2, True, 3)])
def test_repeated_rounds_and_random_inits_match_reference(reference_zopt, objective, max_iters, random_inits, bs):
    """The GUI's calling pattern (GUI.py:1596-1660): one Z_optimizer object, optimize() called round after round (cur_iter
    and Adam's state carry over, :555-562, :645), the convergence window (max_iters < 0, :564-571), and random Z
    initialisations (Randomize_Z, :307-313; the RNG is torch's, so on CPU both sides draw the same numbers)."""
    from oracle.gen_golden import RefModel
    lr, z0 = synth.make_inputs(1, 8, 8, seed=4)
    out = []
    for cls, model_cls in ((reference_zopt.Z_optimizer, RefModel), (Z_optimizer, SRModelShim)):
        netG = zc.StubGenerator()
        model = model_cls(netG)
        data = {"LR": lr.repeat(bs, 1, 1, 1), "Z": (0.3 * z0).repeat(bs, 1, 1, 1), "STD_increment": 0.03}
        model.feed_data(data)
        with torch.no_grad():
            model.fake_H = netG(model.model_input)
        torch.manual_seed(11)
        opt = cls(objective=objective, Z_size=[32, 32], model=model, Z_range=1.0, max_iters=max_iters, data=data, initial_LR=0.05,
                  batch_size=bs, random_Z_inits=random_inits, initial_Z=0.3 * z0)
        rounds = []
        for _ in range(3):
            Z = opt.optimize()
            rounds.append((list(opt.loss_values), list(np.array(opt.latest_Z_loss_values).reshape(-1)), opt.cur_iter, Z.clone()))
        pre_tanh, optimizer = opt.ReturnStatus()
        out.append((rounds, pre_tanh.clone(), optimizer.state_dict()["state"][0]["exp_avg"].clone()))
    (r0, p0, m0), (r1, p1, m1) = out
    for (l0, a0, c0, Z0), (l1, a1, c1, Z1) in zip(r0, r1):
        assert len(l0) == len(l1) and c0 == c1
        np.testing.assert_allclose(l1, l0, rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(a1, a0, rtol=1e-5, atol=1e-8)
        assert float((Z0 - Z1).abs().max()) < 1e-5
    assert float((p0 - p1).abs().max()) < 1e-4 and float((m0 - m1).abs().max()) < 1e-6


def test_feed_desired_hist_im_matches_reference(reference_zopt, oracle_density):
    """Z_optimizer.feed_data's histogram branch (Z_optimization.py:547-548 -> Feed_Desired_Hist_Im, :97-104): new desired
    images replace the target histograms of an existing loss."""
    img, desired, dmasks, im = zc.hist_inputs("hist")
    new_desired = torch.cat([zc.smooth_image(401, 1, *zc.HIST_HW), zc.smooth_image(402, 1, *zc.HIST_HW)], 0)      # [2, 3, H, W]
    out = []
    for cls in (reference_zopt.SoftHistogramLoss, zo.SoftHistogramLoss):
        loss = cls(bins=256, min=0, max=1, desired_hist_image=[d.clone() for d in desired], desired_hist_image_mask=[m.copy() for m in dmasks],
                   input_im_HR_mask=im.clone(), gray_scale=True, patch_size=1, temperature=5e-4)
        before = float(loss(img))
        loss.Feed_Desired_Hist_Im(new_desired.clone())
        x = img.clone().requires_grad_(True)
        after = loss(x)
        grad, = torch.autograd.grad(after, x)
        out.append((before, float(after), grad))
    (b0, a0, g0), (b1, a1, g1) = out
    assert abs(b0 - b1) <= 1e-6 * max(1.0, abs(b0)) and abs(a0 - a1) <= 1e-6 * max(1.0, abs(a0)) and abs(a0 - b0) > 1e-9
    np.testing.assert_allclose(g1.numpy(), g0.numpy(), rtol=1e-4, atol=1e-7 * float(g0.abs().max()))


@pytest.mark.parametrize("objective", ["l1", "dict"])
def test_training_mode_with_hr_unpadder_matches_reference(reference_zopt, oracle_density, objective):
    """The per-batch optimal-Z search of the training loop (SRRaGAN_model.py:144-147): HR_unpadder given, so every
    iteration's output is cropped before the loss (:575-576), no per-image values are kept (:621) and one more
    un-cropped forward follows the loop (:646-649)."""
    from oracle.gen_golden import RefModel
    lr, z0 = synth.make_inputs(2, 8, 8, seed=6)
    crop = lambda t: t[:, :, 4:-4, 4:-4]          # noqa: E731
    out = []
    for cls, model_cls in ((reference_zopt.Z_optimizer, RefModel), (Z_optimizer, SRModelShim)):
        netG = zc.StubGenerator()
        model = model_cls(netG)
        data = {"LR": lr, "Z": 0.2 * z0, "HR": crop(zc.smooth_image(77, 2, 32, 32))}
        if objective == "dict":
            data["HR"] = [zc.smooth_image(78, 1, 24, 24)]
            data["Desired_Im_Mask"] = [np.ones((24, 24), dtype=bool)]
        torch.manual_seed(5)
        opt = cls(objective=objective, Z_size=[32, 32], model=model, Z_range=1.0, max_iters=3, data=data, initial_LR=0.05,
                  batch_size=2, HR_unpadder=crop)
        Z = opt.optimize()
        out.append((list(opt.loss_values), Z, model.fake_H.detach().clone(), opt.cur_iter))
    (l0, Z0, f0, c0), (l1, Z1, f1, c1) = out
    assert c0 == c1 and tuple(f0.shape) == tuple(f1.shape) == (2, 3, 32, 32)          # the final forward is un-cropped
    np.testing.assert_allclose(l1, l0, rtol=1e-5, atol=1e-8)
    assert float((Z0 - Z1).abs().max()) < 1e-5 and float((f0 - f1).abs().max()) < 1e-5


def test_loggers_receive_the_same_records_as_in_the_reference(reference_zopt):
    """loggers (one per image, GUI.py:1590-1594): every iteration each logger gets its image's loss (:609-613); with
    loggers the host reads every value as it goes (no deferred reads)."""
    from oracle.gen_golden import RefModel

    class Recorder:
        def __init__(self):
            self.records = []

        def print_format_results(self, mode, rlt, dont_print=False):
            self.records.append((mode, rlt["iters"], rlt["lr"], rlt["Z_loss"], dont_print))

    lr, z0 = synth.make_inputs(1, 8, 8, seed=8)
    out = []
    for cls, model_cls in ((reference_zopt.Z_optimizer, RefModel), (Z_optimizer, SRModelShim)):
        netG = zc.StubGenerator()
        model = model_cls(netG)
        data = {"LR": lr.repeat(2, 1, 1, 1), "Z": (0.3 * z0).repeat(2, 1, 1, 1)}
        model.feed_data(data)
        with torch.no_grad():
            model.fake_H = netG(model.model_input)
        loggers = [Recorder(), Recorder()]
        opt = cls(objective="TV", Z_size=[32, 32], model=model, Z_range=1.0, max_iters=3, data=data, initial_LR=0.05, batch_size=2,
                  loggers=loggers, initial_Z=0.3 * z0)
        opt.random_Z_inits = False
        opt.Z_model.Z.data.copy_(zc.zopt3_z_init("random_l1")[:2])
        opt.optimize()
        out.append([lg.records for lg in loggers])
    for ref_records, records in zip(*out):
        assert len(ref_records) == len(records) == 3
        for a, b in zip(ref_records, records):
            assert a[:3] == b[:3] and a[4] == b[4] and abs(a[3] - b[3]) <= 1e-6 * max(1.0, abs(a[3]))


class _OracleG(torch.nn.Module):
    """The CPU oracle of G+CEM (oracle/rrdbnet.GCEMOracle, fp32) behind the nn.Module surface Z_optimizer expects."""

    def __init__(self, nb=2, seed=5):
        super().__init__()
        from oracle.rrdbnet import GCEMOracle
        self.anchor = torch.nn.Parameter(torch.zeros(1))            # Z_optimizer reads the device from a parameter
        self.oracle = GCEMOracle(synth.make_weights("default", seed=seed, nb=nb), nb=nb)

    def forward(self, x):
        return self.oracle.forward(x)


@pytest.mark.parametrize("name", [n for n in zc.ZOPT3_CASES if n not in zc.NO_REFERENCE_RUN])
def test_objectives_around_the_oracle_generator_match_reference_golden(golden, oracle_density, name):
    """The trajectories oracle/gen_golden.py recorded from the reference's Z_optimizer around the reference's nb = 2 G+CEM
    (tests/golden/zobjectives.npz), reproduced by this package's Z_optimizer around the CPU oracle of G+CEM: needs
    neither /root/reference nor a GPU.  tests/test_gpu_zobjectives.py runs the same cases through the CUDA generator."""
    g = golden("zobjectives")
    netG = _OracleG()
    lr, z0 = synth.make_inputs(1, zc.ZOPT3_HW[0], zc.ZOPT3_HW[1], seed=11)
    opt, Z = zc.run_zopt_case(Z_optimizer, SRModelShim(netG), netG, name, lr, z0, z_init=zc.zopt3_z_init(name))
    ref = g["zopt_%s_loss" % name]
    np.testing.assert_allclose(opt.initial_STD.numpy(), g["zopt_%s_initial_STD" % name], rtol=1e-4, atol=1e-6)
    np.testing.assert_allclose(np.array(opt.loss_values), ref, rtol=2e-4, atol=2e-4 * np.abs(ref).max())
    np.testing.assert_allclose(np.array(opt.latest_Z_loss_values).reshape(-1), g["zopt_%s_latest" % name], rtol=2e-4,
                               atol=2e-4 * np.abs(ref).max())
    assert float((Z - torch.from_numpy(g["zopt_%s_Z" % name])).abs().mean()) < 2e-3


ZOPT2 = [("max_std", "max_STD", 4, False, 1), ("min_std", "min_STD", 3, False, 1), ("std_increase", "STD_increase", 4, False, 1),
         ("std_decrease_mult", "STD_decrease", 3, False, 1), ("tv_converge", "TV", -3, False, 1), ("tv_masked", "TV", 4, True, 1),
         ("tv_batch2", "TV", 3, False, 2)]


@pytest.mark.parametrize("name,objective,max_iters,masked,bs", ZOPT2)
def test_built_in_objectives_around_the_oracle_generator_match_reference_golden(golden, name, objective, max_iters, masked, bs):
    """tests/golden/zopt2.npz (the reference's Z_optimizer on TV / STD objectives, convergence mode, masks, batch 2) through
    this package's generic loop around the CPU oracle of G+CEM: the host logic of a17-a19 without a GPU.  The GPU test of
    the same cases (tests/test_gpu_zopt.py) runs the CUDA generator and, where it applies, the single-graph loop."""
    g = golden("zopt2")
    netG = _OracleG()
    lr, z0 = synth.make_inputs(1, 8, 8, seed=5)
    model = SRModelShim(netG)
    data = {"LR": lr.repeat(bs, 1, 1, 1), "Z": (0.5 * z0).repeat(bs, 1, 1, 1)}
    if "increase" in objective or "decrease" in objective:
        data["STD_increment"] = None if name.endswith("_mult") else 0.02
    model.feed_data(data)
    with torch.no_grad():
        model.fake_H = netG(model.model_input)
    kw = {}
    if masked:
        im = np.zeros((32, 32), dtype=np.float32)
        im[8:24, 4:20] = 1
        zm = np.zeros((32, 32), dtype=np.float32)
        zm[4:28, :24] = 1
        kw = dict(image_mask=im, Z_mask=zm, initial_Z=0.5 * z0)
    opt = Z_optimizer(objective=objective, Z_size=[32, 32], model=model, Z_range=1.0, max_iters=max_iters, data=data,
                      initial_LR=0.1, batch_size=bs, **kw)
    np.testing.assert_allclose(opt.initial_STD.numpy(), g[name + "_initial_STD"], rtol=1e-4)
    if bs > 1:
        opt.random_Z_inits = False
        opt.Z_model.Z.data.copy_(torch.from_numpy(g[name + "_Zinit"]))
    Z = opt.optimize()
    ref = g[name + "_loss"]
    assert len(opt.loss_values) == len(ref) and opt.cur_iter == int(g[name + "_cur_iter"])
    rtol = 2e-2 if ("increase" in objective or "decrease" in objective) else 5e-4       # squared differences of nearly equal STDs
    np.testing.assert_allclose(np.array(opt.loss_values), ref, rtol=rtol, atol=1e-9)
    np.testing.assert_allclose(np.array(opt.latest_Z_loss_values), g[name + "_latest"], rtol=rtol, atol=1e-9)
    assert float((Z - torch.from_numpy(g[name + "_Z"])).abs().mean()) < 5e-3


def test_l1_training_mode_around_the_oracle_generator_matches_reference_golden(golden):
    """tests/golden/zopt.npz, the training-mode l1 search (no pre-pad, HR_unpadder given) around the CPU oracle."""
    g = golden("zopt")
    netG = _OracleG()
    netG.oracle.pre_pad = False                                  # netG.train(): CEM_PyTorch pads in eval mode only (CEMnet.py:170-181)
    lr, z0 = synth.make_inputs(1, 8, 8, seed=5)
    model = SRModelShim(netG)
    data = {"LR": lr, "Z": 0.5 * z0, "HR": torch.from_numpy(g["l1_train_target"])}
    model.feed_data(data)
    opt = Z_optimizer(objective="l1", Z_size=[32, 32], model=model, Z_range=1.0, max_iters=4, data=data, initial_LR=0.1,
                      batch_size=1, HR_unpadder=lambda t: t)
    opt.feed_data(data)
    opt.random_Z_inits = False
    opt.Z_model.Z.data.copy_(torch.from_numpy(g["l1_train_Zinit"]))
    Z = opt.optimize()
    np.testing.assert_allclose(np.array(opt.loss_values), g["l1_train_loss"], rtol=5e-4)
    assert float((Z - torch.from_numpy(g["l1_train_Z"])).abs().mean()) < 5e-3

"""The reference's OWN training step (SRRaGANModel.optimize_parameters, codes/models/SRRaGAN_model.py:307-547) on top of
this package's generator: discriminator, WGAN-GP penalty, range loss, optimisers and step logic are the reference's torch
code (from /root/reference or the vendored oracle/_ref), G + CEM forward, data gradient and weight gradients are this
package's kernels behind ``netG(model_input)`` / ``l_g_total.backward()``.

  python tools/train_ref_model.py --impl compat|reference [--device cuda|cpu] [--steps 3] [--nb 1] [--patch 64] [--batch 2]

Prints one JSON line: per-step log values, parameter-change norms and (first generator step) a few gradient norms.
TEST INFRASTRUCTURE (imports oracle/): used by tests/test_compat_reference_model.py."""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(args):
    import torch
    from oracle import ref_shims
    ref_shims.install(compat_first=args.impl == "compat")
    if args.impl == "reference" and args.device == "cpu":
        ref_shims.load_reference()
    import options.options as option
    from models import create_model
    if args.impl == "reference" and args.gan > 0:
        # the shipped define_D names Discriminator_VGG_128 but passes the nb= that only Discriminator_VGG_128_ takes
        # (networks.py:119 vs architecture.py:182,222): the reference arm gets the class its options mean
        import models.modules.architecture as ref_arch
        ref_arch.Discriminator_VGG_128 = ref_arch.Discriminator_VGG_128_
    opt = option.parse(os.path.join(ref_shims.REF_ROOT, "options", "train", "train_esrgan_CEM.json"), is_train=True)
    tmp = tempfile.mkdtemp()
    for k in ("root", "experiments_root", "models", "log", "val_images"):
        opt["path"][k] = os.path.join(tmp, k)
        os.makedirs(opt["path"][k], exist_ok=True)
    opt["gpu_ids"] = [0] if args.device == "cuda" else None
    opt["network_G"]["nb"] = args.nb
    opt["network_G"]["latent_channels"] = 3
    opt["network_D"]["nf"] = args.nf_d
    opt["datasets"]["train"].update(batch_size=args.batch, batch_size_4_grads_G=args.batch * args.accum, batch_size_4_grads_D=args.batch * args.accum,
                                    patch_size=args.patch)
    opt["train"].update(grad_accumulation_steps_G=args.accum, grad_accumulation_steps_D=args.accum, D_update_ratio=1, D_verification=None,
                        D_valid_Steps_4_G_update=0, pixel_weight=1e-2, lr_G=1e-4, lr_D=1e-4, resume=0, gan_weight=args.gan, highpass_weight=0, shift_invariant_weight=0)
    opt = option.dict_to_nonedict(opt)
    torch.manual_seed(0)
    model = create_model(opt)
    return model, opt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="compat")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--nb", type=int, default=1)
    ap.add_argument("--patch", type=int, default=64)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--accum", type=int, default=1)
    ap.add_argument("--gan", type=float, default=0.0, help="gan_weight; > 0 needs the reference's discriminator")
    ap.add_argument("--nf-d", type=int, default=64, help="network_D.nf")
    ap.add_argument("--weights", default="")
    ap.add_argument("--save-weights", default="")
    args = ap.parse_args()
    import torch
    model, opt = build(args)
    G = model.netG.module if hasattr(model.netG, "module") else model.netG
    D = None
    if model.D_exists:
        D = model.netD.module if hasattr(model.netD, "module") else model.netD
    res = {"impl": args.impl, "G_class": type(G.generated_image_model).__module__, "D_class": type(D).__module__ if D is not None else None,
           "model_file": sys.modules[type(model).__module__].__file__}
    if args.weights:                      # same initial G and D in both arms
        ck = torch.load(args.weights)
        G.load_state_dict(ck["G"])
        if D is not None:
            D.load_state_dict(ck["D"])
    if args.save_weights:
        torch.save({"G": G.state_dict(), "D": D.state_dict() if D is not None else None}, args.save_weights)
    gen = torch.Generator().manual_seed(1)
    sf = opt["scale"]
    g0 = {k: v.detach().clone() for k, v in G.generated_image_model.named_parameters()}
    d0 = {k: v.detach().clone() for k, v in D.named_parameters()} if D is not None else {}
    steps = []
    for it in range(args.steps):
        hr = torch.rand(args.batch, 3, args.patch, args.patch, generator=gen)
        lr = torch.nn.functional.avg_pool2d(hr, sf)
        z = 2 * torch.rand(args.batch, 3, args.patch, args.patch, generator=gen) - 1
        torch.manual_seed(100 + it)       # the gradient-penalty interpolation points
        model.feed_data({"LR": lr, "HR": hr, "Z": z.to(model.device)})
        model.optimize_parameters()
        ent = {"generator_step": bool(model.generator_step), "fake_H": list(model.fake_H.shape),
               "fake_mean": float(model.fake_H.detach().float().mean()), "fake_std": float(model.fake_H.detach().float().std())}
        if model.generator_step and "grads" not in res:
            res["grads"] = {k: float(p.grad.norm()) for k, p in G.generated_image_model.named_parameters() if p.grad is not None}
            res["grad_sample"] = {k: p.grad.detach().flatten()[:8].cpu().tolist() for k, p in G.generated_image_model.named_parameters()
                                  if p.grad is not None and (k.endswith("model.0.weight") or "RDB2.convs.2.0.weight" in k)}
            if args.save_weights:
                torch.save({k: p.grad.detach().cpu() for k, p in G.generated_image_model.named_parameters() if p.grad is not None},
                           args.save_weights + ".grads")
        steps.append(ent)
    res["steps"] = steps
    res["log"] = {k: [float(v[1]) for v in vals] for k, vals in model.log_dict.items() if len(vals) and k in
                  ("l_g_pix", "l_g_range", "l_g_gan", "l_d_real", "l_d_fake", "l_d_gp", "D_real", "D_fake")}
    res["G_change"] = float(sum((v.detach() - g0[k]).norm() ** 2 for k, v in G.generated_image_model.named_parameters()) ** 0.5)
    if D is not None:
        res["D_change"] = float(sum((v.detach() - d0[k]).norm() ** 2 for k, v in D.named_parameters()) ** 0.5)
    print("RESULT " + json.dumps(res))


if __name__ == "__main__":
    main()

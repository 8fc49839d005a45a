"""Repeats the two arms of tests/test_compat_reference_model.py::test_reference_gan_training_step_runs_on_b200_generator and
prints, per repetition, the generator-gradient agreement per parameter (worst first) and the logged losses: how much of the
disagreement is run-to-run noise of the torch critic (cuDNN, BatchNorm on a batch of 2) and how much is the engines'.
GPU box:  python tools/gan_step_check.py [repetitions]"""
import json
import os
import subprocess
import sys
import tempfile

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def arm(impl, extra):
    cmd = [sys.executable, os.path.join(ROOT, "tools", "train_ref_model.py"), "--impl", impl, "--device", "cuda", "--steps", "3",
           "--nb", "2", "--patch", "208", "--batch", "2", "--gan", "5e-3", "--nf-d", "16"] + extra
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")]
    if not line:
        raise RuntimeError(p.stdout[-2000:] + p.stderr[-2000:])
    return json.loads(line[-1][7:])


def compare(a, b):
    rows = []
    for k, want in a.items():
        have = b[k]
        rel = float((have - want).norm() / want.norm().clamp_min(1e-30))
        cos = float((have * want).sum() / (have.norm() * want.norm()).clamp_min(1e-30))
        rows.append((rel, cos, k))
    return sorted(rows, reverse=True)


def main():
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    tmp = tempfile.mkdtemp()
    w = os.path.join(tmp, "init.pth")
    ref = arm("reference", ["--save-weights", w])
    g_ref = torch.load(w + ".grads")
    print("reference log", {k: ref["log"][k] for k in ("l_d_gp", "l_g_gan", "l_g_pix")})
    w2 = os.path.join(tmp, "ref2.pth")
    arm("reference", ["--weights", w, "--save-weights", w2])
    print("reference vs reference (run-to-run):", [(round(r, 4), round(c, 5), k) for r, c, k in compare(g_ref, torch.load(w2 + ".grads"))[:4]])
    for i in range(reps):
        out = os.path.join(tmp, "compat%d.pth" % i)
        got = arm("compat", ["--weights", w, "--save-weights", out])
        rows = compare(g_ref, torch.load(out + ".grads"))
        print("compat run", i, "log", {k: got["log"][k] for k in ("l_d_gp", "l_g_gan", "l_g_pix")})
        print("  worst:", [(round(r, 4), round(c, 5), k) for r, c, k in rows[:6]])
        print("  median rel:", round(rows[len(rows) // 2][0], 4))


if __name__ == "__main__":
    main()

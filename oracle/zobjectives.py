"""CPU restatement of the pairwise density arithmetic of the reference's SoftHistogramLoss.ComputeSoftHistogram
(/root/reference/codes/Z_optimization.py:184-195) - TEST INFRASTRUCTURE, the checker for csrc/zobj.cu's esr_kde_sums /
esr_kde_grad.  Only tests/ may import this; the product (z_objectives.kde_sums) runs the CUDA kernels and raises on CPU
tensors.

Pinned: tests/test_zobjectives.py runs the product's SoftHistogramLoss / objective classes with THIS function in place of
the kernels against the unmodified reference classes (live when /root/reference exists) and against
tests/golden/zobjectives.npz, which oracle/gen_golden.py recorded from the unmodified reference."""
import torch


def kde_sums(samples, bins, period, temperature, eps=1e-7, per_bin=False):
    """samples [D, N] (any float dtype, differentiable), bins [D, M] -> fp64 sums over bins per sample ([N]) or over
    samples per bin ([M]) of exp(mean_d(-(wrapped distance + eps)^2 / temperature)).
    :184 `image.unsqueeze(-1).double()`, :185-187 the three-way min (distance on a circle of length `period` = self.max),
    :189 `-((hist + SQRT_EPSILON) ** 2) / temperature`, :190 mean over the D values, :194 / :195 exp and the sum."""
    image = samples.double().unsqueeze(-1)                     # [D, N, 1]
    b = bins.double().unsqueeze(1)                             # [D, 1, M]
    dist = (image - b).abs()
    dist = torch.min(dist, (image - b - period).abs())
    dist = torch.min(dist, (image - b + period).abs())
    logits = (-((dist + eps) ** 2) / temperature).mean(0)      # [N, M]
    return torch.exp(logits).sum(0 if per_bin else 1)

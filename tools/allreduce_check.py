"""torchrun --nproc-per-node 2 tools/allreduce_check.py: the bucketed NCCL all-reduce (AVG) of GeneratorTrainer.backward.
Both ranks feed the same data, rank 1 scales its output gradient by 4 (a power of two: exact through the bf16 gradient
chain): the averaged weight gradients must be 2.5 times the gradients of a local backward."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
dev = torch.device("cuda", local)
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
from esr_b200 import synth
from esr_b200.training import GeneratorTrainer
from oracle.cem_ops import concat_latent
from tests.test_gpu_net import build_product_G
wts = synth.make_weights("kaiming", seed=0, nb=3)
netG = build_product_G(dev, 3, "all_layers_HR_downscaled", wts, train=True)
tr = GeneratorTrainer(netG, bucket_bytes=1 << 20)
lr, z = synth.make_inputs(4, 32, 32, seed=0)
mi = concat_latent(lr, z).to(dev)
g = torch.randn(4, 3, 128, 128, generator=torch.Generator().manual_seed(1)).to(dev)
tr.forward(mi)
tr.backward(g, all_reduce=False)
local_flat = tr.flat.clone()
tr.forward(mi)
tr.backward(g * (4.0 if rank == 1 else 1.0), all_reduce=True)
torch.cuda.synchronize()
err = float((tr.flat - 2.5 * local_flat).norm() / (2.5 * local_flat).norm())
print("rank %d: %d buckets, groups %s, relative difference %.3e" % (rank, len(tr.buckets), tr._group_cuts, err))
assert err < 1e-5, err
dist.destroy_process_group()

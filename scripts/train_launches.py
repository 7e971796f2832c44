"""One eager training step (BASELINE config 4 shape) between cudaProfilerStart/Stop, for an ncu launch list:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python scripts/train_launches.py
Development aid."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import manipose_b200 as mb  # noqa: E402
from manipose_b200 import metrics  # noqa: E402
from manipose_b200.optim import FusedAdam  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 27
B = int(sys.argv[2]) if len(sys.argv) > 2 else 30
dev = torch.device("cuda")
torch.manual_seed(42)
model = mb.RMCLManifoldMixSTE(mb.h36m17_skeleton(), num_frame=T, n_hyp=5, drop_path_rate=0.1).to(dev).train().set_compute_dtype("bf16")
opt = FusedAdam(model, lr=4e-5, weight_decay=1e-6)
g = torch.Generator().manual_seed(1234)
x = (0.3 * torch.randn(B, T, 17, 2, generator=g)).to(dev)
y = (0.3 * torch.randn(B, T, 17, 3, generator=g)).to(dev)


def step():
    opt.zero_grad()
    poses, scores = model(x)
    loss, _ = metrics.losses.training_loss(poses, scores, y)
    loss.backward()
    opt.step()
    return loss


for _ in range(4):
    step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("ok")

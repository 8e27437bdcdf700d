"""ncu target: the first Conv3D's weight gradient (conv_wgrad_kernel) and the second Conv3D's data gradient (a halo plan on
dy) at the bench batch of 8 sweeps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lisec_b200.train import ConvDgrad, ConvWgrad

dev = "cuda"
x = torch.randn((8, 8, 200, 400, 64), device=dev).to(torch.bfloat16)
dy = torch.randn((8, 4, 200, 400, 64), device=dev).to(torch.bfloat16)
wg = ConvWgrad(x, dy, (3, 3, 3), 2, (1, 1, 1))
for _ in range(3):
    wg.run()
torch.cuda.synchronize()
dy2 = torch.randn((8, 2, 200, 400, 64), device=dev).to(torch.bfloat16)
dg = ConvDgrad(dy2, torch.randn((27, 64, 64), device=dev) * 0.05, (3, 3, 3), (0, 1, 1))
for _ in range(3):
    dg.run()
torch.cuda.synchronize()
print("ok")

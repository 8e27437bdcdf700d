"""The first Conv3D's weight gradient at the bench batch (8 sweeps: x [8,8,200,400,64], dy [8,4,200,400,64], 3x3x3, stride
2 in depth), timed with CUDA events; LISEC_WGRAD_HALO=0 selects the per-tap kernel.    python tools/wgrad_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lisec_b200.train import ConvWgrad  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for name, shape_x, shape_dy, sd, pad in (("conv3d   (stride_d 2)", (B, 8, 200, 400, 64), (B, 4, 200, 400, 64), 2, (1, 1, 1)),
                                         ("conv3d_1 (stride_d 1)", (B, 4, 200, 400, 64), (B, 2, 200, 400, 64), 1, (0, 1, 1))):
    x = torch.randn(shape_x, device="cuda").to(torch.bfloat16)
    dy = torch.randn(shape_dy, device="cuda").to(torch.bfloat16)
    wg = ConvWgrad(x, dy, (3, 3, 3), sd, pad)
    for _ in range(3):
        wg.run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        wg.run()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    fl = 2.0 * shape_dy[0] * shape_dy[1] * shape_dy[2] * shape_dy[3] * 27 * 64 * 64
    print("%s halo=%s: %.3f ms (kernel + slice reduction), %.0f TFLOP/s" % (name, os.environ.get("LISEC_WGRAD_HALO", "1"), ms,
                                                                           fl / ms / 1e9))
    wg.close()
    del x, dy, wg
    torch.cuda.empty_cache()

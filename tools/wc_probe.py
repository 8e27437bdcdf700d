"""H2D copies of the front end's 9.6 MB of points from ordinary pinned memory against WRITE-COMBINED pinned memory
(cudaHostAllocWriteCombined: the host only ever writes these buffers), alone and in three 3.2 MB pieces."""
import ctypes as C
import warnings

import numpy as np
import torch

warnings.simplefilter("ignore")
from cuda import cudart  # noqa: E402


def timed(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


torch.cuda.init()
n = 9_600_000
d = torch.empty(n, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for name, flags in (("pinned", cudart.cudaHostAllocDefault), ("write-combined", cudart.cudaHostAllocWriteCombined)):
    err, ptr = cudart.cudaHostAlloc(n, flags)
    assert err == cudart.cudaError_t.cudaSuccess, err
    C.memset(ptr, 1, n)

    def whole():
        cudart.cudaMemcpyAsync(d.data_ptr(), ptr, n, cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, st)

    def pieces():
        for k in range(3):
            cudart.cudaMemcpyAsync(d.data_ptr() + k * (n // 3), ptr + k * (n // 3), n // 3,
                                   cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, st)

    a, b = timed(whole), timed(pieces)
    print("%-15s one copy %.3f ms %5.1f GB/s | three pieces %.3f ms %5.1f GB/s" % (name, a, n / a / 1e6, b, n / b / 1e6))
    cudart.cudaFreeHost(ptr)

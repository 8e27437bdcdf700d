"""Host <-> device copy bandwidth on this box, by size (pinned host memory, CUDA events), and a 9.6 MB H2D split over
two streams — what bounds the end-to-end front-end step (bench.py `e2e`)."""
import torch


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for mb in (2.4, 4.8, 9.6, 12.0, 16.0, 19.2, 38.4, 256):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    ms = timed(lambda: d.copy_(h, non_blocking=True))
    ms2 = timed(lambda: h.copy_(d, non_blocking=True))
    print("%6.1f MB: H2D %.3f ms %5.1f GB/s | D2H %.3f ms %5.1f GB/s" % (mb, ms, n / ms / 1e6, ms2, n / ms2 / 1e6))

n = 9_600_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
ev = torch.cuda.Event()


def split():
    cur = torch.cuda.current_stream()
    ev.record(cur)
    for s, (a, b) in ((s1, (0, n // 2)), (s2, (n // 2, n))):
        s.wait_event(ev)
        with torch.cuda.stream(s):
            d[a:b].copy_(h[a:b], non_blocking=True)
        cur.wait_stream(s)


ms = timed(split)
print("9.6 MB H2D as two halves on two streams: %.3f ms %5.1f GB/s" % (ms, n / ms / 1e6))


# the same 9.6 MB as SEQUENTIAL pieces on one stream (the per-copy rate above is not monotonic in the size)
for piece in (1_200_000, 2_400_000, 3_200_000, 4_800_000):
    def chunks(piece=piece):
        for a in range(0, n, piece):
            d[a:a + piece].copy_(h[a:a + piece], non_blocking=True)
    ms = timed(chunks)
    print("9.6 MB H2D as sequential pieces of %.1f MB: %.3f ms %5.1f GB/s" % (piece / 1e6, ms, n / ms / 1e6))
# and with the host buffer freshly written by the CPU before every copy (as a real producer would leave it)
import numpy as np
hn = h.numpy()
def fresh():
    hn[::4096] += 1
    d.copy_(h, non_blocking=True)
print("9.6 MB H2D, host buffer touched before each copy: %.3f ms" % timed(fresh))

"""What a call leaves behind for the next one (voxelize.cu / vfe.cu): no call starts with a memset — the point pass's drop
counters are moved and zeroed by scan_down, the chunk marks and the scans' group totals are put back by the order pass,
and the fused kernel's background writers claim their cells from a counter that the last writer warp zeroes. These tests
drive the sequences that would go wrong if one of them were left dirty: the fused stage several times on one
voxelization, batches of different sizes in alternation (a different number of chunks, scan blocks and writer batches
every call), clouds with and without dropped points, and a replay from a CUDA graph (no host-side state takes part).
Every result must equal, bit for bit, what a fresh handle computes for the same input."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from lisec_b200 import synth  # noqa: E402
from lisec_b200.weights import synthetic_vfe_pack  # noqa: E402

pytestmark = pytest.mark.gpu


def _frontend(max_points=450_000, max_sweeps=4):
    from lisec_b200 import Frontend

    f = Frontend(device=0, max_points=max_points, max_sweeps=max_sweeps)
    f.set_weights(synthetic_vfe_pack(1))
    return f


def _fresh(points, off):
    f = _frontend()
    try:
        grid = f.forward(points, off)
        torch.cuda.synchronize()
        return grid.clone(), f.counts()
    finally:
        f.close()


def _batch(n_sweeps, n_points, seed):
    pts, off = synth.sweep_batch(n_sweeps, n_points, seed0=seed)
    return pts, off


def test_fused_stage_repeats_on_one_voxelization():
    pts, off = _batch(2, 60_000, 11)
    want, _ = _fresh(pts, off)
    f = _frontend()
    try:
        f.voxelize(pts, off)
        out = f.new_grid(2)
        for _ in range(4):
            out.fill_(float("nan"))  # every cell must be written again, by exactly the same values
            f.vfe_scatter_fused(out=out)
            torch.cuda.synchronize()
            assert torch.equal(out, want)
    finally:
        f.close()


def test_alternating_batch_sizes_and_dropped_points_leave_nothing_behind():
    far = np.full((5_000, 3), 1.0e3, np.float32)  # every point out of range: drop counters without a single voxel
    nonfinite = np.array([[np.nan, 0.0, 1.0], [0.0, np.inf, 1.0], [1.0, 2.0, 0.5]], np.float32)
    cases = [_batch(4, 100_000, 3), _batch(1, 20_001, 5), (far, [0, len(far)]), _batch(3, 70_000, 7),
             (nonfinite, [0, 3]), _batch(1, 100_000, 9), _batch(4, 100_000, 3)]
    wants = [_fresh(p, o) for p, o in cases]
    f = _frontend()
    try:
        for rep in range(2):
            for (p, o), (want, want_counts) in zip(cases, wants):
                got = f.forward(p, o)
                torch.cuda.synchronize()
                assert torch.equal(got, want)
                got_counts = f.counts()
                assert [np.asarray(a).tolist() for a in got_counts] == [np.asarray(a).tolist() for a in want_counts]
    finally:
        f.close()


def test_front_end_replays_from_a_cuda_graph():
    pts, off = _batch(2, 80_000, 21)
    want, _ = _fresh(pts, off)
    f = _frontend()
    try:
        dev = torch.from_numpy(pts).cuda()
        out = f.new_grid(2)
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            f.forward(dev, off, out=out)  # warm-up outside the capture (first-call memsets, lazy module loading)
            side.synchronize()
            graph = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(graph, stream=side):
                    f.forward(dev, off, out=out)
            except Exception as e:  # a driver that cannot capture programmatic launches: not this library's claim
                pytest.skip("stream capture of the front end is not available here: %s" % str(e).splitlines()[0])
        for _ in range(3):
            out.fill_(float("nan"))
            graph.replay()
            torch.cuda.synchronize()
            assert torch.equal(out, want)
        got = f.forward(dev, off)  # and an ordinary call after the replays
        torch.cuda.synchronize()
        assert torch.equal(got, want)
    finally:
        f.close()


def test_forward_graph_option_equals_eager_calls():
    """Frontend.forward(graph=True): first use of a buffer eager, second use captured, then replays — with the buffer's
    CONTENTS changing between calls."""
    a, off = _batch(2, 80_000, 31)
    b, _ = _batch(2, 80_000, 33)
    want_a, _ = _fresh(a, off)
    want_b, _ = _fresh(b, off)
    f = _frontend()
    try:
        buf = torch.from_numpy(a).cuda()
        out = f.new_grid(2)
        for i in range(6):
            src, want = (a, want_a) if i % 2 == 0 else (b, want_b)
            buf.copy_(torch.from_numpy(src))
            out.fill_(float("nan"))
            got = f.forward(buf, off, out=out, graph=True)
            torch.cuda.synchronize()
            assert got is out and torch.equal(out, want)
        assert len(f._graphs) == 1 and next(iter(f._graphs.values()))[1] is not None
        fresh = _frontend()
        try:
            fresh.voxelize(b, off)
            want_counts = fresh.counts()
        finally:
            fresh.close()
        got_counts = f.counts()  # the totals on the device are the last replay's (contents = b)
        assert [np.asarray(x).tolist() for x in got_counts] == [np.asarray(x).tolist() for x in want_counts]
    finally:
        f.close()

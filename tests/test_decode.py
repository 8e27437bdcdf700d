"""RPN decode + NMS: rpnToRegion / nonMaxSuppressionFast (rpnToRegion.py:18-164).

CPU: the oracle against the golden file minted from the reference's own lines; the vectorised oracle against the loop
oracle. GPU: decode_kernel and nms_kernel through the C ABI against the oracle."""
import os

import numpy as np
import pytest

from lisec_b200 import synth
from oracle import decode_oracle as DO


@pytest.fixture(scope="module")
def golden(golden_dir):
    return np.load(os.path.join(golden_dir, "decode.npz"))


def small_case(seed, out_x=24, out_y=30):
    cls, reg = synth.synthetic_rpn_output(seed, out_x, out_y, n_objects=6)
    boxes, prob = DO.decode_boxes(cls, reg)
    boxes[:, 0] += 20.0  # keep most of the small map inside the 0..100 range window
    boxes[:, 1] += 30.0
    return boxes, prob


def test_oracle_decode_matches_the_reference_lines(golden):
    cls, reg = synth.synthetic_rpn_output(0)
    boxes, prob = DO.decode_boxes(cls, reg)
    assert boxes.shape == (40000, 7) and boxes.dtype == np.float64 and prob.dtype == np.float32
    assert boxes[golden["s0_boxinfo_rows"]].tobytes() == golden["s0_boxinfo_sample"].tobytes()


def test_vectorised_oracle_nms_equals_the_loop_oracle_and_the_golden_file(golden):
    for seed in (3, 4):
        boxes, prob = small_case(seed)
        for thresh, mb in ((0., 20), (0.3, 300), (0.9, 7)):
            _, _, want = DO.non_max_suppression(boxes, prob, thresh, mb)
            _, _, got = DO.non_max_suppression_vec(boxes, prob, thresh, mb)
            assert got == want and len(got) > 1
    for seed in (0, 1):  # full size: the reference's own lines (np.delete taking the collected candidates)
        cls, reg = synth.synthetic_rpn_output(seed)
        boxes, prob = DO.decode_boxes(cls, reg)
        b, p, pick = DO.non_max_suppression_vec(boxes, prob, 0., 20)
        assert len(pick) == 21
        assert b.tobytes() == golden["s%d_boxes" % seed].tobytes() and p.tobytes() == golden["s%d_probs" % seed].tobytes()


def test_legacy_position_semantics_are_a_different_function(golden):
    """What line :68 literally did under numpy < 1.19 is stated in the oracle and differs from the intent after pick 1."""
    assert golden["s0_legacy_probs"][0] == golden["s0_probs"][0]
    assert not np.array_equal(golden["s0_legacy_probs"], golden["s0_probs"])
    boxes, prob = small_case(5)
    _, _, a = DO.non_max_suppression(boxes, prob, 0., 20, delete="legacy_positions")
    _, _, b = DO.non_max_suppression(boxes, prob, 0., 20, delete="by_value")
    assert a[0] == b[0]


def test_polygon_area_restatement():
    sq = [(0, 0), (2, 0), (2, 2), (0, 2)]
    assert DO.quad_intersection_area(sq, [(1, 1), (3, 1), (3, 3), (1, 3)]) == 1.0
    assert DO.quad_intersection_area(sq, [(1, 3), (3, 3), (3, 1), (1, 1)]) == 1.0  # clockwise clip polygon
    assert DO.quad_intersection_area(sq, [(5, 5), (6, 5), (6, 6), (5, 6)]) == 0.0
    assert DO.quad_intersection_area(sq, [(2, 0), (4, 0), (4, 2), (2, 2)]) == 0.0  # shared edge
    d = [(1, -1), (3, 1), (1, 3), (-1, 1)]  # diamond |x-1| + |y-1| <= 2: contains the square
    assert abs(DO.quad_intersection_area(sq, d) - 4.0) < 1e-12
    assert abs(DO.quad_intersection_area(sq, d) - DO.quad_intersection_area(d, sq)) < 1e-12
    # boxToShapely: yaw 0 spans width along x, length along y (serialize_data.py:151-163)
    p = DO.box_to_polygon([10.0, 20.0, 1.0, 4.0, 2.0, 1.5, 0.0]).points
    assert sorted(p) == [(9.0, 18.0), (9.0, 22.0), (11.0, 18.0), (11.0, 22.0)]
    assert DO.calculate_iou([0, 0, 1, 4, 2, 1.5, 0.0], [0, 0, 1, 4, 2, 1.5, 0.0]) > 1.0  # the reference's z extents are +-h


# ---- GPU -----------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_decode_matches_the_oracle():
    import torch

    from lisec_b200.decode import RegionDecoder

    dec = RegionDecoder()
    cls = np.stack([synth.synthetic_rpn_output(s)[0] for s in (0, 1, 2)])
    reg = np.stack([synth.synthetic_rpn_output(s)[1] for s in (0, 1, 2)])
    reg[0, 3, 5, 3] = 30.0  # exp overflow side: 1e13 m long
    reg[0, 3, 6, 4] = -90.0  # exp underflow to a denormal float32
    # through a fused 16-channel head buffer, as the network writes it (prob = channels 0-1, regress = 2-15)
    heads = torch.from_numpy(np.concatenate([cls, reg], axis=-1)).cuda()
    boxes, scores = dec.decode(heads[..., :2], heads[..., 2:])
    boxes, scores = boxes.cpu().numpy(), scores.cpu().numpy()
    for s in range(3):
        want_b, want_p = DO.decode_boxes(cls[s], reg[s])
        assert scores[s].tobytes() == want_p.tobytes()
        # x, y, z, yaw: float64 multiply-add of exactly representable inputs — bit for bit
        assert boxes[s][:, [0, 1, 2, 6]].tobytes() == np.ascontiguousarray(want_b[:, [0, 1, 2, 6]]).tobytes()
        # l, w, h carry a float32 exp: numpy's and CUDA's expf agree to 3 ulp of float32 (4e-7 relative; absolute for
        # results in float32's denormal range)
        err = np.abs(boxes[s][:, 3:6] - want_b[:, 3:6])
        assert (err <= 4e-7 * want_b[:, 3:6] + 1e-37).all(), (err / want_b[:, 3:6]).max()


@pytest.mark.gpu
@pytest.mark.parametrize("thresh,max_boxes", [(0., 20), (0.3, 300), (0.9, 40)])
def test_gpu_nms_matches_the_oracle_small_maps(thresh, max_boxes):
    import torch

    from lisec_b200.decode import RegionDecoder

    dec = RegionDecoder()
    cases = [small_case(s, ox, oy) for s, ox, oy in ((3, 24, 30), (4, 24, 30), (6, 17, 29), (7, 8, 4))]
    for boxes, prob in cases:
        b = torch.from_numpy(boxes).cuda()[None]
        s = torch.from_numpy(prob).cuda()[None]
        picks, n_picks, out_b, out_s = dec.nms(b, s, thresh, max_boxes)
        k = int(n_picks[0])
        wb, wp, want = DO.non_max_suppression(boxes, prob, thresh, max_boxes)
        got = picks[0].cpu().numpy()
        assert k == len(want) and got[:k].tolist() == want and (got[k:] == -1).all()
        assert out_b[0, :k].cpu().numpy().tobytes() == wb.tobytes() and out_s[0, :k].cpu().numpy().tobytes() == wp.tobytes()


@pytest.mark.gpu
def test_gpu_rpn_to_region_full_size_against_golden_and_oracle(golden):
    import torch

    from lisec_b200.decode import RegionDecoder, nonMaxSuppressionFast, rpnToRegion

    for seed in (0, 1):
        cls, reg = synth.synthetic_rpn_output(seed)
        boxes, probs = rpnToRegion(cls, reg)
        gb, gp = golden["s%d_boxes" % seed], golden["s%d_probs" % seed]
        assert boxes.shape == gb.shape == (21, 7) and probs.dtype == np.float32
        assert probs.tobytes() == gp.tobytes()  # the same 21 candidates in the same order
        assert boxes[:, [0, 1, 2, 6]].tobytes() == np.ascontiguousarray(gb[:, [0, 1, 2, 6]]).tobytes()
        assert (np.abs(boxes[:, 3:6] - gb[:, 3:6]) / gb[:, 3:6]).max() <= 4e-7
    # batched, larger pick budget, on the boxes the GPU itself decoded: pick lists equal the oracle's on the same boxes
    dec = RegionDecoder()
    cls = np.stack([synth.synthetic_rpn_output(s)[0] for s in (2, 3, 4)])
    reg = np.stack([synth.synthetic_rpn_output(s)[1] for s in (2, 3, 4)])
    b, s = dec.decode(torch.from_numpy(cls).cuda(), torch.from_numpy(reg).cuda())
    picks, n_picks, out_b, out_s = dec.nms(b, s, 0.1, 60)
    for i in range(3):
        _, _, want = DO.non_max_suppression_vec(b[i].cpu().numpy(), s[i].cpu().numpy(), 0.1, 60)
        assert picks[i, :int(n_picks[i])].cpu().numpy().tolist() == want
    # numpy drop-in of nonMaxSuppressionFast, and the empty input of :25-26
    bb, pp = nonMaxSuppressionFast(b[0].cpu().numpy(), s[0].cpu().numpy(), 0.1, 60)
    assert bb.tobytes() == out_b[0, :int(n_picks[0])].cpu().numpy().tobytes() and len(pp) == int(n_picks[0])
    assert nonMaxSuppressionFast(np.zeros((0, 7)), np.zeros((0,), np.float32)) == ([], [])


@pytest.mark.gpu
def test_gpu_nms_edge_cases():
    import torch

    from lisec_b200.decode import RegionDecoder

    dec = RegionDecoder()
    # identical boxes (union == intersect would be 0/0 only for zero volume), everything out of range, one candidate
    box = np.array([[50.0, 50.0, 1.0, 1.6, 3.9, 1.56, 0.0]])
    same = np.repeat(box, 5, axis=0)
    p = np.array([0.1, 0.5, 0.3, 0.5, 0.2], dtype=np.float32)  # tie between 1 and 3: the larger index first
    picks, n, _, _ = dec.nms(torch.from_numpy(same).cuda()[None], torch.from_numpy(p).cuda()[None], 0., 20)
    assert int(n[0]) == 1 and int(picks[0, 0]) == 3
    far = same.copy()
    far[:, 0] = 500.0  # out of range: only the first pick survives the range test it never takes (:53-58)
    picks, n, _, _ = dec.nms(torch.from_numpy(far).cuda()[None], torch.from_numpy(p).cuda()[None], 0., 20)
    assert int(n[0]) == 1 and int(picks[0, 0]) == 3
    apart = same.copy()
    apart[:, 0] = [10, 30, 50, 70, 90]
    picks, n, ob, os_ = dec.nms(torch.from_numpy(apart).cuda()[None], torch.from_numpy(p).cuda()[None], 0., 2)
    _, _, want = DO.non_max_suppression(apart, p, 0., 2)
    assert picks[0, :int(n[0])].cpu().numpy().tolist() == want and len(want) == 3  # len(pick) > maxBoxes stops at 3
    # z-disjoint boxes do not suppress each other (iou < 0), touching footprints neither
    z = same.copy()
    z[:, 2] = [0, 10, 20, 30, 40]
    picks, n, _, _ = dec.nms(torch.from_numpy(z).cuda()[None], torch.from_numpy(p).cuda()[None], 0., 20)
    assert int(n[0]) == 5
    _, _, want = DO.non_max_suppression(z, p, 0., 20)
    assert picks[0, :5].cpu().numpy().tolist() == want
    with pytest.raises(Exception):
        dec.nms(torch.from_numpy(same).cuda()[None], torch.from_numpy(p).cuda()[None], -0.5, 20)

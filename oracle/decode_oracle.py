"""CPU oracle for RPN decode + non-maximum suppression.  *** TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT ***

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.

Restated, with the reference lines each function follows (bot15498/Lisec):
  applyRegrssion / applyRegrssionNP       rpnToRegion.py:77-112
  decode_boxes                            rpnToRegion.py:115-152   (anchors, regression, anchor-major flattening)
  box_to_polygon                          serialize_data.py:151-163 (boxToShapely)
  calculate_iou                           serialize_data.py:140-181 (calculateIntersection / calculateUnion / calculateIoU)
  non_max_suppression                     rpnToRegion.py:18-74     (nonMaxSuppressionFast, loop for loop)
  rpn_to_region                           rpnToRegion.py:115-164
  quad_intersection_area / Polygon        shapely (GEOS) — a dependency absent from /root/reference and from this image,
                                          version unpinned by the reference (`from shapely.geometry import Polygon`,
                                          serialize_data.py:13). Polygon(p).intersection(Polygon(q)).area is restated
                                          as Sutherland-Hodgman clipping of one convex quadrilateral by the other plus
                                          the shoelace formula: the same set and the same area in exact arithmetic;
                                          GEOS' own floating-point sequence is not reproduced (PARITY UNPINNED for
                                          that step; it matters only for boxes that touch to within rounding).

literal_functions(): the reference's OWN lines rpnToRegion.py:18-164 and serialize_data.py:140-181, read from
/root/reference at run time and executed unmodified with `Polygon` bound to the restatement above (build container
only; tests/golden/make_golden_decode.py mints tests/golden/decode.npz with it).

Tie order: the reference sorts with np.argsort's default quicksort, which leaves the order of equal scores
unspecified; stable=True below uses a stable sort (equal scores: the larger index is picked first), the product's rule.
"""
from __future__ import annotations

import math
import os

import numpy as np

REF_DIR = "/root/reference"
ANCHORS = [[1.6, 3.9, 1.56, 0], [1.6, 3.9, 1.56, math.pi / 2]]  # Constants.py:17
NX, NY = 200, 400  # Constants.py:12-13
VOXEL_X, VOXEL_Y = 0.5, 0.25  # Constants.py:7-8


# ---- shapely stand-in ------------------------------------------------------------------------------------------------
def _cross(ax, ay, bx, by, cx, cy):
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax)


def quad_intersection_area(subj, clip):
    """Area of the intersection of two convex polygons given as vertex lists (any orientation)."""
    sa = 0.0
    for i in range(len(clip)):
        j = (i + 1) % len(clip)
        sa = sa + (clip[i][0] * clip[j][1] - clip[j][0] * clip[i][1])
    if sa == 0.0:
        return 0.0
    orient = 1.0 if sa > 0.0 else -1.0
    poly = [(float(p[0]), float(p[1])) for p in subj]
    for e in range(len(clip)):
        ax, ay = clip[e]
        bx, by = clip[(e + 1) % len(clip)]
        new = []
        n = len(poly)
        for k in range(n):
            cx, cy = poly[k]
            dx, dy = poly[(k + 1) % n]
            sc = orient * _cross(ax, ay, bx, by, cx, cy)
            sd = orient * _cross(ax, ay, bx, by, dx, dy)
            in_c, in_d = sc >= 0.0, sd >= 0.0
            if in_c:
                new.append((cx, cy))
            if in_c != in_d:
                t = sc / (sc - sd)
                new.append((cx + t * (dx - cx), cy + t * (dy - cy)))
        poly = new
        if not poly:
            return 0.0
    a2 = 0.0
    n = len(poly)
    for k in range(n):
        k1 = (k + 1) % n
        a2 = a2 + (poly[k][0] * poly[k1][1] - poly[k1][0] * poly[k][1])
    return 0.5 * abs(a2)


class _Area:
    def __init__(self, area):
        self.area = area


class Polygon:
    """The slice of shapely.geometry.Polygon the reference touches: Polygon(points).intersection(other).area."""

    def __init__(self, points):
        self.points = [(float(p[0]), float(p[1])) for p in points]
        xs = [p[0] for p in self.points]
        ys = [p[1] for p in self.points]
        self.cx, self.cy = sum(xs) / len(xs), sum(ys) / len(ys)
        self.r = max(math.hypot(p[0] - self.cx, p[1] - self.cy) for p in self.points)

    def intersection(self, other):
        # disjoint bounding circles: the intersection is empty (this is an exact shortcut, not an approximation)
        if math.hypot(self.cx - other.cx, self.cy - other.cy) > (self.r + other.r) * (1 + 1e-9) + 1e-9:
            return _Area(0.0)
        return _Area(quad_intersection_area(self.points, other.points))


# ---- serialize_data.py:140-181 ----------------------------------------------------------------------------------------
def box_to_polygon(box):
    theta = box[6]
    length = box[3]
    width = box[4]
    rr = (box[0] + math.cos(theta) * (width / 2), box[1] - math.sin(theta) * (width / 2))
    rl = (box[0] - math.cos(theta) * (width / 2), box[1] + math.sin(theta) * (width / 2))
    top_right = [rr[0] + math.sin(theta) * (length / 2), rr[1] + math.cos(theta) * (length / 2)]
    bot_right = [rr[0] - math.sin(theta) * (length / 2), rr[1] - math.cos(theta) * (length / 2)]
    top_left = [rl[0] + math.sin(theta) * (length / 2), rl[1] + math.cos(theta) * (length / 2)]
    bot_left = [rl[0] - math.sin(theta) * (length / 2), rl[1] - math.cos(theta) * (length / 2)]
    return Polygon([top_right, bot_right, bot_left, top_left])


def calculate_iou(box1, box2):
    area = box_to_polygon(box1).intersection(box_to_polygon(box2)).area
    bot_z = max(box1[2] - box1[5], box2[2] - box2[5])
    top_z = min(box1[2] + box1[5], box2[2] + box2[5])
    intersect = (top_z - bot_z) * area
    union = box1[3] * box1[4] * box1[5] + box2[3] * box2[4] * box2[5] - intersect
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.float64(intersect) / np.float64(union)


# ---- rpnToRegion.py:77-152 --------------------------------------------------------------------------------------------
def decode_boxes(labels_class, labels_regress, anchors=ANCHORS, voxel_x=VOXEL_X, voxel_y=VOXEL_Y):
    """(boxInfo [N,7] float64, probInfo [N]) of rpnToRegion.py:150-152; N = n_anchors * outX * outY, anchor-major."""
    out_x, out_y = labels_class.shape[0], labels_class.shape[1]
    vx, vy = voxel_x * 2, voxel_y * 2
    A = np.zeros((7,) + labels_class.shape[:2] + (len(anchors),))
    X, Y = np.meshgrid(np.arange(out_x), np.arange(out_y))
    for i, anchor in enumerate(anchors):
        reg = np.transpose(labels_regress[:, :, i * 7:i * 7 + 7], (2, 0, 1))
        A[0, :, :, i] = X.T * vx + vx / 2
        A[1, :, :, i] = Y.T * vy + vy / 2
        A[2, :, :, i] = 1.
        A[3, :, :, i] = anchor[0]
        A[4, :, :, i] = anchor[1]
        A[5, :, :, i] = anchor[2]
        A[6, :, :, i] = anchor[3]
        x, y, z, l, w, h, theta = (A[k, :, :, i] for k in range(7))
        tx, ty, tz, tl, tw, th, tyaw = (reg[k] for k in range(7))
        A[:, :, :, i] = np.stack((tx * l + x, ty * w + y, tz * h + z, np.exp(tl) * l, np.exp(tw) * w, np.exp(th) * h,
                                  tyaw + theta))
    prob = labels_class.transpose((2, 0, 1)).reshape((-1))
    boxes = np.reshape(A.transpose((0, 3, 1, 2)), (7, -1)).transpose((1, 0))
    return np.ascontiguousarray(boxes), np.ascontiguousarray(prob)


# ---- rpnToRegion.py:18-74 ---------------------------------------------------------------------------------------------
def non_max_suppression(box_info, prob_info, overlap_thresh=0.9, max_boxes=300, anchors=ANCHORS, limit=(100, 100),
                        stable=True, delete="by_value"):
    """nonMaxSuppressionFast, loop for loop. Returns (boxes, probs, pick).

    delete="by_value" (the product's contract): the candidates collected in toDelete are removed from idxs — what the
    function's own header comment describes (:19-23).
    delete="legacy_positions": what the reference's line :68 literally does. toDelete holds candidate ids (the VALUES
    subI, :59/:66) but np.delete(idxs, toDelete) removes POSITIONS; under the numpy of the reference's time (< 1.19)
    positions past the end were ignored with a DeprecationWarning, under current numpy the line raises IndexError in the
    first round (probed in the build container, numpy 2.3). Kept so the deviation is stated in executable form."""
    if len(prob_info) == 0:
        return [], [], []
    x_info, y_info = box_info[:, 0], box_info[:, 1]
    pick = []
    idxs = np.argsort(prob_info, kind="stable") if stable else np.argsort(prob_info)
    mx, my = anchors[0][0], anchors[0][1]
    while len(idxs) > 0:
        last = len(idxs) - 1
        cur = idxs[last]
        pick.append(int(cur))
        last_box = [box_info[cur, k] for k in range(7)]
        to_delete = []
        for sub in idxs[:last]:
            if x_info[sub] - mx < 0 or x_info[sub] + mx > limit[0] or y_info[sub] - my < 0 or y_info[sub] + my > limit[1]:
                to_delete.append(sub)
            else:
                box = [box_info[sub, k] for k in range(7)]
                if calculate_iou(last_box, box) > overlap_thresh:
                    to_delete.append(sub)
        idxs = np.delete(idxs, (last,))
        if delete == "by_value":
            idxs = idxs[~np.isin(idxs, np.asarray(to_delete, dtype=idxs.dtype))]
        else:
            idxs = np.delete(idxs, [d for d in to_delete if d < len(idxs)])
        if len(pick) > max_boxes:
            break
    return box_info[pick], prob_info[pick], pick


def non_max_suppression_vec(box_info, prob_info, overlap_thresh=0.9, max_boxes=300, anchors=ANCHORS, limit=(100, 100)):
    """non_max_suppression(delete="by_value", stable=True) with the per-candidate Python loop replaced by numpy for the
    range test and for the bounding-circle shortcut of Polygon.intersection; only candidates whose circles touch the
    pick's go through calculate_iou. Same picks (tests/test_decode.py checks it against the loop version)."""
    if len(prob_info) == 0:
        return [], [], []
    n = len(prob_info)
    order = np.argsort(prob_info, kind="stable")
    alive = np.ones(n, dtype=bool)
    x, y = box_info[:, 0], box_info[:, 1]
    mx, my = anchors[0][0], anchors[0][1]
    out_of_range = (x - mx < 0) | (x + mx > limit[0]) | (y - my < 0) | (y + my > limit[1])
    polys = {}

    def poly(i):
        if i not in polys:
            polys[i] = box_to_polygon(box_info[i])
        return polys[i]

    # centre and radius exactly as Polygon.__init__ computes them are only needed for the survivors of a looser test
    rad = 0.5 * np.sqrt(box_info[:, 3] ** 2 + box_info[:, 4] ** 2) * (1 + 1e-6) + 1e-6
    pick = []
    ptr = n - 1
    while True:
        while ptr >= 0 and not alive[order[ptr]]:
            ptr -= 1
        if ptr < 0:
            break
        cur = int(order[ptr])
        pick.append(cur)
        alive[cur] = False
        alive &= ~out_of_range
        d = np.hypot(x - x[cur], y - y[cur])
        close = np.nonzero(alive & (d <= (rad + rad[cur]) * (1 + 1e-6) + 1e-6))[0]
        last_box = [box_info[cur, k] for k in range(7)]
        for sub in close:
            if calculate_iou(last_box, [box_info[sub, k] for k in range(7)]) > overlap_thresh:
                alive[sub] = False
        if len(pick) > max_boxes:
            break
    return box_info[pick], prob_info[pick], pick


def rpn_to_region(labels_class, labels_regress, max_boxes=20, overlap_thresh=0., delete="by_value"):
    boxes, prob = decode_boxes(labels_class, labels_regress)
    bad = np.where((boxes[:, 3] < 0) | (boxes[:, 4] < 0) | (boxes[:, 5] < 0))
    if len(bad[0]) > 0:
        boxes = np.delete(boxes, bad, 0)
        prob = np.delete(prob, bad, 0)
    b, p, _ = non_max_suppression(boxes, prob, max_boxes=max_boxes, overlap_thresh=overlap_thresh, delete=delete)
    return b, p


# ---- literal mode -----------------------------------------------------------------------------------------------------
def literal_available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "rpnToRegion.py"))


def literal_functions(quiet: bool = True, delete: str = "by_value"):
    """(rpnToRegion, nonMaxSuppressionFast, calculateIoU) compiled from the reference's own lines, unmodified. The
    knob is on a DEPENDENCY: `np` is a shim whose delete() takes the list handed over at :68 either as the candidates it
    holds ("by_value": the evident intent, the product's contract) or as positions with out-of-range entries ignored
    ("legacy_positions": numpy < 1.19, the reference's era). With the real numpy 2.x the line raises IndexError."""
    import types

    real_delete = np.delete

    def shim_delete(arr, obj, axis=None):
        if isinstance(obj, list) and axis is None:  # only the toDelete call of :68 passes a list
            if delete == "by_value":
                return arr[~np.isin(arr, np.asarray(obj, dtype=arr.dtype))] if len(obj) else arr
            return real_delete(arr, [d for d in obj if d < len(arr)])
        return real_delete(arr, obj, axis)

    np_shim = types.ModuleType("np_shim")
    for k in dir(np):
        if not k.startswith("__"):
            try:
                setattr(np_shim, k, getattr(np, k))
            except Exception:
                pass
    np_shim.delete = shim_delete

    with open(os.path.join(REF_DIR, "Constants.py")) as f:
        consts = types.ModuleType("Constants")
        exec(compile(f.read(), "Constants.py", "exec"), consts.__dict__)
    with open(os.path.join(REF_DIR, "serialize_data.py")) as f:
        lines = f.readlines()
    sd = types.ModuleType("serialize_data_140_181")
    sd.__dict__.update({"math": math, "Polygon": Polygon, "np": np})
    exec(compile("".join(lines[139:181]), "serialize_data.py:140-181", "exec"), sd.__dict__)
    with open(os.path.join(REF_DIR, "rpnToRegion.py")) as f:
        lines = f.readlines()
    ns = {"np": np_shim, "math": math, "Constants": consts, "LoadDataModule": sd}
    if quiet:
        ns["print"] = lambda *a, **k: None
    exec(compile("".join(lines[17:164]), "rpnToRegion.py:18-164", "exec"), ns)
    return ns["rpnToRegion"], ns["nonMaxSuppressionFast"], sd.calculateIoU

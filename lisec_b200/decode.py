"""RPN decode + non-maximum suppression on the GPU — the step after the path.

    reference (rpnToRegion.py)                                          here
    ------------------------------------------------------------------- ---------------------------------------------
    rpnToRegion(labelsClass, labelsRegress)                  :115-164   rpnToRegion(...) -> (boxes [k,7] float64, probs [k])
    nonMaxSuppressionFast(boxInfo, probInfo, overlapThresh,  :18-74     nonMaxSuppressionFast(...)
                          maxBoxes)
                                                                        RegionDecoder.decode / .nms / .regions  (batched,
                                                                        device tensors in, device tensors out)

All arithmetic runs in decode_kernel / nms_kernel behind lisec_rpn_decode / lisec_nms_rotated
(lisec_b200/csrc/decode.cu). There is no CPU fallback.

Two things differ from what the reference's lines literally do, both stated in executable form in
oracle/decode_oracle.py: (1) line :68 hands np.delete the candidate ids it collected where positions are expected —
numpy >= 1.19 raises IndexError there, older numpy silently deleted other entries; here the collected candidates are
the ones removed, as the function's header comment (:19-23) says; (2) equal scores: np.argsort's quicksort leaves
their order unspecified; here the larger flat index is picked first.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as N
from . import constants as K

ANCHORS = [[1.6, 3.9, 1.56, 0.0], [1.6, 3.9, 1.56, math.pi / 2]]  # Constants.py:17


class RegionDecoder:
    def __init__(self, device: int = 0, out_x: int = K.nx // 2, out_y: int = K.ny // 2, anchors: Sequence = ANCHORS,
                 cell=(K.voxelx * 2, K.voxely * 2), limit=(100.0, 100.0)):
        if not torch.cuda.is_available():
            raise RuntimeError("lisec_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        if len(anchors) > N.LISEC_MAX_ANCHORS:
            raise ValueError("at most %d anchors" % N.LISEC_MAX_ANCHORS)
        self._lib = N.load()
        self.device = torch.device("cuda", device)
        self.out_x, self.out_y, self.n_anchors = out_x, out_y, len(anchors)
        self.n = out_x * out_y * len(anchors)
        d = N.lisec_rpn_desc(out_x=out_x, out_y=out_y, n_anchors=len(anchors), reserved=0, cell_x=cell[0], cell_y=cell[1],
                             anchor_z=1.0)
        for i, a in enumerate(anchors):
            for k in range(4):
                d.anchors[i][k] = float(a[k])
        self._desc = d
        self._margin = (float(anchors[0][0]), float(anchors[0][1]))  # Constants.anchors[0][0], [0][1] (:55-58)
        self._limit = (float(limit[0]), float(limit[1]))

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _fail(self, st):
        raise N.LisecError(st, self._lib.lisec_decode_last_error().decode("utf-8", "replace"))

    def decode(self, prob: torch.Tensor, regress: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """prob [B,out_x,out_y,A] and regress [B,out_x,out_y,7A] float32 on the device (any channel pitch: views of the
        network's fused head buffer are fine). Returns (boxInfo [B,N,7] float64, probInfo [B,N] float32)."""
        for t, ch in ((prob, self.n_anchors), (regress, 7 * self.n_anchors)):
            if t.dtype != torch.float32 or not t.is_cuda or t.dim() != 4 or t.shape[1:] != (self.out_x, self.out_y, ch):
                raise ValueError("expected cuda float32 [B,%d,%d,%d], got %s %s" % (self.out_x, self.out_y, ch,
                                                                                   t.dtype, tuple(t.shape)))
            if t.stride(3) != 1 or t.stride(1) != self.out_y * t.stride(2):
                raise ValueError("positions must be contiguous with a constant channel pitch")
        B = prob.shape[0]
        boxes = torch.empty((B, self.n, 7), dtype=torch.float64, device=self.device)
        scores = torch.empty((B, self.n), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            st = self._lib.lisec_rpn_decode(C.byref(self._desc), C.c_void_p(prob.data_ptr()), prob.stride(2),
                                            prob.stride(0), C.c_void_p(regress.data_ptr()), regress.stride(2),
                                            regress.stride(0), B, C.c_void_p(boxes.data_ptr()),
                                            C.c_void_p(scores.data_ptr()), self._stream())
        if st != N.LISEC_OK:
            self._fail(st)
        return boxes, scores

    def nms(self, boxes: torch.Tensor, scores: torch.Tensor, overlap_thresh: float = 0.9, max_boxes: int = 300):
        """boxes [B,n,7] float64, scores [B,n] float32 on the device. Returns (picks [B,max_boxes+1] int32, -1 padded;
        n_picks [B] int32; picked boxes [B,max_boxes+1,7] float64; picked scores [B,max_boxes+1] float32)."""
        if boxes.dtype != torch.float64 or scores.dtype != torch.float32 or boxes.dim() != 3 or boxes.shape[2] != 7 or \
                scores.shape != boxes.shape[:2]:
            raise ValueError("boxes [B,n,7] float64 and scores [B,n] float32 expected")
        boxes, scores = boxes.contiguous(), scores.contiguous()
        B, n = scores.shape
        d = N.lisec_nms_desc(overlap_thresh=float(overlap_thresh), max_boxes=int(max_boxes), reserved=0,
                             margin_x=self._margin[0], margin_y=self._margin[1], limit_x=self._limit[0],
                             limit_y=self._limit[1])
        dev = self.device
        picks = torch.empty((B, max_boxes + 1), dtype=torch.int32, device=dev)
        n_picks = torch.empty((B,), dtype=torch.int32, device=dev)
        out_b = torch.zeros((B, max_boxes + 1, 7), dtype=torch.float64, device=dev)
        out_s = torch.zeros((B, max_boxes + 1), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            st = self._lib.lisec_nms_rotated(C.byref(d), C.c_void_p(boxes.data_ptr()), C.c_void_p(scores.data_ptr()), n, B,
                                             C.c_void_p(picks.data_ptr()), C.c_void_p(n_picks.data_ptr()),
                                             C.c_void_p(out_b.data_ptr()), C.c_void_p(out_s.data_ptr()), self._stream())
        if st != N.LISEC_OK:
            self._fail(st)
        return picks, n_picks, out_b, out_s

    def regions(self, prob: torch.Tensor, regress: torch.Tensor, overlap_thresh: float = 0., max_boxes: int = 20):
        """rpnToRegion for a batch, everything on the device: decode + NMS with the reference's call-site arguments
        (maxBoxes=20, overlapThresh=0., rpnToRegion.py:163)."""
        boxes, scores = self.decode(prob, regress)
        return self.nms(boxes, scores, overlap_thresh, max_boxes)


_DECODERS: dict = {}


def _decoder(out_x: int, out_y: int, device: int = 0) -> RegionDecoder:
    key = (out_x, out_y, device)
    if key not in _DECODERS:
        _DECODERS[key] = RegionDecoder(device, out_x, out_y)
    return _DECODERS[key]


def nonMaxSuppressionFast(boxInfo, probInfo, overlapThresh=0.9, maxBoxes=300):  # noqa: N802,N803 - reference names
    """Drop-in for nonMaxSuppressionFast (rpnToRegion.py:18-74): numpy in, (boxes, probs) numpy out."""
    if len(probInfo) == 0:
        return [], []
    dec = _decoder(K.nx // 2, K.ny // 2)
    # the kernel ranks float32 scores (what the network's heads emit and Predict.predictMain saves). Scores that do not
    # survive the cast — float64 values between float32 neighbours, NaN — would be ranked differently from the
    # reference's np.argsort on the original array: refuse them instead of answering with another order.
    p64 = np.asarray(probInfo, dtype=np.float64)
    p32 = p64.astype(np.float32)
    if np.isnan(p64).any() or not np.array_equal(p32.astype(np.float64), p64):
        raise ValueError("nonMaxSuppressionFast: scores must be finite-or-inf float32-representable values (the "
                         "network's float32 head outputs); got values a float32 ranking would order differently")
    b = torch.from_numpy(np.ascontiguousarray(boxInfo, dtype=np.float64)).to(dec.device)[None]
    s = torch.from_numpy(np.ascontiguousarray(p32)).to(dec.device)[None]
    picks, n_picks, out_b, _ = dec.nms(b, s, overlapThresh, maxBoxes)
    k = int(n_picks[0])
    pick = picks[0, :k].cpu().numpy()
    return out_b[0, :k].cpu().numpy(), np.asarray(probInfo)[pick]


def rpnToRegion(labelsClass, labelsRegress):  # noqa: N802,N803 - reference names
    """Drop-in for rpnToRegion (rpnToRegion.py:115-164): labelsClass (outX,outY,2), labelsRegress (outX,outY,14) as
    saved by Predict.predictMain -> (boxes [k,7] float64, probs [k]), k <= 21."""
    cls = np.ascontiguousarray(labelsClass, dtype=np.float32)
    reg = np.ascontiguousarray(labelsRegress, dtype=np.float32)
    dec = _decoder(cls.shape[0], cls.shape[1])
    picks, n_picks, out_b, out_s = dec.regions(torch.from_numpy(cls).to(dec.device)[None],
                                               torch.from_numpy(reg).to(dec.device)[None])
    k = int(n_picks[0])
    return out_b[0, :k].cpu().numpy(), out_s[0, :k].cpu().numpy().astype(np.asarray(labelsClass).dtype, copy=False)

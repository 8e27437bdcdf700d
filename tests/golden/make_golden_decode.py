"""Mint tests/golden/decode.npz by running the reference's own lines rpnToRegion.py:18-164 (+ serialize_data.py:140-181)
unmodified, via oracle/decode_oracle.literal_functions(). Build container only:

    python tests/golden/make_golden_decode.py

Inputs are regenerated from lisec_b200.synth.synthetic_rpn_output(seed) by the tests; stored are the outputs:
  s{seed}_boxes / s{seed}_probs       rpnToRegion(labelsClass, labelsRegress) with np.delete at :68 taking the collected
                                      candidates (the evident intent; the product's contract)
  s{seed}_legacy_boxes / _probs       the same lines with np.delete behaving as numpy < 1.19 did (positions, out-of-range
                                      ignored): what the reference printed in its own era — documentation of the defect
  s0_boxinfo_sha / s0_probinfo_sha    digests of the decoded boxInfo / probInfo (:150-152), and a sample of rows
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from lisec_b200 import synth  # noqa: E402
from oracle import decode_oracle as DO  # noqa: E402


def main():
    assert DO.literal_available(), "needs /root/reference"
    out = {}
    for seed in (0, 1):
        cls, reg = synth.synthetic_rpn_output(seed)
        rpn, _, _ = DO.literal_functions(delete="by_value")
        b, p = rpn(cls, reg)
        out["s%d_boxes" % seed], out["s%d_probs" % seed] = np.asarray(b), np.asarray(p)
        if seed == 0:
            rpn, _, _ = DO.literal_functions(delete="legacy_positions")
            b, p = rpn(cls, reg)
            out["s0_legacy_boxes"], out["s0_legacy_probs"] = np.asarray(b), np.asarray(p)
            boxes, prob = DO.decode_boxes(cls, reg)
            out["s0_boxinfo_rows"] = np.arange(0, len(boxes), 397)
            out["s0_boxinfo_sample"] = boxes[::397]
            out["s0_boxinfo_sha_xyzyaw"] = np.frombuffer(hashlib.sha256(
                np.ascontiguousarray(boxes[:, [0, 1, 2, 6]]).tobytes()).digest(), dtype=np.uint8)
    np.savez_compressed(os.path.join(HERE, "decode.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

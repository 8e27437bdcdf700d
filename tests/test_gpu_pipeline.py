"""The reference's whole inference flow through the drop-ins, from files to boxes (Predict.py:9-59, rpnToRegion.py:276):
load_model(.h5) -> predictMain(samples, outPath, level5Data, model) with the GPU lidar ingest -> rpnToRegion on the saved
.npy files. Each stage is checked against its oracle on the stage's own input."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_files_to_boxes(tmp_path):
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from h5_writer import write_h5

    from lisec_b200 import compat, synth
    from lisec_b200.decode import rpnToRegion
    from lisec_b200.weights import synthetic_model_pack
    from oracle import decode_oracle as DO
    from oracle import ingest_oracle as IO
    from oracle import lisec_oracle as O
    from oracle import network_oracle as NO

    pack = synthetic_model_pack(4)
    h5 = str(tmp_path / "15SampleEpoch0.h5")
    write_h5(h5, {"model_weights/%s/%s:0" % (k.split("/")[0], k): v for k, v in pack.items()})
    model = compat.load_model(h5, custom_objects={"RepeatLayer": compat.RepeatLayer,
                                                  "MaxPoolingVFELayer": compat.MaxPoolingVFELayer})   # Predict.py:51-52
    data_dir = str(tmp_path / "lyft")
    samples, tables = [], synth.SyntheticLyftTables()
    for k in range(2):
        sample, t = synth.synthetic_lyft_sample(data_dir, n_points=60_000, seed=40 + k)
        for name in tables.tables:
            tables.tables[name].update(t.tables[name])
        samples.append(sample)
    out = tmp_path / "out"
    out.mkdir()
    compat.predictMain(samples, str(out), tables, model, dataDir=data_dir)                             # Predict.py:9-40
    for i, sample in enumerate(samples):
        prob = np.load(out / ("sample%d_label.npy" % i))
        reg = np.load(out / ("sample%d_regress.npy" % i))
        assert prob.shape == (1, 100, 200, 2) and reg.shape == (1, 100, 200, 14)
        # oracle chain on the same files: combine_lidar_data -> voxelize -> VFE -> scatter -> dense network
        pts = IO.combine_lidar_data(sample, data_dir, tables)
        vox = O.voxelize_np(pts, xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=35, maxVoxelX=100, maxVoxelY=200, maxVoxelZ=8)
        feat = O.vfe_forward(vox["features"].astype(np.float32), pack)
        grid = O.scatter_dense(vox["coords"], feat, O.c_empty(pack, 35), (8, 200, 400), dtype=np.float32)
        import torch

        want_p, want_r = NO.network_forward(grid[None], pack, dtype=torch.float32)
        for got, want in ((prob, want_p), (reg, want_r)):
            err = np.abs(got.astype(np.float64) - want)
            assert err.max() / np.abs(want).max() <= 2e-2  # north_star's bf16 bar
        # rpnToRegion on the saved files (rpnToRegion.py:270-276): the picks are the oracle's on the same tensors
        boxes, probs = rpnToRegion(prob[0], reg[0])
        wb, wp, _ = DO.non_max_suppression_vec(*DO.decode_boxes(prob[0], reg[0]), 0., 20)
        assert probs.tobytes() == np.asarray(wp).tobytes() and boxes.shape == wb.shape == (21, 7)
        assert boxes[:, [0, 1, 2, 6]].tobytes() == np.ascontiguousarray(wb[:, [0, 1, 2, 6]]).tobytes()

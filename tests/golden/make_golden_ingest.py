"""Mint tests/golden/ingest.npz by running the reference's own lines model_training.py:65-98 (rotate_points,
combine_lidar_data) unmodified, via oracle/ingest_oracle.literal_functions(). Build container only:

    python tests/golden/make_golden_ingest.py

Stored: the raw float32 records of the three synthetic sensor files, their quaternions and translations, and the
float64 (n,3) array the reference's combine_lidar_data returned for them; plus a second sample without LIDAR_FRONT_LEFT
(the "not all samples have all lidar data" branch, :76-79) and rotate_points(..., inverse=True) on the first file."""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from lisec_b200 import synth  # noqa: E402
from oracle import ingest_oracle as IO  # noqa: E402


def pack(sample, tables, data_dir):
    recs, quats, trans = [], [], []
    for s in IO.SENSOR_TYPES:
        if s not in sample["data"]:
            continue
        frame = tables.get("sample_data", sample["data"][s])
        cs = tables.get("calibrated_sensor", frame["calibrated_sensor_token"])
        recs.append(np.fromfile(os.path.join(data_dir, frame["filename"]), dtype=np.float32).reshape(-1, 5))
        quats.append(cs["rotation"])
        trans.append(cs["translation"])
    off = np.cumsum([0] + [len(r) for r in recs])
    return np.concatenate(recs), off, np.asarray(quats, dtype=np.float64), np.asarray(trans, dtype=np.float64)


def main():
    assert IO.literal_available(), "needs /root/reference"
    rotate_points, combine_lidar_data = IO.literal_functions()
    out = {}
    with tempfile.TemporaryDirectory() as d:
        sample, tables = synth.synthetic_lyft_sample(d, n_points=2500, seed=5)
        out["a_points"] = combine_lidar_data(sample, d, tables)
        out["a_records"], out["a_offsets"], out["a_quats"], out["a_trans"] = pack(sample, tables, d)
        out["a_inverse_first"] = rotate_points(out["a_records"][:out["a_offsets"][1], :3], out["a_quats"][0], inverse=True)
    with tempfile.TemporaryDirectory() as d:
        sample, tables = synth.synthetic_lyft_sample(d, n_points=700, seed=6, sensors=("LIDAR_TOP", "LIDAR_FRONT_RIGHT"))
        out["b_points"] = combine_lidar_data(sample, d, tables)
        out["b_records"], out["b_offsets"], out["b_quats"], out["b_trans"] = pack(sample, tables, d)
    assert out["a_points"].dtype == np.float64 and out["a_points"].shape == (2500, 3)
    np.savez_compressed(os.path.join(HERE, "ingest.npz"), **out)
    print("wrote ingest.npz:", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()

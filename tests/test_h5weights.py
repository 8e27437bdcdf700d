"""lisec_b200/h5weights.py (Keras .h5 ingestion without h5py) against files from the independent minimal writer in
tests/h5_writer.py. PARITY UNPINNED: no h5py-written file exists in this image (see the module docstring)."""
import numpy as np
import pytest

from h5_writer import write_h5
from lisec_b200 import h5weights as H
from lisec_b200.weights import synthetic_model_pack


def keras_paths(pack, optimizer=True):
    """The dataset paths model.save() produces: /model_weights/<layer>/<layer>/<weight>:0 (+ optimizer state)."""
    out = {}
    for k, v in pack.items():
        layer = k.split("/")[0]
        out["model_weights/%s/%s:0" % (layer, k)] = v
    if optimizer:
        out["optimizer_weights/training/SGD/iter:0"] = np.asarray(180, dtype=np.int64)
        out["optimizer_weights/training/SGD/dense/kernel/momentum:0"] = np.zeros((6, 16), np.float32)
    return out


@pytest.mark.parametrize("snod,tree,ver", [(8, 32, 0), (2, 2, 0), (3, 4, 1)])
def test_keras_model_file_round_trip(tmp_path, snod, tree, ver):
    pack = synthetic_model_pack(7)
    paths = keras_paths(pack)
    some = list(paths)
    path = str(tmp_path / "model.h5")
    write_h5(path, paths, snod_entries=snod, tree_children=tree, compact={some[1], some[40]}, split={some[2], some[77]},
             superblock_version=ver)
    got = H.read_keras_weights(path)
    assert set(got) == set(pack) and len(got) == 142
    for k, v in pack.items():
        assert got[k].dtype == v.dtype and got[k].shape == v.shape and got[k].tobytes() == v.tobytes(), k
    every = H.read_datasets(path)
    assert int(every["optimizer_weights/training/SGD/iter:0"].reshape(-1)[0]) == 180 and len(every) == 144


def test_save_weights_layout_float64_and_errors(tmp_path):
    # model.save_weights(): layer groups at the root; a float64 and an empty dataset
    data = {"dense/dense/kernel:0": np.arange(96, dtype=np.float64).reshape(6, 16),
            "batch_normalization/batch_normalization/gamma:0": np.ones(16, np.float32),
            "empty/empty/bias:0": np.zeros((0,), np.float32)}
    path = str(tmp_path / "w.h5")
    write_h5(path, data)
    got = H.read_keras_weights(path)
    assert got["dense/kernel"].dtype == np.float64 and np.array_equal(got["dense/kernel"], data["dense/dense/kernel:0"])
    assert got["empty/bias"].shape == (0,)
    bad = tmp_path / "bad.h5"
    bad.write_bytes(b"not an hdf5 file at all" * 10)
    with pytest.raises(H.H5FormatError):
        H.read_keras_weights(str(bad))
    raw = bytearray(open(path, "rb").read())
    raw[8] = 2  # superblock version 2: libver='latest' files are refused with a reason, not misread
    (tmp_path / "v2.h5").write_bytes(bytes(raw))
    with pytest.raises(H.H5FormatError, match="superblock version 2"):
        H.read_keras_weights(str(tmp_path / "v2.h5"))


def test_load_model_takes_h5(tmp_path, monkeypatch):
    """compat.load_model(path.h5) goes through the reader (the model object itself needs a GPU, so stop before it)."""
    from lisec_b200 import compat

    pack = synthetic_model_pack(2)
    path = str(tmp_path / "15SampleEpoch0.h5")
    write_h5(path, keras_paths(pack, optimizer=False))
    seen = {}
    monkeypatch.setattr(compat, "VoxelNetFrontEnd", lambda nx, ny, nz, T, w: seen.update(w) or "model")
    assert compat.load_model(path, custom_objects={"RepeatLayer": None, "MaxPoolingVFELayer": None}) == "model"
    assert set(seen) == set(pack) and all(np.array_equal(seen[k], pack[k]) for k in pack)

"""CPU oracle for the Lisec VoxelNet front end.  *** TEST INFRASTRUCTURE — NOT PART OF THE PRODUCT ***

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
The product (lisec_b200/) never does: it has no CPU path and fails loudly without its CUDA library.

What is restated, and from where (file:line into the reference checkout, bot15498/Lisec):
  get_voxel                       model_training.py:103-107
  vfe_preprocessing_loops         model_training.py:112-152   (loop-for-loop; sampler is pluggable, see below)
  voxelize_np                     the same arithmetic, vectorised; checked against the loops version
  to_dense                        Predict.py:29-30, model_training.py:279 (tf.sparse.to_dense, default 0)
  vfe_forward                     model_training.py:32-61 (RepeatLayer, MaxPoolingVFELayer), :155-186 (addVFELayer,
                                  addFCN, addDenseLayer), :229-235 (wiring), Keras defaults for BatchNormalization
  scatter_dense                   what MaxPoolingVFELayer(combine=True) leaves in [N,nz,nx,ny,C3] (:235)

Pinning status
  * Integer / float64 half (voxel keys, range test, grouping, centroid features, COO layout): PINNED. The reference's
    own source lines model_training.py:103-152 are executed verbatim by oracle/literal_reference.py in the build
    container, and tests/golden/ holds their outputs; tests/test_oracle.py checks this module against them.
  * Floating-point half (Dense/BN/ReLU/max/concat): PARITY UNPINNED. The arithmetic lives in TensorFlow/Keras
    (version unpinned by the reference: "Tensorflow (With Keras included)", README.md:5-10), which is not installed
    here and cannot be; the reference ships no tests, golden outputs or weights (.MISSING_LARGE_BLOBS). vfe_forward
    restates the published Keras layer semantics (Dense = x @ kernel, no bias; BatchNormalization inference
    y = (x - mean) * gamma / sqrt(var + 1e-3) + beta; ReLU; max over axis -2 INCLUDING pad rows; repeat; concat
    [pooled, pointwise]) and is evaluated in float64 as ground truth and in float32 as a noise-floor witness.

Sampling contract: the reference subsamples with the UNSEEDED global RNG (np.random.choice, :132), so its output is
not a function of its input. sampler="first_T" keeps the first T indices in point order (the product's rule);
sampler="numpy_rng" calls np.random.choice exactly as the reference does (seed np.random first to reproduce a run).
float32 points are up-cast to float64 before anything else, as combine_lidar_data's output would be (:93-94).
"""
from __future__ import annotations

from math import floor

import numpy as np

DENSE = ("dense", "dense_1", "dense_2")
BN = ("batch_normalization", "batch_normalization_1", "batch_normalization_2")
BN_EPS = 1e-3  # Keras BatchNormalization default; model_training.py:171 passes no arguments


# ---------------------------------------------------------------------------------------------------------------
# model_training.py:103-107
def get_voxel(point, xSize, ySize, zSize):
    x = floor(point[0] / xSize)
    y = floor(point[1] / ySize)
    z = floor(point[2] / zSize)
    return (x, y, z)


class SparseTensorLike:
    """The three fields of tf.SparseTensor that the reference's callers touch (model_training.py:151-152)."""

    def __init__(self, indices, values, dense_shape):
        self.indices = indices
        self.values = values
        self.dense_shape = list(dense_shape)
        self.shape = tuple(dense_shape)


def _sample(lst, s, sampler):
    if sampler == "first_T":
        return np.asarray(lst[:s], dtype=np.int64)
    if sampler == "numpy_rng":
        return np.random.choice(lst, size=s, replace=False)  # model_training.py:132, verbatim call
    return np.asarray(sampler(lst, s))


# model_training.py:112-152, loop for loop. Slow (about 15 s per 100 k-point sweep), like the original.
def vfe_preprocessing_loops(points, xSize, ySize, zSize, sampleSize, maxVoxelX, maxVoxelY, maxVoxelZ,
                            sampler="first_T", return_groups=False):
    points = np.asarray(points, dtype=np.float64)
    clusteredPoints = {}
    for idx, point in enumerate(points):  # :115
        if not np.isfinite(point).all():
            continue  # the reference raises in math.floor here; the product drops and counts such points
        key = get_voxel(point, xSize, ySize, zSize)
        if -maxVoxelX < key[0] and key[0] < maxVoxelX \
                and -maxVoxelY < key[1] and key[1] < maxVoxelY \
                and 0 < key[2] and key[2] < maxVoxelZ:  # :118-120, strict on both sides
            fixedKey = (key[0] + maxVoxelX, key[1] + maxVoxelY, key[2])  # :122
            if fixedKey in clusteredPoints:
                clusteredPoints[fixedKey].append(idx)
            else:
                clusteredPoints[fixedKey] = [idx]
    appendedPoints = {}
    sampled = {}
    for voxel in clusteredPoints:  # :129
        s = sampleSize if len(clusteredPoints[voxel]) > sampleSize else len(clusteredPoints[voxel])
        sampleIdx = _sample(clusteredPoints[voxel], s, sampler)
        currPoints = points[sampleIdx]
        centroid = np.mean(currPoints, axis=0)  # :135
        centroidX = currPoints[:, 0:1] - centroid[0]
        centroidY = currPoints[:, 1:2] - centroid[1]
        centroidZ = currPoints[:, 2:3] - centroid[2]
        concat = np.hstack((currPoints, centroidX, centroidY, centroidZ))
        buffer = np.vstack((concat, np.zeros((sampleSize - s, 6))))  # :141
        appendedPoints[voxel] = buffer
        sampled[voxel] = sampleIdx
    indices = []
    values = []
    for voxel in appendedPoints:  # :145-149
        for i in range(len(appendedPoints[voxel])):
            for j in range(len(appendedPoints[voxel][i])):
                indices.append((voxel[2],) + voxel[:2] + (i, j))  # (z, x, y, i, j)
                values.append(appendedPoints[voxel][i][j])
    st = SparseTensorLike(indices, values, [maxVoxelZ, maxVoxelX * 2, maxVoxelY * 2, sampleSize, 6])
    if return_groups:
        return st, clusteredPoints, sampled, appendedPoints
    return st


# ---------------------------------------------------------------------------------------------------------------
def voxelize_np(points, xSize, ySize, zSize, sampleSize, maxVoxelX, maxVoxelY, maxVoxelZ):
    """Vectorised restatement of model_training.py:113-142 under the first_T sampling contract.

    Returns a dict; voxel order is ascending linear cell id ((z*nx + x)*ny + y), i.e. the product's order:
      coords      int64 [V,3]   (z, x, y) after the +maxVoxel shift
      counts      int64 [V]     len(clusteredPoints[voxel])
      point_idx   int64 [V,T]   first min(count,T) point indices, ascending, -1 padded
      features    float64 [V,T,6]  the `buffer` rows of :141 ([x,y,z,x-cx,y-cy,z-cz], zero padded)
      first_idx   int64 [V]     smallest point index of the voxel (sorting on it gives the reference's dict order)
      n_nonfinite, n_out_of_range
    """
    pts = np.asarray(points, dtype=np.float64).reshape(-1, 3)
    T = int(sampleSize)
    nx, ny = 2 * maxVoxelX, 2 * maxVoxelY
    finite = np.isfinite(pts).all(axis=1)
    with np.errstate(invalid="ignore", over="ignore"):
        kx = np.floor(pts[:, 0] / xSize)
        ky = np.floor(pts[:, 1] / ySize)
        kz = np.floor(pts[:, 2] / zSize)
        keep = finite & (kx > -maxVoxelX) & (kx < maxVoxelX) & (ky > -maxVoxelY) & (ky < maxVoxelY) \
            & (kz > 0) & (kz < maxVoxelZ)
    idx = np.nonzero(keep)[0]
    fx = kx[idx].astype(np.int64) + maxVoxelX
    fy = ky[idx].astype(np.int64) + maxVoxelY
    fz = kz[idx].astype(np.int64)
    lin = (fz * nx + fx) * ny + fy
    order = np.argsort(lin, kind="stable")  # stable: ascending point order inside a voxel (:123-126)
    lin_s, pid = lin[order], idx[order]
    uniq, start, counts = np.unique(lin_s, return_index=True, return_counts=True)
    V = len(uniq)
    kept = np.minimum(counts, T)
    rank = np.arange(len(lin_s)) - np.repeat(start, counts)
    sel = rank < T
    vrow = np.repeat(np.arange(V), counts)[sel]
    point_idx = np.full((V, T), -1, dtype=np.int64)
    point_idx[vrow, rank[sel]] = pid[sel]
    # np.mean(currPoints, axis=0): float64 adds in row order starting from the additive identity, then one divide
    sums = np.zeros((V, 3), dtype=np.float64)
    for r in range(T):
        m = kept > r
        if not m.any():
            break
        sums[m] += pts[point_idx[m, r]]
    centroid = sums / np.maximum(kept, 1)[:, None]
    features = np.zeros((V, T, 6), dtype=np.float64)
    real = point_idx >= 0
    p = pts[np.where(real, point_idx, 0)]
    features[..., 0:3] = np.where(real[..., None], p, 0.0)
    features[..., 3:6] = np.where(real[..., None], p - centroid[:, None, :], 0.0)
    coords = np.stack([uniq // (nx * ny), (uniq // ny) % nx, uniq % ny], axis=1)
    return {
        "coords": coords,
        "counts": counts.astype(np.int64),
        "point_idx": point_idx,
        "features": features,
        "first_idx": pid[start] if V else np.zeros(0, np.int64),
        "linear": uniq,
        "n_nonfinite": int((~finite).sum()),
        "n_out_of_range": int((finite & ~keep).sum()),
    }


def coo_from_voxels(vox, sampleSize):
    """The indices/values lists of model_training.py:143-149 (dict = first-appearance order), from voxelize_np output."""
    order = np.argsort(vox["first_idx"], kind="stable")
    T = int(sampleSize)
    V = len(order)
    c = vox["coords"][order]
    ii, jj = np.meshgrid(np.arange(T), np.arange(6), indexing="ij")
    indices = np.empty((V, T, 6, 5), dtype=np.int64)
    indices[..., 0] = c[:, 0, None, None]
    indices[..., 1] = c[:, 1, None, None]
    indices[..., 2] = c[:, 2, None, None]
    indices[..., 3] = ii[None]
    indices[..., 4] = jj[None]
    values = vox["features"][order]
    return indices.reshape(-1, 5), values.reshape(-1)


# Predict.py:29-30 / model_training.py:279: sparse.to_dense(default_value=0., validate_indices=False)
def to_dense(indices, values, dense_shape, dtype=np.float64):
    dense = np.zeros(tuple(dense_shape), dtype=dtype)
    ind = np.asarray(indices, dtype=np.int64).reshape(-1, len(dense_shape))
    if len(ind):
        dense[tuple(ind.T)] = np.asarray(values, dtype=dtype)
    return dense


# ---------------------------------------------------------------------------------------------------------------
def _bn(x, pack, name, dtype):
    # Keras BatchNormalization, inference: tf.nn.batch_normalization(x, mean, var, beta, gamma, eps)
    g = pack[name + "/gamma"].astype(dtype)
    b = pack[name + "/beta"].astype(dtype)
    m = pack[name + "/moving_mean"].astype(dtype)
    v = pack[name + "/moving_variance"].astype(dtype)
    inv = g / np.sqrt(v + dtype(BN_EPS))
    return x * inv + (b - m * inv)


def _dense(x, pack, name, dtype):
    # addDenseLayer (:178-186): a bias-free Dense over the last axis (the Reshape pair around it changes nothing)
    k = pack[name + "/kernel"].astype(dtype)
    return (x.reshape(-1, x.shape[-1]) @ k).reshape(x.shape[:-1] + (k.shape[1],))


def _fcn_names(pack, post_dense):
    """Keras names of the three FCNs' layers, [(dense, bn, second dense or None)]: Dense layers are numbered in creation
    order, so the Dense -> BN -> Dense variant (the line commented out at :172, the graph model.png shows) owns two each."""
    if post_dense is None:  # tell the two graphs apart by dense_1: (2*c1, c2) in the current code, (c1, c1) in the other
        post_dense = pack["dense_1/kernel"].shape[0] == pack["dense/kernel"].shape[1]
    d = ["dense"] + ["dense_%d" % i for i in range(1, 6)]
    if post_dense:
        return [(d[0], BN[0], d[1]), (d[2], BN[1], d[3]), (d[4], BN[2], d[5])]
    return [(d[0], BN[0], None), (d[1], BN[1], None), (d[2], BN[2], None)]


def _fcn(x, pack, names, dtype):
    # addFCN (:169-174): addDenseLayer -> BatchNormalization -> ReLU; in the variant at :172, -> Dense(units, relu) instead
    dense, bn, post = names
    y = _bn(_dense(x, pack, dense, dtype), pack, bn, dtype)
    if post is not None:
        y = _dense(y, pack, post, dtype)
    return np.maximum(y, dtype(0))


def _vfe_layer(x, pack, names, dtype):
    # addVFELayer (:155-166): FCN -> MaxPoolingVFELayer (max over axis -2, keepdims) -> RepeatLayer -> Concatenate
    layer = _fcn(x, pack, names, dtype)
    pooling = layer.max(axis=-2, keepdims=True)  # includes the pad rows: nothing is masked anywhere
    pooling = np.repeat(pooling, layer.shape[-2], axis=-2)
    return np.concatenate([pooling, layer], axis=-1)  # [pooled, pointwise] (:164-165)


def vfe_forward(x, pack, dtype=np.float64, post_dense=None):
    """model_training.py:229-235 on a tensor [..., T, 6] (the dense [N,nz,nx,ny,T,6] input, or any batch of voxels).
    Returns [..., C3] — MaxPoolingVFELayer(combine=True) output. The widths come from the kernels' shapes; post_dense
    selects the FCN variant (None: read it off the shapes)."""
    names = _fcn_names(pack, post_dense)
    x = np.asarray(x).astype(dtype)  # the Keras model casts its input to float32
    out = _vfe_layer(x, pack, names[0], dtype)    # addVFELayer(in, 6, 32)
    out = _vfe_layer(out, pack, names[1], dtype)  # addVFELayer(., 32, 64)
    out = _fcn(out, pack, names[2], dtype)        # addFCN(., 64, 64)
    return out.max(axis=-2)                       # MaxPoolingVFELayer(combine=True)


def c_empty(pack, T, dtype=np.float64):
    """What the unmasked network leaves in a voxel whose T rows are all zero."""
    return vfe_forward(np.zeros((1, T, 6)), pack, dtype)[0]


def scatter_dense(coords_zxy, voxel_feat, background, grid_zxy, dtype=np.float32):
    nz, nx, ny = grid_zxy
    grid = np.empty((nz, nx, ny, voxel_feat.shape[-1]), dtype=dtype)
    grid[...] = background.astype(dtype)
    if len(coords_zxy):
        grid[coords_zxy[:, 0], coords_zxy[:, 1], coords_zxy[:, 2]] = voxel_feat.astype(dtype)
    return grid


def vfe_forward_dense_chunked(dense_fn, pack, grid_zxy, T, dtype=np.float32, chunk_z=1):
    """The reference's actual work load — the VFE stack on ALL nz*nx*ny*T slots (model.predict on the dense input) —
    evaluated one z-slab at a time so the float32 intermediates (5.7 GB per tensor at full size) stay bounded.
    dense_fn(z0, z1) returns the dense input slab [z1-z0, nx, ny, T, 6]."""
    nz, nx, ny = grid_zxy
    c3 = pack[_fcn_names(pack, None)[2][0] + "/kernel"].shape[1]
    out = np.empty((nz, nx, ny, c3), dtype=dtype)
    for z0 in range(0, nz, chunk_z):
        z1 = min(nz, z0 + chunk_z)
        out[z0:z1] = vfe_forward(dense_fn(z0, z1), pack, dtype)
    return out

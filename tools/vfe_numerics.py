"""CPU emulation of the arithmetic the VFE kernel uses, against the float64 oracle — run BEFORE spending GPU time on a
new formulation. Emulates, operation by operation in float32:

  VFE-1   z1 = base_v + small_row: base_v = sum_i origin_i * w_i (voxel origin, float64, kept as a float32 hi + lo pair),
          small_row = a 6-term float32 FMA chain over the in-voxel coordinates and the centroid offsets (all < 1 voxel)
  VFE-2   3xTF32 on the tensor core, pooled half and pointwise half in SEPARATE accumulators, added in float32
  FCN     3xTF32, one accumulator (K = 64)

The tensor core's accumulation is emulated pessimistically: every K = 8 MMA adds its exact partial dot product to the
float32 accumulator with truncation (round toward zero). Metric = tests/test_gpu_parity.py's.

    python tools/vfe_numerics.py [n_weight_packs]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lisec_b200 import synth  # noqa: E402
from lisec_b200.weights import synthetic_vfe_pack  # noqa: E402
from oracle import lisec_oracle as O  # noqa: E402

REF = dict(xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=35, maxVoxelX=100, maxVoxelY=200, maxVoxelZ=8)
f32 = np.float32


def within(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    floor = np.sqrt(np.mean(ref * ref))
    return float((np.abs(np.asarray(got, dtype=np.float64) - ref) / np.maximum(np.abs(ref), floor)).max())


def tf32_rna(x):
    u = np.asarray(x, dtype=f32).view(np.uint32)
    u = ((u.astype(np.uint64) + 0x1000) & 0xFFFFE000).astype(np.uint32)
    return u.view(f32)


def split(x):
    hi = tf32_rna(x)
    lo = tf32_rna((x.astype(f32) - hi).astype(f32))
    return hi, lo


def trunc32(x64):
    """float64 -> float32, round toward zero"""
    r = x64.astype(f32)
    over = np.abs(r.astype(np.float64)) > np.abs(x64)
    r[over] = np.nextafter(r[over], f32(0))
    return r


def mma3x(x, w, trunc=True):
    """x [n,K] float32, w [K,N] float32 -> float32 [n,N]: Wh*Xl + Wl*Xh + Wh*Xh in K = 8 steps, one accumulator"""
    xh, xl = split(x)
    wh, wl = split(w)
    acc = np.zeros((x.shape[0], w.shape[1]), dtype=f32)
    first = True
    for a, b in ((xl, wh), (xh, wl), (xh, wh)):
        for k0 in range(0, x.shape[1], 8):
            part = a[:, k0:k0 + 8].astype(np.float64) @ b[k0:k0 + 8].astype(np.float64)
            s = acc.astype(np.float64) + part
            acc = trunc32(s) if (trunc and not first) else s.astype(f32)
            first = False
    return acc


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(f32)


def bn_fold(pack, name):
    a = (f32(1.0) / np.sqrt(pack[name + "/moving_variance"] + f32(1e-3))).astype(f32) * pack[name + "/gamma"]
    b = (pack[name + "/beta"] - pack[name + "/moving_mean"] * a).astype(f32)
    return a.astype(f32), b


def emulate(vox, pack, mode="new"):
    feat = vox["features"].astype(f32)  # [V,T,6]
    V, T, _ = feat.shape
    kept = np.minimum(vox["counts"], T)
    real = np.arange(T)[None, :] < kept[:, None]
    w1 = pack["dense/kernel"]
    a1, b1 = bn_fold(pack, "batch_normalization")
    a2, b2 = bn_fold(pack, "batch_normalization_1")
    a3, b3 = bn_fold(pack, "batch_normalization_2")
    # ---- VFE-1 ----
    co = vox["coords"]  # z, x, y after the shift
    origin = np.stack([(co[:, 1] - REF["maxVoxelX"]) * REF["xSize"], (co[:, 2] - REF["maxVoxelY"]) * REF["ySize"],
                       co[:, 0] * REF["zSize"]], axis=1)  # float64, exact
    base64 = origin @ w1[:3].astype(np.float64)  # [V,16] (three DFMAs per output)
    bh = base64.astype(f32)
    bl = (base64 - bh.astype(np.float64)).astype(f32)
    # exact except for |x| << voxel size on the negative side (0.5 - 1e-30 rounds to 0.5): error <= 2^-25 voxel sizes
    local = (feat[..., :3].astype(np.float64) - origin[:, None, :]).astype(f32)
    small = np.zeros((V, T, 16), dtype=f32)
    terms = [local[..., 0], local[..., 1], local[..., 2], feat[..., 3], feat[..., 4], feat[..., 5]]
    for i, t in enumerate(terms):
        small = fma(t[..., None], np.broadcast_to(w1[i], small.shape), small)
    z1 = (bh[:, None, :] + (bl[:, None, :] + small).astype(f32)).astype(f32)
    h1 = np.maximum(fma(z1, np.broadcast_to(a1, z1.shape), np.broadcast_to(b1, z1.shape)), f32(0))
    h1 = np.where(real[..., None], h1, np.maximum(b1, f32(0)))  # pad rows: zero input
    p1 = h1.max(axis=1)  # [V,16] (all T rows: pad rows take part whenever kept < T)
    # ---- VFE-2 ----
    w2 = pack["dense_1/kernel"]
    q = mma3x(p1, w2[:16])  # pooled half, per voxel
    xw = mma3x(h1.reshape(-1, 16), w2[16:]).reshape(V, T, 32)
    z2 = (q[:, None, :] + xw).astype(f32)
    h2 = np.maximum(fma(z2, np.broadcast_to(a2, z2.shape), np.broadcast_to(b2, z2.shape)), f32(0))
    p2 = h2.max(axis=1)
    # ---- FCN ----
    w3 = pack["dense_2/kernel"]
    x3 = np.concatenate([np.broadcast_to(p2[:, None, :], h2.shape), h2], axis=-1).reshape(-1, 64)
    z3 = mma3x(x3, w3).reshape(V, T, 64)
    h3 = np.maximum(fma(z3, np.broadcast_to(a3, z3.shape), np.broadcast_to(b3, z3.shape)), f32(0))
    return h3.max(axis=1)


if __name__ == "__main__":
    clouds = {
        "lyft 100k": synth.lyft_like_sweep(100_000, seed=3),
        "lyft 20k": synth.lyft_like_sweep(20_000, seed=0),
        "saturated 150k": synth.saturated_cloud(150_000, n_sweeps=3, theta=2.0),
    }
    vox = {k: O.voxelize_np(p, **REF) for k, p in clouds.items()}
    worst = 0.0
    for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
        pack = synthetic_vfe_pack(seed)
        for name in clouds:
            ref = O.vfe_forward(vox[name]["features"].astype(np.float32), pack, np.float64)
            ref32 = O.vfe_forward(vox[name]["features"].astype(np.float32), pack, np.float32)
            e = within(emulate(vox[name], pack), ref)
            worst = max(worst, e)
            print("seed %d  %-15s %6d voxels  emulated err %.2e   (numpy float32 forward %.2e)" %
                  (seed, name, len(ref), e, within(ref32, ref)))
    print("worst %.2e (bar 1e-5)" % worst)

/*
 * lisec_b200.h — C ABI of liblisec_b200.so: the B200-native VoxelNet front end of bot15498/Lisec.
 *
 * One data-parallel hot path, and nothing else:
 *     points (n,3)  ->  point-to-voxel grouping  ->  stacked VFE  ->  dense [N,nz,nx,ny,C3] voxel grid
 *
 * Each entry point names the reference interface it replaces (file:line into the reference checkout).
 * The reference has no FFI of its own (it is pure Python + Keras); the binding a maintainer would add is a
 * ctypes stub, shown in INTEGRATION.md and shipped as lisec_b200/_native.py.
 *
 * Conventions
 *   - every function returns LISEC_OK (0) or a negative lisec_status; nothing throws, aborts or prints;
 *     lisec_last_error(h) returns the text of the last failure on that handle.
 *   - "device pointer" = CUDA global memory on the handle's device, caller-owned. "host pointer" = ordinary
 *     (preferably pinned) host memory. The library owns only the workspace it allocates in lisec_create().
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream). All device work is
 *     enqueued on it; functions marked [async] return before the work has finished.
 *   - one handle = one CUDA device; calls on one handle are not re-entrant; distinct handles are independent
 *     (one per GPU is what the multi-GPU sweep sharding uses). No global mutable state.
 *   - there is no CPU fallback: without a usable CUDA device lisec_create() fails with LISEC_ERR_CUDA.
 */
#ifndef LISEC_B200_H_
#define LISEC_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LISEC_ABI_VERSION 2
#define LISEC_MAX_SWEEPS 64 /* sweeps per call (sweep_offsets travels as a kernel parameter) */

typedef enum lisec_status {
  LISEC_OK = 0,
  LISEC_ERR_BAD_ARG = -1,   /* null pointer, negative size, unsorted sweep_offsets ... */
  LISEC_ERR_BAD_CONFIG = -2,/* unsupported grid / widths / dtype */
  LISEC_ERR_CAPACITY = -3,  /* more points or sweeps than the handle was created for */
  LISEC_ERR_CUDA = -4,      /* a CUDA runtime call failed; see lisec_last_error() */
  LISEC_ERR_STATE = -5,     /* call order violated (e.g. VFE before weights were set) */
  LISEC_ERR_UNSUPPORTED = -6
} lisec_status;

typedef enum lisec_dtype {
  LISEC_F32 = 0,
  LISEC_F64 = 1,
  LISEC_BF16 = 2
} lisec_dtype;

/*
 * Shapes and caps. The first seven fields are exactly the positional arguments of the reference's
 *   VFE_preprocessing(points, xSize, ySize, zSize, sampleSize, maxVoxelX, maxVoxelY, maxVoxelZ)
 * (model_training.py:112; called with Constants.py:7-20 values at Predict.py:21-28, model_training.py:270-277).
 * The grid is nz = max_voxel_z, nx = 2*max_voxel_x, ny = 2*max_voxel_y (model_training.py:151-152).
 */
typedef struct lisec_config {
  double voxel_x, voxel_y, voxel_z;  /* xSize, ySize, zSize            (Constants.py:7-9: 0.5, 0.25, 0.25) */
  int32_t sample_size;               /* sampleSize = T                 (Constants.py:20: 35)               */
  int32_t max_voxel_x;               /* maxVoxelX = nx/2               (100)                               */
  int32_t max_voxel_y;               /* maxVoxelY = ny/2               (200)                               */
  int32_t max_voxel_z;               /* maxVoxelZ = nz                 (8)                                 */
  int32_t c1, c2, c3;                /* VFE-1 / VFE-2 / FCN output widths: (16, 32, 64) = model_training.py:231-233 as it
                                        stands, or (16, 64, 128) = the graph model.png shows             */
  int32_t grid_dtype;                /* lisec_dtype of the dense grid: LISEC_F32 or LISEC_BF16             */
  int32_t max_sweeps;                /* capacity: sweeps per call, <= LISEC_MAX_SWEEPS                     */
  int64_t max_points;                /* capacity: points per call, summed over its sweeps                  */
  int32_t device;                    /* CUDA device ordinal                                                */
  int32_t fcn_post_dense;            /* 0: addFCN = Dense -> BN -> ReLU (model_training.py:169-174 as it stands);
                                        1: Dense -> BN -> Dense(units, relu, no bias) — the line commented out at :172,
                                        the graph model.png shows. (16, 32, 64) with 0 runs the tensor-core kernel
                                        (vfe.cu); every other supported combination the float32 kernel (vfe_generic.cu). */
} lisec_config;

/*
 * VFE parameters in Keras layout and creation order (model_training.py:229-235):
 *   dense    kernel (6,  c1)   + batch_normalization   {gamma, beta, moving_mean, moving_variance}[c1]
 *   dense_1  kernel (2*c1, c2) + batch_normalization_1 {...}[c2]       rows 0..c1-1 multiply the POOLED half
 *   dense_2  kernel (2*c2, c3) + batch_normalization_2 {...}[c3]       rows 0..c2-1 multiply the POOLED half
 * Kernels are row-major (C_in, C_out), bias-free (model_training.py:184). BatchNormalization uses
 * bn_epsilon (Keras default 1e-3, model_training.py:171). All pointers are HOST pointers to float32.
 * With lisec_config.fcn_post_dense = 1 every FCN has a second Dense behind its BatchNormalization (Keras then names the
 * six kernels dense, dense_1 | dense_2, dense_3 | dense_4, dense_5): post_dense_kernel[l] is (c_l, c_l), row-major,
 * bias-free, followed by the ReLU; it is ignored (may be NULL) with fcn_post_dense = 0.
 */
typedef struct lisec_vfe_weights {
  const float* dense_kernel[3];
  const float* bn_gamma[3];
  const float* bn_beta[3];
  const float* bn_mean[3];
  const float* bn_var[3];
  float bn_epsilon;
  int32_t reserved;
  const float* post_dense_kernel[3];
} lisec_vfe_weights;

typedef struct lisec_handle lisec_handle;

/* ---- lifetime ---------------------------------------------------------------------------------------- */

int32_t lisec_abi_version(void);

/* Allocates every workspace once (cell tables, point lists, voxel rows). Replaces nothing in the reference
 * (which re-creates Python dicts per call, model_training.py:113,128); exists so the hot path never allocates. */
/* On failure *out is still a valid (inert) handle so lisec_last_error(*out) explains why; free it with lisec_destroy(). */
int32_t lisec_create(const lisec_config* cfg, lisec_handle** out);
void lisec_destroy(lisec_handle* h);
const char* lisec_last_error(const lisec_handle* h);
/* Bytes of device workspace owned by the handle. */
int64_t lisec_workspace_bytes(const lisec_handle* h);

/* Replaces load_model(...)/createModel(...) for the first 23 Keras layers (Predict.py:51-52,
 * model_training.py:229-235, 337-338). Folds BN to (scale, shift), uploads, and recomputes c_empty — the C3-vector
 * the reference's unmasked network produces for a voxel holding only zero rows. Synchronous. */
int32_t lisec_set_vfe_weights(lisec_handle* h, const lisec_vfe_weights* w, void* stream);
/* Device-to-host copy of c_empty (c3 floats, always float32). Synchronous. */
int32_t lisec_get_c_empty(lisec_handle* h, float* c_empty_host);

/* ---- a1-a3: get_voxel + VFE_preprocessing phases 1-2 (model_training.py:103-142) ---------------------- */

/*
 * [async] Point -> voxel grouping for n_sweeps concatenated sweeps.
 *   points            device pointer, row-major (n,3) of points_dtype (LISEC_F32 or LISEC_F64), 16-byte aligned
 *   sweep_offsets     HOST pointer, n_sweeps+1 non-decreasing int64; sweep s owns points [off[s], off[s+1])
 * Key = (floor(x/voxel_x), floor(y/voxel_y), floor(z/voxel_z)) in float64, strict range test on both sides,
 * shift by (+max_voxel_x, +max_voxel_y, 0)  (model_training.py:103-107, 117-122). Per voxel the point list is
 * in ascending point order (:123-126); the kept rows are the first min(count, T) of it — the deterministic
 * replacement for np.random.choice at :132. Non-finite points are dropped and counted.
 * Results stay in the handle's workspace; read them back with lisec_voxels_export().
 */
int32_t lisec_voxelize(lisec_handle* h, const void* points, int32_t points_dtype,
                       const int64_t* sweep_offsets, int32_t n_sweeps, void* stream);

/* Voxel / row totals of the last lisec_voxelize() on this handle. Synchronises `stream`.
 *   n_voxels_per_sweep  host pointer, n_sweeps int32 (may be NULL)
 *   n_voxels, n_points_in_range (before the T cap), n_dropped_out_of_range, n_dropped_nonfinite
 *                       host pointers (may be NULL) */
int32_t lisec_voxel_counts(lisec_handle* h, int32_t* n_voxels_per_sweep, int64_t* n_voxels,
                           int64_t* n_points_in_range, int64_t* n_dropped_out_of_range,
                           int64_t* n_dropped_nonfinite, void* stream);

/*
 * [async] Export the grouping of the last lisec_voxelize(). Voxel order = ascending (sweep, (z*nx + x)*ny + y);
 * the reference's dict order (first appearance, model_training.py:123-126) is recovered by sorting voxels of a
 * sweep on point_idx[v][0]. Any output pointer may be NULL. Device pointers, sized for lisec_voxel_counts().n_voxels:
 *   coords     int32 [V,4]  (sweep, z, x, y)        — the (z,x,y) of model_training.py:148
 *   counts     int32 [V]    points that fell in the voxel BEFORE the T cap (len(clusteredPoints[voxel]), :131)
 *   point_idx  int32 [V,T]  kept point indices (index into the sweep's own points, as `idx` at :115), ascending, -1 padded
 *   features   float32 [V,T,6]  [x, y, z, x-cx, y-cy, z-cz]; centroid = float64 mean of the kept points in
 *              list order, offsets in float64, one rounding to float32 (model_training.py:134-141 + the Keras
 *              input cast); pad rows are zeros (:141)
 */
int32_t lisec_voxels_export(lisec_handle* h, int32_t* coords, int32_t* counts, int32_t* point_idx,
                            float* features, void* stream);

/* ---- a4-a5: the reference-shaped dense input (tests / tiny grids only) -------------------------------- */

/* [async] dense float32 [n_sweeps,nz,nx,ny,T,6] exactly as sparse.to_dense(VFE_preprocessing(...)) + tf.stack
 * build it (model_training.py:143-152, 279, 285; Predict.py:29-30). 537.6 MB per sweep at the real grid —
 * this is the blow-up the product path exists to avoid; it is here so the drop-in can be checked end to end. */
int32_t lisec_emit_dense_input(lisec_handle* h, float* dense, void* stream);

/* ---- a6-a11: stacked VFE (model_training.py:155-186, 229-235) ----------------------------------------- */

/* [async] voxel_feat: device float32 [V,c3] — the row MaxPoolingVFELayer(combine=True) (model_training.py:235)
 * leaves in each occupied voxel; pad rows take part in every max exactly as in the unmasked reference. */
int32_t lisec_vfe_forward(lisec_handle* h, float* voxel_feat, void* stream);

/* [async] grid: device [n_sweeps,nz,nx,ny,c3] of cfg.grid_dtype — the tensor the first Conv3D consumes
 * (model_training.py:235-236). Every element is written exactly once: voxel_feat[v] where occupied, c_empty elsewhere. */
int32_t lisec_scatter_dense(lisec_handle* h, const float* voxel_feat, void* grid, void* stream);

/* [async] lisec_vfe_forward + lisec_scatter_dense as ONE kernel on the grouping of the last lisec_voxelize(): the voxel
 * rows go straight from the tensor-core accumulators to their cells while writer warps stream c_empty into the empty
 * cells by TMA bulk stores; no voxel_feat round trip, every grid element written exactly once. */
int32_t lisec_vfe_scatter_fused(lisec_handle* h, void* grid, void* stream);

/* [async] lisec_voxelize + lisec_vfe_scatter_fused without host round trips: what
 * VFE_preprocessing -> sparse.to_dense -> model.predict's first 23 layers do (Predict.py:21-38). */
int32_t lisec_frontend_forward(lisec_handle* h, const void* points, int32_t points_dtype,
                               const int64_t* sweep_offsets, int32_t n_sweeps, void* grid, void* stream);

/* Same, from HOST points (pinned for a truly asynchronous copy). The copy runs on the handle's own copy stream into one
 * of two staging buffers, so the H2D copy of call i+1 overlaps the kernels of call i; the kernels run on `stream`.
 * The host buffer must stay valid until the copy has completed (e.g. until `stream` has been synchronised).
 * This is the call the Python drop-in makes when it is handed a numpy array. [async] */
int32_t lisec_frontend_forward_host(lisec_handle* h, const void* points_host, int32_t points_dtype,
                                    const int64_t* sweep_offsets, int32_t n_sweeps, void* grid, void* stream);

/* [async] lisec_voxel_counts() without the synchronisation: enqueues the device-to-host copy of the totals on `stream`
 * into caller-owned pinned memory. Layout: int64[8] = {n_voxels, n_points_in_range, n_vfe_rows, n_vfe_tiles,
 * n_dropped_nonfinite, n_dropped_out_of_range, 0, 0}, then int32[n_sweeps+1] = exclusive prefix of voxels per sweep. */
int32_t lisec_voxel_counts_async(lisec_handle* h, void* pinned_out, int64_t pinned_bytes, void* stream);

/* Device time (ms, CUDA events on the call's stream) of the fused VFE + grid kernel inside the last
 * lisec_frontend_forward* / lisec_vfe_scatter_fused call. Synchronises on that kernel's end. */
int32_t lisec_last_fused_kernel_ms(lisec_handle* h, float* ms);

/* Debug aid, inert unless the environment had LISEC_TRACE=1 at lisec_create(): cycle counters of the VFE kernel's pipeline
 * stages in the last VFE launch, int64 [256 CTAs][16 slots] (slot meaning: lisec_b200/csrc/vfe.cu). Synchronous. */
/* ---- the VFE stack in TRAINING mode (model_training.py:229-235 under fit(), :295-299) -------------------------------
 * Forward with batch statistics and the backward pass on the grouping of the last lisec_voxelize(), evaluated on rows
 * with multiplicities (kept points, one virtual pad row per non-full voxel, one row for all empty voxels) — the exact
 * restatement of the dense graph that oracle/train_oracle.py: forward_train_rows() states. All pointers are DEVICE
 * pointers to float32; kernels are row-major (C_in, C_out) as Keras stores them. */
typedef struct lisec_vfe_train_params {
  const float* dense_kernel[3];
  const float* bn_gamma[3];
  const float* bn_beta[3];
  float* moving_mean[3]; /* updated in place: m <- m * momentum + batch * (1 - momentum); NULL: not tracked */
  float* moving_var[3];
  float bn_epsilon;  /* Keras default 1e-3 */
  float bn_momentum; /* Keras default 0.99 */
} lisec_vfe_train_params;
typedef struct lisec_vfe_train_grads {
  float* dkernel[3];
  float* dgamma[3];
  float* dbeta[3];
} lisec_vfe_train_grads;
/* [async] grid [n_sweeps, nz, nx, ny, 64] in the handle's grid_dtype: the VFE output in training mode, the empty voxels'
 * (batch-statistics) output as the background. Keeps the activations for lisec_vfe_train_backward(). */
int32_t lisec_vfe_train_forward(lisec_handle* h, const lisec_vfe_train_params* p, void* grid, void* stream);
/* [async] dgrid: float32 [n_sweeps, nz, nx, ny, 64], the loss gradient w.r.t. the grid. Writes the parameter gradients. */
int32_t lisec_vfe_train_backward(lisec_handle* h, const lisec_vfe_train_params* p, const float* dgrid,
                                 const lisec_vfe_train_grads* g, void* stream);
/* Test aid: per-voxel output rows of `layer` ([n_rows][C_layer], row n_voxels = the empty voxels) and its batch mean /
 * inverse standard deviation. Synchronous. */
int32_t lisec_vfe_train_read(lisec_handle* h, int32_t layer, float* out_rows, int64_t n_rows, float* mean, float* inv_std);

/* Device pointers of the handle's occupancy map (of the last lisec_voxelize()) and c_empty, and its voxel capacity. */
int32_t lisec_workspace_pointers(lisec_handle* h, const int32_t** cell_voxel, const float** c_empty, int64_t* max_voxels);
int32_t lisec_debug_trace(lisec_handle* h, int64_t* out, int64_t n);
/* Debug / test aid: copy one grouping table to the host (synchronous). which: 0 row_start, 1 row_voxel, 2 tile_first,
   3 tile_row0, 4 chunk_ntiles, 5 chunk_first, 6 voxel_cell. */
int32_t lisec_debug_table(lisec_handle* h, int32_t which, int32_t* out, int64_t n);

/* Number of kernels the last call on this handle launched (bench.py's gpu_launches). */
int32_t lisec_last_launch_count(const lisec_handle* h);

/* ---- dense layers behind the voxel grid: middle Conv3D stack and RPN --------------------------
 *
 * One plan = one Keras layer group of the reference's createModel (model_training.py:236-256), run as an implicit GEMM
 * on the tensor cores (lisec_b200/csrc/conv.cu). A plan binds caller-owned DEVICE buffers:
 *   in       bf16 [batch, in_d, in_h, in_w, in_c]                 channels-last (Conv2D layers: in_d = 1)
 *   weights  bf16 [kd*kh*kw][n_tiles*out_c][in_c]                 tap-major, K (= in_c) contiguous
 *   scale, shift  float32 [out_c] (or [n_tiles*out_c] without shuffle)
 *   out      bf16 or float32 [batch, out_d, out_h, out_w, out_pitch], written at channel offset out_ch_off
 * and computes  out = act((in (*) weights) * scale + shift)  with zero padding (pad_d, pad_h, pad_w) — the reference's
 * ZeroPadding3D/2D + 'valid' convolution (:192-193, :202-203) — and strides (stride_d, stride_hw, stride_hw).
 * BatchNormalization, the convolution bias and, for the Conv3D blocks, the bias-free Dense that follows the BN (:194-195)
 * are affine and are folded into (weights, scale, shift) by the host (lisec_b200/network.py).
 * shuffle = s > 1: a Conv2DTranspose whose kernel equals its stride s (:248, :251): a 1x1 GEMM with s*s groups of
 * columns; group (i, j) lands at output position (s*h + i, s*w + j). out_h/out_w are then s x larger. A group is
 * n_tiles / (s*s) consecutive N-tiles of out_c columns each (usually 1).
 * tile_w x tile_h = 128 output positions per M-tile (tile_w a power of two).
 * in_dtype = LISEC_F32: every operand is a PAIR of float32 planes (hi = x rounded to tf32, lo = x - hi):
 *   in [2][batch, in_d, in_h, in_w, in_c], weights [2][taps][n_tiles*out_c][in_c], out [2][...] when out_split = 1;
 * each product is evaluated as Ah*Bl + Al*Bh + Ah*Bh on the tf32 tensor-core path with float32 accumulation, which
 * reproduces a float32 FMA chain (tools/umma_probe.cu); the partial sums leave the tensor core every 64 channels and are
 * added in float32 registers. m_tiles = 1, group_kh = 0, out_dtype = LISEC_F32, out_c <= 128 (wider layers: n_tiles). */
typedef struct lisec_conv_desc {
  int32_t batch, in_d, in_h, in_w, in_c;
  int32_t kd, kh, kw;
  int32_t stride_d, stride_hw;
  int32_t pad_d, pad_h, pad_w;
  int32_t out_c, n_tiles, shuffle;
  int32_t out_pitch, out_ch_off;
  int32_t relu;      /* 1: ReLU after the affine */
  int32_t out_dtype; /* LISEC_BF16 or LISEC_F32 */
  int32_t tile_w, tile_h;
  int32_t m_tiles;   /* 1 or 2 M-tiles stacked along H per CTA tile: they share every weight box (needs tile_w >= 8) */
  int32_t in_dtype;  /* LISEC_BF16, or LISEC_F32 = 3xTF32 (float32-grade products, see below) */
  int32_t out_split; /* float32 plans: 1 = write hi / lo planes for the next float32 plan, 0 = plain float32 */
  int32_t group_kh;  /* 1: the kh taps of one (kd, kw) come from ONE input box with a kh-1 row halo; the weights are then
                        ordered [kd][kw][kh][n_tiles*out_c][in_c] (stride_hw = 1 only).
                        2: "halo" plan for 3x3 taps, stride_hw = 1, out_c <= 128, tile 8 x 16: ONE box with a 1-position
                        halo per (kd, 64 channels) serves all nine (kh, kw) taps; m_tiles stack along W; weights in
                        the plain [kd][kh][kw] order */
  int32_t reserved;  /* bit 0: scale / shift are rewritten between runs (a trainable bias): the epilogue reads them through
                        L2 instead of the read-only path. Other bits: 0 */
} lisec_conv_desc;

typedef struct lisec_conv_plan lisec_conv_plan;

/* Validates the description, encodes the TMA tensor maps of `in` and `weights`. The current CUDA device is the plan's. */
int32_t lisec_conv_plan_create(const lisec_conv_desc* desc, const void* in, const void* weights, const float* scale,
                               const float* shift, void* out, lisec_conv_plan** plan);
/* [async] One kernel launch on `stream`. */
int32_t lisec_conv_plan_run(lisec_conv_plan* plan, void* stream);
/* SURVEY §8f rank 1 (model_training.py:235-236): the first Conv3D reads the front end's SPARSE output — occupancy map
 * (cell -> voxel row or -1, lisec_workspace_pointers), float32 voxel rows [V,64] (lisec_vfe_forward), c_empty — and builds
 * its input boxes in shared memory itself; the dense voxel grid is never written or read. For halo plans (group_kh = 2)
 * with in_c = 64; the plan's `in` pointer is then unused. Bit-identical to the same plan on the materialised bf16 grid. */
int32_t lisec_conv_plan_set_gather(lisec_conv_plan* plan, const int32_t* cell_voxel, const float* voxel_feat,
                                   const float* c_empty);
/* [async] x[n] float32 -> hi[n], lo[n]: the operand planes a float32 plan reads (n a multiple of 4). */
int32_t lisec_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream);
/* [async] The tail of createModel without the 768-channel concat tensor: Conv2DTranspose (no activation, model_training.py:
 * 247-251) -> Concatenate (:252) -> ClassificationLayer / RegressionLayer (1x1, :253-254) is one linear map per RPN block.
 * With each transposed kernel folded into its 256 rows of the head kernels on the host, three plans leave float32 tensors
 *   c1 [batch, out_h, out_w, n_ch]                         (the k3 s1 block; its shift carries every bias)
 *   c2 [batch, out_h/s2, out_w/s2, s2*s2*n_ch], c3 likewise with s3: channel group (i*s + j) -> pixel (s*h + i, s*w + j)
 * and this call adds them into out [batch, out_h, out_w, n_ch] (n_ch = 2 + 14). All device pointers. */
int32_t lisec_heads_combine(const float* c1, const float* c2, int32_t s2, const float* c3, int32_t s3, float* out,
                            int32_t batch, int32_t out_h, int32_t out_w, int32_t n_ch, void* stream);
/* out_dhw[3] = output depth, height, width (after the pixel shuffle). */
int32_t lisec_conv_plan_output_shape(const lisec_conv_plan* plan, int32_t* out_dhw);
void lisec_conv_plan_destroy(lisec_conv_plan* plan);
/* Text of the last lisec_conv_* failure on the calling thread. */
const char* lisec_conv_last_error(void);

/* ---- lidar ingest: the step before the path ---------------------------------------------------
 *
 * Replaces the arithmetic of combine_lidar_data + rotate_points (model_training.py:65-98):
 *     rawPoints = np.fromfile(path, float32).reshape(-1, 5)[:, :3]                  (:87-90)
 *     points    = np.dot(Quaternion(rotation).rotation_matrix, rawPoints.T).T       (:65-69, :93)
 *     points    = points + np.array(translation);  np.concatenate over the sensors  (:94, :96)
 * A segment = the records of one sensor file; segments are concatenated in the order the reference appends them
 * (LIDAR_TOP, LIDAR_FRONT_RIGHT, LIDAR_FRONT_LEFT per sweep, :74), any number of sweeps behind one another.
 *   rotation     row-major 3x3 float64 = Quaternion(sensor['rotation']).rotation_matrix (built on the host,
 *                lisec_b200/ingest.py: three quaternions per sweep are not device work)
 *   translation  sensor['translation'] as float64 */
typedef struct lisec_sensor_pose {
  double rotation[9];
  double translation[3];
} lisec_sensor_pose;

/* [async] records: device float32 [n, record_floats] (record_floats = 5 for the Lyft .bin files; the first three are
 * x, y, z); segment_offsets: HOST int64 [n_segments + 1], in points, starting at 0; poses: HOST [n_segments];
 * points: device float64 [n, 3] — exactly the array combine_lidar_data returns, ready for lisec_voxelize(...,
 * LISEC_F64, ...). Bit-identical to numpy: float32 -> float64 widening, k-ascending FMA chain, then the float64 add.
 * launches_out (may be NULL) receives the number of kernels launched (one per 24 segments). Stateless; errors through
 * lisec_ingest_last_error() on the calling thread. */
int32_t lisec_ingest_lidar(const float* records, int32_t record_floats, const int64_t* segment_offsets,
                           const lisec_sensor_pose* poses, int32_t n_segments, double* points, void* stream,
                           int32_t* launches_out);
const char* lisec_ingest_last_error(void);

/* ---- RPN decode + non-maximum suppression: the step after the path ----------------------------
 *
 * Replaces rpnToRegion(labelsClass, labelsRegress) (rpnToRegion.py:115-164) on the two tensors model.predict returns:
 * anchors at the cell centres (:137-144), applyRegrssion (:77-88), anchor-major flattening (:150-152), then
 * nonMaxSuppressionFast(boxInfo, probInfo, maxBoxes=20, overlapThresh=0.) (:18-74) with calculateIoU of
 * serialize_data.py:140-181 (rotated footprints, the reference's full-height z extents). */
#define LISEC_MAX_ANCHORS 4

typedef struct lisec_rpn_desc {
  int32_t out_x, out_y, n_anchors;           /* Constants.nx // 2, Constants.ny // 2, len(Constants.anchors): 100, 200, 2 */
  int32_t reserved;
  double cell_x, cell_y;                     /* voxelXSize, voxelYSize of the output map: 2 * voxelx, 2 * voxely (:122-123) */
  double anchor_z;                           /* A[2] = 1. (:139) */
  double anchors[LISEC_MAX_ANCHORS][4];      /* Constants.anchors rows: length, width, height, yaw (Constants.py:17) */
} lisec_rpn_desc;

/* [async] prob: device float32, element (s, a, b, i) at prob[s*prob_batch_stride + (a*out_y + b)*prob_pitch + i];
 * regress likewise with 7*n_anchors channels. boxes: device float64 [batch, N, 7] (x, y, z, l, w, h, yaw), scores:
 * device float32 [batch, N], N = n_anchors*out_x*out_y, candidate index = i*out_x*out_y + a*out_y + b — boxInfo and
 * probInfo of rpnToRegion.py:150-152. float64 arithmetic as numpy evaluates it; the exp is float32's. */
int32_t lisec_rpn_decode(const lisec_rpn_desc* desc, const float* prob, int64_t prob_pitch, int64_t prob_batch_stride,
                         const float* regress, int64_t reg_pitch, int64_t reg_batch_stride, int32_t batch,
                         double* boxes, float* scores, void* stream);

typedef struct lisec_nms_desc {
  double overlap_thresh;   /* overlapThresh: a survivor is deleted when iou > overlap_thresh (>= 0)            */
  int32_t max_boxes;       /* maxBoxes: the loop stops once len(pick) > maxBoxes, i.e. at most max_boxes + 1  */
  int32_t reserved;
  double margin_x, margin_y, limit_x, limit_y; /* the range test of :55-58: Constants.anchors[0][0], [0][1], 100, 100 */
} lisec_nms_desc;

/* [async] nonMaxSuppressionFast on `batch` independent samples (one thread-block cluster each). boxes: device float64
 * [batch, n, 7]; scores: device float32 [batch, n]. Outputs (device): picks int32 [batch, max_boxes + 1] (flat candidate
 * indices in pick order, -1 padded), n_picks int32 [batch], out_boxes float64 [batch, max_boxes + 1, 7], out_scores
 * float32 [batch, max_boxes + 1] — boxInfo[pick], probInfo[pick] of :72-74. Score ties go to the larger index; NaN
 * scores are picked last (numpy's argsort would pick them first). */
int32_t lisec_nms_rotated(const lisec_nms_desc* desc, const double* boxes, const float* scores, int32_t n, int32_t batch,
                          int32_t* picks, int32_t* n_picks, double* out_boxes, float* out_scores, void* stream);
const char* lisec_decode_last_error(void);

/* ---- training step: the pieces that exist (BASELINE configs[4]) ---------------------------------------
 *
 * model_training.train() (model_training.py:260-302) compiles the model with
 *     optimizers.SGD(lr=0.01, decay=1e-6, momentum=0.9, nesterov=True), loss=['mse', 'mse']        (:295-296)
 * Built: the loss head and the optimizer update over the flat parameter buffer, around the one collective of the step
 * (NCCL all-reduce of the 6 491 024-element float32 gradient, lisec_b200/train.py). NOT built: the backward pass. */

/* [async] One Keras SGD update (optimizer_v2 / resource_apply_keras_momentum) on n float32 parameters:
 *   g = grad * grad_scale;  step = lr_t * g;  accum = accum * momentum - step;
 *   var += nesterov ? accum * momentum - step : accum          with lr_t = lr / (1 + decay * iterations) from the host.
 * grad_scale = 1 / world_size after a summing all-reduce. Device pointers, 16-byte aligned; float32 operation by
 * operation (no FMA contraction): bit-identical to the numpy float32 restatement in oracle/train_oracle.py. */
int32_t lisec_sgd_nesterov(float* var, float* accum, const float* grad, int64_t n, float grad_scale, float lr_t,
                           float momentum, int32_t nesterov, void* stream);
/* [async] 'mse' on one output tensor: *sum_sq += sum (y - target)^2 (double, device; the loss is sum_sq / n) and, when dy
 * is not NULL, dy = 2 (y - target) / n — d loss / d y, where the backward pass starts. */
int32_t lisec_mse_loss_grad(const float* y, const float* target, int64_t n, float* dy, double* sum_sq, void* stream);
/* [async] The weight operand of a data-gradient plan: the data gradient of a stride-1 convolution is a convolution of dy
 * (lisec_conv_plan_* with pad' = k - 1 - pad, in_c' = out_c, out_c' = in_c) with the kernel flipped in d, h, w and its
 * channel roles swapped. w: device float32 [kd*kh*kw][out_c][in_c] (master weights); out_bf16: device bf16
 * [kd*kh*kw][in_c][out_c]. */
int32_t lisec_weights_flip_transpose(const float* w, int32_t kd, int32_t kh, int32_t kw, int32_t out_c, int32_t in_c,
                                     void* out_bf16, void* stream);
/* [async] out = dy where y > 0, else 0 (bf16, n a multiple of 8): ReLU backward where no BatchNormalization precedes it. */
int32_t lisec_relu_backward(const void* dy, const void* y, int64_t n, void* out, void* stream);
/* [async] the same from a float32 dy (out stays bf16: it is a tensor-core operand). */
int32_t lisec_relu_backward_f32(const float* dy, const void* y, int64_t n, void* out, void* stream);
/* [async] Zero-dilation of a bf16 gradient tensor [batch, d, h, w, channels] into out [batch, out_d, out_h, out_w,
 * channels] (cleared once by the caller): element (d, h, w) lands at (d*stride_d, h*stride_hw, w*stride_hw). The data
 * gradient of a strided convolution is the stride-1 data gradient of the dilated dy. */
int32_t lisec_dilate(const void* in, int32_t batch, int32_t d, int32_t h, int32_t w, int32_t channels, int32_t stride_d,
                     int32_t stride_hw, int32_t out_d, int32_t out_h, int32_t out_w, void* out, void* stream);
/* [async] float32 [positions][c_in] -> bf16 [positions][c_out], channels c_in.. zero: the heads' 16-column gradient widened
 * to the 64 channels a tensor-core operand box holds. */
int32_t lisec_pad_channels_bf16(const float* in, int64_t positions, int32_t c_in, int32_t c_out, void* out_bf16, void* stream);
/* [async] a += b (bf16 tensors of n elements, n a multiple of 8; the sum is formed in float32): accumulation of the
 * gradients of a tensor that has two consumers. */
int32_t lisec_add_bf16(void* a, const void* b, int64_t n, void* stream);
/* [async] float32 master weights -> the bf16 operand copy the plans read. */
int32_t lisec_cast_f32_to_bf16(const float* w, int64_t n, void* out_bf16, void* stream);
/* [async] All of a training step's operand refreshes in one launch. entries: DEVICE array of n (<= 256) entries, `first`
 * = the running sum of the entries' element counts (kd*kh*kw*out_c*in_c), total = its end. mode 0: dst[j] = bf16(src[j]);
 * mode 1: lisec_weights_flip_transpose of src into dst. */
typedef struct lisec_refresh_entry {
  const float* src;
  void* dst;
  int64_t first;
  int32_t kd, kh, kw, out_c, in_c, mode;
} lisec_refresh_entry;
int32_t lisec_refresh_operands(const lisec_refresh_entry* entries, int32_t n, int64_t total, void* stream);
const char* lisec_train_last_error(void);

/* Training-mode BatchNormalization on channels-last bf16 activations [positions][channels] (lisec_b200/csrc/bn.cu): the
 * batch statistics of model.fit (model_training.py:171/194/204 under :299), deterministic two-stage reductions.
 * channels in {8, 16, 32, 64, 128, 256}. workspace: lisec_bn_workspace_bytes(positions, channels) bytes. All pointers
 * are device pointers; per-channel vectors are float32.
 * forward  [async]: mean, invstd (of the biased batch variance + eps), scale = gamma * invstd, shift = beta - mean * scale
 *                   are written; y = x * scale + shift (ReLU when relu & 1), bf16; moving_mean / moving_var (may be NULL)
 *                   <- momentum * moving + (1 - momentum) * batch. relu & 2: the moving variance takes the Bessel-
 *                   corrected batch variance var * P / (P - 1), as Keras's FUSED BatchNormalization does (rank-4 inputs).
 * backward [async]: g = dy (masked by y > 0 when relu != 0); dgamma = sum g * xhat, dbeta = sum g,
 *                   dx = gamma * invstd * (g - mean(g) - xhat * mean(g * xhat)), bf16; mean_g / mean_gx are scratch outputs. */
int64_t lisec_bn_workspace_bytes(int64_t positions, int32_t channels);
int32_t lisec_bn_train_forward(const void* x, int64_t positions, int32_t channels, const float* gamma, const float* beta,
                               float eps, float momentum, float* moving_mean, float* moving_var, int32_t relu, void* y,
                               float* mean, float* invstd, float* scale, float* shift, void* workspace, void* stream);
int32_t lisec_bn_train_backward(const void* x, const void* dy, const void* y, int64_t positions, int32_t channels,
                                const float* gamma, const float* mean, const float* invstd, int32_t relu, void* dx,
                                float* dgamma, float* dbeta, float* mean_g, float* mean_gx, void* workspace,
                                void* stream);
/* The same backward pass from a FLOAT32 gradient tensor dy: the gradient arriving at a BatchNormalization has a large
 * per-channel common mode that the backward pass removes — a bf16 copy of it keeps 8 bits of the wrong part. */
int32_t lisec_bn_train_backward_f32(const void* x, const float* dy, const void* y, int64_t positions, int32_t channels,
                                    const float* gamma, const float* mean, const float* invstd, int32_t relu, void* dx,
                                    float* dgamma, float* dbeta, float* mean_g, float* mean_gx, void* workspace,
                                    void* stream);
/* [async] sums[c] = sum over positions of x[p][c] (bf16 in, float32 out): the bias gradient of a convolution from dy. */
int32_t lisec_channel_sums(const void* x, int64_t positions, int32_t channels, float* sums, void* workspace, void* stream);
const char* lisec_bn_last_error(void);

/* Weight gradient of one convolution layer on the tensor cores (lisec_b200/csrc/wgrad.cu) — the first backward kernel:
 *   dw[tap][co][ci] = sum over output positions p of dy[p][co] * x[p * stride + tap - pad][ci]
 * `desc` describes the FORWARD convolution (geometry fields, tile_w x tile_h = 128 positions; first version: bf16,
 * stride_hw 1 or 2, in_c and out_c multiples of 64, out_c <= 256, in_c <= 1024, n_tiles = shuffle = 1). Device pointers:
 *   x   bf16 [batch, in_d, in_h, in_w, in_c]      the layer's input
 *   dy  bf16 [batch, out_d, out_h, out_w, out_c]  the gradient with respect to the convolution's output
 *   dw  float32 [kd*kh*kw][out_c][in_c]           the layout the forward plans read their weights in
 *   workspace  float32, lisec_conv_wgrad_workspace_bytes(desc) bytes (per-CTA partial sums, added in a fixed order:
 *              the result is deterministic)
 * [async] lisec_conv_wgrad_plan_run launches two kernels on `stream`. */
typedef struct lisec_wgrad_plan lisec_wgrad_plan;
int64_t lisec_conv_wgrad_workspace_bytes(const lisec_conv_desc* desc);
int32_t lisec_conv_wgrad_plan_create(const lisec_conv_desc* desc, const void* x, const void* dy, float* workspace,
                                     float* dw, lisec_wgrad_plan** plan);
int32_t lisec_conv_wgrad_plan_run(lisec_wgrad_plan* plan, void* stream);
void lisec_conv_wgrad_plan_destroy(lisec_wgrad_plan* plan);
const char* lisec_wgrad_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* LISEC_B200_H_ */

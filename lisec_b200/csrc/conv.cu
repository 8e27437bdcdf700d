// Dense convolutions of the VoxelNet middle layers and RPN on the sm_100a tensor cores: one implicit-GEMM kernel
// (TMA -> shared memory -> tcgen05.mma -> TMEM -> epilogue) behind lisec_conv_plan_*.
//
// Replaces, one plan per layer, the Keras layers of reference model_training.py:
//   addConv3DLayer  :191-196  ZeroPadding3D + Conv3D(64, k3, valid) + BatchNormalization + Dense(64, relu, no bias)
//   addConv2DLayer  :201-208  ZeroPadding2D + Conv2D(k3, stride) + BatchNormalization + ReLU
//   Conv2DTranspose :245,248,251 (k3 s1 / k2 s2 / k4 s4, padding='same'), Concatenate :252, the two 1x1 heads :253-254
// The host side (lisec_b200/network.py) folds each layer's affine tail into (weights, scale, shift); see there.
//
// Formulation. Activations are channels-last bf16 [B, D, H, W, C]. An output tile is 128 output positions: a
// tile_w x tile_h box in (W, H) at one (b, d). For every filter tap and every 64-channel block of C the A operand
// [128 positions x 64 ch] is ONE TMA box load whose start coordinate is the tile origin shifted by the tap — the
// zero padding of ZeroPadding3D/2D is the TMA's out-of-bounds fill, nothing is materialised — and the B operand
// [N x 64 ch] is one box of the [taps][N][C] weight tensor. Both land K-major with the 128-byte swizzle the tensor
// core reads (umma.cuh). Stride-2 layers read a space-to-depth VIEW of the same memory: [H/2][2][W/2][2C], where a
// tap picks a parity plane and a 64-channel window of the 2C pair, so a strided tap is again a plain box.
// Transposed convolutions with kernel == stride do not overlap: they are 1x1 GEMMs with N = k*k*C_out, run as k*k
// N-tiles whose epilogue writes the pixel-shuffled position. The k3 s1 one is a 3x3 convolution with the kernel
// flipped (host side).
//
// One persistent CTA per SM, 10 warps: warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane; owns TMEM),
// warps 2-9 = epilogue (TMEM lane quadrant = warp id % 4, two warps per quadrant split the columns). Three pipelines: shared-memory stages (full/empty
// mbarriers), two TMEM accumulators (acc_full/acc_empty) so a tile's epilogue overlaps the next tile's MMAs, and
// the static tile schedule (tile = blockIdx.x + i * gridDim.x, N-tile fastest so neighbours share A in L2).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "umma.cuh"

namespace lisec {
namespace {

constexpr int kConvThreads = 352;  // warp 0 TMA, warps 1 and 10 MMA (one M-tile each), warps 2-9 epilogue
constexpr int kSecondIssuerWarp = 10;
constexpr int kMaxTaps = 27;
constexpr int kMaxStages = 8;
constexpr uint32_t kABytes = 128 * 128;  // 128 positions x 64 bf16 channels

struct ConvParams {
  // tile schedule
  int tiles_w, tiles_h, out_d, batch, n_tiles;
  long long total_tiles;
  int bw, bw_log2, bh;
  // K loop: n_taps shared-memory stages per 64-channel block; a stage carries `group` filter taps (the kh taps of one
  // (kd, kw), built from row offsets of ONE input box with a kh-1 row halo) for `mt` M-tiles stacked along H
  int n_taps, c_blocks, N, stages, group, mt;
  uint32_t a_bytes, stage_bytes;  // a_bytes: ONE A operand image; f32 mode stages are [A hi][A lo][B hi][B lo]
  // float32 mode (3xTF32): 32 channels per 128-byte row; the lo planes are the same tensor maps at batch + lo_batch /
  // tap + lo_tap; the epilogue splits its float32 result into hi / lo planes out_lo_off elements apart (0: no split)
  int f32, cpb, lo_batch, lo_tap;
  int nsub;  // N-tiles per pixel-shuffle group (1 without shuffle)
  // halo mode (group_kh = 2): one input box with a 1-position halo per (kd, 64-channel block) serves all kh x kw taps;
  // the weights stream through their own ring, three kw taps per slot
  int halo, a_slots, b_slots, box_w, kd_n;
  int step_w, step_h;  // CTA tile pitch in output positions (M-tiles stack along H, or along W in halo mode)
  uint32_t a_slot_bytes, a_box_bytes;
  long long out_lo_off;
  // TMA coordinates of a stage: (c_off[t] + 64 cb, b1 + t1[t], b2 + t2[t], b3 + t3[t], b)
  int s2d, stride_d;
  short c_off[kMaxTaps];
  signed char t1[kMaxTaps], t2[kMaxTaps], t3[kMaxTaps];
  // epilogue: y = acc * scale[n] + shift[n], optional ReLU, to out[(b, d, h*shuffle + i, w*shuffle + j)][ch_off + n]
  int out_h, out_w, shuffle, relu, out_f32, out_ch_off;
  long long out_pitch;
  const float* scale;
  const float* shift;
  int affine_rewritten;  // lisec_conv_desc.reserved & 1: scale / shift change between runs (a trainable bias): L2 loads
  void* out;
  // gather source (halo plans of the FIRST Conv3D, lisec_conv_plan_set_gather): the input boxes are built in shared
  // memory from the front end's sparse output — occupancy map, voxel rows, c_empty — instead of being read from a dense
  // grid: the 82 MB-per-sweep bf16 grid is never written or read (model_training.py:235-236)
  int gather, in_d, in_h, in_w;
  const int* g_cell_voxel;    // [batch * in_d * in_h * in_w] voxel row of the cell or -1
  const float* g_voxel_feat;  // [voxels][64] float32 VFE rows
  const float* g_c_empty;     // [64]
};

// ---- PTX helpers not in umma.cuh ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t mbar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t mbar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// kind::f16 with bf16 operands, float32 accumulation, both operands K-major
__device__ __host__ constexpr uint32_t make_idesc_bf16_k(int M, int N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same, with the descriptors given as (low word, shared high word): the operand address lives in the low 14 bits of the
// low word, so stepping through a stage is a 32-bit add per operand. The single issuing thread's instruction stream is
// what paces the tensor pipe when the MMAs are small (N = 64), so every instruction per MMA counts.
__device__ __forceinline__ void mma_bf16_ss_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// the MMAs of one shared-memory stage: GROUP taps x MT M-tiles x 4 k-steps of 16 channels
template <int GROUP>
__device__ __forceinline__ void issue_stage(uint32_t a0, uint32_t b0, uint32_t d, uint32_t N, uint32_t h_row16,
                                            uint32_t m_rows, uint32_t idesc, uint32_t first) {
  const uint64_t proto = umma::make_desc_k_sw128(0);
  const uint32_t hi = (uint32_t)(proto >> 32), lo0 = (uint32_t)proto;
  const uint32_t a_lo0 = lo0 + (a0 >> 4) + m_rows * h_row16, b_lo0 = lo0 + (b0 >> 4), b_tap16 = N * 8u;  // N * 128 bytes / 16
#pragma unroll
  for (int g = 0; g < GROUP; ++g) {
    const uint32_t a_lo = a_lo0 + (uint32_t)g * h_row16, b_lo = b_lo0 + (uint32_t)g * b_tap16;
#pragma unroll
    for (int j = 0; j < 4; ++j) mma_bf16_ss_lo(d, a_lo + 2 * j, b_lo + 2 * j, hi, idesc, (g == 0 && j == 0) ? first : 1u);
  }
}

__device__ __forceinline__ void mma_tf32_ss_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      ".reg .b64 da, db;\n\t"
      "mov.b64 da, {%1, %3};\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %4, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// float32 mode: one stage = 32 channels = 4 k-steps of 8; per k-step the 3xTF32 triple, small terms first so that their
// sum is not rounded against the large one (as in the VFE kernel's FCN): Ah*Bl, Al*Bh, Ah*Bh.
__device__ __forceinline__ void issue_stage_tf32(uint32_t a0, uint32_t a_bytes, uint32_t b0, uint32_t b_bytes, uint32_t d0,
                                                 uint32_t idesc, uint32_t first) {
  const uint64_t proto = umma::make_desc_k_sw128(0);
  const uint32_t hi = (uint32_t)(proto >> 32), lo0 = (uint32_t)proto;
  const uint32_t ah = lo0 + (a0 >> 4), al = ah + (a_bytes >> 4), bh = lo0 + (b0 >> 4), bl = bh + (b_bytes >> 4);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    mma_tf32_ss_lo(d0, ah + 2 * j, bl + 2 * j, hi, idesc, j == 0 ? first : 1u);
    mma_tf32_ss_lo(d0, al + 2 * j, bh + 2 * j, hi, idesc, 1u);
    mma_tf32_ss_lo(d0, ah + 2 * j, bh + 2 * j, hi, idesc, 1u);
  }
}

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

struct TileCoord {
  int nt, ow0, oh0, od, b;
};
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& P, long long tile) {
  TileCoord t;
  t.nt = (int)(tile % P.n_tiles);
  long long r = tile / P.n_tiles;
  t.ow0 = (int)(r % P.tiles_w) * P.step_w;
  r /= P.tiles_w;
  t.oh0 = (int)(r % P.tiles_h) * P.step_h;
  r /= P.tiles_h;
  t.od = (int)(r % P.out_d);
  t.b = (int)(r / P.out_d);
  return t;
}

__device__ __forceinline__ unsigned pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<unsigned*>(&h);
}

// one chunk of NV consecutive channels of one output position: affine + ReLU + store
template <int NV>
__device__ __forceinline__ void store_chunk(const ConvParams& P, const float (&x)[NV], int sidx0, size_t elem0,
                                            bool valid) {
  float y[NV];
  const float4* sc = reinterpret_cast<const float4*>(P.scale + sidx0);  // sidx0 is a multiple of 16: 16-byte loads
  const float4* sf = reinterpret_cast<const float4*>(P.shift + sidx0);
#pragma unroll
  for (int i = 0; i < NV / 4; ++i) {
    // (inference: constants, through L1. A plan whose bias an optimizer rewrites between runs reads them through L2 — the
    // kernel is launched under programmatic dependent launch, where an L1 line of an earlier run can be served again)
    const float4 a = P.affine_rewritten ? __ldcg(sc + i) : __ldg(sc + i), b = P.affine_rewritten ? __ldcg(sf + i) : __ldg(sf + i);
    y[4 * i] = fmaf(x[4 * i], a.x, b.x);
    y[4 * i + 1] = fmaf(x[4 * i + 1], a.y, b.y);
    y[4 * i + 2] = fmaf(x[4 * i + 2], a.z, b.z);
    y[4 * i + 3] = fmaf(x[4 * i + 3], a.w, b.w);
  }
  if (P.relu) {
#pragma unroll
    for (int i = 0; i < NV; ++i) y[i] = fmaxf(y[i], 0.f);
  }
  if (!valid) return;
  if (P.out_f32) {
    float4* dst = reinterpret_cast<float4*>(static_cast<float*>(P.out) + elem0);
    if (P.out_lo_off) {  // the next layer's 3xTF32 operands: hi = rn_tf32(y), lo = rn_tf32(y - hi)
      float4* dlo = reinterpret_cast<float4*>(static_cast<float*>(P.out) + elem0 + P.out_lo_off);
#pragma unroll
      for (int i = 0; i < NV / 4; ++i) {
        float h[4], l[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) umma::tf32_split(y[4 * i + k], h[k], l[k]);
        dst[i] = make_float4(h[0], h[1], h[2], h[3]);
        dlo[i] = make_float4(l[0], l[1], l[2], l[3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV / 4; ++i) dst[i] = make_float4(y[4 * i], y[4 * i + 1], y[4 * i + 2], y[4 * i + 3]);
    }
  } else {
    uint4* dst = reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(P.out) + elem0);
#pragma unroll
    for (int i = 0; i < NV / 8; ++i)
      dst[i] = make_uint4(pack2(y[8 * i], y[8 * i + 1]), pack2(y[8 * i + 2], y[8 * i + 3]),
                          pack2(y[8 * i + 4], y[8 * i + 5]), pack2(y[8 * i + 6], y[8 * i + 7]));
  }
}

// one output position (oh, ow) of tile t: ncols accumulator columns from taddr on -> affine (+ReLU) -> global
__device__ __forceinline__ void epilogue_rows(const ConvParams& P, const TileCoord& t, int oh, int ow, int si, int sj,
                                              int n0, int col0, int ncols, uint32_t taddr) {
  const int sh = P.shuffle;
  const bool valid = oh < P.out_h && ow < P.out_w;
  const size_t pix = (((size_t)t.b * P.out_d + t.od) * ((size_t)P.out_h * sh) + (size_t)oh * sh + si) *
                         ((size_t)P.out_w * sh) + (size_t)ow * sh + sj;
  const size_t elem = pix * (size_t)P.out_pitch + P.out_ch_off + n0 + col0;
  if (ncols == 16) {
    float x[16];
    tmem_ld_32x16(taddr, x);
    store_chunk<16>(P, x, n0 + col0, elem, valid);
  } else {
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      float x[32];
      umma::tmem_ld_32x32(taddr + c0, x);
      store_chunk<32>(P, x, n0 + col0 + c0, elem + c0, valid);
    }
  }
}

// ===== TMA producer (one thread): every stage of every tile of this CTA, in order =====
__device__ __forceinline__ void producer_loop(const ConvParams& P, const CUtensorMap& map_a, const CUtensorMap& map_b,
                                              uint32_t base, uint32_t bar0) {
  const uint32_t b_tap_bytes = (uint32_t)P.N * 128u, stage_bytes = P.stage_bytes;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
  int s = 0;
  uint32_t ph = 0;
  for (long long tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
    const TileCoord t = decode_tile(P, tile);
    const int b1 = t.ow0, b2 = P.s2d ? 0 : t.oh0, b3 = P.s2d ? t.oh0 : t.od * P.stride_d;
    for (int tap = 0; tap < P.n_taps; ++tap) {
      const int c1 = b1 + P.t1[tap], c2 = b2 + P.t2[tap], c3 = b3 + P.t3[tap];
      for (int cb = 0; cb < P.c_blocks; ++cb) {
        umma::mbar_wait(bar_empty(s), ph ^ 1u);
        mbar_arrive_expect_tx(bar_full(s), stage_bytes);
        const uint32_t dst = base + (uint32_t)s * stage_bytes;
        const int c0 = P.c_off[tap] + P.cpb * cb;
        if (!P.f32) {
          tma_load_5d(dst, &map_a, bar_full(s), c0, c1, c2, c3, t.b);
          tma_load_3d(dst + P.a_bytes, &map_b, bar_full(s), P.cpb * cb, t.nt * P.N, tap * P.group);
        } else {
          tma_load_5d(dst, &map_a, bar_full(s), c0, c1, c2, c3, t.b);
          tma_load_5d(dst + P.a_bytes, &map_a, bar_full(s), c0, c1, c2, c3, t.b + P.lo_batch);
          tma_load_3d(dst + 2 * P.a_bytes, &map_b, bar_full(s), P.cpb * cb, t.nt * P.N, tap);
          tma_load_3d(dst + 2 * P.a_bytes + b_tap_bytes, &map_b, bar_full(s), P.cpb * cb, t.nt * P.N, tap + P.lo_tap);
        }
        if (++s == P.stages) { s = 0; ph ^= 1u; }
      }
    }
  }
}

__global__ void __launch_bounds__(kConvThreads, 1)
    conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      const __grid_constant__ ConvParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = umma::smem_u32(smem);
  if (base & 1023u) __trap();
  const uint32_t b_tap_bytes = (uint32_t)P.N * 128u;
  const uint32_t stage_bytes = P.stage_bytes;
  const uint32_t bar0 = base + (uint32_t)P.stages * stage_bytes;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
  auto bar_acc_full = [&](int a) { return bar0 + 8u * (2 * kMaxStages + a); };
  auto bar_acc_empty = [&](int a) { return bar0 + 8u * (2 * kMaxStages + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (size_t)P.stages * stage_bytes + 8 * (2 * kMaxStages + 4));

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    for (int s = 0; s < P.stages; ++s) {
      umma::mbar_init(bar_full(s), 1);
      umma::mbar_init(bar_empty(s), P.mt);  // one tcgen05.commit per issuing thread
    }
    for (int a = 0; a < 2; ++a) {
      umma::mbar_init(bar_acc_full(a), P.mt);
      umma::mbar_init(bar_acc_empty(a), 8);
    }
    umma::mbar_init_fence();
  }
  if (warp == 1) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();  // the previous layer's output is complete and visible from here on

  const int k_blocks = P.n_taps * P.c_blocks;
  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) producer_loop(P, map_a, map_b, base, bar0);
  } else if (warp == 1 || warp == kSecondIssuerWarp) {
    // ===== MMA issuers: one thread per M-tile =====
    // A 128x64x16 MMA is worth ~32 tensor-pipe cycles, but one thread needs ~80 cycles of dependent uniform-datapath
    // instructions to issue it; with two M-tiles per CTA tile each gets its own issuing thread (own accumulators, the
    // same shared-memory stage; a stage is released by both commits).
    const int m = warp == 1 ? 0 : 1;
    if (lane == 0 && m < P.mt) {
      const uint32_t idesc = make_idesc_bf16_k(128, P.N);
      int s = 0, acc = 0;
      uint32_t ph = 0, acc_ph = 0;
      const uint32_t h_row16 = (uint32_t)P.bw * 8u;  // one H row of the A box in 16-byte units (a multiple of 1 KB when used)
      for (long long tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        umma::mbar_wait(bar_acc_empty(acc), acc_ph ^ 1u);
        umma::fence_after_sync();
        const uint32_t d = tmem_base + (uint32_t)((acc * P.mt + m) * P.N);
        for (int kb = 0; kb < k_blocks; ++kb) {
          umma::mbar_wait(bar_full(s), ph);
          umma::fence_after_sync();
          const uint32_t a0 = base + (uint32_t)s * stage_bytes, b0 = a0 + P.a_bytes;
          const uint32_t first = kb != 0;
          if (P.group == 3) issue_stage<3>(a0, b0, d, (uint32_t)P.N, h_row16, (uint32_t)(m * P.bh), idesc, first);
          else issue_stage<1>(a0, b0, d, (uint32_t)P.N, h_row16, (uint32_t)(m * P.bh), idesc, first);
          umma::mma_commit(bar_empty(s));  // the stage is free again once both issuers' MMAs have read it
          if (++s == P.stages) { s = 0; ph ^= 1u; }
        }
        umma::mma_commit(bar_acc_full(acc));
        if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
      }
    }
  } else {
    // ===== epilogue: TMEM -> registers -> affine (+ReLU) -> global =====
    // 8 warps: warp w reads TMEM lanes 32 (w % 4) .. +31 (= tile rows) and owns half of the N columns. One warp per
    // scheduler could not hide its own dependency latencies: with N = 256 and a short K loop (the transposed
    // convolutions) the epilogue, not the tensor pipe, set the pace (ncu: 12 cycles per issued instruction).
    const int q = warp & 3, row = 32 * q + lane, half = (warp - 2) >> 2;
    const int ncols = P.N >= 64 ? P.N / 2 : (half == 0 ? P.N : 0), col0 = P.N >= 64 ? half * (P.N / 2) : 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    const int sh = P.shuffle;
    for (long long tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(P, tile);
      const int grp = t.nt / P.nsub;  // pixel-shuffle group (i, j); its out_c channels may span nsub N-tiles
      const int si = sh > 1 ? grp / sh : 0, sj = sh > 1 ? grp % sh : 0;
      const int n0 = sh > 1 ? (t.nt % P.nsub) * P.N : t.nt * P.N;  // scale/shift index and output channel of column 0
      umma::mbar_wait(bar_acc_full(acc), acc_ph);
      umma::fence_after_sync();
      for (int m = 0; m < P.mt; ++m) {
        const int oh = t.oh0 + m * P.bh + (row >> P.bw_log2), ow = t.ow0 + (row & (P.bw - 1));
        epilogue_rows(P, t, oh, ow, si, sj, n0, col0, ncols,
                      tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)((acc * P.mt + m) * P.N + col0));
      }
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(bar_acc_empty(acc));
      if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc<512>(tmem_base);
}

// ---- halo plans (group_kh = 2): stride-1 3x3(x3) convolutions with every kh x kw tap read from ONE input box ---------
// tcgen05.mma applies the 128-byte swizzle to the ABSOLUTE shared-memory address (tools/umma_shift_probe.cu: a K-major
// SWIZZLE_128B operand may start at any 128-byte row, with its 8-row groups any multiple of 16 bytes apart, base offset
// 0). So a tap is just a descriptor: the box [18 x 18 positions x 64 ch] (a 16 x 16 tile plus a 1-position halo) is
// loaded once per (kd, 64-channel block); M-tile m (8 wide, 16 high, two side by side) under tap (kh, kw) starts at box
// row kh * 18 + kw + 8 m and its 8-row groups are one box line (18 rows = 2304 bytes) apart. Input traffic per output
// drops 2.7x against the kh-halo plans; the weights stream through their own ring, three kw taps per slot.
// GATHER variant (first Conv3D behind the VFE stack): up to four more warps (11-14) build the input boxes themselves. A box is
// 18 x 18 positions x 64 bf16 channels = 324 rows of 128 bytes in the layout the TMA would have delivered (row r =
// h * 18 + w, 16-byte chunks XOR-swizzled with bits 7..9 of the row's absolute shared-memory address). A lane takes rows
// lane, lane + 32, ...: cell -> voxel through the occupancy map, then the voxel's float32 row rounded to bf16 (the same
// one rounding the dense bf16 grid holds), or c_empty, or zeros outside the grid (ZeroPadding3D, :192). 92 % of the cells
// are empty: a slot is filled with c_empty rows once and a box rewrites only what differs (occupied rows, rows outside
// the grid, rows its predecessor in the slot changed). Gather plans have three box slots by default (LISEC_GATHER_SLOTS);
// box n goes to warp and slot n mod 3 (a box is two dependent L2 round trips — occupancy words, then the occupied rows,
// compacted over the warp so that they cost one trip instead of eleven); the weights keep coming through warp 0's TMA.
constexpr int kGatherWarps = 4;  // one per input-box slot
constexpr int kGatherList = 352;  // (row, voxel) entries of a warp's list of occupied rows: every row of a box
constexpr int kConvGatherThreads = kConvThreads + 32 * kGatherWarps;
template <bool GATHER>
__global__ void __launch_bounds__(GATHER ? kConvGatherThreads : kConvThreads, 1)
    conv_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     const __grid_constant__ ConvParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = umma::smem_u32(smem);
  if (base & 1023u) __trap();
  const uint32_t b_slot_bytes = 3u * (uint32_t)P.N * 128u;
  const uint32_t b_base = base + (uint32_t)P.a_slots * P.a_slot_bytes;
  const uint32_t bar0 = b_base + (uint32_t)P.b_slots * b_slot_bytes;
  auto bar_a_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_a_empty = [&](int s) { return bar0 + 8u * (4 + s); };
  auto bar_b_full = [&](int s) { return bar0 + 8u * (8 + s); };
  auto bar_b_empty = [&](int s) { return bar0 + 8u * (16 + s); };
  auto bar_acc_full = [&](int a) { return bar0 + 8u * (24 + a); };
  auto bar_acc_empty = [&](int a) { return bar0 + 8u * (26 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (bar0 - base) + 8 * 28);
  uint4* s_cempty = reinterpret_cast<uint4*>(smem + (bar0 - base) + 8 * 28 + 16);  // GATHER: c_empty as 64 bf16
  if (GATHER && warp == 11 && lane < 8) {
    const float* c = P.g_c_empty + 8 * lane;
    s_cempty[lane] = make_uint4(pack2(c[0], c[1]), pack2(c[2], c[3]), pack2(c[4], c[5]), pack2(c[6], c[7]));
  }
  if (warp == 0 && lane == 0) {
    if (!GATHER) prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    for (int s = 0; s < P.a_slots; ++s) {
      umma::mbar_init(bar_a_full(s), 1);
      umma::mbar_init(bar_a_empty(s), P.mt);  // one commit per issuing thread
    }
    for (int s = 0; s < P.b_slots; ++s) {
      umma::mbar_init(bar_b_full(s), 1);
      umma::mbar_init(bar_b_empty(s), P.mt);
    }
    for (int a = 0; a < 2; ++a) {
      umma::mbar_init(bar_acc_full(a), P.mt);
      umma::mbar_init(bar_acc_empty(a), 8);
    }
    umma::mbar_init_fence();
  }
  if (warp == 1) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int a_steps = P.kd_n * P.c_blocks;  // input boxes per tile; each is followed by 3 weight slots (kh = 0, 1, 2)
  if (warp == 0) {
    if (lane == 0) {
      int as = 0, bs = 0;
      uint32_t aph = 0, bph = 0;
      for (long long tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile(P, tile);
        for (int kd = 0; kd < P.kd_n; ++kd)
          for (int cb = 0; cb < P.c_blocks; ++cb) {
            if (!GATHER) {
              umma::mbar_wait(bar_a_empty(as), aph ^ 1u);
              mbar_arrive_expect_tx(bar_a_full(as), P.a_box_bytes);
              tma_load_5d(base + (uint32_t)as * P.a_slot_bytes, &map_a, bar_a_full(as), 64 * cb, t.ow0 + P.t1[0],
                          t.oh0 + P.t2[0], t.od * P.stride_d + kd + P.t3[0], t.b);
              if (++as == P.a_slots) { as = 0; aph ^= 1u; }
            }
            for (int kh = 0; kh < 3; ++kh) {
              umma::mbar_wait(bar_b_empty(bs), bph ^ 1u);
              mbar_arrive_expect_tx(bar_b_full(bs), b_slot_bytes);
              tma_load_3d(b_base + (uint32_t)bs * b_slot_bytes, &map_b, bar_b_full(bs), 64 * cb, t.nt * P.N,
                          (kd * 3 + kh) * 3);
              if (++bs == P.b_slots) { bs = 0; bph ^= 1u; }
            }
          }
      }
    }
  } else if (warp == 1 || warp == kSecondIssuerWarp) {
    const int m = warp == 1 ? 0 : 1;  // one issuing thread per M-tile (see conv_igemm_kernel)
    if (lane == 0 && m < P.mt) {
      const uint32_t idesc = make_idesc_bf16_k(128, P.N);
      const uint64_t proto = umma::make_desc_k_sw128(0);
      const uint32_t lo0 = (uint32_t)proto, hi_b = (uint32_t)(proto >> 32);
      // A: 8-row groups one box line apart (stride byte offset = box_w * 128, in 16-byte units in bits 0..13 of the high word)
      const uint32_t hi_a = (hi_b & ~0x3fffu) | (((uint32_t)P.box_w * 128u) >> 4);
      const uint32_t line16 = (uint32_t)P.box_w * 8u, b_tap16 = (uint32_t)P.N * 8u;
      int as = 0, bs = 0, acc = 0;
      uint32_t aph = 0, bph = 0, acc_ph = 0;
      for (long long tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
        umma::mbar_wait(bar_acc_empty(acc), acc_ph ^ 1u);
        umma::fence_after_sync();
        const uint32_t d = tmem_base + (uint32_t)((acc * P.mt + m) * P.N);
        for (int st = 0; st < a_steps; ++st) {
          umma::mbar_wait(bar_a_full(as), aph);
          const uint32_t a_lo0 = lo0 + ((base + (uint32_t)as * P.a_slot_bytes) >> 4);
          for (int kh = 0; kh < 3; ++kh) {
            umma::mbar_wait(bar_b_full(bs), bph);
            umma::fence_after_sync();
            const uint32_t b_lo0 = lo0 + ((b_base + (uint32_t)bs * b_slot_bytes) >> 4);
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
              const uint32_t a_lo = a_lo0 + (uint32_t)kh * line16 + (uint32_t)(kw + 8 * m) * 8u;
              const uint32_t b_lo = b_lo0 + (uint32_t)kw * b_tap16;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t accum = (st | kh | kw | j) != 0;
                asm volatile(
                    "{\n\t"
                    ".reg .pred p;\n\t"
                    ".reg .b64 da, db;\n\t"
                    "mov.b64 da, {%1, %3};\n\t"
                    "mov.b64 db, {%2, %4};\n\t"
                    "setp.ne.b32 p, %6, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
                    "}\n" ::"r"(d),
                    "r"(a_lo + 2 * j), "r"(b_lo + 2 * j), "r"(hi_a), "r"(hi_b), "r"(idesc), "r"(accum)
                    : "memory");
              }
            }
            umma::mma_commit(bar_b_empty(bs));
            if (++bs == P.b_slots) { bs = 0; bph ^= 1u; }
          }
          umma::mma_commit(bar_a_empty(as));
          if (++as == P.a_slots) { as = 0; aph ^= 1u; }
        }
        umma::mma_commit(bar_acc_full(acc));
        if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
      }
    }
  } else if (GATHER && warp >= 11 && warp - 11 < P.a_slots) {
    const int g = warp - 11;  // boxes g, g + a_slots, ... of this CTA's sequence, always into slot g
    const int rows = P.box_w * (P.bh + 2);
    unsigned char* slot = smem + (size_t)g * P.a_slot_bytes;
    const uint32_t slot_addr = base + (uint32_t)g * P.a_slot_bytes;
    long long n = 0;     // box ordinal of this CTA
    uint32_t use = 0;    // how often this warp has filled its slot
    // 92 % of the cells are empty: the slot is filled with c_empty rows ONCE, and a box then only writes the rows that are
    // occupied or outside the grid, plus the rows its predecessor in this slot left different from c_empty (`dirty`: bit k
    // = the lane's row lane + 32 k) — ~60 rows of 8 stores instead of 324.
    unsigned dirty = 0;
    int2* s_list = reinterpret_cast<int2*>(reinterpret_cast<unsigned char*>(s_cempty) + 128) + (size_t)g * kGatherList;
    // a row's 16-byte chunks are XOR-swizzled with bits 7..9 of its absolute address; slots are 1 KB aligned, so that is
    // row & 7 — for the lane's own rows (lane + 32 k) simply lane & 7
    auto store_row = [&](int r, const uint4 (&ch)[8]) {
      const uint32_t sw = (uint32_t)r & 7u;
      unsigned char* row_ptr = slot + (size_t)r * 128;
#pragma unroll
      for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(row_ptr + (((uint32_t)c ^ sw) << 4)) = ch[c];
    };
    // what does not depend on the box: (h, w) of the lane's rows inside a box (the only division), packed h << 16 | w
    int hw[11];
    unsigned valid = 0;
#pragma unroll
    for (int k = 0; k < 11; ++k) {
      const int r = lane + 32 * k;
      const int hh = r / P.box_w;
      hw[k] = (hh << 16) | (r - hh * P.box_w);
      if (r < rows) valid |= 1u << k;
    }
    {
      uint4 ce[8];  // (s_cempty was written before the set-up barrier)
#pragma unroll
      for (int c = 0; c < 8; ++c) ce[c] = s_cempty[c];
      for (int r = lane; r < rows; r += 32) store_row(r, ce);
    }
    for (long long tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(P, tile);
      for (int kd = 0; kd < P.kd_n; ++kd, ++n) {  // (c_blocks == 1 for gather plans)
        if ((int)(n % P.a_slots) != g) continue;
        const int w0 = t.ow0 + P.t1[0], h0 = t.oh0 + P.t2[0], dz = t.od * P.stride_d + kd + P.t3[0];
        const bool d_ok = dz >= 0 && dz < P.in_d;
        const int* cells = P.g_cell_voxel + (((long long)t.b * P.in_d + dz) * P.in_h + h0) * P.in_w + w0;
        // the occupancy words of the lane's rows first (independent loads), before waiting for the slot
        int vox[11];
#pragma unroll
        for (int k = 0; k < 11; ++k) {
          const int hh = hw[k] >> 16, ww = hw[k] & 0xffff;
          const bool ok = ((valid >> k) & 1u) && d_ok && (unsigned)(h0 + hh) < (unsigned)P.in_h &&
                          (unsigned)(w0 + ww) < (unsigned)P.in_w;
          vox[k] = ok ? __ldcg(cells + hh * P.in_w + ww) : -2;  // -1: empty cell, -2: outside the grid
        }
        umma::mbar_wait(bar_a_empty(g), (use & 1u) ^ 1u);
        // rows without a voxel behind them: zeros outside the grid, c_empty back where the slot's last box left something
        // else (one store path for both). The occupied rows are only LISTED here (row, voxel), compacted over the warp:
        // fetching them inside this loop costs one dependent L2 round trip per k for the whole warp (some lane has an
        // occupied row at almost every k) — eleven per box. (Fetching the list's first 32 rows BEFORE the wait, so that
        // the slot is not held for a round trip, measured slower: 0.78 against 0.70 ms.)
        int n_occ = 0;
        const uint32_t lane_sw = (uint32_t)lane & 7u;
#pragma unroll
        for (int k = 0; k < 11; ++k) {
          const unsigned bit = 1u << k;
          const bool live = (valid & bit) != 0;
          const bool occ = live && vox[k] >= 0;
          const unsigned m = __ballot_sync(0xffffffffu, occ);
          if (occ) s_list[n_occ + __popc(m & ((1u << lane) - 1u))] = make_int2(lane + 32 * k, vox[k]);
          n_occ += __popc(m);
          const bool zero = live && vox[k] == -2;
          if (zero || (live && !occ && (dirty & bit))) {
            unsigned char* row_ptr = slot + (size_t)(lane + 32 * k) * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c)
              *reinterpret_cast<uint4*>(row_ptr + (((uint32_t)c ^ lane_sw) << 4)) = zero ? make_uint4(0u, 0u, 0u, 0u) : s_cempty[c];
          }
          dirty = (occ || zero) ? (dirty | bit) : (dirty & ~bit);
        }
        __syncwarp();
        for (int i0 = 0; i0 < n_occ; i0 += 32) {  // one occupied row per lane and round: a single round trip for ~26 rows
          const int i = i0 + lane;
          if (i < n_occ) {
            const int2 e = s_list[i];
            const float4* src = reinterpret_cast<const float4*>(P.g_voxel_feat + (size_t)e.y * 64);
            float4 f[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) f[c] = __ldcg(src + c);
            uint4 ch[8];
#pragma unroll
            for (int c = 0; c < 8; ++c)
              ch[c] = make_uint4(pack2(f[2 * c].x, f[2 * c].y), pack2(f[2 * c].z, f[2 * c].w), pack2(f[2 * c + 1].x, f[2 * c + 1].y),
                                 pack2(f[2 * c + 1].z, f[2 * c + 1].w));
            store_row(e.x, ch);
          }
        }
        __syncwarp();  // (the list is rewritten by the next box)
        umma::fence_async_smem();  // generic-proxy writes -> visible to the tensor core
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(bar_a_full(g));
        ++use;
      }
    }
  } else if (GATHER && warp >= 11) {
    // a gather warp without a box slot (a_slots < kGatherWarps): nothing to do
  } else {
    const int q = warp & 3, row = 32 * q + lane, half = (warp - 2) >> 2;
    const int ncols = P.N >= 64 ? P.N / 2 : (half == 0 ? P.N : 0), col0 = P.N >= 64 ? half * (P.N / 2) : 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    for (long long tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile(P, tile);
      umma::mbar_wait(bar_acc_full(acc), acc_ph);
      umma::fence_after_sync();
      for (int m = 0; m < P.mt; ++m)  // M-tile m: 8 positions wide, 16 high, at w offset 8 m
        epilogue_rows(P, t, t.oh0 + (row >> 3), t.ow0 + 8 * m + (row & 7), 0, 0, t.nt * P.N, col0, ncols,
                      tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)((acc * P.mt + m) * P.N + col0));
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(bar_acc_empty(acc));
      if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc<512>(tmem_base);
}

// ---- float32 plans: 3xTF32 with the partial sums promoted to the FP32 pipe every 64 channels --------------------
// The tensor core's float32 accumulator does not round to nearest: a sum of several hundred MMAs into one accumulator
// drifts by ~1e-5..1e-4 of its magnitude (measured 6e-5 on the K = 1728 Conv3D), past north_star's 1e-5. So an
// accumulator only ever takes ONE 32-channel stage (12 MMAs; 24 per accumulator left the whole network at 1.0e-5, on
// the bar); the epilogue warps add the chunks in registers (round-to-nearest FADD) while the next chunk's MMAs run into the other
// accumulator. A thread keeps its row's N running sums in registers, so float32 plans take out_c <= 128; wider layers
// run as several N-tiles (n_tiles, also inside a pixel-shuffle group).
constexpr int kConvF32Threads = 192;

__global__ void __launch_bounds__(kConvF32Threads, 1)
    conv_igemm_f32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                          const __grid_constant__ ConvParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = umma::smem_u32(smem);
  if (base & 1023u) __trap();
  const uint32_t b_tap_bytes = (uint32_t)P.N * 128u;
  const uint32_t stage_bytes = P.stage_bytes;
  const uint32_t bar0 = base + (uint32_t)P.stages * stage_bytes;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kMaxStages + s); };
  auto bar_acc_full = [&](int a) { return bar0 + 8u * (2 * kMaxStages + a); };
  auto bar_acc_empty = [&](int a) { return bar0 + 8u * (2 * kMaxStages + 2 + a); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (size_t)P.stages * stage_bytes + 8 * (2 * kMaxStages + 4));
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_b);
    for (int s = 0; s < P.stages; ++s) {
      umma::mbar_init(bar_full(s), 1);
      umma::mbar_init(bar_empty(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      umma::mbar_init(bar_acc_full(a), 1);
      umma::mbar_init(bar_acc_empty(a), 4);
    }
    umma::mbar_init_fence();
  }
  if (warp == 1) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int chunks = P.n_taps * P.c_blocks;  // one 32-channel stage per chunk
  if (warp == 0) {
    if (lane == 0) producer_loop(P, map_a, map_b, base, bar0);
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma::make_idesc_tf32_k(128, P.N);
      int s = 0, acc = 0;
      uint32_t ph = 0, acc_ph = 0;
      for (long long tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x)
        for (int c = 0; c < chunks; ++c) {
          umma::mbar_wait(bar_acc_empty(acc), acc_ph ^ 1u);
          umma::fence_after_sync();
          const uint32_t d0 = tmem_base + (uint32_t)(acc * P.N);
          umma::mbar_wait(bar_full(s), ph);
          umma::fence_after_sync();
          const uint32_t a0 = base + (uint32_t)s * stage_bytes;
          issue_stage_tf32(a0, P.a_bytes, a0 + 2 * P.a_bytes, b_tap_bytes, d0, idesc, 0u);
          umma::mma_commit(bar_empty(s));
          if (++s == P.stages) { s = 0; ph ^= 1u; }
          umma::mma_commit(bar_acc_full(acc));
          if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
        }
    }
  } else {
    const int q = warp & 3, row = 32 * q + lane;
    const int ncols = P.N, col0 = 0;
    int acc = 0;
    uint32_t acc_ph = 0;
    const int sh = P.shuffle;
    for (long long tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
      float sum[8][16];
      for (int c = 0; c < chunks; ++c) {
        umma::mbar_wait(bar_acc_full(acc), acc_ph);
        umma::fence_after_sync();
        const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * P.N + col0);
        if (ncols == 16) {
          float x[16];
          tmem_ld_32x16(taddr, x);
#pragma unroll
          for (int i = 0; i < 16; ++i) sum[0][i] = c ? __fadd_rn(sum[0][i], x[i]) : x[i];
        } else {
#pragma unroll
          for (int g = 0; g < 4; ++g)  // 32 columns per TMEM load (128 running sums + 32 fresh values in registers)
            if (32 * g < ncols) {
              float x[32];
              umma::tmem_ld_32x32(taddr + 32 * g, x);
#pragma unroll
              for (int i = 0; i < 32; ++i) sum[2 * g + (i >> 4)][i & 15] = c ? __fadd_rn(sum[2 * g + (i >> 4)][i & 15], x[i]) : x[i];
            }
        }
        umma::fence_before_sync();
        __syncwarp();
        if (lane == 0) umma::mbar_arrive(bar_acc_empty(acc));
        if (++acc == 2) { acc = 0; acc_ph ^= 1u; }
      }
      const TileCoord t = decode_tile(P, tile);
      const int grp = t.nt / P.nsub;
      const int si = sh > 1 ? grp / sh : 0, sj = sh > 1 ? grp % sh : 0;
      const int n0 = (sh > 1 ? (t.nt % P.nsub) * P.N : t.nt * P.N) + col0;
      const int oh = t.oh0 + (row >> P.bw_log2), ow = t.ow0 + (row & (P.bw - 1));
      const bool valid = oh < P.out_h && ow < P.out_w;
      const size_t pix = (((size_t)t.b * P.out_d + t.od) * ((size_t)P.out_h * sh) + (size_t)oh * sh + si) *
                             ((size_t)P.out_w * sh) + (size_t)ow * sh + sj;
      const size_t elem = pix * (size_t)P.out_pitch + P.out_ch_off + n0;
#pragma unroll
      for (int g = 0; g < 8; ++g)
        if (16 * g < ncols) store_chunk<16>(P, sum[g], n0 + 16 * g, elem + 16 * g, valid);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc<512>(tmem_base);
}

// x -> (hi, lo) tf32 operand planes of the float32 plans (the front end's float32 grid feeds the first Conv3D)
__global__ void __launch_bounds__(256) split_tf32_kernel(const float4* __restrict__ x, float4* __restrict__ hi,
                                                         float4* __restrict__ lo, long long n4) {
  pdl_launch_dependents();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldcg(x + i);
    float4 h, l;
    umma::tf32_split(v.x, h.x, l.x);
    umma::tf32_split(v.y, h.y, l.y);
    umma::tf32_split(v.z, h.z, l.z);
    umma::tf32_split(v.w, h.w, l.w);
    hi[i] = h;
    lo[i] = l;
  }
}

// ---- host side ------------------------------------------------------------------------------------------------
// ---- heads combine -------------------------------------------------------------------------------------------------
// Conv2DTranspose (no activation, :247-251) -> Concatenate (:252) -> the two 1x1 head convolutions (:253-254) is ONE
// linear map per RPN block, so the host folds each transposed kernel with its 256 rows of the head kernels
// (lisec_b200/network.py) and the three blocks leave small float32 tensors: c1 [B,H,W,n] at the output resolution (from
// the k3 s1 block; carries every bias), c2 [B,H/s2,W/s2,s2*s2*n] and c3 [B,H/s3,W/s3,s3*s3*n] whose channel group
// (i*s + j) belongs to output pixel (s*h + i, s*w + j). This kernel adds them: the 768-channel concat tensor (245 MB per 8
// sweeps, written once and read once) never exists. Thread = (pixel, 4 channels).
__global__ void __launch_bounds__(256)
    heads_combine_kernel(const float* __restrict__ c1, const float* __restrict__ c2, int s2, const float* __restrict__ c3,
                         int s3, float* __restrict__ out, int batch, int H, int W, int n) {
  pdl_launch_dependents();
  const int q = n >> 2;
  const long long total = (long long)batch * H * W * q;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  pdl_wait();
  if (gid >= total) return;
  const int c = (int)(gid % q);
  const long long pix = gid / q;
  const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((long long)W * H));
  float4 v = __ldcg(reinterpret_cast<const float4*>(c1 + pix * n) + c);  // (the parts were written by the plans in front)
  {
    const int h2 = H / s2, w2 = W / s2;
    const size_t p = ((size_t)b * h2 + y / s2) * w2 + x / s2;
    const float4 a = __ldcg(reinterpret_cast<const float4*>(c2 + (p * (s2 * s2) + (y % s2) * s2 + (x % s2)) * n) + c);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
  }
  {
    const int h3 = H / s3, w3 = W / s3;
    const size_t p = ((size_t)b * h3 + y / s3) * w3 + x / s3;
    const float4 a = __ldcg(reinterpret_cast<const float4*>(c3 + (p * (s3 * s3) + (y % s3) * s3 + (x % s3)) * n) + c);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
  }
  reinterpret_cast<float4*>(out + pix * n)[c] = v;
}

thread_local char g_conv_error[512] = "";

int conv_fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_conv_error, sizeof(g_conv_error), fmt, ap);
  va_end(ap);
  return status;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int floor_div(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

}  // namespace
}  // namespace lisec

using namespace lisec;

struct lisec_conv_plan {
  CUtensorMap map_a, map_b;
  ConvParams p;
  int grid, smem, device;
};

extern "C" {

const char* lisec_conv_last_error(void) { return g_conv_error; }

int32_t lisec_conv_plan_create(const lisec_conv_desc* d, const void* in, const void* weights, const float* scale,
                               const float* shift, void* out, lisec_conv_plan** plan_out) {
  if (!d || !in || !weights || !scale || !shift || !out || !plan_out)
    return conv_fail(LISEC_ERR_BAD_ARG, "null argument");
  *plan_out = nullptr;
  const int C = d->in_c, N = d->out_c;
  const bool f32 = d->in_dtype == LISEC_F32;
  if (!f32 && d->in_dtype != LISEC_BF16) return conv_fail(LISEC_ERR_BAD_CONFIG, "in_dtype: LISEC_BF16 or LISEC_F32");
  if (f32 && (d->m_tiles > 1 || d->group_kh || d->out_dtype != LISEC_F32))
    return conv_fail(LISEC_ERR_BAD_CONFIG, "float32 plans: m_tiles = 1, group_kh = 0, float32 output");
  if (!f32 && d->out_split) return conv_fail(LISEC_ERR_BAD_CONFIG, "out_split is the float32 plans' hi/lo output");
  const int cpb = f32 ? 32 : 64;
  if (C < 64 || C % 64) return conv_fail(LISEC_ERR_BAD_CONFIG, "in_c = %d: need a multiple of 64", C);
  if (N < 16 || N > 256 || N % 16) return conv_fail(LISEC_ERR_BAD_CONFIG, "out_c = %d: need 16..256, multiple of 16", N);
  if (N != 16 && N % 32) return conv_fail(LISEC_ERR_BAD_CONFIG, "out_c = %d: need 16 or a multiple of 32", N);
  const int taps = d->kd * d->kh * d->kw;
  if (taps < 1 || taps > kMaxTaps) return conv_fail(LISEC_ERR_BAD_CONFIG, "%d taps: at most %d", taps, kMaxTaps);
  const int s = d->stride_hw;
  if (s != 1 && s != 2) return conv_fail(LISEC_ERR_BAD_CONFIG, "stride_hw = %d: 1 or 2", s);
  if (s == 2 && (d->in_d != 1 || d->kd != 1 || d->in_h % 2 || d->in_w % 2 || 2 * C > 32767))
    return conv_fail(LISEC_ERR_BAD_CONFIG, "stride-2 layers are 2-D with even H and W");
  if (d->tile_w * d->tile_h != 128 || (d->tile_w & (d->tile_w - 1)) || d->tile_w > 128)
    return conv_fail(LISEC_ERR_BAD_CONFIG, "tile %d x %d: need a power-of-two width and 128 positions", d->tile_w,
                     d->tile_h);
  const int n_tiles = d->n_tiles < 1 ? 1 : d->n_tiles;
  const int shuffle = d->shuffle < 1 ? 1 : d->shuffle;
  const int mt = d->m_tiles < 1 ? 1 : d->m_tiles;
  const bool halo = d->group_kh == 2;
  const int group = (d->group_kh == 1) ? d->kh : 1;
  if (group != 1 && group != 3) return conv_fail(LISEC_ERR_BAD_CONFIG, "group_kh needs kh = 3");
  if (halo && (f32 || s != 1 || d->kh != 3 || d->kw != 3 || d->kd > 3 || N > 128 || shuffle != 1 || d->tile_w != 8 ||
               d->tile_h != 16))
    return conv_fail(LISEC_ERR_BAD_CONFIG, "halo plans: bf16, stride_hw = 1, 3x3 taps, out_c <= 128, tile 8 x 16");
  if (mt > 2 || 2 * mt * N > 512) return conv_fail(LISEC_ERR_BAD_CONFIG, "m_tiles = %d with out_c = %d: 2 * m_tiles * out_c accumulator columns must fit 512", mt, N);
  if ((mt > 1 || group > 1) && d->tile_w < 8)
    return conv_fail(LISEC_ERR_BAD_CONFIG, "m_tiles / group_kh need tile_w >= 8 (1 KB-aligned H rows)");
  if (group > 1 && (s != 1 || d->kh < 2)) return conv_fail(LISEC_ERR_BAD_CONFIG, "group_kh needs stride_hw = 1 and kh > 1");
  if (d->tile_h * mt + group - 1 > 256) return conv_fail(LISEC_ERR_BAD_CONFIG, "input box taller than 256 rows");
  if (shuffle > 1 && (n_tiles % (shuffle * shuffle) || taps != 1))
    return conv_fail(LISEC_ERR_BAD_CONFIG, "pixel shuffle %d needs a multiple of %d N-tiles and a 1x1 kernel", shuffle,
                     shuffle * shuffle);
  const int nsub = shuffle > 1 ? n_tiles / (shuffle * shuffle) : 1;
  if (f32 && (N > 128 || (N != 16 && N % 64)))
    return conv_fail(LISEC_ERR_BAD_CONFIG, "float32 plans take out_c = 16, 64 or 128 per N-tile (got %d)", N);
  if (d->batch < 1 || d->in_d < 1 || d->in_h < 1 || d->in_w < 1 || d->stride_d < 1)
    return conv_fail(LISEC_ERR_BAD_ARG, "bad input shape");
  const int OD = (d->in_d + 2 * d->pad_d - d->kd) / d->stride_d + 1;
  const int OH = (d->in_h + 2 * d->pad_h - d->kh) / s + 1;
  const int OW = (d->in_w + 2 * d->pad_w - d->kw) / s + 1;
  if (OD < 1 || OH < 1 || OW < 1) return conv_fail(LISEC_ERR_BAD_CONFIG, "empty output");
  const int align = d->out_dtype == LISEC_F32 ? 4 : 8;
  if (d->out_pitch % align || d->out_ch_off % align || d->out_pitch < d->out_ch_off + (shuffle > 1 ? nsub * N : n_tiles * N))
    return conv_fail(LISEC_ERR_BAD_CONFIG, "out_pitch %d / out_ch_off %d: need multiples of %d and room for %d channels",
                     d->out_pitch, d->out_ch_off, align, N);
  if (d->out_dtype != LISEC_F32 && d->out_dtype != LISEC_BF16) return conv_fail(LISEC_ERR_BAD_CONFIG, "out_dtype");
  EncodeTiledFn encode = encode_tiled_fn();
  if (!encode) return conv_fail(LISEC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");

  lisec_conv_plan* pl = new lisec_conv_plan();
  memset(pl, 0, sizeof(*pl));
  ConvParams& p = pl->p;
  p.bw = d->tile_w;
  p.bh = d->tile_h;
  p.bw_log2 = 0;
  while ((1 << p.bw_log2) < p.bw) ++p.bw_log2;
  p.tiles_w = (OW + p.bw - 1) / p.bw;
  p.tiles_h = (OH + p.bh * mt - 1) / (p.bh * mt);
  p.step_w = p.bw;
  p.step_h = p.bh * mt;
  if (halo) {  // M-tiles side by side along W
    p.step_w = p.bw * mt;
    p.step_h = p.bh;
    p.tiles_w = (OW + p.step_w - 1) / p.step_w;
    p.tiles_h = (OH + p.step_h - 1) / p.step_h;
  }
  p.halo = halo;
  p.mt = mt;
  p.group = group;
  p.out_d = OD;
  p.batch = d->batch;
  p.n_tiles = n_tiles;
  p.total_tiles = (long long)n_tiles * p.tiles_w * p.tiles_h * OD * d->batch;
  p.n_taps = taps / group;
  p.c_blocks = C / cpb;
  p.f32 = f32;
  p.cpb = cpb;
  p.lo_batch = d->batch;
  p.lo_tap = taps;
  p.nsub = nsub;
  p.N = N;
  p.s2d = s == 2;
  p.stride_d = d->stride_d;
  int t = 0;
  if (halo) {  // one box per kd: its origin is the tile origin minus the padding
    p.c_off[0] = 0;
    p.t1[0] = (signed char)(-d->pad_w);
    p.t2[0] = (signed char)(-d->pad_h);
    p.t3[0] = (signed char)(-d->pad_d);
    p.kd_n = d->kd;
    p.box_w = p.bw * mt + 2;
  } else if (group > 1) {  // stages in (kd, kw) order; the box starts at the kh = 0 row and carries the kh-1 halo rows
    for (int kd = 0; kd < d->kd; ++kd)
      for (int kw = 0; kw < d->kw; ++kw, ++t) {
        p.c_off[t] = 0;
        p.t1[t] = (signed char)(kw - d->pad_w);
        p.t2[t] = (signed char)(-d->pad_h);
        p.t3[t] = (signed char)(kd - d->pad_d);
      }
  } else {
    for (int kd = 0; kd < d->kd; ++kd)
      for (int kh = 0; kh < d->kh; ++kh)
        for (int kw = 0; kw < d->kw; ++kw, ++t) {
          if (s == 1) {
            p.c_off[t] = 0;
            p.t1[t] = (signed char)(kw - d->pad_w);
            p.t2[t] = (signed char)(kh - d->pad_h);
            p.t3[t] = (signed char)(kd - d->pad_d);
          } else {  // input position 2*o + (k - pad) = 2*(o + q) + parity
            const int uw = kw - d->pad_w, uh = kh - d->pad_h;
            const int qw = floor_div(uw, 2), qh = floor_div(uh, 2);
            p.c_off[t] = (short)((uw - 2 * qw) * C);
            p.t1[t] = (signed char)qw;
            p.t2[t] = (signed char)(uh - 2 * qh);
            p.t3[t] = (signed char)qh;
          }
        }
  }
  p.out_h = OH;
  p.out_w = OW;
  p.shuffle = shuffle;
  p.relu = d->relu != 0;
  p.out_f32 = d->out_dtype == LISEC_F32;
  p.out_ch_off = d->out_ch_off;
  p.out_pitch = d->out_pitch;
  p.scale = scale;
  p.shift = shift;
  p.affine_rewritten = d->reserved & 1;
  p.out = out;
  p.gather = 0;
  p.in_d = d->in_d;
  p.in_h = d->in_h;
  p.in_w = d->in_w;
  p.g_cell_voxel = nullptr;
  p.g_voxel_feat = nullptr;
  p.g_c_empty = nullptr;
  // hi / lo output planes: [2][batch, out_d, out_h*shuffle, out_w*shuffle, out_pitch]
  p.out_lo_off = d->out_split ? (long long)d->batch * OD * ((long long)OH * shuffle) * ((long long)OW * shuffle) * d->out_pitch : 0;
  int box_h = p.bh * mt + group - 1, box_w = p.bw;
  if (halo) {
    box_h = p.bh + 2;
    box_w = p.box_w;
    p.a_box_bytes = (uint32_t)(box_w * box_h) * 128u;
    p.a_slot_bytes = (p.a_box_bytes + 1023u) & ~1023u;
    p.a_slots = 2;
    const uint32_t b_slot = 3u * (uint32_t)N * 128u;
    const uint32_t room = 220u * 1024u - 2u * p.a_slot_bytes;
    int b_slots = (int)(room / b_slot);
    if (b_slots > 8) b_slots = 8;
    if (const char* e = getenv("LISEC_CONV_B_SLOTS")) {  // experiment: a shallower weight ring
      const int v = atoi(e);
      if (v >= 2 && v < b_slots) b_slots = v;
    }
    if (b_slots < 2) {
      delete pl;
      return conv_fail(LISEC_ERR_BAD_CONFIG, "halo plan: no room for two weight slots of %u bytes", b_slot);
    }
    p.b_slots = b_slots;
    p.a_bytes = p.a_box_bytes;
    p.stage_bytes = 0;
    p.stages = 0;
    pl->smem = (int)(2u * p.a_slot_bytes + (uint32_t)b_slots * b_slot) + 8 * 28 + 16 + 128;  // (+ c_empty for gather plans)
  } else {
    p.a_bytes = (uint32_t)(p.bw * box_h) * 128u;
    p.stage_bytes = (p.a_bytes + (uint32_t)(group * N) * 128u) * (f32 ? 2u : 1u);
    const uint32_t stage_bytes = p.stage_bytes;
    if (stage_bytes % 1024u) {
      delete pl;
      return conv_fail(LISEC_ERR_BAD_CONFIG, "stage of %u bytes is not 1 KB-aligned", stage_bytes);
    }
    int stages = (int)((220u * 1024u) / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
    if (stages < 2) {
      delete pl;
      return conv_fail(LISEC_ERR_BAD_CONFIG, "a stage of %u bytes leaves room for fewer than 2 stages", stage_bytes);
    }
    p.stages = stages;
    pl->smem = stages * (int)stage_bytes + 8 * (2 * kMaxStages + 4) + 16;
  }

  // tensor maps (bf16, 128-byte swizzle, zero fill out of bounds)
  const cuuint64_t eb = f32 ? 4 : 2;
  const CUtensorMapDataType dt = f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const cuuint64_t planes = f32 ? 2 : 1;  // hi / lo operand planes ride on the outermost dimension
  cuuint64_t dims[5], strides[4];
  cuuint32_t box[5], estr[5] = {1, 1, 1, 1, 1};
  const cuuint64_t W = d->in_w, H = d->in_h, D = d->in_d, B = d->batch;
  if (s == 1) {
    dims[0] = C; dims[1] = W; dims[2] = H; dims[3] = D; dims[4] = B * planes;
    strides[0] = C * eb; strides[1] = W * C * eb; strides[2] = H * W * C * eb; strides[3] = D * H * W * C * eb;
    box[0] = cpb; box[1] = box_w; box[2] = box_h; box[3] = 1; box[4] = 1;
  } else {
    dims[0] = 2 * C; dims[1] = W / 2; dims[2] = 2; dims[3] = H / 2; dims[4] = B * planes;
    strides[0] = 2 * C * eb; strides[1] = W * C * eb; strides[2] = 2 * W * C * eb; strides[3] = H * W * C * eb;
    box[0] = cpb; box[1] = p.bw; box[2] = 1; box[3] = box_h; box[4] = 1;
  }
  CUresult r = encode(&pl->map_a, dt, 5, const_cast<void*>(in), dims, strides, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    delete pl;
    return conv_fail(LISEC_ERR_CUDA, "cuTensorMapEncodeTiled(input) failed: CUresult %d", (int)r);
  }
  const cuuint64_t NT = (cuuint64_t)n_tiles * N;
  cuuint64_t wdims[3] = {(cuuint64_t)C, NT, (cuuint64_t)taps * planes};
  cuuint64_t wstr[2] = {C * eb, NT * C * eb};
  cuuint32_t wbox[3] = {(cuuint32_t)cpb, (cuuint32_t)N, (cuuint32_t)(halo ? 3 : group)}, westr[3] = {1, 1, 1};
  r = encode(&pl->map_b, dt, 3, const_cast<void*>(weights), wdims, wstr, wbox, westr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    delete pl;
    return conv_fail(LISEC_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: CUresult %d", (int)r);
  }
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);  // per function, not per plan: the opt-in maximum
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_igemm_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_halo_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(conv_halo_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e != cudaSuccess) {
    delete pl;
    return conv_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  }
  pl->device = dev;
  pl->grid = (int)(p.total_tiles < sms ? p.total_tiles : sms);
  *plan_out = pl;
  return LISEC_OK;
}

int32_t lisec_conv_plan_set_gather(lisec_conv_plan* pl, const int32_t* cell_voxel, const float* voxel_feat,
                                   const float* c_empty) {
  if (!pl || !cell_voxel || !voxel_feat || !c_empty) return conv_fail(LISEC_ERR_BAD_ARG, "null argument");
  if (!pl->p.halo || pl->p.c_blocks != 1 || pl->p.box_w * (pl->p.bh + 2) > 11 * 32)
    return conv_fail(LISEC_ERR_BAD_CONFIG, "gather source: a halo plan with 64 input channels and a box of <= 352 positions");
  // box slots (one per active gather warp) and what is left for the weight ring: 2 / 3 / 4 box slots measured 0.77 / 0.70 /
  // 0.71 ms (8 sweeps; 0.48 ms from the dense grid). The weight ring's depth is not what limits it (the dense plan with
  // two weight slots instead of five: 0.492 against 0.486 ms).
  int a_slots = 3;
  if (const char* e = getenv("LISEC_GATHER_SLOTS")) {
    const int v = atoi(e);
    if (v >= 2 && v <= kGatherWarps) a_slots = v;
  }
  const uint32_t b_slot = 3u * (uint32_t)pl->p.N * 128u;
  const uint32_t room = 214u * 1024u - (uint32_t)a_slots * pl->p.a_slot_bytes;
  int b_slots = (int)(room / b_slot);
  if (b_slots > 8) b_slots = 8;
  if (b_slots < 2) return conv_fail(LISEC_ERR_BAD_CONFIG, "gather source: no room for two weight slots beside the box slots");
  pl->p.a_slots = a_slots;
  pl->p.b_slots = b_slots;
  pl->smem = (int)((uint32_t)a_slots * pl->p.a_slot_bytes + (uint32_t)b_slots * b_slot) + 8 * 28 + 16 + 128 +
             kGatherWarps * kGatherList * (int)sizeof(int2);
  if (pl->smem > 232448) return conv_fail(LISEC_ERR_BAD_CONFIG, "gather source: %d bytes of shared memory", pl->smem);
  pl->p.gather = 1;
  pl->p.g_cell_voxel = cell_voxel;
  pl->p.g_voxel_feat = voxel_feat;
  pl->p.g_c_empty = c_empty;
  return LISEC_OK;
}

int32_t lisec_conv_plan_run(lisec_conv_plan* pl, void* stream) {
  if (!pl) return conv_fail(LISEC_ERR_BAD_ARG, "null plan");
  cudaError_t e = pl->p.halo && pl->p.gather
                      ? launch_pdl(conv_halo_kernel<true>, pl->grid, kConvGatherThreads, (size_t)pl->smem,
                                   static_cast<cudaStream_t>(stream), pl->map_a, pl->map_b, pl->p)
                  : pl->p.halo ? launch_pdl(conv_halo_kernel<false>, pl->grid, kConvThreads, (size_t)pl->smem,
                                          static_cast<cudaStream_t>(stream), pl->map_a, pl->map_b, pl->p)
                  : pl->p.f32 ? launch_pdl(conv_igemm_f32_kernel, pl->grid, kConvF32Threads, (size_t)pl->smem,
                                         static_cast<cudaStream_t>(stream), pl->map_a, pl->map_b, pl->p)
                            : launch_pdl(conv_igemm_kernel, pl->grid, kConvThreads, (size_t)pl->smem,
                                         static_cast<cudaStream_t>(stream), pl->map_a, pl->map_b, pl->p);
  if (e != cudaSuccess) return conv_fail(LISEC_ERR_CUDA, "conv launch: %s", cudaGetErrorString(e));
  return LISEC_OK;
}

int32_t lisec_heads_combine(const float* c1, const float* c2, int32_t s2, const float* c3, int32_t s3, float* out,
                            int32_t batch, int32_t out_h, int32_t out_w, int32_t n_ch, void* stream) {
  if (!c1 || !c2 || !c3 || !out) return conv_fail(LISEC_ERR_BAD_ARG, "null argument");
  if (batch < 0 || out_h <= 0 || out_w <= 0 || n_ch <= 0 || n_ch % 4 || s2 < 1 || s3 < 1 || out_h % s2 || out_w % s2 ||
      out_h % s3 || out_w % s3)
    return conv_fail(LISEC_ERR_BAD_CONFIG, "heads combine: n_ch a multiple of 4, out_h and out_w multiples of both strides");
  const long long total = (long long)batch * out_h * out_w * (n_ch / 4);
  if (total == 0) return LISEC_OK;
  cudaError_t e = launch_pdl(heads_combine_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0,
                             static_cast<cudaStream_t>(stream), c1, c2, (int)s2, c3, (int)s3, out, (int)batch, (int)out_h,
                             (int)out_w, (int)n_ch);
  if (e != cudaSuccess) return conv_fail(LISEC_ERR_CUDA, "heads combine: %s", cudaGetErrorString(e));
  return LISEC_OK;
}

int32_t lisec_split_tf32(const float* x, float* hi, float* lo, int64_t n, void* stream) {
  if (!x || !hi || !lo || n < 0 || n % 4) return conv_fail(LISEC_ERR_BAD_ARG, "split: null pointer or n not a multiple of 4");
  if (n == 0) return LISEC_OK;
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess)
    e = launch_pdl(split_tf32_kernel, sms * 8, 256, 0, static_cast<cudaStream_t>(stream),
                   reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(hi), reinterpret_cast<float4*>(lo),
                   (long long)(n / 4));
  if (e != cudaSuccess) return conv_fail(LISEC_ERR_CUDA, "split launch: %s", cudaGetErrorString(e));
  return LISEC_OK;
}

int32_t lisec_conv_plan_output_shape(const lisec_conv_plan* pl, int32_t* odhw) {
  if (!pl || !odhw) return conv_fail(LISEC_ERR_BAD_ARG, "null argument");
  odhw[0] = pl->p.out_d;
  odhw[1] = pl->p.out_h * pl->p.shuffle;
  odhw[2] = pl->p.out_w * pl->p.shuffle;
  return LISEC_OK;
}

void lisec_conv_plan_destroy(lisec_conv_plan* pl) { delete pl; }

}  // extern "C"

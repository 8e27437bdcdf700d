// Weight gradient of a convolution layer on the tensor cores — the first backward kernel of the training step (DESIGN.md
// §4e; reference: the gradients Keras computes for Conv3D / Conv2D kernels under model.fit, model_training.py:299).
//
//     dW[tap][co][ci] = sum over output positions p of  dY[p][co] * X[p * stride + tap - pad][ci]
//
// A GEMM whose K dimension is the POSITIONS: per 128-position tile, A = X^T (M = input channels), B = dY (N = output
// channels). Both tensors are channels-last, so both operands are "MN-major" — and a TMA box [128 positions][64 channels]
// with the 128-byte swizzle is exactly the canonical MN-major SWIZZLE_128B operand (tools/umma_mn_probe.cu: exact with
// instruction-descriptor bits 15/16 set, stride byte offset 1024, leading byte offset = distance between 64-channel boxes,
// +2048 bytes per K = 16 step). Nothing is transposed or gathered by threads; ZeroPadding is the TMA's zero fill.
//
// Work decomposition. A "unit" is one (filter tap, 64-input-channel block): one X box per tile. Two units stacked make
// the M = 128 rows of an MMA; N = all output channels (<= 256). A pair's accumulator D[128][N] lives in TMEM for a whole
// PASS over the CTA's tiles (512 columns hold 512 / N pairs); a layer needs ceil(pairs / (512 / N)) passes. Per (pass,
// tile) the dY boxes (N / 64 of them) land once in one of two dY slots and are shared by the pass's pairs; per pair one
// shared-memory stage carries [X unit a][X unit b] and feeds 8 MMAs (128 positions).
// Warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue: at the end of a pass they move the accumulators to
// this CTA's slice of a float32 partial buffer [grid][taps][N][C]; wgrad_reduce_kernel then adds the slices in a fixed
// order (deterministic — no atomics), giving dW in the layout the forward plans read their weights in.
//
// First version: bf16 operands, float32 accumulation, stride_hw 1 or 2 (the TMA's element stride), C and N multiples of
// 64, N <= 256. Each X box feeds
// only 8 MMAs, so this kernel is bound by L2 -> shared-memory delivery like the first forward kernel was
// (profiles/conv_r1j_summary.txt). conv_wgrad_halo_kernel (below) applies the halo-box trick of conv_halo_kernel to the
// layers that dominate the step — the 3 x 3 x 3, 64 -> 64 Conv3D blocks: ONE [18][10][64] input box per (kd, tile) serves
// all nine (kh, kw) taps, each tap a descriptor into the same box.
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

constexpr int kWgThreads = 192;
constexpr int kWgMaxStages = 6;
constexpr uint32_t kBox = 128 * 128;  // one [128 positions][64 channels] bf16 box

struct WgradParams {
  int tiles_w, tiles_h, out_d, batch;
  long long total_tiles;
  int bw, bh;
  int units, pairs, pairs_per_pass, passes, nb, N, C, taps, stages;
  uint32_t stage_bytes;  // one X stage: two boxes
  uint32_t dy_bytes;     // one dY slot: nb boxes (two slots, loaded once per tile and shared by the pass's pairs)
  int stride_d, stride_hw;
  signed char t1[27], t2[27], t3[27];  // X box origin of a tap relative to the tile origin (w, h, d)
  float* partial;  // [grid][taps][N][C]
  long long slice;  // taps * N * C
  // halo mode: one X box per (kd, 64-channel block, tile); kd_n * cblocks boxes = passes
  int halo, boxes, pad_w, pad_h, pad_d, kd_n;
};

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
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a), "l"(b), "r"(idesc), "r"(acc)
      : "memory");
}
// the same MMA with the descriptors as (low word, shared high word): the operand address lives in the low 14 bits of the
// low word, so a K step is a 32-bit add (the issuing thread's instruction stream paces small MMAs)
__device__ __forceinline__ void mma_bf16_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                            uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "setp.ne.b32 p, %5, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc), "r"(acc)
      : "memory");
}
// MN-major SWIZZLE_128B operand: 8 k-rows per 1 KB group (stride byte offset), 64-channel boxes `lbo` bytes apart
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t addr, uint32_t lbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// the same with an explicit stride between the 8-row k groups (a halo box: rows of 8 positions are a box row apart)
__device__ __forceinline__ uint64_t make_desc_mn2(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

struct WgTile {
  int ow0, oh0, od, b;
};
__device__ __forceinline__ WgTile wg_tile(const WgradParams& P, long long tile) {
  WgTile t;
  t.ow0 = (int)(tile % P.tiles_w) * P.bw;
  long long r = tile / P.tiles_w;
  t.oh0 = (int)(r % P.tiles_h) * P.bh;
  r /= P.tiles_h;
  t.od = (int)(r % P.out_d);
  t.b = (int)(r / P.out_d);
  return t;
}

__global__ void __launch_bounds__(kWgThreads, 1)
    conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                      const __grid_constant__ WgradParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = umma::smem_u32(smem);
  if (base & 1023u) __trap();
  const uint32_t dy0 = base + (uint32_t)P.stages * P.stage_bytes;  // the two dY slots sit behind the X stages
  const uint32_t bar0 = dy0 + 2u * P.dy_bytes;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kWgMaxStages + s); };
  auto bar_dy_full = [&](int s) { return bar0 + 8u * (2 * kWgMaxStages + s); };
  auto bar_dy_empty = [&](int s) { return bar0 + 8u * (2 * kWgMaxStages + 2 + s); };
  const uint32_t bar_acc_full = bar0 + 8u * (2 * kWgMaxStages + 4), bar_acc_empty = bar_acc_full + 8u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (size_t)P.stages * P.stage_bytes + 2 * (size_t)P.dy_bytes +
                                                    8 * (2 * kWgMaxStages + 6));
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
    for (int s = 0; s < P.stages; ++s) {
      umma::mbar_init(bar_full(s), 1);
      umma::mbar_init(bar_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      umma::mbar_init(bar_dy_full(s), 1);
      umma::mbar_init(bar_dy_empty(s), 1);
    }
    umma::mbar_init(bar_acc_full, 1);
    umma::mbar_init(bar_acc_empty, 4);
    umma::mbar_init_fence();
  }
  if (warp == 1) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int cblocks = P.C / 64;
  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      int s = 0, ds = 0;
      uint32_t ph = 0, dph = 0;
      for (int pass = 0; pass < P.passes; ++pass) {
        const int pair0 = pass * P.pairs_per_pass, pair1 = min(P.pairs, pair0 + P.pairs_per_pass);
        for (long long tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
          const WgTile t = wg_tile(P, tile);
          umma::mbar_wait(bar_dy_empty(ds), dph ^ 1u);  // dY: once per tile, shared by every pair of the pass
          mbar_arrive_expect_tx(bar_dy_full(ds), P.dy_bytes);
          for (int n = 0; n < P.nb; ++n)
            tma_load_5d(dy0 + (uint32_t)ds * P.dy_bytes + n * kBox, &map_dy, bar_dy_full(ds), 64 * n, t.ow0, t.oh0, t.od, t.b);
          if (++ds == 2) { ds = 0; dph ^= 1u; }
          for (int pair = pair0; pair < pair1; ++pair) {
            umma::mbar_wait(bar_empty(s), ph ^ 1u);
            mbar_arrive_expect_tx(bar_full(s), P.stage_bytes);
            const uint32_t dst = base + (uint32_t)s * P.stage_bytes;
            for (int h = 0; h < 2; ++h) {
              const int u = min(2 * pair + h, P.units - 1);  // an odd unit count: the last pair loads its unit twice
              const int tap = u / cblocks, cb = u - tap * cblocks;
              tma_load_5d(dst + h * kBox, &map_x, bar_full(s), 64 * cb, t.ow0 * P.stride_hw + P.t1[tap],
                          t.oh0 * P.stride_hw + P.t2[tap], t.od * P.stride_d + P.t3[tap], t.b);
            }
            if (++s == P.stages) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      // kind::f16, bf16 x bf16 -> f32, a_major = b_major = MN, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(P.N >> 3) << 17) |
                             ((128u >> 4) << 24);
      const uint64_t proto = make_desc_mn(0, kBox);
      const uint32_t lo0 = (uint32_t)proto, hi = (uint32_t)(proto >> 32);
      int s = 0, ds = 0;
      uint32_t ph = 0, dph = 0, acc_ph = 0;
      for (int pass = 0; pass < P.passes; ++pass) {
        const int pair0 = pass * P.pairs_per_pass, pair1 = min(P.pairs, pair0 + P.pairs_per_pass);
        umma::mbar_wait(bar_acc_empty, acc_ph ^ 1u);  // the epilogue has drained the previous pass
        umma::fence_after_sync();
        bool first_tile = true;
        for (long long tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
          umma::mbar_wait(bar_dy_full(ds), dph);
          const uint32_t b0 = dy0 + (uint32_t)ds * P.dy_bytes;
          for (int pair = pair0; pair < pair1; ++pair) {
            umma::mbar_wait(bar_full(s), ph);
            umma::fence_after_sync();
            const uint32_t a0 = base + (uint32_t)s * P.stage_bytes;
            const uint32_t d = tmem_base + (uint32_t)((pair - pair0) * P.N);
            const uint32_t a_lo = lo0 + (a0 >> 4), b_lo = lo0 + (b0 >> 4);
#pragma unroll
            for (int j = 0; j < 8; ++j)
              mma_bf16_lo(d, a_lo + 128u * j, b_lo + 128u * j, hi, idesc, (first_tile && j == 0) ? 0u : 1u);
            umma::mma_commit(bar_empty(s));
            if (++s == P.stages) { s = 0; ph ^= 1u; }
          }
          umma::mma_commit(bar_dy_empty(ds));  // every MMA that reads this dY slot has been issued
          if (++ds == 2) { ds = 0; dph ^= 1u; }
          first_tile = false;
        }
        umma::mma_commit(bar_acc_full);
        acc_ph ^= 1u;
      }
    }
  } else {
    // ===== epilogue: once per pass, TMEM -> this CTA's slice of the partial buffer =====
    const int q = warp & 3, m = 32 * q + lane;  // TMEM lane = row of the pair: unit 2p + (m >= 64), input channel m % 64
    float* slice = P.partial + (long long)blockIdx.x * P.slice;
    uint32_t acc_ph = 0;
    for (int pass = 0; pass < P.passes; ++pass) {
      const int pair0 = pass * P.pairs_per_pass, pair1 = min(P.pairs, pair0 + P.pairs_per_pass);
      umma::mbar_wait(bar_acc_full, acc_ph);
      umma::fence_after_sync();
      for (int pair = pair0; pair < pair1; ++pair) {
        const int u = 2 * pair + (m >> 6);
        const bool valid = u < P.units;
        const int uu = valid ? u : 0;
        const int tap = uu / cblocks, cb = uu - tap * cblocks;
        float* dst = slice + ((long long)tap * P.N) * P.C + 64 * cb + (m & 63);
        for (int c0 = 0; c0 < P.N; c0 += 32) {
          float x[32];
          umma::tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)((pair - pair0) * P.N + c0), x);
          if (valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[(long long)(c0 + i) * P.C] = x[i];  // a warp: 32 consecutive input channels
          }
        }
      }
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(bar_acc_empty);
      acc_ph ^= 1u;
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc<512>(tmem_base);
}


// ---- halo mode: 3 x 3 (kh, kw) taps out of one input box -----------------------------------------------------------
// Tile = 16 (h) x 8 (w) output positions; the input box of a (kd, 64-channel block) is [18][10] positions x 64 channels,
// 128 bytes per position (SWIZZLE_128B, written by one TMA load, zero-filled outside the tensor = the convolution's
// padding). The K dimension of the MMAs is positions: an 8-position k group = the 8 w positions of one h row = 8
// consecutive 128-byte rows of the box, the next k group is one BOX ROW further (10 x 128 bytes: the descriptor's stride
// byte offset), and tap (kh, kw) is the same walk started (kh * 10 + kw) positions into the box. The two taps of a pair
// sit in the two 64-row halves of M: the descriptor's leading byte offset is the distance between their starts. Nine taps
// = five pairs (the last one half empty) = 320 accumulator columns; one pass over the tiles per box.
constexpr int kHaloBw = 8, kHaloBh = 16, kHaloRow = (kHaloBw + 2) * 128;       // bytes of one box row
constexpr uint32_t kHaloBoxBytes = (kHaloBw + 2) * (kHaloBh + 2) * 128;          // 23 040 delivered by the TMA
constexpr uint32_t kHaloStage = (kHaloBoxBytes + 1023u) & ~1023u;                // 23 552: stages stay 1 KB aligned

__global__ void __launch_bounds__(kWgThreads, 1)
    conv_wgrad_halo_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_dy,
                           const __grid_constant__ WgradParams P) {
  extern __shared__ __align__(1024) unsigned char smem[];
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = umma::smem_u32(smem);
  if (base & 1023u) __trap();
  const uint32_t dy0 = base + (uint32_t)P.stages * P.stage_bytes;
  const uint32_t bar0 = dy0 + 2u * P.dy_bytes;
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_empty = [&](int s) { return bar0 + 8u * (kWgMaxStages + s); };
  auto bar_dy_full = [&](int s) { return bar0 + 8u * (2 * kWgMaxStages + s); };
  auto bar_dy_empty = [&](int s) { return bar0 + 8u * (2 * kWgMaxStages + 2 + s); };
  const uint32_t bar_acc_full = bar0 + 8u * (2 * kWgMaxStages + 4), bar_acc_empty = bar_acc_full + 8u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (size_t)P.stages * P.stage_bytes + 2 * (size_t)P.dy_bytes +
                                                    8 * (2 * kWgMaxStages + 6));
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_dy) : "memory");
    for (int s = 0; s < P.stages; ++s) {
      umma::mbar_init(bar_full(s), 1);
      umma::mbar_init(bar_empty(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      umma::mbar_init(bar_dy_full(s), 1);
      umma::mbar_init(bar_dy_empty(s), 1);
    }
    umma::mbar_init(bar_acc_full, 1);
    umma::mbar_init(bar_acc_empty, 4);
    umma::mbar_init_fence();
  }
  if (warp == 1) umma::tmem_alloc<512>(tmem_slot);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  const int cblocks = P.C / 64;
  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer: per box (= pass), per tile: the dY tile and ONE halo box =====
      int s = 0, ds = 0;
      uint32_t ph = 0, dph = 0;
      for (int box = 0; box < P.boxes; ++box) {
        const int kd = box / cblocks, cb = box - kd * cblocks;
        for (long long tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
          const WgTile t = wg_tile(P, tile);
          umma::mbar_wait(bar_dy_empty(ds), dph ^ 1u);
          mbar_arrive_expect_tx(bar_dy_full(ds), P.dy_bytes);
          for (int n = 0; n < P.nb; ++n)
            tma_load_5d(dy0 + (uint32_t)ds * P.dy_bytes + n * kBox, &map_dy, bar_dy_full(ds), 64 * n, t.ow0, t.oh0, t.od, t.b);
          if (++ds == 2) { ds = 0; dph ^= 1u; }
          umma::mbar_wait(bar_empty(s), ph ^ 1u);
          mbar_arrive_expect_tx(bar_full(s), kHaloBoxBytes);
          tma_load_5d(base + (uint32_t)s * P.stage_bytes, &map_x, bar_full(s), 64 * cb, t.ow0 - P.pad_w, t.oh0 - P.pad_h,
                      t.od * P.stride_d + kd - P.pad_d, t.b);
          if (++s == P.stages) { s = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(P.N >> 3) << 17) |
                             ((128u >> 4) << 24);
      int s = 0, ds = 0;
      uint32_t ph = 0, dph = 0, acc_ph = 0;
      for (int box = 0; box < P.boxes; ++box) {
        umma::mbar_wait(bar_acc_empty, acc_ph ^ 1u);  // the epilogue has drained the previous box's accumulators
        umma::fence_after_sync();
        bool first_tile = true;
        for (long long tile = blockIdx.x; tile < P.total_tiles; tile += gridDim.x) {
          umma::mbar_wait(bar_dy_full(ds), dph);
          umma::mbar_wait(bar_full(s), ph);
          umma::fence_after_sync();
          const uint32_t a0 = base + (uint32_t)s * P.stage_bytes, b0 = dy0 + (uint32_t)ds * P.dy_bytes;
          const uint64_t bdesc = make_desc_mn2(b0, kBox, 1024);
#pragma unroll
          for (int p = 0; p < 5; ++p) {
            const int ta = 2 * p, tb = p < 4 ? 2 * p + 1 : 8;
            const uint32_t offa = (uint32_t)((ta / 3) * (kHaloBw + 2) + ta % 3) * 128u;
            const uint32_t offb = (uint32_t)((tb / 3) * (kHaloBw + 2) + tb % 3) * 128u;
            const uint32_t lbo = p < 4 ? offb - offa : 128u;  // (the last pair's upper half is not read back)
            const uint64_t adesc = make_desc_mn2(a0 + offa, lbo, kHaloRow);
            const uint32_t d = tmem_base + (uint32_t)(p * P.N);
#pragma unroll
            for (int j = 0; j < 8; ++j)  // K = 16 positions = two box rows of 8
              mma_bf16(d, adesc + (uint64_t)((2u * kHaloRow * j) >> 4), bdesc + (uint64_t)(128u * j), idesc,
                       (first_tile && j == 0) ? 0u : 1u);
          }
          umma::mma_commit(bar_empty(s));
          if (++s == P.stages) { s = 0; ph ^= 1u; }
          umma::mma_commit(bar_dy_empty(ds));
          if (++ds == 2) { ds = 0; dph ^= 1u; }
          first_tile = false;
        }
        umma::mma_commit(bar_acc_full);
        acc_ph ^= 1u;
      }
    }
  } else {
    // ===== epilogue: once per box, TMEM -> this CTA's slice of the partial buffer =====
    const int q = warp & 3, m = 32 * q + lane;  // TMEM lane: tap 2p + (m >= 64) of the box, input channel m % 64
    float* slice = P.partial + (long long)blockIdx.x * P.slice;
    uint32_t acc_ph = 0;
    for (int box = 0; box < P.boxes; ++box) {
      const int kd = box / cblocks, cb = box - kd * cblocks;
      umma::mbar_wait(bar_acc_full, acc_ph);
      umma::fence_after_sync();
      for (int p = 0; p < 5; ++p) {
        const int tin = 2 * p + (m >> 6);
        const bool valid = tin < 9;
        const int tap = kd * 9 + (valid ? tin : 0);
        float* dst = slice + ((long long)tap * P.N) * P.C + 64 * cb + (m & 63);
        for (int c0 = 0; c0 < P.N; c0 += 32) {
          float x[32];
          umma::tmem_ld_32x32(tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(p * P.N + c0), x);
          if (valid) {
#pragma unroll
            for (int i = 0; i < 32; ++i) dst[(long long)(c0 + i) * P.C] = x[i];
          }
        }
      }
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) umma::mbar_arrive(bar_acc_empty);
      acc_ph ^= 1u;
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_dealloc<512>(tmem_base);
}

// dW[i] = sum over the CTAs' slices, in slice order
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int slices, long long n,
                                                           float* __restrict__ dw) {
  pdl_launch_dependents();
  pdl_wait();
  const long long i4 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i4 * 4 >= n) return;
  // four running sums (slices g = 0, 1, 2, 3 mod 4) so that four loads are in flight; combined in a fixed order
  float4 a4[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) a4[u] = make_float4(0.f, 0.f, 0.f, 0.f);
  int g = 0;
  for (; g + 3 < slices; g += 4) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(partial + (long long)(g + u) * n) + i4);
      a4[u].x += v.x; a4[u].y += v.y; a4[u].z += v.z; a4[u].w += v.w;
    }
  }
#pragma unroll
  for (int u = 0; u < 3; ++u)
    if (g + u < slices) {
      const float4 v = __ldcg(reinterpret_cast<const float4*>(partial + (long long)(g + u) * n) + i4);
      a4[u].x += v.x; a4[u].y += v.y; a4[u].z += v.z; a4[u].w += v.w;
    }
  float4 acc;
  acc.x = (a4[0].x + a4[1].x) + (a4[2].x + a4[3].x);
  acc.y = (a4[0].y + a4[1].y) + (a4[2].y + a4[3].y);
  acc.z = (a4[0].z + a4[1].z) + (a4[2].z + a4[3].z);
  acc.w = (a4[0].w + a4[1].w) + (a4[2].w + a4[3].w);
  reinterpret_cast<float4*>(dw)[i4] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn wg_encode_fn() {
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

thread_local char g_wg_error[384] = "";
int wg_fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_wg_error, sizeof(g_wg_error), fmt, ap);
  va_end(ap);
  return status;
}

}  // namespace
}  // namespace lisec

using namespace lisec;

struct lisec_wgrad_plan {
  CUtensorMap map_x, map_dy;
  WgradParams p;
  int grid, smem;
  float* dw;
};

extern "C" {

const char* lisec_wgrad_last_error(void) { return g_wg_error; }

int64_t lisec_conv_wgrad_workspace_bytes(const lisec_conv_desc* d) {
  if (!d) return -1;
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (int64_t)sms * d->kd * d->kh * d->kw * d->out_c * d->in_c * 4;
}

int32_t lisec_conv_wgrad_plan_create(const lisec_conv_desc* d, const void* x, const void* dy, float* workspace, float* dw,
                                     lisec_wgrad_plan** out) {
  if (!d || !x || !dy || !workspace || !dw || !out) return wg_fail(LISEC_ERR_BAD_ARG, "null argument");
  *out = nullptr;
  const int C = d->in_c, N = d->out_c, taps = d->kd * d->kh * d->kw;
  const int shw = d->stride_hw;
  if (shw != 1 && shw != 2) return wg_fail(LISEC_ERR_UNSUPPORTED, "wgrad: stride_hw 1 or 2");
  if (C % 64 || N % 64 || N > 256 || C > 1024)
    return wg_fail(LISEC_ERR_BAD_CONFIG, "wgrad: in_c, out_c multiples of 64; out_c <= 256, in_c <= 1024");
  if (taps > 27 || taps < 1) return wg_fail(LISEC_ERR_BAD_CONFIG, "wgrad: at most 27 taps");
  if (d->tile_w * d->tile_h != 128 || d->tile_w < 8 || (d->tile_w & (d->tile_w - 1)))
    return wg_fail(LISEC_ERR_BAD_CONFIG, "wgrad: tile_w * tile_h = 128, tile_w a power of two >= 8");
  if (d->n_tiles != 1 || d->shuffle > 1) return wg_fail(LISEC_ERR_UNSUPPORTED, "wgrad: plain convolutions only");
  const int OD = (d->in_d + 2 * d->pad_d - d->kd) / d->stride_d + 1;
  const int OH = (d->in_h + 2 * d->pad_h - d->kh) / shw + 1, OW = (d->in_w + 2 * d->pad_w - d->kw) / shw + 1;
  if (OD < 1 || OH < 1 || OW < 1) return wg_fail(LISEC_ERR_BAD_CONFIG, "empty output");
  EncodeTiledFn encode = wg_encode_fn();
  if (!encode) return wg_fail(LISEC_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  lisec_wgrad_plan* pl = new lisec_wgrad_plan();
  memset(pl, 0, sizeof(*pl));
  WgradParams& p = pl->p;
  p.bw = d->tile_w;
  p.bh = d->tile_h;
  // halo mode (conv_wgrad_halo_kernel): the 3 x 3 (kh, kw), stride-1, 64-output-channel layers = the Conv3D blocks
  {
    const char* e = getenv("LISEC_WGRAD_HALO");
    p.halo = (shw == 1 && d->kh == 3 && d->kw == 3 && d->pad_h == 1 && d->pad_w == 1 && N == 64 && 5 * N <= 512 &&
              !(e && e[0] == '0')) ? 1 : 0;
  }
  if (p.halo) {
    p.bw = kHaloBw;
    p.bh = kHaloBh;
  }
  p.tiles_w = (OW + p.bw - 1) / p.bw;
  p.tiles_h = (OH + p.bh - 1) / p.bh;
  p.out_d = OD;
  p.batch = d->batch;
  p.total_tiles = (long long)p.tiles_w * p.tiles_h * OD * d->batch;
  p.C = C;
  p.N = N;
  p.taps = taps;
  p.nb = N / 64;
  p.units = taps * (C / 64);
  p.pairs = (p.units + 1) / 2;
  p.pairs_per_pass = 512 / N;
  p.passes = (p.pairs + p.pairs_per_pass - 1) / p.pairs_per_pass;
  p.stage_bytes = p.halo ? kHaloStage : 2u * kBox;
  p.dy_bytes = (uint32_t)p.nb * kBox;
  p.boxes = d->kd * (C / 64);
  p.kd_n = d->kd;
  p.pad_w = d->pad_w;
  p.pad_h = d->pad_h;
  p.pad_d = d->pad_d;
  p.stages = (int)((200u * 1024u - 2u * p.dy_bytes) / p.stage_bytes);
  if (p.stages > kWgMaxStages) p.stages = kWgMaxStages;
  if (p.stages < 2) {
    delete pl;
    return wg_fail(LISEC_ERR_BAD_CONFIG, "wgrad: no room for two stages");
  }
  p.stride_d = d->stride_d;
  p.stride_hw = shw;
  int t = 0;
  for (int kd = 0; kd < d->kd; ++kd)
    for (int kh = 0; kh < d->kh; ++kh)
      for (int kw = 0; kw < d->kw; ++kw, ++t) {
        p.t1[t] = (signed char)(kw - d->pad_w);
        p.t2[t] = (signed char)(kh - d->pad_h);
        p.t3[t] = (signed char)(kd - d->pad_d);
      }
  p.partial = workspace;
  p.slice = (long long)taps * N * C;
  pl->dw = dw;
  pl->smem = p.stages * (int)p.stage_bytes + 2 * (int)p.dy_bytes + 8 * (2 * kWgMaxStages + 6) + 16;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  {
    const cuuint64_t W = d->in_w, H = d->in_h, D = d->in_d, B = d->batch;
    cuuint64_t dims[5] = {(cuuint64_t)C, W, H, D, B};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, W * C * 2, H * W * C * 2, D * H * W * C * 2};
    // stride_hw = 2: the TMA walks W and H with an element stride of 2 — the box spans 2 * tile positions of the input and
    // delivers the tile's 128 positions
    cuuint32_t box[5] = {64, (cuuint32_t)(p.bw * shw), (cuuint32_t)(p.bh * shw), 1, 1};
    if (p.halo) {  // the tile plus one position on every side in W and H
      box[1] = (cuuint32_t)(p.bw + 2);
      box[2] = (cuuint32_t)(p.bh + 2);
    }
    cuuint32_t xstr[5] = {1, (cuuint32_t)shw, (cuuint32_t)shw, 1, 1};
    CUresult r = encode(&pl->map_x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(x), dims, strides, box, xstr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete pl;
      return wg_fail(LISEC_ERR_CUDA, "cuTensorMapEncodeTiled(x) failed: CUresult %d", (int)r);
    }
  }
  {
    const cuuint64_t W = OW, H = OH, D = OD, B = d->batch;
    cuuint64_t dims[5] = {(cuuint64_t)N, W, H, D, B};
    cuuint64_t strides[4] = {(cuuint64_t)N * 2, W * N * 2, H * W * N * 2, D * H * W * N * 2};
    cuuint32_t box[5] = {64, (cuuint32_t)p.bw, (cuuint32_t)p.bh, 1, 1};
    CUresult r = encode(&pl->map_dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(dy), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      delete pl;
      return wg_fail(LISEC_ERR_CUDA, "cuTensorMapEncodeTiled(dy) failed: CUresult %d", (int)r);
    }
  }
  int dev = 0, sms = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_wgrad_halo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448);
  if (e != cudaSuccess) {
    delete pl;
    return wg_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  }
  pl->grid = (int)(p.total_tiles < sms ? p.total_tiles : sms);
  *out = pl;
  return LISEC_OK;
}

int32_t lisec_conv_wgrad_plan_run(lisec_wgrad_plan* pl, void* stream) {
  if (!pl) return wg_fail(LISEC_ERR_BAD_ARG, "null plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  cudaError_t e = pl->p.halo ? launch_pdl(conv_wgrad_halo_kernel, pl->grid, kWgThreads, (size_t)pl->smem, st, pl->map_x,
                                          pl->map_dy, pl->p)
                             : launch_pdl(conv_wgrad_kernel, pl->grid, kWgThreads, (size_t)pl->smem, st, pl->map_x, pl->map_dy,
                                          pl->p);
  if (e == cudaSuccess) {
    const long long n = pl->p.slice;
    e = launch_pdl(wgrad_reduce_kernel, dim3((unsigned)((n / 4 + 255) / 256)), dim3(256), 0, st,
                   (const float*)pl->p.partial, pl->grid, n, pl->dw);
  }
  if (e != cudaSuccess) return wg_fail(LISEC_ERR_CUDA, "%s", cudaGetErrorString(e));
  return LISEC_OK;
}

void lisec_conv_wgrad_plan_destroy(lisec_wgrad_plan* pl) { delete pl; }

}  // extern "C"

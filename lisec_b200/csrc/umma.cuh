// tcgen05 / TMEM / mbarrier helpers for sm_100a, as used by the VFE kernel's dense_2 stage (vfe.cu) and probed in
// isolation by tools/umma_probe.cu. Everything here is inline PTX; no library code.
//
// Operand layout used by the product: "K-major, 128-byte swizzle" — the layout every dense sm_100 GEMM uses. For a tf32
// operand Op[mn][k] (mn = the M index of A or the N index of B) one k-block of 32 channels is a slab of 128-byte rows:
//     byte(mn, k) = (k / 32) * slab_bytes + mn * 128 + ((((k % 32) / 4) ^ (mn % 8)) * 16) + (k % 4) * 4
// i.e. row mn holds its 32 k-values in eight 16-byte chunks XOR-swizzled with mn % 8 (Swizzle<3,4,3> on the byte
// address; slabs are 1 KB-aligned). One MMA has K = 8: k-step j inside a slab is the same descriptor advanced by 32
// bytes. Measured with tools/umma_probe.cu (M=64, N=256, K=64): exact placement, 3xTF32 error 1.1e-6 of rms.
// An MN-major tf32 operand needs the separate SWIZZLE_128B_BASE32B layout (plain SWIZZLE_128B MN-major yields zeros for
// 32-bit types; probe variants 0/1, whose helpers live in the probe) — not used here. For bf16 the plain MN-major
// SWIZZLE_128B layout works as the canonical form says (tools/umma_mn_probe.cu, exact on B200): a TMA box [k rows][64
// channels] IS an MN-major operand — instruction descriptor bits 15/16 (a_major/b_major) = 1, stride byte offset 1024
// (8 k-rows), leading byte offset = distance between 64-channel boxes, one K = 16 step = +2048 bytes, and the K rows may
// start at any row of the box (row shifts 1, 3, 19 exact: the swizzle follows the absolute address). This is what the
// weight-gradient GEMM of the training step will read.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace lisec {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (mn, k) of a K-major SW128 operand whose 32-channel slabs are slab_bytes apart
__device__ __host__ __forceinline__ uint32_t kmajor_offset(int mn, int k, uint32_t slab_bytes) {
  return (uint32_t)(k >> 5) * slab_bytes + (uint32_t)mn * 128u + (uint32_t)((((k & 31) >> 2) ^ (mn & 7)) << 4) +
         (uint32_t)(k & 3) * 4u;
}
// descriptor of a K-major SW128 operand: 8-row groups are 1 KB apart (stride byte offset); the leading offset is unused
__device__ __forceinline__ uint64_t make_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::tf32, float32 accumulation, both operands K-major.
__device__ __host__ constexpr uint32_t make_idesc_tf32_k(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void mma_tf32_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// All MMAs issued so far by this thread -> one arrival on the mbarrier when they have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint32_t mbar_smem) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar_smem)
               : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads, TMA)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM allocation: one full warp, result (base address) lands in shared memory --------------------------
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS) : "memory");
}

// ---- TMEM -> registers: the warp's 32 lanes (lane quadrant = warp id % 4) x 32 consecutive columns ---------
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// Two 8-column loads of the warp's 32 lanes (columns c0.. and c1..), one wait. The wait names every destination
// register as an in/out operand, so no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld_2x8(uint32_t taddr0, uint32_t taddr1, float (&a)[8], float (&b)[8]) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr0)
               : "memory");
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr1)
               : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = __uint_as_float(r[i]);
    b[i] = __uint_as_float(r[8 + i]);
  }
}

// ---- mbarrier ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t mbar_smem, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar_smem), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t mbar_smem) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(mbar_smem) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes or the hint (ns) runs out,
// instead of burning issue slots in a polling loop (ncu: 10 % of the VFE kernel's instructions were such polls)
__device__ __forceinline__ bool mbar_try_wait(uint32_t mbar_smem, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(mbar_smem), "r"(parity), "r"(1000000u)
      : "memory");
  return ok != 0;
}
// non-blocking test (the tensor thread polls several barriers)
__device__ __forceinline__ bool mbar_test_wait(uint32_t mbar_smem, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(mbar_smem), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded: a protocol bug must surface as a CUDA error (trap), never as a hung GPU. try_wait sleeps in hardware for up
// to ~1 ms per attempt, so the cap is minutes of wall clock — far beyond any legitimate wait in these kernels.
__device__ __forceinline__ void mbar_wait(uint32_t mbar_smem, uint32_t parity) {
  for (unsigned spins = 0; !mbar_try_wait(mbar_smem, parity); ++spins)
    if (spins > (1u << 22)) __trap();
}

// ---- 3xTF32 operand split: x = hi + lo with hi = rn_tf32(x), lo = rn_tf32(x - hi) --------------------------
// (the tensor core reads only the top 19 bits of each 32-bit container; rounding here instead of letting it truncate
// halves the representation error: |x - hi - lo| <= 2^-23 |x|)
// cvt.rna.tf32.f32 (round to nearest, ties away) on a FINITE value as two integer instructions; the PTX conversion
// itself compiles to four (it also passes inf / nan through), and the VFE kernel does ~40 of these per tile row.
__device__ __forceinline__ float tf32_rn(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void tf32_split(float x, float& hi, float& lo) {
  hi = tf32_rn(x);
  lo = tf32_rn(x - hi);
}

}  // namespace umma
}  // namespace lisec

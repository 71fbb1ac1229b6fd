// Operand pack of the planned fused forward (generations 6 and 7: fused_fwd6.cu, fused_fwd7.cu) and the tf32 TS-form MMA.
#pragma once
#include "fused_common.cuh"

namespace imp {

struct FusedPack6 {  // one (tower, step)
  static constexpr int WC_BYTES = FZ_D * (FZ_D * FZ_K) * 2;   // Wc in two K halves, as FusedPack (16 KiB)
  static constexpr int BZRH_BYTES = 2 * FZ_D * FZ_D * 2;      // [Wr_h | Wz_h]^T  f16   [64 x 32]
  static constexpr int BZRA_BYTES = 2 * FZ_D * FZ_D * 4;      // [Wr_a | Wz_a]^T  tf32  [64 x 32]
  static constexpr int BHH_BYTES = FZ_D * FZ_D * 2;           // Wh_h^T           f16   [32 x 32]
  static constexpr int BHA_BYTES = FZ_D * FZ_D * 4;           // Wh_a^T           tf32  [32 x 32]
  static constexpr int OFF_BZRH = WC_BYTES;
  static constexpr int OFF_BZRA = OFF_BZRH + BZRH_BYTES;
  static constexpr int OFF_BHH = OFF_BZRA + BZRA_BYTES;
  static constexpr int OFF_BHA = OFF_BHH + BHH_BYTES;
  static constexpr int OFF_BIAS = OFF_BHA + BHA_BYTES;        // gamma[32], beta[32] (fp32)
  static constexpr int OFF_BBZR = OFF_BIAS + 2 * FZ_D * 4;    // [64 x 16] f16, column 0 = 0.5 (br | bz)
  static constexpr int OFF_BBH = OFF_BBZR + 2 * FZ_D * 16 * 2;  // [32 x 16] f16, column 0 = bh
  static constexpr int BYTES = OFF_BBH + FZ_D * 16 * 2;
};
static_assert(FusedPack6::BYTES % 128 == 0 && FusedPack6::OFF_BIAS % 16 == 0, "pack alignment");

__device__ __forceinline__ uint32_t f32_to_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return r;
}

// A operand of a kind::tf32 MMA from tensor memory: row i in lane i, 8 consecutive 32-bit columns = K = 8.
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

}  // namespace imp

// Split-precision tcgen05 GEMMs: an fp32 operand x is carried as TWO 16-bit planes  x = hi + lo  (hi = rn16(x), lo = rn16(x - hi)),
// and a product is three MMAs into the same fp32 TMEM accumulator:  A B^T ~= A_hi B_hi^T + A_lo B_hi^T + A_hi B_lo^T.
// The dropped lo x lo term and the rounding of lo are both 2^-2p relative (p = 8 for bf16 planes, 11 for fp16 planes), so the
// contraction is good to ~2^-16 (bf16) / ~2^-21 (fp16) per product instead of 2^-9 -- what the 0.05 mm joint bar needs
// (DESIGN.md section 2).  An operand that is exactly representable in 16 bits (a bf16 feature map) has no lo plane and its lo
// MMAs are skipped.  The instruction descriptor has separate A / B format fields, but an fp16 x bf16 pairing traps as an illegal
// instruction on sm_100a (measured), so both operands of an MMA use ONE format: where a bf16 feature map is an operand, the other
// side is carried as THREE bf16 planes (24 bits, split3_bf16) instead of two fp16 planes.
//
// A may also come from TENSOR MEMORY (tcgen05.mma [d], [a_tmem], b_desc): row m of A = TMEM lane m, K runs along the columns,
// two 16-bit elements per 32-bit column (low half = even k).  Weights that stay resident for a whole kernel live there, which
// frees their shared memory for the activation planes and removes the A-side shared-memory read from every MMA.
#pragma once
#include "tmem_ldst.cuh"
#include <cuda_fp16.h>

namespace kpf {

constexpr int FMT_F16 = 0, FMT_BF16 = 1;   // = the instruction descriptor's a_format / b_format encodings for kind::f16

// 32-bit instruction descriptor for kind::f16, fp32 accumulate, element formats per operand
__host__ __device__ constexpr uint32_t umma_idesc_f16(int M, int N, bool a_mn_major, bool b_mn_major, int a_fmt, int b_fmt) {
    return (1u << 4) /*D=f32*/ | ((uint32_t)a_fmt << 7) | ((uint32_t)b_fmt << 10) | ((a_mn_major ? 1u : 0u) << 15) |
           ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// A from tensor memory, B from a shared-memory descriptor
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// One operand plane pair in shared memory.  lo == 0: the operand is exact in 16 bits (no lo plane).
struct SmemOp {
    uint32_t hi, lo;      // shared-memory byte addresses of the two planes
    uint32_t lbo, sbo;    // descriptor strides (umma.cuh)
};
// A operand planes in tensor memory: 16-bit elements, K/2 columns per plane
struct TmemOp {
    uint32_t hi, lo;      // TMEM addresses (lane 0 of the operand, first column); lo == 0xffffffff: no lo plane
};
constexpr uint32_t NO_PLANE = 0xffffffffu;

// ---- MMA issue with the 64-bit shared-memory descriptor kept as two 32-bit halves.  Only the low half (start address | LBO) moves
// along K, by a constant that folds into an immediate once the K loop is unrolled; the 64-bit add-with-carry, the per-iteration
// R2UR of the accumulator address / instruction descriptor and the loop control of the rolled form cost ~16 uniform-datapath
// instructions per MMA (~42 cycles per MMA measured, against a 16-cycle tensor-pipe floor at N = 32).
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_bytes) {
    return ((smem_addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16);
}
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }

__device__ __forceinline__ void umma_issue_ss(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_issue_ts(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b64 db;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// one plane pair over KS = K / 16 steps, fully unrolled
template <int KS>
__device__ __forceinline__ void umma_planes_ss(uint32_t tmem_d, uint32_t a_addr, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_addr,
                                               uint32_t b_lbo, uint32_t b_sbo, uint32_t idesc, uint32_t& acc) {
    const uint32_t al = desc_lo(a_addr, a_lbo), ah = desc_hi(a_sbo), bl = desc_lo(b_addr, b_lbo), bh = desc_hi(b_sbo);
    const uint32_t a_step = (2 * a_lbo) >> 4, b_step = (2 * b_lbo) >> 4;   // start-address field, 16-byte units
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        umma_issue_ss(tmem_d, al + ks * a_step, ah, bl + ks * b_step, bh, idesc, acc);
        acc = 1u;
    }
}
template <int KS>
__device__ __forceinline__ void umma_planes_ts(uint32_t tmem_d, uint32_t a_tmem, uint32_t b_addr, uint32_t b_lbo, uint32_t b_sbo,
                                               uint32_t idesc, uint32_t& acc) {
    const uint32_t bl = desc_lo(b_addr, b_lbo), bh = desc_hi(b_sbo);
    const uint32_t b_step = (2 * b_lbo) >> 4;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        umma_issue_ts(tmem_d, a_tmem + 8 * ks, bl + ks * b_step, bh, idesc, acc);
        acc = 1u;
    }
}

// D[128 x N] (+)= A B^T over K = 16 KS, split precision, both operands from shared memory.  Issued by ONE thread.
// Order: the two small cross terms first, then hi x hi.
template <int KS>
__device__ __forceinline__ void umma_gemm3_ss_t(uint32_t tmem_d, const SmemOp a, const SmemOp b, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u;
    if (a.lo) umma_planes_ss<KS>(tmem_d, a.lo, a.lbo, a.sbo, b.hi, b.lbo, b.sbo, idesc, acc);
    if (b.lo) umma_planes_ss<KS>(tmem_d, a.hi, a.lbo, a.sbo, b.lo, b.lbo, b.sbo, idesc, acc);
    umma_planes_ss<KS>(tmem_d, a.hi, a.lbo, a.sbo, b.hi, b.lbo, b.sbo, idesc, acc);
}
// Same with A in tensor memory (8 columns per K = 16 step).
template <int KS>
__device__ __forceinline__ void umma_gemm3_ts_t(uint32_t tmem_d, const TmemOp a, const SmemOp b, uint32_t idesc, bool accumulate) {
    uint32_t acc = accumulate ? 1u : 0u;
    if (a.lo != NO_PLANE) umma_planes_ts<KS>(tmem_d, a.lo, b.hi, b.lbo, b.sbo, idesc, acc);
    if (b.lo) umma_planes_ts<KS>(tmem_d, a.hi, b.lo, b.lbo, b.sbo, idesc, acc);
    umma_planes_ts<KS>(tmem_d, a.hi, b.hi, b.lbo, b.sbo, idesc, acc);
}
// run-time K front ends (K in {16, 32, 64, 128, 256})
__device__ __forceinline__ void umma_gemm3_ss(uint32_t tmem_d, const SmemOp a, const SmemOp b, uint32_t idesc, int K, bool accumulate) {
    switch (K) {
        case 16: umma_gemm3_ss_t<1>(tmem_d, a, b, idesc, accumulate); break;
        case 32: umma_gemm3_ss_t<2>(tmem_d, a, b, idesc, accumulate); break;
        case 64: umma_gemm3_ss_t<4>(tmem_d, a, b, idesc, accumulate); break;
        case 128: umma_gemm3_ss_t<8>(tmem_d, a, b, idesc, accumulate); break;
        default: umma_gemm3_ss_t<16>(tmem_d, a, b, idesc, accumulate); break;
    }
}
__device__ __forceinline__ void umma_gemm3_ts(uint32_t tmem_d, const TmemOp a, const SmemOp b, uint32_t idesc, int K, bool accumulate) {
    switch (K) {
        case 16: umma_gemm3_ts_t<1>(tmem_d, a, b, idesc, accumulate); break;
        case 32: umma_gemm3_ts_t<2>(tmem_d, a, b, idesc, accumulate); break;
        case 64: umma_gemm3_ts_t<4>(tmem_d, a, b, idesc, accumulate); break;
        case 128: umma_gemm3_ts_t<8>(tmem_d, a, b, idesc, accumulate); break;
        default: umma_gemm3_ts_t<16>(tmem_d, a, b, idesc, accumulate); break;
    }
}

// ---- fp32 -> (hi, lo) 16-bit planes ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack2_f16(float a, float b) {
    const __half2 v = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}
// two floats -> one 32-bit word of each plane (low half = first element)
template <int FMT>
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
    if constexpr (FMT == FMT_BF16) {
        const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        const float2 hf = __bfloat1622float2(h);
        hi = *reinterpret_cast<const uint32_t*>(&h);
        lo = pack2_bf16(a - hf.x, b - hf.y);
    } else {
        const __half2 h = __floats2half2_rn(a, b);
        const float2 hf = __half22float2(h);
        hi = *reinterpret_cast<const uint32_t*>(&h);
        lo = pack2_f16(a - hf.x, b - hf.y);
    }
}
__device__ __forceinline__ void split2(int fmt, float a, float b, uint32_t& hi, uint32_t& lo) {
    if (fmt == FMT_BF16) split2<FMT_BF16>(a, b, hi, lo);
    else split2<FMT_F16>(a, b, hi, lo);
}
// eight floats -> one 16-byte operand chunk of each plane
__device__ __forceinline__ void split8(int fmt, const float* v, uint4& hi, uint4& lo) {
    split2(fmt, v[0], v[1], hi.x, lo.x);
    split2(fmt, v[2], v[3], hi.y, lo.y);
    split2(fmt, v[4], v[5], hi.z, lo.z);
    split2(fmt, v[6], v[7], hi.w, lo.w);
}
// fp32 -> three bf16 planes (hi + mid + lo carries 24 mantissa bits, full fp32 range): the partner of a bf16 feature-map operand
__device__ __forceinline__ void split3_bf16(float a, float b, uint32_t& hi, uint32_t& mid, uint32_t& lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    const float2 hf = __bfloat1622float2(h);
    const float ra = a - hf.x, rb = b - hf.y;
    const __nv_bfloat162 m = __floats2bfloat162_rn(ra, rb);
    const float2 mf = __bfloat1622float2(m);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    mid = *reinterpret_cast<const uint32_t*>(&m);
    lo = pack2_bf16(ra - mf.x, rb - mf.y);
}
__device__ __forceinline__ void split8x3_bf16(const float* v, uint4& hi, uint4& mid, uint4& lo) {
    split3_bf16(v[0], v[1], hi.x, mid.x, lo.x);
    split3_bf16(v[2], v[3], hi.y, mid.y, lo.y);
    split3_bf16(v[4], v[5], hi.z, mid.z, lo.z);
    split3_bf16(v[6], v[7], hi.w, mid.w, lo.w);
}
// D (+)= A B^T with A = a bf16 feature-map operand (one exact plane, or hi + lo of an fp32 map) and B = three bf16 planes at
// b_addr + {0, 1, 2} * b_plane_bytes: A_hi (B_hi + B_mid + B_lo) [+ A_lo (B_hi + B_mid)], smallest terms first.
__device__ __forceinline__ void umma_gemm_map_x3(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo /* 0: none */, uint32_t a_lbo, uint32_t a_sbo,
                                                 uint32_t b_addr, uint32_t b_plane_bytes, uint32_t b_lbo, uint32_t b_sbo, uint32_t idesc, int K,
                                                 bool accumulate) {
    umma_gemm(tmem_d, a_hi, a_lbo, a_sbo, b_addr + 2 * b_plane_bytes, b_lbo, b_sbo, idesc, K, accumulate);
    if (a_lo) umma_gemm(tmem_d, a_lo, a_lbo, a_sbo, b_addr + b_plane_bytes, b_lbo, b_sbo, idesc, K, true);
    umma_gemm(tmem_d, a_hi, a_lbo, a_sbo, b_addr + b_plane_bytes, b_lbo, b_sbo, idesc, K, true);
    if (a_lo) umma_gemm(tmem_d, a_lo, a_lbo, a_sbo, b_addr, b_lbo, b_sbo, idesc, K, true);
    umma_gemm(tmem_d, a_hi, a_lbo, a_sbo, b_addr, b_lbo, b_sbo, idesc, K, true);
}

// the value the MMAs see for x (hi + lo), for callers that must stay consistent with an operand they wrote
__device__ __forceinline__ float split_value(int fmt, float x) {
    if (fmt == FMT_BF16) {
        const float h = __bfloat162float(__float2bfloat16_rn(x));
        return h + __bfloat162float(__float2bfloat16_rn(x - h));
    }
    const float h = __half2float(__float2half_rn(x));
    return h + __half2float(__float2half_rn(x - h));
}

}  // namespace kpf

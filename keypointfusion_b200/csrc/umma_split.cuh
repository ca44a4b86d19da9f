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

// D[128 x N] (+)= A B^T over K (multiple of 16), split precision, both operands from shared memory.  Issued by ONE thread.
// Order: the two small cross terms first, then hi x hi.
__device__ __forceinline__ void umma_gemm3_ss(uint32_t tmem_d, const SmemOp a, const SmemOp b, uint32_t idesc, int K, bool accumulate) {
    const uint64_t a_step = (uint64_t)((2 * a.lbo) >> 4), b_step = (uint64_t)((2 * b.lbo) >> 4);
    bool acc = accumulate;
    if (a.lo) {
        uint64_t ad = umma_smem_desc(a.lo, a.lbo, a.sbo), bd = umma_smem_desc(b.hi, b.lbo, b.sbo);
#pragma unroll 1
        for (int ks = 0; ks < K / 16; ++ks) {
            umma_bf16(tmem_d, ad, bd, idesc, acc);
            acc = true;
            ad += a_step;
            bd += b_step;
        }
    }
    if (b.lo) {
        uint64_t ad = umma_smem_desc(a.hi, a.lbo, a.sbo), bd = umma_smem_desc(b.lo, b.lbo, b.sbo);
#pragma unroll 1
        for (int ks = 0; ks < K / 16; ++ks) {
            umma_bf16(tmem_d, ad, bd, idesc, acc);
            acc = true;
            ad += a_step;
            bd += b_step;
        }
    }
    {
        uint64_t ad = umma_smem_desc(a.hi, a.lbo, a.sbo), bd = umma_smem_desc(b.hi, b.lbo, b.sbo);
#pragma unroll 1
        for (int ks = 0; ks < K / 16; ++ks) {
            umma_bf16(tmem_d, ad, bd, idesc, acc);
            acc = true;
            ad += a_step;
            bd += b_step;
        }
    }
}

// Same with A in tensor memory (8 columns per K = 16 step).
__device__ __forceinline__ void umma_gemm3_ts(uint32_t tmem_d, const TmemOp a, const SmemOp b, uint32_t idesc, int K, bool accumulate) {
    const uint64_t b_step = (uint64_t)((2 * b.lbo) >> 4);
    bool acc = accumulate;
    if (a.lo != NO_PLANE) {
        uint64_t bd = umma_smem_desc(b.hi, b.lbo, b.sbo);
        uint32_t at = a.lo;
#pragma unroll 1
        for (int ks = 0; ks < K / 16; ++ks) {
            umma_f16_ts(tmem_d, at, bd, idesc, acc);
            acc = true;
            at += 8;
            bd += b_step;
        }
    }
    if (b.lo) {
        uint64_t bd = umma_smem_desc(b.lo, b.lbo, b.sbo);
        uint32_t at = a.hi;
#pragma unroll 1
        for (int ks = 0; ks < K / 16; ++ks) {
            umma_f16_ts(tmem_d, at, bd, idesc, acc);
            acc = true;
            at += 8;
            bd += b_step;
        }
    }
    {
        uint64_t bd = umma_smem_desc(b.hi, b.lbo, b.sbo);
        uint32_t at = a.hi;
#pragma unroll 1
        for (int ks = 0; ks < K / 16; ++ks) {
            umma_f16_ts(tmem_d, at, bd, idesc, acc);
            acc = true;
            at += 8;
            bd += b_step;
        }
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

// Keypoint-token transformer stacks on tcgen05 tensor cores (bf16 operands, fp32 accumulation in TMEM).
//
//   mode 0 "encoder": KP_Interaction_TR.forward, model/model.py:45-126 (transformers 4.25.1 BertLayer x L):
//        h = pos_emb + x W_emb^T + b ; L x { self-attention, +res, LN, FFN(gelu), +res, LN } ;
//        pred = cls_head(h) + residual(x)
//   mode 1 "cross"  : the live TransformerDecoderLayer of updatedDecoder, model/transfusion_head.py:684-708/:132-173:
//        Q from anchor + self_posembed, K = V from tokens + cross_posembed, +res(anchor), LN2, FFN(relu), +res, LN3
//
// 21 joint tokens per sample are far below the MMA tile height, so SIX samples (126 rows) are packed into one
// M = 128 tile.  Projections / FFNs are [128 x K] x [K x N] MMAs against weights streamed into shared memory by the
// TMA engine (cp.async.bulk, 2-slot ring, prefetched two GEMMs ahead).  Attention runs on the tensor cores too:
// per head S = Q_h K_h^T is one 128 x 128 x 32 MMA over the packed tile; only the 21 x 21 block-diagonal per sample is
// kept (softmax in registers, one thread per row), the rest of P is written as zeros, and O_h = P V_h is a
// 128 x 32 x 128 MMA with V as an MN-major B operand.  The fp32 residual stream lives in TMEM columns [384,512).
//
// Thread t owns row t of the tile (TMEM lane t).  One CTA = 128 threads; thread 0 issues MMAs and TMA copies.
#include "umma.cuh"

namespace kpf {

constexpr int TS_C = 128;              // hidden size
constexpr int TS_SLOT = 2048;          // uint4 per weight slot (32 KB)
constexpr uint32_t TS_LBO = 128 * 16;  // K-major operand with 128 rows: bytes between 8-k groups
constexpr uint32_t ACC0 = 0, ACC1 = 128, ACC2 = 256, RESID = 384;

struct TokParams {
    const float* x;       // encoder: [B,J,D] ; cross: anchor [B,J,C]
    const float* y;       // cross: tokens [B,J,C]
    const uint4* wmat;    // bf16 canonical matrices (see pack_token_* in ops.py)
    const float* wvec;    // fp32 vectors
    float* tokens_out;    // encoder: [B,J,C] or null
    float* pred_out;      // encoder: [B,J,3] or null
    float* out_cj;        // cross: [B,C,J] or null
    float* out_jc;        // cross: element (b,t,c) at out_jc[(b*J+t)*stride + c0 + c] or null
    int out_jc_stride, out_jc_c0;
    int B, J, D, L, F, mode, act;
    float eps;
};

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

// Weight matrix g of the consumption sequence -> (offset in uint4, size in uint4).  Encoder: [emb] + L x {q,k,v,o,w1,w2};
// cross: L x {q,k,v,o,w1,w2}.  The encoder's optional K-tail of W_emb (D > 128) sits right after the main part.
struct WSeq {
    int emb, tail, F;  // emb: 1 if an embedding matrix leads the sequence; tail: uint4 count of its K-tail (0 or 256)
    __device__ __forceinline__ void get(int g, uint32_t& off, uint32_t& n) const {
        const uint32_t per_layer = 4 * TS_SLOT + 16 * F + (F / 8) * 128;
        uint32_t base = 0;
        if (emb) {
            if (g == 0) {
                off = 0;
                n = TS_SLOT;
                return;
            }
            base = TS_SLOT + tail;
            g -= 1;
        }
        const int l = g / 6, j = g - l * 6;
        off = base + l * per_layer;
        if (j < 4) {
            off += j * TS_SLOT;
            n = TS_SLOT;
        } else if (j == 4) {
            off += 4 * TS_SLOT;
            n = 16 * F;
        } else {
            off += 4 * TS_SLOT + 16 * F;
            n = (F / 8) * 128;
        }
    }
};

__global__ void __launch_bounds__(128, 1) token_stack_kernel(const TokParams p) {
    extern __shared__ __align__(128) unsigned char ts_smem[];
    uint4* wslot = reinterpret_cast<uint4*>(ts_smem);             // [2][2048]
    uint4* wtail = wslot + 2 * TS_SLOT;                           // [256]
    uint4* bufA = wtail + 256;                                    // [16][128]  K-major A operand (h / P / O / LN out)
    uint4* bufAt = bufA + 2048;                                   // [2][128]   K-tail of the embedding input
    uint4* bufQ = bufAt + 256;                                    // [16][128]
    uint4* bufK = bufQ + 2048;                                    // [16][128]
    uint4* bufV = bufK + 2048;                                    // MN-major [16][16][8]
    float* sVec = reinterpret_cast<float*>(bufV + 2048);          // [10][128] per-layer vectors (+ scratch)
    __shared__ __align__(8) uint64_t full[2], mma_bar, tail_bar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5;
    const int J = p.J, C = TS_C, F = p.F;
    const int SPT = 128 / J;                 // samples per tile (6 for J = 21)
    const int s_loc = tid / J, tok = tid - s_loc * J;
    const int b = blockIdx.x * SPT + s_loc;
    const bool valid = s_loc < SPT && b < p.B;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    WSeq ws{p.mode == 0 ? 1 : 0, (p.mode == 0 && p.D > C) ? 256 : 0, F};
    const int G = ws.emb + 6 * p.L;

    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        mbar_init(&mma_bar, 1);
        mbar_init(&tail_bar, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot + lane_off;  // this thread's lane window
    const uint32_t tmem0 = tmem_slot;
    uint32_t mma_phase = 0;

    auto load_w = [&](int g) {  // thread 0 only
        uint32_t off, n;
        ws.get(g, off, n);
        mbar_expect_tx(&full[g & 1], n * 16);
        tma_bulk_g2s(wslot + (g & 1) * TS_SLOT, p.wmat + off, n * 16, &full[g & 1]);
    };
    if (tid == 0) {
        load_w(0);
        if (G > 1) load_w(1);
        if (ws.tail) {
            mbar_expect_tx(&tail_bar, ws.tail * 16);
            tma_bulk_g2s(wtail, p.wmat + TS_SLOT, ws.tail * 16, &tail_bar);
        }
    }

    // GEMM g of the sequence: acc[128 x N] = A[128 x K] * W_g^T.  All threads call; returns once the MMAs have completed.
    auto run_gemm = [&](int g, const uint4* a_buf, int N, int K, uint32_t acc_col, bool with_tail) {
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            mbar_wait(&full[g & 1], (g >> 1) & 1);
            const uint32_t idesc = umma_idesc_bf16(128, N, false, false);
            umma_gemm(tmem0 + acc_col, smem_u32(a_buf), TS_LBO, 128, smem_u32(wslot + (g & 1) * TS_SLOT), (uint32_t)N * 16, 128, idesc, K,
                      false);
            if (with_tail) {
                mbar_wait(&tail_bar, 0);
                umma_gemm(tmem0 + acc_col, smem_u32(bufAt), TS_LBO, 128, smem_u32(wtail), (uint32_t)N * 16, 128, idesc, 16, true);
            }
            umma_commit(&mma_bar);
        }
        mbar_wait(&mma_bar, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
        if (tid == 0 && g + 2 < G) load_w(g + 2);  // slot g&1 is free again
    };
    auto load_vecs = [&](const float* src, int n) {
        __syncthreads();
        for (int i = tid; i < n; i += 128) sVec[i] = src[i];
        __syncthreads();
    };
    // row of fp32 -> K-major canonical chunks (this thread's row)
    auto store_row_chunks = [&](uint4* buf, int kc0, const float* v, int nchunks) {
#pragma unroll
        for (int c = 0; c < nchunks; ++c) buf[(kc0 + c) * 128 + tid] = pack8_bf16(v + 8 * c);
    };

    int g = 0;
    const float* vec = p.wvec;
    float head_x[3] = {0.f, 0.f, 0.f};  // residual(x) part of the regression head (fp32, exact)

    if (p.mode == 0) {
        // ---- embedding: h = pos_emb[tok] + x W_emb^T + b_emb      (model.py:56, :88-89)
        const int D = p.D, Dp = D > C ? C + 16 : C;
        const float* pos = vec;                       // [J][128]
        const float* bemb = pos + J * C;              // [128]
        const float* Wres = bemb + C;                 // [3][D]
        const float* bres = Wres + 3 * D;             // [3] (+1 pad)
        // (Wcls [3][128], bcls[3] follow; used at the end)
        const int shift = D - C;                      // leading non-feature inputs (3 joint coords when D = 131)
        const float* xr = p.x + ((size_t)b * J + tok) * D;
        float v[32];
        for (int c0 = 0; c0 < C; c0 += 32) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = valid ? __ldg(xr + shift + c0 + i) : 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                head_x[0] += v[i] * __ldg(Wres + shift + c0 + i);
                head_x[1] += v[i] * __ldg(Wres + D + shift + c0 + i);
                head_x[2] += v[i] * __ldg(Wres + 2 * D + shift + c0 + i);
            }
            store_row_chunks(bufA, c0 / 8, v, 4);
        }
        if (shift > 0) {
            float t[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) t[i] = (valid && i < shift) ? __ldg(xr + i) : 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (i < shift) {
                    head_x[0] += t[i] * __ldg(Wres + i);
                    head_x[1] += t[i] * __ldg(Wres + D + i);
                    head_x[2] += t[i] * __ldg(Wres + 2 * D + i);
                }
            }
            bufAt[tid] = pack8_bf16(t);
            bufAt[128 + tid] = pack8_bf16(t + 8);
        }
        head_x[0] += bres[0];
        head_x[1] += bres[1];
        head_x[2] += bres[2];
        run_gemm(g++, bufA, C, C, ACC0, shift > 0);
        (void)Dp;
        for (int c0 = 0; c0 < C; c0 += 32) {
            float a[32];
            tmem_ld32(tmem + ACC0 + c0, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = valid ? a[i] + __ldg(bemb + c0 + i) + __ldg(pos + tok * C + c0 + i) : 0.f;
            tmem_st32(tmem + RESID + c0, a);
            store_row_chunks(bufA, c0 / 8, a, 4);
        }
        vec = bres + 4 + 3 * C + 4;  // skip Wcls [3][128] + bcls[3] (+1 pad)
    } else {
        // ---- cross layer inputs: q_in = anchor + self_pos (bufA), k_in = tokens + cross_pos (bufV), resid = anchor
        const float* qpos = vec;
        const float* kpos = qpos + J * C;
        const float* ar = p.x + ((size_t)b * J + tok) * C;
        const float* yr = p.y + ((size_t)b * J + tok) * C;
        for (int c0 = 0; c0 < C; c0 += 32) {
            float a[32], q[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                a[i] = valid ? __ldg(ar + c0 + i) : 0.f;
                q[i] = valid ? a[i] + __ldg(qpos + tok * C + c0 + i) : 0.f;
            }
            tmem_st32(tmem + RESID + c0, a);
            store_row_chunks(bufA, c0 / 8, q, 4);
#pragma unroll
            for (int i = 0; i < 32; ++i) q[i] = valid ? __ldg(yr + c0 + i) + __ldg(kpos + tok * C + c0 + i) : 0.f;
            store_row_chunks(bufV, c0 / 8, q, 4);  // bufV temporarily holds k_in as a K-major operand
        }
        vec = kpos + J * C;
    }

    const float qscale = rsqrtf((float)(C / 4));  // head_dim^-0.5 with 4 heads
    const int blk_lo = s_loc * J, blk_hi = blk_lo + J;  // this row's key block in the packed tile
    const int wlo = ((32 * warp) / J) * J;               // warp-uniform span of key columns used by this warp's rows
    const int whi = min(128, ((32 * warp + 31) / J) * J + J);
    float inv_sum[4];

    for (int l = 0; l < p.L; ++l) {
        // per-layer vectors: bq bk bv bo ln1g ln1b b1 b2 ln2g ln2b
        load_vecs(vec, 10 * C);
        vec += 10 * C;
        const float *bq = sVec, *bk = sVec + C, *bv = sVec + 2 * C, *bo = sVec + 3 * C, *g1 = sVec + 4 * C, *be1 = sVec + 5 * C,
                    *b1 = sVec + 6 * C, *b2 = sVec + 7 * C, *g2 = sVec + 8 * C, *be2 = sVec + 9 * C;
        const uint4* kv_src = p.mode == 1 ? bufV : bufA;
        // ---- Q
        run_gemm(g++, bufA, C, C, ACC0, false);
        for (int c0 = 0; c0 < C; c0 += 32) {
            float a[32];
            tmem_ld32(tmem + ACC0 + c0, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = (a[i] + bq[c0 + i]) * qscale;
            store_row_chunks(bufQ, c0 / 8, a, 4);
        }
        // ---- K
        run_gemm(g++, kv_src, C, C, ACC1, false);
        for (int c0 = 0; c0 < C; c0 += 32) {
            float a[32];
            tmem_ld32(tmem + ACC1 + c0, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] += bk[c0 + i];
            store_row_chunks(bufK, c0 / 8, a, 4);
        }
        // ---- V  (MN-major B operand for P V: [token][dim], dim contiguous)
        run_gemm(g++, kv_src, C, C, ACC2, false);
        for (int c0 = 0; c0 < C; c0 += 32) {
            float a[32];
            tmem_ld32(tmem + ACC2 + c0, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] += bv[c0 + i];
#pragma unroll
            for (int c = 0; c < 4; ++c) bufV[(tid >> 3) * 128 + (c0 / 8 + c) * 8 + (tid & 7)] = pack8_bf16(a + 8 * c);
        }
        // ---- attention, one head at a time: S_h -> block-diagonal softmax -> P (bufA) -> O_h
#pragma unroll
        for (int h = 0; h < 4; ++h) {
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                umma_gemm(tmem0 + ACC0, smem_u32(bufQ) + h * 4 * TS_LBO, TS_LBO, 128, smem_u32(bufK) + h * 4 * TS_LBO, TS_LBO, 128,
                          umma_idesc_bf16(128, 128, false, false), 32, false);
                umma_commit(&mma_bar);
            }
            mbar_wait(&mma_bar, mma_phase);
            mma_phase ^= 1;
            tc_fence_after();
            // tcgen05.ld addresses are warp-uniform, so each warp sweeps the 32-column chunks that intersect the key
            // blocks of ITS rows; chunks outside that span are written as zeros without touching TMEM.
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (32 * c + 32 <= wlo || 32 * c >= whi) continue;  // warp-uniform
                float sv[32];
                tmem_ld32(tmem + ACC0 + 32 * c, sv);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int n = 32 * c + i;
                    if (n >= blk_lo && n < blk_hi) mx = fmaxf(mx, sv[i]);
                }
            }
            float sum = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float sv[32];
                if (32 * c + 32 <= wlo || 32 * c >= whi) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) bufA[(4 * c + j) * 128 + tid] = make_uint4(0, 0, 0, 0);
                    continue;
                }
                tmem_ld32(tmem + ACC0 + 32 * c, sv);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const int n = 32 * c + i;
                    sv[i] = (valid && n >= blk_lo && n < blk_hi) ? __expf(sv[i] - mx) : 0.f;
                    sum += sv[i];
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) bufA[(4 * c + j) * 128 + tid] = pack8_bf16(sv + 8 * j);
            }
            inv_sum[h] = valid ? 1.f / sum : 0.f;  // P is left un-normalised; O_h is scaled when it is read out
            fence_proxy_async();
            tc_fence_before();
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
                // O_h[128 x 32] = P[128 x 128] * V_h ; V_h = N-slice [32h, 32h+32) of the MN-major buffer
                umma_gemm(tmem0 + ACC1 + 32 * h, smem_u32(bufA), TS_LBO, 128, smem_u32(bufV) + h * 4 * 128, 16 * 128, 128,
                          umma_idesc_bf16(128, 32, false, true), 128, false);
                umma_commit(&mma_bar);
            }
            mbar_wait(&mma_bar, mma_phase);  // P (bufA) is rewritten by the next head / the O epilogue
            mma_phase ^= 1;
            tc_fence_after();
        }
        // ---- O -> bf16 A operand
#pragma unroll
        for (int hh = 0; hh < 4; ++hh) {
            float a[32];
            tmem_ld32(tmem + ACC1 + 32 * hh, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] *= inv_sum[hh];
            store_row_chunks(bufA, 4 * hh, a, 4);
        }
        // ---- attention output projection + residual + LayerNorm
        auto resid_ln = [&](uint32_t acc, const float* bias, const float* gam, const float* bet) {
            float sum = 0.f, sq = 0.f;
            for (int c0 = 0; c0 < C; c0 += 32) {
                float a[32], r[32];
                tmem_ld32(tmem + acc + c0, a);
                tmem_ld32(tmem + RESID + c0, r);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    a[i] += bias[c0 + i] + r[i];
                    sum += a[i];
                    sq += a[i] * a[i];
                }
                tmem_st32(tmem + RESID + c0, a);
            }
            const float mean = sum * (1.f / C);
            const float rstd = rsqrtf(fmaxf(sq * (1.f / C) - mean * mean, 0.f) + p.eps);
            for (int c0 = 0; c0 < C; c0 += 32) {
                float a[32];
                tmem_ld32(tmem + RESID + c0, a);
#pragma unroll
                for (int i = 0; i < 32; ++i) a[i] = valid ? (a[i] - mean) * rstd * gam[c0 + i] + bet[c0 + i] : 0.f;
                tmem_st32(tmem + RESID + c0, a);
                store_row_chunks(bufA, c0 / 8, a, 4);
            }
        };
        run_gemm(g++, bufA, C, C, ACC0, false);
        resid_ln(ACC0, bo, g1, be1);
        // ---- FFN
        run_gemm(g++, bufA, F, C, ACC2, false);
        for (int c0 = 0; c0 < F; c0 += 32) {
            float a[32];
            if (F - c0 >= 32) {
                tmem_ld32(tmem + ACC2 + c0, a);
            } else {
                tmem_ld16(tmem + ACC2 + c0, a);
#pragma unroll
                for (int i = 16; i < 32; ++i) a[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float t = a[i] + (c0 + i < F ? b1[c0 + i] : 0.f);
                a[i] = p.act == 1 ? gelu_erf(t) : fmaxf(t, 0.f);
            }
            const int nch = (F - c0 >= 32) ? 4 : 2;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < nch) bufQ[(c0 / 8 + c) * 128 + tid] = pack8_bf16(a + 8 * c);
        }
        run_gemm(g++, bufQ, C, F, ACC0, false);
        resid_ln(ACC0, b2, g2, be2);
    }

    // ---- outputs
    if (p.mode == 0) {
        const float* Wcls = p.wvec + J * C + C + 3 * p.D + 4;  // [3][128], then bcls[3]
        const float* bcls = Wcls + 3 * C;
        float pr[3] = {head_x[0] + bcls[0], head_x[1] + bcls[1], head_x[2] + bcls[2]};
        for (int c0 = 0; c0 < C; c0 += 32) {
            float a[32];
            tmem_ld32(tmem + RESID + c0, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                pr[0] += a[i] * __ldg(Wcls + c0 + i);
                pr[1] += a[i] * __ldg(Wcls + C + c0 + i);
                pr[2] += a[i] * __ldg(Wcls + 2 * C + c0 + i);
            }
            if (valid && p.tokens_out) {
                float4* o = reinterpret_cast<float4*>(p.tokens_out + ((size_t)b * J + tok) * C + c0);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
            }
        }
        if (valid && p.pred_out) {
            float* o = p.pred_out + ((size_t)b * J + tok) * 3;
            o[0] = pr[0];
            o[1] = pr[1];
            o[2] = pr[2];
        }
    } else {
        for (int c0 = 0; c0 < C; c0 += 32) {
            float a[32];
            tmem_ld32(tmem + RESID + c0, a);
            if (valid) {
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    if (p.out_cj) p.out_cj[((size_t)b * C + c0 + i) * J + tok] = a[i];
                    if (p.out_jc) p.out_jc[((size_t)b * J + tok) * p.out_jc_stride + p.out_jc_c0 + c0 + i] = a[i];
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem0, 512);
}

constexpr size_t TS_SMEM = (size_t)(2 * TS_SLOT + 256 + 2048 + 256 + 3 * 2048) * 16 + 10 * TS_C * 4;

}  // namespace kpf

extern "C" int kpf_token_stack(const float* x, const float* y, const void* wmat, const float* wvec, int mode, int B, int J, int D, int L,
                               int F, int act, float eps, float* tokens_out, float* pred_out, float* out_cj, float* out_jc,
                               int out_jc_stride, int out_jc_c0, cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(B >= 0 && J >= 1 && J <= 64 && L >= 1 && (F == 16 || F == 32 || F == 64 || F == 128));
    KPF_REQUIRE(mode == 0 ? (D == TS_C || (D > TS_C && D <= TS_C + 16)) : (mode == 1 && y != nullptr && L == 1));
    KPF_REQUIRE(((uintptr_t)wmat % 16) == 0);
    if (B == 0) return 0;
    TokParams p;
    p.x = x; p.y = y; p.wmat = (const uint4*)wmat; p.wvec = wvec; p.tokens_out = tokens_out; p.pred_out = pred_out;
    p.out_cj = out_cj; p.out_jc = out_jc; p.out_jc_stride = out_jc_stride; p.out_jc_c0 = out_jc_c0;
    p.B = B; p.J = J; p.D = D; p.L = L; p.F = F; p.mode = mode; p.act = act; p.eps = eps;
    const int spt = 128 / J;
    cudaError_t e = cudaFuncSetAttribute(token_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_SMEM);
    if (e != cudaSuccess) return (int)e;
    token_stack_kernel<<<(B + spt - 1) / spt, 128, TS_SMEM, stream>>>(p);
    KPF_CHECK_LAUNCH();
    return 0;
}

// Keypoint-token transformer stacks on tcgen05 tensor cores (bf16 operands, fp32 accumulation in TMEM).
//
//   encoder : KP_Interaction_TR.forward, model/model.py:45-126 (transformers 4.25.1 BertLayer x L):
//             h = pos_emb + x W_emb^T + b ; L x { self-attention, +res, LN, FFN(gelu), +res, LN } ; pred = cls_head(h) + residual(x)
//             optional prologue: x = relu(W_fu [desa_0 | desa_1 | desa_2 | jf] + b)   (DESA's fusion conv, model.py:160-164, :203)
//   cross   : the live TransformerDecoderLayer of updatedDecoder, model/transfusion_head.py:684-708 / :132-173:
//             Q from anchor + self_posembed, K = V from tokens + cross_posembed, +res(anchor), LN2, FFN(relu), +res, LN3
//   cross + encoder fused: crossTR followed by final_TR on cat([r3d, cross_out]) (model.py:347-349) without leaving the SM.
//
// Tile = 4 samples x 32 rows (J <= 32 joint tokens per sample, zero padded): warp q of each warpgroup owns sample q, so the
// 32 x 32 key block of a row is ONE 32-column TMEM chunk at a warp-uniform address.  256 threads: two threads per row
// (tid and tid+128) split every epilogue's columns; for attention the two warpgroups take different heads.
// Projections / FFNs are [128 x K] x [K x N] MMAs against weights streamed by the TMA engine (cp.async.bulk, 2-slot ring).
// Per head S = Q_h K_h^T is a 128 x 128 x 32 MMA of which the block diagonal is kept (softmax in registers), P is written
// with zeros elsewhere, and O_h = P V_h is a 128 x 32 x 128 MMA with V as an MN-major B operand.  The fp32 residual stream
// lives in TMEM columns [384,512).  Thread 0 issues MMAs and TMA copies.
#include "umma.cuh"

namespace kpf {

constexpr int TS_C = 128;              // hidden size
constexpr int TS_SLOT = 2048;          // uint4 per weight slot (32 KB)
constexpr uint32_t TS_LBO = 128 * 16;  // K-major operand with 128 rows: bytes between 8-k groups
constexpr uint32_t ACC0 = 0, ACC1 = 128, ACC2 = 256, RESID = 384;

struct TokParams {
    const float* x;        // encoder input [B,J,D] (no prologue) | cross: anchor [B,J,C]
    const float* y;        // cross: tokens [B,J,C]
    const float* r3d;      // cross+encoder: leading D-128 inputs of the encoder [B,J,D-128]
    const float* desa;     // prologue: [B,3,J,C]
    const float* jf;       // prologue: [B,J,C]
    const uint4* wmat;     // bf16 canonical matrices
    const int2* wseq;      // (offset, count) in uint4 of weight g of the consumption sequence
    const float* wvec;     // fp32 vectors
    float* tokens_out;     // [B,J,C] or null (final hidden states)
    float* pred_out;       // [B,J,3] or null
    float* out_cj;         // cross only: [B,C,J] or null
    float* out_jc;         // cross only: element (b,t,c) at out_jc[(b*J+t)*stride + c0 + c] or null
    int out_jc_stride, out_jc_c0;
    int B, J, D, L, F, pre, cross, Fc, G;
    long long* dbg;        // optional: clock64 stamps of CTA 0 (profiling aid)
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

// erf-GELU with Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7): the exact erff costs ~2k cycles per 16-wide FFN epilogue
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    const float t = __fdividef(1.f, 1.f + 0.3275911f * z);
    const float poly = ((((1.061405429f * t - 1.453152027f) * t + 1.421413741f) * t - 0.284496736f) * t + 0.254829592f) * t;
    const float erf_abs = 1.f - poly * __expf(-z * z);
    return 0.5f * x * (1.f + copysignf(erf_abs, x));
}

// 32 consecutive floats of a row; 128-bit loads when the row is 16-byte aligned (it is, except for D = 131 inputs)
__device__ __forceinline__ void load_row32(const float* __restrict__ src, float* v, bool valid) {
    if (!valid) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0.f;
    } else if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 t = __ldg(s4 + i);
            v[4 * i] = t.x;
            v[4 * i + 1] = t.y;
            v[4 * i + 2] = t.z;
            v[4 * i + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __ldg(src + i);
    }
}

__global__ void __launch_bounds__(256, 1) token_stack_kernel(const TokParams p) {
    extern __shared__ __align__(128) unsigned char ts_smem[];
    uint4* wslot = reinterpret_cast<uint4*>(ts_smem);             // [2][2048]
    uint4* wtail = wslot + 2 * TS_SLOT;                           // [256]
    uint4* bufA = wtail + 256;                                    // [16][128]  K-major A operand (h / P / O / LN out)
    uint4* bufAt = bufA + 2048;                                   // [2][128]   K-tail of the embedding input
    uint4* bufQ = bufAt + 256;                                    // [16][128]
    uint4* bufK = bufQ + 2048;                                    // [16][128]
    uint4* bufV = bufK + 2048;                                    // MN-major [16][16][8]
    float* sVec = reinterpret_cast<float*>(bufV + 2048);          // [10][128] per-layer vectors
    float* sRed = sVec + 10 * TS_C;                               // LayerNorm / head partials exchanged by the two threads of a row
    __shared__ __align__(8) uint64_t full[2], mma_bar, tail_bar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, row = tid & 127, half = tid >> 7, wq = (tid >> 5) & 3;
    const int J = p.J, C = TS_C;
    const int tok = row & 31;
    const int b = blockIdx.x * 4 + wq;                   // warp-uniform sample
    const bool valid = tok < J && b < p.B;
    const int cb = half * 64;                            // this thread's column half of every 128-wide epilogue
    const int G = p.G;

    if (tid < 32) tmem_alloc(&tmem_slot, 512);
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
    const uint32_t tmem0 = tmem_slot;
    const uint32_t tmem = tmem0 + ((uint32_t)(wq * 32) << 16);  // this thread's lane window
    uint32_t mma_phase = 0;
    // sequence index of the embedding GEMM that owns the K-tail (its tail weights follow its main part in wmat)
    const int tail_g = (p.L > 0 && p.D > C) ? (p.cross ? 6 : 0) + (p.pre ? 4 : 0) : -1;

    auto load_w = [&](int gi) {  // thread 0 only
        const int2 s = p.wseq[gi];
        mbar_expect_tx(&full[gi & 1], (uint32_t)s.y * 16);
        tma_bulk_g2s(wslot + (gi & 1) * TS_SLOT, p.wmat + s.x, (uint32_t)s.y * 16, &full[gi & 1]);
    };
    if (tid == 0) {
        load_w(0);
        if (G > 1) load_w(1);
        if (tail_g >= 0) {
            const int2 s = p.wseq[tail_g];
            mbar_expect_tx(&tail_bar, 256 * 16);
            tma_bulk_g2s(wtail, p.wmat + s.x + s.y, 256 * 16, &tail_bar);
        }
    }
    // operand writes -> async proxy, everybody's TMEM reads done, then one elected thread issues
    auto sync_for_mma = [&]() {
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
    };
    auto wait_mma = [&]() {
        mbar_wait(&mma_bar, mma_phase);
        mma_phase ^= 1;
        tc_fence_after();
    };
    // GEMM gi of the sequence: acc[128 x N] (+)= A[128 x K] * W^T.  All threads call; returns once the MMAs have completed.
    auto run_gemm = [&](int gi, const uint4* a_buf, int N, int K, uint32_t acc_col, bool with_tail, bool accumulate) {
        sync_for_mma();
        if (tid == 0) {
            tc_fence_after();
            mbar_wait(&full[gi & 1], (gi >> 1) & 1);
            const uint32_t idesc = umma_idesc_bf16(128, N, false, false);
            umma_gemm(tmem0 + acc_col, smem_u32(a_buf), TS_LBO, 128, smem_u32(wslot + (gi & 1) * TS_SLOT), (uint32_t)N * 16, 128, idesc, K,
                      accumulate);
            if (with_tail) {
                mbar_wait(&tail_bar, 0);
                umma_gemm(tmem0 + acc_col, smem_u32(bufAt), TS_LBO, 128, smem_u32(wtail), (uint32_t)N * 16, 128, idesc, 16, true);
            }
            umma_commit(&mma_bar);
        }
        wait_mma();
        if (tid == 0 && gi + 2 < G) load_w(gi + 2);  // slot gi&1 is free again
    };
    auto load_vecs = [&](const float* src, int n) {
        __syncthreads();
        for (int i = tid; i < n; i += 256) sVec[i] = src[i];
        __syncthreads();
    };
    auto store_row_chunks = [&](uint4* buf, int kc0, const float* v) {  // 32 fp32 of this row -> 4 K-major chunks
#pragma unroll
        for (int c = 0; c < 4; ++c) buf[(kc0 + c) * 128 + row] = pack8_bf16(v + 8 * c);
    };

    int g = 0, n_stamp = 0;
    auto stamp = [&]() {
        if (p.dbg && blockIdx.x == 0 && tid == 0 && n_stamp < 64) p.dbg[n_stamp] = clock64();
        ++n_stamp;
    };
    stamp();
    const float* vec = p.wvec;
    float head_x[3] = {0.f, 0.f, 0.f};  // this thread's share of residual(x) of the regression head (fp32)
    const float qscale = rsqrtf((float)(C / 4));  // head_dim^-0.5, 4 heads

    // One transformer layer.  kv_src: K-major operand the K / V projections read (bufA for self-attention).
    // Wres_feat != null: accumulate the encoder's residual() head over this layer's output (cross -> final_TR fusion).
    auto layer = [&](const uint4* kv_src, int F, int act, float eps, const float* Wres_feat) {
        const float *bq = sVec, *bk = sVec + C, *bv = sVec + 2 * C, *bo = sVec + 3 * C, *g1 = sVec + 4 * C, *be1 = sVec + 5 * C,
                    *b1 = sVec + 6 * C, *b2 = sVec + 7 * C, *g2 = sVec + 8 * C, *be2 = sVec + 9 * C;
        // ---- Q, K, V projections (three weight tiles); each thread drains its 64 columns
        run_gemm(g++, bufA, C, C, ACC0, false, false);
        {
            float a[64];
            tmem_ld64(tmem + ACC0 + cb, a);
#pragma unroll
            for (int i = 0; i < 64; ++i) a[i] = (a[i] + bq[cb + i]) * qscale;
#pragma unroll
            for (int c = 0; c < 8; ++c) bufQ[(cb / 8 + c) * 128 + row] = pack8_bf16(a + 8 * c);
        }
        run_gemm(g++, kv_src, C, C, ACC1, false, false);
        {
            float a[64];
            tmem_ld64(tmem + ACC1 + cb, a);
#pragma unroll
            for (int i = 0; i < 64; ++i) a[i] += bk[cb + i];
#pragma unroll
            for (int c = 0; c < 8; ++c) bufK[(cb / 8 + c) * 128 + row] = pack8_bf16(a + 8 * c);
        }
        run_gemm(g++, kv_src, C, C, ACC2, false, false);
        {   // V: MN-major B operand for P V ([token][dim], dim contiguous)
            float a[64];
            tmem_ld64(tmem + ACC2 + cb, a);
#pragma unroll
            for (int i = 0; i < 64; ++i) a[i] += bv[cb + i];
#pragma unroll
            for (int c = 0; c < 8; ++c) bufV[(row >> 3) * 128 + (cb / 8 + c) * 8 + (row & 7)] = pack8_bf16(a + 8 * c);
        }
        stamp();
        // ---- attention.  Warpgroup `half` owns heads 2*half and 2*half+1 (= its 64 output columns).  Round pr handles head pr
        //      (group 0, S in ACC0) and head 2+pr (group 1, S in ACC1) concurrently; the P V MMAs take turns on bufA.
        float inv_sum[2];
#pragma unroll
        for (int pr = 0; pr < 2; ++pr) {
            sync_for_mma();
            if (tid == 0) {
                tc_fence_after();
                const uint32_t idS = umma_idesc_bf16(128, 128, false, false);
                umma_gemm(tmem0 + ACC0, smem_u32(bufQ) + pr * 4 * TS_LBO, TS_LBO, 128, smem_u32(bufK) + pr * 4 * TS_LBO, TS_LBO, 128, idS, 32,
                          false);
                umma_gemm(tmem0 + ACC1, smem_u32(bufQ) + (2 + pr) * 4 * TS_LBO, TS_LBO, 128, smem_u32(bufK) + (2 + pr) * 4 * TS_LBO, TS_LBO,
                          128, idS, 32, false);
                umma_commit(&mma_bar);
            }
            wait_mma();
            // block-diagonal softmax: this row's keys are columns [32*wq, 32*wq + J) of its head's S
            float pv[32];
            tmem_ld32(tmem + (half ? ACC1 : ACC0) + 32 * wq, pv);
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, i < J ? pv[i] : -INFINITY);
            float sum = 0.f;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                pv[i] = (valid && i < J) ? __expf(pv[i] - mx) : 0.f;
                sum += pv[i];
            }
            inv_sum[pr] = valid ? 1.f / sum : 0.f;  // P stays un-normalised; O_h is scaled when it is read out
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                if (half == hh) {  // this warpgroup's P -> bufA; zeros outside the row's 32-column block are laid down once per
                                   // layer (first head): later heads overwrite the same 4 chunks and nothing else
                    if (pr == 0 && hh == 0) {
#pragma unroll
                        for (int kc = 0; kc < 16; ++kc) bufA[kc * 128 + row] = make_uint4(0, 0, 0, 0);
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c) bufA[(4 * wq + c) * 128 + row] = pack8_bf16(pv + 8 * c);
                }
                sync_for_mma();
                if (tid == 0) {
                    tc_fence_after();
                    const int h = 2 * hh + pr;  // O_h[128 x 32] = P V_h, V_h = N-slice [32h, 32h+32) of the MN-major buffer
                    umma_gemm(tmem0 + ACC2 + 32 * h, smem_u32(bufA), TS_LBO, 128, smem_u32(bufV) + h * 4 * 128, 16 * 128, 128,
                              umma_idesc_bf16(128, 32, false, true), 128, false);
                    umma_commit(&mma_bar);
                }
                wait_mma();  // bufA is rewritten next
            }
        }
        stamp();
        // ---- O -> bf16 A operand: thread drains heads 2*half (+0, +1) = columns [cb, cb+64), scaling by its softmax sums
        {
            float a[64];
            tmem_ld64(tmem + ACC2 + cb, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                a[i] *= inv_sum[0];
                a[32 + i] *= inv_sum[1];
            }
#pragma unroll
            for (int c = 0; c < 8; ++c) bufA[(cb / 8 + c) * 128 + row] = pack8_bf16(a + 8 * c);
        }
        // ---- residual + LayerNorm on a 128-wide accumulator; the two threads of a row exchange partial statistics
        auto resid_ln = [&](uint32_t acc, const float* bias, const float* gam, const float* bet, const float* Wrf) {
            float y[64];
            float sum = 0.f, sq = 0.f;
            {
                float r[64];
                tmem_ld64(tmem + acc + cb, y);
                tmem_ld64(tmem + RESID + cb, r);
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    y[i] += bias[cb + i] + r[i];
                    sum += y[i];
                    sq += y[i] * y[i];
                }
            }
            sRed[(half * 128 + row) * 2] = sum;
            sRed[(half * 128 + row) * 2 + 1] = sq;
            __syncthreads();
            sum += sRed[((half ^ 1) * 128 + row) * 2];
            sq += sRed[((half ^ 1) * 128 + row) * 2 + 1];
            __syncthreads();
            const float mean = sum * (1.f / C);
            const float rstd = rsqrtf(fmaxf(sq * (1.f / C) - mean * mean, 0.f) + eps);
#pragma unroll
            for (int i = 0; i < 64; ++i) y[i] = valid ? (y[i] - mean) * rstd * gam[cb + i] + bet[cb + i] : 0.f;
            tmem_st32(tmem + RESID + cb, y);
            tmem_st32(tmem + RESID + cb + 32, y + 32);
#pragma unroll
            for (int c = 0; c < 8; ++c) bufA[(cb / 8 + c) * 128 + row] = pack8_bf16(y + 8 * c);
            if (Wrf) {
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    head_x[0] += y[i] * __ldg(Wrf + cb + i);
                    head_x[1] += y[i] * __ldg(Wrf + p.D + cb + i);
                    head_x[2] += y[i] * __ldg(Wrf + 2 * p.D + cb + i);
                }
            }
        };
        run_gemm(g++, bufA, C, C, ACC0, false, false);
        resid_ln(ACC0, bo, g1, be1, nullptr);
        stamp();
        // ---- FFN
        run_gemm(g++, bufA, F, C, ACC2, false, false);
        for (int c0 = cb; c0 < cb + 64 && c0 < F; c0 += 32) {
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
                a[i] = act == 1 ? gelu_erf(t) : fmaxf(t, 0.f);
            }
            const int nch = (F - c0 >= 32) ? 4 : 2;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                if (c < nch) bufQ[(c0 / 8 + c) * 128 + row] = pack8_bf16(a + 8 * c);
        }
        run_gemm(g++, bufQ, C, F, ACC0, false, false);
        resid_ln(ACC0, b2, g2, be2, Wres_feat);
        stamp();
    };

    // =========================== cross-attention layer (crossTR) ===========================
    if (p.cross) {
        const float* qpos = vec;
        const float* kpos = qpos + J * C;
        const float* ar = p.x + ((size_t)b * J + tok) * C;
        const float* yr = p.y + ((size_t)b * J + tok) * C;
        for (int c0 = cb; c0 < cb + 64; c0 += 32) {
            float a[32], q[32], e[32];
            load_row32(ar + c0, a, valid);
            load_row32(qpos + tok * C + c0, e, valid);
#pragma unroll
            for (int i = 0; i < 32; ++i) q[i] = a[i] + e[i];
            tmem_st32(tmem + RESID + c0, a);           // residual = anchor (transfusion_head.py:164)
            store_row_chunks(bufA, c0 / 8, q);
            load_row32(yr + c0, q, valid);
            load_row32(kpos + tok * C + c0, e, valid);
#pragma unroll
            for (int i = 0; i < 32; ++i) q[i] += e[i];
            store_row_chunks(bufV, c0 / 8, q);         // bufV temporarily holds k_in as a K-major operand
        }
        vec = kpos + J * C;
        load_vecs(vec, 10 * C);
        vec += 10 * C;
        // the fused encoder's residual.weight feature columns ([3][D], features start at column D-128)
        const float* Wres_feat = p.L > 0 ? vec + (size_t)J * C + C + (p.D - C) : nullptr;
        layer(bufV, p.Fc, 0, 1e-5f, Wres_feat);
        if (p.L == 0) {
            for (int c0 = cb; c0 < cb + 64; c0 += 32) {
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
    }

    // =========================== encoder (KP_Interaction_TR) ===========================
    if (p.L > 0) {
        const int D = p.D, shift = D - C;
        if (p.pre) {
            // ---- DESA fusion conv: x = relu(W_fu [desa_0 | desa_1 | desa_2 | jf] + b_fu), four accumulating K = 128 steps
            const float* bfu = vec;
            vec += C;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                uint4* dst = s == 0 ? bufA : (s == 1 ? bufQ : (s == 2 ? bufK : bufV));
                const float* src = s < 3 ? p.desa + (((size_t)b * 3 + s) * J + tok) * C : p.jf + ((size_t)b * J + tok) * C;
                for (int c0 = cb; c0 < cb + 64; c0 += 32) {
                    float a[32];
                    load_row32(src + c0, a, valid);
                    store_row_chunks(dst, c0 / 8, a);
                }
            }
            run_gemm(g++, bufA, C, C, ACC0, false, false);
            run_gemm(g++, bufQ, C, C, ACC0, false, true);
            run_gemm(g++, bufK, C, C, ACC0, false, true);
            run_gemm(g++, bufV, C, C, ACC0, false, true);
            const float* Wres0 = vec + (size_t)J * C + C;
            for (int c0 = cb; c0 < cb + 64; c0 += 32) {
                float a[32];
                tmem_ld32(tmem + ACC0 + c0, a);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    a[i] = valid ? fmaxf(a[i] + __ldg(bfu + c0 + i), 0.f) : 0.f;
                    head_x[0] += a[i] * __ldg(Wres0 + c0 + i);
                    head_x[1] += a[i] * __ldg(Wres0 + D + c0 + i);
                    head_x[2] += a[i] * __ldg(Wres0 + 2 * D + c0 + i);
                }
                store_row_chunks(bufA, c0 / 8, a);
            }
        }
        const float* pos = vec;                       // [J][128]
        const float* bemb = pos + J * C;              // [128]
        const float* Wres = bemb + C;                 // [3][D]
        const float* bres = Wres + 3 * D;             // [3] (+1 pad)
        const float* Wcls = bres + 4;                 // [3][128]
        const float* bcls = Wcls + 3 * C;             // [3] (+1 pad)
        if (!p.pre && !p.cross) {
            const float* xr = p.x + ((size_t)b * J + tok) * D;
            for (int c0 = cb; c0 < cb + 64; c0 += 32) {
                float v[32];
                load_row32(xr + shift + c0, v, valid);
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    head_x[0] += v[i] * __ldg(Wres + shift + c0 + i);
                    head_x[1] += v[i] * __ldg(Wres + D + shift + c0 + i);
                    head_x[2] += v[i] * __ldg(Wres + 2 * D + shift + c0 + i);
                }
                store_row_chunks(bufA, c0 / 8, v);
            }
        }
        if (shift > 0 && half == 0) {  // leading (D - 128) inputs: joint coordinates
            const float* lead = p.cross ? p.r3d + ((size_t)b * J + tok) * shift : p.x + ((size_t)b * J + tok) * D;
            float t[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) t[i] = (valid && i < shift) ? __ldg(lead + i) : 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                if (i < shift) {
                    head_x[0] += t[i] * __ldg(Wres + i);
                    head_x[1] += t[i] * __ldg(Wres + D + i);
                    head_x[2] += t[i] * __ldg(Wres + 2 * D + i);
                }
            }
            bufAt[row] = pack8_bf16(t);
            bufAt[128 + row] = pack8_bf16(t + 8);
        }
        // ---- embedding: h = pos_emb[tok] + x W_emb^T + b_emb      (model.py:56, :88-89)
        run_gemm(g++, bufA, C, C, ACC0, shift > 0, false);
        for (int c0 = cb; c0 < cb + 64; c0 += 32) {
            float a[32], e[32];
            tmem_ld32(tmem + ACC0 + c0, a);
            load_row32(pos + tok * C + c0, e, valid);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = valid ? a[i] + __ldg(bemb + c0 + i) + e[i] : 0.f;
            tmem_st32(tmem + RESID + c0, a);
            store_row_chunks(bufA, c0 / 8, a);
        }
        vec = bcls + 4;
        stamp();
        for (int l = 0; l < p.L; ++l) {
            load_vecs(vec, 10 * C);
            vec += 10 * C;
            layer(bufA, p.F, 1, 1e-12f, nullptr);
        }
        // ---- regression head: pred = cls_head(h) + residual(x)   (model.py:122-124), fp32, two threads per row
        float pr3[3] = {head_x[0], head_x[1], head_x[2]};
        for (int c0 = cb; c0 < cb + 64; c0 += 32) {
            float a[32];
            tmem_ld32(tmem + RESID + c0, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                pr3[0] += a[i] * __ldg(Wcls + c0 + i);
                pr3[1] += a[i] * __ldg(Wcls + C + c0 + i);
                pr3[2] += a[i] * __ldg(Wcls + 2 * C + c0 + i);
            }
            if (valid && p.tokens_out) {
                float4* o = reinterpret_cast<float4*>(p.tokens_out + ((size_t)b * J + tok) * C + c0);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
            }
        }
        __syncthreads();
        if (half == 1) {
            sRed[row * 3] = pr3[0];
            sRed[row * 3 + 1] = pr3[1];
            sRed[row * 3 + 2] = pr3[2];
        }
        __syncthreads();
        if (half == 0 && valid && p.pred_out) {
            float* o = p.pred_out + ((size_t)b * J + tok) * 3;
            o[0] = pr3[0] + sRed[row * 3] + bres[0] + bcls[0];
            o[1] = pr3[1] + sRed[row * 3 + 1] + bres[1] + bcls[1];
            o[2] = pr3[2] + sRed[row * 3 + 2] + bres[2] + bcls[2];
        }
    }
    stamp();
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem0, 512);
}

constexpr size_t TS_SMEM = (size_t)(2 * TS_SLOT + 256 + 2048 + 256 + 3 * 2048) * 16 + (10 * TS_C + 2 * 128 * 3) * 4;

}  // namespace kpf

extern "C" int kpf_token_stack(const float* x, const float* y, const float* r3d, const float* desa, const float* jf, const void* wmat,
                               const void* wseq, const float* wvec, int n_weights, int cross, int pre, int B, int J, int D, int L, int F,
                               int Fc, float* tokens_out, float* pred_out, float* out_cj, float* out_jc, int out_jc_stride, int out_jc_c0,
                               long long* dbg, cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(B >= 0 && J >= 1 && J <= 32 && L >= 0 && (cross || L > 0));
    KPF_REQUIRE(L == 0 || F == 16 || F == 32 || F == 64 || F == 128);
    KPF_REQUIRE(!cross || (y != nullptr && (Fc == 16 || Fc == 32 || Fc == 64 || Fc == 128)));
    KPF_REQUIRE(L == 0 || D == TS_C || (D > TS_C && D <= TS_C + 16));
    KPF_REQUIRE(!pre || (desa != nullptr && jf != nullptr && !cross && D == TS_C));
    KPF_REQUIRE(!(cross && L > 0) || (r3d != nullptr && D > TS_C));
    KPF_REQUIRE(n_weights == (cross ? 6 : 0) + (pre ? 4 : 0) + (L > 0 ? 1 + 6 * L : 0));
    KPF_REQUIRE(((uintptr_t)wmat % 16) == 0 && ((uintptr_t)wseq % 8) == 0);
    if (B == 0) return 0;
    TokParams p;
    p.x = x; p.y = y; p.r3d = r3d; p.desa = desa; p.jf = jf; p.wmat = (const uint4*)wmat; p.wseq = (const int2*)wseq; p.wvec = wvec;
    p.tokens_out = tokens_out; p.pred_out = pred_out; p.out_cj = out_cj; p.out_jc = out_jc; p.out_jc_stride = out_jc_stride;
    p.out_jc_c0 = out_jc_c0; p.B = B; p.J = J; p.D = D; p.L = L; p.F = F; p.pre = pre; p.cross = cross; p.Fc = Fc; p.G = n_weights; p.dbg = dbg;
    cudaError_t e = cudaFuncSetAttribute(token_stack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TS_SMEM);
    if (e != cudaSuccess) return (int)e;
    token_stack_kernel<<<(B + 3) / 4, 256, TS_SMEM, stream>>>(p);
    KPF_CHECK_LAUNCH();
    return 0;
}

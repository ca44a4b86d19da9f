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
// 32 x 32 key block of a row is ONE 32-column TMEM chunk at a warp-uniform address.  TS_NT threads: TS_CG threads per row
// (tid, tid+128, ...) split every epilogue's columns; the softmax of the two heads of a round runs in column groups 0 / 1.
// Projections / FFNs are [128 x K] x [K x N] MMAs against weights streamed by the TMA engine (cp.async.bulk, 2-slot ring).
// Per head S = Q_h K_h^T is a 128 x 128 x 32 MMA of which the block diagonal is kept (softmax in registers), P is written
// with zeros elsewhere, and O_h = P V_h is a 128 x 32 x 128 MMA with V as an MN-major B operand.  The fp32 residual stream
// lives in TMEM columns [384,512).  Thread 0 issues MMAs and TMA copies.
#include "tmem_ldst.cuh"

namespace kpf {

constexpr int TS_C = 128;              // hidden size
constexpr int TS_SLOT = 2048;          // uint4 per weight slot (32 KB)
constexpr uint32_t TS_LBO = 128 * 16;  // K-major operand with 128 rows: bytes between 8-k groups
constexpr int TS_MAXG = 64;            // weight tiles per program
constexpr uint32_t ACC0 = 0, ACC1 = 128, ACC2 = 256, RESID = 384;
// Threads per CTA.  128 rows x TS_CG threads per row: every 128-wide epilogue is split into TS_CW-column pieces, so the
// serial instruction stream of a thread (and the SASS the SM has to fetch) shrinks with TS_CG while the warps per
// scheduler that hide TMEM / shared-memory latency grow with it.
constexpr int TS_NT = 512;
constexpr int TS_CG = TS_NT / 128;
constexpr int TS_CW = 128 / TS_CG;
constexpr int TS_CH = TS_CW < 32 ? TS_CW : 32;  // columns per prologue piece
constexpr int TS_FU = TS_CG >= 8 ? 2 : (TS_CG == 4 ? 4 : 8);  // FFN hidden columns per thread per pass

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

// erf-GELU with Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7): the exact erff costs ~2k cycles per 16-wide FFN epilogue
__device__ __forceinline__ float gelu_erf(float x) {
    const float z = fabsf(x) * 0.70710678118654752f;
    const float t = __fdividef(1.f, 1.f + 0.3275911f * z);
    const float poly = ((((1.061405429f * t - 1.453152027f) * t + 1.421413741f) * t - 0.284496736f) * t + 0.254829592f) * t;
    const float erf_abs = 1.f - poly * __expf(-z * z);
    return 0.5f * x * (1.f + copysignf(erf_abs, x));
}

// N consecutive floats of a row.  ALIGNED: the row start is 16-byte aligned (everything except the D = 131 inputs).
template <int N, bool ALIGNED>
__device__ __forceinline__ void load_row(const float* __restrict__ src, float* v, bool valid) {
    if (!valid) {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = 0.f;
    } else if (ALIGNED) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            const float4 t = __ldg(s4 + i);
            v[4 * i] = t.x;
            v[4 * i + 1] = t.y;
            v[4 * i + 2] = t.z;
            v[4 * i + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] = __ldg(src + i);
    }
}

// acc[k] += sum_i v[i] * W[k][i] for the three 16-byte aligned rows W[k] = w + k * ld (the fp32 regression-head shares)
template <int N>
__device__ __forceinline__ void head_acc(float* acc, const float* v, const float* __restrict__ w, int ld) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float4* w4 = reinterpret_cast<const float4*>(w + (size_t)k * ld);
#pragma unroll
        for (int i = 0; i < N / 4; ++i) {
            const float4 t = __ldg(w4 + i);
            acc[k] += v[4 * i] * t.x + v[4 * i + 1] * t.y + v[4 * i + 2] * t.z + v[4 * i + 3] * t.w;
        }
    }
}

__global__ void __launch_bounds__(TS_NT, 1) token_stack_kernel(const TokParams p) {
    constexpr int CG = TS_CG, CW = TS_CW, CH = TS_CH, NCH = TS_CW / 8;
    extern __shared__ __align__(128) unsigned char ts_smem[];
    uint4* wslot = reinterpret_cast<uint4*>(ts_smem);             // [2][2048]
    uint4* wtail = wslot + 2 * TS_SLOT;                           // [256]
    uint4* bufA = wtail + 256;                                    // [16][128]  K-major A operand (h / P / O / LN out)
    uint4* bufAt = bufA + 2048;                                   // [2][128]   K-tail of the embedding input
    uint4* bufQ = bufAt + 256;                                    // [16][128]
    uint4* bufK = bufQ + 2048;                                    // [16][128]
    uint4* bufV = bufK + 2048;                                    // MN-major [16][16][8]
    float* sVec = reinterpret_cast<float*>(bufV + 2048);          // [10][128] per-layer vectors
    float* sInv = sVec + 10 * TS_C;                               // [4][128] 1 / softmax sum of (head, row)
    float* sRed = sInv + 4 * TS_C;                                // [2][CG][128][2] LayerNorm partials (double buffered)
    __shared__ __align__(8) uint64_t full[2], mma_bar, tail_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ int2 sSeq[TS_MAXG];   // the weight sequence table: the issuing lane must not sit behind a global load per GEMM

    const int tid = threadIdx.x, row = tid & 127, cg = tid >> 7, wq = (tid >> 5) & 3;
    const int J = p.J, C = TS_C;
    const int tok = row & 31;
    const int b = blockIdx.x * 4 + wq;                   // warp-uniform sample
    const bool valid = tok < J && b < p.B;
    const int cb = cg * CW;                              // this thread's columns of every 128-wide epilogue
    const int G = p.G;
    const int warp_u = warp_index_uniform();             // MMA / TMA issue: one elected lane of warp 0

    if (tid < 32) tmem_alloc(&tmem_slot, 512);
    if (tid >= 32 && tid < 32 + G) sSeq[tid - 32] = p.wseq[tid - 32];
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
    int n_fine = 0;
    bool fine_on = false;
    auto fine = [&]() {   // second stamp series (dbg[32..63]): inside the first encoder layer
        if (p.dbg && fine_on && blockIdx.x == 0 && tid == 0 && n_fine < 32) p.dbg[32 + n_fine++] = clock64();
    };
    // sequence index of the embedding GEMM that owns the K-tail (its tail weights follow its main part in wmat)
    const int tail_g = (p.L > 0 && p.D > C) ? (p.cross ? 6 : 0) + (p.pre ? 4 : 0) : -1;

    auto load_w = [&](int gi) {  // one elected lane of warp 0
        const int2 s = sSeq[gi];
        mbar_expect_tx(&full[gi & 1], (uint32_t)s.y * 16);
        tma_bulk_g2s(wslot + (gi & 1) * TS_SLOT, p.wmat + s.x, (uint32_t)s.y * 16, &full[gi & 1]);
    };
    if (warp_u == 0) {
        if (elect_one()) {
            load_w(0);
            if (G > 1) load_w(1);
            if (tail_g >= 0) {
                const int2 s = sSeq[tail_g];
                mbar_expect_tx(&tail_bar, 256 * 16);
                tma_bulk_g2s(wtail, p.wmat + s.x + s.y, 256 * 16, &tail_bar);
            }
        }
        __syncwarp();
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
        fine();
        if (warp_u == 0) {
            tc_fence_after();
            mbar_wait(&full[gi & 1], (gi >> 1) & 1);
            if (with_tail) mbar_wait(&tail_bar, 0);
            fine();
            if (elect_one()) {
                const uint32_t idesc = umma_idesc_bf16(128, N, false, false);
                umma_gemm(tmem0 + acc_col, smem_u32(a_buf), TS_LBO, 128, smem_u32(wslot + (gi & 1) * TS_SLOT), (uint32_t)N * 16, 128, idesc,
                          K, accumulate);
                if (with_tail) umma_gemm(tmem0 + acc_col, smem_u32(bufAt), TS_LBO, 128, smem_u32(wtail), (uint32_t)N * 16, 128, idesc, 16, true);
                umma_commit(&mma_bar);
            }
            __syncwarp();
            fine();
        }
        wait_mma();
        fine();
        if (warp_u == 0 && gi + 2 < G) {  // slot gi&1 is free again
            if (elect_one()) load_w(gi + 2);
            __syncwarp();
        }
    };
    auto load_vecs = [&](const float* src, int n) {
        __syncthreads();
        for (int i = tid; i < n / 4; i += TS_NT) reinterpret_cast<float4*>(sVec)[i] = __ldg(reinterpret_cast<const float4*>(src) + i);
        __syncthreads();
    };
    // N fp32 of this row starting at column c0 -> K-major bf16 chunks
    auto store_chunks = [&](uint4* buf, int c0, const float* v, int n) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (c < n / 8) buf[(c0 / 8 + c) * 128 + row] = pack8_bf16(v + 8 * c);
    };
    // acc[:, cb .. cb+CW) + bias (* scale) -> bf16 chunks of a K-major buffer
    auto drain_kmajor = [&](uint32_t acc, const float* bias, float scale, uint4* dst) {
        float a[CW];
        tmem_ld<CW>(tmem + acc + cb, a);
#pragma unroll
        for (int i = 0; i < CW; ++i) a[i] = (a[i] + bias[cb + i]) * scale;
#pragma unroll
        for (int c = 0; c < NCH; ++c) dst[(cb / 8 + c) * 128 + row] = pack8_bf16(a + 8 * c);
    };

    int g = 0, n_stamp = 0, red_par = 0;
    auto stamp = [&]() {
        if (p.dbg && blockIdx.x == 0 && tid == 0 && n_stamp < 32) p.dbg[n_stamp] = clock64();
        ++n_stamp;
    };
    stamp();
    const float* vec = p.wvec;
    float head_x[3] = {0.f, 0.f, 0.f};  // this thread's share of residual(x) of the regression head (fp32)
    const float qscale = rsqrtf((float)(C / 4));  // head_dim^-0.5, 4 heads
    const int D = p.D, shift = p.L > 0 ? D - C : 0;
    const float *pos = nullptr, *bemb = nullptr, *Wres_lead = nullptr, *Wres_feat = nullptr, *bres = nullptr, *Wcls = nullptr,
                *bcls = nullptr;

    // =========================== cross-attention inputs (crossTR) ===========================
    if (p.cross) {
        const float* qpos = vec;
        const float* kpos = qpos + J * C;
        const float* ar = p.x + ((size_t)b * J + tok) * C;
        const float* yr = p.y + ((size_t)b * J + tok) * C;
        for (int c0 = cb; c0 < cb + CW; c0 += CH) {
            float a[CH], q[CH], e[CH];
            load_row<CH, true>(ar + c0, a, valid);
            load_row<CH, true>(qpos + tok * C + c0, e, valid);
#pragma unroll
            for (int i = 0; i < CH; ++i) q[i] = a[i] + e[i];
            tmem_st<CH>(tmem + RESID + c0, a);           // residual = anchor (transfusion_head.py:164)
            store_chunks(bufA, c0, q, CH);
            load_row<CH, true>(yr + c0, q, valid);
            load_row<CH, true>(kpos + tok * C + c0, e, valid);
#pragma unroll
            for (int i = 0; i < CH; ++i) q[i] += e[i];
            store_chunks(bufV, c0, q, CH);               // bufV temporarily holds k_in as a K-major operand
        }
        vec = kpos + J * C;
    }

    // One loop over every transformer layer of the program: iteration 0 is the cross layer when there is one, the encoder's
    // input stage runs in front of its first layer.  (One instance of the layer body keeps the kernel's SASS small.)
    const int n_layers = (p.cross ? 1 : 0) + p.L;
    for (int it = 0; it < n_layers; ++it) {
        const bool is_cross = p.cross && it == 0;
        if (it == (p.cross ? 1 : 0) && p.L > 0) {
            // =========================== encoder input stage (KP_Interaction_TR) ===========================
            if (p.pre) {
                // ---- DESA fusion conv: x = relu(W_fu [desa_0 | desa_1 | desa_2 | jf] + b_fu), four accumulating K = 128 steps
                const float* bfu = vec;
                vec += C;
#pragma unroll 1
                for (int s = 0; s < 4; ++s) {
                    uint4* dst = s == 0 ? bufA : (s == 1 ? bufQ : (s == 2 ? bufK : bufV));
                    const float* src = s < 3 ? p.desa + (((size_t)b * 3 + s) * J + tok) * C : p.jf + ((size_t)b * J + tok) * C;
                    for (int c0 = cb; c0 < cb + CW; c0 += CH) {
                        float a[CH];
                        load_row<CH, true>(src + c0, a, valid);
                        store_chunks(dst, c0, a, CH);
                    }
                }
                run_gemm(g++, bufA, C, C, ACC0, false, false);
                run_gemm(g++, bufQ, C, C, ACC0, false, true);
                run_gemm(g++, bufK, C, C, ACC0, false, true);
                run_gemm(g++, bufV, C, C, ACC0, false, true);
                const float* Wrf0 = vec + (size_t)J * C + C + 48;
                for (int c0 = cb; c0 < cb + CW; c0 += CH) {
                    float a[CH], bb[CH];
                    tmem_ld_nw<CH>(tmem + ACC0 + c0, a);
                    load_row<CH, true>(bfu + c0, bb, true);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < CH; ++i) a[i] = valid ? fmaxf(a[i] + bb[i], 0.f) : 0.f;
                    head_acc<CH>(head_x, a, Wrf0 + c0, C);
                    store_chunks(bufA, c0, a, CH);
                }
            }
            pos = vec;                       // [J][128]
            bemb = pos + J * C;              // [128]
            Wres_lead = bemb + C;            // [3][16]  residual.weight columns of the leading D-128 inputs (zero padded)
            Wres_feat = Wres_lead + 48;      // [3][128] residual.weight columns of the 128 features
            bres = Wres_feat + 3 * C;        // [3] (+1 pad)
            Wcls = bres + 4;                 // [3][128]
            bcls = Wcls + 3 * C;             // [3] (+1 pad)
            if (!p.pre && !p.cross) {
                const float* xr = p.x + ((size_t)b * J + tok) * D;
                const bool al = (D & 3) == 0 && shift == 0;
                for (int c0 = cb; c0 < cb + CW; c0 += CH) {
                    float v[CH];
                    if (al)
                        load_row<CH, true>(xr + shift + c0, v, valid);
                    else
                        load_row<CH, false>(xr + shift + c0, v, valid);
                    head_acc<CH>(head_x, v, Wres_feat + c0, C);
                    store_chunks(bufA, c0, v, CH);
                }
            }
            if (shift > 0 && cg == 0) {  // leading (D - 128) inputs: joint coordinates
                const float* lead = p.cross ? p.r3d + ((size_t)b * J + tok) * shift : p.x + ((size_t)b * J + tok) * D;
                float t[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) t[i] = (valid && i < shift) ? __ldg(lead + i) : 0.f;
                head_acc<16>(head_x, t, Wres_lead, 16);
                bufAt[row] = pack8_bf16(t);
                bufAt[128 + row] = pack8_bf16(t + 8);
            }
            // ---- embedding: h = pos_emb[tok] + x W_emb^T + b_emb      (model.py:56, :88-89)
            run_gemm(g++, bufA, C, C, ACC0, shift > 0, false);
            for (int c0 = cb; c0 < cb + CW; c0 += CH) {
                float a[CH], e[CH], bb[CH];
                tmem_ld_nw<CH>(tmem + ACC0 + c0, a);
                load_row<CH, true>(pos + tok * C + c0, e, valid);
                load_row<CH, true>(bemb + c0, bb, true);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < CH; ++i) a[i] = valid ? a[i] + bb[i] + e[i] : 0.f;
                tmem_st<CH>(tmem + RESID + c0, a);
                store_chunks(bufA, c0, a, CH);
            }
            vec = bcls + 4;
            stamp();
        }

        fine_on = (it == (p.cross ? 1 : 0));
        fine();
        load_vecs(vec, 10 * C);
        fine();
        vec += 10 * C;
        const uint4* kv_src = is_cross ? bufV : bufA;     // K-major operand the K / V projections read
        const int F = is_cross ? p.Fc : p.F;
        const int act = is_cross ? 0 : 1;                 // relu | erf-gelu
        const float eps = is_cross ? 1e-5f : 1e-12f;
        // the fused encoder's residual() head accumulates over the cross layer's output (cross -> final_TR fusion);
        // its feature columns sit behind [pos | bemb | lead] of the encoder block that follows this layer's vectors
        const float* Wrf = (is_cross && p.L > 0) ? vec + (size_t)J * C + C + 48 : nullptr;

        const float *bq = sVec, *bk = sVec + C, *bv = sVec + 2 * C, *bo = sVec + 3 * C, *g1 = sVec + 4 * C, *be1 = sVec + 5 * C,
                    *b1 = sVec + 6 * C, *b2 = sVec + 7 * C, *g2 = sVec + 8 * C, *be2 = sVec + 9 * C;
        // ---- Q, K, V projections (three weight tiles); each thread drains its CW columns
        run_gemm(g++, bufA, C, C, ACC0, false, false);
        drain_kmajor(ACC0, bq, qscale, bufQ);
        run_gemm(g++, kv_src, C, C, ACC1, false, false);
        drain_kmajor(ACC1, bk, 1.f, bufK);
        run_gemm(g++, kv_src, C, C, ACC2, false, false);
        {   // V: MN-major B operand for P V ([token][dim], dim contiguous)
            float a[CW];
            tmem_ld<CW>(tmem + ACC2 + cb, a);
#pragma unroll
            for (int i = 0; i < CW; ++i) a[i] += bv[cb + i];
#pragma unroll
            for (int c = 0; c < NCH; ++c) bufV[(row >> 3) * 128 + (cb / 8 + c) * 8 + (row & 7)] = pack8_bf16(a + 8 * c);
            // P is block diagonal: zero the 12 chunks of this row outside its own 32-key block once per layer (the row's own
            // block is overwritten by every head's P); bufA is free, its last readers were the projections above
            for (int kc = cg; kc < 16; kc += CG)
                if ((kc >> 2) != wq) bufA[kc * 128 + row] = make_uint4(0, 0, 0, 0);
        }
        stamp();
        // ---- attention.  Column groups 0 / 1 own the softmax of heads pr / 2+pr of round pr (S in ACC0 / ACC1); the P V
        //      MMAs take turns on bufA.
#pragma unroll 1
        for (int pr = 0; pr < 2; ++pr) {
            sync_for_mma();
            if (warp_u == 0) {
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t idS = umma_idesc_bf16(128, 128, false, false);
                    umma_gemm(tmem0 + ACC0, smem_u32(bufQ) + pr * 4 * TS_LBO, TS_LBO, 128, smem_u32(bufK) + pr * 4 * TS_LBO, TS_LBO, 128, idS,
                              32, false);
                    umma_gemm(tmem0 + ACC1, smem_u32(bufQ) + (2 + pr) * 4 * TS_LBO, TS_LBO, 128, smem_u32(bufK) + (2 + pr) * 4 * TS_LBO,
                              TS_LBO, 128, idS, 32, false);
                    umma_commit(&mma_bar);
                }
                __syncwarp();
            }
            wait_mma();
            fine();
            // block-diagonal softmax: this row's keys are columns [32*wq, 32*wq + J) of its head's S
            float pv[32];
            if (cg < 2) {  // warp-uniform
                tmem_ld<32>(tmem + (cg ? ACC1 : ACC0) + 32 * wq, pv);
                float mx = -INFINITY;
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmaxf(mx, i < J ? pv[i] : -INFINITY);
                float sum = 0.f;
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    pv[i] = (valid && i < J) ? __expf(pv[i] - mx) : 0.f;
                    sum += pv[i];
                }
                // P stays un-normalised; O_h is scaled when it is read out
                sInv[(2 * cg + pr) * 128 + row] = valid ? 1.f / sum : 0.f;
            }
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {
                if (cg == hh) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) bufA[(4 * wq + c) * 128 + row] = pack8_bf16(pv + 8 * c);
                }
                sync_for_mma();
                if (warp_u == 0) {
                    tc_fence_after();
                    if (elect_one()) {
                        const int h = 2 * hh + pr;  // O_h[128 x 32] = P V_h, V_h = N-slice [32h, 32h+32) of the MN-major buffer
                        umma_gemm(tmem0 + ACC2 + 32 * h, smem_u32(bufA), TS_LBO, 128, smem_u32(bufV) + h * 4 * 128, 16 * 128, 128,
                                  umma_idesc_bf16(128, 32, false, true), 128, false);
                        umma_commit(&mma_bar);
                    }
                    __syncwarp();
                }
                wait_mma();  // bufA is rewritten next
                fine();
            }
        }
        stamp();
        // ---- O -> bf16 A operand, scaled by the softmax sums of the head each column belongs to
        {
            float a[CW];
            tmem_ld<CW>(tmem + ACC2 + cb, a);
            constexpr int HS = CW < 32 ? CW : 32;
#pragma unroll
            for (int s0 = 0; s0 < CW; s0 += HS) {
                const float inv = sInv[((cb + s0) >> 5) * 128 + row];
#pragma unroll
                for (int i = 0; i < HS; ++i) a[s0 + i] *= inv;
            }
#pragma unroll
            for (int c = 0; c < NCH; ++c) bufA[(cb / 8 + c) * 128 + row] = pack8_bf16(a + 8 * c);
        }
        // ---- residual + LayerNorm on a 128-wide accumulator; the CG threads of a row exchange partial statistics
        auto resid_ln = [&](uint32_t acc, const float* bias, const float* gam, const float* bet, const float* Wr) {
            float y[CW];
            float sum = 0.f, sq = 0.f;
            {
                float r[CW];
                tmem_ld_nw<CW>(tmem + acc + cb, y);
                tmem_ld_nw<CW>(tmem + RESID + cb, r);
                tmem_wait_ld();
                fine();
#pragma unroll
                for (int i = 0; i < CW; ++i) {
                    y[i] += bias[cb + i] + r[i];
                    sum += y[i];
                    sq += y[i] * y[i];
                }
            }
            float2* red = reinterpret_cast<float2*>(sRed) + red_par * (CG * 128);
            red_par ^= 1;
            red[cg * 128 + row] = make_float2(sum, sq);
            fine();
            __syncthreads();
            fine();
            sum = 0.f;
            sq = 0.f;
#pragma unroll
            for (int k = 0; k < CG; ++k) {   // same order in every thread of the row: identical statistics
                const float2 t = red[k * 128 + row];
                sum += t.x;
                sq += t.y;
            }
            const float mean = sum * (1.f / C);
            const float rstd = rsqrtf(fmaxf(sq * (1.f / C) - mean * mean, 0.f) + eps);
#pragma unroll
            for (int i = 0; i < CW; ++i) y[i] = valid ? (y[i] - mean) * rstd * gam[cb + i] + bet[cb + i] : 0.f;
            fine();
            tmem_st_nw<CW>(tmem + RESID + cb, y);
#pragma unroll
            for (int c = 0; c < NCH; ++c) bufA[(cb / 8 + c) * 128 + row] = pack8_bf16(y + 8 * c);
            if (Wr) head_acc<CW>(head_x, y, Wr + cb, C);
            tmem_wait_st();
            fine();
        };
        run_gemm(g++, bufA, C, C, ACC0, false, false);
        resid_ln(ACC0, bo, g1, be1, nullptr);
        stamp();
        // ---- FFN: hidden columns in TS_FU-wide pieces spread over the column groups
        run_gemm(g++, bufA, F, C, ACC2, false, false);
#pragma unroll 1
        for (int c0 = TS_FU * cg; c0 < F; c0 += TS_FU * CG) {
            float a[TS_FU];
            tmem_ld<TS_FU>(tmem + ACC2 + c0, a);
#pragma unroll
            for (int i = 0; i < TS_FU; ++i) {
                const float t = a[i] + b1[c0 + i];
                a[i] = act == 1 ? gelu_erf(t) : fmaxf(t, 0.f);
            }
            unsigned char* dst = reinterpret_cast<unsigned char*>(bufQ + (c0 >> 3) * 128 + row) + (c0 & 7) * 2;
            if constexpr (TS_FU == 8) {
                *reinterpret_cast<uint4*>(dst) = pack8_bf16(a);
            } else {
#pragma unroll
                for (int i = 0; i < TS_FU; i += 2) {
                    const __nv_bfloat162 t2 = __floats2bfloat162_rn(a[i], a[i + 1]);
                    *reinterpret_cast<uint32_t*>(dst + 2 * i) = *reinterpret_cast<const uint32_t*>(&t2);
                }
            }
        }
        run_gemm(g++, bufQ, C, F, ACC0, false, false);
        resid_ln(ACC0, b2, g2, be2, Wrf);
        stamp();
    }

    if (p.cross && p.L == 0) {
        for (int c0 = cb; c0 < cb + CW; c0 += CH) {
            float a[CH];
            tmem_ld<CH>(tmem + RESID + c0, a);
            if (valid) {
#pragma unroll
                for (int i = 0; i < CH; ++i) {
                    if (p.out_cj) p.out_cj[((size_t)b * C + c0 + i) * J + tok] = a[i];
                    if (p.out_jc) p.out_jc[((size_t)b * J + tok) * p.out_jc_stride + p.out_jc_c0 + c0 + i] = a[i];
                }
            }
        }
    }
    if (p.L > 0) {
        // ---- regression head: pred = cls_head(h) + residual(x)   (model.py:122-124), fp32, CG threads per row
        float pr3[3] = {head_x[0], head_x[1], head_x[2]};
        for (int c0 = cb; c0 < cb + CW; c0 += CH) {
            float a[CH];
            tmem_ld<CH>(tmem + RESID + c0, a);
            head_acc<CH>(pr3, a, Wcls + c0, C);
            if (valid && p.tokens_out) {
                float4* o = reinterpret_cast<float4*>(p.tokens_out + ((size_t)b * J + tok) * C + c0);
#pragma unroll
                for (int i = 0; i < CH / 4; ++i) o[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
            }
        }
        __syncthreads();
        float* red3 = sRed;   // [CG][128][3]
        red3[(cg * 128 + row) * 3] = pr3[0];
        red3[(cg * 128 + row) * 3 + 1] = pr3[1];
        red3[(cg * 128 + row) * 3 + 2] = pr3[2];
        __syncthreads();
        if (cg == 0 && valid && p.pred_out) {
            float o3[3] = {bres[0] + bcls[0], bres[1] + bcls[1], bres[2] + bcls[2]};
#pragma unroll
            for (int k = 0; k < CG; ++k) {
                o3[0] += red3[(k * 128 + row) * 3];
                o3[1] += red3[(k * 128 + row) * 3 + 1];
                o3[2] += red3[(k * 128 + row) * 3 + 2];
            }
            float* o = p.pred_out + ((size_t)b * J + tok) * 3;
            o[0] = o3[0];
            o[1] = o3[1];
            o[2] = o3[2];
        }
    }
    stamp();
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem0, 512);
}

constexpr size_t TS_SMEM = (size_t)(2 * TS_SLOT + 256 + 2048 + 256 + 3 * 2048) * 16 + (10 * TS_C + 4 * TS_C + 2 * TS_CG * 128 * 2) * 4;
static_assert(2 * TS_CG * 128 * 2 >= TS_CG * 128 * 3, "head reduction reuses the LayerNorm exchange buffer");

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
    KPF_REQUIRE(n_weights <= TS_MAXG);
    KPF_REQUIRE(n_weights == (cross ? 6 : 0) + (pre ? 4 : 0) + (L > 0 ? 1 + 6 * L : 0));
    KPF_REQUIRE(((uintptr_t)wmat % 16) == 0 && ((uintptr_t)wseq % 8) == 0);
    if (B == 0) return 0;
    TokParams p;
    p.x = x; p.y = y; p.r3d = r3d; p.desa = desa; p.jf = jf; p.wmat = (const uint4*)wmat; p.wseq = (const int2*)wseq; p.wvec = wvec;
    p.tokens_out = tokens_out; p.pred_out = pred_out; p.out_cj = out_cj; p.out_jc = out_jc; p.out_jc_stride = out_jc_stride;
    p.out_jc_c0 = out_jc_c0; p.B = B; p.J = J; p.D = D; p.L = L; p.F = F; p.pre = pre; p.cross = cross; p.Fc = Fc; p.G = n_weights; p.dbg = dbg;
    cudaError_t e = kpf::set_smem(token_stack_kernel, TS_SMEM);
    if (e != cudaSuccess) return (int)e;
    token_stack_kernel<<<(B + 3) / 4, TS_NT, TS_SMEM, stream>>>(p);
    KPF_CHECK_LAUNCH();
    return 0;
}

// Keypoint-token transformer stacks on tcgen05 tensor cores (bf16 operands, fp32 accumulation in TMEM).
//
//   encoder : KP_Interaction_TR.forward, model/model.py:45-126 (transformers 4.25.1 BertLayer x L):
//             h = pos_emb + x W_emb^T + b ; L x { self-attention, +res, LN, FFN(gelu), +res, LN } ; pred = cls_head(h) + residual(x)
//             optional prologue: x = relu(W_fu [desa_0 | desa_1 | desa_2 | jf] + b)   (DESA's fusion conv, model.py:160-164, :203)
//   cross   : the live TransformerDecoderLayer of updatedDecoder, model/transfusion_head.py:684-708 / :132-173:
//             Q from anchor + self_posembed, K = V from tokens + cross_posembed, +res(anchor), LN2, FFN(relu), +res, LN3
//   cross + encoder fused: crossTR followed by final_TR on cat([r3d, cross_out]) (model.py:347-349) without leaving the SM.
//
// These stacks are latency chains (tiny GEMMs, long dependent sequences), so the design minimises the chain of ONE sample and
// spreads samples over SMs: one CTA = one sample, 512 threads.
//   * The M = 128 MMA tile holds the sample's J <= 32 token rows FOUR times (rows 32r + t, r = 0..3).  A warp may only touch the
//     32 TMEM lanes of its quarter (warp % 4), so replication is what lets all 16 warps work on the same 32 tokens: thread
//     (quarter q, column group c, token t) owns columns [32q + 8c, +8) of every 128-wide epilogue = ONE 16-byte operand chunk.
//   * Attention for all four heads is two MMA pairs.  Q is written "head major" (row 32h + t = head h of token t, K = 32), K
//     likewise as the B operand, so ONE 128x128x32 MMA pair yields S_h in lanes/columns [32h, 32h+32) - the block diagonal.
//     P_h goes back to rows 32h + t (K = 32 keys), V is an MN-major [128 dims x 32 keys] B operand, and ONE 128x128x32 MMA
//     pair yields O_h in the same diagonal blocks, which are exactly the columns each quarter owns.
//   * Q and K|V (one N = 256 tile) are issued back to back; weights stream through three 32 KB slots by cp.async.bulk, each
//     transfer gated on the GEMM that last used its slot (table from ops.pack_token_program); per-layer vectors are
//     double-buffered the same way.  The fp32 residual stream stays in registers: a thread owns the same 8 columns of the same
//     token for the whole program.
// One elected lane of warp 0 issues MMAs and TMA copies from warp-uniform code (umma.cuh: elect_one).
#include "tmem_ldst.cuh"

namespace kpf {

constexpr int TS_C = 128;              // hidden size
constexpr int TS_NT = 512;             // 16 warps = 4 lane quarters x 4 column groups
constexpr int TS_SLOT = 2048;          // uint4 per weight slot (32 KB), three slots
constexpr int TS_MAXG = 64;            // weight tiles per program
constexpr int TS_NB = 8;               // weight-arrival barriers (transfer g uses g % TS_NB)
constexpr int TS_VEC = 10 * TS_C;      // floats of per-layer vectors
constexpr uint32_t TS_LBO = 128 * 16;  // K-major operand with 128 rows: bytes between 8-k groups
constexpr uint32_t ACC0 = 0, ACC1 = 128, ACC2 = 256;   // TMEM accumulator columns (ACC1..ACC2 are one N = 256 tile for K|V)

struct TokParams {
    const float* x;        // encoder input [B,J,D] (no prologue) | cross: anchor [B,J,C]
    const float* y;        // cross: tokens [B,J,C]
    const float* r3d;      // cross+encoder: leading D-128 inputs of the encoder [B,J,D-128]
    const float* desa;     // prologue: [B,3,J,C]
    const float* jf;       // prologue: [B,J,C]
    const uint4* wmat;     // bf16 canonical matrices
    const int4* wseq;      // per weight tile g: (source offset, count, slot offset) in uint4, GEMM whose completion frees the slot
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

// 8 consecutive floats of a row.  ALIGNED: 16-byte aligned (everything except the D = 131 inputs).
template <bool ALIGNED>
__device__ __forceinline__ void load8(const float* __restrict__ src, float* v, bool valid) {
    if (!valid) {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
    } else if (ALIGNED) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(src)), b = __ldg(reinterpret_cast<const float4*>(src) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __ldg(src + i);
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
    extern __shared__ __align__(128) unsigned char ts_smem[];
    uint4* wslot = reinterpret_cast<uint4*>(ts_smem);             // [3][2048]  weight slots
    uint4* bufA = wslot + 3 * TS_SLOT;                            // [16][128]  K-major A operand, rows replicated x4 (h / O / LN out)
    uint4* bufB = bufA + 2048;                                    // [16][128]  second full operand: prologue source, cross k_in, FFN hidden
    uint4* bufQ = bufB + 2048;                                    // [4][128]   Q head-major (row 32h+t, K = 32)
    uint4* bufK = bufQ + 512;                                     // [4][128]   K head-major (B operand of S)
    uint4* bufV = bufK + 512;                                     // [4][16][8] V MN-major (B operand of P V): 128 dims x 32 keys
    uint4* bufP = bufV + 512;                                     // [4][128]   P (row 32h+t, K = 32 keys)
    uint4* bufAt = bufP + 512;                                    // [2][128]   K-tail of the embedding input
    uint4* wtail = bufAt + 256;                                   // [256]      K-tail of the embedding weight
    float* sVec = reinterpret_cast<float*>(wtail + 256);          // [2][10][128] per-layer vectors (double buffered)
    float2* sRed = reinterpret_cast<float2*>(sVec + 2 * TS_VEC);  // [2][16][32] LayerNorm partials (double buffered)
    float* sSum = reinterpret_cast<float*>(sRed + 2 * 16 * 32);   // [4][128]   softmax partial sums of (column group, row)
    __shared__ __align__(8) uint64_t full[TS_NB], vec_bar[2], mma_bar, aux_bar, tail_bar;
    __shared__ uint32_t tmem_slot;
    __shared__ int4 sSeq[TS_MAXG];

    const int tid = threadIdx.x, t = tid & 31, w = tid >> 5, q = w & 3, c = w >> 2;
    const int row = 32 * q + t;   // this thread's TMEM lane = operand row of its quarter's replica
    const int col0 = 32 * q + 8 * c;  // its 8 columns of every 128-wide epilogue
    const int ck = 4 * q + c;         // = col0 / 8: its 16-byte chunk of a K-major row
    const int J = p.J, C = TS_C, G = p.G;
    const int b = blockIdx.x;
    const bool valid = t < J;
    const int warp_u = warp_index_uniform();

    pdl_launch_dependents();
    if (tid < 32) tmem_alloc(&tmem_slot, 512);
    if (tid >= 32 && tid < 32 + G) sSeq[tid - 32] = p.wseq[tid - 32];
    if (tid == 0) {
        for (int i = 0; i < TS_NB; ++i) mbar_init(&full[i], 1);
        mbar_init(&vec_bar[0], 1);
        mbar_init(&vec_bar[1], 1);
        mbar_init(&mma_bar, 1);
        mbar_init(&aux_bar, 1);
        mbar_init(&tail_bar, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = tmem_slot;
    const uint32_t tmem = tmem0 + ((uint32_t)(32 * q) << 16);  // this thread's lane window
    uint32_t mma_phase = 0, aux_phase = 0;

    // ---- layout of the fp32 vector blob (ops.pack_token_program)
    const int D = p.D, shift = p.L > 0 ? D - C : 0;
    const float* vp = p.wvec;
    const float *qpos = nullptr, *kpos = nullptr, *cross_vec = nullptr, *bfu = nullptr, *pos = nullptr, *bemb = nullptr,
                *Wres_lead = nullptr, *Wres_feat = nullptr, *bres = nullptr, *Wcls = nullptr, *bcls = nullptr, *enc_vec = nullptr;
    if (p.cross) {
        qpos = vp;
        kpos = qpos + J * C;
        cross_vec = kpos + J * C;
        vp = cross_vec + TS_VEC;
    }
    if (p.pre) {
        bfu = vp;
        vp += C;
    }
    if (p.L > 0) {
        pos = vp;                        // [J][128]
        bemb = pos + J * C;              // [128]
        Wres_lead = bemb + C;            // [3][16]  residual.weight columns of the leading D-128 inputs (zero padded)
        Wres_feat = Wres_lead + 48;      // [3][128] residual.weight columns of the 128 features
        bres = Wres_feat + 3 * C;        // [3] (+1 pad)
        Wcls = bres + 4;                 // [3][128]
        bcls = Wcls + 3 * C;             // [3] (+1 pad)
        enc_vec = bcls + 4;              // [L][10][128]
    }
    const int n_layers = (p.cross ? 1 : 0) + p.L;
    auto layer_vec = [&](int it) { return (p.cross && it == 0) ? cross_vec : enc_vec + (size_t)(it - (p.cross ? 1 : 0)) * TS_VEC; };
    // sequence index of the embedding GEMM that owns the K-tail (its tail weights follow its main part in wmat)
    const int tail_g = (p.L > 0 && D > C) ? (p.cross ? 5 : 0) + (p.pre ? 4 : 0) : -1;

    // ---- weight streaming (elected lane of warp 0)
    int nxt = 0;  // next transfer to issue
    auto load_w = [&](int gi) {
        const int4 s = sSeq[gi];
        mbar_expect_tx(&full[gi % TS_NB], (uint32_t)s.y * 16);
        tma_bulk_g2s(wslot + s.z, p.wmat + s.x, (uint32_t)s.y * 16, &full[gi % TS_NB]);
    };
    // GEMM `done` has completed: start every transfer whose slot it (or an earlier GEMM) released
    auto after_gemm = [&](int done) {
        if (warp_u == 0) {
            const bool lead = elect_one();
            while (nxt < G && sSeq[nxt].w <= done) {
                if (lead) load_w(nxt);
                ++nxt;
            }
            __syncwarp();
        }
    };
    auto load_vec = [&](int it) {  // elected lane
        mbar_expect_tx(&vec_bar[it & 1], TS_VEC * 4);
        tma_bulk_g2s(sVec + (it & 1) * TS_VEC, layer_vec(it), TS_VEC * 4, &vec_bar[it & 1]);
    };
    auto wait_w = [&](int gi) { mbar_wait(&full[gi % TS_NB], (gi / TS_NB) & 1); };
    auto wslot_of = [&](int gi) { return smem_u32(wslot + sSeq[gi].z); };
    if (warp_u == 0) {
        if (elect_one()) {
            if (n_layers > 0) load_vec(0);
            if (tail_g >= 0) {
                const int4 s = sSeq[tail_g];
                mbar_expect_tx(&tail_bar, 256 * 16);
                tma_bulk_g2s(wtail, p.wmat + s.x + s.y, 256 * 16, &tail_bar);
            }
        }
        __syncwarp();
    }
    after_gemm(-1);

    // operand writes -> async proxy, everybody's TMEM reads done, then the elected lane issues
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
    auto wait_aux = [&]() {
        mbar_wait(&aux_bar, aux_phase);
        aux_phase ^= 1;
        tc_fence_after();
    };
    // this thread's chunk of a 128-wide K-major row, written to the four row replicas
    auto store_rep = [&](uint4* buf, int chunk, const uint4 v) {
#pragma unroll
        for (int r = 0; r < 4; ++r) buf[chunk * 128 + 32 * r + t] = v;
    };

    int g = 0, n_stamp = 0, red_par = 0;
    auto stamp = [&]() {
        if (p.dbg && blockIdx.x == 0 && tid == 0 && n_stamp < 64) p.dbg[n_stamp] = clock64();
        ++n_stamp;
    };
    stamp();
    pdl_wait();   // everything above touched only weights; the activations below come from the previous kernel
    float head_x[3] = {0.f, 0.f, 0.f};  // this thread's share of residual(x) of the regression head (fp32)
    float resid[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // fp32 residual stream: this thread's 8 columns of its token, in
                                                               // registers for the whole program (a thread always owns the same ones)
    const float qscale = rsqrtf((float)(C / 4));  // head_dim^-0.5, 4 heads

    // =========================== cross-attention inputs (crossTR) ===========================
    if (p.cross) {
        float a[8], e[8], kin[8];
        load8<true>(p.x + ((size_t)b * J + t) * C + col0, a, valid);
        load8<true>(qpos + t * C + col0, e, valid);
        load8<true>(p.y + ((size_t)b * J + t) * C + col0, kin, valid);
#pragma unroll
        for (int i = 0; i < 8; ++i) resid[i] = a[i];     // residual = anchor (transfusion_head.py:164)
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] += a[i];
        store_rep(bufA, ck, pack8_bf16(e));              // q_in = anchor + self_posembed
        load8<true>(kpos + t * C + col0, e, valid);
#pragma unroll
        for (int i = 0; i < 8; ++i) kin[i] += e[i];
        store_rep(bufB, ck, pack8_bf16(kin));            // k_in = tokens + cross_posembed
    }

    // One loop over every transformer layer of the program: iteration 0 is the cross layer when there is one; the encoder's
    // input stage runs in front of its first layer.
    for (int it = 0; it < n_layers; ++it) {
        const bool is_cross = p.cross && it == 0;
        if (it == (p.cross ? 1 : 0) && p.L > 0) {
            // =========================== encoder input stage (KP_Interaction_TR) ===========================
            if (p.pre) {
                // ---- DESA fusion conv: x = relu(W_fu [desa_0 | desa_1 | desa_2 | jf] + b_fu): four accumulating K = 128 GEMMs,
                //      two operand buffers, two in flight
                float v0[8], v1[8], v2[8], v3[8];
                const float* dsrc = p.desa + ((size_t)b * 3 * J + t) * C + col0;
                load8<true>(dsrc, v0, valid);
                load8<true>(dsrc + (size_t)J * C, v1, valid);
                load8<true>(dsrc + (size_t)2 * J * C, v2, valid);
                load8<true>(p.jf + ((size_t)b * J + t) * C + col0, v3, valid);
                store_rep(bufA, ck, pack8_bf16(v0));
                store_rep(bufB, ck, pack8_bf16(v1));
                sync_for_mma();
                if (warp_u == 0) {
                    tc_fence_after();
                    wait_w(g);
                    wait_w(g + 1);
                    if (elect_one()) {
                        const uint32_t id = umma_idesc_bf16(128, 128, false, false);
                        umma_gemm(tmem0 + ACC0, smem_u32(bufA), TS_LBO, 128, wslot_of(g), TS_LBO, 128, id, C, false);
                        umma_commit(&aux_bar);
                        umma_gemm(tmem0 + ACC0, smem_u32(bufB), TS_LBO, 128, wslot_of(g + 1), TS_LBO, 128, id, C, true);
                        umma_commit(&mma_bar);
                    }
                    __syncwarp();
                }
                wait_aux();
                after_gemm(g);
                store_rep(bufA, ck, pack8_bf16(v2));
                wait_mma();
                after_gemm(g + 1);
                store_rep(bufB, ck, pack8_bf16(v3));
                sync_for_mma();
                if (warp_u == 0) {
                    tc_fence_after();
                    wait_w(g + 2);
                    wait_w(g + 3);
                    if (elect_one()) {
                        const uint32_t id = umma_idesc_bf16(128, 128, false, false);
                        umma_gemm(tmem0 + ACC0, smem_u32(bufA), TS_LBO, 128, wslot_of(g + 2), TS_LBO, 128, id, C, true);
                        umma_gemm(tmem0 + ACC0, smem_u32(bufB), TS_LBO, 128, wslot_of(g + 3), TS_LBO, 128, id, C, true);
                        umma_commit(&mma_bar);
                    }
                    __syncwarp();
                }
                float bb[8];
                load8<true>(bfu + col0, bb, true);
                wait_mma();
                after_gemm(g + 3);
                g += 4;
                float a[8];
                tmem_ld<8>(tmem + ACC0 + col0, a);
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = valid ? fmaxf(a[i] + bb[i], 0.f) : 0.f;
                head_acc<8>(head_x, a, Wres_feat + col0, C);
                store_rep(bufA, ck, pack8_bf16(a));
            }
            if (!p.pre && !p.cross) {
                const float* xr = p.x + ((size_t)b * J + t) * D + shift + col0;
                float v[8];
                if ((D & 3) == 0 && shift == 0)
                    load8<true>(xr, v, valid);
                else
                    load8<false>(xr, v, valid);
                head_acc<8>(head_x, v, Wres_feat + col0, C);
                store_rep(bufA, ck, pack8_bf16(v));
            }
            if (shift > 0 && ck == 0) {  // leading (D - 128) inputs: joint coordinates (one thread per token)
                const float* lead = p.cross ? p.r3d + ((size_t)b * J + t) * shift : p.x + ((size_t)b * J + t) * D;
                float tl[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) tl[i] = (valid && i < shift) ? __ldg(lead + i) : 0.f;
                head_acc<16>(head_x, tl, Wres_lead, 16);
                const uint4 lo = pack8_bf16(tl), hi = pack8_bf16(tl + 8);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    bufAt[32 * r + t] = lo;
                    bufAt[128 + 32 * r + t] = hi;
                }
            }
            // ---- embedding: h = pos_emb[tok] + x W_emb^T + b_emb      (model.py:56, :88-89)
            sync_for_mma();
            if (warp_u == 0) {
                tc_fence_after();
                wait_w(g);
                if (shift > 0) mbar_wait(&tail_bar, 0);
                if (elect_one()) {
                    const uint32_t id = umma_idesc_bf16(128, 128, false, false);
                    umma_gemm(tmem0 + ACC0, smem_u32(bufA), TS_LBO, 128, wslot_of(g), TS_LBO, 128, id, C, false);
                    if (shift > 0) umma_gemm(tmem0 + ACC0, smem_u32(bufAt), TS_LBO, 128, smem_u32(wtail), TS_LBO, 128, id, 16, true);
                    umma_commit(&mma_bar);
                }
                __syncwarp();
            }
            float e[8], bb[8];
            load8<true>(pos + t * C + col0, e, valid);
            load8<true>(bemb + col0, bb, true);
            wait_mma();
            after_gemm(g);
            ++g;
            float a[8];
            tmem_ld<8>(tmem + ACC0 + col0, a);
#pragma unroll
            for (int i = 0; i < 8; ++i) resid[i] = a[i] = valid ? a[i] + bb[i] + e[i] : 0.f;
            store_rep(bufA, ck, pack8_bf16(a));
            stamp();
        }

        const float* sv = sVec + (it & 1) * TS_VEC;
        const float *bq = sv, *bk = sv + C, *bv = sv + 2 * C, *bo = sv + 3 * C, *g1 = sv + 4 * C, *be1 = sv + 5 * C, *b1 = sv + 6 * C,
                    *b2 = sv + 7 * C, *g2 = sv + 8 * C, *be2 = sv + 9 * C;
        const uint4* kv_src = is_cross ? bufB : bufA;     // K-major operand the K / V projections read
        const int F = is_cross ? p.Fc : p.F;
        const int act = is_cross ? 0 : 1;                 // relu | erf-gelu
        const float eps = is_cross ? 1e-5f : 1e-12f;
        // the fused encoder's residual() head accumulates over the cross layer's output (cross -> final_TR fusion)
        const float* Wrf = (is_cross && p.L > 0) ? Wres_feat : nullptr;

        // ---- Q and K|V projections, issued back to back (weights g, g+1 = one N = 256 tile)
        sync_for_mma();
        if (warp_u == 0) {
            tc_fence_after();
            wait_w(g);
            wait_w(g + 1);
            if (elect_one()) {
                if (it + 1 < n_layers) load_vec(it + 1);   // the other vector buffer was last read before the barrier above
                umma_gemm(tmem0 + ACC0, smem_u32(bufA), TS_LBO, 128, wslot_of(g), TS_LBO, 128, umma_idesc_bf16(128, 128, false, false), C,
                          false);
                umma_commit(&aux_bar);
                umma_gemm(tmem0 + ACC1, smem_u32(kv_src), TS_LBO, 128, wslot_of(g + 1), 256 * 16, 128, umma_idesc_bf16(128, 256, false, false),
                          C, false);
                umma_commit(&mma_bar);
            }
            __syncwarp();
        }
        mbar_wait(&vec_bar[it & 1], (it >> 1) & 1);       // this layer's vectors have landed
        wait_aux();
        after_gemm(g);
        {   // Q -> head-major A operand of S: row 32q+t = head q of token t, chunk c = its dims [8c, 8c+8)
            float a[8];
            tmem_ld<8>(tmem + ACC0 + col0, a);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = (a[i] + bq[col0 + i]) * qscale;
            bufQ[c * 128 + row] = pack8_bf16(a);
        }
        wait_mma();
        after_gemm(g + 1);
        float kv[16];
        tmem_ld_nw<8>(tmem + ACC1 + col0, kv);
        tmem_ld_nw<8>(tmem + ACC2 + col0, kv + 8);
        tmem_wait_ld();
        {   // K likewise, as the B operand of S
#pragma unroll
            for (int i = 0; i < 8; ++i) kv[i] += bk[col0 + i];
            bufK[c * 128 + row] = pack8_bf16(kv);
        }
        stamp();
        // ---- S for all four heads: D[32h+t][32h'+k] = Q_h[t] . K_h'[k]; the diagonal blocks h = h' are the scores
        sync_for_mma();
        if (warp_u == 0) {
            tc_fence_after();
            if (elect_one()) {
                umma_gemm(tmem0 + ACC0, smem_u32(bufQ), TS_LBO, 128, smem_u32(bufK), TS_LBO, 128, umma_idesc_bf16(128, 128, false, false), 32,
                          false);
                umma_commit(&mma_bar);
            }
            __syncwarp();
        }
        {   // V -> MN-major B operand of P V (dims contiguous): key t, dims [32q+8c, +8); overlaps the S MMAs
#pragma unroll
            for (int i = 0; i < 8; ++i) kv[8 + i] += bv[col0 + i];
            bufV[(t >> 3) * 128 + ck * 8 + (t & 7)] = pack8_bf16(kv + 8);
        }
        wait_mma();
        {   // softmax of row t of head q: every column group takes the row maximum over all keys and exponentiates its 8 keys;
            // P stays un-normalised, the partial sums meet when O is read out
            float sa[32], own[8];
            tmem_ld_nw<32>(tmem + ACC0 + 32 * q, sa);
            tmem_ld_nw<8>(tmem + ACC0 + col0, own);
            tmem_wait_ld();
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; ++i) mx = fmaxf(mx, i < J ? sa[i] : -INFINITY);
            float psum = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                own[i] = (valid && 8 * c + i < J) ? __expf(own[i] - mx) : 0.f;
                psum += own[i];
            }
            sSum[c * 128 + row] = psum;
            bufP[c * 128 + row] = pack8_bf16(own);
        }
        // ---- O for all four heads: D[32h+t][n] = P_h[t] . V[:, n]; columns [32h, 32h+32) are head h's output
        sync_for_mma();
        if (warp_u == 0) {
            tc_fence_after();
            if (elect_one()) {
                umma_gemm(tmem0 + ACC1, smem_u32(bufP), TS_LBO, 128, smem_u32(bufV), 16 * 128, 128, umma_idesc_bf16(128, 128, false, true), 32,
                          false);
                umma_commit(&mma_bar);
            }
            __syncwarp();
        }
        float inv;
        {
            const float ssum = sSum[row] + sSum[128 + row] + sSum[256 + row] + sSum[384 + row];
            inv = valid ? 1.f / ssum : 0.f;
        }
        wait_mma();
        stamp();
        {   // O -> replicated A operand of the output projection
            float a[8];
            tmem_ld<8>(tmem + ACC1 + col0, a);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] *= inv;
            store_rep(bufA, ck, pack8_bf16(a));
        }
        // ---- residual + LayerNorm on a 128-wide accumulator; the 16 threads of a token exchange partial statistics
        auto resid_ln = [&](uint32_t acc, const float* bias, const float* gam, const float* bet, const float* Wr) {
            float y[8];
            tmem_ld<8>(tmem + acc + col0, y);
            float sum = 0.f, sq = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                y[i] += bias[col0 + i] + resid[i];
                sum += y[i];
                sq += y[i] * y[i];
            }
            float2* red = sRed + red_par * (16 * 32);
            red_par ^= 1;
            red[ck * 32 + t] = make_float2(sum, sq);
            __syncthreads();
            sum = 0.f;
            sq = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) {   // same order in every thread of the token: identical statistics in all replicas
                const float2 u = red[k * 32 + t];
                sum += u.x;
                sq += u.y;
            }
            const float mean = sum * (1.f / C);
            const float rstd = rsqrtf(fmaxf(sq * (1.f / C) - mean * mean, 0.f) + eps);
#pragma unroll
            for (int i = 0; i < 8; ++i) resid[i] = y[i] = valid ? (y[i] - mean) * rstd * gam[col0 + i] + bet[col0 + i] : 0.f;
            store_rep(bufA, ck, pack8_bf16(y));
            if (Wr) head_acc<8>(head_x, y, Wr + col0, C);
        };
        sync_for_mma();
        if (warp_u == 0) {
            tc_fence_after();
            wait_w(g + 2);
            if (elect_one()) {
                umma_gemm(tmem0 + ACC0, smem_u32(bufA), TS_LBO, 128, wslot_of(g + 2), TS_LBO, 128, umma_idesc_bf16(128, 128, false, false), C,
                          false);
                umma_commit(&mma_bar);
            }
            __syncwarp();
        }
        wait_mma();
        after_gemm(g + 2);
        resid_ln(ACC0, bo, g1, be1, nullptr);
        stamp();
        // ---- FFN: hidden column(s) ck * F/16 ... of the token, written to the four row replicas of the hidden operand
        sync_for_mma();
        if (warp_u == 0) {
            tc_fence_after();
            wait_w(g + 3);
            if (elect_one()) {
                umma_gemm(tmem0 + ACC2, smem_u32(bufA), TS_LBO, 128, wslot_of(g + 3), (uint32_t)F * 16, 128, umma_idesc_bf16(128, F, false, false),
                          C, false);
                umma_commit(&mma_bar);
            }
            __syncwarp();
        }
        wait_mma();
        after_gemm(g + 3);
        {
            const int fp = F >> 4;
#pragma unroll 1
            for (int i = 0; i < fp; ++i) {
                const int col = ck * fp + i;
                float a1[2];
                tmem_ld<1>(tmem + ACC2 + col, a1);
                const float u = a1[0] + b1[col];
                const __nv_bfloat16 hv = __float2bfloat16(act == 1 ? gelu_erf(u) : fmaxf(u, 0.f));
#pragma unroll
                for (int r = 0; r < 4; ++r) reinterpret_cast<__nv_bfloat16*>(bufB + (col >> 3) * 128 + 32 * r + t)[col & 7] = hv;
            }
        }
        sync_for_mma();
        if (warp_u == 0) {
            tc_fence_after();
            wait_w(g + 4);
            if (elect_one()) {
                umma_gemm(tmem0 + ACC0, smem_u32(bufB), TS_LBO, 128, wslot_of(g + 4), TS_LBO, 128, umma_idesc_bf16(128, 128, false, false), F,
                          false);
                umma_commit(&mma_bar);
            }
            __syncwarp();
        }
        wait_mma();
        after_gemm(g + 4);
        resid_ln(ACC0, b2, g2, be2, Wrf);
        g += 5;
        stamp();
    }

    if (p.cross && p.L == 0) {
        const float* a = resid;
        if (valid) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (p.out_cj) p.out_cj[((size_t)b * C + col0 + i) * J + t] = a[i];
                if (p.out_jc) p.out_jc[((size_t)b * J + t) * p.out_jc_stride + p.out_jc_c0 + col0 + i] = a[i];
            }
        }
    }
    if (p.L > 0) {
        // ---- regression head: pred = cls_head(h) + residual(x)   (model.py:122-124), fp32, 16 threads per token
        float pr3[3] = {head_x[0], head_x[1], head_x[2]};
        const float* a = resid;
        head_acc<8>(pr3, a, Wcls + col0, C);
        if (valid && p.tokens_out) {
            float4* o = reinterpret_cast<float4*>(p.tokens_out + ((size_t)b * J + t) * C + col0);
            o[0] = make_float4(a[0], a[1], a[2], a[3]);
            o[1] = make_float4(a[4], a[5], a[6], a[7]);
        }
        __syncthreads();
        float* red3 = reinterpret_cast<float*>(sRed);   // [16][32][3]
        red3[(ck * 32 + t) * 3] = pr3[0];
        red3[(ck * 32 + t) * 3 + 1] = pr3[1];
        red3[(ck * 32 + t) * 3 + 2] = pr3[2];
        __syncthreads();
        if (ck == 0 && valid && p.pred_out) {
            float o3[3] = {bres[0] + bcls[0], bres[1] + bcls[1], bres[2] + bcls[2]};
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                o3[0] += red3[(k * 32 + t) * 3];
                o3[1] += red3[(k * 32 + t) * 3 + 1];
                o3[2] += red3[(k * 32 + t) * 3 + 2];
            }
            float* o = p.pred_out + ((size_t)b * J + t) * 3;
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

constexpr size_t TS_SMEM = (size_t)(3 * TS_SLOT + 2 * 2048 + 4 * 512 + 2 * 256) * 16 + (size_t)(2 * TS_VEC) * 4 + 2 * 16 * 32 * 8 + 4 * 128 * 4;
static_assert(2 * 16 * 32 * 8 >= 16 * 32 * 3 * 4, "head reduction reuses the LayerNorm exchange buffer");

}  // namespace kpf

extern "C" int kpf_token_stack(const float* x, const float* y, const float* r3d, const float* desa, const float* jf, const void* wmat,
                               const void* wseq, const float* wvec, int n_weights, int cross, int pre, int B, int J, int D, int L, int F,
                               int Fc, float* tokens_out, float* pred_out, float* out_cj, float* out_jc, int out_jc_stride, int out_jc_c0,
                               long long* dbg, cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(B >= 0 && J >= 1 && J <= 32 && L >= 0 && (cross || L > 0));
    if (B == 0) return 0;   // an empty batch has no buffers to validate
    KPF_REQUIRE(L == 0 || F == 16 || F == 32 || F == 64 || F == 128);
    KPF_REQUIRE(!cross || (y != nullptr && (Fc == 16 || Fc == 32 || Fc == 64 || Fc == 128)));
    KPF_REQUIRE(L == 0 || D == TS_C || (D > TS_C && D <= TS_C + 16));
    KPF_REQUIRE(!pre || (desa != nullptr && jf != nullptr && !cross && D == TS_C));
    KPF_REQUIRE(!(cross && L > 0) || (r3d != nullptr && D > TS_C));
    KPF_REQUIRE(n_weights <= TS_MAXG);
    KPF_REQUIRE(n_weights == (cross ? 5 : 0) + (pre ? 4 : 0) + (L > 0 ? 1 + 5 * L : 0));
    KPF_REQUIRE(((uintptr_t)wmat % 16) == 0 && ((uintptr_t)wseq % 16) == 0);
    TokParams p;
    p.x = x; p.y = y; p.r3d = r3d; p.desa = desa; p.jf = jf; p.wmat = (const uint4*)wmat; p.wseq = (const int4*)wseq; p.wvec = wvec;
    p.tokens_out = tokens_out; p.pred_out = pred_out; p.out_cj = out_cj; p.out_jc = out_jc; p.out_jc_stride = out_jc_stride;
    p.out_jc_c0 = out_jc_c0; p.B = B; p.J = J; p.D = D; p.L = L; p.F = F; p.pre = pre; p.cross = cross; p.Fc = Fc; p.G = n_weights; p.dbg = dbg;
    cudaError_t e = kpf::set_smem(token_stack_kernel, TS_SMEM);
    if (e != cudaSuccess) return (int)e;
    e = kpf::launch_pdl(token_stack_kernel, dim3(B), dim3(TS_NT), TS_SMEM, stream, p);
    if (e != cudaSuccess) return (int)e;
    KPF_CHECK_LAUNCH();
    return 0;
}

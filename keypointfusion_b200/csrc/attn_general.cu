// General-shape attention building blocks (fp32, CUDA cores) for the parts of model/transfusion_head.py that KPFusion itself never
// runs but a user of the reference can import: MultiheadAttention.forward for any (L, S) with masks and averaged
// weights (:176-300, :303-556), TransformerDecoderLayer with self-attention (:94-173), detrDecoder (:560-632: 21 queries x H*W keys),
// spatial_aggregate_TR (:711-783: H*W queries x 21 keys), PositionEmbeddingLearned (:16-32), DetrSinePositionEmbedding (:57-91).
// The live 21 x 21 decoder layer has its own fused kernels (cross_attn.cu, token_stack.cu); these are the general ones:
//   linear_rows_kernel   Y[(b,p)][o] = act(((X[(b,p)][:] + pos[(b,p)][:]) . W[o][:] + bias[o]) * scale), every operand strided
//   mha_core_kernel      softmax(Q K^T + masks) V per head, online softmax over 32-key chunks staged in shared memory
//   mha_weights_kernel   the head-averaged attention weights the reference returns with need_weights=True (:551-554)
//   add_layernorm_kernel LayerNorm(x + r) with a strided (e.g. channel-major [B,C,P], :172) output
//   sine_posembed_kernel DetrSinePositionEmbedding.forward
#include "common.cuh"

namespace kpf {

__device__ __forceinline__ float ag_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float ag_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------------ rows GEMM
constexpr int LR_T = 64, LR_K = 16, LR_LD = LR_T + 4;

struct LinearRows {
    const float* X;  long long x_bs, x_ps, x_ks;       // element (b, p, k) at X[b*x_bs + p*x_ps + k*x_ks]
    const float* pos; long long pos_bs, pos_ps, pos_ks; // optional term added to X before the product (position embedding)
    const long long* pos_idx;                           // optional [B*P]: row of `pos` to add (nn.Embedding lookup), instead of (b, p)
    const float* W;                                     // [O][K] (nn.Linear.weight)
    const float* bias;                                  // [O] or null
    float* Y; long long y_bs, y_ps, y_os;
    int B, P, K, O, relu;
    float scale;
};

__global__ void __launch_bounds__(256) linear_rows_kernel(const LinearRows a) {
    __shared__ __align__(16) float As[LR_K][LR_LD], Bs[LR_K][LR_LD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long long R = (long long)a.B * a.P;
    const long long row0 = (long long)blockIdx.x * LR_T;
    const int col0 = blockIdx.y * LR_T;
    // loader mapping: thread -> (row / out column lr, four consecutive k)
    const int lr = tid >> 2, lk = (tid & 3) * 4;
    const long long r_ld = row0 + lr;
    const bool r_ok = r_ld < R;
    const int bb = r_ok ? (int)(r_ld / a.P) : 0, pp = r_ok ? (int)(r_ld - (long long)bb * a.P) : 0;
    const float* xrow = a.X + (long long)bb * a.x_bs + (long long)pp * a.x_ps;
    const float* prow = nullptr;
    if (a.pos && r_ok)
        prow = a.pos_idx ? a.pos + a.pos_idx[r_ld] * a.pos_ps : a.pos + (long long)bb * a.pos_bs + (long long)pp * a.pos_ps;
    const int o_ld = col0 + lr;
    const float* wrow = a.W + (long long)o_ld * a.K;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < a.K; k0 += LR_K) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = k0 + lk + i;
            float xv = 0.f, wv = 0.f;
            if (k < a.K) {
                if (r_ok) {
                    xv = __ldg(xrow + (long long)k * a.x_ks);
                    if (prow) xv += __ldg(prow + (long long)k * a.pos_ks);
                }
                if (o_ld < a.O) wv = __ldg(wrow + k);
            }
            As[lk + i][lr] = xv;
            Bs[lk + i][lr] = wv;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < LR_K; ++k) {
            const float4 av = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 bv = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] += ar[i] * br[j];
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long r = row0 + ty * 4 + i;
        if (r >= R) continue;
        const int b = (int)(r / a.P), p = (int)(r - (long long)b * a.P);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int o = col0 + tx * 4 + j;
            if (o >= a.O) continue;
            float v = (acc[i][j] + (a.bias ? __ldg(a.bias + o) : 0.f)) * a.scale;
            if (a.relu) v = fmaxf(v, 0.f);
            a.Y[(long long)b * a.y_bs + (long long)p * a.y_ps + (long long)o * a.y_os] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------ attention core
constexpr int MH_QT = 32, MH_KT = 32, MH_HD = 64;   // queries per CTA (8 warps x 4), keys per chunk, maximum head dimension

struct MhaCore {
    const float *Q, *K, *V;              // element (b, p, c) at base[b*bs + p*ps + c]; Q already scaled (:468)
    long long q_bs, q_ps, k_bs, k_ps, v_bs, v_ps;
    const float* attn_mask;              // additive [Pq][Pk] or null (:527-529)
    const unsigned char* key_pad;        // [B][Pk], non-zero = masked with -inf (:531-537), or null
    float* O; long long o_bs, o_ps;      // [B][Pq][C]
    float* stats;                        // optional [B][Pq][H][2] = (row maximum, row sum) for mha_weights_kernel
    int B, Pq, Pk, C, H;
};

__global__ void __launch_bounds__(256) mha_core_kernel(const MhaCore a) {
    __shared__ float Qs[MH_QT][MH_HD], Ks[MH_KT][MH_HD + 1], Vs[MH_KT][MH_HD];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y, q0 = blockIdx.x * MH_QT, hd = a.C / a.H;
    for (int h = 0; h < a.H; ++h) {
        __syncthreads();   // the previous head's Qs / Ks / Vs readers are done
        for (int i = tid; i < MH_QT * hd; i += 256) {
            const int qi = i / hd, d = i - qi * hd;
            Qs[qi][d] = q0 + qi < a.Pq ? __ldg(a.Q + (long long)b * a.q_bs + (long long)(q0 + qi) * a.q_ps + h * hd + d) : 0.f;
        }
        float m[4], l[4], acc[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            m[i] = -INFINITY;
            l[i] = 0.f;
            acc[i][0] = acc[i][1] = 0.f;
        }
        for (int k0 = 0; k0 < a.Pk; k0 += MH_KT) {
            __syncthreads();   // Qs written (first chunk) / the previous chunk's readers are done
            for (int i = tid; i < MH_KT * hd; i += 256) {
                const int kk = i / hd, d = i - kk * hd;
                const bool ok = k0 + kk < a.Pk;
                Ks[kk][d] = ok ? __ldg(a.K + (long long)b * a.k_bs + (long long)(k0 + kk) * a.k_ps + h * hd + d) : 0.f;
                Vs[kk][d] = ok ? __ldg(a.V + (long long)b * a.v_bs + (long long)(k0 + kk) * a.v_ps + h * hd + d) : 0.f;
            }
            __syncthreads();
            const int key = k0 + lane;
            const bool key_ok = key < a.Pk && !(a.key_pad && a.key_pad[(long long)b * a.Pk + key]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int qi = warp * 4 + i;
                float s = 0.f;
                for (int d = 0; d < hd; ++d) s += Qs[qi][d] * Ks[lane][d];
                if (a.attn_mask && key < a.Pk && q0 + qi < a.Pq) s += __ldg(a.attn_mask + (long long)(q0 + qi) * a.Pk + key);
                if (!key_ok) s = -INFINITY;
                const float m_new = fmaxf(m[i], ag_warp_max(s));
                float p = 0.f, corr = 1.f;
                if (m_new != -INFINITY) {   // otherwise every key so far is masked: nothing to add, nothing to rescale
                    p = expf(s - m_new);
                    corr = expf(m[i] - m_new);
                }
                m[i] = m_new;
                l[i] = l[i] * corr + ag_warp_sum(p);
                acc[i][0] *= corr;
                acc[i][1] *= corr;
#pragma unroll 8
                for (int kk = 0; kk < MH_KT; ++kk) {
                    const float pk = __shfl_sync(0xffffffffu, p, kk);
                    acc[i][0] += pk * Vs[kk][lane];
                    if (hd > 32) acc[i][1] += pk * Vs[kk][lane + 32 < MH_HD ? lane + 32 : lane];
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int q = q0 + warp * 4 + i;
            if (q >= a.Pq) continue;
            float* o = a.O + (long long)b * a.o_bs + (long long)q * a.o_ps + h * hd;
            if (lane < hd) o[lane] = acc[i][0] / l[i];            // a fully masked row is 0 / 0 = NaN, like softmax over -inf
            if (lane + 32 < hd) o[lane + 32] = acc[i][1] / l[i];
            if (a.stats && lane == 0) {
                float* st = a.stats + (((long long)b * a.Pq + q) * a.H + h) * 2;
                st[0] = m[i];
                st[1] = l[i];
            }
        }
    }
}

// w[b][q][k] = (1 / H) sum_h exp(s_h[q][k] - max_h[q]) / sum_h[q]     (transfusion_head.py:551-554)
__global__ void __launch_bounds__(128) mha_weights_kernel(const MhaCore a, float* __restrict__ w) {
    extern __shared__ float wq[];   // [C] the query row, then [H][2] its statistics
    const int q = blockIdx.x, b = blockIdx.y, hd = a.C / a.H;
    for (int i = threadIdx.x; i < a.C; i += blockDim.x) wq[i] = __ldg(a.Q + (long long)b * a.q_bs + (long long)q * a.q_ps + i);
    for (int i = threadIdx.x; i < 2 * a.H; i += blockDim.x) wq[a.C + i] = a.stats[((long long)b * a.Pq + q) * a.H * 2 + i];
    __syncthreads();
    for (int key = threadIdx.x; key < a.Pk; key += blockDim.x) {
        const float* kr = a.K + (long long)b * a.k_bs + (long long)key * a.k_ps;
        const bool masked = a.key_pad && a.key_pad[(long long)b * a.Pk + key];
        const float am = a.attn_mask ? __ldg(a.attn_mask + (long long)q * a.Pk + key) : 0.f;
        float tot = 0.f;
        for (int h = 0; h < a.H; ++h) {
            float s = 0.f;
            for (int d = 0; d < hd; ++d) s += wq[h * hd + d] * __ldg(kr + h * hd + d);
            s += am;
            tot += masked ? 0.f / wq[a.C + 2 * h + 1] : expf(s - wq[a.C + 2 * h]) / wq[a.C + 2 * h + 1];
        }
        w[((long long)b * a.Pq + q) * a.Pk + key] = tot / (float)a.H;
    }
}

// ------------------------------------------------------------------------------------------------ residual + LayerNorm
constexpr int LN_ROWS = 32;
struct AddLayerNorm {
    const float* x; long long x_bs, x_ps, x_cs;   // element (b, p, c) at x[b*x_bs + p*x_ps + c*x_cs]
    const float* r;                               // [B*P][C] contiguous, or null
    const float *gamma, *beta;
    float* y; long long y_bs, y_ps, y_cs;
    int B, P, C;
    float eps;
};

// 32 rows per CTA through a shared-memory tile, so that a channel-major operand (x_cs or y_cs != 1: a [B,C,H*W] feature map read as
// [B,H*W,C] rows, or the reference's [B,C,P] result) is accessed with the lanes running over consecutive rows
__global__ void __launch_bounds__(256) add_layernorm_kernel(const AddLayerNorm a) {
    extern __shared__ float tile[];   // [LN_ROWS][C + 1]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, ld = a.C + 1;
    const long long R = (long long)a.B * a.P, row0 = (long long)blockIdx.x * LN_ROWS;
    if (a.x_cs == 1) {
        for (int i = 0; i < LN_ROWS / 8; ++i) {
            const int lr = warp * (LN_ROWS / 8) + i;
            const long long row = row0 + lr;
            if (row >= R) continue;   // warp-uniform
            const int b = (int)(row / a.P), p = (int)(row - (long long)b * a.P);
            const float* xr = a.x + (long long)b * a.x_bs + (long long)p * a.x_ps;
            for (int c = lane; c < a.C; c += 32) tile[lr * ld + c] = __ldg(xr + c);
        }
    } else {
        const long long row = row0 + lane;
        if (row < R) {
            const int b = (int)(row / a.P), p = (int)(row - (long long)b * a.P);
            const float* xr = a.x + (long long)b * a.x_bs + (long long)p * a.x_ps;
            for (int c = warp; c < a.C; c += 8) tile[lane * ld + c] = __ldg(xr + (long long)c * a.x_cs);
        }
    }
    __syncthreads();
    for (int i = 0; i < LN_ROWS / 8; ++i) {
        const int lr = warp * (LN_ROWS / 8) + i;
        const long long row = row0 + lr;
        if (row >= R) continue;
        float s = 0.f;
        for (int c = lane; c < a.C; c += 32) {
            const float v = tile[lr * ld + c] + (a.r ? __ldg(a.r + row * a.C + c) : 0.f);
            tile[lr * ld + c] = v;
            s += v;
        }
        const float mean = ag_warp_sum(s) / (float)a.C;
        float q = 0.f;
        for (int c = lane; c < a.C; c += 32) {
            const float d = tile[lr * ld + c] - mean;
            q += d * d;
        }
        const float rstd = rsqrtf(ag_warp_sum(q) / (float)a.C + a.eps);
        for (int c = lane; c < a.C; c += 32)
            tile[lr * ld + c] = (tile[lr * ld + c] - mean) * rstd * __ldg(a.gamma + c) + __ldg(a.beta + c);
    }
    __syncthreads();
    if (a.y_cs == 1) {
        for (int i = 0; i < LN_ROWS / 8; ++i) {
            const int lr = warp * (LN_ROWS / 8) + i;
            const long long row = row0 + lr;
            if (row >= R) continue;
            const int b = (int)(row / a.P), p = (int)(row - (long long)b * a.P);
            for (int c = lane; c < a.C; c += 32) a.y[(long long)b * a.y_bs + (long long)p * a.y_ps + c] = tile[lr * ld + c];
        }
    } else {
        const long long row = row0 + lane;
        if (row < R) {
            const int b = (int)(row / a.P), p = (int)(row - (long long)b * a.P);
            for (int c = warp; c < a.C; c += 8)
                a.y[(long long)b * a.y_bs + (long long)p * a.y_ps + (long long)c * a.y_cs] = tile[lane * ld + c];
        }
    }
}

// ------------------------------------------------------------------------------------------------ DetrSinePositionEmbedding
// mask [B][Hh][Ww] f32 (null = all ones), dim_t [D] = temperature ** (2 * (i // 2) / D) -> out [B][2D][Hh][Ww]   (:75-91)
__global__ void __launch_bounds__(256) sine_posembed_kernel(const float* __restrict__ mask, const float* __restrict__ dim_t, int B, int Hh,
                                                            int Ww, int D, int normalize, float scale, float* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * Hh * Ww) return;
    const int x = (int)(i % Ww), y = (int)((i / Ww) % Hh), b = (int)(i / ((long long)Ww * Hh));
    float ye = 0.f, xe = 0.f, ytot = 0.f, xtot = 0.f;
    for (int r = 0; r < Hh; ++r) {   // cumulative sums along the rows / the columns (:78-79)
        const float v = mask ? __ldg(mask + ((long long)b * Hh + r) * Ww + x) : 1.f;
        ytot += v;
        if (r == y) ye = ytot;
    }
    for (int c = 0; c < Ww; ++c) {
        const float v = mask ? __ldg(mask + ((long long)b * Hh + y) * Ww + c) : 1.f;
        xtot += v;
        if (c == x) xe = xtot;
    }
    if (normalize) {   // :80-82
        ye = ye / (ytot + 1e-6f) * scale;
        xe = xe / (xtot + 1e-6f) * scale;
    }
    const long long plane = (long long)Hh * Ww, o0 = (long long)b * 2 * D * plane + (long long)y * Ww + x;
    for (int k = 0; k < D; ++k) {   // even channels sin, odd channels cos; pos_y first, then pos_x (:87-90)
        const float t = __ldg(dim_t + k), py = ye / t, px = xe / t;
        out[o0 + (long long)k * plane] = (k & 1) ? cosf(py) : sinf(py);
        out[o0 + (long long)(D + k) * plane] = (k & 1) ? cosf(px) : sinf(px);
    }
}

}  // namespace kpf

extern "C" int kpf_linear_rows(const float* X, long long x_bs, long long x_ps, long long x_ks, const float* pos, long long pos_bs,
                               long long pos_ps, long long pos_ks, const long long* pos_idx, const float* W, const float* bias, int B,
                               int P, int K, int O, float scale, int relu, float* Y, long long y_bs, long long y_ps, long long y_os,
                               cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(B >= 0 && P >= 0 && K >= 1 && O >= 1 && X != nullptr && W != nullptr && Y != nullptr);
    KPF_REQUIRE(pos_idx == nullptr || pos != nullptr);
    const long long R = (long long)B * P;
    if (R == 0) return 0;
    LinearRows a;
    a.X = X; a.x_bs = x_bs; a.x_ps = x_ps; a.x_ks = x_ks; a.pos = pos; a.pos_bs = pos_bs; a.pos_ps = pos_ps; a.pos_ks = pos_ks;
    a.pos_idx = pos_idx; a.W = W; a.bias = bias; a.Y = Y; a.y_bs = y_bs; a.y_ps = y_ps; a.y_os = y_os;
    a.B = B; a.P = P; a.K = K; a.O = O; a.relu = relu; a.scale = scale;
    const dim3 grid((unsigned)((R + LR_T - 1) / LR_T), (unsigned)((O + LR_T - 1) / LR_T));
    linear_rows_kernel<<<grid, 256, 0, stream>>>(a);
    KPF_CHECK_LAUNCH();
    return 0;
}

extern "C" int kpf_mha_core(const float* Q, long long q_bs, long long q_ps, const float* K, long long k_bs, long long k_ps, const float* V,
                            long long v_bs, long long v_ps, const float* attn_mask, const unsigned char* key_padding_mask, int B, int Pq,
                            int Pk, int C, int H, float* O, long long o_bs, long long o_ps, float* stats, float* weights_out,
                            cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(B >= 0 && Pq >= 0 && Pk >= 1 && H >= 1 && C >= H && C % H == 0 && C / H <= MH_HD);
    KPF_REQUIRE(Q != nullptr && K != nullptr && V != nullptr && O != nullptr);
    KPF_REQUIRE(weights_out == nullptr || stats != nullptr);
    if (B == 0 || Pq == 0) return 0;
    MhaCore a;
    a.Q = Q; a.K = K; a.V = V; a.q_bs = q_bs; a.q_ps = q_ps; a.k_bs = k_bs; a.k_ps = k_ps; a.v_bs = v_bs; a.v_ps = v_ps;
    a.attn_mask = attn_mask; a.key_pad = key_padding_mask; a.O = O; a.o_bs = o_bs; a.o_ps = o_ps; a.stats = stats;
    a.B = B; a.Pq = Pq; a.Pk = Pk; a.C = C; a.H = H;
    mha_core_kernel<<<dim3((unsigned)((Pq + MH_QT - 1) / MH_QT), (unsigned)B), 256, 0, stream>>>(a);
    KPF_CHECK_LAUNCH();
    if (weights_out) {
        mha_weights_kernel<<<dim3((unsigned)Pq, (unsigned)B), 128, (size_t)(C + 2 * H) * sizeof(float), stream>>>(a, weights_out);
        KPF_CHECK_LAUNCH();
    }
    return 0;
}

extern "C" int kpf_add_layernorm_rows(const float* x, long long x_bs, long long x_ps, long long x_cs, const float* r, const float* gamma,
                                      const float* beta, int B, int P, int C, float eps, float* y, long long y_bs, long long y_ps,
                                      long long y_cs, cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(B >= 0 && P >= 0 && C >= 1 && C <= 1024 && x != nullptr && gamma != nullptr && beta != nullptr && y != nullptr);
    const long long R = (long long)B * P;
    if (R == 0) return 0;
    AddLayerNorm a;
    a.x = x; a.x_bs = x_bs; a.x_ps = x_ps; a.x_cs = x_cs; a.r = r; a.gamma = gamma; a.beta = beta; a.y = y; a.y_bs = y_bs; a.y_ps = y_ps; a.y_cs = y_cs;
    a.B = B; a.P = P; a.C = C; a.eps = eps;
    const size_t smem = (size_t)LN_ROWS * (C + 1) * sizeof(float);
    cudaError_t e = kpf::set_smem(add_layernorm_kernel, smem);
    if (e != cudaSuccess) return (int)e;
    add_layernorm_kernel<<<(unsigned)((R + LN_ROWS - 1) / LN_ROWS), 256, smem, stream>>>(a);
    KPF_CHECK_LAUNCH();
    return 0;
}

extern "C" int kpf_sine_posembed(const float* mask, const float* dim_t, int B, int H, int W, int D, int normalize, float scale, float* out,
                                 cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(B >= 0 && H >= 1 && W >= 1 && D >= 1 && dim_t != nullptr && out != nullptr);
    if (B == 0) return 0;
    const long long n = (long long)B * H * W;
    sine_posembed_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(mask, dim_t, B, H, W, D, normalize, scale, out);
    KPF_CHECK_LAUNCH();
    return 0;
}

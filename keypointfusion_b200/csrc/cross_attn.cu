// K6 / a13: keypoint-token cross-attention decoder layer (model/transfusion_head.py:684-708 -> :132-173 -> :303-556).
// updatedDecoder feeds every layer the same inputs and returns only the last layer's output (:705-708), so a
// forward is exactly ONE TransformerDecoderLayer(cross_only=True):
//   q_in = anchor + self_posembed ; k_in = v_in = tokens + cross_posembed                 (:144-163)
//   Q = (q_in Wq^T + bq) * hd^-0.5 ; [K|V] = k_in Wkv^T + bkv                              (:403, :418, :468)
//   per head: softmax(Q K^T) V ; out-proj ; x = LN2(anchor + .) ; x = LN3(x + W2 relu(W1 x + b1) + b2)
//   output [B, C, J]                                                                        (:172)
// fp32 CUDA-core version: one CTA per sample, activations in shared memory, weights streamed (pre-transposed,
// coalesced) from L2.  The packed weight blob layout is documented in include/kpf_b200.h.
#include "common.cuh"

namespace kpf {

// Y[t][o] = X[t][:] . Wt[:, o] + bias[o]   for t < J, o < O;   X in smem [J][K], Wt global [K][O] (transposed weight)
__device__ __forceinline__ void token_linear(const float* __restrict__ X, int ldx, const float* __restrict__ Wt,
                                             const float* __restrict__ bias, int J, int K, int O, float* __restrict__ Y, int ldy,
                                             float scale, bool relu) {
    // thread -> output column o, token group tg: tokens tg, tg+G, ...  (needs G*32 >= J)
    const int G = blockDim.x / O > 0 ? blockDim.x / O : 1;
    for (int idx = threadIdx.x; idx < O * G; idx += blockDim.x) {
        const int o = idx % O, tg = idx / O;
        float acc[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) acc[i] = 0.f;
        for (int k = 0; k < K; ++k) {
            const float w = __ldg(Wt + (size_t)k * O + o);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const int t = tg + i * G;
                if (t < J) acc[i] += X[t * ldx + k] * w;
            }
        }
        const float bo = bias[o];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int t = tg + i * G;
            if (t < J) {
                float v = (acc[i] + bo) * scale;
                if (relu) v = fmaxf(v, 0.f);
                Y[t * ldy + o] = v;
            }
        }
    }
}

// x[t][:] = LayerNorm(x[t][:] + r[t][:]) * g + b   (one warp per token), eps = 1e-5
__device__ __forceinline__ void add_layernorm(float* __restrict__ x, const float* __restrict__ r, int J, int C,
                                              const float* __restrict__ g, const float* __restrict__ bta, float eps) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int t = warp; t < J; t += nw) {
        float s = 0.f;
        for (int k = lane; k < C; k += 32) {
            const float v = x[t * C + k] + r[t * C + k];
            x[t * C + k] = v;
            s += v;
        }
        const float mean = warp_sum(s) / (float)C;
        float q = 0.f;
        for (int k = lane; k < C; k += 32) {
            const float d = x[t * C + k] - mean;
            q += d * d;
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
        for (int k = lane; k < C; k += 32) x[t * C + k] = (x[t * C + k] - mean) * rstd * g[k] + bta[k];
    }
}

__global__ void __launch_bounds__(256)
cross_decoder_kernel(const float* __restrict__ anchor, const float* __restrict__ tokens, const float* __restrict__ wp, int J, int C,
                     int F, int heads, float* __restrict__ out_cj, float* __restrict__ out_jc, int out_jc_stride, int out_jc_c0) {
    extern __shared__ __align__(16) float k6sm[];
    const int b = blockIdx.x, tid = threadIdx.x, hd = C / heads;
    // packed weights (floats) -- see include/kpf_b200.h
    const float* self_pos = wp;
    const float* cross_pos = self_pos + (size_t)J * C;
    const float* WqT = cross_pos + (size_t)J * C;
    const float* bq = WqT + (size_t)C * C;
    const float* WkvT = bq + C;
    const float* bkv = WkvT + (size_t)C * 2 * C;
    const float* WoT = bkv + 2 * C;
    const float* bo = WoT + (size_t)C * C;
    const float* g2 = bo + C;
    const float* b2n = g2 + C;
    const float* W1T = b2n + C;
    const float* b1 = W1T + (size_t)C * F;
    const float* W2T = b1 + F;
    const float* b2 = W2T + (size_t)F * C;
    const float* g3 = b2 + C;
    const float* b3n = g3 + C;

    float* sX = k6sm;                       // [J][C]   anchor (residual stream)
    float* sA = sX + (size_t)J * C;         // [J][C]   q_in, later attention output / ffn hidden
    float* sB = sA + (size_t)J * (C > F ? C : F);  // [J][C] k_in, later projection outputs
    float* sQ = sB + (size_t)J * C;         // [J][C]
    float* sKV = sQ + (size_t)J * C;        // [J][2C]
    float* sP = sKV + (size_t)J * 2 * C;    // [heads][J][J]

    for (int i = tid; i < J * C; i += blockDim.x) {
        const float a = anchor[(size_t)b * J * C + i];
        sX[i] = a;
        sA[i] = a + self_pos[i];
        sB[i] = tokens[(size_t)b * J * C + i] + cross_pos[i];
    }
    __syncthreads();
    token_linear(sA, C, WqT, bq, J, C, C, sQ, C, rsqrtf((float)hd), false);
    token_linear(sB, C, WkvT, bkv, J, C, 2 * C, sKV, 2 * C, 1.f, false);
    __syncthreads();
    // scores + softmax: one warp per (head, query) row
    {
        const int lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
        for (int row = warp; row < heads * J; row += nw) {
            const int h = row / J, i = row - h * J;
            float sc[2] = {-INFINITY, -INFINITY};
            for (int jj = lane, s = 0; jj < J; jj += 32, ++s) {
                float d = 0.f;
                for (int k = 0; k < hd; ++k) d += sQ[i * C + h * hd + k] * sKV[jj * 2 * C + h * hd + k];
                sc[s] = d;
            }
            const float mx = warp_max(fmaxf(sc[0], sc[1]));
            float e0 = lane < J ? expf(sc[0] - mx) : 0.f, e1 = lane + 32 < J ? expf(sc[1] - mx) : 0.f;
            const float inv = 1.f / warp_sum(e0 + e1);
            if (lane < J) sP[(h * J + i) * J + lane] = e0 * inv;
            if (lane + 32 < J) sP[(h * J + i) * J + lane + 32] = e1 * inv;
        }
    }
    __syncthreads();
    // attention output -> sA [J][C]
    for (int i = tid; i < J * C; i += blockDim.x) {
        const int t = i / C, cc = i - t * C, h = cc / hd;
        float a = 0.f;
        for (int jj = 0; jj < J; ++jj) a += sP[(h * J + t) * J + jj] * sKV[jj * 2 * C + C + cc];
        sA[i] = a;
    }
    __syncthreads();
    token_linear(sA, C, WoT, bo, J, C, C, sB, C, 1.f, false);
    __syncthreads();
    add_layernorm(sX, sB, J, C, g2, b2n, 1e-5f);  // norm2(query + attn)  :164-165
    __syncthreads();
    token_linear(sX, C, W1T, b1, J, C, F, sA, F, 1.f, true);
    __syncthreads();
    token_linear(sA, F, W2T, b2, J, F, C, sB, C, 1.f, false);
    __syncthreads();
    add_layernorm(sX, sB, J, C, g3, b3n, 1e-5f);  // norm3  :167-169
    __syncthreads();
    for (int i = tid; i < J * C; i += blockDim.x) {
        const int t = i / C, cc = i - t * C;
        if (out_cj) out_cj[((size_t)b * C + cc) * J + t] = sX[i];                                   // [B,C,J]  :172
        if (out_jc) out_jc[((size_t)b * J + t) * out_jc_stride + out_jc_c0 + cc] = sX[i];           // [B,J,C] view for the next stage
    }
}

}  // namespace kpf

extern "C" int kpf_cross_decoder_layer(const float* anchor, const float* tokens, const float* wpack, int B, int J, int C, int F,
                                       int heads, float* out_cj, float* out_jc, int out_jc_stride, int out_jc_c0,
                                       cudaStream_t stream) {
    using namespace kpf;
    // token_linear keeps <= 32 tokens per thread
    KPF_REQUIRE(B >= 0 && J >= 1 && J <= 32 && C >= 32 && C <= 256 && F >= 1 && F <= 256 && heads >= 1 && C % heads == 0);
    if (B == 0) return 0;
    const int mx = C > F ? C : F;
    const size_t smem = ((size_t)J * C * 3 + (size_t)J * mx + (size_t)J * 2 * C + (size_t)heads * J * J) * sizeof(float);
    cudaError_t e = kpf::set_smem(cross_decoder_kernel, smem);
    if (e != cudaSuccess) return (int)e;
    cross_decoder_kernel<<<B, 256, smem, stream>>>(anchor, tokens, wpack, J, C, F, heads, out_cj, out_jc, out_jc_stride, out_jc_c0);
    KPF_CHECK_LAUNCH();
    return 0;
}

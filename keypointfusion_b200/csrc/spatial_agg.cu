// K5 / a12 (+ fused K4c = a10, a11): depth-keypoint spatial attention and aggregation, model/model.py:334-344.
//   hm   = joint2heatmap(r3d[:, :, :2], 0.8, H, sigma=1)            generateFeature.py:584-600
//   GAM  = img2anchor_dis(r3d, img_down, ...)                       loader.py:791-819
//   sw   = sigmoid(Conv1x1_{C+J -> J}(cat[F_rgb, hm]))              -> returned "spatial_weight_loss"
//   w    = sig(weight_dis)*GAM + (1-sig(weight_dis))*sw             (>= 0, so relu(w*f) == w*relu(f))
//   out[j,c] = sum_hw fc_w[hw]*w[j,hw]*relu(F[c,hw]) + fc_b ;  stage 2: relu((out + prev)/2)
// The reference materialises a [B,J,C,HW] product (704 MB fp32 at B=64); here it is two small per-sample
// contractions ([J x (C+J)] x [(C+J) x HW] and [J x HW] x [HW x C]) and nothing is materialised.
// fp32 CUDA-core version: one CTA per sample sweeps HW in tiles of 64 cells staged in shared memory.
#include "common.cuh"

namespace kpf {

constexpr int K5_TH = 64;   // cells per tile
constexpr int K5_LD = 65;   // padded leading dimension of the staged feature tile

template <typename T>
__global__ void __launch_bounds__(256)
spatial_aggregate_kernel(const T* __restrict__ feat_rgb, const float* __restrict__ joints /*[B,J,3] uvd*/,
                         const float* __restrict__ depth, long long depth_bs, int depth_rs, int depth_cs,
                         const float* __restrict__ center, const float* __restrict__ M, const float* __restrict__ cube,
                         const float* __restrict__ cam, const float* __restrict__ Wa /*[J][C+J]*/, const float* __restrict__ ba,
                         const float* __restrict__ weight_dis, const float* __restrict__ fc_w /*[HW]*/,
                         const float* __restrict__ fc_b, const float* __restrict__ prev /*[B,J,C] or null*/, int C, int J, int fs,
                         float img_size, float flip, float hm_std, float hm_sigma, float gamma, float* __restrict__ sw_out,
                         float* __restrict__ feat_j_out, float* __restrict__ hm_out, float* __restrict__ gam_out) {
    extern __shared__ __align__(16) float k5sm[];
    const int b = blockIdx.x, tid = threadIdx.x, HW = fs * fs, CJ = C + J;
    float* sF = k5sm;                         // [C][K5_LD]
    float* sWa = sF + (size_t)C * K5_LD;      // [J][CJ]
    float* sHm = sWa + (size_t)J * CJ;        // [J][K5_TH]
    float* sG = sHm + (size_t)J * K5_TH;      // [J][K5_TH]
    float* sJ = sG + (size_t)J * K5_TH;       // [J][8]: heat-map centre (x,y), xyz, pad
    __shared__ CamF c;
    if (tid == 0) load_cam(c, b, center, M, cube, cam, img_size, flip);
    for (int i = tid; i < J * CJ; i += blockDim.x) sWa[i] = Wa[i];
    __syncthreads();
    for (int j = tid; j < J; j += blockDim.x) {
        const float* s = joints + ((size_t)b * J + j) * 3;
        sJ[8 * j + 0] = (s[0] + 1.f) / 2.f * (float)fs;  // generateFeature.py:592-593
        sJ[8 * j + 1] = (s[1] + 1.f) / 2.f * (float)fs;
        const float3 q = uvd2xyz(c, s[0], s[1], s[2]);   // loader.py:800
        sJ[8 * j + 2] = q.x;
        sJ[8 * j + 3] = q.y;
        sJ[8 * j + 4] = q.z;
    }
    const float sg = 1.f / (1.f + expf(-weight_dis[0]));
    const float inv2s2 = 1.f / (2.f * hm_sigma * hm_sigma);
    const float ffs = (float)fs;
    // phase-B ownership: channel cB, joints jB0, jB0+nhalf, ...  (blockDim = 256 >= C -> 256/C joint groups)
    const int groups = blockDim.x / C > 0 ? blockDim.x / C : 1;  // C=128 -> 2
    const int cB = tid % C, gB = tid / C;
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    const T* fb = feat_rgb + (size_t)b * C * HW;

    for (int t0 = 0; t0 < HW; t0 += K5_TH) {
        const int th = min(K5_TH, HW - t0);
        __syncthreads();
        // stage F tile (coalesced along hw)
        for (int i = tid; i < C * K5_TH; i += blockDim.x) {
            const int cc = i / K5_TH, h = i - cc * K5_TH;
            sF[cc * K5_LD + h] = h < th ? to_f32(fb[(size_t)cc * HW + t0 + h]) : 0.f;
        }
        // heat-map + GAM for the tile
        for (int i = tid; i < J * K5_TH; i += blockDim.x) {
            const int j = i / K5_TH, h = i - j * K5_TH;
            float hmv = 0.f, gv = 0.f;
            if (h < th) {
                const int m = t0 + h, r = m / fs, col = m - r * fs;
                const float dx = ((float)col + 0.5f - sJ[8 * j]) / hm_std, dy = ((float)r + 0.5f - sJ[8 * j + 1]) / hm_std;
                hmv = expf(-(dx * dx + dy * dy) * inv2s2);
                const float d = __ldg(depth + (size_t)b * depth_bs + (size_t)r * depth_rs + (size_t)col * depth_cs);
                const float3 q = uvd2xyz(c, cell_coord(col, ffs), cell_coord(r, ffs), d);
                const float ex = q.x - sJ[8 * j + 2], ey = q.y - sJ[8 * j + 3], ez = q.z - sJ[8 * j + 4];
                gv = 1.f / (gamma * (ex * ex + ey * ey + ez * ez) + 1.f);
                if (hm_out) hm_out[((size_t)b * J + j) * HW + m] = hmv;
                if (gam_out) gam_out[((size_t)b * J + j) * HW + m] = gv;
            }
            sHm[i] = hmv;
            sG[i] = gv;  // GAM for now; overwritten by the blended weight below
        }
        __syncthreads();
        // phase A: S1[j][h] for j = jg, jg+4, ...   (thread: h = tid%64, jg = tid/64)
        {
            const int h = tid % K5_TH, jg = tid / K5_TH, nj = blockDim.x / K5_TH;  // 4 joint groups
            float s1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) s1[i] = 0.f;
            for (int cc = 0; cc < C; ++cc) {
                const float f = sF[cc * K5_LD + h];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = jg + i * nj;
                    if (j < J) s1[i] += sWa[j * CJ + cc] * f;
                }
            }
            for (int jj = 0; jj < J; ++jj) {
                const float f = sHm[jj * K5_TH + h];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int j = jg + i * nj;
                    if (j < J) s1[i] += sWa[j * CJ + C + jj] * f;
                }
            }
            __syncthreads();  // all reads of sG (as GAM) happen below per-thread on own entries only
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int j = jg + i * nj;
                if (j < J) {
                    const float swv = 1.f / (1.f + expf(-(s1[i] + ba[j])));
                    const float w = sg * sG[j * K5_TH + h] + (1.f - sg) * swv;  // model.py:337-338
                    if (h < th) {
                        sw_out[((size_t)b * J + j) * HW + t0 + h] = swv;
                        sG[j * K5_TH + h] = w * fc_w[t0 + h];
                    } else {
                        sG[j * K5_TH + h] = 0.f;
                    }
                }
            }
        }
        __syncthreads();
        // phase B: acc[j][c] += G[j][h] * relu(F[c][h])
        if (tid < groups * C) {
            for (int h = 0; h < K5_TH; ++h) {
                const float f = fmaxf(sF[cB * K5_LD + h], 0.f);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const int j = gB + i * groups;
                    if (j < J) acc[i] += sG[j * K5_TH + h] * f;
                }
            }
        }
    }
    if (tid < groups * C) {
        const float fb0 = fc_b[0];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int j = gB + i * groups;
            if (j < J) {
                float v = acc[i] + fb0;
                const size_t o = ((size_t)b * J + j) * C + cB;
                if (prev) v = fmaxf((v + prev[o]) * 0.5f, 0.f);  // model.py:343-344
                feat_j_out[o] = v;
            }
        }
    }
}

}  // namespace kpf

extern "C" int kpf_spatial_aggregate(const void* feat_rgb, int dtype, const float* joints, const float* depth, long long depth_bs,
                                     int depth_rs, int depth_cs, const float* center, const float* M, const float* cube,
                                     const float* cam, const float* Wa, const float* ba, const float* weight_dis, const float* fc_w,
                                     const float* fc_b, const float* prev, int B, int C, int J, int fs, float img_size, float flip,
                                     float hm_std, float hm_sigma, float gamma, float* sw_out, float* feat_j_out, float* hm_out,
                                     float* gam_out, cudaStream_t stream) {
    using namespace kpf;
    // 256 threads: 4 joint groups x 8 (phase A) and 256/C groups x 16 (phase B) must cover J
    KPF_REQUIRE(B >= 0 && C >= 1 && C <= 256 && J >= 1 && J <= 32 && fs >= 1);
    KPF_REQUIRE((256 / C) * 16 >= J);
    if (B == 0) return 0;
    const size_t smem = ((size_t)C * K5_LD + (size_t)J * (C + J) + 2 * (size_t)J * K5_TH + (size_t)J * 8) * sizeof(float);
    if (dtype == KPF_F32) {
        cudaError_t e = kpf::set_smem(spatial_aggregate_kernel<float>, smem);
        if (e != cudaSuccess) return (int)e;
        spatial_aggregate_kernel<float><<<B, 256, smem, stream>>>((const float*)feat_rgb, joints, depth, depth_bs, depth_rs, depth_cs,
                                                                 center, M, cube, cam, Wa, ba, weight_dis, fc_w, fc_b, prev, C, J, fs,
                                                                 img_size, flip, hm_std, hm_sigma, gamma, sw_out, feat_j_out, hm_out,
                                                                 gam_out);
    } else if (dtype == KPF_BF16) {
        cudaError_t e =
            kpf::set_smem(spatial_aggregate_kernel<__nv_bfloat16>, smem);
        if (e != cudaSuccess) return (int)e;
        spatial_aggregate_kernel<__nv_bfloat16><<<B, 256, smem, stream>>>(
            (const __nv_bfloat16*)feat_rgb, joints, depth, depth_bs, depth_rs, depth_cs, center, M, cube, cam, Wa, ba, weight_dis, fc_w,
            fc_b, prev, C, J, fs, img_size, flip, hm_std, hm_sigma, gamma, sw_out, feat_j_out, hm_out, gam_out);
    } else {
        return KPF_ERR_UNSUPPORTED;
    }
    KPF_CHECK_LAUNCH();
    return 0;
}

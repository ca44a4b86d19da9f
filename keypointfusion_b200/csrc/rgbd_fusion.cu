// K7 / a14-a15: feature-level RGB-D fusion modules (model/fusion_layer.py).  Pure HBM-bound elementwise work
// with one tiny per-pixel (RGBDFusion) or per-channel (ACFusion / FSP) gate.
#include "common.cuh"

namespace kpf {

// RGBDFusion.forward fusion_layer.py:56-83.  CTA = (32-pixel tile, sample); 8 warps split the channels for the
// two gate dot products, then sweep the channels again (L2-resident) to write the three outputs.
template <typename T>
__global__ void __launch_bounds__(256)
rgbd_fusion_kernel(const T* __restrict__ rgb, const T* __restrict__ depth, const float* __restrict__ gate_w /*[2][2C]*/,
                   const float* __restrict__ gate_b /*[2]*/, int C, int HW, T* __restrict__ rgb_out, T* __restrict__ depth_out,
                   T* __restrict__ merge_out, float* __restrict__ attn_sum /*[2] or null*/) {
    __shared__ float part[8][2][32];
    __shared__ float att[2][32];
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane;
    const bool ok = p < HW;
    const size_t base = (size_t)b * C * HW;
    float l = 0.f, r = 0.f;
    for (int c = warp; c < C; c += 8) {
        const float xr = ok ? to_f32(rgb[base + (size_t)c * HW + p]) : 0.f;
        const float xd = ok ? to_f32(depth[base + (size_t)c * HW + p]) : 0.f;
        l += gate_w[c] * xr + gate_w[C + c] * xd;              // gate_rgb(cat)    :61
        r += gate_w[2 * C + c] * xr + gate_w[3 * C + c] * xd;  // gate_depth(cat)  :62
    }
    part[warp][0][lane] = l;
    part[warp][1][lane] = r;
    __syncthreads();
    if (warp == 0) {
        float sl = gate_b[0], sr = gate_b[1];
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            sl += part[w][0][lane];
            sr += part[w][1][lane];
        }
        const float m = fmaxf(sl, sr);
        const float el = expf(sl - m), er = expf(sr - m);
        const float inv = 1.f / (el + er);
        att[0][lane] = el * inv;  // softmax over the 2 gates  :65
        att[1][lane] = er * inv;
        if (attn_sum) {  // optional train_writer statistic (:68-72): mean of the two attention maps
            const float a0 = warp_sum(ok ? el * inv : 0.f), a1 = warp_sum(ok ? er * inv : 0.f);
            if (lane == 0) {
                atomicAdd(attn_sum + 0, a0);
                atomicAdd(attn_sum + 1, a1);
            }
        }
    }
    __syncthreads();
    if (!ok) return;
    const float al = att[0][lane], ar = att[1][lane];
    for (int c = warp; c < C; c += 8) {
        const size_t i = base + (size_t)c * HW + p;
        const float xr = to_f32(rgb[i]), xd = to_f32(depth[i]);
        const float mg = xr * al + xd * ar;                      // :74
        merge_out[i] = from_f32<T>(mg);
        rgb_out[i] = from_f32<T>(fmaxf((xr + mg) * 0.5f, 0.f));  // :76,:80
        depth_out[i] = from_f32<T>(fmaxf((xd + mg) * 0.5f, 0.f));
    }
}

// 16-byte-chunk version of the kernel above for maps with HW % (16 / sizeof(T)) == 0 (every ResNet stage shape): thread = (chunk column
// of PPC consecutive pixels, channel lane), so a warp reads up to 512 contiguous bytes of a channel row per load instead of 64, and the
// deep, small stages (8x8, 4x4: one or two chunks per row) still keep all 256 threads busy by spreading them over the channels.
// CC = chunk columns per CTA; the per-pixel gate sums meet through a fixed-order reduction (shuffles, then the 8 warps' partials).
template <typename T> struct Vec16;
template <> struct Vec16<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void unpack(const uint4& v, float* f) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 t = __bfloat1622float2(h[i]);
            f[2 * i] = t.x;
            f[2 * i + 1] = t.y;
        }
    }
    static __device__ __forceinline__ uint4 pack(const float* f) {
        uint4 o;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
        return o;
    }
};
template <> struct Vec16<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void unpack(const uint4& v, float* f) {
        const float* s = reinterpret_cast<const float*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) f[i] = s[i];
    }
    static __device__ __forceinline__ uint4 pack(const float* f) { return *reinterpret_cast<const uint4*>(f); }
};

template <typename T, int CC>
__global__ void __launch_bounds__(256)
rgbd_fusion_vec_kernel(const T* __restrict__ rgb, const T* __restrict__ depth, const float* __restrict__ gate_w, const float* __restrict__ gate_b,
                       int C, int HW, T* __restrict__ rgb_out, T* __restrict__ depth_out, T* __restrict__ merge_out,
                       float* __restrict__ attn_sum) {
    constexpr int PPC = Vec16<T>::N, CL = 256 / CC;
    __shared__ float part[8][CC][2 * PPC];
    __shared__ float att[2][CC * PPC];
    const int b = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, cc = tid % CC, cl = tid / CC;
    const int chunk = blockIdx.x * CC + cc;
    const bool ok = chunk < HW / PPC;
    const size_t base = (size_t)b * C * HW + (size_t)chunk * PPC;
    float l[PPC], r[PPC];
#pragma unroll
    for (int i = 0; i < PPC; ++i) l[i] = r[i] = 0.f;
    if (ok) {
#pragma unroll 2
        for (int c = cl; c < C; c += CL) {
            float xr[PPC], xd[PPC];
            Vec16<T>::unpack(__ldg(reinterpret_cast<const uint4*>(rgb + base + (size_t)c * HW)), xr);
            Vec16<T>::unpack(__ldg(reinterpret_cast<const uint4*>(depth + base + (size_t)c * HW)), xd);
            const float w0 = __ldg(gate_w + c), w1 = __ldg(gate_w + C + c), w2 = __ldg(gate_w + 2 * C + c), w3 = __ldg(gate_w + 3 * C + c);
#pragma unroll
            for (int i = 0; i < PPC; ++i) {
                l[i] += w0 * xr[i] + w1 * xd[i];   // gate_rgb(cat)    :61
                r[i] += w2 * xr[i] + w3 * xd[i];   // gate_depth(cat)  :62
            }
        }
    }
#pragma unroll
    for (int off = CC; off < 32; off <<= 1)   // the warp's channel lanes of one chunk column sit CC lanes apart
#pragma unroll
        for (int i = 0; i < PPC; ++i) {
            l[i] += __shfl_xor_sync(0xffffffffu, l[i], off);
            r[i] += __shfl_xor_sync(0xffffffffu, r[i], off);
        }
    if (lane < CC) {
#pragma unroll
        for (int i = 0; i < PPC; ++i) {
            part[warp][lane][i] = l[i];
            part[warp][lane][PPC + i] = r[i];
        }
    }
    __syncthreads();
    if (warp * 32 < CC * PPC) {   // one thread per pixel of the tile (warp-uniform branch: the statistic below is a warp reduction)
        const bool mine = tid < CC * PPC;
        const int c2 = mine ? tid / PPC : 0, i = mine ? tid - c2 * PPC : 0;
        float sl = gate_b[0], sr = gate_b[1];
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            sl += part[w][c2][i];
            sr += part[w][c2][PPC + i];
        }
        const float m = fmaxf(sl, sr);
        const float el = expf(sl - m), er = expf(sr - m);
        const float inv = 1.f / (el + er);
        if (mine) {
            att[0][tid] = el * inv;   // softmax over the 2 gates  :65
            att[1][tid] = er * inv;
        }
        if (attn_sum) {               // optional train_writer statistic (:68-72)
            const bool valid = mine && blockIdx.x * CC + c2 < HW / PPC;
            const float a0 = warp_sum(valid ? el * inv : 0.f), a1 = warp_sum(valid ? er * inv : 0.f);
            if (lane == 0) {
                atomicAdd(attn_sum + 0, a0);
                atomicAdd(attn_sum + 1, a1);
            }
        }
    }
    __syncthreads();
    if (!ok) return;
    float al[PPC], ar[PPC];
#pragma unroll
    for (int i = 0; i < PPC; ++i) {
        al[i] = att[0][cc * PPC + i];
        ar[i] = att[1][cc * PPC + i];
    }
#pragma unroll 2
    for (int c = cl; c < C; c += CL) {   // second sweep: the tile's channels are L2 / L1 resident
        const size_t o = base + (size_t)c * HW;
        float xr[PPC], xd[PPC], mg[PPC], orr[PPC], od[PPC];
        Vec16<T>::unpack(__ldg(reinterpret_cast<const uint4*>(rgb + o)), xr);
        Vec16<T>::unpack(__ldg(reinterpret_cast<const uint4*>(depth + o)), xd);
#pragma unroll
        for (int i = 0; i < PPC; ++i) {
            mg[i] = xr[i] * al[i] + xd[i] * ar[i];          // :74
            orr[i] = fmaxf((xr[i] + mg[i]) * 0.5f, 0.f);    // :76, :80
            od[i] = fmaxf((xd[i] + mg[i]) * 0.5f, 0.f);
        }
        *reinterpret_cast<uint4*>(merge_out + o) = Vec16<T>::pack(mg);
        *reinterpret_cast<uint4*>(rgb_out + o) = Vec16<T>::pack(orr);
        *reinterpret_cast<uint4*>(depth_out + o) = Vec16<T>::pack(od);
    }
}

template <typename T>
static cudaError_t launch_rgbd_vec(const T* rgb, const T* depth, const float* gate_w, const float* gate_b, int B, int C, int HW, T* rgb_out,
                                   T* depth_out, T* merge_out, float* attn_sum, cudaStream_t stream) {
    const int cpr = HW / Vec16<T>::N;   // chunks per channel row
    int cc = 2;                         // chunk columns per CTA: a quarter of the row (>= 4 CTAs per sample), between 2 and 32
    while (cc < 32 && cc * 2 <= cpr / 4) cc *= 2;
    const dim3 grid((cpr + cc - 1) / cc, B);
#define KPF_RV(CCV) rgbd_fusion_vec_kernel<T, CCV><<<grid, 256, 0, stream>>>(rgb, depth, gate_w, gate_b, C, HW, rgb_out, depth_out, merge_out, attn_sum)
    switch (cc) {
        case 2: KPF_RV(2); break;
        case 4: KPF_RV(4); break;
        case 8: KPF_RV(8); break;
        case 16: KPF_RV(16); break;
        default: KPF_RV(32); break;
    }
#undef KPF_RV
    return cudaGetLastError();
}

// AdaptiveAvgPool2d(1): one warp per (b, c) row.
template <typename T>
__global__ void __launch_bounds__(256)
channel_mean_kernel(const T* __restrict__ x, int rows, int HW, float* __restrict__ out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const T* p = x + (size_t)row * HW;
    float s = 0.f;
    constexpr int PPC = Vec16<T>::N;
    if (HW % PPC == 0 && (((uintptr_t)x) & 15) == 0) {
        for (int ch = lane; ch < HW / PPC; ch += 32) {
            float v[PPC];
            Vec16<T>::unpack(__ldg(reinterpret_cast<const uint4*>(p + (size_t)ch * PPC)), v);
#pragma unroll
            for (int i = 0; i < PPC; ++i) s += v[i];
        }
    } else {
        for (int i = lane; i < HW; i += 32) s += to_f32(p[i]);
    }
    s = warp_sum(s);
    if (lane == 0) out[row] = s / (float)HW;
}

// ACFusion.forward fusion_layer.py:101-116.  CTA = (8-channel tile, sample): warp w computes the two channel
// gates sigmoid(W[c,:] . mean + b) for its channel, then streams that channel's HW pixels.
template <typename T>
__global__ void __launch_bounds__(256)
ac_fusion_kernel(const T* __restrict__ rgb, const T* __restrict__ depth, const float* __restrict__ mean_rgb,
                 const float* __restrict__ mean_depth, const float* __restrict__ w_rgb, const float* __restrict__ b_rgb,
                 const float* __restrict__ w_depth, const float* __restrict__ b_depth, int C, int HW, T* __restrict__ rgb_out,
                 T* __restrict__ depth_out, T* __restrict__ merge_out) {
    const int b = blockIdx.y, lane = threadIdx.x & 31, c = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c >= C) return;
    float gr = 0.f, gd = 0.f;
    for (int k = lane; k < C; k += 32) {
        gr += w_rgb[(size_t)c * C + k] * mean_rgb[(size_t)b * C + k];
        gd += w_depth[(size_t)c * C + k] * mean_depth[(size_t)b * C + k];
    }
    gr = 1.f / (1.f + expf(-(warp_sum(gr) + b_rgb[c])));
    gd = 1.f / (1.f + expf(-(warp_sum(gd) + b_depth[c])));
    const size_t base = ((size_t)b * C + c) * HW;
    constexpr int PPC = Vec16<T>::N;
    if (HW % PPC == 0 && (((uintptr_t)rgb | (uintptr_t)depth | (uintptr_t)rgb_out | (uintptr_t)depth_out | (uintptr_t)merge_out) & 15) == 0) {
        for (int ch = lane; ch < HW / PPC; ch += 32) {   // 16-byte chunks of the channel's row
            const size_t o = base + (size_t)ch * PPC;
            float xr[PPC], xd[PPC], mg[PPC], orr[PPC], od[PPC];
            Vec16<T>::unpack(__ldg(reinterpret_cast<const uint4*>(rgb + o)), xr);
            Vec16<T>::unpack(__ldg(reinterpret_cast<const uint4*>(depth + o)), xd);
#pragma unroll
            for (int i = 0; i < PPC; ++i) {
                mg[i] = gr * xr[i] + gd * xd[i];
                orr[i] = fmaxf((xr[i] + mg[i]) * 0.5f, 0.f);
                od[i] = fmaxf((xd[i] + mg[i]) * 0.5f, 0.f);
            }
            *reinterpret_cast<uint4*>(merge_out + o) = Vec16<T>::pack(mg);
            *reinterpret_cast<uint4*>(rgb_out + o) = Vec16<T>::pack(orr);
            *reinterpret_cast<uint4*>(depth_out + o) = Vec16<T>::pack(od);
        }
        return;
    }
    for (int i = lane; i < HW; i += 32) {
        const float xr = to_f32(rgb[base + i]), xd = to_f32(depth[base + i]);
        const float mg = gr * xr + gd * xd;
        merge_out[base + i] = from_f32<T>(mg);
        rgb_out[base + i] = from_f32<T>(fmaxf((xr + mg) * 0.5f, 0.f));
        depth_out[base + i] = from_f32<T>(fmaxf((xd + mg) * 0.5f, 0.f));
    }
}

// FSP.forward fusion_layer.py:33-37 with FilterLayer :18-22.  mean_cat = avgpool(cat(guide, main)) [B,2C].
// CTA = (8-channel tile, sample); the hidden layer (C/r wide) is recomputed per CTA in shared memory.
template <typename T>
__global__ void __launch_bounds__(256)
fsp_kernel(const T* __restrict__ guide, const T* __restrict__ mainp, const float* __restrict__ mean_guide,
           const float* __restrict__ mean_main, const float* __restrict__ w0 /*[Hd][2C]*/, const float* __restrict__ b0,
           const float* __restrict__ w2 /*[C][Hd]*/, const float* __restrict__ b2, int C, int Hd, int HW, T* __restrict__ out) {
    extern __shared__ float hid[];  // [Hd]
    const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int h = warp; h < Hd; h += 8) {
        float s = 0.f;
        for (int k = lane; k < 2 * C; k += 32) {
            const float m = k < C ? mean_guide[(size_t)b * C + k] : mean_main[(size_t)b * C + k - C];
            s += w0[(size_t)h * 2 * C + k] * m;
        }
        s = warp_sum(s);
        if (lane == 0) hid[h] = fmaxf(s + b0[h], 0.f);
    }
    __syncthreads();
    const int c = blockIdx.x * 8 + warp;
    if (c >= C) return;
    float g = 0.f;
    for (int h = lane; h < Hd; h += 32) g += w2[(size_t)c * Hd + h] * hid[h];
    g = 1.f / (1.f + expf(-(warp_sum(g) + b2[c])));
    const size_t base = ((size_t)b * C + c) * HW;
    constexpr int PPC = Vec16<T>::N;
    if (HW % PPC == 0 && (((uintptr_t)guide | (uintptr_t)mainp | (uintptr_t)out) & 15) == 0) {
        for (int ch = lane; ch < HW / PPC; ch += 32) {   // 16-byte chunks of the channel's row
            const size_t o = base + (size_t)ch * PPC;
            float xm[PPC], xg[PPC];
            Vec16<T>::unpack(__ldg(reinterpret_cast<const uint4*>(mainp + o)), xm);
            Vec16<T>::unpack(__ldg(reinterpret_cast<const uint4*>(guide + o)), xg);
#pragma unroll
            for (int i = 0; i < PPC; ++i) xm[i] = xm[i] + g * xg[i];
            *reinterpret_cast<uint4*>(out + o) = Vec16<T>::pack(xm);
        }
        return;
    }
    for (int i = lane; i < HW; i += 32) out[base + i] = from_f32<T>(to_f32(mainp[base + i]) + g * to_f32(guide[base + i]));
}

}  // namespace kpf

using namespace kpf;

#define KPF_DISPATCH_DTYPE(dtype, ...)                \
    if (dtype == KPF_F32) {                           \
        using T = float;                              \
        __VA_ARGS__;                                  \
    } else if (dtype == KPF_BF16) {                   \
        using T = __nv_bfloat16;                      \
        __VA_ARGS__;                                  \
    } else                                            \
        return KPF_ERR_UNSUPPORTED;

extern "C" int kpf_rgbd_fusion(const void* rgb, const void* depth, int dtype, const float* gate_w, const float* gate_b, int B,
                               int C, int HW, void* rgb_out, void* depth_out, void* merge_out, float* attn_sum,
                               cudaStream_t stream) {
    KPF_REQUIRE(B >= 0 && C >= 1 && HW >= 1);
    if (B == 0) return 0;
    const int ppc = dtype == KPF_F32 ? 4 : 8;
    const bool vec = HW % ppc == 0 && (((uintptr_t)rgb | (uintptr_t)depth | (uintptr_t)rgb_out | (uintptr_t)depth_out | (uintptr_t)merge_out) & 15) == 0;
    if (vec) {
        cudaError_t e;
        KPF_DISPATCH_DTYPE(dtype, (e = launch_rgbd_vec<T>((const T*)rgb, (const T*)depth, gate_w, gate_b, B, C, HW, (T*)rgb_out, (T*)depth_out,
                                                          (T*)merge_out, attn_sum, stream)));
        return e == cudaSuccess ? 0 : (int)e;
    }
    dim3 grid((HW + 31) / 32, B);
    KPF_DISPATCH_DTYPE(dtype, (rgbd_fusion_kernel<T><<<grid, 256, 0, stream>>>((const T*)rgb, (const T*)depth, gate_w, gate_b, C, HW,
                                                                              (T*)rgb_out, (T*)depth_out, (T*)merge_out, attn_sum)));
    KPF_CHECK_LAUNCH();
    return 0;
}

extern "C" int kpf_channel_mean(const void* x, int dtype, int rows, int HW, float* out, cudaStream_t stream) {
    KPF_REQUIRE(rows >= 0 && HW >= 1);
    if (rows == 0) return 0;
    KPF_DISPATCH_DTYPE(dtype, (channel_mean_kernel<T><<<(rows + 7) / 8, 256, 0, stream>>>((const T*)x, rows, HW, out)));
    KPF_CHECK_LAUNCH();
    return 0;
}

extern "C" int kpf_ac_fusion(const void* rgb, const void* depth, int dtype, const float* mean_rgb, const float* mean_depth,
                             const float* w_rgb, const float* b_rgb, const float* w_depth, const float* b_depth, int B, int C,
                             int HW, void* rgb_out, void* depth_out, void* merge_out, cudaStream_t stream) {
    KPF_REQUIRE(B >= 0 && C >= 1 && HW >= 1);
    if (B == 0) return 0;
    dim3 grid((C + 7) / 8, B);
    KPF_DISPATCH_DTYPE(dtype, (ac_fusion_kernel<T><<<grid, 256, 0, stream>>>((const T*)rgb, (const T*)depth, mean_rgb, mean_depth, w_rgb,
                                                                            b_rgb, w_depth, b_depth, C, HW, (T*)rgb_out,
                                                                            (T*)depth_out, (T*)merge_out)));
    KPF_CHECK_LAUNCH();
    return 0;
}

extern "C" int kpf_fsp(const void* guide, const void* mainp, int dtype, const float* mean_guide, const float* mean_main,
                       const float* w0, const float* b0, const float* w2, const float* b2, int B, int C, int Hd, int HW, void* out,
                       cudaStream_t stream) {
    KPF_REQUIRE(B >= 0 && C >= 1 && HW >= 1 && Hd >= 1);
    if (B == 0) return 0;
    dim3 grid((C + 7) / 8, B);
    KPF_DISPATCH_DTYPE(dtype, (fsp_kernel<T><<<grid, 256, (size_t)Hd * sizeof(float), stream>>>(
                                  (const T*)guide, (const T*)mainp, mean_guide, mean_main, w0, b0, w2, b2, C, Hd, HW, (T*)out)));
    KPF_CHECK_LAUNCH();
    return 0;
}

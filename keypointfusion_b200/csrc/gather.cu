// K3 / a8: K-tap weighted gather of feature maps at the nearest-cell indices (model/model.py:297-306).
//   out[b,n,c] = sum_k closeness[b,n,k] * feat[b,c,index[b,n,k]]
// One CTA per (channel tile, sample).  The [CT x HW] feature tile is contiguous in the NCHW map, so it is
// staged into shared memory by the TMA engine with 1-D bulk copies (cp.async.bulk -> SASS UBLKCP) completing
// on an mbarrier; the random tap reads then hit shared memory, and each point's CT outputs are written as
// one contiguous run of the [B,N,C] result.  HBM-bound: (C*HW + N*C)*e + N*K*12 bytes per sample.
#include "common.cuh"

namespace kpf {

template <typename T, typename I, int CT>
__global__ void __launch_bounds__(256)
gather_taps_kernel(const T* __restrict__ feat, long long feat_bs, int C, int HW, const I* __restrict__ index,
                   const float* __restrict__ closeness, int N, int K, T* __restrict__ out, int out_stride, int out_c0) {
    extern __shared__ __align__(128) unsigned char gsm[];
    T* tile = reinterpret_cast<T*>(gsm);
    __shared__ __align__(8) uint64_t bar;
    const int b = blockIdx.y, c0 = blockIdx.x * CT;
    const int ct = min(CT, C - c0);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t row_bytes = (uint32_t)HW * sizeof(T);
        mbar_expect_tx(&bar, row_bytes * ct);
        const T* src = feat + (size_t)b * feat_bs + (size_t)c0 * HW;
        for (int c = 0; c < ct; ++c) tma_bulk_g2s(tile + (size_t)c * HW, src + (size_t)c * HW, row_bytes, &bar);
    }
    mbar_wait(&bar, 0);
    for (int n = threadIdx.x; n < N; n += blockDim.x) {
        const size_t o = ((size_t)b * N + n) * K;
        float acc[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) acc[c] = 0.f;
        for (int k = 0; k < K; ++k) {
            const int idx = (int)index[o + k];
            const float w = closeness[o + k];
#pragma unroll
            for (int c = 0; c < CT; ++c)
                if (c < ct) acc[c] += to_f32(tile[c * HW + idx]) * w;
        }
        T* dst = out + ((size_t)b * N + n) * out_stride + out_c0 + c0;
#pragma unroll
        for (int c = 0; c < CT; ++c)
            if (c < ct) dst[c] = from_f32<T>(acc[c]);
    }
}

template <typename T, typename I>
static int launch_gather(const T* feat, long long feat_bs, int B, int C, int HW, const I* index, const float* closeness, int N,
                         int K, T* out, int out_stride, int out_c0, cudaStream_t stream) {
    const size_t row = (size_t)HW * sizeof(T);
    if (row % 16 != 0 || ((uintptr_t)feat % 16) != 0 || ((size_t)feat_bs * sizeof(T)) % 16 != 0) return KPF_ERR_BAD_ARGUMENT;
#define KPF_GATHER(CTV)                                                                                                      \
    {                                                                                                                        \
        const size_t smem = row * CTV;                                                                                       \
        cudaError_t e = kpf::set_smem(gather_taps_kernel<T, I, CTV>, smem); \
        if (e != cudaSuccess) return (int)e;                                                                                 \
        dim3 grid((C + CTV - 1) / CTV, B);                                                                                   \
        gather_taps_kernel<T, I, CTV><<<grid, 256, smem, stream>>>(feat, feat_bs, C, HW, index, closeness, N, K, out, out_stride, out_c0); \
    }
    if (row * 16 <= 96 * 1024) KPF_GATHER(16)
    else if (row * 8 <= 128 * 1024) KPF_GATHER(8)
    else if (row * 2 <= 200 * 1024) KPF_GATHER(2)
    else return KPF_ERR_UNSUPPORTED;
#undef KPF_GATHER
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace kpf

extern "C" int kpf_gather_taps(const void* feat, int dtype, long long feat_batch_stride, int B, int C, int HW, const void* index,
                               int index_is_i64, const float* closeness, int N, int K, void* out, int out_stride, int out_c0,
                               cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(B >= 0 && C >= 1 && HW >= 1 && N >= 0 && K >= 1 && out_stride >= C + out_c0);
    if (B == 0 || N == 0) return 0;
    if (dtype == KPF_F32) {
        if (index_is_i64)
            return launch_gather<float, long long>((const float*)feat, feat_batch_stride, B, C, HW, (const long long*)index, closeness,
                                                   N, K, (float*)out, out_stride, out_c0, stream);
        return launch_gather<float, int32_t>((const float*)feat, feat_batch_stride, B, C, HW, (const int32_t*)index, closeness, N, K,
                                             (float*)out, out_stride, out_c0, stream);
    }
    if (dtype == KPF_BF16) {
        if (index_is_i64)
            return launch_gather<__nv_bfloat16, long long>((const __nv_bfloat16*)feat, feat_batch_stride, B, C, HW,
                                                           (const long long*)index, closeness, N, K, (__nv_bfloat16*)out, out_stride,
                                                           out_c0, stream);
        return launch_gather<__nv_bfloat16, int32_t>((const __nv_bfloat16*)feat, feat_batch_stride, B, C, HW, (const int32_t*)index,
                                                     closeness, N, K, (__nv_bfloat16*)out, out_stride, out_c0, stream);
    }
    return KPF_ERR_UNSUPPORTED;
}

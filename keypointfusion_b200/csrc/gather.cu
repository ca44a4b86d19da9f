// K3 / a8: K-tap weighted gather of feature maps at the nearest-cell indices (model/model.py:297-306).
//   out[b,n,c] = sum_k closeness[b,n,k] * feat[b,c,index[b,n,k]]
// The map is channel-major (NCHW) and the result point-major, so a tap of one point is C values HW apart.
// Main path, ONE kernel (gather_slab_kernel): a CTA owns (sample, block of CHB channels); it reads the block's [CHB x HW] slab once,
// transposing it on the way into shared memory as channels-last rows [HW][CHB] (lane = channel: the 2-byte stores of a warp fill one
// 64-byte row, conflict-free); a tap is then one 64-byte row of shared memory read with 16-byte loads by the point's 4 lanes, the K
// taps are accumulated in fp32 and the point's CHB outputs leave as one 64-byte run.  HBM traffic = the algorithmic bytes: the map
// read once, the result written once (the two-kernel version below moved the map three times and ran at 18-22 % of HBM).
// Fallback for maps whose slab does not fit shared memory (HW * 16 B > 96 KB), two kernels through a global workspace:
//   rows_kernel   : [B,C,HW] -> [B,HW,Cp] (Cp = C rounded up to 8): a [32 channels x 64 cells] tile per CTA staged by the TMA engine
//                   (1-D bulk copies of the 64-cell runs, cp.async.bulk -> SASS UBLKCP, completing on an mbarrier), transposed out of
//                   shared memory with 16-byte stores.  One read + one write of the map at copy speed.
//   gather_rows   : a tap is now ONE contiguous row: the lanes of a point read 16-byte chunks of its K rows (served by L2: the rows
//                   were just written), accumulate in fp32 and write the point's C outputs as one contiguous run.
// HBM-bound: (C*HW + N*C)*e + N*K*12 algorithmic bytes per sample (the row copy adds one L2-resident round trip of C*HW*e).
// (Round 1's single kernel staged [16 x HW] channel tiles and read the taps out of shared memory: every lane of a warp hit the same
//  bank -- the rows are 2048 B apart -- and it ran at 5 % of HBM.)
#include "common.cuh"

namespace kpf {

constexpr int GR_CH = 32, GR_CELLS = 64;

template <typename T>
__global__ void __launch_bounds__(256)
rows_kernel(const T* __restrict__ feat, long long feat_bs, int C, int HW, int Cp, T* __restrict__ rows) {
    __shared__ __align__(128) T tile[GR_CH][GR_CELLS];
    __shared__ __align__(8) uint64_t bar;
    const int b = blockIdx.z, c0 = blockIdx.y * GR_CH, h0 = blockIdx.x * GR_CELLS, tid = threadIdx.x;
    const int ct = min(GR_CH, C - c0), cells = min(GR_CELLS, HW - h0);
    const T* src = feat + (size_t)b * feat_bs + (size_t)c0 * HW + h0;
    const bool tma_ok = cells == GR_CELLS && ((size_t)HW * sizeof(T)) % 16 == 0 && ((uintptr_t)src % 16) == 0;
    if (tma_ok) {
        if (tid == 0) {
            mbar_init(&bar, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (tid == 0) {
            mbar_expect_tx(&bar, (uint32_t)(ct * GR_CELLS * sizeof(T)));
            for (int c = 0; c < ct; ++c) tma_bulk_g2s(&tile[c][0], src + (size_t)c * HW, GR_CELLS * sizeof(T), &bar);
        }
        mbar_wait(&bar, 0);
    } else {
        for (int i = tid; i < ct * GR_CELLS; i += 256) {
            const int c = i / GR_CELLS, h = i - c * GR_CELLS;
            tile[c][h] = h < cells ? src[(size_t)c * HW + h] : T(0);
        }
        __syncthreads();
    }
    // out: thread -> (cell, 8-channel group): a 16-byte (bf16) / 32-byte (f32) run of the cell's row
    for (int i = tid; i < GR_CELLS * (GR_CH / 8); i += 256) {
        const int g = i & (GR_CH / 8 - 1), cell = i >> 2;
        if (cell >= cells || c0 + 8 * g >= Cp) continue;
        T v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = (8 * g + k < ct) ? tile[8 * g + k][cell] : T(0);
        T* dst = rows + ((size_t)b * HW + h0 + cell) * Cp + c0 + 8 * g;
        if (sizeof(T) == 2) {
            *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
        } else {
            reinterpret_cast<uint4*>(dst)[0] = reinterpret_cast<const uint4*>(v)[0];
            reinterpret_cast<uint4*>(dst)[1] = reinterpret_cast<const uint4*>(v)[1];
        }
    }
}

template <typename T> struct Chunk;   // 16 bytes of a row
template <> struct Chunk<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void fma(float* acc, const uint4& v, float w) {
        const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(h[i]);
            acc[2 * i] += f.x * w;
            acc[2 * i + 1] += f.y * w;
        }
    }
    static __device__ __forceinline__ uint4 pack(const float* a) {
        uint4 o;
        __nv_bfloat162 t;
        t = __floats2bfloat162_rn(a[0], a[1]); o.x = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(a[2], a[3]); o.y = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(a[4], a[5]); o.z = *reinterpret_cast<uint32_t*>(&t);
        t = __floats2bfloat162_rn(a[6], a[7]); o.w = *reinterpret_cast<uint32_t*>(&t);
        return o;
    }
};
template <> struct Chunk<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void fma(float* acc, const uint4& v, float w) {
        const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] += f[i] * w;
    }
    static __device__ __forceinline__ uint4 pack(const float* a) { return *reinterpret_cast<const uint4*>(a); }
};

// thread -> (point, 16-byte chunk of the row); K <= 16 taps, four row loads in flight
template <typename T, typename I>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const T* __restrict__ rows, int Cp, int C, int HW, const I* __restrict__ index, const float* __restrict__ closeness,
                   long long n_points, int N, int K, T* __restrict__ out, int out_stride, int out_c0) {
    constexpr int CN = Chunk<T>::N;
    const int chunks = Cp / CN;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long pt = t / chunks;
    const int ck = (int)(t - pt * chunks);
    if (pt >= n_points) return;
    const long long b = pt / N;
    const T* base = rows + (size_t)b * HW * Cp + (size_t)ck * CN;
    float acc[CN];
#pragma unroll
    for (int i = 0; i < CN; ++i) acc[i] = 0.f;
    const I* ix = index + pt * K;
    const float* cw = closeness + pt * K;
    for (int k0 = 0; k0 < K; k0 += 4) {
        uint4 v[4];
        float w[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool ok = k0 + u < K;
            w[u] = ok ? __ldg(cw + k0 + u) : 0.f;
            const long long cell = ok ? (long long)ix[k0 + u] : 0;
            v[u] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)cell * Cp));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) Chunk<T>::fma(acc, v[u], w[u]);
    }
    T* dst = out + (size_t)pt * out_stride + out_c0 + ck * CN;
    if (ck * CN + CN <= C && ((uintptr_t)dst % 16) == 0) {
        *reinterpret_cast<uint4*>(dst) = Chunk<T>::pack(acc);
    } else {
#pragma unroll
        for (int i = 0; i < CN; ++i)
            if (ck * CN + i < C) dst[i] = from_f32<T>(acc[i]);
    }
}

// ---- main path: one kernel, slab in shared memory --------------------------------------------------------------------------------
template <typename T, typename I, int CHB>
__global__ void __launch_bounds__(256)
gather_slab_kernel(const T* __restrict__ feat, long long feat_bs, int C, int HW, const I* __restrict__ index,
                   const float* __restrict__ closeness, int N, int K, T* __restrict__ out, int out_stride, int out_c0) {
    extern __shared__ __align__(128) unsigned char gs_smem[];
    T* rows = reinterpret_cast<T*>(gs_smem);   // [HW][CHB]
    constexpr int CN = Chunk<T>::N;            // elements per 16-byte chunk
    constexpr int CHUNKS = CHB / CN;           // 16-byte chunks per row (4, 2 or 1)
    const int b = blockIdx.y, c0 = blockIdx.x * CHB, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ct = min(CHB, C - c0);
    const T* src = feat + (size_t)b * feat_bs + (size_t)c0 * HW;
    // ---- slab in, transposed.  lane = channel (+ sub-run when CHB < 32); a lane reads 32 contiguous bytes of its channel row (a full
    //      sector, two 16-byte loads) and stores the 2 * CN cells one by one: the lanes of one store instruction cover consecutive
    //      channels of a cell's row (64 contiguous bytes, conflict-free)
    {
        constexpr int SUB = 32 / CHB, RUN = 2 * CN;   // runs a warp handles per step (1, 2 or 4); cells per run
        const int c = lane % CHB, sub = lane / CHB;
        const bool vec_ok = (HW % RUN) == 0 && (((uintptr_t)src) % 32) == 0 && (((size_t)HW * sizeof(T)) % 32) == 0;
        constexpr int STEP = 8 * SUB * RUN, DEPTH = 2;   // cells all warps cover per step; steps whose loads are in flight together
        for (int hb = (warp * SUB + sub) * RUN; hb < HW; hb += DEPTH * STEP) {
            T v[DEPTH][RUN];
#pragma unroll
            for (int dpt = 0; dpt < DEPTH; ++dpt) {   // all the loads first: the DRAM round trip is paid once per DEPTH steps
                const int h0 = hb + dpt * STEP;
                if (h0 < HW && c < ct && vec_ok) {
                    const uint4* g = reinterpret_cast<const uint4*>(src + (size_t)c * HW + h0);
                    reinterpret_cast<uint4*>(v[dpt])[0] = __ldg(g);
                    reinterpret_cast<uint4*>(v[dpt])[1] = __ldg(g + 1);
                } else {
#pragma unroll
                    for (int i = 0; i < RUN; ++i) v[dpt][i] = (c < ct && h0 + i < HW) ? src[(size_t)c * HW + h0 + i] : T(0.f);
                }
            }
#pragma unroll
            for (int dpt = 0; dpt < DEPTH; ++dpt) {
                const int h0 = hb + dpt * STEP;
#pragma unroll
                for (int i = 0; i < RUN; ++i)
                    if (h0 + i < HW) rows[(size_t)(h0 + i) * CHB + c] = v[dpt][i];
            }
        }
    }
    __syncthreads();
    // ---- taps out of shared memory: thread = (point, 16-byte chunk of its row).  The indices and weights of a point are fetched four
    //      taps at a time, and those of the thread's NEXT point before the current point's rows are read: the global round trip
    //      (an L2 hit) would otherwise sit in front of every shared-memory read
    const int chunk = tid % CHUNKS;
    constexpr int PSTEP = 256 / CHUNKS;
    const bool k4 = K == 4 && sizeof(I) == 4 && (((uintptr_t)index) & 15) == 0 && (((uintptr_t)closeness) & 15) == 0;
    auto fetch = [&](int n, int* cell, float* w, int k0) {
        const size_t pt = (size_t)b * N + n;
        if (k4) {   // the common case (model.py:297-306 uses 4 taps): one 16-byte load each
            const int4 ci = n < N ? __ldg(reinterpret_cast<const int4*>(index) + pt) : make_int4(0, 0, 0, 0);
            const float4 cw = n < N ? __ldg(reinterpret_cast<const float4*>(closeness) + pt) : make_float4(0.f, 0.f, 0.f, 0.f);
            cell[0] = ci.x; cell[1] = ci.y; cell[2] = ci.z; cell[3] = ci.w;
            w[0] = cw.x; w[1] = cw.y; w[2] = cw.z; w[3] = cw.w;
            return;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const bool ok = n < N && k0 + u < K;
            cell[u] = ok ? (int)index[pt * K + k0 + u] : 0;
            w[u] = ok ? __ldg(closeness + pt * K + k0 + u) : 0.f;
        }
    };
    // two groups of four taps are always on their way (ring of two): one iteration of a warp is shorter than an L2 round trip
    int cell_q[2][4];
    float w_q[2][4];
    const int groups = (K + 3) / 4;
    int n_f = tid / CHUNKS, g_f = 0;   // the (point, tap group) the next fetch is for
    auto fetch_next = [&](int slot) {
        fetch(n_f, cell_q[slot], w_q[slot], 4 * g_f);
        if (++g_f == groups) {
            g_f = 0;
            n_f += PSTEP;
        }
    };
    fetch_next(0);
    fetch_next(1);
    int slot = 0;
    for (int n = tid / CHUNKS; n < N; n += PSTEP) {
        const size_t pt = (size_t)b * N + n;
        float acc[CN];
#pragma unroll
        for (int i = 0; i < CN; ++i) acc[i] = 0.f;
        for (int k0 = 0; k0 < K; k0 += 4) {
            int cell[4];
            float w[4];
            if (slot == 0) {
#pragma unroll
                for (int u = 0; u < 4; ++u) { cell[u] = cell_q[0][u]; w[u] = w_q[0][u]; }
                fetch_next(0);
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u) { cell[u] = cell_q[1][u]; w[u] = w_q[1][u]; }
                fetch_next(1);
            }
            slot ^= 1;
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const uint4*>(rows + (size_t)cell[u] * CHB + chunk * CN);
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k0 + u < K) Chunk<T>::fma(acc, v[u], w[u]);   // (padding taps read cell 0's row but never enter the sum)
        }
        const int cc = chunk * CN;   // first channel of this chunk inside the block
        T* dst = out + pt * out_stride + out_c0 + c0 + cc;
        if (cc + CN <= ct && ((uintptr_t)dst % 16) == 0) {
            *reinterpret_cast<uint4*>(dst) = Chunk<T>::pack(acc);
        } else {
#pragma unroll
            for (int i = 0; i < CN; ++i)
                if (cc + i < ct) dst[i] = from_f32<T>(acc[i]);
        }
    }
}

template <typename T, typename I, int CHB>
static int launch_slab(const T* feat, long long feat_bs, int B, int C, int HW, const I* index, const float* closeness, int N, int K, T* out,
                       int out_stride, int out_c0, cudaStream_t stream) {
    const size_t smem = (size_t)HW * CHB * sizeof(T);
    cudaError_t e = kpf::set_smem(gather_slab_kernel<T, I, CHB>, smem);
    if (e != cudaSuccess) return (int)e;
    gather_slab_kernel<T, I, CHB><<<dim3((C + CHB - 1) / CHB, B), 256, smem, stream>>>(feat, feat_bs, C, HW, index, closeness, N, K, out,
                                                                                      out_stride, out_c0);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

template <typename T, typename I>
static int launch_gather(const T* feat, long long feat_bs, int B, int C, int HW, const I* index, const float* closeness, int N,
                         int K, T* out, int out_stride, int out_c0, T* rows, cudaStream_t stream) {
    // rows of 64 / 32 / 16 bytes: the widest whose [HW] slab leaves room for two CTAs per SM
    constexpr int W = 64 / (int)sizeof(T);
    const size_t cap = 96 * 1024;
    if ((size_t)HW * W * sizeof(T) <= cap)
        return launch_slab<T, I, W>(feat, feat_bs, B, C, HW, index, closeness, N, K, out, out_stride, out_c0, stream);
    if ((size_t)HW * (W / 2) * sizeof(T) <= cap)
        return launch_slab<T, I, W / 2>(feat, feat_bs, B, C, HW, index, closeness, N, K, out, out_stride, out_c0, stream);
    if ((size_t)HW * (W / 4) * sizeof(T) <= cap)
        return launch_slab<T, I, W / 4>(feat, feat_bs, B, C, HW, index, closeness, N, K, out, out_stride, out_c0, stream);
    if (rows == nullptr || ((uintptr_t)rows % 16) != 0) return KPF_ERR_BAD_ARGUMENT;   // the fallback needs its workspace
    const int Cp = (C + 7) / 8 * 8;
    dim3 g1((HW + GR_CELLS - 1) / GR_CELLS, (C + GR_CH - 1) / GR_CH, B);
    rows_kernel<T><<<g1, 256, 0, stream>>>(feat, feat_bs, C, HW, Cp, rows);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    const long long n_points = (long long)B * N, threads = n_points * (Cp / Chunk<T>::N);
    gather_rows_kernel<T, I><<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(rows, Cp, C, HW, index, closeness, n_points, N, K, out,
                                                                                   out_stride, out_c0);
    e = cudaGetLastError();
    return e == cudaSuccess ? 0 : (int)e;
}

}  // namespace kpf

extern "C" int kpf_gather_taps(const void* feat, int dtype, long long feat_batch_stride, int B, int C, int HW, const void* index,
                               int index_is_i64, const float* closeness, int N, int K, void* out, int out_stride, int out_c0,
                               void* workspace, cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(B >= 0 && C >= 1 && HW >= 1 && N >= 0 && K >= 1 && out_stride >= C + out_c0);
    if (B == 0 || N == 0) return 0;
    if (dtype == KPF_F32) {
        if (index_is_i64)
            return launch_gather<float, long long>((const float*)feat, feat_batch_stride, B, C, HW, (const long long*)index, closeness,
                                                   N, K, (float*)out, out_stride, out_c0, (float*)workspace, stream);
        return launch_gather<float, int32_t>((const float*)feat, feat_batch_stride, B, C, HW, (const int32_t*)index, closeness, N, K,
                                             (float*)out, out_stride, out_c0, (float*)workspace, stream);
    }
    if (dtype == KPF_BF16) {
        if (index_is_i64)
            return launch_gather<__nv_bfloat16, long long>((const __nv_bfloat16*)feat, feat_batch_stride, B, C, HW,
                                                           (const long long*)index, closeness, N, K, (__nv_bfloat16*)out, out_stride,
                                                           out_c0, (__nv_bfloat16*)workspace, stream);
        return launch_gather<__nv_bfloat16, int32_t>((const __nv_bfloat16*)feat, feat_batch_stride, B, C, HW, (const int32_t*)index,
                                                     closeness, N, K, (__nv_bfloat16*)out, out_stride, out_c0,
                                                     (__nv_bfloat16*)workspace, stream);
    }
    return KPF_ERR_UNSUPPORTED;
}

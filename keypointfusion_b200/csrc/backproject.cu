// K1: depth crop -> normalised point cloud.
//   reference: dataloader/loader.py:843-853 (getpcl), :874-893 (depthToPCL), :1173-1186 (resample);
//   API shell util/img2pcl.py:11-40 (Pcl_utils.getpcl).  Oracle: oracle/kpf_oracle.py getpcl/getpcl_sample.
// One CTA per sample.  Pass 1 builds the row-major ordered list of valid pixels in shared memory (a run of
// consecutive pixels per thread, 16-byte loads, one block scan of the threads' counts); pass 2 back-projects only the `sample_num` selected points in fp64
// (matching the reference's float64 numpy arithmetic) and writes them as fp32.  HBM-bound: S*S*4 B read,
// sample_num*12 B written per sample.
#include <cmath>

#include "common.cuh"

namespace kpf {

constexpr int K1_THREADS = 1024;
constexpr int K1_WARPS = K1_THREADS / 32;

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x85EBCA6Bu;
    x ^= x >> 13;
    x *= 0xC2B2AE35u;
    x ^= x >> 16;
    return x;
}

// bijection on [0,n): 4-round balanced Feistel over the next even-bit power of two + cycle walking
__device__ __forceinline__ uint32_t feistel_perm(uint32_t i, uint32_t n, uint32_t key) {
    int bits = 32 - __clz((int)(n - 1));
    if (n <= 1) bits = 0;
    if (bits < 2) bits = 2;
    bits += bits & 1;
    const int half = bits >> 1;
    const uint32_t mask = (1u << half) - 1u;
    uint32_t x = i;
    while (true) {
        uint32_t l = x >> half, r = x & mask;
#pragma unroll
        for (uint32_t rnd = 0; rnd < 4; ++rnd) {
            const uint32_t t = l ^ (mix32(r ^ key ^ (rnd * 0x9E3779B9u)) & mask);
            l = r;
            r = t;
        }
        x = (l << half) | r;
        if (x < n) return x;
    }
}

// The reference's float64 np.isclose bands, restated as float thresholds: for a float v, |double(v) - 1| <= band  <=>  lo <= v <= hi
// with lo / hi the extreme floats that satisfy the double predicate (found by testing the predicate itself on neighbouring floats),
// and |double(d)| <= 1e-8  <=>  |d| <= z.  Same decisions for every float (NaN included), no fp64 work per pixel.
struct PixelBands {
    float lo, hi, z;
};
// Evaluated ONCE on the host (the launchers below) and passed by value: every thread used to walk these nextafterf loops itself,
// ~200 of the kernel's ~1500 instructions per warp.  Host and device doubles are the same IEEE arithmetic, so the floats are the same.
static inline bool bg_pred(float v) {
    const double BG_BAND = 1e-8 + 1e-5 * 1.0;  // np.isclose(x, 1)  loader.py:844
    return fabs((double)v - 1.0) <= BG_BAND;
}
static inline PixelBands make_bands() {
    PixelBands b;
    b.hi = (float)(1.0 + (1e-8 + 1e-5));
    while (!bg_pred(b.hi)) b.hi = nextafterf(b.hi, 0.f);
    while (bg_pred(nextafterf(b.hi, 2.f))) b.hi = nextafterf(b.hi, 2.f);
    b.lo = (float)(1.0 - (1e-8 + 1e-5));
    while (!bg_pred(b.lo)) b.lo = nextafterf(b.lo, 2.f);
    while (bg_pred(nextafterf(b.lo, 0.f))) b.lo = nextafterf(b.lo, 0.f);
    b.z = (float)1e-8;
    while ((double)b.z > 1e-8) b.z = nextafterf(b.z, 0.f);
    while ((double)nextafterf(b.z, 1.f) <= 1e-8) b.z = nextafterf(b.z, 1.f);
    return b;
}
__device__ __forceinline__ bool pixel_valid(float v, float hz, float cz, const PixelBands& bands, float& dpt) {
    const bool bg = v >= bands.lo && v <= bands.hi;   // np.isclose(x, 1)  loader.py:844
    dpt = bg ? 0.0f : xadd(xmul(v, hz), cz);          // loader.py:845-847
    return !(fabsf(dpt) <= bands.z);                  // ~np.isclose(dpt, 0)  loader.py:880
}

struct BackprojCam {
    double mi[9], com[3], half[3], fx, fy, fu, fv, flip;
};

__device__ __forceinline__ void backproject_point(const BackprojCam& c, int pix, int S, float dpt, float* out, bool clamp) {
    const int r = pix / S, col = pix - r * S;
    const double u = (double)col + 0.5, v = (double)r + 0.5;  // loader.py:881
    double qx = xadd(xadd(xmul(c.mi[0], u), xmul(c.mi[1], v)), c.mi[2]);
    double qy = xadd(xadd(xmul(c.mi[3], u), xmul(c.mi[4], v)), c.mi[5]);
    const double qz = xadd(xadd(xmul(c.mi[6], u), xmul(c.mi[7], v)), c.mi[8]);
    qx = xdiv(qx, qz);
    qy = xdiv(qy, qz);
    const double d = (double)dpt;
    const double x = xmul(xdiv(xsub(qx, c.fu), c.fx), d);               // loader.py:889
    const double y = xmul(xdiv(xmul(c.flip, xsub(qy, c.fv)), c.fy), d);  // loader.py:890
    float ox = (float)xdiv(xsub(x, c.com[0]), c.half[0]);                 // loader.py:849-852
    float oy = (float)xdiv(xsub(y, c.com[1]), c.half[1]);
    float oz = (float)xdiv(xsub(d, c.com[2]), c.half[2]);
    if (clamp) {
        ox = fminf(fmaxf(ox, -1.f), 1.f);
        oy = fminf(fmaxf(oy, -1.f), 1.f);
        oz = fminf(fmaxf(oz, -1.f), 1.f);
    }
    out[0] = ox;
    out[1] = oy;
    out[2] = oz;
}

// MODE 0: fixed-size sample [B,sample_num,3];  MODE 1: every valid point, ordered, [B,S*S,3] + pixel ids
template <int MODE>
__global__ void __launch_bounds__(K1_THREADS, 1)
backproject_kernel(const float* __restrict__ img, const float* __restrict__ com3D, const float* __restrict__ cube,
                   const float* __restrict__ M, const float* __restrict__ cam, int S, int sample_num,
                   const int32_t* __restrict__ ranks, uint32_t seed, int clamp, float flip, const PixelBands bands,
                   float* __restrict__ pcl_out, int32_t* __restrict__ pix_out, int32_t* __restrict__ count_out) {
    extern __shared__ __align__(16) unsigned char k1_smem[];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int npix = S * S;
    uint16_t* list = reinterpret_cast<uint16_t*>(k1_smem);                               // [npix]
    __shared__ uint32_t warp_tot[K1_WARPS], total;
    __shared__ BackprojCam sc;

    const float hz = xdiv(cube[3 * b + 2], 2.0f), cz = com3D[3 * b + 2];
    const float* im = img + (size_t)b * npix;

    if (tid == 0) {
        inv3x3_f64(M + 9 * b, sc.mi);
        for (int k = 0; k < 3; ++k) {
            sc.com[k] = (double)com3D[3 * b + k];
            sc.half[k] = xdiv((double)cube[3 * b + k], 2.0);
        }
        sc.fx = cam[4 * b + 0];
        sc.fy = cam[4 * b + 1];
        sc.fu = cam[4 * b + 2];
        sc.fv = cam[4 * b + 3];
        sc.flip = (double)flip;
    }

    // ---- pass 1: a thread owns a run of `ppt` consecutive pixels (<= 64: S <= 256), so row-major order = thread order and the ordered
    //      compaction needs ONE exclusive scan over the threads' counts (the round-robin assignment it replaces took a ballot + popc
    //      per pixel round twice over: ~650 of the kernel's ~1500 instructions per warp, profiles/stalls_r2_final.txt)
    const int ppt = (npix + K1_THREADS - 1) / K1_THREADS;
    const int p0 = tid * ppt;
    uint64_t vbits = 0;
    const bool vec = (ppt & 3) == 0 && (npix & 3) == 0 && (((uintptr_t)im) & 15) == 0;
    if (vec) {
        for (int i0 = 0; i0 < ppt; i0 += 16) {   // up to four 16-byte loads in flight
            float4 vv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int p = p0 + i0 + 4 * u;
                vv[u] = (i0 + 4 * u < ppt && p < npix) ? __ldg(reinterpret_cast<const float4*>(im + p)) : make_float4(1.f, 1.f, 1.f, 1.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float f[4] = {vv[u].x, vv[u].y, vv[u].z, vv[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float dpt;
                    const int i = i0 + 4 * u + e;
                    const bool ok = i < ppt && p0 + i < npix && pixel_valid(f[e], hz, cz, bands, dpt);   // padding = 1.0 = background
                    vbits |= (uint64_t)ok << i;
                }
            }
        }
    } else {
        for (int i = 0; i < ppt; ++i) {
            float dpt;
            const bool ok = p0 + i < npix && pixel_valid(__ldg(im + p0 + i), hz, cz, bands, dpt);
            vbits |= (uint64_t)ok << i;
        }
    }
    const uint32_t mine = (uint32_t)__popcll(vbits);
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_tot[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        warp_tot[lane] = wi - w;  // exclusive warp offsets
        if (lane == 31) total = wi;
    }
    __syncthreads();
    const int P = (int)total;
    {
        uint32_t at = warp_tot[warp] + incl - mine;
        uint64_t bits = vbits;
        while (bits) {   // the thread's valid pixels in ascending order
            const int i = __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            list[at++] = (uint16_t)(p0 + i);
        }
    }
    __syncthreads();
    if (tid == 0 && count_out) count_out[b] = P;

    // ---- pass 2
    if (MODE == 0) {
        const uint32_t key = mix32(seed ^ mix32((uint32_t)b + 0x85EBCA6Bu));
        for (int j = tid; j < sample_num; j += K1_THREADS) {
            float* o = pcl_out + ((size_t)b * sample_num + j) * 3;
            if (P == 0) {  // loader.py:1176-1177
                o[0] = o[1] = o[2] = 0.f;
                continue;
            }
            int rank;
            if (ranks) {
                rank = ranks[(size_t)b * sample_num + j];
                rank = rank < 0 ? 0 : (rank > P - 1 ? P - 1 : rank);
            } else if (P >= sample_num) {  // loader.py:1184: random subset, random order
                rank = (int)feistel_perm((uint32_t)j, (uint32_t)P, key);
            } else {  // loader.py:1179-1183: floor(n/P) copies of each index + distinct random remainder, shuffled
                const uint32_t tmp = (uint32_t)sample_num / (uint32_t)P;
                const uint32_t t = feistel_perm((uint32_t)j, (uint32_t)sample_num, key ^ 0x1234567u);
                rank = t < tmp * (uint32_t)P ? (int)(t / tmp) : (int)feistel_perm(t - tmp * (uint32_t)P, (uint32_t)P, key);
            }
            const int pix = list[rank];
            float dpt;
            pixel_valid(__ldg(im + pix), hz, cz, bands, dpt);
            backproject_point(sc, pix, S, dpt, o, clamp != 0);
        }
    } else {
        for (int r = tid; r < npix; r += K1_THREADS) {
            float* o = pcl_out + ((size_t)b * npix + r) * 3;
            if (r < P) {
                const int pix = list[r];
                float dpt;
                pixel_valid(__ldg(im + pix), hz, cz, bands, dpt);
                backproject_point(sc, pix, S, dpt, o, clamp != 0);
                if (pix_out) pix_out[(size_t)b * npix + r] = pix;
            } else {
                o[0] = o[1] = o[2] = 0.f;
                if (pix_out) pix_out[(size_t)b * npix + r] = -1;
            }
        }
    }
}

static size_t k1_smem_bytes(int S) { return (((size_t)S * S * 2 + 15) & ~(size_t)15); }

}  // namespace kpf

extern "C" int kpf_getpcl(const float* img, const float* com3D, const float* cube, const float* M, const float* cam, int B,
                          int S, int sample_num, const int32_t* ranks, uint32_t seed, int clamp, float flip, float* pcl_out,
                          int32_t* count_out, cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(B >= 0 && S >= 1 && S <= 256 && sample_num >= 1);
    if (B == 0) return 0;
    const size_t smem = k1_smem_bytes(S);
    cudaError_t e = kpf::set_smem(backproject_kernel<0>, smem);
    if (e != cudaSuccess) return (int)e;
    static const PixelBands bands = make_bands();
    backproject_kernel<0><<<B, K1_THREADS, smem, stream>>>(img, com3D, cube, M, cam, S, sample_num, ranks, seed, clamp, flip, bands,
                                                          pcl_out, nullptr, count_out);
    KPF_CHECK_LAUNCH();
    return 0;
}

extern "C" int kpf_backproject_all(const float* img, const float* com3D, const float* cube, const float* M, const float* cam,
                                   int B, int S, float flip, float* xyz_out, int32_t* pix_out, int32_t* count_out,
                                   cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(B >= 0 && S >= 1 && S <= 256);
    if (B == 0) return 0;
    const size_t smem = k1_smem_bytes(S);
    cudaError_t e = kpf::set_smem(backproject_kernel<1>, smem);
    if (e != cudaSuccess) return (int)e;
    static const PixelBands bands = make_bands();
    backproject_kernel<1><<<B, K1_THREADS, smem, stream>>>(img, com3D, cube, M, cam, S, 0, nullptr, 0u, 0, flip, bands, xyz_out, pix_out,
                                                          count_out);
    KPF_CHECK_LAUNCH();
    return 0;
}

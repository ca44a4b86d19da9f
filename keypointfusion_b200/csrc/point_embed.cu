// Point stage of Block_KPFusion on tensor cores (model/model.py:295-320):
//   K4b pcl_joint2offset + K3 4-tap gathers + the four folded Conv1d+BN point embeddings + relu/relu
//   + the softmax-over-points aggregation numerators, in ONE persistent kernel; nothing of [B,N,*] but the final
//   point features e (bf16) ever reaches HBM.
//
//   kpf_repack_features : NCHW (img_feat | img_feat_rgb | img_offset[4J:]) -> channels-last bf16 rows [B,HW,288]
//                         so that one tap of a point is ONE contiguous 576-byte row.
//   kpf_point_embed     : per 128-point tile (512 threads: 4 per point in the gather, 4 column groups per row in the epilogues):
//       A1[128 x 256] = [gather(img_feat) | gather(weight map) | (unit offset xyz, closeness) per joint, xyz]   (bf16, smem)
//       A2[128 x 128] =  gather(img_feat_rgb)
//       e  = relu( relu(A1 W1^T + b1) + A2 W2^T + b2 )                     tcgen05, fp32 accumulators in TMEM
//       p  = exp(w - max_tile w)  (softmax numerators of the gathered weight map, per joint)
//       D[c][j] = sum_n e[n][c] p[n][j]                                    tcgen05 with MN-major operands
//     outputs: e [B,N,128] bf16, and per tile (D, max, sum) partials that the DESA kernel combines flash-style.
//   Weights (96 KB bf16) stay resident in shared memory; CTAs are persistent over tiles.
#include "tmem_ldst.cuh"

namespace kpf {

constexpr int PE_CP = 288;   // channels per repacked row: 128 depth-branch + 128 rgb-branch + 32 (J weight channels, zero padded)
constexpr int PE_CH = PE_CP / 8;

template <typename T>
__global__ void __launch_bounds__(256)
repack_kernel(const T* __restrict__ f_d, const T* __restrict__ f_rgb, const T* __restrict__ f_w, long long w_bs, int C, int J, int HW,
              __nv_bfloat16* __restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, c0 = blockIdx.y * 32, h0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, h = h0 + tx;
        float v = 0.f;
        if (h < HW) {
            if (c < C) v = to_f32(f_d[((size_t)b * C + c) * HW + h]);
            else if (c < 2 * C) v = to_f32(f_rgb[((size_t)b * C + (c - C)) * HW + h]);
            else if (c - 2 * C < J) v = to_f32(f_w[(size_t)b * w_bs + (size_t)(c - 2 * C) * HW + h]);
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int h = h0 + r, c = c0 + tx;
        if (h < HW && c < PE_CP) out[((size_t)b * HW + h) * PE_CP + c] = __float2bfloat16_rn(tile[tx][r]);
    }
}

// bf16 fast path of the repack (C = 128, HW % 128 == 0): a CTA transposes a [64 channels x 128 cells] block through shared memory
// with 16-byte global accesses on both sides (8 cells of a channel in, 8 channels of a cell out) -- the generic kernel above
// moves 2 bytes per lane.  Channel blocks 0,1 = depth branch, 2,3 = rgb branch, 4 = the J weight channels zero-padded to 32.
constexpr int RP_LD = 128 + 2;   // bf16 elements per staged channel row (+2: consecutive channels land in consecutive banks)
__global__ void __launch_bounds__(256)
repack_bf16_kernel(const __nv_bfloat16* __restrict__ f_d, const __nv_bfloat16* __restrict__ f_rgb, const __nv_bfloat16* __restrict__ f_w,
                   long long w_bs, int J, int HW, __nv_bfloat16* __restrict__ out) {
    __shared__ __align__(16) __nv_bfloat16 tile[64 * RP_LD];
    const int b = blockIdx.z, cb = blockIdx.y, h0 = blockIdx.x * 128, tid = threadIdx.x;
    const int nch = cb < 4 ? 64 : 32;
    // load: thread -> (channel, 8 consecutive cells); a warp reads 2 channels x 256 B
    for (int i = tid; i < nch * 16; i += 256) {
        const int ch = i >> 4, seg = i & 15;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (cb < 2) v = __ldg(reinterpret_cast<const uint4*>(f_d + ((size_t)b * 128 + cb * 64 + ch) * HW + h0) + seg);
        else if (cb < 4) v = __ldg(reinterpret_cast<const uint4*>(f_rgb + ((size_t)b * 128 + (cb - 2) * 64 + ch) * HW + h0) + seg);
        else if (ch < J) v = __ldg(reinterpret_cast<const uint4*>(f_w + (size_t)b * w_bs + (size_t)ch * HW + h0) + seg);
        uint32_t* dst = reinterpret_cast<uint32_t*>(tile + ch * RP_LD + seg * 8);   // 4-byte aligned (RP_LD even)
        dst[0] = v.x;
        dst[1] = v.y;
        dst[2] = v.z;
        dst[3] = v.w;
    }
    __syncthreads();
    // store: thread -> (cell, 8 consecutive channels); a warp writes 128 contiguous bytes of each of 4 (or 8) cell rows
    const int gs = cb < 4 ? 3 : 2, ngrp = 1 << gs;   // 8-channel groups per cell
    for (int i = tid; i < 128 * ngrp; i += 256) {
        const int g = i & (ngrp - 1), cell = i >> gs;
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint16_t lo = *reinterpret_cast<const uint16_t*>(tile + (8 * g + 2 * k) * RP_LD + cell);
            const uint16_t hi = *reinterpret_cast<const uint16_t*>(tile + (8 * g + 2 * k + 1) * RP_LD + cell);
            w[k] = (uint32_t)lo | ((uint32_t)hi << 16);
        }
        *reinterpret_cast<uint4*>(out + ((size_t)b * HW + h0 + cell) * PE_CP + cb * 64 + 8 * g) = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

struct PointParams {
    const uint4* featT;      // [B,HW,36] uint4 (288 bf16)
    const int32_t* idx;      // [B,N,4]
    const float* clos;       // [B,N,4]
    const float* pcl;        // [B,N,3]
    const float* joint;      // [B,J,3]
    const int32_t* order;    // [B,N] processing order of the points (kpf_spatial_order) or null = identity
    const uint4* wmat;       // W1a, W1b, W2: 3 x [16][128] uint4
    const float* wvec;       // b1[128], b2[128]
    __nv_bfloat16* e_out;    // [B,N,128] with batch stride e_bs elements (>= N*128: DESA appends its joint rows behind the points)
    long long e_bs;
    float* part_acc;         // [B,T,128,32]
    float* part_ms;          // [B,T,2,32]  (max, sum)
    int B, N, J, HW;
    float kernel_size;
    long long* dbg;
};

__device__ __forceinline__ void bf16x8_fma(float* acc, const uint4& v, float w) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 f = __bfloat1622float2(h[i]);
        acc[2 * i] += f.x * w;
        acc[2 * i + 1] += f.y * w;
    }
}

constexpr int PE_NT = 512;   // gather: warp = 8 points x 4 chunk lanes; epilogues: 4 lane quarters x 4 column groups

__global__ void __launch_bounds__(PE_NT, 1) point_embed_kernel(const PointParams p) {
    extern __shared__ __align__(128) unsigned char pe_smem[];
    uint4* sW = reinterpret_cast<uint4*>(pe_smem);   // [3][2048]
    uint4* sA1 = sW + 3 * 2048;                       // K-major [16 row groups][32 k-chunks][8 rows] (K = 256); after the MMAs: sP [16][4][8]
    uint4* sA2 = sA1 + 4096;                          // K-major [16][16][8] (K = 128) = e^T MN-major [16][16][8] after the MMAs
    float* sJ = reinterpret_cast<float*>(sA2 + 2048); // [32][4] joints of the current sample
    float* sRed = sJ + 128;                           // [32] per-joint tile maxima
    float* sB = sRed + 128;                           // b1[128], b2[128]
    float* sT = sB + 256;                             // [21][129] transposed softmax scratch
    int* sN = reinterpret_cast<int*>(sT + 21 * 129 + 3);  // [2][128] point ids of the tile in flight / being prefetched
    __shared__ __align__(8) uint64_t wbar, mma_bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int warp_u = warp_index_uniform();  // MMA issue: one elected lane of warp 0 from warp-uniform code (umma.cuh)
    // gather / offsets: a warp takes 8 points; lane = (point p8, sub): the 4 `sub` lanes of a point read 64 contiguous bytes of a
    // tap row per load (8 L1 wavefronts per warp load instead of 32 with one lane per row) and the 8 points of a warp fill the
    // 8 rows x 16 B core matrices of the operand, so the stores are conflict free
    const int r = 8 * warp + (lane & 7), sub = lane >> 3;
    const int q = warp & 3, cg = warp >> 2, row = 32 * q + lane; // epilogues: TMEM lane `row`, columns [32cg, 32cg + 32)
    const int J = p.J, N = p.N, T = N / 128;

    pdl_launch_dependents();
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        mbar_init(&wbar, 1);
        mbar_init(&mma_bar, 1);
        fence_mbar_init();
        mbar_expect_tx(&wbar, 3 * 2048 * 16);
        tma_bulk_g2s(sW, p.wmat, 3 * 2048 * 16, &wbar);
    }
    for (int i = tid; i < 256; i += PE_NT) sB[i] = p.wvec[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = tmem_slot, tmem = tmem0 + ((uint32_t)(32 * q) << 16);
    const uint32_t ACC1 = 0, ACC2 = 128, ACC3 = 256;
    uint32_t phase = 0;
    bool w_ready = false;
    const float inv_ks = 1.f / p.kernel_size;
    int n_stamp = 0;
    auto stamp = [&]() {
        if (p.dbg && blockIdx.x == 0 && tid == 0 && n_stamp < 64) p.dbg[n_stamp] = clock64();
        ++n_stamp;
    };
    stamp();

    // per-point inputs of the tile about to be gathered (prefetched during the previous tile's MMAs)
    int4 id = make_int4(0, 0, 0, 0);
    float4 cw = make_float4(0.f, 0.f, 0.f, 0.f);
    float px = 0.f, py = 0.f, pz = 0.f;
    int n_id = 0, tile_par = 0;
    auto fetch_point = [&](int tile, int par) {
        const int b = tile / T, t = tile - b * T;
        n_id = p.order ? __ldg(p.order + (size_t)b * N + t * 128 + r) : t * 128 + r;   // tile = 128 consecutive points of the order
        if (sub == 0) sN[par * 128 + r] = n_id;
        const size_t pn = (size_t)b * N + n_id;
        id = __ldg(reinterpret_cast<const int4*>(p.idx + pn * 4));
        cw = __ldg(reinterpret_cast<const float4*>(p.clos + pn * 4));
        px = __ldg(p.pcl + pn * 3);
        py = __ldg(p.pcl + pn * 3 + 1);
        pz = __ldg(p.pcl + pn * 3 + 2);
    };
    pdl_wait();   // weights only so far; points, indices and the repacked maps come from the previous kernels
    if ((int)blockIdx.x < p.B * T) fetch_point(blockIdx.x, 0);

    for (int tile = blockIdx.x; tile < p.B * T; tile += gridDim.x) {
        const int b = tile / T, t = tile - b * T;
        __syncthreads();  // previous tile's readers of sJ / sRed / sT are done (its MMAs were waited for)
        if (tid < J) {
            sJ[4 * tid] = p.joint[((size_t)b * J + tid) * 3];
            sJ[4 * tid + 1] = p.joint[((size_t)b * J + tid) * 3 + 1];
            sJ[4 * tid + 2] = p.joint[((size_t)b * J + tid) * 3 + 2];
        }
        const uint4* r0 = p.featT + ((size_t)b * p.HW + id.x) * PE_CH;
        const uint4* r1 = p.featT + ((size_t)b * p.HW + id.y) * PE_CH;
        const uint4* r2 = p.featT + ((size_t)b * p.HW + id.z) * PE_CH;
        const uint4* r3 = p.featT + ((size_t)b * p.HW + id.w) * PE_CH;
        stamp();
        // ---- K3: 4-tap gathers: chunk 4i + sub of the 36-chunk row in iteration i (depth 0-15 -> A1, rgb 16-31 -> A2, weight map
        //      32-35 -> A1 chunks 16-19 and the softmax), three iterations (12 x 16 B) in flight
        float wraw[8];   // gathered weight-map channels [8 sub, 8 sub + 8) of this point
        uint4* a1r = sA1 + (r >> 3) * 256 + (r & 7);
        uint4* a2r = sA2 + (r >> 3) * 128 + (r & 7);
#pragma unroll
        for (int i0 = 0; i0 < 9; i0 += 3) {
            uint4 v[3][4];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int cc = 4 * (i0 + u) + sub;
                v[u][0] = __ldg(r0 + cc);
                v[u][1] = __ldg(r1 + cc);
                v[u][2] = __ldg(r2 + cc);
                v[u][3] = __ldg(r3 + cc);
            }
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                bf16x8_fma(acc, v[u][0], cw.x);
                bf16x8_fma(acc, v[u][1], cw.y);
                bf16x8_fma(acc, v[u][2], cw.z);
                bf16x8_fma(acc, v[u][3], cw.w);
                const int i = i0 + u;   // compile-time after unrolling
                if (i < 4) {
                    a1r[(4 * i + sub) * 8] = pack8_bf16(acc);
                } else if (i < 8) {
                    a2r[(4 * (i - 4) + sub) * 8] = pack8_bf16(acc);
                } else {
                    a1r[(16 + sub) * 8] = pack8_bf16(acc);
#pragma unroll
                    for (int k = 0; k < 8; ++k) wraw[k] = acc[k];
                }
            }
        }
        // softmax over the tile's points, step 1: the gathered weights, transposed, for the per-joint maxima
#pragma unroll
        for (int k = 0; k < 8; ++k)
            if (8 * sub + k < J) sT[(8 * sub + k) * 129 + r] = wraw[k];
        stamp();
        __syncthreads();  // sJ, sT visible
        // ---- K4b: [unit offset xyz, closeness] of joints [6sub, 6sub + 6), then xyz -> chunks 20 + 3sub .. of A1 (ops.pack_point_embed
        //      orders W1's columns to match)
        {
            float buf[24];
#pragma unroll
            for (int jj = 0; jj < 6; ++jj) {
                const int j = 6 * sub + jj;
                float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
                if (j < J) {
                    const float ox = sJ[4 * j] - px, oy = sJ[4 * j + 1] - py, oz = sJ[4 * j + 2] - pz;
                    const float d2 = ox * ox + oy * oy + oz * oz;
                    const float dis = d2 * rsqrtf(fmaxf(d2, 1e-30f));       // bf16 operand: approximate sqrt / divide are ample
                    const float inv = __fdividef(1.f, dis + 1e-8f);
                    const float heat = (p.kernel_size - dis) * inv_ks;
                    const float msk = (heat >= 0.f && pz < 0.99f) ? 1.f : 0.f;
                    o0 = ox * inv * msk;
                    o1 = oy * inv * msk;
                    o2 = oz * inv * msk;
                    o3 = heat * msk;
                } else if (j == J) {
                    o0 = px;
                    o1 = py;
                    o2 = pz;
                }
                buf[4 * jj] = o0;
                buf[4 * jj + 1] = o1;
                buf[4 * jj + 2] = o2;
                buf[4 * jj + 3] = o3;
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) a1r[(20 + 3 * sub + c) * 8] = pack8_bf16(buf + 8 * c);
        }
        // softmax step 1b: per-joint maximum, 16 threads per joint
        {
            const int rj = tid >> 4, rs = tid & 15;
            float m = -INFINITY;
            if (rj < J) {
#pragma unroll
                for (int i = 0; i < 8; ++i) m = fmaxf(m, sT[rj * 129 + rs + 16 * i]);
            }
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if (rs == 0 && rj < 32) sRed[rj] = rj < J ? m : -INFINITY;
        }
        stamp();
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (warp_u == 0) {
            tc_fence_after();
            if (!w_ready) mbar_wait(&wbar, 0);
            if (elect_one()) {
                const uint32_t id128 = umma_idesc_bf16(128, 128, false, false);
                // A operands: 128 B between k-chunks, 4096 / 2048 B between 8-row groups
                umma_gemm(tmem0 + ACC1, smem_u32(sA1), 128, 4096, smem_u32(sW), 2048, 128, id128, 128, false);
                umma_gemm(tmem0 + ACC1, smem_u32(sA1 + 128), 128, 4096, smem_u32(sW + 2048), 2048, 128, id128, 128, true);
                umma_gemm(tmem0 + ACC2, smem_u32(sA2), 128, 2048, smem_u32(sW + 4096), 2048, 128, id128, 128, false);
                umma_commit(&mma_bar);
            }
            __syncwarp();
        }
        w_ready = true;
        // ---- while the MMAs run: next tile's point inputs, and the softmax numerators p = exp(w - max) (bf16-rounded, as the MMA
        //      will see them) with their per-joint sums
        if (tile + (int)gridDim.x < p.B * T) fetch_point(tile + gridDim.x, tile_par ^ 1);
        float* ms = p.part_ms + ((size_t)b * T + t) * 64;
        float pj[8];   // joints [8 sub, 8 sub + 8) of this point
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int j = 8 * sub + k;
            pj[k] = j < J ? __bfloat162float(__float2bfloat16_rn(__expf(wraw[k] - sRed[j]))) : 0.f;
            if (j < J) sT[j * 129 + r] = pj[k];
        }
        if (tid < 32) ms[tid] = sRed[tid];
        __syncthreads();
        {
            const int rj = tid >> 4, rs = tid & 15;
            float sm_ = 0.f;
            if (rj < J) {
#pragma unroll
                for (int i = 0; i < 8; ++i) sm_ += sT[rj * 129 + rs + 16 * i];
            }
#pragma unroll
            for (int o = 1; o < 16; o <<= 1) sm_ += __shfl_xor_sync(0xffffffffu, sm_, o);
            if (rs == 0 && rj < 32) ms[32 + rj] = rj < J ? sm_ : 0.f;
        }
        stamp();
        mbar_wait(&mma_bar, phase);
        phase ^= 1;
        tc_fence_after();
        stamp();
        // ---- epilogue: e = relu(relu(acc1 + b1) + acc2 + b2) -> global (bf16) and MN-major A operand (sA2 region)
        {
            __nv_bfloat16* eo = p.e_out + (size_t)b * p.e_bs + (size_t)sN[tile_par * 128 + row] * 128 + 32 * cg;
            float a[32], rr[32];
            tmem_ld_nw<32>(tmem + ACC1 + 32 * cg, a);
            tmem_ld_nw<32>(tmem + ACC2 + 32 * cg, rr);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = fmaxf(fmaxf(a[i] + sB[32 * cg + i], 0.f) + rr[i] + sB[128 + 32 * cg + i], 0.f);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const uint4 v = pack8_bf16(a + 8 * c);
                *reinterpret_cast<uint4*>(eo + 8 * c) = v;
                sA2[(row >> 3) * 128 + (4 * cg + c) * 8 + (row & 7)] = v;  // e^T: M = channel contiguous
            }
        }
        // p as MN-major B operand [K = 128 points][N = 32 joints] over the (dead) head of sA1
        sA1[(r >> 3) * 32 + sub * 8 + (r & 7)] = pack8_bf16(pj);
        stamp();
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (warp_u == 0) {
            tc_fence_after();
            if (elect_one()) {
                umma_gemm(tmem0 + ACC3, smem_u32(sA2), 2048, 128, smem_u32(sA1), 512, 128, umma_idesc_bf16(128, 32, true, true), 128, false);
                umma_commit(&mma_bar);
            }
            __syncwarp();
        }
        mbar_wait(&mma_bar, phase);
        phase ^= 1;
        tc_fence_after();
        {   // D[channel = row][joints 8cg .. 8cg + 8)
            float a[8];
            tmem_ld<8>(tmem + ACC3 + 8 * cg, a);
            float4* o = reinterpret_cast<float4*>(p.part_acc + (((size_t)b * T + t) * 128 + row) * 32 + 8 * cg);
            o[0] = make_float4(a[0], a[1], a[2], a[3]);
            o[1] = make_float4(a[4], a[5], a[6], a[7]);
        }
        tc_fence_before();
        tile_par ^= 1;
        stamp();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem0, 512);
}

constexpr size_t PE_SMEM = (size_t)(3 * 2048 + 4096 + 2048) * 16 + (128 + 128 + 256 + 21 * 129 + 3 + 256) * 4;

}  // namespace kpf

extern "C" int kpf_repack_features(const void* f_d, const void* f_rgb, const void* f_w, long long w_batch_stride, int dtype, int B, int C,
                                   int J, int HW, void* out, cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(B >= 0 && C == 128 && J >= 1 && J <= 32 && HW >= 1);
    if (B == 0) return 0;
    dim3 grid((HW + 31) / 32, PE_CP / 32, B);
    if (dtype == KPF_F32) {
        kpf::set_smem(repack_kernel<float>, 0);
        repack_kernel<float><<<grid, 256, 0, stream>>>((const float*)f_d, (const float*)f_rgb, (const float*)f_w, w_batch_stride, C, J, HW,
                                                      (__nv_bfloat16*)out);
    } else if (dtype == KPF_BF16 && HW % 128 == 0 && ((uintptr_t)f_d % 16) == 0 && ((uintptr_t)f_rgb % 16) == 0 && ((uintptr_t)f_w % 16) == 0 &&
               w_batch_stride % 8 == 0) {
        kpf::set_smem(repack_bf16_kernel, 0);
        repack_bf16_kernel<<<dim3(HW / 128, 5, B), 256, 0, stream>>>((const __nv_bfloat16*)f_d, (const __nv_bfloat16*)f_rgb,
                                                                    (const __nv_bfloat16*)f_w, w_batch_stride, J, HW, (__nv_bfloat16*)out);
    } else if (dtype == KPF_BF16) {
        kpf::set_smem(repack_kernel<__nv_bfloat16>, 0);
        repack_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)f_d, (const __nv_bfloat16*)f_rgb,
                                                              (const __nv_bfloat16*)f_w, w_batch_stride, C, J, HW, (__nv_bfloat16*)out);
    } else {
        return KPF_ERR_UNSUPPORTED;
    }
    KPF_CHECK_LAUNCH();
    return 0;
}

extern "C" int kpf_point_embed(const void* featT, const int32_t* idx, const float* clos, const float* pcl, const float* joint,
                               const int32_t* order, const void* wmat, const float* wvec, int B, int N, int J, int HW, float kernel_size, void* e_out,
                               long long e_batch_stride,
                               float* part_acc, float* part_ms, int num_sms, long long* dbg, cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(B >= 0 && N >= 128 && N % 128 == 0 && J >= 1 && J <= 21 && HW >= 1 && num_sms >= 1);
    KPF_REQUIRE(((uintptr_t)featT % 16) == 0 && ((uintptr_t)wmat % 16) == 0 && ((uintptr_t)idx % 16) == 0 && ((uintptr_t)clos % 16) == 0);
    if (B == 0) return 0;
    PointParams p;
    p.featT = (const uint4*)featT; p.idx = idx; p.clos = clos; p.pcl = pcl; p.joint = joint; p.order = order; p.wmat = (const uint4*)wmat; p.wvec = wvec;
    KPF_REQUIRE(e_batch_stride >= (long long)N * 128 && e_batch_stride % 8 == 0);
    p.e_out = (__nv_bfloat16*)e_out; p.e_bs = e_batch_stride; p.part_acc = part_acc; p.part_ms = part_ms; p.B = B; p.N = N; p.J = J; p.HW = HW;
    p.kernel_size = kernel_size;
    p.dbg = dbg;
    cudaError_t e = kpf::set_smem(point_embed_kernel, PE_SMEM);
    if (e != cudaSuccess) return (int)e;
    const int tiles = B * (N / 128);
    e = kpf::launch_pdl(point_embed_kernel, dim3(tiles < num_sms ? tiles : num_sms), dim3(PE_NT), PE_SMEM, stream, p);
    if (e != cudaSuccess) return (int)e;
    KPF_CHECK_LAUNCH();
    return 0;
}

// DESA on tensor cores (model/model.py:129-204 + the joint embeddings :323-325), SURVEY.md 8f-1.
// One CTA per (sample, scale); thread c owns OUTPUT CHANNEL c (TMEM lane c), so every GEMM is computed transposed --
// D[c][row] = sum_k W[c][k] X[row][k] with the (BatchNorm-folded) weight as the M=128 A operand and the activations as the
// N operand -- which turns DESA's max-pool over the 64 grouped points of a joint into a per-thread register reduction.
//
//   prologue  combine the point stage's softmax partials (flash-style) -> joint_agg[J][128]
//             jf = relu(Wj [joint_agg | joint_xyz] + b)                     (model.py:323-325)   tcgen05 + fp32 xyz term
//             ball query (pointnet2_ops semantics) of the J joints over the N points + the J joints themselves
//   per tile  (2 joints x 64 grouped points = 128 rows):
//             X = [feat[idx] - jf[j] | (xyz[idx] - c_j)/r]  ->  h = relu(W1 X + b1)  ->  relu(W2 h + b2)  -> max over 64
//   output    desa_part[b][scale][j][:]  (+ jf[b][j][:] from the scale-0 CTA); the 512->128 fusion conv follows.
#include "umma.cuh"

namespace kpf {

struct DesaParams {
    const __nv_bfloat16* e;   // [B,N,128] point features (kpf_point_embed)
    const float* part_acc;    // [B,T,128,32]
    const float* part_ms;     // [B,T,2,32]
    const float* pcl;         // [B,N,3]
    const float* joint;       // [B,J,3]
    const uint4* wmat;        // Wj [16][128] ; per scale: W1 main [16][128], W1 tail [2][128], W2 [16][128]
    const float* wvec;        // bj[128], Wjx[128][4] ; per scale: b1[128], b2[128]
    float* desa_part;         // [B,S,J,128]
    float* jf_out;            // [B,J,128]
    int B, N, J, T, S, nsample;
    float radius[4];
    long long* dbg;
};

constexpr int DS_MAT_PER_SCALE = 2048 + 256 + 2048;

__global__ void __launch_bounds__(128, 1) desa_fused_kernel(const DesaParams p) {
    extern __shared__ __align__(128) unsigned char ds_smem[];
    uint4* sW1 = reinterpret_cast<uint4*>(ds_smem);   // [16][128] + tail [2][128]
    uint4* sW1t = sW1 + 2048;
    uint4* sW2 = sW1t + 256;                           // [16][128]
    uint4* sX = sW2 + 2048;                            // [16][128] K-major activations; prologue: Wj
    uint4* sXt = sX + 2048;                            // [2][128]
    uint4* sH = sXt + 256;                             // MN-major [16][16][8]; prologue: joint_agg operand [16][4][8]
    float4* sPcl = reinterpret_cast<float4*>(sH + 2048);   // [N + J] xyz
    float* sJF = reinterpret_cast<float*>(sPcl + (p.N + p.J + 3) / 4 * 4);  // [J][128] fp32 joint features
    float* sOut = sJF + p.J * 128;                     // [J][128]
    float* sMS = sOut + p.J * 128;                     // [T][2][32] partial max/sum, then [T][32] scale factors + den[32]
    uint32_t* sMask = reinterpret_cast<uint32_t*>(sMS + p.T * 64 + 64);  // [J][(N+J+31)/32] ball-query hit words (bit = point)
    uint16_t* sIdx = reinterpret_cast<uint16_t*>(sMask + ((p.J * ((p.N + p.J + 31) / 32) + 3) / 4 * 4));  // [J][nsample]
    __shared__ __align__(8) uint64_t wbar[2], mma_bar;
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int warp_u = warp_index_uniform();  // MMA issue: one elected lane of warp 0 from warp-uniform code (umma.cuh)
    const int b = blockIdx.x / p.S, sc = blockIdx.x - b * p.S;
    const int J = p.J, N = p.N, T = p.T, NS = p.nsample;
    const float radius = p.radius[sc];
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const uint32_t ACCE = 0, ACC1 = 128, ACC2 = 256;
    int n_stamp = 0;
    auto stamp = [&]() {
        if (p.dbg && blockIdx.x == 0 && tid == 0 && n_stamp < 64) p.dbg[n_stamp] = clock64();
        ++n_stamp;
    };
    stamp();

    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        mbar_init(&wbar[0], 1);
        mbar_init(&wbar[1], 1);
        mbar_init(&mma_bar, 1);
        fence_mbar_init();
        mbar_expect_tx(&wbar[0], 2048 * 16);                      // Wj -> sX
        tma_bulk_g2s(sX, p.wmat, 2048 * 16, &wbar[0]);
        const uint4* ws = p.wmat + 2048 + (size_t)sc * DS_MAT_PER_SCALE;
        mbar_expect_tx(&wbar[1], DS_MAT_PER_SCALE * 16);          // W1 (+tail), W2 of this scale
        tma_bulk_g2s(sW1, ws, (2048 + 256) * 16, &wbar[1]);
        tma_bulk_g2s(sW2, ws + 2048 + 256, 2048 * 16, &wbar[1]);
    }
    // ---- stage xyz of the point set (N points + J joints) and the partial (max, sum) table
    for (int i = tid; i < N + J; i += 128) {
        const float* s = i < N ? p.pcl + ((size_t)b * N + i) * 3 : p.joint + ((size_t)b * J + (i - N)) * 3;
        sPcl[i] = make_float4(s[0], s[1], s[2], 0.f);
    }
    for (int i = tid; i < T * 64; i += 128) sMS[i] = p.part_ms[(size_t)b * T * 64 + i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = tmem_slot, tmem = tmem0 + lane_off;
    uint32_t phase = 0;
    stamp();
    // scale factors exp(m_t - m) and the softmax denominator, per joint
    if (tid < 32) {
        float m = -INFINITY;
        for (int t = 0; t < T; ++t) m = fmaxf(m, sMS[t * 64 + tid]);
        float den = 0.f;
        for (int t = 0; t < T; ++t) {
            const float f = __expf(sMS[t * 64 + tid] - m);
            den += sMS[t * 64 + 32 + tid] * f;
            sMS[t * 64 + tid] = f;
        }
        sMS[T * 64 + tid] = den;
    }
    __syncthreads();
    // ---- joint_agg[c = tid][j] (softmax over all N points of the gathered weight map, model.py:319-320)
    {
        float agg[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) agg[j] = 0.f;
#pragma unroll 4
        for (int t = 0; t < T; ++t) {
            const float4* a = reinterpret_cast<const float4*>(p.part_acc + (((size_t)b * T + t) * 128 + tid) * 32);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 v = __ldg(a + q);
                agg[4 * q] += v.x * sMS[t * 64 + 4 * q];
                agg[4 * q + 1] += v.y * sMS[t * 64 + 4 * q + 1];
                agg[4 * q + 2] += v.z * sMS[t * 64 + 4 * q + 2];
                agg[4 * q + 3] += v.w * sMS[t * 64 + 4 * q + 3];
            }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) agg[j] = j < J ? agg[j] / sMS[T * 64 + j] : 0.f;
        // B operand, MN-major [K = 128 channels][N = 32 joints]: thread k = tid writes its 32 joints
#pragma unroll
        for (int c = 0; c < 4; ++c) sH[(tid >> 3) * 32 + c * 8 + (tid & 7)] = pack8_bf16(agg + 8 * c);
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (warp_u == 0) {
        tc_fence_after();
        mbar_wait(&wbar[0], 0);
        if (elect_one()) {
            umma_gemm(tmem0 + ACCE, smem_u32(sX), 2048, 128, smem_u32(sH), 512, 128, umma_idesc_bf16(128, 32, false, true), 128, false);
            umma_commit(&mma_bar);
        }
        __syncwarp();
    }
    mbar_wait(&mma_bar, phase);
    phase ^= 1;
    tc_fence_after();
    {
        float d[32];
        tmem_ld32(tmem + ACCE, d);
        const float bj = p.wvec[tid];
        const float4 wx = *reinterpret_cast<const float4*>(p.wvec + 128 + 4 * tid);
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (j < J) {
                const float4 q = sPcl[N + j];
                const float v = fmaxf(d[j] + bj + wx.x * q.x + wx.y * q.y + wx.z * q.z, 0.f);
                sJF[j * 128 + tid] = v;
                if (sc == 0 && p.jf_out) p.jf_out[((size_t)b * J + j) * 128 + tid] = v;
            }
        }
    }
    stamp();
    // ---- ball query (pointnet2_ops: first NS hits in index order, pad with the first hit).
    // Phase 1: one thread per point tests all J centres (exact fp32 op order) -> hit bit mask per point.
    // Phase 2: one warp per centre compacts the set bits in index order with ballots (no arithmetic in the serial loop).
    {
        const float r2 = xmul(radius, radius);
        const int NW = (N + J + 31) / 32;   // hit words per centre
        // Phase 1: rounds of 8 points per thread held in registers; centres stream through (one LDS.128 per centre per round).
        // A warp's 32 lanes hold 32 CONSECUTIVE points, so one ballot per (centre, point group) IS the transposed hit word.
        for (int base = 0; base < N + J; base += 128 * 8) {
            float4 q[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int n = base + u * 128 + tid;
                q[u] = n < N + J ? sPcl[n] : make_float4(1e30f, 1e30f, 1e30f, 0.f);
            }
            for (int j = 0; j < J; ++j) {
                const float4 c = sPcl[N + j];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float dx = xsub(c.x, q[u].x), dy = xsub(c.y, q[u].y), dz = xsub(c.z, q[u].z);
                    const bool hit = xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz)) < r2;
                    const uint32_t bal = __ballot_sync(0xffffffffu, hit);
                    const int wi = (base + u * 128) / 32 + warp;
                    if (lane == 0 && wi < NW) sMask[j * NW + wi] = bal;
                }
            }
        }
        __syncthreads();
        stamp();
        // Phase 2: one warp per centre: popcount prefix over its hit words, then every lane expands the set bits of its word(s)
        // into their slots; first NS hits in index order, the rest padded with the first hit (pointnet2_ops semantics).
        for (int j = warp; j < J; j += 4) {
            const uint32_t* wj = sMask + j * NW;
            int carry = 0, first = -1;
            for (int w0 = 0; w0 < NW && carry < NS; w0 += 32) {
                const int wi = w0 + lane;
                uint32_t word = wi < NW ? wj[wi] : 0u;
                const int cntw = __popc(word);
                int incl = cntw;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += t;
                }
                int slot = carry + incl - cntw;
                const uint32_t nz = __ballot_sync(0xffffffffu, word != 0u);
                if (first < 0 && nz) {
                    const int fl = __ffs(nz) - 1;
                    const uint32_t fw = __shfl_sync(0xffffffffu, word, fl);
                    first = (w0 + fl) * 32 + __ffs(fw) - 1;
                }
                while (word && slot < NS) {
                    const int bit = __ffs(word) - 1;
                    sIdx[j * NS + slot] = (uint16_t)(wi * 32 + bit);
                    word &= word - 1;
                    ++slot;
                }
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (first < 0) first = 0;
            const int cnt = carry < NS ? carry : NS;
            for (int s2 = cnt + lane; s2 < NS; s2 += 32) sIdx[j * NS + s2] = (uint16_t)first;
        }
    }
    stamp();
    const float b1 = p.wvec[128 + 512 + sc * 256 + tid], b2 = p.wvec[128 + 512 + sc * 256 + 128 + tid];
    const float inv_r = 1.f / radius;
    __syncthreads();  // sJF, sIdx ready; the joint-embedding MMA (reader of sX, sH) has completed
    stamp();
    if (warp_u == 0) mbar_wait(&wbar[1], 0);

    const int JPT = 128 / NS;  // joints per tile (2 for nsample = 64)
    stamp();
    // point-feature row of this thread's grouped point for tile j0 (16 x 16 B, all in flight); rows >= N are joints (smem)
    auto row_index = [&](int j0_) {
        const int jj_ = j0_ + tid / NS;
        return jj_ < J ? (int)sIdx[jj_ * NS + (tid % NS)] : 0;
    };
    uint4 pre[16];
    {
        const int i0 = row_index(0);
        const uint4* src = reinterpret_cast<const uint4*>(p.e + ((size_t)b * N + (i0 < N ? i0 : 0)) * 128);
#pragma unroll
        for (int kc = 0; kc < 16; ++kc) pre[kc] = __ldg(src + kc);
    }
    for (int j0 = 0; j0 < J; j0 += JPT) {
        // ---- gather: row r = tid -> (joint j0 + r/NS, slot r%NS); the global rows were prefetched during the previous tile
        {
            const int jj = j0 + tid / NS;
            const bool ok = jj < J;
            const int ii = ok ? sIdx[jj * NS + (tid % NS)] : 0;
            const float* cf = sJF + (ok ? jj : 0) * 128;
            if (ii < N) {
#pragma unroll
                for (int kc = 0; kc < 16; ++kc) {
                    const uint4 v = pre[kc];
                    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
                    float f[8];
                    const float4 c0 = *reinterpret_cast<const float4*>(cf + kc * 8), c1 = *reinterpret_cast<const float4*>(cf + kc * 8 + 4);
                    const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 t2 = __bfloat1622float2(h[i]);
                        f[2 * i] = ok ? t2.x - cc[2 * i] : 0.f;
                        f[2 * i + 1] = ok ? t2.y - cc[2 * i + 1] : 0.f;
                    }
                    sX[kc * 128 + tid] = pack8_bf16(f);
                }
            } else {  // one of the J joints appended to the point set (model.py:168-169)
                const float* sf = sJF + (ii - N) * 128;
#pragma unroll 4
                for (int kc = 0; kc < 16; ++kc) {
                    float f[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] = ok ? sf[kc * 8 + i] - cf[kc * 8 + i] : 0.f;
                    sX[kc * 128 + tid] = pack8_bf16(f);
                }
            }
            if (j0 + JPT < J) {  // prefetch the next tile's rows; they land while this tile's MMAs / epilogues run
                const int i1 = row_index(j0 + JPT);
                const uint4* src = reinterpret_cast<const uint4*>(p.e + ((size_t)b * N + (i1 < N ? i1 : 0)) * 128);
#pragma unroll
                for (int kc = 0; kc < 16; ++kc) pre[kc] = __ldg(src + kc);
            }
            const float4 q = sPcl[ii], c = sPcl[N + (ok ? jj : 0)];
            float t8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (ok) {
                t8[0] = (q.x - c.x) * inv_r;   // group_xyz_norm = (xyz[idx] - centre) / radius   model.py:177
                t8[1] = (q.y - c.y) * inv_r;
                t8[2] = (q.z - c.z) * inv_r;
            }
            sXt[tid] = pack8_bf16(t8);
            sXt[128 + tid] = make_uint4(0, 0, 0, 0);
        }
        if (j0 == 0) stamp();
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        const uint32_t id128 = umma_idesc_bf16(128, 128, false, false);
        if (warp_u == 0) {
            tc_fence_after();
            if (elect_one()) {
                umma_gemm(tmem0 + ACC1, smem_u32(sW1), 2048, 128, smem_u32(sX), 2048, 128, id128, 128, false);
                umma_gemm(tmem0 + ACC1, smem_u32(sW1t), 2048, 128, smem_u32(sXt), 2048, 128, id128, 16, true);
                umma_commit(&mma_bar);
            }
            __syncwarp();
        }
        mbar_wait(&mma_bar, phase);
        phase ^= 1;
        tc_fence_after();
        if (j0 == 0) stamp();
        // ---- layer-1 epilogue: h[c][r] = relu(D1 + b1) -> MN-major B operand [K = channel][N = row]
        for (int c0 = 0; c0 < 128; c0 += 32) {
            float a[32];
            tmem_ld32(tmem + ACC1 + c0, a);
#pragma unroll
            for (int i = 0; i < 32; ++i) a[i] = fmaxf(a[i] + b1, 0.f);
#pragma unroll
            for (int c = 0; c < 4; ++c) sH[(tid >> 3) * 128 + (c0 / 8 + c) * 8 + (tid & 7)] = pack8_bf16(a + 8 * c);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (warp_u == 0) {
            tc_fence_after();
            if (elect_one()) {
                umma_gemm(tmem0 + ACC2, smem_u32(sW2), 2048, 128, smem_u32(sH), 2048, 128, umma_idesc_bf16(128, 128, false, true), 128, false);
                umma_commit(&mma_bar);
            }
            __syncwarp();
        }
        mbar_wait(&mma_bar, phase);
        phase ^= 1;
        tc_fence_after();
        if (j0 == 0) stamp();
        // ---- layer-2 epilogue: max over the NS grouped points of each joint  (model.py:197-198)
        for (int g = 0; g < JPT; ++g) {
            float mx = 0.f;  // relu output >= 0
            for (int c0 = g * NS; c0 < (g + 1) * NS; c0 += 32) {
                float a[32];
                tmem_ld32(tmem + ACC2 + c0, a);
#pragma unroll
                for (int i = 0; i < 32; ++i) mx = fmaxf(mx, a[i] + b2);
            }
            if (j0 + g < J) sOut[(j0 + g) * 128 + tid] = mx;
        }
        tc_fence_before();
        if (j0 == 0) stamp();
    }
    stamp();
    __syncthreads();
    for (int i = tid; i < J * 128; i += 128) p.desa_part[(((size_t)b * p.S + sc) * J) * 128 + i] = sOut[i];
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem0, 512);
}

}  // namespace kpf

extern "C" int kpf_desa_fused(const void* e, const float* part_acc, const float* part_ms, const float* pcl, const float* joint,
                              const void* wmat, const float* wvec, int B, int N, int J, int S, int nsample, float r0, float r1, float r2,
                              float r3, float* desa_part, float* jf_out, long long* dbg, cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(B >= 0 && N >= 128 && N % 128 == 0 && N + J <= 65535 && J >= 1 && J <= 32 && S >= 1 && S <= 4);
    KPF_REQUIRE(nsample == 32 || nsample == 64 || nsample == 128);
    KPF_REQUIRE(((uintptr_t)wmat % 16) == 0 && ((uintptr_t)e % 16) == 0 && ((uintptr_t)part_acc % 16) == 0 && ((uintptr_t)wvec % 16) == 0);
    if (B == 0) return 0;
    DesaParams p;
    p.e = (const __nv_bfloat16*)e; p.part_acc = part_acc; p.part_ms = part_ms; p.pcl = pcl; p.joint = joint; p.wmat = (const uint4*)wmat;
    p.wvec = wvec; p.desa_part = desa_part; p.jf_out = jf_out; p.B = B; p.N = N; p.J = J; p.T = N / 128; p.S = S; p.nsample = nsample;
    p.dbg = dbg;
    p.radius[0] = r0; p.radius[1] = r1; p.radius[2] = r2; p.radius[3] = r3;
    const size_t smem = (size_t)(2048 + 256 + 2048 + 2048 + 256 + 2048) * 16 + (size_t)((N + J + 3) / 4 * 4) * 16 + (size_t)J * 128 * 4 * 2 +
                        (size_t)(p.T * 64 + 64) * 4 + (size_t)((J * ((N + J + 31) / 32) + 3) / 4 * 4) * 4 + (size_t)J * nsample * 2 + 64;
    KPF_REQUIRE(smem <= 227 * 1024);
    cudaError_t err = kpf::set_smem(desa_fused_kernel, smem);
    if (err != cudaSuccess) return (int)err;
    desa_fused_kernel<<<B * S, 128, smem, stream>>>(p);
    KPF_CHECK_LAUNCH();
    return 0;
}

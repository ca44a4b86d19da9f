// DESA on tensor cores (model/model.py:129-204 + the joint embeddings :323-325), SURVEY.md 8f-1.  Two kernels:
//
// desa_prep_kernel   512 threads, two independent roles side by side in one launch:
//             CTA per sample:  combine the point stage's softmax partials (flash-style) -> joint_agg[J][128]
//                              jf = relu(Wj [joint_agg | joint_xyz] + b)    (model.py:323-325)   tcgen05 + fp32 xyz term
//             CTA per (sample, scale): ball query (pointnet2_ops semantics) of the J joints over the N points + the J joints
//                              themselves -> idx[b][scale][j][nsample]
// desa_tile_kernel   persistent, one CTA per SM, 512 threads.  Work item = (scale, sample, tile of 128/nsample joints); every
//             CTA takes a contiguous, scale-major range so the scale's weights stay resident:
//             X = [feat[idx] - jf[j] | (xyz[idx] - c_j)/r]  ->  h = relu(W1 X + b1)  ->  relu(W2 h + b2)  -> max over nsample
//             Both GEMMs are computed transposed -- D[c][row] = sum_k W[c][k] X[row][k], weight = M=128 A operand, activations =
//             N operand -- so thread (lane quarter, column group) owns OUTPUT CHANNEL c and DESA's max-pool over the grouped
//             points of a joint is a per-thread register reduction.  Software pipeline per iteration s: GEMM2(s) and GEMM1(s+1)
//             are issued together; while they run the threads write X(s+2) (rows prefetched one iteration earlier, indices two);
//             then epilogue 2 of s and epilogue 1 of s+1.  One __syncthreads per tile.
//   output    desa_part[b][scale][j][:], jf[b][j][:]; the 512->128 fusion conv follows in kpf_token_stack.
#include "tmem_ldst.cuh"

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
    float* ctx;               // scratch [B][J*128 + 128]: jf (fp32) | joint xyz padded to [32][4]
    uint16_t* idx;            // scratch [B,S,J,nsample] ball-query indices (>= N: one of the joints)
    int B, N, J, T, S, nsample;
    float radius[4];
    long long* dbg;
};

constexpr int DS_NT = 512;
constexpr int DS_MAT_PER_SCALE = 2048 + 256 + 2048;
constexpr int DS_XBUF = 2048 + 256;   // uint4 per activation buffer (main + K tail)

// ================================================================================================ prep
// Two roles in one launch: CTAs [0, B) embed the joints of one sample (softmax-partial combine + tcgen05 GEMM); CTAs
// [B, B + B*S) run the ball query of one (sample, scale).  The roles are independent and run side by side on different SMs.
__global__ void __launch_bounds__(DS_NT, 1) desa_prep_kernel(const DesaParams p) {
    extern __shared__ __align__(128) unsigned char ds_smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int J = p.J, N = p.N, T = p.T, NS = p.nsample, S = p.S;
    int n_stamp = 0;
    auto stamp = [&]() {
        if (p.dbg && (blockIdx.x == 0 || (int)blockIdx.x == p.B) && tid == 0 && n_stamp < 8) p.dbg[(blockIdx.x == 0 ? 0 : 8) + n_stamp] = clock64();
        ++n_stamp;
    };
    stamp();
    pdl_launch_dependents();

    if ((int)blockIdx.x >= p.B) {
        // ================= ball query (pointnet2_ops: first NS hits in index order, padded with the first hit) =================
        const int pi = blockIdx.x - p.B, b = pi / S, sc = pi - b * S;
        pdl_wait();
        float4* sPcl = reinterpret_cast<float4*>(ds_smem);                          // [N + J] xyz
        uint32_t* sMask = reinterpret_cast<uint32_t*>(sPcl + (N + J + 3) / 4 * 4);   // [J][NW] hit words (bit = point)
        for (int i = tid; i < N + J; i += DS_NT) {
            const float* s = i < N ? p.pcl + ((size_t)b * N + i) * 3 : p.joint + ((size_t)b * J + (i - N)) * 3;
            sPcl[i] = make_float4(s[0], s[1], s[2], 0.f);
        }
        __syncthreads();
        stamp();
        // Phase 1: one thread per point tests all J centres (exact fp32 op order).  A warp's 32 lanes hold 32 CONSECUTIVE
        // points, so one ballot per (centre, point group) IS the hit word.
        const int NW = (N + J + 31) / 32;
        const float r2 = xmul(p.radius[sc], p.radius[sc]);
        for (int base = 0; base + 32 * warp < N + J; base += DS_NT) {   // warp-uniform: warps without points skip the round
            const int n = base + tid;
            const float4 q = n < N + J ? sPcl[n] : make_float4(1e30f, 1e30f, 1e30f, 0.f);
            const int wi = (base >> 5) + warp;
#pragma unroll 3
            for (int j = 0; j < J; ++j) {
                const float4 c = sPcl[N + j];
                const float dx = xsub(c.x, q.x), dy = xsub(c.y, q.y), dz = xsub(c.z, q.z);
                const float d2 = xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz));
                const uint32_t bal = __ballot_sync(0xffffffffu, d2 < r2);
                if (lane == 0) sMask[j * NW + wi] = bal;
            }
        }
        __syncthreads();
        stamp();
        // Phase 2: one warp per centre: popcount prefix over its hit words, then every lane expands the set bits of its word(s)
        // into their slots
        for (int j = warp; j < J; j += DS_NT / 32) {
            const uint32_t* wj = sMask + j * NW;
            uint16_t* out = p.idx + (((size_t)b * S + sc) * J + j) * NS;
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
                    out[slot] = (uint16_t)(wi * 32 + bit);
                    word &= word - 1;
                    ++slot;
                }
                carry += __shfl_sync(0xffffffffu, incl, 31);
            }
            if (first < 0) first = 0;
            const int cnt = carry < NS ? carry : NS;
            for (int s2 = cnt + lane; s2 < NS; s2 += 32) out[s2] = (uint16_t)first;
        }
        stamp();
        return;
    }

    // ================= joint embedding: jf = relu(Wj [joint_agg | joint_xyz] + b)  (model.py:319-325) =================
    uint4* sWj = reinterpret_cast<uint4*>(ds_smem);             // [16][128] K-major A operand
    uint4* sAgg = sWj + 2048;                                    // MN-major B operand [16][4][8]: joint_agg[channel][joint]
    float* sMS = reinterpret_cast<float*>(sAgg + 512);           // [T][2][32] partial max/sum -> [T][32] factors + den[32]
    float4* sJ = reinterpret_cast<float4*>(sMS + T * 64 + 64);   // [32] joint xyz
    __shared__ __align__(8) uint64_t wbar, mma_bar;
    __shared__ uint32_t tmem_slot;
    const int warp_u = warp_index_uniform();
    const int b = blockIdx.x;
    if (warp == 0) tmem_alloc(&tmem_slot, 32);
    if (tid == 0) {
        mbar_init(&wbar, 1);
        mbar_init(&mma_bar, 1);
        fence_mbar_init();
        mbar_expect_tx(&wbar, 2048 * 16);
        tma_bulk_g2s(sWj, p.wmat, 2048 * 16, &wbar);
    }
    pdl_wait();   // weights only so far
    if (tid < 32) {
        const float* s = p.joint + ((size_t)b * J + (tid < J ? tid : 0)) * 3;
        sJ[tid] = tid < J ? make_float4(s[0], s[1], s[2], 0.f) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    for (int i = tid; i < T * 64; i += DS_NT) sMS[i] = p.part_ms[(size_t)b * T * 64 + i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = tmem_slot;
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
    stamp();
    // ---- joint_agg[c][j] (softmax over all N points of the gathered weight map): 8 lanes read one channel's 128-byte row of
    //      a partial per load (4 L1 wavefronts per warp load); thread = (channel, 4 joints), two channel halves
    {
        const int jq = lane & 7;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int ch = 64 * half + 4 * warp + (lane >> 3);
            float agg[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 8
            for (int t = 0; t < T; ++t) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(p.part_acc + (((size_t)b * T + t) * 128 + ch) * 32) + jq);
                const float* f = sMS + t * 64 + 4 * jq;
                agg[0] += v.x * f[0];
                agg[1] += v.y * f[1];
                agg[2] += v.z * f[2];
                agg[3] += v.w * f[3];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) agg[j] = (4 * jq + j) < J ? agg[j] / sMS[T * 64 + 4 * jq + j] : 0.f;
            const __nv_bfloat162 lo = __floats2bfloat162_rn(agg[0], agg[1]), hi = __floats2bfloat162_rn(agg[2], agg[3]);
            uint2 o;
            o.x = *reinterpret_cast<const uint32_t*>(&lo);
            o.y = *reinterpret_cast<const uint32_t*>(&hi);
            // (k = channel, n = joint), n contiguous: chunk jq / 2 of the row, half jq & 1
            reinterpret_cast<uint2*>(sAgg + (ch >> 3) * 32 + (jq >> 1) * 8 + (ch & 7))[jq & 1] = o;
        }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if (warp_u == 0) {
        tc_fence_after();
        mbar_wait(&wbar, 0);
        if (elect_one()) {
            umma_gemm(tmem0, smem_u32(sWj), 2048, 128, smem_u32(sAgg), 512, 128, umma_idesc_bf16(128, 32, false, true), 128, false);
            umma_commit(&mma_bar);
        }
        __syncwarp();
    }
    stamp();
    mbar_wait(&mma_bar, 0);
    tc_fence_after();
    {   // jf[j][ch]: thread (lane quarter q, column group cg) -> channel 32q + lane, joints [8cg, 8cg + 8)
        const int q = warp & 3, cg = warp >> 2, ch = 32 * q + lane;
        float d[8];
        tmem_ld<8>(tmem0 + ((uint32_t)(32 * q) << 16) + 8 * cg, d);
        const float bj = p.wvec[ch];
        const float4 wx = *reinterpret_cast<const float4*>(p.wvec + 128 + 4 * ch);
        float* ctx = p.ctx + (size_t)b * (J * 128 + 128);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int j = 8 * cg + i;
            if (j < J) {
                const float4 c = sJ[j];
                const float v = fmaxf(d[i] + bj + wx.x * c.x + wx.y * c.y + wx.z * c.z, 0.f);
                ctx[j * 128 + ch] = v;
                if (p.jf_out) p.jf_out[((size_t)b * J + j) * 128 + ch] = v;
            }
        }
        if (tid < 32) reinterpret_cast<float4*>(ctx + J * 128)[tid] = sJ[tid];
    }
    stamp();
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem0, 32);
}

// ================================================================================================ tiles
constexpr int DS_TILE_NT = DS_NT + 32;   // 16 worker warps + one warp that only issues MMAs / TMA copies

struct DesaItem {   // (scale, sample, first joint) of a work item, advanced incrementally (no divisions in the loop)
    int sc, b, j0;
};

__global__ void __launch_bounds__(DS_TILE_NT, 1) desa_tile_kernel(const DesaParams p) {
    extern __shared__ __align__(128) unsigned char ds_smem[];
    uint4* sW1 = reinterpret_cast<uint4*>(ds_smem);   // [16][128] + tail [2][128]
    uint4* sW2 = sW1 + 2048 + 256;                     // [16][128]
    uint4* sX = sW2 + 2048;                            // [2] x ([16][128] K-major activations + tail [2][128])
    uint4* sH = sX + 2 * DS_XBUF;                      // MN-major [16][16][8]
    float* sCtx = reinterpret_cast<float*>(sH + 2048); // [2][J*128 + 128] jf | joint xyz of the sample(s) in flight
    const int ctx_n = p.J * 128 + 128;
    float* sPart = sCtx + 2 * ctx_n;                   // [2][4][128] per-column-group maxima
    __shared__ __align__(8) uint64_t wbar, g1_bar, g2_bar, ctx_bar[2];
    __shared__ uint32_t tmem_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int warp_u = warp_index_uniform();
    const bool issuer = warp_u == DS_NT / 32;                     // warp 16: MMA / TMA issue only
    const int q = warp & 3, cg = (warp >> 2) & 3, ch = 32 * q + lane;   // epilogues: channel ch, tile rows [32cg, 32cg + 32)
    const int r = tid & 127, kq = (tid >> 7) & 3;                 // gather: tile row r, channel chunks [4kq, 4kq + 4)
    const int J = p.J, N = p.N, NS = p.nsample, S = p.S, B = p.B;
    const int JPT = 128 / NS, TPS = (J + JPT - 1) / JPT;          // joints per tile, tiles per (sample, scale)
    const int jrow = r >> (31 - __clz(NS));                       // r / NS (NS is a power of two): joint of this gather row within the tile
    const int total = S * B * TPS;
    const int it0 = (int)((long long)total * blockIdx.x / gridDim.x), it1 = (int)((long long)total * (blockIdx.x + 1) / gridDim.x);
    const uint32_t ACC1 = 0, ACC2 = 128;
    int n_stamp = 0;
    auto stamp = [&]() {
        if (p.dbg && blockIdx.x == 0 && tid == 0 && n_stamp < 48) p.dbg[16 + n_stamp] = clock64();
        ++n_stamp;
    };
    stamp();
    pdl_launch_dependents();
    if (warp == 0) tmem_alloc(&tmem_slot, 256);
    if (tid == 0) {
        mbar_init(&wbar, 1);
        mbar_init(&g1_bar, 1);
        mbar_init(&g2_bar, 1);
        mbar_init(&ctx_bar[0], 1);
        mbar_init(&ctx_bar[1], 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = tmem_slot, tmem = tmem0 + ((uint32_t)(32 * q) << 16);
    uint32_t g1_phase = 0, g2_phase = 0, w_phase = 0;
    pdl_wait();   // indices, contexts and point features come from the previous kernels

    auto decode = [&](int item) {
        DesaItem it;
        it.sc = item / (B * TPS);
        const int rem = item - it.sc * (B * TPS);
        it.b = rem / TPS;
        it.j0 = (rem - it.b * TPS) * JPT;
        return it;
    };
    auto advance = [&](DesaItem& it) {
        it.j0 += JPT;
        if (it.j0 >= J) {
            it.j0 = 0;
            if (++it.b == B) {
                it.b = 0;
                ++it.sc;
            }
        }
    };
    // ---- gather-side register pipeline (worker warps); every stage walks the items with its own cursor
    DesaItem c_idx, c_rows, c_store, c_max;
    int ii_n = 0;            // ball-query index of this thread's row for the item whose rows are fetched next
    int ii_r = 0;            // ... for the item whose rows are in `pre`
    uint4 pre[4];            // 4 x 16 B of the point-feature row (channels [32kq, 32kq + 32))
    float3 pxyz = make_float3(0.f, 0.f, 0.f);
    int ctx_loads = 0;       // sample contexts requested so far (slot = count & 1)
    int ctx_b_load = -1;     // sample of the most recent request
    int ctx_k = -1;          // index of the context the store stage uses
    int ctx_b_store = -1;
    float inv_r = 1.f;       // 1 / radius of the run's scale

    auto fetch_idx = [&]() {
        const int jj = c_idx.j0 + jrow;
        ii_n = jj < J ? (int)__ldg(p.idx + (((size_t)c_idx.b * S + c_idx.sc) * J + c_idx.j0) * NS + r) : 0;
        advance(c_idx);
    };
    auto fetch_rows = [&]() {   // uses ii_n; all warps track the context requests, the issuer warp makes them
        const int b = c_rows.b;
        if (b != ctx_b_load) {   // first item of a sample on the gather side (block-uniform branch)
            if (issuer) {
                if (elect_one()) {
                    mbar_expect_tx(&ctx_bar[ctx_loads & 1], (uint32_t)ctx_n * 4);
                    tma_bulk_g2s(sCtx + (ctx_loads & 1) * ctx_n, p.ctx + (size_t)b * ctx_n, (uint32_t)ctx_n * 4, &ctx_bar[ctx_loads & 1]);
                }
                __syncwarp();
            }
            ctx_b_load = b;
            ++ctx_loads;
        }
        advance(c_rows);
        if (issuer) return;
        ii_r = ii_n;
        const int i0 = ii_r < N ? ii_r : 0;
        const uint4* src = reinterpret_cast<const uint4*>(p.e + ((size_t)b * N + i0) * 128) + 4 * kq;
#pragma unroll
        for (int k = 0; k < 4; ++k) pre[k] = __ldg(src + k);
        if (kq == 0) {
            const float* s = p.pcl + ((size_t)b * N + i0) * 3;
            pxyz = make_float3(__ldg(s), __ldg(s + 1), __ldg(s + 2));
        }
    };
    auto store_x = [&](int item) {      // uses pre / ii_r / pxyz
        const DesaItem it = c_store;
        advance(c_store);
        if (it.b != ctx_b_store) {   // first tile of a sample on the store side: its context must have landed
            ctx_b_store = it.b;
            ++ctx_k;
            mbar_wait(&ctx_bar[ctx_k & 1], (ctx_k >> 1) & 1);
        }
        const float* cx = sCtx + (ctx_k & 1) * ctx_n;
        uint4* X = sX + (item & 1) * DS_XBUF;
        const int jj = it.j0 + jrow;
        const bool ok = jj < J;
        const float* cf = cx + (ok ? jj : 0) * 128 + 32 * kq;
        if (!ok) {
#pragma unroll
            for (int k = 0; k < 4; ++k) X[(4 * kq + k) * 128 + r] = make_uint4(0, 0, 0, 0);
        } else if (ii_r < N) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&pre[k]);
                const float4 c0 = *reinterpret_cast<const float4*>(cf + k * 8), c1 = *reinterpret_cast<const float4*>(cf + k * 8 + 4);
                const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                float f[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 t2 = __bfloat1622float2(h[i]);
                    f[2 * i] = t2.x - cc[2 * i];
                    f[2 * i + 1] = t2.y - cc[2 * i + 1];
                }
                X[(4 * kq + k) * 128 + r] = pack8_bf16(f);
            }
        } else {  // one of the J joints appended to the point set (model.py:168-169)
            const float* sf = cx + (ii_r - N) * 128 + 32 * kq;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = sf[k * 8 + i] - cf[k * 8 + i];
                X[(4 * kq + k) * 128 + r] = pack8_bf16(f);
            }
        }
        if (kq == 0) {
            const float4* cxyz = reinterpret_cast<const float4*>(cx + J * 128);
            const float4 c = cxyz[ok ? jj : 0];
            float3 pq = pxyz;
            if (ii_r >= N) {
                const float4 t4 = cxyz[ii_r - N];
                pq = make_float3(t4.x, t4.y, t4.z);
            }
            float t8[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (ok) {
                t8[0] = (pq.x - c.x) * inv_r;   // group_xyz_norm = (xyz[idx] - centre) / radius   model.py:177
                t8[1] = (pq.y - c.y) * inv_r;
                t8[2] = (pq.z - c.z) * inv_r;
            }
            X[2048 + r] = pack8_bf16(t8);
            X[2048 + 128 + r] = make_uint4(0, 0, 0, 0);
        }
    };
    // maxima of a finished tile: combine the column groups of each joint, store
    auto store_max = [&](int item) {
        const DesaItem it = c_max;
        advance(c_max);
        if (tid < JPT * 128) {
            const int g = tid >> 7, c = tid & 127, nc = NS / 32;
            const float* pp = sPart + (item & 1) * 512 + (g * nc) * 128 + c;
            float m = pp[0];
            for (int k = 1; k < nc; ++k) m = fmaxf(m, pp[k * 128]);
            if (it.j0 + g < J) p.desa_part[(((size_t)it.b * S + it.sc) * J + it.j0 + g) * 128 + c] = m;
        }
    };

    // ---- runs of items that share a scale (= weights)
    for (int i0 = it0; i0 < it1;) {
        const DesaItem first = decode(i0);
        const int sc0 = first.sc;
        const int run_end = (sc0 + 1) * B * TPS;
        const int i1 = it1 < run_end ? it1 : run_end;
        c_idx = c_rows = c_store = c_max = first;
        // every MMA of the previous run has completed (its epilogues ran), so the weight buffers are free
        if (issuer) {
            if (elect_one()) {
                const uint4* ws = p.wmat + 2048 + (size_t)sc0 * DS_MAT_PER_SCALE;
                mbar_expect_tx(&wbar, DS_MAT_PER_SCALE * 16);
                tma_bulk_g2s(sW1, ws, DS_MAT_PER_SCALE * 16, &wbar);   // W1 | W1 tail | W2 are contiguous on both sides
            }
            __syncwarp();
        }
        const float b1 = p.wvec[128 + 512 + sc0 * 256 + ch], b2 = p.wvec[128 + 512 + sc0 * 256 + 128 + ch];
        inv_r = 1.f / p.radius[sc0];
        bool w_ready = false;
        // fill: indices of i0, rows of i0, indices of i0 + 1
        if (!issuer) fetch_idx();
        fetch_rows();
        if (!issuer && i0 + 1 < i1) fetch_idx();
        for (int s = i0 - 2; s < i1; ++s) {
            if (s >= i0 - 1) {
                fence_proxy_async();
                tc_fence_before();
                __syncthreads();
                if (issuer) {
                    tc_fence_after();
                    if (!w_ready) mbar_wait(&wbar, w_phase);
                    if (elect_one()) {
                        if (s >= i0) {   // layer 2 of tile s: D2[c][row] = W2 h
                            umma_gemm(tmem0 + ACC2, smem_u32(sW2), 2048, 128, smem_u32(sH), 2048, 128, umma_idesc_bf16(128, 128, false, true),
                                      128, false);
                            umma_commit(&g2_bar);
                        }
                        if (s + 1 < i1) {   // layer 1 of tile s + 1: D1[c][row] = W1 [feat - jf | xyz]
                            const uint4* X = sX + ((s + 1) & 1) * DS_XBUF;
                            const uint32_t id128 = umma_idesc_bf16(128, 128, false, false);
                            umma_gemm(tmem0 + ACC1, smem_u32(sW1), 2048, 128, smem_u32(X), 2048, 128, id128, 128, false);
                            umma_gemm(tmem0 + ACC1, smem_u32(sW1 + 2048), 2048, 128, smem_u32(X + 2048), 2048, 128, id128, 16, true);
                            umma_commit(&g1_bar);
                        }
                    }
                    __syncwarp();
                }
                w_ready = true;
                if (s - 1 >= i0) store_max(s - 1);   // written before the barrier above
            }
            // gather side, two / three / four tiles ahead
            if (!issuer && s + 2 < i1) store_x(s + 2);
            if (s + 3 < i1) fetch_rows();
            if (!issuer) {
                if (s + 4 < i1) fetch_idx();
                if (s >= i0) {   // layer-2 epilogue of tile s: max over this thread's 32 grouped points  (model.py:197-198)
                    mbar_wait(&g2_bar, g2_phase);
                    g2_phase ^= 1;
                    tc_fence_after();
                        float a[32];
                    tmem_ld<32>(tmem + ACC2 + 32 * cg, a);
                    float mx = a[0];
#pragma unroll
                    for (int i = 1; i < 32; ++i) mx = fmaxf(mx, a[i]);
                    sPart[(s & 1) * 512 + cg * 128 + ch] = fmaxf(mx + b2, 0.f);   // max_i relu(a_i + b2)
                    tc_fence_before();
                    }
                if (s >= i0 - 1 && s + 1 < i1) {   // layer-1 epilogue of tile s + 1: h[c][row] = relu(D1 + b1) -> MN-major B operand
                    mbar_wait(&g1_bar, g1_phase);
                    g1_phase ^= 1;
                    tc_fence_after();
                        float a[32];
                    tmem_ld<32>(tmem + ACC1 + 32 * cg, a);
#pragma unroll
                    for (int i = 0; i < 32; ++i) a[i] = fmaxf(a[i] + b1, 0.f);
#pragma unroll
                    for (int c = 0; c < 4; ++c) sH[(ch >> 3) * 128 + (4 * cg + c) * 8 + (ch & 7)] = pack8_bf16(a + 8 * c);
                }
            }
            if (s <= i0 + 3) stamp();
        }
        __syncthreads();
        store_max(i1 - 1);
        w_phase ^= 1;
        i0 = i1;
        stamp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem0, 256);
}

}  // namespace kpf

extern "C" int kpf_desa_fused(const void* e, const float* part_acc, const float* part_ms, const float* pcl, const float* joint,
                              const void* wmat, const float* wvec, int B, int N, int J, int S, int nsample, float r0, float r1, float r2,
                              float r3, float* desa_part, float* jf_out, void* scratch, int num_sms, long long* dbg, cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(B >= 0 && N >= 128 && N % 128 == 0 && N + J <= 65535 && J >= 1 && J <= 32 && S >= 1 && S <= 4);
    KPF_REQUIRE(nsample == 32 || nsample == 64 || nsample == 128);
    KPF_REQUIRE(((uintptr_t)wmat % 16) == 0 && ((uintptr_t)e % 16) == 0 && ((uintptr_t)part_acc % 16) == 0 && ((uintptr_t)wvec % 16) == 0);
    KPF_REQUIRE(scratch != nullptr && ((uintptr_t)scratch % 16) == 0 && num_sms >= 1);
    if (B == 0) return 0;
    DesaParams p;
    p.e = (const __nv_bfloat16*)e; p.part_acc = part_acc; p.part_ms = part_ms; p.pcl = pcl; p.joint = joint; p.wmat = (const uint4*)wmat;
    p.wvec = wvec; p.desa_part = desa_part; p.jf_out = jf_out; p.B = B; p.N = N; p.J = J; p.T = N / 128; p.S = S; p.nsample = nsample;
    p.dbg = dbg;
    p.radius[0] = r0; p.radius[1] = r1; p.radius[2] = r2; p.radius[3] = r3;
    const size_t ctx_n = (size_t)J * 128 + 128;
    p.ctx = (float*)scratch;
    p.idx = (uint16_t*)((char*)scratch + (size_t)B * ctx_n * 4);
    const int NW = (N + J + 31) / 32;
    const size_t smem_jf = (size_t)(2048 + 512) * 16 + (size_t)(p.T * 64 + 64) * 4 + 32 * 16 + 64;
    const size_t smem_bq = (size_t)((N + J + 3) / 4 * 4) * 16 + (size_t)J * NW * 4 + 64;
    const size_t smem_a = smem_jf > smem_bq ? smem_jf : smem_bq;
    const size_t smem_b = (size_t)(DS_MAT_PER_SCALE + 2 * DS_XBUF + 2048) * 16 + 2 * ctx_n * 4 + 2 * 512 * 4 + 64;
    KPF_REQUIRE(smem_a <= 227 * 1024 && smem_b <= 227 * 1024);
    cudaError_t err = kpf::set_smem(desa_prep_kernel, smem_a);
    if (err != cudaSuccess) return (int)err;
    err = kpf::set_smem(desa_tile_kernel, smem_b);
    if (err != cudaSuccess) return (int)err;
    err = kpf::launch_pdl(desa_prep_kernel, dim3(B + B * S), dim3(DS_NT), smem_a, stream, p);
    if (err != cudaSuccess) return (int)err;
    KPF_CHECK_LAUNCH();
    const int JPT = 128 / nsample, total = S * B * ((J + JPT - 1) / JPT);
    err = kpf::launch_pdl(desa_tile_kernel, dim3(total < num_sms ? total : num_sms), dim3(DS_TILE_NT), smem_b, stream, p);
    if (err != cudaSuccess) return (int)err;
    KPF_CHECK_LAUNCH();
    return 0;
}

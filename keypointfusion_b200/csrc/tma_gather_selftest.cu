// Self-test + cycle probe of the TMA row gather (csrc/tma_gather.cuh) feeding a SWIZZLE_128B K-major B operand (tests/test_umma_gpu.py):
//   D[128 x N] = A[128 x 128] X[idx[n]][0..128)^T    A fp32 carried as fp16 planes (shared memory, no swizzle), X = rows of a
//   [R][256] 16-bit table ([hi 128 | lo 128] fp16 planes: the layout of kpf_point_embed's e), gathered four rows per instruction.
//   cycles[0] = gather of the N rows (both planes), cold; [1] = one GEMM (24 MMAs); [2] = the gather again (L2-hot); [3] = eight GEMMs.
#include "tma_gather.cuh"
#include "umma_split.cuh"

namespace kpf {

__global__ void __launch_bounds__(128)
tma_gather_selftest_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ A, const int* __restrict__ idx, float* __restrict__ D,
                           int N, int a_tmem, long long* cycles) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar, gbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int warp_u = warp_index_uniform();
    // X planes: [plane hi, lo][k half 0, 1][N rows][128 B]  (swizzled 8-row atoms); then A planes [16 k-chunks][128] hi, lo
    unsigned char* sXb = sm;
    uint4* sAh = reinterpret_cast<uint4*>(sm + 4 * N * 128);
    uint4* sAl = sAh + 16 * 128;
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_init(&gbar, 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot, lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    for (int kc = 0; kc < 16; ++kc) {   // A: thread = row
        float v[8];
        for (int i = 0; i < 8; ++i) v[i] = A[(size_t)tid * 128 + kc * 8 + i];
        uint4 hi, lo;
        split8(FMT_F16, v, hi, lo);
        sAh[kc * 128 + tid] = hi;
        sAl[kc * 128 + tid] = lo;
        if (a_tmem) {   // the same planes as tensor-memory A operands (columns 256.., 320..)
            tmem_st_nw<4>(lane_base + 256 + kc * 4, reinterpret_cast<const float*>(&hi));
            tmem_st_nw<4>(lane_base + 320 + kc * 4, reinterpret_cast<const float*>(&lo));
        }
    }
    if (a_tmem) tmem_wait_st();
    tc_fence_before();
    fence_proxy_async();
    __syncthreads();
    long long t0 = 0, t1 = 0, t_hot = 0;
    for (int rep = 0; rep < 3; ++rep) {   // rep 0: cold (the table comes from DRAM); rep 2: L2-hot steady state
        __syncthreads();
        t0 = clock64();
        if (warp_u == 0) {   // the gather: lane L fetches row groups L, L + 32, ... (4 rows each), four 64-element segments per row
            const int lane = tid & 31;
            if (lane == 0) mbar_expect_tx(&gbar, (uint32_t)N * 512);
            __syncwarp();
            for (int g = lane; g < N / 4; g += 32) {
                const int4 r = *reinterpret_cast<const int4*>(idx + 4 * g);
#pragma unroll
                for (int seg = 0; seg < 4; ++seg)   // seg = 2 plane + k half
                    tma_gather4(sXb + (size_t)seg * N * 128 + (size_t)g * 512, &tmap, 64 * seg, r.x, r.y, r.z, r.w, &gbar);
            }
        }
        mbar_wait(&gbar, rep & 1);
        t1 = clock64();
        if (rep == 0) t_hot = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0 && cycles) {
        cycles[0] = t_hot;      // cold gather
        cycles[2] = t1 - t0;    // L2-hot gather
    }
    for (int reps = 1; reps <= 8; reps *= 8) {   // one GEMM, then eight back to back (steady-state cycles per MMA); same result
        __syncthreads();
        const long long t2 = clock64();
        if (warp_u == 0) {
            if (elect_one()) {
                const uint32_t idesc = umma_idesc_f16(128, N, false, false, FMT_F16, FMT_F16);
                const uint32_t xb = smem_u32(sXb), bh = desc_hi_sw128(1024);
                const uint32_t ah = desc_hi(128);
                for (int r = 0; r < reps; ++r) {
                    uint32_t acc = 0;
                    // three plane products: A_lo X_hi, A_hi X_lo, A_hi X_hi
                    for (int term = 0; term < 3; ++term) {
                        const uint32_t a_addr = smem_u32(term == 0 ? sAl : sAh);
                        const uint32_t a_t = tmem + (term == 0 ? 320u : 256u);
                        const uint32_t x_addr = xb + (term == 1 ? 2u : 0u) * (uint32_t)N * 128;
#pragma unroll
                        for (int ks = 0; ks < 8; ++ks) {
                            const uint32_t bl = desc_lo_sw128(x_addr + (ks >> 2) * (uint32_t)N * 128 + (ks & 3) * 32);
                            if (a_tmem) umma_issue_ts(tmem, a_t + 8 * ks, bl, bh, idesc, acc);
                            else umma_issue_ss(tmem, desc_lo(a_addr + ks * 2 * 2048, 2048), ah, bl, bh, idesc, acc);
                            acc = 1u;
                        }
                    }
                }
                umma_commit(&bar);
            }
            __syncwarp();
        }
        mbar_wait(&bar, reps == 1 ? 0 : 1);
        tc_fence_after();
        const long long t3 = clock64();
        if (tid == 0 && cycles) cycles[reps == 1 ? 1 : 3] = t3 - t2;
    }
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld<16>(lane_base + c0, v);
        for (int i = 0; i < 16; ++i) D[(size_t)tid * N + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}


// Cycle probe: every CTA gathers the same-sized set of 512-byte rows (n_rows, indices idx[cta][n_rows]) into shared memory four times;
// out[cta] = cycles of the last (L2-hot) repetition.  mode 0: TMA gather4 (warp 0 issues); 1: 1-D bulk copies of 512 B (warp 0);
// 2: cp.async 16 B per lane, 8 lanes per 128-byte line (all 512 threads, the DESA mapping); 3: LDG.128 + STS.128 (all threads);
// 4: cp.async 16 B per lane, a warp instruction = one whole 512-byte row.
__global__ void __launch_bounds__(512)
gather_probe_kernel(const __grid_constant__ CUtensorMap tmap, const uint4* __restrict__ table, const int* __restrict__ idx, int n_rows, int mode,
                    long long* out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ __align__(8) uint64_t gbar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int* my = idx + (size_t)blockIdx.x * n_rows;
    uint4* dst = reinterpret_cast<uint4*>(sm);
    if (tid == 0) {
        mbar_init(&gbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    long long t0 = 0, t1 = 0;
    uint32_t phase = 0;
    for (int rep = 0; rep < 4; ++rep) {
        __syncthreads();
        t0 = clock64();
        if (mode == 0) {
            if (warp == 0) {
                if (lane == 0) mbar_expect_tx(&gbar, (uint32_t)n_rows * 512);
                __syncwarp();
                for (int g = lane; g < n_rows / 4; g += 32) {
                    const int4 r = *reinterpret_cast<const int4*>(my + 4 * g);
#pragma unroll
                    for (int seg = 0; seg < 4; ++seg)
                        tma_gather4(sm + (size_t)seg * n_rows * 128 + (size_t)g * 512, &tmap, 64 * seg, r.x, r.y, r.z, r.w, &gbar);
                }
            }
            mbar_wait(&gbar, phase);
            phase ^= 1;
        } else if (mode == 1) {
            if (warp == 0) {
                if (lane == 0) mbar_expect_tx(&gbar, (uint32_t)n_rows * 512);
                __syncwarp();
                for (int r = lane; r < n_rows; r += 32) tma_bulk_g2s(sm + (size_t)r * 512, table + (size_t)my[r] * 32, 512, &gbar);
            }
            mbar_wait(&gbar, phase);
            phase ^= 1;
        } else if (mode == 2) {
            // thread -> rows 4 * warp + (lane & 3) + 64 h, 16-byte chunks (lane >> 2) + 8 k
            for (int h = 0; h * 64 < n_rows; ++h) {
                const int row = 4 * warp + (lane & 3) + 64 * h;
                if (row < n_rows) {
                    const uint4* src = table + (size_t)my[row] * 32 + (lane >> 2);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + row * 32 + (lane >> 2) + 8 * k)), "l"(src + 8 * k) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        } else if (mode == 3) {
            for (int h = 0; h * 64 < n_rows; ++h) {
                const int row = 4 * warp + (lane & 3) + 64 * h;
                if (row < n_rows) {
                    const uint4* src = table + (size_t)my[row] * 32 + (lane >> 2);
                    uint4 v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = __ldcg(src + 8 * k);
#pragma unroll
                    for (int k = 0; k < 4; ++k) dst[row * 32 + (lane >> 2) + 8 * k] = v[k];
                }
            }
        } else {
            for (int row = warp; row < n_rows; row += 16)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + row * 32 + lane)), "l"(table + (size_t)my[row] * 32 + lane) : "memory");
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();
        t1 = clock64();
    }
    if (tid == 0) out[blockIdx.x] = t1 - t0;
}

}  // namespace kpf

extern "C" int kpf_tma_gather_selftest(const void* table, long long rows, const float* A, const int* idx, float* D, int N, int a_tmem, long long* cycles,
                                       cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(N >= 16 && N <= 256 && N % 16 == 0 && rows >= 1 && ((uintptr_t)table % 16) == 0 && ((uintptr_t)idx % 16) == 0);
    CUtensorMap map;
    int rc = make_row_gather_map(&map, table, (unsigned long long)rows, 256, 512);
    if (rc) return rc;
    const size_t smem = (size_t)4 * N * 128 + 2 * 16 * 128 * 16 + 1024;
    cudaError_t e = kpf::set_smem(tma_gather_selftest_kernel, smem);
    if (e != cudaSuccess) return (int)e;
    tma_gather_selftest_kernel<<<1, 128, smem, stream>>>(map, A, idx, D, N, a_tmem, cycles);
    KPF_CHECK_LAUNCH();
    return 0;
}

extern "C" int kpf_gather_probe(const void* table, long long rows, const int* idx, int n_rows, int ctas, int mode, long long* out, cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(n_rows >= 4 && n_rows <= 256 && n_rows % 4 == 0 && ctas >= 1 && mode >= 0 && mode <= 4);
    CUtensorMap map;
    int rc = make_row_gather_map(&map, table, (unsigned long long)rows, 256, 512);
    if (rc) return rc;
    const size_t smem = (size_t)n_rows * 512 + 1024;
    cudaError_t e = kpf::set_smem(gather_probe_kernel, smem);
    if (e != cudaSuccess) return (int)e;
    gather_probe_kernel<<<ctas, 512, smem, stream>>>(map, (const uint4*)table, idx, n_rows, mode, out);
    KPF_CHECK_LAUNCH();
    return 0;
}

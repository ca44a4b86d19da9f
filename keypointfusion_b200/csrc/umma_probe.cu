// Micro-benchmarks of the tcgen05 building blocks (cycles, one CTA): used to direct optimisation, not on the product path.
#include "umma.cuh"

namespace kpf {

__global__ void __launch_bounds__(128) umma_probe_kernel(long long* out, int N, int K, int reps) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    uint4* sA = reinterpret_cast<uint4*>(sm);
    uint4* sB = sA + 128 * (K / 8);
    for (int i = tid; i < (128 + N) * (K / 8); i += 128) sA[i] = make_uint4(0x3c003c00u, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u);
    if (warp == 0) tmem_alloc(&tmem_slot, 512);
    if (tid == 0) {
        mbar_init(&bar, 1);
        fence_mbar_init();
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    uint32_t phase = 0;
    long long t0, t1;
    // (0) one GEMM (K/16 MMAs) issue -> commit -> wait, repeated
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (tid == 0) {
            umma_gemm(tmem, smem_u32(sA), 128 * 16, 128, smem_u32(sB), N * 16, 128, umma_idesc_bf16(128, N, false, false), K, false);
            umma_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        tc_fence_after();
    }
    t1 = clock64();
    if (tid == 0) out[0] = (t1 - t0) / reps;
    // (1) 8 GEMMs back to back before one commit
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (tid == 0) {
            for (int q = 0; q < 8; ++q)
                umma_gemm(tmem, smem_u32(sA), 128 * 16, 128, smem_u32(sB), N * 16, 128, umma_idesc_bf16(128, N, false, false), K, false);
            umma_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        tc_fence_after();
    }
    t1 = clock64();
    if (tid == 0) out[1] = (t1 - t0) / reps;
    // (2) sync round trip without MMA work: fence + syncthreads + commit of nothing + wait
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            umma_commit(&bar);
        }
        mbar_wait(&bar, phase);
        phase ^= 1;
        tc_fence_after();
    }
    t1 = clock64();
    if (tid == 0) out[2] = (t1 - t0) / reps;
    // (3) tmem_ld32 x4 (128 columns) + 128 FMAs
    float acc = 0.f;
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        for (int c0 = 0; c0 < 128; c0 += 32) {
            float v[32];
            tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) acc += v[i];
        }
    }
    t1 = clock64();
    if (tid == 0) out[3] = (t1 - t0) / reps;
    // (4) 16 x 16-byte shared stores (one operand row) + fence.proxy.async
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int c = 0; c < 16; ++c) sA[(c % (K / 8)) * 128 + tid] = make_uint4(r, c, tid, 0);
        fence_proxy_async();
    }
    t1 = clock64();
    if (tid == 0) out[4] = (t1 - t0) / reps;
    // (5) __syncthreads alone
    t0 = clock64();
    for (int r = 0; r < reps; ++r) __syncthreads();
    t1 = clock64();
    if (tid == 0) out[5] = (t1 - t0) / reps;
    if (acc == 12345.f) out[7] = 1;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace kpf

extern "C" int kpf_umma_probe(long long* out, int N, int K, int reps, cudaStream_t stream) {
    using namespace kpf;
    const size_t smem = (size_t)(128 + N) * K * 2;
    cudaError_t e = kpf::set_smem(umma_probe_kernel, smem);
    if (e != cudaSuccess) return (int)e;
    umma_probe_kernel<<<1, 128, smem, stream>>>(out, N, K, reps);
    KPF_CHECK_LAUNCH();
    return 0;
}

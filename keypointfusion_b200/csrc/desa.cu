// DESA local point aggregation (model/model.py:129-204), SURVEY.md 8f-1.  The reference groups points with
// pointnet2_ops 3.0.0 QueryAndGroup (CUDA extension, not vendored); this file provides the B200 replacement.
// Oracle: oracle/kpf_oracle.py ball_query / desa ("parity unpinned": no reference test covers DESA).
#include "common.cuh"

namespace kpf {

// ball_query: for each centre the first `nsample` point indices (ascending) with d2 < r*r; remaining slots = first
// hit; zero if no hit.  One warp per (sample, centre): ballot-ordered compaction over the point stream.
__global__ void __launch_bounds__(256)
ball_query_kernel(const float* __restrict__ xyz, const float* __restrict__ centers, int Np, int J, float radius, int nsample,
                  int32_t* __restrict__ idx_out) {
    const int b = blockIdx.y, lane = threadIdx.x & 31;
    const int j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (j >= J) return;
    const float* c = centers + ((size_t)b * J + j) * 3;
    const float cx = c[0], cy = c[1], cz = c[2];
    const float r2 = xmul(radius, radius);
    int32_t* o = idx_out + ((size_t)b * J + j) * nsample;
    int cnt = 0, first = 0;
    for (int base = 0; base < Np && cnt < nsample; base += 32) {
        const int n = base + lane;
        bool hit = false;
        if (n < Np) {
            const float* p = xyz + ((size_t)b * Np + n) * 3;
            const float dx = xsub(cx, p[0]), dy = xsub(cy, p[1]), dz = xsub(cz, p[2]);
            hit = xadd(xadd(xmul(dx, dx), xmul(dy, dy)), xmul(dz, dz)) < r2;
        }
        const uint32_t bal = __ballot_sync(0xffffffffu, hit);
        if (bal) {
            if (cnt == 0) first = base + __ffs(bal) - 1;
            const int slot = cnt + __popc(bal & ((1u << lane) - 1u));
            if (hit && slot < nsample) o[slot] = n;
            cnt += __popc(bal);
        }
    }
    if (cnt > nsample) cnt = nsample;
    for (int s = cnt + lane; s < nsample; s += 32) o[s] = first;
}

}  // namespace kpf

extern "C" int kpf_ball_query(const float* xyz, const float* centers, int B, int Np, int J, float radius, int nsample,
                              int32_t* idx_out, cudaStream_t stream) {
    using namespace kpf;
    KPF_REQUIRE(B >= 0 && Np >= 1 && J >= 1 && nsample >= 1);
    if (B == 0) return 0;
    dim3 grid((J + 7) / 8, B);
    ball_query_kernel<<<grid, 256, 0, stream>>>(xyz, centers, Np, J, radius, nsample, idx_out);
    KPF_CHECK_LAUNCH();
    return 0;
}

// allreduce.cu -- one-shot all-reduce (sum) of the flat gradient buffer over NVLink peer memory.
//
// The data-parallel step exchanges ONE small buffer per step (54 K floats at config D: latency-bound).
// Every rank keeps a symmetric "send" buffer that its gradient kernels write; this kernel then
//   1. tells every peer that the local buffer is complete (release store of the step's epoch into the peer's flag
//      slot, system scope) and waits until all peers have said the same,
//   2. reads all `world` buffers through peer pointers (P2P loads over NVLink / NVSwitch) and writes their sum, taken
//      in rank order on every rank -- bit-identical results everywhere, run to run -- into the local result buffer,
//   3. (last CTA) tells every peer that its reads are done and waits for the peers' reads of the local buffer, so the
//      next step may overwrite it as soon as this kernel has completed.
// One launch, no host synchronisation, no NCCL call on the path.
#include "common.cuh"

namespace nnue {

constexpr int kArMaxWorld = 16;
struct ArPeers {
    const float *buf[kArMaxWorld];
    int *flags[kArMaxWorld];  // per rank: int[2][kArMaxWorld] -- [0][r] "r's buffer is ready", [1][r] "r has read mine"
};

__device__ __forceinline__ void st_release_sys(int *p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_peer_f4(const float4 *p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_peer_f(const float *p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(512)
allreduce_oneshot_kernel(const ArPeers p, int rank, int world, size_t n, float *__restrict__ out, int epoch,
                         unsigned *__restrict__ counter) {
    __shared__ bool last;
    // 1. my buffer is complete (written by earlier kernels of this stream): publish, then wait for every peer
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();
        st_release_sys(p.flags[threadIdx.x] + rank, epoch);
    }
    if (threadIdx.x < world)
        while (ld_acquire_sys(p.flags[rank] + threadIdx.x) < epoch) {}
    __syncthreads();
    // 2. sum in rank order
    const size_t n4 = n / 4, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 acc = ld_peer_f4(reinterpret_cast<const float4 *>(p.buf[0]) + i);
        for (int r = 1; r < world; ++r) acc = f4_add(acc, ld_peer_f4(reinterpret_cast<const float4 *>(p.buf[r]) + i));
        reinterpret_cast<float4 *>(out)[i] = acc;
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        float acc = ld_peer_f(p.buf[0] + i);
        for (int r = 1; r < world; ++r) acc += ld_peer_f(p.buf[r] + i);
        out[i] = acc;
    }
    // 3. the last CTA to finish tells the peers that this rank's reads are done and waits for theirs
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last) {
        if (threadIdx.x < world) {
            st_release_sys(p.flags[threadIdx.x] + kArMaxWorld + rank, epoch);
            while (ld_acquire_sys(p.flags[rank] + kArMaxWorld + threadIdx.x) < epoch) {}
        }
        __syncthreads();
        if (threadIdx.x == 0) *counter = 0u;
    }
}

}  // namespace nnue

using namespace nnue;

extern "C" {

int nnue_allreduce_max_world(void) { return kArMaxWorld; }

int nnue_allreduce_oneshot(int world, int rank, const void *const *peer_bufs_h, void *const *peer_flags_h, void *counter_d,
                           size_t n, float *out_d, int epoch, void *stream) {
    if (world < 1 || world > kArMaxWorld || rank < 0 || rank >= world || !peer_bufs_h || !peer_flags_h || !counter_d ||
        !out_d || n < 1 || epoch < 1)
        return NNUE_ERR_INVALID_ARG;
    ArPeers p{};
    for (int r = 0; r < world; ++r) {
        p.buf[r] = static_cast<const float *>(peer_bufs_h[r]);
        p.flags[r] = static_cast<int *>(peer_flags_h[r]);
        if (!p.buf[r] || !p.flags[r]) return NNUE_ERR_INVALID_ARG;
    }
    size_t want = (n / 4 + 511) / 512;
    int grid = (int)(want < 1 ? 1 : want > 64 ? 64 : want);  // every CTA spins on the flags: keep them co-resident
    allreduce_oneshot_kernel<<<grid, 512, 0, static_cast<cudaStream_t>(stream)>>>(
        p, rank, world, n, out_d, epoch, static_cast<unsigned *>(counter_d));
    NNUE_CHECK_LAUNCH("allreduce_oneshot_kernel");
    return NNUE_OK;
}

}  // extern "C"

// allreduce.cu -- one-shot all-reduce (sum) of the flat gradient buffer over NVLink peer memory.
//
// The data-parallel step exchanges ONE small buffer per step (54 K floats at config D: latency-bound).
// Push form: every rank owns a symmetric receive area of 2 (epoch parity) x world slots.  One launch
//   1. copies the local buffer into slot [parity][rank] of EVERY rank (posted stores over NVLink / NVSwitch: one-way
//      latency, no read round trip),
//   2. when all CTAs have pushed, publishes the step's epoch in every peer's flag slot (release, system scope) and
//      waits until all peers' epochs have arrived,
//   3. sums the `world` local slots in rank order -- bit-identical results on every rank, run to run -- in place.
// The parity double-buffering makes a trailing barrier unnecessary: a slot is rewritten two steps later, and a rank
// can only get there after every peer has published the epoch in between, which it does after finishing its sums.
// No host synchronisation, no NCCL call on the path.  The step counter (epoch) lives in DEVICE memory and is advanced by
// the kernel itself, so the launch has no per-step host argument and can sit inside a captured CUDA graph; a call
// reduces any 16-byte-aligned slice of the buffer, so the step can exchange the gradients that are final early (head,
// feature transformer) on a side stream while the conv gradient is still running, and only the last few hundred floats
// at the end (train.py DataParallelStep).
#include "common.cuh"

namespace nnue {

constexpr int kArMaxWorld = 16;
struct ArPeers {
    float *recv[kArMaxWorld];  // rank r's receive area: float[2][world][n]
    int *flags[kArMaxWorld];   // rank r's flags: int[kArMaxWorld], flags[s] = last epoch whose data from rank s is complete
};

__device__ __forceinline__ void st_release_sys(int *p, int v) {
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int *p) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_sys_f4(const float4 *p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_sys_f4(float4 *p, float4 v) {
    asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// n4 = float4 elements (the buffer is padded to a multiple of 4 floats); every CTA owns the same index range in the
// push and in the sum, so the sum may overwrite `buf` in place
__global__ void __launch_bounds__(256)
allreduce_push_kernel(const ArPeers p, int rank, int world, size_t n4, float4 *__restrict__ buf,
                      unsigned *__restrict__ state) {
    __shared__ bool last;
    unsigned *counter = state;  // CTAs that have pushed
    // state[1] = epoch of the previous launch: every CTA reads it before its own arrival on `counter`, the last CTA to
    // arrive writes the new value -- after every read
    const int epoch = (int)*reinterpret_cast<volatile unsigned *>(state + 1) + 1;
    const size_t stride = (size_t)gridDim.x * blockDim.x, first = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t slot0 = (size_t)(epoch & 1) * world * n4;
    // 1. push my values into slot [parity][rank] of every rank
    for (size_t i = first; i < n4; i += stride) {
        const float4 v = buf[i];
        for (int r = 0; r < world; ++r) st_sys_f4(reinterpret_cast<float4 *>(p.recv[r]) + slot0 + (size_t)rank * n4 + i, v);
    }
    // 2. all CTAs pushed -> publish the epoch to every peer; then wait for every peer's epoch
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last) {
        __threadfence_system();  // the other CTAs' pushes (ordered before their counter arrivals) before the flags go out
        if (threadIdx.x < world) st_release_sys(p.flags[threadIdx.x] + rank, epoch);
        if (threadIdx.x == 0) {
            *counter = 0u;
            state[1] = (unsigned)epoch;
        }
    }
    if (threadIdx.x < world)
        while (ld_acquire_sys(p.flags[rank] + threadIdx.x) < epoch) {}
    __syncthreads();
    // 3. sum my slots in rank order
    const float4 *mine = reinterpret_cast<const float4 *>(p.recv[rank]) + slot0;
    for (size_t i = first; i < n4; i += stride) {
        float4 acc = ld_sys_f4(mine + i);
        for (int r = 1; r < world; ++r) acc = f4_add(acc, ld_sys_f4(mine + (size_t)r * n4 + i));
        buf[i] = acc;
    }
}

// ---- small slices: data and flag travel in ONE 8-byte store (the "LL" idea of NCCL's low-latency protocol) ----------
// Every element is pushed as {value bits, epoch} with one 8-byte system-scope store per peer (8-byte stores are single
// transactions on NVLink), into slot [epoch parity][source rank][element] of the receiver.  The receiver spins on the
// element itself until its epoch word matches and sums in rank order.  No fence, no flag round, no inter-CTA dependency:
// the latency is one kernel start plus one one-way NVLink store (measured: profiles/r2_exchange_latency.json), at twice
// the bytes -- the right trade below a few hundred KB.  Parity double-buffering as above: a rank reaches epoch e + 2 only
// after it has received every peer's e + 1, which a peer sends only after it has consumed epoch e.
__device__ __forceinline__ void st_sys_u2(uint2 *p, uint2 v) {
    asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(v.x), "r"(v.y) : "memory");
}
__device__ __forceinline__ uint2 ld_sys_u2(const uint2 *p) {
    uint2 v;
    asm volatile("ld.relaxed.sys.global.v2.u32 {%0,%1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
constexpr int kLlPerThread = 4;  // elements per thread: all pushed before the first poll
// W = upper bound of `world` the instance is unrolled for (2, 4, 8 or 16).  The polls of a round -- up to 16 (element, peer)
// pairs; more would cost registers, and the kernel must stay small enough to fit beside the conv-gradient kernel that is
// still running -- are issued TOGETHER and only the pairs whose epoch has not arrived are polled again: the first version
// polled pair after pair (28 dependent system-scope loads per thread at 8 ranks: 16 us for 55 K floats, of which one
// load latency was the wire and the rest the loop).
// One launch serves up to two slices (CTAs [0, a.grid) take slice a, the rest slice b), each in one of three phases:
//   0 push + collect (the whole exchange),  1 push only,  2 collect only (poll + rank-order sum; advances the epoch).
// The data-parallel step pushes the early slice (feature transformer, head) on the side stream as soon as the table
// gradient is final -- a push waits for nobody -- and collects it at the END of the step in the same launch that
// exchanges the last few hundred floats: by then the peers' values have long arrived, so the step has ONE point where
// ranks wait for each other instead of two, and the early slice costs its local sum (measured at 8 GPUs: both slices as
// whole exchanges 18 us per step, of which the early one -- nominally overlapped -- 8.7).
struct LlSlice {
    ArPeers p;
    size_t n;
    float *buf;
    unsigned *state;
    int phase, grid;
};
template <int W>
__global__ void __launch_bounds__(256, 3)
allreduce_ll_kernel(const LlSlice sa, const LlSlice sb, int rank, int world) {
    constexpr int KR = 16 / W < kLlPerThread ? 16 / W : kLlPerThread;  // elements polled per round
    const bool second = (int)blockIdx.x >= sa.grid;
    const LlSlice &sl = second ? sb : sa;
    const int bid = second ? (int)blockIdx.x - sa.grid : (int)blockIdx.x;
    const size_t n = sl.n;
    float *__restrict__ buf = sl.buf;
    unsigned *state = sl.state;
    const bool push = sl.phase != 2, collect = sl.phase != 1;
    const unsigned epoch = *reinterpret_cast<volatile unsigned *>(state + 1) + 1u;
    const size_t slot0 = (size_t)(epoch & 1u) * world * n;
    const size_t stride = (size_t)sl.grid * blockDim.x;
    for (size_t i0 = (size_t)bid * blockDim.x + threadIdx.x; i0 < n; i0 += stride * kLlPerThread) {
        float mine[kLlPerThread];
#pragma unroll
        for (int k = 0; k < kLlPerThread; ++k) {
            const size_t i = i0 + k * stride;
            if (i < n) {
                mine[k] = buf[i];
                if (push) {
                    const uint2 pk = make_uint2(__float_as_uint(mine[k]), epoch);
                    for (int r = 0; r < world; ++r)
                        if (r != rank) st_sys_u2(reinterpret_cast<uint2 *>(sl.p.recv[r]) + slot0 + (size_t)rank * n + i, pk);
                }
            }
        }
        if (!collect) continue;
        const uint2 *in = reinterpret_cast<const uint2 *>(sl.p.recv[rank]) + slot0;
#pragma unroll
        for (int k0 = 0; k0 < kLlPerThread; k0 += KR) {
            float v[KR][W];
            unsigned pending = 0u;  // bit kk * W + r: the value of element k0 + kk from rank r has not arrived yet
#pragma unroll
            for (int kk = 0; kk < KR; ++kk)
#pragma unroll
                for (int r = 0; r < W; ++r) {
                    v[kk][r] = 0.0f;
                    if (i0 + (size_t)(k0 + kk) * stride < n && r < world && r != rank) pending |= 1u << (kk * W + r);
                }
            while (pending) {
                uint2 got[KR][W];
#pragma unroll
                for (int kk = 0; kk < KR; ++kk)
#pragma unroll
                    for (int r = 0; r < W; ++r)
                        if (pending >> (kk * W + r) & 1u) got[kk][r] = ld_sys_u2(in + (size_t)r * n + i0 + (size_t)(k0 + kk) * stride);
#pragma unroll
                for (int kk = 0; kk < KR; ++kk)
#pragma unroll
                    for (int r = 0; r < W; ++r)
                        if ((pending >> (kk * W + r) & 1u) && got[kk][r].y == epoch) {
                            v[kk][r] = __uint_as_float(got[kk][r].x);
                            pending &= ~(1u << (kk * W + r));
                        }
            }
#pragma unroll
            for (int kk = 0; kk < KR; ++kk) {
                const size_t i = i0 + (size_t)(k0 + kk) * stride;
                if (i < n) {
                    float acc = 0.0f;
#pragma unroll
                    for (int r = 0; r < W; ++r)  // rank order: bit-identical on every rank
                        if (r < world) {
                            const float x = r == rank ? mine[k0 + kk] : v[kk][r];
                            acc = r == 0 ? x : acc + x;
                        }
                    buf[i] = acc;
                }
            }
        }
    }
    // the last CTA of a slice to finish its collect advances the slice's device-side epoch (every CTA read it at the top;
    // a push-only launch leaves it alone: the collect that follows computes the same epoch)
    __syncthreads();
    if (collect && threadIdx.x == 0 && atomicAdd(state, 1u) == (unsigned)sl.grid - 1) {
        state[0] = 0u;
        state[1] = epoch;
    }
}

static int ll_grid(size_t n) {
    const size_t want = (n + 256 * kLlPerThread - 1) / (256 * kLlPerThread);
    return (int)(want < 1 ? 1 : want > 64 ? 64 : want);
}
static int launch_ll(const LlSlice &a, const LlSlice &b, int rank, int world, cudaStream_t st) {
    const int grid = a.grid + b.grid;
    if (world <= 2) allreduce_ll_kernel<2><<<grid, 256, 0, st>>>(a, b, rank, world);
    else if (world <= 4) allreduce_ll_kernel<4><<<grid, 256, 0, st>>>(a, b, rank, world);
    else if (world <= 8) allreduce_ll_kernel<8><<<grid, 256, 0, st>>>(a, b, rank, world);
    else allreduce_ll_kernel<16><<<grid, 256, 0, st>>>(a, b, rank, world);
    NNUE_CHECK_LAUNCH("allreduce_ll_kernel");
    return NNUE_OK;
}

}  // namespace nnue

using namespace nnue;

extern "C" {

int nnue_allreduce_max_world(void) { return kArMaxWorld; }

// receive area: two parities x world slots of n elements; small slices (the flagged 8-byte form) need 8 bytes per element
size_t nnue_allreduce_ll_max_floats(void) { return 65536; }
size_t nnue_allreduce_recv_floats(int world, size_t n) {
    const size_t per = (n + 3) / 4 * 4;
    return 2 * (size_t)world * per * (n <= nnue_allreduce_ll_max_floats() ? 2 : 1);
}

int nnue_allreduce_oneshot(int world, int rank, void *const *peer_recv_h, void *const *peer_flags_h, void *state_d,
                           size_t n, float *buf_d, void *stream) {
    if (world < 1 || world > kArMaxWorld || rank < 0 || rank >= world || !peer_recv_h || !peer_flags_h || !state_d ||
        !buf_d || n < 1)
        return NNUE_ERR_INVALID_ARG;
    const bool ll = n <= nnue_allreduce_ll_max_floats();
    if (!ll && ((n & 3) || (reinterpret_cast<uintptr_t>(buf_d) & 15))) return NNUE_ERR_INVALID_ARG;
    ArPeers p{};
    for (int r = 0; r < world; ++r) {
        p.recv[r] = static_cast<float *>(peer_recv_h[r]);
        p.flags[r] = static_cast<int *>(peer_flags_h[r]);
        if (!p.recv[r] || !p.flags[r]) return NNUE_ERR_INVALID_ARG;
    }
    if (ll) {
        LlSlice a{}, none{};
        a.p = p; a.n = n; a.buf = buf_d; a.state = static_cast<unsigned *>(state_d); a.phase = 0; a.grid = ll_grid(n);
        return launch_ll(a, none, rank, world, static_cast<cudaStream_t>(stream));
    }
    // every CTA spins on the flags: few, small CTAs, so that they fit beside a kernel of the step that is still running
    const size_t n4 = n / 4, want = (n4 + 255) / 256;
    const int grid = (int)(want < 1 ? 1 : want > 64 ? 64 : want);
    allreduce_push_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        p, rank, world, n4, reinterpret_cast<float4 *>(buf_d), static_cast<unsigned *>(state_d));
    NNUE_CHECK_LAUNCH("allreduce_push_kernel");
    return NNUE_OK;
}


int nnue_allreduce_oneshot_slices(int world, int rank, const nnue_allreduce_slice *slices, int n_slices, void *stream) {
    if (world < 1 || world > kArMaxWorld || rank < 0 || rank >= world || !slices || n_slices < 1 || n_slices > 2)
        return NNUE_ERR_INVALID_ARG;
    LlSlice sl[2] = {};
    for (int k = 0; k < n_slices; ++k) {
        const nnue_allreduce_slice &in = slices[k];
        if (!in.peer_recv_h || !in.peer_flags_h || !in.state_d || !in.buf_d || in.n < 1 || in.phase < 0 || in.phase > 2)
            return NNUE_ERR_INVALID_ARG;
        if (in.n > nnue_allreduce_ll_max_floats()) return NNUE_ERR_UNSUPPORTED;  // phases exist for the flagged form only
        for (int r = 0; r < world; ++r) {
            sl[k].p.recv[r] = static_cast<float *>(in.peer_recv_h[r]);
            sl[k].p.flags[r] = static_cast<int *>(in.peer_flags_h[r]);
            if (!sl[k].p.recv[r]) return NNUE_ERR_INVALID_ARG;
        }
        sl[k].n = in.n; sl[k].buf = in.buf_d; sl[k].state = static_cast<unsigned *>(in.state_d); sl[k].phase = in.phase;
        sl[k].grid = ll_grid(in.n);
    }
    return launch_ll(sl[0], sl[1], rank, world, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

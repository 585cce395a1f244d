// p2p.cu — the multi-GPU optimizer step as ONE kernel over NVLink peer memory.
//
// Replaces, for several ranks on one NVSwitch box, the sequence the reference's step would run under DDP
// (train.py:197-201 + an NCCL all-reduce of every gradient): pack -> all-reduce -> finite check -> Adam.  Every
// rank owns 1/N of the hash table (the optimizer shards of NGPTrainer).  ngp_adam_step_p2p, on rank r:
//     g   = sum over ranks p of grad_p[i]            i in the slice r owns — peer loads over NVLink (reduce-scatter)
//     p,m,v <- Adam(g)                                on the owned slice only (1/N of the optimizer sweep)
//     shadow_p[i] = fp16(p[i]) for every rank p       peer stores over NVLink (all-gather of what the kernels read)
// and the (tiny, replicated) MLP weights are updated by every rank from the same peer sums in the same order, so they
// stay bit-identical without a broadcast.  No NCCL kernel, no staging buffer, no separate pack / check passes: the
// gradient crosses NVLink once as it is summed, the parameters once as fp16.
//
// Synchronisation: ngp_p2p_barrier is a one-warp kernel — every rank stores its epoch (and its local GradScaler inf
// bit) into each peer's flag block and spins until all peers' epochs have arrived; the OR of the inf bits becomes the
// step's found_inf on every rank.  One barrier before the fused kernel (all backward passes are done), one after it
// (all peers have read my gradient and written my shadow), then the local gradient buffer is cleared.  Flag slots
// alternate with the epoch's parity: a peer can be at most one barrier ahead, so a value is never overwritten before
// it has been read.  A spin that exceeds ~20 s raises a sticky error word instead of hanging the GPU.
//
// Buffers that peers touch (gradient, fp16 shadow, flags) live in cudaMalloc memory exported with CUDA IPC
// (ngp_p2p_alloc / ngp_p2p_open); the host side exchanges the 64-byte handles through torch.distributed.
#include "common.cuh"

namespace {

struct Peers {
    void* p[NGP_MAX_PEERS];
};

constexpr int kFlagWords = 2 * NGP_MAX_PEERS + 2;   // [parity][source rank], then: error word, spare

__device__ __forceinline__ void st_release_sys(uint32_t* addr, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* addr) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
    return v;
}

__global__ void p2p_barrier_kernel(Peers flags, int rank, int world, uint32_t* __restrict__ epoch_dev,
                                   int32_t* __restrict__ found_inf, long long timeout_cycles) {
    const int lane = threadIdx.x;
    const uint32_t e = *epoch_dev + 1u;
    const uint32_t bit = (found_inf != nullptr && *found_inf != 0) ? 1u : 0u;
    const int slot = (int)(e & 1u) * NGP_MAX_PEERS;
    __syncwarp();
    __threadfence_system();
    if (lane < world) st_release_sys(reinterpret_cast<uint32_t*>(flags.p[lane]) + slot + rank, e * 2u + bit);
    uint32_t any = bit;
    bool timed_out = false;
    // a barrier that already gave up once does not wait again (the run is broken; the host reports it)
    const bool broken = reinterpret_cast<const volatile uint32_t*>(flags.p[rank])[2 * NGP_MAX_PEERS] != 0u;
    if (lane < world && !broken) {
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(flags.p[rank]) + slot + lane;
        const long long t0 = clock64();
        uint32_t v = ld_acquire_sys(mine);
        while ((v >> 1) != e) {
            if (clock64() - t0 > timeout_cycles) {
                timed_out = true;
                break;
            }
            __nanosleep(64);
            v = ld_acquire_sys(mine);
        }
        any |= v & 1u;
    }
    any = __any_sync(0xffffffffu, any != 0u) ? 1u : 0u;
    timed_out = __any_sync(0xffffffffu, timed_out);
    if (lane == 0) {
        *epoch_dev = e;
        if (found_inf != nullptr) *found_inf = (int32_t)any;
        if (timed_out) reinterpret_cast<uint32_t*>(flags.p[rank])[2 * NGP_MAX_PEERS] = e;   // error word
    }
    __threadfence_system();
}

struct AdamP2PArgs {
    float beta1, beta2, eps;
    int world;
};

__device__ __forceinline__ void adam1(float& p, float g, float& m, float& v, float lr_over_bc1, float bc2_sqrt,
                                      float inv_scale, const AdamP2PArgs& a) {
    const float gg = g * inv_scale;                       // same arithmetic as optim.cu::adam1
    m = m + (gg - m) * (1.0f - a.beta1);
    v = v * a.beta2 + (1.0f - a.beta2) * gg * gg;
    const float denom = sqrtf(v) / bc2_sqrt + a.eps;
    p = p - lr_over_bc1 * (m / denom);
}

// [lo4, hi4): float4 range of the owned table slice (shadow broadcast to all peers);
// [rep_lo4, rep_hi4): float4 range updated by every rank (MLP weights; shadow written locally only)
__global__ void __launch_bounds__(256) adam_p2p_kernel(float* __restrict__ param, Peers grads,
                                                       float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                       Peers shadows, int rank, const int32_t* __restrict__ found_inf,
                                                       const float* __restrict__ hyper, AdamP2PArgs a, int64_t lo4,
                                                       int64_t hi4, int64_t rep_lo4, int64_t rep_hi4) {
    if (found_inf != nullptr && *found_inf != 0) return;   // GradScaler: the step is skipped on every rank
    const float lr_over_bc1 = hyper[0], bc2_sqrt = hyper[1], inv_scale = hyper[2];
    const int64_t own = hi4 - lo4, total = own + (rep_hi4 - rep_lo4);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < total; k += stride) {
        const bool owned = k < own;
        const int64_t i = owned ? lo4 + k : rep_lo4 + (k - own);
        float4 g[NGP_MAX_PEERS];
#pragma unroll
        for (int p = 0; p < NGP_MAX_PEERS; ++p)            // all peer loads in flight before the first use
            if (p < a.world) g[p] = reinterpret_cast<const float4*>(grads.p[p])[i];
        float4 s = g[0];
#pragma unroll
        for (int p = 1; p < NGP_MAX_PEERS; ++p)            // fixed order 0..N-1: identical on every rank
            if (p < a.world) {
                s.x += g[p].x;
                s.y += g[p].y;
                s.z += g[p].z;
                s.w += g[p].w;
            }
        float4 w = reinterpret_cast<float4*>(param)[i];
        float4 m = reinterpret_cast<float4*>(exp_avg)[i];
        float4 v = reinterpret_cast<float4*>(exp_avg_sq)[i];
        adam1(w.x, s.x, m.x, v.x, lr_over_bc1, bc2_sqrt, inv_scale, a);
        adam1(w.y, s.y, m.y, v.y, lr_over_bc1, bc2_sqrt, inv_scale, a);
        adam1(w.z, s.z, m.z, v.z, lr_over_bc1, bc2_sqrt, inv_scale, a);
        adam1(w.w, s.w, m.w, v.w, lr_over_bc1, bc2_sqrt, inv_scale, a);
        reinterpret_cast<float4*>(param)[i] = w;
        reinterpret_cast<float4*>(exp_avg)[i] = m;
        reinterpret_cast<float4*>(exp_avg_sq)[i] = v;
        const __half2 h0 = __floats2half2_rn(w.x, w.y), h1 = __floats2half2_rn(w.z, w.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&h0);
        pk.y = *reinterpret_cast<const uint32_t*>(&h1);
        if (owned) {
#pragma unroll
            for (int p = 0; p < NGP_MAX_PEERS; ++p)
                if (p < a.world) reinterpret_cast<uint2*>(shadows.p[p])[i] = pk;
        } else {
            reinterpret_cast<uint2*>(shadows.p[rank])[i] = pk;
        }
    }
}

int fill_peers(Peers* out, void* const* ptrs, int world) {
    NGP_REQUIRE(ptrs != nullptr, "null peer pointer table");
    for (int p = 0; p < NGP_MAX_PEERS; ++p) out->p[p] = p < world ? ptrs[p] : nullptr;
    for (int p = 0; p < world; ++p) NGP_REQUIRE(out->p[p] != nullptr, "null peer pointer");
    return 0;
}

}  // namespace

extern "C" {

int ngp_p2p_alloc(int64_t bytes, void** dev_ptr, uint8_t* handle64) {
    NGP_REQUIRE(bytes > 0 && dev_ptr && handle64, "bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == NGP_IPC_HANDLE_BYTES, "IPC handle size");
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, (size_t)bytes);
    if (e == cudaSuccess) e = cudaMemset(p, 0, (size_t)bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        if (p) cudaFree(p);
        cudaGetLastError();
        ngp::set_error("ngp_p2p_alloc: %s", cudaGetErrorString(e));
        return (int)e;
    }
    memcpy(handle64, &h, sizeof(h));
    *dev_ptr = p;
    return 0;
}

int ngp_p2p_open(const uint8_t* handle64, void** peer_ptr) {
    NGP_REQUIRE(handle64 && peer_ptr, "bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
        cudaGetLastError();
        ngp::set_error("ngp_p2p_open: %s", cudaGetErrorString(e));
        return (int)e;
    }
    *peer_ptr = p;
    return 0;
}

int ngp_p2p_close(void* peer_ptr) {
    if (peer_ptr == nullptr) return 0;
    cudaError_t e = cudaIpcCloseMemHandle(peer_ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        ngp::set_error("ngp_p2p_close: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

int ngp_p2p_free(void* dev_ptr) {
    if (dev_ptr == nullptr) return 0;
    cudaError_t e = cudaFree(dev_ptr);
    if (e != cudaSuccess) {
        cudaGetLastError();
        ngp::set_error("ngp_p2p_free: %s", cudaGetErrorString(e));
        return (int)e;
    }
    return 0;
}

int64_t ngp_p2p_flag_bytes(void) { return (int64_t)kFlagWords * (int64_t)sizeof(uint32_t); }

int ngp_p2p_barrier(void* const* flag_blocks, int rank, int world, uint32_t* epoch_dev, int32_t* found_inf_or_null,
                    void* stream) {
    NGP_REQUIRE(world >= 1 && world <= NGP_MAX_PEERS && rank >= 0 && rank < world, "bad rank / world");
    NGP_REQUIRE(epoch_dev != nullptr, "null epoch counter");
    Peers f;
    if (int rc = fill_peers(&f, flag_blocks, world)) return rc;
    p2p_barrier_kernel<<<1, 32, 0, ngp::as_stream(stream)>>>(f, rank, world, epoch_dev, found_inf_or_null,
                                                             40000000000LL /* ~20 s at 1.9 GHz */);
    NGP_LAUNCHED("p2p_barrier_kernel");
    return 0;
}

int ngp_adam_step_p2p(float* param, void* const* grad_peers, float* exp_avg, float* exp_avg_sq,
                      void* const* shadow_peers, int rank, int world, const int32_t* found_inf, const float* hyper_dev,
                      float beta1, float beta2, float eps, int64_t own_begin, int64_t own_end, int64_t rep_begin,
                      int64_t rep_end, void* stream) {
    NGP_REQUIRE(world >= 1 && world <= NGP_MAX_PEERS && rank >= 0 && rank < world, "bad rank / world");
    NGP_REQUIRE(param && exp_avg && exp_avg_sq && hyper_dev, "null pointer");
    NGP_REQUIRE(own_begin >= 0 && own_begin <= own_end && rep_begin >= 0 && rep_begin <= rep_end, "bad ranges");
    NGP_REQUIRE(((own_begin | own_end | rep_begin | rep_end) & 3) == 0, "ranges must be multiples of 4 elements");
    Peers g, s;
    if (int rc = fill_peers(&g, grad_peers, world)) return rc;
    if (int rc = fill_peers(&s, shadow_peers, world)) return rc;
    for (int p = 0; p < world; ++p)
        NGP_REQUIRE((reinterpret_cast<uintptr_t>(g.p[p]) & 15) == 0 && (reinterpret_cast<uintptr_t>(s.p[p]) & 7) == 0,
                    "peer buffers must be 16-byte (gradient) / 8-byte (shadow) aligned");
    const int64_t work = (own_end - own_begin + rep_end - rep_begin) / 4;
    if (work == 0) return 0;
    const int64_t max_blocks = (int64_t)ngp::sm_count() * 8;
    const unsigned grid = (unsigned)min((work + 255) / 256, max_blocks);
    AdamP2PArgs a{beta1, beta2, eps, world};
    adam_p2p_kernel<<<grid, 256, 0, ngp::as_stream(stream)>>>(param, g, exp_avg, exp_avg_sq, s, rank, found_inf,
                                                             hyper_dev, a, own_begin / 4, own_end / 4, rep_begin / 4,
                                                             rep_end / 4);
    NGP_LAUNCHED("adam_p2p_kernel");
    return 0;
}

}  // extern "C"

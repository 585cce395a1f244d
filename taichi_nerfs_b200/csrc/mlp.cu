// mlp.cu — the fused NGP MLP (sigma net 32->64->16, SH, rgb net 32->64->64->3) on the
// 5th-generation tensor cores: tcgen05.mma with operands in shared memory and fp32 accumulators
// in tensor memory (TMEM), one 128-sample tile per CTA iteration.
//
// Replaces, in the reference, five torch.nn.Linear (cuBLAS) calls under torch.autocast(fp16) plus
// ~10 elementwise/cat/cast kernels: modules/networks.py:136-166 (NGP.density/forward), :369-380
// (MLP.forward), :18-30 (TruncExp) and modules/spherical_harmonics.py:16-42 (SH is fused into the
// rgb-net input stage).  Numerics follow autocast: fp16 operands, fp32 accumulate, every layer
// output rounded to fp16, TruncExp / direction normalisation / SH in fp32.
//
// Data layout.  All MMA operands live in shared memory in the canonical UMMA *no-swizzle*
// ("interleaved") layout: 8x8 fp16 core matrices of 128 contiguous bytes (row stride 16 B);
// core matrices adjacent along K are LBO = 128 B apart, 8-row groups are SBO = (K/8)*128 B apart.
// A thread owns one sample row, so an epilogue writes 16-byte chunks that are bank-conflict free
// per quarter warp, and the same buffer can later be consumed K-major (forward) or MN-major
// (weight gradients) by swapping LBO/SBO.  Weights (9,408 fp16 = 18.4 KB, W5 zero-padded to 16
// rows) stay resident in shared memory for the lifetime of the persistent CTA.
//
// Roofline: tensor pipe for the MMAs (18,816 FLOP/sample forward) but K is only 32/64, so the
// kernel is bounded by the TMEM->register->shared epilogue round trips and by 86 B/sample of HBM
// traffic (emb 64 B + dir 12 B in, sigma 4 B + rgb 6 B out); see DESIGN.md.
#include <stdlib.h>

#include "tcgen05.cuh"

namespace ngp {
int mlp_fwd_v2_launch(const void* emb_f16, const float* dirs, const ngp_mlp_weights* w, float* sigmas, void* rgbs,
                      void* save, int64_t n, const int32_t* n_dev, cudaStream_t st);
}

namespace {
using namespace ngp_tc;

constexpr int kRows = 128;          // sample rows per tile == TMEM lanes
constexpr int kThreads = 256;       // TWO threads per row: warps 0-3 own accumulator columns [0,32), warps 4-7 own [32,64)
constexpr int kThreadsBwd = 288;    // backward: + warp 8, which only issues MMAs (the 8-MMA weight-gradient batches of a
                                    // round are issued while warps 0-7 run that round's epilogue)
                                    // (a warp may touch TMEM lanes 32*(warp%4)..+31), halving every epilogue's latency
constexpr uint32_t kTmemCols = 64;  // fp32 accumulator columns (max N = 64)

// shared memory map (bytes).  Weights first (tcgen05.cuh, shared by all MLP kernels), then activation buffers.
constexpr int kB64 = kTile * 64 * 2;         // one [128 x 64] fp16 operand buffer
constexpr int kB32 = kTile * 32 * 2;
constexpr int kB16 = kTile * 16 * 2;
// forward kernel: two ping-pong buffers
constexpr int kBufA = kAct;
constexpr int kBufB = kBufA + kB64;
constexpr int kBar = kBufB + kB64;           // mbarrier (8 B) + tmem base (4 B)
constexpr int kSmemBytes = kBar + 16;
// backward kernel: every activation of the recomputed forward stays resident
constexpr int kE = kAct;                     // X = emb            [128 x 32]
constexpr int kH1 = kE + kB32;               // relu(X W1^T)       [128 x 64]   (later: dH1)
constexpr int kX3 = kH1 + kB64;              // [SH | h]           [128 x 32]
constexpr int kH3 = kX3 + kB32;              // relu(X3 W3^T)      [128 x 64]   (later: dH1)
constexpr int kH4 = kH3 + kB64;              // relu(H3 W4^T)      [128 x 64]   (later: dH3)
constexpr int kDH4 = kH4 + kB64;             // dL/dH4             [128 x 64]
constexpr int kDO = kDH4 + kB64;             // dL/do (3 of 16)    [128 x 16]
constexpr int kDH = kDO + kB16;              // dL/dh              [128 x 16]
constexpr int kBarBwd = kDH + kB16;
constexpr int kSmemBytesBwd = kBarBwd + 32;  // 110,624 B -> 2 CTAs / SM
// TMEM columns of the backward kernel: per-tile accumulator + persistent weight-gradient accumulators
constexpr uint32_t kTmemColsBwd = 256;
constexpr uint32_t kColDW4 = 64, kColDW1 = 128, kColDW3 = 160, kColDW2T = 192, kColDW5T = 208;

// ---- optional per-round clock trace (scripts/mlp_round_trace.py builds a separate library with -DNGP_MLP_TRACE)
#ifdef NGP_MLP_TRACE
__device__ long long g_mlp_trace[2 * 4 * 32];
#define NGP_TR(k)                                                                                   \
    do {                                                                                            \
        if (blockIdx.x == 0 && (tid == 0 || tid == 160) && tile_no < 4)                             \
            g_mlp_trace[(tid ? 128 : 0) + tile_no * 32 + (k)] = clock64();                          \
    } while (0)
#else
#define NGP_TR(k) ((void)0)
#endif

// per-thread inputs of one tile row, fetched one tile ahead so the ~1 us DRAM latency overlaps the
// previous tile's MMA / epilogue rounds
struct RowIn {
    uint4 e[2];          // this thread's 2 of the row's 4 embedding chunks (16 fp16 values)
    float dx, dy, dz;    // used by hh == 1 (SH)
    float dsig, dr[3];   // backward only, used by hh == 0
    uint4 hs[2];         // backward with saved activations: h (16 fp16) of this row, used by hh == 0
    uint2 rgb;           // ... and the forward's fp16 rgb output (3 used)
};
// activations saved by the forward for the backward (ngp_mlp_save_bytes): [n_max x 16] fp16 h = sigma-net output,
// then [n_max x 4] fp16 rgb (the sigmoid output, as torch's sigmoid backward keeps it)
__device__ __forceinline__ const __half* save_rgb_ptr(const __half* save, int64_t n_max) { return save + n_max * 16; }
__device__ __forceinline__ __half* save_rgb_ptr(__half* save, int64_t n_max) { return save + n_max * 16; }

template <typename TEmb, bool kBwd, bool kSaved = false>
__device__ __forceinline__ RowIn load_row(const TEmb* __restrict__ emb, const float* __restrict__ dirs,
                                           const float* __restrict__ dsigmas, const __half* __restrict__ drgbs,
                                           int64_t i, bool valid, int hh, const __half* __restrict__ save = nullptr,
                                           int64_t n_max = 0) {
    RowIn r;
    r.e[0] = r.e[1] = make_uint4(0, 0, 0, 0);
    r.hs[0] = r.hs[1] = make_uint4(0, 0, 0, 0);
    r.rgb = make_uint2(0, 0);
    r.dx = 0.f; r.dy = 0.f; r.dz = 1.f; r.dsig = 0.f; r.dr[0] = r.dr[1] = r.dr[2] = 0.f;
    if (valid) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int kc = hh * 2 + q;
            if constexpr (sizeof(TEmb) == 2) {
                r.e[q] = __ldg(reinterpret_cast<const uint4*>(emb + i * 32) + kc);
            } else {
                const float4 a = __ldg(reinterpret_cast<const float4*>(emb + i * 32 + kc * 8));
                const float4 b = __ldg(reinterpret_cast<const float4*>(emb + i * 32 + kc * 8 + 4));
                r.e[q] = make_uint4(pack_h2(a.x, a.y), pack_h2(a.z, a.w), pack_h2(b.x, b.y), pack_h2(b.z, b.w));
            }
        }
        if (hh == 1) {
            r.dx = __ldg(dirs + i * 3 + 0);
            r.dy = __ldg(dirs + i * 3 + 1);
            r.dz = __ldg(dirs + i * 3 + 2);
        } else if constexpr (kBwd) {
            r.dsig = __ldg(dsigmas + i);
#pragma unroll
            for (int c = 0; c < 3; ++c) r.dr[c] = __half2float(drgbs[i * 3 + c]);
            if constexpr (kSaved) {
                r.hs[0] = __ldg(reinterpret_cast<const uint4*>(save + i * 16));
                r.hs[1] = __ldg(reinterpret_cast<const uint4*>(save + i * 16) + 1);
                r.rgb = __ldg(reinterpret_cast<const uint2*>(save_rgb_ptr(save, n_max) + i * 4));
            }
        }
    }
    return r;
}

template <typename TEmb>
__global__ void __launch_bounds__(kThreads, 4) mlp_fwd_kernel(const TEmb* __restrict__ emb, const float* __restrict__ dirs,
                                                           ngp_mlp_weights w, float* __restrict__ sigmas,
                                                           __half* __restrict__ rgbs, __half* __restrict__ save,
                                                           int64_t n_max, const int32_t* __restrict__ n_dev) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int64_t n = n_dev ? min(n_max, max((int64_t)*n_dev, (int64_t)0)) : n_max;
    const int tid = threadIdx.x, warp = tid >> 5, row = tid & (kRows - 1), hh = tid >> 7;
    const uint32_t bar = smem_u32(smem + kBar);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kBar + 8);

    // ---- one-time setup: weights -> smem, mbarrier, TMEM allocation
    stage_weight(smem + kW1, w.w1, 64, 64, 32, kThreads);
    stage_weight(smem + kW2, w.w2, 16, 16, 64, kThreads);
    stage_weight(smem + kW3, w.w3, 64, 64, 32, kThreads);
    stage_weight(smem + kW4, w.w4, 64, 64, 64, kThreads);
    stage_weight(smem + kW5, w.w5, 3, 16, 64, kThreads);
    if (tid == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);  // this warp's 32 lanes
    uint32_t phase = 0;

    const uint32_t aA = smem_u32(smem + kBufA), aB = smem_u32(smem + kBufB);
    const uint32_t aW1 = smem_u32(smem + kW1), aW2 = smem_u32(smem + kW2), aW3 = smem_u32(smem + kW3),
                   aW4 = smem_u32(smem + kW4), aW5 = smem_u32(smem + kW5);

    const int64_t n_tiles = (n + kTile - 1) / kTile;
    RowIn cur = load_row<TEmb, false>(emb, dirs, nullptr, nullptr, (int64_t)blockIdx.x * kTile + row,
                                      (int64_t)blockIdx.x * kTile + row < n, hh);
    [[maybe_unused]] int tile_no = -1;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        ++tile_no;
        NGP_TR(31);
        const int64_t i = tile * kTile + row;
        const bool valid = i < n;

        // ---- stage X = emb[i, 0:32] as fp16 into bufA (K = 32 layout); inputs were fetched one tile ago
#pragma unroll
        for (int q = 0; q < 2; ++q) *reinterpret_cast<uint4*>(smem + kBufA + chunk_off(row, hh * 2 + q, 32)) = cur.e[q];
        const float dx = cur.dx, dy = cur.dy, dz = cur.dz;
        {   // prefetch the next tile of this CTA
            const int64_t in = (tile + gridDim.x) * kTile + row;
            cur = load_row<TEmb, false>(emb, dirs, nullptr, nullptr, in, in < n, hh);
        }
        fence_proxy_async();
        tc_fence_before();
        __syncthreads();

        // ---- layer 1: H1 = relu(X W1^T)                       [128x32]x[32x64]
        NGP_TR(0);
        if (tid == 0) {
            tc_fence_after();
            issue_layer(tmem_base, aA, aW1, 32, 64, bar);
        }
        NGP_TR(1);
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        NGP_TR(2);
        epilogue_hidden(tmem_row, smem + kBufB, row, hh);
        NGP_TR(3);
        fence_proxy_async();
        tc_fence_before();
        NGP_TR(4);
        __syncthreads();
        NGP_TR(5);

        // ---- layer 2: h = H1 W2^T                             [128x64]x[64x16]
        NGP_TR(6);
        if (tid == 0) {
            tc_fence_after();
            issue_layer(tmem_base, aB, aW2, 64, 16, bar);
        }
        NGP_TR(7);
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        NGP_TR(8);
        if (hh == 0) {
            // sigma = TruncExp(h[:,0]) (networks.py:22-24, :146) and the geometry half of X3 = [SH | h]
            float h[16];
            tmem_ld16(tmem_row, h);
            const float h0 = __half2float(__float2half_rn(h[0]));
            if (valid) sigmas[i] = expf(h0);
            uint8_t* dst = smem + kBufA;
            const uint4 lo = make_uint4(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]), pack_h2(h[4], h[5]), pack_h2(h[6], h[7]));
            const uint4 hi = make_uint4(pack_h2(h[8], h[9]), pack_h2(h[10], h[11]), pack_h2(h[12], h[13]), pack_h2(h[14], h[15]));
            *reinterpret_cast<uint4*>(dst + chunk_off(row, 2, 32)) = lo;
            *reinterpret_cast<uint4*>(dst + chunk_off(row, 3, 32)) = hi;
            if (save != nullptr && valid) {   // the backward restarts from h instead of recomputing layers 1-2 serially
                reinterpret_cast<uint4*>(save + i * 16)[0] = lo;
                reinterpret_cast<uint4*>(save + i * 16)[1] = hi;
            }
        } else {
            // the direction half, in parallel: SH16((d/|d| + 1)/2)  (networks.py:162-164)
            const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
            float e[16];
            sh16((dx * inv + 1.0f) / 2.0f, (dy * inv + 1.0f) / 2.0f, (dz * inv + 1.0f) / 2.0f, e);
            uint8_t* dst = smem + kBufA;
            *reinterpret_cast<uint4*>(dst + chunk_off(row, 0, 32)) =
                make_uint4(pack_h2(e[0], e[1]), pack_h2(e[2], e[3]), pack_h2(e[4], e[5]), pack_h2(e[6], e[7]));
            *reinterpret_cast<uint4*>(dst + chunk_off(row, 1, 32)) =
                make_uint4(pack_h2(e[8], e[9]), pack_h2(e[10], e[11]), pack_h2(e[12], e[13]), pack_h2(e[14], e[15]));
        }
        NGP_TR(9);
        fence_proxy_async();
        tc_fence_before();
        NGP_TR(10);
        __syncthreads();
        NGP_TR(11);

        // ---- layer 3: H3 = relu(X3 W3^T)                      [128x32]x[32x64]
        NGP_TR(12);
        if (tid == 0) {
            tc_fence_after();
            issue_layer(tmem_base, aA, aW3, 32, 64, bar);
        }
        NGP_TR(13);
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        NGP_TR(14);
        epilogue_hidden(tmem_row, smem + kBufB, row, hh);
        NGP_TR(15);
        fence_proxy_async();
        tc_fence_before();
        NGP_TR(16);
        __syncthreads();
        NGP_TR(17);

        // ---- layer 4: H4 = relu(H3 W4^T)                      [128x64]x[64x64]
        NGP_TR(18);
        if (tid == 0) {
            tc_fence_after();
            issue_layer(tmem_base, aB, aW4, 64, 64, bar);
        }
        NGP_TR(19);
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        NGP_TR(20);
        epilogue_hidden(tmem_row, smem + kBufA, row, hh);
        NGP_TR(21);
        fence_proxy_async();
        tc_fence_before();
        NGP_TR(22);
        __syncthreads();
        NGP_TR(23);

        // ---- layer 5: rgb = sigmoid(H4 W5^T)                  [128x64]x[64x16(3 used)]
        NGP_TR(24);
        if (tid == 0) {
            tc_fence_after();
            issue_layer(tmem_base, aA, aW5, 64, 16, bar);
        }
        NGP_TR(25);
        mbar_wait(bar, phase);
        phase ^= 1;
        tc_fence_after();
        NGP_TR(26);
        if (hh == 0) {
            float o[16];
            tmem_ld16(tmem_row, o);
            if (valid) {
                __half out[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float oc = __half2float(__float2half_rn(o[c]));
                    out[c] = __float2half_rn(1.0f / (1.0f + expf(-oc)));
                    rgbs[i * 3 + c] = out[c];
                }
                if (save != nullptr) {
                    const __half2 a = __halves2half2(out[0], out[1]), b = __halves2half2(out[2], __float2half_rn(0.0f));
                    *reinterpret_cast<uint2*>(save_rgb_ptr(save, n_max) + i * 4) =
                        make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
                }
            }
        }
        NGP_TR(27);
        tc_fence_before();
        __syncthreads();  // bufA / TMEM are reused by the next tile
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// =====================================================================================================
// Backward.  Replaces the autograd graph of the five nn.Linear layers (cuBLAS dX + dW GEMMs), ReLU /
// Sigmoid / TruncExp backward (modules/networks.py:26-30) under autocast.  Per 128-sample tile:
//   1. recompute the forward activations into shared memory (nothing was saved by the forward);
//   2. chain  dO -> dH4 -> dH3 -> dX3 -> dH -> dH1 -> dE  with  dX = dY * W  as tcgen05 MMAs whose B
//      operand is the SAME shared-memory weight copy read MN-major (i.e. transposed by descriptor);
//   3. weight gradients  dW = dY^T * X  as M=64 MMAs with K = the 128 samples of the tile, both operands
//      read MN-major from the activation buffers, accumulated in TMEM across ALL tiles of the persistent
//      CTA and flushed once at the end with fp32 atomics (9,408 per CTA).
// Gradient operands of invalid (tail) rows are zero, so they do not contribute to dW.

// generic GEMM issue: operand = (start address, LBO, SBO, bytes to advance per 16-wide K step)
struct Operand {
    uint32_t addr, lbo, sbo, kstep;
};
// operand stored as [row][K cols] (row-block stride K*16 B) and consumed K-major (rows = M or N)
__device__ __forceinline__ Operand op_kmajor(uint32_t addr, int K) { return {addr, 128u, (uint32_t)K * 16u, 256u}; }
// the same storage consumed MN-major: MN = the stored columns, K = the stored rows
__device__ __forceinline__ Operand op_mnmajor(uint32_t addr, int K) { return {addr, (uint32_t)K * 16u, 128u, (uint32_t)K * 32u}; }

__host__ __device__ constexpr uint32_t idesc_full(int m, int n, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, const Operand& a, const Operand& b, int ksteps,
                                           uint32_t idesc, bool accumulate) {
    for (int k = 0; k < ksteps; ++k) {
        const uint64_t da = smem_desc(a.addr + k * a.kstep, a.lbo, a.sbo);
        const uint64_t db = smem_desc(b.addr + k * b.kstep, b.lbo, b.sbo);
        umma_f16(tmem_d, da, db, idesc, (accumulate || k > 0) ? 1u : 0u);
    }
}

// backward hidden epilogue: TMEM [128 x 64] fp32 -> fp16, masked by relu'(act) where `act` holds the
// post-ReLU forward activation of this thread's row -> dst (K = 64 layout).  dst may alias act.
__device__ __forceinline__ void epilogue_relu_bwd(uint32_t tmem_row, const uint8_t* act, uint8_t* dst, int row, int hh) {
    float v[32];
    tmem_ld32(tmem_row + hh * 32, v);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int kc = hh * 4 + q;
        const uint4 a = *reinterpret_cast<const uint4*>(act + chunk_off(row, kc, 64));
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
        uint32_t o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // relu'(act) as a per-half bit mask (0xffff where act > 0): select semantics like threshold_backward,
            // one HSET2.BM + LOP3 per pair instead of two unpack / compare / select chains
            const __half2 ah = *reinterpret_cast<const __half2*>(&aw[j]);
            o[j] = pack_h2(v[q * 8 + 2 * j], v[q * 8 + 2 * j + 1]) & __hgt2_mask(ah, __float2half2_rn(0.0f));
        }
        *reinterpret_cast<uint4*>(dst + chunk_off(row, kc, 64)) = make_uint4(o[0], o[1], o[2], o[3]);
    }
}

template <typename TEmb, bool kSaved>
__global__ void __launch_bounds__(kThreadsBwd, 2) mlp_bwd_kernel(const TEmb* __restrict__ emb, const float* __restrict__ dirs,
                                                           ngp_mlp_weights w, const __half* __restrict__ save,
                                                           const float* __restrict__ dsigmas,
                                                           const __half* __restrict__ drgbs, TEmb* __restrict__ demb,
                                                           float* __restrict__ grad_w, int64_t n_max,
                                                           const int32_t* __restrict__ n_dev,
                                                           int32_t* __restrict__ found_inf) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int64_t n = n_dev ? min(n_max, max((int64_t)*n_dev, (int64_t)0)) : n_max;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, row = tid & (kRows - 1), hh = tid >> 7;
    const bool worker = tid < kThreads;   // warps 0-7: two threads per tile row; warp 8 (hh == 2): MMA issue only
    const uint32_t bar = smem_u32(smem + kBarBwd);
    const uint32_t bar2 = smem_u32(smem + kBarBwd + 16);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kBarBwd + 8);
    uint32_t phase2 = 0;

    stage_weight(smem + kW1, w.w1, 64, 64, 32, kThreads);
    stage_weight(smem + kW2, w.w2, 16, 16, 64, kThreads);
    stage_weight(smem + kW3, w.w3, 64, 64, 32, kThreads);
    stage_weight(smem + kW4, w.w4, 64, 64, 64, kThreads);
    stage_weight(smem + kW5, w.w5, 3, 16, 64, kThreads);
    if (tid == 0) {
        mbar_init(bar, 1);
        mbar_init(bar2, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_slot), kTmemColsBwd);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_row = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t phase = 0;

    const uint32_t aW1 = smem_u32(smem + kW1), aW2 = smem_u32(smem + kW2), aW3 = smem_u32(smem + kW3),
                   aW4 = smem_u32(smem + kW4), aW5 = smem_u32(smem + kW5);
    const uint32_t aE = smem_u32(smem + kE), aH1 = smem_u32(smem + kH1), aX3 = smem_u32(smem + kX3),
                   aH3 = smem_u32(smem + kH3), aH4 = smem_u32(smem + kH4), aDH4 = smem_u32(smem + kDH4),
                   aDO = smem_u32(smem + kDO), aDH = smem_u32(smem + kDH);

// One MMA round.  DX feeds the next epilogue and is committed to `bar`; DW (weight-gradient MMAs, may be
// empty) is issued right behind it and runs while the epilogue executes.  tcgen05.commit tracks ALL prior
// MMAs of the issuing thread, so the commit of round r+1 also covers DW of round r: every buffer a DW
// reads is only overwritten after a later round's wait — except the last round's, which commits to `bar2`.
#define NGP_ROUND2(DX, DW, LAST)               \
    fence_proxy_async();                       \
    tc_fence_before();                         \
    __syncthreads();                           \
    if (tid == kThreads) {                     \
        tc_fence_after();                      \
        DX;                                    \
        umma_commit(bar);                      \
        DW;                                    \
        if (LAST) umma_commit(bar2);           \
    }                                          \
    mbar_wait(bar, phase);                     \
    phase ^= 1;                                \
    tc_fence_after();
#define NGP_ROUND(ISSUE) NGP_ROUND2(ISSUE, (void)0, false)

    bool first = true;  // first tile of this CTA: weight-gradient accumulators start from zero
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    RowIn cur = load_row<TEmb, true, kSaved>(emb, dirs, dsigmas, drgbs, (int64_t)blockIdx.x * kTile + row,
                                             worker && (int64_t)blockIdx.x * kTile + row < n, hh, save, n_max);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t i = tile * kTile + row;
        const bool valid = worker && i < n;

        // ================= forward recompute =================
        if (worker) {
#pragma unroll
            for (int q = 0; q < 2; ++q) *reinterpret_cast<uint4*>(smem + kE + chunk_off(row, hh * 2 + q, 32)) = cur.e[q];
        }
        const float dx = cur.dx, dy = cur.dy, dz = cur.dz, dsig = cur.dsig;  // dirs: hh == 1, dsig/dr: hh == 0
        const float dr[3] = {cur.dr[0], cur.dr[1], cur.dr[2]};
        [[maybe_unused]] const uint4 hs0 = cur.hs[0], hs1 = cur.hs[1];
        [[maybe_unused]] const uint2 rgb_saved = cur.rgb;
        {   // prefetch the next tile of this CTA (consumed one iteration later)
            const int64_t in = (tile + gridDim.x) * kTile + row;
            cur = load_row<TEmb, true, kSaved>(emb, dirs, dsigmas, drgbs, in, worker && in < n, hh, save, n_max);
        }
        float h0 = 0.0f;
        if constexpr (kSaved) {
            // X3 = [SH | h] straight from the saved h: layers 2 and 5 are not recomputed (8 MMA rounds instead of 10)
            if (hh == 0) {
                const __half2 h01 = *reinterpret_cast<const __half2*>(&hs0.x);
                h0 = __low2float(h01);
                *reinterpret_cast<uint4*>(smem + kX3 + chunk_off(row, 2, 32)) = hs0;
                *reinterpret_cast<uint4*>(smem + kX3 + chunk_off(row, 3, 32)) = hs1;
            }
        }
        NGP_ROUND(issue_layer_mma(tmem_base, aE, aW1, 32, 64))          // H1 = relu(E W1^T)
        if (worker) epilogue_hidden(tmem_row, smem + kH1, row, hh);
        if constexpr (!kSaved) {
            NGP_ROUND(issue_layer_mma(tmem_base, aH1, aW2, 64, 16))     // h = H1 W2^T
            if (hh == 0) {
                float h[16];
                tmem_ld16(tmem_row, h);
                h0 = __half2float(__float2half_rn(h[0]));
                uint8_t* dst = smem + kX3;
                *reinterpret_cast<uint4*>(dst + chunk_off(row, 2, 32)) =
                    make_uint4(pack_h2(h[0], h[1]), pack_h2(h[2], h[3]), pack_h2(h[4], h[5]), pack_h2(h[6], h[7]));
                *reinterpret_cast<uint4*>(dst + chunk_off(row, 3, 32)) =
                    make_uint4(pack_h2(h[8], h[9]), pack_h2(h[10], h[11]), pack_h2(h[12], h[13]), pack_h2(h[14], h[15]));
            }
        }
        if (hh == 1) {
            const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
            float e[16];
            sh16((dx * inv + 1.0f) / 2.0f, (dy * inv + 1.0f) / 2.0f, (dz * inv + 1.0f) / 2.0f, e);
            uint8_t* dst = smem + kX3;
            *reinterpret_cast<uint4*>(dst + chunk_off(row, 0, 32)) =
                make_uint4(pack_h2(e[0], e[1]), pack_h2(e[2], e[3]), pack_h2(e[4], e[5]), pack_h2(e[6], e[7]));
            *reinterpret_cast<uint4*>(dst + chunk_off(row, 1, 32)) =
                make_uint4(pack_h2(e[8], e[9]), pack_h2(e[10], e[11]), pack_h2(e[12], e[13]), pack_h2(e[14], e[15]));
        }
        NGP_ROUND(issue_layer_mma(tmem_base, aX3, aW3, 32, 64))         // H3 = relu(X3 W3^T)
        if (worker) epilogue_hidden(tmem_row, smem + kH3, row, hh);
        NGP_ROUND(issue_layer_mma(tmem_base, aH3, aW4, 64, 64))         // H4 = relu(H3 W4^T)
        if (worker) epilogue_hidden(tmem_row, smem + kH4, row, hh);
        if constexpr (!kSaved) {
            NGP_ROUND(issue_layer_mma(tmem_base, aH4, aW5, 64, 16))     // o = H4 W5^T
        }
        if (hh == 0) {
            // dL/do = dL/drgb * rgb (1 - rgb), rounded to fp16 like the autocast graph
            float rgbv[3];
            if constexpr (kSaved) {   // torch's sigmoid backward also uses the saved fp16 output
                const __half2 a = *reinterpret_cast<const __half2*>(&rgb_saved.x);
                const __half2 b = *reinterpret_cast<const __half2*>(&rgb_saved.y);
                rgbv[0] = __low2float(a);
                rgbv[1] = __high2float(a);
                rgbv[2] = __low2float(b);
            } else {
                float o[16];
                tmem_ld16(tmem_row, o);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float oc = __half2float(__float2half_rn(o[c]));
                    rgbv[c] = __half2float(__float2half_rn(1.0f / (1.0f + expf(-oc))));
                }
            }
            float d_o[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) d_o[c] = dr[c] * rgbv[c] * (1.0f - rgbv[c]);
            *reinterpret_cast<uint4*>(smem + kDO + chunk_off(row, 0, 16)) =
                make_uint4(pack_h2(d_o[0], d_o[1]), pack_h2(d_o[2], 0.0f), 0u, 0u);
        } else if (hh == 1) {
            *reinterpret_cast<uint4*>(smem + kDO + chunk_off(row, 1, 16)) = make_uint4(0u, 0u, 0u, 0u);
        }

        // ================= backward =================
        // R1: dH4pre = dO W5 ;  dW5^T += H4^T dO
        NGP_ROUND2(
            issue_gemm(tmem_base, op_kmajor(aDO, 16), op_mnmajor(aW5, 64), 1, idesc_full(128, 64, 0, 1), false),
            issue_gemm(tmem_base + kColDW5T, op_mnmajor(aH4, 64), op_mnmajor(aDO, 16), 8, idesc_full(64, 16, 1, 1), !first),
            false)
        if (worker) epilogue_relu_bwd(tmem_row, smem + kH4, smem + kDH4, row, hh);
        // R2: dH3pre = dH4 W4 ;  dW4 += dH4^T H3
        NGP_ROUND2(
            issue_gemm(tmem_base, op_kmajor(aDH4, 64), op_mnmajor(aW4, 64), 4, idesc_full(128, 64, 0, 1), false),
            issue_gemm(tmem_base + kColDW4, op_mnmajor(aDH4, 64), op_mnmajor(aH3, 64), 8, idesc_full(64, 64, 1, 1), !first),
            false)
        if (worker) epilogue_relu_bwd(tmem_row, smem + kH3, smem + kH4, row, hh);      // dH3 -> H4's buffer (H4 is dead)
        // R3: dX3 = dH3 W3 ;  dW3 += dH3^T X3
        NGP_ROUND2(
            issue_gemm(tmem_base, op_kmajor(aH4, 64), op_mnmajor(aW3, 32), 4, idesc_full(128, 32, 0, 1), false),
            issue_gemm(tmem_base + kColDW3, op_mnmajor(aH4, 64), op_mnmajor(aX3, 32), 8, idesc_full(64, 32, 1, 1), !first),
            false)
        if (hh == 0) {
            // dh = dX3[:, 16:32] (+ TruncExp backward on h[:,0], networks.py:26-30), fp16
            float g[16];
            tmem_ld16(tmem_row + 16, g);
            const float ds = __half2float(__float2half_rn(dsig * expf(fminf(fmaxf(h0, -15.0f), 15.0f))));
            g[0] = __half2float(__float2half_rn(g[0])) + ds;
            *reinterpret_cast<uint4*>(smem + kDH + chunk_off(row, 0, 16)) =
                make_uint4(pack_h2(g[0], g[1]), pack_h2(g[2], g[3]), pack_h2(g[4], g[5]), pack_h2(g[6], g[7]));
            *reinterpret_cast<uint4*>(smem + kDH + chunk_off(row, 1, 16)) =
                make_uint4(pack_h2(g[8], g[9]), pack_h2(g[10], g[11]), pack_h2(g[12], g[13]), pack_h2(g[14], g[15]));
        }
        // R4: dH1pre = dh W2 ;  dW2^T += H1^T dh
        NGP_ROUND2(
            issue_gemm(tmem_base, op_kmajor(aDH, 16), op_mnmajor(aW2, 64), 1, idesc_full(128, 64, 0, 1), false),
            issue_gemm(tmem_base + kColDW2T, op_mnmajor(aH1, 64), op_mnmajor(aDH, 16), 8, idesc_full(64, 16, 1, 1), !first),
            false)
        if (worker) epilogue_relu_bwd(tmem_row, smem + kH1, smem + kH3, row, hh);      // dH1 -> H3's buffer (H3 is dead)
        // R5: dE = dH1 W1 ;  dW1 += dH1^T E
        NGP_ROUND2(
            issue_gemm(tmem_base, op_kmajor(aH3, 64), op_mnmajor(aW1, 32), 4, idesc_full(128, 32, 0, 1), false),
            issue_gemm(tmem_base + kColDW1, op_mnmajor(aH3, 64), op_mnmajor(aE, 32), 8, idesc_full(64, 32, 1, 1), !first),
            true)
        if (worker) {
            const int g = hh;  // columns [16*hh, 16*hh+16) of dE
            float v[16];
            tmem_ld16(tmem_row + g * 16, v);
            if (valid) {
                if constexpr (sizeof(TEmb) == 2) {
                    uint4* o = reinterpret_cast<uint4*>(demb + i * 32 + g * 16);
                    o[0] = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
                    o[1] = make_uint4(pack_h2(v[8], v[9]), pack_h2(v[10], v[11]), pack_h2(v[12], v[13]), pack_h2(v[14], v[15]));
                } else {
                    float4* o = reinterpret_cast<float4*>(demb + i * 32 + g * 16);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {  // fp16-rounded like the autocast graph, stored as fp32
                        o[q] = make_float4(__half2float(__float2half_rn(v[4 * q])), __half2float(__float2half_rn(v[4 * q + 1])),
                                           __half2float(__float2half_rn(v[4 * q + 2])), __half2float(__float2half_rn(v[4 * q + 3])));
                    }
                }
            }
        }
        first = false;
        mbar_wait(bar2, phase2);  // dW1 (reads E and dH1) must finish before the next tile restages them
        phase2 ^= 1;
        tc_fence_before();
        __syncthreads();
    }
#undef NGP_ROUND
#undef NGP_ROUND2

    // ---- flush the weight-gradient accumulators: M = 64 rows live on TMEM lanes (m%16) + 32*(m/16)
    if (!first && warp < 4) {
        const int m = warp * 16 + lane;  // row held by this thread when lane < 16
        const bool has_row = lane < 16;
        float v[16];
        bool bad = false;   // non-finite weight gradient (GradScaler's inf check, raised at the source)
        auto chk = [&]() {
#pragma unroll
            for (int j = 0; j < 16; ++j) bad = bad || !(fabsf(v[j]) < INFINITY);
        };
        // dW4 [64 out x 64 in]
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            tmem_ld16(tmem_row + kColDW4 + g * 16, v);
            if (has_row) chk();
            if (has_row)
#pragma unroll
                for (int j = 0; j < 16; ++j) atomicAdd(grad_w + (NGP_MLP_W1 + NGP_MLP_W2 + NGP_MLP_W3) + m * 64 + g * 16 + j, v[j]);
        }
        // dW1 [64 out x 32 in], dW3 [64 out x 32 in]
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            tmem_ld16(tmem_row + kColDW1 + g * 16, v);
            if (has_row) chk();
            if (has_row)
#pragma unroll
                for (int j = 0; j < 16; ++j) atomicAdd(grad_w + m * 32 + g * 16 + j, v[j]);
            tmem_ld16(tmem_row + kColDW3 + g * 16, v);
            if (has_row) chk();
            if (has_row)
#pragma unroll
                for (int j = 0; j < 16; ++j) atomicAdd(grad_w + (NGP_MLP_W1 + NGP_MLP_W2) + m * 32 + g * 16 + j, v[j]);
        }
        // dW2^T [64 in x 16 out] -> W2 is [16 out x 64 in]
        tmem_ld16(tmem_row + kColDW2T, v);
        if (has_row) chk();
        if (has_row)
#pragma unroll
            for (int j = 0; j < 16; ++j) atomicAdd(grad_w + NGP_MLP_W1 + j * 64 + m, v[j]);
        // dW5^T [64 in x 16 (3 used)] -> W5 is [3 out x 64 in]
        tmem_ld16(tmem_row + kColDW5T, v);
        if (has_row) chk();   // columns 3..15 hold products with the zero padding of dO: finite unless dO is not
        if (has_row)
#pragma unroll
            for (int j = 0; j < 3; ++j) atomicAdd(grad_w + (NGP_MLP_W1 + NGP_MLP_W2 + NGP_MLP_W3 + NGP_MLP_W4) + j * 64 + m, v[j]);
        if (bad && found_inf != nullptr) *found_inf = 1;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kTmemColsBwd);
}

template <typename TEmb, bool kSaved>
int launch_bwd(const void* emb, const float* dirs, const ngp_mlp_weights* w, const void* save, const float* dsigmas,
               const void* drgbs, void* demb, float* grad_w, int64_t n, const int32_t* n_dev, int32_t* found_inf,
               cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(mlp_bwd_kernel<TEmb, kSaved>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             kSmemBytesBwd);
        if (e != cudaSuccess) {
            ngp::set_error("mlp_bwd_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    const int64_t max_ctas = (int64_t)ngp::sm_count() * 2;  // 108 KB smem + 256 TMEM columns per CTA
    const unsigned grid = (unsigned)(n_tiles < max_ctas ? n_tiles : max_ctas);
    mlp_bwd_kernel<TEmb, kSaved><<<grid, kThreadsBwd, kSmemBytesBwd, st>>>((const TEmb*)emb, dirs, *w, (const __half*)save,
                                                                       dsigmas, (const __half*)drgbs, (TEmb*)demb,
                                                                       grad_w, n, n_dev, found_inf);
    NGP_LAUNCHED("mlp_bwd_kernel");
    return 0;
}

// forward implementation: 0 = auto (v2 where it applies), 1 = v1 (shared-memory activations), 2 = v2 (TMEM activations)
int g_fwd_impl = 0;

template <typename TEmb>
int launch_fwd(const void* emb, const float* dirs, const ngp_mlp_weights* w, float* sigmas, void* rgbs, void* save,
               int64_t n, const int32_t* n_dev, cudaStream_t st) {
    // the environment variable NGP_MLP_FWD overrides ngp_mlp_set_impl (A/B runs of unmodified scripts)
    static const int env_impl = [] { const char* e = getenv("NGP_MLP_FWD"); return e ? atoi(e) : -1; }();
    const int impl = env_impl >= 0 ? env_impl : g_fwd_impl;
    if (sizeof(TEmb) == 2 && impl != 1) {
        const int rc = ngp::mlp_fwd_v2_launch(emb, dirs, w, sigmas, rgbs, save, n, n_dev, st);
        if (rc != -2) return rc;   // -2: not applicable (tiny n / no tensor-map entry point) -> v1 below
        if (impl == 2 && n >= kTile) return -1;   // explicitly requested: do not fall back silently
    }
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(mlp_fwd_kernel<TEmb>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
        if (e != cudaSuccess) {
            ngp::set_error("mlp_fwd_kernel: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return (int)e;
        }
        configured = true;
    }
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    const int64_t max_ctas = (int64_t)ngp::sm_count() * 4;  // 52 KB smem + 64 TMEM columns per CTA
    const unsigned grid = (unsigned)(n_tiles < max_ctas ? n_tiles : max_ctas);
    mlp_fwd_kernel<TEmb><<<grid, kThreads, kSmemBytes, st>>>((const TEmb*)emb, dirs, *w, sigmas, (__half*)rgbs,
                                                            (__half*)save, n, n_dev);
    NGP_LAUNCHED("mlp_fwd_kernel");
    return 0;
}

}  // namespace

#ifdef NGP_MLP_TRACE
extern "C" int ngp_debug_mlp_trace(long long* out_host) {
    return (int)cudaMemcpyFromSymbol(out_host, g_mlp_trace, sizeof(long long) * 256);
}
#endif

extern "C" {

int ngp_mlp_set_impl(int fwd_impl) {
    NGP_REQUIRE(fwd_impl >= 0 && fwd_impl <= 2, "fwd_impl must be 0 (auto), 1 (v1) or 2 (v2)");
    g_fwd_impl = fwd_impl;
    return 0;
}

int64_t ngp_mlp_save_bytes(int64_t n) {
    // [n x 16] fp16 h (sigma-net output) + [n x 4] fp16 rgb (3 used): with them the backward skips layers 2 and 5
    return n > 0 ? n * 40 : 0;
}

static int check_mlp_args(const void* emb, int emb_dtype, const float* dirs, const ngp_mlp_weights* w) {
    NGP_REQUIRE(emb_dtype == NGP_F32 || emb_dtype == NGP_F16, "bad dtype");
    NGP_REQUIRE(emb && dirs && w, "null pointer");
    NGP_REQUIRE(w->w1 && w->w2 && w->w3 && w->w4 && w->w5, "null weight pointer");
    NGP_REQUIRE((reinterpret_cast<uintptr_t>(emb) & 15) == 0, "emb must be 16-byte aligned");
    const uintptr_t wal = reinterpret_cast<uintptr_t>(w->w1) | reinterpret_cast<uintptr_t>(w->w2) |
                          reinterpret_cast<uintptr_t>(w->w3) | reinterpret_cast<uintptr_t>(w->w4) |
                          reinterpret_cast<uintptr_t>(w->w5);
    NGP_REQUIRE((wal & 15) == 0, "weights must be 16-byte aligned");
    return 0;
}

int ngp_mlp_fwd(const void* emb, int emb_dtype, const float* dirs, const ngp_mlp_weights* w, float* sigmas,
                void* rgbs_f16, void* save, int64_t n, void* stream) {
    return ngp_mlp_fwd_dyn(emb, emb_dtype, dirs, w, sigmas, rgbs_f16, save, n, nullptr, stream);
}

int ngp_mlp_fwd_dyn(const void* emb, int emb_dtype, const float* dirs, const ngp_mlp_weights* w, float* sigmas,
                    void* rgbs_f16, void* save, int64_t n, const int32_t* n_dev, void* stream) {
    NGP_REQUIRE(n >= 0, "negative n");
    if (n == 0) return 0;
    if (int rc = check_mlp_args(emb, emb_dtype, dirs, w)) return rc;
    NGP_REQUIRE(sigmas && rgbs_f16, "null pointer");
    NGP_REQUIRE((reinterpret_cast<uintptr_t>(save) & 15) == 0, "save must be 16-byte aligned");
    cudaStream_t st = ngp::as_stream(stream);
    if (emb_dtype == NGP_F16) return launch_fwd<__half>(emb, dirs, w, sigmas, rgbs_f16, save, n, n_dev, st);
    return launch_fwd<float>(emb, dirs, w, sigmas, rgbs_f16, save, n, n_dev, st);
}

int ngp_mlp_bwd(const void* emb, int emb_dtype, const float* dirs, const ngp_mlp_weights* w, const void* save,
                const float* dsigmas, const void* drgbs_f16, void* demb, float* grad_w, int64_t n, void* stream) {
    return ngp_mlp_bwd_dyn(emb, emb_dtype, dirs, w, save, dsigmas, drgbs_f16, demb, grad_w, n, nullptr, nullptr, stream);
}

int ngp_mlp_bwd_dyn(const void* emb, int emb_dtype, const float* dirs, const ngp_mlp_weights* w, const void* save,
                    const float* dsigmas, const void* drgbs_f16, void* demb, float* grad_w, int64_t n,
                    const int32_t* n_dev, int32_t* found_inf_or_null, void* stream) {
    NGP_REQUIRE(n >= 0, "negative n");
    if (n == 0) return 0;
    if (int rc = check_mlp_args(emb, emb_dtype, dirs, w)) return rc;
    NGP_REQUIRE(dsigmas && drgbs_f16 && demb && grad_w, "null pointer");
    NGP_REQUIRE((reinterpret_cast<uintptr_t>(demb) & 15) == 0, "demb must be 16-byte aligned");
    NGP_REQUIRE((reinterpret_cast<uintptr_t>(save) & 15) == 0, "save must be 16-byte aligned");
    cudaStream_t st = ngp::as_stream(stream);
    if (emb_dtype == NGP_F16) {
        return save ? launch_bwd<__half, true>(emb, dirs, w, save, dsigmas, drgbs_f16, demb, grad_w, n, n_dev, found_inf_or_null, st)
                    : launch_bwd<__half, false>(emb, dirs, w, save, dsigmas, drgbs_f16, demb, grad_w, n, n_dev, found_inf_or_null, st);
    }
    return save ? launch_bwd<float, true>(emb, dirs, w, save, dsigmas, drgbs_f16, demb, grad_w, n, n_dev, found_inf_or_null, st)
                : launch_bwd<float, false>(emb, dirs, w, save, dsigmas, drgbs_f16, demb, grad_w, n, n_dev, found_inf_or_null, st);
}

}  // extern "C"

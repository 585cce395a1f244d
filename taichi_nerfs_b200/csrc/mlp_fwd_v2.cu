// mlp_fwd_v2.cu — warp-specialised forward of the fused NGP MLP: TMA-fed, activations resident in TENSOR MEMORY.
//
// Same math and rounding as mlp_fwd_kernel (mlp.cu) / the reference's five nn.Linear under autocast
// (modules/networks.py:136-166, 369-380, 18-30; SH: modules/spherical_harmonics.py:16-42) — what changes is where the
// bytes move.  v1 keeps every activation in shared memory: per 128-sample tile ~150 KB cross the 128 B/clk
// shared-memory port (STS of each layer's output, UMMA reads of every A operand and of the weights), which is what
// ncu shows as the top unit (l1tex 60 %) at 17 % tensor pipe.  Here
//   * the embedding tile [128 x 32] fp16 arrives by TMA (cp.async.bulk.tensor.2d, four 16-byte-wide column boxes ->
//     the UMMA no-swizzle K-major core-matrix layout, no register staging) into a 2-stage ring per tile slot;
//   * layer 1 is an SS MMA (A = that tile, B = W1 in shared memory); every later layer is a TS MMA whose A operand is
//     the previous layer's activation in TMEM: the epilogue reads the fp32 accumulators (tcgen05.ld), applies
//     ReLU / [SH | h] / TruncExp, packs fp16 pairs and writes them back with tcgen05.st — activations never touch
//     shared memory, only the 20 KB of weights are read from it (per tile ~28 KB instead of ~150 KB);
//   * roles: warp 8 issues the TMA loads (lane 0) and all MMAs (converged, one elected lane, descriptors from constant
//     memory — see the table below the constants); warps 0-3 / 4-7 are the epilogue warpgroups of tile slots
//     0 / 1 (one thread per sample row).  The issuer alternates between the slots, so slot B's MMA runs under slot A's
//     epilogue; synchronisation is per slot through mbarriers (acc[s]: tcgen05.commit -> epilogue, rdy[s]: 128
//     epilogue arrivals -> issuer), there is no CTA-wide barrier inside the tile loop.
// TMEM per slot (128 columns): D = fp32 accumulators [0,64), A = fp16-pair activations [64,96).  2 slots per CTA,
// 2 CTAs per SM = 512 columns = 4 tiles in flight per SM.
#include <cuda.h>

#include "tcgen05.cuh"

namespace {
using namespace ngp_tc;

constexpr int kSlots = 2;
constexpr int kStages = 2;
constexpr int kThreadsV2 = kSlots * 128 + 32;  // 288
constexpr int kIssuerWarp = kSlots * 4;        // warp 8
constexpr uint32_t kSlotCols = 128, kColA = 64;
constexpr uint32_t kTmemColsV2 = kSlots * kSlotCols;  // 256

constexpr int kEmbBytes = kTile * 32 * 2;                       // 8192
constexpr int kEmb = kAct;                                      // [slot][stage][4 k-chunks][128 rows][16 B]
constexpr int kBarV2 = kEmb + kSlots * kStages * kEmbBytes;     // 53,248
// mbarriers: full[slot][stage] (4), acc[slot] (2), rdy[slot] (2); then the TMEM base address
constexpr int kBarFull = kBarV2, kBarAcc = kBarV2 + 32, kBarRdy = kBarV2 + 48, kTmemSlot = kBarV2 + 64;
constexpr int kSmemV2 = kBarV2 + 80;

// The shared-memory descriptors of every MMA operand (weights: one per 16-wide K step; the layer-1 A operand: two per
// TMA stage buffer) are constants of the kernel.  They live in constant memory, filled once by the host from the
// kernel's shared-memory base address: the issuing warp — converged, one elected lane — fetches them straight into
// uniform registers (LDCU) instead of building them in ordinary registers and moving every 32-bit half with R2UR
// (~90 cycles per tcgen05.mma for a single active lane; see the backward, mlp.cu).
__constant__ uint64_t c_f2_w[16];                          // W1: 0-1, W2: 2-5, W3: 6-7, W4: 8-11, W5: 12-15
__constant__ uint64_t c_f2_a[kSlots * kStages][2];         // layer-1 A: [slot * kStages + stage][k step]
inline void build_f2_desc(uint32_t smem0, uint64_t* w, uint64_t (*a)[2]) {
    for (int k = 0; k < 2; ++k) {
        w[0 + k] = smem_desc(smem0 + kW1 + k * 256, 128, 32 * 16);
        w[6 + k] = smem_desc(smem0 + kW3 + k * 256, 128, 32 * 16);
    }
    for (int k = 0; k < 4; ++k) {
        w[2 + k] = smem_desc(smem0 + kW2 + k * 256, 128, 64 * 16);
        w[8 + k] = smem_desc(smem0 + kW4 + k * 256, 128, 64 * 16);
        w[12 + k] = smem_desc(smem0 + kW5 + k * 256, 128, 64 * 16);
    }
    for (int g = 0; g < kSlots * kStages; ++g)
        for (int k = 0; k < 2; ++k)   // TMA layout: k-chunk stride 2048 (LBO), 8-row group stride 128 (SBO)
            a[g][k] = smem_desc(smem0 + kEmb + g * kEmbBytes + k * 4096, 2048, 128);
}

// D[128 x 64] fp32 -> relu -> fp16 pairs -> A[128 x 64] (32 TMEM columns) of the same lane
__device__ __forceinline__ void epi_hidden_ts(uint32_t d_addr, uint32_t a_addr) {
    uint32_t v[64];
    tmem_ld32_issue_v(d_addr, v);
    tmem_ld32_issue_v(d_addr + 32, v + 32);
    tmem_ld_wait();
    uint32_t p[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) p[c] = pack_h2_relu(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1]));
    tmem_st32(a_addr, p);
}

__global__ void __launch_bounds__(kThreadsV2, 2)
mlp_fwd_v2_kernel(const __grid_constant__ CUtensorMap tmap_emb, const float* __restrict__ dirs, ngp_mlp_weights w,
                  float* __restrict__ sigmas, __half* __restrict__ rgbs, __half* __restrict__ save, int64_t n_max,
                  const int32_t* __restrict__ n_dev, uint32_t* __restrict__ probe_smem_base) {
    extern __shared__ __align__(128) uint8_t smem[];
    if (probe_smem_base != nullptr) {   // set-up launch: where this kernel's dynamic shared memory starts
        if (threadIdx.x == 0) *probe_smem_base = smem_u32(smem);
        return;
    }
    const int64_t n = n_dev ? min(n_max, max((int64_t)*n_dev, (int64_t)0)) : n_max;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    stage_weight(smem + kW1, w.w1, 64, 64, 32, kThreadsV2);
    stage_weight(smem + kW2, w.w2, 16, 16, 64, kThreadsV2);
    stage_weight(smem + kW3, w.w3, 64, 64, 32, kThreadsV2);
    stage_weight(smem + kW4, w.w4, 64, 64, 64, kThreadsV2);
    stage_weight(smem + kW5, w.w5, 3, 16, 64, kThreadsV2);
    if (tid == 0) {
        for (int b = 0; b < kSlots * kStages; ++b) mbar_init(smem_u32(smem + kBarFull + 8 * b), 1);
        for (int s = 0; s < kSlots; ++s) {
            mbar_init(smem_u32(smem + kBarAcc + 8 * s), 1);
            mbar_init(smem_u32(smem + kBarRdy + 8 * s), 128);
        }
        fence_barrier_init();
    }
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + kTmemSlot);
    if (warp == kIssuerWarp) tmem_alloc(smem_u32(tmem_slot), kTmemColsV2);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int64_t n_tiles = (n + kTile - 1) / kTile;
    const int64_t G = gridDim.x;
    // tile j of slot s (this CTA): blockIdx.x + (kSlots * j + s) * G

    if (warp == kIssuerWarp) {
        // ============ TMA producer + MMA issuer: the whole warp, converged; lane 0 issues the TMA loads, one
        // elected lane the MMAs (descriptors from constant memory) ============
        if (lane == 0) tma_prefetch_desc(&tmap_emb);
        int64_t cnt[kSlots];
#pragma unroll
        for (int s = 0; s < kSlots; ++s) {
            const int64_t first = (int64_t)blockIdx.x + s * G;
            cnt[s] = first < n_tiles ? (n_tiles - 1 - first) / (kSlots * G) + 1 : 0;
        }
        auto issue_tma = [&](int s, int64_t j) {   // lane 0 only
            const int64_t tile = (int64_t)blockIdx.x + (kSlots * j + s) * G;
            const uint32_t bar = smem_u32(smem + kBarFull + 8 * (s * kStages + (int)(j & 1)));
            const uint32_t dst = smem_u32(smem + kEmb + (s * kStages + (int)(j & 1)) * kEmbBytes);
            mbar_arrive_expect_tx(bar, kEmbBytes);
#pragma unroll
            for (int kc = 0; kc < 4; ++kc)   // 8 fp16 columns x 128 rows -> one 2 KB column of core matrices
                tma_load_2d(dst + kc * 2048, &tmap_emb, kc * 8, (int32_t)(tile * kTile), bar);
        };
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < kSlots; ++s)
                for (int64_t j = 0; j < kStages && j < cnt[s]; ++j) issue_tma(s, j);
        }
        __syncwarp();
        constexpr uint32_t id64 = idesc_f16(kTile, 64), id16 = idesc_f16(kTile, 16);

        uint32_t rdy_phase[kSlots] = {0, 0};
        const int64_t jmax = cnt[0] > cnt[1] ? cnt[0] : cnt[1];
        for (int64_t j = 0; j < jmax; ++j) {
            const bool odd = (j & 1) != 0;   // which TMA stage this tile's embedding is in (warp-uniform)
#pragma unroll
            for (int l = 0; l < 5; ++l) {
#pragma unroll
                for (int s = 0; s < kSlots; ++s) {
                    if (j >= cnt[s]) continue;
                    const uint32_t D = tmem_base + s * kSlotCols, A = D + kColA;
                    const uint32_t acc = smem_u32(smem + kBarAcc + 8 * s);
                    // l == 0: the slot is free (epilogue of its previous tile has read its outputs);
                    // l >= 1: the epilogue has written layer l's A operand into TMEM
                    mbar_wait_bounded(smem_u32(smem + kBarRdy + 8 * s), rdy_phase[s]);
                    rdy_phase[s] ^= 1;
                    if (l == 0) {
                        const int stg = s * kStages + (odd ? 1 : 0);
                        mbar_wait_bounded(smem_u32(smem + kBarFull + 8 * stg), (uint32_t)((j >> 1) & 1));
                        tc_fence_after();
                        if (odd) {
#pragma unroll
                            for (int k = 0; k < 2; ++k) umma_f16_w(D, c_f2_a[s * kStages + 1][k], c_f2_w[0 + k], id64, k > 0);
                        } else {
#pragma unroll
                            for (int k = 0; k < 2; ++k) umma_f16_w(D, c_f2_a[s * kStages + 0][k], c_f2_w[0 + k], id64, k > 0);
                        }
                    } else {
                        tc_fence_after();
                        if (l == 1) {
                            // layer 1 of tile j has completed (its epilogue ran): the stage is free for tile j + 2
                            if (lane == 0 && j + kStages < cnt[s]) issue_tma(s, j + kStages);
                            __syncwarp();
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_f16_ts_w(D, A + k * 8, c_f2_w[2 + k], id16, k > 0);
                        } else if (l == 2) {
#pragma unroll
                            for (int k = 0; k < 2; ++k) umma_f16_ts_w(D, A + k * 8, c_f2_w[6 + k], id64, k > 0);
                        } else if (l == 3) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_f16_ts_w(D, A + k * 8, c_f2_w[8 + k], id64, k > 0);
                        } else {
#pragma unroll
                            for (int k = 0; k < 4; ++k) umma_f16_ts_w(D, A + k * 8, c_f2_w[12 + k], id16, k > 0);
                        }
                    }
                    umma_commit_w(acc);
                }
            }
        }
        __syncwarp();
    } else {
        // ======================= epilogue warpgroup of slot s: one thread per sample row =======================
        const int s = warp >> 2, row = tid & 127;
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;   // a warp may touch TMEM lanes 32*(warp%4)..+31
        const uint32_t D = tmem_base + s * kSlotCols + lane_base, A = D + kColA;
        const uint32_t acc = smem_u32(smem + kBarAcc + 8 * s), rdy = smem_u32(smem + kBarRdy + 8 * s);
        uint32_t acc_phase = 0;
        mbar_arrive(rdy);   // slot free
        for (int64_t j = 0;; ++j) {
            const int64_t tile = (int64_t)blockIdx.x + (kSlots * j + s) * G;
            if (tile >= n_tiles) break;
            const int64_t i = tile * kTile + row;
            const bool valid = i < n;
            float dx = 0.f, dy = 0.f, dz = 1.f;
            if (valid) {
                dx = __ldg(dirs + i * 3 + 0);
                dy = __ldg(dirs + i * 3 + 1);
                dz = __ldg(dirs + i * 3 + 2);
            }
            // ---- layer 1: H1 = relu(X W1^T) -> A
            mbar_wait_bounded(acc, acc_phase);
            acc_phase ^= 1;
            tc_fence_after();
            epi_hidden_ts(D, A);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(rdy);
            // the direction half of X3 while layer 2 runs: SH16((d/|d| + 1)/2)  (networks.py:162-164)
            uint32_t x3[16];
            {
                const float inv = 1.0f / sqrtf(dx * dx + dy * dy + dz * dz);
                float e[16];
                sh16((dx * inv + 1.0f) / 2.0f, (dy * inv + 1.0f) / 2.0f, (dz * inv + 1.0f) / 2.0f, e);
#pragma unroll
                for (int c = 0; c < 8; ++c) x3[c] = pack_h2(e[2 * c], e[2 * c + 1]);
            }
            // ---- layer 2: h = H1 W2^T; sigma = TruncExp(h[:,0]) (networks.py:22-24, :146); X3 = [SH | h] -> A
            mbar_wait_bounded(acc, acc_phase);
            acc_phase ^= 1;
            tc_fence_after();
            float h[16];
            tmem_ld16(D, h);
#pragma unroll
            for (int c = 0; c < 8; ++c) x3[8 + c] = pack_h2(h[2 * c], h[2 * c + 1]);
            tmem_st16(A, x3);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(rdy);
            if (valid) {
                const float h0 = __half2float(__float2half_rn(h[0]));
                sigmas[i] = expf(h0);
                if (save != nullptr) {   // the backward restarts from h instead of recomputing layers 1-2 serially
                    reinterpret_cast<uint4*>(save + i * 16)[0] = make_uint4(x3[8], x3[9], x3[10], x3[11]);
                    reinterpret_cast<uint4*>(save + i * 16)[1] = make_uint4(x3[12], x3[13], x3[14], x3[15]);
                }
            }
            // ---- layer 3: H3 = relu(X3 W3^T), layer 4: H4 = relu(H3 W4^T)
#pragma unroll
            for (int l = 0; l < 2; ++l) {
                mbar_wait_bounded(acc, acc_phase);
                acc_phase ^= 1;
                tc_fence_after();
                epi_hidden_ts(D, A);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(rdy);
            }
            // ---- layer 5: rgb = sigmoid(H4 W5^T) (3 of 16 columns used)
            mbar_wait_bounded(acc, acc_phase);
            acc_phase ^= 1;
            tc_fence_after();
            uint32_t o[4];
            tmem_ld4_issue_v(D, o);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(rdy);   // slot free: the next tile's layer 1 may overwrite D
            if (valid) {
                __half out[3];
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    const float oc = __half2float(__float2half_rn(__uint_as_float(o[c])));
                    out[c] = __float2half_rn(1.0f / (1.0f + expf(-oc)));
                    rgbs[i * 3 + c] = out[c];
                }
                if (save != nullptr) {
                    const __half2 a = __halves2half2(out[0], out[1]), b = __halves2half2(out[2], __float2half_rn(0.0f));
                    *reinterpret_cast<uint2*>(save + n_max * 16 + i * 4) =
                        make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == kIssuerWarp) tmem_dealloc(tmem_base, kTmemColsV2);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            (void)cudaGetLastError();
    }
    return fn;
}

}  // namespace

namespace ngp {

// returns 0 on launch, > 0 = cudaError, -2 = "not applicable here, use the v1 kernel" (no error is recorded)
int mlp_fwd_v2_launch(const void* emb_f16, const float* dirs, const ngp_mlp_weights* w, float* sigmas, void* rgbs,
                      void* save, int64_t n, const int32_t* n_dev, cudaStream_t st) {
    if (n < kTile || n >= (int64_t)1 << 31) return -2;
    EncodeTiledFn enc = encode_fn();
    if (enc == nullptr) return -2;
    // the embedding as a 2-D tensor [n rows x 32 fp16]; one box = 8 columns (16 B) x 128 rows, which lands as 128
    // consecutive 16-byte rows = 16 stacked 8x8 core matrices of the UMMA no-swizzle layout
    CUtensorMap tmap;
    const cuuint64_t dims[2] = {32, (cuuint64_t)n};
    const cuuint64_t strides[1] = {64};
    const cuuint32_t box[2] = {8, (cuuint32_t)kTile};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(emb_f16), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("mlp_fwd_v2: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return -2;
    }
    // (the descriptor table lives in constant memory, i.e. per device: one flag per device of this process)
    static bool configured_dev[64] = {};
    int dev_id = 0;
    if (cudaGetDevice(&dev_id) != cudaSuccess || dev_id < 0 || dev_id >= 64) dev_id = 0;
    if (!configured_dev[dev_id]) {
        // the set-up synchronises the stream: under stream capture this launch goes to the v1 kernel (-2) and the set-up
        // waits for the first eager call (every warm-up makes one)
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) {
            cudaGetLastError();
            return -2;
        }
        cudaError_t e = cudaFuncSetAttribute(mlp_fwd_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemV2);
        uint32_t* d_base = nullptr;
        uint32_t h_base = 0;
        if (e == cudaSuccess) e = cudaMalloc(&d_base, sizeof(uint32_t));
        if (e == cudaSuccess) {
            mlp_fwd_v2_kernel<<<1, kThreadsV2, kSmemV2, st>>>(tmap, nullptr, *w, nullptr, nullptr, nullptr, 0, nullptr, d_base);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(&h_base, d_base, sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) {
            static uint64_t hw[16], ha[kSlots * kStages][2];
            build_f2_desc(h_base, hw, ha);
            e = cudaMemcpyToSymbolAsync(c_f2_w, hw, sizeof(hw), 0, cudaMemcpyHostToDevice, st);
            if (e == cudaSuccess) e = cudaMemcpyToSymbolAsync(c_f2_a, ha, sizeof(ha), 0, cudaMemcpyHostToDevice, st);
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (d_base) cudaFree(d_base);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("mlp_fwd_v2_kernel: set-up: %s", cudaGetErrorString(e));
            return (int)e;
        }
        configured_dev[dev_id] = true;
    }
    const int64_t n_tiles = (n + kTile - 1) / kTile;
    const int64_t want = (n_tiles + kSlots - 1) / kSlots;
    const int64_t max_ctas = (int64_t)sm_count() * 2;   // 256 TMEM columns + 53 KB shared memory per CTA
    const unsigned grid = (unsigned)(want < max_ctas ? want : max_ctas);
    mlp_fwd_v2_kernel<<<grid, kThreadsV2, kSmemV2, st>>>(tmap, dirs, *w, sigmas, (__half*)rgbs, (__half*)save, n, n_dev,
                                                         nullptr);
    count_launch();
    return check_launch("mlp_fwd_v2_kernel");
}

}  // namespace ngp

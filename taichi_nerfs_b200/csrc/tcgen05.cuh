// tcgen05.cuh — PTX wrappers for the 5th-generation tensor cores (tcgen05.mma / TMEM / mbarrier / TMA bulk
// copies), UMMA descriptors for the no-swizzle layout and the few device helpers shared by the fused MLP
// kernels (mlp.cu: forward v1 + backward, mlp_fwd_v2.cu: warp-specialised forward).  sm_100a only.
#pragma once
#include "common.cuh"

namespace ngp_tc {

constexpr int kTile = 128;          // samples per tile == UMMA M == TMEM lanes

// weights in shared memory (fp16, UMMA no-swizzle K-major layout), identical in every MLP kernel
constexpr int kW1 = 0;                       // [64 x 32]
constexpr int kW2 = kW1 + 64 * 32 * 2;       // [16 x 64]
constexpr int kW3 = kW2 + 16 * 64 * 2;       // [64 x 32]
constexpr int kW4 = kW3 + 64 * 32 * 2;       // [64 x 64]
constexpr int kW5 = kW4 + 64 * 64 * 2;       // [16 x 64] (rows 3..15 zero)
constexpr int kAct = kW5 + 16 * 64 * 2;      // 20480: first byte after the weights

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16, issued by ONE thread
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                         uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// The same two instructions for a CONVERGED warp: every lane executes the surrounding (uniform) descriptor arithmetic,
// one elected lane issues.  With a single active lane (`if (lane == 0)`) the compiler must assume divergent values and
// moves every descriptor into uniform registers one 32-bit half at a time (R2UR): ~90 cycles per tcgen05.mma measured
// (profiles/r2_mlp_bwd_v2_trace_single_thread_issue.txt).  Converged, the descriptors are computed on the uniform
// datapath directly.  elect.sync picks the same lane for the same member mask, so the commit tracks these MMAs.
__device__ __forceinline__ void umma_f16_w(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_w(uint32_t bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
        ::"r"(bar)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 16 consecutive fp32 columns -> 16 registers per thread (thread i of warp w = lane 32w+i)
// issue only; the registers are valid after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t r[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
    uint32_t r[16];
    tmem_ld16_issue(taddr, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// 32 accumulator columns (two pipelined loads, one wait)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float v[32]) {
    uint32_t r[32];
    tmem_ld16_issue(taddr, r);
    tmem_ld16_issue(taddr + 16, r + 16);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// ---- descriptors -------------------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (sm_100)
__host__ __device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor: D=f32, A=B=f16, both K-major, M=128, N
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// byte offset of the 16-byte chunk (row r, k-chunk kc) inside an operand with K columns
__device__ __forceinline__ int chunk_off(int r, int kc, int K) { return (r >> 3) * (K * 16) + kc * 128 + (r & 7) * 16; }

// one GEMM layer: D[128 x N] = A[128 x K] * W[N x K]^T   (K in {16,32,64})
__device__ __forceinline__ void issue_layer_mma(uint32_t tmem_d, uint32_t a_addr, uint32_t w_addr, int K, int N) {
    const uint32_t idesc = idesc_f16(kTile, N);
    const uint32_t sbo = (uint32_t)K * 16;  // (K/8)*128
    for (int k = 0; k < K / 16; ++k) {
        const uint64_t da = smem_desc(a_addr + k * 256, 128, sbo);
        const uint64_t db = smem_desc(w_addr + k * 256, 128, sbo);
        umma_f16(tmem_d, da, db, idesc, k > 0 ? 1u : 0u);
    }
}
__device__ __forceinline__ void issue_layer(uint32_t tmem_d, uint32_t a_addr, uint32_t w_addr, int K, int N,
                                            uint32_t bar) {
    issue_layer_mma(tmem_d, a_addr, w_addr, K, N);
    umma_commit(bar);
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// relu fused into the conversion (one F2FP instead of two FMNMX + F2FP); low half = a
__device__ __forceinline__ uint32_t pack_h2_relu(float a, float b) {
    uint32_t r;
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}

// copy a row-major fp32 weight [rows x K] into the interleaved fp16 operand layout (rows_pad rows)
__device__ __forceinline__ void stage_weight(uint8_t* smem, const float* __restrict__ w, int rows, int rows_pad, int K,
                                             int n_threads) {
    const int kchunks = K / 8;
    if ((int)threadIdx.x >= n_threads) return;  // e.g. the backward's MMA-issue warp does not stage
    for (int c = threadIdx.x; c < rows_pad * kchunks; c += n_threads) {
        const int r = c / kchunks, kc = c % kchunks;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < rows) {
            const float4 a = __ldg(reinterpret_cast<const float4*>(w + r * K + kc * 8));
            const float4 b = __ldg(reinterpret_cast<const float4*>(w + r * K + kc * 8 + 4));
            v = make_uint4(pack_h2(a.x, a.y), pack_h2(a.z, a.w), pack_h2(b.x, b.y), pack_h2(b.z, b.w));
        }
        *reinterpret_cast<uint4*>(smem + chunk_off(r, kc, K)) = v;
    }
}

__device__ __forceinline__ void sh16(float x, float y, float z, float* e) {  // spherical_harmonics.py:16-42
    const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
    e[0] = 0.28209479177387814f;
    e[1] = -0.48860251190291987f * y;
    e[2] = 0.48860251190291987f * z;
    e[3] = -0.48860251190291987f * x;
    e[4] = 1.0925484305920792f * xy;
    e[5] = -1.0925484305920792f * yz;
    e[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
    e[7] = -1.0925484305920792f * xz;
    e[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
    e[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
    e[10] = 2.8906114426405538f * xy * z;
    e[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
    e[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
    e[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
    e[14] = 1.4453057213202769f * z * (x2 - y2);
    e[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

// hidden-layer epilogue: TMEM [128 x 64] fp32 -> relu -> fp16 -> next operand buffer (K = 64 layout);
// thread (row, hh) converts columns [32*hh, 32*hh+32) = chunks 4*hh .. 4*hh+3
__device__ __forceinline__ void epilogue_hidden(uint32_t tmem_row, uint8_t* dst, int row, int hh) {
    float v[32];
    tmem_ld32(tmem_row + hh * 32, v);
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) {
        const float* q = v + kc * 8;
        *reinterpret_cast<uint4*>(dst + chunk_off(row, hh * 4 + kc, 64)) =
            make_uint4(pack_h2_relu(q[0], q[1]), pack_h2_relu(q[2], q[3]), pack_h2_relu(q[4], q[5]),
                       pack_h2_relu(q[6], q[7]));
    }
}

// ---- additions for the warp-specialised kernels (mlp_fwd_v2.cu) -----------------------------------------
// mbarrier wait that traps instead of hanging the GPU if a producer never arrives (a bug, not a run-time state)
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    uint32_t done, spins = 0;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && ++spins > (1u << 20)) __trap();
    } while (!done);
}
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// TMA: 2-D tiled bulk tensor copy global -> shared, completion on an mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_2d(uint32_t dst_smem, const void* tmap, int32_t c0, int32_t c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst_smem), "l"(tmap), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// TMA: 1-D bulk copy global -> shared (SASS: UBLKCP); bytes and both addresses multiples of 16
__device__ __forceinline__ void bulk_load_1d(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T: the A operand (fp16 pairs packed along K, lane = row) comes from tensor memory
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// converged-warp variant (see umma_f16_w)
__device__ __forceinline__ void umma_f16_ts_w(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld4_issue_v(uint32_t taddr, uint32_t r[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32_issue_v(uint32_t taddr, uint32_t r[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t r[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
                 : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t r[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
                 : "memory");
}

}  // namespace ngp_tc

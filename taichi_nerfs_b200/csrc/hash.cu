// hash.cu — multiresolution hash-grid encoding: forward, table backward, input backward.
//
// Semantics: modules/hash_encoder.py:43-143 (fp32) and modules/hash_encoder_half.py:112-213
// (fp16 table, fp16 accumulate) of the reference.  Per-level scale/resolution come from the
// host-built ngp_hash_layout (see include/ngp_b200.h) instead of a per-thread expf.
//
// B200 mapping.  The reference launches one thread per (sample, level) with the level as the
// fastest index and block_dim=16, so a warp touches 2 samples x 16 unrelated table regions.
// Here a CTA owns a tile of 128 consecutive samples; each warp = 32 consecutive samples at ONE
// level, so the 8 corner gathers of neighbouring samples (consecutive samples of a ray are
// <= 1 finest-level cell apart) fall into the same L1/L2 sectors.  The fp16 table (21.8 MiB)
// and fp32 gradient (43.6 MiB) both fit B200's 126 MB L2, so these kernels are L2-gather /
// L2-atomic bound, not HBM bound.  The [tile, L*F] result is staged in shared memory and
// written with 16-byte coalesced stores (the reference's output layout, [n, L*F] row-major).
#include "common.cuh"

namespace {

constexpr int kTile = 128;      // samples per CTA
constexpr int kGroups = 4;      // level groups per CTA (warp-uniform)
constexpr int kThreads = kTile * kGroups;

struct LevelMeta {
    uint32_t offset;     // entries
    uint32_t size;       // entries
    uint32_t mask;       // size-1 if power of two else 0
    uint32_t res;
    float scale;
    int dense;
};

__device__ __forceinline__ LevelMeta level_meta(const ngp_hash_layout& lay, int l) {
    LevelMeta m;
    m.offset = (uint32_t)lay.offsets[l];
    m.size = (uint32_t)lay.map_sizes[l];
    m.mask = (m.size & (m.size - 1)) == 0 ? m.size - 1 : 0u;
    m.res = lay.resolutions[l];
    m.scale = lay.scales[l];
    m.dense = l < lay.begin_fast_hash_level;
    return m;
}

// grid coordinate + fractional position of x at this level (hash_encoder.py:108-110)
template <bool kFracF16>
__device__ __forceinline__ void grid_pos(const float x[3], const LevelMeta& m, uint32_t g[3], float pos[3]) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float p = f_add(f_mul(x[d], m.scale), 0.5f);
        const float fl = floorf(p);
        g[d] = (uint32_t)(int32_t)fl;
        // half kernel: pos -= cast(pos_grid, f16) (hash_encoder_half.py:132) — exact below 2048
        const float gf = kFracF16 ? __half2float(__float2half_rn((float)g[d])) : (float)g[d];
        pos[d] = f_sub(p, gf);
    }
}

__device__ __forceinline__ uint32_t corner_index(const LevelMeta& m, const uint32_t g[3], int c) {
    const uint32_t px = g[0] + (c & 1), py = g[1] + ((c >> 1) & 1), pz = g[2] + ((c >> 2) & 1);
    uint32_t h;
    if (m.dense) h = px + py * m.res + pz * (m.res * m.res);               // under_hash, hash_encoder.py:53-60
    else h = px ^ (py * 2654435761u) ^ (pz * 805459861u);                  // fast_hash,  hash_encoder.py:43-51
    h = m.mask ? (h & m.mask) : (h % m.size);                              // hash_encoder.py:71
    return m.offset + h;
}

__device__ __forceinline__ float corner_weight(const float pos[3], int c) {  // hash_encoder.py:116-126
    float w = 1.0f;
#pragma unroll
    for (int d = 0; d < 3; ++d) w = f_mul(w, (c & (1 << d)) ? pos[d] : f_sub(1.0f, pos[d]));
    return w;
}

// ---- forward -------------------------------------------------------------------------------
template <typename T>
struct Vec2;
template <>
struct Vec2<float> { using type = float2; };
template <>
struct Vec2<__half> { using type = __half2; };

template <typename T>
__global__ void __launch_bounds__(kThreads) hash_fwd_kernel(const float* __restrict__ xyz,
                                                            const T* __restrict__ table,
                                                            const __grid_constant__ ngp_hash_layout lay,
                                                            T* __restrict__ out, int64_t n) {
    using V2 = typename Vec2<T>::type;
    constexpr bool kHalf = sizeof(T) == 2;
    __shared__ __align__(16) V2 tile[kTile][NGP_MAX_LEVELS + 1];  // +1: bank spread for the transpose

    const int s = threadIdx.x & (kTile - 1);
    const int grp = threadIdx.x >> 7;
    const int64_t i = (int64_t)blockIdx.x * kTile + s;
    const int L = lay.n_levels;

    if (i < n) {
        const float x[3] = {xyz[i * 3 + 0], xyz[i * 3 + 1], xyz[i * 3 + 2]};
        for (int l = grp; l < L; l += kGroups) {
            const LevelMeta m = level_meta(lay, l);
            uint32_t g[3];
            float pos[3];
            grid_pos<kHalf>(x, m, g, pos);
            V2 v[8];
#pragma unroll
            for (int c = 0; c < 8; ++c) v[c] = __ldg(reinterpret_cast<const V2*>(table) + corner_index(m, g, c));
            if constexpr (kHalf) {
                __half2 acc = __float2half2_rn(0.0f);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float w = corner_weight(pos, c);
                    const float2 t = __half22float2(v[c]);
                    // local += cast(w * table[idx], f16); f16 accumulate (hash_encoder_half.py:159)
                    acc = __hadd2(acc, __floats2half2_rn(f_mul(w, t.x), f_mul(w, t.y)));
                }
                tile[s][l] = acc;
            } else {
                float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const float w = corner_weight(pos, c);
                    acc.x = f_add(acc.x, f_mul(w, v[c].x));
                    acc.y = f_add(acc.y, f_mul(w, v[c].y));
                }
                tile[s][l] = acc;
            }
        }
    }
    __syncthreads();
    // coalesced write-out: row i holds L V2 values
    const int64_t base = (int64_t)blockIdx.x * kTile;
    const int rows = (int)min((int64_t)kTile, n - base);
    V2* o2 = reinterpret_cast<V2*>(out) + base * L;
    for (int k = threadIdx.x; k < rows * L; k += kThreads) o2[k] = tile[k / L][k % L];
}

// ---- backward wrt table ----------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) hash_bwd_kernel(const float* __restrict__ xyz,
                                                            const T* __restrict__ dout,
                                                            const __grid_constant__ ngp_hash_layout lay,
                                                            float* __restrict__ grad_table, int64_t n) {
    using V2 = typename Vec2<T>::type;
    constexpr bool kHalf = sizeof(T) == 2;
    __shared__ __align__(16) V2 tile[kTile][NGP_MAX_LEVELS + 1];

    const int L = lay.n_levels;
    const int64_t base = (int64_t)blockIdx.x * kTile;
    const int rows = (int)min((int64_t)kTile, n - base);
    const V2* d2 = reinterpret_cast<const V2*>(dout) + base * L;
    for (int k = threadIdx.x; k < rows * L; k += kThreads) tile[k / L][k % L] = d2[k];
    __syncthreads();

    const int s = threadIdx.x & (kTile - 1);
    const int grp = threadIdx.x >> 7;
    const int64_t i = base + s;
    if (i >= n) return;
    const float x[3] = {xyz[i * 3 + 0], xyz[i * 3 + 1], xyz[i * 3 + 2]};
    float2* g2 = reinterpret_cast<float2*>(grad_table);
    for (int l = grp; l < L; l += kGroups) {
        float2 dy;
        if constexpr (kHalf) dy = __half22float2(tile[s][l]);
        else dy = tile[s][l];
        if (dy.x == 0.0f && dy.y == 0.0f) continue;  // hash_encoder_half.py:210
        const LevelMeta m = level_meta(lay, l);
        uint32_t g[3];
        float pos[3];
        grid_pos<kHalf>(x, m, g, pos);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const float w = corner_weight(pos, c);
            atomicAdd(g2 + corner_index(m, g, c), make_float2(w * dy.x, w * dy.y));  // red.global.add.v2.f32
        }
    }
}

// ---- backward wrt input position ----------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) hash_bwd_input_kernel(const float* __restrict__ xyz,
                                                             const T* __restrict__ table,
                                                             const T* __restrict__ dout,
                                                             const __grid_constant__ ngp_hash_layout lay,
                                                             float* __restrict__ dx, int64_t n) {
    using V2 = typename Vec2<T>::type;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int L = lay.n_levels;
    const float x[3] = {xyz[i * 3 + 0], xyz[i * 3 + 1], xyz[i * 3 + 2]};
    float gx = 0.f, gy = 0.f, gz = 0.f;
    for (int l = 0; l < L; ++l) {
        const LevelMeta m = level_meta(lay, l);
        uint32_t g[3];
        float pos[3];
        grid_pos<false>(x, m, g, pos);
        const V2 dyv = reinterpret_cast<const V2*>(dout)[i * L + l];
        float2 dy;
        if constexpr (sizeof(T) == 2) dy = __half22float2(dyv);
        else dy = dyv;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const V2 tv = __ldg(reinterpret_cast<const V2*>(table) + corner_index(m, g, c));
            float2 t;
            if constexpr (sizeof(T) == 2) t = __half22float2(tv);
            else t = tv;
            const float v = (t.x * dy.x + t.y * dy.y) * m.scale;
            const float wx = (c & 1) ? pos[0] : 1.0f - pos[0];
            const float wy = (c & 2) ? pos[1] : 1.0f - pos[1];
            const float wz = (c & 4) ? pos[2] : 1.0f - pos[2];
            gx += ((c & 1) ? v : -v) * wy * wz;
            gy += ((c & 2) ? v : -v) * wx * wz;
            gz += ((c & 4) ? v : -v) * wx * wy;
        }
    }
    dx[i * 3 + 0] = gx;
    dx[i * 3 + 1] = gy;
    dx[i * 3 + 2] = gz;
}

int check_layout(const ngp_hash_layout* lay) {
    NGP_REQUIRE(lay != nullptr, "null layout");
    NGP_REQUIRE(lay->n_levels >= 1 && lay->n_levels <= NGP_MAX_LEVELS, "n_levels out of range");
    NGP_REQUIRE(lay->feat_dim == 2, "CUDA path supports feature_per_level == 2 only");
    return 0;
}

}  // namespace

extern "C" {

int ngp_hash_encode_fwd(const float* xyz, const void* table, const ngp_hash_layout* layout, void* out,
                        int dtype, int64_t n, void* stream) {
    if (int rc = check_layout(layout)) return rc;
    NGP_REQUIRE(n >= 0, "negative n");
    NGP_REQUIRE(dtype == NGP_F32 || dtype == NGP_F16, "bad dtype");
    if (n == 0) return 0;
    NGP_REQUIRE(xyz && table && out, "null pointer");
    const unsigned grid = (unsigned)((n + kTile - 1) / kTile);
    cudaStream_t st = ngp::as_stream(stream);
    if (dtype == NGP_F16)
        hash_fwd_kernel<__half><<<grid, kThreads, 0, st>>>(xyz, (const __half*)table, *layout, (__half*)out, n);
    else
        hash_fwd_kernel<float><<<grid, kThreads, 0, st>>>(xyz, (const float*)table, *layout, (float*)out, n);
    NGP_LAUNCHED("hash_fwd_kernel");
    return 0;
}

int ngp_hash_encode_bwd(const float* xyz, const void* dout, int dout_dtype, const ngp_hash_layout* layout,
                        float* grad_table, int64_t n, void* stream) {
    if (int rc = check_layout(layout)) return rc;
    NGP_REQUIRE(n >= 0, "negative n");
    NGP_REQUIRE(dout_dtype == NGP_F32 || dout_dtype == NGP_F16, "bad dtype");
    if (n == 0) return 0;
    NGP_REQUIRE(xyz && dout && grad_table, "null pointer");
    NGP_REQUIRE((reinterpret_cast<uintptr_t>(grad_table) & 7) == 0, "grad_table must be 8-byte aligned");
    const unsigned grid = (unsigned)((n + kTile - 1) / kTile);
    cudaStream_t st = ngp::as_stream(stream);
    if (dout_dtype == NGP_F16)
        hash_bwd_kernel<__half><<<grid, kThreads, 0, st>>>(xyz, (const __half*)dout, *layout, grad_table, n);
    else
        hash_bwd_kernel<float><<<grid, kThreads, 0, st>>>(xyz, (const float*)dout, *layout, grad_table, n);
    NGP_LAUNCHED("hash_bwd_kernel");
    return 0;
}

int ngp_hash_encode_bwd_input(const float* xyz, const void* table, const void* dout, int dtype,
                              const ngp_hash_layout* layout, float* dx, int64_t n, void* stream) {
    if (int rc = check_layout(layout)) return rc;
    NGP_REQUIRE(n >= 0, "negative n");
    NGP_REQUIRE(dtype == NGP_F32 || dtype == NGP_F16, "bad dtype");
    if (n == 0) return 0;
    NGP_REQUIRE(xyz && table && dout && dx, "null pointer");
    const unsigned grid = (unsigned)((n + 255) / 256);
    cudaStream_t st = ngp::as_stream(stream);
    if (dtype == NGP_F16)
        hash_bwd_input_kernel<__half><<<grid, 256, 0, st>>>(xyz, (const __half*)table, (const __half*)dout, *layout, dx, n);
    else
        hash_bwd_input_kernel<float><<<grid, 256, 0, st>>>(xyz, (const float*)table, (const float*)dout, *layout, dx, n);
    NGP_LAUNCHED("hash_bwd_input_kernel");
    return 0;
}

}  // extern "C"

// grid.cu — occupancy-grid helpers: packbits and Morton encode/decode.
// Semantics: modules/utils.py:95-169 of the reference (no ti.sync() host syncs here).
#include "common.cuh"

namespace {

__device__ __forceinline__ uint32_t expand_bits(uint32_t v) {  // utils.py:95-100
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ int32_t compact_bits(uint32_t x) {  // utils.py:110-117
    x = x & 0x49249249u;
    x = (x | (x >> 2)) & 0xc30c30c3u;
    x = (x | (x >> 4)) & 0x0f00f00fu;
    x = (x | (x >> 8)) & 0xff0000ffu;
    x = (x | (x >> 16)) & 0x0000ffffu;
    return (int32_t)x;
}

// one thread per output byte; the 8 floats it tests are two aligned float4 loads
__global__ void __launch_bounds__(256) packbits_kernel(const float* __restrict__ grid, float thr,
                                                       const float* __restrict__ mean_dev,
                                                       uint8_t* __restrict__ bits, int64_t n_bytes) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= n_bytes) return;
    if (mean_dev != nullptr) {  // python: min(mean_density, density_threshold) -> NaN mean stays NaN
        const float m = *mean_dev;
        thr = (thr < m) ? thr : m;
    }
    const float4 a = reinterpret_cast<const float4*>(grid)[2 * n];
    const float4 b = reinterpret_cast<const float4*>(grid)[2 * n + 1];
    uint32_t v = 0;
    v |= (a.x > thr) ? 1u : 0u;
    v |= (a.y > thr) ? 2u : 0u;
    v |= (a.z > thr) ? 4u : 0u;
    v |= (a.w > thr) ? 8u : 0u;
    v |= (b.x > thr) ? 16u : 0u;
    v |= (b.y > thr) ? 32u : 0u;
    v |= (b.z > thr) ? 64u : 0u;
    v |= (b.w > thr) ? 128u : 0u;
    bits[n] = (uint8_t)v;
}

__global__ void __launch_bounds__(256) morton3d_kernel(const int32_t* __restrict__ coords,
                                                       int32_t* __restrict__ indices, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t x = (uint32_t)coords[i * 3 + 0], y = (uint32_t)coords[i * 3 + 1], z = (uint32_t)coords[i * 3 + 2];
    indices[i] = (int32_t)(expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2));
}

__global__ void __launch_bounds__(256) morton3d_invert_kernel(const int32_t* __restrict__ indices,
                                                              int32_t* __restrict__ coords, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t ind = (uint32_t)indices[i];
    coords[i * 3 + 0] = compact_bits(ind >> 0);
    coords[i * 3 + 1] = compact_bits(ind >> 1);
    coords[i * 3 + 2] = compact_bits(ind >> 2);
}

}  // namespace

extern "C" {

int ngp_packbits(const float* density_grid, float density_threshold, uint8_t* density_bitfield, int64_t n_bytes,
                 void* stream) {
    NGP_REQUIRE(n_bytes >= 0, "negative n_bytes");
    if (n_bytes == 0) return 0;
    NGP_REQUIRE(density_grid && density_bitfield, "null pointer");
    NGP_REQUIRE((reinterpret_cast<uintptr_t>(density_grid) & 15) == 0, "density_grid must be 16-byte aligned");
    packbits_kernel<<<(unsigned)((n_bytes + 255) / 256), 256, 0, ngp::as_stream(stream)>>>(
        density_grid, density_threshold, nullptr, density_bitfield, n_bytes);
    NGP_LAUNCHED("packbits_kernel");
    return 0;
}

int ngp_packbits_dev(const float* density_grid, const float* mean_density_dev, float density_threshold,
                     uint8_t* density_bitfield, int64_t n_bytes, void* stream) {
    NGP_REQUIRE(n_bytes >= 0, "negative n_bytes");
    if (n_bytes == 0) return 0;
    NGP_REQUIRE(density_grid && density_bitfield && mean_density_dev, "null pointer");
    NGP_REQUIRE((reinterpret_cast<uintptr_t>(density_grid) & 15) == 0, "density_grid must be 16-byte aligned");
    packbits_kernel<<<(unsigned)((n_bytes + 255) / 256), 256, 0, ngp::as_stream(stream)>>>(
        density_grid, density_threshold, mean_density_dev, density_bitfield, n_bytes);
    NGP_LAUNCHED("packbits_kernel");
    return 0;
}

int ngp_morton3d(const int32_t* coords, int32_t* indices, int64_t n, void* stream) {
    NGP_REQUIRE(n >= 0, "negative n");
    if (n == 0) return 0;
    NGP_REQUIRE(coords && indices, "null pointer");
    morton3d_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ngp::as_stream(stream)>>>(coords, indices, n);
    NGP_LAUNCHED("morton3d_kernel");
    return 0;
}

int ngp_morton3d_invert(const int32_t* indices, int32_t* coords, int64_t n, void* stream) {
    NGP_REQUIRE(n >= 0, "negative n");
    if (n == 0) return 0;
    NGP_REQUIRE(coords && indices, "null pointer");
    morton3d_invert_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ngp::as_stream(stream)>>>(indices, coords, n);
    NGP_LAUNCHED("morton3d_invert_kernel");
    return 0;
}

}  // extern "C"

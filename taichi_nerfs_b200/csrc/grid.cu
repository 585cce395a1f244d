// grid.cu — occupancy-grid helpers: packbits and Morton encode/decode.
// Semantics: modules/utils.py:95-169 of the reference (no ti.sync() host syncs here).
#include "common.cuh"
#include "philox.cuh"

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


// =====================================================================================================
// Fused occupancy-grid update.  Replaces, per update, the reference's
// sample_uniform_and_occupied_cells / get_all_cells (modules/networks.py:168-209: torch.randint, torch.nonzero +
// len() host sync, morton3D / morton3D_invert kernels each followed by ti.sync()), the jittered cell positions
// (:263-271), the scatter tmp[c, indices] = density (:272), the EMA-max (:276-279), the mean over positive cells with
// its .item() host sync (:286) and packbits (:288-290) by a fixed chain of launches without any host read:
//   occupied-cell words + per-block counts -> block scan -> cell pick + positions  |  (hash + sigma net, existing kernels)
//   -> scatter-max -> EMA + partial sums -> mean + packbits.
constexpr int kCellsPerBlock = 1024;   // cells per counting block (32 words of 32 cells)

// one warp = 32 consecutive cells -> one mask word; one 1024-thread block -> one count
__global__ void __launch_bounds__(1024) grid_occ_count_kernel(const float* __restrict__ grid, float thr,
                                                              uint32_t* __restrict__ words,
                                                              int32_t* __restrict__ block_count) {
    const int64_t i = (int64_t)blockIdx.x * kCellsPerBlock + threadIdx.x;
    const unsigned m = __ballot_sync(0xffffffffu, grid[i] > thr);
    __shared__ int32_t wc[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        words[i >> 5] = m;
        wc[wid] = __popc(m);
    }
    __syncthreads();
    if (wid == 0) {
        int32_t v = wc[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) block_count[blockIdx.x] = v;
    }
}

// exclusive scan of each cascade's block counts (one CTA per cascade); prefix has n_blocks + 1 entries per cascade
__global__ void __launch_bounds__(1024) grid_occ_scan_kernel(const int32_t* __restrict__ block_count,
                                                             int32_t* __restrict__ prefix, int n_blocks) {
    __shared__ int32_t warp_tot[32];
    __shared__ int32_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int32_t* in = block_count + (int64_t)blockIdx.x * n_blocks;
    int32_t* out = prefix + (int64_t)blockIdx.x * (n_blocks + 1);
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < n_blocks; base += 1024) {
        const int k = base + tid;
        const int32_t v = k < n_blocks ? in[k] : 0;
        int32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t nb = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += nb;
        }
        if (lane == 31) warp_tot[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            int32_t w = warp_tot[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int32_t nb = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += nb;
            }
            warp_tot[lane] = w;
        }
        __syncthreads();
        const int32_t carry = carry_s;
        const int32_t excl = carry + (wid ? warp_tot[wid - 1] : 0) + incl - v;
        if (k < n_blocks) out[k] = excl;
        __syncthreads();
        if (tid == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (tid == 0) out[n_blocks] = carry_s;
}

// thread (cascade c, slot i): which cell to evaluate and where inside it.
//   mode 0 (warm-up, get_all_cells): slot i = Morton index i, every cell once.
//   mode 1: slots [0, M) uniform cells (torch.randint coords, :187-189), slots [M, 2M) uniformly among the cells whose
//           density exceeds the threshold (torch.nonzero + randint pick, :190-198); no occupied cell -> index -1.
// Position (:263-271): ((coords / (G-1)) * 2 - 1) * (s - s/G) + (u * 2 - 1) * (s/G), u ~ U[0,1)^3, strict fp32.
__global__ void __launch_bounds__(256) grid_sample_cells_kernel(const uint32_t* __restrict__ words,
                                                                const int32_t* __restrict__ prefix, int n_blocks,
                                                                int cascades, int G, float scale, int mode, int64_t M,
                                                                int64_t per_cascade, uint64_t seed, uint32_t step,
                                                                int32_t* __restrict__ cell_idx, float* __restrict__ xyz) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= per_cascade * cascades) return;
    const int c = (int)(gid / per_cascade);
    const int64_t i = gid - (int64_t)c * per_cascade;
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    const Philox4 ra = philox4x32_10((uint32_t)i, (uint32_t)c, step, 0u, k0, k1);
    const Philox4 rb = philox4x32_10((uint32_t)i, (uint32_t)c, step, 1u, k0, k1);
    int32_t idx;
    uint32_t cx, cy, cz;
    if (mode == 0) {
        idx = (int32_t)i;
        cx = (uint32_t)compact_bits((uint32_t)idx >> 0);
        cy = (uint32_t)compact_bits((uint32_t)idx >> 1);
        cz = (uint32_t)compact_bits((uint32_t)idx >> 2);
    } else if (i < M) {
        cx = philox_below(ra.v[0], (uint32_t)G);
        cy = philox_below(ra.v[1], (uint32_t)G);
        cz = philox_below(ra.v[2], (uint32_t)G);
        idx = (int32_t)(expand_bits(cx) | (expand_bits(cy) << 1) | (expand_bits(cz) << 2));
    } else {
        const int32_t* pf = prefix + (int64_t)c * (n_blocks + 1);
        const int32_t total = pf[n_blocks];
        if (total <= 0) {
            cell_idx[gid] = -1;
            xyz[gid * 3 + 0] = xyz[gid * 3 + 1] = xyz[gid * 3 + 2] = 0.0f;
            return;
        }
        const int32_t k = (int32_t)philox_below(ra.v[3], (uint32_t)total);   // k-th occupied cell in index order
        int lo = 0, hi = n_blocks;                                          // last block with pf[b] <= k
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (pf[mid] <= k) lo = mid;
            else hi = mid;
        }
        int32_t rem = k - pf[lo];
        const uint32_t* w = words + ((int64_t)c * n_blocks + lo) * (kCellsPerBlock / 32);
        int wi = 0;
        uint32_t word = w[0];
        while (rem >= __popc(word)) {
            rem -= __popc(word);
            word = w[++wi];
        }
        idx = lo * kCellsPerBlock + wi * 32 + (int32_t)__fns(word, 0, rem + 1);
        cx = (uint32_t)compact_bits((uint32_t)idx >> 0);
        cy = (uint32_t)compact_bits((uint32_t)idx >> 1);
        cz = (uint32_t)compact_bits((uint32_t)idx >> 2);
    }
    const float s = fminf(exp2f((float)(c - 1)), scale);
    const float hg = f_div(s, (float)G);
    const float span = f_sub(s, hg);
    const uint32_t cc[3] = {cx, cy, cz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float base = f_mul(f_sub(f_mul(f_div((float)cc[d], (float)(G - 1)), 2.0f), 1.0f), span);
        const float jit = f_mul(f_sub(f_mul(philox_unit(rb.v[d]), 2.0f), 1.0f), hg);
        xyz[gid * 3 + d] = f_add(base, jit);
    }
    cell_idx[gid] = idx;
}

// tmp[c, idx] = max over the slots that picked the cell (torch's indexed assignment keeps an arbitrary one of the
// duplicates; the maximum is the deterministic choice).  Densities are exp() outputs, i.e. non-negative, so the
// integer order of their bit patterns is their numeric order.
__global__ void __launch_bounds__(256) grid_scatter_max_kernel(const int32_t* __restrict__ cell_idx,
                                                               const float* __restrict__ dens, int64_t per_cascade,
                                                               int64_t cells, int64_t n, int32_t* __restrict__ tmp) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= n) return;
    const int32_t idx = cell_idx[gid];
    const float d = dens[gid];
    if (idx < 0 || !(d >= 0.0f)) return;   // no occupied cell to pick / NaN
    atomicMax(tmp + (gid / per_cascade) * cells + idx, __float_as_int(d));
}

// grid = grid < 0 ? grid : max(grid * decay, tmp)  (:276-279, erode: per-cell decay clamp(decay^(1/count), 0.1, 0.95));
// each block also leaves (sum, count) of its positive cells for the deterministic mean
__global__ void __launch_bounds__(1024) grid_ema_kernel(float* __restrict__ grid, const float* __restrict__ tmp,
                                                        const float* __restrict__ count_grid, float decay,
                                                        double* __restrict__ part_sum, int32_t* __restrict__ part_cnt) {
    const int64_t i = (int64_t)blockIdx.x * 1024 + threadIdx.x;
    float g = grid[i];
    if (!(g < 0.0f)) {
        float dc = decay;
        if (count_grid != nullptr) dc = fminf(fmaxf(powf(decay, 1.0f / count_grid[i]), 0.1f), 0.95f);
        g = fmaxf(g * dc, tmp[i]);
        grid[i] = g;
    }
    const bool pos = g > 0.0f;
    double v = pos ? (double)g : 0.0;
    int32_t n = pos ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, o);
        n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    __shared__ double sv[32];
    __shared__ int32_t sn[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        sv[wid] = v;
        sn[wid] = n;
    }
    __syncthreads();
    if (wid == 0) {
        v = sv[lane];
        n = sn[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, o);
            n += __shfl_xor_sync(0xffffffffu, n, o);
        }
        if (lane == 0) {
            part_sum[blockIdx.x] = v;
            part_cnt[blockIdx.x] = n;
        }
    }
}

// mean over the positive cells in a fixed order (one CTA) -> *mean_out  (:286, without the .item())
__global__ void __launch_bounds__(1024) grid_mean_kernel(const double* __restrict__ part_sum,
                                                         const int32_t* __restrict__ part_cnt, int n_parts,
                                                         float* __restrict__ mean_out) {
    double v = 0.0;
    int64_t n = 0;
    for (int k = threadIdx.x; k < n_parts; k += 1024) {
        v += part_sum[k];
        n += part_cnt[k];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        v += __shfl_xor_sync(0xffffffffu, v, o);
        n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    __shared__ double sv[32];
    __shared__ int64_t sn[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        sv[wid] = v;
        sn[wid] = n;
    }
    __syncthreads();
    if (wid == 0) {
        v = sv[lane];
        n = sn[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            v += __shfl_xor_sync(0xffffffffu, v, o);
            n += __shfl_xor_sync(0xffffffffu, n, o);
        }
        if (lane == 0) *mean_out = (float)(v / (double)n);   // no positive cell: 0/0 = NaN, like torch's empty mean
    }
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


int64_t ngp_grid_workspace_bytes(int cascades, int grid_size) {
    const int64_t cells = (int64_t)grid_size * grid_size * grid_size;
    const int64_t blocks = cells / kCellsPerBlock;
    // mask words | block counts | block prefix (+1 per cascade) | tmp grid | partial sums | partial counts
    return cascades * (cells / 32 * 4 + blocks * 4 + (blocks + 1) * 4 + cells * 4 + blocks * 8 + blocks * 4) + 256;
}

namespace {
struct GridWs {
    uint32_t* words;
    int32_t *block_count, *prefix, *tmp, *part_cnt;
    double* part_sum;
};
GridWs carve(void* ws, int cascades, int64_t cells) {
    const int64_t blocks = cells / kCellsPerBlock;
    uint8_t* p = reinterpret_cast<uint8_t*>(ws);
    GridWs g;
    g.part_sum = reinterpret_cast<double*>(p);       p += cascades * blocks * 8;
    g.tmp = reinterpret_cast<int32_t*>(p);           p += cascades * cells * 4;
    g.words = reinterpret_cast<uint32_t*>(p);        p += cascades * (cells / 32) * 4;
    g.block_count = reinterpret_cast<int32_t*>(p);   p += cascades * blocks * 4;
    g.prefix = reinterpret_cast<int32_t*>(p);        p += cascades * (blocks + 1) * 4;
    g.part_cnt = reinterpret_cast<int32_t*>(p);
    return g;
}
}  // namespace

int ngp_grid_sample_cells(const float* density_grid, int cascades, int grid_size, float scale, float density_threshold,
                          int mode, int64_t M, uint64_t seed, uint32_t step, void* workspace, int32_t* cell_idx,
                          float* xyz, void* stream) {
    NGP_REQUIRE(cascades >= 1 && grid_size >= 32 && grid_size <= 1024 && (grid_size & (grid_size - 1)) == 0,
                "grid_size must be a power of two in [32, 1024]");
    NGP_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (all cells) or 1 (uniform + occupied)");
    NGP_REQUIRE(density_grid && workspace && cell_idx && xyz, "null pointer");
    NGP_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "workspace must be 16-byte aligned");
    const int64_t cells = (int64_t)grid_size * grid_size * grid_size;
    const int64_t per = mode == 0 ? cells : 2 * M;
    NGP_REQUIRE(mode == 0 || (M > 0 && M < (1ll << 31)), "M out of range");
    cudaStream_t st = ngp::as_stream(stream);
    const GridWs g = carve(workspace, cascades, cells);
    const int n_blocks = (int)(cells / kCellsPerBlock);
    if (mode == 1) {
        grid_occ_count_kernel<<<(unsigned)(cascades * n_blocks), 1024, 0, st>>>(density_grid, density_threshold, g.words,
                                                                              g.block_count);
        NGP_LAUNCHED("grid_occ_count_kernel");
        grid_occ_scan_kernel<<<(unsigned)cascades, 1024, 0, st>>>(g.block_count, g.prefix, n_blocks);
        NGP_LAUNCHED("grid_occ_scan_kernel");
    }
    const int64_t n = per * cascades;
    grid_sample_cells_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g.words, g.prefix, n_blocks, cascades, grid_size,
                                                                         scale, mode, M, per, seed, step, cell_idx, xyz);
    NGP_LAUNCHED("grid_sample_cells_kernel");
    return 0;
}

int ngp_grid_update(float* density_grid, const int32_t* cell_idx, const float* densities, int64_t per_cascade,
                    int cascades, int grid_size, const float* count_grid_or_null, float decay,
                    float density_threshold, void* workspace, float* mean_out, uint8_t* density_bitfield,
                    void* stream) {
    NGP_REQUIRE(cascades >= 1 && grid_size >= 32 && (grid_size & (grid_size - 1)) == 0, "bad grid_size");
    NGP_REQUIRE(density_grid && cell_idx && densities && workspace && mean_out && density_bitfield, "null pointer");
    NGP_REQUIRE((reinterpret_cast<uintptr_t>(density_grid) & 15) == 0, "density_grid must be 16-byte aligned");
    const int64_t cells = (int64_t)grid_size * grid_size * grid_size;
    cudaStream_t st = ngp::as_stream(stream);
    const GridWs g = carve(workspace, cascades, cells);
    cudaError_t e = cudaMemsetAsync(g.tmp, 0, (size_t)(cascades * cells * 4), st);
    if (e != cudaSuccess) {
        ngp::set_error("ngp_grid_update: cudaMemsetAsync: %s", cudaGetErrorString(e));
        return (int)e;
    }
    const int64_t n = per_cascade * cascades;
    grid_scatter_max_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(cell_idx, densities, per_cascade, cells, n, g.tmp);
    NGP_LAUNCHED("grid_scatter_max_kernel");
    const int n_parts = (int)(cascades * cells / 1024);
    grid_ema_kernel<<<(unsigned)n_parts, 1024, 0, st>>>(density_grid, reinterpret_cast<const float*>(g.tmp),
                                                       count_grid_or_null, decay, g.part_sum, g.part_cnt);
    NGP_LAUNCHED("grid_ema_kernel");
    grid_mean_kernel<<<1, 1024, 0, st>>>(g.part_sum, g.part_cnt, n_parts, mean_out);
    NGP_LAUNCHED("grid_mean_kernel");
    const int64_t n_bytes = cascades * cells / 8;
    packbits_kernel<<<(unsigned)((n_bytes + 255) / 256), 256, 0, st>>>(density_grid, density_threshold, mean_out,
                                                                      density_bitfield, n_bytes);
    NGP_LAUNCHED("packbits_kernel");
    return 0;
}

}  // extern "C"

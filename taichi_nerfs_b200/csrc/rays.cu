// rays.cu — training ray batch sampling: pixel selection, target gather,
// camera-to-world rotation and the per-ray marching jitter in one launch.
//
// Replaces, per step, BaseDataset.__getitem__ (datasets/base.py:34-61: two torch.randint, fancy-index
// gathers of rays / poses / directions) + get_rays (datasets/ray_utils.py:51-80) + the
// torch.rand_like jitter of RayMarcher (modules/ray_march.py:166).  The indices either come from the
// caller (the reference's own `img_idxs` / `pix_idxs`) or from Philox4x32-10 keyed by
// (seed, step, ray) so that the launch is CUDA-graph replayable: `step` is read from device memory.
#include "common.cuh"
#include "philox.cuh"

namespace {

__global__ void sample_ray_batch_kernel(const float* __restrict__ image_bank, int channels,
                                        const float* __restrict__ poses, const float* __restrict__ directions,
                                        int64_t n_img, int64_t n_pix, const int64_t* __restrict__ img_in,
                                        const int64_t* __restrict__ pix_in, int64_t fixed_img, uint64_t seed,
                                        const int32_t* __restrict__ step_dev, int32_t step_host,
                                        float* __restrict__ rays_o, float* __restrict__ rays_d,
                                        float* __restrict__ rgb, float* __restrict__ noise,
                                        int64_t* __restrict__ img_out, int64_t* __restrict__ pix_out,
                                        int64_t n_rays) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_rays) return;
    const uint32_t step = (uint32_t)(step_dev ? *step_dev : step_host);
    const Philox4 r = philox4x32_10((uint32_t)i, (uint32_t)((uint64_t)i >> 32), step, 0u, (uint32_t)seed,
                                    (uint32_t)(seed >> 32));
    // multiply-shift range reduction: floor(r * n / 2^32)
    int64_t img = img_in ? img_in[i] : (fixed_img >= 0 ? fixed_img : (int64_t)(((uint64_t)r.v[0] * (uint64_t)n_img) >> 32));
    int64_t pix = pix_in ? pix_in[i] : (int64_t)(((uint64_t)r.v[1] * (uint64_t)n_pix) >> 32);
    img = min(max(img, (int64_t)0), n_img - 1);
    pix = min(max(pix, (int64_t)0), n_pix - 1);

    const float* P = poses + img * 12;  // row-major [3][4] camera-to-world
    const float d0 = directions[pix * 3 + 0], d1 = directions[pix * 3 + 1], d2 = directions[pix * 3 + 2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        // rays_d[a] = sum_c d[c] * R[a][c]   (directions @ c2w[:, :3].T), strict fp32 in source order
        const float v = f_add(f_add(f_mul(d0, P[a * 4 + 0]), f_mul(d1, P[a * 4 + 1])), f_mul(d2, P[a * 4 + 2]));
        rays_d[i * 3 + a] = v;
        rays_o[i * 3 + a] = P[a * 4 + 3];
    }
    if (rgb) {
        const float* px = image_bank + (img * n_pix + pix) * channels;
        rgb[i * 3 + 0] = px[0];
        rgb[i * 3 + 1] = px[1];
        rgb[i * 3 + 2] = px[2];
    }
    if (noise) noise[i] = (float)(r.v[2] >> 8) * 5.9604644775390625e-8f;  // 24 bits -> [0, 1)
    if (img_out) img_out[i] = img;
    if (pix_out) pix_out[i] = pix;
}

}  // namespace

extern "C" int ngp_sample_ray_batch(const float* image_bank, int channels, const float* poses, const float* directions,
                                    int64_t n_img, int64_t n_pix, const int64_t* img_idxs, const int64_t* pix_idxs,
                                    int64_t fixed_img, uint64_t seed, const int32_t* step_dev, int32_t step_host,
                                    float* rays_o, float* rays_d, float* rgb, float* noise, int64_t* img_out,
                                    int64_t* pix_out, int64_t n_rays, void* stream) {
    NGP_REQUIRE(n_rays >= 0, "negative n_rays");
    if (n_rays == 0) return 0;
    NGP_REQUIRE(poses && directions && rays_o && rays_d, "null pointer");
    NGP_REQUIRE(n_img > 0 && n_pix > 0 && n_img < (1ll << 32) && n_pix < (1ll << 32), "n_img / n_pix out of range");
    NGP_REQUIRE(fixed_img < n_img, "fixed_img out of range");
    NGP_REQUIRE(!rgb || (image_bank && channels >= 3), "rgb output needs an image bank with >= 3 channels");
    sample_ray_batch_kernel<<<(unsigned)((n_rays + 255) / 256), 256, 0, ngp::as_stream(stream)>>>(
        image_bank, channels, poses, directions, n_img, n_pix, img_idxs, pix_idxs, fixed_img, seed, step_dev,
        step_host, rays_o, rays_d, rgb, noise, img_out, pix_out, n_rays);
    NGP_LAUNCHED("sample_ray_batch_kernel");
    return 0;
}

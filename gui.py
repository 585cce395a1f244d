"""Viewer entry point — mirrors the reference's gui.py: OrbitCamera (:28-74) and
NGPGUI(hparams, model_config, K, img_wh, poses, radius).render_cam() (:77-145).  The Taichi GGUI window
(Vulkan) is out of scope: render() drives the same per-frame path headlessly
along an orbit and writes PNG frames instead of presenting them."""
import os
import time
import warnings

import numpy as np
import torch

from datasets.ray_utils import get_ray_directions, get_rays
from modules.networks import NGP
from modules.rendering import render
from modules.utils import depth2img

warnings.filterwarnings("ignore")


def _rotvec_to_matrix(v):
    """Rodrigues formula (the reference uses scipy.spatial.transform.Rotation.from_rotvec)."""
    theta = np.linalg.norm(v)
    if theta < 1e-12:
        return np.eye(3)
    k = v / theta
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(theta) * K + (1 - np.cos(theta)) * (K @ K)


class OrbitCamera:
    def __init__(self, K, img_wh, poses, r):
        self.K = K
        self.W, self.H = img_wh
        self.radius = r
        self.center = np.zeros(3)
        pose_np = poses.cpu().numpy()
        self.rot = pose_np[0][:3, :3]  # initial rotation = first training pose
        self.rotate_speed = 0.8
        self.res_defalut = pose_np[0]

    @property
    def pose(self):
        res = np.eye(4)
        res[2, 3] -= self.radius       # move the camera back to the orbit radius
        rot = np.eye(4)
        rot[:3, :3] = self.rot
        res = rot @ res
        res[:3, 3] -= self.center
        return res

    def reset(self, pose=None):
        self.rot = np.eye(3)
        self.center = np.zeros(3)
        self.radius = 2.0
        if pose is not None:
            self.rot = pose.cpu().numpy()[:3, :3]

    def orbit(self, dx, dy):
        rx = self.rot[:, 1] * np.radians(100 * self.rotate_speed * dx)
        ry = self.rot[:, 0] * np.radians(-100 * self.rotate_speed * dy)
        self.rot = _rotvec_to_matrix(ry) @ _rotvec_to_matrix(rx) @ self.rot

    def scale(self, delta):
        self.radius *= 1.1 ** (-delta)

    def pan(self, dx, dy, dz=0):
        self.center += 1e-4 * self.rot @ np.array([dx, dy, dz])


class NGPGUI:
    def __init__(self, hparams, model_config, K, img_wh, poses, radius=4.5):
        self.hparams = hparams
        self.model = NGP(**model_config).cuda()
        if getattr(hparams, 'ckpt_path', None):
            print(f"loading ckpt from: {hparams.ckpt_path}")
            self.model.load_state_dict(torch.load(hparams.ckpt_path, map_location='cuda'))
        self.poses = poses
        self.cam = OrbitCamera(K, img_wh, poses, r=radius)
        self.W, self.H = img_wh
        self.exp_step_factor = 1 / 256 if hparams.dataset_name in ['colmap', 'nerfpp'] else 0
        self.dt = 0
        self.mean_samples = 0
        self.img_mode = 0
        self._directions = None

    @torch.no_grad()
    def render_cam(self):
        """One frame through render(test_time=True) (reference: gui.py:115-145)."""
        t = time.time()
        with torch.autocast(device_type='cuda', dtype=torch.float16):
            if self._directions is None:  # the reference rebuilds the pixel grid every frame (:118-123)
                self._directions = get_ray_directions(self.cam.H, self.cam.W, self.cam.K, device='cuda')
            pose = torch.tensor(self.cam.pose[:3], dtype=torch.float32, device='cuda')
            rays_o, rays_d = get_rays(self._directions, pose)
            results = render(self.model, rays_o, rays_d, test_time=True, exp_step_factor=self.exp_step_factor)
        rgb = results["rgb"].reshape(self.H, self.W, 3)
        depth = results["depth"].reshape(self.H, self.W)
        torch.cuda.synchronize()
        self.dt = time.time() - t
        self.mean_samples = results['total_samples'] / len(rays_o)
        if self.img_mode == 0:
            return rgb
        return torch.from_numpy(depth2img(depth.cpu().numpy()).astype(np.float32) / 255.0)

    def render(self, n_frames=8, out_dir='results/gui'):
        """Headless stand-in for the GGUI loop (gui.py:174-218): orbit the camera, dump frames."""
        from PIL import Image
        os.makedirs(out_dir, exist_ok=True)
        for i in range(n_frames):
            frame = self.render_cam()
            img = (frame.float().clamp(0, 1).cpu().numpy() * 255).astype(np.uint8)
            Image.fromarray(img).save(os.path.join(out_dir, f'frame_{i:03d}.png'))
            print(f'frame {i}: {1000 * self.dt:.2f} ms, {float(self.mean_samples):.2f} samples/ray')
            self.cam.orbit(0.05, 0.0)


if __name__ == "__main__":
    from datasets import dataset_dict
    from opt import get_opts
    from train import build_model_config
    hparams = get_opts()
    dataset = dataset_dict[hparams.dataset_name](root_dir=hparams.root_dir, downsample=hparams.downsample)
    NGPGUI(hparams, build_model_config(hparams), dataset.K, dataset.img_wh, dataset.poses).render()

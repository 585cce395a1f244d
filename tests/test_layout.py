"""Hash layout vs the constants the reference itself prints (the only numeric pins it ships)."""
import json
import math
import os

from conftest import GOLDEN
from taichi_nerfs_b200.layout import CHashLayout, make_hash_layout


def _consts():
    with open(os.path.join(GOLDEN, "layout_constants.json")) as f:
        return json.load(f)


def test_lego_layout_matches_notebook_printout():
    c = _consts()["lego_16_1024"]
    lay = make_hash_layout(2 ** 19, 16, 16, 1024, 2)
    assert math.exp(lay.log_b) == c["per_level_scale"]
    assert lay.total_entries == c["total_entries"]
    assert lay.total_param_size == c["total_params"]
    assert lay.begin_fast_hash_level == 6
    assert lay.layout_res == [16, 22, 28, 37, 49, 64, 85, 112, 148, 195, 256, 338, 446, 589, 777, 1024]
    # f32 kernel constants agree with the f64 sizing on every level (SURVEY §7 hard part 2)
    assert lay.resolutions == lay.layout_res
    assert lay.scales[0] == 15.0 and lay.scales[5] == 63.0 and lay.scales[10] == 255.0 and lay.scales[15] == 1023.0


def test_deployment_layout_matches_app_fp32():
    c = _consts()["deployment"]
    lay = make_hash_layout(2 ** 21, 4, 32, 128, 4)
    assert lay.total_param_size == c["total_params"]
    assert lay.offsets == c["offsets_entries"]
    assert lay.begin_fast_hash_level == 4  # all levels dense


def test_garden_layout():
    lay = make_hash_layout(2 ** 19, 16, 16, 4096, 2)
    assert lay.total_entries == 6299960
    assert lay.begin_fast_hash_level == 5
    assert lay.resolutions == lay.layout_res


def test_ctypes_mirror_size():
    import ctypes
    assert ctypes.sizeof(CHashLayout) == 16 + 4 * 16 * 4
    cl = make_hash_layout(2 ** 19, 16, 16, 1024, 2).as_ctypes()
    assert cl.n_levels == 16 and cl.offsets[15] == 5185744 and cl.map_sizes[0] == 4096


def test_synthetic_dataset_sampling_strategies():
    """datasets/base.py:34-52: 'all_images' draws an image per ray, 'same_image' takes every ray from image idx."""
    import torch
    from datasets.synthetic import SyntheticLego
    ds = SyntheticLego(n_images=7, img_wh=(40, 30), focal=50.0, batch_size=256)
    b = ds[3]
    assert b['rgb'].shape == (256, 3) and len(torch.unique(b['img_idxs'])) > 1
    ds.ray_sampling_strategy = 'same_image'
    b = ds[3]
    assert bool((b['img_idxs'] == 3).all()) and int(b['pix_idxs'].max()) < 1200
    assert torch.equal(b['pose'], ds.poses[3].expand(256, 3, 4))

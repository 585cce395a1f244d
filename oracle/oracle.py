"""ctypes/numpy binding of the CPU oracle (oracle/ngp_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs — never from the product
packages (taichi_nerfs_b200/, modules/).  Parity pin status: see the header of
ngp_oracle.c ("unpinned by the reference" per kernel; layout constants and the
shipped Lego deployment model are the pins).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "build", "libngp_oracle.so")

F32, F16 = 0, 1


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "ngp_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "ngp_b200.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"] + (["-B"] if force else []), check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.ngp_mlp_save_bytes_cpu.restype = C.c_int64
        _lib.ngp_mlp_save_bytes_cpu.argtypes = [C.c_int64]
    return _lib


def set_threads(n: int) -> int:
    """Force the OpenMP team size (torchrun exports OMP_NUM_THREADS=1) and return the number of threads a parallel
    region REALLY runs with (counted, not assumed)."""
    lib().ngp_oracle_set_threads(int(n))
    return int(lib().ngp_oracle_threads_used())


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def _dtype_tag(a):
    if a.dtype == np.float16:
        return F16
    if a.dtype == np.float32:
        return F32
    raise TypeError(a.dtype)


class MlpWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("w1", "w2", "w3", "w4", "w5")]


def _chk(rc):
    if rc != 0:
        raise RuntimeError(f"oracle returned {rc}")


# ---------------------------------------------------------------------------
def ray_aabb_intersect(rays_o, rays_d, scale):
    o, d = _c(rays_o, np.float32), _c(rays_d, np.float32)
    hits = np.empty((o.shape[0], 2), np.float32)
    _chk(lib().ngp_ray_aabb_intersect_cpu(_p(o), _p(d), C.c_float(scale), _p(hits), C.c_int64(o.shape[0])))
    return hits


def raymarching_train(rays_o, rays_d, hits_t, bitfield, noise, cascades, scale, exp_step_factor,
                      grid_size, max_samples):
    o, d, h = _c(rays_o, np.float32), _c(rays_d, np.float32), _c(hits_t, np.float32)
    bf, nz = _c(bitfield, np.uint8), _c(noise, np.float32)
    n = o.shape[0]
    counter = np.zeros(2, np.int32)
    rays_a = np.zeros((n, 3), np.int32)
    L = lib()
    _chk(L.ngp_raymarching_train_count_cpu(_p(o), _p(d), _p(h), _p(bf), _p(nz), C.c_int(cascades),
                                           C.c_int(grid_size), C.c_float(scale), C.c_float(exp_step_factor),
                                           C.c_int(int(max_samples)), _p(counter), _p(rays_a), C.c_int64(n)))
    S = int(counter[0])
    xyzs = np.empty((S, 3), np.float32)
    dirs = np.empty((S, 3), np.float32)
    deltas = np.empty(S, np.float32)
    ts = np.empty(S, np.float32)
    _chk(L.ngp_raymarching_train_write_cpu(_p(o), _p(d), _p(h), _p(bf), _p(nz), C.c_int(cascades),
                                           C.c_int(grid_size), C.c_float(scale), C.c_float(exp_step_factor),
                                           _p(counter), _p(rays_a), _p(xyzs), _p(dirs), _p(deltas), _p(ts),
                                           C.c_int64(n), C.c_int64(S)))
    return rays_a, xyzs, dirs, deltas, ts, S


def raymarching_test(rays_o, rays_d, hits_t, alive_indices, bitfield, cascades, scale, exp_step_factor,
                     grid_size, max_samples):
    """hits_t is updated in place (must be a contiguous float32 array)."""
    o, d = _c(rays_o, np.float32), _c(rays_d, np.float32)
    assert hits_t.dtype == np.float32 and hits_t.flags.c_contiguous
    alive = _c(alive_indices, np.int64)
    bf = _c(bitfield, np.uint8)
    A = alive.shape[0]
    m = int(max_samples)
    ray_indices = np.zeros(A * m, np.int64)
    valid = np.zeros(A * m, np.uint8)
    deltas = np.zeros(A * m, np.float32)
    ts = np.zeros(A * m, np.float32)
    cnt = np.zeros(A, np.int32)
    _chk(lib().ngp_raymarching_test_cpu(_p(o), _p(d), _p(hits_t), _p(alive), _p(bf), C.c_int(cascades),
                                        C.c_int(grid_size), C.c_float(scale), C.c_float(exp_step_factor),
                                        C.c_int(m), _p(ray_indices), _p(valid), _p(deltas), _p(ts), _p(cnt),
                                        C.c_int64(A)))
    return ray_indices, valid, deltas, ts, cnt


def hash_encode_fwd(xyz, table, layout):
    x = _c(xyz, np.float32)
    tag = _dtype_tag(table)
    tab = np.ascontiguousarray(table)
    n = x.shape[0]
    out = np.empty((n, layout.out_dim), tab.dtype)
    cl = layout.as_ctypes()
    _chk(lib().ngp_hash_encode_fwd_cpu(_p(x), _p(tab), C.byref(cl), _p(out), C.c_int(tag), C.c_int64(n)))
    return out


def hash_encode_bwd(xyz, dout, layout, grad_table=None):
    x = _c(xyz, np.float32)
    dy = np.ascontiguousarray(dout)
    if grad_table is None:
        grad_table = np.zeros(layout.total_param_size, np.float32)
    cl = layout.as_ctypes()
    _chk(lib().ngp_hash_encode_bwd_cpu(_p(x), _p(dy), C.c_int(_dtype_tag(dy)), C.byref(cl), _p(grad_table),
                                       C.c_int64(x.shape[0])))
    return grad_table


def hash_encode_bwd_input(xyz, table, dout, layout):
    x = _c(xyz, np.float32)
    tab, dy = np.ascontiguousarray(table), np.ascontiguousarray(dout)
    assert tab.dtype == dy.dtype
    dx = np.empty((x.shape[0], 3), np.float32)
    cl = layout.as_ctypes()
    _chk(lib().ngp_hash_encode_bwd_input_cpu(_p(x), _p(tab), _p(dy), C.c_int(_dtype_tag(tab)), C.byref(cl),
                                             _p(dx), C.c_int64(x.shape[0])))
    return dx


def dir_encode(dirs):
    d = _c(dirs, np.float32)
    out = np.empty((d.shape[0], 16), np.float32)
    _chk(lib().ngp_dir_encode_cpu(_p(d), _p(out), C.c_int64(d.shape[0])))
    return out


def _weights(ws):
    arrs = [_c(w, np.float32) for w in ws]
    shapes = [(64, 32), (16, 64), (64, 32), (64, 64), (3, 64)]
    for a, s in zip(arrs, shapes):
        assert a.shape == s, (a.shape, s)
    st = MlpWeights(*[a.ctypes.data for a in arrs])
    return st, arrs


def mlp_fwd(emb, dirs, weights, save=False):
    e = np.ascontiguousarray(emb)
    d = _c(dirs, np.float32)
    n = e.shape[0]
    st, keep = _weights(weights)
    sig = np.empty(n, np.float32)
    rgb = np.empty((n, 3), np.float16)
    sv = np.empty(lib().ngp_mlp_save_bytes_cpu(n), np.uint8) if save else None
    _chk(lib().ngp_mlp_fwd_cpu(_p(e), C.c_int(_dtype_tag(e)), _p(d), C.byref(st), _p(sig), _p(rgb), _p(sv),
                               C.c_int64(n)))
    return (sig, rgb, sv) if save else (sig, rgb)


def mlp_bwd(emb, dirs, weights, dsigmas, drgbs, save=None):
    e = np.ascontiguousarray(emb)
    d = _c(dirs, np.float32)
    n = e.shape[0]
    st, keep = _weights(weights)
    ds = _c(dsigmas, np.float32)
    dr = _c(drgbs, np.float16)
    demb = np.empty((n, 32), e.dtype)
    gw = np.zeros(9408, np.float32)
    _chk(lib().ngp_mlp_bwd_cpu(_p(e), C.c_int(_dtype_tag(e)), _p(d), C.byref(st), _p(save), _p(ds), _p(dr),
                               _p(demb), _p(gw), C.c_int64(n)))
    return demb, gw


def composite_train_fwd(sigmas, rgbs, deltas, ts, rays_a, T_threshold):
    sg, dl, t = _c(sigmas, np.float32), _c(deltas, np.float32), _c(ts, np.float32)
    rg = np.ascontiguousarray(rgbs)
    ra = _c(rays_a, np.int32)
    n, S = ra.shape[0], sg.shape[0]
    tot = np.zeros(n, np.int32)
    op = np.zeros(n, np.float32)
    dp = np.zeros(n, np.float32)
    rgb = np.zeros((n, 3), np.float32)
    ws = np.zeros(S, np.float32)
    _chk(lib().ngp_composite_train_fwd_cpu(_p(sg), _p(rg), C.c_int(_dtype_tag(rg)), _p(dl), _p(t), _p(ra),
                                           C.c_float(T_threshold), _p(tot), _p(op), _p(dp), _p(rgb), _p(ws),
                                           C.c_int64(n), C.c_int64(S)))
    return tot, op, dp, rgb, ws


def composite_train_bwd(dL_dopacity, dL_ddepth, dL_drgb, dL_dws, sigmas, rgbs, deltas, ts, rays_a,
                        T_threshold):
    sg, dl, t = _c(sigmas, np.float32), _c(deltas, np.float32), _c(ts, np.float32)
    rg = np.ascontiguousarray(rgbs)
    ra = _c(rays_a, np.int32)
    n, S = ra.shape[0], sg.shape[0]
    go, gd, gr, gw = (_c(a, np.float32) for a in (dL_dopacity, dL_ddepth, dL_drgb, dL_dws))
    dsg = np.zeros(S, np.float32)
    drg = np.zeros((S, 3), rg.dtype)
    _chk(lib().ngp_composite_train_bwd_cpu(_p(go), _p(gd), _p(gr), _p(gw), _p(sg), _p(rg),
                                           C.c_int(_dtype_tag(rg)), _p(dl), _p(t), _p(ra), None, None, None,
                                           C.c_float(T_threshold), _p(dsg), _p(drg), C.c_int64(n), C.c_int64(S)))
    return dsg, drg


def composite_test(sigmas, rgbs, deltas, ts, pack_info, alive_indices, T_threshold, opacity, depth, rgb):
    """opacity/depth/rgb/alive_indices are updated in place."""
    sg, dl, t = _c(sigmas, np.float32), _c(deltas, np.float32), _c(ts, np.float32)
    rg = np.ascontiguousarray(rgbs)
    pk = _c(pack_info, np.int64)
    for a, dt in ((alive_indices, np.int64), (opacity, np.float32), (depth, np.float32), (rgb, np.float32)):
        assert a.dtype == dt and a.flags.c_contiguous
    _chk(lib().ngp_composite_test_cpu(_p(sg), _p(rg), C.c_int(_dtype_tag(rg)), _p(dl), _p(t), _p(pk),
                                      _p(alive_indices), C.c_float(T_threshold), _p(opacity), _p(depth),
                                      _p(rgb), C.c_int64(alive_indices.shape[0])))


def distortion_fwd(ws, deltas, ts, rays_a):
    w, dl, t = _c(ws, np.float32), _c(deltas, np.float32), _c(ts, np.float32)
    ra = _c(rays_a, np.int32)
    loss = np.zeros(ra.shape[0], np.float32)
    _chk(lib().ngp_distortion_fwd_cpu(_p(w), _p(dl), _p(t), _p(ra), _p(loss), C.c_int64(ra.shape[0]),
                                      C.c_int64(w.shape[0])))
    return loss


def distortion_bwd(dL_dloss, ws, deltas, ts, rays_a):
    g, w, dl, t = _c(dL_dloss, np.float32), _c(ws, np.float32), _c(deltas, np.float32), _c(ts, np.float32)
    ra = _c(rays_a, np.int32)
    out = np.zeros(w.shape[0], np.float32)
    _chk(lib().ngp_distortion_bwd_cpu(_p(g), _p(w), _p(dl), _p(t), _p(ra), _p(out), C.c_int64(ra.shape[0]),
                                      C.c_int64(w.shape[0])))
    return out


def packbits(density_grid, threshold):
    g = _c(density_grid, np.float32).reshape(-1)
    out = np.empty(g.shape[0] // 8, np.uint8)
    _chk(lib().ngp_packbits_cpu(_p(g), C.c_float(threshold), _p(out), C.c_int64(out.shape[0])))
    return out


def morton3d(coords):
    c = _c(coords, np.int32)
    out = np.empty(c.shape[0], np.int32)
    _chk(lib().ngp_morton3d_cpu(_p(c), _p(out), C.c_int64(c.shape[0])))
    return out


def morton3d_invert(indices):
    i = _c(indices, np.int32)
    out = np.empty((i.shape[0], 3), np.int32)
    _chk(lib().ngp_morton3d_invert_cpu(_p(i), _p(out), C.c_int64(i.shape[0])))
    return out


def check_finite(grad):
    g = _c(grad, np.float32).reshape(-1)
    flag = np.zeros(1, np.int32)
    _chk(lib().ngp_check_finite_cpu(_p(g), C.c_int64(g.shape[0]), _p(flag)))
    return int(flag[0])


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, step, beta1=0.9, beta2=0.999, eps=1e-15,
              inv_scale=1.0, param_f16=None, found_inf=None, zero_grad=False):
    """All arrays are flat contiguous float32 and updated in place."""
    for a in (param, grad, exp_avg, exp_avg_sq):
        assert a.dtype == np.float32 and a.flags.c_contiguous
    fi = None if found_inf is None else np.asarray([found_inf], np.int32)
    _chk(lib().ngp_adam_step_cpu(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), _p(param_f16), _p(fi),
                                 C.c_float(lr), C.c_float(beta1), C.c_float(beta2), C.c_float(eps),
                                 C.c_float(inv_scale), C.c_int32(step), C.c_int(int(zero_grad)),
                                 C.c_int64(param.size)))


def philox4x32_10(ctr, key):
    c, k = _c(ctr, np.uint32), _c(key, np.uint32)
    out = np.empty(4, np.uint32)
    lib().ngp_philox4x32_10_cpu(_p(c), _p(k), _p(out))
    return out


def sample_ray_batch(image_bank, poses, directions, n_rays, img_idxs=None, pix_idxs=None, fixed_img=-1, seed=0,
                     step=0):
    """datasets/base.py:34-61 + datasets/ray_utils.py:51-80 (+ the marching jitter) for one batch."""
    bank = None if image_bank is None else _c(image_bank, np.float32)
    P, D = _c(poses, np.float32), _c(directions, np.float32)
    ii = None if img_idxs is None else _c(img_idxs, np.int64)
    pi = None if pix_idxs is None else _c(pix_idxs, np.int64)
    o, d = np.empty((n_rays, 3), np.float32), np.empty((n_rays, 3), np.float32)
    rgb = None if bank is None else np.empty((n_rays, 3), np.float32)
    noise = np.empty(n_rays, np.float32)
    io, po = np.empty(n_rays, np.int64), np.empty(n_rays, np.int64)
    n = lambda a: None if a is None else _p(a)  # noqa: E731
    lib().ngp_sample_ray_batch_cpu.restype = C.c_int
    _chk(lib().ngp_sample_ray_batch_cpu(n(bank), C.c_int(0 if bank is None else bank.shape[2]), _p(P), _p(D),
                                        C.c_int64(P.shape[0]), C.c_int64(D.shape[0]), n(ii), n(pi),
                                        C.c_int64(fixed_img), C.c_uint64(seed), C.c_int32(step), _p(o), _p(d), n(rgb),
                                        _p(noise), _p(io), _p(po), C.c_int64(n_rays)))
    return {"rays_o": o, "rays_d": d, "rgb": rgb, "noise": noise, "img_idxs": io, "pix_idxs": po}


def raymarching_cellstep(rays_o, rays_d, hits_t, bitfield, noise, cascades, scale, grid_size, max_samples, rays_a,
                         n_samples):
    """Cell-stepping blueprint (exp_step_factor == 0): sample times in the layout of ``rays_a`` + loop statistics."""
    o, d, h = _c(rays_o, np.float32), _c(rays_d, np.float32), _c(hits_t, np.float32)
    bf, nz, ra = _c(bitfield, np.uint8), _c(noise, np.float32), _c(rays_a, np.int32)
    n = o.shape[0]
    ts = np.zeros(n_samples, np.float32)
    counts = np.zeros(n, np.int32)
    stats = np.zeros(2, np.int64)
    _chk(lib().ngp_raymarching_cellstep_cpu(_p(o), _p(d), _p(h), _p(bf), _p(nz), C.c_int(cascades), C.c_int(grid_size),
                                            C.c_float(scale), C.c_int(max_samples), _p(ra), _p(ts), _p(counts),
                                            _p(stats), C.c_int64(n)))
    return ts, counts, {"iterations": int(stats[0]), "real_adds": int(stats[1])}


def raymarching_lanes(rays_o, rays_d, hits_t, bitfield, noise, scale, grid_size, max_samples, rays_a, n_samples):
    """Lane-level emulation of the planned marching fast path (cascades == 1, exp_step_factor == 0)."""
    o, d, h = _c(rays_o, np.float32), _c(rays_d, np.float32), _c(hits_t, np.float32)
    bf, nz, ra = _c(bitfield, np.uint8), _c(noise, np.float32), _c(rays_a, np.int32)
    n = o.shape[0]
    ts = np.zeros(n_samples, np.float32)
    counts = np.zeros(n, np.int32)
    stats = np.zeros(2, np.int64)
    _chk(lib().ngp_raymarching_lanes_cpu(_p(o), _p(d), _p(h), _p(bf), _p(nz), C.c_int(grid_size), C.c_float(scale),
                                         C.c_int(max_samples), _p(ra), _p(ts), _p(counts), _p(stats), C.c_int64(n)))
    return ts, counts, {"regular_chunks": int(stats[0]), "general_chunks": int(stats[1])}


# ---- fused occupancy-grid update (restates modules/networks.py:168-209, 255-290 with the Philox draws of grid.cu) ---
def philox4x32_10_np(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 (same rounds as ngp_philox4x32_10_cpu, pinned on the Random123 vectors by
    tests/test_oracle.py).  Inputs broadcastable uint32 arrays; returns 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) for v in np.broadcast_arrays(c0, c1, c2, c3))
    k0 = np.uint64(k0)
    k1 = np.uint64(k1)
    M = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & M, p1 >> np.uint64(32), p1 & M
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & M, lo1, (hi0 ^ c3 ^ k1) & M, lo0
        k0 = (k0 + np.uint64(0x9E3779B9)) & M
        k1 = (k1 + np.uint64(0xBB67AE85)) & M
    return tuple(v.astype(np.uint32) for v in (c0, c1, c2, c3))


def grid_sample_cells(density_grid, scale, density_threshold, warmup, M, seed, step):
    """(cell_idx [C*slots] int32, xyz [C*slots, 3] f32): get_all_cells (networks.py:168-179) in warm-up mode, else
    sample_uniform_and_occupied_cells (:181-209) - M uniform cells then M picks among nonzero(grid > thr) in index
    order - and the jittered positions of update_density_grid (:263-271), strict fp32."""
    f = np.float32
    C_, cells = density_grid.shape
    G = round(cells ** (1 / 3))
    slots = cells if warmup else 2 * M
    k0, k1 = seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF
    idx_all, xyz_all = [], []
    i = np.arange(slots, dtype=np.uint32)
    for c in range(C_):
        ra = philox4x32_10_np(i, np.uint32(c), np.uint32(step), np.uint32(0), k0, k1)
        rb = philox4x32_10_np(i, np.uint32(c), np.uint32(step), np.uint32(1), k0, k1)
        if warmup:
            idx = i.astype(np.int64)
        else:
            below = lambda r, n: ((r.astype(np.uint64) * np.uint64(n)) >> np.uint64(32)).astype(np.int64)  # noqa: E731
            coords1 = np.stack([below(ra[d][:M], G) for d in range(3)], -1).astype(np.int32)
            idx1 = morton3d(coords1).astype(np.int64)
            occ = np.nonzero(density_grid[c] > f(density_threshold))[0]
            if len(occ) > 0:
                idx2 = occ[below(ra[3][M:], len(occ))]
            else:
                idx2 = np.full(M, -1, np.int64)
            idx = np.concatenate([idx1, idx2])
        coords = morton3d_invert(np.maximum(idx, 0).astype(np.int32)).astype(f)
        s = f(min(2.0 ** (c - 1), scale))
        hg = f(s / f(G))
        span = f(s - hg)
        u = np.stack([(rb[d] >> np.uint32(8)).astype(f) * f(5.9604644775390625e-8) for d in range(3)], -1)
        xyz = ((coords / f(G - 1)) * f(2) - f(1)) * span + (u * f(2) - f(1)) * hg
        xyz[idx < 0] = 0
        idx_all.append(idx.astype(np.int32))
        xyz_all.append(xyz.astype(f))
    return np.concatenate(idx_all), np.concatenate(xyz_all)


def grid_update(density_grid, cell_idx, densities, density_threshold, decay=0.95, count_grid=None):
    """(new grid [C, cells] f32, mean f32, bitfield u8): tmp[c, idx] = density (max over duplicates), EMA-max
    (networks.py:276-279), mean of the positive cells (:286), packbits with min(mean, threshold) (:288-290)."""
    f = np.float32
    C_, cells = density_grid.shape
    per = cell_idx.size // C_
    tmp = np.zeros_like(density_grid)
    for c in range(C_):
        idx, d = cell_idx[c * per:(c + 1) * per], densities[c * per:(c + 1) * per]
        ok = (idx >= 0) & (d >= 0)
        np.maximum.at(tmp[c], idx[ok], d[ok].astype(f))
    dec = f(decay)
    if count_grid is not None:
        dec = np.clip(np.power(f(decay), f(1) / count_grid.astype(f)).astype(f), f(0.1), f(0.95))
    new = np.where(density_grid < 0, density_grid, np.maximum((density_grid * dec).astype(f), tmp)).astype(f)
    pos = new > 0
    mean = f(new[pos].astype(np.float64).sum() / pos.sum()) if pos.any() else f(np.nan)
    thr = min(float(mean), float(density_threshold)) if pos.any() else float(density_threshold)
    return new, mean, packbits(new.reshape(-1), thr)


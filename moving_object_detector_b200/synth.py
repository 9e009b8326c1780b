"""Seeded synthetic stereo / flow / ego-motion data for tests and bench.py (builder's choice --
the reference pins no data).  Plumbing only: torch ops on CPU or CUDA, never on the measured path.

A sequence is rendered analytically: ground plane + back wall seen from a camera that yaws and drives forward,
plus axis-aligned moving boxes.  Frame i's disparity is the scene at pose i; pair i = (prev = frame i,
now = frame i+1) gets the ego-induced ("static") flow computed from the noise-free previous depth, overwritten by
the box displacement inside the boxes' now-rectangles, which is how the reference samples it (flow indexed at the
now pixel, scene_flow_constructor.h:205-211).  Disparity noise, invalid encodings {0, -1, NaN, +inf, > max},
NaN flow and out-of-image border flow are sprinkled in so every branch of the per-pixel state machine is exercised.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Tuple

import numpy as np
import torch


@dataclass
class SynthConfig:
    w: int
    h: int
    fx: float
    cx: float
    cy: float
    baseline: float
    dt: float
    n_movers: int = 2
    z_wall: float = 20.0
    cam_height: float = 1.5
    max_yaw_deg: float = 0.5
    max_forward: float = 0.02
    min_disparity: float = 0.0
    max_disparity: float = 128.0
    disp_noise: float = 0.25
    flow_noise: float = 0.3
    invalid_frac: float = 0.02
    n_invalid_rects: int = 3
    nan_flow_frac: float = 0.001
    border_flow_frac: float = 0.01
    identity_ego: bool = False
    seed: int = 0

    @property
    def cam(self) -> Tuple[float, float, float, float, float, float]:
        return (self.fx, self.fx, self.cx, self.cy, 0.0, 0.0)


# The five BASELINE.json configurations.
CONFIGS = {
    "gazebo_640x480": SynthConfig(640, 480, 554.25, 320.5, 240.5, 0.07, 1.0 / 30.0, n_movers=2, seed=1001),
    "zed_1280x720": SynthConfig(1280, 720, 700.0, 640.0, 360.0, 0.12, 1.0 / 15.0, n_movers=4, max_yaw_deg=1.0,
                                max_forward=0.05, seed=2002),
    "kitti_1242x375": SynthConfig(1242, 375, 721.5377, 609.5593, 172.854, 0.5372, 0.1, n_movers=6, z_wall=80.0,
                                  cam_height=1.65, max_forward=0.8, seed=3003),
    "zed2k_2208x1242": SynthConfig(2208, 1242, 1400.0, 1104.0, 621.0, 0.12, 1.0 / 15.0, n_movers=8, seed=4004),
    "uhd_3840x2160": SynthConfig(3840, 2160, 2800.0, 1920.0, 1080.0, 0.12, 1.0 / 15.0, n_movers=6, z_wall=80.0,
                                 seed=5005),
}


def _tri(x: float, lo: float, hi: float) -> float:
    """Triangle wave: bounces x between lo and hi."""
    span = hi - lo
    if span <= 0:
        return lo
    y = (x - lo) % (2 * span)
    return lo + (y if y <= span else 2 * span - y)


def ros_dt(dt_sec: float) -> float:
    """dt as a ROS node sees it: header stamps have nanosecond resolution and (stamp_now - stamp_prev).toSec() is
    (double)sec + 1e-9 * (double)nsec (scene_flow_constructor.cpp:162-164)."""
    sec, nsec = divmod(int(round(float(dt_sec) * 1e9)), 1000000000)
    return float(sec) + 1e-9 * float(nsec)


def make_sequence(cfg: SynthConfig, n_frames: int, device="cpu"):
    """Returns dict(disp (F,H,W) f32, flows (F-1,H,W,2) f32, transforms [(tx,ty,tz,qx,qy,qz,qw)], dts, cam, f, T,
    min_disparity, max_disparity)."""
    assert n_frames >= 2
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(cfg.seed)
    rng = np.random.RandomState(cfg.seed)
    W, H, F = cfg.w, cfg.h, n_frames
    f64 = torch.float64

    u = torch.arange(W, device=dev, dtype=f64)
    v = torch.arange(H, device=dev, dtype=f64)
    rx = ((u - cfg.cx) / cfg.fx)[None, :].expand(H, W)
    ry = ((v - cfg.cy) / cfg.fx)[:, None].expand(H, W)

    # ---- ego trajectory: P_now = R(yaw) P_prev + t, t = (small lateral, 0, -forward)
    transforms, dts, Rs, ts = [], [], [], []
    for i in range(F - 1):
        if cfg.identity_ego:
            yaw, fwd, lat = 0.0, 0.0, 0.0
        else:
            yaw = math.radians(cfg.max_yaw_deg) * math.sin(0.37 * i + 0.5) * rng.uniform(0.5, 1.0)
            fwd = cfg.max_forward * rng.uniform(0.5, 1.0)
            lat = 0.1 * cfg.max_forward * rng.uniform(-1.0, 1.0)
        c, s = math.cos(yaw), math.sin(yaw)
        R = np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])
        t = np.array([lat, 0.0, -fwd])
        Rs.append(R)
        ts.append(t)
        transforms.append((t[0], t[1], t[2], 0.0, math.sin(yaw / 2), 0.0, math.cos(yaw / 2)))
        dts.append(ros_dt(cfg.dt))

    # ---- back wall n.P = c in each frame (ground stays Y = cam_height under yaw + planar translation)
    n_wall, c_wall = [np.array([0.0, 0.0, 1.0])], [cfg.z_wall]
    for i in range(F - 1):
        n_next = Rs[i] @ n_wall[-1]
        n_wall.append(n_next)
        c_wall.append(c_wall[-1] + float(n_next @ ts[i]))

    # ---- movers: rectangles bouncing inside the image, constant depth drift
    movers = []
    for k in range(cfg.n_movers):
        hw = int(W * rng.uniform(0.05, 0.09))
        hh = int(H * rng.uniform(0.10, 0.16))
        movers.append(dict(
            hw=hw, hh=hh,
            x0=rng.uniform(hw, W - hw), y0=rng.uniform(hh, H - hh),
            mx=rng.choice([-1.0, 1.0]) * rng.uniform(12.0, 24.0), my=rng.uniform(-2.0, 2.0),
            z0=rng.uniform(0.2, 0.45) * cfg.z_wall, dz=rng.uniform(-0.03, 0.03)))

    def mover_pos(m, i):
        return (_tri(m["x0"] + i * m["mx"], m["hw"], W - m["hw"]), _tri(m["y0"] + i * m["my"], m["hh"], H - m["hh"]),
                m["z0"] + i * m["dz"])

    disp = torch.empty((F, H, W), dtype=torch.float32, device=dev)
    flows = torch.empty((F - 1, H, W, 2), dtype=torch.float32, device=dev)
    big = torch.tensor(1e9, dtype=f64, device=dev)
    z_prev_clean = None
    for i in range(F):
        ground = torch.where(ry > 1e-6, cfg.cam_height / ry.clamp_min(1e-6), big)
        nw = n_wall[i]
        denom = nw[0] * rx + nw[1] * ry + nw[2]
        wall = torch.where(denom > 1e-6, c_wall[i] / denom.clamp_min(1e-6), big)
        Z = torch.minimum(ground, wall).clamp(0.5, 4.0 * cfg.z_wall)
        for m in movers:
            bx, by, bz = mover_pos(m, i)
            x0, x1 = int(round(bx - m["hw"])), int(round(bx + m["hw"]))
            y0, y1 = int(round(by - m["hh"])), int(round(by + m["hh"]))
            x0, y0 = max(x0, 0), max(y0, 0)
            Z[y0:y1, x0:x1] = torch.minimum(Z[y0:y1, x0:x1], torch.tensor(bz, dtype=f64, device=dev))
        d = cfg.fx * cfg.baseline / Z + cfg.disp_noise * torch.randn((H, W), generator=gen, device=dev, dtype=f64)
        d = d.clamp(cfg.min_disparity, cfg.max_disparity).to(torch.float32)
        # invalid encodings: speckles + rectangles
        enc = torch.tensor([0.0, -1.0, float("nan"), float("inf"), cfg.max_disparity + 72.0], device=dev)
        if cfg.invalid_frac > 0:
            r = torch.rand((H, W), generator=gen, device=dev)
            which = torch.randint(0, 5, (H, W), generator=gen, device=dev)
            d = torch.where(r < cfg.invalid_frac, enc[which], d)
        for k in range(cfg.n_invalid_rects):
            rw, rh = max(2, int(W * rng.uniform(0.01, 0.04))), max(2, int(H * rng.uniform(0.01, 0.05)))
            x0, y0 = rng.randint(0, max(1, W - rw)), rng.randint(0, max(1, H - rh))
            d[y0:y0 + rh, x0:x0 + rw] = enc[(i + k) % 5]
        disp[i] = d

        if i > 0:
            # pair i-1: prev = frame i-1, now = frame i.  Ego-induced flow from the clean previous depth.
            R, t = Rs[i - 1], ts[i - 1]
            Zp = z_prev_clean
            X, Y = rx * Zp, ry * Zp
            qx = R[0, 0] * X + R[0, 1] * Y + R[0, 2] * Zp + t[0]
            qy = R[1, 0] * X + R[1, 1] * Y + R[1, 2] * Zp + t[1]
            qz = R[2, 0] * X + R[2, 1] * Y + R[2, 2] * Zp + t[2]
            fu = cfg.fx * qx / qz + cfg.cx - u[None, :]
            fv = cfg.fx * qy / qz + cfg.cy - v[:, None]
            fl = torch.stack([fu, fv], dim=-1)
            for m in movers:
                bx, by, _ = mover_pos(m, i)
                pbx, pby, _ = mover_pos(m, i - 1)
                x0, x1 = max(int(round(bx - m["hw"])), 0), int(round(bx + m["hw"]))
                y0, y1 = max(int(round(by - m["hh"])), 0), int(round(by + m["hh"]))
                fl[y0:y1, x0:x1, 0] = bx - pbx
                fl[y0:y1, x0:x1, 1] = by - pby
            fl = fl + cfg.flow_noise * torch.randn((H, W, 2), generator=gen, device=dev, dtype=f64)
            fl = fl.to(torch.float32)
            if cfg.nan_flow_frac > 0:
                r = torch.rand((H, W), generator=gen, device=dev)
                fl[..., 0] = torch.where(r < cfg.nan_flow_frac, torch.tensor(float("nan"), device=dev), fl[..., 0])
                r = torch.rand((H, W), generator=gen, device=dev)
                fl[..., 1] = torch.where(r < cfg.nan_flow_frac, torch.tensor(float("nan"), device=dev), fl[..., 1])
            if cfg.border_flow_frac > 0:
                band = 8
                r = torch.rand((H, W), generator=gen, device=dev)
                border = torch.zeros((H, W), dtype=torch.bool, device=dev)
                border[:band] = True
                border[-band:] = True
                border[:, :band] = True
                border[:, -band:] = True
                out = border & (r < cfg.border_flow_frac * 10)
                fl[..., 0] = torch.where(out, torch.tensor(float(W + 50), device=dev), fl[..., 0])
            flows[i - 1] = fl
        z_prev_clean = Z

    return dict(disp=disp, flows=flows, transforms=transforms, dts=dts, cam=cfg.cam, f=float(np.float32(cfg.fx)),
                T=float(np.float32(cfg.baseline)), min_disparity=cfg.min_disparity, max_disparity=cfg.max_disparity,
                cfg=cfg)


def make_stress_cloud(w: int, h: int, seed: int = 4004, dynamic_frac: float = 0.45, cluster_size: int = 2500,
                      fx: float = 1400.0):
    """Clusterer stress input (config 4): a PointXYZVelocity cloud (H,W,8) float32 numpy array with
    (a) a frame-spanning serpentine, (b) blobs straddling cluster_size, (c) thousands of specks,
    (d) checkerboard patches (anti-diagonal contact only -- must not connect at N=1),
    (e) blobs with a depth step that must split at depth_diff, (f) NaN holes."""
    rng = np.random.RandomState(seed)
    dyn = np.zeros((h, w), bool)
    z = np.full((h, w), 6.0, np.float32)
    z += (np.arange(w, dtype=np.float32) / w)[None, :] * 0.5  # gentle slope: neighbours differ by << depth_diff

    # (a) serpentine in the top 40% of the frame: horizontal bars joined alternately left / right
    top = int(0.4 * h)
    bar, gap = max(3, h // 200), max(12, h // 60)
    y, k = 4, 0
    while y + bar < top:
        dyn[y:y + bar, 8:w - 8] = True
        if y + bar + gap + bar < top:
            xs = slice(w - 8 - bar, w - 8) if k % 2 == 0 else slice(8, 8 + bar)
            dyn[y + bar:y + bar + gap, xs] = True
        y += bar + gap
        k += 1

    # (b) blobs with areas around cluster_size (some just below, some just above, some exactly equal)
    y0 = top + 10
    x = 10
    side = int(math.sqrt(cluster_size))
    row_h = 0
    areas = []
    for j in range(200):
        delta = [-3, -1, 0, 1, 3, 40, -40][j % 7]
        bw = side + rng.randint(-5, 6)
        bh = max(2, (cluster_size + delta) // bw)
        rem = (cluster_size + delta) - bw * bh
        if x + bw + 12 >= w:
            x = 10
            y0 += row_h + 12
            row_h = 0
        if y0 + bh + 2 >= int(0.8 * h):
            break
        dyn[y0:y0 + bh, x:x + bw] = True
        if rem > 0:
            dyn[y0 + bh, x:x + rem] = True
        areas.append(bw * bh + max(rem, 0))
        row_h = max(row_h, bh + 1)
        x += bw + 12

    # (c) specks + (d) checkerboards + (e) depth-stepped blobs in the bottom 20%
    yb = int(0.8 * h) + 6
    hb = h - yb - 4
    if hb > 20:
        third = w // 3
        sp = rng.rand(hb, third) < 0.08
        dyn[yb:yb + hb, 4:4 + third] |= sp
        yy, xx = np.mgrid[0:hb, 0:third]
        dyn[yb:yb + hb, third + 4:2 * third + 4] |= ((yy + xx) % 2 == 0)
        blk = dyn[yb:yb + hb, 2 * third + 8:w - 8]
        blk[:, :] = True
        zz = z[yb:yb + hb, 2 * third + 8:w - 8]
        step_cols = np.arange(zz.shape[1]) // max(8, zz.shape[1] // 6)
        zz += (step_cols.astype(np.float32) * 0.5)[None, :]  # 0.5 m steps > depth_diff 0.15: must split per column band

    # speeds: dynamic pixels 0.5..3 m/s (pseudo-random, many ties), static pixels a mix of 0 and NaN velocity
    speed = (0.5 + 0.01 * rng.randint(0, 250, size=(h, w))).astype(np.float32)
    pts = np.zeros((h, w, 8), np.float32)
    uu = (np.arange(w, dtype=np.float32) - w / 2) / fx
    vv = (np.arange(h, dtype=np.float32) - h / 2) / fx
    pts[..., 0] = uu[None, :] * z
    pts[..., 1] = vv[:, None] * z
    pts[..., 2] = z
    ang = rng.rand(h, w).astype(np.float32) * 6.2831853
    pts[..., 4] = np.where(dyn, speed * np.cos(ang), 0.0)
    pts[..., 5] = np.where(dyn, speed * 0.1, 0.0)
    pts[..., 6] = np.where(dyn, speed * np.sin(ang), 0.0)
    holes = (rng.rand(h, w) < 0.01) & ~dyn
    pts[holes, :3] = np.nan
    pts[holes, 4:7] = np.nan
    return pts, dyn

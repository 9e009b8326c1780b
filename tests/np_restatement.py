"""A SECOND, independent restatement of the hot path in numpy / pure Python, written from SURVEY.md Appendix A
(not from oracle/mod_oracle.cpp).  The oracle is pinned to the reference's own code by tests/test_reference_build.py
(oracle/_ref/libmod_ref.so); this second restatement, derived separately, additionally cross-checks the dependency
arithmetic that the reference build has to shim.  Test infrastructure only; pure-Python loops are used for the clusterer, so keep images small.

numpy float32 / float64 arithmetic is IEEE round-to-nearest-even per operation and never contracts to FMA, which is
exactly the model of Appendix A ("f32(...)" narrowing, f64 ops without contraction).
"""
import numpy as np

f32 = np.float32
f64 = np.float64


def rotation_from_quaternion(q):
    """Eigen::Quaterniond::toRotationMatrix order of operations (Appendix B.2).  q = (x, y, z, w) doubles."""
    x, y, z, w = (f64(v) for v in q)
    tx, ty, tz = f64(2.0) * x, f64(2.0) * y, f64(2.0) * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1.0 - (tyy + tzz), txy - twz, txz + twy],
                     [txy + twz, 1.0 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1.0 - (txx + tyy)]], dtype=f64)


def back_project(disp, f, T, dmin, dmax, cam):
    """A.1 for a whole image: returns (x, y, z) float32 images, NaN where the reference keeps its default."""
    fx, fy, cx, cy, Tx, Ty = (f64(v) for v in cam)
    h, w = disp.shape
    d = disp.astype(f32)
    with np.errstate(all="ignore"):
        invalid = (f32(dmax) < d) | (f32(dmin) > d) | (d == f32(0.0))   # NaN passes the range tests
        fT = f32(f32(f) * f32(T))
        z = (fT / d).astype(f32)
        rx = ((np.arange(w, dtype=f64) - cx) - Tx) / fx
        ry = ((np.arange(h, dtype=f64) - cy) - Ty) / fy
        x = (rx[None, :] * z.astype(f64)).astype(f32)
        y = (ry[:, None] * z.astype(f64)).astype(f32)
    nan = f32(np.nan)
    return np.where(invalid, nan, x), np.where(invalid, nan, y), np.where(invalid, nan, z)


def ego_transform(x, y, z, R, t):
    """A.2 Q(): NaN x passes the point through untouched, else f32(((R0*px + R1*py) + R2*pz) + t)."""
    px, py, pz = x.astype(f64), y.astype(f64), z.astype(f64)
    out = []
    with np.errstate(all="ignore"):
        for r in range(3):
            s = ((R[r, 0] * px + R[r, 1] * py) + R[r, 2] * pz) + f64(t[r])
            out.append(s.astype(f32))
    passthrough = np.isnan(x)
    return (np.where(passthrough, x, out[0]), np.where(passthrough, y, out[1]), np.where(passthrough, z, out[2]))


def static_flow(qx, qy, qz, cam):
    """A.2 S(): project3dToPixel minus the pixel, narrowed; NaN q.x -> (NaN, NaN)."""
    fx, fy, cx, cy, Tx, Ty = (f64(v) for v in cam)
    h, w = qx.shape
    X, Y, Z = qx.astype(f64), qy.astype(f64), qz.astype(f64)
    with np.errstate(all="ignore"):
        u = (fx * X + Tx) / Z + cx
        v = (fy * Y + Ty) / Z + cy
        sx = (u - np.arange(w, dtype=f64)[None, :]).astype(f32)
        sy = (v - np.arange(h, dtype=f64)[:, None]).astype(f32)
    bad = np.isnan(qx)
    nan = f32(np.nan)
    return np.where(bad, nan, sx), np.where(bad, nan, sy)


def round_half_away(a):
    """std::round(float): half away from zero, in float."""
    a = a.astype(f32)
    with np.errstate(all="ignore"):
        r = np.trunc(a)
        frac = np.abs(a - r)           # exact in float for |a| < 2^23
        r = np.where(frac >= f32(0.5), r + np.sign(a).astype(f32), r)
    return r.astype(f32)


def to_int_x86(r):
    """cvttss2si: NaN / out of range -> INT_MIN."""
    ok = (r >= f32(-2147483648.0)) & (r < f32(2147483648.0))
    return np.where(ok, np.where(ok, r, 0).astype(np.int64), np.int64(-2 ** 31))


def scene_flow(disp_now, disp_prev, flow, cam, transform, dt, dynamic_flow_diff=5, f=1.0, T=1.0, min_disparity=0.0,
               max_disparity=128.0):
    """A.3 for one frame pair.  Returns dict(points32 (H,W,8), depth, static_flow (H,W,2), cls (H,W) uint8)."""
    h, w = disp_now.shape
    nan = f32(np.nan)
    R = rotation_from_quaternion(transform[3:])
    t = [f64(v) for v in transform[:3]]

    nx, ny, nz = back_project(disp_now, f, T, min_disparity, max_disparity, cam)
    px, py, pz = back_project(disp_prev, f, T, min_disparity, max_disparity, cam)
    qx, qy, qz = ego_transform(px, py, pz, R, t)
    sx, sy = static_flow(qx, qy, qz, cam)

    out = np.full((h, w, 8), nan, f32)
    out[..., 3] = 0
    out[..., 7] = 0
    cls = np.zeros((h, w), np.uint8)

    valid_now = ~np.isnan(nx) & ~np.isinf(nx)
    out[..., 0] = np.where(valid_now, nx, nan)
    out[..., 1] = np.where(valid_now, ny, nan)
    out[..., 2] = np.where(valid_now, nz, nan)
    cls[valid_now] = 1

    Fx, Fy = flow[..., 0].astype(f32), flow[..., 1].astype(f32)
    xs = np.arange(w, dtype=f32)[None, :].repeat(h, 0)
    ys = np.arange(h, dtype=f32)[:, None].repeat(w, 1)
    with np.errstate(all="ignore"):
        pxi = to_int_x86(round_half_away(xs - Fx))
        pyi = to_int_x86(round_half_away(ys - Fy))
    chain = valid_now & ~np.isnan(Fx) & ~np.isnan(Fy)

    def usable(d, dmin, dmax):   # getDisparity range + getRightPoint extras
        with np.errstate(all="ignore"):
            return ~(f32(dmax) < d) & ~(f32(dmin) > d) & ~np.isnan(d) & ~np.isinf(d) & ~(d < 0)

    chain &= usable(disp_now.astype(f32), min_disparity, max_disparity)
    inside = (pxi >= 0) & (pxi < w) & (pyi >= 0) & (pyi < h)
    chain &= inside
    gx = np.where(inside, pxi, 0).astype(np.int64)
    gy = np.where(inside, pyi, 0).astype(np.int64)
    chain &= usable(disp_prev.astype(f32)[gy, gx], min_disparity, max_disparity)
    mx, my, mz = qx[gy, gx], qy[gy, gx], qz[gy, gx]
    chain &= ~np.isnan(mx) & ~np.isinf(mx)
    chain &= ~np.isnan(sx)

    with np.errstate(all="ignore"):
        ex, ey = (Fx - sx).astype(f32), (Fy - sy).astype(f32)
        dot = (f32(0.0) + ex * ex).astype(f32)
        dot = (dot + ey * ey).astype(f32)
        residual = np.sqrt(dot).astype(f32)
        moving = chain & (residual >= f32(dynamic_flow_diff))
        vx = ((nx - mx).astype(f32).astype(f64) / f64(dt)).astype(f32)
        vy = ((ny - my).astype(f32).astype(f64) / f64(dt)).astype(f32)
        vz = ((nz - mz).astype(f32).astype(f64) / f64(dt)).astype(f32)
    zero = f32(0.0)
    out[..., 4] = np.where(moving, vx, np.where(chain, zero, nan))
    out[..., 5] = np.where(moving, vy, np.where(chain, zero, nan))
    out[..., 6] = np.where(moving, vz, np.where(chain, zero, nan))
    cls[chain] = 2
    cls[moving] = 3
    depth = nz.copy()   # toDepthImage: z wherever getPoint3D succeeds (not gated by isValid)
    return {"points32": out, "depth": depth, "static_flow": np.stack([sx, sy], -1), "class": cls}


def norm3(vx, vy, vz):
    with np.errstate(all="ignore"):
        return np.sqrt(((vx * vx).astype(f32) + (vy * vy).astype(f32)).astype(f32) + (vz * vz).astype(f32)).astype(f32)


def cluster(points32, cluster_size=2500, depth_diff=0.15, dynamic_speed=0.3, neighbor_distance=4):
    """A.4 with pure-Python loops: the raster scan with a label-equivalence table, exactly as the reference walks it.
    Returns dict(labels (H,W) int32, objects [dict], dynamic (H,W) bool)."""
    pts = np.asarray(points32, f32)
    h, w = pts.shape[:2]
    N = int(neighbor_distance)
    speed = norm3(pts[..., 4], pts[..., 5], pts[..., 6])
    with np.errstate(all="ignore"):
        dyn = speed.astype(f64) >= f64(dynamic_speed)          # NaN -> False
    z = pts[..., 2]
    table = []

    def lookup(a):
        while table[a] != a:
            table[a] = table[table[a]]
            a = table[a]
        return a

    def link(a, b):
        ra, rb = lookup(a), lookup(b)
        if ra > rb:
            table[ra] = rb
        else:
            table[rb] = ra

    lab = np.full((h, w), -1, np.int64)
    for v in range(h):
        for u in range(w):
            if not dyn[v, u]:
                continue
            for dv in range(-N, 1):
                for du in range(-N, 1):
                    if du == 0 and dv == 0:
                        continue
                    uu, vv = u + du, v + dv
                    if uu < 0 or vv < 0 or uu >= w or vv >= h or not dyn[vv, uu]:
                        continue
                    with np.errstate(all="ignore"):
                        dz = np.abs(f32(z[v, u] - z[vv, uu]))
                    if f64(dz) > f64(depth_diff):
                        continue
                    a, b = lab[v, u], lab[vv, uu]
                    if a < 0 and b < 0:
                        table.append(len(table))
                        lab[v, u] = lab[vv, uu] = len(table) - 1
                    elif a < 0:
                        lab[v, u] = b
                    elif b < 0:
                        lab[vv, uu] = a
                    else:
                        link(a, b)
    # integrateConnectedClusters + removeSmallClusters
    roots = np.full((h, w), -1, np.int64)
    for v in range(h):
        for u in range(w):
            if lab[v, u] >= 0:
                roots[v, u] = lookup(lab[v, u])
    n_clusters = int(roots.max()) + 1 if (roots >= 0).any() else 0
    sizes = np.bincount(roots[roots >= 0], minlength=n_clusters) if n_clusters else np.zeros(0, np.int64)
    remap = np.full(n_clusters, -1, np.int64)
    nxt = 0
    for c in range(n_clusters):
        keep = False if cluster_size < 0 else (sizes[c] >= cluster_size)   # size_t < (size_t)negative removes all
        if keep:
            remap[c] = nxt
            nxt += 1
    labels = np.where(roots >= 0, remap[np.maximum(roots, 0)], -1).astype(np.int32)
    objects = []
    for c in range(nxt):
        # clusterMap2IndicesCluster: u outer, v inner
        vv, uu = np.nonzero(labels.T == c)[::-1]
        m = pts[vv, uu]
        mn, mx = m[:, :3].min(0), m[:, :3].max(0)
        bbox = (mx - mn).astype(f32)
        center = ((mn + mx).astype(f32) / f32(2.0)).astype(f32)
        sp = norm3(m[:, 4], m[:, 5], m[:, 6])
        order = np.argsort(-sp, kind="stable")
        vel = m[order[len(m) // 2], 4:7]
        if f64(norm3(vel[0:1], vel[1:2], vel[2:3])[0]) < f64(dynamic_speed):
            continue
        objects.append({"id": len(objects), "center": center, "velocity": vel.copy(), "bbox": bbox,
                        "n_points": len(m), "median_speed": sp[order[len(m) // 2]]})
    return {"labels": labels, "objects": objects, "dynamic": dyn}

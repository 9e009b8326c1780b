import sys, os, glob
import numpy as np
sys.path[:0] = ['/root/repo', '/root/repo/tests', '/root/repo/oracle']
from moving_object_detector_b200 import Context, ClustererConfig

def check(ctx, g, tag):
    cfg = ClustererConfig(int(g['params'][0]), float(g['thresholds'][0]), float(g['thresholds'][1]), int(g['params'][1]))
    pts_in = np.ascontiguousarray(g['points32'])
    got = ctx.cluster(pts_in, cfg)
    pts = pts_in.reshape(-1, 8); lab = g['labels'].ravel()
    bad = []
    if not np.array_equal(got.labels, g['labels']): bad.append('LABELS')
    for o, w in zip(got.objects, g['objects']):
        m = pts[lab == w['id']][:, 4:7]
        sp = np.sqrt((m[:, 0] * m[:, 0] + m[:, 1] * m[:, 1]) + m[:, 2] * m[:, 2]).astype(np.float32)
        srt = np.sort(sp)[::-1]
        vg = o['velocity']; ng = np.sqrt(np.float32(np.float32(vg[0]*vg[0] + vg[1]*vg[1]) + vg[2]*vg[2]))
        if ng != srt[w['n_points'] // 2] or o['n_points'] != w['n_points']:
            r = np.nonzero(srt == ng)[0]
            bad.append((int(w['id']), int(w['n_points']), int(o['n_points']), float(ng), r[:2].tolist(), bool((m == vg[None]).all(1).any())))
    print(tag, 'OK' if not bad else bad, flush=True)
    return not bad

small = np.load('/root/repo/tests/golden/cl_small_n4.npz')
for name in ('cl_stress_n1', 'cl_stress_n4', 'cl_stress_n10'):
    g = np.load(f'/root/repo/tests/golden/{name}.npz')
    g = {k: g[k] for k in g.files}
    nfail_a = nfail_b = 0
    for trial in range(12):
        ctx = Context(256, 192, max_batch=1)
        nfail_a += not check(ctx, g, f'{name} fresh-ctx first call  t{trial}')
        nfail_a += 0
        check(ctx, g, f'{name} fresh-ctx second call t{trial}')
        ctx.close()
    for trial in range(12):
        ctx = Context(256, 192, max_batch=1)
        check(ctx, {k: small[k] for k in small.files}, f'{name} warm small            t{trial}')
        nfail_b += not check(ctx, g, f'{name} after-small first call t{trial}')
        ctx.close()
    print('SUMMARY', name, 'fresh first-call failures', nfail_a, '/12; after-small failures', nfail_b, '/12', flush=True)

import sys, os, glob
import numpy as np
sys.path[:0] = ['/root/repo', '/root/repo/tests', '/root/repo/oracle']
from moving_object_detector_b200 import Context, ClustererConfig
ctx = Context(256, 192, max_batch=1)
for name in sorted(glob.glob('/root/repo/tests/golden/cl_*.npz')):
    g = np.load(name)
    cfg = ClustererConfig(int(g['params'][0]), float(g['thresholds'][0]), float(g['thresholds'][1]), int(g['params'][1]))
    for rep in range(2):
        got = ctx.cluster(g['points32'], cfg)
        pts = g['points32'].reshape(-1, 8); lab = g['labels'].ravel()
        print(os.path.basename(name), 'rep', rep, 'labels_equal', np.array_equal(got.labels, g['labels']), 'n', len(got.objects), len(g['objects']))
        for o, w in zip(got.objects, g['objects']):
            m = pts[lab == w['id']][:, 4:7]
            sp = np.sqrt((m[:, 0] * m[:, 0] + m[:, 1] * m[:, 1]) + m[:, 2] * m[:, 2]).astype(np.float32)
            srt = np.sort(sp)[::-1]
            vg = o['velocity']; ng = np.sqrt(np.float32(np.float32(vg[0]*vg[0] + vg[1]*vg[1]) + vg[2]*vg[2]))
            ranks = np.nonzero(srt == ng)[0]
            print('   id', w['id'], 'n', w['n_points'], 'k', w['n_points'] // 2, 'gpu speed', ng, 'ranks of gpu speed', ranks[:3], len(ranks),
                  'want', srt[w['n_points'] // 2], 'is_member', bool((m == vg[None]).all(1).any()))
ctx.close()

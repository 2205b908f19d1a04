"""Free-running per-iteration error of the benchmark pair (448x512, 12 iterations, pair 0) against the oracle, plus a
checksum of the final flow: run under different environment switches to see which component moves the result.
usage: python scripts/parity_probe.py [tag]"""
import hashlib, os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import cases
from oracle import raft_torch as rt, weights
import tf_raft_b200 as T
tag = sys.argv[1] if len(sys.argv) > 1 else ''
p = weights.init_params('raft', 1234)
im1, im2 = cases.images(4, 448, 512, 0, 1)
cache = '/tmp/oracle_448x512.pt'
if os.path.exists(cache):
    want = torch.load(cache)
else:
    want = rt.forward(p, im1[:1], im2[:1], 'raft', 12)
    torch.save(want, cache)
m = T.RAFT(iters=12, iters_pred=12, precision='f16x2')
m.load_params(p)
got = m([torch.from_numpy(im1[:1]).cuda(), torch.from_numpy(im2[:1]).cuda()], training=False)
per = [float((g.cpu() - o).abs().max()) for g, o in zip(got, want)]
cb = m._last['corr_block']
h = hashlib.sha1(got[-1].cpu().numpy().tobytes()).hexdigest()[:12]
hp = hashlib.sha1(cb.corr_pyramid[0].cpu().numpy().tobytes()).hexdigest()[:12]
print(f'{tag:<14} final {per[-1]:.3e} pyr0 {hp} flow {h} per-iteration ' + ' '.join(f'{e:.1e}' for e in per))

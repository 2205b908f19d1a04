"""Two forwards of the bench configuration (batch 4, 448x512, 12 iterations) for ncu launch lists."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import cases
from oracle import weights
import tf_raft_b200 as T
prec = sys.argv[1] if len(sys.argv) > 1 else 'f16x2'
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
model = T.RAFT(iters=12, iters_pred=12, precision=prec)
model.load_params(weights.init_params('raft', 1234))
im1, im2 = cases.images(4, 448, 512)
a, b = torch.from_numpy(im1).cuda(), torch.from_numpy(im2).cuda()
for _ in range(n):
    out = model([a, b], training=False, last_only=True)
torch.cuda.synchronize()
print('ok', float(out[-1].abs().max()))

"""Tiny forwards of both variants for compute-sanitizer (memcheck / racecheck)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import cases
from oracle import weights
import tf_raft_b200 as T
for variant, cls, (h, w) in (('small', T.SmallRAFT, (64, 64)), ('raft', T.RAFT, (64, 72))):
    model = cls(iters=1, iters_pred=1, precision='f16x2')
    model.load_params(weights.init_params(variant, 3, bias_scale=0.05))
    im1, im2 = cases.images(1, h, w)
    out = model([torch.from_numpy(im1).cuda(), torch.from_numpy(im2).cuda()], training=False)
    torch.cuda.synchronize()
    print(variant, 'ok', float(out[-1].abs().max()))

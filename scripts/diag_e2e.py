"""Free-running (not teacher-forced) per-iteration error trace of the CUDA loop vs the oracle at 448x512."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import cases
from oracle import raft_torch as rt, weights
import tf_raft_b200 as T

def pct(e, q):
    e = e.flatten()
    return float(e.kthvalue(max(1, int(q * e.numel()))).values)

p = weights.init_params('raft', 1234)
im1, im2 = cases.images(1, 448, 512)
preds, inter = rt.forward(p, im1, im2, 'raft', 12, return_intermediates=True)
grid = rt.coords_grid(1, 56, 64)
for precision in ('fp32', 'f16x2'):
    model = T.RAFT(iters=12, iters_pred=12, precision=precision)
    model.load_params(p)
    a, b = torch.from_numpy(im1).cuda(), torch.from_numpy(im2).cuda()
    fmap1, fmap2, net, inp = model._encode(a, b, False)
    cb = T.CorrBlock(fmap1, fmap2, 4, 4, precision=precision)
    coords1 = T.coords_grid(1, 56, 64)
    g = coords1.clone()
    print(f'--- {precision}: manual loop through the public ops')
    for i in range(12):
        corr = cb.retrieve(coords1)
        net, mask, delta = model.update_block([net, inp, corr, coords1 - g])
        coords1 = coords1 + delta
        up = model.upsample_flow(coords1 - g, mask)
        ec = (coords1.cpu() - inter['coords'][i]).abs().amax(dim=-1)
        ecorr = (corr.cpu() - inter['corr'][i]).abs()
        eu = (up.cpu() - preds[i]).abs()
        print(f'it {i:2d}: coarse coords err p50 {pct(ec, .5):.2e} p99 {pct(ec, .99):.2e} max {float(ec.max()):.2e} '
              f'n(>1e-3)={int((ec > 1e-3).sum())} | corr feat max {float(ecorr.max()):.2e} n(>0.1)={int((ecorr > 0.1).sum())} '
              f'| flow_up p50 {pct(eu, .5):.2e} p99.9 {pct(eu, .999):.2e} max {float(eu.max()):.2e}')
    full = model([a, b], training=False)
    print('forward_loop vs manual loop (last):', float((full[-1] - up).abs().max()))
    for i in (0, 3, 7, 11):
        eu = (full[i].cpu() - preds[i]).abs()
        print(f'forward_loop it {i}: flow_up p50 {pct(eu, .5):.2e} p99 {pct(eu, .99):.2e} p99.9 {pct(eu, .999):.2e} max {float(eu.max()):.2e}')

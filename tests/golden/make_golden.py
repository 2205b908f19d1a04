"""Regenerate tests/golden/*.npz from the CPU oracle.

    python tests/golden/make_golden.py

PARITY UNPINNED at model level: TensorFlow is not installable here, so these vectors come from the
restated oracle (oracle/), not from the reference itself.  If a TensorFlow install ever becomes
available, regenerate them from tf_raft and compare.  Inputs are seeded (tests/cases.py); only
outputs are stored, kept small enough to commit.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..'))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import cases  # noqa: E402
from oracle import corr_np, raft_torch as rt, weights  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
torch.set_num_threads(max(1, os.cpu_count() or 1))


def corr_lookup():
    out = {}
    for tag, (b, h, w, c, r) in {'a': (2, 8, 12, 64, 4), 'b': (1, 9, 7, 128, 3)}.items():
        f1, f2 = cases.fmaps(b, h, w, c)
        levels = 4 if tag == 'a' else 3
        cb = corr_np.CorrBlock(f1, f2, levels, r)
        for l, p in enumerate(cb.corr_pyramid):
            out[f'{tag}_pyr{l}'] = p
        for kind in ('grid', 'jitter', 'edge'):
            out[f'{tag}_lookup_{kind}'] = cb.retrieve(cases.lookup_coords(b, h, w, kind))
    np.savez_compressed(os.path.join(OUT, 'corr_lookup.npz'), **out)


def update_blocks():
    out = {}
    for variant in ('raft', 'small'):
        p = weights.init_params(variant, 1234, bias_scale=0.05)
        ops = rt.Ops(p)
        net, inp, corr, flow = cases.update_inputs(variant, 1, 8, 8)
        t = [torch.from_numpy(a).permute(0, 3, 1, 2) for a in (net, inp, corr, flow)]
        fn = rt.basic_update_block if variant == 'raft' else rt.small_update_block
        n2, mask, delta = fn(ops, *t)
        out[f'{variant}_net'] = n2.permute(0, 2, 3, 1).numpy()
        out[f'{variant}_delta'] = delta.permute(0, 2, 3, 1).numpy()
        if mask is not None:
            out[f'{variant}_mask'] = mask.permute(0, 2, 3, 1).numpy()
    np.savez_compressed(os.path.join(OUT, 'update_blocks.npz'), **out)


def models():
    out = {}
    # BASELINE.json configs[0]: SmallRAFT, 1 pair 64x128, iters=3
    p = weights.init_params('small', 1234, bias_scale=0.05, norm_jitter=0.1)
    im1, im2 = cases.images(1, 64, 128)
    preds = rt.forward(p, im1, im2, 'small', 3)
    out['small_64x128_it3'] = np.stack([q.numpy() for q in preds])
    # RAFT on the reference test's image size (tests/test_model.py:10-11), B=1 to keep the file small
    p = weights.init_params('raft', 1234, bias_scale=0.05, norm_jitter=0.1)
    im1, im2 = cases.images(1, 64, 96)
    preds = rt.forward(p, im1, im2, 'raft', 4)
    out['raft_64x96_it4'] = np.stack([q.numpy() for q in preds])
    np.savez_compressed(os.path.join(OUT, 'models.npz'), **out)


if __name__ == '__main__':
    corr_lookup()
    update_blocks()
    models()
    for f in sorted(os.listdir(OUT)):
        if f.endswith('.npz'):
            print(f, os.path.getsize(os.path.join(OUT, f)))

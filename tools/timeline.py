"""Timeline of one CTA of a tensor-core update-block layer (clock64 stamps written by conv_tc_kernel)."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import cases
from oracle import weights
import tf_raft_b200 as T
from tf_raft_b200 import _lib
layer = int(sys.argv[1]) if len(sys.argv) > 1 else 4
blk = T.BasicUpdateBlock(precision='f16x2')
blk.load_params(weights.init_params('raft', 1234), 'update_block.')
net, inp, corr, flow = [torch.from_numpy(a).cuda() for a in cases.update_inputs('raft', 4, 56, 64)]
buf = torch.zeros(2048, dtype=torch.int64, device='cuda')
for _ in range(3):
    blk([net, inp, corr, flow])
_lib.lib().raft_b200_debug_timeline(layer, _lib.ptr(buf))
blk([net, inp, corr, flow])
torch.cuda.synchronize()
_lib.lib().raft_b200_debug_timeline(-1, None)
t = buf.cpu().numpy().reshape(4, 512)
n_it = int((t[0] > 0).sum()); n_g = int((t[2][:256] > 0).sum())   # rows 2/3: [0,256) epilogue-warp stamps, [256,512) issuer / per-tile
t0 = t[0][0]
print(f'layer {layer}: {n_it} chunks, {n_g} groups; cycles relative to first slot-free')
print('chunk  slot_free  data_landed  (landed-free)')
for i in range(min(n_it, 16)):
    print(f'{i:5d} {t[0][i]-t0:10d} {t[1][i]-t0:12d} {t[1][i]-t[0][i]:10d}')
print('group  retired   drained  (drain time)  (retire gap)')
for g in range(min(n_g, 10)):
    gap = t[2][g] - t[2][g-1] if g else 0
    print(f'{g:5d} {t[2][g]-t0:9d} {t[3][g]-t0:9d} {t[3][g]-t[2][g]:10d} {gap:12d}')
d = np.diff(t[1][:n_it]); print('landed-to-landed per chunk: median', int(np.median(d)), 'mean', int(d.mean()))
print('load latency (landed - slot_free): median', int(np.median(t[1][:n_it] - t[0][:n_it])))
print('total mainloop cycles', int(t[3][n_g-1] - t0), '| epilogue cycles (last drain -> tile done)', int(t[3][511] - t[3][n_g-1]))

"""Timeline of CTA 0 of the k-th tensor-core convolution of the feature encoder (clock64 stamps of conv_tc_kernel).

usage: python tools/timeline_enc.py [k=1] [n_images=8]      (k = 0 is the stem, 1..4 the 64-channel layer-1 convs)
Rows of the stamp buffer: 0 slot-free / 1 data-landed per K chunk; 2 group retired [0,256) + issuer-owns-buffer
[256,512); 3 group drained [0,256) + tile epilogue done [256,511).
"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
from oracle import weights
import tf_raft_b200 as T
from tf_raft_b200 import _lib
k = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
enc = T.BasicEncoder(output_dim=256, norm_type='instance', drop_rate=0.0)
enc.load_params(weights.init_params('raft', 1234), 'fnet.')
g = torch.Generator().manual_seed(0)
img = (torch.rand((n, 448, 512, 3), generator=g) * 255).cuda()
buf = torch.zeros(2048, dtype=torch.int64, device='cuda')
for _ in range(2):
    enc(img, training=False, raw_image=True)
_lib.lib().raft_b200_debug_timeline(1000 + k, _lib.ptr(buf))
enc(img, training=False, raw_image=True)
torch.cuda.synchronize()
_lib.lib().raft_b200_debug_timeline(-1, None)
t = buf.cpu().numpy().reshape(4, 512)
n_it = int((t[0] > 0).sum()); n_g = int((t[2][:256] > 0).sum()); n_t = int((t[3][256:511] > 0).sum())
t0 = t[0][0]
cpt = n_it // max(n_t, 1); gpt = n_g // max(n_t, 1)
print(f'encoder conv {k}: {n_it} chunks, {n_g} groups, {n_t} tiles on CTA 0 ({cpt} chunks, {gpt} groups per tile); cycles from first slot-free')
tile_end = t[3][256:256 + n_t] - t0
print('tile  first_landed  last_retired  last_drained  epilogue_done  (tile period)  issuer_g0 issuer_g1 issuer_g2')
for i in range(min(n_t, 8)):
    c0, g0 = i * cpt, i * gpt
    per = tile_end[i] - tile_end[i - 1] if i else tile_end[0]
    iss = [int(t[2][256 + g0 + j] - t0) for j in range(min(3, gpt))]
    print(f'{i:4d} {t[1][c0]-t0:13d} {t[2][g0+gpt-1]-t0:13d} {t[3][g0+gpt-1]-t0:13d} {tile_end[i]:14d} {per:14d}  ', *iss)
if n_t > 2:
    per = np.diff(tile_end)
    epi = tile_end - (t[3][np.arange(n_t) * gpt + gpt - 1] - t0)
    drain = (t[3][:n_g] - t[2][:n_g])
    print('tile period: median', int(np.median(per)), '| epilogue (last drain -> done): median', int(np.median(epi)),
          '| drain per group: median', int(np.median(drain)))
    land = np.diff(t[1][:n_it]); print('landed-to-landed per chunk: median', int(np.median(land)), 'mean', int(land.mean()))
    print('load latency (landed - slot_free): median', int(np.median(t[1][:n_it] - t[0][:n_it])))

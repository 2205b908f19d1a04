"""Timeline of CTA 0 of corr_tc_kernel (clock64 stamps): chunk slot-free / data-landed, per tile accumulator-ready,
TMEM buffer handed back, stores done.  usage: python tools/timeline_corr.py [B=4] [h=56] [w=64]"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT)
import numpy as np, torch
import tf_raft_b200 as T
from tf_raft_b200 import _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
h = int(sys.argv[2]) if len(sys.argv) > 2 else 56
w = int(sys.argv[3]) if len(sys.argv) > 3 else 64
g = torch.Generator().manual_seed(0)
f1 = torch.randn((B, h, w, 256), generator=g).cuda()
f2 = torch.randn((B, h, w, 256), generator=g).cuda()
buf = torch.zeros(2048, dtype=torch.int64, device='cuda')
for _ in range(2):
    T.CorrBlock(f1, f2, 4, 4, precision='f16x2')
_lib.lib().raft_b200_debug_timeline(2000, _lib.ptr(buf))
T.CorrBlock(f1, f2, 4, 4, precision='f16x2')
torch.cuda.synchronize()
_lib.lib().raft_b200_debug_timeline(-1, None)
t = buf.cpu().numpy().reshape(4, 512)
n_it = int((t[0] > 0).sum()); n_t = int((t[2][:256] > 0).sum())
t0 = t[0][0]
print(f'corr_tc_kernel CTA 0: {n_it} chunks, {n_t} tiles; cycles from the first slot-free')
print('tile  first_landed  acc_ready  tmem_released  stores_done  (period)')
done = t[3][256:256 + n_t] - t0
for i in range(min(n_t, 18)):
    per = done[i] - done[i - 1] if i else done[0]
    print(f'{i:4d} {t[1][4*i]-t0:13d} {t[2][i]-t0:10d} {t[3][i]-t0:14d} {done[i]:12d} {per:9d}')
if n_t > 3:
    print('tile period: median', int(np.median(np.diff(done))), '| stores (acc ready -> done): median', int(np.median(done - (t[2][:n_t] - t0))),
          '| TMEM read-out (acc ready -> released): median', int(np.median(t[3][:n_t] - t[2][:n_t])))
    land = np.diff(t[1][:n_it]); print('landed-to-landed per chunk: median', int(np.median(land)), 'mean', int(land.mean()))
    print('load latency (landed - slot_free): median', int(np.median(t[1][:n_it] - t[0][:n_it])))

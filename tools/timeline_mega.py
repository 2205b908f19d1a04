"""Schedule trace of update_mega_kernel: per CTA and claimed item, globaltimer stamps of
claim / dependencies satisfied / ring ready / all loads issued (producer) and first group ready / last group drained /
epilogue done (epilogue warp 2).  Prints per-layer statistics and the critical chain of one pixel tile.

usage: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DRAFT_MEGA_TRACE \
           tf_raft_b200/csrc/api.cu -o tools/epi_exp/mega_trace.so          (the stamps are compiled out of the product build)
       RAFT_B200_LIB=$PWD/tools/epi_exp/mega_trace.so python tools/timeline_mega.py     (RAFT_B200_PAIR=0: single-CTA form)
"""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np, torch
import cases
from oracle import weights
import tf_raft_b200 as T
from tf_raft_b200 import _lib
blk = T.BasicUpdateBlock(precision='f16x2')
blk.load_params(weights.init_params('raft', 1234), 'update_block.')
net, inp, corr, flow = [torch.from_numpy(a).cuda() for a in cases.update_inputs('raft', 4, 56, 64)]
NC, NI = 148, 16
buf = torch.zeros(NC * NI * 8, dtype=torch.int64, device='cuda')
for _ in range(3):
    blk([net, inp, corr, flow])
_lib.lib().raft_b200_debug_timeline(3000, _lib.ptr(buf))
blk([net, inp, corr, flow])
torch.cuda.synchronize()
_lib.lib().raft_b200_debug_timeline(-1, None)
t = buf.cpu().numpy().reshape(NC, NI, 8)
valid = t[:, :, 0] > 0
t0 = t[:, :, 1][valid].min()
pair = os.environ.get('RAFT_B200_PAIR', '1') != '0'
units = 56 if pair else 112
names = ['c1', 'f1', 'c2', 'f2', 'conv', 'zr1', 'q1', 'zr2', 'q2', 'fh1', 'fh2', 'mask']
rows = []
for c in range(NC):
    for k in range(NI):
        if t[c, k, 0] > 0:
            item = int(t[c, k, 0]) - 1
            rows.append((item // units, item % units, c, *[(int(x) - t0) / 1000.0 for x in t[c, k, 1:8]]))
rows.sort()
print(f'{len(rows)} items traced; pair={pair}; times in us from the first claim')
print('layer  n   claim(min..max)   deps_ok(med)  wait_deps(med)  loads_done(med)  first_group(med)  drained(med)  epi_done(med..max)  mainloop(med) epilogue(med)')
for L in sorted(set(r[0] for r in rows)):
    rr = np.array([r[3:] for r in rows if r[0] == L])
    claim, deps, ring, loads, g0, drained, done = rr.T
    nm = names[L] if L < len(names) else str(L)
    print(f'{nm:>5} {len(rr):3d}  {claim.min():7.1f}..{claim.max():7.1f}  {np.median(deps):9.1f}  {np.median(deps - claim):9.1f}  {np.median(loads):12.1f}  {np.median(g0):12.1f}  '
          f'{np.median(drained):10.1f}  {np.median(done):8.1f}..{done.max():7.1f}  {np.median(drained - deps):9.1f}  {np.median(done - drained):9.1f}')
print('end of kernel (last epilogue done):', max(r[9] for r in rows))
# per-CTA idle accounting for the producer: time between loads_done of item k and claim/deps of item k+1

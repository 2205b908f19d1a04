"""One paired update-block application (B=2, 56x64: even tile count -> update_mega_kernel<true>) for compute-sanitizer."""
import os, sys
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import cases
from oracle import weights
import tf_raft_b200 as T
blk = T.BasicUpdateBlock(precision='f16x2')
blk.load_params(weights.init_params('raft', 1234), 'update_block.')
net, inp, corr, flow = [torch.from_numpy(a).cuda() for a in cases.update_inputs('raft', 2, 56, 64)]
out = blk([net, inp, corr, flow])
torch.cuda.synchronize()
print('pair update block ok', [float(t.abs().max()) for t in out])

#!/usr/bin/env python
"""Headline benchmark: frame-pairs/sec of the RAFT forward at 448x512, iters_pred=12, batch 4 per GPU
(BASELINE.json `metric`, configs[1]); final-flow max-abs vs the oracle reported beside it.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm on the host CPU cores

N > 1 is launched by torchrun (one rank per GPU): the batch axis shards with no data-path collective
(weak scaling: 4 pairs per GPU).  Prints ONE JSON line (rank 0).
"""
import argparse
import faulthandler
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import torch  # noqa: E402

H, W, ITERS, B_PER_GPU = 448, 512, 12, 4
METRIC = 'frame-pairs/sec at 448x512 iters=12; final-flow max-abs vs ref'
WORKLOAD = 'RAFT inference, batch 4 per GPU, 448x512, iters_pred=12 (BASELINE.json configs[1])'
N_ROTATE = 12            # distinct input batches cycled through: 12 x 2 x 11 MB = 264 MB > 126 MB L2

# Algorithmic work; px = (H/8)*(W/8) per pair
PX = (H // 8) * (W // 8)
UPDATE_MAC_PER_PX = 3_118_336                     # BasicUpdateBlock incl. mask head, update.py:128-153
MASK_MAC_PER_PX = 294_912 + 147_456               # mask[0] 3x3 128->256 + mask[2] 1x1 256->576 (update.py:137-141): only
                                                  # executed on iterations whose prediction is upsampled
CORR_FLOP_PER_PAIR = 2 * PX * PX * 256
CORR_BYTES_PER_PAIR = 4 * sum(PX * ((H // 8) >> l) * ((W // 8) >> l) for l in range(4)) + 8 * PX * 256
LOOKUP_BYTES_PER_PAIR_ITER = PX * 2904


def log(msg):
    """progress on stderr (stdout carries only the JSON line)"""
    print(f'[bench {time.strftime("%H:%M:%S")}] {msg}', file=sys.stderr, flush=True)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        d = json.load(open(path))
        return dict(hbm_gbs=float(d['hbm_gbs']), bf16_tflops=float(d.get('bf16_tflops_sustained', d['bf16_tflops'])),
                    source='measured (MEASURED_PEAKS.json)')
    except Exception:
        return dict(hbm_gbs=6650.0, bf16_tflops=1400.0, source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r for r in self.rows if len(r) >= 6 and r[0].isdigit()]
        if not rows:
            return None
        sm = sorted(int(r[0]) for r in rows)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith('active') for r in rows)]
        return dict(sm_mhz=sm[len(sm) // 2], sm_max_mhz=int(rows[0][1]), reasons=reasons, samples=len(rows))


def oracle_forward_time(n_pairs, steps, warmup):
    """Time the CPU restatement of the reference forward (oracle/raft_torch.py) on all host cores."""
    import cases
    from oracle import raft_torch as rt, weights
    cores = torch.get_num_threads()      # PyTorch's own intra-op pool size; resizing it after use can stall oneDNN
    p = weights.init_params('raft', 1234)
    im1, im2 = cases.images(n_pairs, H, W)
    for _ in range(warmup):
        rt.forward(p, im1, im2, 'raft', ITERS)
    t0 = time.perf_counter()
    for _ in range(steps):
        out = rt.forward(p, im1, im2, 'raft', ITERS)
    dt = time.perf_counter() - t0
    return n_pairs * steps / dt, dt / steps, cores, out[-1]


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return None
    pps, sec, cores, _ = oracle_forward_time(1, args.steps, args.warmup)
    sample = f'1 pair per step ({H}x{W}, {ITERS} iterations), {args.steps} steps after {args.warmup} warm-up'
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': pps, 'unit': 'pairs/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': sec * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'note': 'reference algorithm restated on PyTorch-CPU (oracle/raft_torch.py): '
                   'TensorFlow 2.3 is not installable in this image, so tf_raft itself cannot run'},
        'cpu_baseline': {'value': pps, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': pps, 'unit': 'pairs/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    return line


def run_ours(args):
    import torch.distributed as dist
    import cases
    from oracle import weights
    import tf_raft_b200 as T
    from tf_raft_b200 import _lib, parallel

    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.gpus > 1 and world == 1:
        raise SystemExit('--gpus N > 1 must be launched with torchrun (one rank per GPU)')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)

    precision = args.precision
    model = T.RAFT(iters=ITERS, iters_pred=ITERS, precision=precision, device=device, use_graph=not args.no_graph)
    params = weights.init_params('raft', 1234)          # seeded Glorot-uniform (keras defaults), SURVEY 8(d)
    model.load_params(params)

    # synthetic inputs: rotating set of distinct batches, each rank its own seeds (weak scaling)
    host = [tuple(torch.from_numpy(a).pin_memory() for a in
                  cases.images(B_PER_GPU, H, W, 1000 * rank + 2 * i, 1000 * rank + 2 * i + 1)) for i in range(N_ROTATE)]
    dev_in = [(a.to(device), b.to(device)) for a, b in host]

    def step_resident(i):
        a, b = dev_in[i % N_ROTATE]
        return model([a, b], training=False, last_only=True)[-1]

    def step_e2e(i):
        a, b = host[i % N_ROTATE]
        out = model.predict_step((a.to(device, non_blocking=True), b.to(device, non_blocking=True)))
        return out.to('cpu', non_blocking=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        dev_s = e0.elapsed_time(e1) / 1e3
        barrier()
        return parallel.max_over_ranks(dev_s, device), parallel.max_over_ranks(wall, device)

    log('model and inputs ready; warm-up')
    for i in range(args.warmup):
        step_resident(i)
    torch.cuda.synchronize()
    log('warm-up done; timing device-resident steps')
    # kernels per step: counted on one directly-launched forward (a CUDA-graph replay launches the same kernel
    # nodes without passing through the host-side counter)
    _lib.launch_count_reset()
    model._forward(dev_in[0][0], dev_in[0][1], False, True)
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    dev_s, _ = timed(step_resident, args.steps)
    clocks = sampler.stop() if sampler else None
    launches = launches_per_step * args.steps
    value = world * B_PER_GPU * args.steps / dev_s

    log(f'resident: {value:.1f} pairs/s; timing end-to-end steps')
    for i in range(min(args.warmup, 2)):
        step_e2e(i)
    if args.pipeline:      # experiment: same copies per step, overlapped with the compute of the neighbouring steps
        def run_pipelined(steps):
            barrier()
            t0 = time.perf_counter()
            n = 0
            for out in parallel.predict_stream(lambda a, b: model.predict_step((a, b)),
                                               (host[i % N_ROTATE] for i in range(steps)), device):
                n += out.shape[0]
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            assert n == steps * B_PER_GPU
            barrier()
            return parallel.max_over_ranks(wall, device)
        run_pipelined(2)
        e2e_wall = run_pipelined(args.steps)
    else:
        _, e2e_wall = timed(step_e2e, args.steps)
    e2e_value = world * B_PER_GPU * args.steps / e2e_wall
    h2d = 2 * B_PER_GPU * H * W * 3 * 4
    d2h = B_PER_GPU * H * W * 2 * 4

    line = None
    if rank == 0:
        peaks = load_peaks()
        log(f'e2e: {e2e_value:.1f} pairs/s; kernel-level timings')
        # --- kernel-level timing for the roofline objects (rank 0, CUDA events on the launching stream) ---
        a, b = dev_in[0]
        fmap1, fmap2, net, inp = model._encode(a, b, False)
        h, w = H // 8, W // 8

        def ev_time(fn, reps=5):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / 1e3 / reps

        cb = T.CorrBlock(fmap1, fmap2, 4, 4, precision=precision)
        pyr_ptrs = _lib.ptr_array(cb.corr_pyramid)

        def build_corr():           # same buffers every time: times the kernels, not the allocator
            _lib.check(_lib.lib().raft_b200_corr_pyramid_build(
                _lib.ptr(fmap1), _lib.ptr(fmap2), B_PER_GPU, h, w, 256, 4, pyr_ptrs, _lib.ptr(cb._ws), cb._ws.numel(),
                cb.precision, _lib.stream()), 'corr_pyramid_build')
        t_corr = ev_time(build_corr)
        coords = T.coords_grid(B_PER_GPU, h, w, device) + 0.37
        lk_out = torch.empty((B_PER_GPU, h, w, 324), device=device)

        def lookup():
            _lib.check(_lib.lib().raft_b200_corr_lookup(pyr_ptrs, _lib.ptr(coords), B_PER_GPU, h, w, 4, 4, _lib.ptr(lk_out),
                                                        324, _lib.stream()), 'corr_lookup')
        t_lookup = ev_time(lookup)
        preds = [None] * (ITERS - 1) + [torch.empty((B_PER_GPU, H, W, 2), device=device)]

        def loop():
            c1 = T.coords_grid(B_PER_GPU, h, w, device)
            model._loop(cb, net.clone(), inp, c1, preds, B_PER_GPU, h, w)
        t_loop = ev_time(loop, reps=3)
        n_upsampled = sum(1 for q in preds if q is not None)      # the timed loop upsamples the last prediction only
        upd_flops = 2.0 * ((UPDATE_MAC_PER_PX - MASK_MAC_PER_PX) * ITERS + MASK_MAC_PER_PX * n_upsampled) * PX * B_PER_GPU
        ach_tflops = upd_flops / max(t_loop - ITERS * t_lookup, 1e-9) / 1e12
        corr_bytes = B_PER_GPU * (CORR_BYTES_PER_PAIR + ITERS * LOOKUP_BYTES_PER_PAIR_ITER)
        ach_gbs = corr_bytes / (t_corr + ITERS * t_lookup) / 1e9

        if args.quick or world > 1:     # the CPU legs (oracle parity, CPU baseline) belong to the N=1 line only
            parity, max_abs, cpu_pps, cores = {'skipped': '--quick' if args.quick else 'reported at N=1'}, None, None, 0
        else:
            log('parity check of the timed configuration against the oracle')
            # --- parity of the timed configuration against the oracle (one pair of batch 0) ---
            from oracle import raft_torch as rt
            im1, im2 = cases.images(B_PER_GPU, H, W, 0, 1)
            want = rt.forward(params, im1[:1], im2[:1], 'raft', ITERS)
            check_model = T.RAFT(iters=ITERS, iters_pred=ITERS, precision=precision, device=device)
            check_model.load_params(params)
            got = check_model([dev_in[0][0][:1], dev_in[0][1][:1]], training=False)
            per_iter = [float((g.cpu() - o).abs().max()) for g, o in zip(got, want)]
            final_err = (got[-1].cpu() - want[-1]).abs()
            max_abs = per_iter[-1]
            within = 0
            while within < ITERS and per_iter[within] <= 1e-3:
                within += 1
            parity = {'max_abs': max_abs, 'median_abs': float(final_err.flatten().median()),
                      'frac_px_within_1e-3': float((final_err <= 1e-3).float().mean()),
                      'iterations_within_1e-3': within, 'max_abs_per_iteration': per_iter,
                      'flow_magnitude_px': float(want[-1].abs().max()),
                      'note': 'free-running vs the CPU oracle on pair 0; the reference sampler is discontinuous at integer / '
                              'border coordinates (corr.py:45-60), so once one tap crosses, that pixel legitimately diverges '
                              '(DESIGN.md section 4); teacher-forced stage parity is in tests/test_gpu_stages.py'}

            log(f'max-abs {max_abs:.2e}; CPU baseline')
            # --- CPU baseline: the restated reference on the host cores, bounded sample ---
            cpu_pps, cpu_sec, cores, _ = oracle_forward_time(1, 3, 1)

        line = {
            'metric': METRIC, 'value': value, 'unit': 'pairs/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dev_s / args.steps * 1e3, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'global_batch': world * B_PER_GPU, 'parallelism': f'dp{world}',
                       'arithmetic': {'f16x2': 'tcgen05 fp16 hi/lo split, 3 passes, fp32 accumulate (fp32-grade)',
                                      'fp32': 'CUDA-core FFMA'}[precision],
                       'encoders': {'f16x2': 'native: the same tcgen05 implicit-GEMM kernel (stride-2 TMA boxes, fused norm affine)',
                                    'fp32': 'cuDNN IEEE fp32 via PyTorch'}[precision],
                       'cuda_graph': not args.no_graph, 'last_only': True,
                       'l2': f'inputs rotate over {N_ROTATE} distinct batches (264 MB) and every step rewrites the '
                             '273 MB correlation pyramid: working set > 126 MB L2'},
            'final_flow_max_abs_vs_oracle': max_abs,
            'parity': parity,
            'e2e': {'value': e2e_value, 'unit': 'pairs/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'api': 'parallel.predict_stream (pipelined)' if args.pipeline else 'RAFT.predict_step, one synchronous call per step'},
            'gpu_launches': int(launches),
            'clocks': clocks,
            'roofline': {'bound': 'tensor', 'kernel': 'conv_tc_kernel (update-block implicit GEMMs, 12 iterations)',
                         'achieved': ach_tflops, 'peak': peaks['bf16_tflops'], 'unit': 'TFLOP/s',
                         'frac': ach_tflops / peaks['bf16_tflops'],
                         'traffic': {'bytes_per_launch': 31.5e6, 'launch': 'GRU z||r layer (14336 px x 256 x 1920)',
                                     'source': 'profiles/r01_ncu_conv_tc.txt (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full; '
                                               'caches flushed by ncu -- operands are L2-resident in the real run)',
                                     'algorithmic_bytes_per_launch': 14336 * (384 * 5 * 4 + 256 * 4) + 256 * 1920 * 4},
                         'peak_source': peaks['source'],
                         'note': 'achieved = algorithmic fp32 FLOPs the timed loop executes (2*2,675,968 MAC/px on every '
                                 'iteration + 2*442,368 MAC/px of mask head on the one upsampled iteration) / CUDA-event '
                                 'time of the loop minus lookups; the kernel executes 3 fp16 MMA passes per FLOP, so the '
                                 'tensor pipe is 3x busier than `frac`',
                         'executed_frac': 3 * ach_tflops / peaks['bf16_tflops']},
            'roofline_corr_lookup': {'bound': 'hbm', 'kernel': 'correlation pyramid build + 12 lookups',
                                     'achieved': ach_gbs, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                                     'frac': ach_gbs / peaks['hbm_gbs'],
                                     'traffic': {'lookup_bytes_per_launch': 49.2e6, 'lookup_algorithmic_bytes': B_PER_GPU * LOOKUP_BYTES_PER_PAIR_ITER,
                                                 'source': 'profiles/r01_ncu_lookup.txt'},
                                     'ms': {'pyramid_build': t_corr * 1e3, 'lookup': t_lookup * 1e3},
                                     'peak_source': peaks['source']},
            'cpu_baseline': {'value': cpu_pps, 'unit': 'pairs/s', 'cores': cores, 'kind': 'port',
                             'sample': f'1 pair ({H}x{W}, {ITERS} iterations) x 3 steps after 1 warm-up, '
                                       'oracle/raft_torch.py on all host cores'},
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--no-graph', action='store_true', help='launch kernels directly instead of replaying a CUDA graph')
    ap.add_argument('--pipeline', action='store_true',
                    help='end-to-end leg through parallel.predict_stream (upload of pair i+1 / read-back of pair i overlap the '
                         'compute; experiment, not yet run on hardware) instead of one synchronous predict_step per step')
    ap.add_argument('--quick', action='store_true', help='timing only: skip the parity and CPU-baseline legs (A/B runs)')
    ap.add_argument('--precision', default=os.environ.get('RAFT_B200_PRECISION', 'f16x2'), choices=['f16x2', 'fp32'])
    args = ap.parse_args()
    # stdout must carry exactly one JSON line: libraries (NCCL prints its version banner there) get stderr instead
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    # torchrun exports OMP_NUM_THREADS=1; the CPU baseline / reference arm should use the host's cores.  Size the
    # intra-op pool once, before its first use (resizing a pool that is already in use stalled oneDNN for minutes).
    torch.set_num_threads(max(1, min(host_cores(), 64)))
    faulthandler.enable()
    faulthandler.dump_traceback_later(420, exit=False)     # a hang leaves stack traces on stderr
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else max(args.warmup, 1)
    line = run_reference(args) if args.impl == 'reference' else run_ours(args)
    if line is not None:
        os.write(real_stdout, (json.dumps(line) + '\n').encode())


if __name__ == '__main__':
    main()

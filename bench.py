"""Benchmark of the two hot paths on synthetic data of BASELINE.json's shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (N GPUs, weak scaling): config C2 -- CAE 128x128x1, 12 latents, 256 frames per
GPU, one ``AE.loss(data, accumulate_grad=True)`` call per step (forward + fused MSE + backward over
all reference chunks, plus the gradient all-reduce when N > 1), frames resident in HBM.
The same JSON line carries an ``arhmm`` object for config C4 (K=16, lag 2, D=12, 2048 trials x 1000
steps per GPU, one E-step per step).

``--impl reference`` times the CPU oracle port of the reference path (oracle/) on the host cores.
"""

import argparse
import datetime
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic work
CAE_C2_TRAIN_GFLOP_PER_FRAME = 2.078
ARHMM_BYTES_PER_TIMESTEP = 113.0
# dram__bytes_read + write per launch (profiles/r01_g_ncu_full.txt)
IGEMM128_KERNEL_TRAFFIC_BYTES = 72.1e6      # encoder conv2 forward: 67 MB input image read once
HALO_KERNEL_TRAFFIC_BYTES = 144.2e6
CAE_BATCH_PER_GPU = 256
ARHMM_TRIALS_PER_GPU, ARHMM_T, ARHMM_K, ARHMM_D, ARHMM_LAGS = 2048, 1000, 16, 12, 2


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region: the sampler runs from
    before the warm-up, every row carries nvidia-smi's own timestamp, and only rows inside
    [mark_start, mark_stop] are summarised."""

    Q = ('timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '20'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def mark_start(self):
        self.t0 = datetime.datetime.now()

    def mark_stop(self):
        self.t1 = datetime.datetime.now()

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        parsed = []
        for r in self.rows:
            f = [v.strip() for v in r.split(',')]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], '%Y/%m/%d %H:%M:%S.%f')
                parsed.append((ts, float(f[1]), float(f[2]), f[4:8]))
            except ValueError:
                continue
        inside = [q for q in parsed if self.t0 is not None and self.t0 <= q[0] <= self.t1]
        # nvidia-smi's sampling period is coarser than a short timed region: fall back to the rows
        # closest to it (the GPU is under the same load during warm-up and the follow-up loops)
        where = 'timed region'
        if len(inside) < 2:
            inside = [q for q in parsed if self.t0 is not None and
                      abs((q[0] - self.t0).total_seconds()) < 1.0 + (self.t1 - self.t0).total_seconds()]
            where = 'timed region +-1 s (region shorter than the sampling period)'
        sm = [q[1] for q in inside]
        mx = [q[2] for q in inside]
        reasons = set()
        for q in inside:
            for n, v in zip(names, q[3]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': float(np.max(mx)) if mx else None,
                'samples': len(sm), 'window': where, 'reasons': sorted(reasons)}


def make_cae(device):
    import copy
    from behavenet_b200.models import AE
    from oracle import cae_oracle as co          # only for the seeded synthetic parameters
    hp = co.make_hparams(1, 128, 128, 12)
    model = AE(copy.deepcopy(hp))
    model.load_state_dict(co.init_state_dict(hp, seed=0))
    model.to(device)
    return model, hp


def timed(fn, steps, warmup, flush=None):
    """CUDA-event timing of ``steps`` calls after ``warmup`` calls; returns list of ms per step."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(steps):
        if flush is not None:
            flush()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    return ms


def time_layer_kernel(model, device, side, layer, op, name, traffic_bytes, iters=20):
    """One layer kernel timed alone through bn_cae_layer_op (CUDA events on the launch stream, after
    warm-up): side 0/1 = encoder/decoder layer, op 0/1 = forward / backward-data."""
    from behavenet_b200 import _lib
    drv, rt = model._driver, model._rt
    hp = model.hparams
    n = CAE_BATCH_PER_GPU
    params = model._kernel_params()
    packed = drv.packed(rt, params, device)
    ws = drv.workspace(rt, n, device)
    if side == 0:
        cb, hb, wb = ((hp['ae_input_dim'][0], hp['ae_input_dim'][1], hp['ae_input_dim'][2]) if layer == 0 else
                      (hp['ae_encoding_n_channels'][layer - 1], hp['ae_encoding_y_dim'][layer - 1],
                       hp['ae_encoding_x_dim'][layer - 1]))
        cs, hs, wsm = (hp['ae_encoding_n_channels'][layer], hp['ae_encoding_y_dim'][layer],
                       hp['ae_encoding_x_dim'][layer])
    else:
        c0, h0, w0 = hp['ae_decoding_starting_dim']
        cs, hs, wsm = ((c0, h0, w0) if layer == 0 else
                       (hp['ae_decoding_n_channels'][layer - 1], hp['ae_decoding_y_dim'][layer - 1],
                        hp['ae_decoding_x_dim'][layer - 1]))
        cb, hb, wb = (hp['ae_decoding_n_channels'][layer], hp['ae_decoding_y_dim'][layer],
                      hp['ae_decoding_x_dim'][layer])
    fprop_form = (side == 0 and op == 0) or (side == 1 and op == 1)
    src = torch.rand((n, hb, wb, cb) if fprop_form else (n, hs, wsm, cs), device=device)
    out = torch.empty((n, hs, wsm, cs) if fprop_form else (n, hb, wb, cb), device=device)
    lib = _lib.lib()

    def run():
        _lib.check(lib.bn_cae_layer_op(drv.plan(device), side, layer, op, n, src.data_ptr(), None, out.data_ptr(),
                                       drv.table(params), packed.data_ptr(), ws.data_ptr(),
                                       _lib.stream_ptr()), 'bn_cae_layer_op')
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        run()
    e1.record()
    e1.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / iters
    gflop = 2.0 * n * hs * wsm * cs * 25 * cb * 1e-9
    return {'kernel': name, 'us': us, 'gflop': gflop, 'tflops': gflop / us * 1e3, 'launches': iters,
            'traffic_bytes': traffic_bytes}


def time_dominant_kernel(model, device):
    """The kernel with the largest share of the step in the committed launch list
    (profiles/r01_g_launches_summary.txt: igemm_tma_kernel<128,3>, 8 launches, 10.1 %): its encoder
    conv2 forward instance (64 -> 128 channels, 32x32 -> 16x16, k5 s2, 256 frames)."""
    return time_layer_kernel(model, device, 0, 2, 0,
                             'igemm_tma_kernel<128,3> (encoder conv2 forward, M=65536 N=128 K=1600)',
                             IGEMM128_KERNEL_TRAFFIC_BYTES)


def measure_cublas_tf32(device):
    """cuBLAS TF32 GEMM throughput on this GPU, for context next to the roofline denominator."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        a = torch.randn(8192, 8192, device=device)
        b = torch.randn(8192, 8192, device=device)
        for _ in range(3):
            a @ b
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            a @ b
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return 2 * 8192 ** 3 / (best * 1e-3) * 1e-12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def dist_max(value, device):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        t = torch.tensor([value], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return value


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def run_ours(args):
    from behavenet_b200 import _lib, parallel
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        parallel.init('nccl')
    peaks, peak_src = measured_peaks()
    lib = _lib.lib()
    lib.bn_set_tensor_core_mode(1)

    # ---------------- CAE (C2) : weak scaling, 256 frames per GPU
    model, hp = make_cae(device)
    model.data_parallel = world > 1
    B = CAE_BATCH_PER_GPU * world
    g = torch.Generator().manual_seed(0)
    x_host = torch.rand(B, 1, 128, 128, generator=g).pin_memory()
    x_dev = x_host.to(device)
    data = {'images': x_dev[None]}
    opt = torch.optim.Adam(model.get_parameters(), lr=1e-4, amsgrad=True)
    l2_flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)   # > 126 MB L2

    def flush():
        l2_flush.zero_()

    def step():
        opt.zero_grad()
        # a training step follows an optimizer step: the GEMM-ordered weight copies are stale and are
        # re-packed inside the timed call (the optimizer itself is timed separately)
        model.invalidate_packed()
        model.loss(data, accumulate_grad=True)

    lo, hi = parallel.shard_range(B, world, rank)
    x_shard_host = x_host[lo:hi].clone().pin_memory()       # this rank's frames only

    # End to end as a training loop would run it: every step's frames come from pinned host memory,
    # the copy of step i+1 is issued on a side stream before the loss of step i (two device input
    # buffers), and every loss() call ends with its device->host read of the per-chunk sums.
    copy_stream = torch.cuda.Stream(device=device)
    in_bufs = [torch.empty_like(x_dev[lo:hi]) for _ in range(2)]
    in_ready = [torch.cuda.Event() for _ in range(2)]

    def prefetch(i):
        with torch.cuda.stream(copy_stream):
            in_bufs[i & 1].copy_(x_shard_host, non_blocking=True)
            in_ready[i & 1].record(copy_stream)

    def run_e2e(n_steps):
        prefetch(0)
        for i in range(n_steps):
            torch.cuda.current_stream().wait_event(in_ready[i & 1])
            prefetch(i + 1)          # buffer (i+1)&1 was last read by step i-1, which has returned
            opt.zero_grad()
            model.invalidate_packed()
            model.loss({'images': in_bufs[i & 1][None], 'shard': (lo, B)}, accumulate_grad=True)
        copy_stream.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    barrier()
    launches0 = _lib.launch_count()
    # timed region: exactly K steps, barrier + synchronize on both sides
    torch.cuda.synchronize()
    sampler.mark_start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    sampler.mark_stop()
    barrier()
    total_ms = dist_max(e0.elapsed_time(e1), device)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    cae_value = B / (ms_per_step * 1e-3)
    # per-step distribution with an L2 flush between steps (inputs are 16.8 MB/GPU < L2)
    ms_flushed = timed(step, max(3, min(args.steps, 10)), 0, flush)
    opt_ms = timed(lambda: opt.step(), 3, 1)
    e2e_steps = max(3, min(args.steps, 20))
    run_e2e(2)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    run_e2e(e2e_steps)
    torch.cuda.synchronize()
    e2e_ms = dist_max((time.perf_counter() - t0) * 1e3 / e2e_steps, device)

    tf = cae_value * CAE_C2_TRAIN_GFLOP_PER_FRAME * 1e-3 / world         # TFLOP/s per GPU
    # TF32 dense peak: MEASURED_PEAKS.json has no TF32 entry; the hardware ratio to bf16 is 1/2, so
    # the denominator is half of the measured bf16 burst (kernel timed alone) / sustained (whole step)
    tf32_peak_burst = peaks['bf16_tflops'] / 2.0
    tf32_peak_sust = peaks.get('bf16_tflops_sustained', peaks['bf16_tflops']) / 2.0
    dom = time_dominant_kernel(model, device)
    dom2 = time_layer_kernel(model, device, 0, 1, 0,
                             'igemm_tma_kernel<64,4> (encoder conv1 forward, M=262144 N=64 K=800)', 51.7e6)
    dom3 = time_layer_kernel(model, device, 1, 3, 0,
                             'dgrad_halo_kernel<32,3> (decoder convtranspose3 forward, 4 x [M=262144 N=32 K<=576])',
                             HALO_KERNEL_TRAFFIC_BYTES)
    cublas_tf32 = measure_cublas_tf32(device) if rank == 0 else None

    # ---------------- ARHMM (C4): weak scaling, 2048 trials per GPU
    from behavenet_b200.ssm import HMM
    from oracle import arhmm_oracle as ao         # synthetic parameters / data generator only
    p = ao.synth_params(ARHMM_K, ARHMM_D, ARHMM_LAGS, seed=0)
    hmm = HMM(ARHMM_K, ARHMM_D, observations='ar', observation_kwargs={'lags': ARHMM_LAGS})
    hmm.init_state_distn.log_pi0, hmm.transitions.log_Ps = p.log_pi0, p.log_Ps
    hmm.observations.As, hmm.observations.bs, hmm.observations.Sigmas = p.As, p.bs, p.Sigmas
    X = ao.sample_batch(p, ARHMM_TRIALS_PER_GPU, ARHMM_T, seed=rank)
    trials = [X[i] for i in range(X.shape[0])]
    st = hmm._stage(trials)
    n_ts = ARHMM_TRIALS_PER_GPU * ARHMM_T

    def estep():
        Ez, Ezz, logZ = hmm._run_estep(st, True)
        if world > 1:
            stats = torch.cat([Ezz.sum(0).double().reshape(-1), logZ.sum().reshape(1)])
            parallel.all_reduce_sum(stats)

    for _ in range(max(3, args.warmup)):
        estep()
    torch.cuda.synchronize()
    barrier()
    h0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        estep()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    hmm_ms = dist_max(e0.elapsed_time(e1), device) / args.steps
    hmm_launches = _lib.launch_count() - h0
    hmm_value = n_ts * world / (hmm_ms * 1e-3)

    def estep_e2e():
        hmm.clear_cache()
        s2 = hmm._stage(trials)                       # host -> device copy of the latents
        Ez, Ezz, logZ = hmm._run_estep(s2, True)
        return float(logZ.sum().item()), Ezz.sum(0).cpu()
    hmm_e2e = float(np.median(timed(estep_e2e, 3, 1)))

    # ---------------- PS-VAE (C3): 128x128x2, 16 latents, 4 labels, 512 frames per GPU (weak scaling)
    psvae = None
    try:
        psvae = bench_psvae(device, world, rank, args)
    except Exception as exc:                     # the headline line must not depend on this extra leg
        psvae = {'error': repr(exc)[:200]}

    if rank != 0:
        return
    cpu = cpu_baselines() if world == 1 else None
    hbm_achieved = hmm_value / world * ARHMM_BYTES_PER_TIMESTEP / 1e9
    line = {
        'metric': 'CAE train frames/sec (C2: 128x128x1, 12 latents, fwd+loss+bwd)',
        'value': cae_value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'tf32 (fp32 accumulate)' if lib.bn_get_tensor_core_mode() else 'f32',
        'data': 'synthetic',
        'config': {'workload': 'C2: CAE 128x128x1, 12 latents, batch=256 per GPU, AE.loss fwd+bwd, '
                               'default 5-layer arch', 'global_batch': B,
                   'l2': 'inputs 16.8 MB/GPU < L2; ms_per_step_l2_flushed reports the flushed timing',
                   'parallelism': 'dp%d' % world},
        'ms_per_step_l2_flushed': float(np.median(ms_flushed)),
        'optimizer_step_ms': float(np.median(opt_ms)),
        'gpu_launches': int(launches),
        'clocks': clocks,
        'e2e': {'value': B / (e2e_ms * 1e-3), 'unit': 'frames/s',
                'h2d_bytes_per_step': int(x_host.numel() * 4), 'd2h_bytes_per_step': 16 * world,
                'steps': e2e_steps,
                'note': 'AE.loss on frames copied from pinned host memory every step (the copy of step i+1 '
                        'overlaps the compute of step i on a side stream; one extra prefetch is inside the '
                        'timed region) + the per-step loss read-back; host wall clock, max over ranks'},
        'roofline': {'bound': 'tensor', 'achieved': dom['tflops'], 'peak': tf32_peak_burst,
                     'unit': 'TFLOP/s', 'frac': dom['tflops'] / tf32_peak_burst,
                     'traffic': dom['traffic_bytes'],
                     'kernel': dom['kernel'], 'kernel_us': dom['us'], 'kernel_gflop': dom['gflop'],
                     'note': 'dominant kernel timed alone with CUDA events on the launch stream (%d '
                             'launches after warm-up); peak = TF32 dense = half of the %s bf16 burst '
                             'peak (MEASURED_PEAKS.json has no TF32 entry); traffic = dram read+write '
                             'bytes per launch from profiles/r01_g_ncu_full.txt'
                             % (dom['launches'], peak_src)},
        'step_roofline': {'bound': 'tensor', 'achieved': tf, 'peak': tf32_peak_sust, 'unit': 'TFLOP/s',
                          'frac': tf / tf32_peak_sust,
                          'note': 'whole step: %.3f GFLOP/frame algorithmic / step time, per GPU; peak = '
                                  'half of the %s bf16 sustained peak' % (CAE_C2_TRAIN_GFLOP_PER_FRAME, peak_src)},
        'kernel_rooflines': [{'kernel': d['kernel'], 'us': d['us'], 'tflops': d['tflops'],
                              'frac': d['tflops'] / tf32_peak_burst, 'traffic': d['traffic_bytes']}
                             for d in (dom, dom2, dom3)],
        'cublas_tf32_tflops_here': cublas_tf32,
        'arhmm': {
            'metric': 'ARHMM E-step timesteps/sec (C4: K=16, lag 2, D=12, 2048 trials x 1000 per GPU)',
            'value': hmm_value, 'unit': 'timesteps/s', 'ms_per_step': hmm_ms,
            'gpu_launches': int(hmm_launches), 'dtype': 'f32 (scaled messages), f64 log-normaliser',
            'e2e': {'value': n_ts / (hmm_e2e * 1e-3), 'unit': 'timesteps/s',
                    'h2d_bytes_per_step': int(n_ts * ARHMM_D * 4),
                    'd2h_bytes_per_step': int(8 + ARHMM_K * ARHMM_K * 4)},
            'roofline': {'bound': 'hbm', 'achieved': hbm_achieved, 'peak': peaks['hbm_gbs'],
                         'unit': 'GB/s', 'frac': hbm_achieved / peaks['hbm_gbs'], 'traffic': 588.4e6,
                         'note': '113 B/timestep algorithmic (48 B latents in + 64 B posteriors out '
                                 '+ per-trial outputs), whole E-step time (emission + scan kernels); peak %s; '
                                 'traffic = dram read+write of the dominant kernel (scan2_kernel, 0.31 of the 0.57 ms; '
                                 'emission_tc_kernel adds 187 MB), profiles/r01_g_ncu_full.txt; the binding limits '
                                 'are the T-step serial chain and the 3-pass emission GEMM, not HBM' % peak_src},
        },
    }
    line['psvae'] = psvae
    if cpu is not None:
        line['cpu_baseline'] = cpu['cae']
        line['arhmm']['cpu_baseline'] = cpu['arhmm']
    _OUT.write(json.dumps(line) + '\n')
    _OUT.flush()


PSVAE_GFLOP_PER_FRAME = 2.111          # train step, 128x128x2, 16 latents
PSVAE_BATCH_PER_GPU = 512


def bench_psvae(device, world, rank, args):
    """Config C3: one PSVAE.loss(accumulate_grad=True) per step on 512 frames per GPU (chunks of 200;
    whole chunks are sharded over ranks, so every rank gets its own 512-frame batch slice)."""
    import copy
    from behavenet_b200.models import PSVAE
    from oracle import cae_oracle as co          # seeded synthetic parameters only
    np.random.seed(0)
    hp = co.make_hparams(2, 128, 128, 16, 'ps-vae', 4)
    model = PSVAE(copy.deepcopy(hp))
    model.load_state_dict(co.init_state_dict(hp, seed=0))
    model.to(device)
    model.curr_epoch = 1
    # 600 frames per rank = three whole reference chunks of 200 per rank would change the config;
    # keep 512 per GPU and let the chunk sharding split [200, 200, 112] x world over the ranks
    B = PSVAE_BATCH_PER_GPU * world
    model.data_parallel = world > 1
    g = torch.Generator().manual_seed(1)
    x = torch.rand(B, 2, 128, 128, generator=g).to(device)
    y = torch.randn(B, 4, generator=g).to(device)
    eps = torch.randn(B, 16, generator=g).to(device)
    data = {'images': x[None], 'labels': y[None]}

    def step():
        model.zero_grad()
        model.invalidate_packed()
        model.loss(data, accumulate_grad=True, eps=eps)

    steps = max(3, min(args.steps, 10))
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = dist_max(e0.elapsed_time(e1), device) / steps
    value = B / (ms * 1e-3)
    return {'metric': 'PS-VAE train frames/sec (C3: 128x128x2, 16 latents, 4 labels, fwd+loss+bwd)',
            'value': value, 'unit': 'frames/s', 'ms_per_step': ms, 'steps': steps, 'global_batch': B,
            'tflops_per_gpu': value * PSVAE_GFLOP_PER_FRAME * 1e-3 / world,
            'note': 'includes sklearn r2_score on the host and the per-chunk latent-block launches, as the '
                    'reference loss() does; frames resident in HBM'}


def import_reference():
    """The UNMODIFIED reference package from baseline/_ref (installed by baseline/install_ref.sh; git-ignored,
    travels with gpurun).  ``commentjson`` -- imported at the top of ae_model_architecture_generator.py
    but only used to read arch json files -- is absent from this image and is stubbed.  Returns the
    ``behavenet.models`` module or None when the install is missing."""
    import types
    ref = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref, 'behavenet', 'models')):
        return None
    sys.modules.setdefault('commentjson', types.ModuleType('commentjson'))
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import behavenet.models as ref_models
    return ref_models


def reference_ae(device):
    """behavenet.models.AE (reference aes.py:616-773) with the seeded parameters of the GPU arm."""
    import copy
    from oracle import cae_oracle as co          # seeded synthetic parameters only
    ref_models = import_reference()
    if ref_models is None:
        return None, None
    hp = co.make_hparams(1, 128, 128, 12)
    hp_ref = copy.deepcopy(hp)
    hp_ref['device'] = str(device)
    model = ref_models.AE(hp_ref)
    model.load_state_dict(co.init_state_dict(hp, seed=0))
    model.to(device)
    return model, hp


def cpu_cae_rate(batch, steps, warmup, threads):
    """frames/s of the reference's own AE.loss(data, accumulate_grad=True) on the host cores (the oracle
    port when baseline/_ref is missing) -> (frames/s, seconds per step, kind)."""
    torch.set_num_threads(threads)
    x = torch.rand(batch, 1, 128, 128, generator=torch.Generator().manual_seed(0))
    model, hp = reference_ae('cpu')
    if model is not None:
        kind = 'reference'

        def step():
            model.zero_grad()
            model.loss({'images': x[None]}, accumulate_grad=True)
    else:
        from oracle import cae_oracle as co
        kind = 'port'
        hp = co.make_hparams(1, 128, 128, 12)
        sd = co.init_state_dict(hp, seed=0)

        def step():
            co.ae_loss(sd, hp, x)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return batch / dt, dt, kind


CPU_SAMPLE = {'reference': 'behavenet.models.AE.loss(accumulate_grad=True) of the unmodified reference (baseline/_ref, '
                           'PyTorch eager fp32, all host threads) on the full 256-frame batch (chunks 200+56)',
              'port': 'oracle/cae_oracle.ae_loss (torch eager fp32, all host threads; baseline/_ref not installed) on '
                      'the full 256-frame batch (chunks 200+56)'}


def cpu_arhmm_rate(reps):
    from oracle import arhmm_oracle as ao
    p = ao.synth_params(ARHMM_K, ARHMM_D, ARHMM_LAGS, seed=0)
    X = ao.sample_batch(p, 16, ARHMM_T, seed=0)
    ao.e_step(p, [X[0]])                        # numba compile
    t0 = time.perf_counter()
    for _ in range(reps):
        ao.e_step(p, [X[i] for i in range(16)])
    dt = (time.perf_counter() - t0) / reps
    return 16 * ARHMM_T / dt, dt


ARHMM_CPU_SAMPLE = ('16 of the 2048 trials x 1000 steps, oracle/arhmm_oracle.e_step (numpy emissions + numba fp64 '
                    'log-space messages, python loop over trials = ssm execution model; ssm itself is not installable)')


def cpu_baselines():
    """The reference paths on this box's host cores (bounded samples), rank 0 at N = 1."""
    cores = os.cpu_count() or 1
    v, dt, kind = cpu_cae_rate(CAE_BATCH_PER_GPU, 3, 1, cores)
    cae = {'value': v, 'unit': 'frames/s', 'cores': cores, 'kind': kind,
           'sample': '3 steps after 1 warm-up: ' + CPU_SAMPLE[kind]}
    hv, _ = cpu_arhmm_rate(1)
    hm = {'value': hv, 'unit': 'timesteps/s', 'cores': 1, 'kind': 'port', 'sample': ARHMM_CPU_SAMPLE}
    return {'cae': cae, 'arhmm': hm}


def reference_eager_b200(device):
    """Informational: the unmodified reference classes executed by eager PyTorch + cuDNN on this B200 (what
    the reference itself does on a GPU; SURVEY.md section 2.1 calls it "the bar to beat"), same batch,
    same metric, TF32 convolutions on (torch's default) and off."""
    model, _ = reference_ae(device)
    if model is None:
        return {'unavailable': 'baseline/_ref not installed'}
    x = torch.rand(CAE_BATCH_PER_GPU, 1, 128, 128, generator=torch.Generator().manual_seed(0)).to(device)
    out = {}
    old = torch.backends.cudnn.allow_tf32
    try:
        for tag, flag in (('tf32', True), ('fp32', False)):
            torch.backends.cudnn.allow_tf32 = flag

            def step():
                model.zero_grad()
                model.loss({'images': x[None]}, accumulate_grad=True)
            ms = timed(step, 10, 3)
            out[tag] = {'frames_per_s': CAE_BATCH_PER_GPU / (float(np.median(ms)) * 1e-3), 'ms_per_step': float(np.median(ms))}
    finally:
        torch.backends.cudnn.allow_tf32 = old
    out['note'] = ('behavenet.models.AE.loss from baseline/_ref on cuda, eager PyTorch %s + cuDNN, 256 frames resident '
                   'in HBM, median of 10 steps after 3 warm-ups (CUDA events)' % torch.__version__)
    return out


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path -- behavenet.models.AE.loss from
    baseline/_ref, eager PyTorch on all host threads -- on the GPU arm's config (full 256-frame batch,
    chunks 200 + 56), K timed steps after W warm-ups.  Rank 0 only; other ranks exit without work."""
    if int(os.environ.get('RANK', '0')) != 0:
        return
    cores = os.cpu_count() or 1
    value, dt, kind = cpu_cae_rate(CAE_BATCH_PER_GPU, args.steps, args.warmup, cores)
    hvalue, hdt = cpu_arhmm_rate(max(1, min(args.steps, 3)))
    _OUT.write(json.dumps({
        'impl': 'reference',
        'metric': 'CAE train frames/sec (C2: 128x128x1, 12 latents, fwd+loss+bwd)',
        'value': value, 'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': dt * 1e3, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'C2: CAE 128x128x1, 12 latents, batch=256 per GPU, AE.loss fwd+bwd, '
                               'default 5-layer arch', 'global_batch': CAE_BATCH_PER_GPU,
                   'parallelism': 'host cores (one process)'},
        'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': cores, 'kind': kind,
                         'sample': '%d steps after %d warm-ups: %s' % (args.steps, args.warmup, CPU_SAMPLE[kind])},
        'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'arhmm': {'metric': 'ARHMM E-step timesteps/sec (C4: K=16, lag 2, D=12)', 'value': hvalue,
                  'unit': 'timesteps/s', 'ms_per_step': hdt * 1e3, 'dtype': 'f64',
                  'cpu_baseline': {'value': hvalue, 'unit': 'timesteps/s', 'cores': 1, 'kind': 'port',
                                   'sample': ARHMM_CPU_SAMPLE},
                  'e2e': {'value': hvalue, 'unit': 'timesteps/s', 'h2d_bytes_per_step': 0,
                          'd2h_bytes_per_step': 0}},
    }) + '\n')
    _OUT.flush()


def _claim_stdout():
    """NCCL / torchrun helpers may print to fd 1; the driver expects ONE JSON line there.  Route fd 1
    to stderr for the whole run and return a writer on the original stdout for the final line."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(saved, 'w')


def main():
    global _OUT
    _OUT = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=50)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)
        from behavenet_b200 import parallel
        parallel.shutdown()


if __name__ == '__main__':
    main()

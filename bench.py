"""Benchmark of the two hot paths on synthetic data of BASELINE.json's shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Headline workload (N GPUs, weak scaling): config C2 -- CAE 128x128x1, 12 latents, 256 frames per
GPU, one ``AE.loss(data, accumulate_grad=True)`` call per step (forward + fused MSE + backward over
all reference chunks, plus the bucketed gradient all-reduce when N > 1), frames resident in HBM.
The same JSON line carries: ``sustained`` (the same step for >= 2 s), ``arhmm`` (config C4: K=16, lag 2,
D=12; weak = 2048 trials x 1000 steps per GPU, strong = 2048 trials in total), ``psvae`` (config C3, weak
and strong), ``c5`` (config C5: 1,000,188 uint8 frames resident in HBM -> encoder -> E-step),
``matmul_peaks_here`` (cuBLAS TF32 / bf16 measured in this run: the roofline denominators) and
``reference_eager_b200`` (the unmodified reference on this GPU through eager PyTorch + cuDNN, informational).

``--impl reference`` times the UNMODIFIED reference (baseline/_ref, installed by baseline/install_ref.sh) on
the host cores: behavenet.models.AE.loss on the full 256-frame batch.
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


def timed_block(fn, steps, device):
    """K back-to-back calls between a barrier + synchronize on both sides; ms per call, max over ranks."""
    torch.cuda.synchronize()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    return dist_max(e0.elapsed_time(e1), device) / steps


def sustained(fn, seconds, device):
    """Back-to-back calls for at least ``seconds`` of device time (the power / clock state of a long run, like
    the sustained figure of MEASURED_PEAKS.json) -> (ms per call, calls)."""
    torch.cuda.synchronize()
    barrier()
    calls, total = 0, 0.0
    while total < seconds * 1e3:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        e1.synchronize()
        total += e0.elapsed_time(e1)
        calls += 20
        # every rank must leave the loop after the same number of calls (collectives inside fn)
        if dist_max(total, device) >= seconds * 1e3:
            break
    return dist_max(total, device) / calls, calls


def kernel_traffic():
    """dram read + write bytes per launch of the profiled kernels, from the ncu export of the CURRENT build
    (profiles/r02_kernel_traffic.json, written by scripts/ncu_traffic.py); {} when no capture is committed."""
    p = os.path.join(ROOT, 'profiles', 'r02_kernel_traffic.json')
    if os.path.exists(p):
        return json.load(open(p))
    return {}


def time_layer_kernel(model, device, side, layer, op, name, iters=20):
    """One layer kernel timed alone through bn_cae_layer_op (CUDA events on the launch stream, after
    warm-up): side 0/1 = encoder/decoder layer, op 0/1/2 = forward / backward-data / weight gradient."""
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
        wshape = params[2 * layer].shape
    else:
        c0, h0, w0 = hp['ae_decoding_starting_dim']
        cs, hs, wsm = ((c0, h0, w0) if layer == 0 else
                       (hp['ae_decoding_n_channels'][layer - 1], hp['ae_decoding_y_dim'][layer - 1],
                        hp['ae_decoding_x_dim'][layer - 1]))
        cb, hb, wb = (hp['ae_decoding_n_channels'][layer], hp['ae_decoding_y_dim'][layer],
                      hp['ae_decoding_x_dim'][layer])
        wshape = params[2 * drv.n_layers + 6 + 2 * layer].shape
    fprop_form = (side == 0 and op == 0) or (side == 1 and op == 1)
    big = torch.rand((n, hb, wb, cb), device=device)
    small = torch.rand((n, hs, wsm, cs), device=device)
    if op == 2:
        src, src2, out = big, small, torch.zeros(wshape, device=device)
    else:
        src, src2 = (big if fprop_form else small), None
        out = torch.empty((n, hs, wsm, cs) if fprop_form else (n, hb, wb, cb), device=device)
    lib = _lib.lib()

    def run():
        _lib.check(lib.bn_cae_layer_op(drv.plan(device), side, layer, op, n, src.data_ptr(), _lib.ptr(src2),
                                       out.data_ptr(), drv.table(params), packed.data_ptr(), ws.data_ptr(),
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
    key = name.split(' ')[0]
    return {'kernel': name, 'us': us, 'gflop': gflop, 'tflops': gflop / us * 1e3, 'launches': iters,
            'traffic_bytes': kernel_traffic().get(key)}


def measure_matmul_peaks(device, sustained_s=2.0):
    """cuBLAS dense GEMM throughput on THIS GPU, measured the way MEASURED_PEAKS.json measures bf16: torch.matmul
    8192^3, best of 10 (burst) and back to back for ``sustained_s`` seconds (sustained), for TF32 operands (the
    MMA kind the conv kernels use) and for bf16 (cross-check against the driver's file)."""
    out = {}
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        for tag, dt in (('tf32', torch.float32), ('bf16', torch.bfloat16)):
            a = torch.randn(8192, 8192, device=device, dtype=dt)
            b = torch.randn(8192, 8192, device=device, dtype=dt)
            for _ in range(3):
                a @ b
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                a @ b
                e1.record()
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1))
            flop = 2 * 8192 ** 3
            out[tag + '_tflops'] = flop / (best * 1e-3) * 1e-12
            n, tot = 0, 0.0
            while tot < sustained_s * 1e3:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(20):
                    a @ b
                e1.record()
                e1.synchronize()
                tot += e0.elapsed_time(e1)
                n += 20
            out[tag + '_tflops_sustained'] = flop * n / (tot * 1e-3) * 1e-12
            del a, b
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    out['how'] = ('torch.matmul 8192^3 on this GPU in this run: best of 10 (burst) and back to back for %.0f s (sustained); '
                  'tf32 = fp32 tensors with torch.backends.cuda.matmul.allow_tf32' % sustained_s)
    return out


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


# ----------------------------------------------------------------------------------------------------
# CAE C2 (headline): weak scaling, 256 frames per GPU
# ----------------------------------------------------------------------------------------------------
def bench_cae(args, device, world, rank, local):
    from behavenet_b200 import _lib, parallel
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
    # the same step for >= 2 s: the clock / power state of a long training run
    sus_sampler = ClockSampler(local)
    if rank == 0:
        sus_sampler.start()
    sus_sampler.mark_start()
    sus_ms, sus_calls = sustained(step, 2.0, device)
    sus_sampler.mark_stop()
    sus_clocks = sus_sampler.stop() if rank == 0 else None
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
    return dict(model=model, hp=hp, B=B, ms_per_step=ms_per_step, launches=launches, clocks=clocks,
                sustained_ms=sus_ms, sustained_calls=sus_calls, sustained_clocks=sus_clocks,
                ms_flushed=float(np.median(ms_flushed)), opt_ms=float(np.median(opt_ms)), e2e_ms=e2e_ms,
                e2e_steps=e2e_steps, h2d=int(x_host.numel() * 4))


# ----------------------------------------------------------------------------------------------------
# ARHMM C4: weak (2048 trials per GPU) and strong (2048 trials in total) scaling
# ----------------------------------------------------------------------------------------------------
def make_hmm():
    from behavenet_b200.ssm import HMM
    from oracle import arhmm_oracle as ao         # synthetic parameters / data generator only
    p = ao.synth_params(ARHMM_K, ARHMM_D, ARHMM_LAGS, seed=0)
    hmm = HMM(ARHMM_K, ARHMM_D, observations='ar', observation_kwargs={'lags': ARHMM_LAGS})
    hmm.init_state_distn.log_pi0, hmm.transitions.log_Ps = p.log_pi0, p.log_Ps
    hmm.observations.As, hmm.observations.bs, hmm.observations.Sigmas = p.As, p.bs, p.Sigmas
    return hmm, p


def bench_arhmm(args, device, world, rank):
    from behavenet_b200 import _lib, parallel
    from oracle import arhmm_oracle as ao
    hmm, p = make_hmm()
    out = {}
    X = ao.sample_batch(p, ARHMM_TRIALS_PER_GPU, ARHMM_T, seed=rank)
    trials = [X[i] for i in range(X.shape[0])]

    def make_step(st, shard):
        def estep():
            Ez, Ezz, logZ = hmm._run_estep(st, True, shard)
            if world > 1:
                stats = torch.cat([Ezz.sum(0).double().reshape(-1), logZ.sum().reshape(1)])
                parallel.all_reduce_sum(stats)
        return estep

    # weak: every rank owns 2048 trials
    st = hmm._stage(trials)
    step = make_step(st, None)
    for _ in range(max(3, args.warmup)):
        step()
    h0 = _lib.launch_count()
    ms = timed_block(step, args.steps, device)
    out['launches'] = _lib.launch_count() - h0
    n_ts = ARHMM_TRIALS_PER_GPU * ARHMM_T
    out['weak_ms'] = ms
    out['weak_value'] = n_ts * world / (ms * 1e-3)
    # strong: the 2048 trials of BASELINE config 4 sharded over the ranks (rank 0's trials are "the" data set)
    if world > 1:
        lo, hi = parallel.shard_range(ARHMM_TRIALS_PER_GPU, world, rank)
        st_s = hmm.stage_device(st.x[lo * ARHMM_T:hi * ARHMM_T].contiguous(), [ARHMM_T] * (hi - lo))
        step_s = make_step(st_s, None)
        for _ in range(3):
            step_s()
        ms_s = timed_block(step_s, args.steps, device)
    else:
        ms_s = ms
    out['strong_ms'] = ms_s
    out['strong_value'] = n_ts / (ms_s * 1e-3)

    # end to end from the python list of host arrays the reference hands to ssm (arhmm_grid_search.py:170):
    # (a) what EM does -- the list is staged ONCE (gather into pinned memory + H2D), then every iteration's
    # E-step re-uses it (hmm.py caches by identity + fingerprint) and reads its statistics back;
    # (b) cold: staging + one E-step + read-back per call.
    n_iter = 20                  # the reference's default number of EM iterations (configs/arhmm_jsons/arhmm_training.json:17)

    def em_like():
        hmm.clear_cache()
        s2 = hmm._stage(trials)
        tot = 0.0
        for _ in range(n_iter):
            Ez, Ezz, logZ = hmm._run_estep(s2, True)
            tot += float(logZ.sum().item())
            Ezz.sum(0).cpu()
        return tot

    def cold():
        hmm.clear_cache()
        s2 = hmm._stage(trials)
        Ez, Ezz, logZ = hmm._run_estep(s2, True)
        return float(logZ.sum().item()), Ezz.sum(0).cpu()
    em_like()
    barrier()
    t0 = time.perf_counter()
    em_like()
    torch.cuda.synchronize()
    em_s = dist_max(time.perf_counter() - t0, device)
    cold()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        cold()
    torch.cuda.synchronize()
    cold_s = dist_max((time.perf_counter() - t0) / 3, device)
    out['e2e_value'] = n_ts * world * n_iter / em_s
    out['e2e_cold_value'] = n_ts * world / cold_s
    out['e2e_iters'] = n_iter
    out['n_ts'] = n_ts
    return out


# ----------------------------------------------------------------------------------------------------
# PS-VAE C3: 128x128x2, 16 latents, 4 labels; weak (512 frames per GPU) and strong (512 frames in total)
# ----------------------------------------------------------------------------------------------------
PSVAE_GFLOP_PER_FRAME = 2.111          # train step, 128x128x2, 16 latents
PSVAE_BATCH_PER_GPU = 512


def bench_psvae(device, world, rank, args):
    """One PSVAE.loss(accumulate_grad=True) per step: contiguous frame shards per rank, reference chunks of
    200 frames that span ranks exchange their (FF output, logvar, eps) rows (one small all-reduce), one
    bucketed gradient all-reduce."""
    import copy
    from behavenet_b200.models import PSVAE
    from oracle import cae_oracle as co          # seeded synthetic parameters only
    np.random.seed(0)
    hp = co.make_hparams(2, 128, 128, 16, 'ps-vae', 4)
    model = PSVAE(copy.deepcopy(hp))
    model.load_state_dict(co.init_state_dict(hp, seed=0))
    model.to(device)
    model.curr_epoch = 1
    model.data_parallel = world > 1
    steps = max(3, min(args.steps, 10))
    res = {}
    for tag, B in (('weak', PSVAE_BATCH_PER_GPU * world), ('strong', PSVAE_BATCH_PER_GPU)):
        if tag == 'strong' and world == 1:
            res['strong'] = dict(res['weak'])
            break
        g = torch.Generator().manual_seed(1)
        x = torch.rand(B, 2, 128, 128, generator=g).to(device)
        y = torch.randn(B, 4, generator=g).to(device)
        eps = torch.randn(B, 16, generator=g).to(device)
        data = {'images': x[None], 'labels': y[None]}

        def step():
            model.zero_grad()
            model.invalidate_packed()
            model.loss(data, accumulate_grad=True, eps=eps)
        for _ in range(3):
            step()
        ms = timed_block(step, steps, device)
        res[tag] = {'value': B / (ms * 1e-3), 'ms_per_step': ms, 'global_batch': B,
                    'frames_per_gpu': B // world, 'tflops_per_gpu': B / (ms * 1e-3) * PSVAE_GFLOP_PER_FRAME * 1e-3 / world}
        del x, y, eps, data
    return {'metric': 'PS-VAE train frames/sec (C3: 128x128x2, 16 latents, 4 labels, fwd+loss+bwd)',
            'unit': 'frames/s', 'value': res['weak']['value'], 'ms_per_step': res['weak']['ms_per_step'],
            'steps': steps, 'global_batch': res['weak']['global_batch'],
            'tflops_per_gpu': res['weak']['tflops_per_gpu'], 'weak': res['weak'], 'strong': res['strong'],
            'note': 'weak = 512 frames per GPU, strong = the 512-frame batch of BASELINE config 3 sharded over the '
                    'ranks; includes the label R^2 on the host and the per-chunk latent-block launches, as the '
                    'reference loss() does; frames resident in HBM'}


# ----------------------------------------------------------------------------------------------------
# C5: encode-only export -> ARHMM E-step on device-resident latents (Musall-shape synthetic video)
# ----------------------------------------------------------------------------------------------------
C5_TRIALS, C5_T, C5_LATENTS = 5292, 189, 12
C5_ENCODE_GFLOP_PER_FRAME = 0.3539        # 2-channel encoder incl. FF


def bench_c5(device, world, rank):
    """BASELINE config 5: 5292 trials x 189 frames of (2, 128, 128) uint8 video (= 1,000,188 frames) sharded by
    trial over the ranks and RESIDENT in HBM as bytes; every trial goes through the encoder (uint8 -> /255 in
    the first layer's loader, reference data_generator.py:258-263 + eval.py:6-118) and the latents stay on the
    GPU that produced them for one ARHMM E-step (K=16, lag 2; arhmm_grid_search.py:170).  No inter-GPU traffic
    between the stages; one all-reduce of the E-step statistics."""
    import copy
    from behavenet_b200 import parallel
    from behavenet_b200.models import AE
    from oracle import cae_oracle as co
    hp = co.make_hparams(2, 128, 128, C5_LATENTS)
    model = AE(copy.deepcopy(hp))
    model.load_state_dict(co.init_state_dict(hp, seed=0))
    model.to(device)
    model.eval()
    lo, hi = parallel.shard_range(C5_TRIALS, world, rank)
    n_trials = hi - lo
    n_frames = n_trials * C5_T
    # noise plus structure, generated on the device in slabs (a 33 GB pinned host copy would only time malloc)
    frames = torch.empty(n_frames, 2, 128, 128, dtype=torch.uint8, device=device)
    yy = torch.arange(128, device=device, dtype=torch.float32)[:, None]
    xx = torch.arange(128, device=device, dtype=torch.float32)[None, :]
    gen = torch.Generator(device=device).manual_seed(1000 + rank)
    slab = 8192
    for o in range(0, n_frames, slab):
        n = min(slab, n_frames - o)
        ph = torch.rand(n, 2, 1, 1, device=device, generator=gen) * 6.28
        base = 110 + 70 * torch.sin(yy * 0.07 + ph) * torch.cos(xx * 0.05 + 2 * ph)
        noise = torch.randint(0, 48, (n, 2, 128, 128), device=device, generator=gen, dtype=torch.int16)
        frames[o:o + n] = (base + noise).clamp_(0, 255).to(torch.uint8)
        del ph, base, noise
    hmm, _ = make_hmm()
    chunk = 4096
    lat = torch.empty(n_frames, C5_LATENTS, dtype=torch.float32, device=device)

    def encode():
        with torch.no_grad():
            for o in range(0, n_frames, chunk):
                lat[o:o + chunk] = model.encoding(frames[o:o + chunk])[0]

    def estep():
        st = hmm.stage_device(lat, [C5_T] * n_trials)
        Ez, Ezz, logZ = hmm.expected_states_device(st)
        stats = torch.cat([Ezz.sum(0).double().reshape(-1), logZ.sum().reshape(1)])
        if world > 1:
            parallel.all_reduce_sum(stats)
        return stats
    encode()
    estep()
    enc_ms = timed_block(encode, 2, device)
    es_ms = timed_block(estep, 5, device)
    total_frames = C5_TRIALS * C5_T
    enc_rate = total_frames / (enc_ms * 1e-3)
    tfl = enc_rate * C5_ENCODE_GFLOP_PER_FRAME * 1e-3 / world
    # end to end on a bounded sample: 64 trials of uint8 frames from pinned host memory (one byte per pixel
    # across PCIe), encoder, E-step on their latents, statistics read back
    ns = min(64, n_trials)
    host = frames[:ns * C5_T].cpu().pin_memory()
    host_trials = [host[i * C5_T:(i + 1) * C5_T] for i in range(ns)]
    from behavenet_b200.fitting.eval import encode_trials

    def e2e():
        # the public call of the export path: groups of trials copied on a side stream one group ahead of the encoder
        z, lens = encode_trials(model, host_trials, frames_per_launch=2048, device=device)
        st = hmm.stage_device(z, lens)
        Ez, Ezz, logZ = hmm.expected_states_device(st)
        return float(logZ.sum().item()), Ezz.sum(0).cpu()
    e2e()
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        e2e()
    torch.cuda.synchronize()
    e2e_s = dist_max((time.perf_counter() - t0) / 3, device)
    res = {
        'metric': 'C5: encode-only export -> ARHMM E-step, 5292 trials x 189 frames of 2x128x128 uint8 (1,000,188 frames)',
        'frames': total_frames, 'frames_per_gpu': n_frames, 'resident_bytes_per_gpu': int(frames.numel()),
        'encode': {'value': enc_rate, 'unit': 'frames/s', 'ms': enc_ms, 'tflops_per_gpu': tfl},
        'estep': {'value': total_frames / (es_ms * 1e-3), 'unit': 'timesteps/s', 'ms': es_ms},
        'pipeline': {'value': total_frames / ((enc_ms + es_ms) * 1e-3), 'unit': 'frames/s',
                     'note': 'encode + E-step, frames resident in HBM as uint8, latents never leave the GPU'},
        'e2e': {'value': ns * C5_T * world / e2e_s, 'unit': 'frames/s', 'h2d_bytes_per_step': int(host.numel()),
                'd2h_bytes_per_step': int(8 + ARHMM_K * ARHMM_K * 4),
                'sample': '%d trials (%d frames) per rank from pinned host memory per call through '
                          'fitting.eval.encode_trials (H2D at one byte per pixel on a side stream, one 2048-frame group '
                          'ahead of the encoder) + E-step + read-back; host wall clock, max over ranks' % (ns, ns * C5_T)},
    }
    del frames, lat, host, host_trials
    torch.cuda.empty_cache()
    return res


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

    cae = bench_cae(args, device, world, rank, local)
    model, hp, B = cae['model'], cae['hp'], cae['B']
    ms_per_step = cae['ms_per_step']
    cae_value = B / (ms_per_step * 1e-3)
    # (kernels are timed alone BEFORE the seconds-long cuBLAS loops below heat the part up)
    kernels = [
        time_layer_kernel(model, device, 0, 2, 0, 'igemm_tma_kernel<128,2,2> (encoder conv2 forward, M=65536 N=128 K=1600)'),
        time_layer_kernel(model, device, 0, 1, 0, 'fprop_halo_pair_kernel (encoder conv1 forward, M=262144 N=64 K=800)'),
        time_layer_kernel(model, device, 0, 1, 2, 'wgrad_tma_kernel<64,2,2> (encoder conv1 weight gradient, 800x64 over 262144 pixels)'),
        time_layer_kernel(model, device, 1, 3, 0, 'dgrad_halo_pair_kernel (decoder convtranspose3 forward, 4 x [M=262144 N=32 K<=576])'),
        time_layer_kernel(model, device, 1, 2, 0, 'dgrad_halo_pair64_kernel (decoder convtranspose2 forward, 4 x [M=65536 N=64 K<=1152])'),
        time_layer_kernel(model, device, 0, 3, 0, 'igemm_tma_kernel<256,4,1> (encoder conv3 forward, M=16384 N=256 K=3200)'),
    ]
    dom = kernels[0]

    legs = {'cae', 'arhmm', 'psvae', 'c5', 'eager', 'cpu'} if args.legs == 'all' else set(args.legs.split(','))
    hm = bench_arhmm(args, device, world, rank) if 'arhmm' in legs else None
    psvae = c5 = None
    if 'psvae' in legs:
        try:
            psvae = bench_psvae(device, world, rank, args)
        except Exception as exc:                     # the headline line must not depend on the extra legs
            psvae = {'error': repr(exc)[:300]}
    if 'c5' in legs:
        try:
            c5 = bench_c5(device, world, rank)
        except Exception as exc:
            c5 = {'error': repr(exc)[:300]}
    # roofline denominators: cuBLAS TF32 measured in THIS run the way MEASURED_PEAKS.json measures bf16
    # (burst for a kernel timed alone, sustained for the step); half of the driver's bf16 figures beside it.
    # Last, so that the 2 x 2 s of full-power GEMMs do not pre-heat the legs above.
    mm = measure_matmul_peaks(device)
    tf32_burst, tf32_sust = mm['tf32_tflops'], mm['tf32_tflops_sustained']
    half_bf16_burst = peaks['bf16_tflops'] / 2.0
    half_bf16_sust = peaks.get('bf16_tflops_sustained', peaks['bf16_tflops']) / 2.0
    tf = cae_value * CAE_C2_TRAIN_GFLOP_PER_FRAME * 1e-3 / world         # TFLOP/s per GPU
    tf_sus = B / (cae['sustained_ms'] * 1e-3) * CAE_C2_TRAIN_GFLOP_PER_FRAME * 1e-3 / world
    if isinstance(c5, dict) and 'encode' in c5:
        tfl = c5['encode']['tflops_per_gpu']
        c5['encode']['roofline'] = {'bound': 'tensor', 'achieved': tfl, 'peak': tf32_sust, 'unit': 'TFLOP/s',
                                    'frac': tfl / tf32_sust,
                                    'note': '%.4f GFLOP/frame algorithmic (SURVEY 8d) / encode time; peak = cuBLAS TF32 '
                                            'sustained measured in this run' % C5_ENCODE_GFLOP_PER_FRAME}
    eager = None
    if rank == 0 and world == 1 and 'eager' in legs:
        try:
            eager = reference_eager_b200(device)
        except Exception as exc:
            eager = {'error': repr(exc)[:300]}

    if rank != 0:
        return
    cpu = cpu_baselines() if (world == 1 and 'cpu' in legs) else None
    hbm_achieved = hm['weak_value'] / world * ARHMM_BYTES_PER_TIMESTEP / 1e9 if hm else None
    line = {
        'metric': 'CAE train frames/sec (C2: 128x128x1, 12 latents, fwd+loss+bwd)',
        'value': cae_value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'tf32 (fp32 accumulate)' if lib.bn_get_tensor_core_mode() else 'f32',
        'data': 'synthetic',
        'config': {'workload': 'C2: CAE 128x128x1, 12 latents, batch=256 per GPU, AE.loss fwd+bwd, '
                               'default 5-layer arch', 'global_batch': B,
                   'l2': 'inputs 16.8 MB/GPU < L2; ms_per_step_l2_flushed is the same step with a 256 MB L2 flush '
                         'between steps', 'parallelism': 'dp%d' % world},
        'ms_per_step_l2_flushed': cae['ms_flushed'],
        'sustained': {'ms_per_step': cae['sustained_ms'], 'value': B / (cae['sustained_ms'] * 1e-3), 'unit': 'frames/s',
                      'steps': cae['sustained_calls'], 'clocks': cae['sustained_clocks'],
                      'note': 'the same step back to back for >= 2 s of device time (the burst figure above covers '
                              '%d steps = %.0f ms)' % (args.steps, ms_per_step * args.steps)},
        'optimizer_step_ms': cae['opt_ms'],
        'gpu_launches': int(cae['launches']),
        'clocks': cae['clocks'],
        'e2e': {'value': B / (cae['e2e_ms'] * 1e-3), 'unit': 'frames/s',
                'h2d_bytes_per_step': cae['h2d'], 'd2h_bytes_per_step': 16 * world,
                'steps': cae['e2e_steps'],
                'note': 'AE.loss on frames copied from pinned host memory every step (the copy of step i+1 '
                        'overlaps the compute of step i on a side stream; one extra prefetch is inside the '
                        'timed region) + the per-step loss read-back; host wall clock, max over ranks'},
        'roofline': {'bound': 'tensor', 'achieved': dom['tflops'], 'peak': tf32_burst,
                     'unit': 'TFLOP/s', 'frac': dom['tflops'] / tf32_burst,
                     'frac_vs_half_bf16_burst': dom['tflops'] / half_bf16_burst,
                     'traffic': dom['traffic_bytes'],
                     'kernel': dom['kernel'], 'kernel_us': dom['us'], 'kernel_gflop': dom['gflop'],
                     'note': 'dominant kernel timed alone with CUDA events on the launch stream (%d launches after '
                             'warm-up); peak = cuBLAS TF32 8192^3 burst measured in this run (%.1f TFLOP/s; half of the %s '
                             'bf16 burst of MEASURED_PEAKS.json would be %.1f); traffic = dram read+write bytes per '
                             'launch from the ncu export of this build (profiles/r02_kernel_traffic.json) or null'
                             % (dom['launches'], tf32_burst, peak_src, half_bf16_burst)},
        'step_roofline': {'bound': 'tensor', 'achieved': tf, 'peak': tf32_sust, 'unit': 'TFLOP/s',
                          'frac': tf / tf32_sust, 'frac_vs_half_bf16_sustained': tf / half_bf16_sust,
                          'sustained_achieved': tf_sus, 'sustained_frac': tf_sus / tf32_sust,
                          'note': 'whole step: %.3f GFLOP/frame algorithmic / step time, per GPU; peak = cuBLAS TF32 '
                                  'sustained measured in this run (%.1f TFLOP/s; half of the %s bf16 sustained peak would '
                                  'be %.1f)' % (CAE_C2_TRAIN_GFLOP_PER_FRAME, tf32_sust, peak_src, half_bf16_sust)},
        'kernel_rooflines': [{'kernel': d['kernel'], 'us': d['us'], 'tflops': d['tflops'],
                              'frac': d['tflops'] / tf32_burst, 'traffic': d['traffic_bytes']} for d in kernels],
        'matmul_peaks_here': mm,
        'reference_eager_b200': eager,
        'arhmm': None,
        'psvae': psvae,
        'c5': c5,
    }
    if hm is not None:
        line['arhmm'] = {
        'metric': 'ARHMM E-step timesteps/sec (C4: K=16, lag 2, D=12, 2048 trials x 1000 per GPU)',
        'value': hm['weak_value'], 'unit': 'timesteps/s', 'ms_per_step': hm['weak_ms'],
        'scaling': 'weak',
        'strong': {'value': hm['strong_value'], 'ms_per_step': hm['strong_ms'],
                   'note': 'the 2048 trials of BASELINE config 4 in total, sharded by trial over the ranks'},
        'gpu_launches': int(hm['launches']), 'dtype': 'f32 (scaled messages), f64 log-normaliser',
        'e2e': {'value': hm['e2e_value'], 'unit': 'timesteps/s',
                'h2d_bytes_per_step': int(hm['n_ts'] * ARHMM_D * 4 / hm['e2e_iters']),
                'd2h_bytes_per_step': int(8 + ARHMM_K * ARHMM_K * 4),
                'cold_value': hm['e2e_cold_value'],
                'note': 'whole job (all ranks), host wall clock, max over ranks.  value = what EM does: the python '
                        'list of 2048 host arrays per rank is staged once (pooled pinned gather + H2D) and %d '
                        'E-steps re-use it, each reading its statistics back; cold_value = staging + ONE E-step '
                        '+ read-back per call' % hm['e2e_iters']},
        'roofline': {'bound': 'hbm', 'achieved': hbm_achieved, 'peak': peaks['hbm_gbs'],
                     'unit': 'GB/s', 'frac': hbm_achieved / peaks['hbm_gbs'],
                     'traffic': kernel_traffic().get('arhmm_estep'),
                     'other_bounds': arhmm_other_bounds(hm['weak_ms'] if hm else None),
                     'note': '113 B/timestep algorithmic (48 B latents in + 64 B posteriors out + per-trial outputs), '
                             'whole E-step time; peak %s; traffic = dram read+write of all E-step kernels from the ncu '
                             'export of this build or null.  The binding limits are the T-step serial chain / '
                             'issue slots of the scan and the 3-pass emission GEMM, not HBM (DESIGN.md section 4)'
                             % peak_src},
    }
    if cpu is not None:
        line['cpu_baseline'] = cpu['cae']
        if line['arhmm'] is not None:
            line['arhmm']['cpu_baseline'] = cpu['arhmm']
    _OUT.write(json.dumps(line) + '\n')
    _OUT.flush()


def arhmm_other_bounds(estep_ms):
    """The limits that actually bind the E-step (SURVEY hard part 1), from the ncu export of this build
    (profiles/r02_kernel_traffic.json, 'arhmm_estep_bounds'): the ISSUE-SLOT floor of each kernel = executed warp
    instructions / (SMs x 4 schedulers) cycles, and the TENSOR floor of the emission GEMM = its tensor-pipe-active
    cycles; frac = (sum of the larger floor per kernel) / measured E-step time."""
    b = kernel_traffic().get('arhmm_estep_bounds')
    if not b or not estep_ms:
        return None
    out, floor_us = {}, 0.0
    for name, k in b.items():
        clk_mhz = k['sm_cycles'] / k['duration_us']
        issue_us = k['warp_instructions'] / (148 * 4) / clk_mhz
        tensor_us = k['tensor_pipe_active_pct'] / 100.0 * k['duration_us']
        out[name] = {'issue_slot_floor_us': issue_us, 'tensor_floor_us': tensor_us, 'ncu_duration_us': k['duration_us'],
                     'issue_active_pct': k['issue_active_pct']}
        floor_us += max(issue_us, tensor_us)
    out['floor_us'] = floor_us
    out['frac'] = floor_us / (estep_ms * 1e3)
    return out


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
    ap.add_argument('--legs', default='all', help="'all' or a comma list of cae,arhmm,psvae,c5,eager,cpu (diagnostic runs)")
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

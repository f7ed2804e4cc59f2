#!/usr/bin/env python
"""Benchmark of the soundgen source-filter synthesis path on B200.

  python bench.py --gpus N --steps K --warmup W [--config 3] [--batch B]
  python bench.py --impl reference ...      (the reference algorithm on the host cores)

A "step" is one pass of the whole hot path (control stage, amplitude matrices, additive
synthesis, epoch joining, envelopes, noise, fused STFT filter, mixing) over one batch of
synthetic soundgen() calls of the chosen BASELINE.json config.  `value` is audio seconds
synthesised per second with inputs resident in HBM; `e2e` includes the H2D copy of the
step's inputs from pinned host memory and the D2H read of every waveform.
One JSON line is printed by rank 0.
"""
from __future__ import annotations

import argparse
import faulthandler
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

T_START = time.perf_counter()
METRIC = 'audio_seconds_synthesised_per_second'
UNIT = 'audio-s/s'


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def run(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(',')])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for i, nm in enumerate(names):
                    if r[3 + i].lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                continue
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': mx or None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def oracle_one(kw):
    """One soundgen() call through the CPU oracle; returns (seconds of audio, samples)."""
    from oracle import soundgen_oracle as so
    from oracle.soundgen_call import soundgen as osg
    kw = dict(kw)
    z = kw.pop('z', None)
    u = kw.pop('u', None)
    if 'seed' in kw:      # cfg0 / cfg4: R's own stream after set.seed(seed)
        from oracle.rrng import RRng
        rng = RRng(kw.pop('seed'))
    else:
        rng = so.RStream(z=np.concatenate(z) if z else None, u=np.concatenate(u) if u else None)
    y = osg(rng=rng, **kw)
    return y.size / float(kw.get('samplingRate', 16000))


def run_reference(args, rank, world):
    """The reference algorithm (CPU restatement under oracle/) on all host cores, bounded sample."""
    if rank != 0:
        return
    import multiprocessing as mp
    from soundgen_beta_b200 import workloads
    cores = os.cpu_count() or 1
    per_step = max(cores, 2 * cores if args.config in (1,) else cores)
    if args.config == 4:
        per_step = max(per_step, 33)
    calls = workloads.CONFIGS[args.config](n=per_step)
    with mp.Pool(cores) as pool:
        for _ in range(args.warmup):
            pool.map(oracle_one, calls[:cores])
        t0 = time.perf_counter()
        audio = 0.0
        for _ in range(args.steps):
            audio += sum(pool.map(oracle_one, calls))
        dt = time.perf_counter() - t0
    val = audio / dt
    line = {'impl': 'reference', 'metric': METRIC, 'value': val, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic', 'gpu_launches': 0,
            'config': {'workload': workloads.NAMES[args.config], 'calls_per_step': len(calls)},
            'cpu_baseline': {'value': val, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': '%d calls of %s per step (numpy restatement of the R reference; '
                                       'R itself is not installed)' % (len(calls), workloads.NAMES[args.config])},
            'e2e': {'value': val, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--config', type=int, default=3)
    ap.add_argument('--batch', type=int, default=0, help='calls per GPU per step (default: the config size)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--pipeline', type=int, default=8,
                    help='sub-batches (handles / CUDA streams) of the end-to-end measurement')
    ap.add_argument('--runners', type=int, default=3,
                    help='host threads that run kernels in the end-to-end pipeline (plus one uploader, one fetcher)')
    ap.add_argument('--watchdog', type=int, default=1500, help='seconds after which a stuck run dumps its stacks and exits')
    args = ap.parse_args()
    faulthandler.dump_traceback_later(args.watchdog, exit=True)   # a stuck run reports where, instead of hanging the box
    rank, world, local = env_int('RANK', 0), env_int('WORLD_SIZE', 1), env_int('LOCAL_RANK', 0)
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return

    import __graft_entry__ as ge
    ge.build()
    import soundgen_beta_b200 as sg
    from soundgen_beta_b200 import _abi, workloads
    L = _abi.load()
    if L.sgb_device_count() < 1:
        raise SystemExit('bench.py: no CUDA device (the product path has no CPU fallback)')
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    assert L.sgb_set_device(local) == 0
    from soundgen_beta_b200 import sharding
    numa_cpus = sharding.bind_to_gpu_numa(L, local) if world > 1 else []

    sizes = {0: 1, 1: 1024, 2: 4096, 3: 8192, 4: max(1, 65536 // world)}   # cfg4: one sweep shared by the ranks
    n = args.batch or sizes[args.config]
    gen = workloads.CONFIGS[args.config]
    from soundgen_beta_b200 import sharding
    calls = gen(n=n, seed=sharding.shard_seed(args.config, rank))
    srs = np.array([float(kw.get('samplingRate', 16000)) for kw in calls])

    def add_calls(builder, some):
        for kw in some:
            builder.add_soundgen(**kw)

    bb = sg.BatchBuilder(u_dtype=np.float32)   # uniforms travel as float32 (halves the PCIe bytes)
    add_calls(bb, calls)
    desc = bb.build()
    sg.pin_desc(desc)   # page-lock the host pools
    bt = sg.Batch()
    bt.upload(desc)

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- warm-up (also sizes the device pools) ----
    info = None
    for _ in range(max(args.warmup, 1)):
        info = bt.run()
    lens = bt.lengths()
    audio_s = float(np.sum(lens / srs))
    out = np.zeros(int(lens.sum()), dtype=np.float32)
    L.sgb_pin(out.ctypes.data, out.nbytes)
    bt.fetch(np.float32, out=out)

    # ---- timed: inputs resident in HBM ----
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    stage_ms = np.zeros(len(_abi.T_NAMES))
    t0 = time.perf_counter()
    launches = 0
    for _ in range(args.steps):
        info = bt.run()   # ends with a stream synchronise
        stage_ms += np.array(list(info.ms))
        launches += info.kernel_launches
    barrier()
    dt_wall = time.perf_counter() - t0
    # device time of the K steps: CUDA events recorded on the launching stream around every step
    # (first launch to last kernel, host-side layout between the stages included)
    dt = float(stage_ms[_abi.T_NAMES.index('total')]) * 1e-3
    # ---- timed: end to end through the public batch API (H2D + run + D2H of every waveform) ----
    # the batch is cut into `--pipeline` sub-batches that flow through an upload / run / fetch software
    # pipeline (one uploader, `--runners` kernel threads, one fetcher), so the transfers of one
    # sub-batch overlap the kernels of others (no work is skipped)
    npipe = max(1, min(args.pipeline, len(calls)))
    subs = []
    for i in range(npipe):
        lo, hi = sharding.shard_range(len(calls), i, npipe)
        sb = sg.BatchBuilder(u_dtype=np.float32)
        add_calls(sb, calls[lo:hi])
        sd = sb.build()
        sg.pin_desc(sd)
        subs.append(sd)
    pipe = sg.PipelinedBatches(subs, runners=args.runners)

    def report_stuck():   # fires shortly before the watchdog: where is every handle waiting?
        for i, hb in enumerate([bt] + pipe.batches):
            st = np.zeros(12, dtype=np.int32)
            L.sgb_batch_debug_state(hb.h, st.ctypes.data, 12)
            print('handle %d: wait %d, stage events done %s' % (i, st[0], st[1:].tolist()), file=sys.stderr, flush=True)
    stuck = threading.Timer(max(3, args.watchdog - 8 - (time.perf_counter() - T_START)), report_stuck)
    stuck.daemon = True
    stuck.start()
    pipe.run_steps(2)    # warm-up: sizes the pools of every handle, pins the outputs
    barrier()
    t1 = time.perf_counter()
    pipe.run_steps(args.steps)
    barrier()
    dt_e2e = time.perf_counter() - t1
    d2h_f32 = sum(o.nbytes for o in pipe.outs)
    # the same end to end, fetching 16-bit PCM (what the reference's savePath branch writes): half the D2H bytes
    pipe.run_steps(1, dtype=np.int16)
    barrier()
    t2 = time.perf_counter()
    pipe.run_steps(args.steps, dtype=np.int16)
    barrier()
    dt_wav = time.perf_counter() - t2
    d2h_i16 = sum(o.nbytes for o in pipe.outs)
    stuck.cancel()
    d2h_bytes = d2h_f32
    clocks = sampler.stop()
    stage_ms /= args.steps

    dt, dt_e2e, audio_total, dt_wav = sharding.aggregate(dist, dt, dt_e2e, audio_s,
                                                         device='cuda' if dist is not None else None, extra=(dt_wav,))
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = audio_total * args.steps / dt
    e2e = audio_total * args.steps / dt_e2e
    # ---- roofline of the dominant kernel (K1 additive synthesis: FP32-pipe bound) ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    import ctypes
    fp32 = ctypes.c_double(0)
    L.sgb_measure_fp32_peak(ctypes.byref(fp32))
    ms_synth = float(stage_ms[_abi.T_NAMES.index('synth')])
    ms_filter = float(stage_ms[_abi.T_NAMES.index('filter')])
    ms_noise = float(stage_ms[_abi.T_NAMES.index('noise')])
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, 'profiles', 'r01_traffic.json')))
        key = 'cfg%d_%d' % (args.config, len(calls))
        traffic = tj.get(key, {}).get('k_synth', {}).get('dram_bytes_per_launch')
    except Exception:
        pass
    roofline = None
    if ms_synth > 0 and info.synth_partials > 0:
        ach = 6.0 * info.synth_partials / (ms_synth * 1e-3) / 1e12
        roofline = {'kernel': 'k_synth', 'bound': 'fp32', 'achieved': ach, 'peak': fp32.value,
                    'unit': 'TFLOP/s', 'frac': ach / fp32.value if fp32.value else None, 'traffic': traffic,
                    'peak_source': 'measured in this run: sgb_measure_fp32_peak (FFMA2 chains)',
                    'algorithmic': '6 flop per partial-sample x %d partial-samples per launch' % info.synth_partials}
    hbm = peaks.get('hbm_gbs', 6650.0)
    roof_filter = None
    if ms_filter > 0 and info.filter_samples > 0:
        ach = 8.0 * info.filter_samples / (ms_filter * 1e-3) / 1e9
        roof_filter = {'kernel': 'k_stft<filter>', 'bound': 'hbm', 'achieved': ach, 'peak': hbm, 'unit': 'GB/s',
                       'frac': ach / hbm, 'traffic': None,
                       'peak_source': 'MEASURED_PEAKS.json' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s',
                       'algorithmic': '8 B per output sample x %d samples per launch' % info.filter_samples}
    if roofline is None:
        roofline = roof_filter

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        m = {0: 8, 1: 24, 2: 4, 3: 6, 4: 33}[args.config]
        sample = (calls * m)[:m]
        tc = time.perf_counter()
        a = sum(oracle_one(kw) for kw in sample)
        tc = time.perf_counter() - tc
        cpu = {'value': a / tc, 'unit': UNIT, 'cores': 1, 'kind': 'port',
               'sample': 'first %d calls of the workload, numpy restatement of the R reference '
                         '(R is not installed on this image)' % len(sample)}
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3, 'wall_ms_per_step': dt_wall / args.steps * 1e3,
            'timing': 'cuda events on the launching stream, max over ranks', 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 (f64 phase / control)', 'data': 'synthetic',
            'config': {'workload': workloads.NAMES[args.config], 'calls_per_gpu': len(calls),
                       'audio_seconds_per_gpu_step': audio_s, 'l2': 'inputs and intermediates larger than L2',
                       'uniforms': 'float32'},
            'e2e': {'value': e2e, 'unit': UNIT, 'h2d_bytes_per_step': int(bb.h2d_bytes()),
                    'd2h_bytes_per_step': int(d2h_bytes), 'ms_per_step': dt_e2e / args.steps * 1e3,
                    'pipeline': npipe, 'runners': args.runners, 'cpus_bound': len(numa_cpus)},
            'e2e_wav16': {'value': audio_total * args.steps / dt_wav, 'unit': UNIT, 'ms_per_step': dt_wav / args.steps * 1e3,
                          'd2h_bytes_per_step': int(d2h_i16),
                          'note': 'same pipeline, waveforms fetched as 16-bit PCM (the savePath / WAV format)'},
            'gpu_launches': int(launches), 'clocks': clocks, 'roofline': roofline,
            'roofline_filter': roof_filter, 'cpu_baseline': cpu,
            'stage_ms': {nm: float(v) for nm, v in zip(_abi.T_NAMES, stage_ms)},
            'failed_calls': int(info.n_failed)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

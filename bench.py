"""Benchmark of the 2G-GCN hot path (BASELINE.json: video frames/s of the forward on MPHOI-72-shaped data).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one forward pass (eval, no_grad) of the drop-in TGGCN over one padded batch of synthetic
videos: B=8 videos per GPU, T=128 frames, H=2 humans, O=4 object slots, V=26 nodes, hidden 512,
stage-2 settings (learned Gumbel gates + local-maximum filter) — BASELINE.json configs[1].  Videos are
independent, so N GPUs run N disjoint batches with no data-path collective ("scaling": "weak").

`value`      : frames/s with inputs (and the pre-drawn Gumbel noise) resident in HBM, CUDA events, max over ranks.
`e2e`        : the same through the public call with HOST buffers: every step copies one full input batch from pinned host
               memory (double-buffered on a side stream, 2g-gcn_b200/feeder.py), draws the Gumbel noise on the CPU like the
               reference does and copies it, runs the forward and copies all outputs back to pinned host memory, where the
               host reads them one step later (while the next step is already running).
`roofline`   : the dominant kernel (largest share of the step, per-stage CUDA events measured live).
`cpu_baseline`: the CPU oracle port of the reference path (oracle/tggcn_oracle.py) timed on this box's host cores.
`--impl reference` times that CPU port as the reference arm (the reference itself is pure PyTorch and is not
present on the GPU box; its CPU algorithm is restated 1:1, loops included, in the oracle).
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# BASELINE.json's metric, verbatim; `value` is the inference forward, the training step is reported under `train_step`
METRIC = 'video frames/sec (fwd and train step) at 1/2/4/8 B200 vs host-CPU torch ref'
METRIC_DETAIL = ('value / e2e = 2G-GCN inference forward (eval, no_grad) on MPHOI-72-shaped synthetic videos; '
                 'train_step.value / train_step.e2e = one training step (forward + criterion + backward + Adam) on the same shape')
UNIT = 'frames/s'


def load_peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p['hbm_gbs'], tf_burst=p['bf16_tflops'], tf_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    source='measured')
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source='fallback')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.path = gpu_index, None, None

    def __enter__(self):
        try:
            f = tempfile.NamedTemporaryFile('w', suffix='.csv', delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '200',
                                          '-i', str(self.idx)], stdout=f, stderr=subprocess.DEVNULL)
            time.sleep(0.25)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        if self.proc is None or not self.path or not os.path.exists(self.path):
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['unavailable'])
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            c = [x.strip() for x in line.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), c[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=['no samples'])
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons))


def workload(args):
    pkg = importlib.import_module('2g-gcn_b200')
    shape = pkg.synth.SHAPES[args.shape]
    kwargs = pkg.synth.model_kwargs(shape, hidden_size=args.D, stage=2)
    return pkg, shape, kwargs


def stage_work(shape, B, T, D, hh=True):
    """Algorithmic FLOPs and HBM bytes per forward of each stage (SURVEY.md §8d closed forms; message MLPs
    counted once per sender and kind; fp32 operands).  Returns {stage: (flops, bytes)}."""
    H, O, V = shape.H, shape.O, shape.V
    N = B * T
    nkh = 2 if hh else 1
    f = 4
    gemm = lambda M, Nn, K: (2.0 * M * Nn * K, f * (M * K + Nn * K + M * Nn))
    add = lambda *xs: (sum(x[0] for x in xs), sum(x[1] for x in xs))
    w = {}
    w['geo_gcn'] = (N * (2 * V * (4 * 64 + 64 * 64 + 2 * 64 * 128 + 64 * 128) + 2 * V * V * (128 + 64)),
                    f * N * (4 * V + 128 * V) + f * 28000)
    w['gemm_embed'] = add(gemm(N * H, D, 2048), gemm(N * O, D, 2048), gemm(N, 2048, 128 * V))
    w['gemm_geo2'] = gemm(N, D, 2048)
    w['gemm_gi'] = add(*[gemm(N * E, 3 * D, D) for E in (H, H, O, O, 1, 1)])
    rows = B * (H + O + 1)
    w['bigru'] = (T * 2 * rows * 2 * 3 * D * D, f * (N * (H + O + 1) * (6 * D + 2 * D) + 6 * 3 * D * D))
    w['gemm_bd'] = add(*[gemm(N * E, D, 2 * D) for E in (H, O, 1)])
    w['gemm_msg'] = add(*[gemm(N * E, D, 2 * D) for E in ((H,) if hh else ()) + (H, O, O, 1)])
    w['frame_msg'] = (N * (H + O) * (H + O) * 2 * D, f * N * ((H + O) * 2 * D + (nkh * H + 2 * H + 2 * O + 1) * D
                                                                + H * (1 + nkh) * D + O * 4 * D))
    w['gate_post'] = (0.0, f * N * (H + O) * 3)
    w['gemm_gs'] = add(gemm(N * H, 3 * D, (1 + nkh) * D), gemm(N * H, 3 * D, (1 + nkh) * D), gemm(N * O, 3 * D, 4 * D),
                       gemm(N * O, 3 * D, 4 * D))
    seg_flops = T * 2 * (2.0 * B * H * 3 * D * (nkh * D + D) + 2.0 * B * O * 3 * D * (2 * D + D)
                         + 2.0 * B * (nkh * H + 2 * O) * D * D)
    seg_bytes = f * (N * (H + O) * (6 * D + 2 * D) + 2 * (3 * D * (nkh * D + D) + 3 * D * 3 * D + (nkh + 3) * D * D))
    w['segment'] = (seg_flops, seg_bytes)
    w['heads'] = (N * H * 4 * 2 * 2 * D * shape.num_classes[0], f * N * H * (2 * 2 * D + 4 * shape.num_classes[0]))
    return w


def run_ours(args, rank, world, local_rank):
    import torch.distributed as dist
    pkg, shape, kwargs = workload(args)
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    torch.manual_seed(0)
    model = pkg.TGGCN(**kwargs).to(dev).eval()
    model.gemm_path = args.gemm_path
    B, T, D = args.B, args.T, args.D
    host = pkg.synth.make_batch(shape, B, T, seed=1234 + rank)
    pinned = {k: host[k].pin_memory() for k in ('x_human', 'x_objects', 'objects_mask')}
    resident = {k: v.to(dev) for k, v in pinned.items()}
    n_calls = T * (shape.H + shape.O)
    noise_dev = pkg.TGGCN.draw_gumbel_noise(n_calls, B).to(dev)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    lib = pkg.abi.lib()

    def step_resident():
        return model(x_human=resident['x_human'], x_objects=resident['x_objects'], objects_mask=resident['objects_mask'])

    pipe = pkg.feeder.DeviceBatchPipeline(dev, pinned)
    out_host = [None, None]                    # two pinned sets: step i's outputs land while step i+1 is being launched
    out_done = [None, None]
    e2e_count = [0]

    def step_e2e():
        # every step: one H2D of a full input batch from pinned memory (it lands in the other slot while this step computes),
        # the per-call CPU Gumbel draw + its H2D, the forward, and the D2H of all outputs into pinned memory.  The host reads
        # step i's outputs after it has launched step i+1 (one step of look-ahead keeps the GPU busy while Python prepares the
        # next launch); the closing synchronize of the timed region covers the last step.
        i = e2e_count[0]
        e2e_count[0] += 1
        xs = pipe.get()
        out = model(x_human=xs['x_human'], x_objects=xs['x_objects'], objects_mask=xs['objects_mask'])
        # the next batch's copy is queued AFTER this forward: the forward's own small H2D (the Gumbel draws) would otherwise wait
        # behind 51 MB on the copy engine; the batch copy still has the whole forward to hide under
        pipe.submit(pinned)
        pipe.release()
        slot = i & 1
        # all outputs (gates + four heads) come back with ONE copy: in inference they are views of one buffer (OutputList.flat)
        if out_host[slot] is None:
            out_host[slot] = torch.empty(out.flat.shape, dtype=out.flat.dtype, pin_memory=True)
        out_host[slot].copy_(out.flat, non_blocking=True)
        out_done[slot] = torch.cuda.Event()
        out_done[slot].record()
        if out_done[slot ^ 1] is not None:
            out_done[slot ^ 1].synchronize()   # the previous step's result is on the host now
            return float(out_host[slot ^ 1][-1])
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        evs = []
        launches0 = lib.tggcn_launch_count()
        for _ in range(steps):
            flush.zero_()                      # evict L2 between timed iterations (outside the event pair)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        launches = lib.tggcn_launch_count() - launches0
        ms = [a.elapsed_time(b) for a, b in evs]
        total = torch.tensor([sum(ms)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.MAX)
        return float(total.item()), ms, launches

    with torch.no_grad():
        model.set_gumbel_noise(noise_dev)
        with ClockSampler(local_rank) as clk:
            total_ms, ms_list, launches = timed(step_resident, args.steps, args.warmup)
        clocks = clk.summary()
        model.check_persistent_kernels()
        model.set_gumbel_noise(None)           # e2e: per-call CPU draw + H2D, like the reference
        pipe.submit(pinned)                    # prime the pipeline: from here on one batch is always in flight
        e2e_total_ms, _, _ = timed(step_e2e, args.steps, max(3, args.warmup))
        # per-stage device time (CUDA events on the launching stream, inside this process)
        model.set_gumbel_noise(noise_dev)
        rows = []
        for i in range(6):
            flush.zero_()
            _, st = model.forward_profile(resident['x_human'], resident['x_objects'], resident['objects_mask'])
            if i:
                rows.append(st)
    train = train_bf16 = None
    if not args.no_train:
        del model
        train = train_step_throughput(args, pkg, shape, kwargs, dev, rank, world, flush, barrier)
        train_bf16 = train_step_throughput(args, pkg, shape, kwargs, dev, rank, world, flush, barrier, precision='bf16')
    others = other_configs(args, pkg, dev, flush) if (world == 1 and not args.no_extras) else None
    frames = world * B * T * args.steps
    value = frames / (total_ms / 1e3)
    e2e_value = frames / (e2e_total_ms / 1e3)
    if rank != 0:
        return
    stages = {k: statistics.median(r[k] for r in rows) for k in rows[0]}
    step_ms = sum(stages.values())
    work = stage_work(shape, B, T, D, shape.hh)
    peaks = load_peaks()
    top = max(stages, key=stages.get)
    flops, nbytes = work[top]
    intensity = flops / max(nbytes, 1.0)
    balance = peaks['tf_sustained'] * 1e12 / (peaks['hbm_gbs'] * 1e9)
    sec = stages[top] / 1e3
    if intensity >= balance or top in ('segment', 'bigru') or top.startswith('gemm'):
        roof = dict(bound='tensor', achieved=flops / sec / 1e12, peak=peaks['tf_sustained'], unit='TFLOP/s')
    else:
        roof = dict(bound='hbm', achieved=nbytes / sec / 1e9, peak=peaks['hbm_gbs'], unit='GB/s')
    roof['frac'] = roof['achieved'] / roof['peak']
    roof['traffic'] = None
    roof.update(kernel=top, kernel_ms=stages[top], share_of_step=stages[top] / step_ms, peak_source=peaks['source'],
                us_per_recurrent_step=(stages[top] * 1e3 / T) if top in ('segment', 'bigru') else None)
    tr = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
    traffic = json.load(open(tr)) if os.path.exists(tr) else {}
    roof['traffic'] = traffic.get(top)
    if top in ('segment', 'bigru'):
        # why the fraction is what it is: a chain of T dependent steps, each two grid barriers + L2 round trips;
        # ncu: tensor pipe 17 % active, 127 MB of DRAM traffic for the whole launch, all weights resident on chip
        roof['limited_by'] = ('latency chain: T dependent recurrent steps (2 grid barriers + L2 round trips each, 3-product fp16-split '
                              'mma.sync on 16-32 activation rows); the on-chip storage of all 148 SMs is full of weights (46 MB as 4-byte words)')
    # the same two ratios for every stage (algorithmic FLOPs and bytes of stage_work over the live stage time): tensor-bound
    # stages are quoted against the sustained bf16 peak although they compute fp32-accurate products (3 MMAs per product)
    per_stage = {}
    for k, ms in stages.items():
        fl, by = work[k]
        if ms <= 0:
            continue
        per_stage[k] = {'ms': round(ms, 4), 'tflops': round(fl / (ms / 1e3) / 1e12, 2), 'gbs': round(by / (ms / 1e3) / 1e9, 1),
                        'frac_tensor': round(fl / (ms / 1e3) / 1e12 / peaks['tf_sustained'], 4),
                        'frac_hbm': round(by / (ms / 1e3) / 1e9 / peaks['hbm_gbs'], 4)}
    h2d = sum(v.numel() * v.element_size() for v in pinned.values()) + n_calls * B * 2 * 4
    d2h = out_host[0].numel() * out_host[0].element_size()      # every output tensor of the forward (counted from the buffer copied)
    line = {
        'metric': METRIC, 'metric_detail': METRIC_DETAIL, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'2G-GCN inference forward, {shape.name.upper()} shape, stage-2 settings', 'videos_per_gpu': B,
                   'frames_per_video': T, 'humans': shape.H, 'objects': shape.O, 'gcn_node': shape.V, 'hidden_size': D,
                   'weights': 'reference default init, torch.manual_seed(0)', 'l2': 'flushed (256 MB memset) between timed steps',
                   'projections': 'TMA-fed tcgen05 kind::f16 on fp16 (hi, lo) operand planes, 3-term split (fp32-class accuracy)' if args.gemm_path else 'fp32 SIMT',
                   'recurrences': 'persistent kernels, on-chip resident weights, mma.sync 3xFP16 split (fp32-class accuracy); '
                                  'large-batch tcgen05 + TMA step kernels from 128 rows per step (segment level) / 192 (BiGRUs) (other_configs)',
                   'parallelism': f'replicas x{world} (videos sharded)'},
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                'ms_per_step': e2e_total_ms / args.steps},
        'gpu_launches': int(launches),
        'roofline': roof,
        'stages_ms': {k: round(v, 4) for k, v in stages.items()},
        'stage_roofline': per_stage,
    }
    if train is not None:
        line['train_step'] = train
        line['train_step_bf16'] = train_bf16         # BASELINE.json configs[2]
    if others is not None:
        line['other_configs'] = others               # BASELINE.json configs[3], [4]
    if world == 1 and not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_port_throughput(args, shape, kwargs, warmup=1, steps=2)
        if train is not None:
            line['cpu_baseline']['train_step'] = cpu_port_train_throughput(args, shape, kwargs)
    print(json.dumps(line), flush=True)


def train_step_throughput(args, pkg, shape, kwargs, dev, rank, world, flush, barrier, precision='fp32'):
    """One training step = forward (BatchNorm in train mode, activations saved) + the fused criterion (drop-in for
    vhoi/losses.py:8, stage-2 weights: BCE on the soft gates + 2 NLL terms) + the hand-written backward (tggcn_backward_ex) +
    gradient all-reduce (N > 1: four buckets launched from the backward's stage boundaries on a side stream, loss terms weighted
    by each rank's share of the valid targets) + Adam, in the sequence of train_utils.py:143-154."""
    import torch.distributed as dist
    torch.manual_seed(0)
    model = pkg.TGGCN(**kwargs).to(dev).train()
    model.gemm_path = args.gemm_path
    model.set_precision(precision)
    # Adam with train.py:40's hyper-parameters: the training driver's one-launch form over the flat parameter / gradient buffers
    # (2g-gcn_b200/optim.py, same update and state_dict as torch.optim.Adam: tests/test_gpu_optim.py); --torch-adam = torch's own
    opt = (torch.optim.Adam(model.parameters(), lr=1e-4, fused=True) if getattr(args, 'torch_adam', False)
           else pkg.optim.FlatAdam(model, lr=1e-4))
    reducer = pkg.dp.GradientAllReduce(model).attach()
    reducer.sync_parameters()
    B, T = args.B, args.T
    host = pkg.synth.make_batch(shape, B, T, seed=1234 + rank)
    x = {k: host[k].to(dev) for k in ('x_human', 'x_objects', 'objects_mask')}
    targets = [t.to(dev) for t in pkg.synth.target_list(shape, pkg.synth.make_targets(shape, host['lengths'], T, seed=77 + rank))]
    model.set_gumbel_noise(pkg.TGGCN.draw_gumbel_noise(T * (shape.H + shape.O), B).to(dev))
    lib = pkg.abi.lib()

    class Cfg(dict):                                   # omegaconf-1.4 style node, as vhoi/losses.py:10 reads it
        def get(self, k, default_value=None):
            return dict.get(self, k, default_value)
    # conf/models/2G-GCN_stage2.yaml:37-54: segmentation BCE on, both label NLLs on
    criterion, _ = pkg.losses.select_loss('2G-GCN', 'multiple', shape.dataset,
                                          Cfg(misc=dict(segmentation_loss=dict(add=True, sigma=4.0, weight=1.0))))

    def step():
        opt.zero_grad(set_to_none=True)
        out = model(**x)
        losses = criterion(out, targets, reduction='mean')
        if world > 1:                                  # each rank's terms weighted by its share of the valid targets
            w = pkg.dp.loss_term_weights(targets)
            losses = [l * wi for l, wi in zip(losses, w.unbind(0))]
        loss = sum(losses)
        loss.backward()                                # queues the backward, then the bucket all-reduces on the side stream
        if world > 1:
            reducer.reduce()                           # joins the streams
        opt.step()
        return loss

    steps, warmup = max(3, min(args.steps, 10)), 3
    for _ in range(warmup):
        step()
    barrier()
    evs = []
    l0 = lib.tggcn_launch_count()
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loss = step()
        e1.record()
        evs.append((e0, e1))
    barrier()
    model.check_persistent_kernels()
    launches = lib.tggcn_launch_count() - l0
    total = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    ms = float(total.item()) / steps

    # the same step end to end: inputs AND targets come from pinned host memory every step (double-buffered on the feeder's
    # side stream), the Gumbel noise is drawn on the CPU per call, and the loss value is read back on the host one step later
    model.set_gumbel_noise(None)
    host_step = {k: host[k].pin_memory() for k in ('x_human', 'x_objects', 'objects_mask')}
    host_tg = pkg.synth.target_list(shape, pkg.synth.make_targets(shape, host['lengths'], T, seed=77 + rank))
    for i, t in enumerate(host_tg):
        host_step[f'target{i}'] = t.pin_memory()
    pipe = pkg.feeder.DeviceBatchPipeline(dev, host_step)
    loss_host = [torch.empty(1, pin_memory=True), torch.empty(1, pin_memory=True)]
    loss_done = [None, None]
    count = [0]

    def step_e2e():
        i = count[0]
        count[0] += 1
        cur = pipe.get()
        opt.zero_grad(set_to_none=True)
        out = model(x_human=cur['x_human'], x_objects=cur['x_objects'], objects_mask=cur['objects_mask'])
        pipe.submit(host_step)                         # after the forward is queued (its noise H2D goes first on the copy engine)
        tg = [cur[f'target{j}'] for j in range(len(host_tg))]
        losses = criterion(out, tg, reduction='mean')
        if world > 1:
            w = pkg.dp.loss_term_weights(tg)
            losses = [l * wi for l, wi in zip(losses, w.unbind(0))]
        loss = sum(losses)
        loss.backward()
        pipe.release()                                 # after the backward: tggcn_backward reads the inputs from the slot
        if world > 1:
            reducer.reduce()
        opt.step()
        loss_host[i & 1].copy_(loss.detach().reshape(1), non_blocking=True)
        loss_done[i & 1] = torch.cuda.Event()
        loss_done[i & 1].record()
        if loss_done[(i & 1) ^ 1] is not None:
            loss_done[(i & 1) ^ 1].synchronize()
            return float(loss_host[(i & 1) ^ 1][0])
        return None

    pipe.submit(host_step)
    for _ in range(warmup):
        step_e2e()
    barrier()
    evs = []
    for _ in range(steps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_e2e()
        e1.record()
        evs.append((e0, e1))
    barrier()
    total = torch.tensor([sum(a.elapsed_time(b) for a, b in evs)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    e2e_ms = float(total.item()) / steps
    h2d = sum(v.numel() * v.element_size() for v in host_step.values()) + T * (shape.H + shape.O) * B * 2 * 4
    return {'value': world * B * T / (ms / 1e3), 'unit': UNIT, 'ms_per_step': ms, 'steps': steps, 'warmup': warmup,
            'e2e': {'value': world * B * T / (e2e_ms / 1e3), 'unit': UNIT, 'ms_per_step': e2e_ms, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': 4},
            'gpu_launches_per_step': int(launches // steps), 'loss': float(loss.detach()),
            'dtype': 'bf16' if precision == 'bf16' else 'f32',
            'what': 'forward(train-mode BN, saves) + fused criterion (BCE + 2x NLL) + tggcn_backward_ex + '
                    + ('bucketed NCCL all-reduce of the flat gradient overlapped with the backward + ' if world > 1 else '')
                    + ('torch.optim.Adam(fused=True) step; ' if getattr(args, 'torch_adam', False) else 'Adam step (one launch over the flat buffers, optim.FlatAdam); ')
                    + ('bf16 operands / fp32 accumulation in every projection and weight-gradient GEMM, fp32 gates, recurrences and optimiser'
                       if precision == 'bf16' else 'fp32-accurate products (3xTF32 / 3xFP16 split)'),
            'global_batch_videos': world * B}


def other_configs(args, pkg, dev, flush):
    """BASELINE.json configs[3] and [4] on this GPU (N = 1): the CAD-120-shaped inference sweep over the batch size (T = 512) and
    the Bimanual-shaped training step (B = 32, T = 256, hidden 64 as shipped and 512).  Each entry: frames/s (CUDA events, median
    of 3 after 2 warm-ups, L2 flushed between iterations) and the roofline of its largest stage from the stage_work closed forms."""
    peaks = load_peaks()
    res = {'cad120_inference_sweep': [], 'bimanual_training': []}

    def med(fn, iters=3, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        ms = []
        for _ in range(iters):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize(dev)
            ms.append(e0.elapsed_time(e1))
        return statistics.median(ms)

    cad = pkg.synth.SHAPES['cad120']
    T = 512
    sweep = [b for b in (8, 16, 32, 64, 128, 256) if b <= args.max_sweep_batch]
    big = pkg.synth.make_batch(cad, max(sweep), T, seed=1234)                 # generated once, sliced for the smaller batches
    torch.manual_seed(0)
    model = pkg.TGGCN(**pkg.synth.model_kwargs(cad, hidden_size=512, stage=2)).to(dev).eval()
    for b in sweep:
        x = {k: big[k][:b].to(dev) for k in ('x_human', 'x_objects', 'objects_mask')}
        model.set_gumbel_noise(pkg.TGGCN.draw_gumbel_noise(T * (cad.H + cad.O), b).to(dev))
        with torch.no_grad():
            ms = med(lambda: model(**x))
            _, st = model.forward_profile(x['x_human'], x['x_objects'], x['objects_mask'])
        model.check_persistent_kernels()
        top = max(st, key=st.get)
        fl, by = stage_work(cad, b, T, 512, cad.hh)[top]
        rows = b * max(cad.H, cad.O)
        res['cad120_inference_sweep'].append({
            'videos': b, 'frames_per_video': T, 'ms': round(ms, 3), 'frames_per_s': round(b * T / ms * 1e3),
            'recurrent_path': ('large-batch tcgen05 + TMA step kernels' if rows >= 192 else
                               'segment level: step kernels; BiGRUs: cluster / resident kernels' if rows >= 128 else 'persistent latency path'),
            'top_stage': top, 'top_stage_ms': round(st[top], 3), 'us_per_recurrent_step': round(st[top] * 1e3 / T, 2) if top in ('segment', 'bigru') else None,
            'roofline': {'bound': 'tensor', 'achieved': round(fl / (st[top] / 1e3) / 1e12, 2), 'peak': peaks['tf_sustained'], 'unit': 'TFLOP/s',
                         'frac': round(fl / (st[top] / 1e3) / 1e12 / peaks['tf_sustained'], 4)}})
        del x
    del model, big
    torch.cuda.empty_cache()

    class Cfg(dict):
        def get(self, k, default_value=None):
            return dict.get(self, k, default_value)
    bim = pkg.synth.SHAPES['bimanual']
    Bb, Tb = 32, 256
    host = pkg.synth.make_batch(bim, Bb, Tb, seed=1234)
    x = {k: host[k].to(dev) for k in ('x_human', 'x_objects', 'objects_mask')}
    targets = [t.to(dev) for t in pkg.synth.target_list(bim, pkg.synth.make_targets(bim, host['lengths'], Tb, seed=5))]
    criterion, _ = pkg.losses.select_loss('2G-GCN', 'multiple', bim.dataset, Cfg(misc=dict(segmentation_loss=dict(add=True, sigma=4.0, weight=1.0))))
    for D in (64, 512):
        torch.manual_seed(0)
        m = pkg.TGGCN(**pkg.synth.model_kwargs(bim, hidden_size=D, stage=2)).to(dev).train()
        opt = pkg.optim.FlatAdam(m, lr=1e-4)
        m.set_gumbel_noise(pkg.TGGCN.draw_gumbel_noise(Tb * (bim.H + bim.O), Bb).to(dev))

        def step():
            opt.zero_grad(set_to_none=True)
            sum(criterion(m(**x), targets, reduction='mean')).backward()
            opt.step()
        ms = med(step)
        m.check_persistent_kernels()
        fl = 3.0 * sum(v[0] for v in stage_work(bim, Bb, Tb, D, bim.hh).values())          # forward + 2x backward
        res['bimanual_training'].append({
            'videos': Bb, 'frames_per_video': Tb, 'hidden_size': D, 'ms': round(ms, 3), 'frames_per_s': round(Bb * Tb / ms * 1e3),
            'roofline': {'bound': 'tensor', 'achieved': round(fl / (ms / 1e3) / 1e12, 2), 'peak': peaks['tf_sustained'], 'unit': 'TFLOP/s',
                         'frac': round(fl / (ms / 1e3) / 1e12 / peaks['tf_sustained'], 4), 'of': 'whole training step'}})
        del m, opt
    torch.cuda.empty_cache()
    return res


def cpu_port_throughput(args, shape, kwargs, warmup, steps):
    """The oracle port of the reference's CPU path, timed on this box's host cores (whole workload per step)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import tggcn_oracle as orc
    pkg = importlib.import_module('2g-gcn_b200')
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    params = {k: v.detach() for k, v in pkg.TGGCN(**kwargs).state_dict().items()}
    B, T = args.B, args.T
    batch = pkg.synth.make_batch(shape, B, T, seed=1234)
    cfg = orc.OracleConfig(args.D, shape.V, shape.num_classes, shape.hh, True, kwargs['update_segment_threshold'])
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            orc.forward(params, cfg, batch['x_human'], batch['x_objects'], batch['objects_mask'], None, None, None)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {'value': B * T / sec, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'whole workload (B={B}, T={T}, hidden {args.D}), {steps} timed forwards after {warmup} warm-up, '
                      f'{cores} torch threads', 'seconds_per_step': sec}


def cpu_port_train_throughput(args, shape, kwargs, T_sample=None, steps=1):
    """CPU train step of the reference path (oracle port: train-mode forward + criterion + autograd backward + Adam) on the
    WHOLE workload (all B videos, all T frames; ~5-10 s per step on 16 cores), one timed step after one warm-up."""
    T_sample = args.T if T_sample is None else T_sample
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import tggcn_oracle as orc
    pkg = importlib.import_module('2g-gcn_b200')
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    sd = pkg.TGGCN(**kwargs).state_dict()
    params = {k: (v.detach().clone().requires_grad_(True) if (v.is_floating_point() and 'running' not in k) else v.clone())
              for k, v in sd.items()}
    opt = torch.optim.Adam([v for v in params.values() if v.requires_grad], lr=1e-4)
    B, T = args.B, min(T_sample, args.T)
    batch = pkg.synth.make_batch(shape, B, T, seed=1234)
    targets = pkg.synth.target_list(shape, pkg.synth.make_targets(shape, batch['lengths'], T, seed=77))
    cfg = orc.OracleConfig(args.D, shape.V, shape.num_classes, shape.hh, True, kwargs['update_segment_threshold'])
    times = []
    for i in range(1 + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        out = orc.forward(params, cfg, batch['x_human'], batch['x_objects'], batch['objects_mask'], None, None, None, training=True)
        loss = sum(orc.multi_task_loss(out, targets, shape.dataset, 2))
        loss.backward()
        opt.step()
        if i >= 1:
            times.append(time.perf_counter() - t0)
    sec = sum(times) / len(times)
    return {'value': B * T / sec, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'seconds_per_step': sec,
            'sample': f'whole workload: train step on B={B} videos x T={T} frames (hidden {args.D}), {steps} timed step(s) after 1 warm-up, '
                      f'{cores} torch threads'}


def run_reference(args, rank, world):
    if rank != 0:
        return
    pkg, shape, kwargs = workload(args)
    ran_steps, ran_warmup = min(args.steps, 3), min(args.warmup, 1)     # ~1.2 s per forward on 16 cores: bounded so the arm ends in seconds
    res = cpu_port_throughput(args, shape, kwargs, warmup=ran_warmup, steps=ran_steps)
    line = {
        'impl': 'reference', 'metric': METRIC, 'metric_detail': METRIC_DETAIL, 'value': res['value'], 'unit': UNIT, 'n_gpus': world, 'steps': ran_steps,
        'warmup': ran_warmup, 'steps_requested': args.steps, 'warmup_requested': args.warmup, 'ms_per_step': res['seconds_per_step'] * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'2G-GCN inference forward, {shape.name.upper()} shape, stage-2 settings', 'videos_per_gpu': args.B,
                   'frames_per_video': args.T, 'humans': shape.H, 'objects': shape.O, 'gcn_node': shape.V, 'hidden_size': args.D,
                   'note': 'CPU port of the reference path (oracle); steps / warmup are what was run (timed steps capped at 3)'},
        'cpu_baseline': {k: res[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
        'e2e': {'value': res['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    if not args.no_train:
        line['train_step'] = cpu_port_train_throughput(args, shape, kwargs)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--shape', default='mphoi')
    ap.add_argument('--B', type=int, default=8)
    ap.add_argument('--T', type=int, default=128)
    ap.add_argument('--D', type=int, default=512)
    ap.add_argument('--gemm-path', type=int, default=2)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train', action='store_true', help='skip the training-step measurement')
    ap.add_argument('--torch-adam', action='store_true', help='training step with torch.optim.Adam(fused=True) instead of optim.FlatAdam')
    ap.add_argument('--max-sweep-batch', type=int, default=256, help='largest batch of the CAD-120 sweep in other_configs (46 GB at 256)')
    ap.add_argument('--no-extras', action='store_true', help='skip the CAD-120 sweep and the Bimanual training configs (other_configs)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == '__main__':
    main()

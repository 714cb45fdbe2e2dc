"""BASELINE.json configs[3] and [4] on one B200: the CAD-120-shaped inference sweep over the batch size and the
Bimanual-shaped training step, both on synthetic tensors, timed with CUDA events (median of `--iters` after 3 warm-ups).

    python tools/sweep_configs.py [--what cad120,bimanual,mphoi] [--iters 5] [--max-gb 100] > profiles/rNN_sweep.txt

A case whose inputs + workspace would exceed --max-gb of device memory is reported as skipped instead of being attempted.
"""
import argparse
import ctypes as C
import importlib
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('2g-gcn_b200')
abi = pkg.abi


def dims_for(model, shape, B, T, train):
    n_sub, n_aff = shape.num_classes
    return abi.Dims(B=B, T=T, H=shape.H, O=shape.O, V=shape.V, D=model.hidden_size, Fh=shape.Fh, C_sub=n_sub,
                    C_aff=0 if n_aff is None else n_aff, hh=int(shape.hh), filter=int(model.filter_discrete_updates),
                    bn_train=int(train), human_seg_given=0, object_seg_given=0, inspect=0, persistent=1, gemm_path=2,
                    thr=model.update_segment_threshold, save_for_backward=int(train))


def footprint_gb(model, shape, B, T, train):
    d = dims_for(model, shape, B, T, train)
    ws = abi.workspace_bytes(d)
    if train:
        ws += int(abi.lib().tggcn_backward_workspace_bytes(C.byref(d)))
    inputs = 4 * B * T * (shape.H * shape.Fh + shape.O * 2048)
    return (ws + inputs) / 2 ** 30


def time_ms(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return statistics.median(ms)


def inference_case(shape, B, T, D, iters, max_gb):
    torch.manual_seed(0)
    model = pkg.TGGCN(**pkg.synth.model_kwargs(shape, hidden_size=D, stage=2)).cuda().eval()
    gb = footprint_gb(model, shape, B, T, False)
    tag = f'{shape.name:9s} inference B={B:<4d} T={T:<4d} D={D:<4d}'
    if gb > max_gb:
        print(f'{tag} skipped: inputs + workspace = {gb:.1f} GB > {max_gb} GB', flush=True)
        return
    batch = pkg.synth.make_batch(shape, B, T, seed=1234)
    x = {k: batch[k].cuda() for k in ('x_human', 'x_objects', 'objects_mask')}
    model.set_gumbel_noise(pkg.TGGCN.draw_gumbel_noise(T * (shape.H + shape.O), B).cuda())
    with torch.no_grad():
        ms = time_ms(lambda: model(**x), iters)
        _, st = model.forward_profile(x['x_human'], x['x_objects'], x['objects_mask'])
    model.check_persistent_kernels()
    top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f'{tag} {ms:9.3f} ms  {B * T / ms * 1e3:12,.0f} frames/s  {gb:6.1f} GB   top: '
          + ', '.join(f'{k} {v:.2f}' for k, v in top), flush=True)
    del model, x
    torch.cuda.empty_cache()


def train_case(shape, B, T, D, iters, max_gb):
    torch.manual_seed(0)
    model = pkg.TGGCN(**pkg.synth.model_kwargs(shape, hidden_size=D, stage=2)).cuda().train()
    gb = footprint_gb(model, shape, B, T, True)
    tag = f'{shape.name:9s} train     B={B:<4d} T={T:<4d} D={D:<4d}'
    if gb > max_gb:
        print(f'{tag} skipped: inputs + workspaces = {gb:.1f} GB > {max_gb} GB', flush=True)
        return
    opt = pkg.optim.FlatAdam(model, lr=1e-4)
    batch = pkg.synth.make_batch(shape, B, T, seed=1234)
    x = {k: batch[k].cuda() for k in ('x_human', 'x_objects', 'objects_mask')}
    targets = [t.cuda() for t in pkg.synth.target_list(shape, pkg.synth.make_targets(shape, batch['lengths'], T, seed=5))]
    model.set_gumbel_noise(pkg.TGGCN.draw_gumbel_noise(T * (shape.H + shape.O), B).cuda())

    class Cfg(dict):
        def get(self, k, default_value=None):
            return dict.get(self, k, default_value)
    criterion, _ = pkg.losses.select_loss('2G-GCN', 'multiple', shape.dataset,
                                          Cfg(misc=dict(segmentation_loss=dict(add=True, sigma=4.0, weight=1.0))))

    def step():
        opt.zero_grad(set_to_none=True)
        loss = sum(criterion(model(**x), targets, reduction='mean'))
        loss.backward()
        opt.step()
    ms = time_ms(step, iters)
    model.check_persistent_kernels()
    print(f'{tag} {ms:9.3f} ms  {B * T / ms * 1e3:12,.0f} frames/s  {gb:6.1f} GB   (forward + fused criterion + backward + Adam)', flush=True)
    del model, opt, x, targets
    torch.cuda.empty_cache()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--what', default='cad120,bimanual,mphoi')
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--max-gb', type=float, default=100.0)
    ap.add_argument('--batches', default='8,16,32,64,128,256', help='batch sizes of the CAD-120 sweep')
    a = ap.parse_args()
    what = a.what.split(',')
    print(f'# TGGCN_RECURRENT_MODE={os.environ.get("TGGCN_RECURRENT_MODE", "0")} (0: by rows per step, 1: latency path, 2: large-batch path)')
    print(f'# {torch.cuda.get_device_name(0)}; median of {a.iters} timed runs after 3 warm-ups, CUDA events; synthetic tensors (2g-gcn_b200/synth.py)')
    if 'cad120' in what:       # BASELINE.json configs[3]: CAD-120 shape (1 human, 5 objects, long sequences), inference sweep bs 8 -> 256
        for B in [int(b) for b in a.batches.split(',')]:
            inference_case(pkg.synth.SHAPES['cad120'], B, 512, 512, a.iters, a.max_gb)
    if 'bimanual' in what:     # configs[4]: Bimanual shape (2 hands, 9 objects), training; D = 64 is the shipped yaml comment, D = 512 the others' size
        for D in (64, 512):
            train_case(pkg.synth.SHAPES['bimanual'], 32, 256, D, a.iters, a.max_gb)
            inference_case(pkg.synth.SHAPES['bimanual'], 32, 256, D, a.iters, a.max_gb)
    if 'mphoi' in what:        # the headline shape at larger batches (the recurrences are latency-bound at B = 8)
        for B in (8, 32, 128):
            inference_case(pkg.synth.SHAPES['mphoi'], B, 128, 512, a.iters, a.max_gb)
        train_case(pkg.synth.SHAPES['mphoi'], 8, 128, 512, a.iters, a.max_gb)


if __name__ == '__main__':
    main()

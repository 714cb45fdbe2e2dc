"""Run a few forwards (or training steps) of one synthetic configuration — the command line ncu wraps.

    python tools/run_once.py --shape cad120 --B 64 --T 16 --D 512 [--mode 2] [--train] [--iters 2] [--precision bf16]
"""
import argparse
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('2g-gcn_b200')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--shape', default='mphoi')
    ap.add_argument('--B', type=int, default=8)
    ap.add_argument('--T', type=int, default=128)
    ap.add_argument('--D', type=int, default=512)
    ap.add_argument('--mode', type=int, default=0)
    ap.add_argument('--iters', type=int, default=2)
    ap.add_argument('--train', action='store_true')
    ap.add_argument('--precision', default='fp32')
    a = ap.parse_args()
    shape = pkg.synth.SHAPES[a.shape]
    torch.manual_seed(0)
    model = pkg.TGGCN(**pkg.synth.model_kwargs(shape, hidden_size=a.D, stage=2)).cuda()
    model.recurrent_mode = a.mode
    model.set_precision(a.precision)
    model.train(a.train)
    batch = pkg.synth.make_batch(shape, a.B, a.T, seed=1234)
    x = {k: batch[k].cuda() for k in ('x_human', 'x_objects', 'objects_mask')}
    model.set_gumbel_noise(pkg.TGGCN.draw_gumbel_noise(a.T * (shape.H + shape.O), a.B).cuda())
    targets = [t.cuda() for t in pkg.synth.target_list(shape, pkg.synth.make_targets(shape, batch['lengths'], a.T, seed=5))]

    class Cfg(dict):
        def get(self, k, default_value=None):
            return dict.get(self, k, default_value)
    criterion, _ = pkg.losses.select_loss('2G-GCN', 'multiple', shape.dataset,
                                          Cfg(misc=dict(segmentation_loss=dict(add=True, sigma=4.0, weight=1.0))))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
    for i in range(a.iters):
        ev[i].record()
        if a.train:
            model.zero_grad(set_to_none=True)
            sum(criterion(model(**x), targets, reduction='mean')).backward()
        else:
            with torch.no_grad():
                model(**x)
    ev[a.iters].record()
    torch.cuda.synchronize()
    model.check_persistent_kernels()
    print('ms per iteration:', [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(a.iters)])


if __name__ == '__main__':
    main()

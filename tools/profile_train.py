"""Times the phases of one training step (forward with saves, criterion, hand-written backward, Adam) with CUDA events.

    python tools/profile_train.py [--B 8 --T 128 --D 512 --shape mphoi --iters 5]
"""
import argparse
import importlib
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--shape', default='mphoi')
    ap.add_argument('--B', type=int, default=8)
    ap.add_argument('--T', type=int, default=128)
    ap.add_argument('--D', type=int, default=512)
    ap.add_argument('--stage', type=int, default=2)
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--torch-criterion', action='store_true', help='plain torch ops (what the unchanged reference criterion launches)')
    args = ap.parse_args()
    pkg = importlib.import_module('2g-gcn_b200')
    shape = pkg.synth.SHAPES[args.shape]
    kwargs = pkg.synth.model_kwargs(shape, hidden_size=args.D, stage=args.stage)
    dev = torch.device('cuda', 0)
    torch.manual_seed(0)
    model = pkg.TGGCN(**kwargs).to(dev).train()
    opt = pkg.optim.FlatAdam(model, lr=1e-4) if os.environ.get('TGGCN_TORCH_ADAM') != '1' else torch.optim.Adam(model.parameters(), lr=1e-4)
    B, T = args.B, args.T
    batch = pkg.synth.make_batch(shape, B, T, seed=1234)
    x = {k: batch[k].to(dev) for k in ('x_human', 'x_objects', 'objects_mask')}
    targets = [t.to(dev) for t in pkg.synth.target_list(shape, pkg.synth.make_targets(shape, batch['lengths'], T, seed=5))]
    hseg = torch.ones(B, T, shape.H, device=dev) if args.stage == 1 else None
    n_s = T * ((0 if args.stage == 1 else shape.H) + shape.O)
    noise = pkg.TGGCN.draw_gumbel_noise(n_s, B).to(dev)
    model.set_gumbel_noise(noise)
    class Cfg(dict):
        def get(self, k, default_value=None):
            return dict.get(self, k, default_value)
    misc = dict(segmentation_loss=dict(add=args.stage == 2, sigma=4.0, weight=1.0))
    criterion, _ = pkg.losses.select_loss('2G-GCN', 'multiple', shape.dataset, Cfg(misc=misc))
    if args.torch_criterion:
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import tggcn_oracle as orc
        criterion = lambda o, t, reduction='mean': orc.multi_task_loss(o, t, shape.dataset, args.stage)
    rows = []
    for it in range(args.iters + 2):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        opt.zero_grad(set_to_none=True)
        ev[0].record()
        out = model(human_segmentation=hseg, **x)
        ev[1].record()
        loss = sum(criterion(out, targets, reduction='mean'))
        ev[2].record()
        loss.backward()
        ev[3].record()
        opt.step()
        ev[4].record()
        torch.cuda.synchronize()
        if it >= 2:
            rows.append([ev[i].elapsed_time(ev[i + 1]) for i in range(4)])
    model.check_persistent_kernels()
    names = ['forward(save)', 'criterion', 'backward', 'adam']
    med = [statistics.median(r[i] for r in rows) for i in range(4)]
    for n, m in zip(names, med):
        print(f'{n:16s} {m:9.3f} ms')
    tot = sum(med)
    print(f'{"train step":16s} {tot:9.3f} ms   {B * T / tot * 1e3:10.0f} frames/s   loss {float(loss):.5f}')


if __name__ == '__main__':
    main()

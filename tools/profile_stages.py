"""Per-stage device time of one forward (CUDA events around each stage, via tggcn_forward_profile).

    python tools/profile_stages.py [--shape mphoi] [--B 8] [--T 128] [--D 512] [--stage 2] [--iters 5]
"""
import argparse
import ctypes as C
import importlib
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module('2g-gcn_b200')


def stage_times(model, batch, iters=5, hseg=None, oseg=None):
    """Returns (names, median ms per stage, median total ms)."""
    rows = []
    with torch.no_grad():
        for i in range(iters + 1):
            _, ms = model.forward_profile(batch['x_human'], batch['x_objects'], batch['objects_mask'], hseg, oseg)
            if i:
                rows.append([ms[n] for n in pkg.abi.STAGE_NAMES])
    t = torch.tensor(rows)
    return pkg.abi.STAGE_NAMES, t.median(dim=0).values.tolist(), float(t.sum(dim=1).median())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--shape', default='mphoi')
    ap.add_argument('--B', type=int, default=8)
    ap.add_argument('--T', type=int, default=128)
    ap.add_argument('--D', type=int, default=512)
    ap.add_argument('--stage', type=int, default=2)
    ap.add_argument('--iters', type=int, default=5)
    ap.add_argument('--per-step', action='store_true', help='one launch per recurrent step instead of persistent kernels')
    ap.add_argument('--gemm-path', type=int, default=2)
    a = ap.parse_args()
    shape = pkg.synth.SHAPES[a.shape]
    torch.manual_seed(0)
    model = pkg.TGGCN(**pkg.synth.model_kwargs(shape, hidden_size=a.D, stage=a.stage)).cuda().eval()
    model.persistent_kernels = not a.per_step
    model.gemm_path = a.gemm_path
    batch = {k: v.cuda() for k, v in pkg.synth.make_batch(shape, a.B, a.T).items()}
    hseg = torch.ones(a.B, a.T, shape.H, device='cuda') if a.stage == 1 else None
    oseg = torch.ones(a.B, a.T, shape.O, device='cuda') if (a.stage == 1 and shape.dataset == 'cad120') else None
    names, med, total = stage_times(model, batch, a.iters, hseg, oseg)
    model.check_persistent_kernels()
    frames = a.B * a.T
    print(f'{a.shape} B={a.B} T={a.T} D={a.D} stage={a.stage} persistent={not a.per_step}: '
          f'{total:.3f} ms/forward -> {frames / total * 1e3:,.0f} frames/s')
    for n, m in zip(names, med):
        print(f'  {n:12s} {m:9.3f} ms  {100 * m / total:5.1f}%')


if __name__ == '__main__':
    main()

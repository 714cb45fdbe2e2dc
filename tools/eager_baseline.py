"""The "same box" bar of SURVEY.md §8d: the reference's algorithm as plain eager PyTorch ON THE GPU (the oracle port run with CUDA
tensors — thousands of small library kernels per forward, the way the unmodified reference would run with device='cuda'),
next to the hand-written path, same tensors.  Measurement aid only; nothing in the product path imports the oracle.

    python tools/eager_baseline.py [--B 8 --T 128 --D 512 --shape mphoi] > profiles/rNN_eager_gpu_baseline.txt
"""
import argparse
import importlib
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import tggcn_oracle as orc  # noqa: E402

pkg = importlib.import_module('2g-gcn_b200')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--shape', default='mphoi')
    ap.add_argument('--B', type=int, default=8)
    ap.add_argument('--T', type=int, default=128)
    ap.add_argument('--D', type=int, default=512)
    ap.add_argument('--iters', type=int, default=3)
    a = ap.parse_args()
    shape = pkg.synth.SHAPES[a.shape]
    kw = pkg.synth.model_kwargs(shape, hidden_size=a.D, stage=2)
    torch.manual_seed(0)
    model = pkg.TGGCN(**kw).cuda().eval()
    batch = {k: v.cuda() for k, v in pkg.synth.make_batch(shape, a.B, a.T).items() if torch.is_tensor(v)}
    noise = pkg.TGGCN.draw_gumbel_noise(a.T * (shape.H + shape.O), a.B).cuda()
    params = {k: v.detach() for k, v in model.state_dict().items()}
    cfg = orc.OracleConfig(a.D, shape.V, shape.num_classes, shape.hh, True, kw['update_segment_threshold'])

    def eager():
        with torch.no_grad():
            return orc.forward(params, cfg, batch['x_human'], batch['x_objects'], batch['objects_mask'], None, None, noise)

    def ours():
        model.set_gumbel_noise(noise)
        with torch.no_grad():
            return model(x_human=batch['x_human'], x_objects=batch['x_objects'], objects_mask=batch['objects_mask'])

    def wall_ms(fn, n):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(n):
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
        return statistics.median(ts)

    ref, got = eager(), ours()
    worst = max(float((r - g).abs().max()) for r, g in zip(ref, got))
    e_ms, o_ms = wall_ms(eager, a.iters), wall_ms(ours, 20)
    frames = a.B * a.T
    print(f'# {torch.cuda.get_device_name(0)}, {a.shape} B={a.B} T={a.T} D={a.D} stage-2 settings, wall clock incl. launch overhead, median')
    print(f'eager PyTorch port of the reference algorithm on the GPU : {e_ms:9.2f} ms / forward  {frames / e_ms * 1e3:12,.0f} frames/s')
    print(f'hand-written sm_100a path (this repo)                    : {o_ms:9.2f} ms / forward  {frames / o_ms * 1e3:12,.0f} frames/s')
    print(f'ratio {e_ms / o_ms:.1f}x; max |difference| of the outputs on the same tensors and noise: {worst:.2e}')


if __name__ == '__main__':
    main()

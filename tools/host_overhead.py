"""Host-side enqueue time of a forward / training step against its device time (is the step launch-bound?).

    python tools/host_overhead.py [--train]
"""
import argparse, importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('2g-gcn_b200')

ap = argparse.ArgumentParser()
ap.add_argument('--train', action='store_true')
ap.add_argument('--profile', action='store_true')
a = ap.parse_args()
shape = pkg.synth.SHAPES['mphoi']
B, T, D = 8, 128, 512
torch.manual_seed(0)
model = pkg.TGGCN(**pkg.synth.model_kwargs(shape, hidden_size=D, stage=2)).cuda()
model.train(a.train)
# fixed device-resident noise: the default path's pinned ring makes the host wait for the forward three calls back (back-pressure,
# not enqueue cost)
model.set_gumbel_noise(pkg.TGGCN.draw_gumbel_noise(T * (shape.H + shape.O), B).cuda())
batch = pkg.synth.make_batch(shape, B, T, seed=1234)
x = {k: batch[k].cuda() for k in ('x_human', 'x_objects', 'objects_mask')}
targets = [t.cuda() for t in pkg.synth.target_list(shape, pkg.synth.make_targets(shape, batch['lengths'], T, seed=5))]


class Cfg(dict):
    def get(self, k, default_value=None):
        return dict.get(self, k, default_value)


criterion, _ = pkg.losses.select_loss('2G-GCN', 'multiple', shape.dataset, Cfg(misc=dict(segmentation_loss=dict(add=True, sigma=4.0, weight=1.0))))


def step():
    t0 = time.perf_counter()
    if a.train:
        model.zero_grad(set_to_none=True)
        out = model(**x)
        t1 = time.perf_counter()
        loss = sum(criterion(out, targets, reduction='mean'))
        t2 = time.perf_counter()
        loss.backward()
        t3 = time.perf_counter()
        return t1 - t0, t2 - t1, t3 - t2
    with torch.no_grad():
        model(**x)
    return (time.perf_counter() - t0,)


for _ in range(5):
    step()
torch.cuda.synchronize()
rows = []
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
if a.profile:
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable()
for _ in range(20):
    rows.append(step())
if a.profile:
    pr.disable()
ev1.record()
torch.cuda.synchronize()
names = ('forward', 'criterion', 'backward') if a.train else ('forward',)
for i, n in enumerate(names):
    print(f'host enqueue {n:10s} {1e3 * sum(r[i] for r in rows) / len(rows):7.3f} ms')
print(f'device time per step      {ev0.elapsed_time(ev1) / 20:7.3f} ms')
if a.profile:
    pstats.Stats(pr).sort_stats('cumulative').print_stats(25)

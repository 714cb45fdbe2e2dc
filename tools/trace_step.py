"""Timing experiment (build with TGGCN_NVCC_DEFS=-DST_TRACE): clock64 stamps of CTA 0 of the LAST step_tc_kernel launch.
    TGGCN_NVCC_DEFS=-DST_TRACE python tools/trace_step.py --shape cad120 --B 64 --T 4 [--bigru-only]"""
import argparse, ctypes as C, importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = importlib.import_module('2g-gcn_b200')
ap = argparse.ArgumentParser()
ap.add_argument('--shape', default='cad120'); ap.add_argument('--B', type=int, default=64); ap.add_argument('--T', type=int, default=4)
ap.add_argument('--D', type=int, default=512)
ap.add_argument('--kind', type=int, default=0, help='0 = cell GEMM, 1 = BiGRU step, 2 = message MLPs')
a = ap.parse_args()
shape = pkg.synth.SHAPES[a.shape]
torch.manual_seed(0)
model = pkg.TGGCN(**pkg.synth.model_kwargs(shape, hidden_size=a.D, stage=2)).cuda().eval()
model.recurrent_mode = 2
batch = pkg.synth.make_batch(shape, a.B, a.T, seed=1)
x = {k: batch[k].cuda() for k in ('x_human', 'x_objects', 'objects_mask')}
lib = pkg.abi.lib()
assert lib.tggcn_debug_trace_kind(a.kind) == 0
for it in range(3):
    with torch.no_grad():
        model(**x)
    torch.cuda.synchronize()
buf = (C.c_longlong * 64)()
assert lib.tggcn_debug_trace(buf) == 0
t = list(buf)
t0 = t[0]
print('kind', a.kind, 'shape', a.shape, 'B', a.B, 'TGGCN_STEP_CG', os.environ.get('TGGCN_STEP_CG', '2'))
print('setup done      ', t[1] - t0)
print('producer issue  ', [v - t0 for v in t[4:32] if v > t0])
print('mma sees full   ', [v - t0 for v in t[32:60] if v > t0])
print('epilogue starts ', t[2] - t0)
print('epilogue ends   ', t[60] - t0, '(warp 2)')
print('all done        ', t[3] - t0)

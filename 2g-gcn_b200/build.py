"""Builds lib2ggcn_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python 2g-gcn_b200/build.py [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels to the GPU box with the
repo snapshot.  No torch dependency: the library's interface is include/tggcn_b200.h only.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'lib2ggcn_b200.so')
STAMP = LIB + '.srchash'
SOURCES = ['api.cu', 'api_bwd.cu', 'backward.cu', 'recurrent_bwd.cu', 'geo_gcn_bwd.cu', 'geo_gcn.cu', 'gemm_simt.cu', 'gemm_tc.cu', 'gemm16.cu', 'gemm_bwd.cu', 'bigru.cu', 'bigru_res.cu', 'bigru_cl.cu', 'bigru_bwd.cu', 'segment.cu', 'step_tc.cu', 'frame.cu', 'loss.cu', 'evaluate.cu', 'optim.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr',
              '-Xptxas', '-v'] + os.environ.get('TGGCN_NVCC_DEFS', '').split()       # e.g. "-DSEG_NO_SHADOW" for A/B experiments


def _source_hash() -> str:
    h = hashlib.sha256()
    files = sorted(os.listdir(CSRC)) + ['../../include/tggcn_b200.h']
    for f in files:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, 'rb').read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def build(force: bool = False, verbose: bool = False) -> str:
    want = _source_hash()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == want:
        return LIB
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, 'build', src.replace('.cu', '.o'))
        objs.append(obj)
        cmd = [nvcc_path()] + NVCC_FLAGS + ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f'--- {src} ---\n{out}')
        failed |= p.returncode != 0
    with open(os.path.join(HERE, 'build', 'nvcc.log'), 'w') as f:
        f.write('\n'.join(log))
    if failed or verbose:
        sys.stderr.write('\n'.join(log))
    if failed:
        raise RuntimeError('nvcc failed; see 2g-gcn_b200/build/nvcc.log')
    cmd = [nvcc_path(), '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcuda']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        # libcuda stubs may be absent in a GPU-less container; the driver API is resolved at run time anyway
        cmd = [c for c in cmd if c != '-lcuda']
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError('link failed:\n' + r.stdout)
    with open(STAMP, 'w') as f:
        f.write(want)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))

"""Build recipe for liblarvanet_b200.so (sm_100a only, in-tree so the .so travels with `gpurun`).

    python -m larvanet_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  `-lineinfo` keeps ncu's source page usable; `-Xptxas -v` output is written to
`larvanet_b200/csrc/build/ptxas.log` for register / spill / shared-memory review.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
BUILD = os.path.join(CSRC, 'build')
LIB = os.path.join(HERE, 'liblarvanet_b200.so')
SOURCES = ['capi.cu', 'conv_tc.cu', 'conv_row.cu', 'conv_row_cp.cu', 'conv_chain.cu', 'conv_simt.cu', 'wgrad.cu', 'head.cu', 'layout.cu']
# measured-slower experiments of round 1 (cluster-resident strips, ky-stacked tiles with a shuffle epilogue) live in
# tools/experiments/ and are only compiled into the library with LARVANET_B200_EXPERIMENTAL=1 (adds -DLV_EXPERIMENTAL)
EXPERIMENTAL = os.environ.get('LARVANET_B200_EXPERIMENTAL', '0') == '1'
EXPERIMENT_DIR = os.path.join(os.path.dirname(HERE), 'tools', 'experiments')
EXPERIMENT_SOURCES = ['conv_strip.cu', 'conv_tc_ky.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17', '-lineinfo',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v', '--expt-relaxed-constexpr']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; larvanet_b200 needs the CUDA 12.9 toolkit to build')


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h'))]
    headers.append(os.path.join(os.path.dirname(HERE), 'include', 'larvanet_b200.h'))
    objs, jobs = [], []
    flags = list(NVCC_FLAGS)
    tag = ''
    sources = [os.path.join(CSRC, f) for f in SOURCES]
    if EXPERIMENTAL:
        flags += ['-DLV_EXPERIMENTAL', '-I', CSRC]
        tag = '.exp'                       # separate objects: the flag changes capi.cu / conv_chain.cu
        sources += [os.path.join(EXPERIMENT_DIR, f) for f in EXPERIMENT_SOURCES]
    for s in sources:
        o = os.path.join(BUILD, os.path.basename(s).replace('.cu', tag + '.o'))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append((s, o))
    marker = os.path.join(BUILD, 'flavour.txt')
    flavour = 'experimental' if EXPERIMENTAL else 'product'
    relink = not os.path.exists(marker) or open(marker).read() != flavour

    def compile_one(job):
        s, o = job
        cmd = [nvcc] + flags + ['-c', s, '-o', o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return s, r

    logs = []
    with ThreadPoolExecutor(max_workers=min(6, max(1, len(jobs)))) as ex:
        for s, r in ex.map(compile_one, jobs):
            logs.append(f'==== {os.path.basename(s)} ====\n{r.stderr}\n{r.stdout}')
            if r.returncode != 0:
                sys.stderr.write(logs[-1])
                raise RuntimeError(f'nvcc failed on {s}')
            if verbose:
                print(logs[-1])
    if jobs:
        with open(os.path.join(BUILD, 'ptxas.log'), 'w') as f:
            f.write('\n'.join(logs))
    if force or jobs or relink or not os.path.exists(LIB):
        cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart']
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stderr)
            raise RuntimeError('link failed')
        with open(marker, 'w') as f:
            f.write(flavour)
    return LIB


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--force', action='store_true')
    ap.add_argument('--verbose', action='store_true')
    a = ap.parse_args()
    print(build(a.force, a.verbose))

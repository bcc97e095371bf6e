"""nvcc recipe for libuvip_orb.so (sm_100a only, -lineinfo so ncu's source page maps to our code).
The library is built IN-TREE (u-vip-slam_b200/libuvip_orb.so) so it travels to the GPU box with the snapshot."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SO = os.path.join(HERE, 'libuvip_orb.so')
SOURCES = ['matcher.cu', 'extractor.cu', 'klt.cu', 'synth.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-Wall', '--shared', '-cudart', 'static',
              '-fmad=false',            # float parity: the reference build has no FMA contraction
              ]


def nvcc_path():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', shutil.which('nvcc')):
        if c and os.path.exists(c):
            return c
    raise RuntimeError('nvcc not found: libuvip_orb.so cannot be built and there is no CPU fallback')


def _deps():
    out = [os.path.join(HERE, '..', 'include', 'uvip_orb.h')]
    for f in os.listdir(CSRC):
        out.append(os.path.join(CSRC, f))
    return out


def build_cuda(force=False, verbose=False, extra=(), out=None):
    """out: alternative output path (A/B variants built with extra -D flags; selected at run time with UVIP_LIB)"""
    srcs = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if out is None and not force and os.path.exists(SO) and all(os.path.getmtime(d) <= os.path.getmtime(SO) for d in _deps()):
        return SO
    env = dict(os.environ)
    env.pop('CC', None); env.pop('CXX', None)
    cmd = [nvcc_path()] + NVCC_FLAGS + ['-ccbin', '/usr/bin/g++' if os.path.exists('/usr/bin/g++') else 'g++'] + \
        list(extra) + ['-o', out or SO] + srcs
    if verbose:
        print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd, env=env)
    return out or SO

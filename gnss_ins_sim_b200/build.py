"""Build csrc/libb2ins.so in-tree with nvcc for sm_100a.

    python -m gnss_ins_sim_b200.build [--force]

The shared library is a plain C-ABI library (include/b2ins.h); it links the static CUDA
runtime only, so it can be loaded with ctypes next to PyTorch (which supplies device
memory and streams) or from any other host language.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
# B2INS_LIB lets tools load an experimental build (tools/variants.sh); the product uses the default
LIB = os.environ.get('B2INS_LIB') or os.path.join(CSRC, 'libb2ins.so')
SOURCES = ['b2ins_api.cu']
DEPS = ['b2ins_api.cu', 'common.cuh', 'fastmath64.cuh', 'mech.cuh', 'mc_kernel.cuh', 'noise_kernel.cuh',
        'stats_kernel.cuh', 'allan_kernel.cuh', 'psd_kernel.cuh', 'gps_kernel.cuh', 'pathgen_host.h', os.path.join('..', '..', 'include', 'b2ins.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '-shared', '-Xcompiler', '-fPIC']


def find_nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found (set NVCC=/path/to/nvcc)')


def stale():
    if os.environ.get('B2INS_LIB'):
        return False
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, d)) > t for d in DEPS)


def build(force=False, verbose=False):
    """Compile the CUDA library if it is missing or older than its sources."""
    if not force and not stale():
        return LIB
    cmd = [find_nvcc()] + NVCC_FLAGS + ['-o', LIB] + SOURCES
    if verbose:
        cmd.insert(1, '-Xptxas')
        cmd.insert(2, '-v')
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed:\n' + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

"""Build the sm_100a CUDA library in-tree (arah_release_b200/libarah_b200.so) with nvcc.

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
SO = os.path.join(HERE, 'libarah_b200.so')
SOURCES = ['arah_api.cu', 'arah_mesh.cu', 'arah_hyper.cu', 'arah_rays.cu', 'arah_image.cu', 'arah_loss.cu']
HEADERS = ['arah_math.cuh', 'arah_tile.cuh', 'arah_kernels.cuh', 'arah_umma.cuh', 'arah_shade_tc.cuh', 'arah_corr_tc.cuh', 'arah_tc2.cuh', 'arah_shade_tc2.cuh', 'arah_corr_tc2.cuh', 'arah_shade_tc3.cuh', 'arah_corr_tc3.cuh', 'arah_sdf3x.cuh', 'arah_shade_tc4.cuh', 'arah_corr_tc4.cuh', 'arah_corr_tc5.cuh', 'arah_iso_init_tc.cuh', 'arah_train.h', 'arah_train_cuda.cuh', 'arah_train_tc.cuh', 'arah_image_core.h', 'arah_loss_core.h', os.path.join('..', '..', 'include', 'arah_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '--shared',
              '-Xcompiler', '-fPIC', '-Xptxas', '-v', '--threads', '4']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def needs_build():
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    cmd = [_nvcc()] + NVCC_FLAGS + ['-o', SO] + [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed building libarah_b200.so')
    with open(os.path.join(HERE, 'ptxas_info.txt'), 'w') as f:
        f.write(res.stderr)
    return SO


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))

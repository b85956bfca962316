"""Builds libb200ir.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

Usage: python -m image_restoration_b200.build [--force]
The .so lands next to this file so it travels with the repo snapshot to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libb200ir.so')
SOURCES = ['api.cu', 'conv_igemm.cu', 'conv_epi_generic.cu', 'conv_epi0.cu', 'conv_epi1.cu', 'conv_epi2.cu', 'conv_epi3.cu',
           'conv_epi4.cu', 'conv_epi5.cu', 'conv_epi6.cu', 'pointwise.cu', 'fir_tma.cu', 'resample_tma.cu', 'sr_ops.cu', 'degrade.cu', 'degrade_full.cu', 'wgrad.cu',
           'backward.cu', 'train_ops.cu']
HEADERS = ['ptx.cuh', 'host_common.h', 'conv_common.cuh', 'blur_taps.cuh', os.path.join('..', '..', 'include', 'b200ir.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--use_fast_math', '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    deps += [os.path.join(CSRC, f) for f in os.listdir(CSRC)]          # any new header counts too
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)

    def compile_one(src):
        obj = os.path.join(HERE, 'build', src.replace('.cu', '.o'))
        cmd = [nvcc, *NVCC_FLAGS, '-c', os.path.join(CSRC, src), '-o', obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, cmd, r

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:   # one nvcc process per translation unit
        results = list(ex.map(compile_one, SOURCES))
    objs = []
    for src, obj, cmd, r in results:
        if verbose or r.returncode != 0:
            sys.stderr.write(' '.join(cmd) + '\n' + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}')
        objs.append(obj)
    # shared CUDA runtime: the process already holds torch's libcudart.so.12 (one runtime instance instead of a second, static
    # one inside this library); the rpath covers a host program that loads the library without torch
    cuda_lib = os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(nvcc))), 'lib64')
    cmd = [nvcc, '-shared', '--cudart', 'shared', '-o', LIB, *objs, '-Xlinker', f'-rpath={cuda_lib}', '-ldl', '-lrt', '-lpthread']
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError('link failed')
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose=True))

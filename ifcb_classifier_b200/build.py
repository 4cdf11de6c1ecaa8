"""Builds the CUDA library in-tree: ifcb_classifier_b200/libifcb_b200.so (sm_100a only).

The shared object is git-ignored but travels to the GPU box with the repo
snapshot.  nvcc cross-compiles without a GPU, so this also runs on the CPU-only
build container (``__graft_entry__.build()``).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libifcb_b200.so')
SOURCES = ['runtime.cu', 'preprocess.cu', 'conv_umma.cu', 'conv_wgrad.cu', 'train_ops.cu', 'stem_pool_head.cu', 'plan.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC']


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + \
           [os.path.join(HERE, '..', 'include', 'ifcb_b200.h'), os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, 'build'), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, 'build', src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get('IFCB_NVCC_EXTRA', '').split() + (['-Xptxas', '-v'] if verbose else []) + \
              ['-c', os.path.join(CSRC, src), '-o', obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    tmp = LIB + '.tmp.%d' % os.getpid()              # link aside, then rename: a reader never sees a half-written library
    cmd = [nvcc, '-shared', '-o', tmp] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a']
    subprocess.check_call(cmd)
    os.replace(tmp, LIB)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

"""Build libspi_b200.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C ABI)."""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libspi_b200.so')
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '--use_fast_math=false',
         '-Xcompiler', '-fPIC', '-Xcompiler', '-O2']
FLAGS = [f for f in FLAGS if f != '--use_fast_math=false']


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def _stamp():
    h = hashlib.sha1()
    for f in _sources() + sorted(glob.glob(os.path.join(CSRC, '*.cuh'))):
        h.update(open(f, 'rb').read())
    h.update(' '.join(FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=True):
    stamp_file = LIB + '.stamp'
    stamp = _stamp()
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in _sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        cmd = [NVCC] + FLAGS + ['-I', os.path.join(os.path.dirname(HERE), 'include'), '-c', src, '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError(f'nvcc failed on {src}:\n{out}')
        if verbose and out.strip():
            print(out)
    subprocess.check_call([NVCC, '-shared', '-o', LIB] + objs + ['-lcudart'])
    open(stamp_file, 'w').write(stamp)
    if verbose:
        print('built', LIB)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv)

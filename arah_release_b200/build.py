"""Build the sm_100a CUDA library in-tree (arah_release_b200/libarah_b200.so) with nvcc.

nvcc cross-compiles without a GPU; the .so travels to the GPU box with the repo snapshot.
Every translation unit is compiled to its own object under build/ (git-ignored) and re-compiled only when the CONTENT of the
source, of any header under csrc/ or include/, or the flag set changes (a hash, not mtimes: a stale shipped .so is never
silently reused).  The ptxas resource log of each unit lands next to its object (build/*.ptxas.txt).
"""
import hashlib
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, 'csrc')
INC = os.path.join(ROOT, 'include')
OBJ = os.path.join(ROOT, 'build')
SO = os.path.join(HERE, 'libarah_b200.so')
SOURCES = ['arah_api.cu', 'arah_root.cu', 'arah_shade.cu', 'arah_mesh.cu', 'arah_hyper.cu', 'arah_rays.cu', 'arah_image.cu', 'arah_loss.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC', '-Xptxas', '-v']


def _nvcc():
    for c in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return 'nvcc'


def _headers():
    hs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(('.cuh', '.h'))]
    hs += [os.path.join(INC, f) for f in sorted(os.listdir(INC)) if f.endswith('.h')]
    return hs


_INC_RE = re.compile(r'^\s*#\s*include\s+"([^"]+)"', re.M)


def _deps(path, seen=None):
    """Quoted includes of `path`, transitively (paths that exist in the tree only)."""
    seen = set() if seen is None else seen
    with open(path) as f:
        text = f.read()
    for inc in _INC_RE.findall(text):
        q = os.path.normpath(os.path.join(os.path.dirname(path), inc))
        if os.path.exists(q) and q not in seen:
            seen.add(q)
            _deps(q, seen)
    return sorted(seen)


def _digest(paths, extra=''):
    h = hashlib.sha256(extra.encode())
    for p in paths:
        h.update(os.path.basename(p).encode())
        with open(p, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()


def source_hash():
    """Content hash of everything the library is built from (sources, headers, flags)."""
    return _digest([os.path.join(CSRC, s) for s in SOURCES] + _headers(), ' '.join(NVCC_FLAGS))


def _stamp_path():
    return os.path.join(OBJ, 'libarah_b200.hash')


def needs_build():
    if not os.path.exists(SO) or not os.path.exists(_stamp_path()):
        return True
    with open(_stamp_path()) as f:
        return f.read().strip() != source_hash()


def _compile_one(src, verbose):
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ, src + '.o')
    stamp = obj + '.hash'
    want = _digest([path] + _deps(path), ' '.join(NVCC_FLAGS))
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == want:
        return obj, ''
    cmd = [_nvcc()] + NVCC_FLAGS + ['-c', '-o', obj, path]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError('nvcc failed compiling ' + src)
    with open(obj + '.ptxas.txt', 'w') as f:
        f.write(res.stderr)
    with open(stamp, 'w') as f:
        f.write(want)
    return obj, res.stderr


def build(force=False, verbose=False):
    if not force and not needs_build():
        return SO
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            if f.endswith('.o.hash'):
                os.remove(os.path.join(OBJ, f))
    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = [o for o, _ in ex.map(lambda s: _compile_one(s, verbose), SOURCES)]
    cmd = [_nvcc(), '-gencode', 'arch=compute_100a,code=sm_100a', '--shared', '-Xcompiler', '-fPIC', '-o', SO] + objs
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError('nvcc failed linking libarah_b200.so')
    with open(_stamp_path(), 'w') as f:
        f.write(source_hash())
    return SO


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))

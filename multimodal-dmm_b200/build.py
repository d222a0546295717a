"""In-tree nvcc build of libbfvi_b200.so for sm_100a (B200)."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libbfvi_b200.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
              '--shared', '-Xcompiler', '-fPIC', '-Xcompiler', '-O3', '--threads', '2']


# translation units of the one library (everything else under csrc/ is a header they include)
UNITS = [os.path.join(CSRC, 'bfvi_api.cu'), os.path.join(CSRC, 'bfvi_conv.cu')]


def sources():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [
        os.path.join(os.path.dirname(HERE), 'include', 'bfvi.h')]


def source_id():
    """16 hex digits of the SHA-256 over csrc/ (name order) and include/bfvi.h: stamped into the library as
    bfvi_build_id() so that a binary can be matched to the sources it was built from."""
    import hashlib
    h = hashlib.sha256()
    for path in sources():
        h.update(os.path.basename(path).encode() + b'\0')
        with open(path, 'rb') as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def up_to_date():
    return os.path.exists(OUT) and all(
        os.path.getmtime(OUT) >= os.path.getmtime(s) for s in sources())


def build(force=False, verbose=False, defines=(), out=None):
    """Compiles every CUDA source into one shared library next to this file.
    `defines` / `out` build a tuning variant (tools/variants.py) beside the product."""
    if out is not None:
        nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
        subprocess.run([nvcc] + NVCC_FLAGS + ['-D%s' % d for d in defines] +
                       UNITS + ['-o', out], check=True)
        return out
    if not force and up_to_date():
        return OUT
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: cannot build libbfvi_b200.so')
    cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + [
        '-DBFVI_SOURCE_ID="%s"' % source_id()] + UNITS + ['-o', OUT]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == '__main__':
    print(build(force=True, verbose=True))

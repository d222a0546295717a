"""Builds tests/emu/_build/libbfvi_emu.so: the CUDA kernel sources compiled as host
C++ against cuda_emu.h (development / test tool, see that header)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'multimodal-dmm_b200', 'csrc')
OUT = os.path.join(HERE, '_build', 'libbfvi_emu.so')


def build(force=False):
    srcs = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [
        os.path.join(HERE, 'cuda_emu.h'), os.path.join(ROOT, 'include', 'bfvi.h')]
    if not force and os.path.exists(OUT) and all(
            os.path.getmtime(OUT) >= os.path.getmtime(s) for s in srcs):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    cmd = ['g++', '-std=c++17', '-O1', '-g', '-fPIC', '-shared', '-DBFVI_EMU', '-x', 'c++',
           '-I', HERE, '-I', CSRC, os.path.join(CSRC, 'bfvi_api.cu'), os.path.join(CSRC, 'bfvi_conv.cu'), '-o', OUT]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == '__main__':
    print(build(force=True))

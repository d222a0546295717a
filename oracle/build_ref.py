"""Byte-compile the UNMODIFIED reference (ztangent/multimodal-dmm) from where it lies into oracle/_ref/.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/bfvi_oracle.py).  The reference is a flat Python script repo
(no setup.py, nothing to `pip install`), and `/root/reference` does not exist on the GPU box.  This recipe compiles
the path's own modules — models/{__init__,common,dgts,dmm,dks,vrnn,losses}.py and, for the trainer-level drop-in
test, trainer.py, spirals.py, utils.py, datasets/{__init__,multiseq,spirals}.py — with `py_compile` into SOURCELESS
bytecode files (`oracle/_ref/models/dmm.pybc` …: the .pyc format under an extension of its own, because the GPU
runner's snapshot drops `*.pyc`; oracle/ref_shim.py installs the importer that reads them), the Python analogue of compiling a C reference into
`oracle/_ref/*.so`: outputs only, no reference source is copied, `oracle/_ref/` is git-ignored (and not
gpurun-ignored, so it travels to the GPU box like the built .so files).

Users (never the product path):
  * `bench.py --impl reference` and the `cpu_baseline` leg: the reference's own `MultiDMM.step` + backward on the host
    cores (`cpu_baseline.kind = "reference"`);
  * tests/test_reference_dropin.py: the reference's `SpiralsTrainer` with `models` swapped for this package.

    python oracle/build_ref.py        (no-op when /root/reference is absent: prebuilt files are used as they are)
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, '_ref')
REFERENCE_ROOT = '/root/reference'
EXT = '.pybc'
FILES = ['models/__init__.py', 'models/common.py', 'models/dgts.py', 'models/dmm.py', 'models/dks.py',
         'models/vrnn.py', 'models/losses.py', 'trainer.py', 'spirals.py', 'utils.py',
         'datasets/__init__.py', 'datasets/multiseq.py', 'datasets/spirals.py']


def available():
    return os.path.exists(os.path.join(OUT, 'models', 'dmm' + EXT))


def build(force=False):
    """Returns the list of bytecode files (empty when neither the reference nor a prebuilt copy exists)."""
    outs = [os.path.join(OUT, f[:-3] + EXT) for f in FILES]
    if not os.path.isdir(REFERENCE_ROOT):
        return [o for o in outs if os.path.exists(o)]
    tag = os.path.join(OUT, 'PYTHON_VERSION')
    same_python = os.path.exists(tag) and open(tag).read().strip() == sys.version.split()[0]
    for f, o in zip(FILES, outs):
        src = os.path.join(REFERENCE_ROOT, f)
        if not force and same_python and os.path.exists(o) and os.path.getmtime(o) >= os.path.getmtime(src):
            continue
        os.makedirs(os.path.dirname(o), exist_ok=True)
        # unchecked-hash pyc: valid for a sourceless import whatever the mtime of the (absent) source
        py_compile.compile(src, cfile=o, dfile=f, doraise=True,
                           invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    with open(tag, 'w') as fh:
        fh.write(sys.version.split()[0] + '\n')
    return outs


if __name__ == '__main__':
    for p in build(force=True):
        print(p)

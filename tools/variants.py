"""Build tuning variants of libbfvi_b200.so (different launch geometry macros) into
tools/_variants/ so ONE gpurun call can time them all (tools/quick_time.py with
BFVI_LIB_PATH).  usage: python tools/variants.py name:DEF=V,DEF=V ..."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location('_b', os.path.join(ROOT, 'multimodal-dmm_b200', 'build.py'))
b = importlib.util.module_from_spec(spec)
spec.loader.exec_module(b)
out_dir = os.path.join(ROOT, 'tools', '_variants')
os.makedirs(out_dir, exist_ok=True)
for arg in sys.argv[1:]:
    name, _, defs = arg.partition(':')
    defines = [d for d in defs.split(',') if d]
    print(b.build(defines=defines, out=os.path.join(out_dir, 'libbfvi_%s.so' % name)))

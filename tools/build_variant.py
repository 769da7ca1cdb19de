"""Build a variant of libbkf.so with extra -D flags on selected translation units (debug / profiling builds).

    python tools/build_variant.py NAME "FLAGS" file1.cu [file2.cu ...]   ->  pymc_experimental_b200/libbkf_NAME.so
The other objects come from the product build (pymc_experimental_b200/build).  Use with DBG_LIB=libbkf_NAME.so.
"""
import os, subprocess, sys
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pymc_experimental_b200 import build as B
name, flags, files = sys.argv[1], sys.argv[2].split(), sys.argv[3:]
B.build_library()
vdir = os.path.join(B.HERE, "build", "var_" + name)
os.makedirs(vdir, exist_ok=True)
def comp(f):
    obj = os.path.join(vdir, f[:-3] + ".o")
    r = subprocess.run(["nvcc", *B.NVCC_FLAGS, *flags, "-c", os.path.join(B.CSRC, f), "-o", obj], capture_output=True, text=True)
    if r.returncode: raise RuntimeError(r.stderr)
    return obj
with ThreadPoolExecutor(8) as ex: vobjs = list(ex.map(comp, files))
objs = [os.path.join(B.OBJDIR, os.path.basename(s)[:-3] + ".o") for s in B.sources() if os.path.basename(s) not in files] + vobjs
lib = os.path.join(B.HERE, f"libbkf_{name}.so")
r = subprocess.run(["nvcc", "-shared", "-o", lib, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"], capture_output=True, text=True)
if r.returncode: raise RuntimeError(r.stderr)
print(lib)

#!/usr/bin/env python
"""Build a variant of the product library with extra nvcc flags (kernel experiments).

    python tools/exp/variant.py g4 -DWQ_GROUPS_N=4          -> tools/exp/variants/lib_g4.so
    gpurun -- 'LIEVAE_LIB=tools/exp/variants/lib_g4.so python tools/time_wigner.py'
Only wigner.cu is rebuilt per variant unless --all is given; the other objects come from lie_vae_b200/build/.
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lie_vae_b200 import _build  # noqa: E402

tag, flags = sys.argv[1], sys.argv[2:]
srcs = ["wigner.cu"] if "--gemm" not in sys.argv else ["gemm_tf32.cu"]
if "--gemm" in flags:
    flags.remove("--gemm")
if "--all" in flags:
    flags.remove("--all")
    srcs = list(_build.SOURCES)
_build.build()
out = os.path.join(ROOT, "tools", "exp", "variants")
os.makedirs(out, exist_ok=True)
objs = []
for s in _build.SOURCES:
    if s in srcs:
        obj = os.path.join(out, "%s_%s.o" % (s[:-3], tag))
        cmd = [_build._nvcc()] + [f for f in _build.NVCC_FLAGS if f != "-shared"] + flags + ["-Xptxas", "-v", "-c", "-o", obj, s]
        pr = subprocess.run(cmd, cwd=_build.CSRC, capture_output=True, text=True)
        if pr.returncode:
            sys.exit(pr.stderr[-6000:])
        log = pr.stderr.splitlines()
        for i, l in enumerate(log):
            if "bwd_tma_kernelILi10ELi8" in l and "Function properties" in l:
                print("\n".join(x.strip() for x in log[i + 1:i + 3]))
    else:
        obj = os.path.join(_build.HERE, "build", s.replace(".cu", ".o"))
    objs.append(obj)
lib = os.path.join(out, "lib_%s.so" % tag)
subprocess.check_call([_build._nvcc(), "-shared", "-o", lib] + objs)
print(lib)

"""Build liblievae_sm100a.so in-tree with nvcc (sm_100a only; cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "liblievae_sm100a.so")
SOURCES = ["cabi.cu", "elementwise.cu", "reparam.cu", "wigner.cu", "wigner_generic.cu", "head_reparam.cu", "gemm_tf32.cu"]
HEADERS = ["common.cuh", "wigner_gen.cuh", "wigner_bwd_dg.cuh", "reparam_core.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def is_stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def _compile_one(args):
    src, obj, verbose = args
    extra = os.environ.get("LV_NVCC_EXTRA", "").split()      # build-time experiments only
    cmd = [_nvcc()] + [f for f in NVCC_FLAGS if f != "-shared"] + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
    proc = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    return src, proc.returncode, proc.stdout + proc.stderr


def build(force=False, verbose=False):
    """Compile every CUDA source (in parallel, one object per file) and link lie_vae_b200/liblievae_sm100a.so."""
    if not force and not is_stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    jobs = [(s, os.path.join(objdir, s.replace(".cu", ".o")), verbose) for s in SOURCES]
    with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
        results = list(ex.map(_compile_one, jobs))
    for src, rc, log in results:
        if rc != 0:
            sys.stderr.write(log[-12000:])
            raise RuntimeError("nvcc failed compiling %s" % src)
        if verbose:
            sys.stderr.write("==== %s\n%s" % (src, log))
    link = [_nvcc(), "-shared", "-Xlinker", "-soname=" + os.path.basename(LIB), "-o", LIB] + [j[1] for j in jobs]
    proc = subprocess.run(link, cwd=CSRC, capture_output=True, text=True)
    if proc.returncode != 0:
        sys.stderr.write(proc.stdout[-4000:] + proc.stderr[-8000:])
        raise RuntimeError("nvcc failed linking %s" % LIB)
    return LIB


TORCH_EXT_NAME = "lievae_torch"
TORCH_EXT_DIR = os.path.join(HERE, "build_torch")
TORCH_EXT_LIB = os.path.join(TORCH_EXT_DIR, TORCH_EXT_NAME + ".so")
TORCH_EXT_SRC = os.path.join(CSRC, "torch_binding.cpp")


def _torch_ext_src_hash():
    import hashlib
    h = hashlib.sha256()
    for path in (TORCH_EXT_SRC, os.path.join(os.path.dirname(HERE), "include", "lievae.h")):
        with open(path, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def torch_ext_is_stale():
    """True unless the built extension carries the hash of the sources it was compiled from (content, not mtimes: a snapshot
    or checkout of the tree may reset those)."""
    try:
        with open(TORCH_EXT_LIB + ".srchash") as f:
            return not os.path.exists(TORCH_EXT_LIB) or f.read().strip() != _torch_ext_src_hash()
    except OSError:
        return True


def build_torch_ext(force=False, verbose=False):
    """Compile csrc/torch_binding.cpp (C++ autograd Functions over the C ABI; host code only, no device code) in-tree with
    torch.utils.cpp_extension; linked against liblievae_sm100a.so next to it (rpath $ORIGIN/..)."""
    if not force and not torch_ext_is_stale():
        return TORCH_EXT_LIB
    build()
    from torch.utils import cpp_extension
    os.makedirs(TORCH_EXT_DIR, exist_ok=True)
    cpp_extension.load(name=TORCH_EXT_NAME, sources=[TORCH_EXT_SRC], extra_cflags=["-O2", "-std=c++17"],
                       extra_ldflags=["-L" + HERE, "-l:" + os.path.basename(LIB), "-Wl,-rpath,'$$ORIGIN/..'"],    # $$ for ninja, quotes for sh
                       build_directory=TORCH_EXT_DIR, with_cuda=True, verbose=verbose)
    with open(TORCH_EXT_LIB + ".srchash", "w") as f:
        f.write(_torch_ext_src_hash())
    return TORCH_EXT_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    if "--no-torch-ext" not in sys.argv:
        print(build_torch_ext(force="--force" in sys.argv, verbose="-v" in sys.argv))

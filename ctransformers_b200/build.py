"""Build recipe for the native library (explicit nvcc, in-tree output).

    python -m ctransformers_b200.build            # builds ctransformers_b200/lib/libctransformers.so

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  sm_100a only.
"""
import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OUT_DIR = PKG / "lib"
LIB = OUT_DIR / "libctransformers.so"
OBJ_DIR = PKG.parent / "build" / "obj"
SOURCES = ["engine.cu", "llm_abi.cu", "ops_abi.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function,-ffp-contract=off", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(exe).exists():
        raise RuntimeError("nvcc not found: the native library cannot be built")
    return exe


def _digest():
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*")) + [PKG.parent / "include" / "ctransformers_b200.h", Path(__file__)]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    return h.hexdigest()


def build(force=False, verbose=False, variant=None, defines=()):
    """Compile every CUDA source for sm_100a and link libctransformers.so.  Returns the library path.
    variant + defines: an A/B build (tools/ab.py) with extra -D flags, written to lib/libctransformers_<variant>.so."""
    OUT_DIR.mkdir(exist_ok=True)
    obj_dir = OBJ_DIR if not variant else OBJ_DIR.parent / f"obj_{variant}"
    lib = LIB if not variant else OUT_DIR / f"libctransformers_{variant}.so"
    obj_dir.mkdir(parents=True, exist_ok=True)
    stamp = obj_dir / "digest"
    dig = _digest() + "|" + " ".join(defines)
    if not force and lib.exists() and stamp.exists() and stamp.read_text() == dig:
        return lib
    exe = nvcc()

    def compile_one(src):
        obj = obj_dir / (src + ".o")
        cmd = [exe, *NVCC_FLAGS, *defines, "-c", str(CSRC / src), "-o", str(obj)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        (obj_dir / (src + ".ptxas.log")).write_text(r.stderr)
        if verbose:
            print(r.stderr, file=sys.stderr)
        return obj

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [exe, "-shared", "-o", str(lib), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    stamp.write_text(dig)
    return lib


if __name__ == "__main__":
    var = next((a.split("=", 1)[1] for a in sys.argv if a.startswith("--variant=")), None)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, variant=var, defines=[a for a in sys.argv[1:] if a.startswith("-D")]))

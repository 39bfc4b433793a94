"""Build recipe for libcsd_b200.so (in-tree, sm_100a only).

`python -m conditional_score_diffusion_b200.build` compiles every csrc/*.cu with
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo and links one shared library next to this
file. nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the tree.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "_build")
LIB_PATH = os.path.join(HERE, "libcsd_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libcsd_b200 cannot be built")
    return nvcc


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS + _extra_flags()).encode())
    return h.hexdigest()


def _all_inputs():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    files.append(os.path.join(HERE, "..", "include", "csd_b200.h"))
    return [os.path.abspath(f) for f in files]


def _extra_flags():
    """Developer-only extra nvcc flags (e.g. CSD_NVCC_EXTRA=-DCSD_ENABLE_PHASE_TIMESTAMPS)."""
    return os.environ.get("CSD_NVCC_EXTRA", "").split()


def _compile_one(src):
    obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
    cmd = [_nvcc()] + NVCC_FLAGS + _extra_flags() + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, obj, r.returncode, r.stdout + r.stderr


def build(force=False, verbose=False):
    """Compile and link libcsd_b200.so if sources changed. Returns the library path."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    stamp = os.path.join(OBJ_DIR, "stamp.txt")
    digest = _digest(_all_inputs())
    if not force and os.path.exists(LIB_PATH) and os.path.exists(stamp):
        if open(stamp).read().strip() == digest:
            return LIB_PATH
    srcs = _sources()
    objs, logs = [], []
    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        for src, obj, rc, log in ex.map(_compile_one, srcs):
            logs.append((src, log))
            if rc != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{log}")
            objs.append(obj)
    with open(os.path.join(OBJ_DIR, "ptxas.log"), "w") as f:
        for src, log in logs:
            f.write(f"==== {src}\n{log}\n")
    if verbose:
        for src, log in logs:
            print(f"==== {src}\n{log}")
    cmd = [_nvcc(), "-shared", "-o", LIB_PATH] + objs + ["-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return LIB_PATH


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)

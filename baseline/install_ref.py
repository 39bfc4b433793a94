"""Install the UNMODIFIED reference hot path under baseline/_ref/ so it can be timed on the GPU box.

Run in the build container (where /root/reference is mounted read-only):

    python baseline/install_ref.py

The reference (GBATZOLIS/conditional_score_diffusion) is a flat script tree without setup.py / pyproject.toml, so the
base contract's `pip install --target baseline/_ref /root/reference` has nothing to install; this script does what
that command would have done for the files the benchmarked path imports: it copies them byte for byte (sha256
recorded in baseline/_ref/MANIFEST.json) and pre-builds the reference's two JIT CUDA extensions (op/upfirdn2d*,
op/fused_bias_act*; `torch.utils.cpp_extension.load` at `import op`, op/upfirdn2d.py:10-16, op/fused_act.py:11-17)
for compute_100 so the GPU box does not spend ~2 minutes compiling them. baseline/_ref/ is git-ignored (the
reference's sources never enter this repository's history) but NOT gpurun-ignored, so it travels with the snapshot.

Nothing under baseline/ is imported by the product package; only bench.py's reference arm / stock-GPU leg use it.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
REF = os.environ.get("CSD_REFERENCE", "/root/reference")

FILES = [
    "sde_lib.py", "losses.py", "likelihood.py",
    "models/__init__.py", "models/utils.py", "models/ncsnpp.py", "models/ddpm.py", "models/layers.py",
    "models/layerspp.py", "models/up_or_down_sampling.py", "models/normalization.py", "models/ema.py",
    "op/__init__.py", "op/upfirdn2d.py", "op/upfirdn2d.cpp", "op/upfirdn2d_kernel.cu",
    "op/fused_act.py", "op/fused_bias_act.cpp", "op/fused_bias_act_kernel.cu",
    "sampling/__init__.py", "sampling/conditional.py", "sampling/unconditional.py", "sampling/predictors.py",
    "sampling/correctors.py",
]


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def install(prebuild=True, verbose=True):
    if not os.path.isdir(REF):
        raise RuntimeError(f"reference tree {REF} not present (the GPU box uses the prebuilt baseline/_ref)")
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or _sha(src) != _sha(dst):
            shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DST, "MANIFEST.json"), "w") as f:
        json.dump({"reference": REF, "sha256": manifest}, f, indent=1)
    if prebuild:
        prebuild_extensions(verbose)
    return DST


def prebuild_extensions(verbose=True):
    """Build the reference's own extensions exactly as its `load(...)` calls would (same sources, default flags),
    with the arch list the B200 needs, into baseline/_ref/_ext/<name>/<name>.so."""
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    from torch.utils import cpp_extension as ext
    for name, srcs in (("upfirdn2d", ["op/upfirdn2d.cpp", "op/upfirdn2d_kernel.cu"]),
                       ("fused", ["op/fused_bias_act.cpp", "op/fused_bias_act_kernel.cu"])):
        bdir = os.path.join(DST, "_ext", name)
        so = os.path.join(bdir, name + ".so")
        srcs_abs = [os.path.join(DST, s) for s in srcs]
        stamp = os.path.join(bdir, "stamp.txt")
        digest = "".join(_sha(s) for s in srcs_abs)
        if os.path.exists(so) and os.path.exists(stamp) and open(stamp).read() == digest:
            continue
        os.makedirs(bdir, exist_ok=True)
        if verbose:
            print(f"[baseline] building the reference's '{name}' extension (about a minute)...", file=sys.stderr)
        ext.load(name, sources=srcs_abs, build_directory=bdir, verbose=False)
        with open(stamp, "w") as f:
            f.write(digest)


if __name__ == "__main__":
    print(install())

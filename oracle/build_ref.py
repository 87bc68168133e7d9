"""oracle/_ref recipe: the UNMODIFIED reference, byte-compiled where it lies.  TEST INFRASTRUCTURE.

    python oracle/build_ref.py [--reference /root/reference]

The reference is Python: "building" it means `py_compile` of the handful of files on the hot path,
straight from the sources under /root/reference into oracle/_ref/*.pyc (git-ignored, NOT gpurun-ignored:
the bytecode travels to the GPU box like our own .so, the sources never enter the repo). The GPU box
runs the same image (same CPython magic number), so the .pyc files import there unchanged -- that is
how `-m gpu` tests, smoke() and `bench.py --impl reference` execute the real reference on the box
(`cpu_baseline.kind = "reference"`) although /root/reference does not exist there.

Only tests/, __graft_entry__ and bench.py's reference/cpu_baseline legs load these (oracle/ref_loader.py);
the product package never does.
"""
import argparse
import json
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")

# logical name -> path under the reference tree
FILES = {
    # the solver itself (every row cites this file)
    "dpm_solver_pytorch": "dpm_solver_pytorch.py",
    # Stable-Diffusion adapter + its vendored (older) solver copy (SURVEY 8b "who calls it")
    "sd_sampler": "examples/stable-diffusion/ldm/models/diffusion/dpm_solver/sampler.py",
    "sd_dpm_solver": "examples/stable-diffusion/ldm/models/diffusion/dpm_solver/dpm_solver.py",
    # score_sde glue (get_dpm_solver_sampler)
    "score_sde_sampling": "examples/score_sde_pytorch/sampling.py",
    # guided-diffusion runner (Diffusion.sample_image: classifier guidance + dynamic thresholding) and
    # the solver copy it imports
    "guided_runner": "examples/ddpm_and_guided-diffusion/runners/diffusion.py",
    "guided_sampler": "examples/ddpm_and_guided-diffusion/dpm_solver/sampler.py",
}


def build_ref(reference="/root/reference", quiet=False):
    """Byte-compile FILES into oracle/_ref/. Returns the manifest, or None when the reference tree is
    absent (the GPU box: the prebuilt files are used as they are)."""
    if not os.path.isfile(os.path.join(reference, FILES["dpm_solver_pytorch"])):
        return None
    os.makedirs(OUT, exist_ok=True)
    manifest = {"python": sys.version.split()[0], "files": {}}
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", SyntaxWarning)      # the reference's docstrings contain "\h" etc.
        for name, rel in FILES.items():
            src = os.path.join(reference, rel)
            dst = os.path.join(OUT, name + ".pyc")
            py_compile.compile(src, cfile=dst, dfile=rel, doraise=True)
            manifest["files"][name] = {"source": rel, "bytes": os.path.getsize(src)}
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    if not quiet:
        print("oracle/_ref: compiled", ", ".join(FILES))
    return manifest


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("DPM_REFERENCE", "/root/reference"))
    a = ap.parse_args()
    if build_ref(a.reference) is None:
        print("reference tree not found at", a.reference, "- keeping the prebuilt oracle/_ref", file=sys.stderr)

"""Recipe for oracle/_ref: the UNMODIFIED reference modules of the hot path, compiled to bytecode.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py): nothing under pytorch-video-caption-rationale_b200/ may
import this.  The reference is pure Python (no native code), so "compiling the reference from the
sources where they lie" means `py_compile`: every file of the path is compiled from /root/reference into a
bytecode file `<module>.bc` under oracle/_ref/ (git-ignored, not gpurun-ignored: it travels to the GPU box like our own
built `.so`; the extension is not `.pyc` because snapshot tools commonly drop `*.pyc`).  No reference source text is copied into the repository; the GPU box (same image, same CPython 3.12) imports
the bytecode.  Run by `__graft_entry__.build()` whenever /root/reference is present.

    python oracle/build_ref.py            # -> oracle/_ref/{utils,train_utils}.bc, oracle/_ref/model/*.bc
"""
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE = os.environ.get("PVCR_REFERENCE", "/root/reference")
# the files of the hot path and its callers' loss contract; model/__init__.py is empty
FILES = ["utils.py", "train_utils.py", "model/__init__.py", "model/S2VTModel.py", "model/S2VTAttModel.py",
         "model/RationaleNet.py", "model/SpatialNet.py"]


def build(force=False):
    """-> list of written bytecode paths ([] when the reference tree is absent, e.g. on the GPU box)."""
    if not os.path.isdir(REFERENCE):
        return []
    written = []
    for rel in FILES:
        src = os.path.join(REFERENCE, rel)
        dst = os.path.join(OUT, rel[:-3] + ".bc")
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            py_compile.compile(src, cfile=dst, dfile="reference/" + rel, doraise=True,
                               invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
        written.append(dst)
    with open(os.path.join(OUT, "PYTHON_TAG"), "w") as f:
        f.write(sys.implementation.cache_tag + "\n")
    return written


if __name__ == "__main__":
    for p in build(force="--force" in sys.argv):
        print(p)

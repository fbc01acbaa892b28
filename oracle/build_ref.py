"""TEST INFRASTRUCTURE ONLY.  Build oracle/_ref/aitref_C.so = the reference's own CPU
`model._C` (nms, roi_align_forward) from the sources under /root/reference.

Only runs where /root/reference exists (the build container).  The produced .so is
git-ignored but travels to the GPU box with the snapshot; there it is loaded prebuilt by
`oracle.ref_ops.load_ref_C()` (checker + `bench.py --impl reference` CPU arm only).
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("AIT_REFERENCE_ROOT", "/root/reference")
OUT_DIR = os.path.join(HERE, "_ref")
NAME = "aitref_C"


def so_path():
    return os.path.join(OUT_DIR, NAME + ".so")


def build(verbose=False):
    csrc = os.path.join(REF_ROOT, "lib", "model", "csrc")
    if not os.path.isdir(csrc):
        if os.path.exists(so_path()):
            return so_path()
        raise FileNotFoundError(
            "reference sources not found at %s and no prebuilt %s" % (csrc, so_path()))
    os.makedirs(OUT_DIR, exist_ok=True)
    src = os.path.join(HERE, "ref_wrap.cpp")
    if os.path.exists(so_path()) and os.path.getmtime(so_path()) >= os.path.getmtime(src):
        return so_path()
    from torch.utils.cpp_extension import load
    load(name=NAME, sources=[src], extra_include_paths=[csrc],
         extra_cflags=["-O2", "-w"], build_directory=OUT_DIR, verbose=verbose)
    return so_path()


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))

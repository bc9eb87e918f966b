"""Stage the reference's tf2_examples scripts for the GPU box.

The reference checkout (/root/reference) exists only in the build container, and its sources may not be committed to this
repo.  `tests/test_gpu_reference_examples.py` runs the UNMODIFIED scripts against liborx.so on the GPU, so they have to
travel with the gpurun snapshot: this copies them into tests/_ref_examples/, which is git-ignored (never enters history)
but not gpurun-ignored.  Called by __graft_entry__.build() whenever the reference is present.
    python tools/stage_reference_examples.py"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/tf2_examples"
DST = os.path.join(ROOT, "tests", "_ref_examples")
FILES = ("bpr_citeulike.py", "dlrm_criteo.py", "dataloader.py")


def stage():
    if not os.path.isdir(SRC):
        return None
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
    return DST


if __name__ == "__main__":
    d = stage()
    print(d if d else "reference checkout not present: nothing staged")
    sys.exit(0)

"""TEST / BASELINE INFRASTRUCTURE ONLY -- recipe that makes the *unmodified* reference travel to the GPU box.

The reference is pure Python (no build step): ``build_ref()`` copies the package ``/root/reference/aaerec`` file
by file into ``oracle/_ref/aaerec`` (git-ignored, so no reference source enters the history; NOT gpurun-ignored,
so it ships with the snapshot like the built ``.so``).  It runs in the authoring container, where
``/root/reference`` exists (``__graft_entry__.build()`` calls it); on the GPU box the prebuilt copy is used.

Users of ``oracle/_ref``: ``bench.py --impl reference`` (the reference's own ``AdversarialAutoEncoder`` timed on the
host cores with CUDA hidden, ``cpu_baseline.kind = "reference"``), ``bench.py``'s ``gpu_baseline`` leg (the same
unmodified code on the B200 through stock PyTorch) and the harness parity test (``tests/test_gpu_harness.py``).
The product package never imports it.
"""
import filecmp
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("AAE_REFERENCE_ROOT", "/root/reference")
REF_DST = os.path.join(HERE, "_ref")


def build_ref(verbose=True):
    src = os.path.join(REF_SRC, "aaerec")
    if not os.path.isdir(src):
        if verbose:
            print("oracle/build_ref: %s not present; keeping %s as it is" % (src, REF_DST))
        return os.path.isdir(os.path.join(REF_DST, "aaerec"))
    dst = os.path.join(REF_DST, "aaerec")
    os.makedirs(dst, exist_ok=True)
    n = 0
    for name in sorted(os.listdir(src)):
        if not name.endswith(".py"):
            continue
        a, b = os.path.join(src, name), os.path.join(dst, name)
        if not os.path.exists(b) or not filecmp.cmp(a, b, shallow=False):
            shutil.copyfile(a, b)
            n += 1
    with open(os.path.join(REF_DST, "README"), "w") as fh:
        fh.write("Unmodified copy of %s/aaerec made by oracle/build_ref.py (git-ignored; ships to the GPU box).\n" % REF_SRC)
    if verbose:
        print("oracle/build_ref: %d file(s) refreshed in %s" % (n, dst))
    return True


if __name__ == "__main__":
    build_ref()

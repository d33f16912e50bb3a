"""TEST / BENCH INFRASTRUCTURE -- recipe that vendors the Python reference for timing.

Copies the reference's own .py files, byte for byte, from /root/reference into
oracle/_ref/ (git-ignored, NOT gpurun-ignored: it travels to the GPU box like the
built .so files, where /root/reference does not exist).  Nothing is edited; the
copies are only ever imported by oracle/ref_baseline.py -- the `cpu_baseline` /
`--impl reference` arm of bench.py, which times the unmodified reference on the
box's host cores.  Never imported by thesis_b200/.

    python oracle/vendor_ref.py            # done by __graft_entry__.build() when the reference is present
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("THESIS_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def ref_files():
    """Every top-level module of the reference: main.py (whose body is under __main__, so that
    main.resample is importable) imports all of its loaders at the top."""
    return sorted(f for f in os.listdir(SRC) if f.endswith(".py"))


def vendor():
    if not os.path.isfile(os.path.join(SRC, "robot.py")):
        return None
    os.makedirs(DST, exist_ok=True)
    files = ref_files()
    for f in files:
        shutil.copyfile(os.path.join(SRC, f), os.path.join(DST, f))
        assert filecmp.cmp(os.path.join(SRC, f), os.path.join(DST, f), shallow=False)
    with open(os.path.join(DST, "VENDORED.txt"), "w") as fh:
        fh.write("verbatim copies of %s from %s (oracle/vendor_ref.py); do not edit, do not commit\n" % (", ".join(files), SRC))
    return DST


if __name__ == "__main__":
    d = vendor()
    print(d if d else "reference tree not present at %s" % SRC)
    sys.exit(0 if d else 1)

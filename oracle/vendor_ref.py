"""Recipe that makes the UNMODIFIED reference available on the GPU box: copies the Python sources of the path's
packages from /root/reference/codes into oracle/_ref/codes (git-ignored build output, like a compiled reference
binary would be; it travels with the gpurun snapshot, never into history).

TEST / BASELINE INFRASTRUCTURE.  Used by: ``bench.py --impl reference`` and its ``cpu_baseline`` leg (the reference's
own modules timed on the host cores, ``cpu_baseline.kind = "reference"``), and tests/test_compat_reference_model.py
(the reference's SRRaGANModel running on top of this package).  Nothing in the product imports it.
Run:  python -m oracle.vendor_ref        (``__graft_entry__.build()`` does when /root/reference exists)
"""
import os
import shutil
import sys

SRC = os.environ.get("ESR_REFERENCE_SRC", "/root/reference/codes")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "codes")
WANTED = ["CEM", "models", "utils", "options", "Z_optimization.py"]
KEEP_EXT = (".py", ".json")


def vendor(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "CEM")):
        if verbose:
            print("vendor_ref: %s not present, nothing to do" % SRC)
        return None
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for item in WANTED:
        s = os.path.join(SRC, item)
        if os.path.isfile(s):
            os.makedirs(DST, exist_ok=True)
            shutil.copy2(s, os.path.join(DST, item))
            n += 1
            continue
        for root, _, files in os.walk(s):
            for f in files:
                if f.endswith(KEEP_EXT):
                    rel = os.path.relpath(os.path.join(root, f), SRC)
                    os.makedirs(os.path.dirname(os.path.join(DST, rel)), exist_ok=True)
                    shutil.copy2(os.path.join(root, f), os.path.join(DST, rel))
                    n += 1
    if verbose:
        print("vendor_ref: %d files -> %s" % (n, DST))
    return DST


if __name__ == "__main__":
    sys.exit(0 if vendor() or True else 1)

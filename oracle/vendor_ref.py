"""TEST INFRASTRUCTURE ONLY -- makes the *reference's own Python modules* for the hot path available on
the GPU box, next to its CUDA extension (oracle/build_ref_ext.py), so that the -m gpu parity tests and
bench.py's reference legs can run the UNMODIFIED reference there (`/root/reference` does not exist on
the box; `oracle/_ref/` travels with the snapshot and is git-ignored, so no reference source enters the
history).

    python oracle/vendor_ref.py        (also run by __graft_entry__.build() when /root/reference exists)

What is copied: exactly the files Python loads when the reference's agents are imported through
oracle/ref_shim.py (the transitive closure of `networks.posenet_agent` + `networks.pts_encoder.pointnet2`
inside /root/reference, found from sys.modules -- about 35 files), byte for byte, keeping their relative
paths, into oracle/_ref/refpkg/.  Nothing is edited; the missing third-party imports (ipdb, tensorboardX,
cutoop, matplotlib) are stubbed at import time by ref_shim, as SURVEY.md section 8(c) describes.
The product (genpose2_b200/) never imports anything from here.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC_ROOT = "/root/reference"
OUT = os.path.join(HERE, "_ref", "refpkg")

# data files the modules read at import / construction time
EXTRA = ["configs/xyzibd_trans_mean.npy", "configs/xyzibd_trans_std.npy"]


def vendor(verbose=False):
    if not os.path.isdir(os.path.join(SRC_ROOT, "networks")):
        return None
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    os.environ["GENPOSE2_REFERENCE_ROOT"] = SRC_ROOT
    from oracle import ref_shim
    import importlib

    ref_shim.load()
    importlib.import_module("networks.pts_encoder.pointnet2")
    files = sorted({m.__file__ for m in list(sys.modules.values())
                    if getattr(m, "__file__", None) and m.__file__.startswith(SRC_ROOT + os.sep)})
    manifest = {}
    for f in files + [os.path.join(SRC_ROOT, e) for e in EXTRA]:
        if not os.path.exists(f):
            continue
        rel = os.path.relpath(f, SRC_ROOT)
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(f, dst)
        manifest[rel] = hashlib.sha256(open(f, "rb").read()).hexdigest()
        if verbose:
            print("vendored", rel)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as fh:
        json.dump({"source": SRC_ROOT, "files": manifest}, fh, indent=1, sort_keys=True)
    return OUT


if __name__ == "__main__":
    print(vendor(verbose="-v" in sys.argv))

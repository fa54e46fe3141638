"""TEST INFRASTRUCTURE ONLY -- imports the *reference* (PythonerJOJO/GenPose2) from
/root/reference inside THIS container so that golden vectors can be generated and
the oracle restatements can be pinned against it.

Nothing in the product path (genpose2_b200/) may import this module.  /root/reference does not
exist on the GPU box; there the -m gpu parity tests and bench.py's reference legs import the
byte-for-byte copy of the same modules that oracle/vendor_ref.py put into oracle/_ref/refpkg
(next to the reference's CUDA extension built by oracle/build_ref_ext.py).

The reference parses argv at import time (configs/config.py, executed from
networks/pts_encoder/pointnet2.py:28) and imports a handful of packages that are not
installed here (ipdb, tensorboardX, cutoop, matplotlib); those are stubbed through
sys.modules exactly as SURVEY.md section 8(c) describes.  The reference CUDA extension
`pointnet2_cuda` is stubbed too unless oracle/_ref holds a built copy (GPU box only).
"""
import os
import sys
import types

_REF_EXT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
# The build container has the reference tree; the GPU box only has the byte-for-byte copy of the hot-path
# modules that oracle/vendor_ref.py placed in oracle/_ref/refpkg (git-ignored, travels with the snapshot).
VENDORED_ROOT = os.path.join(_REF_EXT_DIR, "refpkg")


def _default_root():
    env = os.environ.get("GENPOSE2_REFERENCE_ROOT")
    if env:
        return env
    if os.path.isdir("/root/reference/networks"):
        return "/root/reference"
    return VENDORED_ROOT


REFERENCE_ROOT = _default_root()


class _Anything(types.ModuleType):
    """A module whose every attribute is a harmless callable/class stub."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)

        class _Stub:
            def __init__(self, *a, **k):
                pass

            def __call__(self, *a, **k):
                return None

            def __getattr__(self, n):
                return _Stub()

        _Stub.__name__ = name
        return _Stub


def install_stubs():
    for name in (
        "ipdb",
        "tensorboardX",
        "cutoop",
        "cutoop.rotation",
        "cutoop.eval_utils",
        "cutoop.data_loader",
        "cutoop.transform",
        "cutoop.utils",
        "cutoop.data_types",
        "matplotlib",
        "matplotlib.pyplot",
        "matplotlib.cm",
        "cv2",
        "open3d",
    ):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Anything(name)
    if "pointnet2_cuda" not in sys.modules:
        if os.path.isdir(_REF_EXT_DIR) and _REF_EXT_DIR not in sys.path:
            sys.path.insert(0, _REF_EXT_DIR)
        try:
            import torch  # noqa: F401  (the ext links against libtorch)
            __import__("pointnet2_cuda")
        except Exception:
            sys.modules["pointnet2_cuda"] = _Anything("pointnet2_cuda")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "networks"))


def has_cuda_ext():
    """True when the reference's own CUDA extension (oracle/_ref/pointnet2_cuda.so) is importable: the reference
    encoder can then run for real (GPU box)."""
    install_stubs()
    return not isinstance(sys.modules.get("pointnet2_cuda"), _Anything)


def load(argv=("x", "--dino", "none", "--sampler_mode", "ode")):
    """Return a namespace with the reference modules on the hot path."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    old_argv = sys.argv
    sys.argv = list(argv)
    try:
        import importlib

        ns = types.SimpleNamespace()
        ns.config = importlib.import_module("configs.config")
        ns.sde = importlib.import_module("networks.gf_algorithms.sde")
        ns.samplers = importlib.import_module("networks.gf_algorithms.samplers")
        ns.scorenet = importlib.import_module("networks.gf_algorithms.scorenet")
        ns.energynet = importlib.import_module("networks.gf_algorithms.energynet")
        ns.scalenet = importlib.import_module("networks.scalenet")
        ns.reward = importlib.import_module("networks.reward")
        ns.misc = importlib.import_module("utils.misc")
        ns.genpose_utils = importlib.import_module("utils.genpose_utils")
        ns.rotconv = importlib.import_module("utils.transforms.rotation_conversions")
        ns.posenet = importlib.import_module("networks.posenet")
        ns.posenet_agent = importlib.import_module("networks.posenet_agent")
        ns.pointnet2 = importlib.import_module("networks.pts_encoder.pointnet2")
        ns.pointnet2_utils = importlib.import_module(
            "networks.pts_encoder.pointnet2_utils.pointnet2.pointnet2_utils")
        ns.cfg = ns.config.get_config()
    finally:
        sys.argv = old_argv
    return ns

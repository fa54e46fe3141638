"""TEST INFRASTRUCTURE ONLY -- builds the *reference's own* PointNet++ CUDA extension
(`pointnet2_cuda`, /root/reference/networks/pts_encoder/pointnet2_utils/pointnet2/src)
from the sources where they lie, with the reference's flags (setup.py:7-21: cxx -g,
nvcc -O2), for sm_100, into oracle/_ref/ (git-ignored, travels to the GPU box).

It is the final arbiter for "bit-exact FPS / ball-query indices" in the -m gpu tests
and is never imported by the product path.  No reference source is copied.
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRC = "/root/reference/networks/pts_encoder/pointnet2_utils/pointnet2/src"
FILES = [
    "pointnet2_api.cpp",
    "ball_query.cpp",
    "ball_query_gpu.cu",
    "group_points.cpp",
    "group_points_gpu.cu",
    "interpolate.cpp",
    "interpolate_gpu.cu",
    "sampling.cpp",
    "sampling_gpu.cu",
]


def build(verbose=False):
    if not os.path.isdir(SRC):
        return None
    os.makedirs(OUT, exist_ok=True)
    so = [f for f in os.listdir(OUT) if f.startswith("pointnet2_cuda") and f.endswith(".so")]
    if so:
        return os.path.join(OUT, so[0])
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.setdefault("MAX_JOBS", "8")
    from torch.utils.cpp_extension import load

    load(
        name="pointnet2_cuda",
        sources=[os.path.join(SRC, f) for f in FILES],
        extra_cflags=["-g"],
        extra_cuda_cflags=["-O2"],
        build_directory=OUT,
        verbose=verbose,
        is_python_module=False,  # just build; importing needs no GPU but keep it inert
    )
    so = [f for f in os.listdir(OUT) if f.startswith("pointnet2_cuda") and f.endswith(".so")]
    return os.path.join(OUT, so[0]) if so else None


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))

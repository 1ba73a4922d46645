"""Hash of the CUDA sources libmsda_b200.so is built from: stored beside profiles/ncu_summary.json so that bench.py can
say whether its `roofline.traffic` (an ncu capture) belongs to the build it is timing."""
import glob
import hashlib
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "depth-fusion-in-transformer-based-video-object-detection_b200", "csrc")


def gather_kernels_hash():
    """The drop-in op's kernels only (forward / backward gather + their shared header)."""
    h = hashlib.sha256()
    for name in ("msda_common.cuh", "msda_forward.cu", "msda_backward.cu"):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


if __name__ == "__main__":
    print(gather_kernels_hash())

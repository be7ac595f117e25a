"""Recipe for ``oracle/_ref/``: an UNMODIFIED copy of the reference files on the hot path, made from /root/reference where
it exists (the build container).  ``oracle/_ref/`` is git-ignored (no reference source enters the history) but not
gpurun-ignored, so it travels to the GPU box, where ``bench.py --impl reference`` and the ``cpu_baseline`` leg time the
reference's own ``ddnm_diffusion`` + ``Model`` + ``SuperResolution`` on the host cores (oracle/ref_runner.py).

    python -m oracle.make_ref            # (also run by __graft_entry__.build())

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py): nothing under ddnm_b200/ may import it.
"""
import filecmp
import os
import shutil

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
# the path's files + what they import
FILES = ["functions/__init__.py", "functions/svd_ddnm.py", "functions/svd_operators.py", "functions/ckpt_util.py",
         "guided_diffusion/models.py", "guided_diffusion/unet.py", "guided_diffusion/nn.py", "guided_diffusion/fp16_util.py",
         "guided_diffusion/logger.py", "guided_diffusion/script_util.py", "exp/inp_masks/mask.npy"]


def make_ref(verbose=False):
    """Copy the files byte for byte; returns the destination, or None when /root/reference is absent (GPU box: the prebuilt
    copy that travelled with the snapshot is used as is)."""
    if not os.path.isdir(REF):
        return DST if os.path.isdir(DST) else None
    for rel in FILES:
        src, dst = os.path.join(REF, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
            if verbose:
                print("copied", rel)
    return DST


def available():
    return all(os.path.exists(os.path.join(DST, rel)) for rel in FILES)


if __name__ == "__main__":
    print(make_ref(verbose=True))

"""TEST / BENCH INFRASTRUCTURE ONLY -- recipe that stages the UNMODIFIED reference under ``oracle/_ref/``.

The reference (ingra14m/RobIR, /root/reference) is pure Python: there is nothing to compile, so "building" it for the
GPU box means copying the files the hot path imports, byte for byte, into ``oracle/_ref/``.  That directory is
git-ignored (the reference's sources never enter this repository's history) but NOT gpurun-ignored, so it travels
to the B200 box like the in-tree ``librobir_b200.so`` does.  ``__graft_entry__.build()`` runs this where
/root/reference exists (the build container); on the GPU box the staged copy is used as it is.

Used by (and only by): ``bench.py --impl reference`` / ``--impl reference-cuda`` (the reference's own implementation
timed on the host cores / eagerly on the B200), and the ``-m gpu`` drop-in test that runs ``robir_b200.install()`` on a
live reference ``IDRNetwork``.  Nothing under ``robir_b200/`` reads it.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
SRC_DEFAULT = "/root/reference"

# what the stage-2 hot path (and the stage-1 renderer the f4 row is checked against) imports
DIRS = ["model", "utils", "training", "confs_sg", "datasets"]
FILES = ["neus/volume_render/sdf_render.py", "envmaps/envmap3/sg_128.npy", "envmaps/envmap6/sg_128.npy",
         "envmaps/envmap12/sg_128.npy", "LICENSE"]
SUFFIXES = (".py", ".conf", ".npy")


def stage(src=SRC_DEFAULT, dest=DEST):
    """Copy the file list; returns the number of files staged.  Raises if the reference tree is missing."""
    if not os.path.isdir(os.path.join(src, "model")):
        raise RuntimeError("reference tree not present at %s" % src)
    n = 0
    for d in DIRS:
        for root, _, names in os.walk(os.path.join(src, d)):
            for name in names:
                if name.endswith(SUFFIXES):
                    rel = os.path.relpath(os.path.join(root, name), src)
                    n += _copy(os.path.join(src, rel), os.path.join(dest, rel))
    for rel in FILES:
        n += _copy(os.path.join(src, rel), os.path.join(dest, rel))
    with open(os.path.join(dest, "STAGED_FROM"), "w") as f:
        f.write("byte-for-byte copy of the hot-path files of %s made by oracle/stage_ref.py; not part of the repository\n"
                % src)
    return n


def _copy(a, b):
    os.makedirs(os.path.dirname(b), exist_ok=True)
    if not (os.path.exists(b) and filecmp.cmp(a, b, shallow=False)):
        shutil.copyfile(a, b)
        os.chmod(b, 0o644)
    return 1


def staged() -> bool:
    return os.path.isdir(os.path.join(DEST, "model"))


if __name__ == "__main__":
    print("staged %d files into %s" % (stage(*(sys.argv[1:2] or [SRC_DEFAULT])), DEST))

"""Place the UNMODIFIED reference files of the hot path under the git-ignored ``baseline/_ref/`` so that
``bench.py --impl reference`` and the same-box torch-GPU baseline can run the reference itself on the GPU box
(``/root/reference`` does not exist there; ``baseline/_ref`` travels with the snapshot).

The contract's ``pip install --target baseline/_ref /root/reference`` fails ("Neither 'setup.py' nor 'pyproject.toml'
found": the reference is a script tree, not a package), so the files are copied verbatim, tree preserved.  Nothing is
copied into tracked paths.  Run in the build container:  python scripts/install_reference.py
"""
import hashlib
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("SAM2_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = [
    "sam2_video/model/losses.py",
    "sam2_video/model/modeling/memory_attention.py",
    "sam2_video/model/modeling/memory_encoder.py",
    "sam2_video/model/modeling/position_encoding.py",
    "sam2_video/model/modeling/sam2_utils.py",
    "sam2_video/model/modeling/sam/transformer.py",
    "configs/sam2/sam2.1_hiera_t.yaml",
]


def install(verbose=True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print(f"install_reference: {SRC} not present (GPU box?) -- keeping whatever is in {DST}")
        return os.path.isfile(os.path.join(DST, FILES[1]))
    manifest = []
    for rel in FILES:
        src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest.append(f"{hashlib.sha256(open(dst, 'rb').read()).hexdigest()}  {rel}")
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    if verbose:
        print(f"install_reference: {len(FILES)} files -> {DST}")
    return True


if __name__ == "__main__":
    sys.exit(0 if install() else 1)

"""Install the UNMODIFIED reference modules for the CPU arm of bench.py (``--impl reference`` / ``cpu_baseline``).

The reference (Wadaboa/titanet @ 7b77053) is a flat directory of Python files with no setup.py / pyproject.toml, so
``pip install --target baseline/_ref /root/reference`` has nothing to build ("neither 'setup.py' nor 'pyproject.toml'
found").  The install is therefore a byte-for-byte copy of the four hot-path modules (and parameters.yml, which names the
defaults the bench uses) from the reference checkout into ``baseline/_ref/`` -- a git-ignored directory (the reference's
sources never enter this repository's history) that is NOT gpurun-ignored, so it travels to the GPU box, where
``/root/reference`` does not exist.  ``__graft_entry__.build()`` runs this when the checkout is present.

    python baseline/install_ref.py [/root/reference]
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ["src/modules.py", "src/models.py", "src/losses.py", "src/transforms.py", "parameters.yml"]


def install(ref_root: str = "/root/reference") -> bool:
    if not os.path.isdir(os.path.join(ref_root, "src")):
        return False
    os.makedirs(DEST, exist_ok=True)
    manifest = {}
    for rel in FILES:
        src = os.path.join(ref_root, rel)
        dst = os.path.join(DEST, os.path.basename(rel))
        shutil.copyfile(src, dst)
        manifest[os.path.basename(rel)] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": ref_root, "sha256": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    ok = install(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("installed" if ok else "reference checkout not found; nothing installed", DEST)

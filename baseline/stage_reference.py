"""Stages the UNMODIFIED reference package under the git-ignored ``baseline/_ref/`` so that it travels to the GPU box
(git-ignored files do, ``.gpurunignore``d files do not) and the reference arm of ``bench.py`` and the reference-wrapper
GPU tests can import the real ``cwm`` there.  Run in the build container (needs the read-only /root/reference mount);
``__graft_entry__.build()`` calls it when the mount exists.

Step 1 is the contract's offline install,
    python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
        --target baseline/_ref <copy of /root/reference under /tmp>
(from a /tmp copy because the build writes ``build/`` and ``*.egg-info`` into the source tree and /root/reference is
read-only; ``--no-deps`` because the pinned ``matplotlib==3.5.2``, ``timm``, ``kornia`` ... are not in the wheelhouse).
Outcome: the wheel builds and installs, but it only contains ``cwm/{__init__,interface,version,vis_utils}.py`` --
``cwm/models`` and ``cwm/data`` have no ``__init__.py`` (they are namespace packages) and ``setup.py`` uses
``find_packages()``, which skips them.  Step 2 therefore completes the installed tree with the missing sub-packages,
copied byte for byte (``*.py`` only; notebooks, images and checkpoints are not needed).  Nothing is edited.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference"
DST = os.path.join(ROOT, "baseline", "_ref")


def stage(verbose=True):
    if not os.path.isdir(os.path.join(SRC, "cwm")):
        return False
    marker = os.path.join(DST, "cwm", "models", "VideoMAE", "vmae.py")
    if os.path.exists(marker):
        return True
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(DST, exist_ok=True)
    pip_ok = False
    with tempfile.TemporaryDirectory() as tmp:
        work = os.path.join(tmp, "reference")
        shutil.copytree(SRC, work, ignore=shutil.ignore_patterns("demo", "*.png", ".git", ".#*"))
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--quiet",
               "--find-links", "/opt/wheelhouse", "--target", DST, work]
        pip_ok = subprocess.run(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode == 0
    n = 0
    for dirpath, dirnames, filenames in os.walk(os.path.join(SRC, "cwm")):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        rel = os.path.relpath(dirpath, SRC)
        for fn in filenames:
            if not fn.endswith(".py") or fn.startswith("."):
                continue
            out = os.path.join(DST, rel, fn)
            if os.path.exists(out):
                continue
            os.makedirs(os.path.dirname(out), exist_ok=True)
            shutil.copyfile(os.path.join(dirpath, fn), out)
            n += 1
    with open(os.path.join(DST, "STAGED.txt"), "w") as f:
        f.write(f"pip install --target: {'ok' if pip_ok else 'failed'}; {n} module files of the namespace "
                "sub-packages (cwm/models, cwm/data) added unmodified from /root/reference\n")
    if verbose:
        print(f"baseline/_ref: pip {'ok' if pip_ok else 'FAILED'}, +{n} files")
    return True


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)

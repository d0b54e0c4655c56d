"""Recipe: put the UNMODIFIED reference where bench.py can time it on the GPU box.

    python baseline/install_reference.py        (build container only: needs /root/reference)

1. ``pip install --no-index --no-build-isolation --no-deps --target baseline/_ref`` of a scratch copy of
   /root/reference (the tree itself is read-only and setuptools writes build/ next to setup.py).  ``--no-deps``:
   moorpy / pyhams / ccblade / wisdem / openmdao / matplotlib are not in the wheelhouse; none of them is on the
   per-frequency hot path, and oracle/ref_harness.py stubs their import lines.
2. copy the design INPUTS the baseline configurations read (YAML files and the WAMIT coefficient tables of
   configs[2]; data, not code) to baseline/_ref/inputs/.

baseline/_ref/ is git-ignored (never part of the history) but not gpurun-ignored, so it travels to the GPU box with
the snapshot like the built .so files.  Nothing under raft_b200/ reads it; only bench.py's CPU baseline legs do
(oracle/ref_timing.py).  Outcome of the install is recorded in DESIGN.md section 7.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("RAFT_REFERENCE_SRC", "/root/reference")
DST = os.path.join(HERE, "_ref")
INPUTS = ["designs/OC3spar.yaml", "designs/VolturnUS-S.yaml", "designs/VolturnUS-S_farm.yaml",
          "examples/OC4semi-WAMIT_Coefs.yaml", "examples/OC4semi-WAMIT_Coefs/marin_semi.1", "examples/OC4semi-WAMIT_Coefs/marin_semi.3"]


def installed():
    return os.path.isdir(os.path.join(DST, "raft")) and os.path.isdir(os.path.join(DST, "inputs"))


def main():
    if not os.path.isdir(os.path.join(REF, "raft")):
        print("reference tree %s not present: nothing to install" % REF)
        return 1
    if installed():
        print("baseline/_ref already installed")
        return 0
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF, src, ignore=shutil.ignore_patterns("docs", ".git", "examples"))
        subprocess.check_call(["chmod", "-R", "u+w", src])
        subprocess.check_call([sys.executable, "-m", "pip", "install", "-q", "--no-index", "--no-build-isolation", "--no-deps",
                               "--find-links", "/opt/wheelhouse", "--target", DST, "--upgrade", src])
    for rel in INPUTS:
        s = os.path.join(REF, rel)
        if os.path.exists(s):
            d = os.path.join(DST, "inputs", rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
    print("installed the unmodified reference into", DST)
    return 0


if __name__ == "__main__":
    sys.exit(main())

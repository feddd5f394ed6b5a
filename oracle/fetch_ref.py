"""oracle/fetch_ref.py -- stages the UNMODIFIED reference hot path under git-ignored oracle/_ref/.

Test / measurement infrastructure only (never imported by phnn_mpc_b200/).  The reference is pure
Python: nothing is compiled, the files are copied byte for byte from the read-only checkout so that
the GPU box (which has no /root/reference) can time and check against the reference's own PyTorch
implementation (`bench.py --impl reference`, the `cpu_baseline` leg, the unchanged-driver tests).
oracle/_ref/ is listed in .gitignore (the sources never enter this repository's history) but not in
.gpurunignore, so it travels with the tree like the built .so files.

    python oracle/fetch_ref.py [--ref /root/reference]

Copied: the 8 hot-path modules + the plant, the two MPC drivers the
unchanged-driver test runs (scripts/run_cartpole_mpc.py, scripts/run_mpc_canonical.py), the three
YAML configs and the shipped pendulum checkpoint.  MANIFEST.json records sha256 of every file.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
DEFAULT_REF = "/root/reference"

FILES = [
    "src/pHNN.py", "src/pHNN_canonical.py", "src/NN.py", "src/mass_matrix.py", "src/coordinate_transforms.py",
    "src/integrators.py", "src/mpc_controller.py", "src/mpc_controller_canonical.py", "src/cartpole_simulator.py",
    "scripts/run_cartpole_mpc.py", "scripts/run_mpc_canonical.py",
    "cartpole_mpc_config.yaml", "pole_stabilization_config.yaml", "pendulum_config.yaml",
    "pendulum_pHNN_weights.pth",
]


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 20), b""):
            h.update(blk)
    return h.hexdigest()


def available():
    """True when a complete staged copy exists (every file of the manifest present)."""
    man = os.path.join(DEST, "MANIFEST.json")
    if not os.path.exists(man):
        return False
    try:
        m = json.load(open(man))
    except Exception:
        return False
    return all(os.path.exists(os.path.join(DEST, f)) for f in m.get("files", {}))


def fetch(ref=DEFAULT_REF, quiet=False):
    """Copy the files (no-op when the reference checkout is absent and a staged copy exists)."""
    if not os.path.isdir(ref):
        if available():
            return DEST
        raise RuntimeError("reference checkout %s not found and oracle/_ref is not staged" % ref)
    manifest = {"source": ref, "files": {}}
    for rel in FILES:
        src = os.path.join(ref, rel)
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not os.path.exists(dst) or sha256(dst) != sha256(src):
            shutil.copyfile(src, dst)
        manifest["files"][rel] = sha256(dst)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    if not quiet:
        print("staged %d reference files under %s" % (len(FILES), DEST))
    return DEST


if __name__ == "__main__":
    ref = DEFAULT_REF
    if "--ref" in sys.argv:
        ref = sys.argv[sys.argv.index("--ref") + 1]
    fetch(ref)

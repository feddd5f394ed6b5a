"""Checkpoint <-> packed-weights file.

* ``load_reference_checkpoint(path)``: the reference's ``.pth`` layouts -- a raw ``state_dict`` or a dict with
  ``'model_state_dict'`` (scripts/run_cartpole_mpc.py:38-44) -- to a plain ``{name: float32 ndarray}`` mapping.
* ``save_packed`` / ``load_packed``: a flat little-endian binary file that needs no pickle and maps 1:1 onto
  ``phnn_model_desc`` (include/phnn_mpc.h):

      magic  b"PHNNPK01"
      int32  kind (0 pHNN, 1 canonical), n, m, h, learned_G, n_arrays
      per array: int32 name_len, name (utf-8), int32 ndim, int32 dims[ndim], float32 data (C order)

  Array names are the reference's state_dict keys, so a packed file round-trips to a state_dict that
  ``load_state_dict`` of the (reference or drop-in) modules accepts.
* ``PackedModel.from_file`` (packing.py) builds the device image straight from such a file.
"""
import struct

import numpy as np

MAGIC = b"PHNNPK01"


def _to_np(v):
    if hasattr(v, "detach"):
        v = v.detach().cpu().numpy()
    return np.asarray(v, dtype=np.float32, order="C")  # (ascontiguousarray would turn 0-d scalars into shape (1,))


def load_reference_checkpoint(path):
    """torch.load with weights_only=True; accepts both layouts the reference's drivers accept."""
    import torch
    obj = torch.load(path, map_location="cpu", weights_only=True)
    if isinstance(obj, dict) and "model_state_dict" in obj:
        obj = obj["model_state_dict"]
    if not isinstance(obj, dict):
        raise ValueError("%s does not hold a state_dict" % path)
    return {k: _to_np(v) for k, v in obj.items()}


def describe(state_dict):
    from .packing import normalize_state_dict
    sd = normalize_state_dict(state_dict)
    kind = 1 if "R_diag_raw" in sd else 0
    W1 = _to_np(sd["H_net.net.0.weight"])
    h, n = int(W1.shape[0]), int(W1.shape[1])
    if kind == 1:
        m, learned = int(_to_np(sd["G"]).shape[1]), 0
    elif "G_fixed" in sd:
        m, learned = int(_to_np(sd["G_fixed"]).shape[1]), 0
    else:
        m, learned = int(_to_np(sd["G_net.net.2.weight"]).shape[0]) // n, 1
    return kind, n, m, h, learned


def save_packed(path, state_dict):
    sd = {k: _to_np(v) for k, v in state_dict.items()}
    kind, n, m, h, learned = describe(sd)
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<6i", kind, n, m, h, learned, len(sd)))
        for name in sorted(sd):
            a = sd[name]
            nb = name.encode("utf-8")
            f.write(struct.pack("<i", len(nb)))
            f.write(nb)
            f.write(struct.pack("<i", a.ndim))
            f.write(struct.pack("<%di" % a.ndim, *a.shape))
            f.write(a.astype("<f4").tobytes())
    return path


def load_packed(path):
    with open(path, "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError("%s is not a PHNNPK01 file" % path)
        kind, n, m, h, learned, count = struct.unpack("<6i", f.read(24))
        sd = {}
        for _ in range(count):
            (ln,) = struct.unpack("<i", f.read(4))
            name = f.read(ln).decode("utf-8")
            (nd,) = struct.unpack("<i", f.read(4))
            dims = struct.unpack("<%di" % nd, f.read(4 * nd)) if nd else ()
            cnt = int(np.prod(dims)) if nd else 1
            sd[name] = np.frombuffer(f.read(4 * cnt), dtype="<f4").reshape(dims).copy()
    if describe(sd) != (kind, n, m, h, learned):
        raise ValueError("%s: header does not match its arrays" % path)
    return sd, {"kind": "canonical" if kind else "phnn", "n": n, "m": m, "h": h, "learned_G": bool(learned)}

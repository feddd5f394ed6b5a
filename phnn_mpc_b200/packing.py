"""Weights of a pHNN / pHNN_Canonical module -> packed device image (phnn_pack).

Accepts any module (or state_dict) with the reference's parameter names
(src/pHNN.py:13-38: J, G_fixed | G_net.net.{0,2}.*, R_net.net.{0,2}.*, H_net.net.{0,2,4}.*;
 src/pHNN_canonical.py:57-110: J, G, R_diag_raw, M_net.{log_a,b,log_c}, H_net.net.{0,2,4}.*),
so reference checkpoints load unchanged.
"""
import ctypes
import weakref

import numpy as np
import torch

from . import _lib


def _f32(t):
    if isinstance(t, torch.Tensor):
        t = t.detach().to("cpu", torch.float32).numpy()
    return np.ascontiguousarray(np.asarray(t, dtype=np.float32))


def infer_kind(sd):
    return "canonical" if "R_diag_raw" in sd else "phnn"


def normalize_state_dict(sd):
    """Reference MLPs built with ``dropout > 0`` (src/NN.py:16-25) interleave ``nn.Dropout`` modules, which shifts the
    Sequential indices of the Linear layers (``net.0, net.3, net.6`` instead of ``net.0, net.2, net.4``).  Dropout is the
    identity in eval mode -- the mode both controllers put the model in (src/mpc_controller.py:44,
    src/mpc_controller_canonical.py:54) -- so such a checkpoint is the same function: rename the k-th Linear of every net
    to ``net.<2k>``.  A LayerNorm (1-D ``weight`` under ``net.<i>``) is not the identity and raises."""
    out = dict(sd)
    for prefix in ("H_net", "R_net", "G_net"):
        idx = sorted({int(k.split(".")[2]) for k in sd if k.startswith(prefix + ".net.") and k.endswith(".weight")})
        if not idx:
            continue
        for i in idx:
            w = sd["%s.net.%d.weight" % (prefix, i)]
            if len(tuple(w.shape)) != 2:
                raise NotImplementedError("%s.net.%d is not a Linear layer (LayerNorm?): not built for the CUDA path" % (prefix, i))
        if idx != [2 * j for j in range(len(idx))]:
            for k in [k for k in out if k.startswith(prefix + ".net.")]:
                del out[k]
            for j, i in enumerate(idx):
                for leaf in ("weight", "bias"):
                    src = "%s.net.%d.%s" % (prefix, i, leaf)
                    if src in sd:
                        out["%s.net.%d.%s" % (prefix, 2 * j, leaf)] = sd[src]
        # ``bias: false`` (src/NN.py:19,25): a Linear without bias is the same function as one with a zero bias
        for j in range(len(idx)):
            if "%s.net.%d.bias" % (prefix, 2 * j) not in out:
                out["%s.net.%d.bias" % (prefix, 2 * j)] = np.zeros(int(out["%s.net.%d.weight" % (prefix, 2 * j)].shape[0]), np.float32)
    return out


def constant_mass_abc(L_tril):
    """(a, b, c) of the constant mass matrix M = L L^T = [[a, b], [b, c]] of MassMatrixNetwork(mass_type='constant')
    (src/mass_matrix.py:130-147): L = tril(L_tril) with softplus(diag) + 1e-3 on the diagonal, float32 arithmetic."""
    L = torch.tril(torch.as_tensor(L_tril).detach().float().cpu())
    if tuple(L.shape) != (2, 2):
        raise NotImplementedError("constant mass matrix: q_dim = 2 (the cart-pole case the kernels are built for)")
    l00 = torch.nn.functional.softplus(L[0, 0]) + 1e-3
    l11 = torch.nn.functional.softplus(L[1, 1]) + 1e-3
    l10 = L[1, 0]
    return float(l00 * l00), float(l10 * l00), float(l10 * l10 + l11 * l11)


class PackedModel:
    """Owns one phnn_pack handle (immutable device copy of the weights)."""

    def __init__(self, state_dict, kind=None, device=None):
        L = _lib.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("phnn_mpc_b200 needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("phnn_mpc_b200 runs on CUDA devices only, got %s" % self.device)
        sd = normalize_state_dict(dict(state_dict))
        kind = kind or infer_kind(sd)
        keep = []

        def ptr(name):
            a = _f32(sd[name])
            keep.append(a)
            return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))

        for i in (0, 2, 4):
            if "H_net.net.%d.weight" % i not in sd:
                raise RuntimeError("H_net must be Linear-Tanh-Linear-Tanh-Linear (two hidden layers, no LayerNorm/"
                                   "Dropout); key H_net.net.%d.weight missing" % i)
        if "H_net.net.6.weight" in sd:
            raise RuntimeError("H_net with more than two hidden layers is not supported by the CUDA kernels")
        W1, W2 = sd["H_net.net.0.weight"], sd["H_net.net.2.weight"]
        d = _lib.ModelDesc()
        d.n, d.h = int(W1.shape[1]), int(W1.shape[0])
        if tuple(W2.shape) != (d.h, d.h):
            raise RuntimeError("H_net hidden layers must have equal width, got %s" % (tuple(W2.shape),))
        d.W1, d.b1 = ptr("H_net.net.0.weight"), ptr("H_net.net.0.bias")
        d.W2, d.b2 = ptr("H_net.net.2.weight"), ptr("H_net.net.2.bias")
        d.W3, d.b3 = ptr("H_net.net.4.weight"), ptr("H_net.net.4.bias")
        d.J = ptr("J")
        if kind == "phnn":
            d.kind = _lib.PHNN_KIND_PHNN
            if "R_net.net.4.weight" in sd or int(sd["R_net.net.0.weight"].shape[0]) != d.h:
                raise RuntimeError("R_net must have one hidden layer of the same width as H_net (%d)" % d.h)
            d.Wr1, d.br1 = ptr("R_net.net.0.weight"), ptr("R_net.net.0.bias")
            d.Wr2, d.br2 = ptr("R_net.net.2.weight"), ptr("R_net.net.2.bias")
            if "G_fixed" in sd:
                d.learned_G = 0
                d.m = int(sd["G_fixed"].shape[1])
                d.G = ptr("G_fixed")
            else:
                d.learned_G = 1
                if "G_net.net.4.weight" in sd or int(sd["G_net.net.0.weight"].shape[0]) != d.h:
                    raise RuntimeError("G_net must have one hidden layer of the same width as H_net (%d)" % d.h)
                d.m = int(sd["G_net.net.2.weight"].shape[0]) // d.n
                d.Wg1, d.bg1 = ptr("G_net.net.0.weight"), ptr("G_net.net.0.bias")
                d.Wg2, d.bg2 = ptr("G_net.net.2.weight"), ptr("G_net.net.2.bias")
        elif kind == "canonical":
            d.kind = _lib.PHNN_KIND_CANONICAL
            d.learned_G = 0
            d.m = int(sd["G"].shape[1])
            d.G = ptr("G")
            # src/mass_matrix.py:283-285 (float32 arithmetic) and src/pHNN_canonical.py:162
            if "M_net.L_tril" in sd:
                # MassMatrixNetwork 'constant' (src/mass_matrix.py:130-147): M = L L^T, diag(L) = softplus(.) + 1e-3
                d.mass_a, d.mass_b, d.mass_c = constant_mass_abc(sd["M_net.L_tril"])
                d.mass_const = 1
            elif any(k.startswith("M_net.mlp.") for k in sd):
                raise NotImplementedError("configuration-dependent MassMatrixNetwork ('diagonal' / 'full') is not built; "
                                          "the kernels cover the cart-pole and the constant mass matrix")
            else:
                la = torch.as_tensor(sd["M_net.log_a"]).detach().float()
                lc = torch.as_tensor(sd["M_net.log_c"]).detach().float()
                d.mass_a = float(torch.exp(la) + 1e-3)
                d.mass_b = float(torch.as_tensor(sd["M_net.b"]).detach().float())
                d.mass_c = float(torch.exp(lc) + 1e-3)
            rd = torch.nn.functional.softplus(torch.as_tensor(sd["R_diag_raw"]).detach().float()) + 1e-4
            sd["__r_diag"] = rd
            d.r_diag = ptr("__r_diag")
        else:
            raise ValueError("unknown model kind %r" % (kind,))
        self.kind, self.n, self.m, self.h = kind, d.n, d.m, d.h
        handle = ctypes.c_void_p()
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        _lib.check(L.phnn_pack_create(ctypes.byref(d), self.device.index, ctypes.byref(handle)),
                   "phnn_pack_create")
        self.handle = handle.value
        self._fin = weakref.finalize(self, L.phnn_pack_destroy, ctypes.c_void_p(self.handle))

    @classmethod
    def from_file(cls, path, device=None):
        """device image straight from a PHNNPK01 file or a reference ``.pth`` checkpoint (weights_io.py)"""
        from . import weights_io
        with open(path, "rb") as f:
            packed = f.read(8) == weights_io.MAGIC
        sd = weights_io.load_packed(path)[0] if packed else weights_io.load_reference_checkpoint(path)
        return cls({k: torch.from_numpy(v) for k, v in sd.items()}, device=device)

    def set_option(self, key, value):
        """kernel selection knobs (include/phnn_mpc.h): 'tensor_mode' (0 FP32-FMA, 4 tcgen05 3 x FP16 hi/lo with A in TMEM
        (default), 2 tcgen05 TF32 + BF16 correction, 3 tcgen05 3xTF32, 1 tcgen05 TF32), 'tensor_min_batch',
        'tensor_fwd_min_batch' (n = 2 models: forward-only tcgen05 kernel), 'latency_max_batch'"""
        _lib.check(_lib.lib().phnn_pack_set_option(ctypes.c_void_p(self.handle), key.encode(), int(value)),
                   "phnn_pack_set_option")

    def get_option(self, key):
        return int(_lib.lib().phnn_pack_get_option(ctypes.c_void_p(self.handle), key.encode()))

    def __int__(self):
        return self.handle


_cache = weakref.WeakKeyDictionary()


def _fingerprint(t):
    """cheap content fingerprint: catches in-place writes through ``param.data`` (``p.data.copy_``, EMA code, older
    optimizers), which bump neither ``_version`` nor ``data_ptr``.  The models here are < 0.3 MB of weights."""
    d = t.detach()
    if d.numel() == 0:
        return (0.0, 0.0)
    d = d.double()
    return (float(d.sum()), float(d.abs().sum()))


def invalidate(module):
    """drop the cached device image of `module` (the next pack_of repacks)"""
    _cache.pop(module, None)


def pack_of(module, device=None):
    """Packed image of an nn.Module, rebuilt when any parameter/buffer changed (tracked by the tensors' _version
    counters, data pointers and a content fingerprint)."""
    tensors = list(module.state_dict(keep_vars=True).items())
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type == "cuda" and dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    key = (str(dev),) + tuple((k, t._version, t.data_ptr()) + _fingerprint(t) for k, t in tensors)
    hit = _cache.get(module)
    if hit is not None and hit[0] == key:
        return hit[1]
    pk = PackedModel({k: t for k, t in tensors}, device=dev)
    _cache[module] = (key, pk)
    return pk

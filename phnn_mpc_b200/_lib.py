"""ctypes binding of libphnn_mpc.so (include/phnn_mpc.h).  There is no fallback: if the CUDA
library is missing or a call fails, the caller gets an exception."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PHNN_MPC_LIB points the binding at another build of the same library (A/B experiments, tools/gpu_ab.sh)
LIB_PATH = os.environ.get("PHNN_MPC_LIB") or os.path.join(_HERE, "libphnn_mpc.so")

PHNN_KIND_PHNN, PHNN_KIND_CANONICAL = 0, 1
PHNN_EULER, PHNN_RK4 = 0, 1
E_ARG, E_UNSUPPORTED, E_INTEGRATOR, E_WORKSPACE = -1, -2, -3, -4

_fp = ctypes.POINTER(ctypes.c_float)


class ModelDesc(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int), ("n", ctypes.c_int), ("m", ctypes.c_int), ("h", ctypes.c_int),
                ("learned_G", ctypes.c_int),
                ("W1", _fp), ("b1", _fp), ("W2", _fp), ("b2", _fp), ("W3", _fp), ("b3", _fp),
                ("Wr1", _fp), ("br1", _fp), ("Wr2", _fp), ("br2", _fp),
                ("Wg1", _fp), ("bg1", _fp), ("Wg2", _fp), ("bg2", _fp),
                ("J", _fp), ("G", _fp),
                ("mass_a", ctypes.c_float), ("mass_b", ctypes.c_float), ("mass_c", ctypes.c_float),
                ("r_diag", _fp), ("mass_const", ctypes.c_int)]


class CostDesc(ctypes.Structure):
    _fields_ = [("Q", _fp), ("R", _fp), ("x_target", _fp), ("has_u_bounds", ctypes.c_int),
                ("u_min", ctypes.c_float), ("u_max", ctypes.c_float), ("x_min", _fp), ("x_max", _fp),
                ("barrier_weight", ctypes.c_float)]


class Episode(ctypes.Structure):
    _fields_ = [("state", ctypes.c_void_p), ("traj", ctypes.c_void_p), ("controls", ctypes.c_void_p),
                ("done_step", ctypes.c_void_p), ("stable_start", ctypes.c_void_p), ("stable_duration", ctypes.c_void_p),
                ("stability_achieved", ctypes.c_void_p), ("steps", ctypes.c_int)]


class ParamGrads(ctypes.Structure):
    _fields_ = [(k, ctypes.c_void_p) for k in ("W1", "b1", "W2", "b2", "W3", "Wr1", "br1", "Wr2", "br2", "Wg1", "bg1", "Wg2",
                                               "bg2", "J", "r_diag")]


class PeerDesc(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int), ("offset", ctypes.c_longlong), ("U", ctypes.c_void_p * 8), ("cost", ctypes.c_void_p * 8)]


_lib = None


def lib():
    """Load the CUDA library; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "phnn_mpc_b200: %s is missing - build it with `python -m phnn_mpc_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, ll, ci, cd = ctypes.c_void_p, ctypes.c_long, ctypes.c_int, ctypes.c_double
    L.phnn_last_error.restype = ctypes.c_char_p
    L.phnn_version.restype = ci
    L.phnn_pack_create.argtypes = [ctypes.POINTER(ModelDesc), ci, ctypes.POINTER(vp)]
    L.phnn_pack_destroy.argtypes = [vp]
    L.phnn_pack_dims.argtypes = [vp] + [ctypes.POINTER(ci)] * 4
    L.phnn_forward.argtypes = [vp, vp, vp, vp, vp, ll, vp]
    L.phnn_vjp.argtypes = [vp, vp, vp, vp, vp, vp, ll, vp]
    L.phnn_rollout.argtypes = [vp, vp, vp, vp, vp, ll, ci, cd, ci, ci, vp]
    L.phnn_workspace_bytes.argtypes = [vp, ll, ci, ci]
    L.phnn_workspace_bytes.restype = ctypes.c_size_t
    L.phnn_cost_grad.argtypes = [vp, ctypes.POINTER(CostDesc), vp, vp, vp, vp, vp, ll, ci, cd, ci, vp,
                                 ctypes.c_size_t, vp]
    L.phnn_mpc_solve.argtypes = [vp, ctypes.POINTER(CostDesc), vp, vp, vp, vp, ll, ci, cd, ci, cd, cd, cd, cd, ci, ci,
                                 vp, ctypes.c_size_t, vp]
    L.phnn_mpc_solve_peer.argtypes = [vp, ctypes.POINTER(CostDesc), vp, vp, vp, vp, ll, ci, cd, ci, cd, cd, cd, cd, ci, ci,
                                      vp, ctypes.c_size_t, ctypes.POINTER(PeerDesc), vp]
    L.phnn_mpc_solve_peer.restype = ci
    L.phnn_peer_alloc.argtypes = [ctypes.c_size_t, ci, ctypes.POINTER(vp), ctypes.c_char_p]
    L.phnn_peer_open.argtypes = [ctypes.c_char_p, ci, ctypes.POINTER(vp)]
    L.phnn_peer_close.argtypes = [vp, ci]
    L.phnn_peer_free.argtypes = [vp, ci]
    for f in ("phnn_peer_alloc", "phnn_peer_open", "phnn_peer_close", "phnn_peer_free"):
        getattr(L, f).restype = ci
    L.phnn_rollout_vjp_workspace_bytes.argtypes = [vp, ll, ci, ci]
    L.phnn_rollout_vjp_workspace_bytes.restype = ctypes.c_size_t
    L.phnn_rollout_vjp.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.POINTER(ParamGrads), ll, ci, cd, ci, vp, ctypes.c_size_t, vp]
    L.phnn_rollout_vjp.restype = ci
    L.phnn_pack_set_option.argtypes = [vp, ctypes.c_char_p, ll]
    L.phnn_pack_set_option.restype = ci
    L.phnn_pack_get_option.argtypes = [vp, ctypes.c_char_p]
    L.phnn_pack_get_option.restype = ll
    dp = ctypes.POINTER(ctypes.c_double)
    L.phnn_plant_step.argtypes = [ctypes.POINTER(Episode), vp, ll, ci, cd, dp, dp, cd, ll, vp]
    L.phnn_plant_step.restype = ci
    L.phnn_state_to_f32.argtypes = [vp, vp, vp, ci, ll, vp]
    L.phnn_state_to_f32.restype = ci
    L.phnn_shift_controls.argtypes = [vp, vp, ll, ci, vp]
    L.phnn_shift_controls.restype = ci
    L.phnn_ffma_probe.argtypes = [vp, ci, ci, vp, ctypes.POINTER(cd)]
    L.phnn_ffma_probe.restype = ci
    L.phnn_tf32_probe.argtypes = [vp, ci, ci, vp, ctypes.POINTER(cd)]
    L.phnn_tf32_probe.restype = ci
    for f in ("phnn_pack_create", "phnn_pack_destroy", "phnn_pack_dims", "phnn_forward", "phnn_vjp", "phnn_rollout",
              "phnn_cost_grad", "phnn_mpc_solve"):
        getattr(L, f).restype = ci
    _lib = L
    return L


EXPORTS = ["phnn_last_error", "phnn_version", "phnn_pack_create", "phnn_pack_destroy", "phnn_pack_dims",
           "phnn_forward", "phnn_vjp", "phnn_rollout", "phnn_workspace_bytes", "phnn_cost_grad", "phnn_mpc_solve", "phnn_ffma_probe", "phnn_tf32_probe",
           "phnn_pack_set_option", "phnn_pack_get_option", "phnn_plant_step", "phnn_state_to_f32", "phnn_shift_controls",
           "phnn_rollout_vjp", "phnn_rollout_vjp_workspace_bytes", "phnn_mpc_solve_peer", "phnn_peer_alloc", "phnn_peer_open", "phnn_peer_close", "phnn_peer_free"]


def check(rc, what):
    """Map a C-ABI status to the reference's error conventions (ValueError for a bad integrator
    as src/integrators.py:172,226; RuntimeError otherwise)."""
    if rc == 0:
        return
    msg = lib().phnn_last_error().decode("utf-8", "replace")
    if rc == E_INTEGRATOR:
        raise ValueError(msg)
    if rc in (E_ARG, E_UNSUPPORTED, E_WORKSPACE):
        raise RuntimeError("%s: %s" % (what, msg))
    raise RuntimeError("%s: CUDA error %d: %s" % (what, rc, msg))

"""ctypes binding of librankaae_b200.so (the C ABI in include/rankaae_b200.h).

There is no fallback: if the shared library has not been built (`python -c "import __graft_entry__ as g;
g.build()"`), importing the compute path raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librankaae_b200.so")

MAX_LAYERS = 8
HIDDEN = 64
NUM_PHASES = 5
NUM_NETS = 3
ZPAD = 8
PHASES = ("adversarial", "correlation", "reconstruction", "mutual_info", "smoothness")
NETS = ("E", "D", "S")

HP_LR0, HP_BETA1, HP_BETA2, HP_WD = 0, 5, 10, 15
HP_DROPOUT, HP_DIS_DROPOUT, HP_DIS_NOISE, HP_SPEC_NOISE = 20, 21, 22, 23
HP_ALPHA_FLAT_STEP, HP_ALPHA_LIMIT, HP_SCH_FACTOR, HP_SCH_PATIENCE = 24, 25, 26, 27
HP_EPOCH_STOP_SMOOTH, HP_MAX_EPOCH, HP_SEED, HP_COUNT = 28, 29, 30, 32

_i32 = C.c_int32
_L = _i32 * MAX_LAYERS


class Config(C.Structure):
    _fields_ = [(n, _i32) for n in (
        "dim_in", "dim_out", "nstyle", "n_aux", "n_layers", "dis_layers", "batch_size", "n_trials",
        "kendall_activation", "use_flex_spec_target", "decoder_softplus", "max_rows", "ctas_per_trial",
        "tensor_cores")] + [("reserved", _i32 * 2)]


class NetLayout(C.Structure):
    _fields_ = [("n_linear", _i32), ("in_dim", _L), ("out_dim", _L), ("w_off", _L), ("b_off", _L), ("a_off", _L),
                ("rm_off", _L), ("rv_off", _L), ("param_off", _i32), ("n_params", _i32), ("nbt_off", _i32)]


class OptLayout(C.Structure):
    _fields_ = [("m_off", _i32), ("v_off", _i32), ("n", _i32), ("net_off", _i32 * NUM_NETS), ("scalar_off", _i32)]


class Layout(C.Structure):
    _fields_ = [("net", NetLayout * NUM_NETS), ("opt", OptLayout * NUM_PHASES), ("misc_off", _i32),
                ("state_floats", _i32), ("scratch_floats", _i32)]


_p = C.c_void_p


class DebugIO(C.Structure):
    _fields_ = [("x_noisy", _p), ("aux", _p), ("rows", _i32), ("epoch", _i32), ("phase_mask", _i32),
                ("apply_updates", _i32),
                ("mask_enc", (_p * MAX_LAYERS) * 6), ("mask_dec", (_p * MAX_LAYERS) * 4),
                ("mask_dis", (_p * MAX_LAYERS) * 2),
                ("z_real", _p), ("dis_eps_real", _p), ("dis_eps_fake", _p), ("z_sample", _p),
                ("losses", _p), ("grads", _p * NUM_PHASES), ("styles", _p)]


class ValIO(C.Structure):
    _fields_ = [("z_sample", _p), ("z_real", _p), ("epoch", _i32), ("avg_mutual_info", C.c_float),
                ("losses", _p), ("metrics", _p), ("z", _p), ("row_mae", _p), ("per_trial", _i32), ("reserved", _i32)]


EXPORTS = ("raae_last_error", "raae_version", "raae_query_layout", "raae_create", "raae_destroy", "raae_max_clusters", "raae_bind_state",
           "raae_bind_dataset", "raae_bind_shapiro_weights", "raae_reset_optimizers", "raae_step_debug",
           "raae_validate", "raae_train_epochs", "raae_launch_count", "raae_set_profile_buffer",
           "raae_train_phase", "raae_apply_adam", "raae_validate_epoch", "raae_evaluate_trials",
           "raae_debug_plateau",
           "raae_peer_alloc", "raae_peer_connect", "raae_peer_grad_ptr", "raae_apply_adam_peer", "raae_peer_status",
           "raae_peer_free")
MAX_PEERS = 8
IPC_HANDLE_BYTES = 64

_lib = None


class RaaeError(RuntimeError):
    pass


def load():
    """Loads the shared library once; raises if it has not been built (no CPU / eager fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RaaeError(f"{LIB_PATH} is missing: build it with __graft_entry__.build(); "
                        "rankaae_b200 has no fallback compute path")
    lib = C.CDLL(LIB_PATH)
    lib.raae_last_error.restype = C.c_char_p
    lib.raae_version.restype = C.c_int
    lib.raae_query_layout.argtypes = [C.POINTER(Config), C.POINTER(Layout)]
    lib.raae_create.argtypes = [C.POINTER(Config), C.c_int, C.POINTER(_p)]
    lib.raae_destroy.argtypes = [_p]
    lib.raae_max_clusters.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.raae_bind_state.argtypes = [_p, _p, _p, _p]
    lib.raae_bind_dataset.argtypes = [_p, _p, _p, C.c_int, _p, _p, C.c_int]
    lib.raae_bind_shapiro_weights.argtypes = [_p, _p, C.c_int]
    lib.raae_reset_optimizers.argtypes = [_p, _p]
    lib.raae_step_debug.argtypes = [_p, C.c_int, C.POINTER(DebugIO), _p]
    lib.raae_validate.argtypes = [_p, C.c_int, C.POINTER(ValIO), _p]
    lib.raae_train_epochs.argtypes = [_p, C.c_int, C.c_int, _p, _p, _p, _p]
    lib.raae_launch_count.argtypes = [_p]
    lib.raae_launch_count.restype = C.c_int64
    lib.raae_set_profile_buffer.argtypes = [_p, _p]
    lib.raae_train_phase.argtypes = [_p, C.c_int, C.c_int, C.c_int, _p, C.POINTER(_p), _p]
    lib.raae_apply_adam.argtypes = [_p, C.c_int, _p, _p]
    lib.raae_validate_epoch.argtypes = [_p, C.c_int, _p, _p, _p]
    lib.raae_evaluate_trials.argtypes = [_p, C.c_int, _p, _p, _p, _p, _p]
    lib.raae_debug_plateau.argtypes = [_p, C.c_int, _p, C.c_int, _p, _p]
    lib.raae_peer_alloc.argtypes = [_p, C.c_int, C.c_int, C.c_char_p]
    lib.raae_peer_connect.argtypes = [_p, C.c_char_p]
    lib.raae_peer_grad_ptr.argtypes = [_p, C.c_int, C.POINTER(_p)]
    lib.raae_apply_adam_peer.argtypes = [_p, C.c_int, _p]
    lib.raae_peer_status.argtypes = [_p, C.POINTER(C.c_uint)]
    lib.raae_peer_free.argtypes = [_p]
    _lib = lib
    return lib


def max_clusters(ctas_per_trial, device=0):
    """Co-resident clusters of `ctas_per_trial` CTAs of the train kernel on `device` (148 for one CTA per trial on a B200)."""
    if ctas_per_trial == 1:
        import torch
        return torch.cuda.get_device_properties(device).multi_processor_count
    n = C.c_int(0)
    check(load().raae_max_clusters(int(ctas_per_trial), int(device), C.byref(n)))
    return n.value


def check(rc):
    if rc != 0:
        raise RaaeError(f"rankaae_b200 error {rc}: {load().raae_last_error().decode()}")


def query_layout(cfg: Config) -> Layout:
    lay = Layout()
    check(load().raae_query_layout(C.byref(cfg), C.byref(lay)))
    return lay

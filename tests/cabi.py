"""ctypes mirror of the reference C ABI (struct layouts + prototypes).

The same mirror binds BOTH libraries, because the product keeps the reference's struct layouts
and symbol names:
  * ``oracle/_ref/libppo_ref.so``  — the unmodified reference, compiled by oracle/Makefile
  * ``ppo.c_b200/libppo_b200.so``  — the product (CUDA), see include/ppo_b200.h

Layouts follow /root/reference/include: neural_network.h:18-53 (Layer, NeuralNetwork),
policy.h:13-24, trajectory_buffer.h:15-62, adam.h:10-21, ppo.h:15-28, env.h:7-15,
activation_function.h:10-13.
"""
import ctypes as C
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libppo_ref.so")
ORACLE_SO = os.path.join(ROOT, "oracle", "libppo_oracle.so")
B200_SO = os.path.join(ROOT, "ppo.c_b200", "libppo_b200.so")

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
c_bool_p = C.POINTER(C.c_bool)


class ActivationFunction(C.Structure):
    _fields_ = [("activation", C.c_void_p), ("activation_derivative", C.c_void_p)]


class Layer(C.Structure):
    _fields_ = [
        ("weights", c_float_p), ("biases", c_float_p), ("grad_weights", c_float_p),
        ("grad_biases", c_float_p), ("input", c_float_p),
        ("d_weights", c_float_p), ("d_biases", c_float_p), ("d_grad_weights", c_float_p),
        ("d_grad_biases", c_float_p), ("d_input", c_float_p), ("d_grad_x", c_float_p),
        ("activation_function", C.POINTER(ActivationFunction)),
        ("d_activation_function", C.POINTER(ActivationFunction)),
        ("input_size", C.c_int), ("output_size", C.c_int),
    ]


class NeuralNetwork(C.Structure):
    _fields_ = [
        ("layers", C.POINTER(Layer)), ("num_layers", C.c_int), ("output_size", C.c_int),
        ("cache_m_forward", C.c_int), ("cache_m_backward", C.c_int),
        ("output", c_float_p), ("d_output", c_float_p),
        ("activation_functions", C.POINTER(C.c_char_p)),
        ("cublas_handle", C.c_void_p),
    ]


class GaussianPolicy(C.Structure):
    _fields_ = [
        ("mu", C.POINTER(NeuralNetwork)), ("log_std", c_float_p), ("log_std_grad", c_float_p),
        ("d_log_std", c_float_p), ("d_log_std_grad", c_float_p),
        ("state_size", C.c_int), ("action_size", C.c_int),
        ("input_action", c_float_p), ("d_input_action", c_float_p),
    ]


class TrajectoryBuffer(C.Structure):
    pass


_ACC_F = C.CFUNCTYPE(c_float_p, C.POINTER(TrajectoryBuffer), C.c_int)
_ACC_B = C.CFUNCTYPE(c_bool_p, C.POINTER(TrajectoryBuffer), C.c_int)
_F9 = ["state_p", "action_p", "next_state_p", "reward_p", "logprob_p", "advantage_p", "adv_target_p"]
TrajectoryBuffer._fields_ = (
    [(n, c_float_p) for n in _F9] + [("terminated_p", c_bool_p), ("truncated_p", c_bool_p)]
    + [("h_" + n, c_float_p) for n in _F9] + [("h_terminated_p", c_bool_p), ("h_truncated_p", c_bool_p)]
    + [("d_" + n, c_float_p) for n in _F9] + [("d_terminated_p", c_bool_p), ("d_truncated_p", c_bool_p)]
    + [("random_idx", c_int_p), ("state_size", C.c_int), ("action_size", C.c_int),
       ("capacity", C.c_int), ("idx", C.c_int), ("full", C.c_bool)]
    + [(n, _ACC_F) for n in ["state", "action", "next_state", "reward", "logprob", "advantage", "adv_target"]]
    + [("terminated", _ACC_B), ("truncated", _ACC_B)]
)


class Adam(C.Structure):
    _fields_ = [
        ("weights", C.POINTER(c_float_p)), ("grad_weights", C.POINTER(c_float_p)),
        ("lengths", c_int_p), ("m", c_float_p), ("v", c_float_p),
        ("beta1", C.c_float), ("beta2", C.c_float), ("time_step", C.c_int),
        ("size", C.c_int), ("num_layers", C.c_int),
    ]


RESET_FN = C.CFUNCTYPE(None, c_float_p)
STEP_FN = C.CFUNCTYPE(None, c_float_p, c_float_p, c_float_p, c_bool_p, c_bool_p, C.c_int)
FREE_FN = C.CFUNCTYPE(None)


class Env(C.Structure):
    _fields_ = [
        ("free_env", FREE_FN), ("reset_env", RESET_FN), ("step_env", STEP_FN),
        ("state_size", C.c_int), ("action_size", C.c_int), ("horizon", C.c_int),
        ("gamma", C.c_float),
    ]


class PPO(C.Structure):
    _fields_ = [
        ("buffer", C.POINTER(TrajectoryBuffer)), ("policy", C.POINTER(GaussianPolicy)),
        ("V", C.POINTER(NeuralNetwork)), ("adam_policy", C.POINTER(Adam)),
        ("adam_V", C.POINTER(Adam)), ("adam_entropy", C.POINTER(Adam)),
        ("lambda_", C.c_float), ("epsilon", C.c_float), ("ent_coeff", C.c_float),
        ("lr_policy", C.c_float), ("lr_V", C.c_float), ("use_cuda", C.c_bool),
    ]


NNp, GPp, TBp, ADp, PPOp, ENVp = (C.POINTER(t) for t in
                                  (NeuralNetwork, GaussianPolicy, TrajectoryBuffer, Adam, PPO, Env))

# name -> (restype, argtypes): every symbol of the reference's header set.
REFERENCE_API = {
    # ppo.h:30-47
    "create_ppo": (PPOp, [C.POINTER(C.c_char_p), c_int_p, C.c_int, C.c_int, C.c_float, C.c_float,
                          C.c_float, C.c_float, C.c_float, C.c_float, C.c_bool]),
    "free_ppo": (None, [PPOp]),
    "collect_trajectories": (None, [TBp, ENVp, GPp, C.c_int]),
    "compute_gae": (None, [NNp, TBp, C.c_float, C.c_float]),
    "policy_loss_and_grad": (C.c_float, [c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                         C.c_float, C.c_float, C.c_float, C.c_int]),
    "compute_gae_cuda": (None, [NNp, TBp, C.c_float, C.c_float, C.c_int]),
    "policy_loss_and_grad_cuda": (C.c_float, [c_float_p, c_float_p, c_float_p, c_float_p, c_float_p,
                                              C.c_float, C.c_float, C.c_float, C.c_int]),
    "train_ppo_epoch": (None, [PPOp, ENVp, C.c_int, C.c_int, C.c_int, C.c_int]),
    "eval_ppo": (None, [PPOp, ENVp, C.c_int]),
    "save_ppo": (None, [PPOp, C.c_char_p]),
    "load_ppo": (PPOp, [C.c_char_p, C.c_bool]),
    # policy.h:26-41
    "create_gaussian_policy": (GPp, [c_int_p, C.POINTER(C.c_char_p), C.c_int, C.c_float]),
    "free_gaussian_policy": (None, [GPp]),
    "sample_action": (None, [GPp, c_float_p, c_float_p, c_float_p, C.c_int]),
    "compute_log_prob": (None, [GPp, c_float_p, c_float_p, c_float_p, C.c_int]),
    "log_prob_backwards": (None, [GPp, c_float_p, c_float_p, c_float_p, C.c_int]),
    "compute_log_prob_cuda": (None, [GPp, c_float_p, c_float_p, c_float_p, C.c_int]),
    "log_prob_backwards_cuda": (None, [GPp, c_float_p, c_float_p, c_float_p, C.c_int]),
    "compute_entropy_cuda": (C.c_float, [GPp]),
    "compute_entropy": (C.c_float, [GPp]),
    "policy_to_host": (None, [GPp]),
    "save_policy": (None, [GPp, C.c_void_p]),
    "load_policy": (GPp, [C.c_void_p, C.c_int, C.c_int]),
    # neural_network.h:60-72
    "create_neural_network": (NNp, [c_int_p, C.POINTER(C.c_char_p), C.c_int]),
    "forward_propagation": (None, [NNp, c_float_p, C.c_int]),
    "free_neural_network": (None, [NNp]),
    "backward_propagation": (None, [NNp, c_float_p, C.c_int]),
    "forward_propagation_cuda": (None, [NNp, c_float_p, C.c_int]),
    "backward_propagation_cuda": (None, [NNp, c_float_p, C.c_int]),
    "nn_write_weights_to_device": (None, [NNp]),
    "nn_write_weights_to_host": (None, [NNp]),
    "save_neural_network": (None, [NNp, C.c_void_p]),
    "load_neural_network": (NNp, [C.c_void_p]),
    # trajectory_buffer.h:66-79
    "create_trajectory_buffer": (TBp, [C.c_int, C.c_int, C.c_int]),
    "free_trajectory_buffer": (None, [TBp, C.c_bool]),
    "shuffle_buffer": (None, [TBp]),
    "get_batch": (None, [TBp, C.c_int, C.c_int] + [c_float_p] * 5),
    "shuffle_buffer_cuda": (None, [TBp]),
    "get_batch_cuda": (None, [TBp, C.c_int, C.c_int] + [c_float_p] * 5),
    "reset_buffer": (None, [TBp]),
    "buffer_to_device": (None, [TBp]),
    "buffer_to_host": (None, [TBp]),
    # adam.h:24-38
    "create_adam": (ADp, [C.POINTER(c_float_p), C.POINTER(c_float_p), c_int_p, C.c_int, C.c_int,
                          C.c_float, C.c_float]),
    "create_adam_from_nn": (ADp, [NNp, C.c_float, C.c_float]),
    "free_adam": (None, [ADp]),
    "adam_update": (None, [ADp, C.c_float]),
    "create_adam_cuda": (ADp, [C.POINTER(c_float_p), C.POINTER(c_float_p), c_int_p, C.c_int, C.c_int,
                               C.c_float, C.c_float]),
    "create_adam_from_nn_cuda": (ADp, [NNp, C.c_float, C.c_float]),
    "free_adam_cuda": (None, [ADp]),
    "adam_update_cuda": (None, [ADp, C.c_float]),
    "save_adam": (None, [ADp, C.c_void_p, C.c_bool]),
    "load_adam": (ADp, [C.c_void_p, C.POINTER(c_float_p), C.POINTER(c_float_p), c_int_p, C.c_bool]),
    "load_adam_from_nn": (ADp, [C.c_void_p, NNp, C.c_bool]),
    # loss.h:10-14
    "mean_squared_error": (C.c_float, [c_float_p, c_float_p, C.c_int, C.c_int]),
    "mean_squared_error_derivative": (None, [c_float_p, c_float_p, c_float_p, C.c_int, C.c_int]),
    "mean_squared_error_cuda": (C.c_float, [c_float_p, c_float_p, C.c_int, C.c_int]),
    "mean_squared_error_derivative_cuda": (None, [c_float_p, c_float_p, c_float_p, C.c_int, C.c_int]),
    # mat_mul.h:16-20
    "mat_mul": (None, [c_float_p] * 4 + [C.c_int] * 3),
    "mat_mul_backwards": (None, [c_float_p] * 5 + [C.c_int] * 3),
    "mat_mul_cuda": (None, [C.c_void_p] + [c_float_p] * 4 + [C.c_int] * 3),
    "mat_mul_backwards_cuda": (None, [C.c_void_p] + [c_float_p] * 5 + [C.c_int] * 3),
    # activation_function.h:15-22
    "ReLU": (None, [c_float_p, C.c_int, C.c_int]),
    "ReLU_derivative": (None, [c_float_p, c_float_p, C.c_int, C.c_int]),
    "ReLU_cuda": (None, [c_float_p, C.c_int, C.c_int]),
    "ReLU_derivative_cuda": (None, [c_float_p, c_float_p, C.c_int, C.c_int]),
    "build_activation_function": (C.POINTER(ActivationFunction), [C.c_char_p]),
    "build_activation_function_cuda": (C.POINTER(ActivationFunction), [C.c_char_p]),
    # env.h:18
    "create_simple_env": (ENVp, [C.c_int, C.c_int]),
}


def bind(lib, table, strict=True):
    missing = []
    for name, (res, args) in table.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if strict and missing:
        raise RuntimeError("missing symbols: %s" % missing)
    return missing


_libc = C.CDLL(None)
_libc.srand.argtypes = [C.c_uint]
_libc.rand.restype = C.c_int
_libc.fopen.restype = C.c_void_p
_libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
_libc.fclose.argtypes = [C.c_void_p]


def srand(seed):
    _libc.srand(seed)


def rand():
    return _libc.rand()


def fopen(path, mode):
    f = _libc.fopen(path.encode(), mode.encode())
    if not f:
        raise OSError("fopen failed: %s" % path)
    return f


def fclose(f):
    _libc.fclose(f)


def unlimit_stack():
    """The reference keeps B-sized VLAs on the stack (ppo.cu:329-338,374-383)."""
    import resource
    soft, hard = resource.getrlimit(resource.RLIMIT_STACK)
    try:
        resource.setrlimit(resource.RLIMIT_STACK, (hard, hard))
    except (ValueError, OSError):
        pass


def load_ref():
    """The unmodified reference (plain-C twins only are usable without a GPU)."""
    lib = C.CDLL(REF_SO, mode=C.RTLD_LOCAL)
    bind(lib, REFERENCE_API)
    return lib


REF_BLAS_SO = os.path.join(ROOT, "oracle", "_ref", "libppo_ref_blas.so")
_OPENBLAS_DIR = os.path.join(os.path.dirname(os.path.dirname(C.__file__)), "site-packages", "opencv_python_headless.libs")


def load_openblas():
    """The OpenBLAS 0.3.15 bundled with the image's opencv wheel (unprefixed cblas_* symbols), RTLD_GLOBAL so that
    oracle/_ref/libppo_ref_blas.so binds its undefined cblas_sgemm / cblas_sgemv to it.  Returns the handle or None."""
    import glob
    import sys
    dirs = [_OPENBLAS_DIR] + [os.path.join(p, "opencv_python_headless.libs") for p in sys.path if p.endswith("site-packages")]
    for d in dirs:
        libs = sorted(glob.glob(os.path.join(d, "libopenblas*.so")))
        if not libs:
            continue
        try:
            for pat in ("libquadmath*", "libgfortran*"):
                for f in sorted(glob.glob(os.path.join(d, pat))):
                    C.CDLL(f, mode=C.RTLD_GLOBAL)
            h = C.CDLL(libs[0], mode=C.RTLD_GLOBAL)
            h.openblas_set_num_threads.argtypes = [C.c_int]
            h.openblas_set_num_threads(1)          # src/main.c:18
            return h
        except OSError:
            continue
    return None


def load_ref_blas():
    """The unmodified reference linked against a real BLAS (the way its own Makefile:5 links it), or None."""
    if not os.path.exists(REF_BLAS_SO) or load_openblas() is None:
        return None
    lib = C.CDLL(REF_BLAS_SO, mode=C.RTLD_LOCAL)
    bind(lib, REFERENCE_API)
    return lib


def oracle_pendulum_env(oracle_lib):
    """An `Env` (include/env.h:7-15) whose hooks are the C Pendulum of oracle/ppo_oracle.c: no Python in the loop."""
    env = Env()
    env.free_env = C.cast(oracle_lib.orc_pendulum_hook_free, FREE_FN)
    env.reset_env = C.cast(oracle_lib.orc_pendulum_hook_reset, RESET_FN)
    env.step_env = C.cast(oracle_lib.orc_pendulum_hook_step, STEP_FN)
    env.state_size, env.action_size, env.horizon, env.gamma = 3, 1, 200, 0.99
    return env


def fptr(arr):
    return arr.ctypes.data_as(c_float_p)


def iptr(arr):
    return arr.ctypes.data_as(c_int_p)


def bptr(arr):
    return arr.ctypes.data_as(c_bool_p)


def cstr_array(strs):
    arr = (C.c_char_p * len(strs))()
    arr[:] = [s.encode() for s in strs]
    return arr


def int_array(vals):
    arr = (C.c_int * len(vals))()
    arr[:] = list(vals)
    return arr

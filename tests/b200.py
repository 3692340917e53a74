"""Test-side access to the product library: imports the dotted package by path, binds the reference
API (tests/cabi.py) plus the extension API onto ONE ctypes handle."""
import ctypes as C
import importlib.util
import os
import sys

import numpy as np

import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def package():
    if "ppo_c_b200" in sys.modules:
        return sys.modules["ppo_c_b200"]
    spec = importlib.util.spec_from_file_location("ppo_c_b200", os.path.join(ROOT, "ppo.c_b200", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ppo_c_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


_lib = None


def lib():
    global _lib
    if _lib is None:
        pk = package()
        _lib = pk.load_library()
        cabi.bind(_lib, cabi.REFERENCE_API)
        # typed overrides used by the tests (structs instead of void*)
        _lib.create_pendulum_env.restype = cabi.ENVp
        _lib.create_pendulum_env_cuda.restype = cabi.ENVp
        _lib.create_gym_env.restype = cabi.ENVp
        for name in ["ppo_b200_update", "ppo_b200_update_device"]:
            getattr(_lib, name).argtypes = [cabi.PPOp, C.c_float, C.c_int, C.c_int, C.c_int]
        for name in ["ppo_b200_buffer_upload", "ppo_b200_sync_host"]:
            getattr(_lib, name).argtypes = [cabi.PPOp]
        _lib.ppo_b200_train_iterations.argtypes = [cabi.PPOp, cabi.ENVp, C.c_int, C.c_int, C.c_int, C.c_int]
        _lib.ppo_b200_set_permutation_mode.argtypes = [cabi.PPOp, C.c_int, C.c_ulonglong]
        for name in ["ppo_b200_last_mean_return", "ppo_b200_last_value_loss", "ppo_b200_last_policy_loss"]:
            getattr(_lib, name).argtypes = [cabi.PPOp]
        _lib.ppo_b200_env_is_device.argtypes = [cabi.ENVp]
        _lib.ppo_b200_env_num_envs.argtypes = [cabi.ENVp]
    return _lib


def dev(arr):
    return package().DeviceArray.from_numpy(lib(), np.ascontiguousarray(arr))


def dev_empty(shape, dtype=np.float32):
    return package().DeviceArray(lib(), shape, dtype)


def has_gpu():
    try:
        return lib().ppo_b200_device_count() > 0
    except Exception:
        return False


# ---- struct helpers shared by the GPU tests ------------------------------------------------------
def nn_set_params(l, nn, flat):
    """Write a flat W0,b0,W1,b1.. vector into the HOST arrays and push them to the device."""
    n = nn.contents
    o = 0
    for i in range(n.num_layers - 1):
        L = n.layers[i]
        k = L.input_size * L.output_size
        np.ctypeslib.as_array(L.weights, shape=(k,))[:] = flat[o:o + k]
        o += k
        np.ctypeslib.as_array(L.biases, shape=(L.output_size,))[:] = flat[o:o + L.output_size]
        o += L.output_size
    l.nn_write_weights_to_device(nn)


def nn_get_params(l, nn, sync=True):
    if sync:
        l.nn_write_weights_to_host(nn)
    n = nn.contents
    out = []
    for i in range(n.num_layers - 1):
        L = n.layers[i]
        out.append(np.ctypeslib.as_array(L.weights, shape=(L.input_size * L.output_size,)).copy())
        out.append(np.ctypeslib.as_array(L.biases, shape=(L.output_size,)).copy())
    return np.concatenate(out)


def nn_get_device_grads(l, nn):
    n = nn.contents
    out = []
    for i in range(n.num_layers - 1):
        L = n.layers[i]
        for ptr, cnt in ((L.d_grad_weights, L.input_size * L.output_size), (L.d_grad_biases, L.output_size)):
            a = np.empty(cnt, np.float32)
            l.ppo_b200_d2h(a.ctypes.data, C.cast(ptr, C.c_void_p), a.nbytes)
            out.append(a)
    return np.concatenate(out)


def d2h(l, ptr, shape, dtype=np.float32):
    a = np.empty(shape, dtype)
    l.ppo_b200_d2h(a.ctypes.data, C.cast(ptr, C.c_void_p), a.nbytes)
    return a

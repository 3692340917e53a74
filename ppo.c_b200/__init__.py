"""ppo.c_b200 — host-side loader for the B200-native PPO training path.

The product is the C-ABI shared library ``libppo_b200.so`` next to this file (sources in ``csrc/``,
boundary in ``include/ppo_b200.h``).  This module only builds it, loads it with ctypes and attaches
prototypes; it contains no arithmetic and no fallback: if the library is missing or no CUDA device is
present, the calls fail loudly.

The directory name contains a dot (the upstream project is called ``ppo.c``), so import it by path:

    import importlib.util, sys
    spec = importlib.util.spec_from_file_location("ppo_c_b200", "<repo>/ppo.c_b200/__init__.py")
    mod = importlib.util.module_from_spec(spec); sys.modules["ppo_c_b200"] = mod; spec.loader.exec_module(mod)
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SO_PATH = os.path.join(HERE, "libppo_b200.so")

c_float_p = C.POINTER(C.c_float)
c_int_p = C.POINTER(C.c_int)
c_bool_p = C.POINTER(C.c_bool)
c_double_p = C.POINTER(C.c_double)
vp = C.c_void_p

# Additive entry points of include/ppo_b200.h section (2): name -> (restype, argtypes).
# Device pointers are passed as void* (integers).
EXTENSION_API = {
    "create_pendulum_env": (vp, [C.c_int, C.c_int]),
    "create_pendulum_env_cuda": (vp, [C.c_int, C.c_int]),
    "create_gym_env": (vp, [C.c_int, C.c_int]),
    "openblas_set_num_threads": (None, [C.c_int]),
    "ppo_b200_env_is_device": (C.c_int, [vp]),
    "ppo_b200_env_num_envs": (C.c_int, [vp]),
    "ppo_b200_device_count": (C.c_int, []),
    "ppo_b200_set_device": (None, [C.c_int]),
    "ppo_b200_set_stream": (None, [vp]),
    "ppo_b200_malloc": (vp, [C.c_size_t]),
    "ppo_b200_free": (None, [vp]),
    "ppo_b200_malloc_host": (vp, [C.c_size_t]),
    "ppo_b200_free_host": (None, [vp]),
    "ppo_b200_h2d": (None, [vp, vp, C.c_size_t]),
    "ppo_b200_d2h": (None, [vp, vp, C.c_size_t]),
    "ppo_b200_memset": (None, [vp, C.c_int, C.c_size_t]),
    "ppo_b200_sync": (None, []),
    "ppo_b200_launch_count": (C.c_ulonglong, []),
    "ppo_b200_version": (C.c_char_p, []),
    "ppo_b200_measure_fp32_peak": (C.c_double, [C.c_int]),
    "ppo_b200_profile_begin": (None, []),
    "ppo_b200_profile_end": (C.c_int, [C.c_char_p, C.c_int]),
    "ppo_b200_debug_phase_stamps": (None, [vp, C.c_int]),
    "ppo_b200_gae": (None, [vp, vp, vp, vp, vp, C.c_int, C.c_float, C.c_float, vp, vp, C.c_int, vp]),
    "ppo_b200_adam_flat": (None, [vp, vp, vp, vp, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]),
    "ppo_b200_gather": (None, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [vp] * 10),
    "ppo_b200_packed_row_floats": (C.c_int, [C.c_int, C.c_int]),
    "ppo_b200_pack_rows": (None, [vp, C.c_longlong, C.c_int, C.c_int] + [vp] * 5),
    "ppo_b200_gather_packed": (None, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int] + [vp] * 6),
    "ppo_b200_permutation": (None, [vp, C.c_int, C.c_ulonglong, C.c_ulonglong]),
    "ppo_b200_pendulum_step": (None, [vp, vp, vp, vp, vp, C.c_int]),
    "ppo_b200_update": (None, [vp, C.c_float, C.c_int, C.c_int, C.c_int]),
    "ppo_b200_update_device": (None, [vp, C.c_float, C.c_int, C.c_int, C.c_int]),
    "ppo_b200_buffer_upload": (None, [vp]),
    "ppo_b200_sync_host": (None, [vp]),
    "ppo_b200_train_iterations": (None, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ppo_b200_set_permutation_mode": (None, [vp, C.c_int, C.c_ulonglong]),
    "ppo_b200_set_obs_norm": (None, [vp, C.c_int]),
    "ppo_b200_get_obs_norm": (None, [vp, vp, vp, vp]),
    "ppo_b200_set_kernel_path": (None, [C.c_int]),
    "ppo_b200_set_matmul_precision": (None, [C.c_int]),
    "ppo_b200_tc_linear": (None, [C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ppo_b200_tc_linear_x3": (None, [C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ppo_b200_tc_linear_bf16": (None, [C.c_int, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "ppo_b200_last_mean_return": (C.c_float, [vp]),
    "ppo_b200_last_eval": (None, [vp, c_float_p, c_float_p, c_int_p]),
    "ppo_b200_last_value_loss": (C.c_float, [vp]),
    "ppo_b200_last_policy_loss": (C.c_float, [vp]),
    "ppo_b200_dist_unique_id": (None, [C.c_char_p]),
    "ppo_b200_dist_init": (None, [C.c_char_p, C.c_int, C.c_int]),
    "ppo_b200_dist_finalize": (None, []),
    "ppo_b200_dist_rank": (C.c_int, []),
    "ppo_b200_dist_world": (C.c_int, []),
    "ppo_b200_dist_set_shard_mode": (None, [C.c_int]),
    "Tanh_cuda": (None, [vp, C.c_int, C.c_int]),
    "Tanh_derivative_cuda": (None, [vp, vp, C.c_int, C.c_int]),
}


def build(verbose=False):
    """Compile every CUDA source for sm_100a into libppo_b200.so (in-tree; nvcc cross-compiles)."""
    cmd = ["make", "-C", HERE, "-j", str(os.cpu_count() or 4)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libppo_b200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout[-2000:])
    return SO_PATH


_lib = None


def load_library():
    """dlopen libppo_b200.so (RTLD_LOCAL) and attach the extension prototypes.  No fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError("%s is missing: run __graft_entry__.build() (there is no CPU fallback)" % SO_PATH)
    lib = C.CDLL(SO_PATH, mode=C.RTLD_LOCAL)
    for name, (res, args) in EXTENSION_API.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class DeviceArray:
    """A device allocation made through the library's own allocator (numpy in / numpy out)."""

    def __init__(self, lib, shape, dtype):
        self.lib, self.shape, self.dtype = lib, tuple(np.atleast_1d(shape)), np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        self.ptr = lib.ppo_b200_malloc(max(self.nbytes, 4))

    @classmethod
    def from_numpy(cls, lib, arr):
        arr = np.ascontiguousarray(arr)
        d = cls(lib, arr.shape, arr.dtype)
        if d.nbytes:
            lib.ppo_b200_h2d(d.ptr, arr.ctypes.data, d.nbytes)
        return d

    def numpy(self):
        out = np.empty(self.shape, self.dtype)
        if self.nbytes:
            self.lib.ppo_b200_d2h(out.ctypes.data, self.ptr, self.nbytes)
        return out

    def fp(self):
        return C.cast(self.ptr, c_float_p)

    def free(self):
        if self.ptr:
            self.lib.ppo_b200_free(self.ptr)
            self.ptr = None


# ---- host-side data-parallel helpers (pure index arithmetic; tested under gloo on CPU) -----------
def shard_rows(batch_size, rank, world, mode=0):
    """Rows of a global minibatch owned by `rank` (mirrors update_device in csrc/ppo.cu)."""
    if world <= 1:
        return 0, batch_size, batch_size
    if mode == 0:
        if batch_size % world:
            raise ValueError("batch_size %d not divisible by world %d" % (batch_size, world))
        local = batch_size // world
        return rank * local, local, batch_size
    return 0, batch_size, batch_size * world


def welford_merge(triples):
    """Ordered merge of (mean, M2, n) triples — the arithmetic of gae_merge_ranks_kernel (float64)."""
    mean, m2, cnt = 0.0, 0.0, 0.0
    for mb, m2b, nb in triples:
        if nb > 0:
            delta = mb - mean
            nn = cnt + nb
            mean += delta * nb / nn
            m2 += m2b + delta * delta * cnt * nb / nn
            cnt = nn
    return mean, m2, cnt


def dist_init_from_torch(lib):
    """Create the library's NCCL communicator using torch.distributed only to broadcast the id."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = C.create_string_buffer(128)
    if rank == 0:
        lib.ppo_b200_dist_unique_id(buf)
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().numpy().tobytes())
    lib.ppo_b200_dist_init(C.create_string_buffer(raw, 128), rank, world)
    return rank, world

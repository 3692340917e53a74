"""CPU-only checks of the drop-in boundary: the product library loads without a GPU, exports every
symbol include/ppo_b200.h declares, keeps the reference's struct layouts, and the reference's own
caller (src/main.c) compiles and links against it unmodified."""
import ctypes as C
import os
import re
import subprocess

import pytest

import b200
import cabi

ROOT = cabi.ROOT


def declared_functions():
    src = open(os.path.join(ROOT, "include", "ppo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set()
    for m in re.finditer(r"^[A-Za-z_][\w \*]*?\b(\w+)\s*\([^;{]*\)\s*;", src, flags=re.M):
        name = m.group(1)
        if "(*" in m.group(0).split(name)[0]:
            continue
        names.add(name)
    return names


def test_library_loads_and_exports_every_declared_symbol():
    b200.package().build()
    lib = b200.lib()      # binds the reference API + extension API (raises on a missing symbol)
    names = declared_functions()
    assert len(names) > 90
    missing = [n for n in sorted(names) if not hasattr(lib, n)]
    assert not missing, missing
    for n in cabi.REFERENCE_API:
        assert n in names, n
    assert b"sm_100a" in lib.ppo_b200_version()
    # ... and nothing else: every unmangled (C ABI) function the library exports is declared in the boundary header
    out = subprocess.check_output(["nm", "-D", "--defined-only", os.path.join(ROOT, "ppo.c_b200", "libppo_b200.so")], text=True)
    exported = {ln.split()[2] for ln in out.splitlines() if len(ln.split()) == 3 and ln.split()[1] == "T" and not ln.split()[2].startswith("_Z")}
    assert exported - names == set(), sorted(exported - names)


def test_packed_row_width_is_whole_sectors():
    """Host-side layout rule of the row-packed gather mirror (no device needed): a row holds S + A + 3 floats rounded up to
    whole 32-byte sectors."""
    lib = b200.lib()
    for S, A in [(3, 1), (17, 6), (1, 1), (24, 4), (100, 25), (5, 0)]:
        pw = lib.ppo_b200_packed_row_floats(S, A)
        assert pw % 8 == 0 and S + A + 3 <= pw < S + A + 3 + 8


def test_struct_layouts_match_reference_abi():
    """sizeof/offsetof of the public structs, as compiled from include/ppo_b200.h by gcc, must equal
    the ctypes mirror of the REFERENCE headers (tests/cabi.py)."""
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "ppo.h"
#define P(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void) {
  printf("Layer %zu\nNeuralNetwork %zu\nGaussianPolicy %zu\nTrajectoryBuffer %zu\nAdam %zu\nEnv %zu\nPPO %zu\n",
         sizeof(Layer), sizeof(NeuralNetwork), sizeof(GaussianPolicy), sizeof(TrajectoryBuffer), sizeof(Adam), sizeof(Env), sizeof(PPO));
  P(Layer, d_grad_x); P(Layer, input_size); P(NeuralNetwork, d_output); P(NeuralNetwork, cublas_handle);
  P(GaussianPolicy, d_input_action); P(TrajectoryBuffer, random_idx); P(TrajectoryBuffer, full);
  P(TrajectoryBuffer, truncated); P(Adam, time_step); P(Env, gamma); P(PPO, use_cuda); P(PPO, lambda);
  return 0; }
'''
    tmp = os.path.join(ROOT, "build", "abi_probe")
    os.makedirs(tmp, exist_ok=True)
    open(os.path.join(tmp, "probe.c"), "w").write(prog)
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(tmp, "probe.c"), "-o", os.path.join(tmp, "probe")])
    out = dict(line.rsplit(" ", 1) for line in subprocess.check_output([os.path.join(tmp, "probe")], text=True).splitlines())
    sizes = {"Layer": cabi.Layer, "NeuralNetwork": cabi.NeuralNetwork, "GaussianPolicy": cabi.GaussianPolicy,
             "TrajectoryBuffer": cabi.TrajectoryBuffer, "Adam": cabi.Adam, "Env": cabi.Env, "PPO": cabi.PPO}
    for name, t in sizes.items():
        assert int(out[name]) == C.sizeof(t), name
    for key, (t, f) in {"Layer.d_grad_x": (cabi.Layer, "d_grad_x"), "Layer.input_size": (cabi.Layer, "input_size"),
                        "NeuralNetwork.d_output": (cabi.NeuralNetwork, "d_output"),
                        "NeuralNetwork.cublas_handle": (cabi.NeuralNetwork, "cublas_handle"),
                        "GaussianPolicy.d_input_action": (cabi.GaussianPolicy, "d_input_action"),
                        "TrajectoryBuffer.random_idx": (cabi.TrajectoryBuffer, "random_idx"),
                        "TrajectoryBuffer.full": (cabi.TrajectoryBuffer, "full"),
                        "TrajectoryBuffer.truncated": (cabi.TrajectoryBuffer, "truncated"),
                        "Adam.time_step": (cabi.Adam, "time_step"), "Env.gamma": (cabi.Env, "gamma"),
                        "PPO.use_cuda": (cabi.PPO, "use_cuda"), "PPO.lambda": (cabi.PPO, "lambda_")}.items():
        assert int(out[key]) == getattr(t, f).offset, key


@pytest.mark.skipif(not os.path.exists("/root/reference/src/main.c"), reason="reference not mounted")
def test_reference_main_compiles_and_links_unmodified():
    """The reference's own caller, src/main.c, compiled where it lies against OUR headers and linked
    against OUR library (it is not run here: it needs a GPU)."""
    b200.package().build()
    out = os.path.join(ROOT, "build", "abi_probe", "ref_main")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["gcc", "-std=gnu11", "-Wno-implicit-function-declaration", "-Wno-discarded-qualifiers",
                           "-I", os.path.join(ROOT, "include"), "/root/reference/src/main.c", "-o", out,
                           "-L", os.path.join(ROOT, "ppo.c_b200"), "-lppo_b200", "-lm",
                           "-Wl,-rpath," + os.path.join(ROOT, "ppo.c_b200")])
    assert os.path.exists(out)


def _build_example():
    b200.package().build()
    out = os.path.join(ROOT, "build", "abi_probe", "pendulum_b200")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["gcc", "-std=gnu11", "-O2", "-Wall", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "examples", "pendulum_b200.c"), "-o", out,
                           "-L", os.path.join(ROOT, "ppo.c_b200"), "-lppo_b200", "-lm",
                           "-Wl,-rpath," + os.path.join(ROOT, "ppo.c_b200")])
    return out


def test_c_example_compiles_against_the_boundary_headers():
    """A plain-C caller (gcc, no nvcc, no CUDA headers) builds against include/ and links the C-ABI library."""
    assert os.path.exists(_build_example())


@pytest.mark.gpu
def test_c_example_solves_pendulum():
    """The C caller trains Pendulum-v1 through train_ppo_epoch on the device env and exits 0 iff the mean return of
    some iteration exceeded -200 (the north star's "solved" threshold)."""
    res = subprocess.run([_build_example(), "40", "4096"], capture_output=True, text=True, timeout=300)
    print(res.stdout[-1500:])
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-1500:]


REF_MAIN = os.path.join(ROOT, "build", "abi_probe", "ref_main")


@pytest.mark.gpu
def test_reference_main_runs_unmodified_on_the_drop_in():
    """The reference's own src/main.c, compiled UNMODIFIED against include/ and linked against libppo_b200.so by
    test_reference_main_compiles_and_links_unmodified (in the build container, where /root/reference is mounted; the
    binary travels with the snapshot), trains Pendulum for its 10 epochs x 30000 steps at minibatch 64 through the
    reference call sequence (create_gym_env -> create_ppo -> eval_ppo / train_ppo_epoch -> save_ppo).  The mean
    episode return printed by eval_ppo ("R:") must improve clearly."""
    if not os.path.exists(REF_MAIN):
        pytest.skip("build/abi_probe/ref_main not prebuilt (needs /root/reference at build time)")
    res = subprocess.run([REF_MAIN, "64"], capture_output=True, text=True, timeout=600, cwd=os.path.join(ROOT, "build", "abi_probe"))
    print(res.stdout[-3000:])
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    rs = [float(m.group(1)) for m in re.finditer(r"R: (-?[0-9.]+)", res.stdout)]
    assert len(rs) >= 11, res.stdout[-2000:]
    assert max(rs[1:]) > rs[0] + 300, rs
    assert os.path.exists(os.path.join(ROOT, "build", "abi_probe", "ppo_model.bin"))


def test_boundary_header_compiles_as_c_and_cxx():
    """include/ppo_b200.h (and every thin compatibility header) is valid C11 and C++17 on its own, warnings as errors."""
    tmp = os.path.join(ROOT, "build", "abi_probe")
    os.makedirs(tmp, exist_ok=True)
    heads = ["ppo.h", "policy.h", "neural_network.h", "trajectory_buffer.h", "adam.h", "loss.h", "mat_mul.h",
             "activation_function.h", "env.h", "gym_env.h", "ppo_b200.h"]
    src = "".join('#include "%s"\n' % h for h in heads) + "int main(void) { return (int)sizeof(PPO) == 0; }\n"
    for name, cc, std in (("hdr.c", "gcc", "-std=c11"), ("hdr.cpp", "g++", "-std=c++17")):
        path = os.path.join(tmp, name)
        open(path, "w").write(src)
        subprocess.check_call([cc, std, "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), path])

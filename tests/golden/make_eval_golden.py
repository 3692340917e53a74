"""Mint tests/golden/eval_golden.json from the UNMODIFIED reference's eval_ppo (src/ppo.cu:560-583).

    python tests/golden/make_eval_golden.py          (build container only: needs oracle/_ref/libppo_ref.so)

For each case: srand(seed) -> create_simple_env + create_ppo(use_cuda=false) -> [train_ppo_epoch] -> eval_ppo.
eval_ppo only prints, so file descriptor 1 is redirected around the call and the printed line is the golden
("J: %f R: %f Episodes: %d").  The rand() value drawn right after is recorded too (stream position).
"""
import ctypes as C
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import cabi  # noqa: E402
import refdrive  # noqa: E402

# mu_bias: the output bias of the policy mean is overwritten after creation (a mean action near +1 walks the toy env to its
# goal, so rewards / terminations / discounting all show up in J and R; the untouched initial policy almost never scores).
CASES = [dict(seed=21, hidden=8, capacity=302, steps=302, train_epochs=0, mu_bias=None),
         dict(seed=22, hidden=16, capacity=600, steps=450, train_epochs=1, mu_bias=0.9),
         dict(seed=23, hidden=8, capacity=302, steps=31, train_epochs=0, mu_bias=0.5),
         dict(seed=24, hidden=8, capacity=2000, steps=2000, train_epochs=0, mu_bias=2.5)]


def capture_stdout(fn):
    libc = C.CDLL(None)
    sys.stdout.flush()
    libc.fflush(None)
    saved = os.dup(1)
    with tempfile.TemporaryFile() as tmp:
        os.dup2(tmp.fileno(), 1)
        try:
            fn()
            libc.fflush(None)
        finally:
            os.dup2(saved, 1)
            os.close(saved)
        tmp.seek(0)
        return tmp.read().decode()


def main():
    R = refdrive.Ref()
    out = []
    for c in CASES:
        cabi.srand(c["seed"])
        env = R.lib.create_simple_env(0, c["seed"])
        sizes = [1, c["hidden"], c["hidden"], 1]
        ppo = R.lib.create_ppo(cabi.cstr_array(["relu", "relu", "none"]), cabi.int_array(sizes), 4, c["capacity"],
                               C.c_float(3e-4), C.c_float(3e-4), C.c_float(0.95), C.c_float(0.2), C.c_float(0.0),
                               C.c_float(1.0), False)
        if c["mu_bias"] is not None:
            mu = ppo.contents.policy.contents.mu
            p = R.nn_get_params(mu)
            p[-1] = c["mu_bias"]
            R.nn_set_params(mu, p)
        for _ in range(c["train_epochs"]):
            R.lib.train_ppo_epoch(ppo, env, c["capacity"], 64, 1, 2)
        line = capture_stdout(lambda: R.lib.eval_ppo(ppo, env, c["steps"])).strip()
        out.append(dict(c, line=line, rand_after=cabi.rand()))
        print(c, "->", line)
    with open(os.path.join(HERE, "eval_golden.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

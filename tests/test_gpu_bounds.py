"""Out-of-bounds guard: every array the kernels read or write is a slice in the middle of a larger buffer filled with a
sentinel; after each entry point (K1 lean / full / scalar path, step_many, K3 reset, K2 fast / generic / both sweep kernels)
the pads on either side must be untouched.  (compute-sanitizer is not available on the GPU pool; this is the bounds check
the kernels get instead.)"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
PAD = 64   # elements: keeps every slice 16-byte aligned, so even batches still take the vector path
SENT = {"torch.float64": 12345.678, "torch.int32": 0x5A5A5A5A, "torch.uint8": 0xA5}


class Guarded:
    def __init__(self):
        self.bufs = []

    def make(self, shape, dtype):
        import torch

        n = int(np.prod(shape))
        buf = torch.full((n + 2 * PAD,), SENT[str(dtype)], dtype=dtype, device="cuda")
        self.bufs.append((buf, n))
        return buf[PAD:PAD + n].view(*shape)

    def check(self, what):
        for k, (buf, n) in enumerate(self.bufs):
            s = SENT[str(buf.dtype)]
            lo, hi = buf[:PAD], buf[PAD + n:]
            assert bool((lo == s).all()) and bool((hi == s).all()), f"{what}: guard region of buffer {k} was written"


def guarded_batch(kind, n, g, **kw):
    import torch

    from gym_fishing_b200.engine import EnvBatch

    b = EnvBatch(kind, n, **kw)
    b.state = g.make((n,), torch.float64)
    b.years = g.make((n,), torch.int32)
    b.years.zero_()
    if b.K_env is not None:
        b.K_env = g.make((n,), torch.float64)
        b.r_env = g.make((n,), torch.float64)
    return b


@pytest.mark.parametrize("kind", [1, 2, 4])
@pytest.mark.parametrize("n", [4096, 4098, 4097, 62, 1])
def test_k1_and_reset_stay_inside_their_arrays(kind, n):
    import torch

    g = Guarded()
    b = guarded_batch(kind, n, g, auto_reset=True, r=0.3, K=1.0 if kind != 2 else 1.5, sigma=0.05, x0=0.75, Tmax=3, seed=5)
    b.reset()
    g.check("reset")
    T = 6
    acts = g.make((T, n), torch.float64)
    acts.copy_(torch.rand(T, n, dtype=torch.float64, device="cuda") * 2 - 1)
    obs, rew = g.make((T, n), torch.float64), g.make((T, n), torch.float64)
    done, term = g.make((T, n), torch.uint8), g.make((T, n), torch.float64)
    noise = g.make((T, n), torch.float64)
    noise.normal_()
    state0 = b.state
    for t in range(T):     # full kernel: injected noise + terminal observations; state array kept apart from the outputs
        b.state = state0
        b.step(acts[t], noise=noise[t], out_obs=obs[t], out_reward=rew[t], out_done=done[t], out_terminal=term[t])
        state0.copy_(obs[t])
    g.check("K1 with optional arrays")
    b.state = state0
    for t in range(T):     # lean kernel: in-kernel Philox, no optional arrays
        b.state = state0
        b.step(acts[t], out_obs=obs[t], out_reward=rew[t], out_done=done[t])
        state0.copy_(obs[t])
    g.check("K1 lean")
    b.state = state0
    b.step_many(acts, obs, rew, done)
    g.check("step_many")
    assert bool(done.any()) and bool(torch.isfinite(rew).all())


@pytest.mark.parametrize("kind", [1, 4])
@pytest.mark.parametrize("levels", [0, 32, 8, 3])
def test_k2_stays_inside_its_arrays(kind, levels):
    import torch

    reps = 6
    n = (levels * reps) if levels else 4098
    T = 40
    for mode in ("plain", "auto_reset", "freeze"):
        for traced in (False, True):
            g = Guarded()
            b = guarded_batch(kind, n, g, auto_reset=(mode == "auto_reset"), r=0.3, K=1.0, sigma=0.1, x0=0.75, Tmax=9, seed=2)
            b.reset()
            ret, ds = g.make((n,), torch.float64), g.make((n,), torch.int32)
            S_env = None
            if levels:
                S_env = g.make((n,), torch.float64)
                S_env.copy_(torch.linspace(0, 1, levels, dtype=torch.float64, device="cuda").repeat(reps))
            b.rollout(T, policy="escapement", S=None if levels else 0.5, S_env=S_env, noise_id_div=levels,
                      freeze_done=(mode == "freeze"), trace=traced, out_return=ret, out_done_step=ds)
            torch.cuda.synchronize()
            g.check(f"K2 kind={kind} levels={levels} mode={mode} traced={traced}")
            assert bool(torch.isfinite(ret).all())

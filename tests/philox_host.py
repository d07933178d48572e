"""TEST INFRASTRUCTURE — the random draws of the Philox-mode CUDA kernels, rebuilt on the host.

The product kernels draw from Philox4x32-7 keyed by the env seed with counter (global env id | quad id, step
counter, stream) and map the 32-bit words to env decisions through Walker alias tables / multiply-high
(csrc/gpt_common.cuh, gpt_taxi.cu ``taxi_fix_inline``, gpt_rooms_kernel.cuh, gpt_msrooms.cu).  ``PhiloxDraws``
recomputes exactly those decisions with numpy and hands them to the oracle envs through the oracle's draw-source
protocol (oracle/draws.py), so that the oracle — the restatement of the REFERENCE's step — can be stepped beside
the Philox kernels (single-step and fused multi-step launches) and compared bit for bit.

Independent of the product code: Philox is restated here from the published algorithm (Salmon et al., SC'11;
known-answer vectors of Random123 in ``test_philox_known_answers``); the alias tables are read back from the
handle with ``gpt_table_read`` because they are *data* the kernel consumes, and the law they encode is checked
separately (chi-square tests in test_taxi_gpu.py / test_rooms_gpu.py).
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


RESET_ROUNDS = STEP_ROUNDS = 7   # csrc/gpt_common.cuh kRounds: every draw is Philox4x32-7


def philox4x32(c0, c1, c2, c3, k0, k1, rounds=10):
    """Vectorised Philox4x32-R (default 10 rounds).  c*: uint32-valued arrays (any broadcastable shapes), k0/k1: python ints."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK for c in (c0, c1, c2, c3))
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    for r in range(rounds):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        kk0 = np.uint64((k0 + r * W0) & 0xFFFFFFFF)
        kk1 = np.uint64((k1 + r * W1) & 0xFFFFFFFF)
        c0, c1, c2, c3 = hi1 ^ c1 ^ kk0, lo1, hi0 ^ c3 ^ kk1, lo0
    return c0, c1, c2, c3


def mulhi(u, n):
    """floor(u * n / 2^32) — the kernels' ``bounded``."""
    return (np.asarray(u, dtype=np.uint64) * np.uint64(n)) >> np.uint64(32)


class PhiloxDraws:
    """Draw source for the oracle envs that returns what the CUDA kernels draw at Philox step ``counter``.

    Set ``counter`` to the handle's step counter (``env.rng_counter``) before every oracle ``reset`` / ``step``
    (inside a fused launch: counter of the launch + t)."""

    def __init__(self, env, seed, family, env_offset=0):
        self.family = family
        self.seed = int(seed) & (2**64 - 1)
        self.k0, self.k1 = self.seed & 0xFFFFFFFF, self.seed >> 32
        self.counter = 0
        self.b = env.num_envs
        self.gid = np.arange(self.b, dtype=np.uint64) + np.uint64(int(env_offset))
        if family == "taxi":
            al = env.read_table("reset_alias", np.uint32).reshape(-1, 2)
            self.alias_thr, self.alias_val = al[:, 0].astype(np.uint64), al[:, 1]
            self.n_valid = len(al)
            self.nlocs = env.nlocs
        else:
            n = env.single_action_space.n
            al = env.read_table("slip_alias", np.uint32).reshape(n, 8, 2)
            self.slip_thr, self.slip_val = al[..., 0].astype(np.uint64), al[..., 1]
            self.n_actions = n
            self.log2n = {4: 2, 8: 3}[n]
            self.spawn = env.read_table("spawn_cells", np.uint16).astype(np.int64)
            self.goals = env.read_table("goal_cells", np.uint16).astype(np.int64)
        self._cache = {}

    def reseed(self, seed):
        self.seed = int(seed) & (2**64 - 1)
        self.k0, self.k1 = self.seed & 0xFFFFFFFF, self.seed >> 32

    # one Philox block per (id, step counter, stream), memoised per counter value
    def _block(self, ids, stream, rounds=RESET_ROUNDS):
        key = (self.counter, stream, len(ids), rounds)
        if key not in self._cache:
            self._cache = {k: v for k, v in self._cache.items() if k[0] == self.counter}
            ctr_lo = self.counter & 0xFFFFFFFF
            ctr_hi = ((self.counter >> 32) & 0x00FFFFFF) ^ (stream << 24)
            self._cache[key] = philox4x32(ids & MASK, ids >> np.uint64(32), ctr_lo, ctr_hi, self.k0, self.k1, rounds)
        return self._cache[key]

    # ---- Taxi (gpt_taxi.cu taxi_fix_inline) -----------------------------------------------------
    def multinomial_argmax(self, n, pvals, b, where=None, **ctx):
        x = self._block(self.gid, 0)[0][where]
        w = x * np.uint64(self.n_valid)                       # 64-bit product: column = high word, fraction = low word
        col = (w >> np.uint64(32)).astype(np.int64)
        own = (w & MASK) < self.alias_thr[col]
        v = self.alias_val[col]
        return np.where(own, v & 0xFFFF, v >> 16).astype(np.int64)

    def integers(self, high, size=None, where=None, kind=None, **ctx):
        if self.family == "taxi":
            _, y, z, _ = self._block(self.gid, 0)
            p = mulhi(y[where], self.nlocs).astype(np.int64)
            if kind == "new_p":
                return p
            if kind == "new_d":   # uniform over the other locations: never clashes, the reference's redraw loop is a no-op
                d = mulhi(z[where], self.nlocs - 1).astype(np.int64)
                return d + (d >= p)
        raise AssertionError(f"unexpected integers() call kind={kind}")

    # ---- ROOMS / MSROOMS (gpt_rooms_kernel.cuh, gpt_msrooms.cu) ------------------------------------
    def random(self, b, kind=None, action=None, cumsum=None, **ctx):
        assert kind == "slip" and b == self.b
        quad = self.gid >> np.uint64(2)
        blk = self._block(quad, 0, STEP_ROUNDS)                # one Philox4x32-7 block per quad of envs; word k belongs to env 4*quad + k
        lane = (self.gid & np.uint64(3)).astype(np.int64)
        u32 = np.choose(lane, blk)
        a = np.asarray(action).astype(np.int64) & (self.n_actions - 1)
        col = (u32 >> np.uint64(32 - self.log2n)).astype(np.int64)
        frac = (u32 << np.uint64(self.log2n)) & MASK
        own = frac < self.slip_thr[a, col]
        v = self.slip_val[a, col]
        d8 = np.where(own, v & 0xFF, v >> 8).astype(np.int64)  # ordinal-direction units
        a2 = d8 >> (1 if self.n_actions == 4 else 0)
        # a float64 u the reference's sampler maps to a2: cumsum[a2-1] < u <= cumsum[a2]  (action_utils.py:84-90)
        rows = cumsum[a]
        lo = np.where(a2 > 0, rows[np.arange(b), np.maximum(a2 - 1, 0)], 0.0)
        hi = rows[np.arange(b), a2]
        return (lo + hi) / 2

    def choice(self, values, b, where=None, kind=None, **ctx):
        x, y, _, _ = self._block(self.gid, 1)
        if kind == "reset_goal":
            return self.goals[mulhi(y[where], len(self.goals)).astype(np.int64)]
        if kind == "reset_agent":
            return self.spawn[mulhi(x[where], len(self.spawn)).astype(np.int64)]
        raise AssertionError(f"unexpected choice() call kind={kind}")

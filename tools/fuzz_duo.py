"""Randomised cross-check of duo mode (run on a GPU box): random state counts, component counts and weights, random trees
(multifurcations, unary nodes, zero-length branches), random site ranges around superblock boundaries -- the tiled kernel
against the generic one (same bytes expected) and, every few rounds, against the executable spec on the CPU.

  python tools/fuzz_duo.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import elynx_b200 as eb  # noqa: E402
from helpers import random_tree, spec_freerun  # noqa: E402


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
    t_end = time.time() + budget
    rounds = specs = 0
    while time.time() < t_end:
        k = int(rng.integers(5, 23))
        c = int(rng.choice([1, 2, 3, 4, 5, 8, 12, 17, 40]))
        ds = rng.dirichlet(np.ones(k) * rng.uniform(0.3, 5.0), size=c)
        es = rng.uniform(0.05, 3.0, size=(c, k, k))
        es = (es + es.transpose(0, 2, 1)) / 2
        for i in range(c):
            np.fill_diagonal(es[i], 0.0)
        ws = rng.uniform(0.05, 1.0, size=c) if rng.random() < 0.7 else np.ones(c)
        is_mix = c > 1 or rng.random() < 0.5
        tree = random_tree(int(rng.integers(2, 70)), rng, multifurcate=bool(rng.random() < 0.4), zero_frac=float(rng.choice([0.0, 0.1])))
        sb = 1 << (20 if c <= 16 else 22)
        site0 = int(rng.choice([0, 17, sb - int(rng.integers(1, 5000)), 3 * sb + int(rng.integers(0, sb)), (1 << 33) + sb - 3]))
        n = int(rng.choice([1, 15, 16, 17, 511, 4097, int(rng.integers(1, 300_000))]))
        seed = int(rng.integers(0, 2 ** 63))
        model = (is_mix, ws, ds, es, None)
        outs = []
        for fg in (False, True):
            with eb.Simulator(ws if is_mix else None, ds, es, tree, is_mixture=is_mix, force_generic=fg) as sim:
                assert sim.duo_groups > 0 or tree.n_nodes == 1
                leaves, comps, _ = sim.simulate(seed, n, site0=site0, want_comps=True)
                outs.append((leaves, comps))
                if not fg and rounds % 5 == 0 and n <= 20000:
                    l0, c0, _ = spec_freerun(sim, model, tree, seed, site0, n)
                    assert np.array_equal(leaves, l0) and np.array_equal(comps, c0), ("spec", k, c, tree.n_leaves, site0, n, seed)
                    specs += 1
        assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1]), ("tiled != generic", k, c, tree.n_leaves, site0, n, seed)
        rounds += 1
    print("fuzz_duo ok: %d rounds (%d against the spec) in %.0f s" % (rounds, specs, budget))


if __name__ == "__main__":
    main()

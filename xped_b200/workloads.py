"""Synthetic bases of the benchmark workloads (BASELINE.json configs): host metadata only, shared by bench.py and the tests."""
import math


def hubbard_bond_basis(m=2000, sigma_n=2.0, sigma_s=2.0, n0=0):
    """configs[2]: bond basis of a U(1) x U(1) (fermion number N, 2 Sz) Hubbard MPS with ~m states: sector dimensions follow a Gaussian over (N - n0, 2 Sz),
    only N + 2 Sz even (what fusing four-state Hubbard sites produces); sectors whose share is below 0.2 states are dropped, the others have at least one
    state: 89 sectors, 2019 states, largest block 159 for m = 2000 ("88-89 sectors, max block 159").  Returns [((N, 2Sz), dim)] sorted."""
    weights = {}
    for n in range(n0 - 10, n0 + 11):
        for s in range(-10, 11):
            if (n + s) % 2:
                continue
            weights[(n, s)] = math.exp(-0.5 * ((n - n0) / sigma_n) ** 2 - 0.5 * (s / sigma_s) ** 2)
    tot = sum(weights.values())
    out = []
    for q in sorted(weights):
        x = m * weights[q] / tot
        if x >= 0.2:
            out.append((q, max(1, int(round(x)))))
    return out

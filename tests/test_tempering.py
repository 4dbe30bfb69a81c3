"""Parallel-tempering swap logic (extension; host-side, CPU): pure-function invariants."""
import numpy as np


def test_ladder_and_energy():
    from tonga_b200.tempering import energy, geometric_ladder
    b = geometric_ladder(32, 50.0)
    assert b[0] == 1.0 and abs(b[-1] - 1 / 50) < 1e-15 and (np.diff(b) < 0).all()
    assert np.allclose(b[1:] / b[:-1], b[1] / b[0])
    assert np.allclose(energy([10.0], [1.0], 381), [5.0]) and np.allclose(energy([10.0], [np.e], 381), [5.0 + 381])


def test_swap_step_invariants():
    from tonga_b200.tempering import geometric_ladder, swap_step
    rng = np.random.default_rng(0)
    T, L = 8, 5
    beta0 = np.tile(geometric_ladder(T, 20.0), L)
    E = rng.uniform(50, 400, T * L)
    b1, acc, att = swap_step(E, beta0, T, step=0, seed=3)
    b1b, _, _ = swap_step(E, beta0, T, step=0, seed=3)
    assert np.array_equal(b1, b1b)                      # deterministic: every rank takes the same decisions
    assert att == L * 4 and 0 <= acc <= att             # even sweep: pairs (0,1),(2,3),(4,5),(6,7)
    for l in range(L):                                  # betas are permuted inside each ladder only
        assert np.array_equal(np.sort(b1[l * T:(l + 1) * T]), np.sort(beta0[:T]))
    _, _, att_odd = swap_step(E, beta0, T, step=1, seed=3)
    assert att_odd == L * 3                             # odd sweep: (1,2),(3,4),(5,6)
    # equal energies -> every attempted swap is accepted (log alpha = 0)
    b2, acc2, att2 = swap_step(np.full(T * L, 7.0), beta0, T, step=0, seed=1)
    assert acc2 == att2
    # a colder replica with the HIGHER energy always swaps with its hotter neighbour (log alpha > 0)
    E3 = np.tile(np.arange(T, 0, -1.0) * 100, L)
    _, acc3, att3 = swap_step(E3, beta0, T, step=0, seed=5)
    assert acc3 == att3


def test_swap_acceptance_rate_matches_formula():
    from tonga_b200.tempering import swap_step
    beta = np.array([1.0, 0.5])
    E = np.array([10.0, 12.0])  # log alpha = (1 - 0.5) * (10 - 12) = -1
    acc = sum(swap_step(E, beta, 2, step=2 * s, seed=9)[1] for s in range(4000))
    assert abs(acc / 4000 - np.exp(-1.0)) < 0.03

"""CPU check of the numerical DESIGN of the GPU path: tests/algo_model.py restates the kernels' algorithm
(Taylor-series trajectory → per-cell polynomials of F → lookup items → per-tree sums, and the moment
contraction) in numpy and is compared with the oracle.  This validates the order/step rule without a GPU;
the CUDA kernels themselves are compared with the oracle in test_gpu_parity.py."""
import numpy as np
import pytest

import algo_model as am
import gcdyn
from helpers import load_golden, rel


@pytest.mark.parametrize("name,key", [("g1_config1", "ode_closed"), ("g2_config2", "ode_branch"),
                                      ("g5_batched_discrete", "ode_shared"), ("g4_no_extinction", "stadler")])
def test_taylor_lookup_model_matches_oracle(name, key):
    flat, params, T, exp, par = load_golden(name)
    items = am.build_items(flat, T)
    want = exp[key]
    for p in range(want.shape[0]):
        S = am.auto_steps(T, par["lam"][p], par["mu"][p], par["delta"][p], par["Gamma"][p], 16)
        out, err = am.loglik_per_tree(flat, items, par["lam"][p], par["mu"][p], par["delta"][p], par["Gamma"][p],
                                      par["rho"][p], par["sigma"][p], T, S=S, M=16)
        assert err <= 1e-11          # the library's default acceptance threshold for the indicator
        assert rel(out, want[p]) <= 1e-10


def test_moment_contraction_equals_item_sweep():
    flat, params, T, exp, par = load_golden("g2_config2")
    items = am.build_items(flat, T)
    S, M, K = 24, 16, flat.n_types
    Mom = am.moments(items, T, S, M, K)
    Fc, pT, err, bad = am.taylor_table(par["lam"][0], par["mu"][0], par["delta"][0], par["Gamma"][0],
                                       par["rho"][0], par["sigma"][0], T, S, M)
    per_tree, _ = am.loglik_per_tree(flat, items, par["lam"][0], par["mu"][0], par["delta"][0], par["Gamma"][0],
                                     par["rho"][0], par["sigma"][0], T, S, M)
    with np.errstate(divide="ignore"):
        loglam = np.log(par["lam"][0])
        logG = np.log(par["delta"][0] * par["Gamma"][0]).reshape(-1)
    node_terms = (am._xlogy_count(items["nb"].sum(0), loglam).sum() + am._xlogy_count(items["ntc"].sum(0), logG).sum()
                  + items["nd"].sum() * (np.log(par["mu"][0]) + np.log(par["sigma"][0]))
                  + items["ns"].sum() * np.log(par["rho"][0]))
    total = (Fc * Mom).sum() + node_terms - np.log1p(-pT)[flat.root_type_idx].sum()
    assert abs(total - per_tree.sum()) <= 1e-12 * abs(total)
    assert abs(total - exp["ode_branch"].sum()) <= 1e-10 * abs(total)


def test_item_count_is_below_node_count():
    """Regrouping by node time needs no lookup for survivors: items ≈ births + deaths + 2·type changes + 1."""
    flat, _, T, exp, _ = load_golden("g2_config2")
    items = am.build_items(flat, T)
    c = exp["event_counts"].sum(axis=0)
    assert len(items["time"]) == c[1] + c[2] + 2 * c[4] + flat.n_trees
    assert len(items["time"]) < flat.n_nodes


@pytest.mark.parametrize("name,key", [("g1_config1", "ode_closed"), ("g2_config2", "ode_branch"),
                                      ("g4_no_extinction", "stadler"), ("g5_batched_discrete", "ode_shared")])
def test_chebyshev_moment_model_matches_oracle(name, key):
    """The GEMM formulation: one Chebyshev series of F_x per pset × pset-independent per-tree Chebyshev
    moments.  Degree 32 already reproduces the golden per-tree values; the tail of the series bounds the error."""
    flat, params, T, exp, par = load_golden(name)
    items = am.build_items(flat, T)
    Mt = am.cheb_tree_moments(flat, items, T, 64)
    for p in range(exp[key].shape[0]):
        S = am.auto_steps(T, par["lam"][p], par["mu"][p], par["delta"][p], par["Gamma"][p], 16)
        coarse, tail16 = am.loglik_per_tree_cheb(flat, items, Mt, par["lam"][p], par["mu"][p], par["delta"][p],
                                                 par["Gamma"][p], par["rho"][p], par["sigma"][p], T, S, 16, 16)
        out, tail = am.loglik_per_tree_cheb(flat, items, Mt, par["lam"][p], par["mu"][p], par["delta"][p],
                                            par["Gamma"][p], par["rho"][p], par["sigma"][p], T, S, 16, 48)
        assert tail <= 1e-11 and rel(out, exp[key][p]) <= 1e-10
        # an unresolved series announces itself through its tail (what triggers the sweep fallback on the GPU)
        assert tail16 > 1e-9 and rel(coarse, exp[key][p]) > rel(out, exp[key][p])


def test_high_word_log2_used_by_the_step_controller():
    """traj_kernel's step control reads log2(ind) off the HIGH WORD of the double: (hi − 0x3ff00000)/2^20 = exponent +
    linear mantissa.  Error < 0.0861 (the classic bound of log2(1+m) − m), monotone in x, so (tol/ind)^(1/M) is within
    0.8 % for M ≥ 16 and the accept test hi(ind) < hi(tol) is never looser than ind < tol."""
    import numpy as np
    x = np.concatenate([np.exp(np.random.default_rng(0).uniform(np.log(1e-300), np.log(1e300), 200000)),
                        [1e-11, 2.3546e-12, 1.0, 0.5, 2.0 ** -1022]])
    hi = (x.view(np.uint64) >> np.uint64(32)).astype(np.int64)
    approx = (hi - 0x3FF00000) / 2.0 ** 20
    err = np.log2(x) - approx
    assert err.min() >= -1e-6 and err.max() < 0.0861          # the truncated low word only lowers approx by < 2^-20
    order = np.argsort(x)
    assert np.all(np.diff(hi[order]) >= 0)                     # monotone: integer compare == ordering of the doubles
    for M in (8, 12, 16, 20, 24):
        worst = 2.0 ** (0.0861 * 2 / M) - 1.0                  # both tol and ind are approximated
        assert worst < (0.016 if M == 8 else 0.011 if M == 12 else 0.0076)
    tol = 1e-11
    tol_hi = int(np.array([tol]).view(np.uint64)[0] >> np.uint64(32))
    accepted = x[hi < tol_hi]
    assert accepted.max() < tol                                # never accepts what ind ≤ tol would reject
    rejected_below_tol = x[(hi >= tol_hi) & (x < tol)]
    assert rejected_below_tol.size == 0 or rejected_below_tol.min() > tol * (1 - 2e-6)

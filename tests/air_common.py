"""Shared by tests/test_air_host.py (CPU) and tests/test_gpu_air.py (B200): the oracle's coefficient-form
Stark::prove on the real Rescue-Prime AIR, instrumented to hand out what the evaluation-form route consumes
(committed codewords, weights, shifts, zerofiers, interpolants) and what it must reproduce (transition
quotients, the combined codeword)."""
import numpy as np

from oracle import cbind as C, ntt as N
from oracle.stark import Backend, RPSSS, deterministic_rng


def recording_backend(log):
    """oracle Backend that records what Stark::prove computes between the commits and FRI"""
    class Rec(Backend):
        @staticmethod
        def fast_coset_evaluate(w, n, off, p):
            out = N.fast_coset_evaluate(w, n, off, p)
            log.setdefault("lde", []).append(out)
            return out

        @staticmethod
        def fast_coset_divide(root, order, off, a, b):
            out = N.fast_coset_divide(root, order, off, a, b)
            log.setdefault("div", []).append(out)
            return out
    return Rec


def prover_intermediates(seed=b"air"):
    """Runs the oracle RPSSS signature and returns everything the evaluation-form kernel consumes / must reproduce."""
    log = {}
    r = RPSSS(4, 64, 128, 3, backend=recording_backend(log))
    st = r.stark
    orig = st.sample_weights
    st.sample_weights = lambda number, randomness: log.setdefault("weights", orig(number, randomness))
    rng = deterministic_rng(seed)
    sk, pk = r.keygen(rng)
    tcs = r.transition_constraints()
    boundary = r.rp.boundary_constraints(pk)
    sig = r.sign(sk, b"doc", rng)
    nr, nc, n = st.num_registers, len(tcs), st.fri_domain_length
    trace_len = st.original_trace_length + st.num_randomizers
    zerofiers = st.boundary_zerofiers(boundary)
    tcd = st.max_degree(tcs)
    shifts = [tcd - b for b in st.transition_quotient_degree_bounds(tcs)] + [tcd - b for b in st.boundary_quotient_degree_bounds(trace_len, boundary)]
    return dict(stark=st, tcs=tcs, boundary=boundary, signature=sig, nr=nr, nc=nc, n=n,
                bq_cws=log["lde"][:nr], rnd_cw=log["lde"][nr], combined=log["lde"][nr + 1],
                bq_polys=log["div"][:nr], tq_polys=log["div"][nr:nr + nc], weights=log["weights"], shifts=shifts,
                zerofiers=zerofiers, interpolants=st.boundary_interpolants(boundary), tz=st.transition_zerofier())


def flatten(tcs, nr):
    nvars = 1 + 2 * nr
    counts, coefs, exps = [], [], []
    for tc in tcs:
        counts.append(len(tc.dictionary))
        for key, coef in tc.dictionary.items():
            exps.append(list(key[:nvars]) + [0] * (nvars - len(key)))
            coefs.append(coef)
    return np.asarray(counts, dtype=np.uint32), C.to_arr(coefs), np.asarray(exps, dtype=np.uint32)

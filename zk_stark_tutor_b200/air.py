"""Evaluation-form transition quotients and nonlinear combination (zkb_air_combination, csrc/air.cu):
the middle of Stark::prove (src/stark/stark.rs:388-519) computed pointwise on the FRI coset from the
committed codewords, without leaving HBM.  A transition constraint is the reference's MPolynomial
dictionary (src/m_polynomial.rs): {tuple of exponents over (x, registers now, registers next): coefficient}."""
import ctypes

import numpy as np

from . import _lib
from .context import Vec, default_context, pack


def flatten_constraints(constraints, num_registers):
    """[{exponent tuple: coefficient}] -> (term_counts, coefs (T, 2) uint64, exps (T, 1 + 2*num_registers) uint32).
    Keys shorter than the variable count are zero padded, as MPolynomial::evaluate treats them (m_polynomial.rs:97-126)."""
    nvars = 1 + 2 * num_registers
    counts, coefs, exps = [], [], []
    for c in constraints:
        d = getattr(c, "dictionary", c)
        counts.append(len(d))
        for key, coef in d.items():
            key = tuple(key)
            if len(key) > nvars and any(key[nvars:]):
                raise ValueError("constraint uses more than 1 + 2*num_registers variables")
            exps.append(list(key[:nvars]) + [0] * (nvars - len(key)))
            coefs.append(int(coef))
    return (np.asarray(counts, dtype=np.uint32), pack(coefs),
            np.asarray(exps, dtype=np.uint32).reshape(len(coefs), nvars))


def air_combination(offset, omega, domain_length, expansion_factor, constraints, boundary_zerofiers, boundary_interpolants,
                    transition_zerofier, weights, shifts, bq_codewords, randomizer_codeword, want_quotients=False, ctx=None):
    """bq_codewords: CUDA tensor (num_registers, domain_length, 2); randomizer_codeword: CUDA tensor (domain_length, 2).
    Returns the combined codeword as a CUDA tensor (and the transition-quotient codewords if asked)."""
    import torch
    ctx = ctx or default_context()
    nr, n = bq_codewords.shape[0], domain_length
    assert tuple(bq_codewords.shape) == (nr, n, 2) and bq_codewords.is_cuda and bq_codewords.is_contiguous()
    assert tuple(randomizer_codeword.shape) == (n, 2) and randomizer_codeword.is_cuda and randomizer_codeword.is_contiguous()
    # a (term_counts, coefs, exps) tuple is taken as already flattened (a prover flattens its AIR once, not per proof)
    counts, coefs, exps = constraints if isinstance(constraints, tuple) else flatten_constraints(constraints, nr)
    nc = len(counts)
    assert len(weights) == 1 + 2 * nc + 2 * nr and len(shifts) == nc + nr
    d = _lib.AirDesc()
    d.offset[:] = list(int(offset).to_bytes(16, "little"))
    d.omega[:] = list(int(omega).to_bytes(16, "little"))
    d.domain_length, d.expansion_factor, d.num_registers, d.num_constraints = n, expansion_factor, nr, nc
    keep = [counts, coefs, exps]
    d.term_counts = counts.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
    d.coefs = coefs.ctypes.data_as(_lib.c_u8p)
    d.exps = exps.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))

    def poly_list(polys):
        vecs = [Vec(list(p) if not isinstance(p, np.ndarray) else p) for p in polys]
        ptrs = (ctypes.c_void_p * nr)(*[v.ptr if v.n else None for v in vecs])
        lens = (ctypes.c_size_t * nr)(*[v.n for v in vecs])
        keep.extend([vecs, ptrs, lens])
        return ptrs, lens
    assert len(boundary_zerofiers) == nr and len(boundary_interpolants) == nr
    d.boundary_zerofiers, d.boundary_zerofier_lens = poly_list(boundary_zerofiers)
    d.boundary_interpolants, d.boundary_interpolant_lens = poly_list(boundary_interpolants)
    tz = Vec(list(transition_zerofier))
    d.transition_zerofier, d.transition_zerofier_len = tz.ptr, tz.n
    w = pack([int(x) for x in weights])
    sh = np.asarray([int(x) for x in shifts], dtype=np.uint64)
    keep.extend([tz, w, sh])
    d.weights = w.ctypes.data_as(_lib.c_u8p)
    d.shifts = sh.ctypes.data_as(_lib.c_u64p)
    out = torch.empty((n, 2), dtype=torch.int64, device=bq_codewords.device)
    tq = torch.empty((nc, n, 2), dtype=torch.int64, device=bq_codewords.device) if want_quotients else None
    ctx.check(ctx.lib.zkb_air_combination(ctx.h, ctypes.byref(d), bq_codewords.data_ptr(), n, randomizer_codeword.data_ptr(),
                                          out.data_ptr(), tq.data_ptr() if want_quotients else None))
    del keep
    return (out, tq) if want_quotients else out

"""FRI::prove (src/fri.rs:115-248) assembled from the C oracle's kernels so that whole
transcripts can be produced at 2^16..2^24 in seconds (test infrastructure only).

Same semantics as oracle.fri.FRI.prove - which is the literal restatement and is what this
module is checked against in tests/test_oracle_c.py - but each tree is built once
(zo_merkle with all nodes) instead of once per opening.
"""
import numpy as np

from . import cbind as C
from . import field as F
from . import proof_stream as PS
from .fri import FRI


def _vals(arr, idx):
    return int(arr[idx, 0]) | (int(arr[idx, 1]) << 64)


class Tree:
    def __init__(self, cw):
        self.n = len(cw)
        self.root, self.nodes = C.merkle(cw, want_nodes=True)     # levels bottom-up, concatenated
        self.off, o, m = [], 0, self.n
        while m >= 1:
            self.off.append(o)
            o += m
            m //= 2

    def open(self, index):
        path = []
        for lvl in range(len(self.off) - 1):
            path.append(self.nodes[self.off[lvl] + (index ^ 1)].tobytes())
            index >>= 1
        return path


def commit(fri: FRI, codeword, ps):
    """-> (codewords, trees, alphas); pushes Roots and the last Codeword to ps."""
    omega, offset = fri.omega, fri.offset
    rounds = fri.num_rounds()
    cw = np.ascontiguousarray(codeword, dtype=np.uint64).reshape(-1, 2)
    codewords, trees, alphas = [], [], []
    for r in range(rounds):
        t = Tree(cw)
        ps.push((PS.ROOT, t.root))
        trees.append(t)
        codewords.append(cw)
        if r == rounds - 1:
            break
        alpha = F.sample(ps.fiat_shamir_prover(PS.PROOF_BYTES))
        alphas.append(alpha)
        cw = C.fri_fold(cw, alpha, offset, omega)
        omega = F.mul(omega, omega)
        offset = F.mul(offset, offset)
    ps.push((PS.CODEWORD, C.from_arr(cw)))
    return codewords, trees, alphas


def prove(fri: FRI, codeword, ps):
    assert fri.domain_length == len(codeword)
    codewords, trees, _ = commit(fri, codeword, ps)
    ncc = fri.num_colinearity_tests
    top = FRI.sample_indices(ps.fiat_shamir_prover(PS.PROOF_BYTES), len(codewords[1]), len(codewords[-1]), ncc)
    indices = list(top)
    for i in range(len(codewords) - 1):
        half = len(codewords[i]) // 2
        indices = [j % half for j in indices]
        cur, nxt = codewords[i], codewords[i + 1]
        for s in range(ncc):
            ps.push((PS.LEAFS, (_vals(cur, indices[s]), _vals(cur, indices[s] + half), _vals(nxt, indices[s]))))
        for s in range(ncc):
            ps.push((PS.PATH, trees[i].open(indices[s])))
            ps.push((PS.PATH, trees[i].open(indices[s] + half)))
            ps.push((PS.PATH, trees[i + 1].open(indices[s])))
    return top, codewords, trees

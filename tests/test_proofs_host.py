"""Host-side logic of the proof-batch path (zk_stark_tutor_b200/proofs.py): sizes, the packed coefficient
layout, the vectorised quadrupled indices against the literal stark.rs:534-542 form, and the round-robin
deal of proofs to ranks.  CPU only."""
import random

import numpy as np

from zk_stark_tutor_b200 import proofs


def test_shape_matches_the_tutorial_parameters():
    s = proofs.ProofShape()
    assert (s.omicron_len, s.fri_len, s.ef, s.ncc) == (1024, 4096, 4, 64)
    assert s.column_lengths() == [282, 282, 1024] and s.comb_len == 1024 and s.max_degree == 1023


def test_quadrupled_indices_vectorised_form():
    r = random.Random(3)
    n, ef, ncc, B = 4096, 4, 64, 9
    top = np.array([[r.randrange(n // 2) for _ in range(ncc)] for _ in range(B)], dtype=np.uint64)
    nn, e = np.uint64(n), np.uint64(ef)
    dup = np.concatenate([top, (top + e) % nn], axis=1)
    quad = np.sort(np.concatenate([dup, (dup + nn // np.uint64(2)) % nn], axis=1), axis=1)
    for b in range(B):
        want = proofs.quadrupled_indices([int(x) for x in top[b]], n, ef)
        assert quad[b].tolist() == want
        assert len(want) == 4 * ncc and want == sorted(want)          # sorted, NOT deduplicated


def test_pack_batch_layout():
    s = proofs.ProofShape()
    rng = np.random.default_rng(1)
    batch = []
    for _ in range(3):
        cols = [rng.integers(0, 1 << 62, size=(ln, 2), dtype=np.uint64) for ln in s.column_lengths()]
        batch.append((cols, rng.integers(0, 1 << 62, size=(s.comb_len, 2), dtype=np.uint64)))
    packed = proofs.pack_batch(s, batch)
    K = len(s.column_lengths())
    assert packed.shape == (K + 1, 3, s.comb_len, 2) and packed.dtype == np.uint64
    for b, (cols, comb) in enumerate(batch):
        for t, col in enumerate(cols):
            assert np.array_equal(packed[t, b, :col.shape[0]], col)
            assert not packed[t, b, col.shape[0]:].any()                # zero padding: same LDE
        assert np.array_equal(packed[K, b], comb)


def test_partition_covers_every_proof_once():
    for world in (1, 2, 3, 8):
        seen = sorted(i for r in range(world) for i in proofs.partition(37, world, r))
        assert seen == list(range(37))

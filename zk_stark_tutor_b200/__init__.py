"""zk_stark_tutor_b200 - B200-native NTT / coset LDE / Merkle / FRI engine behind the
function signatures of SpekalsG3/zk-stark-tutor's hot path (src/fft, src/merkle_root.rs,
src/fri.rs).  The compute lives in lib/libzkb200.so (hand-written CUDA for sm_100a behind
the C ABI of include/zkb200.h); these modules are the Python mirror of the reference's
module layout.  No CPU fallback: without the built library and a CUDA device every
compute call raises."""
from .context import Context, ZkbError, default_context, pack, unpack, P      # noqa: F401
from .field import Field, FIELD_PRIME                                           # noqa: F401
from .fft import (ntt, intt, ntt_batch, scale, fast_coset_evaluate, coset_lde_batch,   # noqa: F401
                  fast_multiply, fast_coset_divide, fast_zerofier, fast_evaluate_domain, fast_interpolate_domain)
from .merkle_root import MerkleRoot, MerkleTree                                  # noqa: F401
from .proof_stream import IndependentProofStream, SignatureProofStream           # noqa: F401
from .fri import FRI, FriLayers                                                  # noqa: F401
from .air import air_combination                                                 # noqa: F401
from .stark import Stark                                                         # noqa: F401

"""ctypes loader for lib/libzkb200.so.  There is NO fallback: if the CUDA library is
missing or has no device to run on, every product call fails loudly."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libzkb200.so")
CSRC = os.path.join(_HERE, "csrc")
_lib = None

c_u8p = ctypes.POINTER(ctypes.c_uint8)
c_u64p = ctypes.POINTER(ctypes.c_uint64)
vp = ctypes.c_void_p
sz = ctypes.c_size_t
u64 = ctypes.c_uint64


class FriParams(ctypes.Structure):
    _fields_ = [("offset", ctypes.c_uint8 * 16), ("omega", ctypes.c_uint8 * 16),
                ("domain_length", u64), ("expansion_factor", u64), ("num_colinearity_tests", u64)]


class AirDesc(ctypes.Structure):
    _fields_ = [("offset", ctypes.c_uint8 * 16), ("omega", ctypes.c_uint8 * 16),
                ("domain_length", u64), ("expansion_factor", u64),
                ("num_registers", ctypes.c_uint32), ("num_constraints", ctypes.c_uint32),
                ("term_counts", ctypes.POINTER(ctypes.c_uint32)), ("coefs", c_u8p), ("exps", ctypes.POINTER(ctypes.c_uint32)),
                ("boundary_zerofiers", ctypes.POINTER(vp)), ("boundary_zerofier_lens", ctypes.POINTER(sz)),
                ("boundary_interpolants", ctypes.POINTER(vp)), ("boundary_interpolant_lens", ctypes.POINTER(sz)),
                ("transition_zerofier", vp), ("transition_zerofier_len", sz),
                ("weights", c_u8p), ("shifts", c_u64p)]


class AirShape(ctypes.Structure):
    _fields_ = [("offset", ctypes.c_uint8 * 16), ("omega", ctypes.c_uint8 * 16),
                ("domain_length", u64), ("expansion_factor", u64),
                ("num_registers", ctypes.c_uint32), ("num_constraints", ctypes.c_uint32),
                ("term_counts", ctypes.POINTER(ctypes.c_uint32)), ("coefs", c_u8p), ("exps", ctypes.POINTER(ctypes.c_uint32)),
                ("boundary_zerofiers", ctypes.POINTER(vp)), ("boundary_zerofier_lens", ctypes.POINTER(sz)),
                ("transition_zerofier", vp), ("transition_zerofier_len", sz), ("shifts", c_u64p)]


class StarkShape(ctypes.Structure):
    _fields_ = [("omicron", ctypes.c_uint8 * 16), ("omicron_order", u64), ("trace_length", u64), ("num_randomizers", u64),
                ("rnd_poly_len", u64), ("num_registers", ctypes.c_uint32), ("num_constraints", ctypes.c_uint32),
                ("tq_degree_bounds", ctypes.POINTER(ctypes.c_int64)), ("num_boundary", ctypes.c_uint32),
                ("boundary_register", ctypes.POINTER(ctypes.c_uint32)), ("lagrange", vp), ("fri", FriParams), ("proof_bytes", u64)]


FS_CALLBACK = ctypes.CFUNCTYPE(ctypes.c_int, vp, ctypes.c_uint32, c_u8p, ctypes.c_int, c_u8p)

# name -> (restype, argtypes); mirrors include/zkb200.h one to one
PROTOTYPES = {
    "zkb_ctx_create": (ctypes.c_int, [ctypes.c_int, vp, ctypes.POINTER(vp)]),
    "zkb_ctx_destroy": (None, [vp]),
    "zkb_last_error": (ctypes.c_char_p, [vp]),
    "zkb_ctx_sync": (ctypes.c_int, [vp]),
    "zkb_ctx_launches": (u64, [vp]),
    "zkb_version": (ctypes.c_char_p, []),
    "zkb_ctx_zero_copy_inputs": (ctypes.c_int, [vp, ctypes.c_int]),
    "zkb_ctx_assembly_threads": (ctypes.c_int, [vp, ctypes.c_int]),
    "zkb_ctx_tail_threads": (ctypes.c_int, [vp, ctypes.c_int]),
    "zkb_ctx_blocking_sync": (ctypes.c_int, [vp, ctypes.c_int]),
    "zkb_ctx_profile": (ctypes.c_int, [vp, ctypes.c_int]),
    "zkb_ctx_profile_read": (ctypes.c_int, [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_double), c_u64p]),
    "zkb_kernel_name": (ctypes.c_char_p, [ctypes.c_int]),
    "zkb_dev_alloc": (ctypes.c_int, [vp, sz, ctypes.POINTER(vp)]),
    "zkb_dev_free": (ctypes.c_int, [vp, vp]),
    "zkb_memcpy": (ctypes.c_int, [vp, vp, vp, sz]),
    "zkb_primitive_nth_root": (ctypes.c_int, [u64, c_u8p]),
    "zkb_field_generator": (None, [c_u8p]),
    "zkb_field_mul": (None, [c_u8p, c_u8p, c_u8p]),
    "zkb_field_inv": (None, [c_u8p, c_u8p]),
    "zkb_field_pow": (None, [c_u8p, u64, c_u8p]),
    "zkb_field_sample": (None, [c_u8p, sz, c_u8p]),
    "zkb_ntt": (ctypes.c_int, [vp, c_u8p, vp, sz, vp]),
    "zkb_intt": (ctypes.c_int, [vp, c_u8p, vp, sz, vp]),
    "zkb_ntt_batch": (ctypes.c_int, [vp, c_u8p, ctypes.c_int, vp, sz, sz, vp, sz, sz]),
    "zkb_ntt_strided": (ctypes.c_int, [vp, c_u8p, ctypes.c_int, vp, sz, sz, sz, vp]),
    "zkb_ntt4_create": (ctypes.c_int, [vp, ctypes.c_uint32, ctypes.c_uint32, sz, ctypes.POINTER(vp)]),
    "zkb_ntt4_free": (None, [vp]),
    "zkb_ntt4_connect_local": (ctypes.c_int, [ctypes.POINTER(vp), sz]),
    "zkb_ntt4_export": (ctypes.c_int, [vp, c_u8p]),
    "zkb_ntt4_connect_ipc": (ctypes.c_int, [vp, c_u8p]),
    "zkb_ntt4_scatter": (ctypes.c_int, [vp, c_u8p, ctypes.c_int, vp]),
    "zkb_ntt4_finish": (ctypes.c_int, [vp, vp]),
    "zkb_ntt4_run": (ctypes.c_int, [ctypes.POINTER(vp), sz, c_u8p, ctypes.c_int, ctypes.POINTER(vp), ctypes.POINTER(vp)]),
    "zkb_ntt_4step": (ctypes.c_int, [ctypes.POINTER(vp), sz, c_u8p, ctypes.c_int, ctypes.POINTER(vp), sz, ctypes.POINTER(vp)]),
    "zkb_lde_commit_batch": (ctypes.c_int, [ctypes.POINTER(vp), sz, ctypes.POINTER(FriParams), ctypes.POINTER(vp), sz, sz, c_u8p]),
    "zkb_poly_scale": (ctypes.c_int, [vp, c_u8p, vp, sz, vp]),
    "zkb_coset_lde": (ctypes.c_int, [vp, c_u8p, u64, c_u8p, vp, sz, vp]),
    "zkb_coset_lde_batch": (ctypes.c_int, [vp, c_u8p, u64, c_u8p, vp, sz, sz, vp, sz, sz]),
    "zkb_poly_mul": (ctypes.c_int, [vp, c_u8p, u64, vp, sz, vp, sz, vp, ctypes.POINTER(sz)]),
    "zkb_coset_div": (ctypes.c_int, [vp, c_u8p, u64, c_u8p, vp, sz, vp, sz, vp, ctypes.POINTER(sz)]),
    "zkb_merkle_commit": (ctypes.c_int, [vp, vp, sz, c_u8p]),
    "zkb_merkle_build": (ctypes.c_int, [vp, vp, sz, ctypes.POINTER(vp)]),
    "zkb_merkle_root": (ctypes.c_int, [vp, c_u8p]),
    "zkb_merkle_open": (ctypes.c_int, [vp, c_u64p, sz, c_u8p]),
    "zkb_merkle_open_ps": (ctypes.c_int, [vp, c_u64p, sz, vp]),
    "zkb_merkle_build_batch": (ctypes.c_int, [vp, vp, sz, sz, sz, ctypes.POINTER(vp), ctypes.POINTER(vp)]),
    "zkb_fri_prove_batch": (ctypes.c_int, [vp, ctypes.POINTER(FriParams), vp, sz, sz, sz, ctypes.POINTER(vp), c_u64p]),
    "zkb_merkle_open_ps_batch": (ctypes.c_int, [ctypes.POINTER(vp), sz, c_u64p, sz, ctypes.POINTER(vp)]),
    "zkb_merkle_free": (None, [vp]),
    "zkb_merkle_verify": (ctypes.c_int, [c_u8p, u64, c_u8p, sz, c_u8p]),
    "zkb_blake2b512": (None, [c_u8p, sz, c_u8p]),
    "zkb_shake256": (None, [c_u8p, sz, c_u8p, sz]),
    "zkb_shake256_device": (ctypes.c_int, [vp, c_u8p, sz, c_u8p, sz]),
    "zkb_fri_num_rounds": (u64, [ctypes.POINTER(FriParams)]),
    "zkb_fri_fold": (ctypes.c_int, [vp, vp, sz, c_u8p, c_u8p, c_u8p, vp]),
    "zkb_fri_commit": (ctypes.c_int, [vp, ctypes.POINTER(FriParams), vp, sz, FS_CALLBACK, vp, ctypes.POINTER(vp)]),
    "zkb_lde_fri_commit": (ctypes.c_int, [vp, ctypes.POINTER(FriParams), vp, sz, FS_CALLBACK, vp, ctypes.POINTER(vp)]),
    "zkb_fri_commit_ps": (ctypes.c_int, [vp, ctypes.POINTER(FriParams), vp, sz, vp, ctypes.POINTER(vp)]),
    "zkb_lde_fri_commit_ps": (ctypes.c_int, [vp, ctypes.POINTER(FriParams), vp, sz, vp, ctypes.POINTER(vp)]),
    "zkb_fri_layer_count": (u64, [vp]),
    "zkb_fri_layer_len": (u64, [vp, u64]),
    "zkb_fri_layer_root": (ctypes.c_int, [vp, u64, c_u8p]),
    "zkb_fri_layer_codeword": (ctypes.c_int, [vp, u64, vp]),
    "zkb_fri_layer_device_ptr": (vp, [vp, u64]),
    "zkb_fri_query": (ctypes.c_int, [vp, u64, c_u64p, sz, c_u8p, c_u8p]),
    "zkb_fri_layers_free": (None, [vp]),
    "zkb_fri_sample_indices": (ctypes.c_int, [c_u8p, sz, u64, u64, u64, c_u64p]),
    "zkb_ps_create": (ctypes.c_int, [c_u8p, sz, ctypes.c_int, ctypes.POINTER(vp)]),
    "zkb_ps_free": (None, [vp]),
    "zkb_ps_push_root": (ctypes.c_int, [vp, c_u8p, sz]),
    "zkb_ps_push_codeword": (ctypes.c_int, [vp, vp, sz]),
    "zkb_ps_push_path": (ctypes.c_int, [vp, c_u8p, sz]),
    "zkb_ps_push_leafs": (ctypes.c_int, [vp, c_u8p, c_u8p, c_u8p]),
    "zkb_ps_push_value": (ctypes.c_int, [vp, c_u8p]),
    "zkb_ps_push_object": (ctypes.c_int, [vp, ctypes.c_uint8, c_u8p, sz]),
    "zkb_ps_digest": (sz, [vp, c_u8p, sz]),
    "zkb_ps_fiat_shamir": (ctypes.c_int, [vp, sz, c_u8p]),
    "zkb_fri_prove": (ctypes.c_int, [vp, ctypes.POINTER(FriParams), vp, sz, vp, c_u64p]),
    "zkb_air_combination": (ctypes.c_int, [vp, ctypes.POINTER(AirDesc), vp, sz, vp, vp, vp]),
    "zkb_air_create": (ctypes.c_int, [vp, ctypes.POINTER(AirShape), ctypes.POINTER(vp)]),
    "zkb_air_free": (None, [vp]),
    "zkb_air_set_interpolants": (ctypes.c_int, [vp, sz, vp, sz]),
    "zkb_air_boundary_quotients": (ctypes.c_int, [vp, sz, vp, sz, sz, vp, sz, sz]),
    "zkb_air_combine": (ctypes.c_int, [vp, sz, c_u8p, vp, sz, sz, vp, sz, vp, sz, vp, sz]),
    "zkb_trace_lde_batch": (ctypes.c_int, [vp, c_u8p, u64, u64, c_u8p, u64, c_u8p, vp, sz, sz, vp, sz, vp]),
    "zkb_coset_degree_batch": (ctypes.c_int, [vp, c_u8p, vp, sz, sz, sz, ctypes.POINTER(ctypes.c_int64)]),
    "zkb_stark_prove_batch": (ctypes.c_int, [vp, vp, ctypes.POINTER(StarkShape), sz, vp, vp, vp, ctypes.POINTER(vp), c_u64p]),
}


def build(verbose=False):
    """Compile csrc/ into lib/libzkb200.so with nvcc for sm_100a (cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-C", CSRC, "-j8"] + ([] if verbose else ["-s"]))
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "libzkb200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU fallback." % LIB_PATH)
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)          # AttributeError here == header/library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib

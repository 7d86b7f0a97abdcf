"""ctypes binding of libb200pos.so (include/b200pos.h).  There is no CPU fallback: if the shared
library is missing, or no sm_100 GPU is visible, importing callers get a loud error."""
import ctypes
import os
from ctypes import POINTER, c_char_p, c_int, c_int32, c_uint8, c_uint32, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# B2_LIB: load another build of the SAME library (tools/sanitize.sh: the ASan/UBSan-instrumented host side); not a fallback
LIB_PATH = os.environ.get("B2_LIB") or os.path.join(HERE, "libb200pos.so")

B2_OK, B2_EINVAL, B2_ECUDA, B2_ENODEVICE, B2_ENOMEM = 0, -1, -2, -3, -4
_ERRNAMES = {B2_EINVAL: "B2_EINVAL", B2_ECUDA: "B2_ECUDA", B2_ENODEVICE: "B2_ENODEVICE", B2_ENOMEM: "B2_ENOMEM"}


class B2Error(RuntimeError):
    def __init__(self, code, text):
        super().__init__("%s: %s" % (_ERRNAMES.get(code, code), text))
        self.code = code


_lib = None

u8p, u32p, u64p, i32p = POINTER(c_uint8), POINTER(c_uint32), POINTER(c_uint64), POINTER(c_int32)

# name -> (restype, argtypes); every symbol include/b200pos.h declares
SIGNATURES = {
    "b2_init": (c_int, [c_int, POINTER(c_void_p)]),
    "b2_destroy": (None, [c_void_p]),
    "b2_last_error": (c_char_p, [c_void_p]),
    "b2_sync": (c_int, [c_void_p]),
    "b2_launch_count": (c_uint64, [c_void_p]),
    "b2_registry_load": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p]),
    "b2_key_validate": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p]),
    "b2_registry_update_balances": (c_int, [c_void_p, c_void_p, c_void_p, c_uint64]),
    "b2_g1_aggregate": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_uint32, c_void_p, c_void_p]),
    "b2_aggregate": (c_int, [c_void_p, c_void_p, c_void_p, c_uint32, c_void_p, c_void_p]),
    "b2_fast_aggregate_verify": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p, c_void_p, c_uint32, c_void_p]),
    "b2_fast_aggregate_verify_pks": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p]),
    "b2_sk_to_pk": (c_int, [c_void_p, c_void_p, c_uint64, c_void_p]),
    "b2_sign": (c_int, [c_void_p, c_void_p, c_void_p, c_uint64, c_void_p, c_uint32, c_void_p]),
    "b2_hash_to_g2": (c_int, [c_void_p, c_void_p, c_uint32, c_void_p]),
    "b2_sha256_batch": (c_int, [c_void_p, c_void_p, c_uint32, c_uint64, c_void_p]),
    "b2_signing_roots": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_uint32, c_void_p]),
    "b2_shuffle_committees": (c_int, [c_void_p, c_void_p, c_void_p, c_uint32, c_uint32, c_void_p]),
    "b2_shuffle_committees_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_uint32, c_uint32, c_void_p, c_void_p]),
    "b2_latest_messages_reset": (c_int, [c_void_p]),
    "b2_latest_messages_load": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64]),
    "b2_latest_messages_read": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint64]),
    "b2_latest_messages_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p, c_void_p, c_void_p, c_uint32]),
    "b2_participation_load": (c_int, [c_void_p, c_int, c_void_p, c_uint64]),
    "b2_participation_read": (c_int, [c_void_p, c_int, c_void_p, c_uint64]),
    "b2_participation_update": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p, c_void_p, c_uint32, c_uint64, c_uint64, c_void_p]),
    "b2_ffg_balances": (c_int, [c_void_p, c_uint32, c_void_p]),
    "b2_attestations_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_uint32, c_uint32, c_uint32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2_set_fork_choice_params": (c_int, [c_void_p, c_uint64, c_int]),
    "b2_set_verify_mode": (c_int, [c_void_p, c_int, c_void_p]),
    "b2_epoch_set_pairing_form": (c_int, [c_void_p, c_int]),
    "b2_on_attester_slashing": (c_int, [c_void_p, c_void_p, c_uint32, c_void_p, c_uint32]),
    "b2_tree_load": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint32]),
    "b2_get_weights": (c_int, [c_void_p, c_int32, c_uint64, c_void_p]),
    "b2_get_head": (c_int, [c_void_p, c_uint32, c_int32, c_uint64, c_void_p]),
    "b2_aggregate_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_uint32, c_uint64, c_void_p, c_void_p, c_void_p]),
    "b2_fast_aggregate_verify_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p, c_void_p, c_uint32, c_void_p, c_void_p]),
    "b2_latest_messages_update_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p]),
    "b2_epoch_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p, c_void_p, c_void_p, c_uint32, c_uint64,
                             c_void_p, c_void_p, c_void_p, c_void_p]),
    "b2_epoch_start_dev": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p, c_uint32, c_uint64, c_void_p, c_void_p]),
    "b2_epoch_tail_dev": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_uint32, c_void_p, c_void_p, c_uint32, c_void_p, c_void_p, c_void_p]),
    "b2_epoch_wait_dev": (c_int, [c_void_p, c_int, c_void_p]),
    "b2_vote_weights_dev": (c_int, [c_void_p, c_void_p, c_void_p]),
    "b2_vote_weights_range_dev": (c_int, [c_void_p, c_uint64, c_uint64, c_void_p, c_void_p]),
    "b2_guard_flags": (c_int, [c_void_p, c_void_p]),
    "b2_gather_probe_dev": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_uint32, c_uint32, c_int, c_void_p, c_void_p]),
    "b2_head_from_votes_dev": (c_int, [c_void_p, c_void_p, c_uint32, c_int32, c_uint64, c_void_p, c_void_p, c_void_p]),
    "b2_tree_size": (c_uint32, [c_void_p]),
    "b2_fc_exchange_export": (c_int, [c_void_p, c_void_p]),
    "b2_fc_exchange_open": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "b2_get_head_multi": (c_int, [c_void_p, c_uint64, c_uint64, c_uint32, c_int32, c_uint64, c_void_p]),
    "b2_debug_head_clocks": (c_int, [c_void_p, c_void_p]),
}


def load():
    """Load the shared library (once) and declare every prototype.  Needs no GPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `python -m pos_evolution_b200.build` "
                          "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib

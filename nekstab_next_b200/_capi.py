"""ctypes declarations for include/nekstab_b200.h.  No torch types cross this boundary."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / 'libnekstab_b200.so'

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)
c_i64_p = C.POINTER(C.c_int64)
c_u64_p = C.POINTER(C.c_uint64)
c_void_pp = C.POINTER(C.c_void_p)
c_dpp = C.POINTER(c_double_p)
H = C.c_void_p  # opaque handle

HOST_MATVEC = C.CFUNCTYPE(C.c_int, C.c_void_p, c_dpp, C.c_double, c_dpp, c_double_p)

# name -> (restype, argtypes); every symbol the header declares
PROTOTYPES = {
    'nsb_last_error': (C.c_char_p, []),
    'nsb_version': (C.c_int, []),
    'nsb_build_id': (C.c_char_p, []),
    'nsb_get_unique_id': (C.c_int, [C.c_void_p]),
    'nsb_init': (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_void_p, c_void_pp]),
    'nsb_finalize': (C.c_int, [H]),
    'nsb_p2p_mailbox_create': (C.c_int, [H, C.c_int64, C.c_void_p]),
    'nsb_p2p_mailbox_connect': (C.c_int, [H, C.c_void_p]),
    'nsb_p2p_enabled': (C.c_int, [H, c_int_p]),
    'nsb_sync': (C.c_int, [H]),
    'nsb_rank': (C.c_int, [H, c_int_p, c_int_p]),
    'nsb_stream': (C.c_int, [H, c_u64_p]),
    'nsb_launch_count': (C.c_int, [H, c_i64_p]),
    'nsb_timer_start': (C.c_int, [H]),
    'nsb_timer_stop': (C.c_int, [H, c_double_p]),
    'nsb_prof_enable': (C.c_int, [H, C.c_int]),
    'nsb_prof_get': (C.c_int, [H, C.c_int, c_double_p, c_i64_p, c_double_p]),
    'nsb_profiler_start': (C.c_int, []),
    'nsb_profiler_stop': (C.c_int, []),
    'nsb_allreduce_host': (C.c_int, [H, c_double_p, C.c_int]),
    'nsb_flush_l2': (C.c_int, [H]),
    'nsb_layout_create': (C.c_int, [H, C.c_int, c_i64_p, c_int_p, C.c_int, c_void_pp]),
    'nsb_layout_create_c0': (C.c_int, [H, H, C.c_int, c_i64_p, c_int_p, C.c_int, C.c_int, c_void_pp]),
    'nsb_layout_is_c0': (C.c_int, [H, c_int_p, c_i64_p]),
    'nsb_layout_destroy': (C.c_int, [H]),
    'nsb_layout_info': (C.c_int, [H, c_i64_p, c_i64_p, c_i64_p]),
    'nsb_layout_set_weight': (C.c_int, [H, c_dpp]),
    'nsb_basis_create': (C.c_int, [H, C.c_int, c_void_pp]),
    'nsb_basis_destroy': (C.c_int, [H]),
    'nsb_basis_ncols': (C.c_int, [H, c_int_p]),
    'nsb_basis_col_ptr': (C.c_int, [H, C.c_int, c_u64_p]),
    'nsb_vec_upload': (C.c_int, [H, C.c_int, c_dpp, C.c_double]),
    'nsb_vec_download': (C.c_int, [H, C.c_int, c_dpp, c_double_p]),
    'nsb_vec_zero': (C.c_int, [H, C.c_int]),
    'nsb_vec_copy': (C.c_int, [H, C.c_int, H, C.c_int]),
    'nsb_vec_scal': (C.c_int, [H, C.c_int, C.c_double]),
    'nsb_vec_axpby': (C.c_int, [H, C.c_int, C.c_double, H, C.c_int, C.c_double, C.c_int]),
    'nsb_vec_add2': (C.c_int, [H, C.c_int, H, C.c_int]),
    'nsb_vec_sub2': (C.c_int, [H, C.c_int, H, C.c_int]),
    'nsb_vec_sub3': (C.c_int, [H, C.c_int, H, C.c_int, H, C.c_int]),
    'nsb_vec_dot': (C.c_int, [H, C.c_int, H, C.c_int, c_double_p]),
    'nsb_vec_norm': (C.c_int, [H, C.c_int, c_double_p]),
    'nsb_vec_normalize': (C.c_int, [H, C.c_int, c_double_p]),
    'nsb_set_dgks_eta': (C.c_int, [H, C.c_double]),
    'nsb_orthonormalize': (C.c_int, [H, C.c_int, C.c_int, C.c_int, c_double_p, c_int_p]),
    'nsb_orthonormalize_async': (C.c_int, [H, C.c_int, C.c_int, C.c_int, c_double_p]),
    'nsb_host_alloc': (C.c_int, [c_void_pp, C.c_int64]),
    'nsb_host_free': (C.c_int, [C.c_void_p]),
    'nsb_basis_gram': (C.c_int, [H, C.c_int, c_double_p, C.c_int]),
    'nsb_basis_qr': (C.c_int, [H, C.c_int, C.c_int, c_double_p, C.c_int]),
    'nsb_basis_gemv': (C.c_int, [H, C.c_int, c_double_p, H, C.c_int]),
    'nsb_basis_rotate': (C.c_int, [H, C.c_int, c_double_p, C.c_int, C.c_int]),
    'nsb_gll': (C.c_int, [C.c_int, c_double_p, c_double_p, c_double_p]),
    'nsb_sem_create': (C.c_int, [H, C.c_int, C.c_int, C.c_int64, c_double_p, c_double_p, c_double_p,
                                 c_double_p, c_i64_p, c_void_pp]),
    'nsb_sem_destroy': (C.c_int, [H]),
    'nsb_sem_get': (C.c_int, [H, C.c_int, c_double_p]),
    'nsb_sem_npts': (C.c_int64, [H]),
    'nsb_sem_setup_exchange': (C.c_int, [H]),
    'nsb_host_gs_plan': (C.c_int, [C.c_int, C.c_int, C.c_int64, c_i64_p, c_i64_p, c_i64_p, c_i64_p,
                                   C.POINTER(C.c_int32), c_i64_p]),
    'nsb_host_exchange_plan': (C.c_int, [C.c_int, C.c_int, C.c_int64, c_i64_p, c_i64_p, c_i64_p, C.c_int64,
                                         c_i64_p, c_i64_p, c_i64_p, C.POINTER(C.c_int32)]),
    'nsb_sem_axhelm': (C.c_int, [H, H, C.c_int, H, C.c_int, C.c_int, C.c_double, C.c_double]),
    'nsb_sem_dssum': (C.c_int, [H, H, C.c_int, C.c_int]),
    'nsb_sem_col2': (C.c_int, [H, H, C.c_int, C.c_int, C.c_int]),
    'nsb_sem_ax': (C.c_int, [H, H, C.c_int, H, C.c_int, C.c_int, C.c_double, C.c_double]),
    'nsb_sem_hmholtz': (C.c_int, [H, H, C.c_int, H, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                  c_int_p, c_double_p]),
    'nsb_sem_hmholtz_vec': (C.c_int, [H, H, C.c_int, H, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                      C.c_int, c_int_p, c_double_p]),
    'nsb_dealias_matrices': (C.c_int, [C.c_int, C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    'nsb_sem_dealias_setup': (C.c_int, [H, C.c_int]),
    'nsb_sem_set_convect': (C.c_int, [H, C.c_int, H, C.c_int, C.c_int]),
    'nsb_sem_convect': (C.c_int, [H, C.c_int, H, C.c_int, H, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]),
    'nsb_sem_convect_t': (C.c_int, [H, C.c_int, H, C.c_int, H, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]),
    'nsb_sem_bdf_ext': (C.c_int, [H, H, C.c_int, C.c_int, C.c_int, c_int_p, C.c_int, C.c_int, C.c_int, c_double_p,
                                  c_double_p, C.c_double]),
    'nsb_op_create_sem': (C.c_int, [H, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                    c_double_p, c_double_p, c_double_p, c_void_pp]),
    'nsb_op_create_host': (C.c_int, [H, HOST_MATVEC, C.c_void_p, c_void_pp]),
    'nsb_op_set_linear': (C.c_int, [H, C.c_int]),
    'nsb_op_create_compose': (C.c_int, [H, H, H, c_void_pp]),
    'nsb_op_create_axpby': (C.c_int, [H, H, H, C.c_double, C.c_double, c_void_pp]),
    'nsb_op_create_frechet_fd': (C.c_int, [H, H, H, C.c_int, C.c_int, c_void_pp]),
    'nsb_op_frechet_set_epsilon': (C.c_int, [H, C.c_double]),
    'nsb_op_create_stepper': (C.c_int, [H, H, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                        C.c_double, C.c_int, c_void_pp]),
    'nsb_op_create_stepper_adjoint': (C.c_int, [H, H, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int,
                                                C.c_double, C.c_int, c_void_pp]),
    'nsb_pressure_matrices': (C.c_int, [C.c_int, c_double_p, c_double_p, c_double_p, c_double_p]),
    'nsb_fdm_matrices': (C.c_int, [C.c_int, c_double_p, c_double_p]),
    'nsb_sem_pressure_setup': (C.c_int, [H]),
    'nsb_sem_npres': (C.c_int64, [H]),
    'nsb_sem_pressure_get': (C.c_int, [H, C.c_int, c_double_p]),
    'nsb_sem_opdiv': (C.c_int, [H, H, C.c_int, H, C.c_int]),
    'nsb_sem_opgradt': (C.c_int, [H, H, C.c_int, H, C.c_int]),
    'nsb_sem_cdabdtp': (C.c_int, [H, H, C.c_int, H, C.c_int]),
    'nsb_sem_esolve': (C.c_int, [H, H, C.c_int, H, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, c_int_p,
                                 c_double_p]),
    'nsb_sem_norm_grad': (C.c_int, [H, H, C.c_int, c_double_p]),
    'nsb_sem_cfl': (C.c_int, [H, H, C.c_int, C.c_double, c_double_p]),
    'nsb_op_create_ns_stepper': (C.c_int, [H, H, H, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double,
                                           C.c_int, C.c_int, C.c_int, c_void_pp]),
    'nsb_op_create_ns_stepper_adjoint': (C.c_int, [H, H, H, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double,
                                                   C.c_double, C.c_int, C.c_int, C.c_int, c_void_pp]),
    'nsb_op_ns_set_orbit': (C.c_int, [H, H, C.c_int, C.c_int]),
    'nsb_op_ns_iterations': (C.c_int, [H, c_i64_p, c_i64_p]),
    'nsb_op_destroy': (C.c_int, [H]),
    'nsb_op_apply': (C.c_int, [H, H, C.c_int, H, C.c_int]),
    'nsb_op_count': (C.c_int, [H, c_i64_p]),
    'nsb_arnoldi': (C.c_int, [H, H, C.c_int, C.c_int, C.c_int, c_double_p, C.c_int]),
    'nsb_arnoldi_passes': (C.c_int, [H, C.c_int, C.c_int, C.c_int, c_int_p]),
    'nsb_set_lapack': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'nsb_set_lapack_svd': (C.c_int, [C.c_void_p]),
    'nsb_svd': (C.c_int, [c_double_p, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p, c_double_p]),
    'nsb_eig': (C.c_int, [c_double_p, C.c_int, C.c_int, c_double_p, c_double_p]),
    'nsb_schur': (C.c_int, [c_double_p, C.c_int, C.c_int, c_double_p, c_double_p]),
    'nsb_ordschur': (C.c_int, [c_double_p, C.c_int, c_double_p, C.c_int, c_int_p, C.c_int]),
    'nsb_lstsq': (C.c_int, [c_double_p, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p]),
    'nsb_select_eigenvalues': (C.c_int, [c_int_p, c_int_p, c_double_p, C.c_double, C.c_int, C.c_int]),
    'nsb_schur_condensation': (C.c_int, [H, c_int_p, c_double_p, C.c_int, C.c_int, C.c_double, C.c_int]),
    'nsb_krylov_schur': (C.c_int, [H, H, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                   c_double_p, C.c_int, c_double_p, c_double_p, c_double_p, c_int_p,
                                   c_int_p]),
    'nsb_eigs': (C.c_int, [H, H, C.c_int, C.c_int, C.c_double, C.c_int, c_double_p, C.c_int, c_double_p,
                           c_double_p, c_double_p, c_int_p, c_int_p]),
    'nsb_newton_krylov': (C.c_int, [H, H, H, H, C.c_int, H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double,
                                    C.c_int, c_int_p, c_double_p, c_int_p]),
    'nsb_ritz_vector': (C.c_int, [H, C.c_int, c_double_p, H, C.c_int, C.c_int, C.c_int, c_double_p, c_double_p]),
    'nsb_svds': (C.c_int, [H, H, H, H, C.c_int, C.c_int, C.c_double, C.c_int, c_double_p, C.c_int, c_double_p,
                           c_double_p, c_double_p, c_double_p, c_int_p, c_int_p]),
    'nsb_hessenberg_write': (C.c_int, [C.c_char_p, c_double_p, C.c_int, C.c_int]),
    'nsb_hessenberg_read': (C.c_int, [C.c_char_p, C.c_int, C.c_int, c_double_p, C.c_int]),
    'nsb_fld_read_into': (C.c_int, [H, C.c_int, C.c_char_p, c_i64_p, C.c_int64, C.c_int, C.c_int, C.c_int, c_double_p]),
    'nsb_restart_load': (C.c_int, [H, C.c_char_p, C.c_char_p, C.c_int, C.c_int, c_i64_p, C.c_int64, C.c_int, C.c_int,
                                   C.c_int, c_double_p, C.c_int, c_int_p]),
    'nsb_ts_gmres': (C.c_int, [H, H, H, C.c_int, H, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int,
                               c_int_p, c_double_p, c_int_p]),
}

_lib = None


class NsbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f'nekstab_b200 error {code}: {msg}')
        self.code = code


def load():
    """Load libnekstab_b200.so; fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f'{LIB_PATH} is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                          f'or `python -m nekstab_next_b200.build`; there is no CPU fallback')
    lib = C.CDLL(str(LIB_PATH), mode=C.RTLD_GLOBAL)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export the symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code):
    if code != 0:
        msg = load().nsb_last_error()
        raise NsbError(code, msg.decode() if msg else '')
    return code

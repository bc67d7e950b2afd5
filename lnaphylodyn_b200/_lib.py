"""ctypes loader for liblnaphylodyn_b200.so (the C ABI in include/lnaphylodyn_b200.h).

Fails loudly: if the CUDA library is missing it is built with nvcc (in-tree); if that is impossible
an ImportError/RuntimeError is raised.  There is no CPU fallback anywhere in this package.
"""
import ctypes as C
import os

from . import build as _build

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_vp = C.c_void_p


class LnaCfg(C.Structure):
    _fields_ = [("model", C.c_int), ("transX", C.c_int), ("nch", C.c_int), ("nparam", C.c_int),
                ("iR0", C.c_int), ("iGamma", C.c_int), ("gridsize", C.c_int), ("nt", C.c_int),
                ("times", _dp), ("x_r", _dp), ("t_correct", C.c_double)]


class LnaInit(C.Structure):
    _fields_ = [("E", C.c_int), ("ng", C.c_int), ("C", _dp), ("D", _dp), ("y", _dp), ("gridrep", _dp)]


class LnaMcmcOpts(C.Structure):
    _fields_ = [("prior8", C.c_double * 8), ("gamma_prior", C.c_double * 2), ("gamma_prop", C.c_double),
                ("ESS_vec", C.c_int32 * 7), ("update_traj", C.c_int32), ("hyper_noop_mh", C.c_int32),
                ("spec_width", C.c_int32), ("n_mh", C.c_int32), ("mh_kind", C.c_int32 * 4), ("mh_index", C.c_int32 * 4),
                ("mh_is_log", C.c_int32 * 4), ("mh_prior", (C.c_double * 2) * 4), ("mh_prop", C.c_double * 4),
                ("use_ess2", C.c_int32), ("ESS_vec2", C.c_int32 * 7)]


class LnaMcmc2Opts(C.Structure):
    _fields_ = [("n_mh", C.c_int32), ("mh_index", C.c_int32 * 4), ("mh_is_log", C.c_int32 * 4),
                ("mh_prior_kind", C.c_int32 * 4), ("mh_prior", (C.c_double * 2) * 4), ("mh_prop", C.c_double * 4),
                ("update_hyper", C.c_int32), ("hyper_prop", C.c_double), ("hyper_prior", C.c_double * 2),
                ("update_chpt", C.c_int32), ("update_traj", C.c_int32), ("spec_width", C.c_int32),
                ("joint", C.c_int32), ("joint_d", C.c_int32), ("joint_index", C.c_int32 * 4),
                ("joint_is_log", C.c_int32 * 4), ("joint_prior_kind", C.c_int32 * 4), ("joint_prior", (C.c_double * 2) * 4)]


# name -> (restype, argtypes); pointers are passed as void* so that host (numpy) and device (torch
# data_ptr) addresses go through the same declarations.
_I64 = C.c_int64
_SIGS = {
    "lna_abi_version": (C.c_int, []),
    "lna_last_error": (C.c_char_p, []),
    "lna_device_count": (C.c_int, []),
    "lna_shard": (C.c_int, [_I64, C.c_int, C.c_int, C.POINTER(_I64), C.POINTER(_I64)]),
    "lna_problem_create": (_vp, [C.POINTER(LnaCfg), C.POINTER(LnaInit), C.c_int]),
    "lna_problem_destroy": (None, [_vp]),
    "lna_problem_sync": (C.c_int, [_vp]),
    "lna_problem_stream": (_vp, [_vp]),
    "lna_problem_set_stream": (C.c_int, [_vp, _vp]),
    "lna_problem_dims": (C.c_int, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "lna_problem_launch_count": (_I64, [_vp]),
    "lna_problem_cell_sums": (C.c_int, [_vp, _vp, _vp, _vp]),
    "lna_betaTs": (C.c_int, [_vp, _I64, _vp, _vp, C.c_int, _vp]),
    "lna_param_transform": (C.c_int, [_vp, _I64, _vp, _vp, _vp]),
    "lna_ode_rk45": (C.c_int, [_vp, _I64, _vp, _vp, _vp]),
    "lna_kf_param": (C.c_int, [_vp, _I64, _vp, _vp, C.c_int, _vp, _vp, _vp]),
    "lna_ode_coarse_slicer": (C.c_int, [_vp, _I64, _vp, _vp]),
    "lna_new_param_list": (C.c_int, [_vp, _I64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lna_transform_traj": (C.c_int, [_vp, _I64, _vp, _vp, _vp, _vp, _vp]),
    "lna_traj_sim_noncentral": (C.c_int, [_vp, _I64] + [_vp] * 8),
    "lna_volz_loglik_nh2": (C.c_int, [_vp, _I64, _vp, _vp, _vp, _vp]),
    "lna_volz_loglik_nh3": (C.c_int, [_vp, _I64, _vp, _vp, _vp, _vp]),
    "lna_coal_loglik": (C.c_int, [_vp, _I64, _vp, _vp, C.c_int, C.c_int, _vp]),
    "lna_full_loglik": (C.c_int, [_vp, _I64] + [_vp] * 10),
    "lna_full_loglik_dev": (C.c_int, [_vp, _I64] + [_vp] * 10),
    "lna_full_loglik_async": (C.c_int, [_vp, _I64] + [_vp] * 5),
    "lna_full_loglik_shared": (C.c_int, [_vp, _I64, C.c_int] + [_vp] * 5),
    "lna_full_loglik_shared_dev": (C.c_int, [_vp, _I64, C.c_int] + [_vp] * 5),
    "lna_full_loglik_shared_async": (C.c_int, [_vp, _I64, C.c_int] + [_vp] * 5),
    "lna_sirs_loglik": (C.c_int, [_vp, _I64, _vp, _vp, C.c_int, _vp, C.c_int] + [_vp] * 5),
    "lna_sirs_loglik_dev": (C.c_int, [_vp, _I64, _vp, _vp, C.c_int, _vp, C.c_int] + [_vp] * 5),
    "lna_update_param": (C.c_int, [_vp, _I64] + [_vp] * 14),
    "lna_param_slice_update": (C.c_int, [_vp, _I64, _vp, _vp, _vp, _vp]),
    "lna_param_slice_update_all2": (C.c_int, [_vp, _I64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lna_eslice_par_general": (C.c_int, [_vp, _I64] + [_vp] * 7 + [C.c_int] + [_vp] * 10),
    "lna_eslice_par_general_ex": (C.c_int, [_vp, _I64] + [_vp] * 7 + [C.c_int] + [_vp] * 11),
    "lna_eslice_general_nc": (C.c_int, [_vp, _I64] + [_vp] * 14),
    "lna_eslice_general_nc_kingman": (C.c_int, [_vp, _I64] + [_vp] * 15),
    "lna_eslice_change_points": (C.c_int, [_vp, _I64] + [_vp] * 6 + [C.c_int] + [_vp] * 10),
    "lna_log_like_traj_general2": (C.c_int, [_vp, _I64] + [_vp] * 5),
    "lna_log_like_traj_sirs": (C.c_int, [_vp, _I64] + [_vp] * 5),
    "lna_traj_sim_general3": (C.c_int, [_vp, _I64] + [_vp] * 7),
    "lna_traj_sim_sirs": (C.c_int, [_vp, _I64] + [_vp] * 8),
    "lna_eslice_general2": (C.c_int, [_vp, _I64] + [_vp] * 10),
    "lna_log_like_traj_general": (C.c_int, [_vp, _I64] + [_vp] * 5),
    "lna_traj_sim_general": (C.c_int, [_vp, _I64] + [_vp] * 6),
    "lna_volz_loglik_nh": (C.c_int, [_vp, _I64] + [_vp] * 3),
    "lna_eslice_legacy": (C.c_int, [_vp, _I64, C.c_int, C.c_int, C.c_int] + [_vp] * 13),
    "lna_foo": (C.c_int, [_vp, _I64, C.c_int, C.c_int] + [_vp] * 7),
    "lna_volz_thin_sir": (C.c_int, [_vp, _I64, C.c_int, _I64] + [_vp] * 4 + [C.c_int, _vp, _vp, C.c_double, C.c_double,
                                    C.c_uint64, _I64, C.c_uint64, C.c_int] + [_vp] * 5),
    "lna_thin_draws": (C.c_int, [_vp, C.c_uint64, C.c_uint64, C.c_uint64, _I64, _vp, _vp]),
    "lna_chains_create": (_vp, [_vp, _I64]),
    "lna_chains_destroy": (None, [_vp]),
    "lna_chains_init": (C.c_int, [_vp, _vp, _vp]),
    "lna_chains_step": (C.c_int, [_vp, C.POINTER(LnaMcmcOpts), _vp, _vp]),
    "lna_chains_step_dev": (C.c_int, [_vp, C.POINTER(LnaMcmcOpts), _vp, _vp]),
    "lna_chains_step_mcmc2": (C.c_int, [_vp, C.POINTER(LnaMcmc2Opts), _vp, _vp]),
    "lna_chains_step_mcmc2_dev": (C.c_int, [_vp, C.POINTER(LnaMcmc2Opts), _vp, _vp]),
    "lna_chains_set_joint_transform": (C.c_int, [_vp, C.c_int, _vp]),
    "lna_chains_step_mcmc2_rng": (C.c_int, [_vp, C.POINTER(LnaMcmc2Opts), C.c_uint64, C.c_uint64, C.c_int64]),
    "lna_chains_step_rng": (C.c_int, [_vp, C.POINTER(LnaMcmcOpts), C.c_uint64, C.c_uint64, _I64]),
    "lna_draw_streams": (C.c_int, [_vp, _I64, _I64, C.c_int, C.c_int, C.c_uint64, C.c_uint64, _vp, _vp]),
    "lna_chains_get": (C.c_int, [_vp] + [_vp] * 6),
    "lna_chains_summary_dev": (_vp, [_vp]),
    "lna_chains_set_free_running": (C.c_int, [_vp, C.c_int]),
    "lna_chains_join": (C.c_int, [_vp]),
    "lna_chains_sync": (C.c_int, [_vp]),
    "lna_chains_groups": (C.c_int, [_vp]),
    "lna_chains_exec_counts": (C.c_int, [_vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "lna_chains_history_begin": (C.c_int, [_vp, C.c_int64]),
    "lna_chains_history_record": (C.c_int, [_vp, C.c_int64]),
    "lna_chains_history_get": (C.c_int, [_vp, _vp, _vp, _vp]),
    "lna_comm_available": (C.c_int, []),
    "lna_comm_unique_id": (C.c_int, [_vp]),
    "lna_comm_create": (_vp, [C.c_int, C.c_int, _vp, C.c_int]),
    "lna_comm_destroy": (None, [_vp]),
    "lna_comm_rank": (C.c_int, [_vp]),
    "lna_comm_size": (C.c_int, [_vp]),
    "lna_comm_allgather": (C.c_int, [_vp, _vp, _vp, C.POINTER(_I64), _vp]),
    "lna_chains_allgather_summaries": (C.c_int, [_vp, _vp, _I64, _vp]),
    "lna_multi_create": (_vp, [C.POINTER(LnaCfg), C.POINTER(LnaInit), C.c_int, C.POINTER(C.c_int), _I64]),
    "lna_multi_destroy": (None, [_vp]),
    "lna_multi_devices": (C.c_int, [_vp]),
    "lna_multi_chains": (_vp, [_vp, C.c_int]),
    "lna_multi_problem": (_vp, [_vp, C.c_int]),
    "lna_multi_init": (C.c_int, [_vp, _vp, _vp]),
    "lna_multi_set_free_running": (C.c_int, [_vp, C.c_int]),
    "lna_multi_step_rng": (C.c_int, [_vp, C.POINTER(LnaMcmcOpts), C.c_uint64, C.c_uint64, _I64]),
    "lna_multi_step_mcmc2_rng": (C.c_int, [_vp, C.POINTER(LnaMcmc2Opts), C.c_uint64, C.c_uint64, _I64]),
    "lna_multi_set_joint_transform": (C.c_int, [_vp, C.c_int, _vp]),
    "lna_multi_sync": (C.c_int, [_vp]),
    "lna_multi_get": (C.c_int, [_vp] + [_vp] * 6),
    "lna_multi_history_begin": (C.c_int, [_vp, _I64]),
    "lna_multi_history_record": (C.c_int, [_vp, _I64]),
    "lna_multi_history_get": (C.c_int, [_vp, _I64, _vp, _vp, _vp]),
    "lna_multi_gather_summaries": (C.c_int, [_vp, C.c_int, _vp]),
    "lna_multi_gathered_dev": (_vp, [_vp, C.c_int]),
    "lna_measure_fp64_peak": (C.c_double, [C.c_int, C.c_int]),
    "lna_debug_timeline_dump": (C.c_int, [C.c_char_p]),
}

_lib = None


def lib():
    """The loaded CUDA library; builds it in-tree on first use.  Never falls back to CPU code."""
    global _lib
    if _lib is None:
        so = _build.build()
        if not os.path.exists(so):
            raise ImportError(f"{so} is missing and could not be built: lnaphylodyn_b200 has no CPU fallback")
        L = C.CDLL(so)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)  # AttributeError here = the .so does not match the header
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def exported_symbols():
    return sorted(_SIGS)


class LnaError(RuntimeError):
    pass


def check(rc, what=""):
    if rc != 0:
        msg = lib().lna_last_error()
        raise LnaError(f"{what}: {msg.decode() if msg else 'unknown error'}")

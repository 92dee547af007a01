"""ctypes binding of libbgm_b200.so (include/bgm_b200.h).

There is NO fallback: if the library is missing or a call fails, the product path
raises.  torch tensors are used only as device-memory containers; the library sees
raw pointers and a raw cudaStream_t.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbgm_b200.so")

_lib = None


class BgmError(RuntimeError):
    """A C-ABI call returned a negative bgm_status."""

    def __init__(self, fn, code, msg):
        super().__init__("%s failed (%d): %s" % (fn, code, msg))
        self.code = code


class NetDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int), ("dims", C.POINTER(C.c_int)), ("params", C.POINTER(C.c_float))]


class MhArgs(C.Structure):
    _fields_ = [
        ("x_dev", C.c_void_p), ("y_dev", C.c_void_p), ("v_dev", C.c_void_p),
        ("ldv", C.c_int), ("n", C.c_int),
        ("vproj_dev", C.c_void_p), ("r0_dev", C.c_void_p), ("ldvproj", C.c_int), ("sched_dev", C.c_void_p),
        ("z_state_dev", C.c_void_p), ("lp_state_dev", C.c_void_p),
        ("init_mode", C.c_int), ("t_begin", C.c_int), ("t_end", C.c_int), ("burn_in", C.c_int),
        ("q_sd_dev", C.c_void_p), ("eps_dev", C.c_void_p), ("u_dev", C.c_void_p),
        ("seed", C.c_uint64), ("row_offset", C.c_int64),
        ("out_samples_dev", C.c_void_p), ("accept_count_dev", C.c_void_p),
        ("accept_mask_dev", C.c_void_p), ("lp_trace_dev", C.c_void_p),
        ("prior_dev", C.c_void_p), ("ldprior", C.c_int),
    ]


class MtState(C.Structure):
    _fields_ = [("key", C.c_uint32 * 624), ("pos", C.c_int), ("has_gauss", C.c_int), ("gauss", C.c_double)]


class BnnNetDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int), ("dims", C.POINTER(C.c_int)), ("bn", C.POINTER(C.c_float)),
                ("params", C.POINTER(C.c_float))]


class VarNetDesc(C.Structure):
    _fields_ = [("z_dim", C.c_int), ("x_dim", C.c_int), ("n_hidden", C.c_int),
                ("units", C.POINTER(C.c_int)), ("bn", C.POINTER(C.c_float)),
                ("hidden_params", C.POINTER(C.c_float)), ("mean_params", C.POINTER(C.c_float)),
                ("var_params", C.POINTER(C.c_float))]


class DiscDesc(C.Structure):
    _fields_ = [("n_hidden", C.c_int), ("dims", C.POINTER(C.c_int)), ("params", C.POINTER(C.c_float))]


class HmcArgs(C.Structure):
    _fields_ = [
        ("x_dev", C.c_void_p), ("ldx", C.c_int), ("n", C.c_int),
        ("z_state_dev", C.c_void_p), ("g_state_dev", C.c_void_p), ("lp_state_dev", C.c_void_p),
        ("init_mode", C.c_int), ("t_begin", C.c_int), ("t_end", C.c_int), ("burn_in", C.c_int),
        ("num_leapfrog", C.c_int),
        ("step_dev", C.c_void_p), ("mom_dev", C.c_void_p), ("logu_dev", C.c_void_p),
        ("seed", C.c_uint64), ("row_offset", C.c_int64),
        ("out_samples_dev", C.c_void_p), ("accept_stat_dev", C.c_void_p), ("accept_count_dev", C.c_void_p),
        ("accept_mask_dev", C.c_void_p), ("log_accept_dev", C.c_void_p),
    ]


# name -> (restype, argtypes); every symbol include/bgm_b200.h declares
SYMBOLS = {
    "bgm_last_error": (C.c_char_p, []),
    "bgm_version": (C.c_int, []),
    "bgm_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "bgm_causal_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int,
                                    C.c_float, C.c_float, C.c_float,
                                    C.POINTER(NetDesc), C.POINTER(NetDesc), C.POINTER(NetDesc)]),
    "bgm_causal_destroy": (None, [C.c_void_p]),
    "bgm_causal_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                  C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.POINTER(C.c_int)]),
    "bgm_causal_set_sampler": (C.c_int, [C.c_void_p, C.c_int]),
    "bgm_causal_sampler_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                          C.POINTER(C.c_longlong)]),
    "bgm_causal_kernel_name": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
    "bgm_causal_project": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_void_p]),
    "bgm_causal_logpost": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "bgm_causal_logpost_cond": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgm_causal_prior_rows": (C.c_int, [C.POINTER(NetDesc), C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                        C.c_void_p]),
    "bgm_causal_mh": (C.c_int, [C.c_void_p, C.POINTER(MhArgs), C.c_void_p]),
    "bgm_mh_adapt_qsd": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_longlong, C.c_double, C.c_double,
                                   C.c_void_p, C.c_void_p]),
    "bgm_mh_noise": (C.c_int, [C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgm_causal_effect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                    C.c_int, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "bgm_causal_effect_index": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "bgm_causal_effect_compact": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p]),
    "bgm_causal_effect_heads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                          C.c_void_p]),
    "bgm_causal_effect_combine": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_uint64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p]),
    "bgm_fp32_peak_tflops": (C.c_int, [C.POINTER(C.c_double), C.c_void_p]),
    "bgm_trainer_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int,
                                     C.POINTER(NetDesc), C.POINTER(NetDesc), C.POINTER(NetDesc),
                                     C.POINTER(NetDesc), C.POINTER(DiscDesc), C.c_float, C.c_float, C.c_float]),
    "bgm_trainer_destroy": (None, [C.c_void_p]),
    "bgm_trainer_buffers": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p),
                                      C.POINTER(C.c_void_p)]),
    "bgm_trainer_get_params": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "bgm_trainer_set_params": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "bgm_train_disc_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float,
                                      C.c_void_p, C.c_void_p]),
    "bgm_train_gen_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_void_p]),
    "bgm_train_adam": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_void_p]),
    "bgm_bgmtrainer_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(VarNetDesc), C.POINTER(NetDesc),
                                        C.POINTER(DiscDesc), C.POINTER(DiscDesc), C.c_float, C.c_float, C.c_float,
                                        C.c_float, C.c_float]),
    "bgm_trainer_bn_moving": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "bgm_bgm_train_disc_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgm_bgm_train_gen_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]),
    "bgm_trainer_set_iter": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]),
    "bgm_train_iter_nets": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "bgm_train_iter_latent": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                        C.c_void_p]),
    "bgm_causal_evaluate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgm_bgmtrainer_set_iter": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
    "bgm_bgm_iter_g": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                 C.c_void_p, C.c_void_p]),
    "bgm_bgm_iter_latent": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "bgm_bgm_evaluate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "bgm_gather_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "bgm_hmc_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(VarNetDesc)]),
    "bgm_hmc_destroy": (None, [C.c_void_p]),
    "bgm_hmc_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_longlong),
                               C.POINTER(C.c_longlong)]),
    "bgm_hmc_logpost_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "bgm_hmc_run": (C.c_int, [C.c_void_p, C.POINTER(HmcArgs), C.c_void_p]),
    "bgm_hmc_adapt": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "bgm_hmc_noise": (C.c_int, [C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgm_hmc_predict": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_int64,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgm_bnn_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_float, C.c_float,
                                 C.c_float, C.POINTER(BnnNetDesc), C.POINTER(BnnNetDesc), C.POINTER(BnnNetDesc)]),
    "bgm_bnn_destroy": (None, [C.c_void_p]),
    "bgm_bnn_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_longlong)]),
    "bgm_bnn_set_plan": (C.c_int, [C.c_void_p, C.c_int]),
    "bgm_bnn_scratch_doubles": (C.c_longlong, [C.c_void_p, C.c_int]),
    "bgm_bnn_logpost": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                  C.c_uint64, C.c_int, C.c_int64, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgm_bnn_mh": (C.c_int, [C.c_void_p, C.POINTER(MhArgs), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgm_bnn_effect": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_uint64,
                                 C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgm_bnn_noise": (C.c_int, [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_int64, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgm_lt_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.c_int, C.c_int, C.c_int, C.c_int,
                                C.POINTER(BnnNetDesc), C.POINTER(BnnNetDesc), C.POINTER(BnnNetDesc), C.POINTER(BnnNetDesc),
                                C.POINTER(DiscDesc), C.c_float, C.c_float, C.c_float, C.c_float, C.c_uint64]),
    "bgm_lt_destroy": (None, [C.c_void_p]),
    "bgm_lt_buffers": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "bgm_lt_get_params": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "bgm_lt_set_params": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "bgm_lt_set_call": (C.c_int, [C.c_void_p, C.c_uint32]),
    "bgm_lt_disc_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                   C.c_void_p]),
    "bgm_lt_gen_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                  C.c_void_p]),
    "bgm_lt_adam": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_void_p]),
    "bgm_lt_set_iter": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float]),
    "bgm_lt_iter_nets": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                   C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "bgm_lt_iter_latent": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p]),
    "bgm_lt_evaluate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "bgm_ltb_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(VarNetDesc), C.POINTER(NetDesc),
                                 C.POINTER(DiscDesc), C.POINTER(DiscDesc), C.c_float, C.c_float, C.c_float,
                                 C.c_float, C.c_float]),
    "bgm_ltb_destroy": (None, [C.c_void_p]),
    "bgm_ltb_buffers": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "bgm_ltb_get_params": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "bgm_ltb_bn_moving": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "bgm_ltb_adam": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_void_p]),
    "bgm_ltb_disc_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "bgm_ltb_gen_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p]),
    "bgm_ltb_set_iter": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
    "bgm_ltb_iter_g": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                 C.c_void_p, C.c_void_p]),
    "bgm_ltb_iter_latent": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "bgm_ltb_evaluate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "bgm_ltb_encode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "bgm_host_choice": (C.c_int, [C.POINTER(MtState), C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "bgm_host_normal": (C.c_int, [C.POINTER(MtState), C.c_double, C.c_double, C.c_longlong, C.c_void_p]),
    "bgm_host_rand": (C.c_int, [C.POINTER(MtState), C.c_longlong, C.c_void_p]),
    "bgm_host_egm_stream": (C.c_int, [C.POINTER(MtState), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p]),
    "bgm_hmc_set_engine": (C.c_int, [C.c_void_p, C.c_int]),
    "bgm_hmc_engine_info": (C.c_int, [C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.POINTER(C.c_longlong)]),
    "bgm_column_quantiles": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_double, C.c_double, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p]),
    "bgm_hmc_heads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
}


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "bayesgm_b200: %s is missing.  Build it with `python -m bayesgm_b200._build` "
            "(nvcc, sm_100a).  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the header and the .so disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(fn_name, rc):
    if rc != 0:
        msg = load().bgm_last_error()
        raise BgmError(fn_name, rc, msg.decode() if msg else "?")


def call(fn_name, *args):
    check(fn_name, getattr(load(), fn_name)(*args))


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("bayesgm_b200 needs a CUDA device (sm_100a); there is no CPU fallback.")
    return torch


def ptr(t):
    """Raw device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_MEM_CAP = {}


def free_memory_estimate(torch):
    """Bytes this process can still allocate on the current device, WITHOUT a driver call per use:
    `cudaMemGetInfo` is an ioctl under the driver's global lock and was measured to block the calling thread for
    10-90 ms at a time on a shared multi-GPU host (profiles/r02Y_jitter*.log), which made one predict() call in
    three take 60-160 ms instead of 52.  The driver is asked once per device; afterwards the figure follows the
    allocator's own counters (capacity seen at the first call minus what torch currently holds in live tensors)."""
    dev = torch.cuda.current_device()
    cap = _MEM_CAP.get(dev)
    if cap is None:
        cap = _MEM_CAP[dev] = torch.cuda.mem_get_info()[0] + torch.cuda.memory_reserved()
    return max(cap - torch.cuda.memory_allocated(), 0)

"""ctypes binding of the dktb200 C ABI (include/dktb200.h).

The product path has NO CPU fallback: ``load()`` opens the nvcc-built ``libdktb200.so`` that sits
next to this file (built in-tree by ``deep_kernel_transfer_b200.build`` / ``__graft_entry__.build``)
and raises if it is missing.  ``DktbLib(path)`` can also wrap another library exporting the same ABI --
the unit tests use that to drive the g++ emulation build of the same kernel sources on CPU tensors.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libdktb200.so")

# signature strings: p = device pointer, i = int, l = long, f = float, s = cudaStream_t
SIGNATURES = {
    "dktb_version": ("", ctypes.c_int),
    "dktb_conv1_tiles": ("ii", ctypes.c_int),
    "dktb_conv1_fwd": ("pppppiiis", ctypes.c_int),
    "dktb_conv1_wgrad_nsplit": ("", ctypes.c_int),
    "dktb_conv1_wgrad": ("pppppiiis", ctypes.c_int),
    "dktb_conv1_bwd_fused": ("pppppppppppiiiiis", ctypes.c_int),
    "dktb_conv1_bwd_fused_mma": ("pppppppppppiiiiis", ctypes.c_int),
    "dktb_conv1_wgrad_reduce": ("pipps", ctypes.c_int),
    "dktb_prep_weights": ("ppps", ctypes.c_int),
    "dktb_conv3x3_tiles": ("ii", ctypes.c_int),
    "dktb_conv3x3_fwd": ("pppppiiis", ctypes.c_int),
    "dktb_conv3x3_wgrad_nsplit": ("", ctypes.c_int),
    "dktb_conv3x3_wgrad_scratch_floats": ("", ctypes.c_long),
    "dktb_conv3x3_wgrad": ("pppppiiis", ctypes.c_int),
    "dktb_prep_weights_tc": ("ppps", ctypes.c_int),
    "dktb_conv3x3_tc_weight_floats": ("", ctypes.c_long),
    "dktb_prep_weights_conv1_tc": ("pps", ctypes.c_int),
    "dktb_conv1_tc": ("pppppppppppiiiiis", ctypes.c_int),
    "dktb_conv3x3_tc_fwd": ("ppppppiiis", ctypes.c_int),
    "dktb_conv3x3_wgrad_tc": ("ppppppiiis", ctypes.c_int),
    "dktb_conv_tcg_ok": ("iiiiiii", ctypes.c_int),
    "dktb_conv_tcg_weight_floats": ("iii", ctypes.c_long),
    "dktb_prep_weights_tcg": ("pppiiis", ctypes.c_int),
    "dktb_conv_tcg": ("pppppiiiiiis", ctypes.c_int),
    "dktb_wgrad_tcg_scratch_floats": ("iiiiii", ctypes.c_long),
    "dktb_wgrad_tcg": ("ppppppiiiiiis", ctypes.c_int),
    "dktb_conv_tcg_s2_ok": ("iiii", ctypes.c_int),
    "dktb_conv_tcg_s2_weight_floats": ("ii", ctypes.c_long),
    "dktb_prep_weights_tcg_s2": ("pppiis", ctypes.c_int),
    "dktb_s2d": ("ppiiiiiis", ctypes.c_int),
    "dktb_subsample2": ("ppiiiiiis", ctypes.c_int),
    "dktb_conv_tcg_s2": ("pppppiiiiiis", ctypes.c_int),
    "dktb_wgrad_tcg_s2_ok": ("iiii", ctypes.c_int),
    "dktb_wgrad_tcg_s2_scratch_floats": ("iiiii", ctypes.c_long),
    "dktb_wgrad_tcg_s2": ("ppppppiiiiis", ctypes.c_int),
    "dktb_stem_tc_ok": ("iiiiiiii", ctypes.c_int),
    "dktb_stem_tc_weight_floats": ("", ctypes.c_long),
    "dktb_prep_weights_stem_tc": ("pps", ctypes.c_int),
    "dktb_stem_tc": ("pppppiiis", ctypes.c_int),
    "dktb_stem_wgrad_tc_scratch_floats": ("ii", ctypes.c_long),
    "dktb_stem_wgrad_tc": ("pppppiiis", ctypes.c_int),
    "dktb_zero_border": ("piiiis", ctypes.c_int),
    "dktb_pad_copy": ("ppiiiiis", ctypes.c_int),
    "dktb_conv3x3_wgrad_reduce": ("pipps", ctypes.c_int),
    "dktb_bn_finalize": ("piiiipppppffs", ctypes.c_int),
    "dktb_bn_eval_prepare": ("ppppifs", ctypes.c_int),
    "dktb_bn_relu_pool_fwd": ("ppppppiiiiiiis", ctypes.c_int),
    "dktb_bn_bwd_chunks": ("iii", ctypes.c_int),
    "dktb_bn_relu_pool_bwd": ("ppppppppppppiiiiiiis", ctypes.c_int),
    "dktb_bn_scratch_doubles": ("i", ctypes.c_long),
    "dktb_bn1d_fwd": ("pppppppppiiiiiiiffs", ctypes.c_int),
    "dktb_bn1d_bwd": ("pppppppppiiiiis", ctypes.c_int),
    "dktb_l2norm_fwd": ("ppplifs", ctypes.c_int),
    "dktb_l2norm_bwd": ("pppplis", ctypes.c_int),
    "dktb_gram": ("pppiiiis", ctypes.c_int),
    "dktb_gram_tc_ok": ("iiii", ctypes.c_int),
    "dktb_gram_tc_scratch_floats": ("iiii", ctypes.c_long),
    "dktb_gram_tc": ("pppppiiiis", ctypes.c_int),
    "dktb_gp_max_n": ("", ctypes.c_int),
    "dktb_gp_fit": ("plplpppppppppffiiis", ctypes.c_int),
    "dktb_gp_large_max_n": ("", ctypes.c_int),
    "dktb_gp_large_work_floats": ("iii", ctypes.c_long),
    "dktb_gp_fit_large": ("plplppppppppppffiiis", ctypes.c_int),
    "dktb_gp_reduce": ("ppppiis", ctypes.c_int),
    "dktb_gp_info_accumulate": ("ppis", ctypes.c_int),
    "dktb_gram_bwd": ("pppiiiifs", ctypes.c_int),
    "dktb_gp_predict": ("plpppppiiiis", ctypes.c_int),
    "dktb_center_rows": ("pppiiiis", ctypes.c_int),
    "dktb_row_sqnorm": ("pplis", ctypes.c_int),
    "dktb_sqdist": ("pppiiiis", ctypes.c_int),
    "dktb_kernel_fwd": ("ippppiiiis", ctypes.c_int),
    "dktb_kernel_bwd": ("ipppppppiiis", ctypes.c_int),
    "dktb_gp_predict_var": ("plplppppiiiis", ctypes.c_int),
    "dktb_conv2d_out_size": ("iiiii", ctypes.c_int),
    "dktb_conv2d_fwd": ("ppppiiiiiiiiiiis", ctypes.c_int),
    "dktb_conv2d_dgrad": ("ppppiiiiiiiiiiis", ctypes.c_int),
    "dktb_conv2d_wgrad_nsplit": ("l", ctypes.c_int),
    "dktb_conv2d_mma_ok": ("ii", ctypes.c_int),
    "dktb_conv2d_prep_mma": ("pppiiiis", ctypes.c_int),
    "dktb_conv2d_fwd_mma": ("ppppiiiiiiiiiis", ctypes.c_int),
    "dktb_conv2d_dgrad_mma": ("pppiiiiiiiiiis", ctypes.c_int),
    "dktb_conv2d_flat_k": ("iii", ctypes.c_int),
    "dktb_conv2d_prep_flat_mma": ("ppiiiis", ctypes.c_int),
    "dktb_conv2d_fwd_flat_mma": ("ppppiiiiiiiiiis", ctypes.c_int),
    "dktb_conv2d_wgrad": ("ppppppiiiiiiiiiiis", ctypes.c_int),
    "dktb_nchw_to_nhwc": ("ppiiiis", ctypes.c_int),
    "dktb_spectral_fwd": ("pppppppiiiiiiis", ctypes.c_int),
    "dktb_spectral_bwd": ("ppppppppppiiiiiis", ctypes.c_int),
    "dktb_bn2d_partial_floats": ("iii", ctypes.c_long),
    "dktb_bn2d_stats": ("ppppppiiiiffs", ctypes.c_int),
    "dktb_bn2d_stats_l": ("ppppppiiiiffiis", ctypes.c_int),
    "dktb_bn2d_apply_l": ("pppppppiiiiiiis", ctypes.c_int),
    "dktb_bn2d_bwd_l": ("ppppppppppppiiiiiiis", ctypes.c_int),
    "dktb_add_inplace_l": ("ppiiiiis", ctypes.c_int),
    "dktb_bn2d_apply": ("pppppppiiiiis", ctypes.c_int),
    "dktb_bn2d_bwd": ("ppppppppppppiiiiis", ctypes.c_int),
    "dktb_maxpool3_fwd": ("pppiiiis", ctypes.c_int),
    "dktb_maxpool3_bwd": ("pppiiiis", ctypes.c_int),
    "dktb_avgpool_fwd": ("ppiiis", ctypes.c_int),
    "dktb_avgpool_bwd": ("ppiiis", ctypes.c_int),
    "dktb_add_inplace": ("ppls", ctypes.c_int),
    "dktb_adam_step": ("pppplffffifs", ctypes.c_int),
    "dktb_adam_step_dev": ("pppplffffpfs", ctypes.c_int),
    "dktb_counter_add": ("pis", ctypes.c_int),
    "dktb_scale": ("plfs", ctypes.c_int),
    "dktb_episode_transform_smem": ("iii", ctypes.c_long),
    "dktb_episode_transform": ("pplpppiiiiiiiiffffffps", ctypes.c_int),
}

_CT = {"p": ctypes.c_void_p, "i": ctypes.c_int, "l": ctypes.c_long, "f": ctypes.c_float, "s": ctypes.c_void_p}


class DktbError(RuntimeError):
    pass


class DktbLib:
    def __init__(self, path, signatures=None):
        if not os.path.exists(path):
            raise DktbError(
                "dktb200 native library not found at %s -- build it with `python -m deep_kernel_transfer_b200.build` "
                "(there is no CPU fallback)" % path)
        self.path = path
        self._c = ctypes.CDLL(path)
        self._sig = {}
        self.launches = 0
        for name, (sig, res) in (signatures or SIGNATURES).items():
            try:
                fn = getattr(self._c, name)
            except AttributeError:
                continue
            fn.argtypes = [_CT[ch] for ch in sig]
            fn.restype = res
            self._sig[name] = (fn, sig)

    def has(self, name):
        return name in self._sig

    def exported(self):
        return sorted(self._sig)

    def __getattr__(self, name):
        if name.startswith("dktb_") or name in ("_sig",):
            key = name
        else:
            key = "dktb_" + name
        sig = self.__dict__.get("_sig", {})
        if key not in sig:
            raise AttributeError(name)
        fn, s = sig[key]

        def call(*args):
            if len(args) != len(s):
                raise TypeError("%s expects %d arguments, got %d" % (key, len(s), len(args)))
            conv = []
            dev = None
            for ch, a in zip(s, args):
                if ch == "p":
                    if a is None:
                        conv.append(None)
                    elif isinstance(a, torch.Tensor):
                        if not a.is_contiguous():
                            raise DktbError("%s: non-contiguous tensor argument" % key)
                        if a.is_cuda:
                            # raw pointers carry no device: every tensor of one call must live on the same GPU
                            if dev is None:
                                dev = a.device
                            elif a.device != dev:
                                raise DktbError("%s: tensor arguments on different devices (%s, %s)" % (key, dev, a.device))
                        conv.append(ctypes.c_void_p(a.data_ptr()))
                    else:
                        conv.append(ctypes.c_void_p(int(a)))
                elif ch == "s":
                    conv.append(ctypes.c_void_p(int(a)) if a else None)
                elif ch == "f":
                    conv.append(float(a))
                else:
                    conv.append(int(a))
            if dev is not None and dev.index != torch.cuda.current_device():
                # launch (and cudaFuncSetAttribute) on the tensors' device, not on the thread's current one
                with torch.cuda.device(dev):
                    r = fn(*conv)
            else:
                r = fn(*conv)
            if s.endswith("s") and r != 0:
                raise DktbError("%s failed with status %d (%s)" % (
                    key, r, "bad argument" if r < 0 else "CUDA error"))
            if s.endswith("s"):
                self.launches += 1
            return r
        return call


_LIB = None


def load():
    """The product library.  Fails loudly when the CUDA extension has not been built."""
    global _LIB
    if _LIB is None:
        _LIB = DktbLib(LIB_PATH)
    return _LIB

"""ctypes binding of libidv_b200.so (the C ABI declared in include/idv.h).

There is deliberately no fallback: if the library is missing (and cannot be built because nvcc is
absent) or a call fails, a RuntimeError is raised.  Tensors are passed as raw device pointers."""
import ctypes
import os

import torch

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
ABI_VERSION = 8        # IDV_ABI_VERSION of include/idv.h this binding was written against

c_f32p = ctypes.c_void_p
i32, i64, u64, f32, vp = ctypes.c_int, ctypes.c_int64, ctypes.c_uint64, ctypes.c_float, ctypes.c_void_p

# name -> argtypes (all return int status)
SIGNATURES = {
    "idv_device_sm_count": [ctypes.POINTER(ctypes.c_int)],
    "idv_tapgemm_f32": [vp, i32, i64, vp, i32, i64, i32, i32, vp, vp, i32, vp, vp, i32, vp, i32, i64, i32, f32, i32, vp],
    "idv_stft_fwd": [vp, i32, i32, vp, i32, i32, i32, vp, vp],
    "idv_istft_fwd": [vp, i32, i32, vp, vp, i32, i32, i32, vp, vp, vp],
    "idv_tapgemm_tc": [vp, i32, i32, vp, i32, i32, i32, i32, vp, i32, i32, vp, i32, vp, vp, i32, vp, i32, i64, i64,
                       i32, i32, f32, i32, vp],
    "idv_tapgemm_tc_b2": [vp, i32, i32, vp, i32, i32, i32, i32, vp, i32, i32, vp, vp, i32, vp, vp, i32, vp, i32, i64, i64,
                          i32, i32, f32, i32, vp],
    "idv_tapgemm_tc_splitk": [vp, i32, i32, vp, i32, i32, i32, i32, vp, i32, i32, vp, vp, i32, vp, vp, i32, vp, i32, i64,
                              i64, i32, i32, f32, i32, i32, vp],
    "idv_tapgemm_tc_head": [vp, i32, i32, vp, i32, i32, i32, i32, vp, i32, i32, vp, i32, vp, vp, i32, vp, i32, i64, i64,
                            i32, i32, f32, i32, i32, i32, i32, vp, vp, i32, vp],
    "idv_stft_frames_split": [vp, i32, i32, i32, i32, i32, i32, vp, vp, vp],
    "idv_spec_rows_split": [vp, i32, i32, i32, i32, vp, vp],
    "idv_ola_fwd": [vp, i32, vp, i32, i32, i32, i32, i32, vp, vp, vp],
    "idv_enc0_fwd": [vp, i32, i32, i32, vp, vp, i32, f32, vp, i32, i32, i32, vp, i32, vp],
    "idv_dec5_head_fwd": [vp, i32, vp, i32, i32, i32, i32, i32, vp, vp, f32, i32, vp, vp, i32, i32, vp],
    "idv_lstm_recurrent_fwd": [vp, i64, i64, i32, vp, i32, i32, i32, vp, vp, vp, i32, vp],
    "idv_lstm_recurrent_tc": [vp, i64, i64, i32, vp, i32, i32, i32, vp, vp, vp, vp, i32, vp],
    "idv_lstm2_wave_tc": [vp, i64, i64, i32, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, i32, vp],
    "idv_lstm_layer_pair_tc": [vp, i64, i64, i32, vp, i32, i32, i32, vp, vp, vp, vp, i32, vp],
    "idv_lstm2_cluster_tc": [vp, i64, i64, i32, vp, vp, vp, vp, i32, i32, i32, vp, vp, vp, i32, vp],
    "idv_lstm_h1_fwd": [vp, i32, vp, i32, i32, i32, i32, vp, vp],
    "idv_lstm_combine_fwd": [vp, i32, i32, i32, vp, i32, vp],
    "idv_reparam_fwd": [vp, i32, i32, i32, i32, i32, i32, vp, vp, u64, u64, vp, i32, vp, vp],
    "idv_latent_fwd": [vp, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, vp, u64, u64, vp, vp, vp, vp, vp, i32, i32, vp],
    "idv_lstm_combine_planes": [vp, i32, i32, i32, i32, vp, vp, i32, vp],
    "idv_bin_affine": [vp, i32, i32, i32, vp, vp, i32, vp, vp],
    "idv_planes_to_user": [vp, i32, i32, i32, i32, i32, vp, i32, vp],
    "idv_user_to_planes": [vp, i32, i32, i32, i32, vp, i32, i32, vp],
    "idv_z_to_planes": [vp, i32, i32, i32, i32, i32, vp, i32, i32, vp],
    "idv_cbn_eval_user": [vp, i64, i32, i64, vp, vp, vp],
    "idv_cbn_stats_planes": [vp, i32, i32, i32, i32, i32, vp, i32, vp],
    "idv_cbn_train_finalize": [vp, ctypes.c_double, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, i32, vp, vp, vp],
    "idv_cbn_apply_planes": [vp, i32, i32, i32, i32, i32, vp, i32, f32, i32, vp, i32, vp],
    "idv_planes_transpose_split": [vp, i32, i32, i32, i32, i32, i32, vp, vp],
    "idv_f32_to_split": [vp, i64, vp, vp],
    "idv_cbn_bwd_reduce": [vp, i32, vp, i32, i32, i32, i32, i32, vp, vp, f32, vp, i32, vp],
    "idv_cbn_bwd_finalize": [vp, ctypes.c_double, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "idv_cbn_bwd_apply": [vp, i32, vp, i32, i32, i32, i32, i32, vp, vp, vp, f32, vp, i32, i32, vp],
    "idv_lstm_combine_bwd": [vp, i32, i32, i32, vp, i32, vp],
    "idv_lstm_scan_c": [vp, i32, i32, i32, vp, i32, vp],
    "idv_lstm_cell_bwd_step": [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp],
    "idv_colsum_add": [vp, i64, i32, i32, vp, vp],
    "idv_enc0_wgrad": [vp, vp, i32, i32, i32, i32, i32, vp, vp],
    "idv_adam_step": [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, vp],
    "idv_kl_fwd_bwd": [vp, i32, i32, vp, i32, i32, i64, i32, f32, f32, vp, vp, vp],
    "idv_sisnr_fwd_bwd": [vp, vp, i32, i32, f32, vp, vp, vp, vp],
    "idv_spec_loss_fwd_bwd": [vp, vp, i64, f32, f32, f32, vp, vp, vp],
    "idv_ola_bwd": [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp],
    "idv_head_bwd": [vp, vp, f32, i32, vp, vp, i32, vp, i32, i32, i32, vp, vp, vp],
    "idv_dec5_dgrad": [vp, vp, i32, i32, i32, i32, i32, i32, vp, vp],
    "idv_dec5_wgrad": [vp, i32, vp, i32, i32, i32, i32, i32, i32, vp, vp],
    "idv_axpy": [vp, vp, f32, i64, vp],
    "idv_reparam_bwd": [vp, i32, i32, i32, i32, i32, vp, vp, vp, vp, vp],
    "idv_cbn_stats_user": [vp, i64, i32, i64, vp, vp],
    "idv_head_user": [vp, i64, i64, f32, i32, vp, i32, vp],
    "idv_stream_frames_split": [vp, vp, i32, i32, i64, i32, i32, i32, vp, vp],
    "idv_stream_hist_shift": [vp, vp, i32, i32, i32, i32, vp],
    "idv_lstm_cell_step": [vp, i64, i64, i32, vp, i32, i32, i32, i32, vp, vp, vp, vp],
    "idv_carry_rows": [vp, i32, vp, vp],
    "idv_stream_last_frame": [vp, i32, i32, i32, vp, vp],
    "idv_stream_ola": [vp, i32, vp, vp, i32, i32, i64, i32, i32, vp, vp],
    "idv_stream_tail": [vp, i32, vp, vp, vp, vp, i32, vp, vp, i32, vp, vp, i64, vp, i32, i32, i32, i32, vp],
}
EXPORTS = ["idv_abi_version", "idv_last_error", "idv_lstm_tc_config", "idv_lstm2_wave_config",
           "idv_lstm_layer_pair_config", "idv_lstm2_cluster_config", "idv_lstm2_cluster_concurrency", "idv_set_option"] + \
    list(SIGNATURES)


def lib_path():
    return _build.LIB


def load():
    """Load (building first if the sources are newer and nvcc exists) and type the library."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if _build.needs_build():
        if os.path.exists(_build.NVCC):
            _build.build()
        elif not os.path.exists(path):
            raise RuntimeError("libidv_b200.so is missing and nvcc is not available: the CUDA extension is "
                               "mandatory, there is no CPU or PyTorch fallback")
    lib = ctypes.CDLL(path)
    lib.idv_abi_version.restype = ctypes.c_int
    lib.idv_last_error.restype = ctypes.c_char_p
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = ctypes.c_int
    _LIB = lib
    # IDV_OPTIONS="name=value,name=value": idv_set_option switches for A/B measurements (tools, bench runs)
    for item in filter(None, os.environ.get("IDV_OPTIONS", "").split(",")):
        name, _, value = item.partition("=")
        set_option(name.strip(), int(value))
    return lib


# kernels launched by each entry point (memset nodes not counted)
KERNELS_PER_CALL = {"idv_istft_fwd": 2, "idv_sisnr_fwd_bwd": 2}


class CarryEntry(ctypes.Structure):
    """idv_carry_t (include/idv.h)."""
    _fields_ = [("base", ctypes.c_uint64), ("n_planes", ctypes.c_int64), ("plane_bytes", ctypes.c_int64),
                ("row_bytes", ctypes.c_int32), ("NB", ctypes.c_int32), ("Tp", ctypes.c_int32),
                ("src_row", ctypes.c_int32)]
LAUNCHES = [0]          # running count of kernel launches issued through call()
_PROFILE_HOOK = None


def set_profile_hook(hook):
    """hook(name, (start_event, end_event)) is invoked for every call (CUDA events recorded on the launching
    stream); None disables.  Used by bench.py for the per-kernel device times."""
    global _PROFILE_HOOK
    _PROFILE_HOOK = hook


def resolve_profile(prof):
    """{name: [(e0, e1), ...]} -> {name: [ms, ...]} (call after a device synchronize)."""
    return {k: [e0.elapsed_time(e1) for (e0, e1) in v] for k, v in prof.items()}


E_RESOURCE = 3         # IDV_E_RESOURCE


def call(name, *args, soft_resource=False):
    """Call a C-ABI entry point.  torch tensors are passed as device pointers (they must be contiguous
    CUDA tensors); the current CUDA stream is appended as the trailing ``stream`` argument.  With ``soft_resource`` an
    IDV_E_RESOURCE status (the kernel cannot be made resident on this device) is returned as False instead of raised, for
    entry points whose contract names another entry point to use in that case."""
    lib = load()
    conv = [ptr(a) if isinstance(a, torch.Tensor) else a for a in args]
    conv.append(torch.cuda.current_stream().cuda_stream)
    hook = _PROFILE_HOOK
    if hook is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = getattr(lib, name)(*conv)
    if rc == E_RESOURCE and soft_resource:
        return False
    if rc != 0:
        raise RuntimeError("%s failed (code %d): %s" % (name, rc, lib.idv_last_error().decode()))
    LAUNCHES[0] += KERNELS_PER_CALL.get(name, 1)
    if hook is not None:
        e1.record()
        hook(name, (e0, e1))
    return True


OPTIONS = {}            # options set through set_option in this process (name -> value)


def set_option(name, value):
    lib = load()
    lib.idv_set_option.argtypes = [ctypes.c_char_p, ctypes.c_int]
    lib.idv_set_option.restype = ctypes.c_int
    if lib.idv_set_option(name.encode(), int(value)) != 0:
        raise RuntimeError("idv_set_option failed: %s" % lib.idv_last_error().decode())
    OPTIONS[name] = int(value)


def lstm_tc_config(H):
    """(gate columns per CTA, CTAs per module) of the tensor-core recurrence, or None if H is unsupported."""
    lib = load()
    n, c = ctypes.c_int(0), ctypes.c_int(0)
    lib.idv_lstm_tc_config.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
    lib.idv_lstm_tc_config.restype = ctypes.c_int
    if lib.idv_lstm_tc_config(int(H), ctypes.byref(n), ctypes.byref(c)) != 0:
        return None
    return n.value, c.value


def lstm2_wave_config(H):
    """(gate columns per CTA, CTAs per (module, role), workspace bytes) of the 2-layer wavefront kernel or None."""
    lib = load()
    n, c, w = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int64(0)
    lib.idv_lstm2_wave_config.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                          ctypes.POINTER(ctypes.c_int64)]
    lib.idv_lstm2_wave_config.restype = ctypes.c_int
    if lib.idv_lstm2_wave_config(int(H), ctypes.byref(n), ctypes.byref(c), ctypes.byref(w)) != 0:
        return None
    return n.value, c.value, w.value


def lstm_layer_pair_config(H):
    """(gate columns per CTA, CTAs per module, workspace bytes) of the one-layer CTA-pair recurrence, or None."""
    lib = load()
    n, c, w = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int64(0)
    lib.idv_lstm_layer_pair_config.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int),
                                               ctypes.POINTER(ctypes.c_int64)]
    lib.idv_lstm_layer_pair_config.restype = ctypes.c_int
    if lib.idv_lstm_layer_pair_config(int(H), ctypes.byref(n), ctypes.byref(c), ctypes.byref(w)) != 0:
        return None
    return n.value, c.value, w.value


def lstm2_cluster_config(H, NB, T):
    """(hidden units per CTA, CTAs per cluster, workspace bytes) of the small-batch cluster recurrence, or None."""
    lib = load()
    u, c, w = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int64(0)
    lib.idv_lstm2_cluster_config.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                             ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int64)]
    lib.idv_lstm2_cluster_config.restype = ctypes.c_int
    if lib.idv_lstm2_cluster_config(int(H), int(NB), int(T), ctypes.byref(u), ctypes.byref(c), ctypes.byref(w)) != 0:
        return None
    return u.value, c.value, w.value


def lstm2_cluster_concurrency(H, NB):
    """Clusters of the (H, NB) shape the CURRENT CUDA device holds at once, or None (unsupported shape / no device)."""
    lib = load()
    n = ctypes.c_int(0)
    lib.idv_lstm2_cluster_concurrency.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
    lib.idv_lstm2_cluster_concurrency.restype = ctypes.c_int
    if lib.idv_lstm2_cluster_concurrency(int(H), int(NB), ctypes.byref(n)) != 0:
        return None
    return n.value


_EXCLUSIVE = {}


def check_exclusive_device(device_index):
    """The CTA-pair LSTM kernels (clusters of 2) are launched WITHOUT the cooperative guarantee (Nsight Compute cannot
    replay a launch that carries both the cooperative and the cluster attribute): their CTAs wait on one another, so the
    whole grid must be resident, which holds when this process has the GPU to itself.  Checked once per device through
    NVML: if other compute processes are on the device (MPS clients, a second job), the pair kernels are switched off
    process-wide ("lstm_wave_cta_pairs" = 0) and the recurrences run on the cooperative one-CTA-per-tile kernels, whose
    launch fails cleanly instead of waiting when the grid cannot be co-resident.  Kernels of several streams of THIS
    process sharing the GPU: set the "gemm_dynamic_tiles" option (pipeline.StreamPipeline does), which has the same
    effect.  Returns True when the device is exclusive (or NVML is unavailable: nothing is known, nothing is changed)."""
    if device_index in _EXCLUSIVE:
        return _EXCLUSIVE[device_index]
    ok = True
    try:
        import pynvml
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(device_index)
        h = None
        uuid = getattr(props, "uuid", None)
        if uuid is not None:
            try:
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:
                h = None
        if h is None:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = device_index
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if device_index < len(ids) and ids[device_index].isdigit():
                    phys = int(ids[device_index])
            h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        others = [p for p in pynvml.nvmlDeviceGetComputeRunningProcesses(h) if p.pid != os.getpid()]
        if others:
            ok = False
            import warnings
            warnings.warn("idccrn_b200: %d other compute process(es) share GPU %d - the CTA-pair LSTM kernels need the "
                          "device to themselves and are switched off (cooperative one-CTA kernels are used instead)"
                          % (len(others), device_index))
            set_option("lstm_wave_cta_pairs", 0)
    except Exception:
        ok = True
    _EXCLUSIVE[device_index] = ok
    return ok


def ptr(t):
    """Device pointer of a contiguous fp32/int32 CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("idccrn_b200 ops need CUDA tensors (no CPU fallback); got device %s" % t.device)
    if not t.is_contiguous():
        raise RuntimeError("idccrn_b200 ops need contiguous tensors")
    return t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_f32_cuda(t, what):
    if not isinstance(t, torch.Tensor) or not t.is_cuda or t.dtype != torch.float32:
        raise RuntimeError("%s must be a float32 CUDA tensor (no CPU fallback), got %s %s" % (
            what, getattr(t, "dtype", type(t)), getattr(t, "device", "")))
    return t.contiguous()

"""ctypes binding of librib_b200.so (C ABI declared in include/rib_b200.h).

The extension is the product: there is no Python/PyTorch fallback.  If the shared library is
missing or fails to load, importing this module raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RIB_LIB selects another build of the same library (kernel-variant experiments under tools/); never set in tests or bench
LIB_PATH = os.environ.get('RIB_LIB') or os.path.join(_HERE, 'librib_b200.so')

# every symbol include/rib_b200.h declares
SYMBOLS = [
    'rib_last_error', 'rib_abi_version', 'rib_kernel_launch_count', 'rib_rasterize_workspace_bytes', 'rib_rasterize', 'rib_warp', 'rib_composite',
    'rib_generator_create', 'rib_generator_destroy', 'rib_generator_workspace_bytes', 'rib_generator_bind', 'rib_generator_forward',
    'rib_debug_set_simt', 'rib_debug_get_simt', 'rib_generator_debug_tensor', 'rib_generator_plan_text', 'rib_act_is_fp16',
    'rib_conv_test_scratch_bytes', 'rib_conv_test', 'rib_profile_enable', 'rib_profile_collect', 'rib_profile_collect_launches',
    'rib_tune_log', 'rib_tune_export', 'rib_tune_import', 'rib_conv_test_ex',
    'rib_plan_dry_run', 'rib_frames_from_u8', 'rib_resize_cubic_u8', 'rib_avgpool_test',
    'rib_motion_create', 'rib_motion_destroy', 'rib_motion_workspace_bytes', 'rib_motion_forward',
]


class GenConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in (
        'label_nc', 'img_nc', 'nf', 'maxf', 'n_down', 'n_res', 'emb_nf', 'emb_max', 'emb_down',
        'mask_nf', 'mask_max', 'mask_down', 'mask_res')]


class MotionConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ('input_joints', 'hidden_dim', 'nheads', 'dim_feedforward', 'enc_layers', 'dec_layers')]


class Tensor(C.Structure):
    _fields_ = [('name', C.c_char_p), ('data', C.c_void_p), ('numel', C.c_longlong)]


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            'rib: %s not found - build it with `python -c "import __graft_entry__ as g; g.build()"` '
            '(there is no fallback path)' % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, f64 = C.c_void_p, C.c_int, C.c_longlong, C.c_double
    lib.rib_last_error.restype = C.c_char_p
    lib.rib_last_error.argtypes = []
    lib.rib_abi_version.restype = i32
    lib.rib_kernel_launch_count.restype = i64
    lib.rib_rasterize.restype = i32
    lib.rib_rasterize.argtypes = [vp, i32, i32, i32, C.POINTER(f64), f64, f64, vp, vp, vp, i64, vp]
    lib.rib_rasterize_workspace_bytes.restype = i64
    lib.rib_rasterize_workspace_bytes.argtypes = [i32, i32, i32]
    lib.rib_warp.restype = i32
    lib.rib_warp.argtypes = [vp, vp, i32, vp, i32, i32, i32, i32, i64, i64, i64, vp]
    lib.rib_composite.restype = i32
    lib.rib_composite.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i64, i64, i64, vp]
    lib.rib_resize_cubic_u8.restype = i32
    lib.rib_resize_cubic_u8.argtypes = [vp, vp, i32, i32, i32, i32, i32, i64, i64, vp]
    lib.rib_frames_from_u8.restype = i32
    lib.rib_frames_from_u8.argtypes = [vp, vp, vp, i32, i32, i32, i64, i64, i64, vp]
    lib.rib_generator_create.restype = i32
    lib.rib_generator_create.argtypes = [C.POINTER(GenConfig), C.POINTER(Tensor), i32, vp, C.POINTER(vp)]
    lib.rib_generator_destroy.restype = None
    lib.rib_generator_destroy.argtypes = [vp]
    lib.rib_generator_workspace_bytes.restype = i64
    lib.rib_generator_workspace_bytes.argtypes = [vp, i32, i32, i32]
    lib.rib_generator_bind.restype = i32
    lib.rib_generator_bind.argtypes = [vp, i32, i32, i32, vp, i64, C.POINTER(vp), vp]
    lib.rib_generator_forward.restype = i32
    lib.rib_generator_forward.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp, vp, vp, i64, vp]
    lib.rib_debug_set_simt.restype = None
    lib.rib_debug_set_simt.argtypes = [i32]
    lib.rib_debug_get_simt.restype = i32
    lib.rib_generator_debug_tensor.restype = i32
    lib.rib_generator_debug_tensor.argtypes = [vp, C.c_char_p, C.POINTER(vp)] + [C.POINTER(i32)] * 5
    lib.rib_tune_log.restype = i32
    lib.rib_tune_log.argtypes = [C.c_char_p, i64]
    lib.rib_tune_export.restype = i32
    lib.rib_tune_export.argtypes = [C.c_char_p, i64]
    lib.rib_tune_import.restype = i32
    lib.rib_tune_import.argtypes = [C.c_char_p]
    lib.rib_generator_plan_text.restype = i32
    lib.rib_generator_plan_text.argtypes = [vp, C.c_char_p, i64]
    lib.rib_act_is_fp16.restype = i32
    lib.rib_profile_enable.restype = None
    lib.rib_profile_enable.argtypes = [i32]
    lib.rib_profile_collect.restype = i32
    lib.rib_profile_collect.argtypes = [C.POINTER(f64), C.POINTER(i64)]
    lib.rib_profile_collect_launches.restype = i32
    lib.rib_profile_collect_launches.argtypes = [C.POINTER(f64), C.POINTER(i64), C.POINTER(C.c_float), i64]
    lib.rib_conv_test_scratch_bytes.restype = i64
    lib.rib_conv_test_scratch_bytes.argtypes = [i32, i32, i32]
    lib.rib_conv_test.restype = i32
    lib.rib_conv_test.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.rib_plan_dry_run.restype = i32
    lib.rib_plan_dry_run.argtypes = [C.POINTER(GenConfig), i32, i32, i32, C.POINTER(i64), C.c_char_p, i64]
    lib.rib_conv_test_ex.restype = i32
    lib.rib_conv_test_ex.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp, vp, i32, vp, vp]
    lib.rib_motion_create.restype = i32
    lib.rib_motion_create.argtypes = [C.POINTER(MotionConfig), C.POINTER(Tensor), i32, vp, C.POINTER(vp)]
    lib.rib_motion_destroy.restype = None
    lib.rib_motion_destroy.argtypes = [vp]
    lib.rib_motion_workspace_bytes.restype = i64
    lib.rib_motion_workspace_bytes.argtypes = [vp, i32, i32]
    lib.rib_motion_forward.restype = i32
    lib.rib_motion_forward.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, i32, vp, vp, vp, i64, vp]
    lib.rib_avgpool_test.restype = i32
    lib.rib_avgpool_test.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    return lib


lib = _load()

# Tuning table (see rib_tune_import): the shipped table for the benchmark shapes, then the user's file.  With
# RIB_TUNE_FILE set, everything tuned in this process is written back to that file at exit.
TUNE_TABLE = os.path.join(_HERE, 'tune_b200.txt')


def _load_tune_tables():
    import atexit
    user = os.environ.get('RIB_TUNE_FILE')
    for path in ((None if os.environ.get('RIB_NO_TUNE_TABLE') else TUNE_TABLE), user):
        if path and os.path.isfile(path):
            with open(path, 'rb') as f:
                lib.rib_tune_import(f.read() + b'\0')
    if user:
        def _save():
            buf = C.create_string_buffer(1 << 20)
            if lib.rib_tune_export(buf, len(buf)) == 0 and buf.value:
                with open(user, 'wb') as f:
                    f.write(buf.value)
        atexit.register(_save)


if os.environ.get('RIB_AUTOTUNE', '1') != '0':
    _load_tune_tables()


def check(rc, what):
    """Non-zero C return codes become RuntimeError (the reference raises Python exceptions)."""
    if rc != 0:
        msg = lib.rib_last_error()
        raise RuntimeError('%s failed (%d): %s' % (what, rc, msg.decode() if msg else ''))

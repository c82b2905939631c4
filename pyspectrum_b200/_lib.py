"""ctypes binding of libpsb200.so (include/psb200.h).  No fallback: if the library or a CUDA device is
missing, compute calls raise -- the product path never routes through a CPU implementation."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, 'csrc')
_SO = os.environ.get('PSB200_LIB') or os.path.join(_CSRC, 'libpsb200.so')      # PSB200_LIB: A/B runs of two builds in one session
_LIB = None

ERRORS = {0: 'PSB_OK', -1: 'PSB_ERR_ARG', -2: 'PSB_ERR_UNSUPPORTED_N', -3: 'PSB_ERR_CUDA', -4: 'PSB_ERR_WORKSPACE'}


class PsbError(RuntimeError):
    def __init__(self, code, where):
        self.code = code
        msg = lib().psb_error_string(code).decode() if _LIB is not None else ''
        RuntimeError.__init__(self, '%s failed: %s (%s)' % (where, ERRORS.get(code, code), msg))


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a into csrc/libpsb200.so (nvcc cross-compiles without a GPU)."""
    cmd = ['make', '-C', _CSRC, '-j', str(min(8, os.cpu_count() or 1))]
    if force:
        subprocess.check_call(['make', '-C', _CSRC, 'clean'])
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode:
        print(out.stdout)
    if out.returncode:
        raise RuntimeError('building libpsb200.so failed')
    return _SO


_c = ctypes
_vp, _i, _i64, _f, _d, _sz = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float, _c.c_double, _c.c_size_t

# name -> (restype, argtypes); mirrors include/psb200.h one to one (tests/test_cabi_and_host.py checks it)
PROTOTYPES = {
    'psb_version': (_i, []),
    'psb_error_string': (_c.c_char_p, [_i]),
    'psb_twiddles_f32': (_i, [_i, _vp]),
    'psb_twiddles_f64': (_i, [_i, _vp]),
    'psb_fcomb_tables': (_i, [_i, _vp, _vp]),
    'psb_rsd_trig': (_i, [_i, _vp]),
    'psb_rsd_bin_table': (_i, [_i, _i, _i, _vp]),
    'psb_irk_table_f32': (_i, [_f, _i, _vp]),
    'psb_assign_workspace_bytes': (_sz, [_i64, _i]),
    'psb_assign_pcs_interlaced': (_i, [_vp, _i, _i, _vp, _i, _i64, _i, _d, _f, _f, _vp, _i, _vp, _sz, _vp, _vp]),
    'psb_slab_route_count': (_i, [_vp, _i, _i, _vp, _i, _i64, _i, _d, _f, _f, _i, _i, _vp, _vp, _vp]),
    'psb_slab_route_scatter': (_i, [_vp, _i, _i, _vp, _i, _i64, _i, _d, _f, _f, _i, _i, _vp, _vp, _vp, _vp]),
    'psb_slab_route_scatter_peer': (_i, [_vp, _i, _i, _vp, _i, _i64, _i, _d, _f, _f, _i, _i, _vp, _vp, _vp]),
    'psb_assign_slab': (_i, [_vp, _i64, _i, _f, _f, _i, _i, _vp, _i, _vp, _sz, _vp, _vp]),
    'psb_survey_prepare': (_i, [_vp, _vp, _vp, _i64, _vp, _i, _d, _d, _vp, _vp, _vp, _vp]),
    'psb_apply_rsd': (_i, [_vp, _vp, _i64, _i, _d, _d, _vp, _vp]),
    'psb_fft_mesh_to_delta': (_i, [_vp, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    'psb_fft_c2c_3d': (_i, [_vp, _i, _i, _vp, _vp]),
    'psb_fcomb': (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _vp]),
    'psb_fft_slab_xy': (_i, [_vp, _i, _i, _i, _vp, _vp]),
    'psb_fft_slab_z': (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    'psb_slab_split_ab': (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    'psb_slab_split_ab_routed': (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp]),
    'psb_slab_fcomb': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp]),
    'psb_pk_monopole': (_i, [_vp, _i, _vp, _i, _d, _vp, _vp]),
    'psb_pk_multipoles': (_i, [_vp, _i, _vp, _i, _i, _f, _vp, _vp, _vp]),
    'psb_pk_monopole_slab': (_i, [_vp, _i, _i, _i, _vp, _i, _d, _vp, _vp]),
    'psb_pk_multipoles_slab': (_i, [_vp, _i, _i, _i, _vp, _i, _i, _f, _vp, _vp, _vp]),
    'psb_half_extract': (_i, [_vp, _i, _i, _i, _vp, _i, _vp]),
    'psb_pk_kmu_python': (_i, [_vp, _i, _vp, _i, _i, _d, _vp, _vp, _vp]),
    'psb_shell_mode_counts': (_i, [_i, _vp, _i, _vp, _vp]),
    'psb_bk_shell_pair_f32': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    'psb_bk_shell_pair_f32_routed': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    'psb_bk_shell_power': (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    'psb_bk_shell_scales': (_i, [_vp, _i, _f, _vp, _vp]),
    'psb_bk_shell_pair_f64': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    'psb_bk_triangle_workspace_bytes': (_sz, [_i]),
    'psb_bk_triangle_sums_f32': (_i, [_vp, _i, _i64, _vp, _i, _vp, _vp, _sz, _i, _vp]),
    'psb_bk_triangle_sums_f64': (_i, [_vp, _i, _i64, _vp, _i, _vp, _vp, _sz, _vp]),
    'psb_bk_triangle_tc_workspace_bytes': (_sz, [_i, _i]),
    'psb_bk_triangle_sums_tc': (_i, [_vp, _i, _i64, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _sz, _vp]),
    'psb_bk_build_tiles': (_i, [_vp, _i, _i, _vp, _vp]),
    'psb_host_assign_quad': (_i, [_vp, _vp, _vp, _i64, _i, _f, _f, _i, _i, _i, _i]),
    'psb_host_fcomb_periodic': (_i, [_vp, _f, _i]),
    'psb_host_fcomb_survey': (_i, [_vp, _i]),
    'psb_host_ffting': (_i, [_vp, _i]),
    'psb_host_pk_pbox_rsd': (_i, [_vp] * 10 + [_i] * 5),
    'psb_host_bk_counts': (_i, [_vp, _i, _f, _i, _i]),
    'psb_host_fivedelta2g_1': (_i, [_vp, _vp, _vp, _i]),
    'psb_host_fivedelta2g_2': (_i, [_vp, _vp, _vp, _vp, _vp, _i]),
    'psb_host_build_quad': (_i, [_vp, _vp, _i, _i]),
    'psb_quad_weights': (_i, [_vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp]),
    'psb_quad_fields': (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp]),
}


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.isfile(_SO):
            raise RuntimeError('%s is missing: run `python -c "import __graft_entry__ as g; g.build()"` '
                               '(there is no CPU fallback)' % _SO)
        import torch  # noqa: F401  (loads libcudart.so.12 so the library shares torch's CUDA runtime)
        L = ctypes.CDLL(_SO)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _LIB = L
    return _LIB


def check(code, where):
    if code != 0:
        raise PsbError(code, where)

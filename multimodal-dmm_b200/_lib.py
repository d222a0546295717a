"""ctypes binding of include/bfvi.h (libbfvi_b200.so).

The product path is CUDA only: `load()` raises when the nvcc-built library is
missing, and every model entry point raises when its tensors are not on a CUDA
device.  There is no CPU implementation behind these calls.
"""
import ctypes as C
import os

MAX_MODS = 16
MAX_SETS = MAX_MODS + 1
MAX_EXPERTS = MAX_MODS + 2

DIST_CODES = {'Normal': 0, 'Bernoulli': 1, 'Categorical': 2}
DIR_FWD, DIR_BWD = 0, 1
MODE_CODES = {'bfilter': 0, 'ffilter': 1, 'fsmooth': 2, 'bsmooth': 3}
EXPERT_TENSOR, EXPERT_INV_PRIOR = 0, 1
PRECISION_CODES = {'tf32x3': 0, 'tf32': 1, 'fused': 2}
PHASES = ('match', 'encode_fwd', 'filter_f_fwd', 'filter_s_flt_fwd', 'filter_s_smt_fwd', 'decode_nll',
          'filter_s_smt_bwd', 'filter_s_flt_bwd', 'filter_f_bwd', 'encode_bwd', 'finalize')

_HERE = os.path.dirname(os.path.abspath(__file__))
# BFVI_LIB_PATH: tuning aid (tools/variants.py builds differently-tiled variants of the same library)
LIB_PATH = os.environ.get('BFVI_LIB_PATH') or os.path.join(_HERE, 'libbfvi_b200.so')


class Model(C.Structure):
    _fields_ = [('n_mods', C.c_int32), ('z_dim', C.c_int32), ('h_dim', C.c_int32),
                ('dims', C.c_int32 * MAX_MODS), ('dists', C.c_int32 * MAX_MODS),
                ('min_std', C.c_float)]


class MlpLayout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                ('in_to_h_w', 'in_to_h_b', 'mean_w', 'mean_b', 'std_w', 'std_b', 'begin', 'end')]


class GtfLayout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                ('gate0_w', 'gate0_b', 'gate2_w', 'gate2_b', 'lin_w', 'lin_b',
                 'nonlin0_w', 'nonlin0_b', 'nonlin2_w', 'nonlin2_b', 'std_w', 'std_b',
                 'begin', 'end')]


class Layout(C.Structure):
    _fields_ = [('z0_mean', C.c_int64), ('z0_log_std', C.c_int64),
                ('enc', MlpLayout * MAX_MODS), ('dec', MlpLayout * MAX_MODS),
                ('trans', GtfLayout * 2), ('total', C.c_int64)]


class Expert(C.Structure):
    _fields_ = [('mean', C.c_void_p), ('std', C.c_void_p), ('mask', C.c_void_p),
                ('stride_s', C.c_int64), ('stride_t', C.c_int64), ('stride_b', C.c_int64),
                ('mstride_s', C.c_int64), ('mstride_t', C.c_int64), ('mstride_b', C.c_int64),
                ('d_mean', C.c_void_p), ('d_std', C.c_void_p),
                ('kind', C.c_int32), ('zero_mask_last_t', C.c_int32)]


class Noise(C.Structure):
    _fields_ = [('eps', C.c_void_p), ('seed', C.c_uint64), ('stream_id', C.c_uint32),
                ('b_offset', C.c_uint32), ('seed_dev', C.c_void_p)]


class FilterArgs(C.Structure):
    _fields_ = [('T', C.c_int32), ('B', C.c_int32), ('S', C.c_int32), ('n_experts', C.c_int32),
                ('experts', Expert * MAX_EXPERTS),
                ('set_expert_bits', C.c_uint32 * MAX_SETS),
                ('direction', C.c_int32), ('n_particles', C.c_int32),
                ('sample', C.c_int32), ('sample_init', C.c_int32),
                ('noise', Noise),
                ('infer_mean', C.c_void_p), ('infer_std', C.c_void_p),
                ('prior_mean', C.c_void_p), ('prior_std', C.c_void_p),
                ('samples', C.c_void_p),
                ('seq_mask', C.c_void_p), ('kl_weight', C.c_float), ('loss_acc', C.c_void_p),
                ('d_infer_mean', C.c_void_p), ('d_infer_std', C.c_void_p),
                ('d_prior_mean', C.c_void_p), ('d_prior_std', C.c_void_p),
                ('d_samples', C.c_void_p),
                ('workspace', C.c_void_p), ('workspace_bytes', C.c_size_t)]


class StepArgs(C.Structure):
    _fields_ = [('T', C.c_int32), ('B', C.c_int32),
                ('inputs', C.c_void_p * MAX_MODS), ('targets', C.c_void_p * MAX_MODS),
                ('seq_mask', C.c_void_p),
                ('kld_mult', C.c_float), ('rec_mults', C.c_float * MAX_MODS),
                ('uni_loss', C.c_int32), ('f_mode', C.c_int32), ('s_mode', C.c_int32),
                ('f_mult', C.c_float), ('s_mult', C.c_float), ('match_mult', C.c_float),
                ('train_particles', C.c_int32), ('match_particles', C.c_int32),
                ('sample', C.c_int32), ('sample_init', C.c_int32),
                ('eps_match', C.c_void_p), ('eps_filt', C.c_void_p),
                ('eps_sflt', C.c_void_p), ('eps_ssmt', C.c_void_p),
                ('seed', C.c_uint64), ('b_offset', C.c_uint32), ('match_count', C.c_float),
                ('seed_dev', C.c_void_p),
                ('precision', C.c_int32), ('batch_tile', C.c_int32)]


class ForwardArgs(C.Structure):
    _fields_ = [('T', C.c_int32), ('B', C.c_int32),
                ('inputs', C.c_void_p * MAX_MODS),
                ('mode', C.c_int32), ('sample', C.c_int32), ('sample_init', C.c_int32),
                ('flt_particles', C.c_int32), ('smt_particles', C.c_int32),
                ('eps_flt', C.c_void_p), ('eps_smt', C.c_void_p),
                ('seed', C.c_uint64), ('b_offset', C.c_uint32), ('precision', C.c_int32),
                ('infer_mean', C.c_void_p), ('infer_std', C.c_void_p),
                ('prior_mean', C.c_void_p), ('prior_std', C.c_void_p),
                ('recon_mean', C.c_void_p * MAX_MODS), ('recon_std', C.c_void_p * MAX_MODS)]


class MlpDesc(C.Structure):           # bfvi_mlp_desc
    _fields_ = [('emb', C.c_void_p), ('w1', C.c_void_p), ('b1', C.c_void_p), ('wa', C.c_void_p), ('ba', C.c_void_p),
                ('wb', C.c_void_p), ('bb', C.c_void_p),
                ('n_in', C.c_int32), ('h_dim', C.c_int32), ('n_out', C.c_int32), ('n_classes', C.c_int32),
                ('head', C.c_int32), ('nan_mask', C.c_int32), ('min_std', C.c_float), ('pad_', C.c_int32)]


class MlpGrads(C.Structure):          # bfvi_mlp_grads
    _fields_ = [(k, C.c_void_p) for k in ('emb', 'w1', 'b1', 'wa', 'ba', 'wb', 'bb')]


HEAD_GAUSSIAN, HEAD_SOFTMAX = 0, 1
ACT_NONE, ACT_SIGMOID = 0, 2


class ConvGeom(C.Structure):
    """bfvi_conv_geom: small map = Conv2d output / ConvTranspose2d input, big map = the other side."""
    _fields_ = [(n, C.c_int32) for n in ('n', 'c_small', 'h_small', 'w_small', 'c_big', 'h_big', 'w_big',
                                         'kernel', 'stride', 'padding')]


class BfviError(RuntimeError):
    pass


# order of the BFVI_STRUCT_* enum (bfvi_sizeof)
STRUCTS = (Model, Layout, Expert, Noise, FilterArgs, StepArgs, ForwardArgs, ConvGeom)


# every symbol include/bfvi.h declares (tests check that the library exports them all)
SYMBOLS = {
    'bfvi_version': (C.c_int, []),
    'bfvi_build_id': (C.c_char_p, []),
    'bfvi_last_error': (C.c_char_p, []),
    'bfvi_last_dispatch': (C.c_char_p, []),
    'bfvi_sizeof': (C.c_size_t, [C.c_int32]),
    'bfvi_param_layout': (C.c_int, [C.POINTER(Model), C.POINTER(Layout)]),
    'bfvi_kernel_family': (C.c_int, [C.POINTER(Model)]),
    'bfvi_encode_fwd': (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'bfvi_encode_bwd': (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                  C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    'bfvi_decode_fwd': (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    'bfvi_decode_bwd': (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                  C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'bfvi_decode_nll': (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    'bfvi_filter_workspace': (C.c_int, [C.POINTER(Model), C.POINTER(FilterArgs), C.POINTER(C.c_size_t)]),
    'bfvi_filter_fwd': (C.c_int, [C.POINTER(Model), C.c_void_p, C.POINTER(FilterArgs), C.c_void_p]),
    'bfvi_filter_bwd': (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_void_p, C.POINTER(FilterArgs),
                                  C.c_void_p]),
    'bfvi_kld_fwd': (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    'bfvi_kld_bwd': (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_int32, C.c_float] + [C.c_void_p] * 5),
    'bfvi_nll_gauss_fwd': (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    'bfvi_nll_gauss_bwd': (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int32, C.c_float]
                           + [C.c_void_p] * 3),
    'bfvi_nll_bernoulli_fwd': (C.c_int, [C.c_void_p] * 3 + [C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    'bfvi_nll_bernoulli_bwd': (C.c_int, [C.c_void_p] * 3 + [C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]),
    'bfvi_nll_categorical_fwd': (C.c_int, [C.c_void_p] * 3 + [C.c_int64, C.c_int32, C.c_void_p, C.c_void_p]),
    'bfvi_nll_categorical_bwd': (C.c_int, [C.c_void_p] * 3 + [C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_void_p]),
    'bfvi_len_to_mask': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    'bfvi_pad_merge': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    'bfvi_unpad': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_void_p,
                             C.c_void_p]),
    'bfvi_ssim_scratch': (C.c_size_t, [C.c_int32] * 5),
    'bfvi_ssim': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float),
                            C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'bfvi_seq_mse_splits': (C.c_int, [C.c_int32, C.c_int32]),
    'bfvi_seq_mse': (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64), C.c_int32,
                               C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                               C.c_void_p]),
    'bfvi_delete_rows': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    'bfvi_delete_spans': (C.c_int, [C.c_void_p] * 4 + [C.c_int32, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p]),
    'bfvi_draw_deletions': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_uint64, C.c_uint32,
                                      C.c_uint32, C.c_void_p, C.c_void_p]),
    'bfvi_step_workspace': (C.c_int, [C.POINTER(Model), C.POINTER(StepArgs), C.POINTER(C.c_size_t)]),
    'bfvi_step_fwd_bwd': (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_void_p, C.POINTER(StepArgs),
                                    C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_int32),
                                    C.c_void_p]),
    'bfvi_step_profile': (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_void_p, C.POINTER(StepArgs),
                                    C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_float),
                                    C.c_void_p]),
    'bfvi_forward_workspace': (C.c_int, [C.POINTER(Model), C.POINTER(ForwardArgs), C.POINTER(C.c_size_t)]),
    'bfvi_forward': (C.c_int, [C.POINTER(Model), C.c_void_p, C.POINTER(ForwardArgs), C.c_void_p, C.c_size_t,
                               C.c_void_p]),
    'bfvi_linear_tf32': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                   C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    'bfvi_wgrad_tf32': (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                  C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    'bfvi_mlp_workspace': (C.c_int, [C.POINTER(MlpDesc), C.c_int64, C.c_int32, C.POINTER(C.c_size_t)]),
    'bfvi_mlp_fwd': (C.c_int, [C.POINTER(MlpDesc), C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_size_t, C.c_void_p]),
    'bfvi_mlp_bwd': (C.c_int, [C.POINTER(MlpDesc), C.POINTER(MlpGrads), C.c_void_p, C.c_int64] + [C.c_void_p] * 5
                     + [C.c_void_p, C.c_size_t, C.c_void_p]),
    'bfvi_gtf_workspace': (C.c_size_t, [C.POINTER(Model), C.c_int64]),
    'bfvi_gtf_fwd': (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_int32, C.c_void_p, C.c_int64] + [C.c_void_p] * 4
                     + [C.c_int32, C.c_void_p, C.c_size_t, C.c_void_p]),
    'bfvi_gtf_bwd': (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64]
                     + [C.c_void_p] * 5 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    'bfvi_gtf_probe': (C.c_int, [C.POINTER(Model), C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_int32,
                                 C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_float), C.c_void_p]),
    'bfvi_adam_step': (C.c_int, [C.c_void_p] * 4 + [C.c_int64] + [C.c_float] * 5 + [C.c_int32, C.c_float, C.c_float,
                                                                                    C.c_void_p, C.c_void_p]),
    'bfvi_conv_gather': (C.c_int, [C.POINTER(ConvGeom)] + [C.c_void_p] * 4 + [C.c_int32, C.c_void_p]),
    'bfvi_conv_scatter': (C.c_int, [C.POINTER(ConvGeom)] + [C.c_void_p] * 4 + [C.c_int32, C.c_void_p]),
    'bfvi_conv_wgrad': (C.c_int, [C.POINTER(ConvGeom)] + [C.c_void_p] * 3 + [C.c_void_p]),
    'bfvi_chan_scratch': (C.c_size_t, [C.c_int32]),
    'bfvi_chan_bias_grad': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_size_t,
                                      C.c_void_p]),
    'bfvi_bn2d_fwd': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64] + [C.c_void_p] * 4
                      + [C.c_int32, C.c_float, C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                         C.c_void_p]),
    'bfvi_bn2d_bwd': (C.c_int, [C.c_void_p] * 5 + [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32]
                      + [C.c_void_p] * 3 + [C.c_void_p, C.c_size_t, C.c_void_p]),
    'bfvi_sigmoid_bwd': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    'bfvi_dense_fwd': (C.c_int, [C.c_void_p] * 4 + [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]),
    'bfvi_dense_bwd': (C.c_int, [C.c_void_p] * 5 + [C.c_int64, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 3
                       + [C.c_void_p, C.c_size_t, C.c_void_p]),
    'bfvi_ffma_probe': (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    'bfvi_dump_noise': (C.c_int, [C.c_uint64, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
}


class Library(object):
    """A bound libbfvi: `lib.call('bfvi_x', ...)` raises BfviError on failure."""

    def __init__(self, path):
        self.path = path
        self.dll = C.CDLL(path)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(self.dll, name)
            fn.restype, fn.argtypes = res, args
        # a stale binding must fail here, not hand the library a short struct
        for which, cls in enumerate(STRUCTS):
            want = self.dll.bfvi_sizeof(which)
            if want != C.sizeof(cls):
                raise BfviError('%s: ctypes struct %s is %d bytes, the library expects %d (binding out of date)'
                                % (path, cls.__name__, C.sizeof(cls), want))

    def build_id(self):
        """digest of the sources this binary was compiled from ('unknown' when the build did not stamp one)"""
        return self.dll.bfvi_build_id().decode('ascii', 'replace')

    def last_dispatch(self):
        return self.dll.bfvi_last_dispatch().decode('utf-8', 'replace').split(';')

    def call(self, name, *args):
        rc = getattr(self.dll, name)(*args)
        if rc != 0:
            raise BfviError('%s failed (%d): %s' % (
                name, rc, self.dll.bfvi_last_error().decode('utf-8', 'replace')))

    def layout(self, model):
        lay = Layout()
        self.call('bfvi_param_layout', C.byref(model), C.byref(lay))
        return lay


def source_id():
    """The digest build.py stamps into the library, recomputed from the source tree beside this file (None when the
    sources are not there)."""
    import importlib.util
    path = os.path.join(_HERE, 'build.py')
    if not os.path.isdir(os.path.join(_HERE, 'csrc')) or not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location('_bfvi_build_id', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.source_id()


_lib = None


def load():
    """Loads the in-tree CUDA library; raises when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BfviError(
                '%s is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                '(nvcc, sm_100a).  There is no CPU fallback.' % LIB_PATH)
        _lib = Library(LIB_PATH)
        want = source_id() if LIB_PATH.startswith(_HERE) else None
        if want is not None and _lib.build_id() not in ('unknown', want):
            import warnings
            warnings.warn('%s was built from other sources (library %s, tree %s): rebuild with '
                          '`python -c "import __graft_entry__ as g; g.build()"`' % (LIB_PATH, _lib.build_id(), want))
    return _lib


def make_model(dims, dists, z_dim, h_dim, min_std):
    if len(dims) > MAX_MODS:
        raise BfviError('at most %d modalities are supported' % MAX_MODS)
    m = Model()
    m.n_mods, m.z_dim, m.h_dim, m.min_std = len(dims), int(z_dim), int(h_dim), float(min_std)
    for i, (d, dist) in enumerate(zip(dims, dists)):
        m.dims[i] = int(d)
        m.dists[i] = DIST_CODES[dist]
    return m


def param_slots(modalities, dists, lay):
    """[(state_dict key, offset, shape-less)] for the flat buffer, in layout order.
    Only default Gaussian MLP encoders/decoders live in the flat buffer."""
    slots = [('z0_mean', lay.z0_mean), ('z0_log_std', lay.z0_log_std)]

    def mlp(prefix, l):
        return [(prefix + '.in_to_h.0.weight', l.in_to_h_w), (prefix + '.in_to_h.0.bias', l.in_to_h_b),
                (prefix + '.h_to_mean.weight', l.mean_w), (prefix + '.h_to_mean.bias', l.mean_b),
                (prefix + '.h_to_std.0.weight', l.std_w), (prefix + '.h_to_std.0.bias', l.std_b)]
    for i, m in enumerate(modalities):
        if dists[i] != 'Categorical':
            slots += mlp('enc.%s' % m, lay.enc[i])
    for i, m in enumerate(modalities):
        if dists[i] == 'Normal':
            slots += mlp('dec.%s' % m, lay.dec[i])
    for d, name in ((0, 'fwd'), (1, 'bwd')):
        l, p = lay.trans[d], 'trans.%s' % name
        slots += [(p + '.z_to_gate.0.weight', l.gate0_w), (p + '.z_to_gate.0.bias', l.gate0_b),
                  (p + '.z_to_gate.2.weight', l.gate2_w), (p + '.z_to_gate.2.bias', l.gate2_b),
                  (p + '.z_lin.weight', l.lin_w), (p + '.z_lin.bias', l.lin_b),
                  (p + '.z_nonlin.0.weight', l.nonlin0_w), (p + '.z_nonlin.0.bias', l.nonlin0_b),
                  (p + '.z_nonlin.2.weight', l.nonlin2_w), (p + '.z_nonlin.2.bias', l.nonlin2_b),
                  (p + '.z_to_std.0.weight', l.std_w), (p + '.z_to_std.0.bias', l.std_b)]
    return slots


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())

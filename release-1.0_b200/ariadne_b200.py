"""ctypes binding of libariadne_b200.so (include/ariadne_b200.h) plus thin classes named after the reference's
interfaces for the grid-reachability path:

  HybridReachabilityAnalyser.upper_reach / upper_evolve / upper_reach_evolve / outer_chain_reach
      (reference: include/reachability_analyser.h:122-172, src/reachability_analyser.cpp:278-493)
  Evolver.reach_evolve        (reference: include/evolver_interface.h:114-124; batch form of include/workers.h:51-81)
  Evolver.flow_bounds         (reference: include/integrator_interface.h:45-50)

There is no CPU fallback: loading fails if the shared library has not been built, and every compute call fails with
Ab200Error(AB200_E_CUDA) when no B200 is present.
"""
import ctypes as C
import os

import numpy as np

MAX_DIM = 5
BOX_STRIDE = 2 * MAX_DIM

E_NAMES = {
    0: "OK", 1: "SqrtNumericException", 2: "DivideByZeroException", 3: "DomainException", 4: "BounderAssertion",
    5: "ArenaExhausted", 6: "OutputExhausted", 7: "Unsupported", 8: "ImplicitFunctionException",
    9: "WorkingSetTooLargeException", 100: "CudaError", 101: "ArgumentError", 102: "KeySpaceError", 103: "TimeoutException",
}

AB200_OK = 0
AB200_E_SQRT, AB200_E_DIVZERO, AB200_E_DOMAIN, AB200_E_BOUNDER, AB200_E_ARENA, AB200_E_OUTPUT, AB200_E_UNSUPPORTED, \
    AB200_E_IMPLICIT, AB200_E_WORKING_SETS = 1, 2, 3, 4, 5, 6, 7, 8, 9
AB200_E_CUDA, AB200_E_ARGUMENT, AB200_E_KEYSPACE, AB200_E_TIMEOUT = 100, 101, 102, 103

EXPORTS = [
    "ab200_default_settings", "ab200_ctx_create", "ab200_ctx_destroy", "ab200_last_error", "ab200_set_system_text",
    "ab200_system_info", "ab200_system_check", "ab200_location_index", "ab200_set_settings", "ab200_set_location_tuning", "ab200_set_accuracy", "ab200_set_accuracy_for_domain",
    "ab200_upper_reach_evolve", "ab200_cellset_view", "ab200_cellset_free", "ab200_key_decode", "ab200_reach_evolve_boxes",
    "ab200_boxlist_view", "ab200_boxlist_free", "ab200_flow_bounds", "ab200_analyser_create", "ab200_analyser_destroy",
    "ab200_analyser_info", "ab200_analyser_set_lock_time", "ab200_analyser_upper_reach_evolve",
    "ab200_analyser_outer_chain_reach", "ab200_analyser_lower_reach_evolve", "ab200_ctx_get_settings", "ab200_analyser_discretise", "ab200_analyser_cells_evolved", "ab200_analyser_enumerate", "ab200_probe_fp64", "ab200_probe_fp64_fma", "ab200_comm_unique_id", "ab200_comm_init", "ab200_comm_destroy", "ab200_comm_info",
    "ab200_integrator_flow_bounds", "ab200_integrator_flow", "ab200_taylor_models_view", "ab200_taylor_models_free", "ab200_taylor_model_op", "ab200_interval_op", "ab200_analyser_discretise_constraint", "ab200_analyser_upper_reach_evolve_from", "ab200_analyser_outer_chain_reach_from",
    "ab200_reach_evolve_enclosures", "ab200_enclosures_view", "ab200_enclosures_model", "ab200_enclosures_select", "ab200_enclosures_free", "ab200_enclosures_create",
    "ab200_analyser_epsilon_lower_chain_reach", "ab200_analyser_lower_reach_evolve_constraint", "ab200_analyser_epsilon_lower_chain_reach_constraint",
]


class Settings(C.Structure):
    _fields_ = [("temporal_order", C.c_int32), ("spacial_order", C.c_int32), ("sweep_threshold", C.c_double),
                ("maximum_enclosure_widths_ratio", C.c_double), ("max_boxes", C.c_int32), ("arena_words", C.c_int32),
                ("hash_log2", C.c_int32), ("batch_cells", C.c_int32), ("ttl_seconds", C.c_double),
                ("semantics", C.c_int32), ("max_working_sets", C.c_int32)]


class Status(C.Structure):
    _fields_ = [("code", C.c_int32), ("cell", C.c_int64), ("steps", C.c_int64), ("keys_emitted", C.c_int64),
                ("arena_peak", C.c_int32), ("evolve_ms", C.c_double), ("raster_ms", C.c_double), ("launches", C.c_int64),
                ("build", C.c_int32), ("libm_sensitive", C.c_int32), ("alg_flops", C.c_double)]


class Ab200Error(RuntimeError):
    def __init__(self, code, message=""):
        self.code = code
        RuntimeError.__init__(self, "%s (%d): %s" % (E_NAMES.get(code, "error"), code, message))


def library_path():
    if os.environ.get("AB200_LIBRARY"):      # development hook: a differently tuned build of the same sources
        return os.environ["AB200_LIBRARY"]
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libariadne_b200.so")


_lib = None


def load_library():
    """Load libariadne_b200.so; raises if it has not been built (no fallback of any kind)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not os.path.exists(path):
        raise FileNotFoundError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'`" % path)
    lib = C.CDLL(path)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    P = C.POINTER
    lib.ab200_default_settings.argtypes = [P(Settings)]; lib.ab200_default_settings.restype = None
    lib.ab200_ctx_create.argtypes = [P(vp), C.c_int]; lib.ab200_ctx_create.restype = C.c_int
    lib.ab200_ctx_destroy.argtypes = [vp]; lib.ab200_ctx_destroy.restype = None
    lib.ab200_last_error.argtypes = [vp]; lib.ab200_last_error.restype = C.c_char_p
    lib.ab200_set_system_text.argtypes = [vp, C.c_char_p]; lib.ab200_set_system_text.restype = C.c_int
    lib.ab200_system_info.argtypes = [vp, P(i32), P(i32)]; lib.ab200_system_info.restype = C.c_int
    lib.ab200_location_index.argtypes = [vp, C.c_char_p]; lib.ab200_location_index.restype = C.c_int
    lib.ab200_set_settings.argtypes = [vp, P(Settings)]; lib.ab200_set_settings.restype = C.c_int
    lib.ab200_set_location_tuning.argtypes = [vp, i32, dbl, P(dbl), P(dbl), P(dbl)]; lib.ab200_set_location_tuning.restype = C.c_int
    lib.ab200_set_accuracy.argtypes = [vp, i32]; lib.ab200_set_accuracy.restype = C.c_int
    lib.ab200_analyser_discretise_constraint.argtypes = [vp, C.c_char_p, P(vp)]; lib.ab200_analyser_discretise_constraint.restype = C.c_int
    lib.ab200_analyser_upper_reach_evolve_from.argtypes = [vp, vp, dbl, i32, C.c_int, P(vp), P(vp), P(Status)]; lib.ab200_analyser_upper_reach_evolve_from.restype = C.c_int
    lib.ab200_analyser_outer_chain_reach_from.argtypes = [vp, vp, i32, P(vp), P(Status)]; lib.ab200_analyser_outer_chain_reach_from.restype = C.c_int
    lib.ab200_taylor_model_op.argtypes = [vp, i32, i32, i32, P(C.c_uint8), P(dbl), dbl, i32, P(C.c_uint8), P(dbl), dbl, i32, i32, dbl, P(dbl), P(dbl), P(vp)]
    lib.ab200_taylor_model_op.restype = C.c_int
    lib.ab200_interval_op.argtypes = [vp, i32, P(dbl), P(dbl)]; lib.ab200_interval_op.restype = C.c_int
    lib.ab200_set_accuracy_for_domain.argtypes = [vp, i32, P(dbl)]; lib.ab200_set_accuracy_for_domain.restype = C.c_int
    lib.ab200_upper_reach_evolve.argtypes = [vp, u64, P(i32), P(dbl), dbl, i32, C.c_int, C.c_int, P(vp), P(vp), P(Status)]
    lib.ab200_upper_reach_evolve.restype = C.c_int
    lib.ab200_cellset_view.argtypes = [vp, P(P(u64)), P(u64)]; lib.ab200_cellset_view.restype = C.c_int
    lib.ab200_cellset_free.argtypes = [vp]; lib.ab200_cellset_free.restype = None
    lib.ab200_key_decode.argtypes = [vp, u64, P(i32), P(i64)]; lib.ab200_key_decode.restype = C.c_int
    lib.ab200_reach_evolve_boxes.argtypes = [vp, u64, P(i32), P(dbl), dbl, i32, C.c_int, C.c_int, P(vp), P(vp), P(Status)]
    lib.ab200_reach_evolve_boxes.restype = C.c_int
    lib.ab200_boxlist_view.argtypes = [vp, P(P(u64)), P(P(dbl)), P(P(i32))]; lib.ab200_boxlist_view.restype = C.c_int
    lib.ab200_boxlist_free.argtypes = [vp]; lib.ab200_boxlist_free.restype = None
    lib.ab200_flow_bounds.argtypes = [vp, u64, P(i32), P(dbl), dbl, P(dbl), P(dbl), P(i32)]; lib.ab200_flow_bounds.restype = C.c_int
    lib.ab200_analyser_create.argtypes = [P(vp), vp, P(dbl), i32]; lib.ab200_analyser_create.restype = C.c_int
    lib.ab200_analyser_destroy.argtypes = [vp]; lib.ab200_analyser_destroy.restype = None
    lib.ab200_analyser_info.argtypes = [vp, P(dbl), P(dbl), P(dbl), P(dbl)]; lib.ab200_analyser_info.restype = C.c_int
    lib.ab200_analyser_set_lock_time.argtypes = [vp, dbl]; lib.ab200_analyser_set_lock_time.restype = C.c_int
    lib.ab200_analyser_upper_reach_evolve.argtypes = [vp, u64, P(i32), P(dbl), dbl, i32, C.c_int, P(vp), P(vp), P(Status)]
    lib.ab200_analyser_upper_reach_evolve.restype = C.c_int
    lib.ab200_analyser_lower_reach_evolve.argtypes = [vp, u64, P(i32), P(dbl), dbl, i32, P(vp), P(vp), P(Status)]
    lib.ab200_analyser_lower_reach_evolve.restype = C.c_int
    lib.ab200_analyser_epsilon_lower_chain_reach.argtypes = [vp, u64, P(i32), P(dbl), i32, C.c_char_p, P(vp), P(dbl), P(i32), P(Status)]
    lib.ab200_analyser_epsilon_lower_chain_reach.restype = C.c_int
    lib.ab200_analyser_lower_reach_evolve_constraint.argtypes = [vp, C.c_char_p, dbl, i32, P(vp), P(vp), P(Status)]
    lib.ab200_analyser_lower_reach_evolve_constraint.restype = C.c_int
    lib.ab200_analyser_epsilon_lower_chain_reach_constraint.argtypes = [vp, C.c_char_p, i32, C.c_char_p, P(vp), P(dbl), P(i32), P(Status)]
    lib.ab200_analyser_epsilon_lower_chain_reach_constraint.restype = C.c_int
    lib.ab200_reach_evolve_enclosures.argtypes = [vp, u64, P(i32), P(dbl), vp, dbl, i32, C.c_int, C.c_int, i32, P(vp), P(vp), P(vp), P(Status)]
    lib.ab200_reach_evolve_enclosures.restype = C.c_int
    lib.ab200_enclosures_view.argtypes = [vp, P(u64), P(P(i32)), P(P(u64)), P(u64)]; lib.ab200_enclosures_view.restype = C.c_int
    lib.ab200_enclosures_model.argtypes = [vp, u64, i32, P(i32), P(i32), P(dbl), P(C.c_uint8), P(dbl)]; lib.ab200_enclosures_model.restype = C.c_int
    lib.ab200_enclosures_select.argtypes = [vp, u64, P(u64), P(vp)]; lib.ab200_enclosures_select.restype = C.c_int
    lib.ab200_enclosures_free.argtypes = [vp]; lib.ab200_enclosures_free.restype = None
    lib.ab200_enclosures_create.argtypes = [vp, u64, P(i32), P(i32), P(i32), P(dbl), P(C.c_uint8), P(dbl), P(vp)]; lib.ab200_enclosures_create.restype = C.c_int
    lib.ab200_ctx_get_settings.argtypes = [vp, P(Settings)]; lib.ab200_ctx_get_settings.restype = C.c_int
    lib.ab200_analyser_outer_chain_reach.argtypes = [vp, u64, P(i32), P(dbl), i32, P(vp), P(Status)]
    lib.ab200_analyser_outer_chain_reach.restype = C.c_int
    lib.ab200_analyser_discretise.argtypes = [vp, u64, P(i32), P(dbl), P(vp)]; lib.ab200_analyser_discretise.restype = C.c_int
    lib.ab200_analyser_cells_evolved.argtypes = [vp]; lib.ab200_analyser_cells_evolved.restype = i64
    lib.ab200_analyser_enumerate.argtypes = [vp, vp, C.c_int, P(u64), P(i32), P(dbl)]; lib.ab200_analyser_enumerate.restype = C.c_int
    lib.ab200_probe_fp64.argtypes = [vp, P(dbl)]; lib.ab200_probe_fp64.restype = C.c_int
    lib.ab200_probe_fp64_fma.argtypes = [vp, P(dbl)]; lib.ab200_probe_fp64_fma.restype = C.c_int
    lib.ab200_integrator_flow_bounds.argtypes = [vp, u64, P(i32), P(dbl), dbl, P(dbl), P(dbl), P(i32)]; lib.ab200_integrator_flow_bounds.restype = C.c_int
    lib.ab200_integrator_flow.argtypes = [vp, i32, P(dbl), dbl, i32, P(dbl), P(dbl), P(vp)]; lib.ab200_integrator_flow.restype = C.c_int
    lib.ab200_taylor_models_view.argtypes = [vp, i32, P(i32), P(i32), P(dbl), P(P(C.c_uint8)), P(P(dbl))]; lib.ab200_taylor_models_view.restype = C.c_int
    lib.ab200_taylor_models_free.argtypes = [vp]; lib.ab200_taylor_models_free.restype = None
    lib.ab200_comm_unique_id.argtypes = [C.c_char_p]; lib.ab200_comm_unique_id.restype = C.c_int
    lib.ab200_comm_init.argtypes = [vp, i32, i32, C.c_char_p]; lib.ab200_comm_init.restype = C.c_int
    lib.ab200_comm_destroy.argtypes = [vp]; lib.ab200_comm_destroy.restype = None
    lib.ab200_comm_info.argtypes = [vp, P(i32), P(i32)]; lib.ab200_comm_info.restype = C.c_int
    lib.ab200_system_check.argtypes = [C.c_char_p, P(i32), P(i32), P(i32), C.c_char_p, i32]; lib.ab200_system_check.restype = C.c_int
    _lib = lib
    return lib


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int32))


def pack_boxes(boxes, dim):
    """[[lo0,hi0,lo1,hi1,...], ...] -> contiguous float64 array with BOX_STRIDE doubles per box."""
    b = np.asarray(boxes, dtype=np.float64).reshape(-1, 2 * dim)
    out = np.zeros((b.shape[0], BOX_STRIDE), dtype=np.float64)
    out[:, :2 * dim] = b
    return np.ascontiguousarray(out)


class CellSet(object):
    """Sorted, duplicate-free packed cell keys of a HybridDenotableSet (host copy)."""

    def __init__(self, keys, lib=None, handle=None):
        self.keys = keys
        self._lib = lib
        self._handle = handle        # the library object (kept for HybridReachabilityAnalyser.enumerate), freed with this one

    def __len__(self):
        return int(self.keys.shape[0])

    def release(self):
        if self._handle is not None and self._lib is not None:
            self._lib.ab200_cellset_free(self._handle)
        self._handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def _take_cellset(lib, handle):
    keys = C.POINTER(C.c_uint64)(); n = C.c_uint64(0)
    lib.ab200_cellset_view(handle, C.byref(keys), C.byref(n))
    arr = np.ctypeslib.as_array(keys, shape=(n.value,)).copy() if n.value else np.zeros((0,), dtype=np.uint64)
    return CellSet(arr, lib, handle)


def _take_boxlist(lib, handle, n_cells):
    off = C.POINTER(C.c_uint64)(); bx = C.POINTER(C.c_double)(); lc = C.POINTER(C.c_int32)()
    lib.ab200_boxlist_view(handle, C.byref(off), C.byref(bx), C.byref(lc))
    offsets = np.ctypeslib.as_array(off, shape=(n_cells + 1,)).copy()
    total = int(offsets[-1])
    boxes = np.ctypeslib.as_array(bx, shape=(total, BOX_STRIDE)).copy() if total else np.zeros((0, BOX_STRIDE))
    locs = np.ctypeslib.as_array(lc, shape=(total,)).copy() if total else np.zeros((0,), dtype=np.int32)
    lib.ab200_boxlist_free(handle)
    return offsets, boxes, locs


class Enclosures(object):
    """A list of TaylorSet enclosures (ab200_enclosures) as the evolution returns its final sets: the hand-over between
    the laps of LowerReachEpsilonWorker (include/workers.h:299-302)."""

    def __init__(self, lib, handle):
        self.lib = lib; self._handle = handle
        n = C.c_uint64(0); lc = C.POINTER(C.c_int32)(); uo = C.POINTER(C.c_uint64)(); nu = C.c_uint64(0)
        lib.ab200_enclosures_view(handle, C.byref(n), C.byref(lc), C.byref(uo), C.byref(nu))
        self.locs = np.ctypeslib.as_array(lc, shape=(n.value,)).copy() if n.value else np.zeros((0,), dtype=np.int32)
        self.unit_offsets = np.ctypeslib.as_array(uo, shape=(nu.value + 1,)).copy() if uo else np.zeros((1,), dtype=np.uint64)

    def __len__(self):
        return int(self.locs.shape[0])

    def model(self, i, k):
        """Coordinate model k of enclosure i: (exponents [terms, nargs] uint8, coefficients [terms], error)."""
        na = C.c_int32(0); nt = C.c_int32(0); err = C.c_double(0.0)
        rc = self.lib.ab200_enclosures_model(self._handle, i, k, C.byref(na), C.byref(nt), C.byref(err), None, None)
        if rc:
            raise Ab200Error(rc, "no such enclosure model")
        ex = np.zeros((nt.value, na.value), dtype=np.uint8); co = np.zeros((nt.value,), dtype=np.float64)
        self.lib.ab200_enclosures_model(self._handle, i, k, C.byref(na), C.byref(nt), C.byref(err),
                                        ex.ctypes.data_as(C.POINTER(C.c_uint8)), _dptr(co))
        return ex, co, err.value

    def select(self, idx):
        idx = np.ascontiguousarray(np.asarray(idx, dtype=np.uint64))
        out = C.c_void_p()
        rc = self.lib.ab200_enclosures_select(self._handle, idx.shape[0], idx.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(out))
        if rc:
            raise Ab200Error(rc, "enclosure index out of range")
        return Enclosures(self.lib, out)

    def release(self):
        if self._handle is not None:
            self.lib.ab200_enclosures_free(self._handle); self._handle = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def system_check(system_text):
    """Parse a system text on the host (no GPU): returns (n_locations, n_events, n_ops) or raises Ab200Error."""
    lib = load_library()
    nl = C.c_int32(0); ne = C.c_int32(0); no = C.c_int32(0)
    err = C.create_string_buffer(512)
    rc = lib.ab200_system_check(system_text.encode(), C.byref(nl), C.byref(ne), C.byref(no), err, 512)
    if rc != 0:
        raise Ab200Error(rc, err.value.decode())
    return nl.value, ne.value, no.value


class Evolver(object):
    """Device evolver for one system (the role of ImageSetHybridEvolver behind EvolverInterface)."""

    def __init__(self, system_text, device=0, settings=None):
        self.lib = load_library()
        self.handle = C.c_void_p()
        self.device = device
        rc = self.lib.ab200_ctx_create(C.byref(self.handle), device)
        if rc != 0:
            self.handle = None
            raise Ab200Error(rc, "no usable sm_100 GPU (this library has no CPU path)")
        self._check(self.lib.ab200_set_system_text(self.handle, system_text.encode()))
        n = C.c_int32(0)
        self.lib.ab200_system_info(self.handle, C.byref(n), None)
        dims = (C.c_int32 * n.value)()
        self.lib.ab200_system_info(self.handle, C.byref(n), dims)
        self.n_locations = n.value
        self.dims = list(dims)
        self.settings = Settings()
        self.lib.ab200_default_settings(C.byref(self.settings))
        if settings:
            for k, v in settings.items():
                setattr(self.settings, k, v)
            self._check(self.lib.ab200_set_settings(self.handle, C.byref(self.settings)))

    def _check(self, rc):
        if rc != 0:
            raise Ab200Error(rc, self.lib.ab200_last_error(self.handle).decode())

    def close(self):
        if self.handle:
            self.lib.ab200_ctx_destroy(self.handle)
            self.handle = None

    def location(self, name):
        i = self.lib.ab200_location_index(self.handle, name.encode())
        if i < 0:
            raise KeyError(name)
        return i

    def get_settings(self):
        """The context's current settings (the analyser may have raised arena_words, see ab200_analyser.cu: with_arena_growth)."""
        s = Settings()
        self._check(self.lib.ab200_ctx_get_settings(self.handle, C.byref(s)))
        return s

    def set_settings(self, **kw):
        for k, v in kw.items():
            setattr(self.settings, k, v)
        self._check(self.lib.ab200_set_settings(self.handle, C.byref(self.settings)))

    def tune_settings(self, loc, max_step, ref_widths, grid_origin, grid_lengths):
        """ImageSetHybridEvolver::tune_settings for one location (src/hybrid_evolver-image.cpp:99-113)."""
        a = [np.ascontiguousarray(np.asarray(v, dtype=np.float64)) for v in (ref_widths, grid_origin, grid_lengths)]
        self._check(self.lib.ab200_set_location_tuning(self.handle, loc, float(max_step), _dptr(a[0]), _dptr(a[1]), _dptr(a[2])))

    def set_accuracy(self, accuracy):
        self._check(self.lib.ab200_set_accuracy(self.handle, accuracy))

    def reach_evolve(self, locs, boxes, time, max_events, ignore_activations=False, backward=False):
        """Batch EvolverInterface::reach_evolve: returns (reach, final) as (offsets, boxes, locs) triples."""
        locs = np.ascontiguousarray(np.asarray(locs, dtype=np.int32))
        boxes = np.ascontiguousarray(np.asarray(boxes, dtype=np.float64))
        n = locs.shape[0]
        r = C.c_void_p(); f = C.c_void_p(); st = Status()
        rc = self.lib.ab200_reach_evolve_boxes(self.handle, n, _iptr(locs), _dptr(boxes), float(time), int(max_events),
                                               int(bool(ignore_activations)), int(bool(backward)), C.byref(r), C.byref(f), C.byref(st))
        self.last_status = st
        self._check(rc)
        return _take_boxlist(self.lib, r, n), _take_boxlist(self.lib, f, n)

    def reach_evolve_enclosures(self, locs, boxes, time, max_events, initial=None, ignore_activations=False, backward=False, max_enclosure_words=0):
        """EvolverInterface::reach_evolve with the final sets handed back as TaylorSets (and, with `initial`, started from
        TaylorSets): returns (reach boxlist, final boxlist, Enclosures)."""
        if initial is None:
            locs = np.ascontiguousarray(np.asarray(locs, dtype=np.int32))
            boxes = np.ascontiguousarray(np.asarray(boxes, dtype=np.float64))
            n = locs.shape[0]; lp = _iptr(locs); bp = _dptr(boxes); ih = None
        else:
            n = len(initial); lp = None; bp = None; ih = initial._handle
        r = C.c_void_p(); f = C.c_void_p(); e = C.c_void_p(); st = Status()
        rc = self.lib.ab200_reach_evolve_enclosures(self.handle, n, lp, bp, ih, float(time), int(max_events), int(bool(ignore_activations)),
                                                    int(bool(backward)), int(max_enclosure_words), C.byref(r), C.byref(f), C.byref(e), C.byref(st))
        self.last_status = st
        self._check(rc)
        return _take_boxlist(self.lib, r, n), _take_boxlist(self.lib, f, n), Enclosures(self.lib, e)

    def enclosures_create(self, sets):
        """Enclosures from TaylorSets given as [(location index, [(exponents [terms, nargs], coefficients [terms], error), ...])]."""
        locs = []; nargs = []; nterms = []; errors = []; exps = []; coefs = []
        for loc, models in sets:
            locs.append(int(loc)); nargs.append(int(np.asarray(models[0][0]).reshape(len(models[0][1]), -1).shape[1]) if len(models[0][1]) else 0)
            for ex, co, err in models:
                co = np.asarray(co, dtype=np.float64); ex = np.asarray(ex, dtype=np.uint8).reshape(co.shape[0], -1)
                nterms.append(co.shape[0]); errors.append(float(err)); exps.append(ex.reshape(-1)); coefs.append(co)
        locs = np.asarray(locs, dtype=np.int32); nargs = np.asarray(nargs, dtype=np.int32); nterms = np.asarray(nterms, dtype=np.int32)
        errors = np.asarray(errors, dtype=np.float64)
        exps = np.ascontiguousarray(np.concatenate(exps) if exps else np.zeros((0,), dtype=np.uint8)).astype(np.uint8)
        coefs = np.ascontiguousarray(np.concatenate(coefs) if coefs else np.zeros((0,), dtype=np.float64))
        out = C.c_void_p()
        rc = self.lib.ab200_enclosures_create(self.handle, locs.shape[0], _iptr(locs), _iptr(nargs), _iptr(nterms), _dptr(errors),
                                              exps.ctypes.data_as(C.POINTER(C.c_uint8)), _dptr(coefs), C.byref(out))
        self._check(rc)
        return Enclosures(self.lib, out)

    def upper_reach_evolve(self, locs, boxes, time, max_events, ignore_activations=False, backward=False):
        """UpperReachEvolveWorker::get_result(): (reach CellSet, evolve CellSet)."""
        locs = np.ascontiguousarray(np.asarray(locs, dtype=np.int32))
        boxes = np.ascontiguousarray(np.asarray(boxes, dtype=np.float64))
        r = C.c_void_p(); e = C.c_void_p(); st = Status()
        rc = self.lib.ab200_upper_reach_evolve(self.handle, locs.shape[0], _iptr(locs), _dptr(boxes), float(time), int(max_events),
                                               int(bool(ignore_activations)), int(bool(backward)), C.byref(r), C.byref(e), C.byref(st))
        self.last_status = st
        self._check(rc)
        return _take_cellset(self.lib, r), _take_cellset(self.lib, e)

    def flow_bounds(self, locs, boxes, hmax):
        locs = np.ascontiguousarray(np.asarray(locs, dtype=np.int32))
        boxes = np.ascontiguousarray(np.asarray(boxes, dtype=np.float64))
        n = locs.shape[0]
        h = np.zeros((n,), dtype=np.float64); b = np.zeros((n, BOX_STRIDE), dtype=np.float64); s = np.zeros((n,), dtype=np.int32)
        self._check(self.lib.ab200_flow_bounds(self.handle, n, _iptr(locs), _dptr(boxes), float(hmax), _dptr(h), _dptr(b), _iptr(s)))
        return h, b, s

    def decode(self, keys):
        out = np.zeros((len(keys), 1 + max(self.dims)), dtype=np.int64)
        loc = C.c_int32(0); idx = (C.c_int64 * MAX_DIM)()
        for k, key in enumerate(keys):
            self.lib.ab200_key_decode(self.handle, int(key), C.byref(loc), idx)
            out[k, 0] = loc.value
            for i in range(self.dims[0]):
                out[k, 1 + i] = idx[i]
        return out

    def integrator_flow_bounds(self, loc, box, hmax):
        """IntegratorBase::flow_bounds(vf, box, hmax) -> (h, bounding box) for the dynamics of a location."""
        n = self.dims[loc]
        b = pack_boxes([box], n)
        h = np.zeros((1,), dtype=np.float64); out = np.zeros((1, 2 * MAX_DIM), dtype=np.float64); st = np.zeros((1,), dtype=np.int32)
        locs = np.array([loc], dtype=np.int32)
        self._check(self.lib.ab200_integrator_flow_bounds(self.handle, 1, locs.ctypes.data_as(C.POINTER(C.c_int32)), b.ctypes.data_as(C.POINTER(C.c_double)), hmax,
                                                          h.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data_as(C.POINTER(C.c_double)), st.ctypes.data_as(C.POINTER(C.c_int32))))
        if st[0] != 0:
            raise Ab200Error(int(st[0]), "flow_bounds")
        return float(h[0]), out[0, :2 * n].copy()

    def integrator_flow(self, loc, box, hmax, temporal_order):
        """TaylorIntegrator(temporal_order).flow(vf, box, hmax) -> (h, domain, [(exponents[n_terms][n+1], coefficients, error)] per
        state variable)."""
        n = self.dims[loc]
        b = np.ascontiguousarray(np.asarray(box, dtype=np.float64))
        h = C.c_double(0); dom = np.zeros((2 * (n + 1),), dtype=np.float64); m = C.c_void_p()
        self._check(self.lib.ab200_integrator_flow(self.handle, loc, b.ctypes.data_as(C.POINTER(C.c_double)), hmax, temporal_order, C.byref(h),
                                                   dom.ctypes.data_as(C.POINTER(C.c_double)), C.byref(m)))
        models = []
        try:
            for i in range(n):
                na = C.c_int32(0); nt = C.c_int32(0); err = C.c_double(0); ex = C.POINTER(C.c_uint8)(); cf = C.POINTER(C.c_double)()
                self.lib.ab200_taylor_models_view(m, i, C.byref(na), C.byref(nt), C.byref(err), C.byref(ex), C.byref(cf))
                e = np.ctypeslib.as_array(ex, shape=(nt.value, na.value)).copy() if nt.value else np.zeros((0, na.value), dtype=np.uint8)
                v = np.ctypeslib.as_array(cf, shape=(nt.value,)).copy() if nt.value else np.zeros((0,), dtype=np.float64)
                models.append((e, v, err.value))
        finally:
            self.lib.ab200_taylor_models_free(m)
        return h.value, dom, models

    TM_OPS = {"add": 0, "sub": 1, "mul": 2, "addf": 3, "subf": 4, "mulf": 5, "divf": 6, "addi": 7, "subi": 8, "muli": 9, "divi": 10, "neg": 11,
              "intersection": 12, "split": 13, "antiderivative": 14, "rec": 15, "sqrt": 16, "abs": 17, "max": 18, "range": 19, "exp": 20, "log": 21, "sin": 22, "cos": 23}
    IV_OPS = {"add": 0, "sub": 1, "mul": 2, "div": 3, "addf": 4, "subf": 5, "mulf": 6, "divf": 7, "fsub": 8, "fdiv": 9, "sqr": 10, "sqrt": 11,
              "pow": 12, "rec": 13, "abs": 14, "neg": 15, "max": 16, "min": 17, "exp": 18, "log": 19, "sin": 20, "cos": 21}

    def taylor_model_op(self, op, nargs, x, y=None, k=0, flag=0, s=0.0, interval=(0.0, 0.0)):
        """One TaylorModel operator / function on the device. x, y: (exponents[n][nargs] uint8, coefficients[n], error).
        Returns (exponents, coefficients, error) -- or the range interval for op "range"."""
        def unpack(m):
            if m is None:
                return 0, None, None, 0.0, None
            e = np.ascontiguousarray(np.asarray(m[0], dtype=np.uint8).reshape(-1, nargs)); v = np.ascontiguousarray(np.asarray(m[1], dtype=np.float64))
            n = int(v.shape[0])
            return n, (e.ctypes.data_as(C.POINTER(C.c_uint8)) if n else None), (v.ctypes.data_as(C.POINTER(C.c_double)) if n else None), float(m[2]), (e, v)
        nx, xe, xv, xerr, keepx = unpack(x); ny, ye, yv, yerr, keepy = unpack(y)
        iv = (C.c_double * 2)(float(interval[0]), float(interval[1])); rng = (C.c_double * 2)(0.0, 0.0); m = C.c_void_p()
        self._check(self.lib.ab200_taylor_model_op(self.handle, self.TM_OPS[op], nargs, nx, xe, xv, xerr, ny, ye, yv, yerr, int(k), int(flag), float(s), iv, rng, C.byref(m)))
        try:
            na = C.c_int32(0); nt = C.c_int32(0); err = C.c_double(0); ex = C.POINTER(C.c_uint8)(); cf = C.POINTER(C.c_double)()
            self.lib.ab200_taylor_models_view(m, 0, C.byref(na), C.byref(nt), C.byref(err), C.byref(ex), C.byref(cf))
            e = np.ctypeslib.as_array(ex, shape=(nt.value, na.value)).copy() if nt.value else np.zeros((0, nargs), dtype=np.uint8)
            v = np.ctypeslib.as_array(cf, shape=(nt.value,)).copy() if nt.value else np.zeros((0,), dtype=np.float64)
        finally:
            self.lib.ab200_taylor_models_free(m)
        if op == "range":
            return (rng[0], rng[1])
        return e, v, err.value

    def interval_ops(self, ops):
        """Batch of Interval operations: ops = [(name, m, (al, au), (bl, bu))] -> array [n][3] = lower, upper, status."""
        a = np.ascontiguousarray(np.asarray([[self.IV_OPS[o], m, x[0], x[1], y[0], y[1]] for o, m, x, y in ops], dtype=np.float64).reshape(-1, 6))
        out = np.zeros((a.shape[0], 3), dtype=np.float64)
        self._check(self.lib.ab200_interval_op(self.handle, a.shape[0], _dptr(a), _dptr(out)))
        return out

    def init_comm(self, dist):
        """Join the ranks of an initialised torch.distributed process group into the library's NCCL communicator: rank 0's
        ncclUniqueId is broadcast as bytes; afterwards every lap of this evolver is sharded over the ranks."""
        import torch
        world, rank = dist.get_world_size(), dist.get_rank()
        buf = C.create_string_buffer(128)
        if rank == 0:
            self._check(self.lib.ab200_comm_unique_id(buf))
        dev = torch.device("cuda", self.device) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
        dist.broadcast(t, 0)
        raw = bytes(t.cpu().tolist())
        self._check(self.lib.ab200_comm_init(self.handle, world, rank, raw))

    def probe_fp64(self, fma=False):
        """Measured FP64 rate in GFLOP/s: separately rounded multiply + add chains (what interval arithmetic can issue), or
        fused multiply-adds counted as two operations (the nominal peak)."""
        g = C.c_double(0)
        self._check((self.lib.ab200_probe_fp64_fma if fma else self.lib.ab200_probe_fp64)(self.handle, C.byref(g)))
        return g.value


class HybridReachabilityAnalyser(object):
    """HybridReachabilityAnalyser(system, domain, accuracy) (src/reachability_analyser.cpp:83-100)."""

    def __init__(self, system_text, domain, accuracy, device=0, settings=None):
        self.evolver = Evolver(system_text, device=device, settings=settings)
        self.lib = self.evolver.lib
        dim = self.evolver.dims[0]
        if isinstance(domain, dict):
            dom = np.zeros((self.evolver.n_locations, BOX_STRIDE))
            for name, box in domain.items():
                dom[self.evolver.location(name), :2 * dim] = box
        else:
            dom = np.tile(pack_boxes([domain], dim), (self.evolver.n_locations, 1))
        dom = np.ascontiguousarray(dom)
        self.handle = C.c_void_p()
        self.evolver._check(self.lib.ab200_analyser_create(C.byref(self.handle), self.evolver.handle, _dptr(dom), int(accuracy)))
        self.accuracy = accuracy
        self.dim = dim
        lock = C.c_double(0)
        L = self.evolver.n_locations
        self.grid_origin = np.zeros((L, MAX_DIM)); self.grid_lengths = np.zeros((L, MAX_DIM)); self.max_step = np.zeros((L,))
        self.lib.ab200_analyser_info(self.handle, C.byref(lock), _dptr(self.grid_origin), _dptr(self.grid_lengths), _dptr(self.max_step))
        self.lock_to_grid_time = lock.value
        self.last_status = None

    def close(self):
        if self.handle:
            self.lib.ab200_analyser_destroy(self.handle)
            self.handle = None
        self.evolver.close()

    def set_lock_to_grid_time(self, t):
        self.evolver._check(self.lib.ab200_analyser_set_lock_time(self.handle, float(t)))
        self.lock_to_grid_time = float(t)

    def _initial(self, initial_set):
        locs, boxes = [], []
        for name, box in initial_set.items():
            locs.append(self.evolver.location(name) if isinstance(name, str) else int(name))
            boxes.append(box)
        return np.ascontiguousarray(np.asarray(locs, dtype=np.int32)), pack_boxes(boxes, self.dim)

    def upper_reach_evolve(self, initial_set, time, max_events=5, mince=False):
        locs, boxes = self._initial(initial_set)
        r = C.c_void_p(); e = C.c_void_p(); st = Status()
        rc = self.lib.ab200_analyser_upper_reach_evolve(self.handle, locs.shape[0], _iptr(locs), _dptr(boxes), float(time), int(max_events),
                                                        int(bool(mince)), C.byref(r), C.byref(e), C.byref(st))
        self.last_status = st
        self.evolver._check(rc)
        return self._wrap(r), self._wrap(e)

    def upper_reach(self, initial_set, time, max_events=5, mince=False):
        return self.upper_reach_evolve(initial_set, time, max_events, mince)[0]

    def upper_evolve(self, initial_set, time, max_events=5, mince=False):
        return self.upper_reach_evolve(initial_set, time, max_events, mince)[1]

    def lower_reach_evolve(self, initial_set, time, max_events=5):
        """HybridReachabilityAnalyser::lower_reach_evolve (src/reachability_analyser.cpp:230-275)."""
        locs, boxes = self._initial(initial_set)
        r = C.c_void_p(); e = C.c_void_p(); st = Status()
        rc = self.lib.ab200_analyser_lower_reach_evolve(self.handle, locs.shape[0], _iptr(locs), _dptr(boxes), float(time), int(max_events),
                                                        C.byref(r), C.byref(e), C.byref(st))
        self.last_status = st
        self.evolver._check(rc)
        return self._wrap(r), self._wrap(e)

    def lower_reach_evolve_constraint(self, constraint_text, time, max_events=5):
        """lower_reach_evolve from a constraint-set initial set (the text discretise_constraint takes)."""
        r = C.c_void_p(); e = C.c_void_p(); st = Status()
        rc = self.lib.ab200_analyser_lower_reach_evolve_constraint(self.handle, constraint_text.encode(), float(time), int(max_events),
                                                                   C.byref(r), C.byref(e), C.byref(st))
        self.last_status = st
        self.evolver._check(rc)
        return self._wrap(r), self._wrap(e)

    def lower_reach(self, initial_set, time, max_events=5):
        return self.lower_reach_evolve(initial_set, time, max_events)[0]

    def lower_evolve(self, initial_set, time, max_events=5):
        return self.lower_reach_evolve(initial_set, time, max_events)[1]

    def epsilon_lower_chain_reach(self, initial_set, prune=True, substituted_system_text=None):
        """HybridReachabilityAnalyser::epsilon_lower_chain_reach (src/reachability_analyser.cpp:497-627) without a constraint
        set: returns (reach CellSet, epsilon [n_locations, dim], laps)."""
        r = C.c_void_p(); st = Status(); laps = C.c_int32(0)
        eps = np.zeros((self.evolver.n_locations, MAX_DIM), dtype=np.float64)
        sub = substituted_system_text.encode() if substituted_system_text is not None else None
        if isinstance(initial_set, str):              # a constraint-set initial set (the text discretise_constraint takes)
            rc = self.lib.ab200_analyser_epsilon_lower_chain_reach_constraint(self.handle, initial_set.encode(), int(bool(prune)), sub,
                                                                              C.byref(r), _dptr(eps), C.byref(laps), C.byref(st))
        else:
            locs, boxes = self._initial(initial_set)
            rc = self.lib.ab200_analyser_epsilon_lower_chain_reach(self.handle, locs.shape[0], _iptr(locs), _dptr(boxes), int(bool(prune)), sub,
                                                                   C.byref(r), _dptr(eps), C.byref(laps), C.byref(st))
        self.last_status = st
        self.evolver._check(rc)
        return self._wrap(r), eps, laps.value

    def outer_chain_reach(self, initial_set, backward=False):
        locs, boxes = self._initial(initial_set)
        r = C.c_void_p(); st = Status()
        rc = self.lib.ab200_analyser_outer_chain_reach(self.handle, locs.shape[0], _iptr(locs), _dptr(boxes), int(bool(backward)), C.byref(r), C.byref(st))
        self.last_status = st
        self.evolver._check(rc)
        return self._wrap(r)

    def discretise(self, initial_set):
        """The discretised initial set (outer_intersection_with) as a CellSet."""
        locs, boxes = self._initial(initial_set)
        r = C.c_void_p()
        self.evolver._check(self.lib.ab200_analyser_discretise(self.handle, locs.shape[0], _iptr(locs), _dptr(boxes), C.byref(r)))
        return self._wrap(r)

    def discretise_constraint(self, text):
        """outer_intersection_with(HybridBoundedConstraintSet): the initial cells of a constraint-set initial set."""
        r = C.c_void_p()
        self.evolver._check(self.lib.ab200_analyser_discretise_constraint(self.handle, text.encode(), C.byref(r)))
        return self._wrap(r)

    def upper_reach_evolve_from(self, initial_cells, time, max_events=5, mince=False):
        r = C.c_void_p(); e = C.c_void_p(); st = Status()
        rc = self.lib.ab200_analyser_upper_reach_evolve_from(self.handle, initial_cells._handle, float(time), int(max_events), int(bool(mince)),
                                                             C.byref(r), C.byref(e), C.byref(st))
        self.last_status = st
        self.evolver._check(rc)
        return self._wrap(r), self._wrap(e)

    def outer_chain_reach_from(self, initial_cells, backward=False):
        r = C.c_void_p(); st = Status()
        rc = self.lib.ab200_analyser_outer_chain_reach_from(self.handle, initial_cells._handle, int(bool(backward)), C.byref(r), C.byref(st))
        self.last_status = st
        self.evolver._check(rc)
        return self._wrap(r)

    def cells_evolved(self):
        return int(self.lib.ab200_analyser_cells_evolved(self.handle))

    def _wrap(self, handle):
        return _take_cellset(self.lib, handle)

    def enumerate(self, cellset, mince=False):
        """Cells of a set as the reference iterates them: (locs, boxes[count, 2*dim])."""
        n = C.c_uint64(0)
        self.evolver._check(self.lib.ab200_analyser_enumerate(self.handle, cellset._handle, int(bool(mince)), C.byref(n), None, None))
        locs = np.zeros((n.value,), dtype=np.int32); boxes = np.zeros((n.value, BOX_STRIDE))
        if n.value:
            self.evolver._check(self.lib.ab200_analyser_enumerate(self.handle, cellset._handle, int(bool(mince)), C.byref(n), _iptr(locs), _dptr(boxes)))
        return locs, boxes[:, :2 * self.dim]

    def free(self, cellset):
        cellset.release()

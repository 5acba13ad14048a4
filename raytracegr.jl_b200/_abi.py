"""ctypes mirror of include/raytracegr_cuda.h (the POD structs of the C ABI).

Kept free of any library loading so that both the product binding (this package) and the
test-side oracle binding (tests/oracle_lib.py) can share the struct layouts.
"""
import ctypes as C

RTGR_MINKOWSKI, RTGR_KERR_SCHILD = 0, 1
RTGR_R_AS_WRITTEN, RTGR_R_CORRECTED = 0, 1
RTGR_PLANE, RTGR_SPHERE = 0, 1
RTGR_TILE_W = RTGR_TILE_H = 32
RTGR_MAX_OBJECTS = 16
RTGR_IPC_HANDLE_BYTES = 64   # size of the handle rtgr_frame_create exports (a cudaIpcMemHandle_t)
RTGR_USER_METRIC_BASE = 16   # rtgr_params.metric >= this: a metric compiled with rtgr_metric_compile

STATUS_EVENT, STATUS_LAMBDA_END, STATUS_MAXITERS, STATUS_DT_MIN, STATUS_NONFINITE = range(5)

#: eps(Float64)^(3/4), the reference's reltol = abstol (src/RayTraceGR.jl:485)
REFERENCE_TOL = (2.0 ** -52) ** 0.75


class rtgr_object(C.Structure):
    _fields_ = [("kind", C.c_int32), ("_pad", C.c_int32), ("time", C.c_double),
                ("pos", C.c_double * 4), ("vel", C.c_double * 4), ("radius", C.c_double)]


class rtgr_params(C.Structure):
    _fields_ = [("metric", C.c_int32), ("r_formula", C.c_int32), ("M", C.c_double), ("a", C.c_double),
                ("lambda0", C.c_double), ("lambda1", C.c_double), ("reltol", C.c_double),
                ("abstol", C.c_double), ("hit_threshold", C.c_double),
                ("interp_points", C.c_int32), ("maxiters", C.c_int32)]


class rtgr_camera(C.Structure):
    _fields_ = [("pos", C.c_double * 4), ("widthx", C.c_double * 4), ("widthy", C.c_double * 4),
                ("normal", C.c_double * 4), ("ni", C.c_int32), ("nj", C.c_int32)]


class rtgr_stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("rhs_evals", C.c_uint64), ("steps_accepted", C.c_uint64),
                ("steps_rejected", C.c_uint64), ("kernel_ms", C.c_double), ("total_ms", C.c_double),
                ("drain_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def default_params(metric, M=1.0, a=0.0, r_formula=RTGR_R_AS_WRITTEN, tol=REFERENCE_TOL,
                   lambda0=0.0, lambda1=100.0, hit_threshold=0.01, interp_points=10, maxiters=100000):
    """The reference's hard-coded values (src:275-276, :485, :497, :519) as a params struct."""
    return rtgr_params(int(metric), int(r_formula), float(M), float(a), float(lambda0), float(lambda1),
                       float(tol), float(tol), float(hit_threshold), int(interp_points), int(maxiters))


def make_objects(objs):
    """Marshal a list of (kind, dict) host objects into a flat rtgr_object array."""
    arr = (rtgr_object * max(1, len(objs)))()
    for i, o in enumerate(objs):
        arr[i] = o
    return arr

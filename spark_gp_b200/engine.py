"""`ProjectedProcessEngine`: the Python face of the C-ABI (one context = one GPU).

Mirrors the reference seam `ProjectedGaussianProcessHelper` (commons/ProjectedGaussianProcessHelper.scala):
  getMatrixKmnKnmAndVectorKmny(experts, activeSet)  ->  begin / accumulate* / finish
  getMagicVector(kernel, G, b, ...)                 ->  magic()
and `GaussianProjectedProcessRawPredictor.predict` (commons/GaussianProcessCommons.scala:121-125) -> predict().
Error codes are mapped back to the reference's exception types."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as N
from .kernels import Kernel


class NotPositiveDefiniteException(Exception):
    """ProjectedGaussianProcessHelper.scala:9-11."""


class TrainingVectorsNotInitializedException(Exception):
    """kernel/Kernel.scala:116-117 (raised here for call-order violations)."""


class MatrixSingularException(Exception):
    """breeze MatrixSingularException (commons/util/logDetAndInv.scala:27-28)."""


class SgpError(RuntimeError):
    pass


class OperandRangeError(SgpError):
    """SGP_E_RANGE: the tcgen05 int8 path cannot represent these coordinates; rerun in SGP_PREC_F64."""


def _make_desc(kernel: Kernel, d: int):
    terms = kernel.flatten()
    arr = (N.KernelTerm * len(terms))()
    keep = []
    for i, t in enumerate(terms):
        arr[i].type = t["type"]
        arr[i].scale = t["scale"]
        arr[i].sigma = t.get("sigma", 0.0)
        if t["type"] == N.SGP_TERM_ARD:
            beta = np.ascontiguousarray(t["beta"], dtype=np.float64)
            if len(beta) != d:
                raise ValueError("ARDRBFKernel has %d betas but the data has %d features" % (len(beta), d))
            keep.append(beta)
            arr[i].beta = beta.ctypes.data_as(C.POINTER(C.c_double))
    desc = N.KernelDesc(len(terms), 0, arr)
    return desc, (arr, keep)


class ProjectedProcessEngine:
    def __init__(self, device: int = 0):
        self._lib = N.load()
        self._h = C.c_void_p()
        rc = self._lib.sgp_ctx_create(C.byref(self._h), device)
        if rc != N.SGP_OK:
            raise SgpError("sgp_ctx_create failed (%d): %s" % (rc, self._lib.sgp_last_error(None).decode()))
        self.m = self.d = 0
        self._device = device
        self._has_comm = False

    # ---- a small per-device pool: creating a context costs ~20 ms (streams, cuBLAS, cuSOLVER) and a fresh context
    #      re-allocates every grow-only workspace; an estimator that fits repeatedly gets its contexts back from here ----
    _idle = {}
    _IDLE_MAX = 4

    @classmethod
    def acquire(cls, device: int = 0):
        pool = cls._idle.get(device)
        if pool:
            eng = pool.pop()
            eng.set_precision(N.SGP_PREC_AUTO)
            return eng
        return cls(device)

    def release(self):
        """Hand the context back to the pool (or destroy it when the pool is full or it carries a communicator)."""
        if not self._h:
            return
        pool = ProjectedProcessEngine._idle.setdefault(self._device, [])
        if self._has_comm or len(pool) >= ProjectedProcessEngine._IDLE_MAX:
            self.close()
        else:
            pool.append(self)

    def close(self):
        if self._h:
            self._lib.sgp_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc == N.SGP_OK:
            return
        msg = self._lib.sgp_last_error(self._h).decode()
        if rc == N.SGP_E_NOT_PD:
            raise NotPositiveDefiniteException(msg)
        if rc == N.SGP_E_STATE:
            raise TrainingVectorsNotInitializedException(msg)
        if rc == N.SGP_E_SINGULAR:
            raise MatrixSingularException(msg)
        if rc == N.SGP_E_BADARG:
            raise ValueError(msg)
        if rc == N.SGP_E_RANGE:
            raise OperandRangeError(msg)
        raise SgpError("sgp error %d: %s" % (rc, msg))

    # ---- configuration ---------------------------------------------------------------------------
    def set_precision(self, mode: int):
        self._check(self._lib.sgp_set_precision(self._h, mode))

    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(N.SGP_UNIQUE_ID_BYTES)
        rc = N.load().sgp_comm_unique_id(buf)
        if rc != N.SGP_OK:
            raise SgpError("sgp_comm_unique_id failed")
        return buf.raw

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(unique_id, N.SGP_UNIQUE_ID_BYTES)
        self._check(self._lib.sgp_comm_init(self._h, buf, rank, nranks))
        self._has_comm = True

    # ---- getMatrixKmnKnmAndVectorKmny ---------------------------------------------------------------
    def begin(self, kernel: Kernel, active_set):
        Z = np.ascontiguousarray(active_set, dtype=np.float64)
        if Z.ndim != 2:
            raise ValueError("activeSet must be m x d")
        self.m, self.d = Z.shape
        desc, keep = _make_desc(kernel, self.d)
        self._check(self._lib.sgp_stats_begin(self._h, C.byref(desc), N.ptr(Z), self.m, self.d))
        del keep

    def accumulate(self, X, y):
        """One shard of points in host memory.  X: n x d (float32 or float64), y: n."""
        X = np.asarray(X)
        if X.dtype != np.float32:
            X = np.asarray(X, dtype=np.float64)
        X = np.ascontiguousarray(X)
        y = np.ascontiguousarray(y, dtype=np.float64)
        if X.ndim != 2 or X.shape[1] != self.d or len(y) != len(X):
            raise ValueError("shard shape mismatch")
        self._check(self._lib.sgp_stats_accumulate(self._h, N.ptr(X), int(X.dtype == np.float32), N.ptr(y), len(X)))

    def accumulate_ptr(self, x_ptr: int, x_is_f32: bool, y_ptr: int, n: int, device: bool):
        fn = self._lib.sgp_stats_accumulate_device if device else self._lib.sgp_stats_accumulate
        self._check(fn(self._h, C.c_void_p(x_ptr), int(x_is_f32), C.c_void_p(y_ptr), n))

    def finish(self, copy_out: bool = True):
        if not copy_out:
            self._check(self._lib.sgp_stats_finish(self._h, None, None))
            return None, None
        G = np.empty((self.m, self.m))
        b = np.empty(self.m)
        self._check(self._lib.sgp_stats_finish(self._h, N.ptr(G), N.ptr(b)))
        return G, b

    def sync(self):
        self._check(self._lib.sgp_sync(self._h))

    def statistics(self, kernel: Kernel, active_set, X, y, shard_points: int = 1 << 22):
        """`getMatrixKmnKnmAndVectorKmny` over host data in shards (PGPH:23 broadcast, PGPH:25-35 seqOp over shards).
        When the tcgen05 int8 path reports SGP_E_RANGE (coordinates outside its operand range, or AUTO's magnitude
        budget exceeded over the whole window) the pass is repeated on the fp64 DMMA kernel -- still on the GPU."""
        def run():
            self.begin(kernel, active_set)
            for s in range(0, len(X), shard_points):
                self.accumulate(X[s:s + shard_points], y[s:s + shard_points])
            return self.finish()
        try:
            return run()
        except OperandRangeError:
            # the int8 Gram refused the shard (coordinates outside its operand range, or scaled norms above AUTO's budget:
            # tiny kernel values, where its fixed-point elements are not parity-grade): the fp64 DMMA kernel takes anything
            self.set_precision(N.SGP_PREC_F64)
            return run()

    # ---- the BCM hyper-parameter objective (GPR:55-68 over all experts, GPC:73-78) ----------------------------------
    def experts_upload(self, X, y, offsets):
        """Experts packed expert-major: expert e owns rows offsets[e]..offsets[e+1]-1."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        self._check(self._lib.sgp_experts_upload(self._h, N.ptr(X), N.ptr(y), N.ptr(off), len(off) - 1, X.shape[1]))
        self._experts_d = X.shape[1]

    def experts_upload_grouped(self, X, y, dataset_size_for_expert: int):
        """Groups the points into experts ON THE DEVICE (GPC:26-31: point i -> expert i % E) while they stream in; X may be
        float32 or float64, row-major as given.  Returns the number of experts."""
        X = np.asarray(X)
        if X.dtype != np.float32:
            X = np.asarray(X, dtype=np.float64)
        X = np.ascontiguousarray(X)
        y = np.ascontiguousarray(y, dtype=np.float64)
        self._check(self._lib.sgp_experts_upload_grouped(self._h, N.ptr(X), int(X.dtype == np.float32), N.ptr(y), len(X),
                                                         X.shape[1], int(dataset_size_for_expert)))
        self._experts_d = X.shape[1]
        return int(np.floor(len(X) / dataset_size_for_expert + 0.5))

    def _hyper_array(self, kernel: Kernel, nterms: int):
        hd = kernel.hyper_descriptors()
        arr = (N.Hyper * max(len(hd), 1))()
        coefs = []
        for i, h in enumerate(hd):
            arr[i].kind = h["kind"]
            arr[i].term = h.get("term", 0)
            arr[i].dim = h.get("dim", 0)
            arr[i].value = h.get("value", 0.0)
            if h["kind"] == N.SGP_HYPER_SCALE:
                cf = np.zeros(nterms)
                for t, v in h["coef"].items():
                    cf[t] = v
                coefs.append(cf)
                arr[i].coef = cf.ctypes.data_as(C.POINTER(C.c_double))
        return arr, len(hd), coefs

    def laplace_nll(self, kernel: Kernel, tol: float):
        """Binary classification: (-log Z summed over experts, its gradient).  Updates the device-resident latent
        modes f of the uploaded experts (warm start), as GaussianProcessClassifier.likelihoodAndGradient does."""
        desc, keep = _make_desc(kernel, self._experts_d)
        arr, nh, coefs = self._hyper_array(kernel, desc.n_terms)
        val = C.c_double()
        grad = np.zeros(max(nh, 1))
        self._check(self._lib.sgp_laplace_nll(self._h, C.byref(desc), arr, nh, float(tol), C.byref(val), N.ptr(grad)))
        del keep, coefs
        return val.value, grad[:nh]

    def experts_f(self, n: int):
        f = np.empty(n)
        self._check(self._lib.sgp_experts_get_f(self._h, N.ptr(f)))
        return f

    def bcm_nll(self, kernel: Kernel):
        """(sum over experts of 1/2 y'K^-1 y + 1/2 log|K|, gradient w.r.t. kernel.getHyperparameters())."""
        desc, keep = _make_desc(kernel, self._experts_d)
        hd = kernel.hyper_descriptors()
        nterms = desc.n_terms
        arr = (N.Hyper * max(len(hd), 1))()
        coefs = []
        for i, h in enumerate(hd):
            arr[i].kind = h["kind"]
            arr[i].term = h.get("term", 0)
            arr[i].dim = h.get("dim", 0)
            arr[i].value = h.get("value", 0.0)
            if h["kind"] == N.SGP_HYPER_SCALE:
                cf = np.zeros(nterms)
                for t, v in h["coef"].items():
                    cf[t] = v
                coefs.append(cf)
                arr[i].coef = cf.ctypes.data_as(C.POINTER(C.c_double))
        nll = C.c_double()
        grad = np.zeros(max(len(hd), 1))
        self._check(self._lib.sgp_bcm_nll(self._h, C.byref(desc), arr, len(hd), C.byref(nll), N.ptr(grad)))
        del keep, coefs
        return nll.value, grad[:len(hd)]

    # ---- getMagicVector -------------------------------------------------------------------------------
    def magic(self, G=None, b=None, copy_out: bool = True):
        mv = np.empty(self.m) if copy_out else None
        mm = np.empty((self.m, self.m)) if copy_out else None
        if G is not None:
            G = np.ascontiguousarray(G, dtype=np.float64)
            b = np.ascontiguousarray(b, dtype=np.float64)
        self._check(self._lib.sgp_magic(self._h, N.ptr(G) if G is not None else None,
                                        N.ptr(b) if b is not None else None,
                                        N.ptr(mv) if copy_out else None, N.ptr(mm) if copy_out else None))
        return mv, mm

    # ---- predict --------------------------------------------------------------------------------------
    def set_magic(self, vector, matrix):
        """Installs a caller-supplied (vector, symmetric matrix) for `predict`:  mean = k.v,  var = selfKernel + k M k^T
        (the per-point quadratic forms of the greedy active-set provider, ActiveSetProvider.scala:109-113)."""
        v = np.ascontiguousarray(vector, dtype=np.float64)
        M = np.ascontiguousarray(matrix, dtype=np.float64)
        if v.shape != (self.m,) or M.shape != (self.m, self.m):
            raise ValueError("set_magic: vector must be m, matrix m x m")
        self._check(self._lib.sgp_set_magic(self._h, N.ptr(v), N.ptr(M)))

    def predict(self, X, with_variance: bool = True):
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[None, :]
        mean = np.empty(len(X))
        var = np.empty(len(X)) if with_variance else None
        self._check(self._lib.sgp_predict(self._h, N.ptr(X), len(X), N.ptr(mean),
                                          N.ptr(var) if with_variance else None))
        return mean, var

    def cross_kernel(self, X):
        """kernel.crossKernel(test) with the active set as training vectors: len(test) x m."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[None, :]
        K = np.empty((len(X), self.m))
        self._check(self._lib.sgp_cross_kernel(self._h, N.ptr(X), len(X), N.ptr(K)))
        return K

    def greedy_active_set(self, kernel: Kernel, X, y, n_experts: int, first_index: int, m_target: int):
        """Row indices of the points GreedilyOptimizingActiveSetProvider selects (ActiveSetProvider.scala:58-139), computed
        with rank-1 updates on the device (`sgp_greedy_active_set`): O(n m) per round instead of a statistics pass."""
        X = np.ascontiguousarray(X, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        desc, keep = _make_desc(kernel, X.shape[1])
        idx = np.empty(int(m_target), dtype=np.int64)
        self._check(self._lib.sgp_greedy_active_set(self._h, C.byref(desc), N.ptr(X), N.ptr(y), len(X), X.shape[1],
                                                    int(n_experts), int(first_index), int(m_target), N.ptr(idx)))
        del keep
        return idx

    def kmn_sweep(self, X):
        """fp32 K[i][j] = k(x_i, z_j) (n x m) through the tensor-core sweep kernel (one non-Eye term, d <= 32)."""
        X = np.asarray(X)
        if X.dtype != np.float32:
            X = np.asarray(X, dtype=np.float64)
        X = np.ascontiguousarray(X)
        K = np.empty((len(X), self.m), dtype=np.float32)
        self._check(self._lib.sgp_kmn_sweep(self._h, N.ptr(X), int(X.dtype == np.float32), len(X), N.ptr(K)))
        return K

    def kmn_sweep_device(self, x_ptr: int, x_is_f32: bool, n: int, k_ptr: int):
        self._check(self._lib.sgp_kmn_sweep_device(self._h, C.c_void_p(x_ptr), int(x_is_f32), n, C.c_void_p(k_ptr)))

    def debug_i8_tile(self):
        """Arms (first call) / reads back the SGP_PREC_I8 debug dump: (T [128x64] fp32, words [128x64] uint32)."""
        T = np.zeros((128, 64), dtype=np.float32)
        w = np.zeros((128, 64), dtype=np.uint32)
        self._check(self._lib.sgp_debug_i8_tile(self._h, N.ptr(T), N.ptr(w)))
        return T, w

    def debug_i8_timeline(self):
        out = np.zeros(2560 + 148 * 32, dtype=np.int64)
        self._check(self._lib.sgp_debug_i8_timeline(self._h, N.ptr(out)))
        self.i8_progress = out[2560:].reshape(148, 32)        # [CTA][unit / 128] clock64 of the Gram issuer
        return out[:2560].reshape(2, 5, 32, 8)

    # ---- introspection ---------------------------------------------------------------------------------
    def last_path(self) -> int:
        """SGP_PREC_F64 / SGP_PREC_F64_STRICT / SGP_PREC_I8: the kernel the last statistics launch ran."""
        return int(self._lib.sgp_last_path(self._h))

    def last_bcm_path(self) -> int:
        """0: on-chip Cholesky objective kernel; 1: global-memory LU path (large experts / not positive definite)."""
        return int(self._lib.sgp_last_bcm_path(self._h))

    def last_tail_path(self) -> int:
        """1: the tail ran on Cholesky factors (PD check included); 0: the reference's dsyevd + LU sequence; -1: never ran."""
        return int(self._lib.sgp_last_tail_path(self._h))

    def launch_count(self) -> int:
        return int(self._lib.sgp_launch_count(self._h))

    def event_record(self, slot: int):
        self._check(self._lib.sgp_event_record(self._h, slot))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_double()
        self._check(self._lib.sgp_event_elapsed_ms(self._h, a, b, C.byref(ms)))
        return ms.value

    def gram_kernel_time(self):
        ms = C.c_double()
        n = C.c_int64()
        self._check(self._lib.sgp_gram_kernel_time(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

"""TEST INFRASTRUCTURE: run the CUDA kernel *source* on the host (one-lane warps).

``numbakit-ode_b200/csrc/nbk_program.cu`` is compiled with g++ against
``cuda_host_shim.h`` so that the logic of the device code -- trajectory queue, retire,
t_bound guard, controller, dense output, multistep start-up -- is covered by
``pytest -m "not gpu"``.  It checks the same source the GPU runs, but it is NOT a product
path: nothing under ``nbkode_b200`` imports it, and it says nothing about GPU-only aspects
(warp refill with 32 lanes, MUFU seeds, FMA contraction), which the ``-m gpu`` parity tests cover.
"""

from __future__ import annotations

import ctypes as C
import math
import os
import re
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "numbakit-ode_b200", "csrc")
BUILD = os.path.join(HERE, "_build")

_CTYPES = {
    "long long": C.c_longlong, "int": C.c_int, "double": C.c_double,
}


def _kargs_struct():
    """ctypes mirror of struct nbk_kargs, parsed from the header so it cannot drift."""
    text = open(os.path.join(CSRC, "nbk_kargs.h")).read()
    body = text[text.index("struct nbk_kargs {") + len("struct nbk_kargs {"): text.index("};")]
    body = re.sub(r"//[^\n]*", "", body)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if not decl:
            continue
        m = re.match(r"(const\s+)?(unsigned long long|long long|double|int)\s*(.*)", decl)
        assert m, decl
        base = m.group(2)
        for name in m.group(3).split(","):
            name = name.strip()
            arr = re.match(r"(\w+)\[(\d+)\]$", name)
            if name.startswith("*"):
                fields.append((name.lstrip("* "), C.c_void_p))
            elif arr:
                fields.append((arr.group(1), _CTYPES[base] * int(arr.group(2))))
            else:
                fields.append((name, _CTYPES[base]))
    return type("KArgs", (C.Structure,), {"_fields_": fields})


KArgs = _kargs_struct()

RHS_TYPES = {"lorenz": ("nbk::rhs_lorenz", 3, 3), "vdp": ("nbk::rhs_vdp", 2, 1)}
METHOD_IDS = {
    "RungeKutta23": 23, "RungeKutta45": 45, "DOP853": 853, "ForwardEuler": 1, "AdamsBashforth1": 1, "AdamsBashforth2": 2,
    "AdamsBashforth3": 3, "AdamsBashforth4": 4, "AdamsBashforth5": 5,
    "Runge2": 202, "Runge3": 203, "Heun3": 213, "RungeKutta4": 204, "RungeKutta3_8": 238,
}
METHOD_IDS.update({"BackwardEuler": 11, **{f"AdamsMoulton{k}": 10 + k for k in range(1, 6)}, **{f"BDF{k}": 30 + k for k in range(1, 7)}})
ERK_STAGES = {202: 2, 203: 4, 213: 3, 204: 4, 238: 4}


def implicit_order(mid):  # size of A for AdamsMoulton1..5 (11..15) / BDF1..6 (31..36), else 0
    if 11 <= mid <= 15:
        return 1 if mid <= 12 else mid - 11
    if 31 <= mid <= 36:
        return mid - 30
    return 0
_libs = {}


def _program(method, rhs, n, npar, events=None):
    """events: (n_events, CUDA C source defining nbk_event) -> also builds the run_events kernel (kernel 3)"""
    if rhs == "decay":
        # typedef through a forced include: a -D value cannot carry the template comma portably
        rtype, tag = f"nbk::rhs_decay<{n},{npar}>", f"m{method}_decay_{n}_{npar}"
    else:
        rtype, tag = RHS_TYPES[rhs][0], f"m{method}_{rhs}"
    if events:
        import hashlib

        tag += "_ev" + hashlib.sha1(events[1].encode()).hexdigest()[:10]
    if tag in _libs:
        return _libs[tag]
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, f"libemul_{tag}.so")
    deps = [os.path.join(CSRC, f) for f in ("nbk_device.cuh", "nbk_kargs.h", "nbk_program.cu", "nbk_dop853_coeffs.h")]
    deps += [os.path.join(HERE, f) for f in ("nbk_emul.cpp", "cuda_host_shim.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        pre = os.path.join(BUILD, f"pre_{tag}.h")
        with open(pre, "w") as fh:
            fh.write(f"#define NBK_METHOD {method}\n#define NBK_RHS_T {rtype}\n#define NBK_TAG {tag}\n")
            if events:
                fh.write(f"#define NBK_NEVENTS {events[0]}\n#define NBK_EVENT static inline double\n{events[1]}\n")
        subprocess.check_call([
            "g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-ffp-contract=off",
            "-I", CSRC, "-I", HERE, "-D__CUDACC_RTC__", "-include", pre,
            os.path.join(HERE, "nbk_emul.cpp"), "-o", so + f".{os.getpid()}.tmp",
        ])
        os.replace(so + f".{os.getpid()}.tmp", so)  # atomic: pytest-xdist workers may build the same program
    lib = C.CDLL(so)
    fn = getattr(lib, f"emul_{tag}")
    fn.argtypes = [C.c_int, C.POINTER(KArgs), C.c_longlong]
    _libs[tag] = fn
    return fn


class EmulEnsemble:
    """Host mirror of what nbkode_b200.core.Solver does around the C-ABI, on NumPy arrays."""

    def __init__(self, method, rhs, y0, params, *, t0=0.0, h=None, first_step=None, rtol=1e-3, atol=1e-6,
                 min_factor=0.2, max_factor=10.0, safety_factor=0.9, first_stepper=45, params_shared=False):
        self.mid = METHOD_IDS[method] if isinstance(method, str) else method
        y0 = np.ascontiguousarray(np.atleast_2d(np.asarray(y0, float)))
        self.B, self.n = y0.shape
        params = np.ascontiguousarray(np.asarray(params, float))
        params = params.reshape(1, -1) if params_shared else params.reshape(self.B, -1)
        self.np_ = params.shape[1]
        self._rhs_name = rhs
        self.fn = _program(self.mid, rhs, self.n, self.np_)
        adaptive = self.mid in (23, 45, 853)
        self.L = 2 if (adaptive or self.mid in ERK_STAGES) else max(implicit_order(self.mid) or self.mid, 2)
        self.SR = {23: 4, 45: 7, 853: 13, **ERK_STAGES}.get(self.mid, 0)
        B, n = self.B, self.n
        self.ts = np.zeros((self.L, B)); self.ys = np.zeros((self.L, n, B)); self.fs = np.zeros((self.L, n, B))
        self.K = np.zeros((max(self.SR, 1), n, B)); self.h = np.zeros(B)
        self.p = np.zeros((self.np_,) if params_shared else (self.np_, B))
        if params_shared:
            self.p[:] = params[0]
        self.nacc = np.zeros(B, np.int64); self.nrej = np.zeros(B, np.int64)
        self.status = np.zeros(B, np.int32); self.nout = np.zeros(B, np.int32)
        self.counter = np.zeros(2, np.uint64)
        root = {45: 10, 853: 16}.get(self.mid, 6)
        self.base = dict(
            B=B, ts=self.ts.ctypes.data, ys=self.ys.ctypes.data, fs=self.fs.ctypes.data, K=self.K.ctypes.data,
            h=self.h.ctypes.data, params=self.p.ctypes.data, n_acc=self.nacc.ctypes.data, n_rej=self.nrej.ctypes.data,
            status=self.status.ctypes.data, n_out=self.nout.ctypes.data, params_shared=int(params_shared),
            first_stepper=first_stepper or 0, rtol=rtol, atol=atol, min_factor=min_factor, max_factor=max_factor,
            safety=safety_factor, x_lo=0.5 * (safety_factor / max_factor) ** root,
            x_hi=2.0 * (safety_factor / min_factor) ** root, bound=math.inf, max_attempts=2**31 - 1,
            work_counter=self.counter.ctypes.data,
        )
        hh = first_step if adaptive else (1.0 if h is None else h)
        a = KArgs(**self.base)
        a.y0, a.y0_sb, a.y0_sc = y0.ctypes.data, n, 1
        a.p0, a.p0_sb, a.p0_sc = params.ctypes.data, self.np_, 1
        a.t0, a.h_init = t0, (math.nan if hh is None else hh)
        self._keep = (y0, params)
        self.fn(0, C.byref(a), B)
        self._pristine = True  # mirrors nbk_api.cu: the first advance after init may skip loading K / old history

    def run(self, ts, t_bound=math.inf):
        ts = np.ascontiguousarray(np.atleast_1d(np.asarray(ts, float)))
        m = ts.size
        out = np.full((self.B, m, self.n), np.nan)
        a = KArgs(**self.base)
        a.bound, a.tgrid, a.m = t_bound, ts.ctypes.data, m
        a.pristine, self._pristine = int(self._pristine), False
        a.y_out, a.o_sb, a.o_sk, a.o_sc = out.ctypes.data, m * self.n, self.n, 1
        self.counter[:] = 0
        self.fn(1, C.byref(a), self.B)
        return out

    def run_events(self, ts, events, cap=64, t_bound=math.inf):
        """events: nbkode_b200.rhs.CudaEvents (e.g. from rhs.trace_events).  Mirrors Solver.run_events' buffers."""
        fn = _program(self.mid, self._rhs_name, self.n, self.np_, (events.n_events, events.source))
        ts = np.ascontiguousarray(np.atleast_1d(np.asarray(ts, float)))
        m, NE = ts.size, events.n_events
        out = np.full((self.B, m, self.n), np.nan)
        ev_t = np.full((self.B, NE, cap), np.nan); ev_y = np.full((self.B, NE, cap, self.n), np.nan)
        ev_c = np.zeros((self.B, NE), np.int32); t_term = np.full(self.B, np.nan)
        a = KArgs(**self.base)
        a.bound, a.tgrid, a.m = t_bound, ts.ctypes.data, m
        a.pristine, self._pristine = int(self._pristine), False
        a.y_out, a.o_sb, a.o_sk, a.o_sc = out.ctypes.data, m * self.n, self.n, 1
        a.ev_cap, a.ev_terminal = cap, sum(1 << k for k, x in enumerate(events.terminal) if x)
        for k, d in enumerate(events.direction):
            a.ev_dir[k] = d
        a.ev_t, a.ev_y, a.ev_count, a.ev_tterm = ev_t.ctypes.data, ev_y.ctypes.data, ev_c.ctypes.data, t_term.ctypes.data
        self.counter[:] = 0
        fn(3, C.byref(a), self.B)
        return out, ev_t, ev_y, ev_c, t_term

    def step(self, n_steps, bound=math.inf, emit=True):
        t_out = np.full((self.B, max(n_steps, 1)), np.nan)
        y_out = np.full((self.B, max(n_steps, 1), self.n), np.nan)
        a = KArgs(**self.base)
        a.bound, a.n_steps, a.emit = bound, n_steps, int(emit)
        a.pristine, self._pristine = int(self._pristine), False
        a.t_out, a.to_sb = t_out.ctypes.data, max(n_steps, 1)
        a.y_out, a.o_sb, a.o_sk, a.o_sc = y_out.ctypes.data, max(n_steps, 1) * self.n, self.n, 1
        self.counter[:] = 0
        self.fn(2, C.byref(a), self.B)
        return t_out, y_out, self.nout.copy()

    t = property(lambda self: self.ts[-1].copy())
    y = property(lambda self: self.ys[-1].T.copy())
    f = property(lambda self: self.fs[-1].T.copy())


# --------------------------------------------------------------------------------------------------------------------
# CTA-per-trajectory regime (csrc/nbk_cta.cuh) on the host: one block of T POSIX threads (nbk_emul_cta.cpp)
# --------------------------------------------------------------------------------------------------------------------
CTA_RHS_TYPES = {"rd1d": ("nbk::rhs_rd1d_cta", 2), "linear": ("nbk::rhs_linear_cta", -1)}


def cta_shape(mid, rhs, n):
    """Mirror of cta_shape() / cta_block() in csrc/nbk_api.cu -> (components per thread, block, full)."""
    if n > 2048:
        e = (n + 255) // 256
    elif mid == 853:
        e = 1 if n <= 512 else (2 if n <= 1024 else 4)
    elif rhs == "linear":
        e = 1 if n <= 512 else (2 if n <= 1024 else 4)
    else:
        e = 1 if n <= 64 else (2 if n < 128 else 4)
    t = ((n + e - 1) // e + 31) // 32 * 32
    return e, t, int(mid in (23, 45, 853) and t * e == n)


def _program_cta(method, rhs, elems, full, user_src=None, n_params=0, events=None):
    """events: (n_events, CUDA C source defining nbk_event) -> also builds the run_events kernel (kernel 3)"""
    import hashlib

    tag = f"m{method}_{rhs}_e{elems}{'f' if full else ''}"
    extra = []
    if user_src is not None:
        tag += "_u" + hashlib.sha1((user_src + str(n_params)).encode()).hexdigest()[:10]
    if events:
        tag += "_ev" + hashlib.sha1(events[1].encode()).hexdigest()[:10]
    if ("cta", tag) in _libs:
        return _libs[("cta", tag)]
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, f"libemul_cta_{tag}.so")
    if user_src is not None:
        # the translation-unit prelude of nbk_jit_compile() for a CTA user right-hand side
        pre = os.path.join(BUILD, f"pre_cta_{tag}.h")
        with open(pre, "w") as fh:
            fh.write("#define NBK_RHS_I static inline double\n"
                     "namespace nbk { template <int E> struct smem_vec; }\ntypedef nbk::smem_vec<NBK_E> nbk_vec;\n"
                     + (f"#define NBK_CTA_NP {n_params}\n" if 0 < n_params <= 16 else "")
                     + "#define NBK_USER_RHS_CTA 1\n"
                     "NBK_RHS_I nbk_rhs_i(double t, nbk_vec y, const double* p, int i, int n);\n"
                     '#include "nbk_cta.cuh"\n' + user_src + "\n")
        extra = [f'-DEMUL_USER_PRELUDE="{pre}"']
    if events:
        pre_ev = os.path.join(BUILD, f"pre_cta_ev_{tag}.{os.getpid()}.h")
        with open(pre_ev, "w") as fh:
            fh.write(f"#define NBK_NEVENTS {events[0]}\n#define NBK_EVENT static inline double\n{events[1]}\n")
        extra += ["-include", pre_ev]
    deps = [os.path.join(CSRC, f) for f in ("nbk_device.cuh", "nbk_cta.cuh", "nbk_kargs.h", "nbk_program_cta.cu", "nbk_dop853_coeffs.h")]
    deps += [os.path.join(HERE, f) for f in ("nbk_emul_cta.cpp", "cuda_host_shim.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call([
            "g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-ffp-contract=off",
            "-I", CSRC, "-I", HERE, "-D__CUDACC_RTC__", f"-DNBK_METHOD={method}",
            f"-DNBK_RHS_T={'nbk::rhs_user_cta' if user_src is not None else CTA_RHS_TYPES[rhs][0]}",
            f"-DNBK_TAG={tag}", f"-DNBK_E={elems}", "-DNBK_CTA_MAXT=512", f"-DNBK_CTA_FULL={full}", *extra,
            os.path.join(HERE, "nbk_emul_cta.cpp"), "-o", so + f".{os.getpid()}.tmp",
        ])
        os.replace(so + f".{os.getpid()}.tmp", so)  # atomic: pytest-xdist workers may build the same program
    fn = getattr(C.CDLL(so), f"emul_cta_{tag}")
    fn.argtypes = [C.c_int, C.POINTER(KArgs), C.c_int]
    _libs[("cta", tag)] = fn
    return fn


class EmulCtaEnsemble:
    """Host mirror of the Solver around the C-ABI for the CTA regime (trajectory-major state: ts [B][L], ys/fs [B][L][n],
    K [B][rows][n]); same methods as EmulEnsemble."""

    def __init__(self, method, rhs, y0, params, *, t0=0.0, h=None, first_step=None, rtol=1e-3, atol=1e-6,
                 min_factor=0.2, max_factor=10.0, safety_factor=0.9, first_stepper=45, params_shared=False):
        """rhs: a built-in name ("rd1d", "linear") or CUDA C source defining nbk_rhs_i (e.g. rhs.trace_cta(...).source)"""
        self.mid = METHOD_IDS[method]
        y0 = np.ascontiguousarray(np.atleast_2d(np.asarray(y0, float)))
        self.B, self.n = y0.shape
        params = np.ascontiguousarray(np.asarray(params, float))
        params = params.reshape(1, -1) if params_shared else params.reshape(self.B, -1)
        self.np_ = params.shape[1]
        user_src = rhs if "nbk_rhs_i" in rhs else None
        if user_src is not None:  # shape rule of csrc/nbk_api.cu for source: a loop means dense coupling
            rhs = "linear" if ("for (" in user_src or "for(" in user_src or "while" in user_src) else "rd1d"
        self._user_src, self._rhs_name = user_src, rhs
        self.E, self.T, self.full = cta_shape(self.mid, rhs, self.n)
        self.fn = _program_cta(self.mid, "user" if user_src is not None else rhs, self.E, self.full, user_src, self.np_)
        adaptive = self.mid in (23, 45, 853)
        self.L = 2 if (adaptive or self.mid in ERK_STAGES) else max(self.mid, 2)
        self.SR = {23: 4, 45: 7, 853: 13, **ERK_STAGES}.get(self.mid, 0)
        B, n = self.B, self.n
        self.ts = np.zeros((B, self.L)); self.ys = np.zeros((B, self.L, n)); self.fs = np.zeros((B, self.L, n))
        self.K = np.zeros((B, max(self.SR, 1), n)); self.h = np.zeros(B)
        self.p = np.zeros((self.np_,) if params_shared else (B, self.np_))
        if params_shared:
            self.p[:] = params[0]
        self.nacc = np.zeros(B, np.int64); self.nrej = np.zeros(B, np.int64)
        self.status = np.zeros(B, np.int32); self.nout = np.zeros(B, np.int32)
        self.counter = np.zeros(2, np.uint64)
        root = {45: 10, 853: 16}.get(self.mid, 6)
        self.base = dict(
            B=B, n=n, n_params=self.np_, ts=self.ts.ctypes.data, ys=self.ys.ctypes.data, fs=self.fs.ctypes.data,
            K=self.K.ctypes.data, h=self.h.ctypes.data, params=self.p.ctypes.data, n_acc=self.nacc.ctypes.data,
            n_rej=self.nrej.ctypes.data, status=self.status.ctypes.data, n_out=self.nout.ctypes.data,
            params_shared=int(params_shared), first_stepper=first_stepper or 0, rtol=rtol, atol=atol,
            min_factor=min_factor, max_factor=max_factor, safety=safety_factor,
            x_lo=0.5 * (safety_factor / max_factor) ** root, x_hi=2.0 * (safety_factor / min_factor) ** root,
            bound=math.inf, max_attempts=2**31 - 1, work_counter=self.counter.ctypes.data,
        )
        hh = first_step if adaptive else (1.0 if h is None else h)
        a = KArgs(**self.base)
        a.y0, a.y0_sb, a.y0_sc = y0.ctypes.data, n, 1
        a.p0, a.p0_sb, a.p0_sc = params.ctypes.data, self.np_, 1
        a.t0, a.h_init = t0, (math.nan if hh is None else hh)
        self._keep = (y0, params)
        assert self.fn(0, C.byref(a), self.T) == 0
        self._pristine = True

    def run(self, ts, t_bound=math.inf):
        ts = np.ascontiguousarray(np.atleast_1d(np.asarray(ts, float)))
        m = ts.size
        out = np.full((self.B, m, self.n), np.nan)
        a = KArgs(**self.base)
        a.bound, a.tgrid, a.m = t_bound, ts.ctypes.data, m
        a.pristine, self._pristine = int(self._pristine), False
        a.y_out, a.o_sb, a.o_sk, a.o_sc = out.ctypes.data, m * self.n, self.n, 1
        self.counter[:] = 0
        assert self.fn(1, C.byref(a), self.T) == 0
        return out

    def step(self, n_steps, bound=math.inf, emit=True):
        t_out = np.full((self.B, max(n_steps, 1)), np.nan)
        y_out = np.full((self.B, max(n_steps, 1), self.n), np.nan)
        a = KArgs(**self.base)
        a.bound, a.n_steps, a.emit = bound, n_steps, int(emit)
        a.pristine, self._pristine = int(self._pristine), False
        a.t_out, a.to_sb = t_out.ctypes.data, max(n_steps, 1)
        a.y_out, a.o_sb, a.o_sk, a.o_sc = y_out.ctypes.data, max(n_steps, 1) * self.n, self.n, 1
        self.counter[:] = 0
        assert self.fn(2, C.byref(a), self.T) == 0
        return t_out, y_out, self.nout.copy()

    def _events_program(self, events):
        user = getattr(self, "_user_src", None)
        return _program_cta(self.mid, "user" if user is not None else self._rhs_name, self.E, self.full, user, self.np_,
                            (events.n_events, events.source))

    def run_events(self, ts, events, cap=64, t_bound=math.inf):
        """events: nbkode_b200.rhs.CudaEvents (e.g. from rhs.trace_events).  Mirrors Solver.run_events' buffers."""
        fn = self._events_program(events)
        ts = np.ascontiguousarray(np.atleast_1d(np.asarray(ts, float)))
        m, NE = ts.size, events.n_events
        out = np.full((self.B, m, self.n), np.nan)
        ev_t = np.full((self.B, NE, cap), np.nan); ev_y = np.full((self.B, NE, cap, self.n), np.nan)
        ev_c = np.zeros((self.B, NE), np.int32); t_term = np.full(self.B, np.nan)
        a = KArgs(**self.base)
        a.bound, a.tgrid, a.m = t_bound, ts.ctypes.data, m
        a.pristine, self._pristine = int(self._pristine), False
        a.y_out, a.o_sb, a.o_sk, a.o_sc = out.ctypes.data, m * self.n, self.n, 1
        a.ev_cap, a.ev_terminal = cap, sum(1 << k for k, x in enumerate(events.terminal) if x)
        for k, d in enumerate(events.direction):
            a.ev_dir[k] = d
        a.ev_t, a.ev_y, a.ev_count, a.ev_tterm = ev_t.ctypes.data, ev_y.ctypes.data, ev_c.ctypes.data, t_term.ctypes.data
        self.counter[:] = 0
        assert fn(3, C.byref(a), self.T) == 0
        return out, ev_t, ev_y, ev_c, t_term

    t = property(lambda self: self.ts[:, -1].copy())
    y = property(lambda self: self.ys[:, -1].copy())
    f = property(lambda self: self.fs[:, -1].copy())


# --------------------------------------------------------------------------------------------------------------------
# Sub-warp regime (csrc/nbk_group.cuh) on the host: the run / step kernels as one warp of 32 POSIX threads
# (nbk_emul_group.cpp); the construction kernel is the CTA regime's
# --------------------------------------------------------------------------------------------------------------------
GROUP_RHS_TYPES = {"rd1d": "nbk::rhs_rd1d_cta", "decay": "nbk::rhs_decay_cta<false>", "decay1": "nbk::rhs_decay_cta<true>"}


def group_shape(n, lanes=None, emax=4):
    """Mirror of group_shape() in csrc/nbk_api.cu -> (lanes per trajectory, components per lane, construction block);
    emax: components per lane, 4 (2 under DOP853)."""
    g = 2
    while g < 32 and (n + g - 1) // g > emax:
        g *= 2
    g = lanes or g
    e = (n + g - 1) // g
    return g, e, ((n + e - 1) // e + 31) // 32 * 32


def _program_group(method, rhs, lanes, elems, user_src=None, n_params=0, events=None, full=0):
    """events: (n_events, CUDA C source defining nbk_event) -> also builds the run_events kernel (kernel 3);
    full: 1 = every lane of a group owns `elems` valid components (n == lanes x elems): no bounds checks compiled in"""
    import hashlib

    tag = f"m{method}_{rhs}_g{lanes}e{elems}{'f' if full else ''}"
    extra = []
    if user_src is not None:
        tag += "_u" + hashlib.sha1((user_src + str(n_params)).encode()).hexdigest()[:10]
    if events:
        tag += "_ev" + hashlib.sha1(events[1].encode()).hexdigest()[:10]
    if ("group", tag) in _libs:
        return _libs[("group", tag)]
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, f"libemul_group_{tag}.so")
    pre = os.path.join(BUILD, f"pre_group_{tag}.{os.getpid()}.h")
    with open(pre, "w") as fh:
        if user_src is not None:
            fh.write("#define NBK_RHS_I static inline double\n"
                     "namespace nbk { template <int E> struct smem_vec; }\ntypedef nbk::smem_vec<NBK_E> nbk_vec;\n"
                     + (f"#define NBK_CTA_NP {n_params}\n" if 0 < n_params <= 16 else "")
                     + "#define NBK_USER_RHS_CTA 1\n"
                     "NBK_RHS_I nbk_rhs_i(double t, nbk_vec y, const double* p, int i, int n);\n"
                     '#include "nbk_cta.cuh"\n' + user_src + "\n#define NBK_RHS_T nbk::rhs_user_cta\n")
        else:
            fh.write(f"#define NBK_RHS_T {GROUP_RHS_TYPES[rhs]}\n")  # (a -D value cannot carry the template brackets portably)
        if events:
            fh.write(f"#define NBK_NEVENTS {events[0]}\n#define NBK_EVENT static inline double\n{events[1]}\n")
    extra = [f'-DEMUL_USER_PRELUDE="{pre}"']
    deps = [os.path.join(CSRC, f) for f in ("nbk_device.cuh", "nbk_cta.cuh", "nbk_group.cuh", "nbk_kargs.h", "nbk_program_cta.cu")]
    deps += [os.path.join(HERE, f) for f in ("nbk_emul_group.cpp", "cuda_host_shim.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call([
            "g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-ffp-contract=off",
            "-I", CSRC, "-I", HERE, "-D__CUDACC_RTC__", f"-DNBK_METHOD={method}", f"-DNBK_TAG={tag}", f"-DNBK_E={elems}",
            f"-DNBK_GROUP={lanes}", "-DNBK_GRP_BLOCK=32", "-DNBK_CTA_MAXT=128", f"-DNBK_CTA_FULL={int(full)}", *extra,
            os.path.join(HERE, "nbk_emul_group.cpp"), "-o", so + f".{os.getpid()}.tmp",
        ])
        os.replace(so + f".{os.getpid()}.tmp", so)  # atomic: pytest-xdist workers may build the same program
    fn = getattr(C.CDLL(so), f"emul_group_{tag}")
    fn.argtypes = [C.c_int, C.POINTER(KArgs), C.c_int]
    _libs[("group", tag)] = fn
    return fn


class EmulGroupEnsemble(EmulCtaEnsemble):
    """The sub-warp regime in one emulated warp: up to 32 / G trajectories in flight, the rest pulled from the queue (or, for
    the fixed-step classes, taken in turn).  RungeKutta23 / RungeKutta45 run the dedicated kernels of nbk_group.cuh, every
    other class of the CTA regime its source with sub-warp teams (NBK_TEAM_G).  `rhs`: "rd1d", "decay" (n rates), "decay1"
    (one rate) or nbk_rhs_i source."""

    def __init__(self, method, rhs, y0, params, lanes=None, **kw):
        n = np.atleast_2d(np.asarray(y0)).shape[1]
        self._shape = group_shape(n, lanes, emax=2 if method == "DOP853" else 4)
        self._rhs = rhs
        super().__init__(method, rhs if "nbk_rhs_i" in rhs else "rd1d", y0, params, **kw)

    def __setattr__(self, name, value):
        if name in ("E", "T", "full") and "_shape" in self.__dict__:
            value = {"E": self._shape[1], "T": self._shape[2], "full": 0}[name]
        if name == "fn" and "_shape" in self.__dict__:
            user = self._rhs if "nbk_rhs_i" in self._rhs else None
            value = _program_group(self.mid, "user" if user else self._rhs, self._shape[0], self._shape[1], user, self.np_,
                                   full=int(self._shape[0] * self._shape[1] == self.n))
        object.__setattr__(self, name, value)

    def _events_program(self, events):
        user = self._rhs if "nbk_rhs_i" in self._rhs else None
        return _program_group(self.mid, "user" if user else self._rhs, self._shape[0], self._shape[1], user, self.np_,
                              (events.n_events, events.source), full=int(self._shape[0] * self._shape[1] == self.n))


# --------------------------------------------------------------------------------------------------------------------
# K4 (csrc/nbk_mma.cuh): shared-matrix linear systems, right-hand side as FP64 DMMA tiles -- host build with the
# fragment exchange emulated (nbk_emul_mma.cpp)
# --------------------------------------------------------------------------------------------------------------------
def _program_mma(method):
    tag = f"m{method}_linear_mma"
    if ("mma", tag) in _libs:
        return _libs[("mma", tag)]
    os.makedirs(BUILD, exist_ok=True)
    so = os.path.join(BUILD, f"libemul_mma_{tag}.so")
    deps = [os.path.join(CSRC, f) for f in ("nbk_device.cuh", "nbk_cta.cuh", "nbk_mma.cuh", "nbk_kargs.h", "nbk_program_mma.cu")]
    deps += [os.path.join(HERE, f) for f in ("nbk_emul_mma.cpp", "cuda_host_shim.h")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call([
            "g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wno-unknown-pragmas", "-ffp-contract=off",
            "-I", CSRC, "-I", HERE, "-D__CUDACC_RTC__", f"-DNBK_METHOD={method}", f"-DNBK_TAG={tag}",
            os.path.join(HERE, "nbk_emul_mma.cpp"), "-o", so + f".{os.getpid()}.tmp",
        ])
        os.replace(so + f".{os.getpid()}.tmp", so)  # atomic: pytest-xdist workers may build the same program
    fn = getattr(C.CDLL(so), f"emul_mma_{tag}")
    fn.argtypes = [C.c_int, C.POINTER(KArgs)]
    _libs[("mma", tag)] = fn
    return fn


class EmulMmaEnsemble(EmulCtaEnsemble):
    """y' = A y, one shared dense A (n <= 128), through the DMMA tile kernel; `A` is given as the matrix itself."""

    def __init__(self, method, y0, A, **kw):
        fn = _program_mma(METHOD_IDS[method])
        self._mma = fn
        super().__init__(method, "linear", y0, np.ascontiguousarray(np.asarray(A, float).T), params_shared=True, **kw)

    def __setattr__(self, name, value):
        if name == "fn" and "_mma" in self.__dict__:
            # the construction kernel is the CTA regime's (256 threads), run / step are the tile kernels: all three
            # live in the K4 program, launched with 256 threads
            mma = self.__dict__["_mma"]
            value = lambda kernel, a, _t: mma(kernel, a)  # noqa: E731
        object.__setattr__(self, name, value)

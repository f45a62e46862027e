"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/bacon_ivp.h declares, and the builder reproduces the reference's validation rules
(src/ivp/rk.rs:168-256, identical in bdf.rs:176-264).  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from bacon_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "bacon_ivp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+)?[A-Za-z_][\w\s\*]*?\b(bacon_\w+)\s*\(", src, flags=re.M)
    return sorted(set(n for n in names if n != "bacon_launch_fn"))


def test_header_and_python_mirror_agree():
    assert declared_functions() == sorted(_abi.EXPORTED_SYMBOLS)


def test_rust_sys_bindings_declare_the_same_abi():
    """rust/bacon-ivp-sys cannot be compiled here (no cargo): at least its extern block names every entry point of the
    header except the nvcc-side plug-in registration, and its #[repr(C)] structs list the header's fields in order."""
    rs = open(os.path.join(ROOT, "rust", "bacon-ivp-sys", "src", "lib.rs")).read()
    assert f"ABI version {_abi.ABI_VERSION}" in rs
    fns = set(re.findall(r"pub fn (bacon_\w+)", rs))
    assert fns == set(declared_functions()) - {"bacon_rhs_register"}
    for struct, mirror in (("bacon_ivp_config", _abi.Config), ("bacon_ivp_result", _abi.Result),
                           ("bacon_ivp_options", _abi.Options), ("bacon_ivp_launch_info", _abi.LaunchInfo)):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % struct, rs, flags=re.S).group(1)
        assert re.findall(r"pub (\w+):", body) == [n for n, _ in mirror._fields_], struct


def test_library_exports_every_declared_symbol(engine):
    from bacon_b200._lib import lib
    L = lib()
    for name in declared_functions():
        assert hasattr(L, name), f"libbacon_ivp.so does not export {name}"
    assert L.bacon_abi_version() == _abi.ABI_VERSION
    for code, name in _abi.STATUS_NAMES.items():
        assert L.bacon_status_name(code).decode() == name


def test_struct_layouts_match_header(tmp_path):
    """The ctypes mirrors against the header itself: a C program compiled from include/bacon_ivp.h prints the sizes and
    field offsets gcc gives the structs."""
    import subprocess
    fields = {"Config": ("bacon_ivp_config", [n for n, _ in _abi.Config._fields_]),
              "Result": ("bacon_ivp_result", [n for n, _ in _abi.Result._fields_]),
              "Options": ("bacon_ivp_options", [n for n, _ in _abi.Options._fields_]),
              "LaunchInfo": ("bacon_ivp_launch_info", [n for n, _ in _abi.LaunchInfo._fields_])}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "bacon_ivp.h"', 'int main(void) {']
    for py, (c, names) in fields.items():
        lines.append(f'printf("{py} size %zu\\n", sizeof({c}));')
        for n in names:
            lines.append(f'printf("{py} {n} %zu\\n", offsetof({c}, {n}));')
    lines.append('printf("abi %d\\n", BACON_IVP_ABI_VERSION); printf("dyn %d\\n", BACON_DIM_DYN);')
    lines.append('printf("stopped %d\\n", (int)BACON_STOPPED_AT_EVENT); return 0; }')
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = str(tmp_path / "layout")
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe], check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout.split("\n")
    seen = 0
    for l in out:
        w = l.split()
        if len(w) == 3 and w[1] == "size":
            assert C.sizeof(getattr(_abi, w[0])) == int(w[2]), l
            seen += 1
        elif len(w) == 3:
            assert getattr(getattr(_abi, w[0]), w[1]).offset == int(w[2]), l
            seen += 1
        elif len(w) == 2:
            assert {"abi": _abi.ABI_VERSION, "dyn": _abi.DIM_DYN, "stopped": _abi.STOPPED_AT_EVENT}[w[0]] == int(w[1]), l
            seen += 1
    assert seen == 4 + sum(len(n) for _, n in fields.values()) + 3


def test_dimension_errors_of_the_constructors(engine):
    """IVPSolver::new / new_dyn with the reference's `Dimension` (src/lib.rs:53-76, ivp.rs:72-88): new() on a Dyn solver
    is StaticOnDynamic, new_dyn(size) on a Const<C> solver is DynamicOnStatic; the matching calls build a solver."""
    E = engine.IVPError
    for cls in (engine.RungeKutta45, engine.BDF6, engine.Adams5, engine.Euler):
        assert cls.new(3).dim() == 3                      # RK45::<U3>::new()
        assert cls.new_dyn(2).dim() == 2                  # RK45::<Dyn>::new_dyn(2)   (ivp.rs:562)
        with pytest.raises(E) as e:
            cls.new(dim_type=cls.DYN)                     # RK45::<Dyn>::new()
        assert e.value.variant == "StaticOnDynamic" and e.value.code == 12
        with pytest.raises(E) as e:
            cls.new_dyn(3, dim_type=3)                    # RK45::<U3>::new_dyn(3)
        assert e.value.variant == "DynamicOnStatic" and e.value.code == 11
    # the messages are the reference's (lib.rs:47-50)
    try:
        engine.RK45.new(dim_type=engine.RK45.DYN)
    except E as e:
        assert "static solver with dynamic dimension" in str(e)


def test_initial_dt_setter(engine):
    E = engine.IVPError
    s = (engine.RK45.new(1).with_maximum_dt(0.1).with_minimum_dt(0.01).with_tolerance(1e-4).with_initial_time(0.0)
         .with_ending_time(1.0))
    assert s._config(0).dt_init == 0.0                    # the reference's (dt_max + dt_min)/2
    assert s.with_initial_dt(0.05)._config(0).dt_init == 0.05
    for bad in (0.0, -1.0, float("nan")):
        with pytest.raises(E) as e:
            s.with_initial_dt(bad)
        assert e.value.variant == "TimeDeltaOOB"


def test_builtin_rhs_registry(engine):
    from bacon_b200._lib import lib
    L = lib()
    want = {"lorenz": (3, 3), "vdp": (2, 1), "robertson": (3, 3), "exp": (1, 0), "decay": (1, 0),
            "quadratic": (1, 0), "cos": (1, 0), "harmonic": (2, 1), "linear4": (4, 16), "linear32": (32, 1024)}
    for name, (dim, npar) in want.items():
        rid = L.bacon_rhs_lookup(name.encode())
        assert rid >= 0, name
        d, p = C.c_int(), C.c_int()
        assert L.bacon_rhs_info(rid, None, C.byref(d), C.byref(p)) == 0
        assert (d.value, p.value) == (dim, npar)
    assert L.bacon_rhs_lookup(b"no_such_rhs") == -1


@pytest.mark.parametrize("cls_name", ["RungeKutta45", "RungeKutta23", "BDF6", "BDF2"])
def test_builder_validation_rules(engine, cls_name):
    cls = getattr(engine, cls_name)
    E = engine.IVPError
    # with_tolerance: tol <= 0 -> ToleranceOOB (rk.rs:168-174)
    for bad in (0.0, -1e-3):
        with pytest.raises(E) as e:
            cls.new(1).with_tolerance(bad)
        assert e.value.variant == "ToleranceOOB"
    # with_maximum_dt / with_minimum_dt: <= 0 -> TimeDeltaOOB (rk.rs:179-210)
    for setter in ("with_maximum_dt", "with_minimum_dt", "with_dt_max", "with_dt_min"):
        with pytest.raises(E) as e:
            getattr(cls.new(1), setter)(0.0)
        assert e.value.variant == "TimeDeltaOOB"
    # ordering: a later max below min pulls min down; a later min above max pushes max up
    s = cls.new(1).with_minimum_dt(0.5).with_maximum_dt(0.1).with_tolerance(1e-3).with_initial_time(0).with_ending_time(1)
    cfg = s._config(0)
    assert (cfg.dt_min, cfg.dt_max) == (0.1, 0.1)
    s = cls.new(1).with_maximum_dt(0.1).with_minimum_dt(0.5).with_tolerance(1e-3).with_initial_time(0).with_ending_time(1)
    cfg = s._config(0)
    assert (cfg.dt_min, cfg.dt_max) == (0.5, 0.5)
    # with_initial_time after end: end <= initial -> TimeStartOOB; with_ending_time: initial >= ending -> TimeEndOOB
    with pytest.raises(E) as e:
        cls.new(1).with_ending_time(1.0).with_initial_time(1.0)
    assert e.value.variant == "TimeStartOOB"
    with pytest.raises(E) as e:
        cls.new(1).with_initial_time(2.0).with_ending_time(1.0)
    assert e.value.variant == "TimeEndOOB"
    with pytest.raises(E) as e:
        cls.new(1).with_start(2.0).with_end(2.0)
    assert e.value.variant == "TimeEndOOB"
    # solve with anything unset -> MissingParameters (rk.rs:249-256)
    full = dict(with_maximum_dt=0.1, with_minimum_dt=0.01, with_tolerance=1e-4, with_initial_time=0.0, with_ending_time=1.0)
    for missing in full:
        s = cls.new(1)
        for k, v in full.items():
            if k != missing:
                getattr(s, k)(v)
        with pytest.raises(E) as e:
            s._config(0)
        assert e.value.variant == "MissingParameters"
    s = cls.new(1)
    for k, v in full.items():
        getattr(s, k)(v)
    with pytest.raises(E) as e:  # no initial conditions
        s.with_derivative("exp").solve()
    assert e.value.variant == "MissingParameters"
    with pytest.raises(E) as e:  # no derivative
        cls.new(1).with_initial_conditions([1.0]).solve_ivp_ensemble([[1.0]])
    assert e.value.variant == "MissingParameters"


def test_validate_config_codes(engine):
    from bacon_b200._lib import lib
    L = lib()
    ok = dict(method=_abi.RK45, dim=3, n_params=3, semantics=0, flags=0, history_capacity=0, dt_min=1e-9, dt_max=0.1,
              tol=1e-8, t_start=0.0, t_end=5.0, max_attempts=0)
    assert L.bacon_ivp_validate(C.byref(_abi.Config(**ok))) == 0
    for patch, code in [(dict(tol=0.0), _abi.E_TOLERANCE_OOB), (dict(dt_min=-1.0), _abi.E_TIME_DELTA_OOB),
                        (dict(dt_min=1.0, dt_max=0.5), _abi.E_TIME_DELTA_OOB), (dict(t_end=0.0), _abi.E_TIME_END_OOB),
                        (dict(method=9), _abi.E_BAD_ARGUMENT), (dict(dim=0), _abi.E_BAD_ARGUMENT),
                        (dict(semantics=5), _abi.E_BAD_ARGUMENT)]:
        assert L.bacon_ivp_validate(C.byref(_abi.Config(**{**ok, **patch}))) == code
        assert L.bacon_last_error()


def test_solve_rejects_mismatched_arguments_before_touching_the_gpu(engine):
    """Argument errors are reported by code without any CUDA call (works on a CPU-only box)."""
    import numpy as np
    s = (engine.RK45.new(3).with_dt_min(1e-9).with_dt_max(0.1).with_tolerance(1e-8).with_start(0).with_end(1)
         .with_derivative("lorenz"))
    with pytest.raises(engine.IVPError) as e:
        s.solve_ivp_ensemble(np.zeros((2, 4)), np.zeros((3, 4)))
    assert e.value.variant == "BadArgument"
    with pytest.raises(engine.IVPError) as e:
        s.solve_ivp_ensemble(np.zeros((3, 4)))  # lorenz needs params
    assert e.value.variant == "MissingParameters"
    with pytest.raises(engine.IVPError) as e:
        engine.RK45.new(2).with_derivative("lorenz")._rhs_info()
    assert e.value.variant == "BadArgument"
    with pytest.raises(engine.IVPError):
        engine.RK45.new(3).with_derivative("nope")


def test_no_cpu_fallback_in_product_package():
    """The product package never imports the oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "bacon_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in text and "from oracle" not in text and "import oracle" not in text, f
                assert not re.search(r'#\s*include\s*[<"][^>"]*oracle', text), f


def test_user_rhs_plugin_builds_and_registers(engine):
    """North-star: user right-hand sides plug in through the same C ABI.  examples/user_rhs.cu is compiled
    against include/bacon_ivp_rhs.cuh for sm_100a (nvcc cross-compiles here) and registers on load."""
    import subprocess
    from bacon_b200._lib import lib
    subprocess.run(["make", "-C", os.path.join(ROOT, "examples")], check=True, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    L = lib()
    C.CDLL(os.path.join(ROOT, "examples", "libuser_rhs.so"), mode=C.RTLD_GLOBAL)
    for name, dim, npar in (("brusselator", 2, 2), ("pendulum", 2, 1)):
        rid = L.bacon_rhs_lookup(name.encode())
        assert rid >= 0
        d, p = C.c_int(), C.c_int()
        L.bacon_rhs_info(rid, None, C.byref(d), C.byref(p))
        assert (d.value, p.value) == (dim, npar)


def test_cpp_facade_validation():
    """include/bacon_ivp.hpp (the compiled-language mirror of the reference builder): validation rules, no GPU."""
    import subprocess
    exe = "/tmp/bacon_cpp_facade_test"
    subprocess.run(["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp_facade_test.cpp"),
                    "-o", exe, "-L" + os.path.join(ROOT, "bacon_b200"), "-lbacon_ivp", "-Wl,-rpath," + os.path.join(ROOT, "bacon_b200")],
                   check=True)
    out = subprocess.run([exe], check=True, capture_output=True, text=True).stdout
    assert out.strip().endswith("ok")


def test_every_kernel_exists_in_exactly_one_object(engine):
    """The library links a fast (FMA) and a strict (-fmad=false) build of the same headers.  A kernel name that
    appears in both objects would be launched from whichever copy the runtime registered last (found once: the
    headline kernel ran without FMA contraction in its right-hand side)."""
    import shutil
    import subprocess
    from bacon_b200 import _lib
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-symbols", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    entries = [l.split()[-1] for l in out.splitlines() if "STT_FUNC" in l and "STO_ENTRY" in l]
    assert len(entries) > 100
    dup = sorted({e for e in entries if entries.count(e) > 1})
    assert not dup, dup[:4]


def test_runtime_compiled_rhs_is_checked_at_registration():
    """bacon_rhs_register_source (NVRTC): a good functor registers and can be looked up, a broken one is rejected with
    the compiler's log as IVPError(UserError) — the analogue of a `Derivative` closure that does not type-check.  No GPU
    needed: compiling is not launching."""
    import bacon_b200 as B
    from bacon_b200 import _abi
    from bacon_b200._lib import lib
    good = """
    struct Brusselator {
        static constexpr int DIM = 2, NPARAM = 2;
        __device__ void operator()(double, const double (&y)[2], const double* p, double (&dy)[2]) const {
            dy[0] = p[0] + y[0] * y[0] * y[1] - (p[1] + 1.0) * y[0];
            dy[1] = p[1] * y[0] - y[0] * y[0] * y[1];
        }
    };"""
    try:
        rid = B.register_rhs_source("brusselator_rtc_abi", "Brusselator", good, 2, 2)
    except B.IVPError as e:
        if e.code == _abi.E_UNSUPPORTED:
            pytest.skip(f"no libnvrtc here: {e}")
        raise
    assert rid >= 0 and lib().bacon_rhs_lookup(b"brusselator_rtc_abi") == rid
    with pytest.raises(B.IVPError) as bad:
        B.register_rhs_source("broken_rtc_abi", "Broken", good.replace("Brusselator", "Broken").replace("p[1] * y[0]", "q[1] * y[0]"), 2, 2)
    assert bad.value.code == _abi.E_USER and "q" in str(bad.value) and "undefined" in str(bad.value)
    with pytest.raises(B.IVPError):  # DIM of the functor and of the registration must agree
        B.register_rhs_source("wrongdim_rtc_abi", "Brusselator", good, 3, 2)
    # the path-query kernels (path_query.cuh) go through NVRTC too, fast and strict (compiled lazily otherwise)
    os.environ["BACON_RTC_EAGER_PATHS"] = "1"
    try:
        assert B.register_rhs_source("brusselator_rtc_paths", "Brusselator", good, 2, 2) >= 0
    finally:
        del os.environ["BACON_RTC_EAGER_PATHS"]


def test_path_query_argument_checks_need_no_gpu():
    """bacon_ivp_sample_paths / bacon_ivp_locate_events reject a solve without history, a missing buffer and an unknown
    direction before anything touches the device."""
    import ctypes as C

    import numpy as np

    from bacon_b200 import RungeKutta45
    from bacon_b200._lib import lib
    L = lib()
    s = (RungeKutta45.new(3).with_minimum_dt(1e-9).with_maximum_dt(0.1).with_tolerance(1e-8).with_initial_time(0.0)
         .with_ending_time(1.0).with_derivative("lorenz"))
    rid = L.bacon_rhs_lookup(b"lorenz")
    y0 = np.ones((3, 2))
    p = np.ones((3, 2))
    hist = np.zeros((2, 8, 4))
    hl = np.zeros(2, dtype=np.uint32)
    out = np.zeros((2, 1, 3))
    t = np.zeros(1)
    res = _abi.Result(hist=hist.ctypes.data, hist_len=hl.ctypes.data)
    cfg = s._config(3)  # history_capacity == 0
    assert L.bacon_ivp_sample_paths(C.byref(cfg), rid, 2, y0.ctypes.data, p.ctypes.data, C.byref(res), 1, t.ctypes.data,
                                    out.ctypes.data) == _abi.E_BAD_ARGUMENT
    cfg = s.with_history(8)._config(3)
    assert L.bacon_ivp_sample_paths(C.byref(cfg), rid, 2, y0.ctypes.data, p.ctypes.data, C.byref(res), 1, None,
                                    out.ctypes.data) == _abi.E_BAD_ARGUMENT
    w = np.ones(3)
    cnt = np.zeros(2, dtype=np.uint32)
    ev = np.zeros((2, 2, 4))
    assert L.bacon_ivp_locate_events(C.byref(cfg), rid, 2, y0.ctypes.data, p.ctypes.data, C.byref(res), w.ctypes.data, 0.0,
                                     2, 2, ev.ctypes.data, cnt.ctypes.data) == _abi.E_BAD_ARGUMENT
    assert L.bacon_ivp_locate_events(C.byref(cfg), rid, 2, y0.ctypes.data, p.ctypes.data, C.byref(res), None, 0.0,
                                     0, 2, ev.ctypes.data, cnt.ctypes.data) == _abi.E_BAD_ARGUMENT
    # n == 0 is a no-op success
    assert L.bacon_ivp_sample_paths(C.byref(cfg), rid, 0, y0.ctypes.data, p.ctypes.data, C.byref(res), 1, t.ctypes.data,
                                    out.ctypes.data) == 0


def test_product_coefficient_tables_are_the_reference_sources():
    """The PRODUCT's tables — the constexpr tableaux the fast kernels are compiled from and the runtime tableaux the
    strict kernels' constant memory is filled with (bacon_b200/csrc/tableaux.cuh, adams.cuh), dumped by a host program
    compiled from those headers (tests/tableau_dump.cu) — against the numbers parsed out of the reference's own source
    text (tests/golden/reference_coefficients.json), bit for bit, in both semantics.  No GPU: the headers' table code is
    host code too.  (tests/test_oracle.py holds the oracle's tables against the same file.)"""
    import json
    import shutil
    import subprocess
    import numpy as np
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = "/tmp/bacon_tableau_dump"
    subprocess.run([nvcc, "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "bacon_b200", "csrc"),
                    os.path.join(ROOT, "tests", "tableau_dump.cu"), "-o", exe], check=True, capture_output=True)
    got = {}
    for line in subprocess.run([exe], check=True, capture_output=True, text=True).stdout.splitlines():
        key, *words = line.split()
        got[key] = np.array([int(w, 16) for w in words], dtype=np.uint64)
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_coefficients.json")))

    def bits(x):
        return np.ascontiguousarray(np.asarray(x, dtype=np.float64)).view(np.uint64).ravel()

    for name, o in (("RK45", 6), ("RK23", 4)):
        listed = np.array(ref[name]["k_coefficients"]["values"]).reshape(o, o)
        corrected = listed.copy()
        if name == "RK45":
            assert corrected[5, 3] == 1859.0 / 4014.0  # as written
            corrected[5, 3] = 1859.0 / 4104.0
        for flavour, want_a, safety in (("fast", corrected, 0.84), ("corrected", corrected, 0.84), ("literal", listed.T, 1.0)):
            assert np.array_equal(got[f"{name}.{flavour}.c"], bits(ref[name]["t_coefficients"]["values"])), (name, flavour)
            assert np.array_equal(got[f"{name}.{flavour}.b"], bits(ref[name]["avg_coefficients"]["values"])), (name, flavour)
            assert np.array_equal(got[f"{name}.{flavour}.e"], bits(ref[name]["error_coefficients"]["values"])), (name, flavour)
            assert np.array_equal(got[f"{name}.{flavour}.A"], bits(want_a)), (name, flavour)
            assert np.array_equal(got[f"{name}.{flavour}.safety"], bits([84.0 / 100.0 if safety == 0.84 else 1.0])), (name, flavour)
        assert ref["RK_safety"]["values"] == [1.0]  # what rk.rs:266-268 computes (D3); REF_LITERAL keeps it
    for name in ("BDF6", "BDF2"):
        assert np.array_equal(got[f"{name}.higher"], bits(ref[name]["higher_coefficients"]["values"]))
        assert np.array_equal(got[f"{name}.lower"], bits(ref[name]["lower_coefficients"]["values"]))
    for name in ("Adams5", "Adams3"):
        assert np.array_equal(got[f"{name}.predictor"], bits(ref[name]["predictor_coefficients"]["values"]))
        assert np.array_equal(got[f"{name}.corrector"], bits(ref[name]["corrector_coefficients"]["values"]))
        assert np.array_equal(got[f"{name}.error"], bits(ref[name]["error_coefficient"]["values"]))

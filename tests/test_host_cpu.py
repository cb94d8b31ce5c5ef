"""CPU tests of the host side: model compilation to the rate table, the C-ABI library loads and exports every symbol
of include/dpomp.h, and compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_case


def test_library_exports_every_declared_symbol(built_lib, dp):
    header = open(os.path.join(ROOT, "include", "dpomp.h")).read()
    declared = set(re.findall(r"\b(dpomp_[a-z0-9_]+)\s*\(", header))
    declared -= {"dpomp_model_desc", "dpomp_status"}
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(built_lib, name), f"{name} declared in include/dpomp.h but not exported"
    assert declared == set(dp._capi.EXPORTED_SYMBOLS), declared ^ set(dp._capi.EXPORTED_SYMBOLS)


def test_model_desc_layout_matches_header(built_lib, dp):
    # sizeof(dpomp_model_desc) as laid out by ctypes must match the C struct: build a model and read it back through
    # the library's own validation (a mismatched layout makes the observation pointers garbage)
    model, y, hmm, theta = load_case(dp, "sis_pooley")
    dm = dp.device_model(hmm)  # dpomp_model_create needs no GPU
    d = dm.compiled.desc
    assert (d.n_compartments, d.n_events, d.n_params, d.n_obs, d.n_obs_vals) == (2, 2, 2, 5, 2)


def test_no_cpu_fallback(built_lib, dp):
    n = C.c_int()
    rc = built_lib.dpomp_device_count(C.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    model, y, hmm, theta = load_case(dp, "sis_pooley")
    with pytest.raises(dp._capi.DpompError) as ei:
        dp.get_particle_filter_lpdf(model, y)(theta)
    assert ei.value.code == -2
    with pytest.raises(dp._capi.DpompError):
        dp.rs_systematic(np.ones(8), u=0.5)


def test_predefined_models_compile_to_rate_tables(dp):
    rng = np.random.default_rng(3)
    cases = [("SI", [10, 1]), ("SIR", [100, 1, 0]), ("SIS", [100, 1]), ("SEI", [10, 1, 1]), ("SEIR", [100, 0, 1, 0]),
             ("SEIS", [10, 1, 1]), ("LOTKA", [70, 70]), ("ROSSMAC", [10, 1, 10, 1])]
    for name, ic in cases:
        for fd in (False, True):
            if fd and name in ("LOTKA", "ROSSMAC"):
                continue
            m = dp.generate_model(name, ic, freq_dep=fd)
            e_n, c = m.m_transition.shape
            tab = dp.compile_rate_table(m.rate_function, e_n, len(m.prior), c)
            for _ in range(20):
                th = rng.uniform(0.01, 1.0, size=len(m.prior)); x = rng.integers(1, 300, size=c)
                want = np.zeros(e_n); m.rate_function(want, th, x)
                assert np.allclose(tab.evaluate(th, x), want, rtol=1e-13), (name, fd)
    # mass-action products keep the reference's association (theta * x_a) * x_b with a < b
    tab = dp.compile_rate_table(dp.generate_model("SIR", [100, 1, 0]).rate_function, 2, 2, 3)
    assert tab.f1[0].tolist() == [1, 0, 0] and tab.f2[0].tolist() == [0, 1, 0]
    assert dp.generate_model("SEIRS", [1, 1, 1, 1]) is None  # src/hmm_examples.jl:205-207


def test_custom_model_of_reference_tests_compiles(dp):
    # test/runtests.jl:74-100
    def sis_rf(output, parameters, population):
        output[0] = parameters[0] * population[0] * population[1]
        output[1] = parameters[1] * population[1]

    def si_gaussian(y, population, theta):
        obs_err = 2
        tmp1 = np.log(1 / (np.sqrt(2 * np.pi) * obs_err))
        tmp2 = 2 * obs_err * obs_err
        obs_diff = y.val[1] - population[1]
        return tmp1 - ((obs_diff * obs_diff) / tmp2)

    prior = dp.UniformProduct([0, 0], [0.1, 0.5])
    model = dp.DPOMPModel("SIS", sis_rf, [100, 1], [[-1, 1], [1, -1]], dp.dmy_obs_fn, si_gaussian, prior, 0)
    y = dp.get_observations(os.path.join(GOLDEN, "pooley.csv"))
    cm = dp.compile_model(model, y)
    assert cm.obs.sigma == pytest.approx(2.0) and cm.obs.xmask.tolist() == [0, 1] and cm.obs.ymask.tolist() == [0, 1]


def test_unrepresentable_closures_raise(dp):
    def bad_rate(output, parameters, population):
        output[0] = parameters[0] * np.sqrt(population[0])

    with pytest.raises(dp.ModelCompileError):
        dp.compile_rate_table(bad_rate, 1, 1, 2)

    def two_params(output, parameters, population):
        output[0] = parameters[0] * parameters[1] * population[0]

    with pytest.raises(dp.ModelCompileError):
        dp.compile_rate_table(two_params, 1, 2, 2)


def test_get_observations_matches_reference_format(dp):
    y = dp.get_observations(os.path.join(GOLDEN, "pooley.csv"))
    assert [o.time for o in y] == [20.0, 40.0, 60.0, 80.0, 100.0]
    assert [int(o.val[1]) for o in y] == [18, 65, 70, 66, 67] and all(o.obs_id == 1 and o.prop == 1.0 for o in y)
    y2 = dp.get_observations(np.array([[2.0, 5, 1], [1.0, 3, 0]]), type_col=3, val_seq=[2])
    assert [o.time for o in y2] == [1.0, 2.0] and [o.obs_id for o in y2] == [0, 1]


def test_prior_and_ess(dp):
    pr = dp.UniformProduct([0, 0], [0.01, 0.5])
    assert pr.logpdf([0.005, 0.1]) == pytest.approx(-np.log(0.01 * 0.5))
    assert pr.logpdf([0.02, 0.1]) == -np.inf
    assert pr.rand(7, np.random.default_rng(1)).shape == (2, 7)
    assert dp.compute_ess(np.ones(10)) == pytest.approx(10.0)


def test_bad_arguments_are_rejected_by_the_abi(built_lib, dp):
    model, y, hmm, theta = load_case(dp, "sis_pooley")
    cm = dp.compile_model(model, y)
    h = C.c_void_p()
    bad = dp._capi.ModelDesc.from_buffer_copy(bytes(cm.desc))
    bad.n_events = 99
    assert built_lib.dpomp_model_create(C.byref(bad), C.byref(h)) == -3
    assert b"limits" in built_lib.dpomp_last_error()
    assert built_lib.dpomp_model_create(None, C.byref(h)) == -1
    out = np.zeros(4, dtype=np.int64)
    w = np.ones(4)
    assert built_lib.dpomp_resample_indices(7, 0, w.ctypes.data, 4, w.ctypes.data, 4, 4, out.ctypes.data, -1) == -1


def test_c_driver_compiles_and_links_against_the_header(built_lib, tmp_path):
    """tests/c_driver.c (the no-Python host of the GPU suite) compiles as C11 against include/dpomp.h and links against the
    in-tree library; without a GPU it stops at the first device query with a clean error, not a crash."""
    import shutil
    import subprocess
    from conftest import ROOT
    libdir = os.path.join(ROOT, "discretepomp.jl_b200", "lib")
    exe = str(tmp_path / "c_driver")
    subprocess.run([shutil.which("gcc"), "-O1", "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_driver.c"), "-o", exe, "-L", libdir, "-ldpomp", "-lm", f"-Wl,-rpath,{libdir}"],
                   check=True, capture_output=True, text=True)
    import torch
    if not torch.cuda.is_available():
        res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert res.returncode == 1 and "FAIL" in res.stderr


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        depth += ch in "({["
        depth -= ch in ")}]"
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_julia_shim_agrees_with_the_header(built_lib, dp):
    """The ccall shim a DiscretePOMP.jl maintainer adds (untested here: no Julia in the image) must at least agree with
    include/dpomp.h symbol for symbol: every ccall names an exported entry point, passes as many arguments as the C
    declaration takes, with Julia types of the C argument's width and kind, and the `DpompModelDesc` struct lists the
    fields of `dpomp_model_desc` in order with the same element types and counts."""
    import re

    hdr = re.sub(r"/\*.*?\*/", " ", open(os.path.join(ROOT, "include", "dpomp.h")).read(), flags=re.S)
    jl = open(os.path.join(ROOT, "discretepomp.jl_b200", "julia", "DiscretePOMPGPU.jl")).read()
    decls = {}
    for m in re.finditer(r"\b(?:int|const char\s*\*)\s+(dpomp_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        args = [a.strip() for a in m.group(2).split(",")]
        decls[m.group(1)] = [] if args in (["void"], [""]) else args
    assert set(decls) == set(dp._capi.EXPORTED_SYMBOLS)  # the regex sees what the ctypes table lists

    def kind(c_arg):
        c = re.sub(r"\bconst\b", "", c_arg).strip()
        base = re.sub(r"\s+\w+$", "", c).replace(" ", "")  # drop the parameter name
        return base

    ok = {
        "int32_t": {"Int32", "Cint"}, "int64_t": {"Int64"}, "uint64_t": {"UInt64"},
        "double*": {"Ptr{Float64}", "Ref{Float64}"}, "int64_t*": {"Ptr{Int64}", "Ref{Int64}"},
        "int32_t*": {"Ptr{Int32}", "Ref{Int32}"}, "uint8_t*": {"Ptr{UInt8}"}, "void*": {"Ptr{UInt8}", "Ptr{Cvoid}"},
        "dpomp_model_desc*": {"Ref{DpompModelDesc}"},
    }
    calls = 0
    for m in re.finditer(r"ccall\(\(:(dpomp_[a-z0-9_]+),\s*LIBDPOMP\),\s*(\w+),\s*\(", jl):
        name, ret = m.group(1), m.group(2)
        assert name in decls, name
        i = j = m.end()
        depth = 1
        while depth:
            depth += jl[j] == "("
            depth -= jl[j] == ")"
            j += 1
        jtypes = _split_top(jl[i:j - 1])
        cargs = decls[name]
        assert len(jtypes) == len(cargs), (name, jtypes, cargs)
        assert ret == ("Cstring" if name == "dpomp_last_error" else "Cint"), (name, ret)
        for jt, ca in zip(jtypes, cargs):
            k = kind(ca)
            if re.fullmatch(r"dpomp_(model|pf|comm|mbp)\*", k):
                assert jt == "Ptr{Cvoid}", (name, jt, ca)  # opaque handle
            elif re.fullmatch(r"dpomp_(model|pf|comm|mbp)\*\*", k):
                assert jt == "Ref{Ptr{Cvoid}}", (name, jt, ca)
            else:
                assert jt in ok[k], (name, jt, ca)
        calls += 1
    assert calls >= 25
    # struct dpomp_model_desc vs struct DpompModelDesc: same fields, order, element type and count
    body = re.search(r"typedef struct dpomp_model_desc \{(.*?)\} dpomp_model_desc;", hdr, flags=re.S).group(1)
    dims = {"DPOMP_MAX_EVENTS": 8, "DPOMP_MAX_COMPARTMENTS": 8, "DPOMP_MAX_OBS_VALS": 8}
    for k in dims:
        assert re.search(rf"#define {k}\s+{dims[k]}\b", hdr), k
    c_fields = []
    for m in re.finditer(r"(const\s+)?(\w+)\s*(\*)?\s*(\w+)((?:\[\w+\])*)\s*;", body):
        n = 1
        for d in re.findall(r"\[(\w+)\]", m.group(5)):
            n *= dims[d]
        c_fields.append((m.group(4), m.group(2) + ("*" if m.group(3) else ""), n))
    jbody = re.search(r"struct DpompModelDesc\n(.*?)\nend", jl, flags=re.S).group(1)
    jmap = {"Int32": "int32_t", "Int64": "int64_t", "Float64": "double", "Ptr{Float64}": "double*", "Ptr{Int32}": "int32_t*",
            "Ptr{Int64}": "int64_t*"}
    j_fields = []
    for f in re.split(r"[;\n]", jbody):
        f = f.strip()
        if not f:
            continue
        nm, ty = f.split("::")
        t = re.fullmatch(r"NTuple\{(\d+),(\w+)\}", ty)
        j_fields.append((nm, jmap[t.group(2)], int(t.group(1))) if t else (nm, jmap[ty], 1))
    assert j_fields == c_fields, (j_fields, c_fields)

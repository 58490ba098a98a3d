"""Static checks of julia/AdvancedVIB200.jl -- the `@ccall` glue a maintainer adds to AdvancedVI.jl.

Julia is not installed in this image, so the file cannot be executed here; what CAN be verified without Julia is that
every foreign call in it binds an entry point that include/avi.h declares and libavi_b200.so exports, with the
declared number of arguments and argument types that agree with the C prototype (pointer / 32-bit / 64-bit / float),
that its block structure is balanced, that no name is used before it is imported, and -- when the reference checkout
is present -- that every AdvancedVI function it extends and every name it imports exists in the reference's sources."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GLUE = os.path.join(ROOT, "julia", "AdvancedVIB200.jl")
HEADER = os.path.join(ROOT, "include", "avi.h")
REF = "/root/reference"


def _strip_c_comments(s):
    return re.sub(r"/\*.*?\*/", " ", s, flags=re.S)


def header_prototypes():
    """name -> list of C parameter type strings"""
    src = _strip_c_comments(open(HEADER).read())
    protos = {}
    for m in re.finditer(r"\b(int32_t|int64_t|const char\s*\*|void)\s+(avi_\w+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        args = [a.strip() for a in m.group(3).replace("\n", " ").split(",")] if m.group(3).strip() not in ("", "void") else []
        protos[m.group(2)] = args
    return protos


def _strip_julia(s):
    """comments and string literals out (keeps line structure)"""
    s = re.sub(r'""".*?"""', lambda m: '""' + "\n" * m.group(0).count("\n"), s, flags=re.S)   # docstrings
    s = re.sub(r"#=.*?=#", lambda m: "\n" * m.group(0).count("\n"), s, flags=re.S)             # block comments
    out = []
    for line in s.split("\n"):
        buf, i, in_str = [], 0, False
        while i < len(line):
            ch = line[i]
            if in_str:
                if ch == "\\":
                    i += 2
                    continue
                if ch == '"':
                    in_str = False
                i += 1
                continue
            if ch == '"':
                in_str = True
                buf.append('""')
                i += 1
                continue
            if ch == "#":
                break
            buf.append(ch)
            i += 1
        out.append("".join(buf))
    return "\n".join(out)


def _split_top_level(s, sep=","):
    parts, depth, cur = [], 0, []
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == sep and depth == 0:
            parts.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    if "".join(cur).strip():
        parts.append("".join(cur))
    return [p.strip() for p in parts]


def glue_ccalls(text=None):
    """[(symbol, [julia arg type, ...], return type)]"""
    src = _strip_julia(open(GLUE).read() if text is None else text)
    calls = []
    for m in re.finditer(r"@ccall\s*\(?\s*libavi\.(avi_\w+)\(", src):
        i = m.end()
        depth, j = 1, i
        while depth:
            depth += {"(": 1, ")": -1}.get(src[j], 0)
            j += 1
        args = _split_top_level(src[i:j - 1].replace("\n", " "))
        ret = re.match(r"\s*::\s*(\w+(\{[^}]*\})?)", src[j:])
        types = []
        for a in args:
            depth, k = 0, len(a) - 1        # the LAST top-level `::` of the argument
            pos = -1
            while k > 0:
                depth += {")": 1, "]": 1, "}": 1, "(": -1, "[": -1, "{": -1}.get(a[k], 0)
                if depth == 0 and a[k - 1:k + 1] == "::":
                    pos = k - 1
                    break
                k -= 1
            assert pos >= 0, f"{m.group(1)}: argument without a type annotation: {a!r}"
            types.append(a[pos + 2:].strip())
        calls.append((m.group(1), types, ret.group(1) if ret else None))
    return calls


def _c_class(t):
    t = t.replace("const ", "").strip()
    if "*" in t or "avi_allreduce_fn" in t or "avi_logdensity_fn" in t or re.search(r"_fn\b", t):
        return "ptr"
    base = t.split()[0]
    return {"int32_t": "i32", "int64_t": "i64", "uint64_t": "u64", "uint32_t": "u32", "float": "f32", "double": "f64"}[base]


def _jl_class(t):
    if t.startswith(("Ptr{", "Ref{")) or t in ("Cstring", "Ptr"):
        return "ptr"
    return {"Int32": "i32", "Cint": "i32", "Int64": "i64", "Clonglong": "i64", "UInt64": "u64", "UInt32": "u32",
            "Float32": "f32", "Cfloat": "f32", "Float64": "f64", "Cdouble": "f64"}[t]


def check_ccalls(calls, protos, lib):
    for name, jl_types, ret in calls:
        assert name in protos, f"{name} is not declared in include/avi.h"
        assert hasattr(lib, name), f"{name} is not exported by libavi_b200.so"
        c_args = protos[name]
        assert len(c_args) == len(jl_types), f"{name}: {len(jl_types)} arguments in the glue, {len(c_args)} in avi.h"
        for k, (ct, jt) in enumerate(zip(c_args, jl_types)):
            assert _c_class(ct) == _jl_class(jt), f"{name}: argument {k} is `{ct}` in avi.h but `{jt}` in the glue"
        assert ret in ("Int32", "Int64", "Cstring", "Cvoid"), (name, ret)


def test_every_ccall_binds_a_declared_and_exported_symbol_with_matching_signature():
    protos = header_prototypes()
    assert len(protos) > 50
    lib = ctypes.CDLL(os.path.join(ROOT, "advancedvi.jl_b200", "libavi_b200.so"))
    calls = glue_ccalls()
    assert len(calls) >= 25
    check_ccalls(calls, protos, lib)


def test_the_checker_catches_broken_bindings():
    """The checks above are only worth something if they fail on a broken glue: an undeclared symbol, a dropped argument,
    a 32-bit integer where the header has 64 bits, a float passed where a pointer is expected."""
    protos = header_prototypes()
    lib = ctypes.CDLL(os.path.join(ROOT, "advancedvi.jl_b200", "libavi_b200.so"))
    good = 'check(@ccall(libavi.avi_obj_set_base(st.h::Ptr{Cvoid}, bc::Int32, bp::Float32)::Int32), c.h)'
    check_ccalls(glue_ccalls(good), protos, lib)
    for bad in (good.replace("avi_obj_set_base", "avi_obj_set_basis"),                       # no such entry point
                good.replace(", bp::Float32", ""),                                            # argument dropped
                good.replace("bc::Int32", "bc::Int64"),                                       # wrong integer width
                good.replace("st.h::Ptr{Cvoid}", "st.h::Float32"),                            # handle passed by value
                'check(@ccall(libavi.avi_model_subsample(p.h::Ptr{Cvoid}, p.idx::Ptr{Int32}, length(p.idx)::Int32)::Int32), c)'):
        with pytest.raises(AssertionError):
            check_ccalls(glue_ccalls(bad), protos, lib)


def test_block_structure_is_balanced():
    """Every function / if / for / while / let / struct / module / begin / try / do / quote / macro has its `end`
    (`a[end]` and generators / comprehensions live inside brackets and are not block structure)."""
    for path in (GLUE, os.path.join(ROOT, "julia", "test", "runtests.jl")):
        src = _strip_julia(open(path).read())
        depth = 0          # nesting of ( and [
        opens = ends = 0
        for tok in re.finditer(r"[\[\]()]|\b(function|if|for|while|let|struct|module|begin|try|do|quote|macro|end)\b", src):
            t = tok.group(0)
            if t in "([":
                depth += 1
            elif t in ")]":
                depth -= 1
                assert depth >= 0, f"{os.path.basename(path)}: unbalanced bracket near offset {tok.start()}"
            elif t == "end":
                if depth == 0:
                    ends += 1
            elif t in ("for", "if"):
                if depth == 0:
                    opens += 1
            else:
                opens += 1          # (`function` / `begin` / `do` blocks may sit inside a call's parentheses)
                if depth > 0 and t in ("function", "begin", "do", "let", "try", "quote"):
                    ends -= 1       # ... and so does their `end`: account for it here
        assert depth == 0, f"{os.path.basename(path)}: unbalanced brackets"
        assert opens == ends, f"{os.path.basename(path)}: {opens} block openers vs {ends} `end`"


def test_names_are_imported_before_use():
    """The round-1 glue used Optimisers, LinearAlgebra.AbstractTriangular and Normal without importing them."""
    raw = open(GLUE).read()
    src = _strip_julia(raw)
    header = src[:src.index("const libavi")]
    for mod in ("Optimisers", "LinearAlgebra", "Random", "DiffResults", "LogDensityProblems", "ADTypes", "AdvancedVI"):
        if re.search(rf"\b{mod}\.", src):
            assert re.search(rf"\busing\b[^\n]*\b{mod}\b", header) or re.search(rf"\bimport\b[^\n]*\b{mod}\b", header), mod
    imported = set(re.findall(r"\b[A-Z]\w+", " ".join(re.findall(r"using \w+: ([^\n]*(?:\n {4,}[^\n]*)*)", header))))
    for name in ("Normal", "Laplace", "TDist", "Diagonal", "LowerTriangular", "RepGradELBO", "ScoreGradELBO", "SubsampledObjective",
                 "MvLocationScale", "MvLocationScaleLowRank", "ClipScale", "IdentityOperator", "PolynomialAveraging", "DoG", "DoWG"):
        if re.search(rf"(?<![\w.]){name}\b", src[len(header):]):
            assert name in imported, f"{name} is used but not imported"
    # every struct field read through `.c.` exists on Ctx (round 1 read prob.c.device, which Ctx did not have)
    ctx_def = re.search(r"mutable struct Ctx(.*?)\nend", src, flags=re.S).group(1)
    ctx_fields = set(re.findall(r"^\s*(\w+)::", ctx_def, flags=re.M))
    for f in set(re.findall(r"\.c\.(\w+)", src)):
        assert f in ctx_fields, f"Ctx has no field `{f}`"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference checkout not present")
def test_extended_functions_and_imported_names_exist_in_the_reference():
    ref_src = ""
    for dp, _, fns in os.walk(os.path.join(REF, "src")):
        for fn in fns:
            if fn.endswith(".jl"):
                ref_src += open(os.path.join(dp, fn)).read() + "\n"
    src = _strip_julia(open(GLUE).read())
    for fn in set(re.findall(r"function AdvancedVI\.(\w+!?)\(", src)) | set(re.findall(r"^AdvancedVI\.(\w+!?)\(", src, flags=re.M)):
        assert re.search(rf"function {re.escape(fn)}\s*[({{]|^{re.escape(fn)}\(|\b{re.escape(fn)}\(.*\) =", ref_src, flags=re.M), \
            f"AdvancedVI.{fn} is extended by the glue but not defined in the reference"
    header = src[:src.index("const libavi")]
    names = re.findall(r"\b\w+", " ".join(re.findall(r"using AdvancedVI: ([^\n]*(?:\n {4,}[^\n]*)*)", header)))
    assert len(names) > 10
    for name in names:
        assert re.search(rf"\b(struct|function|abstract type)\s+{name}\b|^{name}\(|const {name}\b", ref_src, flags=re.M), \
            f"`{name}` is imported from AdvancedVI but not defined in the reference"


def test_project_toml_lists_every_package_the_glue_and_its_tests_use():
    """julia/Project.toml: every `using X` of the glue and of runtests.jl is a dependency, with the UUID the reference's
    own Project.toml / test/Project.toml gives that package (when the reference checkout is present)."""
    toml = open(os.path.join(ROOT, "julia", "Project.toml")).read()
    deps = dict(re.findall(r'^(\w+) = "([0-9a-f-]{36})"', toml, flags=re.M))
    used = set()
    for path in (GLUE, os.path.join(ROOT, "julia", "test", "runtests.jl")):
        for line in _strip_julia(open(path).read()).split("\n"):
            m = re.match(r"\s*using\s+(.*)", line)
            if m and not m.group(1).startswith("."):
                for part in m.group(1).split(":")[0].split(","):
                    used.add(part.strip())
    assert used and used <= set(deps), used - set(deps)
    if os.path.isdir(REF):
        ref = open(os.path.join(REF, "Project.toml")).read() + open(os.path.join(REF, "test", "Project.toml")).read()
        ref_ids = dict(re.findall(r'^(\w+) = "([0-9a-f-]{36})"', ref, flags=re.M))
        for name, uid in deps.items():
            if name in ref_ids:
                assert ref_ids[name] == uid, name
        assert re.search(r'^uuid = "' + deps["AdvancedVI"] + '"', ref, flags=re.M)


_BASE_FUNCTIONS = set("""length size zeros ones fill similar copy collect map sum abs2 rand randn isnothing isa get get! haskey
push! throw error string unsafe_string finalizer reinterpret pointer convert eltype typeof min max minimum maximum any all first
last isempty vec reshape transpose adjoint view copyto! unsafe_copyto! div rem mod ceil floor round sqrt exp log abs println print
show repr zero one iszero isfinite isnan ntuple getfield setfield! getproperty setproperty! hasproperty hasfield fieldnames nameof
eachindex axes range vcat hcat cat getindex setindex! in keys values pairs tuple something ifelse merge unsafe_wrap unsafe_load
unsafe_store! deepcopy identity foreach filter reduce mapreduce findfirst sort sort! unique reverse dof params diag tril triu
inv det logdet dot norm mul! ldiv! sizeof include joinpath dirname abspath parse float new""".split())


def test_every_called_function_is_defined_imported_or_base():
    """A name that is called but neither defined in the module, nor a closure / keyword argument, nor a Base /
    LinearAlgebra / Distributions function the module imports: the round-1 glue had several."""
    src = _strip_julia(open(GLUE).read())
    called = set(re.findall(r"(?<![\w.@:])([a-z_][A-Za-z0-9_!]*)\(", src))
    defined = set(re.findall(r"function\s+(?:\w+\.)?([A-Za-z_][\w!]*)", src))
    defined |= set(re.findall(r"^\s*(?:\w+\.)?([a-z_][\w!]*)\([^=\n]*\)\s*(?:where[^=\n]*)?=(?!=)", src, flags=re.M))
    local = set(re.findall(r"\b([a-z_]\w*)\s*(?:=|::)", src)) | set(re.findall(r"[(,;]\s*([a-z_]\w*)\s*(?:[,;)]|::)", src))
    unknown = sorted(called - defined - _BASE_FUNCTIONS - local)
    assert not unknown, f"called but never defined / imported: {unknown}"

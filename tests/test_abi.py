"""The C-ABI library loads, exports every symbol include/gcdyn_b200.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device.  No compute happens here."""
import ctypes
import os
import re

import pytest

import gcdyn
from gcdyn import likelihood

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "gcdyn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gcdyn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = _declared()
    assert len(names) >= 15
    lib = ctypes.CDLL(likelihood.library_path())
    for n in names:
        assert hasattr(lib, n), n
    import re
    header = open(os.path.join(ROOT, "include", "gcdyn_b200.h")).read()
    declared = int(re.search(r"#define GCDYN_B200_VERSION (\d+)", header).group(1))
    assert likelihood._lib().gcdyn_version() == declared >= 110
    assert likelihood._lib().gcdyn_last_error(None) is not None


def test_ctypes_structs_match_header_layout():
    # field order/size of the structs the Python binding mirrors (a Julia `struct` would mirror the same)
    assert ctypes.sizeof(likelihood.ForestDesc) == 8 + 8 + 4 + 4 + 13 * 8
    assert ctypes.sizeof(likelihood.Params) == 4 * 4 + 11 * 8
    assert ctypes.sizeof(likelihood.Opts) == 4 + 4 + 8 + 4 + 4


def test_ctypes_field_offsets_equal_the_c_compiler_s(tmp_path):
    """Offsets of every field of the four boundary structs as gcc lays them out == offsets of the ctypes mirrors, in order."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    header = open(os.path.join(ROOT, "include", "gcdyn_b200.h")).read()
    structs = {"gcdyn_forest_desc": likelihood.ForestDesc, "gcdyn_params": likelihood.Params,
               "gcdyn_opts": likelihood.Opts, "gcdyn_mh_desc": likelihood.MhDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "gcdyn_b200.h"', 'int main(void) {']
    names = {}
    for cname in structs:
        end = header.index("} %s;" % cname)
        start = header.rindex("typedef struct {", 0, end) + len("typedef struct {")
        body = header[start:end]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):                      # "double rho, sigma" declares two fields
                fields.append(re.findall(r"[A-Za-z_][A-Za-z_0-9]*", part)[-1])
        names[cname] = fields
        for f in fields:
            lines.append('  printf("%s %s %%zu\\n", offsetof(%s, %s));' % (cname, f, cname, f))
        lines.append('  printf("%s sizeof %%zu\\n", sizeof(%s));' % (cname, cname))
    lines += ["  return 0;", "}"]
    src = tmp_path / "off.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "off"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = {}
    for line in out.splitlines():
        c, f, v = line.split()
        got.setdefault(c, []).append((f, int(v)))
    for cname, cls in structs.items():
        c_offsets = [v for f, v in got[cname] if f != "sizeof"]
        py_offsets = [getattr(cls, f).offset for f, _ in cls._fields_]
        assert c_offsets == py_offsets, (cname, names[cname], c_offsets, py_offsets)
        assert dict(got[cname])["sizeof"] == ctypes.sizeof(cls), cname


def test_no_cpu_fallback_without_a_device():
    lib = likelihood._lib()
    if lib.gcdyn_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(gcdyn.GcdynError) as e:
        gcdyn.Context(0)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)
    model = gcdyn.ConstantBranchingProcess(1.8, 1.0, 1.0, 1, 0, [1, 2])
    r = gcdyn.TreeNode("root", 0, 1)
    gcdyn.attach(r, gcdyn.TreeNode("sampled_survival", 2.0, 1))
    with pytest.raises(gcdyn.GcdynError):
        gcdyn.loglikelihood(model, r, 2.0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "gcdyn.jl_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                text = open(os.path.join(d, f), encoding="utf-8").read()
                assert not re.search(r"(import|from)\s+(oracle|refcpu|algo_model)\b", text), f
                assert "librefcpu" not in text, f


def test_header_is_plain_c99_and_cxx11(tmp_path):
    """The boundary is a C ABI: the header must compile on its own as strict C99 (what a Julia/C consumer sees) and as
    C++11, warnings as errors."""
    import shutil
    import subprocess
    if not shutil.which("gcc"):
        pytest.skip("no gcc")
    src = tmp_path / "hdr.c"
    src.write_text('#include "gcdyn_b200.h"\n'
                   'int main(void) { gcdyn_opts o = {0, 0, 0.0, 0, 0}; gcdyn_mh_desc d; gcdyn_params p; gcdyn_forest_desc f;\n'
                   '  (void)o; (void)d; (void)p; (void)f; return GCDYN_B200_VERSION >= 110 ? 0 : 1; }\n')
    inc = os.path.join(ROOT, "include")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", inc, "-c", str(src),
                    "-o", str(tmp_path / "a.o")], check=True)
    subprocess.run(["g++", "-std=c++11", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", inc, "-x", "c++", "-c", str(src),
                    "-o", str(tmp_path / "b.o")], check=True)


def test_julia_struct_mirrors_follow_the_header():
    """The Julia backend cannot be executed here; at least its `struct` mirrors must list the header's fields in the header's
    order with layout-compatible types (a `ccall` passes them by reference)."""
    header = open(os.path.join(ROOT, "include", "gcdyn_b200.h")).read()
    julia = open(os.path.join(ROOT, "gcdyn.jl_b200", "julia", "b200_backend.jl")).read()
    pairs = {"gcdyn_forest_desc": "GcdynForestDesc", "gcdyn_params": "GcdynParams", "gcdyn_opts": "GcdynOpts",
             "gcdyn_mh_desc": "GcdynMhDesc"}
    ctype = {"int32_t": "Int32", "int64_t": "Int64", "uint64_t": "UInt64", "double": "Float64", "uint8_t": "UInt8"}
    for cname, jname in pairs.items():
        end = header.index("} %s;" % cname)
        body = header[header.rindex("typedef struct {", 0, end) + len("typedef struct {"):end]
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        want = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            base = re.match(r"(const\s+)?([A-Za-z_0-9]+)", decl).group(2)
            for part in decl.split(","):
                name = re.findall(r"[A-Za-z_][A-Za-z_0-9]*", part)[-1]
                is_ptr = "*" in part or ("*" in decl.split(",")[0] and part is decl.split(",")[0])
                want.append((name, "Ptr{%s}" % ctype[base] if is_ptr else ctype[base]))
        jbody = re.search(r"^struct %s\n(.*?)^end" % jname, julia, re.S | re.M).group(1)
        got = [tuple(x.strip().split("::")) for x in jbody.strip().splitlines()]
        assert [g[0] for g in got] == [w[0] for w in want], (cname, got, want)
        for (gn, gt), (wn, wt) in zip(got, want):
            assert gt == wt or (wt.startswith("Ptr") and gt.startswith("Ptr")), (cname, gn, gt, wt)


def test_julia_ccalls_pass_as_many_arguments_as_the_header_declares():
    header = open(os.path.join(ROOT, "include", "gcdyn_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    julia = open(os.path.join(ROOT, "gcdyn.jl_b200", "julia", "b200_backend.jl")).read()
    declared = {}
    for m in re.finditer(r"\b(gcdyn_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, re.S):
        args = m.group(2).strip()
        declared[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    used = 0
    for m in re.finditer(r"ccall\(\(:(gcdyn_[a-z0-9_]+), LIBGCDYN\),\s*[A-Za-z]+,\s*\(", julia):
        name = m.group(1)
        i, depth, start = m.end(), 1, m.end()
        while depth:                                   # the matching ')' of the argument-type tuple
            depth += {"(": 1, ")": -1}.get(julia[i], 0)
            i += 1
        types = julia[start:i - 1]
        flat, d = [], 0
        for ch in types:                               # commas at depth 0 separate types (Ptr{...}, Ref{...} have none inside)
            d += {"{": 1, "}": -1}.get(ch, 0)
            flat.append("," if (ch == "," and d == 0) else ch if ch != "," else ";")
        n = len([t for t in "".join(flat).split(",") if t.strip()])
        assert name in declared, name
        assert n == declared[name], (name, n, declared[name], types)
        used += 1
    assert used >= 12

"""Static consistency checks of fortran/libGPU.f90 (the iso_c_binding shim) against include/volcanor_b200.h.

No Fortran compiler exists in this image ("Boundary"), so the shim cannot be compiled here; its tested twin
is tests/native/case_gpu_hooks.c.  What CAN be checked without a compiler is everything a binding typically gets wrong:
every `bind(C, name=...)` interface names a function the header declares, with the same number of arguments in the same
order, the same C types (int <-> integer(c_int), int64_t <-> integer(c_int64_t), double <-> real(c_double), pointers <->
arrays / type(c_ptr)), scalars passed by `value` and arrays by reference; every `import` covers the kinds the interface
uses; blocks are balanced; every public procedure is defined; the named constants equal the header's enums."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
F90 = (ROOT / "fortran" / "libGPU.f90").read_text()
HDR = (ROOT / "include" / "volcanor_b200.h").read_text()


def _join_continuations(src: str) -> list[str]:
    """Free-form Fortran: strip comments, join `&` continuation lines."""
    out, cur = [], ""
    for raw in src.splitlines():
        line = raw.split("!")[0].rstrip() if "'" not in raw.split("!")[0] or raw.count("'") % 2 == 0 else raw.rstrip()
        if not line.strip():
            continue
        s = line.strip()
        if s.startswith("&"):
            s = s[1:].lstrip()
        if s.endswith("&"):
            cur += s[:-1].rstrip() + " "
            continue
        out.append(cur + s)
        cur = ""
    return out


LINES = _join_continuations(F90)


def c_prototypes() -> dict:
    """name -> (return type, [(type, is_pointer, name)]) for every vlc_* function of the header."""
    txt = re.sub(r"/\*.*?\*/", "", HDR, flags=re.S)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(vlc_\w+)\s*\(([^)]*)\)\s*;", txt):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                ptr = "*" in a
                toks = a.replace("*", " ").split()
                pname = toks[-1]
                base = " ".join(t for t in toks[:-1] if t != "const")
                params.append((base, ptr, pname))
        protos[name] = (ret, params)
    return protos


def fortran_interfaces() -> dict:
    """name -> dict(result=..., args=[names], decl={name: (type, has_value, is_array)}, imports=set())."""
    out, cur = {}, None
    inside = False
    for l in LINES:
        low = l.lower()
        if low == "interface":
            inside = True
            continue
        if low == "end interface":
            inside = False
            continue
        if not inside:
            continue
        m = re.match(r"(.+?)\s+function\s+(\w+)\s*\(([^)]*)\)\s*bind\s*\(\s*c\s*,\s*name\s*=\s*'(\w+)'\s*\)", l, flags=re.I)
        if m:
            cur = dict(result=m.group(1).strip().lower(), fname=m.group(2), args=[a.strip() for a in m.group(3).split(",") if a.strip()],
                       decl={}, imports=set())
            out[m.group(4)] = cur
            continue
        if low.startswith("end function"):
            cur = None
            continue
        if cur is None:
            continue
        if low.startswith("import"):
            cur["imports"] |= {t.strip().lower() for t in l.split("::")[1].split(",")}
            continue
        typ, names = l.split("::")
        attrs = [a.strip().lower() for a in re.split(r",(?![^()]*\))", typ)]
        for n in re.split(r",(?![^()]*\))", names):
            n = n.strip()
            base = re.match(r"\w+", n).group(0)
            cur["decl"][base] = (attrs[0], "value" in attrs, "(" in n or any(a.startswith("dimension") for a in attrs))
    return out


PROTOS = c_prototypes()
IFACES = fortran_interfaces()
KIND = {"int": "integer(c_int)", "int64_t": "integer(c_int64_t)", "double": "real(c_double)"}


def test_shim_binds_only_declared_functions_and_covers_the_case_path():
    assert len(IFACES) >= 35
    for name, f in IFACES.items():
        assert name in PROTOS, f"{name} is not declared in include/volcanor_b200.h"
        assert f["fname"] == name
    # what the tested C twin (tests/native/case_gpu_hooks.c) calls, the Fortran shim binds as well
    twin = (ROOT / "tests" / "native" / "case_gpu_hooks.c").read_text()
    used = set(re.findall(r"\b(vlc_\w+)\s*\(", twin))
    sharded = {"vlc_wake_sweep_count", "vlc_wake_sweep_slice", "vlc_wake_sweep_scatter", "vlc_sync"}  # MPI variant: comment in the shim
    read_back = {"vlc_rotor_get_wakevel"}                                       # the twin's test-only download of velocity arrays
    missing = sorted(used - set(IFACES) - sharded - read_back)
    assert not missing, missing


@pytest.mark.parametrize("name", sorted(IFACES))
def test_interface_matches_the_c_prototype(name):
    ret, params = PROTOS[name]
    f = IFACES[name]
    assert len(f["args"]) == len(params), (name, f["args"], [p[2] for p in params])
    want_ret = {"int": "integer(c_int)", "const char": "type(c_ptr)", "int64_t": "integer(c_int64_t)"}[ret.replace("*", "").strip()]
    assert f["result"] == want_ret, (name, f["result"], ret)
    for arg, (ctype, is_ptr, cname) in zip(f["args"], params):
        assert arg in f["decl"], (name, arg, "dummy argument without a declaration")
        ftype, by_value, is_array = f["decl"][arg]
        if not is_ptr:                                   # scalar by value
            assert ftype == KIND[ctype] and by_value and not is_array, (name, arg, ftype, ctype)
        elif ctype in ("vlc_ctx", "void"):               # opaque handles
            assert ftype == "type(c_ptr)", (name, arg, ftype)
            assert by_value == (cname != "out" and not (name == "vlc_create" and arg == "out")), (name, arg)
        else:                                            # arrays by reference
            assert ftype == KIND[ctype] and not by_value, (name, arg, ftype, ctype)
            assert is_array, (name, arg, "pointer argument must be an array dummy")
    used = {t for d in f["decl"].values() for t in re.findall(r"c_\w+", d[0])} | set(re.findall(r"c_\w+", f["result"]))
    assert used <= f["imports"], (name, used - f["imports"])


def test_named_constants_equal_the_header_enums():
    txt = re.sub(r"/\*.*?\*/", "", HDR, flags=re.S)
    every = {k: int(v) for k, v in re.findall(r"\b(VLC_VEL_\w+)\s*=\s*(\d+)", txt)}
    enums = {k: v for k, v in every.items() if not k.startswith("VLC_VEL_ARRAY")}
    arrays = {k: v for k, v in every.items() if k.startswith("VLC_VEL_ARRAY")}
    assert len(enums) == 6 and len(arrays) == 6
    consts = {}
    for l in LINES:
        if "parameter" in l.lower() and "::" in l:
            for k, v in re.findall(r"(\w+)\s*=\s*(-?\d+)", l.split("::")[1]):
                consts[k] = int(v)
    for k, v in enums.items():
        assert consts.get(k.replace("VLC_", "")) == v, (k, v, consts.get(k.replace("VLC_", "")))
    assert (consts["GPU_BYWING"], consts["GPU_BYWAKE"], consts["GPU_BOTH"], consts["GPU_BOUNDVORTICES"]) == (0, 1, 2, 3)
    names = {"VLC_VEL_ARRAY": "ARR_VEL", "VLC_VEL_ARRAY_1": "ARR_VEL1", "VLC_VEL_ARRAY_PREDICTED": "ARR_PREDICTED",
             "VLC_VEL_ARRAY_STEP": "ARR_STEP", "VLC_VEL_ARRAY_2": "ARR_VEL2", "VLC_VEL_ARRAY_3": "ARR_VEL3"}
    for k, v in arrays.items():
        assert consts.get(names[k]) == v, (k, v)


def test_blocks_balance_and_public_procedures_exist():
    opens = {"do": 0, "if": 0, "select": 0, "associate": 0, "interface": 0}
    procs, ends = [], []
    for l in LINES:
        low = l.lower()
        if re.match(r"(end\s*do|enddo)\b", low):
            opens["do"] -= 1
        elif re.match(r"do\b", low):
            opens["do"] += 1
        if re.match(r"(end\s*if|endif)\b", low):
            opens["if"] -= 1
        elif re.match(r"(else\s*)?if\s*\(.*\)\s*then$", low) and not low.startswith("else"):
            opens["if"] += 1
        if low.startswith("end select"):
            opens["select"] -= 1
        elif low.startswith("select case"):
            opens["select"] += 1
        if low.startswith("end associate"):
            opens["associate"] -= 1
        elif low.startswith("associate"):
            opens["associate"] += 1
        if low == "end interface":
            opens["interface"] -= 1
        elif low == "interface":
            opens["interface"] += 1
        m = re.match(r"(?:subroutine|function)\s+(\w+)", low)
        if m and "bind" not in low:
            procs.append(m.group(1))
        m = re.match(r"end\s+(?:subroutine|function)\s+(\w+)", low)
        if m:
            ends.append(m.group(1))
    assert all(v == 0 for v in opens.values()), opens
    assert procs == ends, (procs, ends)                 # every procedure closed under its own name, in order
    public = set()
    for l in LINES:
        if l.lower().startswith("public ::"):
            public |= {t.strip().lower() for t in l.split("::")[1].split(",")}
    defined = set(procs) | {"gpu_wing", "gpu_wake_c", "gpu_wake_p"}
    assert public <= defined, public - defined
    assert {"gpu_cp_rhs_solve", "gpu_cp_forces", "gpu_wake_convect", "gpu_vind_onnwake_byrotor"} <= public


def test_every_called_binding_is_declared_in_the_interface_block():
    body = "\n".join(LINES)
    called = set(re.findall(r"\b(vlc_\w+)\s*\(", body))
    assert called <= set(IFACES), called - set(IFACES)
    # and with the right number of actual arguments
    for l in LINES:
        for m in re.finditer(r"\b(vlc_\w+)\s*\(", l):
            if "function" in l.lower():
                continue
            depth, i, n, start = 1, m.end(), 1, m.end()
            while depth and i < len(l):
                ch = l[i]
                depth += ch in "(["
                depth -= ch in ")]"
                if ch == "," and depth == 1:
                    n += 1
                i += 1
            if l[start:i - 1].strip() == "":
                n = 0
            assert n == len(IFACES[m.group(1)]["args"]), (l, n, len(IFACES[m.group(1)]["args"]))


REF_CLASSDEF = Path("/root/reference/src/classdef.f90")


@pytest.mark.skipif(not REF_CLASSDEF.exists(), reason="the reference tree is only present in the build container")
def test_every_derived_type_member_the_shim_touches_exists_in_the_reference():
    """`rotor%blade(ib)%secTauCapChord`, `rotor%nbConvect`, `...%wiP(ic, is)%velCP`, `call rotor(ir)%map_gam()` ...: each
    component / type-bound procedure name used after a `%` is declared in the reference's classdef.f90 (in the type that
    owns it: rotor_class, blade_class, wingpanel_class, pFwake_class)."""
    ref = REF_CLASSDEF.read_text()
    types = {}
    for m in re.finditer(r"^\s*type(?:\s*,\s*\w+)*\s*(?:::)?\s*(\w+_class)\s*$(.*?)^\s*end type", ref, flags=re.S | re.M | re.I):
        body = m.group(2)
        names = set()
        for l in body.splitlines():
            l = l.split("!")[0]
            if "::" in l:
                for n in re.split(r",(?![^()]*\))", l.split("::", 1)[1]):
                    n = n.strip()
                    if "=>" in n:
                        n = n.split("=>")[0].strip()
                    mm = re.match(r"\w+", n)
                    if mm:
                        names.add(mm.group(0).lower())
        types[m.group(1).lower()] = names
    assert {"rotor_class", "blade_class", "wingpanel_class"} <= set(types)
    owner = {"rotor": "rotor_class", "blade": "blade_class", "wip": "wingpanel_class", "b": "blade_class",
             "wapf": "pfwake_class", "wapfpredicted": "pfwake_class"}
    checked = 0
    for l in LINES:
        for m in re.finditer(r"\b(\w+)(?:\([^()]*(?:\([^()]*\)[^()]*)*\))?%(\w+)", l):
            base, member = m.group(1).lower(), m.group(2).lower()
            if base in owner:
                assert member in types[owner[base]], (l.strip(), base, member)
                checked += 1
    assert checked > 80, checked

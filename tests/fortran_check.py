"""A free-form Fortran checker for the generated ``bind(c)`` interface module -- TEST INFRASTRUCTURE.

This image has no Fortran compiler, so the acceptance the reference gets by compiling and running
test/py_ftn_interface/data/fortran_program.f90 against its generated module is replaced by two checks a compiler would
make, done here by parsing:

1. the source is legal free-form Fortran 2008 *of the subset the generator emits*: physical lines <= 132 columns, at
   most 255 continuation lines per statement, no tabs, identifiers <= 63 characters, balanced module / interface /
   function / program blocks with matching names, every statement one of the known forms;
2. every interface is a correct interoperable declaration of the C prototype of the same name in
   include/b200stencil.h: same number of dummies in the same order, every dummy declared exactly once (names compared
   case-insensitively, as Fortran does), C scalars passed ``value`` with the matching kind, pointers either an opaque
   ``type(c_ptr), value`` or a by-reference entity of the pointee's interoperable type, result kind matching the C
   return type, every kind named in ``import`` and in the module's ``use iso_c_binding, only:`` list.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

IDENT = re.compile(r"^[A-Za-z][A-Za-z0-9_]*$")


class FortranError(AssertionError):
    pass


def strip_comment(line: str) -> str:
    out, quote = [], None
    for ch in line:
        if quote:
            out.append(ch)
            if ch == quote:
                quote = None
        elif ch in "'\"":
            quote = ch
            out.append(ch)
        elif ch == "!":
            break
        else:
            out.append(ch)
    return "".join(out).rstrip()


def statements(text: str) -> List[Tuple[int, str]]:
    """Logical statements (continuations joined, comments removed) with the line number they start on."""
    out, cur, start, ncont = [], "", 0, 0
    for no, raw in enumerate(text.splitlines(), 1):
        if "\t" in raw:
            raise FortranError(f"line {no}: tab character")
        if len(raw) > 132:
            raise FortranError(f"line {no}: {len(raw)} columns (free-form limit 132)")
        line = strip_comment(raw).strip()
        if not line:
            continue
        if cur:
            if line.startswith("&"):
                line = line[1:].lstrip()
            ncont += 1
            if ncont > 255:
                raise FortranError(f"line {start}: more than 255 continuation lines")
        else:
            start, ncont = no, 0
        if line.endswith("&"):
            cur += line[:-1].rstrip() + " "
            continue
        out.append((start, (cur + line).strip()))
        cur = ""
    if cur:
        raise FortranError(f"line {start}: continuation without a following line")
    return out


def split_top(s: str) -> List[str]:
    """Split on commas outside parentheses / brackets / strings."""
    parts, depth, cur, quote = [], 0, "", None
    for ch in s:
        if quote:
            cur += ch
            if ch == quote:
                quote = None
            continue
        if ch in "'\"":
            quote = ch
        if ch in "([":
            depth += 1
        elif ch in ")]":
            depth -= 1
        if ch == "," and depth == 0:
            parts.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        parts.append(cur.strip())
    return parts


@dataclass
class Dummy:
    name: str
    base: str          # integer | real | character | type(c_ptr)
    kind: str          # c_int ... ("" for type(c_ptr))
    attrs: List[str]   # value | intent(in) ...
    shape: str         # "" | "*" | "6"


@dataclass
class Interface:
    name: str
    bind_name: str
    dummies: List[str]
    decls: Dict[str, Dummy] = field(default_factory=dict)
    result: Dummy = None
    imports: List[str] = field(default_factory=list)


@dataclass
class Module:
    name: str
    kinds: List[str]
    publics: List[str]
    interfaces: Dict[str, Interface]


def _ident(name: str, no: int) -> str:
    if not IDENT.match(name) or len(name) > 63:
        raise FortranError(f"line {no}: '{name}' is not a legal Fortran name (letter first, <= 63 characters)")
    return name.lower()


_DECL = re.compile(r"^(integer|real|character)\s*\(\s*kind\s*=\s*(\w+)\s*\)(.*?)::\s*(\w+)\s*(\((.*)\))?$|^type\s*\(\s*c_ptr\s*\)(.*?)::\s*(\w+)$", re.I)


def parse_declaration(stmt: str, no: int) -> Dummy:
    m = _DECL.match(stmt)
    if not m:
        raise FortranError(f"line {no}: unrecognised declaration '{stmt}'")
    if m.group(1):
        base, kind, attrs, name, shape = m.group(1).lower(), m.group(2).lower(), m.group(3), m.group(4), m.group(6) or ""
    else:
        base, kind, attrs, name, shape = "type(c_ptr)", "", m.group(7), m.group(8), ""
    attr_list = [a.strip().lower().replace(" ", "") for a in attrs.split(",") if a.strip()]
    for a in attr_list:
        if a not in ("value", "intent(in)", "intent(out)", "intent(inout)", "parameter"):
            raise FortranError(f"line {no}: attribute '{a}' is not one the generator emits")
    if "value" in attr_list and (shape or any(a.startswith("intent(out") or a.startswith("intent(inout") for a in attr_list)):
        raise FortranError(f"line {no}: VALUE on an array or an intent(out) dummy ('{name}')")
    return Dummy(_ident(name, no), base, kind, attr_list, shape.strip())


def parse_module(text: str) -> Module:
    st = statements(text)
    pos = 0

    def nxt():
        nonlocal pos
        if pos >= len(st):
            raise FortranError("unexpected end of file")
        pos += 1
        return st[pos - 1]

    no, s = nxt()
    m = re.match(r"^module\s+(\w+)$", s, re.I)
    if not m:
        raise FortranError(f"line {no}: expected 'module NAME', got '{s}'")
    mod = Module(_ident(m.group(1), no), [], [], {})
    no, s = nxt()
    m = re.match(r"^use\s+iso_c_binding\s*,\s*only\s*:\s*(.+)$", s, re.I)
    if not m:
        raise FortranError(f"line {no}: expected 'use iso_c_binding, only: ...'")
    mod.kinds = [_ident(k, no) for k in split_top(m.group(1))]
    no, s = nxt()
    if s.lower() != "implicit none":
        raise FortranError(f"line {no}: expected 'implicit none'")
    no, s = nxt()
    if s.lower() != "private":
        raise FortranError(f"line {no}: expected 'private'")
    while True:
        no, s = nxt()
        m = re.match(r"^public\s*::\s*(\w+)$", s, re.I)
        if not m:
            break
        mod.publics.append(_ident(m.group(1), no))
    if s.lower() != "interface":
        raise FortranError(f"line {no}: expected 'interface', got '{s}'")
    while True:
        no, s = nxt()
        if s.lower() == "end interface":
            break
        m = re.match(r"^function\s+(\w+)\s*\((.*?)\)\s*bind\s*\(\s*c\s*,\s*name\s*=\s*'(\w+)'\s*\)\s*result\s*\(\s*(\w+)\s*\)$", s, re.I)
        if not m:
            raise FortranError(f"line {no}: expected 'function NAME(...) bind(c, name=...) result(...)', got '{s[:80]}'")
        itf = Interface(_ident(m.group(1), no), m.group(3), [_ident(d, no) for d in split_top(m.group(2))] if m.group(2).strip() else [])
        result_name = _ident(m.group(4), no)
        if len(set(itf.dummies)) != len(itf.dummies):
            dup = sorted(d for d in set(itf.dummies) if itf.dummies.count(d) > 1)
            raise FortranError(f"line {no}: {itf.name}: dummy names clash (Fortran names are case-insensitive): {dup}")
        no, s = nxt()
        m = re.match(r"^import\s+(.+)$", s, re.I)
        if not m:
            raise FortranError(f"line {no}: {itf.name}: an interface body needs 'import' to see the kinds")
        itf.imports = [_ident(k, no) for k in split_top(m.group(1))]
        for k in itf.imports:
            if k not in mod.kinds:
                raise FortranError(f"line {no}: {itf.name}: imports '{k}', which the module does not use from iso_c_binding")
        no, s = nxt()
        if s.lower() != "implicit none":
            raise FortranError(f"line {no}: {itf.name}: expected 'implicit none'")
        while True:
            no, s = nxt()
            m = re.match(r"^end\s+function\s+(\w+)$", s, re.I)
            if m:
                if m.group(1).lower() != itf.name:
                    raise FortranError(f"line {no}: 'end function {m.group(1)}' closes 'function {itf.name}'")
                break
            d = parse_declaration(s, no)
            if d.kind and d.kind not in itf.imports:
                raise FortranError(f"line {no}: {itf.name}: kind '{d.kind}' is not imported")
            if d.base == "type(c_ptr)" and "c_ptr" not in itf.imports:
                raise FortranError(f"line {no}: {itf.name}: c_ptr is not imported")
            if d.name == result_name:
                if itf.result is not None or d.attrs or d.shape:
                    raise FortranError(f"line {no}: {itf.name}: bad result declaration")
                itf.result = d
            elif d.name in itf.decls:
                raise FortranError(f"line {no}: {itf.name}: '{d.name}' declared twice")
            elif d.name not in itf.dummies:
                raise FortranError(f"line {no}: {itf.name}: '{d.name}' is declared but is not a dummy argument")
            else:
                itf.decls[d.name] = d
        missing = [d for d in itf.dummies if d not in itf.decls]
        if missing:
            raise FortranError(f"{itf.name}: dummies without a declaration (implicit none): {missing}")
        if itf.result is None:
            raise FortranError(f"{itf.name}: result '{result_name}' is not declared")
        if itf.name in mod.interfaces:
            raise FortranError(f"{itf.name}: declared twice")
        mod.interfaces[itf.name] = itf
    no, s = nxt()
    m = re.match(r"^end\s+module\s+(\w+)$", s, re.I)
    if not m or m.group(1).lower() != mod.name:
        raise FortranError(f"line {no}: expected 'end module {mod.name}'")
    if pos != len(st):
        raise FortranError(f"line {st[pos][0]}: text after 'end module'")
    for p in mod.publics:
        if p not in mod.interfaces:
            raise FortranError(f"public :: {p} has no interface")
    for n in mod.interfaces:
        if n not in mod.publics:
            raise FortranError(f"interface {n} is private: callers cannot use it")
    return mod


# ---- cross-check against the C header ---------------------------------------------------------------------------

_C_SCALAR = {"int": ("integer", "c_int"), "int64_t": ("integer", "c_int64_t"), "float": ("real", "c_float"), "double": ("real", "c_double")}
_C_POINTEE = dict(_C_SCALAR, char=("character", "c_char"))


def c_prototypes(header_text: str) -> Dict[str, Tuple[str, List[Tuple[str, str]]]]:
    out = {}
    for m in re.finditer(r"^B2S_API\s+(.+?)\s*\b(\w+)\s*\(([^)]*)\)\s*;", header_text, re.M):
        ret, name, params = re.sub(r"\s*\*", "*", m.group(1).strip()), m.group(2), m.group(3).strip()
        plist = []
        if params and params != "void":
            for p in params.split(","):
                pm = re.match(r"\s*(.+?)\s*(\w+)\s*$", p)
                plist.append((re.sub(r"\s*\*", "*", pm.group(1)).strip(), pm.group(2)))
        out[name] = (ret, plist)
    return out


def check_against_header(mod: Module, header_text: str) -> int:
    protos = c_prototypes(header_text)
    checked = 0
    for name, itf in mod.interfaces.items():
        if itf.bind_name not in protos:
            raise FortranError(f"{name}: binds to '{itf.bind_name}', which include/b200stencil.h does not declare")
        ret, params = protos[itf.bind_name]
        if len(params) != len(itf.dummies):
            raise FortranError(f"{name}: {len(itf.dummies)} dummies, the C prototype has {len(params)} parameters")
        for (ctype, cname), dname in zip(params, itf.dummies):
            d = itf.decls[dname]
            if cname.lower() != dname:
                raise FortranError(f"{name}: dummy '{dname}' sits where the C prototype has '{cname}'")
            if ctype in _C_SCALAR:
                if (d.base, d.kind) != _C_SCALAR[ctype] or "value" not in d.attrs or d.shape:
                    raise FortranError(f"{name}: C passes '{ctype} {cname}' by value; Fortran declares {d}")
            elif ctype.endswith("**"):
                if d.base != "type(c_ptr)" or "value" in d.attrs:
                    raise FortranError(f"{name}: '{ctype} {cname}' must be a type(c_ptr) passed by reference; got {d}")
            elif ctype.endswith("*"):
                pointee = ctype[:-1].replace("const ", "").strip()
                if d.base == "type(c_ptr)":
                    if "value" not in d.attrs:
                        raise FortranError(f"{name}: '{ctype} {cname}' as type(c_ptr) must be passed by value; got {d}")
                else:
                    if pointee not in _C_POINTEE or (d.base, d.kind) != _C_POINTEE[pointee] or "value" in d.attrs:
                        raise FortranError(f"{name}: '{ctype} {cname}' by reference needs the interoperable type of '{pointee}'; got {d}")
                    if ctype.startswith("const ") and "intent(in)" not in d.attrs:
                        raise FortranError(f"{name}: 'const' pointer '{cname}' should be intent(in); got {d.attrs}")
            else:
                raise FortranError(f"{name}: no rule for C type '{ctype}'")
        want = {"int": ("integer", "c_int"), "int64_t": ("integer", "c_int64_t"), "const char*": ("type(c_ptr)", "")}[ret]
        if (itf.result.base, itf.result.kind) != want:
            raise FortranError(f"{name}: C returns '{ret}'; Fortran result is {itf.result}")
        checked += 1
    missing = sorted(set(protos) - {i.bind_name for i in mod.interfaces.values()})
    if missing:
        raise FortranError(f"C entry points without a Fortran interface: {missing}")
    return checked


# ---- the acceptance program -------------------------------------------------------------------------------------


def check_program(text: str, mod: Module) -> List[str]:
    """Line rules + block structure of the example program; every module function it calls must be imported with
    ``use ..., only:`` and called with as many actual arguments as the interface has dummies.  Returns the calls."""
    st = statements(text)
    no, s = st[0]
    m = re.match(r"^program\s+(\w+)$", s, re.I)
    if not m:
        raise FortranError(f"line {no}: expected 'program NAME'")
    pname = m.group(1).lower()
    no, s = st[-1]
    m = re.match(r"^end\s+program\s+(\w+)$", s, re.I)
    if not m or m.group(1).lower() != pname:
        raise FortranError(f"line {no}: expected 'end program {pname}'")
    used: List[str] = []
    for no, s in st:
        m = re.match(rf"^use\s+{mod.name}\s*,\s*only\s*:\s*(.+)$", s, re.I)
        if m:
            used = [_ident(u, no) for u in split_top(m.group(1))]
            for u in used:
                if u not in mod.publics:
                    raise FortranError(f"line {no}: '{u}' is not a public entity of {mod.name}")
    if not any(s.lower() == "implicit none" for _, s in st):
        raise FortranError("the program has no 'implicit none'")
    calls = []
    for no, s in st:
        for name, itf in mod.interfaces.items():
            for m in re.finditer(rf"\b{name}\s*\(", s, re.I):
                if re.match(r"^use\b", s, re.I):
                    continue
                depth, i = 1, m.end()
                while i < len(s) and depth:
                    depth += s[i] == "("
                    depth -= s[i] == ")"
                    i += 1
                args = split_top(s[m.end():i - 1])
                if name not in used:
                    raise FortranError(f"line {no}: calls {name} without importing it from {mod.name}")
                if len(args) != len(itf.dummies):
                    raise FortranError(f"line {no}: {name} called with {len(args)} arguments, the interface has {len(itf.dummies)}")
                calls.append(name)
    return calls

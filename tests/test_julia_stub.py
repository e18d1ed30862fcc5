"""julia/VoronoiRTB200.jl cannot be executed here (no Julia in the image); at least every `ccall` in it must name a symbol
that include/vrt.h declares, with the same number of arguments."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def split_top_level(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def header_arity():
    h = open(os.path.join(ROOT, "include", "vrt.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    ar = {}
    for m in re.finditer(r"\b(?:int|void|const char\*)\s+(vrt_[A-Za-z_0-9]+)\s*\(([^;]*?)\)\s*;", h, flags=re.S):
        args = m.group(2).strip()
        ar[m.group(1)] = 0 if args in ("", "void") else len(split_top_level(args))
    return ar


def test_every_ccall_matches_the_header():
    jl = open(os.path.join(ROOT, "julia", "VoronoiRTB200.jl")).read()
    ar = header_arity()
    assert len(ar) >= 35
    seen = 0
    for m in re.finditer(r"ccall\(\(:(vrt_[A-Za-z_0-9]+),\s*libvrt\),\s*([A-Za-z]+),\s*\(", jl):
        name = m.group(1)
        i = m.end()                      # just after the opening parenthesis of the type tuple
        depth, j = 1, i
        while depth:
            depth += {"(": 1, ")": -1}.get(jl[j], 0)
            j += 1
        types = split_top_level(jl[i:j - 1])
        types = [t for t in types if t]
        assert name in ar, f"{name} is not declared in include/vrt.h"
        assert len(types) == ar[name], f"{name}: {len(types)} ccall argument types, header has {ar[name]}"
        seen += 1
    assert seen >= 10

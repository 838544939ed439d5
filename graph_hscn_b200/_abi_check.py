"""Parses include/ghscn.h and cross-checks it against _lib.SIGNATURES (names, arity and argument type classes)."""
from __future__ import annotations

import re
from pathlib import Path
from typing import Dict, List

HEADER = Path(__file__).resolve().parent.parent / "include" / "ghscn.h"


def _type_class(decl: str) -> str:
    """'const float* x' -> 'P', 'int64_t n' -> 'I64', 'double lr' -> 'F64', ..."""
    decl = decl.strip()
    if "*" in decl or decl.startswith("ghscn_stream_t"):
        return "P"
    for key, cls in (("int64_t", "I64"), ("int32_t", "I32"), ("size_t", "SZ"), ("double", "F64"), ("float", "F32"),
                     ("int", "I32")):
        if re.search(rf"\b{key}\b", decl):
            return cls
    return "?"


def header_signatures() -> Dict[str, List[str]]:
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    out: Dict[str, List[str]] = {}
    for m in re.finditer(r"GHSCN_API\s+[\w\s\*]+?\b(ghscn_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        out[name] = [] if args in ("", "void") else [_type_class(a) for a in args.split(",")]
    return out


def header_functions() -> Dict[str, int]:
    return {n: len(a) for n, a in header_signatures().items()}


def check() -> None:
    from . import _lib
    names = {_lib.P: "P", _lib.I64: "I64", _lib.I32: "I32", _lib.F32: "F32", _lib.F64: "F64", _lib.SZ: "SZ"}
    hdr = header_signatures()
    sigs = _lib.SIGNATURES
    missing = sorted(set(hdr) - set(sigs))
    extra = sorted(set(sigs) - set(hdr))
    bad = sorted(n for n in hdr if n in sigs and len(hdr[n]) != len(sigs[n][1]))
    wrong = sorted(n for n in hdr if n in sigs and n not in bad
                   and [names.get(t, "?") for t in sigs[n][1]] != hdr[n])
    if missing or extra or bad or wrong:
        raise AssertionError(f"ABI mismatch: not bound {missing}, not declared {extra}, arity differs {bad}, "
                             f"argument types differ {wrong}")

"""Parses include/ghscn.h and cross-checks it against _lib.SIGNATURES (names and arity)."""
from __future__ import annotations

import re
from pathlib import Path
from typing import Dict

HEADER = Path(__file__).resolve().parent.parent / "include" / "ghscn.h"


def header_functions() -> Dict[str, int]:
    text = HEADER.read_text()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    out: Dict[str, int] = {}
    for m in re.finditer(r"GHSCN_API\s+[\w\s\*]+?\b(ghscn_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        out[name] = 0 if args in ("", "void") else args.count(",") + 1
    return out


def check() -> None:
    from ._lib import SIGNATURES
    hdr = header_functions()
    missing = sorted(set(hdr) - set(SIGNATURES))
    extra = sorted(set(SIGNATURES) - set(hdr))
    bad = sorted(n for n in hdr if n in SIGNATURES and hdr[n] != len(SIGNATURES[n][1]))
    if missing or extra or bad:
        raise AssertionError(f"ABI mismatch: not bound {missing}, not declared {extra}, arity differs {bad}")

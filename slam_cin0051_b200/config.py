"""Reader for the OpenCV-FileStorage YAML files the reference's constructors parse.

The reference reads its configuration through cv::FileStorage (feature_detector.hpp:54-94,
feature_matcher.cpp:19-59, common.hpp:78-95).  Those files are `%YAML:1.0` documents holding flat
`key: value` scalars, inline `[a, b]` sequences and `!!opencv-matrix` blocks; this module parses exactly
that subset without depending on OpenCV.
"""
from __future__ import annotations

import os
import re

import numpy as np


def _scalar(tok: str):
    tok = tok.strip()
    if (tok.startswith('"') and tok.endswith('"')) or (tok.startswith("'") and tok.endswith("'")):
        return tok[1:-1]
    try:
        return int(tok)
    except ValueError:
        pass
    try:
        return float(tok)
    except ValueError:
        return tok


def _strip_comment(line: str) -> str:
    out, quote = [], None
    for ch in line:
        if quote:
            if ch == quote:
                quote = None
        elif ch in "\"'":
            quote = ch
        elif ch == "#":
            break
        out.append(ch)
    return "".join(out).rstrip()


def read_yaml(path) -> dict:
    """Returns {key: scalar | list | np.ndarray}.  Raises RuntimeError if the file cannot be opened."""
    path = os.fspath(path)
    try:
        with open(path, "r") as f:
            text = f.read()
    except OSError as e:
        raise RuntimeError(f"Could not open file: {path}") from e
    lines = [_strip_comment(l) for l in text.splitlines()]
    lines = [l for l in lines if l.strip() and not l.startswith("%") and l.strip() != "---"]
    out: dict = {}
    i = 0
    while i < len(lines):
        line = lines[i]
        m = re.match(r"^([A-Za-z_][A-Za-z0-9_]*)\s*:\s*(.*)$", line)
        if not m:
            i += 1
            continue
        key, rest = m.group(1), m.group(2).strip()
        if rest.startswith("!!opencv-matrix"):
            block = {}
            i += 1
            while i < len(lines) and lines[i].startswith((" ", "\t")):
                sub = lines[i].strip()
                sm = re.match(r"^([a-z]+)\s*:\s*(.*)$", sub)
                if sm:
                    k, v = sm.group(1), sm.group(2).strip()
                    if k == "data":
                        while v.count("[") > v.count("]") and i + 1 < len(lines):
                            i += 1
                            v += " " + lines[i].strip()
                        block["data"] = [float(t) for t in v.strip("[] ").split(",") if t.strip()]
                    else:
                        block[k] = _scalar(v)
                i += 1
            dt = {"d": np.float64, "f": np.float32, "i": np.int32, "u": np.uint8}.get(str(block.get("dt", "d")), np.float64)
            out[key] = np.array(block.get("data", []), dtype=dt).reshape(int(block.get("rows", 1)), int(block.get("cols", 1)))
            continue
        if rest.startswith("["):
            while rest.count("[") > rest.count("]") and i + 1 < len(lines):
                i += 1
                rest += " " + lines[i].strip()
            out[key] = [_scalar(t) for t in rest.strip("[] ").split(",") if t.strip()]
        elif rest:
            out[key] = _scalar(rest)
        i += 1
    return out


def get_int(cfg: dict, key: str) -> int:
    """cv::FileNode >> int semantics: missing key -> 0, real -> rounded."""
    v = cfg.get(key, 0)
    if isinstance(v, float):
        return int(round(v))
    return int(v) if isinstance(v, int) else 0


def get_float(cfg: dict, key: str) -> float:
    v = cfg.get(key, 0.0)
    return float(v) if isinstance(v, (int, float)) else 0.0


def get_str(cfg: dict, key: str) -> str:
    v = cfg.get(key, "")
    return v if isinstance(v, str) else str(v)

"""Shared helpers for the tests: golden-fixture access, flag parsing, text-DB parsing."""
import json
import os
import tarfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")

_cache = {}


def golden_case(name, tmp_root):
    """Unpack tests/golden/<name>.tar.xz once per session; returns (dir, manifest)."""
    if name not in _cache:
        d = os.path.join(str(tmp_root), name)
        os.makedirs(d, exist_ok=True)
        with tarfile.open(os.path.join(GOLDEN, name + ".tar.xz")) as tf:
            tf.extractall(d, filter="data")
        _cache[name] = (d, json.load(open(os.path.join(GOLDEN, name + ".json"))))
    return _cache[name]


def parse_flags(args):
    """argv tail -> dict(flag -> string value), as the reference's Parameters would see it."""
    out = {}
    i = 0
    while i < len(args):
        out[args[i]] = args[i + 1]
        i += 2
    return out


def multi(v, nucl):
    """MultiParam 'nucl:0.200,aa:0.000' (mm/commons/MultiParam.cpp) -> the applicable value."""
    if ":" not in v:
        return v
    d = dict(x.split(":") for x in v.split(","))
    return d["nucl" if nucl else "aa"]


def parse_pref_entry(key, entry):
    """prefilter entry bytes -> list of (target, score, diag) WITHOUT the leading self line."""
    lines = entry.decode().splitlines()
    assert lines[0] == "%d\t0\t0" % key, (key, lines[:2])
    return [tuple(int(x) for x in ln.split("\t")) for ln in lines[1:]]

"""Shared helpers for the parity tests."""
import hashlib
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
ALGOS = {"quicked": 0, "windowed": 1, "banded": 2, "hirschberg": 3}


def sha(s):
    return hashlib.sha1((s or "").encode()).hexdigest()[:16]


def load_golden(name):
    return json.load(open(os.path.join(GOLDEN, name)))


def golden_kw(gold, aname):
    kw = dict(gold["params"][aname])
    kw.setdefault("algo", ALGOS.get(aname, 0))
    return kw


def expand_rle(cigar):
    ops, num = [], 0
    for ch in cigar:
        if ch.isdigit():
            num = num * 10 + int(ch)
        else:
            ops.append(ch * num)
            num = 0
    return "".join(ops)


def replay(ops, pattern, text):
    """cigar_check_alignment (quicked_utils/src/cigar.c:363-434): the ops must replay onto the pair.
    Returns the edit cost, raises AssertionError on an invalid alignment."""
    i = j = cost = 0
    for op in ops:
        if op == "M":
            assert pattern[i] == text[j], f"M over a mismatch at p[{i}] t[{j}]"
            i += 1; j += 1
        elif op == "X":
            assert pattern[i] != text[j], f"X over a match at p[{i}] t[{j}]"
            i += 1; j += 1; cost += 1
        elif op == "D":
            i += 1; cost += 1
        elif op == "I":
            j += 1; cost += 1
        else:
            raise AssertionError(f"unknown op {op!r}")
    assert i == len(pattern) and j == len(text), "alignment does not span both sequences"
    return cost

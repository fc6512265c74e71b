"""Assertions for the reference known-answer tables (tests/ref_cases.py), implementation-agnostic."""
import math

import numpy as np

from ref_cases import EPS
from frankensearch_b200.types import fnv1a_hash


def check_scan_case(impl, case):
    rows = np.array([v for _, v in case["rows"]], dtype=np.float32)
    ids = [d for d, _ in case["rows"]]
    got_rows, got_scores = impl.search(rows, case["query"], case["k"], tombstones=case.get("tombstones"))
    if "expect_ids" in case:
        assert [ids[r] for r in got_rows] == case["expect_ids"], case["name"]
    if "expect_rows" in case:
        assert got_rows == case["expect_rows"], case["name"]
    if case.get("all_nan"):
        assert len(got_scores) == len(case["expect_rows"]) and all(math.isnan(float(s)) for s in got_scores)


def check_rrf_case(impl, case):
    out = impl.rrf(case["lexical"], case["semantic"], case["limit"], case.get("offset", 0), case.get("k", 60.0),
                   case.get("w_lex", 1.0), case.get("w_sem", 1.0), case.get("tiebreak", "LexicalThenId"))
    name = case["name"]
    if "expect" in case:
        assert len(out) == len(case["expect"]), name
        for (doc, score, *_), (edoc, escore) in zip(out, case["expect"]):
            assert doc == edoc, name
            assert abs(score - escore) < 1e-12, (name, score, escore)
    if "expect_both" in case:
        assert [o[4] for o in out] == case["expect_both"], name
    if "expect_order" in case:
        assert [o[0] for o in out] == case["expect_order"], name
    if "expect_len" in case:
        assert len(out) == case["expect_len"], name
    if "expect_score" in case:
        got = {o[0]: o[1] for o in out}
        for doc, s in case["expect_score"].items():
            assert abs(got[doc] - s) < 1e-12, name
    if "expect_hash_order" in case:
        a, b = case["expect_hash_order"]
        first = a if fnv1a_hash(a.encode()) <= fnv1a_hash(b.encode()) else b
        assert out[0][0] == first, name
        assert abs(out[0][1] - out[1][1]) < 1e-12


def check_blend_case(impl, case):
    out = impl.blend(case["fast"], case["quality"], case["alpha"])
    name = case["name"]
    score = {d: float(s) for d, _, s in out}
    if "expect_score" in case:
        for d, s in case["expect_score"].items():
            assert abs(score[d] - s) <= EPS, (name, d, score[d], s)
    if case.get("expect_finite"):
        assert all(math.isfinite(v) for v in score.values()), name
    if "expect_order" in case:
        assert [d for d, _, _ in out] == case["expect_order"], name
    if "expect_len" in case:
        assert len(out) == case["expect_len"], name
    if "expect_equal" in case:
        a, b = case["expect_equal"]
        assert abs(score[a] - score[b]) <= EPS, name
    if "same_as_alpha" in case:
        ref = impl.blend(case["fast"], case["quality"], case["same_as_alpha"])
        assert [(d, i, np.float32(s).view(np.uint32)) for d, i, s in out] == \
               [(d, i, np.float32(s).view(np.uint32)) for d, i, s in ref], name


def check_blend_aligned(impl, table):
    fast, scores = table["fast"], table["scores"]
    for alpha in table["alphas"]:
        subset = [(d, i, s) for (d, i, _), s in zip(fast, scores) if s is not None]
        materialized = impl.blend(fast, subset, alpha)
        aligned = impl.blend_aligned(fast, scores, alpha)
        assert len(materialized) == len(aligned)
        for m, a in zip(materialized, aligned):
            assert m[0] == a[0] and m[1] == a[1]
            assert np.float32(m[2]).view(np.uint32) == np.float32(a[2]).view(np.uint32), (alpha, m, a)

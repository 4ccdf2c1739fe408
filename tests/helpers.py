"""Shared helpers of the test-suite (test infrastructure; may use oracle/)."""
from __future__ import annotations

import json
import tempfile
from pathlib import Path

import numpy as np

from leann_b200 import csr, synth

GOLDEN_CASES = [  # (ef, beam, batch, check_rel, k) — keep in sync with tests/golden/make_golden.py
    (64, 1, 0, 1, 10), (16, 1, 0, 1, 10), (8, 1, 0, 1, 20), (32, 4, 0, 1, 10), (32, 16, 0, 1, 5),
    (32, 1, 24, 1, 10), (24, 2, 0, 0, 10), (128, 1, 0, 1, 1), (64, 3, 40, 0, 7),
]


def golden_key(tag, ef, beam, batch, cr, k):
    return f"{tag}_ef{ef}_b{beam}_bs{batch}_cr{cr}_k{k}"


from leann_b200.tooling import open_encoder_only, recall_at_k, stub_graph, write_leann_index  # noqa: E402,F401

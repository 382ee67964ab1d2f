"""Host-side mirror of the reference's src/utils/common_utils.py (the live part, :62-103).

parse_encode_dict is pure host logic (pinned by the reference's tests/test_utils.py:11-32).
encode_sequence keeps the reference's signature and return shape (L, C) but computes the one-hot
on the GPU with the dataset kernel (hb_encode_haplotypes with no variant records); the reference's
own implementation returns all zeros (pandas labels b'A' vs 'A') -- the intent
(tests/test_utils.py:38-65: shape (L, 5), one 1 per row) is what is implemented.
"""
from __future__ import annotations

import numpy as np


def parse_encode_dict(encode_spec):
    """common_utils.py:62-79, verbatim semantics."""
    if not encode_spec:
        return {"A": 0, "C": 1, "G": 2, "T": 3, "N": 4}
    elif isinstance(encode_spec, (list, tuple, str)):
        return {base: i for i, base in enumerate(encode_spec)}
    elif isinstance(encode_spec, dict):
        return encode_spec
    else:
        raise TypeError("Please input as dict, list or string!")


def build_lut(encode_spec, ignore_case: bool = True) -> np.ndarray:
    """256-entry byte -> class index table: ACGT (and lower case when ignore_case) map to their
    encode_spec index, everything else to 'N' (common_utils.py:84-85); -1 (all-zero row) when the
    spec has no such key."""
    spec = parse_encode_dict(encode_spec)
    lut = np.full(256, spec.get("N", -1), dtype=np.int8)
    for b in "ACGT":
        if b in spec:
            lut[ord(b)] = spec[b]
            if ignore_case:
                lut[ord(b.lower())] = spec[b]
    return lut


def encode_sequence(seq_data, encode_spec=None, ignore_case=True):
    """common_utils.py:88-103: str or |S1 ndarray -> (L, C) one-hot, columns in encode_spec order."""
    if isinstance(seq_data, str):
        raw = np.frombuffer(seq_data.encode("latin-1"), np.uint8)
    elif isinstance(seq_data, np.ndarray):
        if seq_data.dtype != "|S1":
            seq_data = seq_data.astype("|S1")
        raw = seq_data.view(np.uint8)
    else:
        raise TypeError("Please input as string or numpy array!")
    from .haplotype_dataset import onehot_windows
    spec = parse_encode_dict(encode_spec)
    out = onehot_windows([raw], len(raw), spec, ignore_case=ignore_case)
    return out[0].to("cpu").numpy().astype(np.uint8)

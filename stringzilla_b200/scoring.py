"""Substitution models for the weighted engines, in the reference's compact layout: a 256-entry byte -> class map and
a 32x32 int8 class table (reference serial.hpp:186-200). `blosum62()` plays the role of the reference's
`error_costs_32x32_t::blosum62()` (serial.hpp:223-249): the standard NCBI BLOSUM62 values, restated here."""

from __future__ import annotations

import numpy as np

PROTEIN_ALPHABET = "ARNDCQEGHILKMFPSTWYV"  # the 20 amino acids the synthetic protein corpora draw from
_BLOSUM62_ORDER = "ARNDCQEGHILKMFPSTWYVBZX*"
_BLOSUM62_ROWS = """
 4 -1 -2 -2  0 -1 -1  0 -2 -1 -1 -1 -1 -2 -1  1  0 -3 -2  0 -2 -1  0 -4
-1  5  0 -2 -3  1  0 -2  0 -3 -2  2 -1 -3 -2 -1 -1 -3 -2 -3 -1  0 -1 -4
-2  0  6  1 -3  0  0  0  1 -3 -3  0 -2 -3 -2  1  0 -4 -2 -3  3  0 -1 -4
-2 -2  1  6 -3  0  2 -1 -1 -3 -4 -1 -3 -3 -1  0 -1 -4 -3 -3  4  1 -1 -4
 0 -3 -3 -3  9 -3 -4 -3 -3 -1 -1 -3 -1 -2 -3 -1 -1 -2 -2 -1 -3 -3 -2 -4
-1  1  0  0 -3  5  2 -2  0 -3 -2  1  0 -3 -1  0 -1 -2 -1 -2  0  3 -1 -4
-1  0  0  2 -4  2  5 -2  0 -3 -3  1 -2 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
 0 -2  0 -1 -3 -2 -2  6 -2 -4 -4 -2 -3 -3 -2  0 -2 -2 -3 -3 -1 -2 -1 -4
-2  0  1 -1 -3  0  0 -2  8 -3 -3 -1 -2 -1 -2 -1 -2 -2  2 -3  0  0 -1 -4
-1 -3 -3 -3 -1 -3 -3 -4 -3  4  2 -3  1  0 -3 -2 -1 -3 -1  3 -3 -3 -1 -4
-1 -2 -3 -4 -1 -2 -3 -4 -3  2  4 -2  2  0 -3 -2 -1 -2 -1  1 -4 -3 -1 -4
-1  2  0 -1 -3  1  1 -2 -1 -3 -2  5 -1 -3 -1  0 -1 -3 -2 -2  0  1 -1 -4
-1 -1 -2 -3 -1  0 -2 -3 -2  1  2 -1  5  0 -2 -1 -1 -1 -1  1 -3 -1 -1 -4
-2 -3 -3 -3 -2 -3 -3 -3 -1  0  0 -3  0  6 -4 -2 -2  1  3 -1 -3 -3 -1 -4
-1 -2 -2 -1 -3 -1 -1 -2 -2 -3 -3 -1 -2 -4  7 -1 -1 -4 -3 -2 -2 -1 -2 -4
 1 -1  1  0 -1  0  0  0 -1 -2 -2  0 -1 -2 -1  4  1 -3 -2 -2  0  0  0 -4
 0 -1  0 -1 -1 -1 -1 -2 -2 -1 -1 -1 -1 -2 -1  1  5 -2 -2  0 -1 -1  0 -4
-3 -3 -4 -4 -2 -2 -3 -2 -2 -3 -2 -3 -1  1 -4 -3 -2 11  2 -3 -4 -3 -2 -4
-2 -2 -2 -3 -2 -1 -2 -3  2 -1 -1 -2 -1  3 -3 -2 -2  2  7 -1 -3 -2 -1 -4
 0 -3 -3 -3 -1 -2 -2 -3 -3  3  1 -2  1 -1 -2 -2  0 -3 -1  4 -3 -2 -1 -4
-2 -1  3  4 -3  0  1 -1  0 -3 -4  0 -3 -3 -2  0 -1 -4 -3 -3  4  1 -1 -4
-1  0  0  1 -3  3  4 -2  0 -3 -3  1 -1 -3 -1  0 -1 -3 -2 -2  1  4 -1 -4
 0 -1 -1 -1 -2 -1 -1 -1 -1 -1 -1 -1 -1 -1 -2  0  0 -2 -1 -1 -1 -1 -1 -4
-4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4 -4  1
"""


def blosum62() -> tuple[np.ndarray, np.ndarray]:
    """Returns (byte_to_class[256] uint8, class_substitution_costs[32, 32] int8). Class 0 is the catch-all bucket
    (scored like `X`); residue i of "ARNDCQEGHILKMFPSTWYVBZX*" takes class i + 1, upper and lower case alike."""
    matrix = np.array([[int(v) for v in row.split()] for row in _BLOSUM62_ROWS.strip().splitlines()], dtype=np.int8)
    assert matrix.shape == (24, 24) and (matrix == matrix.T).all()
    byte_to_class = np.zeros(256, dtype=np.uint8)
    for index, residue in enumerate(_BLOSUM62_ORDER):
        byte_to_class[ord(residue)] = index + 1
        byte_to_class[ord(residue.lower())] = index + 1
    costs = np.zeros((32, 32), dtype=np.int8)
    costs[1:25, 1:25] = matrix
    unknown = _BLOSUM62_ORDER.index("X")
    costs[0, 1:25] = matrix[unknown]
    costs[1:25, 0] = matrix[:, unknown]
    costs[0, 0] = matrix[unknown, unknown]
    return byte_to_class, costs


def uniform(alphabet: bytes, match: int, mismatch: int) -> tuple[np.ndarray, np.ndarray]:
    """A match/mismatch model over `alphabet` (<= 31 symbols) in the weighted engines' layout; every other byte falls
    into class 0, which mismatches everything including itself."""
    if len(alphabet) > 31:
        raise ValueError("at most 31 symbols fit beside the catch-all class")
    byte_to_class = np.zeros(256, dtype=np.uint8)
    for index, symbol in enumerate(alphabet):
        byte_to_class[symbol] = index + 1
    costs = np.full((32, 32), mismatch, dtype=np.int8)
    for index in range(1, len(alphabet) + 1):
        costs[index, index] = match
    return byte_to_class, costs

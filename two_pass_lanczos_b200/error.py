"""Error types of the reference (src/error.rs:16-58, src/utils/data_loader.rs:16-43) on the Python side."""
from __future__ import annotations


class LanczosError(Exception):
    """`LanczosError(LanczosErrorKind)`; `kind` names the variant, str(e) equals the reference's Display."""

    KINDS = {1: "Breakdown", 2: "DimensionMismatch", 3: "InputError", 4: "ParameterMismatch", 5: "EvdError",
             6: "SolverError", 7: "Panic"}

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code
        self.kind = self.KINDS.get(code, f"Status{code}")
        self.message = message


class DataLoaderError(Exception):
    KINDS = {101: "Io", 102: "ParseInt", 103: "ParseFloat", 104: "ProblemLineMissing", 105: "UnexpectedEof",
             106: "ArcCountMismatch", 107: "SparseMatrixConstructionError", 108: "InvalidDimacsNodeIndex",
             109: "MalformedArcLine"}

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code
        self.kind = self.KINDS.get(code, f"Status{code}")
        self.message = message


class CudaError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code
        self.message = message


def raise_for_status(code: int, message: str):
    if code == 0:
        return
    if 100 < code < 200:
        raise DataLoaderError(code, message)
    if code >= 200:
        raise CudaError(code, message)
    raise LanczosError(code, message)

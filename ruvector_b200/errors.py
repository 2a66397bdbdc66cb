"""GnnError mirror (crates/ruvector-gnn/src/error.rs:9-60)."""
from __future__ import annotations


class GnnError(Exception):
    """Base class; ``str(e)`` matches the reference's ``#[error(...)]`` text."""

    code = -1


class DimensionMismatch(GnnError):
    code = 1


class InvalidShape(GnnError):
    code = 2

    def __str__(self):
        return f"Invalid tensor shape: {self.args[0]}"


class LayerConfig(GnnError):
    """``GnnError::LayerConfig`` -- raised by ``RuvectorLayer(...)`` (layer.rs:101-106,353-358)."""

    code = 3

    def __str__(self):
        return f"Layer configuration error: {self.args[0]}"


class InvalidInput(GnnError):
    code = 7

    def __str__(self):
        return f"Invalid input: {self.args[0]}"


class CudaError(GnnError):
    code = 100


class NcclError(GnnError):
    code = 101


class OutOfMemory(GnnError):
    code = 102


_BY_CODE = {c.code: c for c in (DimensionMismatch, InvalidShape, LayerConfig, InvalidInput, CudaError, NcclError, OutOfMemory)}


def from_status(code: int, message: str) -> GnnError:
    return _BY_CODE.get(code, GnnError)(message)

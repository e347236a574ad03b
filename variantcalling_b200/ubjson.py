"""Minimal UBJSON (draft 12) reader -- the binary flavour of xgboost's model files
(``Booster.save_model("model.ubj")``, the default since xgboost 2.1): big-endian scalars, strings as
length + UTF-8, containers optionally "optimised" with a ``$`` element type and a ``#`` count
(xgboost writes its numeric vectors that way).  Decoding only; the result is the same dict the JSON
flavour gives, which ``model_compiler._lower_xgboost_json`` lowers.

PARITY UNPINNED: xgboost is not installed in the build container; the format follows the published
specification (ubjson.org) and is exercised in the tests by an independent encoder.
"""
from __future__ import annotations

import struct

import numpy as np

_SCALARS = {"i": (">b", 1), "U": (">B", 1), "I": (">h", 2), "l": (">i", 4), "L": (">q", 8), "d": (">f", 4), "D": (">d", 8)}
_NP = {"i": ">i1", "U": ">u1", "I": ">i2", "l": ">i4", "L": ">i8", "d": ">f4", "D": ">f8"}


class UbjsonError(ValueError):
    pass


class _Reader:
    def __init__(self, data: bytes):
        self.b, self.at = memoryview(data), 0

    def take(self, n: int) -> memoryview:
        if self.at + n > len(self.b):
            raise UbjsonError("truncated UBJSON document")
        out = self.b[self.at:self.at + n]
        self.at += n
        return out

    def marker(self) -> str:
        while True:
            m = chr(self.take(1)[0])
            if m != "N":  # no-op
                return m

    def length(self) -> int:
        m = self.marker()
        if m not in "iUIlL":
            raise UbjsonError(f"bad length marker {m!r}")
        n = struct.unpack(_SCALARS[m][0], self.take(_SCALARS[m][1]))[0]
        if n < 0:
            raise UbjsonError("negative length")
        return n

    def string(self) -> str:
        return bytes(self.take(self.length())).decode()

    def value(self, m: str | None = None):  # noqa: C901, PLR0911, PLR0912
        m = m or self.marker()
        if m in _SCALARS:
            fmt, size = _SCALARS[m]
            return struct.unpack(fmt, self.take(size))[0]
        if m == "Z":
            return None
        if m == "T":
            return True
        if m == "F":
            return False
        if m == "C":
            return chr(self.take(1)[0])
        if m == "S":
            return self.string()
        if m == "H":
            text = self.string()
            return float(text) if any(c in text for c in ".eE") else int(text)
        if m in "[{":
            return self.container(m == "{")
        raise UbjsonError(f"unknown UBJSON marker {m!r} at byte {self.at - 1}")

    def container(self, is_object: bool):
        etype, count = None, None
        nxt = chr(self.b[self.at]) if self.at < len(self.b) else ""
        if nxt == "$":
            self.at += 1
            etype = chr(self.take(1)[0])
            if chr(self.take(1)[0]) != "#":
                raise UbjsonError("a typed container needs a count")
            count = self.length()
        elif nxt == "#":
            self.at += 1
            count = self.length()
        if not is_object and etype in _NP and count is not None:  # numeric vector in one piece
            size = np.dtype(_NP[etype]).itemsize
            return np.frombuffer(self.take(count * size), dtype=_NP[etype]).tolist()
        out = {} if is_object else []
        end = "}" if is_object else "]"
        i = 0
        while True:
            if count is not None:
                if i == count:
                    break
            elif chr(self.b[self.at]) == end:
                self.at += 1
                break
            if is_object:
                key = self.string()
                out[key] = self.value(etype)
            else:
                out.append(self.value(etype))
            i += 1
        return out


def loads(data: bytes):
    r = _Reader(data)
    doc = r.value()
    return doc


def load(path: str):
    with open(path, "rb") as fh:
        return loads(fh.read())

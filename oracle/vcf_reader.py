"""ORACLE (test infrastructure, not product code).

A pure-Python VCF reader that hands out records with the value semantics of
``pysam.VariantFile`` (pysam 0.22.1 over htslib 1.20 -- the un-vendored third
party engine behind the reference's loader).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it.

PARITY STATUS: *unpinned* for the htslib typed-decode boundary.  pysam/htslib
are not installed in the build container and every reference test that pins
this boundary (``ugbio_utils/src/core/tests/unit/vcfbed/test_vcftools.py:110-``)
runs on git-LFS fixtures that are pointer stubs here.  (Only the scalar-versus-tuple rule and the
Integer / Float / String typing are pinned, by the handful of literal expectations in the reference's
in-code CNV unit tests -- see tests/test_oracle_golden.py.)  The behaviour below is
restated from the VCF 4.2 specification plus the documented htslib/pysam rules
(SURVEY.md Appendix B):

* the typed value of a tag is decided by the header ``Type``/``Number``;
* ``Number=1`` -> scalar, any other ``Number`` -> tuple (also for one element);
* ``Type=Flag`` -> ``True``; ``Integer`` -> int; ``Float`` -> the Python float
  of a **float32** (htslib stores INFO/FORMAT floats and QUAL as float32, i.e.
  ``float(float32(strtod(text)))``); ``String`` -> str, comma-split when
  ``Number != 1``;
* a ``.`` numeric element -> ``None``; a missing trailing FORMAT sub-field ->
  ``None`` / ``(None,)``;
* ``GT`` -> tuple of allele indices, ``None`` for ``.``, phasing ignored;
* ``ALT`` ``.`` -> ``alleles == (REF,)``; ``ID`` ``.`` -> ``None``;
  ``FILTER`` ``.`` -> no keys, ``PASS`` -> ``["PASS"]``.

Call sites of the reference that consume these semantics:
``ugbio_utils/src/core/ugbio_core/vcfbed/vcftools.py:63-89`` and
``ugbio_utils/src/filtering/ugbio_filtering/filter_variants_pipeline.py:106-228``.
"""
from __future__ import annotations

import gzip
import re
from collections import OrderedDict

import numpy as np

_META_RE = re.compile(r"^##(INFO|FORMAT|FILTER|contig)=<(.*)>\s*$")


def _split_meta(body: str) -> dict:
    """Split ``ID=X,Number=1,Description="a, b"`` honouring quotes."""
    out, key, buf, in_q, is_key = {}, "", [], False, True
    for ch in body:
        if in_q:
            if ch == '"':
                in_q = False
            else:
                buf.append(ch)
        elif ch == '"':
            in_q = True
        elif ch == "=" and is_key:
            key, buf, is_key = "".join(buf), [], False
        elif ch == ",":
            out[key] = "".join(buf)
            key, buf, is_key = "", [], True
        else:
            buf.append(ch)
    if key:
        out[key] = "".join(buf)
    return out


class OracleHeader:
    """Typed view of the ``##`` meta lines (what ``pysam.VariantHeader`` exposes)."""

    def __init__(self, lines: list[str]):
        self.lines = [ln.rstrip("\n") for ln in lines]
        self.info: "OrderedDict[str, tuple[str, str]]" = OrderedDict()
        self.formats: "OrderedDict[str, tuple[str, str]]" = OrderedDict()
        self.filters: "OrderedDict[str, str]" = OrderedDict()
        self.contigs: "OrderedDict[str, int]" = OrderedDict()
        self.samples: list[str] = []
        for ln in self.lines:
            m = _META_RE.match(ln)
            if m:
                kind, d = m.group(1), _split_meta(m.group(2))
                if kind == "INFO":
                    self.info[d["ID"]] = (d.get("Number", "."), d.get("Type", "String"))
                elif kind == "FORMAT":
                    self.formats[d["ID"]] = (d.get("Number", "."), d.get("Type", "String"))
                elif kind == "FILTER":
                    self.filters[d["ID"]] = d.get("Description", "")
                else:
                    self.contigs[d["ID"]] = int(d.get("length", 0) or 0)
            elif ln.startswith("#CHROM"):
                self.samples = ln.split("\t")[9:]
        # htslib always knows PASS
        if "PASS" not in self.filters:
            self.filters["PASS"] = "All filters passed"


def _f32(text: str) -> float:
    return float(np.float32(float(text)))


def _typed_elem(text: str, vtype: str):
    if vtype == "Integer":
        return None if text == "." else int(text)
    if vtype == "Float":
        return None if text == "." else _f32(text)
    return text  # String / Character keep the literal (a lone "." stays ".")


def typed_value(text: str | None, number: str, vtype: str):
    """One INFO/FORMAT value -> Python object, by header Number/Type."""
    if vtype == "Flag":
        return True
    scalar = number == "1"
    if text is None or text == "":
        return None if scalar else ()
    if scalar:
        if vtype == "String":
            return text
        return _typed_elem(text, vtype)
    return tuple(_typed_elem(t, vtype) for t in text.split(","))


def parse_gt(text: str):
    if text is None:
        return (None,)
    out = []
    for tok in re.split(r"[/|]", text):
        out.append(None if tok in (".", "") else int(tok))
    return tuple(out)


class OracleRecord:
    """One data line with ``pysam.VariantRecord``-like attributes."""

    __slots__ = ("line", "chrom", "pos", "id", "ref", "alts", "alleles", "qual", "filter_keys",
                 "info", "format_keys", "sample")

    def __init__(self, line: str, header: OracleHeader):
        self.line = line
        cols = line.rstrip("\n").split("\t")
        self.chrom = cols[0]
        self.pos = int(cols[1])
        self.id = None if cols[2] == "." else cols[2]
        self.ref = cols[3]
        self.alts = None if cols[4] == "." else tuple(cols[4].split(","))
        self.alleles = (self.ref,) + (self.alts or ())
        self.qual = None if cols[5] == "." else _f32(cols[5])
        self.filter_keys = [] if cols[6] == "." else cols[6].split(";")
        self.info: "OrderedDict[str, object]" = OrderedDict()
        if cols[7] != ".":
            for kv in cols[7].split(";"):
                if not kv:
                    continue
                key, sep, val = kv.partition("=")
                number, vtype = header.info.get(key, (".", "String"))
                self.info[key] = typed_value(val if sep else None, number, vtype)
        self.format_keys = cols[8].split(":") if len(cols) > 8 and cols[8] != "." else []  # noqa: PLR2004
        self.sample: "OrderedDict[str, object]" = OrderedDict()
        if self.format_keys and len(cols) > 9:  # noqa: PLR2004
            vals = cols[9].split(":")
            for i, key in enumerate(self.format_keys):
                text = vals[i] if i < len(vals) else None
                if key == "GT":
                    self.sample[key] = parse_gt(text if text is not None else ".")
                    continue
                number, vtype = header.formats.get(key, (".", "String"))
                if text is None or text == ".":
                    self.sample[key] = None if number == "1" else (None,)
                else:
                    self.sample[key] = typed_value(text, number, vtype)


class OracleVariantFile:
    """Iterate a plain or gzip/bgzip VCF (whole file or one contig)."""

    def __init__(self, path_or_text):
        if isinstance(path_or_text, (bytes, bytearray)):
            text = bytes(path_or_text).decode()
        else:
            with open(path_or_text, "rb") as fh:
                magic = fh.read(2)
            opener = gzip.open if magic == b"\x1f\x8b" else open
            with opener(path_or_text, "rb") as fh:
                text = fh.read().decode()
        lines = text.split("\n")
        if lines and lines[-1] == "":
            lines.pop()
        self.header_lines = [ln for ln in lines if ln.startswith("#")]
        self.data_lines = [ln for ln in lines if ln and not ln.startswith("#")]
        self.header = OracleHeader(self.header_lines)

    def fetch(self, contig: str | None = None):
        for ln in self.data_lines:
            if contig is None or ln.split("\t", 1)[0] == contig:
                yield OracleRecord(ln, self.header)

    def __iter__(self):
        return self.fetch(None)

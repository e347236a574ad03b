"""VCF header model of the product path (host side).

Gives the plan compiler what ``pysam.VariantHeader`` gives the reference loader:
the ``Number``/``Type`` of every INFO and FORMAT tag, the contig order and the
sample names, plus the header edits of
``ugbio_utils/src/filtering/ugbio_filtering/filter_variants_pipeline.py:106-113``.
"""
from __future__ import annotations

import re
from collections import OrderedDict

_META_RE = re.compile(r"^##(INFO|FORMAT|FILTER|contig)=<(.*)>\s*$")

# The loader's tag whitelist (reference vcftools.py:90-189): only these tags (plus
# --custom_annotations) ever become DataFrame columns, lower-cased.
LOADER_COLUMNS = (
    "GT PL DP AD MQ MMQ SOR AF DP_R DP_F AD_R AD_F TLOD VAF STRANDQ FPR GROUP TREE_SCORE VARIANT_TYPE DB "
    "AS_SOR AS_SORP FS VQR_VAL QD hiConfDeNovo loConfDeNovo GQ PGT PID PS AC AN BaseQRankSum ExcessHet "
    "MLEAC MLEAF MQRankSum ReadPosRankSum XC ID GNOMAD_AF gnomad.AF NLOD NALOD X_IC X_IL X_HIL X_HIN X_LM "
    "X_RM X_GCC X_CSS RPA RU STR AVERAGE_TREE_SCORE VQSLOD BLACKLST SCORE CALL BASE TVAF HighConfidence "
    "BG_AD BG_DP BG_SB DP4 INDEL IDV IMF VDB RPBZ MQBZ BQBZ MQSBZ NM SCBZ SGB MQ0F MinDP ADF ADR GP SYNC "
    "ML_PROB ASSEMBLED_HAPLOTYPES EXOME FILTERED_HAPS HAPCOMP HAPDOM HEC SB MQ0C SCL SCR NMC AFR"
).split()
FIXED_COLUMNS = ("chrom", "pos", "qual", "ref", "alleles", "filter", "id", "indel")


def _split_meta(body: str) -> dict:
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


class VcfHeader:
    def __init__(self, text: str | bytes):
        if isinstance(text, (bytes, bytearray)):
            text = bytes(text).decode()
        self.lines = [ln for ln in text.split("\n") if ln.startswith("#")]
        self.info: "OrderedDict[str, tuple[str, str]]" = OrderedDict()
        self.formats: "OrderedDict[str, tuple[str, str]]" = OrderedDict()
        self.filters: "OrderedDict[str, str]" = OrderedDict()
        self.contigs: "OrderedDict[str, int]" = OrderedDict()
        self.samples: list[str] = []
        for ln in self.lines:
            m = _META_RE.match(ln)
            if m:
                kind, d = m.group(1), _split_meta(m.group(2))
                if "ID" not in d:
                    continue
                if kind == "INFO":
                    self.info[d["ID"]] = (d.get("Number", "."), d.get("Type", "String"))
                elif kind == "FORMAT":
                    self.formats[d["ID"]] = (d.get("Number", "."), d.get("Type", "String"))
                elif kind == "FILTER":
                    self.filters[d["ID"]] = d.get("Description", "")
                else:
                    try:
                        self.contigs[d["ID"]] = int(d.get("length", 0) or 0)
                    except ValueError:
                        self.contigs[d["ID"]] = 0
            elif ln.startswith("#CHROM"):
                self.samples = ln.split("\t")[9:]

    def loader_columns(self, custom_info_fields: list[str] | None = None) -> dict:
        """lower-cased DataFrame column -> VCF tag, as the reference loader selects
        them (vcftools.py:190-207: whitelist + custom fields, intersected with the
        header, case-sensitive)."""
        cols = list(LOADER_COLUMNS)
        for cf in custom_info_fields or []:
            if cf not in cols:
                cols.append(cf)
        known = set(self.info) | set(self.formats)
        return {c.lower(): c for c in cols if c in known}

    def edited_lines(self, *, with_model: bool, with_blacklist: bool) -> list[str]:
        """Header lines of the output file (filter_variants_pipeline.py:106-113)."""
        add = []
        if with_model and "LOW_SCORE" not in self.filters:
            add.append('##FILTER=<ID=LOW_SCORE,Description="Low decision tree score">')
        if with_blacklist and "BLACKLST" not in self.info:
            add.append('##INFO=<ID=BLACKLST,Number=.,Type=String,Description="blacklist">')
        if with_model and "TREE_SCORE" not in self.info:
            add.append('##INFO=<ID=TREE_SCORE,Number=1,Type=Float,Description="Filtering score">')
        lines = list(self.lines)
        if "PASS" not in self.filters:
            # pysam / htslib always emit the PASS filter line (first, after ##fileformat); the records written here do carry PASS
            lines.insert(1 if lines and lines[0].startswith("##fileformat") else 0,
                         '##FILTER=<ID=PASS,Description="All filters passed">')
        at = next(i for i, ln in enumerate(lines) if ln.startswith("#CHROM"))
        return lines[:at] + add + lines[at:]

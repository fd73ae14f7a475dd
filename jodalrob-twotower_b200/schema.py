"""Embedding-table configuration: key order and vocab sizes.

Mirrors what the reference derives at model construction:
  * key list/order: data/column_classifier.py:67-130 (metadata rows in file order, use=Y, not PK,
    text|char(1) with categorical=Y) == the key order of meta/{notice,company}_categorical_config.json;
  * vocab size: src/towers/cat_embed.py:50-85 — int(category count) + 10, or 1000 when the count is
    missing / the column is unknown.
Implemented with the csv module (no pandas dependency on the hot path).
"""
from __future__ import annotations

import csv
import json
import re
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

SAFETY_MARGIN = 10
DEFAULT_VOCAB = 1000

_YES = {"y", "yes", "true", "1", "t"}
_COLS = {
    "table": ["테이블명", "table", "TABLE"],
    "column": ["컬럼명", "컬럼", "column", "COLUMN", "필드명"],
    "use": ["사용 여부", "사용여부", "use", "USE"],
    "pk": ["PK", "pk", "Pk"],
    "dtype": ["타입", "데이터타입", "type", "TYPE", "data_type"],
    "is_cat": ["범주형 여부", "범주형여부", "categorical", "IS_CATEGORICAL"],
    "count": ["범주 갯수"],
}


def _find(header: Sequence[str], names: Sequence[str]) -> Optional[str]:
    for n in names:
        if n in header:
            return n
    norm = {re.sub(r"\s+", "", h).lower(): h for h in header}
    for n in names:
        k = re.sub(r"\s+", "", n).lower()
        if k in norm:
            return norm[k]
    return None


def _read_meta(path: str) -> Tuple[List[dict], Dict[str, Optional[str]]]:
    with open(path, newline="", encoding="utf-8-sig") as f:
        rd = csv.DictReader(f)
        rows = list(rd)
        header = rd.fieldnames or []
    return rows, {k: _find(header, v) for k, v in _COLS.items()}


def categorical_keys_from_metadata(metadata_path: str, table: str) -> List[str]:
    rows, col = _read_meta(metadata_path)
    for need in ("table", "column", "use", "pk", "dtype", "is_cat"):
        if col[need] is None:
            raise KeyError(f"metadata csv lacks a '{need}' column")
    yes = lambda v: str(v if v is not None else "").strip().lower() in _YES
    keys = []
    for r in rows:
        if str(r[col["table"]]).strip() != table or not yes(r[col["use"]]) or yes(r[col["pk"]]):
            continue
        dt = str(r[col["dtype"]]).strip().lower()
        is_text = dt == "text" or dt.startswith("text") or dt == "varchar" or dt.startswith("character varying")
        is_char1 = re.fullmatch(r"(character|char|character varying|varchar)\s*\(\s*1\s*\)", dt) is not None
        if dt in ("bigint", "double precision", "numeric", "integer"):
            continue
        if (is_text or is_char1) and yes(r[col["is_cat"]]):
            keys.append(str(r[col["column"]]).strip())
    return keys


def vocab_sizes_from_metadata(metadata_path: str, table: str, keys: Sequence[str],
                              safety_margin: int = SAFETY_MARGIN) -> Dict[str, int]:
    try:
        rows, col = _read_meta(metadata_path)
        if col["table"] is None or col["column"] is None or col["count"] is None:
            raise KeyError("metadata csv lacks table/column/count columns")
    except Exception:
        # cat_embed.py:82-85: any failure -> default vocab for every key
        return {k: DEFAULT_VOCAB for k in keys}
    out = {}
    for k in keys:
        hit = [r for r in rows if str(r[col["table"]]) == table and str(r[col["column"]]) == k]
        cnt = hit[0][col["count"]] if hit else None
        try:
            out[k] = int(float(cnt)) + safety_margin if cnt not in (None, "") else DEFAULT_VOCAB
        except (TypeError, ValueError):
            out[k] = DEFAULT_VOCAB
    return out


@dataclass
class TowerTables:
    name: str
    keys: List[str]
    vocab: List[int]


def tower_tables(name: str, keys: Sequence[str], metadata_path: Optional[str] = None,
                 vocab: Optional[Dict[str, int]] = None) -> TowerTables:
    if vocab is None:
        vocab = vocab_sizes_from_metadata(metadata_path, name, keys)
    return TowerTables(name, list(keys), [int(vocab[k]) for k in keys])


def load_schema_json(path: str) -> dict:
    """{"notice": [[key, vocab], ...], "company": [...]} (tests/golden/schema_readme_default.json)."""
    with open(path) as f:
        s = json.load(f)
    return {"notice": [(k, int(v)) for k, v in s["notice"]], "company": [(k, int(v)) for k, v in s["company"]]}


def keys_from_categorical_config(path: str) -> List[str]:
    """Key order of meta/{notice,company}_categorical_config.json (insertion order of the JSON object)."""
    with open(path) as f:
        return list(json.load(f).keys())

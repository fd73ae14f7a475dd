"""Shared test utilities (golden loading, schema)."""
import json
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TINY_CASES = ["base", "oob_temp", "deep_ls", "nohidden"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    cfg = json.loads(bytes(d.pop("config_json")).decode())
    return d, cfg


def readme_schema():
    with open(os.path.join(GOLDEN, "schema_readme_default.json")) as f:
        s = json.load(f)
    return {"notice": s["notice"], "company": s["company"]}


def sd_from(d, prefix):
    out = {}
    for k, v in d.items():
        if k.startswith(prefix):
            out[k[len(prefix):]] = torch.from_numpy(np.array(v))
    return out


def batch_inputs(d, i):
    t = lambda n: torch.from_numpy(np.array(d[f"in{i}.{n}"]))
    return t("notice_ids").long(), t("company_ids").long(), t("notice_dense"), t("company_dense")


def rel_err(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))

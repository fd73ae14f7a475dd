"""TEST INFRASTRUCTURE ONLY — loader for the *unmodified* reference implementation.

This module imports the reference's own Python modules from ``/root/reference``
(read-only) so that (a) ``oracle/make_golden.py`` can generate golden vectors and
(b) CPU tests can pin ``oracle/two_tower_oracle.py`` against the real thing.
``/root/reference`` does not exist on the GPU box: nothing that runs there
(``-m gpu`` tests, ``smoke()``, ``bench.py``) may import this file.

Two shims, no reference file is edited:

1. a stub ``torchrec`` package exporting a minimal ``KeyedJaggedTensor``
   (``/root/reference/src/towers/tower/base_tower.py:5`` and
   ``/root/reference/src/towers/cat_embed.py:8`` import it at module top; the
   path only touches ``.values() / .lengths() / .keys() / .to()``,
   ``cat_embed.py:92-94``, ``base_tower.py:130``);
2. ``nn.Module.to`` is wrapped so the hard-coded ``"cuda:0"``
   (``base_tower.py:56,69``, ``cat_embed.py:48``, ``two_tower_model.py:41``)
   lands on the CPU.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types
from typing import List

import torch
import torch.nn as nn

# /root/reference in the build container; on the GPU box the verbatim copy that oracle/make_ref.py leaves in oracle/_ref/
_LOCAL_COPY = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REFERENCE_ROOT = os.environ.get("TT_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference") else _LOCAL_COPY)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "towers", "two_tower_train_task.py"))


class _StubKJT:
    """Minimal stand-in for torchrec.KeyedJaggedTensor (container only)."""

    def __init__(self, keys: List[str], values: torch.Tensor, lengths: torch.Tensor):
        self._keys, self._values, self._lengths = list(keys), values, lengths

    @classmethod
    def from_lengths_sync(cls, keys, values, lengths):
        return cls(keys, values, lengths)

    def keys(self):
        return self._keys

    def values(self):
        return self._values

    def lengths(self):
        return self._lengths

    def to(self, device, non_blocking: bool = False):
        device = _redirect(device)
        return _StubKJT(self._keys, self._values.to(device), self._lengths.to(device))

    def device(self):
        return self._values.device


_ORACLE_DEVICE = torch.device("cpu")


def _redirect(dev):
    if isinstance(dev, str) and dev.startswith("cuda"):
        return _ORACLE_DEVICE
    if isinstance(dev, torch.device) and dev.type == "cuda":
        return _ORACLE_DEVICE
    return dev


_installed = False


def install_shims() -> None:
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    if "torchrec" not in sys.modules:
        pkg = types.ModuleType("torchrec")
        sparse = types.ModuleType("torchrec.sparse")
        jt = types.ModuleType("torchrec.sparse.jagged_tensor")
        pkg.KeyedJaggedTensor = _StubKJT
        jt.KeyedJaggedTensor = _StubKJT
        pkg.sparse = sparse
        sparse.jagged_tensor = jt
        sys.modules["torchrec"] = pkg
        sys.modules["torchrec.sparse"] = sparse
        sys.modules["torchrec.sparse.jagged_tensor"] = jt
    if not torch.cuda.is_available():
        orig_to = nn.Module.to

        def to(self, *args, **kwargs):
            args = tuple(_redirect(a) for a in args)
            if "device" in kwargs:
                kwargs["device"] = _redirect(kwargs["device"])
            return orig_to(self, *args, **kwargs)

        nn.Module.to = to
        orig_tensor_to = torch.Tensor.to

        def tensor_to(self, *args, **kwargs):
            args = tuple(_redirect(a) for a in args)
            if "device" in kwargs:
                kwargs["device"] = _redirect(kwargs["device"])
            return orig_tensor_to(self, *args, **kwargs)

        torch.Tensor.to = tensor_to
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _installed = True


def metadata_path() -> str:
    return os.path.join(REFERENCE_ROOT, "meta", "metadata.csv")


def reference_schema(meta: str | None = None):
    """(notice_keys, company_keys) exactly as scripts/train.py:172,204-205 derives them."""
    install_shims()
    meta = meta or metadata_path()
    from src.torchrec_preprocess.schema import build_torchrec_schema_from_meta

    schema = build_torchrec_schema_from_meta(
        notice_table="notice",
        company_table="company",
        pair_table="bid_two_tower",
        pair_notice_id_cols=["bidntceno", "bidntceord"],
        pair_company_id_cols=["bizno"],
        metadata_path=meta,
    )
    return list(schema.notice.categorical), list(schema.company.categorical)


def build_reference_task(
    categorical_embedding_dim=32,
    notice_dense_input_dim=256,
    company_dense_input_dim=128,
    tower_hidden_dims=(128, 64),
    final_embedding_dim=64,
    dropout_rate=0.0,
    temperature=1.0,
    loss_type="cross_entropy",
    seed=0,
    quiet=True,
    meta: str | None = None,
    label_smoothing: float = 0.0,
):
    """Build the reference TwoTowerTrainTask on CPU (two_tower_train_task.py:211-249)."""
    install_shims()
    from src.towers.two_tower_train_task import create_two_tower_train_task

    meta = meta or metadata_path()
    nk, ck = reference_schema(meta)
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext():
        task = create_two_tower_train_task(
            notice_categorical_keys=nk,
            company_categorical_keys=ck,
            metadata_path=meta,
            categorical_embedding_dim=categorical_embedding_dim,
            notice_dense_input_dim=notice_dense_input_dim,
            company_dense_input_dim=company_dense_input_dim,
            tower_hidden_dims=list(tower_hidden_dims),
            final_embedding_dim=final_embedding_dim,
            dropout_rate=dropout_rate,
            temperature=temperature,
            loss_type=loss_type,
            device="cpu",
        )
    # the one-time pair-alignment print (two_tower_train_task.py:82-84) is a diagnostic
    task._pair_check_done = True
    task.label_smoothing = label_smoothing  # ctor kwarg never forwarded by the factory (:243-247)
    return task, nk, ck


def make_reference_batch(nk, ck, notice_ids, company_ids, notice_dense, company_dense):
    """Batch dict of unified_bid_data_loader.py:630-684 with the sample-major KJT of :827-841."""
    install_shims()

    def kjt(keys, ids):
        b, k = ids.shape
        return _StubKJT(list(keys), ids.reshape(-1).clone(), torch.ones(b * k, dtype=torch.long))

    return {
        "notice": {"dense": notice_dense, "kjt": kjt(nk, notice_ids)},
        "company": {"dense": company_dense, "kjt": kjt(ck, company_ids)},
    }


@contextlib.contextmanager
def silenced():
    with contextlib.redirect_stdout(io.StringIO()):
        yield

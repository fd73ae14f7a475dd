"""Duck-typed KeyedJaggedTensor with exactly the surface the reference touches on the hot path
(/root/reference: src/towers/cat_embed.py:92-94 `.values()/.lengths()/.keys()`,
src/towers/tower/base_tower.py:130 `.to(device)`,
src/towers/pairs/unified_bid_data_loader.py:837 `from_lengths_sync`,
src/training/fast_kjt_builder.py:58-65 ctor + `.pin_memory()`,
src/torchrec_preprocess/torchrec_inputs.py:286 `.device()`).

torchrec is not required; a real torchrec KJT is accepted anywhere this class is (duck typing on
`.values()` / `.lengths()` / `.keys()`).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch


class KeyedJaggedTensor:
    def __init__(self, keys: Sequence[str], values: torch.Tensor, lengths: Optional[torch.Tensor] = None,
                 offsets: Optional[torch.Tensor] = None, layout: str = "sample_major"):
        self._keys = list(keys)
        self._values = values
        self._lengths = lengths
        self._offsets = offsets
        # the reference flattens a [B,K] grid sample-major (unified_bid_data_loader.py:832) while keeping
        # TorchRec's container; canonical TorchRec data is key-major. The flag travels with the tensor.
        self.layout = layout

    @classmethod
    def from_lengths_sync(cls, keys, values, lengths, layout: str = "sample_major"):
        return cls(keys, values, lengths=lengths, layout=layout)

    @classmethod
    def from_offsets_sync(cls, keys, values, offsets, layout: str = "sample_major"):
        return cls(keys, values, offsets=offsets, layout=layout)

    def keys(self) -> List[str]:
        return self._keys

    def values(self) -> torch.Tensor:
        return self._values

    def lengths(self) -> torch.Tensor:
        if self._lengths is None:
            if self._offsets is not None:
                self._lengths = self._offsets[1:] - self._offsets[:-1]
            else:
                self._lengths = torch.ones_like(self._values)
        return self._lengths

    def offsets(self) -> torch.Tensor:
        if self._offsets is None:
            ln = self.lengths()
            off = torch.zeros(ln.numel() + 1, dtype=torch.int64, device=ln.device)
            off[1:] = torch.cumsum(ln, 0)
            self._offsets = off
        return self._offsets

    def lengths_or_none(self):
        return self._lengths

    def offsets_or_none(self):
        return self._offsets

    def device(self) -> torch.device:
        return self._values.device

    def to(self, device, non_blocking: bool = False) -> "KeyedJaggedTensor":
        mv = lambda t: None if t is None else t.to(device, non_blocking=non_blocking)
        return KeyedJaggedTensor(self._keys, mv(self._values), mv(self._lengths), mv(self._offsets), self.layout)

    def pin_memory(self) -> "KeyedJaggedTensor":
        pin = lambda t: None if t is None else t.pin_memory()
        return KeyedJaggedTensor(self._keys, pin(self._values), pin(self._lengths), pin(self._offsets), self.layout)

    def __repr__(self):
        return f"KeyedJaggedTensor(keys={len(self._keys)}, values={tuple(self._values.shape)}, layout={self.layout})"


def is_unit_length(kjt) -> bool:
    """True when every bag holds exactly one id — the reference's only case
    (unified_bid_data_loader.py:835 lengths = ones). Decided without a device sync."""
    n_keys = len(kjt.keys())
    off = kjt.offsets_or_none() if hasattr(kjt, "offsets_or_none") else None
    if off is not None:
        return False
    ln = kjt.lengths_or_none() if hasattr(kjt, "lengths_or_none") else kjt.lengths()
    if ln is None:
        return True
    # ones(B*K) by construction in the reference; a jagged KJT must come with offsets or be flagged
    return ln.numel() == kjt.values().numel() and kjt.values().numel() % max(n_keys, 1) == 0 and \
        not getattr(kjt, "jagged", False)

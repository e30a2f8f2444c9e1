"""Lazy, version-tracked pre-pack of a module's parameters into a device-resident VsgPack."""
from __future__ import annotations

import torch

from ... import _lib


class PackedModuleMixin:
    """Gives an nn.Module a `_pack()` that (re)builds the VsgPack when its parameters change."""

    _vsg_pack = None
    _vsg_key = None

    def _vsg_config(self) -> _lib.VsgConfig:  # pragma: no cover - overridden
        raise NotImplementedError

    def _vsg_prefixes(self):  # (flow_prefix, dec_prefix); None = part absent
        raise NotImplementedError

    def _pack(self) -> _lib.Pack:
        params = list(self.parameters())
        if not params:
            raise RuntimeError("module has no parameters")
        dev = params[0].device
        key = (str(dev),) + tuple((p.data_ptr(), p._version) for p in params)
        if self._vsg_pack is None or self._vsg_key != key:
            sd = {k: v for k, v in self.state_dict().items()}
            fp, dp = self._vsg_prefixes()
            self._vsg_pack = _lib.Pack(self._vsg_config(), sd, fp, dp, dev)
            self._vsg_key = key
        return self._vsg_pack

    def invalidate_pack(self) -> None:
        """Force a re-pack on the next forward (e.g. after editing weights through .data in place)."""
        self._vsg_pack = None
        self._vsg_key = None

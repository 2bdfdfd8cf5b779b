# SPDX-License-Identifier: Apache-2.0
"""Per-geometry caches of search results (warpconvnet/geometry/coords/search/cache.py:24-163)."""
from __future__ import annotations

from typing import Optional, Tuple

from torch import Tensor

from .search_results import IntSearchResult, RealSearchResult


def _tensor_key(t: Tensor) -> Tuple[int, ...]:
    """Hashable copy of a (CPU) offsets tensor; memoised on the tensor object — the same offsets
    tensor is turned into a key several times per conv (the reference walks ``offsets.tolist()``
    on every lookup, cache.py:126-136)."""
    key = getattr(t, "_wcn_key", None)
    if key is None or key[0] != t._version:
        key = (t._version, tuple(int(v) for v in t.detach().cpu().reshape(-1).tolist()))
        try:
            t._wcn_key = key
        except Exception:  # tensor subclasses without a __dict__
            pass
    return key[1]


class IntSearchCacheKey:
    def __init__(self, kernel_size, kernel_dilation, transposed, generative, stride_mode,
                 skip_symmetric_kernel_map, in_offsets, out_offsets):
        self.kernel_size = tuple(kernel_size)
        self.kernel_dilation = tuple(kernel_dilation)
        self.transposed = bool(transposed)
        self.generative = bool(generative)
        self.stride_mode = str(stride_mode)
        self.skip_symmetric_kernel_map = bool(skip_symmetric_kernel_map)
        self.in_offsets = _tensor_key(in_offsets)
        self.out_offsets = _tensor_key(out_offsets)
        self._key = (self.kernel_size, self.kernel_dilation, self.transposed, self.generative,
                     self.stride_mode, self.skip_symmetric_kernel_map, self.in_offsets,
                     self.out_offsets)

    def __hash__(self):
        return hash(self._key)

    def __eq__(self, other):
        return isinstance(other, IntSearchCacheKey) and self._key == other._key

    def __repr__(self):
        return (f"IntSearchCacheKey(kernel_size={self.kernel_size}, "
                f"kernel_dilation={self.kernel_dilation}, transposed={self.transposed}, "
                f"generative={self.generative}, stride_mode={self.stride_mode}, "
                f"num_in={self.in_offsets[-1]}, num_out={self.out_offsets[-1]})")


class IntSearchCache(dict):
    def get(self, key: IntSearchCacheKey) -> Optional[IntSearchResult]:
        return super().get(key, None)

    def put(self, key: IntSearchCacheKey, value: IntSearchResult):
        super().__setitem__(key, value)

    def __repr__(self):
        return f"{self.__class__.__name__}({len(self)} keys)"


class RealSearchCache:
    def __init__(self):
        self._search_cache = {}

    @staticmethod
    def _key(search_args, ref_offsets, query_offsets):
        return (search_args, _tensor_key(ref_offsets), _tensor_key(query_offsets))

    def get(self, search_args, ref_offsets, query_offsets) -> Optional[RealSearchResult]:
        return self._search_cache.get(self._key(search_args, ref_offsets, query_offsets))

    def put(self, search_args, ref_offsets, query_offsets, result: RealSearchResult):
        self._search_cache[self._key(search_args, ref_offsets, query_offsets)] = result

    def __repr__(self):
        return f"Cache({len(self._search_cache)} keys)"

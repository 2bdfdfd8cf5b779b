# SPDX-License-Identifier: Apache-2.0
"""Import the REFERENCE's own explicit gather-matmul-scatter on CPU.  TEST INFRASTRUCTURE ONLY.

Works only where /root/reference is mounted (this container, not the GPU box). It is used by
``tests/golden/make_golden.py`` to generate the committed fixtures and by one ``-m "not gpu"``
test that is skipped when the reference tree is absent. Nothing under ``warpconvnet_b200`` and
nothing that runs on the GPU box may import it.

The reference package does not import without its compiled extension, so two modules are stubbed
before the import: ``warpconvnet._C`` and ``torch_scatter``.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("WCN_REFERENCE_ROOT", "/root/reference")


class _Stub(types.ModuleType):
    """Module whose every attribute is another stub (so ``_C.cuhash.foo`` resolves lazily)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        sub = _Stub(f"{self.__name__}.{name}")
        setattr(self, name, sub)
        return sub

    # Import-time probes of the native extension ("is backend X compiled in?") must answer
    # "no": a call returns a falsy, empty stub. No native compute is ever reached on CPU because
    # only the pure-PyTorch explicit path is used.
    def __call__(self, *a, **k):
        return _Stub(f"{self.__name__}()")

    def __bool__(self):
        return False

    def __int__(self):
        return 0

    def __index__(self):
        return 0

    def __iter__(self):
        return iter(())

    def __len__(self):
        return 0


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "warpconvnet"))


_loaded = None


def load():
    """Returns (explicit_forward, explicit_backward, IntSearchResult) from the reference tree."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference tree not mounted")
    # keep the reference's autotune cache out of $HOME
    os.environ.setdefault("WARPCONVNET_BENCHMARK_CACHE_DIR", "/tmp/wcn_ref_cache")
    for name in ("warpconvnet._C", "torch_scatter"):
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    explicit = importlib.import_module("warpconvnet.nn.functional.sparse_conv.detail.explicit")
    sr = importlib.import_module("warpconvnet.geometry.coords.search.search_results")
    _loaded = (explicit._explicit_gemm_forward_logic, explicit._explicit_gemm_backward_logic,
               sr.IntSearchResult)
    return _loaded

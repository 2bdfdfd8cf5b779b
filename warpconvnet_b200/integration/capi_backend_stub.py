# SPDX-License-Identifier: Apache-2.0
"""INTEGRATION.md seam A written directly against the C-ABI: the file a maintainer of the
REFERENCE would drop into ``warpconvnet/nn/functional/sparse_conv/detail/`` to get the kernels of
``libwcn_b200.so`` as one more backend of its own dispatcher. It is deliberately SELF-CONTAINED:
ctypes + torch only, nothing from the ``warpconvnet_b200`` Python package — every call below is
an entry point of ``include/wcn_b200.h`` with the argument order of that header.

    import capi_backend_stub as wcn
    wcn.register("/path/to/libwcn_b200.so")           # adds "wcn_b200_capi" to the registries
    conv = SparseConv3d(64, 64, 3, fwd_algo=["wcn_b200_capi"], dgrad_algo=["wcn_b200_capi"],
                        wgrad_algo=["wcn_b200_capi"])

Contract of the reference's seam (detail/backends.py:90-131, 443-463, 489-510):
``FORWARD_BACKENDS[name](FwdCtx) -> Tensor | int`` and
``BACKWARD_BACKENDS[name](BwdCtx) -> (Tensor | int | None, Tensor | None)``; a negative int means
"not supported, try the next candidate". Everything derived from a kernel map (pair tables, tile
plans) is cached on the reference's ``IntSearchResult`` object, like its own ``_mask_data``
(detail/mask_gemm.py:127-276). Dense (groups = 1) convolutions in bf16 / fp16; other cases return
the unsupported status. Exercised on the GPU by
``tests/test_gpu_round2.py::test_reference_sparseconv3d_through_the_capi_stub``.
"""
import ctypes
from ctypes import c_float, c_int, c_longlong, c_size_t, c_void_p

import torch

NAME = "wcn_b200_capi"
UNSUPPORTED = -1
_DT = {torch.bfloat16: 0, torch.float16: 1}
_lib = None


def _load(path):
    global _lib
    lib = ctypes.CDLL(path)
    P, I, LL = c_void_p, c_int, c_longlong
    sig = {
        "wcn_kernel_map_num_blocks": (I, [I]),
        "wcn_csr_to_pair_table": (I, [P, P, P, I, I, I, P, P]),
        "wcn_mask_keys": (I, [P, I, I, P, P]),
        "wcn_sort_workspace_bytes": (c_size_t, [I]),
        "wcn_sort_rows_by_key": (I, [P, I, I, P, P, c_size_t, P]),
        "wcn_build_tiles": (I, [P, I, I, P, I, I, P, P, P, P, P, I, P, P]),
        "wcn_weight_image_bytes": (c_size_t, [I, I, I, I, I, I, P, P]),
        "wcn_weight_image": (I, [P, P, I, I, I, I, I, I, P]),
        "wcn_gather_gemm": (I, [P, I, LL, P, P, LL, P, P, P, P, P, I, I, I, I, I, I, I, I, P, I, I,
                                I, P, I, P, P]),
        "wcn_wgrad": (I, [P, LL, P, LL, P, P, P, P, I, I, I, I, I, c_float, I, I, P, I, I, I, I, P,
                          LL, LL, P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _empty(shape, dtype, dev):
    return torch.empty(shape, dtype=dtype, device=dev)


class _Plan:
    """Mask-sorted tiles of one [K, n_rows] neighbour table (wcn_build_tiles)."""

    def __init__(self, table):
        K, M = table.shape
        dev = table.device
        keys = _empty(M, torch.int64, dev)
        assert _lib.wcn_mask_keys(table.data_ptr(), K, M, keys.data_ptr(), _stream()) == 0
        rows_sorted = _empty(M, torch.int32, dev)
        ws_bytes = _lib.wcn_sort_workspace_bytes(M)
        ws = _empty(ws_bytes, torch.uint8, dev)
        assert _lib.wcn_sort_rows_by_key(keys.data_ptr(), M, K, rows_sorted.data_ptr(),
                                         ws.data_ptr(), ws_bytes, _stream()) == 0
        self.tile_rows = 256 if M >= 148 * 256 else 128
        self.m_pad = (M + self.tile_rows - 1) // self.tile_rows * self.tile_rows
        self.num_tiles = self.m_pad // self.tile_rows
        nt = max(self.num_tiles, 1)
        self.step_nbr = _empty((nt, K, self.tile_rows), torch.int32, dev)
        self.step_k = _empty((nt, K), torch.int32, dev)
        self.rows = _empty(max(self.m_pad, 1), torch.int32, dev)
        self.tile_nk = _empty(nt, torch.int32, dev)
        self.tile_cum = _empty(self.num_tiles + 1, torch.int32, dev)
        self.K, self.n_rows = K, M
        assert _lib.wcn_build_tiles(table.data_ptr(), K, M, rows_sorted.data_ptr(), self.tile_rows,
                                    self.m_pad, self.step_nbr.data_ptr(), self.step_k.data_ptr(),
                                    self.rows.data_ptr(), self.tile_nk.data_ptr(),
                                    self.tile_cum.data_ptr(), 0, None, _stream()) == 0


def _state(kernel_map):
    """Per-map cache on the reference's IntSearchResult: device offsets, int32 CSR lists, plans."""
    st = getattr(kernel_map, "_wcn_capi", None)
    if st is None:
        dev = kernel_map.in_maps.device
        st = {"in": kernel_map.in_maps.int().contiguous(), "out": kernel_map.out_maps.int().contiguous(),
              "offs": kernel_map.offsets.to(device=dev, dtype=torch.int32).contiguous()}
        kernel_map._wcn_capi = st
    return st


def _table(st, which, n_rows):
    """[K, n_rows] table: rows = output rows holding input rows ("fwd") or the reverse ("bwd")."""
    key = ("table", which, n_rows)
    if key not in st:
        vals, rows = (st["in"], st["out"]) if which == "fwd" else (st["out"], st["in"])
        K = st["offs"].numel() - 1
        t = _empty((K, n_rows), torch.int32, vals.device)
        assert _lib.wcn_csr_to_pair_table(vals.data_ptr(), rows.data_ptr(), st["offs"].data_ptr(), K,
                                          n_rows, vals.numel(), t.data_ptr(), _stream()) == 0
        st[key] = t
        st[("plan", which, n_rows)] = _Plan(t)
    return st[key], st[("plan", which, n_rows)]


def _image(w, K, cin, cout, transpose):
    code = _DT[w.dtype]
    nbytes = _lib.wcn_weight_image_bytes(K, 1, cin, cout, code, int(transpose), None, None)
    if nbytes == 0:
        return None
    img = _empty(nbytes, torch.uint8, w.device)
    assert _lib.wcn_weight_image(w.data_ptr(), img.data_ptr(), K, 1, cin, cout, code,
                                 int(transpose), _stream()) == 0
    return img


def _gather_gemm(x, img, plan, cin, cout):
    out = _empty((plan.n_rows, cout), x.dtype, x.device)   # every row is written: no zero-fill
    status = _lib.wcn_gather_gemm(
        x.data_ptr(), x.shape[0], x.stride(0), img.data_ptr(), out.data_ptr(), out.stride(0),
        plan.step_nbr.data_ptr(), plan.step_k.data_ptr(), plan.rows.data_ptr(),
        plan.tile_nk.data_ptr(), plan.tile_cum.data_ptr(), plan.num_tiles, plan.tile_rows,
        plan.m_pad, plan.K, 1, cin, cout, _DT[x.dtype], None, 0, 0, 0, None, 0, None, _stream())
    return out if status == 0 else status


def _supported(x, w, groups):
    return (groups == 1 and w.dim() == 3 and x.dtype in _DT and w.shape[1] % 16 == 0
            and w.shape[2] % 16 == 0 and x.is_cuda)


def forward(ctx):
    dt = ctx.compute_dtype or ctx.in_features.dtype
    x = ctx.in_features.to(dt).contiguous()
    w = ctx.weight.to(dt).contiguous()
    if not _supported(x, w, ctx.groups):
        return UNSUPPORTED
    K, cin, cout = w.shape
    _, plan = _table(_state(ctx.kernel_map), "fwd", ctx.num_out_coords)
    img = _image(w, K, cin, cout, transpose=False)
    if img is None:
        return UNSUPPORTED
    y = _gather_gemm(x, img, plan, cin, cout)
    return y if isinstance(y, int) else y.to(ctx.in_features.dtype)


def backward(ctx):
    dt = ctx.compute_dtype or ctx.in_features.dtype
    x = ctx.in_features.to(dt).contiguous()
    gy = ctx.grad_output.to(dt).contiguous()
    w = ctx.weight.to(dt).contiguous()
    if not _supported(x, w, ctx.groups):
        return UNSUPPORTED, None
    K, cin, cout = w.shape
    st = _state(ctx.kernel_map)
    grad_in = grad_w = None
    if ctx.needs_input_grad[0]:                            # dgrad = same kernel on the reverse table
        _, plan = _table(st, "bwd", x.shape[0])
        img_t = _image(w, K, cin, cout, transpose=True)
        if img_t is None:
            return UNSUPPORTED, None
        grad_in = _gather_gemm(gy, img_t, plan, cout, cin)
        if isinstance(grad_in, int):
            return grad_in, None
        grad_in = grad_in.to(ctx.in_features.dtype)
    if len(ctx.needs_input_grad) > 1 and ctx.needs_input_grad[1]:
        dw = torch.zeros((K, cin, cout), dtype=torch.float32, device=x.device)
        status = _lib.wcn_wgrad(x.data_ptr(), x.stride(0), gy.data_ptr(), gy.stride(0), dw.data_ptr(),
                                st["in"].data_ptr(), st["out"].data_ptr(), st["offs"].data_ptr(), K, 1,
                                cin, cout, _DT[x.dtype], c_float(1.0), 0, 0, None, 0, 1, 1, -1, None,
                                x.shape[0], gy.shape[0], _stream())
        if status != 0:
            return status, None
        grad_w = dw.to(ctx.weight.dtype)
    return grad_in, grad_w


def register(lib_path, name=NAME):
    from warpconvnet.nn.functional.sparse_conv.detail import algo_params, backends
    _load(lib_path)
    backends.FORWARD_BACKENDS[name] = forward
    backends.BACKWARD_BACKENDS[name] = backward
    for pool in (algo_params._ALL_AB_PARAMS, algo_params._ALL_ATB_PARAMS):
        if not any(str(tag) == name for tag, _ in pool):
            pool.append((name, {}))
    return name

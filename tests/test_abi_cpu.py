"""CPU-side checks of the C ABI: the shared library loads, exports every symbol the header declares,
and validates its arguments (no kernel is launched, so no GPU is needed)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from learning_embeddings_b200 import _native


def header_symbols():
    src = open(os.path.join(ROOT, "include", "lec_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(lec_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    lib = _native.lib()
    names = header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_native.EXPORTS) == names
    assert lib.lec_abi_version() == _native.ABI_VERSION == 14


def test_argument_validation_codes():
    lib = _native.lib()
    null = ctypes.c_void_p(0)
    fake = ctypes.c_void_p(0x1000)      # 16-byte aligned, never dereferenced: validation fails first
    odd = ctypes.c_void_p(0x1004)
    # NULL rows
    assert lib.lec_pairs_flat(0, 0, null, fake, 10, 4, 4, fake, fake, 8, null, null, 5, 3.0, 1.0, fake, null, null, 1, null) == -1
    # ld not a multiple of 4 / smaller than D
    assert lib.lec_pairs_flat(0, 0, fake, fake, 10, 4, 6, fake, fake, 8, null, null, 5, 3.0, 1.0, fake, null, null, 1, null) == -2
    assert lib.lec_pairs_flat(0, 0, fake, fake, 10, 8, 4, fake, fake, 8, null, null, 5, 3.0, 1.0, fake, null, null, 1, null) == -2
    # unknown geometry / index width
    assert lib.lec_pairs_flat(7, 0, fake, fake, 10, 4, 4, fake, fake, 8, null, null, 5, 3.0, 1.0, fake, null, null, 1, null) == -3
    assert lib.lec_pairs_flat(0, 0, fake, fake, 10, 4, 4, fake, fake, 3, null, null, 5, 3.0, 1.0, fake, null, null, 1, null) == -3
    # negative count, misaligned rows
    assert lib.lec_pairs_flat(0, 0, fake, fake, 10, 4, 4, fake, fake, 8, null, null, -1, 3.0, 1.0, fake, null, null, 1, null) == -4
    assert lib.lec_pairs_flat(0, 0, odd, fake, 10, 4, 4, fake, fake, 8, null, null, 5, 3.0, 1.0, fake, null, null, 1, null) == -5
    # empty batch is a no-op success
    assert lib.lec_pairs_flat(0, 0, fake, fake, 10, 4, 4, null, null, 8, null, null, 0, 3.0, 1.0, null, null, null, 1, null) == 0
    assert lib.lec_pairs_grouped(1, 1, fake, fake, 10, 4, 4, null, null, null, null, 4, 0, 5, null, null, 0.1, 1.0, null,
                                 null, null, null, 1, null) == 0
    # cone energies need the per-row aux terms
    assert lib.lec_pairs_flat(1, 0, fake, null, 10, 4, 4, fake, fake, 8, null, null, 5, 0.1, 1.0, fake, null, null, 1, null) == -1
    assert lib.lec_rows_fwd(fake, 5, 4, 1, 9, 3.0, fake, 4, fake, null, 0, 0, null, null) == -3
    assert lib.lec_rows_fwd(fake, 5, 4, 1, 0, 3.0, fake, 4, fake, fake, 2, 8, null, null) == -4      # replica stride < n * ld
    assert lib.lec_cone_step(None, null) == -1
    # gradient requested with zero replicas
    assert lib.lec_pairs_flat(0, 0, fake, fake, 10, 4, 4, fake, fake, 8, null, null, 5, 3.0, 1.0, fake, null, fake, 0, null) == -7
    assert lib.lec_rows_bwd(fake, fake, 0, 0, 5, 4, 4, 1, 3.0, fake, 0, null) == -7
    assert lib.lec_score_topk(1, 0, fake, 5, fake, 5, 10, 0.1, null, null, 4, 9, null, fake, null, null) == -6
    assert lib.lec_rsgd_update(fake, fake, 1, 5, 0, 0, 0.1, 0.1, 0, null, null) == -2
    # fused update (+ exchange) and the whole-step call (ABI 13)
    u = _native.LecUpdate()
    assert lib.lec_update_rows(None, None, null) == -1
    u.rule, u.row_mode, u.geom = _native.UPD_RSGD, _native.ROWS_HYP_SHELL, 1
    u.table, u.grad_rows, u.grad_replicas, u.n, u.D, u.ld = 0x1000, 0x1000, 1, 5, 4, 4
    u.n = 0
    assert lib.lec_update_rows(ctypes.byref(u), None, null) == 0            # empty table: no launch
    u.n = 5
    u.rule = 7
    assert lib.lec_update_rows(ctypes.byref(u), None, null) == -3           # unknown rule
    u.rule, u.row_mode = _native.UPD_RSGD, _native.ROWS_HYP_TANH_FEAT
    assert lib.lec_update_rows(ctypes.byref(u), None, null) == -3           # FeatNet tail is not a table mode
    u.row_mode, u.grad_replicas = _native.ROWS_HYP_SHELL, 0
    assert lib.lec_update_rows(ctypes.byref(u), None, null) == -7
    u.grad_replicas, u.ld = 1, 6
    assert lib.lec_update_rows(ctypes.byref(u), None, null) == -2           # ld % 4
    u.ld, u.lambda_mode = 4, 2
    assert lib.lec_update_rows(ctypes.byref(u), None, null) == -3
    u.lambda_mode, u.grad_rows = 0, 0x1004
    assert lib.lec_update_rows(ctypes.byref(u), None, null) == -5           # misaligned replicas
    u.grad_rows, u.rule = 0x1000, _native.UPD_ADAM
    assert lib.lec_update_rows(ctypes.byref(u), None, null) == -1           # Adam without its moments
    u.rule, u.momentum = _native.UPD_SGD, 0.9
    assert lib.lec_update_rows(ctypes.byref(u), None, null) == -1           # momentum without its buffer
    u.rule, u.momentum = _native.UPD_RSGD, 0.0
    arr = (ctypes.c_void_p * 2)(0x1000, 0x2000)
    x = _native.LecExchange()
    x.peer_bufs, x.world, x.rank, x.slot, x.tag = ctypes.cast(arr, ctypes.c_void_p), 2, 5, 0, 1
    x.slot_packets = lib.lec_exchange_packets(5, 4)
    assert x.slot_packets == 11
    assert lib.lec_update_rows(ctypes.byref(u), ctypes.byref(x), null) == -8   # rank out of range
    x.rank, x.slot = 0, 3
    assert lib.lec_update_rows(ctypes.byref(u), ctypes.byref(x), null) == -8   # slot
    x.slot, x.tag = 0, 0
    assert lib.lec_update_rows(ctypes.byref(u), ctypes.byref(x), null) == -8   # tag 0 is the empty-buffer value
    x.tag, x.slot_packets = 1, 10
    assert lib.lec_update_rows(ctypes.byref(u), ctypes.byref(x), null) == -8   # source region too small
    x.slot_packets, x.peer_bufs = 11, None
    assert lib.lec_update_rows(ctypes.byref(u), ctypes.byref(x), null) == -1
    # FeatNet kernels
    assert lib.lec_featnet_supported(2048, 10) == 1 and lib.lec_featnet_supported(2048, 50) == 0
    assert lib.lec_featnet_supported(2050, 10) == 0
    assert lib.lec_featnet_fwd(fake, 100, 2048, fake, 8, 10, fake, null, 50, fake, null) == -2      # D > 16
    assert lib.lec_featnet_fwd(fake, 100, 2048, fake, 2, 10, fake, null, 10, fake, null) == -3      # index width
    assert lib.lec_featnet_fwd(null, 100, 2048, fake, 8, 10, fake, null, 10, fake, null) == -1
    assert lib.lec_featnet_fwd(fake, 100, 2048, null, 8, 200, fake, null, 10, fake, null) == -4     # identity selection past the pool
    assert lib.lec_featnet_fwd(fake, 100, 2048, fake, 8, 0, null, null, 10, null, null) == 0        # empty step
    assert lib.lec_featnet_wgrad(fake, 100, 2048, fake, 8, 10, fake, 10, fake, 0, 0, null) == -7
    assert lib.lec_featnet_wgrad(fake, 100, 2048, fake, 8, 10, fake, 10, fake, 4, 1000, null) == -4  # replica stride too small
    assert lib.lec_set_pdl(0) in (0, 1)
    step = _native.LecStep()
    assert lib.lec_cone_step(ctypes.byref(step), null) == -1    # no table
    step.upd = u
    assert lib.lec_cone_step(ctypes.byref(step), null) == -1    # rows_out / loss_acc missing
    assert lib.lec_index_errors(None, 0, null) != -1            # count_out is optional (no device here: a CUDA error code)
    assert lib.lec_caption_hinge(null, fake, 4, 3, 1.0, null, fake, null, null, null) == -1
    assert lib.lec_caption_hinge(fake, fake, -1, 3, 1.0, null, fake, null, null, null) == -4
    assert lib.lec_caption_hinge(null, null, 0, 3, 1.0, null, null, null, null, null) == 0
    assert b"16-byte" in lib.lec_error_string(-5)
    # host pipe (ABI 14): argument errors are reported before any CUDA call
    h = ctypes.c_void_p()
    assert lib.lec_host_pipe_create(None, 2) == -1
    assert lib.lec_host_pipe_create(ctypes.byref(h), 0) == -4 and lib.lec_host_pipe_create(ctypes.byref(h), 17) == -4
    assert lib.lec_host_pipe_submit(None, 0, ctypes.byref(step), None, fake, 8, fake, fake, null, null) == -1
    assert lib.lec_host_pipe_wait(None, 0) == -1
    lib.lec_host_pipe_destroy(None)


def test_ops_refuse_cpu_tensors():
    import torch
    from learning_embeddings_b200 import ops
    with pytest.raises(_native.LecError):
        ops.energy(torch.zeros(4, 3), torch.zeros(4, 3), "euc", 3.0)
    with pytest.raises(_native.LecError):
        ops.rows_forward(torch.zeros(4, 3), _native.ROWS_EUC_SOFTCLIP, 3.0, 'euc')

"""CPU-only: the C-ABI library loads, exports every symbol include/pacoh_b200.h declares, and its host-only
entry points (layout / prior / workspace sizing / argument validation) behave.  No kernel is launched here."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from meta_learning_pacoh_b200 import _lib, engine as eng
from oracle import pacoh_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "pacoh_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pacoh_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported():
    names = _header_functions()
    assert sorted(_lib.EXPORTED_SYMBOLS) == names
    for n in names:
        assert hasattr(_lib.lib, n), n
    assert _lib.lib.pacoh_abi_version() == 1


def test_param_count_and_layout_match_reference(golden_dir):
    import json
    ref = json.load(open(os.path.join(golden_dir, "layout.json")))
    cases = {"default_d1": eng.GPArch(1), "arch_d2": eng.GPArch(2, mean_layers=(16,), kernel_layers=(8, 24, 16)),
             "const_se_d2": eng.GPArch(2, mean_kind="constant", covar_kind="SE")}
    for key, arch in cases.items():
        ent = arch.entries()
        assert list(ent.keys()) == list(ref[key].keys())
        assert [b - a for a, b in ent.values()] == [v[0] for v in ref[key].values()]
    assert eng.GPArch(1).D == 2342
    assert eng.GPArch(1, mean_layers=(32,) * 4, kernel_layers=(32,) * 4).D == 6566
    assert eng.GPArch(1, outputscale=True, noise_floor=1e-3).D == 2343


@pytest.mark.parametrize("kw", [dict(input_dim=1), dict(input_dim=3, mean_layers=(8, 8, 8), kernel_layers=(64,), feature_dim=5),
                                dict(input_dim=2, mean_kind="constant", covar_kind="SE"),
                                dict(input_dim=2, mean_kind="zero", covar_kind="NN")])
def test_hyper_prior_params_match_oracle(kw):
    arch = eng.GPArch(**kw)
    okw = dict(kw)
    lay = orc.Layout(**okw)
    assert lay.D == arch.D
    mu, sigma = arch.hyper_prior(0.4, 2.5)
    mu2, sigma2 = orc.hyper_prior_params(lay, 0.4, 2.5)
    assert torch.equal(mu, mu2) and torch.equal(sigma, sigma2)


def test_pre_factor_matches_reference_formula():
    assert abs(eng.pre_factor([5] * 20) - 5.0 / 25.0) < 1e-12
    assert abs(eng.pre_factor([4, 12]) - orc.pre_factor([4, 12])) < 1e-15


def test_argument_validation_and_workspace_sizes():
    a = eng.GPArch(1).c_struct()
    nbytes = _lib.lib.pacoh_workspace_bytes(ctypes.byref(a), 64, 4096, 50)
    assert nbytes > 64 * 4096 * 50 * 4 * 6          # m, z(2), dm, dz(2)
    assert nbytes < 2 * 1024 ** 3
    assert _lib.lib.pacoh_workspace_bytes(ctypes.byref(a), 0, 4, 5) == _lib.PACOH_ERR_INVALID
    assert _lib.lib.pacoh_workspace_bytes(ctypes.byref(a), 2, 4, 128) > 0              # extended range: one matrix per CTA
    assert _lib.lib.pacoh_workspace_bytes(ctypes.byref(a), 2, 4, 129) > 2 * 4 * 256 * 256 * 4    # blocked large-n path: L / U tiles
    assert _lib.lib.pacoh_workspace_bytes(ctypes.byref(a), 1, 8, 2048) > 8 * 2048 * 2048 * 4      # BASELINE config #5
    assert _lib.lib.pacoh_workspace_bytes(ctypes.byref(a), 2, 4, 4097) == _lib.PACOH_ERR_UNSUPPORTED
    assert b"not supported" in _lib.lib.pacoh_last_error()
    bad = eng.GPArch(1).c_struct()
    bad.mean_kind = 7
    assert _lib.lib.pacoh_param_count(ctypes.byref(bad)) == _lib.PACOH_ERR_INVALID
    with pytest.raises(_lib.PacohError):
        _lib.check(_lib.lib.pacoh_svgd_phi(4, 10, None, None, 0.0, 0, None, None, None, 0, None))
    assert _lib.lib.pacoh_svgd_workspace_bytes(64, 2342) >= 4 * (74 * 64 * 64 + 64 * 64 + 64)


def test_engine_refuses_cpu_device():
    with pytest.raises(RuntimeError):
        eng.MetaMLLEngine(eng.GPArch(1), np.zeros((2, 5, 1), np.float32), np.zeros((2, 5), np.float32), device="cpu")


def test_round2_entry_points_validate_arguments_and_size_workspaces():
    """Host-side behaviour of the entry points added in round 2 (no device work): large-n layout, posterior workspace, step state."""
    lib = _lib.lib
    a = eng.GPArch(1, outputscale=True, noise_floor=1e-3).c_struct()
    out = (ctypes.c_int64 * 14)()
    # BASELINE config #5: 1024 tasks x 2048 points, P = 1 -> 16 tiles per side, one pass, factors = n^2 * 4 B per matrix
    assert lib.pacoh_debug_big_layout(ctypes.byref(a), 1, 1024, 2048, out) == 0
    off_big, nb, npad, batch, off_L, off_SB = out[0], out[1], out[2], out[3], out[4], out[5]
    assert (nb, npad, batch) == (16, 2048, 1024) and off_L == 0 and off_SB >= 1024 * 2048 * 2048 * 4
    total = lib.pacoh_workspace_bytes(ctypes.byref(a), 1, 1024, 2048)
    assert off_big + out[13] <= total < 24 * 1024 ** 3
    # n = 130: two tiles, padded to 256; n = 50: a small-matrix kernel, no large-n scratch at all
    assert lib.pacoh_debug_big_layout(ctypes.byref(a), 2, 3, 130, out) == 0 and (out[1], out[2], out[3]) == (2, 256, 6)
    assert lib.pacoh_debug_big_layout(ctypes.byref(a), 2, 3, 50, out) == 0 and list(out) == [0] * 14
    # a batch that does not fit the 24 GB scratch budget is processed in passes
    assert lib.pacoh_debug_big_layout(ctypes.byref(a), 4, 1024, 2048, out) == 0 and 1 <= out[3] < 4096
    assert lib.pacoh_workspace_bytes(ctypes.byref(a), 4, 1024, 2048) < 26 * 1024 ** 3

    # predictive path: context sets up to 128 points, n_c + n* up to 4096, the covariance scratch only on request
    w0 = lib.pacoh_gp_posterior_workspace_bytes(ctypes.byref(a), 10, 20, 5, 50, 0)
    w1 = lib.pacoh_gp_posterior_workspace_bytes(ctypes.byref(a), 10, 20, 5, 50, 1)
    assert 0 < w0 < w1 and w1 - w0 >= 10 * 20 * 5 * 50 * 4
    assert lib.pacoh_gp_posterior_workspace_bytes(ctypes.byref(a), 2, 3, 50, 500, 0) > 2 * 3 * 640 * 640 * 4      # joint sets > 64: Cholesky tiles
    assert lib.pacoh_gp_posterior_workspace_bytes(ctypes.byref(a), 2, 3, 129, 10, 0) == _lib.PACOH_ERR_UNSUPPORTED
    assert lib.pacoh_gp_posterior_workspace_bytes(ctypes.byref(a), 2, 3, 100, 4000, 0) == _lib.PACOH_ERR_UNSUPPORTED
    assert lib.pacoh_gp_posterior_workspace_bytes(ctypes.byref(a), 0, 3, 5, 10, 0) == _lib.PACOH_ERR_INVALID
    assert lib.pacoh_gp_posterior(ctypes.byref(a), 2, 3, 5, 10, None, None, None, None, None, None, None, None, None, None, None, None,
                                  None, 0, None) == _lib.PACOH_ERR_INVALID
    assert lib.pacoh_pred_metrics(2, 3, 10, None, None, None, None, None, 1.0, None, None) == _lib.PACOH_ERR_INVALID

    # step state / fused optimizers: null pointers are refused before any launch
    assert lib.pacoh_step_prepare(None, 1, 4, None, None, 0, None, None, 1e-3, 1.0, 0, 0.9, 0.999, None) == _lib.PACOH_ERR_INVALID
    assert lib.pacoh_adam_step_dev(10, None, None, 1.0, None, None, 0.9, 0.999, 1e-8, None, None) == _lib.PACOH_ERR_INVALID
    assert lib.pacoh_adamw_step_dev(10, None, None, 1.0, None, None, 0.9, 0.999, 1e-8, 0.1, None, None, None) == _lib.PACOH_ERR_INVALID
    assert lib.pacoh_peer_allreduce_finalize_dev(9, 0, None, None, 0, None, None, 4, 10, None, None, None, 0.01, 0.5, None, None, None) == _lib.PACOH_ERR_INVALID


@pytest.mark.parametrize("n", [65, 128, 129, 300, 512, 1000, 1024, 2047, 2048, 3000, 4096])
def test_large_n_workspace_layout_invariants(n):
    """Property sweep over the large-n scratch layout (host arithmetic only): tiles cover n, the sub-buffers are 256-byte
    aligned, disjoint and in order, the per-pass batch respects the 24 GiB budget and never exceeds the matrices asked for,
    and the caller-visible workspace size covers the layout for every batch size."""
    lib = _lib.lib
    a = eng.GPArch(1, outputscale=True, noise_floor=1e-3).c_struct()
    out = (ctypes.c_int64 * 14)()
    prev_total = 0
    for P, T in ((1, 1), (1, 7), (3, 50), (1, 1024), (8, 4096)):
        assert lib.pacoh_debug_big_layout(ctypes.byref(a), P, T, n, out) == 0
        off_big, nb, npad, batch = out[0], out[1], out[2], out[3]
        offs, total = list(out[4:13]), out[13]
        assert nb == -(-n // 128) and npad == 128 * nb and npad >= n
        assert 1 <= batch <= P * T
        assert offs[0] == 0 and all(o % 256 == 0 for o in offs) and offs == sorted(offs) and len(set(offs)) == len(offs)
        assert offs[1] - offs[0] >= batch * npad * npad * 4                    # factor tiles of every matrix of a pass
        assert offs[2] - offs[1] >= batch * nb * 2 * 128 * 128 * 4             # inverse diagonal tiles (L_kk^-1 and U_kk)
        assert offs[-1] < total <= (24 << 30) + (64 << 20) or batch == 1       # budget (one matrix always goes through)
        if batch < P * T:                                                      # passes: the budget is what limits the batch
            assert (batch + 1) * (npad * npad * 4 + nb * 2 * 128 * 128 * 4) > (24 << 30) * 0.9
        ws = lib.pacoh_workspace_bytes(ctypes.byref(a), P, T, n)
        assert ws >= off_big + total and off_big % 256 == 0
        assert total >= prev_total                                             # monotone in the number of matrices
        prev_total = total

"""CPU: the C-ABI library loads and exports every symbol include/its_b200.h declares."""
import ctypes
import os
import re

from tests.conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "its_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(its_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    syms = declared_symbols()
    assert len(syms) >= 18
    missing = [s for s in syms if not hasattr(built_lib, s)]
    assert not missing, missing


def test_binding_lists_every_symbol(built_lib):
    from its_b200 import _lib
    assert sorted(_lib.SYMBOLS) == declared_symbols()


def test_version_and_struct_layout(built_lib):
    from its_b200 import _lib
    assert built_lib.its_version() >= 100
    assert built_lib.its_abi_sizeof(0) == ctypes.sizeof(_lib.ConvDesc)
    assert built_lib.its_abi_sizeof(1) == ctypes.sizeof(_lib.Src)
    assert built_lib.its_abi_sizeof(2) == ctypes.sizeof(_lib.Phase)


def test_invalid_arguments_return_error_codes(built_lib):
    """Validation happens before any CUDA call, so it is testable without a GPU."""
    rc = built_lib.its_ddpm_step(None, None, None, None, 0, 1, 4, None, None, 0.0, 0, 0, None, 0, None)
    assert rc == 1
    assert b"null pointer" in built_lib.its_last_error_string()
    rc = built_lib.its_group_norm(1, 1, 12, None, 0, 1, 1, 1, 16, 32, 1e-5, 1, 1, 1, 0, None)
    assert rc == 1 and b"multiples of 8" in built_lib.its_last_error_string()
    from its_b200._lib import ConvDesc
    d = ConvDesc()
    rc = built_lib.its_conv_igemm(ctypes.byref(d), 0, None)
    assert rc == 1


def test_product_fails_loudly_without_cuda(built_lib):
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from its_b200.Diffusion import UNet, GaussianDiffusionSampler
    net = UNet(T=10, ch=64, ch_mult=[1, 2], attn=[1], num_res_blocks=1, dropout=0.0)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(1, 3, 16, 16), torch.zeros(1, dtype=torch.long))
    with pytest.raises(RuntimeError, match="CUDA"):
        GaussianDiffusionSampler(net, 1e-4, 0.02, 10)(torch.zeros(1, 3, 16, 16))
    from its_b200.search.verifier import OracleVerifier
    with pytest.raises(RuntimeError, match="CUDA"):
        OracleVerifier().score(torch.zeros(2, 3, 8, 8))


def test_no_product_import_of_oracle():
    """The oracle is test infrastructure: nothing under the package may import it."""
    pkg = os.path.join(ROOT, "inference-time-scaling-for-diffusion-models-beyond-scaling-denoising-steps_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                # the reference's verifier is called OracleVerifier ("oracle" as a config value); what must
                # not appear is an import of, or a path into, the repo's oracle/ directory
                src = re.sub(r"OracleVerifier|KIND_ORACLE|oracle_|Oracle|[\"']oracle[\"']|\boracle \|", "", src)
                assert not re.search(r"(from|import)\s+\.*oracle\b|oracle[/\\.]|\boracle\b", src), (dirpath, f)

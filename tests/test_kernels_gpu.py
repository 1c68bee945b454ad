"""Per-kernel GPU parity through the C ABI (sam3b_gemm / sam3b_attention_* / sam3b_layernorm_* / ...):
each case compares one hand-written sm_100a kernel with a plain PyTorch fp32 computation of the same
op on the same seeded inputs (TF32 disabled).  Tolerances: fp32 outputs 1e-5 relative to the output
range; 16-bit outputs 1.5e-3 (fp16) / 1.2e-2 (bf16) = a few ulps of the storage format."""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tools"))

pytestmark = pytest.mark.gpu

F32_TOL, F16_TOL, BF16_TOL = 1e-5, 1.5e-3, 1.2e-2


@pytest.mark.parametrize("case", ["k_tile1", "k_tile1_bn64", "k_multi", "k_ragged", "k_persist", "k_big_f16", "k_big_bf16",
                                  "epi_residual", "p_tile1", "p_multi", "p_ragged", "p_persist", "p_big_f16", "p_big_bf16"])
def test_gemm_fp32_outputs(case):
    import bringup_gemm

    r = bringup_gemm.run_case(case)
    assert r["err"] < F32_TOL, r


@pytest.mark.parametrize("case", ["k_store16", "k_skinny", "k_skinny_full", "k_skinny_ragged", "epi_gelu", "epi_dgelu", "epi_rope", "p_store16"])
def test_gemm_16bit_epilogues(case):
    import bringup_gemm

    r = bringup_gemm.run_case(case)
    assert r["err"] < F16_TOL, r
    if "pad_untouched" in r:
        assert r["pad_untouched"]       # the K-extension columns next to the output are not clobbered


def test_gemm_mn_major_splitk_atomics():
    import bringup_gemm

    r = bringup_gemm.run_case("mn_8192_1024")
    assert r["err"] < F32_TOL and r["err_trans"] < F32_TOL, r


@pytest.mark.parametrize("case", ["attn_bwd_576_2_2", "attn_bwd_64_3_2", "attn_fwd_192_1_1", "attn_fwd_1152_1_2", "attn_bwd_5184_1_16"])
def test_attention_fp16(case):
    import bringup_ops

    r = bringup_ops.run_case(case)
    assert r["err_O"] < F16_TOL and r["err_lse"] < 1e-4 and r["pad_untouched"], r
    for k in ("err_dq", "err_dk", "err_dv"):
        if k in r:
            assert r[k] < 2e-3, r
    if "err_delta" in r:
        assert r["err_delta"] < 1e-5, r


def test_attention_bf16():
    import bringup_ops

    r = bringup_ops.run_case("attn_bwd_576_2_2_bf16")
    assert r["err_O"] < BF16_TOL and max(r["err_dq"], r["err_dk"], r["err_dv"]) < BF16_TOL, r


@pytest.mark.parametrize("case", ["ln_128", "ln_1024"])
def test_layernorm(case):
    import bringup_ops

    r = bringup_ops.run_case(case)
    assert r["err_fwd"] < F16_TOL and r["err_bwd16"] < F16_TOL and r["err_bwd"] < F32_TOL and r["err_mean"] < F32_TOL, r
    assert r["err_cast"] == 0.0


def test_patch_gather_and_layout_roundtrip():
    import bringup_ops

    r = bringup_ops.run_case("patch")
    assert r["err"] == 0.0 and r["pad_zero"] and r["err_to_nchw"] == 0.0 and r["err_roundtrip"] == 0.0, r


def test_lora_pack_unpack_bit_exact():
    import bringup_ops

    r = bringup_ops.run_case("lora_pack")
    assert r["ok"] and r["ok_unpack"], r


def test_adamw_matches_torch():
    import bringup_ops

    r = bringup_ops.run_case("adamw")
    assert r["err"] < 1e-6, r

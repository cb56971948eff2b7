"""CPU: the C-ABI library loads and exports every symbol include/slnlp_b200.h declares
(no compute calls - there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "slnlp_b200.h")


def declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"(?:int64_t|int|const char\*)\s+(slnlp_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        out[m.group(1)] = n
    return out


def test_header_declares_the_hot_path():
    d = declared()
    for name in ("slnlp_embed_gather_fwd", "slnlp_gemm_f32", "slnlp_rnn_layer_fwd", "slnlp_rnn_layer_bwd",
                 "slnlp_attn_step_fwd", "slnlp_attn_step_bwd", "slnlp_ce_on_logp", "slnlp_gradnorm",
                 "slnlp_sgd_momentum_clip", "slnlp_last_error_string"):
        assert name in d


def test_library_exports_every_declared_symbol():
    from slnlp_b200 import _lib
    d = declared()
    assert len(d) >= 25
    for name, nargs in d.items():
        fn = getattr(_lib.lib, name)            # AttributeError = missing export
        assert name in _lib.SIGNATURES, name
        assert len(_lib.SIGNATURES[name]) == nargs, (name, nargs, len(_lib.SIGNATURES[name]))
        assert fn is not None
    assert _lib.lib.slnlp_abi_version() == 1
    assert isinstance(_lib.last_error(), str)


def test_argument_errors_are_reported_without_a_gpu():
    from slnlp_b200 import _lib
    rc = _lib.lib.slnlp_gemm_f32(0, 0, 4, 4, 4, None, 4, None, 4, None, 4, None, 0.0, None, 0, None)
    assert rc != 0 and "null" in _lib.last_error()
    rc = _lib.lib.slnlp_rnn_layer_fwd(7, 0, 1, 1, 1, 1, None, None, None, None, None, None, None, None, None, None)
    assert rc != 0 and "mode" in _lib.last_error()


def test_fused_decoder_entry_points_validate_their_arguments_without_a_gpu():
    from slnlp_b200 import _lib
    L = _lib.lib
    assert L.slnlp_dec_cell_bwd_supported(0, 50, 128, 384) == 1 and L.slnlp_dec_cell_bwd_supported(0, 50, 20, 64) == 0
    assert L.slnlp_dec_head_supported(64, 50, 128, 2, 0, 0) == 1
    assert L.slnlp_dec_head_supported(64, 50, 128, 2, 0, 1) == 0          # a fused bridge needs the fused query
    assert L.slnlp_dec_head_supported(64, 50, 512, 6, 1, 1) == 1 and L.slnlp_dec_head_supported(20000, 50, 128, 2, 0, 0) == 0
    one = 8      # any non-null address: the checks below fail before anything is dereferenced or launched
    rc = L.slnlp_dec_cell_bwd(0, 50, 20, 64, one, one, one, one, one, one, one, one, None, one, one, 0.0, None, 0, None)
    assert rc != 0 and "multiples of 16" in _lib.last_error()
    rc = L.slnlp_dec_cell_bwd(1, 50, 32, 64, one, one, one, None, one, one, one, one, None, one, one, 0.0, None, 0, None)
    assert rc != 0 and "GRU" in _lib.last_error()
    rc = L.slnlp_dec_head_fwd(64, 50, 128, 2, 128, None, None, None, None, None, None, None, None, 1, None, None, None, None, None,
                              None, None)
    assert rc != 0 and "null" in _lib.last_error()
    rc = L.slnlp_dec_head_bwd(64, 50, 128, 2, 128, one, one, one, one, one, one, None, None, one, one, one, one, one, None, None, None)
    assert rc != 0 and "fused bridge" in _lib.last_error()
    rc = L.slnlp_relu_dropout_fwd(one, 16, 1.5, one, 0, None)
    assert rc != 0 and "relu_dropout_fwd" in _lib.last_error()


def test_library_is_cuda_only_sm100a():
    """The .so carries sm_100a SASS (and nothing for another arch)."""
    import subprocess
    from slnlp_b200 import _lib
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        import pytest
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_(\d+a?)", out.stdout))
    assert archs == {"100a"}, archs


def test_hot_kernels_use_the_blackwell_paths_in_sass():
    """The built library's SASS shows the hardware path each hot kernel claims (B200_PROFILING.md:
    tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG, cluster barrier -> UCGABAR,
    warp-level MMA -> HMMA): a guard against a silent regression to CUDA-core code."""
    import subprocess
    import sys
    script = os.path.join(ROOT, "profiles", "sass_mnemonics.py")
    out = subprocess.run([sys.executable, script], capture_output=True, text=True)
    if out.returncode != 0 or not out.stdout.strip():
        import pytest
        pytest.skip("cuobjdump unavailable")
    rows = {}
    for line in out.stdout.splitlines()[1:]:
        name, _, rest = line.partition("  ")
        rows[name.strip()] = dict(kv.split("=") for kv in rest.split())
    def need(prefix, *ops):
        hits = [v for k, v in rows.items() if k.startswith(prefix)]
        assert hits, prefix
        for v in hits:
            for op in ops:
                assert int(v.get(op, 0)) > 0, (prefix, op, v)
    need("rnn_persistent_fwd_kernel", "UTCHMMA", "LDTM", "STTM")      # W_hh resident in TMEM, A operand from TMEM
    need("rnn_persistent_bwd_kernel", "UTCHMMA", "LDTM", "STTM")
    need("rnn_cluster_fwd_kernel", "UTCHMMA", "LDTM", "UCGABAR")      # thread-block cluster + DSMEM exchange
    need("rnn_cluster_bwd_kernel", "UTCHMMA", "LDTM", "UCGABAR")
    need("rnn_step_fwd_tc_kernel", "UTCHMMA", "UTMALDG", "LDTM")      # TMA-fed tf32 step kernels
    need("rnn_step_bwd_tc_kernel", "UTCHMMA", "UTMALDG", "LDTM")
    need("gemm_tma_kernel", "UTCHMMA", "UTMALDG", "LDTM")             # TMA-fed tcgen05 GEMM
    need("gemm_pair_kernel", "UTCHMMA.2CTA", "UTMALDG", "LDTM", "UCGABAR")            # CTA-pair (cta_group::2) bf16 GEMM
    need("lstm_step_fwd_pair_kernel", "UTCHMMA.2CTA", "UTMALDG", "LDTM", "UCGABAR")   # CTA-pair recurrent step kernels
    need("lstm_step_bwd_pair_kernel", "UTCHMMA.2CTA", "UTMALDG", "LDTM", "UCGABAR")
    need("mha_tc_fwd_kernel", "HMMA")                                 # warp-level tf32 MMA attention
    need("mha_tc_bwd_kernel", "HMMA")

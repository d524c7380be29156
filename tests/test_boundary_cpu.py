"""CPU-side tests of the drop-in boundary: the C-ABI library loads and exports every
symbol include/immtsf.h declares (with matching arity), the `fusions` mirror keeps the
reference's registries / constructor contracts / state_dict, and nothing computes on CPU."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from helpers import golden_names, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "immtsf.h")
C2CT = {"int": "c_int", "float": "c_float", "uint32_t": "c_uint", "uint64_t": "c_ulong", "size_t": "c_ulong", "long": "c_long"}


def _header_decls():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|const char\*|unsigned long long|size_t)\s+(immtsf_\w+)\s*\(([^)]*)\)\s*;", src):
        name, args = m.group(1), m.group(2).strip()
        decls[name] = [] if args in ("", "void") else [a.strip() for a in args.split(",")]
    return decls


def test_header_declares_something():
    d = _header_decls()
    assert len(d) >= 25 and "immtsf_gemm" in d and "immtsf_csr_build" in d


def test_library_exports_every_declared_symbol():
    from immtsf import _lib

    lib = _lib.load()
    for name in _header_decls():
        assert hasattr(lib, name), f"{name} declared in include/immtsf.h but not exported"


def test_binding_signatures_match_header():
    from immtsf import _lib

    decls = _header_decls()
    for name, argtypes in _lib.SIGNATURES.items():
        assert name in decls, name
        cargs = decls[name]
        assert len(cargs) == len(argtypes), (name, len(cargs), len(argtypes))
        for carg, ct in zip(cargs, argtypes):
            if "*" in carg:
                assert ct is ctypes.c_void_p, (name, carg)
            else:
                base = carg.split()[-2] if len(carg.split()) >= 2 else carg
                assert ct.__name__ == C2CT[base], (name, carg, ct.__name__)
    assert set(decls) - set(_lib.SIGNATURES) == {"immtsf_last_error_string", "immtsf_launch_count", "immtsf_gemm_workspace_bytes",
                                                   "immtsf_gemm_batched_workspace_bytes", "immtsf_masked_mse_workspace_bytes",
                                                   "immtsf_t2vq_bwd_workspace_bytes", "immtsf_xattn_rank_fused_bwd_workspace_bytes", "immtsf_nvls_flag_bytes"}


def test_library_contains_sm100a_code_only():
    from immtsf import _lib

    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_arg_validation_without_a_gpu():
    """Argument errors are reported before any launch, so they are testable on a CPU box."""
    from immtsf import _lib

    lib = _lib.load()
    rc = lib.immtsf_gemm(0, 1, 4, 4, 4, 1.0, None, 4, None, 4, 0.0, None, 4, None, None, 0, 0, None, 0, None)
    assert rc == -1 and b"null operand" in lib.immtsf_last_error_string()
    rc = lib.immtsf_recavg_pool_fwd(1, 6, 1, 1, 1, 0, 1, 1, 1, 2, 2, 6, 4, 1e-5, 0, 0, 1, None, None, None, None, None)
    assert rc == -1 and b"multiple of 4" in lib.immtsf_last_error_string()


def test_registries_and_identity_dispatch():
    import fusions.FusionModel as FM
    from fusions.TTF_RecAvg import TTF_RecAvg
    from fusions.TTF_T2V_XAttn import TTF_T2V_XAttn
    from fusions.MMF_GR_Add import MMF_GR_Add
    from fusions.MMF_XAttn_Add import MMF_XAttn_Add

    from fusions.TTF_T2V_XAttn_old import TTF_T2V_XAttn as PerQuery

    # the reference's two names (fusions/FusionModel.py:14-17) plus the per-(note, query) variant
    assert FM._TTF_CLASSES == {"TTF_RecAvg": TTF_RecAvg, "TTF_T2V_XAttn": TTF_T2V_XAttn, "TTF_T2V_XAttn_old": PerQuery}
    assert FM._MMF_CLASSES == {"MMF_GR_Add": MMF_GR_Add, "MMF_XAttn_Add": MMF_XAttn_Add}
    from fusions.load_llm import get_context_window_size, get_d_model  # main.py:40 imports this name

    assert get_d_model("GPT2") == 768 and get_d_model("Llama") == 4096 and get_context_window_size("BERT") == 512


def test_perquery_state_dict_contract():
    """The per-(note, query) module loads the state_dict of the reference's TTF_T2V_XAttn_old class strictly."""
    import fusions.load_llm as L
    from fusions.TTF_T2V_XAttn_old import TTF_T2V_XAttn
    from test_perquery_cpu import load_pq

    cfg, params, inp, _ = load_pq("pq_h4")
    L.register_d_model("TINY", inp["notes"].shape[2])
    m = TTF_T2V_XAttn("TINY", 1, n_heads_fusion=cfg["H"])
    m.load_state_dict({k[len("ttf."):]: v for k, v in params.items()}, strict=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(inp["notes"], inp["tau"], inp["t_hat"])  # no CPU fallback


@pytest.mark.parametrize("name", golden_names())
def test_state_dict_contract(name):
    """Every parameter name and shape of the reference's state_dict (captured in the golden file)
    loads strictly into the drop-in module."""
    import gpu_common as G
    import fusions.load_llm as L
    from fusions.FusionModel import FusionModel

    cfg, params, inp, _ = load_golden(name)
    L.register_d_model("TINY", inp["notes"].shape[2])
    args = G.make_args(cfg["ttf"], cfg["mmf"], "TINY", cfg["d_txt"], cfg["C"], cfg["H"], cfg["kappa"], 0.1)
    args.device = "cpu"
    fm = FusionModel(args)
    fm.load_state_dict(params, strict=True)
    assert {k: tuple(v.shape) for k, v in fm.state_dict().items()} == {k: tuple(v.shape) for k, v in params.items()}
    assert fm.ttf.d_txt == (cfg["d_txt"] or inp["notes"].shape[2])
    # class objects are accepted in place of strings (FusionModel.py:45-50)
    args.TTF_module, args.MMF_module = type(fm.ttf), type(fm.mmf)
    fm2 = FusionModel(args)
    assert type(fm2.ttf) is type(fm.ttf) and type(fm2.mmf) is type(fm.mmf)


def test_constructor_asserts_like_reference():
    from fusions.TTF_RecAvg import TTF_RecAvg
    from fusions.TTF_T2V_XAttn import TTF_T2V_XAttn

    with pytest.raises(AssertionError):
        TTF_RecAvg("GPT2", 1, recency_sigma=0.0)
    with pytest.raises(AssertionError):
        TTF_T2V_XAttn("GPT2", 1, d_txt=2)  # d_tau = 1


def test_no_cpu_fallback():
    import gpu_common as G
    import fusions.load_llm as L
    from fusions.FusionModel import FusionModel

    L.register_d_model("TINY", 16)
    args = G.make_args("TTF_RecAvg", "MMF_GR_Add", "TINY", 8, 3, 1, 0.5, 0.0)
    fm = FusionModel(args)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        fm(torch.randn(2, 3, 16), torch.zeros(2, 3), torch.zeros(2, 4), torch.randn(2, 4, 3))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "imm-tsf_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "immtsf_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f


def test_philox_mirror_matches_device_code(tmp_path):
    """tests/philox_ref.py (numpy) == csrc/common.cuh (compiled for the host)."""
    import numpy as np
    import philox_ref as P

    src = tmp_path / "t.cu"
    src.write_text(
        '#include "common.cuh"\nvoid immtsf_set_error(const char*, ...) {}\n'
        "int main(){unsigned long long s=0x123456789abcdefull;for(unsigned long long c=0;c<4;++c){"
        "unsigned long long cc=c*0x100000001ull+7;Philox4 r=philox4x32_10((uint32_t)cc,(uint32_t)(cc>>32),3u,0u,(uint32_t)s,(uint32_t)(s>>32));"
        'printf("%llu %u %u %u %u\\n",cc,r.x,r.y,r.z,r.w);}}\n')
    exe = tmp_path / "t"
    r = subprocess.run(["nvcc", "-I", os.path.join(ROOT, "imm-tsf_b200", "csrc"), "-o", str(exe), str(src)], capture_output=True)
    if r.returncode != 0:
        pytest.skip("nvcc unavailable")
    seed = 0x123456789ABCDEF
    for line in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines():
        cc, *w = [int(x) for x in line.split()]
        got = P.philox4x32_10(np.array([cc & 0xFFFFFFFF], np.uint32), np.array([cc >> 32], np.uint32),
                              np.array([3], np.uint32), np.array([0], np.uint32), seed & 0xFFFFFFFF, seed >> 32)
        assert [int(x[0]) for x in got] == w


def test_keyed_dropout_mask_equals_the_per_call_one(tmp_path):
    """csrc/common.cuh compiled for the host: dropout_scale8_sw (Philox round keys computed once, thresholds on shifted words,
    the float4 halves in the lane's order -- what the one-launch RecAvg backward uses) draws exactly the mask of
    dropout_scale8 (one full Philox call per chunk -- what every forward kernel uses), for p = 0 (natural order) and p = 1
    (halves swapped), at drop thresholds 0, 1, 0.1 * 2^16 and 0xFFFF."""
    src = tmp_path / "k.cu"
    src.write_text(
        '#include "common.cuh"\nvoid immtsf_set_error(const char*, ...) {}\n'
        "int main(){int bad=0;const uint32_t thrs[4]={0u,1u,6553u,65535u};"
        "for(int t=0;t<4;++t)for(uint64_t s=1;s<40;s+=13){const uint64_t seed=s*0x9E3779B97F4A7C15ull;const PhiloxKeys k=philox_keys(seed);"
        "const float ik=thrs[t]==0u?1.f:(float)(1.0/(1.0-(double)thrs[t]/65536.0));"
        "for(uint64_t c=0;c<2000;++c){const uint64_t idx=c*0x10001ull+(c<<33);float a[8],b0[8],b1[8];"
        "dropout_scale8(seed,1u,idx,thrs[t],ik,a);dropout_scale8_sw(k,1u,idx,thrs[t],ik,0,b0);dropout_scale8_sw(k,1u,idx,thrs[t],ik,1,b1);"
        "for(int e=0;e<8;++e){if(a[e]!=b0[e])++bad;if(a[e]!=b1[e^4])++bad;}}}"
        'printf("%d\\n",bad);return 0;}\n')
    exe = tmp_path / "k"
    r = subprocess.run(["nvcc", "-I", os.path.join(ROOT, "imm-tsf_b200", "csrc"), "-o", str(exe), str(src)], capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("nvcc unavailable: " + r.stderr[-300:])
    assert subprocess.run([str(exe)], capture_output=True, text=True).stdout.strip() == "0"


def test_launcher_swaps_fusions_under_an_unmodified_script(tmp_path):
    """tools/run_with_immtsf.py: a script that lives next to its own `fusions` package (as the reference's
    main.py does) must get the B200 drop-in from the same import statements (main.py:39-40)."""
    import subprocess

    decoy = tmp_path / "fusions"
    decoy.mkdir()
    (decoy / "__init__.py").write_text("")
    (decoy / "FusionModel.py").write_text("class FusionModel: pass\n")
    (decoy / "load_llm.py").write_text("def get_context_window_size(*a, **k): return -1\n")
    (tmp_path / "main.py").write_text(
        "import sys\n"
        "from fusions.FusionModel import FusionModel\n"
        "from fusions.load_llm import get_context_window_size\n"
        "import fusions.FusionModel as M\n"
        "print('FILE', M.__file__)\n"
        "print('CTX', get_context_window_size('GPT2'))\n"
        "print('ARGS', sys.argv[1:])\n"
        "print('REG', sorted(M._TTF_CLASSES), sorted(M._MMF_CLASSES))\n")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "run_with_immtsf.py"), str(tmp_path / "main.py"), "--x", "1"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert os.path.join("imm-tsf_b200", "fusions", "FusionModel.py") in r.stdout
    assert "CTX 1024" in r.stdout and "ARGS ['--x', '1']" in r.stdout
    assert "['TTF_RecAvg', 'TTF_T2V_XAttn', 'TTF_T2V_XAttn_old'] ['MMF_GR_Add', 'MMF_XAttn_Add']" in r.stdout


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="the reference tree exists in the build container only")
def test_reference_caller_reaches_the_dropin():
    """lib/evaluation.py:72-164 `compute_all_losses`, imported UNMODIFIED, calls whatever `fusion` object it is given positionally
    (:95-100).  Handed the drop-in FusionModel on a CPU box it must reach the drop-in's forward -- which refuses CPU tensors
    (no CPU fallback) -- i.e. the caller-side wiring needs no change.  (The numerical check of this call sequence runs on the
    GPU: tests/test_gpu_caller.py, against golden vectors produced by this very function with the reference FusionModel.)"""
    code = r"""
import sys, types, torch
sys.path[:0] = ["%(root)s/imm-tsf_b200", "%(root)s/tools", "/root/reference"]
import run_with_immtsf; run_with_immtsf.install()
import fusions.load_llm as L; L.register_d_model("TINY", 48)
from fusions.FusionModel import FusionModel
import lib.evaluation as E
assert "imm-tsf_b200" in sys.modules["fusions.FusionModel"].__file__
a = types.SimpleNamespace(TTF_module="TTF_T2V_XAttn", MMF_module="MMF_XAttn_Add", llm_model_fusion="TINY", llm_layers_fusion=1, max_length=1024,
                          device="cpu", use_text_embeddings=True, recency_sigma=1.0, dropout=0.0, d_txt=32, n_heads_fusion=1, C=4, kappa=0.5)
fm = FusionModel(a)
class M(torch.nn.Module):
    def forecasting(self, tp, x, t, m): return torch.zeros(tp.shape[0], tp.shape[1], 4)
b = dict(notes_embeddings=torch.randn(2, 3, 48), tau=torch.rand(2, 3), tp_to_predict=torch.rand(2, 5), observed_data=torch.zeros(2, 4, 4),
         observed_tp=torch.zeros(2, 4), observed_mask=torch.ones(2, 4, 4), data_to_predict=torch.zeros(2, 5, 4), mask_predicted_data=torch.ones(2, 5, 4))
try:
    E.compute_all_losses(M(), fm, b)
except RuntimeError as e:
    assert "no CPU fallback" in str(e), e
    print("REACHED_DROPIN")
""" % dict(root=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "REACHED_DROPIN" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]

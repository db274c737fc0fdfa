"""What the in-tree library is made of, read from its SASS on the host (no GPU): every kernel family DESIGN.md describes as tcgen05 / TMEM /
TMA really carries those instructions, the library holds sm_100a code only, and the op's backward reduces with 16-byte vector reductions.
Mnemonics as in /opt/skills/guides/B200_PROFILING.md: UTC*MMA = tcgen05.mma (.2CTA = cta_group::2), LDTM = tcgen05.ld, UTMALDG / UTMASTG /
UTMAREDG = cp.async.bulk.tensor load / store / reduce, UBLKCP = cp.async.bulk, HMMA.1688.F32.TF32 = mma.sync.m16n8k8 tf32."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "transoar_b200", "libmsda3d.so")

pytestmark = pytest.mark.skipif(shutil.which("cuobjdump") is None or not os.path.exists(LIB), reason="needs cuobjdump and the built library")

_MNEMONIC = re.compile(r"\b(UTC[A-Z]*MMA(?:\.2CTA)?|UTMALDG\.[0-9]D(?:\.2CTA)?|UTMASTG\.[0-9]D|UTMAREDG\.[0-9]D|LDTM\.x[0-9]+|UBLKCP|HMMA\.1688\.F32\.TF32|"
                       r"REDG\.E\.ADD\.F32x4|REDG\.E\.ADD\.F32x2|UTCBAR(?:\.2CTA\.MULTICAST)?|FFMA2|FMUL2)\b")


@pytest.fixture(scope="module")
def sass():
    """namespace -> {mangled kernel name -> Counter of the mnemonics above}"""
    try:
        proc = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True, timeout=600)
    except (subprocess.SubprocessError, OSError) as exc:          # the tool itself is not under test
        pytest.skip(f"cuobjdump -sass failed: {exc}")
    table, cur = collections.defaultdict(dict), None
    for line in proc.stdout.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            m = re.match(r"_ZN(\d+)", name)
            ns = name[3 + len(m.group(1)):3 + len(m.group(1)) + int(m.group(1))] if m else "?"
            cur = table[ns].setdefault(name, collections.Counter())
        elif cur is not None:
            for t in _MNEMONIC.findall(line):
                cur[t] += 1
    return table


def _kernels(table, ns, fragment):
    return {k: v for k, v in table[ns].items() if fragment in k}


def _total(kernels, prefix):
    return sum(n for c in kernels.values() for t, n in c.items() if t.startswith(prefix))


def test_library_holds_sm_100a_code_only():
    try:
        out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True, check=True, timeout=600).stdout
    except (subprocess.SubprocessError, OSError) as exc:
        pytest.skip(f"cuobjdump -lelf failed: {exc}")
    cubins = [l for l in out.splitlines() if "ELF file" in l]
    assert len(cubins) >= 10 and all("sm_100a" in l for l in cubins), cubins


def test_every_kernel_family_is_present(sass):
    assert {"msda3d", "roiattn", "tcgemm", "convtc", "convgen", "instnorm", "stemconv", "fusedln", "winattn", "crit"} <= set(sass)


def test_gemm_kernels_are_tcgen05_tmem_tma(sass):
    for frag in ("gemm_tf32_kernel", "gemm_tf32_pair_kernel", "gemm_bf16_kernel", "gemm_bf16_pair_kernel"):
        ks = _kernels(sass, "tcgemm", frag)
        assert ks, frag
        for name, c in ks.items():
            assert _total({name: c}, "UTC") > 0 and _total({name: c}, "UTMALDG") > 0 and _total({name: c}, "LDTM") > 0, name
    pair = _kernels(sass, "tcgemm", "pair_kernel")
    assert all(c["UTCHMMA.2CTA"] > 0 and c["UTMALDG.2D.2CTA"] > 0 and c["UTCBAR.2CTA.MULTICAST"] > 0 for c in pair.values())   # cta_group::2
    assert _total(sass["tcgemm"], "UTMASTG") > 0                                   # the TMA-store epilogue


def test_convolution_kernels_are_tcgen05_fed_by_tma(sass):
    for ns, frags in (("convtc", ("conv3d_k3_kernel", "conv3d_k3_wgrad_kernel")), ("convgen", ("conv_halo_kernel", "conv_kmajor_kernel", "conv_wgrad_kernel"))):
        for frag in frags:
            ks = _kernels(sass, ns, frag)
            assert ks, (ns, frag)
            for name, c in ks.items():
                assert c["UTCHMMA"] > 0 and _total({name: c}, "UTMALDG") > 0, name
    assert _total(sass["convgen"], "UTMAREDG") > 0                                 # split-K by cp.reduce.async.bulk.tensor
    assert _total(sass["convgen"], "UTMASTG") > 0


def test_roi_attention_runs_its_contractions_on_tensor_cores(sass):
    for frag in ("fwd_tc_kernel", "bwd_tc_kernel"):
        ks = _kernels(sass, "roiattn", frag)
        assert ks and all(c["HMMA.1688.F32.TF32"] > 0 for c in ks.values()), frag
    assert all(c["REDG.E.ADD.F32x2"] > 0 for c in _kernels(sass, "roiattn", "bwd_tc_kernel").values())      # dk / dv as red.global.add.v2.f32


def test_op_backward_uses_vector_reductions_and_the_staged_forward_uses_bulk_copies(sass):
    ints = lambda name: [int(v) for v in re.findall(r"Li(\d+)E", name.split("EEv")[0])]
    # template arguments: bwd_duo_kernel<FUSED, ROT, SKIP, ...>, bwd_vec_kernel<float, G, NV, MINB, SKIP, ...>; SKIP = 1 is the diagnostic
    # instantiation with the grad_value reductions compiled out (profiles/r02_experiments.md 2, 13f)
    duo = {k: c for k, c in _kernels(sass, "msda3d", "bwd_duo_kernel").items() if ints(k)[2] != 1}
    vec = {k: c for k, c in _kernels(sass, "msda3d", "bwd_vec_kernelIf").items() if ints(k)[3] != 1}
    assert duo and vec
    assert all(c["REDG.E.ADD.F32x4"] > 0 for c in duo.values()) and all(c["REDG.E.ADD.F32x4"] > 0 for c in vec.values())
    assert all(c["UBLKCP"] > 0 for c in _kernels(sass, "msda3d", "fwd_stage_kernel").values())
    # the op's kernels run on CUDA cores: no tensor-core / tensor-map instruction, and no packed fp32 arithmetic yet
    assert _total(sass["msda3d"], "UTC") == 0 and _total(sass["msda3d"], "UTMALDG") == 0
    assert _total(sass["msda3d"], "FFMA2") == 0 and _total(sass["msda3d"], "FMUL2") == 0


def test_register_and_stack_budgets_of_the_shipped_kernels():
    """`cuobjdump -res-usage`: the tcgen05 kernels keep everything in registers (no stack frame = no spills), and the op's default
    kernels fit the occupancy DESIGN.md quotes: forward <= 64 registers (4 CTAs of 256 threads per SM), one-unit backward <= 80 (3 CTAs),
    pair kernel <= 128 (2 CTAs)."""
    try:
        out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True, timeout=600).stdout
    except (subprocess.SubprocessError, OSError) as exc:
        pytest.skip(f"cuobjdump -res-usage failed: {exc}")
    rows = {n: (int(reg), int(stack)) for n, reg, stack in re.findall(r"Function (\S+):\n\s+REG:(\d+) STACK:(\d+)", out)}
    assert len(rows) >= 400
    for prefix in ("_ZN6tcgemm", "_ZN6convtc", "_ZN7convgen", "_ZN7winattn", "_ZN7fusedln", "_ZN4crit"):
        fam = {n: v for n, v in rows.items() if n.startswith(prefix)}
        assert fam and all(stack == 0 for _, stack in fam.values()), prefix
    reg = lambda frag: [v[0] for n, v in rows.items() if frag in n]
    assert reg("fwd_vec_kernelIfLi16ELi1ELi4E") and max(reg("fwd_vec_kernelIfLi16ELi1ELi4E")) <= 64
    assert reg("bwd_vec_kernelIfLi16ELi1ELi3E") and max(reg("bwd_vec_kernelIfLi16ELi1ELi3E")) <= 80
    shipped_duo = [v for n, v in rows.items() if "bwd_duo_kernel" in n and "Li256ELi2E" in n]
    assert shipped_duo and max(r for r, _ in shipped_duo) <= 128 and max(s for _, s in shipped_duo) <= 16

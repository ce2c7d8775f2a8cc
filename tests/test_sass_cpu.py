"""Static checks on the SASS of the built library (no GPU needed): the tensor-core kernels really
are tcgen05 + TMA + TMEM kernels, and the per-tile hand-back of the TMEM accumulators carries no
GPU-scope memory barrier (profiles/README.md, `arrive_ab_r1.log`)."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cplxmodule_b200", "csrc", "libcplxk.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

TC_KERNELS = ("fwd_tc_kernel", "fwd_tc3_kernel", "lin_tc3_kernel", "conv_tc_kernel",
              "conv_tc_persistent_kernel", "conv_tc_pair_kernel")
CLUSTER_KERNELS = ("fwd_tc3_kernel", "lin_tc3_kernel", "conv_tc_pair_kernel")


@pytest.fixture(scope="module")
def sass():
    if not os.path.exists(LIB) or not os.path.exists(CUOBJDUMP):
        pytest.skip("library or cuobjdump not present")
    out = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, timeout=300).stdout
    per_fn, name = collections.defaultdict(collections.Counter), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
        if m and name:
            per_fn[name][m.group(1)] += 1
            per_fn[name][m.group(1) + m.group(2)] += 1
    return per_fn


def kernels_of(per_fn, stem):
    # mangled: _ZN5cplxk<len><name>I...
    return {k: v for k, v in per_fn.items() if re.search(r"cplxk\d+" + stem + "I", k)}


def test_sm100a_only():
    if not os.path.exists(LIB) or not os.path.exists(CUOBJDUMP):
        pytest.skip("library or cuobjdump not present")
    out = subprocess.run([CUOBJDUMP, "-lelf", LIB], capture_output=True, text=True, timeout=120).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("stem", TC_KERNELS)
def test_tensor_core_kernels_use_tcgen05_tma_tmem(sass, stem):
    ks = kernels_of(sass, stem)
    assert ks, f"no instantiation of {stem} in the library"
    for name, ops in ks.items():
        assert ops["UTCHMMA"] > 0, f"{name}: no tcgen05.mma (UTCHMMA)"
        assert ops["UTMALDG"] > 0, f"{name}: no TMA load (UTMALDG)"
        assert ops["LDTM"] > 0, f"{name}: no tcgen05.ld (LDTM)"
        assert ops["HMMA"] == 0 and ops["IMMA"] == 0, f"{name}: legacy mma.sync present"


@pytest.mark.parametrize("stem", CLUSTER_KERNELS)
def test_cta_pair_kernels(sass, stem):
    for name, ops in kernels_of(sass, stem).items():
        assert ops["UTCHMMA.2CTA"] > 0, f"{name}: MMAs are not cta_group::2"
        # the two cluster-wide syncs (after barrier init, before TMEM dealloc) and nothing per tile
        assert ops["MEMBAR.ALL.GPU"] <= 2, f"{name}: GPU-scope membar inside the tile loop"


def test_real_and_grouped_conv_variants_are_built(sass):
    """conv_tc_kernel<T, VD, real> for both plane types, with and without the
    variational part, complex and real planes -- eight tcgen05 kernels"""
    names = kernels_of(sass, "conv_tc_kernel")
    assert len(names) == 8, sorted(names)
    # real-plane instantiations issue fewer MMAs per k-step than the complex ones (one A plane)
    real = [v["UTCHMMA"] for k, v in names.items() if k.endswith("Lb1EEEv14CUtensorMap_stS2_S2_S2_S2_S2_NS_10ConvTcGeomENS_9ConvTcEpiE")
            or "Lb1EEEv14CUtensorMap_stS1_" in k]
    assert real, "no real-plane instantiation found"


def test_guard_and_fingerprint_kernels_exist(sass):
    assert any("kl_guard_kernel" in k for k in sass), "kl_guard_kernel missing"
    assert any("vd_grad_s2_torch_kernel" in k for k in sass), "vd_grad_s2_torch_kernel missing"


def test_pair_conv_kernel_variants_are_built(sass):
    """conv_tc_pair_kernel<T, half operands, row mode, real planes>: complex bf16 / tf32 / scaled-fp16
    and real bf16 / tf32, each with per-tap and per-row activation loads -- ten cta_group::2 kernels;
    a row-mode instantiation issues its taps from one loaded tile (no extra TMA opcode per tap)."""
    names = kernels_of(sass, "conv_tc_pair_kernel")
    assert len(names) == 10, sorted(names)
    row = [k for k in names if re.search(r"Lb[01]ELb1ELb[01]EEEv", k)]
    real = [k for k in names if re.search(r"Lb[01]ELb[01]ELb1EEEv", k)]
    assert len(row) == 5 and len(real) == 4, (row, real)
    assert any("vd_combine_torch_flat_kernel" in k for k in sass), "flat noise kernel missing"
    assert sum("conv_nhwc_f16_v4_kernel" in k for k in sass) == 3, "optimistic / fix-up conversion kernels"

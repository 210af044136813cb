"""CPU: the C-ABI library loads, exports every symbol include/fieldconv_b200.h declares, and its host-side
argument checking / workspace queries behave (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT
from fieldconv_b200 import _lib


def header_symbols():
    src = open(os.path.join(ROOT, "include", "fieldconv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fcb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    names = header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), "missing export " + n


def test_binding_covers_header():
    bound = set(_lib.SIGNATURES) | {"fcb_last_error", "fcb_launch_count", "fcb_pk_supported", "fcb_gemm_tc_feasible", "fcb_fused_supported"}
    assert set(header_symbols()) <= bound


def header_prototypes():
    """name -> list of parameter declarations of every `int fcb_*(...)` prototype in the header"""
    src = open(os.path.join(ROOT, "include", "fieldconv_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|const char\*|unsigned long long)\s+(fcb_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", src, flags=re.S):
        params = [a.strip() for a in m.group(2).split(",")]
        out[m.group(1)] = [] if params in (["void"], [""]) else params
    return out


def _kind(decl):
    """pointer / i64 / int / float / size of one C parameter declaration, as the ctypes binding has to mirror it"""
    if "*" in decl:
        return "ptr"
    if "int64_t" in decl:
        return "i64"
    if "size_t" in decl:
        return "size"
    if "float" in decl:
        return "float"
    return "int"


def test_ctypes_signatures_match_the_header_prototypes():
    """Every binding in _lib.SIGNATURES has the argument count and the argument kinds (pointer / int64 / int / float / size_t)
    of its prototype: a signature changed in the header but not in the binding shifts every later argument silently."""
    kinds = {_lib._P: "ptr", _lib._PSZ: "ptr", ctypes.c_char_p: "ptr", ctypes.POINTER(ctypes.c_float): "ptr", ctypes.POINTER(ctypes.c_int): "ptr",
             _lib._I64: "i64", _lib._I: "int", _lib._F: "float", _lib._SZ: "size"}
    protos = header_prototypes()
    checked = 0
    for name, argtypes in _lib.SIGNATURES.items():
        assert name in protos, name + " is bound but not declared in the header"
        want = [_kind(d) for d in protos[name]]
        got = [kinds[a] for a in argtypes]
        assert got == want, "%s: binding %s, header %s" % (name, got, want)
        checked += 1
    assert checked >= 30


def test_version_and_counter():
    lib = _lib.load()
    assert lib.fcb_version() >= 100
    assert _lib.launch_count() >= 0


def test_workspace_queries_and_errors():
    n = _lib.query_bytes("fcb_fwd_workspace_bytes", 1000, 32, 32, 1, 6, 0)
    assert n >= 4 * (6 * 32 * 3) * 32 * 4
    # flags 0: nothing kept by the forward -> gW from G and xhat (room for the xhat operand, never for an N x K contrib);
    # 0x100 (FCB_FLAG_HAVE_CONTRIB): gW from the saved contrib
    from_g = _lib.query_bytes("fcb_bwd_workspace_bytes", 1000, 32, 32, 1, 6, 3)
    have = _lib.query_bytes("fcb_bwd_workspace_bytes", 1000, 32, 32, 1, 6, 3 | 0x100)
    g_bytes = 1024 * 6 * 32 * 3 * 8
    assert from_g >= g_bytes + 1000 * 3 * 32 * 8 and have >= g_bytes
    assert from_g < have + 2 * g_bytes        # no recomputed-contrib buffer on top of G
    assert _lib.query_bytes("fcb_plan_workspace_bytes", 10000, 500, 6) > 7 * 10000 * 4
    assert _lib.query_bytes("fcb_sort_workspace_bytes", 0) > 0
    with pytest.raises(RuntimeError, match="even"):
        _lib.query_bytes("fcb_fwd_workspace_bytes", 10, 3, 4, 1, 6, 0)          # odd channel count
    with pytest.raises(RuntimeError, match="band_limit"):
        _lib.query_bytes("fcb_fwd_workspace_bytes", 10, 4, 4, 9, 6, 0)
    out = ctypes.c_size_t(0)
    rc = _lib.load().fcb_plan_workspace_bytes(-1, 5, 6, ctypes.byref(out))
    assert rc == -1 and b"bad arguments" in _lib.load().fcb_last_error()


def test_null_pointer_calls_are_rejected_without_touching_the_gpu():
    lib = _lib.load()
    rc = lib.fcb_fwd_f32(None, None, None, None, None, None, None, None, 10, 4, 4, 1, 6, 0, None, 0, None)
    assert rc == -1
    rc = lib.fcb_plan_build(None, None, None, None, None, None, 1.0, 10, 10, 1, None, None, None, None,
                            None, None, None, None, None, 0, None)
    assert rc == -5      # n_rings = 1 is unsupported (the reference divides by n_rings-1)
    rc = lib.fcb_gemm_f32(None, None, None, 4, 4, 4, 4, 4, 4, 0, 1, 0, 0, 0, 1, None, None, None, 0, 0, None)
    assert rc == -1


def test_tensor_core_accumulation_plan():
    """auto precision: tensor cores only where no TMEM accumulator would take more than 400 accumulating MMAs."""
    assert _lib.tc_feasible(96, 2880)                 # cfg 2 forward (C=48, B=2, R=6)
    assert _lib.tc_feasible(256, 7680)                # cfg 3 forward (C=128): two 128-column chunks
    assert _lib.tc_feasible(512, 21504)               # C=256, B=3: 64-column chunks, 7 accumulators
    assert not _lib.tc_feasible(96, 2880, flags=0)    # FP32-FMA mode never uses tensor cores
    assert _lib.tc_feasible(96, 80000, trans_a=1, split_k=26)
    assert not _lib.tc_feasible(64, 10 ** 6)          # too many accumulating steps for any plan
    # 2xFP16 mode (flags=3): 16 reals per MMA, (main, cross) accumulator pairs
    assert _lib.tc_feasible(96, 2880, flags=3)
    assert _lib.tc_feasible(256, 7680, flags=3)       # two 128-column chunks, two pairs each
    assert _lib.tc_feasible(96, 80000, trans_a=1, split_k=26, flags=3)
    assert not _lib.tc_feasible(64, 10 ** 6, flags=3)


def test_fused_forward_shape_support():
    # host-side predicate only: band_limit <= 1, Ci a multiple of 32, Co even and <= 128, <= 400 accumulating MMAs
    assert _lib.fused_supported(32, 32, 1, 6) and _lib.fused_supported(64, 64, 1, 6) and _lib.fused_supported(128, 128, 1, 6)
    assert _lib.fused_supported(32, 16, 1, 2) and _lib.fused_supported(64, 48, 1, 2)
    assert not _lib.fused_supported(48, 48, 1, 6)          # Ci not a multiple of 32
    assert not _lib.fused_supported(32, 32, 2, 6)          # band_limit 2: the ring burst does not fit shared memory
    assert not _lib.fused_supported(256, 256, 1, 6)        # 576 accumulating MMAs / Co too wide for one TMEM tile
    assert _lib.query_bytes("fcb_fwd_fused_workspace_bytes", 32, 32, 1, 6) >= 18 * 2 * 64 * 64 * 2
    with pytest.raises(RuntimeError, match="not supported"):
        _lib.query_bytes("fcb_fwd_fused_workspace_bytes", 48, 48, 2, 6)


def test_packed_path_shape_support_and_sizes():
    # 2*R*M*Ci must be a multiple of 64 and the 2xFP16 accumulation plans must fit (host-side predicates only)
    assert _lib.pk_supported(80656, 48, 48, 2, 6)          # BASELINE cfg 2 layer
    assert _lib.pk_supported(5041, 32, 32, 1, 6)           # cfg 1
    assert _lib.pk_supported(6889, 128, 128, 2, 6)         # cfg 3
    assert not _lib.pk_supported(5041, 6, 6, 1, 3)         # 2*3*3*6 = 108 columns: not a multiple of 64
    n, ci, b, r = 1000, 32, 1, 6
    nbytes = _lib.query_bytes("fcb_pk_contrib_bytes", n, ci, b, r)
    assert nbytes == 1024 * (r * ci * 3) * 8               # same bytes as the fp32 layout, rows padded to 128
    lib = _lib.load()
    rc = lib.fcb_fwd_pk_f32(None, None, None, None, None, None, None, None, None, None, 10, 32, 32, 1, 6, 3, None, 0, None)
    assert rc == -1
    rc = lib.fcb_plan_norm(None, None, 10, None, None)
    assert rc == -1


def test_default_aggregation_variants_are_compiled_without_heavy_spills():
    """The variants agg_variant() selects by default (aggregate_kernel.cuh) exist in the library for band limits 1 and 2,
    both directions, fp32 and packed output, and stay within the spill budget they were measured with."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-res-usage", _lib.LIB_PATH], capture_output=True, text=True).stdout
    lines = out.splitlines()
    usage = {}
    for i, l in enumerate(lines):
        m = re.search(r"Function _ZN3fcb11k_aggregateILi(\d)ELb(\d)ELb(\d)ELi(\d)ELb(\d)EEEv", l)
        if m and i + 1 < len(lines):
            u = re.search(r"REG:(\d+) STACK:(\d+)", lines[i + 1])
            usage[tuple(int(x) for x in m.groups())] = (int(u.group(1)), int(u.group(2)))
    # (band limit, transpose, packed, resident CTAs, pointer-increment packed store) — the dispatcher of aggregate_kernel.cuh
    ctas = {0: 3, 1: 4, 2: 3, 3: 2}
    defaults = [(b, t, pk, ctas[b], 0) for b in (0, 1, 2, 3) for t in (0, 1) for pk in (0, 1)]
    defaults += [(b, t, 1, ctas[b], 1) for b in (1, 2) for t in (0, 1)]
    for key in defaults:
        assert key in usage, "missing kernel variant %s" % (key,)
        regs, stack = usage[key]
        assert regs <= {2: 128, 3: 80, 4: 64}[key[3]] and stack <= 80, (key, regs, stack)


def _edge_loop_histogram(mangled):
    """Opcode histogram of the (two-edge unrolled) edge loop of one k_aggregate instantiation in the built library: the
    smallest backward-branch loop that holds >= 40 FFMA2 outside its inner loops (the inlined ring-retire blocks)."""
    import collections
    import subprocess
    out = subprocess.run(["cuobjdump", "-sass", "-fun", mangled, _lib.LIB_PATH], capture_output=True, text=True).stdout
    ins = []
    for l in out.splitlines():
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    loops = []
    for a, t in ins:
        m = re.search(r"BRA\S*\s+(?:\S+,\s+)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            loops.append((int(m.group(1), 16), a))
    best = None
    for lo, hi in loops:
        inner = [(a, b) for a, b in loops if lo < a and b < hi]
        body = [t for a, t in ins if lo <= a <= hi and not any(x <= a <= y for x, y in inner)]
        if sum("FFMA2" in t for t in body) >= 40 and (best is None or hi - lo < best[0]):
            best = (hi - lo, body)
    assert best is not None, "edge loop not found in " + mangled
    c = collections.Counter()
    for t in best[1]:
        tt = t.split()
        c[(tt[1] if tt[0].startswith("@") else tt[0]).split(".")[0]] += 1
    return c


def test_aggregation_edge_loop_instruction_budget():
    """The cfg-2 aggregation kernels are bound by instruction issue / the FMA pipe, so the instruction count of
    their edge loop IS their speed: per two edges (the loop is unrolled twice) the arithmetic is fixed by the formulation
    (52 FFMA2; 64 resp. 32 scalar FMUL/FFMA) and the overhead must stay where round 2 left it (forward 174, transposed 116;
    they were 200 and 143 before the r04d trimming).  Static check on the built library, no GPU needed."""
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    sig = "EEEvPK6float4PKiPK4int4PK6float2PS1_liiPjPKfSF_Pf"
    fwd = _edge_loop_histogram("_ZN3fcb11k_aggregateILi2ELb0ELb0ELi3ELb0" + sig)       # forward, fp32 contrib
    tr = _edge_loop_histogram("_ZN3fcb11k_aggregateILi2ELb1ELb1ELi3ELb1" + sig)        # transposed, PK G, pointer-increment store
    for name, c, scalar, budget in (("forward", fwd, 64, 182), ("transposed", tr, 32, 122)):
        assert c["FFMA2"] == 52, (name, dict(c))
        assert c["FMUL"] + c["FFMA"] == scalar, (name, dict(c))
        assert c["LDG"] == 2 and c["LDS"] == 6, (name, dict(c))            # one gather, two record halves + the next offset per edge
        assert c["LDL"] == 0 and c["STL"] == 0, (name, dict(c))            # no spill traffic inside the loop
        assert sum(c.values()) <= budget, (name, sum(c.values()), dict(c))


def test_contraction_kernels_carry_tcgen05_and_bulk_copy_sass():
    """The 2xFP16 / 3xTF32 contraction kernels and the fused forward really are tcgen05 kernels: their SASS holds UTCHMMA
    (tcgen05.mma), LDTM (tcgen05.ld from TMEM), UBLKCP (cp.async.bulk) and SYNCS (mbarrier); the aggregation kernels hold
    packed FFMA2.  Static check on the built library (profiles/r04_sass_summary.txt is the same listing per kernel)."""
    import shutil
    import subprocess
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    per_fn, fn = {}, None
    for l in out.splitlines():
        m = re.search(r"Function : (\S+)", l)
        if m:
            fn = m.group(1)
            per_fn[fn] = {"UTCHMMA": 0, "LDTM": 0, "UBLKCP": 0, "SYNCS": 0, "FFMA2": 0}
            continue
        if fn:
            for k in per_fn[fn]:
                if re.search(r"\b%s\b" % k, l):
                    per_fn[fn][k] += 1
    def total(pattern, key):
        return sum(c[key] for f, c in per_fn.items() if pattern in f)
    for kernel in ("k_gemm_h_nn", "k_gemm_h_tn", "k_gemm_h_nn_small", "k_gemm_tc_nn", "k_gemm_tc_tn", "k_fused_fwd"):
        assert total(kernel, "UTCHMMA") > 0 and total(kernel, "LDTM") > 0 and total(kernel, "SYNCS") > 0, kernel
    assert total("k_gemm_h_nn", "UBLKCP") > 0 and total("k_gemm_h_tn", "UBLKCP") > 0
    assert total("k_aggregate", "FFMA2") > 0 and total("k_aggregate", "UTCHMMA") == 0

"""CPU-side checks of the product's host logic: the C-ABI library loads and exports every symbol of
include/sia_b200.h, the banded resize tables equal the oracle's operator, the count -> metrics
arithmetic equals the oracle / reference fixtures, and the product path refuses to run without CUDA."""
import contextlib
import io
import json
import os
import re

import numpy as np
import pytest
import torch

from oracle import analysis as oa
from oracle import resize as R
from tests import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header_name):
    header = open(os.path.join(ROOT, "include", header_name)).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    return re.findall(r"\b(sia_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", header, flags=re.S)


def test_library_exports_every_declared_symbol():
    """include/sia_b200.h == the product library's bound symbols (plus the three instrumentation switches of
    sia_b200_debug.h group 1); the probes of group 2 live in libsia_b200_debug.so and NOT in the product library."""
    from skin_image_analysis_b200 import _lib, build
    build.build()
    lib, dbg = _lib.load(), _lib.load_debug()
    product = {n for n, _ in _declared("sia_b200.h")}
    debug = {n for n, _ in _declared("sia_b200_debug.h")}
    assert product and debug and not (product & debug)
    switches = {"sia_debug_tv_force_generic", "sia_debug_set_trace", "sia_debug_set_stats", "sia_debug_set_mma_warps",
                "sia_debug_set_programmatic_launch", "sia_debug_set_tail_impl"}
    assert not any(n.startswith("sia_debug") for n in product)
    for name in sorted(product | switches):
        assert hasattr(lib, name), f"{name} declared but not exported by libsia_b200.so"
    for name in sorted(debug - switches):
        assert hasattr(dbg, name), f"{name} declared but not exported by libsia_b200_debug.so"
        assert not hasattr(lib, name), f"probe {name} leaked into the product library"
    assert product | switches == set(_lib.SIGNATURES), (product | switches) ^ set(_lib.SIGNATURES)
    assert debug - switches == set(_lib.DEBUG_SIGNATURES), (debug - switches) ^ set(_lib.DEBUG_SIGNATURES)
    assert lib.sia_version() == 100
    assert lib.sia_error_string(-2).decode().startswith("sia:")
    assert lib.sia_pack_conv7x7_c3_bytes() == 128 * 256 * 2
    assert lib.sia_pack_conv3x3_bytes(64, 128) == 9 * 64 * 128 * 2


def test_ctypes_signatures_match_the_header_prototypes():
    """Every prototype of include/sia_b200.h and sia_b200_debug.h against _lib.SIGNATURES / DEBUG_SIGNATURES: same
    number of parameters, pointers bound as pointers and integers as integers (a missing argument shifts the stream
    handle and crashes on the GPU box)."""
    import ctypes
    from skin_image_analysis_b200 import _lib
    protos = _declared("sia_b200.h") + _declared("sia_b200_debug.h")
    bound = {**_lib.SIGNATURES, **_lib.DEBUG_SIGNATURES}
    assert len(protos) == len(bound)
    for name, params in protos:
        params = " ".join(params.split())
        plist = [] if params in ("", "void") else [q.strip() for q in params.split(",")]
        _res, args = bound[name]
        assert len(plist) == len(args), f"{name}: header has {len(plist)} parameters, _lib binds {len(args)}"
        for q, a in zip(plist, args):
            is_ptr = "*" in q
            bound_ptr = a is ctypes.c_void_p or a is ctypes.c_char_p or isinstance(a, type(ctypes.POINTER(ctypes.c_int)))
            assert is_ptr == bound_ptr, f"{name}: parameter '{q}' bound as {a}"


@pytest.mark.parametrize("shape", [(450, 600, 224, 224), (450, 600, 512, 512), (97, 131, 64, 86), (600, 450, 298, 224)])
def test_resize_tables_equal_oracle_operator(shape):
    from skin_image_analysis_b200 import resize_weights as rw
    h, w, oh, ow = shape
    t = rw.build_tables(h, w, oh, ow, keep_dense=True)
    aa = oh < h or ow < w
    assert np.abs(t.wy_dense - R.axis_weight_matrix(h, oh, aa)).max() < 1e-15
    assert np.abs(t.wx_dense - R.axis_weight_matrix(w, ow, aa)).max() < 1e-7     # x_w is stored as float32
    assert t.x_taps in (8, 16) and np.all(t.x_off >= 0) and np.all(t.x_off + t.x_taps <= w)
    # every output row is emitted exactly once, after its last contributing source row
    emitted = t.row_emit[t.row_emit >= 0]
    assert sorted(emitted.tolist()) == list(range(oh))
    for r in range(h):
        for s in range(4):
            e = t.row_emit[r, s]
            if e >= 0:
                assert t.y_first_last[e, 1] == r and e % 4 == s
    # emulate the kernel's slot schedule on a random column
    col = np.random.default_rng(0).random(h)
    acc, out = np.zeros(4), np.zeros(oh)
    for r in range(h):
        for s in range(4):
            acc[s] += float(t.row_w[r, s]) * col[r]
            if t.row_emit[r, s] >= 0:
                out[t.row_emit[r, s]] = acc[s]
                acc[s] = 0.0
    assert np.abs(out * 255.0 - R.axis_weight_matrix(h, oh, aa) @ col).max() < 1e-6


def test_resize_tables_reject_unsupported_geometry():
    from skin_image_analysis_b200 import resize_weights as rw
    with pytest.raises(ValueError):
        rw.build_tables(2000, 2000, 100, 100)          # 20x shrink: window far beyond 16 taps
    with pytest.raises(ValueError):
        rw.build_tables(4, 4, 8, 8)                    # narrower than the 8-tap window
    assert rw.rescale_size(450, 600, 224) == (224, 298)


@pytest.mark.parametrize("case", ["n500", "n64", "n1087"])
def test_results_from_counts_equals_reference_fixture(golden_dir, case):
    """Counts built on the CPU by the oracle -> product host arithmetic -> the reference's exact dict+stdout."""
    from skin_image_analysis_b200 import tone_bias_test as tt
    with open(os.path.join(golden_dir, "analysis_synth.json")) as f:
        g = json.load(f)[case]
    inst = helpers.synthetic_instances(g["n"], g["seed"], g["with_oddities"])
    keys, pred, label, groups = tt.encode_instances(inst)
    counts = np.zeros((len(tt.ATTRIBUTES), tt.N_GROUPS, 2, 2), np.int64)
    for a in range(groups.shape[0]):
        ok = groups[a] < tt.N_GROUPS
        np.add.at(counts, (a, groups[a][ok], label[ok], pred[ok]), 1)
    tab = oa.counts_table(inst, {"skin_tone": ["light", "dark"], "sex": ["male", "female"]})
    assert counts[0, :2].tolist() == tab["skin_tone"] and counts[1, :2].tolist() == tab["sex"]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        res = tt.results_from_counts(counts)
    assert json.loads(json.dumps(res)) == g["result"]
    assert buf.getvalue() == g["stdout"]
    # the compact [3][6][2][2] tensor of the sharded engine maps to the same result
    compact = np.stack([counts[3], counts[1], counts[2]])
    res2 = tt.results_from_type_counts(compact, out=lambda *a: None)
    assert json.loads(json.dumps(res2)) == g["result"]


def test_notebook_known_answer_through_product_arithmetic(golden_dir):
    from skin_image_analysis_b200 import tone_bias_test as tt
    with open(os.path.join(golden_dir, "notebook_di.json")) as f:
        nb = json.load(f)
    for attr in ("tone", "sex"):
        c = nb[attr]["cells"]
        mn = [[c["tn_min"], c["fp_min"]], [c["fn_min"], c["tp_min"]]]
        mj = [[c["tn_maj"], c["fp_maj"]], [c["fn_maj"], c["tp_maj"]]]
        r = tt.di_from_tables(mn, mj)
        assert r == oa.di_from_cells(c["tp_min"], c["tn_min"], c["fp_min"], c["fn_min"],
                                     c["tp_maj"], c["tn_maj"], c["fp_maj"], c["fn_maj"])
        assert f"{r['di']:.3f}" == f"{nb[attr]['printed']['di']:.3f}"
        assert list(r.keys()) == list(tt.DI_KEYS) and len(r) == 27


def test_model_drop_in_surface_on_cpu():
    """Construction, state_dict compatibility and save/load work anywhere; forward needs the GPU."""
    from oracle import model as om
    from skin_image_analysis_b200 import jgi_hiba_2022_model as hm
    from skin_image_analysis_b200 import tone_bias_model as tm
    from skin_image_analysis_b200._lib import SiaError
    for kind, cls in ((om.LIST_MODEL, tm.SkinCancerListModel), (om.FOUR_CONV_MODEL, hm.SkinCancerModel)):
        m = cls(helpers.CLASS_NAMES)
        shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
        assert shapes == om.param_shapes(kind)
        assert m.get_class_names() == helpers.CLASS_NAMES
    assert isinstance(tm.create_model(helpers.CLASS_NAMES), tm.SkinCancerModel)
    assert isinstance(tm.create_loss_function(), torch.nn.NLLLoss)
    m = tm.SkinCancerListModel(helpers.CLASS_NAMES).eval()
    if not torch.cuda.is_available():
        with pytest.raises(SiaError):
            m(torch.rand(1, 3, 224, 224))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "skin_image_analysis_b200")
    for dirpath, _d, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src, fn


def test_no_cuda_no_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from skin_image_analysis_b200 import tone_bias_test as tt
    from skin_image_analysis_b200._lib import SiaError
    from skin_image_analysis_b200.tone_bias_dataset import Rescale
    with pytest.raises(SiaError):
        tt.analyse_predictions(helpers.synthetic_instances(10, 1, False))
    with pytest.raises(SiaError):
        Rescale((8, 8))((np.zeros((16, 16, 3), np.float32), 0, 0))


@pytest.mark.parametrize("shape", [(450, 600, 224, 224), (450, 600, 512, 512), (96, 128, 40, 56), (100, 64, 50, 30)])
def test_tensor_core_tables_reproduce_the_oracle_operator(shape):
    """The numpy model of csrc/preprocess_tc.cu (fp16 vertical weights, 4-slot horizontal schedule) stays
    within 2.5e-4 of full scale of the oracle resize, and every output pixel is emitted exactly once."""
    import numpy as np
    from oracle import resize as R
    from skin_image_analysis_b200 import resize_weights as rw
    h, w, oh, ow = shape
    t = rw.build_tc_tables(h, w, oh, ow)
    assert t.n_tiles * t.tile_rows >= oh and t.tile_rows <= 128
    info = t.items.view(np.int32)
    emitted = info[info[:, 1] >= 0, 1]
    assert np.array_equal(emitted, np.arange(ow))                 # every output column exactly once, in order
    assert np.all(info[:, 0] >= 0) and np.all(info[:, 0] + rw.TC_ITEM_LOAD <= rw.TC_BLOCK_COLS)
    assert np.all(np.diff(info[:, 2]) >= 0) and info[-1, 2] == t.n_blocks - 1 and t.n_items % 8 == 0
    vec = info[0::4, 3]
    assert np.all((vec == -1) | ((vec >= 0) & (vec % 4 == 0) & (vec + 4 <= ow + 8)))
    assert t.a_packed.shape == (t.n_tiles, 128 * 256 * 2)
    rng = np.random.default_rng(h + ow)
    im = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    got = rw.tc_emulate(im, t, oh, ow)
    want = R.transform_u8(im, (oh, ow)).transpose(1, 2, 0)
    assert np.abs(got - want).max() <= 2.5e-4
    # A operand image: element (l, k) of tile 0 sits where the UMMA K-major core-matrix layout expects it
    a = t.a_packed[0].view(np.float16)
    for l, k in ((0, 0), (5, 3), (9, 17), (127, 255), (64, 100)):
        off = ((l // 8) * rw.TC_A_SBO + (k // 8) * rw.TC_A_LBO + (l % 8) * 16 + (k % 8) * 2) // 2
        assert float(a[off]) == t.a_dense[0, l, k]


def test_tensor_core_tables_reject_unsupported_geometry():
    from skin_image_analysis_b200 import resize_weights as rw
    for shape in ((97, 131, 64, 86), (1024, 1024, 224, 224), (64, 64, 128, 128), (450, 601, 224, 224)):
        with pytest.raises(ValueError):
            rw.build_tc_tables(*shape)


def test_tone_classifier_surface_on_cpu():
    """cnn_trial_dataset mirrors notebooks/ToneClassifier/CNNTrialDataset.py: the label converter (:11-25) and a
    transform that refuses CPU tensors instead of falling back."""
    import torch
    from skin_image_analysis_b200 import _lib
    from skin_image_analysis_b200.cnn_trial_dataset import TestTransforms, fitzpatrick_converter
    assert [fitzpatrick_converter(t) for t in ("I", "II", "III", "IV", "V", "VI")] == [0, 0, 1, 1, 1, 1]
    assert fitzpatrick_converter("VII") == "Error" and fitzpatrick_converter(float("nan")) == "Error"
    tf = TestTransforms()
    with pytest.raises(TypeError):
        tf(torch.zeros(3, 8, 8))
    if not torch.cuda.is_available():
        with pytest.raises(_lib.SiaError):
            tf(torch.zeros(3, 8, 8, dtype=torch.uint8))


@pytest.mark.parametrize("shape", [(450, 600, 224, 224), (480, 640, 224, 224)])
def test_two_product_tables_reproduce_the_oracle_operator(shape):
    """resize_weights.build_tc2_tables: the numpy model of the two-product kernel (fp16 V, fp16 Wx slices, register
    carry between column blocks) stays within 2.6e-4 of full scale of the oracle; every output column is completed
    exactly once, in column order, by a contiguous slot range; slots 17..20 of a block feed slots 0..3 of the next."""
    from skin_image_analysis_b200 import resize_weights as rw
    sh, sw, oh, ow = shape
    t = rw.build_tc2_tables(sh, sw, oh, ow)
    u8 = helpers.synthetic_u8_image(sh, sw, 77, "noise")
    em = rw.tc2_emulate(u8, t, oh, ow)
    want = R.transform_u8(u8, (oh, ow)).transpose(1, 2, 0)
    assert not np.isnan(em).any() and np.abs(em - want).max() <= 2.6e-4
    done = []
    for b in range(t.n_blocks):
        s_lo, s_hi, j_lo, _ = (int(v) for v in t.block_meta[b])
        assert 0 <= s_lo <= s_hi <= rw.TC2_SLOTS
        assert [int(c) for c in t.slot_col[b, s_lo:s_hi]] == list(range(j_lo, j_lo + s_hi - s_lo))
        assert (t.slot_col[b] >= 0).sum() == s_hi - s_lo
        done += list(range(j_lo, j_lo + s_hi - s_lo))
    assert done == list(range(ow))
    flat = rw.tc2_emulate(np.full((sh, sw, 3), 200, np.uint8), t, oh, ow)
    assert np.abs(flat - np.float32(200 / 255)).max() <= 2e-7


def test_two_product_tables_reject_unsupported_geometry():
    from skin_image_analysis_b200 import resize_weights as rw
    for shape in [(450, 600, 512, 512), (300, 400, 224, 224)]:
        with pytest.raises(ValueError):
            rw.build_tc2_tables(*shape)


@pytest.mark.parametrize("shape", [(450, 600, 224, 224), (450, 600, 512, 512), (300, 400, 224, 224), (96, 128, 64, 88),
                                   (480, 640, 224, 224), (200, 200, 224, 224)])
def test_warp_mma_tables_reproduce_the_oracle_operator(shape):
    """resize_weights.build_mma_tables: the fp16 weights are the exact operator rounded to nearest (row / column sums
    within 5e-4 of one), the fragment tables hold exactly the dense fp16 operators (through the row permutation and the
    K order of the second product), and the numpy model of the kernel's arithmetic is within 8e-4 of full scale of
    the oracle and within one bf16 ulp of it after rounding -- also on the sparse 0 / 255 pattern."""
    from skin_image_analysis_b200 import resize_weights as rw
    h, w, oh, ow = shape
    t = rw.build_mma_tables(h, w, oh, ow)
    assert t.kv in (2, 3) and t.n_tiles == (ow + 8) // 8 and t.n_groups == -(-w // 32)
    assert np.abs(t.wy16[:oh].sum(1) - 1.0).max() <= 5e-4 and np.abs(t.wx16[1:ow + 1].sum(1) - 1.0).max() <= 5e-4
    assert np.all(t.wx16[0] == 0) and np.all(t.wx16[ow + 1:] == 0)
    lane = np.arange(32)
    g, q = lane // 4, lane % 4
    qs, cs = rw.MMA_ROW_MAPS[t.row_map]
    rows_read = sorted({qs * qq + c for qq in range(4) for c in cs})
    assert rows_read == list(range(16))                    # the permutation covers the 16 rows of a chunk exactly once
    if w == 600:
        assert t.row_map == "spread"                       # 450-word rows: conflict-free only with the spread map
    # A fragments -> dense Wy: register (a0, a1, a2, a3) = rows (g, g+8, g, g+8), K slots (0/1, 0/1, 2/3, 2/3)
    for m in (0, t.n_msteps // 2, t.n_msteps - 1):
        dense = np.zeros((16, 16 * t.kv))
        for kc in range(t.kv):
            for reg, (rr, c_lo) in enumerate(((g, 0), (g + 8, 0), (g, 2), (g + 8, 2))):
                word = t.wy_frag[m, kc, :, reg]
                dense[rr, 16 * kc + qs * q + cs[c_lo]] = (word & 0xFFFF).astype(np.uint16).view(np.float16)
                dense[rr, 16 * kc + qs * q + cs[c_lo + 1]] = (word >> 16).astype(np.uint16).view(np.float16)
        rows = np.arange(16 * m, min(16 * m + 16, oh))
        cols = np.arange(t.r0[m], min(t.r0[m] + 16 * t.kv, h))
        assert np.array_equal(dense[:len(rows), :len(cols)] * 2.0 ** -15, t.wy16[rows][:, cols])
        assert t.r0[m] % 8 == 0 and not t.wy16[rows][:, :t.r0[m]].any() and not t.wy16[rows][:, t.r0[m] + 16 * t.kv:].any()
    # B fragments -> dense Wx: every non-zero weight appears exactly once, in the fragment of its source group
    seen = np.zeros_like(t.wx16)
    tile_last = np.searchsorted(t.tile_begin, np.arange(t.n_tiles), side="right") - 1
    assert t.tile_begin[0] == 0 and t.tile_begin[-1] == t.n_tiles and np.all(np.diff(t.tile_begin) >= 0)
    for tile in range(t.n_tiles):
        for rel in range(2):
            grp = int(tile_last[tile]) - 1 + rel
            for x in range(2):
                frag = t.wx_frag[tile, rel, x]
                assert bool(frag.any()) == bool((t.wx_mask[tile] >> (rel * 2 + x)) & 1)
                if grp < 0:
                    assert not frag.any()
                    continue
                for reg in range(2):
                    for half in range(2):
                        k = 2 * q + half + 8 * reg
                        px = 32 * grp + 4 * (k % 8) + 2 * x + k // 8
                        val = ((frag[:, reg] >> (16 * half)) & 0xFFFF).astype(np.uint16).view(np.float16).astype(np.float64)
                        ok = px < w
                        np.add.at(seen, ((8 * tile + g)[ok], px[ok]), val[ok])
                        assert not val[~ok].any()
    assert np.array_equal(seen, t.wx16)
    for kind in ("noise", "extremes", "smooth"):
        u8 = helpers.synthetic_u8_image(h, w, 77, kind)
        got = rw.mma_emulate(u8, t, oh, ow)
        want = R.transform_u8(u8, (oh, ow)).transpose(1, 2, 0)
        assert np.abs(got - want).max() <= 8e-4
        got_bf, want_bf = (torch.from_numpy(np.ascontiguousarray(a)).to(torch.bfloat16).float().numpy() for a in (got, want))
        assert np.all(np.abs(got_bf - want_bf) <= np.maximum(np.abs(want_bf), 2.0 ** -126) * 2.0 ** -7), kind
    flat = np.full((h, w, 3), 200, np.uint8)
    assert np.abs(rw.mma_emulate(flat, t, oh, ow) / (200.0 / 255.0) - 1.0).max() <= 1e-3


def test_warp_mma_tables_reject_unsupported_geometry():
    from skin_image_analysis_b200 import resize_weights as rw
    for shape in [(97, 131, 64, 64), (450, 600, 112, 112), (450, 600, 224, 220), (451, 600, 224, 224)]:
        with pytest.raises(ValueError):
            rw.build_mma_tables(*shape)

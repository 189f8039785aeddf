"""CPU: the C-ABI library loads, exports every symbol include/dppo.h declares, and its host-only
entry points (layout, record size, MT19937 permutation) are correct.  No device compute here."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "dppo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dppo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from diamond import _native as N
    lib = N.load_library()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"libdppo.so does not export {s}"
    assert sorted(N.EXPORTS) == syms
    assert lib.dppo_version() == 100


def test_every_option_of_dppo_set_option_is_documented_in_the_header():
    """dppo_set_option is the ABI's one stringly-typed entry point: every option name api.cu accepts must be described in the
    comment above its declaration in include/dppo.h (and nothing is documented that the library does not accept)."""
    api = open(os.path.join(ROOT, "diamond-ppo_b200", "csrc", "api.cu")).read()
    body = api[api.index('extern "C" int dppo_set_option'):]
    body = body[:body.index("\n}\n")]
    accepted = set(re.findall(r'strcmp\(name, "([a-z_0-9]+)"\)', body))
    assert {"tensor_cores", "row_sweep", "tc_prefetch", "gae_variant", "gae_inputs_settled"} <= accepted
    hdr = open(os.path.join(ROOT, "include", "dppo.h")).read()
    doc = hdr[:hdr.index("int dppo_set_option(")]
    doc = doc[doc.rindex("/*"):]
    documented = set(re.findall(r'^ \*   "([a-z_0-9]+)"', doc, flags=re.M))
    # "tc_debug" exists only in timing builds (-DDPPO_TIMING_SWITCHES) and is rejected by the release library
    assert documented == accepted - {"tc_debug"}, (sorted(documented), sorted(accepted))


def test_only_sm100a_code_is_embedded():
    import subprocess
    from diamond import _native as N
    out = subprocess.run(["cuobjdump", "--list-elf", N.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_layout_matches_parameter_counts():
    from diamond.flat import FlatMlp
    # SURVEY §8c KAT 4: parameter counts of the reference networks
    for (D, H, A, cont), count in {(4, 64, 2, False): 12995, (8, 64, 4, False): 13381, (3, 64, 1, True): 12867,
                                   (64, 256, 4, False): 215301}.items():
        fm = FlatMlp(D, H, A, cont)
        n = sum(int(np.prod(shape)) for _, shape in fm.slices.values())
        assert n == count
        offs = sorted((off, int(np.prod(shape))) for off, shape in fm.slices.values())
        for (o1, n1), (o2, _) in zip(offs, offs[1:]):
            assert o1 + n1 <= o2                      # no overlap
        assert fm.total >= offs[-1][0] + offs[-1][1]
        # critic_head.0 directly follows actor_head.0 (one [2H,H] product)
        head = "actor_mean_head" if cont else "actor_head"
        assert fm.slices["critic_head.0.weight"][0] == fm.slices[f"{head}.0.weight"][0] + H * H


def test_no_gpu_means_error_not_fallback():
    import torch
    from diamond import _native as N
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(N.NativeError):
        N.Context(0)


def test_host_permutation_bit_exact_vs_numpy_and_golden():
    from diamond import _native as N
    g = np.load(os.path.join(ROOT, "tests", "golden", "perm.npz"))
    for seed, n in ((42, 1024), (42, 524288), (123, 1024), (0, 1), (0, 2), (7, 1000), (5, 65537)):
        key, pos = N.mt19937_seed(seed)
        p, pos = N.permutation_mt19937(key, pos, n)
        assert p.dtype == np.int32
        np.testing.assert_array_equal(p[:32], g[f"s{seed}_n{n}.head"])
        np.testing.assert_array_equal(p[-32:], g[f"s{seed}_n{n}.tail"])
        assert int((p.astype(np.int64) * np.arange(1, n + 1)).sum()) == int(g[f"s{seed}_n{n}.checksum"])
    # continuing numpy's GLOBAL stream: same draws as np.random.permutation, and the stream ends up where numpy's would
    np.random.seed(123)
    a = N.numpy_global_permutations(4096, 2)
    np.testing.assert_array_equal(a[0], g["s123_n4096_x2.first"])
    np.testing.assert_array_equal(a[1], g["s123_n4096_x2.second"])
    after = np.random.randint(0, 2**31, size=4)
    np.random.seed(123)
    np.random.permutation(4096); np.random.permutation(4096)
    np.testing.assert_array_equal(after, np.random.randint(0, 2**31, size=4))
    # mid-stream start (pos not at a block boundary) and empty/ragged sizes
    np.random.seed(9); np.random.random(17)
    ref_state = np.random.get_state()
    ref = [np.random.permutation(n) for n in (0, 1, 3, 1000)]
    np.random.set_state(ref_state)
    got = [N.numpy_global_permutations(n, 1)[0] for n in (0, 1, 3, 1000)]
    for r, q in zip(ref, got):
        np.testing.assert_array_equal(r, q)


def test_step_record_bytes():
    from diamond import _native as N
    lib = N.load_library()
    assert lib.dppo_step_record_bytes(8, 4, 2, 0) == 2 * 8 * 4 * 4 + 8 * 8 + 8 * 8 + 2 * 8
    assert lib.dppo_step_record_bytes(3, 3, 2, 1) == 2 * 9 * 4 + 3 * 8 + 3 * 2 * 4 + 2 * 3


def test_rnn_layout_matches_reference_parameter_count():
    """SURVEY §8c KAT 4: the recurrent network of config R has 6 627 parameters; dppo_rnn_layout places the 14 tensors of
    recurrent_ppo.py:101-125 (nn.GRU names, gate order r,z,n) without overlap, every offset a multiple of 4 floats."""
    from diamond import _native as N
    from diamond.recurrent import rnn_param_slices
    desc = N.RnnDesc(4, 64, 16, 2)
    lay = N.rnn_layout(desc)
    sl = rnn_param_slices(desc, lay)
    assert len(sl) == 14
    assert sum(int(np.prod(shape)) for _, shape in sl.values()) == 6627
    offs = sorted((off, int(np.prod(shape))) for off, shape in sl.values())
    for (o1, n1), (o2, _) in zip(offs, offs[1:]):
        assert o1 + n1 <= o2 and o1 % 4 == 0
    assert lay.total >= offs[-1][0] + offs[-1][1]
    assert sl["critic_head.0.weight"][0] == sl["actor_head.0.weight"][0] + 64 * 16       # one [2H, Hg] product

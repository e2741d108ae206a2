"""The oracle's counter RNG: Philox4x32-10 against the Random123 known-answer vectors, and the integer mapping
against libstdc++'s own std::uniform_int_distribution (fixture produced by oracle/lemire_check.cpp)."""
import ctypes as C

import numpy as np

from oracle import pyoracle
from oracle.pyoracle import Oracle

# Random123 kat_vectors, philox4x32 10 rounds: (counter, key) -> output
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def test_philox_known_answers_python():
    for ctr, key, want in KAT:
        assert pyoracle.philox4x32_10(ctr, key) == want


def test_philox_known_answers_c():
    L = Oracle.lib()
    for ctr, key, want in KAT:
        c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
        L.orc_philox_export(c, k, o)
        assert tuple(o) == want


def test_integer_mapping_matches_libstdcxx(golden):
    """rows: stream step id a b value words_consumed, from the real std::uniform_int_distribution<int>."""
    L = Oracle.lib()
    L.orc_draw_export.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
    tab = golden["lemire"]
    seed = int(golden["seed"])
    plain = tab[tab[:, 0] != 99]
    assert len(plain) == 2400
    for stream, step, fid, a, b, val, _k in plain:
        assert L.orc_draw_export(seed, 0, int(step), int(fid), int(stream), 0, int(a), int(b)) == val
    for stream, step, fid, a, b, val, _k in plain[:300]:
        assert pyoracle.draw(seed, 0, int(step), int(fid), int(stream), int(a), int(b)) == val
    # crafted first words around the rejection threshold: libstdc++ consumed exactly one word iff we accept it
    crafted = tab[tab[:, 0] == 99]
    assert (crafted[:, 6] > 1).sum() > 10, "fixture must exercise the rejection path"
    L.orc_lemire_accept_export.argtypes = [C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_int)]
    for _s, word, _z, a, b, val, k in crafted:
        out = C.c_int(0)
        acc = L.orc_lemire_accept_export(int(word), int(a), int(b), C.byref(out))
        assert bool(acc) == (k == 1), (word, a, b, k)
        if acc:
            assert out.value == val


def test_draw_ranges_and_uniformity():
    L = Oracle.lib()
    L.orc_draw_export.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
    v = np.array([L.orc_draw_export(5, 0, 1, i, 1, 0, -10, 10) for i in range(21000)])
    assert v.min() == -10 and v.max() == 10
    counts = np.bincount(v + 10, minlength=21)
    assert np.abs(counts - 1000).max() < 150


def test_wide_k_is_folded_into_the_replicate_word():
    """k >= 2^24 (a neighbour id of a very large school in a tie draw) must address its own counter: the C oracle and the
    pure-Python restatement agree, and k and k + 2^24 differ"""
    import ctypes as C
    from oracle.pyoracle import Oracle, draw
    L = Oracle.lib()
    L.orc_draw_export.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_int]
    a = [L.orc_draw_export(99, 3, s, 5, 2, 7, -45, 45) for s in range(32)]
    b = [L.orc_draw_export(99, 3, s, 5, 2, 7 + (1 << 24), -45, 45) for s in range(32)]
    assert a != b
    assert b == [draw(99, 3, s, 5, 2, -45, 45, k0=7 + (1 << 24)) for s in range(32)]
    assert a == [draw(99, 3, s, 5, 2, -45, 45, k0=7) for s in range(32)]

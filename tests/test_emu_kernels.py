"""Kernel logic on the CPU: the product's .cu sources compiled against the SIMT emulator
(tests/emu, test infrastructure) and compared with the oracle.  Small inputs, reduced block size
(131072) so that several blocks, batches, lookback and the twice-inserted position are exercised;
the GPU suite repeats the comparison at the real block size on the real kernels."""
import pytest

from emu_lib import emu_compressor
from oracle_lib import oracle_compress
from smallz4_b200 import corpus

BS = 131072


@pytest.fixture(scope="module")
def emu():
    c = emu_compressor(block_size=BS, batch_blocks=2)
    yield c
    c.close()


def check(emu, data, level, legacy=False, dictionary=None):
    got = emu.compress(data, level=level, use_legacy_format=legacy, dictionary=dictionary)
    want, _ = oracle_compress(data, level, legacy, dictionary, block_size=BS)
    assert got == want


@pytest.mark.parametrize("n", [0, 1, 4, 5, 11, 12, 13, 19, 31, 32, 33, 64, 255, 1000])
def test_edge_sizes(emu, n):
    check(emu, corpus.make("text", n, 7).tobytes(), 9)
    check(emu, corpus.make("text", n, 7).tobytes(), 2)


@pytest.mark.parametrize("level", range(10))
def test_levels_text(emu, level):
    check(emu, corpus.make("text", 60_000, 11).tobytes(), level)


@pytest.mark.parametrize("level", [1, 4, 7, 9])
def test_levels_binary_and_runs(emu, level):
    check(emu, corpus.make("binary", 50_000, 11).tobytes(), level)
    check(emu, corpus.make("runs", 150_000, 11).tobytes(), level)


@pytest.mark.parametrize("level", [3, 6, 9])
def test_blocks_batches_and_tails(emu, level):
    # 3 blocks + a short tail, two batches; tails around the 12-byte lookback rule
    check(emu, corpus.make("mixed", 3 * BS + 777, 5).tobytes(), level)
    for extra in (5, 12):
        check(emu, corpus.make("text", BS + extra, 9).tobytes(), level)


@pytest.mark.parametrize("level", [2, 5, 9])
def test_long_runs_take_the_shortcut(emu, level):
    """smallz4.h:632: runs above 65 299 bytes; also the position right behind a skipped stretch."""
    check(emu, bytes(300_000), level)
    mid = corpus.make("text", 20_000, 3).tobytes() + b"\x07" * 90_000 + corpus.make("text", 9_000, 4).tobytes()
    check(emu, mid, level)


def test_legacy_frames(emu):
    check(emu, corpus.make("text", 200_000, 5).tobytes(), 9, legacy=True)
    check(emu, corpus.make("text", 200_000, 5).tobytes(), 0, legacy=True)     # reference writes empty blocks here
    check(emu, bytes(200_000), 5, legacy=True)


@pytest.mark.parametrize("dict_len", [1000, 65535, 65536, 70000])
def test_dictionary_ring_shift(emu, dict_len):
    """Q-dict: reads of the chain ring are shifted by one slot when a dictionary is loaded."""
    d = corpus.make("text", dict_len, 21, offset=1 << 40).tobytes()
    data = corpus.make("text", 40_000, 21).tobytes()
    for level in (1, 5, 9):
        check(emu, data, level, dictionary=d)


def test_dictionary_with_long_runs_is_fenced(emu):
    """-D plus a run long enough for the reference's long-run shortcut: refused by default (the device finds the run),
    replayed exactly by one device thread with allow_scalar_dict=1."""
    from smallz4_b200.api import Sz4Error
    with pytest.raises(Sz4Error, match="allow_scalar_dict"):
        emu.compress(bytes(150_000), level=9, dictionary=bytes(65536))
    emu.set_option("allow_scalar_dict", 1)
    try:
        check(emu, bytes(150_000), 9, dictionary=bytes(65536))
        check(emu, bytes(150_000), 3, dictionary=bytes(30000))
    finally:
        emu.set_option("allow_scalar_dict", 0)
    check(emu, corpus.make("text", 30_000, 5).tobytes(), 9, dictionary=bytes(1000))     # a short run of zeros is fine


def test_periodic_data_leaves_the_cost_ring(emu):
    """matches far longer than the DP's shared-memory ring (8192 positions) and many length classes"""
    check(emu, (b"abcdefg" * 20000)[:120_000], 9)


@pytest.mark.parametrize("age", [0, 2])
def test_long_walks_are_handed_to_k_long(age):
    """k_search hands walks that go on for many rounds to k_long (one warp per walk over the sorted arrays); with the
    threshold at 0 or 2 rounds nearly every walk takes that path.  Two chunks of the sort, a block border included."""
    c = emu_compressor(block_size=BS, batch_blocks=2, long_age=age, tail_lanes=8 * age)
    try:
        cases = [("mixed", BS + 20_000, 9), ("binary", 50_000, 4)] if age == 0 else [("binary", 70_000, 9), ("text", 40_000, 7)]
        for kind, n, level in cases:
            check(c, corpus.make(kind, n, 13).tobytes(), level)
        check(c, (b"abcdefg" * 20000)[:30_000], 9)
        check(c, (b"abcdefgh12345678" * 8000)[:40_000], 8)
    finally:
        c.close()


@pytest.mark.parametrize("n,level,legacy", [(0, 9, False), (5, 9, False), (2 * BS, 9, False), (2 * BS + 1, 3, False),
                                            (5 * BS + 777, 9, False), (3 * BS + 5, 5, True), (300_000, 0, False), (300_000, 0, True)])
def test_lz4_streams_batch_by_batch(n, level, legacy):
    """sz4_lz4 pulls and pushes batch by batch (two blocks per batch here) with ragged get_bytes returns; same frame."""
    c = emu_compressor(block_size=BS, batch_blocks=2, stream_blocks=2)
    try:
        data = corpus.make("mixed", n, 21).tobytes()
        pos, out = [0], []

        def get(k):
            chunk = data[pos[0]: pos[0] + min(k, 50_000)]
            pos[0] += len(chunk)
            return chunk

        c.lz4(get, out.append, max_chain_length=(65535 if level == 9 else level), use_legacy_format=legacy)
        want, _ = oracle_compress(data, level, legacy, block_size=BS)
        assert b"".join(out) == want
        assert len(out) >= 2 + (n > 2 * BS)            # header, records batch by batch, end mark: not one big push
    finally:
        c.close()


def test_input_arriving_in_pieces(emu):
    """From 16 sort chunks on (1 MiB in this build) the input of a batch is copied in four pieces and the histogram, the
    digit offsets and pass 1 of the sort run piece by piece behind them: same frame."""
    check(emu, corpus.make("mixed", 1_150_000, 19).tobytes(), 1)


def _class_reaching_in_front_of_the_batch(tail_bytes):
    """150 times "seven zeros, the input's first byte, nine random bytes": the 8-byte class of those positions also holds
    the anchor whose key is seven bytes of the zero padding in front of the batch plus the first byte -- a position at -7."""
    r = corpus.make("random", 2000, 5).tobytes()
    data = r[:20]
    for i in range(150):
        data += b"\0" * 7 + r[:1] + r[20 + 9 * i: 29 + 9 * i]
    return data + corpus.make("text", tail_bytes, 5).tobytes()


def test_k_long_stops_at_the_first_position_of_the_batch():
    """The sorted arrays of the LSD sort begin with a few anchors whose position lies in front of the batch (their keys
    contain the zero padding).  They sit at the far end of their class: the tables leave them out, and so must the walk
    over the sorted arrays -- it used to read in front of the batch (found by tools/emu_asan_check.py)."""
    c = emu_compressor(block_size=BS, batch_blocks=2, long_age=0)
    try:
        for level in (9, 2):
            check(c, _class_reaching_in_front_of_the_batch(3000), level)
    finally:
        c.close()

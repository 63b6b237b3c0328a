"""The oracle (oracle/smallz4_oracle.c) against digests of the unmodified reference's output.

tests/golden/golden.json was produced by tests/golden/make_golden.py from oracle/_ref
(the reference compiled from /root/reference).  This pins the oracle without needing the
reference at run time.
"""
import json
import os

import pytest

from golden_util import case_dict, case_id, case_input, digest
from oracle_lib import oracle_compress, oracle_decompress

with open(os.path.join(os.path.dirname(__file__), "golden", "golden.json")) as _f:
    CASES = json.load(_f)["cases"]

# level -9 of a few MiB costs the scalar oracle tens of seconds; keep the default CPU suite short
FAST = [c for c in CASES if c["size"] <= 400_000 or c["level"] <= 3]
SLOW = [c for c in CASES if c not in FAST]


def _check(c):
    data = case_input(c)
    d = case_dict(c)
    frame, stats = oracle_compress(data, c["level"], c["legacy"], d)
    assert len(frame) == c["frame_size"]
    assert digest(frame) == c["sha256"]
    if "frame_hex" in c:
        assert frame.hex() == c["frame_hex"]
    # the decoder restatement (smallz4cat) must give the input back.  Dictionary frames are
    # excluded: the reference's -D output does not decode with its own smallz4cat (DESIGN.md
    # "Q-dict", shown by test_reference_dictionary_frames_do_not_round_trip below)
    if not c["dict"] and not (c["legacy"] and c["level"] == 0):
        assert oracle_decompress(frame, len(data), d) == data
    assert stats["oob_first_reads"] == 0      # no case depends on the reference's out-of-bounds read


@pytest.mark.parametrize("c", FAST, ids=case_id)
def test_oracle_matches_reference_digest(c):
    _check(c)


@pytest.mark.slow
@pytest.mark.parametrize("c", SLOW, ids=case_id)
def test_oracle_matches_reference_digest_slow(c):
    _check(c)


def test_reference_dictionary_frames_do_not_round_trip():
    """Documents reference behaviour Q-dict: with -D the reference writes its chain ring with the
    block-relative index (smallz4.h:656) but reads it with the absolute one (smallz4.h:190), which a
    dictionary shifts by 65535.  The frames it produces are reproduced bit-exactly (digests above)
    and, exactly like the reference's own, they do not decode back to the input."""
    c = next(c for c in CASES if c["dict"] and c["dict"][2] == 65536 and c["level"] == 9 and c["kind"] == "text")
    data, d = case_input(c), case_dict(c)
    frame, _ = oracle_compress(data, c["level"], c["legacy"], d)
    assert digest(frame) == c["sha256"]
    try:
        back = oracle_decompress(frame, len(data) + 65536, d)
    except AssertionError:
        back = None
    assert back != data


# ---------------------------------------------------------------------------------------------
# The decoder restatement (oracle/lz4cat_oracle.c) against the reference's own smallz4cat
# (oracle/_ref/smallz4cat, compiled from /root/reference/smallz4cat.c by oracle/Makefile).
# ---------------------------------------------------------------------------------------------
def _cat_cases():
    return [c for c in CASES if not c["dict"] and not (c["legacy"] and c["level"] == 0) and c["size"] <= 400_000
            and c["level"] in (0, 1, 9) and (c["size"] <= 1000 or c["kind"] != "mixed")]


def test_decoder_restatement_is_pinned_to_the_reference_decoder():
    """Every golden frame decodes to the same bytes with the real smallz4cat and with the restatement, and
    both give the input back (smallz4cat.c:112-360)."""
    from oracle_lib import reference_cat, reference_decompress
    if reference_cat() is None:
        pytest.skip("oracle/_ref/smallz4cat not built (no /root/reference here)")
    n = 0
    for c in _cat_cases():
        data = case_input(c)
        frame, _ = oracle_compress(data, c["level"], c["legacy"])
        assert digest(frame) == c["sha256"]
        real = reference_decompress(frame)
        assert real == data, case_id(c)
        assert oracle_decompress(frame, len(data)) == real, case_id(c)
        n += 1
    assert n >= 30


def test_reference_decoder_rejects_what_the_restatement_rejects():
    """A truncated frame: neither decoder returns the input."""
    from oracle_lib import reference_cat
    import subprocess
    if reference_cat() is None:
        pytest.skip("oracle/_ref/smallz4cat not built")
    c = next(c for c in CASES if c["kind"] == "text" and c["size"] == 300_000 and c["level"] == 9)
    data = case_input(c)
    frame, _ = oracle_compress(data, 9)
    cut = frame[: len(frame) // 2]
    r = subprocess.run([reference_cat()], input=cut, capture_output=True)
    assert r.stdout != data
    try:
        back = oracle_decompress(cut, len(data))
    except AssertionError:
        back = None
    assert back != data
